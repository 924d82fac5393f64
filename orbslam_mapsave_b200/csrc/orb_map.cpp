// orb_map.cpp — the fork's on-disk map (System::SaveMap / LoadMap, src/System.cc:552-574) as a descriptor source for
// the GPU matcher (second half).  Host-only code: no kernel lives here; the archive is parsed into flat
// tables that the orbm_* / orbv_* entry points consume.
//
// The file is a Boost.Serialization *binary_oarchive* opened with `no_header` holding `oa << mpMap` (a `Map*`).  Boost
// is not part of this image, so the byte layout is restated from Boost's published archive format (the fork's
// CMakeLists pins ROS kinetic = Ubuntu 16.04 = Boost 1.58, archive library version 12; the layout below is the one of
// every library version >= 7, i.e. Boost >= 1.44):
//   * primitives are raw little-endian bytes of their C++ size on x86-64 (int 4, long/size_t 8, float 4, double 8,
//     bool 1);
//   * an object saved through a POINTER (`oa << mpMap`) is preceded by class_id (int16; -1 = NULL pointer) and, the
//     first time its class appears, tracking_type (1 byte) + version_type (uint32); a tracked object then carries an
//     object_id (uint32);
//   * an object of class type saved BY VALUE (`ar & *pMapPoint`, every cv::Mat, cv::KeyPoint, std::vector<...>) is
//     preceded, only the first time its class appears in the archive, by tracking_type + version_type — the optional
//     class id is never written by binary archives.  Nothing but `Map` is ever saved through a pointer in the fork, so
//     under Boost's default `track_selectively` these classes are untracked (tracking byte 0, no object ids).  A set
//     tracking byte is still honoured (object_id per object, id == running count means "new object");
//   * std::vector<arithmetic> = collection_size_type (uint64) + the raw elements (array optimisation of binary
//     archives) and NO class preamble (BOOST_SERIALIZATION_COLLECTION_TRAITS makes them `object_serializable`); any other
//     std::vector = [first time: tracking + version] + collection_size_type + item_version_type (uint32) + the items.
// What follows that framing is the field order of Map::load (src/Map.cc:76-134), MapPoint::load (src/MapPoint.cc:142-213),
// KeyFrame::load (src/KeyFrame.cc:308-510), the free cv::Mat save/load and the cv::KeyPoint serialize of
// include/MapPoint.h:198-247 (which stores `response` twice and never `size`).
//
// PARITY: unpinned — the reference ships no .bin map and Boost cannot be run here.  tests/map_archive_writer.py is a second,
// independent (Python struct.pack) statement of the same layout; reader and writer below share ONE field list (the
// `io()` templates), so a round trip proves symmetry, and the Python writer proves the two statements agree byte for byte.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/orb_b200.h"

void orb_set_error(const char* fmt, ...);

namespace {

enum ClassId { C_MAP, C_MAPPOINT, C_KEYFRAME, C_MAT, C_KEYPOINT, C_VEC_KEYPOINT, C_VEC_VEC_SIZE, C_VEC_VEC_VEC_SIZE, C_COUNT };
const char* kClassName[C_COUNT] = {"Map", "MapPoint", "KeyFrame", "cv::Mat", "cv::KeyPoint", "vector<KeyPoint>", "vector<vector<size_t>>",
                                   "vector<vector<vector<size_t>>>"};

struct ClassState {
    bool seen = false, tracked = false;
    uint32_t version = 0;
};

struct Mat {                    // include/MapPoint.h:214-247
    int32_t cols = 0, rows = 0;
    uint64_t elem_size = 0, elem_type = 0;
    std::vector<uint8_t> data;
};

struct KeyPoint {               // include/MapPoint.h:198-209: angle, class_id, octave, response, response, pt.x, pt.y
    float angle = -1.f;
    int32_t class_id = -1, octave = 0;
    float response = 0.f, response2 = 0.f, x = 0.f, y = 0.f;
};

struct IdRef {
    bool valid = false;
    uint64_t id = 0;
};
struct Observation {
    bool valid = false;
    uint64_t kf = 0, idx = 0;
};
struct Connection {
    bool valid = false;
    uint64_t kf = 0;
    int32_t weight = 0;
};

struct MapPointRec {
    uint64_t id = 0, next_id = 0;
    int64_t first_kf = 0, first_frame = 0;
    int32_t n_obs = 0;
    float proj_x = 0, proj_y = 0, proj_xr = 0;
    bool track_in_view = false;
    int32_t track_scale_level = 0;
    float track_view_cos = 0;
    uint64_t track_ref_frame = 0, last_frame_seen = 0, ba_local_kf = 0, fuse_candidate_kf = 0, loop_point_kf = 0, corrected_by_kf = 0,
             corrected_ref = 0;
    Mat pos_gba;
    uint64_t ba_global_kf = 0;
    Mat world_pos;
    std::vector<Observation> obs;
    Mat normal, desc;
    IdRef ref_kf;
    int32_t visible = 1, found = 1;
    bool bad = false;
    float min_dist = 0, max_dist = 0;
};

struct KeyFrameRec {
    uint64_t next_id = 0, id = 0, frame_id = 0;
    double timestamp = 0;
    int32_t grid_cols = 0, grid_rows = 0;
    float grid_inv_w = 0, grid_inv_h = 0;
    uint64_t track_ref_frame = 0, fuse_target_kf = 0, ba_local_kf = 0, ba_fixed_kf = 0, loop_query = 0;
    int32_t loop_words = 0;
    float loop_score = 0;
    uint64_t reloc_query = 0;
    int32_t reloc_words = 0;
    float reloc_score = 0;
    Mat tcw_gba, tcw_bef_gba;
    uint64_t ba_global_kf = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, bf = 0, b = 0, th_depth = 0;
    int32_t n = 0;
    std::vector<KeyPoint> keys, keys_un;
    std::vector<float> uright, depth;
    Mat desc, tcp;
    int32_t n_levels = 0;
    float scale_factor = 0, log_scale_factor = 0;
    std::vector<float> scale_factors, level_sigma2, inv_level_sigma2;
    int32_t min_x = 0, min_y = 0, max_x = 0, max_y = 0;
    Mat K, Tcw, Twc, Ow, Cw;
    std::vector<IdRef> mappoints;
    std::vector<std::vector<std::vector<uint64_t>>> grid;
    std::vector<Connection> connected;
    std::vector<IdRef> ordered;
    std::vector<int32_t> ordered_weights;
    bool first_connection = true;
    IdRef parent;
    std::vector<IdRef> children, loop_edges;
    bool not_erase = false, to_be_erased = false, is_bad = false;
    float half_baseline = 0;
};

struct MapRec {
    std::vector<MapPointRec> mappoints;
    std::vector<KeyFrameRec> keyframes, origins;
    uint64_t max_kf_id = 0;
    uint32_t test_data = 0xdeadbeefu;    // src/Map.cc:22
    bool tracked[C_COUNT] = {};          // tracking flag per class as read from the file (kept so that a save reproduces it)
    int64_t trailing_bytes = 0;          // what Map::load leaves unread (Map::save appends the map points a second time, :68-73)
    std::map<uint64_t, int> kf_index;    // mnId -> index into keyframes
};

// ---- archive streams ------------------------------------------------------------------------------------------------------
struct Reader {
    static constexpr bool reading = true;
    const uint8_t *p, *end;
    ClassState cls[C_COUNT];
    uint32_t n_objects = 0;
    void raw(void* dst, size_t n) {
        if ((size_t)(end - p) < n) throw std::runtime_error("map archive truncated");
        if (n) memcpy(dst, p, n);
        p += n;
    }
    template <class T> void prim(T& v) { static_assert(std::is_arithmetic<T>::value, ""); raw(&v, sizeof(T)); }
    void boolean(bool& v) {
        uint8_t b;
        raw(&b, 1);
        if (b > 1) throw std::runtime_error("map archive: bool byte is neither 0 nor 1 (layout mismatch)");
        v = b != 0;
    }
    // returns false when the object is a back-reference to one already read (tracked classes only)
    bool begin_object(ClassId c) {
        ClassState& s = cls[c];
        if (!s.seen) {
            boolean(s.tracked);
            prim(s.version);
            s.seen = true;
            if (s.version != 0) throw std::runtime_error(std::string("map archive: class version != 0 for ") + kClassName[c]);
        }
        if (s.tracked) {
            uint32_t oid;
            prim(oid);
            if (oid != n_objects) return false;
            ++n_objects;
        }
        return true;
    }
    size_t remaining() const { return (size_t)(end - p); }
};

struct Writer {
    static constexpr bool reading = false;
    std::vector<uint8_t> out;
    ClassState cls[C_COUNT];
    bool tracked_in[C_COUNT] = {};   // per-class tracking flags of the file this map was loaded from (all false for a built map)
    bool framing = true;             // false: field payload only (orbmap_mappoint_record)
    uint32_t n_objects = 1;          // object ids handed out so far (the Map itself is object 0)
    void raw(const void* src, size_t n) {
        const uint8_t* s = (const uint8_t*)src;
        out.insert(out.end(), s, s + n);
    }
    template <class T> void prim(T& v) { static_assert(std::is_arithmetic<T>::value, ""); raw(&v, sizeof(T)); }
    void boolean(bool& v) {
        uint8_t b = v ? 1 : 0;
        raw(&b, 1);
    }
    bool begin_object(ClassId c) {
        if (!framing) return true;
        ClassState& s = cls[c];
        if (!s.seen) {
            bool t = tracked_in[c];  // by-value classes are untracked in the fork (no pointer serialisation of them anywhere); a loaded
            uint32_t ver = 0;        // file's own flags are written back as they were read
            boolean(t);
            prim(ver);
            s.seen = true;
            s.tracked = t;
        }
        if (s.tracked) {             // every by-value object of a tracked class is a new object
            uint32_t oid = n_objects++;
            prim(oid);
        }
        return true;
    }
    size_t remaining() const { return 0; }
};

constexpr uint64_t kMaxCount = 1ull << 31;

template <class Ar> void io(Ar& ar, Mat& m) {
    if (!ar.begin_object(C_MAT)) throw std::runtime_error("map archive: back-referenced cv::Mat");
    ar.prim(m.cols);
    ar.prim(m.rows);
    ar.prim(m.elem_size);
    ar.prim(m.elem_type);
    if (Ar::reading) {
        if (m.cols < 0 || m.rows < 0 || m.elem_size > 1024) throw std::runtime_error("map archive: implausible cv::Mat header");
        uint64_t bytes = (uint64_t)m.cols * (uint64_t)m.rows * m.elem_size;
        if (bytes > ar.remaining()) throw std::runtime_error("map archive truncated inside a cv::Mat");
        m.data.resize(bytes);
    }
    if (!m.data.empty()) ar.raw(m.data.data(), m.data.size());
}

template <class Ar> void io(Ar& ar, KeyPoint& k) {
    if (!ar.begin_object(C_KEYPOINT)) throw std::runtime_error("map archive: back-referenced cv::KeyPoint");
    ar.prim(k.angle);
    ar.prim(k.class_id);
    ar.prim(k.octave);
    ar.prim(k.response);
    ar.prim(k.response2);
    ar.prim(k.x);
    ar.prim(k.y);
}

template <class Ar> uint64_t io_count(Ar& ar, uint64_t n) {
    ar.prim(n);
    if (Ar::reading && (n > kMaxCount || n > ar.remaining())) throw std::runtime_error("map archive: implausible collection size");
    return n;
}

// std::vector of an arithmetic type: Boost's serialization/vector.hpp ends with BOOST_SERIALIZATION_COLLECTION_TRAITS(std::vector), which
// gives vector<bool ... double> the implementation level `object_serializable` — no tracking byte, no version is ever written for
// them — and binary archives store the elements as one raw array.
template <class Ar, class T> void io_pod_vector(Ar& ar, std::vector<T>& v) {
    uint64_t n = io_count(ar, v.size());
    if (Ar::reading) {
        if (n * sizeof(T) > ar.remaining()) throw std::runtime_error("map archive truncated inside a vector");
        v.resize(n);
    }
    if (n) ar.raw(v.data(), n * sizeof(T));
}

template <class Ar, class T, class F> void io_obj_vector(Ar& ar, ClassId c, std::vector<T>& v, F&& item) {
    if (!ar.begin_object(c)) throw std::runtime_error("map archive: back-referenced vector");
    uint64_t n = io_count(ar, v.size());
    uint32_t item_version = 0;
    ar.prim(item_version);
    if (Ar::reading) v.resize(n);
    for (uint64_t i = 0; i < n; ++i) item(v[i]);
}

template <class Ar> int32_t io_nitems(Ar& ar, size_t have) {
    int32_t n = (int32_t)have;
    ar.prim(n);
    if (Ar::reading && (n < 0 || (size_t)n > ar.remaining())) throw std::runtime_error("map archive: implausible nItems");
    return n;
}

template <class Ar> void io_idrefs(Ar& ar, std::vector<IdRef>& v) {
    int32_t n = io_nitems(ar, v.size());
    if (Ar::reading) v.resize(n);
    for (auto& r : v) {
        ar.boolean(r.valid);
        if (r.valid) ar.prim(r.id);
        else r.id = 0;
    }
}

template <class Ar> void io(Ar& ar, MapPointRec& m) {       // src/MapPoint.cc:142-213
    if (!ar.begin_object(C_MAPPOINT)) throw std::runtime_error("map archive: back-referenced MapPoint");
    ar.prim(m.id);
    ar.prim(m.next_id);
    ar.prim(m.first_kf);
    ar.prim(m.first_frame);
    ar.prim(m.n_obs);
    ar.prim(m.proj_x);
    ar.prim(m.proj_y);
    ar.prim(m.proj_xr);
    ar.boolean(m.track_in_view);
    ar.prim(m.track_scale_level);
    ar.prim(m.track_view_cos);
    ar.prim(m.track_ref_frame);
    ar.prim(m.last_frame_seen);
    ar.prim(m.ba_local_kf);
    ar.prim(m.fuse_candidate_kf);
    ar.prim(m.loop_point_kf);
    ar.prim(m.corrected_by_kf);
    ar.prim(m.corrected_ref);
    io(ar, m.pos_gba);
    ar.prim(m.ba_global_kf);
    io(ar, m.world_pos);
    uint32_t n = (uint32_t)m.obs.size();
    ar.prim(n);
    if (Ar::reading) {
        if (n > ar.remaining()) throw std::runtime_error("map archive: implausible observation count");
        m.obs.resize(n);
    }
    for (auto& o : m.obs) {
        ar.boolean(o.valid);
        if (o.valid) {
            ar.prim(o.kf);
            ar.prim(o.idx);
        }
    }
    io(ar, m.normal);
    io(ar, m.desc);
    ar.boolean(m.ref_kf.valid);
    if (m.ref_kf.valid) ar.prim(m.ref_kf.id);
    ar.prim(m.visible);
    ar.prim(m.found);
    ar.boolean(m.bad);
    ar.prim(m.min_dist);
    ar.prim(m.max_dist);
}

template <class Ar> void io(Ar& ar, KeyFrameRec& k) {       // src/KeyFrame.cc:308-510
    if (!ar.begin_object(C_KEYFRAME)) throw std::runtime_error("map archive: back-referenced KeyFrame");
    ar.prim(k.next_id);
    ar.prim(k.id);
    ar.prim(k.frame_id);
    ar.prim(k.timestamp);
    ar.prim(k.grid_cols);
    ar.prim(k.grid_rows);
    ar.prim(k.grid_inv_w);
    ar.prim(k.grid_inv_h);
    ar.prim(k.track_ref_frame);
    ar.prim(k.fuse_target_kf);
    ar.prim(k.ba_local_kf);
    ar.prim(k.ba_fixed_kf);
    ar.prim(k.loop_query);
    ar.prim(k.loop_words);
    ar.prim(k.loop_score);
    ar.prim(k.reloc_query);
    ar.prim(k.reloc_words);
    ar.prim(k.reloc_score);
    io(ar, k.tcw_gba);
    io(ar, k.tcw_bef_gba);
    ar.prim(k.ba_global_kf);
    ar.prim(k.fx);
    ar.prim(k.fy);
    ar.prim(k.cx);
    ar.prim(k.cy);
    ar.prim(k.invfx);
    ar.prim(k.invfy);
    ar.prim(k.bf);
    ar.prim(k.b);
    ar.prim(k.th_depth);
    ar.prim(k.n);
    io_obj_vector(ar, C_VEC_KEYPOINT, k.keys, [&](KeyPoint& p) { io(ar, p); });
    io_obj_vector(ar, C_VEC_KEYPOINT, k.keys_un, [&](KeyPoint& p) { io(ar, p); });
    io_pod_vector(ar, k.uright);
    io_pod_vector(ar, k.depth);
    io(ar, k.desc);
    io(ar, k.tcp);
    ar.prim(k.n_levels);
    ar.prim(k.scale_factor);
    ar.prim(k.log_scale_factor);
    io_pod_vector(ar, k.scale_factors);
    io_pod_vector(ar, k.level_sigma2);
    io_pod_vector(ar, k.inv_level_sigma2);
    ar.prim(k.min_x);
    ar.prim(k.min_y);
    ar.prim(k.max_x);
    ar.prim(k.max_y);
    io(ar, k.K);
    io(ar, k.Tcw);
    io(ar, k.Twc);
    io(ar, k.Ow);
    io(ar, k.Cw);
    io_idrefs(ar, k.mappoints);
    io_obj_vector(ar, C_VEC_VEC_VEC_SIZE, k.grid, [&](std::vector<std::vector<uint64_t>>& col) {
        io_obj_vector(ar, C_VEC_VEC_SIZE, col, [&](std::vector<uint64_t>& cell) { io_pod_vector(ar, cell); });
    });
    int32_t nc = io_nitems(ar, k.connected.size());
    if (Ar::reading) k.connected.resize(nc);
    for (auto& c : k.connected) {
        ar.boolean(c.valid);
        if (c.valid) {
            ar.prim(c.kf);
            ar.prim(c.weight);
        }
    }
    io_idrefs(ar, k.ordered);
    io_pod_vector(ar, k.ordered_weights);
    ar.boolean(k.first_connection);
    ar.boolean(k.parent.valid);
    if (k.parent.valid) ar.prim(k.parent.id);
    else k.parent.id = 0;
    io_idrefs(ar, k.children);
    io_idrefs(ar, k.loop_edges);
    ar.boolean(k.not_erase);
    ar.boolean(k.to_be_erased);
    ar.boolean(k.is_bad);
    ar.prim(k.half_baseline);
}

// Map::load, src/Map.cc:76-134.  The pointer preamble of `ia >> mpMap` comes first.
void read_map(Reader& ar, MapRec& m) {
    int16_t class_id;
    ar.prim(class_id);
    if (class_id == -1) return;                                  // NULL Map*
    if (class_id != 0) throw std::runtime_error("map archive: first class id is not 0 (not a `oa << mpMap` archive)");
    ClassState& s = ar.cls[C_MAP];
    ar.boolean(s.tracked);
    ar.prim(s.version);
    s.seen = true;
    if (s.tracked) {
        uint32_t oid;
        ar.prim(oid);
        if (oid != 0) throw std::runtime_error("map archive: Map object id is not 0");
        ar.n_objects = 1;
    }
    int32_t n = io_nitems(ar, 0);
    m.mappoints.resize(n);
    for (auto& p : m.mappoints) io(ar, p);
    n = io_nitems(ar, 0);
    m.keyframes.resize(n);
    for (auto& k : m.keyframes) io(ar, k);
    n = io_nitems(ar, 0);
    m.origins.resize(n);
    for (auto& k : m.origins) io(ar, k);
    ar.prim(m.max_kf_id);
    ar.prim(m.test_data);
    m.trailing_bytes = (int64_t)ar.remaining();
    for (int c = 0; c < C_COUNT; ++c) m.tracked[c] = ar.cls[c].tracked;
}

// Map::save, src/Map.cc:31-74 (incl. the second copy of the map points that load never reads).
void write_map(Writer& ar, MapRec& m) {
    int16_t class_id = 0;
    bool tracked = true;          // a class saved through a pointer is tracked
    uint32_t version = 0, oid = 0;
    for (int c = 0; c < C_COUNT; ++c) ar.tracked_in[c] = m.tracked[c];
    ar.prim(class_id);
    ar.boolean(tracked);
    ar.prim(version);
    ar.prim(oid);
    io_nitems(ar, m.mappoints.size());
    for (auto& p : m.mappoints) io(ar, p);
    io_nitems(ar, m.keyframes.size());
    for (auto& k : m.keyframes) io(ar, k);
    io_nitems(ar, m.origins.size());
    for (auto& k : m.origins) io(ar, k);
    ar.prim(m.max_kf_id);
    ar.prim(m.test_data);
    io_nitems(ar, m.mappoints.size());
    for (auto& p : m.mappoints) io(ar, p);
}

void index_keyframes(MapRec& m) {
    m.kf_index.clear();
    for (size_t i = 0; i < m.keyframes.size(); ++i) m.kf_index.emplace(m.keyframes[i].id, (int)i);   // first one wins, like std::map::emplace
}

Mat make_mat_f32(int rows, int cols, const float* v) {
    Mat m;
    m.rows = rows;
    m.cols = cols;
    m.elem_size = 4;
    m.elem_type = 5;        // CV_32FC1
    m.data.resize((size_t)rows * cols * 4);
    if (v) memcpy(m.data.data(), v, m.data.size());
    return m;
}

void copy_mat_f32(const Mat& m, float* dst, int want) {
    // absent / differently typed matrices read as zeros so that the caller's buffer is always defined
    for (int i = 0; i < want; ++i) dst[i] = 0.f;
    if (m.elem_size == 4 && (m.elem_type & 7) == 5 && (int64_t)m.rows * m.cols >= want) memcpy(dst, m.data.data(), (size_t)want * 4);
}

const KeyFrameRec* pick_kf(const MapRec* m, int group, int i) {
    const std::vector<KeyFrameRec>& v = group == 0 ? m->keyframes : m->origins;
    if ((group != 0 && group != 1) || i < 0 || (size_t)i >= v.size()) return nullptr;
    return &v[i];
}

}  // namespace

struct orbmap_archive {
    MapRec map;
};

#define ORBMAP_TRY(body)                              \
    try {                                             \
        body                                          \
    } catch (const std::exception& e) {               \
        orb_set_error("%s", e.what());                \
        return ORB_ERR_ARG;                           \
    }

extern "C" {

int orbmap_create(orbmap_archive** out) {
    if (!out) {
        orb_set_error("orbmap_create: null out");
        return ORB_ERR_ARG;
    }
    *out = new orbmap_archive();
    return ORB_OK;
}

int orbmap_load(orbmap_archive** out, const char* path) {
    if (!out || !path) {
        orb_set_error("orbmap_load: null argument");
        return ORB_ERR_ARG;
    }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) {
        orb_set_error("orbmap_load: cannot open %s", path);
        return ORB_ERR_ARG;
    }
    std::vector<uint8_t> buf;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    if (got != buf.size()) {
        orb_set_error("orbmap_load: short read on %s", path);
        return ORB_ERR_ARG;
    }
    orbmap_archive* a = new orbmap_archive();
    try {
        Reader r;
        r.p = buf.data();
        r.end = buf.data() + buf.size();
        read_map(r, a->map);
        index_keyframes(a->map);
    } catch (const std::exception& e) {
        delete a;
        orb_set_error("orbmap_load(%s): %s", path, e.what());
        return ORB_ERR_ARG;
    }
    *out = a;
    return ORB_OK;
}

int orbmap_save(const orbmap_archive* ar, const char* path) {
    if (!ar || !path) {
        orb_set_error("orbmap_save: null argument");
        return ORB_ERR_ARG;
    }
    Writer w;
    ORBMAP_TRY(write_map(w, const_cast<MapRec&>(ar->map));)
    FILE* f = fopen(path, "wb");
    if (!f) {
        orb_set_error("orbmap_save: cannot open %s", path);
        return ORB_ERR_ARG;
    }
    size_t put = w.out.empty() ? 0 : fwrite(w.out.data(), 1, w.out.size(), f);
    int rc = fclose(f);
    if (put != w.out.size() || rc != 0) {
        orb_set_error("orbmap_save: short write on %s", path);
        return ORB_ERR_ARG;
    }
    return ORB_OK;
}

void orbmap_destroy(orbmap_archive* ar) { delete ar; }

int orbmap_get_info(const orbmap_archive* ar, orbmap_info* info) {
    if (!ar || !info) {
        orb_set_error("orbmap_get_info: null argument");
        return ORB_ERR_ARG;
    }
    const MapRec& m = ar->map;
    memset(info, 0, sizeof(*info));
    info->n_mappoints = (int32_t)m.mappoints.size();
    info->n_keyframes = (int32_t)m.keyframes.size();
    info->n_origins = (int32_t)m.origins.size();
    info->test_data = m.test_data;
    info->max_kf_id = m.max_kf_id;
    info->trailing_bytes = m.trailing_bytes;
    for (const auto& k : m.keyframes) info->total_features += k.desc.rows;
    for (const auto& p : m.mappoints) info->total_observations += (int64_t)p.obs.size();
    return ORB_OK;
}

int orbmap_keyframe_get_info(const orbmap_archive* ar, int group, int i, orbmap_keyframe_info* o) {
    const KeyFrameRec* k = ar ? pick_kf(&ar->map, group, i) : nullptr;
    if (!k || !o) {
        orb_set_error("orbmap_keyframe_get_info: bad archive / group / index");
        return ORB_ERR_ARG;
    }
    memset(o, 0, sizeof(*o));
    o->id = k->id;
    o->frame_id = k->frame_id;
    o->next_id = k->next_id;
    o->parent_id = k->parent.id;
    o->timestamp = k->timestamp;
    o->n = k->n;
    o->n_keys = (int32_t)k->keys.size();
    o->n_keys_un = (int32_t)k->keys_un.size();
    o->n_uright = (int32_t)k->uright.size();
    o->n_depth = (int32_t)k->depth.size();
    o->desc_rows = k->desc.rows;
    o->desc_cols = (int32_t)(k->desc.cols * k->desc.elem_size);
    o->n_mappoint_slots = (int32_t)k->mappoints.size();
    o->n_levels = k->n_levels;
    o->n_scale_factors = (int32_t)k->scale_factors.size();
    o->grid_cols = k->grid_cols;
    o->grid_rows = k->grid_rows;
    o->min_x = k->min_x;
    o->min_y = k->min_y;
    o->max_x = k->max_x;
    o->max_y = k->max_y;
    o->n_connected = (int32_t)k->connected.size();
    o->n_ordered = (int32_t)k->ordered.size();
    o->n_children = (int32_t)k->children.size();
    o->n_loop_edges = (int32_t)k->loop_edges.size();
    o->has_parent = k->parent.valid;
    o->is_bad = k->is_bad;
    o->not_erase = k->not_erase;
    o->to_be_erased = k->to_be_erased;
    o->first_connection = k->first_connection;
    o->scale_factor = k->scale_factor;
    o->log_scale_factor = k->log_scale_factor;
    o->fx = k->fx;
    o->fy = k->fy;
    o->cx = k->cx;
    o->cy = k->cy;
    o->invfx = k->invfx;
    o->invfy = k->invfy;
    o->bf = k->bf;
    o->b = k->b;
    o->th_depth = k->th_depth;
    o->grid_inv_w = k->grid_inv_w;
    o->grid_inv_h = k->grid_inv_h;
    o->half_baseline = k->half_baseline;
    return ORB_OK;
}

static void export_keys(const std::vector<KeyPoint>& src, orbx_keypoint* dst) {
    for (size_t i = 0; i < src.size(); ++i) {
        dst[i].x = src[i].x;
        dst[i].y = src[i].y;
        dst[i].size = 0.f;              // never stored by the fork's cv::KeyPoint serialize (include/MapPoint.h:198-209)
        dst[i].angle = src[i].angle;
        dst[i].response = src[i].response2;     // the second read overwrites the first on load
        dst[i].octave = src[i].octave;
        dst[i].class_id = src[i].class_id;
    }
}

int orbmap_keyframe_arrays(const orbmap_archive* ar, int group, int i, orbx_keypoint* keys, orbx_keypoint* keys_un, float* uright,
                           float* depth, uint8_t* desc, int64_t* mappoint_ids, float* scale_factors, float* level_sigma2,
                           float* inv_level_sigma2, float* Tcw, float* K) {
    const KeyFrameRec* k = ar ? pick_kf(&ar->map, group, i) : nullptr;
    if (!k) {
        orb_set_error("orbmap_keyframe_arrays: bad archive / group / index");
        return ORB_ERR_ARG;
    }
    if (keys) export_keys(k->keys, keys);
    if (keys_un) export_keys(k->keys_un, keys_un);
    if (uright && !k->uright.empty()) memcpy(uright, k->uright.data(), k->uright.size() * 4);
    if (depth && !k->depth.empty()) memcpy(depth, k->depth.data(), k->depth.size() * 4);
    if (desc && !k->desc.data.empty()) memcpy(desc, k->desc.data.data(), k->desc.data.size());
    if (mappoint_ids)
        for (size_t j = 0; j < k->mappoints.size(); ++j) mappoint_ids[j] = k->mappoints[j].valid ? (int64_t)k->mappoints[j].id : -1;
    if (scale_factors && !k->scale_factors.empty()) memcpy(scale_factors, k->scale_factors.data(), k->scale_factors.size() * 4);
    // the three scale tables are independent vectors in the file; the caller sizes all of them from n_scale_factors
    const size_t nsf = k->scale_factors.size();
    if (level_sigma2)
        for (size_t j = 0; j < nsf; ++j) level_sigma2[j] = j < k->level_sigma2.size() ? k->level_sigma2[j] : 0.f;
    if (inv_level_sigma2)
        for (size_t j = 0; j < nsf; ++j) inv_level_sigma2[j] = j < k->inv_level_sigma2.size() ? k->inv_level_sigma2[j] : 0.f;
    if (Tcw) copy_mat_f32(k->Tcw, Tcw, 16);
    if (K) copy_mat_f32(k->K, K, 9);
    return ORB_OK;
}

int orbmap_keyframe_links(const orbmap_archive* ar, int group, int i, int64_t* connected_ids, int32_t* connected_weights,
                          int64_t* ordered_ids, int32_t* ordered_weights, int64_t* children_ids, int64_t* loop_edge_ids) {
    const KeyFrameRec* k = ar ? pick_kf(&ar->map, group, i) : nullptr;
    if (!k) {
        orb_set_error("orbmap_keyframe_links: bad archive / group / index");
        return ORB_ERR_ARG;
    }
    for (size_t j = 0; j < k->connected.size(); ++j) {
        if (connected_ids) connected_ids[j] = k->connected[j].valid ? (int64_t)k->connected[j].kf : -1;
        if (connected_weights) connected_weights[j] = k->connected[j].valid ? k->connected[j].weight : 0;
    }
    auto refs = [](const std::vector<IdRef>& v, int64_t* dst) {
        if (dst)
            for (size_t j = 0; j < v.size(); ++j) dst[j] = v[j].valid ? (int64_t)v[j].id : -1;
    };
    refs(k->ordered, ordered_ids);
    if (ordered_weights)         // mvOrderedWeights is a vector of its own in the file; the caller sizes it from n_ordered
        for (size_t j = 0; j < k->ordered.size(); ++j) ordered_weights[j] = j < k->ordered_weights.size() ? k->ordered_weights[j] : 0;
    refs(k->children, children_ids);
    refs(k->loop_edges, loop_edge_ids);
    return ORB_OK;
}

int orbmap_keyframe_grid(const orbmap_archive* ar, int group, int i, int32_t* cell_offsets, int32_t* cell_features, int capacity,
                         int32_t* n_cells, int32_t* n_entries) {
    const KeyFrameRec* k = ar ? pick_kf(&ar->map, group, i) : nullptr;
    if (!k) {
        orb_set_error("orbmap_keyframe_grid: bad archive / group / index");
        return ORB_ERR_ARG;
    }
    // cells in the reference's mGrid[col][row] order (column-major), the CSR layout orbm_grid_view takes
    int cells = 0;
    int64_t entries = 0;
    for (const auto& col : k->grid) {
        cells += (int)col.size();
        for (const auto& cell : col) entries += (int64_t)cell.size();
    }
    if (n_cells) *n_cells = cells;
    if (n_entries) *n_entries = (int32_t)entries;
    if (!cell_offsets && !cell_features) return ORB_OK;
    if (entries > capacity) {
        orb_set_error("orbmap_keyframe_grid: capacity %d < %lld grid entries", capacity, (long long)entries);
        return ORB_ERR_CAPACITY;
    }
    int c = 0, e = 0;
    for (const auto& col : k->grid)
        for (const auto& cell : col) {
            if (cell_offsets) cell_offsets[c] = e;
            for (uint64_t f : cell) {
                if (cell_features) cell_features[e] = (int32_t)f;
                ++e;
            }
            ++c;
        }
    if (cell_offsets) cell_offsets[c] = e;
    return ORB_OK;
}

int orbmap_mappoints(const orbmap_archive* ar, uint64_t* ids, float* world_pos, float* normal, uint8_t* desc, int64_t* ref_kf,
                     uint8_t* bad, int32_t* n_obs, int32_t* visible, int32_t* found, float* min_dist, float* max_dist,
                     int32_t* obs_offsets) {
    if (!ar) {
        orb_set_error("orbmap_mappoints: null archive");
        return ORB_ERR_ARG;
    }
    const MapRec& m = ar->map;
    int64_t off = 0;
    for (size_t i = 0; i < m.mappoints.size(); ++i) {
        const MapPointRec& p = m.mappoints[i];
        if (ids) ids[i] = p.id;
        if (world_pos) copy_mat_f32(p.world_pos, world_pos + 3 * i, 3);
        if (normal) copy_mat_f32(p.normal, normal + 3 * i, 3);
        if (desc) {
            memset(desc + 32 * i, 0, 32);
            if (p.desc.data.size() >= 32) memcpy(desc + 32 * i, p.desc.data.data(), 32);
        }
        if (ref_kf) ref_kf[i] = p.ref_kf.valid ? (int64_t)p.ref_kf.id : -1;
        if (bad) bad[i] = p.bad;
        if (n_obs) n_obs[i] = p.n_obs;
        if (visible) visible[i] = p.visible;
        if (found) found[i] = p.found;
        if (min_dist) min_dist[i] = p.min_dist;
        if (max_dist) max_dist[i] = p.max_dist;
        if (obs_offsets) obs_offsets[i] = (int32_t)off;
        off += (int64_t)p.obs.size();
    }
    if (obs_offsets) obs_offsets[m.mappoints.size()] = (int32_t)off;
    return ORB_OK;
}

int orbmap_observations(const orbmap_archive* ar, int64_t* kf_ids, int64_t* feature_idx) {
    if (!ar) {
        orb_set_error("orbmap_observations: null archive");
        return ORB_ERR_ARG;
    }
    size_t e = 0;
    for (const auto& p : ar->map.mappoints)
        for (const auto& o : p.obs) {
            if (kf_ids) kf_ids[e] = o.valid ? (int64_t)o.kf : -1;
            if (feature_idx) feature_idx[e] = o.valid ? (int64_t)o.idx : -1;
            ++e;
        }
    return ORB_OK;
}

/* The gather loop of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:495-510) for every map point of the archive:
 * observations in the stored order (= std::map<KeyFrame*, size_t> order at save time), skipping invalid entries, keyframes
 * that are not in the map, bad keyframes and out-of-range rows. */
int orbmap_mappoint_record(const orbmap_archive* ar, int i, uint8_t* out, int64_t capacity, int64_t* n_bytes) {
    if (!ar || i < 0 || (size_t)i >= ar->map.mappoints.size() || !n_bytes) {
        orb_set_error("orbmap_mappoint_record: bad archive / index");
        return ORB_ERR_ARG;
    }
    Writer w;
    w.framing = false;
    MapPointRec rec = ar->map.mappoints[i];
    io(w, rec);
    *n_bytes = (int64_t)w.out.size();
    if (out) {
        if ((int64_t)w.out.size() > capacity) {
            orb_set_error("orbmap_mappoint_record: capacity %lld too small", (long long)capacity);
            return ORB_ERR_CAPACITY;
        }
        memcpy(out, w.out.data(), w.out.size());
    }
    return ORB_OK;
}

int orbmap_observed_descriptors(const orbmap_archive* ar, uint8_t* desc, int32_t* offsets, int64_t capacity, int64_t* n_total) {
    if (!ar) {
        orb_set_error("orbmap_observed_descriptors: null archive");
        return ORB_ERR_ARG;
    }
    const MapRec& m = ar->map;
    int64_t n = 0;
    for (size_t i = 0; i < m.mappoints.size(); ++i) {
        if (offsets) offsets[i] = (int32_t)n;
        for (const auto& o : m.mappoints[i].obs) {
            if (!o.valid) continue;
            auto it = m.kf_index.find(o.kf);
            if (it == m.kf_index.end()) continue;
            const KeyFrameRec& k = m.keyframes[it->second];
            if (k.is_bad || k.desc.elem_size != 1 || k.desc.cols != 32 || o.idx >= (uint64_t)k.desc.rows) continue;
            if (desc) {
                if (n >= capacity) {
                    orb_set_error("orbmap_observed_descriptors: capacity %lld too small", (long long)capacity);
                    return ORB_ERR_CAPACITY;
                }
                memcpy(desc + 32 * n, k.desc.data.data() + 32 * o.idx, 32);
            }
            ++n;
        }
    }
    if (offsets) offsets[m.mappoints.size()] = (int32_t)n;
    if (n_total) *n_total = n;
    return ORB_OK;
}

/* ---- building an archive from flat arrays (the MapSave direction) --------------------------------------------------- */

int orbmap_add_mappoint(orbmap_archive* ar, uint64_t id, int64_t first_kf_id, const float* world_pos, const float* normal,
                        const uint8_t* desc, int64_t ref_kf_id, int n_obs, const int64_t* obs_kf_ids, const int64_t* obs_feature_idx,
                        int visible, int found, float min_dist, float max_dist) {
    if (!ar || !world_pos || !normal || !desc || n_obs < 0 || (n_obs > 0 && (!obs_kf_ids || !obs_feature_idx))) {
        orb_set_error("orbmap_add_mappoint: bad argument");
        return ORB_ERR_ARG;
    }
    MapPointRec p;
    p.id = id;
    p.first_kf = first_kf_id;
    p.first_frame = first_kf_id;
    p.n_obs = n_obs;
    p.world_pos = make_mat_f32(3, 1, world_pos);
    p.normal = make_mat_f32(3, 1, normal);
    p.desc.rows = 1;
    p.desc.cols = 32;
    p.desc.elem_size = 1;
    p.desc.elem_type = 0;       // CV_8UC1
    p.desc.data.assign(desc, desc + 32);
    p.ref_kf.valid = ref_kf_id >= 0;
    p.ref_kf.id = ref_kf_id >= 0 ? (uint64_t)ref_kf_id : 0;
    for (int i = 0; i < n_obs; ++i) {
        Observation o;
        o.valid = obs_kf_ids[i] >= 0;
        o.kf = o.valid ? (uint64_t)obs_kf_ids[i] : 0;
        o.idx = o.valid ? (uint64_t)obs_feature_idx[i] : 0;
        p.obs.push_back(o);
    }
    p.visible = visible;
    p.found = found;
    p.min_dist = min_dist;
    p.max_dist = max_dist;
    ar->map.mappoints.push_back(std::move(p));
    MapRec& m = ar->map;
    uint64_t next = 0;
    for (const auto& q : m.mappoints) next = q.id + 1 > next ? q.id + 1 : next;
    for (auto& q : m.mappoints) q.next_id = next;           // the static MapPoint::nNextId every record carries
    return ORB_OK;
}

int orbmap_add_keyframe(orbmap_archive* ar, const orbmap_keyframe_info* info, const orbx_keypoint* keys, const orbx_keypoint* keys_un,
                        const float* uright, const float* depth, const uint8_t* desc, const int64_t* mappoint_ids,
                        const float* scale_factors, const float* level_sigma2, const float* inv_level_sigma2, const float* Tcw,
                        const float* K) {
    if (!ar || !info || info->n < 0 || (info->n > 0 && (!keys || !keys_un || !desc)) || !Tcw || !K || info->n_levels < 0 ||
        (info->n_levels > 0 && (!scale_factors || !level_sigma2 || !inv_level_sigma2))) {
        orb_set_error("orbmap_add_keyframe: bad argument");
        return ORB_ERR_ARG;
    }
    KeyFrameRec k;
    const int n = info->n;
    k.id = info->id;
    k.frame_id = info->frame_id;
    k.timestamp = info->timestamp;
    k.grid_cols = info->grid_cols > 0 ? info->grid_cols : 64;         // FRAME_GRID_COLS / ROWS, include/Frame.h:37-38
    k.grid_rows = info->grid_rows > 0 ? info->grid_rows : 48;
    k.min_x = info->min_x;
    k.min_y = info->min_y;
    k.max_x = info->max_x;
    k.max_y = info->max_y;
    // Frame::Frame: mfGridElementWidthInv = FRAME_GRID_COLS / (mnMaxX - mnMinX) in float (src/Frame.cc:156-157)
    k.grid_inv_w = info->grid_inv_w != 0.f ? info->grid_inv_w : (float)k.grid_cols / (float)(k.max_x - k.min_x);
    k.grid_inv_h = info->grid_inv_h != 0.f ? info->grid_inv_h : (float)k.grid_rows / (float)(k.max_y - k.min_y);
    k.fx = info->fx;
    k.fy = info->fy;
    k.cx = info->cx;
    k.cy = info->cy;
    k.invfx = info->invfx != 0.f ? info->invfx : 1.0f / info->fx;
    k.invfy = info->invfy != 0.f ? info->invfy : 1.0f / info->fy;
    k.bf = info->bf;
    k.b = info->b;
    k.th_depth = info->th_depth;
    k.half_baseline = info->half_baseline;
    k.n = n;
    auto import_keys = [n](const orbx_keypoint* src, std::vector<KeyPoint>& dst) {
        dst.resize(n);
        for (int i = 0; i < n; ++i) {
            dst[i].angle = src[i].angle;
            dst[i].class_id = src[i].class_id;
            dst[i].octave = src[i].octave;
            dst[i].response = dst[i].response2 = src[i].response;
            dst[i].x = src[i].x;
            dst[i].y = src[i].y;
        }
    };
    import_keys(keys, k.keys);
    import_keys(keys_un, k.keys_un);
    k.uright.assign(n, -1.f);
    k.depth.assign(n, -1.f);
    if (uright) k.uright.assign(uright, uright + n);
    if (depth) k.depth.assign(depth, depth + n);
    k.desc.rows = n;
    k.desc.cols = 32;
    k.desc.elem_size = 1;
    k.desc.elem_type = 0;
    if (n) k.desc.data.assign(desc, desc + (size_t)n * 32);
    k.n_levels = info->n_levels;
    k.scale_factor = info->scale_factor;
    k.log_scale_factor = info->log_scale_factor;
    if (info->n_levels) {
        k.scale_factors.assign(scale_factors, scale_factors + info->n_levels);
        k.level_sigma2.assign(level_sigma2, level_sigma2 + info->n_levels);
        k.inv_level_sigma2.assign(inv_level_sigma2, inv_level_sigma2 + info->n_levels);
    }
    k.K = make_mat_f32(3, 3, K);
    // KeyFrame::SetPose (src/KeyFrame.cc:792-806): Twc = [Rwc | Ow], Ow = -Rwc * tcw, Cw = Twc * (half_baseline, 0, 0, 1);
    // float products accumulated left to right, like cv::gemm on 3x3 floats.
    k.Tcw = make_mat_f32(4, 4, Tcw);
    float Twc[16] = {0}, Ow[3], Cw[4];
    for (int r = 0; r < 3; ++r) {
        float acc = 0.f;
        for (int c = 0; c < 3; ++c) {
            Twc[r * 4 + c] = Tcw[c * 4 + r];
            acc += Tcw[c * 4 + r] * Tcw[c * 4 + 3];
        }
        Ow[r] = -acc;
        Twc[r * 4 + 3] = Ow[r];
    }
    Twc[15] = 1.f;
    const float center[4] = {k.half_baseline, 0.f, 0.f, 1.f};
    for (int r = 0; r < 4; ++r) {
        float acc = 0.f;
        for (int c = 0; c < 4; ++c) acc += Twc[r * 4 + c] * center[c];
        Cw[r] = acc;
    }
    k.Twc = make_mat_f32(4, 4, Twc);
    k.Ow = make_mat_f32(3, 1, Ow);
    k.Cw = make_mat_f32(4, 1, Cw);
    k.mappoints.resize(n);
    for (int i = 0; i < n; ++i) {
        k.mappoints[i].valid = mappoint_ids && mappoint_ids[i] >= 0;
        k.mappoints[i].id = k.mappoints[i].valid ? (uint64_t)mappoint_ids[i] : 0;
    }
    // Frame::AssignFeaturesToGrid / PosInGrid (src/Frame.cc:341-356, 500-510): cell = round((kpUn - min) * inv), features that
    // round outside the grid are dropped
    k.grid.assign(k.grid_cols, std::vector<std::vector<uint64_t>>(k.grid_rows));
    for (int i = 0; i < n; ++i) {
        int gx = (int)__builtin_roundf((k.keys_un[i].x - (float)k.min_x) * k.grid_inv_w);
        int gy = (int)__builtin_roundf((k.keys_un[i].y - (float)k.min_y) * k.grid_inv_h);
        if (gx < 0 || gx >= k.grid_cols || gy < 0 || gy >= k.grid_rows) continue;
        k.grid[gx][gy].push_back((uint64_t)i);
    }
    k.parent.valid = info->has_parent != 0;
    k.parent.id = info->has_parent ? info->parent_id : 0;
    k.first_connection = info->first_connection != 0;
    k.is_bad = info->is_bad != 0;
    k.not_erase = info->not_erase != 0;
    k.to_be_erased = info->to_be_erased != 0;
    MapRec& m = ar->map;
    m.keyframes.push_back(std::move(k));
    if (info->id > m.max_kf_id) m.max_kf_id = info->id;
    uint64_t next = 0;
    for (const auto& q : m.keyframes) next = q.id + 1 > next ? q.id + 1 : next;
    for (auto& q : m.keyframes) q.next_id = next;
    index_keyframes(m);
    return ORB_OK;
}

int orbmap_set_keyframe_links(orbmap_archive* ar, int i, int n_connected, const int64_t* connected_ids, const int32_t* connected_weights,
                              int n_ordered, const int64_t* ordered_ids, const int32_t* ordered_weights, int n_children,
                              const int64_t* children_ids, int n_loop_edges, const int64_t* loop_edge_ids) {
    if (!ar || i < 0 || (size_t)i >= ar->map.keyframes.size() || n_connected < 0 || n_ordered < 0 || n_children < 0 || n_loop_edges < 0) {
        orb_set_error("orbmap_set_keyframe_links: bad argument");
        return ORB_ERR_ARG;
    }
    KeyFrameRec& k = ar->map.keyframes[i];
    k.connected.resize(n_connected);
    for (int j = 0; j < n_connected; ++j) {
        k.connected[j].valid = connected_ids[j] >= 0;
        k.connected[j].kf = k.connected[j].valid ? (uint64_t)connected_ids[j] : 0;
        k.connected[j].weight = k.connected[j].valid ? connected_weights[j] : 0;
    }
    auto refs = [](int n, const int64_t* src, std::vector<IdRef>& dst) {
        dst.resize(n);
        for (int j = 0; j < n; ++j) {
            dst[j].valid = src[j] >= 0;
            dst[j].id = dst[j].valid ? (uint64_t)src[j] : 0;
        }
    };
    refs(n_ordered, ordered_ids, k.ordered);
    k.ordered_weights.assign(ordered_weights, ordered_weights + (ordered_weights ? n_ordered : 0));
    refs(n_children, children_ids, k.children);
    refs(n_loop_edges, loop_edge_ids, k.loop_edges);
    return ORB_OK;
}

int orbmap_add_origin(orbmap_archive* ar, int keyframe_index) {
    if (!ar || keyframe_index < 0 || (size_t)keyframe_index >= ar->map.keyframes.size()) {
        orb_set_error("orbmap_add_origin: bad argument");
        return ORB_ERR_ARG;
    }
    ar->map.origins.push_back(ar->map.keyframes[keyframe_index]);     // Map::save writes *pKeyFrameOrigin again (src/Map.cc:52-56)
    return ORB_OK;
}

}  // extern "C"
