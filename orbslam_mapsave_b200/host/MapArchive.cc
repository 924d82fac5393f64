// MapArchive.cc — see MapArchive.h.  Thin C++ view over the orbmap_* C ABI.
#include "MapArchive.h"

#include <cstring>

#include "../../include/orb_b200.h"

namespace ORB_SLAM2 {

MapArchiveB200::~MapArchiveB200() {
    if (mHandle) orbmap_destroy(mHandle);
}

bool MapArchiveB200::Load(const std::string& filename) {
    if (mHandle) {
        orbmap_destroy(mHandle);
        mHandle = nullptr;
    }
    if (orbmap_load(&mHandle, filename.c_str()) != ORB_OK) {
        mError = orb_last_error();
        mHandle = nullptr;
        return false;
    }
    return true;
}

bool MapArchiveB200::Save(const std::string& filename) const {
    if (!mHandle || orbmap_save(mHandle, filename.c_str()) != ORB_OK) {
        mError = mHandle ? orb_last_error() : "no map loaded";
        return false;
    }
    return true;
}

static orbmap_info info_of(const orbmap_archive* h) {
    orbmap_info i;
    std::memset(&i, 0, sizeof(i));
    if (h) orbmap_get_info(h, &i);
    return i;
}

long unsigned int MapArchiveB200::KeyFramesInMap() const { return (long unsigned int)info_of(mHandle).n_keyframes; }
long unsigned int MapArchiveB200::MapPointsInMap() const { return (long unsigned int)info_of(mHandle).n_mappoints; }
long unsigned int MapArchiveB200::GetMaxKFid() const { return (long unsigned int)info_of(mHandle).max_kf_id; }
bool MapArchiveB200::LoadValidated() const { return mHandle && info_of(mHandle).test_data == 0xdeadbeefu; }

static void to_cv(const std::vector<orbx_keypoint>& src, std::vector<cv::KeyPoint>& dst) {
    dst.resize(src.size());
    for (size_t i = 0; i < src.size(); i++) {
        dst[i].pt.x = src[i].x;
        dst[i].pt.y = src[i].y;
        dst[i].size = src[i].size;
        dst[i].angle = src[i].angle;
        dst[i].response = src[i].response;
        dst[i].octave = src[i].octave;
        dst[i].class_id = src[i].class_id;
    }
}

static void to_long(const std::vector<int64_t>& src, std::vector<long>& dst) { dst.assign(src.begin(), src.end()); }

bool MapArchiveB200::GetKeyFrame(size_t i, KeyFrameData& k) const {
    orbmap_keyframe_info inf;
    if (!mHandle || orbmap_keyframe_get_info(mHandle, 0, (int)i, &inf) != ORB_OK) {
        mError = mHandle ? orb_last_error() : "no map loaded";
        return false;
    }
    k = KeyFrameData();
    k.mnId = inf.id;
    k.mnFrameId = inf.frame_id;
    k.mTimeStamp = inf.timestamp;
    k.N = inf.n;
    k.mnScaleLevels = inf.n_levels;
    k.mfScaleFactor = inf.scale_factor;
    k.mfLogScaleFactor = inf.log_scale_factor;
    k.fx = inf.fx; k.fy = inf.fy; k.cx = inf.cx; k.cy = inf.cy; k.invfx = inf.invfx; k.invfy = inf.invfy;
    k.mbf = inf.bf; k.mb = inf.b; k.mThDepth = inf.th_depth;
    k.mnMinX = inf.min_x; k.mnMinY = inf.min_y; k.mnMaxX = inf.max_x; k.mnMaxY = inf.max_y;
    k.mnGridCols = inf.grid_cols; k.mnGridRows = inf.grid_rows;
    k.mfGridElementWidthInv = inf.grid_inv_w; k.mfGridElementHeightInv = inf.grid_inv_h;
    k.hasParent = inf.has_parent != 0;
    k.parentId = inf.parent_id;
    k.mbBad = inf.is_bad != 0;

    std::vector<orbx_keypoint> keys(inf.n_keys), keysUn(inf.n_keys_un);
    std::vector<int64_t> mp(inf.n_mappoint_slots);
    k.mvuRight.resize(inf.n_uright);
    k.mvDepth.resize(inf.n_depth);
    k.mvScaleFactors.resize(inf.n_scale_factors);
    k.mvLevelSigma2.resize(inf.n_scale_factors);
    k.mvInvLevelSigma2.resize(inf.n_scale_factors);
    if (inf.desc_rows > 0 && inf.desc_cols > 0) k.mDescriptors.create(inf.desc_rows, inf.desc_cols, CV_8U);
    k.Tcw.create(4, 4, CV_32F);
    k.mK.create(3, 3, CV_32F);
    if (orbmap_keyframe_arrays(mHandle, 0, (int)i, keys.data(), keysUn.data(), k.mvuRight.data(), k.mvDepth.data(),
                               k.mDescriptors.empty() ? nullptr : k.mDescriptors.ptr(), mp.data(), k.mvScaleFactors.data(),
                               k.mvLevelSigma2.data(), k.mvInvLevelSigma2.data(), k.Tcw.ptr<float>(), k.mK.ptr<float>()) != ORB_OK) {
        mError = orb_last_error();
        return false;
    }
    to_cv(keys, k.mvKeys);
    to_cv(keysUn, k.mvKeysUn);
    to_long(mp, k.mvpMapPointIds);

    std::vector<int64_t> conn(inf.n_connected), ord(inf.n_ordered), ch(inf.n_children), le(inf.n_loop_edges);
    std::vector<int32_t> connW(inf.n_connected);
    k.mvOrderedWeights.resize(inf.n_ordered);
    if (orbmap_keyframe_links(mHandle, 0, (int)i, conn.data(), connW.data(), ord.data(), k.mvOrderedWeights.data(), ch.data(),
                              le.data()) != ORB_OK) {
        mError = orb_last_error();
        return false;
    }
    k.mConnectedKeyFrameWeights.resize(conn.size());
    for (size_t j = 0; j < conn.size(); j++) k.mConnectedKeyFrameWeights[j] = std::make_pair((long)conn[j], (int)connW[j]);
    to_long(ord, k.mvpOrderedConnectedKeyFrames);
    to_long(ch, k.mspChildrens);
    to_long(le, k.mspLoopEdges);

    int32_t nCells = 0, nEntries = 0;
    if (orbmap_keyframe_grid(mHandle, 0, (int)i, nullptr, nullptr, 0, &nCells, &nEntries) != ORB_OK) {
        mError = orb_last_error();
        return false;
    }
    k.gridOffsets.assign(nCells + 1, 0);
    k.gridFeatures.assign(nEntries > 0 ? nEntries : 1, 0);
    if (orbmap_keyframe_grid(mHandle, 0, (int)i, k.gridOffsets.data(), k.gridFeatures.data(), nEntries, nullptr, nullptr) != ORB_OK) {
        mError = orb_last_error();
        return false;
    }
    k.gridFeatures.resize(nEntries);
    return true;
}

std::vector<MapArchiveB200::MapPointData> MapArchiveB200::GetAllMapPoints() const {
    std::vector<MapPointData> out;
    if (!mHandle) return out;
    const int n = info_of(mHandle).n_mappoints;
    std::vector<uint64_t> ids(n);
    std::vector<float> wp(3 * (size_t)n), nv(3 * (size_t)n), dmin(n), dmax(n);
    std::vector<uint8_t> desc(32 * (size_t)n), bad(n);
    std::vector<int64_t> ref(n);
    std::vector<int32_t> nobs(n), vis(n), found(n), off(n + 1);
    if (orbmap_mappoints(mHandle, ids.data(), wp.data(), nv.data(), desc.data(), ref.data(), bad.data(), nobs.data(), vis.data(),
                         found.data(), dmin.data(), dmax.data(), off.data()) != ORB_OK) {
        mError = orb_last_error();
        return out;
    }
    std::vector<int64_t> okf(off[n] > 0 ? off[n] : 1), oidx(off[n] > 0 ? off[n] : 1);
    if (orbmap_observations(mHandle, okf.data(), oidx.data()) != ORB_OK) {
        mError = orb_last_error();
        return out;
    }
    out.resize(n);
    for (int i = 0; i < n; i++) {
        MapPointData& p = out[i];
        p.mnId = ids[i];
        p.mWorldPos.create(3, 1, CV_32F);
        p.mNormalVector.create(3, 1, CV_32F);
        p.mDescriptor.create(1, 32, CV_8U);
        for (int c = 0; c < 3; c++) {
            p.mWorldPos.at<float>(c) = wp[3 * (size_t)i + c];
            p.mNormalVector.at<float>(c) = nv[3 * (size_t)i + c];
        }
        std::memcpy(p.mDescriptor.ptr(), desc.data() + 32 * (size_t)i, 32);
        p.refKFId = (long)ref[i];
        p.nObs = nobs[i];
        p.mnVisible = vis[i];
        p.mnFound = found[i];
        p.mbBad = bad[i] != 0;
        p.mfMinDistance = dmin[i];
        p.mfMaxDistance = dmax[i];
        for (int e = off[i]; e < off[i + 1]; e++) p.mObservations.push_back(std::make_pair((long)okf[e], (long)oidx[e]));
    }
    return out;
}

bool MapArchiveB200::ObservedDescriptors(cv::Mat& descriptors, std::vector<int>& offsets) const {
    if (!mHandle) {
        mError = "no map loaded";
        return false;
    }
    const int n = info_of(mHandle).n_mappoints;
    offsets.assign(n + 1, 0);
    int64_t total = 0;
    if (orbmap_observed_descriptors(mHandle, nullptr, offsets.data(), 0, &total) != ORB_OK) {
        mError = orb_last_error();
        return false;
    }
    descriptors.release();
    if (total == 0) return true;
    descriptors.create((int)total, 32, CV_8U);
    if (orbmap_observed_descriptors(mHandle, descriptors.ptr(), offsets.data(), total, nullptr) != ORB_OK) {
        mError = orb_last_error();
        return false;
    }
    return true;
}

}  // namespace ORB_SLAM2
