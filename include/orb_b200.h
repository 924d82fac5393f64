/* orb_b200.h — C ABI of the B200-native ORB front-end (liborb_b200.so).
 *
 * The reference (skaegy/ORBSLAM_MapSave) is a C++ library with no FFI; the drop-in boundary for its feature hot path is
 * the pair of C++ classes ORB_SLAM2::ORBextractor (include/ORBextractor.h:50-116) and ORB_SLAM2::ORBmatcher
 * (include/ORBmatcher.h:37-101).  orbslam_mapsave_b200/host/ re-implements those classes with the same signatures on
 * top of THIS header; every entry point below names the reference interface it replaces.
 *
 * Conventions: plain pointers + sizes, no C++/torch types.  All functions return 0 (ORB_OK) or a negative orb_status.
 * "host" pointers are ordinary CPU memory (pinned memory makes the copies asynchronous and faster); "device"
 * pointers (suffix _device entry points) are CUDA device memory on the handle's device and the call is asynchronous on
 * the given cudaStream_t (passed as void*; NULL = the handle's own stream).  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with ORB_ERR_CUDA.
 */
#ifndef ORB_B200_H_
#define ORB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum orb_status {
    ORB_OK = 0,
    ORB_ERR_CUDA = -1,          /* a CUDA runtime call failed; orb_last_error() has the text */
    ORB_ERR_ARG = -2,           /* bad argument (null pointer, size mismatch, unsupported geometry) */
    ORB_ERR_CAPACITY = -3,      /* an output capacity given by the caller is too small */
    ORB_ERR_OVERFLOW = -4,      /* a candidate buffer capped by ORBX_CAND_PER_CELL overflowed (default sizing cannot) */
    ORB_ERR_GEOMETRY = -5       /* a pyramid level is < 62 px: the reference divides by zero there (ORBextractor.cc:783-786) */
} orb_status;

/* Same 28-byte layout as cv::KeyPoint {Point2f pt; float size, angle, response; int octave, class_id}. */
typedef struct orbx_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_keypoint;

/* FAST candidate as fed to the octree (debug / parity view of ORBextractor.cc:788-828 `vToDistributeKeys`). */
typedef struct orbx_candidate {
    int32_t x, y, response;     /* level pixel coordinates */
} orbx_candidate;

typedef struct orbx_extractor orbx_extractor;

const char* orb_last_error(void);           /* thread-local text of the last failure */
int orb_device_count(void);                 /* number of CUDA devices, 0 if none / no driver */

/* ------------------------------------------------------------------------------------------------------------------
 * ORBextractor
 * ---------------------------------------------------------------------------------------------------------------- */

/* Replaces ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (include/ORBextractor.h:56-57, src/ORBextractor.cc:409-469).  width/height fix the image geometry the handle is
 * planned for (the C++ class re-plans when operator() sees another size); max_batch = frames processed per device
 * pass (workspace is allocated for that many).  device = CUDA ordinal. */
int orbx_create(orbx_extractor** out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                int width, int height, int max_batch, int device);
void orbx_destroy(orbx_extractor* ex);

/* Replaces GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:74-88) and exposes mnFeaturesPerLevel (src/ORBextractor.cc:434-445).  Any pointer may be
 * NULL; arrays have nlevels entries. */
int orbx_tables(const orbx_extractor* ex, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota);
/* The same tables without a handle (no CUDA device is touched): what the ORBextractor constructor computes (src/ORBextractor.cc:414-445).
 * Lets the drop-in class answer GetScaleFactors() ... before the first image has fixed the geometry. */
int orbx_compute_tables(int nfeatures, float scale_factor, int nlevels, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                        int* quota);
/* Level sizes (src/ORBextractor.cc:1114-1115). */
int orbx_level_size(const orbx_extractor* ex, int level, int* w, int* h);
/* Upper bound on keypoints per frame: nfeatures + 3 per level (the octree may overshoot each quota). */
int orbx_max_keypoints(const orbx_extractor* ex);

/* Replaces ORBextractor::operator()(image, mask, keypoints, descriptors) (src/ORBextractor.cc:1042-1108) for ONE host
 * image.  image: 8-bit gray, `stride` bytes per row.  mask: NULL/empty = none, else 8-bit, same size: pixels where
 * mask==0 are zeroed before the pyramid (this fork's behaviour, :1048-1053).  kp_out[cap], desc_out[cap*32].
 * *n_out = number of keypoints (level-major order, as the reference).  Blocks until the result is in host memory. */
int orbx_extract(orbx_extractor* ex, const uint8_t* image, int width, int height, int stride,
                 const uint8_t* mask, int mask_stride,
                 orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out);

/* Batch form (not in the reference: N frames per call for throughput).  Frames are `frame_stride` bytes apart, rows
 * `stride` bytes; masks (or NULL) laid out likewise with mask_frame_stride/mask_stride.  kp_out[n_frames*cap],
 * desc_out[n_frames*cap*32], n_out[n_frames].  Host pointers (pinned memory makes the copies true DMA); blocks until done.
 * More than max_batch frames are pipelined in chunks of max_batch (first and last chunk 64 frames): copy-in, compute and copy-out
 * run on separate streams over six device staging slots, and a helper handle (a second workspace of the same size, created on
 * first use and owned by `ex`) computes every other chunk.  The "last pass" views (orbx_get_pyramid*, orbx_stereo_matches) refer to
 * the final chunk.  Ordinary (pageable) caller memory — what cv::Mat and std::vector hold — is not handed to the driver's
 * single-threaded staging: a pool of host threads owned by the handle copies each chunk into pinned staging with non-temporal
 * stores (packing strided rows on the way) and copies the valid keypoints / descriptors of every frame back out (measured at
 * 640x480 on a 16-thread host: 134 k frames/s against 170 k from page-locked buffers and 28 k through the driver's staging).
 * ORBX_HOST_THREADS sets the pool size (default: 3/4 of the hardware threads, divided by LOCAL_WORLD_SIZE, at most 16).
 * Environment switches for measurements: ORBX_E2E_STREAMS (1..4), ORBX_E2E_SLOTS, ORBX_E2E_RAMP=0, ORBX_HOST_NT=0 (plain memcpy
 * in the pool), ORBX_E2E_TRACE=1 (per-chunk device timeline on stderr). */
int orbx_extract_batch(orbx_extractor* ex, const uint8_t* images, int n_frames, int width, int height, int stride,
                       size_t frame_stride, const uint8_t* masks, int mask_stride, size_t mask_frame_stride,
                       orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out);

/* The same for frames held as SEPARATE allocations (std::vector<cv::Mat>): images[f] points at frame f (all width x height, rows `stride`
 * bytes apart); no masks.  Pinned frames are copied by DMA one 2-D copy each, pageable ones go through the host staging pool. */
int orbx_extract_batch_ptrs(orbx_extractor* ex, const uint8_t* const* images, int n_frames, int width, int height, int stride,
                            orbx_keypoint* kp_out, uint8_t* desc_out, int cap, int* n_out);

/* Page-lock (cudaHostRegister) / unlock a caller-owned host buffer that is handed to orbx_extract_batch repeatedly: frames and results
 * then move by DMA straight from / into the caller's memory (no host copy at all; pageable buffers cost the host-pool copy above).
 * A one-off cost (the pages are pinned one by one), so register long-lived buffers, not per call. */
int orbx_host_register(void* p, size_t bytes);
int orbx_host_unregister(void* p);

/* Same, with every buffer already in device memory (frames tightly packed: stride = width, frame_stride = w*h).
 * n_frames may exceed max_batch (processed in passes).  Asynchronous on `stream`. */
int orbx_extract_batch_device(orbx_extractor* ex, const uint8_t* d_images, int n_frames,
                              const uint8_t* d_masks, orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap,
                              int* d_n_out, void* stream);
/* Poll the device-side status word after a *_device call has completed (ORB_OK / ORB_ERR_OVERFLOW / ORB_ERR_CAPACITY). */
int orbx_check_status(orbx_extractor* ex);

/* Replaces the public member `std::vector<cv::Mat> mvImagePyramid` (include/ORBextractor.h:90; read by
 * Frame::ComputeStereoMatches, src/Frame.cc:589-696): copies level `level` of frame `frame` of the LAST pass to host.
 * bordered != 0 copies the (w+38)x(h+38) buffer with its REFLECT_101 border (src/ORBextractor.cc:1125-1132), else
 * the w x h image. */
int orbx_get_pyramid_level(orbx_extractor* ex, int frame, int level, int bordered, uint8_t* dst, int dst_stride);
/* All levels of frame `frame` in one call (one border launch, nlevels asynchronous copies, one synchronisation):
 * dst[l] / dst_stride[l] as for orbx_get_pyramid_level.  This is what the C++ class uses to fill mvImagePyramid. */
int orbx_get_pyramid(orbx_extractor* ex, int frame, int bordered, uint8_t* const* dst, const int* dst_stride);
/* Parity/debug views of the last pass: the blurred level (src/ORBextractor.cc:1088-1089) and the FAST candidates of a
 * level in the reference's octree input order (:788-828).  *n_out receives the count (may exceed cap). */
int orbx_get_blurred_level(orbx_extractor* ex, int frame, int level, uint8_t* dst, int dst_stride);
int orbx_get_candidates(orbx_extractor* ex, int frame, int level, orbx_candidate* out, int cap, int* n_out);
/* Number of kernels launched by this handle so far (bench bookkeeping). */
long long orbx_launch_count(const orbx_extractor* ex);

/* Stage timing hooks for bench.py: run only some stages of a device pass (bit mask) on data of the last full pass. */
enum { ORBX_STAGE_PYRAMID = 1, ORBX_STAGE_FAST = 2, ORBX_STAGE_OCTREE = 4, ORBX_STAGE_BLUR = 8, ORBX_STAGE_DESCRIBE = 16,
       ORBX_STAGE_ALL = 31 };
int orbx_run_stages_device(orbx_extractor* ex, const uint8_t* d_images, int n_frames, const uint8_t* d_masks,
                           orbx_keypoint* d_kp_out, uint8_t* d_desc_out, int cap, int* d_n_out, int stage_mask,
                           void* stream);

/* Replaces void Frame::ComputeStereoMatches() (src/Frame.cc:584-756), the only reader of mvImagePyramid: row-band
 * candidates among the right keypoints, descriptor distance < TH_HIGH, 11x11 SAD over shifts -5..+5 on the pyramid level of
 * the left keypoint, parabola fit, disparity / depth, and the final cut at 1.5*1.4*median SAD.  The pyramids are the
 * device-resident ones of the LAST pass of the two handles (frame_left / frame_right inside those passes), so a stereo
 * front-end never downloads them.  kp_left / kp_right: mvKeys / mvKeysRight; u_right / depth: mvuRight / mvDepth (n_left each,
 * -1 where no match).  Host pointers; blocks until done. */
int orbx_stereo_matches(orbx_extractor* left, orbx_extractor* right, int frame_left, int frame_right,
                        const orbx_keypoint* kp_left, const uint8_t* desc_left, int n_left,
                        const orbx_keypoint* kp_right, const uint8_t* desc_right, int n_right,
                        float mbf, float mb, float* u_right, float* depth);

/* ------------------------------------------------------------------------------------------------------------------
 * ORBmatcher
 * ---------------------------------------------------------------------------------------------------------------- */
#define ORBM_TH_HIGH 100        /* ORBmatcher::TH_HIGH  (src/ORBmatcher.cc:37) */
#define ORBM_TH_LOW 50          /* ORBmatcher::TH_LOW   (src/ORBmatcher.cc:38) */
#define ORBM_HISTO_LENGTH 30    /* ORBmatcher::HISTO_LENGTH (src/ORBmatcher.cc:39) */

/* Replaces static ORBmatcher::DescriptorDistance(a, b) (src/ORBmatcher.cc:1650-1666) for n descriptor pairs:
 * dist[i] = Hamming(a[i], b[i]) over 256 bits.  Host pointers. */
int orbm_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int* dist, int device);

/* The inner loop of SearchByBoW (src/ORBmatcher.cc:205-229) over whole sets: for each of nq query descriptors the
 * first database index attaining the minimum distance, that distance, and the second-smallest distance (256 if
 * ndb < 2; index -1 and 256 if ndb == 0).  Host pointers. */
int orbm_hamming_top2(const uint8_t* q, int nq, const uint8_t* db, int ndb,
                      int* best_idx, int* best_dist, int* second_dist, int device);
/* Batched device form: `npairs` independent (query set, db set) problems.  Query set p = d_q + q_off[p]*32 with
 * q_cnt[p] descriptors, likewise db; outputs at d_out_* + q_off[p].  q_off/q_cnt/db_off/db_cnt are DEVICE int arrays.
 * max_q = max over p of q_cnt[p] (sizes the grid).  Asynchronous on `stream`. */
int orbm_hamming_top2_batch_device(const uint8_t* d_q, const int* d_q_off, const int* d_q_cnt,
                                   const uint8_t* d_db, const int* d_db_off, const int* d_db_cnt,
                                   int npairs, int max_q, int* d_best_idx, int* d_best_dist, int* d_second_dist,
                                   void* stream);
/* All-pairs keyframe matching (BASELINE.json config 4): every keyframe holds `per_kf` descriptors; query keyframes
 * [q_begin, q_end) of d_desc are matched against all n_kf keyframes.  d_count[(q-q_begin)*n_kf + k] = number of query
 * descriptors of q whose top-2 against keyframe k passes best <= th_low && best < ratio*second (uint16; per_kf <= 65535).  d_count must be
 * 4-byte aligned and hold an EVEN number of entries (round (q_end - q_begin) * n_kf up): the kernel updates it as 32-bit words.
 * If d_best_kf / d_best_dist are non-NULL they receive, per query descriptor, the keyframe (!= q) holding its
 * overall nearest descriptor and that distance.  Asynchronous on `stream`. */
int orbm_allpairs_device(const uint8_t* d_desc, int n_kf, int per_kf, int q_begin, int q_end, int th_low, float ratio,
                         uint16_t* d_count, int* d_best_kf, int* d_best_dist, void* stream);

/* The same on several GPUs of one node from ONE process (the multi-GPU form of SURVEY §8e for a C / C++ host; under torchrun the
 * per-rank call above plus one all-gather does the same).  desc: HOST pointer, n_kf x per_kf x 32 bytes; devices[n_devices]: CUDA ordinals
 * (a device may be listed more than once: its shards run on separate streams).  The database is uploaded once and replicated device
 * to device (NVLink where peer access exists); query keyframes are sharded contiguously and evenly over the listed devices, one host thread
 * each; the gather of the shard tables into count_out — HOST, n_kf x n_kf uint16, row = query keyframe — is the only exchange.
 * best_kf_out / best_dist_out (HOST, n_kf x per_kf ints, or both NULL) as in orbm_allpairs_device.  Blocks until done. */
int orbm_allpairs_multi(const uint8_t* desc, int n_kf, int per_kf, int th_low, float ratio, const int* devices, int n_devices,
                        uint16_t* count_out, int* best_kf_out, int* best_dist_out);

/* A DBoW2::FeatureVector (Thirdparty/DBoW2/DBoW2/FeatureVector.h:21-22, std::map<NodeId, vector<unsigned>>) flattened:
 * node ids ascending, CSR offsets (n_nodes+1), feature indices (ascending inside a node, as transform() appends). */
typedef struct orbm_featvec {
    int n_nodes;
    const int* node_ids;
    const int* offsets;
    const int* features;
} orbm_featvec;

/* One side of a search: the fields of KeyFrame / Frame the reference reads (include/KeyFrame.h:179-204,
 * include/Frame.h:141-189), snapshotted by the caller under the reference's own locks. */
typedef struct orbm_view {
    int n;                      /* N */
    const uint8_t* desc;        /* mDescriptors, n x 32 */
    const uint8_t* flag;        /* per feature: SearchByBoW: 1 = has a MapPoint that is not bad;
                                   SearchForTriangulation: 1 = GetMapPoint(i) != NULL.  May be NULL where unused. */
    const float* angle;         /* mvKeysUn[i].angle (KeyFrame) / mvKeys[i].angle (Frame) */
    const float* x;             /* mvKeysUn[i].pt.x   (triangulation only, else NULL) */
    const float* y;             /* mvKeysUn[i].pt.y   (triangulation only) */
    const int* octave;          /* mvKeysUn[i].octave (triangulation only) */
    const float* uright;        /* mvuRight[i]        (triangulation only) */
    orbm_featvec fv;            /* mFeatVec */
} orbm_view;

/* Replaces int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
 * (src/ORBmatcher.cc:159-291).  match21[F.n]: index of the KF feature whose MapPoint the reference stores in
 * vpMapPointMatches[j], or -1.  *n_matches = return value of the reference. */
int orbm_search_by_bow_kf_frame(const orbm_view* kf, const orbm_view* frame, float nnratio, int check_orientation,
                                int* match21, int* n_matches, int device);
/* Replaces int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
 * (src/ORBmatcher.cc:525-658).  match12[kf1.n]: index into KF2 (whose MapPoint the reference stores) or -1. */
int orbm_search_by_bow_kf_kf(const orbm_view* kf1, const orbm_view* kf2, float nnratio, int check_orientation,
                             int* match12, int* n_matches, int device);
/* Replaces int ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo)
 * (src/ORBmatcher.cc:660-826, CheckDistEpipolarLine :140-157).  F12: row-major 3x3 float.  (ex, ey): epipole of KF1's
 * centre in KF2 (:667-673).  scale_factors2 / level_sigma2_2: pKF2->mvScaleFactors / mvLevelSigma2 (n_levels2 entries).
 * pairs_out[2*kf1.n] receives (idx1, idx2) ascending idx1; *n_pairs their number; *n_matches the reference's return. */
int orbm_search_for_triangulation(const orbm_view* kf1, const orbm_view* kf2, const float* F12, float ex, float ey,
                                  const float* scale_factors2, const float* level_sigma2_2, int n_levels2,
                                  int only_stereo, int check_orientation,
                                  int* pairs_out, int* n_pairs, int* n_matches, int device);
/* Batched forms of the node-constrained searches (not in the reference, which calls them in loops): one view against n_others views
 * in ONE call — two kernel launches for the whole batch (grid = nodes x candidates), one upload, one download.  Results are identical to
 * calling the single-pair entry points in a loop.
 * orbm_search_by_bow_batch:
 *   mode 0: SearchByBoW(KeyFrame* others[i], Frame& anchor) for every i — the candidate loop of Tracking::Relocalization
 *           (src/Tracking.cc:1621-1643); match_out + i * anchor->n = match21 of candidate i;
 *   mode 1: SearchByBoW(KeyFrame* anchor, KeyFrame* others[i]) — LoopClosing::ComputeSim3 (src/LoopClosing.cc:240-266);
 *           match_out + i * anchor->n = match12 of candidate i.
 *   n_matches[n_others] = the reference's return values.
 * orbm_search_for_triangulation_batch: SearchForTriangulation(anchor, others[i], F12[i], ...) for every i — the loop over up to 20
 *   neighbours of LocalMapping::CreateNewMapPoints (src/LocalMapping.cc:215-268).  F12: n_others x 9; epipoles: n_others x 2 (ex, ey);
 *   scale_factors2 / level_sigma2_2: n_others x n_levels2; pairs_out + 2 * i * anchor->n = the (idx1, idx2) pairs of neighbour i,
 *   n_pairs[i] of them; n_matches[i] = the reference's return value. */
int orbm_search_by_bow_batch(const orbm_view* anchor, const orbm_view* others, int n_others, int mode, float nnratio,
                             int check_orientation, int* match_out, int* n_matches, int device);
int orbm_search_for_triangulation_batch(const orbm_view* anchor, const orbm_view* others, int n_others, const float* F12,
                                        const float* epipoles, const float* scale_factors2, const float* level_sigma2_2, int n_levels2,
                                        int only_stereo, int check_orientation, int* pairs_out, int* n_pairs, int* n_matches, int device);

/* Replaces ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1604-1645) on bin sizes; runs on the device as part of the
 * searches above; this entry exposes it for tests.  ind[3]. */
int orbm_three_maxima(const int* histo, int n_bins, int* ind, int device);

/* ------------------------------------------------------------------------------------------------------------------
 * ORB vocabulary: BoW assignment feeding the matcher (the first "next" row after the hot path)
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct orbv_vocabulary orbv_vocabulary;

/* Build a vocabulary tree from flat node arrays, the state TemplatedVocabulary::loadFromTextFile / loadFromBinaryFile
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1351-1440, 1467-1512) leave in m_nodes: node 0 is the root; parent[i] < i
 * for i >= 1 (children keep file order); descriptors n_nodes x 32 bytes (row 0 unused); weights[i] = Node::weight;
 * is_leaf[i] != 0 marks words, numbered in node order like m_words.  scoring / weighting: DBoW2 ScoringType / WeightingType. */
int orbv_create(orbv_vocabulary** out, int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                const uint8_t* descriptors, const double* weights, const uint8_t* is_leaf, int device);
/* Same from the two on-disk formats of the reference (text: ORBvoc.txt; binary: the fork's ORBvoc.bin, tools/bin_vocabulary.cc). */
int orbv_load_text(orbv_vocabulary** out, const char* path, int device);
int orbv_load_binary(orbv_vocabulary** out, const char* path, int device);
/* Writes the fork's binary format (TemplatedVocabulary::saveToBinaryFile, :1515-1536). */
int orbv_save_binary(const orbv_vocabulary* voc, const char* path);
void orbv_destroy(orbv_vocabulary* voc);
int orbv_info(const orbv_vocabulary* voc, int* k, int* L, int* n_nodes, int* n_words, int* scoring, int* weighting);

/* Replaces TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup) (:1231-1272) for n descriptors:
 * tree descent with FORB::distance (FORB.cpp:81-101), first child attaining the minimum wins; node_id = the node on the path
 * at level L - levelsup (0 = root when that is <= 0).  Host pointers; any output may be NULL. */
int orbv_transform(const orbv_vocabulary* voc, const uint8_t* desc, int n, int levelsup, int* word_id, double* weight, int* node_id);
/* Device-resident form, asynchronous on `stream`. */
int orbv_transform_device(const orbv_vocabulary* voc, const uint8_t* d_desc, int n, int levelsup, int* d_word_id, double* d_weight,
                          int* d_node_id, void* stream);

/* Replaces the selection loop of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:483-548) for a batch
 * of map points: set s holds descriptors [offsets[s], offsets[s+1]) of `desc` (the observations' descriptors in map order);
 * best_idx[s] = index inside the set of the descriptor with the least median Hamming distance to the set
 * (median = sorted row[int(0.5*(N-1))] incl. the zero self-distance, first index wins ties), -1 for an empty set.
 * Sets of more than 1024 descriptors are rejected (ORB_ERR_ARG).  Host pointers. */
int orbm_distinctive_descriptors(const uint8_t* desc, const int* offsets, int n_sets, int* best_idx, int device);

/* ------------------------------------------------------------------------------------------------------------------
 * Projection / window searches (the first "next" row): the ORBmatcher searches that look for a descriptor
 * inside a window of a Frame's / KeyFrame's feature grid (GetFeaturesInArea) instead of inside a vocabulary node.
 * ---------------------------------------------------------------------------------------------------------------- */

/* The side that is searched: the fields of Frame (include/Frame.h:98,141-197) / KeyFrame (include/KeyFrame.h:146-204) that
 * GetFeaturesInArea (src/Frame.cc:445-498, src/KeyFrame.cc:1311-1350) and the candidate loops read.  mGrid is flattened to
 * CSR: cell (ix, iy) = mGrid[ix][iy] has index ix*grid_rows + iy and holds cell_features[cell_offsets[c] .. cell_offsets[c+1])
 * in push order (AssignFeaturesToGrid, src/Frame.cc:341-356, note PosInGrid ROUNDS, :500-510). */
typedef struct orbm_grid_view {
    int n;                          /* N */
    const uint8_t* desc;            /* mDescriptors, n x 32 */
    const float* x;                 /* mvKeysUn[i].pt.x */
    const float* y;                 /* mvKeysUn[i].pt.y */
    const int* octave;              /* mvKeysUn[i].octave */
    const float* angle;             /* mvKeysUn[i].angle (NULL when the orientation check is off) */
    const float* uright;            /* mvuRight[i] (NULL = monocular, no stereo gate) */
    const uint8_t* blocked;         /* per feature, 1 = the candidate loop `continue`s on it from the start
                                       (e.g. mvpMapPoints[i] && mvpMapPoints[i]->Observations() > 0); NULL = none */
    int grid_cols, grid_rows;       /* FRAME_GRID_COLS x FRAME_GRID_ROWS (64 x 48) / mnGridCols x mnGridRows */
    float min_x, min_y, max_x, max_y;   /* mnMinX, mnMinY, mnMaxX, mnMaxY */
    float inv_w, inv_h;             /* mfGridElementWidthInv, mfGridElementHeightInv */
    const int* cell_offsets;        /* grid_cols*grid_rows + 1 */
    const int* cell_features;
    const float* scale_factors;     /* mvScaleFactors */
    int n_levels;
} orbm_grid_view;

/* Replaces int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)
 * (src/ORBmatcher.cc:45-129; caller Tracking::SearchLocalPoints).  Per map point, as left by Frame::isInFrustum
 * (src/Frame.cc:373-443): in_view = mbTrackInView && !isBad(); proj_x/proj_y/proj_xr = mTrackProjX/Y/XR;
 * level = mnTrackScaleLevel; view_cos = mTrackViewCos; desc = GetDescriptor(); claims = Observations() > 0 (a feature
 * given to such a point is skipped by later points, :87-89).
 * owner[F.n] receives, per feature, the index of the LAST map point the reference stores in F.mvpMapPoints[idx], or -1 if
 * the reference leaves that entry alone.  *n_matches = the reference's return value. */
int orbm_search_by_projection_map(const orbm_grid_view* frame, int n_points, const uint8_t* in_view, const float* proj_x,
                                  const float* proj_y, const float* proj_xr, const int* level, const float* view_cos,
                                  const uint8_t* desc, const uint8_t* claims, float th, float nnratio,
                                  int* owner, int* n_matches, int device);

/* Replaces int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
 * (src/ORBmatcher.cc:1331-1463; caller Tracking::TrackWithMotionModel, the per-frame tracking search).
 * cur: CurrentFrame (blocked = mvpMapPoints[i] && Observations() > 0).  Tcw_cur / Tcw_last: rows 0..2 of mTcw, row-major
 * 3x4 float.  Per LastFrame feature i (n_last of them): has_point = mvpMapPoints[i] && !mvbOutlier[i]; world = GetWorldPos()
 * (3 floats); octave = LastFrame.mvKeys[i].octave; angle = LastFrame.mvKeysUn[i].angle; desc = the MAP POINT's
 * GetDescriptor(); claims = Observations() > 0.
 * owner[cur.n]: index i of the LastFrame feature whose point ends up in CurrentFrame.mvpMapPoints[idx]; -1 = entry left
 * alone; -2 = entry set to NULL by the rotation-consistency cull (:1446-1458).  *n_matches = the return value. */
int orbm_search_by_projection_frame(const orbm_grid_view* cur, const float* Tcw_cur, const float* Tcw_last,
                                    float fx, float fy, float cx, float cy, float mbf, float mb,
                                    int n_last, const uint8_t* has_point, const float* world, const int* octave,
                                    const float* angle, const uint8_t* desc, const uint8_t* claims,
                                    float th, int mono, int check_orientation, int* owner, int* n_matches, int device);

/* Many-frame form of orbm_search_by_projection_frame (not in the reference, which tracks one camera per process): n_jobs independent
 * (CurrentFrame, LastFrame) pairs — several cameras / sessions, or a recorded sequence whose poses are already predicted — in ONE call:
 * one upload, three kernel launches (the serial claim replay of every pair runs in its own CTA), one download.  Every field has the
 * meaning of the same-named argument above; owner (cur->n entries) and n_matches are the outputs of job k.  Results are identical to
 * calling orbm_search_by_projection_frame once per job. */
typedef struct orbm_frame_search_job {
    const orbm_grid_view* cur;
    const float* Tcw_cur;
    const float* Tcw_last;
    float fx, fy, cx, cy, mbf, mb;
    int n_last;
    const uint8_t* has_point;
    const float* world;
    const int* octave;
    const float* angle;
    const uint8_t* desc;
    const uint8_t* claims;
    float th;
    int mono;
    int* owner;                     /* out */
    int n_matches;                  /* out */
} orbm_frame_search_job;
int orbm_search_by_projection_frame_batch(orbm_frame_search_job* jobs, int n_jobs, int check_orientation, int device);

/* Replaces int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched,
 * vector<int>& vnMatches12, int windowSize) (src/ORBmatcher.cc:408-523; monocular initialisation).
 * f2: F2.  Per F1 feature (n1): desc1, octave1 = mvKeysUn[i].octave, angle1, prev_xy = vbPrevMatched (2 floats each, updated
 * in place for matched features, :516-519).  matches12[n1] = vnMatches12.  *n_matches = the return value. */
int orbm_search_for_initialization(const orbm_grid_view* f2, int n1, const uint8_t* desc1, const int* octave1,
                                   const float* angle1, float* prev_xy, int window_size, float nnratio,
                                   int check_orientation, int* matches12, int* n_matches, int device);

/* The shared core of the two remaining SearchByProjection overloads — (Frame& CurrentFrame, KeyFrame*, sAlreadyFound, th,
 * ORBdist) src/ORBmatcher.cc:1465-1602 (relocalisation) and (KeyFrame*, Scw, vpPoints, vpMatched, th) :293-406 (loop closing):
 * per query an explicit window (u, v, r) with the level bounds given to GetFeaturesInArea (:1536 / the test at :380-381), the
 * nearest descriptor among features that are neither `target->blocked` nor taken by an earlier query (first wins ties),
 * accepted when its distance <= th_dist; an accepted match blocks its feature.  The per-point projection (Rcw*x+tcw, the depth
 * and viewing-angle tests, MapPoint::PredictScale) is done by the C++ class where the MapPoint objects live.
 * owner[target->n]: query index; -1 = entry left alone; -2 = set to NULL by the rotation cull (:1583-1599).  */
int orbm_search_windows(const orbm_grid_view* target, int nq, const uint8_t* active, const float* u, const float* v,
                        const float* r, const int* min_level, const int* max_level, const uint8_t* desc, const float* angle,
                        int th_dist, int check_orientation, int* owner, int* n_matches, int device);

/* Independent form of the window search, for the searches in which a match does not change what later queries may take:
 * SearchBySim3 (src/ORBmatcher.cc:1105-1329, both directions) and the candidate loops of Fuse (:828-972) and Fuse with a
 * similarity (:974-1103).  best_idx[q] = the feature in q's window (levels as above) with the smallest descriptor distance
 * (first wins ties) if that distance <= th_dist, else -1.  With ur / inv_level_sigma2 (both or neither) every candidate first
 * passes Fuse's reprojection gate (:913-944): e2 * inv_level_sigma2[octave] <= 7.8 with e2 = ex^2+ey^2+(ur-uright)^2 when the
 * feature has uright >= 0, else <= 5.99 with e2 = ex^2+ey^2.  target->blocked is ignored. */
int orbm_search_windows_best(const orbm_grid_view* target, int nq, const uint8_t* active, const float* u, const float* v,
                             const float* r, const int* min_level, const int* max_level, const uint8_t* desc, const float* ur,
                             const float* inv_level_sigma2, int th_dist, int* best_idx, int device);

/* POPC issue-rate microbenchmark (defines the matching roofline): returns measured 32-bit POPC results
 * per second on `device` over a register-resident loop. */
int orbm_popc_peak(int device, double* popc_per_second, double* sm_clock_hz_used);

/* Platform probe (a measurement aid for bench.py, not on the product path): the host->device rate the box sustains when n_dev GPUs
 * stream frames at once from pinned memory — one thread per GPU, plain cudaMemcpyAsync per chunk, all released together.
 * flags: 1 = write-combined pinned memory, 2 = also 20 % of the volume device->host concurrently, 4 = bind every copy thread (and its
 * first-touched pinned buffer) to the NUMA node of its GPU.  gbs_each[n_dev], *gbs_total: GB/s; numa_nodes[n_dev] (may be NULL): the
 * sysfs numa_node of every GPU (-1: the kernel reports none; -3: sysfs entry not visible). */
int orb_h2d_probe(int n_dev, const int* devices, size_t bytes_per_step, size_t chunk_bytes, int steps, int flags, double* gbs_each,
                  double* gbs_total, int* numa_nodes);
/* The same probe with a common start time for several processes (one per GPU, the layout of bench.py's e2e measurement): every process
 * allocates and warms up on its own and then waits for start_unix_ns (CLOCK_REALTIME, nanoseconds) before its timed copies, so that the
 * timed windows of all processes coincide.  *late_ms: by how much this process missed the start (0 = on time). */
int orb_h2d_probe_at(int n_dev, const int* devices, size_t bytes_per_step, size_t chunk_bytes, int steps, int flags,
                     long long start_unix_ns, double* gbs_each, double* gbs_total, int* numa_nodes, double* late_ms);

/* ------------------------------------------------------------------------------------------------------------------
 * Map archive: the fork's System::SaveMap / LoadMap file (src/System.cc:552-574) as a source of real keyframe
 * descriptor sets for the matcher.  Host-only: parses / writes the Boost binary archive (`no_header`, `oa << mpMap`) whose
 * field order is Map::save/load (src/Map.cc:31-134), MapPoint::save/load (src/MapPoint.cc:59-213), KeyFrame::save/load
 * (src/KeyFrame.cc:86-510) and the cv::Mat / cv::KeyPoint serializers of include/MapPoint.h:198-247.  The byte framing is
 * documented in orbslam_mapsave_b200/csrc/orb_map.cpp.  Boost is absent from this image: the class-info framing is restated from Boost's documentation (unpinned); the field sequence of MapPoint is pinned to the reference's own MapPoint::save (orbmap_mappoint_record).
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct orbmap_archive orbmap_archive;

typedef struct orbmap_info {
    int32_t n_mappoints, n_keyframes, n_origins;     /* mspMapPoints, mspKeyFrames, mvpKeyFrameOrigins */
    uint32_t test_data;                              /* 0xdeadbeef when the load sequence matched (src/Map.cc:127-131) */
    uint64_t max_kf_id;                              /* mnMaxKFid */
    int64_t total_features;                          /* sum of mDescriptors.rows over the keyframes */
    int64_t total_observations;                      /* sum of mObservations sizes over the map points */
    int64_t trailing_bytes;                          /* bytes Map::load leaves unread (Map::save's second map-point block) */
} orbmap_info;

typedef struct orbmap_keyframe_info {
    uint64_t id, frame_id, next_id, parent_id;       /* mnId, mnFrameId, KeyFrame::nNextId, mpParent->mnId */
    double timestamp;
    int32_t n;                                       /* N */
    int32_t n_keys, n_keys_un, n_uright, n_depth;    /* vector sizes as stored */
    int32_t desc_rows, desc_cols;                    /* mDescriptors: rows x bytes per row (32) */
    int32_t n_mappoint_slots;                        /* mvpMapPoints.size() */
    int32_t n_levels, n_scale_factors;               /* mnScaleLevels, mvScaleFactors.size() */
    int32_t grid_cols, grid_rows, min_x, min_y, max_x, max_y;
    int32_t n_connected, n_ordered, n_children, n_loop_edges;
    int32_t has_parent, is_bad, not_erase, to_be_erased, first_connection;
    float scale_factor, log_scale_factor, fx, fy, cx, cy, invfx, invfy, bf, b, th_depth, grid_inv_w, grid_inv_h, half_baseline;
} orbmap_keyframe_info;

/* Replaces System::LoadMap (src/System.cc:552-563) as far as the data goes: every field of the archive is kept. */
int orbmap_load(orbmap_archive** out, const char* path);
/* Replaces System::SaveMap (src/System.cc:565-574): writes the archive back, byte-identical for a loaded file. */
int orbmap_save(const orbmap_archive* ar, const char* path);
int orbmap_create(orbmap_archive** out);             /* an empty map to fill with orbmap_add_* */
void orbmap_destroy(orbmap_archive* ar);
int orbmap_get_info(const orbmap_archive* ar, orbmap_info* info);
/* group 0 = mspKeyFrames, 1 = mvpKeyFrameOrigins (stored as full KeyFrame records, src/Map.cc:52-56); i = position in the file */
int orbmap_keyframe_get_info(const orbmap_archive* ar, int group, int i, orbmap_keyframe_info* info);
/* Copies of the per-feature arrays; any pointer may be NULL.  keys / keys_un: n_keys / n_keys_un records (`size` is 0: the
 * fork's cv::KeyPoint serializer never stores it); desc: desc_rows x 32; mappoint_ids[n_mappoint_slots]: mnId or -1;
 * scale_factors / level_sigma2 / inv_level_sigma2: n_scale_factors floats; Tcw: 16 floats; K: 9 floats. */
int orbmap_keyframe_arrays(const orbmap_archive* ar, int group, int i, orbx_keypoint* keys, orbx_keypoint* keys_un, float* uright,
                           float* depth, uint8_t* desc, int64_t* mappoint_ids, float* scale_factors, float* level_sigma2,
                           float* inv_level_sigma2, float* Tcw, float* K);
/* Covisibility / spanning-tree ids (-1 = entry stored without an id). */
int orbmap_keyframe_links(const orbmap_archive* ar, int group, int i, int64_t* connected_ids, int32_t* connected_weights,
                          int64_t* ordered_ids, int32_t* ordered_weights, int64_t* children_ids, int64_t* loop_edge_ids);
/* mGrid as the CSR that orbm_grid_view takes (cells column-major like mGrid[col][row]); with both arrays NULL only the
 * counts are returned. */
int orbmap_keyframe_grid(const orbmap_archive* ar, int group, int i, int32_t* cell_offsets, int32_t* cell_features, int capacity,
                         int32_t* n_cells, int32_t* n_entries);
/* All map points in file order; any pointer may be NULL.  world_pos / normal: 3 floats each; desc: 32 bytes each;
 * ref_kf: mnId or -1; obs_offsets[n_mappoints + 1]: CSR offsets into orbmap_observations' arrays. */
int orbmap_mappoints(const orbmap_archive* ar, uint64_t* ids, float* world_pos, float* normal, uint8_t* desc, int64_t* ref_kf,
                     uint8_t* bad, int32_t* n_obs, int32_t* visible, int32_t* found, float* min_dist, float* max_dist,
                     int32_t* obs_offsets);
int orbmap_observations(const orbmap_archive* ar, int64_t* kf_ids, int64_t* feature_idx);
/* The fields of map point i exactly as MapPoint::save (src/MapPoint.cc:58-140) hands them to the archive — raw little-endian values in
 * call order, WITHOUT Boost's class-info framing.  tests compare it with a recording of the reference's own MapPoint::save.
 * out may be NULL (size query). */
int orbmap_mappoint_record(const orbmap_archive* ar, int i, uint8_t* out, int64_t capacity, int64_t* n_bytes);
/* The gather loop of MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:495-510) for every map point: the observed
 * keyframe rows back to back (capacity in descriptors), offsets[n_mappoints + 1]; the input of orbm_distinctive_descriptors.
 * With desc == NULL only the offsets / total are produced. */
int orbmap_observed_descriptors(const orbmap_archive* ar, uint8_t* desc, int32_t* offsets, int64_t capacity, int64_t* n_total);
/* Building a map from flat arrays (what Map::save would be fed from a live Map). */
int orbmap_add_mappoint(orbmap_archive* ar, uint64_t id, int64_t first_kf_id, const float* world_pos, const float* normal,
                        const uint8_t* desc, int64_t ref_kf_id, int n_obs, const int64_t* obs_kf_ids, const int64_t* obs_feature_idx,
                        int visible, int found, float min_dist, float max_dist);
/* Uses of `info`: id, frame_id, timestamp, n, n_levels, scale_factor, log_scale_factor, intrinsics, bounds, grid size (0 = 64 x 48),
 * has_parent / parent_id and the flags.  Twc / Ow / Cw follow KeyFrame::SetPose (src/KeyFrame.cc:792-806), mGrid follows
 * Frame::AssignFeaturesToGrid (src/Frame.cc:341-356). */
int orbmap_add_keyframe(orbmap_archive* ar, const orbmap_keyframe_info* info, const orbx_keypoint* keys, const orbx_keypoint* keys_un,
                        const float* uright, const float* depth, const uint8_t* desc, const int64_t* mappoint_ids,
                        const float* scale_factors, const float* level_sigma2, const float* inv_level_sigma2, const float* Tcw,
                        const float* K);
int orbmap_set_keyframe_links(orbmap_archive* ar, int i, int n_connected, const int64_t* connected_ids, const int32_t* connected_weights,
                              int n_ordered, const int64_t* ordered_ids, const int32_t* ordered_weights, int n_children,
                              const int64_t* children_ids, int n_loop_edges, const int64_t* loop_edge_ids);
int orbmap_add_origin(orbmap_archive* ar, int keyframe_index);

#ifdef __cplusplus
}
#endif
#endif /* ORB_B200_H_ */
