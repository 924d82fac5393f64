// ref_mappoint_driver.cpp — TEST INFRASTRUCTURE.  C entry points around the reference's own ORB_SLAM2::MapPoint
// (/root/reference/src/MapPoint.cc + include/MapPoint.h compiled unmodified, see prelude.hpp).  Pinned through these:
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:483-548) with the descriptor distance of the reference's own
// ORBmatcher::DescriptorDistance (forwarded to oracle/_ref/libref_orbmatcher.so), and the FIELD SEQUENCE of MapPoint::save (:58-140).
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include "MapPoint.h"

namespace ORB_SLAM2 {
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
    typedef int (*fn_t)(const unsigned char*, const unsigned char*);
    static fn_t fn = [] {
        Dl_info info;
        dladdr((void*)&ORBmatcher::DescriptorDistance, &info);
        std::string dir(info.dli_fname);
        dir = dir.substr(0, dir.find_last_of('/'));
        void* h = dlopen((dir + "/libref_orbmatcher.so").c_str(), RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
        fn_t f = h ? (fn_t)dlsym(h, "refm_descriptor_distance") : nullptr;
        if (!f) { fprintf(stderr, "ref_mappoint: cannot bind refm_descriptor_distance: %s\n", dlerror()); abort(); }
        return f;
    }();
    return fn(a.ptr(), b.ptr());
}
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

extern "C" {

// One map point observed by n keyframes; observation i is feature 0 of keyframe i, which holds descriptor desc[i] and may be bad.
// The reference walks its std::map<KeyFrame*, size_t> in POINTER order: the keyframes are allocated in one array so that pointer order
// is index order (the order the oracle and the product take).  Returns the index (among ALL n observations) of the keyframe whose row
// the reference copies into mDescriptor, or -1 when it leaves mDescriptor alone; out32 receives mDescriptor.
int refp_distinctive(const unsigned char* desc, const unsigned char* kf_bad, int n, int point_bad, unsigned char* out32) {
    Map map;
    std::vector<KeyFrame> kfs(n > 0 ? n : 1);
    for (int i = 0; i < n; i++) {
        kfs[i].mnId = i;
        kfs[i].mvuRight.assign(1, -1.f);
        kfs[i].mvpMapPoints.assign(1, static_cast<MapPoint*>(NULL));
        kfs[i].mDescriptors.create(1, 32, CV_8U);
        memcpy(kfs[i].mDescriptors.ptr(), desc + 32 * (size_t)i, 32);
        kfs[i].mbBad = kf_bad && kf_bad[i];
    }
    cv::Mat pos(3, 1, CV_32F);
    MapPoint mp(pos, &kfs[0], &map);
    for (int i = 0; i < n; i++) mp.AddObservation(&kfs[i], 0);
    if (point_bad) mp.SetBadFlag();
    mp.ComputeDistinctiveDescriptors();
    cv::Mat d = mp.GetDescriptor();
    if (d.empty()) return -1;
    memcpy(out32, d.ptr(), 32);
    for (int i = 0; i < n; i++)
        if (!kfs[i].mbBad && memcmp(kfs[i].mDescriptors.ptr(), out32, 32) == 0) return i;
    return -2;
}

// The byte / field sequence MapPoint::save (src/MapPoint.cc:58-140) produces for a point with `n_obs` observations, through the
// recording archive (raw little-endian values in call order, no Boost framing).  fields: one tag per item, ';'-separated.
int refp_save_fields(int n_obs, int has_ref, int track, unsigned char* bytes, int cap, char* fields, int fcap) {
    Map map;
    std::vector<KeyFrame> kfs(n_obs > 0 ? n_obs : 1);
    for (int i = 0; i < (int)kfs.size(); i++) {
        kfs[i].mnId = 100 + 7 * i;
        kfs[i].mnFrameId = kfs[i].mnId;
        kfs[i].mvuRight.assign(4, -1.f);
        kfs[i].mvpMapPoints.assign(4, static_cast<MapPoint*>(NULL));
        kfs[i].mDescriptors.create(4, 32, CV_8U);
        for (int r = 0; r < 4; r++) memset(kfs[i].mDescriptors.ptr(r), 16 * i + r, 32);
    }
    cv::Mat pos(3, 1, CV_32F);
    pos.at<float>(0) = 1.5f; pos.at<float>(1) = -2.25f; pos.at<float>(2) = 8.f;
    MapPoint::nNextId = 41;
    MapPoint mp(pos, &kfs[0], &map);
    for (int i = 0; i < n_obs; i++) mp.AddObservation(&kfs[i], (size_t)(i % 4));
    mp.ComputeDistinctiveDescriptors();
    // per-frame tracking scratch (the constructor leaves most of it uninitialised): `track` != 0 fills it with distinct values so that a
    // test can see each field land in its slot; 0 zeroes it, which is what a map built through orbmap_add_mappoint holds
    const float ft = track ? 1.f : 0.f;
    const unsigned long ut = track ? 1u : 0u;
    mp.mTrackProjX = 11.f * ft; mp.mTrackProjY = 12.f * ft; mp.mTrackProjXR = 13.f * ft; mp.mnTrackScaleLevel = 3 * (int)ut; mp.mTrackViewCos = 0.5f * ft;
    mp.mnTrackReferenceForFrame = 21 * ut; mp.mnLastFrameSeen = 22 * ut; mp.mnBALocalForKF = 23 * ut; mp.mnFuseCandidateForKF = 24 * ut;
    mp.mnLoopPointForKF = 25 * ut; mp.mnCorrectedByKF = 26 * ut; mp.mnCorrectedReference = 27 * ut; mp.mnBAGlobalForKF = 28 * ut;
    if (!has_ref) { std::vector<KeyFrame*> none; mp.SetObservations(none); }      // leaves mpRefKF NULL (:272-281)
    boost::archive::binary_oarchive oa;
    boost::serialization::access::save(oa, mp, 0u);
    std::string f;
    for (size_t i = 0; i < oa.fields.size(); i++) { f += oa.fields[i]; f += ';'; }
    if ((int)oa.bytes.size() > cap || (int)f.size() + 1 > fcap) return -(int)oa.bytes.size();
    memcpy(bytes, oa.bytes.data(), oa.bytes.size());
    memcpy(fields, f.c_str(), f.size() + 1);
    return (int)oa.bytes.size();
}

}  // extern "C"
