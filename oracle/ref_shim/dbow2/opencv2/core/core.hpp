// TEST INFRASTRUCTURE (oracle/_ref DBoW2 build).  Stand-in for <opencv2/core/core.hpp> as far as the reference's
// Thirdparty/DBoW2/DBoW2/{TemplatedVocabulary.h, FORB.cpp, ScoringObject.cpp, BowVector.cpp, FeatureVector.cpp} need it: cv::Mat is the
// host's stand-in (orbslam_mapsave_b200/host/cv_compat.h); cv::FileStorage / cv::FileNode exist only so that the YAML save()/load()
// members of the class template compile (they are virtual, hence instantiated) — the text and binary loaders the reference actually
// uses (loadFromTextFile :1351-1440, loadFromBinaryFile :1467-1512) do not touch them, and calling them aborts.
#pragma once
#define ORB_B200_FORCE_CV_SHIM 1
#include <math.h>            // the real core.hpp brings these in; TemplatedVocabulary.h relies on it (pow, log, stringstream ...)
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include "../../../../../orbslam_mapsave_b200/host/cv_compat.h"

namespace cv {
struct FileNode {
    FileNode operator[](const char*) const { std::abort(); }
    FileNode operator[](const std::string&) const { std::abort(); }
    FileNode operator[](int) const { std::abort(); }
    size_t size() const { std::abort(); }
    operator int() const { std::abort(); }
    operator double() const { std::abort(); }
    operator std::string() const { std::abort(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { std::abort(); }
};
template <class T> inline FileStorage& operator<<(FileStorage&, const T&) { std::abort(); }
}  // namespace cv
