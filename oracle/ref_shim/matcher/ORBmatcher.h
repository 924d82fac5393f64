// TEST INFRASTRUCTURE (oracle/_ref matcher build).  Found before the reference's include/ORBmatcher.h on the include path: brings in
// the stand-in KeyFrame / Frame / MapPoint (ref_types.hpp), pre-defines the include guards of the reference's MapPoint.h / KeyFrame.h /
// Frame.h so that those three (which need Boost, DBoW2's vocabulary, Eigen ...) expand to nothing, and then continues to the
// reference's own, unmodified ORBmatcher.h.
#pragma once
#include "ref_types.hpp"
#define MAPPOINT_H
#define KEYFRAME_H
#define FRAME_H
#include_next "ORBmatcher.h"
