// ORACLE — TEST INFRASTRUCTURE ONLY. include/PoseEstimation.hpp = g2o pose optimisation / bundle adjustment (out of scope, needs
// g2o + opencv calib3d). src/Map.cpp:145 and src/Tracker.cpp:73-130 call it from control flow the parity tests never enter
// (Map::AddKeyFrame / Tracker::Track); the matching front-ends under test (Map::Fuse, Tracker::Wnd_Track / Bow_Track,
// Map::ComputeF12) do not. Calling any of these aborts.
#pragma once
#include <cstdlib>
#include <unordered_set>
#include "BAoptimizer.hpp"
namespace cv { enum { FM_7POINT = 1, FM_8POINT = 2, FM_LMEDS = 4, FM_RANSAC = 8 }; }
namespace MCVSLAM {
#define MIN_DISPARITY 1
#define CHI2_STEREO_THRESHOLD 7.815
#define CHI2_MONO_THRESHOLD 5.991
class PoseEstimation {
   public:
    static cv::Mat _2d2d(const ObjectRef&, const ObjectRef&, const std::vector<cv::DMatch>&, uint = cv::FM_8POINT) { std::abort(); }
    static int PoseOptimization(const KeyFrame&) { std::abort(); }
    template <typename F> static int BoundleAdjustment(const std::unordered_set<KeyFrame>&, const std::unordered_set<KeyFrame>&, uint, F&&) { std::abort(); }
    static bool FilterCallBack_Chi2(const BAoptimizer::EdgeInfoMation&) { std::abort(); }
};
}  // namespace MCVSLAM
