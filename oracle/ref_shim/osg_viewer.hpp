// ORACLE — TEST INFRASTRUCTURE ONLY. include/Map.hpp:12 includes the OpenSceneGraph viewer (modules/osg_viewer, needs OSG + boost);
// Map.hpp itself only needs boost::shared_mutex from what that header drags in. Nothing of the viewer is on the hot path.
#pragma once
#include <boost/thread.hpp>
#include <opencv2/core.hpp>
#include <string>
#include <vector>
namespace MCVSLAM {
// interface of modules/osg_viewer/osg_viewer.hpp:40-61 as far as src/Tracker.cpp:84-97,150-194 calls it; every method is a no-op
class osg_viewer {
   public:
    osg_viewer() {}
    explicit osg_viewer(const std::string&) {}
    void Draw(const cv::Mat, unsigned, unsigned, unsigned) {}
    void Draw(const std::vector<cv::Mat>, unsigned, unsigned, unsigned) {}
    void DrawCam(const cv::Mat, bool, unsigned, unsigned, unsigned) {}
    void DrawEssentialGraph(const std::vector<std::pair<cv::Mat, cv::Mat>>&, unsigned = 90, unsigned = 90, unsigned = 150) {}
    void Commit() {}
    bool IsStoped() { return false; }
    void RequestStop() {}
    void DrawPredictTracjectories(const std::vector<cv::Mat>&, const std::vector<bool>&, unsigned, unsigned, unsigned) {}
    void SetCurrentCamera(const cv::Mat) {}
    void SetCurViewFollow(const cv::Mat) {}
};
}  // namespace MCVSLAM
