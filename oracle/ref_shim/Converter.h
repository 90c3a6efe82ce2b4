// ORACLE — TEST INFRASTRUCTURE ONLY. include/Converter.h needs g2o + Eigen/Dense. src/Map.cpp:563 (Map::ComputeF12) uses exactly one
// member, toSkewSymmetricMatrix (defined in src/Converter.cc:182-184, a translation unit that cannot be compiled here for the same
// reason); the 3x3 cross-product matrix [v]x of a CV_32F 3-vector is therefore provided inline.
#pragma once
#include <opencv2/core/core.hpp>
namespace MCVSLAM {
class Converter {
   public:
    static cv::Mat toSkewSymmetricMatrix(const cv::Mat& v) {
        return (cv::Mat_<float>(3, 3) << 0, -v.at<float>(2), v.at<float>(1), v.at<float>(2), 0, -v.at<float>(0), -v.at<float>(1), v.at<float>(0), 0);
    }
};
}  // namespace MCVSLAM
