// ORACLE — TEST INFRASTRUCTURE ONLY. include/BAoptimizer.hpp is the g2o bundle-adjustment wrapper (out of scope;
// g2o is not installed). src/Map.cpp:12 includes it; only the type name is needed for PoseEstimation's callback signature.
#pragma once
#include "Frame.hpp"
#include "Map.hpp"
#include "MapPoint.hpp"
#include "Object.hpp"
namespace MCVSLAM {
class BAoptimizer {
   public:
    struct EdgeInfoMation { double chi2 = 0; bool is_stereo = false; };
};
}  // namespace MCVSLAM
