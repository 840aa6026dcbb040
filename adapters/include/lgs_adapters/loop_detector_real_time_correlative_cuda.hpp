/* loop_detector_real_time_correlative_cuda.hpp
 *
 * Drop-in replacement for MyLidarGraphSlam::Mapping::LoopDetectorRealTimeCorrelative
 * (mapping/loop_detector_real_time_correlative.hpp:17-50): same constructor shape (a matcher +
 * the normalised score threshold), same LoopDetector interface, same results.  Every query
 * (one finished local map + its candidate nodes) becomes ONE device batch: the local map is
 * uploaded and its coarse map computed once (the reference's ComputeCoarserMap,
 * loop_detector_real_time_correlative.cpp:49-57), then all nodes are matched together.
 * mPrecomputedMaps is left empty -- the coarse map lives on the device -- and mPrecomputed is set,
 * which is all LidarGraphSlam::UpdatePrecomputedGridMaps copies back (lidar_graph_slam.cpp:285-303).
 * Selected by the type string "RealTimeCorrelativeCuda" in Backend.LoopDetectorType. */
#ifndef LGS_ADAPTERS_LOOP_DETECTOR_REAL_TIME_CORRELATIVE_CUDA_HPP
#define LGS_ADAPTERS_LOOP_DETECTOR_REAL_TIME_CORRELATIVE_CUDA_HPP

#include <memory>

#include "lgs_adapters/scan_matcher_real_time_correlative_cuda.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

class LoopDetectorRealTimeCorrelativeCuda final : public LoopDetector
{
public:
    LoopDetectorRealTimeCorrelativeCuda(
        const std::shared_ptr<ScanMatcherRealTimeCorrelativeCuda>& scanMatcher,
        const double scoreThreshold);
    ~LoopDetectorRealTimeCorrelativeCuda() = default;

    void Detect(LoopDetectionQueryVector& loopDetectionQueries,
                LoopDetectionResultVector& loopDetectionResults) override;

private:
    std::shared_ptr<ScanMatcherRealTimeCorrelativeCuda> mScanMatcher;
    const double                                        mScoreThreshold;
};

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */

#endif
