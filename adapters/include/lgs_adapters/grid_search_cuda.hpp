/* grid_search_cuda.hpp
 *
 * Drop-in replacements for MyLidarGraphSlam::Mapping::ScanMatcherGridSearch
 * (mapping/scan_matcher_grid_search.hpp) and LoopDetectorGridSearch
 * (mapping/loop_detector_grid_search.hpp:17-47) whose exhaustive (y, x, theta) search runs on a
 * B200 through lgs_gs_match.  Same constructor parameters -- the ScorePixelAccurate object of the
 * reference is replaced by its two constructor arguments, as in LoopDetectorBranchBoundCuda -- same
 * ScanMatcher / LoopDetector interfaces, same results: the winner is returned as the loop counters
 * of the reference's accumulating loops and the pose is rebuilt by replaying those loops on the
 * host, so estimated poses are bit-identical.  The tail (Cost, ComputeCovariance) is the reference's
 * own host code, or lgs_cost_tail after UseDeviceCost().
 * Selected by the type strings "GridSearchCuda" (see create_cuda_backends.hpp / INTEGRATION.md). */
#ifndef LGS_ADAPTERS_GRID_SEARCH_CUDA_HPP
#define LGS_ADAPTERS_GRID_SEARCH_CUDA_HPP

#include <memory>
#include <vector>

#include "lgs_b200.h"
#include "my_lidar_graph_slam/mapping/cost_function.hpp"
#include "my_lidar_graph_slam/mapping/grid_map_builder.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

class ScanMatcherGridSearchCuda final : public ScanMatcher
{
public:
    /* ScorePixelAccurate(scoreUsableRangeMin, scoreUsableRangeMax), then the reference matcher's
     * own arguments (scan_matcher_grid_search.cpp:9-27) */
    ScanMatcherGridSearchCuda(const double scoreUsableRangeMin,
                              const double scoreUsableRangeMax,
                              const CostFuncPtr& costFunc,
                              const double rangeX,
                              const double rangeY,
                              const double rangeTheta,
                              const double stepX,
                              const double stepY,
                              const double stepTheta,
                              const int device = 0);
    ~ScanMatcherGridSearchCuda();

    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& queryInfo) override;
    ScanMatchingSummary OptimizePose(const GridMapType& gridMap,
                                     const Sensor::ScanDataPtr<double>& scanData,
                                     const RobotPose2D<double>& initialPose,
                                     const double normalizedScoreThreshold);
    /* All (scan, initial pose) pairs against ONE map in one device batch */
    std::vector<ScanMatchingSummary> OptimizePoses(
        const GridMapType& gridMap,
        const std::vector<Sensor::ScanDataPtr<double>>& scans,
        const std::vector<RobotPose2D<double>>& initialPoses,
        const double normalizedScoreThreshold);

    /* Evaluate the tail on the device (see ScanMatcherRealTimeCorrelativeCuda::UseDeviceCost) */
    void UseDeviceCost(const lgs_cost_params& params)
    { this->mCostParams = params; this->mDeviceCost = true; }

    const std::vector<lgs_match_result>& LastResults() const { return this->mLast; }

private:
    const CostFuncPtr             mCostFunc;
    const lgs_gs_params           mParams;
    lgs_ctx*                      mCtx;
    lgs_grid*                     mGrid;
    std::vector<double>           mDense;
    std::vector<double>           mOffsetsX, mOffsetsY, mOffsetsTheta;   /* the loops' dx, dy, dt */
    std::vector<lgs_match_result> mLast;
    bool                          mDeviceCost;
    lgs_cost_params               mCostParams;
};

class LoopDetectorGridSearchCuda final : public LoopDetector
{
public:
    LoopDetectorGridSearchCuda(const std::shared_ptr<ScanMatcherGridSearchCuda>& scanMatcher,
                               const double scoreThreshold);
    ~LoopDetectorGridSearchCuda() = default;

    void Detect(LoopDetectionQueryVector& loopDetectionQueries,
                LoopDetectionResultVector& loopDetectionResults) override;

private:
    std::shared_ptr<ScanMatcherGridSearchCuda> mScanMatcher;
    const double                               mScoreThreshold;
};

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */

#endif
