/* scan_matcher_real_time_correlative_cuda.hpp
 *
 * Drop-in replacement for MyLidarGraphSlam::Mapping::ScanMatcherRealTimeCorrelative
 * (mapping/scan_matcher_real_time_correlative.hpp:16-93) whose exhaustive (x, y, theta) sweep
 * runs on a B200 through the C ABI in include/lgs_b200.h.  Same constructor parameters, same
 * ScanMatcher interface, same results (winning indices bit-exact, hence identical poses).  The
 * tail (Cost, ComputeCovariance) is the reference's own code on the host unless UseDeviceCost()
 * hands the CostGreedyEndpoint parameters over, in which case it runs on the device too
 * (lgs_cost_tail, bit-identical); MoveBackward stays the reference's.
 * Selected by the type string "RealTimeCorrelativeCuda" (see INTEGRATION.md). */
#ifndef LGS_ADAPTERS_SCAN_MATCHER_REAL_TIME_CORRELATIVE_CUDA_HPP
#define LGS_ADAPTERS_SCAN_MATCHER_REAL_TIME_CORRELATIVE_CUDA_HPP

#include <vector>

#include "lgs_b200.h"
#include "my_lidar_graph_slam/mapping/cost_function.hpp"
#include "my_lidar_graph_slam/mapping/grid_map_builder.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

class ScanMatcherRealTimeCorrelativeCuda final : public ScanMatcher
{
public:
    ScanMatcherRealTimeCorrelativeCuda(const CostFuncPtr& costFunc,
                                       const int lowResolution,
                                       const double rangeX,
                                       const double rangeY,
                                       const double rangeTheta,
                                       const double scanRangeMax,
                                       const int device = 0);
    ~ScanMatcherRealTimeCorrelativeCuda();

    /* Optimize the robot pose by scan matching (threshold = DBL_MIN, whole window) */
    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& queryInfo) override;

    /* Same as the reference's 5-argument overload minus the precomputed map, which now lives
     * on the device */
    ScanMatchingSummary OptimizePose(const GridMapType& gridMap,
                                     const Sensor::ScanDataPtr<double>& scanData,
                                     const RobotPose2D<double>& initialPose,
                                     const double normalizedScoreThreshold);

    /* Same match against a map that is ALREADY on the device (GridMapBuilderCuda::DeviceLatestMap):
     * no flattening and no upload of the map, one device-to-device copy instead.  Needs
     * UseDeviceCost(), because there is no host map for the reference's cost function to read.
     * The caller guarantees that `deviceMap` is not being integrated into during the call (in the
     * reference: take it under the same lock as GetLatestPoseAndMap, lidar_graph_slam.cpp:306-315) */
    ScanMatchingSummary OptimizePose(const lgs_grid* deviceMap,
                                     const Sensor::ScanDataPtr<double>& scanData,
                                     const RobotPose2D<double>& initialPose,
                                     const double normalizedScoreThreshold);

    /* All (scan, initial pose) pairs against ONE map in one device batch (the map is uploaded and
     * its coarse map computed once): what LoopDetectorRealTimeCorrelativeCuda issues per query */
    std::vector<ScanMatchingSummary> OptimizePoses(
        const GridMapType& gridMap,
        const std::vector<Sensor::ScanDataPtr<double>>& scans,
        const std::vector<RobotPose2D<double>>& initialPoses,
        const double normalizedScoreThreshold);

    /* Evaluate the tail (normalised cost + covariance) on the device from now on.  `params` are
     * the constructor arguments of the CostGreedyEndpoint this matcher was given (the class keeps
     * them private, so the factory that built it passes them here as well). */
    void UseDeviceCost(const lgs_cost_params& params)
    { this->mCostParams = params; this->mDeviceCost = true; }

    /* Details of the last match (window indices, score, device counters) */
    const lgs_match_result& LastResult() const { return this->mLast; }

private:
    void EnsureGrids(int nx, int ny, double minX, double minY, double res);
    void UploadMap(const GridMapType& gridMap);
    /* sweep + tail against the map in mGrid / mCoarse; hostMap may be null with the device tail */
    ScanMatchingSummary MatchUploaded(const GridMapType* hostMap,
                                      const Sensor::ScanDataPtr<double>& scanData,
                                      const RobotPose2D<double>& initialPose,
                                      const double normalizedScoreThreshold);
    /* Normalised cost + covariance of every (scan, best sensor pose) pair on the device */
    void DeviceTail(const lgs_scan_batch& scans, const std::vector<double>& bestPoses,
                    std::vector<double>& normalizedCosts, std::vector<double>& covariances);

    const CostFuncPtr   mCostFunc;
    const int           mLowResolution;
    const double        mRangeX;
    const double        mRangeY;
    const double        mRangeTheta;
    const double        mScanRangeMax;
    lgs_ctx*            mCtx;
    lgs_grid*           mGrid;
    lgs_grid*           mCoarse;
    lgs_rtcsm_batch*    mBatch;
    std::vector<double> mDense;
    lgs_match_result    mLast;
    bool                mDeviceCost;
    lgs_cost_params     mCostParams;
};

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */

#endif
