/* loop_detector_branch_bound_cuda.hpp
 *
 * Drop-in replacement for LoopDetectorBranchBound + ScanMatcherBranchBound + ScorePixelAccurate
 * (mapping/loop_detector_branch_bound.hpp:17-51, scan_matcher_branch_bound.hpp:19-98,
 * score_function_pixel_accurate.hpp) behind the unchanged LoopDetector interface
 * (mapping/loop_detector.hpp:92-107).  All (node, local map) pairs of one Detect() call are
 * searched as ONE batch; the win-max pyramids stay on the device, cached per local map index and
 * rebuilt when the builder resets LocalMapInfo::mPrecomputed (grid_map_builder.cpp:70-72).
 * Selected by the type string "BranchBoundCuda".
 *
 * Several GPUs ("Devices": [0, 1, ...]): local map i lives on device i mod G, the pairs of a Detect()
 * call run where their local map lives -- all devices concurrently, driven from this one process by
 * one host thread per device (lgs_group) -- and every device's kernel stores its result records
 * straight into the first device's gather buffer over NVLink.  The results do not depend on the
 * number of devices. */
#ifndef LGS_ADAPTERS_LOOP_DETECTOR_BRANCH_BOUND_CUDA_HPP
#define LGS_ADAPTERS_LOOP_DETECTOR_BRANCH_BOUND_CUDA_HPP

#include <map>
#include <vector>

#include "lgs_b200.h"
#include "my_lidar_graph_slam/mapping/cost_function.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

class LoopDetectorBranchBoundCuda final : public LoopDetector
{
public:
    /* Parameters: ScorePixelAccurate(usableRangeMin, usableRangeMax), then
     * ScanMatcherBranchBound(costFunc, nodeHeightMax, rangeX, rangeY, rangeTheta, scanRangeMax),
     * then LoopDetectorBranchBound(scoreThreshold) -- in the reference's own order */
    LoopDetectorBranchBoundCuda(const double scoreUsableRangeMin,
                                const double scoreUsableRangeMax,
                                const CostFuncPtr& costFunc,
                                const int nodeHeightMax,
                                const double rangeX,
                                const double rangeY,
                                const double rangeTheta,
                                const double scanRangeMax,
                                const double scoreThreshold,
                                const int device = 0);
    /* Same, on several devices of one box */
    LoopDetectorBranchBoundCuda(const double scoreUsableRangeMin,
                                const double scoreUsableRangeMax,
                                const CostFuncPtr& costFunc,
                                const int nodeHeightMax,
                                const double rangeX,
                                const double rangeY,
                                const double rangeTheta,
                                const double scanRangeMax,
                                const double scoreThreshold,
                                const std::vector<int>& devices);
    ~LoopDetectorBranchBoundCuda();

    void Detect(LoopDetectionQueryVector& loopDetectionQueries,
                LoopDetectionResultVector& loopDetectionResults) override;

    /* Evaluate the covariance of the accepted matches on the device from now on; `params` are the
     * constructor arguments of the CostGreedyEndpoint this detector was given */
    void UseDeviceCost(const lgs_cost_params& params)
    { this->mCostParams = params; this->mDeviceCost = true; }

    /* Per-pair device results of the last Detect() (query-major, node-minor order) */
    const std::vector<lgs_match_result>& LastResults() const { return this->mLast; }

private:
    struct DeviceMap { lgs_grid* mGrid; lgs_pyramid* mPyramid; int mMember; };

    lgs_pyramid* PyramidFor(LocalMapInfo& localMapInfo, std::vector<int>& builtThisCall);

    const CostFuncPtr             mCostFunc;
    const lgs_bb_params           mParams;
    const double                  mScoreThreshold;
    lgs_group*                    mGroup;
    lgs_group_bb*                 mDetector;
    std::map<int, DeviceMap>      mDeviceMaps;
    std::vector<double>           mDense;
    std::vector<lgs_match_result> mLast;
    /* Nodes scored so far against the local maps of every device: a new local map goes to the
     * device that has carried the least work (the true loop candidates cost ~40x the others) */
    std::vector<double>           mMemberLoad;
    std::vector<int>              mMemberMaps;
    bool                          mDeviceCost;
    lgs_cost_params               mCostParams;
};

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */

#endif
