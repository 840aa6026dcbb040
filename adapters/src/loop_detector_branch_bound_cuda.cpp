/* loop_detector_branch_bound_cuda.cpp */

#include "lgs_adapters/loop_detector_branch_bound_cuda.hpp"

#include <algorithm>
#include <cassert>

#include "lgs_adapters/grid_map_flatten.hpp"
#include "my_lidar_graph_slam/util.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

namespace {
void Check(lgs_ctx* ctx, int rc, const char* what)
{
    if (rc != LGS_OK) {
        std::cerr << "lgs_b200: " << what << " failed (" << rc << "): "
                  << (ctx ? lgs_ctx_last_error(ctx) : "no context") << std::endl;
        std::abort();
    }
}
void CheckGroup(lgs_group* group, int rc, const char* what)
{
    if (rc != LGS_OK) {
        std::cerr << "lgs_b200: " << what << " failed (" << rc << "): "
                  << (group ? lgs_group_last_error(group) : "no device group") << std::endl;
        std::abort();
    }
}
} /* namespace */

LoopDetectorBranchBoundCuda::LoopDetectorBranchBoundCuda(
    const double scoreUsableRangeMin, const double scoreUsableRangeMax,
    const CostFuncPtr& costFunc, const int nodeHeightMax, const double rangeX,
    const double rangeY, const double rangeTheta, const double scanRangeMax,
    const double scoreThreshold, const int device) :
    LoopDetectorBranchBoundCuda(scoreUsableRangeMin, scoreUsableRangeMax, costFunc, nodeHeightMax,
                                rangeX, rangeY, rangeTheta, scanRangeMax, scoreThreshold,
                                std::vector<int> { device })
{
}

LoopDetectorBranchBoundCuda::LoopDetectorBranchBoundCuda(
    const double scoreUsableRangeMin, const double scoreUsableRangeMax,
    const CostFuncPtr& costFunc, const int nodeHeightMax, const double rangeX,
    const double rangeY, const double rangeTheta, const double scanRangeMax,
    const double scoreThreshold, const std::vector<int>& devices) :
    mCostFunc(costFunc),
    mParams { nodeHeightMax, rangeX, rangeY, rangeTheta, scanRangeMax,
              scoreUsableRangeMin, scoreUsableRangeMax },
    mScoreThreshold(scoreThreshold), mGroup(nullptr), mDetector(nullptr),
    mDeviceCost(false), mCostParams()
{
    assert(scoreThreshold > 0.0);      /* loop_detector_branch_bound.cpp:20-21 */
    assert(scoreThreshold <= 1.0);
    assert(!devices.empty());
    CheckGroup(nullptr, lgs_group_create(devices.data(), static_cast<int>(devices.size()),
               &this->mGroup), "lgs_group_create (peer-capable B200s are required)");
    CheckGroup(this->mGroup, lgs_group_bb_create(this->mGroup, &this->mParams, &this->mDetector),
               "lgs_group_bb_create");
    const int numOfMembers = lgs_group_size(this->mGroup);
    this->mMemberLoad.assign(numOfMembers, 0.0);
    this->mMemberMaps.assign(numOfMembers, 0);
    /* lgs_match_result::n_scored per pair (not per batch): the cost the placement balances */
    for (int m = 0; m < numOfMembers; ++m)
        Check(lgs_group_ctx(this->mGroup, m), lgs_ctx_set_option(lgs_group_ctx(this->mGroup, m),
              "bb_count_nodes", 1.0), "lgs_ctx_set_option(bb_count_nodes)");
}

LoopDetectorBranchBoundCuda::~LoopDetectorBranchBoundCuda()
{
    lgs_group_bb_destroy(this->mDetector);
    for (auto& kv : this->mDeviceMaps) {
        lgs_pyramid_destroy(kv.second.mPyramid);
        lgs_grid_destroy(kv.second.mGrid);
    }
    lgs_group_destroy(this->mGroup);
}

/* Device pyramid of a local map: built on first use, rebuilt when the builder has reset
 * mPrecomputed after a loop closure (loop_detector_branch_bound.cpp:51-60).  A new local map goes
 * to the device that has scored the fewest nodes so far (ties: fewest maps, lowest index -- round
 * robin while nothing has been measured); a rebuilt one stays where it is.  Every LoopDetectionQuery carries its own COPY of LocalMapInfo, so a second
 * query of the same Detect() call that names the same local map still says mPrecomputed == false:
 * `builtThisCall` keeps it from destroying the pyramid the batch under construction points to. */
lgs_pyramid* LoopDetectorBranchBoundCuda::PyramidFor(
    LocalMapInfo& localMapInfo, std::vector<int>& builtThisCall)
{
    auto it = this->mDeviceMaps.find(localMapInfo.mIdx);
    const bool fresh = std::find(builtThisCall.begin(), builtThisCall.end(), localMapInfo.mIdx) !=
                       builtThisCall.end();
    if (it != this->mDeviceMaps.end() && (localMapInfo.mPrecomputed || fresh)) {
        localMapInfo.mPrecomputed = true;
        return it->second.mPyramid;
    }
    const int numOfMembers = lgs_group_size(this->mGroup);
    int member = 0;
    if (it != this->mDeviceMaps.end()) {
        member = it->second.mMember;
        lgs_pyramid_destroy(it->second.mPyramid);
        lgs_grid_destroy(it->second.mGrid);
        this->mDeviceMaps.erase(it);
    } else {
        for (int m = 1; m < numOfMembers; ++m)
            if (this->mMemberLoad[m] < this->mMemberLoad[member] ||
                (this->mMemberLoad[m] == this->mMemberLoad[member] &&
                 this->mMemberMaps[m] < this->mMemberMaps[member]))
                member = m;
        this->mMemberMaps[member]++;
    }
    const GridMapType& map = localMapInfo.mMap;
    lgs_ctx* ctx = lgs_group_ctx(this->mGroup, member);
    DeviceMap dev { nullptr, nullptr, member };
    Check(ctx, lgs_grid_create(ctx, map.NumOfGridCellsX(),
          map.NumOfGridCellsY(), map.MinPos().mX, map.MinPos().mY, map.Resolution(), 1,
          &dev.mGrid), "lgs_grid_create");
    LgsB200::FlattenGridMap(map, this->mDense);
    Check(ctx, lgs_grid_upload(dev.mGrid, this->mDense.data()), "lgs_grid_upload");
    Check(ctx, lgs_pyramid_create(ctx, dev.mGrid, this->mParams.node_height_max,
          &dev.mPyramid), "lgs_pyramid_create");
    this->mDeviceMaps.emplace(localMapInfo.mIdx, dev);
    builtThisCall.push_back(localMapInfo.mIdx);
    /* The host-side pyramids stay empty; the flag alone is what the SLAM parent copies back
     * (lidar_graph_slam.cpp:285-303) */
    localMapInfo.mPrecomputed = true;
    return dev.mPyramid;
}

void LoopDetectorBranchBoundCuda::Detect(
    LoopDetectionQueryVector& loopDetectionQueries,
    LoopDetectionResultVector& loopDetectionResults)
{
    loopDetectionResults.clear();
    this->mLast.clear();

    if (loopDetectionQueries.empty())
        return;

    /* Gather every (node, local map) pair of every query into one device batch */
    std::vector<int> beamBegin { 0 };
    std::vector<double> angles, ranges, poses, rangeMin, rangeMax, thresholds;
    std::vector<lgs_pyramid*> pyramids;
    std::vector<int> builtThisCall;

    for (auto& query : loopDetectionQueries) {
        auto& localMapInfo = query.mLocalMapInfo;
        assert(query.mLocalMapNode.Index() >= localMapInfo.mPoseGraphNodeIdxMin &&
               query.mLocalMapNode.Index() <= localMapInfo.mPoseGraphNodeIdxMax);
        assert(localMapInfo.mFinished);
        lgs_pyramid* pyramid = this->PyramidFor(localMapInfo, builtThisCall);

        for (const auto& node : query.mPoseGraphNodes) {
            const auto& scanData = node.ScanData();
            const RobotPose2D<double> sensorPose =
                Compound(node.Pose(), scanData->RelativeSensorPose());
            angles.insert(angles.end(), scanData->Angles().begin(), scanData->Angles().end());
            ranges.insert(ranges.end(), scanData->Ranges().begin(), scanData->Ranges().end());
            beamBegin.push_back(static_cast<int>(ranges.size()));
            poses.insert(poses.end(), { sensorPose.mX, sensorPose.mY, sensorPose.mTheta });
            rangeMin.push_back(scanData->MinRange());
            rangeMax.push_back(scanData->MaxRange());
            thresholds.push_back(this->mScoreThreshold);
            pyramids.push_back(pyramid);
        }
    }

    const int numOfPairs = static_cast<int>(pyramids.size());
    this->mLast.resize(numOfPairs);
    if (numOfPairs > 0) {
        const lgs_scan_batch scans { numOfPairs, beamBegin.data(), angles.data(), ranges.data(),
                                     poses.data(), rangeMin.data(), rangeMax.data() };
        /* Every pair on the device of its local map, all devices at once; one kernel launch per
         * device, records exchanged on the devices (lgs_group_bb_detect) */
        CheckGroup(this->mGroup, lgs_group_bb_detect(this->mDetector, &scans, numOfPairs, nullptr,
                   pyramids.data(), thresholds.data(), this->mLast.data()), "lgs_group_bb_detect");
    }

    /* What the pairs cost goes to the load of the device their local map lives on */
    {
        int pairIdx = 0;
        for (const auto& query : loopDetectionQueries) {
            const auto it = this->mDeviceMaps.find(query.mLocalMapInfo.mIdx);
            for (size_t k = 0; k < query.mPoseGraphNodes.size(); ++k, ++pairIdx)
                if (it != this->mDeviceMaps.end())
                    this->mMemberLoad[it->second.mMember] +=
                        static_cast<double>(this->mLast[pairIdx].n_scored);
        }
    }

    /* Host tail and loop closing edges in the reference's order
     * (loop_detector_branch_bound.cpp:63-88, scan_matcher_branch_bound.cpp:143-162) */
    int pairIdx = 0;
    for (auto& query : loopDetectionQueries) {
        const auto& localMap = query.mLocalMapInfo.mMap;
        const auto& localMapNode = query.mLocalMapNode;

        /* Accepted matches of this query (one local map) */
        std::vector<const PoseGraph::Node*> foundNodes;
        std::vector<double> best;
        for (const auto& node : query.mPoseGraphNodes) {
            const lgs_match_result& r = this->mLast[pairIdx++];
            if (!r.found)
                continue;
            const auto& scanData = node.ScanData();
            const RobotPose2D<double> sensorPose =
                Compound(node.Pose(), scanData->RelativeSensorPose());
            foundNodes.push_back(&node);
            best.insert(best.end(), { sensorPose.mX + r.ix * r.step_x,
                                      sensorPose.mY + r.iy * r.step_y,
                                      sensorPose.mTheta + r.it * r.step_t });
        }
        if (foundNodes.empty())
            continue;

        /* ComputeCovariance of every accepted match: on the device in one call against the local
         * map's device grid (lgs_cost_tail), or the reference's own code on the host */
        std::vector<double> covariances;
        if (this->mDeviceCost) {
            std::vector<int> tailBegin { 0 };
            std::vector<double> tailAngles, tailRanges, tailMin, tailMax;
            for (const auto* node : foundNodes) {
                const auto& scanData = node->ScanData();
                tailAngles.insert(tailAngles.end(), scanData->Angles().begin(), scanData->Angles().end());
                tailRanges.insert(tailRanges.end(), scanData->Ranges().begin(), scanData->Ranges().end());
                tailBegin.push_back(static_cast<int>(tailRanges.size()));
                tailMin.push_back(scanData->MinRange());
                tailMax.push_back(scanData->MaxRange());
            }
            const int numOfFound = static_cast<int>(foundNodes.size());
            const lgs_scan_batch tailScans { numOfFound, tailBegin.data(), tailAngles.data(),
                                             tailRanges.data(), best.data(), tailMin.data(), tailMax.data() };
            covariances.resize(9 * foundNodes.size());
            const DeviceMap& deviceMap = this->mDeviceMaps.at(query.mLocalMapInfo.mIdx);
            lgs_ctx* ctx = lgs_group_ctx(this->mGroup, deviceMap.mMember);
            Check(ctx, lgs_cost_tail(ctx, deviceMap.mGrid, &this->mCostParams,
                  &tailScans, best.data(), nullptr, covariances.data(), nullptr), "lgs_cost_tail");
        }

        for (std::size_t k = 0; k < foundNodes.size(); ++k) {
            const auto& node = *foundNodes[k];
            const auto& scanData = node.ScanData();
            const RobotPose2D<double> bestSensorPose { best[3 * k], best[3 * k + 1], best[3 * k + 2] };
            const RobotPose2D<double> correspondingPose =
                MoveBackward(bestSensorPose, scanData->RelativeSensorPose());
            Eigen::Matrix3d covarianceMatrix;
            if (this->mDeviceCost) {
                const double* c = covariances.data() + 9 * k;
                covarianceMatrix << c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8];
            } else {
                covarianceMatrix =
                    this->mCostFunc->ComputeCovariance(localMap, scanData, bestSensorPose);
            }

            const RobotPose2D<double> relativePose =
                InverseCompound(localMapNode.Pose(), correspondingPose);
            loopDetectionResults.emplace_back(
                relativePose, localMapNode.Pose(),
                localMapNode.Index(), node.Index(), covarianceMatrix);
        }
    }
}

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */
