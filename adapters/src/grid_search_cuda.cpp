/* grid_search_cuda.cpp -- see the header */
#include "lgs_adapters/grid_search_cuda.hpp"

#include <cassert>
#include <cstdlib>
#include <iostream>
#include <limits>

#include "lgs_adapters/grid_map_flatten.hpp"
#include "my_lidar_graph_slam/util.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

namespace {
void Check(lgs_ctx* ctx, int rc, const char* what)
{
    /* Same failure behaviour as the reference: assertion + abort (util.hpp:29-38) */
    if (rc != LGS_OK) {
        std::cerr << "lgs_b200: " << what << " failed (" << rc << "): "
                  << (ctx ? lgs_ctx_last_error(ctx) : "no context") << std::endl;
        std::abort();
    }
}

/* The values the reference's loop `for (d = -range / 2; d <= range / 2; d += step)` visits
 * (scan_matcher_grid_search.cpp:59-76), replayed with the same running sum */
std::vector<double> LoopOffsets(const double range, const double step)
{
    std::vector<double> offsets;
    const double radius = range / 2.0;
    for (double d = -radius; d <= radius; d += step)
        offsets.push_back(d);
    return offsets;
}
} /* namespace */

ScanMatcherGridSearchCuda::ScanMatcherGridSearchCuda(
    const double scoreUsableRangeMin, const double scoreUsableRangeMax,
    const CostFuncPtr& costFunc, const double rangeX, const double rangeY, const double rangeTheta,
    const double stepX, const double stepY, const double stepTheta, const int device) :
    mCostFunc(costFunc),
    mParams { rangeX, rangeY, rangeTheta, stepX, stepY, stepTheta,
              scoreUsableRangeMin, scoreUsableRangeMax },
    mCtx(nullptr), mGrid(nullptr),
    mOffsetsX(LoopOffsets(rangeX, stepX)), mOffsetsY(LoopOffsets(rangeY, stepY)),
    mOffsetsTheta(LoopOffsets(rangeTheta, stepTheta)),
    mDeviceCost(false), mCostParams()
{
    Check(nullptr, lgs_ctx_create(device, &this->mCtx), "lgs_ctx_create (a B200 is required)");
}

ScanMatcherGridSearchCuda::~ScanMatcherGridSearchCuda()
{
    lgs_grid_destroy(this->mGrid);
    lgs_ctx_destroy(this->mCtx);
}

ScanMatchingSummary ScanMatcherGridSearchCuda::OptimizePose(const ScanMatchingQuery& queryInfo)
{
    return this->OptimizePose(queryInfo.mGridMap, queryInfo.mScanData, queryInfo.mInitialPose,
                              std::numeric_limits<double>::min());
}

ScanMatchingSummary ScanMatcherGridSearchCuda::OptimizePose(
    const GridMapType& gridMap, const Sensor::ScanDataPtr<double>& scanData,
    const RobotPose2D<double>& initialPose, const double normalizedScoreThreshold)
{
    return this->OptimizePoses(gridMap, { scanData }, { initialPose }, normalizedScoreThreshold).front();
}

std::vector<ScanMatchingSummary> ScanMatcherGridSearchCuda::OptimizePoses(
    const GridMapType& gridMap,
    const std::vector<Sensor::ScanDataPtr<double>>& scans,
    const std::vector<RobotPose2D<double>>& initialPoses,
    const double normalizedScoreThreshold)
{
    std::vector<ScanMatchingSummary> summaries;
    const int n = static_cast<int>(scans.size());
    if (n == 0)
        return summaries;

    /* the map: a dense device copy with a one-cell zero apron (GridMap::Value's unknown) */
    const int nx = gridMap.NumOfGridCellsX(), ny = gridMap.NumOfGridCellsY();
    int curNx = -1, curNy = -1;
    double curRes = 0.0;
    if (this->mGrid != nullptr)
        lgs_grid_info(this->mGrid, &curNx, &curNy, nullptr, nullptr, &curRes, nullptr);
    if (this->mGrid == nullptr || curRes != gridMap.Resolution()) {
        lgs_grid_destroy(this->mGrid);
        this->mGrid = nullptr;
        Check(this->mCtx, lgs_grid_create(this->mCtx, nx, ny, gridMap.MinPos().mX, gridMap.MinPos().mY,
              gridMap.Resolution(), 1, &this->mGrid), "lgs_grid_create");
    } else {
        /* same allocation when the size allows, new placement, all cells unknown */
        Check(this->mCtx, lgs_grid_resize(this->mGrid, nx, ny, gridMap.MinPos().mX, gridMap.MinPos().mY,
              1 << 30, 1 << 30), "lgs_grid_resize");
    }
    LgsB200::FlattenGridMap(gridMap, this->mDense);
    Check(this->mCtx, lgs_grid_upload(this->mGrid, this->mDense.data()), "lgs_grid_upload");

    std::vector<int> beamBegin(1, 0);
    std::vector<double> angles, ranges, poses, scanMin, scanMax, thresholds(n, normalizedScoreThreshold);
    std::vector<RobotPose2D<double>> sensorPoses;
    for (int k = 0; k < n; ++k) {
        const RobotPose2D<double> sensorPose = Compound(initialPoses[k], scans[k]->RelativeSensorPose());
        sensorPoses.push_back(sensorPose);
        poses.insert(poses.end(), { sensorPose.mX, sensorPose.mY, sensorPose.mTheta });
        angles.insert(angles.end(), scans[k]->Angles().begin(), scans[k]->Angles().end());
        ranges.insert(ranges.end(), scans[k]->Ranges().begin(), scans[k]->Ranges().end());
        beamBegin.push_back(static_cast<int>(angles.size()));
        scanMin.push_back(scans[k]->MinRange());
        scanMax.push_back(scans[k]->MaxRange());
    }
    const lgs_scan_batch batch { n, beamBegin.data(), angles.data(), ranges.data(), poses.data(),
                                 scanMin.data(), scanMax.data() };
    std::vector<const lgs_grid*> grids(n, this->mGrid);
    this->mLast.assign(n, lgs_match_result());
    Check(this->mCtx, lgs_gs_match(this->mCtx, &this->mParams, &batch, grids.data(), thresholds.data(),
          this->mLast.data(), nullptr), "lgs_gs_match");

    /* best sensor poses: the winning loop values added like the reference does (:78-80); the
     * initial sensor pose when nothing exceeded the threshold (:71) */
    std::vector<double> best;
    for (int k = 0; k < n; ++k) {
        const lgs_match_result& r = this->mLast[k];
        if (r.found)
            best.insert(best.end(), { sensorPoses[k].mX + this->mOffsetsX.at(r.ix),
                                      sensorPoses[k].mY + this->mOffsetsY.at(r.iy),
                                      sensorPoses[k].mTheta + this->mOffsetsTheta.at(r.it) });
        else
            best.insert(best.end(), { sensorPoses[k].mX, sensorPoses[k].mY, sensorPoses[k].mTheta });
    }

    /* tail (scan_matcher_grid_search.cpp:96-113): on the device, or the reference's host code */
    std::vector<double> normalizedCosts, covariances;
    if (this->mDeviceCost) {
        normalizedCosts.resize(n);
        covariances.resize(9 * static_cast<std::size_t>(n));
        Check(this->mCtx, lgs_cost_tail(this->mCtx, this->mGrid, &this->mCostParams, &batch, best.data(),
              normalizedCosts.data(), covariances.data(), nullptr), "lgs_cost_tail");
    }
    for (int k = 0; k < n; ++k) {
        const RobotPose2D<double> bestSensorPose { best[3 * k], best[3 * k + 1], best[3 * k + 2] };
        const RobotPose2D<double> estimatedPose =
            MoveBackward(bestSensorPose, scans[k]->RelativeSensorPose());
        if (this->mDeviceCost) {
            const double* c = covariances.data() + 9 * k;
            Eigen::Matrix3d covariance;
            covariance << c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8];
            summaries.emplace_back(this->mLast[k].found != 0, normalizedCosts[k], initialPoses[k],
                                   estimatedPose, covariance);
        } else {
            const double costVal = this->mCostFunc->Cost(gridMap, scans[k], bestSensorPose);
            summaries.emplace_back(
                this->mLast[k].found != 0, costVal / scans[k]->NumOfScans(), initialPoses[k], estimatedPose,
                this->mCostFunc->ComputeCovariance(gridMap, scans[k], bestSensorPose));
        }
    }
    return summaries;
}

LoopDetectorGridSearchCuda::LoopDetectorGridSearchCuda(
    const std::shared_ptr<ScanMatcherGridSearchCuda>& scanMatcher, const double scoreThreshold) :
    mScanMatcher(scanMatcher),
    mScoreThreshold(scoreThreshold)
{
    assert(scoreThreshold > 0.0);      /* loop_detector_grid_search.cpp:20-21 */
    assert(scoreThreshold <= 1.0);
}

void LoopDetectorGridSearchCuda::Detect(
    LoopDetectionQueryVector& loopDetectionQueries,
    LoopDetectionResultVector& loopDetectionResults)
{
    loopDetectionResults.clear();

    for (auto& query : loopDetectionQueries) {
        const auto& localMapInfo = query.mLocalMapInfo;
        const auto& localMapNode = query.mLocalMapNode;
        assert(localMapNode.Index() >= localMapInfo.mPoseGraphNodeIdxMin &&
               localMapNode.Index() <= localMapInfo.mPoseGraphNodeIdxMax);

        /* every node of the query against the query's local map, one device batch */
        std::vector<Sensor::ScanDataPtr<double>> scans;
        std::vector<RobotPose2D<double>> poses;
        for (const auto& node : query.mPoseGraphNodes) {
            scans.push_back(node.ScanData());
            poses.push_back(node.Pose());
        }
        const std::vector<ScanMatchingSummary> summaries =
            this->mScanMatcher->OptimizePoses(localMapInfo.mMap, scans, poses, this->mScoreThreshold);

        /* one loop closing edge per detected node, in node order (loop_detector_grid_search.cpp:48-78) */
        for (std::size_t k = 0; k < summaries.size(); ++k) {
            if (!summaries[k].mPoseFound)
                continue;
            loopDetectionResults.emplace_back(
                InverseCompound(localMapNode.Pose(), summaries[k].mEstimatedPose), localMapNode.Pose(),
                localMapNode.Index(), query.mPoseGraphNodes[k].Index(), summaries[k].mEstimatedCovariance);
        }
    }
}

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */
