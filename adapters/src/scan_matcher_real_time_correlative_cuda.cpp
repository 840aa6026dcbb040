/* scan_matcher_real_time_correlative_cuda.cpp */

#include "lgs_adapters/scan_matcher_real_time_correlative_cuda.hpp"

#include <cmath>
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
} /* namespace */

ScanMatcherRealTimeCorrelativeCuda::ScanMatcherRealTimeCorrelativeCuda(
    const CostFuncPtr& costFunc, const int lowResolution, const double rangeX,
    const double rangeY, const double rangeTheta, const double scanRangeMax,
    const int device) :
    mCostFunc(costFunc), mLowResolution(lowResolution), mRangeX(rangeX),
    mRangeY(rangeY), mRangeTheta(rangeTheta), mScanRangeMax(scanRangeMax),
    mCtx(nullptr), mGrid(nullptr), mCoarse(nullptr), mBatch(nullptr), mLast(),
    mDeviceCost(false), mCostParams()
{
    Check(nullptr, lgs_ctx_create(device, &this->mCtx), "lgs_ctx_create (a B200 is required)");
    const lgs_rtcsm_params params { lowResolution, rangeX, rangeY, rangeTheta, scanRangeMax };
    Check(this->mCtx, lgs_rtcsm_batch_create(this->mCtx, &params, &this->mBatch),
          "lgs_rtcsm_batch_create");
}

ScanMatcherRealTimeCorrelativeCuda::~ScanMatcherRealTimeCorrelativeCuda()
{
    lgs_rtcsm_batch_destroy(this->mBatch);
    lgs_grid_destroy(this->mCoarse);
    lgs_grid_destroy(this->mGrid);
    lgs_ctx_destroy(this->mCtx);
}

void ScanMatcherRealTimeCorrelativeCuda::EnsureGrids(
    const int nx, const int ny, const double minX, const double minY, const double res)
{
    int curNx = -1, curNy = -1;
    double curMinX = 0.0, curMinY = 0.0, curRes = 0.0;
    if (this->mGrid != nullptr)
        lgs_grid_info(this->mGrid, &curNx, &curNy, &curMinX, &curMinY, &curRes, nullptr);
    if (curNx != nx || curNy != ny || curMinX != minX || curMinY != minY || curRes != res) {
        lgs_grid_destroy(this->mCoarse);
        lgs_grid_destroy(this->mGrid);
        /* The zero apron must cover the fine search window (H2: it is asymmetric) */
        const int winX = static_cast<int>(std::ceil(0.5 * this->mRangeX / res));
        const int winY = static_cast<int>(std::ceil(0.5 * this->mRangeY / res));
        const int spanX = ((2 * winX) / this->mLowResolution + 1) * this->mLowResolution;
        const int spanY = ((2 * winY) / this->mLowResolution + 1) * this->mLowResolution;
        const int apron = std::max(spanX, spanY);
        Check(this->mCtx, lgs_grid_create(this->mCtx, nx, ny, minX, minY, res, apron, &this->mGrid),
              "lgs_grid_create");
        Check(this->mCtx, lgs_grid_create(this->mCtx, nx, ny, minX, minY, res, apron, &this->mCoarse),
              "lgs_grid_create");
    }
}

void ScanMatcherRealTimeCorrelativeCuda::UploadMap(const GridMapType& gridMap)
{
    this->EnsureGrids(gridMap.NumOfGridCellsX(), gridMap.NumOfGridCellsY(), gridMap.MinPos().mX,
                      gridMap.MinPos().mY, gridMap.Resolution());
    LgsB200::FlattenGridMap(gridMap, this->mDense);
    Check(this->mCtx, lgs_grid_upload(this->mGrid, this->mDense.data()), "lgs_grid_upload");
    /* ComputeCoarserMap (scan_matcher_real_time_correlative.cpp:148-153) on the device */
    Check(this->mCtx, lgs_precompute(this->mCtx, this->mGrid, this->mLowResolution,
          this->mCoarse), "lgs_precompute");
}

namespace {
Eigen::Matrix3d ToMatrix(const double* c)
{
    Eigen::Matrix3d m;
    m << c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8];
    return m;
}
} /* namespace */

void ScanMatcherRealTimeCorrelativeCuda::DeviceTail(
    const lgs_scan_batch& scans, const std::vector<double>& bestPoses,
    std::vector<double>& normalizedCosts, std::vector<double>& covariances)
{
    normalizedCosts.resize(scans.n_scans);
    covariances.resize(9 * static_cast<std::size_t>(scans.n_scans));
    Check(this->mCtx, lgs_cost_tail(this->mCtx, this->mGrid, &this->mCostParams, &scans,
          bestPoses.data(), normalizedCosts.data(), covariances.data(), nullptr), "lgs_cost_tail");
}

ScanMatchingSummary ScanMatcherRealTimeCorrelativeCuda::OptimizePose(
    const ScanMatchingQuery& queryInfo)
{
    return this->OptimizePose(queryInfo.mGridMap, queryInfo.mScanData,
                              queryInfo.mInitialPose,
                              std::numeric_limits<double>::min());
}

ScanMatchingSummary ScanMatcherRealTimeCorrelativeCuda::OptimizePose(
    const GridMapType& gridMap,
    const Sensor::ScanDataPtr<double>& scanData,
    const RobotPose2D<double>& initialPose,
    const double normalizedScoreThreshold)
{
    this->UploadMap(gridMap);
    return this->MatchUploaded(&gridMap, scanData, initialPose, normalizedScoreThreshold);
}

ScanMatchingSummary ScanMatcherRealTimeCorrelativeCuda::OptimizePose(
    const lgs_grid* deviceMap,
    const Sensor::ScanDataPtr<double>& scanData,
    const RobotPose2D<double>& initialPose,
    const double normalizedScoreThreshold)
{
    if (!this->mDeviceCost) {
        std::cerr << "lgs_b200: matching against a device-resident map needs UseDeviceCost() "
                     "(there is no host map for the reference's cost function)" << std::endl;
        std::abort();
    }
    int nx = 0, ny = 0;
    double minX = 0.0, minY = 0.0, res = 0.0;
    Check(this->mCtx, lgs_grid_info(deviceMap, &nx, &ny, &minX, &minY, &res, nullptr), "lgs_grid_info");
    this->EnsureGrids(nx, ny, minX, minY, res);
    Check(this->mCtx, lgs_grid_copy(deviceMap, this->mGrid), "lgs_grid_copy");
    Check(this->mCtx, lgs_precompute(this->mCtx, this->mGrid, this->mLowResolution,
          this->mCoarse), "lgs_precompute");
    return this->MatchUploaded(nullptr, scanData, initialPose, normalizedScoreThreshold);
}

ScanMatchingSummary ScanMatcherRealTimeCorrelativeCuda::MatchUploaded(
    const GridMapType* hostMap,
    const Sensor::ScanDataPtr<double>& scanData,
    const RobotPose2D<double>& initialPose,
    const double normalizedScoreThreshold)
{

    const RobotPose2D<double> sensorPose =
        Compound(initialPose, scanData->RelativeSensorPose());
    const int beamBegin[2] = { 0, static_cast<int>(scanData->NumOfScans()) };
    const double pose[3] = { sensorPose.mX, sensorPose.mY, sensorPose.mTheta };
    const lgs_scan_batch scans { 1, beamBegin, scanData->Angles().data(),
                                 scanData->Ranges().data(), pose, nullptr, nullptr };
    Check(this->mCtx, lgs_rtcsm_batch_upload(this->mBatch, this->mGrid, &scans,
          &normalizedScoreThreshold), "lgs_rtcsm_batch_upload");
    Check(this->mCtx, lgs_rtcsm_batch_run(this->mBatch, this->mGrid, this->mCoarse),
          "lgs_rtcsm_batch_run");
    Check(this->mCtx, lgs_rtcsm_batch_results(this->mBatch, this->mGrid, this->mCoarse,
          &this->mLast), "lgs_rtcsm_batch_results");

    /* From here on: the reference's own host tail
     * (scan_matcher_real_time_correlative.cpp:118-144) */
    const bool poseFound = this->mLast.found != 0;
    const RobotPose2D<double> bestSensorPose {
        sensorPose.mX + this->mLast.ix * this->mLast.step_x,
        sensorPose.mY + this->mLast.iy * this->mLast.step_y,
        sensorPose.mTheta + this->mLast.it * this->mLast.step_t };
    if (this->mDeviceCost) {
        /* The same tail on the device (cost_function_greedy_endpoint.cpp:32-171) */
        const double scanMin = scanData->MinRange(), scanMax = scanData->MaxRange();
        const lgs_scan_batch tailScans { 1, beamBegin, scanData->Angles().data(),
                                         scanData->Ranges().data(), pose, &scanMin, &scanMax };
        std::vector<double> normalizedCosts, covariances;
        this->DeviceTail(tailScans, { bestSensorPose.mX, bestSensorPose.mY, bestSensorPose.mTheta },
                         normalizedCosts, covariances);
        return ScanMatchingSummary {
            poseFound, normalizedCosts[0], initialPose,
            MoveBackward(bestSensorPose, scanData->RelativeSensorPose()),
            ToMatrix(covariances.data()) };
    }
    const GridMapType& gridMap = *hostMap;
    const double costVal = this->mCostFunc->Cost(gridMap, scanData, bestSensorPose);
    const double normalizedCost = costVal / scanData->NumOfScans();
    const RobotPose2D<double> estimatedPose =
        MoveBackward(bestSensorPose, scanData->RelativeSensorPose());
    const Eigen::Matrix3d estimatedCovariance =
        this->mCostFunc->ComputeCovariance(gridMap, scanData, bestSensorPose);

    return ScanMatchingSummary {
        poseFound, normalizedCost, initialPose, estimatedPose, estimatedCovariance };
}

std::vector<ScanMatchingSummary> ScanMatcherRealTimeCorrelativeCuda::OptimizePoses(
    const GridMapType& gridMap,
    const std::vector<Sensor::ScanDataPtr<double>>& scans,
    const std::vector<RobotPose2D<double>>& initialPoses,
    const double normalizedScoreThreshold)
{
    std::vector<ScanMatchingSummary> summaries;
    const int n = static_cast<int>(scans.size());
    if (n == 0)
        return summaries;
    this->UploadMap(gridMap);

    std::vector<int> beamBegin(1, 0);
    std::vector<double> angles, ranges, poses, thresholds(n, normalizedScoreThreshold);
    std::vector<RobotPose2D<double>> sensorPoses;
    for (int k = 0; k < n; ++k) {
        const RobotPose2D<double> sensorPose = Compound(initialPoses[k], scans[k]->RelativeSensorPose());
        sensorPoses.push_back(sensorPose);
        poses.insert(poses.end(), { sensorPose.mX, sensorPose.mY, sensorPose.mTheta });
        angles.insert(angles.end(), scans[k]->Angles().begin(), scans[k]->Angles().end());
        ranges.insert(ranges.end(), scans[k]->Ranges().begin(), scans[k]->Ranges().end());
        beamBegin.push_back(static_cast<int>(angles.size()));
    }
    const lgs_scan_batch batch { n, beamBegin.data(), angles.data(), ranges.data(), poses.data(),
                                 nullptr, nullptr };
    std::vector<lgs_match_result> results(n);
    Check(this->mCtx, lgs_rtcsm_batch_upload(this->mBatch, this->mGrid, &batch, thresholds.data()),
          "lgs_rtcsm_batch_upload");
    Check(this->mCtx, lgs_rtcsm_batch_run(this->mBatch, this->mGrid, this->mCoarse), "lgs_rtcsm_batch_run");
    Check(this->mCtx, lgs_rtcsm_batch_results(this->mBatch, this->mGrid, this->mCoarse, results.data()),
          "lgs_rtcsm_batch_results");

    if (this->mDeviceCost) {
        std::vector<double> best, scanMin, scanMax, normalizedCosts, covariances;
        for (int k = 0; k < n; ++k) {
            const lgs_match_result& r = results[k];
            best.insert(best.end(), { sensorPoses[k].mX + r.ix * r.step_x,
                                      sensorPoses[k].mY + r.iy * r.step_y,
                                      sensorPoses[k].mTheta + r.it * r.step_t });
            scanMin.push_back(scans[k]->MinRange());
            scanMax.push_back(scans[k]->MaxRange());
        }
        const lgs_scan_batch tailScans { n, beamBegin.data(), angles.data(), ranges.data(), poses.data(),
                                         scanMin.data(), scanMax.data() };
        this->DeviceTail(tailScans, best, normalizedCosts, covariances);
        for (int k = 0; k < n; ++k)
            summaries.emplace_back(
                results[k].found != 0, normalizedCosts[k], initialPoses[k],
                MoveBackward(RobotPose2D<double>(best[3 * k], best[3 * k + 1], best[3 * k + 2]),
                             scans[k]->RelativeSensorPose()),
                ToMatrix(covariances.data() + 9 * k));
        this->mLast = results.back();
        return summaries;
    }

    /* per match: the reference's own host tail (scan_matcher_real_time_correlative.cpp:118-144) */
    for (int k = 0; k < n; ++k) {
        const lgs_match_result& r = results[k];
        const RobotPose2D<double> bestSensorPose {
            sensorPoses[k].mX + r.ix * r.step_x, sensorPoses[k].mY + r.iy * r.step_y,
            sensorPoses[k].mTheta + r.it * r.step_t };
        const double costVal = this->mCostFunc->Cost(gridMap, scans[k], bestSensorPose);
        summaries.emplace_back(
            r.found != 0, costVal / scans[k]->NumOfScans(), initialPoses[k],
            MoveBackward(bestSensorPose, scans[k]->RelativeSensorPose()),
            this->mCostFunc->ComputeCovariance(gridMap, scans[k], bestSensorPose));
    }
    this->mLast = results.back();
    return summaries;
}

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */
