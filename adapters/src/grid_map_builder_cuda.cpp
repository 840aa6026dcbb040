/* grid_map_builder_cuda.cpp -- see grid_map_builder_cuda.hpp */
#include "lgs_adapters/grid_map_builder_cuda.hpp"
#include "lgs_adapters/grid_map_flatten.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

namespace MyLidarGraphSlam {
namespace Mapping {

namespace {

void Check(lgs_ctx* ctx, int rc, const char* what)
{
    if (rc == LGS_OK)
        return;
    /* like Assert() in util.hpp:29-38: the reference aborts on broken invariants */
    std::fprintf(stderr, "GridMapBuilderCuda: %s failed (%d): %s\n", what, rc,
                 ctx != nullptr ? lgs_ctx_last_error(ctx) : "no context");
    std::abort();
}

inline double MsSince(const std::chrono::steady_clock::time_point& t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

/* Write the device value into one host cell through the cell's public API (see the header) */
inline void StoreCell(BinaryBayesGridCell<double>& cell, const double value)
{
    if (cell.Value() == value)
        return;
    cell.Reset();
    if (value != 0.0)
        cell.Update(value);
}

} /* namespace */

GridMapBuilderCuda::GridMapBuilderCuda(
    double mapResolution, int patchSize, int numOfScansForLatestMap, double travelDistThreshold,
    double usableRangeMin, double usableRangeMax, double probHit, double probMiss, int device) :
    mResolution(mapResolution),
    mPatchSize(patchSize),
    mLatestMap(mapResolution, patchSize, 0, 0, Point2D<double>(0.0, 0.0)),
    mAccumTravelDist(0.0),
    mNumOfScansForLatestMap(numOfScansForLatestMap),
    mLatestScanIdxMin(0),
    mLatestScanIdxMax(0),
    mLastRobotPose(0.0, 0.0, 0.0),
    mTravelDistLastLocalMap(0.0),
    mRobotPoseLastLocalMap(0.0, 0.0, 0.0),
    mTravelDistThreshold(travelDistThreshold),
    mUsableRangeMin(usableRangeMin),
    mUsableRangeMax(usableRangeMax),
    mProbHit(probHit),
    mProbMiss(probMiss),
    mCtx(nullptr),
    mDevLocal(nullptr),
    mDevScratch(nullptr),
    mDensePinned(false),
    mLazy(false),
    mLocalPending(false),
    mLocalBox { 0, 0, -1, -1 },
    mLatestPending(false),
    mScratchIsLatest(false),
    mNumOfUpdates(0),
    mTimingsMs { 0.0, 0.0, 0.0, 0.0 }
{
    Check(nullptr, lgs_ctx_create(device, &this->mCtx), "lgs_ctx_create (no usable B200: no CPU fallback)");
}

GridMapBuilderCuda::~GridMapBuilderCuda()
{
    if (this->mDensePinned)
        lgs_host_unpin(this->mCtx, this->mDense.data());
    lgs_grid_destroy(this->mDevLocal);
    lgs_grid_destroy(this->mDevScratch);
    lgs_ctx_destroy(this->mCtx);
}

bool GridMapBuilderCuda::AppendScan(const std::shared_ptr<PoseGraph>& poseGraph)
{
    const bool localMapCreated = this->UpdateGridMap(poseGraph);
    this->UpdateLatestMap(poseGraph);
    return localMapCreated;
}

void GridMapBuilderCuda::AfterLoopClosure(const std::shared_ptr<PoseGraph>& poseGraph)
{
    /* every local map is rebuilt from the corrected poses (grid_map_builder.cpp:62-80); stale host cells of
     * the current local map are about to be recomputed anyway */
    this->mLocalPending = false;
    this->mLatestPending = false;
    const bool lazy = this->mLazy;
    this->mLazy = false;                        /* the rebuilds share one scratch grid: synchronise each at once */
    for (auto& info : this->mLocalMaps) {
        this->ConstructMapFromScans(info.mMap, poseGraph,
                                    info.mPoseGraphNodeIdxMin, info.mPoseGraphNodeIdxMax);
        info.mPrecomputedMaps.clear();
        info.mPrecomputed = false;
    }
    /* the mirror of the current local map follows its rebuilt host map */
    if (!this->mLocalMaps.empty()) {
        GridMapType& current = this->mLocalMaps.back().mMap;
        this->MirrorGeometry(this->mDevLocal, current);
        if (current.NumOfGridCellsX() > 0 && current.NumOfGridCellsY() > 0) {
            std::vector<double> flat;
            LgsB200::FlattenGridMap(current, flat);
            this->ReserveDense(flat.size());
            std::copy(flat.begin(), flat.end(), this->mDense.begin());
            Check(this->mCtx, lgs_grid_upload(this->mDevLocal, this->mDense.data()), "lgs_grid_upload");
        }
    }
    this->mLazy = lazy;
    this->UpdateLatestMap(poseGraph);

    /* accumulated travel distance from the corrected poses (:210-224) */
    this->mAccumTravelDist = 0.0;
    const int numOfNodes = static_cast<int>(poseGraph->Nodes().size());
    for (int i = 1; i < numOfNodes; ++i)
        this->mAccumTravelDist += Distance(poseGraph->NodeAt(i - 1).Pose(), poseGraph->NodeAt(i).Pose());
}

GridMapType GridMapBuilderCuda::ConstructGlobalMap(const std::shared_ptr<PoseGraph>& poseGraph)
{
    GridMapType gridMap { this->mResolution, this->mPatchSize, 0, 0, Point2D<double>(0.0, 0.0) };
    this->FlushHostMaps();                      /* the scratch grid is about to hold the global map */
    const bool lazy = this->mLazy;
    this->mLazy = false;
    this->ConstructMapFromScans(gridMap, poseGraph, poseGraph->Nodes().front().Index(),
                                poseGraph->Nodes().back().Index());
    this->mLazy = lazy;
    return gridMap;
}

void GridMapBuilderCuda::AppendNodeHits(const PoseGraph::Node& node, double* bbox)
{
    const auto& scanData = node.ScanData();
    const RobotPose2D<double> sensorPose = Compound(node.Pose(), scanData->RelativeSensorPose());
    /* usable range window of this scan (grid_map_builder.cpp:359-362) */
    const double minRange = std::max(this->mUsableRangeMin, scanData->MinRange());
    const double maxRange = std::min(this->mUsableRangeMax, scanData->MaxRange());
    const int numOfScans = static_cast<int>(scanData->NumOfScans());
    const double pose[3] = { sensorPose.mX, sensorPose.mY, sensorPose.mTheta };
    const std::size_t first = this->mHitXY.size();
    this->mHitXY.resize(first + 2 * static_cast<std::size_t>(numOfScans));
    int numOfHits = 0;
    double scanBox[4];
    /* range filter + ScanData::HitPoint + bounding box (sensor position included), host IEEE */
    Check(this->mCtx, lgs_scan_hit_points(pose, numOfScans, scanData->Angles().data(),
          scanData->Ranges().data(), minRange, maxRange, this->mHitXY.data() + first, &numOfHits,
          scanBox), "lgs_scan_hit_points");
    this->mHitXY.resize(first + 2 * static_cast<std::size_t>(numOfHits));
    this->mSensorXY.push_back(sensorPose.mX);
    this->mSensorXY.push_back(sensorPose.mY);
    this->mHitBegin.push_back(static_cast<int>(this->mHitXY.size() / 2));
    bbox[0] = std::min(bbox[0], scanBox[0]); bbox[1] = std::min(bbox[1], scanBox[1]);
    bbox[2] = std::max(bbox[2], scanBox[2]); bbox[3] = std::max(bbox[3], scanBox[3]);
}

void GridMapBuilderCuda::MirrorGeometry(lgs_grid*& grid, const GridMapType& map)
{
    const int nx = map.NumOfGridCellsX(), ny = map.NumOfGridCellsY();
    if (grid != nullptr) {
        /* same allocation, new placement, all cells unknown */
        Check(this->mCtx, lgs_grid_resize(grid, nx, ny, map.MinPos().mX, map.MinPos().mY,
              1 << 30, 1 << 30), "lgs_grid_resize");
        return;
    }
    Check(this->mCtx, lgs_grid_create(this->mCtx, nx, ny, map.MinPos().mX, map.MinPos().mY,
          map.Resolution(), 1, &grid), "lgs_grid_create");
}

void GridMapBuilderCuda::ReserveDense(std::size_t cells) const
{
    /* page-locked staging for the map downloads; re-registered only when it has to grow */
    if (cells > this->mDense.capacity()) {
        if (this->mDensePinned)
            lgs_host_unpin(this->mCtx, this->mDense.data());
        this->mDense.clear();
        this->mDense.reserve(cells + cells / 2);
        this->mDensePinned = lgs_host_pin(this->mCtx, this->mDense.data(),
                                          this->mDense.capacity() * sizeof(double)) == LGS_OK;
    }
    this->mDense.resize(cells);
}

void GridMapBuilderCuda::Integrate(lgs_grid* grid)
{
    const lgs_hit_batch batch { static_cast<int>(this->mSensorXY.size() / 2), this->mSensorXY.data(),
                                this->mHitBegin.data(), this->mHitXY.data() };
    long long updates = 0;
    const auto t0 = std::chrono::steady_clock::now();
    Check(this->mCtx, lgs_grid_integrate_scans(this->mCtx, grid, &batch, this->mProbHit,
          this->mProbMiss, &updates), "lgs_grid_integrate_scans");
    this->mNumOfUpdates += updates;
    this->mTimingsMs[1] += MsSince(t0);
}

void GridMapBuilderCuda::FlushHostMaps() const
{
    if (this->mLocalPending) {
        this->mLocalPending = false;
        GridMapType& map = const_cast<GridMapType&>(this->mLocalMaps.back().mMap);
        this->SyncRegion(this->mDevLocal, map, this->mLocalBox[0], this->mLocalBox[1],
                         this->mLocalBox[2], this->mLocalBox[3]);
    }
    if (this->mLatestPending) {
        this->mLatestPending = false;
        GridMapType& map = const_cast<GridMapType&>(this->mLatestMap);
        this->SyncRegion(this->mDevScratch, map, 0, 0, map.NumOfGridCellsX() - 1, map.NumOfGridCellsY() - 1);
    }
}

void GridMapBuilderCuda::SyncRegion(const lgs_grid* grid, GridMapType& map,
                                    int x0, int y0, int x1, int y1) const
{
    const int nx = map.NumOfGridCellsX(), ny = map.NumOfGridCellsY();
    if (nx == 0 || ny == 0)
        return;
    auto t0 = std::chrono::steady_clock::now();
    x0 = std::max(x0, 0); y0 = std::max(y0, 0);
    x1 = std::min(x1, nx - 1); y1 = std::min(y1, ny - 1);
    if (x1 < x0 || y1 < y0)
        return;
    /* only the region that can have changed comes back, into its place in the full-map staging */
    this->ReserveDense(static_cast<std::size_t>(nx) * ny);
    Check(this->mCtx, lgs_grid_download_region(grid, x0, y0, x1 - x0 + 1, y1 - y0 + 1,
          this->mDense.data() + static_cast<std::size_t>(y0) * nx + x0, nx), "lgs_grid_download_region");
    this->mTimingsMs[2] += MsSince(t0);
    t0 = std::chrono::steady_clock::now();
    const int patch = this->mPatchSize;
    for (int py = y0 / patch; py <= y1 / patch; ++py)
        for (int px = x0 / patch; px <= x1 / patch; ++px) {
            const int cx0 = std::max(x0, px * patch), cx1 = std::min(x1, px * patch + patch - 1);
            const int cy0 = std::max(y0, py * patch), cy1 = std::min(y1, py * patch + patch - 1);
            if (!map.PatchIsAllocated(px, py)) {
                /* the first observed cell allocates the patch exactly like the CPU's first Update */
                bool allocated = false;
                for (int y = cy0; y <= cy1 && !allocated; ++y)
                    for (int x = cx0; x <= cx1 && !allocated; ++x) {
                        const double v = this->mDense[static_cast<std::size_t>(y) * nx + x];
                        if (v != 0.0) { map.Update(x, y, v); allocated = true; }
                    }
                if (!allocated)
                    continue;
            }
            auto* cells = map.PatchAt(px, py).Data();     /* row-major y * size + x */
            for (int y = cy0; y <= cy1; ++y)
                for (int x = cx0; x <= cx1; ++x)
                    StoreCell(cells[(y - py * patch) * patch + (x - px * patch)],
                              this->mDense[static_cast<std::size_t>(y) * nx + x]);
        }
    this->mTimingsMs[3] += MsSince(t0);
}

bool GridMapBuilderCuda::UpdateGridMap(const std::shared_ptr<PoseGraph>& poseGraph)
{
    const PoseGraph::Node& node = poseGraph->LatestNode();
    const RobotPose2D<double>& robotPose = node.Pose();
    const int nodeIdx = node.Index();

    /* travel distance bookkeeping (grid_map_builder.cpp:108-126) */
    const bool isFirstScan = this->mLocalMaps.empty();
    const double moved = isFirstScan ? 0.0 : Distance(InverseCompound(this->mLastRobotPose, robotPose));
    this->mLastRobotPose = robotPose;
    this->mAccumTravelDist += moved;
    this->mTravelDistLastLocalMap += moved;
    const bool createNewLocalMap =
        isFirstScan || this->mTravelDistLastLocalMap >= this->mTravelDistThreshold;

    if (createNewLocalMap) {
        this->FlushHostMaps();                  /* the device mirror is about to follow the NEW local map */
        if (!isFirstScan)
            this->mLocalMaps.back().mFinished = true;
        GridMapType newLocalMap { this->mResolution, this->mPatchSize, 0, 0,
                                  Point2D<double>(robotPose.mX, robotPose.mY) };
        this->mLocalMaps.emplace_back(static_cast<int>(this->mLocalMaps.size()),
                                      std::move(newLocalMap), nodeIdx);
        this->mTravelDistLastLocalMap = 0.0;
        this->mRobotPoseLastLocalMap = robotPose;
    }

    LocalMapInfo& info = this->mLocalMaps.back();
    GridMapType& localMap = info.mMap;

    /* scan points + bounding box, then grow the host map like the reference does (:152-160) */
    this->mSensorXY.clear(); this->mHitXY.clear(); this->mHitBegin.assign(1, 0);
    double bbox[4] = { std::numeric_limits<double>::max(), std::numeric_limits<double>::max(),
                       std::numeric_limits<double>::lowest(), std::numeric_limits<double>::lowest() };
    this->AppendNodeHits(node, bbox);
    const Point2D<double> oldMin = localMap.MinPos();
    const int oldNx = localMap.NumOfGridCellsX(), oldNy = localMap.NumOfGridCellsY();
    localMap.Expand(bbox[0], bbox[1], bbox[2], bbox[3]);

    /* the device mirror follows: new map -> fresh grid; grown map -> shifted copy */
    const int nx = localMap.NumOfGridCellsX(), ny = localMap.NumOfGridCellsY();
    if (createNewLocalMap || this->mDevLocal == nullptr) {
        this->MirrorGeometry(this->mDevLocal, localMap);
    } else if (nx != oldNx || ny != oldNy || localMap.MinPos().mX != oldMin.mX ||
               localMap.MinPos().mY != oldMin.mY) {
        /* the map grows by whole patches: the offset of the old origin is an exact cell count */
        const int shiftX = static_cast<int>(std::lround((localMap.MinPos().mX - oldMin.mX) / this->mResolution));
        const int shiftY = static_cast<int>(std::lround((localMap.MinPos().mY - oldMin.mY) / this->mResolution));
        Check(this->mCtx, lgs_grid_resize(this->mDevLocal, nx, ny, localMap.MinPos().mX,
              localMap.MinPos().mY, shiftX, shiftY), "lgs_grid_resize");
        /* new cell (x, y) holds old cell (x + shift, y + shift): stale host cells move with the map */
        this->mLocalBox[0] -= shiftX; this->mLocalBox[2] -= shiftX;
        this->mLocalBox[1] -= shiftY; this->mLocalBox[3] -= shiftY;
    }

    /* one scan of integration on the device; only the scan's bounding box can have changed */
    const Point2D<int> c0 = localMap.WorldCoordinateToGridCellIndex(bbox[0], bbox[1]);
    const Point2D<int> c1 = localMap.WorldCoordinateToGridCellIndex(bbox[2], bbox[3]);
    this->Integrate(this->mDevLocal);
    if (this->mLazy) {
        if (!this->mLocalPending) {
            this->mLocalBox[0] = c0.mX; this->mLocalBox[1] = c0.mY; this->mLocalBox[2] = c1.mX; this->mLocalBox[3] = c1.mY;
        } else {
            this->mLocalBox[0] = std::min(this->mLocalBox[0], c0.mX); this->mLocalBox[1] = std::min(this->mLocalBox[1], c0.mY);
            this->mLocalBox[2] = std::max(this->mLocalBox[2], c1.mX); this->mLocalBox[3] = std::max(this->mLocalBox[3], c1.mY);
        }
        this->mLocalPending = true;
    } else {
        this->SyncRegion(this->mDevLocal, localMap, c0.mX, c0.mY, c1.mX, c1.mY);
    }

    info.mPoseGraphNodeIdxMax = nodeIdx;
    return createNewLocalMap;
}

void GridMapBuilderCuda::UpdateLatestMap(const std::shared_ptr<PoseGraph>& poseGraph)
{
    this->mLatestScanIdxMin = std::max(
        0, poseGraph->LatestNode().Index() - this->mNumOfScansForLatestMap + 1);
    this->mLatestScanIdxMax = poseGraph->LatestNode().Index();
    this->ConstructMapFromScans(this->mLatestMap, poseGraph,
                                this->mLatestScanIdxMin, this->mLatestScanIdxMax);
    this->mScratchIsLatest = true;
}

void GridMapBuilderCuda::ConstructMapFromScans(
    GridMapType& gridMap, const std::shared_ptr<PoseGraph>& poseGraph, int nodeIdxMin, int nodeIdxMax)
{
    /* hit points of all nodes and their bounding box; the reference seeds the upper bound with
     * numeric_limits<double>::min(), the smallest POSITIVE double (grid_map_builder.cpp:236-237) */
    this->mSensorXY.clear(); this->mHitXY.clear(); this->mHitBegin.assign(1, 0);
    double bbox[4] = { std::numeric_limits<double>::max(), std::numeric_limits<double>::max(),
                       std::numeric_limits<double>::min(), std::numeric_limits<double>::min() };
    for (int nodeIdx = nodeIdxMin; nodeIdx <= nodeIdxMax; ++nodeIdx)
        this->AppendNodeHits(poseGraph->NodeAt(nodeIdx), bbox);

    /* geometry and the all-unknown state from the reference's own methods (:279-282) */
    gridMap.Resize(bbox[0], bbox[1], bbox[2], bbox[3]);
    gridMap.Reset();

    /* all scans in one device batch, applied in node order like the CPU loop (:285-329) */
    this->mScratchIsLatest = false;
    this->mLatestPending = false;               /* whatever the scratch grid held is being replaced */
    this->MirrorGeometry(this->mDevScratch, gridMap);
    this->Integrate(this->mDevScratch);
    if (this->mLazy && &gridMap == &this->mLatestMap)
        this->mLatestPending = true;            /* the host latest map follows when somebody reads it */
    else
        this->SyncRegion(this->mDevScratch, gridMap, 0, 0,
                         gridMap.NumOfGridCellsX() - 1, gridMap.NumOfGridCellsY() - 1);
}

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */
