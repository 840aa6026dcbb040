// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Flat extern "C" access to the UNMODIFIED reference hot path, compiled from the
// sources where they lie under /root/reference (see oracle/Makefile) into
// oracle/_ref/liblgs_ref.so.  Nothing here restates an algorithm: every result comes
// from the reference's own object code (GridMapBuilder, PrecomputeGridMap(s),
// ScanMatcherRealTimeCorrelative, ScanMatcherBranchBound, ScorePixelAccurate,
// CostGreedyEndpoint, Bresenham, BinaryBayesGridCell).  This translation unit is built
// with -fno-access-control so the private helpers of those classes can be called
// directly (scan_matcher_real_time_correlative.hpp:49-76, grid_map_builder.hpp:158-190).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <vector>

#include "my_lidar_graph_slam/mapping/cost_function_greedy_endpoint.hpp"
#include "my_lidar_graph_slam/mapping/grid_map_builder.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_real_time_correlative.hpp"
#include "my_lidar_graph_slam/mapping/pose_graph.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_grid_search.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_real_time_correlative.hpp"
#include "my_lidar_graph_slam/mapping/score_function_pixel_accurate.hpp"
#include "my_lidar_graph_slam/util.hpp"

using namespace MyLidarGraphSlam;
using namespace MyLidarGraphSlam::Mapping;

namespace {

struct RefMap { GridMapType m; };
struct RefPre { PrecomputedMapType m; };

struct RefBuilder {
    std::shared_ptr<PoseGraph> pg;
    GridMapBuilder b;
    RefBuilder(double res, int patch, int nLatest, double travel, double rmin,
               double rmax, double pHit, double pMiss)
        : pg(std::make_shared<PoseGraph>()),
          b(res, patch, nLatest, travel, rmin, rmax, pHit, pMiss) {}
};

Sensor::ScanDataPtr<double> MakeScan(const double* pose, const double* rel, int n,
                                     const double* angles, const double* ranges,
                                     double scanMinRange, double scanMaxRange) {
    std::vector<double> a(angles, angles + n), r(ranges, ranges + n);
    const double minAngle = n > 0 ? angles[0] : 0.0;
    const double maxAngle = n > 0 ? angles[n - 1] : 0.0;
    return std::make_shared<Sensor::ScanData<double>>(
        "oracle", 0.0, RobotPose2D<double>(pose[0], pose[1], pose[2]),
        RobotPose2D<double>(0.0, 0.0, 0.0),
        RobotPose2D<double>(rel[0], rel[1], rel[2]), scanMinRange, scanMaxRange,
        minAngle, maxAngle, std::move(a), std::move(r));
}

template <typename MapT>
void Dense(const MapT& m, double* out) {
    const int nx = m.NumOfGridCellsX(), ny = m.NumOfGridCellsY();
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x)
            out[static_cast<size_t>(y) * nx + x] = m.Value(x, y, 0.0);
}

std::shared_ptr<CostGreedyEndpoint> MakeCost(const double* c) {
    // c = {usableRangeMin, usableRangeMax, hitAndMissedDist, occupancyThreshold,
    //      kernelSize, arg6, arg7} passed positionally exactly as slam_launcher.cpp:70-72
    // does (the launcher swaps the last two relative to the header's names; kept as is).
    return std::make_shared<CostGreedyEndpoint>(c[0], c[1], c[2], c[3],
                                                static_cast<int>(c[4]), c[5], c[6]);
}

}  // namespace

extern "C" {

// ---- result record shared by both matchers ------------------------------------------
struct RefMatchResult {
    int found;
    int ix, iy, it;           // winning window indices (recovered from the pose)
    int winX, winY, winT;     // window half sizes the reference derived
    int pad;
    double stepX, stepY, stepT;
    double score;             // reference score function evaluated at the winner
    double sensorPose[3];     // Compound(initialPose, relSensorPose)
    double bestSensorPose[3];
    double estPose[3];        // ScanMatchingSummary::mEstimatedPose
    double normalizedCost;    // ScanMatchingSummary::mNormalizedCost
    double cov[9];            // ScanMatchingSummary::mEstimatedCovariance, row-major
};

// ---- primitives ---------------------------------------------------------------------
int ref_bresenham(int x0, int y0, int x1, int y1, int* outXY, int cap) {
    const auto v = Bresenham(Point2D<int>(x0, y0), Point2D<int>(x1, y1));
    const int n = static_cast<int>(v.size());
    for (int i = 0; i < n && i < cap; ++i) {
        outXY[2 * i] = v[i].mX;
        outXY[2 * i + 1] = v[i].mY;
    }
    return n;
}

double ref_bayes_update(double value, double prob) {
    // Drives BinaryBayesGridCell<double>::Update (binary_bayes_grid_cell.hpp:75-92)
    // from a cell currently holding `value` (0.0 = unknown).
    BinaryBayesGridCell<double> cell;
    if (value != 0.0) cell.mValue = value;
    cell.Update(prob);
    return cell.Value();
}

void ref_sliding_window_max(const double* in, int n, int w, double* out) {
    // SlidingWindowMax (util.hpp:199-253); reads past the end return 0.0 like
    // GridMap::Value(x, y, unknown) does for out-of-range cells (grid_map.hpp:859-873).
    std::function<double(int)> inF = [&](int i) { return (i >= 0 && i < n) ? in[i] : 0.0; };
    std::function<void(int, double)> outF = [&](int i, double v) { out[i] = v; };
    SlidingWindowMax(inF, outF, n, w);
}

void ref_compound(const double* a, const double* b, double* o) {
    const auto r = Compound(RobotPose2D<double>(a[0], a[1], a[2]),
                            RobotPose2D<double>(b[0], b[1], b[2]));
    o[0] = r.mX; o[1] = r.mY; o[2] = r.mTheta;
}
void ref_move_backward(const double* a, const double* b, double* o) {
    const auto r = MoveBackward(RobotPose2D<double>(a[0], a[1], a[2]),
                                RobotPose2D<double>(b[0], b[1], b[2]));
    o[0] = r.mX; o[1] = r.mY; o[2] = r.mTheta;
}
void ref_inverse_compound(const double* a, const double* b, double* o) {
    const auto r = InverseCompound(RobotPose2D<double>(a[0], a[1], a[2]),
                                   RobotPose2D<double>(b[0], b[1], b[2]));
    o[0] = r.mX; o[1] = r.mY; o[2] = r.mTheta;
}

// ---- maps ---------------------------------------------------------------------------
RefMap* ref_map_from_dense(double res, int patch, int nx, int ny, double minX,
                           double minY, const double* dense) {
    // nx, ny must be multiples of `patch`.  A non-zero value v is planted through the
    // cell's first-touch path Update(v) -> clamp(v) (binary_bayes_grid_cell.hpp:79-83),
    // so v must lie in [1e-3, 0.999]; 0.0 leaves the cell (and possibly the whole patch)
    // unobserved/unallocated.
    if (nx % patch || ny % patch) return nullptr;
    auto* h = new RefMap{GridMapType(res, patch, nx / patch, ny / patch, minX, minY)};
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            const double v = dense[static_cast<size_t>(y) * nx + x];
            if (v != 0.0) h->m.Update(x, y, v);
        }
    return h;
}
void ref_map_geometry(const RefMap* h, int* nx, int* ny, double* minX, double* minY,
                      double* res) {
    *nx = h->m.NumOfGridCellsX(); *ny = h->m.NumOfGridCellsY();
    *minX = h->m.MinPos().mX; *minY = h->m.MinPos().mY; *res = h->m.Resolution();
}
void ref_map_dense(const RefMap* h, double* out) { Dense(h->m, out); }
void ref_map_destroy(RefMap* h) { delete h; }

RefPre* ref_precompute(const RefMap* h, int win) {
    return new RefPre{PrecomputeGridMap(h->m, win)};            // grid_map_builder.cpp:518
}
int ref_precompute_pyramid(const RefMap* h, int nodeHeightMax, RefPre** out) {
    std::map<int, PrecomputedMapType> pyr;
    PrecomputeGridMaps(h->m, pyr, nodeHeightMax);               // grid_map_builder.cpp:471
    for (int l = 0; l <= nodeHeightMax; ++l) out[l] = new RefPre{std::move(pyr.at(l))};
    return nodeHeightMax + 1;
}
void ref_pre_dense(const RefPre* p, double* out) { Dense(p->m, out); }
void ref_pre_destroy(RefPre* p) { delete p; }

// ---- builder ------------------------------------------------------------------------
RefBuilder* ref_builder_create(double res, int patch, int nLatest, double travelThr,
                               double rmin, double rmax, double pHit, double pMiss) {
    return new RefBuilder(res, patch, nLatest, travelThr, rmin, rmax, pHit, pMiss);
}
int ref_builder_append_scan(RefBuilder* b, const double* pose, const double* rel, int n,
                            const double* angles, const double* ranges,
                            double scanMinRange, double scanMaxRange) {
    b->pg->AppendNode(RobotPose2D<double>(pose[0], pose[1], pose[2]),
                      MakeScan(pose, rel, n, angles, ranges, scanMinRange, scanMaxRange));
    return b->b.AppendScan(b->pg) ? 1 : 0;                      // grid_map_builder.cpp:48
}
// Append a node without touching any map (lets a caller time / drive the two halves of
// AppendScan separately).
void ref_builder_append_node_only(RefBuilder* b, const double* pose, const double* rel,
                                  int n, const double* angles, const double* ranges,
                                  double scanMinRange, double scanMaxRange) {
    b->pg->AppendNode(RobotPose2D<double>(pose[0], pose[1], pose[2]),
                      MakeScan(pose, rel, n, angles, ranges, scanMinRange, scanMaxRange));
}
int ref_builder_update_grid_map(RefBuilder* b) {               // grid_map_builder.cpp:98
    return b->b.UpdateGridMap(b->pg) ? 1 : 0;
}
void ref_builder_update_latest_map(RefBuilder* b) {            // grid_map_builder.cpp:196
    b->b.UpdateLatestMap(b->pg);
}
// ConstructMapFromScans over nodes [lo, hi] into a fresh map (grid_map_builder.cpp:227).
RefMap* ref_builder_construct_map(RefBuilder* b, int lo, int hi) {
    auto* h = new RefMap{GridMapType(b->b.mResolution, b->b.mPatchSize, 0, 0,
                                     Point2D<double>(0.0, 0.0))};
    b->b.ConstructMapFromScans(h->m, b->pg, lo, hi);
    return h;
}
void ref_builder_set_node_pose(RefBuilder* b, int idx, const double* pose) {
    b->pg->NodeAt(idx).Pose() = RobotPose2D<double>(pose[0], pose[1], pose[2]);
}
void ref_builder_after_loop_closure(RefBuilder* b) {           // grid_map_builder.cpp:62
    b->b.AfterLoopClosure(b->pg);
}
int ref_builder_num_local_maps(const RefBuilder* b) {
    return static_cast<int>(b->b.LocalMaps().size());
}
RefMap* ref_builder_local_map(const RefBuilder* b, int i) {
    return new RefMap{b->b.LocalMapAt(i).mMap};
}
void ref_builder_local_map_nodes(const RefBuilder* b, int i, int* lo, int* hi) {
    *lo = b->b.LocalMapAt(i).mPoseGraphNodeIdxMin;
    *hi = b->b.LocalMapAt(i).mPoseGraphNodeIdxMax;
}
RefMap* ref_builder_latest_map(const RefBuilder* b) { return new RefMap{b->b.LatestMap()}; }
void ref_builder_destroy(RefBuilder* b) { delete b; }

// ---- real-time correlative matcher ----------------------------------------------------
// cost[7]: CostGreedyEndpoint ctor arguments in positional order.
int ref_rtcsm_match(const RefMap* map, const RefPre* preOrNull, int lowRes, double rangeX,
                    double rangeY, double rangeTheta, double scanRangeMax,
                    const double* cost, const double* initPose, const double* rel, int n,
                    const double* angles, const double* ranges, double scanMinRange,
                    double scanMaxRange, double normThreshold, RefMatchResult* out) {
    ScanMatcherRealTimeCorrelative rt(MakeCost(cost), lowRes, rangeX, rangeY, rangeTheta,
                                      scanRangeMax);
    const auto scan = MakeScan(initPose, rel, n, angles, ranges, scanMinRange, scanMaxRange);
    const RobotPose2D<double> init(initPose[0], initPose[1], initPose[2]);
    std::unique_ptr<RefPre> own;
    if (!preOrNull) { own.reset(new RefPre{rt.ComputeCoarserMap(map->m)}); preOrNull = own.get(); }

    const ScanMatchingSummary s =
        rt.OptimizePose(map->m, preOrNull->m, scan, init, normThreshold);

    // Everything below only *reads back* what the reference decided.
    const RobotPose2D<double> sensorPose = Compound(init, scan->RelativeSensorPose());
    double sx, sy, st;
    rt.ComputeSearchStep(map->m, scan, sx, sy, st);
    const RobotPose2D<double> best = Compound(s.mEstimatedPose, scan->RelativeSensorPose());
    out->found = s.mPoseFound ? 1 : 0;
    out->winX = static_cast<int>(std::ceil(0.5 * rangeX / sx));
    out->winY = static_cast<int>(std::ceil(0.5 * rangeY / sy));
    out->winT = static_cast<int>(std::ceil(0.5 * rangeTheta / st));
    out->stepX = sx; out->stepY = sy; out->stepT = st;
    out->ix = static_cast<int>(std::lround((best.mX - sensorPose.mX) / sx));
    out->iy = static_cast<int>(std::lround((best.mY - sensorPose.mY) / sy));
    out->it = static_cast<int>(std::lround((best.mTheta - sensorPose.mTheta) / st));
    out->sensorPose[0] = sensorPose.mX; out->sensorPose[1] = sensorPose.mY;
    out->sensorPose[2] = sensorPose.mTheta;
    out->bestSensorPose[0] = sensorPose.mX + out->ix * sx;
    out->bestSensorPose[1] = sensorPose.mY + out->iy * sy;
    out->bestSensorPose[2] = sensorPose.mTheta + out->it * st;
    out->estPose[0] = s.mEstimatedPose.mX; out->estPose[1] = s.mEstimatedPose.mY;
    out->estPose[2] = s.mEstimatedPose.mTheta;
    out->normalizedCost = s.mNormalizedCost;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out->cov[3 * i + j] = s.mEstimatedCovariance(i, j);
    // Score of the winner through the reference's own projection + gather.
    std::vector<Point2D<int>> idx;
    rt.ComputeScanIndices(preOrNull->m,
        RobotPose2D<double>(sensorPose.mX, sensorPose.mY, sensorPose.mTheta + st * out->it),
        scan, idx);
    out->score = rt.ComputeScore(map->m, idx, out->ix, out->iy);
    return 0;
}

// Exhaustive score table through the reference's ComputeScanIndices / ComputeScore:
// table[(t+winT) * nyw * nxw + (y - yLo) * nxw + (x - xLo)] for x in [xLo, xLo+nxw),
// y in [yLo, yLo+nyw); `useCoarse` selects the win-max map (must be given) instead of
// the fine grid.  Also returns, per theta, the projected indices (idxOut: [nT][n][2],
// entries beyond the kept count are left untouched) and kept counts (cntOut[nT]).
int ref_rtcsm_score_table(const RefMap* map, const RefPre* pre, int useCoarse, int lowRes,
                          double scanRangeMax, const double* sensorPose, int n,
                          const double* angles, const double* ranges, double stepT,
                          int winT, int xLo, int nxw, int yLo, int nyw, double* table,
                          int* idxOut, int* cntOut) {
    const double cost[7] = {0.01, 20.0, 0.075, 0.1, 1, 0.05, 1.0};
    ScanMatcherRealTimeCorrelative rt(MakeCost(cost), lowRes, 1.0, 1.0, 1.0, scanRangeMax);
    const double zero[3] = {0, 0, 0};
    const auto scan = MakeScan(sensorPose, zero, n, angles, ranges, 0.0, 1e9);
    std::vector<Point2D<int>> idx;
    for (int t = -winT; t <= winT; ++t) {
        rt.ComputeScanIndices(pre->m,
            RobotPose2D<double>(sensorPose[0], sensorPose[1], sensorPose[2] + stepT * t),
            scan, idx);
        const size_t tt = static_cast<size_t>(t + winT);
        if (cntOut) cntOut[tt] = static_cast<int>(idx.size());
        if (idxOut)
            for (size_t i = 0; i < idx.size(); ++i) {
                idxOut[(tt * n + i) * 2] = idx[i].mX;
                idxOut[(tt * n + i) * 2 + 1] = idx[i].mY;
            }
        if (!table) continue;
        for (int y = 0; y < nyw; ++y)
            for (int x = 0; x < nxw; ++x) {
                const double s = useCoarse
                    ? rt.ComputeScore(pre->m, idx, xLo + x, yLo + y)
                    : rt.ComputeScore(map->m, idx, xLo + x, yLo + y);
                table[(tt * nyw + y) * nxw + x] = s;
            }
    }
    return 0;
}

// ---- branch-and-bound matcher ---------------------------------------------------------
int ref_bb_match(const RefMap* map, RefPre* const* pyr, int nodeHeightMax, double rangeX,
                 double rangeY, double rangeTheta, double scanRangeMax,
                 double scoreRangeMin, double scoreRangeMax, const double* cost,
                 const double* initPose, const double* rel, int n, const double* angles,
                 const double* ranges, double scanMinRange, double scanMaxRange,
                 double normThreshold, RefMatchResult* out) {
    auto sf = std::make_shared<ScorePixelAccurate>(scoreRangeMin, scoreRangeMax);
    ScanMatcherBranchBound bb(sf, MakeCost(cost), nodeHeightMax, rangeX, rangeY,
                              rangeTheta, scanRangeMax);
    const auto scan = MakeScan(initPose, rel, n, angles, ranges, scanMinRange, scanMaxRange);
    const RobotPose2D<double> init(initPose[0], initPose[1], initPose[2]);
    std::map<int, PrecomputedMapType> maps;
    if (pyr) {
        for (int l = 0; l <= nodeHeightMax; ++l) maps.emplace(l, pyr[l]->m);
    } else {
        maps = bb.ComputeCoarserMaps(map->m);
    }
    const ScanMatchingSummary s = bb.OptimizePose(map->m, maps, scan, init, normThreshold);

    const RobotPose2D<double> sensorPose = Compound(init, scan->RelativeSensorPose());
    double sx, sy, st;
    bb.ComputeSearchStep(map->m, scan, sx, sy, st);
    const RobotPose2D<double> best = Compound(s.mEstimatedPose, scan->RelativeSensorPose());
    out->found = s.mPoseFound ? 1 : 0;
    out->winX = static_cast<int>(std::ceil(0.5 * rangeX / sx));
    out->winY = static_cast<int>(std::ceil(0.5 * rangeY / sy));
    out->winT = static_cast<int>(std::ceil(0.5 * rangeTheta / st));
    out->stepX = sx; out->stepY = sy; out->stepT = st;
    out->ix = static_cast<int>(std::lround((best.mX - sensorPose.mX) / sx));
    out->iy = static_cast<int>(std::lround((best.mY - sensorPose.mY) / sy));
    out->it = static_cast<int>(std::lround((best.mTheta - sensorPose.mTheta) / st));
    out->sensorPose[0] = sensorPose.mX; out->sensorPose[1] = sensorPose.mY;
    out->sensorPose[2] = sensorPose.mTheta;
    if (s.mPoseFound) {
        out->bestSensorPose[0] = sensorPose.mX + out->ix * sx;
        out->bestSensorPose[1] = sensorPose.mY + out->iy * sy;
        out->bestSensorPose[2] = sensorPose.mTheta + out->it * st;
    } else {  // scan_matcher_branch_bound.cpp:80: bestSensorPose starts as sensorPose
        out->bestSensorPose[0] = sensorPose.mX; out->bestSensorPose[1] = sensorPose.mY;
        out->bestSensorPose[2] = sensorPose.mTheta;
    }
    out->estPose[0] = s.mEstimatedPose.mX; out->estPose[1] = s.mEstimatedPose.mY;
    out->estPose[2] = s.mEstimatedPose.mTheta;
    out->normalizedCost = s.mNormalizedCost;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out->cov[3 * i + j] = s.mEstimatedCovariance(i, j);
    ScoreFunction::Summary sum;
    sf->Score(maps.at(0), scan,
              RobotPose2D<double>(out->bestSensorPose[0], out->bestSensorPose[1],
                                  out->bestSensorPose[2]), sum);
    out->score = sum.mScore;
    return 0;
}

// ScanMatcherGridSearch::OptimizePose (scan_matcher_grid_search.cpp:45-114).  ix / iy / it are the
// loop counters of the winning dx / dy / dt (recovered by replaying the reference's own accumulating
// loops, :74-76); winX / winY / winT receive the numbers of steps of the three loops.
int ref_gs_match(const RefMap* map, double rangeX, double rangeY, double rangeTheta, double stepX,
                 double stepY, double stepTheta, double scoreRangeMin, double scoreRangeMax,
                 const double* cost, const double* initPose, const double* rel, int n,
                 const double* angles, const double* ranges, double scanMinRange, double scanMaxRange,
                 double normThreshold, RefMatchResult* out) {
    auto sf = std::make_shared<ScorePixelAccurate>(scoreRangeMin, scoreRangeMax);
    ScanMatcherGridSearch gs(sf, MakeCost(cost), rangeX, rangeY, rangeTheta, stepX, stepY, stepTheta);
    const auto scan = MakeScan(initPose, rel, n, angles, ranges, scanMinRange, scanMaxRange);
    const RobotPose2D<double> init(initPose[0], initPose[1], initPose[2]);
    const ScanMatchingSummary s = gs.OptimizePose(map->m, scan, init, normThreshold);

    const RobotPose2D<double> sensorPose = Compound(init, scan->RelativeSensorPose());
    const RobotPose2D<double> best = Compound(s.mEstimatedPose, scan->RelativeSensorPose());
    std::memset(out, 0, sizeof(*out));
    out->found = s.mPoseFound ? 1 : 0;
    out->stepX = stepX; out->stepY = stepY; out->stepT = stepTheta;
    out->sensorPose[0] = sensorPose.mX; out->sensorPose[1] = sensorPose.mY;
    out->sensorPose[2] = sensorPose.mTheta;
    out->ix = out->iy = out->it = -1;
    // the winner as the matcher itself computed it: MoveBackward then Compound is not an exact
    // round trip, so search the loop values for the pose whose MoveBackward gives the summary's
    const double rx = rangeX / 2.0, ry = rangeY / 2.0, rt = rangeTheta / 2.0;
    int ny = 0;
    for (double dy = -ry; dy <= ry; dy += stepY, ++ny) {
        int nx = 0;
        for (double dx = -rx; dx <= rx; dx += stepX, ++nx) {
            int nt = 0;
            for (double dt = -rt; dt <= rt; dt += stepTheta, ++nt) {
                if (!s.mPoseFound || out->ix >= 0) continue;
                const RobotPose2D<double> pose { sensorPose.mX + dx, sensorPose.mY + dy, sensorPose.mTheta + dt };
                const RobotPose2D<double> est = MoveBackward(pose, scan->RelativeSensorPose());
                if (est.mX == s.mEstimatedPose.mX && est.mY == s.mEstimatedPose.mY &&
                    est.mTheta == s.mEstimatedPose.mTheta) {
                    out->ix = nx; out->iy = ny; out->it = nt;
                    out->bestSensorPose[0] = pose.mX; out->bestSensorPose[1] = pose.mY;
                    out->bestSensorPose[2] = pose.mTheta;
                }
            }
            out->winT = nt;
        }
        out->winX = nx;
    }
    out->winY = ny;
    if (!s.mPoseFound) {   // :71: bestSensorPose starts as sensorPose
        out->bestSensorPose[0] = sensorPose.mX; out->bestSensorPose[1] = sensorPose.mY;
        out->bestSensorPose[2] = sensorPose.mTheta;
    }
    (void)best;
    out->estPose[0] = s.mEstimatedPose.mX; out->estPose[1] = s.mEstimatedPose.mY;
    out->estPose[2] = s.mEstimatedPose.mTheta;
    out->normalizedCost = s.mNormalizedCost;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) out->cov[3 * i + j] = s.mEstimatedCovariance(i, j);
    ScoreFunction::Summary sum;
    sf->Score(map->m, scan, RobotPose2D<double>(out->bestSensorPose[0], out->bestSensorPose[1],
                                                out->bestSensorPose[2]), sum);
    out->score = sum.mScore;
    return 0;
}

// ScorePixelAccurate::Score on one pyramid level (score_function_pixel_accurate.cpp:19).
double ref_pixel_accurate_score(const RefPre* level, double scoreRangeMin,
                                double scoreRangeMax, const double* sensorPose, int n,
                                const double* angles, const double* ranges,
                                double scanMinRange, double scanMaxRange) {
    ScorePixelAccurate sf(scoreRangeMin, scoreRangeMax);
    const double zero[3] = {0, 0, 0};
    const auto scan = MakeScan(sensorPose, zero, n, angles, ranges, scanMinRange, scanMaxRange);
    ScoreFunction::Summary sum;
    sf.Score(level->m, scan,
             RobotPose2D<double>(sensorPose[0], sensorPose[1], sensorPose[2]), sum);
    return sum.mScore;
}

// Host tail exactly as both matchers run it (scan_matcher_real_time_correlative.cpp:128-138).
void ref_host_tail(const RefMap* map, const double* cost, const double* bestSensorPose,
                   const double* rel, int n, const double* angles, const double* ranges,
                   double scanMinRange, double scanMaxRange, double* normalizedCost,
                   double* estPose, double* cov) {
    auto cf = MakeCost(cost);
    const auto scan = MakeScan(bestSensorPose, rel, n, angles, ranges, scanMinRange, scanMaxRange);
    const RobotPose2D<double> best(bestSensorPose[0], bestSensorPose[1], bestSensorPose[2]);
    *normalizedCost = cf->Cost(map->m, scan, best) / scan->NumOfScans();
    const auto est = MoveBackward(best, scan->RelativeSensorPose());
    estPose[0] = est.mX; estPose[1] = est.mY; estPose[2] = est.mTheta;
    const Eigen::Matrix3d c = cf->ComputeCovariance(map->m, scan, best);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov[3 * i + j] = c(i, j);
}

// CostGreedyEndpoint::Cost at an arbitrary sensor pose (cost_function_greedy_endpoint.cpp:32-110).
double ref_cost_greedy_endpoint(const RefMap* map, const double* cost, const double* sensorPose, int n,
                                const double* angles, const double* ranges, double scanMinRange,
                                double scanMaxRange) {
    auto cf = MakeCost(cost);
    const double zero[3] = {0, 0, 0};
    const auto scan = MakeScan(sensorPose, zero, n, angles, ranges, scanMinRange, scanMaxRange);
    return cf->Cost(map->m, scan, RobotPose2D<double>(sensorPose[0], sensorPose[1], sensorPose[2]));
}

// Range filter + HitPoint + bounding box exactly as GridMapBuilder does it
// (ComputeBoundingBoxAndScanPoints, grid_map_builder.cpp:335-380).  usable = {min, max}.
int ref_hit_points(const double* robotPose, const double* rel, int n, const double* angles,
                   const double* ranges, double scanMinRange, double scanMaxRange,
                   double usableMin, double usableMax, double* sensorPose, double* hitXY,
                   double* bbox) {
    GridMapBuilder b(0.05, 64, 10, 20.0, usableMin, usableMax, 0.6, 0.45);
    const auto scan = MakeScan(robotPose, rel, n, angles, ranges, scanMinRange, scanMaxRange);
    const RobotPose2D<double> rp(robotPose[0], robotPose[1], robotPose[2]);
    Point2D<double> bl, tr;
    std::vector<Point2D<double>> hits;
    b.ComputeBoundingBoxAndScanPoints(rp, scan, bl, tr, hits);
    const auto sp = Compound(rp, scan->RelativeSensorPose());
    sensorPose[0] = sp.mX; sensorPose[1] = sp.mY; sensorPose[2] = sp.mTheta;
    for (size_t i = 0; i < hits.size(); ++i) { hitXY[2 * i] = hits[i].mX; hitXY[2 * i + 1] = hits[i].mY; }
    bbox[0] = bl.mX; bbox[1] = bl.mY; bbox[2] = tr.mX; bbox[3] = tr.mY;
    return static_cast<int>(hits.size());
}

// The per-scan integration loop of UpdateGridMap (grid_map_builder.cpp:159-186) applied to an
// existing map with caller-supplied sensor position and hit points (every touched cell must be
// inside the map): WorldCoordinateToGridCellIndex, ComputeMissedGridCellIndices (Bresenham
// minus the last cell), Update(miss)..., Update(hit) -- all the reference's own functions.
int ref_map_integrate_hits(RefMap* h, const double* sensorXY, int n, const double* hitXY,
                           double pHit, double pMiss) {
    GridMapBuilder b(h->m.Resolution(), h->m.PatchSize(), 10, 20.0, 0.01, 20.0, pHit, pMiss);
    const Point2D<int> s = h->m.WorldCoordinateToGridCellIndex(sensorXY[0], sensorXY[1]);
    int updates = 0;
    for (int i = 0; i < n; ++i) {
        const Point2D<int> e = h->m.WorldCoordinateToGridCellIndex(hitXY[2 * i], hitXY[2 * i + 1]);
        const std::vector<Point2D<int>> missed = b.ComputeMissedGridCellIndices(s, e);
        for (const auto& c : missed) {
            if (!h->m.IsInside(c)) return -1;
            h->m.Update(c, pMiss);
        }
        if (!h->m.IsInside(e)) return -1;
        h->m.Update(e, pHit);
        updates += static_cast<int>(missed.size()) + 1;
    }
    return updates;
}

// GridMap::Resize / Expand / Reset on an oracle map (grid_map.hpp:652-736, :740-752).
void ref_map_resize(RefMap* h, double minX, double minY, double maxX, double maxY) {
    h->m.Resize(minX, minY, maxX, maxY);
}
void ref_map_expand(RefMap* h, double minX, double minY, double maxX, double maxY, double step) {
    h->m.Expand(minX, minY, maxX, maxY, step);
}
void ref_map_reset(RefMap* h) { h->m.Reset(); }
RefMap* ref_map_create_empty(double res, int patch, double centerX, double centerY) {
    return new RefMap{GridMapType(res, patch, 0, 0, Point2D<double>(centerX, centerY))};
}

const char* ref_version(void) { return "my-lidar-graph-slam reference objects (unmodified)"; }

}  // extern "C"
