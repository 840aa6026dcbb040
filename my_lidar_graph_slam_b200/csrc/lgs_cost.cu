// lgs_cost.cu -- the matchers' host tail on the device: CostGreedyEndpoint::Cost and the finite
// difference gradient / covariance built from it.
//
// Reference: cost_function_greedy_endpoint.cpp:32-110 (Cost), :113-145 (ComputeGradient),
// :148-171 (ComputeCovariance); ScanData::HitAndMissedPoint sensor_data.hpp:177-198;
// GridMap::SquaredDistance grid_map.hpp:895-902; the call site is the tail of both matchers
// (scan_matcher_real_time_correlative.cpp:126-138, scan_matcher_branch_bound.cpp:143-162).
// On the CPU one tail is 7 Cost() evaluations = 7 x beams x (sincos + 18 virtual cell reads + exp),
// about a millisecond per match -- 40 times the device sweep it follows.
//
// Bit-exactness: the per-beam term is exp(-0.5 * d / variance) where d is the squared distance to
// the nearest admissible cell of a (2K+1)^2 window, i.e. one of (K+1)^2 + 1 doubles known up
// front.  The host evaluates those with glibc's exp (the reference's own arithmetic) and the
// kernel only selects among them, so no device transcendental reaches the result.  The beam's
// hit / missed cells come from device sincos; a coordinate within the edge guard band of a cell
// boundary is flagged and re-derived on the host with glibc (as in lgs_csm.cu), and the affected
// poses are re-evaluated.  The sum runs in beam order, one subtraction at a time, like the CPU.
//
// One launch per call in the common case: block = one sensor pose; all threads evaluate their
// beams' terms into shared memory, thread 0 folds them in beam order.
#include <cmath>

#include "lgs_internal.cuh"

namespace {

constexpr int kMaxKernel = 7;                       // window half size K (reference default 1)
constexpr int kTab = (kMaxKernel + 1) * (kMaxKernel + 1);
constexpr int kChunk = 2048;                        // beams folded per shared-memory round
constexpr int kFlagCap = 1 << 16;
constexpr int kThreads = 256;

struct CostTables {                                 // kernel parameter (1 KB)
    double sq[kTab];                                // SquaredDistance for (|ky|, |kx|)
    double ex[kTab];                                // exp(-0.5 * sq / variance), host glibc
    double sqDefault, exDefault;                    // no admissible cell: distance (K+1, K+1)
    double occThr, dist, scaling;
    int K;
};

struct CostPose {                                   // one evaluation
    double x, y, t;
    double minRange, maxRange;                      // already combined with the scan's own limits
    int beamBegin, nBeams;
};

struct CostFlag { int pose, beam; };
struct CostPatch { int pose, beam, hx, hy, mx, my; };

struct GridView {
    const double* origin;
    double minX, minY, res;
    int nx, ny, pitch, offX, offY;
};

__device__ __forceinline__ double cellValue(const GridView& g, int x, int y) {
    x -= g.offX; y -= g.offY;                       // GridMap::Value: 0.0 outside the map
    return (x >= 0 && x < g.nx && y >= 0 && y < g.ny) ? __ldg(g.origin + (long long)y * g.pitch + x) : 0.0;
}

// grid = one block per evaluated pose (poseList maps block -> pose on the fix-up pass).  KT >= 0 fixes
// the window half size at compile time (the reference default 1: all 18 cell reads of a beam are
// issued together instead of one dependent L2 round trip per window cell); KT < 0 reads it from tab.
template <bool PATCHED, int KT>
__global__ void __launch_bounds__(kThreads)
cost_kernel(const CostPose* __restrict__ poses, const int* __restrict__ poseList,
            const double* __restrict__ angles, const double* __restrict__ ranges, GridView g,
            CostTables tab, double eps, const CostPatch* __restrict__ patches,
            const int* __restrict__ patchBegin, CostFlag* __restrict__ flags,
            int* __restrict__ flagCount, double* __restrict__ cost) {
    __shared__ double sTerm[kChunk];
    const int p = poseList ? poseList[blockIdx.x] : blockIdx.x;
    const CostPose ps = poses[p];
    int pLo = 0, pHi = 0;
    if (PATCHED) { pLo = patchBegin[blockIdx.x]; pHi = patchBegin[blockIdx.x + 1]; }
    double value = 0.0;                             // thread 0 only
    for (int c0 = 0; c0 < ps.nBeams; c0 += kChunk) {
        const int cn = min(kChunk, ps.nBeams - c0);
        for (int j = threadIdx.x; j < cn; j += kThreads) {
            const int i = c0 + j;
            const double r = ranges[ps.beamBegin + i];
            double term = 0.0;                      // skipped beam: value - 0.0 == value
            if (!(r >= ps.maxRange || r <= ps.minRange)) {                  // :49-50
                int hx, hy, mx, my;
                bool havePatch = false;
                if (PATCHED) {
                    for (int k = pLo; k < pHi; ++k)
                        if (patches[k].beam == i) {
                            hx = patches[k].hx; hy = patches[k].hy; mx = patches[k].mx; my = patches[k].my;
                            havePatch = true;
                        }
                }
                if (!havePatch) {
                    double s, c;
                    sincos(__dadd_rn(ps.t, angles[ps.beamBegin + i]), &s, &c);
                    const double rm = __dsub_rn(r, tab.dist);
                    const double q0 = __ddiv_rn(__dsub_rn(__dadd_rn(ps.x, __dmul_rn(r, c)), g.minX), g.res);
                    const double q1 = __ddiv_rn(__dsub_rn(__dadd_rn(ps.y, __dmul_rn(r, s)), g.minY), g.res);
                    const double q2 = __ddiv_rn(__dsub_rn(__dadd_rn(ps.x, __dmul_rn(rm, c)), g.minX), g.res);
                    const double q3 = __ddiv_rn(__dsub_rn(__dadd_rn(ps.y, __dmul_rn(rm, s)), g.minY), g.res);
                    const double f0 = q0 - floor(q0), f1 = q1 - floor(q1), f2 = q2 - floor(q2), f3 = q3 - floor(q3);
                    const double lo = fmin(fmin(f0, f1), fmin(f2, f3)), hi = fmax(fmax(f0, f1), fmax(f2, f3));
                    if (!PATCHED && !(lo >= eps && hi <= 1.0 - eps)) {      // also catches NaN
                        const int k = atomicAdd(flagCount, 1);
                        if (k < kFlagCap) flags[k] = CostFlag{p, i};
                    }
                    hx = __double2int_rd(q0); hy = __double2int_rd(q1);
                    mx = __double2int_rd(q2); my = __double2int_rd(q3);
                }
                double best = tab.sqDefault;        // :64-66
                int arg = -1;
                if constexpr (KT >= 0) {
                    constexpr int W = 2 * KT + 1;
                    double hv[W * W], mv[W * W];
#pragma unroll
                    for (int w = 0; w < W * W; ++w) {
                        hv[w] = cellValue(g, hx + w % W - KT, hy + w / W - KT);
                        mv[w] = cellValue(g, mx + w % W - KT, my + w / W - KT);
                    }
#pragma unroll
                    for (int w = 0; w < W * W; ++w) {                       // ky outer, kx inner (:68-69)
                        const int e = abs(w / W - KT) * (KT + 1) + abs(w % W - KT);
                        const bool ok = !(hv[w] == 0.0 || mv[w] == 0.0) &&              // :81-83
                                        !(hv[w] < tab.occThr || mv[w] > tab.occThr);    // :89-91
                        if (ok && tab.sq[e] < best) { best = tab.sq[e]; arg = e; }
                    }
                } else {
                    const int K = tab.K;
                    for (int ky = -K; ky <= K; ++ky)
                        for (int kx = -K; kx <= K; ++kx) {
                            const double hv = cellValue(g, hx + kx, hy + ky);
                            const double mv = cellValue(g, mx + kx, my + ky);
                            if (hv == 0.0 || mv == 0.0) continue;               // :81-83
                            if (hv < tab.occThr || mv > tab.occThr) continue;   // :89-91
                            const int e = abs(ky) * (K + 1) + abs(kx);
                            if (tab.sq[e] < best) { best = tab.sq[e]; arg = e; }
                        }
                }
                term = arg < 0 ? tab.exDefault : tab.ex[arg];
            }
            sTerm[j] = term;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll 8
            for (int j = 0; j < cn; ++j) value = __dsub_rn(value, sTerm[j]);   // :103, beam order
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) cost[p] = __dmul_rn(value, tab.scaling);          // :107
}

}  // namespace

// Persistent scratch of the cost entry points (grown on demand, freed with the context).
struct lgs_cost_ws {
    DevBuf<CostPose> dPoses;
    PinBuf<CostPose> hPoses;
    DevBuf<double> dBeams;                           // angles, then ranges
    DevBuf<double> dCost;
    PinBuf<double> hCost;
    DevBuf<CostFlag> dFlags;
    DevBuf<int> dFlagCount;
    PinBuf<int> hFlagCount;
    void release() {
        dPoses.release(); hPoses.release(); dBeams.release(); dCost.release(); hCost.release(); dFlags.release();
        dFlagCount.release(); hFlagCount.release();
    }
};

void lgs_cost_ws_destroy(lgs_cost_ws* ws) {
    if (!ws) return;
    ws->release();
    delete ws;
}

namespace {

int buildTables(lgs_ctx* c, const lgs_grid* g, const lgs_cost_params* p, CostTables* t) {
    if (p->kernel_size < 0 || p->kernel_size > kMaxKernel)
        return lgs_fail(c, LGS_ERR_INVALID, "cost: kernel size %d outside [0, %d]", p->kernel_size, kMaxKernel);
    const int K = p->kernel_size;
    const double variance = p->standard_deviation * p->standard_deviation;  // ctor, :24
    for (int ky = 0; ky <= K; ++ky)
        for (int kx = 0; kx <= K; ++kx) {
            const double dx = kx * g->res, dy = ky * g->res;                // grid_map.hpp:899-901
            const double sq = dx * dx + dy * dy;
            t->sq[ky * (K + 1) + kx] = sq;
            t->ex[ky * (K + 1) + kx] = std::exp(-0.5 * sq / variance);      // :103
        }
    const double d0 = (K + 1) * g->res;
    t->sqDefault = d0 * d0 + d0 * d0;
    t->exDefault = std::exp(-0.5 * t->sqDefault / variance);
    t->occThr = p->occupancy_threshold; t->dist = p->hit_and_missed_dist; t->scaling = p->scaling_factor;
    t->K = K;
    return LGS_OK;
}

inline int worldToCell(double v, double minV, double res) {
    return static_cast<int>(std::floor((v - minV) / res));                  // grid_map.hpp:779-790
}

}  // namespace

extern "C" {

int lgs_cost_greedy_endpoint(lgs_ctx* c, const lgs_grid* grid, const lgs_cost_params* params,
                             const lgs_scan_batch* scans, int nPoses, const int* poseScan,
                             const double* poses, double* cost, int* nFixups) {
    if (!c || !grid || !params || !scans || nPoses < 0 || (nPoses > 0 && (!poses || !cost)))
        return LGS_ERR_INVALID;
    if (nFixups) *nFixups = 0;
    if (nPoses == 0) return LGS_OK;
    if (scans->n_scans <= 0 || !scans->beam_begin || !scans->angles || !scans->ranges)
        return lgs_fail(c, LGS_ERR_INVALID, "cost: empty scan batch");
    if (!poseScan && nPoses != scans->n_scans)
        return lgs_fail(c, LGS_ERR_INVALID, "cost: %d poses for %d scans need pose_scan", nPoses, scans->n_scans);
    CostTables tab;
    int rc = buildTables(c, grid, params, &tab);
    if (rc != LGS_OK) return rc;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, lgs_grid_acquire(c, grid));
    if (!c->cost) c->cost = new lgs_cost_ws();
    lgs_cost_ws* ws = c->cost;

    const int nScans = scans->n_scans;
    const size_t nBeams = (size_t)scans->beam_begin[nScans];
    LGS_CUDA(c, ws->hPoses.reserve(nPoses));
    LGS_CUDA(c, ws->dPoses.reserve(nPoses));
    LGS_CUDA(c, ws->dBeams.reserve(2 * nBeams + 1));
    LGS_CUDA(c, ws->dCost.reserve(nPoses));
    LGS_CUDA(c, ws->hCost.reserve(nPoses));
    LGS_CUDA(c, ws->dFlags.reserve(kFlagCap));
    LGS_CUDA(c, ws->dFlagCount.reserve(1));
    LGS_CUDA(c, ws->hFlagCount.reserve(1));
    CostPose* hp = ws->hPoses.p;
    for (int p = 0; p < nPoses; ++p) {
        const int s = poseScan ? poseScan[p] : p;
        if (s < 0 || s >= nScans) return lgs_fail(c, LGS_ERR_INVALID, "cost: pose %d names scan %d", p, s);
        const double scanMin = scans->range_min ? scans->range_min[s] : 0.0;
        const double scanMax = scans->range_max ? scans->range_max[s] : INFINITY;
        hp[p] = CostPose{poses[3 * p], poses[3 * p + 1], poses[3 * p + 2],
                         std::max(params->usable_range_min, scanMin),       // :39-42
                         std::min(params->usable_range_max, scanMax),
                         scans->beam_begin[s], scans->beam_begin[s + 1] - scans->beam_begin[s]};
    }
    // The beams go to the device straight from the caller's arrays (full speed when the caller page
    // locked them with lgs_host_pin, driver staged otherwise); the call is synchronous, so the rare
    // host fix-up below reads the same arrays.
    const double* hAngles = scans->angles;
    const double* hRanges = scans->ranges;
    const CostPose* dPoses = ws->dPoses.p;
    double* dAngles = ws->dBeams.p;
    double* dRanges = dAngles + nBeams;

    GridView gv{grid->origin(), grid->min_x, grid->min_y, grid->res, grid->nx, grid->ny, grid->pitch,
                grid->off_x, grid->off_y};
    LGS_CUDA(c, cudaMemcpyAsync(ws->dPoses.p, hp, nPoses * sizeof(CostPose), cudaMemcpyHostToDevice, c->stream));
    if (nBeams > 0) {
        LGS_CUDA(c, cudaMemcpyAsync(dAngles, hAngles, nBeams * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(dRanges, hRanges, nBeams * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    LGS_CUDA(c, cudaMemsetAsync(ws->dFlagCount.p, 0, sizeof(int), c->stream));
    if (tab.K == 1)
        cost_kernel<false, 1><<<nPoses, kThreads, 0, c->stream>>>(dPoses, nullptr, dAngles, dRanges, gv, tab,
                                                                  c->opt.edgeEps, nullptr, nullptr, ws->dFlags.p,
                                                                  ws->dFlagCount.p, ws->dCost.p);
    else
        cost_kernel<false, -1><<<nPoses, kThreads, 0, c->stream>>>(dPoses, nullptr, dAngles, dRanges, gv, tab,
                                                                   c->opt.edgeEps, nullptr, nullptr, ws->dFlags.p,
                                                                   ws->dFlagCount.p, ws->dCost.p);
    LGS_LAUNCH_CHECK(c);
    LGS_CUDA(c, cudaMemcpyAsync(ws->hCost.p, ws->dCost.p, nPoses * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaMemcpyAsync(ws->hFlagCount.p, ws->dFlagCount.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));

    const int nFlag = *ws->hFlagCount.p;
    if (nFlag > 0) {
        // Rare path: re-derive the flagged beams' cells with the host's libm (the reference's own
        // arithmetic, sensor_data.hpp:184-195 + grid_map.hpp:779-790) and redo the affected poses.
        if (nFlag > kFlagCap)
            return lgs_fail(c, LGS_ERR_OVERFLOW, "cost: %d near-edge beams exceed the fix-up list", nFlag);
        std::vector<CostFlag> fl(nFlag);
        LGS_CUDA(c, cudaMemcpy(fl.data(), ws->dFlags.p, nFlag * sizeof(CostFlag), cudaMemcpyDeviceToHost));
        std::sort(fl.begin(), fl.end(), [](const CostFlag& a, const CostFlag& b) {
            return a.pose != b.pose ? a.pose < b.pose : a.beam < b.beam; });
        std::vector<CostPatch> patches(nFlag);
        std::vector<int> poseList, patchBegin;
        for (int k = 0; k < nFlag; ++k) {
            const CostPose& ps = hp[fl[k].pose];
            const int i = fl[k].beam;
            const double r = hRanges[ps.beamBegin + i], a = hAngles[ps.beamBegin + i];
            const double cosT = std::cos(ps.t + a), sinT = std::sin(ps.t + a);
            const double hx = ps.x + r * cosT, hy = ps.y + r * sinT;
            const double mx = ps.x + (r - params->hit_and_missed_dist) * cosT;
            const double my = ps.y + (r - params->hit_and_missed_dist) * sinT;
            patches[k] = CostPatch{fl[k].pose, i, worldToCell(hx, grid->min_x, grid->res),
                                   worldToCell(hy, grid->min_y, grid->res),
                                   worldToCell(mx, grid->min_x, grid->res),
                                   worldToCell(my, grid->min_y, grid->res)};
            if (poseList.empty() || poseList.back() != fl[k].pose) {
                poseList.push_back(fl[k].pose);
                patchBegin.push_back(k);
            }
        }
        patchBegin.push_back(nFlag);
        const int nRedo = (int)poseList.size();
        CostPatch* dPatches = nullptr; int* dList = nullptr; int* dBegin = nullptr;
        LGS_CUDA(c, lgs_alloc_async(c, &dPatches, nFlag * sizeof(CostPatch)));
        LGS_CUDA(c, lgs_alloc_async(c, &dList, nRedo * sizeof(int)));
        LGS_CUDA(c, lgs_alloc_async(c, &dBegin, (nRedo + 1) * sizeof(int)));
        LGS_CUDA(c, cudaMemcpyAsync(dPatches, patches.data(), nFlag * sizeof(CostPatch), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(dList, poseList.data(), nRedo * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(dBegin, patchBegin.data(), (nRedo + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        cost_kernel<true, -1><<<nRedo, kThreads, 0, c->stream>>>(dPoses, dList, dAngles, dRanges, gv, tab,
                                                                 c->opt.edgeEps, dPatches, dBegin, nullptr,
                                                                 nullptr, ws->dCost.p);
        LGS_LAUNCH_CHECK(c);
        LGS_CUDA(c, cudaMemcpyAsync(ws->hCost.p, ws->dCost.p, nPoses * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        LGS_CUDA(c, cudaFreeAsync(dPatches, c->stream));
        LGS_CUDA(c, cudaFreeAsync(dList, c->stream));
        LGS_CUDA(c, cudaFreeAsync(dBegin, c->stream));
        LGS_CUDA(c, cudaStreamSynchronize(c->stream));   // the pageable sources above stay alive until here
        if (nFixups) *nFixups = nFlag;
    }
    std::memcpy(cost, ws->hCost.p, nPoses * sizeof(double));
    return LGS_OK;
}

int lgs_cost_tail(lgs_ctx* c, const lgs_grid* grid, const lgs_cost_params* params,
                  const lgs_scan_batch* scans, const double* best, double* normalizedCost,
                  double* covariance, int* nFixups) {
    if (!c || !grid || !params || !scans || scans->n_scans < 0) return LGS_ERR_INVALID;
    const int n = scans->n_scans;
    if (nFixups) *nFixups = 0;
    if (n == 0) return LGS_OK;
    if (!best || !scans->beam_begin) return LGS_ERR_INVALID;
    // 7 evaluations per match: the pose itself, then pose +/- delta per axis in ComputeGradient's
    // order (:124-136); RobotPose2D +/- adds every component (pose.hpp:60-74), zeros included.
    const double diff[3] = {grid->res, grid->res, 1e-2};                    // :120-121
    std::vector<double> poses((size_t)n * 21), costs((size_t)n * 7);
    std::vector<int> poseScan((size_t)n * 7);
    for (int m = 0; m < n; ++m) {
        double* q = poses.data() + (size_t)m * 21;
        for (int k = 0; k < 3; ++k) q[k] = best[3 * m + k];
        for (int a = 0; a < 3; ++a)
            for (int k = 0; k < 3; ++k) {
                const double d = k == a ? diff[a] : 0.0;
                q[3 * (1 + 2 * a) + k] = best[3 * m + k] + d;
                q[3 * (2 + 2 * a) + k] = best[3 * m + k] - d;
            }
        for (int j = 0; j < 7; ++j) poseScan[(size_t)m * 7 + j] = m;
    }
    const int rc = lgs_cost_greedy_endpoint(c, grid, params, scans, 7 * n, poseScan.data(), poses.data(),
                                            costs.data(), nFixups);
    if (rc != LGS_OK) return rc;
    for (int m = 0; m < n; ++m) {
        const double* v = costs.data() + (size_t)m * 7;
        const size_t numOfScans = (size_t)(scans->beam_begin[m + 1] - scans->beam_begin[m]);
        if (normalizedCost) normalizedCost[m] = v[0] / numOfScans;          // rtcsm :129-130
        if (covariance) {
            double grad[3];
            for (int a = 0; a < 3; ++a) grad[a] = 0.5 * (v[1 + 2 * a] - v[2 + 2 * a]) / diff[a];   // :139-141
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    covariance[9 * m + 3 * i + j] = i == j ? grad[i] * grad[j] + 0.01 : grad[i] * grad[j];   // :161-166
        }
    }
    return LGS_OK;
}

}  // extern "C"
