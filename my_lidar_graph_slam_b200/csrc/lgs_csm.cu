// lgs_csm.cu -- real-time correlative scan matcher on sm_100a.
//
// Replaces ScanMatcherRealTimeCorrelative::OptimizePose(grid, coarse, scan, pose, thr)
// (mapping/scan_matcher_real_time_correlative.cpp:50-145) and its helpers ComputeSearchStep
// (:156-175), ComputeScanIndices (:178-203), ComputeScore (:207-224) and
// EvaluateHighResolutionMap (:227-256).
//
// Exactness strategy:
//  * csm_project_kernel projects every kept beam once per theta with the reference's exact
//    operation order (IEEE add/mul/div intrinsics, never fused).  Only sin/cos can differ from
//    glibc by an ulp, which can change floor() only within the edge guard band of a cell edge; such
//    points are flagged and re-derived on the host with glibc before the result is accepted.
//  * the sweep (csm_sweep_rows_kernel: lanes = the x offsets of one window row, 4-5 rows per
//    thread; csm_sweep_kernel: one thread per hypothesis, for small windows) sums the gathered
//    cells of every hypothesis in beam order in double: the score is bit-identical to the CPU loop
//    by construction, so no epsilon logic is needed anywhere.
//  * csm_select_kernel computes each coarse block's fine maximum with first-visit argmax and
//    then reproduces the CPU's pruned visit sequence: a parallel (score desc, visit asc)
//    reduction when every coarse score bounds its block, otherwise a sequential replay of
//    `if (C[b] > best && F[b] > best) best = F[b]` in the CPU's visit order.
//
// Data layout: projected points are stored per (match, theta) as int32 offsets into the
// apron-padded grid, clamped so that every out-of-map read lands in the zero apron; the sweep
// stages one theta's offsets in shared memory and each thread gathers grid[base + off[i]].
#include <cfloat>
#include <cmath>

#include "lgs_internal.cuh"

namespace {

constexpr int kBeamPad = 8;          // kept-beam count is padded to a multiple of this
constexpr int kFlagCap = 1 << 16;    // initial capacity of the near-edge fix-up list (grown on demand)

struct CsmDesc {
    double sx, sy, st;       // sensor pose
    double stepT;            // angular step
    double thrAbs;           // normalizedScoreThreshold * NumOfScans()
    int winT, nT;            // theta half window, 2 * winT + 1
    int nKept, nKeptPad;     // beams with range < scanRangeMax
    int beamBegin;           // into the kept angle / range arrays
    int pad0;
    long long offBegin;      // into offs  : nT * nKeptPad
    long long cellBegin;     // into cells : nT * nKept (int2)
    long long fineBegin;     // into fine table   : nT * nyw * nxw
    long long coarseBegin;   // into coarse table : nT * nbx * nby
};

struct CsmWindow {           // identical for every match of a batch (depends on grid + params)
    int winX, winY;
    int nbx, nby;            // coarse blocks per axis
    int nxw, nyw;            // fine hypotheses per axis = nb * lowRes
    int lowRes;
    int tilesFine, tilesCoarse;   // thread blocks per theta
    int block;               // threads per block
    int hpt;                 // fine hypotheses per thread (1 or 4)
    int slots;               // thread slots per theta = ceil(nHyp / hpt) rounded up to a warp
    int rows;                // > 0: row-mapped sweep, fine rows per thread (4 or 5)
    int xChunks, rowGroups;  // warps per theta = xChunks * rowGroups (lanes = 32 consecutive x)
};

struct GridGeom {
    double minX, minY, res;
    int nx, ny, pitch;
};

struct DevResult {
    double score;
    int found, ix, iy, it;
    int exactReplay, pad;
};

struct FlagEntry { int m, t, i; };

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- K1: scan projection ----------------------------------------------------------------------
// One thread per (match, theta, kept beam).  scan_matcher_real_time_correlative.cpp:88-96,
// :178-203; sensor_data.hpp:162-173; grid_map.hpp:779-790.
__global__ void csm_project_kernel(const CsmDesc* __restrict__ descs,
                                   const double* __restrict__ angles,
                                   const double* __restrict__ ranges, GridGeom g, CsmWindow w, double eps,
                                   int* __restrict__ offs, int2* __restrict__ cells,
                                   FlagEntry* __restrict__ flags, int flagCap, int* __restrict__ flagCount) {
    const CsmDesc d = descs[blockIdx.y];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)d.nT * d.nKeptPad) return;
    const int t = (int)(idx / d.nKeptPad);
    const int i = (int)(idx - (long long)t * d.nKeptPad);
    // Clamp targets: every window offset of a clamped point reads the zero apron.
    const int xLo = -w.winX, xHi = -w.winX + w.nxw - 1;
    const int yLo = -w.winY, yHi = -w.winY + w.nyw - 1;
    if (i >= d.nKept) {   // padding beam: contributes exactly +0.0
        offs[d.offBegin + idx] = (-yHi - 1) * g.pitch + (-xHi - 1);
        return;
    }
    const double theta = __dadd_rn(d.st, __dmul_rn(d.stepT, (double)(t - d.winT)));
    const double a = __dadd_rn(theta, angles[d.beamBegin + i]);
    double s, c;
    sincos(a, &s, &c);
    const double r = ranges[d.beamBegin + i];
    const double hx = __dadd_rn(d.sx, __dmul_rn(r, c));
    const double hy = __dadd_rn(d.sy, __dmul_rn(r, s));
    const double qx = __ddiv_rn(__dsub_rn(hx, g.minX), g.res);
    const double qy = __ddiv_rn(__dsub_rn(hy, g.minY), g.res);
    const double fx = floor(qx), fy = floor(qy);
    const double rx = qx - fx, ry = qy - fy;
    const bool edge = !(rx >= eps && rx <= 1.0 - eps && ry >= eps && ry <= 1.0 - eps);
    if (edge) {
        const int k = atomicAdd(flagCount, 1);
        if (k < flagCap) flags[k] = FlagEntry{(int)blockIdx.y, t, i};
    }
    const int cx = __double2int_rd(qx), cy = __double2int_rd(qy);
    cells[d.cellBegin + (long long)t * d.nKept + i] = make_int2(cx, cy);
    const int ccx = clampi(cx, -xHi - 1, g.nx - xLo);
    const int ccy = clampi(cy, -yHi - 1, g.ny - yLo);
    offs[d.offBegin + idx] = ccy * g.pitch + ccx;
}

// ---- K2: the sweep (hot kernel) -------------------------------------------------------------------
// grid = (tilesFine + tilesCoarse, maxNT, matches).  One thread per hypothesis; the theta's
// offsets are staged in shared memory; each thread sums grid[base + off[i]] in beam order.
// Lanes are consecutive x offsets so a warp's gathers for one beam fall in 1-3 cache lines.
// scan_matcher_real_time_correlative.cpp:98-102 (coarse), :207-224, :239-244 (fine).
template <int UNROLL, int HPT>
__global__ void __launch_bounds__(256)
csm_sweep_kernel(const CsmDesc* __restrict__ descs, const double* __restrict__ fineGrid,
                 const double* __restrict__ coarseGrid, int pitch, CsmWindow w,
                 const int* __restrict__ offs, double* __restrict__ fineTab,
                 double* __restrict__ coarseTab) {
    extern __shared__ int sOff[];
    const CsmDesc d = descs[blockIdx.z];
    const int t = blockIdx.y;
    if (t >= d.nT) return;
    {   // stage this theta's offsets (nKeptPad is a multiple of 4 ints = 16 B)
        const int4* src = reinterpret_cast<const int4*>(offs + d.offBegin + (long long)t * d.nKeptPad);
        int4* dst = reinterpret_cast<int4*>(sOff);
        for (int k = threadIdx.x; k < d.nKeptPad / 4; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const int n = d.nKeptPad;   // multiple of kBeamPad >= UNROLL

    if (blockIdx.x >= (unsigned)w.tilesFine) {
        // Coarse hypotheses (stride lowRes on the win-max map), one per thread.
        const int h = (blockIdx.x - w.tilesFine) * blockDim.x + threadIdx.x;
        if (h >= w.nbx * w.nby) return;
        const int oy = h / w.nbx, ox = h - oy * w.nbx;
        const double* __restrict__ gp =
            coarseGrid + (-w.winY + oy * w.lowRes) * pitch + (-w.winX + ox * w.lowRes);
        double acc = 0.0;
#pragma unroll 1
        for (int i = 0; i < n; i += UNROLL) {
            double v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(gp + sOff[i + u]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc = __dadd_rn(acc, v[u]);
        }
        // stored in the CPU's visit order: x outer, y inner
        coarseTab[d.coarseBegin + ((long long)t * w.nbx + ox) * w.nby + oy] = acc;
        return;
    }

    // Fine hypotheses: thread slot j owns hypotheses j, j + S, ..., j + (HPT-1) S of the flattened
    // (y, x) window, so one shared-memory offset load feeds HPT gathers and a warp's lanes stay
    // on consecutive x cells for every one of them.
    const int nHyp = w.nxw * w.nyw;
    const int S = w.slots;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    const double* __restrict__ gp[HPT];
    double acc[HPT];
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
        const int h = min(j + k * S, nHyp - 1);          // out-of-range slots alias the last one
        const int oy = h / w.nxw, ox = h - oy * w.nxw;
        gp[k] = fineGrid + (-w.winY + oy) * pitch + (-w.winX + ox);
        acc[k] = 0.0;
    }
#pragma unroll 1
    for (int i = 0; i < n; i += UNROLL) {
        double v[UNROLL][HPT];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int o = sOff[i + u];
#pragma unroll
            for (int k = 0; k < HPT; ++k) v[u][k] = __ldg(gp[k] + o);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int k = 0; k < HPT; ++k) acc[k] = __dadd_rn(acc[k], v[u][k]);   // beam order, like the CPU
    }
#pragma unroll
    for (int k = 0; k < HPT; ++k) {
        const int h = j + k * S;
        if (h < nHyp) {
            const int oy = h / w.nxw, ox = h - oy * w.nxw;
            fineTab[d.fineBegin + ((long long)t * w.nyw + oy) * w.nxw + ox] = acc[k];
        }
    }
}

// ---- K2b: the sweep, row mapped --------------------------------------------------------------------
// Same arithmetic as csm_sweep_kernel, different ownership: a warp's lanes are the x offsets of ONE
// window row (25 of 32 lanes for config C2) and a thread owns ROWS consecutive rows.  A 32-lane
// request then touches the 2-3 cache lines of a single 25-cell row (2.5 L1 wavefronts on average)
// instead of the 3.8 lines of 1.28 rows of the flattened mapping: ~15 % fewer L1 wavefronts for the
// same gathers, and the L1 data pipe is what bounds this kernel (profiles/r1_kernels_v2.md).
template <int UNROLL, int ROWS>
__global__ void __launch_bounds__(256)
csm_sweep_rows_kernel(const CsmDesc* __restrict__ descs, const double* __restrict__ fineGrid,
                      const double* __restrict__ coarseGrid, int pitch, CsmWindow w,
                      const int* __restrict__ offs, double* __restrict__ fineTab,
                      double* __restrict__ coarseTab) {
    extern __shared__ int sOff[];
    const CsmDesc d = descs[blockIdx.z];
    const int t = blockIdx.y;
    if (t >= d.nT) return;
    {   // stage this theta's offsets (nKeptPad is a multiple of 4 ints = 16 B)
        const int4* src = reinterpret_cast<const int4*>(offs + d.offBegin + (long long)t * d.nKeptPad);
        int4* dst = reinterpret_cast<int4*>(sOff);
        for (int k = threadIdx.x; k < d.nKeptPad / 4; k += blockDim.x) dst[k] = src[k];
    }
    __syncthreads();
    const int n = d.nKeptPad;   // multiple of kBeamPad >= UNROLL

    if (blockIdx.x >= (unsigned)w.tilesFine) {
        // Coarse hypotheses (stride lowRes on the win-max map), one per thread.
        const int h = (blockIdx.x - w.tilesFine) * blockDim.x + threadIdx.x;
        if (h >= w.nbx * w.nby) return;
        const int oy = h / w.nbx, ox = h - oy * w.nbx;
        const double* __restrict__ gp =
            coarseGrid + (-w.winY + oy * w.lowRes) * pitch + (-w.winX + ox * w.lowRes);
        double acc = 0.0;
#pragma unroll 1
        for (int i = 0; i < n; i += UNROLL) {
            double v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(gp + sOff[i + u]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) acc = __dadd_rn(acc, v[u]);
        }
        coarseTab[d.coarseBegin + ((long long)t * w.nbx + ox) * w.nby + oy] = acc;
        return;
    }

    const int wg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int xc = wg % w.xChunks, rg = wg / w.xChunks;
    const int ox = xc * 32 + (threadIdx.x & 31);
    if (rg >= w.rowGroups || ox >= w.nxw) return;
    const int oy0 = rg * ROWS;
    const double* __restrict__ gp = fineGrid + (-w.winY + oy0) * pitch + (-w.winX + ox);
    double acc[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) acc[k] = 0.0;
    const int nRows = min(ROWS, w.nyw - oy0);
    if (nRows == ROWS) {
#pragma unroll 1
        for (int i = 0; i < n; i += UNROLL) {
            double v[UNROLL][ROWS];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const double* __restrict__ q = gp + sOff[i + u];
#pragma unroll
                for (int k = 0; k < ROWS; ++k) v[u][k] = __ldg(q + k * pitch);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int k = 0; k < ROWS; ++k) acc[k] = __dadd_rn(acc[k], v[u][k]);   // beam order, like the CPU
        }
    } else {                                   // last row group of a window whose height is no multiple of ROWS
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const double* __restrict__ q = gp + sOff[i];
#pragma unroll
            for (int k = 0; k < ROWS; ++k)
                if (k < nRows) acc[k] = __dadd_rn(acc[k], __ldg(q + k * pitch));
        }
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k)
        if (k < nRows) fineTab[d.fineBegin + ((long long)t * w.nyw + oy0 + k) * w.nxw + ox] = acc[k];
}

// ---- K3: block maxima + CPU-order selection -----------------------------------------------------
// One thread block per match.
__global__ void __launch_bounds__(256)
csm_select_kernel(const CsmDesc* __restrict__ descs, CsmWindow w,
                  const double* __restrict__ fineTab, const double* __restrict__ coarseTab,
                  double* __restrict__ blockMax, int* __restrict__ blockArg,
                  DevResult* __restrict__ results) {
    const CsmDesc d = descs[blockIdx.x];
    const int nB = d.nT * w.nbx * w.nby;
    double* F = blockMax + d.coarseBegin;
    int* A = blockArg + d.coarseBegin;
    const double* C = coarseTab + d.coarseBegin;
    const double* fine = fineTab + d.fineBegin;
    __shared__ int sInvalid;
    __shared__ double sBest[256];
    __shared__ int sIdx[256];
    if (threadIdx.x == 0) sInvalid = 0;
    __syncthreads();

    // Phase 1: F[v] = max of the lowRes x lowRes fine scores of block v, first visit wins
    // (x outer, y inner; strict '<' update: scan_matcher_real_time_correlative.cpp:239-251).
    double myBest = -1.0;
    int myIdx = 0x7fffffff;
    bool invalid = false;
    for (int v = threadIdx.x; v < nB; v += blockDim.x) {
        const int by = v % w.nby;
        const int bx = (v / w.nby) % w.nbx;
        const int t = v / (w.nby * w.nbx);
        const double* ft = fine + (long long)t * w.nyw * w.nxw;
        double m = -1.0;
        int arg = 0;
        for (int fx = 0; fx < w.lowRes; ++fx)
            for (int fy = 0; fy < w.lowRes; ++fy) {
                const double s = ft[(by * w.lowRes + fy) * w.nxw + bx * w.lowRes + fx];
                if (m < s) { m = s; arg = fx * w.lowRes + fy; }
            }
        F[v] = m;
        A[v] = arg;
        if (C[v] < m) invalid = true;   // the coarse value is not an upper bound (H12)
        if (m > myBest) { myBest = m; myIdx = v; }   // v ascending per thread: first visit kept
    }
    if (invalid) atomicOr(&sInvalid, 1);
    sBest[threadIdx.x] = myBest;
    sIdx[threadIdx.x] = myIdx;
    __syncthreads();

    DevResult r;
    r.pad = 0;
    if (!sInvalid) {
        // Every C[v] >= F[v]: the pruned CPU loop returns the first visit of the global maximum.
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                const double ob = sBest[threadIdx.x + s];
                const int oi = sIdx[threadIdx.x + s];
                if (ob > sBest[threadIdx.x] || (ob == sBest[threadIdx.x] && oi < sIdx[threadIdx.x])) {
                    sBest[threadIdx.x] = ob;
                    sIdx[threadIdx.x] = oi;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x != 0) return;
        r.exactReplay = 0;
        if (nB > 0 && sBest[0] > d.thrAbs) {
            const int v = sIdx[0];
            const int by = v % w.nby, bx = (v / w.nby) % w.nbx, t = v / (w.nby * w.nbx);
            r.found = 1; r.score = sBest[0];
            r.ix = -w.winX + bx * w.lowRes + A[v] / w.lowRes;
            r.iy = -w.winY + by * w.lowRes + A[v] % w.lowRes;
            r.it = t - d.winT;
        } else {
            r.found = 0; r.score = d.thrAbs;
            r.ix = -w.winX; r.iy = -w.winY; r.it = -d.winT;
        }
        results[blockIdx.x] = r;
        return;
    }
    // Sequential replay of the CPU's visit sequence (:88-114).
    if (threadIdx.x != 0) return;
    double best = d.thrAbs;
    int bv = -1;
    for (int v = 0; v < nB; ++v) {
        if (C[v] <= best) continue;
        if (best < F[v]) { best = F[v]; bv = v; }
    }
    r.exactReplay = 1;
    if (bv >= 0) {
        const int by = bv % w.nby, bx = (bv / w.nby) % w.nbx, t = bv / (w.nby * w.nbx);
        r.found = 1; r.score = best;
        r.ix = -w.winX + bx * w.lowRes + A[bv] / w.lowRes;
        r.iy = -w.winY + by * w.lowRes + A[bv] % w.lowRes;
        r.it = t - d.winT;
    } else {
        r.found = 0; r.score = d.thrAbs;
        r.ix = -w.winX; r.iy = -w.winY; r.it = -d.winT;
    }
    results[blockIdx.x] = r;
}

__global__ void csm_patch_kernel(const int* __restrict__ where, const int* __restrict__ offVal,
                                 const int2* __restrict__ cellVal, const long long* cellWhere,
                                 int n, int* offs, int2* cells) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    (void)where;
    offs[cellWhere[2 * k]] = offVal[k];
    cells[cellWhere[2 * k + 1]] = cellVal[k];
}

int pick_block(int nHyp) {
    // Threads per block: the multiple of 32 in [64, 256] wasting the fewest lanes, larger wins ties.
    int best = 128;
    long long bestWaste = 1LL << 60;
    for (int b = 256; b >= 64; b -= 32) {
        const long long tiles = (nHyp + b - 1) / b;
        const long long waste = tiles * b - nHyp;
        if (waste < bestWaste) { bestWaste = waste; best = b; }
    }
    return best;
}

}  // namespace

struct lgs_rtcsm_batch {
    lgs_ctx* ctx = nullptr;
    lgs_rtcsm_params params{};
    int nMatch = 0;
    int maxNT = 0, maxKeptPad = 0;
    CsmWindow win{};
    GridGeom geom{};
    std::vector<CsmDesc> descs;
    std::vector<double> hAngles, hRanges;        // kept beams, host copy (for fix-ups)
    std::vector<double> stepT;
    std::vector<int> fixups;
    long long nOff = 0, nCell = 0, nFine = 0, nCoarse = 0;
    long long workHyp = 0, workGather = 0;
    DevBuf<CsmDesc> dDescs;
    DevBuf<double> dAngles, dRanges, dFine, dCoarse, dBlockMax;
    DevBuf<int> dOffs, dBlockArg, dFlagCount;
    DevBuf<int2> dCells;
    DevBuf<FlagEntry> dFlags;
    int flagCap = kFlagCap;
    // near-edge fix-up values (results()): kept with the batch, no per-call cudaMalloc
    DevBuf<int> dFixOff;
    DevBuf<int2> dFixCell;
    DevBuf<long long> dFixWhere;
    DevBuf<DevResult> dResults;
    PinBuf<DevResult> hResults;
    PinBuf<int> hFlagCount;
    PinBuf<CsmDesc> hDescs;
    PinBuf<double> hStage;
    bool uploaded = false, ran = false;
};

static int csm_launch_sweep(lgs_rtcsm_batch* b, const lgs_grid* grid, const lgs_grid* coarse) {
    lgs_ctx* c = b->ctx;
    const CsmWindow& w = b->win;
    for (int m0 = 0; m0 < b->nMatch; m0 += 65535) {
        const int nm = std::min(65535, b->nMatch - m0);
        dim3 gridDim(w.tilesFine + w.tilesCoarse, b->maxNT, nm);
        const size_t smem = (size_t)b->maxKeptPad * sizeof(int);
        if (w.rows == 5)
            csm_sweep_rows_kernel<4, 5><<<gridDim, w.block, smem, c->stream>>>(
                b->dDescs.p + m0, grid->origin(), coarse->origin(), grid->pitch, w, b->dOffs.p,
                b->dFine.p, b->dCoarse.p);
        else if (w.rows == 4)
            csm_sweep_rows_kernel<4, 4><<<gridDim, w.block, smem, c->stream>>>(
                b->dDescs.p + m0, grid->origin(), coarse->origin(), grid->pitch, w, b->dOffs.p,
                b->dFine.p, b->dCoarse.p);
        else if (w.hpt == 4)
            csm_sweep_kernel<4, 4><<<gridDim, w.block, smem, c->stream>>>(
                b->dDescs.p + m0, grid->origin(), coarse->origin(), grid->pitch, w, b->dOffs.p,
                b->dFine.p, b->dCoarse.p);
        else
            csm_sweep_kernel<kBeamPad, 1><<<gridDim, w.block, smem, c->stream>>>(
                b->dDescs.p + m0, grid->origin(), coarse->origin(), grid->pitch, w, b->dOffs.p,
                b->dFine.p, b->dCoarse.p);
        LGS_LAUNCH_CHECK(c);
    }
    return LGS_OK;
}

static int csm_launch_select(lgs_rtcsm_batch* b) {
    lgs_ctx* c = b->ctx;
    csm_select_kernel<<<b->nMatch, 256, 0, c->stream>>>(b->dDescs.p, b->win, b->dFine.p,
                                                        b->dCoarse.p, b->dBlockMax.p,
                                                        b->dBlockArg.p, b->dResults.p);
    LGS_LAUNCH_CHECK(c);
    return LGS_OK;
}

static int csm_launch_sweep_select(lgs_rtcsm_batch* b, const lgs_grid* grid, const lgs_grid* coarse) {
    const int rc = csm_launch_sweep(b, grid, coarse);
    return rc != LGS_OK ? rc : csm_launch_select(b);
}

extern "C" {

int lgs_rtcsm_batch_create(lgs_ctx* ctx, const lgs_rtcsm_params* p, lgs_rtcsm_batch** out) {
    if (!ctx || !p || !out) return LGS_ERR_INVALID;
    *out = nullptr;
    if (p->low_res < 1 || !(p->range_x >= 0) || !(p->range_y >= 0) || !(p->range_theta >= 0))
        return lgs_fail(ctx, LGS_ERR_INVALID, "rtcsm_batch_create: bad parameters");
    lgs_rtcsm_batch* b = new lgs_rtcsm_batch();
    b->ctx = ctx;
    b->params = *p;
    *out = b;
    return LGS_OK;
}

int lgs_rtcsm_batch_destroy(lgs_rtcsm_batch* b) {
    if (!b) return LGS_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    b->dDescs.release(); b->dAngles.release(); b->dRanges.release(); b->dFine.release();
    b->dCoarse.release(); b->dBlockMax.release(); b->dOffs.release(); b->dBlockArg.release();
    b->dFlagCount.release(); b->dCells.release(); b->dFlags.release(); b->dResults.release();
    b->dFixOff.release(); b->dFixCell.release(); b->dFixWhere.release();
    b->hResults.release(); b->hFlagCount.release(); b->hDescs.release(); b->hStage.release();
    delete b;
    return LGS_OK;
}

int lgs_rtcsm_batch_upload(lgs_rtcsm_batch* b, const lgs_grid* grid, const lgs_scan_batch* scans,
                           const double* normThr) {
    if (!b || !grid || !scans) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    const int n = scans->n_scans;
    if (n < 0 || (n > 0 && (!scans->beam_begin || !scans->sensor_pose)))
        return lgs_fail(c, LGS_ERR_INVALID, "rtcsm_batch_upload: bad scan batch");
    if (grid->off_x || grid->off_y)
        return lgs_fail(c, LGS_ERR_INVALID, "rtcsm_batch_upload: windowed grids (lgs_grid_set_window) are not supported");
    LGS_CUDA(c, cudaSetDevice(c->device));
    const lgs_rtcsm_params& p = b->params;
    b->uploaded = false; b->ran = false;
    b->nMatch = n;

    // Search step and window: scan_matcher_real_time_correlative.cpp:61-73, :156-175.
    const double stepX = grid->res, stepY = grid->res;
    CsmWindow& w = b->win;
    w.lowRes = p.low_res;
    w.winX = static_cast<int>(std::ceil(0.5 * p.range_x / stepX));
    w.winY = static_cast<int>(std::ceil(0.5 * p.range_y / stepY));
    // for (x = -winX; x <= winX; x += lowRes): SURVEY.md H2 (asymmetric window).
    w.nbx = (2 * w.winX) / w.lowRes + 1;
    w.nby = (2 * w.winY) / w.lowRes + 1;
    w.nxw = w.nbx * w.lowRes;
    w.nyw = w.nby * w.lowRes;
    if (std::max(w.nxw, w.nyw) > grid->apron)
        return lgs_fail(c, LGS_ERR_APRON, "rtcsm: window %dx%d cells needs apron >= %d, grid has %d",
                        w.nxw, w.nyw, std::max(w.nxw, w.nyw), grid->apron);
    w.hpt = (w.nxw * w.nyw >= 512) ? 4 : 1;
    w.slots = (((w.nxw * w.nyw + w.hpt - 1) / w.hpt) + 31) / 32 * 32;
    w.block = pick_block(w.slots);
    w.tilesFine = (w.slots + w.block - 1) / w.block;
    w.rows = 0; w.xChunks = 1; w.rowGroups = 0;
    if (w.hpt == 4 && !c->opt.csmFlat) {
        // row-mapped sweep: lanes = x offsets of one window row, ROWS rows per thread
        w.rows = (w.nyw % 5 == 0) ? 5 : 4;
        w.xChunks = (w.nxw + 31) / 32;
        w.rowGroups = (w.nyw + w.rows - 1) / w.rows;
        const int warps = w.xChunks * w.rowGroups;
        int wpb = std::min(8, warps);                        // warps per block: fewest idle warps, larger wins
        long long bestWaste = 1LL << 60;
        for (int cand = std::min(8, warps); cand >= 2; --cand) {
            const long long waste = (long long)((warps + cand - 1) / cand) * cand - warps;
            if (waste < bestWaste) { bestWaste = waste; wpb = cand; }
        }
        w.block = wpb * 32;
        w.tilesFine = (warps + wpb - 1) / wpb;
    }
    w.tilesCoarse = (w.nbx * w.nby + w.block - 1) / w.block;
    b->geom = GridGeom{grid->min_x, grid->min_y, grid->res, grid->nx, grid->ny, grid->pitch};

    b->descs.assign(n, CsmDesc{});
    b->stepT.assign(n, 0.0);
    b->fixups.assign(n, 0);
    b->hAngles.clear(); b->hRanges.clear();
    b->maxNT = 0; b->maxKeptPad = kBeamPad;
    long long nOff = 0, nCell = 0, nFine = 0, nCoarse = 0;
    b->workHyp = 0; b->workGather = 0;
    for (int m = 0; m < n; ++m) {
        const int b0 = scans->beam_begin[m], b1 = scans->beam_begin[m + 1];
        const int nb = b1 - b0;
        if (nb <= 0) return lgs_fail(c, LGS_ERR_INVALID, "rtcsm: scan %d has no beams", m);
        CsmDesc& d = b->descs[m];
        d.sx = scans->sensor_pose[3 * m]; d.sy = scans->sensor_pose[3 * m + 1];
        d.st = scans->sensor_pose[3 * m + 2];
        double maxR = scans->ranges[b0];
        for (int i = b0 + 1; i < b1; ++i) maxR = std::max(maxR, scans->ranges[i]);   // :163-164
        const double maxRange = std::min(maxR, p.scan_range_max);                        // :165
        const double th = grid->res / maxRange;                                          // :166
        d.stepT = std::acos(1.0 - 0.5 * th * th);                                       // :170
        b->stepT[m] = d.stepT;
        d.winT = static_cast<int>(std::ceil(0.5 * p.range_theta / d.stepT));             // :72-73
        if (!(d.stepT > 0.0) || d.winT < 0 || d.winT > (1 << 20))
            return lgs_fail(c, LGS_ERR_INVALID, "rtcsm: scan %d gives stepTheta=%g winTheta=%d", m,
                            d.stepT, d.winT);
        d.nT = 2 * d.winT + 1;
        const double thr = normThr ? normThr[m] : DBL_MIN;                               // :46-47
        d.thrAbs = thr * static_cast<double>(static_cast<size_t>(nb));                   // :77-78
        d.beamBegin = (int)b->hAngles.size();
        for (int i = b0; i < b1; ++i) {
            if (scans->ranges[i] >= p.scan_range_max) continue;                          // :192-193
            b->hAngles.push_back(scans->angles[i]);
            b->hRanges.push_back(scans->ranges[i]);
        }
        d.nKept = (int)b->hAngles.size() - d.beamBegin;
        d.nKeptPad = std::max(kBeamPad, (d.nKept + kBeamPad - 1) / kBeamPad * kBeamPad);
        d.offBegin = nOff; d.cellBegin = nCell; d.fineBegin = nFine; d.coarseBegin = nCoarse;
        nOff += (long long)d.nT * d.nKeptPad;
        nCell += (long long)d.nT * d.nKept;
        nFine += (long long)d.nT * w.nxw * w.nyw;
        nCoarse += (long long)d.nT * w.nbx * w.nby;
        b->maxNT = std::max(b->maxNT, d.nT);
        b->maxKeptPad = std::max(b->maxKeptPad, d.nKeptPad);
        const long long hyp = (long long)d.nT * ((long long)w.nxw * w.nyw + (long long)w.nbx * w.nby);
        b->workHyp += hyp;
        b->workGather += hyp * d.nKept;
    }
    if (b->maxNT > 65535)
        return lgs_fail(c, LGS_ERR_INVALID, "rtcsm: %d theta slices exceed the launch grid", b->maxNT);
    if ((size_t)b->maxKeptPad * sizeof(int) > 200 * 1024)
        return lgs_fail(c, LGS_ERR_INVALID, "rtcsm: %d beams exceed shared memory", b->maxKeptPad);
    b->nOff = nOff; b->nCell = nCell; b->nFine = nFine; b->nCoarse = nCoarse;
    if (n == 0) { b->uploaded = true; return LGS_OK; }

    const size_t nk = b->hAngles.size();
    LGS_CUDA(c, b->dDescs.reserve(n));
    LGS_CUDA(c, b->dAngles.reserve(std::max<size_t>(nk, 1)));
    LGS_CUDA(c, b->dRanges.reserve(std::max<size_t>(nk, 1)));
    LGS_CUDA(c, b->dOffs.reserve(nOff));
    LGS_CUDA(c, b->dCells.reserve(std::max<long long>(nCell, 1)));
    LGS_CUDA(c, b->dFine.reserve(nFine));
    LGS_CUDA(c, b->dCoarse.reserve(nCoarse));
    LGS_CUDA(c, b->dBlockMax.reserve(nCoarse));
    LGS_CUDA(c, b->dBlockArg.reserve(nCoarse));
    LGS_CUDA(c, b->dFlagCount.reserve(1));
    LGS_CUDA(c, b->dFlags.reserve(b->flagCap));
    LGS_CUDA(c, b->dResults.reserve(n));
    LGS_CUDA(c, b->hResults.reserve(n));
    LGS_CUDA(c, b->hFlagCount.reserve(1));
    LGS_CUDA(c, b->hDescs.reserve(n));
    LGS_CUDA(c, b->hStage.reserve(2 * std::max<size_t>(nk, 1)));
    // Stage through pinned memory so the copies are truly asynchronous.
    memcpy(b->hDescs.p, b->descs.data(), n * sizeof(CsmDesc));
    memcpy(b->hStage.p, b->hAngles.data(), nk * sizeof(double));
    memcpy(b->hStage.p + nk, b->hRanges.data(), nk * sizeof(double));
    LGS_CUDA(c, cudaMemcpyAsync(b->dDescs.p, b->hDescs.p, n * sizeof(CsmDesc),
                                cudaMemcpyHostToDevice, c->stream));
    if (nk) {
        LGS_CUDA(c, cudaMemcpyAsync(b->dAngles.p, b->hStage.p, nk * sizeof(double),
                                    cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(b->dRanges.p, b->hStage.p + nk, nk * sizeof(double),
                                    cudaMemcpyHostToDevice, c->stream));
    }
    if ((size_t)b->maxKeptPad * sizeof(int) > 48 * 1024)
    {
        LGS_CUDA(c, cudaFuncSetAttribute(csm_sweep_kernel<kBeamPad, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         b->maxKeptPad * (int)sizeof(int)));
        LGS_CUDA(c, cudaFuncSetAttribute(csm_sweep_kernel<4, 4>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         b->maxKeptPad * (int)sizeof(int)));
        LGS_CUDA(c, cudaFuncSetAttribute(csm_sweep_rows_kernel<4, 5>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         b->maxKeptPad * (int)sizeof(int)));
        LGS_CUDA(c, cudaFuncSetAttribute(csm_sweep_rows_kernel<4, 4>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         b->maxKeptPad * (int)sizeof(int)));
    }
    b->uploaded = true;
    return LGS_OK;
}

static int csm_launch_project(lgs_rtcsm_batch* b) {
    lgs_ctx* c = b->ctx;
    for (int m0 = 0; m0 < b->nMatch; m0 += 65535) {
        const int nm = std::min(65535, b->nMatch - m0);
        const long long per = (long long)b->maxNT * b->maxKeptPad;
        dim3 gridDim((unsigned)((per + 255) / 256), nm);
        csm_project_kernel<<<gridDim, 256, 0, c->stream>>>(b->dDescs.p + m0, b->dAngles.p,
                                                           b->dRanges.p, b->geom, b->win,
                                                           c->opt.edgeEps, b->dOffs.p, b->dCells.p, b->dFlags.p,
                                                           b->flagCap, b->dFlagCount.p);
        LGS_LAUNCH_CHECK(c);
    }
    return LGS_OK;
}

static int csm_run_impl(lgs_rtcsm_batch* b, const lgs_grid* grid, const lgs_grid* coarse, float* ms) {
    if (!b || !grid || !coarse) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->uploaded) return lgs_fail(c, LGS_ERR_INVALID, "rtcsm_batch_run before upload");
    if (coarse->nx != grid->nx || coarse->ny != grid->ny || coarse->pitch != grid->pitch ||
        grid->pitch != b->geom.pitch || grid->nx != b->geom.nx || grid->ny != b->geom.ny)
        return lgs_fail(c, LGS_ERR_INVALID, "rtcsm_batch_run: grid geometry mismatch");
    if (ms) ms[0] = ms[1] = ms[2] = 0.f;
    if (b->nMatch == 0) { b->ran = true; return LGS_OK; }
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, lgs_grid_acquire(c, grid));     // maps owned by another context (the builder's latest map)
    LGS_CUDA(c, lgs_grid_acquire(c, coarse));
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (ms) for (auto& e : ev) LGS_CUDA(c, cudaEventCreate(&e));
    LGS_CUDA(c, cudaMemsetAsync(b->dFlagCount.p, 0, sizeof(int), c->stream));
    if (ms) LGS_CUDA(c, cudaEventRecord(ev[0], c->stream));
    int rc = csm_launch_project(b);
    if (rc != LGS_OK) return rc;
    if (ms) LGS_CUDA(c, cudaEventRecord(ev[1], c->stream));
    rc = csm_launch_sweep(b, grid, coarse);
    if (rc != LGS_OK) return rc;
    if (ms) LGS_CUDA(c, cudaEventRecord(ev[2], c->stream));
    rc = csm_launch_select(b);
    if (rc != LGS_OK) return rc;
    if (ms) {
        LGS_CUDA(c, cudaEventRecord(ev[3], c->stream));
        LGS_CUDA(c, cudaEventSynchronize(ev[3]));
        for (int k = 0; k < 3; ++k) LGS_CUDA(c, cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]));
        for (auto& e : ev) cudaEventDestroy(e);
    }
    b->ran = true;
    return LGS_OK;
}

int lgs_rtcsm_batch_run(lgs_rtcsm_batch* b, const lgs_grid* grid, const lgs_grid* coarse) {
    return csm_run_impl(b, grid, coarse, nullptr);
}

int lgs_rtcsm_batch_run_timed(lgs_rtcsm_batch* b, const lgs_grid* grid, const lgs_grid* coarse,
                              float* ms) {
    if (!ms) return LGS_ERR_INVALID;
    return csm_run_impl(b, grid, coarse, ms);
}

int lgs_rtcsm_batch_results(lgs_rtcsm_batch* b, const lgs_grid* grid, const lgs_grid* coarse,
                            lgs_match_result* out) {
    if (!b || !grid || !coarse || (!out && b->nMatch > 0)) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->ran) return lgs_fail(c, LGS_ERR_INVALID, "rtcsm_batch_results before run");
    if (b->nMatch == 0) return LGS_OK;
    LGS_CUDA(c, cudaSetDevice(c->device));
    const CsmWindow& w = b->win;
    LGS_CUDA(c, cudaMemcpyAsync(b->hFlagCount.p, b->dFlagCount.p, sizeof(int),
                                cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaMemcpyAsync(b->hResults.p, b->dResults.p, b->nMatch * sizeof(DevResult),
                                cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    int nFlag = *b->hFlagCount.p;
    std::fill(b->fixups.begin(), b->fixups.end(), 0);
    if (nFlag > b->flagCap) {
        // More near-edge points than the list holds (only with a widened guard band): grow the list
        // and project again -- the projection is deterministic, so it now records every one of them.
        b->flagCap = nFlag + nFlag / 8;
        LGS_CUDA(c, b->dFlags.reserve(b->flagCap));
        LGS_CUDA(c, cudaMemsetAsync(b->dFlagCount.p, 0, sizeof(int), c->stream));
        const int rcp = csm_launch_project(b);
        if (rcp != LGS_OK) return rcp;
        LGS_CUDA(c, cudaMemcpyAsync(b->hFlagCount.p, b->dFlagCount.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LGS_CUDA(c, cudaStreamSynchronize(c->stream));
        nFlag = *b->hFlagCount.p;
        if (nFlag > b->flagCap)
            return lgs_fail(c, LGS_ERR_OVERFLOW, "rtcsm: %d near-edge points exceed the regrown fix-up list", nFlag);
    }
    if (nFlag > 0) {
        // Rare path: re-derive the flagged points with the host's libm (the reference's own
        // arithmetic, sensor_data.hpp:162-173 + grid_map.hpp:779-790), patch, and redo the sweep.
        std::vector<FlagEntry> fl(nFlag);
        LGS_CUDA(c, cudaMemcpy(fl.data(), b->dFlags.p, nFlag * sizeof(FlagEntry), cudaMemcpyDeviceToHost));
        std::vector<int> offVal(nFlag);
        std::vector<int2> cellVal(nFlag);
        std::vector<long long> where(2 * (size_t)nFlag);
        const int xLo = -w.winX, xHi = -w.winX + w.nxw - 1, yLo = -w.winY, yHi = -w.winY + w.nyw - 1;
        for (int k = 0; k < nFlag; ++k) {
            const CsmDesc& d = b->descs[fl[k].m];
            const int t = fl[k].t, i = fl[k].i;
            const double theta = d.st + d.stepT * static_cast<double>(t - d.winT);
            const double a = theta + b->hAngles[d.beamBegin + i];
            const double cosT = std::cos(a), sinT = std::sin(a);
            const double r = b->hRanges[d.beamBegin + i];
            const double hx = d.sx + r * cosT, hy = d.sy + r * sinT;
            const int cx = static_cast<int>(std::floor((hx - b->geom.minX) / b->geom.res));
            const int cy = static_cast<int>(std::floor((hy - b->geom.minY) / b->geom.res));
            const int ccx = std::min(std::max(cx, -xHi - 1), b->geom.nx - xLo);
            const int ccy = std::min(std::max(cy, -yHi - 1), b->geom.ny - yLo);
            offVal[k] = ccy * b->geom.pitch + ccx;
            cellVal[k] = make_int2(cx, cy);
            where[2 * k] = d.offBegin + (long long)t * d.nKeptPad + i;
            where[2 * k + 1] = d.cellBegin + (long long)t * d.nKept + i;
            b->fixups[fl[k].m]++;
        }
        LGS_CUDA(c, b->dFixOff.reserve(nFlag));
        LGS_CUDA(c, b->dFixCell.reserve(nFlag));
        LGS_CUDA(c, b->dFixWhere.reserve(2 * (size_t)nFlag));
        int* dOffVal = b->dFixOff.p; int2* dCellVal = b->dFixCell.p; long long* dWhere = b->dFixWhere.p;
        LGS_CUDA(c, cudaMemcpy(dOffVal, offVal.data(), nFlag * sizeof(int), cudaMemcpyHostToDevice));
        LGS_CUDA(c, cudaMemcpy(dCellVal, cellVal.data(), nFlag * sizeof(int2), cudaMemcpyHostToDevice));
        LGS_CUDA(c, cudaMemcpy(dWhere, where.data(), 2 * (size_t)nFlag * sizeof(long long), cudaMemcpyHostToDevice));
        csm_patch_kernel<<<(nFlag + 127) / 128, 128, 0, c->stream>>>(nullptr, dOffVal, dCellVal, dWhere,
                                                                     nFlag, b->dOffs.p, b->dCells.p);
        LGS_LAUNCH_CHECK(c);
        const int rc = csm_launch_sweep_select(b, grid, coarse);
        if (rc != LGS_OK) return rc;
        LGS_CUDA(c, cudaMemcpyAsync(b->hResults.p, b->dResults.p, b->nMatch * sizeof(DevResult),
                                    cudaMemcpyDeviceToHost, c->stream));
        LGS_CUDA(c, cudaStreamSynchronize(c->stream));
        // The patched offsets stay valid for a re-run of results(); a new run() re-projects.
        LGS_CUDA(c, cudaMemsetAsync(b->dFlagCount.p, 0, sizeof(int), c->stream));
    }
    for (int m = 0; m < b->nMatch; ++m) {
        const DevResult& r = b->hResults.p[m];
        const CsmDesc& d = b->descs[m];
        lgs_match_result& o = out[m];
        o.found = r.found; o.ix = r.ix; o.iy = r.iy; o.it = r.it;
        o.win_x = w.winX; o.win_y = w.winY; o.win_t = d.winT;
        o.n_fixups = b->fixups[m];
        o.step_x = b->geom.res; o.step_y = b->geom.res; o.step_t = d.stepT;
        o.score = r.score;
        o.n_scored = (long long)d.nT * ((long long)w.nxw * w.nyw + (long long)w.nbx * w.nby);
        o.exact_replay = r.exactReplay;
        o.reserved = 0;
    }
    return LGS_OK;
}

int lgs_rtcsm_batch_debug(lgs_rtcsm_batch* b, int m, int* dims, double* fine, double* coarse,
                          int* cells) {
    if (!b || m < 0 || m >= b->nMatch) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->ran) return lgs_fail(c, LGS_ERR_INVALID, "rtcsm_batch_debug before run");
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    const CsmDesc& d = b->descs[m];
    const CsmWindow& w = b->win;
    if (dims) { dims[0] = d.nT; dims[1] = w.nxw; dims[2] = w.nyw; dims[3] = w.nbx; dims[4] = w.nby; dims[5] = d.nKept; }
    if (fine)
        LGS_CUDA(c, cudaMemcpy(fine, b->dFine.p + d.fineBegin,
                               (size_t)d.nT * w.nxw * w.nyw * sizeof(double), cudaMemcpyDeviceToHost));
    if (coarse)
        LGS_CUDA(c, cudaMemcpy(coarse, b->dCoarse.p + d.coarseBegin,
                               (size_t)d.nT * w.nbx * w.nby * sizeof(double), cudaMemcpyDeviceToHost));
    if (cells && d.nKept > 0)
        LGS_CUDA(c, cudaMemcpy(cells, b->dCells.p + d.cellBegin,
                               (size_t)d.nT * d.nKept * sizeof(int2), cudaMemcpyDeviceToHost));
    return LGS_OK;
}

int lgs_rtcsm_batch_work(const lgs_rtcsm_batch* b, long long* hyp, long long* gathers) {
    if (!b) return LGS_ERR_INVALID;
    if (hyp) *hyp = b->workHyp;
    if (gathers) *gathers = b->workGather;
    return LGS_OK;
}

int lgs_rtcsm_match(lgs_ctx* ctx, const lgs_grid* grid, const lgs_grid* coarse,
                    const lgs_rtcsm_params* params, const lgs_scan_batch* scans,
                    const double* normThr, lgs_match_result* out) {
    lgs_rtcsm_batch* b = nullptr;
    int rc = lgs_rtcsm_batch_create(ctx, params, &b);
    if (rc != LGS_OK) return rc;
    rc = lgs_rtcsm_batch_upload(b, grid, scans, normThr);
    if (rc == LGS_OK) rc = lgs_rtcsm_batch_run(b, grid, coarse);
    if (rc == LGS_OK) rc = lgs_rtcsm_batch_results(b, grid, coarse, out);
    lgs_rtcsm_batch_destroy(b);
    return rc;
}

}  // extern "C"
