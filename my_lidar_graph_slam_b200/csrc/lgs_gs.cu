// lgs_gs.cu -- exhaustive grid-search matcher.
//
// Reference: ScanMatcherGridSearch::OptimizePose scan_matcher_grid_search.cpp:45-114 with
// ScorePixelAccurate::Score score_function_pixel_accurate.cpp:19-76.  The CPU evaluates
//   for (dy = -ry; dy <= ry; dy += sy) for (dx ...) for (dt ...)      // ACCUMULATED offsets
//       score(sensorPose + (dx, dy, dt))                               // full projection per pose
// and keeps the first strictly greater score.  Steps are free parameters (not the map resolution),
// so a hypothesis is not "base cell + integer offset" as in the correlative sweep.  But the hit
// point separates:  hit.x = (sx + dx_i) + r cos(theta_k + a),  hit.y = (sy + dy_j) + r sin(...),
// so per (theta_k, beam) there are nX candidate columns and nY candidate rows:
//   K1 gs_project_kernel  one thread per (query, theta_k, beam): sincos once, then the nX column
//                         indices and nY row offsets with the CPU's own IEEE sequence; a coordinate
//                         inside the guard band of a cell edge flags the (query, k, beam) and the host
//                         re-derives its indices with glibc (as in lgs_csm.cu);
//   K2 gs_score_kernel    one thread per hypothesis (j, i) of one theta_k: sum over beams, in beam
//                         order, of grid[row[b][j] + col[b][i]] (lanes = consecutive i, so a warp
//                         gathers a short row segment per beam, like the sweep);
//   K3 gs_select_kernel   one block per query: max score, ties to the earliest visit (y, x, theta
//                         loop order), found = max > threshold.
// The offset lists dx_i, dy_j, dt_k are produced on the host by the reference's own accumulating
// loops, so the loop lengths (which depend on the rounding of the running sums) are the CPU's.
// Unknown cells and out-of-range beams add exactly +0.0 (sums are non-negative, so x + 0.0 == x),
// which is what the CPU's `continue` does.  8 B per (hypothesis, beam) algorithmic, as for the sweep.
#include <cmath>

#include "lgs_internal.cuh"

namespace {

constexpr int kFlagCap = 1 << 16;
constexpr size_t kTableBudget = size_t(3) << 30;     // bytes of index tables per device chunk

struct GsQuery {
    double sx, sy, st;              // sensor pose
    double minRange, maxRange;      // usable range combined with the scan's own limits
    double minX, minY, res;
    const double* origin;           // grid cell (0, 0)
    int nx, ny, pitch, offX, offY;
    int beamBegin, nBeams;
    long long colBegin, rowBegin;   // into the index tables: [k][beam][i], [k][beam][j]
    long long scoreBegin;           // [k][j][i]
    double threshold;
};

struct GsFlag { int q, k, b; };

struct GsResult { double score; int found, ix, iy, it; };

__global__ void gs_project_kernel(const GsQuery* __restrict__ queries, const double* __restrict__ angles,
                                  const double* __restrict__ ranges, const double* __restrict__ dX,
                                  const double* __restrict__ dY, const double* __restrict__ dT, int nX, int nY,
                                  int nT, double eps, int* __restrict__ col, int* __restrict__ row,
                                  GsFlag* __restrict__ flags, int* __restrict__ flagCount) {
    const GsQuery q = queries[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nT * q.nBeams) return;
    const int k = idx / q.nBeams, b = idx - k * q.nBeams;
    int* c = col + q.colBegin + ((long long)k * q.nBeams + b) * nX;
    int* rw = row + q.rowBegin + ((long long)k * q.nBeams + b) * nY;
    const double r = ranges[q.beamBegin + b];
    if (r >= q.maxRange || r <= q.minRange) {          // score_function_pixel_accurate.cpp:38-39
        for (int i = 0; i < nX; ++i) c[i] = -1;        // (-1, -1) is an apron cell: adds +0.0
        for (int j = 0; j < nY; ++j) rw[j] = -q.pitch;
        return;
    }
    const double theta = __dadd_rn(q.st, dT[k]);       // scan_matcher_grid_search.cpp:78-80
    double s, cs;
    sincos(__dadd_rn(theta, angles[q.beamBegin + b]), &s, &cs);   // sensor_data.hpp:168-169
    const double rc = __dmul_rn(r, cs), rs = __dmul_rn(r, s);
    bool edge = false;
    for (int i = 0; i < nX; ++i) {
        const double hx = __dadd_rn(__dadd_rn(q.sx, dX[i]), rc);
        const double qx = __ddiv_rn(__dsub_rn(hx, q.minX), q.res);          // grid_map.hpp:779-790
        const double f = qx - floor(qx);
        edge |= !(f >= eps && f <= 1.0 - eps);
        const int cx = __double2int_rd(qx) - q.offX;
        c[i] = min(max(cx, -1), q.nx);
    }
    for (int j = 0; j < nY; ++j) {
        const double hy = __dadd_rn(__dadd_rn(q.sy, dY[j]), rs);
        const double qy = __ddiv_rn(__dsub_rn(hy, q.minY), q.res);
        const double f = qy - floor(qy);
        edge |= !(f >= eps && f <= 1.0 - eps);
        const int cy = __double2int_rd(qy) - q.offY;
        rw[j] = min(max(cy, -1), q.ny) * q.pitch;
    }
    if (edge) {
        const int n = atomicAdd(flagCount, 1);
        if (n < kFlagCap) flags[n] = GsFlag{(int)blockIdx.y, k, b};
    }
}

// grid = (ceil(nX * nY / 256), nT, queries)
__global__ void __launch_bounds__(256)
gs_score_kernel(const GsQuery* __restrict__ queries, const int* __restrict__ col, const int* __restrict__ row,
                int nX, int nY, double* __restrict__ scores) {
    const GsQuery q = queries[blockIdx.z];
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nX * nY) return;
    const int k = blockIdx.y;
    const int j = h / nX, i = h - j * nX;
    const int* c = col + q.colBegin + (long long)k * q.nBeams * nX + i;
    const int* rw = row + q.rowBegin + (long long)k * q.nBeams * nY + j;
    const double* g = q.origin;
    double sum = 0.0;
    int b = 0;
    for (; b + 4 <= q.nBeams; b += 4) {                 // gathers of four beams in flight, adds in order
        const double v0 = __ldg(g + (rw[0] + c[0]));
        const double v1 = __ldg(g + (rw[nY] + c[nX]));
        const double v2 = __ldg(g + (rw[2 * nY] + c[2 * nX]));
        const double v3 = __ldg(g + (rw[3 * nY] + c[3 * nX]));
        sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(sum, v0), v1), v2), v3);
        c += 4 * nX; rw += 4 * nY;
    }
    for (; b < q.nBeams; ++b) {
        sum = __dadd_rn(sum, __ldg(g + (rw[0] + c[0])));
        c += nX; rw += nY;
    }
    scores[q.scoreBegin + ((long long)k * nY + j) * nX + i] = sum;
}

__global__ void __launch_bounds__(256)
gs_select_kernel(const GsQuery* __restrict__ queries, const double* __restrict__ scores, int nX, int nY,
                 int nT, GsResult* __restrict__ results) {
    const GsQuery q = queries[blockIdx.x];
    const long long total = (long long)nT * nY * nX;
    double best = -1.0;
    long long bestVisit = 0x7fffffffffffffffLL;
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
        const int k = (int)(e / ((long long)nY * nX));
        const int ji = (int)(e - (long long)k * nY * nX);      // j * nX + i
        const long long visit = (long long)ji * nT + k;        // loop order y, x, theta (:74-76)
        const double s = scores[q.scoreBegin + e];
        if (s > best || (s == best && visit < bestVisit)) { best = s; bestVisit = visit; }
    }
    __shared__ double sBest[256];
    __shared__ long long sVisit[256];
    sBest[threadIdx.x] = best; sVisit[threadIdx.x] = bestVisit;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const double s = sBest[threadIdx.x + w];
            const long long v = sVisit[threadIdx.x + w];
            if (s > sBest[threadIdx.x] || (s == sBest[threadIdx.x] && v < sVisit[threadIdx.x])) {
                sBest[threadIdx.x] = s; sVisit[threadIdx.x] = v;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        GsResult r;
        r.found = total > 0 && sBest[0] > q.threshold;   // strict, first visit among equals (:85, :94)
        r.score = r.found ? sBest[0] : q.threshold;
        const long long v = r.found ? sVisit[0] : 0;
        const int ji = (int)(v / nT);
        r.it = r.found ? (int)(v - (long long)ji * nT) : -1;
        r.iy = r.found ? ji / nX : -1;
        r.ix = r.found ? ji - (ji / nX) * nX : -1;
        results[blockIdx.x] = r;
    }
}

// The reference's accumulating loop (scan_matcher_grid_search.cpp:74-76): values of d with
// `for (d = -r; d <= r; d += s)`.
std::vector<double> offsets(double range, double step) {
    std::vector<double> v;
    for (double d = -range; d <= range; d += step) {
        v.push_back(d);
        if (v.size() > (1u << 20)) break;
    }
    return v;
}

inline int worldToCell(double p, double minP, double res) {
    return static_cast<int>(std::floor((p - minP) / res));                  // grid_map.hpp:779-790
}

}  // namespace

extern "C" {

int lgs_gs_offsets(double range, double step, double* out, int cap, int* n) {
    if (!(step > 0.0) || !(range >= 0.0) || !n) return LGS_ERR_INVALID;
    const std::vector<double> v = offsets(range / 2.0, step);               // :59-61: radius = range / 2
    *n = (int)v.size();
    if (out) std::memcpy(out, v.data(), std::min<size_t>(v.size(), cap > 0 ? cap : 0) * sizeof(double));
    return LGS_OK;
}

int lgs_gs_match(lgs_ctx* c, const lgs_gs_params* p, const lgs_scan_batch* scans,
                 const lgs_grid* const* grids, const double* normThreshold, lgs_match_result* out,
                 double* scoreTable) {
    if (!c || !p || !scans || scans->n_scans < 0) return LGS_ERR_INVALID;
    const int nQ = scans->n_scans;
    if (nQ == 0) return LGS_OK;
    if (!grids || !out || !scans->beam_begin || !scans->sensor_pose) return LGS_ERR_INVALID;
    if (!(p->step_x > 0.0) || !(p->step_y > 0.0) || !(p->step_theta > 0.0) || !(p->range_x >= 0.0) ||
        !(p->range_y >= 0.0) || !(p->range_theta >= 0.0))
        return lgs_fail(c, LGS_ERR_INVALID, "gs_match: ranges must be >= 0 and steps > 0");
    const std::vector<double> dX = offsets(p->range_x / 2.0, p->step_x);
    const std::vector<double> dY = offsets(p->range_y / 2.0, p->step_y);
    const std::vector<double> dT = offsets(p->range_theta / 2.0, p->step_theta);
    const int nX = (int)dX.size(), nY = (int)dY.size(), nT = (int)dT.size();
    if ((long long)nX * nY * nT > (1LL << 31) || nX > (1 << 20) || nY > (1 << 20) || nT > (1 << 20))
        return lgs_fail(c, LGS_ERR_INVALID, "gs_match: %d x %d x %d hypotheses per query", nX, nY, nT);
    if (scoreTable && nQ != 1) return lgs_fail(c, LGS_ERR_INVALID, "gs_match: the score table is a 1-query diagnostic");
    LGS_CUDA(c, cudaSetDevice(c->device));
    const double eps = g_lgs_edge_eps;
    const size_t nBeamsAll = (size_t)scans->beam_begin[nQ];

    double *dAngles = nullptr, *dRanges = nullptr, *dOff = nullptr;
    LGS_CUDA(c, cudaMallocAsync(&dAngles, std::max<size_t>(nBeamsAll, 1) * sizeof(double), c->stream));
    LGS_CUDA(c, cudaMallocAsync(&dRanges, std::max<size_t>(nBeamsAll, 1) * sizeof(double), c->stream));
    LGS_CUDA(c, cudaMallocAsync(&dOff, (size_t)(nX + nY + nT + 1) * sizeof(double), c->stream));
    if (nBeamsAll) {
        LGS_CUDA(c, cudaMemcpyAsync(dAngles, scans->angles, nBeamsAll * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(dRanges, scans->ranges, nBeamsAll * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    std::vector<double> hOff;
    hOff.insert(hOff.end(), dX.begin(), dX.end());
    hOff.insert(hOff.end(), dY.begin(), dY.end());
    hOff.insert(hOff.end(), dT.begin(), dT.end());
    if (!hOff.empty())
        LGS_CUDA(c, cudaMemcpyAsync(dOff, hOff.data(), hOff.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const double *ddX = dOff, *ddY = dOff + nX, *ddT = dOff + nX + nY;

    int rcOut = LGS_OK;
    int q0 = 0;
    while (q0 < nQ && rcOut == LGS_OK) {
        // chunk of queries whose index tables fit the budget
        std::vector<GsQuery> hq;
        long long colCells = 0, rowCells = 0, scoreCells = 0;
        int maxBeams = 0;
        int q1 = q0;
        for (; q1 < nQ; ++q1) {
            const lgs_grid* g = grids[q1];
            if (!g) { rcOut = lgs_fail(c, LGS_ERR_INVALID, "gs_match: query %d has no grid", q1); break; }
            if (g->ctx->device != c->device) { rcOut = lgs_fail(c, LGS_ERR_INVALID, "gs_match: grid of query %d on another device", q1); break; }
            const int nb = scans->beam_begin[q1 + 1] - scans->beam_begin[q1];
            const long long cc = (long long)nT * nb * nX, rr = (long long)nT * nb * nY;
            if (q1 > q0 && (size_t)(colCells + rowCells + cc + rr) * sizeof(int) > kTableBudget) break;
            GsQuery h;
            h.sx = scans->sensor_pose[3 * q1]; h.sy = scans->sensor_pose[3 * q1 + 1]; h.st = scans->sensor_pose[3 * q1 + 2];
            const double scanMin = scans->range_min ? scans->range_min[q1] : 0.0;
            const double scanMax = scans->range_max ? scans->range_max[q1] : INFINITY;
            h.minRange = std::max(p->score_range_min, scanMin);              // score_function_pixel_accurate.cpp:27-30
            h.maxRange = std::min(p->score_range_max, scanMax);
            h.minX = g->min_x; h.minY = g->min_y; h.res = g->res; h.origin = g->origin();
            h.nx = g->nx; h.ny = g->ny; h.pitch = g->pitch; h.offX = g->off_x; h.offY = g->off_y;
            h.beamBegin = scans->beam_begin[q1]; h.nBeams = nb;
            h.colBegin = colCells; h.rowBegin = rowCells; h.scoreBegin = scoreCells;
            h.threshold = (normThreshold ? normThreshold[q1] : 2.2250738585072014e-308) * (double)(size_t)nb;   // :67-68
            colCells += cc; rowCells += rr; scoreCells += (long long)nT * nY * nX;
            maxBeams = std::max(maxBeams, nb);
            hq.push_back(h);
            if (g->ctx != c) cudaStreamSynchronize(g->ctx->stream);          // maps produced on another context
        }
        if (rcOut != LGS_OK) break;
        const int nq = (int)hq.size();
        GsQuery* dQ = nullptr; int *dCol = nullptr, *dRow = nullptr, *dFlagCount = nullptr;
        double* dScores = nullptr; GsFlag* dFlags = nullptr; GsResult* dRes = nullptr;
        auto freeAll = [&]() {
            cudaFreeAsync(dQ, c->stream); cudaFreeAsync(dCol, c->stream); cudaFreeAsync(dRow, c->stream);
            cudaFreeAsync(dScores, c->stream); cudaFreeAsync(dFlags, c->stream);
            cudaFreeAsync(dFlagCount, c->stream); cudaFreeAsync(dRes, c->stream);
        };
#define GS_TRY(call)                                                                               \
        do { cudaError_t e__ = (call);                                                             \
             if (e__ != cudaSuccess) { freeAll();                                                  \
                 rcOut = lgs_fail(c, e__ == cudaErrorMemoryAllocation ? LGS_ERR_NOMEM : LGS_ERR_CUDA, \
                                  "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); } } while (0)
        GS_TRY(cudaMallocAsync(&dQ, nq * sizeof(GsQuery), c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMallocAsync(&dCol, std::max<long long>(colCells, 1) * sizeof(int), c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMallocAsync(&dRow, std::max<long long>(rowCells, 1) * sizeof(int), c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMallocAsync(&dScores, std::max<long long>(scoreCells, 1) * sizeof(double), c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMallocAsync(&dFlags, kFlagCap * sizeof(GsFlag), c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMallocAsync(&dFlagCount, sizeof(int), c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMallocAsync(&dRes, nq * sizeof(GsResult), c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMemcpyAsync(dQ, hq.data(), nq * sizeof(GsQuery), cudaMemcpyHostToDevice, c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMemsetAsync(dFlagCount, 0, sizeof(int), c->stream));
        if (rcOut != LGS_OK) break;

        const long long hyp = (long long)nX * nY;
        if (maxBeams > 0 && nT > 0) {
            dim3 gp((unsigned)(((long long)nT * maxBeams + 127) / 128), nq);
            gs_project_kernel<<<gp, 128, 0, c->stream>>>(dQ, dAngles, dRanges, ddX, ddY, ddT, nX, nY, nT, eps,
                                                         dCol, dRow, dFlags, dFlagCount);
            c->launches++;
        }
        int nFlag = 0;
        GS_TRY(cudaMemcpyAsync(&nFlag, dFlagCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaStreamSynchronize(c->stream));
        if (rcOut != LGS_OK) break;
        if (nFlag > kFlagCap) { freeAll(); rcOut = lgs_fail(c, LGS_ERR_OVERFLOW, "gs_match: %d near-edge beams exceed the fix-up list", nFlag); break; }
        std::vector<int> fixups(nq, 0);
        if (nFlag > 0) {
            // Rare path: the flagged beams' columns and rows with the host's libm, the reference's
            // own arithmetic (sensor_data.hpp:162-173 + grid_map.hpp:779-790)
            std::vector<GsFlag> fl(nFlag);
            GS_TRY(cudaMemcpy(fl.data(), dFlags, nFlag * sizeof(GsFlag), cudaMemcpyDeviceToHost));
            std::vector<int> hc(nX), hr(nY);
            for (int f = 0; f < nFlag && rcOut == LGS_OK; ++f) {
                const GsQuery& h = hq[fl[f].q];
                const int k = fl[f].k, b = fl[f].b;
                const double r = scans->ranges[h.beamBegin + b];
                const double theta = h.st + dT[k];
                const double cosT = std::cos(theta + scans->angles[h.beamBegin + b]);
                const double sinT = std::sin(theta + scans->angles[h.beamBegin + b]);
                for (int i = 0; i < nX; ++i) {
                    const double hx = (h.sx + dX[i]) + r * cosT;
                    hc[i] = std::min(std::max(worldToCell(hx, h.minX, h.res) - h.offX, -1), h.nx);
                }
                for (int j = 0; j < nY; ++j) {
                    const double hy = (h.sy + dY[j]) + r * sinT;
                    hr[j] = std::min(std::max(worldToCell(hy, h.minY, h.res) - h.offY, -1), h.ny) * h.pitch;
                }
                GS_TRY(cudaMemcpy(dCol + h.colBegin + ((long long)k * h.nBeams + b) * nX, hc.data(), nX * sizeof(int), cudaMemcpyHostToDevice));
                if (rcOut == LGS_OK)
                    GS_TRY(cudaMemcpy(dRow + h.rowBegin + ((long long)k * h.nBeams + b) * nY, hr.data(), nY * sizeof(int), cudaMemcpyHostToDevice));
                fixups[fl[f].q]++;
            }
            if (rcOut != LGS_OK) break;
        }
        if (hyp > 0 && nT > 0) {
            dim3 gs((unsigned)((hyp + 255) / 256), nT, nq);
            gs_score_kernel<<<gs, 256, 0, c->stream>>>(dQ, dCol, dRow, nX, nY, dScores);
            c->launches++;
        }
        gs_select_kernel<<<nq, 256, 0, c->stream>>>(dQ, dScores, nX, nY, nT, dRes);
        c->launches++;
        GS_TRY(cudaGetLastError());
        std::vector<GsResult> hr(nq);
        if (rcOut == LGS_OK) GS_TRY(cudaMemcpyAsync(hr.data(), dRes, nq * sizeof(GsResult), cudaMemcpyDeviceToHost, c->stream));
        if (rcOut == LGS_OK && scoreTable && scoreCells > 0)
            GS_TRY(cudaMemcpyAsync(scoreTable, dScores, scoreCells * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaStreamSynchronize(c->stream));
        if (rcOut != LGS_OK) break;
        freeAll();
#undef GS_TRY
        for (int k = 0; k < nq; ++k) {
            lgs_match_result& o = out[q0 + k];
            std::memset(&o, 0, sizeof(o));
            o.found = hr[k].found; o.ix = hr[k].ix; o.iy = hr[k].iy; o.it = hr[k].it;
            o.win_x = nX; o.win_y = nY; o.win_t = nT;      // loop lengths (not half sizes) for this matcher
            o.n_fixups = fixups[k];
            o.step_x = p->step_x; o.step_y = p->step_y; o.step_t = p->step_theta;
            o.score = hr[k].score;
            o.n_scored = (long long)nX * nY * nT;
        }
        q0 = q1;
    }
    cudaFreeAsync(dAngles, c->stream); cudaFreeAsync(dRanges, c->stream); cudaFreeAsync(dOff, c->stream);
    cudaStreamSynchronize(c->stream);
    return rcOut;
}

}  // extern "C"
