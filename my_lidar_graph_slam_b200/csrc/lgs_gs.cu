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
// When the hypotheses of one theta fit one block (the launcher's 41 x 41 window does), K1 and K2 run
// fused (gs_fused_kernel): the index runs of 64 beams at a time live in shared memory only, nothing
// but the block maxima reaches HBM; a near-edge beam sends the chunk through the table path above.
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

// Block = 128 consecutive (theta_k, beam) entries of one query, whose column / row index runs are
// contiguous in the tables: with STAGED the block builds them in shared memory and writes them out
// with coalesced stores (a thread's own run is nX ints, so direct stores scatter over 128 lines).
template <bool STAGED>
__global__ void __launch_bounds__(128)
gs_project_kernel(const GsQuery* __restrict__ queries, const double* __restrict__ angles,
                  const double* __restrict__ ranges, const double* __restrict__ dX,
                  const double* __restrict__ dY, const double* __restrict__ dT, int nX, int nY,
                  int nYp, int nT, double eps, int* __restrict__ col, int* __restrict__ row,
                  GsFlag* __restrict__ flags, int* __restrict__ flagCount) {
    extern __shared__ int sh[];
    const GsQuery q = queries[blockIdx.y];
    const int total = nT * q.nBeams;
    const int base = blockIdx.x * blockDim.x;
    if (base >= total) return;                         // whole block
    const int idx = base + threadIdx.x;
    const bool live = idx < total;
    const int k = live ? idx / q.nBeams : 0, b = live ? idx - k * q.nBeams : 0;
    int* c = STAGED ? sh + threadIdx.x * nX : col + q.colBegin + (long long)idx * nX;
    int* rw = STAGED ? sh + blockDim.x * nX + threadIdx.x * nYp   // rows padded to a multiple of 4
                     : row + q.rowBegin + (long long)idx * nYp;
    const double r = live ? ranges[q.beamBegin + b] : 0.0;
    if (live && (r >= q.maxRange || r <= q.minRange)) {          // score_function_pixel_accurate.cpp:38-39
        for (int i = 0; i < nX; ++i) c[i] = -1;        // (-1, -1) is an apron cell: adds +0.0
        for (int j = 0; j < nYp; ++j) rw[j] = -q.pitch;
    } else if (live) {
        const double theta = __dadd_rn(q.st, dT[k]);   // scan_matcher_grid_search.cpp:78-80
        double s, cs;
        sincos(__dadd_rn(theta, angles[q.beamBegin + b]), &s, &cs);   // sensor_data.hpp:168-169
        const double rc = __dmul_rn(r, cs), rs = __dmul_rn(r, s);
        // floor((h - min) / res), grid_map.hpp:779-790.  The quotient is first estimated with a
        // multiply (relative error a few ulp, i.e. < 1e-9 cells for |q| < 1e6); only an estimate
        // within eps + 1e-6 of a cell edge -- or a huge one -- pays for the IEEE division that decides
        // the floor and the guard band.
        const double invRes = __ddiv_rn(1.0, q.res);
        const double fast = eps + 1e-6;                // the estimate's band contains the guard band
        bool edge = false;
        for (int i = 0; i < nX; ++i) {
            const double num = __dsub_rn(__dadd_rn(__dadd_rn(q.sx, dX[i]), rc), q.minX);
            double qx = __dmul_rn(num, invRes);
            double f = qx - floor(qx);
            if (!(f >= fast && f <= 1.0 - fast && fabs(qx) < 1e6)) {
                qx = __ddiv_rn(num, q.res);
                f = qx - floor(qx);
                edge |= !(f >= eps && f <= 1.0 - eps);
            }
            const int cx = __double2int_rd(qx) - q.offX;
            c[i] = min(max(cx, -1), q.nx);
        }
        for (int j = 0; j < nY; ++j) {
            const double num = __dsub_rn(__dadd_rn(__dadd_rn(q.sy, dY[j]), rs), q.minY);
            double qy = __dmul_rn(num, invRes);
            double f = qy - floor(qy);
            if (!(f >= fast && f <= 1.0 - fast && fabs(qy) < 1e6)) {
                qy = __ddiv_rn(num, q.res);
                f = qy - floor(qy);
                edge |= !(f >= eps && f <= 1.0 - eps);
            }
            const int cy = __double2int_rd(qy) - q.offY;
            rw[j] = min(max(cy, -1), q.ny) * q.pitch;
        }
        for (int j = nY; j < nYp; ++j) rw[j] = -q.pitch;   // padding: never stored, must stay readable
        if (edge) {
            const int n = atomicAdd(flagCount, 1);
            if (n < kFlagCap) flags[n] = GsFlag{(int)blockIdx.y, k, b};
        }
    }
    if (STAGED) {
        __syncthreads();
        const int n = min((int)blockDim.x, total - base);
        int* gc = col + q.colBegin + (long long)base * nX;
        int* gr = row + q.rowBegin + (long long)base * nYp;
        for (int e = threadIdx.x; e < n * nX; e += blockDim.x) gc[e] = sh[e];
        for (int e = threadIdx.x; e < n * nYp; e += blockDim.x) gr[e] = sh[blockDim.x * nX + e];
    }
}

constexpr int kRows = 4;          // hypotheses (consecutive y offsets) per thread in the score kernel
static_assert(kRows == 4, "the score kernel loads a thread's row offsets as one int4");

struct GsPartial { double score; long long visit; };

__device__ __forceinline__ bool gsBetter(double s, long long v, double bs, long long bv) {
    return s > bs || (s == bs && v < bv);           // strict ">" in visit order (:85)
}

// grid = (ceil(nX * ceil(nY / kRows) / 256), nT, queries).  A thread owns column offset i and kRows
// consecutive row offsets: one column-index load serves kRows gathers, and kRows x 4 gathers are in
// flight per thread.  Each block leaves its best (score, visit) in `partials`; the full table is
// written only when a caller asks for it.
__global__ void __launch_bounds__(256)
gs_score_kernel(const GsQuery* __restrict__ queries, const int* __restrict__ col, const int* __restrict__ row,
                int nX, int nY, int nYp, int nT, double* __restrict__ scores, GsPartial* __restrict__ partials) {
    const GsQuery* q = queries + blockIdx.z;
    const int nBeams = q->nBeams;
    const int groups = (nY + kRows - 1) / kRows;
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = h < nX * groups;
    const int k = blockIdx.y;
    const int jg = active ? h / nX : 0, i = active ? h - jg * nX : 0;
    const int j0 = jg * kRows;
    const int* c = col + q->colBegin + (long long)k * nBeams * nX + i;
    // the kRows = 4 row offsets of a beam are one aligned 16-byte load (rows are padded to nYp)
    const int4* rw = reinterpret_cast<const int4*>(row + q->rowBegin + (long long)k * nBeams * nYp + j0);
    const int rowStride = nYp / 4;
    const double* g = q->origin;
    double sum[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) sum[r] = 0.0;
    int b = 0;
    if (active) {
        for (; b + 4 <= nBeams; b += 4) {               // 4 beams x kRows gathers in flight, adds in beam order
            double v[4][kRows];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ci = c[u * nX];
                const int4 ro = rw[u * rowStride];
                v[u][0] = __ldg(g + (ro.x + ci)); v[u][1] = __ldg(g + (ro.y + ci));
                v[u][2] = __ldg(g + (ro.z + ci)); v[u][3] = __ldg(g + (ro.w + ci));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int r = 0; r < kRows; ++r) sum[r] = __dadd_rn(sum[r], v[u][r]);
            c += 4 * nX; rw += 4 * rowStride;
        }
        for (; b < nBeams; ++b) {
            const int ci = c[0];
            const int4 ro = rw[0];
            sum[0] = __dadd_rn(sum[0], __ldg(g + (ro.x + ci))); sum[1] = __dadd_rn(sum[1], __ldg(g + (ro.y + ci)));
            sum[2] = __dadd_rn(sum[2], __ldg(g + (ro.z + ci))); sum[3] = __dadd_rn(sum[3], __ldg(g + (ro.w + ci)));
            c += nX; rw += rowStride;
        }
    }
    double best = -1.0;
    long long bestVisit = 0x7fffffffffffffffLL;
    if (active) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int j = j0 + r;
            if (j >= nY) break;
            if (scores) scores[q->scoreBegin + ((long long)k * nY + j) * nX + i] = sum[r];
            const long long visit = ((long long)j * nX + i) * nT + k;      // loop order y, x, theta (:74-76)
            if (gsBetter(sum[r], visit, best, bestVisit)) { best = sum[r]; bestVisit = visit; }
        }
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        const double s = __shfl_xor_sync(0xffffffffu, best, w);
        const long long v = __shfl_xor_sync(0xffffffffu, bestVisit, w);
        if (gsBetter(s, v, best, bestVisit)) { best = s; bestVisit = v; }
    }
    __shared__ double sBest[8];
    __shared__ long long sVisit[8];
    if ((threadIdx.x & 31) == 0) { sBest[threadIdx.x >> 5] = best; sVisit[threadIdx.x >> 5] = bestVisit; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (gsBetter(sBest[w], sVisit[w], best, bestVisit)) { best = sBest[w]; bestVisit = sVisit[w]; }
        partials[((long long)blockIdx.z * nT + k) * gridDim.x + blockIdx.x] = GsPartial{best, bestVisit};
    }
}

// Fused form for windows whose hypotheses of one theta fit one block (nX * ceil(nY / 4) <= THREADS,
// e.g. the launcher's 41 x 41): block = (theta_k, query).  The column / row index runs of kFusedBeams
// beams at a time are built in shared memory and consumed from there, so no index table ever reaches
// HBM and there is no separate projection launch.  A beam inside the guard band only raises
// `flagCount`; the host then redoes the chunk through the table path, which has the exact fix-up.
constexpr int kFusedBeams = 64;

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
gs_fused_kernel(const GsQuery* __restrict__ queries, const double* __restrict__ angles,
                const double* __restrict__ ranges, const double* __restrict__ dX,
                const double* __restrict__ dY, const double* __restrict__ dT, int nX, int nY, int nYp,
                int nT, double eps, double* __restrict__ scores, GsPartial* __restrict__ partials,
                int* __restrict__ flagCount) {
    extern __shared__ __align__(16) unsigned char shRaw[];
    double* sRc = reinterpret_cast<double*>(shRaw);                    // r cos, NaN = beam out of range
    double* sRs = sRc + kFusedBeams;
    int* sRow = reinterpret_cast<int*>(sRs + kFusedBeams);            // [beam][nYp], 16-byte aligned rows
    int* sCol = sRow + kFusedBeams * nYp;                             // [beam][nX]
    __shared__ int sEdge;
    const GsQuery* q = queries + blockIdx.y;
    const int k = blockIdx.x;
    const int nBeams = q->nBeams;
    const int groups = (nY + kRows - 1) / kRows;
    const bool active = (int)threadIdx.x < nX * groups;
    const int jg = active ? threadIdx.x / nX : 0, i = active ? threadIdx.x - jg * nX : 0;
    const int j0 = jg * kRows;
    const double sx = q->sx, sy = q->sy, minX = q->minX, minY = q->minY, res = q->res;
    const double theta = __dadd_rn(q->st, dT[k]);                     // scan_matcher_grid_search.cpp:78-80
    const double invRes = __ddiv_rn(1.0, res);
    const double fast = eps + 1e-6;
    const int gnx = q->nx, gny = q->ny, pitch = q->pitch, offX = q->offX, offY = q->offY;
    const double* g = q->origin;
    if (threadIdx.x == 0) sEdge = 0;
    double sum[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) sum[r] = 0.0;
    const int per = nX + nYp;
    for (int c0 = 0; c0 < nBeams; c0 += kFusedBeams) {
        const int nb = min(kFusedBeams, nBeams - c0);
        __syncthreads();                                               // previous chunk consumed
        if ((int)threadIdx.x < nb) {
            const int b = q->beamBegin + c0 + threadIdx.x;
            const double r = ranges[b];
            double rc = nan(""), rs = 0.0;
            if (!(r >= q->maxRange || r <= q->minRange)) {             // score_function_pixel_accurate.cpp:38-39
                double s, cs;
                sincos(__dadd_rn(theta, angles[b]), &s, &cs);          // sensor_data.hpp:168-169
                rc = __dmul_rn(r, cs); rs = __dmul_rn(r, s);
            }
            sRc[threadIdx.x] = rc; sRs[threadIdx.x] = rs;
        }
        __syncthreads();
        bool edge = false;
        for (int e = threadIdx.x; e < nb * per; e += THREADS) {
            const int t = e / per, o = e - t * per;
            const double rc = sRc[t];
            if (o < nX) {
                int v = -1;                                            // (-1, -1) is an apron cell: adds +0.0
                if (rc == rc) {
                    const double num = __dsub_rn(__dadd_rn(__dadd_rn(sx, dX[o]), rc), minX);
                    double qx = __dmul_rn(num, invRes);                // estimate; see gs_project_kernel
                    double f = qx - floor(qx);
                    if (!(f >= fast && f <= 1.0 - fast && fabs(qx) < 1e6)) {
                        qx = __ddiv_rn(num, res);                      // grid_map.hpp:779-790
                        f = qx - floor(qx);
                        edge |= !(f >= eps && f <= 1.0 - eps);
                    }
                    v = min(max(__double2int_rd(qx) - offX, -1), gnx);
                }
                sCol[t * nX + o] = v;
            } else {
                const int j = o - nX;
                int v = -pitch;
                if (rc == rc && j < nY) {
                    const double num = __dsub_rn(__dadd_rn(__dadd_rn(sy, dY[j]), sRs[t]), minY);
                    double qy = __dmul_rn(num, invRes);
                    double f = qy - floor(qy);
                    if (!(f >= fast && f <= 1.0 - fast && fabs(qy) < 1e6)) {
                        qy = __ddiv_rn(num, res);
                        f = qy - floor(qy);
                        edge |= !(f >= eps && f <= 1.0 - eps);
                    }
                    v = min(max(__double2int_rd(qy) - offY, -1), gny) * pitch;
                }
                sRow[t * nYp + j] = v;
            }
        }
        if (edge) sEdge = 1;
        __syncthreads();
        if (active) {
            const int* c = sCol + i;
            const int4* rw = reinterpret_cast<const int4*>(sRow + j0);
            const int rowStride = nYp / 4;
            int t = 0;
            for (; t + 4 <= nb; t += 4) {                              // 16 gathers in flight, adds in beam order
                double v[4][kRows];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ci = c[(t + u) * nX];
                    const int4 ro = rw[(t + u) * rowStride];
                    v[u][0] = __ldg(g + (ro.x + ci)); v[u][1] = __ldg(g + (ro.y + ci));
                    v[u][2] = __ldg(g + (ro.z + ci)); v[u][3] = __ldg(g + (ro.w + ci));
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int r = 0; r < kRows; ++r) sum[r] = __dadd_rn(sum[r], v[u][r]);
            }
            for (; t < nb; ++t) {
                const int ci = c[t * nX];
                const int4 ro = rw[t * rowStride];
                sum[0] = __dadd_rn(sum[0], __ldg(g + (ro.x + ci))); sum[1] = __dadd_rn(sum[1], __ldg(g + (ro.y + ci)));
                sum[2] = __dadd_rn(sum[2], __ldg(g + (ro.z + ci))); sum[3] = __dadd_rn(sum[3], __ldg(g + (ro.w + ci)));
            }
        }
    }
    double best = -1.0;
    long long bestVisit = 0x7fffffffffffffffLL;
    if (active) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const int j = j0 + r;
            if (j >= nY) break;
            if (scores) scores[q->scoreBegin + ((long long)k * nY + j) * nX + i] = sum[r];
            const long long visit = ((long long)j * nX + i) * nT + k;      // loop order y, x, theta (:74-76)
            if (gsBetter(sum[r], visit, best, bestVisit)) { best = sum[r]; bestVisit = visit; }
        }
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        const double s = __shfl_xor_sync(0xffffffffu, best, w);
        const long long v = __shfl_xor_sync(0xffffffffu, bestVisit, w);
        if (gsBetter(s, v, best, bestVisit)) { best = s; bestVisit = v; }
    }
    __shared__ double sBest[THREADS / 32];
    __shared__ long long sVisit[THREADS / 32];
    if ((threadIdx.x & 31) == 0) { sBest[threadIdx.x >> 5] = best; sVisit[threadIdx.x >> 5] = bestVisit; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < THREADS / 32; ++w)
            if (gsBetter(sBest[w], sVisit[w], best, bestVisit)) { best = sBest[w]; bestVisit = sVisit[w]; }
        partials[(long long)blockIdx.y * nT + k] = GsPartial{best, bestVisit};
        if (sEdge) atomicAdd(flagCount, 1);
    }
}

// one block per query over its nT x tiles block partials
__global__ void __launch_bounds__(256)
gs_select_kernel(const GsQuery* __restrict__ queries, const GsPartial* __restrict__ partials, int perQuery,
                 int nX, int nT, GsResult* __restrict__ results) {
    const double threshold = queries[blockIdx.x].threshold;
    double best = -1.0;
    long long bestVisit = 0x7fffffffffffffffLL;
    for (int e = threadIdx.x; e < perQuery; e += blockDim.x) {
        const GsPartial p = partials[(long long)blockIdx.x * perQuery + e];
        if (gsBetter(p.score, p.visit, best, bestVisit)) { best = p.score; bestVisit = p.visit; }
    }
    __shared__ double sBest[256];
    __shared__ long long sVisit[256];
    sBest[threadIdx.x] = best; sVisit[threadIdx.x] = bestVisit;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w && gsBetter(sBest[threadIdx.x + w], sVisit[threadIdx.x + w], sBest[threadIdx.x],
                                        sVisit[threadIdx.x])) {
            sBest[threadIdx.x] = sBest[threadIdx.x + w]; sVisit[threadIdx.x] = sVisit[threadIdx.x + w];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        GsResult r;
        r.found = perQuery > 0 && sBest[0] > threshold;  // strict, first visit among equals (:85, :94)
        r.score = r.found ? sBest[0] : threshold;
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
    const int nYp = (nY + 3) & ~3;                     // row offsets padded for 16-byte loads
    if ((long long)nX * nY * nT > (1LL << 31) || nX > (1 << 20) || nY > (1 << 20) || nT > (1 << 20))
        return lgs_fail(c, LGS_ERR_INVALID, "gs_match: %d x %d x %d hypotheses per query", nX, nY, nT);
    if (scoreTable && nQ != 1) return lgs_fail(c, LGS_ERR_INVALID, "gs_match: the score table is a 1-query diagnostic");
    LGS_CUDA(c, cudaSetDevice(c->device));
    const double eps = c->opt.edgeEps;
    const size_t nBeamsAll = (size_t)scans->beam_begin[nQ];

    double *dAngles = nullptr, *dRanges = nullptr, *dOff = nullptr;
    LGS_CUDA(c, lgs_alloc_async(c, &dAngles, std::max<size_t>(nBeamsAll, 1) * sizeof(double)));
    LGS_CUDA(c, lgs_alloc_async(c, &dRanges, std::max<size_t>(nBeamsAll, 1) * sizeof(double)));
    LGS_CUDA(c, lgs_alloc_async(c, &dOff, (size_t)(nX + nY + nT + 1) * sizeof(double)));
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
            if (lgs_grid_acquire(c, g) != cudaSuccess) { rcOut = lgs_fail(c, LGS_ERR_CUDA, "gs_match: cannot order after the grid of query %d", q1); break; }
            const int nb = scans->beam_begin[q1 + 1] - scans->beam_begin[q1];
            const long long cc = (long long)nT * nb * nX, rr = (long long)nT * nb * nYp;
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
        double* dScores = nullptr; GsFlag* dFlags = nullptr; GsResult* dRes = nullptr; GsPartial* dPart = nullptr;
        const int groups = (nY + kRows - 1) / kRows;
        const int tiles = (int)(((long long)nX * groups + 255) / 256);
        const int perQuery = nT * tiles;
        auto freeAll = [&]() {
            cudaFreeAsync(dQ, c->stream); cudaFreeAsync(dCol, c->stream); cudaFreeAsync(dRow, c->stream);
            cudaFreeAsync(dScores, c->stream); cudaFreeAsync(dFlags, c->stream);
            cudaFreeAsync(dFlagCount, c->stream); cudaFreeAsync(dRes, c->stream); cudaFreeAsync(dPart, c->stream);
        };
#define GS_TRY(call)                                                                               \
        do { cudaError_t e__ = (call);                                                             \
             if (e__ != cudaSuccess) { freeAll();                                                  \
                 rcOut = lgs_fail(c, e__ == cudaErrorMemoryAllocation ? LGS_ERR_NOMEM : LGS_ERR_CUDA, \
                                  "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); } } while (0)
        GS_TRY(lgs_alloc_async(c, &dQ, nq * sizeof(GsQuery)));
        if (rcOut == LGS_OK && scoreTable) GS_TRY(lgs_alloc_async(c, &dScores, std::max<long long>(scoreCells, 1) * sizeof(double)));
        if (rcOut == LGS_OK) GS_TRY(lgs_alloc_async(c, &dPart, std::max<size_t>((size_t)nq * perQuery, 1) * sizeof(GsPartial)));
        if (rcOut == LGS_OK) GS_TRY(lgs_alloc_async(c, &dFlags, kFlagCap * sizeof(GsFlag)));
        if (rcOut == LGS_OK) GS_TRY(lgs_alloc_async(c, &dFlagCount, sizeof(int)));
        if (rcOut == LGS_OK) GS_TRY(lgs_alloc_async(c, &dRes, nq * sizeof(GsResult)));
        if (rcOut == LGS_OK) GS_TRY(cudaMemcpyAsync(dQ, hq.data(), nq * sizeof(GsQuery), cudaMemcpyHostToDevice, c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaMemsetAsync(dFlagCount, 0, sizeof(int), c->stream));
        if (rcOut != LGS_OK) break;

        const long long hyp = (long long)nX * nY;
        // Fast path: projection fused into the scoring block (no tables); any near-edge beam sends the
        // chunk through the table path below, which carries the exact host fix-up.
        const size_t fusedSmem = (size_t)kFusedBeams * (2 * sizeof(double) + (size_t)(nX + nYp) * sizeof(int));
        const long long fusedThreads = (long long)nX * groups;
        const bool tablesOnly = c->opt.gsTables != 0;   // test / diagnostic hook: force the table path
        bool done = false;
        if (!tablesOnly && hyp > 0 && nT > 0 && maxBeams > 0 && fusedThreads <= 1024 && fusedSmem <= 48 * 1024) {
            dim3 gf(nT, nq);
            if (fusedThreads <= 512)
                gs_fused_kernel<512><<<gf, 512, fusedSmem, c->stream>>>(dQ, dAngles, dRanges, ddX, ddY, ddT, nX, nY, nYp,
                                                                        nT, eps, dScores, dPart, dFlagCount);
            else
                gs_fused_kernel<1024><<<gf, 1024, fusedSmem, c->stream>>>(dQ, dAngles, dRanges, ddX, ddY, ddT, nX, nY, nYp,
                                                                          nT, eps, dScores, dPart, dFlagCount);
            c->launches++;
            int fusedFlags = 0;
            GS_TRY(cudaMemcpyAsync(&fusedFlags, dFlagCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            if (rcOut == LGS_OK) GS_TRY(cudaStreamSynchronize(c->stream));
            if (rcOut != LGS_OK) break;
            if (fusedFlags == 0) {
                gs_select_kernel<<<nq, 256, 0, c->stream>>>(dQ, dPart, nT, nX, nT, dRes);
                c->launches++;
                done = true;
            } else {
                GS_TRY(cudaMemsetAsync(dFlagCount, 0, sizeof(int), c->stream));
                if (rcOut != LGS_OK) break;
            }
        }
        std::vector<int> fixups(nq, 0);
        if (!done) {
        GS_TRY(lgs_alloc_async(c, &dCol, std::max<long long>(colCells, 1) * sizeof(int)));
        if (rcOut == LGS_OK) GS_TRY(lgs_alloc_async(c, &dRow, std::max<long long>(rowCells, 1) * sizeof(int)));
        if (rcOut != LGS_OK) break;
        if (maxBeams > 0 && nT > 0) {
            dim3 gp((unsigned)(((long long)nT * maxBeams + 127) / 128), nq);
            const size_t stage = (size_t)128 * (nX + nYp) * sizeof(int);
            if (stage <= 48 * 1024)
                gs_project_kernel<true><<<gp, 128, stage, c->stream>>>(dQ, dAngles, dRanges, ddX, ddY, ddT, nX, nY,
                                                                       nYp, nT, eps, dCol, dRow, dFlags, dFlagCount);
            else
                gs_project_kernel<false><<<gp, 128, 0, c->stream>>>(dQ, dAngles, dRanges, ddX, ddY, ddT, nX, nY,
                                                                    nYp, nT, eps, dCol, dRow, dFlags, dFlagCount);
            c->launches++;
        }
        int nFlag = 0;
        GS_TRY(cudaMemcpyAsync(&nFlag, dFlagCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if (rcOut == LGS_OK) GS_TRY(cudaStreamSynchronize(c->stream));
        if (rcOut != LGS_OK) break;
        if (nFlag > kFlagCap) { freeAll(); rcOut = lgs_fail(c, LGS_ERR_OVERFLOW, "gs_match: %d near-edge beams exceed the fix-up list", nFlag); break; }
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
                    GS_TRY(cudaMemcpy(dRow + h.rowBegin + ((long long)k * h.nBeams + b) * nYp, hr.data(), nY * sizeof(int), cudaMemcpyHostToDevice));
                fixups[fl[f].q]++;
            }
            if (rcOut != LGS_OK) break;
        }
        if (hyp > 0 && nT > 0) {
            dim3 gs(tiles, nT, nq);
            gs_score_kernel<<<gs, 256, 0, c->stream>>>(dQ, dCol, dRow, nX, nY, nYp, nT, dScores, dPart);
            c->launches++;
        }
        gs_select_kernel<<<nq, 256, 0, c->stream>>>(dQ, dPart, (hyp > 0 && nT > 0) ? perQuery : 0, nX, nT, dRes);
        c->launches++;
        }   // table path
        GS_TRY(cudaGetLastError());
        std::vector<GsResult> hr(nq);
        if (rcOut == LGS_OK) GS_TRY(cudaMemcpyAsync(hr.data(), dRes, nq * sizeof(GsResult), cudaMemcpyDeviceToHost, c->stream));
        if (rcOut == LGS_OK && scoreTable && dScores && scoreCells > 0)
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
