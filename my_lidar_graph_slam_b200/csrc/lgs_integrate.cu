// lgs_integrate.cu -- occupancy-grid scan integration on sm_100a.
//
// Replaces the integration loops of GridMapBuilder::UpdateGridMap / ConstructMapFromScans
// (mapping/grid_map_builder.cpp:170-186, :311-328): for every beam that passed the range
// filter, the cells on Bresenham(sensorCell -> hitCell) (util.hpp:257-303) except the last get
// Update(pMiss), the last gets Update(pHit) (BinaryBayesGridCell::Update,
// grid_map/binary_bayes_grid_cell.hpp:75-119).
//
// The cell update is order dependent at the bit level: each cell must see its
// touches in (scan, beam) order.  Instead of scattering updates (atomics cannot be ordered),
// every grid cell is OWNED by one thread, which walks the scans of the batch in order and, for
// each scan, the beams that can touch the cell in beam order, deciding membership in O(1) from
// the closed form of the reference's Bresenham:
//     x-major (|dx| > |dy|):  y_k = y0 + sy * floor((2|dy|k + |dx|) / (2|dx|)),  k = 0..|dx|
//     y-major (otherwise)  :  x_k = x0 + sx * floor((2|dx|k + |dy|) / (2|dy|)),  k = 0..|dy|
// (cells k < length are misses, k = length is the hit).  That is conflict free by
// construction and applies exactly the CPU's IEEE sequence (div/mul/add intrinsics, no FMA).
//
// Finding the candidate beams of a cell:
//  * far field (Chebyshev distance to the sensor cell > kNear): beams are angularly sorted, so
//    the candidates are a binary-searched window of +-(2.2 / d + 0.001) rad around the cell's
//    direction -- a proven superset (cell centres on a Bresenham line lie within 0.5 cell of
//    the centre-to-centre segment, whose end points are within 0.71 cell of the true ray);
//  * near field (<= kNear cells): nearly every beam passes, so one warp per (scan, near cell)
//    tests all beams with ballots once and stores the ordered touch sequence run-length
//    encoded (M^a H^b M^c ...); the owning thread then just applies the runs, stopping a run
//    early once the value reaches its fixed point.
// Scans whose beams are not angularly monotone fall back to testing every beam (still exact).
#include <cmath>

#include "lgs_internal.cuh"

namespace {

constexpr int kNear = 16;                      // near-field half width (cells)
constexpr int kNearW = 2 * kNear + 1;
constexpr int kRuns = 15;                      // RLE runs kept per (scan, near cell)
constexpr unsigned short kRleOverflow = 0xFFFF;
constexpr float kTwoPi = 6.28318530717958647692f;

struct ScanMeta {
    int sx, sy;          // sensor cell
    int beamBegin, n;    // beams of this scan
    int maxLen;          // max Chebyshev ray length in cells
    int unsorted;        // beams not angularly monotone -> test all beams
    float ang0;          // world angle of beam 0
    int bad;             // a touched cell lies outside the grid
};

struct GridRef {
    double* origin;
    int nx, ny, pitch;
    double minX, minY, res;
};

__device__ __forceinline__ double clampProb(double v) {
    // std::clamp(v, ProbabilityMin, ProbabilityMax)  (binary_bayes_grid_cell.hpp:50-52, :97-101)
    const double lo = 1e-3, hi = 1.0 - 1e-3;
    return v < lo ? lo : (hi < v ? hi : v);
}

// BinaryBayesGridCell<double>::Update (binary_bayes_grid_cell.hpp:75-92); oddsP = ValueToOdds(p).
__device__ __forceinline__ double bayesUpdate(double v, double p, double oddsP) {
    if (v == 0.0) return clampProb(p);
    const double cv = clampProb(v);
    const double oldOdds = __ddiv_rn(cv, __dsub_rn(1.0, cv));           // ValueToOdds
    const double o = __dmul_rn(oldOdds, oddsP);
    const double nv = clampProb(__ddiv_rn(o, __dadd_rn(1.0, o)));       // OddsToValue
    return clampProb(nv);
}

// 0 = not on the ray, 1 = miss cell, 2 = hit cell.  (rx, ry) = cell - sensorCell, (ex, ey) = hitCell - sensorCell.
// Division free: with k the step along the major axis, the reference's minor coordinate is
// floor((2|minor| k + |major|) / (2|major|)) (util.hpp:276-299), i.e. the cell is on the ray iff
//     2|major| t <= 2|minor| k + |major| < 2|major| (t + 1),   t = signed minor offset >= 0.
__device__ __forceinline__ int rayTouch(int rx, int ry, int ex, int ey) {
    const int ax = abs(ex), ay = abs(ey);
    int k, t, amaj, amin;
    if (ax > ay) {                                   // x-major (util.hpp:276-287)
        k = ex < 0 ? -rx : rx; t = ey < 0 ? -ry : ry; amaj = ax; amin = ay;
    } else {                                         // y-major (util.hpp:288-299), also dx == dy == 0
        k = ey < 0 ? -ry : ry; t = ex < 0 ? -rx : rx; amaj = ay; amin = ax;
    }
    if (k < 0 || k > amaj || t < 0) return 0;
    const int lhs = 2 * amin * k + amaj;             // < 2^31 for maps below 16k cells per side
    const int m2 = 2 * amaj;
    if (amaj == 0) return (t == 0) ? 2 : 0;          // zero-length ray: the sensor cell is the hit
    if (m2 * t > lhs || lhs >= m2 * (t + 1)) return 0;
    return k == amaj ? 2 : 1;
}

__device__ __forceinline__ float wrapBeta(float a) {   // into [-0.01, 2*pi - 0.01)
    a = fmodf(a, kTwoPi);
    if (a < -0.01f) a += kTwoPi;
    if (a >= kTwoPi - 0.01f) a -= kTwoPi;
    return a;
}

// ---- pre-pass A: sensor cells ---------------------------------------------------------------------
__global__ void integ_sensor_kernel(const double* __restrict__ sensorXY, const int* __restrict__ hitBegin,
                                    const double* __restrict__ hitXY, int nScans, GridRef g,
                                    ScanMeta* __restrict__ meta) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nScans) return;
    ScanMeta m;
    // WorldCoordinateToGridCellIndex (grid_map.hpp:779-790)
    m.sx = __double2int_rd(__ddiv_rn(__dsub_rn(sensorXY[2 * s], g.minX), g.res));
    m.sy = __double2int_rd(__ddiv_rn(__dsub_rn(sensorXY[2 * s + 1], g.minY), g.res));
    m.beamBegin = hitBegin[s];
    m.n = hitBegin[s + 1] - hitBegin[s];
    m.maxLen = 0; m.unsorted = 0; m.bad = 0;
    m.ang0 = 0.f;
    if (m.n > 0)
        m.ang0 = atan2f((float)(hitXY[2 * (size_t)m.beamBegin + 1] - sensorXY[2 * s + 1]),
                        (float)(hitXY[2 * (size_t)m.beamBegin] - sensorXY[2 * s]));
    if (m.sx < 0 || m.sx >= g.nx || m.sy < 0 || m.sy >= g.ny) m.bad = 1;
    meta[s] = m;
}

// ---- pre-pass B: per beam end cell (relative), angle ------------------------------------------------
__global__ void integ_beam_kernel(const double* __restrict__ sensorXY, const double* __restrict__ hitXY,
                                  int nScans, GridRef g, ScanMeta* __restrict__ meta,
                                  int2* __restrict__ rel, float* __restrict__ beta) {
    const int s = blockIdx.y;
    const ScanMeta m = meta[s];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.n) return;
    const size_t b = (size_t)m.beamBegin + i;
    const double hx = hitXY[2 * b], hy = hitXY[2 * b + 1];
    const int ex = __double2int_rd(__ddiv_rn(__dsub_rn(hx, g.minX), g.res));
    const int ey = __double2int_rd(__ddiv_rn(__dsub_rn(hy, g.minY), g.res));
    if (ex < 0 || ex >= g.nx || ey < 0 || ey >= g.ny) atomicOr(&meta[s].bad, 1);
    const int dx = ex - m.sx, dy = ey - m.sy;
    rel[b] = make_int2(dx, dy);
    atomicMax(&meta[s].maxLen, max(abs(dx), abs(dy)));
    const float a = atan2f((float)(hy - sensorXY[2 * s + 1]), (float)(hx - sensorXY[2 * s]));
    beta[b] = i == 0 ? 0.f : wrapBeta(a - m.ang0);
}

__global__ void integ_sorted_kernel(ScanMeta* __restrict__ meta, const float* __restrict__ beta) {
    const int s = blockIdx.y;
    const ScanMeta m = meta[s];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 1 || i >= m.n) return;
    const size_t b = (size_t)m.beamBegin + i;
    if (beta[b] < beta[b - 1]) atomicOr(&meta[s].unsorted, 1);
}

// ---- pre-pass C: near-field touch sequences, run-length encoded ---------------------------------------
// One warp per (scan, near cell).  Entry = (type << 15) | count with type 1 = hit; 0 terminates.
__global__ void __launch_bounds__(128)
integ_near_kernel(const ScanMeta* __restrict__ meta, const int2* __restrict__ rel, int nScans,
                  unsigned short* __restrict__ table) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= nScans * kNearW * kNearW) return;
    const int s = warp / (kNearW * kNearW);
    const int c = warp - s * (kNearW * kNearW);
    const int rx = c % kNearW - kNear, ry = c / kNearW - kNear;
    const ScanMeta m = meta[s];
    unsigned short* out = table + (size_t)warp * (kRuns + 1);
    int nRuns = 0, curType = -1, curCount = 0;
    bool overflow = false;
    for (int base = 0; base < m.n; base += 32) {
        const int i = base + lane;
        int ty = 0;
        if (i < m.n) {
            const int2 e = __ldg(rel + m.beamBegin + i);
            ty = rayTouch(rx, ry, e.x, e.y);
        }
        unsigned touched = __ballot_sync(0xffffffffu, ty != 0);
        const unsigned hits = __ballot_sync(0xffffffffu, ty == 2);
        while (touched) {                       // uniform across the warp
            const int bit = __ffs(touched) - 1;
            const int type = (hits >> bit) & 1;
            // length of the run of equal type among the touched bits starting at `bit`
            const unsigned same = type ? (touched & hits) : (touched & ~hits);
            const unsigned other = touched & ~same;
            const unsigned upto = other ? ((1u << (__ffs(other) - 1)) - 1u) : 0xffffffffu;
            const unsigned runBits = same & upto;
            const int cnt = __popc(runBits);
            if (type == curType) {
                curCount += cnt;
            } else {
                if (curType >= 0) {
                    if (nRuns < kRuns) { if (lane == 0) out[nRuns] = (unsigned short)((curType << 15) | curCount); }
                    else overflow = true;
                    ++nRuns;
                }
                curType = type; curCount = cnt;
            }
            touched &= ~runBits;
        }
    }
    if (curType >= 0) {
        if (nRuns < kRuns) { if (lane == 0) out[nRuns] = (unsigned short)((curType << 15) | curCount); }
        else overflow = true;
        ++nRuns;
    }
    if (lane == 0) {
        if (overflow || m.n > 32767) out[0] = kRleOverflow;
        else out[nRuns] = 0;
    }
}

// ---- ray walks: counting sort of the far-field touches by cell ----------------------------------------
// One thread per (scan, beam) walks the reference's Bresenham (util.hpp:257-303) twice.  Pass 1
// (FILL = false) ORs the scan's bit into every touched cell and counts the far-field touches per
// cell; after an exclusive prefix sum over the counts, pass 2 (FILL = true) drops one 32-bit key per
// far-field touch into the cell's own segment.  Near-field touches (<= kNear cells from the sensor
// cell: thousands per cell) are not recorded; the run-length table covers them.
//   key = 1 + ((scan << 17) | (beam << 1) | isHit): ascending key == the CPU's (scan, beam) order.
template <bool FILL>
__global__ void __launch_bounds__(128)
integ_walk_kernel(const ScanMeta* __restrict__ meta, const int2* __restrict__ rel, int nScans,
                  int x0, int y0, int rw, unsigned long long* __restrict__ mask,
                  unsigned* __restrict__ cnt, const unsigned* __restrict__ local,
                  const unsigned* __restrict__ blockSums, unsigned* __restrict__ keys) {
    const int s = blockIdx.y;
    const ScanMeta m = meta[s];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m.n) return;
    const int2 e = __ldg(rel + m.beamBegin + i);
    const unsigned long long bit = 1ull << s;
    int x = m.sx, y = m.sy;
    const int x1 = m.sx + e.x, y1 = m.sy + e.y;
    const int sx = e.x < 0 ? -1 : 1, sy = e.y < 0 ? -1 : 1;
    const int dx = abs(e.x * 2), dy = abs(e.y * 2);
    const unsigned keyMiss = 1u + (((unsigned)s << 17) | ((unsigned)i << 1));
    auto touch = [&]() {
        const size_t c = (size_t)(y - y0) * rw + (x - x0);
        const bool nearField = max(abs(x - m.sx), abs(y - m.sy)) <= kNear;
        if (!FILL) {
            if (!(mask[c] & bit)) atomicOr(mask + c, bit);      // racy pre-check only saves atomics
            if (!nearField) atomicAdd(cnt + c, 1u);
        } else if (!nearField) {
            const unsigned slot = local[c] + blockSums[c >> 10] + atomicAdd(cnt + c, 1u);
            keys[slot] = keyMiss + ((x == x1 && y == y1) ? 1u : 0u);
        }
    };
    touch();
    if (dx > dy) {
        int err = dy - dx / 2;
        while (x != x1) {
            if (err >= 0) { y += sy; err -= dx; }
            x += sx; err += dy;
            touch();
        }
    } else {
        int err = dx - dy / 2;
        while (y != y1) {
            if (err >= 0) { x += sx; err -= dy; }
            y += sy; err += dx;
            touch();
        }
    }
}

// Exclusive prefix sum over the per-cell counts: 1024 cells per block, then the block totals.
__global__ void __launch_bounds__(256)
integ_scan_blocks_kernel(const unsigned* __restrict__ cnt, size_t n, unsigned* __restrict__ local,
                         unsigned* __restrict__ blockSums) {
    __shared__ unsigned sWarp[8];
    const size_t base = (size_t)blockIdx.x * 1024 + threadIdx.x * 4;
    unsigned v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = base + k < n ? cnt[base + k] : 0u; sum += v[k]; }
    unsigned inc = sum;                                   // inclusive scan of the per-thread sums
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) sWarp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned w = lane < 8 ? sWarp[lane] : 0u;
        for (int o = 1; o < 8; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
        if (lane < 8) sWarp[lane] = w;
    }
    __syncthreads();
    unsigned run = inc - sum + (wid ? sWarp[wid - 1] : 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < n) local[base + k] = run; run += v[k]; }
    if (threadIdx.x == 255) blockSums[blockIdx.x] = sWarp[7];
}

__global__ void __launch_bounds__(1024)
integ_scan_sums_kernel(unsigned* __restrict__ blockSums, int nBlocks) {
    __shared__ unsigned sWarp[32];
    __shared__ unsigned sCarry;
    if (threadIdx.x == 0) sCarry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nBlocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nBlocks ? blockSums[i] : 0u;
        unsigned inc = v;
        for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) sWarp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned w = sWarp[lane];
            for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
            sWarp[lane] = w;
        }
        __syncthreads();
        const unsigned excl = sCarry + inc - v + (wid ? sWarp[wid - 1] : 0u);
        if (i < nBlocks) blockSums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) sCarry = excl + v;
        __syncthreads();
    }
}

// ---- apply pass: one thread owns one cell ----------------------------------------------------------
constexpr unsigned kSortCap = 160;     // recorded touches a thread orders by repeated-minimum selection

struct ApplyArgs {
    const ScanMeta* meta;
    const int2* rel;
    const float* beta;
    const unsigned short* nearTab;
    const unsigned long long* mask;
    const unsigned* cnt;          // far-field touches per cell
    const unsigned* local;        // exclusive prefix sum within 1024-cell blocks
    const unsigned* blockSums;    // exclusive prefix sum of the block totals
    const unsigned* keys;
    double pHit, pMiss, oddsHit, oddsMiss;
};

__device__ __forceinline__ double applyTouch(double v, bool hit, const ApplyArgs& a) {
    return hit ? bayesUpdate(v, a.pHit, a.oddsHit) : bayesUpdate(v, a.pMiss, a.oddsMiss);
}

// All touches of scan `m` on the cell at (rx, ry) from its sensor cell, in beam order, by testing
// every beam (exact, slow): the fallback when the fast paths cannot be used or disagree.
__device__ __forceinline__ double applyAllBeams(double v, const ScanMeta& m, const int2* __restrict__ e,
                                                int rx, int ry, const ApplyArgs& a, unsigned& count) {
    for (int i = 0; i < m.n; ++i) {
        const int2 ee = __ldg(e + i);
        const int ty = rayTouch(rx, ry, ee.x, ee.y);
        if (ty) { v = applyTouch(v, ty == 2, a); ++count; }
    }
    return v;
}

// Near field of scan s: apply the run-length encoded touch sequence.
__device__ __forceinline__ double applyNear(double v, int s, const ScanMeta& m, int rx, int ry,
                                            const ApplyArgs& a, unsigned& count) {
    const unsigned short* t = a.nearTab + ((size_t)s * kNearW * kNearW + (ry + kNear) * kNearW + (rx + kNear)) * (kRuns + 1);
    if (t[0] == kRleOverflow) return applyAllBeams(v, m, a.rel + m.beamBegin, rx, ry, a, count);
    for (int k = 0; k < kRuns; ++k) {
        const unsigned short ent = t[k];
        if (ent == 0) break;
        const int cnt = ent & 0x7fff;
        const bool hit = (ent >> 15) != 0;
        count += cnt;
        for (int j = 0; j < cnt; ++j) {
            const double nv = applyTouch(v, hit, a);
            if (nv == v) break;        // fixed point: the rest of the run is a no-op
            v = nv;
        }
    }
    return v;
}

// Far field of scan s without a recorded list: binary-searched angular window of candidate beams.
__device__ __forceinline__ double applyFarSearch(double v, const ScanMeta& m, int rx, int ry,
                                                 const ApplyArgs& a, unsigned& count) {
    const int2* __restrict__ e = a.rel + m.beamBegin;
    if (m.unsorted) return applyAllBeams(v, m, e, rx, ry, a, count);
    const float* __restrict__ bt = a.beta + m.beamBegin;
    const float d = sqrtf((float)(rx * rx + ry * ry));
    const float delta = 2.2f / d + 1e-3f;
    const float bc = wrapBeta(atan2f((float)ry, (float)rx) - m.ang0);
    const float bLast = __ldg(bt + m.n - 1);
#pragma unroll 1
    for (int w = -1; w <= 1; ++w) {
        const float lo = bc + w * kTwoPi - delta, hi = bc + w * kTwoPi + delta;
        if (hi < -0.01f || lo > bLast) continue;
        int lb = 0, ub = m.n;                  // first i with beta[i] >= lo
        while (lb < ub) {
            const int mid = (lb + ub) >> 1;
            if (__ldg(bt + mid) < lo) lb = mid + 1; else ub = mid;
        }
        for (int i = lb; i < m.n && __ldg(bt + i) <= hi; ++i) {
            const int2 ee = __ldg(e + i);
            const int ty = rayTouch(rx, ry, ee.x, ee.y);
            if (ty) { v = applyTouch(v, ty == 2, a); ++count; }
        }
    }
    return v;
}

__global__ void __launch_bounds__(256)
integ_apply_kernel(ApplyArgs a, GridRef g, int x0, int y0, int x1, int y1,
                   unsigned long long* __restrict__ counters /* [0] updates, [1] fallback cells */) {
    const int cx = x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = y0 + blockIdx.y * blockDim.y + threadIdx.y;
    unsigned total = 0;
    if (cx < x1 && cy < y1) {
        const size_t ridx = (size_t)(cy - y0) * (x1 - x0) + (cx - x0);
        const unsigned long long mk = a.mask[ridx];
        if (mk) {
            double* cell = g.origin + (size_t)cy * g.pitch + cx;
            const double v0 = *cell;
            double v = v0;
            const unsigned nFar = a.cnt[ridx];
            const unsigned* __restrict__ K = a.keys + a.local[ridx] + a.blockSums[ridx >> 10];
            unsigned nearCount = 0, farCount = 0;
            unsigned prev = 0;
            for (unsigned long long r = mk; r; r &= r - 1) {           // scans in order
                const int s = __ffsll((long long)r) - 1;
                const ScanMeta m = a.meta[s];
                const int rx = cx - m.sx, ry = cy - m.sy;
                if (max(abs(rx), abs(ry)) <= kNear) { v = applyNear(v, s, m, rx, ry, a, nearCount); continue; }
                if (nFar > kSortCap) { v = applyFarSearch(v, m, rx, ry, a, farCount); continue; }
                // this scan's recorded touches in ascending key (= beam) order, by repeated minimum
                const unsigned hiKey = (((unsigned)s + 1u) << 17);     // keys of scan s are in (s<<17, hiKey]
                prev = max(prev, (unsigned)s << 17);
                for (;;) {
                    unsigned bestKey = 0xffffffffu;
                    for (unsigned j = 0; j < nFar; ++j) {
                        const unsigned key = __ldg(K + j);
                        if (key > prev && key < bestKey) bestKey = key;
                    }
                    if (bestKey > hiKey) break;
                    prev = bestKey;
                    v = applyTouch(v, ((bestKey - 1u) & 1u) != 0u, a);
                    ++farCount;
                }
            }
            if (farCount != nFar) {
                // The fast paths disagree with the count pass: redo this cell the slow, exact way.
                v = v0; nearCount = 0; farCount = 0;
                for (unsigned long long r = mk; r; r &= r - 1) {
                    const int s = __ffsll((long long)r) - 1;
                    const ScanMeta m = a.meta[s];
                    v = applyAllBeams(v, m, a.rel + m.beamBegin, cx - m.sx, cy - m.sy, a, farCount);
                }
                atomicAdd(counters + 1, 1ull);
            }
            total = nearCount + farCount;
            if (v != v0) *cell = v;
        }
    }
    __shared__ unsigned sCount[8];
    unsigned c = total;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if ((tid & 31) == 0) sCount[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        unsigned long long tot = 0;
        for (int k = 0; k < (int)(blockDim.x * blockDim.y + 31) / 32; ++k) tot += sCount[k];
        if (tot) atomicAdd(counters, tot);
    }
}

__global__ void grid_shift_copy_kernel(const double* __restrict__ src, int srcNx, int srcNy, int srcPitch,
                                       double* __restrict__ dst, int dstNx, int dstNy, int dstPitch,
                                       int shiftX, int shiftY) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dstNx || y >= dstNy) return;
    const int ox = x + shiftX, oy = y + shiftY;
    double v = 0.0;
    if (ox >= 0 && ox < srcNx && oy >= 0 && oy < srcNy) v = src[(size_t)oy * srcPitch + ox];
    dst[(size_t)y * dstPitch + x] = v;
}

}  // namespace

extern "C" {

int lgs_grid_integrate_scans(lgs_ctx* c, lgs_grid* grid, const lgs_hit_batch* scans, double pHit,
                             double pMiss, long long* nUpdatesOut) {
    if (!c || !grid || !scans) return LGS_ERR_INVALID;
    if (nUpdatesOut) *nUpdatesOut = 0;
    const int n = scans->n_scans;
    if (n < 0 || (n > 0 && (!scans->sensor_xy || !scans->hit_begin)))
        return lgs_fail(c, LGS_ERR_INVALID, "integrate: bad scan batch");
    if (n == 0) return LGS_OK;
    const long long total = scans->hit_begin[n];
    if (total > 0 && !scans->hit_xy) return lgs_fail(c, LGS_ERR_INVALID, "integrate: hit_xy is NULL");
    LGS_CUDA(c, cudaSetDevice(c->device));
    if (!c->integ) c->integ = new lgs_integ_ws();
    lgs_integ_ws& w = *c->integ;

    // Stage the whole batch once; the kernels then run in sub-batches of <= 64 scans (one mask bit each).
    LGS_CUDA(c, w.sensor.reserve((size_t)n * 2));
    LGS_CUDA(c, w.hit.reserve(std::max<size_t>((size_t)total, 1) * 2));
    LGS_CUDA(c, w.begin.reserve((size_t)n + 1));
    LGS_CUDA(c, w.meta.reserve((size_t)n * sizeof(ScanMeta)));
    LGS_CUDA(c, w.rel.reserve(std::max<size_t>((size_t)total, 1)));
    LGS_CUDA(c, w.beta.reserve(std::max<size_t>((size_t)total, 1)));
    LGS_CUDA(c, w.nearTab.reserve((size_t)std::min(n, 64) * kNearW * kNearW * (kRuns + 1)));
    LGS_CUDA(c, w.counters.reserve(8));
    LGS_CUDA(c, w.hMeta.reserve((size_t)n * sizeof(ScanMeta)));
    LGS_CUDA(c, w.hCounters.reserve(8));
    LGS_CUDA(c, cudaMemcpyAsync(w.sensor.p, scans->sensor_xy, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (total)
        LGS_CUDA(c, cudaMemcpyAsync(w.hit.p, scans->hit_xy, (size_t)total * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    LGS_CUDA(c, cudaMemcpyAsync(w.begin.p, scans->hit_begin, (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    LGS_CUDA(c, cudaMemsetAsync(w.counters.p, 0, 8 * sizeof(unsigned long long), c->stream));

    int maxBeams = 0;
    for (int s = 0; s < n; ++s) maxBeams = std::max(maxBeams, scans->hit_begin[s + 1] - scans->hit_begin[s]);
    if (maxBeams > 32767) return lgs_fail(c, LGS_ERR_INVALID, "integrate: %d beams in one scan (limit 32767)", maxBeams);
    ScanMeta* dMeta = reinterpret_cast<ScanMeta*>(w.meta.p);
    ScanMeta* hMeta = reinterpret_cast<ScanMeta*>(w.hMeta.p);
    GridRef g{grid->origin(), grid->nx, grid->ny, grid->pitch, grid->min_x, grid->min_y, grid->res};
    // Pre-pass over the whole batch: sensor cells, relative end cells, beam angles, sortedness.
    integ_sensor_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(w.sensor.p, w.begin.p, w.hit.p, n, g, dMeta);
    LGS_LAUNCH_CHECK(c);
    if (maxBeams > 0) {
        dim3 gb((maxBeams + 127) / 128, n);
        integ_beam_kernel<<<gb, 128, 0, c->stream>>>(w.sensor.p, w.hit.p, n, g, dMeta, w.rel.p, w.beta.p);
        LGS_LAUNCH_CHECK(c);
        integ_sorted_kernel<<<gb, 128, 0, c->stream>>>(dMeta, w.beta.p);
        LGS_LAUNCH_CHECK(c);
    }
    LGS_CUDA(c, cudaMemcpyAsync(hMeta, dMeta, (size_t)n * sizeof(ScanMeta), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int s = 0; s < n; ++s)
        if (hMeta[s].bad)
            return lgs_fail(c, LGS_ERR_INVALID, "integrate: scan %d touches cells outside the %dx%d grid "
                            "(expand the map first, as GridMap::Expand does)", s, grid->nx, grid->ny);

    // ValueToOdds(prob) for the two observations (binary_bayes_grid_cell.hpp:104-113), host IEEE.
    auto clampP = [](double v) { const double lo = 1e-3, hi = 1.0 - 1e-3; return v < lo ? lo : (hi < v ? hi : v); };
    const double oddsHit = clampP(pHit) / (1.0 - clampP(pHit));
    const double oddsMiss = clampP(pMiss) / (1.0 - clampP(pMiss));

    // Sub-batches: every scan owns one mask bit (<= 64).
    int sub = 64;
    if (const char* e = getenv("LGS_INTEG_SUB")) sub = std::min(64, std::max(1, atoi(e)));   // tuning hook
    for (int s0 = 0; s0 < n; s0 += sub) {
        const int ns = std::min(sub, n - s0);
        int x0 = grid->nx, y0 = grid->ny, x1 = 0, y1 = 0, subBeams = 0;
        for (int s = s0; s < s0 + ns; ++s) {
            if (hMeta[s].n == 0) continue;
            subBeams = std::max(subBeams, hMeta[s].n);
            x0 = std::min(x0, hMeta[s].sx - hMeta[s].maxLen); x1 = std::max(x1, hMeta[s].sx + hMeta[s].maxLen + 1);
            y0 = std::min(y0, hMeta[s].sy - hMeta[s].maxLen); y1 = std::max(y1, hMeta[s].sy + hMeta[s].maxLen + 1);
        }
        x0 = std::max(x0, 0); y0 = std::max(y0, 0); x1 = std::min(x1, grid->nx); y1 = std::min(y1, grid->ny);
        if (subBeams == 0 || x1 <= x0 || y1 <= y0) continue;
        const size_t region = (size_t)(x1 - x0) * (y1 - y0);
        const int nScanBlocks = (int)((region + 1023) / 1024);
        long long subTouches = 0;                       // upper bound of recorded keys: sum of ray lengths
        for (int s = s0; s < s0 + ns; ++s) subTouches += (long long)hMeta[s].n * (hMeta[s].maxLen + 1);
        LGS_CUDA(c, w.mask.reserve(region));
        LGS_CUDA(c, w.expect.reserve(region));
        LGS_CUDA(c, w.local.reserve(region));
        LGS_CUDA(c, w.blockSums.reserve((size_t)nScanBlocks));
        LGS_CUDA(c, w.lists.reserve((size_t)std::max<long long>(subTouches, 1)));
        LGS_CUDA(c, cudaMemsetAsync(w.mask.p, 0, region * sizeof(unsigned long long), c->stream));
        LGS_CUDA(c, cudaMemsetAsync(w.expect.p, 0, region * sizeof(unsigned), c->stream));
        dim3 gm((subBeams + 127) / 128, ns);
        integ_walk_kernel<false><<<gm, 128, 0, c->stream>>>(dMeta + s0, w.rel.p, ns, x0, y0, x1 - x0, w.mask.p,
                                                            w.expect.p, nullptr, nullptr, nullptr);
        LGS_LAUNCH_CHECK(c);
        integ_scan_blocks_kernel<<<nScanBlocks, 256, 0, c->stream>>>(w.expect.p, region, w.local.p, w.blockSums.p);
        LGS_LAUNCH_CHECK(c);
        integ_scan_sums_kernel<<<1, 1024, 0, c->stream>>>(w.blockSums.p, nScanBlocks);
        LGS_LAUNCH_CHECK(c);
        // pass 2 re-uses the count array as the per-cell cursor; the apply kernel needs the counts,
        // which the cursors equal again once every touch has been recorded
        LGS_CUDA(c, cudaMemsetAsync(w.expect.p, 0, region * sizeof(unsigned), c->stream));
        integ_walk_kernel<true><<<gm, 128, 0, c->stream>>>(dMeta + s0, w.rel.p, ns, x0, y0, x1 - x0, w.mask.p,
                                                           w.expect.p, w.local.p, w.blockSums.p, w.lists.p);
        LGS_LAUNCH_CHECK(c);
        const long long warps = (long long)ns * kNearW * kNearW;
        integ_near_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, c->stream>>>(dMeta + s0, w.rel.p, ns, w.nearTab.p);
        LGS_LAUNCH_CHECK(c);
        ApplyArgs a{dMeta + s0, w.rel.p, w.beta.p, w.nearTab.p, w.mask.p, w.expect.p, w.local.p, w.blockSums.p,
                    w.lists.p, pHit, pMiss, oddsHit, oddsMiss};
        dim3 block(32, 8), gridDim((x1 - x0 + 31) / 32, (y1 - y0 + 7) / 8);
        integ_apply_kernel<<<gridDim, block, 0, c->stream>>>(a, g, x0, y0, x1, y1, w.counters.p);
        LGS_LAUNCH_CHECK(c);
    }
    LGS_CUDA(c, cudaMemcpyAsync(w.hCounters.p, w.counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (nUpdatesOut) *nUpdatesOut = (long long)w.hCounters.p[0];
    w.fallbackCells += (long long)w.hCounters.p[1];
    return LGS_OK;
}

long long lgs_ctx_integrate_fallback_cells(const lgs_ctx* c) { return (c && c->integ) ? c->integ->fallbackCells : 0; }

int lgs_grid_resize(lgs_grid* g, int nx, int ny, double minX, double minY, int shiftX, int shiftY) {
    if (!g) return LGS_ERR_INVALID;
    lgs_ctx* c = g->ctx;
    if (nx < 0 || ny < 0) return lgs_fail(c, LGS_ERR_INVALID, "grid_resize: %dx%d", nx, ny);
    const long long pitch = (long long)nx + 2LL * g->apron, rows = (long long)ny + 2LL * g->apron;
    if (pitch * rows >= (1LL << 31)) return lgs_fail(c, LGS_ERR_INVALID, "grid_resize: too many cells");
    LGS_CUDA(c, cudaSetDevice(c->device));
    double* nd = nullptr;
    const size_t bytes = (size_t)pitch * rows * sizeof(double);
    cudaError_t e = cudaMalloc(&nd, bytes);
    if (e != cudaSuccess) return lgs_fail(c, LGS_ERR_NOMEM, "grid_resize: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    LGS_CUDA(c, cudaMemsetAsync(nd, 0, bytes, c->stream));
    if (nx > 0 && ny > 0) {
        dim3 gridDim((nx + 255) / 256, ny);
        grid_shift_copy_kernel<<<gridDim, 256, 0, c->stream>>>(g->origin(), g->nx, g->ny, g->pitch,
                                                               nd + (size_t)g->apron * pitch + g->apron,
                                                               nx, ny, (int)pitch, shiftX, shiftY);
        LGS_LAUNCH_CHECK(c);
    }
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(g->d);
    g->d = nd; g->nx = nx; g->ny = ny; g->pitch = (int)pitch; g->rows = (int)rows;
    g->min_x = minX; g->min_y = minY;
    return LGS_OK;
}

int lgs_grid_clear(lgs_grid* g) {
    if (!g) return LGS_ERR_INVALID;
    lgs_ctx* c = g->ctx;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaMemsetAsync(g->d, 0, (size_t)g->pitch * g->rows * sizeof(double), c->stream));
    return LGS_OK;
}

}  // extern "C"
