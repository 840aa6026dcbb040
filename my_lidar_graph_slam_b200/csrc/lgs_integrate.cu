// lgs_integrate.cu -- occupancy-grid scan integration on sm_100a.
//
// Replaces the integration loops of GridMapBuilder::UpdateGridMap / ConstructMapFromScans
// (mapping/grid_map_builder.cpp:170-186, :311-328): for every beam that passed the range
// filter, the cells on Bresenham(sensorCell -> hitCell) (util.hpp:257-303) except the last get
// Update(pMiss), the last gets Update(pHit) (BinaryBayesGridCell::Update,
// grid_map/binary_bayes_grid_cell.hpp:75-119).
//
// The cell update is order dependent at the bit level: each cell must see its
// touches in (scan, beam) order, so updates cannot be scattered with atomics.  But WHICH touches a
// cell sees, and in which order, does not depend on the cell values.  The work is therefore split
// like a tiled rasteriser, over chunks of <= 64 scans:
//
//   1. mark   one thread per (scan, beam) walks the reference's Bresenham and, whenever the ray
//             enters another 16x16-cell tile, folds its beam index into the tile's [kmin, kmax)
//             range for that scan (idempotent atomics: min / max / or -- order free);
//   2. pairs  every tile turns its scan mask into a contiguous, scan-ordered run of
//             (tile, scan, kmin, kmax) pairs;
//   3. touch  fully parallel over the pairs: one block per pair drops the cells of the beams
//             kmin..kmax that fall into the tile into per-cell bitmaps in shared memory (bit = beam,
//             so ascending bits ARE the CPU's beam order), using the closed form of the
//             reference's Bresenham (step k of a ray is independent of step k - 1):
//                 x-major (|dx| > |dy|):  y_k = y0 + sy * floor((2|dy|k + |dx|) / (2|dx|)),  k = 0..|dx|
//                 y-major (otherwise)  :  x_k = x0 + sx * floor((2|dx|k + |dy|) / (2|dy|)),  k = 0..|dy|
//             (cells k < length are misses, k = length is the hit), then every cell packs its
//             ordered miss/hit sequence into one 32-bit record (raw bits up to 26 touches,
//             three run lengths beyond -- the near field is M^a H^b M^c);
//   4. fold   one thread OWNS one cell and applies the records of its tile in scan order with
//             exactly the CPU's IEEE sequence (div/mul/add intrinsics, no FMA), leaving a run early
//             once the value reaches its fixed point (the probability clamp makes runs idempotent).
//
// Nothing depends on beams being angularly sorted: an unsorted scan only widens [kmin, kmax).
// A record that fits neither format (alternating hits and misses in a cell crossed by > 26 beams)
// is marked, and the owner re-derives that (cell, scan) by testing the beams kmin..kmax itself.
#include <cmath>

#include "lgs_internal.cuh"

namespace {

constexpr int kTile = 16;                      // tile side (cells)
constexpr int kTileShift = 4;
constexpr int kTileCells = kTile * kTile;      // = threads per block of the touch / fold passes
constexpr int kChunk = 64;                     // scans per chunk (one mask bit each)
constexpr int kSeg = 128;                      // beams per shared-memory bitmap segment
constexpr int kSegWords = kSeg / 32;
constexpr unsigned kRawMax = 26;               // touches a raw record holds
constexpr unsigned kRecOverflow = 0xFFFFFFFFu;
constexpr unsigned kRecSide = 31u << 26;
constexpr size_t kRecordBudget = (size_t)1 << 30;   // bytes of records per chunk before it is split

struct ScanMeta {
    int sx, sy;          // sensor cell
    int beamBegin, n;    // beams of this scan
    int maxLen;          // max Chebyshev ray length in cells
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

// ---- pre-pass A: sensor cells ---------------------------------------------------------------------
__global__ void integ_sensor_kernel(const double* __restrict__ sensorXY, const int* __restrict__ hitBegin,
                                    int nScans, GridRef g, ScanMeta* __restrict__ meta) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nScans) return;
    ScanMeta m;
    // WorldCoordinateToGridCellIndex (grid_map.hpp:779-790)
    m.sx = __double2int_rd(__ddiv_rn(__dsub_rn(sensorXY[2 * s], g.minX), g.res));
    m.sy = __double2int_rd(__ddiv_rn(__dsub_rn(sensorXY[2 * s + 1], g.minY), g.res));
    m.beamBegin = hitBegin[s];
    m.n = hitBegin[s + 1] - hitBegin[s];
    m.maxLen = 0;
    m.bad = (m.sx < 0 || m.sx >= g.nx || m.sy < 0 || m.sy >= g.ny) ? 1 : 0;
    meta[s] = m;
}

// ---- pre-pass B: per beam end cell relative to the sensor cell ---------------------------------------
__global__ void integ_beam_kernel(const double* __restrict__ hitXY, GridRef g, ScanMeta* __restrict__ meta,
                                  int2* __restrict__ rel) {
    const int s = blockIdx.y;
    const int n = meta[s].n, beamBegin = meta[s].beamBegin, sx = meta[s].sx, sy = meta[s].sy;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (i < n) {
        const size_t b = (size_t)beamBegin + i;
        const double2 h = reinterpret_cast<const double2*>(hitXY)[b];
        const int ex = __double2int_rd(__ddiv_rn(__dsub_rn(h.x, g.minX), g.res));
        const int ey = __double2int_rd(__ddiv_rn(__dsub_rn(h.y, g.minY), g.res));
        if (ex < 0 || ex >= g.nx || ey < 0 || ey >= g.ny) atomicOr(&meta[s].bad, 1);
        const int dx = ex - sx, dy = ey - sy;
        rel[b] = make_int2(dx, dy);
        len = max(abs(dx), abs(dy));
    }
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(&meta[s].maxLen, len);
}

// ---- 1. mark: beam index range per (tile, scan) ------------------------------------------------------
// One thread per (scan, beam, 16-step piece of the ray); consecutive lanes = consecutive beams at
// the same distance from the sensor, so a warp's atomics mostly hit the same few addresses.
// (x0, y0) = tile-aligned origin of the chunk's region, tw = tiles per region row.  The racy
// pre-reads only save atomics: kmin only ever decreases and kmax only ever increases.
__global__ void __launch_bounds__(128)
integ_mark_kernel(const ScanMeta* __restrict__ meta, const int2* __restrict__ rel, int x0, int y0, int tw,
                  int beamsPad, unsigned* __restrict__ kmin, unsigned* __restrict__ kmax) {
    const int s = blockIdx.y;
    const ScanMeta m = meta[s];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t % beamsPad, piece = t / beamsPad;
    if (i >= m.n) return;
    const int2 e = __ldg(rel + m.beamBegin + i);
    const int ax = abs(e.x), ay = abs(e.y);
    const bool xMajor = ax > ay;                            // util.hpp:276 vs :288
    const int amaj = xMajor ? ax : ay, amin = xMajor ? ay : ax;
    const int kBegin = piece * kTile;
    if (kBegin > amaj) return;
    const int kEnd = min(kBegin + kTile - 1, amaj);
    const int sMaj = (xMajor ? e.x : e.y) < 0 ? -1 : 1, sMin = (xMajor ? e.y : e.x) < 0 ? -1 : 1;
    const int bx = m.sx - x0, by = m.sy - y0;
    int last = -1;
    for (int k = kBegin; k <= kEnd; ++k) {
        const int minor = amaj ? (int)((unsigned)(2 * amin * k + amaj) / (unsigned)(2 * amaj)) : 0;
        const int x = bx + (xMajor ? sMaj * k : sMin * minor);
        const int y = by + (xMajor ? sMin * minor : sMaj * k);
        const int tile = (y >> kTileShift) * tw + (x >> kTileShift);
        if (tile == last) continue;
        last = tile;
        const size_t idx = (size_t)tile * kChunk + s;
        if (__ldcg(kmin + idx) > (unsigned)i) atomicMin(kmin + idx, (unsigned)i);
        if (__ldcg(kmax + idx) < (unsigned)i + 1u) atomicMax(kmax + idx, (unsigned)i + 1u);
    }
}

// ---- 2. pairs: scan-ordered (tile, scan, k0, k1) runs; resets the mark buffers -------------------------
// One warp per tile: lanes read the tile's 64 kmax entries (non-zero = the scan touches the tile).
__global__ void __launch_bounds__(128)
integ_pairs_kernel(int nTiles, unsigned* __restrict__ kmin, unsigned* __restrict__ kmax,
                   uint2* __restrict__ tileInfo, int4* __restrict__ pairs, unsigned* __restrict__ nPairs) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= nTiles) return;
    const size_t idx = (size_t)t * kChunk + lane;
    const unsigned hiA = kmax[idx], hiB = kmax[idx + 32];
    const unsigned mA = __ballot_sync(0xffffffffu, hiA != 0u), mB = __ballot_sync(0xffffffffu, hiB != 0u);
    const unsigned cnt = (unsigned)(__popc(mA) + __popc(mB));
    unsigned base = 0;
    if (cnt && lane == 0) base = atomicAdd(nPairs, cnt);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane == 0) tileInfo[t] = make_uint2(base, cnt);
    const unsigned below = (1u << lane) - 1u;
    if (hiA) {
        pairs[base + __popc(mA & below)] = make_int4(t, lane, (int)kmin[idx], (int)hiA);
        kmin[idx] = 0xFFFFFFFFu; kmax[idx] = 0u;
    }
    if (hiB) {
        pairs[base + __popc(mA) + __popc(mB & below)] = make_int4(t, lane + 32, (int)kmin[idx + 32], (int)hiB);
        kmin[idx + 32] = 0xFFFFFFFFu; kmax[idx + 32] = 0u;
    }
}

// ---- 3. touch: ordered miss/hit sequence of every cell of a (tile, scan) pair ----------------------------
// Record formats (32 bit):   0                      no touch
//   raw  [31] = 0, [30:26] = count (1..26), [25:0] = types in order (1 = hit)
//   runs [31] = 1, [30] = type of the first run, [29:20] [19:10] [9:0] = three alternating run lengths
//   side [31] = 0, [30:26] = 31, [25:0] = offset (units of 8 words) of the cell's raw touch / hit
//        bitmap words in the side buffer: sequences that fit neither format (a near cell whose
//        beams alternate between ending in it and passing through)
//   0xFFFFFFFF               side buffer exhausted: the owner re-derives the sequence from the beams
struct TouchSeq {
    unsigned cnt = 0, raw = 0, r0 = 0, r1 = 0, r2 = 0;
    int nrun = 0, cur = -1, first = 0;
    __device__ __forceinline__ void append(int type, unsigned n) {
        if (cnt < kRawMax && type) raw |= (n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u)) << cnt;
        cnt += n;
        if (type != cur) { cur = type; if (++nrun == 1) first = type; }
        if (nrun == 1) r0 += n; else if (nrun == 2) r1 += n; else if (nrun == 3) r2 += n;
    }
    __device__ __forceinline__ unsigned record() const {
        if (cnt == 0) return 0u;
        if (cnt <= kRawMax) return (cnt << 26) | (raw & ((1u << 26) - 1u));
        if (nrun <= 3 && r0 < 1023u && r1 < 1023u && r2 < 1023u)
            return 0x80000000u | ((unsigned)first << 30) | (r0 << 20) | (r1 << 10) | r2;
        return kRecOverflow;
    }
};

__global__ void __launch_bounds__(kTileCells)
integ_touch_kernel(const ScanMeta* __restrict__ meta, const int2* __restrict__ rel,
                   const int4* __restrict__ pairs, const unsigned* __restrict__ nPairs, int x0, int y0,
                   int tw, unsigned* __restrict__ records, unsigned* __restrict__ side,
                   unsigned* __restrict__ sideCursor, unsigned sideCap,
                   unsigned long long* __restrict__ counters) {
    __shared__ unsigned sTouch[kSegWords * kTileCells];     // [word][cell]
    __shared__ unsigned sHit[kSegWords * kTileCells];
    __shared__ unsigned sSum[kTileCells / 32];
    const int tid = threadIdx.x;
    const unsigned nP = *nPairs;
    unsigned total = 0, overflow = 0;
    for (unsigned p = blockIdx.x; p < nP; p += gridDim.x) {
        const int4 pr = pairs[p];
        const ScanMeta m = meta[pr.y];
        const int ox = x0 + (pr.x % tw) * kTile - m.sx;     // tile origin relative to the sensor cell
        const int oy = y0 + (pr.x / tw) * kTile - m.sy;
        const int2* __restrict__ E = rel + m.beamBegin;
        TouchSeq seq;
        unsigned rec = 0u, sideOff = 0u;
        // pass 0 builds the records; pass 1 (only if a cell of this pair fits no record format)
        // repeats the bitmaps and streams the raw words of those cells to the side buffer
        for (int pass = 0; pass < 2; ++pass) {
            for (int seg = pr.z; seg < pr.w; seg += kSeg) {
#pragma unroll
                for (int w = 0; w < kSegWords; ++w) { sTouch[w * kTileCells + tid] = 0u; sHit[w * kTileCells + tid] = 0u; }
                __syncthreads();
                const int nb = min(kSeg, pr.w - seg);
                for (int it = tid; it < nb * kTile; it += kTileCells) {
                    const int b = it >> kTileShift, j = it & (kTile - 1);
                    const int2 e = __ldg(E + seg + b);
                    const int ax = abs(e.x), ay = abs(e.y);
                    const bool xMajor = ax > ay;
                    const int amaj = xMajor ? ax : ay, amin = xMajor ? ay : ax;
                    const int eMaj = xMajor ? e.x : e.y, eMin = xMajor ? e.y : e.x;
                    const int oMaj = xMajor ? ox : oy, oMin = xMajor ? oy : ox;
                    const int rMaj = oMaj + j;                               // this item's column (row)
                    const int k = eMaj < 0 ? -rMaj : rMaj;                   // step along the major axis
                    if (k < 0 || k > amaj) continue;
                    const int minor = amaj ? (int)((unsigned)(2 * amin * k + amaj) / (unsigned)(2 * amaj)) : 0;
                    const int lMin = (eMin < 0 ? -minor : minor) - oMin;
                    if ((unsigned)lMin >= (unsigned)kTile) continue;
                    const int cell = xMajor ? (lMin * kTile + j) : (j * kTile + lMin);
                    const unsigned bitb = 1u << (b & 31);
                    atomicOr(&sTouch[(b >> 5) * kTileCells + cell], bitb);
                    if (k == amaj) atomicOr(&sHit[(b >> 5) * kTileCells + cell], bitb);
                }
                __syncthreads();
                if (pass == 0) {
#pragma unroll
                    for (int w = 0; w < kSegWords; ++w) {
                        unsigned t = sTouch[w * kTileCells + tid];
                        if (t == 0u) continue;
                        const unsigned h = sHit[w * kTileCells + tid];
                        while (t) {                              // maximal runs of equal type, ascending beams
                            const int type = (h >> (__ffs(t) - 1)) & 1;
                            const unsigned same = type ? (t & h) : (t & ~h);
                            const unsigned other = t & ~same;
                            const unsigned upto = other ? ((1u << (__ffs(other) - 1)) - 1u) : 0xFFFFFFFFu;
                            const unsigned run = same & upto;
                            seq.append(type, (unsigned)__popc(run));
                            t &= ~run;
                        }
                    }
                } else if ((rec >> 26) == 31u) {
                    unsigned* out = side + (size_t)sideOff + (size_t)((seg - pr.z) / kSeg) * (2 * kSegWords);
#pragma unroll
                    for (int w = 0; w < kSegWords; ++w) {
                        out[2 * w] = sTouch[w * kTileCells + tid];
                        out[2 * w + 1] = sHit[w * kTileCells + tid];
                    }
                }
                __syncthreads();
            }
            if (pass == 0) {
                rec = seq.record();
                if (rec == kRecOverflow) {
                    const unsigned words = (unsigned)((pr.w - pr.z + kSeg - 1) / kSeg) * (2 * kSegWords);
                    sideOff = atomicAdd(sideCursor, words);
                    if (sideOff + words <= sideCap) rec = kRecSide | (sideOff >> 3);
                }
                if (!__syncthreads_or((rec >> 26) == 31u)) break;
            }
        }
        records[(size_t)p * kTileCells + tid] = rec;
        total += seq.cnt;
        overflow += rec == kRecOverflow ? 1u : 0u;
    }
    // one atomic per block: touches (= Update calls of the CPU) and overflowed records
    unsigned v = total | 0u;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    unsigned ov = overflow;
    for (int o = 16; o > 0; o >>= 1) ov += __shfl_down_sync(0xffffffffu, ov, o);
    if ((tid & 31) == 0) sSum[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        unsigned long long tot = 0;
        for (int k = 0; k < kTileCells / 32; ++k) tot += sSum[k];
        if (tot) atomicAdd(counters, tot);
    }
    if ((tid & 31) == 0 && ov) atomicAdd(counters + 1, (unsigned long long)ov);
}

// ---- 4. fold: one thread owns one cell ---------------------------------------------------------------
struct FoldArgs {
    const ScanMeta* meta;
    const int2* rel;
    const uint2* tileInfo;
    const int4* pairs;
    const unsigned* records;
    const unsigned* side;
    double pHit, pMiss, oddsHit, oddsMiss;
};

__device__ __forceinline__ double applyRun(double v, bool hit, unsigned n, const FoldArgs& a) {
    const double p = hit ? a.pHit : a.pMiss, odds = hit ? a.oddsHit : a.oddsMiss;
    for (unsigned j = 0; j < n; ++j) {
        const double nv = bayesUpdate(v, p, odds);
        if (nv == v) break;                // fixed point: the rest of the run is a no-op
        v = nv;
    }
    return v;
}

__global__ void __launch_bounds__(kTileCells)
integ_fold_kernel(FoldArgs a, GridRef g, int x0, int y0, int tw) {
    const uint2 info = a.tileInfo[blockIdx.x];
    if (info.y == 0u) return;
    const int tid = threadIdx.x;
    const int cx = x0 + (blockIdx.x % tw) * kTile + (tid & (kTile - 1));
    const int cy = y0 + (blockIdx.x / tw) * kTile + (tid >> kTileShift);
    const bool inside = cx < g.nx && cy < g.ny;
    double* cell = g.origin + (size_t)cy * g.pitch + cx;
    const double v0 = inside ? *cell : 0.0;
    double v = v0;
    const unsigned* __restrict__ R = a.records + (size_t)info.x * kTileCells + tid;
    unsigned next = __ldcs(R);
    for (unsigned j = 0; j < info.y; ++j) {
        const unsigned rec = next;
        if (j + 1 < info.y) next = __ldcs(R + (size_t)(j + 1) * kTileCells);
        if (rec == 0u) continue;
        if (rec == kRecOverflow) {
            const int4 pr = a.pairs[info.x + j];
            const ScanMeta m = a.meta[pr.y];
            const int2* __restrict__ E = a.rel + m.beamBegin;
            const int rx = cx - m.sx, ry = cy - m.sy;
            for (int i = pr.z; i < pr.w; ++i) {
                const int2 e = __ldg(E + i);
                const int ty = rayTouch(rx, ry, e.x, e.y);
                if (ty) v = ty == 2 ? bayesUpdate(v, a.pHit, a.oddsHit) : bayesUpdate(v, a.pMiss, a.oddsMiss);
            }
        } else if ((rec >> 26) == 31u) {
            const int4 pr = a.pairs[info.x + j];
            const unsigned* __restrict__ S = a.side + (size_t)(rec & ((1u << 26) - 1u)) * 8u;
            const int nW = ((pr.w - pr.z + kSeg - 1) / kSeg) * kSegWords;
            for (int w = 0; w < nW; ++w) {
                unsigned t = S[2 * w];
                const unsigned h = S[2 * w + 1];
                while (t) {
                    const bool hit = (h >> (__ffs(t) - 1)) & 1u;
                    const unsigned same = hit ? (t & h) : (t & ~h);
                    const unsigned other = t & ~same;
                    const unsigned upto = other ? ((1u << (__ffs(other) - 1)) - 1u) : 0xFFFFFFFFu;
                    const unsigned run = same & upto;
                    v = applyRun(v, hit, (unsigned)__popc(run), a);
                    t &= ~run;
                }
            }
        } else if (rec >> 31) {
            const bool first = (rec >> 30) & 1u;
            v = applyRun(v, first, (rec >> 20) & 1023u, a);
            v = applyRun(v, !first, (rec >> 10) & 1023u, a);
            v = applyRun(v, first, rec & 1023u, a);
        } else {
            unsigned cnt = rec >> 26, bits = rec & ((1u << 26) - 1u);
            while (cnt) {
                const bool hit = bits & 1u;
                const unsigned diff = hit ? ~bits : bits;        // first touch of the other type
                const unsigned n = min(cnt, diff ? (unsigned)(__ffs(diff) - 1) : 32u);
                v = applyRun(v, hit, n, a);
                bits >>= n; cnt -= n;
            }
        }
    }
    if (inside && v != v0) *cell = v;
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

    // Stage the whole batch once; the passes then run over chunks of <= 64 scans (one mask bit each).
    LGS_CUDA(c, w.sensor.reserve((size_t)n * 2));
    LGS_CUDA(c, w.hit.reserve(std::max<size_t>((size_t)total, 1) * 2));
    LGS_CUDA(c, w.begin.reserve((size_t)n + 1));
    LGS_CUDA(c, w.meta.reserve((size_t)n * sizeof(ScanMeta)));
    LGS_CUDA(c, w.rel.reserve(std::max<size_t>((size_t)total, 1)));
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
    ScanMeta* dMeta = reinterpret_cast<ScanMeta*>(w.meta.p);
    ScanMeta* hMeta = reinterpret_cast<ScanMeta*>(w.hMeta.p);
    GridRef g{grid->origin(), grid->nx, grid->ny, grid->pitch, grid->min_x, grid->min_y, grid->res};
    // Pre-pass over the whole batch: sensor cells, relative end cells, ray lengths, bounds check.
    integ_sensor_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(w.sensor.p, w.begin.p, n, g, dMeta);
    LGS_LAUNCH_CHECK(c);
    if (maxBeams > 0) {
        dim3 gb((maxBeams + 127) / 128, n);
        integ_beam_kernel<<<gb, 128, 0, c->stream>>>(w.hit.p, g, dMeta, w.rel.p);
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

    // Side buffer for touch sequences that fit no record format (words; LGS_INTEG_SIDE_WORDS is a test
    // hook that shrinks it to exercise the exhaustive fallback).
    size_t sideCap = (size_t)8 << 20;
    if (const char* e = getenv("LGS_INTEG_SIDE_WORDS")) sideCap = (size_t)std::max(8, atoi(e));
    // tiles of scan s's reach: the square of half side maxLen around the sensor cell
    auto tileSpan = [](int lo, int hi) { return (hi >> kTileShift) - (lo >> kTileShift) + 1; };
    int s0 = 0;
    while (s0 < n) {
        // grow the chunk while it stays within 64 scans and the record budget
        int x0 = grid->nx, y0 = grid->ny, x1 = -1, y1 = -1, subBeams = 0, subLen = 0, ns = 0;
        size_t pairBound = 0;
        while (s0 + ns < n && ns < kChunk) {
            const ScanMeta& m = hMeta[s0 + ns];
            if (m.n > 0) {
                const int lx = std::max(m.sx - m.maxLen, 0), hx = std::min(m.sx + m.maxLen, grid->nx - 1);
                const int ly = std::max(m.sy - m.maxLen, 0), hy = std::min(m.sy + m.maxLen, grid->ny - 1);
                const size_t pb = (size_t)tileSpan(lx, hx) * tileSpan(ly, hy);
                if (ns > 0 && (pairBound + pb) * kTileCells * sizeof(unsigned) > kRecordBudget) break;
                pairBound += pb;
                x0 = std::min(x0, lx); x1 = std::max(x1, hx);
                y0 = std::min(y0, ly); y1 = std::max(y1, hy);
                subBeams = std::max(subBeams, m.n);
                subLen = std::max(subLen, m.maxLen);
            }
            ++ns;
        }
        const int cs = s0;
        s0 += ns;
        if (subBeams == 0 || x1 < x0 || y1 < y0) continue;
        x0 &= ~(kTile - 1); y0 &= ~(kTile - 1);                       // tile-aligned region origin
        const int tw = tileSpan(x0, x1), th = tileSpan(y0, y1);
        const size_t nTiles = (size_t)tw * th;
        if (nTiles * kChunk >= ((size_t)1 << 31)) return lgs_fail(c, LGS_ERR_INVALID, "integrate: region of %dx%d tiles is too large", tw, th);
        pairBound = std::min(pairBound, nTiles * ns);

        // mark buffers: kept in their reset state by the pair pass; (re)initialise what is new
        if (w.dirty) { w.cleanTiles = 0; w.dirty = false; }
        if (nTiles * kChunk > w.kmin.cap) w.cleanTiles = 0;           // reserve() reallocates
        LGS_CUDA(c, w.kmin.reserve(nTiles * kChunk));
        LGS_CUDA(c, w.kmax.reserve(nTiles * kChunk));
        LGS_CUDA(c, w.tileInfo.reserve(nTiles));
        LGS_CUDA(c, w.pairs.reserve(pairBound));
        LGS_CUDA(c, w.records.reserve(pairBound * kTileCells));
        if (nTiles > w.cleanTiles) {
            const size_t a0 = w.cleanTiles, cnt = w.kmin.cap / kChunk - a0;   // initialise up to the capacity
            LGS_CUDA(c, cudaMemsetAsync(w.kmin.p + a0 * kChunk, 0xFF, cnt * kChunk * sizeof(unsigned), c->stream));
            LGS_CUDA(c, cudaMemsetAsync(w.kmax.p + a0 * kChunk, 0, cnt * kChunk * sizeof(unsigned), c->stream));
            w.cleanTiles = w.kmin.cap / kChunk;
        }
        LGS_CUDA(c, w.side.reserve(sideCap));
        unsigned* nPairs = reinterpret_cast<unsigned*>(w.counters.p + 4);   // [0] pairs, [1] side-buffer cursor
        LGS_CUDA(c, cudaMemsetAsync(nPairs, 0, 2 * sizeof(unsigned), c->stream));

        w.dirty = true;
        const int beamsPad = (subBeams + 31) & ~31, pieces = subLen / kTile + 1;
        dim3 gm((unsigned)(((size_t)beamsPad * pieces + 127) / 128), ns);
        integ_mark_kernel<<<gm, 128, 0, c->stream>>>(dMeta + cs, w.rel.p, x0, y0, tw, beamsPad, w.kmin.p, w.kmax.p);
        LGS_LAUNCH_CHECK(c);
        integ_pairs_kernel<<<(unsigned)((nTiles * 32 + 127) / 128), 128, 0, c->stream>>>((int)nTiles, w.kmin.p, w.kmax.p,
                                                                                       w.tileInfo.p, w.pairs.p, nPairs);
        LGS_LAUNCH_CHECK(c);
        w.dirty = false;
        const unsigned touchBlocks = (unsigned)std::min<size_t>(pairBound, (size_t)c->sm_count * 8);
        integ_touch_kernel<<<touchBlocks, kTileCells, 0, c->stream>>>(dMeta + cs, w.rel.p, w.pairs.p, nPairs, x0, y0, tw,
                                                                      w.records.p, w.side.p, nPairs + 1, (unsigned)sideCap,
                                                                      w.counters.p);
        LGS_LAUNCH_CHECK(c);
        FoldArgs a{dMeta + cs, w.rel.p, w.tileInfo.p, w.pairs.p, w.records.p, w.side.p, pHit, pMiss, oddsHit, oddsMiss};
        integ_fold_kernel<<<(unsigned)nTiles, kTileCells, 0, c->stream>>>(a, g, x0, y0, tw);
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
