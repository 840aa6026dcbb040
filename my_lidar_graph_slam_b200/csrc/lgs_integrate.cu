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

#include <chrono>

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
constexpr int kFoldBatch = 8;                  // records a fold thread keeps in flight
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

// Minor-axis offset of step k of a ray: floor((2 amin k + amaj) / (2 amaj)) (util.hpp:276-299), amaj > 0.
// For rays up to 2047 cells the dividend is below 2^24, i.e. exact in float: one reciprocal, one
// multiply and an integer correction replace the 32-bit division; longer rays divide.
__device__ __forceinline__ int minorAt(int amin, int amaj, int k) {
    const int lhs = 2 * amin * k + amaj, m2 = 2 * amaj;
    if (amaj > 2047) return (int)((unsigned)lhs / (unsigned)m2);
    int q = __float2int_rz(__fmul_rn((float)lhs, __frcp_rn((float)m2)));
    const int r = lhs - q * m2;
    if (r < 0) --q; else if (r >= m2) ++q;
    return q;
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
// One thread per (scan, beam, 16-step piece of the ray); a warp = 32 consecutive beams at the same
// distance from the sensor, which mostly cross the same tiles.  A 16-step piece of a (monotone)
// Bresenham path visits at most 3 tiles; for each of them the warp groups the lanes that share the
// tile and issues ONE atomicMin (lowest beam of the group) and ONE atomicMax (highest beam).
// (x0, y0) = tile-aligned origin of the chunk's region, tw = tiles per region row.
__global__ void __launch_bounds__(128)
integ_mark_kernel(const ScanMeta* __restrict__ meta, const int2* __restrict__ rel, int x0, int y0, int tw,
                  int beamsPad, unsigned* __restrict__ kmin, unsigned* __restrict__ kmax) {
    const int s = blockIdx.y;
    const ScanMeta m = meta[s];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;        // beamsPad is a multiple of 32:
    const int i = t % beamsPad, piece = t / beamsPad;           // a warp never straddles two pieces
    const int lane = threadIdx.x & 31;
    int tiles[3] = {-1, -1, -1};
    if (i < m.n) {
        const int2 e = __ldg(rel + m.beamBegin + i);
        const int ax = abs(e.x), ay = abs(e.y);
        const bool xMajor = ax > ay;                            // util.hpp:276 vs :288
        const int amaj = xMajor ? ax : ay, amin = xMajor ? ay : ax;
        const int kBegin = piece * kTile;
        if (kBegin <= amaj) {
            const int kEnd = min(kBegin + kTile - 1, amaj);
            const int sMaj = (xMajor ? e.x : e.y) < 0 ? -1 : 1, sMin = (xMajor ? e.y : e.x) < 0 ? -1 : 1;
            const int bx = m.sx - x0, by = m.sy - y0;
            const int m2 = 2 * amaj, d2 = 2 * amin;
            int minor = amaj ? minorAt(amin, amaj, kBegin) : 0;
            int rem = d2 * kBegin + amaj - minor * m2;
            int nt = 0, last = -1;
            for (int k = kBegin; k <= kEnd; ++k) {
                const int x = bx + (xMajor ? sMaj * k : sMin * minor);
                const int y = by + (xMajor ? sMin * minor : sMaj * k);
                const int tile = (y >> kTileShift) * tw + (x >> kTileShift);
                if (tile != last) {
                    last = tile;
                    if (nt == 0) tiles[0] = tile; else if (nt == 1) tiles[1] = tile; else tiles[2] = tile;
                    ++nt;
                }
                rem += d2;
                if (rem >= m2) { rem -= m2; ++minor; }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int mine = tiles[u];
        unsigned pending = __ballot_sync(0xffffffffu, mine >= 0);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const int tl = __shfl_sync(0xffffffffu, mine, leader);
            const unsigned grp = __ballot_sync(0xffffffffu, mine == tl);
            if (lane == leader) {
                const size_t idx = (size_t)tl * kChunk + s;
                atomicMin(kmin + idx, (unsigned)i);                             // lowest lane = lowest beam
                atomicMax(kmax + idx, (unsigned)(i + (31 - __clz(grp)) - lane) + 1u);
            }
            pending &= ~grp;
        }
    }
}

// ---- 2. pairs: scan-ordered pair descriptors per tile; resets the mark buffers --------------------------
// Descriptor = two int4: {ox, oy, k0, k1} (tile origin relative to the scan's sensor cell, beam range)
// and {beamBegin, tile, scan, 0}, so the touch pass needs no further dependent look-ups.
// One warp per tile: lanes read the tile's 64 kmax entries (non-zero = the scan touches the tile).
__global__ void __launch_bounds__(128)
integ_pairs_kernel(int nTiles, const ScanMeta* __restrict__ meta, int x0, int y0, int tw,
                   unsigned* __restrict__ kmin, unsigned* __restrict__ kmax,
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
    const int tx = x0 + (t % tw) * kTile, ty = y0 + (t / tw) * kTile;
    auto emit = [&](unsigned slot, int s, unsigned lo, unsigned hi) {
        const ScanMeta m = meta[s];
        pairs[2 * (size_t)slot] = make_int4(tx - m.sx, ty - m.sy, (int)lo, (int)hi);
        pairs[2 * (size_t)slot + 1] = make_int4(m.beamBegin, t, s, 0);
    };
    if (hiA) {
        emit(base + __popc(mA & below), lane, kmin[idx], hiA);
        kmin[idx] = 0xFFFFFFFFu; kmax[idx] = 0u;
    }
    if (hiB) {
        emit(base + __popc(mA) + __popc(mB & below), lane + 32, kmin[idx + 32], hiB);
        kmin[idx + 32] = 0xFFFFFFFFu; kmax[idx + 32] = 0u;
    }
}

// ---- 3. touch: ordered miss/hit sequence of every cell of a (tile, scan) pair ----------------------------
// Record formats (32 bit):   0                      no touch
//   raw  [31] = 0, [30:26] = count (1..26), [25:0] = types in order (1 = hit)
//   runs [31] = 1, [30] = type of the first run, [29:20] [19:10] [9:0] = three alternating run lengths
//   side [31] = 0, [30:26] = 31, [25:0] = offset (units of 8 words) of the cell's non-empty raw
//        touch / hit bitmap words in the side buffer: sequences that fit neither format (a near cell whose
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

__global__ void __launch_bounds__(kTileCells, 4)
integ_touch_kernel(const int2* __restrict__ rel, const int4* __restrict__ pairs,
                   const unsigned* __restrict__ nPairs, unsigned* __restrict__ records, unsigned* __restrict__ side,
                   unsigned* __restrict__ sideCursor, unsigned sideCap,
                   unsigned long long* __restrict__ counters) {
    __shared__ unsigned sBits[2][2 * kSegWords * kTileCells];   // double buffered; [touch | hit][word][cell]
    int buf = 0;
    __shared__ unsigned sSum[kTileCells / 32];
    const int tid = threadIdx.x;
    const unsigned nP = *nPairs;
    unsigned total = 0, overflow = 0;
    int4 nextA = make_int4(0, 0, 0, 0), nextB = nextA;
    // Threads per beam in the scatter step: 4 (4 steps each) for segments of <= 64 beams, else 2.
    auto lanesPerBeam = [](int nb) { return nb <= kTileCells / 4 ? 4 : 2; };
    int2 eNext = make_int2(0, 0);                            // this thread's beam of the next pair's first segment
    if (blockIdx.x < nP) {
        nextA = pairs[2 * (size_t)blockIdx.x]; nextB = pairs[2 * (size_t)blockIdx.x + 1];
        const int nb = min(kSeg, nextA.w - nextA.z), b = tid / lanesPerBeam(nb);
        if (b < nb) eNext = __ldg(rel + nextB.x + nextA.z + b);
    }
    for (unsigned p = blockIdx.x; p < nP; p += gridDim.x) {
        const int4 pr = nextA;                              // .x/.y tile origin - sensor cell, .z/.w beams
        const int2* __restrict__ E = rel + nextB.x;
        const int2 eFirst = eNext;
        const bool more = p + gridDim.x < nP;
        if (more) {                                         // the next descriptor is in flight meanwhile
            nextA = pairs[2 * (size_t)(p + gridDim.x)];
            nextB = pairs[2 * (size_t)(p + gridDim.x) + 1];
        }
        bool fetched = false;
        const int ox = pr.x, oy = pr.y;
        TouchSeq seq;
        unsigned rec = 0u, sideOff = 0u;
        int wFirst = 0x7fffffff, wLast = -1;                   // non-empty bitmap words of this cell
        // pass 0 builds the records; pass 1 (only if a cell of this pair fits no record format)
        // repeats the bitmaps and streams the raw words of those cells to the side buffer
        for (int pass = 0; pass < 2; ++pass) {
            for (int seg = pr.z; seg < pr.w; seg += kSeg) {
                // Two barriers per segment: the buffer cleared here was last read two segments ago.
                buf ^= 1;
                unsigned* sTouch = sBits[buf];
                unsigned* sHit = sBits[buf] + kSegWords * kTileCells;
                const int nW = (min(kSeg, pr.w - seg) + 31) >> 5;
                for (int w = 0; w < nW; ++w) { sTouch[w * kTileCells + tid] = 0u; sHit[w * kTileCells + tid] = 0u; }
                __syncthreads();
                const int nb = min(kSeg, pr.w - seg);
                const int per = lanesPerBeam(nb), b = tid / per;
                if (b < nb) {
                    // `per` threads per beam share the <= 16 steps whose major coordinate lies inside
                    // the tile; the minor coordinate floor((2 amin k + amaj) / (2 amaj)) is divided
                    // out once per thread and then carried with its remainder (2 amin <= 2 amaj: at
                    // most one increment a step)
                    const int2 e = (pass == 0 && seg == pr.z) ? eFirst : __ldg(E + seg + b);
                    const int ax = abs(e.x), ay = abs(e.y);
                    const bool xMajor = ax > ay;
                    const int amaj = xMajor ? ax : ay, amin = xMajor ? ay : ax;
                    const int eMaj = xMajor ? e.x : e.y, eMin = xMajor ? e.y : e.x;
                    const int oMaj = xMajor ? ox : oy, oMin = xMajor ? oy : ox;
                    const int steps = kTile / per, sub = tid - b * per;
                    const int kBase = (eMaj >= 0 ? oMaj : -oMaj - (kTile - 1)) + sub * steps;
                    const int kLo = max(kBase, 0);
                    const int kHi = min(kBase + steps - 1, amaj);
                    if (kLo <= kHi) {
                        const int m2 = 2 * amaj, d2 = 2 * amin;
                        int minor = amaj ? minorAt(amin, amaj, kLo) : 0;
                        int rem = d2 * kLo + amaj - minor * m2;
                        const unsigned bitb = 1u << (b & 31);
                        unsigned* tw_ = sTouch + (b >> 5) * kTileCells;
                        unsigned* hw_ = sHit + (b >> 5) * kTileCells;
                        for (int k = kLo; k <= kHi; ++k) {
                            const int j = eMaj >= 0 ? k - oMaj : -k - oMaj;              // column (row) in the tile
                            const int lMin = (eMin < 0 ? -minor : minor) - oMin;
                            if ((unsigned)lMin < (unsigned)kTile) {
                                const int cell = xMajor ? (lMin * kTile + j) : (j * kTile + lMin);
                                atomicOr(tw_ + cell, bitb);
                                if (k == amaj) atomicOr(hw_ + cell, bitb);
                            }
                            rem += d2;
                            if (rem >= m2) { rem -= m2; ++minor; }
                        }
                    }
                }
                if (more && !fetched) {                     // next pair's beams: in flight during the extraction
                    fetched = true;
                    const int nbN = min(kSeg, nextA.w - nextA.z), bN = tid / lanesPerBeam(nbN);
                    if (bN < nbN) eNext = __ldg(rel + nextB.x + nextA.z + bN);
                }
                __syncthreads();
                if (pass == 0) {
                    for (int w = 0; w < nW; ++w) {
                        unsigned t = sTouch[w * kTileCells + tid];
                        if (t == 0u) continue;
                        const int gw = ((seg - pr.z) / kSeg) * kSegWords + w;
                        wFirst = min(wFirst, gw); wLast = gw;
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
                    unsigned* out = side + (size_t)sideOff + 2;
                    for (int w = 0; w < nW; ++w) {
                        const int gw = ((seg - pr.z) / kSeg) * kSegWords + w;
                        if (gw < wFirst || gw > wLast) continue;
                        out[2 * (gw - wFirst)] = sTouch[w * kTileCells + tid];
                        out[2 * (gw - wFirst) + 1] = sHit[w * kTileCells + tid];
                    }
                }
            }
            if (pass == 0) {
                rec = seq.record();
                if (rec == kRecOverflow) {
                    // header {first word | words << 16, -} + (touch, hit) word pairs, in units of 8 words
                    const unsigned nWords = (unsigned)(wLast - wFirst + 1);
                    const unsigned words = (2u + 2u * nWords + 7u) & ~7u;
                    sideOff = atomicAdd(sideCursor, words);
                    if (sideOff + words <= sideCap && nWords < 65536u) {
                        rec = kRecSide | (sideOff >> 3);
                        side[sideOff] = (unsigned)wFirst | (nWords << 16);
                    }
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
    const int2* rel;
    const uint2* tileInfo;
    const int4* pairs;
    const unsigned* records;
    const unsigned* side;
    unsigned long long* diag;      // LGS_INTEG_TIMING: [2] max / [3] sum of computed updates per thread
    double pHit, pMiss, oddsHit, oddsMiss;
};

// One thread owns one cell.  Raw records (the common case) are queued per lane and applied only when
// some lane of the warp is about to overflow its 64-touch queue: neighbouring cells see similar
// numbers of touches over a chunk, so the lanes of a warp then have similar amounts of work, while
// applying every record at once would run the ~100-instruction update under heavy divergence.
// A touch that cannot change the value (cell already at the clamp the observation pushes towards:
// the clamp makes the update idempotent there) is skipped without arithmetic.
struct Folder {
    const FoldArgs& a;
    double v;
    unsigned long long q = 0ull;
    int qn = 0;
    bool missSat, hitSat;
    unsigned computed = 0;
    __device__ __forceinline__ Folder(const FoldArgs& args, double v0) : a(args), v(v0) {
        const double lo = 1e-3, hi = 1.0 - 1e-3;
        missSat = bayesUpdate(lo, a.pMiss, a.oddsMiss) == lo;
        hitSat = bayesUpdate(hi, a.pHit, a.oddsHit) == hi;
    }
    __device__ __forceinline__ bool saturated(bool hit) const {
        return hit ? (hitSat && v == 1.0 - 1e-3) : (missSat && v == 1e-3);
    }
    __device__ __forceinline__ void touch(bool hit) {
        if (!saturated(hit)) { v = bayesUpdate(v, hit ? a.pHit : a.pMiss, hit ? a.oddsHit : a.oddsMiss); ++computed; }
    }
    // Every iteration performs exactly one COMPUTED update per lane: a lane first drops, without
    // iterating, the leading touches that cannot change a value sitting on a clamp, so a warp
    // iterates max-over-lanes(computed updates) times, not max-over-lanes(queue length) times.
    __device__ __forceinline__ void drain() {
        while (qn) {
            const bool atLo = missSat && v == 1e-3, atHi = hitSat && v == 1.0 - 1e-3;
            if (atLo || atHi) {
                const unsigned long long x = atLo ? q : ~q;            // first touch of the other type
                const int n = min(qn, x ? __ffsll((long long)x) - 1 : 64);
                q = n >= 64 ? 0ull : q >> n;
                qn -= n;
                if (!qn) break;
            }
            const bool hit = q & 1ull;
            v = bayesUpdate(v, hit ? a.pHit : a.pMiss, hit ? a.oddsHit : a.oddsMiss);
            ++computed;
            q >>= 1; --qn;
        }
    }
    __device__ __forceinline__ void run(bool hit, unsigned n) {
        for (unsigned j = 0; j < n; ++j) {
            if (saturated(hit)) break;
            const double nv = bayesUpdate(v, hit ? a.pHit : a.pMiss, hit ? a.oddsHit : a.oddsMiss);
            ++computed;
            if (nv == v) break;            // fixed point: the rest of the run is a no-op
            v = nv;
        }
    }
};

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
    Folder f(a, v0);
    // Records are prefetched kFoldBatch pairs at a time (they do not depend on the cell value): the
    // loads of the next batch are in flight while this thread's own column of the shared staging
    // buffer is consumed, so a tile touched by all 64 scans pays 8 memory round trips, not 64.
    __shared__ unsigned sRec[kFoldBatch * kTileCells];
    const unsigned* __restrict__ R = a.records + (size_t)info.x * kTileCells + tid;
    unsigned pre[kFoldBatch];
#pragma unroll
    for (int u = 0; u < kFoldBatch; ++u) pre[u] = (unsigned)u < info.y ? __ldcs(R + (size_t)u * kTileCells) : 0u;
    for (unsigned j = 0; j < info.y; ++j) {
        const unsigned u0 = j % kFoldBatch;
        if (u0 == 0u) {
#pragma unroll
            for (int u = 0; u < kFoldBatch; ++u) sRec[u * kTileCells + tid] = pre[u];
#pragma unroll
            for (int u = 0; u < kFoldBatch; ++u) {
                const unsigned jj = j + kFoldBatch + u;
                pre[u] = jj < info.y ? __ldcs(R + (size_t)jj * kTileCells) : 0u;
            }
        }
        const unsigned rec = sRec[u0 * kTileCells + tid];
        if (__any_sync(0xffffffffu, f.qn > 64 - (int)kRawMax)) f.drain();
        const unsigned cnt = rec >> 26;
        if (cnt <= kRawMax) {                                   // raw (or empty): queue
            f.q |= (unsigned long long)(rec & ((1u << 26) - 1u)) << f.qn;
            f.qn += (int)cnt;
            continue;
        }
        f.drain();
        if (rec == kRecOverflow) {
            const int4 pr = a.pairs[2 * (size_t)(info.x + j)];
            const int2* __restrict__ E = a.rel + a.pairs[2 * (size_t)(info.x + j) + 1].x;
            const int rx = pr.x + (tid & (kTile - 1)), ry = pr.y + (tid >> kTileShift);
            for (int i = pr.z; i < pr.w; ++i) {
                const int2 e = __ldg(E + i);
                const int ty = rayTouch(rx, ry, e.x, e.y);
                if (ty) f.touch(ty == 2);
            }
        } else if (cnt == 31u) {
            const unsigned* __restrict__ S = a.side + (size_t)(rec & ((1u << 26) - 1u)) * 8u;
            const unsigned nWords = S[0] >> 16;
            const uint2* __restrict__ W = reinterpret_cast<const uint2*>(S + 2);
            uint2 nx = W[0];
            for (unsigned w = 0; w < nWords; ++w) {
                unsigned t = nx.x;
                const unsigned h = nx.y;
                if (w + 1 < nWords) nx = W[w + 1];
                while (t) {
                    const bool hit = (h >> (__ffs(t) - 1)) & 1u;
                    const unsigned same = hit ? (t & h) : (t & ~h);
                    const unsigned other = t & ~same;
                    const unsigned upto = other ? ((1u << (__ffs(other) - 1)) - 1u) : 0xFFFFFFFFu;
                    const unsigned run = same & upto;
                    f.run(hit, (unsigned)__popc(run));
                    t &= ~run;
                }
            }
        } else {                                                // three alternating runs
            const bool first = (rec >> 30) & 1u;
            f.run(first, (rec >> 20) & 1023u);
            f.run(!first, (rec >> 10) & 1023u);
            f.run(first, rec & 1023u);
        }
    }
    f.drain();
    if (a.diag) { atomicMax(a.diag + 2, (unsigned long long)f.computed); atomicAdd(a.diag + 3, (unsigned long long)f.computed); }
    if (inside && f.v != v0) *cell = f.v;
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

}  // extern "C"

namespace {

// Completion of the OLDEST submitted call: its update count, the fallback statistics, its staging set.
int integ_wait(lgs_ctx* c, long long* nUpdatesOut) {
    if (nUpdatesOut) *nUpdatesOut = 0;
    if (!c->integ || c->integ->nPending == 0) return LGS_OK;
    lgs_integ_ws& w = *c->integ;
    lgs_integ_ws::Stage& st = w.stage[(w.calls - (unsigned long long)w.nPending) & 1];
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaEventSynchronize(st.evDone));
    st.pending = false;
    --w.nPending;
    if (nUpdatesOut) *nUpdatesOut = (long long)st.hCounters.p[0];
    w.fallbackCells += (long long)st.hCounters.p[1];
    return LGS_OK;
}

// Everything of a call up to the last enqueue; the host waits only for the pre-pass (scan bounds).
int integ_submit(lgs_ctx* c, lgs_grid* grid, const lgs_hit_batch* scans, double pHit, double pMiss) {
    const int n = scans->n_scans;
    if (n < 0 || (n > 0 && (!scans->sensor_xy || !scans->hit_begin)))
        return lgs_fail(c, LGS_ERR_INVALID, "integrate: bad scan batch");
    if (n == 0) return LGS_OK;                       // nothing in flight: the matching wait reports 0 updates
    if (grid->off_x || grid->off_y)
        return lgs_fail(c, LGS_ERR_INVALID, "integrate: windowed grids (lgs_grid_set_window) are not supported");
    const long long total = n > 0 ? scans->hit_begin[n] : 0;
    if (total > 0 && !scans->hit_xy) return lgs_fail(c, LGS_ERR_INVALID, "integrate: hit_xy is NULL");
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, lgs_grid_acquire(c, grid));
    if (!c->integ) c->integ = new lgs_integ_ws();
    lgs_integ_ws& ws = *c->integ;
    if (ws.nPending >= 2 || ws.stage[ws.calls & 1].pending)
        return lgs_fail(c, LGS_ERR_INVALID, "integrate: two calls are in flight, wait for the older one first");
    lgs_integ_ws::Stage& w0 = ws.stage[ws.calls & 1];
    if (!ws.copyStream) LGS_CUDA(c, cudaStreamCreateWithFlags(&ws.copyStream, cudaStreamNonBlocking));
    if (!w0.evDone) {
        LGS_CUDA(c, cudaEventCreateWithFlags(&w0.evDone, cudaEventDisableTiming));
        LGS_CUDA(c, cudaEventCreateWithFlags(&w0.evCounters, cudaEventDisableTiming));
    }
    // `w` = the shared workspace with this call's staging set in front of it
    struct View {
        DevBuf<double>& sensor; DevBuf<double>& hit; DevBuf<int>& begin; DevBuf<char>& meta; DevBuf<int2>& rel;
        DevBuf<unsigned long long>& counters; PinBuf<char>& hMeta; PinBuf<unsigned long long>& hCounters;
        DevBuf<unsigned>& kmin; DevBuf<unsigned>& kmax;
        DevBuf<uint2>* tileInfo; DevBuf<int4>* pairs; DevBuf<unsigned>* records; DevBuf<unsigned>* side;
        cudaStream_t& foldStream; cudaEvent_t* evTouch; cudaEvent_t* evFold;
        size_t& cleanTiles; bool& dirty;
    } w{w0.sensor, w0.hit, w0.begin, w0.meta, w0.rel, w0.counters, w0.hMeta, w0.hCounters, ws.kmin, ws.kmax,
        ws.tileInfo, ws.pairs, ws.records, ws.side, ws.foldStream, ws.evTouch, ws.evFold, ws.cleanTiles, ws.dirty};
    cudaStream_t cs0 = ws.copyStream;
    // LGS_INTEG_HOSTTIMING=1 (diagnostic): host wall time of the call's phases to stderr when a call
    // takes longer than a millisecond.
    const bool hostTiming = c->opt.integHostTiming != 0;
    const auto tStart = std::chrono::steady_clock::now();
    auto msSince = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    double tStage = 0.0, tPre = 0.0, tLoop = 0.0;

    // Stage the whole batch once, on the copy stream (under the chunks of an earlier call that is still
    // running); the passes then run over chunks of <= 64 scans (one mask bit each).
    LGS_CUDA(c, w.sensor.reserve((size_t)n * 2));
    LGS_CUDA(c, w.hit.reserve(std::max<size_t>((size_t)total, 1) * 2));
    LGS_CUDA(c, w.begin.reserve((size_t)n + 1));
    LGS_CUDA(c, w.meta.reserve((size_t)n * sizeof(ScanMeta)));
    LGS_CUDA(c, w.rel.reserve(std::max<size_t>((size_t)total, 1)));
    LGS_CUDA(c, w.counters.reserve(8));
    LGS_CUDA(c, w.hMeta.reserve((size_t)n * sizeof(ScanMeta)));
    LGS_CUDA(c, w.hCounters.reserve(8));
    LGS_CUDA(c, cudaMemcpyAsync(w.sensor.p, scans->sensor_xy, (size_t)n * 2 * sizeof(double), cudaMemcpyHostToDevice, cs0));
    if (total)
        LGS_CUDA(c, cudaMemcpyAsync(w.hit.p, scans->hit_xy, (size_t)total * 2 * sizeof(double), cudaMemcpyHostToDevice, cs0));
    LGS_CUDA(c, cudaMemcpyAsync(w.begin.p, scans->hit_begin, (size_t)(n + 1) * sizeof(int), cudaMemcpyHostToDevice, cs0));
    LGS_CUDA(c, cudaMemsetAsync(w.counters.p, 0, 8 * sizeof(unsigned long long), cs0));
    tStage = msSince(tStart);

    int maxBeams = 0;
    for (int s = 0; s < n; ++s) maxBeams = std::max(maxBeams, scans->hit_begin[s + 1] - scans->hit_begin[s]);
    ScanMeta* dMeta = reinterpret_cast<ScanMeta*>(w.meta.p);
    ScanMeta* hMeta = reinterpret_cast<ScanMeta*>(w.hMeta.p);
    GridRef g{grid->origin(), grid->nx, grid->ny, grid->pitch, grid->min_x, grid->min_y, grid->res};
    // Pre-pass over the whole batch: sensor cells, relative end cells, ray lengths, bounds check.
    integ_sensor_kernel<<<(n + 127) / 128, 128, 0, cs0>>>(w.sensor.p, w.begin.p, n, g, dMeta);
    LGS_LAUNCH_CHECK(c);
    if (maxBeams > 0) {
        dim3 gb((maxBeams + 127) / 128, n);
        integ_beam_kernel<<<gb, 128, 0, cs0>>>(w.hit.p, g, dMeta, w.rel.p);
        LGS_LAUNCH_CHECK(c);
    }
    LGS_CUDA(c, cudaMemcpyAsync(hMeta, dMeta, (size_t)n * sizeof(ScanMeta), cudaMemcpyDeviceToHost, cs0));
    LGS_CUDA(c, cudaStreamSynchronize(cs0));
    tPre = msSince(tStart);
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
    if (c->opt.integSideWords > 0) sideCap = (size_t)std::max<long long>(8, c->opt.integSideWords);
    // tiles of scan s's reach: the square of half side maxLen around the sensor cell
    auto tileSpan = [](int lo, int hi) { return (hi >> kTileShift) - (lo >> kTileShift) + 1; };
    // LGS_INTEG_TIMING=1 (diagnostic): CUDA-event time of every pass, summed over the call, to stderr.
    const bool timing = c->opt.integTiming != 0;
    const bool diag = timing && c->opt.integDiag != 0;   // + per-cell update counts (slows the fold)
    std::vector<cudaEvent_t> evs;
    auto stamp = [&]() { if (timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, c->stream); evs.push_back(e); } };
    // The fold pass runs on a second stream so that it overlaps the next chunk's touch passes
    // (LGS_INTEG_TIMING serialises everything on the context stream to time the passes).
    const bool overlap = !timing;
    if (!w.foldStream) {
        LGS_CUDA(c, cudaStreamCreateWithFlags(&w.foldStream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            LGS_CUDA(c, cudaEventCreateWithFlags(&w.evTouch[b], cudaEventDisableTiming));
            LGS_CUDA(c, cudaEventCreateWithFlags(&w.evFold[b], cudaEventDisableTiming));
        }
    }
    int lastBuf = -1;
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
        const int buf = (int)(ws.chunks & 1);
        ++ws.chunks;
        // this buffer's previous fold (chunk k - 2, possibly of the previous call) must be done before it is
        // overwritten: the context stream waits for it; only a buffer that has to grow makes the host wait
        // (reallocation frees)
        if (ws.usedBuf[buf]) {
            const bool grows = nTiles > w.tileInfo[buf].cap || 2 * pairBound > w.pairs[buf].cap ||
                               pairBound * kTileCells > w.records[buf].cap || sideCap > w.side[buf].cap;
            if (grows) LGS_CUDA(c, cudaEventSynchronize(w.evFold[buf]));
            else LGS_CUDA(c, cudaStreamWaitEvent(c->stream, w.evFold[buf], 0));
        }
        LGS_CUDA(c, w.tileInfo[buf].reserve(nTiles));
        LGS_CUDA(c, w.pairs[buf].reserve(2 * pairBound));
        LGS_CUDA(c, w.records[buf].reserve(pairBound * kTileCells));
        LGS_CUDA(c, w.side[buf].reserve(sideCap));
        if (nTiles > w.cleanTiles) {
            const size_t a0 = w.cleanTiles, cnt = w.kmin.cap / kChunk - a0;   // initialise up to the capacity
            LGS_CUDA(c, cudaMemsetAsync(w.kmin.p + a0 * kChunk, 0xFF, cnt * kChunk * sizeof(unsigned), c->stream));
            LGS_CUDA(c, cudaMemsetAsync(w.kmax.p + a0 * kChunk, 0, cnt * kChunk * sizeof(unsigned), c->stream));
            w.cleanTiles = w.kmin.cap / kChunk;
        }
        unsigned* nPairs = reinterpret_cast<unsigned*>(w.counters.p + 4 + buf);   // [0] pairs, [1] side-buffer cursor
        LGS_CUDA(c, cudaMemsetAsync(nPairs, 0, 2 * sizeof(unsigned), c->stream));

        w.dirty = true;
        stamp();
        const int beamsPad = (subBeams + 31) & ~31, pieces = subLen / kTile + 1;
        dim3 gm((unsigned)(((size_t)beamsPad * pieces + 127) / 128), ns);
        integ_mark_kernel<<<gm, 128, 0, c->stream>>>(dMeta + cs, w.rel.p, x0, y0, tw, beamsPad, w.kmin.p, w.kmax.p);
        LGS_LAUNCH_CHECK(c);
        stamp();
        integ_pairs_kernel<<<(unsigned)((nTiles * 32 + 127) / 128), 128, 0, c->stream>>>((int)nTiles, dMeta + cs, x0, y0, tw, w.kmin.p, w.kmax.p,
                                                                                       w.tileInfo[buf].p, w.pairs[buf].p, nPairs);
        LGS_LAUNCH_CHECK(c);
        stamp();
        w.dirty = false;
        const unsigned touchBlocks = (unsigned)std::min<size_t>(pairBound, (size_t)c->sm_count * 4);
        integ_touch_kernel<<<touchBlocks, kTileCells, 0, c->stream>>>(w.rel.p, w.pairs[buf].p, nPairs, w.records[buf].p,
                                                                      w.side[buf].p, nPairs + 1, (unsigned)sideCap,
                                                                      w.counters.p);
        LGS_LAUNCH_CHECK(c);
        stamp();
        // fold on its own stream: it only has to follow this chunk's touch pass and the previous fold
        LGS_CUDA(c, cudaEventRecord(w.evTouch[buf], c->stream));
        cudaStream_t fs = overlap ? w.foldStream : c->stream;
        if (overlap) LGS_CUDA(c, cudaStreamWaitEvent(fs, w.evTouch[buf], 0));
        FoldArgs a{w.rel.p, w.tileInfo[buf].p, w.pairs[buf].p, w.records[buf].p, w.side[buf].p,
                   diag ? w.counters.p : nullptr, pHit, pMiss, oddsHit, oddsMiss};
        integ_fold_kernel<<<(unsigned)nTiles, kTileCells, 0, fs>>>(a, g, x0, y0, tw);
        LGS_LAUNCH_CHECK(c);
        LGS_CUDA(c, cudaEventRecord(w.evFold[buf], fs));
        ws.usedBuf[buf] = true;
        lastBuf = buf;
        stamp();
    }
    tLoop = msSince(tStart);
    // The call is complete when its counters are back AND its last fold is done.  The context stream does
    // not wait for that fold: the next call's first chunk starts under it (the record buffers and the
    // fold stream's own order keep the cells right); lgs_grid_integrate_wait is what orders everything
    // else behind the call.
    LGS_CUDA(c, cudaMemcpyAsync(w.hCounters.p, w.counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    if (overlap && lastBuf >= 0) {
        LGS_CUDA(c, cudaEventRecord(w0.evCounters, c->stream));
        LGS_CUDA(c, cudaStreamWaitEvent(w.foldStream, w0.evCounters, 0));
        LGS_CUDA(c, cudaEventRecord(w0.evDone, w.foldStream));
    } else {
        LGS_CUDA(c, cudaEventRecord(w0.evDone, c->stream));
    }
    w0.pending = true;
    ++ws.calls;
    ++ws.nPending;
    if (hostTiming && msSince(tStart) > c->opt.integHostTimingMinMs)
        fprintf(stderr, "[lgs integrate host] %d scans into %dx%d: staged %.3f ms, pre-pass synced %.3f ms, chunks queued "
                "%.3f ms\n", n, grid->nx, grid->ny, tStage, tPre, tLoop);
    if (timing) {
        LGS_CUDA(c, cudaStreamSynchronize(c->stream));
        float t[4] = {0, 0, 0, 0};
        for (size_t k = 0; k + 5 <= evs.size(); k += 5)      // 5 stamps per chunk
            for (int j = 0; j < 4; ++j) { float ms = 0; cudaEventElapsedTime(&ms, evs[k + j], evs[k + j + 1]); t[j] += ms; }
        fprintf(stderr, "[lgs integrate] %d scans: mark %.3f ms, pairs %.3f ms, touch %.3f ms, fold %.3f ms; computed updates: "
                "max per cell and call %llu, total %llu of %llu touches\n", n, t[0], t[1], t[2], t[3],
                w.hCounters.p[2], w.hCounters.p[3], w.hCounters.p[0]);
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
    }
    return LGS_OK;
}

}  // namespace

extern "C" {

int lgs_grid_integrate_scans(lgs_ctx* c, lgs_grid* grid, const lgs_hit_batch* scans, double pHit,
                             double pMiss, long long* nUpdatesOut) {
    if (!c || !grid || !scans) return LGS_ERR_INVALID;
    if (nUpdatesOut) *nUpdatesOut = 0;
    // calls submitted asynchronously before this one finish first (their counts are dropped)
    while (c->integ && c->integ->nPending > 0) { const int rc = integ_wait(c, nullptr); if (rc != LGS_OK) return rc; }
    if (scans->n_scans == 0) return LGS_OK;
    const int rc = integ_submit(c, grid, scans, pHit, pMiss);
    if (rc != LGS_OK) return rc;
    return integ_wait(c, nUpdatesOut);
}

int lgs_grid_integrate_submit(lgs_ctx* c, lgs_grid* grid, const lgs_hit_batch* scans, double pHit, double pMiss) {
    if (!c || !grid || !scans) return LGS_ERR_INVALID;
    return integ_submit(c, grid, scans, pHit, pMiss);
}

int lgs_grid_integrate_wait(lgs_ctx* c, long long* nUpdatesOut) {
    if (!c) return LGS_ERR_INVALID;
    return integ_wait(c, nUpdatesOut);
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
    // stream-ordered: the new buffer comes from the pool and the old one returns to it after the
    // copy, without synchronising the device (a map that grows every few frames pays microseconds)
    cudaError_t e = lgs_alloc_async(c, &nd, std::max<size_t>(bytes, 8));
    if (e != cudaSuccess) return lgs_fail(c, LGS_ERR_NOMEM, "grid_resize: cudaMallocAsync(%zu) -> %s", bytes, cudaGetErrorString(e));
    LGS_CUDA(c, cudaMemsetAsync(nd, 0, bytes, c->stream));
    if (nx > 0 && ny > 0) {
        dim3 gridDim((nx + 255) / 256, ny);
        grid_shift_copy_kernel<<<gridDim, 256, 0, c->stream>>>(g->origin(), g->nx, g->ny, g->pitch,
                                                               nd + (size_t)g->apron * pitch + g->apron,
                                                               nx, ny, (int)pitch, shiftX, shiftY);
        LGS_LAUNCH_CHECK(c);
    }
    LGS_CUDA(c, cudaFreeAsync(g->d, c->stream));
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
