// lgs_bb_run.cu -- the device-only branch-and-bound run: ONE persistent cooperative kernel.
//
// Replaces ScanMatcherBranchBound::OptimizePose (mapping/scan_matcher_branch_bound.cpp:47-163) with
// ScorePixelAccurate::Score (mapping/score_function_pixel_accurate.cpp:19-76) for a whole batch of
// (scan, submap) queries; see lgs_bb.cu for why a breadth-first expansion of the static-threshold
// superset plus a verified (score desc, LIFO rank asc) winner reproduces the CPU's depth-first search.
//
// One launch per run.  The grid (a few CTAs per SM, all co-resident) walks the phases
//     hit points -> root level -> levels H-1 .. 0 -> winner -> verify / replay / records
// separated by grid-wide barriers; node counts never leave the device, and the host is not in the
// loop at all (the level-synchronous path of lgs_bb.cu needs one round trip per level plus one for
// the near-edge points).
//
// Cell of a beam without floating point in the inner loop: the hit point of (scan, beam, theta) at
// node offset (0, 0) is stored once as 12.20 FIXED-POINT CELLS relative to the scan's own origin cell,
// H = rn((h / res - origin) * 2^20)  (8 bytes per point, shared by all queries of the scan; usable
// beams reach < 2000 cells), a query stores  M = rn((min / res - origin) * 2^20) + (window offset << 20)
// split into its 20 fraction bits Mlo and whole cells Mhi, and a node offset (x, y) is a whole number
// of cells (stepX = stepY = res), so with 32-bit integer arithmetic only
//     cell = ((H - Mlo) >> 20) - Mhi + (x, y)          frac = (H - Mlo) & (2^20 - 1).
// That differs from the CPU's  floor(((sx + x * step) + r * cos - minX) / res)  only through rounding
// (< 1e-10 cells in double, 1e-6 cells of quantisation), so it is exact unless frac lies within the
// guard band of a cell edge.  Those points (about one per 1e5) are decided on the spot by evaluating the
// CPU's own expression, in the CPU's operation order, for the two ends of an interval that contains
// glibc's cos / sin (device value +- `resolveUlps` ulps; CUDA's double sincos is within 2 ulps of the
// true value, glibc's within 1): every operation is monotone, so if both ends floor to the same cell
// that IS the CPU's cell.  If they do not (probability ~1e-7 per near-edge point) the run counts an
// unresolved point and lgs_bb_batch_results repeats it on the exact path, where the host supplies
// glibc's values.
//
// What binds the scoring phases is the L1 request path (ncu: data-pipe wavefronts, not DRAM or L2
// bandwidth), so the hit points are as narrow as the guard band allows (8 instead of 16 bytes) and
// exist in two layouts, [beam][theta] and [theta][beam], so that every warp mapping below reads
// contiguous runs.
//
// Scoring keeps the CPU's summation order bit for bit: a node's cells are added in beam order in
// double.  How a warp is spent on that depends on the level's size (chosen on the device from the
// level's node count, per-pass cost model in RunArgs::costUs):
//   G = 1   one lane per node, 16 beams in flight per lane (large levels: throughput);
//   G = 4/8 4 / 8 lanes fetch the beams of one node, values are parked in shared memory in beam order
//           and one lane per node adds them in that order (mid-size levels: memory parallelism);
//   G = 32  the whole warp fetches one node (small levels: latency).
// Lanes that share a warp are neighbours in theta, so their hit points are contiguous and their map
// cells share sectors.
#include <cooperative_groups.h>

#include <cfloat>
#include <chrono>
#include <cmath>

#include "lgs_bb.cuh"

namespace cg = cooperative_groups;
using namespace lgsbb;

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
// 16 beams in flight per lane, 2 CTAs per SM (128 registers).
template <int KU> struct StageRow { static constexpr int value = 32 * KU + 8; };   // doubles of staging per warp: (32 / G) rows of KU * G + 1

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct NodeRef {
    int q, t, x, y;
    bool active;
};

// The CPU's cell of beam i of (query, theta, node offset), as an offset into the level's padded grid --
// or, if the interval around the device's cos / sin straddles a cell edge, an unresolved point
// (counted; the host repeats the run exactly).  Called BEFORE the gathers of a round are issued, when
// only the round's offsets are live (a call with the cell values and the prefetched hit points in
// registers spills them).
__device__ __noinline__ int bb_exact_offset(const RunArgs& a, int q, int t, int i, int x, int y, int h) {
    const BbQuery& d = a.qs[q];
    const BbScan& u = a.us[d.scan];
    // nodePose (scan_matcher_branch_bound.cpp:96-99), HitPoint (sensor_data.hpp:162-173),
    // WorldCoordinateToGridCellIndex (grid_map.hpp:779-790), in the CPU's operation order
    const double theta = __dadd_rn(u.st, __dmul_rn((double)(t - u.winT), u.stepT));
    const double ang = __dadd_rn(theta, a.angles[u.beamBegin + i]);
    double sn, cs;
    sincos(ang, &sn, &cs);
    const double r = a.ranges[u.beamBegin + i];
    const double px = __dadd_rn(u.sx, __dmul_rn((double)x, d.res));
    const double py = __dadd_rn(u.sy, __dmul_rn((double)y, d.res));
    const double k = (double)a.resolveUlps * 2.220446049250313e-16;
    const double wc = __dadd_rn(__dmul_rn(fabs(cs), k), 4.9406564584124654e-324);
    const double ws = __dadd_rn(__dmul_rn(fabs(sn), k), 4.9406564584124654e-324);
    auto cell = [&](double p, double trig, double mn) {
        return __double2int_rd(__ddiv_rn(__dsub_rn(__dadd_rn(p, __dmul_rn(r, trig)), mn), d.res));
    };
    const int ix0 = cell(px, __dsub_rn(cs, wc), d.minX), ix1 = cell(px, __dadd_rn(cs, wc), d.minX);
    const int iy0 = cell(py, __dsub_rn(sn, ws), d.minY), iy1 = cell(py, __dadd_rn(sn, ws), d.minY);
    if (ix0 != ix1 || iy0 != iy1) atomicAdd(a.ctr + kCtrUnresolved, 1);
    if (h == a.H) atomicAdd(&a.best[q].fixups, 1);
    const int ix = min(max(ix0 - d.offX, -1), d.nx);
    const int iy = min(max(iy0 - d.offY, -1), d.ny);
    return iy * d.pitch + ix;           // relative to the level's cell (0, 0)
}

// Lane roles of the warp mappings.  G = 1 / 4 read the [beam][theta] hit array: lanes that are
// neighbours in theta sit next to each other (row = lane % NPW), so a request covers NPW contiguous
// hit points per beam.  G = 8 / 32 read the [theta][beam] copy: the G lanes of a node are neighbours
// (sub = lane % G) and fetch G consecutive beams, one 64- / 256-byte run per node.
template <int G> struct LaneMap {
    static constexpr int NPW = 32 / G;
    static constexpr bool kBeamMajor = G >= 8;
    __device__ static __forceinline__ int row(int lane) { return kBeamMajor ? lane / G : lane % NPW; }
    __device__ static __forceinline__ int sub(int lane) { return kBeamMajor ? lane % G : lane / NPW; }
};

// ScorePixelAccurate::Score of the (up to 32 / G) nodes of this warp on pyramid level h.  Lane
// `lane` works for node LaneMap<G>::row(lane) and fetches the beams congruent to LaneMap<G>::sub(lane)
// modulo G; the sum is returned in the lanes with sub == 0.
template <int G, int kU>
__device__ __forceinline__ double score_group(const RunArgs& a, const NodeRef& n, int h, int lane, double* svw, int& skipped) {
    const int sub = LaneMap<G>::sub(lane), row = LaneMap<G>::row(lane);
    (void)sub; (void)row;
    const BbQuery* dq = a.qs + (n.active ? n.q : 0);
    const BbScan* su = a.us + dq->scan;
    const int t = n.active ? n.t : 0;
    // hit point of beam b: hb[b * stride]
    const int2* hb = LaneMap<G>::kBeamMajor ? a.hitsT + su->hitTBegin + (long long)t * su->beamPad
                                            : a.hits + su->hitBegin + t;
    const unsigned stride = LaneMap<G>::kBeamMajor ? 1u : (unsigned)su->nTpad;
    // cell + 1 = ((H - Mlo) >> 20) - (Mhi - node offset - 1), frac = (H - Mlo) & (2^20 - 1).  The + 1
    // turns "clamp into [-1, n]" (out of the map -> zero apron) into ONE unsigned min: a cell left of /
    // below the map wraps to a huge unsigned value and lands in the right / upper apron, which is zero too.
    const int mlx = dq->MloX, mly = dq->MloY;
    const int cx = dq->MhiX - n.x - 1, cy = dq->MhiY - n.y - 1;
    const int pitch = dq->pitch;
    const unsigned gx1 = (unsigned)dq->nx + 1u, gy1 = (unsigned)dq->ny + 1u;
    const double* __restrict__ lvl = dq->level[h] - pitch - 1;        // cell (-1, -1) of the padded level
    const int nb = n.active ? dq->nUse : 0;
    // near-edge test on the 20 fraction bits moved to the top of the word: ((d << 12) + (E << 12)) < (2E << 12)
    // in wrapping unsigned arithmetic -- one shift-add (LEA) per axis instead of add + mask
    const unsigned E = a.edgeUnits << 12, E2 = (2u * a.edgeUnits) << 12;
    double acc = 0.0;
    // cell values are probabilities (<= 1): remaining beams bound the rest of the sum ("bb_early_reject")
    const double lim = a.earlyReject ? dq->thrAbs - 1e-6 : -1.0e300;

    // offset of the beam's cell; `near` keeps the smallest distance-to-edge measure seen so far
    // (one compare per round instead of one per beam and axis)
    auto offset = [&](const int2 hp, unsigned& near) -> unsigned {
        const int dx = hp.x - mlx, dy = hp.y - mly;
        near = min(near, min(((unsigned)dx << 12) + E, ((unsigned)dy << 12) + E));
        const unsigned ix = min((unsigned)((dx >> 20) - cx), gx1);
        const unsigned iy = min((unsigned)((dy >> 20) - cy), gy1);
        return iy * (unsigned)pitch + ix;
    };
    auto is_near = [&](const int2 hp) -> bool {
        const int dx = hp.x - mlx, dy = hp.y - mly;
        return ((((unsigned)dx << 12) + E) < E2) | ((((unsigned)dy << 12) + E) < E2);
    };
    // near-edge beam: the CPU's own cell, + 1 per axis like `offset`
    auto exact = [&](int beam) -> unsigned {
        return (unsigned)(bb_exact_offset(a, n.q, n.t, beam, n.x, n.y, h) + pitch + 1);
    };

    if constexpr (G == 1) {
        if (!n.active) return 0.0;
        const int nFull = nb / kU;
        const int2* hr = hb;                           // hit points of the current round: hr[u * stride]
        const size_t roundStep = (size_t)kU * stride;
        int2 hp[kU];
#pragma unroll 1
        for (int r = 0; r < nFull; ++r) {
            unsigned off[kU];
            double v[kU];
            unsigned near = 0xffffffffu;
#pragma unroll
            for (int u = 0; u < kU; ++u) hp[u] = hr[(unsigned)u * stride];
#pragma unroll
            for (int u = 0; u < kU; ++u) off[u] = offset(hp[u], near);
            if (near < E2) {                 // rare: one call site, the patch is a select chain (off[] stays in registers)
#pragma unroll 1
                for (int u = 0; u < kU; ++u) {
                    int2 hq = hp[0];
#pragma unroll
                    for (int w = 1; w < kU; ++w) hq = (w == u) ? hp[w] : hq;
                    if (!is_near(hq)) continue;
                    const unsigned o = exact(r * kU + u);
#pragma unroll
                    for (int w = 0; w < kU; ++w) off[w] = (w == u) ? o : off[w];
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) v[u] = __ldg(lvl + off[u]);
            hr += roundStep;
#pragma unroll
            for (int u = 0; u < kU; ++u) acc = __dadd_rn(acc, v[u]);      // beam order; unknown cells add 0.0
            // early rejection: even if every remaining beam hit a cell of value 1 the node would stay at
            // or below the threshold -- the CPU prunes it whatever the rest of its sum is, and a stored
            // partial sum (<= threshold) makes every later test on this node come out the same
            if (acc + (double)(nb - (r + 1) * kU) <= lim) { skipped += (nb - (r + 1) * kU) / kU; return acc; }
        }
        for (int i = nFull * kU; i < nb; ++i) {
            unsigned near = 0xffffffffu;
            const int2 hq = hb[(unsigned)i * stride];
            unsigned off = offset(hq, near);
            if (near < E2) off = exact(i);
            acc = __dadd_rn(acc, __ldg(lvl + off));
        }
        return acc;
    } else {
        // Stage s = beams [s * S, (s + 1) * S): every lane fetches kV of them, the values are parked in
        // shared memory in beam order and the node's first lane adds them in that order.  (Measured and
        // dropped: a software pipeline that keeps the next stage's cells in flight under the adding lanes.
        // The scoring phases are bound by L1 request wavefronts, not by exposed latency; halving the
        // stage to afford the second set of registers doubled the per-stage overhead instead.)
        constexpr int kV = kU;                     // beams per lane per stage
        constexpr int S = kV * G;                  // beams per stage
        constexpr int RS = S + 1;                  // row stride (doubles): odd, so the adding lanes hit distinct banks
        const int nbMax = __reduce_max_sync(0xffffffffu, nb);
        const int nStages = (nbMax + S - 1) / S;
        const int nbAll = nb;
        int nbLive = nb;                           // 0 once the node is rejected early (see G == 1)
        const int leader = LaneMap<G>::kBeamMajor ? row * G : row;
        int2 hp[kV];
        double v[kV];
        unsigned okm = 0;
        auto load_hits = [&](int s) {
            okm = 0;
#pragma unroll
            for (int u = 0; u < kV; ++u) {
                const int b = s * S + u * G + sub;
                const bool ok = b < nbLive;
                okm |= (ok ? 1u : 0u) << u;
                hp[u] = ok ? hb[(unsigned)b * stride] : make_int2(mlx + 0x80000, mly + 0x80000);   // mid-cell: never near an edge
            }
        };
        auto gather = [&](int s) {      // consumes hp / okm of stage s, leaves the requests for v in flight
            unsigned off[kV];
            unsigned near = 0xffffffffu;
#pragma unroll
            for (int u = 0; u < kV; ++u) off[u] = offset(hp[u], near);
            if (near < E2) {                 // rare: one call site, the patch is a select chain
#pragma unroll 1
                for (int u = 0; u < kV; ++u) {
                    int2 hq = hp[0];
#pragma unroll
                    for (int w = 1; w < kV; ++w) hq = (w == u) ? hp[w] : hq;
                    if (!((okm >> u) & 1u) || !is_near(hq)) continue;
                    const unsigned o = exact(s * S + u * G + sub);
#pragma unroll
                    for (int w = 0; w < kV; ++w) off[w] = (w == u) ? o : off[w];
                }
            }
#pragma unroll
            for (int u = 0; u < kV; ++u) v[u] = ((okm >> u) & 1u) ? __ldg(lvl + off[u]) : 0.0;
        };
        auto park = [&](int s) {   // v -> shared memory, beam order
            double* rowp = svw + row * RS;
#pragma unroll
            for (int u = 0; u < kV; ++u) rowp[u * G + sub] = v[u];
        };
        auto chain = [&](int s) {  // the node's first lane adds stage s in beam order
            if (sub == 0 && nbLive > s * S) {
                const double* rowp = svw + row * RS;
                const int m = min(S, nbLive - s * S);
                int j = 0;
                for (; j + 8 <= m; j += 8) {
                    double t8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) t8[u] = rowp[j + u];
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, t8[u]);
                }
                for (; j < m; ++j) acc = __dadd_rn(acc, rowp[j]);
            }
        };
#pragma unroll 1
            for (int s = 0; s < nStages; ++s) {
                load_hits(s);
                gather(s);
                park(s);
                __syncwarp();
                chain(s);
                __syncwarp();
                const bool out = nbLive == 0 || (sub == 0 && acc + (double)max(nbAll - (s + 1) * S, 0) <= lim);
                if (__shfl_sync(0xffffffffu, out, leader)) { if (sub == 0 && nbLive) skipped += max(nbAll - (s + 1) * S, 0) / 16; nbLive = 0; }
                if (__all_sync(0xffffffffu, nbLive <= (s + 1) * S)) break;
            }
        return acc;
    }
}

// Threshold test, winner bookkeeping and child allocation of the nodes a warp has just scored.
// Survivors of a warp allocate their children together (one atomic per warp) and store them
// child-major, so the next level's lanes again walk neighbouring thetas with equal offsets.
// Visit (pop) order of the children: (x+w, y+w), (x, y+w), (x+w, y), (x, y)
// (scan_matcher_branch_bound.cpp:134-137).
template <int G>
__device__ __forceinline__ void expand(const RunArgs& a, int h, int k, const NodeRef& n, long long rank,
                                       bool isRoot, double acc, int lane) {
    const bool holder = n.active && LaneMap<G>::sub(lane) == 0;
    bool survive = false;
    if (holder) {
        a.scores[h][k] = acc;
        if (acc > a.qs[n.q].thrAbs) {                                   // :108 with scoreMax >= threshold
            if (h == 0)
                atomicMax(&a.best[n.q].scoreBits, (unsigned long long)__double_as_longlong(acc));
            else
                survive = true;
        }
    }
    if (a.countNodes && survive) atomicAdd(&a.best[n.q].pad, 4);        // nodes of the deeper levels, per query (placement weights)
    int childBase = -1, childStride = 0;
    const unsigned m = __ballot_sync(0xffffffffu, survive);
    if (m != 0u) {
        const int cnt = __popc(m), leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(a.ctr + kCtrChild + h, 4 * cnt);
        base = __shfl_sync(0xffffffffu, base, leader);
        if (survive) {
            const int r = __popc(m & ((1u << lane) - 1u));
            childBase = base + r;
            childStride = cnt;
            if (base + 4 * cnt <= a.cap[h - 1]) {
                const int w = 1 << (h - 1);
                const int dx[4] = {w, 0, w, 0}, dy[4] = {w, w, 0, 0};
                Node* next = a.nodes[h - 1];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    Node ch;
                    ch.x = (short)(n.x + dx[c]); ch.y = (short)(n.y + dy[c]); ch.t = n.t; ch.q = n.q;
                    ch.rank = rank * 4 + c;
                    ch.parent = k; ch.childBase = -1; ch.childStride = 0;
                    next[base + c * cnt + r] = ch;
                }
            } else if (lane == leader) {
                atomicOr(a.ctr + kCtrOverflow, 1 << (h - 1));           // the host repeats the run with larger pools
            }
        }
    }
    if (holder) {
        if (isRoot) {
            Node nd;
            nd.x = (short)n.x; nd.y = (short)n.y; nd.t = n.t; nd.q = n.q;
            nd.parent = -1; nd.childBase = childBase; nd.childStride = childStride; nd.rank = rank;
            a.nodes[h][k] = nd;
        } else if (h > 0) {
            a.nodes[h][k].childBase = childBase;
            a.nodes[h][k].childStride = childStride;
        }
    }
}

// Root level: every (query, root cell, theta).  A warp tile = 32 / G consecutive thetas of one (query,
// root cell); consecutive tiles are consecutive QUERIES of the same scan and theta range, so the warps
// of a CTA read the same hit-point lines (L1) while gathering from different submaps.
// One atomic per warp and phase: rounds of 16 beams left out by early rejection (diagnostic).
__device__ __forceinline__ void flush_skipped(const RunArgs& a, int skipped) {
    skipped = __reduce_add_sync(0xffffffffu, skipped);
    if ((threadIdx.x & 31) == 0 && skipped) atomicAdd(a.ctr + kCtrSkipped, skipped);
}

template <int G, int kU>
__device__ __forceinline__ void root_phase(const RunArgs& a, int lane, int gw, int tw, double* svw) {
    constexpr int NPW = 32 / G;
    const int H = a.H;
    int skipped = 0;
    for (int tile = gw; tile < a.rootTiles; tile += tw) {
        int lo = 0, hi = a.nu;                          // scan of this tile: tileBegin[lo] <= tile < tileBegin[lo + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.tileBegin[mid] <= tile) lo = mid; else hi = mid;
        }
        const BbScan& s = a.us[lo];
        int rel = tile - a.tileBegin[lo];
        const int mq = rel % s.qCount; rel /= s.qCount;
        const int nrxy = s.nrx * s.nry;
        const int r = rel % nrxy, j = rel / nrxy;
        const int q = a.qlist[s.qBegin + mq];
        const int t = j * NPW + LaneMap<G>::row(lane);
        const int kx = r / s.nry, ky = r % s.nry;
        NodeRef n;
        n.q = q; n.t = t; n.x = -s.winX + (kx << H); n.y = -s.winY + (ky << H);
        n.active = t < s.nT;
        // push order x asc, y asc, theta asc (scan_matcher_branch_bound.cpp:85-88); LIFO pops reverse it
        const int kLocal = (kx * s.nry + ky) * s.nT + t;
        const int k = a.qs[q].rootBegin + kLocal;
        const double acc = score_group<G, kU>(a, n, H, lane, svw, skipped);
        expand<G>(a, H, k, n, (long long)(nrxy * s.nT - 1 - kLocal), true, acc, lane);
    }
    flush_skipped(a, skipped);
}

template <int G, int kU>
__device__ __forceinline__ void level_phase(const RunArgs& a, int h, int nNodes, int lane, int gw, int tw, double* svw) {
    constexpr int NPW = 32 / G;
    int skipped = 0;
    const int nGroups = (nNodes + NPW - 1) / NPW;
    for (int g = gw; g < nGroups; g += tw) {
        const int k = g * NPW + LaneMap<G>::row(lane);
        NodeRef n;
        n.active = k < nNodes;
        long long rank = 0;
        n.q = 0; n.t = 0; n.x = 0; n.y = 0;
        if (n.active) {
            const Node nd = a.nodes[h][k];
            n.q = nd.q; n.t = nd.t; n.x = nd.x; n.y = nd.y;
            rank = nd.rank;
        }
        const double acc = score_group<G, kU>(a, n, h, lane, svw, skipped);
        expand<G>(a, h, k, n, rank, false, acc, lane);
    }
    flush_skipped(a, skipped);
}

// Which warp mapping serves a level of n nodes fastest on `tw` warps: passes x per-pass cost.
__device__ __forceinline__ int pick_mapping(const RunArgs& a, int n, int tw) {
    const int gs[4] = {1, 4, 8, 32};
    float bestCost = 3.0e38f;
    int g = 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long perPass = (long long)tw * (32 / gs[i]);
        const float cost = (float)((n + perPass - 1) / perPass) * a.costUs[i];
        if (cost < bestCost) { bestCost = cost; g = gs[i]; }
    }
    return g;
}

template <int kU, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) bb_run_kernel(const __grid_constant__ RunArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sv[kWarps][StageRow<kU>::value];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * kWarps + wib, tw = gridDim.x * kWarps;
    const long long gt = (long long)blockIdx.x * kThreads + threadIdx.x, nt = (long long)gridDim.x * kThreads;
    double* svw = sv[wib];
    const int H = a.H;

    int phase = 0;
    auto stamp = [&]() { if (gt == 0) a.phaseNs[phase] = global_ns(); ++phase; };
    stamp();
    // ---- phase 0: clear the next run's counters, reset the winners, project the scans -------------
    for (long long i = gt; i < kCounters; i += nt) a.ctrNext[i] = 0;
    for (long long q = gt; q < a.nq; q += nt) {
        BbBest b;
        b.scoreBits = 0ull; b.rank = 0x7fffffffffffffffLL; b.leaf = -1; b.needReplay = 0;
        b.rankLeaf = ~0ull; b.fixups = 0; b.pad = 0;
        a.best[q] = b;
    }
    // Hit points: a warp takes an 8 x 32 tile of (beam, theta) of one scan; a lane owns one theta.  The
    // [beam][theta] copy is written straight from the registers (a row of thetas per beam), the
    // [theta][beam] copy through the warp's staging rows (64-byte runs of eight beams per theta).
    for (int tile = gw; tile < a.projTiles; tile += tw) {
        int lo = 0, hi = a.nu;                          // scan of this tile: projBegin[lo] <= tile < projBegin[lo + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (a.us[mid].projBegin <= tile) lo = mid; else hi = mid;
        }
        const BbScan& s = a.us[lo];
        const int rel = tile - s.projBegin;
        const int nTT = (s.nT + 31) >> 5;
        const int i0 = (rel / nTT) << 3, t = ((rel % nTT) << 5) + lane;
        // nodePose.mTheta = sensorPose.mTheta + node.mTheta * stepTheta (scan_matcher_branch_bound.cpp:96-99)
        const double theta = __dadd_rn(s.st, __dmul_rn((double)(t - s.winT), s.stepT));
        const bool tOk = t < s.nT;
        int2* stage = reinterpret_cast<int2*>(svw);     // [8 beams][33]
        double ang8[8], r8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = min(i0 + u, s.nUse - 1);
            ang8[u] = a.angles[s.beamBegin + i];
            r8[u] = a.ranges[s.beamBegin + i];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u;
            int2 hp = make_int2(0, 0);
            if (i < s.nUse && tOk) {
                double sn, cs;
                sincos(__dadd_rn(theta, ang8[u]), &sn, &cs);
                const double hx = __dadd_rn(s.sx, __dmul_rn(r8[u], cs));          // sensor_data.hpp:171-172
                const double hy = __dadd_rn(s.sy, __dmul_rn(r8[u], sn));
                // 12.20 fixed-point cells relative to the scan origin
                hp = make_int2(__double2int_rn(__dmul_rn(__dsub_rn(__dmul_rn(hx, s.invRes), s.originX), 1048576.0)),
                               __double2int_rn(__dmul_rn(__dsub_rn(__dmul_rn(hy, s.invRes), s.originY), 1048576.0)));
                a.hits[s.hitBegin + (long long)i * s.nTpad + t] = hp;
            }
            stage[u * 33 + lane] = hp;
        }
        __syncwarp();
        const int ub = lane & 7;                        // beam within the group of eight
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int tl = (lane >> 3) + 4 * k, tt = (t - lane) + tl;
            const int i = i0 + ub;
            if (tt < s.nT && i < s.nUse) a.hitsT[s.hitTBegin + (long long)tt * s.beamPad + i] = stage[ub * 33 + tl];
        }
        __syncwarp();
    }
    grid.sync();
    stamp();

    // ---- root level ----------------------------------------------------------------------------------
    switch (a.rootG) {
        case 1: root_phase<1, kU>(a, lane, gw, tw, svw); break;
        case 4: root_phase<4, kU>(a, lane, gw, tw, svw); break;
        case 8: root_phase<8, kU>(a, lane, gw, tw, svw); break;
        default: root_phase<32, kU>(a, lane, gw, tw, svw); break;
    }
    grid.sync();
    stamp();
    if (gt == 0) a.phaseG[H] = a.rootG;

    // ---- levels H-1 .. 0 -------------------------------------------------------------------------------
    for (int h = H - 1; h >= 0; --h) {
        int nNodes = __ldcg(a.ctr + kCtrChild + h + 1);
        if (nNodes > a.cap[h]) nNodes = 0;      // overflowed pool (flagged by its allocator): nothing valid to read
        const int g = pick_mapping(a, nNodes, tw);
        if (gt == 0) a.phaseG[h] = g;
        switch (g) {
            case 1: level_phase<1, kU>(a, h, nNodes, lane, gw, tw, svw); break;
            case 4: level_phase<4, kU>(a, h, nNodes, lane, gw, tw, svw); break;
            case 8: level_phase<8, kU>(a, h, nNodes, lane, gw, tw, svw); break;
            default: level_phase<32, kU>(a, h, nNodes, lane, gw, tw, svw); break;
        }
        grid.sync();
        stamp();
    }

    // ---- winner among the leaves: (score desc, rank asc) --------------------------------------------------
    {
        int nLeaves = H == 0 ? a.totalRoots : __ldcg(a.ctr + kCtrChild + 1);
        if (nLeaves > a.cap[0]) nLeaves = 0;
        for (long long k = gt; k < nLeaves; k += nt) {
            const Node nd = a.nodes[0][k];
            const unsigned long long bits = (unsigned long long)__double_as_longlong(a.scores[0][k]);
            const unsigned long long top = a.best[nd.q].scoreBits;
            if (top != 0ull && bits == top)
                atomicMin(&a.best[nd.q].rankLeaf, ((unsigned long long)nd.rank << 24) | (unsigned long long)k);
        }
    }
    grid.sync();
    stamp();

    // ---- per query: ancestor check, CPU-order replay if needed, result + 32-byte record ----------------------
    for (long long q = gt; q < a.nq; q += nt) {
        const BbQuery& d = a.qs[q];
        const BbBest b = a.best[q];
        BbResult r;
        r.exactReplay = 0; r.fixups = b.fixups; r.nodes = b.pad; r.pad = 0;
        r.found = 0; r.score = d.thrAbs; r.ix = 0; r.iy = 0; r.it = 0;
        bool replay = false;
        if (b.scoreBits != 0ull && b.rankLeaf != ~0ull) {
            // If every ancestor of the best leaf scores >= the leaf, the CPU's search provably returns it.
            const double s = __longlong_as_double((long long)b.scoreBits);
            const int leafIdx = (int)(b.rankLeaf & 0xffffffull);
            int idx = leafIdx;
            bool ok = true;
            for (int h = 0; h < H; ++h) {
                idx = a.nodes[h][idx].parent;
                if (a.scores[h + 1][idx] < s) ok = false;
            }
            const Node leaf = a.nodes[0][leafIdx];
            r.found = 1; r.score = s; r.ix = leaf.x; r.iy = leaf.y; r.it = leaf.t - d.winT;
            replay = !ok || a.forceReplay != 0;
        }
        if (replay) {
            // Sequential replay of the CPU's LIFO search over the stored superset scores (the win-max
            // maps are not upper bounds where a window index is negative).
            const int nRoots = d.nrx * d.nry * d.nT;
            double bestScore = d.thrAbs;
            int bestLeaf = -1;
            int stackIdx[4 * kMaxLevels];
            int stackH[4 * kMaxLevels];
            for (int rv = 0; rv < nRoots; ++rv) {              // roots in pop order: rank == rv
                int sp = 0;
                stackIdx[sp] = d.rootBegin + (nRoots - 1 - rv); stackH[sp] = H; ++sp;
                while (sp > 0) {
                    --sp;
                    const int idx = stackIdx[sp], h = stackH[sp];
                    const double s = a.scores[h][idx];
                    if (s <= bestScore) continue;                               // :108
                    if (h == 0) { bestScore = s; bestLeaf = idx; continue; }    // :114-120
                    const Node nd = a.nodes[h][idx];                            // s > best >= thr => expanded
                    for (int c = 3; c >= 0; --c) { stackIdx[sp] = nd.childBase + c * nd.childStride; stackH[sp] = h - 1; ++sp; }
                }
            }
            r.exactReplay = 1;
            if (bestLeaf >= 0) {
                const Node leaf = a.nodes[0][bestLeaf];
                r.found = 1; r.score = bestScore; r.ix = leaf.x; r.iy = leaf.y; r.it = leaf.t - d.winT;
            } else {
                r.found = 0; r.score = d.thrAbs; r.ix = 0; r.iy = 0; r.it = 0;
            }
        }
        a.res[q] = r;
        if (a.rec) {
            lgs_loop_record rc;
            rc.found = r.found; rc.ix = r.ix; rc.iy = r.iy; rc.it = r.it; rc.score = r.score;
            rc.id = a.recIds ? a.recIds[q] : q;
            a.rec[a.recFirst + q] = rc;
        }
    }
    if (gt == 0) {
        if (a.recStatus) {
            // Status record behind the batch's records: consumers of an exchanged buffer (other ranks) learn
            // from it whether this run has to be repeated on the exact path before its records count.
            const int overflow = __ldcg(a.ctr + kCtrOverflow), unresolved = __ldcg(a.ctr + kCtrUnresolved);
            lgs_loop_record st;
            st.found = (overflow == 0 && unresolved == 0) ? 1 : -1;
            st.ix = overflow; st.iy = unresolved; st.it = 0; st.score = 0.0; st.id = -1;
            a.rec[a.recFirst + a.nq] = st;
        }
        a.ctr[kCtrDone] = 1;
        a.phaseNs[phase] = global_ns();
    }
}

}  // namespace

// Enqueue one device-only run on the context stream (kernel + the small D2H of its counters and
// results); lgs_bb_batch_results waits and validates.
int lgs_bb_launch_device_run(lgs_bb_batch* b) {
    lgs_ctx* c = b->ctx;
    const int H = b->H, n = b->nq;
    // Pools: the largest count seen so far + 25 %, or a share of the root count before any run has
    // been measured.  A pool that is still too small is detected on the device and the run repeated.
    for (int h = H - 1; h >= 0; --h) {
        const long long want = std::max<long long>(b->hint[h] + b->hint[h] / 4 + 1024, b->totalRoots / 2 + 4096);
        if ((long long)b->dNodes[h].cap < want) LGS_CUDA(c, b->dNodes[h].reserve((size_t)want));
        LGS_CUDA(c, b->dScores[h].reserve(b->dNodes[h].cap));
    }
    LGS_CUDA(c, b->dNodes[H].reserve(b->totalRoots));
    LGS_CUDA(c, b->dScores[H].reserve(b->totalRoots));
    LGS_CUDA(c, b->dHitsFix.reserve(std::max<long long>(b->nHits, 1)));
    LGS_CUDA(c, b->dHitsFixT.reserve(std::max<long long>(b->nHitsT, 1)));
    LGS_CUDA(c, b->dRec.reserve(n + 1));
    LGS_CUDA(c, b->hRec.reserve(n + 1));
    if (b->dCtr.cap == 0) {
        LGS_CUDA(c, b->dCtr.reserve(2 * kCounters));
        LGS_CUDA(c, cudaMemsetAsync(b->dCtr.p, 0, 2 * kCounters * sizeof(int), c->stream));
        b->parity = 0;
    }
    const void* kernel = (const void*)bb_run_kernel<16, 2>;
    if (c->bbBlocks == 0) {
        int perSm = 0;
        LGS_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, bb_run_kernel<16, 2>, kThreads, 0));
        if (perSm < 1) return lgs_fail(c, LGS_ERR_CUDA, "bb: the persistent kernel does not fit an SM");
        if (c->opt.bbBlocksPerSm > 0) perSm = std::min(perSm, c->opt.bbBlocksPerSm);
        c->bbBlocks = perSm * c->sm_count;
    }
    const int tw = c->bbBlocks * kWarps;
    LGS_CUDA(c, b->dPhase.reserve(kPhases + kMaxLevels));
    LGS_CUDA(c, b->hPhase.reserve(kPhases + kMaxLevels));
    RunArgs a{};
    a.phaseNs = b->dPhase.p;
    a.phaseG = reinterpret_cast<int*>(b->dPhase.p + kPhases);
    char* blob = b->dBlob.p;
    a.qs = reinterpret_cast<const BbQuery*>(blob + b->offQs);
    a.us = reinterpret_cast<const BbScan*>(blob + b->offUs);
    a.qlist = reinterpret_cast<const int*>(blob + b->offQlist);
    a.angles = reinterpret_cast<const double*>(blob + b->offAngles);
    a.ranges = reinterpret_cast<const double*>(blob + b->offRanges);
    a.recIds = b->ids.empty() ? nullptr : reinterpret_cast<const long long*>(blob + b->offIds);
    a.hits = b->dHitsFix.p;
    a.hitsT = b->dHitsFixT.p;
    for (int h = 0; h <= H; ++h) {
        a.nodes[h] = b->dNodes[h].p;
        a.scores[h] = b->dScores[h].p;
        a.cap[h] = (int)std::min<size_t>(b->dNodes[h].cap, (size_t)1 << 24);
    }
    a.ctr = b->dCtr.p + b->parity * kCounters;
    a.ctrNext = b->dCtr.p + (1 - b->parity) * kCounters;
    a.best = b->dBest.p;
    a.res = b->dRes.p;
    a.rec = b->sink ? b->sink : b->dRec.p;
    a.recFirst = b->sink ? b->sinkFirst : 0;
    a.recStatus = 1;
    a.nq = n; a.nu = (int)b->us.size(); a.H = H;
    a.totalRoots = b->totalRoots;
    a.projTiles = b->projTiles;
    a.costUs[0] = (float)c->opt.bbCost[0]; a.costUs[1] = (float)c->opt.bbCost[1];
    a.costUs[2] = (float)c->opt.bbCost[2]; a.costUs[3] = (float)c->opt.bbCost[3];
    {   // root mapping: same cost model as the device uses for the deeper levels
        const int gs[4] = {1, 4, 8, 32};
        double bestCost = 1e300;
        int gi = 0;
        for (int i = 0; i < 4; ++i) {
            const long long tiles = b->tileBegin[i].back();
            const double cost = (double)((tiles + tw - 1) / tw) * c->opt.bbCost[i];
            if (cost < bestCost) { bestCost = cost; gi = i; }
        }
        a.rootG = gs[gi];
        a.rootTiles = b->tileBegin[gi].back();
        a.tileBegin = reinterpret_cast<const int*>(blob + b->offTiles) + (size_t)gi * (b->us.size() + 1);
        b->rootG = a.rootG;
    }
    {   // guard band in 2^-20 cells: the caller's band + quantisation of H and M (half a unit each) + double
        // rounding at the batch's largest coordinate (16 roundings of 2^-53 relative)
        const double units = std::ceil(c->opt.edgeEps * 1048576.0) + 2.0 +
                             std::ceil(b->maxAbsCells * 1.1102230246251565e-16 * 16.0 * 1048576.0);
        a.edgeUnits = (unsigned)std::min(units, 524287.0);
    }
    a.resolveUlps = std::max(c->opt.bbResolveUlps, 0);
    a.forceReplay = b->forceReplay ? 1 : 0;
    a.countNodes = (c->opt.bbHostTiming || c->opt.bbCountNodes) ? 1 : 0;
    a.earlyReject = c->opt.bbEarlyReject ? 1 : 0;
    void* params[] = {&a};
    LGS_CUDA(c, cudaLaunchCooperativeKernel(kernel, dim3(c->bbBlocks), dim3(kThreads), params, 0,
                                            c->stream));
    c->launches++;
    LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p, a.ctr, kCounters * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaMemcpyAsync(b->hRes.p, b->dRes.p, n * sizeof(BbResult), cudaMemcpyDeviceToHost, c->stream));
    if (c->opt.bbHostTiming)
        LGS_CUDA(c, cudaMemcpyAsync(b->hPhase.p, b->dPhase.p, (kPhases + kMaxLevels) * sizeof(unsigned long long),
                                    cudaMemcpyDeviceToHost, c->stream));
    b->parity ^= 1;
    b->pendingValidate = true;
    b->lastRunDevice = true;
    b->deviceRuns++;
    return LGS_OK;
}
