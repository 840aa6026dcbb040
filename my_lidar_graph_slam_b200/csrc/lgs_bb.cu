// lgs_bb.cu -- branch-and-bound scan matcher (loop detection) on sm_100a.
//
// Replaces ScanMatcherBranchBound::OptimizePose(grid, pyramids, scan, pose, thr)
// (mapping/scan_matcher_branch_bound.cpp:47-163) with ScorePixelAccurate::Score
// (mapping/score_function_pixel_accurate.cpp:19-76) as node score, for a whole batch of
// (scan, submap) queries at once.
//
// How the CPU's depth-first search is reproduced exactly by a breadth-first one:
//  * Node score S(n) is a pure function of (x, y, theta, height).  The CPU visits a node only
//    if every ancestor scored above the running best, which never drops below the static
//    threshold thr * NumOfScans().  bb_score_kernel therefore expands, level by level, the
//    SUPERSET of nodes whose ancestors all score above the static threshold, one thread per
//    node summing the gathered cells in beam order (bit-identical to the CPU sum).
//  * Let L* be the superset leaf with the highest score, ties broken by the CPU's LIFO visit
//    order (carried as an explicit rank: roots are popped x desc, y desc, theta desc; children
//    (x+w,y+w), (x,y+w), (x+w,y), (x,y); scan_matcher_branch_bound.cpp:85-88, :134-137).  If every
//    ancestor A of L* has S(A) >= S(L*), the CPU search provably returns L*: no earlier leaf
//    reaches S(L*), so no ancestor of L* is pruned, and no later leaf beats it.  bb_verify_kernel
//    checks exactly that.
//  * Otherwise (the win-max maps are not upper bounds where a window index is negative,
//    SURVEY.md H12) bb_replay_kernel replays the CPU's stack discipline sequentially over the
//    stored superset scores; every node the CPU can touch is in the superset.
//
// World-coordinate re-projection (H4): the CPU recomputes
//     ix = floor(((sx + nx*step) + r*cos(theta_t + a_i) - minX) / res)
// per node and beam.  In exact arithmetic that is I0 + nx with I0 the index at nx = 0; all
// rounding errors together stay below 1e-11 cells, so I0 + nx is exact unless the fractional
// cell coordinate lies within the guard band of an edge.  Such (theta, beam) pairs are flagged
// (bb_flagsearch_kernel / bb_index_kernel) and get per-offset index tables computed on the host
// with the CPU's own expression (and glibc sin/cos, H5).
//
// Two ways to get a beam's cell at node offset (0, 0), chosen per run:
//  * several queries share a scan (1 scan x 500 submaps): the root level, which visits every
//    (query, theta), converts the shared hit points on the fly (bb_score_root_kernel) and only the
//    surviving (query, theta) pairs get index rows (bb_index_slots_kernel); the near-edge flags come
//    from a sorted-fraction search over the scan's queries instead of an all-pairs pass;
//  * one scan per query, or more than kFlagInline flagged points: a full per-query index table
//    (bb_index_kernel), read by every level.
// After a batch object's first run the levels are launched speculatively over their pools'
// capacities with device-side node counts (no host round trip per level) and validated afterwards.
#include <cfloat>
#include <cmath>

#include <chrono>

#include "lgs_internal.cuh"

namespace {

constexpr int kFlagCapBB = 1 << 16;
constexpr int kMaxLevels = 21;
constexpr int kWarpPerNodeBelow = 8192;     // levels with fewer nodes score one warp per node
constexpr int kDeepUnrollBelow = 1 << 30;  // levels with fewer nodes keep 32 instead of 16 beams in flight (measured: always better)

struct BbScan {                     // one per DISTINCT (scan, sensor pose): hit points are map independent
    double sx, sy, st, stepT;
    int winT, nT, nTpad;            // theta slices, padded to a multiple of 4
    int nUse, beamBegin;            // usable beams (ScorePixelAccurate range filter)
    int pad;
    long long hitBegin;             // into hits: nUse * nTpad (double2), beam-major
    double invRes;                  // 1 / resolution of the maps this scan is matched against
    int sortBegin, sortCount;       // this scan's queries sorted by frac(min * invRes) (flag search)
};

struct BbQuery {
    double thrAbs;
    double minX, minY, res, invRes;
    int nx, ny, pitch;              // submap geometry
    int offX, offY;                 // window origin (cells) when the grid is a band of a larger map
    int winX, winY, winT, nT, nTpad;
    int nrx, nry;                   // roots per axis
    int nUse, scan;                 // usable beams, index of the distinct scan
    int rootBegin;                  // first root of this query in the level-H pool
    long long tabBegin;             // into the base-index table: nUse * nTpad int2, beam-major
    const double* level[kMaxLevels];// origin() of every pyramid level
};

struct Node {            // 32 bytes
    short x, y;          // window offsets of the node's lower-left corner
    int t;               // theta index 0..nT-1
    int q;               // query
    int parent;          // index in the pool one level up (-1 for roots)
    int childBase;       // child c (visit order) lives at childBase + c * childStride, -1 if pruned
    int childStride;
    long long rank;      // CPU visit order among nodes of the same height (lower = earlier)
};

struct BbBest {          // per query
    unsigned long long scoreBits;   // max leaf score above threshold (as ordered bits)
    long long rank;                 // visit rank of the winning leaf
    int leaf;                       // its index in the level-0 pool
    int needReplay;
};

struct BbResult {
    double score;
    int found, ix, iy, it;
    int exactReplay, pad;
};

struct BbFlag { int q, t, i; };
constexpr int kFlagInline = 8;      // the root-from-hit-points path handles up to this many near-edge points

// ---- stage A: world hit point of every (distinct scan, usable beam, theta) at node offset (0, 0) ----
// Threads are theta-fastest so neighbouring lanes differ by one angular step: their hit points are
// a fraction of a cell apart, which keeps every later table read and map gather coalesced.
__global__ void bb_hit_kernel(const BbScan* __restrict__ scans, const double* __restrict__ angles,
                              const double* __restrict__ ranges, double2* __restrict__ hits) {
    // grid = (theta blocks, beams, distinct scans): no index arithmetic, theta-fastest
    const BbScan& u = scans[blockIdx.z];
    const int i = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u.nUse || t >= u.nTpad) return;
    const long long idx = (long long)i * u.nTpad + t;
    if (t >= u.nT) { hits[u.hitBegin + idx] = make_double2(0.0, 0.0); return; }
    // nodePose.mTheta = sensorPose.mTheta + node.mTheta * stepTheta  (scan_matcher_branch_bound.cpp:96-99)
    const double theta = __dadd_rn(u.st, __dmul_rn((double)(t - u.winT), u.stepT));
    const double a = __dadd_rn(theta, angles[u.beamBegin + i]);
    double s, c;
    sincos(a, &s, &c);
    const double r = ranges[u.beamBegin + i];
    hits[u.hitBegin + idx] = make_double2(__dadd_rn(u.sx, __dmul_rn(r, c)),      // sensor_data.hpp:171-172
                                          __dadd_rn(u.sy, __dmul_rn(r, s)));
}

// ---- stage B: base cell index in every query's submap + near-edge flags -----------------------------
// One thread loads a hit point once and converts it for up to kIdxChunk queries that share the
// scan, so the (L2 resident) hit array is read once per chunk instead of once per query.
constexpr int kIdxChunk = 16;
struct IdxChunk { int scan, begin, count; };

__global__ void bb_index_kernel(const BbQuery* __restrict__ qs, const BbScan* __restrict__ scans,
                                const IdxChunk* __restrict__ chunks, const int* __restrict__ qlist,
                                const double2* __restrict__ hits, double eps, int2* __restrict__ tab,
                                BbFlag* __restrict__ flags, int* __restrict__ flagCount) {
    const IdxChunk ch = chunks[blockIdx.z];
    const BbScan& u = scans[ch.scan];
    const int i = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u.nUse || t >= u.nTpad) return;
    const long long idx = (long long)i * u.nTpad + t;
    const bool padT = t >= u.nT;
    const double2 h = padT ? make_double2(0.0, 0.0) : hits[u.hitBegin + idx];
    for (int k = 0; k < ch.count; ++k) {
        const int q = __ldg(qlist + ch.begin + k);
        const BbQuery& d = qs[q];
        if (padT) { tab[d.tabBegin + idx] = make_int2(-(1 << 28), -(1 << 28)); continue; }
        // grid_map.hpp:784-787 divides by the resolution; multiplying by its reciprocal differs from
        // that by < 1e-12 cells, far inside the guard band, so floor() agrees for every unflagged
        // point and flagged ones are re-derived on the host with the real division anyway.
        const double qx = __dmul_rn(__dsub_rn(h.x, d.minX), d.invRes);
        const double qy = __dmul_rn(__dsub_rn(h.y, d.minY), d.invRes);
        const double fx = floor(qx), fy = floor(qy);
        const double rx = qx - fx, ry = qy - fy;
        const bool edge = !(rx >= eps && rx <= 1.0 - eps && ry >= eps && ry <= 1.0 - eps);
        int2 v = make_int2(__double2int_rd(qx) - d.offX, __double2int_rd(qy) - d.offY);
        if (edge) {
            const int f = atomicAdd(flagCount, 1);
            if (f < kFlagCapBB) {
                flags[f] = BbFlag{q, t, i};
                v = make_int2(INT_MIN, f);      // sentinel: use the exact per-offset table f
            }
        }
        tab[d.tabBegin + idx] = v;
    }
}

// ---- stage B': near-edge flags without the all-pairs pass ------------------------------------------------
// A point is near an edge in query q iff frac((h - min_q) / res) is within eps of 0 or 1, i.e. iff
// frac(h / res) and frac(min_q / res) are within eps of each other (mod 1, up to ~1e-10 of rounding).
// The queries of a scan are sorted by frac(min_q / res) per axis on the host; a hit point binary
// searches the +-delta window (delta = eps + 1e-7, a superset) and runs the exact test of the
// reference expression only on those candidates.  Every near-edge (query, theta, beam) is registered
// exactly once: through the x list if x is near an edge, else through the y list.
__global__ void bb_flagsearch_kernel(const BbScan* __restrict__ scans, const BbQuery* __restrict__ qs,
                                     const double2* __restrict__ hits, const double* __restrict__ sortFx,
                                     const int* __restrict__ sortQx, const double* __restrict__ sortFy,
                                     const int* __restrict__ sortQy, double eps, double delta,
                                     BbFlag* __restrict__ flags, int* __restrict__ flagCount) {
    const BbScan& u = scans[blockIdx.z];
    const int i = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= u.nUse || t >= u.nT) return;
    const double2 h = hits[u.hitBegin + (long long)i * u.nTpad + t];
    auto test = [&](int q, bool viaX) {
        const BbQuery& d = qs[q];
        const double qx = __dmul_rn(__dsub_rn(h.x, d.minX), d.invRes);
        const double qy = __dmul_rn(__dsub_rn(h.y, d.minY), d.invRes);
        const double rx = qx - floor(qx), ry = qy - floor(qy);
        const bool edgeX = !(rx >= eps && rx <= 1.0 - eps), edgeY = !(ry >= eps && ry <= 1.0 - eps);
        if (viaX ? !edgeX : (!edgeY || edgeX)) return;
        const int f = atomicAdd(flagCount, 1);
        if (f < kFlagCapBB) flags[f] = BbFlag{q, t, i};
    };
    auto window = [&](const double* __restrict__ F, const int* __restrict__ Q, double lo, double hi, bool viaX) {
        int a = 0, b = u.sortCount;                          // first entry >= lo
        while (a < b) { const int m = (a + b) >> 1; if (F[u.sortBegin + m] < lo) a = m + 1; else b = m; }
        for (; a < u.sortCount && F[u.sortBegin + a] <= hi; ++a) test(Q[u.sortBegin + a], viaX);
    };
    auto axis = [&](double coord, const double* __restrict__ F, const int* __restrict__ Q, bool viaX) {
        const double a = __dmul_rn(coord, u.invRes);
        const double fa = a - floor(a);
        window(F, Q, fa - delta, fa + delta, viaX);
        if (fa - delta < 0.0) window(F, Q, fa - delta + 1.0, 2.0, viaX);     // wrap around 0 / 1
        if (fa + delta >= 1.0) window(F, Q, -1.0, fa + delta - 1.0, viaX);
    };
    axis(h.x, sortFx, sortQx, true);
    axis(h.y, sortFy, sortQy, false);
}

// ---- root level straight from the hit points ------------------------------------------------------------
// The root level visits every (query, theta), so it would read the whole per-query index table once;
// instead it converts the shared hit points on the fly (the reference's expression in double: every
// point outside the eps band floors identically, the <= kFlagInline points inside it come from the
// host's tables) and only the (query, theta) rows that survive get table rows (bb_index_slots_kernel).
// A survivor's row slot is childBase / 4: children are allocated four per survivor.
template <int U>
__global__ void __launch_bounds__(128)
bb_score_root_kernel(const BbQuery* __restrict__ qs, const BbScan* __restrict__ scans,
                     const double2* __restrict__ hits, const BbFlag* __restrict__ flags, int nFlag,
                     const int* __restrict__ exactIdx, int exactSpanX, int exactSpanY, int height,
                     Node* __restrict__ nodes, double* __restrict__ scores, int nNodes,
                     Node* __restrict__ next, int nextCap, int* __restrict__ nextCount,
                     int2* __restrict__ slotQT, int* __restrict__ slotOut, BbBest* __restrict__ best) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < nNodes;
    Node n;
    bool survive = false;
    if (active) {
        n = nodes[k];
        const BbQuery& d = qs[n.q];
        const double* __restrict__ lvl = d.level[height];
        const int stride = d.nTpad;
        const double2* __restrict__ hb = hits + scans[d.scan].hitBegin + n.t;
        const int pitch = d.pitch, gx = d.nx, gy = d.ny, nb = d.nUse;
        const double minX = d.minX, minY = d.minY, invRes = d.invRes;
        const int ox = n.x - d.offX, oy = n.y - d.offY;       // node offset minus window origin
        bool hasFlag = false;
        for (int f = 0; f < nFlag; ++f) hasFlag |= flags[f].q == n.q && flags[f].t == n.t;
        double acc = 0.0;
        auto cellAt = [&](int i) -> const double* {           // slow path: a flagged beam may be among them
            int ix, iy, ff = -1;
            for (int f = 0; f < nFlag; ++f)
                if (flags[f].q == n.q && flags[f].t == n.t && flags[f].i == i) ff = f;
            if (ff < 0) {
                const double2 h = __ldg(hb + (long long)i * stride);
                ix = __double2int_rd(__dmul_rn(__dsub_rn(h.x, minX), invRes)) + ox;
                iy = __double2int_rd(__dmul_rn(__dsub_rn(h.y, minY), invRes)) + oy;
            } else {
                const int* e = exactIdx + (long long)ff * (exactSpanX + exactSpanY);
                ix = e[n.x + d.winX];
                iy = e[exactSpanX + n.y + d.winY];
            }
            ix = min(max(ix, -1), gx);
            iy = min(max(iy, -1), gy);
            return lvl + (long long)iy * pitch + ix;
        };
        int i = 0;
        bool hasNaN = false;
        if (!hasFlag) {
#pragma unroll 1
            for (; i + U <= nb; i += U) {
                // three separate unrolled loops, so that U hit points, then U map cells are in flight
                // together (one fused loop lets the compiler serialise the two dependent latencies)
                double2 hp[U];
                long long off[U];
                double v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) hp[u] = __ldg(hb + (long long)(i + u) * stride);
                // a value that depends on ALL loads gates the rest, so the U loads are issued back to
                // back (the full-table kernel gets the same effect from its sentinel test)
                bool odd = false;
#pragma unroll
                for (int u = 0; u < U; ++u) odd |= !(hp[u].x == hp[u].x);
                if (odd) { hasNaN = true; break; }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int ix = min(max(__double2int_rd(__dmul_rn(__dsub_rn(hp[u].x, minX), invRes)) + ox, -1), gx);
                    const int iy = min(max(__double2int_rd(__dmul_rn(__dsub_rn(hp[u].y, minY), invRes)) + oy, -1), gy);
                    off[u] = (long long)iy * pitch + ix;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = __ldg(lvl + off[u]);
#pragma unroll
                for (int u = 0; u < U; ++u) acc = __dadd_rn(acc, v[u]);   // unknown cells add 0.0
            }
        }
        (void)hasNaN;                        // a NaN hit point simply continues on the one-beam path below
        for (; i < nb; ++i) acc = __dadd_rn(acc, __ldg(cellAt(i)));
        scores[k] = acc;
        if (acc > d.thrAbs) {                                     // :108 with scoreMax >= threshold
            if (height == 0)
                atomicMax(&best[n.q].scoreBits, (unsigned long long)__double_as_longlong(acc));
            else
                survive = true;
        } else {
            nodes[k].childBase = -1;
        }
    }
    if (height == 0) return;
    const unsigned m = __ballot_sync(0xffffffffu, survive);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int cnt = __popc(m);
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(nextCount, 4 * cnt);
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (!survive) return;
    const int r = __popc(m & ((1u << lane) - 1u));
    nodes[k].childBase = base + r;
    nodes[k].childStride = cnt;
    if (base + 4 * cnt > nextCap) return;       // the pool is too small: this run is repeated
    const int slot = (base >> 2) + r;           // nextCount only ever grows by multiples of four
    slotQT[slot] = make_int2(n.q, n.t);
    const int w = 1 << (height - 1);
    const int dx[4] = {w, 0, w, 0}, dy[4] = {w, w, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        Node ch;
        ch.x = (short)(n.x + dx[c]); ch.y = (short)(n.y + dy[c]); ch.t = n.t; ch.q = n.q;
        ch.rank = n.rank * 4 + c;
        ch.parent = k; ch.childBase = -1; ch.childStride = 0;
        next[base + c * cnt + r] = ch;
        slotOut[base + c * cnt + r] = slot;
    }
}

// ---- index rows of the surviving (query, theta) pairs ------------------------------------------------------
// tab2[beam * slotStride + slot] = cell of the beam at node offset (0, 0); slots of neighbouring thetas
// are neighbours (warp-aggregated allocation), so the deeper levels' table reads stay coalesced.
__global__ void bb_index_slots_kernel(const BbQuery* __restrict__ qs, const BbScan* __restrict__ scans,
                                      const double2* __restrict__ hits, const int2* __restrict__ slotQT,
                                      const int* __restrict__ childCount, int slotStride,
                                      const BbFlag* __restrict__ flags, int nFlag, int2* __restrict__ tab2) {
    constexpr int kBeamsPerThread = 8;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= min(__ldg(childCount) >> 2, slotStride)) return;
    const int2 qt = slotQT[slot];
    const BbQuery& d = qs[qt.x];
    const double2* __restrict__ hb = hits + scans[d.scan].hitBegin + qt.y;
    const double minX = d.minX, minY = d.minY, invRes = d.invRes;
    const int offX = d.offX, offY = d.offY, nb = d.nUse, stride = d.nTpad;
#pragma unroll
    for (int j = 0; j < kBeamsPerThread; ++j) {
        const int i = blockIdx.y * kBeamsPerThread + j;
        if (i >= nb) break;
        const double2 h = __ldg(hb + (long long)i * stride);
        int2 v = make_int2(__double2int_rd(__dmul_rn(__dsub_rn(h.x, minX), invRes)) - offX,
                           __double2int_rd(__dmul_rn(__dsub_rn(h.y, minY), invRes)) - offY);
        for (int f = 0; f < nFlag; ++f)
            if (flags[f].q == qt.x && flags[f].t == qt.y && flags[f].i == i) v = make_int2(INT_MIN, f);
        tab2[(long long)i * slotStride + slot] = v;
    }
}

// ---- roots --------------------------------------------------------------------------------------
__global__ void bb_roots_kernel(const BbQuery* __restrict__ qs, int nq, int height,
                                Node* __restrict__ pool) {
    const int q = blockIdx.y;
    const BbQuery& d = qs[q];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int nRoots = d.nrx * d.nry * d.nT;
    if (k >= nRoots) return;
    // push order: x asc, y asc, theta asc (scan_matcher_branch_bound.cpp:85-88); LIFO pops reverse it.
    const int t = k % d.nT;
    const int ky = (k / d.nT) % d.nry;
    const int kx = k / (d.nT * d.nry);
    Node n;
    n.x = (short)(-d.winX + (kx << height));
    n.y = (short)(-d.winY + (ky << height));
    n.t = t; n.q = q;
    n.rank = (long long)(nRoots - 1 - k);
    n.parent = -1; n.childBase = -1; n.childStride = 0;
    pool[d.rootBegin + k] = n;
}

// ---- node scoring + expansion (hot kernel) -------------------------------------------------------
// One thread per node; ScorePixelAccurate::Score on pyramid level `height`, summed in beam order.
// Survivors of a warp allocate their children together (one atomic per warp) and store them
// child-major, so the next level's lanes again walk neighbouring thetas with equal offsets.
template <int U>                           // beams in flight per thread (two dependent latencies each)
__global__ void __launch_bounds__(128)
bb_score_kernel(const BbQuery* __restrict__ qs, const int2* __restrict__ tab,
                const int* __restrict__ exactIdx, int exactSpanX, int exactSpanY, int height,
                Node* __restrict__ nodes, double* __restrict__ scores, int nMax,
                const int* __restrict__ nDev, Node* __restrict__ next, int nextCap,
                int* __restrict__ nextCount, BbBest* __restrict__ best,
                const int* __restrict__ slotIn, int* __restrict__ slotOut, int slotStride) {
    // slotIn != nullptr: `tab` holds one row per surviving (query, theta) (bb_index_slots_kernel) and
    // slotIn[k] is node k's row; otherwise `tab` is the full per-query table.
    // nDev (speculative, sync-free runs): the level's node count lives on the device; the launch
    // covers the pool capacity nMax and surplus threads leave here
    const int nNodes = nDev ? min(__ldg(nDev), nMax) : nMax;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < nNodes;
    Node n;
    bool survive = false;
    if (active) {
        n = nodes[k];
        const BbQuery& d = qs[n.q];
        const double* __restrict__ lvl = d.level[height];
        const int stride = slotIn ? slotStride : d.nTpad;
        const int2* __restrict__ tb = slotIn ? tab + slotIn[k] : tab + d.tabBegin + n.t;
        const int pitch = d.pitch, gx = d.nx, gy = d.ny, nb = d.nUse;
        const int nxo = n.x, nyo = n.y;
        double acc = 0.0;
        auto cellOf = [&](const int2 c) -> const double* {
            int ix, iy;
            if (c.x != INT_MIN) {
                ix = c.x + nxo; iy = c.y + nyo;
            } else {   // near-edge beam: indices from the host-computed exact table
                const int* e = exactIdx + (long long)c.y * (exactSpanX + exactSpanY);
                ix = e[nxo + d.winX];
                iy = e[exactSpanX + nyo + d.winY];
            }
            ix = min(max(ix, -1), gx);          // out of the map -> zero apron (Value(idx, unknown))
            iy = min(max(iy, -1), gy);
            return lvl + (long long)iy * pitch + ix;
        };
        int i = 0;
#pragma unroll 1
        for (; i + U <= nb; i += U) {
            int2 c[U];
            double v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) c[u] = __ldg(tb + (long long)(i + u) * stride);
            bool exact = false;
#pragma unroll
            for (int u = 0; u < U; ++u) exact |= c[u].x == INT_MIN;
            if (!exact) {      // branch free: all U gathers in flight together
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int ix = min(max(c[u].x + nxo, -1), gx);
                    const int iy = min(max(c[u].y + nyo, -1), gy);
                    v[u] = __ldg(lvl + (long long)iy * pitch + ix);
                }
            } else {
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = __ldg(cellOf(c[u]));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc = __dadd_rn(acc, v[u]);   // unknown cells add 0.0
        }
        for (; i < nb; ++i) acc = __dadd_rn(acc, __ldg(cellOf(__ldg(tb + (long long)i * stride))));
        scores[k] = acc;
        if (acc > d.thrAbs) {                                     // :108 with scoreMax >= threshold
            if (height == 0)
                atomicMax(&best[n.q].scoreBits, (unsigned long long)__double_as_longlong(acc));
            else
                survive = true;
        } else {
            nodes[k].childBase = -1;
        }
    }
    if (height == 0) return;
    const unsigned m = __ballot_sync(0xffffffffu, survive);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int cnt = __popc(m);
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(nextCount, 4 * cnt);
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (!survive) return;
    const int r = __popc(m & ((1u << lane) - 1u));
    nodes[k].childBase = base + r;
    nodes[k].childStride = cnt;
    if (base + 4 * cnt > nextCap) return;       // host grows the pool and re-runs this level
    const int w = 1 << (height - 1);
    // visit (pop) order: (x+w, y+w), (x, y+w), (x+w, y), (x, y)   (:134-137)
    const int dx[4] = {w, 0, w, 0}, dy[4] = {w, w, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        Node ch;
        ch.x = (short)(n.x + dx[c]); ch.y = (short)(n.y + dy[c]); ch.t = n.t; ch.q = n.q;
        ch.rank = n.rank * 4 + c;
        ch.parent = k; ch.childBase = -1; ch.childStride = 0;
        next[base + c * cnt + r] = ch;
        if (slotIn) slotOut[base + c * cnt + r] = slotIn[k];
    }
}

// ---- node scoring, one WARP per node (small levels) ------------------------------------------------
// A level with few nodes cannot hide the two dependent memory latencies (index table, then map
// cell) of a 1000-beam walk with one thread per node.  Here the 32 lanes fetch the node's beams 512
// at a time (16 independent load pairs in flight per lane), park the values in shared memory in
// beam order, and lane 0 then adds them strictly in that order: still bit-identical to the CPU sum,
// but a level costs ~2 memory latencies + one dependent add chain instead of ~70 latencies.
__global__ void __launch_bounds__(128)
bb_score_warp_kernel(const BbQuery* __restrict__ qs, const int2* __restrict__ tab,
                     const int* __restrict__ exactIdx, int exactSpanX, int exactSpanY, int height,
                     Node* __restrict__ nodes, double* __restrict__ scores, int nMax,
                     const int* __restrict__ nDev, Node* __restrict__ next, int nextCap,
                     int* __restrict__ nextCount, BbBest* __restrict__ best,
                     const int* __restrict__ slotIn, int* __restrict__ slotOut, int slotStride) {
    constexpr int CH = 16, STAGE = 32 * CH;
    const int nNodes = nDev ? min(__ldg(nDev), nMax) : nMax;
    __shared__ double sv[4][STAGE];
    const int wib = threadIdx.x >> 5;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= nNodes) return;                       // whole warp
    const Node n = nodes[k];
    const BbQuery& d = qs[n.q];
    const double* __restrict__ lvl = d.level[height];
    const int stride = slotIn ? slotStride : d.nTpad;
    const int2* __restrict__ tb = slotIn ? tab + slotIn[k] : tab + d.tabBegin + n.t;
    const int pitch = d.pitch, gx = d.nx, gy = d.ny, nb = d.nUse;
    const int nxo = n.x, nyo = n.y;
    double acc = 0.0;
#pragma unroll 1
    for (int base = 0; base < nb; base += STAGE) {
        int2 c[CH];
        double v[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int i = base + u * 32 + lane;
            // a beam past the end projects far outside -> clamps into the zero apron (adds 0.0)
            c[u] = i < nb ? __ldg(tb + (long long)i * stride) : make_int2(-(1 << 28), -(1 << 28));
        }
        bool exact = false;
#pragma unroll
        for (int u = 0; u < CH; ++u) exact |= c[u].x == INT_MIN;
        if (!__any_sync(0xffffffffu, exact)) {      // branch free: all gathers in flight together
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                const int ix = min(max(c[u].x + nxo, -1), gx);
                const int iy = min(max(c[u].y + nyo, -1), gy);
                v[u] = __ldg(lvl + (long long)iy * pitch + ix);
            }
        } else {
#pragma unroll
            for (int u = 0; u < CH; ++u) {
                int ix, iy;
                if (c[u].x != INT_MIN) {
                    ix = c[u].x + nxo; iy = c[u].y + nyo;
                } else {   // near-edge beam: indices from the host-computed exact table
                    const int* e = exactIdx + (long long)c[u].y * (exactSpanX + exactSpanY);
                    ix = e[nxo + d.winX];
                    iy = e[exactSpanX + nyo + d.winY];
                }
                ix = min(max(ix, -1), gx);
                iy = min(max(iy, -1), gy);
                v[u] = __ldg(lvl + (long long)iy * pitch + ix);
            }
        }
#pragma unroll
        for (int u = 0; u < CH; ++u) sv[wib][u * 32 + lane] = v[u];
        __syncwarp();
        if (lane == 0) {
            const int m = min(STAGE, nb - base);
            int j = 0;
            for (; j + 8 <= m; j += 8) {
                double t8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t8[u] = sv[wib][j + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, t8[u]);      // beam order
            }
            for (; j < m; ++j) acc = __dadd_rn(acc, sv[wib][j]);
        }
        __syncwarp();
    }
    if (lane != 0) return;
    scores[k] = acc;
    if (!(acc > d.thrAbs)) { nodes[k].childBase = -1; return; }
    if (height == 0) {
        atomicMax(&best[n.q].scoreBits, (unsigned long long)__double_as_longlong(acc));
        return;
    }
    const int slot = atomicAdd(nextCount, 4);
    nodes[k].childBase = slot;
    nodes[k].childStride = 1;
    if (slot + 4 > nextCap) return;
    const int w = 1 << (height - 1);
    const int dx[4] = {w, 0, w, 0}, dy[4] = {w, w, 0, 0};
#pragma unroll
    for (int cidx = 0; cidx < 4; ++cidx) {
        Node ch;
        ch.x = (short)(n.x + dx[cidx]); ch.y = (short)(n.y + dy[cidx]); ch.t = n.t; ch.q = n.q;
        ch.rank = n.rank * 4 + cidx;
        ch.parent = k; ch.childBase = -1; ch.childStride = 0;
        next[slot + cidx] = ch;
        if (slotIn) slotOut[slot + cidx] = slotIn[k];
    }
}

// ---- winner among the leaves: (score desc, rank asc) ---------------------------------------------
__global__ void bb_leaf_rank_kernel(const Node* __restrict__ leaves, const double* __restrict__ scores,
                                    int nMax, const int* __restrict__ nDev, BbBest* __restrict__ best) {
    const int n = nDev ? min(__ldg(nDev), nMax) : nMax;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = leaves[k].q;
    if ((unsigned long long)__double_as_longlong(scores[k]) == best[q].scoreBits && best[q].scoreBits != 0ull)
        atomicMin((unsigned long long*)&best[q].rank, (unsigned long long)leaves[k].rank);
}

__global__ void bb_leaf_pick_kernel(const Node* __restrict__ leaves, const double* __restrict__ scores,
                                    int nMax, const int* __restrict__ nDev, BbBest* __restrict__ best) {
    const int n = nDev ? min(__ldg(nDev), nMax) : nMax;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = leaves[k].q;
    if ((unsigned long long)__double_as_longlong(scores[k]) == best[q].scoreBits &&
        best[q].scoreBits != 0ull && leaves[k].rank == best[q].rank)
        best[q].leaf = k;
}

struct LevelView { const Node* nodes; const double* scores; };
struct LevelViews { LevelView v[kMaxLevels]; };

// ---- verification of the winner's ancestor chain + result record -----------------------------------
__global__ void bb_verify_kernel(const BbQuery* __restrict__ qs, int nq, int heightMax, LevelViews lv,
                                 BbBest* __restrict__ best, BbResult* __restrict__ res, int forceReplay) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    BbResult r;
    r.pad = 0; r.exactReplay = 0;
    BbBest b = best[q];
    if (b.scoreBits == 0ull || b.leaf < 0) {
        // No superset leaf above the threshold: the CPU cannot accept any leaf either.
        r.found = 0; r.score = qs[q].thrAbs; r.ix = 0; r.iy = 0; r.it = 0;
        res[q] = r;
        best[q].needReplay = 0;
        return;
    }
    const double s = __longlong_as_double((long long)b.scoreBits);
    int idx = b.leaf;
    bool ok = true;
    for (int h = 0; h < heightMax; ++h) {
        idx = lv.v[h].nodes[idx].parent;
        if (lv.v[h + 1].scores[idx] < s) ok = false;
    }
    const Node leaf = lv.v[0].nodes[b.leaf];
    r.found = 1; r.score = s; r.ix = leaf.x; r.iy = leaf.y; r.it = leaf.t - qs[q].winT;
    res[q] = r;
    best[q].needReplay = (!ok || forceReplay) ? 1 : 0;
}

// ---- sequential replay of the CPU's LIFO search over the stored superset scores --------------------
__global__ void bb_replay_kernel(const BbQuery* __restrict__ qs, int nq, int heightMax, LevelViews lv,
                                 const BbBest* __restrict__ best, BbResult* __restrict__ res) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq || !best[q].needReplay) return;
    const BbQuery& d = qs[q];
    const int nRoots = d.nrx * d.nry * d.nT;
    double bestScore = d.thrAbs;
    int bestLeaf = -1;
    int stackIdx[4 * kMaxLevels];
    int stackH[4 * kMaxLevels];
    for (int rv = 0; rv < nRoots; ++rv) {              // roots in pop order: rank == rv
        int sp = 0;
        stackIdx[sp] = d.rootBegin + (nRoots - 1 - rv); stackH[sp] = heightMax; ++sp;
        while (sp > 0) {
            --sp;
            const int idx = stackIdx[sp], h = stackH[sp];
            const double s = lv.v[h].scores[idx];
            if (s <= bestScore) continue;                               // :108
            if (h == 0) { bestScore = s; bestLeaf = idx; continue; }    // :114-120
            const Node nd = lv.v[h].nodes[idx];                         // s > best >= thr => expanded
            for (int c = 3; c >= 0; --c) { stackIdx[sp] = nd.childBase + c * nd.childStride; stackH[sp] = h - 1; ++sp; }
        }
    }
    BbResult r;
    r.pad = 0; r.exactReplay = 1;
    if (bestLeaf >= 0) {
        const Node leaf = lv.v[0].nodes[bestLeaf];
        r.found = 1; r.score = bestScore; r.ix = leaf.x; r.iy = leaf.y; r.it = leaf.t - d.winT;
    } else {
        r.found = 0; r.score = d.thrAbs; r.ix = 0; r.iy = 0; r.it = 0;
    }
    res[q] = r;
}

__global__ void bb_init_best_kernel(BbBest* best, int nq) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    best[q].scoreBits = 0ull; best[q].rank = 0x7fffffffffffffffLL; best[q].leaf = -1; best[q].needReplay = 0;
}

}  // namespace

struct lgs_bb_batch {
    lgs_ctx* ctx = nullptr;
    lgs_bb_params params{};
    int nq = 0, H = 0;
    int maxRoots = 0;
    int maxNTpad = 0, maxUse = 0;   // launch extents of the projection kernels
    int spanX = 0, spanY = 0;
    std::vector<BbQuery> qs;
    std::vector<BbScan> us;         // distinct (scan, pose) pairs
    std::vector<IdxChunk> chunks;   // <= kIdxChunk queries of one scan each
    std::vector<int> qlist;         // queries grouped by scan
    DevBuf<IdxChunk> dChunks;
    DevBuf<int> dQlist;
    std::vector<double> hAngles, hRanges;
    std::vector<int> fixups;
    long long nTab = 0;
    int totalRoots = 0;
    bool uploaded = false, ran = false, forceReplay = false;
    // Speculative (sync-free) runs: launches are sized by the node pools' capacities and read the
    // level counts on the device; lgs_bb_batch_results validates (no pool overflow) and otherwise
    // repeats the run level-synchronously.  Hints = counts of the last run.
    bool haveHints = false, pendingValidate = false;
    cudaGraphExec_t graphExec = nullptr;  // LGS_BB_GRAPH: the speculative level chain, updated in place run after run
    double hostMs[3] = {0.0, 0.0, 0.0};   // LGS_BB_HOSTTIMING: preamble enqueue, flag-count wait, level enqueue
    long long hostRuns = 0;
    long long hint[kMaxLevels] = {0};
    long long nodesPerLevel[kMaxLevels] = {0};
    long long gathers = 0;
    DevBuf<BbQuery> dQs;
    DevBuf<BbScan> dUs;
    DevBuf<double2> dHits;
    DevBuf<double> dAngles, dRanges;
    DevBuf<int2> dTab;               // full per-query index table (only the many-flags fallback path)
    DevBuf<int2> dTab2, dSlotQT;     // rows of the surviving (query, theta) pairs; their (q, t)
    DevBuf<int> dSlot[kMaxLevels];   // row slot of every node of a level
    DevBuf<double> dSortFx, dSortFy; // per scan: its queries sorted by frac(min * invRes), per axis
    DevBuf<int> dSortQx, dSortQy;
    std::vector<double> hSortFx, hSortFy;
    std::vector<int> hSortQx, hSortQy;
    bool slotPath = false;           // the last run scored the root level from the hit points
    int nFlagRun = 0;                // near-edge points of the last run
    int slotLaunch = 0;              // slots the speculative run built index rows for
    DevBuf<BbFlag> dFlags;
    DevBuf<int> dCounters;          // [0] flag count, [1 + h] node count of level h
    DevBuf<int> dExact;
    DevBuf<Node> dNodes[kMaxLevels];
    DevBuf<double> dScores[kMaxLevels];
    DevBuf<BbBest> dBest;
    DevBuf<BbResult> dRes;
    PinBuf<BbResult> hRes;
    PinBuf<int> hCounters;
};

extern "C" {

int lgs_bb_batch_create(lgs_ctx* ctx, const lgs_bb_params* p, lgs_bb_batch** out) {
    if (!ctx || !p || !out) return LGS_ERR_INVALID;
    *out = nullptr;
    if (p->node_height_max < 0 || p->node_height_max >= kMaxLevels - 1 || !(p->range_x >= 0) ||
        !(p->range_y >= 0) || !(p->range_theta >= 0))
        return lgs_fail(ctx, LGS_ERR_INVALID, "bb_batch_create: bad parameters");
    lgs_bb_batch* b = new lgs_bb_batch();
    b->ctx = ctx; b->params = *p; b->H = p->node_height_max;
    *out = b;
    return LGS_OK;
}

int lgs_bb_batch_destroy(lgs_bb_batch* b) {
    if (b && b->hostRuns > 0 && getenv("LGS_BB_HOSTTIMING"))
        fprintf(stderr, "[lgs bb host] %lld speculative runs of %d queries: per run preamble enqueue %.3f ms, flag-count "
                "wait %.3f ms, level enqueue %.3f ms\n", b->hostRuns, b->nq, b->hostMs[0] / b->hostRuns,
                b->hostMs[1] / b->hostRuns, b->hostMs[2] / b->hostRuns);
    if (!b) return LGS_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->graphExec) cudaGraphExecDestroy(b->graphExec);
    b->dQs.release(); b->dUs.release(); b->dHits.release(); b->dChunks.release(); b->dQlist.release(); b->dAngles.release(); b->dRanges.release(); b->dTab.release(); b->dTab2.release(); b->dSlotQT.release();
    for (auto& sl : b->dSlot) sl.release();
    b->dSortFx.release(); b->dSortFy.release(); b->dSortQx.release(); b->dSortQy.release();
    b->dFlags.release(); b->dCounters.release(); b->dExact.release(); b->dBest.release();
    b->dRes.release(); b->hRes.release(); b->hCounters.release();
    for (int h = 0; h < kMaxLevels; ++h) { b->dNodes[h].release(); b->dScores[h].release(); }
    delete b;
    return LGS_OK;
}

int lgs_bb_batch_force_replay(lgs_bb_batch* b, int on) {
    if (!b) return LGS_ERR_INVALID;
    b->forceReplay = on != 0;
    return LGS_OK;
}

int lgs_bb_batch_upload(lgs_bb_batch* b, const lgs_scan_batch* scans, lgs_pyramid* const* pyramids,
                        const double* normThr) {
    if (!b || !scans) return LGS_ERR_INVALID;
    return lgs_bb_batch_upload_pairs(b, scans, scans->n_scans, nullptr, pyramids, normThr);
}

int lgs_bb_batch_upload_pairs(lgs_bb_batch* b, const lgs_scan_batch* scans, int nPairs,
                              const int* pairScan, lgs_pyramid* const* pyramids,
                              const double* normThr) {
    if (!b || !scans) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    const int n = nPairs;
    if (n < 0 || scans->n_scans < 0 || (n > 0 && (!scans->beam_begin || !scans->sensor_pose || !pyramids)))
        return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_upload: bad arguments");
    for (int q = 0; q < n; ++q) {
        const int sq = pairScan ? pairScan[q] : q;
        if (sq < 0 || sq >= scans->n_scans)
            return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_upload: pair %d names scan %d of %d", q, sq, scans->n_scans);
    }
    LGS_CUDA(c, cudaSetDevice(c->device));
    const lgs_bb_params& p = b->params;
    const int H = b->H;
    b->uploaded = false; b->ran = false;
    b->nq = n;
    b->qs.assign(n, BbQuery{});
    b->us.clear();
    b->fixups.assign(n, 0);
    b->hAngles.clear(); b->hRanges.clear();
    b->maxRoots = 0; b->maxNTpad = 0; b->maxUse = 0; b->spanX = 0; b->spanY = 0;
    long long nTab = 0, roots = 0, nHits = 0;
    const int winSizeMax = 1 << H;
    // Pairs that name the same scan share its projected hit points (1 scan x many submaps).
    std::vector<int> scanToUnique(std::max(scans->n_scans, 1), -1);
    std::vector<double> scanMaxR(std::max(scans->n_scans, 1), 0.0);
    std::vector<char> haveMaxR(std::max(scans->n_scans, 1), 0);
    for (int q = 0; q < n; ++q) {
        const int sq = pairScan ? pairScan[q] : q;
        const int b0 = scans->beam_begin[sq], b1 = scans->beam_begin[sq + 1];
        const int nb = b1 - b0;
        if (nb <= 0) return lgs_fail(c, LGS_ERR_INVALID, "bb: scan %d has no beams", q);
        const lgs_pyramid* pyr = pyramids[q];
        if (!pyr || lgs_pyramid_levels(pyr) < H + 1)
            return lgs_fail(c, LGS_ERR_INVALID, "bb: query %d needs a pyramid with %d levels", q, H + 1);
        const lgs_grid* g0 = lgs_pyramid_level(pyr, 0);
        lgs_pyramid_note_user(pyr, c);
        if (g0->ctx->device != c->device)
            return lgs_fail(c, LGS_ERR_INVALID, "bb: pyramid of query %d lives on another device", q);
        BbQuery& d = b->qs[q];
        d.minX = g0->min_x; d.minY = g0->min_y; d.res = g0->res; d.invRes = 1.0 / g0->res;
        d.nx = g0->nx; d.ny = g0->ny; d.pitch = g0->pitch;
        d.offX = g0->off_x; d.offY = g0->off_y;
        for (int h = 0; h <= H; ++h) d.level[h] = lgs_pyramid_level(pyr, h)->origin();
        // ComputeSearchStep (scan_matcher_branch_bound.cpp:178-197)
        if (!haveMaxR[sq]) {           // std::max_element over the scan, once per scan
            double m = scans->ranges[b0];
            for (int i = b0 + 1; i < b1; ++i) m = std::max(m, scans->ranges[i]);
            scanMaxR[sq] = m;
            haveMaxR[sq] = 1;
        }
        const double maxR = scanMaxR[sq];
        const double maxRange = std::min(maxR, p.scan_range_max);
        const double th = d.res / maxRange;
        const double stepX = d.res, stepY = d.res;
        const double stepT = std::acos(1.0 - 0.5 * th * th);
        d.winX = static_cast<int>(std::ceil(0.5 * p.range_x / stepX));                 // :68-73
        d.winY = static_cast<int>(std::ceil(0.5 * p.range_y / stepY));
        d.winT = static_cast<int>(std::ceil(0.5 * p.range_theta / stepT));
        if (!(stepT > 0.0) || d.winT < 0 || d.winT > (1 << 20))
            return lgs_fail(c, LGS_ERR_INVALID, "bb: scan %d gives stepTheta=%g winTheta=%d", q, stepT, d.winT);
        if (d.winX + 2 * winSizeMax > 32000 || d.winY + 2 * winSizeMax > 32000)
            return lgs_fail(c, LGS_ERR_INVALID, "bb: search window too large for 16-bit node offsets");
        d.nT = 2 * d.winT + 1;
        d.nTpad = (d.nT + 3) / 4 * 4;
        d.nrx = (2 * d.winX) / winSizeMax + 1;                                         // :85-86
        d.nry = (2 * d.winY) / winSizeMax + 1;
        const double thr = normThr ? normThr[q] : DBL_MIN;
        d.thrAbs = thr * static_cast<double>(static_cast<size_t>(nb));                 // :75-76
        // ScorePixelAccurate range filter (score_function_pixel_accurate.cpp:27-41)
        const double sMin = scans->range_min ? scans->range_min[sq] : 0.0;
        const double sMax = scans->range_max ? scans->range_max[sq] : HUGE_VAL;
        const double minRange = std::max(p.score_range_min, sMin);
        const double maxRangeS = std::min(p.score_range_max, sMax);
        int uidx = scanToUnique[sq];
        if (uidx >= 0 && (b->us[uidx].stepT != stepT || b->us[uidx].invRes != d.invRes)) uidx = -1;   // other map resolution
        if (uidx < 0) {
            BbScan u{};
            u.sx = scans->sensor_pose[3 * sq]; u.sy = scans->sensor_pose[3 * sq + 1];
            u.st = scans->sensor_pose[3 * sq + 2];
            u.stepT = stepT; u.winT = d.winT; u.nT = d.nT; u.nTpad = d.nTpad;
            u.invRes = d.invRes; u.sortBegin = 0; u.sortCount = 0;
            u.beamBegin = (int)b->hAngles.size();
            for (int i = b0; i < b1; ++i) {
                const double r = scans->ranges[i];
                if (r >= maxRangeS || r <= minRange) continue;
                b->hAngles.push_back(scans->angles[i]);
                b->hRanges.push_back(r);
            }
            u.nUse = (int)b->hAngles.size() - u.beamBegin;
            u.hitBegin = nHits;
            nHits += (long long)u.nUse * u.nTpad;
            uidx = (int)b->us.size();
            b->us.push_back(u);
            scanToUnique[sq] = uidx;
        }
        d.scan = uidx;
        d.nUse = b->us[uidx].nUse;
        d.tabBegin = nTab;
        nTab += (long long)d.nUse * d.nTpad;
        const long long nr = (long long)d.nrx * d.nry * d.nT;
        if (roots + nr > (1LL << 30)) return lgs_fail(c, LGS_ERR_INVALID, "bb: too many root nodes");
        d.rootBegin = (int)roots;
        roots += nr;
        b->maxNTpad = std::max(b->maxNTpad, d.nTpad);
        b->maxUse = std::max(b->maxUse, d.nUse);
        b->maxRoots = std::max<long long>(b->maxRoots, nr);
        b->spanX = std::max(b->spanX, d.nrx * winSizeMax);
        b->spanY = std::max(b->spanY, d.nry * winSizeMax);
    }
    b->nTab = nTab;
    b->totalRoots = (int)roots;
    {   // group the queries by distinct scan, kIdxChunk per index-kernel block
        std::vector<std::vector<int>> byScan(b->us.size());
        for (int q = 0; q < n; ++q) byScan[b->qs[q].scan].push_back(q);
        b->qlist.clear(); b->chunks.clear();
        // flag search: the queries of every scan sorted by the fractional part of min * invRes
        b->hSortFx.clear(); b->hSortFy.clear(); b->hSortQx.clear(); b->hSortQy.clear();
        for (size_t u = 0; u < byScan.size(); ++u) {
            std::vector<std::pair<double, int>> fx, fy;
            for (int q : byScan[u]) {
                const BbQuery& d = b->qs[q];
                const double ax = d.minX * d.invRes, ay = d.minY * d.invRes;
                fx.emplace_back(ax - std::floor(ax), q);
                fy.emplace_back(ay - std::floor(ay), q);
            }
            std::sort(fx.begin(), fx.end());
            std::sort(fy.begin(), fy.end());
            b->us[u].sortBegin = (int)b->hSortFx.size();
            b->us[u].sortCount = (int)fx.size();
            for (size_t k = 0; k < fx.size(); ++k) {
                b->hSortFx.push_back(fx[k].first); b->hSortQx.push_back(fx[k].second);
                b->hSortFy.push_back(fy[k].first); b->hSortQy.push_back(fy[k].second);
            }
        }
        for (size_t u = 0; u < byScan.size(); ++u)
            for (size_t k = 0; k < byScan[u].size(); k += kIdxChunk) {
                const int cnt = (int)std::min<size_t>(kIdxChunk, byScan[u].size() - k);
                b->chunks.push_back(IdxChunk{(int)u, (int)b->qlist.size(), cnt});
                b->qlist.insert(b->qlist.end(), byScan[u].begin() + k, byScan[u].begin() + k + cnt);
            }
    }
    if (n == 0) { b->uploaded = true; return LGS_OK; }
    const size_t nk = b->hAngles.size();
    LGS_CUDA(c, b->dQs.reserve(n));
    LGS_CUDA(c, b->dUs.reserve(b->us.size()));
    LGS_CUDA(c, b->dChunks.reserve(std::max<size_t>(b->chunks.size(), 1)));
    LGS_CUDA(c, b->dQlist.reserve(std::max<size_t>(b->qlist.size(), 1)));
    LGS_CUDA(c, b->dHits.reserve(std::max<long long>(nHits, 1)));
    LGS_CUDA(c, b->dAngles.reserve(std::max<size_t>(nk, 1)));
    LGS_CUDA(c, b->dRanges.reserve(std::max<size_t>(nk, 1)));
    LGS_CUDA(c, b->dSortFx.reserve(std::max<size_t>(b->hSortFx.size(), 1)));
    LGS_CUDA(c, b->dSortFy.reserve(std::max<size_t>(b->hSortFx.size(), 1)));
    LGS_CUDA(c, b->dSortQx.reserve(std::max<size_t>(b->hSortFx.size(), 1)));
    LGS_CUDA(c, b->dSortQy.reserve(std::max<size_t>(b->hSortFx.size(), 1)));
    LGS_CUDA(c, b->dFlags.reserve(kFlagCapBB));
    LGS_CUDA(c, b->dCounters.reserve(2 + kMaxLevels));
    LGS_CUDA(c, b->hCounters.reserve(2 + kMaxLevels));
    LGS_CUDA(c, b->dBest.reserve(n));
    LGS_CUDA(c, b->dRes.reserve(n));
    LGS_CUDA(c, b->hRes.reserve(n));
    LGS_CUDA(c, b->dNodes[H].reserve(roots));
    LGS_CUDA(c, b->dScores[H].reserve(roots));
    LGS_CUDA(c, cudaMemcpyAsync(b->dQs.p, b->qs.data(), n * sizeof(BbQuery), cudaMemcpyHostToDevice, c->stream));
    LGS_CUDA(c, cudaMemcpyAsync(b->dUs.p, b->us.data(), b->us.size() * sizeof(BbScan), cudaMemcpyHostToDevice, c->stream));
    LGS_CUDA(c, cudaMemcpyAsync(b->dChunks.p, b->chunks.data(), b->chunks.size() * sizeof(IdxChunk), cudaMemcpyHostToDevice, c->stream));
    LGS_CUDA(c, cudaMemcpyAsync(b->dQlist.p, b->qlist.data(), b->qlist.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (!b->hSortFx.empty()) {
        const size_t ns = b->hSortFx.size();
        LGS_CUDA(c, cudaMemcpyAsync(b->dSortFx.p, b->hSortFx.data(), ns * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(b->dSortFy.p, b->hSortFy.data(), ns * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(b->dSortQx.p, b->hSortQx.data(), ns * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(b->dSortQy.p, b->hSortQy.data(), ns * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    if (nk) {
        LGS_CUDA(c, cudaMemcpyAsync(b->dAngles.p, b->hAngles.data(), nk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(b->dRanges.p, b->hRanges.data(), nk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));   // host vectors may be reused by the caller
    b->uploaded = true;
    return LGS_OK;
}

static int bb_run_impl(lgs_bb_batch* b, bool spec) {
    lgs_ctx* c = b->ctx;
    if (!b->uploaded) return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_run before upload");
    const int H = b->H, n = b->nq;
    for (int h = 0; h < kMaxLevels; ++h) b->nodesPerLevel[h] = 0;
    b->gathers = 0;
    if (n == 0) { b->ran = true; return LGS_OK; }
    LGS_CUDA(c, cudaSetDevice(c->device));
    // LGS_BB_HOSTTIMING=1 (diagnostic): host wall time of a run's three phases, summed per batch
    // object and printed when it is destroyed.
    const auto hostT0 = std::chrono::steady_clock::now();
    auto hostMsSince = [](std::chrono::steady_clock::time_point t) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    LGS_CUDA(c, cudaMemsetAsync(b->dCounters.p, 0, (2 + kMaxLevels) * sizeof(int), c->stream));
    // Scoring the root level from the hit points pays off when several queries share a scan (the
    // 16-byte hit points are then L2 hits); with one scan per query the 8-byte table rows are cheaper.
    const bool wantSlots = H >= 1 && getenv("LGS_BB_TABLE") == nullptr &&
                           ((size_t)n >= 4 * b->us.size() || getenv("LGS_BB_SLOTS") != nullptr);
    {
        const unsigned gx = (unsigned)((b->maxNTpad + 127) / 128), gy = (unsigned)std::max(b->maxUse, 1);
        for (size_t u0 = 0; u0 < b->us.size(); u0 += 65535) {
            const unsigned nu = (unsigned)std::min<size_t>(65535, b->us.size() - u0);
            bb_hit_kernel<<<dim3(gx, gy, nu), 128, 0, c->stream>>>(b->dUs.p + u0, b->dAngles.p, b->dRanges.p, b->dHits.p);
            LGS_LAUNCH_CHECK(c);
        }
        const double eps = g_lgs_edge_eps;
        for (size_t u0 = 0; wantSlots && u0 < b->us.size(); u0 += 65535) {
            const unsigned nu = (unsigned)std::min<size_t>(65535, b->us.size() - u0);
            bb_flagsearch_kernel<<<dim3(gx, gy, nu), 128, 0, c->stream>>>(
                b->dUs.p + u0, b->dQs.p, b->dHits.p, b->dSortFx.p, b->dSortQx.p, b->dSortFy.p, b->dSortQy.p,
                eps, eps + 1e-7, b->dFlags.p, b->dCounters.p);
            LGS_LAUNCH_CHECK(c);
        }
        dim3 gridR((b->maxRoots + 127) / 128, n);
        bb_roots_kernel<<<gridR, 128, 0, c->stream>>>(b->dQs.p, n, H, b->dNodes[H].p);
        LGS_LAUNCH_CHECK(c);
        bb_init_best_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(b->dBest.p, n);
        LGS_LAUNCH_CHECK(c);
    }
    const int spanX = b->spanX, spanY = b->spanY;
    {   // the one host round trip a speculative run keeps (the flag count is almost always 0)
        // Near-edge beams: exact per-offset index tables from the host (CPU expression + glibc).
        int nFlag = 0;
        b->slotPath = false;
        if (wantSlots) {
            LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p, b->dCounters.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            b->hostMs[0] += hostMsSince(hostT0);
            const auto w0 = std::chrono::steady_clock::now();
            LGS_CUDA(c, cudaStreamSynchronize(c->stream));
            b->hostMs[1] += hostMsSince(w0);
            nFlag = b->hCounters.p[0];
            b->slotPath = nFlag <= kFlagInline;
        }
        if (!b->slotPath) {
            // one scan per query, many near-edge points, or forced: the full per-query table, whose
            // kernel flags the near-edge points itself
            LGS_CUDA(c, b->dTab.reserve(std::max<long long>(b->nTab, 1)));
            LGS_CUDA(c, cudaMemsetAsync(b->dCounters.p, 0, sizeof(int), c->stream));
            const unsigned gx = (unsigned)((b->maxNTpad + 127) / 128), gy = (unsigned)std::max(b->maxUse, 1);
            for (size_t c0 = 0; c0 < b->chunks.size(); c0 += 65535) {
                const unsigned nc = (unsigned)std::min<size_t>(65535, b->chunks.size() - c0);
                bb_index_kernel<<<dim3(gx, gy, nc), 128, 0, c->stream>>>(b->dQs.p, b->dUs.p, b->dChunks.p + c0, b->dQlist.p,
                                                                         b->dHits.p, g_lgs_edge_eps, b->dTab.p,
                                                                         b->dFlags.p, b->dCounters.p);
                LGS_LAUNCH_CHECK(c);
            }
            LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p, b->dCounters.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            LGS_CUDA(c, cudaStreamSynchronize(c->stream));
            nFlag = b->hCounters.p[0];
        }
        b->nFlagRun = nFlag;
        std::fill(b->fixups.begin(), b->fixups.end(), 0);
        if (nFlag > kFlagCapBB)
            return lgs_fail(c, LGS_ERR_OVERFLOW, "bb: %d near-edge points exceed the fix-up list", nFlag);
        if (nFlag > 0) {
            std::vector<BbFlag> fl(nFlag);
            LGS_CUDA(c, cudaMemcpy(fl.data(), b->dFlags.p, nFlag * sizeof(BbFlag), cudaMemcpyDeviceToHost));
            std::vector<int> exact((size_t)nFlag * (spanX + spanY), 0);
            for (int k = 0; k < nFlag; ++k) {
                const BbQuery& d = b->qs[fl[k].q];
                const BbScan& u = b->us[d.scan];
                const double stepX = d.res, stepY = d.res;
                const double theta = u.st + static_cast<double>(fl[k].t - d.winT) * u.stepT;
                const double a = theta + b->hAngles[u.beamBegin + fl[k].i];
                const double cosT = std::cos(a), sinT = std::sin(a);
                const double r = b->hRanges[u.beamBegin + fl[k].i];
                int* e = exact.data() + (size_t)k * (spanX + spanY);
                for (int o = 0; o < spanX; ++o) {
                    const double px = u.sx + static_cast<double>(o - d.winX) * stepX;       // :96-97
                    e[o] = static_cast<int>(std::floor(((px + r * cosT) - d.minX) / d.res)) - d.offX;
                }
                for (int o = 0; o < spanY; ++o) {
                    const double py = u.sy + static_cast<double>(o - d.winY) * stepY;
                    e[spanX + o] = static_cast<int>(std::floor(((py + r * sinT) - d.minY) / d.res)) - d.offY;
                }
                b->fixups[fl[k].q]++;
            }
            LGS_CUDA(c, b->dExact.reserve(exact.size()));
            LGS_CUDA(c, cudaMemcpy(b->dExact.p, exact.data(), exact.size() * sizeof(int), cudaMemcpyHostToDevice));
        } else {
            LGS_CUDA(c, b->dExact.reserve(1));
        }

    }
    // Level-synchronous expansion of the static-threshold superset.
    int warpBelow = kWarpPerNodeBelow;
    int deepBelow = kDeepUnrollBelow;
    if (const char* e = getenv("LGS_BB_WARP_BELOW")) warpBelow = atoi(e);     // tuning hooks
    if (const char* e = getenv("LGS_BB_DEEP_BELOW")) deepBelow = atoi(e);
    const bool slots = b->slotPath;
    const int nFlag = b->nFlagRun;
    // One level: nLaunch threads / warps, the node count either exact (nDev == nullptr) or on the device.
    auto launchLevel = [&](int h, int nLaunch, const int* nDev, long long expect) -> int {
        LGS_CUDA(c, b->dScores[h].reserve(nLaunch));
        Node* next = h > 0 ? b->dNodes[h - 1].p : nullptr;
        const int nextCap = h > 0 ? (int)b->dNodes[h - 1].cap : 0;
        int* nextCount = b->dCounters.p + 1 + h;
        const int slotStride = (int)(b->dNodes[H - (H > 0 ? 1 : 0)].cap / 4);
        if (slots && h > 0) LGS_CUDA(c, b->dSlot[h - 1].reserve(std::max(nextCap, 1)));
        if (slots && h == H) {
            LGS_CUDA(c, b->dSlotQT.reserve(std::max(slotStride, 1)));
            bb_score_root_kernel<8><<<(nLaunch + 127) / 128, 128, 0, c->stream>>>(
                b->dQs.p, b->dUs.p, b->dHits.p, b->dFlags.p, nFlag, b->dExact.p, spanX, spanY, h, b->dNodes[h].p,
                b->dScores[h].p, nLaunch, next, nextCap, nextCount, b->dSlotQT.p, b->dSlot[h - 1].p, b->dBest.p);
            LGS_LAUNCH_CHECK(c);
            return LGS_OK;
        }
        const int2* tab = slots ? b->dTab2.p : b->dTab.p;
        const int* slotIn = slots ? b->dSlot[h].p : nullptr;
        int* slotOut = slots && h > 0 ? b->dSlot[h - 1].p : nullptr;
        if (expect >= warpBelow && expect >= deepBelow)
            bb_score_kernel<16><<<(nLaunch + 127) / 128, 128, 0, c->stream>>>(
                b->dQs.p, tab, b->dExact.p, spanX, spanY, h, b->dNodes[h].p, b->dScores[h].p,
                nLaunch, nDev, next, nextCap, nextCount, b->dBest.p, slotIn, slotOut, slotStride);
        else if (expect >= warpBelow)
            bb_score_kernel<32><<<(nLaunch + 127) / 128, 128, 0, c->stream>>>(
                b->dQs.p, tab, b->dExact.p, spanX, spanY, h, b->dNodes[h].p, b->dScores[h].p,
                nLaunch, nDev, next, nextCap, nextCount, b->dBest.p, slotIn, slotOut, slotStride);
        else
            bb_score_warp_kernel<<<(nLaunch + 3) / 4, 128, 0, c->stream>>>(
                b->dQs.p, tab, b->dExact.p, spanX, spanY, h, b->dNodes[h].p, b->dScores[h].p,
                nLaunch, nDev, next, nextCap, nextCount, b->dBest.p, slotIn, slotOut, slotStride);
        LGS_LAUNCH_CHECK(c);
        return LGS_OK;
    };
    // After the root level of the slot path: index rows for the survivors (slot count = children / 4).
    auto launchSlotIndex = [&](int nSlotsLaunch) -> int {
        const int slotStride = (int)(b->dNodes[H - 1].cap / 4);
        if (slotStride == 0 || nSlotsLaunch == 0) return LGS_OK;
        LGS_CUDA(c, b->dTab2.reserve((size_t)slotStride * std::max(b->maxUse, 1)));
        dim3 gi((nSlotsLaunch + 127) / 128, (std::max(b->maxUse, 1) + 7) / 8);
        bb_index_slots_kernel<<<gi, 128, 0, c->stream>>>(b->dQs.p, b->dUs.p, b->dHits.p, b->dSlotQT.p,
                                                        b->dCounters.p + 1 + H, slotStride, b->dFlags.p, nFlag,
                                                        b->dTab2.p);
        LGS_LAUNCH_CHECK(c);
        return LGS_OK;
    };
    auto finish = [&](int leafLaunch, const int* leafDev) -> int {   // winner, verification, replay
        if (leafLaunch > 0) {
            bb_leaf_rank_kernel<<<(leafLaunch + 127) / 128, 128, 0, c->stream>>>(b->dNodes[0].p, b->dScores[0].p, leafLaunch, leafDev, b->dBest.p);
            LGS_LAUNCH_CHECK(c);
            bb_leaf_pick_kernel<<<(leafLaunch + 127) / 128, 128, 0, c->stream>>>(b->dNodes[0].p, b->dScores[0].p, leafLaunch, leafDev, b->dBest.p);
            LGS_LAUNCH_CHECK(c);
        }
        LevelViews lv;
        for (int h = 0; h < kMaxLevels; ++h) lv.v[h] = LevelView{b->dNodes[h].p, b->dScores[h].p};
        bb_verify_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(b->dQs.p, n, H, lv, b->dBest.p, b->dRes.p,
                                                              b->forceReplay ? 1 : 0);
        LGS_LAUNCH_CHECK(c);
        bb_replay_kernel<<<(n + 31) / 32, 32, 0, c->stream>>>(b->dQs.p, n, H, lv, b->dBest.p, b->dRes.p);
        LGS_LAUNCH_CHECK(c);
        return LGS_OK;
    };

    if (spec) {
        const auto l0 = std::chrono::steady_clock::now();
        // Every buffer the level launches need, grown BEFORE anything is enqueued (the launch code's own
        // reserve() calls are then no-ops), so that the chain below can also be recorded into a CUDA graph.
        for (int h = H; h >= 0; --h) {
            const long long expect = h == H ? b->totalRoots : b->hint[h];
            if (h < H && b->dNodes[h].cap == 0) break;
            const int nMax = h == H ? b->totalRoots : (int)b->dNodes[h].cap;
            if (nMax == 0) break;
            if (h > 0) {
                const size_t want = (size_t)std::min<long long>(std::max<long long>(4LL * expect, 64), 1 << 16);
                if (b->dNodes[h - 1].cap < want) LGS_CUDA(c, b->dNodes[h - 1].reserve(want));
            }
            LGS_CUDA(c, b->dScores[h].reserve(nMax));
            if (slots && h > 0) LGS_CUDA(c, b->dSlot[h - 1].reserve(std::max<size_t>(b->dNodes[h - 1].cap, 1)));
        }
        if (slots && H > 0) {
            const size_t slotStride = b->dNodes[H - 1].cap / 4;
            LGS_CUDA(c, b->dSlotQT.reserve(std::max<size_t>(slotStride, 1)));
            if (slotStride > 0) LGS_CUDA(c, b->dTab2.reserve(slotStride * std::max(b->maxUse, 1)));
        }
        auto enqueue = [&]() -> int {
            for (int h = H; h >= 0; --h) {
                const long long expect = h == H ? b->totalRoots : b->hint[h];
                if (h < H && b->dNodes[h].cap == 0) break;
                const int nMax = h == H ? b->totalRoots : (int)b->dNodes[h].cap;
                const int* nDev = h == H ? nullptr : b->dCounters.p + 2 + h;      // children of level h + 1
                if (nMax == 0) break;
                int rc = launchLevel(h, nMax, nDev, expect);
                if (rc != LGS_OK) return rc;
                if (slots && h == H) {
                    // survivors expected from the last run (+25 %), never more than the pool can hold
                    b->slotLaunch = (int)std::min<long long>((long long)(b->dNodes[H - 1].cap / 4),
                                                             b->hint[H - 1] / 4 + b->hint[H - 1] / 16 + 256);
                    rc = launchSlotIndex(b->slotLaunch);
                    if (rc != LGS_OK) return rc;
                }
            }
            const int rc = finish(H == 0 ? b->totalRoots : (int)b->dNodes[0].cap, H == 0 ? nullptr : b->dCounters.p + 2);
            if (rc != LGS_OK) return rc;
            LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p, b->dCounters.p, (2 + kMaxLevels) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            return LGS_OK;
        };
        // LGS_BB_GRAPH=1 (experimental, off by default): the ~25 launches of the chain become one graph
        // launch -- they carry no host decision (grid sizes are pool capacities, counts stay on the
        // device).  Meant for many ranks driven from one host, where the launch path is contended
        // (profiles/r1_scaling.md).  Any capture / instantiate problem falls back to direct launches.
        bool launched = false;
        if (getenv("LGS_BB_GRAPH") != nullptr &&
            cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const long long launchesBefore = c->launches;
            const int rcCap = enqueue();
            cudaGraph_t graph = nullptr;
            const cudaError_t eEnd = cudaStreamEndCapture(c->stream, &graph);
            if (rcCap == LGS_OK && eEnd == cudaSuccess && graph != nullptr) {
                cudaGraphExecUpdateResultInfo info;
                if (b->graphExec != nullptr && cudaGraphExecUpdate(b->graphExec, graph, &info) != cudaSuccess) {
                    cudaGetLastError();
                    cudaGraphExecDestroy(b->graphExec);
                    b->graphExec = nullptr;
                }
                if (b->graphExec == nullptr && cudaGraphInstantiate(&b->graphExec, graph, 0) != cudaSuccess) {
                    cudaGetLastError();
                    b->graphExec = nullptr;
                }
                if (b->graphExec != nullptr && cudaGraphLaunch(b->graphExec, c->stream) == cudaSuccess) launched = true;
            }
            if (graph != nullptr) cudaGraphDestroy(graph);
            if (!launched) { cudaGetLastError(); c->launches = launchesBefore; }
        }
        if (!launched) {
            const int rc = enqueue();
            if (rc != LGS_OK) return rc;
        }
        b->pendingValidate = true;
        b->ran = true;
        b->hostMs[2] += hostMsSince(l0);
        b->hostRuns++;
        return LGS_OK;
    }
    int nNodes = b->totalRoots;
    for (int h = H; h >= 0; --h) {
        b->nodesPerLevel[h] = nNodes;
        if (nNodes == 0) break;
        int* nextCount = b->dCounters.p + 1 + h;
        if (h > 0 && b->dNodes[h - 1].cap < (size_t)std::min<long long>(4LL * nNodes, 1 << 16))
            LGS_CUDA(c, b->dNodes[h - 1].reserve(std::min<long long>(4LL * nNodes, 1 << 16)));
        for (int attempt = 0; attempt < 2; ++attempt) {
            LGS_CUDA(c, cudaMemsetAsync(nextCount, 0, sizeof(int), c->stream));
            const int rc = launchLevel(h, nNodes, nullptr, nNodes);
            if (rc != LGS_OK) return rc;
            if (h == 0) break;
            LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p + 1 + h, nextCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            LGS_CUDA(c, cudaStreamSynchronize(c->stream));
            const int want = b->hCounters.p[1 + h];
            if ((size_t)want <= b->dNodes[h - 1].cap) break;
            if (attempt == 1) return lgs_fail(c, LGS_ERR_OVERFLOW, "bb: level %d pool overflow", h - 1);
            LGS_CUDA(c, b->dNodes[h - 1].reserve((size_t)want + want / 4));   // grow, redo this level
        }
        nNodes = h > 0 ? b->hCounters.p[1 + h] : 0;
        if (slots && h == H && nNodes > 0) {
            const int rc = launchSlotIndex(nNodes / 4);
            if (rc != LGS_OK) return rc;
        }
    }
    for (int h = 0; h <= H; ++h) b->gathers += b->nodesPerLevel[h];   // refined per query below
    {
        const int rc = finish((int)b->nodesPerLevel[0], nullptr);
        if (rc != LGS_OK) return rc;
    }
    b->ran = true;
    b->pendingValidate = false;
    for (int h = 0; h < kMaxLevels; ++h) b->hint[h] = b->nodesPerLevel[h];
    b->haveHints = true;
    return LGS_OK;
}

// Wait for a speculative run and validate it; repeat level-synchronously if it cannot be trusted.
static int bb_finish(lgs_bb_batch* b) {
    lgs_ctx* c = b->ctx;
    if (!b->pendingValidate) return LGS_OK;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    b->pendingValidate = false;
    const int H = b->H;
    bool ok = !b->slotPath || H < 1 || b->hCounters.p[1 + H] / 4 <= b->slotLaunch;   // every survivor got its row
    for (int h = 1; h <= H && ok; ++h)
        ok = (size_t)b->hCounters.p[1 + h] <= b->dNodes[h - 1].cap;   // no pool overflow
    if (!ok) return bb_run_impl(b, false);
    for (int h = 0; h < kMaxLevels; ++h) b->nodesPerLevel[h] = 0;
    b->nodesPerLevel[H] = b->totalRoots;
    for (int h = 1; h <= H; ++h) b->nodesPerLevel[h - 1] = b->hCounters.p[1 + h];
    for (int h = 0; h < kMaxLevels; ++h) b->hint[h] = b->nodesPerLevel[h];
    return LGS_OK;
}

int lgs_bb_batch_run(lgs_bb_batch* b) {
    if (!b) return LGS_ERR_INVALID;
    // LGS_BB_SYNC=1 (diagnostic) forces the level-synchronous path
    const bool spec = b->haveHints && !b->forceReplay && getenv("LGS_BB_SYNC") == nullptr;
    return bb_run_impl(b, spec);
}


int lgs_bb_batch_results(lgs_bb_batch* b, lgs_match_result* out) {
    if (!b || (!out && b->nq > 0)) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->ran) return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_results before run");
    if (b->nq == 0) return LGS_OK;
    { const int rc = bb_finish(b); if (rc != LGS_OK) return rc; }
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaMemcpyAsync(b->hRes.p, b->dRes.p, b->nq * sizeof(BbResult), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    long long total = 0;
    for (int h = 0; h <= b->H; ++h) total += b->nodesPerLevel[h];
    for (int q = 0; q < b->nq; ++q) {
        const BbResult& r = b->hRes.p[q];
        const BbQuery& d = b->qs[q];
        lgs_match_result& o = out[q];
        o.found = r.found; o.ix = r.ix; o.iy = r.iy; o.it = r.it;
        o.win_x = d.winX; o.win_y = d.winY; o.win_t = d.winT;
        o.n_fixups = b->fixups[q];
        o.step_x = d.res; o.step_y = d.res; o.step_t = b->us[d.scan].stepT;
        o.score = r.score;
        o.n_scored = total;          // batch-wide count of nodes scored (all queries)
        o.exact_replay = r.exactReplay;
        o.reserved = 0;
    }
    return LGS_OK;
}

int lgs_bb_batch_work(const lgs_bb_batch* b, long long* nodesPerLevel, int nLevels, long long* gathers) {
    if (!b) return LGS_ERR_INVALID;
    { const int rc = bb_finish(const_cast<lgs_bb_batch*>(b)); if (rc != LGS_OK) return rc; }
    long long total = 0;
    for (int h = 0; h < kMaxLevels; ++h) {
        if (nodesPerLevel && h < nLevels) nodesPerLevel[h] = b->nodesPerLevel[h];
        total += b->nodesPerLevel[h];
    }
    if (gathers) {
        // every node sums the usable beams of its query; batches here share one beam count
        double avgUse = 0;
        for (const BbQuery& d : b->qs) avgUse += d.nUse;
        avgUse = b->qs.empty() ? 0 : avgUse / b->qs.size();
        *gathers = (long long)(total * avgUse);
    }
    return LGS_OK;
}

int lgs_bb_match(lgs_ctx* ctx, const lgs_bb_params* params, const lgs_scan_batch* scans,
                 lgs_pyramid* const* pyramids, const double* normThr, lgs_match_result* out) {
    lgs_bb_batch* b = nullptr;
    int rc = lgs_bb_batch_create(ctx, params, &b);
    if (rc != LGS_OK) return rc;
    rc = lgs_bb_batch_upload(b, scans, pyramids, normThr);
    if (rc == LGS_OK) rc = lgs_bb_batch_run(b);
    if (rc == LGS_OK) rc = lgs_bb_batch_results(b, out);
    lgs_bb_batch_destroy(b);
    return rc;
}

}  // extern "C"
