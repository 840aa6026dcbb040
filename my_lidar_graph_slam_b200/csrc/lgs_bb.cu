// lgs_bb.cu -- branch-and-bound scan matcher (loop detection) on sm_100a: C ABI, host preparation and
// the level-synchronous EXACT path.  The device-only run (one persistent kernel) is lgs_bb_run.cu.
//
// Replaces ScanMatcherBranchBound::OptimizePose(grid, pyramids, scan, pose, thr)
// (mapping/scan_matcher_branch_bound.cpp:47-163) with ScorePixelAccurate::Score
// (mapping/score_function_pixel_accurate.cpp:19-76) as node score, for a whole batch of
// (scan, submap) queries at once.
//
// How the CPU's depth-first search is reproduced exactly by a breadth-first one:
//  * Node score S(n) is a pure function of (x, y, theta, height).  The CPU visits a node only
//    if every ancestor scored above the running best, which never drops below the static
//    threshold thr * NumOfScans().  The level kernels therefore expand, level by level, the
//    SUPERSET of nodes whose ancestors all score above the static threshold, summing the gathered
//    cells of a node in beam order (bit-identical to the CPU sum).
//  * Let L* be the superset leaf with the highest score, ties broken by the CPU's LIFO visit
//    order (carried as an explicit rank: roots are popped x desc, y desc, theta desc; children
//    (x+w,y+w), (x,y+w), (x+w,y), (x,y); scan_matcher_branch_bound.cpp:85-88, :134-137).  If every
//    ancestor A of L* has S(A) >= S(L*), the CPU search provably returns L*: no earlier leaf
//    reaches S(L*), so no ancestor of L* is pruned, and no later leaf beats it.  The verify step
//    checks exactly that.
//  * Otherwise (the win-max maps are not upper bounds where a window index is negative,
//    SURVEY.md H12) the replay step repeats the CPU's stack discipline sequentially over the
//    stored superset scores; every node the CPU can touch is in the superset.
//
// World-coordinate re-projection (H4): the CPU recomputes
//     ix = floor(((sx + nx*step) + r*cos(theta_t + a_i) - minX) / res)
// per node and beam.  In exact arithmetic that is I0 + nx with I0 the index at nx = 0; all
// rounding errors together stay below 1e-11 cells, so I0 + nx is exact unless the fractional
// cell coordinate lies within the guard band of an edge.
//
// Two run paths share the node pools, the verify / replay logic and the result records:
//  * device-only (default, lgs_bb_run.cu): near-edge points are decided on the device by interval
//    evaluation of the CPU's expression; no host round trip inside a run;
//  * exact (this file; "bb_sync" / "bb_table" options, and the automatic fallback whenever a
//    device-only run reports an undecided near-edge point or an overflowed node pool): a full
//    per-query index table (bb_index_kernel) whose near-edge entries come from per-offset tables
//    computed on the HOST with the CPU's own expression and glibc sin / cos (H5); one host round
//    trip per level.
#include <cfloat>
#include <cmath>
#include <cstdint>

#include <chrono>

#include "lgs_bb.cuh"

using namespace lgsbb;

namespace {

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

// ---- node scoring + expansion, one thread per node (exact path) -----------------------------------
// ScorePixelAccurate::Score on pyramid level `height`, summed in beam order, through the full
// per-query index table.  Survivors of a warp allocate their children together (one atomic per warp)
// and store them child-major, so the next level's lanes again walk neighbouring thetas.
template <int U>                           // beams in flight per thread (two dependent latencies each)
__global__ void __launch_bounds__(128)
bb_score_kernel(const BbQuery* __restrict__ qs, const int2* __restrict__ tab,
                const int* __restrict__ exactIdx, int exactSpanX, int exactSpanY, int height,
                Node* __restrict__ nodes, double* __restrict__ scores, int nNodes,
                Node* __restrict__ next, int nextCap, int* __restrict__ nextCount, BbBest* __restrict__ best) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < nNodes;
    Node n;
    bool survive = false;
    if (active) {
        n = nodes[k];
        const BbQuery& d = qs[n.q];
        const double* __restrict__ lvl = d.level[height];
        const int stride = d.nTpad;
        const int2* __restrict__ tb = tab + d.tabBegin + n.t;
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
    }
}

// ---- node scoring, one WARP per node (exact path, small levels) ----------------------------------------
// The 32 lanes fetch the node's beams 512 at a time, park the values in shared memory in beam order,
// and lane 0 adds them strictly in that order: still bit-identical to the CPU sum.
__global__ void __launch_bounds__(128)
bb_score_warp_kernel(const BbQuery* __restrict__ qs, const int2* __restrict__ tab,
                     const int* __restrict__ exactIdx, int exactSpanX, int exactSpanY, int height,
                     Node* __restrict__ nodes, double* __restrict__ scores, int nNodes,
                     Node* __restrict__ next, int nextCap, int* __restrict__ nextCount, BbBest* __restrict__ best) {
    constexpr int CH = 16, STAGE = 32 * CH;
    __shared__ double sv[4][STAGE];
    const int wib = threadIdx.x >> 5;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= nNodes) return;                       // whole warp
    const Node n = nodes[k];
    const BbQuery& d = qs[n.q];
    const double* __restrict__ lvl = d.level[height];
    const int stride = d.nTpad;
    const int2* __restrict__ tb = tab + d.tabBegin + n.t;
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
    }
}

// ---- winner among the leaves: (score desc, rank asc) ---------------------------------------------
__global__ void bb_leaf_rank_kernel(const Node* __restrict__ leaves, const double* __restrict__ scores,
                                    int n, BbBest* __restrict__ best) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = leaves[k].q;
    if ((unsigned long long)__double_as_longlong(scores[k]) == best[q].scoreBits && best[q].scoreBits != 0ull)
        atomicMin((unsigned long long*)&best[q].rank, (unsigned long long)leaves[k].rank);
}

__global__ void bb_leaf_pick_kernel(const Node* __restrict__ leaves, const double* __restrict__ scores,
                                    int n, BbBest* __restrict__ best) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = leaves[k].q;
    if ((unsigned long long)__double_as_longlong(scores[k]) == best[q].scoreBits &&
        best[q].scoreBits != 0ull && leaves[k].rank == best[q].rank)
        best[q].leaf = k;
}

// ---- verification of the winner's ancestor chain + result record -----------------------------------
__global__ void bb_verify_kernel(const BbQuery* __restrict__ qs, int nq, int heightMax, LevelViews lv,
                                 BbBest* __restrict__ best, BbResult* __restrict__ res, int forceReplay) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    BbResult r;
    r.fixups = 0; r.exactReplay = 0;
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
    r.fixups = 0; r.exactReplay = 1;
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
    BbBest b;
    b.scoreBits = 0ull; b.rank = 0x7fffffffffffffffLL; b.leaf = -1; b.needReplay = 0;
    b.rankLeaf = ~0ull; b.fixups = 0; b.pad = 0;
    best[q] = b;
}

}  // namespace

namespace {

double ms_since(std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
}

size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

// The level-synchronous exact path: full per-query index table, near-edge points from the host (the
// CPU's expression with glibc sin / cos), one host round trip per level.  Leaves the results in
// hRes once the context stream has drained.
int bb_run_exact(lgs_bb_batch* b) {
    lgs_ctx* c = b->ctx;
    const int H = b->H, n = b->nq;
    for (int h = 0; h < kMaxLevels; ++h) b->nodesPerLevel[h] = 0;
    b->gathers = 0;
    b->pendingValidate = false;
    b->lastRunDevice = false;
    b->needDeliver = true;
    b->exactRuns++;
    char* blob = b->dBlob.p;
    const BbQuery* dQs = reinterpret_cast<const BbQuery*>(blob + b->offQs);
    const BbScan* dUs = reinterpret_cast<const BbScan*>(blob + b->offUs);
    const int* dQlist = reinterpret_cast<const int*>(blob + b->offQlist);
    const IdxChunk* dChunks = reinterpret_cast<const IdxChunk*>(blob + b->offChunks);
    const double* dAngles = reinterpret_cast<const double*>(blob + b->offAngles);
    const double* dRanges = reinterpret_cast<const double*>(blob + b->offRanges);
    LGS_CUDA(c, b->dCounters.reserve(kCounters));
    LGS_CUDA(c, b->dFlags.reserve(kFlagCapBB));
    LGS_CUDA(c, b->dHits.reserve(std::max<long long>(b->nHits, 1)));
    LGS_CUDA(c, b->dTab.reserve(std::max<long long>(b->nTab, 1)));
    LGS_CUDA(c, b->dNodes[H].reserve(b->totalRoots));
    LGS_CUDA(c, b->dScores[H].reserve(b->totalRoots));
    LGS_CUDA(c, cudaMemsetAsync(b->dCounters.p, 0, kCounters * sizeof(int), c->stream));
    const unsigned gx = (unsigned)((b->maxNTpad + 127) / 128), gy = (unsigned)std::max(b->maxUse, 1);
    for (size_t u0 = 0; u0 < b->us.size(); u0 += 65535) {
        const unsigned nu = (unsigned)std::min<size_t>(65535, b->us.size() - u0);
        bb_hit_kernel<<<dim3(gx, gy, nu), 128, 0, c->stream>>>(dUs + u0, dAngles, dRanges, b->dHits.p);
        LGS_LAUNCH_CHECK(c);
    }
    {
        dim3 gridR((b->maxRoots + 127) / 128, n);
        bb_roots_kernel<<<gridR, 128, 0, c->stream>>>(dQs, n, H, b->dNodes[H].p);
        LGS_LAUNCH_CHECK(c);
        bb_init_best_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(b->dBest.p, n);
        LGS_LAUNCH_CHECK(c);
    }
    for (size_t c0 = 0; c0 < b->chunks.size(); c0 += 65535) {
        const unsigned nc = (unsigned)std::min<size_t>(65535, b->chunks.size() - c0);
        bb_index_kernel<<<dim3(gx, gy, nc), 128, 0, c->stream>>>(dQs, dUs, dChunks + c0, dQlist, b->dHits.p,
                                                                 c->opt.edgeEps, b->dTab.p, b->dFlags.p,
                                                                 b->dCounters.p + kCtrFlags);
        LGS_LAUNCH_CHECK(c);
    }
    LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p, b->dCounters.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    const int nFlag = b->hCounters.p[kCtrFlags];
    const int spanX = b->spanX, spanY = b->spanY;
    std::fill(b->fixups.begin(), b->fixups.end(), 0);
    if (nFlag > kFlagCapBB)
        return lgs_fail(c, LGS_ERR_OVERFLOW, "bb: %d near-edge points exceed the fix-up list", nFlag);
    if (nFlag > 0) {
        // Near-edge beams: exact per-offset index tables from the host (CPU expression + glibc).
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
    const int warpBelow = c->opt.bbWarpBelow;
    int nNodes = b->totalRoots;
    for (int h = H; h >= 0; --h) {
        b->nodesPerLevel[h] = nNodes;
        if (nNodes == 0) break;
        int* nextCount = b->dCounters.p + kCtrChild + h;
        if (h > 0 && b->dNodes[h - 1].cap < (size_t)std::min<long long>(4LL * nNodes, 1 << 16))
            LGS_CUDA(c, b->dNodes[h - 1].reserve(std::min<long long>(4LL * nNodes, 1 << 16)));
        LGS_CUDA(c, b->dScores[h].reserve(std::max<size_t>(nNodes, b->dNodes[h].cap)));
        for (int attempt = 0; attempt < 2; ++attempt) {
            LGS_CUDA(c, cudaMemsetAsync(nextCount, 0, sizeof(int), c->stream));
            Node* next = h > 0 ? b->dNodes[h - 1].p : nullptr;
            const int nextCap = h > 0 ? (int)std::min<size_t>(b->dNodes[h - 1].cap, 0x7fffffff) : 0;
            if (nNodes >= warpBelow)
                bb_score_kernel<32><<<(nNodes + 127) / 128, 128, 0, c->stream>>>(
                    dQs, b->dTab.p, b->dExact.p, spanX, spanY, h, b->dNodes[h].p, b->dScores[h].p, nNodes, next,
                    nextCap, nextCount, b->dBest.p);
            else
                bb_score_warp_kernel<<<(nNodes + 3) / 4, 128, 0, c->stream>>>(
                    dQs, b->dTab.p, b->dExact.p, spanX, spanY, h, b->dNodes[h].p, b->dScores[h].p, nNodes, next,
                    nextCap, nextCount, b->dBest.p);
            LGS_LAUNCH_CHECK(c);
            if (h == 0) break;
            LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p + kCtrChild + h, nextCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            LGS_CUDA(c, cudaStreamSynchronize(c->stream));
            const int want = b->hCounters.p[kCtrChild + h];
            if ((size_t)want <= b->dNodes[h - 1].cap) break;
            if (attempt == 1) return lgs_fail(c, LGS_ERR_OVERFLOW, "bb: level %d pool overflow", h - 1);
            LGS_CUDA(c, b->dNodes[h - 1].reserve((size_t)want + want / 4));   // grow, redo this level
        }
        nNodes = h > 0 ? b->hCounters.p[kCtrChild + h] : 0;
    }
    const int nLeaves = (int)b->nodesPerLevel[0];
    if (nLeaves > 0) {
        bb_leaf_rank_kernel<<<(nLeaves + 127) / 128, 128, 0, c->stream>>>(b->dNodes[0].p, b->dScores[0].p, nLeaves, b->dBest.p);
        LGS_LAUNCH_CHECK(c);
        bb_leaf_pick_kernel<<<(nLeaves + 127) / 128, 128, 0, c->stream>>>(b->dNodes[0].p, b->dScores[0].p, nLeaves, b->dBest.p);
        LGS_LAUNCH_CHECK(c);
    }
    LevelViews lv;
    for (int h = 0; h < kMaxLevels; ++h) lv.v[h] = LevelView{b->dNodes[h].p, b->dScores[h].p};
    bb_verify_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(dQs, n, H, lv, b->dBest.p, b->dRes.p, b->forceReplay ? 1 : 0);
    LGS_LAUNCH_CHECK(c);
    bb_replay_kernel<<<(n + 31) / 32, 32, 0, c->stream>>>(dQs, n, H, lv, b->dBest.p, b->dRes.p);
    LGS_LAUNCH_CHECK(c);
    LGS_CUDA(c, cudaMemcpyAsync(b->hRes.p, b->dRes.p, n * sizeof(BbResult), cudaMemcpyDeviceToHost, c->stream));
    for (int h = 0; h <= H; ++h) b->hint[h] = std::max(b->hint[h], b->nodesPerLevel[h]);
    return LGS_OK;
}

void bb_fill_records(const lgs_bb_batch* b, lgs_loop_record* out) {
    for (int q = 0; q < b->nq; ++q) {
        const BbResult& r = b->hRes.p[q];
        out[q].found = r.found; out[q].ix = r.ix; out[q].iy = r.iy; out[q].it = r.it;
        out[q].score = r.score;
        out[q].id = b->ids.empty() ? (long long)q : b->ids[q];
    }
}

// The exact path leaves its results on the host; the 32-byte records go to the same place the
// device-only run's finalize phase writes them (the sink, or the batch's own record buffer).
int bb_deliver_records(lgs_bb_batch* b) {
    lgs_ctx* c = b->ctx;
    LGS_CUDA(c, b->dRec.reserve(b->nq + 1));
    LGS_CUDA(c, b->hRec.reserve(b->nq + 1));
    bb_fill_records(b, b->hRec.p);
    lgs_loop_record& st = b->hRec.p[b->nq];          // status record: valid
    st.found = 1; st.ix = 0; st.iy = 0; st.it = 0; st.score = 0.0; st.id = -1;
    lgs_loop_record* dst = b->sink ? b->sink + b->sinkFirst : b->dRec.p;
    LGS_CUDA(c, cudaMemcpyAsync(dst, b->hRec.p, (b->nq + 1) * sizeof(lgs_loop_record), cudaMemcpyDefault, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    return LGS_OK;
}

// Wait for the last run; a device-only run that could not decide a near-edge point or overflowed a
// node pool is repeated on the exact path.
int bb_settle(lgs_bb_batch* b) {
    lgs_ctx* c = b->ctx;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    if (b->pendingValidate) {
        b->pendingValidate = false;
        const int H = b->H;
        const int* hc = b->hCounters.p;
        const bool ok = hc[kCtrDone] == 1 && hc[kCtrOverflow] == 0 && hc[kCtrUnresolved] == 0;
        for (int h = 0; h < kMaxLevels; ++h) b->nodesPerLevel[h] = 0;
        b->nodesPerLevel[H] = b->totalRoots;
        for (int h = 0; h < H; ++h) b->nodesPerLevel[h] = hc[kCtrChild + h + 1];
        b->skippedBeams = 16LL * hc[kCtrSkipped];
        for (int h = 0; h <= H; ++h) b->hint[h] = std::max(b->hint[h], b->nodesPerLevel[h]);
        if (ok) {
            for (int q = 0; q < b->nq; ++q) b->fixups[q] = b->hRes.p[q].fixups;
        } else {
            const int rc = bb_run_exact(b);
            if (rc != LGS_OK) return rc;
            LGS_CUDA(c, cudaStreamSynchronize(c->stream));
        }
    }
    if (b->needDeliver) {
        b->needDeliver = false;
        return bb_deliver_records(b);
    }
    return LGS_OK;
}

}  // namespace

extern "C" {

int lgs_bb_batch_create(lgs_ctx* ctx, const lgs_bb_params* p, lgs_bb_batch** out) {
    if (!ctx || !p || !out) return LGS_ERR_INVALID;
    *out = nullptr;
    if (p->node_height_max < 0 || p->node_height_max >= kMaxLevels - 1 || !(p->range_x >= 0) ||
        !(p->range_y >= 0) || !(p->range_theta >= 0))
        return lgs_fail(ctx, LGS_ERR_INVALID, "bb_batch_create: bad parameters");
    LGS_CUDA(ctx, cudaSetDevice(ctx->device));
    lgs_bb_batch* b = new lgs_bb_batch();
    b->ctx = ctx; b->params = *p; b->H = p->node_height_max;
    if (cudaEventCreateWithFlags(&b->evUpload, cudaEventDisableTiming) != cudaSuccess) {
        delete b;
        return lgs_fail(ctx, LGS_ERR_CUDA, "bb_batch_create: cudaEventCreate");
    }
    *out = b;
    return LGS_OK;
}

int lgs_bb_batch_destroy(lgs_bb_batch* b) {
    if (!b) return LGS_OK;
    if (b->hostRuns > 0 && b->ctx->opt.bbHostTiming)
        fprintf(stderr, "[lgs bb host] %lld runs of %d queries (%lld device-only, %lld exact): per run upload %.3f ms, "
                "run enqueue %.3f ms, results wait %.3f ms\n", b->hostRuns, b->nq, b->deviceRuns, b->exactRuns,
                b->hostMs[0] / b->hostRuns, b->hostMs[1] / b->hostRuns, b->hostMs[2] / b->hostRuns);
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    if (b->evUpload) cudaEventDestroy(b->evUpload);
    b->hBlob.release(); b->dBlob.release(); b->dHitsFix.release(); b->dHitsFixT.release(); b->dHits.release(); b->dTab.release();
    b->dChunks.release(); b->dFlags.release(); b->dCounters.release(); b->dCtr.release(); b->dExact.release();
    b->dBest.release(); b->dRes.release(); b->dRec.release(); b->hRes.release(); b->hCounters.release();
    b->hRec.release(); b->dPhase.release(); b->hPhase.release();
    for (int h = 0; h < kMaxLevels; ++h) { b->dNodes[h].release(); b->dScores[h].release(); }
    delete b;
    return LGS_OK;
}

int lgs_bb_batch_force_replay(lgs_bb_batch* b, int on) {
    if (!b) return LGS_ERR_INVALID;
    b->forceReplay = on != 0;
    return LGS_OK;
}

int lgs_bb_batch_set_record_ids(lgs_bb_batch* b, const long long* ids, int n) {
    if (!b || n < 0 || (n > 0 && !ids)) return LGS_ERR_INVALID;
    b->ids.assign(ids, ids + n);
    b->uploaded = false;            // shipped with the next upload
    return LGS_OK;
}

int lgs_bb_batch_set_record_sink(lgs_bb_batch* b, lgs_loop_record* deviceRecords, long long firstSlot) {
    if (!b || firstSlot < 0) return LGS_ERR_INVALID;
    b->sink = deviceRecords;
    b->sinkFirst = deviceRecords ? firstSlot : 0;
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
    const auto t0 = std::chrono::steady_clock::now();
    const int n = nPairs;
    if (n < 0 || scans->n_scans < 0 || (n > 0 && (!scans->beam_begin || !scans->sensor_pose || !pyramids)))
        return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_upload: bad arguments");
    for (int q = 0; q < n; ++q) {
        const int sq = pairScan ? pairScan[q] : q;
        if (sq < 0 || sq >= scans->n_scans)
            return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_upload: pair %d names scan %d of %d", q, sq, scans->n_scans);
    }
    if (!b->ids.empty() && (int)b->ids.size() != n)
        return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_upload: %zu record ids for %d pairs", b->ids.size(), n);
    LGS_CUDA(c, cudaSetDevice(c->device));
    const lgs_bb_params& p = b->params;
    const int H = b->H;
    b->uploaded = false; b->ran = false; b->pendingValidate = false;
    b->nq = n;
    b->qs.resize(n);                  // every field is written below (template copy + scan-dependent part)
    b->us.clear();
    b->fixups.assign(n, 0);
    b->hAngles.clear(); b->hRanges.clear();
    b->maxRoots = 0; b->maxNTpad = 0; b->maxUse = 0; b->spanX = 0; b->spanY = 0;
    b->maxAbsCells = 0.0; b->maxReachCells = 0.0;
    long long nTab = 0, roots = 0, nHits = 0, nHitsT = 0;
    int projTiles = 0;
    const int winSizeMax = 1 << H;
    // Pairs that name the same scan share its projected hit points (1 scan x many submaps).
    std::vector<int> scanToUnique(std::max(scans->n_scans, 1), -1);
    std::vector<double> scanMaxR(std::max(scans->n_scans, 1), 0.0);
    std::vector<char> haveMaxR(std::max(scans->n_scans, 1), 0);
    // Host preparation is on the end-to-end path of every step, so everything that depends only on the
    // submap (geometry, level pointers) or only on the (scan, resolution) (search step, window, usable
    // beams) is derived once and copied: a 64-scan x 63-submap share is 4032 pairs but 63 + 64 of those.
    struct PyrInfo { const lgs_pyramid* pyr; BbQuery tmpl; };
    std::vector<PyrInfo> pyrCache;
    pyrCache.reserve(64);
    std::vector<int> pyrSlot(1024, -1);                // open-addressed pointer hash -> pyrCache index
    auto pyrLookup = [&](const lgs_pyramid* pyr) -> int {
        size_t mask = pyrSlot.size() - 1;
        size_t h = ((size_t)reinterpret_cast<uintptr_t>(pyr) >> 4) * 0x9E3779B97F4A7C15ull;
        for (size_t k = h & mask;; k = (k + 1) & mask) {
            if (pyrSlot[k] < 0) return -(int)k - 1;
            if (pyrCache[pyrSlot[k]].pyr == pyr) return pyrSlot[k];
        }
    };
    for (int q = 0; q < n; ++q) {
        const int sq = pairScan ? pairScan[q] : q;
        const int b0 = scans->beam_begin[sq], b1 = scans->beam_begin[sq + 1];
        const int nb = b1 - b0;
        if (nb <= 0) return lgs_fail(c, LGS_ERR_INVALID, "bb: scan %d has no beams", q);
        const lgs_pyramid* pyr = pyramids[q];
        int slot = pyr ? pyrLookup(pyr) : 0;
        if (!pyr || slot < 0) {
            if (!pyr || lgs_pyramid_levels(pyr) < H + 1)
                return lgs_fail(c, LGS_ERR_INVALID, "bb: query %d needs a pyramid with %d levels", q, H + 1);
            const lgs_grid* g0 = lgs_pyramid_level(pyr, 0);
            lgs_pyramid_note_user(pyr, c);
            if (g0->ctx->device != c->device)
                return lgs_fail(c, LGS_ERR_INVALID, "bb: pyramid of query %d lives on another device", q);
            PyrInfo info;
            info.pyr = pyr;
            BbQuery& t = info.tmpl;
            t = BbQuery{};
            t.minX = g0->min_x; t.minY = g0->min_y; t.res = g0->res; t.invRes = 1.0 / g0->res;
            t.nx = g0->nx; t.ny = g0->ny; t.pitch = g0->pitch;
            t.offX = g0->off_x; t.offY = g0->off_y;
            for (int h = 0; h <= H; ++h) t.level[h] = lgs_pyramid_level(pyr, h)->origin();
            if (2 * (pyrCache.size() + 1) > pyrSlot.size()) {          // grow + rehash
                std::vector<int> old;
                old.swap(pyrSlot);
                pyrSlot.assign(old.size() * 4, -1);
                for (int idx : old) if (idx >= 0) pyrSlot[(size_t)(-pyrLookup(pyrCache[idx].pyr) - 1)] = idx;
                slot = pyrLookup(pyr);
            }
            pyrSlot[(size_t)(-slot - 1)] = (int)pyrCache.size();
            slot = (int)pyrCache.size();
            pyrCache.push_back(info);
        }
        BbQuery& d = b->qs[q];
        d = pyrCache[slot].tmpl;
        int uidx = scanToUnique[sq];
        if (uidx >= 0 && b->us[uidx].invRes != d.invRes) uidx = -1;        // another map resolution
        if (uidx < 0) {
            // ComputeSearchStep (scan_matcher_branch_bound.cpp:178-197)
            if (!haveMaxR[sq]) {           // std::max_element over the scan, once per scan
                double m = scans->ranges[b0];
                for (int i = b0 + 1; i < b1; ++i) m = std::max(m, scans->ranges[i]);
                scanMaxR[sq] = m;
                haveMaxR[sq] = 1;
            }
            const double maxRange = std::min(scanMaxR[sq], p.scan_range_max);
            const double th = d.res / maxRange;
            const double stepX = d.res, stepY = d.res;
            const double stepT = std::acos(1.0 - 0.5 * th * th);
            BbScan u{};
            u.winX = static_cast<int>(std::ceil(0.5 * p.range_x / stepX));                 // :68-73
            u.winY = static_cast<int>(std::ceil(0.5 * p.range_y / stepY));
            u.winT = static_cast<int>(std::ceil(0.5 * p.range_theta / stepT));
            if (!(stepT > 0.0) || u.winT < 0 || u.winT > (1 << 20))
                return lgs_fail(c, LGS_ERR_INVALID, "bb: scan %d gives stepTheta=%g winTheta=%d", q, stepT, u.winT);
            if (u.winX + 2 * winSizeMax > 32000 || u.winY + 2 * winSizeMax > 32000)
                return lgs_fail(c, LGS_ERR_INVALID, "bb: search window too large for 16-bit node offsets");
            u.nT = 2 * u.winT + 1;
            u.nTpad = (u.nT + 3) / 4 * 4;
            u.nrx = (2 * u.winX) / winSizeMax + 1;                                         // :85-86
            u.nry = (2 * u.winY) / winSizeMax + 1;
            // ScorePixelAccurate range filter (score_function_pixel_accurate.cpp:27-41)
            const double sMin = scans->range_min ? scans->range_min[sq] : 0.0;
            const double sMax = scans->range_max ? scans->range_max[sq] : HUGE_VAL;
            const double minRange = std::max(p.score_range_min, sMin);
            const double maxRangeS = std::min(p.score_range_max, sMax);
            u.sx = scans->sensor_pose[3 * sq]; u.sy = scans->sensor_pose[3 * sq + 1];
            u.st = scans->sensor_pose[3 * sq + 2];
            u.stepT = stepT;
            u.invRes = d.invRes;
            u.originX = std::floor(u.sx * d.invRes); u.originY = std::floor(u.sy * d.invRes);
            u.qBegin = 0; u.qCount = 0;
            u.beamBegin = (int)b->hAngles.size();
            double reach = 0.0;
            for (int i = b0; i < b1; ++i) {
                const double r = scans->ranges[i];
                if (r >= maxRangeS || r <= minRange) continue;
                b->hAngles.push_back(scans->angles[i]);
                b->hRanges.push_back(r);
                reach = std::max(reach, std::fabs(r));
            }
            u.nUse = (int)b->hAngles.size() - u.beamBegin;
            u.hitBegin = nHits;
            nHits += (long long)u.nUse * u.nTpad;
            u.beamPad = (u.nUse + 3) / 4 * 4;
            u.projBegin = projTiles;
            projTiles += ((u.nUse + 7) / 8) * ((u.nT + 31) / 32);
            u.hitTBegin = nHitsT;
            nHitsT += (long long)u.nT * u.beamPad;
            b->maxAbsCells = std::max(b->maxAbsCells, (std::max(std::fabs(u.sx), std::fabs(u.sy)) + reach) * d.invRes);
            b->maxReachCells = std::max(b->maxReachCells, reach * d.invRes);
            uidx = (int)b->us.size();
            b->us.push_back(u);
            // a scan matched against maps of several resolutions keeps one entry per resolution; the
            // lookup remembers the last one (maps of one resolution are the rule)
            scanToUnique[sq] = uidx;
        }
        const BbScan& us = b->us[uidx];
        d.winX = us.winX; d.winY = us.winY; d.winT = us.winT; d.nT = us.nT; d.nTpad = us.nTpad;
        d.nrx = us.nrx; d.nry = us.nry;
        const double thr = normThr ? normThr[q] : DBL_MIN;
        d.thrAbs = thr * static_cast<double>(static_cast<size_t>(nb));                 // :75-76
        d.scan = uidx;
        d.nUse = us.nUse;
        d.tabBegin = nTab;
        nTab += (long long)d.nUse * d.nTpad;
        {   // 12.20 fixed-point origin of the query's map relative to the scan origin (+ the window offset of a
            // banded map), split into fraction bits and whole cells (lgs_bb_run.cu)
            const long long mx = std::llrint((d.minX * d.invRes - us.originX) * 1048576.0) + ((long long)d.offX << 20);
            const long long my = std::llrint((d.minY * d.invRes - us.originY) * 1048576.0) + ((long long)d.offY << 20);
            d.MloX = (int)(mx & 0xfffff); d.MhiX = (int)(mx >> 20);
            d.MloY = (int)(my & 0xfffff); d.MhiY = (int)(my >> 20);
        }
        b->maxAbsCells = std::max(b->maxAbsCells, std::max(std::fabs(d.minX), std::fabs(d.minY)) * d.invRes +
                                                  std::max(d.offX, d.offY) + std::max(d.nx, d.ny));
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
    b->nHits = nHits;
    b->nHitsT = nHitsT;
    b->projTiles = projTiles;
    b->totalRoots = (int)roots;
    const size_t nu = b->us.size();
    {   // group the queries by distinct scan; kIdxChunk per index-kernel block (exact path)
        std::vector<int> count(nu, 0);
        for (int q = 0; q < n; ++q) count[b->qs[q].scan]++;
        int run = 0;
        for (size_t u = 0; u < nu; ++u) { b->us[u].qBegin = run; b->us[u].qCount = 0; run += count[u]; }
        b->qlist.assign(n, 0);
        for (int q = 0; q < n; ++q) { BbScan& u = b->us[b->qs[q].scan]; b->qlist[u.qBegin + u.qCount++] = q; }
        b->chunks.clear();
        for (size_t u = 0; u < nu; ++u)
            for (int k = 0; k < b->us[u].qCount; k += kIdxChunk)
                b->chunks.push_back(IdxChunk{(int)u, b->us[u].qBegin + k, std::min(kIdxChunk, b->us[u].qCount - k)});
        // root-level warp tiles per scan for the four warp mappings (lgs_bb_run.cu)
        const int gs[4] = {1, 4, 8, 32};
        bool tilesOk = true;
        for (int i = 0; i < 4; ++i) {
            const int npw = 32 / gs[i];
            b->tileBegin[i].assign(nu + 1, 0);
            long long acc = 0;
            for (size_t u = 0; u < nu; ++u) {
                b->tileBegin[i][u] = (int)acc;
                acc += (long long)((b->us[u].nT + npw - 1) / npw) * b->us[u].nrx * b->us[u].nry * b->us[u].qCount;
                if (acc > 0x7fffffffLL) { tilesOk = false; acc = 0x7fffffffLL; }
            }
            b->tileBegin[i][nu] = (int)acc;
        }
        // device-only run: coordinates inside the fixed-point range, rank << 24 | leaf fits 64 bits
        // device-only run: beams inside the 12.20 fixed-point range around the scan origin, whole-cell parts
        // inside int32, rank << 24 | leaf inside 64 bits, 32-bit hit-array indices
        b->deviceOk = tilesOk && b->maxReachCells < 2000.0 && b->maxAbsCells < 5.0e8 && (H > 0 || roots <= (1 << 24)) &&
                      (double)b->maxRoots * std::pow(4.0, H) < 1099511627776.0 && nHits < 0x7fffffffLL &&
                      nHitsT < 0x7fffffffLL;
    }
    if (n == 0) { b->uploaded = true; return LGS_OK; }
    // One staging blob, one H2D copy.
    const size_t nk = b->hAngles.size();
    size_t off = 0;
    b->offQs = off;      off = align16(off + (size_t)n * sizeof(BbQuery));
    b->offUs = off;      off = align16(off + nu * sizeof(BbScan));
    b->offQlist = off;   off = align16(off + (size_t)n * sizeof(int));
    b->offTiles = off;   off = align16(off + 4 * (nu + 1) * sizeof(int));
    b->offChunks = off;  off = align16(off + b->chunks.size() * sizeof(IdxChunk));
    b->offAngles = off;  off = align16(off + nk * sizeof(double));
    b->offRanges = off;  off = align16(off + nk * sizeof(double));
    b->offIds = off;     off = align16(off + b->ids.size() * sizeof(long long));
    LGS_CUDA(c, cudaEventSynchronize(b->evUpload));        // the previous upload's copy has left the pinned blob
    LGS_CUDA(c, b->hBlob.reserve(off));
    if (b->dBlob.cap < off) {
        LGS_CUDA(c, cudaStreamSynchronize(c->stream));      // a run may still read the old blob
        LGS_CUDA(c, b->dBlob.reserve(off));
    }
    char* hb = b->hBlob.p;
    memcpy(hb + b->offQs, b->qs.data(), (size_t)n * sizeof(BbQuery));
    memcpy(hb + b->offUs, b->us.data(), nu * sizeof(BbScan));
    memcpy(hb + b->offQlist, b->qlist.data(), (size_t)n * sizeof(int));
    for (int i = 0; i < 4; ++i)
        memcpy(hb + b->offTiles + (size_t)i * (nu + 1) * sizeof(int), b->tileBegin[i].data(), (nu + 1) * sizeof(int));
    if (!b->chunks.empty()) memcpy(hb + b->offChunks, b->chunks.data(), b->chunks.size() * sizeof(IdxChunk));
    if (nk) {
        memcpy(hb + b->offAngles, b->hAngles.data(), nk * sizeof(double));
        memcpy(hb + b->offRanges, b->hRanges.data(), nk * sizeof(double));
    }
    if (!b->ids.empty()) memcpy(hb + b->offIds, b->ids.data(), b->ids.size() * sizeof(long long));
    LGS_CUDA(c, b->hCounters.reserve(kCounters));
    LGS_CUDA(c, b->dBest.reserve(n));
    LGS_CUDA(c, b->dRes.reserve(n));
    LGS_CUDA(c, b->hRes.reserve(n));
    LGS_CUDA(c, cudaMemcpyAsync(b->dBlob.p, hb, off, cudaMemcpyHostToDevice, c->stream));
    LGS_CUDA(c, cudaEventRecord(b->evUpload, c->stream));
    b->uploaded = true;
    b->hostMs[0] += ms_since(t0);
    return LGS_OK;
}

int lgs_bb_batch_run(lgs_bb_batch* b) {
    if (!b) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->uploaded) return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_run before upload");
    const auto t0 = std::chrono::steady_clock::now();
    for (int h = 0; h < kMaxLevels; ++h) b->nodesPerLevel[h] = 0;
    if (b->nq == 0) { b->ran = true; return LGS_OK; }
    LGS_CUDA(c, cudaSetDevice(c->device));
    // "bb_sync" / "bb_table" (diagnostic and test hooks) force the level-synchronous exact path
    const bool device = b->deviceOk && !c->opt.bbSync && !c->opt.bbTable;
    const int rc = device ? lgs_bb_launch_device_run(b) : bb_run_exact(b);
    if (rc != LGS_OK) return rc;
    b->ran = true;
    b->hostMs[1] += ms_since(t0);
    b->hostRuns++;
    return LGS_OK;
}

int lgs_bb_batch_results(lgs_bb_batch* b, lgs_match_result* out) {
    if (!b || (!out && b->nq > 0)) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->ran) return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_results before run");
    if (b->nq == 0) return LGS_OK;
    const auto t0 = std::chrono::steady_clock::now();
    { const int rc = bb_settle(b); if (rc != LGS_OK) return rc; }
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
        // nodes scored: of this query when the run counted them ("bb_count_nodes"), else of the whole batch
        const bool perQuery = b->lastRunDevice && (c->opt.bbCountNodes || c->opt.bbHostTiming);
        o.n_scored = perQuery ? (long long)d.nrx * d.nry * d.nT + r.nodes : total;
        o.exact_replay = r.exactReplay;
        o.reserved = b->lastRunDevice ? 0 : 1;     // 1: the level-synchronous exact path produced this result
    }
    b->hostMs[2] += ms_since(t0);
    return LGS_OK;
}

int lgs_bb_batch_records(lgs_bb_batch* b, lgs_loop_record* out) {
    if (!b || (!out && b->nq > 0)) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->ran) return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_records before run");
    if (b->nq == 0) return LGS_OK;
    const auto t0 = std::chrono::steady_clock::now();
    { const int rc = bb_settle(b); if (rc != LGS_OK) return rc; }
    bb_fill_records(b, out);
    b->hostMs[2] += ms_since(t0);
    return LGS_OK;
}

int lgs_bb_batch_settle(lgs_bb_batch* b) {
    if (!b) return LGS_ERR_INVALID;
    if (!b->ran) return lgs_fail(b->ctx, LGS_ERR_INVALID, "bb_batch_settle before run");
    if (b->nq == 0) return LGS_OK;
    return bb_settle(b);
}

void* lgs_bb_batch_device_records(lgs_bb_batch* b) { return b ? (void*)b->dRec.p : nullptr; }

int lgs_bb_batch_work(const lgs_bb_batch* b, long long* nodesPerLevel, int nLevels, long long* gathers) {
    if (!b) return LGS_ERR_INVALID;
    if (b->ran && b->nq > 0) { const int rc = bb_settle(const_cast<lgs_bb_batch*>(b)); if (rc != LGS_OK) return rc; }
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

// Beams the last device-only run did not have to gather because the node was rejected early (a lower
// bound, counted in rounds of 16): algorithmic gathers (lgs_bb_batch_work) minus these were issued.
long long lgs_bb_batch_skipped_gathers(const lgs_bb_batch* b) {
    if (!b) return 0;
    if (b->ran && b->nq > 0 && bb_settle(const_cast<lgs_bb_batch*>(b)) != LGS_OK) return 0;
    return b->lastRunDevice ? b->skippedBeams : 0;
}

// Diagnostic ("bb_host_timing" must be on before the run): microseconds between the phase boundaries
// of the last device-only run -- hit points, root level, levels H-1..0, winner, finalize -- and the
// warp mapping (lanes per node) each level used.
int lgs_bb_batch_phase_times(lgs_bb_batch* b, double* us, int* mapping, int n) {
    if (!b || !us || n < 0) return LGS_ERR_INVALID;
    if (!b->lastRunDevice || !b->ctx->opt.bbHostTiming || !b->hPhase.p) return LGS_ERR_INVALID;
    { const int rc = bb_settle(b); if (rc != LGS_OK) return rc; }
    const int phases = b->H + 4;
    for (int k = 0; k < n; ++k) {
        us[k] = k < phases ? (double)(b->hPhase.p[k + 1] - b->hPhase.p[k]) * 1e-3 : 0.0;
        if (mapping) mapping[k] = 0;
    }
    if (mapping) {
        const int* g = reinterpret_cast<const int*>(b->hPhase.p + kPhases);
        for (int h = b->H; h >= 0; --h) { const int k = 1 + (b->H - h); if (k < n) mapping[k] = g[h]; }
    }
    return LGS_OK;
}

// Diagnostic ("bb_host_timing" must be on before the run): nodes[q] = nodes of the levels below the root
// that the last device-only run scored for query q (the root level is the same for every query of a
// scan).  What a cost-aware placement of submaps on devices weighs (sharding.balanced_placement).
int lgs_bb_batch_query_nodes(lgs_bb_batch* b, long long* nodes, int n) {
    if (!b || !nodes || n < b->nq) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->lastRunDevice || !(c->opt.bbHostTiming || c->opt.bbCountNodes))
        return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_query_nodes: needs a device-only run with bb_count_nodes (or bb_host_timing) on");
    { const int rc = bb_settle(b); if (rc != LGS_OK) return rc; }
    for (int q = 0; q < b->nq; ++q) nodes[q] = b->hRes.p[q].nodes;
    return LGS_OK;
}

int lgs_bb_batch_path(const lgs_bb_batch* b, long long* deviceRuns, long long* exactRuns) {
    if (!b) return LGS_ERR_INVALID;
    if (deviceRuns) *deviceRuns = b->deviceRuns;
    if (exactRuns) *exactRuns = b->exactRuns;
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
