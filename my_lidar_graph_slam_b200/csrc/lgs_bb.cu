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
// by bb_project_kernel and get per-offset index tables computed on the host with the CPU's
// own expression (and glibc sin/cos, H5).
#include <cfloat>
#include <cmath>

#include "lgs_internal.cuh"

namespace {

constexpr int kBeamPadBB = 4;
constexpr int kFlagCapBB = 1 << 16;
constexpr int kMaxLevels = 21;

struct BbQuery {
    double sx, sy, st, stepT;
    double thrAbs;
    double minX, minY, res;
    int nx, ny, pitch;              // submap geometry
    int winX, winY, winT, nT;
    int nrx, nry;                   // roots per axis
    int nUse, nUsePad, beamBegin;   // usable beams
    int rootBegin;                  // first root of this query in the level-H pool
    long long tabBegin;             // into base-index table: nT * nUsePad int2
    const double* level[kMaxLevels];// origin() of every pyramid level
};

struct Node {            // 32 bytes
    int x, y;            // window offsets of the node's lower-left corner
    int t;               // theta index 0..nT-1
    int q;               // query
    long long rank;      // CPU visit order among nodes of the same height (lower = earlier)
    int parent;          // index in the pool one level up (-1 for roots)
    int childBase;       // first of 4 children (visit order) one level down, -1 if pruned
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

// ---- projection: base cell index of every (query, theta, usable beam) at node offset (0, 0) ----
__global__ void bb_project_kernel(const BbQuery* __restrict__ qs, const double* __restrict__ angles,
                                  const double* __restrict__ ranges, double eps,
                                  int2* __restrict__ tab, BbFlag* __restrict__ flags,
                                  int* __restrict__ flagCount) {
    const BbQuery& d = qs[blockIdx.y];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)d.nT * d.nUsePad) return;
    const int t = (int)(idx / d.nUsePad);
    const int i = (int)(idx - (long long)t * d.nUsePad);
    if (i >= d.nUse) {   // padding beam: far outside every map -> reads the zero apron
        tab[d.tabBegin + idx] = make_int2(-(1 << 28), -(1 << 28));
        return;
    }
    // nodePose.mTheta = sensorPose.mTheta + node.mTheta * stepTheta  (scan_matcher_branch_bound.cpp:96-99)
    const double theta = __dadd_rn(d.st, __dmul_rn((double)(t - d.winT), d.stepT));
    const double a = __dadd_rn(theta, angles[d.beamBegin + i]);
    double s, c;
    sincos(a, &s, &c);
    const double r = ranges[d.beamBegin + i];
    const double hx = __dadd_rn(d.sx, __dmul_rn(r, c));          // sensor_data.hpp:171-172
    const double hy = __dadd_rn(d.sy, __dmul_rn(r, s));
    const double qx = __ddiv_rn(__dsub_rn(hx, d.minX), d.res);    // grid_map.hpp:784-787
    const double qy = __ddiv_rn(__dsub_rn(hy, d.minY), d.res);
    const double fx = floor(qx), fy = floor(qy);
    const double rx = qx - fx, ry = qy - fy;
    const bool edge = !(rx >= eps && rx <= 1.0 - eps && ry >= eps && ry <= 1.0 - eps);
    int2 v = make_int2(__double2int_rd(qx), __double2int_rd(qy));
    if (edge) {
        const int k = atomicAdd(flagCount, 1);
        if (k < kFlagCapBB) {
            flags[k] = BbFlag{(int)blockIdx.y, t, i};
            v = make_int2(INT_MIN, k);      // sentinel: use the exact per-offset table k
        }
    }
    tab[d.tabBegin + idx] = v;
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
    n.x = -d.winX + (kx << height);
    n.y = -d.winY + (ky << height);
    n.t = t; n.q = q;
    n.rank = (long long)(nRoots - 1 - k);
    n.parent = -1; n.childBase = -1;
    pool[d.rootBegin + k] = n;
}

// ---- node scoring + expansion (hot kernel) -------------------------------------------------------
// One thread per node; ScorePixelAccurate::Score on pyramid level `height`.
__global__ void __launch_bounds__(128)
bb_score_kernel(const BbQuery* __restrict__ qs, const int2* __restrict__ tab,
                const int* __restrict__ exactIdx, int exactSpanX, int exactSpanY, int height,
                Node* __restrict__ nodes, double* __restrict__ scores, int nNodes,
                Node* __restrict__ next, int nextCap, int* __restrict__ nextCount,
                BbBest* __restrict__ best) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nNodes) return;
    Node n = nodes[k];
    const BbQuery& d = qs[n.q];
    const double* __restrict__ lvl = d.level[height];
    const int2* __restrict__ tb = tab + d.tabBegin + (long long)n.t * d.nUsePad;
    const int pitch = d.pitch, gx = d.nx, gy = d.ny;
    double acc = 0.0;
    const int nb = d.nUsePad;
#pragma unroll 4
    for (int i = 0; i < nb; ++i) {
        const int2 c = __ldg(tb + i);
        int ix, iy;
        if (c.x != INT_MIN) {
            ix = c.x + n.x; iy = c.y + n.y;
        } else {   // near-edge beam: indices from the host-computed exact table
            const int* e = exactIdx + (long long)c.y * (exactSpanX + exactSpanY);
            ix = e[n.x + d.winX];
            iy = e[exactSpanX + n.y + d.winY];
        }
        ix = min(max(ix, -1), gx);          // out of the map -> zero apron (Value(idx, unknown))
        iy = min(max(iy, -1), gy);
        acc = __dadd_rn(acc, __ldg(lvl + (long long)iy * pitch + ix));   // unknown cells add 0.0
    }
    scores[k] = acc;
    if (!(acc > d.thrAbs)) { nodes[k].childBase = -1; return; }   // :108 with scoreMax >= threshold
    if (height == 0) {
        atomicMax(&best[n.q].scoreBits, (unsigned long long)__double_as_longlong(acc));
        return;
    }
    const int slot = atomicAdd(nextCount, 4);
    nodes[k].childBase = slot;
    if (slot + 4 > nextCap) return;           // host grows the pool and re-runs this level
    const int w = 1 << (height - 1);
    // visit (pop) order: (x+w, y+w), (x, y+w), (x+w, y), (x, y)   (:134-137)
    const int dx[4] = {w, 0, w, 0}, dy[4] = {w, w, 0, 0};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        Node m;
        m.x = n.x + dx[c]; m.y = n.y + dy[c]; m.t = n.t; m.q = n.q;
        m.rank = n.rank * 4 + c;
        m.parent = k; m.childBase = -1;
        next[slot + c] = m;
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
            const int cb = lv.v[h].nodes[idx].childBase;                // s > best >= thr => expanded
            for (int c = 3; c >= 0; --c) { stackIdx[sp] = cb + c; stackH[sp] = h - 1; ++sp; }
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
    int maxNT = 0, maxUsePad = 0, maxRoots = 0;
    int spanX = 0, spanY = 0;
    std::vector<BbQuery> qs;
    std::vector<double> hAngles, hRanges;
    std::vector<int> fixups;
    long long nTab = 0;
    int totalRoots = 0;
    bool uploaded = false, ran = false, forceReplay = false;
    long long nodesPerLevel[kMaxLevels] = {0};
    long long gathers = 0;
    DevBuf<BbQuery> dQs;
    DevBuf<double> dAngles, dRanges;
    DevBuf<int2> dTab;
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
    if (!b) return LGS_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    b->dQs.release(); b->dAngles.release(); b->dRanges.release(); b->dTab.release();
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
    lgs_ctx* c = b->ctx;
    const int n = scans->n_scans;
    if (n < 0 || (n > 0 && (!scans->beam_begin || !scans->sensor_pose || !pyramids)))
        return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_upload: bad arguments");
    LGS_CUDA(c, cudaSetDevice(c->device));
    const lgs_bb_params& p = b->params;
    const int H = b->H;
    b->uploaded = false; b->ran = false;
    b->nq = n;
    b->qs.assign(n, BbQuery{});
    b->fixups.assign(n, 0);
    b->hAngles.clear(); b->hRanges.clear();
    b->maxNT = 0; b->maxUsePad = kBeamPadBB; b->maxRoots = 0; b->spanX = 0; b->spanY = 0;
    long long nTab = 0, roots = 0;
    const int winSizeMax = 1 << H;
    for (int q = 0; q < n; ++q) {
        const int b0 = scans->beam_begin[q], b1 = scans->beam_begin[q + 1];
        const int nb = b1 - b0;
        if (nb <= 0) return lgs_fail(c, LGS_ERR_INVALID, "bb: scan %d has no beams", q);
        const lgs_pyramid* pyr = pyramids[q];
        if (!pyr || lgs_pyramid_levels(pyr) < H + 1)
            return lgs_fail(c, LGS_ERR_INVALID, "bb: query %d needs a pyramid with %d levels", q, H + 1);
        const lgs_grid* g0 = lgs_pyramid_level(pyr, 0);
        if (g0->ctx->device != c->device)
            return lgs_fail(c, LGS_ERR_INVALID, "bb: pyramid of query %d lives on another device", q);
        BbQuery& d = b->qs[q];
        d.sx = scans->sensor_pose[3 * q]; d.sy = scans->sensor_pose[3 * q + 1];
        d.st = scans->sensor_pose[3 * q + 2];
        d.minX = g0->min_x; d.minY = g0->min_y; d.res = g0->res;
        d.nx = g0->nx; d.ny = g0->ny; d.pitch = g0->pitch;
        for (int h = 0; h <= H; ++h) d.level[h] = lgs_pyramid_level(pyr, h)->origin();
        // ComputeSearchStep (scan_matcher_branch_bound.cpp:178-197)
        double maxR = scans->ranges[b0];
        for (int i = b0 + 1; i < b1; ++i) maxR = std::max(maxR, scans->ranges[i]);
        const double maxRange = std::min(maxR, p.scan_range_max);
        const double th = d.res / maxRange;
        const double stepX = d.res, stepY = d.res;
        d.stepT = std::acos(1.0 - 0.5 * th * th);
        d.winX = static_cast<int>(std::ceil(0.5 * p.range_x / stepX));                 // :68-73
        d.winY = static_cast<int>(std::ceil(0.5 * p.range_y / stepY));
        d.winT = static_cast<int>(std::ceil(0.5 * p.range_theta / d.stepT));
        if (!(d.stepT > 0.0) || d.winT < 0 || d.winT > (1 << 20))
            return lgs_fail(c, LGS_ERR_INVALID, "bb: scan %d gives stepTheta=%g winTheta=%d", q, d.stepT, d.winT);
        d.nT = 2 * d.winT + 1;
        d.nrx = (2 * d.winX) / winSizeMax + 1;                                         // :85-86
        d.nry = (2 * d.winY) / winSizeMax + 1;
        const double thr = normThr ? normThr[q] : DBL_MIN;
        d.thrAbs = thr * static_cast<double>(static_cast<size_t>(nb));                 // :75-76
        // ScorePixelAccurate range filter (score_function_pixel_accurate.cpp:27-41)
        const double sMin = scans->range_min ? scans->range_min[q] : 0.0;
        const double sMax = scans->range_max ? scans->range_max[q] : HUGE_VAL;
        const double minRange = std::max(p.score_range_min, sMin);
        const double maxRangeS = std::min(p.score_range_max, sMax);
        d.beamBegin = (int)b->hAngles.size();
        for (int i = b0; i < b1; ++i) {
            const double r = scans->ranges[i];
            if (r >= maxRangeS || r <= minRange) continue;
            b->hAngles.push_back(scans->angles[i]);
            b->hRanges.push_back(r);
        }
        d.nUse = (int)b->hAngles.size() - d.beamBegin;
        d.nUsePad = std::max(kBeamPadBB, (d.nUse + kBeamPadBB - 1) / kBeamPadBB * kBeamPadBB);
        d.tabBegin = nTab;
        nTab += (long long)d.nT * d.nUsePad;
        const long long nr = (long long)d.nrx * d.nry * d.nT;
        if (roots + nr > (1LL << 30)) return lgs_fail(c, LGS_ERR_INVALID, "bb: too many root nodes");
        d.rootBegin = (int)roots;
        roots += nr;
        b->maxNT = std::max(b->maxNT, d.nT);
        b->maxUsePad = std::max(b->maxUsePad, d.nUsePad);
        b->maxRoots = std::max<long long>(b->maxRoots, nr);
        b->spanX = std::max(b->spanX, d.nrx * winSizeMax);
        b->spanY = std::max(b->spanY, d.nry * winSizeMax);
    }
    b->nTab = nTab;
    b->totalRoots = (int)roots;
    if (n == 0) { b->uploaded = true; return LGS_OK; }
    const size_t nk = b->hAngles.size();
    LGS_CUDA(c, b->dQs.reserve(n));
    LGS_CUDA(c, b->dAngles.reserve(std::max<size_t>(nk, 1)));
    LGS_CUDA(c, b->dRanges.reserve(std::max<size_t>(nk, 1)));
    LGS_CUDA(c, b->dTab.reserve(nTab));
    LGS_CUDA(c, b->dFlags.reserve(kFlagCapBB));
    LGS_CUDA(c, b->dCounters.reserve(2 + kMaxLevels));
    LGS_CUDA(c, b->hCounters.reserve(2 + kMaxLevels));
    LGS_CUDA(c, b->dBest.reserve(n));
    LGS_CUDA(c, b->dRes.reserve(n));
    LGS_CUDA(c, b->hRes.reserve(n));
    LGS_CUDA(c, b->dNodes[H].reserve(roots));
    LGS_CUDA(c, b->dScores[H].reserve(roots));
    LGS_CUDA(c, cudaMemcpyAsync(b->dQs.p, b->qs.data(), n * sizeof(BbQuery), cudaMemcpyHostToDevice, c->stream));
    if (nk) {
        LGS_CUDA(c, cudaMemcpyAsync(b->dAngles.p, b->hAngles.data(), nk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        LGS_CUDA(c, cudaMemcpyAsync(b->dRanges.p, b->hRanges.data(), nk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));   // host vectors may be reused by the caller
    b->uploaded = true;
    return LGS_OK;
}

int lgs_bb_batch_run(lgs_bb_batch* b) {
    if (!b) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->uploaded) return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_run before upload");
    const int H = b->H, n = b->nq;
    for (int h = 0; h < kMaxLevels; ++h) b->nodesPerLevel[h] = 0;
    b->gathers = 0;
    if (n == 0) { b->ran = true; return LGS_OK; }
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaMemsetAsync(b->dCounters.p, 0, (2 + kMaxLevels) * sizeof(int), c->stream));
    {
        const long long per = (long long)b->maxNT * b->maxUsePad;
        dim3 gridDim((unsigned)((per + 255) / 256), n);
        bb_project_kernel<<<gridDim, 256, 0, c->stream>>>(b->dQs.p, b->dAngles.p, b->dRanges.p,
                                                          g_lgs_edge_eps, b->dTab.p, b->dFlags.p,
                                                          b->dCounters.p);
        LGS_LAUNCH_CHECK(c);
        dim3 gridR((b->maxRoots + 127) / 128, n);
        bb_roots_kernel<<<gridR, 128, 0, c->stream>>>(b->dQs.p, n, H, b->dNodes[H].p);
        LGS_LAUNCH_CHECK(c);
        bb_init_best_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(b->dBest.p, n);
        LGS_LAUNCH_CHECK(c);
    }
    // Near-edge beams: exact per-offset index tables from the host (CPU expression + glibc).
    LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p, b->dCounters.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    const int nFlag = b->hCounters.p[0];
    std::fill(b->fixups.begin(), b->fixups.end(), 0);
    const int spanX = b->spanX, spanY = b->spanY;
    if (nFlag > kFlagCapBB)
        return lgs_fail(c, LGS_ERR_OVERFLOW, "bb: %d near-edge points exceed the fix-up list", nFlag);
    if (nFlag > 0) {
        std::vector<BbFlag> fl(nFlag);
        LGS_CUDA(c, cudaMemcpy(fl.data(), b->dFlags.p, nFlag * sizeof(BbFlag), cudaMemcpyDeviceToHost));
        std::vector<int> exact((size_t)nFlag * (spanX + spanY), 0);
        for (int k = 0; k < nFlag; ++k) {
            const BbQuery& d = b->qs[fl[k].q];
            const double stepX = d.res, stepY = d.res;
            const double theta = d.st + static_cast<double>(fl[k].t - d.winT) * d.stepT;
            const double a = theta + b->hAngles[d.beamBegin + fl[k].i];
            const double cosT = std::cos(a), sinT = std::sin(a);
            const double r = b->hRanges[d.beamBegin + fl[k].i];
            int* e = exact.data() + (size_t)k * (spanX + spanY);
            for (int o = 0; o < spanX; ++o) {
                const double px = d.sx + static_cast<double>(o - d.winX) * stepX;       // :96-97
                e[o] = static_cast<int>(std::floor(((px + r * cosT) - d.minX) / d.res));
            }
            for (int o = 0; o < spanY; ++o) {
                const double py = d.sy + static_cast<double>(o - d.winY) * stepY;
                e[spanX + o] = static_cast<int>(std::floor(((py + r * sinT) - d.minY) / d.res));
            }
            b->fixups[fl[k].q]++;
        }
        LGS_CUDA(c, b->dExact.reserve(exact.size()));
        LGS_CUDA(c, cudaMemcpy(b->dExact.p, exact.data(), exact.size() * sizeof(int), cudaMemcpyHostToDevice));
    } else {
        LGS_CUDA(c, b->dExact.reserve(1));
    }

    // Level-synchronous expansion of the static-threshold superset.
    int nNodes = b->totalRoots;
    for (int h = H; h >= 0; --h) {
        b->nodesPerLevel[h] = nNodes;
        if (nNodes == 0) break;
        LGS_CUDA(c, b->dScores[h].reserve(nNodes));
        int* nextCount = b->dCounters.p + 1 + h;
        if (h > 0 && b->dNodes[h - 1].cap < (size_t)std::min<long long>(4LL * nNodes, 1 << 16))
            LGS_CUDA(c, b->dNodes[h - 1].reserve(std::min<long long>(4LL * nNodes, 1 << 16)));
        for (int attempt = 0; attempt < 2; ++attempt) {
            LGS_CUDA(c, cudaMemsetAsync(nextCount, 0, sizeof(int), c->stream));
            bb_score_kernel<<<(nNodes + 127) / 128, 128, 0, c->stream>>>(
                b->dQs.p, b->dTab.p, b->dExact.p, spanX, spanY, h, b->dNodes[h].p, b->dScores[h].p,
                nNodes, h > 0 ? b->dNodes[h - 1].p : nullptr, h > 0 ? (int)b->dNodes[h - 1].cap : 0,
                nextCount, b->dBest.p);
            LGS_LAUNCH_CHECK(c);
            if (h == 0) break;
            LGS_CUDA(c, cudaMemcpyAsync(b->hCounters.p + 1 + h, nextCount, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            LGS_CUDA(c, cudaStreamSynchronize(c->stream));
            const int want = b->hCounters.p[1 + h];
            if ((size_t)want <= b->dNodes[h - 1].cap) break;
            if (attempt == 1) return lgs_fail(c, LGS_ERR_OVERFLOW, "bb: level %d pool overflow", h - 1);
            LGS_CUDA(c, b->dNodes[h - 1].reserve((size_t)want + want / 4));   // grow, redo this level
        }
        nNodes = h > 0 ? b->hCounters.p[1 + h] : 0;
    }
    for (int h = 0; h <= H; ++h) b->gathers += b->nodesPerLevel[h];   // refined per query below
    // Winner, verification, replay.
    const int nLeaves = (int)b->nodesPerLevel[0];
    if (nLeaves > 0) {
        bb_leaf_rank_kernel<<<(nLeaves + 127) / 128, 128, 0, c->stream>>>(b->dNodes[0].p, b->dScores[0].p, nLeaves, b->dBest.p);
        LGS_LAUNCH_CHECK(c);
        bb_leaf_pick_kernel<<<(nLeaves + 127) / 128, 128, 0, c->stream>>>(b->dNodes[0].p, b->dScores[0].p, nLeaves, b->dBest.p);
        LGS_LAUNCH_CHECK(c);
    }
    LevelViews lv;
    for (int h = 0; h < kMaxLevels; ++h) lv.v[h] = LevelView{b->dNodes[h].p, b->dScores[h].p};
    bb_verify_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(b->dQs.p, n, H, lv, b->dBest.p, b->dRes.p,
                                                          b->forceReplay ? 1 : 0);
    LGS_LAUNCH_CHECK(c);
    bb_replay_kernel<<<(n + 31) / 32, 32, 0, c->stream>>>(b->dQs.p, n, H, lv, b->dBest.p, b->dRes.p);
    LGS_LAUNCH_CHECK(c);
    b->ran = true;
    return LGS_OK;
}

int lgs_bb_batch_results(lgs_bb_batch* b, lgs_match_result* out) {
    if (!b || (!out && b->nq > 0)) return LGS_ERR_INVALID;
    lgs_ctx* c = b->ctx;
    if (!b->ran) return lgs_fail(c, LGS_ERR_INVALID, "bb_batch_results before run");
    if (b->nq == 0) return LGS_OK;
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
        o.step_x = d.res; o.step_y = d.res; o.step_t = d.stepT;
        o.score = r.score;
        o.n_scored = total;          // batch-wide count of nodes scored (all queries)
        o.exact_replay = r.exactReplay;
        o.reserved = 0;
    }
    return LGS_OK;
}

int lgs_bb_batch_work(const lgs_bb_batch* b, long long* nodesPerLevel, int nLevels, long long* gathers) {
    if (!b) return LGS_ERR_INVALID;
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
