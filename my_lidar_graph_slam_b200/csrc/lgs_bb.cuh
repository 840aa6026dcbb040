// lgs_bb.cuh -- types shared by the two halves of the branch-and-bound matcher:
//   lgs_bb.cu      C ABI, host preparation, the level-synchronous exact path (host-computed index
//                  tables for near-edge points; also the fallback when a device-only run cannot be
//                  trusted) and the winner / verification / replay kernels;
//   lgs_bb_run.cu  the device-only run: ONE persistent cooperative kernel per batch run.
#pragma once
#include "lgs_internal.cuh"

namespace lgsbb {

constexpr int kFlagCapBB = 1 << 16;
constexpr int kMaxLevels = 21;
constexpr int kWarpPerNodeBelow = 8192;     // level-synchronous path: levels with fewer nodes score one warp per node

// Counter block of one run (ints).  The persistent kernel alternates between two blocks: run k
// counts in block k & 1 and clears the other one for run k + 1.
constexpr int kCtrFlags = 0;                // level-synchronous path: near-edge points
constexpr int kCtrChild = 1;                // [kCtrChild + h]: children allocated by level h = nodes of level h - 1
constexpr int kCtrUnresolved = 24;          // near-edge points the device could not decide (-> exact path)
constexpr int kCtrOverflow = 25;            // bit h: the pool of level h was too small
constexpr int kCtrFixups = 26;              // near-edge points resolved on the device (root level)
constexpr int kCtrDone = 27;                // 1 once the finalize phase has written every record
constexpr int kCtrSkipped = 28;             // rounds of 16 beams that early rejection did not have to gather
constexpr int kCounters = 32;

struct BbScan {                     // one per DISTINCT (scan, sensor pose, map resolution): hit points are map independent
    double sx, sy, st, stepT;
    int winT, nT, nTpad;            // theta slices, padded to a multiple of 4
    int nUse, beamBegin;            // usable beams (ScorePixelAccurate range filter)
    int qBegin, qCount;             // this scan's queries in qlist
    int nrx, nry, winX, winY;       // root lattice of its queries (same resolution => same window)
    long long hitBegin;             // into the hit arrays: nUse * nTpad, beam-major ([beam][theta])
    long long hitTBegin;            // into the transposed hit array: nT * beamPad ([theta][beam]), device-only run
    int beamPad, projBegin;         // nUse padded to a multiple of 4; first 8 x 32 (beam, theta) projection tile of this scan
    double invRes;                  // 1 / resolution of the maps this scan is matched against
    double originX, originY;        // floor(sensor * invRes): the cell the 12.20 fixed-point hit points are relative to
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
    long long tabBegin;             // into the base-index table: nUse * nTpad int2, beam-major (exact path)
    // device-only run: M = (min * invRes - scan origin) * 2^20 + (window offset << 20), split into its
    // 20 fraction bits (lo, 0 .. 2^20 - 1) and the whole cells (hi)
    int MloX, MhiX, MloY, MhiY;
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

struct BbBest {          // per query, 40 bytes
    unsigned long long scoreBits;   // max leaf score above threshold (as ordered bits)
    long long rank;                 // visit rank of the winning leaf
    int leaf;                       // its index in the level-0 pool
    int needReplay;
    unsigned long long rankLeaf;    // device-only run: (rank << 24 | leaf), minimised over the best-scoring leaves
    int fixups, pad;                // device-only run: near-edge points of the query's root level
};

struct BbResult {
    double score;
    int found, ix, iy, it;
    int exactReplay, fixups;
    int nodes, pad;                 // device-only run with node counting on: nodes of the levels below the root
};

struct BbFlag { int q, t, i; };                // exact path: a near-edge (query, theta, beam)
constexpr int kIdxChunk = 16;
struct IdxChunk { int scan, begin, count; };   // exact path: <= kIdxChunk queries of one scan per index-kernel block

struct LevelView { const Node* nodes; const double* scores; };
struct LevelViews { LevelView v[kMaxLevels]; };

// Everything the persistent kernel needs (passed by value as a __grid_constant__ parameter).
struct RunArgs {
    const BbQuery* qs;
    const BbScan* us;
    const int* qlist;               // queries grouped by scan
    const int* tileBegin;           // [nu + 1] prefix of root-level warp tiles per scan (for rootG)
    const double* angles;
    const double* ranges;
    int2* hits;                     // 12.20 fixed-point hit points (cells relative to the scan origin), [beam][theta] per scan
    int2* hitsT;                    // the same points, [theta][beam] per scan (mappings whose lanes walk the beams)
    Node* nodes[kMaxLevels];
    double* scores[kMaxLevels];
    int cap[kMaxLevels];            // pool capacities (nodes)
    int* ctr;                       // this run's counter block
    int* ctrNext;                   // the other block: cleared for the next run
    BbBest* best;
    BbResult* res;
    lgs_loop_record* rec;           // optional record sink (may be peer memory of another GPU), slot = recFirst + q
    const long long* recIds;        // optional ids copied into the records (NULL: the query index)
    long long recFirst;
    int recStatus;                  // also write the run's status record at slot recFirst + nq
    int nq, nu, H;
    int totalRoots, rootTiles, rootG;
    int projTiles;                  // 8 x 32 (beam, theta) tiles of the hit-point projection, all scans
    unsigned edgeUnits;             // guard band in 2^-20 cells
    int resolveUlps;
    int forceReplay;
    int countNodes;                 // count the nodes of the deeper levels per query (BbBest::pad)
    int earlyReject;                // stop a node's sum once the remaining beams cannot lift it over the threshold
    float costUs[4];                // per-pass cost model of the G = 1 / 4 / 8 / 32 mappings (us)
    unsigned long long* phaseNs;    // [kPhases] globaltimer at the phase boundaries (diagnostic)
    int* phaseG;                    // [kMaxLevels] warp mapping used per level (diagnostic)
};
constexpr int kPhases = kMaxLevels + 6;   // start, hits, root, levels H-1..0, winner, finalize

}  // namespace lgsbb

struct lgs_bb_batch {
    lgs_ctx* ctx = nullptr;
    lgs_bb_params params{};
    int nq = 0, H = 0;
    int maxRoots = 0;
    int maxNTpad = 0, maxUse = 0;   // launch extents of the projection kernels
    int spanX = 0, spanY = 0;
    std::vector<lgsbb::BbQuery> qs;
    std::vector<lgsbb::BbScan> us;  // distinct (scan, pose) pairs
    std::vector<int> qlist;         // queries grouped by scan
    std::vector<int> tileBegin[4];  // root tiles per scan for G = 1, 4, 8, 32
    std::vector<double> hAngles, hRanges;
    std::vector<long long> ids;     // record ids (empty: the query index)
    std::vector<int> fixups;
    long long nTab = 0, nHits = 0, nHitsT = 0;
    long long skippedBeams = 0;     // beams early rejection left out in the last device-only run
    int projTiles = 0;              // 8 x 32 (beam, theta) projection tiles of the batch's distinct scans
    double maxReachCells = 0.0;     // longest usable beam of the batch in cells (12.20 fixed-point range check)
    int totalRoots = 0;
    double maxAbsCells = 0.0;       // largest |coordinate| * invRes of the batch (fixed-point range check)
    bool uploaded = false, ran = false, forceReplay = false;
    bool pendingValidate = false;   // a device-only run is in flight / unvalidated
    bool lastRunDevice = false;
    bool needDeliver = false;       // an exact-path run whose records have not reached the sink / record buffer yet
    bool deviceOk = false;          // the batch fits the device-only run (fixed-point range, rank / leaf packing)
    cudaEvent_t evUpload = nullptr; // the staging blob's H2D copy (the pinned blob is reused by the next upload)
    int parity = 0;                 // counter block of the NEXT device-only run
    int rootG = 1;
    long long hint[lgsbb::kMaxLevels] = {0};       // largest node count seen per level (pool sizing)
    long long nodesPerLevel[lgsbb::kMaxLevels] = {0};
    long long gathers = 0;
    long long deviceRuns = 0, exactRuns = 0;
    double hostMs[3] = {0.0, 0.0, 0.0};   // "bb_host_timing": upload prep, run enqueue, results wait
    long long hostRuns = 0;
    // one staging blob per upload: [qs | us | qlist | tileBegin | angles | ranges | ids]
    PinBuf<char> hBlob;
    DevBuf<char> dBlob;
    size_t offQs = 0, offUs = 0, offQlist = 0, offTiles = 0, offAngles = 0, offRanges = 0, offIds = 0, offChunks = 0;
    DevBuf<int2> dHitsFix, dHitsFixT;   // device-only run: [beam][theta] and [theta][beam]
    DevBuf<double2> dHits;          // exact path
    DevBuf<int2> dTab;              // exact path: full per-query index table
    std::vector<lgsbb::IdxChunk> chunks;
    DevBuf<lgsbb::IdxChunk> dChunks;
    DevBuf<lgsbb::BbFlag> dFlags;   // exact path: near-edge list
    DevBuf<int> dCounters;          // exact path counters
    DevBuf<int> dCtr;               // device-only run: two counter blocks
    DevBuf<int> dExact;
    DevBuf<lgsbb::Node> dNodes[lgsbb::kMaxLevels];
    DevBuf<double> dScores[lgsbb::kMaxLevels];
    DevBuf<lgsbb::BbBest> dBest;
    DevBuf<lgsbb::BbResult> dRes;
    DevBuf<lgs_loop_record> dRec;   // the batch's own record buffer (default sink)
    lgs_loop_record* sink = nullptr;// external sink (peer memory / collective send buffer), NULL: dRec
    long long sinkFirst = 0;
    PinBuf<lgsbb::BbResult> hRes;
    PinBuf<int> hCounters;
    PinBuf<lgs_loop_record> hRec;
    DevBuf<unsigned long long> dPhase;      // kPhases timestamps + kMaxLevels mappings (as 64-bit words)
    PinBuf<unsigned long long> hPhase;
};

// lgs_bb_run.cu
int lgs_bb_launch_device_run(lgs_bb_batch* b);
