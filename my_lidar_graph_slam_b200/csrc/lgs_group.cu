// lgs_group.cu -- loop detection across the GPUs of one box.
//
// The reference's LoopDetectorBranchBound::Detect walks independent (pose-graph node, local map)
// pairs one after the other (mapping/loop_detector_branch_bound.cpp:38-90); nothing is carried from
// one pair to the next, so the pairs shard by the submap they name.  Two ways to drive that:
//
//  * lgs_group_*   ONE process, one lgs_ctx per device, one persistent host thread per device (this is
//                  what the C++ adapter uses: eight ranks' launch chains issued from eight processes of
//                  one host contend for the launch path, eight threads of one process do not).  Peer
//                  access is enabled between all members; every member's finalize phase stores its
//                  32-byte records STRAIGHT INTO THE ROOT DEVICE'S gather buffer over NVLink (plain peer
//                  stores from the kernel, lgs_bb_batch_set_record_sink) -- no host staging, no separate
//                  copy, no collective library for 16 KB.  One D2H of the gathered records follows.
//  * lgs_comm_*    one process PER device (torchrun-style launch): an all-gather of the records on the
//                  context stream, device to device, through NCCL.  The batch's record sink is the
//                  rank's own slice of the receive buffer, so the all-gather runs in place on data the
//                  kernel wrote there itself.  libnccl.so.2 is resolved at run time (dlopen), so the
//                  library has no link-time dependency on it and shares the copy a host framework may
//                  already have loaded.
#include <dlfcn.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "lgs_bb.cuh"

using namespace lgsbb;

namespace {

// One persistent host thread per group member: runs the closures handed to it, in order.
struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void()> task;
    bool hasTask = false, done = true, quit = false;

    void start() {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(mu);
            for (;;) {
                cv.wait(lk, [this] { return hasTask || quit; });
                if (quit) return;
                std::function<void()> t = std::move(task);
                hasTask = false;
                lk.unlock();
                t();
                lk.lock();
                done = true;
                cv.notify_all();
            }
        });
    }
    void submit(std::function<void()> t) {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return done; });
        task = std::move(t);
        hasTask = true;
        done = false;
        cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return done; });
    }
    void stop() {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [this] { return done; });
            quit = true;
            cv.notify_all();
        }
        if (th.joinable()) th.join();
    }
};

}  // namespace

struct lgs_group {
    std::vector<int> devices;
    std::vector<lgs_ctx*> ctx;
    std::vector<Worker*> workers;
    lgs_loop_record* gather = nullptr;      // on devices[0], peer-mapped into every member
    size_t gatherCap = 0;
    char err[512] = {0};
};

struct lgs_group_bb {
    lgs_group* g = nullptr;
    lgs_bb_params params{};
    std::vector<lgs_bb_batch*> batch;       // one per member
    // per-member share of the last detect
    struct Share {
        std::vector<int> pairs;             // global pair indices, in member order
        std::vector<int> pairScan;          // index into the member's own scan list
        std::vector<lgs_pyramid*> pyr;
        std::vector<double> thr;
        std::vector<long long> ids;
        std::vector<int> beamBegin;
        std::vector<double> angles, ranges, poses, rmin, rmax;
        int rc = LGS_OK;
    };
    std::vector<Share> share;
    std::vector<lgs_loop_record> host;      // gathered records (member order)
    int nPairs = 0;
};

static int group_fail(lgs_group* g, int code, const char* fmt, ...) {
    if (g) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(g->err, sizeof(g->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

extern "C" {

int lgs_group_create(const int* devices, int n, lgs_group** out) {
    if (!devices || n < 1 || !out) return LGS_ERR_INVALID;
    *out = nullptr;
    lgs_group* g = new lgs_group();
    for (int k = 0; k < n; ++k) {
        for (int j = 0; j < k; ++j)
            if (devices[j] == devices[k]) { lgs_group_destroy(g); return LGS_ERR_INVALID; }
        lgs_ctx* c = nullptr;
        const int rc = lgs_ctx_create(devices[k], &c);
        if (rc != LGS_OK) { lgs_group_destroy(g); return rc; }
        g->devices.push_back(devices[k]);
        g->ctx.push_back(c);
    }
    // every member may store into (and read from) every other member's memory
    for (int k = 0; k < n; ++k) {
        cudaSetDevice(devices[k]);
        for (int j = 0; j < n; ++j) {
            if (j == k) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[k], devices[j]);
            if (!can) { lgs_group_destroy(g); return LGS_ERR_CUDA; }       // no host-staged fallback: fail loudly
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { lgs_group_destroy(g); return LGS_ERR_CUDA; }
            cudaGetLastError();
        }
    }
    for (int k = 0; k < n; ++k) {
        Worker* w = new Worker();
        w->start();
        g->workers.push_back(w);
    }
    *out = g;
    return LGS_OK;
}

int lgs_group_destroy(lgs_group* g) {
    if (!g) return LGS_OK;
    for (Worker* w : g->workers) { w->stop(); delete w; }
    if (g->gather && !g->ctx.empty()) {
        cudaSetDevice(g->devices[0]);
        cudaDeviceSynchronize();
        cudaFree(g->gather);
    }
    for (lgs_ctx* c : g->ctx) lgs_ctx_destroy(c);
    delete g;
    return LGS_OK;
}

int lgs_group_size(const lgs_group* g) { return g ? (int)g->ctx.size() : 0; }

lgs_ctx* lgs_group_ctx(lgs_group* g, int member) {
    return (g && member >= 0 && member < (int)g->ctx.size()) ? g->ctx[member] : nullptr;
}

const char* lgs_group_last_error(const lgs_group* g) { return g ? g->err : "null group"; }

int lgs_group_bb_create(lgs_group* g, const lgs_bb_params* params, lgs_group_bb** out) {
    if (!g || !params || !out) return LGS_ERR_INVALID;
    *out = nullptr;
    lgs_group_bb* d = new lgs_group_bb();
    d->g = g;
    d->params = *params;
    d->share.resize(g->ctx.size());
    for (size_t m = 0; m < g->ctx.size(); ++m) {
        lgs_bb_batch* b = nullptr;
        const int rc = lgs_bb_batch_create(g->ctx[m], params, &b);
        if (rc != LGS_OK) {
            group_fail(g, rc, "group_bb_create: member %zu: %s", m, lgs_ctx_last_error(g->ctx[m]));
            lgs_group_bb_destroy(d);
            return rc;
        }
        d->batch.push_back(b);
    }
    *out = d;
    return LGS_OK;
}

int lgs_group_bb_destroy(lgs_group_bb* d) {
    if (!d) return LGS_OK;
    for (Worker* w : d->g->workers) w->wait();
    for (lgs_bb_batch* b : d->batch) lgs_bb_batch_destroy(b);
    delete d;
    return LGS_OK;
}

int lgs_group_bb_detect(lgs_group_bb* d, const lgs_scan_batch* scans, int nPairs, const int* pairScan,
                        lgs_pyramid* const* pyramids, const double* normThr, lgs_match_result* out) {
    if (!d || !scans || nPairs < 0 || (nPairs > 0 && (!pyramids || !out))) return LGS_ERR_INVALID;
    lgs_group* g = d->g;
    const int nm = (int)g->ctx.size();
    d->nPairs = nPairs;
    if (nPairs == 0) return LGS_OK;
    // ---- shard: a pair runs where its submap's pyramid lives ------------------------------------------
    for (auto& s : d->share) {
        s.pairs.clear(); s.pairScan.clear(); s.pyr.clear(); s.thr.clear(); s.ids.clear();
        s.beamBegin.assign(1, 0); s.angles.clear(); s.ranges.clear(); s.poses.clear(); s.rmin.clear(); s.rmax.clear();
        s.rc = LGS_OK;
    }
    std::vector<std::vector<int>> scanLocal(nm, std::vector<int>(std::max(scans->n_scans, 1), -1));
    for (int q = 0; q < nPairs; ++q) {
        const int sq = pairScan ? pairScan[q] : q;
        if (sq < 0 || sq >= scans->n_scans) return group_fail(g, LGS_ERR_INVALID, "group_bb_detect: pair %d names scan %d", q, sq);
        const lgs_grid* g0 = lgs_pyramid_level(pyramids[q], 0);
        if (!g0) return group_fail(g, LGS_ERR_INVALID, "group_bb_detect: pair %d has no pyramid", q);
        int m = -1;
        for (int k = 0; k < nm; ++k) if (g->ctx[k] == g0->ctx) m = k;
        if (m < 0)                               // built on a foreign context: any member on the same device will do
            for (int k = 0; k < nm; ++k) if (g->devices[k] == g0->ctx->device) m = k;
        if (m < 0) return group_fail(g, LGS_ERR_INVALID, "group_bb_detect: the pyramid of pair %d lives on device %d, "
                                     "which is not in the group", q, g0->ctx->device);
        auto& s = d->share[m];
        int& ls = scanLocal[m][sq];
        if (ls < 0) {                            // first use of this scan on this member: append it to the member's list
            ls = (int)s.beamBegin.size() - 1;
            const int b0 = scans->beam_begin[sq], b1 = scans->beam_begin[sq + 1];
            s.angles.insert(s.angles.end(), scans->angles + b0, scans->angles + b1);
            s.ranges.insert(s.ranges.end(), scans->ranges + b0, scans->ranges + b1);
            s.beamBegin.push_back((int)s.ranges.size());
            for (int k = 0; k < 3; ++k) s.poses.push_back(scans->sensor_pose[3 * sq + k]);
            s.rmin.push_back(scans->range_min ? scans->range_min[sq] : 0.0);
            s.rmax.push_back(scans->range_max ? scans->range_max[sq] : HUGE_VAL);
        }
        s.pairs.push_back(q);
        s.pairScan.push_back(ls);
        s.pyr.push_back(pyramids[q]);
        if (normThr) s.thr.push_back(normThr[q]);
        s.ids.push_back(q);
    }
    // ---- the root's gather buffer: member m's records at [first[m], first[m] + share size) ---------------
    const size_t needSlots = (size_t)nPairs + (size_t)nm;      // + one status record behind every member's share
    if (needSlots > g->gatherCap) {
        cudaSetDevice(g->devices[0]);
        for (Worker* w : g->workers) w->wait();
        if (g->gather) { cudaDeviceSynchronize(); cudaFree(g->gather); g->gather = nullptr; g->gatherCap = 0; }
        const size_t want = needSlots + needSlots / 2 + 64;
        if (cudaMalloc(&g->gather, want * sizeof(lgs_loop_record)) != cudaSuccess)
            return group_fail(g, LGS_ERR_NOMEM, "group_bb_detect: gather buffer of %zu records", want);
        g->gatherCap = want;
    }
    std::vector<long long> first(nm, 0);
    { long long run = 0; for (int m = 0; m < nm; ++m) { first[m] = run; run += (long long)d->share[m].pairs.size() + 1; } }
    // ---- one host thread per member: upload, ONE kernel launch, wait + validate ----------------------------
    for (int m = 0; m < nm; ++m) {
        auto& s = d->share[m];
        if (s.pairs.empty()) continue;
        lgs_bb_batch* b = d->batch[m];
        lgs_loop_record* sink = g->gather;
        const long long f = first[m];
        g->workers[m]->submit([&s, b, sink, f]() {
            const int n = (int)s.pairs.size();
            const lgs_scan_batch sb{(int)s.beamBegin.size() - 1, s.beamBegin.data(), s.angles.data(), s.ranges.data(),
                                    s.poses.data(), s.rmin.data(), s.rmax.data()};
            int rc = lgs_bb_batch_set_record_ids(b, s.ids.data(), n);
            if (rc == LGS_OK) rc = lgs_bb_batch_set_record_sink(b, sink, f);
            if (rc == LGS_OK) rc = lgs_bb_batch_upload_pairs(b, &sb, n, s.pairScan.data(), s.pyr.data(),
                                                             s.thr.empty() ? nullptr : s.thr.data());
            if (rc == LGS_OK) rc = lgs_bb_batch_run(b);
            if (rc == LGS_OK) rc = lgs_bb_batch_settle(b);       // the records are in the root's buffer when this returns
            s.rc = rc;
        });
    }
    int rcAll = LGS_OK;
    for (int m = 0; m < nm; ++m) {
        if (d->share[m].pairs.empty()) continue;
        g->workers[m]->wait();
        if (d->share[m].rc != LGS_OK && rcAll == LGS_OK) {
            rcAll = d->share[m].rc;
            group_fail(g, rcAll, "group_bb_detect: member %d (device %d): %s", m, g->devices[m], lgs_ctx_last_error(g->ctx[m]));
        }
    }
    if (rcAll != LGS_OK) return rcAll;
    // ---- one D2H of the gathered records ------------------------------------------------------------------------
    d->host.resize(needSlots);
    {
        const int rc = lgs_device_download(g->ctx[0], g->gather, d->host.data(), (unsigned long long)needSlots * sizeof(lgs_loop_record));
        if (rc != LGS_OK) return group_fail(g, rc, "group_bb_detect: download of the gathered records: %s", lgs_ctx_last_error(g->ctx[0]));
    }
    // found / indices / score come from the exchanged records; window sizes and steps are host-side
    // facts of the member's batch (no device involvement)
    for (int m = 0; m < nm; ++m) {
        const auto& s = d->share[m];
        const lgs_bb_batch* b = d->batch[m];
        long long total = 0;
        for (int h = 0; h <= b->H; ++h) total += b->nodesPerLevel[h];
        for (size_t k = 0; k < s.pairs.size(); ++k) {
            const lgs_loop_record& r = d->host[first[m] + (long long)k];
            if (r.id != (long long)s.pairs[k])
                return group_fail(g, LGS_ERR_CUDA, "group_bb_detect: record %lld of member %d carries id %lld, expected %d",
                                  first[m] + (long long)k, m, r.id, s.pairs[k]);
            const BbQuery& qd = b->qs[k];
            lgs_match_result& o = out[s.pairs[k]];
            o.found = r.found; o.ix = r.ix; o.iy = r.iy; o.it = r.it;
            o.win_x = qd.winX; o.win_y = qd.winY; o.win_t = qd.winT;
            o.n_fixups = b->fixups[k];
            o.step_x = qd.res; o.step_y = qd.res; o.step_t = b->us[qd.scan].stepT;
            o.score = r.score;
            const bool perQuery = b->lastRunDevice && (b->ctx->opt.bbCountNodes || b->ctx->opt.bbHostTiming);
            o.n_scored = perQuery ? (long long)qd.nrx * qd.nry * qd.nT + b->hRes.p[k].nodes : total;
            o.exact_replay = b->hRes.p[k].exactReplay;
            o.reserved = b->lastRunDevice ? 0 : 1;
        }
    }
    return LGS_OK;
}

int lgs_group_bb_records(const lgs_group_bb* d, lgs_loop_record* out) {
    if (!d || (!out && d->nPairs > 0)) return LGS_ERR_INVALID;
    // pair order (the gather buffer is in member order)
    for (const lgs_loop_record& r : d->host)
        if (r.id >= 0 && r.id < d->nPairs) out[r.id] = r;
    return LGS_OK;
}

// ---- one process per device: NCCL all-gather of the records ---------------------------------------------------
}  // extern "C"

namespace {

struct NcclApi {
    struct Id128 { char b[128]; };          // ncclUniqueId: passed by value
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy) api.lib = h;
    });
    return api.lib ? &api : nullptr;
}

}  // namespace

struct lgs_comm {
    lgs_ctx* ctx = nullptr;
    void* comm = nullptr;
    int world = 1, rank = 0;
};

extern "C" {

int lgs_comm_unique_id(void* id128) {
    if (!id128) return LGS_ERR_INVALID;
    NcclApi* api = nccl_api();
    if (!api) return LGS_ERR_CUDA;
    return api->GetUniqueId(id128) == 0 ? LGS_OK : LGS_ERR_CUDA;
}

int lgs_comm_create(lgs_ctx* ctx, int world, int rank, const void* id128, lgs_comm** out) {
    if (!ctx || !out || world < 1 || rank < 0 || rank >= world || !id128) return LGS_ERR_INVALID;
    *out = nullptr;
    NcclApi* api = nccl_api();
    if (!api) return lgs_fail(ctx, LGS_ERR_CUDA, "comm_create: libnccl.so.2 not found");
    LGS_CUDA(ctx, cudaSetDevice(ctx->device));
    lgs_comm* c = new lgs_comm();
    c->ctx = ctx; c->world = world; c->rank = rank;
    NcclApi::Id128 id;
    memcpy(id.b, id128, sizeof(id.b));
    const int rc = api->CommInitRank(&c->comm, world, id, rank);
    if (rc != 0) {
        delete c;
        return lgs_fail(ctx, LGS_ERR_CUDA, "comm_create: ncclCommInitRank -> %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
    }
    *out = c;
    return LGS_OK;
}

int lgs_comm_destroy(lgs_comm* c) {
    if (!c) return LGS_OK;
    NcclApi* api = nccl_api();
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (api && c->comm) api->CommDestroy(c->comm);
    delete c;
    return LGS_OK;
}

int lgs_comm_all_gather_records(lgs_comm* c, const void* sendDevice, void* recvDevice, int count) {
    if (!c || !recvDevice || count < 0) return LGS_ERR_INVALID;
    if (count == 0) return LGS_OK;
    NcclApi* api = nccl_api();
    lgs_ctx* ctx = c->ctx;
    if (!api) return lgs_fail(ctx, LGS_ERR_CUDA, "comm_all_gather: libnccl.so.2 not found");
    LGS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)count * sizeof(lgs_loop_record);
    // in place when the caller's kernels wrote the rank's records into its own slice of recvDevice
    const void* send = sendDevice ? sendDevice : static_cast<const char*>(recvDevice) + (size_t)c->rank * bytes;
    const int rc = api->AllGather(send, recvDevice, bytes, /* ncclUint8 */ 1, c->comm, ctx->stream);
    if (rc != 0) return lgs_fail(ctx, LGS_ERR_CUDA, "comm_all_gather: ncclAllGather -> %s", api->GetErrorString ? api->GetErrorString(rc) : "error");
    return LGS_OK;
}

}  // extern "C"
