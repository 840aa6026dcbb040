// lgs_core.cu -- context and dense device grid of the sm_100a backend.
//
// The device grid replaces GridMap<T>'s patch-tiled storage (grid_map/grid_map.hpp:22-318,
// grid_map_patch.hpp:15-75) on the hot path.  Reads of unallocated patches and of cells
// outside the map both return the unknown value 0.0 through Value(x, y, unknown)
// (grid_map.hpp:859-873), so a dense row-major array whose unknown cells hold 0.0 and which
// is surrounded by a zero apron is value-equivalent for every reader on the path.
#include "lgs_internal.cuh"

int lgs_fail(lgs_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

double g_lgs_edge_eps = LGS_EDGE_EPS_DEFAULT;

extern "C" {

void lgs_set_edge_eps(double eps) { g_lgs_edge_eps = (eps > 0.0 && eps < 0.5) ? eps : LGS_EDGE_EPS_DEFAULT; }
double lgs_get_edge_eps(void) { return g_lgs_edge_eps; }

const char* lgs_version(void) { return "lgs_b200 0.1 (sm_100a)"; }

int lgs_ctx_create(int device, lgs_ctx** out) {
    if (!out) return LGS_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n)
        return LGS_ERR_CUDA;   // no CPU fallback: the caller must fail loudly
    lgs_ctx* c = new lgs_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return LGS_ERR_CUDA; }
    if (prop.major < 10) { delete c; return LGS_ERR_CUDA; }   // sm_100a code only
    c->sm_count = prop.multiProcessorCount;
    {   // keep freed stream-ordered allocations (pyramid slabs) in the pool instead of returning them
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        delete c;
        return LGS_ERR_CUDA;
    }
    *out = c;
    return LGS_OK;
}

int lgs_ctx_destroy(lgs_ctx* c) {
    if (!c) return LGS_OK;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    c->scratch.release();
    if (c->integ) { c->integ->release(); delete c->integ; }
    lgs_cost_ws_destroy(c->cost);
    delete c;
    return LGS_OK;
}

const char* lgs_ctx_last_error(const lgs_ctx* c) { return c ? c->err : "null context"; }

int lgs_ctx_synchronize(lgs_ctx* c) {
    if (!c) return LGS_ERR_INVALID;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    return LGS_OK;
}

int lgs_host_pin(lgs_ctx* c, void* ptr, unsigned long long bytes) {
    if (!c || !ptr || bytes == 0) return LGS_ERR_INVALID;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return LGS_OK;
}

int lgs_host_unpin(lgs_ctx* c, void* ptr) {
    if (!c || !ptr) return LGS_ERR_INVALID;
    LGS_CUDA(c, cudaHostUnregister(ptr));
    return LGS_OK;
}

void* lgs_ctx_stream(lgs_ctx* c) { return c ? (void*)c->stream : nullptr; }

int lgs_ctx_timer_start(lgs_ctx* c) {
    if (!c) return LGS_ERR_INVALID;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    return LGS_OK;
}

int lgs_ctx_timer_stop(lgs_ctx* c, float* ms) {
    if (!c || !ms) return LGS_ERR_INVALID;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    LGS_CUDA(c, cudaEventSynchronize(c->ev1));
    LGS_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return LGS_OK;
}

long long lgs_ctx_launch_count(const lgs_ctx* c) { return c ? c->launches : 0; }

int lgs_grid_create(lgs_ctx* c, int nx, int ny, double min_x, double min_y, double res,
                    int apron, lgs_grid** out) {
    if (!c || !out) return LGS_ERR_INVALID;
    *out = nullptr;
    if (nx < 0 || ny < 0 || apron < 1 || !(res > 0.0))
        return lgs_fail(c, LGS_ERR_INVALID, "grid_create: nx=%d ny=%d apron=%d res=%g", nx, ny,
                        apron, res);
    const long long pitch = (long long)nx + 2LL * apron, rows = (long long)ny + 2LL * apron;
    if (pitch * rows >= (1LL << 31))
        return lgs_fail(c, LGS_ERR_INVALID, "grid_create: %lld cells exceed int32 indexing",
                        pitch * rows);
    LGS_CUDA(c, cudaSetDevice(c->device));
    lgs_grid* g = new lgs_grid();
    g->ctx = c; g->nx = nx; g->ny = ny; g->apron = apron;
    g->pitch = (int)pitch; g->rows = (int)rows;
    g->min_x = min_x; g->min_y = min_y; g->res = res;
    const size_t bytes = (size_t)pitch * rows * sizeof(double);
    // stream-ordered pool allocation: lgs_grid_resize can then swap buffers without a device sync
    cudaError_t e = cudaMallocAsync(&g->d, std::max<size_t>(bytes, 8), c->stream);
    if (e != cudaSuccess) {
        delete g;
        return lgs_fail(c, LGS_ERR_NOMEM, "grid_create: cudaMallocAsync(%zu) -> %s", bytes,
                        cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(g->d, 0, bytes, c->stream);
    if (e != cudaSuccess) {
        cudaFree(g->d); delete g;
        return lgs_fail(c, LGS_ERR_CUDA, "grid_create: memset -> %s", cudaGetErrorString(e));
    }
    *out = g;
    return LGS_OK;
}

int lgs_grid_destroy(lgs_grid* g) {
    if (!g) return LGS_OK;
    cudaSetDevice(g->ctx->device);
    if (g->d && g->owns) {
        cudaDeviceSynchronize();            // readers on other contexts' streams (like cudaFree did)
        cudaFreeAsync(g->d, g->ctx->stream);
    }
    delete g;
    return LGS_OK;
}

int lgs_grid_upload(lgs_grid* g, const double* dense) {
    if (!g || !dense) return LGS_ERR_INVALID;
    lgs_ctx* c = g->ctx;
    if (g->nx == 0 || g->ny == 0) return LGS_OK;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaMemcpy2DAsync(g->origin(), (size_t)g->pitch * sizeof(double), dense,
                                  (size_t)g->nx * sizeof(double), (size_t)g->nx * sizeof(double),
                                  g->ny, cudaMemcpyHostToDevice, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    return LGS_OK;
}

int lgs_grid_download(const lgs_grid* g, double* dense) {
    if (!g || !dense) return LGS_ERR_INVALID;
    lgs_ctx* c = g->ctx;
    if (g->nx == 0 || g->ny == 0) return LGS_OK;
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaMemcpy2DAsync(dense, (size_t)g->nx * sizeof(double), g->origin(),
                                  (size_t)g->pitch * sizeof(double),
                                  (size_t)g->nx * sizeof(double), g->ny, cudaMemcpyDeviceToHost,
                                  c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    return LGS_OK;
}

int lgs_grid_download_region(const lgs_grid* g, int x0, int y0, int w, int h, double* dst, long long dstPitch) {
    if (!g || !dst) return LGS_ERR_INVALID;
    lgs_ctx* c = g->ctx;
    if (w == 0 || h == 0) return LGS_OK;
    if (x0 < 0 || y0 < 0 || w < 0 || h < 0 || x0 + (long long)w > g->nx || y0 + (long long)h > g->ny || dstPitch < w)
        return lgs_fail(c, LGS_ERR_INVALID, "grid_download_region: [%d, %d) x [%d, %d) of a %dx%d grid, pitch %lld",
                        x0, x0 + w, y0, y0 + h, g->nx, g->ny, dstPitch);
    LGS_CUDA(c, cudaSetDevice(c->device));
    LGS_CUDA(c, cudaMemcpy2DAsync(dst, (size_t)dstPitch * sizeof(double), g->origin() + (size_t)y0 * g->pitch + x0,
                                  (size_t)g->pitch * sizeof(double), (size_t)w * sizeof(double), h,
                                  cudaMemcpyDeviceToHost, c->stream));
    LGS_CUDA(c, cudaStreamSynchronize(c->stream));
    return LGS_OK;
}

int lgs_grid_copy(const lgs_grid* src, lgs_grid* dst) {
    if (!src || !dst || src == dst) return LGS_ERR_INVALID;
    lgs_ctx* c = dst->ctx;
    if (src->ctx->device != c->device)
        return lgs_fail(c, LGS_ERR_INVALID, "grid_copy: grids live on devices %d and %d", src->ctx->device, c->device);
    if (src->res != dst->res)
        return lgs_fail(c, LGS_ERR_INVALID, "grid_copy: resolutions %g and %g differ", src->res, dst->res);
    LGS_CUDA(c, cudaSetDevice(c->device));
    if (dst->nx != src->nx || dst->ny != src->ny) {
        // new placement, everything unknown (shifts beyond any map size)
        const int rc = lgs_grid_resize(dst, src->nx, src->ny, src->min_x, src->min_y, 1 << 30, 1 << 30);
        if (rc != LGS_OK) return rc;
    }
    dst->min_x = src->min_x; dst->min_y = src->min_y;
    dst->off_x = src->off_x; dst->off_y = src->off_y;
    // everything queued on the source's stream (its last integration) must have landed
    if (src->ctx != c) LGS_CUDA(c, cudaStreamSynchronize(src->ctx->stream));
    if (src->nx > 0 && src->ny > 0)
        LGS_CUDA(c, cudaMemcpy2DAsync(dst->origin(), (size_t)dst->pitch * sizeof(double), src->origin(),
                                      (size_t)src->pitch * sizeof(double), (size_t)src->nx * sizeof(double),
                                      src->ny, cudaMemcpyDeviceToDevice, c->stream));
    return LGS_OK;
}

int lgs_grid_set_window(lgs_grid* g, int off_x, int off_y) {
    if (!g) return LGS_ERR_INVALID;
    if (off_x < 0 || off_y < 0) return lgs_fail(g->ctx, LGS_ERR_INVALID, "grid_set_window: negative offset");
    g->off_x = off_x; g->off_y = off_y;
    return LGS_OK;
}

int lgs_grid_info(const lgs_grid* g, int* nx, int* ny, double* min_x, double* min_y, double* res,
                  int* apron) {
    if (!g) return LGS_ERR_INVALID;
    if (nx) *nx = g->nx;
    if (ny) *ny = g->ny;
    if (min_x) *min_x = g->min_x;
    if (min_y) *min_y = g->min_y;
    if (res) *res = g->res;
    if (apron) *apron = g->apron;
    return LGS_OK;
}

}  // extern "C"
