// lgs_peaks.cu -- measured gather ceilings for the scoring kernels (diagnostic, used by bench.py).
//
// the correlative / branch-and-bound scoring kernels are bound by 8-byte
// gathers from a map that is cache resident by design, so the HBM copy peak is not their roofline.
// MEASURED_PEAKS.json has no L1/L2 figure; this micro-benchmark measures one with the access width
// and shapes of the sweep: every warp repeatedly loads a ROW of consecutive doubles at a data
// dependent position of a map-sized array and adds it to a per-lane accumulator.
//   rows = 32, aligned to 256 B : the best a 32-lane 8-byte gather can do (2 L1 wavefronts)
//   rows = 25, any 8-byte offset: the sweep's shape (one window row per request, 2-3 lines)
//   locality = 1: successive rows move by a few cells (neighbouring beams) -> L1 hits
//   locality = 0: successive rows are uniformly random over the array      -> L2 hits
#include "lgs_internal.cuh"

namespace {

template <int LANES>
__global__ void __launch_bounds__(256)
gather_peak_kernel(const double* __restrict__ a, unsigned nCells, int pitch, int iters, int aligned,
                   int local, double* __restrict__ sink) {
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned state = warp * 2654435761u + 12345u;
    unsigned pos = (state >> 4) % (nCells - 8u * (unsigned)pitch - 64u);
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    const bool on = lane < LANES;
    for (int it = 0; it < iters; ++it) {
        state = state * 1664525u + 1013904223u;
        if (local) {                         // next beam: a few cells along the wall, maybe a row up / down
            pos += ((state >> 8) & 7u) + (((state >> 12) & 3u) == 0u ? (unsigned)pitch : 0u);
            if (pos >= nCells - 8u * (unsigned)pitch - 64u) pos = (state >> 4) % (nCells - 8u * (unsigned)pitch - 64u);
        } else {
            pos = (state >> 4) % (nCells - 8u * (unsigned)pitch - 64u);
        }
        const unsigned p = aligned ? (pos & ~31u) : pos;
        if (on) {                            // eight window rows per position (the sweep: 4 hypotheses / thread, unrolled)
            const double* q = a + p + lane;
            const double v0 = __ldg(q), v1 = __ldg(q + pitch), v2 = __ldg(q + 2 * pitch), v3 = __ldg(q + 3 * pitch);
            const double v4 = __ldg(q + 4 * pitch), v5 = __ldg(q + 5 * pitch), v6 = __ldg(q + 6 * pitch), v7 = __ldg(q + 7 * pitch);
            acc0 += v0; acc1 += v1; acc2 += v2; acc3 += v3;
            acc0 += v4; acc1 += v5; acc2 += v6; acc3 += v7;
        }
    }
    if (on && acc0 + acc1 + acc2 + acc3 == 123.456) sink[0] = acc0;   // keeps the loads alive
}

}  // namespace

extern "C" int lgs_measure_gather_peak(lgs_ctx* c, int nx, int ny, int rowLanes, int aligned, int local,
                                       double* gbps) {
    if (!c || !gbps || nx < 64 || ny < 16 || (rowLanes != 25 && rowLanes != 32))
        return c ? lgs_fail(c, LGS_ERR_INVALID, "measure_gather_peak: bad arguments") : LGS_ERR_INVALID;
    LGS_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)nx * ny;
    double* a = nullptr;
    LGS_CUDA(c, cudaMalloc(&a, (n + 1) * sizeof(double)));
    LGS_CUDA(c, cudaMemsetAsync(a, 0, (n + 1) * sizeof(double), c->stream));
    const int iters = 2048, blocks = c->sm_count * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {      // rep 0 warms the caches
        LGS_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        if (rowLanes == 32)
            gather_peak_kernel<32><<<blocks, 256, 0, c->stream>>>(a, (unsigned)n, nx, iters, aligned, local, a + n);
        else
            gather_peak_kernel<25><<<blocks, 256, 0, c->stream>>>(a, (unsigned)n, nx, iters, aligned, local, a + n);
        c->launches++;
        LGS_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        LGS_CUDA(c, cudaEventSynchronize(c->ev1));
        float ms = 0;
        LGS_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0) best = std::min(best, ms);
    }
    cudaFree(a);
    const double bytes = (double)blocks * 8 /* warps */ * iters * 8.0 * rowLanes * 8.0;
    *gbps = bytes / (best * 1e-3) / 1e9;
    return LGS_OK;
}
