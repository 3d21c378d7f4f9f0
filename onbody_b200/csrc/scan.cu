/*
 * scan.cu - exclusive prefix sum over uint32 counts (interaction-list offsets). Three small kernels:
 * per-block scan + block totals, recursive scan of the totals, add-back. HBM-bound, trivial share of a step.
 */
#include "onb_internal.h"

namespace {
constexpr int SCAN_T = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_T * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_T) k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ tile_sums, uint32_t n) {
    __shared__ uint32_t s_warp[SCAN_T / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS]; uint32_t sum = 0;
    #pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { const uint32_t i = base + k; v[k] = i < n ? in[i] : 0u; sum += v[k]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = sum;
    #pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, total = 0;
    #pragma unroll
    for (int w = 0; w < SCAN_T / 32; ++w) { const uint32_t t = s_warp[w]; if (w < warp) wbase += t; total += t; }
    uint32_t run = wbase + inc - sum;
    #pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { const uint32_t i = base + k; if (i < n) out[i] = run; run += v[k]; }
    if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = total;
}
__global__ void k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_off, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_off[i / SCAN_TILE];
}
__global__ void k_total(const uint32_t* __restrict__ in, const uint32_t* __restrict__ out, uint32_t n, unsigned long long* total) {
    *total = (unsigned long long)out[n - 1] + (unsigned long long)in[n - 1];
}
}  // namespace

// out[i] = sum of in[0..i); if total != null it receives the grand total (host value, synchronises the stream).
// in and out may alias.
int onb_exclusive_scan_u32(onb_context* c, const uint32_t* in, uint32_t* out, uint32_t n, uint64_t* total) {
    if (n == 0) { if (total) *total = 0; return ONB_OK; }
    unsigned long long* d_total = nullptr;
    uint32_t last_in = 0;
    if (total) ONB_CUDA(cudaMemcpyAsync(&last_in, in + (n - 1), 4, cudaMemcpyDeviceToHost, ONB_ST(c)));
    (void)d_total;
    const uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t* sums = nullptr;
    if (ntiles > 1) ONB_CUDA(onb_dmalloc(c, (void**)&sums, (size_t)ntiles * 4));
    k_scan_tiles<<<ntiles, SCAN_T, 0, ONB_ST(c)>>>(in, out, sums, n); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    if (ntiles > 1) {
        int rc = onb_exclusive_scan_u32(c, sums, sums, ntiles, nullptr);
        if (rc) { onb_dfree(c, sums); return rc; }
        k_scan_add<<<(n + 255) / 256, 256, 0, ONB_ST(c)>>>(out, sums, n); ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
    }
    if (total) {
        uint32_t last_out = 0;
        ONB_CUDA(cudaMemcpyAsync(&last_out, out + (n - 1), 4, cudaMemcpyDeviceToHost, ONB_ST(c)));
        ONB_CUDA(cudaStreamSynchronize(ONB_ST(c)));
        *total = (uint64_t)last_out + (uint64_t)last_in;
    }
    (void)sums;      // arena scratch: released at the next phase
    return ONB_OK;
}
