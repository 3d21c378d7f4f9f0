/*
 * scan.cu - exclusive prefix sum over uint32 counts (interaction-list offsets). Three small kernels:
 * per-block scan + block totals, recursive scan of the totals, add-back. HBM-bound, trivial share of a step.
 */
#include "onb_internal.h"
#include <algorithm>

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
// ---- longest-list-first order of the work items of a pair-kernel launch --------------------------------------------
// A launch ends when its longest list has been walked; with items in tree order the long lists are scattered and the tail of
// every launch is a few CTAs finishing alone. A counting sort of the items by list length (256 bins of 4 entries, descending;
// exact order inside a bin does not matter) lets the hardware scheduler hand out the long lists first. Per-item results do
// not depend on which CTA runs when, so nothing changes but the time.
constexpr int LPT_BINS = 256;
__device__ __forceinline__ uint32_t lpt_bin(uint32_t len) { return (LPT_BINS - 1) - min((uint32_t)(LPT_BINS - 1), (len + 3u) >> 2); }   // len 0 -> last bin
__global__ void __launch_bounds__(256) k_lpt_hist(const uint32_t* __restrict__ start, uint32_t n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[LPT_BINS];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(&sh[lpt_bin(start[i + 1] - start[i])], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void __launch_bounds__(LPT_BINS) k_lpt_offsets(uint32_t* hist) {      // exclusive scan of the 256 bins, in place
    __shared__ uint32_t sh[LPT_BINS];
    const uint32_t v = hist[threadIdx.x];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < LPT_BINS; o <<= 1) { const uint32_t t = threadIdx.x >= (unsigned)o ? sh[threadIdx.x - o] : 0u; __syncthreads(); sh[threadIdx.x] += t; __syncthreads(); }
    hist[threadIdx.x] = sh[threadIdx.x] - v;
}
__global__ void __launch_bounds__(256) k_lpt_scatter(const uint32_t* __restrict__ start, uint32_t n, uint32_t* __restrict__ offs, uint32_t* __restrict__ order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[atomicAdd(&offs[lpt_bin(start[i + 1] - start[i])], 1u)] = i;
}
}  // namespace

// order[k] = k-th work item in (binned) descending list length; start has n+1 entries. Enqueued on ONB_ST(c), no host sync.
int onb_lpt_order(onb_context* c, const uint32_t* start, uint32_t n, uint32_t** order_out) {
    *order_out = nullptr;
    if (n < 2) return ONB_OK;
    uint32_t *hist = nullptr, *order = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&hist, LPT_BINS * 4));
    ONB_CUDA(onb_dmalloc(c, (void**)&order, (size_t)n * 4));
    ONB_CUDA(cudaMemsetAsync(hist, 0, LPT_BINS * 4, ONB_ST(c)));
    const uint32_t blocks = std::min<uint32_t>((n + 255) / 256, 1024u);
    k_lpt_hist<<<blocks, 256, 0, ONB_ST(c)>>>(start, n, hist); ONB_LAUNCH(c);
    k_lpt_offsets<<<1, LPT_BINS, 0, ONB_ST(c)>>>(hist); ONB_LAUNCH(c);
    k_lpt_scatter<<<(n + 255) / 256, 256, 0, ONB_ST(c)>>>(start, n, hist, order); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    *order_out = order;
    return ONB_OK;
}

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
