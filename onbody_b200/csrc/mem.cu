/*
 * mem.cu - device memory of a context: the scratch arena, persistent planes, and SPARSE planes.
 *
 * Sparse planes (lean memory mode, N = 1e9 on 8 GPUs): a plane keeps its full virtual extent - so every kernel indexes it
 * exactly as on one GPU - but physical HBM is mapped only under the index ranges this rank touches (its shard of the
 * target outputs; per tree level, the equivalent-target blocks of the nodes that overlap its shard). Built on the CUDA
 * virtual-memory-management driver API (cuMemAddressReserve / cuMemCreate / cuMemMap), reached through
 * cudaGetDriverEntryPoint because the library links the runtime statically and must not link libcuda.
 */
#include "onb_internal.h"
#include <cuda.h>
#include <algorithm>

namespace {

struct Vmm {
    CUresult (*getGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*addressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    bool ok = false, tried = false;
};
Vmm g_vmm;

bool vmm_load() {
    if (g_vmm.tried) return g_vmm.ok;
    g_vmm.tried = true;
    struct { const char* name; void** fn; } syms[] = {
        {"cuMemGetAllocationGranularity", (void**)&g_vmm.getGranularity}, {"cuMemAddressReserve", (void**)&g_vmm.addressReserve},
        {"cuMemAddressFree", (void**)&g_vmm.addressFree}, {"cuMemCreate", (void**)&g_vmm.create}, {"cuMemRelease", (void**)&g_vmm.release},
        {"cuMemMap", (void**)&g_vmm.map}, {"cuMemUnmap", (void**)&g_vmm.unmap}, {"cuMemSetAccess", (void**)&g_vmm.setAccess} };
    for (auto& s : syms) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint(s.name, s.fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*s.fn) { cudaGetLastError(); return false; }
    }
    g_vmm.ok = true;
    return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// scratch arena
// ---------------------------------------------------------------------------------------------
cudaError_t onb_dmalloc(onb_context* c, void** p, size_t bytes) {
    bytes = (std::max<size_t>(bytes, 4) + 255) & ~(size_t)255;
    while (c->slab_cur < c->slabs.size()) {
        onb_context::Slab& s = c->slabs[c->slab_cur];
        if (c->slab_off + bytes <= s.cap) { *p = s.p + c->slab_off; c->slab_off += bytes; return cudaSuccess; }
        ++c->slab_cur; c->slab_off = 0;
    }
    onb_context::Slab s; s.cap = std::max<size_t>(bytes, (size_t)64 << 20); s.p = nullptr;
    cudaSetDevice(c->device);
    cudaError_t e = cudaMalloc((void**)&s.p, s.cap);
    if (e != cudaSuccess) return e;
    c->slabs.push_back(s); c->slab_cur = c->slabs.size() - 1; c->slab_off = bytes;
    *p = s.p;
    return cudaSuccess;
}

void onb_scratch_reset(onb_context* c) {
    cudaSetDevice(c->device);         // every public phase call starts here, possibly on a fresh host thread (the -g drivers): the
                                      // current device is per-thread state and all allocations below must land on OUR device
    c->cur_stream = nullptr;          // (an error return may have left a phase's secondary stream selected)
    onb_join_copies(c);
    if (c->slabs.size() > 1) {          // coalesce what the last call needed into one slab
        cudaStreamSynchronize(c->stream);
        size_t total = 0;
        for (auto& s : c->slabs) { total += s.cap; cudaFree(s.p); }
        c->slabs.clear();
        onb_context::Slab s; s.cap = total; s.p = nullptr;
        if (cudaMalloc((void**)&s.p, s.cap) == cudaSuccess) c->slabs.push_back(s);
    }
    c->slab_cur = 0; c->slab_off = 0;
}

// lean memory mode: hand the arena back to the driver (the big users are the tree builds: ~40 B per particle)
void onb_scratch_trim(onb_context* c) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    for (auto& s : c->slabs) cudaFree(s.p);
    c->slabs.clear(); c->slab_cur = 0; c->slab_off = 0;
}

// ---------------------------------------------------------------------------------------------
// sparse planes
// ---------------------------------------------------------------------------------------------
// Reserves `total_bytes` of virtual address space and maps physical memory under the given byte ranges (rounded out to the
// allocation granularity, overlapping ranges merged). The mapped memory is zero-filled on `st`.
cudaError_t onb_sparse_alloc(onb_context* c, void** p, size_t total_bytes, const std::vector<std::pair<size_t, size_t>>& ranges, cudaStream_t st) {
    *p = nullptr;
    if (!vmm_load()) return cudaErrorNotSupported;
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = c->device;
    size_t gran = 0;
    if (g_vmm.getGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return cudaErrorNotSupported;
    const size_t va = ((std::max<size_t>(total_bytes, 1) + gran - 1) / gran) * gran;
    std::vector<std::pair<size_t, size_t>> rs;      // [begin, end) in bytes, granularity aligned, merged
    for (auto& r : ranges) {
        if (r.second == 0) continue;
        const size_t b = (r.first / gran) * gran, e = std::min(va, ((r.first + r.second + gran - 1) / gran) * gran);
        if (e > b) rs.push_back({b, e});
    }
    std::sort(rs.begin(), rs.end());
    std::vector<std::pair<size_t, size_t>> merged;
    for (auto& r : rs) { if (!merged.empty() && r.first <= merged.back().second) merged.back().second = std::max(merged.back().second, r.second); else merged.push_back(r); }
    onb_context::Sparse sp;
    CUdeviceptr base = 0;
    if (g_vmm.addressReserve(&base, va, 0, 0, 0) != CUDA_SUCCESS) return cudaErrorMemoryAllocation;
    sp.base = (char*)base; sp.va = va;
    CUmemAccessDesc acc = {};
    acc.location = prop.location; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    bool fail = false;
    for (auto& r : merged) {
        CUmemGenericAllocationHandle h = 0;
        const size_t sz = r.second - r.first;
        if (g_vmm.create(&h, sz, &prop, 0) != CUDA_SUCCESS) { fail = true; break; }
        if (g_vmm.map(base + r.first, sz, 0, h, 0) != CUDA_SUCCESS) { g_vmm.release(h); fail = true; break; }
        sp.off.push_back(r.first); sp.len.push_back(sz); sp.handle.push_back((unsigned long long)h);
        if (g_vmm.setAccess(base + r.first, sz, &acc, 1) != CUDA_SUCCESS) { fail = true; break; }
    }
    if (fail) {
        for (size_t k = 0; k < sp.off.size(); ++k) { g_vmm.unmap(base + sp.off[k], sp.len[k]); g_vmm.release((CUmemGenericAllocationHandle)sp.handle[k]); }
        g_vmm.addressFree(base, va);
        return cudaErrorMemoryAllocation;
    }
    for (size_t k = 0; k < sp.off.size(); ++k) {
        cudaError_t e = cudaMemsetAsync(sp.base + sp.off[k], 0, sp.len[k], st);
        if (e != cudaSuccess) return e;
    }
    c->sparse[(void*)sp.base] = sp;
    *p = (void*)sp.base;
    return cudaSuccess;
}

const onb_context::Sparse* onb_sparse_info(const onb_context* c, const void* p) {
    auto it = c->sparse.find(const_cast<void*>(p));
    return it == c->sparse.end() ? nullptr : &it->second;
}

void onb_pfree(onb_context* c, void* p) {
    if (!p) return;
    auto it = c->sparse.find(p);
    if (it == c->sparse.end()) { cudaFree(p); return; }
    cudaDeviceSynchronize();
    const onb_context::Sparse& sp = it->second;
    for (size_t k = 0; k < sp.off.size(); ++k) { g_vmm.unmap((CUdeviceptr)(sp.base + sp.off[k]), sp.len[k]); g_vmm.release((CUmemGenericAllocationHandle)sp.handle[k]); }
    g_vmm.addressFree((CUdeviceptr)sp.base, sp.va);
    c->sparse.erase(it);
}

// copies [first, first+count) elements of a (possibly sparse) float plane to the host, skipping what is not mapped
cudaError_t onb_copy_plane_to_host(onb_context* c, float* dst, const float* src, size_t first, size_t count, cudaStream_t st) {
    const onb_context::Sparse* sp = onb_sparse_info(c, src);
    if (!sp) return cudaMemcpyAsync(dst + first, src + first, count * sizeof(float), cudaMemcpyDefault, st);
    const size_t b0 = first * sizeof(float), b1 = (first + count) * sizeof(float);
    for (size_t k = 0; k < sp->off.size(); ++k) {
        const size_t a = std::max(b0, sp->off[k]), b = std::min(b1, sp->off[k] + sp->len[k]);
        if (b <= a) continue;
        cudaError_t e = cudaMemcpyAsync((char*)dst + a, (const char*)src + a, b - a, cudaMemcpyDefault, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// zero a (possibly sparse) float plane
int onb_memset_plane(onb_context* c, float* p, size_t count, cudaStream_t st) {
    const onb_context::Sparse* sp = onb_sparse_info(c, p);
    if (!sp) { ONB_CUDA(cudaMemsetAsync(p, 0, count * sizeof(float), st)); return ONB_OK; }
    for (size_t k = 0; k < sp->off.size(); ++k) ONB_CUDA(cudaMemsetAsync(sp->base + sp->off[k], 0, sp->len[k], st));
    return ONB_OK;
}

// events are created once per context and reused (creating and destroying them in the hot path costs, and leaks on early returns)
cudaEvent_t onb_cached_event(onb_context* c, size_t i) {
    if (c->ev_cache.size() <= i) c->ev_cache.resize(i + 1, nullptr);
    if (!c->ev_cache[i]) cudaEventCreate(&c->ev_cache[i]);
    return c->ev_cache[i];
}
