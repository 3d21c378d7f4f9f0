/*
 * onb_internal.h - device data model and launch plumbing shared by the .cu files.
 *
 * Layout in HBM (all SoA, float32, indices uint32 - enough for N < 2^32 per GPU):
 *   particles : x[PD][n], r[n], s[SD][n] (sources), u[OD][n] (targets), gidx[n] (targets: original index)
 *   tree      : implicit binary tree, root 1, children 2i/2i+1, 2^levels slots (reference Tree.hpp:44-76):
 *               x[PD] centre of |s|, nc[PD] bbox centre, ns[PD] bbox size, nr half diagonal, pr mean radius,
 *               s[SD] summed strength, ioffset/num (first particle, count)
 *   equivalent particles of node i live at [i*ebs, i*ebs + numEqps), ebs = 128-padded (order+1)^PD
 *   packed source tiles for the pair kernels: float4 planes built once per source set (p2p.cu)
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <map>

#include "../../include/onbody_b200.h"

#define ONB_MAX_PD 3
#define ONB_MAX_SD 3
#define ONB_MAX_OD 12
#define ONB_MAX_ORDER 20

struct DParts {
    uint32_t n = 0;
    uint32_t cap = 0;        // allocated elements per plane (>= n, padded so 16-byte bulk copies may overrun)
    int PD = 3, SD = 1, OD = 3;
    bool are_sources = false;
    float* x[ONB_MAX_PD] = {nullptr, nullptr, nullptr};
    float* r = nullptr;
    float* s[ONB_MAX_SD] = {nullptr, nullptr, nullptr};
    float* u[ONB_MAX_OD] = {nullptr};
    uint32_t* gidx = nullptr;
    uint32_t* gidx_spare = nullptr;   // the index plane of the previous build, kept for reuse (gidx == nullptr means "no tree order yet")
    // packed planes for the pair kernels (see p2p.cu):
    //   grav3d          pk0 = (x,y,z,s)      pk2 = r^2
    //   vort3d/vortgrad pk0 = (x,y,z,r^2)    pk1 = (sx,sy,sz,0)
    //   vort2d/2dtr     pk0 = (x,y,r^2,s)
    float4* pk0 = nullptr;
    float4* pk1 = nullptr;
    float*  pk2 = nullptr;
    bool packed_valid = false;
    uint32_t build_lo = 0, build_hi = 0;     // particle range the last tree build was restricted to
};

struct DTree {
    int levels = 0, numnodes = 0;
    float* x[ONB_MAX_PD] = {nullptr, nullptr, nullptr};
    float* nc[ONB_MAX_PD] = {nullptr, nullptr, nullptr};
    float* ns[ONB_MAX_PD] = {nullptr, nullptr, nullptr};
    float* nr = nullptr;
    float* pr = nullptr;
    float* s[ONB_MAX_SD] = {nullptr, nullptr, nullptr};
    uint32_t* ioffset = nullptr;
    uint32_t* num = nullptr;
    bool built = false;
};

// by-value kernel argument views
struct PartsView {
    uint32_t n;
    float* x[ONB_MAX_PD];
    float* r;
    float* s[ONB_MAX_SD];
    float* u[ONB_MAX_OD];
    uint32_t* gidx;
};
struct TreeView {
    int levels, numnodes;
    float* x[ONB_MAX_PD];
    float* nc[ONB_MAX_PD];
    float* ns[ONB_MAX_PD];
    float* nr;
    float* pr;
    float* s[ONB_MAX_SD];
    uint32_t* ioffset;
    uint32_t* num;
};

struct onb_context {
    int physics = 0, device = 0;
    int PD = 3, SD = 1, OD = 3, flops_per_pair = 19;
    bool has_tr = false, has_fastsumm = true;
    int block = 128, order = 4, arith = ONB_ARITH_FAST;
    int ncp = 5, num_eqps = 125, ebs = 128;
    // legacy equivalents (-o omitted, order = -1, barneshut.hpp:946-1061): per-node counts of pair-merged equivalents
    bool legacy = false;
    uint32_t* d_epnum = nullptr; uint32_t epnum_cap = 0, root_epnum = 0;
    int shard_rank = 0, shard_n = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    DParts parts[4];
    DTree trees[2];
    uint64_t stats[9] = {0};
    uint64_t last_pairs = 0;
    uint64_t launches = 0;
    std::map<std::string, double> phase_ms;
    std::string err;
    // error flag written by kernels (capacity overflow etc.)
    int* d_flag = nullptr;
    int* h_flag = nullptr;   // pinned
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    // scratch arena: grow-only device slabs, bump-allocated inside one API call and reset at the start of the next.
    // After the first (warm-up) call of a phase no allocation reaches the driver any more.
    struct Slab { char* p; size_t cap; };
    std::vector<Slab> slabs;
    size_t slab_cur = 0, slab_off = 0;
    unsigned long long* d_build_stats = nullptr;   // selects, passes, stalls, scanned, tie sorts (x2: second slot for a concurrent build)
    cudaStream_t stream2 = nullptr, cur_stream = nullptr;   // second stream for building both trees concurrently
    int cur_stats_off = 0;
    bool concurrent_builds = false;
    std::vector<uint64_t> dtt_sizes;      // per level: interaction / deferred list sizes of the last dual-tree evaluation
    bool dtt_sizes_valid = false;
    // opt-in asynchronous input copies (onb_set_async_inputs): the target planes arrive on stream2
    bool async_inputs = false, tgt_copy_pending = false;
    cudaEvent_t ev_copy = nullptr, ev_tgt_ready = nullptr;
};
int onb_join_copies(onb_context* c);   // make the context stream wait for a pending asynchronous target copy

#define ONB_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    c->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
    return ONB_ERR_CUDA; } } while (0)

#define ONB_LAUNCH(c) ((c)->launches++)
// the stream / statistics slot of the work being enqueued: onb_make_trees enqueues two builds on two streams, the dual
// tree builds its interaction lists on the second stream ahead of the pair kernels
#define ONB_ST(c) ((c)->cur_stream ? (c)->cur_stream : (c)->stream)
#define ONB_STATS(c) ((c)->d_build_stats + (c)->cur_stats_off)

// scratch (per-call) device memory: bump allocation from the context's arena; "free" is a no-op, the arena is reset by
// onb_scratch_reset() at the start of every public phase call. Persistent arrays use onb_pmalloc / onb_pfree.
cudaError_t onb_dmalloc(onb_context* c, void** p, size_t bytes);
static inline void onb_dfree(onb_context*, void*) {}
void onb_scratch_reset(onb_context* c);
static inline cudaError_t onb_pmalloc(onb_context*, void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 4); }
static inline void onb_pfree(onb_context*, void* p) { if (p) cudaFree(p); }

static inline PartsView view_of(const DParts& p) {
    PartsView v; v.n = p.n;
    for (int d = 0; d < ONB_MAX_PD; ++d) v.x[d] = p.x[d];
    v.r = p.r;
    for (int d = 0; d < ONB_MAX_SD; ++d) v.s[d] = p.s[d];
    for (int d = 0; d < ONB_MAX_OD; ++d) v.u[d] = p.u[d];
    v.gidx = p.gidx;
    return v;
}
static inline TreeView view_of(const DTree& t) {
    TreeView v; v.levels = t.levels; v.numnodes = t.numnodes;
    for (int d = 0; d < ONB_MAX_PD; ++d) { v.x[d] = t.x[d]; v.nc[d] = t.nc[d]; v.ns[d] = t.ns[d]; }
    v.nr = t.nr; v.pr = t.pr;
    for (int d = 0; d < ONB_MAX_SD; ++d) v.s[d] = t.s[d];
    v.ioffset = t.ioffset; v.num = t.num;
    return v;
}

// phase timer: CUDA events on the context stream
struct PhaseTimer {
    onb_context* c; const char* name; cudaEvent_t a, b; bool on;
    PhaseTimer(onb_context* ctx, const char* nm) : c(ctx), name(nm), on(true) {
        cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->stream);
    }
    double stop() {
        if (!on) return 0.0;
        cudaEventRecord(b, c->stream); cudaEventSynchronize(b);
        float ms = 0.f; cudaEventElapsedTime(&ms, a, b);
        cudaEventDestroy(a); cudaEventDestroy(b); on = false;
        c->phase_ms[name] = ms; return ms;
    }
    ~PhaseTimer() { if (on) stop(); }
};

// ---- implemented in the .cu files -------------------------------------------------------------
// memory
int onb_alloc_parts(onb_context* c, DParts& p, uint32_t n, bool are_sources);
void onb_free_parts(onb_context* c, DParts& p);
int onb_alloc_tree(onb_context* c, DTree& t, uint32_t n, int block);
void onb_free_tree(onb_context* c, DTree& t);
int onb_check_flag(onb_context* c, const char* what);
void onb_shard_range(const onb_context* c, uint32_t* lo, uint32_t* hi);   // particle index range of this context's target shard
// tree.cu
int onb_tree_build(onb_context* c, DParts& p, DTree& t, uint32_t blo, uint32_t bhi);
int onb_tree_finish_from_particles(onb_context* c, DParts& p, DTree& t);
int onb_tree_refine(onb_context* c, DParts& p, DTree& t, bool check_now = true);
// bary.cu
int onb_bary_upward(onb_context* c, DParts& p, DParts& ep, DTree& t);
int onb_bary_downward_level(onb_context* c, int level);
int onb_legacy_equivalents(onb_context* c, DParts& p, DParts& ep, DTree& t);
// p2p.cu
int onb_pack_sources(onb_context* c, DParts& p);
int onb_p2p_direct(onb_context* c, uint64_t tskip);
// interaction lists are CSR over "target work items": item w covers targets [tgt_off[w], +tgt_cnt[w]) of
// parts[tgt_which], and interacts with entries [start[w], start[w+1]); entry = source node id, bit 31 set = use the
// node's equivalent particles, clear = its real particles.
struct WorkList {
    uint32_t nitems = 0;
    uint32_t* tgt_node = nullptr;  // target tree node id per item; null = node_base + item
    uint32_t node_base = 0;
    uint32_t* start = nullptr;     // nitems+1
    uint32_t* entries = nullptr;
    uint64_t nentries = 0;
};
int onb_p2p_lists(onb_context* c, const WorkList& wl, int tgt_which_leaf, int tgt_which_box, bool accumulate, uint32_t nsplit = 1);
void onb_free_worklist(onb_context* c, WorkList& wl);
// traverse.cu
int onb_lists_boxwise(onb_context* c, float theta, WorkList& wl);
int onb_run_treecode2(onb_context* c, float theta, int variant);   // fused pointwise traversal + pair kernels (variant 1 = treecode1)
int onb_run_fastsumm(onb_context* c, float theta);
// scan.cu
int onb_exclusive_scan_u32(onb_context* c, const uint32_t* in, uint32_t* out, uint32_t n, uint64_t* total);
