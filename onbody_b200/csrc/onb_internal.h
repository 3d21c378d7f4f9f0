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
#define ONB_MAX_RANKS 16

// multi-GPU partition derived from the tree shape alone (plan.cu)
struct ShardPlan {
    bool valid = false;
    uint64_t n = 0, lo = 0, hi = 0, chunk = 0;      // particles; this rank's range; particles per rank (leaf aligned, equal for all ranks)
    int block = 0, nranks = 1, rank = 0, levels = 0;
    std::vector<uint32_t> own_lo, own_hi;           // per level: node ids [own_lo, own_hi) completely inside this rank's range
    std::vector<uint32_t> need_lo, need_hi;         // per level: node ids overlapping this rank's range
    std::vector<uint32_t> all_own_lo, all_own_hi;   // [level * nranks + r]: the same for every rank
    std::vector<std::vector<uint32_t>> shared;      // per level: non-leaf nodes straddling a rank boundary
};
int onb_plan_levels(uint64_t n, int block);
uint64_t onb_shard_chunk(uint64_t n, int block, int nranks);
int onb_plan_make(ShardPlan& P, uint64_t n, int block, int nranks, int rank);
struct OnbComm;

struct DParts {
    uint32_t n = 0;
    uint32_t cap = 0;        // allocated elements per plane (>= n, padded so 16-byte bulk copies may overrun)
    int PD = 3, SD = 1, OD = 3;
    bool are_sources = false;
    float* x[ONB_MAX_PD] = {nullptr, nullptr, nullptr};
    float* r = nullptr;
    float* s[ONB_MAX_SD] = {nullptr, nullptr, nullptr};
    float* u[ONB_MAX_OD] = {nullptr};
    double* ud[ONB_MAX_OD] = {nullptr};   // ACCUM = double (onb_set_accum): the outputs are accumulated and kept in fp64, u is their rounded copy
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
    uint64_t sparse_key = 0;                 // != 0: the output planes (and, for equivalent targets, all planes) are sparse, mapped for this partition
    bool unpacked_released = false;          // lean memory mode: x/r/s were freed after packing (sources only)
    uint32_t u_lo = 0, u_hi = 0;             // element range of the output planes that is backed by memory (sparse planes: the shard)
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
    double* ud[ONB_MAX_OD];
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
    bool accum64 = false;   // the reference's ACCUM = double (ongrav3d.cpp:8): fp32 pair arithmetic, fp64 accumulation and outputs
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
    // dual-tree lists: one persistent pool, bump-allocated per level on the device (traverse.cu); grown when a pass overflows
    uint32_t* dtt_pool = nullptr; uint32_t dtt_pool_cap = 0; uint64_t dtt_pool_want = 0;
    uint32_t dtt_qcap = 2048;             // per-warp FIFO of opened source nodes, x8 when a pass overflows it
    std::vector<cudaEvent_t> ev_cache;    // timing / ordering events, created once
    float* h_stage = nullptr; size_t h_stage_cap = 0; cudaEvent_t ev_stage[2] = {nullptr, nullptr};   // pinned double buffer of the result read-back
    // sparse planes (mem.cu): full virtual extent, physical memory only under the ranges this rank touches
    struct Sparse { char* base = nullptr; size_t va = 0; std::vector<size_t> off, len; std::vector<unsigned long long> handle; };
    std::map<void*, Sparse> sparse;
    int mem_mode = 0;                     // ONB_MEM_NORMAL / ONB_MEM_LEAN
    // multi-GPU (comm.cu, dist.cu): communicator, its stream, the partition plans of the two trees
    OnbComm* comm = nullptr;
    ShardPlan plan[2];
    uint32_t* d_shared[2] = {nullptr, nullptr};   // device copy of plan[which].shared, [levels * nranks]
    uint64_t shared_key[2] = {0, 0};
    const uint32_t* d_eqtab[2] = {nullptr, nullptr}; uint32_t eq_chunk_blocks[2] = {0, 0};   // packed exchange of the equivalent strengths (dist.cu)
    float* eq_stage = nullptr; size_t eq_stage_cap = 0;
    // identifies the partition a sparse allocation / an uploaded plan belongs to (never 0)
    uint64_t plan_key(int which) const { return plan_key_for(parts[which].n); }
    uint64_t plan_key_for(uint64_t n) const { return ((n * 131u + (uint64_t)block) * 131u + (uint64_t)shard_n) * 131u + (uint64_t)shard_rank + 1u; }
    bool ag_timed = false, eq_timed = false;
    float* rec_buf[2] = {nullptr, nullptr}; size_t rec_cap[2] = {0, 0};   // leaf records of the two trees (dist.cu)
    cudaEvent_t ev_src_planes = nullptr; bool src_planes_pending = false;   // the source-plane all-gather may outlive onb_make_trees
    // opt-in asynchronous input copies (onb_set_async_inputs): the target planes arrive on stream2
    bool async_inputs = false, tgt_copy_pending = false, sliced_inputs = false;
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
void onb_scratch_trim(onb_context* c);
static inline cudaError_t onb_pmalloc(onb_context*, void** p, size_t bytes) { return cudaMalloc(p, bytes ? bytes : 4); }
void onb_pfree(onb_context* c, void* p);
cudaError_t onb_sparse_alloc(onb_context* c, void** p, size_t total_bytes, const std::vector<std::pair<size_t, size_t>>& ranges, cudaStream_t st);
const onb_context::Sparse* onb_sparse_info(const onb_context* c, const void* p);
cudaError_t onb_copy_plane_to_host(onb_context* c, float* dst, const float* src, size_t first, size_t count, cudaStream_t st);

static inline PartsView view_of(const DParts& p) {
    PartsView v; v.n = p.n;
    for (int d = 0; d < ONB_MAX_PD; ++d) v.x[d] = p.x[d];
    v.r = p.r;
    for (int d = 0; d < ONB_MAX_SD; ++d) v.s[d] = p.s[d];
    for (int d = 0; d < ONB_MAX_OD; ++d) { v.u[d] = p.u[d]; v.ud[d] = p.ud[d]; }
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
int onb_tree_build(onb_context* c, DParts& p, DTree& t, uint32_t blo, uint32_t bhi, bool finish = true);
int onb_leafrec_floats(const onb_context* c, bool are_sources);
int onb_tree_leaf_records(onb_context* c, DParts& p, DTree& t, uint32_t leaf0, uint32_t leaf1, float* rec);
int onb_tree_finish_from_records(onb_context* c, DParts& p, DTree& t, const float* rec);
int onb_tree_finish_from_particles(onb_context* c, DParts& p, DTree& t);
int onb_tree_refine(onb_context* c, DParts& p, DTree& t, bool check_now = true);
// bary.cu
enum { ONB_UP_ALL = 0, ONB_UP_OWN = 1, ONB_UP_SHARED = 2, ONB_UP_POS = 3, ONB_UP_NEED = 4 };
int onb_bary_upward_mode(onb_context* c, DParts& p, DParts& ep, DTree& t, int mode);
int onb_bary_upward(onb_context* c, DParts& p, DParts& ep, DTree& t);
int onb_bary_downward_level(onb_context* c, int level);
int onb_legacy_equivalents(onb_context* c, DParts& p, DParts& ep, DTree& t);
// p2p.cu
int onb_pack_sources(onb_context* c, DParts& p);
int onb_p2p_direct(onb_context* c, uint64_t tskip);
int onb_round_outputs(onb_context* c, DParts& p);      // ACCUM = double: u = (float) ud
// interaction lists are CSR over "target work items": item w covers targets [tgt_off[w], +tgt_cnt[w]) of
// parts[tgt_which], and interacts with entries [start[w], start[w+1]); entry = source node id, bit 31 set = use the
// node's equivalent particles, clear = its real particles.
struct WorkList {
    uint32_t nitems = 0;
    uint32_t* tgt_node = nullptr;  // target tree node id per item; null = node_base + item
    uint32_t node_base = 0;
    uint32_t* start = nullptr;     // nitems+1
    uint32_t* entries = nullptr;
    uint64_t nentries = 0;         // allocated entries (reads are clamped to it)
    const uint32_t* ebase = nullptr;   // optional device-resident offset of the list inside `entries`
    const uint32_t* order = nullptr;   // optional launch order of the items (longest list first, scan.cu onb_lpt_order)
};
int onb_p2p_lists(onb_context* c, const WorkList& wl, int tgt_which_leaf, int tgt_which_box, bool accumulate, uint32_t nsplit = 1);
void onb_free_worklist(onb_context* c, WorkList& wl);
// traverse.cu
int onb_lists_boxwise(onb_context* c, float theta, WorkList& wl);
int onb_run_treecode2(onb_context* c, float theta, int variant);   // fused pointwise traversal + pair kernels (variant 1 = treecode1)
int onb_run_fastsumm(onb_context* c, float theta);
void onb_level_span(onb_context* c, int level, uint32_t* node0, uint32_t* count);
// dist.cu / comm.cu
int onb_plan_upload_shared(onb_context* c, int which);
bool onb_dist_sequential_builds(const onb_context* c);
int onb_dist_make_trees(onb_context* c, int which);
int onb_dist_join_source_planes(onb_context* c, cudaStream_t st);
int onb_dist_upward_sources(onb_context* c);
void onb_dist_record_exchange_times(onb_context* c);
cudaEvent_t onb_cached_event(onb_context* c, size_t i);
int onb_memset_plane(onb_context* c, float* p, size_t count, cudaStream_t st);
// scan.cu
int onb_exclusive_scan_u32(onb_context* c, const uint32_t* in, uint32_t* out, uint32_t n, uint64_t* total);
int onb_lpt_order(onb_context* c, const uint32_t* start, uint32_t n, uint32_t** order_out);
