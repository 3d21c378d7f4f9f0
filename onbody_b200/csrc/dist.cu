/*
 * dist.cu - the phases of the hot path when a communicator is attached (comm.cu): one context per GPU, every context holds
 * the same inputs, the TARGETS are sharded by contiguous leaf ranges and the SOURCES end up replicated (north_star).
 *
 * What crosses NVLink, per step (N particles, R ranks, L tree levels):
 *   leaf records   13 floats per leaf of both trees, one in-place all-gather each (0.4 B per particle)   tree.cu k_leafrec_*
 *   source planes  x[PD], r, s[SD]: each rank's tree-ordered range, one grouped in-place all-gather       (20 B per particle)
 *   eq. strengths  s[SD] of the equivalent sources every rank anterpolated for the nodes inside its range, packed into one
 *                  chunk per rank, one in-place all-gather, scattered back                                (4 B per particle)
 * Target particles are NOT exchanged: a rank sorts, refines and evaluates only its own leaves, and the node arrays of the
 * whole target tree (the ancestors' centres enter the dual-tree MAC, ongrav3d.cpp:338) follow from the leaf records alone.
 *
 * Everything a rank does on its own range needs no communication, so the collectives overlap with it:
 *   stream  : source range build -> own leaf records ............ node arrays -> upward pass of the own nodes .. positions, pack -> straddling nodes -> pack
 *   stream2 : target range build -> own leaf records ............ node arrays -> in-leaf refinement -> equivalent target points of the needed nodes
 *   comm    :                      gather(src rec) gather(tgt rec) gather(source planes) ...................... bcast(eq. strengths)
 * Results are bit-identical to the single-GPU run for every rank count (tools/check_multi.py, tests/test_gpu_dist.py):
 * per-leaf sums, per-node anterpolation and per-target accumulation orders do not depend on who computes them.
 */
#include "onb_internal.h"
#include <algorithm>
#include <cstdlib>

int onb_comm_allgather(onb_context* c, const std::vector<void*>& bufs, const std::vector<size_t>& chunk_bytes);
cudaStream_t onb_comm_stream(const onb_context* c);

static int ensure_plan(onb_context* c, int which) {
    const int rc = onb_plan_make(c->plan[which], c->parts[which].n, c->block, c->shard_n, c->shard_rank);
    if (rc) c->err = "cannot plan the multi-GPU partition (rank count above ONB_MAX_RANKS, or an irregular tree shape)";
    return rc;
}

int onb_plan_upload_shared(onb_context* c, int which) {
    int rc = ensure_plan(c, which); if (rc) return rc;
    const ShardPlan& P = c->plan[which];
    const uint64_t key = c->plan_key(which);
    if (c->d_shared[which] && c->shared_key[which] == key) return ONB_OK;
    if (c->d_shared[which]) { cudaFree(c->d_shared[which]); c->d_shared[which] = nullptr; }
    std::vector<uint32_t> h((size_t)P.levels * P.nranks, 0u);
    for (int l = 0; l < P.levels; ++l) {
        if (P.shared[l].size() > (size_t)P.nranks) { c->err = "plan: more straddling nodes than ranks on one level"; return ONB_ERR_UNSUPPORTED; }
        for (size_t k = 0; k < P.shared[l].size(); ++k) h[(size_t)l * P.nranks + k] = P.shared[l][k];
    }
    // ... followed by the table of the strength exchange: for every rank q and level l the first node it owns and the number of
    // blocks it owns on the levels above (tab[(q*(L+1) + l)*2 + {0,1}]); tab[..L..][1] = all blocks of rank q
    const int L = P.levels;
    std::vector<uint32_t> tab((size_t)P.nranks * (L + 1) * 2, 0u);
    uint32_t most = 0;
    for (int q = 0; q < P.nranks; ++q) {
        uint32_t pre = 0;
        for (int l = 0; l <= L; ++l) {
            tab[((size_t)q * (L + 1) + l) * 2 + 1] = pre;
            if (l < L) {
                const uint32_t a = P.all_own_lo[(size_t)l * P.nranks + q], b = P.all_own_hi[(size_t)l * P.nranks + q];
                tab[((size_t)q * (L + 1) + l) * 2] = a;
                if (l + 1 < L && b > a) pre += b - a;          // the last level holds only leaves: no equivalent particles there
            }
        }
        most = std::max(most, pre);
    }
    c->eq_chunk_blocks[which] = most;
    const size_t shared_words = h.size();
    h.insert(h.end(), tab.begin(), tab.end());
    ONB_CUDA(cudaMalloc((void**)&c->d_shared[which], h.size() * 4 + 4));
    ONB_CUDA(cudaMemcpy(c->d_shared[which], h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    c->d_eqtab[which] = c->d_shared[which] + shared_words;
    c->shared_key[which] = key;
    return ONB_OK;
}

// one build after the other: on request, and in the lean memory mode (one build's scratch at a time)
bool onb_dist_sequential_builds(const onb_context* c) {
    static const bool seq_env = std::getenv("ONB_SEQ_BUILDS") != nullptr;
    return seq_env || c->mem_mode == ONB_MEM_LEAN;
}

namespace {

enum { EV_E0 = 400, EV_BUILT0, EV_BUILT1, EV_REC0, EV_REC1, EV_UP, EV_EQ, EV_JOIN, EV_END, EV_AG0, EV_AG1, EV_EQ0 };

struct RecBuf { float* p = nullptr; size_t chunk_bytes = 0; };

// range-restricted build of one tree on ONB_ST(c) (no node summaries) + the records of the rank's own leaves
int dist_build(onb_context* c, int which, RecBuf& rb) {
    const ShardPlan& P = c->plan[which];
    DParts& p = c->parts[which]; DTree& t = c->trees[which];
    const int K = onb_leafrec_floats(c, p.are_sources);
    const size_t leaves_per_rank = (size_t)(P.chunk / (uint64_t)c->block);
    rb.chunk_bytes = leaves_per_rank * K * sizeof(float);
    // persistent (not arena scratch): with sequential builds the arena is rewound between the two builds while the records of
    // the first tree are still waiting to be turned into node arrays
    const size_t need = rb.chunk_bytes * (size_t)P.nranks;
    if (c->rec_cap[which] < need) {
        if (c->rec_buf[which]) cudaFree(c->rec_buf[which]);
        c->rec_buf[which] = nullptr; c->rec_cap[which] = 0;
        ONB_CUDA(cudaMalloc((void**)&c->rec_buf[which], need));
        c->rec_cap[which] = need;
    }
    rb.p = c->rec_buf[which];
    if (P.hi > P.lo) {
        int rc = onb_tree_build(c, p, t, (uint32_t)P.lo, (uint32_t)P.hi, false); if (rc) return rc;
        rc = onb_tree_leaf_records(c, p, t, (uint32_t)(P.lo / c->block), (uint32_t)((P.hi + c->block - 1) / c->block), rb.p); if (rc) return rc;
    } else {
        // a rank without leaves (fewer leaves than ranks) still needs the shape arrays of the tree: a build of nothing
        int rc = onb_tree_build(c, p, t, 0, std::min<uint32_t>(p.n, (uint32_t)c->block), false); if (rc) return rc;
        p.build_lo = p.build_hi = 0;
    }
    return ONB_OK;
}

int source_plane_list(onb_context* c, std::vector<void*>& bufs, std::vector<size_t>& chunks) {
    DParts& p = c->parts[0];
    const size_t cb = (size_t)c->plan[0].chunk * sizeof(float);
    for (int d = 0; d < c->PD; ++d) { bufs.push_back(p.x[d]); chunks.push_back(cb); }
    bufs.push_back(p.r); chunks.push_back(cb);
    for (int d = 0; d < c->SD; ++d) { bufs.push_back(p.s[d]); chunks.push_back(cb); }
    if ((uint64_t)p.cap < c->plan[0].chunk * (uint64_t)c->shard_n) { c->err = "source planes are too short for the in-place all-gather"; return ONB_ERR_ARG; }
    return ONB_OK;
}

}  // namespace

// both trees (which = -1) or one: range builds, leaf-record exchange, node arrays; the all-gather of the source planes is
// enqueued as well and, for which = -1, left in flight (its consumers wait for c->ev_src_planes)
int onb_dist_make_trees(onb_context* c, int which) {
    const bool both = which < 0;
    const bool do_src = both || which == 0, do_tgt = both || which == 1;
    int rc;
    if (c->legacy) {      // refineTree(srcs) + calcEquivalents walk every source leaf in place: there is no sharded form of them here
        c->err = "the legacy equivalents (-o omitted) are single-GPU only: attach the communicator to a context with a barycentric order";
        return ONB_ERR_UNSUPPORTED;
    }
    if (do_src && (rc = ensure_plan(c, 0))) return rc;
    if (do_tgt && (rc = ensure_plan(c, 1))) return rc;
    if (do_src && (rc = onb_alloc_tree(c, c->trees[0], c->parts[0].n, c->block))) return rc;
    if (do_tgt && (rc = onb_alloc_tree(c, c->trees[1], c->parts[1].n, c->block))) return rc;
    const bool seq = !both || onb_dist_sequential_builds(c);
    cudaStream_t s1 = c->stream, s2 = seq ? c->stream : c->stream2, sc = onb_comm_stream(c);
    ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_E0), s1));
    if (s2 != s1) ONB_CUDA(cudaStreamWaitEvent(s2, onb_cached_event(c, EV_E0), 0));
    ONB_CUDA(cudaStreamWaitEvent(sc, onb_cached_event(c, EV_E0), 0));
    RecBuf rb0, rb1;
    c->concurrent_builds = both && !seq;
    if (do_src) {
        rc = dist_build(c, 0, rb0); if (rc) { c->concurrent_builds = false; return rc; }
        ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_BUILT0), s1));
    }
    auto gather_src_rec = [&]() -> int {
        ONB_CUDA(cudaStreamWaitEvent(sc, onb_cached_event(c, EV_BUILT0), 0));
        int r = onb_comm_allgather(c, {rb0.p}, {rb0.chunk_bytes}); if (r) return r;
        ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_REC0), sc));
        return ONB_OK;
    };
    auto gather_src_planes = [&]() -> int {
        std::vector<void*> bufs; std::vector<size_t> chunks;
        int r = source_plane_list(c, bufs, chunks); if (r) return r;
        ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_AG0), sc));
        r = onb_comm_allgather(c, bufs, chunks); if (r) return r;
        ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_AG1), sc));
        c->ag_timed = true;
        ONB_CUDA(cudaEventRecord(c->ev_src_planes, sc));
        c->src_planes_pending = true;
        return ONB_OK;
    };
    if (do_src && seq) {
        if ((rc = gather_src_rec())) return rc; if ((rc = gather_src_planes())) return rc;                          // under the target build
        c->slab_cur = 0; c->slab_off = 0;     // same stream: the target build may reuse the source build's scratch (~40 B per particle)
    }
    if (do_tgt) {
        if (s2 != s1) { c->cur_stream = s2; c->cur_stats_off = 8; }
        rc = dist_build(c, 1, rb1);
        c->cur_stream = nullptr; c->cur_stats_off = 0; c->concurrent_builds = false;
        if (rc) return rc;
        ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_BUILT1), s2));
    }
    c->concurrent_builds = false;
    if (do_src && !seq) { if ((rc = gather_src_rec())) return rc; }
    if (do_tgt) {
        ONB_CUDA(cudaStreamWaitEvent(sc, onb_cached_event(c, EV_BUILT1), 0));
        std::vector<void*> bufs{rb1.p}; std::vector<size_t> chunks{rb1.chunk_bytes};
        if (c->has_tr) { bufs.push_back(c->parts[1].r); chunks.push_back((size_t)c->plan[1].chunk * sizeof(float)); }   // target radii enter the 2-D kernel
        rc = onb_comm_allgather(c, bufs, chunks); if (rc) return rc;
        ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_REC1), sc));
    }
    if (do_src && !seq) { if ((rc = gather_src_planes())) return rc; }
    // node arrays from the complete records
    if (do_src) {
        ONB_CUDA(cudaStreamWaitEvent(s1, onb_cached_event(c, EV_REC0), 0));
        rc = onb_tree_finish_from_records(c, c->parts[0], c->trees[0], rb0.p); if (rc) return rc;
    }
    if (do_tgt) {
        ONB_CUDA(cudaStreamWaitEvent(s2, onb_cached_event(c, EV_REC1), 0));
        if (s2 != s1) c->cur_stream = s2;
        rc = onb_tree_finish_from_records(c, c->parts[1], c->trees[1], rb1.p);
        c->cur_stream = nullptr;
        if (rc) return rc;
    }
    if (s2 != s1) { ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_JOIN), s2)); ONB_CUDA(cudaStreamWaitEvent(s1, onb_cached_event(c, EV_JOIN), 0)); }
    if (!both && do_src) {      // the separate-call sequence: complete when the call returns
        ONB_CUDA(cudaStreamWaitEvent(s1, c->ev_src_planes, 0)); c->src_planes_pending = false;
    }
    ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_END), s1));
    ONB_CUDA(cudaEventSynchronize(onb_cached_event(c, EV_END)));
    float ms = 0.f; cudaEventElapsedTime(&ms, onb_cached_event(c, EV_E0), onb_cached_event(c, EV_END));
    c->phase_ms["tree"] = ms; if (both) c->phase_ms["trees"] = ms;
    if (c->mem_mode == ONB_MEM_LEAN) onb_scratch_trim(c);       // the build scratch (~40 B per particle) goes back to the driver
    return ONB_OK;
}

// make the stream wait for the all-gather of the source planes if it is still in flight
int onb_dist_join_source_planes(onb_context* c, cudaStream_t st) {
    if (c->src_planes_pending) { ONB_CUDA(cudaStreamWaitEvent(st, c->ev_src_planes, 0)); if (st == c->stream) c->src_planes_pending = false; }
    return ONB_OK;
}

namespace {
// The strengths of the equivalent sources a rank anterpolated for its own nodes sit in one node interval PER LEVEL. Sending them
// as they lie would take one broadcast per (level, owner) - 184 small collectives on 8 GPUs, 1.2 ms at N = 1e7 -, so they are
// packed into one contiguous chunk per rank (blocks in level order), exchanged with ONE in-place all-gather and scattered back.
struct EqXArgs { float* s[ONB_MAX_SD]; float* stage; const uint32_t* tab; uint32_t chunk_blocks, ebs; int SD, L, me, nranks; };
__device__ __forceinline__ uint32_t eqx_node(const EqXArgs& a, int q, uint32_t k) {
    const uint32_t* t = a.tab + (size_t)q * (a.L + 1) * 2;
    int l = 0;
    while (l + 1 < a.L && t[(l + 1) * 2 + 1] <= k) ++l;          // levels are few (<= 32): a linear walk
    return t[l * 2] + (k - t[l * 2 + 1]);
}
__global__ void __launch_bounds__(128) k_eqx_pack(const EqXArgs a) {
    const uint32_t k = blockIdx.x;
    const uint32_t node = eqx_node(a, a.me, k);
    float* out = a.stage + ((size_t)a.me * a.chunk_blocks + k) * a.SD * a.ebs;
    for (int d = 0; d < a.SD; ++d) out[d * a.ebs + threadIdx.x] = a.s[d][(size_t)node * a.ebs + threadIdx.x];
}
__global__ void __launch_bounds__(128) k_eqx_unpack(const EqXArgs a) {
    const int q = (int)(blockIdx.x / a.chunk_blocks);
    const uint32_t k = blockIdx.x - (uint32_t)q * a.chunk_blocks;
    if (q == a.me || k >= a.tab[((size_t)q * (a.L + 1) + a.L) * 2 + 1]) return;
    const uint32_t node = eqx_node(a, q, k);
    const float* in = a.stage + ((size_t)q * a.chunk_blocks + k) * a.SD * a.ebs;
    for (int d = 0; d < a.SD; ++d) a.s[d][(size_t)node * a.ebs + threadIdx.x] = in[d * a.ebs + threadIdx.x];
}
}  // namespace

// barycentric upward pass of the source tree, on c->stream: own nodes -> exchange of their strengths (comm stream) while the
// replicated positions pass and the packing of the real sources run -> the few nodes that straddle rank boundaries -> packing
int onb_dist_upward_sources(onb_context* c) {
    DParts& p = c->parts[0]; DParts& ep = c->parts[2]; DTree& t = c->trees[0];
    int rc = ensure_plan(c, 0); if (rc) return rc;
    const ShardPlan& P = c->plan[0];
    cudaStream_t s1 = c->stream, sc = onb_comm_stream(c);
    rc = onb_bary_upward_mode(c, p, ep, t, ONB_UP_OWN); if (rc) return rc;
    rc = onb_plan_upload_shared(c, 0); if (rc) return rc;
    EqXArgs xa; xa.stage = nullptr;
    for (int d = 0; d < ONB_MAX_SD; ++d) xa.s[d] = ep.s[d];
    xa.tab = c->d_eqtab[0]; xa.chunk_blocks = c->eq_chunk_blocks[0]; xa.ebs = (uint32_t)c->ebs; xa.SD = c->SD; xa.L = P.levels; xa.me = P.rank; xa.nranks = P.nranks;
    const size_t chunk_bytes = (size_t)xa.chunk_blocks * c->SD * c->ebs * sizeof(float);
    if (xa.chunk_blocks) {
        if (c->eq_stage_cap < chunk_bytes * (size_t)P.nranks) {
            if (c->eq_stage) cudaFree(c->eq_stage);
            c->eq_stage = nullptr; c->eq_stage_cap = 0;
            ONB_CUDA(cudaMalloc((void**)&c->eq_stage, chunk_bytes * (size_t)P.nranks));
            c->eq_stage_cap = chunk_bytes * (size_t)P.nranks;
        }
        xa.stage = c->eq_stage;
        uint32_t mine = 0;
        for (int l = 0; l + 1 < P.levels; ++l) mine += P.own_hi[l] - P.own_lo[l];
        if (mine) { k_eqx_pack<<<mine, 128, 0, s1>>>(xa); ONB_LAUNCH(c); ONB_CUDA(cudaGetLastError()); }
    }
    ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_UP), s1));
    ONB_CUDA(cudaStreamWaitEvent(sc, onb_cached_event(c, EV_UP), 0));
    ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_EQ0), sc));
    if (xa.chunk_blocks) { rc = onb_comm_allgather(c, {(void*)c->eq_stage}, {chunk_bytes}); if (rc) return rc; }
    ONB_CUDA(cudaEventRecord(onb_cached_event(c, EV_EQ), sc));
    c->eq_timed = true;
    rc = onb_dist_join_source_planes(c, s1); if (rc) return rc;
    rc = onb_bary_upward_mode(c, p, ep, t, ONB_UP_POS); if (rc) return rc;          // needs r of every node's first particle: after the plane gather
    if (!p.packed_valid) { rc = onb_pack_sources(c, p); if (rc) return rc; }
    ONB_CUDA(cudaStreamWaitEvent(s1, onb_cached_event(c, EV_EQ), 0));
    if (xa.chunk_blocks) { k_eqx_unpack<<<xa.chunk_blocks * (uint32_t)P.nranks, 128, 0, s1>>>(xa); ONB_LAUNCH(c); ONB_CUDA(cudaGetLastError()); }
    rc = onb_bary_upward_mode(c, p, ep, t, ONB_UP_SHARED); if (rc) return rc;
    return onb_pack_sources(c, ep);
}

// device times of the two big exchanges of the last step (valid once the streams have been synchronised): the in-place
// all-gather of the source planes and the broadcasts of the equivalent strengths, with the bytes each rank ends up holding
void onb_dist_record_exchange_times(onb_context* c) {
    float ms = 0.f;
    if (c->ag_timed && cudaEventElapsedTime(&ms, onb_cached_event(c, EV_AG0), onb_cached_event(c, EV_AG1)) == cudaSuccess) {
        c->phase_ms["ag_src_planes"] = ms;
        c->phase_ms["ag_src_planes_bytes"] = (double)c->plan[0].chunk * c->shard_n * sizeof(float) * (c->PD + 1 + c->SD);
    }
    if (c->eq_timed && cudaEventElapsedTime(&ms, onb_cached_event(c, EV_EQ0), onb_cached_event(c, EV_EQ)) == cudaSuccess) {
        c->phase_ms["bcast_eq_strengths"] = ms;          // (name kept from the broadcast form: now one all-gather of packed chunks)
        c->phase_ms["bcast_eq_strengths_bytes"] = (double)c->eq_chunk_blocks[0] * c->shard_n * c->ebs * sizeof(float) * c->SD;
    }
    cudaGetLastError();
    c->ag_timed = c->eq_timed = false;
}
