/*
 * traverse.cu - GPU-side interaction-list generation under the reference's reciprocal multipole-acceptance
 * criteria, and the drivers that pair the lists with the leaf-block kernels of p2p.cu.
 *
 * Replaces (reference): treecode3_block + nbody_treecode3 barneshut.hpp:228-337 (boxwise),
 *                       nbody_fastsumm ongrav3d.cpp:206-452 == onvort3d.cpp:219-465 == onvort2d.cpp:193-439 (dual tree).
 * (pointwise treecode2 / treecode1 live in pointwise.cu)
 *
 * The reference's lists are implicit in its recursion; here they are explicit CSR arrays so that the pair kernel
 * gets one CTA per target block with a contiguous, ORDER-PRESERVING list (accumulation order == the reference's
 * recursion order, which is what lets ARITH_STRICT reproduce its results bit for bit).
 *
 * Every acceptance test is evaluated in the reference's exact IEEE sequence, including the float->double->float
 * round trips that std::pow(float,int) introduces (barneshut.hpp:255, ongrav3d.cpp:338): an accept/reject that
 * flips changes the printed GFlop checksum and moves results at the 1e-4 level.
 *
 * Boxwise: one thread per target leaf walks the source tree depth first with a small stack (count pass, scan,
 * fill pass). Dual tree: level-synchronous over the target tree; one warp per target node
 * consumes its inherited list 32 entries at a time, ballot-compacting into (a) its interaction list, (b) the list its
 * children inherit, (c) a FIFO of opened source nodes that is processed in the following rounds - exactly the
 * "push_back onto the list being iterated" order of the reference.
 */
#include "onb_internal.h"
#include <cstdlib>
#include <algorithm>
#include <cstdio>

namespace {

__device__ __forceinline__ float dist_sq_step(float dist, float v) {
    // dist += std::pow(v, 2) with float dist: promoted to double, rounded back to float each step
    return __double2float_rn(__dadd_rn((double)dist, __dmul_rn((double)v, (double)v)));
}

// ---------------------------------------------------------------------------------------------
// boxwise (treecode3)
// ---------------------------------------------------------------------------------------------
struct BoxArgs {
    TreeView st, tt;
    const uint32_t* leaf_nodes;     // target leaf node ids of this shard, tree order
    uint32_t nleaves;
    uint32_t* counts;               // per leaf (count pass)
    const uint32_t* start;          // per leaf (fill pass)
    uint32_t* entries;
    unsigned long long* stats;      // [0] sltp [1] sbtp [9] pairs
    const uint32_t* s_epnum;        // legacy equivalents: per source node count (null = num_eqps everywhere)
    uint32_t block, num_eqps; int PD; float theta;
};

template <bool FILL>
__global__ void k_boxwise(const BoxArgs a) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= a.nleaves) return;
    const uint32_t T = a.leaf_nodes[w];
    float tc[3]; for (int d = 0; d < a.PD; ++d) tc[d] = a.tt.nc[d][T];
    const float tnr = a.tt.nr[T];
    const uint32_t tcnt = a.tt.num[T];
    uint32_t stack[64]; int sp = 0;
    stack[sp++] = 1;
    uint32_t n = 0, nl = 0, nb = 0; unsigned long long pairs = 0;
    uint32_t* out = FILL ? a.entries + a.start[w] : nullptr;
    while (sp > 0) {
        const uint32_t S = stack[--sp];
        const uint32_t sn = a.st.num[S];
        if (sn <= a.block) {                                                              // barneshut.hpp:242 leaf first, no MAC
            if (FILL) { out[n] = S; pairs += (unsigned long long)sn * tcnt; }
            ++n; ++nl; continue;
        }
        float dist = 0.0f;
        for (int d = 0; d < a.PD; ++d) dist = dist_sq_step(dist, __fsub_rn(a.st.nc[d][S], tc[d]));   // :255
        dist = __fsqrt_rn(dist);
        const float snr = a.st.nr[S];
        const float testrad = __fadd_rn(fmaxf(snr, tnr), __fmul_rn(0.25f, fminf(snr, tnr)));          // :280
        if (__fdiv_rn(dist, __fmul_rn(2.0f, testrad)) > a.theta) {                                     // :283
            if (FILL) { out[n] = S | 0x80000000u; pairs += (unsigned long long)(a.s_epnum ? a.s_epnum[S] : a.num_eqps) * tcnt; }
            ++n; ++nb;
        } else {
            if (sp + 2 > 64) { continue; }      // cannot happen: depth <= levels <= 32
            stack[sp++] = 2 * S + 1;            // :291-292 left child is visited first
            stack[sp++] = 2 * S;
        }
    }
    if (!FILL) a.counts[w] = n;
    else {
        atomicAdd(&a.stats[0], (unsigned long long)nl); atomicAdd(&a.stats[1], (unsigned long long)nb);
        atomicAdd(&a.stats[9], pairs);
    }
}

// leaf table: leaf j of a tree = the node whose particles start at j*block
__global__ void k_leaf_table(TreeView t, uint32_t block, uint32_t* leaf_nodes) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node == 0 || node >= (uint32_t)t.numnodes) return;
    const uint32_t n = t.num[node];
    if (n > 0 && n <= block) leaf_nodes[t.ioffset[node] / block] = node;
}

// ---------------------------------------------------------------------------------------------
// dual tree (fastsumm)
// ---------------------------------------------------------------------------------------------
struct DttArgs {
    TreeView st, tt;
    int level; uint32_t nnodes;             // target nodes of this level handled here: T = node0 + local (the whole level, or the
    uint32_t node0, pnode0;                 // nodes overlapping this rank's shard); pnode0 = node0 of the level above
    // All lists of one evaluation live in ONE device pool, bump-allocated level by level ON THE DEVICE (k_dtt_bases):
    // bases[2*l] / bases[2*l+1] = first pool entry of level l's interaction / deferred lists. The host never needs a list
    // size, so an evaluation is enqueued without a single synchronisation - first call or hundredth, old input or new.
    uint32_t* pool; uint32_t pool_cap;      // every access is clamped to pool_cap (an overflowing pass is redone with a larger pool)
    const uint32_t* bases;
    // inherited lists = the parent's deferred list
    const uint32_t* pc_start;               // per node of level-1 (+1), relative to bases[2*(level-1)+1]; null at the root level
    // outputs
    uint32_t* icount; uint32_t* ccount;     // count pass, per node of this level
    const uint32_t* istart; const uint32_t* cstart;   // fill pass, relative to this level's bases
    uint32_t* queue; uint32_t qcap;         // per-warp FIFO scratch: 2 x qcap entries per warp slot
    unsigned long long* stats;              // [2] sltl [3] sbtl [4] sltb [5] sbtb [6] tlc [7] lpc [8] bpc [9] pairs
    unsigned long long* ctl;                // [0] pool top [1] FIFO overflow flag [2] pool overflow flag
    uint32_t block, num_eqps, shard_lo, shard_hi; int PD; float theta;
};

template <bool FILL>
__global__ void __launch_bounds__(256) k_dtt(const DttArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t wslot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t* q0 = a.queue + (size_t)wslot * 2 * a.qcap;
    uint32_t* q1 = q0 + a.qcap;
    unsigned long long st_sltl = 0, st_sbtl = 0, st_sltb = 0, st_sbtb = 0, st_pairs = 0, st_tlc = 0, st_lpc = 0, st_bpc = 0;

    for (uint32_t local = wslot; local < a.nnodes; local += nwarps) {
        const uint32_t T = a.node0 + local;
        const uint32_t tn = a.tt.num[T];
        const uint32_t tio = a.tt.ioffset[T];
        const bool needed = tn >= 1 && tio < a.shard_hi && tio + tn > a.shard_lo;         // ongrav3d.cpp:221 (+ sharding)
        if (!needed) { if (!FILL && lane == 0) { a.icount[local] = 0; a.ccount[local] = 0; } continue; }
        const bool tleaf = tn <= a.block;                                                 // :224
        const uint32_t tcnt = tleaf ? tn : a.num_eqps;
        float tx[3]; for (int d = 0; d < a.PD; ++d) tx[d] = a.tt.x[d][T];
        const float tnr = a.tt.nr[T];
        if (FILL && lane == 0) { if (tleaf) { ++st_tlc; if (T > 1) ++st_lpc; } else if (T > 1) ++st_bpc; }

        const uint32_t* cur; uint32_t len;
        uint32_t root_list = 1;
        if (a.level == 0) { cur = nullptr; len = 1; }
        else {
            const uint32_t pl = (T >> 1) - a.pnode0;
            const uint32_t pb = a.bases[2 * (a.level - 1) + 1];
            const uint32_t s0 = (uint32_t)min((unsigned long long)pb + a.pc_start[pl], (unsigned long long)a.pool_cap);     // clamp to the pool
            const uint32_t s1 = (uint32_t)min((unsigned long long)pb + a.pc_start[pl + 1], (unsigned long long)a.pool_cap);
            cur = a.pool + s0; len = s1 - s0;
        }
        uint32_t nI = 0, nC = 0;
        const unsigned long long oI = FILL ? (unsigned long long)a.bases[2 * a.level] + a.istart[local] : 0ull;       // first pool entry of my interaction list
        const unsigned long long oC = FILL ? (unsigned long long)a.bases[2 * a.level + 1] + a.cstart[local] : 0ull;   // ... and of the list my children inherit
        uint32_t* nxt = q0;
        bool overflow = false;
        while (len > 0 && !overflow) {
            uint32_t nextlen = 0;
            for (uint32_t base = 0; base < len; base += 32) {                             // :315 list order
                const uint32_t i = base + lane;
                const bool valid = i < len;
                const uint32_t S = valid ? (a.level == 0 && cur == nullptr ? root_list : cur[i]) : 1u;
                int outcome = 0;   // 0 skip, 1 emit real sources, 2 emit equivalent sources, 3 defer to children, 4 open source node
                uint32_t sn = 0;
                if (valid) {
                    sn = a.st.num[S];
                    if (sn >= 1) {                                                        // :319
                        const bool sleaf = sn <= a.block;
                        if (sleaf && tleaf) outcome = 1;                                  // :326 direct, no MAC
                        else {
                            float dist = 0.0f;
                            for (int d = 0; d < a.PD; ++d) dist = dist_sq_step(dist, __fsub_rn(a.st.x[d][S], tx[d]));   // :338
                            dist = __fsqrt_rn(dist);
                            const float snr = a.st.nr[S];
                            const float diag = __fadd_rn(snr, tnr);                       // :340
                            if (__fdiv_rn(dist, diag) > a.theta) outcome = sleaf ? 1 : 2; // :344-365
                            else if (tnr > snr) outcome = tleaf ? 4 : 3;                  // :367-382
                            else outcome = sleaf ? 3 : 4;                                 // :384-399
                        }
                    }
                }
                const uint32_t bE = __ballot_sync(0xffffffffu, outcome == 1 || outcome == 2);
                const uint32_t bC = __ballot_sync(0xffffffffu, outcome == 3);
                const uint32_t bX = __ballot_sync(0xffffffffu, outcome == 4);
                if (nextlen + 2u * __popc(bX) > a.qcap) { overflow = true; break; }
                if (outcome == 4) { const uint32_t pos = nextlen + 2u * __popc(bX & lt_mask); nxt[pos] = 2 * S; nxt[pos + 1] = 2 * S + 1; }   // :374,:395
                if (FILL) {
                    if (outcome == 1 || outcome == 2) {
                        const unsigned long long pos = oI + nI + __popc(bE & lt_mask);
                        if (pos < a.pool_cap) a.pool[pos] = outcome == 2 ? (S | 0x80000000u) : S;
                        st_pairs += (unsigned long long)(outcome == 2 ? a.num_eqps : sn) * tcnt;
                        if (outcome == 1) { if (tleaf) ++st_sltl; else ++st_sltb; } else { if (tleaf) ++st_sbtl; else ++st_sbtb; }
                    }
                    if (outcome == 3) { const unsigned long long pos = oC + nC + __popc(bC & lt_mask); if (pos < a.pool_cap) a.pool[pos] = S; }
                }
                nI += __popc(bE); nC += __popc(bC); nextlen += 2u * __popc(bX);
            }
            __syncwarp();
            cur = nxt; len = nextlen; nxt = (nxt == q0) ? q1 : q0;
        }
        if (overflow && lane == 0) a.ctl[1] = 1ull;
        if (!FILL && lane == 0) { a.icount[local] = nI; a.ccount[local] = tleaf ? 0u : nC; }
    }
    if (FILL) {
        // warp-reduce the statistics, one atomic per warp and counter
        unsigned long long v[8] = { st_sltl, st_sbtl, st_sltb, st_sbtb, st_tlc, st_lpc, st_bpc, st_pairs };
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned long long x = v[k];
            for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            if (lane == 0 && x) atomicAdd(&a.stats[2 + k], x);
        }
    }
}

// device-side bump allocation of one level's two lists: totals are the last elements of the two exclusive scans
__global__ void k_dtt_bases(const uint32_t* istart, const uint32_t* cstart, uint32_t nn, int level, uint32_t* bases, uint32_t* totals,
                            unsigned long long* ctl, uint32_t pool_cap) {
    const unsigned long long it = istart[nn], ct = cstart[nn];
    const unsigned long long top = ctl[0];
    const unsigned long long ib = top, cb = top + it, nt = cb + ct;
    bases[2 * level] = (uint32_t)(ib < pool_cap ? ib : pool_cap);
    bases[2 * level + 1] = (uint32_t)(cb < pool_cap ? cb : pool_cap);
    totals[2 * level] = (uint32_t)it; totals[2 * level + 1] = (uint32_t)ct;
    ctl[0] = nt;
    if (nt > pool_cap) ctl[2] = 1ull;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host drivers
// ---------------------------------------------------------------------------------------------
void onb_shard_range(const onb_context* c, uint32_t* lo, uint32_t* hi) {
    uint64_t a = 0, b = 0;
    onb_shard_range_for(c->parts[1].n, c->block, c->shard_rank, c->shard_n, &a, &b);      // equal leaf-aligned chunks (plan.cu)
    *lo = (uint32_t)a; *hi = (uint32_t)b;
}

// the target nodes of one level a sharded context works on: those that overlap its particle range (plan.cu); unsharded: all
void onb_level_span(onb_context* c, int level, uint32_t* node0, uint32_t* count) {
    *node0 = 1u << level; *count = 1u << level;
    if (c->shard_n <= 1) return;
    if (onb_plan_make(c->plan[1], c->parts[1].n, c->block, c->shard_n, c->shard_rank) != ONB_OK) return;
    const ShardPlan& P = c->plan[1];
    if (level >= P.levels) { *count = 0; return; }
    *node0 = P.need_lo[level]; *count = P.need_hi[level] - P.need_lo[level];
}

static int fetch_stats(onb_context* c, unsigned long long* d_stats) {
    unsigned long long h[10];
    ONB_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 9; ++i) c->stats[i] = h[i];
    c->last_pairs = h[9];
    return ONB_OK;
}

int onb_lists_boxwise(onb_context* c, float theta, WorkList& wl) {
    DTree& st = c->trees[0]; DTree& tt = c->trees[1];
    const uint32_t nleaf_all = (c->parts[1].n + c->block - 1) / c->block;
    uint32_t* leaf_all = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&leaf_all, (size_t)nleaf_all * 4));
    k_leaf_table<<<(tt.numnodes + 255) / 256, 256, 0, c->stream>>>(view_of(tt), c->block, leaf_all); ONB_LAUNCH(c);
    uint32_t lo, hi; onb_shard_range(c, &lo, &hi);
    const uint32_t l0 = lo / c->block, l1 = (hi + c->block - 1) / c->block;
    const uint32_t nl = l1 - l0;
    unsigned long long* d_stats = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&d_stats, 10 * sizeof(unsigned long long)));
    ONB_CUDA(cudaMemsetAsync(d_stats, 0, 10 * sizeof(unsigned long long), c->stream));
    wl = WorkList(); wl.nitems = nl;
    if (nl == 0) { onb_dfree(c, leaf_all); onb_dfree(c, d_stats); memset(c->stats, 0, sizeof(c->stats)); c->last_pairs = 0; return ONB_OK; }
    ONB_CUDA(onb_dmalloc(c, (void**)&wl.tgt_node, (size_t)nl * 4));
    ONB_CUDA(cudaMemcpyAsync(wl.tgt_node, leaf_all + l0, (size_t)nl * 4, cudaMemcpyDeviceToDevice, c->stream));
    ONB_CUDA(onb_dmalloc(c, (void**)&wl.start, (size_t)(nl + 1) * 4));
    BoxArgs a; a.st = view_of(st); a.tt = view_of(tt); a.leaf_nodes = wl.tgt_node; a.nleaves = nl;
    a.counts = wl.start; a.start = wl.start; a.entries = nullptr; a.stats = d_stats;
    a.block = c->block; a.num_eqps = c->num_eqps; a.PD = c->PD; a.theta = theta;
    a.s_epnum = c->legacy ? c->d_epnum : nullptr;
    const int TB = 128;
    k_boxwise<false><<<(nl + TB - 1) / TB, TB, 0, c->stream>>>(a); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    ONB_CUDA(cudaMemsetAsync(wl.start + nl, 0, 4, c->stream));
    uint64_t total = 0;
    int rc = onb_exclusive_scan_u32(c, wl.start, wl.start, nl + 1, &total);
    if (rc) return rc;
    if (total >= 0xffffffffull) { c->err = "boxwise: interaction list exceeds 2^32 entries on one GPU"; return ONB_ERR_CAPACITY; }
    wl.nentries = total;
    ONB_CUDA(onb_dmalloc(c, (void**)&wl.entries, std::max<size_t>(4, (size_t)total * 4)));
    a.entries = wl.entries;
    k_boxwise<true><<<(nl + TB - 1) / TB, TB, 0, c->stream>>>(a); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    { static const bool lpt = std::getenv("ONB_LPT") != nullptr;
      uint32_t* order = nullptr;
      if (lpt && nl >= 4096) { rc = onb_lpt_order(c, wl.start, nl, &order); if (rc) return rc; }
      wl.order = order; }
    rc = fetch_stats(c, d_stats);
    onb_dfree(c, d_stats); onb_dfree(c, leaf_all);
    return rc;
}

// One dual-tree evaluation, enqueued without a single host synchronisation. All interaction and deferred lists live in one
// persistent device pool that is bump-allocated per level on the device (k_dtt_bases), so the host needs no list size: the
// first evaluation of a context costs what every later one costs, and new particle positions cost nothing extra. The
// traversal of level l+1 depends only on the deferred lists of level l, not on the pair kernels, so the lists are built on
// the second stream and run underneath the pair kernels of the levels above. One read-back at the very end fetches the
// statistics and the two overflow flags (pool too small / a per-warp FIFO too small); an overflowing pass - whose every
// pool access was clamped - is redone with larger buffers (geometric growth: at most a few times in a context's life).
namespace {
struct StreamScope { onb_context* c; explicit StreamScope(onb_context* x) : c(x) {} ~StreamScope() { c->cur_stream = nullptr; } };
}  // namespace

static int fastsumm_pass(onb_context* c, float theta, bool* redo) {
    DTree& st = c->trees[0]; DTree& tt = c->trees[1];
    StreamScope scope(c);
    *redo = false;
    uint32_t lo, hi; onb_shard_range(c, &lo, &hi);
    const int L = tt.levels;
    // pool: first guess 512 entries per target leaf of the shard (a uniform cloud at theta 1.4 uses ~220: 17.2 M entries for 78125 leaves at N = 1e7), grown on overflow
    if (!c->dtt_pool) {
        const uint64_t nleaf = (hi > lo ? (uint64_t)(hi - lo) : 0) / (uint64_t)c->block + 1;
        uint64_t want = std::max<uint64_t>(c->dtt_pool_want, std::max<uint64_t>((uint64_t)1 << 20, nleaf * 512));
        if (!c->dtt_pool_want) {                                   // tests: start small so that the overflow path runs
            if (const char* e = std::getenv("ONB_DTT_POOL_INIT")) want = std::max<uint64_t>(16, strtoull(e, nullptr, 10));
            if (const char* e = std::getenv("ONB_DTT_QCAP_INIT")) c->dtt_qcap = (uint32_t)std::max<uint64_t>(4, strtoull(e, nullptr, 10));
        }
        want = std::min<uint64_t>(want, 0xfffffff0ull);
        ONB_CUDA(onb_pmalloc(c, (void**)&c->dtt_pool, (size_t)want * 4));
        c->dtt_pool_cap = (uint32_t)want;
    }
    unsigned long long* d_stats = nullptr;   // [0..9] counters, [10..13] ctl: pool top, FIFO overflow, pool overflow
    ONB_CUDA(onb_dmalloc(c, (void**)&d_stats, 14 * sizeof(unsigned long long)));
    ONB_CUDA(cudaMemsetAsync(d_stats, 0, 14 * sizeof(unsigned long long), c->stream));
    unsigned long long* d_ctl = d_stats + 10;
    const int TB = 256, WPB = TB / 32;
    // per-warp FIFO of opened source nodes: 2 x qcap entries per resident warp, the whole scratch kept under 256 MB
    const uint32_t qcap = c->dtt_qcap;
    uint32_t max_blocks = (uint32_t)c->sm_count * 8u;
    while (max_blocks > 16u && (size_t)max_blocks * WPB * 2 * qcap * 4 > ((size_t)256 << 20)) max_blocks >>= 1;
    max_blocks = std::min<uint32_t>(max_blocks, ((1u << (L - 1)) + WPB - 1) / WPB);
    uint32_t* queue = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&queue, (size_t)max_blocks * WPB * 2 * qcap * 4));
    uint32_t *d_bases = nullptr, *d_totals = nullptr;      // [2*L] each
    ONB_CUDA(onb_dmalloc(c, (void**)&d_bases, (size_t)4 * L * 4));
    ONB_CUDA(cudaMemsetAsync(d_bases, 0, (size_t)4 * L * 4, c->stream));
    d_totals = d_bases + 2 * L;
    static const bool want_prof = std::getenv("ONB_DTT_PROF") != nullptr;      // diagnostics: pairs and pair-kernel time per level
    unsigned long long* d_lvl = nullptr;
    if (want_prof) ONB_CUDA(onb_dmalloc(c, (void**)&d_lvl, (size_t)L * 8));

    static const bool no_ahead = std::getenv("ONB_DTT_ONE_STREAM") != nullptr;
    const bool ahead = !no_ahead;
    auto ev = [&](int lev, int k) { return onb_cached_event(c, (size_t)5 * lev + k); };    // 0..3 phase stamps, 4 = pair stream joined
    if (ahead) {      // the list stream starts behind everything enqueued so far (statistics reset, earlier phases)
        ONB_CUDA(cudaEventRecord(ev(0, 4), c->stream));
        ONB_CUDA(cudaStreamWaitEvent(c->stream2, ev(0, 4), 0));
    }
    uint32_t* pc_start = nullptr;            // previous level's deferred-list offsets
    uint32_t pnode0 = 0;
    int rc = ONB_OK, levels_done = 0;
    for (int lev = 0; lev < L && rc == ONB_OK; ++lev) {
        uint32_t node0, nn; onb_level_span(c, lev, &node0, &nn);      // the whole level, or the nodes that overlap this rank's shard
        if (nn == 0) break;                                            // (nothing below either: children of nothing)
        levels_done = lev + 1;
        if (ahead) c->cur_stream = c->stream2;
        ONB_CUDA(cudaEventRecord(ev(lev, 0), ONB_ST(c)));
        uint32_t *istart = nullptr, *cstart = nullptr;
        ONB_CUDA(onb_dmalloc(c, (void**)&istart, (size_t)(nn + 1) * 4)); ONB_CUDA(onb_dmalloc(c, (void**)&cstart, (size_t)(nn + 1) * 4));
        DttArgs a; a.st = view_of(st); a.tt = view_of(tt); a.level = lev; a.nnodes = nn; a.node0 = node0; a.pnode0 = pnode0;
        a.pool = c->dtt_pool; a.pool_cap = c->dtt_pool_cap; a.bases = d_bases; a.pc_start = pc_start;
        a.icount = istart; a.ccount = cstart; a.istart = istart; a.cstart = cstart;
        a.queue = queue; a.qcap = qcap; a.stats = d_stats; a.ctl = d_ctl;
        a.block = c->block; a.num_eqps = c->num_eqps; a.shard_lo = lo; a.shard_hi = hi; a.PD = c->PD; a.theta = theta;
        const uint32_t blocks = std::min<uint32_t>(max_blocks, (nn + WPB - 1) / WPB);
        k_dtt<false><<<blocks, TB, 0, ONB_ST(c)>>>(a); ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        ONB_CUDA(cudaMemsetAsync(istart + nn, 0, 4, ONB_ST(c))); ONB_CUDA(cudaMemsetAsync(cstart + nn, 0, 4, ONB_ST(c)));
        if ((rc = onb_exclusive_scan_u32(c, istart, istart, nn + 1, nullptr))) break;
        if ((rc = onb_exclusive_scan_u32(c, cstart, cstart, nn + 1, nullptr))) break;
        k_dtt_bases<<<1, 1, 0, ONB_ST(c)>>>(istart, cstart, nn, lev, d_bases, d_totals, d_ctl, c->dtt_pool_cap); ONB_LAUNCH(c);
        k_dtt<true><<<blocks, TB, 0, ONB_ST(c)>>>(a); ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        if (d_lvl) ONB_CUDA(cudaMemcpyAsync(d_lvl + lev, d_stats + 9, 8, cudaMemcpyDeviceToDevice, ONB_ST(c)));
        // optional launch order of this level's pair kernel: longest list first (ONB_LPT=1; built on the list stream). Measured
        // at N = 1e7: no gain on one GPU (79.2 vs 80.1 ms of pair kernels, within box-to-box noise) and none on a 1/8 shard - the
        // items of a level have similar lengths, the loss of small launches is wave quantisation, not stragglers - so it is off.
        uint32_t* order = nullptr;
        static const bool lpt = std::getenv("ONB_LPT") != nullptr;
        if (lpt && nn >= 4096) { if ((rc = onb_lpt_order(c, istart, nn, &order))) break; }
        ONB_CUDA(cudaEventRecord(ev(lev, 1), ONB_ST(c)));
        if (ahead) { c->cur_stream = nullptr; ONB_CUDA(cudaStreamWaitEvent(c->stream, ev(lev, 1), 0)); ONB_CUDA(cudaEventRecord(ev(lev, 4), c->stream)); }
        // node entry: zero + interpolate from the parent (ongrav3d.cpp:232-304)
        if ((rc = onb_bary_downward_level(c, lev))) break;
        ONB_CUDA(cudaEventRecord(ev(lev, 2), c->stream));
        // then this level's interactions, in list order, on top of the interpolated values (:315-402)
        WorkList wl; wl.nitems = nn; wl.tgt_node = nullptr; wl.node_base = node0; wl.start = istart; wl.entries = c->dtt_pool; wl.nentries = c->dtt_pool_cap;
        wl.ebase = d_bases + 2 * lev; wl.order = order;
        // Upper levels have too few target nodes to fill the machine with one warp per node. Cutting every list into nsplit
        // segments (ONB_P2P_SPLIT_TARGET=<CTAs per launch to aim for>) makes the pair kernel 4 % faster at N = 1e7 (79.5 ->
        // 76.1 ms), but it is OFF by default: adding the far-field partial sums in another order than the reference moves
        // the result 2.1e-6 (relative rms) away from the reference treecode - the noise of its own float summation -
        // where the same order stays within 4.4e-7, and the contract is 1e-6. nsplit depends on the LEVEL only (not on
        // this rank's share of it), so a multi-GPU run would still add in exactly the order of the single-GPU run.
        uint32_t nsplit = 1;
        { static int target = -1;
          if (target < 0) { target = 0; if (const char* e = std::getenv("ONB_P2P_SPLIT_TARGET")) target = std::max(0, atoi(e)); }
          while (nsplit < 16u && (unsigned long long)nn * nsplit < (unsigned long long)target) nsplit <<= 1; }
        if ((rc = onb_p2p_lists(c, wl, 1, 3, true, nsplit))) break;
        ONB_CUDA(cudaEventRecord(ev(lev, 3), c->stream));
        pc_start = cstart; pnode0 = node0;
    }
    c->cur_stream = nullptr;
    if (rc != ONB_OK) return rc;
    // the one read-back of the evaluation: counters, overflow flags, per-level totals
    unsigned long long h[14];
    std::vector<uint32_t> h_tot(2 * L);
    ONB_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    ONB_CUDA(cudaMemcpyAsync(h_tot.data(), d_totals, (size_t)2 * L * 4, cudaMemcpyDeviceToHost, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    if (h[11]) {                       // a target node opened more source nodes in one round than its FIFO holds
        if (c->dtt_qcap >= (1u << 22)) { c->err = "dual-tree traversal: per-node source FIFO overflow"; return ONB_ERR_CAPACITY; }
        c->dtt_qcap *= 8; *redo = true; return ONB_OK;
    }
    if (h[12]) {                       // the pool was too small: at least h[10] entries are needed (a lower bound, lists were clamped)
        const uint64_t want = std::max<uint64_t>((uint64_t)c->dtt_pool_cap * 4, h[10] + h[10] / 4);
        if (c->dtt_pool_cap >= 0xfffffff0u) { c->err = "fastsumm: lists exceed 2^32 entries on one GPU"; return ONB_ERR_CAPACITY; }
        ONB_CUDA(cudaStreamSynchronize(c->stream2));
        onb_pfree(c, c->dtt_pool); c->dtt_pool = nullptr; c->dtt_pool_cap = 0;
        c->dtt_pool_want = std::min<uint64_t>(want, 0xfffffff0ull);
        *redo = true; return ONB_OK;
    }
    for (int i = 0; i < 9; ++i) c->stats[i] = h[i];
    c->last_pairs = h[9];
    std::vector<unsigned long long> h_lvl(L, 0);
    if (d_lvl) ONB_CUDA(cudaMemcpy(h_lvl.data(), d_lvl, (size_t)L * 8, cudaMemcpyDeviceToHost));
    double ms_lists = 0.0, ms_p2p = 0.0, ms_down = 0.0;
    for (int lev = 0; lev < levels_done; ++lev) {
        float t01 = 0, t12 = 0, t23 = 0;
        cudaEventElapsedTime(&t01, ev(lev, 0), ev(lev, 1)); cudaEventElapsedTime(&t12, ahead ? ev(lev, 4) : ev(lev, 1), ev(lev, 2)); cudaEventElapsedTime(&t23, ev(lev, 2), ev(lev, 3));
        ms_lists += t01; ms_down += t12; ms_p2p += t23;
        if (d_lvl) {
            const unsigned long long pr = h_lvl[lev] - (lev ? h_lvl[lev - 1] : 0ull);
            fprintf(stderr, "dtt level %2d: entries %10u deferred %10u pairs %14llu lists %7.3f ms down %7.3f ms p2p %8.3f ms -> %7.1f Gpairs/s\n", lev, h_tot[2 * lev], h_tot[2 * lev + 1], pr, t01, t12, t23, t23 > 0 ? pr / t23 * 1e-6 : 0.0);
        }
    }
    c->phase_ms["lists"] = ms_lists; c->phase_ms["downward"] = ms_down; c->phase_ms["p2p"] = ms_p2p;
    c->phase_ms["dtt_pool_used"] = (double)h[10]; c->phase_ms["dtt_pool_cap"] = (double)c->dtt_pool_cap;
    return ONB_OK;
}

int onb_run_fastsumm(onb_context* c, float theta) {
    for (int attempt = 0; attempt < 24; ++attempt) {
        bool redo = false;
        const int rc = fastsumm_pass(c, theta, &redo);
        if (rc != ONB_OK) return rc;
        if (!redo) { c->phase_ms["dtt_attempts"] = attempt + 1; return ONB_OK; }
        onb_scratch_reset(c);          // the discarded pass is complete (its read-back synchronised): reuse its scratch
    }
    c->err = "fastsumm: list buffers kept overflowing"; return ONB_ERR_CAPACITY;
}
