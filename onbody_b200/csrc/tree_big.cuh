/*
 * tree_big.cuh - the same partial select as k_node_split (tree.cu), for nodes too large for one CTA.
 *
 * The top levels of the k-d tree have few, huge nodes (the root holds every particle), so a CTA-per-node kernel would
 * leave 147 SMs idle exactly where most of the data is. Here every "big" node of a level is cut into fixed chunks
 * of 2048 particles and each partition pass (barneshut.hpp:527-586) becomes five grid-wide launches over all chunks
 * of all big nodes of the level:
 *
 *   k_big_count    per chunk: #{v < pivot}, max{v < pivot}, min{v >= pivot}            -> atomics on the node
 *   (misplaced counts per chunk follow from the count and the chunk's position relative to B - no second pass)
 *   k_big_scan     per node : exclusive scan of the chunk counts, k, next window + next pivot (the reference's exit rules)
 *   k_big_compact  per chunk: ordered compaction of the misplaced positions into scr[]
 *   k_big_swap     per chunk: swap pair j of the node, a_j <-> b_j
 *
 * Node state lives in global memory, double buffered by pass parity so that compact/swap still see the window the
 * pass started with while scan has already published the next one. Results are bit-identical to k_node_split: the
 * passes, pivots and exit conditions are the same, only the parallel decomposition differs.
 */
#pragma once
// (tree.cu includes <cooperative_groups.h> at global scope and defines `namespace cg` before including this file)

constexpr int BIG_T = 256;                 // threads per chunk CTA
constexpr int BIG_ROUNDS = 8;              // 8 warps x 8 rounds x 32 lanes
constexpr uint32_t BIG_CH = BIG_T * BIG_ROUNDS;   // 2048 particles per chunk
constexpr uint32_t BIG_LOCAL = 4096;       // a window this small is finished by ONE CTA with block barriers only (cta_select)

struct BigWin { uint32_t wf, wl; float lo, hi, pivot; int done, iters, local; };   // local: finish in one CTA (big_local_finish)
struct BigNode {
    uint32_t node, pf, pl, nless, chunk0, nchunks; int axis; float ideal;
    BigWin w[2];
    uint32_t m, mx_enc, mn_enc;            // atomics of the current pass (reset by k_big_scan)
    uint32_t B, k;
    uint32_t bmin[3], bmax[3];             // order-encoded bounding box atomics
    uint32_t npass, nstall; unsigned long long nscan;
    uint32_t arrived, expect;              // count-phase ticket: the last chunk of the node to arrive runs the node's scan
};

struct BigArgs {
    float* x[3]; TreeView t;
    BigNode* nodes; uint32_t* nbig;        // nbig[0] = number of big nodes of this level, nbig[1] = total chunks, nbig[2] = still active
    uint32_t* chunk_owner;                 // chunk -> index into nodes
    uint32_t* cntA; uint32_t* cntB;        // per chunk: counts, then exclusive offsets
    uint32_t* wl;                          // 2 x max_chunks: the chunks that intersect a live window, by pass parity (count in nbig[4], nbig[5])
    uint32_t* lidx; uint32_t* scr;
    uint8_t* axis_of; uint32_t* pmid;
    unsigned long long* stats;
    uint32_t block, big, max_nodes, max_chunks, blo, bhi; int level, PD, pivot_mode;
    unsigned long long* prof;              // diagnostics (ONB_BIG_PROF): globaltimer at every phase boundary, block 0
};

// one block: list the big nodes of this level (node order) and lay out their chunks - a block-wide exclusive scan of
// (is big, number of chunks) over the nodes of the level
__device__ __forceinline__ void big_list(const BigArgs& a) {                         // one CTA
    __shared__ uint32_t s_wn[BIG_T / 32], s_wc[BIG_T / 32], s_bn, s_bc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_bn = 0; s_bc = 0; }
    __syncthreads();
    const uint32_t first = 1u << a.level, last = 2u << a.level;
    for (uint32_t base = first; base < last; base += BIG_T) {
        const uint32_t node = base + threadIdx.x;
        uint32_t flag = 0, nch = 0, n = 0, pf0 = 0;
        if (node < last) {
            n = a.t.num[node]; pf0 = a.t.ioffset[node];
            if (n > a.big && pf0 < a.bhi && pf0 + n > a.blo) { flag = 1; nch = (n + BIG_CH - 1) / BIG_CH; }
        }
        uint32_t in = flag, ic = nch;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t tn = __shfl_up_sync(0xffffffffu, in, o), tc = __shfl_up_sync(0xffffffffu, ic, o); if (lane >= o) { in += tn; ic += tc; } }
        if (lane == 31) { s_wn[warp] = in; s_wc[warp] = ic; }
        __syncthreads();
        uint32_t pn = s_bn, pc = s_bc, tn = 0, tc = 0;
        for (int q = 0; q < BIG_T / 32; ++q) { if (q < warp) { pn += s_wn[q]; pc += s_wc[q]; } tn += s_wn[q]; tc += s_wc[q]; }
        const uint32_t nb = pn + in - flag, nc = pc + ic - nch;
        if (flag && nb < a.max_nodes) {
            BigNode& b = a.nodes[nb];
            b.node = node; b.pf = pf0; b.pl = pf0 + n; b.chunk0 = nc; b.nchunks = nch;
            for (int d = 0; d < 3; ++d) { b.bmin[d] = 0xffffffffu; b.bmax[d] = 0u; }
            b.m = 0; b.mx_enc = 0u; b.mn_enc = 0xffffffffu; b.npass = 0; b.nstall = 0; b.nscan = 0;
            b.arrived = 0; b.expect = nch;
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_bn += tn; s_bc += tc; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { const uint32_t nb = min(s_bn, a.max_nodes); a.nbig[0] = nb; a.nbig[1] = s_bc; a.nbig[2] = nb; }
}

__device__ __forceinline__ float bw_min(float v) { return warp_min(v); }
__device__ __forceinline__ float bw_max(float v) { return warp_max(v); }
__device__ __forceinline__ uint32_t bw_sum(uint32_t v) { return warp_sum(v); }

// bounding boxes (barneshut.hpp:621-625) and lidx = iota (:516) over this CTA's contiguous share of the level's chunks:
// per-thread min/max accumulate across the chunks of a node with no barrier in between (the loads of consecutive
// chunks overlap); one block reduction + six atomics per (CTA, node)
__device__ __forceinline__ void big_bbox_flush(const BigArgs& a, const uint32_t bi, float* lo, float* hi) {
    __shared__ uint32_t s_red[6][BIG_T / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    #pragma unroll
    for (int d = 0; d < 3; ++d) {
        const uint32_t l = __reduce_min_sync(0xffffffffu, f2ord(lo[d])), h = __reduce_max_sync(0xffffffffu, f2ord(hi[d]));
        if (lane == 0) { s_red[d][warp] = l; s_red[3 + d][warp] = h; }
        lo[d] = INFINITY; hi[d] = -INFINITY;
    }
    __syncthreads();
    if (warp < 6 && warp % 3 < a.PD) {
        const uint32_t v = lane < BIG_T / 32 ? s_red[warp][lane] : (warp < 3 ? 0xffffffffu : 0u);
        const uint32_t r = warp < 3 ? __reduce_min_sync(0xffffffffu, v) : __reduce_max_sync(0xffffffffu, v);
        if (lane == 0) { if (warp < 3) atomicMin(&a.nodes[bi].bmin[warp], r); else atomicMax(&a.nodes[bi].bmax[warp - 3], r); }
    }
    __syncthreads();
}
__device__ __forceinline__ void big_bbox_phase(const BigArgs& a, const uint32_t nchunks) {
    __shared__ uint32_t s_owner[64];
    __shared__ uint32_t s_nd[3];     // pf, pl, chunk0 of the current node
    const uint32_t i0 = (uint32_t)((unsigned long long)nchunks * blockIdx.x / gridDim.x);
    const uint32_t i1 = (uint32_t)((unsigned long long)nchunks * (blockIdx.x + 1) / gridDim.x);
    if (i0 >= i1) return;
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    uint32_t cur = 0xffffffffu;
    for (uint32_t base = i0; base < i1; base += 64) {
        const uint32_t m = min(64u, i1 - base);
        __syncthreads();
        if (threadIdx.x < m) s_owner[threadIdx.x] = a.chunk_owner[base + threadIdx.x];
        __syncthreads();
        for (uint32_t j = 0; j < m; ++j) {
            const uint32_t bi = s_owner[j], chunk = base + j;
            if (bi != cur) {
                if (cur != 0xffffffffu) big_bbox_flush(a, cur, lo, hi);
                __syncthreads();
                if (threadIdx.x == 0) { const BigNode& n = a.nodes[bi]; s_nd[0] = n.pf; s_nd[1] = n.pl; s_nd[2] = n.chunk0; }
                __syncthreads();
                cur = bi;
            }
            const uint32_t c0 = s_nd[0] + (chunk - s_nd[2]) * BIG_CH, c1 = min(s_nd[1], c0 + BIG_CH);
            #pragma unroll
            for (int d = 0; d < 3; ++d) if (d < a.PD) {
                const float* __restrict__ xd = a.x[d];
                float v[BIG_ROUNDS];
                #pragma unroll
                for (int r = 0; r < BIG_ROUNDS; ++r) { const uint32_t i = c0 + (uint32_t)r * BIG_T + threadIdx.x; v[r] = i < c1 ? xd[i] : NAN; }   // 8 loads in flight
                #pragma unroll
                for (int r = 0; r < BIG_ROUNDS; ++r) { lo[d] = fminf(lo[d], v[r]); hi[d] = fmaxf(hi[d], v[r]); }      // fminf/fmaxf ignore the NaN fillers
            }
            #pragma unroll
            for (int r = 0; r < BIG_ROUNDS; ++r) { const uint32_t i = c0 + (uint32_t)r * BIG_T + threadIdx.x; if (i < c1) a.lidx[i] = i; }
        }
    }
    if (cur != 0xffffffffu) big_bbox_flush(a, cur, lo, hi);
}

// per node: node arrays, split axis, first window and pivot (barneshut.hpp:623-663, :519-540)
__device__ __forceinline__ void big_setup(const BigArgs& a, const uint32_t bi) {      // one thread per node
    BigNode& b = a.nodes[bi];
    const uint32_t node = b.node;
    float lo[3], hi[3], bsss = 0.0f;
    int axis = 0; float axsz = -1.0f;
    for (int d = 0; d < a.PD; ++d) {
        lo[d] = ord2f(b.bmin[d]); hi[d] = ord2f(b.bmax[d]);
        const float ns = __fsub_rn(hi[d], lo[d]);
        a.t.ns[d][node] = ns;
        a.t.nc[d][node] = __fmul_rn(0.5f, __fadd_rn(hi[d], lo[d]));
        bsss = __double2float_rn(__dadd_rn((double)bsss, __dmul_rn((double)ns, (double)ns)));
        if (ns > axsz) { axsz = ns; axis = d; }
    }
    a.t.nr[node] = __fmul_rn(0.5f, __fsqrt_rn(bsss));
    const uint32_t n = b.pl - b.pf;
    b.axis = axis;
    b.nless = b.pf + a.block * (1u << (31 - __clz((n - 1) / a.block)));
    b.ideal = __fdiv_rn(__uint2float_rn(b.nless - b.pf), __uint2float_rn(n));
    BigWin w; w.wf = b.pf; w.wl = b.pl - 1; w.lo = lo[axis]; w.hi = hi[axis]; w.done = 0; w.iters = 0; w.local = 0;
    w.pivot = select_pivot(b.nless, w.wf, w.wl, w.lo, w.hi, b.ideal, a.pivot_mode);
    b.w[0] = w; b.w[1] = w;
}

// what a chunk phase needs to know about its node, cached in shared memory while consecutive chunks of a CTA's range
// belong to the same node (the worklist keeps a node's chunks together): one dependent global load chain per node
// change instead of one per chunk
struct BigCache { uint32_t pf, pl, chunk0, B, k; int axis; BigWin w; };

// element range of this chunk inside the node's current window; false if empty / node finished
__device__ __forceinline__ bool big_range(const BigCache& b, uint32_t chunk, uint32_t& i0, uint32_t& i1) {
    if (b.w.done) return false;
    const uint32_t c0 = b.pf + (chunk - b.chunk0) * BIG_CH;
    i0 = max(c0, b.w.wf); i1 = min(min(b.pl, c0 + BIG_CH), b.w.wl + 1);
    return i0 < i1;
}

__device__ __forceinline__ void big_scan(const BigArgs& a, const int it, const uint32_t bi);

// per chunk: #{v < pivot}, max{v < pivot}, min{v >= pivot}; the chunk count goes to cntA, the node totals are
// accumulated in shared memory (s_acc) and flushed once per (CTA, node) by big_count_flush
__device__ __forceinline__ void big_count(const BigArgs& a, const BigCache& b, const uint32_t chunk, uint32_t* s_acc) {
    uint32_t i0, i1;
    if (!big_range(b, chunk, i0, i1)) return;
    const float* key = a.x[b.axis];
    const float pivot = b.w.pivot;
    const uint32_t c0 = b.pf + (chunk - b.chunk0) * BIG_CH;
    uint32_t cnt = 0; float mx = -INFINITY, mn = INFINITY;
    float v[BIG_ROUNDS]; bool ok[BIG_ROUNDS];
    #pragma unroll
    for (int r = 0; r < BIG_ROUNDS; ++r) {            // 8 independent loads in flight per thread
        const uint32_t i = c0 + (uint32_t)r * BIG_T + threadIdx.x;
        ok[r] = i >= i0 && i < i1;
        v[r] = ok[r] ? key[i] : 0.f;
    }
    #pragma unroll
    for (int r = 0; r < BIG_ROUNDS; ++r) if (ok[r]) { if (v[r] < pivot) { ++cnt; mx = fmaxf(mx, v[r]); } else mn = fminf(mn, v[r]); }
    __shared__ uint32_t s_c[BIG_T / 32], s_mx[BIG_T / 32], s_mn[BIG_T / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cnt = bw_sum(cnt);
    const uint32_t emx = __reduce_max_sync(0xffffffffu, f2ord(mx)), emn = __reduce_min_sync(0xffffffffu, f2ord(mn));
    if (lane == 0) { s_c[warp] = cnt; s_mx[warp] = emx; s_mn[warp] = emn; }
    __syncthreads();
    if (warp == 0) {
        const uint32_t c2 = bw_sum(lane < BIG_T / 32 ? s_c[lane] : 0u);
        const uint32_t x2 = __reduce_max_sync(0xffffffffu, lane < BIG_T / 32 ? s_mx[lane] : 0u);
        const uint32_t n2 = __reduce_min_sync(0xffffffffu, lane < BIG_T / 32 ? s_mn[lane] : 0xffffffffu);
        if (lane == 0) {
            a.cntA[chunk] = c2;                        // elements < pivot in this chunk's part of the window
            s_acc[0] += c2; s_acc[1] = max(s_acc[1], x2); s_acc[2] = min(s_acc[2], n2); s_acc[3] += 1u;
        }
    }
}

// once per (CTA, node): node totals by atomics, then the ticket - the last arrival of the node's window runs the node's
// scan right here (one grid-wide barrier less per pass). Release our results, take the ticket, acquire everybody else's.
__device__ __forceinline__ void big_count_flush(const BigArgs& a, const int it, const uint32_t bi, uint32_t* s_acc) {    // one CTA
    __shared__ uint32_t s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        BigNode& b = a.nodes[bi];
        s_last = 0u;
        if (s_acc[3]) {
            if (s_acc[0]) atomicAdd(&b.m, s_acc[0]);
            if (s_acc[1] > f2ord(-INFINITY)) atomicMax(&b.mx_enc, s_acc[1]);
            if (s_acc[2] < f2ord(INFINITY)) atomicMin(&b.mn_enc, s_acc[2]);
            __threadfence();
            s_last = (atomicAdd(&b.arrived, s_acc[3]) + s_acc[3] == b.expect) ? 1u : 0u;
        }
        s_acc[0] = 0u; s_acc[1] = f2ord(-INFINITY); s_acc[2] = f2ord(INFINITY); s_acc[3] = 0u;
    }
    __syncthreads();
    if (s_last) { __threadfence(); big_scan(a, it, bi); }
    __syncthreads();
}

// one CTA per big node. The misplaced counts of a chunk follow from its "< pivot" count and its position relative to
// B = wf + m (left of B: everything not "<" is misplaced; right of B: everything "<" is); only the one chunk that
// straddles B needs a second look at its keys. Then: exclusive offsets per chunk, k, and the reference's window update
// and exit rules (barneshut.hpp:565-585).
__device__ __forceinline__ void big_scan(const BigArgs& a, const int it, const uint32_t bi) {     // one CTA
    BigNode& b = a.nodes[bi];
    const BigWin w = b.w[it & 1];
    if (w.done) { if (threadIdx.x == 0) b.w[(it + 1) & 1] = w; return; }
    __shared__ uint32_t s_wa[8], s_wb[8], s_carry[2], s_ltl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t B = w.wf + *(volatile uint32_t*)&b.m;
    // the straddling chunk: #{v < pivot} among its window elements left of B
    {
        const uint32_t qb = (B - b.pf) / BIG_CH;
        const uint32_t c0 = b.pf + qb * BIG_CH;
        const uint32_t i0 = max(c0, w.wf), i1 = min(min(b.pl, c0 + BIG_CH), w.wl + 1);
        uint32_t ltl = 0;
        if (B > i0 && B < i1) {
            const float* key = a.x[b.axis];
            for (uint32_t i = i0 + threadIdx.x; i < B; i += 256) ltl += (key[i] < w.pivot);
        }
        ltl = bw_sum(ltl);
        if (lane == 0) s_wa[warp] = ltl;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (int q = 0; q < 8; ++q) t += s_wa[q]; s_ltl = t; s_carry[0] = 0; s_carry[1] = 0; }
        __syncthreads();
    }
    const uint32_t wq0 = (w.wf - b.pf) / BIG_CH, wq1 = (w.wl - b.pf) / BIG_CH;      // only the chunks of the window take part
    for (uint32_t base = wq0; base <= wq1; base += 256) {
        const uint32_t q = base + threadIdx.x;
        uint32_t va = 0, vb = 0;
        if (q <= wq1) {
            const uint32_t c0 = b.pf + q * BIG_CH;
            const uint32_t i0 = max(c0, w.wf), i1 = min(min(b.pl, c0 + BIG_CH), w.wl + 1);
            if (i0 < i1) {
                const uint32_t lt = __ldcg(&a.cntA[b.chunk0 + q]);
                if (i1 <= B) va = (i1 - i0) - lt;
                else if (i0 >= B) vb = lt;
                else { va = (B - i0) - s_ltl; vb = lt - s_ltl; }
            }
        }
        uint32_t ia = va, ib = vb;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o); if (lane >= o) { ia += ta; ib += tb; } }
        if (lane == 31) { s_wa[warp] = ia; s_wb[warp] = ib; }
        __syncthreads();
        uint32_t pa = s_carry[0], pb = s_carry[1], ta = 0, tb = 0;
        for (int q2 = 0; q2 < 8; ++q2) { if (q2 < warp) { pa += s_wa[q2]; pb += s_wb[q2]; } ta += s_wa[q2]; tb += s_wb[q2]; }
        if (q <= wq1) { a.cntA[b.chunk0 + q] = pa + ia - va; a.cntB[b.chunk0 + q] = pb + ib - vb; }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] += ta; s_carry[1] += tb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        b.B = B; b.k = s_carry[0];
        const float mx_lt = ord2f(*(volatile uint32_t*)&b.mx_enc), mn_ge = ord2f(*(volatile uint32_t*)&b.mn_enc);
        b.npass += 1; b.nscan += (w.wl - w.wf + 1);
        BigWin nw = w;
        if (B == b.nless) nw.done = 1;
        else {
            if (B < b.nless) { nw.wf = B; nw.lo = mn_ge; } else { nw.wl = B - 1; nw.hi = mx_lt; }
            if (nw.wf == w.wf && nw.wl == w.wl) { nw.done = 1; b.nstall += 1; }
            else {
                nw.iters = w.iters + 1;
                if (!(nw.wl > nw.wf) || nw.iters >= 100) nw.done = 1;                      // loop condition :527
                else {
                    nw.pivot = select_pivot(b.nless, nw.wf, nw.wl, nw.lo, nw.hi, b.ideal, a.pivot_mode);
                    // the remaining passes of a small window cost three grid-wide barriers each for a few microseconds of
                    // work: hand the node to one CTA instead (first chunk of the window, next count phase)
                    if (nw.wl - nw.wf + 1u <= BIG_LOCAL) nw.local = 1;
                }
            }
        }
        b.w[(it + 1) & 1] = nw;
        b.m = 0; b.mx_enc = 0u; b.mn_enc = 0xffffffffu; b.arrived = 0; b.expect = 0;
        if (nw.done) atomicSub(&a.nbig[2], 1u);
        // the chunks the next pass has to visit
        s_carry[0] = 0; s_carry[1] = 0;
        if (!nw.done) {
            const uint32_t q0 = (nw.wf - b.pf) / BIG_CH, q1 = nw.local ? q0 : (nw.wl - b.pf) / BIG_CH;
            s_carry[0] = q1 - q0 + 1; b.expect = q1 - q0 + 1;
            s_carry[1] = atomicAdd(&a.nbig[4 + ((it + 1) & 1)], q1 - q0 + 1);
            s_wa[0] = b.chunk0 + q0;
        }
    }
    __syncthreads();
    {
        const uint32_t nq = s_carry[0], base = s_carry[1], first = s_wa[0];
        uint32_t* out = a.wl + (size_t)((it + 1) & 1) * a.max_chunks + base;
        for (uint32_t j = threadIdx.x; j < nq; j += 256) out[j] = first + j;
    }
}

__device__ __forceinline__ void big_compact(const BigArgs& a, const BigCache& b, const uint32_t chunk) {
    uint32_t i0, i1;
    if (!big_range(b, chunk, i0, i1)) return;
    const uint32_t B = b.B;
    const float* key = a.x[b.axis];
    const float pivot = b.w.pivot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c0 = b.pf + (chunk - b.chunk0) * BIG_CH;
    uint32_t ba[BIG_ROUNDS], bb[BIG_ROUNDS], totA = 0, totB = 0;
    #pragma unroll
    for (int r = 0; r < BIG_ROUNDS; ++r) {
        const uint32_t i = c0 + (uint32_t)warp * 32u * BIG_ROUNDS + (uint32_t)r * 32u + lane;
        const bool valid = i >= i0 && i < i1;
        const float v = valid ? key[i] : 0.f;
        const bool lt = v < pivot;
        ba[r] = __ballot_sync(0xffffffffu, valid && i < B && !lt);
        bb[r] = __ballot_sync(0xffffffffu, valid && i >= B && lt);
        totA += __popc(ba[r]); totB += __popc(bb[r]);
    }
    __shared__ uint32_t s_a[BIG_T / 32], s_b[BIG_T / 32];
    if (lane == 0) { s_a[warp] = totA; s_b[warp] = totB; }
    __syncthreads();
    uint32_t offA = a.cntA[chunk], offB = a.cntB[chunk];
    for (int q = 0; q < warp; ++q) { offA += s_a[q]; offB += s_b[q]; }
    const uint32_t lt_mask = (1u << lane) - 1u;
    #pragma unroll
    for (int r = 0; r < BIG_ROUNDS; ++r) {
        const uint32_t i = c0 + (uint32_t)warp * 32u * BIG_ROUNDS + (uint32_t)r * 32u + lane;
        if ((ba[r] >> lane) & 1u) a.scr[b.pf + offA + __popc(ba[r] & lt_mask)] = i;
        if ((bb[r] >> lane) & 1u) a.scr[b.pl - 1u - (offB + __popc(bb[r] & lt_mask))] = i;
        offA += __popc(ba[r]); offB += __popc(bb[r]);
    }
}

__device__ __forceinline__ void big_swap(const BigArgs& a, const BigCache& b, const uint32_t chunk) {
    if (b.w.done) return;
    // the k swap pairs of the node are dealt out evenly over the chunks of its window (a pass misplaces about a quarter
    // of the window: handing each chunk "its" 2048 pairs would leave three quarters of the CTAs without work)
    const uint32_t q0 = (b.w.wf - b.pf) / BIG_CH, nq = (b.w.wl - b.pf) / BIG_CH - q0 + 1u;
    const uint32_t k = b.k, q = (chunk - b.chunk0) - q0;                              // my rank among the chunks of the node's window
    float* key = a.x[b.axis];
    const uint32_t j0 = (uint32_t)((unsigned long long)k * q / nq), j1 = (uint32_t)((unsigned long long)k * (q + 1u) / nq);
    for (uint32_t j = j0 + threadIdx.x; j < j1; j += BIG_T) {                             // barneshut.hpp:549-556
        const uint32_t pa = a.scr[b.pf + j], pb = a.scr[b.pl - k + j];
        const float va = key[pa], vb = key[pb]; key[pa] = vb; key[pb] = va;
        const uint32_t ia = a.lidx[pa], ib = a.lidx[pb]; a.lidx[pa] = ib; a.lidx[pb] = ia;
    }
}

// all remaining passes of a node whose window has become small, by this CTA alone (same passes, pivots and exit rules:
// cta_select is the loop k_node_split runs). Publishes the node as done in both window slots.
__device__ __forceinline__ void big_local_finish(const BigArgs& a, const int it, const uint32_t bi) {     // one CTA
    BigNode& b = a.nodes[bi];
    const BigWin w = b.w[it & 1];
    const SelStats st = cta_select(a.x[b.axis], a.lidx, a.scr, b.pf, b.pl, b.nless, w.wf, w.wl, w.lo, w.hi, b.ideal, w.iters, a.pivot_mode);
    if (threadIdx.x == 0) {
        b.npass += st.n_pass; b.nstall += st.n_stall; b.nscan += st.n_scan;
        BigWin d = w; d.done = 1;
        b.w[0] = d; b.w[1] = d;
        atomicSub(&a.nbig[2], 1u);
    }
    __syncthreads();
}

// one phase (0 count+scan, 1 compact, 2 swap) of one pass over this CTA's CONTIGUOUS share of the worklist
constexpr uint32_t BIG_PRE = 64;
template <int PHASE>
__device__ __forceinline__ void big_phase(const BigArgs& a, const int it, const uint32_t* wl, const uint32_t nw) {
    __shared__ uint32_t s_chunk[BIG_PRE], s_owner[BIG_PRE];
    __shared__ BigCache s_node;
    __shared__ uint32_t s_acc[4];
    const uint32_t i0 = (uint32_t)((unsigned long long)nw * blockIdx.x / gridDim.x);
    const uint32_t i1 = (uint32_t)((unsigned long long)nw * (blockIdx.x + 1) / gridDim.x);
    if (i0 >= i1) return;
    if (PHASE == 0 && threadIdx.x == 0) { s_acc[0] = 0u; s_acc[1] = f2ord(-INFINITY); s_acc[2] = f2ord(INFINITY); s_acc[3] = 0u; }
    uint32_t cur = 0xffffffffu;
    for (uint32_t base = i0; base < i1; base += BIG_PRE) {
        const uint32_t m = min(BIG_PRE, i1 - base);
        __syncthreads();
        if (threadIdx.x < m) { const uint32_t ch = wl[base + threadIdx.x]; s_chunk[threadIdx.x] = ch; s_owner[threadIdx.x] = a.chunk_owner[ch]; }
        __syncthreads();
        for (uint32_t j = 0; j < m; ++j) {
            const uint32_t bi = s_owner[j], chunk = s_chunk[j];
            if (bi != cur) {
                if (PHASE == 0 && cur != 0xffffffffu) big_count_flush(a, it, cur, s_acc);
                __syncthreads();
                if (threadIdx.x == 0) {
                    const BigNode& n = a.nodes[bi];
                    s_node.pf = n.pf; s_node.pl = n.pl; s_node.chunk0 = n.chunk0; s_node.axis = n.axis; s_node.w = n.w[it & 1];
                    s_node.B = n.B; s_node.k = n.k;
                }
                __syncthreads();
                cur = bi;
            }
            if (PHASE == 0 && s_node.w.local) big_local_finish(a, it, bi);        // (its single worklist entry; phases 1 and 2 then see done)
            else if (PHASE == 0) big_count(a, s_node, chunk, s_acc);
            else if (PHASE == 1) big_compact(a, s_node, chunk);
            else big_swap(a, s_node, chunk);
            __syncthreads();
        }
    }
    if (PHASE == 0 && cur != 0xffffffffu) big_count_flush(a, it, cur, s_acc);
}

// per node: publish the split (barneshut.hpp:702-704) and the statistics
__device__ __forceinline__ void big_finish(const BigArgs& a, const uint32_t bi) {     // one thread per node
    const BigNode& b = a.nodes[bi];
    const uint32_t node = b.node;
    a.axis_of[node] = (uint8_t)b.axis; a.pmid[node] = b.nless;
    a.t.ioffset[2 * node] = b.pf;        a.t.num[2 * node] = b.nless - b.pf;
    a.t.ioffset[2 * node + 1] = b.nless; a.t.num[2 * node + 1] = b.pl - b.nless;
    atomicAdd(&a.stats[0], 1ull); atomicAdd(&a.stats[1], (unsigned long long)b.npass);
    atomicAdd(&a.stats[2], (unsigned long long)b.nstall); atomicAdd(&a.stats[3], b.nscan);
}


// ---- one cooperative launch per level: every phase of every pass, separated by grid-wide barriers -----------------
// (one launch instead of ~60: at these sizes the passes are short enough that launch latency dominated them)
__global__ void __launch_bounds__(BIG_T, 4) k_big_level(const BigArgs a) {
    cg::grid_group grid = cg::this_grid();
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
    int nprof = 0;
    #define BIG_STAMP() do { if (a.prof && gtid == 0 && nprof < 510) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.prof[1 + nprof++] = t_; a.prof[0] = (unsigned long long)nprof; } } while (0)
    BIG_STAMP();
    if (blockIdx.x == 0) big_list(a);
    grid.sync();
    const uint32_t nnodes = a.nbig[0], nchunks = a.nbig[1];
    if (nnodes == 0) return;
    for (uint32_t bi = blockIdx.x; bi < nnodes; bi += gridDim.x) {
        const uint32_t c0 = a.nodes[bi].chunk0, nc = a.nodes[bi].nchunks;
        for (uint32_t q = threadIdx.x; q < nc; q += blockDim.x) a.chunk_owner[c0 + q] = bi;
    }
    grid.sync();
    BIG_STAMP();
    big_bbox_phase(a, nchunks);
    grid.sync();
    BIG_STAMP();
    for (uint32_t bi = gtid; bi < nnodes; bi += gthreads) big_setup(a, bi);
    for (uint32_t i = gtid; i < nchunks; i += gthreads) a.wl[i] = i;            // pass 0 visits every chunk
    if (gtid == 0) { a.nbig[4] = nchunks; a.nbig[5] = 0; }
    grid.sync();
    BIG_STAMP();
    for (int it = 0; it < 104; ++it) {
        const uint32_t* wl = a.wl + (size_t)(it & 1) * a.max_chunks;
        const uint32_t nw = a.nbig[4 + (it & 1)];
        // count, and per node (last chunk to arrive) the scan: next window, next pivot, next worklist (parity it+1)
        big_phase<0>(a, it, wl, nw);
        grid.sync();
        BIG_STAMP();
        big_phase<1>(a, it, wl, nw);
        grid.sync();
        BIG_STAMP();
        if (gtid == 0) a.nbig[4 + (it & 1)] = 0;          // this pass's worklist is consumed (everybody holds nw); pass it+1's scans refill it for pass it+2
        big_phase<2>(a, it, wl, nw);
        grid.sync();
        BIG_STAMP();
        if (a.nbig[2] == 0) break;        // every node of the level has met one of the reference's exit conditions
    }
    for (uint32_t bi = gtid; bi < nnodes; bi += gthreads) big_finish(a, bi);
}
