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

struct BigWin { uint32_t wf, wl; float lo, hi, pivot; int done, iters, pad; };
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

// per chunk: bounding box contribution and lidx = iota (barneshut.hpp:621-625, :516)
__device__ __forceinline__ void big_bbox(const BigArgs& a, const int it, const uint32_t chunk) {
    BigNode& b = a.nodes[a.chunk_owner[chunk]];
    const uint32_t i0 = b.pf + (chunk - b.chunk0) * BIG_CH, i1 = min(b.pl, i0 + BIG_CH);
    __shared__ float s_lo[BIG_T / 32], s_hi[BIG_T / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = i0 + threadIdx.x; i < i1; i += BIG_T) a.lidx[i] = i;
    for (int d = 0; d < a.PD; ++d) {
        float lo = INFINITY, hi = -INFINITY;
        float v[BIG_ROUNDS];
        #pragma unroll
        for (int r = 0; r < BIG_ROUNDS; ++r) { const uint32_t i = i0 + (uint32_t)r * BIG_T + threadIdx.x; v[r] = i < i1 ? a.x[d][i] : NAN; }
        #pragma unroll
        for (int r = 0; r < BIG_ROUNDS; ++r) { lo = fminf(lo, v[r]); hi = fmaxf(hi, v[r]); }      // fminf/fmaxf ignore the NaN fillers
        lo = bw_min(lo); hi = bw_max(hi);
        if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
        __syncthreads();
        if (warp == 0) {
            lo = lane < BIG_T / 32 ? s_lo[lane] : INFINITY; hi = lane < BIG_T / 32 ? s_hi[lane] : -INFINITY;
            lo = bw_min(lo); hi = bw_max(hi);
            if (lane == 0) { atomicMin(&b.bmin[d], f2ord(lo)); atomicMax(&b.bmax[d], f2ord(hi)); }
        }
        __syncthreads();
    }
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
    BigWin w; w.wf = b.pf; w.wl = b.pl - 1; w.lo = lo[axis]; w.hi = hi[axis]; w.done = 0; w.iters = 0; w.pad = 0;
    w.pivot = select_pivot(b.nless, w.wf, w.wl, w.lo, w.hi, b.ideal, a.pivot_mode);
    b.w[0] = w; b.w[1] = w;
}

// element range of this chunk inside the node's current window; false if empty / node finished
__device__ __forceinline__ bool big_range(const BigNode& b, const BigWin& w, uint32_t chunk, uint32_t& i0, uint32_t& i1) {
    if (w.done) return false;
    const uint32_t c0 = b.pf + (chunk - b.chunk0) * BIG_CH;
    i0 = max(c0, w.wf); i1 = min(min(b.pl, c0 + BIG_CH), w.wl + 1);
    return i0 < i1;
}

__device__ __forceinline__ void big_scan(const BigArgs& a, const int it, const uint32_t bi);

__device__ __forceinline__ void big_count(const BigArgs& a, const int it, const uint32_t chunk) {
    const uint32_t bi = a.chunk_owner[chunk];
    BigNode& b = a.nodes[bi];
    const BigWin w = b.w[it & 1];
    uint32_t i0, i1;
    if (!big_range(b, w, chunk, i0, i1)) return;
    const float* key = a.x[b.axis];
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
    for (int r = 0; r < BIG_ROUNDS; ++r) if (ok[r]) { if (v[r] < w.pivot) { ++cnt; mx = fmaxf(mx, v[r]); } else mn = fminf(mn, v[r]); }
    __shared__ uint32_t s_c[BIG_T / 32]; __shared__ float s_mx[BIG_T / 32], s_mn[BIG_T / 32];
    __shared__ uint32_t s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_last = 0u;
    cnt = bw_sum(cnt); mx = bw_max(mx); mn = bw_min(mn);
    if (lane == 0) { s_c[warp] = cnt; s_mx[warp] = mx; s_mn[warp] = mn; }
    __syncthreads();
    if (warp == 0) {
        cnt = lane < BIG_T / 32 ? s_c[lane] : 0u; mx = lane < BIG_T / 32 ? s_mx[lane] : -INFINITY; mn = lane < BIG_T / 32 ? s_mn[lane] : INFINITY;
        cnt = bw_sum(cnt); mx = bw_max(mx); mn = bw_min(mn);
        if (lane == 0) {
            a.cntA[chunk] = cnt;                       // elements < pivot in this chunk's part of the window
            if (cnt) atomicAdd(&b.m, cnt);
            if (mx > -INFINITY) atomicMax(&b.mx_enc, f2ord(mx));
            if (mn < INFINITY) atomicMin(&b.mn_enc, f2ord(mn));
            // ticket: the last chunk of this node's window to arrive runs the node's scan right here (one grid-wide
            // barrier less per pass). Release our results, take the ticket, acquire everybody else's.
            __threadfence();
            s_last = (atomicAdd(&b.arrived, 1u) + 1u == b.expect) ? 1u : 0u;
        }
    }
    __syncthreads();
    if (s_last) { __threadfence(); big_scan(a, it, bi); }
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
    for (uint32_t base = 0; base < b.nchunks; base += 256) {
        const uint32_t q = base + threadIdx.x;
        uint32_t va = 0, vb = 0;
        if (q < b.nchunks) {
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
        if (q < b.nchunks) { a.cntA[b.chunk0 + q] = pa + ia - va; a.cntB[b.chunk0 + q] = pb + ib - vb; }
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
                else nw.pivot = select_pivot(b.nless, nw.wf, nw.wl, nw.lo, nw.hi, b.ideal, a.pivot_mode);
            }
        }
        b.w[(it + 1) & 1] = nw;
        b.m = 0; b.mx_enc = 0u; b.mn_enc = 0xffffffffu; b.arrived = 0; b.expect = 0;
        if (nw.done) atomicSub(&a.nbig[2], 1u);
        // the chunks the next pass has to visit
        s_carry[0] = 0; s_carry[1] = 0;
        if (!nw.done) {
            const uint32_t q0 = (nw.wf - b.pf) / BIG_CH, q1 = (nw.wl - b.pf) / BIG_CH;
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

__device__ __forceinline__ void big_compact(const BigArgs& a, const int it, const uint32_t chunk) {
    BigNode& b = a.nodes[a.chunk_owner[chunk]];
    const BigWin w = b.w[it & 1];
    uint32_t i0, i1;
    if (!big_range(b, w, chunk, i0, i1)) return;
    const uint32_t B = b.B;
    const float* key = a.x[b.axis];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c0 = b.pf + (chunk - b.chunk0) * BIG_CH;
    uint32_t ba[BIG_ROUNDS], bb[BIG_ROUNDS], totA = 0, totB = 0;
    #pragma unroll
    for (int r = 0; r < BIG_ROUNDS; ++r) {
        const uint32_t i = c0 + (uint32_t)warp * 32u * BIG_ROUNDS + (uint32_t)r * 32u + lane;
        const bool valid = i >= i0 && i < i1;
        const float v = valid ? key[i] : 0.f;
        const bool lt = v < w.pivot;
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

__device__ __forceinline__ void big_swap(const BigArgs& a, const int it, const uint32_t chunk) {
    BigNode& b = a.nodes[a.chunk_owner[chunk]];
    const BigWin w = b.w[it & 1];
    if (w.done) return;
    const uint32_t k = b.k, q = (chunk - b.chunk0) - (w.wf - b.pf) / BIG_CH;      // my rank among the chunks of the node's window
    float* key = a.x[b.axis];
    const uint32_t j1 = min(k, (q + 1) * BIG_CH);
    for (uint32_t j = q * BIG_CH + threadIdx.x; j < j1; j += BIG_T) {                     // barneshut.hpp:549-556
        const uint32_t pa = a.scr[b.pf + j], pb = a.scr[b.pl - k + j];
        const float va = key[pa], vb = key[pb]; key[pa] = vb; key[pb] = va;
        const uint32_t ia = a.lidx[pa], ib = a.lidx[pb]; a.lidx[pa] = ib; a.lidx[pb] = ia;
    }
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
    if (blockIdx.x == 0) big_list(a);
    grid.sync();
    const uint32_t nnodes = a.nbig[0], nchunks = a.nbig[1];
    if (nnodes == 0) return;
    for (uint32_t bi = blockIdx.x; bi < nnodes; bi += gridDim.x) {
        const uint32_t c0 = a.nodes[bi].chunk0, nc = a.nodes[bi].nchunks;
        for (uint32_t q = threadIdx.x; q < nc; q += blockDim.x) a.chunk_owner[c0 + q] = bi;
    }
    grid.sync();
    for (uint32_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) { big_bbox(a, 0, ch); __syncthreads(); }
    grid.sync();
    for (uint32_t bi = gtid; bi < nnodes; bi += gthreads) big_setup(a, bi);
    for (uint32_t i = gtid; i < nchunks; i += gthreads) a.wl[i] = i;            // pass 0 visits every chunk
    if (gtid == 0) { a.nbig[4] = nchunks; a.nbig[5] = 0; }
    grid.sync();
    for (int it = 0; it < 104; ++it) {
        const uint32_t* wl = a.wl + (size_t)(it & 1) * a.max_chunks;
        const uint32_t nw = a.nbig[4 + (it & 1)];
        // count, and per node (last chunk to arrive) the scan: next window, next pivot, next worklist (parity it+1)
        for (uint32_t i = blockIdx.x; i < nw; i += gridDim.x) { big_count(a, it, wl[i]); __syncthreads(); }
        grid.sync();
        for (uint32_t i = blockIdx.x; i < nw; i += gridDim.x) { big_compact(a, it, wl[i]); __syncthreads(); }
        grid.sync();
        if (gtid == 0) a.nbig[4 + (it & 1)] = 0;          // this pass's worklist is consumed (everybody holds nw); pass it+1 refills it for it+2
        for (uint32_t i = blockIdx.x; i < nw; i += gridDim.x) { big_swap(a, it, wl[i]); __syncthreads(); }
        grid.sync();
        if (a.nbig[2] == 0) break;        // every node of the level has met one of the reference's exit conditions
    }
    for (uint32_t bi = gtid; bi < nnodes; bi += gthreads) big_finish(a, bi);
}
