/*
 * tree.cu - VAM-split k-d tree build, node summaries and in-leaf refinement on the GPU.
 *
 * Replaces (reference src/barneshut.hpp): makeTree :814-854, splitNode :594-712, partialSortIndexes :505-587,
 * minMaxValue :423-454, reorder :475-500, Parts::reorder_idx Parts.hpp:188-196, finishTree :717-807,
 * refineTree/refineLeaf :860-936.
 *
 * The reference builds the tree by recursive partial selection on the longest axis of the node's tight bounding
 * box. There is no Morton key and no radix sort in it, and the particle order that comes out (which the parity
 * contract requires bit-for-bit) is the composition of its Hoare partition passes. This file therefore runs the
 * SAME passes, level-synchronously and in closed form (proven equal to the recursive
 * two-pointer loop by the test suite's CPU restatement against the compiled reference):
 *
 *   pass:  m = #{v < pivot in window}; B = wf + m;
 *          a_1 < a_2 < ...  positions in [wf,B) holding v >= pivot,   b_1 > b_2 > ...  positions in [B,wl] holding v < pivot
 *          swap (v, idx) at a_j <-> b_j for all j                      (two ordered stream compactions + one scatter)
 *
 * One CTA owns one tree node per level: bounding box (block min/max), pivot arithmetic in the reference's exact IEEE
 * sequence (double for the 9:1 blend), count / ordered-compaction / swap passes until the reference's own exit
 * conditions (done, stall, window of one), then one element-parallel gather kernel per level permutes the remaining
 * planes. All traffic is HBM/L2 streaming: the roofline for this file is memory bandwidth.
 */
#include "onb_internal.h"
#include <cstdlib>
#include <cstdio>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;


namespace {

// order-preserving float <-> uint encoding (min/max of floats as integer min/max: atomics and redux.sync)
__device__ __forceinline__ uint32_t f2ord(float f) { const uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }
// warp reductions in ONE instruction each (redux.sync, sm_80+) instead of five shuffle steps: the select passes are
// chains of short reductions, their latency is what bounds the small-node kernels
__device__ __forceinline__ float warp_min(float v) { return ord2f(__reduce_min_sync(0xffffffffu, f2ord(v))); }
__device__ __forceinline__ float warp_max(float v) { return ord2f(__reduce_max_sync(0xffffffffu, f2ord(v))); }
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }

__device__ __forceinline__ uint32_t log_2(uint32_t x) { return x == 0 ? 0u : 31u - __clz(x); }   // Tree.hpp:30-33

// barneshut.hpp:538-540 in the reference's exact IEEE sequence (mode 0) or as g++ -O3 -ffast-math contracts it (mode 1)
__device__ __forceinline__ float select_pivot(uint32_t nless, uint32_t wf, uint32_t wl, float lo, float hi, float ideal, int mode) {
    const float f0 = __fdiv_rn(__double2float_rn(__dsub_rn(__dsub_rn((double)nless, 0.5), (double)wf)), __uint2float_rn(wl - wf));
    if (mode == 0) {
        const float frac = __double2float_rn(__ddiv_rn(__dadd_rn(__dmul_rn(9.0, (double)f0), __dmul_rn(1.0, (double)ideal)), 10.0));
        return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), frac));
    }
    const float frac = __double2float_rn(__dmul_rn(__fma_rn(9.0, (double)f0, (double)ideal), 0.1));
    return __fmaf_rn(__fsub_rn(hi, lo), frac, lo);
}
constexpr int SPLIT_ROUNDS = 4;   // each warp handles 4 x 32 consecutive elements per chunk

// ---- the select passes of ONE node by ONE CTA, from any window state (barneshut.hpp:527-586) ----
// Used by k_node_split from the full window, and by the big-level kernel to finish a node locally once its window has
// become small (tree_big.cuh). Every thread tracks the (identical) window state in registers.
struct SelStats { uint32_t n_pass, n_stall; unsigned long long n_scan; };
__device__ __forceinline__ SelStats cta_select(float* key, uint32_t* lidx, uint32_t* scr, const uint32_t pf, const uint32_t pl, const uint32_t nless,
                                               uint32_t wf, uint32_t wl, float lo, float hi, const float ideal, int iters, const int pivot_mode) {
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
    __shared__ uint32_t s_cnt[32], s_cnt2[32];
    __shared__ float s_max[32], s_min[32];
    __shared__ float s_piv;
    uint32_t n_pass = 0, n_stall = 0; unsigned long long n_scan = 0;
    // the pivot (:538-540: double arithmetic, software division) is evaluated by warp 0 alone and published across a barrier
    bool go = wl > wf && iters < 100;
    if (warp == 0 && go) { const float pv = select_pivot(nless, wf, wl, lo, hi, ideal, pivot_mode); if (lane == 0) s_piv = pv; }
    __syncthreads();
    float pivot = go ? s_piv : 0.f;
    while (go) {
        // pass 1: m = #{v < pivot}, and the min/max the two possible next windows will have
        uint32_t cnt = 0; float mx_lt = -INFINITY, mn_ge = INFINITY;
        for (uint32_t i = wf + tid; i <= wl; i += T) {
            const float v = key[i];
            if (v < pivot) { ++cnt; mx_lt = fmaxf(mx_lt, v); } else mn_ge = fminf(mn_ge, v);
        }
        cnt = warp_sum(cnt); mx_lt = warp_max(mx_lt); mn_ge = warp_min(mn_ge);
        if (lane == 0) { s_cnt[warp] = cnt; s_max[warp] = mx_lt; s_min[warp] = mn_ge; }
        __syncthreads();
        {
            uint32_t c2 = lane < W ? s_cnt[lane] : 0u; float a2 = lane < W ? s_max[lane] : -INFINITY; float b2 = lane < W ? s_min[lane] : INFINITY;
            cnt = warp_sum(c2); mx_lt = warp_max(a2); mn_ge = warp_min(b2);
        }
        __syncthreads();
        const uint32_t B = wf + cnt;

        // pass 2: ordered compaction of the misplaced positions; a_j ascending from the front of scr[pf..],
        // b ascending-rank jb stored at scr[pl-1-jb] so that b_j (descending rank j) sits at scr[pl-k+j]
        uint32_t carryA = 0, carryB = 0;
        const uint32_t chunk = (uint32_t)W * 32u * SPLIT_ROUNDS;
        for (uint32_t base = wf; base <= wl; base += chunk) {
            uint32_t ba[SPLIT_ROUNDS], bb[SPLIT_ROUNDS];
            uint32_t totA = 0, totB = 0;
            #pragma unroll
            for (int r = 0; r < SPLIT_ROUNDS; ++r) {
                const uint32_t i = base + (uint32_t)warp * 32u * SPLIT_ROUNDS + (uint32_t)r * 32u + lane;
                const bool valid = i <= wl;
                const float v = valid ? key[i] : 0.f;
                const bool lt = v < pivot;
                ba[r] = __ballot_sync(0xffffffffu, valid && i < B && !lt);
                bb[r] = __ballot_sync(0xffffffffu, valid && i >= B && lt);
                totA += __popc(ba[r]); totB += __popc(bb[r]);
            }
            if (lane == 0) { s_cnt[warp] = totA; s_cnt2[warp] = totB; }
            __syncthreads();
            const uint32_t ca = lane < W ? s_cnt[lane] : 0u, cb = lane < W ? s_cnt2[lane] : 0u;
            const uint32_t allA = warp_sum(ca), allB = warp_sum(cb);
            const uint32_t exA = warp_sum(lane < warp ? ca : 0u), exB = warp_sum(lane < warp ? cb : 0u);
            uint32_t offA = carryA + exA, offB = carryB + exB;
            const uint32_t lt_mask = (1u << lane) - 1u;
            #pragma unroll
            for (int r = 0; r < SPLIT_ROUNDS; ++r) {
                const uint32_t i = base + (uint32_t)warp * 32u * SPLIT_ROUNDS + (uint32_t)r * 32u + lane;
                if ((ba[r] >> lane) & 1u) scr[pf + offA + __popc(ba[r] & lt_mask)] = i;
                if ((bb[r] >> lane) & 1u) scr[pl - 1u - (offB + __popc(bb[r] & lt_mask))] = i;
                offA += __popc(ba[r]); offB += __popc(bb[r]);
            }
            carryA += allA; carryB += allB;
            __syncthreads();
        }
        const uint32_t k = carryA;    // == carryB
        // pass 3: the swaps :549-556
        for (uint32_t j = tid; j < k; j += T) {
            const uint32_t pa = scr[pf + j], pb = scr[pl - k + j];
            const float va = key[pa], vb = key[pb]; key[pa] = vb; key[pb] = va;
            const uint32_t ia = lidx[pa], ib = lidx[pb]; lidx[pa] = ib; lidx[pb] = ia;
        }
        ++n_pass; n_scan += (wl - wf + 1);
        // :565-583 - the next window follows from B and the two extrema, all known since the count
        if (B == nless) go = false;
        else {
            const uint32_t owf = wf, owl = wl;
            if (B < nless) { wf = B; lo = mn_ge; } else { wl = B - 1; hi = mx_lt; }
            if (wf == owf && wl == owl) { ++n_stall; go = false; }
            else { ++iters; go = wl > wf && iters < 100; }                                        // loop condition :527
        }
        if (go && warp == 0) { const float pv = select_pivot(nless, wf, wl, lo, hi, ideal, pivot_mode); if (lane == 0) s_piv = pv; }
        __syncthreads();                             // swaps done; next pivot published
        if (go) pivot = s_piv;
    }
    SelStats st; st.n_pass = n_pass; st.n_stall = n_stall; st.n_scan = n_scan;
    return st;
}

struct SplitArgs {
    float* x[3];            // current coordinate planes (the select permutes x[axis] in place)
    TreeView t;
    uint8_t* axis_of;       // per node: split axis
    uint32_t* pmid;         // per node: first index of the right child
    uint32_t* lidx;         // per particle: position before this level's select (barneshut.hpp:516)
    uint32_t* scr;          // per particle scratch for the ordered compactions
    unsigned long long* stats;   // selects, passes, stalls, scanned
    uint32_t block, big;    // nodes with more than `big` particles are split by the grid-wide kernels below
    uint32_t blo, bhi;      // build range: only nodes overlapping particles [blo,bhi) are split (multi-GPU: each rank its share)
    int level, PD, pivot_mode;
};


// ---- one CTA = one node of this level -----------------------------------------------------
__global__ void __launch_bounds__(1024) k_node_split(const SplitArgs a) {
    const uint32_t node = (1u << a.level) + blockIdx.x;
    const uint32_t n = a.t.num[node];
    if (n == 0) return;
    const uint32_t pf = a.t.ioffset[node], pl = pf + n;
    if (!(pf < a.bhi && pl > a.blo)) {
        // outside this rank's build range: not sorted here, but the shape of the tree below it is data independent
        if (threadIdx.x == 0 && n > a.block) {
            const uint32_t pm = pf + a.block * (1u << log_2((n - 1) / a.block));
            a.t.ioffset[2 * node] = pf;     a.t.num[2 * node] = pm - pf;
            a.t.ioffset[2 * node + 1] = pm; a.t.num[2 * node + 1] = pl - pm;
        }
        return;
    }
    if (n > a.big) return;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;

    __shared__ float s_min[32], s_max[32];
    __shared__ uint32_t s_cnt[32], s_cnt2[32];
    __shared__ float s_lo[3], s_hi[3];

    // bounding box :621-625 (exact, order independent)
    for (int d = 0; d < a.PD; ++d) {
        const float* xd = a.x[d];
        float lo = INFINITY, hi = -INFINITY;
        for (uint32_t i = pf + tid; i < pl; i += T) { const float v = xd[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
        lo = warp_min(lo); hi = warp_max(hi);
        if (lane == 0) { s_min[warp] = lo; s_max[warp] = hi; }
        __syncthreads();
        if (warp == 0) {
            lo = lane < W ? s_min[lane] : INFINITY; hi = lane < W ? s_max[lane] : -INFINITY;
            lo = warp_min(lo); hi = warp_max(hi);
            if (lane == 0) { s_lo[d] = lo; s_hi[d] = hi; }
        }
        __syncthreads();
    }
    if (tid == 0) {
        float bsss = 0.0f;
        for (int d = 0; d < a.PD; ++d) {
            const float ns = __fsub_rn(s_hi[d], s_lo[d]);
            a.t.ns[d][node] = ns;
            a.t.nc[d][node] = __fmul_rn(0.5f, __fadd_rn(s_hi[d], s_lo[d]));
            bsss = __double2float_rn(__dadd_rn((double)bsss, __dmul_rn((double)ns, (double)ns)));   // std::pow(float,int) is double :638
        }
        a.t.nr[node] = __fmul_rn(0.5f, __fsqrt_rn(bsss));
    }
    if (n <= a.block) return;                                                            // :644 leaf

    // longest axis :652-659 (first strict maximum)
    int axis = 0; float axsz = -1.0f;
    for (int d = 0; d < a.PD; ++d) { const float ns = __fsub_rn(s_hi[d], s_lo[d]); if (ns > axsz) { axsz = ns; axis = d; } }
    const uint32_t nless = pf + a.block * (1u << log_2((n - 1) / a.block));              // :663
    float* key = a.x[axis];
    uint32_t* lidx = a.lidx;
    uint32_t* scr = a.scr;

    for (uint32_t i = pf + tid; i < pl; i += T) lidx[i] = i;                             // :516
    __syncthreads();

    // partial select :519-586
    const float ideal = __fdiv_rn(__uint2float_rn(nless - pf), __uint2float_rn(pl - pf));     // :522
    const SelStats st = cta_select(key, lidx, scr, pf, pl, nless, pf, pl - 1, s_lo[axis], s_hi[axis], ideal, 0, a.pivot_mode);
    const uint32_t n_pass = st.n_pass, n_stall = st.n_stall; const unsigned long long n_scan = st.n_scan;
    if (tid == 0) {
        a.axis_of[node] = (uint8_t)axis; a.pmid[node] = nless;
        a.t.ioffset[2 * node] = pf;        a.t.num[2 * node] = nless - pf;               // :702-704
        a.t.ioffset[2 * node + 1] = nless; a.t.num[2 * node + 1] = pl - nless;
        atomicAdd(&a.stats[0], 1ull); atomicAdd(&a.stats[1], (unsigned long long)n_pass);
        atomicAdd(&a.stats[2], (unsigned long long)n_stall); atomicAdd(&a.stats[3], n_scan);
    }
}

// ---- per level: permute the other coordinate planes and the composed index plane by lidx (reorder :475-485,
// reorder_idx Parts.hpp:188-196). Radii and strengths do not travel with the levels: they are permuted once at the
// end of the build through the composed index (k_apply_perm in tree_sub.cuh).
struct GatherArgs {
    float* sx[3]; uint32_t* sg;     // current
    float* dx[3]; uint32_t* dg;     // destination
    const uint32_t* lidx; uint32_t* owner;
    const uint8_t* axis_of; const uint32_t* pmid; const uint32_t* num; const uint32_t* ioffset;
    uint32_t n, block, blo, bhi, span_lo; int level, PD;
};
__global__ void k_gather(const GatherArgs a) {
    const uint32_t i = a.span_lo + blockIdx.x * blockDim.x + threadIdx.x;      // n = end of the span covered at this level
    if (i >= a.n) return;
    const uint32_t node = a.owner[i];
    bool active = node >= (1u << a.level) && a.num[node] > a.block;
    if (active) { const uint32_t pf = a.ioffset[node]; active = pf < a.bhi && pf + a.num[node] > a.blo; }
    const uint32_t j = active ? a.lidx[i] : i;
    const int ax = active ? (int)a.axis_of[node] : -1;
    for (int d = 0; d < a.PD; ++d) a.dx[d][i] = a.sx[d][d == ax ? i : j];
    a.dg[i] = a.sg[j];
    if (active) a.owner[i] = 2u * node + (i >= a.pmid[node] ? 1u : 0u);
}

__global__ void k_fill_u32(uint32_t* p, uint32_t n, uint32_t v, int iota) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = iota ? i : v;
}

// ---- finishTree leaves :753-806: one warp per node, one lane per summed quantity, sequential double sums ----
struct FinishArgs { PartsView p; TreeView t; uint32_t block; int PD, SD, are_sources; };

__global__ void k_finish_leaves(const FinishArgs a) {
    const uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (node >= (uint32_t)a.t.numnodes || node == 0) return;
    const uint32_t n = a.t.num[node];
    if (n == 0 || n > a.block) return;
    const uint32_t pf = a.t.ioffset[node], pl = pf + n;
    const int PD = a.PD, SD = a.are_sources ? a.SD : 0;
    // lane roles: [0,PD) weighted coordinate sums; PD weight sum; (PD, PD+SD] strength sums; PD+SD+1 radius sum
    double acc = 0.0;
    if (lane <= PD) {
        for (uint32_t i = pf; i < pl; ++i) {
            float w;
            if (!a.are_sources) w = 1.0f;
            else if (a.SD == 1) w = fabsf(a.p.s[0][i]);
            else {
                w = 0.0f;
                for (int d = 0; d < a.SD; ++d) { const double sd = (double)a.p.s[d][i]; w = __double2float_rn(__dadd_rn((double)w, __dmul_rn(sd, sd))); }   // :775
                w = __fsqrt_rn(w);
            }
            if (lane < PD) acc = __dadd_rn(acc, (double)__fmul_rn(a.p.x[lane][i], w));     // :790 float product, double sum
            else           acc = __dadd_rn(acc, (double)w);                                // :786
        }
    } else if (lane <= PD + SD) {
        const float* __restrict__ sd = a.p.s[lane - PD - 1];
        for (uint32_t i = pf; i < pl; ++i) acc = __dadd_rn(acc, (double)sd[i]);            // :796
    } else if (lane == PD + SD + 1) {
        for (uint32_t i = pf; i < pl; ++i) acc = __dadd_rn(acc, (double)a.p.r[i]);         // :800
    }
    const double wsum = __shfl_sync(0xffffffffu, acc, PD);
    if (lane < PD) {
        const float ooass = __double2float_rn(__ddiv_rn(1.0, __dadd_rn(1.e-20, wsum)));
        a.t.x[lane][node] = __double2float_rn(__dmul_rn((double)ooass, acc));
    } else if (lane > PD && lane <= PD + SD) {
        a.t.s[lane - PD - 1][node] = __double2float_rn(acc);
    } else if (lane == PD + SD + 1) {
        a.t.pr[node] = __fdiv_rn(__double2float_rn(acc), __uint2float_rn(n));              // :801
    }
}

// ---- finishTree parents :721-746, one level per launch (children first) ----
__global__ void k_finish_parents(const FinishArgs a, int level) {
    const uint32_t node = (1u << level) + blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= (2u << level)) return;
    const uint32_t n = a.t.num[node];
    if (n <= a.block) return;
    const uint32_t c1 = 2 * node, c2 = c1 + 1;
    const float n1 = __uint2float_rn(a.t.num[c1]), n2 = __uint2float_rn(a.t.num[c2]);
    const float oonp = __fdiv_rn(1.0f, __uint2float_rn(a.t.num[c1] + a.t.num[c2]));
    for (int d = 0; d < a.PD; ++d)
        a.t.x[d][node] = __fmul_rn(oonp, __fadd_rn(__fmul_rn(n1, a.t.x[d][c1]), __fmul_rn(n2, a.t.x[d][c2])));
    for (int d = 0; d < a.SD; ++d) a.t.s[d][node] = __fadd_rn(a.t.s[d][c1], a.t.s[d][c2]);
    a.t.pr[node] = __fmul_rn(oonp, __fadd_rn(__fmul_rn(n1, a.t.pr[c1]), __fmul_rn(n2, a.t.pr[c2])));
}

// ---------------------------------------------------------------------------------------------
// refineLeaf :860-895. One CTA per leaf (leaf j = particles [j*b, min((j+1)*b, n)): the VAM split gives
// every left child a perfect subtree, so all leaves but the last hold exactly b particles).
// Each recursion level sorts every segment on its longest axis. Without equal keys any sort gives the
// same order, so ranks are counted in parallel; a segment that does contain equal keys is re-sorted by
// one thread with libstdc++ 13's introsort restated literally (std::sort is unstable; the reference's tie
// order is whatever that algorithm leaves - bits/stl_algo.h: threshold 16, median-of-3 to first,
// unguarded partition, final insertion sort).
// ---------------------------------------------------------------------------------------------
struct IdxKey { const float* v; __device__ bool lt(int a, int b) const { return v[a] < v[b]; } };

__device__ void dev_insertion_unguarded(int* last, IdxKey K) {
    const int val = *last; int* next = last - 1;
    while (K.lt(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
__device__ void dev_insertion_sort(int* first, int* last, IdxKey K) {
    if (first == last) return;
    for (int* i = first + 1; i != last; ++i) {
        if (K.lt(*i, *first)) { const int val = *i; for (int* q = i; q != first; --q) *q = *(q - 1); *first = val; }
        else dev_insertion_unguarded(i, K);
    }
}
__device__ bool dev_libstdcxx_sort(int* first, int* last, IdxKey K) {
    const int n = (int)(last - first);
    if (n == 0) return true;
    // __introsort_loop with an explicit stack of pending right parts
    int stk_f[40], stk_l[40], stk_d[40]; int sp = 0;
    stk_f[0] = 0; stk_l[0] = n; stk_d[0] = 2 * (int)log_2((uint32_t)n); sp = 1;
    bool ok = true;
    while (sp > 0) {
        --sp; int* f = first + stk_f[sp]; int* l = first + stk_l[sp]; int depth = stk_d[sp];
        while (l - f > 16) {
            if (depth == 0) { ok = false; break; }      // heapsort fallback of libstdc++: not restated (needs 2 log2 n bad splits)
            --depth;
            int* mid = f + (l - f) / 2;
            int* a = f + 1; int* b = mid; int* c = l - 1;   // __move_median_to_first(f, f+1, mid, l-1)
            int* med;
            if (K.lt(*a, *b)) { if (K.lt(*b, *c)) med = b; else if (K.lt(*a, *c)) med = c; else med = a; }
            else if (K.lt(*a, *c)) med = a; else if (K.lt(*b, *c)) med = c; else med = b;
            { const int tmp = *f; *f = *med; *med = tmp; }
            int* lo = f + 1; int* hi = l;                   // __unguarded_partition(f+1, l, f)
            while (true) {
                while (K.lt(*lo, *f)) ++lo;
                --hi;
                while (K.lt(*f, *hi)) --hi;
                if (!(lo < hi)) break;
                const int tmp = *lo; *lo = *hi; *hi = tmp;
                ++lo;
            }
            if (sp < 40) { stk_f[sp] = (int)(lo - first); stk_l[sp] = (int)(l - first); stk_d[sp] = depth; ++sp; } else ok = false;
            l = lo;
        }
    }
    // __final_insertion_sort
    if (n > 16) { dev_insertion_sort(first, first + 16, K); for (int* i = first + 16; i != last; ++i) dev_insertion_unguarded(i, K); }
    else dev_insertion_sort(first, last, K);
    return ok;
}

struct RefineArgs { PartsView p; uint32_t n, block, leaf0; int PD, SD, OD, are_sources; int* flag; unsigned long long* tie_sorts; };

__global__ void __launch_bounds__(128) k_refine(const RefineArgs a) {
    const uint32_t pf = (a.leaf0 + blockIdx.x) * a.block;
    const int n = (int)min(a.block, a.n - pf);
    const int tid = threadIdx.x;
    __shared__ float sx[3][128];
    __shared__ int sidx[128];
    __shared__ int perm[128];
    __shared__ int segtie[128];
    __shared__ float keys[128];
    __shared__ float s_wb[4][6];
    const bool act = tid < n;
    if (act) { for (int d = 0; d < a.PD; ++d) sx[d][tid] = a.p.x[d][pf + tid]; perm[tid] = tid; }
    int sa = 0, sb = n;             // my segment [sa, sb)
    __syncthreads();
    while (true) {
        const bool work = act && (sb - sa) >= 3;                                          // :864
        if (!__syncthreads_or(work ? 1 : 0)) break;
        int axis = 0; int rank = 0; bool tie = false;
        if (act) segtie[tid] = 0;
        float bs[3] = {0.f, 0.f, 0.f};
        if (n == 128) {
            // a full leaf (all but the last one): every segment of this recursion level is an aligned block of w = sb - sa
            // threads, so the boxes come from warp reductions / xor shuffles instead of one loop over the segment per thread
            const int w = sb - sa, lane = tid & 31, warp = tid >> 5;
            if (w >= 3) {
                float lo[3], hi[3];
                #pragma unroll
                for (int d = 0; d < 3; ++d) { const float v = d < a.PD ? sx[d][tid] : 0.f; lo[d] = v; hi[d] = v; }
                if (w >= 32) {
                    #pragma unroll
                    for (int d = 0; d < 3; ++d) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
                    if (w > 32) {
                        if (lane == 0) {
                            #pragma unroll
                            for (int d = 0; d < 3; ++d) { s_wb[warp][d] = lo[d]; s_wb[warp][3 + d] = hi[d]; }
                        }
                        __syncthreads();
                        const int w0 = sa >> 5, nw = w >> 5;
                        #pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            lo[d] = s_wb[w0][d]; hi[d] = s_wb[w0][3 + d];
                            for (int q = 1; q < nw; ++q) { lo[d] = fminf(lo[d], s_wb[w0 + q][d]); hi[d] = fmaxf(hi[d], s_wb[w0 + q][3 + d]); }
                        }
                    }
                } else {
                    for (int o = w >> 1; o; o >>= 1) {
                        #pragma unroll
                        for (int d = 0; d < 3; ++d) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
                    }
                }
                #pragma unroll
                for (int d = 0; d < 3; ++d) bs[d] = __fsub_rn(hi[d], lo[d]);
            }
        } else if (work) {
            for (int d = 0; d < a.PD; ++d) {
                float lo = sx[d][sa], hi = lo;
                for (int j = sa; j < sb; ++j) { const float v = sx[d][j]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
                bs[d] = __fsub_rn(hi, lo);
            }
        }
        if (work) {
            for (int d = 1; d < a.PD; ++d) if (bs[axis] < bs[d]) axis = d;                // std::max_element :878
            const float key = sx[axis][tid];
            for (int j = sa; j < sb; ++j) { const float v = sx[axis][j]; rank += (v < key); tie |= (v == key && j != tid); }
            keys[tid] = key;
        }
        __syncthreads();
        if (work && tie) segtie[sa] = 1;
        __syncthreads();
        if (work) {
            if (!segtie[sa]) sidx[sa + rank] = tid;
            else if (tid == sa) {
                for (int j = sa; j < sb; ++j) sidx[j] = j;                                // :408
                IdxKey K; K.v = keys;
                if (!dev_libstdcxx_sort(sidx + sa, sidx + sb, K)) atomicExch(a.flag, ONB_ERR_UNSUPPORTED);
                atomicAdd(a.tie_sorts, 1ull);
            }
        }
        __syncthreads();
        float nx[3]; int np = 0;
        if (work) { const int src = sidx[tid]; for (int d = 0; d < a.PD; ++d) nx[d] = sx[d][src]; np = perm[src]; }
        __syncthreads();
        if (work) {
            for (int d = 0; d < a.PD; ++d) sx[d][tid] = nx[d];
            perm[tid] = np;
            const int pm = sa + (1 << log_2((uint32_t)(sb - sa - 1)));                    // :890
            if (tid < pm) sb = pm; else sa = pm;
        }
        __syncthreads();
    }
    // write back: coordinates from shared memory, the other planes through the composed permutation
    float vr = 0.f, vs[3] = {0.f, 0.f, 0.f}; uint32_t vg = 0;
    if (act) {
        const uint32_t src = pf + perm[tid];
        vr = a.p.r[src];
        if (a.are_sources) for (int d = 0; d < a.SD; ++d) vs[d] = a.p.s[d][src];
        if (a.p.gidx) vg = a.p.gidx[src];
    }
    __syncthreads();
    if (act) {
        for (int d = 0; d < a.PD; ++d) a.p.x[d][pf + tid] = sx[d][tid];
        a.p.r[pf + tid] = vr;
        if (a.are_sources) for (int d = 0; d < a.SD; ++d) a.p.s[d][pf + tid] = vs[d];
        if (a.p.gidx) a.p.gidx[pf + tid] = vg;
    }
}

// ---- bottom-up bounding boxes (multi-GPU: after the exchange of a range-restricted build) ----
__global__ void k_bbox_leaves(const FinishArgs a, float* lohi) {
    const uint32_t node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (node >= (uint32_t)a.t.numnodes || node == 0) return;
    const uint32_t n = a.t.num[node];
    if (n == 0 || n > a.block) return;
    const uint32_t pf = a.t.ioffset[node], pl = pf + n;
    const size_t nn = a.t.numnodes;
    for (int d = 0; d < a.PD; ++d) {
        float lo = INFINITY, hi = -INFINITY;
        for (uint32_t i = pf + lane; i < pl; i += 32) { const float v = a.p.x[d][i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
        lo = warp_min(lo); hi = warp_max(hi);
        if (lane == 0) { lohi[(size_t)(2 * d) * nn + node] = lo; lohi[(size_t)(2 * d + 1) * nn + node] = hi; }
    }
}
__global__ void k_bbox_parents(const FinishArgs a, float* lohi, int level) {
    const uint32_t node = (1u << level) + blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= (2u << level)) return;
    if (a.t.num[node] <= a.block) return;
    const size_t nn = a.t.numnodes;
    for (int d = 0; d < a.PD; ++d) {
        lohi[(size_t)(2 * d) * nn + node] = fminf(lohi[(size_t)(2 * d) * nn + 2 * node], lohi[(size_t)(2 * d) * nn + 2 * node + 1]);
        lohi[(size_t)(2 * d + 1) * nn + node] = fmaxf(lohi[(size_t)(2 * d + 1) * nn + 2 * node], lohi[(size_t)(2 * d + 1) * nn + 2 * node + 1]);
    }
}
__global__ void k_bbox_final(const FinishArgs a, const float* lohi) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node == 0 || node >= (uint32_t)a.t.numnodes || a.t.num[node] == 0) return;
    const size_t nn = a.t.numnodes;
    float bsss = 0.0f;
    for (int d = 0; d < a.PD; ++d) {
        const float lo = lohi[(size_t)(2 * d) * nn + node], hi = lohi[(size_t)(2 * d + 1) * nn + node];
        const float ns = __fsub_rn(hi, lo);
        a.t.ns[d][node] = ns;
        a.t.nc[d][node] = __fmul_rn(0.5f, __fadd_rn(hi, lo));
        bsss = __double2float_rn(__dadd_rn((double)bsss, __dmul_rn((double)ns, (double)ns)));
    }
    a.t.nr[node] = __fmul_rn(0.5f, __fsqrt_rn(bsss));
}

// ---- leaf records (multi-GPU): everything finishTree and the bounding boxes need from a leaf, as one fixed-size record, so
// that ranks exchange 13 floats per LEAF instead of 4-5 floats per PARTICLE and no rank ever has to read another rank's
// particles to complete its node arrays. Record of leaf j (particles [j*block, ...)), K = 3*PD + SD + 1 floats:
//   [0,PD) bbox min   [PD,2PD) bbox max   [2PD,3PD) centre (t.x)   [3PD,3PD+SD) strength sums (t.s)   [3PD+SD] mean radius (t.pr)
// The sums are the same sequential double sums as k_finish_leaves (finishTree :753-806), so the values are bit-identical.
__global__ void k_leafrec_make(const FinishArgs a, uint32_t leaf0, uint32_t leaf1, float* __restrict__ rec, int K) {
    const uint32_t leaf = leaf0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (leaf >= leaf1) return;
    const uint32_t pf = leaf * a.block, pl = min(pf + a.block, a.p.n), n = pl - pf;
    const int PD = a.PD, SD = a.are_sources ? a.SD : 0;
    float* out = rec + (size_t)leaf * K;
    for (int d = 0; d < PD; ++d) {
        float lo = INFINITY, hi = -INFINITY;
        for (uint32_t i = pf + lane; i < pl; i += 32) { const float v = a.p.x[d][i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
        lo = warp_min(lo); hi = warp_max(hi);
        if (lane == 0) { out[d] = lo; out[PD + d] = hi; }
    }
    double acc = 0.0;
    if (lane <= PD) {
        for (uint32_t i = pf; i < pl; ++i) {
            float w;
            if (!a.are_sources) w = 1.0f;
            else if (a.SD == 1) w = fabsf(a.p.s[0][i]);
            else {
                w = 0.0f;
                for (int d = 0; d < a.SD; ++d) { const double sd = (double)a.p.s[d][i]; w = __double2float_rn(__dadd_rn((double)w, __dmul_rn(sd, sd))); }   // :775
                w = __fsqrt_rn(w);
            }
            if (lane < PD) acc = __dadd_rn(acc, (double)__fmul_rn(a.p.x[lane][i], w));     // :790
            else           acc = __dadd_rn(acc, (double)w);                                // :786
        }
    } else if (lane <= PD + SD) {
        const float* __restrict__ sd = a.p.s[lane - PD - 1];
        for (uint32_t i = pf; i < pl; ++i) acc = __dadd_rn(acc, (double)sd[i]);            // :796
    } else if (lane == PD + SD + 1) {
        for (uint32_t i = pf; i < pl; ++i) acc = __dadd_rn(acc, (double)a.p.r[i]);         // :800
    }
    const double wsum = __shfl_sync(0xffffffffu, acc, PD);
    if (lane < PD) {
        const float ooass = __double2float_rn(__ddiv_rn(1.0, __dadd_rn(1.e-20, wsum)));
        out[2 * PD + lane] = __double2float_rn(__dmul_rn((double)ooass, acc));
    } else if (lane > PD && lane <= PD + SD) {
        out[3 * PD + (lane - PD - 1)] = __double2float_rn(acc);
    } else if (lane == PD + SD + 1) {
        out[3 * PD + a.SD * a.are_sources] = __fdiv_rn(__double2float_rn(acc), __uint2float_rn(n));   // :801
    }
}
// records (complete after the all-gather) -> node arrays of the leaves + the box scratch of k_bbox_parents
__global__ void k_leafrec_apply(const FinishArgs a, const float* __restrict__ rec, int K, float* lohi) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= (uint32_t)a.t.numnodes || node == 0) return;
    const uint32_t n = a.t.num[node];
    if (n == 0 || n > a.block) return;
    const float* in = rec + (size_t)(a.t.ioffset[node] / a.block) * K;
    const size_t nn = a.t.numnodes;
    const int PD = a.PD, SD = a.are_sources ? a.SD : 0;
    for (int d = 0; d < PD; ++d) {
        lohi[(size_t)(2 * d) * nn + node] = in[d]; lohi[(size_t)(2 * d + 1) * nn + node] = in[PD + d];
        a.t.x[d][node] = in[2 * PD + d];
    }
    for (int d = 0; d < SD; ++d) a.t.s[d][node] = in[3 * PD + d];
    a.t.pr[node] = in[3 * PD + SD];
}

#include "tree_big.cuh"
#include "tree_sub.cuh"

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
extern "C" int onb_get_build_stats(onb_context* c, uint64_t out[5]) {
    unsigned long long h[5];
    ONB_CUDA(cudaSetDevice(c->device));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    ONB_CUDA(cudaMemcpy(h, c->d_build_stats, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 5; ++i) out[i] = h[i];
    return ONB_OK;
}
static int onb_pivot_mode = 0;   // 0 = the reference source evaluated in IEEE order; 1 = as g++ -O3 -ffast-math contracts it
extern "C" void onb_set_pivot_mode(int mode) { onb_pivot_mode = mode ? 1 : 0; }

static uint32_t big_node_threshold() {      // nodes above this are split by the grid-wide kernel of tree_big.cuh
    static uint32_t v = 0;
    if (!v) { v = 32768; if (const char* e = std::getenv("ONB_BIG_NODE")) v = (uint32_t)std::max(2048, atoi(e)); }
    return v;
}
#define BIG_NODE (big_node_threshold())     // nodes above this are split by the grid-wide kernels of tree_big.cuh

static int run_finish(onb_context* c, DParts& p, DTree& t) {      // finishTree :717-807
    FinishArgs fa; fa.p = view_of(p); fa.t = view_of(t); fa.block = c->block; fa.PD = c->PD; fa.SD = c->SD; fa.are_sources = p.are_sources ? 1 : 0;
    k_finish_leaves<<<((size_t)t.numnodes * 32 + 255) / 256, 256, 0, ONB_ST(c)>>>(fa); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    for (int lev = t.levels - 2; lev >= 0; --lev) {
        const uint32_t nn = 1u << lev;
        k_finish_parents<<<(nn + 127) / 128, 128, 0, ONB_ST(c)>>>(fa, lev); ONB_LAUNCH(c);
    }
    ONB_CUDA(cudaGetLastError());
    return ONB_OK;
}

int onb_tree_build(onb_context* c, DParts& p, DTree& t, uint32_t blo, uint32_t bhi, bool finish) {
    const uint32_t n = p.n;
    if (bhi > n) bhi = n;
    if (blo >= bhi) { c->err = "make_tree: empty build range"; return ONB_ERR_ARG; }
    p.build_lo = blo; p.build_hi = bhi;
    const int PD = c->PD, SD = p.are_sources ? c->SD : 0;
    if (n == 0) { c->err = "make_tree: no particles"; return ONB_ERR_ARG; }
    ONB_CUDA(cudaMemsetAsync(ONB_STATS(c), 0, 4 * sizeof(unsigned long long), ONB_ST(c)));

    // scratch: second copy of every plane (ping-pong), per-particle index planes, per-node split records
    const size_t capf = (size_t)p.cap * sizeof(float);
    float* alt_x[3] = {nullptr, nullptr, nullptr}; float* alt_r = nullptr; float* alt_s[3] = {nullptr, nullptr, nullptr};
    uint32_t *alt_g = nullptr, *lidx = nullptr, *scr = nullptr, *owner = nullptr, *pmid = nullptr; uint8_t* axis_of = nullptr;
    for (int d = 0; d < PD; ++d) ONB_CUDA(onb_dmalloc(c, (void**)&alt_x[d], capf));
    ONB_CUDA(onb_dmalloc(c, (void**)&alt_r, capf));
    for (int d = 0; d < SD; ++d) ONB_CUDA(onb_dmalloc(c, (void**)&alt_s[d], capf));
    ONB_CUDA(onb_dmalloc(c, (void**)&alt_g, (size_t)n * 4));
    ONB_CUDA(onb_dmalloc(c, (void**)&lidx, (size_t)n * 4)); ONB_CUDA(onb_dmalloc(c, (void**)&scr, (size_t)n * 4)); ONB_CUDA(onb_dmalloc(c, (void**)&owner, (size_t)n * 4));
    ONB_CUDA(onb_dmalloc(c, (void**)&pmid, (size_t)t.numnodes * 4)); ONB_CUDA(onb_dmalloc(c, (void**)&axis_of, (size_t)t.numnodes));
    // the tree-order index plane is persistent for targets (gidx), scratch for sources (dropped, barneshut.hpp:853)
    uint32_t* own_g = nullptr;
    if (!p.are_sources) {
        if (!p.gidx && p.gidx_spare) { p.gidx = p.gidx_spare; p.gidx_spare = nullptr; }
        if (!p.gidx) ONB_CUDA(onb_pmalloc(c, (void**)&p.gidx, (size_t)n * 4));
        own_g = p.gidx;
    }
    else ONB_CUDA(onb_dmalloc(c, (void**)&own_g, (size_t)n * 4));
    // grid-wide select state for the big nodes of the top levels
    const uint32_t max_big_nodes = n / BIG_NODE + 2, max_chunks = n / BIG_CH + max_big_nodes + 1;
    BigNode* bignodes = nullptr; uint32_t *nbig = nullptr, *chunk_owner = nullptr, *cntA = nullptr, *cntB = nullptr, *bwl = nullptr;
    if (n > BIG_NODE) {
        ONB_CUDA(onb_dmalloc(c, (void**)&bignodes, (size_t)max_big_nodes * sizeof(BigNode)));
        ONB_CUDA(onb_dmalloc(c, (void**)&nbig, 64));
        ONB_CUDA(onb_dmalloc(c, (void**)&chunk_owner, (size_t)max_chunks * 4));
        ONB_CUDA(onb_dmalloc(c, (void**)&cntA, (size_t)max_chunks * 4)); ONB_CUDA(onb_dmalloc(c, (void**)&cntB, (size_t)max_chunks * 4));
        ONB_CUDA(onb_dmalloc(c, (void**)&bwl, (size_t)max_chunks * 8));
    }

    unsigned long long* big_prof = nullptr;      // diagnostics: ONB_BIG_PROF=1 prints the phase timeline of every big level
    static const bool want_prof = std::getenv("ONB_BIG_PROF") != nullptr;
    if (want_prof) { ONB_CUDA(onb_dmalloc(c, (void**)&big_prof, (size_t)t.levels * 512 * 8)); ONB_CUDA(cudaMemsetAsync(big_prof, 0, (size_t)t.levels * 512 * 8, ONB_ST(c))); }

    const int TB = 256; const uint32_t GB = (n + TB - 1) / TB;
    k_fill_u32<<<GB, TB, 0, ONB_ST(c)>>>(own_g, n, 0, 1); ONB_LAUNCH(c);                 // gidx = iota :823
    k_fill_u32<<<GB, TB, 0, ONB_ST(c)>>>(owner, n, 1, 0); ONB_LAUNCH(c);
    ONB_CUDA(cudaMemsetAsync(t.num, 0, (size_t)t.numnodes * 4, ONB_ST(c)));
    ONB_CUDA(cudaMemsetAsync(t.ioffset, 0, (size_t)t.numnodes * 4, ONB_ST(c)));
    { const uint32_t root[1] = { n }; ONB_CUDA(cudaMemcpyAsync(t.num + 1, root, 4, cudaMemcpyHostToDevice, ONB_ST(c))); }

    float* cx[3] = { p.x[0], p.x[1], p.x[2] }; uint32_t* cg = own_g;
    float* ax[3] = { alt_x[0], alt_x[1], alt_x[2] }; uint32_t* ag = alt_g;

    static uint32_t sub_max = 0;     // particles per shared-memory subtree CTA (tree_sub.cuh): 8192 x 1024 threads or 4096 x 512
    if (!sub_max) { sub_max = 8192; if (const char* e = std::getenv("ONB_SUB")) sub_max = atoi(e) == 4096 ? 4096u : 8192u; }
    uint32_t leftmost = n;     // the leftmost node of a level is its largest
    // the nodes of this level that contain the first / last particle of the build range: [spf, epl) is what this level touches
    uint32_t spf = 0, spl = n, epf = 0, epl = n;
    for (int lev = 0; lev < t.levels; ++lev) {
        if (leftmost <= sub_max && (t.levels - lev) <= 10) {
            // every node of this level fits in shared memory: one kernel does all the levels below and writes the
            // final order of the coordinates and the index plane (tree_sub.cuh)
            SubArgs sa;
            for (int d = 0; d < 3; ++d) { sa.x[d] = cx[d]; sa.ox[d] = p.x[d]; }
            sa.g = cg; sa.og = own_g; sa.t = view_of(t); sa.stats = ONB_STATS(c);
            sa.block = c->block; sa.blo = blo; sa.bhi = bhi; sa.level = lev; sa.nsub = t.levels - lev; sa.PD = PD; sa.pivot_mode = onb_pivot_mode;
            const size_t sub_smem = (size_t)sub_max * (3 * sizeof(float) + 2 * sizeof(uint16_t));
            static bool attr_set[64] = {false};      // function attributes are per device (a process may hold contexts on several)
            if (!attr_set[c->device & 63]) {
                ONB_CUDA(cudaFuncSetAttribute(k_subtree<1024, 8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
                ONB_CUDA(cudaFuncSetAttribute(k_subtree<512, 4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
                attr_set[c->device & 63] = true;
            }
            if (sub_max == 8192) k_subtree<1024, 8192><<<1u << lev, 1024, sub_smem, ONB_ST(c)>>>(sa);
            else                 k_subtree<512, 4096><<<1u << lev, 512, sub_smem, ONB_ST(c)>>>(sa);
            ONB_LAUNCH(c);
            ONB_CUDA(cudaGetLastError());
            break;
        }
        if (leftmost > BIG_NODE) {
            BigArgs ba;
            for (int d = 0; d < 3; ++d) ba.x[d] = cx[d];
            ba.t = view_of(t); ba.nodes = bignodes; ba.nbig = nbig; ba.chunk_owner = chunk_owner; ba.cntA = cntA; ba.cntB = cntB; ba.wl = bwl;
            ba.lidx = lidx; ba.scr = scr; ba.axis_of = axis_of; ba.pmid = pmid; ba.stats = ONB_STATS(c);
            ba.block = c->block; ba.big = BIG_NODE; ba.level = lev; ba.PD = PD; ba.pivot_mode = onb_pivot_mode;
            ba.blo = blo; ba.bhi = bhi;
            ba.prof = big_prof ? big_prof + (size_t)lev * 512 : nullptr;
            const uint32_t lev_nodes = std::min<uint64_t>(1ull << lev, max_big_nodes);
            ba.max_nodes = lev_nodes; ba.max_chunks = max_chunks;
            const uint32_t chunks_ub = std::min<uint32_t>(max_chunks, n / BIG_CH + lev_nodes + 1);
            {
                static int coop_max_dev[64] = {0}, coop_want = 4, coop_want_conc = 2;
                int& coop_max = coop_max_dev[c->device & 63];      // occupancy is a per-device property
                if (!coop_max) {
                    ONB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&coop_max, k_big_level, BIG_T, 0));
                    if (const char* e = std::getenv("ONB_BIG_BLOCKS_PER_SM")) coop_want = std::max(1, atoi(e));
                    if (const char* e = std::getenv("ONB_BIG_BLOCKS_PER_SM_CONC")) coop_want_conc = std::max(1, atoi(e));
                    coop_max = std::max(1, coop_max);
                }
                // two builds in flight (onb_make_trees): both cooperative grids must be resident at once, so each takes
                // at most half of what fits on an SM
                const int per_sm = c->concurrent_builds ? std::max(1, std::min(coop_max / 2, coop_want_conc)) : std::min(coop_max, coop_want);
                const uint32_t grid = std::min<uint32_t>(chunks_ub, (uint32_t)(c->sm_count * per_sm));
                void* args[] = { (void*)&ba };
                ONB_CUDA(cudaLaunchCooperativeKernel((void*)k_big_level, dim3(grid), dim3(BIG_T), args, 0, ONB_ST(c))); ONB_LAUNCH(c);
            }
            ONB_CUDA(cudaGetLastError());
        }
        SplitArgs sa;
        for (int d = 0; d < 3; ++d) sa.x[d] = cx[d];
        sa.t = view_of(t); sa.axis_of = axis_of; sa.pmid = pmid; sa.lidx = lidx; sa.scr = scr; sa.stats = ONB_STATS(c);
        sa.block = c->block; sa.big = BIG_NODE; sa.blo = blo; sa.bhi = bhi; sa.level = lev; sa.PD = PD; sa.pivot_mode = onb_pivot_mode;
        // ~8 particles per thread: small nodes get small CTAs so that several share an SM and hide each other's barriers
        int threads = 1024;
        const uint32_t largest_small = std::min(leftmost, BIG_NODE);
        while (threads > 64 && (uint32_t)threads * 8 > largest_small) threads >>= 1;
        k_node_split<<<1u << lev, threads, 0, ONB_ST(c)>>>(sa); ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        GatherArgs ga;
        for (int d = 0; d < 3; ++d) { ga.sx[d] = cx[d]; ga.dx[d] = ax[d]; }
        ga.sg = cg; ga.dg = ag;
        ga.lidx = lidx; ga.owner = owner; ga.axis_of = axis_of; ga.pmid = pmid; ga.num = t.num; ga.ioffset = t.ioffset;
        ga.n = epl; ga.span_lo = spf; ga.blo = blo; ga.bhi = bhi; ga.block = c->block; ga.level = lev; ga.PD = PD;
        k_gather<<<(epl - spf + TB - 1) / TB, TB, 0, ONB_ST(c)>>>(ga); ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        for (int d = 0; d < 3; ++d) std::swap(cx[d], ax[d]);
        std::swap(cg, ag);
        leftmost = (uint32_t)c->block * (1u << (31 - __builtin_clz((leftmost - 1) / c->block)));
        // descend the two boundary nodes (the split position of a node depends on its size only, barneshut.hpp:663)
        if (spl - spf > (uint32_t)c->block) { const uint32_t pm = spf + (uint32_t)c->block * (1u << (31 - __builtin_clz((spl - spf - 1) / c->block))); if (blo < pm) spl = pm; else spf = pm; }
        if (epl - epf > (uint32_t)c->block) { const uint32_t pm = epf + (uint32_t)c->block * (1u << (31 - __builtin_clz((epl - epf - 1) / c->block))); if (bhi - 1 < pm) epl = pm; else epf = pm; }
    }
    // radii and strengths: one gather through the composed index (own_g[i] = position of particle i before the build),
    // out of place into the scratch copies, then home
    {
        PermArgs pa; pa.g = own_g; pa.lo = spf; pa.hi = epl; pa.nplanes = 1 + SD;
        pa.src[0] = p.r; pa.dst[0] = alt_r;
        for (int d = 0; d < 3; ++d) { pa.src[1 + d] = d < SD ? p.s[d] : nullptr; pa.dst[1 + d] = d < SD ? alt_s[d] : nullptr; }
        k_apply_perm<<<(epl - spf + TB - 1) / TB, TB, 0, ONB_ST(c)>>>(pa); ONB_LAUNCH(c);
        PermArgs pb = pa;
        for (int q = 0; q < 4; ++q) { pb.src[q] = pa.dst[q]; pb.dst[q] = const_cast<float*>(pa.src[q]); }
        k_copy_back<<<(epl - spf + TB - 1) / TB, TB, 0, ONB_ST(c)>>>(pb); ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
    }
    p.packed_valid = false;
    if (big_prof) {
        std::vector<unsigned long long> h((size_t)t.levels * 512);
        ONB_CUDA(cudaStreamSynchronize(ONB_ST(c)));
        ONB_CUDA(cudaMemcpy(h.data(), big_prof, h.size() * 8, cudaMemcpyDeviceToHost));
        for (int lev = 0; lev < t.levels; ++lev) {
            const unsigned long long* q = h.data() + (size_t)lev * 512;
            if (!q[0]) continue;
            fprintf(stderr, "big level %d: %llu stamps, total %.1f us; list+owner %.1f bbox %.1f setup %.1f | passes (count+scan, compact, swap):", lev, q[0],
                    (q[q[0]] - q[1]) * 1e-3, (q[2] - q[1]) * 1e-3, (q[3] - q[2]) * 1e-3, (q[4] - q[3]) * 1e-3);
            for (unsigned long long i = 4; i + 3 <= q[0]; i += 3) fprintf(stderr, " [%.1f %.1f %.1f]", (q[i + 1] - q[i]) * 1e-3, (q[i + 2] - q[i + 1]) * 1e-3, (q[i + 3] - q[i + 2]) * 1e-3);
            fprintf(stderr, "\n");
        }
    }

    // (multi-GPU: the node summaries come from exchanged leaf records instead, onb_tree_leaf_records / onb_tree_finish_from_records)
    if (finish) { int rc = run_finish(c, p, t); if (rc) return rc; }
    t.built = true;
    return ONB_OK;
}

// multi-GPU: the records of the leaves [leaf0, leaf1) of a tree-ordered particle range, K floats each, into rec[leaf*K ...]
int onb_leafrec_floats(const onb_context* c, bool are_sources) { return 3 * c->PD + (are_sources ? c->SD : 0) + 1; }
int onb_tree_leaf_records(onb_context* c, DParts& p, DTree& t, uint32_t leaf0, uint32_t leaf1, float* rec) {
    if (leaf1 <= leaf0) return ONB_OK;
    FinishArgs fa; fa.p = view_of(p); fa.t = view_of(t); fa.block = c->block; fa.PD = c->PD; fa.SD = c->SD; fa.are_sources = p.are_sources ? 1 : 0;
    const int K = onb_leafrec_floats(c, p.are_sources);
    k_leafrec_make<<<(unsigned)(((size_t)(leaf1 - leaf0) * 32 + 255) / 256), 256, 0, ONB_ST(c)>>>(fa, leaf0, leaf1, rec, K); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    return ONB_OK;
}
// ... and every node array from the complete set of records: leaves copied, parents bottom-up (boxes as unions - exact
// min/max, bit-identical to the top-down build -, centres / strengths / radii as finishTree :721-746)
int onb_tree_finish_from_records(onb_context* c, DParts& p, DTree& t, const float* rec) {
    if (!t.built) { c->err = "finish_tree: build the tree first"; return ONB_ERR_ARG; }
    float* lohi = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&lohi, (size_t)6 * t.numnodes * sizeof(float)));
    FinishArgs fa; fa.p = view_of(p); fa.t = view_of(t); fa.block = c->block; fa.PD = c->PD; fa.SD = c->SD; fa.are_sources = p.are_sources ? 1 : 0;
    const int K = onb_leafrec_floats(c, p.are_sources);
    k_leafrec_apply<<<(t.numnodes + 255) / 256, 256, 0, ONB_ST(c)>>>(fa, rec, K, lohi); ONB_LAUNCH(c);
    for (int lev = t.levels - 2; lev >= 0; --lev) {
        const uint32_t nn = 1u << lev;
        k_bbox_parents<<<(nn + 127) / 128, 128, 0, ONB_ST(c)>>>(fa, lohi, lev); ONB_LAUNCH(c);
        k_finish_parents<<<(nn + 127) / 128, 128, 0, ONB_ST(c)>>>(fa, lev); ONB_LAUNCH(c);
    }
    k_bbox_final<<<(t.numnodes + 127) / 128, 128, 0, ONB_ST(c)>>>(fa, lohi); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    return ONB_OK;
}

// Recompute EVERY node array bottom-up from complete, tree-ordered particle planes (after the ranks of a multi-GPU
// run have exchanged their shares of a range-restricted build). Tight boxes are exact min/max, so leaf boxes from the
// particles and parent boxes as the union of the children give bit-identical nc/ns/nr to the top-down build.
int onb_tree_finish_from_particles(onb_context* c, DParts& p, DTree& t) {
    if (!t.built) { c->err = "finish_tree: build the tree first"; return ONB_ERR_ARG; }
    float* lohi = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&lohi, (size_t)6 * t.numnodes * sizeof(float)));
    FinishArgs fa; fa.p = view_of(p); fa.t = view_of(t); fa.block = c->block; fa.PD = c->PD; fa.SD = c->SD; fa.are_sources = p.are_sources ? 1 : 0;
    k_bbox_leaves<<<((size_t)t.numnodes * 32 + 255) / 256, 256, 0, ONB_ST(c)>>>(fa, lohi); ONB_LAUNCH(c);
    for (int lev = t.levels - 2; lev >= 0; --lev) {
        const uint32_t nn = 1u << lev;
        k_bbox_parents<<<(nn + 127) / 128, 128, 0, ONB_ST(c)>>>(fa, lohi, lev); ONB_LAUNCH(c);
    }
    k_bbox_final<<<(t.numnodes + 127) / 128, 128, 0, ONB_ST(c)>>>(fa, lohi); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    p.build_lo = 0; p.build_hi = p.n;
    p.packed_valid = false;
    return run_finish(c, p, t);
}

int onb_tree_refine(onb_context* c, DParts& p, DTree& t, bool check_now) {
    if (!t.built) { c->err = "refine: tree not built"; return ONB_ERR_ARG; }
    if (c->block > 128) { c->err = "refine: block size > 128 not supported by the GPU build"; return ONB_ERR_UNSUPPORTED; }
    ONB_CUDA(cudaMemsetAsync(ONB_STATS(c) + 4, 0, sizeof(unsigned long long), ONB_ST(c)));
    RefineArgs ra; ra.p = view_of(p); ra.n = p.n; ra.block = c->block; ra.PD = c->PD; ra.SD = c->SD; ra.OD = c->OD;
    ra.are_sources = p.are_sources ? 1 : 0; ra.flag = c->d_flag; ra.tie_sorts = ONB_STATS(c) + 4;
    // leaves of the build range only (whole tree unless onb_make_tree_range restricted it; ranges are leaf aligned)
    const uint32_t leaf0 = p.build_lo / c->block, leaf1 = (std::min(p.build_hi, p.n) + c->block - 1) / c->block;
    ra.leaf0 = leaf0;
    if (leaf1 <= leaf0) return ONB_OK;                  // a rank without leaves
    k_refine<<<leaf1 - leaf0, 128, 0, ONB_ST(c)>>>(ra); ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    p.packed_valid = false;
    if (!check_now) return ONB_OK;      // the caller checks the flag after joining its streams (onb_prepare_eval)
    return onb_check_flag(c, "refine (introsort depth limit: libstdc++ heapsort fallback is not restated)");
}
