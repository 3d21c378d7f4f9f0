/*
 * bary.cu - barycentric Lagrange equivalent particles: upward pass (anterpolation of strengths onto the
 * (order+1)^PD Chebyshev points of every non-leaf node) and downward pass (interpolation of the parent's
 * equivalent-target values onto a child's points, fused with the zero-fill the dual-tree traversal does at node entry).
 *
 * Replaces reference src/BarycentricLagrange.hpp: set_sk/set_wk :28-48, calcBarycentricLagrange :255-417,
 * calcBarycentricUpward :175-248, calcBarycentricDownward :62-166 (called from ongrav3d.cpp:232,257,270,296).
 *
 * One CTA per tree node, one level per launch (children before parents going up, parents before children going
 * down). Phase A: one thread per contributing point builds its PD x (order+1) one-dimensional weights and the
 * normalising denominator in shared memory. Phase B: one thread per equivalent point walks the contributors in
 * the reference's order (child 2i first, then 2i+1; points in index order) and accumulates. The arithmetic is the
 * reference's IEEE sequence (true divisions, no FMA contraction): these passes move O(N) data once and are
 * HBM-bound, so exact arithmetic costs nothing measurable and makes equivalent strengths bit-comparable.
 */
#include "onb_internal.h"
#include <cmath>

namespace {

constexpr int AM_STRIDE = 23;     // >= PD*(order+1) for every supported (PD, order), odd -> conflict-free rows

struct Cheb { float sk[ONB_MAX_ORDER + 1]; float wk[ONB_MAX_ORDER + 1]; };

struct UpArgs {
    PartsView p, ep; TreeView t; Cheb ch;
    uint32_t block, ebs; int level, PD, SD, ncp, numEqps, are_sources;
    // which nodes this launch covers: node = nodes ? nodes[blockIdx.x] : node0 + blockIdx.x (single GPU: the whole level; multi-GPU:
    // the rank's own interval of the level, or the short list of nodes that straddle a rank boundary - plan.cu)
    uint32_t node0; const uint32_t* nodes;
    int strengths;      // 0: positions and radii only (target trees; the replicated positions pass of a multi-GPU source tree)
};

// 1-D weights of one point against the node's Chebyshev coordinates lsk[d*ncp+k]; row = PD*ncp floats.
// Returns 1/prod_d(sum_k a_dk)   (BarycentricLagrange.hpp:193-225 == :105-137)
__device__ __forceinline__ float bary_row(int PD, int ncp, const float* wk, const float* px, const float* lsk, float* row) {
    float denom = 1.0f;
    for (int d = 0; d < PD; ++d) {
        int flag = -1; float sum = 0.0f;
        for (int k = 0; k < ncp; ++k) {
            row[d * ncp + k] = 0.0f;
            const float dist = __fsub_rn(px[d], lsk[d * ncp + k]);
            if ((double)fabsf(dist) < 1.e-10) flag = k;                                   // CLOSE_THRESH :16
            else { const float am = __fdiv_rn(wk[k], dist); row[d * ncp + k] = am; sum = __fadd_rn(sum, am); }
        }
        if (flag > -1) { sum = 1.0f; for (int k = 0; k < ncp; ++k) row[d * ncp + k] = 0.0f; row[d * ncp + flag] = 1.0f; }
        denom = __fmul_rn(denom, sum);
    }
    return __fdiv_rn(1.0f, denom);
}

// NCP > 0 (3-D, order+1 == NCP <= 5): every contributor also stores the NCP x NCP products (den*a0[k0])*a1[k1] - exactly
// the intermediate of the reference's wgt = ((den*a0)*a1)*a2 - so that the accumulation loop needs one multiplication per
// weight instead of three and two shared-memory reads instead of four. Same rounding sequence: results stay bit-identical.
constexpr int W01_STRIDE = 27;
template <int PD, int SD, int NCP>
__global__ void __launch_bounds__(128) k_upward(const UpArgs a) {
    const uint32_t node = a.nodes ? a.nodes[blockIdx.x] : a.node0 + blockIdx.x;
    if (a.t.num[node] <= a.block) return;                                                 // :266 leaves have no equivalents
    const int tid = threadIdx.x, ncp = a.ncp, numEqps = a.numEqps;
    __shared__ float lsk[3 * (ONB_MAX_ORDER + 1)];
    __shared__ float s_am[128 * AM_STRIDE];
    __shared__ float s_den[128];
    __shared__ float s_str[3][128];
    __shared__ float s_w01[NCP > 0 ? 128 * W01_STRIDE : 1];
    const uint32_t e0 = node * a.ebs;                                                     // :289
    if (tid < PD * ncp) {
        const int d = tid / ncp, k = tid % ncp;
        lsk[tid] = __fadd_rn(a.t.nc[d][node], __fmul_rn(__fmul_rn(0.5f, a.ch.sk[k]), a.t.ns[d][node]));   // :305
    }
    // my equivalent point's Chebyshev indices
    int kd[3] = {0, 0, 0};
    { int q = tid;
      #pragma unroll
      for (int d = 0; d < PD; ++d) { kd[d] = q % ncp; q /= ncp; } }
    const int o0 = kd[0], o1 = ncp + kd[1], o2 = 2 * ncp + kd[2];      // column of my 1-D weight in a contributor's row
    __syncthreads();
    if ((uint32_t)tid < a.ebs) {
        const float rr = a.p.r[a.t.ioffset[node]];                                        // :353
        #pragma unroll
        for (int d = 0; d < PD; ++d)
            a.ep.x[d][e0 + tid] = tid < numEqps ? lsk[d * ncp + kd[d]] : a.t.nc[d][node]; // :330, :336
        a.ep.r[e0 + tid] = rr;
    }
    if (!a.strengths) return;                                                             // target trees: positions only (:378,:400)

    float acc[3] = {0.0f, 0.0f, 0.0f};                                                    // :343-347
    for (uint32_t child = 2 * node; child < 2 * node + 2; ++child) {                      // :360
        const uint32_t cn = a.t.num[child];
        const bool leaf = cn <= a.block;
        const PartsView& sp = leaf ? a.p : a.ep;
        const uint32_t is = leaf ? a.t.ioffset[child] : child * a.ebs;
        const int cnt = leaf ? (int)cn : numEqps;
        if (tid < cnt) {
            float px[3];
            #pragma unroll
            for (int d = 0; d < PD; ++d) px[d] = sp.x[d][is + tid];
            const float den = bary_row(PD, ncp, a.ch.wk, px, lsk, &s_am[tid * AM_STRIDE]);
            s_den[tid] = den;
            if (NCP > 0) {
                const float* row = &s_am[tid * AM_STRIDE];
                #pragma unroll
                for (int k0 = 0; k0 < NCP; ++k0) {
                    const float w0 = __fmul_rn(den, row[k0]);
                    #pragma unroll
                    for (int k1 = 0; k1 < NCP; ++k1) s_w01[tid * W01_STRIDE + k1 * NCP + k0] = __fmul_rn(w0, row[NCP + k1]);
                }
            }
            #pragma unroll
            for (int d = 0; d < SD; ++d) s_str[d][tid] = sp.s[d][is + tid];
        }
        __syncthreads();
        if (NCP > 0) {
            if (tid < numEqps) {
                const int o01 = kd[1] * NCP + kd[0];
                #pragma unroll 4
                for (int j = 0; j < cnt; ++j) {                                           // :190, :228-241
                    const float wgt = __fmul_rn(s_w01[j * W01_STRIDE + o01], s_am[j * AM_STRIDE + o2]);
                    #pragma unroll
                    for (int d = 0; d < SD; ++d) acc[d] = __fadd_rn(acc[d], __fmul_rn(wgt, s_str[d][j]));
                }
            }
        } else if (tid < numEqps) {
            #pragma unroll 4
            for (int j = 0; j < cnt; ++j) {                                               // :190, :228-241
                const float* row = &s_am[j * AM_STRIDE];
                float wgt = __fmul_rn(s_den[j], row[o0]);
                if (PD > 1) wgt = __fmul_rn(wgt, row[o1]);
                if (PD > 2) wgt = __fmul_rn(wgt, row[o2]);
                #pragma unroll
                for (int d = 0; d < SD; ++d) acc[d] = __fadd_rn(acc[d], __fmul_rn(wgt, s_str[d][j]));
            }
        }
        __syncthreads();
    }
    if ((uint32_t)tid < a.ebs) {
        #pragma unroll
        for (int d = 0; d < SD; ++d) a.ep.s[d][e0 + tid] = tid < numEqps ? acc[d] : 0.0f;
    }
}

// ---- downward: zero-fill + interpolation from the parent, one CTA per target node of this level ----
struct DownArgs {
    PartsView tl, tb; TreeView t; Cheb ch;
    uint32_t block, ebs, shard_lo, shard_hi, node0; int level, PD, OD, ncp, numEqps;
};

// NCP > 0 (3-D, OD == 3, order+1 == NCP): the point's weights live in registers as (den*a0[k0])*a1[k1], the loops over the
// parent's equivalent points are fully unrolled and the parent's values are read as one float4 broadcast per point: one
// multiplication + three accumulations per equivalent point. The multiplication order is the reference's, so STRICT
// (separate multiply and add) stays bit-identical; FAST contracts the accumulation into FMAs.
template <int PD, int OD, int NCP, bool FAST>
__global__ void __launch_bounds__(128) k_downward(const DownArgs a) {
    const uint32_t T = a.node0 + blockIdx.x;
    const uint32_t tn = a.t.num[T];
    if (tn < 1) return;                                                                   // ongrav3d.cpp:221
    const uint32_t tio = a.t.ioffset[T];
    if (!(tio < a.shard_hi && tio + tn > a.shard_lo)) return;                             // not needed by this shard
    const int tid = threadIdx.x, ncp = a.ncp, numEqps = a.numEqps;
    const bool leaf = tn <= a.block;
    const PartsView& tp = leaf ? a.tl : a.tb;
    const uint32_t p0 = leaf ? tio : T * a.ebs;
    const int cnt = leaf ? (int)tn : numEqps;
    __shared__ float lsk[3 * (ONB_MAX_ORDER + 1)];
    __shared__ float s_am[128 * AM_STRIDE];
    __shared__ float s_pu[3][128];
    __shared__ float4 s_pu4[NCP > 0 ? 128 : 1];
    float acc[3] = {0.0f, 0.0f, 0.0f};                                                    // :232 / :270 zero fill
    if (T > 1) {
        const uint32_t pe0 = (T >> 1) * a.ebs;                                            // parent's equivalent points
        if (tid < PD * ncp) {                                                             // BarycentricLagrange.hpp:78-89
            const int d = tid / ncp, k = tid % ncp;
            int stride = 1; for (int q = 0; q < d; ++q) stride *= ncp;
            lsk[tid] = a.tb.x[d][pe0 + stride * k];
        }
        if (NCP > 0) { if (tid < numEqps) s_pu4[tid] = make_float4(a.tb.u[0][pe0 + tid], a.tb.u[1][pe0 + tid], a.tb.u[2][pe0 + tid], 0.0f); }
        else if (tid < numEqps) for (int d = 0; d < OD; ++d) s_pu[d][tid] = a.tb.u[d][pe0 + tid];
        __syncthreads();
        if (tid < cnt) {
            float px[3];
            for (int d = 0; d < PD; ++d) px[d] = tp.x[d][p0 + tid];
            float* row = &s_am[tid * AM_STRIDE];
            const float denom = bary_row(PD, ncp, a.ch.wk, px, lsk, row);
            if (NCP > 0) {
                float r01[NCP > 0 ? NCP : 1][NCP > 0 ? NCP : 1], a2[NCP > 0 ? NCP : 1];
                #pragma unroll
                for (int k0 = 0; k0 < NCP; ++k0) {
                    const float w0 = __fmul_rn(denom, row[k0]);
                    #pragma unroll
                    for (int k1 = 0; k1 < NCP; ++k1) r01[k1][k0] = __fmul_rn(w0, row[NCP + k1]);
                }
                #pragma unroll
                for (int k2 = 0; k2 < NCP; ++k2) a2[k2] = row[2 * NCP + k2];
                #pragma unroll
                for (int k2 = 0; k2 < NCP; ++k2) {
                    #pragma unroll
                    for (int k1 = 0; k1 < NCP; ++k1) {
                        #pragma unroll
                        for (int k0 = 0; k0 < NCP; ++k0) {                                // :140-156, i = (k2*ncp + k1)*ncp + k0
                            const float wgt = __fmul_rn(r01[k1][k0], a2[k2]);
                            const float4 pu = s_pu4[(k2 * NCP + k1) * NCP + k0];
                            if (FAST) { acc[0] = fmaf(wgt, pu.x, acc[0]); acc[1] = fmaf(wgt, pu.y, acc[1]); acc[2] = fmaf(wgt, pu.z, acc[2]); }
                            else { acc[0] = __fadd_rn(acc[0], __fmul_rn(wgt, pu.x)); acc[1] = __fadd_rn(acc[1], __fmul_rn(wgt, pu.y)); acc[2] = __fadd_rn(acc[2], __fmul_rn(wgt, pu.z)); }
                        }
                    }
                }
            } else {
            int k0 = 0, k1 = 0, k2 = 0;
            #pragma unroll 5
            for (int i = 0; i < numEqps; ++i) {                                           // :140-156
                float wgt = __fmul_rn(denom, row[k0]);
                if (PD > 1) wgt = __fmul_rn(wgt, row[ncp + k1]);
                if (PD > 2) wgt = __fmul_rn(wgt, row[2 * ncp + k2]);
                #pragma unroll
                for (int d = 0; d < OD; ++d) acc[d] = __fadd_rn(acc[d], __fmul_rn(wgt, s_pu[d][i]));
                if (++k0 == ncp) { k0 = 0; if (++k1 == ncp) { k1 = 0; ++k2; } }
            }
            }
        }
    }
    if (tid < cnt) for (int d = 0; d < OD; ++d) tp.u[d][p0 + tid] = acc[d];
}

// ACCUM = double (onb_set_accum): tp.u[d][ip] += wgt * sp.u[d][iep] with float weights and fp64 values (BarycentricLagrange.hpp:156
// with A = double: the product and the sum are fp64). One CTA per target node, generic in PD / OD / order.
template <int PD, int OD>
__global__ void __launch_bounds__(128) k_downward_a64(const DownArgs a) {
    const uint32_t T = a.node0 + blockIdx.x;
    const uint32_t tn = a.t.num[T];
    if (tn < 1) return;
    const uint32_t tio = a.t.ioffset[T];
    if (!(tio < a.shard_hi && tio + tn > a.shard_lo)) return;
    const int tid = threadIdx.x, ncp = a.ncp, numEqps = a.numEqps;
    const bool leaf = tn <= a.block;
    const PartsView& tp = leaf ? a.tl : a.tb;
    const uint32_t p0 = leaf ? tio : T * a.ebs;
    const int cnt = leaf ? (int)tn : numEqps;
    __shared__ float lsk[3 * (ONB_MAX_ORDER + 1)];
    __shared__ float s_am[128 * AM_STRIDE];
    __shared__ double s_pu[OD][128];
    double acc[OD];
    #pragma unroll
    for (int d = 0; d < OD; ++d) acc[d] = 0.0;
    if (T > 1) {
        const uint32_t pe0 = (T >> 1) * a.ebs;
        if (tid < PD * ncp) {
            const int d = tid / ncp, k = tid % ncp;
            int stride = 1; for (int q = 0; q < d; ++q) stride *= ncp;
            lsk[tid] = a.tb.x[d][pe0 + stride * k];
        }
        if (tid < numEqps) for (int d = 0; d < OD; ++d) s_pu[d][tid] = a.tb.ud[d][pe0 + tid];
        __syncthreads();
        if (tid < cnt) {
            float px[3];
            for (int d = 0; d < PD; ++d) px[d] = tp.x[d][p0 + tid];
            float* row = &s_am[tid * AM_STRIDE];
            const float denom = bary_row(PD, ncp, a.ch.wk, px, lsk, row);
            int k0 = 0, k1 = 0, k2 = 0;
            for (int i = 0; i < numEqps; ++i) {
                float wgt = __fmul_rn(denom, row[k0]);
                if (PD > 1) wgt = __fmul_rn(wgt, row[ncp + k1]);
                if (PD > 2) wgt = __fmul_rn(wgt, row[2 * ncp + k2]);
                #pragma unroll
                for (int d = 0; d < OD; ++d) acc[d] = __dadd_rn(acc[d], __dmul_rn((double)wgt, s_pu[d][i]));
                if (++k0 == ncp) { k0 = 0; if (++k1 == ncp) { k1 = 0; ++k2; } }
            }
        }
    }
    if (tid < cnt) for (int d = 0; d < OD; ++d) tp.ud[d][p0 + tid] = acc[d];
}

// ---------------------------------------------------------------------------------------------
// calcEquivalents barneshut.hpp:946-1061 - the drivers' default when -o is omitted: every non-leaf node gets
// ceil(cnt/2) equivalents per child, each the strength-weighted merge of two consecutive points of the child (its real
// particles if it is a leaf, its own equivalents otherwise); an odd last point is passed up unchanged. One CTA per
// node per level (deepest first), one thread per equivalent slot; the reference's IEEE operation sequence
// (std::pow(float,int) and 1.0/(float) evaluate in double) so the result is bit-identical.
// ---------------------------------------------------------------------------------------------
struct EqArgs { PartsView p, ep; TreeView t; uint32_t* epnum; uint32_t block, ebs; int level, PD, SD; };

__global__ void __launch_bounds__(128) k_equivalents(const EqArgs a) {
    const uint32_t node = (1u << a.level) + blockIdx.x;
    if (a.t.num[node] <= a.block) return;
    const uint32_t half = a.block / 2;                                                    // ep.blockSize/2 :975
    const uint32_t side = threadIdx.x >= half ? 1u : 0u, j = threadIdx.x - side * half;
    const uint32_t child = 2u * node + side;
    uint32_t cnt[2], first[2]; bool leaf[2];
    #pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t ch = 2u * node + q, cn = a.t.num[ch];
        leaf[q] = !(cn > a.block);                                                        // :963
        cnt[q] = leaf[q] ? cn : a.epnum[ch];
        first[q] = leaf[q] ? a.t.ioffset[ch] : ch * a.ebs;
    }
    if (threadIdx.x == 0) a.epnum[node] = (cnt[0] + 1) / 2 + (cnt[1] + 1) / 2;            // :1009, :1056
    const uint32_t c = cnt[side], numEq = (c + 1) / 2;
    if (j >= numEq || threadIdx.x >= a.block) return;
    const PartsView& sp = leaf[side] ? a.p : a.ep;
    const uint32_t i1 = first[side] + 2u * j, i2 = i1 + 1u;
    const uint32_t iep = node * a.ebs + side * half + j;                                  // (ep.blockSize/2) * ichild
    (void)child;
    if (2u * j + 1u < c) {
        float str1, str2;
        if (a.SD == 1) {
            str1 = fmaxf(1.e-20f, fabsf(sp.s[0][i1])); str2 = fmaxf(1.e-20f, fabsf(sp.s[0][i2]));
        } else {
            str1 = 0.0f; str2 = 0.0f;
            for (int d = 0; d < a.SD; ++d) {
                const double s1 = (double)sp.s[d][i1], s2 = (double)sp.s[d][i2];
                str1 = __double2float_rn(__dadd_rn((double)str1, __dmul_rn(s1, s1)));
                str2 = __double2float_rn(__dadd_rn((double)str2, __dmul_rn(s2, s2)));
            }
            str1 = fmaxf(1.e-20f, __fsqrt_rn(str1)); str2 = fmaxf(1.e-20f, __fsqrt_rn(str2));
        }
        const float pairm = __double2float_rn(__ddiv_rn(1.0, (double)__fadd_rn(str1, str2)));
        for (int d = 0; d < a.PD; ++d)
            a.ep.x[d][iep] = __fmul_rn(__fadd_rn(__fmul_rn(sp.x[d][i1], str1), __fmul_rn(sp.x[d][i2], str2)), pairm);
        const double r1 = (double)sp.r[i1], r2 = (double)sp.r[i2];
        const double rr = __dmul_rn(__dadd_rn(__dmul_rn(__dmul_rn(r1, r1), (double)str1), __dmul_rn(__dmul_rn(r2, r2), (double)str2)), (double)pairm);
        a.ep.r[iep] = __double2float_rn(__dsqrt_rn(rr));
        for (int d = 0; d < a.SD; ++d) a.ep.s[d][iep] = __fadd_rn(sp.s[d][i1], sp.s[d][i2]);
    } else {                                                                              // :1003-1008, :1049-1054
        for (int d = 0; d < a.PD; ++d) a.ep.x[d][iep] = sp.x[d][i1];
        for (int d = 0; d < a.SD; ++d) a.ep.s[d][iep] = sp.s[d][i1];
        a.ep.r[iep] = sp.r[i1];
    }
}

Cheb make_cheb(int order) {                                                               // set_sk / set_wk :28-48
    Cheb c;
    for (int k = 0; k <= ONB_MAX_ORDER; ++k) { c.sk[k] = 0.f; c.wk[k] = 0.f; }
    for (int k = 0; k <= order; ++k) c.sk[k] = (float)(-std::cos(k * M_PI / order));
    c.wk[0] = 0.5f;
    for (int k = 1; k < order; ++k) c.wk[k] = (k % 2) ? -1.0f : 1.0f;
    c.wk[order] = 0.5f * ((order % 2) ? -1.0f : 1.0f);
    return c;
}

}  // namespace

// allocation of the equivalent particles of a tree: numnodes/2 blocks of ebs slots (ongrav3d.cpp:645,696). In the lean memory
// mode the equivalent TARGET points of a sharded run are sparse planes: memory only under the blocks of the nodes that overlap
// this rank's shard (plan.need), the full index space everywhere else.
int onb_bary_alloc(onb_context* c, DParts& p, DParts& ep, DTree& t) {
    const uint32_t need = (uint32_t)(t.numnodes / 2) * (uint32_t)c->ebs;
    const bool sparse = !p.are_sources && c->mem_mode == ONB_MEM_LEAN && c->shard_n > 1;
    const bool want64 = !p.are_sources && c->accum64;
    if (ep.n == need && !ep.unpacked_released && ep.sparse_key == (sparse ? c->plan_key(1) : 0ull) && (ep.ud[0] != nullptr) == want64) return ONB_OK;
    if (sparse && want64) { c->err = "ACCUM = double is not available together with the lean memory mode of a sharded run"; return ONB_ERR_UNSUPPORTED; }
    onb_free_parts(c, ep);
    if (!sparse) return onb_alloc_parts(c, ep, need, p.are_sources);
    int rc = onb_plan_make(c->plan[1], p.n, c->block, c->shard_n, c->shard_rank); if (rc) { c->err = "upward: cannot plan the target partition"; return rc; }
    const ShardPlan& P = c->plan[1];
    std::vector<std::pair<size_t, size_t>> ranges;
    for (int l = 0; l < P.levels; ++l)
        if (P.need_hi[l] > P.need_lo[l]) ranges.push_back({(size_t)P.need_lo[l] * c->ebs * sizeof(float), (size_t)(P.need_hi[l] - P.need_lo[l]) * c->ebs * sizeof(float)});
    ep = DParts();
    ep.n = need; ep.cap = ((need + 63u) & ~31u) + 288u; ep.PD = c->PD; ep.SD = c->SD; ep.OD = c->OD; ep.are_sources = false;
    const size_t bytes = (size_t)ep.cap * sizeof(float);
    for (int d = 0; d < c->PD; ++d) ONB_CUDA(onb_sparse_alloc(c, (void**)&ep.x[d], bytes, ranges, ONB_ST(c)));
    ONB_CUDA(onb_sparse_alloc(c, (void**)&ep.r, bytes, ranges, ONB_ST(c)));
    for (int d = 0; d < c->OD; ++d) ONB_CUDA(onb_sparse_alloc(c, (void**)&ep.u[d], bytes, ranges, ONB_ST(c)));
    ep.sparse_key = c->plan_key(1);
    return ONB_OK;
}

template <int STR>
static void launch_upward(onb_context* c, UpArgs a, uint32_t G) {
    a.strengths = STR;
    if (G == 0) return;
    if (c->PD == 3 && c->SD == 1) { if (c->ncp == 5) k_upward<3, 1, 5><<<G, 128, 0, ONB_ST(c)>>>(a); else k_upward<3, 1, 0><<<G, 128, 0, ONB_ST(c)>>>(a); }
    else if (c->PD == 3) { if (c->ncp == 5) k_upward<3, 3, 5><<<G, 128, 0, ONB_ST(c)>>>(a); else k_upward<3, 3, 0><<<G, 128, 0, ONB_ST(c)>>>(a); }
    else k_upward<2, 1, 0><<<G, 128, 0, ONB_ST(c)>>>(a);
    ONB_LAUNCH(c);
}

// mode: ONB_UP_ALL      every non-leaf node, bottom-up (single GPU)
//       ONB_UP_OWN      the nodes completely inside this rank's particle range, bottom-up: needs no other rank's data
//       ONB_UP_SHARED   the nodes that straddle a rank boundary, bottom-up, after the owned strengths have been exchanged
//       ONB_UP_POS      positions and radii of every node (sources: replicated, it only needs the node boxes)
//       ONB_UP_NEED     positions of the nodes that overlap this rank's range (targets of a sharded run)
int onb_bary_upward_mode(onb_context* c, DParts& p, DParts& ep, DTree& t, int mode) {
    if (!t.built) { c->err = "upward: tree not built"; return ONB_ERR_ARG; }
    { int rc = onb_bary_alloc(c, p, ep, t); if (rc) return rc; }
    UpArgs a; a.p = view_of(p); a.ep = view_of(ep); a.t = view_of(t); a.ch = make_cheb(c->order);
    a.block = c->block; a.ebs = c->ebs; a.PD = c->PD; a.SD = c->SD; a.ncp = c->ncp; a.numEqps = c->num_eqps;
    a.are_sources = p.are_sources ? 1 : 0; a.nodes = nullptr; a.node0 = 0; a.strengths = 0;
    const int which = p.are_sources ? 0 : 1;
    const ShardPlan* P = nullptr;
    if (mode != ONB_UP_ALL && mode != ONB_UP_POS) {
        int rc = onb_plan_make(c->plan[which], p.n, c->block, c->shard_n, c->shard_rank); if (rc) { c->err = "upward: cannot plan the partition"; return rc; }
        P = &c->plan[which];
        if (mode == ONB_UP_SHARED) { rc = onb_plan_upload_shared(c, which); if (rc) return rc; }
    }
    for (int lev = t.levels - 2; lev >= 0; --lev) {      // the last level holds only leaves
        a.level = lev;
        switch (mode) {
            case ONB_UP_ALL:    a.node0 = 1u << lev; if (p.are_sources) launch_upward<1>(c, a, 1u << lev); else launch_upward<0>(c, a, 1u << lev); break;
            case ONB_UP_POS:    a.node0 = 1u << lev; launch_upward<0>(c, a, 1u << lev); break;
            case ONB_UP_OWN:    a.node0 = P->own_lo[lev]; launch_upward<1>(c, a, P->own_hi[lev] - P->own_lo[lev]); break;
            case ONB_UP_NEED:   a.node0 = P->need_lo[lev]; launch_upward<0>(c, a, P->need_hi[lev] - P->need_lo[lev]); break;
            case ONB_UP_SHARED: a.nodes = c->d_shared[which] + (size_t)lev * c->shard_n; launch_upward<1>(c, a, (uint32_t)P->shared[lev].size()); break;
        }
    }
    ONB_CUDA(cudaGetLastError());
    ep.packed_valid = false;
    return ONB_OK;
}
int onb_bary_upward(onb_context* c, DParts& p, DParts& ep, DTree& t) { return onb_bary_upward_mode(c, p, ep, t, ONB_UP_ALL); }

int onb_bary_downward_level(onb_context* c, int level) {
    DTree& t = c->trees[1];
    DownArgs a; a.tl = view_of(c->parts[1]); a.tb = view_of(c->parts[3]); a.t = view_of(t); a.ch = make_cheb(c->order);
    a.block = c->block; a.ebs = c->ebs; a.level = level; a.PD = c->PD; a.OD = c->OD; a.ncp = c->ncp; a.numEqps = c->num_eqps;
    // shard range in particle indices (contiguous target leaves)
    onb_shard_range(c, &a.shard_lo, &a.shard_hi);
    const bool fast = c->arith != ONB_ARITH_STRICT;
    uint32_t node0, cnt; onb_level_span(c, level, &node0, &cnt);      // the whole level, or the nodes that overlap this rank's shard
    if (cnt == 0) return ONB_OK;
    a.node0 = node0;
    const dim3 G(cnt);
    if (c->accum64) {
        if (!c->parts[1].ud[0] || !c->parts[3].ud[0]) { c->err = "ACCUM = double: set the targets and run the target upward pass after onb_set_accum"; return ONB_ERR_ARG; }
        if (c->PD == 3) k_downward_a64<3, 3><<<G, 128, 0, c->stream>>>(a); else k_downward_a64<2, 2><<<G, 128, 0, c->stream>>>(a);
        ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        return ONB_OK;
    }
    if (c->PD == 3 && c->ncp == 5) { if (fast) k_downward<3, 3, 5, true><<<G, 128, 0, c->stream>>>(a); else k_downward<3, 3, 5, false><<<G, 128, 0, c->stream>>>(a); }
    else if (c->PD == 3) k_downward<3, 3, 0, false><<<G, 128, 0, c->stream>>>(a);
    else k_downward<2, 2, 0, false><<<G, 128, 0, c->stream>>>(a);
    ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    return ONB_OK;
}

int onb_legacy_equivalents(onb_context* c, DParts& p, DParts& ep, DTree& t) {
    if (!t.built) { c->err = "upward: tree not built"; return ONB_ERR_ARG; }
    if (!p.are_sources) { c->err = "legacy equivalents exist for sources only (barneshut.hpp:953)"; return ONB_ERR_UNSUPPORTED; }
    const uint32_t need = (uint32_t)(t.numnodes / 2) * (uint32_t)c->ebs;                  // ongrav3d.cpp:645
    if (ep.n != need) {
        onb_free_parts(c, ep);
        int rc = onb_alloc_parts(c, ep, need, true);
        if (rc) return rc;
    }
    if (c->epnum_cap < (uint32_t)t.numnodes) {
        if (c->d_epnum) cudaFree(c->d_epnum);
        c->d_epnum = nullptr; c->epnum_cap = 0;
        ONB_CUDA(cudaMalloc((void**)&c->d_epnum, (size_t)t.numnodes * 4));
        c->epnum_cap = (uint32_t)t.numnodes;
    }
    ONB_CUDA(cudaMemsetAsync(c->d_epnum, 0, (size_t)t.numnodes * 4, c->stream));
    // unused slots stay zero like the reference's freshly resized arrays (Parts.hpp resize)
    const size_t fb = (size_t)ep.cap * sizeof(float);
    for (int d = 0; d < c->PD; ++d) ONB_CUDA(cudaMemsetAsync(ep.x[d], 0, fb, c->stream));
    ONB_CUDA(cudaMemsetAsync(ep.r, 0, fb, c->stream));
    for (int d = 0; d < c->SD; ++d) ONB_CUDA(cudaMemsetAsync(ep.s[d], 0, fb, c->stream));
    EqArgs a; a.p = view_of(p); a.ep = view_of(ep); a.t = view_of(t); a.epnum = c->d_epnum;
    a.block = c->block; a.ebs = c->ebs; a.PD = c->PD; a.SD = c->SD;
    for (int lev = t.levels - 2; lev >= 0; --lev) {      // the last level holds only leaves
        a.level = lev;
        k_equivalents<<<1u << lev, 128, 0, c->stream>>>(a); ONB_LAUNCH(c);
    }
    ONB_CUDA(cudaGetLastError());
    ONB_CUDA(cudaMemcpyAsync(&c->root_epnum, c->d_epnum + 1, 4, cudaMemcpyDeviceToHost, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    ep.packed_valid = false;
    return ONB_OK;
}
