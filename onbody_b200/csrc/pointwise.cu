/*
 * pointwise.cu - the per-target treecodes: nbody_treecode2 (equivalent-particle, barneshut.hpp:137-222) and
 * nbody_treecode1 (box as one particle, barneshut.hpp:65-132, tpinter ongrav3d.cpp:174-181).
 *
 * In the reference every target point walks the source tree on its own (the MAC is per point), so no two
 * targets share a list. On the GPU one warp takes 32 consecutive (tree-ordered, hence spatially compact) targets
 * and walks the UNION of their lists depth first, carrying a lane mask per stack entry: a lane that accepted a
 * node's equivalent particles is masked off for that node's subtree, a lane that rejected it stays on. Each lane
 * therefore sees exactly the reference's per-point interaction list in the reference's (pre-order) order, while
 * the source tile is fetched once per warp into shared memory and broadcast - the same inner loop as p2p.cu.
 * The interaction list is generated on the fly and never stored (its length is 32x the boxwise one);
 * the reference's counters sltp/sbtp come out of the lane masks exactly.
 */
#include "onb_internal.h"
#include "pair.cuh"

namespace {

__device__ __forceinline__ float dist_sq_step(float dist, float v) {
    return __double2float_rn(__dadd_rn((double)dist, __dmul_rn((double)v, (double)v)));   // dist += std::pow(v,2)
}

struct PwArgs {
    const float* tx[3]; const float* tr; float* tu[ONB_MAX_OD]; double* tud[ONB_MAX_OD];
    TreeView st;
    const float4* s_pk0; const float4* s_pk1; const float* s_pk2;
    const float4* e_pk0; const float4* e_pk1; const float* e_pk2;
    unsigned long long* stats;     // [0] sltp [1] sbtp [9] pairs
    const uint32_t* s_epnum;       // legacy equivalents: per source node count (null = num_eqps everywhere)
    uint32_t t_lo, t_hi, block, ebs, num_eqps; float theta;
};

constexpr int PW_WARPS = 4;

template <bool A64> struct AccType { typedef float type; };
template <> struct AccType<true> { typedef double type; };

template <int PHYS, bool STRICT, int VARIANT, bool A64>
__global__ void __launch_bounds__(PW_WARPS * 32) k_pointwise(const __grid_constant__ PwArgs a) {
    constexpr int OD = Phys<PHYS>::OD, PD = Phys<PHYS>::PD;
    typedef typename AccType<A64>::type acc_t;      // ACCUM = double: fp64 outputs (onb_set_accum)
    __shared__ float4 sA[PW_WARPS][128];
    __shared__ float4 sB[PW_WARPS][Phys<PHYS>::NF4 > 1 ? 128 : 1];
    __shared__ float  sC[PW_WARPS][Phys<PHYS>::F1 ? 128 : 1];
    __shared__ uint32_t s_node[PW_WARPS][64], s_mask[PW_WARPS][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * PW_WARPS + wib;
    const uint32_t i = a.t_lo + gw * 32u + lane;
    const bool valid = i < a.t_hi;
    const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
    if (vmask == 0) return;
    const uint32_t ti = valid ? i : a.t_lo;
    Tgt tg; tg.x = a.tx[0][ti]; tg.y = a.tx[1][ti]; tg.z = PD > 2 ? a.tx[2][ti] : 0.f; tg.r2 = 0.f;
    if (Phys<PHYS>::TR) { const float r = a.tr[ti]; tg.r2 = __fmul_rn(r, r); }
    const float tpos[3] = { tg.x, tg.y, tg.z };
    acc_t acc[OD];
    #pragma unroll
    for (int d = 0; d < OD; ++d) acc[d] = valid ? (A64 ? (acc_t)a.tud[d][ti] : (acc_t)a.tu[d][ti]) : (acc_t)0;
    uint32_t n_leaf = 0, n_box = 0; unsigned long long pairs = 0;

    int sp = 0;
    if (lane == 0) { s_node[wib][0] = 1; s_mask[wib][0] = vmask; }
    sp = 1;
    __syncwarp();
    while (sp > 0) {
        --sp;
        const uint32_t S = s_node[wib][sp], m = s_mask[wib][sp];
        __syncwarp();
        const bool mine = (m >> lane) & 1u;
        const uint32_t sn = a.st.num[S];
        if (sn <= a.block) {                                                              // barneshut.hpp:75 / :148
            const uint32_t off = a.st.ioffset[S];
            for (uint32_t j = lane; j < sn; j += 32) {
                sA[wib][j] = a.s_pk0[off + j];
                if (Phys<PHYS>::NF4 > 1) sB[wib][j] = a.s_pk1[off + j];
                if (Phys<PHYS>::F1) sC[wib][j] = a.s_pk2[off + j];
            }
            __syncwarp();
            if (mine) {
                #pragma unroll 4
                for (uint32_t j = 0; j < sn; ++j)
                    pair_acc<PHYS, STRICT>(sA[wib][j], Phys<PHYS>::NF4 > 1 ? sB[wib][j] : make_float4(0.f, 0.f, 0.f, 0.f), Phys<PHYS>::F1 ? sC[wib][j] : 0.f, tg, acc);
                ++n_leaf; pairs += sn;
            }
            __syncwarp();
            continue;
        }
        bool accept = false;
        if (mine) {
            float dist = 0.0f;
            if (VARIANT == 2) {                                                           // :161-162
                #pragma unroll
                for (int d = 0; d < PD; ++d) dist = dist_sq_step(dist, __fsub_rn(a.st.nc[d][S], tpos[d]));
            } else {                                                                      // :85
                #pragma unroll
                for (int d = 0; d < PD; ++d) {
                    const float v = fmaxf(0.0f, __fsub_rn(fabsf(__fsub_rn(a.st.x[d][S], tpos[d])), __fmul_rn(0.5f, a.st.ns[d][S])));
                    dist = dist_sq_step(dist, v);
                }
            }
            dist = __fsqrt_rn(dist);
            accept = __ddiv_rn((double)dist, __dmul_rn(2.0, (double)a.st.nr[S])) > (double)a.theta;   // :93 / :175 (double compare)
        }
        const uint32_t am = __ballot_sync(0xffffffffu, accept);
        if (am) {
            if (VARIANT == 2) {                                                           // :177 equivalent particles of the node
                const uint32_t off = S * a.ebs, cnt = a.s_epnum ? a.s_epnum[S] : a.num_eqps;
                for (uint32_t j = lane; j < cnt; j += 32) {
                    sA[wib][j] = a.e_pk0[off + j];
                    if (Phys<PHYS>::NF4 > 1) sB[wib][j] = a.e_pk1[off + j];
                    if (Phys<PHYS>::F1) sC[wib][j] = a.e_pk2[off + j];
                }
                __syncwarp();
                if (accept) {
                    #pragma unroll 4
                    for (uint32_t j = 0; j < cnt; ++j)
                        pair_acc<PHYS, STRICT>(sA[wib][j], Phys<PHYS>::NF4 > 1 ? sB[wib][j] : make_float4(0.f, 0.f, 0.f, 0.f), Phys<PHYS>::F1 ? sC[wib][j] : 0.f, tg, acc);
                    ++n_box; pairs += cnt;
                }
                __syncwarp();
            } else if (accept) {                                                          // :95 tpinter: the node as one particle
                const float pr = a.st.pr[S]; const float r2 = __fmul_rn(pr, pr);
                float4 p0, p1 = make_float4(0.f, 0.f, 0.f, 0.f); float p2 = 0.f;
                if (PHYS == ONB_GRAV3D) { p0 = make_float4(a.st.x[0][S], a.st.x[1][S], a.st.x[2][S], a.st.s[0][S]); p2 = r2; }
                else if (PHYS == ONB_VORT3D || PHYS == ONB_VORTGRAD3D) {
                    p0 = make_float4(a.st.x[0][S], a.st.x[1][S], a.st.x[2][S], r2);
                    p1 = make_float4(a.st.s[0][S], a.st.s[1][S], a.st.s[2][S], 0.f);
                } else p0 = make_float4(a.st.x[0][S], a.st.x[1][S], r2, a.st.s[0][S]);
                pair_acc<PHYS, STRICT>(p0, p1, p2, tg, acc);
                ++n_box; pairs += 1;
            }
        }
        const uint32_t rem = m & ~am;
        if (rem) {                                                                        // :99-100 / :181-182 left child first
            if (lane == 0) { s_node[wib][sp] = 2 * S + 1; s_mask[wib][sp] = rem; s_node[wib][sp + 1] = 2 * S; s_mask[wib][sp + 1] = rem; }
            sp += 2;
            __syncwarp();
        }
    }
    if (valid) {
        #pragma unroll
        for (int d = 0; d < OD; ++d) { if (A64) a.tud[d][ti] = (double)acc[d]; else a.tu[d][ti] = (float)acc[d]; }
    }
    // counters: one atomic per warp
    unsigned long long v[3] = { n_leaf, n_box, pairs };
    #pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned long long x = v[k];
        for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x) atomicAdd(&a.stats[k == 2 ? 9 : k], x);
    }
}

template <int PHYS, bool A64>
void launch_pw_t(onb_context* c, const PwArgs& a, uint32_t blocks, int variant) {
    const bool st = c->arith == ONB_ARITH_STRICT;
    if (variant == 2) { if (st) k_pointwise<PHYS, true, 2, A64><<<blocks, PW_WARPS * 32, 0, c->stream>>>(a); else k_pointwise<PHYS, false, 2, A64><<<blocks, PW_WARPS * 32, 0, c->stream>>>(a); }
    else              { if (st) k_pointwise<PHYS, true, 1, A64><<<blocks, PW_WARPS * 32, 0, c->stream>>>(a); else k_pointwise<PHYS, false, 1, A64><<<blocks, PW_WARPS * 32, 0, c->stream>>>(a); }
}
template <int PHYS>
void launch_pw(onb_context* c, const PwArgs& a, uint32_t blocks, int variant) {
    if (c->accum64) launch_pw_t<PHYS, true>(c, a, blocks, variant); else launch_pw_t<PHYS, false>(c, a, blocks, variant);
}

}  // namespace

int onb_run_treecode2(onb_context* c, float theta, int variant) {
    DParts& srcs = c->parts[0]; DParts& eqs = c->parts[2]; DParts& t = c->parts[1];
    if (!srcs.packed_valid) { int rc = onb_pack_sources(c, srcs); if (rc) return rc; }
    if (variant == 2 && !eqs.packed_valid) { int rc = onb_pack_sources(c, eqs); if (rc) return rc; }
    unsigned long long* d_stats = nullptr;
    ONB_CUDA(onb_dmalloc(c, (void**)&d_stats, 10 * sizeof(unsigned long long)));
    ONB_CUDA(cudaMemsetAsync(d_stats, 0, 10 * sizeof(unsigned long long), c->stream));
    PwArgs a;
    for (int d = 0; d < 3; ++d) a.tx[d] = t.x[d];
    a.tr = t.r;
    for (int d = 0; d < ONB_MAX_OD; ++d) { a.tu[d] = t.u[d]; a.tud[d] = t.ud[d]; }
    if (c->accum64 && !t.ud[0]) { c->err = "ACCUM = double: set the targets after onb_set_accum"; return ONB_ERR_ARG; }
    a.st = view_of(c->trees[0]);
    a.s_pk0 = srcs.pk0; a.s_pk1 = srcs.pk1; a.s_pk2 = srcs.pk2;
    a.e_pk0 = eqs.pk0; a.e_pk1 = eqs.pk1; a.e_pk2 = eqs.pk2;
    a.stats = d_stats;
    // the shard is the same leaf-aligned index range as for the list-driven methods (treecode1/2 need no target tree for it)
    onb_shard_range(c, &a.t_lo, &a.t_hi);
    a.block = c->block; a.ebs = c->ebs; a.num_eqps = c->num_eqps; a.theta = theta;
    a.s_epnum = c->legacy ? c->d_epnum : nullptr;
    const uint32_t nt = a.t_hi - a.t_lo;
    if (nt > 0) {
        const uint32_t blocks = (nt + PW_WARPS * 32 - 1) / (PW_WARPS * 32);
        PhaseTimer tp(c, "p2p");
        switch (c->physics) {
            case ONB_GRAV3D:     launch_pw<ONB_GRAV3D>(c, a, blocks, variant); break;
            case ONB_VORT3D:     launch_pw<ONB_VORT3D>(c, a, blocks, variant); break;
            case ONB_VORTGRAD3D: launch_pw<ONB_VORTGRAD3D>(c, a, blocks, variant); break;
            case ONB_VORT2D:     launch_pw<ONB_VORT2D>(c, a, blocks, variant); break;
            default:             launch_pw<ONB_VORT2DTR>(c, a, blocks, variant); break;
        }
        ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        tp.stop();
    }
    c->phase_ms["lists"] = 0.0;
    unsigned long long h[10];
    ONB_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    ONB_CUDA(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 9; ++i) c->stats[i] = h[i];
    c->last_pairs = h[9];
    onb_dfree(c, d_stats);
    return ONB_OK;
}
