/*
 * p2p.cu - leaf-block particle-particle kernels (the >95 % hot loop of the reference).
 *
 * Replaces the reference's nbody_kernel / ppinter pairs:
 *   grav3d     ongrav3d.cpp:44-58, 118-170        vort3d   onvort3d.cpp:44-59 (+CoreFunc3d.hpp:27-30)
 *   vortgrad3d onvortgrad3d.cpp:45-76 (+CoreFunc3d.hpp:34-40)
 *   vort2d     interface2dvort.cpp:39-50 (+CoreFunc2d.hpp:24-28)   vort2dtr onvort2d.cpp:44-55 (+CoreFunc2d.hpp:31-35)
 *
 * Design (sm_100a): one CTA of 128 threads owns one target block (<=128 targets, one per thread, accumulators
 * in registers) and streams its whole interaction list. Every list entry is one source tile (<=128 sources) that
 * lives contiguously in PACKED float4 planes; thread 0 moves the next tile global->shared with TMA bulk copies
 * (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) into a 2-stage ring while all four warps run the pair
 * loop on the current one. Inside the loop every lane reads the same source (shared-memory broadcast, 1-2 LDS.128),
 * so the loop is 12 FP32-pipe instructions + 1 MUFU per pair for gravity. No atomics: a target's sum is produced
 * by exactly one thread in list order, so results do not depend on the launch geometry.
 *
 * FP32 SIMT is the roofline here (a 128x128 tile moves 2.5 KB for 16384 pairs); tensor cores do not apply.
 *
 * ARITH_STRICT instantiations evaluate the reference's exact IEEE operation sequence (no FMA contraction, sqrt
 * and divide correctly rounded), which makes whole-pipeline results bit-comparable with the strict oracle.
 */
#include "onb_internal.h"
#include <cstdlib>
#include <algorithm>
#include "ptx.cuh"
#include "pair.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// packing: SoA planes -> float4 tiles (done once per source set, after the tree has ordered it)
// ---------------------------------------------------------------------------------------------
template <int PHYS>
__global__ void k_pack(PartsView p, float4* __restrict__ pk0, float4* __restrict__ pk1, float* __restrict__ pk2, uint32_t n, uint32_t cap) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    if (i >= n) {   // padding: far away, zero strength, unit radius
        pk0[i] = make_float4(0.f, 0.f, 0.f, PHYS == ONB_GRAV3D ? 0.f : 1.f);
        if (PHYS == ONB_VORT2D || PHYS == ONB_VORT2DTR) pk0[i] = make_float4(0.f, 0.f, 1.f, 0.f);
        if (Phys<PHYS>::NF4 > 1) pk1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (Phys<PHYS>::F1) pk2[i] = 1.f;
        return;
    }
    const float r = p.r[i];
    const float r2 = __fmul_rn(r, r);
    if (PHYS == ONB_GRAV3D) {
        pk0[i] = make_float4(p.x[0][i], p.x[1][i], p.x[2][i], p.s[0][i]);
        pk2[i] = r2;
    } else if (PHYS == ONB_VORT3D || PHYS == ONB_VORTGRAD3D) {
        pk0[i] = make_float4(p.x[0][i], p.x[1][i], p.x[2][i], r2);
        pk1[i] = make_float4(p.s[0][i], p.s[1][i], p.s[2][i], 0.f);
    } else {
        pk0[i] = make_float4(p.x[0][i], p.x[1][i], r2, p.s[0][i]);
    }
}

// ---------------------------------------------------------------------------------------------
// tile ring
// ---------------------------------------------------------------------------------------------
struct TileRef { const float4* p0; const float4* p1; const float* p2; uint32_t off, cnt; };

template <int PHYS>
struct TileSmem {
    alignas(128) float4 a[2][128];
    alignas(128) float4 b[2][Phys<PHYS>::NF4 > 1 ? 128 : 1];
    alignas(128) float  c[2][Phys<PHYS>::F1 ? 136 : 4];
    alignas(8) unsigned long long bar[2];
};

// thread 0 only: arm the barrier and launch the bulk copies of one tile
template <int PHYS>
__device__ __forceinline__ void tile_issue(TileSmem<PHYS>& sm, int buf, const TileRef& t) {
    const uint32_t b4 = t.cnt * 16u;
    uint32_t total = b4 * Phys<PHYS>::NF4;
    uint32_t c_start = 0, c_bytes = 0;
    if (Phys<PHYS>::F1) {
        c_start = t.off & ~3u;                                   // 16-byte aligned start of the float plane
        c_bytes = (((t.off - c_start) + t.cnt) * 4u + 15u) & ~15u;
        total += c_bytes;
    }
    ptx::mbar_arrive_expect_tx(&sm.bar[buf], total);
    ptx::bulk_g2s(&sm.a[buf][0], t.p0 + t.off, b4, &sm.bar[buf]);
    if (Phys<PHYS>::NF4 > 1) ptx::bulk_g2s(&sm.b[buf][0], t.p1 + t.off, b4, &sm.bar[buf]);
    if (Phys<PHYS>::F1) ptx::bulk_g2s(&sm.c[buf][0], t.p2 + c_start, c_bytes, &sm.bar[buf]);
}

template <int PHYS, bool STRICT>
__device__ __forceinline__ void tile_compute(const TileSmem<PHYS>& sm, int buf, const TileRef& t, const Tgt& tg, float* __restrict__ acc) {
    const float4* __restrict__ A = sm.a[buf];
    const float4* __restrict__ B = sm.b[buf];
    const float*  __restrict__ C = sm.c[buf] + (Phys<PHYS>::F1 ? (t.off & 3u) : 0u);
    const int cnt = (int)t.cnt;
    #pragma unroll 8
    for (int j = 0; j < cnt; ++j) {
        const float4 p0 = A[j];
        const float4 p1 = Phys<PHYS>::NF4 > 1 ? B[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float p2 = Phys<PHYS>::F1 ? C[j] : 0.f;
        pair<PHYS, STRICT>(p0, p1, p2, tg, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// list-driven kernel: one CTA per target work item
// ---------------------------------------------------------------------------------------------
struct P2PArgs {
    const float* tx[3]; const float* tr; float* tu[ONB_MAX_OD];     // leaf targets (real particles)
    const float* bx[3]; const float* br; float* bu[ONB_MAX_OD];     // box targets (equivalent points)
    double* tud[ONB_MAX_OD]; double* bud[ONB_MAX_OD];               // ACCUM = double: the fp64 output planes of the two target sets
    const uint32_t* t_ioffset; const uint32_t* t_num;
    const float4* s_pk0; const float4* s_pk1; const float* s_pk2;   // real sources
    const float4* e_pk0; const float4* e_pk1; const float* e_pk2;   // equivalent sources
    const uint32_t* s_ioffset; const uint32_t* s_num;
    const uint32_t* item_node; const uint32_t* start; const uint32_t* entries;
    const uint32_t* s_epnum;        // legacy equivalents: per source node count (null = num_eqps everywhere)
    float* partial;                 // nsplit > 1: [item][segment][OD][128] partial sums, reduced in segment order by k_p2p_reduce
    const uint32_t* ebase;          // device-resident offset of this work list inside `entries` (dual tree: the level's pool base), or null
    const uint32_t* order;          // launch order of the work items (longest list first), or null = item index order
    uint32_t* work; uint32_t nwork; // persistent mode: atomic item counter (zeroed before the launch) and the number of items x nsplit
    uint32_t* work_counter;         // host side: a zeroed counter the launcher may use for the persistent mode
    uint32_t block, ebs, num_eqps, node_base, nentries, nsplit;
};
// [e0, e1) of work item w, clamped to the allocated entries (a dual-tree pass whose pool overflowed is discarded, not faulted)
__device__ __forceinline__ void item_range(const P2PArgs& a, uint32_t w, uint32_t& e0, uint32_t& e1) {
    const unsigned long long eb = a.ebase ? (unsigned long long)*a.ebase : 0ull;
    e0 = (uint32_t)min(eb + a.start[w], (unsigned long long)a.nentries);
    e1 = (uint32_t)min(eb + a.start[w + 1], (unsigned long long)a.nentries);
}

__device__ __forceinline__ TileRef decode_entry(const P2PArgs& a, uint32_t entry) {
    TileRef t;
    const uint32_t S = entry & 0x7fffffffu;
    if (entry >> 31) { t.p0 = a.e_pk0; t.p1 = a.e_pk1; t.p2 = a.e_pk2; t.off = S * a.ebs; t.cnt = a.s_epnum ? a.s_epnum[S] : a.num_eqps; }
    else             { t.p0 = a.s_pk0; t.p1 = a.s_pk1; t.p2 = a.s_pk2; t.off = a.s_ioffset[S]; t.cnt = a.s_num[S]; }
    return t;
}

// TPT targets per thread (register blocking): the CTA has 128/TPT threads, thread t owns targets t, t+NT, t+2NT, ...
// Every shared-memory broadcast of a source then feeds TPT pairs, which is what lifts the kernel off the shared-memory
// wavefront limit (3 wavefronts per source and warp: ncu showed the LSU data pipe at 66 % with TPT=1) and raises the
// FP32 share of the issue slots. Each target still accumulates its sources in list order, so results do not change.
template <int PHYS, bool STRICT, int TPT, bool PK>
__global__ void __launch_bounds__(128 / TPT) k_p2p_lists(const __grid_constant__ P2PArgs a) {
    constexpr int OD = Phys<PHYS>::OD;
    constexpr int NT = 128 / TPT;
    constexpr int G = PK ? TPT / 2 : 1;          // packed target pairs per thread
    static_assert(!PK || (TPT % 2 == 0 && !STRICT), "packed arithmetic needs an even TPT and the fast mode");
    __shared__ TileSmem<PHYS> sm;
    __shared__ uint32_t s_next;
    const int tid = threadIdx.x;
    if (tid == 0) { ptx::mbar_init(&sm.bar[0], 1); ptx::mbar_init(&sm.bar[1], 1); ptx::fence_mbar_init(); }
    __syncthreads();
    uint32_t tiles_done = 0;      // tiles this CTA has consumed so far: ring slot and mbarrier phase continue across work items
    // Two ways to hand out the work items (one item = one target block with its list):
    //   a.work == null : CTA b takes item b (the hardware scheduler deals CTAs out as slots free up - fine for launches of many waves)
    //   a.work != null : PERSISTENT CTAs pull items from an atomic counter. A launch that fits the machine in one wave is otherwise
    //                    dealt out statically and ends when the most loaded SM sub-partition is done; pulling items balances it.
    for (uint32_t round = 0;; ++round) {
        uint32_t bid;
        if (a.work) {
            if (NT > 32) {
                if (tid == 0) s_next = atomicAdd(a.work, 1u);
                __syncthreads();
                bid = s_next;
                __syncthreads();
            } else {
                uint32_t v = 0;
                if (tid == 0) v = atomicAdd(a.work, 1u);
                bid = __shfl_sync(0xffffffffu, v, 0);
            }
            if (bid >= a.nwork) break;
        } else {
            if (round) break;
            bid = blockIdx.x;
        }
        // nsplit > 1 (optional, fast arithmetic only): the item's list is cut into nsplit contiguous segments, one CTA each,
        // partial sums go to a.partial and are added in segment order afterwards
        const uint32_t wl = bid / a.nsplit, seg = bid - wl * a.nsplit;
        const uint32_t w = a.order ? a.order[wl] : wl;
        uint32_t e0, e1; item_range(a, w, e0, e1);   // never read past the allocated list
        if (e0 >= e1) continue;
        if (a.nsplit > 1) {
            const uint32_t len = e1 - e0;
            const uint32_t s0 = e0 + (uint32_t)((unsigned long long)len * seg / a.nsplit), s1 = e0 + (uint32_t)((unsigned long long)len * (seg + 1u) / a.nsplit);
            e0 = s0; e1 = s1;
            if (e0 >= e1) continue;
        }
        const uint32_t T = a.item_node ? a.item_node[w] : a.node_base + w;
        const uint32_t tn = a.t_num[T];
        const bool leaf = tn <= a.block;
        const uint32_t toff = leaf ? a.t_ioffset[T] : T * a.ebs;
        const uint32_t tcnt = leaf ? tn : a.num_eqps;
        Tgt tg[TPT];
        float acc[TPT][OD];
        #pragma unroll
        for (int q = 0; q < TPT; ++q) {
            const uint32_t slot = (uint32_t)tid + (uint32_t)q * NT;
            const bool valid = slot < tcnt;
            const uint32_t ti = toff + (valid ? slot : 0);
            tg[q].x = leaf ? a.tx[0][ti] : a.bx[0][ti];
            tg[q].y = leaf ? a.tx[1][ti] : a.bx[1][ti];
            tg[q].z = Phys<PHYS>::PD > 2 ? (leaf ? a.tx[2][ti] : a.bx[2][ti]) : 0.f;
            tg[q].r2 = 0.f;
            if (Phys<PHYS>::TR) { const float r = leaf ? a.tr[ti] : a.br[ti]; tg[q].r2 = __fmul_rn(r, r); }
            #pragma unroll
            for (int d = 0; d < OD; ++d) acc[q][d] = (valid && a.nsplit == 1) ? (leaf ? a.tu[d][ti] : a.bu[d][ti]) : 0.f;
        }
        Tgt2 tg2[G];
        f2 acc2[G][OD];
        if (PK) {
            #pragma unroll
            for (int g = 0; g < G; ++g) {
                const Tgt& t0 = tg[2 * g]; const Tgt& t1 = tg[PK ? 2 * g + 1 : 0];
                tg2[g].px = mk2(t0.x, t1.x); tg2[g].py = mk2(t0.y, t1.y); tg2[g].pz = mk2(t0.z, t1.z);
                tg2[g].nx = mk2(-t0.x, -t1.x); tg2[g].ny = mk2(-t0.y, -t1.y); tg2[g].nz = mk2(-t0.z, -t1.z);
                tg2[g].r2 = mk2(t0.r2, t1.r2);
                #pragma unroll
                for (int d = 0; d < OD; ++d) acc2[g][d] = mk2(acc[2 * g][d], acc[PK ? 2 * g + 1 : 0][d]);
            }
        }

        TileRef cur = decode_entry(a, a.entries[e0]);
        if (tid == 0) tile_issue<PHYS>(sm, (int)(tiles_done & 1u), cur);
        for (uint32_t e = e0; e < e1; ++e, ++tiles_done) {
            const int buf = (int)(tiles_done & 1u);
            const uint32_t par = (tiles_done >> 1) & 1u;
            TileRef nxt = cur;
            if (e + 1 < e1) {
                nxt = decode_entry(a, a.entries[e + 1]);
                if (tid == 0) tile_issue<PHYS>(sm, buf ^ 1, nxt);
            }
            ptx::mbar_wait(&sm.bar[buf], par);
            {
                const float4* __restrict__ A = sm.a[buf];
                const float4* __restrict__ B = sm.b[buf];
                const float*  __restrict__ C = sm.c[buf] + (Phys<PHYS>::F1 ? (cur.off & 3u) : 0u);
                const int cnt = (int)cur.cnt;
                #pragma unroll 4
                for (int j = 0; j < cnt; ++j) {
                    const float4 p0 = A[j];
                    const float4 p1 = Phys<PHYS>::NF4 > 1 ? B[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float p2 = Phys<PHYS>::F1 ? C[j] : 0.f;
                    if (PK) {
                        #pragma unroll
                        for (int g = 0; g < G; ++g) pair2<PHYS>(p0, p1, p2, tg2[g], acc2[g]);
                    } else {
                        #pragma unroll
                        for (int q = 0; q < TPT; ++q) pair<PHYS, STRICT>(p0, p1, p2, tg[q], acc[q]);
                    }
                }
            }
            if (NT > 32) __syncthreads(); else __syncwarp();   // everyone is done with `buf` before it is refilled two iterations later
            cur = nxt;
        }
        if (PK) {
            #pragma unroll
            for (int g = 0; g < G; ++g) {
                #pragma unroll
                for (int d = 0; d < OD; ++d) { acc[2 * g][d] = lo2(acc2[g][d]); acc[PK ? 2 * g + 1 : 0][d] = hi2(acc2[g][d]); }
            }
        }
        #pragma unroll
        for (int q = 0; q < TPT; ++q) {
            const uint32_t slot = (uint32_t)tid + (uint32_t)q * NT;
            if (slot < tcnt) {
                const uint32_t ti = toff + slot;
                if (a.nsplit == 1) {
                    #pragma unroll
                    for (int d = 0; d < OD; ++d) { if (leaf) a.tu[d][ti] = acc[q][d]; else a.bu[d][ti] = acc[q][d]; }
                } else {
                    float* __restrict__ part = a.partial + (((size_t)w * a.nsplit + seg) * OD) * 128u;
                    #pragma unroll
                    for (int d = 0; d < OD; ++d) part[d * 128 + slot] = acc[q][d];
                }
            }
        }
    }
}

// ACCUM = double (the reference's STORE = float, ACCUM = double build, ongrav3d.cpp:7-8): the pair arithmetic is the fp32 one of
// pair<> - each contribution is a rounded float, exactly the "(A)(r3 * dx)" of the reference kernels - and every "+=" is an fp64
// addition into fp64 outputs. One target per thread, list order; the fp64 adds and the float->double conversions run at a
// fraction of the FP32 rate on this GPU, which is the price the reference's README describes for the last digits.
template <int PHYS, bool STRICT>
__global__ void __launch_bounds__(128) k_p2p_lists_a64(const __grid_constant__ P2PArgs a) {
    constexpr int OD = Phys<PHYS>::OD;
    __shared__ TileSmem<PHYS> sm;
    const int tid = threadIdx.x;
    const uint32_t w = a.order ? a.order[blockIdx.x] : blockIdx.x;
    uint32_t e0, e1; item_range(a, w, e0, e1);
    if (e0 >= e1) return;
    const uint32_t T = a.item_node ? a.item_node[w] : a.node_base + w;
    const uint32_t tn = a.t_num[T];
    const bool leaf = tn <= a.block;
    const uint32_t toff = leaf ? a.t_ioffset[T] : T * a.ebs;
    const uint32_t tcnt = leaf ? tn : a.num_eqps;
    const bool valid = (uint32_t)tid < tcnt;
    const uint32_t ti = toff + (valid ? (uint32_t)tid : 0u);
    Tgt tg;
    tg.x = leaf ? a.tx[0][ti] : a.bx[0][ti];
    tg.y = leaf ? a.tx[1][ti] : a.bx[1][ti];
    tg.z = Phys<PHYS>::PD > 2 ? (leaf ? a.tx[2][ti] : a.bx[2][ti]) : 0.f;
    tg.r2 = 0.f;
    if (Phys<PHYS>::TR) { const float r = leaf ? a.tr[ti] : a.br[ti]; tg.r2 = __fmul_rn(r, r); }
    double acc[OD];
    #pragma unroll
    for (int d = 0; d < OD; ++d) acc[d] = valid ? (leaf ? a.tud[d][ti] : a.bud[d][ti]) : 0.0;
    if (tid == 0) { ptx::mbar_init(&sm.bar[0], 1); ptx::mbar_init(&sm.bar[1], 1); ptx::fence_mbar_init(); }
    __syncthreads();
    TileRef cur = decode_entry(a, a.entries[e0]);
    if (tid == 0) tile_issue<PHYS>(sm, 0, cur);
    for (uint32_t e = e0; e < e1; ++e) {
        const int buf = (e - e0) & 1;
        const uint32_t par = ((e - e0) >> 1) & 1;
        TileRef nxt = cur;
        if (e + 1 < e1) { nxt = decode_entry(a, a.entries[e + 1]); if (tid == 0) tile_issue<PHYS>(sm, buf ^ 1, nxt); }
        ptx::mbar_wait(&sm.bar[buf], par);
        const float4* __restrict__ A = sm.a[buf];
        const float4* __restrict__ B = sm.b[buf];
        const float*  __restrict__ C = sm.c[buf] + (Phys<PHYS>::F1 ? (cur.off & 3u) : 0u);
        const int cnt = (int)cur.cnt;
        #pragma unroll 4
        for (int j = 0; j < cnt; ++j)
            pair_a64<PHYS, STRICT>(A[j], Phys<PHYS>::NF4 > 1 ? B[j] : make_float4(0.f, 0.f, 0.f, 0.f), Phys<PHYS>::F1 ? C[j] : 0.f, tg, acc);
        __syncthreads();
        cur = nxt;
    }
    if (valid) {
        #pragma unroll
        for (int d = 0; d < OD; ++d) { if (leaf) a.tud[d][ti] = acc[d]; else a.bud[d][ti] = acc[d]; }
    }
}

// ACCUM = double direct sum: every sample target against all sources in source order, one thread per target
struct DirectArgs64 { const float* tx[3]; const float* tr; double* tud[ONB_MAX_OD]; const float4* s_pk0; const float4* s_pk1; const float* s_pk2; uint32_t nsrc, tskip, nt_eff, k0; };
template <int PHYS, bool STRICT>
__global__ void __launch_bounds__(128) k_p2p_direct_a64(const __grid_constant__ DirectArgs64 a) {
    constexpr int OD = Phys<PHYS>::OD;
    __shared__ TileSmem<PHYS> sm;
    const int tid = threadIdx.x;
    const uint32_t k = blockIdx.x * 128u + tid;
    const bool valid = k < a.nt_eff;
    const uint32_t ti = (a.k0 + (valid ? k : 0u)) * a.tskip;
    Tgt tg;
    tg.x = a.tx[0][ti]; tg.y = a.tx[1][ti]; tg.z = Phys<PHYS>::PD > 2 ? a.tx[2][ti] : 0.f; tg.r2 = 0.f;
    if (Phys<PHYS>::TR) { const float r = a.tr[ti]; tg.r2 = __fmul_rn(r, r); }
    double acc[OD];
    #pragma unroll
    for (int d = 0; d < OD; ++d) acc[d] = valid ? a.tud[d][ti] : 0.0;
    const uint32_t ntiles = (a.nsrc + 127u) / 128u;
    if (tid == 0) { ptx::mbar_init(&sm.bar[0], 1); ptx::mbar_init(&sm.bar[1], 1); ptx::fence_mbar_init(); }
    __syncthreads();
    auto tile_of = [&](uint32_t t) { TileRef r; r.p0 = a.s_pk0; r.p1 = a.s_pk1; r.p2 = a.s_pk2; r.off = t * 128u; r.cnt = min(128u, a.nsrc - t * 128u); return r; };
    TileRef cur = tile_of(0);
    if (tid == 0) tile_issue<PHYS>(sm, 0, cur);
    for (uint32_t t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const uint32_t par = (t >> 1) & 1;
        TileRef nxt = cur;
        if (t + 1 < ntiles) { nxt = tile_of(t + 1); if (tid == 0) tile_issue<PHYS>(sm, buf ^ 1, nxt); }
        ptx::mbar_wait(&sm.bar[buf], par);
        const float4* __restrict__ A = sm.a[buf];
        const float4* __restrict__ B = sm.b[buf];
        const float*  __restrict__ C = sm.c[buf];
        const int cnt = (int)cur.cnt;
        #pragma unroll 4
        for (int j = 0; j < cnt; ++j)
            pair_a64<PHYS, STRICT>(A[j], Phys<PHYS>::NF4 > 1 ? B[j] : make_float4(0.f, 0.f, 0.f, 0.f), Phys<PHYS>::F1 ? C[j] : 0.f, tg, acc);
        __syncthreads();
        cur = nxt;
    }
    if (valid) {
        #pragma unroll
        for (int d = 0; d < OD; ++d) a.tud[d][ti] = acc[d];
    }
}

__global__ void k_round_outputs(const double* __restrict__ ud, float* __restrict__ u, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) u[i] = __double2float_rn(ud[i]);
}

// nsplit > 1: acc(target) = ((stored + partial_0) + partial_1) + ... in segment order (deterministic for a given nsplit)
template <int OD>
__global__ void __launch_bounds__(128) k_p2p_reduce(const __grid_constant__ P2PArgs a) {
    const uint32_t w = blockIdx.x, slot = threadIdx.x;
    uint32_t e0, e1; item_range(a, w, e0, e1);
    if (e0 >= e1) return;
    const uint32_t T = a.item_node ? a.item_node[w] : a.node_base + w;
    const uint32_t tn = a.t_num[T];
    const bool leaf = tn <= a.block;
    const uint32_t toff = leaf ? a.t_ioffset[T] : T * a.ebs, tcnt = leaf ? tn : a.num_eqps;
    if (slot >= tcnt) return;
    const uint32_t ti = toff + slot;
    #pragma unroll
    for (int d = 0; d < OD; ++d) {
        float v = leaf ? a.tu[d][ti] : a.bu[d][ti];
        for (uint32_t k = 0; k < a.nsplit; ++k) v += a.partial[(((size_t)w * a.nsplit + k) * OD + d) * 128u + slot];
        if (leaf) a.tu[d][ti] = v; else a.bu[d][ti] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// direct sum (nbody_naive, barneshut.hpp:46-53): every tskip-th target against ALL sources
// ---------------------------------------------------------------------------------------------
struct DirectArgs {
    const float* tx[3]; const float* tr; float* tu[ONB_MAX_OD];
    const float4* s_pk0; const float4* s_pk1; const float* s_pk2;
    float* partial;                // [nsplit][OD][nt_eff] when nsplit > 1
    uint32_t nsrc, ntarg, tskip, nt_eff, nsplit, tiles_per_split;
    uint32_t k0;                   // first sample handled here (a sharded context evaluates the samples inside its target range)
};

template <int PHYS, bool STRICT>
__global__ void __launch_bounds__(128) k_p2p_direct(const __grid_constant__ DirectArgs a) {
    constexpr int OD = Phys<PHYS>::OD;
    __shared__ TileSmem<PHYS> sm;
    const int tid = threadIdx.x;
    const uint32_t k = blockIdx.x * 128u + tid;          // sample index relative to k0
    const bool valid = k < a.nt_eff;
    const uint32_t ti = (a.k0 + (valid ? k : 0u)) * a.tskip;
    Tgt tg;
    tg.x = a.tx[0][ti]; tg.y = a.tx[1][ti]; tg.z = Phys<PHYS>::PD > 2 ? a.tx[2][ti] : 0.f; tg.r2 = 0.f;
    if (Phys<PHYS>::TR) { const float r = a.tr[ti]; tg.r2 = __fmul_rn(r, r); }
    float acc[OD];
    #pragma unroll
    for (int d = 0; d < OD; ++d) acc[d] = (valid && a.nsplit == 1) ? a.tu[d][ti] : 0.f;

    const uint32_t ntiles = (a.nsrc + 127u) / 128u;
    const uint32_t t0 = blockIdx.y * a.tiles_per_split;
    const uint32_t t1 = min(ntiles, t0 + a.tiles_per_split);
    if (tid == 0) { ptx::mbar_init(&sm.bar[0], 1); ptx::mbar_init(&sm.bar[1], 1); ptx::fence_mbar_init(); }
    __syncthreads();
    auto tile_of = [&](uint32_t t) { TileRef r; r.p0 = a.s_pk0; r.p1 = a.s_pk1; r.p2 = a.s_pk2; r.off = t * 128u; r.cnt = min(128u, a.nsrc - t * 128u); return r; };
    if (t0 < t1) {
        TileRef cur = tile_of(t0);
        if (tid == 0) tile_issue<PHYS>(sm, 0, cur);
        for (uint32_t t = t0; t < t1; ++t) {
            const int buf = (t - t0) & 1;
            const uint32_t par = ((t - t0) >> 1) & 1;
            TileRef nxt = cur;
            if (t + 1 < t1) { nxt = tile_of(t + 1); if (tid == 0) tile_issue<PHYS>(sm, buf ^ 1, nxt); }
            ptx::mbar_wait(&sm.bar[buf], par);
            tile_compute<PHYS, STRICT>(sm, buf, cur, tg, acc);
            __syncthreads();
            cur = nxt;
        }
    }
    if (valid) {
        if (a.nsplit == 1) {
            #pragma unroll
            for (int d = 0; d < OD; ++d) a.tu[d][ti] = acc[d];
        } else {
            #pragma unroll
            for (int d = 0; d < OD; ++d) a.partial[((size_t)blockIdx.y * OD + d) * a.nt_eff + k] = acc[d];
        }
    }
}

__global__ void k_reduce_partials(DirectArgs a, int OD) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.nt_eff) return;
    const uint32_t ti = (a.k0 + k) * a.tskip;
    for (int d = 0; d < OD; ++d) {
        float v = a.tu[d][ti];
        for (uint32_t y = 0; y < a.nsplit; ++y) v += a.partial[((size_t)y * OD + d) * a.nt_eff + k];
        a.tu[d][ti] = v;
    }
}

// FP32 FMA issue-rate probe: 8 independent chains per thread
__global__ void k_fma_peak(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f, b = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
            a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

int g_p2p_tpt = 0;     // 0 = per-physics default; 1, 2 or 4 forces the register blocking; +16 = scalar instead of packed f32x2

template <int PHYS, int TPT>
void launch_lists_t(onb_context* c, P2PArgs a, uint32_t nitems, bool packed) {
    uint32_t grid = nitems * a.nsplit;
    // Optional persistent mode (ONB_PERSIST=1) for launches of few waves (the upper and middle dual-tree levels; every level of a
    // multi-GPU shard): a fixed number of CTAs - ONB_PERSIST_WARPS warps per SM (default 16) - pull the items from an atomic
    // counter. OFF by default: measured at N = 1e7 it changes nothing (pair kernels 79.9 vs 80.0 ms on one GPU, 13.5 vs 13.5 ms
    // on a 1/8 shard) - ncu of such a launch (12 warps per SM resident): 53 % of the cycles have no eligible warp (short
    // scoreboard + fixed-latency waits), i.e. these launches are bound by per-warp latency, not by how items are dealt out.
    static const int persist_on = std::getenv("ONB_PERSIST") ? atoi(std::getenv("ONB_PERSIST")) : 0;
    static const int persist_warps = std::getenv("ONB_PERSIST_WARPS") ? std::max(4, atoi(std::getenv("ONB_PERSIST_WARPS"))) : 16;
    static const int persist_waves = std::getenv("ONB_PERSIST_WAVES") ? std::max(1, atoi(std::getenv("ONB_PERSIST_WAVES"))) : 8;
    constexpr int warps_per_cta = (128 / TPT + 31) / 32;
    const uint32_t slots = (uint32_t)c->sm_count * (uint32_t)std::min(32, 64 / warps_per_cta);      // resident CTAs of this shape
    const uint32_t pgrid = (uint32_t)c->sm_count * (uint32_t)std::max(1, persist_warps / warps_per_cta);
    a.work = nullptr; a.nwork = grid;
    if (persist_on && a.work_counter && grid > pgrid && (unsigned long long)grid <= (unsigned long long)persist_waves * slots) { a.work = a.work_counter; grid = pgrid; }
    if (c->arith == ONB_ARITH_STRICT) k_p2p_lists<PHYS, true, TPT, false><<<grid, 128 / TPT, 0, c->stream>>>(a);
    else if (packed && TPT > 1)       k_p2p_lists<PHYS, false, TPT, (TPT > 1)><<<grid, 128 / TPT, 0, c->stream>>>(a);
    else                              k_p2p_lists<PHYS, false, TPT, false><<<grid, 128 / TPT, 0, c->stream>>>(a);
    if (a.nsplit > 1) { k_p2p_reduce<Phys<PHYS>::OD><<<nitems, 128, 0, c->stream>>>(a); ONB_LAUNCH(c); }
}
template <int PHYS>
void launch_lists(onb_context* c, const P2PArgs& a, uint32_t nitems) {
    const bool packed = !(g_p2p_tpt & 16);
    int tpt = (g_p2p_tpt & 15) ? (g_p2p_tpt & 15) : (PHYS == ONB_VORTGRAD3D ? 2 : 4);
    // The top levels of the dual tree have a handful of target nodes: a launch that cannot fill the machine is bound by the
    // time ONE warp needs for the longest list, so spread each target block over more warps there (fewer targets per
    // thread). Per-target sums do not depend on the blocking, results stay bit-identical.
    if (!(g_p2p_tpt & 15) && c->arith != ONB_ARITH_STRICT) {
        // launch sizes up to which TPT = 1 / TPT = 2 are used (tunable: ONB_TPT1_MAX, ONB_TPT2_MAX; otherwise from this device's SM count)
        static const int e1 = std::getenv("ONB_TPT1_MAX") ? std::max(1, atoi(std::getenv("ONB_TPT1_MAX"))) : 0;
        static const int e2 = std::getenv("ONB_TPT2_MAX") ? std::max(1, atoi(std::getenv("ONB_TPT2_MAX"))) : 0;
        const uint32_t t1 = e1 ? (uint32_t)e1 : (uint32_t)c->sm_count * 8u, t2 = e2 ? (uint32_t)e2 : 4u * (uint32_t)c->sm_count * 8u;
        const uint32_t ncta = nitems * a.nsplit;
        if (ncta <= t1) tpt = 1; else if (ncta <= t2 && tpt > 2) tpt = 2;
    }
    if (tpt == 1) launch_lists_t<PHYS, 1>(c, a, nitems, packed);
    else if (tpt == 2) launch_lists_t<PHYS, 2>(c, a, nitems, packed);
    else launch_lists_t<PHYS, 4>(c, a, nitems, packed);
}
template <int PHYS>
void launch_lists_a64(onb_context* c, const P2PArgs& a, uint32_t nitems) {
    if (c->arith == ONB_ARITH_STRICT) k_p2p_lists_a64<PHYS, true><<<nitems, 128, 0, c->stream>>>(a);
    else                              k_p2p_lists_a64<PHYS, false><<<nitems, 128, 0, c->stream>>>(a);
}
template <int PHYS>
void launch_direct_a64(onb_context* c, const DirectArgs64& a) {
    const uint32_t grid = (a.nt_eff + 127) / 128;
    if (c->arith == ONB_ARITH_STRICT) k_p2p_direct_a64<PHYS, true><<<grid, 128, 0, c->stream>>>(a);
    else                              k_p2p_direct_a64<PHYS, false><<<grid, 128, 0, c->stream>>>(a);
}
template <int PHYS>
void launch_direct(onb_context* c, const DirectArgs& a) {
    dim3 grid((a.nt_eff + 127) / 128, a.nsplit);
    if (c->arith == ONB_ARITH_STRICT) k_p2p_direct<PHYS, true><<<grid, 128, 0, c->stream>>>(a);
    else                              k_p2p_direct<PHYS, false><<<grid, 128, 0, c->stream>>>(a);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int onb_pack_sources(onb_context* c, DParts& p) {
    if (!p.are_sources) { c->err = "pack: not a source set"; return ONB_ERR_ARG; }
    if (p.n == 0) return ONB_OK;
    if (p.unpacked_released) { c->err = "pack: the source planes were released (lean memory mode): set the sources again"; return ONB_ERR_ARG; }
    if (&p == &c->parts[0]) { int rc = onb_dist_join_source_planes(c, ONB_ST(c)); if (rc) return rc; }     // a multi-GPU plane gather may still be in flight
    const uint32_t cap = p.cap;
    const int T = 256; const uint32_t G = (cap + T - 1) / T;
    PartsView v = view_of(p);
    switch (c->physics) {
        case ONB_GRAV3D:     k_pack<ONB_GRAV3D><<<G, T, 0, ONB_ST(c)>>>(v, p.pk0, p.pk1, p.pk2, p.n, cap); break;
        case ONB_VORT3D:     k_pack<ONB_VORT3D><<<G, T, 0, ONB_ST(c)>>>(v, p.pk0, p.pk1, p.pk2, p.n, cap); break;
        case ONB_VORTGRAD3D: k_pack<ONB_VORTGRAD3D><<<G, T, 0, ONB_ST(c)>>>(v, p.pk0, p.pk1, p.pk2, p.n, cap); break;
        case ONB_VORT2D:     k_pack<ONB_VORT2D><<<G, T, 0, ONB_ST(c)>>>(v, p.pk0, p.pk1, p.pk2, p.n, cap); break;
        default:             k_pack<ONB_VORT2DTR><<<G, T, 0, ONB_ST(c)>>>(v, p.pk0, p.pk1, p.pk2, p.n, cap); break;
    }
    ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    p.packed_valid = true;
    return ONB_OK;
}

int onb_p2p_lists(onb_context* c, const WorkList& wl, int tgt_which_leaf, int tgt_which_box, bool accumulate, uint32_t nsplit) {
    (void)accumulate;   // the kernel always starts from the stored value, i.e. "+=" like ppinter
    if (wl.nitems == 0) return ONB_OK;
    if (nsplit < 1 || c->arith == ONB_ARITH_STRICT) nsplit = 1;     // the strict mode keeps the reference's summation order
    DParts& srcs = c->parts[0]; DParts& eqs = c->parts[2];
    DParts& tl = c->parts[tgt_which_leaf]; DParts& tb = c->parts[tgt_which_box];
    if (!srcs.packed_valid) { int rc = onb_pack_sources(c, srcs); if (rc) return rc; }
    if (eqs.n && !eqs.packed_valid) { int rc = onb_pack_sources(c, eqs); if (rc) return rc; }
    P2PArgs a;
    for (int d = 0; d < 3; ++d) { a.tx[d] = tl.x[d]; a.bx[d] = tb.x[d]; }
    a.tr = tl.r; a.br = tb.r;
    for (int d = 0; d < ONB_MAX_OD; ++d) { a.tu[d] = tl.u[d]; a.bu[d] = tb.u[d]; }
    a.t_ioffset = c->trees[1].ioffset; a.t_num = c->trees[1].num;
    a.s_pk0 = srcs.pk0; a.s_pk1 = srcs.pk1; a.s_pk2 = srcs.pk2;
    a.e_pk0 = eqs.pk0; a.e_pk1 = eqs.pk1; a.e_pk2 = eqs.pk2;
    a.s_ioffset = c->trees[0].ioffset; a.s_num = c->trees[0].num;
    a.item_node = wl.tgt_node; a.start = wl.start; a.entries = wl.entries;
    a.block = c->block; a.ebs = c->ebs; a.num_eqps = c->num_eqps; a.node_base = wl.node_base; a.nentries = (uint32_t)wl.nentries;
    a.s_epnum = c->legacy ? c->d_epnum : nullptr;
    a.nsplit = nsplit; a.partial = nullptr; a.ebase = wl.ebase; a.order = wl.order;
    for (int d = 0; d < ONB_MAX_OD; ++d) { a.tud[d] = tl.ud[d]; a.bud[d] = tb.ud[d]; }
    a.work = nullptr; a.nwork = 0; a.work_counter = nullptr;
    if (!c->accum64) {
        ONB_CUDA(onb_dmalloc(c, (void**)&a.work_counter, 4));
        ONB_CUDA(cudaMemsetAsync(a.work_counter, 0, 4, c->stream));
    }
    if (c->accum64) {
        if (!tl.ud[0] || (tgt_which_box != tgt_which_leaf && tb.n && !tb.ud[0])) { c->err = "ACCUM = double: set the targets (and build the equivalent target points) after onb_set_accum"; return ONB_ERR_ARG; }
        a.nsplit = 1;
        switch (c->physics) {
            case ONB_GRAV3D:     launch_lists_a64<ONB_GRAV3D>(c, a, wl.nitems); break;
            case ONB_VORT3D:     launch_lists_a64<ONB_VORT3D>(c, a, wl.nitems); break;
            case ONB_VORTGRAD3D: launch_lists_a64<ONB_VORTGRAD3D>(c, a, wl.nitems); break;
            case ONB_VORT2D:     launch_lists_a64<ONB_VORT2D>(c, a, wl.nitems); break;
            default:             launch_lists_a64<ONB_VORT2DTR>(c, a, wl.nitems); break;
        }
        ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        return ONB_OK;
    }
    if (nsplit > 1) ONB_CUDA(onb_dmalloc(c, (void**)&a.partial, (size_t)wl.nitems * nsplit * c->OD * 128u * sizeof(float)));
    switch (c->physics) {
        case ONB_GRAV3D:     launch_lists<ONB_GRAV3D>(c, a, wl.nitems); break;
        case ONB_VORT3D:     launch_lists<ONB_VORT3D>(c, a, wl.nitems); break;
        case ONB_VORTGRAD3D: launch_lists<ONB_VORTGRAD3D>(c, a, wl.nitems); break;
        case ONB_VORT2D:     launch_lists<ONB_VORT2D>(c, a, wl.nitems); break;
        default:             launch_lists<ONB_VORT2DTR>(c, a, wl.nitems); break;
    }
    ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    return ONB_OK;
}

int onb_p2p_direct(onb_context* c, uint64_t tskip) {
    DParts& srcs = c->parts[0]; DParts& targs = c->parts[1];
    if (srcs.n == 0 || targs.n == 0) return ONB_OK;
    if (tskip < 1) tskip = 1;
    if (!srcs.packed_valid) { int rc = onb_pack_sources(c, srcs); if (rc) return rc; }
    DirectArgs a;
    for (int d = 0; d < 3; ++d) a.tx[d] = targs.x[d];
    a.tr = targs.r;
    for (int d = 0; d < ONB_MAX_OD; ++d) a.tu[d] = targs.u[d];
    a.s_pk0 = srcs.pk0; a.s_pk1 = srcs.pk1; a.s_pk2 = srcs.pk2;
    a.nsrc = srcs.n; a.ntarg = targs.n; a.tskip = (uint32_t)tskip;
    // i = 0, tskip, 2 tskip, ... < n; a sharded context takes the samples that fall into its target range
    uint32_t lo = 0, hi = targs.n;
    if (c->shard_n > 1) onb_shard_range(c, &lo, &hi);
    a.k0 = (uint32_t)((lo + tskip - 1) / tskip);
    const uint32_t k1 = (uint32_t)((hi + tskip - 1) / tskip);
    a.nt_eff = k1 > a.k0 ? k1 - a.k0 : 0u;
    if (a.nt_eff == 0) { c->last_pairs = 0; return ONB_OK; }
    if (c->accum64) {
        if (!targs.ud[0]) { c->err = "ACCUM = double: set the targets after onb_set_accum"; return ONB_ERR_ARG; }
        DirectArgs64 b;
        for (int d = 0; d < 3; ++d) b.tx[d] = targs.x[d];
        b.tr = targs.r;
        for (int d = 0; d < ONB_MAX_OD; ++d) b.tud[d] = targs.ud[d];
        b.s_pk0 = srcs.pk0; b.s_pk1 = srcs.pk1; b.s_pk2 = srcs.pk2; b.nsrc = srcs.n; b.tskip = (uint32_t)tskip; b.nt_eff = a.nt_eff; b.k0 = a.k0;
        switch (c->physics) {
            case ONB_GRAV3D:     launch_direct_a64<ONB_GRAV3D>(c, b); break;
            case ONB_VORT3D:     launch_direct_a64<ONB_VORT3D>(c, b); break;
            case ONB_VORTGRAD3D: launch_direct_a64<ONB_VORTGRAD3D>(c, b); break;
            case ONB_VORT2D:     launch_direct_a64<ONB_VORT2D>(c, b); break;
            default:             launch_direct_a64<ONB_VORT2DTR>(c, b); break;
        }
        ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
        c->last_pairs = (uint64_t)a.nt_eff * (uint64_t)srcs.n;
        return ONB_OK;
    }
    const uint32_t ntiles = (srcs.n + 127u) / 128u;
    // the source split follows from the GLOBAL sample count, not from this shard's share of it: a target's partial sums are then
    // the same for every rank count, and so is the result, bit for bit
    const uint32_t tgt_blocks = ((uint32_t)((targs.n + tskip - 1) / tskip) + 127u) / 128u;
    uint32_t nsplit = 1;
    if (c->arith != ONB_ARITH_STRICT) {
        // fill the machine: aim for >= 4 CTAs per SM, keep >= 8 tiles per split
        const uint32_t want = (uint32_t)(4 * c->sm_count);
        if (tgt_blocks < want) nsplit = (want + tgt_blocks - 1) / tgt_blocks;
        if (nsplit > (ntiles + 7) / 8) nsplit = (ntiles + 7) / 8;
        if (nsplit < 1) nsplit = 1;
    }
    a.tiles_per_split = (ntiles + nsplit - 1) / nsplit;
    nsplit = (ntiles + a.tiles_per_split - 1) / a.tiles_per_split;
    a.nsplit = nsplit;
    a.partial = nullptr;
    if (nsplit > 1) ONB_CUDA(onb_dmalloc(c, (void**)&a.partial, (size_t)nsplit * c->OD * a.nt_eff * sizeof(float)));
    switch (c->physics) {
        case ONB_GRAV3D:     launch_direct<ONB_GRAV3D>(c, a); break;
        case ONB_VORT3D:     launch_direct<ONB_VORT3D>(c, a); break;
        case ONB_VORTGRAD3D: launch_direct<ONB_VORTGRAD3D>(c, a); break;
        case ONB_VORT2D:     launch_direct<ONB_VORT2D>(c, a); break;
        default:             launch_direct<ONB_VORT2DTR>(c, a); break;
    }
    ONB_LAUNCH(c);
    ONB_CUDA(cudaGetLastError());
    if (nsplit > 1) {
        k_reduce_partials<<<(a.nt_eff + 255) / 256, 256, 0, c->stream>>>(a, c->OD);
        ONB_LAUNCH(c);
        ONB_CUDA(cudaGetLastError());
    }
    c->last_pairs = (uint64_t)a.nt_eff * (uint64_t)srcs.n;
    return ONB_OK;
}

int onb_round_outputs(onb_context* c, DParts& p) {
    if (!p.ud[0] || p.n == 0) return ONB_OK;
    for (int d = 0; d < c->OD; ++d) { k_round_outputs<<<(p.n + 255) / 256, 256, 0, c->stream>>>(p.ud[d], p.u[d], p.n); ONB_LAUNCH(c); }
    ONB_CUDA(cudaGetLastError());
    return ONB_OK;
}

void onb_free_worklist(onb_context* c, WorkList& wl) {
    if (wl.tgt_node) onb_dfree(c, wl.tgt_node);
    if (wl.start) onb_dfree(c, wl.start);
    if (wl.entries) onb_dfree(c, wl.entries);
    wl = WorkList();
}

extern "C" void onb_set_p2p_tpt(int tpt) { const int t = tpt & 15; g_p2p_tpt = (t == 0 || t == 1 || t == 2 || t == 4) ? tpt : 0; }

extern "C" double onb_measure_fp32_peak(onb_context* c) {
    if (!c) return -1.0;
    onb_scratch_reset(c);      // (selects the context's device)
    const int blocks = c->sm_count * 8, threads = 256, iters = 2000;
    float* d = nullptr;
    if (onb_dmalloc(c, (void**)&d, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0, c->stream);
        k_fma_peak<<<blocks, threads, 0, c->stream>>>(d, iters);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        ONB_LAUNCH(c);
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 8.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = flop / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); onb_dfree(c, d);
    return best;
}
