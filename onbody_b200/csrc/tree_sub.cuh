/*
 * tree_sub.cuh - the bottom of the k-d tree build: every level below a node of <= 8192 (or 4096) particles in ONE kernel.
 *
 * Once a node fits in shared memory there is no reason to go back to HBM between levels: one CTA loads the node's
 * coordinates (plus a 16-bit local index), runs ALL remaining levels of splitNode / partialSortIndexes
 * (barneshut.hpp:505-712) on the shared-memory copy and writes the particles out once. For N = 1e7, b = 128 this
 * replaces 7 of the 18 levels' HBM round trips (k_node_split + k_gather per level) by one read and one write.
 *
 * Inside the CTA the levels are processed one after another; at sub-level k the 2^k nodes are handled concurrently
 * by thread groups of 1024 / 2^k threads (never less than a warp) that synchronise with named barriers
 * (bar.sync id, nthreads) or __syncwarp, so every thread is busy at every level. The passes, pivots and exit rules
 * are those of k_node_split (the reference's, in closed form), hence the permutation is bit-identical.
 *
 * Only the coordinates and the composed index plane move here; radius and strength planes are permuted once at
 * the very end of the build through that index (k_apply_perm).
 */
#pragma once

constexpr int SUB_TAB = 512;                  // max nodes of one sub-level
constexpr int SUB_ROUNDS = 4;

struct SubArgs {
    float* x[3];            // current coordinate planes
    float* ox[3];           // final coordinate planes (may alias x)
    const uint32_t* g;      // current composed index plane
    uint32_t* og;           // final index plane (may alias g)
    TreeView t;
    unsigned long long* stats;
    uint32_t block, blo, bhi;
    int level, nsub, PD, pivot_mode;
};

struct SubGrp { int tig, T, W, w0, id; };     // thread in group, threads, warps, first warp (CTA-wide index), barrier id

template <int SUB_T>
__device__ __forceinline__ void sub_sync(const SubGrp& g) {
    if (g.W == 1) __syncwarp();
    else if (g.T == SUB_T) __syncthreads();
    else asm volatile("bar.sync %0, %1;" :: "r"(g.id), "r"(g.T) : "memory");
}

template <int SUB_T, uint32_t SUB_MAX>
__global__ void __launch_bounds__(SUB_T, 8192 / SUB_MAX) k_subtree(const SubArgs a) {
    constexpr int SUB_PER_T = SUB_MAX / SUB_T;
    extern __shared__ __align__(16) unsigned char sub_smem[];
    float* sx[3];
    sx[0] = reinterpret_cast<float*>(sub_smem); sx[1] = sx[0] + SUB_MAX; sx[2] = sx[1] + SUB_MAX;
    uint16_t* sperm = reinterpret_cast<uint16_t*>(sx[2] + SUB_MAX);
    uint16_t* scr = sperm + SUB_MAX;
    __shared__ uint16_t s_pf[2][SUB_TAB], s_n[2][SUB_TAB];
    __shared__ float s_box[32][6];
    __shared__ uint32_t s_cnt[32], s_cnt2[32];
    __shared__ float s_mx[32], s_mn[32], s_piv[32];
    __shared__ unsigned long long s_stats[4];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t R = (1u << a.level) + blockIdx.x;
    const uint32_t n0 = a.t.num[R];
    if (n0 == 0) return;
    const uint32_t pf0 = a.t.ioffset[R];
    const int PD = a.PD;

    if (!(pf0 < a.bhi && pf0 + n0 > a.blo)) {
        // outside this rank's build range: only the (data independent) shape of the subtree is recorded
        if (tid == 0) { s_pf[0][0] = 0; s_n[0][0] = (uint16_t)n0; }
        __syncthreads();
        for (int k = 0; k + 1 < a.nsub; ++k) {
            const int nn = 1 << k, cur = k & 1, nxt = cur ^ 1;
            for (int j = tid; j < nn; j += SUB_T) {
                const uint32_t n = s_n[cur][j], pf = s_pf[cur][j];
                uint32_t nl = 0, nr = 0, pm = pf;
                if (n > a.block) {
                    pm = pf + a.block * (1u << log_2((n - 1) / a.block));
                    nl = pm - pf; nr = pf + n - pm;
                    const uint32_t node = (R << k) + (uint32_t)j;
                    a.t.ioffset[2 * node] = pf0 + pf;     a.t.num[2 * node] = nl;
                    a.t.ioffset[2 * node + 1] = pf0 + pm; a.t.num[2 * node + 1] = nr;
                }
                s_pf[nxt][2 * j] = (uint16_t)pf; s_n[nxt][2 * j] = (uint16_t)nl;
                s_pf[nxt][2 * j + 1] = (uint16_t)pm; s_n[nxt][2 * j + 1] = (uint16_t)nr;
            }
            __syncthreads();
        }
        return;
    }

    // ---- load the node ----
    for (int d = 0; d < PD; ++d) {
        const float* __restrict__ xd = a.x[d] + pf0;
        #pragma unroll
        for (int r = 0; r < SUB_PER_T; ++r) { const uint32_t i = (uint32_t)r * SUB_T + tid; if (i < n0) sx[d][i] = xd[i]; }
    }
    #pragma unroll
    for (int r = 0; r < SUB_PER_T; ++r) { const uint32_t i = (uint32_t)r * SUB_T + tid; if (i < n0) sperm[i] = (uint16_t)i; }
    if (tid == 0) { s_pf[0][0] = 0; s_n[0][0] = (uint16_t)n0; }
    if (tid < 4) s_stats[tid] = 0ull;
    __syncthreads();

    for (int k = 0; k < a.nsub; ++k) {
        const int nn = 1 << k, cur = k & 1, nxt = cur ^ 1;
        SubGrp g;
        // group size: 1024, 512, 256, 128, 128 (two nodes per group), then one warp per node. Named barriers 1..8 only:
        // barrier 0 belongs to __syncthreads and there are 16 barriers per CTA.
        g.T = SUB_T >> k; if (g.T == 64) g.T = 128; if (g.T < 32) g.T = 32;
        g.W = g.T >> 5;
        const int NG = SUB_T / g.T, gi = tid / g.T;
        g.tig = tid - gi * g.T; g.w0 = gi * g.W; g.id = 1 + gi;
        const int wig = warp - g.w0;          // warp within the group
        const bool last_level = (k + 1 == a.nsub);

        for (int j = gi; j < nn; j += NG) {
            const uint32_t n = s_n[cur][j], pf = s_pf[cur][j], pl = pf + n;
            const uint32_t node = (R << k) + (uint32_t)j;
            if (n == 0) {
                if (!last_level && g.tig == 0) { s_n[nxt][2 * j] = 0; s_n[nxt][2 * j + 1] = 0; s_pf[nxt][2 * j] = 0; s_pf[nxt][2 * j + 1] = 0; }
                continue;
            }
            // bounding box :621-625 (exact, order independent)
            float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
            for (uint32_t i = pf + g.tig; i < pl; i += g.T) {
                #pragma unroll
                for (int d = 0; d < 3; ++d) if (d < PD) { const float v = sx[d][i]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
            }
            #pragma unroll
            for (int d = 0; d < 3; ++d) { lo[d] = warp_min(lo[d]); hi[d] = warp_max(hi[d]); }
            if (g.W > 1) {
                if (lane == 0) {
                    #pragma unroll
                    for (int d = 0; d < 3; ++d) { s_box[warp][d] = lo[d]; s_box[warp][3 + d] = hi[d]; }
                }
                sub_sync<SUB_T>(g);
                #pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float l2 = lane < g.W ? s_box[g.w0 + lane][d] : INFINITY, h2 = lane < g.W ? s_box[g.w0 + lane][3 + d] : -INFINITY;
                    lo[d] = warp_min(l2); hi[d] = warp_max(h2);
                }
                sub_sync<SUB_T>(g);
            }
            if (g.tig == 0) {
                float bsss = 0.0f;
                #pragma unroll
                for (int d = 0; d < 3; ++d) if (d < PD) {
                    const float ns = __fsub_rn(hi[d], lo[d]);
                    a.t.ns[d][node] = ns;
                    a.t.nc[d][node] = __fmul_rn(0.5f, __fadd_rn(hi[d], lo[d]));
                    bsss = __double2float_rn(__dadd_rn((double)bsss, __dmul_rn((double)ns, (double)ns)));   // std::pow(float,int) is double :638
                }
                a.t.nr[node] = __fmul_rn(0.5f, __fsqrt_rn(bsss));
            }
            if (n <= a.block || last_level) {                                                    // :644 leaf
                if (!last_level && g.tig == 0) { s_n[nxt][2 * j] = 0; s_n[nxt][2 * j + 1] = 0; s_pf[nxt][2 * j] = 0; s_pf[nxt][2 * j + 1] = 0; }
                continue;
            }
            // longest axis :652-659 (first strict maximum)
            int axis = 0; float axsz = -1.0f;
            #pragma unroll
            for (int d = 0; d < 3; ++d) if (d < PD) { const float ns = __fsub_rn(hi[d], lo[d]); if (ns > axsz) { axsz = ns; axis = d; } }
            const uint32_t nless = pf + a.block * (1u << log_2((n - 1) / a.block));              // :663
            float* key = axis == 0 ? sx[0] : (axis == 1 ? sx[1] : sx[2]);

            // partial select :519-586 on the shared-memory copy; every thread of the group tracks the window
            uint32_t wf = pf, wl = pl - 1;
            float wlo = axis == 0 ? lo[0] : (axis == 1 ? lo[1] : lo[2]), whi = axis == 0 ? hi[0] : (axis == 1 ? hi[1] : hi[2]);
            const float ideal = __fdiv_rn(__uint2float_rn(nless - pf), __uint2float_rn(pl - pf));     // :522
            int iters = 0;
            uint32_t n_pass = 0, n_stall = 0, n_scan = 0;
            // The pivot (:538-540, double arithmetic with a software division; index differences only) is evaluated by the
            // group's first warp alone and published through shared memory across a barrier the pass needs anyway: the
            // kernel is issue-bound and every warp repeating those ~200 instructions per pass was a quarter of its work.
            bool go = wl > wf;
            float pivot = 0.f;
            if (g.W > 1) {
                if (wig == 0 && go) { const float pv = select_pivot(nless, wf, wl, wlo, whi, ideal, a.pivot_mode); if (lane == 0) s_piv[g.w0] = pv; }
                sub_sync<SUB_T>(g);
                if (go) pivot = s_piv[g.w0];
            } else if (go) pivot = select_pivot(nless, wf, wl, wlo, whi, ideal, a.pivot_mode);
            while (go) {
                // pass 1: m = #{v < pivot}, and the min/max the two possible next windows will have
                uint32_t cnt = 0; float mx_lt = -INFINITY, mn_ge = INFINITY;
                for (uint32_t i = wf + g.tig; i <= wl; i += g.T) {
                    const float v = key[i];
                    if (v < pivot) { ++cnt; mx_lt = fmaxf(mx_lt, v); } else mn_ge = fminf(mn_ge, v);
                }
                cnt = warp_sum(cnt); mx_lt = warp_max(mx_lt); mn_ge = warp_min(mn_ge);
                if (g.W > 1) {
                    if (lane == 0) { s_cnt[warp] = cnt; s_mx[warp] = mx_lt; s_mn[warp] = mn_ge; }
                    sub_sync<SUB_T>(g);
                    const uint32_t c2 = lane < g.W ? s_cnt[g.w0 + lane] : 0u;
                    const float a2 = lane < g.W ? s_mx[g.w0 + lane] : -INFINITY, b2 = lane < g.W ? s_mn[g.w0 + lane] : INFINITY;
                    cnt = warp_sum(c2); mx_lt = warp_max(a2); mn_ge = warp_min(b2);
                    sub_sync<SUB_T>(g);
                }
                const uint32_t B = wf + cnt;

                // pass 2: ordered compaction of the misplaced positions (a ascending from scr[pf], b from scr[pl-1] downwards)
                uint32_t carryA = 0, carryB = 0;
                const uint32_t chunk = (uint32_t)g.T * SUB_ROUNDS;
                for (uint32_t base = wf; base <= wl; base += chunk) {
                    uint32_t ba[SUB_ROUNDS], bb[SUB_ROUNDS];
                    uint32_t totA = 0, totB = 0;
                    const bool warp_in = base + (uint32_t)wig * 32u * SUB_ROUNDS <= wl;       // (warp-uniform) anything of the window here?
                    #pragma unroll
                    for (int r = 0; r < SUB_ROUNDS; ++r) {
                        ba[r] = 0u; bb[r] = 0u;
                        if (warp_in) {
                            const uint32_t i = base + (uint32_t)wig * 32u * SUB_ROUNDS + (uint32_t)r * 32u + lane;
                            const bool valid = i <= wl;
                            const float v = valid ? key[i] : 0.f;
                            const bool lt = v < pivot;
                            ba[r] = __ballot_sync(0xffffffffu, valid && i < B && !lt);
                            bb[r] = __ballot_sync(0xffffffffu, valid && i >= B && lt);
                            totA += __popc(ba[r]); totB += __popc(bb[r]);
                        }
                    }
                    uint32_t exA = 0, exB = 0, allA = totA, allB = totB;
                    if (g.W > 1) {
                        if (lane == 0) { s_cnt[warp] = totA; s_cnt2[warp] = totB; }
                        sub_sync<SUB_T>(g);
                        const uint32_t ca = lane < g.W ? s_cnt[g.w0 + lane] : 0u, cb = lane < g.W ? s_cnt2[g.w0 + lane] : 0u;
                        allA = warp_sum(ca); allB = warp_sum(cb);
                        exA = warp_sum(lane < wig ? ca : 0u); exB = warp_sum(lane < wig ? cb : 0u);
                    }
                    uint32_t offA = carryA + exA, offB = carryB + exB;
                    const uint32_t lt_mask = (1u << lane) - 1u;
                    #pragma unroll
                    for (int r = 0; r < SUB_ROUNDS; ++r) {
                        const uint32_t i = base + (uint32_t)wig * 32u * SUB_ROUNDS + (uint32_t)r * 32u + lane;
                        if ((ba[r] >> lane) & 1u) scr[pf + offA + __popc(ba[r] & lt_mask)] = (uint16_t)i;
                        if ((bb[r] >> lane) & 1u) scr[pl - 1u - (offB + __popc(bb[r] & lt_mask))] = (uint16_t)i;
                        offA += __popc(ba[r]); offB += __popc(bb[r]);
                    }
                    carryA += allA; carryB += allB;
                    sub_sync<SUB_T>(g);
                }
                const uint32_t kk = carryA;    // == carryB
                // pass 3: the swaps :549-556 - coordinates and index move together
                for (uint32_t q = g.tig; q < kk; q += g.T) {
                    const uint32_t pa = scr[pf + q], pb = scr[pl - kk + q];
                    #pragma unroll
                    for (int d = 0; d < 3; ++d) if (d < PD) { const float va = sx[d][pa], vb = sx[d][pb]; sx[d][pa] = vb; sx[d][pb] = va; }
                    const uint16_t ia = sperm[pa], ib = sperm[pb]; sperm[pa] = ib; sperm[pb] = ia;
                }
                ++n_pass; n_scan += (wl - wf + 1);
                // :565-583 - the next window follows from B and the two extrema, all known since the count
                if (B == nless) go = false;
                else {
                    const uint32_t owf = wf, owl = wl;
                    if (B < nless) { wf = B; wlo = mn_ge; } else { wl = B - 1; whi = mx_lt; }
                    if (wf == owf && wl == owl) { ++n_stall; go = false; }
                    else { ++iters; go = wl > wf && iters < 100; }                                // loop condition :527
                }
                if (go && g.W > 1 && wig == 0) { const float pv = select_pivot(nless, wf, wl, wlo, whi, ideal, a.pivot_mode); if (lane == 0) s_piv[g.w0] = pv; }
                sub_sync<SUB_T>(g);                      // swaps done; next pivot published
                if (go) pivot = g.W > 1 ? s_piv[g.w0] : select_pivot(nless, wf, wl, wlo, whi, ideal, a.pivot_mode);
            }
            if (g.tig == 0) {
                a.t.ioffset[2 * node] = pf0 + pf;        a.t.num[2 * node] = nless - pf;           // :702-704
                a.t.ioffset[2 * node + 1] = pf0 + nless; a.t.num[2 * node + 1] = pl - nless;
                if (!last_level) {
                    s_pf[nxt][2 * j] = (uint16_t)pf;        s_n[nxt][2 * j] = (uint16_t)(nless - pf);
                    s_pf[nxt][2 * j + 1] = (uint16_t)nless; s_n[nxt][2 * j + 1] = (uint16_t)(pl - nless);
                }
                atomicAdd(&s_stats[0], 1ull); atomicAdd(&s_stats[1], (unsigned long long)n_pass);
                atomicAdd(&s_stats[2], (unsigned long long)n_stall); atomicAdd(&s_stats[3], (unsigned long long)n_scan);
            }
        }
        __syncthreads();
    }

    // ---- write the node back: coordinates from shared memory, the index plane through the local permutation ----
    for (int d = 0; d < PD; ++d) {
        float* __restrict__ od = a.ox[d] + pf0;
        #pragma unroll
        for (int r = 0; r < SUB_PER_T; ++r) { const uint32_t i = (uint32_t)r * SUB_T + tid; if (i < n0) od[i] = sx[d][i]; }
    }
    uint32_t gv[SUB_PER_T];
    #pragma unroll
    for (int r = 0; r < SUB_PER_T; ++r) { const uint32_t i = (uint32_t)r * SUB_T + tid; gv[r] = i < n0 ? a.g[pf0 + sperm[i]] : 0u; }
    __syncthreads();            // g and og may be the same plane: every read before any write
    #pragma unroll
    for (int r = 0; r < SUB_PER_T; ++r) { const uint32_t i = (uint32_t)r * SUB_T + tid; if (i < n0) a.og[pf0 + i] = gv[r]; }
    if (tid < 4 && s_stats[tid]) atomicAdd(&a.stats[tid], s_stats[tid]);
}

// ---- the planes that did not travel with the levels: dst[i] = src[g[i]] over the span the build touched ----
struct PermArgs { const float* src[4]; float* dst[4]; const uint32_t* g; uint32_t lo, hi; int nplanes; };
__global__ void k_apply_perm(const PermArgs a) {
    const uint32_t i = a.lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.hi) return;
    const uint32_t j = a.g[i];
    #pragma unroll
    for (int q = 0; q < 4; ++q) if (q < a.nplanes) a.dst[q][i] = a.src[q][j];
}
__global__ void k_copy_back(const PermArgs a) {
    const uint32_t i = a.lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.hi) return;
    #pragma unroll
    for (int q = 0; q < 4; ++q) if (q < a.nplanes) a.dst[q][i] = a.src[q][i];
}
