/*
 * plan.cu - host-only arithmetic of the multi-GPU partition (no device code, no CUDA calls: usable without a GPU).
 *
 * The shape of the reference's VAM-split tree (Tree.hpp:83-87, barneshut.hpp:663) depends on the particle count and the
 * block size only: node i holds particles [ioffset, ioffset+num), a node with num > block is cut at
 * ioffset + block * 2^floor(log2((num-1)/block)). Every leaf but the last holds exactly `block` particles, so rank r of R
 * owns the contiguous leaves [r*per, (r+1)*per), per = ceil(nleaf/R): equal chunks, which is what lets every plane exchange
 * be ONE in-place all-gather (comm.cu). From the shape alone this file derives, per tree level,
 *   own[l]    the node interval that lies completely inside the rank's particle range (the rank builds, summarises and
 *             anterpolates these without any communication),
 *   need[l]   the node interval that overlaps the range (the target nodes the rank evaluates; own + the straddling ones),
 *   shared[l] the non-leaf nodes that straddle a rank boundary (at most R-1 per level; every rank recomputes them from
 *             exchanged children).
 */
#include "onb_internal.h"
#include <algorithm>

static inline uint32_t plan_log2(uint32_t x) { return x == 0 ? 0 : 31 - __builtin_clz(x); }

int onb_plan_levels(uint64_t n, int block) {
    if (n == 0 || block < 1) return 0;
    const uint32_t numLeaf = (uint32_t)(1 + (n - 1) / (uint64_t)block);                  // Tree.hpp:83-87
    return 1 + (int)plan_log2(2 * numLeaf - 1);
}

int onb_shard_range_for(uint64_t n, int block, int rank, int nranks, uint64_t* lo, uint64_t* hi) {
    if (nranks < 1 || rank < 0 || rank >= nranks || block < 1) return ONB_ERR_ARG;
    const uint64_t nleaf = (n + block - 1) / block;
    const uint64_t per = (nleaf + nranks - 1) / nranks;
    *lo = std::min<uint64_t>(per * (uint64_t)rank * block, n);
    *hi = std::min<uint64_t>(per * (uint64_t)(rank + 1) * block, n);
    return ONB_OK;
}

uint64_t onb_shard_chunk(uint64_t n, int block, int nranks) {
    const uint64_t nleaf = (n + block - 1) / block;
    return ((nleaf + nranks - 1) / nranks) * (uint64_t)block;
}

int onb_plan_make(ShardPlan& P, uint64_t n, int block, int nranks, int rank) {
    if (n == 0 || block < 1 || nranks < 1 || rank < 0 || rank >= nranks || nranks > ONB_MAX_RANKS) return ONB_ERR_ARG;
    if (P.valid && P.n == n && P.block == block && P.nranks == nranks && P.rank == rank) return ONB_OK;
    P = ShardPlan();
    P.n = n; P.block = block; P.nranks = nranks; P.rank = rank;
    P.levels = onb_plan_levels(n, block);
    P.chunk = onb_shard_chunk(n, block, nranks);
    onb_shard_range_for(n, block, rank, nranks, &P.lo, &P.hi);
    const int L = P.levels;
    P.own_lo.assign(L, 0); P.own_hi.assign(L, 0); P.need_lo.assign(L, 0); P.need_hi.assign(L, 0);
    P.all_own_lo.assign((size_t)L * nranks, 0); P.all_own_hi.assign((size_t)L * nranks, 0);
    P.shared.assign(L, std::vector<uint32_t>());
    // one level at a time: (first particle, count) of the existing nodes, left to right
    std::vector<uint64_t> io(1, 0), num(1, n), nio, nnum;
    for (int l = 0; l < L; ++l) {
        const uint32_t base = 1u << l;
        const size_t cnt = io.size();                 // existing nodes of this level are the ids base .. base+cnt-1
        for (int r = 0; r < nranks; ++r) {
            uint64_t rlo, rhi; onb_shard_range_for(n, block, r, nranks, &rlo, &rhi);
            // nodes are disjoint and ordered: first node starting at or after rlo, first node ending after rhi
            size_t a = std::lower_bound(io.begin(), io.end(), rlo) - io.begin();
            size_t b = a;
            while (b < cnt && io[b] + num[b] <= rhi) ++b;
            if (rhi <= rlo) { a = b = 0; }
            P.all_own_lo[(size_t)l * nranks + r] = base + (uint32_t)a; P.all_own_hi[(size_t)l * nranks + r] = base + (uint32_t)std::max(a, b);
            if (r == rank) {
                P.own_lo[l] = base + (uint32_t)a; P.own_hi[l] = base + (uint32_t)std::max(a, b);
                // overlapping nodes: from the node containing rlo to the node containing rhi-1
                size_t na = std::upper_bound(io.begin(), io.end(), rlo) - io.begin();   // first node starting after rlo
                na = na > 0 ? na - 1 : 0;
                if (io[na] + num[na] <= rlo) ++na;                                       // (cannot happen: nodes tile [0,n) at every level they exist)
                size_t nb = na;
                while (nb < cnt && io[nb] < rhi) ++nb;
                if (rhi <= rlo) { na = nb = 0; }
                P.need_lo[l] = base + (uint32_t)na; P.need_hi[l] = base + (uint32_t)nb;
            }
        }
        // straddling non-leaf nodes: not inside any single rank's range
        for (size_t k = 0; k < cnt; ++k) {
            if (num[k] <= (uint64_t)block) continue;
            const uint64_t r0 = io[k] / P.chunk, r1 = (io[k] + num[k] - 1) / P.chunk;
            if (r0 != r1) P.shared[l].push_back(base + (uint32_t)k);
        }
        // next level
        nio.clear(); nnum.clear();
        for (size_t k = 0; k < cnt; ++k) {
            if (num[k] <= (uint64_t)block) continue;                                      // leaves have no children
            const uint64_t pm = io[k] + (uint64_t)block * (1ull << plan_log2((uint32_t)((num[k] - 1) / block)));   // barneshut.hpp:663
            nio.push_back(io[k]); nnum.push_back(pm - io[k]);
            nio.push_back(pm);    nnum.push_back(io[k] + num[k] - pm);
        }
        // children ids are 2i, 2i+1: the non-leaf nodes of a level are a prefix of it (sizes never grow left to right), so the
        // existing nodes of the next level are again the ids base' .. base'+cnt'-1
        for (size_t k = 0; k + 1 < cnt; ++k) if (num[k] <= (uint64_t)block && num[k + 1] > (uint64_t)block) return ONB_ERR_UNSUPPORTED;
        io.swap(nio); num.swap(nnum);
        if (io.empty()) break;
    }
    P.valid = true;
    return ONB_OK;
}

extern "C" {

uint64_t onb_shard_chunk_for(uint64_t n, int block, int nranks) { return (nranks < 1 || block < 1) ? 0 : onb_shard_chunk(n, block, nranks); }

int onb_plan_query(uint64_t n, int block, int nranks, int rank, int max_levels, uint32_t* own_lo, uint32_t* own_hi,
                   uint32_t* need_lo, uint32_t* need_hi, uint32_t* nshared, uint32_t* shared) {
    ShardPlan P;
    const int rc = onb_plan_make(P, n, block, nranks, rank);
    if (rc) return -rc;
    if (P.levels > max_levels) return -ONB_ERR_ARG;
    for (int l = 0; l < P.levels; ++l) {
        own_lo[l] = P.own_lo[l]; own_hi[l] = P.own_hi[l]; need_lo[l] = P.need_lo[l]; need_hi[l] = P.need_hi[l];
        nshared[l] = (uint32_t)P.shared[l].size();
        for (size_t k = 0; k < P.shared[l].size() && k < (size_t)nranks; ++k) shared[(size_t)l * nranks + k] = P.shared[l][k];
    }
    return P.levels;
}

}  // extern "C"
