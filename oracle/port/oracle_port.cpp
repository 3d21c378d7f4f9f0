/*
 * oracle_port.cpp - CPU restatement ("port") of onbody's summation hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_port.h). Compile with -O2 -ffp-contract=off (oracle/Makefile):
 * every float expression below is then evaluated exactly as written, which is how the parity oracle
 * oracle/_ref/strict (the unmodified reference, same flags) evaluates the reference source. tests/ pin
 * this port to that build array-for-array, bit-for-bit, and to the golden vectors in tests/golden/.
 *
 * This is a restatement, not a copy: the reference's recursive OpenMP-task formulations are rewritten
 * here in the LEVEL-SYNCHRONOUS / CLOSED forms the CUDA kernels use, so that the
 * equivalence "recursive reference == data-parallel formulation" is itself what the tests prove on CPU.
 * Citations are file:line into /root/reference/src.
 */
#include "oracle_port.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

namespace {

enum { GRAV3D = 0, VORT3D = 1, VORTGRAD3D = 2, VORT2D = 3, VORT2DTR = 4 };

struct PParts {                     // Parts.hpp:32-74
    bool are_sources = false;
    size_t n = 0, blockSize = 128;
    int PD = 3, SD = 1, OD = 3;
    std::vector<float> x[3], r, s[3], u[12];
    std::vector<uint64_t> gidx;
    void resize(size_t nn) {        // Parts.hpp:85-92
        n = nn;
        for (int d = 0; d < PD; ++d) x[d].resize(n);
        if (are_sources) for (int d = 0; d < SD; ++d) s[d].resize(n);
        r.resize(n);
        if (!are_sources) for (int d = 0; d < OD; ++d) u[d].resize(n);
    }
};

struct PTree {                      // Tree.hpp:44-76
    int levels = 0, numnodes = 0;
    std::vector<float> x[3], nc[3], ns[3], nr, pr, s[3];
    std::vector<uint64_t> ioffset, num, epoffset, epnum;
};

inline uint32_t log_2(uint32_t v) { return v == 0 ? 0 : 31 - __builtin_clz(v); }   // Tree.hpp:30-33

void tree_alloc(PTree& t, int PD, int SD, size_t n, size_t bs) {                    // Tree.hpp:79-107
    uint32_t numLeaf = 1 + (uint32_t)((n - 1) / bs);
    t.levels = 1 + log_2(2 * numLeaf - 1);
    t.numnodes = 1 << t.levels;
    const size_t nn = t.numnodes;
    for (int d = 0; d < PD; ++d) { t.x[d].assign(nn, 0.f); t.nc[d].assign(nn, 0.f); t.ns[d].assign(nn, 0.f); }
    t.nr.assign(nn, 0.f); t.pr.assign(nn, 0.f);
    for (int d = 0; d < SD; ++d) t.s[d].assign(nn, 0.f);
    t.ioffset.assign(nn, 0); t.num.assign(nn, 0); t.epoffset.assign(nn, 0); t.epnum.assign(nn, 0);
}

struct Session {
    int physics, PD, SD, OD, order, flops;
    bool has_tr, has_fastsumm;
    PParts srcs, targs, eqsrcs, eqtargs;
    PTree stree, ttree;
    uint64_t stats[9] = {0};
    uint64_t bstats[4] = {0};
    uint64_t tie_sorts = 0;
};

// ---------------------------------------------------------------------------------------------
// a-3: the pair kernels, one source on one target, in the reference's own operation order
// ---------------------------------------------------------------------------------------------
struct Src { float x, y, z, r, s0, s1, s2; };

// ongrav3d.cpp:44-58
inline void pair_grav3d(const Src& s, float tx, float ty, float tz, float* u) {
    const float dx = s.x - tx, dy = s.y - ty, dz = s.z - tz;
    float r3 = dx*dx + dy*dy + dz*dz + s.r*s.r;
    r3 = s.s0 / (r3 * std::sqrt(r3));
    u[0] += r3 * dx; u[1] += r3 * dy; u[2] += r3 * dz;
}
// onvort3d.cpp:44-59 with core_func CoreFunc3d.hpp:27-30 and oor1p5 MathHelper.hpp:214-217
inline void pair_vort3d(const Src& s, float tx, float ty, float tz, float* u) {
    const float dx = s.x - tx, dy = s.y - ty, dz = s.z - tz;
    const float r2 = (dx*dx + dy*dy + dz*dz) + s.r*s.r;
    const float r3 = 1.0f / (r2 * std::sqrt(r2));
    const float dxxw = dz*s.s1 - dy*s.s2;
    const float dyxw = dx*s.s2 - dz*s.s0;
    const float dzxw = dy*s.s0 - dx*s.s1;
    u[0] += r3 * dxxw; u[1] += r3 * dyxw; u[2] += r3 * dzxw;
}
// onvortgrad3d.cpp:45-76 with core_func CoreFunc3d.hpp:34-40 (note d = target - source here)
inline void pair_vortgrad3d(const Src& s, float tx, float ty, float tz, float* u) {
    const float dx = tx - s.x, dy = ty - s.y, dz = tz - s.z;
    const float r2 = (dx*dx + dy*dy + dz*dz) + s.r*s.r;
    const float r3 = 1.0f / (r2 * std::sqrt(r2));
    const float bbb = -3.0f * r3 * (1.0f / r2);
    float dxxw = dz*s.s1 - dy*s.s2;
    float dyxw = dx*s.s2 - dz*s.s0;
    float dzxw = dy*s.s0 - dx*s.s1;
    u[0] += r3 * dxxw; u[1] += r3 * dyxw; u[2] += r3 * dzxw;
    dxxw *= bbb; dyxw *= bbb; dzxw *= bbb;
    u[3]  += dx*dxxw;
    u[4]  += dx*dyxw + s.s2*r3;
    u[5]  += dx*dzxw - s.s1*r3;
    u[6]  += dy*dxxw - s.s2*r3;
    u[7]  += dy*dyxw;
    u[8]  += dy*dzxw + s.s0*r3;
    u[9]  += dz*dxxw + s.s1*r3;
    u[10] += dz*dyxw - s.s0*r3;
    u[11] += dz*dzxw;
}
// interface2dvort.cpp:39-50 with core_func CoreFunc2d.hpp:24-28
inline void pair_vort2d(const Src& s, float tx, float ty, float* u) {
    const float dx = tx - s.x, dy = ty - s.y;
    const float r2c = (dx*dx + dy*dy) + s.r*s.r;
    const float r2 = s.s0 * (1.0f / r2c);
    u[0] -= r2 * dy; u[1] += r2 * dx;
}
// onvort2d.cpp:44-55 with core_func CoreFunc2d.hpp:31-35
inline void pair_vort2dtr(const Src& s, float tx, float ty, float tr, float* u) {
    const float dx = tx - s.x, dy = ty - s.y;
    const float r2c = (dx*dx + dy*dy) + s.r*s.r + tr*tr;
    const float r2 = s.s0 * (1.0f / r2c);
    u[0] -= r2 * dy; u[1] += r2 * dx;
}

inline void pair_any(int physics, const Src& s, const float* t /*x,y,z,r*/, float* u) {
    switch (physics) {
        case GRAV3D:     pair_grav3d(s, t[0], t[1], t[2], u); break;
        case VORT3D:     pair_vort3d(s, t[0], t[1], t[2], u); break;
        case VORTGRAD3D: pair_vortgrad3d(s, t[0], t[1], t[2], u); break;
        case VORT2D:     pair_vort2d(s, t[0], t[1], u); break;
        default:         pair_vort2dtr(s, t[0], t[1], t[3], u); break;
    }
}

inline Src load_src(const PParts& p, size_t j) {
    Src s;
    s.x = p.x[0][j]; s.y = p.x[1][j]; s.z = p.PD > 2 ? p.x[2][j] : 0.f; s.r = p.r[j];
    s.s0 = p.s[0][j]; s.s1 = p.SD > 1 ? p.s[1][j] : 0.f; s.s2 = p.SD > 2 ? p.s[2][j] : 0.f;
    return s;
}

// ppinter, block of sources on a block of targets: target loop outside, source loop inside,
// accumulating straight into the target's outputs (ongrav3d.cpp:162-168)
void ppinter(int physics, const PParts& sp, size_t jstart, size_t jend, PParts& tp, size_t istart, size_t iend) {
    for (size_t i = istart; i < iend; ++i) {
        float t[4] = { tp.x[0][i], tp.x[1][i], tp.PD > 2 ? tp.x[2][i] : 0.f, tp.r[i] };
        float u[12];
        for (int d = 0; d < tp.OD; ++d) u[d] = tp.u[d][i];
        for (size_t j = jstart; j < jend; ++j) pair_any(physics, load_src(sp, j), t, u);
        for (int d = 0; d < tp.OD; ++d) tp.u[d][i] = u[d];
    }
}

// tpinter: a tree node as one particle (ongrav3d.cpp:174-181)
void tpinter(int physics, const PTree& st, size_t j, PParts& tp, size_t i) {
    Src s;
    s.x = st.x[0][j]; s.y = st.x[1][j]; s.z = tp.PD > 2 ? st.x[2][j] : 0.f; s.r = st.pr[j];
    s.s0 = st.s[0][j]; s.s1 = tp.SD > 1 ? st.s[1][j] : 0.f; s.s2 = tp.SD > 2 ? st.s[2][j] : 0.f;
    float t[4] = { tp.x[0][i], tp.x[1][i], tp.PD > 2 ? tp.x[2][i] : 0.f, tp.r[i] };
    float u[12];
    for (int d = 0; d < tp.OD; ++d) u[d] = tp.u[d][i];
    pair_any(physics, s, t, u);
    for (int d = 0; d < tp.OD; ++d) tp.u[d][i] = u[d];
}

// ---------------------------------------------------------------------------------------------
// a-1: VAM-split k-d tree, level-synchronous, with the Hoare pass in closed form
// ---------------------------------------------------------------------------------------------

// one partial select on v[istart,istop) so that [istart,nless) holds the nless-istart smallest-side
// elements; idx[i] = position (before the select) of the element now at i. barneshut.hpp:505-587.
void partial_select(std::vector<float>& v, std::vector<uint64_t>& idx, float lo, float hi,
                    size_t istart, size_t nless, size_t istop, uint64_t* bstats) {
    for (size_t i = istart; i < istop; ++i) idx[i] = i;                                  // :516
    size_t wf = istart, wl = istop - 1;                                                  // :519-520
    int iters = 0;
    const float ideal = (float)(nless - istart) / (float)(istop - istart);               // :522
    std::vector<size_t> apos, bpos;
    bstats[0]++;
    while (wl > wf && iters < 100) {                                                     // :527
        if (iters > 0) {                                                                 // :529-531 re-minmax of the window
            lo = v[wf]; hi = v[wf];
            for (size_t i = wf; i <= wl; ++i) { if (v[i] < lo) lo = v[i]; if (v[i] > hi) hi = v[i]; }
        }
        float frac = (float)((double)nless - 0.5 - (double)wf) / (float)(wl - wf);       // :538
        frac = (float)((9.0 * (double)frac + 1.0 * (double)ideal) / 10.0);               // :539
        const float pivot = lo + (hi - lo) * frac;                                       // :540
        // closed form of the two-pointer march :546-560
        size_t m = 0;
        for (size_t i = wf; i <= wl; ++i) m += (v[i] < pivot);
        const size_t B = wf + m;
        apos.clear(); bpos.clear();
        for (size_t i = wf; i < B; ++i) if (!(v[i] < pivot)) apos.push_back(i);          // ascending
        for (size_t i = wl + 1; i-- > B; ) if (v[i] < pivot) bpos.push_back(i);          // descending
        for (size_t j = 0; j < apos.size(); ++j) {
            std::swap(v[apos[j]], v[bpos[j]]);
            std::swap(idx[apos[j]], idx[bpos[j]]);
        }
        bstats[1]++; bstats[3] += (wl - wf + 1);
        // :565-583
        if (B == nless) break;
        const size_t owf = wf, owl = wl;
        if (B < nless) wf = B; else wl = B - 1;
        if (wf == owf && wl == owl) { bstats[2]++; break; }
        iters++;
    }
}

void gather_f(std::vector<float>& a, std::vector<float>& tmp, const std::vector<uint64_t>& idx, size_t pf, size_t pl) {
    std::copy(a.begin() + pf, a.begin() + pl, tmp.begin() + pf);                         // barneshut.hpp:475-485
    for (size_t i = pf; i < pl; ++i) a[i] = tmp[idx[i]];
}
void gather_u(std::vector<uint64_t>& a, std::vector<uint64_t>& tmp, const std::vector<uint64_t>& idx, size_t pf, size_t pl) {
    std::copy(a.begin() + pf, a.begin() + pl, tmp.begin() + pf);                         // Parts.hpp:188-196
    for (size_t i = pf; i < pl; ++i) a[i] = tmp[idx[i]];
}

void make_tree(Session& S, PParts& p, PTree& t) {
    const int PD = p.PD, SD = p.SD;
    const size_t bs = p.blockSize;
    tree_alloc(t, PD, SD, p.n, bs);                                                      // barneshut.hpp:826
    p.gidx.resize(p.n);
    for (size_t i = 0; i < p.n; ++i) p.gidx[i] = i;                                      // :823
    std::vector<uint64_t> lidx(p.n), itemp(p.n);
    std::vector<float> ftemp(p.n);
    memset(S.bstats, 0, sizeof(S.bstats));

    t.ioffset[1] = 0; t.num[1] = p.n;
    // splitNode, one level at a time: children ranges are disjoint so any schedule gives the same arrays
    for (int lev = 0; lev < t.levels; ++lev) {
        for (size_t node = (size_t)1 << lev; node < ((size_t)2 << lev); ++node) {
            if (lev > 0 && t.num[node] == 0) continue;
            const size_t pf = t.ioffset[node], pl = pf + t.num[node];
            float lo[3], hi[3];
            for (int d = 0; d < PD; ++d) {                                               // :621-625
                lo[d] = p.x[d][pf]; hi[d] = p.x[d][pf];
                for (size_t i = pf; i < pl; ++i) { if (p.x[d][i] < lo[d]) lo[d] = p.x[d][i]; if (p.x[d][i] > hi[d]) hi[d] = p.x[d][i]; }
                t.ns[d][node] = hi[d] - lo[d];
                t.nc[d][node] = (float)(0.5 * (double)(hi[d] + lo[d]));
            }
            float bsss = 0.0f;                                                           // :637-639 (std::pow(float,int) is double)
            for (int d = 0; d < PD; ++d) bsss = (float)((double)bsss + (double)t.ns[d][node] * (double)t.ns[d][node]);
            t.nr[node] = (float)(0.5 * (double)std::sqrt(bsss));
            if (t.num[node] <= bs) continue;                                             // :644 leaf
            int axis = 0; float axsz = -1.0f;                                            // :652-659 first strict max
            for (int d = 0; d < PD; ++d) if (t.ns[d][node] > axsz) { axsz = t.ns[d][node]; axis = d; }
            const size_t pm = pf + bs * ((size_t)1 << log_2((uint32_t)((t.num[node] - 1) / bs)));   // :663
            partial_select(p.x[axis], lidx, lo[axis], hi[axis], pf, pm, pl, S.bstats);   // :673
            for (int d = 0; d < PD; ++d) if (d != axis) gather_f(p.x[d], ftemp, lidx, pf, pl);      // :682-684
            if (p.are_sources) for (int d = 0; d < SD; ++d) gather_f(p.s[d], ftemp, lidx, pf, pl);  // :692
            gather_f(p.r, ftemp, lidx, pf, pl);                                          // :693
            gather_u(p.gidx, itemp, lidx, pf, pl);                                       // :694
            t.ioffset[2*node] = pf;   t.num[2*node] = pm - pf;                           // :702-704
            t.ioffset[2*node+1] = pm; t.num[2*node+1] = pl - pm;
        }
    }

    // finishTree (barneshut.hpp:717-807), deepest level first
    for (int lev = t.levels - 1; lev >= 0; --lev) {
        for (size_t node = (size_t)1 << lev; node < ((size_t)2 << lev); ++node) {
            if (t.num[node] == 0) continue;
            if (t.num[node] > bs) {                                                      // :721-746
                const size_t c1 = 2*node, c2 = 2*node+1;
                const float oonp = 1.0f / (float)(t.num[c1] + t.num[c2]);
                for (int d = 0; d < PD; ++d)
                    t.x[d][node] = oonp * ((float)t.num[c1] * t.x[d][c1] + (float)t.num[c2] * t.x[d][c2]);
                for (int d = 0; d < SD; ++d) t.s[d][node] = t.s[d][c1] + t.s[d][c2];
                t.pr[node] = oonp * ((float)t.num[c1] * t.pr[c1] + (float)t.num[c2] * t.pr[c2]);
            } else {                                                                     // :753-806
                const size_t pf = t.ioffset[node], pl = pf + t.num[node];
                std::vector<float> w(pl - pf);
                if (p.are_sources) {
                    if (SD == 1) for (size_t i = pf; i < pl; ++i) w[i-pf] = std::fabs(p.s[0][i]);
                    else {
                        std::fill(w.begin(), w.end(), 0.0f);
                        for (int d = 0; d < SD; ++d) for (size_t i = pf; i < pl; ++i)
                            w[i-pf] = (float)((double)w[i-pf] + (double)p.s[d][i] * (double)p.s[d][i]);
                        for (auto& q : w) q = std::sqrt(q);
                    }
                } else std::fill(w.begin(), w.end(), 1.0f);
                double wsum = 0.0; for (float q : w) wsum = wsum + q;
                const float ooass = (float)(1.0 / (1.e-20 + wsum));                      // :786
                for (int d = 0; d < PD; ++d) {                                           // :790 float product, double sum
                    double acc = 0.0;
                    for (size_t i = pf; i < pl; ++i) acc = acc + (double)(p.x[d][i] * w[i-pf]);
                    t.x[d][node] = (float)((double)ooass * acc);
                }
                if (p.are_sources) for (int d = 0; d < SD; ++d) {                        // :795-797
                    double acc = 0.0; for (size_t i = pf; i < pl; ++i) acc = acc + p.s[d][i];
                    t.s[d][node] = (float)acc;
                }
                double racc = 0.0; for (size_t i = pf; i < pl; ++i) racc = racc + p.r[i];
                const float radsum = (float)racc;                                        // :800-801
                t.pr[node] = radsum / (float)t.num[node];
            }
        }
    }
    if (p.are_sources) p.gidx.clear();                                                   // :853
}

// ---------------------------------------------------------------------------------------------
// a-1c: refineLeaf (barneshut.hpp:860-895). The sort is libstdc++ 13.3's std::sort on an index
// array (barneshut.hpp:411), which is UNSTABLE: equal keys come out in introsort's order. Restated
// from the published algorithm (bits/stl_algo.h of GCC 13.3.0: __introsort_loop threshold 16,
// median-of-three to first, unguarded Hoare partition, then one insertion sort) so ties match.
// ---------------------------------------------------------------------------------------------
struct IdxLess { const float* v; bool operator()(uint64_t a, uint64_t b) const { return v[a] < v[b]; } };

void move_median_to_first(uint64_t* r, uint64_t* a, uint64_t* b, uint64_t* c, IdxLess lt) {
    if (lt(*a, *b)) {
        if (lt(*b, *c)) std::swap(*r, *b); else if (lt(*a, *c)) std::swap(*r, *c); else std::swap(*r, *a);
    } else if (lt(*a, *c)) std::swap(*r, *a);
    else if (lt(*b, *c)) std::swap(*r, *c);
    else std::swap(*r, *b);
}
uint64_t* unguarded_partition(uint64_t* first, uint64_t* last, uint64_t* pivot, IdxLess lt) {
    while (true) {
        while (lt(*first, *pivot)) ++first;
        --last;
        while (lt(*pivot, *last)) --last;
        if (!(first < last)) return first;
        std::swap(*first, *last);
        ++first;
    }
}
void heap_sort_fallback(uint64_t* first, uint64_t* last, IdxLess lt) {
    // depth limit hit (2*floor(log2 n) bad splits): libstdc++ switches to heapsort (__partial_sort).
    // std::make_heap/sort_heap ARE libstdc++ here, so calling them reproduces the reference exactly.
    std::make_heap(first, last, lt); std::sort_heap(first, last, lt);
}
void introsort_loop(uint64_t* first, uint64_t* last, int depth, IdxLess lt) {
    while (last - first > 16) {
        if (depth == 0) { heap_sort_fallback(first, last, lt); return; }
        --depth;
        uint64_t* mid = first + (last - first) / 2;
        move_median_to_first(first, first + 1, mid, last - 1, lt);
        uint64_t* cut = unguarded_partition(first + 1, last, first, lt);
        introsort_loop(cut, last, depth, lt);
        last = cut;
    }
}
void unguarded_linear_insert(uint64_t* last, IdxLess lt) {
    uint64_t val = *last; uint64_t* next = last - 1;
    while (lt(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
void insertion_sort(uint64_t* first, uint64_t* last, IdxLess lt) {
    if (first == last) return;
    for (uint64_t* i = first + 1; i != last; ++i) {
        if (lt(*i, *first)) { uint64_t val = *i; std::memmove(first + 1, first, (i - first) * sizeof(uint64_t)); *first = val; }
        else unguarded_linear_insert(i, lt);
    }
}
void libstdcxx_sort(uint64_t* first, uint64_t* last, IdxLess lt) {
    if (first == last) return;
    introsort_loop(first, last, 2 * (int)log_2((uint32_t)(last - first)), lt);
    if (last - first > 16) {
        insertion_sort(first, first + 16, lt);
        for (uint64_t* i = first + 16; i != last; ++i) unguarded_linear_insert(i, lt);
    } else insertion_sort(first, last, lt);
}

void refine_leaf(Session& S, PParts& p, std::vector<uint64_t>& lidx, std::vector<uint64_t>& itemp,
                 std::vector<float>& ftemp, size_t pf, size_t pl) {
    if (pl - pf < 3) return;                                                             // :864
    float bsz[3]; int axis = 0;
    for (int d = 0; d < p.PD; ++d) {
        float lo = p.x[d][pf], hi = lo;
        for (size_t i = pf; i < pl; ++i) { if (p.x[d][i] < lo) lo = p.x[d][i]; if (p.x[d][i] > hi) hi = p.x[d][i]; }
        bsz[d] = hi - lo;
    }
    for (int d = 1; d < p.PD; ++d) if (bsz[axis] < bsz[d]) axis = d;                     // :878 std::max_element = first max
    for (size_t i = pf; i < pl; ++i) lidx[i] = i;                                        // :408
    libstdcxx_sort(lidx.data() + pf, lidx.data() + pl, IdxLess{ p.x[axis].data() });     // :411
    for (size_t i = pf + 1; i < pl; ++i) if (p.x[axis][lidx[i]] == p.x[axis][lidx[i-1]]) { S.tie_sorts++; break; }
    for (int d = 0; d < p.PD; ++d) gather_f(p.x[d], ftemp, lidx, pf, pl);                // :884-887
    if (p.are_sources) for (int d = 0; d < p.SD; ++d) gather_f(p.s[d], ftemp, lidx, pf, pl);
    gather_f(p.r, ftemp, lidx, pf, pl);
    gather_u(p.gidx, itemp, lidx, pf, pl);
    const size_t pm = pf + ((size_t)1 << log_2((uint32_t)(pl - pf - 1)));                // :890
    refine_leaf(S, p, lidx, itemp, ftemp, pf, pm);
    refine_leaf(S, p, lidx, itemp, ftemp, pm, pl);
}

void refine_tree(Session& S, PParts& p, const PTree& t) {                                // :901-936
    std::vector<uint64_t> lidx(p.n), itemp(p.n);
    std::vector<float> ftemp(p.n);
    const bool had_gidx = p.gidx.size() == p.n;
    if (p.are_sources || !had_gidx) { p.gidx.resize(p.n); for (size_t i = 0; i < p.n; ++i) p.gidx[i] = i; }
    S.tie_sorts = 0;
    for (size_t node = 1; node < (size_t)t.numnodes; ++node)
        if (t.num[node] > 0 && t.num[node] <= p.blockSize)
            refine_leaf(S, p, lidx, itemp, ftemp, t.ioffset[node], t.ioffset[node] + t.num[node]);
    if (p.are_sources) p.gidx.clear();
}

// ---------------------------------------------------------------------------------------------
// a-2 / a-6: barycentric Lagrange upward and downward (BarycentricLagrange.hpp)
// ---------------------------------------------------------------------------------------------
struct Cheb { float sk[21], wk[21]; };
Cheb make_cheb(int order) {                                                              // :28-48
    Cheb c;
    for (int k = 0; k <= order; ++k) c.sk[k] = (float)(-std::cos(k * M_PI / order));
    c.wk[0] = 0.5f;
    for (int k = 1; k < order; ++k) c.wk[k] = (k % 2) ? -1.0f : 1.0f;
    c.wk[order] = 0.5f * ((order % 2) ? -1.0f : 1.0f);
    return c;
}
inline size_t ipow_sz(size_t b, int e) { size_t r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

// the per-point 1D weights (shared by :190-225 and :102-137): amat[d][k], returns 1/prod(sum)
float bary_weights(int PD, size_t ncp, const Cheb& c, const float* px, const float* lsk, float amat[3][21]) {
    float denom = 1.0f;
    for (int d = 0; d < PD; ++d) {
        int flag = -1; float sum = 0.0f;
        for (size_t k = 0; k < ncp; ++k) {
            amat[d][k] = 0.0f;
            const float dist = px[d] - lsk[d*ncp + k];
            if ((double)std::fabs(dist) < 1.e-10) flag = (int)k;                          // CLOSE_THRESH :16
            else { amat[d][k] = c.wk[k] / dist; sum += amat[d][k]; }
        }
        if (flag > -1) { sum = 1.0f; for (size_t k = 0; k < ncp; ++k) amat[d][k] = 0.0f; amat[d][flag] = 1.0f; }
        denom *= sum;
    }
    return 1.0f / denom;
}

void bary_upward_node(Session& S, const PParts& p, PParts& ep, PTree& t, const Cheb& c, size_t node) {
    const int PD = p.PD, SD = p.SD, order = S.order;
    const size_t ncp = order + 1, numEqps = ipow_sz(ncp, PD), ebs = ep.blockSize;
    t.epoffset[node] = node * ebs; t.epnum[node] = 0;                                    // :289-290
    const size_t e0 = t.epoffset[node];
    float lsk[3*21];
    for (int d = 0; d < PD; ++d) for (size_t k = 0; k < ncp; ++k)
        lsk[d*ncp + k] = t.nc[d][node] + 0.5f * c.sk[k] * t.ns[d][node];                 // :305
    for (int d = 0; d < PD; ++d) {                                                       // :325-332
        const size_t div = ipow_sz(ncp, d);
        for (size_t i = 0; i < numEqps; ++i) ep.x[d][e0+i] = t.nc[d][node] + 0.5f * c.sk[(i/div) % ncp] * t.ns[d][node];
    }
    for (size_t i = e0 + numEqps; i < e0 + ebs; ++i) for (int d = 0; d < PD; ++d) ep.x[d][i] = t.nc[d][node];   // :335-337
    if (ep.are_sources) for (size_t i = e0; i < e0 + ebs; ++i) for (int d = 0; d < SD; ++d) ep.s[d][i] = 0.0f;  // :343-347
    for (size_t i = e0; i < e0 + ebs; ++i) ep.r[i] = p.r[t.ioffset[node]];               // :353
    for (size_t child = 2*node; child < 2*node + 2; ++child) {                           // :360-406
        const bool leaf = !(t.num[child] > p.blockSize);
        const PParts& sp = leaf ? p : ep;
        const size_t is = leaf ? t.ioffset[child] : t.epoffset[child];
        const size_t ie = is + (leaf ? t.num[child] : t.epnum[child]);
        if (p.are_sources && ep.are_sources) {
            for (size_t ip = is; ip < ie; ++ip) {                                        // :190-247
                float px[3] = { sp.x[0][ip], sp.x[1][ip], PD > 2 ? sp.x[2][ip] : 0.f };
                float amat[3][21];
                const float denom = bary_weights(PD, ncp, c, px, lsk, amat);
                for (size_t i = 0; i < numEqps; ++i) {
                    float wgt = denom; size_t q = i;
                    for (int d = 0; d < PD; ++d) { wgt *= amat[d][q % ncp]; q /= ncp; }
                    for (int d = 0; d < SD; ++d) ep.s[d][e0+i] += wgt * sp.s[d][ip];
                }
            }
        }
        t.epnum[node] = numEqps;
    }
}

// calcEquivalents barneshut.hpp:946-1061 - the drivers' default when -o is omitted (order = -1): every non-leaf node
// gets ceil(cnt/2) equivalents per child, each the strength-weighted merge of two consecutive points of the child
// (its real particles if it is a leaf, its own equivalents otherwise), an odd last point is passed up unchanged.
// The recursion only orders children before parents, so it is restated level by level, deepest first.
void legacy_merge(const PParts& sp, size_t i1, size_t i2, PParts& ep, size_t iep) {         // :985-1001 == :1031-1043
    const int PD = sp.PD, SD = sp.SD;
    float str1, str2;
    if (SD == 1) {
        str1 = std::max(1.e-20f, std::abs(sp.s[0][i1]));
        str2 = std::max(1.e-20f, std::abs(sp.s[0][i2]));
    } else {
        str1 = 0.0f; for (int d = 0; d < SD; ++d) str1 += std::pow(sp.s[d][i1], 2);
        str1 = std::max(1.e-20f, std::sqrt(str1));
        str2 = 0.0f; for (int d = 0; d < SD; ++d) str2 += std::pow(sp.s[d][i2], 2);
        str2 = std::max(1.e-20f, std::sqrt(str2));
    }
    const float pairm = 1.0 / (str1 + str2);
    for (int d = 0; d < PD; ++d) ep.x[d][iep] = (sp.x[d][i1]*str1 + sp.x[d][i2]*str2) * pairm;
    ep.r[iep] = std::sqrt((std::pow(sp.r[i1],2)*str1 + std::pow(sp.r[i2],2)*str2) * pairm);
    for (int d = 0; d < SD; ++d) ep.s[d][iep] = sp.s[d][i1] + sp.s[d][i2];
}
void legacy_equivalents(Session& S, PParts& p, PParts& ep, PTree& t) {
    (void)S;
    ep.resize((size_t)(t.numnodes / 2) * ep.blockSize);                                  // ongrav3d.cpp:645
    if (!p.are_sources || !ep.are_sources) return;                                       // :953
    const size_t ebs = ep.blockSize;
    for (int lev = t.levels - 1; lev >= 0; --lev)
        for (size_t node = (size_t)1 << lev; node < ((size_t)2 << lev); ++node) {
            if (!(t.num[node] > p.blockSize)) continue;
            t.epoffset[node] = node * ebs; t.epnum[node] = 0;                            // :955-956
            for (size_t child = 2*node; child < 2*node + 2; ++child) {
                const bool leaf = !(t.num[child] > p.blockSize);                         // :963
                const PParts& sp = leaf ? p : ep;
                const size_t first = leaf ? t.ioffset[child] : t.epoffset[child];
                const size_t cnt = leaf ? t.num[child] : t.epnum[child];
                const size_t numEqps = (cnt + 1) / 2, istart = (ebs / 2) * child;        // :974-975, :1020-1021
                for (size_t j = 0; j < numEqps; ++j) {
                    const size_t i1 = first + 2*j, iep = istart + j;
                    if (2*j + 1 < cnt) legacy_merge(sp, i1, i1 + 1, ep, iep);
                    else {                                                               // :1003-1008, :1049-1054
                        for (int d = 0; d < p.PD; ++d) ep.x[d][iep] = sp.x[d][i1];
                        for (int d = 0; d < p.SD; ++d) ep.s[d][iep] = sp.s[d][i1];
                        ep.r[iep] = sp.r[i1];
                    }
                }
                t.epnum[node] += numEqps;
            }
        }
}

void bary_upward(Session& S, PParts& p, PParts& ep, PTree& t) {                          // :255-417, post-order == deepest level first
    ep.resize((size_t)(t.numnodes / 2) * ep.blockSize);                                  // ongrav3d.cpp:645
    const Cheb c = make_cheb(S.order);
    for (int lev = t.levels - 1; lev >= 0; --lev)
        for (size_t node = (size_t)1 << lev; node < ((size_t)2 << lev); ++node)
            if (t.num[node] > p.blockSize) bary_upward_node(S, p, ep, t, c, node);
}

// calcBarycentricDownward (BarycentricLagrange.hpp:62-166): parent eq values -> points [istart,istop) of tp
void bary_downward(Session& S, const PParts& sp, PParts& tp, size_t istart, size_t istop, size_t iepstart) {
    const int PD = sp.PD, OD = sp.OD, order = S.order;
    const size_t ncp = order + 1, numEqps = ipow_sz(ncp, PD);
    const Cheb c = make_cheb(order);
    float lsk[3*21];
    { size_t stride = 1;                                                                 // :78-89 read back from the eq points
      for (int d = 0; d < PD; ++d) { for (size_t k = 0; k < ncp; ++k) lsk[d*ncp+k] = sp.x[d][iepstart + stride*k]; stride *= ncp; } }
    for (size_t ip = istart; ip < istop; ++ip) {
        float px[3] = { tp.x[0][ip], tp.x[1][ip], PD > 2 ? tp.x[2][ip] : 0.f };
        float amat[3][21];
        const float denom = bary_weights(PD, ncp, c, px, lsk, amat);
        for (size_t i = 0; i < numEqps; ++i) {                                           // :140-156
            float wgt = denom; size_t q = i;
            for (int d = 0; d < PD; ++d) { wgt *= amat[d][q % ncp]; q /= ncp; }
            for (int d = 0; d < OD; ++d) tp.u[d][ip] += wgt * sp.u[d][iepstart + i];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a-4: treecodes (barneshut.hpp:65-337); recursion order = accumulation order
// ---------------------------------------------------------------------------------------------
void tc1_block(Session& S, size_t sn, size_t ip, float theta) {                          // :65-102
    const PParts& sp = S.srcs; const PTree& st = S.stree; PParts& tp = S.targs;
    if (st.num[sn] <= sp.blockSize) { ppinter(S.physics, sp, st.ioffset[sn], st.ioffset[sn]+st.num[sn], tp, ip, ip+1); S.stats[0]++; return; }
    float dist = 0.0f;
    for (int d = 0; d < S.PD; ++d) {
        const float v = std::max(0.0f, std::fabs(st.x[d][sn] - tp.x[d][ip]) - 0.5f * st.ns[d][sn]);
        dist = (float)((double)dist + (double)v * (double)v);
    }
    dist = std::sqrt(dist);
    if ((double)dist / (2.0 * (double)st.nr[sn]) > (double)theta) { tpinter(S.physics, st, sn, tp, ip); S.stats[1]++; }
    else { tc1_block(S, 2*sn, ip, theta); tc1_block(S, 2*sn+1, ip, theta); }
}
void tc2_block(Session& S, size_t sn, size_t ip, float theta) {                          // :137-184
    const PParts& sp = S.srcs; const PTree& st = S.stree; PParts& tp = S.targs;
    if (st.num[sn] <= sp.blockSize) { ppinter(S.physics, sp, st.ioffset[sn], st.ioffset[sn]+st.num[sn], tp, ip, ip+1); S.stats[0]++; return; }
    float dist = 0.0f;
    for (int d = 0; d < S.PD; ++d) { const float v = st.nc[d][sn] - tp.x[d][ip]; dist = (float)((double)dist + (double)v * (double)v); }
    dist = std::sqrt(dist);
    if ((double)dist / (2.0 * (double)st.nr[sn]) > (double)theta) {
        ppinter(S.physics, S.eqsrcs, st.epoffset[sn], st.epoffset[sn]+st.epnum[sn], tp, ip, ip+1); S.stats[1]++;
    } else { tc2_block(S, 2*sn, ip, theta); tc2_block(S, 2*sn+1, ip, theta); }
}
void tc3_block(Session& S, size_t sn, size_t tn, float theta) {                          // :228-294
    const PParts& sp = S.srcs; const PTree& st = S.stree; const PTree& tt = S.ttree; PParts& tp = S.targs;
    if (st.num[sn] <= sp.blockSize) {
        ppinter(S.physics, sp, st.ioffset[sn], st.ioffset[sn]+st.num[sn], tp, tt.ioffset[tn], tt.ioffset[tn]+tt.num[tn]); S.stats[0]++; return;
    }
    float dist = 0.0f;
    for (int d = 0; d < S.PD; ++d) { const float v = st.nc[d][sn] - tt.nc[d][tn]; dist = (float)((double)dist + (double)v * (double)v); }
    dist = std::sqrt(dist);
    const float testrad = std::max(st.nr[sn], tt.nr[tn]) + 0.25f * std::min(st.nr[sn], tt.nr[tn]);   // :280
    if (dist / (2.0f * testrad) > theta) {                                               // :283
        ppinter(S.physics, S.eqsrcs, st.epoffset[sn], st.epoffset[sn]+st.epnum[sn], tp, tt.ioffset[tn], tt.ioffset[tn]+tt.num[tn]); S.stats[1]++;
    } else { tc3_block(S, 2*sn, tn, theta); tc3_block(S, 2*sn+1, tn, theta); }
}

// ---------------------------------------------------------------------------------------------
// a-5: dual-tree traversal, level-synchronous (SURVEY App. B == ongrav3d.cpp:206-452)
// ---------------------------------------------------------------------------------------------
void fastsumm(Session& S, float theta) {
    const PParts& srcs = S.srcs; const PParts& eqsrcs = S.eqsrcs; const PTree& st = S.stree;
    PParts& targs = S.targs; PParts& eqtargs = S.eqtargs; const PTree& tt = S.ttree;
    const size_t bs = targs.blockSize;
    std::vector<std::vector<uint64_t>> L(tt.numnodes);
    L[1] = { 1 };
    for (int lev = 0; lev < tt.levels; ++lev) {
        for (size_t T = (size_t)1 << lev; T < ((size_t)2 << lev); ++T) {
            if (tt.num[T] < 1) continue;                                                 // :221
            const bool tleaf = tt.num[T] <= bs;
            PParts& acc = tleaf ? targs : eqtargs;
            const size_t a0 = tleaf ? tt.ioffset[T] : tt.epoffset[T];
            const size_t an = tleaf ? tt.num[T] : tt.epnum[T];
            for (int d = 0; d < S.OD; ++d) std::fill_n(&acc.u[d][a0], an, 0.0f);         // :232,:270
            if (tleaf) S.stats[6]++;
            if (T > 1) {                                                                 // :257,:296
                bary_downward(S, eqtargs, acc, a0, a0 + an, tt.epoffset[T/2]);
                if (tleaf) S.stats[7]++; else S.stats[8]++;
            }
            std::vector<uint64_t> work = L[T], C;
            for (size_t i = 0; i < work.size(); ++i) {                                   // :315 (list grows while iterating)
                const size_t sn = work[i];
                if (st.num[sn] < 1) continue;                                            // :319
                const bool sleaf = st.num[sn] <= srcs.blockSize;
                if (sleaf && tleaf) {                                                    // :326-334
                    ppinter(S.physics, srcs, st.ioffset[sn], st.ioffset[sn]+st.num[sn], targs, a0, a0+an); S.stats[2]++; continue;
                }
                float dist = 0.0f;
                for (int d = 0; d < S.PD; ++d) { const float v = st.x[d][sn] - tt.x[d][T]; dist = (float)((double)dist + (double)v * (double)v); }
                dist = std::sqrt(dist);
                const float diag = st.nr[sn] + tt.nr[T];                                 // :340
                if (dist / diag > theta) {                                               // :344
                    if (sleaf)      { ppinter(S.physics, srcs,   st.ioffset[sn],  st.ioffset[sn]+st.num[sn],    eqtargs, a0, a0+an); S.stats[4]++; }
                    else if (tleaf) { ppinter(S.physics, eqsrcs, st.epoffset[sn], st.epoffset[sn]+st.epnum[sn], targs,   a0, a0+an); S.stats[3]++; }
                    else            { ppinter(S.physics, eqsrcs, st.epoffset[sn], st.epoffset[sn]+st.epnum[sn], eqtargs, a0, a0+an); S.stats[5]++; }
                } else if (tt.nr[T] > st.nr[sn]) {                                       // :367-382
                    if (tleaf) { work.push_back(2*sn); work.push_back(2*sn+1); } else C.push_back(sn);
                } else {                                                                 // :384-399
                    if (sleaf) C.push_back(sn); else { work.push_back(2*sn); work.push_back(2*sn+1); }
                }
            }
            if (!tleaf) { L[2*T] = C; L[2*T+1] = C; }                                    // :418-422
            std::vector<uint64_t>().swap(L[T]);
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C API
// ---------------------------------------------------------------------------------------------
extern "C" {

void* oport_create(int physics, uint64_t nsrc, uint64_t ntarg, int blockSize, int eqBlockSize, int order) {
    static const int PDs[5] = {3,3,3,2,2}, SDs[5] = {1,3,3,1,1}, ODs[5] = {3,3,12,2,2}, FL[5] = {19,28,64,13,15};
    Session* S = new Session();
    S->physics = physics; S->PD = PDs[physics]; S->SD = SDs[physics]; S->OD = ODs[physics]; S->flops = FL[physics];
    S->order = order; S->has_tr = physics == VORT2DTR; S->has_fastsumm = physics != VORTGRAD3D;
    PParts* ps[4] = { &S->srcs, &S->targs, &S->eqsrcs, &S->eqtargs };
    for (int i = 0; i < 4; ++i) { ps[i]->PD = S->PD; ps[i]->SD = S->SD; ps[i]->OD = S->OD; ps[i]->are_sources = (i % 2 == 0); }
    S->srcs.blockSize = S->targs.blockSize = blockSize; S->eqsrcs.blockSize = S->eqtargs.blockSize = eqBlockSize;
    S->srcs.resize(nsrc); S->targs.resize(ntarg);
    return S;
}
void oport_destroy(void* h) { delete (Session*)h; }
void oport_dims(void* h, int* pd, int* sd, int* od, int* hf) { Session* S = (Session*)h; *pd = S->PD; *sd = S->SD; *od = S->OD; *hf = S->has_fastsumm; }

// Parts.hpp:99-109 (engine passed by value) + wave_strengths Parts.hpp:169-176
void oport_init_driver(void* h, int strength_mode) {
    Session* S = (Session*)h;
    for (int pass = 0; pass < 2; ++pass) {
        PParts& p = pass == 0 ? S->srcs : S->targs;
        std::mt19937 eng(12345);
        std::uniform_real_distribution<float> dist(-1.0, 1.0);
        for (int d = 0; d < p.PD; ++d) for (auto& v : p.x[d]) v = dist(eng);
        if (p.are_sources) { const float factor = (float)(1.0 / (float)p.n); for (int d = 0; d < p.SD; ++d) for (auto& v : p.s[d]) v = dist(eng) * factor; }
        const float rad = (float)std::pow((float)p.n, -1.0 / (float)p.PD);
        for (auto& v : p.r) v = rad;
    }
    if (strength_mode == 1) {
        PParts& p = S->srcs; const float factor = (float)(1.0 / (float)p.n);
        for (size_t i = 0; i < p.n; ++i) for (int d = 0; d < p.SD; ++d) p.s[d][i] = (float)((double)factor * std::cos((d + 0.7) * 10.0 * (double)p.x[d][i]));
    }
}
void oport_set_sources(void* h, const float* x, const float* r, const float* s) {
    Session* S = (Session*)h; const size_t n = S->srcs.n;
    for (int d = 0; d < S->PD; ++d) memcpy(S->srcs.x[d].data(), x + d*n, n*4);
    memcpy(S->srcs.r.data(), r, n*4);
    for (int d = 0; d < S->SD; ++d) memcpy(S->srcs.s[d].data(), s + d*n, n*4);
}
void oport_set_targets(void* h, const float* x, const float* r) {
    Session* S = (Session*)h; const size_t n = S->targs.n;
    for (int d = 0; d < S->PD; ++d) memcpy(S->targs.x[d].data(), x + d*n, n*4);
    memcpy(S->targs.r.data(), r, n*4);
}
static PParts& parts_of(Session* S, int which) { return which == 0 ? S->srcs : which == 1 ? S->targs : which == 2 ? S->eqsrcs : S->eqtargs; }
static PTree& tree_of(Session* S, int which) { return which == 0 ? S->stree : S->ttree; }

void oport_make_tree(void* h, int which) { Session* S = (Session*)h; make_tree(*S, parts_of(S, which), tree_of(S, which)); }
void oport_refine(void* h, int which) { Session* S = (Session*)h; refine_tree(*S, parts_of(S, which), tree_of(S, which)); }
void oport_upward(void* h, int which) {
    Session* S = (Session*)h;
    if (S->order < 0) legacy_equivalents(*S, parts_of(S, which), parts_of(S, which+2), tree_of(S, which));
    else bary_upward(*S, parts_of(S, which), parts_of(S, which+2), tree_of(S, which));
}
void oport_zero_vels(void* h) { Session* S = (Session*)h; for (int d = 0; d < S->OD; ++d) std::fill(S->targs.u[d].begin(), S->targs.u[d].end(), 0.0f); }

float oport_naive(void* h, uint64_t tskip) {                                             // barneshut.hpp:46-53
    Session* S = (Session*)h;
    const long n = (long)S->targs.n;
    #pragma omp parallel for schedule(dynamic,16)
    for (long i = 0; i < n; i += (long)tskip) ppinter(S->physics, S->srcs, 0, S->srcs.n, S->targs, i, i+1);
    return (float)(S->targs.n / tskip) * (float)S->srcs.n * (float)S->flops;
}
float oport_treecode1(void* h, float theta) {                                            // :107-132
    Session* S = (Session*)h; memset(S->stats, 0, sizeof(S->stats));
    for (size_t i = 0; i < S->targs.n; ++i) tc1_block(*S, 1, i, theta);
    return (float)S->flops * ((float)S->stats[1] + (float)S->stats[0] * (float)S->srcs.blockSize);
}
float oport_treecode2(void* h, float theta) {                                            // :189-222
    Session* S = (Session*)h; memset(S->stats, 0, sizeof(S->stats));
    for (size_t i = 0; i < S->targs.n; ++i) tc2_block(*S, 1, i, theta);
    return (float)S->flops * ((float)S->stats[0] * (float)S->srcs.blockSize + (float)S->stats[1] * (float)S->stree.epnum[1]);
}
float oport_treecode3(void* h, float theta) {                                            // :299-337
    Session* S = (Session*)h; memset(S->stats, 0, sizeof(S->stats));
    for (size_t ib = 0; ib < (size_t)S->ttree.numnodes; ++ib)
        if (S->ttree.num[ib] <= S->targs.blockSize && S->ttree.num[ib] > 0) tc3_block(*S, 1, ib, theta);
    return (float)S->flops * (float)S->targs.blockSize *
           ((float)S->stats[0] * (float)S->srcs.blockSize + (float)S->stats[1] * (float)S->stree.epnum[1]);
}
int oport_fastsumm(void* h, float theta, int parallel) {
    (void)parallel;
    Session* S = (Session*)h; if (!S->has_fastsumm) return -1;
    memset(S->stats, 0, sizeof(S->stats));
    fastsumm(*S, theta); return 0;
}
uint64_t oport_count(void* h, int which) { return parts_of((Session*)h, which).n; }
void oport_get_parts(void* h, int which, float* x, float* r, float* s, float* u, uint64_t* gidx) {
    Session* S = (Session*)h; PParts& p = parts_of(S, which); const size_t n = p.n;
    if (x) for (int d = 0; d < S->PD; ++d) memcpy(x + d*n, p.x[d].data(), n*4);
    if (r) memcpy(r, p.r.data(), n*4);
    if (s && p.are_sources) for (int d = 0; d < S->SD; ++d) memcpy(s + d*n, p.s[d].data(), n*4);
    if (u && !p.are_sources) for (int d = 0; d < S->OD; ++d) memcpy(u + d*n, p.u[d].data(), n*4);
    if (gidx && p.gidx.size() == n) memcpy(gidx, p.gidx.data(), n*8);
}
void oport_tree_shape(void* h, int which, int* levels, int* numnodes) { PTree& t = tree_of((Session*)h, which); *levels = t.levels; *numnodes = t.numnodes; }
void oport_get_tree(void* h, int which, float* x, float* nc, float* ns, float* nr, float* pr, float* s,
                    uint64_t* ioffset, uint64_t* num, uint64_t* epoffset, uint64_t* epnum) {
    Session* S = (Session*)h; PTree& t = tree_of(S, which); const size_t n = t.numnodes;
    if (x)  for (int d = 0; d < S->PD; ++d) memcpy(x  + d*n, t.x[d].data(),  n*4);
    if (nc) for (int d = 0; d < S->PD; ++d) memcpy(nc + d*n, t.nc[d].data(), n*4);
    if (ns) for (int d = 0; d < S->PD; ++d) memcpy(ns + d*n, t.ns[d].data(), n*4);
    if (nr) memcpy(nr, t.nr.data(), n*4);
    if (pr) memcpy(pr, t.pr.data(), n*4);
    if (s) for (int d = 0; d < S->SD; ++d) memcpy(s + d*n, t.s[d].data(), n*4);
    if (ioffset) memcpy(ioffset, t.ioffset.data(), n*8);
    if (num) memcpy(num, t.num.data(), n*8);
    if (epoffset) memcpy(epoffset, t.epoffset.data(), n*8);
    if (epnum) memcpy(epnum, t.epnum.data(), n*8);
}
void oport_get_stats(void* h, uint64_t out[9]) { memcpy(out, ((Session*)h)->stats, 9*8); }
void oport_get_build_stats(void* h, uint64_t out[4]) { memcpy(out, ((Session*)h)->bstats, 4*8); }
uint64_t oport_refine_tie_sorts(void* h) { return ((Session*)h)->tie_sorts; }
uint64_t oport_fnv1a64(const void* data, uint64_t nbytes) {
    const unsigned char* p = (const unsigned char*)data; uint64_t hsh = 1469598103934665603ULL;
    for (uint64_t i = 0; i < nbytes; ++i) { hsh ^= p[i]; hsh *= 1099511628211ULL; }
    return hsh;
}

}  // extern "C"
