/*
 * hooks_common.hpp - C entry points around the UNMODIFIED reference templates.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path may include, link or call this.
 *
 * Each hooks_<driver>.cpp first includes the reference driver translation unit where it lies
 * under /root/reference/src (with `main` renamed), which brings the driver's own
 * nbody_kernel/ppinter/tpinter (+ nbody_fastsumm where present) and barneshut.hpp /
 * BarycentricLagrange.hpp into scope; this header then wraps the same call sequence the
 * driver's main() performs (ongrav3d.cpp:600-908) behind a handle so that tests can pull out
 * every intermediate array. No reference source is copied: the reference is compiled in place.
 *
 * Macros the including TU must define: OREF_PD, OREF_SD, OREF_OD, OREF_HAS_FASTSUMM (0/1).
 */
#pragma once

#include <cstdint>
#include <cstring>
#include <vector>

// OREF_ACCUM: the accumulator type the templates are instantiated with. The drivers fix it with "#define ACCUM float"
// (ongrav3d.cpp:8); hooks_grav3d_a64.cpp asks for double here, which is the README's fp32-store / fp64-accumulate variant -
// every reference function is a template on <S,A,...>, so no reference source needs touching.
#ifndef OREF_ACCUM
#define OREF_ACCUM ACCUM
#endif
typedef Parts<STORE,OREF_ACCUM,OREF_PD,OREF_SD,OREF_OD> OrefParts;
typedef Tree<STORE,OREF_PD,OREF_SD> OrefTree;

struct OrefSession {
    OrefParts srcs, targs, eqsrcs, eqtargs;
    OrefTree stree, ttree;
    int32_t order;
    OrefSession(size_t ns, size_t nt, size_t bs, size_t ebs, int32_t ord)
        : srcs(ns, true, bs), targs(nt, false, bs), eqsrcs(0, true, ebs), eqtargs(0, false, ebs),
          stree(0), ttree(0), order(ord) {}
};

#define OREF_API extern "C" __attribute__((visibility("default")))

OREF_API void* oref_create(uint64_t nsrc, uint64_t ntarg, int blockSize, int eqBlockSize, int order) {
    return new OrefSession(nsrc, ntarg, blockSize, eqBlockSize, order);
}
OREF_API void oref_destroy(void* h) { delete (OrefSession*)h; }

OREF_API void oref_dims(int* pd, int* sd, int* od, int* has_fastsumm) {
    *pd = OREF_PD; *sd = OREF_SD; *od = OREF_OD; *has_fastsumm = OREF_HAS_FASTSUMM;
}

// same initialisation as the drivers' main(): the engine is passed BY VALUE to both calls
// (Parts.hpp:100), so sources and targets get identical coordinates (ongrav3d.cpp:574-594).
// strength_mode 0: keep random_in_cube strengths (ongrav3d, onvort2d); 1: wave_strengths (onvort3d, onvortgrad3d)
OREF_API void oref_init_driver(void* h, int strength_mode) {
    OrefSession* s = (OrefSession*)h;
    std::mt19937 mt_engine(12345);
    s->srcs.random_in_cube(mt_engine);
    if (strength_mode == 1) s->srcs.wave_strengths();
    s->targs.random_in_cube(mt_engine);
}

// x is [PD][n], s is [SD][n]
OREF_API void oref_set_sources(void* h, const float* x, const float* r, const float* str) {
    OrefSession* s = (OrefSession*)h;
    const size_t n = s->srcs.n;
    for (int d=0; d<OREF_PD; ++d) std::memcpy(s->srcs.x[d].data(), x+d*n, n*sizeof(float));
    std::memcpy(s->srcs.r.data(), r, n*sizeof(float));
    for (int d=0; d<OREF_SD; ++d) std::memcpy(s->srcs.s[d].data(), str+d*n, n*sizeof(float));
}
OREF_API void oref_set_targets(void* h, const float* x, const float* r) {
    OrefSession* s = (OrefSession*)h;
    const size_t n = s->targs.n;
    for (int d=0; d<OREF_PD; ++d) std::memcpy(s->targs.x[d].data(), x+d*n, n*sizeof(float));
    std::memcpy(s->targs.r.data(), r, n*sizeof(float));
}

static OrefParts& oref_parts(OrefSession* s, int which) {
    switch (which) { case 0: return s->srcs; case 1: return s->targs; case 2: return s->eqsrcs; default: return s->eqtargs; }
}
static OrefTree& oref_tree(OrefSession* s, int which) { return which == 0 ? s->stree : s->ttree; }

// which: 0 = sources, 1 = targets
OREF_API void oref_make_tree(void* h, int which) {
    OrefSession* s = (OrefSession*)h;
    (void) makeTree(oref_parts(s, which), oref_tree(s, which));
}

OREF_API void oref_refine(void* h, int which) {
    OrefSession* s = (OrefSession*)h;
    #pragma omp parallel
    #pragma omp single
    (void) refineTree(oref_parts(s, which), oref_tree(s, which), 1);
    #pragma omp taskwait
}

// barycentric upward pass, driver sequence ongrav3d.cpp:645,665 (sources) / :696,719 (targets)
OREF_API void oref_upward(void* h, int which) {
    OrefSession* s = (OrefSession*)h;
    OrefParts& p = oref_parts(s, which);
    OrefParts& ep = oref_parts(s, which+2);
    OrefTree& t = oref_tree(s, which);
    ep.resize((t.numnodes/2) * ep.blockSize);
    if (s->order < 0) {
        // the drivers' default (-o omitted): hierarchical pair-merge equivalents, ongrav3d.cpp:654-659
        (void) calcEquivalents(p, ep, t, 1);
        return;
    }
    #pragma omp parallel
    #pragma omp single
    (void) calcBarycentricLagrange(p, ep, t, s->order, 1);
    #pragma omp taskwait
}

OREF_API void oref_zero_vels(void* h) { ((OrefSession*)h)->targs.zero_vels(); }

OREF_API float oref_naive(void* h, uint64_t tskip) {
    OrefSession* s = (OrefSession*)h;
    return nbody_naive(s->srcs, s->targs, (size_t)tskip);
}
OREF_API float oref_treecode1(void* h, float theta) {
    OrefSession* s = (OrefSession*)h;
    return nbody_treecode1(s->srcs, s->stree, s->targs, theta);
}
OREF_API float oref_treecode2(void* h, float theta) {
    OrefSession* s = (OrefSession*)h;
    return nbody_treecode2(s->srcs, s->eqsrcs, s->stree, s->targs, theta);
}
OREF_API float oref_treecode3(void* h, float theta) {
    OrefSession* s = (OrefSession*)h;
    return nbody_treecode3(s->srcs, s->eqsrcs, s->stree, s->targs, s->ttree, theta);
}

// Dual-tree traversal. parallel=0 calls it OUTSIDE any parallel region: the omp tasks are then
// executed immediately by the encountering thread, which avoids the reference's known race
// (ongrav3d.cpp:416-436, README.md:200) and gives the race-free answer. parallel=1 is the
// driver's own invocation (ongrav3d.cpp:880-884), only good for timing.
OREF_API int oref_fastsumm(void* h, float theta, int parallel) {
#if OREF_HAS_FASTSUMM
    OrefSession* s = (OrefSession*)h;
    std::vector<size_t> source_boxes = {1};
    if (parallel) {
        #pragma omp parallel
        #pragma omp single
        (void) nbody_fastsumm(s->srcs, s->eqsrcs, s->stree, s->targs, s->eqtargs, s->ttree,
                              1, source_boxes, (STORE)theta, s->order);
        #pragma omp taskwait
    } else {
        (void) nbody_fastsumm(s->srcs, s->eqsrcs, s->stree, s->targs, s->eqtargs, s->ttree,
                              1, source_boxes, (STORE)theta, s->order);
    }
    return 0;
#else
    (void)h; (void)theta; (void)parallel;
    return -1;
#endif
}

OREF_API uint64_t oref_count(void* h, int which) { return oref_parts((OrefSession*)h, which).n; }

// copy out particle arrays; any pointer may be null. x:[PD][n] r:[n] str:[SD][n] u:[OD][n] gidx:[n]
OREF_API void oref_get_parts(void* h, int which, float* x, float* r, float* str, float* u, uint64_t* gidx) {
    OrefParts& p = oref_parts((OrefSession*)h, which);
    const size_t n = p.n;
    if (x) for (int d=0; d<OREF_PD; ++d) std::memcpy(x+d*n, p.x[d].data(), n*sizeof(float));
    if (r) std::memcpy(r, p.r.data(), n*sizeof(float));
    if (str && p.are_sources) for (int d=0; d<OREF_SD; ++d) std::memcpy(str+d*n, p.s[d].data(), n*sizeof(float));
    if (u && !p.are_sources) for (int d=0; d<OREF_OD; ++d) for (size_t i=0; i<n; ++i) u[d*n+i] = (float)p.u[d][i];
    if (gidx && p.gidx.size() == n) for (size_t i=0; i<n; ++i) gidx[i] = p.gidx[i];
}
// outputs in the accumulator's own precision (which: 1 targets, 3 equivalent targets); u:[OD][n] doubles
OREF_API void oref_get_u64(void* h, int which, double* u) {
    OrefParts& p = oref_parts((OrefSession*)h, which);
    const size_t n = p.n;
    for (int d=0; d<OREF_OD; ++d) for (size_t i=0; i<n; ++i) u[d*n+i] = (double)p.u[d][i];
}
OREF_API int oref_accum_bytes(void) { return (int)sizeof(OREF_ACCUM); }

OREF_API void oref_tree_shape(void* h, int which, int* levels, int* numnodes) {
    OrefTree& t = oref_tree((OrefSession*)h, which);
    *levels = t.levels; *numnodes = t.numnodes;
}
// x,nc,ns: [PD][numnodes]; nr,pr: [numnodes]; str: [SD][numnodes]; ioffset,num,epoffset,epnum: [numnodes]
OREF_API void oref_get_tree(void* h, int which, float* x, float* nc, float* ns, float* nr, float* pr, float* str,
                            uint64_t* ioffset, uint64_t* num, uint64_t* epoffset, uint64_t* epnum) {
    OrefTree& t = oref_tree((OrefSession*)h, which);
    const size_t n = t.numnodes;
    if (x)  for (int d=0; d<OREF_PD; ++d) std::memcpy(x+d*n,  t.x[d].data(),  n*sizeof(float));
    if (nc) for (int d=0; d<OREF_PD; ++d) std::memcpy(nc+d*n, t.nc[d].data(), n*sizeof(float));
    if (ns) for (int d=0; d<OREF_PD; ++d) std::memcpy(ns+d*n, t.ns[d].data(), n*sizeof(float));
    if (nr) std::memcpy(nr, t.nr.data(), n*sizeof(float));
    if (pr) std::memcpy(pr, t.pr.data(), n*sizeof(float));
    if (str) for (int d=0; d<OREF_SD; ++d) std::memcpy(str+d*n, t.s[d].data(), n*sizeof(float));
    if (ioffset)  for (size_t i=0; i<n; ++i) ioffset[i]  = t.ioffset[i];
    if (num)      for (size_t i=0; i<n; ++i) num[i]      = t.num[i];
    if (epoffset) for (size_t i=0; i<n; ++i) epoffset[i] = t.epoffset[i];
    if (epnum)    for (size_t i=0; i<n; ++i) epnum[i]    = t.epnum[i];
}
