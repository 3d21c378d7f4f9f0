/*
 * oracle_port.h - C API of the CPU restatement ("port") of onbody's summation hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load this library; the product (onbody_b200/) never does.
 *
 * The function set mirrors oracle/ref/hooks_common.hpp (prefix oport_ instead of oref_) so that one
 * Python front-end drives both and they can be diffed array by array.
 *
 * physics ids: 0 grav3d (ongrav3d.cpp:44-58)  1 vort3d (onvort3d.cpp:44-59)
 *              2 vortgrad3d (onvortgrad3d.cpp:45-76)  3 vort2d (interface2dvort.cpp:39-50)
 *              4 vort2dtr (onvort2d.cpp:44-55, target radius in the core)
 */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

void*    oport_create(int physics, uint64_t nsrc, uint64_t ntarg, int blockSize, int eqBlockSize, int order);
void     oport_destroy(void* h);
void     oport_dims(void* h, int* pd, int* sd, int* od, int* has_fastsumm);
void     oport_init_driver(void* h, int strength_mode);
void     oport_set_sources(void* h, const float* x, const float* r, const float* s);
void     oport_set_targets(void* h, const float* x, const float* r);
void     oport_make_tree(void* h, int which);
void     oport_refine(void* h, int which);
void     oport_upward(void* h, int which);
void     oport_zero_vels(void* h);
float    oport_naive(void* h, uint64_t tskip);
float    oport_treecode1(void* h, float theta);
float    oport_treecode2(void* h, float theta);
float    oport_treecode3(void* h, float theta);
int      oport_fastsumm(void* h, float theta, int parallel);
uint64_t oport_count(void* h, int which);
void     oport_get_parts(void* h, int which, float* x, float* r, float* s, float* u, uint64_t* gidx);
void     oport_tree_shape(void* h, int which, int* levels, int* numnodes);
void     oport_get_tree(void* h, int which, float* x, float* nc, float* ns, float* nr, float* pr, float* s,
                        uint64_t* ioffset, uint64_t* num, uint64_t* epoffset, uint64_t* epnum);

/* extras the reference does not expose */
/* counters of the last treecode / dual-tree call: sltp sbtp | sltl sbtl sltb sbtb tlc lpc bpc */
void     oport_get_stats(void* h, uint64_t out[9]);
/* tree-build statistics of the last make_tree: selects, passes, stall exits, elements scanned */
void     oport_get_build_stats(void* h, uint64_t out[4]);
/* number of std::sort calls in the last refine that saw equal keys (where libstdc++'s tie order matters) */
uint64_t oport_refine_tie_sorts(void* h);
uint64_t oport_fnv1a64(const void* data, uint64_t nbytes);

#ifdef __cplusplus
}
#endif
