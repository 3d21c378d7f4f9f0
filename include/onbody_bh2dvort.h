/*
 * onbody_bh2dvort.h - drop-in for the reference's libbh2dvort (CMakeLists.txt:147-149): the four Fortran-callable entry
 * points of src/interface2dvort.cpp and src/interface2dvorttr.cpp, backed by the CUDA library
 * (libbh2dvort_b200.so -> libonbody_b200.so).
 *
 * Conventions kept (interface2dvort.cpp:182-187, 296-305, 324-329; interface2dvorttr.cpp:177-183, 321-327): trailing underscore,
 * scalars by pointer, int counts, caller-owned host arrays, results ACCUMULATED (tu -=/+= ...) in the caller's original target
 * order, return value = flop estimate; solver parameters theta = 1.3, order = 4, block = 128, boxwise (:193-197).
 *
 * One defect of the reference is deliberately NOT reproduced: its bh2dvort target links two translation units that both
 * instantiate ppinter<float,float,2,1,2> with different bodies (with and without target radius); the linker keeps one and
 * run2dvort reports rms 0.12. Here each of the four entry points has its own kernel.
 * No CPU path exists: see onbody_bh3dvortgrads.h for the failure behaviour.
 */
#ifndef ONBODY_BH2DVORT_H
#define ONBODY_BH2DVORT_H
#ifdef __cplusplus
extern "C" {
#endif

/* replaces interface2dvort.cpp:182 */
float external_vel_solver_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                             const int* ntarg, const float* tx, const float* ty, float* tu, float* tv);
/* replaces interface2dvort.cpp:324 */
float external_vel_direct_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                             const int* ntarg, const float* tx, const float* ty, float* tu, float* tv);
/* replaces interface2dvorttr.cpp:177 */
float external_vel_solver_tr_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                                const int* ntarg, const float* tx, const float* ty, const float* tr, float* tu, float* tv);
/* replaces interface2dvorttr.cpp:321 */
float external_vel_direct_tr_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                                const int* ntarg, const float* tx, const float* ty, const float* tr, float* tu, float* tv);

#ifdef __cplusplus
}
#endif
#endif
