// iface2dvort.cpp - libbh2dvort_b200.so: the reference's four 2-D vortex entry points (src/interface2dvort.cpp:182-374,
// src/interface2dvorttr.cpp:177-373) on the B200. See include/onbody_bh2dvort.h. The variants with and without target
// radius use different kernels (ONB_VORT2DTR / ONB_VORT2D): no ODR collision as in the reference's own library.
#include "iface_common.hpp"
#include "onbody_bh2dvort.h"

#define ONB_EXPORT extern "C" __attribute__((visibility("default")))

ONB_EXPORT float external_vel_solver_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                                        const int* ntarg, const float* tx, const float* ty, float* tu, float* tv) {
    const float* X[2] = {sx, sy}; const float* S[1] = {ss}; const float* T[2] = {tx, ty}; float* O[2] = {tu, tv};
    return onb_shim::run(ONB_VORT2D, false, 1.3f /* interface2dvort.cpp:193 */, *nsrc, X, 2, S, 1, sr, *ntarg, T, nullptr, O, 2);
}
ONB_EXPORT float external_vel_direct_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                                        const int* ntarg, const float* tx, const float* ty, float* tu, float* tv) {
    const float* X[2] = {sx, sy}; const float* S[1] = {ss}; const float* T[2] = {tx, ty}; float* O[2] = {tu, tv};
    return onb_shim::run(ONB_VORT2D, true, 0.f, *nsrc, X, 2, S, 1, sr, *ntarg, T, nullptr, O, 2);
}
ONB_EXPORT float external_vel_solver_tr_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                                           const int* ntarg, const float* tx, const float* ty, const float* tr, float* tu, float* tv) {
    const float* X[2] = {sx, sy}; const float* S[1] = {ss}; const float* T[2] = {tx, ty}; float* O[2] = {tu, tv};
    return onb_shim::run(ONB_VORT2DTR, false, 1.3f /* interface2dvorttr.cpp:189 */, *nsrc, X, 2, S, 1, sr, *ntarg, T, tr, O, 2);
}
ONB_EXPORT float external_vel_direct_tr_f_(const int* nsrc, const float* sx, const float* sy, const float* ss, const float* sr,
                                           const int* ntarg, const float* tx, const float* ty, const float* tr, float* tu, float* tv) {
    const float* X[2] = {sx, sy}; const float* S[1] = {ss}; const float* T[2] = {tx, ty}; float* O[2] = {tu, tv};
    return onb_shim::run(ONB_VORT2DTR, true, 0.f, *nsrc, X, 2, S, 1, sr, *ntarg, T, tr, O, 2);
}
