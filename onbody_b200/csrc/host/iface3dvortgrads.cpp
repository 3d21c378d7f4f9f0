// iface3dvortgrads.cpp - libbh3dvortgrads_b200.so: the reference's 3-D velocity + velocity-gradient entry points
// (src/interface3dvortgrads.cpp:247-416 and :422-500) on the B200. See include/onbody_bh3dvortgrads.h.
#include "iface_common.hpp"
#include "onbody_bh3dvortgrads.h"

#define ONB_EXPORT extern "C" __attribute__((visibility("default")))

ONB_EXPORT float external_vel_solver_f_(const int* nsrc, const float* sx, const float* sy, const float* sz,
                                        const float* ssx, const float* ssy, const float* ssz, const float* sr,
                                        const int* ntarg, const float* tx, const float* ty, const float* tz,
                                        float* tu, float* tv, float* tw, float* tux, float* tvx, float* twx,
                                        float* tuy, float* tvy, float* twy, float* tuz, float* tvz, float* twz) {
    const float* X[3] = {sx, sy, sz}; const float* S[3] = {ssx, ssy, ssz}; const float* T[3] = {tx, ty, tz};
    float* O[12] = {tu, tv, tw, tux, tvx, twx, tuy, tvy, twy, tuz, tvz, twz};
    return onb_shim::run(ONB_VORTGRAD3D, false, 1.5f /* :259 */, *nsrc, X, 3, S, 3, sr, *ntarg, T, nullptr, O, 12);
}

ONB_EXPORT float external_vel_direct_f_(const int* nsrc, const float* sx, const float* sy, const float* sz,
                                        const float* ssx, const float* ssy, const float* ssz, const float* sr,
                                        const int* ntarg, const float* tx, const float* ty, const float* tz,
                                        float* tu, float* tv, float* tw, float* tux, float* tvx, float* twx,
                                        float* tuy, float* tvy, float* twy, float* tuz, float* tvz, float* twz) {
    const float* X[3] = {sx, sy, sz}; const float* S[3] = {ssx, ssy, ssz}; const float* T[3] = {tx, ty, tz};
    float* O[12] = {tu, tv, tw, tux, tvx, twx, tuy, tvy, twy, tuz, tvz, twz};
    return onb_shim::run(ONB_VORTGRAD3D, true, 0.f, *nsrc, X, 3, S, 3, sr, *ntarg, T, nullptr, O, 12);
}
