/*
 * onbody_bh3dvortgrads.h - drop-in for the reference's libbh3dvortgrads (CMakeLists.txt:154-156):
 * the two Fortran-callable entry points of src/interface3dvortgrads.cpp, same names, same arguments, same
 * "+=" output convention, backed by the CUDA library (libbh3dvortgrads_b200.so -> libonbody_b200.so).
 *
 * Conventions kept from the reference (interface3dvortgrads.cpp:247-256, 384-395, 422-431, 484-495):
 *   - trailing underscore, every scalar by pointer, int counts; all arrays are caller-owned host arrays
 *   - inputs are copied, never modified; results are ACCUMULATED into tu..twz in the caller's original target order
 *     (the caller pre-zeroes); the return value is the flop estimate
 *   - fixed parameters of the solver: theta = 1.5, order = 4, block = 128, boxwise traversal (:259-263)
 *   - blocking: results are valid on return; the device context is created lazily on first call and reused
 * Differences: no CPU path exists. If no B200 is usable the call prints the reason to stderr and aborts (set
 * ONBODY_B200_ON_ERROR=return to get -1.0f back instead).
 */
#ifndef ONBODY_BH3DVORTGRADS_H
#define ONBODY_BH3DVORTGRADS_H
#ifdef __cplusplus
extern "C" {
#endif

/* replaces interface3dvortgrads.cpp:247 */
float external_vel_solver_f_(const int* nsrc, const float* sx, const float* sy, const float* sz,
                             const float* ssx, const float* ssy, const float* ssz, const float* sr,
                             const int* ntarg, const float* tx, const float* ty, const float* tz,
                             float* tu, float* tv, float* tw, float* tux, float* tvx, float* twx,
                             float* tuy, float* tvy, float* twy, float* tuz, float* tvz, float* twz);
/* replaces interface3dvortgrads.cpp:422 */
float external_vel_direct_f_(const int* nsrc, const float* sx, const float* sy, const float* sz,
                             const float* ssx, const float* ssy, const float* ssz, const float* sr,
                             const int* ntarg, const float* tx, const float* ty, const float* tz,
                             float* tu, float* tv, float* tw, float* tux, float* tvx, float* twx,
                             float* tuy, float* tvy, float* twy, float* tuz, float* tvz, float* twz);

#ifdef __cplusplus
}
#endif
#endif
