// TEST INFRASTRUCTURE ONLY - wraps the unmodified reference interface3dvortgrads.cpp (same kernel as
// onvortgrad3d.cpp:45-240 but without the dead Eigen include); also re-exports the reference's own
// external_vel_solver_f_ / external_vel_direct_f_ (interface3dvortgrads.cpp:247,422).
#include <random>
#pragma GCC visibility push(default)
#include "interface3dvortgrads.cpp"
#pragma GCC visibility pop
#define OREF_PD 3
#define OREF_SD 3
#define OREF_OD 12
#define OREF_HAS_FASTSUMM 0
#include "hooks_common.hpp"
