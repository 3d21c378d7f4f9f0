// TEST INFRASTRUCTURE ONLY - wraps the unmodified reference interface2dvort.cpp (2D, no target radius);
// also re-exports the reference's own external_vel_solver_f_ / external_vel_direct_f_ (interface2dvort.cpp:182,324).
#include <random>
#pragma GCC visibility push(default)
#include "interface2dvort.cpp"
#pragma GCC visibility pop
#define OREF_PD 2
#define OREF_SD 1
#define OREF_OD 2
#define OREF_HAS_FASTSUMM 0
#include "hooks_common.hpp"
