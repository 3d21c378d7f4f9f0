// TEST INFRASTRUCTURE ONLY - the reference's interface2dvorttr.cpp alone (external_vel_solver_tr_f_ /
// external_vel_direct_tr_f_, interface2dvorttr.cpp:177,321), kept in its own library to dodge the ODR
// collision the reference's own bh2dvort target has.
#include <random>
#pragma GCC visibility push(default)
#include "interface2dvorttr.cpp"
#pragma GCC visibility pop
#define OREF_PD 2
#define OREF_SD 1
#define OREF_OD 2
#define OREF_HAS_FASTSUMM 0
#include "hooks_common.hpp"
