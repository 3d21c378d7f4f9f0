// TEST INFRASTRUCTURE ONLY - wraps the unmodified reference driver ongrav3d.cpp (compiled in place).
#define main onbody_ref_unused_main
#include "ongrav3d.cpp"
#undef main
#define OREF_PD 3
#define OREF_SD 1
#define OREF_OD 3
#define OREF_HAS_FASTSUMM 1
#include "hooks_common.hpp"
