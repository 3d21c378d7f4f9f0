// TEST INFRASTRUCTURE ONLY - the unmodified reference driver onvort3d.cpp (compiled in place), ACCUM = double.
#define main onbody_ref_unused_main
#include "onvort3d.cpp"
#undef main
#define OREF_PD 3
#define OREF_SD 3
#define OREF_OD 3
#define OREF_HAS_FASTSUMM 1
#define OREF_ACCUM double
#include "hooks_common.hpp"
