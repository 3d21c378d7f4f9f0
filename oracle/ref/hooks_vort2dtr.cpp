// TEST INFRASTRUCTURE ONLY - wraps the unmodified reference driver onvort2d.cpp (2D, with target radius).
#define main onbody_ref_unused_main
#include "onvort2d.cpp"
#undef main
#define OREF_PD 2
#define OREF_SD 1
#define OREF_OD 2
#define OREF_HAS_FASTSUMM 1
#include "hooks_common.hpp"
