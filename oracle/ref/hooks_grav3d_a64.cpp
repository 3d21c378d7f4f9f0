// TEST INFRASTRUCTURE ONLY - the unmodified reference driver ongrav3d.cpp (compiled in place) with its templates
// instantiated for ACCUM = double (README.md:107-112: fp32 storage, fp64 accumulation).
#define main onbody_ref_unused_main
#include "ongrav3d.cpp"
#undef main
#define OREF_PD 3
#define OREF_SD 1
#define OREF_OD 3
#define OREF_HAS_FASTSUMM 1
#define OREF_ACCUM double
#include "hooks_common.hpp"
