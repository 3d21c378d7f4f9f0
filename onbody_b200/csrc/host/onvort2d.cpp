// onvort2d - B200 build of the reference driver src/onvort2d.cpp (2-D vortex particles with target radius, PD 2 SD 1 OD 2)
#include "driver_common.hpp"
int main(int argc, char* argv[]) {
    static const DriverSpec spec = { "onvort2d", ONB_VORT2DTR, 2, 1, 2, 0, true, true, nullptr, 2.0f, 1.05f, 1.0f, 1.35f };
    return run_driver(argc, argv, spec);
}
