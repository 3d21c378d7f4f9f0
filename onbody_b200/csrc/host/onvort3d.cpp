// onvort3d - B200 build of the reference driver src/onvort3d.cpp (Biot-Savart vortex particles, PD 3 SD 3 OD 3)
#include "driver_common.hpp"
int main(int argc, char* argv[]) {
    static const DriverSpec spec = { "onvort3d", ONB_VORT3D, 3, 3, 3, 1, true, true, nullptr, 2.0f, 1.05f, 1.0f, 1.35f };
    return run_driver(argc, argv, spec);
}
