// ongrav3d - B200 build of the reference driver src/ongrav3d.cpp (gravity / electrostatics, PD 3 SD 1 OD 3)
#include "driver_common.hpp"
int main(int argc, char* argv[]) {
    static const DriverSpec spec = { "ongrav3d", ONB_GRAV3D, 3, 1, 3, 0, true, true,
                                     "  electrostatics simulation with random charges", 2.0f, 1.05f, 1.0f, 1.35f };   // ongrav3d.cpp:477-480,586
    return run_driver(argc, argv, spec);
}
