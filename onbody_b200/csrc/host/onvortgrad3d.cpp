// onvortgrad3d - B200 build of the reference driver src/onvortgrad3d.cpp (velocity + 9 gradients, PD 3 SD 3 OD 12;
// no dual-tree method in the reference: test_iterations = {1,1,1,1,0}, onvortgrad3d.cpp:264; single -t)
#include "driver_common.hpp"
int main(int argc, char* argv[]) {
    static const DriverSpec spec = { "onvortgrad3d", ONB_VORTGRAD3D, 3, 3, 12, 1, false, false, nullptr, 1.0f, 1.0f, 1.0f, 1.0f };
    return run_driver(argc, argv, spec);
}
