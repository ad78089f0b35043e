// scpp_b200/csrc/kernels_inst.cu — explicit instantiation of one group of kernels per translation unit.
// Compile with -DSCPP_KERNEL_MODEL=0|1|2|3 (RocketQuat | Rocket2d | Rocket2dPlugin | RocketQuatRollPlugin) and -DSCPP_KERNEL_GROUP=0..5 (scpp_b200/build.py runs them in parallel).
#define SCPP_KERNEL_INST 1
#include "kernels.cuh"

namespace scpp {
#if SCPP_KERNEL_MODEL == 0
#define SCPP_M RocketQuat
#elif SCPP_KERNEL_MODEL == 1
#define SCPP_M Rocket2d
#elif SCPP_KERNEL_MODEL == 2
#define SCPP_M Rocket2dPlugin
#else
#define SCPP_M RocketQuatRollPlugin
#endif
#if SCPP_KERNEL_GROUP == 0
SCPP_GROUP0(, SCPP_M)
#elif SCPP_KERNEL_GROUP == 1
SCPP_GROUP1(, SCPP_M)
#elif SCPP_KERNEL_GROUP == 2
SCPP_GROUP2(, SCPP_M)
#elif SCPP_KERNEL_GROUP == 3
SCPP_GROUP3(, SCPP_M)
#elif SCPP_KERNEL_GROUP == 4
SCPP_GROUP4(, SCPP_M)
#else
SCPP_GROUP5(, SCPP_M)
#endif
} // namespace scpp
