// kernel_inst.cu -- instantiates the sweep kernels of ONE compile-time shape.
// Built once per shape of riccati_configs.h with
//   -DAB2_NX=<nx> -DAB2_NU=<nu> -DAB2_NC=<nc> -DAB2_G=<lanes per instance>
#include "riccati_launch.cuh"

#define AB2_CAT_(a, b, c, d) a##b##_##c##_##d
#define AB2_CAT(a, b, c, d) AB2_CAT_(a, b, c, d)

namespace ab2 {
extern const KernelEntry AB2_CAT(kEntry_, AB2_NX, AB2_NU, AB2_NC);
const KernelEntry AB2_CAT(kEntry_, AB2_NX, AB2_NU, AB2_NC) = {
    AB2_NX,
    AB2_NU,
    AB2_NC,
    AB2_G,
    Cfg<AB2_NX, AB2_NU, AB2_NC, AB2_G>::SREC_PAD,
    &group_doubles_cfg<AB2_NX, AB2_NU, AB2_NC, AB2_G>,
    &launch_cfg<AB2_NX, AB2_NU, AB2_NC, AB2_G>};
} // namespace ab2
