// inst_ws_d2.cu -- one group of kernel instantiations (see okb_variant.h): SET_D2 = value + the three pure second
// derivatives (D=4), the second pass of the two-pass rho + laplacian (first pass: SET_GRAD with the squared-gradient
// epilogue).  Same warp layout as the gradient kernels: 4 consumer + 8 producer warps, P = 32.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    OKB_WS(SET_D2, 11, 1, 1, 4, 8, 3, SINK_RHO), OKB_WS(SET_D2, 12, 1, 1, 4, 8, 3, SINK_RHO),
    OKB_WS(SET_D2, 3, 1, 1, 4, 8, 3, SINK_RHO),
};
OKB_TABLE(okb_variants_d2, table);

}  // namespace okb
