// inst_ws_d2p.cu -- one group of kernel instantiations (see okb_variant.h): SET_D2P = the three pure second derivatives
// (D=3), the second pass of rho + laplacian when the first pass (SET_GRAD, epilogue 3) left the MO values in HBM.
// Same warp layout as the gradient kernels: 4 consumer + 8 producer warps, P = 32.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    OKB_WS(SET_D2P, 11, 1, 1, 4, 8, 3, SINK_RHO), OKB_WS(SET_D2P, 12, 1, 1, 4, 8, 3, SINK_RHO),
    // (12 producer warps with the 88-wide tile: second pass 146 instead of 143 ms -- not kept)
    OKB_WS(SET_D2P, 3, 1, 1, 4, 12, 3, SINK_RHO), OKB_WS(SET_D2P, 3, 1, 1, 4, 8, 3, SINK_RHO),
    // 80-wide tile: less padding for MO counts such as 222 (3 x 80 instead of 3 x 88) or 160
    OKB_WS(SET_D2P, 10, 1, 1, 4, 8, 3, SINK_RHO),
};
OKB_TABLE(okb_variants_d2p, table);

}  // namespace okb
