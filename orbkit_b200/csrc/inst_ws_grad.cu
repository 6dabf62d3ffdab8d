// inst_ws_grad.cu -- one group of kernel instantiations (see okb_variant.h).
// Warp-specialised DMMA kernels: NPW producer warps + WM x WN consumer warps, NST stages; MO tile
// MC = 8*MB (MB blocks split over the WM warp rows), point tile P = 8*BN*WN.  MO-tile widths per set: a
// wide tile (96), the 88-wide tile that fits the 82 occupied MOs of the ~1000-function benchmark molecule,
// and a narrow tile for small MO counts.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    // value + gradient (D=4): 4 consumer + 8 producer warps, P = 32
    OKB_WS(SET_GRAD, 11, 1, 1, 4, 8, 3, SINK_MO), OKB_WS(SET_GRAD, 11, 1, 1, 4, 8, 3, SINK_RHO),
    OKB_WS(SET_GRAD, 12, 1, 1, 4, 8, 3, SINK_MO), OKB_WS(SET_GRAD, 12, 1, 1, 4, 8, 3, SINK_RHO),
    // narrow MO tiles are bound by the AO generation: 12 producer warps (listed first: wins the tie of the cost model)
    OKB_WS(SET_GRAD, 3, 1, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_GRAD, 3, 1, 1, 4, 12, 3, SINK_RHO),
    OKB_WS(SET_GRAD, 3, 1, 1, 4, 8, 3, SINK_MO), OKB_WS(SET_GRAD, 3, 1, 1, 4, 8, 3, SINK_RHO),
    // (two consumer warps per sub-partition -- WM=2 with 4 or 8 producer warps -- measured 265 / 238 ms against
    // 192 ms on the benchmark: DMMA wins the FP64 pipe arbitration and the producers starve)
    // 80-wide tile: less padding for MO counts such as 222 (3 x 80 instead of 3 x 88) or 160
    OKB_WS(SET_GRAD, 10, 1, 1, 4, 8, 3, SINK_MO), OKB_WS(SET_GRAD, 10, 1, 1, 4, 8, 3, SINK_RHO),
};
OKB_TABLE(okb_variants_grad, table);

}  // namespace okb
