// inst_ws_val.cu -- one group of kernel instantiations (see okb_variant.h).
// Warp-specialised DMMA kernels: NPW producer warps + WM x WN consumer warps, NST stages; MO tile
// MC = 8*MB (MB blocks split over the WM warp rows), point tile P = 8*BN*WN.  MO-tile widths per set: a
// wide tile (96), the 88-wide tile that fits the 82 occupied MOs of the ~1000-function benchmark molecule,
// and a narrow tile for small MO counts.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    // value only (D=1): 4 consumer + 12 producer warps (AO generation dominates), P = 128
    OKB_WS(SET_VAL, 11, 4, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_VAL, 11, 4, 1, 4, 12, 3, SINK_RHO),
    OKB_WS(SET_VAL, 12, 4, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_VAL, 12, 4, 1, 4, 12, 3, SINK_RHO),
    OKB_WS(SET_VAL, 3, 4, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_VAL, 3, 4, 1, 4, 12, 3, SINK_RHO),
    OKB_WS(SET_ONE, 12, 4, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_ONE, 3, 4, 1, 4, 12, 3, SINK_MO),
    // 80-wide tile: less padding for MO counts such as 222 (3 x 80 instead of 3 x 88) or 160
    OKB_WS(SET_VAL, 10, 4, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_VAL, 10, 4, 1, 4, 12, 3, SINK_RHO),
};
OKB_TABLE(okb_variants_val, table);

}  // namespace okb
