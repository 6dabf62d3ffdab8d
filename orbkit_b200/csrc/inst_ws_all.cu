// inst_ws_all.cu -- one group of kernel instantiations (see okb_variant.h).
// Warp-specialised DMMA kernels: NPW producer warps + WM x WN consumer warps, NST stages; MO tile
// MC = 8*MB (MB blocks split over the WM warp rows), point tile P = 8*BN*WN.  MO-tile widths per set: a
// wide tile (96), the 88-wide tile that fits the 82 occupied MOs of the ~1000-function benchmark molecule,
// and a narrow tile for small MO counts.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    // all ten codes (D=10): 8 consumer + 4 producer warps, P = 32
    OKB_WS(SET_ALL, 6, 1, 2, 4, 4, 2, SINK_MO), OKB_WS(SET_ALL, 6, 1, 2, 4, 4, 2, SINK_RHO),
    OKB_WS(SET_ALL, 2, 1, 2, 4, 4, 2, SINK_MO), OKB_WS(SET_ALL, 2, 1, 2, 4, 4, 2, SINK_RHO),
};
OKB_TABLE(okb_variants_all, table);

}  // namespace okb
