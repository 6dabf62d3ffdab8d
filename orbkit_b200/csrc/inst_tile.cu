// inst_tile.cu -- one group of kernel instantiations (see okb_variant.h).
// Warp-specialised DMMA kernels: NPW producer warps + WM x WN consumer warps, NST stages; MO tile
// MC = 8*MB (MB blocks split over the WM warp rows), point tile P = 8*BN*WN.  MO-tile widths per set: a
// wide tile (96), the 88-wide tile that fits the 82 occupied MOs of the ~1000-function benchmark molecule,
// and a narrow tile for small MO counts.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    // AO sinks (no contraction): P = 128 points for values, fewer for the derivative sets
    // (the first entry of a set is the default; OKB_AO_VARIANT=<substring of the name> picks another one for A/B runs)
    OKB_VARIANT_AO(SET_VAL, 4, 8, 2, 2), OKB_VARIANT_AO(SET_VAL, 2, 8, 2, 3),
    OKB_VARIANT_AO(SET_ONE, 4, 8, 2, 2),
    OKB_VARIANT(SET_GRAD, 1, 2, 8, SINK_AO), OKB_VARIANT(SET_LAP, 1, 1, 8, SINK_AO),
    OKB_VARIANT(SET_ALL, 1, 1, 8, SINK_AO),
};
OKB_TABLE(okb_variants_tile, table);

}  // namespace okb
