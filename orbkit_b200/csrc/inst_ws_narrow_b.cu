// inst_ws_narrow_b.cu -- narrow MO tiles of the value set (see inst_ws_narrow_a.cu): 256-point tiles for 24 orbitals,
// a 48-orbital tile for MO counts between 25 and 48.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    // 21 orbitals, 150^3 points: 3.03 -> 2.74 ms with 256-point tiles and two stages
    OKB_WS(SET_VAL, 3, 8, 1, 4, 12, 2, SINK_MO), OKB_WS(SET_VAL, 3, 8, 1, 4, 12, 2, SINK_RHO),
    // (48 orbitals with 256-point tiles: 3.46 against 3.55 ms -- not kept)
    OKB_WS(SET_VAL, 6, 4, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_VAL, 6, 4, 1, 4, 12, 3, SINK_RHO),
    // 32-point tiles for SMALL requests only (pick_variant: fewer points than 32 per SM -- the cubature use case, a
    // thousand new points per call): a 128-point tile would leave all but a few SMs idle and take longer per chunk
    OKB_WS(SET_VAL, 11, 1, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_VAL, 11, 1, 1, 4, 12, 3, SINK_RHO),
    OKB_WS(SET_VAL, 3, 1, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_VAL, 3, 1, 1, 4, 12, 3, SINK_RHO),
};
OKB_TABLE(okb_variants_narrow_b, table);

}  // namespace okb
