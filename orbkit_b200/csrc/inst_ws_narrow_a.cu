// inst_ws_narrow_a.cu -- narrow MO tiles (24 and 48 orbitals: the occupied orbitals of small and medium molecules) of the
// derivative sets.  These kernels are bound by the AO generation (profiles/r02_c2_narrow.txt: producers 75 % of the stall
// samples, consumers wait for full stages), so the point tile is widened at the price of a stage: two points per producer
// thread (independent dependency chains) and half as many tile switches.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    // 24 orbitals: 64-point tiles, two stages (Config-2 shape, 21 orbitals, 150^3 points: rho + grad rho 8.69 -> 7.23 ms);
    // second laplacian pass: 96-point tiles (7.8 -> 6.7 ms).  (D2P with 64-point tiles and three stages needs 249 KB.)
    OKB_WS(SET_GRAD, 3, 2, 1, 4, 12, 2, SINK_MO), OKB_WS(SET_GRAD, 3, 2, 1, 4, 12, 2, SINK_RHO),
    OKB_WS(SET_D2P, 3, 3, 1, 4, 12, 2, SINK_RHO),
    // 48 orbitals (25 .. 48 MOs went to the 80-wide tile before: 40 orbitals 17.4 -> 11.3 ms); 64-point tiles with two
    // stages are slower here (11.46 ms: the coefficient tiles are larger and the consumers need the third stage)
    OKB_WS(SET_GRAD, 6, 1, 1, 4, 12, 3, SINK_MO), OKB_WS(SET_GRAD, 6, 1, 1, 4, 12, 3, SINK_RHO),
    OKB_WS(SET_D2P, 6, 1, 1, 4, 12, 3, SINK_RHO),
};
OKB_TABLE(okb_variants_narrow_a, table);

}  // namespace okb
