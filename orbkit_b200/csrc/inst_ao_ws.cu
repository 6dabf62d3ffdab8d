// inst_ao_ws.cu -- instantiations of the warp-specialised SINK_AO kernel (okb_ao_ws.cuh); see okb_variant.h.
// Default for the derivative sets; for plain values the tile kernel stays the default (same throughput, no alignment
// requirement) and OKB_AO_VARIANT=aows selects these for A/B runs.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    OKB_AO_WS(SET_VAL, 4, 15, 4, 2, 1), OKB_AO_WS(SET_VAL, 2, 7, 4, 2, 2),
    // derivative sets: the default kernels of SINK_AO (gradient set: 2.0 TB/s stored against 1.0 TB/s of the tile kernel)
    OKB_AO_WS(SET_GRAD, 2, 15, 3, 2, 1), OKB_AO_WS(SET_GRAD, 1, 7, 3, 1, 2),
    OKB_AO_WS(SET_ONE, 2, 7, 4, 2, 2), OKB_AO_WS(SET_ONE, 4, 15, 4, 2, 1),
    OKB_AO_WS(SET_LAP, 1, 15, 3, 1, 1),
    // (all ten sets -- the generic generators on the Cartesian layout -- stay with the tile kernel: 8.3 ms against 21.9 ms
    // per 1e5 points x 1000 AOs with two stages of 80 KB here)
};
OKB_TABLE(okb_variants_aows, table);

}  // namespace okb
