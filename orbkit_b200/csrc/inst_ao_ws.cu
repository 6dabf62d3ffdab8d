// inst_ao_ws.cu -- instantiations of the warp-specialised SINK_AO kernel (okb_ao_ws.cuh); see okb_variant.h.
// Not the default (see g_tables in okb200.cu): OKB_AO_VARIANT=aows selects them for A/B runs.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    OKB_AO_WS(SET_VAL, 4, 15, 4, 2, 1), OKB_AO_WS(SET_VAL, 2, 7, 4, 2, 2),
};
OKB_TABLE(okb_variants_aows, table);

}  // namespace okb
