// inst_ws_rem.cu -- one group of kernel instantiations (see okb_variant.h): MO tiles of 8*MB + 2 orbitals whose two
// remainder orbitals are contracted by the producer warps (okb_ws.cuh, REM).  82-wide tile = 10 DMMA blocks + 2: the 82
// occupied MOs of the benchmark molecule run 40 instead of 44 DMMAs per k-step; 246 MOs = 3 x 82.  Gradient and
// second-derivative sets only: for plain values (D = 1) the producers are the bottleneck already.
#include "okb_variant_inst.h"

namespace okb {

static const Variant table[] = {
    OKB_WSR(SET_GRAD, 10, 1, 1, 4, 8, 3, SINK_RHO, 2), OKB_WSR(SET_GRAD, 10, 1, 1, 4, 8, 3, SINK_MO, 2),
    OKB_WSR(SET_D2P, 10, 1, 1, 4, 8, 3, SINK_RHO, 2),
    // 74-wide tile = 9 blocks + 2: 222 MOs (benzene def2-TZVP, BASELINE configs[1]) = 3 x 74 instead of 3 x 80
    OKB_WSR(SET_GRAD, 9, 1, 1, 4, 8, 3, SINK_RHO, 2), OKB_WSR(SET_GRAD, 9, 1, 1, 4, 8, 3, SINK_MO, 2),
    OKB_WSR(SET_D2P, 9, 1, 1, 4, 8, 3, SINK_RHO, 2),
};
OKB_TABLE(okb_variants_rem, table);

}  // namespace okb
