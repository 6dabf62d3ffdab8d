// okb_variant.h -- kernel variant records.  The template instantiations are spread over several
// translation units (inst_*.cu, via okb_variant_inst.h) so that they compile in parallel; okb200.cu only
// walks the per-unit tables declared here.
#pragma once
#include "okb_common.cuh"

namespace okb {

struct Variant {
    const char *name;
    int set, sink, MW, PT, NW;
    int P, MC;
    size_t (*smem)(int meta_stride);
    cudaError_t (*launch)(const KParams &, int grid, size_t smem, cudaStream_t);
    int rem = 0;                   // remainder orbitals of the MO tile contracted by the producer warps (MC includes them)
};
struct VariantTable { const Variant *v; int n; };

// defined in inst_ao_ws.cu, inst_tile.cu, inst_ws_val.cu, inst_ws_grad.cu, inst_ws_lap.cu, inst_ws_all.cu, inst_ws_d2.cu, inst_ws_d2p.cu, inst_ws_rem.cu, inst_ws_narrow_a.cu, inst_ws_narrow_b.cu
extern const VariantTable okb_variants_tile, okb_variants_val, okb_variants_grad, okb_variants_lap, okb_variants_all,
    okb_variants_d2, okb_variants_d2p, okb_variants_aows, okb_variants_rem, okb_variants_narrow_a, okb_variants_narrow_b;

// z-run SINK_AO kernel for regular grids (inst_ao_zrun.cu)
cudaError_t okb_launch_ao_zrun(const KParams &p, int sm_count, cudaStream_t st);
const char *okb_ao_zrun_name();
const char *okb_ao_zrun_code_name(int code);     // kernel name of a launch for derivative code 0..6

}  // namespace okb
