// inst_ao_zrun.cu -- the z-run SINK_AO kernel for regular grids (okb_ao_zrun.cuh) and its launcher.
#include <algorithm>

#include "okb_ao_zrun.cuh"
#include "okb_variant.h"

namespace okb {

const char *okb_ao_zrun_name() { return "zrun/SET_VAL/SINK_AO/J4"; }

// p: a SINK_AO request on a regular grid with axis tables (p.tabx != null), p.one_code = the single code 0, p.slot[]
// set; the points [p.p0, p.p0 + p.npts) may start and end inside a z run
cudaError_t okb_launch_ao_zrun(const KParams &p, int sm_count, cudaStream_t st) {
    constexpr int J = 4;
    const long long row_first = p.p0 / p.nz, row_last = (p.p0 + p.npts - 1) / p.nz;
    const long long ngroups = (row_last - row_first + 1 + J - 1) / J;
    const int block = std::min(256, (p.nz + 31) / 32 * 32);
    const int nzb = (p.nz + block - 1) / block;
    // RG row groups per CTA (the chunk's tabz slices stay in L1 across them): as many as still leave ~8 CTAs per SM
    int RG = (int)std::min<long long>(8, std::max<long long>(1, ngroups * nzb / ((long long)sm_count * 8)));
    const long long grid = (ngroups + RG - 1) / RG * nzb;
    if (grid <= 0 || grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    okb_ao_zrun_kernel<J><<<(unsigned)grid, block, 0, st>>>(p, row_first, row_last, nzb, RG);
    return cudaGetLastError();
}

}  // namespace okb
