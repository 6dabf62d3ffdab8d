// inst_ao_zrun.cu -- the z-run SINK_AO kernel for regular grids (okb_ao_zrun.cuh) and its launcher.
#include <stdlib.h>

#include <algorithm>

#include "okb_ao_zrun.cuh"
#include "okb_variant.h"

namespace okb {

// A/B builds of the kernel: OKB_ZRUN=<J><PF>, e.g. "40" = 4 rows per thread without prefetch; default "44"
static int zrun_variant() {
    static const char *e = getenv("OKB_ZRUN");
    if (e && e[0] && e[1]) return (e[0] - '0') * 10 + (e[1] - '0');
    return 44;
}

const char *okb_ao_zrun_name() {
    switch (zrun_variant()) {
        case 40: return "zrun/SET_VAL/SINK_AO/J4xPF0";
        case 84: return "zrun/SET_VAL/SINK_AO/J8xPF4";
        case 24: return "zrun/SET_VAL/SINK_AO/J2xPF4";
        case 48: return "zrun/SET_VAL/SINK_AO/J4xPF8";
        default: return "zrun/SET_VAL/SINK_AO/J4xPF4";
    }
}

template <int J, int PF>
static cudaError_t launch(const KParams &p, int sm_count, cudaStream_t st) {
    const long long row_first = p.p0 / p.nz, row_last = (p.p0 + p.npts - 1) / p.nz;
    const long long ngroups = (row_last - row_first + 1 + J - 1) / J;
    const int block = std::min(256, (p.nz + 31) / 32 * 32);
    const int nzb = (p.nz + block - 1) / block;
    // RG row groups per CTA (the chunk's tabz slices stay in L1 across them): as many as still leave two waves of CTAs
    static const char *frg = getenv("OKB_ZRUN_RG");
    int RG = (int)std::min<long long>(8, std::max<long long>(1, ngroups * nzb / ((long long)sm_count * 8)));
    if (frg && frg[0]) RG = atoi(frg);
    const long long grid = (ngroups + RG - 1) / RG * nzb;
    if (grid <= 0 || grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    okb_ao_zrun_kernel<J, PF><<<(unsigned)grid, block, 0, st>>>(p, row_first, row_last, nzb, RG);
    return cudaGetLastError();
}

// p: a SINK_AO request on a regular grid with axis tables (p.tabx != null), p.one_code = the single code 0, p.slot[]
// set; the points [p.p0, p.p0 + p.npts) may start and end inside a z run
cudaError_t okb_launch_ao_zrun(const KParams &p, int sm_count, cudaStream_t st) {
    switch (zrun_variant()) {
        case 40: return launch<4, 0>(p, sm_count, st);
        case 84: return launch<8, 4>(p, sm_count, st);
        case 24: return launch<2, 4>(p, sm_count, st);
        case 48: return launch<4, 8>(p, sm_count, st);
        default: return launch<4, 4>(p, sm_count, st);
    }
}

}  // namespace okb
