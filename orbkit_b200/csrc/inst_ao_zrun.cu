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

template <int J, int PF, int NPOLY>
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
    const size_t smem = 2 * sizeof(ZrunSmem<J, NPOLY>);         // two-deep ring of the uniform quantities
    auto kern = okb_ao_zrun_kernel<J, PF, NPOLY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<(unsigned)grid, block, smem, st>>>(p, row_first, row_last, nzb, RG);
    return cudaGetLastError();
}

// rows per thread of the derivative kernels: 4 (first derivatives 3.8 against 2.9 TB/s stored with 2; second derivatives --
// three polynomial sets per row, 77 KB of shared memory per CTA -- 2.5 against 2.25 TB/s).  A/B: OKB_ZRUN_DJ=2|4 forces one.
static int zrun_dj(int npoly) {
    static const char *e = getenv("OKB_ZRUN_DJ");
    if (e && e[0]) return e[0] == '4' ? 4 : 2;
    (void)npoly;
    return 4;
}

// p: a SINK_AO request on a regular grid with axis tables (p.tabx != null), p.one_code = the ONE code 0..6 of this launch,
// p.slot[] set; the points [p.p0, p.p0 + p.npts) may start and end inside a z run
cudaError_t okb_launch_ao_zrun(const KParams &p, int sm_count, cudaStream_t st) {
    if (p.one_code >= 4) return zrun_dj(3) == 4 ? launch<4, 4, 3>(p, sm_count, st) : launch<2, 4, 3>(p, sm_count, st);
    if (p.one_code >= 1) return zrun_dj(2) == 4 ? launch<4, 4, 2>(p, sm_count, st) : launch<2, 4, 2>(p, sm_count, st);
    switch (zrun_variant()) {
        case 40: return launch<4, 0, 1>(p, sm_count, st);
        case 84: return launch<8, 4, 1>(p, sm_count, st);
        case 24: return launch<2, 4, 1>(p, sm_count, st);
        case 48: return launch<4, 8, 1>(p, sm_count, st);
        default: return launch<4, 4, 1>(p, sm_count, st);
    }
}

const char *okb_ao_zrun_code_name(int code) {
    static const char *const names[7] = {nullptr, "zrun/ONE1/SINK_AO", "zrun/ONE2/SINK_AO", "zrun/ONE3/SINK_AO",
                                         "zrun/ONE4/SINK_AO", "zrun/ONE5/SINK_AO", "zrun/ONE6/SINK_AO"};
    return (code >= 1 && code <= 6) ? names[code] : okb_ao_zrun_name();
}

}  // namespace okb
