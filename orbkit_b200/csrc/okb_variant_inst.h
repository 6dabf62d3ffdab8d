// okb_variant_inst.h -- launchers of the kernel templates and the macros that build Variant records;
// included by the inst_*.cu translation units only.
#pragma once
#include <algorithm>

#include "okb_variant.h"
#include "okb_tile_kernel.cuh"
#include "okb_ws.cuh"
#include "okb_ao_ws.cuh"

namespace okb {

// grid < 0: -grid is the number of SMs; the launcher asks the occupancy calculator how many CTAs fit on one
template <int SET, int MW, int PT, int NW, int SINK, int NPT = 0, int MINB = 1>
inline cudaError_t launch_variant(const KParams &p, int grid, size_t smem, cudaStream_t st) {
    auto kern = okb_grid_kernel<SET, MW, PT, NW, SINK, NPT, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (grid < 0) {
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem);
        if (e != cudaSuccess) return e;
        grid = std::min(p.ntiles, -grid * std::max(per_sm, 1));
    }
    kern<<<grid, NW * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <int SET, int MW, int PT, int NW, int SINK>
inline size_t smem_variant(int meta_stride) {
    return Cfg<SET, MW, PT, NW, SINK>::smem_bytes(meta_stride);
}
template <int SET, int MB, int BN, int WM, int WN, int NPW, int NST, int SINK, int REM = 0>
inline cudaError_t launch_ws(const KParams &p, int grid, size_t smem, cudaStream_t st) {
    auto kern = okb_ws_kernel<SET, MB, BN, WM, WN, NPW, NST, SINK, REM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, (WM * WN + NPW) * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <int SET, int MB, int BN, int WM, int WN, int NPW, int NST, int SINK, int REM = 0>
inline size_t smem_ws(int meta_stride) {
    return WsCfg<SET, MB, BN, WM, WN, NPW, NST, SINK, REM>::smem_bytes(meta_stride);
}
// MO tile of 8*MB + REM orbitals: the REM remainder orbitals are contracted by the producer warps (okb_ws.cuh)
#define OKB_WSR(SET, MB, BN, WM, WN, NPW, NST, SINK, REM)                                                       \
    Variant { "ws-dmma/" #SET "/" #SINK "/MB" #MB "R" #REM "xBN" #BN "xWM" #WM "xWN" #WN "xNPW" #NPW "xNST" #NST, SET, \
              SINK, MB, BN, WM * WN, 8 * BN * WN, 8 * MB + REM, smem_ws<SET, MB, BN, WM, WN, NPW, NST, SINK, REM>, \
              launch_ws<SET, MB, BN, WM, WN, NPW, NST, SINK, REM>, REM }
#define OKB_WS(SET, MB, BN, WM, WN, NPW, NST, SINK)                                                           \
    Variant { "ws-dmma/" #SET "/" #SINK "/MB" #MB "xBN" #BN "xWM" #WM "xWN" #WN "xNPW" #NPW "xNST" #NST, SET, SINK, \
              MB, BN, WM * WN, 8 * BN * WN, 8 * MB, smem_ws<SET, MB, BN, WM, WN, NPW, NST, SINK>,               \
              launch_ws<SET, MB, BN, WM, WN, NPW, NST, SINK> }
template <int SET, int PT, int NPW, int NST, int NPT, int MINB>
inline cudaError_t launch_ao_ws(const KParams &p, int grid, size_t smem, cudaStream_t st) {
    auto kern = okb_ao_ws_kernel<SET, PT, NPW, NST, NPT, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (grid < 0) {
        int per_sm = 1;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (NPW + 1) * 32, smem);
        if (e != cudaSuccess) return e;
        grid = std::min(p.ntiles, -grid * std::max(per_sm, 1));
    }
    kern<<<grid, (NPW + 1) * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <int SET, int PT, int NPW, int NST>
inline size_t smem_ao_ws(int meta_stride) {
    return AoWsCfg<SET, PT, NPW, NST>::smem_bytes(meta_stride);
}
// warp-specialised SINK_AO kernel (okb_ao_ws.cuh): names start with "aows/"; needs 16-byte aligned output rows
#define OKB_AO_WS(SET, PT, NPW, NST, NPT, MINB)                                                               \
    Variant { "aows/" #SET "/SINK_AO/PT" #PT "xNPW" #NPW "xNST" #NST "xNP" #NPT "xB" #MINB, SET, SINK_AO, 1, PT, NPW, \
              32 * PT, NPW, smem_ao_ws<SET, PT, NPW, NST>, launch_ao_ws<SET, PT, NPW, NST, NPT, MINB> }
#define OKB_VARIANT_AO(SET, PT, NW, NPT, MINB)                                                     \
    Variant { #SET "/SINK_AO/PT" #PT "xNW" #NW "xNP" #NPT "xB" #MINB, SET, SINK_AO, 1, PT, NW, 32 * PT, NW, \
              smem_variant<SET, 1, PT, NW, SINK_AO>, launch_variant<SET, 1, PT, NW, SINK_AO, NPT, MINB> }
#define OKB_VARIANT(SET, MW, PT, NW, SINK)                                                         \
    Variant { #SET "/" #SINK "/MW" #MW "xPT" #PT "xNW" #NW, SET, SINK, MW, PT, NW, 32 * PT, NW * MW, \
              smem_variant<SET, MW, PT, NW, SINK>, launch_variant<SET, MW, PT, NW, SINK> }



#define OKB_TABLE(NAME, ARR) extern const VariantTable NAME = {ARR, (int)(sizeof(ARR) / sizeof(ARR[0]))}

}  // namespace okb
