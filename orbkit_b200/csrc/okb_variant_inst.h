// okb_variant_inst.h -- launchers of the kernel templates and the macros that build Variant records;
// included by the inst_*.cu translation units only.
#pragma once
#include "okb_variant.h"
#include "okb_tile_kernel.cuh"
#include "okb_ws.cuh"

namespace okb {

template <int SET, int MW, int PT, int NW, int SINK>
inline cudaError_t launch_variant(const KParams &p, int grid, size_t smem, cudaStream_t st) {
    auto kern = okb_grid_kernel<SET, MW, PT, NW, SINK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NW * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <int SET, int MW, int PT, int NW, int SINK>
inline size_t smem_variant(int meta_stride) {
    return Cfg<SET, MW, PT, NW, SINK>::smem_bytes(meta_stride);
}
template <int SET, int MB, int BN, int WM, int WN, int NPW, int NST, int SINK>
inline cudaError_t launch_ws(const KParams &p, int grid, size_t smem, cudaStream_t st) {
    auto kern = okb_ws_kernel<SET, MB, BN, WM, WN, NPW, NST, SINK>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, (WM * WN + NPW) * 32, smem, st>>>(p);
    return cudaGetLastError();
}
template <int SET, int MB, int BN, int WM, int WN, int NPW, int NST, int SINK>
inline size_t smem_ws(int meta_stride) {
    return WsCfg<SET, MB, BN, WM, WN, NPW, NST, SINK>::smem_bytes(meta_stride);
}
#define OKB_WS(SET, MB, BN, WM, WN, NPW, NST, SINK)                                                           \
    Variant { "ws-dmma/" #SET "/" #SINK "/MB" #MB "xBN" #BN "xWM" #WM "xWN" #WN "xNPW" #NPW "xNST" #NST, SET, SINK, \
              MB, BN, WM * WN, 8 * BN * WN, 8 * MB, smem_ws<SET, MB, BN, WM, WN, NPW, NST, SINK>,               \
              launch_ws<SET, MB, BN, WM, WN, NPW, NST, SINK> }
#define OKB_VARIANT(SET, MW, PT, NW, SINK)                                                         \
    Variant { #SET "/" #SINK "/MW" #MW "xPT" #PT "xNW" #NW, SET, SINK, MW, PT, NW, 32 * PT, NW * MW, \
              smem_variant<SET, MW, PT, NW, SINK>, launch_variant<SET, MW, PT, NW, SINK> }



#define OKB_TABLE(NAME, ARR) extern const VariantTable NAME = {ARR, (int)(sizeof(ARR) / sizeof(ARR[0]))}

}  // namespace okb
