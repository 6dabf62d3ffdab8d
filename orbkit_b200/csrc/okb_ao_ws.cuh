// okb_ao_ws.cuh -- warp-specialised kernel for the HBM-store-bound SINK_AO requests (calc_ao, cy_core.aocreator).
//
// The phase-serial tile kernel (okb_tile_kernel.cuh) spends its warps twice per AO value: once to generate it into the
// shared-memory tile and once more to read it back, scale it and store it (LDS + DMUL + STG), with a CTA-wide barrier per
// chunk between the two phases (ncu, profiles/r01_ao_tile_summary.txt: 2.4 TB/s, FP64 pipe 30 %, 13 % of the samples on
// that barrier, 16 warps per SM).  Here the store phase is taken off the warps:
//
//   producers   NPW warps evaluate (shell, 32*NP points) items of chunk g into stage g % NST of a ring of AO tiles
//               (same generators as the fused kernel), then  fence.proxy.async  +  one arrival per thread on full[s].
//               They are not synchronised with each other per chunk (bounded by the ring).
//   store warp  waits for full[s]; every lane owns one output row of the chunk and, when the row is a plain copy of a tile
//               row (one term, coefficient 1 -- every row of Cartesian bases and of spherical bases with standard shells),
//               issues ONE bulk-async copy shared -> global of the row's P*8 bytes per derivative set
//               (cp.async.bulk.global.shared::cta, the TMA engine; no register traffic, no LSU instructions).  Other rows
//               (multi-term spherical rows of non-standard shells, partial last tile) are combined and stored by the warp.
//               A stage is handed back (empty[s]) once the bulk copies of the NEXT chunk have been issued and those of
//               this one have finished reading shared memory (cp.async.bulk.wait_group.read 1).
//   chunk tables a ring of NM slots filled by bulk-async copies (issued by the first producer thread), freed when all
//               producer warps and the store warp have left the chunk.
//
// Requirements (checked by the host, which otherwise launches the tile kernel): the output base address is 16-byte
// aligned and the row stride `ld` is even, so that every row segment starts on a 16-byte boundary.
#pragma once
#include "okb_shell.cuh"
#include "okb_ws.cuh"      // barrier helpers on shared-space addresses

namespace okb {

template <int SET, int PT, int NPW, int NST>
struct AoWsCfg {
    static constexpr int D = set_ncodes(SET);
    static constexpr int P = 32 * PT;
    static constexpr int NT = (NPW + 1) * 32;
    static constexpr int TILE_DOUBLES = D * KC * P;
    static constexpr int NM = 4, LAG = 2;
    static constexpr size_t OFF_BAR = 0;                       // full[NST] empty[NST] mfull[NM] mempty[NM]
    static constexpr size_t OFF_XYZ = 256;
    static constexpr size_t OFF_IJK = OFF_XYZ + (size_t)3 * P * 8;
    static constexpr size_t OFF_META = (OFF_IJK + (size_t)3 * P * 4 + 127) / 128 * 128;
    __host__ __device__ static constexpr size_t off_tile(int meta_stride) {
        return (OFF_META + (size_t)NM * meta_stride + 127) / 128 * 128;
    }
    __host__ __device__ static constexpr size_t smem_bytes(int meta_stride) {
        return off_tile(meta_stride) + (size_t)NST * TILE_DOUBLES * 8;
    }
    static_assert(2 * NST + 2 * NM <= 32, "barrier area");
};

__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

template <int SET, int PT, int NPW, int NST, int NPT_, int MINB>
__global__ void __launch_bounds__((NPW + 1) * 32, MINB) okb_ao_ws_kernel(const KParams p) {
    using C = AoWsCfg<SET, PT, NPW, NST>;
    constexpr int D = C::D, P = C::P, NM = C::NM, LAG = C::LAG;
    constexpr int NPT = NPW * 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t a_full = sbase + (uint32_t)C::OFF_BAR, a_empty = a_full + 8 * NST, a_mfull = a_empty + 8 * NST,
                   a_mempty = a_mfull + 8 * NM;
    double *xs = reinterpret_cast<double *>(smem + C::OFF_XYZ);
    double *ys = xs + P, *zs = ys + P;
    int *isx = reinterpret_cast<int *>(smem + C::OFF_IJK), *isy = isx + P, *isz = isy + P;
    unsigned char *mbase = smem + C::OFF_META;
    double *tbase = reinterpret_cast<double *>(smem + C::off_tile(p.lay.stride));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
        for (int i = 0; i < NST; ++i) {
            mbar_init(&bars[i], NPT);                 // one arrival per producer thread
            mbar_init(&bars[NST + i], 1);             // the store warp hands the stage back
        }
        for (int i = 0; i < NM; ++i) {
            mbar_init(&bars[2 * NST + i], 1);
            mbar_init(&bars[2 * NST + NM + i], NPW + 1);   // producer warps + the store warp left the chunk
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t meta_bytes = (uint32_t)p.lay.stride;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)p.nchunk;

    if (warp >= 1) {
        // ====================================== producers ==========================================
        const int ptid = tid - 32, pwarp = warp - 1;
        const uint32_t a_meta = smem_u32(mbase);
        auto issue_meta = [&](uint32_t gc) {
            const uint32_t bar = a_mfull + 8 * (gc % NM);
            mbar_arrive_expect_tx_a(bar, meta_bytes);
            bulk_g2s_a(a_meta + (gc % NM) * meta_bytes, p.meta + (size_t)(gc % p.nchunk) * meta_bytes, meta_bytes, bar);
        };
        if (ptid == 0)
            for (uint32_t i = 0; i < NM && i < total; ++i) issue_meta(i);
        uint32_t g = 0;
        for (int tile_id = blockIdx.x; tile_id < p.ntiles; tile_id += gridDim.x) {
            const int q0 = tile_id * P;
            named_bar(1, NPT);                        // previous tile's items no longer read xs/ys/zs
            for (int e = ptid; e < P; e += NPT) {
                int q = q0 + e;
                if (q >= p.npts) q = p.npts - 1;
                grid_point(p, p.p0 + q, xs[e], ys[e], zs[e], isx[e], isy[e], isz[e]);
            }
            named_bar(1, NPT);
            for (int c = 0; c < p.nchunk; ++c, ++g) {
                const int s = g % NST;
                if (ptid == 0 && g >= LAG && g - LAG + NM < total) {
                    const uint32_t h = g - LAG;
                    mbar_wait_a(a_mempty + 8 * (h % NM), (h / NM) & 1);
                    issue_meta(h + NM);
                }
                mbar_wait_a(a_mfull + 8 * (g % NM), (g / NM) & 1);
                mbar_wait_a(a_empty + 8 * s, ((g / NST) & 1) ^ 1);       // the store warp released the stage
                const unsigned char *mb = mbase + (size_t)(g % NM) * meta_bytes;
                const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(mb);
                const ShellMeta *shells = reinterpret_cast<const ShellMeta *>(mb + p.lay.off_shell);
                const double2 *prims = reinterpret_cast<const double2 *>(mb + p.lay.off_prim);
                const FnMeta *fns = reinterpret_cast<const FnMeta *>(mb + p.lay.off_fn);
                const double *aux = reinterpret_cast<const double *>(mb + p.lay.off_aux);
                double *tile = tbase + (size_t)s * C::TILE_DOUBLES;
                constexpr int NP = NPT_ > 0 ? NPT_ : (PT % 2 == 0 && SET != SET_LAP && SET != SET_ALL) ? 2 : 1, PG = PT / NP;
                // shells are sorted by descending cost; items are dealt in boustrophedon order, rotated per chunk
                const int nitems = hdr.nshell * PG, wrot = (pwarp + g) % NPW;
                for (int r = 0; r * NPW < nitems; ++r) {
                    const int item = r * NPW + ((r & 1) ? NPW - 1 - wrot : wrot);
                    if (item >= nitems) continue;
                    const int sh = item / PG, pt = (item % PG) * (32 * NP) + lane;
                    const AxTab tab{p.tabx, p.taby, p.tabz, p.nx, p.ny, p.nz, isx + pt, isy + pt, isz + pt};
                    gen_shell_any<SET, P, NP>(shells[sh], prims, fns, aux, xs + pt, ys + pt, zs + pt, tile + pt,
                                              p.one_code, p.exact_mixed, tab);
                }
                // the tile was written through the generic proxy and is read by the async proxy (bulk copy)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_a(a_full + 8 * s);
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_mempty + 8 * (g % NM));
            }
        }
    } else {
        // ====================================== store warp =========================================
        uint32_t g = 0;
        for (int tile_id = blockIdx.x; tile_id < p.ntiles; tile_id += gridDim.x) {
            const int q0 = tile_id * P;
            const bool full_tile = q0 + P <= p.npts;
            for (int c = 0; c < p.nchunk; ++c, ++g) {
                const int s = g % NST;
                mbar_wait_a(a_mfull + 8 * (g % NM), (g / NM) & 1);
                mbar_wait_a(a_full + 8 * s, (g / NST) & 1);
                const unsigned char *mb = mbase + (size_t)(g % NM) * meta_bytes;
                const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(mb);
                const RowMeta *rows = reinterpret_cast<const RowMeta *>(mb + p.lay.off_row);
                const TermMeta *terms = reinterpret_cast<const TermMeta *>(mb + p.lay.off_term);
                const double *tile = tbase + (size_t)s * C::TILE_DOUBLES;
                const uint32_t a_tile = smem_u32(tile);
                bool any_slow = false;
                for (int r = lane; r < hdr.nrow; r += 32) {
                    const RowMeta rm = rows[r];
                    const TermMeta t0 = terms[rm.term_off];
                    if (full_tile && rm.nterm == 1 && t0.coef == 1.0) {
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const int code = (SET == SET_ONE) ? p.one_code : d;
                            const int sl = p.slot[code];
                            if (sl < 0) continue;
                            bulk_s2g(p.out + (size_t)sl * p.slot_stride + (size_t)rm.out_row * p.ld + q0,
                                     a_tile + (uint32_t)(((size_t)d * KC + t0.k) * P * 8), (uint32_t)(P * 8));
                        }
                    } else {
                        any_slow = true;
                    }
                }
                bulk_commit();
                if (__any_sync(0xffffffffu, any_slow)) {
                    // rows that are not plain copies: combine the terms and store, the warp striding over the points
                    for (int r = 0; r < hdr.nrow; ++r) {
                        const RowMeta rm = rows[r];
                        const TermMeta t0 = terms[rm.term_off];
                        if (full_tile && rm.nterm == 1 && t0.coef == 1.0) continue;
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const int code = (SET == SET_ONE) ? p.one_code : d;
                            const int sl = p.slot[code];
                            if (sl < 0) continue;
                            double *orow = p.out + (size_t)sl * p.slot_stride + (size_t)rm.out_row * p.ld + q0;
#pragma unroll
                            for (int j = 0; j < PT; ++j) {
                                double v = t0.coef * tile[((size_t)d * KC + t0.k) * P + j * 32 + lane];
                                for (int t = 1; t < rm.nterm; ++t) {
                                    const TermMeta tm = terms[rm.term_off + t];
                                    v = fma(tm.coef, tile[((size_t)d * KC + tm.k) * P + j * 32 + lane], v);
                                }
                                if (q0 + j * 32 + lane < p.npts) orow[j * 32 + lane] = v;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_mempty + 8 * (g % NM));      // the chunk table is no longer read
                // hand back the PREVIOUS stage: its bulk copies (one group older than the one just committed) are done
                // reading shared memory
                bulk_wait_read<1>();
                __syncwarp();
                if (g >= 1 && lane == 0) mbar_arrive_a(a_empty + 8 * ((g - 1) % NST));
            }
        }
        bulk_wait_all<0>();                            // every row has left shared memory and is written
    }
}

}  // namespace okb
