// okb_tile_kernel.cuh -- phase-serial fused kernel (all warps alternate AO generation and store /
// contraction).  Since the warp-specialised DMMA kernel (okb_ws.cuh) took over SINK_MO and SINK_RHO it
// serves the HBM-store-bound SINK_AO requests (calc_ao, cy_core.aocreator).
#pragma once
#include "okb_shell.cuh"

namespace okb {

// ---- shared memory carve-up (host and device agree through this struct) -----------------------
template <int SET, int MW, int PT, int NW, int SINK>
struct Cfg {
    static constexpr int D = set_ncodes(SET);
    static constexpr int P = 32 * PT;
    static constexpr int MC = NW * MW;
    static constexpr int NT = NW * 32;
    static constexpr int TILE_DOUBLES = D * KC * P;
    static constexpr int CBUF_DOUBLES = (SINK == SINK_AO) ? 0 : KC * MC;
    static constexpr int NOUT = (SINK == SINK_RHO) ? D : 0;      // rho + (D-1) derivative sums
    static constexpr size_t OFF_BAR = 0;                          // 5 mbarriers (3 meta + 2 coef)
    static constexpr size_t OFF_XYZ = 128;
    static constexpr size_t OFF_IJK = OFF_XYZ + (size_t)3 * P * 8;      // axis indices (regular grids)
    static constexpr size_t OFF_META = OFF_IJK + (size_t)3 * P * 4;
    __host__ __device__ static constexpr size_t off_cbuf(int meta_stride) {
        return (OFF_META + (size_t)NMETA * meta_stride + 127) / 128 * 128;
    }
    __host__ __device__ static constexpr size_t off_tile(int meta_stride) {
        return off_cbuf(meta_stride) + (size_t)2 * CBUF_DOUBLES * 8;
    }
    __host__ __device__ static constexpr size_t smem_bytes(int meta_stride) {
        return off_tile(meta_stride) + (size_t)2 * TILE_DOUBLES * 8;
    }
    // the cross-warp reduction scratch of SINK_RHO aliases the AO tiles: NW*NOUT*P doubles
    static_assert(SINK != SINK_RHO || (size_t)NW * D * P <= (size_t)2 * TILE_DOUBLES, "reduction scratch");
};

// ---- the kernel -------------------------------------------------------------------------------------
// NPT: points per thread and shell in the generation phase (0 = automatic: 2 where PT is even and the set is small);
// MINB: CTAs per SM the register allocation must allow (__launch_bounds__)
template <int SET, int MW, int PT, int NW, int SINK, int NPT = 0, int MINB = 1>
__global__ void __launch_bounds__(NW * 32, MINB) okb_grid_kernel(const KParams p) {
    using C = Cfg<SET, MW, PT, NW, SINK>;
    constexpr int D = C::D, P = C::P, MC = C::MC;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);   // [0..2] meta, [3..4] coef
    double *xs = reinterpret_cast<double *>(smem + C::OFF_XYZ);
    double *ys = xs + P, *zs = ys + P;
    int *isx = reinterpret_cast<int *>(smem + C::OFF_IJK), *isy = isx + P, *isz = isy + P;
    unsigned char *mbase = smem + C::OFF_META;
    double *cbase = reinterpret_cast<double *>(smem + C::off_cbuf(p.lay.stride));
    double *tbase = reinterpret_cast<double *>(smem + C::off_tile(p.lay.stride));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < NMETA + 2; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t meta_bytes = (uint32_t)p.lay.stride;
    constexpr uint32_t cbuf_bytes = (uint32_t)C::CBUF_DOUBLES * 8u;
    uint32_t g = 0;   // running chunk sequence number (uniform across the CTA)

    auto issue_meta = [&](int c, uint32_t gc) {
        uint64_t *bar = &bars[gc % NMETA];
        mbar_expect_tx(bar, meta_bytes);
        bulk_g2s(mbase + (size_t)(gc % NMETA) * meta_bytes, p.meta + (size_t)c * meta_bytes, meta_bytes, bar);
    };
    auto issue_coef = [&](int mt, int c, uint32_t gc) {
        if (SINK == SINK_AO) return;
        uint64_t *bar = &bars[NMETA + (gc & 1)];
        mbar_expect_tx(bar, cbuf_bytes);
        bulk_g2s(cbase + (size_t)(gc & 1) * C::CBUF_DOUBLES,
                 p.cblob + ((size_t)mt * p.nchunk + c) * C::CBUF_DOUBLES, cbuf_bytes, bar);
    };
    auto phase_a = [&](uint32_t gc) {
        const unsigned char *mb = mbase + (size_t)(gc % NMETA) * meta_bytes;
        const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(mb);
        const ShellMeta *shells = reinterpret_cast<const ShellMeta *>(mb + p.lay.off_shell);
        const double2 *prims = reinterpret_cast<const double2 *>(mb + p.lay.off_prim);
        const FnMeta *fns = reinterpret_cast<const FnMeta *>(mb + p.lay.off_fn);
        const double *aux = reinterpret_cast<const double *>(mb + p.lay.off_aux);
        double *tile = tbase + (size_t)(gc & 1) * C::TILE_DOUBLES;
        // a thread evaluates NP points (32 apart) of one shell at a time: independent dependency chains
        constexpr int NP = NPT > 0 ? NPT : (PT % 2 == 0 && SET != SET_LAP && SET != SET_ALL) ? 2 : 1, PG = PT / NP;
        const int nitems = hdr.nshell * PG;
        for (int item = warp; item < nitems; item += NW) {
            const int s = item / PG, pt = (item % PG) * (32 * NP) + lane;
            const AxTab tab{p.tabx, p.taby, p.tabz, p.nx, p.ny, p.nz, isx + pt, isy + pt, isz + pt};
            gen_shell_any<SET, P, NP>(shells[s], prims, fns, aux, xs + pt, ys + pt, zs + pt, tile + pt, p.one_code,
                                      p.exact_mixed, tab);
        }
    };

    for (int tile_id = blockIdx.x; tile_id < p.ntiles; tile_id += gridDim.x) {
        const int q0 = tile_id * P;               // launch-local index of the tile's first point
        // stage the coordinates of the tile (previous tile ended with a __syncthreads)
        if (tid < P) {
            int q = q0 + tid;
            if (q >= p.npts) q = p.npts - 1;
            grid_point(p, p.p0 + q, xs[tid], ys[tid], zs[tid], isx[tid], isy[tid], isz[tid]);
        }
        double osum[C::NOUT > 0 ? C::NOUT : 1][PT];
        if (SINK == SINK_RHO) {
#pragma unroll
            for (int o = 0; o < C::NOUT; ++o)
#pragma unroll
                for (int j = 0; j < PT; ++j) osum[o][j] = 0.0;
        }

        for (int mt = 0; mt < p.n_mtile; ++mt) {
            double acc[MW][PT][D];
            if (SINK != SINK_AO) {
#pragma unroll
                for (int i = 0; i < MW; ++i)
#pragma unroll
                    for (int j = 0; j < PT; ++j)
#pragma unroll
                        for (int d = 0; d < D; ++d) acc[i][j][d] = 0.0;
            }
            const uint32_t g0 = g;
            if (tid == 0) {
                issue_meta(0, g0);
                if (p.nchunk > 1) issue_meta(1, g0 + 1);
                issue_coef(mt, 0, g0);
            }
            __syncthreads();                       // coordinates staged
            mbar_wait(&bars[g0 % NMETA], (g0 / NMETA) & 1);
            phase_a(g0);

            for (int c = 0; c < p.nchunk; ++c) {
                const uint32_t gc = g0 + c;
                __syncthreads();                   // A(c) done, B(c-1) done
                if (tid == 0) {
                    if (c + 1 < p.nchunk) issue_coef(mt, c + 1, gc + 1);
                    if (c + 2 < p.nchunk) issue_meta(c + 2, gc + 2);
                }
                const unsigned char *mb = mbase + (size_t)(gc % NMETA) * meta_bytes;
                const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(mb);
                const double *tile = tbase + (size_t)(gc & 1) * C::TILE_DOUBLES;

                if (SINK == SINK_AO) {
                    // ---- phase B (store): optional cart->sph rows, coalesced row stores ----
                    const RowMeta *rows = reinterpret_cast<const RowMeta *>(mb + p.lay.off_row);
                    const TermMeta *terms = reinterpret_cast<const TermMeta *>(mb + p.lay.off_term);
                    for (int r = warp; r < hdr.nrow; r += NW) {
                        const RowMeta rm = rows[r];
                        // rows are warp-uniform: the common single-term row (a Cartesian function, a spherical
                        // row the generator already built, or an s/p function) is a scaled copy with PT*D
                        // independent load -> multiply -> store chains; multi-term rows keep PT accumulators
                        const TermMeta t0 = terms[rm.term_off];
#pragma unroll
                        for (int d = 0; d < D; ++d) {
                            const int code = (SET == SET_ONE) ? p.one_code : d;
                            const int sl = p.slot[code];
                            if (sl < 0) continue;
                            double *orow = p.out + (size_t)sl * p.slot_stride + (size_t)rm.out_row * p.ld + q0;
                            double v[PT];
#pragma unroll
                            for (int j = 0; j < PT; ++j) v[j] = t0.coef * tile[((size_t)d * KC + t0.k) * P + j * 32 + lane];
                            for (int t = 1; t < rm.nterm; ++t) {
                                const TermMeta tm = terms[rm.term_off + t];
#pragma unroll
                                for (int j = 0; j < PT; ++j)
                                    v[j] = fma(tm.coef, tile[((size_t)d * KC + tm.k) * P + j * 32 + lane], v[j]);
                            }
#pragma unroll
                            for (int j = 0; j < PT; ++j)
                                if (q0 + j * 32 + lane < p.npts) orow[j * 32 + lane] = v[j];
                        }
                    }
                } else {
                    // ---- phase B (contract): acc[i][j][d] += C[k][w*MW+i] * ao[d][k][pt_j] ----
                    mbar_wait(&bars[NMETA + (gc & 1)], (gc >> 1) & 1);
                    const double *cs = cbase + (size_t)(gc & 1) * C::CBUF_DOUBLES + warp * MW;
                    const double *tl = tile + lane;
                    const int nfn = hdr.nfn;
#pragma unroll 2
                    for (int k = 0; k < nfn; ++k) {
                        double cv[MW];
                        if (MW % 2 == 0) {
#pragma unroll
                            for (int i = 0; i < MW; i += 2) {
                                const double2 c2 = *reinterpret_cast<const double2 *>(cs + (size_t)k * MC + i);
                                cv[i] = c2.x;
                                cv[i + 1] = c2.y;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < MW; ++i) cv[i] = cs[(size_t)k * MC + i];
                        }
#pragma unroll
                        for (int d = 0; d < D; ++d)
#pragma unroll
                            for (int j = 0; j < PT; ++j) {
                                const double a = tl[((size_t)d * KC + k) * P + j * 32];
#pragma unroll
                                for (int i = 0; i < MW; ++i) acc[i][j][d] = fma(cv[i], a, acc[i][j][d]);
                            }
                    }
                }
                if (c + 1 < p.nchunk) {
                    mbar_wait(&bars[(gc + 1) % NMETA], ((gc + 1) / NMETA) & 1);
                    phase_a(gc + 1);
                }
            }
            g = g0 + p.nchunk;
            __syncthreads();                       // every buffer is free again

            // ---- per-MO-tile epilogues -----------------------------------------------------------
            if (SINK == SINK_MO) {
#pragma unroll
                for (int i = 0; i < MW; ++i) {
                    const int mo = mt * MC + warp * MW + i;
                    if (mo >= p.n_mo) continue;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const int code = (SET == SET_ONE) ? p.one_code : d;
                        const int sl = p.slot[code];
                        if (sl < 0) continue;
                        double *orow = p.out + (size_t)sl * p.slot_stride + (size_t)mo * p.ld + q0;
#pragma unroll
                        for (int j = 0; j < PT; ++j) {
                            const int pt = j * 32 + lane;
                            if (q0 + pt < p.npts) orow[pt] = acc[i][j][d];
                        }
                    }
                }
            }
            if (SINK == SINK_RHO) {
#pragma unroll
                for (int i = 0; i < MW; ++i) {
                    const int mo = mt * MC + warp * MW + i;
                    const double oc = p.occ[mo];          // zero for padding MOs
                    double nrm = 0.0;
#pragma unroll
                    for (int j = 0; j < PT; ++j) {
                        const double phi = acc[i][j][0];
                        const bool valid = (q0 + j * 32 + lane) < p.npts;
                        if (valid) nrm += phi * phi;
                        osum[0][j] += oc * (phi * phi);
                        if (D >= 4) {
                            const double o2 = oc * 2.0;
                            osum[1][j] += o2 * (acc[i][j][1] * phi);
                            osum[2][j] += o2 * (acc[i][j][2] * phi);
                            osum[3][j] += o2 * (acc[i][j][3] * phi);
                            if (D >= 7) {
                                osum[4][j] += o2 * (acc[i][j][4] * phi + acc[i][j][1] * acc[i][j][1]);
                                osum[5][j] += o2 * (acc[i][j][5] * phi + acc[i][j][2] * acc[i][j][2]);
                                osum[6][j] += o2 * (acc[i][j][6] * phi + acc[i][j][3] * acc[i][j][3]);
                            }
                            if (D >= 10) {
                                osum[7][j] += o2 * (acc[i][j][7] * phi + acc[i][j][1] * acc[i][j][2]);
                                osum[8][j] += o2 * (acc[i][j][8] * phi + acc[i][j][1] * acc[i][j][3]);
                                osum[9][j] += o2 * (acc[i][j][9] * phi + acc[i][j][2] * acc[i][j][3]);
                            }
                        }
                    }
                    if (p.mo_norm != nullptr && mo < p.n_mo) {
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, off);
                        if (lane == 0) atomicAdd(p.mo_norm + mo, nrm);
                    }
                }
            }
        }   // mt

        if (SINK == SINK_RHO) {
            // cross-warp reduction through shared memory (aliases the AO tiles; all warps passed
            // the post-loop __syncthreads, so the tiles are dead)
            double *red = tbase;                   // [NW][NOUT][P]
#pragma unroll
            for (int o = 0; o < C::NOUT; ++o)
#pragma unroll
                for (int j = 0; j < PT; ++j) red[((size_t)warp * C::NOUT + o) * P + j * 32 + lane] = osum[o][j];
            __syncthreads();
            for (int e = tid; e < C::NOUT * P; e += C::NT) {
                const int o = e / P, pt = e - o * P;
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += red[((size_t)w * C::NOUT + o) * P + pt];
                if (q0 + pt < p.npts) {
                    if (o == 0) {
                        if (p.rho != nullptr) p.rho[q0 + pt] = s;
                    } else {
                        const int sl = p.slot[o];
                        if (sl >= 0) p.delta[(size_t)sl * p.ld + q0 + pt] = s;
                    }
                }
            }
            __syncthreads();                       // scratch free before the next tile's phase A
        }
    }
}

}  // namespace okb
