// okb_ws.cuh -- warp-specialised fused kernel for SINK_MO / SINK_RHO: AO generation and the FP64
// contraction run CONCURRENTLY on different warps of one persistent CTA per SM.
//
// Why (measured, profiles/r01_*): in the phase-serial kernel the FP64 pipe was 56% active -- all warps
// did the latency-bound AO generation together, then the pipe-bound contraction together.  With
// DFMA consumers the pipe reached 67%: DFMA with varying operands issues at 2.44 cycles/instruction
// (30 TFLOP/s) and the register tile needed an LDS per 4.6 DFMA.  scripts/fp64_micro.cu shows DMMA
// (mma.sync.m8n8k4.f64) sustaining 37.1 TFLOP/s = 16.0 cycles per instruction per SM sub-partition,
// 8x fewer issue slots than DFMA on the same 64 lanes/SM datapath (DMMA wins arbitration against
// DFMA).  So the contraction uses DMMA; the FP64 tensor path of sm_100a is mma.sync (tcgen05 has no
// f64 kind).
//
//   producers  NPW warps (registers cut by setmaxnreg): evaluate (shell, 32 points) items of chunk g
//              into stage g % NST of the AO tile ring, zero the k-padding rows, issue the bulk-async
//              (TMA) copies of the chunk tables and of the coefficient tile of the stage.
//   consumers  NCW = WM x WN warps (registers raised by setmaxnreg): warp (wm, wn) owns the MO blocks
//              [wm*AM, min((wm+1)*AM, MB)) x point blocks [wn*BN, (wn+1)*BN) (blocks of 8), an
//              (8 AM) x (8 BN) x D register tile of 2*AM*BN*D doubles per thread.  Per k-step of 4:
//              ceil(AM/2) LDS.128 (MO blocks are stored in pairs, see mo_blob in okb200.cu) + D*BN LDS.64, all
//              conflict free, feed AM*BN*D DMMAs.  Warps wm = 0..WM-1 of one wn share
//              an SM sub-partition (warp % 4 == wn for WN == 4), so while one waits on a barrier or a
//              fragment load the other keeps the pipe busy.
//   hand-over  mbarriers: full[s] (one arrival per producer thread + the TMA bytes of the coefficient
//              tile), empty[s] (one arrival per consumer warp).  No CTA-wide barrier in the main loop.
//
// mma.m8n8k4 operand mapping (lane T):  A[m][k] = C'[mo0 + T/4][k0 + T%4]   (row-major 8x4)
//                                       B[k][n] = ao[d][k0 + T%4][pt0 + T/4] (col-major 4x8)
//                                       C[m][n] : m = T/4, n = 2*(T%4) + {0,1}
// Shared-memory rows are padded to a stride = 4 (mod 16) doubles so that both fragment loads touch
// 32 distinct 8-byte words per half-warp (2 wavefronts per LDS.64, the minimum; the 16-byte A loads touch every
// 16-byte slot of a 128-byte line exactly twice per half-warp).
#pragma once
#include <type_traits>

#include "okb_shell.cuh"

namespace okb {

#ifndef OKB_CSLACK
#define OKB_CSLACK 56
#endif
__host__ __device__ constexpr int pad_stride(int n) { return n + ((4 - (n % 16)) + 16) % 16; }

// MB: MO blocks of 8 per CTA tile; WM x WN consumer warps; BN point blocks per warp; NPW producer warps;
// REM: remainder orbitals of the tile beyond its MB blocks (MO tile = 8*MB + REM orbitals).  They are contracted by the
// PRODUCER warps with DFMAs on the AO values they have just written (2 DFMA per AO value and remainder orbital, partial
// sums per producer warp handed to the consumers once per pass) instead of costing a whole, mostly empty DMMA block:
// 82 occupied MOs = 10 blocks + 2 run 40 instead of 44 DMMAs per k-step.
template <int SET, int MB, int BN, int WM, int WN, int NPW, int NST, int SINK, int REM = 0>
struct WsCfg {
    static constexpr int D = set_ncodes(SET);
    static constexpr int NCW = WM * WN;
    static constexpr int AM = (MB + WM - 1) / WM;              // MO blocks per consumer warp
    static constexpr int P = 8 * BN * WN;                      // points per CTA tile
    static constexpr int PT = P / 32;
    static constexpr int MC = 8 * MB;                          // MOs per CTA tile contracted by the consumers (DMMA)
    static constexpr int MCT = MC + REM;                       // MOs per CTA tile
    static constexpr int PS = pad_stride(P);                   // AO tile row stride (doubles)
    static constexpr int CS = pad_stride(MC);                  // coefficient tile row stride (doubles)
    static constexpr int NT = (NCW + NPW) * 32;
    // ROT: the consumer walks TWO k-steps (8 rows) per loop iteration with the A fragments in a small ring of register
    // buffers and the B fragments double buffered (see the consumer loop).
#ifdef OKB_NO_ROT
    static constexpr bool ROT = false;
#else
    static constexpr bool ROT = (WM == 1 && NST >= 3 && D <= 4 && (((MB + WM - 1) / WM + 1) / 2) >= 2);
#endif
    // IL: derivative sets interleaved in pairs in the AO tile (TileLay<-RS>): one 16-byte load fetches the B fragments of
    // two sets, one 16-byte store of a producer writes two sets.  RS = row stride; RS/2 = 2 (mod 8) makes the 16-byte
    // fragment loads of a quarter warp (2 point rows x 4 k rows) hit all 8 16-byte slots of a 128-byte line.
#ifdef OKB_NO_IL
    static constexpr bool IL = false;
#else
    static constexpr bool IL = ROT && D >= 2;
#endif
    static constexpr int DP = (D + 1) / 2;                     // set pairs
    static constexpr int RS = 2 * (P + ((2 - (P % 8)) + 8) % 8);
    static constexpr int STR = IL ? -RS : PS;                  // TileLay parameter of the generators
    static constexpr int TILE_DOUBLES = IL ? DP * KC * RS : D * KC * PS;
    static constexpr int CBUF_DOUBLES = KC * CS;
    static constexpr int NOUT = (SINK == SINK_RHO) ? D : 0;
    // register budget: NT threads are launched with 65536/NT registers each; after setmaxnreg
    // NCW*32*CREG + NPW*32*PREG must not exceed what the launch reserved
    static constexpr int LAUNCH_REGS = (65536 / NT) / 8 * 8;
    static constexpr int ACC_REGS = 4 * AM * BN * D;
    // registers of a consumer thread beyond its accumulators: A/B fragments, addresses, loop state.  One consumer
    // warpgroup with a full tile gets 56 (ptxas otherwise recycles a B-fragment register as address register and
    // refills it right in front of its use); big tiles on two warpgroups leave the producers > 100 registers.
    static constexpr int CREG_SLACK = (NCW == 8 && ACC_REGS > 160) ? 32 : (NCW == 4 && ACC_REGS >= 160) ? OKB_CSLACK : 40;
    static constexpr int CREG_WANT = ((ACC_REGS + CREG_SLACK + 7) / 8 * 8 > 248) ? 248 : (ACC_REGS + CREG_SLACK + 7) / 8 * 8;
    static constexpr int PREG_MAX = ((NT * LAUNCH_REGS - NCW * 32 * CREG_WANT) / (NPW * 32)) / 8 * 8;
    static constexpr int PREG_CAP = LAUNCH_REGS < 152 ? LAUNCH_REGS : 152;   // setmaxnreg.dec may only lower
    static constexpr int PREG = PREG_MAX > PREG_CAP ? PREG_CAP : PREG_MAX;
    static constexpr int CREG = ((NT * LAUNCH_REGS - NPW * 32 * PREG) / (NCW * 32)) / 8 * 8 > 248
                                    ? 248
                                    : ((NT * LAUNCH_REGS - NPW * 32 * PREG) / (NCW * 32)) / 8 * 8;
    static constexpr int NM = 4;                               // chunk-table ring slots
    static constexpr int LAG = 2;                              // a slot is refilled LAG chunks after its chunk
    static constexpr size_t OFF_BAR = 0;                       // full[NST] empty[NST] mfull[NM] mempty[NM]
    static constexpr size_t OFF_NFN = 256;                     // int nfn[NST]
    static constexpr size_t OFF_XYZ = 384;
    static constexpr size_t OFF_IJK = OFF_XYZ + (size_t)3 * P * 8;     // axis indices of the tile's points (regular grids)
    static constexpr size_t OFF_RED = OFF_IJK + (size_t)3 * P * 4;
    // remainder orbitals: partial sums [NPW][REM][D][P] of one pass, then (behind the chunk tables) a ring of NM
    // coefficient slots [KC][REM]
    static constexpr size_t OFF_PART = OFF_RED + (size_t)(NOUT > 0 ? WM * NOUT * P * 8 : 0);
    static constexpr size_t OFF_META = OFF_PART + (size_t)NPW * REM * D * P * 8;
    static constexpr uint32_t CREM_BYTES = (uint32_t)KC * REM * 8u;
    __host__ __device__ static constexpr size_t off_crem(int meta_stride) { return OFF_META + (size_t)NM * meta_stride; }
    __host__ __device__ static constexpr size_t off_cbuf(int meta_stride) {
        return (off_crem(meta_stride) + (size_t)NM * CREM_BYTES + 127) / 128 * 128;
    }
    __host__ __device__ static constexpr size_t off_tile(int meta_stride) {
        return off_cbuf(meta_stride) + (size_t)NST * CBUF_DOUBLES * 8;
    }
    __host__ __device__ static constexpr size_t smem_bytes(int meta_stride) {
        return off_tile(meta_stride) + (size_t)NST * TILE_DOUBLES * 8;
    }
    static_assert(2 * NST + 2 * NM + 1 <= 32, "barrier area");
    static_assert(REM == 0 || (REM == 2 && WM == 1 && SINK != SINK_AO), "remainder orbitals: pairs, one consumer warp row");
    static_assert(NCW % 4 == 0 && NPW % 4 == 0, "whole warpgroups (setmaxnreg)");
    // MO blocks are fetched in pairs (one 16-byte load) when every warp row starts on an even block
    static constexpr bool PAIRED = (WM == 1 || AM % 2 == 0);
    // NPAIR block pairs per k-step, NP2 pair slots per double step, NBUF ring buffers (a divisor of NP2)
    static constexpr int NPAIR = (AM + 1) / 2;
    static constexpr int NP2 = 2 * NPAIR;
#ifdef OKB_NBUF
    static constexpr int NBUF = OKB_NBUF, PD = OKB_PD;          // A/B builds
#else
    static constexpr int NBUF = (NP2 % 3 == 0) ? 3 : 2;         // ring buffers (a divisor of NP2)
    static constexpr int PD = 1;                                // the pair of slot q + 1 is fetched during slot q
#endif
    static_assert(NP2 % NBUF == 0 && PD >= 1 && PD < NP2, "ring");
    static constexpr int KSTEP = ROT ? 8 : 4;                  // rows the tile is padded to
    static_assert(P % 32 == 0, "whole warps of points for the producers");
    static_assert(PREG >= 56 && CREG >= LAUNCH_REGS && PREG <= LAUNCH_REGS, "register split");
};

// Remainder-orbital hook of the AO generators (okb_shell.cuh): the producer thread multiplies every AO value it writes
// by the two remainder orbitals' coefficients of that row (one warp-uniform 16-byte load) and keeps the partial sums.
template <int D, int NP>
struct RemAcc {
    static constexpr bool on = true;
    const double2 *cr;                       // [KC] coefficient pairs of the chunk's rows (shared memory)
    double (&acc)[NP][2][D];                 // [point of the thread][orbital][set]
    __device__ __forceinline__ void row(int k, int q, const double (&v)[D]) const {
        const double2 c = cr[k];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            acc[q][0][d] = fma(c.x, v[d], acc[q][0][d]);
            acc[q][1][d] = fma(c.y, v[d], acc[q][1][d]);
        }
    }
};

template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// barrier helpers on precomputed shared-space addresses (no generic->shared conversion in the loops)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_only_a(uint32_t bar, uint32_t bytes) {   // no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "OKB_WAITA_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra OKB_DONEA_%=;\n"
        "bra OKB_WAITA_%=;\n"
        "OKB_DONEA_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// volatile shared load: keeps its place in the hand-scheduled DMMA stream
__device__ __forceinline__ double lds64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds128(uint32_t addr, double &v0, double &v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr));
}
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// one non-blocking probe of a barrier phase (the consumers look at the NEXT chunk's barrier a chunk early, so that the
// blocking wait -- and the ~40 cycles of its predicate -- only happens when the producers are not ahead)
__device__ __forceinline__ bool mbar_test_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

template <int SET, int MB, int BN, int WM, int WN, int NPW, int NST, int SINK, int REM = 0>
__global__ void __launch_bounds__((WM * WN + NPW) * 32, 1) okb_ws_kernel(const KParams p) {
    using C = WsCfg<SET, MB, BN, WM, WN, NPW, NST, SINK, REM>;
    constexpr int D = C::D, P = C::P, PT = C::PT, PS = C::PS, CS = C::CS, MC = C::MC, MCT = C::MCT, NCW = C::NCW, AM = C::AM;
    constexpr int NPT = NPW * 32;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    constexpr int NM = C::NM, LAG = C::LAG;
    const uint32_t a_full = sbase + (uint32_t)C::OFF_BAR, a_empty = a_full + 8 * NST, a_mfull = a_empty + 8 * NST,
                   a_mempty = a_mfull + 8 * NM, a_rfree = a_mempty + 8 * NM;
    int *nfn_s = reinterpret_cast<int *>(smem + C::OFF_NFN);
    double *xs = reinterpret_cast<double *>(smem + C::OFF_XYZ);
    double *ys = xs + P, *zs = ys + P;
    int *isx = reinterpret_cast<int *>(smem + C::OFF_IJK), *isy = isx + P, *isz = isy + P;
    double *red = reinterpret_cast<double *>(smem + C::OFF_RED);
    double *part = reinterpret_cast<double *>(smem + C::OFF_PART);      // [NPW][REM][D][P]
    unsigned char *mbase = smem + C::OFF_META;
    double *crem_s = reinterpret_cast<double *>(smem + C::off_crem(p.lay.stride));   // [NM][KC][REM]
    double *cbase = reinterpret_cast<double *>(smem + C::off_cbuf(p.lay.stride));
    double *tbase = reinterpret_cast<double *>(smem + C::off_tile(p.lay.stride));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
        for (int i = 0; i < NST; ++i) {
            mbar_init(&bars[i], NPT);                 // + TMA bytes of the coefficient tile
            mbar_init(&bars[NST + i], NCW);
        }
        for (int i = 0; i < NM; ++i) {
            mbar_init(&bars[2 * NST + i], 1);
            mbar_init(&bars[2 * NST + NM + i], NPW);  // one arrival per producer warp that left the chunk
        }
        mbar_init(&bars[2 * NST + 2 * NM], NCW);      // rfree: the consumers have read the remainder partial sums
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t meta_bytes = (uint32_t)p.lay.stride;
    constexpr uint32_t cbuf_bytes = (uint32_t)C::CBUF_DOUBLES * 8u;
    constexpr uint32_t tile_bytes = (uint32_t)C::TILE_DOUBLES * 8u;
    const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)p.n_mtile * (uint32_t)p.nchunk;

    if (warp >= NCW) {
        // ====================================== producers ==========================================
        reg_dec<C::PREG>();
        const int ptid = tid - NCW * 32, pwarp = warp - NCW;
        const uint32_t a_meta = smem_u32(mbase), a_cbuf = smem_u32(cbase), a_crem = smem_u32(crem_s);
        auto issue_meta = [&](uint32_t gc) {
            const uint32_t bar = a_mfull + 8 * (gc % NM);
            const uint32_t c = gc % (uint32_t)p.nchunk;
            mbar_arrive_expect_tx_a(bar, meta_bytes + C::CREM_BYTES);
            bulk_g2s_a(a_meta + (gc % NM) * meta_bytes, p.meta + (size_t)c * meta_bytes, meta_bytes, bar);
            if constexpr (REM > 0) {                  // the remainder orbitals' coefficients of this (MO tile, chunk)
                const uint32_t mt = (gc / (uint32_t)p.nchunk) % (uint32_t)p.n_mtile;
                bulk_g2s_a(a_crem + (gc % NM) * C::CREM_BYTES, p.crem + ((size_t)mt * p.nchunk + c) * (KC * REM),
                           C::CREM_BYTES, bar);
            }
        };
        if (ptid == 0)
            for (uint32_t i = 0; i < NM && i < total; ++i) issue_meta(i);
        uint32_t g = 0, npass = 0;
        for (int tile_id = blockIdx.x; tile_id < p.ntiles; tile_id += gridDim.x) {
            const int q0 = tile_id * P;
            named_bar(1, NPT);                        // previous tile's items no longer read xs/ys/zs
            for (int e = ptid; e < P; e += NPT) {
                int q = q0 + e;
                if (q >= p.npts) q = p.npts - 1;
                grid_point(p, p.p0 + q, xs[e], ys[e], zs[e], isx[e], isy[e], isz[e]);
            }
            named_bar(1, NPT);
            // a thread evaluates NP points (32 apart) of one shell at a time (independent dependency
            // chains); warps take items round-robin, rotated per chunk so that the same warp is not
            // always the one with the extra item
            constexpr int NP = (PT % 2 == 0 && SET != SET_LAP && SET != SET_ALL) ? 2 : 1, PG = PT / NP;
            for (int mt = 0; mt < p.n_mtile; ++mt, ++npass) {
                // remainder orbitals: this thread's partial sums over the shells its warp evaluates in this pass,
                // per point group / point / orbital / derivative set
                double racc[PG][NP][REM > 0 ? REM : 1][D];
                if constexpr (REM > 0) {
#pragma unroll
                    for (int a = 0; a < PG; ++a)
#pragma unroll
                        for (int q = 0; q < NP; ++q)
#pragma unroll
                            for (int r = 0; r < REM; ++r)
#pragma unroll
                                for (int d = 0; d < D; ++d) racc[a][q][r][d] = 0.0;
                }
                for (int c = 0; c < p.nchunk; ++c, ++g) {
                    const int s = g % NST;
                    // The producer warps are NOT synchronised per chunk: a warp that is done with its items
                    // of chunk g goes on to chunk g+1 (at most NST-1 chunks ahead of the slowest one, bounded
                    // by the stage ring), so uneven item costs average out over several chunks.  The
                    // chunk-table slot of chunk g-LAG is refilled once every warp has left that chunk.
                    if (ptid == 0 && g >= LAG && g - LAG + NM < total) {
                        const uint32_t h = g - LAG;
                        mbar_wait_a(a_mempty + 8 * (h % NM), (h / NM) & 1);
                        issue_meta(h + NM);
                    }
                    mbar_wait_a(a_mfull + 8 * (g % NM), (g / NM) & 1);
                    mbar_wait_a(a_empty + 8 * s, ((g / NST) & 1) ^ 1);   // consumers released the stage
                    if (ptid == 0) {
                        mbar_expect_tx_only_a(a_full + 8 * s, cbuf_bytes);
                        bulk_g2s_a(a_cbuf + s * cbuf_bytes, p.cblob + ((size_t)mt * p.nchunk + c) * C::CBUF_DOUBLES,
                                   cbuf_bytes, a_full + 8 * s);
                    }
                    const unsigned char *mb = mbase + (size_t)(g % NM) * meta_bytes;
                    const ChunkHdr hdr = *reinterpret_cast<const ChunkHdr *>(mb);
                    const ShellMeta *shells = reinterpret_cast<const ShellMeta *>(mb + p.lay.off_shell);
                    const double2 *prims = reinterpret_cast<const double2 *>(mb + p.lay.off_prim);
                    const FnMeta *fns = reinterpret_cast<const FnMeta *>(mb + p.lay.off_fn);
                    const double *aux = reinterpret_cast<const double *>(mb + p.lay.off_aux);
                    double *tile = tbase + (size_t)s * C::TILE_DOUBLES;
                    // the host sorts the shells of a chunk by descending cost; items are dealt to the warps in
                    // boustrophedon order (0..NPW-1, NPW-1..0, ...), rotated per chunk
                    // (rotated by the chunk's index IN THE PASS, not by the CTA's running chunk counter: which warp
                    // evaluates which shell of a tile must not depend on how a request is cut into launches -- the remainder
                    // orbitals' partial sums are grouped by warp, and results stay bit-identical across slabs and shards)
                    const int nitems = hdr.nshell * PG, wrot = (pwarp + c) % NPW;
                    for (int r = 0; r * NPW < nitems; ++r) {
                        const int item = r * NPW + ((r & 1) ? NPW - 1 - wrot : wrot);
                        if (item >= nitems) continue;
                        const int sh = item / PG, pt = (item % PG) * (32 * NP) + lane;
                        const AxTab tab{p.tabx, p.taby, p.tabz, p.nx, p.ny, p.nz, isx + pt, isy + pt, isz + pt};
                        if constexpr (REM > 0) {
                            // the generators hand every row they write to the remainder orbitals' partial sums
                            const double2 *cr = reinterpret_cast<const double2 *>(crem_s + (size_t)(g % NM) * (KC * REM));
                            static_for<0, PG>([&](auto ac) {
                                constexpr int a = decltype(ac)::value;
                                if (PG > 1 && (item % PG) != a) return;          // warp-uniform
                                const RemAcc<D, NP> ra{cr, racc[a]};
                                gen_shell_any<SET, C::STR, NP>(shells[sh], prims, fns, aux, xs + pt, ys + pt, zs + pt,
                                                               tile + (C::IL ? 2 : 1) * pt, p.one_code, p.exact_mixed, tab, ra);
                            });
                        } else {
                            gen_shell_any<SET, C::STR, NP>(shells[sh], prims, fns, aux, xs + pt, ys + pt, zs + pt,
                                                           tile + (C::IL ? 2 : 1) * pt, p.one_code, p.exact_mixed, tab);
                        }
                    }
                    // zero the rows that pad nfn up to the k-step of the consumer loop (coefficients there are 0,
                    // but stale shared memory could hold NaN/Inf bit patterns)
                    const int kpad = (hdr.nfn + C::KSTEP - 1) & ~(C::KSTEP - 1);
                    for (int e = ptid; e < (kpad - hdr.nfn) * D * P; e += NPT) {
                        const int pt = e % P, r = e / P, d = r % D, k = hdr.nfn + r / D;
                        tile[TileLay<C::STR>::off(d, k) + (C::IL ? 2 : 1) * pt] = 0.0;
                    }
                    if (ptid == 0) nfn_s[s] = kpad;
                    if constexpr (REM > 0) {
                        if (c == p.nchunk - 1) {
                            // last chunk of the pass: hand the partial sums over (the arrival on full[s] below
                            // publishes them); the consumers must have read those of the previous pass
                            mbar_wait_a(a_rfree, (npass & 1) ^ 1);
#pragma unroll
                            for (int a = 0; a < PG; ++a)
#pragma unroll
                                for (int q = 0; q < NP; ++q)
#pragma unroll
                                    for (int r = 0; r < REM; ++r)
#pragma unroll
                                        for (int d = 0; d < D; ++d)
                                            part[(((size_t)pwarp * REM + r) * D + d) * P + a * (32 * NP) + 32 * q + lane] =
                                                racc[a][q][r][d];
                        }
                    }
                    mbar_arrive_a(a_full + 8 * s);                       // release: tile + nfn (+ partial sums) visible
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(a_mempty + 8 * (g % NM));   // this warp no longer reads the chunk table
                }
            }
        }
    } else {
        // ====================================== consumers ==========================================
        reg_inc<C::CREG>();
        const int wm = warp / WN, wn = warp % WN;
        const int tr = lane >> 2, tc = lane & 3;            // fragment row (T/4) and column (T%4)
        const int mo_w = wm * AM * 8;                       // first MO of this warp inside the tile
        const int pt_w = wn * BN * 8;                       // first point of this warp inside the tile
        // number of MO blocks this warp really owns (the last warp row may own fewer: MB % WM != 0)
        const int nblk = (MB - wm * AM) < AM ? (MB - wm * AM) : AM;
        // Stage bookkeeping carried from chunk to chunk: stage index, phase parity and this lane's fragment
        // addresses in the stage.  (Recomputing them from the chunk counter per chunk cost ~140 cycles per chunk:
        // ptxas rematerialised the lane offsets from %tid every time -- profiles/r02_ws_grad_loop.txt.)
        // coefficient rows hold the MO blocks in pairs (okb200.cu: mo_blob): block b, row r at (b/2)*16 + 2r + (b&1);
        // an odd first block (WM > 1 with odd AM) starts in the second slot of its pair
        uint32_t st = 0, ph = 0;
        uint32_t a_st = smem_u32(cbase + (size_t)tc * CS + ((mo_w >> 3) >> 1) * 16 + ((mo_w >> 3) & 1) + 2 * tr);
        uint32_t b_st = C::IL ? smem_u32(tbase + (size_t)tc * C::RS + 2 * (pt_w + tr)) : smem_u32(tbase + (size_t)tc * PS + pt_w + tr);
        auto advance = [&](uint32_t &s_, uint32_t &h_, uint32_t &a_, uint32_t &b_) {
            if (s_ + 1 == NST) {
                s_ = 0; h_ ^= 1u;
                a_ -= (uint32_t)(NST - 1) * cbuf_bytes; b_ -= (uint32_t)(NST - 1) * tile_bytes;
            } else {
                ++s_; a_ += cbuf_bytes; b_ += tile_bytes;
            }
        };
        uint32_t npass = 0;
        for (int tile_id = blockIdx.x; tile_id < p.ntiles; tile_id += gridDim.x) {
            const int q0 = tile_id * P;
            double osum[C::NOUT > 0 ? C::NOUT : 1][BN][2];
            if (SINK == SINK_RHO) {
#pragma unroll
                for (int o = 0; o < C::NOUT; ++o)
#pragma unroll
                    for (int ib = 0; ib < BN; ++ib) osum[o][ib][0] = osum[o][ib][1] = 0.0;
            }
            for (int mt = 0; mt < p.n_mtile; ++mt, ++npass) {
                double acc[AM][BN][D][2];
#pragma unroll
                for (int ia = 0; ia < AM; ++ia)
#pragma unroll
                    for (int ib = 0; ib < BN; ++ib)
#pragma unroll
                        for (int d = 0; d < D; ++d) acc[ia][ib][d][0] = acc[ia][ib][d][1] = 0.0;
                // ---- contraction over the chunks of this MO tile ------------------------------------------------
                // One software pipeline over ALL k-steps of the pass: the fragments of the next step are fetched
                // during the current one, across chunk boundaries too (the last step of chunk c prefetches the
                // first fragments of chunk c+1 from the next stage of the ring), so the DMMA stream is only
                // interrupted once per MO-tile pass.  ncu had shown ~8% of the consumer's time in the per-chunk
                // prologue (barrier wait, pointer set-up, exposed LDS latency of the first 15 fragments).
                // Register use is explicit: A fragments afr[AM], ONE set of B fragments bfr[D][BN]; a fragment
                // register is refilled right after the last DMMA of the step that reads it.
                double afr[AM], bfr[D][BN];
                // One k-step = NB*AM DMMAs (NB = D*BN B fragments outermost, AM MO blocks innermost) on the fragments
                // in registers; a fragment register is refilled behind the last DMMA of the step that reads it:
                // B[j] behind pass j (needed again (NB-1)*AM DMMAs later), A[ia] during the last pass (needed again
                // AM DMMAs later) -- both far above the LDS latency.
                // Known residue (ncu, profiles/r01_ws_grad_refill.txt): a DMMA.8x8x4 reads its source registers
                // only when it enters the pipe, so the LDS that overwrites one of them right behind it waits a few
                // cycles, and with it the in-order warp: ~10% of the consumer's samples are short-scoreboard stalls
                // on these 15 refills per step.  ptxas places a refill directly behind the last reader whatever the
                // source order says (also when the value travels through a temporary and an opaque xor), so with ONE
                // consumer warp per sub-partition this is the floor; two consumer warps per sub-partition hide it.
                // Two timing experiments of round 1 (profiles/README.md): making each refill data dependent on the accumulator
                // of its last reader (so that ptxas schedules it by the DMMA result latency) makes ptxas cluster all refills at
                // the end of the step: 206.6 ms instead of 192.3; dropping the 11 A refills altogether (wrong results): 182.2 ms,
                // i.e. all A refills together cost 5 %, the 15 refills about 8 % -- an upper bound for any refill scheme.
                constexpr int NB = D * BN;
                // unpaired fallback: byte offset of the warp's block ia relative to the slot of its first block
                auto aoff = [&](int ia) -> uint32_t {
                    const int fb = mo_w >> 3, b = fb + ia;
                    return (uint32_t)(((b >> 1) * 16 + (b & 1)) - ((fb >> 1) * 16 + (fb & 1))) * 8u;
                };
                auto step = [&](const uint32_t na, const uint32_t nb) {
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
#pragma unroll
                        for (int ia = 0; ia < AM; ++ia) {
                            const bool owned = !(MB % WM != 0 && ia == AM - 1 && ia >= nblk);   // warp-uniform
                            if (owned)
                                dmma_m8n8k4(acc[ia][j % BN][j / BN][0], acc[ia][j % BN][j / BN][1], afr[ia], bfr[j / BN][j % BN]);
                            if (j == NB - 1) {               // A fragments of the next step: one 16-byte load per block pair
                                if (C::PAIRED) {
                                    if (ia & 1) lds128(na + (uint32_t)((ia >> 1) * 16) * 8u, afr[ia - 1], afr[ia]);
                                    else if (ia == AM - 1) afr[ia] = lds64(na + (uint32_t)((ia >> 1) * 16) * 8u);
                                } else {
                                    afr[ia] = lds64(na + aoff(ia));
                                }
                            }
                        }
                        bfr[j / BN][j % BN] = lds64(nb + (uint32_t)((j / BN) * KC * PS + (j % BN) * 8) * 8u);
                    }
                };
                auto load_all = [&](const uint32_t ca, const uint32_t cb) {     // fragments straight from shared memory
#pragma unroll
                    for (int d = 0; d < D; ++d)
#pragma unroll
                        for (int ib = 0; ib < BN; ++ib) bfr[d][ib] = lds64(cb + (uint32_t)(d * KC * PS + ib * 8) * 8u);
                    if (C::PAIRED) {
#pragma unroll
                        for (int ia = 0; ia + 1 < AM; ia += 2) lds128(ca + (uint32_t)((ia >> 1) * 16) * 8u, afr[ia], afr[ia + 1]);
                        if (AM & 1) afr[AM - 1] = lds64(ca + (uint32_t)((AM >> 1) * 16) * 8u);
                    } else {
#pragma unroll
                        for (int ia = 0; ia < AM; ++ia) afr[ia] = lds64(ca + aoff(ia));
                    }
                };
                uint32_t a_ap = a_st, a_bp = b_st;
                mbar_wait_a(a_full + 8 * st, ph);                        // AO tile + coefficient tile landed
                int nk = nfn_s[st];                                      // multiple of KSTEP, >= KSTEP
                if constexpr (C::ROT) {
                    // ---- rotating-register double step ----------------------------------------------------------------
                    // ncu (profiles/r01_ws_grad_v3_regions.txt, source page): in the one-set scheme above every fragment
                    // load overwrites a register that the DMMA right in front of it reads; the DMMA collects its operands
                    // over several cycles, the load waits for that (short scoreboard) and with it the in-order warp:
                    // ~7 cycles per load, 10 loads per 704-cycle k-step.  Here the A fragments of a k-step are no longer
                    // all resident: the k-step walks the MO block PAIRS (slot = one pair x all B fragments, 4-8 DMMAs) and
                    // the pair of slot q + 1 is fetched during slot q into a small ring of register buffers; the B
                    // fragments are double buffered (the k-step of half h reads set h and fetches set h ^ 1).
                    // Order per accumulator is unchanged (k-steps in sequence), so the results are bit-identical.
                    constexpr int NPAIR = C::NPAIR, NP2 = C::NP2, NBUF = C::NBUF, PD = C::PD;
                    double abuf[NBUF][2], bset[2][NB];
                    constexpr uint32_t A_HALF = (uint32_t)(4 * CS) * 8u, B_HALF = (uint32_t)(4 * (C::IL ? C::RS : PS)) * 8u;
                    // the B fragments (set d, point block ib) = bset[.][d * BN + ib] of one k-step from the tile rows at `base`
                    auto load_b = [&](double (&dst)[NB], const uint32_t base) {
                        if constexpr (C::IL) {
#pragma unroll
                            for (int dp = 0; dp < C::DP; ++dp)
#pragma unroll
                                for (int ib = 0; ib < BN; ++ib) {
                                    const uint32_t a = base + (uint32_t)(dp * KC * C::RS + ib * 16) * 8u;
                                    if (2 * dp + 1 < D) lds128(a, dst[(2 * dp) * BN + ib], dst[(2 * dp + 1) * BN + ib]);
                                    else dst[(2 * dp) * BN + ib] = lds64(a);
                                }
                        } else {
#pragma unroll
                            for (int j = 0; j < NB; ++j) dst[j] = lds64(base + (uint32_t)((j / BN) * KC * PS + (j % BN) * 8) * 8u);
                        }
                    };
                    auto dstep = [&](const uint32_t ca, const uint32_t cb, const uint32_t na, const uint32_t nb) {
                        static_for<0, NP2>([&](auto qc) {
                            constexpr int q = decltype(qc)::value;
                            constexpr int h = q / NPAIR, pq = q % NPAIR;
                            // the B fragments of the next k-step first: issued while set h is still needed by every
                            // DMMA of this half, so the two sets cannot share registers
                            if constexpr (pq == 0) {
                                load_b(bset[h ^ 1], (h == 0) ? cb + B_HALF : nb);
                            }
                            // the pair of slot q + 1
                            constexpr int t = q + PD, tt = t % NP2, th = tt / NPAIR, tp = tt % NPAIR;
                            const uint32_t abase = (t >= NP2 ? na : ca) + (uint32_t)th * A_HALF + (uint32_t)tp * 128u;
                            if constexpr (2 * tp + 1 < AM) lds128(abase, abuf[tt % NBUF][0], abuf[tt % NBUF][1]);
                            else abuf[tt % NBUF][0] = lds64(abase);
#pragma unroll
                            for (int j = 0; j < NB; ++j)
#pragma unroll
                                for (int e = 0; e < 2; ++e)
                                    if (2 * pq + e < AM)
                                        dmma_m8n8k4(acc[2 * pq + e][j % BN][j / BN][0], acc[2 * pq + e][j % BN][j / BN][1],
                                                    abuf[q % NBUF][e], bset[h][j]);
                        });
                    };
                    auto load_first = [&](const uint32_t ca, const uint32_t cb) {
                        load_b(bset[0], cb);
                        static_for<0, PD>([&](auto tc_) {
                            constexpr int t = decltype(tc_)::value, th = t / NPAIR, tp = t % NPAIR;
                            const uint32_t abase = ca + (uint32_t)th * A_HALF + (uint32_t)tp * 128u;
                            if constexpr (2 * tp + 1 < AM) lds128(abase, abuf[t % NBUF][0], abuf[t % NBUF][1]);
                            else abuf[t % NBUF][0] = lds64(abase);
                        });
                    };
                    load_first(a_ap, a_bp);
                    for (int c = 0; c < p.nchunk; ++c) {
                        // the next stage of the ring; its barrier is probed now and only waited for (behind the k-steps
                        // of this chunk) if the producers were not a chunk ahead
                        uint32_t st_n = st, ph_n = ph, a_n = a_st, b_n = b_st;
                        advance(st_n, ph_n, a_n, b_n);
                        const bool more = c + 1 < p.nchunk;
                        // (the probe sits in ONE basic block with the chunk's first double step, peeled out of the loop: its
                        // S2UR / SYNCS / predicate latencies, ~75 cycles, overlap with DMMAs instead of stalling the warp
                        // in front of the loop)
                        bool ready = false;
                        if (nk > 8) {
                            ready = more && mbar_test_a(a_full + 8 * st_n, ph_n);
                            dstep(a_ap, a_bp, a_ap + 2 * A_HALF, a_bp + 2 * B_HALF);
                            a_ap += 2 * A_HALF;
                            a_bp += 2 * B_HALF;
#pragma unroll 1
                            for (int k0 = 16; k0 < nk; k0 += 8) {
                                dstep(a_ap, a_bp, a_ap + 2 * A_HALF, a_bp + 2 * B_HALF);
                                a_ap += 2 * A_HALF;
                                a_bp += 2 * B_HALF;
                            }
                        } else {
                            ready = more && mbar_test_a(a_full + 8 * st_n, ph_n);
                        }
                        uint32_t n_ap = a_ap, n_bp = a_bp;               // behind the last chunk: harmless reloads
                        int nk_next = nk;
                        if (more) {
                            if (!ready) mbar_wait_a(a_full + 8 * st_n, ph_n);
                            nk_next = nfn_s[st_n];
                            n_ap = a_n; n_bp = b_n;
                        }
                        dstep(a_ap, a_bp, n_ap, n_bp);
                        __syncwarp();
                        if (lane == 0) mbar_arrive_a(a_empty + 8 * st);  // every read of this stage has returned
                        a_ap = n_ap; a_bp = n_bp; nk = nk_next;
                        st = st_n; ph = ph_n; a_st = a_n; b_st = b_n;
                    }
                } else {
                load_all(a_ap, a_bp);
                for (int c = 0; c < p.nchunk; ++c) {
                    uint32_t st_n = st, ph_n = ph, a_n = a_st, b_n = b_st;
                    advance(st_n, ph_n, a_n, b_n);
#pragma unroll 1
                    for (int k0 = 4; k0 < nk; k0 += 4)                   // steps 0 .. nk/4-2: next step in this chunk
                        step(a_ap + (uint32_t)(k0 * CS) * 8u, a_bp + (uint32_t)(k0 * PS) * 8u);
                    // last step of the chunk: refill from the next chunk of this pass (or, behind the last chunk,
                    // harmlessly from this one)
                    uint32_t n_ap = a_ap, n_bp = a_bp;
                    int nk_next = nk;
                    // With only two stages the consumer must hand its stage back BEFORE it waits for the next one
                    // (otherwise the producers idle for a k-step per chunk: measured +5% on rho + laplacian), so
                    // the pipeline is seamless only for NST >= 3.
                    constexpr bool SEAMLESS = (NST >= 3);
                    if (SEAMLESS && c + 1 < p.nchunk) {
                        mbar_wait_a(a_full + 8 * st_n, ph_n);
                        nk_next = nfn_s[st_n];
                        n_ap = a_n; n_bp = b_n;
                    }
                    step(n_ap, n_bp);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(a_empty + 8 * st);      // every read of this stage has returned
                    if (!SEAMLESS && c + 1 < p.nchunk) {
                        mbar_wait_a(a_full + 8 * st_n, ph_n);
                        nk_next = nfn_s[st_n];
                        n_ap = a_n; n_bp = b_n;
                        load_all(n_ap, n_bp);
                    }
                    a_ap = n_ap; a_bp = n_bp; nk = nk_next;
                    st = st_n; ph = ph_n; a_st = a_n; b_st = b_n;
                }
                }   // !ROT
                // ---- remainder orbitals: sum the producers' partial sums (lane (tr, tc) of warp row 0 takes orbital
                // tr < REM at its two points per point block, the layout of an accumulator block) ----------------------
                double rv[BN][D][2];
                const bool ract = REM > 0 && wm == 0 && tr < REM;
                if constexpr (REM > 0) {
#pragma unroll
                    for (int ib = 0; ib < BN; ++ib)
#pragma unroll
                        for (int d = 0; d < D; ++d) rv[ib][d][0] = rv[ib][d][1] = 0.0;
                    if (ract) {
#pragma unroll
                        for (int w = 0; w < NPW; ++w)                    // fixed order: results are reproducible
#pragma unroll
                            for (int d = 0; d < D; ++d)
#pragma unroll
                                for (int ib = 0; ib < BN; ++ib) {
                                    const double2 v = *reinterpret_cast<const double2 *>(
                                        part + (((size_t)w * REM + tr) * D + d) * P + pt_w + ib * 8 + 2 * tc);
                                    rv[ib][d][0] += v.x;
                                    rv[ib][d][1] += v.y;
                                }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_a(a_rfree);               // the producers may write the next pass' sums
                }
                // ---- per-MO-tile epilogues: lane holds MO row tr of each block, points 2*tc + {0,1} -----
                // act: the lane holds an orbital (always for the DMMA blocks; lanes tr < REM for the remainder orbitals)
                auto mo_store = [&](const int mo, const bool act, const double (&a)[BN][D][2]) {
                    if (!act || mo >= p.n_mo) return;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
                        const int code = (SET == SET_ONE) ? p.one_code : d;
                        const int sl = p.slot[code];
                        if (sl < 0) continue;
                        double *orow = p.out + (size_t)sl * p.slot_stride + (size_t)mo * p.ld + q0;
#pragma unroll
                        for (int ib = 0; ib < BN; ++ib) {
                            const int pt = pt_w + ib * 8 + 2 * tc;
                            if (q0 + pt < p.npts) orow[pt] = a[ib][d][0];
                            if (q0 + pt + 1 < p.npts) orow[pt + 1] = a[ib][d][1];
                        }
                    }
                };
                auto rho_block = [&](const int mo, const bool act, const double (&a)[BN][D][2]) {
                    const double oc = act ? p.occ[mo] : 0.0;             // zero for padding MOs
                    const double o2 = oc * 2.0;
                    double nrm = 0.0;
#pragma unroll
                    for (int ib = 0; ib < BN; ++ib)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            if (SET == SET_D2P) {
                                // second pass of rho + laplacian on the MO values the first pass left in HBM:
                                // sum_i 2 occ phi d2phi
                                const int q = q0 + pt_w + ib * 8 + 2 * tc + e;
                                const double phi = (act && q < p.npts && mo < p.n_mo) ? __ldg(p.phi + (size_t)mo * p.ldp + q) : 0.0;
#pragma unroll
                                for (int d = 0; d < D; ++d) osum[d][ib][e] += o2 * (a[ib][d][e] * phi);
                                continue;
                            }
                            const double phi = a[ib][0][e];
                            if ((q0 + pt_w + ib * 8 + 2 * tc + e) < p.npts) nrm += phi * phi;
                            if (SET == SET_GRAD && p.epi == 3) {
                                const int q = q0 + pt_w + ib * 8 + 2 * tc + e;
                                if (act && q < p.npts && mo < p.n_mo) p.phi[(size_t)mo * p.ldp + q] = phi;
                            }
                            osum[0][ib][e] += oc * (phi * phi);
                            if (SET == SET_D2) {
                                // second pass of rho + laplacian: sum_i 2 occ phi d2phi (the first pass added
                                // sum_i 2 occ (d phi)^2 and wrote rho)
                                osum[1][ib][e] += o2 * (a[ib][1][e] * phi);
                                osum[2][ib][e] += o2 * (a[ib][2][e] * phi);
                                osum[3][ib][e] += o2 * (a[ib][3][e] * phi);
                            } else if (D >= 4) {
                                const double gx = a[ib][1][e], gy = a[ib][2][e], gz = a[ib][3][e];
                                if (SET == SET_GRAD && p.epi != 0) {         // first pass of rho + laplacian
                                    osum[1][ib][e] += o2 * (gx * gx);
                                    osum[2][ib][e] += o2 * (gy * gy);
                                    osum[3][ib][e] += o2 * (gz * gz);
                                } else {
                                    osum[1][ib][e] += o2 * (gx * phi);
                                    osum[2][ib][e] += o2 * (gy * phi);
                                    osum[3][ib][e] += o2 * (gz * phi);
                                }
                                if (D >= 7) {
                                    osum[4][ib][e] += o2 * (a[ib][4][e] * phi + gx * gx);
                                    osum[5][ib][e] += o2 * (a[ib][5][e] * phi + gy * gy);
                                    osum[6][ib][e] += o2 * (a[ib][6][e] * phi + gz * gz);
                                }
                                if (D >= 10) {
                                    osum[7][ib][e] += o2 * (a[ib][7][e] * phi + gx * gy);
                                    osum[8][ib][e] += o2 * (a[ib][8][e] * phi + gx * gz);
                                    osum[9][ib][e] += o2 * (a[ib][9][e] * phi + gy * gz);
                                }
                            }
                        }
                    if (p.mo_norm != nullptr) {
                        nrm += __shfl_xor_sync(0xffffffffu, nrm, 1);     // over the 4 lanes of a row
                        nrm += __shfl_xor_sync(0xffffffffu, nrm, 2);
                        if (act && tc == 0 && mo < p.n_mo) atomicAdd(p.mo_norm + mo, nrm);
                    }
                };
                if (SINK == SINK_MO) {
#pragma unroll
                    for (int ia = 0; ia < AM; ++ia)
                        if (ia < nblk) mo_store(mt * MCT + mo_w + ia * 8 + tr, true, acc[ia]);
                    if constexpr (REM > 0) mo_store(mt * MCT + MC + (ract ? tr : 0), ract, rv);
                }
                if (SINK == SINK_RHO) {
#pragma unroll
                    for (int ia = 0; ia < AM; ++ia)
                        if (ia < nblk) rho_block(mt * MCT + mo_w + ia * 8 + tr, true, acc[ia]);   // nblk is warp-uniform
                    if constexpr (REM > 0) rho_block(mt * MCT + MC + (ract ? tr : 0), ract, rv);
                }
            }   // mt
            if (SINK == SINK_RHO) {
                // sum over the 8 MO rows held by different lanes (xor 4, 8, 16), then over the WM warp
                // rows through shared memory (own scratch: the producers are already filling stages of
                // the next tile)
#pragma unroll
                for (int o = 0; o < C::NOUT; ++o)
#pragma unroll
                    for (int ib = 0; ib < BN; ++ib)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double v = osum[o][ib][e];
                            v += __shfl_xor_sync(0xffffffffu, v, 4);
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            if (tr == 0) red[((size_t)wm * C::NOUT + o) * P + pt_w + ib * 8 + 2 * tc + e] = v;
                        }
                named_bar(2, NCW * 32);
                for (int e = tid; e < C::NOUT * P; e += NCW * 32) {
                    const int o = e / P, pt = e - o * P;
                    double sum = 0.0;
#pragma unroll
                    for (int w = 0; w < WM; ++w) sum += red[((size_t)w * C::NOUT + o) * P + pt];
                    if (q0 + pt < p.npts) {
                        const bool two_pass = (SET == SET_D2) || (SET == SET_GRAD && p.epi != 0);
                        if (SET == SET_D2P) {                                // all three sums go to the slots of codes 4..6
                            const int sl = p.slot[o + 4];
                            if (sl >= 0) p.delta[(size_t)sl * p.ld + q0 + pt] += sum;
                        } else if (o == 0) {
                            if (p.rho != nullptr && SET != SET_D2) p.rho[q0 + pt] = sum;
                        } else {
                            const int sl = p.slot[two_pass ? o + 3 : o];     // two-pass laplacian: codes 4..6
                            if (sl >= 0) {
                                double *dst = p.delta + (size_t)sl * p.ld + q0 + pt;
                                *dst = (SET == SET_D2) ? *dst + sum : sum;
                            }
                        }
                    }
                }
                named_bar(2, NCW * 32);
            }
        }
    }
}

}  // namespace okb
