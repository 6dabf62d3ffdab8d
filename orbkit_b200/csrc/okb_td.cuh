// okb_td.cuh -- time-dependent detCI contractions (replaces the OpenMP loops of orbkit/detci/cy_ci.pyx:101-122
// get_rho_full and 126-151 get_j_full):
//
//   out[t][x] = sum_k w[t][k] * in[k][x]        t < nt time steps, k < nk state pairs, x < n points
//
// get_rho_full: in = the transition densities rho[count][x] of the state pairs count = (n, m >= n), w[t][count] =
// ReS[t,m,m] (m == n) or 2 ReS[t,m,n]; get_j_full: in = j[count][3][x] read as rows of 3*npts entries, w[t][count] =
// -2 ImS[t,n,m] (0 for m == n).  The weights are packed by the caller (nt x nk doubles); 2 * S is exact, so the packed
// products are the reference's.
//
// A dense (nt x nk) . (nk x n) product with small nk (3 .. a few hundred) and huge n: between the HBM roofline (8 (nk + nt)
// bytes per point; store bound below nk ~ 23) and the FP64 one.  FP64 tensor path: mma.sync.m8n8k4.f64 (DMMA; tcgen05 has
// no f64 kind), the k-steps of 4 state pairs in the reference's order -- the sum runs in the reference's order up to the
// fused rounding inside a DMMA, so the results agree to a few ulp of the largest term, not bit for bit (stated tolerance
// in tests/test_gpu_ci.py).
//
//   CTA = 4 warps, tile = TD_P = 128 points (warp w: points [32 w, 32 w + 32) = 4 point blocks) x all time steps in
//   passes of 8 MTB rows (MTB = 4 blocks of 8: 16 accumulator blocks = 64 registers per thread; MTB = 8 when the k range
//   needs several chunks, i.e. when every pass re-reads `in`: half the passes).  Also the cy_core.mocreator drop-in
//   (okb_mocreator: t = MO, k = AO).
//   The `in` tile [kc][TD_P] is staged ONCE per point tile (whole k range when nk <= TD_KC, else per pass and chunk),
//   the weights of a pass [32][kc] per pass; both row strides = 4 (mod 16) doubles: conflict-free fragment loads.
//   HBM traffic = `in` once + `out` once (the weights, nt x nk doubles, stay in L2).
#pragma once
#include "okb_ws.cuh"

namespace okb {

constexpr int TD_P = 128, TD_KC = 64, TD_NT = 128;
constexpr int TD_PS = TD_P + 4;                 // 132 = 4 (mod 16)
constexpr int TD_WS = TD_KC + 4;                // 68  = 4 (mod 16)
// dynamic shared memory: the `in` tile takes only the rows it needs (few state pairs -> more CTAs per SM)
__host__ __device__ constexpr size_t td_smem(int kp, int mtb) {
    return ((size_t)(kp < TD_KC ? kp : TD_KC) * TD_PS + (size_t)(8 * mtb) * TD_WS) * 8;
}

struct TdParams {
    const double *w;       // [ntp][kp] device: nt rows padded to a multiple of 64, nk to a multiple of 4, zero filled
    const double *in;      // [nk][ldi]
    double *out;           // [nt][ldo]
    long long ldi, ldo, n;
    int nt, nk, kp;
    int vec_ok;            // out rows 16-byte aligned: paired stores
    // RDM form (dense one-particle-matrix contraction of the detCI sums, OKB_FLAG_CI_FAST with many terms per orbital pair):
    //   out[d][x] = sum_a phi[rows[a]][x] * sum_b w[a][b] in_d[rows[b]][x],   in_d = in + d * dstride_in, d < ncomp
    // nt = nk = the number of orbitals the term list refers to; CTA (tile, d) = blockIdx.x = tile * ncomp + d, so the
    // CTAs that share a phi tile run back to back (L2).
    const double *phi;     // [*][ldi]
    const int *rows;       // [nk] row of orbital k in phi / in
    long long dstride_in;  // distance of the derivative blocks of `in`
    int ncomp;
};

template <int MTB, bool RDM = false>
__global__ void __launch_bounds__(TD_NT, MTB == 4 ? 4 : 2) okb_td_kernel(const TdParams p) {
    constexpr int TD_MT = 8 * MTB;
    extern __shared__ __align__(16) unsigned char td_smem_raw[];
    double *rt = reinterpret_cast<double *>(td_smem_raw);       // [min(kp, TD_KC)][TD_PS]
    double *wt = rt + (size_t)(p.kp < TD_KC ? p.kp : TD_KC) * TD_PS;   // [TD_MT][TD_WS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr = lane >> 2, tc = lane & 3;
    const int comp = RDM ? (int)(blockIdx.x % (unsigned)p.ncomp) : 0;
    const long long x0 = (long long)(RDM ? blockIdx.x / (unsigned)p.ncomp : blockIdx.x) * TD_P;
    const double *in = RDM ? p.in + (size_t)comp * p.dstride_in : p.in;
    const int nchunk = (p.kp + TD_KC - 1) / TD_KC;
    auto stage_in = [&](int k0, int kc) {                       // rows [k0, k0 + kc) of `in`, zero beyond nk / n
        for (int e = tid; e < kc * TD_P; e += TD_NT) {
            const int k = e / TD_P, pt = e - k * TD_P;
            const long long x = x0 + pt;
            double v = 0.0;
            if (k0 + k < p.nk && x < p.n) v = __ldg(in + (size_t)(RDM ? __ldg(p.rows + k0 + k) : k0 + k) * p.ldi + x);
            rt[(size_t)k * TD_PS + pt] = v;
        }
    };
    double red[4][2];                                           // RDM: sum over the orbitals a of phi_a (w in)_a
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) red[nb][0] = red[nb][1] = 0.0;
    if (nchunk == 1) stage_in(0, p.kp);
    const uint32_t a_rt = smem_u32(rt) + (uint32_t)((tc * TD_PS + warp * 32 + tr) * 8);
    const uint32_t a_wt = smem_u32(wt) + (uint32_t)((tr * TD_WS + tc) * 8);
    for (int t0 = 0; t0 < p.nt; t0 += TD_MT) {
        double acc[MTB][4][2];
#pragma unroll
        for (int mb = 0; mb < MTB; ++mb)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
        for (int c = 0; c < nchunk; ++c) {
            const int k0 = c * TD_KC, kc = min(TD_KC, p.kp - k0);
            __syncthreads();                                    // previous pass / chunk no longer reads wt (rt)
            if (nchunk > 1) stage_in(k0, kc);
            for (int e = tid; e < TD_MT * kc; e += TD_NT) {     // weights of the pass (rows beyond nt are zero padded)
                const int t = e / kc, k = e - t * kc;
                wt[(size_t)t * TD_WS + k] = __ldg(p.w + (size_t)(t0 + t) * p.kp + k0 + k);
            }
            __syncthreads();
#pragma unroll 2
            for (int ks = 0; ks < kc; ks += 4) {
                double a[MTB], b[4];
#pragma unroll
                for (int mb = 0; mb < MTB; ++mb) a[mb] = lds64(a_wt + (uint32_t)((mb * 8 * TD_WS + ks) * 8));
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) b[nb] = lds64(a_rt + (uint32_t)((ks * TD_PS + nb * 8) * 8));
#pragma unroll
                for (int mb = 0; mb < MTB; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma_m8n8k4(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
            }
        }
        // lane holds out[t0 + 8 mb + tr][x0 + 32 warp + 8 nb + 2 tc + {0, 1}]
        if constexpr (RDM) {
#pragma unroll
            for (int mb = 0; mb < MTB; ++mb) {
                const int t = t0 + mb * 8 + tr;
                if (t >= p.nt) continue;
                const double *prow = p.phi + (size_t)__ldg(p.rows + t) * p.ldi;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const long long x = x0 + warp * 32 + nb * 8 + 2 * tc;
                    double f0 = 0.0, f1 = 0.0;
                    if (p.vec_ok && x + 1 < p.n) {
                        const double2 f = __ldg(reinterpret_cast<const double2 *>(prow + x));
                        f0 = f.x; f1 = f.y;
                    } else {
                        if (x < p.n) f0 = __ldg(prow + x);
                        if (x + 1 < p.n) f1 = __ldg(prow + x + 1);
                    }
                    red[nb][0] = fma(acc[mb][nb][0], f0, red[nb][0]);
                    red[nb][1] = fma(acc[mb][nb][1], f1, red[nb][1]);
                }
            }
        } else {
#pragma unroll
            for (int mb = 0; mb < MTB; ++mb) {
                const int t = t0 + mb * 8 + tr;
                if (t >= p.nt) continue;
                double *orow = p.out + (size_t)t * p.ldo;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const long long x = x0 + warp * 32 + nb * 8 + 2 * tc;
                    if (p.vec_ok && x + 1 < p.n) {
                        *reinterpret_cast<double2 *>(orow + x) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
                    } else {
                        if (x < p.n) orow[x] = acc[mb][nb][0];
                        if (x + 1 < p.n) orow[x + 1] = acc[mb][nb][1];
                    }
                }
            }
        }
    }
    if (RDM) {                                                  // the eight row groups of the lanes, fixed order
        double *orow = p.out + (size_t)comp * p.ldo;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double v = red[nb][h];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                red[nb][h] = v;
            }
            const long long x = x0 + warp * 32 + nb * 8 + 2 * tc;
            if (tr != 0) continue;
            if (x < p.n) orow[x] = red[nb][0];
            if (x + 1 < p.n) orow[x + 1] = red[nb][1];
        }
    }
}

// ---- k ranges of several chunks (more than TD_KC state pairs; cy_core.mocreator: k = AO): pipelined variant -------------
// The kernel above re-stages the `in` chunk and the weights synchronously for every (pass, chunk): with several chunks the
// DMMA pipe idles during the staging (990 functions x 82 rows: 0.21 of the FP64 peak).  Here EIGHT DMMA warps (two per
// sub-partition: 2 warp rows x 4 point groups; a warp row takes every other block of 8 rows of the pass, x 32 points) walk
// the flattened (pass of 64 rows, chunk of 32 k) sequence through a three-stage ring that a NINTH warp fills with 16-byte
// cp.async copies two iterations ahead, handed over through two named barriers in producer / consumer form, staging hidden behind the 128 DMMAs a warp issues
// per chunk.  Row blocks beyond nt are skipped (82 rows: 11
// instead of 16 blocks).  Needs 16-byte aligned `in` rows (else the kernel above).
constexpr int TD2_NT = 288, TD2_KC = 32, TD2_NS = 3, TD2_MT = 64;   // 8 DMMA warps + 1 staging warp
constexpr int TD2_WS = TD2_KC + 4;                              // 36 = 4 (mod 16)
constexpr size_t TD2_STAGE = ((size_t)TD2_KC * TD_PS + (size_t)TD2_MT * TD2_WS) * 8;
constexpr size_t TD2_SMEM = TD2_NS * TD2_STAGE;

template <bool RDM>
__global__ void __launch_bounds__(TD2_NT, 1) okb_td2_kernel(const TdParams p) {
    extern __shared__ __align__(16) unsigned char td2_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tr = lane >> 2, tc = lane & 3, wr = warp >> 2, wc = warp & 3;
    const int comp = RDM ? (int)(blockIdx.x % (unsigned)p.ncomp) : 0;
    const long long x0 = (long long)(RDM ? blockIdx.x / (unsigned)p.ncomp : blockIdx.x) * TD_P;
    const double *in = RDM ? p.in + (size_t)comp * p.dstride_in : p.in;
    const int nchunk = (p.kp + TD2_KC - 1) / TD2_KC, npass = (p.nt + TD2_MT - 1) / TD2_MT, nit = npass * nchunk;
    const uint32_t a_raw = smem_u32(td2_raw);
    if (warp == 8) {
        // ---- staging warp: all copies of the ring (ncu of the first version, in which the eight DMMA warps staged their
        // own share after the barrier: half of all instructions were staging, issued while the DMMA pipe idled).
        // Lane slots, the same in every iteration: `in` row r (r < kc), 16-byte pieces lane and lane + 32; weight rows
        // (lane >> 4) + 2 j (j < 32) at piece lane & 15.
        const uint32_t xb0 = x0 + 2 * lane + 1 < p.n ? 16u : x0 + 2 * lane < p.n ? 8u : 0u;
        const uint32_t xb1 = x0 + 2 * lane + 65 < p.n ? 16u : x0 + 2 * lane + 64 < p.n ? 8u : 0u;
        const double *in_x = in + x0 + 2 * lane;
        const int sw = lane >> 4, pw = lane & 15;
        auto stage = [&](int it) {                              // copies of iteration `it` into stage it % NS, one group
            const int ps = it / nchunk, c = it - ps * nchunk, k0 = c * TD2_KC, kc = min(TD2_KC, p.kp - k0), t0 = ps * TD2_MT;
            const uint32_t a_rt = a_raw + (uint32_t)((it % TD2_NS) * TD2_STAGE) + (uint32_t)(2 * lane * 8);
            const uint32_t a_wt = a_raw + (uint32_t)((it % TD2_NS) * TD2_STAGE) + (uint32_t)(TD2_KC * TD_PS * 8) +
                                  (uint32_t)((sw * TD2_WS + 2 * pw) * 8);
#pragma unroll 4
            for (int k = 0; k < kc; ++k) {                      // `in`: zero filled beyond nk / n
                const int row = k0 + k;
                const bool live = row < p.nk;
                const double *src = live ? in_x + (size_t)(RDM ? __ldg(p.rows + row) : row) * p.ldi : in;
                const uint32_t b0 = live ? xb0 : 0u, b1 = live ? xb1 : 0u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_rt + (uint32_t)(k * TD_PS * 8)),
                             "l"(b0 ? src : in), "r"(b0)
                             : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a_rt + (uint32_t)((k * TD_PS + 64) * 8)),
                             "l"(b1 ? src + 64 : in), "r"(b1)
                             : "memory");
            }
            if (2 * pw < kc) {                                  // weights: 64 rows (zero padded beyond nt) x kc / 2 pieces
                const double *wsrc = p.w + (size_t)(t0 + sw) * p.kp + k0 + 2 * pw;
#pragma unroll 8
                for (int j = 0; j < TD2_MT / 2; ++j)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a_wt + (uint32_t)(2 * j * TD2_WS * 8)),
                                 "l"(wsrc + (size_t)2 * j * p.kp)
                                 : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // hand-over by two named barriers in producer / consumer form (PTX bar.arrive / bar.sync with the thread count):
        // FULL (1): this warp arrives when the copies of iteration `it` have landed, the DMMA warps wait on it;
        // EMPTY (2): the DMMA warps arrive when they are done with an iteration, this warp waits on it before it reuses
        // that stage -- and before it arrives on FULL again, so no barrier ever sees two phases at once.
        stage(0);
        if (nit > 1) stage(1);
        for (int it = 0; it < nit; ++it) {
            if (it >= 1) named_bar(2, TD2_NT);                  // the DMMA warps left iteration it - 1
            if (it + 1 < nit) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __threadfence_block();
            asm volatile("bar.arrive 1, %0;" ::"r"(TD2_NT) : "memory");
            if (it >= 1 && it + 2 < nit) stage(it + 2);         // into the stage of iteration it - 1
            else if (it == 0 && nit > 2) stage(2);              // third stage: never used so far
        }
        return;
    }
    double acc[4][4][2], red[4][2];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        red[nb][0] = red[nb][1] = 0.0;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    }
    for (int it = 0; it < nit; ++it) {
        if (it >= 1) asm volatile("bar.arrive 2, %0;" ::"r"(TD2_NT) : "memory");      // done with iteration it - 1
        named_bar(1, TD2_NT);                                   // FULL: the copies of iteration `it` have landed
        const int ps = it / nchunk, c = it - ps * nchunk, kc = min(TD2_KC, p.kp - c * TD2_KC), t0 = ps * TD2_MT;
        const uint32_t a_rt = a_raw + (uint32_t)((it % TD2_NS) * TD2_STAGE) + (uint32_t)((tc * TD_PS + wc * 32 + tr) * 8);
        // the row blocks of a pass alternate between the two warp rows (block 2 mb + wr): a ragged last pass is shared
        const uint32_t a_wt = a_raw + (uint32_t)((it % TD2_NS) * TD2_STAGE) + (uint32_t)(TD2_KC * TD_PS * 8) +
                              (uint32_t)(((wr * 8 + tr) * TD2_WS + tc) * 8);
        const int nblk = min(8, (p.nt - t0 + 7) >> 3);          // row blocks of this pass that exist
        const int nmbv = (nblk - wr + 1) >> 1;                  // ... of this warp row (warp-uniform)
        auto kstep = [&](int ks) {
            double b[4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) b[nb] = lds64(a_rt + (uint32_t)((ks * TD_PS + nb * 8) * 8));
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                if (mb >= nmbv) continue;
                const double a = lds64(a_wt + (uint32_t)((mb * 16 * TD2_WS + ks) * 8));
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma_m8n8k4(acc[mb][nb][0], acc[mb][nb][1], a, b[nb]);
            }
        };
        if (kc == TD2_KC && nmbv == 4) {                        // full chunk of a full pass: straight-line code, all
#pragma unroll                                                  // fragment loads of a k-step in front of its 16 DMMAs
            for (int ks = 0; ks < TD2_KC; ks += 4) {
                double a[4], b[4];
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) b[nb] = lds64(a_rt + (uint32_t)((ks * TD_PS + nb * 8) * 8));
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) a[mb] = lds64(a_wt + (uint32_t)((mb * 16 * TD2_WS + ks) * 8));
#pragma unroll
                for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                    for (int nb = 0; nb < 4; ++nb) dmma_m8n8k4(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
            }
        } else {
            for (int ks = 0; ks < kc; ks += 4) kstep(ks);
        }
        if (c != nchunk - 1) continue;
        // end of a pass: lane holds out[t0 + 8 (2 mb + wr) + tr][x0 + 32 wc + 8 nb + 2 tc + {0, 1}]
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
            const int t = t0 + (2 * mb + wr) * 8 + tr;
            if (t < p.nt) {
                const double *prow = RDM ? p.phi + (size_t)__ldg(p.rows + t) * p.ldi : nullptr;
                double *orow = RDM ? nullptr : p.out + (size_t)t * p.ldo;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const long long x = x0 + wc * 32 + nb * 8 + 2 * tc;
                    if constexpr (RDM) {
                        double f0 = 0.0, f1 = 0.0;
                        if (p.vec_ok && x + 1 < p.n) {
                            const double2 f = __ldg(reinterpret_cast<const double2 *>(prow + x));
                            f0 = f.x; f1 = f.y;
                        } else {
                            if (x < p.n) f0 = __ldg(prow + x);
                            if (x + 1 < p.n) f1 = __ldg(prow + x + 1);
                        }
                        red[nb][0] = fma(acc[mb][nb][0], f0, red[nb][0]);
                        red[nb][1] = fma(acc[mb][nb][1], f1, red[nb][1]);
                    } else {
                        if (p.vec_ok && x + 1 < p.n) {
                            *reinterpret_cast<double2 *>(orow + x) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
                        } else {
                            if (x < p.n) orow[x] = acc[mb][nb][0];
                            if (x + 1 < p.n) orow[x + 1] = acc[mb][nb][1];
                        }
                    }
                }
            }
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
        }
    }
    if constexpr (RDM) {                                        // the eight row groups of the lanes, then the two row halves
        double *scratch = reinterpret_cast<double *>(td2_raw);
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double v = red[nb][h];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                red[nb][h] = v;
            }
        named_bar(3, TD2_NT - 32);                                        // the stages are no longer read
        if (wr == 1 && tr == 0)
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                scratch[wc * 32 + nb * 8 + 2 * tc] = red[nb][0];
                scratch[wc * 32 + nb * 8 + 2 * tc + 1] = red[nb][1];
            }
        named_bar(3, TD2_NT - 32);
        if (wr == 0 && tr == 0) {
            double *orow = p.out + (size_t)comp * p.ldo;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const long long x = x0 + wc * 32 + nb * 8 + 2 * tc;
                if (x < p.n) orow[x] = red[nb][0] + scratch[wc * 32 + nb * 8 + 2 * tc];
                if (x + 1 < p.n) orow[x + 1] = red[nb][1] + scratch[wc * 32 + nb * 8 + 2 * tc + 1];
            }
        }
    }
}

}  // namespace okb
