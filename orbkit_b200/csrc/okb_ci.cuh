// okb_ci.cuh -- detCI grid contractions over pairs of molecular orbitals (replaces the loops of
// orbkit/detci/cy_ci.pyx:70-97 get_rho, 156-186 get_jab, 211-240 get_a_nabla_b):
//
//   CI_RHO    out[x]    = sum_t c_t mo[a_t][x] mo[b_t][x]
//   CI_JAB    out[d][x] = sum_t -1/2 c_t (mo[a_t][x] dmo[d][b_t][x] - mo[b_t][x] dmo[d][a_t][x])
//   CI_ANB    out[d][x] = sum_t c_t mo[a_t][x] dmo[d][b_t][x]
//   CI_PAIRS  out[t][x] = mo[a_t][x] ket[b_t][x]                    (per-pair products, store bound; ket = mo unless a
//                                                                     second array is given: core.calc_mo_matrix, core.py:925-941)
//   CI_JPAIRS out[d][t][x] = -1/2 (mo[a_t] dmo[d][b_t] - mo[b_t] dmo[d][a_t])   per pair, d < ncomp (extras.calc_jmo,
//                                                                     extras.py:441-493, only the requested pairs)
//   CI_JABF   out[d][x] = sum_t c_t (mo[a_t][x] dmo[d][b_t][x] - mo[b_t][x] dmo[d][a_t][x]),  d < ncomp
//                                                                    (cy_ci.get_jab_full, cy_ci.pyx:186-202: the terms are
//                                                                     the pairs n > m of the state basis, c = ImS[n,m] / mu)
//
// One thread owns one grid point and walks the term list in the caller's order with the reference's
// expression order and NO fused multiply-add (__dmul_rn / __dadd_rn), so for the same MO arrays the
// results are bit-identical to the reference's C loops.  The term list streams through shared memory
// in batches (every lane of a warp needs the same term: broadcast reads); the MO rows are read with
// fully coalesced, L1/L2-cached loads (a warp reads 32 consecutive doubles of row a_t and of row b_t).
// HBM-bound: algorithmic traffic = the MO rows once (8 n_mo n_sets bytes per point) + the output.
#pragma once
#include "okb_common.cuh"

namespace okb {

enum { CI_RHO = 0, CI_JAB = 1, CI_ANB = 2, CI_PAIRS = 3, CI_JPAIRS = 4, CI_JABF = 5 };

struct CiParams {
    const double *mo;          // [n_mo][ld]
    const double *dmo;         // [3][n_mo][ld] (JAB, ANB), [ncomp][n_mo][ld] (JPAIRS), [n_mo][ld] second factor (PAIRS) or null
    long long ld;              // row stride of mo / dmo in points
    long long dstride;         // n_mo * ld: distance between the three derivative blocks
    long long npts;
    int n_terms;
    const double2 *tp;         // term records: x = coefficient, y = the two orbital indices (low word a, high word b) --
                               // one 16-byte load per term instead of three (every kernel here is bound by the load pipe)
    double *out;
    long long ldo;             // row stride of out in points
    int ncomp;                 // JPAIRS, JABF: number of derivative components (1..3)
};

constexpr int CI_NT = 128;     // threads per CTA = points per CTA
constexpr int CI_TB = 1024;    // terms per shared-memory batch

template <int MODE>
__global__ void __launch_bounds__(CI_NT) okb_ci_kernel(const CiParams p) {
    __shared__ double2 s_t[CI_TB];
    const long long x = (long long)blockIdx.x * CI_NT + threadIdx.x;
    const bool live = x < p.npts;
    const long long xc = live ? x : p.npts - 1;              // clamp: every thread takes part in the staging
    const double *mo = p.mo + xc;
    const double *d0 = (MODE == CI_JAB || MODE == CI_ANB || MODE == CI_JPAIRS || MODE == CI_JABF) ? p.dmo + xc
                       : (MODE == CI_PAIRS && p.dmo != nullptr)                 ? p.dmo + xc
                                                                                : mo;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int t0 = 0; t0 < p.n_terms; t0 += CI_TB) {
        const int nb = min(CI_TB, p.n_terms - t0);
        __syncthreads();
        for (int e = threadIdx.x; e < nb; e += CI_NT) {
            s_t[e] = p.tp[t0 + e];
        }
        __syncthreads();
#pragma unroll 4
        for (int e = 0; e < nb; ++e) {
            const double2 rec = s_t[e];
            const double c = rec.x;
            const long long ra = (long long)__double2loint(rec.y) * p.ld, rb = (long long)__double2hiint(rec.y) * p.ld;
            if (MODE == CI_RHO) {
                // rho[x] += citmp*molist[sta,x]*molist[stb,x]        (cy_ci.pyx:88,95)
                acc0 = __dadd_rn(acc0, __dmul_rn(__dmul_rn(c, __ldg(mo + ra)), __ldg(mo + rb)));
            } else if (MODE == CI_PAIRS) {
                // mo_matrix[n,m] = mo_bra[n]*mo_ket[m]                                  (core.py:939-941)
                if (live) p.out[(long long)(t0 + e) * p.ldo + x] = __dmul_rn(__ldg(mo + ra), __ldg(d0 + rb));
            } else if (MODE == CI_JPAIRS) {
                // jmo[:,n] = -0.5*(mo_matrix[:,i,j] - mo_matrix[:,j,i])                 (extras.py:482)
                const double ma = __ldg(mo + ra), mb = __ldg(mo + rb);
                for (int d = 0; d < p.ncomp; ++d) {
                    const double db = __ldg(d0 + d * p.dstride + rb), da = __ldg(d0 + d * p.dstride + ra);
                    if (live)
                        p.out[((long long)d * p.n_terms + t0 + e) * p.ldo + x] =
                            __dmul_rn(-0.5, __dadd_rn(__dmul_rn(ma, db), -__dmul_rn(mb, da)));
                }
            } else if (MODE == CI_JABF) {
                // tmp = tmp + f*ImS[n,m]*(chi_n[n,r]*nabla_chi_n[c,m,r] - chi_n[m,r]*nabla_chi_n[c,n,r])   (cy_ci.pyx:199-201)
                // with a = n, b = m, c = f*ImS[n,m] (the product the reference forms first)
                const double ma = __ldg(mo + ra), mb = __ldg(mo + rb);
                double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
                for (int d = 0; d < 3; ++d)
                    if (d < p.ncomp) {
                        const double db = __ldg(d0 + d * p.dstride + rb), da = __ldg(d0 + d * p.dstride + ra);
                        v[d] = __dmul_rn(c, __dadd_rn(__dmul_rn(ma, db), -__dmul_rn(mb, da)));
                    }
                acc0 = __dadd_rn(acc0, v[0]);
                acc1 = __dadd_rn(acc1, v[1]);
                acc2 = __dadd_rn(acc2, v[2]);
            } else {
                const double ma = __ldg(mo + ra), mb = (MODE == CI_JAB) ? __ldg(mo + rb) : 0.0;
                double v[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const double db = __ldg(d0 + d * p.dstride + rb);
                    if (MODE == CI_JAB) {
                        // jab[d,x] -= 0.5*(citmp*(mo[a]*dmo[d,b] - mo[b]*dmo[d,a]))      (cy_ci.pyx:181-184)
                        const double da = __ldg(d0 + d * p.dstride + ra);
                        v[d] = -__dmul_rn(0.5, __dmul_rn(c, __dadd_rn(__dmul_rn(ma, db), -__dmul_rn(mb, da))));
                    } else {
                        // out[d,x] += citmp*(mo[a]*dmo[d,b])                             (cy_ci.pyx:236-238)
                        v[d] = __dmul_rn(c, __dmul_rn(ma, db));
                    }
                }
                acc0 = __dadd_rn(acc0, v[0]);
                acc1 = __dadd_rn(acc1, v[1]);
                acc2 = __dadd_rn(acc2, v[2]);
            }
        }
    }
    if (!live || MODE == CI_PAIRS || MODE == CI_JPAIRS) return;
    p.out[x] = acc0;
    if (MODE == CI_JABF) {
        if (p.ncomp > 1) p.out[p.ldo + x] = acc1;
        if (p.ncomp > 2) p.out[2 * p.ldo + x] = acc2;
    } else if (MODE != CI_RHO) {
        p.out[p.ldo + x] = acc1;
        p.out[2 * p.ldo + x] = acc2;
    }
}

// ---- fast variant: terms split over the warps of a CTA (OKB_FLAG_CI_FAST) ----------------------------------------------
// The bit-identical kernel above is bound by re-fetched MO rows (3.5x the algorithmic DRAM bytes).  When the caller does not
// need the reference's summation order -- the fused *_from_qc paths, whose MO values come from the device anyway -- the sum
// over the term list is split: a CTA stages ALL n_mo x nsets row segments of a tile of PW points (a power of two <= 32) in
// shared memory with 16-byte async copies (every MO value comes from HBM exactly once; bulk copies of these 32..256-byte
// segments were measured at ~16 cycles per copy, TMA-issue bound), a warp is split into 32 / PW term slots, slot s of warp w
// sums the terms (w 32/PW + s), + NW 32/PW, ... for the tile's points; the slots are added by shuffles, the warps in warp
// order.  Deterministic (the grouping depends on nothing but the term list and PW), agreement with the sequential sum to
// rounding.  Two CTAs per SM: one stages while the other sums.  HBM bound: 8 n_mo nsets bytes per point in, the results out.
struct CiFastParams {
    CiParams p;
    int n_mo, nsets;           // rows per set; sets staged: 1 (rho) or 1 + ncomp (mo + derivative blocks)
    int pw, lpw;               // points per tile = 1 << lpw
    long long ntiles;          // full tiles; the ragged rest of the points is left to the kernel above
};

constexpr int CIF_NW = 16;     // warps per CTA
constexpr size_t CIF_FIXED = (size_t)CIF_NW * 3 * 32 * 8;

template <int MODE>
__global__ void __launch_bounds__(CIF_NW * 32, 2) okb_ci_fast_kernel(const CiFastParams q) {
    extern __shared__ __align__(128) unsigned char cif_smem[];
    const CiParams &p = q.p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int PW = q.pw, nrow = q.n_mo * q.nsets;
    double *part = reinterpret_cast<double *>(cif_smem);                       // [NW][3][PW]
    double *buf = part + CIF_NW * 3 * 32;                                      // [set][mo][PW]
    const uint32_t a_buf = smem_u32(buf);
    const int ncomp = MODE == CI_RHO ? 1 : MODE == CI_JABF ? p.ncomp : 3;
    const int nslot = 32 >> q.lpw, slot = lane >> q.lpw, pt = lane & (PW - 1);
    const int seg16 = PW >> 1;                           // 16-byte pieces of a row segment
    const int ncopy = nrow * seg16;
    const size_t dstr = (size_t)q.n_mo * PW;
    const double *mo = buf + pt, *d0 = mo + dstr;
    for (long long tile = blockIdx.x; tile < q.ntiles; tile += gridDim.x) {
        const long long x0 = tile * PW;
        __syncthreads();                                 // the previous tile (buf, part) is no longer read
        for (int e = tid; e < ncopy; e += CIF_NW * 32) {
            const int r = e >> (q.lpw - 1), piece = e & (seg16 - 1);
            const int sset = r / q.n_mo, m = r - sset * q.n_mo;
            const double *src = (sset == 0 ? p.mo : p.dmo + (size_t)(sset - 1) * p.dstride) + (size_t)m * p.ld + x0 + 2 * piece;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a_buf + (uint32_t)e * 16u), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
#pragma unroll 2
        for (int e = warp * nslot + slot; e < p.n_terms; e += CIF_NW * nslot) {
            const double2 rec = __ldg(p.tp + e);
            const double c = rec.x;
            const int ra = __double2loint(rec.y) * PW, rb = __double2hiint(rec.y) * PW;
            const double ma = mo[ra];
            if (MODE == CI_RHO) {
                acc0 = fma(c * ma, mo[rb], acc0);
            } else if (MODE == CI_ANB) {
                const double cm = c * ma;
                acc0 = fma(cm, d0[rb], acc0);
                acc1 = fma(cm, d0[dstr + rb], acc1);
                acc2 = fma(cm, d0[2 * dstr + rb], acc2);
            } else {
                // JAB: -1/2 c (ma db - mb da);  JABF: c (ma db - mb da)
                const double mb = mo[rb];
                const double f = MODE == CI_JAB ? -0.5 * c : c;
                acc0 = fma(f, fma(ma, d0[rb], -(mb * d0[ra])), acc0);
                if (ncomp > 1) acc1 = fma(f, fma(ma, d0[dstr + rb], -(mb * d0[dstr + ra])), acc1);
                if (ncomp > 2) acc2 = fma(f, fma(ma, d0[2 * dstr + rb], -(mb * d0[2 * dstr + ra])), acc2);
            }
        }
        for (int off = 16; off >= PW; off >>= 1) {       // the term slots of the warp
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, off);
            if (MODE != CI_RHO) {
                acc1 += __shfl_xor_sync(0xffffffffu, acc1, off);
                acc2 += __shfl_xor_sync(0xffffffffu, acc2, off);
            }
        }
        if (lane < PW) {
            part[(warp * 3 + 0) * 32 + lane] = acc0;
            if (MODE != CI_RHO) {
                part[(warp * 3 + 1) * 32 + lane] = acc1;
                part[(warp * 3 + 2) * 32 + lane] = acc2;
            }
        }
        __syncthreads();
        if (warp < ncomp && lane < PW) {                 // warp d adds the partial sums of component d in warp order
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < CIF_NW; ++w) sum += part[(w * 3 + warp) * 32 + lane];
            p.out[(size_t)warp * p.ldo + x0 + lane] = sum;
        }
    }
}

// ---- sequential sums from shared memory: the bit-identical path for MANY terms over FEW rows of several sets -------------
// With thousands of terms per point tile the gather kernel re-reads its rows from L1 / L2 / DRAM over and over (40 MOs x 4
// sets, 20 000 terms: 740 GB of DRAM reads for 1.1 GB of MO values).  When all rows of a tile of 32 NWP points fit in
// shared memory they are staged ONCE (16-byte async copies) and every lane walks the whole term list for its point with the
// reference's expression order -- the same __dmul_rn / __dadd_rn sequence as okb_ci_kernel, so the same bits.  The three
// components of jab / a_nabla_b are independent sums: component d runs on its own warp (NCW = 3 warps per 32 points; one
// warp for all three was measured at 151 instead of 85 ms: too few warps per SM).  Several CTAs per SM: one stages while
// the others sum.  Used for the derivative modes only: for rho (one set) the gather kernel's rows stay in L1 and it is the
// faster one (15.6 against 18.3 ms), as it is for few terms over many rows (see below).
struct CiSeqParams {
    CiParams p;
    int n_mo, nsets, nwp, ncw;  // rows per set, sets staged, 32-point groups per tile, component warps per group
    long long ntiles;
};

template <int MODE>
__global__ void __launch_bounds__(384) okb_ci_seq_kernel(const CiSeqParams q) {
    extern __shared__ __align__(128) unsigned char cis_smem[];
    const CiParams &p = q.p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;
    const int TP = 32 * q.nwp, nrow = q.n_mo * q.nsets;
    double *buf = reinterpret_cast<double *>(cis_smem);                       // [set][mo][TP]
    const uint32_t a_buf = smem_u32(buf);
    const int seg16 = TP >> 1, ncopy = nrow * seg16;
    const int grp = warp / q.ncw, comp = warp - grp * q.ncw;                  // 32-point group, component of this warp
    const size_t dstr = (size_t)q.n_mo * TP;
    const double *mo = buf + grp * 32 + lane;
    const double *dd = mo + dstr + (size_t)comp * dstr;                       // derivative block of this warp's component
    for (long long tile = blockIdx.x; tile < q.ntiles; tile += gridDim.x) {
        const long long x0 = tile * TP;
        __syncthreads();
        for (int e = tid; e < ncopy; e += nthr) {
            const int r = e / seg16, piece = e - r * seg16;
            const int sset = r / q.n_mo, m = r - sset * q.n_mo;
            const double *src = (sset == 0 ? p.mo : p.dmo + (size_t)(sset - 1) * p.dstride) + (size_t)m * p.ld + x0 + 2 * piece;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a_buf + (uint32_t)e * 16u), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        double acc = 0.0;
        const double *trec = reinterpret_cast<const double *>(p.tp);
#pragma unroll 4
        for (int e = 0; e < p.n_terms; ++e) {
            // two 8-byte loads: a warp-uniform 16-byte load was measured slower here (jab 96 against 85 ms)
            const double c = __ldg(trec + 2 * e), ab = __ldg(trec + 2 * e + 1);
            const int ra = __double2loint(ab) * TP, rb = __double2hiint(ab) * TP;
            if (MODE == CI_RHO) {
                acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(c, mo[ra]), mo[rb]));                        // cy_ci.pyx:88,95
            } else if (MODE == CI_ANB) {
                acc = __dadd_rn(acc, __dmul_rn(c, __dmul_rn(mo[ra], dd[rb])));                        // cy_ci.pyx:236-238
            } else {
                const double t = __dadd_rn(__dmul_rn(mo[ra], dd[rb]), -__dmul_rn(mo[rb], dd[ra]));
                if (MODE == CI_JAB) acc = __dadd_rn(acc, -__dmul_rn(0.5, __dmul_rn(c, t)));           // cy_ci.pyx:181-184
                else acc = __dadd_rn(acc, __dmul_rn(c, t));                                           // cy_ci.pyx:199-201
            }
        }
        p.out[(size_t)comp * p.ldo + x0 + grp * 32 + lane] = acc;
    }
}

// Measured alternatives (not kept).  Round 1: a tiled kernel that stages all n_mo row segments of a 24-32 point tile in
// shared memory so that HBM delivers every MO value exactly once was 2x (rho) to 4x (jab) SLOWER than the gather kernel.
// Round 2 (profiles/r02_ci_staged_vs_gather.txt): the same idea with one private tile per WARP, bulk-async copies issued by
// all lanes and the warps of a CTA in different phases (loading / summing): bit-identical, DRAM traffic = algorithmic, but
// 5.55 ms against 2.25 ms (rho) and 59.7 against 9.5 ms (jab) for 884 736 points, 500 MOs, 1000 pairs.  Bit-identical results
// require the sequential term order per point, i.e. parallelism only ACROSS points, and shared memory caps the points in
// flight at ~48 per SM (rho; 12 for the four sets of jab): ~90 cycles per term of exposed load + FP64 latency cannot be hidden
// by three warps, while the gather kernel keeps 2048 points per SM in flight.  The gather kernel runs at 85% of the HBM
// bandwidth but moves 3.5x the algorithmic bytes (ncu: 12.5 GB read for 3.5 GB of MO values; L1/L2 hit rates 1% / 8%): the
// MO rows are re-fetched per term because n_mo KB per CTA times the resident CTAs exceeds the L2.  In the fused path
// (okb_eval_ci) this kernel is < 8% of the time, the MO evaluation dominates.

}  // namespace okb
