// okb_text.cuh -- Gaussian cube text on the device: the data loop of output/cube.py:86-96 (cube_creator).
//
// The reference writes, for every (x, y) pair, the values along z (all data sets of a point next to each other) as
// '%.5E' right-justified in 13 columns, a newline behind every 6th value of the row and one more at the end of the row:
//
//     for rr: for ss: c = 0
//         for tt: for dd in data[:, rr, ss, tt]:  string += ('%.5E' % dd).rjust(13); if c % 6 == 5: string += '\n'; c += 1
//         string += '\n'
//
// '%.5E' of a double never needs more than 13 characters ("-1.23457E-308"), so every value occupies exactly 13 bytes and
// the byte offset of every value is known in closed form: a row of n = Nz * n_sets values takes 13 n + n/6 + 1 bytes.  One
// CTA formats CUBE_VPB consecutive values into shared memory and copies that contiguous piece of the text to global memory
// with aligned 16-byte stores.
//
// '%.5E' is CORRECTLY ROUNDED (round-half-even on the exact binary value), as Python's float formatting is:
//   fast path   y = |v| * 10^(5-k) in double-double arithmetic (table of 10^j as (hi, lo) pairs, okb_pow10.h; relative
//               error < 2^-100, i.e. < 2^-80 absolute for y in [1e5, 1e6)); the six digits are floor(y) or floor(y) + 1
//               depending on the fraction, which is decided from the double-double value whenever it is further than
//               2^-70 from one half;
//   exact path  otherwise (exact ties such as 100000.5, or -- never observed -- a value closer than 2^-70 to a tie):
//               compare  m 2^e 10^(5-k)  with  floor(y) + 1/2  in multi-word integer arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "okb_pow10.h"

namespace okb {

constexpr int CUBE_NT = 256;                  // threads per CTA
constexpr int CUBE_VPT = 4;                   // values per thread
constexpr int CUBE_VPB = CUBE_NT * CUBE_VPT;  // values per CTA
constexpr int CUBE_W = 13;                    // characters per value
// worst case per value: 13 characters + one newline behind every value (rows of one value carry two: n/6 = 0, +1 ... see
// cube_row_bytes) -> 15 bytes; + 16 bytes of alignment slack in front
constexpr int CUBE_SMEM = CUBE_VPB * 15 + 32;

__device__ double g_pow10[2 * OKB_POW10_JMAX + 1][2];      // filled from OKB_POW10_TABLE at context creation

struct CubeParams {
    const double *data;        // [n_sets][nrows][nz]
    long long set_stride;      // doubles between two data sets
    long long nrows;           // Nx * Ny rows of this launch
    int nz, n_sets;
    long long row0;            // first row (absolute), only for addressing `data`: data + (row - row0) * nz
    char *text;                // output, byte 0 = first byte of row `row0`
    long long total_values;    // nrows * nz * n_sets
};

__host__ __device__ inline long long cube_row_bytes(long long n) { return 13 * n + n / 6 + 1; }

// ---- exact comparison (slow path) --------------------------------------------------------------------------------
struct Big {
    uint32_t w[40];
    int n;
};
__device__ inline void big_set(Big &b, uint64_t v) {
    b.w[0] = (uint32_t)v;
    b.w[1] = (uint32_t)(v >> 32);
    b.n = b.w[1] ? 2 : 1;
}
__device__ inline void big_mul(Big &b, uint32_t c) {
    uint64_t carry = 0;
    for (int i = 0; i < b.n; ++i) {
        const uint64_t t = (uint64_t)b.w[i] * c + carry;
        b.w[i] = (uint32_t)t;
        carry = t >> 32;
    }
    if (carry && b.n < 40) b.w[b.n++] = (uint32_t)carry;
}
__device__ inline void big_pow5(Big &b, int p) {
    for (; p >= 13; p -= 13) big_mul(b, 1220703125u);          // 5^13
    uint32_t c = 1;
    for (; p > 0; --p) c *= 5u;
    if (c > 1) big_mul(b, c);
}
__device__ inline void big_shl(Big &b, int s) {
    const int ws = s >> 5, bs = s & 31;
    if (bs) {
        uint32_t carry = 0;
        for (int i = 0; i < b.n; ++i) {
            const uint32_t t = b.w[i];
            b.w[i] = (t << bs) | carry;
            carry = t >> (32 - bs);
        }
        if (carry && b.n < 40) b.w[b.n++] = carry;
    }
    if (ws) {
        for (int i = min(b.n - 1, 39 - ws); i >= 0; --i) b.w[i + ws] = b.w[i];
        for (int i = 0; i < ws; ++i) b.w[i] = 0;
        b.n = min(b.n + ws, 40);
    }
}
__device__ inline int big_cmp(const Big &a, const Big &b) {
    if (a.n != b.n) return a.n > b.n ? 1 : -1;
    for (int i = a.n - 1; i >= 0; --i)
        if (a.w[i] != b.w[i]) return a.w[i] > b.w[i] ? 1 : -1;
    return 0;
}
// sign of  x * 10^j - (q + 1/2)  for x = m * 2^e > 0 (m < 2^53), q < 2^20
__device__ __noinline__ int cube_exact_cmp(uint64_t m, int e, int j, uint32_t q) {
    // m 2^(e+1) 2^j 5^j  vs  (2q + 1): negative exponents move to the other side
    const int a2 = e + 1 + j, a5 = j;
    Big L, R;
    big_set(L, m);
    big_set(R, 2ull * q + 1ull);
    if (a5 >= 0) big_pow5(L, a5); else big_pow5(R, -a5);
    if (a2 >= 0) big_shl(L, a2); else big_shl(R, -a2);
    return big_cmp(L, R);
}

// ---- '%.5E' -----------------------------------------------------------------------------------------------------------
// six significant digits N in [100000, 999999] and the decimal exponent k of x > 0 (finite), correctly rounded
__device__ inline void cube_digits(double x, uint32_t &N, int &k) {
    // decimal exponent estimate from the binary exponent: x in [2^e, 2^(e+1)) -> k in {floor(e log10 2), that + 1}
    // (78913 / 2^18 = 0.301029...; exact floor for |e| <= 1100), corrected by the range check below
    const uint64_t xb = (uint64_t)__double_as_longlong(x);
    const int eb = (int)((xb >> 52) & 0x7ff);
    const int e2 = eb ? eb - 1023 : -1011 - __clzll((long long)(xb & 0xfffffffffffffull));
    k = (e2 * 78913) >> 18;
    double s, t;
    for (int it = 0; it < 3; ++it) {
        int j = 5 - k;
        // y = x * 10^j as (s, t), in one or two double-double multiplications
        int j1 = j, j2 = 0;
        if (j > OKB_POW10_JMAX || j < -OKB_POW10_JMAX) {
            j1 = j / 2;
            j2 = j - j1;
        }
        const double ph = g_pow10[j1 + OKB_POW10_JMAX][0], pl = g_pow10[j1 + OKB_POW10_JMAX][1];
        double p = x * ph;
        double err = fma(x, ph, -p) + x * pl;
        s = p + err;
        t = err - (s - p);
        if (j2 != 0) {
            const double qh = g_pow10[j2 + OKB_POW10_JMAX][0], ql = g_pow10[j2 + OKB_POW10_JMAX][1];
            p = s * qh;
            err = fma(s, qh, -p) + (s * ql + t * qh);
            s = p + err;
            t = err - (s - p);
        }
        if (s >= 1e6) { ++k; continue; }
        if (s < 1e5) { --k; continue; }
        break;
    }
    // s in [1e5, 1e6): integer part and the distance of the fraction from one half
    const double yi = floor(s);
    const double g = (s - yi) - 0.5;                         // exact: multiples of ulp(s) >= 2^-36, |g| <= 1/2
    uint32_t q = (uint32_t)yi;
    int up;
    if (g != 0.0) up = g > 0.0;                              // |g| >= ulp(s) > |t|: the sign of g + t is the sign of g
    else if (fabs(t) > 0x1p-70) up = t > 0.0;
    else {
        // (nearly) a tie: decide exactly; ties go to the even digit string
        const uint64_t bits = (uint64_t)__double_as_longlong(x);
        const int be = (int)((bits >> 52) & 0x7ff);
        const uint64_t m = be ? ((bits & 0xfffffffffffffull) | (1ull << 52)) : (bits & 0xfffffffffffffull);
        const int e = (be ? be : 1) - 1075;
        const int c = cube_exact_cmp(m, e, 5 - k, q);
        up = c > 0 || (c == 0 && (q & 1u));
    }
    q += up;
    if (q >= 1000000u) {
        q = 100000u;
        ++k;
    }
    N = q;
}

// ('%.5E' % v).rjust(13) into 13 bytes
__device__ inline void cube_format(double v, char *o) {
    const uint64_t bits = (uint64_t)__double_as_longlong(v);
    const bool neg = (bits >> 63) != 0;
    const double x = fabs(v);
#pragma unroll
    for (int i = 0; i < CUBE_W; ++i) o[i] = ' ';
    if (!(x <= 1.7976931348623157e308)) {                    // inf / nan
        if (x != x) {                                        // Python prints NAN without a sign
            o[10] = 'N'; o[11] = 'A'; o[12] = 'N';
        } else {
            o[10] = 'I'; o[11] = 'N'; o[12] = 'F';
            if (neg) o[9] = '-';
        }
        return;
    }
    uint32_t N = 0;
    int k = 0;
    if (x != 0.0) cube_digits(x, N, k);
    // d.dddddE+XX (11 characters) or d.dddddE+XXX (12)
    const bool kneg = k < 0;
    const int ka = kneg ? -k : k;
    int pos = CUBE_W - 1;
    if (ka >= 100) {
        o[pos--] = '0' + ka % 10;
        o[pos--] = '0' + (ka / 10) % 10;
        o[pos--] = '0' + ka / 100;
    } else {
        o[pos--] = '0' + ka % 10;
        o[pos--] = '0' + ka / 10;
    }
    o[pos--] = kneg ? '-' : '+';
    o[pos--] = 'E';
    uint32_t n = N;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        o[pos--] = '0' + n % 10;
        n /= 10;
    }
    o[pos--] = '.';
    o[pos--] = '0' + n;
    if (neg) o[pos] = '-';
}

__global__ void __launch_bounds__(CUBE_NT) okb_cube_kernel(const CubeParams p) {
    extern __shared__ __align__(16) char sm[];
    const unsigned n = (unsigned)p.nz * (unsigned)p.n_sets;  // values per row (< 2^31, checked by the host)
    const long long rb = cube_row_bytes(n);
    const long long v0 = (long long)blockIdx.x * CUBE_VPB;
    const int cnt = (int)min((long long)CUBE_VPB, p.total_values - v0);
    // the only 64-bit division: row and position of this CTA's first value; everything else is relative to it
    const long long row0 = v0 / n;
    const unsigned c0 = (unsigned)(v0 - row0 * n);
    const long long in_row0 = 13ll * c0 + c0 / 6;            // offset of the first value inside its row
    const long long b0 = row0 * rb + in_row0;
    // end of this CTA's piece: start of the next CTA's first value, or the end of the text
    const unsigned ce = c0 + (unsigned)cnt, dre = ce / n, cre = ce - dre * n;
    const long long nb = (v0 + cnt < p.total_values ? (long long)dre * rb + 13ll * cre + cre / 6
                                                    : (p.nrows - row0) * rb) - in_row0;
    const int lead = (int)((uintptr_t)(p.text + b0) & 15);   // the smem image is aligned like the global text
    char *img = sm + lead;
    const unsigned n_sets = (unsigned)p.n_sets;
    for (int i = threadIdx.x; i < cnt; i += CUBE_NT) {
        const unsigned cc = c0 + (unsigned)i, dr = cc / n, c = cc - dr * n;
        const unsigned tt = c / n_sets, dd = c - tt * n_sets;
        const double v = __ldg(p.data + dd * p.set_stride + (row0 + dr) * p.nz + tt);
        const unsigned c6 = c / 6;
        char *o = img + ((long long)dr * rb + 13ll * c + c6 - in_row0);
        cube_format(v, o);
        // newline(s) behind the value: after every 6th value of the row, and at the end of the row
        int w = CUBE_W;
        if (c - c6 * 6 == 5) o[w++] = '\n';
        if (c == n - 1) o[w++] = '\n';
    }
    __syncthreads();
    // copy the image [0, nb) to global memory: bytes up to the first 16-byte boundary, 16-byte body, byte tail
    char *dst = p.text + b0;
    const int head = (int)min((long long)((16 - lead) & 15), nb);
    if ((int)threadIdx.x < head) dst[threadIdx.x] = img[threadIdx.x];
    const long long body = (nb - head) / 16;
    const uint4 *s4 = reinterpret_cast<const uint4 *>(img + head);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    for (long long i = threadIdx.x; i < body; i += CUBE_NT) d4[i] = s4[i];
    const long long done = head + body * 16;
    if (done + (long long)threadIdx.x < nb) dst[done + threadIdx.x] = img[done + threadIdx.x];
}

}  // namespace okb
