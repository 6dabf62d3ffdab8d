// okb200.cu -- C ABI (include/okb200.h) + host runtime of the B200 grid path.
//
// Host responsibilities (all O(basis), never O(points)):
//   * flatten the reference's basis arrays (the argument list of cy_core.aocreator,
//     cy_core.pyx:51-61) into shell-sorted device tables: per shell the centre, per primitive
//     (alpha, c * N_L(alpha)) with the primitive norm of c_support.c:177-188 hoisted out of the point
//     loop, per Cartesian function (lx,ly,lz) and its angular norm 1/sqrt((2lx-1)!!(2ly-1)!!(2lz-1)!!);
//   * cut the shells into chunks of <= KC functions and pack one fixed-stride table blob per chunk
//     (a single bulk-async copy per chunk in the kernel);
//   * fold the Cartesian->spherical rows into the MO coefficients (C' = C T, the transform is linear)
//     and lay C' out as [mo-tile][chunk][KC][CS] (CS = MC padded to 4 mod 16) so that each (tile, chunk) is
//     one contiguous TMA copy and the DMMA fragment loads are bank-conflict free;
//   * slab the point range, launch, and stream results back to host buffers.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/okb200.h"
#include "okb_common.cuh"
#include "okb_variant.h"
#include "okb_shell.cuh"   // constexpr shell tables (templates are instantiated in inst_*.cu only)
#include "okb_ws.cuh"      // pad_stride
#include "okb_ao_zrun.cuh" // ZR_MAXL
#include "okb_misc.cuh"
#include "okb_ci.cuh"
#include "okb_td.cuh"
#include "okb_overlap.cuh"
#include "okb_text.cuh"

using namespace okb;

// ---- error channel ------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? OKB_ERR_NOMEM : OKB_ERR_CUDA,            \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char *okb_last_error(void) { return g_err; }
extern "C" int okb_version(void) { return 100; }
extern "C" int okb_device_count(int *n) {
    if (!n) return fail(OKB_ERR_ARG, "okb_device_count: null pointer");
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) {
        *n = 0;
        return fail(OKB_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return OKB_OK;
}

// ---- handles ----------------------------------------------------------------------------------------
struct okb_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_compute[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    long long launches = 0;
    long long h2d_bytes = 0, d2h_bytes = 0;   // bytes moved over PCIe by this context
    std::string last_kernel;
    void *slab[2] = {nullptr, nullptr};      // device staging for host outputs
    size_t slab_bytes = 0;
    double *norm_dev = nullptr;              // mo_norm accumulation
    size_t norm_cap = 0;
    void *ci_buf = nullptr;                  // detCI: term arrays + MO slab + output slab
    size_t ci_bytes = 0;
    void *rdm_buf = nullptr;                 // detCI, OKB_FLAG_CI_FAST: dense orbital-pair matrix + row list
    size_t rdm_bytes = 0;
    double *phi_buf = nullptr;               // rho + laplacian: MO values between the two passes
    size_t phi_bytes = 0;
    // Recycled device buffers of destroyed handles (chunk tables, coefficient tiles, axes, axis tables).
    // cudaFree sporadically takes 5-600 ms next to large page-locked host buffers (measured:
    // scripts/e2e_probe.py, e2e_probe2.py), so handle churn must not free.
    std::vector<std::pair<void *, size_t>> pool;
};

static int pool_get(okb_ctx *ctx, size_t bytes, void **out, size_t *got) {
    bytes = std::max<size_t>(bytes, 256);
    int best = -1;
    for (int i = 0; i < (int)ctx->pool.size(); ++i)
        if (ctx->pool[i].second >= bytes && ctx->pool[i].second <= 2 * bytes + ((size_t)64 << 10) &&
            (best < 0 || ctx->pool[i].second < ctx->pool[best].second))
            best = i;
    if (best >= 0) {
        *out = ctx->pool[best].first;
        *got = ctx->pool[best].second;
        ctx->pool.erase(ctx->pool.begin() + best);
        return OKB_OK;
    }
    CU(cudaMalloc(out, bytes));
    *got = bytes;
    return OKB_OK;
}
// The caller guarantees that no kernel still uses the buffer (the destroy functions drain the stream first:
// a recycled buffer may be overwritten by a synchronous copy that is not ordered with ctx->stream).
static void pool_put(okb_ctx *ctx, void *ptr, size_t bytes) {
    if (!ptr) return;
    if (ctx->pool.size() < 64 && bytes <= ((size_t)256 << 20)) {
        ctx->pool.emplace_back(ptr, bytes);
        return;
    }
    cudaFree(ptr);
}
struct DevShell {
    double c[3];
    int L;
    std::vector<int> prims;                  // indices into prim arrays
    std::vector<double> alpha, cn;
    std::vector<int> fn_row;                 // Cartesian row of each function
    std::vector<int> lx, ly, lz;
    std::vector<double> f;
};

// One way of cutting the device shells into chunks, with its packed chunk tables on the device.
// `cart`: every shell writes its Cartesian rows (SINK_AO applies the spherical rows while storing).
// `mix` : standard d/f/g shells of a spherical basis write their 2L+1 spherical rows themselves
//         (kind 2), so the contraction runs over n_sph instead of n_cart functions; all other shells
//         write Cartesian rows whose transform is folded into the MO coefficients.
struct KRow { int is_sph; int index; };      // what a tile row holds: spherical AO `index` or Cartesian row `index`
struct Layout {
    struct Chunk { int s0, s1, k0, nfn, nprim; };
    std::vector<Chunk> chunks;
    std::vector<KRow> krow;                  // device row -> content
    std::vector<int> shell_sph;              // per device shell: 1 if it writes spherical rows
    std::vector<int> order;                  // chunk c holds the device shells order[s0..s1)
    BlobLayout lay{};
    unsigned char *meta_dev = nullptr;
    size_t meta_bytes = 0;                   // pooled size of meta_dev
    int n_rows = 0;
};

struct okb_basis {
    okb_ctx *ctx = nullptr;
    int n_cart = 0, n_ao = 0;
    bool spherical = false;
    std::vector<DevShell> shells;
    std::vector<int> row_shell;              // Cartesian row -> device shell
    // cart -> sph CSR
    std::vector<int> t_ptr, t_col;
    std::vector<double> t_val;
    Layout cart, mix;
    bool mix_is_cart = true;                 // Cartesian basis: the two layouts coincide
    unsigned long long serial = 0;           // identifies the primitive set (axis-table cache key)
    std::vector<int> shell_gprim;            // device shell -> index of its first primitive (basis-wide)
    int n_prim_dev = 0;
    double *prim5_dev = nullptr;             // [n_prim_dev][5] = alpha, cN, X, Y, Z
    size_t prim5_bytes = 0;
    const Layout &contraction_layout() const { return mix_is_cart ? cart : mix; }
};

struct okb_mo {
    okb_ctx *ctx = nullptr;
    okb_basis *basis = nullptr;
    int n_mo = 0;
    std::vector<double> ccart;               // [n_mo][n_cart] in Cartesian rows (C' = C T)
    std::vector<double> csph;                // [n_mo][n_ao] as given (spherical bases only)
    std::vector<double> occ;
    struct Blob {
        double *c = nullptr, *occ = nullptr, *crem = nullptr;
        int n_mtile = 0;
        size_t c_bytes = 0, occ_bytes = 0, crem_bytes = 0;
    };
    std::map<int, Blob> blobs;               // keyed by 2*(16*MC + rem) + (mix layout ? 1 : 0)
};

struct okb_grid {
    okb_ctx *ctx = nullptr;
    int kind = 0;                            // 0 regular, 1 vector, 2 spherical product, 3 cylindrical product
    bool has_aff = false;                    // product grids: affine map applied to the generated coordinates
    double aff[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    int nx = 0, ny = 0, nz = 0;
    long long npts = 0;
    double *gx = nullptr, *gy = nullptr, *gz = nullptr;
    size_t gbytes[3] = {0, 0, 0};            // pooled sizes of gx, gy, gz
    bool owns = true;
    // regular grids: separable-exponential tables per basis (key: okb_basis::serial), [3 tables back to back]
    struct Tab { double *ptr; size_t bytes; };
    std::map<unsigned long long, Tab> axis_tabs;
};

// ---- context ------------------------------------------------------------------------------------------
extern "C" int okb_ctx_create(int device, okb_ctx **out) {
    if (!out) return fail(OKB_ERR_ARG, "okb_ctx_create: null output pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(OKB_ERR_CUDA, "no CUDA device available (%s): orbkit_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return fail(OKB_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(OKB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    device, prop.major, prop.minor);
    okb_ctx *c = new okb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&c->ev_compute[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_copy[i], cudaEventDisableTiming));
    }
    CU(cudaMemcpyToSymbol(g_pow10, OKB_POW10_TABLE, sizeof(OKB_POW10_TABLE)));   // '%.5E' formatter (okb_text.cuh)
    *out = c;
    return OKB_OK;
}

extern "C" int okb_ctx_destroy(okb_ctx *c) {
    if (!c) return OKB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < 2; ++i) {
        if (c->slab[i]) cudaFree(c->slab[i]);
        if (c->ev_compute[i]) cudaEventDestroy(c->ev_compute[i]);
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
    }
    if (c->norm_dev) cudaFree(c->norm_dev);
    if (c->ci_buf) cudaFree(c->ci_buf);
    if (c->rdm_buf) cudaFree(c->rdm_buf);
    if (c->phi_buf) cudaFree(c->phi_buf);
    for (auto &e : c->pool) cudaFree(e.first);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->copy_stream);
    delete c;
    return OKB_OK;
}

extern "C" int okb_ctx_sync(okb_ctx *c) {
    if (!c) return fail(OKB_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    return OKB_OK;
}
extern "C" void *okb_ctx_stream(okb_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int okb_ctx_launch_count(okb_ctx *c, long long *n) {
    if (!c || !n) return fail(OKB_ERR_ARG, "null pointer");
    *n = c->launches;
    return OKB_OK;
}
extern "C" int okb_ctx_traffic(okb_ctx *c, long long *h2d, long long *d2h) {
    if (!c) return fail(OKB_ERR_ARG, "null context");
    if (h2d) *h2d = c->h2d_bytes;
    if (d2h) *d2h = c->d2h_bytes;
    return OKB_OK;
}
extern "C" int okb_ctx_last_kernel(okb_ctx *c, char *buf, int buflen) {
    if (!c || !buf || buflen <= 0) return fail(OKB_ERR_ARG, "null pointer");
    snprintf(buf, buflen, "%s", c->last_kernel.c_str());
    return OKB_OK;
}

// ---- scalar helpers (host; the reference exposes them through cy_core.aonorm / aoxyz) ------------------
static int dfact(int n) {
    int r = 1;
    for (; n > 0; n -= 2) r *= n;
    return r;
}
static double radial_norm(int L, double alpha) {   // (2/pi)^(3/4) 2^L alpha^((2L+3)/4)
    return pow(2. / M_PI, 0.75) * (pow(2., (double)L) * pow(alpha, (2. * L + 3.) / 4.));
}
static double angular_norm(int lx, int ly, int lz) {
    return 1.0 / sqrt((double)(dfact(2 * lx - 1) * dfact(2 * ly - 1) * dfact(2 * lz - 1)));
}
extern "C" double okb_aonorm(int lx, int ly, int lz, double alpha, int is_normalized) {
    if (is_normalized > 0) return 1.0;
    return radial_norm(lx + ly + lz, alpha) /
           sqrt((double)(dfact(2 * lx - 1) * dfact(2 * ly - 1) * dfact(2 * lz - 1)));
}
static double hpow(double r, int l) {
    double a = 1.0;
    for (; l > 0; l >>= 1) {
        if (l & 1) a *= r;
        r *= r;
    }
    return a;
}
extern "C" double okb_aoxyz(double x, double y, double z, int lx, int ly, int lz, double alpha, int drv) {
    // derivative prefactor of x^lx y^ly z^lz exp(-alpha r^2), reference semantics (c_support.c:28-175),
    // written with the per-axis quantities the kernels use.
    const double r[3] = {x, y, z};
    const int l[3] = {lx, ly, lz};
    double q0[3], qm1[3], qp1[3], qm2[3];
    for (int a = 0; a < 3; ++a) {
        q0[a] = hpow(r[a], l[a]);
        qm1[a] = l[a] > 0 ? l[a] * hpow(r[a], l[a] - 1) : 0.0;
        qp1[a] = q0[a] * r[a];
        qm2[a] = l[a] > 1 ? (double)(l[a] * (l[a] - 1)) * hpow(r[a], l[a] - 2) : 0.0;
    }
    if (drv == 0) return q0[0] * q0[1] * q0[2];
    if (drv >= 1 && drv <= 3) {
        const int a = drv - 1, b = (a + 1) % 3, c = (a + 2) % 3;
        return q0[b] * q0[c] * (qm1[a] - 2.0 * alpha * qp1[a]);
    }
    if (drv >= 4 && drv <= 6) {
        const int a = drv - 4, b = (a + 1) % 3, c = (a + 2) % 3;
        return q0[b] * q0[c] *
               (q0[a] * (4.0 * alpha * alpha * r[a] * r[a] - 2.0 * alpha * (2 * l[a] + 1)) + qm2[a]);
    }
    if (drv >= 7 && drv <= 9) {
        const int a = drv == 9 ? 1 : 0, b = drv == 7 ? 1 : 2, c = 3 - a - b;
        double B = 0.0;
        if (l[a] > 0 || l[b] > 0) B = (l[a] > 0 ? qm1[a] : 1.0) * (l[b] > 0 ? qm1[b] : 1.0);
        return q0[c] * (4.0 * alpha * alpha * qp1[a] * qp1[b] + B);
    }
    return 0.0;
}

// ---- basis ----------------------------------------------------------------------------------------------
// functions in the default (Molden) order of tools.exp[L]?  Those shells take the straight-line AO code.
static bool shell_is_standard(const DevShell &sh) {
    if (sh.L < 0 || sh.L > 4 || (int)sh.fn_row.size() != std_nfn(sh.L)) return false;
    for (int j = 0; j < std_nfn(sh.L); ++j) {
        const int e = std_lxyz(sh.L, j);
        if (sh.lx[j] != (e & 15) || sh.ly[j] != ((e >> 4) & 15) || sh.lz[j] != ((e >> 8) & 15)) return false;
    }
    return true;
}

// Spherical rows of a standard shell: which spherical AOs (CSR rows) are built from exactly this shell's
// Cartesian functions with the compiled-in term pattern of canonical row r?  Returns false (-> fold the
// transform into the coefficients instead) unless all 2L+1 canonical rows are matched exactly once.
static bool match_sph_rows(const okb_basis *b, int s, std::vector<int> &sph_of_canon) {
    const DevShell &sh = b->shells[s];
    const int L = sh.L;
    if (!b->spherical || L < 2 || L > 4 || !shell_is_standard(sh)) return false;
    sph_of_canon.assign(2 * L + 1, -1);
    int found = 0;
    for (int j = 0; j < b->n_ao; ++j) {
        const int t0 = b->t_ptr[j], t1 = b->t_ptr[j + 1];
        if (t1 <= t0 || b->row_shell[b->t_col[t0]] != s) continue;
        int hit = -1;
        for (int r = 0; r < 2 * L + 1 && hit < 0; ++r) {
            if (sph_nterm(L, r) != t1 - t0 || sph_of_canon[r] >= 0) continue;
            bool same = true;
            for (int t = 0; t < t1 - t0 && same; ++t)
                same = (b->t_col[t0 + t] == sh.fn_row[sph_cart(L, r, t)]);
            if (same) hit = r;
        }
        if (hit < 0) return false;          // a spherical function of this shell with an unknown pattern
        sph_of_canon[hit] = j;
        ++found;
    }
    // every Cartesian function of the shell must be used only by these rows
    for (int j = 0; j < b->n_ao; ++j)
        for (int t = b->t_ptr[j]; t < b->t_ptr[j + 1]; ++t)
            if (b->row_shell[b->t_col[t]] == s) {
                bool mine = false;
                for (int r = 0; r < 2 * L + 1; ++r) mine |= (sph_of_canon[r] == j);
                if (!mine) return false;
            }
    return found == 2 * L + 1;
}

static void layout_free(okb_ctx *ctx, Layout &lo) {
    if (lo.meta_dev) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        pool_put(ctx, lo.meta_dev, lo.meta_bytes);
    }
    lo = Layout();
}

static int layout_build(okb_basis *b, bool sph_out, Layout &lo) {
    layout_free(b->ctx, lo);
    const int nshell = (int)b->shells.size();
    // ---- rows written by every shell -------------------------------------------------------------
    std::vector<std::vector<int>> canon(nshell);           // kind-2 shells: spherical AO of canonical row r
    lo.shell_sph.assign(nshell, 0);
    for (int s = 0; s < nshell; ++s)
        if (sph_out && match_sph_rows(b, s, canon[s])) lo.shell_sph[s] = 1;
    // ---- chunking -----------------------------------------------------------------------------------------
    // Shells are taken in the caller's order; when the next one no longer fits, the remaining room of
    // the chunk is filled with the largest later shells that still fit (`order` records the resulting
    // permutation; chunk c holds order[s0..s1)).  The contraction pads every chunk to a multiple of 4
    // rows and pays a fixed hand-over cost per chunk, so full chunks are worth 4-5% on the 1000-AO
    // molecule (38 chunks / 1045 padded rows in plain order -> 32 chunks / 1000 rows).
    const int MAXS = 32, MAXP = 96;
    std::vector<int> nf_of(nshell), np_of(nshell);
    for (int s = 0; s < nshell; ++s) {
        const DevShell &sh = b->shells[s];
        nf_of[s] = lo.shell_sph[s] ? 2 * sh.L + 1 : (int)sh.fn_row.size();
        np_of[s] = (int)sh.alpha.size();
        if (nf_of[s] > KC) return fail(OKB_ERR_UNSUPPORTED, "shell with %d functions exceeds the chunk size %d", nf_of[s], KC);
        if (np_of[s] > MAXP) return fail(OKB_ERR_UNSUPPORTED, "shell with %d primitives exceeds the chunk limit %d", np_of[s], MAXP);
    }
    static const bool plain_order = getenv("OKB_PLAIN_CHUNKS") != nullptr;
    lo.order.clear();
    {
        std::vector<char> used(nshell, 0);
        int first = 0;                                   // first unused shell
        while (first < nshell) {
            Layout::Chunk cur{(int)lo.order.size(), (int)lo.order.size(), 0, 0, 0};
            auto fits = [&](int s) {
                return cur.nfn + nf_of[s] <= KC && cur.s1 - cur.s0 < MAXS && cur.nprim + np_of[s] <= MAXP;
            };
            auto take = [&](int s) {
                used[s] = 1;
                lo.order.push_back(s);
                cur.s1++;
                cur.nfn += nf_of[s];
                cur.nprim += np_of[s];
            };
            int s = first;
            for (; s < nshell; ++s) {                    // the caller's order while it fits
                if (used[s]) continue;
                if (!fits(s)) break;
                take(s);
            }
            while (!plain_order && cur.nfn < KC) {       // fill the remaining room from later shells
                int best = -1;
                for (int t = s; t < nshell; ++t)
                    if (!used[t] && fits(t) && (best < 0 || nf_of[t] > nf_of[best])) best = t;
                if (best < 0) break;
                take(best);
            }
            while (first < nshell && used[first]) ++first;
            // most expensive shells first: the producer warps deal the items of a chunk out in this order
            // (cost ~ exponentials + per-function polynomial work)
            auto cost = [&](int t) { return 5 * np_of[t] + 4 * (int)b->shells[t].fn_row.size(); };
            if (!plain_order)
                std::stable_sort(lo.order.begin() + cur.s0, lo.order.begin() + cur.s1,
                                 [&](int a, int c) { return cost(a) > cost(c); });
            lo.chunks.push_back(cur);
        }
    }
    int k = 0;
    std::vector<int> shell_k0(nshell, 0);
    for (Layout::Chunk &ch : lo.chunks) {
        ch.k0 = k;
        for (int o = ch.s0; o < ch.s1; ++o) {
            const int s = lo.order[o];
            const DevShell &sh = b->shells[s];
            shell_k0[s] = k;
            if (lo.shell_sph[s]) {
                // rows in ascending spherical AO index (keeps the coefficient rows in the caller's order)
                std::vector<int> js(canon[s]);
                std::sort(js.begin(), js.end());
                for (int j : js) lo.krow.push_back(KRow{1, j});
            } else {
                for (int r : sh.fn_row) lo.krow.push_back(KRow{0, r});
            }
            k += nf_of[s];
        }
    }
    lo.n_rows = k;
    const int nchunk = (int)lo.chunks.size();
    // Cartesian row -> (chunk, chunk-local k) for the SINK_AO output rows (cart layout only)
    std::vector<int> fn_chunk(b->n_cart, -1), fn_klocal(b->n_cart, -1);
    for (int c = 0; c < nchunk; ++c)
        for (int kk = 0; kk < lo.chunks[c].nfn; ++kk) {
            const KRow &kr = lo.krow[lo.chunks[c].k0 + kk];
            if (!kr.is_sph) {
                fn_chunk[kr.index] = c;
                fn_klocal[kr.index] = kk;
            }
        }
    // ---- output rows per chunk (used by SINK_AO only; the cart layout) ---------------------------------------
    std::vector<std::vector<RowMeta>> rows(nchunk);
    std::vector<std::vector<TermMeta>> terms(nchunk);
    if (sph_out) {
        // every spherical AO is either a tile row of a kind-2 shell or a combination of the Cartesian rows
        // of another shell
        std::vector<int> sph_chunk(b->n_ao, -1), sph_klocal(b->n_ao, -1);
        for (int c = 0; c < nchunk; ++c)
            for (int kk = 0; kk < lo.chunks[c].nfn; ++kk) {
                const KRow &kr = lo.krow[lo.chunks[c].k0 + kk];
                if (kr.is_sph) {
                    sph_chunk[kr.index] = c;
                    sph_klocal[kr.index] = kk;
                }
            }
        for (int j = 0; j < b->n_ao; ++j) {
            if (sph_chunk[j] >= 0) {
                const int c = sph_chunk[j];
                rows[c].push_back(RowMeta{j, (int)terms[c].size(), 1, 0});
                terms[c].push_back(TermMeta{sph_klocal[j], 0, 1.0});
                continue;
            }
            const int t0 = b->t_ptr[j], t1 = b->t_ptr[j + 1];
            if (t1 <= t0) return fail(OKB_ERR_ARG, "spherical function %d has no Cartesian terms", j);
            const int c = fn_chunk[b->t_col[t0]];
            if (c < 0) return fail(OKB_ERR_UNSUPPORTED, "spherical function %d refers to a transformed shell", j);
            rows[c].push_back(RowMeta{j, (int)terms[c].size(), t1 - t0, 0});
            for (int t = t0; t < t1; ++t) {
                if (fn_chunk[b->t_col[t]] != c)
                    return fail(OKB_ERR_UNSUPPORTED,
                                "spherical function %d mixes Cartesian functions of different shells", j);
                terms[c].push_back(TermMeta{fn_klocal[b->t_col[t]], 0, b->t_val[t]});
            }
        }
    } else {
        if (!b->spherical) {
            for (int c = 0; c < nchunk; ++c)
                for (int kk = 0; kk < lo.chunks[c].nfn; ++kk) {
                    rows[c].push_back(RowMeta{lo.krow[lo.chunks[c].k0 + kk].index, (int)terms[c].size(), 1, 0});
                    terms[c].push_back(TermMeta{kk, 0, 1.0});
                }
        } else {
            for (int j = 0; j < b->n_ao; ++j) {
                const int t0 = b->t_ptr[j], t1 = b->t_ptr[j + 1];
                if (t1 <= t0) return fail(OKB_ERR_ARG, "spherical function %d has no Cartesian terms", j);
                const int c = fn_chunk[b->t_col[t0]];
                rows[c].push_back(RowMeta{j, (int)terms[c].size(), t1 - t0, 0});
                for (int t = t0; t < t1; ++t) {
                    if (fn_chunk[b->t_col[t]] != c)
                        return fail(OKB_ERR_UNSUPPORTED,
                                    "spherical function %d mixes Cartesian functions of different shells", j);
                    terms[c].push_back(TermMeta{fn_klocal[b->t_col[t]], 0, b->t_val[t]});
                }
            }
        }
    }
    // ---- rows grouped by the shell they are built from (the z-run kernel walks shell by shell) ------------------
    std::vector<std::vector<int>> shell_row_off(nchunk), shell_nrow(nchunk);
    for (int c = 0; c < nchunk; ++c) {
        const Layout::Chunk &ch = lo.chunks[c];
        const int ns = ch.s1 - ch.s0;
        std::vector<int> fn_shell(ch.nfn, 0);
        for (int o = ch.s0; o < ch.s1; ++o) {
            const int s = lo.order[o];
            for (int kk = shell_k0[s] - ch.k0; kk < shell_k0[s] - ch.k0 + nf_of[s]; ++kk) fn_shell[kk] = o - ch.s0;
        }
        for (RowMeta &rm : rows[c]) rm.shell = fn_shell[terms[c][rm.term_off].k];
        std::stable_sort(rows[c].begin(), rows[c].end(), [](const RowMeta &a, const RowMeta &b) { return a.shell < b.shell; });
        shell_row_off[c].assign(ns, 0);
        shell_nrow[c].assign(ns, 0);
        for (const RowMeta &rm : rows[c]) shell_nrow[c][rm.shell]++;
        for (int q = 1; q < ns; ++q) shell_row_off[c][q] = shell_row_off[c][q - 1] + shell_nrow[c][q - 1];
    }
    // ---- aux records of the kind-2 shells ----------------------------------------------------------------------
    std::vector<std::vector<double>> aux(nchunk);
    std::vector<int> shell_aux(nshell, 0);
    for (int c = 0; c < nchunk; ++c)
        for (int o = lo.chunks[c].s0; o < lo.chunks[c].s1; ++o) {
            const int s = lo.order[o];
            if (!lo.shell_sph[s]) continue;
            const DevShell &sh = b->shells[s];
            const int L = sh.L;
            shell_aux[s] = (int)aux[c].size();
            for (int j = 0; j < std_nfn(L); ++j) aux[c].push_back(sh.f[j]);
            std::vector<int> js(canon[s]);
            std::sort(js.begin(), js.end());
            for (int r = 0; r < 2 * L + 1; ++r) {
                const int j = canon[s][r];
                const int pos = (int)(std::find(js.begin(), js.end(), j) - js.begin());
                aux[c].push_back((double)pos);
                for (int t = b->t_ptr[j]; t < b->t_ptr[j + 1]; ++t) aux[c].push_back(b->t_val[t]);
            }
        }
    // ---- pack -------------------------------------------------------------------------------------------------------
    int maxS = 1, maxP = 1, maxR = 1, maxT = 1, maxA = 2;
    for (int c = 0; c < nchunk; ++c) {
        maxS = std::max(maxS, lo.chunks[c].s1 - lo.chunks[c].s0);
        maxP = std::max(maxP, lo.chunks[c].nprim);
        maxR = std::max(maxR, (int)rows[c].size());
        maxT = std::max(maxT, (int)terms[c].size());
        maxA = std::max(maxA, (int)aux[c].size());
    }
    BlobLayout &L = lo.lay;
    auto up16 = [](int v) { return (v + 15) / 16 * 16; };
    L.off_shell = 16;
    L.off_prim = up16(L.off_shell + maxS * (int)sizeof(ShellMeta));
    L.off_fn = L.off_prim + maxP * 16;
    L.off_row = L.off_fn + KC * (int)sizeof(FnMeta);
    L.off_term = L.off_row + maxR * (int)sizeof(RowMeta);
    L.off_aux = L.off_term + maxT * (int)sizeof(TermMeta);
    L.stride = (L.off_aux + maxA * 8 + 127) / 128 * 128;
    if (L.stride > 24 * 1024)
        return fail(OKB_ERR_UNSUPPORTED, "chunk table of %d bytes is too large", L.stride);
    std::vector<unsigned char> blob((size_t)L.stride * std::max(nchunk, 1), 0);
    for (int c = 0; c < nchunk; ++c) {
        unsigned char *mb = blob.data() + (size_t)c * L.stride;
        const Layout::Chunk &ch = lo.chunks[c];
        ChunkHdr hdr{ch.s1 - ch.s0, ch.nprim, ch.nfn, (int)rows[c].size()};
        memcpy(mb, &hdr, sizeof(hdr));
        ShellMeta *sm = reinterpret_cast<ShellMeta *>(mb + L.off_shell);
        double2 *pm = reinterpret_cast<double2 *>(mb + L.off_prim);
        FnMeta *fm = reinterpret_cast<FnMeta *>(mb + L.off_fn);
        int po = 0;
        for (int o = ch.s0; o < ch.s1; ++o) {
            const int s = lo.order[o];
            const DevShell &sh = b->shells[s];
            const int fo = shell_k0[s] - ch.k0;
            ShellMeta m{};
            m.cx = sh.c[0]; m.cy = sh.c[1]; m.cz = sh.c[2];
            m.prim_off = po; m.nprim = (int)sh.alpha.size();
            m.fn_off = fo;
            m.L = sh.L;
            m.gprim = b->shell_gprim[s];
            m.row_off = shell_row_off[c][o - ch.s0];
            m.nrow = shell_nrow[c][o - ch.s0];
            if (lo.shell_sph[s]) {
                m.kind = 2;
                m.nfn = 2 * sh.L + 1;
                m.aux_off = shell_aux[s];
            } else {
                m.kind = shell_is_standard(sh) ? 1 : 0;
                m.nfn = (int)sh.fn_row.size();
                for (size_t j = 0; j < sh.fn_row.size(); ++j)
                    fm[fo + j] = FnMeta{sh.lx[j] | (sh.ly[j] << 8) | (sh.lz[j] << 16), 0, sh.f[j]};
            }
            sm[o - ch.s0] = m;
            for (size_t i = 0; i < sh.alpha.size(); ++i) pm[po++] = make_double2(sh.alpha[i], sh.cn[i]);
        }
        if (!rows[c].empty()) memcpy(mb + L.off_row, rows[c].data(), rows[c].size() * sizeof(RowMeta));
        if (!terms[c].empty()) memcpy(mb + L.off_term, terms[c].data(), terms[c].size() * sizeof(TermMeta));
        if (!aux[c].empty()) memcpy(mb + L.off_aux, aux[c].data(), aux[c].size() * sizeof(double));
    }
    CU(cudaSetDevice(b->ctx->device));
    {
        void *raw = nullptr;
        int rcp = pool_get(b->ctx, blob.size(), &raw, &lo.meta_bytes);
        if (rcp != OKB_OK) return rcp;
        lo.meta_dev = reinterpret_cast<unsigned char *>(raw);
    }
    CU(cudaMemcpy(lo.meta_dev, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    b->ctx->h2d_bytes += (long long)blob.size();
    return OKB_OK;
}

// (re)build both layouts after the basis or its spherical transform changed
static int basis_upload(okb_basis *b) {
    static unsigned long long next_serial = 1;
    b->serial = next_serial++;
    {
        std::vector<double> prim5;
        b->shell_gprim.assign(b->shells.size(), 0);
        for (size_t s = 0; s < b->shells.size(); ++s) {
            const DevShell &sh = b->shells[s];
            b->shell_gprim[s] = (int)(prim5.size() / 5);
            for (size_t i = 0; i < sh.alpha.size(); ++i) {
                const double q[5] = {sh.alpha[i], sh.cn[i], sh.c[0], sh.c[1], sh.c[2]};
                prim5.insert(prim5.end(), q, q + 5);
            }
        }
        b->n_prim_dev = (int)(prim5.size() / 5);
        CU(cudaSetDevice(b->ctx->device));
        pool_put(b->ctx, b->prim5_dev, b->prim5_bytes);
        b->prim5_dev = nullptr;
        {
            void *raw = nullptr;
            int rcp = pool_get(b->ctx, std::max<size_t>(prim5.size(), 1) * sizeof(double), &raw, &b->prim5_bytes);
            if (rcp != OKB_OK) return rcp;
            b->prim5_dev = reinterpret_cast<double *>(raw);
        }
        CU(cudaMemcpy(b->prim5_dev, prim5.data(), prim5.size() * sizeof(double), cudaMemcpyHostToDevice));
        b->ctx->h2d_bytes += (long long)(prim5.size() * sizeof(double));
    }
    b->row_shell.assign(b->n_cart, -1);
    for (int s = 0; s < (int)b->shells.size(); ++s)
        for (int r : b->shells[s].fn_row) b->row_shell[r] = s;
    int rc = layout_build(b, false, b->cart);
    if (rc != OKB_OK) return rc;
    b->mix_is_cart = true;
    if (b->spherical && !getenv("OKB_NO_SPH_ROWS")) {
        rc = layout_build(b, true, b->mix);
        if (rc != OKB_OK) return rc;
        bool any = false;
        for (int v : b->mix.shell_sph) any |= (v != 0);
        if (any) b->mix_is_cart = false;
        else layout_free(b->ctx, b->mix);
    }
    return OKB_OK;
}

extern "C" int okb_basis_create(okb_ctx *ctx, const int *lxlylz, const int *assign, const double *ao_coeffs,
                                const int *pnum_list, const double *geo_spec, const int *atom_indices,
                                int n_cont, int n_cart, int n_prim, int n_atoms, int is_normalized,
                                const double *renorm, okb_basis **out) {
    if (!ctx || !out) return fail(OKB_ERR_ARG, "okb_basis_create: null context/output");
    *out = nullptr;
    if (n_cont <= 0 || n_cart <= 0 || n_prim <= 0 || n_atoms <= 0)
        return fail(OKB_ERR_ARG, "okb_basis_create: empty basis");
    if (!lxlylz || !assign || !ao_coeffs || !pnum_list || !geo_spec || !atom_indices)
        return fail(OKB_ERR_ARG, "okb_basis_create: null array");
    okb_basis *b = new okb_basis();
    b->ctx = ctx;
    b->n_cart = n_cart;
    b->n_ao = n_cart;
    int c_ao = 0, c_p = 0;
    for (int s = 0; s < n_cont; ++s) {
        const int nf = assign[s], np = pnum_list[s], at = atom_indices[s];
        if (nf < 0 || np < 0 || c_ao + nf > n_cart || c_p + np > n_prim || at < 0 || at >= n_atoms) {
            delete b;
            return fail(OKB_ERR_ARG, "okb_basis_create: contraction %d inconsistent with array sizes", s);
        }
        // split a contraction with mixed angular momentum into pure-L device shells
        std::vector<int> Ls;
        for (int j = 0; j < nf; ++j) {
            const int *l = lxlylz + 3 * (c_ao + j);
            if (l[0] < 0 || l[1] < 0 || l[2] < 0 || l[0] > 15 || l[1] > 15 || l[2] > 15) {
                delete b;
                return fail(OKB_ERR_ARG, "okb_basis_create: exponent out of range at function %d", c_ao + j);
            }
            const int L = l[0] + l[1] + l[2];
            if (std::find(Ls.begin(), Ls.end(), L) == Ls.end()) Ls.push_back(L);
        }
        for (int L : Ls) {
            DevShell sh;
            sh.L = L;
            for (int a = 0; a < 3; ++a) sh.c[a] = geo_spec[3 * at + a];
            for (int i = 0; i < np; ++i) {
                const double alpha = ao_coeffs[2 * (c_p + i)], coef = ao_coeffs[2 * (c_p + i) + 1];
                sh.alpha.push_back(alpha);
                sh.cn.push_back(is_normalized > 0 ? coef : coef * radial_norm(L, alpha));
            }
            for (int j = 0; j < nf; ++j) {
                const int *l = lxlylz + 3 * (c_ao + j);
                if (l[0] + l[1] + l[2] != L) continue;
                sh.fn_row.push_back(c_ao + j);
                sh.lx.push_back(l[0]); sh.ly.push_back(l[1]); sh.lz.push_back(l[2]);
                double f = is_normalized > 0 ? 1.0 : angular_norm(l[0], l[1], l[2]);
                if (renorm) f *= renorm[c_ao + j];
                sh.f.push_back(f);
            }
            b->shells.push_back(std::move(sh));
        }
        c_ao += nf;
        c_p += np;
    }
    if (c_ao != n_cart) {
        delete b;
        return fail(OKB_ERR_ARG, "okb_basis_create: sum(assign)=%d != n_cart=%d", c_ao, n_cart);
    }
    int rc = basis_upload(b);
    if (rc != OKB_OK) {
        delete b;
        return rc;
    }
    *out = b;
    return OKB_OK;
}

extern "C" int okb_basis_set_cart2sph(okb_basis *b, int n_sph, const int *row_ptr, const int *col,
                                      const double *val) {
    if (!b || !row_ptr || !col || !val || n_sph <= 0) return fail(OKB_ERR_ARG, "okb_basis_set_cart2sph: bad argument");
    const int nnz = row_ptr[n_sph];
    for (int t = 0; t < nnz; ++t)
        if (col[t] < 0 || col[t] >= b->n_cart)
            return fail(OKB_ERR_ARG, "okb_basis_set_cart2sph: column %d out of range", col[t]);
    b->t_ptr.assign(row_ptr, row_ptr + n_sph + 1);
    b->t_col.assign(col, col + nnz);
    b->t_val.assign(val, val + nnz);
    b->n_ao = n_sph;
    b->spherical = true;
    return basis_upload(b);
}

extern "C" int okb_basis_info(okb_basis *b, int *n_cart, int *n_ao, int *n_dev_shells, int *n_chunks) {
    if (!b) return fail(OKB_ERR_ARG, "null basis");
    if (n_cart) *n_cart = b->n_cart;
    if (n_ao) *n_ao = b->n_ao;
    if (n_dev_shells) *n_dev_shells = (int)b->shells.size();
    if (n_chunks) *n_chunks = (int)b->contraction_layout().chunks.size();
    return OKB_OK;
}

extern "C" int okb_basis_destroy(okb_basis *b) {
    if (!b) return OKB_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    layout_free(b->ctx, b->cart);
    layout_free(b->ctx, b->mix);
    pool_put(b->ctx, b->prim5_dev, b->prim5_bytes);
    delete b;
    return OKB_OK;
}

// ---- MO coefficients ------------------------------------------------------------------------------------
extern "C" int okb_mo_create(okb_ctx *ctx, okb_basis *b, int n_mo, const double *coeffs, const double *occ,
                             okb_mo **out) {
    if (!ctx || !b || !out || !coeffs) return fail(OKB_ERR_ARG, "okb_mo_create: null argument");
    *out = nullptr;
    if (n_mo <= 0) return fail(OKB_ERR_ARG, "okb_mo_create: n_mo must be positive");
    okb_mo *m = new okb_mo();
    m->ctx = ctx;
    m->basis = b;
    m->n_mo = n_mo;
    m->occ.assign(n_mo, 0.0);
    if (occ) m->occ.assign(occ, occ + n_mo);
    m->ccart.assign((size_t)n_mo * b->n_cart, 0.0);
    if (!b->spherical) {
        memcpy(m->ccart.data(), coeffs, sizeof(double) * (size_t)n_mo * b->n_cart);
    } else {
        m->csph.assign(coeffs, coeffs + (size_t)n_mo * b->n_ao);
        // C'[i][cart] = sum_j C[i][j] * T[j][cart]; accumulated in (j, term) order like the
        // row axpys of core.py:168-174
        for (int i = 0; i < n_mo; ++i) {
            double *dst = m->ccart.data() + (size_t)i * b->n_cart;
            const double *src = coeffs + (size_t)i * b->n_ao;
            for (int j = 0; j < b->n_ao; ++j)
                for (int t = b->t_ptr[j]; t < b->t_ptr[j + 1]; ++t) dst[b->t_col[t]] += src[j] * b->t_val[t];
        }
    }
    *out = m;
    return OKB_OK;
}

// MC: orbitals of an MO tile contracted by the consumer warps (DMMA blocks), rem: remainder orbitals of the tile contracted
// by the producer warps (okb_ws.cuh); the tile holds MCT = MC + rem orbitals
static int mo_blob(okb_mo *m, int MC, int rem, const Layout &lo, bool is_mix, okb_mo::Blob **out) {
    const int MCT = MC + rem;
    const int key = 2 * (MC * 16 + rem) + (is_mix ? 1 : 0);
    auto it = m->blobs.find(key);
    if (it != m->blobs.end()) {
        *out = &it->second;
        return OKB_OK;
    }
    okb_basis *b = m->basis;
    const int nchunk = (int)lo.chunks.size();
    const int n_mtile = (m->n_mo + MCT - 1) / MCT;
    const int CS = pad_stride(MC);           // row stride = 4 (mod 16) doubles: conflict-free MMA fragment loads
    std::vector<double> blob((size_t)n_mtile * nchunk * KC * CS, 0.0);
    std::vector<double> crem((size_t)n_mtile * nchunk * KC * std::max(rem, 1), 0.0);
    for (int mt = 0; mt < n_mtile; ++mt)
        for (int c = 0; c < nchunk; ++c) {
            double *dst = blob.data() + ((size_t)mt * nchunk + c) * KC * CS;
            double *rdst = crem.data() + ((size_t)mt * nchunk + c) * KC * rem;
            for (int kk = 0; kk < lo.chunks[c].nfn; ++kk) {
                const KRow &kr = lo.krow[lo.chunks[c].k0 + kk];
                for (int r = 0; r < rem; ++r) {          // [row][rem]: one 16-byte load per row for rem = 2
                    const int mo = mt * MCT + MC + r;
                    if (mo >= m->n_mo) continue;
                    rdst[(size_t)kk * rem + r] = kr.is_sph ? m->csph[(size_t)mo * b->n_ao + kr.index]
                                                           : m->ccart[(size_t)mo * b->n_cart + kr.index];
                }
                for (int i = 0; i < MC; ++i) {
                    const int mo = mt * MCT + i;
                    if (mo >= m->n_mo) continue;
                    // MO blocks of 8 are stored in pairs, row by row: (block 2q, row r) and (block 2q+1, row r) are
                    // neighbours, so that a consumer lane fetches its A fragments of two blocks with one 16-byte load
                    const int blk = i >> 3, r = i & 7, pos = (blk >> 1) * 16 + 2 * r + (blk & 1);
                    dst[(size_t)kk * CS + pos] = kr.is_sph ? m->csph[(size_t)mo * b->n_ao + kr.index]
                                                           : m->ccart[(size_t)mo * b->n_cart + kr.index];
                }
            }
        }
    std::vector<double> occ((size_t)n_mtile * MCT, 0.0);
    std::copy(m->occ.begin(), m->occ.end(), occ.begin());
    okb_mo::Blob bl;
    bl.n_mtile = n_mtile;
    CU(cudaSetDevice(m->ctx->device));
    {
        void *raw = nullptr;
        int rcp = pool_get(m->ctx, blob.size() * sizeof(double), &raw, &bl.c_bytes);
        if (rcp != OKB_OK) return rcp;
        bl.c = reinterpret_cast<double *>(raw);
        rcp = pool_get(m->ctx, occ.size() * sizeof(double), &raw, &bl.occ_bytes);
        if (rcp != OKB_OK) return rcp;
        bl.occ = reinterpret_cast<double *>(raw);
        if (rem > 0) {
            rcp = pool_get(m->ctx, crem.size() * sizeof(double), &raw, &bl.crem_bytes);
            if (rcp != OKB_OK) return rcp;
            bl.crem = reinterpret_cast<double *>(raw);
            CU(cudaMemcpy(bl.crem, crem.data(), crem.size() * sizeof(double), cudaMemcpyHostToDevice));
            m->ctx->h2d_bytes += (long long)(crem.size() * sizeof(double));
        }
    }
    CU(cudaMemcpy(bl.c, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(bl.occ, occ.data(), occ.size() * sizeof(double), cudaMemcpyHostToDevice));
    m->ctx->h2d_bytes += (long long)((blob.size() + occ.size()) * sizeof(double));
    m->blobs[key] = bl;
    *out = &m->blobs[key];
    return OKB_OK;
}

extern "C" int okb_mo_destroy(okb_mo *m) {
    if (!m) return OKB_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (auto &kv : m->blobs) {
        pool_put(m->ctx, kv.second.c, kv.second.c_bytes);
        pool_put(m->ctx, kv.second.occ, kv.second.occ_bytes);
        if (kv.second.crem) pool_put(m->ctx, kv.second.crem, kv.second.crem_bytes);
    }
    delete m;
    return OKB_OK;
}

// ---- grids -----------------------------------------------------------------------------------------------
extern "C" int okb_grid_destroy(okb_grid *g);
// destroys a half-built grid on every early return of its constructor (CU() / pool_get failures)
struct GridGuard {
    okb_grid *g;
    ~GridGuard() { if (g) okb_grid_destroy(g); }
    okb_grid *release() { okb_grid *r = g; g = nullptr; return r; }
};
extern "C" int okb_grid_regular(okb_ctx *ctx, const double *x, int nx, const double *y, int ny,
                                const double *z, int nz, okb_grid **out) {
    if (!ctx || !out || !x || !y || !z) return fail(OKB_ERR_ARG, "okb_grid_regular: null argument");
    *out = nullptr;
    if (nx <= 0 || ny <= 0 || nz <= 0) return fail(OKB_ERR_ARG, "okb_grid_regular: empty axis");
    okb_grid *g = new okb_grid();
    GridGuard guard{g};
    g->ctx = ctx;
    g->kind = 0;
    g->nx = nx; g->ny = ny; g->nz = nz;
    g->npts = (long long)nx * ny * nz;
    CU(cudaSetDevice(ctx->device));
    {
        double **dst[3] = {&g->gx, &g->gy, &g->gz};
        const int n[3] = {nx, ny, nz};
        for (int a = 0; a < 3; ++a) {
            void *raw = nullptr;
            int rcp = pool_get(ctx, sizeof(double) * n[a], &raw, &g->gbytes[a]);
            if (rcp != OKB_OK) return rcp;
            *dst[a] = reinterpret_cast<double *>(raw);
        }
    }
    CU(cudaMemcpy(g->gx, x, sizeof(double) * nx, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g->gy, y, sizeof(double) * ny, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(g->gz, z, sizeof(double) * nz, cudaMemcpyHostToDevice));
    ctx->h2d_bytes += (long long)sizeof(double) * (nx + ny + nz);
    *out = guard.release();
    return OKB_OK;
}

extern "C" int okb_grid_vector(okb_ctx *ctx, const double *x, const double *y, const double *z,
                               long long npts, int coords_on_device, okb_grid **out) {
    if (!ctx || !out || !x || !y || !z) return fail(OKB_ERR_ARG, "okb_grid_vector: null argument");
    *out = nullptr;
    if (npts <= 0) return fail(OKB_ERR_ARG, "okb_grid_vector: empty grid");
    okb_grid *g = new okb_grid();
    GridGuard guard{g};
    g->ctx = ctx;
    g->kind = 1;
    g->npts = npts;
    CU(cudaSetDevice(ctx->device));
    if (coords_on_device) {
        g->gx = const_cast<double *>(x);
        g->gy = const_cast<double *>(y);
        g->gz = const_cast<double *>(z);
        g->owns = false;
    } else {
        double **dst[3] = {&g->gx, &g->gy, &g->gz};
        for (int a = 0; a < 3; ++a) {
            void *raw = nullptr;
            int rcp = pool_get(ctx, sizeof(double) * npts, &raw, &g->gbytes[a]);
            if (rcp != OKB_OK) return rcp;
            *dst[a] = reinterpret_cast<double *>(raw);
        }
        CU(cudaMemcpyAsync(g->gx, x, sizeof(double) * npts, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(g->gy, y, sizeof(double) * npts, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(g->gz, z, sizeof(double) * npts, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->h2d_bytes += 3ll * (long long)sizeof(double) * npts;
    }
    *out = guard.release();
    return OKB_OK;
}

// Product grids in non-Cartesian coordinates (cy_grid.sph2cart / cyl2cart, cy_grid.pyx:58-97): the coordinates are
// generated in the kernels from the three axis vectors -- 0 input bytes per point instead of 24.  sin / cos of the
// angular axes are taken here on the host (the same libm the reference's Cython code calls) so that the device
// reproduces the reference's coordinates bit for bit.
extern "C" int okb_grid_product(okb_ctx *ctx, int kind, const double *a0, int n0, const double *a1, int n1,
                                const double *a2, int n2, const double *affine, okb_grid **out) {
    if (!ctx || !out || !a0 || !a1 || !a2) return fail(OKB_ERR_ARG, "okb_grid_product: null argument");
    *out = nullptr;
    if (kind != 2 && kind != 3) return fail(OKB_ERR_ARG, "okb_grid_product: kind must be 2 (spherical) or 3 (cylindrical)");
    if (n0 <= 0 || n1 <= 0 || n2 <= 0) return fail(OKB_ERR_ARG, "okb_grid_product: empty axis");
    okb_grid *g = new okb_grid();
    GridGuard guard{g};
    g->ctx = ctx;
    g->kind = kind;
    g->nx = n0; g->ny = n1; g->nz = n2;
    g->npts = (long long)n0 * n1 * n2;
    if (affine) {
        g->has_aff = true;
        memcpy(g->aff, affine, sizeof(double) * 12);
    }
    std::vector<double> t1((size_t)2 * n1), t2((size_t)2 * n2);
    for (int j = 0; j < n1; ++j) {
        t1[j] = kind == 2 ? sin(a1[j]) : cos(a1[j]);            // spherical: sin(theta), cos(theta)
        t1[n1 + j] = kind == 2 ? cos(a1[j]) : sin(a1[j]);       // cylindrical: cos(phi), sin(phi)
    }
    for (int k = 0; k < n2; ++k) {
        t2[k] = kind == 2 ? cos(a2[k]) : a2[k];                 // spherical: cos(phi), sin(phi); cylindrical: zed
        t2[n2 + k] = kind == 2 ? sin(a2[k]) : 0.0;
    }
    CU(cudaSetDevice(ctx->device));
    double **dst[3] = {&g->gx, &g->gy, &g->gz};
    const double *src[3] = {a0, t1.data(), t2.data()};
    const size_t cnt[3] = {(size_t)n0, (size_t)2 * n1, (size_t)2 * n2};
    for (int a = 0; a < 3; ++a) {
        void *raw = nullptr;
        int rcp = pool_get(ctx, sizeof(double) * cnt[a], &raw, &g->gbytes[a]);
        if (rcp != OKB_OK) return rcp;
        *dst[a] = reinterpret_cast<double *>(raw);
        CU(cudaMemcpy(*dst[a], src[a], sizeof(double) * cnt[a], cudaMemcpyHostToDevice));
        ctx->h2d_bytes += (long long)(sizeof(double) * cnt[a]);
    }
    *out = guard.release();
    return OKB_OK;
}

extern "C" int okb_grid_size(okb_grid *g, long long *npts) {
    if (!g || !npts) return fail(OKB_ERR_ARG, "null pointer");
    *npts = g->npts;
    return OKB_OK;
}

extern "C" int okb_grid_destroy(okb_grid *g) {
    if (!g) return OKB_OK;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    for (auto &kv : g->axis_tabs) pool_put(g->ctx, kv.second.ptr, kv.second.bytes);
    if (g->owns) {
        pool_put(g->ctx, g->gx, g->gbytes[0]);
        pool_put(g->ctx, g->gy, g->gbytes[1]);
        pool_put(g->ctx, g->gz, g->gbytes[2]);
    }
    delete g;
    return OKB_OK;
}

// ---- launch machinery ---------------------------------------------------------------------------------------
// the kernel instantiations live in inst_*.cu (compiled in parallel); okb_variant.h declares their tables
// (SINK_AO: the first matching entry is the default: the warp-specialised "aows/" kernels for the derivative sets -- twice
// the throughput of the tile kernel there -- and the tile kernel for plain values, where both measure the same because the
// AO generators, not the stores, bound calc_ao)
// (ties of the cost model go to the table listed first: the narrow-tile variants with wide point tiles come before the
// older narrow variants of the per-set tables)
static const VariantTable *const g_tables[] = {&okb_variants_aows, &okb_variants_tile, &okb_variants_narrow_a,
                                               &okb_variants_narrow_b, &okb_variants_val, &okb_variants_grad,
                                               &okb_variants_lap, &okb_variants_all, &okb_variants_d2, &okb_variants_d2p,
                                               &okb_variants_rem};

// ao_bulk_ok: the SINK_AO output rows start on 16-byte boundaries (the "aows/" kernels store them with bulk copies)
// meta_stride > 0: skip the MO-tile variants whose shared memory does not fit with chunk tables of that size
// npts / n_sm: points of the request and SMs of the device -- value-set variants with 32-point tiles are for requests of
// at most 96 points per SM only (and are preferred there: a 128-point tile takes ~4x as long as a 32-point one, so up to
// three waves of 32-point tiles beat one partial wave of 128-point tiles)
static const Variant *pick_variant(int set, int sink, int n_mo, bool ao_bulk_ok = false, int meta_stride = 0,
                                   long long npts = -1, int n_sm = 148) {
    const Variant *best = nullptr;
    long long best_cost = 0;
    for (const VariantTable *tab : g_tables)
      for (int iv = 0; iv < tab->n; ++iv) {
        const Variant &v = tab->v[iv];
        if (v.set != set || v.sink != sink) continue;
        if (sink == SINK_AO) {
            static const char *force_ao = getenv("OKB_AO_VARIANT");      // A/B measurements only
            if (force_ao && force_ao[0]) {
                // a forced name that carries a set ("...SET_GRAD...") only applies to requests of that set
                static const char *const set_names[] = {"SET_VAL", "SET_GRAD", "SET_LAP", "SET_ALL", "SET_ONE", "SET_D2", "SET_D2P"};
                const bool applies = !strstr(force_ao, "SET_") || strstr(force_ao, set_names[set]);
                if (applies && !strstr(v.name, force_ao)) continue;
            }
            const bool is_aows = strncmp(v.name, "aows/", 5) == 0;
            if (is_aows && (!ao_bulk_ok || (set == SET_VAL && !(force_ao && force_ao[0])))) continue;
            return &v;
        }
        // OKB_VARIANT=<substring of a variant name> forces a configuration (A/B measurements only)
        static const char *force = getenv("OKB_VARIANT");
        if (force && force[0] && strstr(v.name, force)) return &v;
        if (meta_stride > 0 && v.smem(meta_stride) > 227 * 1024) continue;
        static const char *small_env = getenv("OKB_SMALL_PTS_PER_SM");     // A/B
        static const long long small_per_sm = small_env && small_env[0] ? atoll(small_env) : 96;
        const bool small = npts >= 0 && npts <= small_per_sm * n_sm;
        if (set == SET_VAL && v.P == 32 && !small) continue;
        // Every MO tile regenerates the AO tiles, and the producers run beside the consumers.  Measured per pass over one
        // MO tile on the Config-2 shape (222 AOs; 24- / 48- / 80-wide tiles, profiles/r02_c2_narrow.txt) the time is affine
        // in the tile width MC: rho 3.0 / 3.55 / 4.5 ms, rho + grad 8.7 / 11.3 / 17.4 ms, second laplacian pass
        // 7.8 / 9.55 / 14.4 ms, i.e. ~ MC + K with K = 100 / 35 / 42 orbitals' worth of AO generation per tile.  (The
        // other sets keep the older rule: narrow tiles charged as 48 wide, 24 for SET_ALL.)
        // Prefer the wider tile on ties (fewer AO regenerations).
        // Remainder orbitals (v.rem of the v.MC, contracted by the producer warps) are charged 3 DMMA columns each:
        // 82 MOs go to the 80 + 2 tile (86) instead of the 88-wide one, 83..88 MOs stay on the latter.
        const long long n_tiles = (n_mo + v.MC - 1) / v.MC;
        // (The second laplacian pass, SET_D2P, has 3 sets per row for the consumers but the heaviest generators: there the
        // producers bound the kernel and a remainder orbital costs more than its DMMA columns -- 82 MOs: 88-wide 144 ms,
        // 80 + 2 148 ms, profiles/r02_lap_passes_ab.txt -- so it is charged 5 columns.)
        const int rem_cols = set == SET_D2P ? 5 : 3;
        const int width = v.MC - v.rem + rem_cols * v.rem;
        const int gen = set == SET_VAL ? 100 : set == SET_GRAD ? 35 : set == SET_D2P ? 42 : 0;
        long long cost = n_tiles * (gen > 0 ? width + gen : std::max(width, set == SET_ALL ? 24 : 48)) * 1000 - v.MC;
        if (small && set == SET_VAL) cost += v.P;               // small request: the narrowest point tile of that width
        if (!best || cost < best_cost) {
            best = &v;
            best_cost = cost;
        }
    }
    return best;
}

static int codes_to_set(const int *codes, int n, int *set) {
    int mx = 0;
    for (int i = 0; i < n; ++i) {
        if (codes[i] < 0 || codes[i] > 9) return fail(OKB_ERR_ARG, "derivative code %d not in 0..9", codes[i]);
        mx = std::max(mx, codes[i]);
    }
    *set = mx == 0 ? SET_VAL : mx <= 3 ? SET_GRAD : mx <= 6 ? SET_LAP : SET_ALL;
    return OKB_OK;
}

static int ensure_slabs(okb_ctx *c, size_t bytes) {
    if (c->slab_bytes >= bytes) return OKB_OK;
    for (int i = 0; i < 2; ++i) {
        if (c->slab[i]) CU(cudaFree(c->slab[i]));
        c->slab[i] = nullptr;
    }
    c->slab_bytes = 0;
    for (int i = 0; i < 2; ++i) CU(cudaMalloc(&c->slab[i], bytes));
    c->slab_bytes = bytes;
    return OKB_OK;
}

// Separable-exponential tables of (basis, regular grid): built on first use, cached in the grid handle.
static int ensure_axis_tables(okb_ctx *ctx, okb_basis *b, okb_grid *g, const double **tx, const double **ty,
                              const double **tz) {
    *tx = *ty = *tz = nullptr;
    static const bool off = getenv("OKB_NO_AXIS_TABLES") != nullptr;
    const size_t np = (size_t)b->n_prim_dev;
    const size_t total = np * ((size_t)g->nx + g->ny + g->nz);
    // index arithmetic in the kernels is 32-bit
    if (off || g->kind != 0 || np == 0 || np * (size_t)std::max(g->nx, std::max(g->ny, g->nz)) >= ((size_t)1 << 31))
        return OKB_OK;
    auto it = g->axis_tabs.find(b->serial);
    double *base = nullptr;
    if (it != g->axis_tabs.end()) {
        base = it->second.ptr;
    } else {
        size_t got = 0;
        void *raw = nullptr;
        int rcp = pool_get(ctx, total * sizeof(double), &raw, &got);
        if (rcp != OKB_OK) return rcp;
        base = reinterpret_cast<double *>(raw);
        const double *axes[3] = {g->gx, g->gy, g->gz};
        const int n[3] = {g->nx, g->ny, g->nz};
        double *dst = base;
        for (int a = 0; a < 3; ++a) {
            const long long cnt = (long long)np * n[a];
            okb_axis_table_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, ctx->stream>>>(b->prim5_dev, (int)np, axes[a], n[a],
                                                                                          a, dst);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                pool_put(ctx, base, got);
                return fail(OKB_ERR_CUDA, "axis table kernel: %s", cudaGetErrorString(e));
            }
            ctx->launches++;
            dst += cnt;
        }
        g->axis_tabs[b->serial] = okb_grid::Tab{base, got};
    }
    *tx = base;
    *ty = base + np * g->nx;
    *tz = *ty + np * g->ny;
    return OKB_OK;
}

struct EvalReq {
    int sink;
    okb_basis *basis;
    okb_mo *mo;
    okb_grid *grid;
    long long p0, p1;
    const int *codes;
    int n_codes;
    double *out;        // AO/MO
    double *rho, *delta;
    double *mo_norm;
    unsigned flags;
    long long ld_out = 0;   // row stride (points) of the caller's output arrays; 0 = dense (p1 - p0)
};

static int run_eval(okb_ctx *ctx, const EvalReq &rq) {
    okb_basis *b = rq.basis;
    okb_grid *g = rq.grid;
    if (!ctx || !b || !g) return fail(OKB_ERR_ARG, "eval: null handle");
    if (b->ctx != ctx || g->ctx != ctx) return fail(OKB_ERR_ARG, "eval: handles belong to another context");
    if (rq.p0 < 0 || rq.p1 > g->npts || rq.p1 < rq.p0)
        return fail(OKB_ERR_ARG, "eval: point range [%lld,%lld) outside the grid (%lld points)", rq.p0, rq.p1, g->npts);
    if (rq.n_codes < 0 || rq.n_codes > 64) return fail(OKB_ERR_ARG, "eval: bad number of derivative codes");
    if (rq.n_codes > 0 && !rq.codes) return fail(OKB_ERR_ARG, "eval: null derivative code list");
    const long long ntot = rq.p1 - rq.p0;
    if (rq.ld_out != 0 && rq.ld_out < ntot) return fail(OKB_ERR_ARG, "eval: output row stride %lld < %lld points", rq.ld_out, ntot);
    const long long ldo = rq.ld_out ? rq.ld_out : ntot;      // row stride of the caller's arrays
    if (ntot == 0) return OKB_OK;
    CU(cudaSetDevice(ctx->device));

    // which derivative sets must be evaluated, and where does each requested code go
    int set = SET_VAL;
    int rc = codes_to_set(rq.codes, rq.n_codes, &set);
    if (rc != OKB_OK) return rc;
    const bool dev_out = (rq.flags & OKB_FLAG_OUT_DEVICE) != 0;
    const int n_rows = rq.sink == SINK_AO ? b->n_ao : (rq.sink == SINK_MO ? rq.mo->n_mo : 1);

    // Passes: normally one pass with the smallest covering set.  AO/MO requests whose code list has
    // duplicates or is a single code use one SET_ONE pass per requested slot.
    struct Pass { int set; int one_code; int slot[10]; int epi = 0; };
    std::vector<Pass> passes;
    bool phi_cache = false;                  // two-pass rho + laplacian with the MO values kept in ctx->phi_buf
    bool dup = false;
    {
        int seen[10] = {0};
        for (int i = 0; i < rq.n_codes; ++i) dup |= (seen[rq.codes[i]]++ > 0);
    }
    if (rq.sink == SINK_RHO) {
        if (dup) return fail(OKB_ERR_ARG, "eval_rho: duplicate derivative codes");
        for (int i = 0; i < rq.n_codes; ++i)
            if (rq.codes[i] == 0) return fail(OKB_ERR_ARG, "eval_rho: derivative code 0 is not a derivative");
        Pass ps;
        ps.set = set; ps.one_code = 0;
        for (int k = 0; k < 10; ++k) ps.slot[k] = -1;
        for (int i = 0; i < rq.n_codes; ++i) ps.slot[rq.codes[i]] = i;
        // rho + pure second derivatives only (laplacian=True): two passes of the 4-set kernels instead of one pass
        // of the 7-set kernel -- 8/7 of the flops, but at the gradient kernel's efficiency (the 7-set kernel is
        // starved of registers: 8 consumer + 4 producer warps; measured 438 -> 39x ms on the benchmark).
        //   pass 1  SET_GRAD, epi 1: rho, and sum 2 occ (d_d phi)^2 into the slots of codes 4..6
        //   pass 2  SET_D2,   epi 2: those slots += sum 2 occ phi d_dd phi
        bool pure_second = (set == SET_LAP);
        for (int i = 0; i < rq.n_codes; ++i) pure_second &= (rq.codes[i] >= 4 && rq.codes[i] <= 6);
        static const bool one_pass = getenv("OKB_LAP_ONE_PASS") != nullptr;
        static const bool no_cache = getenv("OKB_LAP_NO_PHI_CACHE") != nullptr;     // A/B measurements only
        if (pure_second && !one_pass && !no_cache) {
            // the first pass leaves the MO values in HBM (8 n_mo bytes per point written once, read once: ~1 % of the
            // pass), so the second pass contracts the three second-derivative sets only: 4 + 3 = 7 sets, the algorithmic
            // count, instead of 4 + 4
            ps.set = SET_GRAD; ps.epi = 3;
            passes.push_back(ps);
            ps.set = SET_D2P; ps.epi = 2;
            passes.push_back(ps);
            phi_cache = true;
        } else if (pure_second && !one_pass) {
            ps.set = SET_GRAD; ps.epi = 1;
            passes.push_back(ps);
            ps.set = SET_D2; ps.epi = 2;
            passes.push_back(ps);
        } else {
            passes.push_back(ps);
        }
    } else if (rq.n_codes == 1 || dup) {
        for (int i = 0; i < rq.n_codes; ++i) {
            Pass ps;
            ps.set = rq.codes[i] == 0 ? SET_VAL : SET_ONE;
            ps.one_code = rq.codes[i];
            for (int k = 0; k < 10; ++k) ps.slot[k] = -1;
            ps.slot[rq.codes[i]] = i;
            passes.push_back(ps);
        }
    } else {
        Pass ps;
        ps.set = set; ps.one_code = 0;
        for (int k = 0; k < 10; ++k) ps.slot[k] = -1;
        for (int i = 0; i < rq.n_codes; ++i) ps.slot[rq.codes[i]] = i;
        passes.push_back(ps);
    }

    // output geometry: rows of `ntot` points; host outputs are produced slab by slab
    const size_t n_out_rows = rq.sink == SINK_RHO ? (size_t)(1 + rq.n_codes) : (size_t)rq.n_codes * n_rows;
    long long slab_pts = ntot;
    if (!dev_out) {
        const size_t budget = (size_t)192 << 20;     // bytes per staging slab
        long long fit = (long long)(budget / (n_out_rows * sizeof(double)));
        fit = std::max<long long>(fit / 1024 * 1024, 1024);
        slab_pts = std::min(ntot, fit);
        rc = ensure_slabs(ctx, (size_t)slab_pts * n_out_rows * sizeof(double));
        if (rc != OKB_OK) return rc;
    }
    slab_pts = std::min<long long>(slab_pts, (long long)1 << 30);

    double *norm_dev = nullptr;
    if (rq.sink == SINK_RHO && rq.mo_norm) {
        if (ctx->norm_cap < (size_t)rq.mo->n_mo) {
            if (ctx->norm_dev) CU(cudaFree(ctx->norm_dev));
            ctx->norm_dev = nullptr;
            CU(cudaMalloc(&ctx->norm_dev, sizeof(double) * rq.mo->n_mo));
            ctx->norm_cap = rq.mo->n_mo;
        }
        norm_dev = ctx->norm_dev;
        CU(cudaMemsetAsync(norm_dev, 0, sizeof(double) * rq.mo->n_mo, ctx->stream));
    }

    const double *tabx = nullptr, *taby = nullptr, *tabz = nullptr;
    // SINK_AO: plain values on a regular grid run the z-run kernel (okb_ao_zrun.cuh), which is built on the tables; the
    // exponential-per-point SINK_AO kernels do not take them (too few warps to hide the table loads: measured 1.6x slower)
    static const bool ao_tables = getenv("OKB_AO_TABLES") != nullptr;      // A/B measurements only
    static const bool no_zrun = getenv("OKB_NO_ZRUN") != nullptr;          // A/B measurements only
    bool zrun_ok = rq.sink == SINK_AO && g->kind == 0 && !no_zrun;
    if (zrun_ok) {
        // values and the derivative codes 1..6 (one launch per code); the mixed second derivatives 7..9 keep the
        // kernels that reproduce the reference's formulas
        bool all_ok = true;
        for (int i = 0; i < rq.n_codes; ++i) all_ok &= (rq.codes[i] >= 0 && rq.codes[i] <= 6);
        for (const DevShell &sh : b->shells) all_ok &= (sh.L <= ZR_MAXL);
        static const bool no_zrun_drv = getenv("OKB_NO_ZRUN_DRV") != nullptr;   // A/B measurements only
        if (no_zrun_drv)
            for (int i = 0; i < rq.n_codes; ++i) all_ok &= (rq.codes[i] == 0);
        zrun_ok = all_ok;
    }
    if (rq.sink != SINK_AO || ao_tables || zrun_ok) {
        rc = ensure_axis_tables(ctx, b, g, &tabx, &taby, &tabz);
        if (rc != OKB_OK) return rc;
    }
    zrun_ok = zrun_ok && tabx != nullptr;

    // Slab boundaries.  Host outputs: the device -> host copy of a slab overlaps the evaluation of the next one, so only the
    // LAST slab's copy is exposed; the last slab is therefore cut short (1/8 of the final stretch, >= 64 K points): a rank of
    // an 8-GPU job that evaluates 1e6 points (32 MB of rho + grad rho) waits for a 4 MB copy instead of the whole 32 MB.
    std::vector<long long> cut;                               // slab start offsets, then ntot
    for (long long s0 = 0; s0 < ntot; s0 += slab_pts) {
        cut.push_back(s0);
        const long long sn = std::min(slab_pts, ntot - s0);
        if (!dev_out && s0 + sn == ntot) {
            const long long tail = std::max<long long>(sn / 8 / 1024 * 1024, 65536);
            if (sn >= 4 * tail) cut.push_back(s0 + sn - tail);
        }
    }
    cut.push_back(ntot);
    for (int slab_idx = 0; slab_idx + 1 < (int)cut.size(); ++slab_idx) {
        const long long s0 = cut[slab_idx], sn = cut[slab_idx + 1] - s0;
        const int buf = slab_idx & 1;
        double *dbase;          // device output base for this slab
        long long ld;
        if (dev_out) {
            ld = ldo;
            dbase = nullptr;
        } else {
            ld = sn;
            dbase = reinterpret_cast<double *>(ctx->slab[buf]);
            if (slab_idx >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[buf], 0));
        }
        // Sub-ranges of the slab: only the two-pass laplacian with cached MO values cuts it further, so that the scratch
        // of MO values (8 n_mo bytes per point) stays bounded; every other request runs the slab in one piece.
        long long sub_pts = sn, ldp = 0;
        if (phi_cache) {
            const char *force_pts = getenv("OKB_PHI_CACHE_PTS");                     // tests: force several sub-ranges
            const long long budget_pts = force_pts ? atoll(force_pts) : (long long)(((size_t)8 << 30) / (8 * (size_t)rq.mo->n_mo));
            sub_pts = std::min(sn, std::max<long long>(budget_pts / 1024 * 1024, 1024));
            ldp = (sub_pts + 1) & ~1LL;
            const size_t need = (size_t)rq.mo->n_mo * (size_t)ldp * sizeof(double);
            if (ctx->phi_bytes < need) {
                CU(cudaStreamSynchronize(ctx->stream));
                if (ctx->phi_buf) CU(cudaFree(ctx->phi_buf));
                ctx->phi_buf = nullptr;
                ctx->phi_bytes = 0;
                CU(cudaMalloc(&ctx->phi_buf, need));
                ctx->phi_bytes = need;
            }
        }
        for (long long u0 = 0; u0 < sn; u0 += sub_pts) {
        const long long un = std::min(sub_pts, sn - u0);
        for (const Pass &ps : passes) {
            bool ao_bulk_ok = false;
            if (rq.sink == SINK_AO) {
                const double *obase = dev_out ? rq.out + s0 : dbase;
                ao_bulk_ok = (reinterpret_cast<uintptr_t>(obase) % 16 == 0) && (ld % 2 == 0);
            }
            // spherical-row shells exist only in the straight-line generators (VAL/GRAD/LAP/D2/D2P and the single codes
            // 1..6 of ONE); the generic sets work on the all-Cartesian layout
            if (zrun_ok) {
                const Layout &lo = b->cart;
                KParams p{};
                p.grid_kind = 0;
                p.gx = g->gx; p.gy = g->gy; p.gz = g->gz;
                p.nx = g->nx; p.ny = g->ny; p.nz = g->nz;
                p.tabx = tabx; p.taby = taby; p.tabz = tabz;
                p.p0 = rq.p0 + s0 + u0;
                p.npts = (int)un;
                p.meta = lo.meta_dev;
                p.lay = lo.lay;
                p.nchunk = (int)lo.chunks.size();
                p.ld = ld;
                p.slot_stride = (long long)n_rows * ld;
                for (int k = 0; k < 10; ++k) p.slot[k] = ps.slot[k];
                p.out = (dev_out ? rq.out + s0 : dbase) + u0;
                for (int code = 0; code <= 6; ++code) {      // every code this pass writes: its own launch
                    const bool in_pass = ps.set == SET_ONE ? code == ps.one_code : ps.slot[code] >= 0;
                    if (!in_pass || ps.slot[code] < 0) continue;
                    p.one_code = code;
                    cudaError_t e = okb_launch_ao_zrun(p, ctx->sm_count, ctx->stream);
                    if (e != cudaSuccess)
                        return fail(OKB_ERR_CUDA, "launch of %s failed: %s", okb_ao_zrun_code_name(code), cudaGetErrorString(e));
                    ctx->launches++;
                    ctx->last_kernel = okb_ao_zrun_code_name(code);
                }
                continue;
            }
            const bool use_mix = !b->mix_is_cart &&
                                 (ps.set == SET_VAL || ps.set == SET_GRAD || ps.set == SET_LAP || ps.set == SET_D2 ||
                                  ps.set == SET_D2P || (ps.set == SET_ONE && ps.one_code >= 1 && ps.one_code <= 6));
            const Layout &lo = use_mix ? b->mix : b->cart;
            const Variant *v = pick_variant(ps.set, rq.sink, rq.sink == SINK_AO ? 1 : rq.mo->n_mo, ao_bulk_ok,
                                            rq.sink == SINK_AO ? 0 : lo.lay.stride, rq.p1 - rq.p0, ctx->sm_count);
            // (a basis with very large chunk tables may not leave the stage ring of an "aows/" kernel enough shared memory)
            if (v && rq.sink == SINK_AO && ao_bulk_ok && v->smem(lo.lay.stride) > 227 * 1024)
                v = pick_variant(ps.set, rq.sink, 1, false);
            if (!v) return fail(OKB_ERR_UNSUPPORTED, "no kernel variant for set %d sink %d", ps.set, rq.sink);
            KParams p{};
            p.grid_kind = g->kind;
            p.gx = g->gx; p.gy = g->gy; p.gz = g->gz;
            p.nx = g->nx; p.ny = g->ny; p.nz = g->nz;
            p.has_aff = g->has_aff ? 1 : 0;
            for (int q = 0; q < 12; ++q) p.aff[q] = g->aff[q];
            p.tabx = tabx; p.taby = taby; p.tabz = tabz;
            p.p0 = rq.p0 + s0 + u0;
            p.npts = (int)un;
            p.ntiles = (int)((un + v->P - 1) / v->P);
            p.meta = lo.meta_dev;
            p.lay = lo.lay;
            p.nchunk = (int)lo.chunks.size();
            p.n_mtile = 1;
            p.n_mo = 0;
            if (rq.sink != SINK_AO) {
                okb_mo::Blob *bl = nullptr;
                rc = mo_blob(rq.mo, v->MC - v->rem, v->rem, lo, use_mix, &bl);
                if (rc != OKB_OK) return rc;
                p.cblob = bl->c;
                p.crem = bl->crem;
                p.occ = bl->occ;
                p.n_mtile = bl->n_mtile;
                p.n_mo = rq.mo->n_mo;
            }
            p.ld = ld;
            p.slot_stride = (long long)n_rows * ld;
            for (int k = 0; k < 10; ++k) p.slot[k] = ps.slot[k];
            p.one_code = ps.one_code;
            p.exact_mixed = (rq.flags & OKB_FLAG_EXACT_MIXED) ? 1 : 0;
            p.epi = ps.epi;
            if (rq.sink == SINK_RHO) {
                p.rho = dev_out ? (rq.rho ? rq.rho + s0 + u0 : nullptr) : dbase + u0;
                p.delta = dev_out ? (rq.delta ? rq.delta + s0 + u0 : nullptr) : dbase + ld + u0;
                p.mo_norm = ps.epi == 2 ? nullptr : norm_dev;    // the second laplacian pass must not add the norms again
                p.phi = ctx->phi_buf;
                p.ldp = ldp;
            } else {
                p.out = (dev_out ? rq.out + s0 : dbase) + u0;
            }
            const size_t smem = v->smem(lo.lay.stride);
            if (smem > 227 * 1024) return fail(OKB_ERR_UNSUPPORTED, "variant %s needs %zu bytes of shared memory", v->name, smem);
            // SINK_AO CTAs are small: several per SM overlap generation and stores (the launcher asks the occupancy
            // calculator how many)
            const int grid = rq.sink == SINK_AO ? -ctx->sm_count : std::min(p.ntiles, ctx->sm_count);
            cudaError_t e = v->launch(p, grid, smem, ctx->stream);
            if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "launch of %s failed: %s", v->name, cudaGetErrorString(e));
            ctx->launches++;
            ctx->last_kernel = v->name;
        }
        }   // sub-ranges
        if (!dev_out) {
            CU(cudaEventRecord(ctx->ev_compute[buf], ctx->stream));
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_compute[buf], 0));
            const size_t wbytes = (size_t)sn * sizeof(double);
            // Few rows: one copy per row, so that a caller may page-lock just the row segments it receives (a pitched
            // copy is rejected when its bounding range mixes page-locked and pageable memory: dist.shared_host_array
            // locks only the rank's own point range of every row).  Many rows: one pitched copy.
            double *hrow0 = rq.sink == SINK_RHO ? rq.delta : rq.out;
            const double *drow0 = rq.sink == SINK_RHO ? dbase + ld : dbase;
            const size_t nrow2d = rq.sink == SINK_RHO ? (size_t)rq.n_codes : n_out_rows;
            if (rq.sink == SINK_RHO && rq.rho)
                CU(cudaMemcpyAsync(rq.rho + s0, dbase, wbytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
            if (nrow2d > 0 && nrow2d <= 64) {
                for (size_t r = 0; r < nrow2d; ++r)
                    CU(cudaMemcpyAsync(hrow0 + r * (size_t)ldo + s0, drow0 + r * (size_t)ld, wbytes, cudaMemcpyDeviceToHost,
                                       ctx->copy_stream));
            } else if (nrow2d > 0) {
                CU(cudaMemcpy2DAsync(hrow0 + s0, (size_t)ldo * sizeof(double), drow0, (size_t)ld * sizeof(double), wbytes,
                                     nrow2d, cudaMemcpyDeviceToHost, ctx->copy_stream));
            }
            CU(cudaEventRecord(ctx->ev_copy[buf], ctx->copy_stream));
            ctx->d2h_bytes += (long long)(wbytes * n_out_rows);
        }
    }
    if (norm_dev) {
        CU(cudaMemcpyAsync(rq.mo_norm, norm_dev, sizeof(double) * rq.mo->n_mo, cudaMemcpyDeviceToHost, ctx->stream));
        ctx->d2h_bytes += (long long)sizeof(double) * rq.mo->n_mo;
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (!dev_out) {
        CU(cudaStreamSynchronize(ctx->stream));
        CU(cudaStreamSynchronize(ctx->copy_stream));
    }
    return OKB_OK;
}

extern "C" int okb_eval_ao_ld(okb_ctx *ctx, okb_basis *basis, okb_grid *grid, long long p0, long long p1,
                              const int *drv_codes, int n_drv, double *out, long long ld_out, unsigned flags) {
    if (!out) return fail(OKB_ERR_ARG, "okb_eval_ao: null output");
    if (n_drv <= 0) return fail(OKB_ERR_ARG, "okb_eval_ao: need at least one derivative code");
    EvalReq rq{SINK_AO, basis, nullptr, grid, p0, p1, drv_codes, n_drv, out, nullptr, nullptr, nullptr, flags, ld_out};
    return run_eval(ctx, rq);
}
extern "C" int okb_eval_ao(okb_ctx *ctx, okb_basis *basis, okb_grid *grid, long long p0, long long p1,
                           const int *drv_codes, int n_drv, double *out, unsigned flags) {
    return okb_eval_ao_ld(ctx, basis, grid, p0, p1, drv_codes, n_drv, out, 0, flags);
}

extern "C" int okb_eval_mo_ld(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                              const int *drv_codes, int n_drv, double *out, long long ld_out, unsigned flags) {
    if (!mo) return fail(OKB_ERR_ARG, "okb_eval_mo: null MO handle");
    if (!out) return fail(OKB_ERR_ARG, "okb_eval_mo: null output");
    if (n_drv <= 0) return fail(OKB_ERR_ARG, "okb_eval_mo: need at least one derivative code");
    EvalReq rq{SINK_MO, mo->basis, mo, grid, p0, p1, drv_codes, n_drv, out, nullptr, nullptr, nullptr, flags, ld_out};
    return run_eval(ctx, rq);
}
extern "C" int okb_eval_mo(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                           const int *drv_codes, int n_drv, double *out, unsigned flags) {
    return okb_eval_mo_ld(ctx, mo, grid, p0, p1, drv_codes, n_drv, out, 0, flags);
}

extern "C" int okb_eval_rho_ld(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                               const int *drv_codes, int n_drv, double *rho, double *delta_rho, long long ld_out,
                               double *mo_norm, unsigned flags) {
    if (!mo) return fail(OKB_ERR_ARG, "okb_eval_rho: null MO handle");
    if (!rho) return fail(OKB_ERR_ARG, "okb_eval_rho: null rho output");
    if (n_drv > 0 && !delta_rho) return fail(OKB_ERR_ARG, "okb_eval_rho: null delta_rho output");
    EvalReq rq{SINK_RHO, mo->basis, mo, grid, p0, p1, drv_codes, n_drv, nullptr, rho, delta_rho, mo_norm, flags, ld_out};
    return run_eval(ctx, rq);
}
extern "C" int okb_eval_rho(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1,
                            const int *drv_codes, int n_drv, double *rho, double *delta_rho, double *mo_norm,
                            unsigned flags) {
    return okb_eval_rho_ld(ctx, mo, grid, p0, p1, drv_codes, n_drv, rho, delta_rho, 0, mo_norm, flags);
}

// ---- detCI grid contractions ------------------------------------------------------------------------------------
static int ci_ncomp(int mode, int n_terms, int nd) {
    return mode == CI_RHO ? 1 : mode == CI_PAIRS ? n_terms : mode == CI_JPAIRS ? nd * n_terms : mode == CI_JABF ? nd : 3;
}

static int ci_reserve(okb_ctx *ctx, size_t bytes) {
    if (ctx->ci_bytes >= bytes) return OKB_OK;
    if (ctx->ci_buf) CU(cudaFree(ctx->ci_buf));
    ctx->ci_buf = nullptr;
    ctx->ci_bytes = 0;
    CU(cudaMalloc(&ctx->ci_buf, bytes));
    ctx->ci_bytes = bytes;
    return OKB_OK;
}

static int ci_check(okb_ctx *ctx, int mode, int n_mo, int n_terms, const double *coef, const int *ia, const int *ib,
                    const double *out) {
    if (!ctx) return fail(OKB_ERR_ARG, "ci: null context");
    if (mode < CI_RHO || mode > CI_JABF) return fail(OKB_ERR_ARG, "ci: unknown mode %d", mode);
    if (n_mo <= 0) return fail(OKB_ERR_ARG, "ci: n_mo must be positive");
    if (n_terms < 0) return fail(OKB_ERR_ARG, "ci: negative term count");
    if (n_terms > 0 && (!ia || !ib || (mode != CI_PAIRS && mode != CI_JPAIRS && !coef)))
        return fail(OKB_ERR_ARG, "ci: null term array");
    if (!out) return fail(OKB_ERR_ARG, "ci: null output");
    for (int t = 0; t < n_terms; ++t)
        if (ia[t] < 0 || ia[t] >= n_mo || ib[t] < 0 || ib[t] >= n_mo)
            return fail(OKB_ERR_ARG, "ci: term %d refers to orbitals (%d,%d) outside 0..%d", t, ia[t], ib[t], n_mo - 1);
    return OKB_OK;
}

// term records at the head of ctx->ci_buf: [n_terms] x {coefficient, (a, b)} of 16 bytes, padded to 256 bytes
static size_t ci_terms_bytes(int n_terms) { return (((size_t)n_terms * 16 + 255) / 256 + 1) * 256; }
static int ci_upload_terms(okb_ctx *ctx, int n_terms, const double *coef, const int *ia, const int *ib, CiParams *p) {
    double2 *tp = reinterpret_cast<double2 *>(ctx->ci_buf);
    if (n_terms > 0) {
        std::vector<double2> rec(n_terms);
        for (int t = 0; t < n_terms; ++t) {
            const long long bits = (long long)(unsigned)ia[t] | ((long long)(unsigned)ib[t] << 32);
            double y;
            memcpy(&y, &bits, 8);
            rec[t] = make_double2(coef ? coef[t] : 0.0, y);
        }
        // pageable source: the copy returns after the vector was read
        CU(cudaMemcpyAsync(tp, rec.data(), (size_t)n_terms * 16, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->h2d_bytes += (long long)n_terms * 16;
    }
    p->n_terms = n_terms;
    p->tp = tp;
    return OKB_OK;
}

static int ci_launch(okb_ctx *ctx, int mode, const CiParams &p);

// OKB_FLAG_CI_FAST, many terms per orbital pair (the usual shape of a CI expansion: 1e4 .. 1e6 determinant pairs over a few
// dozen active orbitals): the terms are summed on the host into the dense matrix of the orbital pairs,
//   RHO   M = D             out    = sum_a phi_a (M phi)_a          D[a][b] = sum of the c_t with (a_t, b_t) = (a, b)
//   ANB   M = D             out[d] = sum_a phi_a (M d_d phi)_a
//   JAB   M = -(D - D^T)/2, JABF: M = D - D^T  (same form as ANB)
// and the grid work becomes n_act^2 fused multiply-adds per point on the FP64 tensor path (okb_td_kernel<MTB, true>)
// instead of n_terms gathers.  Taken when r n_terms >= n_act^2, r = the measured cost of a term of the split kernel in
// matrix entries of the DMMA kernel (rho 4.5, jab 3.8, a_nabla_b 2.5: profiles/r02_ci_fast.txt; OKB_CI_DENSE=0|1 forces
// either).
struct CiDensePlan {
    bool on = false;
    int n_act = 0, kp = 0;
    const double *d_w = nullptr;
    const int *d_rows = nullptr;
};

static int ci_dense_plan(okb_ctx *ctx, int mode, int n_mo, int n_terms, const double *coef, const int *ia, const int *ib,
                         CiDensePlan *pl) {
    pl->on = false;
    if (!(mode == CI_RHO || mode == CI_JAB || mode == CI_ANB || mode == CI_JABF) || n_terms == 0) return OKB_OK;
    std::vector<int> idx(n_mo, -1), rows;
    for (int t = 0; t < n_terms; ++t) {
        if (idx[ia[t]] < 0) idx[ia[t]] = 1;
        if (idx[ib[t]] < 0) idx[ib[t]] = 1;
    }
    for (int m = 0; m < n_mo; ++m)
        if (idx[m] > 0) {
            idx[m] = (int)rows.size();
            rows.push_back(m);
        }
    const int n_act = (int)rows.size();
    static const char *force = getenv("OKB_CI_DENSE");
    const double r = mode == CI_RHO ? 4.5 : mode == CI_ANB ? 2.5 : 3.8;
    bool dense = (double)n_terms * r >= (double)n_act * n_act;
    if (force && force[0]) dense = atoi(force) != 0;
    if (!dense || n_act > 4096) return OKB_OK;
    const int kp = std::max(4, (n_act + 3) / 4 * 4), ntp = (n_act + 63) / 64 * 64;
    std::vector<double> d((size_t)n_act * n_act, 0.0), w((size_t)ntp * kp, 0.0);
    for (int t = 0; t < n_terms; ++t) d[(size_t)idx[ia[t]] * n_act + idx[ib[t]]] += coef[t];
    for (int a = 0; a < n_act; ++a)
        for (int b = 0; b < n_act; ++b) {
            const double dab = d[(size_t)a * n_act + b], dba = d[(size_t)b * n_act + a];
            w[(size_t)a * kp + b] = mode == CI_JAB ? -0.5 * (dab - dba) : mode == CI_JABF ? dab - dba : dab;
        }
    const size_t wbytes = (w.size() * 8 + 255) / 256 * 256, need = wbytes + (size_t)n_act * 4;
    CU(cudaSetDevice(ctx->device));
    if (ctx->rdm_bytes < need) {
        CU(cudaStreamSynchronize(ctx->stream));
        if (ctx->rdm_buf) CU(cudaFree(ctx->rdm_buf));
        ctx->rdm_buf = nullptr;
        ctx->rdm_bytes = 0;
        CU(cudaMalloc(&ctx->rdm_buf, need));
        ctx->rdm_bytes = need;
    }
    unsigned char *base = reinterpret_cast<unsigned char *>(ctx->rdm_buf);
    // pageable sources: the copies return after the host vectors were read
    CU(cudaMemcpyAsync(base, w.data(), w.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(base + wbytes, rows.data(), (size_t)n_act * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->h2d_bytes += (long long)(w.size() * 8 + (size_t)n_act * 4);
    pl->on = true;
    pl->n_act = n_act;
    pl->kp = kp;
    pl->d_w = reinterpret_cast<const double *>(base);
    pl->d_rows = reinterpret_cast<const int *>(base + wbytes);
    return OKB_OK;
}

static int ci_launch_dense(okb_ctx *ctx, int mode, const CiParams &p, const CiDensePlan &pl) {
    TdParams t{};
    t.w = pl.d_w;
    t.rows = pl.d_rows;
    t.phi = p.mo;
    t.in = mode == CI_RHO ? p.mo : p.dmo;
    t.dstride_in = p.dstride;
    t.ncomp = mode == CI_RHO ? 1 : mode == CI_JABF ? p.ncomp : 3;
    t.out = p.out;
    t.ldi = p.ld; t.ldo = p.ldo; t.n = p.npts;
    t.nt = t.nk = pl.n_act;
    t.kp = pl.kp;
    t.vec_ok = (reinterpret_cast<uintptr_t>(p.mo) % 16 == 0 && p.ld % 2 == 0) ? 1 : 0;
    const int mtb = (pl.kp > TD_KC && pl.n_act > 32) ? 8 : 4;
    const long long tiles = (p.npts + TD_P - 1) / TD_P;
    if (tiles * t.ncomp > 0x7fffffffLL) return fail(OKB_ERR_ARG, "ci: too many points for one launch");
    const unsigned grid = (unsigned)(tiles * t.ncomp);
    static const char *no_pipe = getenv("OKB_TD_NOPIPE");       // A/B
    const bool in_ok = reinterpret_cast<uintptr_t>(t.in) % 16 == 0 && t.ldi % 2 == 0 && t.dstride_in % 2 == 0;
    if (pl.kp > TD_KC && in_ok && !(no_pipe && no_pipe[0] == '1')) {      // several k chunks: pipelined variant
        CU(cudaFuncSetAttribute(okb_td2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TD2_SMEM));
        okb_td2_kernel<true><<<grid, TD2_NT, TD2_SMEM, ctx->stream>>>(t);
        cudaError_t e2 = cudaGetLastError();
        if (e2 != cudaSuccess) return fail(OKB_ERR_CUDA, "dense ci kernel launch failed: %s", cudaGetErrorString(e2));
        ctx->launches++;
        ctx->last_kernel = mode == CI_RHO ? "ci-dense/rho" : mode == CI_JAB ? "ci-dense/jab" : mode == CI_ANB ? "ci-dense/a_nabla_b"
                                                                                                         : "ci-dense/jab_full";
        return OKB_OK;
    }
    CU(cudaFuncSetAttribute(okb_td_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)td_smem(TD_KC, 4)));
    CU(cudaFuncSetAttribute(okb_td_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)td_smem(TD_KC, 8)));
    if (mtb == 8) okb_td_kernel<8, true><<<grid, TD_NT, td_smem(pl.kp, 8), ctx->stream>>>(t);
    else okb_td_kernel<4, true><<<grid, TD_NT, td_smem(pl.kp, 4), ctx->stream>>>(t);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "dense ci kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    ctx->last_kernel = mode == CI_RHO ? "ci-dense/rho" : mode == CI_JAB ? "ci-dense/jab" : mode == CI_ANB ? "ci-dense/a_nabla_b"
                                                                                                     : "ci-dense/jab_full";
    return OKB_OK;
}

// OKB_FLAG_CI_FAST: the term list split over the warps of a CTA, every MO row staged once (okb_ci_fast_kernel).  Needs
// 16-byte aligned row segments and at least two points per tile in shared memory; everything else (and the ragged rest of
// the points) goes to the bit-identical gather kernel.
static int ci_launch_fast(okb_ctx *ctx, int mode, const CiParams &p, int n_mo, int nsets, const CiDensePlan &pl) {
    if (pl.on) return ci_launch_dense(ctx, mode, p, pl);
    const bool summed = mode == CI_RHO || mode == CI_JAB || mode == CI_ANB || mode == CI_JABF;
    // tile width: 32 points when all row segments fit twice per SM (two resident CTAs: one stages while the other sums),
    // else 16.  Narrower tiles lose: row segments under 128 bytes waste DRAM pages (500 MOs x 4 sets, 96^3 points: jab
    // 11.6 ms at 4 points, 7.5 ms at 8 points in one CTA per SM, 9.5 ms gather kernel; a_nabla_b 11.2 / 6.4 / 5.0 ms --
    // profiles/r02_ci_fast.txt), so wider row sets stay with the gather kernel.  A/B: OKB_CIF_PW = log2(points).
    const size_t row_bytes = (size_t)n_mo * nsets * 8;
    static const char *force_pw = getenv("OKB_CIF_PW");
    int lpw = CIF_FIXED + (row_bytes << 5) <= (size_t)112 * 1024 ? 5 : 4;
    const bool forced = force_pw && force_pw[0];
    if (forced) lpw = std::max(1, std::min(5, atoi(force_pw)));
    const int pw = 1 << lpw;
    const size_t smem = CIF_FIXED + row_bytes * pw;
    const bool aligned = reinterpret_cast<uintptr_t>(p.mo) % 16 == 0 && p.ld % 2 == 0 &&
                         (nsets == 1 || (reinterpret_cast<uintptr_t>(p.dmo) % 16 == 0 && p.dstride % 2 == 0));
    if (!summed || smem > (size_t)(forced ? 224 : 112) * 1024 || !aligned || p.npts < pw || p.n_terms == 0)
        return ci_launch(ctx, mode, p);
    CiFastParams q{};
    q.p = p;
    q.n_mo = n_mo; q.nsets = nsets; q.pw = pw; q.lpw = lpw;
    q.ntiles = p.npts / pw;
    const int per_sm = smem <= (size_t)112 * 1024 ? 2 : 1;
    const unsigned grid = (unsigned)std::min<long long>(q.ntiles, (long long)ctx->sm_count * per_sm);
    cudaError_t e = cudaSuccess;
#define OKB_CIF_LAUNCH(M)                                                                                            \
    e = cudaFuncSetAttribute(okb_ci_fast_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
    if (e == cudaSuccess) okb_ci_fast_kernel<M><<<grid, CIF_NW * 32, smem, ctx->stream>>>(q)
    switch (mode) {
        case CI_RHO: OKB_CIF_LAUNCH(CI_RHO); break;
        case CI_JAB: OKB_CIF_LAUNCH(CI_JAB); break;
        case CI_ANB: OKB_CIF_LAUNCH(CI_ANB); break;
        default: OKB_CIF_LAUNCH(CI_JABF); break;
    }
#undef OKB_CIF_LAUNCH
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "fast ci kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    const long long done = q.ntiles * pw;
    if (done < p.npts) {                                      // ragged rest: fewer than pw points
        CiParams r = p;
        r.mo = p.mo + done;
        r.dmo = p.dmo ? p.dmo + done : nullptr;
        r.out = p.out + done;
        r.npts = p.npts - done;
        const int rc = ci_launch(ctx, mode, r);
        if (rc != OKB_OK) return rc;
    }
    ctx->last_kernel = mode == CI_RHO ? "ci-fast/rho" : mode == CI_JAB ? "ci-fast/jab" : mode == CI_ANB ? "ci-fast/a_nabla_b"
                                                                                                    : "ci-fast/jab_full";
    return OKB_OK;
}

// Bit-identical sums from shared memory (okb_ci_seq_kernel): many terms over rows that fit.  Returns false when the
// request is not of that shape (the caller then takes the gather kernel).  OKB_CI_SEQ=0|1 forces either (A/B).
static bool ci_seq_wanted(const CiParams &p, int mode, int n_mo, int nsets, int *nwp) {
    const bool summed = mode == CI_RHO || mode == CI_JAB || mode == CI_ANB || mode == CI_JABF;
    if (!summed || p.n_terms == 0) return false;
    static const char *force = getenv("OKB_CI_SEQ");
    if (force && force[0] && atoi(force) == 0) return false;
    if (!(force && force[0]) && nsets == 1) return false;     // rho: the gather kernel's rows stay in L1
    const size_t row_bytes = (size_t)n_mo * nsets * 8;
    const bool aligned = reinterpret_cast<uintptr_t>(p.mo) % 16 == 0 && p.ld % 2 == 0 &&
                         (nsets == 1 || (reinterpret_cast<uintptr_t>(p.dmo) % 16 == 0 && p.dstride % 2 == 0));
    if (!aligned || row_bytes * 32 > (size_t)72 * 1024 || p.npts < 32) return false;
    // staging costs ~ the rows, the sums ~ the terms: wins from 6 terms per staged row on (the smallest ratio measured,
    // 2x there: profiles/r02_ci_seq.txt); below 2 the gather kernel is kept
    if (!(force && force[0]) && (long long)p.n_terms < 2LL * n_mo * nsets) return false;
    int w = 4;                                                // 32-point groups per tile: three CTAs per SM if possible
    while (w > 1 && row_bytes * 32 * w > (size_t)72 * 1024) w >>= 1;
    *nwp = w;
    return true;
}

static int ci_launch(okb_ctx *ctx, int mode, const CiParams &p);
static int ci_launch_seq(okb_ctx *ctx, int mode, const CiParams &p, int n_mo, int nsets, int nwp) {
    CiSeqParams q{};
    q.p = p;
    q.n_mo = n_mo; q.nsets = nsets; q.nwp = nwp;
    q.ncw = mode == CI_RHO ? 1 : mode == CI_JABF ? p.ncomp : 3;
    const int tp = 32 * nwp;
    q.ntiles = p.npts / tp;
    const size_t smem = (size_t)n_mo * nsets * 8 * tp;
    const int threads = 32 * nwp * q.ncw;
    cudaError_t e = cudaSuccess;
    int per_sm = 1;
#define OKB_CIS_LAUNCH(M)                                                                                            \
    e = cudaFuncSetAttribute(okb_ci_seq_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, okb_ci_seq_kernel<M>, threads, smem); \
    if (e == cudaSuccess)                                                                                            \
        okb_ci_seq_kernel<M><<<(unsigned)std::min<long long>(q.ntiles, (long long)ctx->sm_count * std::max(per_sm, 1)), \
                               threads, smem, ctx->stream>>>(q)
    switch (mode) {
        case CI_RHO: OKB_CIS_LAUNCH(CI_RHO); break;
        case CI_JAB: OKB_CIS_LAUNCH(CI_JAB); break;
        case CI_ANB: OKB_CIS_LAUNCH(CI_ANB); break;
        default: OKB_CIS_LAUNCH(CI_JABF); break;
    }
#undef OKB_CIS_LAUNCH
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "sequential ci kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    const long long done = q.ntiles * tp;
    if (done < p.npts) {                                      // ragged rest: fewer than one tile of points
        CiParams r = p;
        r.mo = p.mo + done;
        r.dmo = p.dmo ? p.dmo + done : nullptr;
        r.out = p.out + done;
        r.npts = p.npts - done;
        const int rc = ci_launch(ctx, mode, r);
        if (rc != OKB_OK) return rc;
    }
    ctx->last_kernel = mode == CI_RHO ? "ci-seq/rho" : mode == CI_JAB ? "ci-seq/jab" : mode == CI_ANB ? "ci-seq/a_nabla_b"
                                                                                                  : "ci-seq/jab_full";
    return OKB_OK;
}

static int ci_launch(okb_ctx *ctx, int mode, const CiParams &p) {
    if (p.npts <= 0) return OKB_OK;
    const unsigned grid = (unsigned)((p.npts + CI_NT - 1) / CI_NT);
    // A/B measurements only: OKB_CI_SMEM = bytes of (unused) dynamic shared memory per CTA, i.e. a cap on the resident CTAs
    // per SM (does the working set of fewer CTAs stay in L2?  profiles/r02_ci_occupancy.txt)
    static const char *occ_env = getenv("OKB_CI_SMEM");
    const size_t dsm = occ_env ? (size_t)atol(occ_env) : 0;
#define OKB_CI_LAUNCH(M)                                                                                                  \
    do {                                                                                                                  \
        if (dsm) cudaFuncSetAttribute(okb_ci_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm);           \
        okb_ci_kernel<M><<<grid, CI_NT, dsm, ctx->stream>>>(p);                                                           \
    } while (0)
    switch (mode) {
        case CI_RHO: OKB_CI_LAUNCH(CI_RHO); break;
        case CI_JAB: OKB_CI_LAUNCH(CI_JAB); break;
        case CI_ANB: OKB_CI_LAUNCH(CI_ANB); break;
        case CI_JPAIRS: OKB_CI_LAUNCH(CI_JPAIRS); break;
        case CI_JABF: OKB_CI_LAUNCH(CI_JABF); break;
        default: OKB_CI_LAUNCH(CI_PAIRS); break;
    }
#undef OKB_CI_LAUNCH
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "ci kernel launch failed: %s", cudaGetErrorString(e));
    ctx->launches++;
    ctx->last_kernel = mode == CI_RHO ? "ci/rho" : mode == CI_JAB ? "ci/jab" : mode == CI_ANB ? "ci/a_nabla_b"
                       : mode == CI_JPAIRS ? "ci/jpairs" : mode == CI_JABF ? "ci/jab_full" : "ci/pairs";
    return OKB_OK;
}

// reference summation order: shared-memory kernel when the request has its shape, else the gather kernel
static int ci_launch_exact(okb_ctx *ctx, int mode, const CiParams &p, int n_mo, int nsets) {
    int nwp = 1;
    if (ci_seq_wanted(p, mode, n_mo, nsets, &nwp)) return ci_launch_seq(ctx, mode, p, n_mo, nsets, nwp);
    return ci_launch(ctx, mode, p);
}

// ncomp_in: JABF only -- the number of derivative components in molistdrv (1..3); the other modes fix it
static int ci_contract_impl(okb_ctx *ctx, int mode, int n_mo, long long npts, long long ld_in,
                            const double *molist, const double *molistdrv, int ncomp_in, int n_terms, const double *coef,
                            const int *ia, const int *ib, double *out, long long ld_out, unsigned flags) {
    int rc = ci_check(ctx, mode, n_mo, n_terms, coef, ia, ib, out);
    if (rc != OKB_OK) return rc;
    if (npts < 0) return fail(OKB_ERR_ARG, "okb_ci_contract: negative point count");
    if (ld_in < npts || ld_out < npts) return fail(OKB_ERR_ARG, "okb_ci_contract: row stride smaller than the point count");
    if (npts == 0) return OKB_OK;
    // derivative / second-factor sets that come with molistdrv: 3 for JAB, A_NABLA_B and JPAIRS, 1 (optional) for PAIRS
    const int nd = (mode == CI_JAB || mode == CI_ANB || mode == CI_JPAIRS) ? 3 : mode == CI_JABF ? ncomp_in
                   : (mode == CI_PAIRS && molistdrv) ? 1 : 0;
    const bool need_drv = nd > 0;
    if (!molist || (need_drv && mode != CI_PAIRS && !molistdrv)) return fail(OKB_ERR_ARG, "okb_ci_contract: null MO array");
    const bool in_dev = (flags & OKB_FLAG_IN_DEVICE) != 0, out_dev = (flags & OKB_FLAG_OUT_DEVICE) != 0;
    const int nsets = 1 + nd, ncomp = ci_ncomp(mode, n_terms, nd);
    CU(cudaSetDevice(ctx->device));
    // Host inputs: only the rows the term list refers to are staged -- a CI expansion over a few dozen active orbitals of
    // a few hundred MOs moves n_act / n_mo of the bytes over PCIe (the reference's callers pass the MOs of the whole
    // calculation, ci_core.py:85-136).  The terms keep their order and the rows their values: the sums are unchanged.
    std::vector<int> rows, ia2, ib2;
    const int n_all = n_mo;                                   // rows per set of the caller's arrays
    if (!in_dev && n_terms > 0) {
        std::vector<int> idx(n_mo, -1);
        for (int t = 0; t < n_terms; ++t) idx[ia[t]] = idx[ib[t]] = 0;
        for (int m = 0; m < n_mo; ++m)
            if (idx[m] == 0) {
                idx[m] = (int)rows.size();
                rows.push_back(m);
            }
        if (rows.size() * 4 <= (size_t)n_mo * 3) {            // worth one copy per row instead of one strided copy
            ia2.resize(n_terms);
            ib2.resize(n_terms);
            for (int t = 0; t < n_terms; ++t) {
                ia2[t] = idx[ia[t]];
                ib2[t] = idx[ib[t]];
            }
            ia = ia2.data();
            ib = ib2.data();
            n_mo = (int)rows.size();                          // from here on: rows per set of the staged arrays
        } else {
            rows.clear();
        }
    }
    // slab geometry: device-resident inputs and outputs need no staging at all
    const size_t per_pt = (in_dev ? 0 : (size_t)nsets * n_mo * 8) + (out_dev ? 0 : (size_t)ncomp * 8);
    long long slab = npts;
    if (per_pt > 0) {
        slab = (long long)(((size_t)1 << 30) / per_pt) / 1024 * 1024;
        slab = std::max<long long>(1024, std::min(slab, npts));
    }
    const size_t tbytes = ci_terms_bytes(n_terms);
    const long long lds = (slab + 1) & ~1LL;                 // even row stride of the staged MO rows (16-byte aligned)
    const size_t in_bytes = in_dev ? 0 : ((size_t)nsets * n_mo * lds * 8 + 255) / 256 * 256;
    rc = ci_reserve(ctx, tbytes + in_bytes + (out_dev ? 0 : (size_t)ncomp * slab * 8));
    if (rc != OKB_OK) return rc;
    CiParams p{};
    CiDensePlan dense;
    if (flags & OKB_FLAG_CI_FAST) {
        rc = ci_dense_plan(ctx, mode, n_mo, n_terms, coef, ia, ib, &dense);
        if (rc != OKB_OK) return rc;
    }
    if (dense.on) {                                           // the dense form needs no term list on the device
        p.n_terms = n_terms;
    } else {
        rc = ci_upload_terms(ctx, n_terms, coef, ia, ib, &p);
        if (rc != OKB_OK) return rc;
    }
    unsigned char *base = reinterpret_cast<unsigned char *>(ctx->ci_buf);
    double *d_in = reinterpret_cast<double *>(base + tbytes), *d_out = reinterpret_cast<double *>(base + tbytes + in_bytes);
    for (long long s0 = 0; s0 < npts; s0 += slab) {
        const long long sn = std::min(slab, npts - s0);
        if (in_dev) {
            p.mo = molist + s0;
            p.dmo = need_drv ? molistdrv + s0 : nullptr;
            p.ld = ld_in;
        } else {
            if (!rows.empty()) {
                for (int st = 0; st < nsets; ++st) {
                    const double *src = st == 0 ? molist : molistdrv + (size_t)(st - 1) * n_all * ld_in;
                    for (int r = 0; r < n_mo; ++r)
                        CU(cudaMemcpyAsync(d_in + ((size_t)st * n_mo + r) * lds, src + (size_t)rows[r] * ld_in + s0,
                                           (size_t)sn * 8, cudaMemcpyHostToDevice, ctx->stream));
                }
            } else {
                CU(cudaMemcpy2DAsync(d_in, (size_t)lds * 8, molist + s0, (size_t)ld_in * 8, (size_t)sn * 8, n_mo,
                                     cudaMemcpyHostToDevice, ctx->stream));
                if (need_drv)
                    CU(cudaMemcpy2DAsync(d_in + (size_t)n_mo * lds, (size_t)lds * 8, molistdrv + s0, (size_t)ld_in * 8,
                                         (size_t)sn * 8, (size_t)nd * n_mo, cudaMemcpyHostToDevice, ctx->stream));
            }
            ctx->h2d_bytes += (long long)nsets * n_mo * sn * 8;
            p.mo = d_in;
            p.dmo = need_drv ? d_in + (size_t)n_mo * lds : nullptr;
            p.ld = lds;
        }
        p.dstride = (long long)n_mo * p.ld;
        p.npts = sn;
        p.ncomp = mode == CI_JABF ? nd : 3;
        p.out = out_dev ? out + s0 : d_out;
        p.ldo = out_dev ? ld_out : sn;
        rc = (flags & OKB_FLAG_CI_FAST) ? ci_launch_fast(ctx, mode, p, n_mo, nsets, dense) : ci_launch_exact(ctx, mode, p, n_mo, nsets);
        if (rc != OKB_OK) return rc;
        if (!out_dev) {
            CU(cudaMemcpy2DAsync(out + s0, (size_t)ld_out * 8, d_out, (size_t)sn * 8, (size_t)sn * 8, ncomp,
                                 cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h_bytes += (long long)ncomp * sn * 8;
            CU(cudaStreamSynchronize(ctx->stream));          // the staging buffers are reused by the next slab
        }
    }
    if (!out_dev || !in_dev) CU(cudaStreamSynchronize(ctx->stream));
    return OKB_OK;
}

extern "C" int okb_ci_contract(okb_ctx *ctx, int mode, int n_mo, long long npts, long long ld_in,
                               const double *molist, const double *molistdrv, int n_terms, const double *coef,
                               const int *ia, const int *ib, double *out, long long ld_out, unsigned flags) {
    if (mode == CI_JABF) return fail(OKB_ERR_ARG, "okb_ci_contract: use okb_ci_jab_full for mode %d", mode);
    return ci_contract_impl(ctx, mode, n_mo, npts, ld_in, molist, molistdrv, 3, n_terms, coef, ia, ib, out, ld_out, flags);
}

// cy_ci.get_jab_full (cy_ci.pyx:186-202): the pairs n > m of the state basis in the reference's loop order as terms
// (f ImS[n,m], n, m) of the pair kernel; the sums are bit-identical to the reference's
extern "C" int okb_ci_jab_full(okb_ctx *ctx, int nbasis, int ncomp, long long npts, long long ld_in, const double *ImS,
                               const double *chi, const double *dchi, double mu, double *out, long long ld_out,
                               unsigned flags) {
    if (!ctx) return fail(OKB_ERR_ARG, "okb_ci_jab_full: null context");
    if (nbasis <= 0 || ncomp < 1 || ncomp > 3) return fail(OKB_ERR_ARG, "okb_ci_jab_full: nbasis > 0 and 1 <= ncomp <= 3 expected");
    if (!ImS || !chi || !dchi || !out) return fail(OKB_ERR_ARG, "okb_ci_jab_full: null array");
    const double f = 1. / mu;
    std::vector<double> coef;
    std::vector<int> ia, ib;
    for (int n = 0; n < nbasis; ++n)
        for (int m = 0; m < n; ++m) {
            coef.push_back(f * ImS[(size_t)n * nbasis + m]);
            ia.push_back(n);
            ib.push_back(m);
        }
    if (coef.empty()) {                                       // a single state: the sum is empty
        if (flags & OKB_FLAG_OUT_DEVICE) {
            CU(cudaSetDevice(ctx->device));
            CU(cudaMemset2DAsync(out, (size_t)ld_out * 8, 0, (size_t)npts * 8, ncomp, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
        } else {
            for (int c = 0; c < ncomp; ++c) std::fill(out + (size_t)c * ld_out, out + (size_t)c * ld_out + npts, 0.0);
        }
        return OKB_OK;
    }
    return ci_contract_impl(ctx, CI_JABF, nbasis, npts, ld_in, chi, dchi, ncomp, (int)coef.size(), coef.data(), ia.data(),
                            ib.data(), out, ld_out, flags);
}

// Time-dependent contraction out[t][x] = sum_k w[t][k] in[k][x] (okb_td.cuh; cy_ci.get_rho_full / get_j_full with the
// weights packed by the caller).  w: HOST [nt][nk]; in / out: host (staged in slabs) or device (OKB_FLAG_IN_DEVICE /
// OKB_FLAG_OUT_DEVICE) rows of stride ld_in / ld_out >= n.
extern "C" int okb_ci_td(okb_ctx *ctx, int nt, int nk, long long n, const double *w, const double *in, long long ld_in,
                         double *out, long long ld_out, unsigned flags) {
    if (!ctx) return fail(OKB_ERR_ARG, "okb_ci_td: null context");
    if (nt < 0 || nk < 0 || n < 0) return fail(OKB_ERR_ARG, "okb_ci_td: negative extent");
    if (nt == 0 || n == 0) return OKB_OK;
    if (!out || (nk > 0 && (!w || !in))) return fail(OKB_ERR_ARG, "okb_ci_td: null array");
    if (ld_in < n || ld_out < n) return fail(OKB_ERR_ARG, "okb_ci_td: row stride smaller than the point count");
    const bool in_dev = (flags & OKB_FLAG_IN_DEVICE) != 0, out_dev = (flags & OKB_FLAG_OUT_DEVICE) != 0;
    CU(cudaSetDevice(ctx->device));
    const int kp = std::max(4, (nk + 3) / 4 * 4), ntp = (nt + 63) / 64 * 64;
    // several k chunks: every pass over 8 MTB time steps re-reads `in`, so take the taller pass
    // (A/B: OKB_TD_MTB=4|8 forces one)
    static const char *force_mtb = getenv("OKB_TD_MTB");
    int mtb = (kp > TD_KC && nt > 32) ? 8 : 4;
    if (force_mtb && force_mtb[0]) mtb = atoi(force_mtb) == 8 ? 8 : 4;
    std::vector<double> wp((size_t)ntp * kp, 0.0);
    for (int t = 0; t < nt; ++t)
        for (int k = 0; k < nk; ++k) wp[(size_t)t * kp + k] = w[(size_t)t * nk + k];
    const size_t wbytes = (wp.size() * 8 + 255) / 256 * 256;
    const size_t per_pt = (in_dev ? 0 : (size_t)nk * 8) + (out_dev ? 0 : (size_t)nt * 8);
    long long slab = n;
    if (per_pt > 0) {
        slab = (long long)(((size_t)1 << 30) / per_pt) / 1024 * 1024;
        slab = std::max<long long>(1024, std::min(slab, n));
    }
    const long long lds = (slab + 1) & ~1LL;
    const size_t in_bytes = in_dev ? 0 : ((size_t)nk * lds * 8 + 255) / 256 * 256;
    int rc = ci_reserve(ctx, wbytes + in_bytes + (out_dev ? 0 : (size_t)nt * lds * 8));
    if (rc != OKB_OK) return rc;
    unsigned char *base = reinterpret_cast<unsigned char *>(ctx->ci_buf);
    double *d_w = reinterpret_cast<double *>(base), *d_in = reinterpret_cast<double *>(base + wbytes),
           *d_out = reinterpret_cast<double *>(base + wbytes + in_bytes);
    CU(cudaMemcpyAsync(d_w, wp.data(), wp.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (long long)wp.size() * 8;
    const size_t td_bytes = td_smem(kp, mtb);
    CU(cudaFuncSetAttribute(okb_td_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)td_smem(TD_KC, 4)));
    CU(cudaFuncSetAttribute(okb_td_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)td_smem(TD_KC, 8)));
    for (long long s0 = 0; s0 < n; s0 += slab) {
        const long long sn = std::min(slab, n - s0);
        TdParams p{};
        p.w = d_w;
        p.nt = nt; p.nk = nk; p.kp = kp; p.n = sn;
        if (in_dev || nk == 0) {
            p.in = in ? in + s0 : nullptr;
            p.ldi = ld_in;
        } else {
            CU(cudaMemcpy2DAsync(d_in, (size_t)lds * 8, in + s0, (size_t)ld_in * 8, (size_t)sn * 8, nk, cudaMemcpyHostToDevice,
                                 ctx->stream));
            ctx->h2d_bytes += (long long)nk * sn * 8;
            p.in = d_in;
            p.ldi = lds;
        }
        p.out = out_dev ? out + s0 : d_out;
        p.ldo = out_dev ? ld_out : lds;
        p.vec_ok = (reinterpret_cast<uintptr_t>(p.out) % 16 == 0 && p.ldo % 2 == 0) ? 1 : 0;
        const unsigned grid = (unsigned)((sn + TD_P - 1) / TD_P);
        static const char *no_pipe = getenv("OKB_TD_NOPIPE");   // A/B
        static const char *pipe_k = getenv("OKB_TD_PIPE_K");    // A/B: smallest k range that takes the pipelined variant
        const int kmin = pipe_k && pipe_k[0] ? atoi(pipe_k) : TD_KC;
        const bool pipe = kp > kmin && reinterpret_cast<uintptr_t>(p.in) % 16 == 0 && p.ldi % 2 == 0 &&
                          !(no_pipe && no_pipe[0] == '1');
        if (pipe) {                                             // several k chunks: pipelined variant
            CU(cudaFuncSetAttribute(okb_td2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TD2_SMEM));
            okb_td2_kernel<false><<<grid, TD2_NT, TD2_SMEM, ctx->stream>>>(p);
        } else if (mtb == 8) {
            okb_td_kernel<8><<<grid, TD_NT, td_bytes, ctx->stream>>>(p);
        } else {
            okb_td_kernel<4><<<grid, TD_NT, td_bytes, ctx->stream>>>(p);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "okb_td_kernel launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
        ctx->last_kernel = pipe ? "td-dmma/pipe" : mtb == 8 ? "td-dmma/MT64" : "td-dmma/MT32";
        if (!out_dev) {
            CU(cudaMemcpy2DAsync(out + s0, (size_t)ld_out * 8, d_out, (size_t)lds * 8, (size_t)sn * 8, nt, cudaMemcpyDeviceToHost,
                                 ctx->stream));
            ctx->d2h_bytes += (long long)nt * sn * 8;
            CU(cudaStreamSynchronize(ctx->stream));          // the staging buffers are reused by the next slab
        }
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return OKB_OK;
}

extern "C" int okb_eval_ci(okb_ctx *ctx, okb_mo *mo, okb_grid *grid, long long p0, long long p1, int mode,
                           const int *drv_codes, int n_terms, const double *coef, const int *ia, const int *ib,
                           double *out, unsigned flags) {
    if (!mo) return fail(OKB_ERR_ARG, "okb_eval_ci: null MO handle");
    if (!grid) return fail(OKB_ERR_ARG, "okb_eval_ci: null grid handle");
    int rc = ci_check(ctx, mode, mo->n_mo, n_terms, coef, ia, ib, out);
    if (rc != OKB_OK) return rc;
    if (p0 < 0 || p1 > grid->npts || p1 < p0) return fail(OKB_ERR_ARG, "okb_eval_ci: point range outside the grid");
    const long long npts = p1 - p0;
    if (npts == 0) return OKB_OK;
    // JAB / A_NABLA_B / JPAIRS: three derivative sets drv_codes[0..2]; PAIRS: drv_codes (may be NULL) names ONE set for the
    // second factor (NULL or code 0: the MO values themselves)
    const int nd = (mode == CI_JAB || mode == CI_ANB || mode == CI_JPAIRS) ? 3
                   : (mode == CI_PAIRS && drv_codes && drv_codes[0] != 0)  ? 1 : 0;
    const bool need_drv = nd > 0;
    int codes[4] = {0, 1, 2, 3};
    if (need_drv) {
        if (!drv_codes) return fail(OKB_ERR_ARG, "okb_eval_ci: null derivative codes");
        for (int d = 0; d < nd; ++d) {
            if (drv_codes[d] < 1 || drv_codes[d] > 9) return fail(OKB_ERR_ARG, "okb_eval_ci: derivative code %d not in 1..9", drv_codes[d]);
            for (int e = 0; e < d; ++e)
                if (drv_codes[e] == drv_codes[d]) return fail(OKB_ERR_ARG, "okb_eval_ci: duplicate derivative code");
            codes[1 + d] = drv_codes[d];
        }
    }
    const bool out_dev = (flags & OKB_FLAG_OUT_DEVICE) != 0;
    const int n_mo = mo->n_mo, nsets = 1 + nd, ncomp = ci_ncomp(mode, n_terms, 3);
    CU(cudaSetDevice(ctx->device));
    const size_t per_pt = (size_t)nsets * n_mo * 8 + (out_dev ? 0 : (size_t)ncomp * 8);
    long long slab = (long long)(((size_t)1 << 30) / per_pt) / 1024 * 1024;
    slab = std::max<long long>(1024, std::min(slab, npts));
    const size_t tbytes = ci_terms_bytes(n_terms);
    const long long lds = (slab + 1) & ~1LL;                 // even row stride of the staged MO rows (16-byte aligned)
    const size_t in_bytes = ((size_t)nsets * n_mo * lds * 8 + 255) / 256 * 256;
    rc = ci_reserve(ctx, tbytes + in_bytes + (out_dev ? 0 : (size_t)ncomp * slab * 8));
    if (rc != OKB_OK) return rc;
    CiParams p{};
    CiDensePlan dense;
    if (flags & OKB_FLAG_CI_FAST) {
        rc = ci_dense_plan(ctx, mode, n_mo, n_terms, coef, ia, ib, &dense);
        if (rc != OKB_OK) return rc;
    }
    if (dense.on) {                                           // the dense form needs no term list on the device
        p.n_terms = n_terms;
    } else {
        rc = ci_upload_terms(ctx, n_terms, coef, ia, ib, &p);
        if (rc != OKB_OK) return rc;
    }
    unsigned char *base = reinterpret_cast<unsigned char *>(ctx->ci_buf);
    double *d_in = reinterpret_cast<double *>(base + tbytes), *d_out = reinterpret_cast<double *>(base + tbytes + in_bytes);
    for (long long s0 = 0; s0 < npts; s0 += slab) {
        const long long sn = std::min(slab, npts - s0);
        // MOs (+ derivative sets) of the slab, device resident: [nsets][n_mo][sn]
        EvalReq rq{SINK_MO, mo->basis, mo, grid, p0 + s0, p0 + s0 + sn, codes, nsets, d_in, nullptr, nullptr, nullptr,
                   (flags & OKB_FLAG_EXACT_MIXED) | OKB_FLAG_OUT_DEVICE, lds};
        rc = run_eval(ctx, rq);
        if (rc != OKB_OK) return rc;
        p.mo = d_in;
        p.dmo = need_drv ? d_in + (size_t)n_mo * lds : nullptr;
        p.ld = lds;
        p.dstride = (long long)n_mo * lds;
        p.npts = sn;
        p.ncomp = 3;
        p.out = out_dev ? out + s0 : d_out;
        p.ldo = out_dev ? npts : sn;
        rc = (flags & OKB_FLAG_CI_FAST) ? ci_launch_fast(ctx, mode, p, n_mo, nsets, dense) : ci_launch_exact(ctx, mode, p, n_mo, nsets);
        if (rc != OKB_OK) return rc;
        if (!out_dev) {
            CU(cudaMemcpy2DAsync(out + s0, (size_t)npts * 8, d_out, (size_t)sn * 8, (size_t)sn * 8, ncomp,
                                 cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h_bytes += (long long)ncomp * sn * 8;
            CU(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (!out_dev) CU(cudaStreamSynchronize(ctx->stream));
    return OKB_OK;
}

// ---- cube text (output sink) ------------------------------------------------------------------------------------
extern "C" long long okb_cube_body_bytes(int n_sets, long long nx, long long ny, long long nz) {
    if (n_sets <= 0 || nx < 0 || ny < 0 || nz < 0) return -1;
    return nx * ny * cube_row_bytes(nz * n_sets);
}

extern "C" int okb_format_cube(okb_ctx *ctx, const double *data, int n_sets, long long nx, long long ny, long long nz,
                               char *text, long long capacity, unsigned flags) {
    if (!ctx) return fail(OKB_ERR_ARG, "okb_format_cube: null context");
    if (n_sets <= 0 || nx < 0 || ny < 0 || nz < 0) return fail(OKB_ERR_ARG, "okb_format_cube: bad extents");
    if (nz * n_sets > (1ll << 31) - 1 || nz > (1ll << 31) - 1) return fail(OKB_ERR_ARG, "okb_format_cube: row too long");
    const long long nrows = nx * ny, n = nz * n_sets, rb = cube_row_bytes(n), total = nrows * rb;
    if (capacity < total) return fail(OKB_ERR_ARG, "okb_format_cube: text buffer of %lld bytes, %lld needed", capacity, total);
    if (nrows == 0) return OKB_OK;
    if (!data || !text) return fail(OKB_ERR_ARG, "okb_format_cube: null buffer");
    const bool in_dev = (flags & OKB_FLAG_IN_DEVICE) != 0, out_dev = (flags & OKB_FLAG_OUT_DEVICE) != 0;
    CU(cudaSetDevice(ctx->device));
    static_assert(CUBE_SMEM <= 48 * 1024, "the cube kernel fits the default dynamic shared memory limit");
    // rows per slab: staged input + staged text of at most ~256 MB
    const size_t per_row = (in_dev ? 0 : (size_t)n * 8) + (out_dev ? 0 : (size_t)rb);
    long long slab = per_row ? std::max<long long>(1, (long long)(((size_t)256 << 20) / per_row)) : nrows;
    slab = std::min(slab, nrows);
    // a slab's text must start 16-byte aligned only for speed; any offset is handled by the kernel
    const size_t in_bytes = in_dev ? 0 : (((size_t)slab * n * 8 + 255) / 256) * 256;
    const size_t out_bytes = out_dev ? 0 : (size_t)slab * rb + 16;
    if (in_bytes + out_bytes) {
        int rc = ci_reserve(ctx, in_bytes + out_bytes);
        if (rc != OKB_OK) return rc;
    }
    unsigned char *base = reinterpret_cast<unsigned char *>(ctx->ci_buf);
    for (long long r0 = 0; r0 < nrows; r0 += slab) {
        const long long rn = std::min(slab, nrows - r0);
        CubeParams p{};
        if (in_dev) {
            p.data = data + r0 * nz;
            p.set_stride = nrows * nz;
        } else {
            double *d_in = reinterpret_cast<double *>(base);
            CU(cudaMemcpy2DAsync(d_in, (size_t)rn * nz * 8, data + r0 * nz, (size_t)nrows * nz * 8, (size_t)rn * nz * 8,
                                 n_sets, cudaMemcpyHostToDevice, ctx->stream));
            ctx->h2d_bytes += (long long)n_sets * rn * nz * 8;
            p.data = d_in;
            p.set_stride = rn * nz;
        }
        p.nrows = rn;
        p.nz = (int)nz;
        p.n_sets = n_sets;
        p.row0 = r0;
        p.text = out_dev ? text + r0 * rb : reinterpret_cast<char *>(base + in_bytes);
        p.total_values = rn * n;
        const long long nblk = (p.total_values + CUBE_VPB - 1) / CUBE_VPB;
        okb_cube_kernel<<<(unsigned)nblk, CUBE_NT, CUBE_SMEM, ctx->stream>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "cube kernel launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
        ctx->last_kernel = "cube/format";
        if (!out_dev) {
            CU(cudaMemcpyAsync(text + r0 * rb, p.text, (size_t)rn * rb, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h_bytes += rn * rb;
            CU(cudaStreamSynchronize(ctx->stream));
        }
    }
    if (!out_dev) CU(cudaStreamSynchronize(ctx->stream));
    return OKB_OK;
}

// ---- cy_core drop-ins -----------------------------------------------------------------------------------------
extern "C" int okb_aocreator(okb_ctx *ctx, const int *lxlylz, const int *assign, const double *ao_coeffs,
                             const int *pnum_list, const double *geo_spec, const int *atom_indices,
                             int n_cont, int n_cart, int n_prim, int n_atoms, const double *x,
                             const double *y, const double *z, long long npts, int drv, int is_normalized,
                             unsigned flags, double *out) {
    if (drv < 0 || drv > 9) return fail(OKB_ERR_ARG, "okb_aocreator: drv=%d not in 0..9", drv);
    if (!out) return fail(OKB_ERR_ARG, "okb_aocreator: null output");
    if (flags & OKB_FLAG_OUT_DEVICE) return fail(OKB_ERR_ARG, "okb_aocreator works on host buffers");
    okb_basis *b = nullptr;
    okb_grid *g = nullptr;
    int rc = okb_basis_create(ctx, lxlylz, assign, ao_coeffs, pnum_list, geo_spec, atom_indices, n_cont, n_cart,
                              n_prim, n_atoms, is_normalized, nullptr, &b);
    if (rc == OKB_OK) rc = okb_grid_vector(ctx, x, y, z, npts, 0, &g);
    if (rc == OKB_OK) rc = okb_eval_ao(ctx, b, g, 0, npts, &drv, 1, out, flags);
    okb_grid_destroy(g);
    okb_basis_destroy(b);
    return rc;
}

extern "C" int okb_lcreator(okb_ctx *ctx, double *ao_list, long long row_stride, const int *lxlylz,
                            const double *coeff_list, const double *at_pos, const double *x, const double *y,
                            const double *z, long long npts, int ao_num, int pnum, int drv, int is_normalized,
                            unsigned flags) {
    if (!ao_list) return fail(OKB_ERR_ARG, "okb_lcreator: null output");
    if (row_stride < npts) return fail(OKB_ERR_ARG, "okb_lcreator: row stride smaller than npts");
    const int atom = 0;
    std::vector<double> tmp((size_t)ao_num * npts);
    int rc = okb_aocreator(ctx, lxlylz, &ao_num, coeff_list, &pnum, at_pos, &atom, 1, ao_num, pnum, 1, x, y, z,
                           npts, drv, is_normalized, flags, tmp.data());
    if (rc != OKB_OK) return rc;
    for (int r = 0; r < ao_num; ++r)
        memcpy(ao_list + (size_t)r * row_stride, tmp.data() + (size_t)r * npts, sizeof(double) * npts);
    return OKB_OK;
}

// cy_core.mocreator (cy_core.pyx:82-101): mo[i][x] = sum_j coeffs[i][j] ao[j][x] -- the dense contraction of okb_td.cuh
// (FP64 tensor cores, AOs in the reference's order) on host arrays staged in slabs
extern "C" int okb_mocreator(okb_ctx *ctx, const double *ao, const double *coeffs, int n_ao, long long npts,
                             int n_mo, double *mo) {
    if (!ctx || !ao || !coeffs || !mo) return fail(OKB_ERR_ARG, "okb_mocreator: null argument");
    if (n_ao <= 0 || n_mo <= 0 || npts <= 0) return fail(OKB_ERR_ARG, "okb_mocreator: empty operand");
    return okb_ci_td(ctx, n_mo, n_ao, npts, coeffs, ao, npts, mo, npts, 0);
}

// ---- analytic overlap matrix (cy_overlap.aooverlap, cy_overlap.pyx:75-156) ---------------------------------------
// Same argument list as the Cython function (contractions: assign[i] functions x pnum_list[i] primitives on atom
// atom_indices[i]); aoom is a HOST array [ao_num][ao_num].
extern "C" int okb_aooverlap(okb_ctx *ctx, const double *geo_a, const double *geo_b, int n_atom, const int *lxlylz_a,
                             const int *lxlylz_b, int ao_num, const int *assign, const double *ao_coeffs,
                             const int *pnum_list, const int *atom_indices, int n_cont, int drv, int is_normalized,
                             double *aoom) {
    if (!ctx) return fail(OKB_ERR_ARG, "okb_aooverlap: null context");
    if (!geo_a || !geo_b || !lxlylz_a || !lxlylz_b || !assign || !ao_coeffs || !pnum_list || !atom_indices || !aoom)
        return fail(OKB_ERR_ARG, "okb_aooverlap: null argument");
    if (drv < 0 || drv > 3) return fail(OKB_ERR_ARG, "okb_aooverlap: drv must be 0..3 (only first derivatives)");
    if (ao_num <= 0 || n_cont <= 0 || n_atom <= 0) return fail(OKB_ERR_ARG, "okb_aooverlap: empty basis");
    std::vector<int> fn_atom, fn_e0, fn_ne;
    std::vector<double> e_alpha, e_c, e_n;
    int c_ao = 0, c_p = 0;
    for (int i = 0; i < n_cont; ++i) {
        if (assign[i] < 0 || pnum_list[i] < 0) return fail(OKB_ERR_ARG, "okb_aooverlap: negative count in contraction %d", i);
        if (atom_indices[i] < 0 || atom_indices[i] >= n_atom) return fail(OKB_ERR_ARG, "okb_aooverlap: atom index of contraction %d", i);
        for (int f = 0; f < assign[i]; ++f) {
            if (c_ao + f >= ao_num) return fail(OKB_ERR_ARG, "okb_aooverlap: assign does not match the %d functions", ao_num);
            const int *l = lxlylz_a + 3 * (c_ao + f);
            for (int r = 0; r < 3; ++r) {
                const int lb = lxlylz_b[3 * (c_ao + f) + r];
                if (l[r] < 0 || lb < 0 || l[r] + lb + 1 > 15) return fail(OKB_ERR_UNSUPPORTED, "okb_aooverlap: exponents of function %d", c_ao + f);
            }
            fn_atom.push_back(atom_indices[i]);
            fn_e0.push_back((int)e_alpha.size());
            fn_ne.push_back(pnum_list[i]);
            for (int q = 0; q < pnum_list[i]; ++q) {
                const double alpha = ao_coeffs[2 * (size_t)(c_p + q)];
                e_alpha.push_back(alpha);
                e_c.push_back(ao_coeffs[2 * (size_t)(c_p + q) + 1]);
                e_n.push_back(okb_aonorm(l[0], l[1], l[2], alpha, is_normalized));
            }
        }
        c_ao += assign[i];
        c_p += pnum_list[i];
    }
    if (c_ao != ao_num) return fail(OKB_ERR_ARG, "okb_aooverlap: assign sums to %d functions, %d given", c_ao, ao_num);
    CU(cudaSetDevice(ctx->device));
    const size_t nf = (size_t)ao_num, ne = std::max<size_t>(e_alpha.size(), 1);
    // one scratch allocation: [geo_a | geo_b | e_alpha | e_c | e_n | aoom | la | lb | fn_atom | fn_e0 | fn_ne]
    const size_t nd = 6 * (size_t)n_atom + 3 * ne + nf * nf, ni = 6 * nf + 3 * nf;
    unsigned char *buf = nullptr;
    CU(cudaMalloc(&buf, nd * 8 + ni * 4));
    double *d = reinterpret_cast<double *>(buf);
    int *di = reinterpret_cast<int *>(buf + nd * 8);
    OvParams p{};
    auto up_d = [&](const double *src, size_t n, const double **dst) {
        cudaMemcpyAsync(d, src, n * 8, cudaMemcpyHostToDevice, ctx->stream);
        *dst = d;
        d += n;
    };
    auto up_i = [&](const int *src, size_t n, const int **dst) {
        cudaMemcpyAsync(di, src, n * 4, cudaMemcpyHostToDevice, ctx->stream);
        *dst = di;
        di += n;
    };
    up_d(geo_a, 3 * (size_t)n_atom, &p.geo_a);
    up_d(geo_b, 3 * (size_t)n_atom, &p.geo_b);
    up_d(e_alpha.data(), e_alpha.size(), &p.e_alpha);
    d = const_cast<double *>(p.e_alpha) + ne;
    up_d(e_c.data(), e_c.size(), &p.e_c);
    d = const_cast<double *>(p.e_c) + ne;
    up_d(e_n.data(), e_n.size(), &p.e_n);
    d = const_cast<double *>(p.e_n) + ne;
    p.aoom = d;
    up_i(lxlylz_a, 3 * nf, &p.la);
    up_i(lxlylz_b, 3 * nf, &p.lb);
    up_i(fn_atom.data(), nf, &p.fn_atom);
    up_i(fn_e0.data(), nf, &p.fn_e0);
    up_i(fn_ne.data(), nf, &p.fn_ne);
    p.n_fn = ao_num;
    p.drv = drv;
    const long long total = (long long)ao_num * ao_num;
    okb_overlap_kernel<<<(unsigned)((total + 127) / 128), 128, 0, ctx->stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(aoom, p.aoom, nf * nf * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(OKB_ERR_CUDA, "okb_aooverlap: %s", cudaGetErrorString(e));
    ctx->launches++;
    ctx->last_kernel = "okb_overlap_kernel";
    ctx->h2d_bytes += (long long)(nd * 8 + ni * 4 - nf * nf * 8);
    ctx->d2h_bytes += (long long)(nf * nf * 8);
    return OKB_OK;
}

// ---- FP64 peak measurement (the roofline denominator MEASURED_PEAKS.json does not carry) --------------
extern "C" int okb_measure_fp64(okb_ctx *ctx, int kind, double min_seconds, double *tflops, double *ms_out) {
    if (!ctx || !tflops) return fail(OKB_ERR_ARG, "okb_measure_fp64: null argument");
    if (kind < 0 || kind > 2) return fail(OKB_ERR_ARG, "okb_measure_fp64: kind must be 0 (DFMA), 1 (DMMA) or 2 (mixed)");
    CU(cudaSetDevice(ctx->device));
    double *sink = nullptr;
    CU(cudaMalloc(&sink, 8));
    const int iters = 4096, grid = ctx->sm_count * 8, block = 256;
    // flops per warp-iteration: DFMA 16 chains x 2 x 32 lanes; DMMA 8 mma x (8*8*4*2)
    const double per_warp_iter_dfma = 16.0 * 2.0 * 32.0, per_warp_iter_dmma = 8.0 * 512.0;
    const double warps = (double)grid * block / 32.0;
    double flops;
    if (kind == 0) flops = warps * iters * per_warp_iter_dfma;
    else if (kind == 1) flops = warps * iters * per_warp_iter_dmma;
    else flops = 0.5 * warps * iters * (per_warp_iter_dfma + per_warp_iter_dmma);
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    okb_fp64_peak_kernel<<<grid, block, 0, ctx->stream>>>(sink, iters, kind);   // warm-up
    CU(cudaStreamSynchronize(ctx->stream));
    double result_ms = 0.0, result_flops = 0.0;
    if (min_seconds <= 0.0) {           // burst: best single launch of 4
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CU(cudaEventRecord(e0, ctx->stream));
            okb_fp64_peak_kernel<<<grid, block, 0, ctx->stream>>>(sink, iters, kind);
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaEventSynchronize(e1));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, e0, e1));
            best = std::min(best, ms);
            ctx->launches++;
        }
        result_ms = best;
        result_flops = flops;
    } else {                            // sustained: back-to-back launches for >= min_seconds
        int n = 0;
        CU(cudaEventRecord(e0, ctx->stream));
        float ms = 0.f;
        do {
            for (int k = 0; k < 8; ++k) okb_fp64_peak_kernel<<<grid, block, 0, ctx->stream>>>(sink, iters, kind);
            n += 8;
            ctx->launches += 8;
            CU(cudaEventRecord(e1, ctx->stream));
            CU(cudaEventSynchronize(e1));
            CU(cudaEventElapsedTime(&ms, e0, e1));
        } while (ms < min_seconds * 1e3 && n < 100000);
        result_ms = ms;
        result_flops = flops * n;
    }
    CU(cudaGetLastError());
    *tflops = result_flops / (result_ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = result_ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return OKB_OK;
}
