// Stage A of the 3-D step, FP64 parity flavour, SYMMETRIC variant.  The reference's d^2 chain is
// symmetric in (i, j) bit for bit (commutative products, hi/lo ordering by index) and its pair term is
// exactly antisymmetric, so every unordered pair is evaluated once: 29 FP64-pipe instructions per
// pair = 14.5 per ordered interaction instead of 25.  Same decomposition as accel_sym64.cu: a CTA
// owns a 512-object row block in registers (2 + 2 objects per lane, 8 warps) and walks column blocks
// staged in shared memory by the TMA engine's bulk copy; inside a 64 x 64 sub-block the 32 lanes
// rotate over the column chunks, the column sums travelling through the warp by SHFL; column
// partials go to `colbuf` (one row of three planes per row block), row partials to `partial`, and
// k3_sym64_reduce adds both in a fixed order (deterministic) and integrates.
// Replaces acc_blas on (N,3) input (cosmosim/util/blas.py:4-41, call site universe.py:101) + the
// integrator (universe.py:103-104).
#include <math.h>

#include "state3d.cuh"

namespace cosmo {
namespace s3 {

constexpr int T_BLOCK = 512;
constexpr int T_GROUP = 64;
constexpr int T_WARPS = 8;
constexpr int T_THREADS = T_WARPS * 32;
// dynamic shared memory: 2 stages x 5 planes (x, y, z, h, m) + 3 planes of column sums + 2 mbarriers
constexpr size_t T_SMEM = (size_t)(2 * 5 + 3) * T_BLOCK * sizeof(double) + 2 * sizeof(uint64_t);

int configure_sym_kernels();

// mangled names of the two symmetric kernels: their post-scheduled copies (sym_tuned.cu, sass_tune.py)
static const char *const kTuned64 = "_ZN5cosmo2s316k3_accel_f64_symE12cosmo3_state14cosmo3_scratchiii";
static const char *const kTuned32 = "_ZN5cosmo2s316k3_accel_f32_symE12cosmo3_state14cosmo3_scratchiii";

// launch the tuned copy if the build produced one, else the kernel compiled into the library
static int launch_sym3(const char *mangled, const void *builtin, const char *what, dim3 grid, int threads, size_t smem,
                       const cosmo3_state &st, const cosmo3_scratch &sc, int sh_i, int sh_n, int chunk, cudaStream_t s) {
    const void *tuned = tuned_kernel(mangled, (int)smem);
    cosmo3_state st_v = st;
    cosmo3_scratch sc_v = sc;
    void *args[] = {&st_v, &sc_v, &sh_i, &sh_n, &chunk};
    return check_cuda(cudaLaunchKernel(tuned ? tuned : builtin, grid, dim3(threads), args, smem, s), what);
}

__device__ __forceinline__ double shfl_f64(double v, int src_lane) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src_lane);
    hi = __shfl_sync(0xffffffffu, hi, src_lane);
    return __hiloint2double(hi, lo);
}

// One unordered pair {i, j}.  DIAG: i and j come from the same block (index order varies, self pairs
// occur, only the row side is accumulated).  CHECK0: honour the where= mask (blas.py:31).
template <bool DIAG, bool CHECK0>
__device__ __forceinline__ void pair3_sym(const double (&pi)[3], const double (&qi)[3], double hi, double nmi,
                                          int i, double xj, double yj, double zj, double hj, double mj, int j,
                                          double (&ai)[3], double &axj, double &ayj, double &azj) {
    const double c = chain_c(qi[0], qi[1], qi[2], xj, yj, zj);
    double d2;
    if (!DIAG) {
        d2 = __dadd_rn(__dadd_rn(c, -hj), -hi);   // every column index is above every row index
    } else {
        const bool ge = j >= i;
        d2 = __dadd_rn(__dadd_rn(c, -(ge ? hj : hi)), -(ge ? hi : hj));
    }
    const double dx = xj - pi[0], dy = yj - pi[1], dz = zj - pi[2];
    double y = rsqrt_approx_f64(d2);
    const double t = d2 * y;
    const double e = __fma_rn(-t, y, 1.0);
    const double p = __fma_rn(e, 0.375, 0.5);
    const double ye = y * e;
    y = __fma_rn(ye, p, y);
    const double y3 = (y * y) * y;
    double sj = mj * y3;
    double nsi = nmi * y3;
    if (CHECK0) {
        const bool zero = ((__double2hiint(d2) << 1) | __double2loint(d2)) == 0;
        sj = zero ? mj : sj;
        nsi = zero ? nmi : nsi;
    }
    ai[0] = __fma_rn(dx, sj, ai[0]);
    ai[1] = __fma_rn(dy, sj, ai[1]);
    ai[2] = __fma_rn(dz, sj, ai[2]);
    if (!DIAG) {
        axj = __fma_rn(dx, nsi, axj);
        ayj = __fma_rn(dy, nsi, ayj);
        azj = __fma_rn(dz, nsi, azj);
    }
}

// 64 row objects (2 per lane) x 64 column objects (chunk c = columns 2c, 2c+1): 32 rotation steps
template <bool DIAG, bool CHECK0>
__device__ __forceinline__ void round3_sym(const double *__restrict__ gp, int lane, int jbase,
                                           const double (&pi)[2][3], const double (&qi)[2][3],
                                           const double (&hi)[2], const double (&nmi)[2], const int (&ii)[2],
                                           double (&ra)[2][3], double (&tc)[2][3]) {
    // gp points at plane 0 of the group inside the stage; planes are T_BLOCK doubles apart
#pragma unroll 1
    for (int s = 0; s < 32; ++s) {
        const int c = (lane + s) & 31;
        const double2 X = *reinterpret_cast<const double2 *>(gp + 2 * c);
        const double2 Y = *reinterpret_cast<const double2 *>(gp + T_BLOCK + 2 * c);
        const double2 Z = *reinterpret_cast<const double2 *>(gp + 2 * T_BLOCK + 2 * c);
        const double2 H = *reinterpret_cast<const double2 *>(gp + 3 * T_BLOCK + 2 * c);
        const double2 M = *reinterpret_cast<const double2 *>(gp + 4 * T_BLOCK + 2 * c);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            pair3_sym<DIAG, CHECK0>(pi[k], qi[k], hi[k], nmi[k], ii[k], X.x, Y.x, Z.x, H.x, M.x, jbase + 2 * c,
                                    ra[k], tc[0][0], tc[0][1], tc[0][2]);
            pair3_sym<DIAG, CHECK0>(pi[k], qi[k], hi[k], nmi[k], ii[k], X.y, Y.y, Z.y, H.y, M.y, jbase + 2 * c + 1,
                                    ra[k], tc[1][0], tc[1][1], tc[1][2]);
        }
        if (!DIAG) {
            const int src = (lane + 1) & 31;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int d = 0; d < 3; ++d) tc[q][d] = shfl_f64(tc[q][d], src);
        }
    }
}

__global__ void __launch_bounds__(T_THREADS, 2)
k3_accel_f64_sym(cosmo3_state st, cosmo3_scratch sc, int shard_index, int shard_count, int chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_p = reinterpret_cast<double *>(smem_raw);          // [2][5][T_BLOCK]
    double *s_c = s_p + 2 * 5 * T_BLOCK;                         // [3][T_BLOCK]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_c + 3 * T_BLOCK);

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int nblk = (n + T_BLOCK - 1) / T_BLOCK;
    const int I = shard_index + blockIdx.x * shard_count;   // row blocks dealt round-robin to the ranks
    if (I >= nblk) return;
    const int J0 = I + blockIdx.y * chunk;
    if (J0 >= nblk) return;
    const int J1 = min(nblk, J0 + chunk);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const size_t cap = (size_t)st.cap;

    double pi[2][3], qi[2][3], hi[2], nmi[2], ra[2][3];
    int ii[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int r = I * T_BLOCK + w * T_GROUP + k * 32 + lane;   // slots past n are massless pads (k3_prep)
        ii[k] = r;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            pi[k][d] = st.pos[d * cap + r];
            qi[k][d] = -2.0 * pi[k][d];
            asm volatile("" : "+d"(qi[k][d]));
            ra[k][d] = 0.0;
        }
        hi[k] = sc.h[r];
        nmi[k] = -st.mass[r];
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t kBytes = T_BLOCK * 5 * sizeof(double);
    auto issue = [&](int J, int b) {
        const size_t j0 = (size_t)J * T_BLOCK;
        double *dst = s_p + (size_t)b * 5 * T_BLOCK;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kBytes);
        bulk_g2s(dst, st.pos + j0, T_BLOCK * sizeof(double), &s_bar[b]);
        bulk_g2s(dst + T_BLOCK, st.pos + cap + j0, T_BLOCK * sizeof(double), &s_bar[b]);
        bulk_g2s(dst + 2 * T_BLOCK, st.pos + 2 * cap + j0, T_BLOCK * sizeof(double), &s_bar[b]);
        bulk_g2s(dst + 3 * T_BLOCK, sc.h + j0, T_BLOCK * sizeof(double), &s_bar[b]);
        bulk_g2s(dst + 4 * T_BLOCK, st.mass + j0, T_BLOCK * sizeof(double), &s_bar[b]);
    };
    if (tid == 0) issue(J0, 0);
    double *colrow = sc.colbuf + (size_t)blockIdx.x * 3 * cap;

    for (int J = J0; J < J1; ++J) {
        const int it_j = J - J0;
        const int b = it_j & 1;
        if (tid == 0 && J + 1 < J1) issue(J + 1, b ^ 1);
        mbar_wait(&s_bar[b], (it_j >> 1) & 1);
        const double *stage = s_p + (size_t)b * 5 * T_BLOCK;
        const bool diag = (J == I);
        double ba[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};   // block-local row sums (two-level summation)
        for (int attempt = 0; attempt < 2; ++attempt) {
            for (int q = tid; q < 3 * T_BLOCK; q += T_THREADS) s_c[q] = 0.0;
            __syncthreads();
            for (int it = 0; it < T_WARPS; ++it) {
                const int g = (w + it) & (T_WARPS - 1);
                const double *gp = stage + g * T_GROUP;
                const int jbase = J * T_BLOCK + g * T_GROUP;
                double tc[2][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
                if (diag) round3_sym<true, true>(gp, lane, jbase, pi, qi, hi, nmi, ii, ba, tc);
                else if (attempt == 0) round3_sym<false, false>(gp, lane, jbase, pi, qi, hi, nmi, ii, ba, tc);
                else round3_sym<false, true>(gp, lane, jbase, pi, qi, hi, nmi, ii, ba, tc);
                if (!diag) {   // lane l is home to chunk l: column objects 2l, 2l+1 of group g
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        double2 *cell = reinterpret_cast<double2 *>(&s_c[d * T_BLOCK + g * T_GROUP + 2 * lane]);
                        double2 v = *cell;
                        v.x += tc[0][d];
                        v.y += tc[1][d];
                        *cell = v;
                    }
                }
                __syncthreads();
            }
            if (diag || attempt == 1) break;
            // the off-diagonal fast flavour skips the d2 == 0 mask; a non-finite sum (practically never)
            // means such an entry occurred: redo this block pair with the mask honoured
            bool bad = false;
#pragma unroll
            for (int k = 0; k < 2; ++k) bad = bad || !(isfinite(ba[k][0]) && isfinite(ba[k][1]) && isfinite(ba[k][2]));
            for (int q = tid; q < 3 * T_BLOCK; q += T_THREADS) bad = bad || !isfinite(s_c[q]);
            if (!__syncthreads_or(bad)) break;
#pragma unroll
            for (int k = 0; k < 2; ++k) ba[k][0] = ba[k][1] = ba[k][2] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) { ra[k][0] += ba[k][0]; ra[k][1] += ba[k][1]; ra[k][2] += ba[k][2]; }
        if (!diag) {
            for (int q = tid; q < 3 * T_BLOCK; q += T_THREADS) {
                const int d = q / T_BLOCK, o = q - d * T_BLOCK;
                colrow[d * cap + (size_t)J * T_BLOCK + o] = s_c[q];
            }
        }
        __syncthreads();
    }
    double *part = sc.partial + (size_t)blockIdx.y * 3 * cap;
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d) part[d * cap + ii[k]] = ra[k][d];
}

// row partials of the object's own row block (in chunk order) + column partials of every row block
// above it (in block order): a fixed summation order, independent of scheduling
__global__ void __launch_bounds__(256)
k3_sym64_reduce(cosmo3_state st, cosmo3_scratch sc, int shard_index, int shard_count, int chunk, double G, double dt,
                int store_acc, int defer) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t cap = (size_t)st.cap;
    const int nblk = (n + T_BLOCK - 1) / T_BLOCK;
    const int K = k / T_BLOCK;
    double a[3] = {0.0, 0.0, 0.0};
    if (K % shard_count == shard_index) {
        const int nchunk = (nblk - K + chunk - 1) / chunk;
        for (int c = 0; c < nchunk; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) a[d] += sc.partial[((size_t)c * 3 + d) * cap + k];
    }
    int r = 0;
    for (int I = shard_index; I < K; I += shard_count, ++r)
#pragma unroll
        for (int d = 0; d < 3; ++d) a[d] += __ldcs(&sc.colbuf[((size_t)r * 3 + d) * cap + k]);
    if (defer) {
#pragma unroll
        for (int d = 0; d < 3; ++d) sc.acc[d * cap + k] = a[d];
    } else {
        integrate_store3(st, sc, k, a[0], a[1], a[2], G, dt, store_acc != 0);
    }
}

// multi-GPU: every rank integrates all rows from the all-reduced sums (identical on every rank)
__global__ void __launch_bounds__(256) k3_integrate_all(cosmo3_state st, cosmo3_scratch sc, double G, double dt,
                                                        int store_acc) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t cap = (size_t)st.cap;
    integrate_store3(st, sc, k, sc.acc[k], sc.acc[cap + k], sc.acc[2 * cap + k], G, dt, store_acc != 0);
}

int launch_integrate3(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p, cudaStream_t s) {
    if (p.n_upper <= 0) return COSMO_OK;
    k3_integrate_all<<<(p.n_upper + 255) / 256, 256, 0, s>>>(st, sc, p.G, p.dt, (p.opts & COSMO3_OPT_STORE_ACC) ? 1 : 0);
    return COSMO_LAUNCH_CHECK("k3_integrate_all");
}

int launch_accel3_sym64(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                        cudaStream_t s) {
    const int n = p.n_upper;
    if (n <= 0) return COSMO_OK;
    const int nblk = (n + T_BLOCK - 1) / T_BLOCK;
    int chunk = p.sym_chunk > 0 ? p.sym_chunk : 4;
    while ((nblk + chunk - 1) / chunk > sc.jsplit_max) ++chunk;
    const int sh_n = p.shard_count > 0 ? p.shard_count : 1;
    const int sh_i = p.shard_count > 0 ? p.shard_index : 0;
    const int rows_local = (nblk - sh_i + sh_n - 1) / sh_n;
    if (!sc.colbuf || (long long)rows_local > sc.colbuf_rows) {
        set_error("symmetric 3-D accel: colbuf has %lld rows, %d needed", (long long)sc.colbuf_rows, rows_local);
        return COSMO_E_ARG;
    }
    COSMO_TRY(configure_sym_kernels());
    if (rows_local > 0) {
        dim3 grid(rows_local, (nblk + chunk - 1) / chunk);
        COSMO_TRY(launch_sym3(kTuned64, (const void *)k3_accel_f64_sym, "k3_accel_f64_sym", grid, T_THREADS, T_SMEM, st, sc,
                              sh_i, sh_n, chunk, s));
    }
    k3_sym64_reduce<<<(n + 255) / 256, 256, 0, s>>>(st, sc, sh_i, sh_n, chunk, p.G, p.dt,
                                                    (p.opts & COSMO3_OPT_STORE_ACC) ? 1 : 0,
                                                    (p.opts & COSMO3_OPT_DEFER_INTEGRATE) ? 1 : 0);
    return COSMO_LAUNCH_CHECK("k3_sym64_reduce");
}


// ---------------------------------------------------------------------------------------
// FP32 fast flavour, SYMMETRIC: direct differences on packed f32x2, every unordered pair once --
// 16 packed FMA-pipe instructions per two pairs = 8 per ORDERED interaction (scalar-equivalent)
// and half a MUFU.RSQ instead of 12 and 1.  Decomposition of accel_sym.cu: 1024-object blocks,
// 8 warps, 4 row objects per lane, column chunks of 4 rotating through the warp (12 SHFL per step).
// Column sums in float planes of `colbuf`, row sums folded into FP64 once per column block.
// ---------------------------------------------------------------------------------------
constexpr int U_BLOCK = 1024;
constexpr int U_GROUP = 128;
constexpr int U_WARPS = 8;
constexpr int U_THREADS = U_WARPS * 32;
// dynamic shared memory: 2 stages x 4 planes (x, y, z, m) + 3 planes of column sums (floats) + 2 mbarriers
constexpr size_t U_SMEM = (size_t)(2 * 4 + 3) * U_BLOCK * sizeof(float) + 2 * sizeof(uint64_t);
#define COSMO3_F32_TINY_SYM 8.673617379884035e-19f /* 2^-60 softening (scaled units), see state3d.cu */

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src_lane) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src_lane);
    hi = __shfl_sync(0xffffffffu, hi, src_lane);
    return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(U_THREADS, 2)
k3_accel_f32_sym(cosmo3_state st, cosmo3_scratch sc, int shard_index, int shard_count, int chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *s_p = reinterpret_cast<float *>(smem_raw);            // [2][4][U_BLOCK]
    float *s_c = s_p + 2 * 4 * U_BLOCK;                          // [3][U_BLOCK]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_c + 3 * U_BLOCK);

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int nblk = (n + U_BLOCK - 1) / U_BLOCK;
    const int I = shard_index + blockIdx.x * shard_count;
    if (I >= nblk) return;
    const int J0 = I + blockIdx.y * chunk;
    if (J0 >= nblk) return;
    const int J1 = min(nblk, J0 + chunk);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const size_t cap = (size_t)st.cap;

    float xi[4], yi[4], zi[4], nmi[4];
    double ra[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = I * U_BLOCK + w * U_GROUP + k * 32 + lane;   // slots past n: zero mass (k3_prep)
        xi[k] = sc.pk[r];
        yi[k] = sc.pk[cap + r];
        zi[k] = sc.pk[2 * cap + r];
        nmi[k] = -sc.pk[3 * cap + r];
        ra[k][0] = 0.0; ra[k][1] = 0.0; ra[k][2] = 0.0;
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t kBytes = U_BLOCK * 4 * sizeof(float);
    auto issue = [&](int J, int b) {
        const size_t j0 = (size_t)J * U_BLOCK;
        float *dst = s_p + (size_t)b * 4 * U_BLOCK;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kBytes);
#pragma unroll
        for (int d = 0; d < 4; ++d) bulk_g2s(dst + d * U_BLOCK, sc.pk + d * cap + j0, U_BLOCK * sizeof(float), &s_bar[b]);
    };
    if (tid == 0) issue(J0, 0);
    const uint64_t eps2 = pack2(COSMO3_F32_TINY_SYM, COSMO3_F32_TINY_SYM);
    float *colrow = reinterpret_cast<float *>(sc.colbuf) + (size_t)blockIdx.x * 3 * cap;

    for (int J = J0; J < J1; ++J) {
        const int it_j = J - J0;
        const int b = it_j & 1;
        if (tid == 0 && J + 1 < J1) issue(J + 1, b ^ 1);
        for (int q = tid; q < 3 * U_BLOCK; q += U_THREADS) s_c[q] = 0.f;
        mbar_wait(&s_bar[b], (it_j >> 1) & 1);
        __syncthreads();
        const float *stage = s_p + (size_t)b * 4 * U_BLOCK;
        const bool diag = (J == I);
        uint64_t rx2[4], ry2[4], rz2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { rx2[k] = 0ull; ry2[k] = 0ull; rz2[k] = 0ull; }

        for (int it = 0; it < U_WARPS; ++it) {
            const int g = (w + it) & (U_WARPS - 1);
            const float *gp = stage + g * U_GROUP;
            uint64_t tx2[2] = {0ull, 0ull}, ty2[2] = {0ull, 0ull}, tz2[2] = {0ull, 0ull};
#pragma unroll 1
            for (int s = 0; s < 32; ++s) {
                const int c = (lane + s) & 31;
                const float4 X = *reinterpret_cast<const float4 *>(gp + 4 * c);
                const float4 Y = *reinterpret_cast<const float4 *>(gp + U_BLOCK + 4 * c);
                const float4 Z = *reinterpret_cast<const float4 *>(gp + 2 * U_BLOCK + 4 * c);
                const float4 M = *reinterpret_cast<const float4 *>(gp + 3 * U_BLOCK + 4 * c);
                const uint64_t xj2[2] = {pack2(X.x, X.y), pack2(X.z, X.w)};
                const uint64_t yj2[2] = {pack2(Y.x, Y.y), pack2(Y.z, Y.w)};
                const uint64_t zj2[2] = {pack2(Z.x, Z.y), pack2(Z.z, Z.w)};
                const uint64_t mj2[2] = {pack2(M.x, M.y), pack2(M.z, M.w)};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t xik = pack2(xi[k], xi[k]), yik = pack2(yi[k], yi[k]), zik = pack2(zi[k], zi[k]);
                    const uint64_t nmik = pack2(nmi[k], nmi[k]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t dx = sub2(xj2[h], xik);
                        const uint64_t dy = sub2(yj2[h], yik);
                        const uint64_t dz = sub2(zj2[h], zik);
                        const uint64_t r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, eps2)));
                        float r2a, r2b;
                        unpack2(r2, r2a, r2b);
                        const uint64_t rinv = pack2(rsqrt_approx_f32(r2a), rsqrt_approx_f32(r2b));
                        const uint64_t rinv3 = mul2(mul2(rinv, rinv), rinv);
                        const uint64_t sj = mul2(mj2[h], rinv3);
                        const uint64_t nsi = mul2(rinv3, nmik);
                        rx2[k] = fma2(dx, sj, rx2[k]);
                        ry2[k] = fma2(dy, sj, ry2[k]);
                        rz2[k] = fma2(dz, sj, rz2[k]);
                        tx2[h] = fma2(dx, nsi, tx2[h]);
                        ty2[h] = fma2(dy, nsi, ty2[h]);
                        tz2[h] = fma2(dz, nsi, tz2[h]);
                    }
                }
                const int src = (lane + 1) & 31;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    tx2[h] = shfl_u64(tx2[h], src);
                    ty2[h] = shfl_u64(ty2[h], src);
                    tz2[h] = shfl_u64(tz2[h], src);
                }
            }
            if (!diag) {   // home again: lane l holds the sums of chunk l of group g
                const uint64_t *t2[3] = {tx2, ty2, tz2};
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    float a0, a1, a2, a3;
                    unpack2(t2[d][0], a0, a1);
                    unpack2(t2[d][1], a2, a3);
                    float4 *cell = reinterpret_cast<float4 *>(&s_c[d * U_BLOCK + g * U_GROUP + 4 * lane]);
                    float4 v = *cell;
                    v.x += a0; v.y += a1; v.z += a2; v.w += a3;
                    *cell = v;
                }
            }
            __syncthreads();   // the next round moves every warp to another column group
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // fold this block's FP32 row sums into the FP64 accumulators
            float u, v;
            unpack2(rx2[k], u, v); ra[k][0] += (double)u + (double)v;
            unpack2(ry2[k], u, v); ra[k][1] += (double)u + (double)v;
            unpack2(rz2[k], u, v); ra[k][2] += (double)u + (double)v;
        }
        if (!diag) {
            for (int q = tid; q < 3 * U_BLOCK; q += U_THREADS) {
                const int d = q / U_BLOCK, o = q - d * U_BLOCK;
                colrow[d * cap + (size_t)J * U_BLOCK + o] = s_c[q];
            }
        }
        __syncthreads();   // column accumulators and buffer b are free again
    }
    double *part = sc.partial + (size_t)blockIdx.y * 3 * cap;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = I * U_BLOCK + w * U_GROUP + k * 32 + lane;
#pragma unroll
        for (int d = 0; d < 3; ++d) part[d * cap + r] = ra[k][d];
    }
}

__global__ void __launch_bounds__(256)
k3_sym32_reduce(cosmo3_state st, cosmo3_scratch sc, int shard_index, int shard_count, int chunk, double unscale,
                double G, double dt, int store_acc, int defer, int near) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const size_t cap = (size_t)st.cap;
    const int nblk = (n + U_BLOCK - 1) / U_BLOCK;
    const int K = k / U_BLOCK;
    double a[3] = {0.0, 0.0, 0.0};
    if (K % shard_count == shard_index) {
        const int nchunk = (nblk - K + chunk - 1) / chunk;
        for (int c = 0; c < nchunk; ++c)
#pragma unroll
            for (int d = 0; d < 3; ++d) a[d] += sc.partial[((size_t)c * 3 + d) * cap + k];
    }
    const float *colbuf = reinterpret_cast<const float *>(sc.colbuf);
    int r = 0;
    for (int I = shard_index; I < K; I += shard_count, ++r)
#pragma unroll
        for (int d = 0; d < 3; ++d) a[d] += (double)__ldcs(&colbuf[((size_t)r * 3 + d) * cap + k]);
#pragma unroll
    for (int d = 0; d < 3; ++d) a[d] *= unscale;
    if (near) {   // FP64 near-field corrections of the rows this rank evaluated (nearfield.cu; zero elsewhere)
#pragma unroll
        for (int d = 0; d < 3; ++d) a[d] += sc.corr[d * cap + k];
    }
    if (defer) {
#pragma unroll
        for (int d = 0; d < 3; ++d) sc.acc[d * cap + k] = a[d];
    } else {
        integrate_store3(st, sc, k, a[0], a[1], a[2], G, dt, store_acc != 0);
    }
}

int launch_accel3_sym32(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                        cudaStream_t s) {
    const int n = p.n_upper;
    if (n <= 0) return COSMO_OK;
    const int nblk = (n + U_BLOCK - 1) / U_BLOCK;
    const int sh_n = p.shard_count > 0 ? p.shard_count : 1;
    const int sh_i = p.shard_count > 0 ? p.shard_index : 0;
    const int rows_local = (nblk - sh_i + sh_n - 1) / sh_n;
    int chunk = p.sym_chunk > 0 ? p.sym_chunk : 4;
    while ((nblk + chunk - 1) / chunk > sc.jsplit_max) ++chunk;
    // colbuf rows are counted in units of three planes of `cap` DOUBLES; a float row takes half of one
    if (!sc.colbuf || (long long)(rows_local + 1) / 2 > sc.colbuf_rows) {
        set_error("symmetric 3-D FP32 accel: colbuf has %lld rows, %d needed", (long long)sc.colbuf_rows,
                  (rows_local + 1) / 2);
        return COSMO_E_ARG;
    }
    COSMO_TRY(configure_sym_kernels());
    if (rows_local > 0) {
        dim3 grid(rows_local, (nblk + chunk - 1) / chunk);
        COSMO_TRY(launch_sym3(kTuned32, (const void *)k3_accel_f32_sym, "k3_accel_f32_sym", grid, U_THREADS, U_SMEM, st, sc,
                              sh_i, sh_n, chunk, s));
    }
    const double unscale = scalbn(1.0, p.mass_exp - 2 * p.pos_exp);
    k3_sym32_reduce<<<(n + 255) / 256, 256, 0, s>>>(st, sc, sh_i, sh_n, chunk, unscale, p.G, p.dt,
                                                    (p.opts & COSMO3_OPT_STORE_ACC) ? 1 : 0,
                                                    (p.opts & COSMO3_OPT_DEFER_INTEGRATE) ? 1 : 0, p.near_mode != 0);
    return COSMO_LAUNCH_CHECK("k3_sym32_reduce");
}

// Opt-in to > 48 KB of dynamic shared memory for both symmetric kernels, once per process (one process
// drives one GPU).  Called from every host entry point of the 3-D step, so it has always happened
// before a step is captured into a CUDA graph.
int configure_sym_kernels() {
    static bool configured = false;
    if (configured) return COSMO_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {   // no device (symbol / argument tests on a CPU box)
        cudaGetLastError();
        return COSMO_OK;
    }
    COSMO_TRY(check_cuda(cudaFuncSetAttribute(k3_accel_f64_sym, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)T_SMEM), "cudaFuncSetAttribute(k3_accel_f64_sym)"));
    COSMO_TRY(check_cuda(cudaFuncSetAttribute(k3_accel_f32_sym, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)U_SMEM), "cudaFuncSetAttribute(k3_accel_f32_sym)"));
    // the post-scheduled copies (sym_tuned.cu) are loaded and opted in here too, never inside a capture
    tuned_kernel(kTuned64, (int)T_SMEM);
    tuned_kernel(kTuned32, (int)U_SMEM);
    configured = true;
    return COSMO_OK;
}

}  // namespace s3
}  // namespace cosmo
