// Stage A, FP64 parity flavour, SYMMETRIC variant: the reference's d^2 chain is symmetric in
// (i, j) bit for bit (commutative products, hi/lo ordering by index) and its pair term is exactly
// antisymmetric, so every unordered pair is evaluated once: 23 FP64-pipe instructions per pair =
// 11.5 per ordered interaction instead of 20.  Same decomposition as accel_sym.cu (row block in
// registers, column block in shared memory, 8 rounds x 32 rotation steps, column sums travelling
// through the warp by SHFL), with 512-planet blocks and 2 + 2 planets per lane.  Replaces
// acc_blas (cosmosim/util/blas.py:4-41) + integrator (universe_old.py:177-178) like accel.cu.
#include <math.h>

#include "cosmo_device.cuh"
#include "cosmo_internal.h"
#include "integrate.cuh"

namespace cosmo {

constexpr int S64_BLOCK = 512;
constexpr int S64_GROUP = 64;
constexpr int S64_WARPS = 8;
constexpr int S64_THREADS = S64_WARPS * 32;

__device__ __forceinline__ double shfl_f64(double v, int src_lane) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src_lane);
    hi = __shfl_sync(0xffffffffu, hi, src_lane);
    return __hiloint2double(hi, lo);
}

// One unordered pair {i, j}.  DIAG: i and j come from the same block (index order varies, self
// pairs occur, only the row side is accumulated).  CHECK0: honour the where= mask (blas.py:31).
template <bool DIAG, bool CHECK0>
__device__ __forceinline__ void pair_sym_f64(double xi, double yi, double qix, double qiy, double hi,
                                             double nmi, int i, double xj, double yj, double hj, double mj,
                                             int j, double &axi, double &ayi, double &axj, double &ayj) {
    double rx = __dadd_rn(__dmul_rn(qix, xj), -2.0);
    double ry = __dadd_rn(__dmul_rn(qiy, yj), -2.0);
    double c = __dadd_rn(__dadd_rn(rx, ry), 6.0);
    double d2;
    if (!DIAG) {
        d2 = __dadd_rn(__dadd_rn(c, -hj), -hi);        // every column index is above every row index
    } else {
        const bool ge = j >= i;
        d2 = __dadd_rn(__dadd_rn(c, -(ge ? hj : hi)), -(ge ? hi : hj));
    }
    double dx = xj - xi, dy = yj - yi;
    double y = rsqrt_approx_f64(d2);
    double t = d2 * y;
    double e = __fma_rn(-t, y, 1.0);
    double p = __fma_rn(e, 0.375, 0.5);
    double ye = y * e;
    y = __fma_rn(ye, p, y);
    double y3 = (y * y) * y;
    double sj = mj * y3;
    double nsi = nmi * y3;
    if (CHECK0) {
        const bool zero = ((__double2hiint(d2) << 1) | __double2loint(d2)) == 0;
        sj = zero ? mj : sj;
        nsi = zero ? nmi : nsi;
    }
    axi = __fma_rn(dx, sj, axi);
    ayi = __fma_rn(dy, sj, ayi);
    if (!DIAG) {
        axj = __fma_rn(dx, nsi, axj);
        ayj = __fma_rn(dy, nsi, ayj);
    }
}

template <bool DIAG, bool CHECK0>
__device__ __forceinline__ void round_sym_f64(const double2 *__restrict__ gxy, const double *__restrict__ gh,
                                              const double *__restrict__ gm, int lane, int jbase,
                                              const double (&xi)[2], const double (&yi)[2],
                                              const double (&qix)[2], const double (&qiy)[2],
                                              const double (&hi)[2], const double (&nmi)[2], const int (&ii)[2],
                                              double (&rax)[2], double (&ray)[2], double (&tx)[2], double (&ty)[2]) {
#pragma unroll 1
    for (int s = 0; s < 32; ++s) {
        const int c = (lane + s) & 31;
        // chunk c = column planets 2c, 2c+1 of the group
        const double4 XY = *reinterpret_cast<const double4 *>(gxy + 2 * c);
        const double2 H = *reinterpret_cast<const double2 *>(gh + 2 * c);
        const double2 M = *reinterpret_cast<const double2 *>(gm + 2 * c);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            pair_sym_f64<DIAG, CHECK0>(xi[k], yi[k], qix[k], qiy[k], hi[k], nmi[k], ii[k], XY.x, XY.y, H.x, M.x,
                                       jbase + 2 * c, rax[k], ray[k], tx[0], ty[0]);
            pair_sym_f64<DIAG, CHECK0>(xi[k], yi[k], qix[k], qiy[k], hi[k], nmi[k], ii[k], XY.z, XY.w, H.y, M.y,
                                       jbase + 2 * c + 1, rax[k], ray[k], tx[1], ty[1]);
        }
        if (!DIAG) {
            const int src = (lane + 1) & 31;
            tx[0] = shfl_f64(tx[0], src); tx[1] = shfl_f64(tx[1], src);
            ty[0] = shfl_f64(ty[0], src); ty[1] = shfl_f64(ty[1], src);
        }
    }
}

__global__ void __launch_bounds__(S64_THREADS, 2)
k_accel_f64_sym(cosmo_state st, cosmo_scratch sc, int shard_index, int shard_count, int chunk) {
    __shared__ __align__(16) double2 s_xy[2][S64_BLOCK];
    __shared__ __align__(16) double s_h[2][S64_BLOCK];
    __shared__ __align__(16) double s_m[2][S64_BLOCK];
    __shared__ __align__(16) double s_cx[S64_BLOCK];
    __shared__ __align__(16) double s_cy[S64_BLOCK];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int nblk = (n + S64_BLOCK - 1) / S64_BLOCK;
    const int I = shard_index + blockIdx.x * shard_count;
    if (I >= nblk) return;
    const int J0 = I + blockIdx.y * chunk;
    if (J0 >= nblk) return;
    const int J1 = min(nblk, J0 + chunk);
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler too (no BRA.DIV around the SHFLs)
    const double2 *pos = reinterpret_cast<const double2 *>(st.pos);

    double xi[2], yi[2], qix[2], qiy[2], hi[2], nmi[2], rax[2], ray[2];
    int ii[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int r = I * S64_BLOCK + w * S64_GROUP + k * 32 + lane;
        ii[k] = r;
        const double2 p = pos[r];           // slots past n are zero-filled by k_prep (mass 0, h 0)
        xi[k] = p.x; yi[k] = p.y;
        qix[k] = -2.0 * p.x; qiy[k] = -2.0 * p.y;
        asm volatile("" : "+d"(qix[k]), "+d"(qiy[k]));
        hi[k] = sc.h[r];
        nmi[k] = -st.mass[r];
        rax[k] = 0.0; ray[k] = 0.0;
    }
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t kBytes = S64_BLOCK * (sizeof(double2) + 2 * sizeof(double));
    auto issue = [&](int J, int b) {
        const size_t j0 = (size_t)J * S64_BLOCK;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kBytes);
        bulk_g2s(&s_xy[b][0], pos + j0, S64_BLOCK * sizeof(double2), &s_bar[b]);
        bulk_g2s(&s_h[b][0], sc.h + j0, S64_BLOCK * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_m[b][0], st.mass + j0, S64_BLOCK * sizeof(double), &s_bar[b]);
    };
    if (tid == 0) issue(J0, 0);
    double2 *colbuf = reinterpret_cast<double2 *>(sc.colbuf);
    const size_t col_row = (size_t)blockIdx.x * (size_t)st.cap;

    for (int J = J0; J < J1; ++J) {
        const int it_j = J - J0;
        const int b = it_j & 1;
        if (tid == 0 && J + 1 < J1) issue(J + 1, b ^ 1);
        mbar_wait(&s_bar[b], (it_j >> 1) & 1);
        const bool diag = (J == I);
        // block-local row sums, folded into the totals once per column block (two-level summation)
        double bax[2] = {0.0, 0.0}, bay[2] = {0.0, 0.0};
        for (int attempt = 0; attempt < 2; ++attempt) {
            for (int q = tid; q < S64_BLOCK; q += S64_THREADS) { s_cx[q] = 0.0; s_cy[q] = 0.0; }
            __syncthreads();
            for (int it = 0; it < S64_WARPS; ++it) {
                const int g = (w + it) & (S64_WARPS - 1);
                const double2 *gxy = &s_xy[b][g * S64_GROUP];
                const double *gh = &s_h[b][g * S64_GROUP];
                const double *gm = &s_m[b][g * S64_GROUP];
                const int jbase = J * S64_BLOCK + g * S64_GROUP;
                double tx[2] = {0.0, 0.0}, ty[2] = {0.0, 0.0};
                if (diag)
                    round_sym_f64<true, true>(gxy, gh, gm, lane, jbase, xi, yi, qix, qiy, hi, nmi, ii, bax, bay, tx, ty);
                else if (attempt == 0)
                    round_sym_f64<false, false>(gxy, gh, gm, lane, jbase, xi, yi, qix, qiy, hi, nmi, ii, bax, bay, tx, ty);
                else
                    round_sym_f64<false, true>(gxy, gh, gm, lane, jbase, xi, yi, qix, qiy, hi, nmi, ii, bax, bay, tx, ty);
                if (!diag) {   // lane l is home to chunk l: column planets 2l, 2l+1 of group g
                    double2 cx = *reinterpret_cast<double2 *>(&s_cx[g * S64_GROUP + 2 * lane]);
                    double2 cy = *reinterpret_cast<double2 *>(&s_cy[g * S64_GROUP + 2 * lane]);
                    cx.x += tx[0]; cx.y += tx[1];
                    cy.x += ty[0]; cy.y += ty[1];
                    *reinterpret_cast<double2 *>(&s_cx[g * S64_GROUP + 2 * lane]) = cx;
                    *reinterpret_cast<double2 *>(&s_cy[g * S64_GROUP + 2 * lane]) = cy;
                }
                __syncthreads();
            }
            if (diag || attempt == 1) break;
            // off-diagonal fast flavour skips the d2 == 0 mask; a non-finite sum (practically never)
            // means such an entry occurred: redo this block pair with the mask honoured
            bool bad = !(isfinite(bax[0]) && isfinite(bay[0]) && isfinite(bax[1]) && isfinite(bay[1]));
            for (int q = tid; q < S64_BLOCK; q += S64_THREADS) bad = bad || !(isfinite(s_cx[q]) && isfinite(s_cy[q]));
            if (!__syncthreads_or(bad)) break;
            bax[0] = bax[1] = bay[0] = bay[1] = 0.0;
        }
        rax[0] += bax[0]; rax[1] += bax[1];
        ray[0] += bay[0]; ray[1] += bay[1];
        if (!diag) {
            double2 *dst = colbuf + col_row + (size_t)J * S64_BLOCK;
            for (int q = tid; q < S64_BLOCK; q += S64_THREADS) dst[q] = make_double2(s_cx[q], s_cy[q]);
        }
        __syncthreads();
    }
    double2 *partial = reinterpret_cast<double2 *>(sc.partial);
#pragma unroll
    for (int k = 0; k < 2; ++k)
        partial[(size_t)blockIdx.y * st.cap + ii[k]] = make_double2(rax[k], ray[k]);
}

__global__ void __launch_bounds__(256)
k_sym64_reduce(cosmo_state st, cosmo_scratch sc, int shard_index, int shard_count, int chunk, double G,
               double dt, int store_acc, int defer) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int nblk = (n + S64_BLOCK - 1) / S64_BLOCK;
    const int K = k / S64_BLOCK;
    double ax = 0.0, ay = 0.0;
    if (K % shard_count == shard_index) {
        const double2 *partial = reinterpret_cast<const double2 *>(sc.partial);
        const int nchunk = (nblk - K + chunk - 1) / chunk;
        for (int c = 0; c < nchunk; ++c) {
            const double2 v = partial[(size_t)c * st.cap + k];
            ax += v.x;
            ay += v.y;
        }
    }
    const double2 *colbuf = reinterpret_cast<const double2 *>(sc.colbuf);
    int r = 0;
    for (int I = shard_index; I < K; I += shard_count, ++r) {
        const double2 v = __ldcs(&colbuf[(size_t)r * st.cap + k]);
        ax += v.x;
        ay += v.y;
    }
    if (defer == 2) {
        const double2 v = make_double2(ax, ay);
        for (int r = 0; r < sc.n_peers; ++r)
            reinterpret_cast<double2 *>(sc.peer_acc[r])[(size_t)shard_index * st.cap + k] = v;
    } else if (defer) reinterpret_cast<double2 *>(sc.acc)[k] = make_double2(ax, ay);
    else integrate_store(st, sc, k, ax, ay, G, dt, store_acc != 0);
}

int launch_accel_sym64(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                       cudaStream_t s) {
    const int n = p.n_upper;
    if (n <= 0) return COSMO_OK;
    const int nblk = (n + S64_BLOCK - 1) / S64_BLOCK;
    const int sh_n = p.shard_count > 0 ? p.shard_count : 1;
    const int sh_i = p.shard_count > 0 ? p.shard_index : 0;
    const int rows_local = (nblk - sh_i + sh_n - 1) / sh_n;
    int chunk = p.sym_chunk > 0 ? p.sym_chunk : 16;
    while ((nblk + chunk - 1) / chunk > sc.jsplit_max) ++chunk;
    // colbuf rows are counted in float2 units of `cap` (accel_sym.cu); a double2 row takes two
    if (!sc.colbuf || (long long)rows_local * 2 > sc.colbuf_rows) {
        set_error("symmetric FP64 accel: colbuf has %lld float2 rows, %d needed", (long long)sc.colbuf_rows,
                  rows_local * 2);
        return COSMO_E_ARG;
    }
    if (rows_local > 0) {
        dim3 grid(rows_local, (nblk + chunk - 1) / chunk);
        // the post-scheduled copy of the kernel (sym_tuned.cu, sass_tune.py) if the build produced one
        const void *tuned = tuned_kernel("_ZN5cosmo15k_accel_f64_symE11cosmo_state13cosmo_scratchiii", 0);
        if (tuned) {
            cosmo_state st_v = st;
            cosmo_scratch sc_v = sc;
            int a_i = sh_i, a_n = sh_n, a_c = chunk;
            void *args[] = {&st_v, &sc_v, &a_i, &a_n, &a_c};
            COSMO_TRY(check_cuda(cudaLaunchKernel(tuned, grid, dim3(S64_THREADS), args, 0, s), "k_accel_f64_sym (tuned)"));
        } else {
            k_accel_f64_sym<<<grid, S64_THREADS, 0, s>>>(st, sc, sh_i, sh_n, chunk);
            COSMO_TRY(COSMO_LAUNCH_CHECK("k_accel_f64_sym"));
        }
    }
    const int defer = (p.opts & COSMO_OPT_DEFER_INTEGRATE) ? ((p.opts & COSMO_OPT_PEER_EXCHANGE) ? 2 : 1) : 0;
    k_sym64_reduce<<<(n + 255) / 256, 256, 0, s>>>(st, sc, sh_i, sh_n, chunk, p.G, p.dt,
                                                   (p.opts & COSMO_OPT_STORE_ACC) ? 1 : 0, defer);
    return COSMO_LAUNCH_CHECK("k_sym64_reduce");
}

}  // namespace cosmo
