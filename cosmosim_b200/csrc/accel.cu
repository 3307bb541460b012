// Stage A: all-pairs 2-D gravitational acceleration with the semi-implicit Euler update
// fused into the same pass.  Replaces acc_blas (reference cosmosim/util/blas.py:4-41, call
// site cosmosim/core/universe_old.py:174) and the integrator (universe_old.py:177-178).
//
// Shape of the computation (both precisions):
//   grid = (target blocks, source splits).  A CTA owns BLOCK*IB target planets held in
//   registers (IB per thread) and streams its share of the source planets through shared
//   memory in tiles staged by the TMA engine's 1-D bulk copy (cp.async.bulk -> SASS UBLKCP)
//   into a 2-deep mbarrier ring; every lane of a warp reads the same source element
//   (shared-memory broadcast), so the inner loop is pure register math.  Partial sums of the
//   source splits meet in `partial`; the last CTA to arrive for a target block (threadfence
//   + arrival counter) reduces them in split order -- deterministic -- and integrates.
//
// FP64 parity flavour: reproduces the reference's rounding sequence of d^2 exactly (the
// reference forms |p_i|^2 + |p_j|^2 - 2 p_i.p_j through zhpr/dspr2 and loses ~9 digits to
// cancellation at AU scale; a direct dx^2+dy^2 kernel is 1e-9 away from it).  Every product
// and sum of that chain is an explicit __dmul_rn/__dadd_rn so nothing is contracted.
// 20 FP64-pipe instructions per interaction + 1 MUFU.RSQ64H.
//
// FP32 fast flavour: positions/masses rescaled by powers of two (exact) and packed to
// float once per frame; inner loop on packed FFMA2/FMUL2/FADD2 (two sources per
// instruction) -- 9 FMA-pipe instructions per TWO interactions + 2 MUFU.RSQ; per-tile FP32
// sums are folded into FP64 accumulators; the integrator stays FP64.
#include <math.h>

#include "cosmo_device.cuh"
#include "cosmo_internal.h"
#include "integrate.cuh"

namespace cosmo {

// ---------------------------------------------------------------------------------------
// k_prep: per-frame O(N) preparation over ALL active slots (replicated on every rank).
//   h[k]      half diagonal of the reference's d^2 chain (blas.py:23,27)
//   pk_*      FP32 j-stage: x, y, m scaled by exact powers of two
//   rinf      radii inflated for the stage-B candidate filter
//   alive=1, escaped = |pos| > r_max on START-of-frame positions (universe_old.py:75-76,87)
//   thread 0 resets the per-frame counters.
// Slots in [n, n rounded up to COSMO_PAD) are zero-filled so tail tiles read benign data.
// ---------------------------------------------------------------------------------------
__global__ void k_prep(cosmo_state st, cosmo_scratch sc, double r_max, int pos_exp, int mass_exp,
                       int do_f32) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int n_pad = min(st.cap, (n + COSMO_PAD - 1) / COSMO_PAD * COSMO_PAD);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st.status[COSMO_ST_N_PAIRS] = 0;
        st.status[COSMO_ST_N_DEAD] = 0;
        st.status[COSMO_ST_N_CAND] = 0;
        st.status[COSMO_ST_N_GIANTS] = 0;
        st.status[COSMO_ST_PMAX2] = 0;
        st.status[COSMO_ST_N_COMPLEX] = 0;
        st.status[COSMO_ST_N_HEAVY] = 0;
    }
    const double2 *pos = reinterpret_cast<const double2 *>(st.pos);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pad; k += gridDim.x * blockDim.x) {
        if (k < n) {
            double2 p = pos[k];
            // c(k,k) = fl(fl(fl(fl(-2x*x) - 2) + fl(fl(-2y*y) - 2)) + 6)
            double rx = __dadd_rn(__dmul_rn(-2.0 * p.x, p.x), -2.0);
            double ry = __dadd_rn(__dmul_rn(-2.0 * p.y, p.y), -2.0);
            double c = __dadd_rn(__dadd_rn(rx, ry), 6.0);
            sc.h[k] = 0.5 * c;
            double r = st.radius[k];
            sc.rinf[k] = __dmul_rn(r, 1.0 + 9.313225746154785e-10);  // 1 + 2^-30
            double nrm = sqrt(__dadd_rn(__dmul_rn(p.x, p.x), __dmul_rn(p.y, p.y)));
            sc.escaped[k] = nrm > r_max ? 1 : 0;
            sc.alive[k] = 1;
            if (do_f32) {
                sc.pk_x[k] = (float)scalbn(p.x, -pos_exp);
                sc.pk_y[k] = (float)scalbn(p.y, -pos_exp);
                sc.pk_m[k] = (float)scalbn(st.mass[k], -mass_exp);
            }
        } else {
            // padding slots [n, n_pad): massless planets far outside the scene.  The blocked
            // kernels process whole blocks, so these must be benign sources AND targets: zero mass,
            // a finite position with its own consistent chain diagonal, nothing that can cancel
            // to d2 <= 0 against a real planet.
            const double far = 1.0e100;
            reinterpret_cast<double2 *>(st.pos)[k] = make_double2(far, far);
            reinterpret_cast<double2 *>(st.vel)[k] = make_double2(0.0, 0.0);
            st.mass[k] = 0.0;
            double rx = __dadd_rn(__dmul_rn(-2.0 * far, far), -2.0);
            sc.h[k] = 0.5 * __dadd_rn(__dadd_rn(rx, rx), 6.0);
            sc.rinf[k] = 0.0;
            sc.escaped[k] = 0;
            sc.alive[k] = 0;
            if (do_f32) {
                sc.pk_x[k] = 0.f;
                sc.pk_y[k] = 0.f;
                sc.pk_m[k] = 0.f;
            }
        }
    }
}

// Cross-split reduction: every CTA of a target block publishes its partial sums; the last
// one to arrive adds them in split order and integrates.
template <int BLOCK, int IB>
__device__ __forceinline__ void reduce_and_integrate(const cosmo_state &st, const cosmo_scratch &sc,
                                                     int base, int i_hi, const double (&sx)[IB],
                                                     const double (&sy)[IB], int jsplit, double G,
                                                     double dt, bool store_acc) {
    __shared__ int s_last;
    const int tid = threadIdx.x;
    if (jsplit == 1) {
#pragma unroll
        for (int k = 0; k < IB; ++k) {
            int i = base + k * BLOCK + tid;
            if (i < i_hi) integrate_store(st, sc, i, sx[k], sy[k], G, dt, store_acc);
        }
        return;
    }
    double2 *partial = reinterpret_cast<double2 *>(sc.partial);
    const size_t cap = (size_t)st.cap;
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        if (i < i_hi) __stcg(&partial[blockIdx.y * cap + i], make_double2(sx[k], sy[k]));
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        int old = atomicAdd(&sc.tile_ctr[blockIdx.x], 1);
        s_last = (old == jsplit - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        if (i < i_hi) {
            double ax = 0.0, ay = 0.0;
            for (int s = 0; s < jsplit; ++s) {
                double2 v = __ldcg(&partial[s * cap + i]);
                ax += v.x;
                ay += v.y;
            }
            integrate_store(st, sc, i, ax, ay, G, dt, store_acc);
        }
    }
    if (tid == 0) sc.tile_ctr[blockIdx.x] = 0;  // re-armed for the next frame
}

// ---------------------------------------------------------------------------------------
// FP64 parity kernel
// ---------------------------------------------------------------------------------------
// One interaction i <- j following blas.py:9-41 (see oracle/cosmo_oracle.c for the same
// chain on the CPU):
//   R_k = fl(fl(q_ik p_jk) - 2), q = -2 p_i        (zhpr real part)
//   c   = fl(fl(R_x + R_y) + 6)                     (trck.real.sum(0) + 6)
//   d2  = fl(fl(c - h_hi) - h_lo), hi/lo = max/min(i,j)   (dspr2: column axpy, then row)
//   term = (p_j - p_i) * m_j / (d2 sqrt(d2)) if d2 != 0 else (p_j - p_i) * m_j
// Only the d2 chain needs the reference's exact roundings (it is the ill-conditioned part);
// 1/(d2 sqrt d2) is formed as y^3 with y = rsqrt(d2) refined from the MUFU seed by one
// third-order step (relative error ~2e-16), instead of sqrt + divide.
// ORDER: 0 -- every source index j of the tile is below every target i of the CTA
//             (d2 = (c - h_i) - h_j), 2 -- every j is above every i (d2 = (c - h_j) - h_i),
//        1 -- mixed (diagonal tiles): per-pair select on j >= i.
// CHECK0: honour the reference's where= mask for d2 == 0 (self pair, coincident planets).
// Off-diagonal tiles run with CHECK0 = false -- no select / integer test in the hot loop, the
// FP64 pipe's dispatch slots are the bottleneck -- and are re-run with the checked flavour
// in the (practically never taken) case that an accumulator came out non-finite.
template <int ORDER, bool CHECK0>
__device__ __forceinline__ void pair_f64(double xi, double yi, double qix, double qiy, double hi,
                                         int i, double xj, double yj, double hj, double mj, int j,
                                         double &ax, double &ay) {
    double rx = __dadd_rn(__dmul_rn(qix, xj), -2.0);
    double ry = __dadd_rn(__dmul_rn(qiy, yj), -2.0);
    double c = __dadd_rn(__dadd_rn(rx, ry), 6.0);
    double d2;
    if (ORDER == 0) {
        d2 = __dadd_rn(__dadd_rn(c, -hi), -hj);
    } else if (ORDER == 2) {
        d2 = __dadd_rn(__dadd_rn(c, -hj), -hi);
    } else {
        const bool ge = j >= i;
        double a1 = ge ? hj : hi;
        double a2 = ge ? hi : hj;
        d2 = __dadd_rn(__dadd_rn(c, -a1), -a2);
    }
    double dx = xj - xi, dy = yj - yi;
    double y = rsqrt_approx_f64(d2);
    double t = d2 * y;
    double e = __fma_rn(-t, y, 1.0);       // 1 - d2 y^2
    double p = __fma_rn(e, 0.375, 0.5);    // 1/2 + 3e/8
    double ye = y * e;
    y = __fma_rn(ye, p, y);                // y (1 + e/2 + 3 e^2/8)
    double s = (mj * y) * (y * y);
    // where= mask of blas.py:31: entries with d2 == 0 keep the undivided difference
    if (CHECK0) s = ((__double2hiint(d2) << 1) | __double2loint(d2)) == 0 ? mj : s;
    ax = __fma_rn(dx, s, ax);
    ay = __fma_rn(dy, s, ay);
}

template <int IB, int ORDER, bool CHECK0>
__device__ __forceinline__ void tile_f64(const double2 *__restrict__ sxy, const double *__restrict__ sh,
                                         const double *__restrict__ sm, int cnt, int j0,
                                         const double (&xi)[IB], const double (&yi)[IB],
                                         const double (&qix)[IB], const double (&qiy)[IB],
                                         const double (&hi)[IB], const int (&ii)[IB], double (&ax)[IB],
                                         double (&ay)[IB]) {
#pragma unroll 2
    for (int jj = 0; jj < cnt; ++jj) {
        const double2 pj = sxy[jj];
        const double hj = sh[jj];
        const double mj = sm[jj];
#pragma unroll
        for (int k = 0; k < IB; ++k)
            pair_f64<ORDER, CHECK0>(xi[k], yi[k], qix[k], qiy[k], hi[k], ii[k], pj.x, pj.y, hj, mj, j0 + jj,
                                    ax[k], ay[k]);
    }
}

template <int BLOCK, int IB, int TJ>
__global__ void __launch_bounds__(BLOCK)
k_accel_f64(cosmo_state st, cosmo_scratch sc, int i_begin, int i_end, int jsplit, double G, double dt,
            int store_acc) {
    __shared__ __align__(16) double2 s_xy[2][TJ];
    __shared__ __align__(16) double s_h[2][TJ];
    __shared__ __align__(16) double s_m[2][TJ];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int base = i_begin + blockIdx.x * (BLOCK * IB);
    if (base >= i_hi) return;
    const int tid = threadIdx.x;

    const double2 *pos = reinterpret_cast<const double2 *>(st.pos);
    double xi[IB], yi[IB], qix[IB], qiy[IB], hi[IB], ax[IB], ay[IB];
    int ii[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        ii[k] = i;
        int ic = min(i, i_hi - 1);
        double2 p = pos[ic];
        xi[k] = p.x; yi[k] = p.y;
        qix[k] = -2.0 * p.x; qiy[k] = -2.0 * p.y;
        // keep q in registers: ptxas otherwise recomputes -2*p inside the loop on the FP64 pipe
        asm volatile("" : "+d"(qix[k]), "+d"(qiy[k]));
        hi[k] = sc.h[ic];
        ax[k] = 0.0; ay[k] = 0.0;
    }

    const int tiles = (n + TJ - 1) / TJ;
    const int t0 = (int)(((long long)blockIdx.y * tiles) / jsplit);
    const int t1 = (int)(((long long)(blockIdx.y + 1) * tiles) / jsplit);

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t kTileBytes = TJ * (sizeof(double2) + 2 * sizeof(double));
    auto issue = [&](int t, int b) {
        const size_t j0 = (size_t)t * TJ;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kTileBytes);
        bulk_g2s(&s_xy[b][0], pos + j0, TJ * sizeof(double2), &s_bar[b]);
        bulk_g2s(&s_h[b][0], sc.h + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_m[b][0], st.mass + j0, TJ * sizeof(double), &s_bar[b]);
    };
    if (tid == 0 && t0 < t1) issue(t0, 0);

    for (int t = t0; t < t1; ++t) {
        const int it = t - t0;
        const int b = it & 1;
        if (tid == 0 && t + 1 < t1) issue(t + 1, b ^ 1);
        mbar_wait(&s_bar[b], (it >> 1) & 1);
        const int j0 = t * TJ;
        const int cnt = min(TJ, n - j0);
        const int order = (j0 + TJ <= base) ? 0 : (j0 >= base + BLOCK * IB ? 2 : 1);
        // tile-local sums, folded into the running totals once per tile: two-level summation keeps
        // the rounding noise of rows with large cancelling neighbour terms ~8x below a single chain
        double lx[IB], ly[IB];
#pragma unroll
        for (int k = 0; k < IB; ++k) { lx[k] = 0.0; ly[k] = 0.0; }
        if (order == 1) {
            tile_f64<IB, 1, true>(s_xy[b], s_h[b], s_m[b], cnt, j0, xi, yi, qix, qiy, hi, ii, lx, ly);
        } else {
            if (order == 0)
                tile_f64<IB, 0, false>(s_xy[b], s_h[b], s_m[b], cnt, j0, xi, yi, qix, qiy, hi, ii, lx, ly);
            else
                tile_f64<IB, 2, false>(s_xy[b], s_h[b], s_m[b], cnt, j0, xi, yi, qix, qiy, hi, ii, lx, ly);
            bool bad = false;
#pragma unroll
            for (int k = 0; k < IB; ++k) bad = bad || !(isfinite(lx[k]) && isfinite(ly[k]));
            if (bad) {  // a d2 == 0 (or negative) entry off the diagonal: redo with the mask honoured
#pragma unroll
                for (int k = 0; k < IB; ++k) { lx[k] = 0.0; ly[k] = 0.0; }
                tile_f64<IB, 1, true>(s_xy[b], s_h[b], s_m[b], cnt, j0, xi, yi, qix, qiy, hi, ii, lx, ly);
            }
        }
#pragma unroll
        for (int k = 0; k < IB; ++k) { ax[k] += lx[k]; ay[k] += ly[k]; }
        __syncthreads();  // everyone is done with buffer b before it is refilled
    }
    reduce_and_integrate<BLOCK, IB>(st, sc, base, i_hi, ax, ay, jsplit, G, dt, store_acc != 0);
}

// ---------------------------------------------------------------------------------------
// FP32 fast kernel
// ---------------------------------------------------------------------------------------
// r2 is formed as dx*dx + dy*dy + 2^-60 (scaled units, i.e. a softening length of 2^-30 of
// the position scale: ~2 m for the AU-scale scenes against radii >= 1e6 m).  It keeps
// rsqrt finite for the self pair and for zero-mass padding without a compare/select in the
// hot loop; the self term then contributes dx * s = 0 exactly.
#define COSMO_F32_TINY 8.673617379884035e-19f

template <int BLOCK, int IB, int TJ>
__global__ void __launch_bounds__(BLOCK)
k_accel_f32(cosmo_state st, cosmo_scratch sc, int i_begin, int i_end, int jsplit, double G, double dt,
            double unscale, int store_acc, int near) {
    __shared__ __align__(16) float s_x[2][TJ];
    __shared__ __align__(16) float s_y[2][TJ];
    __shared__ __align__(16) float s_m[2][TJ];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int base = i_begin + blockIdx.x * (BLOCK * IB);
    if (base >= i_hi) return;
    const int tid = threadIdx.x;

    float xi[IB], yi[IB];
    double axd[IB], ayd[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int ic = min(base + k * BLOCK + tid, i_hi - 1);
        xi[k] = sc.pk_x[ic];
        yi[k] = sc.pk_y[ic];
        axd[k] = 0.0; ayd[k] = 0.0;
    }

    const int tiles = (n + TJ - 1) / TJ;
    const int t0 = (int)(((long long)blockIdx.y * tiles) / jsplit);
    const int t1 = (int)(((long long)(blockIdx.y + 1) * tiles) / jsplit);

    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t kTileBytes = TJ * 3 * sizeof(float);
    auto issue = [&](int t, int b) {
        const size_t j0 = (size_t)t * TJ;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kTileBytes);
        bulk_g2s(&s_x[b][0], sc.pk_x + j0, TJ * sizeof(float), &s_bar[b]);
        bulk_g2s(&s_y[b][0], sc.pk_y + j0, TJ * sizeof(float), &s_bar[b]);
        bulk_g2s(&s_m[b][0], sc.pk_m + j0, TJ * sizeof(float), &s_bar[b]);
    };
    if (tid == 0 && t0 < t1) issue(t0, 0);

    for (int t = t0; t < t1; ++t) {
        const int it = t - t0;
        const int b = it & 1;
        if (tid == 0 && t + 1 < t1) issue(t + 1, b ^ 1);
        mbar_wait(&s_bar[b], (it >> 1) & 1);
        // Slots past n inside the tile hold zero mass (k_prep), so full tiles are always
        // processed: no tail bound in the hot loop.
        {
            uint64_t xi2[IB], yi2[IB], ax2[IB], ay2[IB];
            const uint64_t eps2 = pack2(COSMO_F32_TINY, COSMO_F32_TINY);
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                xi2[k] = pack2(xi[k], xi[k]);
                yi2[k] = pack2(yi[k], yi[k]);
                ax2[k] = 0ull; ay2[k] = 0ull;
            }
            const uint64_t *px = reinterpret_cast<const uint64_t *>(&s_x[b][0]);
            const uint64_t *py = reinterpret_cast<const uint64_t *>(&s_y[b][0]);
            const uint64_t *pm = reinterpret_cast<const uint64_t *>(&s_m[b][0]);
#pragma unroll 4
            for (int jj = 0; jj < TJ / 2; ++jj) {
                const uint64_t xj2 = px[jj], yj2 = py[jj], mj2 = pm[jj];
#pragma unroll
                for (int k = 0; k < IB; ++k) {
                    uint64_t dx = sub2(xj2, xi2[k]);
                    uint64_t dy = sub2(yj2, yi2[k]);
                    uint64_t r2 = fma2(dy, dy, fma2(dx, dx, eps2));
                    float r2a, r2b;
                    unpack2(r2, r2a, r2b);
                    float ia = rsqrt_approx_f32(r2a);
                    float ib = rsqrt_approx_f32(r2b);
                    uint64_t rinv = pack2(ia, ib);
                    // same order as the symmetric kernel; nearfield.cu replays it to the bit
                    uint64_t s = mul2(mj2, mul2(mul2(rinv, rinv), rinv));
                    ax2[k] = fma2(dx, s, ax2[k]);
                    ay2[k] = fma2(dy, s, ay2[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                float a, c;
                unpack2(ax2[k], a, c);
                axd[k] += (double)a + (double)c;
                unpack2(ay2[k], a, c);
                ayd[k] += (double)a + (double)c;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < IB; ++k) { axd[k] *= unscale; ayd[k] *= unscale; }
    if (near && blockIdx.y == 0) {   // FP64 near-field corrections (nearfield.cu), once per row
#pragma unroll
        for (int k = 0; k < IB; ++k) {
            const int i = base + k * BLOCK + tid;
            if (i < i_hi) {
                const double2 c = reinterpret_cast<const double2 *>(sc.corr)[i];
                axd[k] += c.x; ayd[k] += c.y;
            }
        }
    }
    reduce_and_integrate<BLOCK, IB>(st, sc, base, i_hi, axd, ayd, jsplit, G, dt, store_acc != 0);
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
constexpr int F64_BLOCK = 128, F64_IB = 2, F64_TJ = 256;
constexpr int F32_BLOCK = 256, F32_IB = 4, F32_TJ = 512;
constexpr int F64S_BLOCK = 64, F64S_IB = 1, F64S_TJ = 64;  // small-N shape: many short CTAs
static_assert(COSMO_PAD % F64_TJ == 0 && COSMO_PAD % F32_TJ == 0, "tiles must divide the slot padding");

// resident CTAs per SM of the all-pairs kernels (0: FP64, otherwise FP32)
int accel_residency(int which, int *rows_per_cta, int *tile) {
    int nb = 0;
    cudaError_t e;
    if (which == 0) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_accel_f64<F64_BLOCK, F64_IB, F64_TJ>, F64_BLOCK, 0);
        *rows_per_cta = F64_BLOCK * F64_IB; *tile = F64_TJ;
    } else {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_accel_f32<F32_BLOCK, F32_IB, F32_TJ>, F32_BLOCK, 0);
        *rows_per_cta = F32_BLOCK * F32_IB; *tile = F32_TJ;
    }
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    return nb;
}

int launch_prep(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                cudaStream_t s) {
    int n_pad = (p.n_upper + COSMO_PAD - 1) / COSMO_PAD * COSMO_PAD;
    if (n_pad > st.cap) n_pad = st.cap;
    int blocks = (n_pad + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_prep<<<blocks, 256, 0, s>>>(st, sc, p.r_max, p.pos_exp, p.mass_exp, (p.opts & COSMO_OPT_FP32) ? 1 : 0);
    return COSMO_LAUNCH_CHECK("k_prep");
}

int launch_accel(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                 cudaStream_t s) {
    if ((p.opts & COSMO_OPT_FP32) && p.near_mode != 0) COSMO_TRY(launch_nearfield(st, sc, p, s));
    if ((p.opts & COSMO_OPT_FP32) && (p.opts & COSMO_OPT_F32_SYM)) return launch_accel_sym(st, sc, p, s);
    if (!(p.opts & COSMO_OPT_FP32) && (p.opts & COSMO_OPT_F64_SYM)) return launch_accel_sym64(st, sc, p, s);
    const int rows = p.i_end - p.i_begin;
    if (rows <= 0) return COSMO_OK;
    int jsplit = p.jsplit < 1 ? 1 : p.jsplit;
    if (jsplit > sc.jsplit_max) jsplit = sc.jsplit_max;
    const int store_acc = (p.opts & COSMO_OPT_STORE_ACC) ? 1 : 0;
    if (p.opts & COSMO_OPT_FP32) {
        const int per = F32_BLOCK * F32_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        // S = S_scaled * 2^(mass_exp - 2 pos_exp)
        const double unscale = scalbn(1.0, p.mass_exp - 2 * p.pos_exp);
        k_accel_f32<F32_BLOCK, F32_IB, F32_TJ><<<grid, F32_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit, p.G, p.dt,
                                                                         unscale, store_acc, p.near_mode != 0);
        return COSMO_LAUNCH_CHECK("k_accel_f32");
    }
    if (p.small_tiles) {
        const int per = F64S_BLOCK * F64S_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        k_accel_f64<F64S_BLOCK, F64S_IB, F64S_TJ><<<grid, F64S_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit,
                                                                            p.G, p.dt, store_acc);
        return COSMO_LAUNCH_CHECK("k_accel_f64(small)");
    }
    const int per = F64_BLOCK * F64_IB;
    dim3 grid((rows + per - 1) / per, jsplit);
    k_accel_f64<F64_BLOCK, F64_IB, F64_TJ><<<grid, F64_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit,
                                                                     p.G, p.dt, store_acc);
    return COSMO_LAUNCH_CHECK("k_accel_f64");
}

}  // namespace cosmo
