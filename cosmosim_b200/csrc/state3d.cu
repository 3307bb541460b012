// The step of the reference's NEW (3-D, batch) API: State.interact(), cosmosim/core/universe.py:94-129
// (SURVEY.md section 8f rank 3).  Same three stages as the 2-D path, self-contained here:
//
//   A  k3_prep + k3_accel_f64 / k3_accel_f32     acc_blas on (N,3) positions (blas.py:4-41, call site
//                                                universe.py:101) + the integrator (universe.py:103-104)
//   B  k3_detect                                 pairwise_distances(p) <= r_i + r_j (universe.py:107-108)
//   C  k3_sort_pairs, k3_resolve, k3_commit,     the ordered absorb loop, Object.absorb/destroy
//      compaction                                (universe.py:111-128, :39-54)
//
// Stage A/B use the tiling of accel.cu / detect.cu: a CTA keeps BLOCK*IB target objects in registers
// and streams the sources through shared memory with the TMA engine's 1-D bulk copy into a 2-deep
// mbarrier ring; vectors live in HBM as three planes (unit-stride doubles), so a tile is five bulk
// copies.  Source splits meet in `partial` and the last CTA to arrive reduces them in split order.
//
// FP64 flavour: the reference's rounding sequence of d^2, now with three dimensions,
//     R_k = fl(fl(q_ik p_jk) - 2), c = fl(fl(fl(R_x + R_y) + R_z) + 6), d2 = fl(fl(c - h_hi) - h_lo)
// 25 FP64-pipe instructions per interaction + 1 MUFU.RSQ64H.  FP32 flavour: direct differences on
// packed f32x2, 12 packed FMA-pipe instructions + 2 MUFU.RSQ per two interactions.
#include <math.h>

#include "cosmo_device.cuh"
#include "cosmo_internal.h"
#include "state3d.cuh"

namespace cosmo {
namespace s3 {

#ifdef __CUDA_ARCH__
#define X_MUL(a, b) __dmul_rn((a), (b))
#define X_ADD(a, b) __dadd_rn((a), (b))
#define X_DIV(a, b) __ddiv_rn((a), (b))
#else
#define X_MUL(a, b) ((a) * (b))
#define X_ADD(a, b) ((a) + (b))
#define X_DIV(a, b) ((a) / (b))
#endif

// ---------------------------------------------------------------------------------------
// Python numbers (int up to 128 bits | float) with Python's arithmetic, for the absorb loop
// ---------------------------------------------------------------------------------------
struct PyNum {
    bool is_int;
    u128 i;
    double f;
};

__host__ __device__ inline PyNum pn_int(u128 v) { PyNum r; r.is_int = true; r.i = v; r.f = 0.0; return r; }
__host__ __device__ inline PyNum pn_float(double v) { PyNum r; r.is_int = false; r.i = 0; r.f = v; return r; }
__host__ __device__ inline double pn_f(const PyNum &a) { return a.is_int ? u128_to_double(a.i) : a.f; }

__host__ __device__ inline int bitlen(u128 v) {
    uint64_t hi = (uint64_t)(v >> 64), lo = (uint64_t)v;
#ifdef __CUDA_ARCH__
    if (hi) return 128 - __clzll((long long)hi);
    return lo ? 64 - __clzll((long long)lo) : 0;
#else
    if (hi) return 128 - __builtin_clzll(hi);
    return lo ? 64 - __builtin_clzll(lo) : 0;
#endif
}

__host__ __device__ inline PyNum pn_add(const PyNum &a, const PyNum &b, bool *ovf) {
    if (a.is_int && b.is_int) {
        u128 s = a.i + b.i;
        if (s < a.i) *ovf = true;
        return pn_int(s);
    }
    return pn_float(X_ADD(pn_f(a), pn_f(b)));
}

__host__ __device__ inline PyNum pn_mul(const PyNum &a, const PyNum &b, bool *ovf) {
    if (a.is_int && b.is_int) {
        if (a.i != 0 && b.i != 0 && bitlen(a.i) + bitlen(b.i) > 128) *ovf = true;
        return pn_int(a.i * b.i);
    }
    return pn_float(X_MUL(pn_f(a), pn_f(b)));
}

// 2^e as a double, e in the normal range
__host__ __device__ inline double pow2i(int e) {
    uint64_t bits = (uint64_t)(e + 1023) << 52;
    double d;
    memcpy(&d, &bits, 8);
    return d;
}

// int / int, correctly rounded (CPython's long true division), b > 0: restoring division that
// yields the leading 55-56 quotient bits plus a sticky bit, then one round-half-even to 53 bits
__host__ __device__ inline double u128_truediv(u128 a, u128 b) {
    if (a == 0) return 0.0;
    const int la = bitlen(a), lb = bitlen(b);
    const int s = 55 - (la - lb);
    const int drop = s < 0 ? -s : 0;
    const int total = la + (s > 0 ? s : 0);
    u128 q = 0, rem = 0;
    for (int k = total - 1; k >= drop; --k) {
        int bit;
        if (s > 0) bit = k >= s ? (int)((a >> (k - s)) & 1) : 0;
        else bit = (int)((a >> k) & 1);
        const bool carry = (rem >> 127) != 0;
        rem = (rem << 1) | (u128)bit;
        q <<= 1;
        if (carry || rem >= b) { rem -= b; q |= 1; }
    }
    bool sticky = rem != 0;
    if (drop && (a & ((((u128)1) << drop) - 1)) != 0) sticky = true;
    const int sh = bitlen(q) - 53;
    uint64_t mant = (uint64_t)(q >> sh);
    const u128 low = q & ((((u128)1) << sh) - 1);
    const u128 half = ((u128)1) << (sh - 1);
    if (low > half || (low == half && (sticky || (mant & 1)))) ++mant;
    // mant <= 2^53; the exponent sh - s lies within [-140, 140] for 128-bit operands
    return X_MUL((double)mant, pow2i(sh - s));
}

__host__ __device__ inline PyNum pn_truediv(const PyNum &a, const PyNum &b) {
    if (a.is_int && b.is_int) return pn_float(u128_truediv(a.i, b.i));
    return pn_float(X_DIV(pn_f(a), pn_f(b)));
}

// Object.get_radius(): ((3*(mass/density))/(4*math.pi))**(1/3)   (universe.py:33-37)
__host__ __device__ inline double radius_of(const PyNum &mass, const PyNum &dens) {
    const double vol = pn_f(pn_truediv(mass, dens));
    const double x = X_DIV(X_MUL(3.0, vol), 12.566370614359172);  // 4*math.pi
    return pow_third(x);
}

__device__ __forceinline__ PyNum load_mass(const cosmo3_state &st, int k) {
    if (st.flags[k] & COSMO3_F_MASS_INT) return pn_int(((u128)st.mass_hi[k] << 64) | st.mass_lo[k]);
    return pn_float(st.mass[k]);
}
__device__ __forceinline__ PyNum load_dens(const cosmo3_state &st, int k) {
    if (st.flags[k] & COSMO3_F_DENS_INT) return pn_int((u128)(uint64_t)st.density[k]);
    return pn_float(st.density[k]);
}

// ---------------------------------------------------------------------------------------
// stage A
// ---------------------------------------------------------------------------------------
__global__ void k3_prep(cosmo3_state st, cosmo3_scratch sc, int pos_exp, int mass_exp, int do_f32) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int n_pad = min(st.cap, (n + COSMO_PAD - 1) / COSMO_PAD * COSMO_PAD);
    const size_t cap = (size_t)st.cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st.status[COSMO_ST_N_PAIRS] = 0;
        st.status[COSMO_ST_N_DEAD] = 0;
        st.status[COSMO_ST_N_CAND] = 0;
        st.status[COSMO_ST_N_GIANTS] = 0;
        st.status[COSMO_ST_PMAX2] = 0;
        st.status[COSMO_ST_N_HEAVY] = 0;
    }
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_pad; k += gridDim.x * blockDim.x) {
        if (k < n) {
            const double x = st.pos[k], y = st.pos[cap + k], z = st.pos[2 * cap + k];
            sc.h[k] = 0.5 * chain_c(-2.0 * x, -2.0 * y, -2.0 * z, x, y, z);
            sc.rinf[k] = __dmul_rn(st.radius[k], 1.0 + 9.313225746154785e-10);  // 1 + 2^-30
            sc.exists[k] = 1;
            if (do_f32) {
                sc.pk[k] = (float)scalbn(x, -pos_exp);
                sc.pk[cap + k] = (float)scalbn(y, -pos_exp);
                sc.pk[2 * cap + k] = (float)scalbn(z, -pos_exp);
                sc.pk[3 * cap + k] = (float)scalbn(st.mass[k], -mass_exp);
            }
        } else {
            // padding slots: massless objects far outside the scene (benign sources and targets)
            const double far = 1.0e100;
            for (int d = 0; d < 3; ++d) { st.pos[d * cap + k] = far; st.vel[d * cap + k] = 0.0; }
            st.mass[k] = 0.0;
            sc.h[k] = 0.5 * chain_c(-2.0 * far, -2.0 * far, -2.0 * far, far, far, far);
            sc.rinf[k] = 0.0;
            sc.exists[k] = 0;
            if (do_f32)
                for (int d = 0; d < 4; ++d) sc.pk[d * cap + k] = 0.f;
        }
    }
}

template <int BLOCK, int IB>
__device__ __forceinline__ void reduce_and_integrate3(const cosmo3_state &st, const cosmo3_scratch &sc, int base,
                                                      int i_hi, const double (&sx)[IB], const double (&sy)[IB],
                                                      const double (&sz)[IB], int jsplit, double G, double dt,
                                                      bool store_acc) {
    __shared__ int s_last;
    const int tid = threadIdx.x;
    if (jsplit == 1) {
#pragma unroll
        for (int k = 0; k < IB; ++k) {
            int i = base + k * BLOCK + tid;
            if (i < i_hi) integrate_store3(st, sc, i, sx[k], sy[k], sz[k], G, dt, store_acc);
        }
        return;
    }
    const size_t cap = (size_t)st.cap;
    double *part = sc.partial + (size_t)blockIdx.y * 3 * cap;
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        if (i < i_hi) {
            __stcg(&part[i], sx[k]);
            __stcg(&part[cap + i], sy[k]);
            __stcg(&part[2 * cap + i], sz[k]);
        }
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
            double a[3] = {0.0, 0.0, 0.0};
            for (int s = 0; s < jsplit; ++s)
#pragma unroll
                for (int d = 0; d < 3; ++d) a[d] += __ldcg(&sc.partial[((size_t)s * 3 + d) * cap + i]);
            integrate_store3(st, sc, i, a[0], a[1], a[2], G, dt, store_acc);
        }
    }
    if (tid == 0) sc.tile_ctr[blockIdx.x] = 0;  // re-armed for the next step
}

// One interaction i <- j (see oracle/state3d_oracle.c for the same chain on the CPU).  Only the
// d2 chain needs the reference's exact roundings; 1/(d2 sqrt d2) is y^3 with y = rsqrt(d2) refined
// from the MUFU seed by one third-order step.  ORDER / CHECK0 as in accel.cu.
template <int ORDER, bool CHECK0>
__device__ __forceinline__ void pair3_f64(const double (&pi)[3], const double (&qi)[3], double hi, int i,
                                          double xj, double yj, double zj, double hj, double mj, int j,
                                          double (&a)[3]) {
    const double c = chain_c(qi[0], qi[1], qi[2], xj, yj, zj);
    double d2;
    if (ORDER == 0) {
        d2 = __dadd_rn(__dadd_rn(c, -hi), -hj);
    } else if (ORDER == 2) {
        d2 = __dadd_rn(__dadd_rn(c, -hj), -hi);
    } else {
        const bool ge = j >= i;
        const double a1 = ge ? hj : hi;
        const double a2 = ge ? hi : hj;
        d2 = __dadd_rn(__dadd_rn(c, -a1), -a2);
    }
    const double dx = xj - pi[0], dy = yj - pi[1], dz = zj - pi[2];
    double y = rsqrt_approx_f64(d2);
    const double t = d2 * y;
    const double e = __fma_rn(-t, y, 1.0);
    const double p = __fma_rn(e, 0.375, 0.5);
    const double ye = y * e;
    y = __fma_rn(ye, p, y);
    double s = (mj * y) * (y * y);
    if (CHECK0) s = ((__double2hiint(d2) << 1) | __double2loint(d2)) == 0 ? mj : s;  // where= mask, blas.py:31
    a[0] = __fma_rn(dx, s, a[0]);
    a[1] = __fma_rn(dy, s, a[1]);
    a[2] = __fma_rn(dz, s, a[2]);
}

template <int IB, int TJ, int ORDER, bool CHECK0>
__device__ __forceinline__ void tile3_f64(const double (*__restrict__ sp)[TJ], int cnt, int j0,
                                          const double (&pi)[IB][3], const double (&qi)[IB][3],
                                          const double (&hi)[IB], const int (&ii)[IB], double (&a)[IB][3]) {
    // sp[0..2] = x, y, z; sp[3] = h; sp[4] = m
#pragma unroll 2
    for (int jj = 0; jj < cnt; ++jj) {
        const double xj = sp[0][jj], yj = sp[1][jj], zj = sp[2][jj], hj = sp[3][jj], mj = sp[4][jj];
#pragma unroll
        for (int k = 0; k < IB; ++k)
            pair3_f64<ORDER, CHECK0>(pi[k], qi[k], hi[k], ii[k], xj, yj, zj, hj, mj, j0 + jj, a[k]);
    }
}

template <int BLOCK, int IB, int TJ>
__global__ void __launch_bounds__(BLOCK)
k3_accel_f64(cosmo3_state st, cosmo3_scratch sc, int i_begin, int i_end, int jsplit, double G, double dt,
             int store_acc) {
    __shared__ __align__(16) double s_p[2][5][TJ];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int base = i_begin + blockIdx.x * (BLOCK * IB);
    if (base >= i_hi) return;
    const int tid = threadIdx.x;
    const size_t cap = (size_t)st.cap;

    double pi[IB][3], qi[IB][3], hi[IB], a[IB][3];
    int ii[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        ii[k] = i;
        int ic = min(i, i_hi - 1);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            pi[k][d] = st.pos[d * cap + ic];
            qi[k][d] = -2.0 * pi[k][d];
            asm volatile("" : "+d"(qi[k][d]));  // keep q in registers (see accel.cu)
            a[k][d] = 0.0;
        }
        hi[k] = sc.h[ic];
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
    constexpr uint32_t kTileBytes = TJ * 5 * sizeof(double);
    auto issue = [&](int t, int b) {
        const size_t j0 = (size_t)t * TJ;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kTileBytes);
        bulk_g2s(&s_p[b][0][0], st.pos + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][1][0], st.pos + cap + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][2][0], st.pos + 2 * cap + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][3][0], sc.h + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][4][0], st.mass + j0, TJ * sizeof(double), &s_bar[b]);
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
        double l[IB][3];  // tile-local sums (two-level summation)
#pragma unroll
        for (int k = 0; k < IB; ++k) { l[k][0] = 0.0; l[k][1] = 0.0; l[k][2] = 0.0; }
        if (order == 1) {
            tile3_f64<IB, TJ, 1, true>(s_p[b], cnt, j0, pi, qi, hi, ii, l);
        } else {
            if (order == 0) tile3_f64<IB, TJ, 0, false>(s_p[b], cnt, j0, pi, qi, hi, ii, l);
            else tile3_f64<IB, TJ, 2, false>(s_p[b], cnt, j0, pi, qi, hi, ii, l);
            bool bad = false;
#pragma unroll
            for (int k = 0; k < IB; ++k) bad = bad || !(isfinite(l[k][0]) && isfinite(l[k][1]) && isfinite(l[k][2]));
            if (bad) {  // a d2 == 0 (or negative) entry off the diagonal: redo with the mask honoured
#pragma unroll
                for (int k = 0; k < IB; ++k) { l[k][0] = 0.0; l[k][1] = 0.0; l[k][2] = 0.0; }
                tile3_f64<IB, TJ, 1, true>(s_p[b], cnt, j0, pi, qi, hi, ii, l);
            }
        }
#pragma unroll
        for (int k = 0; k < IB; ++k) { a[k][0] += l[k][0]; a[k][1] += l[k][1]; a[k][2] += l[k][2]; }
        __syncthreads();
    }
    double sx[IB], sy[IB], sz[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) { sx[k] = a[k][0]; sy[k] = a[k][1]; sz[k] = a[k][2]; }
    reduce_and_integrate3<BLOCK, IB>(st, sc, base, i_hi, sx, sy, sz, jsplit, G, dt, store_acc != 0);
}

// FP32 fast flavour: r2 = dx^2 + dy^2 + dz^2 + 2^-60 (scaled units) keeps rsqrt finite for the self
// pair and the massless padding without a select in the hot loop.
#define COSMO3_F32_TINY 8.673617379884035e-19f

template <int BLOCK, int IB, int TJ>
__global__ void __launch_bounds__(BLOCK)
k3_accel_f32(cosmo3_state st, cosmo3_scratch sc, int i_begin, int i_end, int jsplit, double G, double dt,
             double unscale, int store_acc, int near) {
    __shared__ __align__(16) float s_p[2][4][TJ];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int base = i_begin + blockIdx.x * (BLOCK * IB);
    if (base >= i_hi) return;
    const int tid = threadIdx.x;
    const size_t cap = (size_t)st.cap;

    uint64_t xi2[IB], yi2[IB], zi2[IB];
    double axd[IB], ayd[IB], azd[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int ic = min(base + k * BLOCK + tid, i_hi - 1);
        const float x = sc.pk[ic], y = sc.pk[cap + ic], z = sc.pk[2 * cap + ic];
        xi2[k] = pack2(x, x); yi2[k] = pack2(y, y); zi2[k] = pack2(z, z);
        axd[k] = 0.0; ayd[k] = 0.0; azd[k] = 0.0;
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
    constexpr uint32_t kTileBytes = TJ * 4 * sizeof(float);
    auto issue = [&](int t, int b) {
        const size_t j0 = (size_t)t * TJ;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kTileBytes);
#pragma unroll
        for (int d = 0; d < 4; ++d) bulk_g2s(&s_p[b][d][0], sc.pk + d * cap + j0, TJ * sizeof(float), &s_bar[b]);
    };
    if (tid == 0 && t0 < t1) issue(t0, 0);
    const uint64_t eps2 = pack2(COSMO3_F32_TINY, COSMO3_F32_TINY);
    for (int t = t0; t < t1; ++t) {
        const int it = t - t0;
        const int b = it & 1;
        if (tid == 0 && t + 1 < t1) issue(t + 1, b ^ 1);
        mbar_wait(&s_bar[b], (it >> 1) & 1);
        uint64_t ax2[IB], ay2[IB], az2[IB];
#pragma unroll
        for (int k = 0; k < IB; ++k) { ax2[k] = 0ull; ay2[k] = 0ull; az2[k] = 0ull; }
        const uint64_t *px = reinterpret_cast<const uint64_t *>(&s_p[b][0][0]);
        const uint64_t *py = reinterpret_cast<const uint64_t *>(&s_p[b][1][0]);
        const uint64_t *pz = reinterpret_cast<const uint64_t *>(&s_p[b][2][0]);
        const uint64_t *pm = reinterpret_cast<const uint64_t *>(&s_p[b][3][0]);
#pragma unroll 4
        for (int jj = 0; jj < TJ / 2; ++jj) {
            const uint64_t xj2 = px[jj], yj2 = py[jj], zj2 = pz[jj], mj2 = pm[jj];
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                uint64_t dx = sub2(xj2, xi2[k]);
                uint64_t dy = sub2(yj2, yi2[k]);
                uint64_t dz = sub2(zj2, zi2[k]);
                uint64_t r2 = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, eps2)));
                float r2a, r2b;
                unpack2(r2, r2a, r2b);
                uint64_t rinv = pack2(rsqrt_approx_f32(r2a), rsqrt_approx_f32(r2b));
                // same order as the symmetric kernel; nearfield.cu replays it to the bit
                uint64_t s = mul2(mj2, mul2(mul2(rinv, rinv), rinv));
                ax2[k] = fma2(dx, s, ax2[k]);
                ay2[k] = fma2(dy, s, ay2[k]);
                az2[k] = fma2(dz, s, az2[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < IB; ++k) {
            float u, v;
            unpack2(ax2[k], u, v); axd[k] += (double)u + (double)v;
            unpack2(ay2[k], u, v); ayd[k] += (double)u + (double)v;
            unpack2(az2[k], u, v); azd[k] += (double)u + (double)v;
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < IB; ++k) { axd[k] *= unscale; ayd[k] *= unscale; azd[k] *= unscale; }
    if (near && blockIdx.y == 0) {   // FP64 near-field corrections (nearfield.cu), once per row
        const size_t cap = (size_t)st.cap;
#pragma unroll
        for (int k = 0; k < IB; ++k) {
            const int i = base + k * BLOCK + threadIdx.x;
            if (i < i_hi) { axd[k] += sc.corr[i]; ayd[k] += sc.corr[cap + i]; azd[k] += sc.corr[2 * cap + i]; }
        }
    }
    reduce_and_integrate3<BLOCK, IB>(st, sc, base, i_hi, axd, ayd, azd, jsplit, G, dt, store_acc != 0);
}

// ---------------------------------------------------------------------------------------
// stage B: the reference's predicate on the NEW positions (sklearn euclidean_distances, (N,3))
//   XX = fl(fl(x^2 + z^2) + y^2), dot = fma(z,z', fma(y,y', fl(x x'))), d2 = fl(fl(-2 dot + XX_i) + XX_j)
//   flag = sqrt(max(d2, 0)) <= fl(r_i + r_j)
// q = -2 p_i is exact, so the fma chain on q yields -2 dot bit for bit.  Candidate filter with radii
// inflated by 2^-30 (superset), exact test for the rare candidates.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_pair3(const cosmo3_state &st, const cosmo3_scratch &sc, int i, int j) {
    unsigned long long idx = atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_PAIRS]), 1ull);
    if ((long long)idx < sc.pair_cap)
        sc.pairs[idx] = ((uint64_t)(uint32_t)i << 32) | (uint32_t)j;
    else
        atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                 (unsigned long long)COSMO_ERR_PAIR_OVERFLOW);
}

template <int BLOCK, int IB, int TJ>
__global__ void __launch_bounds__(BLOCK)
k3_detect(cosmo3_state st, cosmo3_scratch sc, int i_begin, int i_end, int jsplit) {
    __shared__ __align__(16) double s_p[2][5][TJ];   // x, y, z, xx, inflated radius
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int base = i_begin + blockIdx.x * (BLOCK * IB);
    if (base >= i_hi) return;
    const int tid = threadIdx.x;
    const size_t cap = (size_t)st.cap;

    double q[IB][3], xxi[IB], ri[IB];
    int ii[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        ii[k] = i < i_hi ? i : -1;
        int ic = min(i, i_hi - 1);
#pragma unroll
        for (int d = 0; d < 3; ++d) q[k][d] = -2.0 * sc.pos_new[d * cap + ic];
        xxi[k] = sc.xx_new[ic];
        ri[k] = sc.rinf[ic];
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
    constexpr uint32_t kTileBytes = TJ * 5 * sizeof(double);
    auto issue = [&](int t, int b) {
        const size_t j0 = (size_t)t * TJ;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kTileBytes);
        bulk_g2s(&s_p[b][0][0], sc.pos_new + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][1][0], sc.pos_new + cap + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][2][0], sc.pos_new + 2 * cap + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][3][0], sc.xx_new + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_p[b][4][0], sc.rinf + j0, TJ * sizeof(double), &s_bar[b]);
    };
    if (tid == 0 && t0 < t1) issue(t0, 0);
    for (int t = t0; t < t1; ++t) {
        const int it = t - t0;
        const int b = it & 1;
        if (tid == 0 && t + 1 < t1) issue(t + 1, b ^ 1);
        mbar_wait(&s_bar[b], (it >> 1) & 1);
        const int j0 = t * TJ;
        const int cnt = min(TJ, n - j0);
#pragma unroll 4
        for (int jj = 0; jj < cnt; ++jj) {
            const double xj = s_p[b][0][jj], yj = s_p[b][1][jj], zj = s_p[b][2][jj];
            const double xxj = s_p[b][3][jj], rj = s_p[b][4][jj];
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                const double m2dot = __fma_rn(q[k][2], zj, __fma_rn(q[k][1], yj, __dmul_rn(q[k][0], xj)));
                const double d2 = __dadd_rn(__dadd_rn(m2dot, xxi[k]), xxj);
                const double rs = ri[k] + rj;
                if (d2 <= rs * rs) {
                    const int i = ii[k], j = j0 + jj;
                    if (i >= 0 && i != j) {
                        const double d = sqrt(fmax(d2, 0.0));
                        if (d <= __dadd_rn(st.radius[i], st.radius[j])) emit_pair3(st, sc, i, j);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// stage C
// ---------------------------------------------------------------------------------------
// Rank sort of the unique keys i<<32|j into row-major order = the order of the reference's double loop.
__global__ void __launch_bounds__(256) k3_sort_pairs(cosmo3_state st, cosmo3_scratch sc) {
    __shared__ uint64_t s_keys[256];
    long long M = st.status[COSMO_ST_N_PAIRS];
    if (M > sc.pair_cap) M = sc.pair_cap;
    if (M == 0) return;
    const long long span = (long long)gridDim.x * blockDim.x;
    const long long rounds = (M + span - 1) / span;
    for (long long r = 0; r < rounds; ++r) {
        const long long k = r * span + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        const uint64_t key = k < M ? sc.pairs[k] : ~0ull;
        long long rank = 0;
        for (long long b0 = 0; b0 < M; b0 += 256) {
            __syncthreads();
            s_keys[threadIdx.x] = b0 + threadIdx.x < M ? sc.pairs[b0 + threadIdx.x] : ~0ull;
            __syncthreads();
            const int lim = (int)min(256ll, M - b0);
            for (int q = 0; q < lim; ++q) rank += s_keys[q] < key;
        }
        if (k < M) sc.pairs_sorted[rank] = key;
    }
}

// The ordered absorb loop (universe.py:111-125).  One thread: the loop is sequential by definition
// (live masses decide every comparison) and the flagged set is small.  State of an object while the
// loop runs: a row object i holds p_new/v_new from the moment it is visited (cur_* once it has
// absorbed something); an object j > i not yet visited still holds its OLD position; a destroyed
// object holds mass 0.0 and zero vectors -- and is absorbed again by any row that still flags it,
// because the reference tests only `obj.mass >= other.mass`.
// One flagged ordered pair (i, j) of the loop: 0 = nothing happens (row object gone, or lighter than
// the other), 1 = obj.absorb(other) of an object that still existed, 2 = of one destroyed earlier.
__device__ int absorb_pair(const cosmo3_state &st, const cosmo3_scratch &sc, int i, int j, int stamp, bool *ovf) {
    if (!sc.exists[i]) return 0;
    const size_t cap = (size_t)st.cap;
    const PyNum mi = load_mass(st, i), mj = load_mass(st, j);
    if (!mass_ge(mi.is_int, mi.i, mi.f, mj.is_int, mj.i, mj.f)) return 0;
    const bool existed = sc.exists[j] != 0;
    const bool ti = sc.touched[i] == stamp;
    const bool tj = sc.touched[j] == stamp;
    const PyNum mt = pn_add(mi, mj, ovf);
    const double fmi = pn_f(mi), fmj = pn_f(mj), fmt = pn_f(mt);
    for (int d = 0; d < 3; ++d) {
        const size_t oi = d * cap + i, oj = d * cap + j;
        const double pi = ti ? sc.cur_pos[oi] : sc.pos_new[oi];
        const double vi = ti ? sc.cur_vel[oi] : sc.vel_new[oi];
        double pj = 0.0, vj = 0.0;
        if (existed) {
            if (j < i) {
                pj = tj ? sc.cur_pos[oj] : sc.pos_new[oj];
                vj = tj ? sc.cur_vel[oj] : sc.vel_new[oj];
            } else {
                pj = st.pos[oj];
                vj = st.vel[oj];
            }
        }
        sc.cur_pos[oi] = __ddiv_rn(__dadd_rn(__dmul_rn(fmi, pi), __dmul_rn(fmj, pj)), fmt);
        sc.cur_vel[oi] = __ddiv_rn(__dadd_rn(__dmul_rn(fmi, vi), __dmul_rn(fmj, vj)), fmt);
    }
    sc.touched[i] = stamp;
    const PyNum nd = pn_truediv(pn_add(pn_mul(mi, load_dens(st, i), ovf), pn_mul(mj, load_dens(st, j), ovf), ovf), mt);
    st.density[i] = nd.f;
    uint8_t fl = st.flags[i] & ~(COSMO3_F_MASS_INT | COSMO3_F_DENS_INT);
    if (mt.is_int) {
        fl |= COSMO3_F_MASS_INT;
        st.mass_hi[i] = (uint64_t)(mt.i >> 64);
        st.mass_lo[i] = (uint64_t)mt.i;
    } else {
        st.mass_hi[i] = 0;
        st.mass_lo[i] = 0;
    }
    st.mass[i] = fmt;
    st.flags[i] = fl;
    st.radius[i] = radius_of(mt, nd);
    // other.destroy() (universe.py:48-54)
    sc.exists[j] = 0;
    st.mass[j] = 0.0;
    st.mass_hi[j] = 0;
    st.mass_lo[j] = 0;
    st.flags[j] &= ~COSMO3_F_MASS_INT;
    return existed ? 1 : 2;
}

__device__ void resolve_loop(const cosmo3_state &st, const cosmo3_scratch &sc) {
    long long M = st.status[COSMO_ST_N_PAIRS];
    if (M > sc.pair_cap) M = sc.pair_cap;
    if (M == 0) return;
    const int stamp = (int)st.status[COSMO_ST_STEP];
    long long nev = st.status[COSMO_ST_N_EVENTS];
    long long dead = 0;
    bool ovf = false;
    for (long long q = 0; q < M; ++q) {
        const uint64_t key = sc.pairs_sorted[q];
        const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
        const int what = absorb_pair(st, sc, i, j, stamp, &ovf);
        if (!what) continue;
        if (nev < st.event_cap) {
            st.events[4 * nev + 0] = stamp;
            st.events[4 * nev + 1] = st.id[i];
            st.events[4 * nev + 2] = st.id[j];
            st.events[4 * nev + 3] = what == 1 ? 1 : 0;
        } else {
            st.status[COSMO_ST_ERROR] |= COSMO_ERR_EVENT_OVERFLOW;
        }
        ++nev;
        if (what == 1) ++dead;
    }
    st.status[COSMO_ST_N_EVENTS] = nev;
    st.status[COSMO_ST_N_DEAD] = dead;
    if (ovf) st.status[COSMO_ST_ERROR] |= COSMO3_ERR_INT_OVERFLOW;
}

__global__ void k3_resolve(cosmo3_state st, cosmo3_scratch sc) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    resolve_loop(st, sc);
}

// ---------------------------------------------------------------------------------------
// Parallel stage C (resolve_mode 1; crowded scenes with 10^3..10^6 flagged pairs per step).
// The absorb loop only ever touches the two objects of a flagged pair, so the connected components
// of the flagged-pair graph are independent of each other; inside a component the reference's
// visitation order decides everything.  Hence:
//   counting sort of the pairs by row + per-row sort by partner  -> row-major order (= the double loop)
//   lock-free union-find over the pair edges                      -> component root of every pair
//   counting sort of the pair indices by root + rank sort         -> every component's pairs in order
//   one thread per component replays its pairs (absorb_pair)      -> a result code per pair
//   ordered compaction of the result codes                        -> the absorb-call log, in loop order
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long pair_count3(const cosmo3_state &st, const cosmo3_scratch &sc) {
    const long long M = st.status[COSMO_ST_N_PAIRS];
    return M > sc.pair_cap ? sc.pair_cap : M;
}

__device__ __forceinline__ int uf_find3(int *parent, int v) {
    int p = *(volatile int *)&parent[v];
    while (p != v) {
        const int gp = *(volatile int *)&parent[p];
        if (gp != p) parent[v] = gp;     // path halving (benign race: only ever moves towards the root)
        v = p;
        p = gp;
    }
    return v;
}

__global__ void __launch_bounds__(256) k3p_count(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st.status[COSMO_ST_N_DEAD] = 0;
        st.status[COSMO_ST_N_HEAVY] = 0;      // re-used as the long-row counter of k3p_sort_rows
    }
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs[q];
        const int i = (int)(key >> 32), j = (int)(key & 0xffffffffu);
        atomicAdd(&sc.row_count[i], 1);
        sc.label[i] = i;
        sc.label[j] = j;
    }
}

__global__ void __launch_bounds__(256) k3p_fill(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs[q];
        const int i = (int)(key >> 32);
        const int slot = sc.row_start[i] + atomicSub(&sc.row_count[i], 1) - 1;   // counts return to zero
        sc.pairs_sorted[slot] = key;
    }
}

// order every row by partner: short rows by the owner of the row's first slot (insertion sort), long
// rows (a giant flagged with thousands of objects) by a rank sort spread over a whole CTA
constexpr int ROW_LONG = 48;

__global__ void __launch_bounds__(256) k3p_sort_rows(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(sc.pairs_sorted[q] >> 32);
        const int s0 = sc.row_start[i], s1 = sc.row_start[i + 1];
        sc.pair_flag[q] = 0;
        if (q != s0 || s1 - s0 < 2) continue;
        if (s1 - s0 > ROW_LONG) {      // handed to k3p_sort_long_rows (N_HEAVY is free after stage B)
            const unsigned long long r =
                atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_HEAVY]), 1ull);
            sc.comp_items[r] = i;
            continue;
        }
        for (int a = s0 + 1; a < s1; ++a) {
            const uint64_t key = sc.pairs_sorted[a];
            int b = a - 1;
            while (b >= s0 && sc.pairs_sorted[b] > key) {
                sc.pairs_sorted[b + 1] = sc.pairs_sorted[b];
                --b;
            }
            sc.pairs_sorted[b + 1] = key;
        }
    }
}

// long rows: CTA b takes the rows i = b, b + gridDim.x, ...; ranks are counted against the row itself
// and the sorted row is staged in `pairs` (free after k3p_fill) before it is copied back
__global__ void __launch_bounds__(256) k3p_sort_long_rows(cosmo3_state st, cosmo3_scratch sc) {
    const int nlong = (int)st.status[COSMO_ST_N_HEAVY];          // rows listed by k3p_sort_rows
    for (int r = blockIdx.x; r < nlong; r += gridDim.x) {
        const int i = sc.comp_items[r];                          // comp_items is free until k3p_comp_fill
        const int s0 = sc.row_start[i], s1 = sc.row_start[i + 1];
        for (int a = s0 + threadIdx.x; a < s1; a += blockDim.x) {
            const uint64_t key = sc.pairs_sorted[a];
            int rank = 0;
            for (int b = s0; b < s1; ++b) rank += sc.pairs_sorted[b] < key;
            sc.pairs[s0 + rank] = key;
        }
        __syncthreads();
        for (int a = s0 + threadIdx.x; a < s1; a += blockDim.x) sc.pairs_sorted[a] = sc.pairs[a];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k3p_union(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const uint64_t key = sc.pairs_sorted[q];
        int a = uf_find3(sc.label, (int)(key >> 32)), b = uf_find3(sc.label, (int)(key & 0xffffffffu));
        while (a != b) {
            const int hi = max(a, b), lo = min(a, b);
            const int old = atomicCAS(&sc.label[hi], hi, lo);    // hook root `hi` under `lo`
            if (old == hi) break;
            a = uf_find3(sc.label, old);                          // someone else moved it: retry from the new roots
            b = uf_find3(sc.label, lo);
        }
    }
}

// flatten (root = smallest object index of the component) and count the pairs per component
__global__ void __launch_bounds__(256) k3p_comp_count(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const int root = uf_find3(sc.label, (int)(sc.pairs_sorted[q] >> 32));
        sc.comp_root[q] = root;
        atomicAdd(&sc.row_count[root], 1);
    }
}

__global__ void __launch_bounds__(256) k3p_comp_fill(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const int root = sc.comp_root[q];
        const int slot = sc.row_start[root] + atomicSub(&sc.row_count[root], 1) - 1;
        sc.comp_items[slot] = (int)q;
    }
}

// order every component's pairs by sorted-pair index: rank = number of smaller indices in the component
__global__ void __launch_bounds__(256) k3p_comp_rank(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const int root = sc.comp_root[q];
        const int s0 = sc.row_start[root], s1 = sc.row_start[root + 1];
        int rank = 0;
        for (int s = s0; s < s1; ++s) rank += sc.comp_items[s] < (int)q;
        sc.scan_a[s0 + rank] = (int)q;           // scan_a is free until the event compaction
    }
}

// one thread per component root: the component's pairs in the loop's order
__global__ void __launch_bounds__(128) k3p_comp_resolve(cosmo3_state st, cosmo3_scratch sc) {
    if (pair_count3(st, sc) == 0) return;
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int s0 = sc.row_start[v], s1 = sc.row_start[v + 1];
    if (s1 <= s0) return;
    const int stamp = (int)st.status[COSMO_ST_STEP];
    bool ovf = false;
    unsigned long long dead = 0;
    for (int a = s0; a < s1; ++a) {
        const int q = sc.scan_a[a];
        const uint64_t key = sc.pairs_sorted[q];
        const int what = absorb_pair(st, sc, (int)(key >> 32), (int)(key & 0xffffffffu), stamp, &ovf);
        sc.pair_flag[q] = what;
        dead += what == 1;
    }
    if (dead) atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_DEAD]), dead);
    if (ovf) atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                      (unsigned long long)COSMO3_ERR_INT_OVERFLOW);
}

__global__ void __launch_bounds__(256) k3p_mark(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    if (blockIdx.x == 0 && threadIdx.x == 0) st.status[COSMO_ST_EV_BASE] = st.status[COSMO_ST_N_EVENTS];
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x)
        sc.scan_a[q] = sc.pair_flag[q] ? 1 : 0;
}

__global__ void __launch_bounds__(256) k3p_scatter(cosmo3_state st, cosmo3_scratch sc) {
    const long long M = pair_count3(st, sc);
    const int stamp = (int)st.status[COSMO_ST_STEP];
    const long long base = st.status[COSMO_ST_EV_BASE];
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < M; q += (long long)gridDim.x * blockDim.x) {
        const int what = sc.pair_flag[q];
        if (!what) continue;
        const long long e = base + sc.scan_a[q];
        if (e < st.event_cap) {
            const uint64_t key = sc.pairs_sorted[q];
            st.events[4 * e + 0] = stamp;
            st.events[4 * e + 1] = st.id[(int)(key >> 32)];
            st.events[4 * e + 2] = st.id[(int)(key & 0xffffffffu)];
            st.events[4 * e + 3] = what == 1 ? 1 : 0;
        } else {
            atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                     (unsigned long long)COSMO_ERR_EVENT_OVERFLOW);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) st.status[COSMO_ST_N_EVENTS] = base + (M > 0 ? sc.scan_a[M] : 0);
}

static int launch_resolve_parallel(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                                   cudaStream_t s) {
    if (!sc.row_count || !sc.row_start || !sc.label || !sc.comp_root || !sc.comp_items || !sc.pair_flag || !sc.scan_a) {
        set_error("3-D stage C: resolve_mode 1 needs the row / component arrays of cosmo3_scratch");
        return COSMO_E_ARG;
    }
    const int pb = 296;
    const int64_t *n_ptr = &st.status[COSMO_ST_N_ACTIVE];
    k3p_count<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_count"));
    COSMO_TRY(launch_scan_exclusive(sc.row_count, sc.row_start, sc.scan_blk, n_ptr, p.n_upper, 0, s));
    k3p_fill<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_fill"));
    k3p_sort_rows<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_sort_rows"));
    k3p_sort_long_rows<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_sort_long_rows"));
    k3p_union<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_union"));
    k3p_comp_count<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_comp_count"));
    COSMO_TRY(launch_scan_exclusive(sc.row_count, sc.row_start, sc.scan_blk, n_ptr, p.n_upper, 0, s));
    k3p_comp_fill<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_comp_fill"));
    k3p_comp_rank<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_comp_rank"));
    k3p_comp_resolve<<<(p.n_upper + 127) / 128, 128, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_comp_resolve"));
    k3p_mark<<<pb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3p_mark"));
    COSMO_TRY(launch_scan_exclusive(sc.scan_a, sc.scan_a, sc.scan_blk, &st.status[COSMO_ST_N_PAIRS], sc.pair_cap, 0, s));
    k3p_scatter<<<pb, 256, 0, s>>>(st, sc);
    return COSMO_LAUNCH_CHECK("k3p_scatter");
}

// obj.velocity = v[i]; obj.position = p[i] (universe.py:113-114) for every object that still exists,
// or the merged values of an absorber; marks the survivors for the compaction.
__global__ void __launch_bounds__(256) k3_commit(cosmo3_state st, cosmo3_scratch sc) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int stamp = (int)st.status[COSMO_ST_STEP];
    const size_t cap = (size_t)st.cap;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int ex = sc.exists[k];
        sc.scan_in[k] = ex;
        if (!ex) continue;
        const bool t = sc.touched[k] == stamp;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            st.pos[d * cap + k] = t ? sc.cur_pos[d * cap + k] : sc.pos_new[d * cap + k];
            st.vel[d * cap + k] = t ? sc.cur_vel[d * cap + k] : sc.vel_new[d * cap + k];
        }
    }
}

__global__ void __launch_bounds__(256) k3_compact_scatter(cosmo3_state st, cosmo3_scratch sc) {
    if (st.status[COSMO_ST_N_DEAD] == 0) return;
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const size_t cap = (size_t)st.cap;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        if (!sc.exists[k]) continue;
        const int d = sc.scan_out[k];
        for (int c = 0; c < 3; ++c) {
            sc.t_pos[c * cap + d] = st.pos[c * cap + k];
            sc.t_vel[c * cap + d] = st.vel[c * cap + k];
        }
        sc.t_mass[d] = st.mass[k];
        sc.t_mass_hi[d] = st.mass_hi[k];
        sc.t_mass_lo[d] = st.mass_lo[k];
        sc.t_density[d] = st.density[k];
        sc.t_radius[d] = st.radius[k];
        sc.t_id[d] = st.id[k];
        sc.t_flags[d] = st.flags[k];
    }
}

__global__ void __launch_bounds__(256) k3_compact_copyback(cosmo3_state st, cosmo3_scratch sc) {
    const long long dead = st.status[COSMO_ST_N_DEAD];
    if (dead == 0) return;
    const int n_new = (int)(st.status[COSMO_ST_N_ACTIVE] - dead);
    const size_t cap = (size_t)st.cap;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_new; k += gridDim.x * blockDim.x) {
        for (int c = 0; c < 3; ++c) {
            st.pos[c * cap + k] = sc.t_pos[c * cap + k];
            st.vel[c * cap + k] = sc.t_vel[c * cap + k];
        }
        st.mass[k] = sc.t_mass[k];
        st.mass_hi[k] = sc.t_mass_hi[k];
        st.mass_lo[k] = sc.t_mass_lo[k];
        st.density[k] = sc.t_density[k];
        st.radius[k] = sc.t_radius[k];
        st.id[k] = sc.t_id[k];
        st.flags[k] = sc.t_flags[k];
    }
}

__global__ void k3_end_step(cosmo3_state st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        st.status[COSMO_ST_N_ACTIVE] -= st.status[COSMO_ST_N_DEAD];
        st.status[COSMO_ST_STEP] += 1;  // State.iteration += 1 (universe.py:129)
    }
}


// ---------------------------------------------------------------------------------------
// Stage C for small scenes (n_upper <= TAIL_N) in ONE launch of ONE CTA: rank sort of the pairs,
// the sequential absorb loop, commit, stable compaction and the step counter.  At the size of the
// reference's example scenes a step is launch-bound; this replaces nine launches.
// ---------------------------------------------------------------------------------------
constexpr int TAIL_N = 4096;
constexpr int TAIL_THREADS = 1024;
constexpr int TAIL_PER = TAIL_N / TAIL_THREADS;

__global__ void __launch_bounds__(TAIL_THREADS) k3_tail_small(cosmo3_state st, cosmo3_scratch sc, int collisions) {
    __shared__ uint64_t s_keys[TAIL_THREADS];
    __shared__ int s_warp[TAIL_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int stamp = (int)st.status[COSMO_ST_STEP];
    const size_t cap = (size_t)st.cap;
    if (collisions) {
        long long M = st.status[COSMO_ST_N_PAIRS];
        if (M > sc.pair_cap) M = sc.pair_cap;
        // rank sort (keys are unique): row-major order of the reference's double loop
        for (long long k0 = 0; k0 < M; k0 += TAIL_THREADS) {
            const long long k = k0 + tid;
            const uint64_t key = k < M ? sc.pairs[k] : ~0ull;
            long long rank = 0;
            for (long long b0 = 0; b0 < M; b0 += TAIL_THREADS) {
                __syncthreads();
                s_keys[tid] = b0 + tid < M ? sc.pairs[b0 + tid] : ~0ull;
                __syncthreads();
                const int lim = (int)min((long long)TAIL_THREADS, M - b0);
                for (int q = 0; q < lim; ++q) rank += s_keys[q] < key;
            }
            if (k < M) sc.pairs_sorted[rank] = key;
        }
        __threadfence_block();
        __syncthreads();
        if (tid == 0 && M > 0) resolve_loop(st, sc);
        __threadfence_block();
        __syncthreads();
    }
    // commit + gather of the survivors into registers (TAIL_PER consecutive slots per thread)
    double pv[TAIL_PER][6], ms[TAIL_PER], dn[TAIL_PER], rd[TAIL_PER];
    uint64_t mh[TAIL_PER], ml[TAIL_PER];
    int idv[TAIL_PER];
    uint8_t fl[TAIL_PER];
    int live[TAIL_PER], mine = 0;
#pragma unroll
    for (int q = 0; q < TAIL_PER; ++q) {
        const int k = tid * TAIL_PER + q;
        live[q] = (k < n && sc.exists[k]) ? 1 : 0;
        if (live[q]) {
            const bool t = sc.touched[k] == stamp;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                pv[q][d] = t ? sc.cur_pos[d * cap + k] : sc.pos_new[d * cap + k];
                pv[q][3 + d] = t ? sc.cur_vel[d * cap + k] : sc.vel_new[d * cap + k];
            }
            ms[q] = st.mass[k]; mh[q] = st.mass_hi[k]; ml[q] = st.mass_lo[k];
            dn[q] = st.density[k]; rd[q] = st.radius[k]; idv[q] = st.id[k]; fl[q] = st.flags[k];
        }
        mine += live[q];
    }
    // block-wide exclusive scan of `mine`
    const int lane = tid & 31, w = tid >> 5;
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        int v = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        s_warp[lane] = v;
        if (lane == 31) s_total = v;
    }
    __syncthreads();   // every source slot has been read into registers above: in-place scatter is safe
    int dst = incl - mine + (w ? s_warp[w - 1] : 0);
#pragma unroll
    for (int q = 0; q < TAIL_PER; ++q) {
        if (!live[q]) continue;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            st.pos[d * cap + dst] = pv[q][d];
            st.vel[d * cap + dst] = pv[q][3 + d];
        }
        st.mass[dst] = ms[q]; st.mass_hi[dst] = mh[q]; st.mass_lo[dst] = ml[q];
        st.density[dst] = dn[q]; st.radius[dst] = rd[q]; st.id[dst] = idv[q]; st.flags[dst] = fl[q];
        ++dst;
    }
    if (tid == 0) {
        st.status[COSMO_ST_N_DEAD] = n - s_total;
        st.status[COSMO_ST_N_ACTIVE] = s_total;
        st.status[COSMO_ST_STEP] += 1;  // State.iteration += 1 (universe.py:129)
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
constexpr int A64_BLOCK = 128, A64_IB = 2, A64_TJ = 256;
constexpr int A64S_BLOCK = 64, A64S_IB = 1, A64S_TJ = 64;   // small-N shape
constexpr int A32_BLOCK = 256, A32_IB = 4, A32_TJ = 512;
constexpr int DET_BLOCK = 128, DET_IB = 2, DET_TJ = 256;
constexpr int DETS_BLOCK = 64, DETS_IB = 1, DETS_TJ = 64;   // small-N shape
constexpr int SMALL_N = 8192;
static_assert(COSMO_PAD % A64_TJ == 0 && COSMO_PAD % A32_TJ == 0 && COSMO_PAD % DET_TJ == 0, "tiles must divide the padding");

static int check_args3(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p) {
    if (!st || !sc || !p || !st->pos || !st->vel || !st->mass || !st->status || st->cap <= 0 ||
        st->cap % COSMO_PAD != 0 || p->n_upper < 0 || p->n_upper > st->cap || p->i_begin < 0 ||
        p->i_end < p->i_begin || p->i_end > st->cap) {
        set_error("bad argument (cosmo3)");
        return COSMO_E_ARG;
    }
    return COSMO_OK;
}

static int launch_stage_a(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                          cudaStream_t s) {
    int n_pad = (p.n_upper + COSMO_PAD - 1) / COSMO_PAD * COSMO_PAD;
    if (n_pad > st.cap) n_pad = st.cap;
    int blocks = (n_pad + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    const int f32 = (p.opts & COSMO3_OPT_FP32) ? 1 : 0;
    k3_prep<<<blocks, 256, 0, s>>>(st, sc, p.pos_exp, p.mass_exp, f32);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3_prep"));
    if (f32 && p.near_mode != 0) {
        // FP64 near field of the FP32 flavours (nearfield.cu) on shadow structs: positions as planes, the
        // grid arrays of stage B (free until then), corrections into scratch.corr
        cosmo_state s2 = {};
        cosmo_scratch c2 = {};
        s2.cap = st.cap; s2.n_total = st.n_total; s2.status = st.status; s2.pos = st.pos; s2.radius = st.radius;
        c2.rinf = sc.rinf; c2.cell_of = sc.cell_of; c2.cell_count = sc.cell_count; c2.cell_start = sc.cell_start;
        c2.cell_items = sc.cell_items; c2.giants = sc.giants; c2.grid_cells = sc.grid_cells; c2.giant_cap = sc.giant_cap;
        c2.chunk_alive = sc.chunk_bounds; c2.scan_blk = sc.scan_blk; c2.grid_dev = sc.grid_dev; c2.tune_hist = sc.tune_hist;
        cosmo_step_params p2 = {};
        p2.n_upper = p.n_upper; p2.i_begin = p.i_begin; p2.i_end = p.i_end; p2.near_mode = p.near_mode;
        p2.pos_exp = p.pos_exp; p2.mass_exp = p.mass_exp;
        p2.shard_index = p.shard_count > 0 ? p.shard_index : 0;
        p2.shard_count = p.shard_count > 0 ? p.shard_count : 1;
        COSMO_TRY(launch_nearfield3(s2, c2, p2, sc.pk, st.mass, sc.corr, (p.opts & COSMO3_OPT_F32_SYM) ? 1 : 0, s));
    }
    // the symmetric kernels take their rows from shard_index / shard_count (round-robin row blocks), not
    // from [i_begin, i_end): a rank whose contiguous range is empty still owns row blocks
    if (f32 && (p.opts & COSMO3_OPT_F32_SYM)) return launch_accel3_sym32(st, sc, p, s);
    if (!f32 && (p.opts & COSMO3_OPT_F64_SYM)) return launch_accel3_sym64(st, sc, p, s);
    const int rows = p.i_end - p.i_begin;
    if (rows <= 0) return COSMO_OK;
    int jsplit = p.jsplit < 1 ? 1 : p.jsplit;
    if (jsplit > sc.jsplit_max) jsplit = sc.jsplit_max;
    const int store_acc = (p.opts & COSMO3_OPT_STORE_ACC) ? 1 : 0;
    if (f32) {
        const int per = A32_BLOCK * A32_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        const double unscale = scalbn(1.0, p.mass_exp - 2 * p.pos_exp);
        k3_accel_f32<A32_BLOCK, A32_IB, A32_TJ><<<grid, A32_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit, p.G,
                                                                           p.dt, unscale, store_acc, p.near_mode != 0);
        return COSMO_LAUNCH_CHECK("k3_accel_f32");
    }
    if (p.n_upper <= SMALL_N) {
        const int per = A64S_BLOCK * A64S_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        k3_accel_f64<A64S_BLOCK, A64S_IB, A64S_TJ><<<grid, A64S_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit,
                                                                               p.G, p.dt, store_acc);
        return COSMO_LAUNCH_CHECK("k3_accel_f64(small)");
    }
    const int per = A64_BLOCK * A64_IB;
    dim3 grid((rows + per - 1) / per, jsplit);
    k3_accel_f64<A64_BLOCK, A64_IB, A64_TJ><<<grid, A64_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit, p.G, p.dt,
                                                                       store_acc);
    return COSMO_LAUNCH_CHECK("k3_accel_f64");
}

static int launch_stage_b(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                          cudaStream_t s) {
    if (!(p.opts & COSMO3_OPT_COLLISIONS)) return COSMO_OK;
    const int rows = p.i_end - p.i_begin;
    if (p.detect_mode != 0) {
        // the 2-D grid kernels of detect.cu on shadow structs: planes for pos_new, 3-D exact predicate
        cosmo_state s2 = {};
        cosmo_scratch c2 = {};
        s2.cap = st.cap; s2.n_total = st.n_total; s2.status = st.status; s2.radius = st.radius;
        c2.pos_new = sc.pos_new; c2.xx_new = sc.xx_new; c2.rinf = sc.rinf;
        c2.pairs = sc.pairs; c2.pair_cap = sc.pair_cap;
        c2.cell_of = sc.cell_of; c2.cell_count = sc.cell_count; c2.cell_start = sc.cell_start;
        c2.cell_items = sc.cell_items; c2.giants = sc.giants; c2.grid_cells = sc.grid_cells;
        c2.giant_cap = sc.giant_cap; c2.chunk_alive = sc.chunk_bounds; c2.label = sc.heavy_rows;
        c2.scan_blk = sc.scan_blk; c2.grid_dev = sc.grid_dev; c2.tune_hist = sc.tune_hist;
        cosmo_step_params p2 = {};
        p2.opts = COSMO_OPT_COLLISIONS; p2.n_upper = p.n_upper; p2.i_begin = p.i_begin; p2.i_end = p.i_end;
        p2.detect_mode = p.detect_mode; p2.cells_per_side = p.cells_per_side;
        p2.cell_size = p.cell_size; p2.grid_origin_x = p.grid_origin_x; p2.grid_origin_y = p.grid_origin_y;
        p2.giant_radius = p.giant_radius;
        p2.shard_index = p.shard_count > 0 ? p.shard_index : 0;
        p2.shard_count = p.shard_count > 0 ? p.shard_count : 1;
        return launch_detect_grid(s2, c2, p2, s, true);
    }
    if (rows <= 0) return COSMO_OK;
    int jsplit = p.jsplit_detect < 1 ? 1 : p.jsplit_detect;
    if (p.n_upper <= SMALL_N) {   // many short CTAs: at this size the pass is latency-bound
        const int per = DETS_BLOCK * DETS_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        k3_detect<DETS_BLOCK, DETS_IB, DETS_TJ><<<grid, DETS_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit);
        return COSMO_LAUNCH_CHECK("k3_detect(small)");
    }
    const int per = DET_BLOCK * DET_IB;
    dim3 grid((rows + per - 1) / per, jsplit);
    k3_detect<DET_BLOCK, DET_IB, DET_TJ><<<grid, DET_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit);
    return COSMO_LAUNCH_CHECK("k3_detect");
}

static int launch_stage_c(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                          cudaStream_t s) {
    if (p.n_upper <= TAIL_N) {
        k3_tail_small<<<1, TAIL_THREADS, 0, s>>>(st, sc, (p.opts & COSMO3_OPT_COLLISIONS) ? 1 : 0);
        return COSMO_LAUNCH_CHECK("k3_tail_small");
    }
    int blocks = (p.n_upper + 255) / 256;
    blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
    if ((p.opts & COSMO3_OPT_COLLISIONS) && p.resolve_mode == 1) {
        COSMO_TRY(launch_resolve_parallel(st, sc, p, s));
    } else if (p.opts & COSMO3_OPT_COLLISIONS) {
        k3_sort_pairs<<<64, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k3_sort_pairs"));
        k3_resolve<<<1, 32, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k3_resolve"));
    }
    k3_commit<<<blocks, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k3_commit"));
    if (p.opts & COSMO3_OPT_COLLISIONS) {
        COSMO_TRY(launch_scan_exclusive(sc.scan_in, sc.scan_out, sc.scan_blk, &st.status[COSMO_ST_N_ACTIVE],
                                        p.n_upper, 0, s));
        k3_compact_scatter<<<blocks, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k3_compact_scatter"));
        k3_compact_copyback<<<blocks, 256, 0, s>>>(st, sc);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k3_compact_copyback"));
    }
    k3_end_step<<<1, 32, 0, s>>>(st);
    return COSMO_LAUNCH_CHECK("k3_end_step");
}

template <typename K>
static int residency(K kernel, int threads) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}

}  // namespace s3
}  // namespace cosmo

using namespace cosmo;
using namespace cosmo::s3;

extern "C" {

int cosmo3_accel_integrate(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p,
                           void *stream) {
    COSMO_TRY(check_args3(st, sc, p));
    return launch_stage_a(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo3_detect_pairs(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p,
                        void *stream) {
    COSMO_TRY(check_args3(st, sc, p));
    return launch_stage_b(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo3_resolve_commit(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p,
                          void *stream) {
    COSMO_TRY(check_args3(st, sc, p));
    return launch_stage_c(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo3_integrate(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p, void *stream) {
    COSMO_TRY(check_args3(st, sc, p));
    return launch_integrate3(*st, *sc, *p, (cudaStream_t)stream);
}

int cosmo3_seal_pair_row(const cosmo3_state *st, int64_t *row, void *stream) {
    if (!st || !st->status || !row) {
        set_error("bad argument (cosmo3_seal_pair_row)");
        return COSMO_E_ARG;
    }
    return launch_seal_pair_row(st->status, row, (cudaStream_t)stream);
}

int cosmo3_append_pair_rows(const cosmo3_state *st, const cosmo3_scratch *sc, const int64_t *rows, int32_t n_rows,
                            int64_t max_count, void *stream) {
    if (!st || !sc || !rows || n_rows < 0 || max_count < 0) {
        set_error("bad argument (cosmo3_append_pair_rows)");
        return COSMO_E_ARG;
    }
    cosmo_state s2 = {};
    cosmo_scratch c2 = {};
    s2.cap = st->cap; s2.status = st->status;
    c2.pairs = sc->pairs; c2.pair_cap = sc->pair_cap;
    return launch_append_pair_rows(s2, c2, rows, n_rows, (long long)max_count, (cudaStream_t)stream);
}

int cosmo3_step(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p, void *stream) {
    COSMO_TRY(check_args3(st, sc, p));
    cudaStream_t s = (cudaStream_t)stream;
    COSMO_TRY(launch_stage_a(*st, *sc, *p, s));
    COSMO_TRY(launch_stage_b(*st, *sc, *p, s));
    return launch_stage_c(*st, *sc, *p, s);
}

int cosmo3_suggest_jsplit(int32_t rows, int32_t n, int kernel, int32_t jsplit_max) {
    configure_sym_kernels();   // first host call of an engine: never inside a graph capture
    if (rows <= 0 || n <= 0) return 1;
    if (jsplit_max < 1) jsplit_max = 1;
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
    } else {
        cudaGetLastError();
    }
    int per, tile, res;
    if (kernel == 1) {
        per = A32_BLOCK * A32_IB; tile = A32_TJ; res = residency(k3_accel_f32<A32_BLOCK, A32_IB, A32_TJ>, A32_BLOCK);
        if (res <= 0) res = 3;
    } else if (kernel == 2 && n <= SMALL_N) {
        per = DETS_BLOCK * DETS_IB; tile = DETS_TJ; res = 16;
    } else if (kernel == 2) {
        per = DET_BLOCK * DET_IB; tile = DET_TJ; res = residency(k3_detect<DET_BLOCK, DET_IB, DET_TJ>, DET_BLOCK);
        if (res <= 0) res = 6;
    } else if (n <= SMALL_N) {
        per = A64S_BLOCK * A64S_IB; tile = A64S_TJ; res = 16;
    } else {
        per = A64_BLOCK * A64_IB; tile = A64_TJ; res = residency(k3_accel_f64<A64_BLOCK, A64_IB, A64_TJ>, A64_BLOCK);
        if (res <= 0) res = 4;
    }
    const long long slots = (long long)sms * res;
    const long long iblocks = (rows + per - 1) / per;
    const long long tiles = (n + tile - 1) / tile;
    long long best_cost = -1;
    int best = 1;
    for (int js = 1; js <= jsplit_max && js <= tiles; ++js) {
        const long long waves = (iblocks * js + slots - 1) / slots;
        const long long cost = waves * ((tiles + js - 1) / js) * 1024 + js;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = js; }
    }
    return best;
}

double cosmo3_host_truediv(uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo) {
    return u128_truediv(((u128)a_hi << 64) | a_lo, ((u128)b_hi << 64) | b_lo);
}

double cosmo3_host_radius(int mass_int, uint64_t mass_hi, uint64_t mass_lo, double mass_f, int dens_int,
                          double dens_f) {
    const PyNum m = mass_int ? pn_int(((u128)mass_hi << 64) | mass_lo) : pn_float(mass_f);
    const PyNum d = dens_int ? pn_int((u128)(uint64_t)dens_f) : pn_float(dens_f);
    return radius_of(m, d);
}

}  // extern "C"
