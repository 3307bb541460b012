// Optional diagnostics: potential / kinetic energy per planet, total energy, angular momentum.
// Replaces reference cosmosim/core/universe_old.py:181-189,210 (the HUD numbers and
// Planet.energy, which decides bound / unbound).  Evaluated like the reference on the NEW
// positions and velocities with the masses as they are BEFORE the frame's merges.
#include <math.h>

#include "pred.cuh"

namespace cosmo {

constexpr int EN_BLOCK = 64, EN_TJ = 256;

// d_inv = 1 / max(sqrt(max(d2, 0)), 1) = min(rsqrt(max(d2, 0)), 1)   (np.clip(d, 1, None), :180-181)
__device__ __forceinline__ double clipped_inverse_distance(double d2) {
    d2 = fmax(d2, 0.0);
    double y = rsqrt_approx_f64(d2);
    double e = __fma_rn(-(d2 * y), y, 1.0);
    y = __fma_rn(y * e, __fma_rn(e, 0.375, 0.5), y);
    // d2 == 0 (coincident planets): seed is +inf, the refinement yields NaN -> clip gives 1
    return (d2 > 0.0) ? fmin(y, 1.0) : 1.0;
}

__global__ void __launch_bounds__(EN_BLOCK)
k_energy(cosmo_state st, cosmo_scratch sc, double G) {
    __shared__ __align__(16) double2 s_xy[EN_TJ];
    __shared__ __align__(16) double s_xx[EN_TJ];
    __shared__ __align__(16) double s_m[EN_TJ];
    __shared__ double s_red[2][EN_BLOCK / 32];
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i = blockIdx.x * EN_BLOCK + threadIdx.x;
    const bool valid = i < n;
    const int ic = valid ? i : max(n - 1, 0);
    const double2 *pn = reinterpret_cast<const double2 *>(sc.pos_new);
    const double2 p = pn[ic];
    const double qx = -2.0 * p.x, qy = -2.0 * p.y, xxi = sc.xx_new[ic];
    double acc = 0.0;
    for (int j0 = 0; j0 < n; j0 += EN_TJ) {
        __syncthreads();
        for (int q = threadIdx.x; q < EN_TJ; q += EN_BLOCK) {
            const int j = min(j0 + q, n - 1);
            s_xy[q] = pn[j];
            s_xx[q] = sc.xx_new[j];
            s_m[q] = st.mass[j];
        }
        __syncthreads();
        const int cnt = min(EN_TJ, n - j0);
#pragma unroll 4
        for (int jj = 0; jj < cnt; ++jj) {
            const double2 pj = s_xy[jj];
            const double dinv = clipped_inverse_distance(pred_d2(qx, qy, xxi, pj.x, pj.y, s_xx[jj]));
            const double term = s_m[jj] * dinv;
            acc += (j0 + jj == i) ? 0.0 : term;          // np.fill_diagonal(V, 0)
        }
    }
    double e = 0.0, l = 0.0;
    if (valid) {
        const double m = st.mass[i];
        const double2 v = reinterpret_cast<const double2 *>(sc.vel_new)[i];
        const double u = m * (-G * acc);                              // U = m * sum(V, axis=1)
        const double k = 0.5 * m * (v.x * v.x + v.y * v.y);           // K = 0.5 m |v|^2
        e = k + u;
        l = m * (p.x * v.y - p.y * v.x);                              // m * cross(p, v)
        st.energy[i] = e;
    }
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        l += __shfl_xor_sync(0xffffffffu, l, o);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = e; s_red[1][threadIdx.x >> 5] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double es = 0.0, ls = 0.0;
        for (int w = 0; w < EN_BLOCK / 32; ++w) { es += s_red[0][w]; ls += s_red[1][w]; }
        atomicAdd(&st.totals[0], es);
        atomicAdd(&st.totals[1], ls);
    }
}

__global__ void k_energy_reset(cosmo_state st) {
    if (threadIdx.x < 4) st.totals[threadIdx.x] = 0.0;
}

int launch_energies(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                    cudaStream_t s) {
    if (!st.energy || !st.totals) {
        set_error("COSMO_OPT_ENERGY needs state.energy / state.totals");
        return COSMO_E_ARG;
    }
    if (p.n_upper <= 0) return COSMO_OK;
    k_energy_reset<<<1, 32, 0, s>>>(st);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_energy_reset"));
    k_energy<<<(p.n_upper + EN_BLOCK - 1) / EN_BLOCK, EN_BLOCK, 0, s>>>(st, sc, p.G);
    return COSMO_LAUNCH_CHECK("k_energy");
}

}  // namespace cosmo
