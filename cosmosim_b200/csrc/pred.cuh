// Collision predicate of the reference and the flagged-pair emitter, shared by the detection
// and resolver translation units.
#pragma once
#include "cosmo_device.cuh"
#include "cosmo_internal.h"

namespace cosmo {

// ---------------------------------------------------------------------------------------
// Exact predicate of the reference on the NEW positions (sklearn euclidean_distances):
//   XX   = fl(fl(x^2) + fl(y^2))                 (xx_new, written by stage A)
//   dot  = fma(y_i, y_j, fl(x_i x_j))            (dgemm, K = 2)
//   d2   = fl(fl(-2 dot + XX_i) + XX_j)          (row norm first => not symmetric in i,j)
//   flag = max(sqrt(max(d2, 0)), 1) <= fl(r_i + r_j)
// q = -2 p_i is exact, so fma(q_y, y_j, fl(q_x x_j)) == -2 dot bit-for-bit.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double pred_d2(double qxi, double qyi, double xxi, double xj, double yj,
                                          double xxj) {
    double m2dot = __fma_rn(qyi, yj, __dmul_rn(qxi, xj));
    return __dadd_rn(__dadd_rn(m2dot, xxi), xxj);
}

__device__ __forceinline__ bool pred_exact(double d2, double ri, double rj) {
    double d = sqrt(fmax(d2, 0.0));  // IEEE sqrt
    d = fmax(d, 1.0);                // np.clip(d, 1, None)
    return d <= __dadd_rn(ri, rj);
}

__device__ __forceinline__ void emit_pair(const cosmo_state &st, const cosmo_scratch &sc, int i, int j) {
    unsigned long long idx =
        atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_PAIRS]), 1ull);
    if ((long long)idx < sc.pair_cap)
        sc.pairs[idx] = ((uint64_t)(uint32_t)i << 32) | (uint32_t)j;
    else
        atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                 (unsigned long long)COSMO_ERR_PAIR_OVERFLOW);
}

}  // namespace cosmo
