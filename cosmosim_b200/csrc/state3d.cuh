// Device helpers shared by the translation units of the 3-D step (state3d.cu, state3d_sym.cu).
#pragma once
#include "cosmo_device.cuh"
#include "cosmo_internal.h"
#include "cosmosim_b200_state3d.h"

namespace cosmo {
namespace s3 {

// c(i,j) of the reference's d^2 chain with three dimensions (blas.py:9-23):
//   R_k = fl(fl(q_ik p_jk) - 2), q = -2 p_i;  c = fl(fl(fl(R_x + R_y) + R_z) + 6)
__device__ __forceinline__ double chain_c(double qx, double qy, double qz, double x, double y, double z) {
    double rx = __dadd_rn(__dmul_rn(qx, x), -2.0);
    double ry = __dadd_rn(__dmul_rn(qy, y), -2.0);
    double rz = __dadd_rn(__dmul_rn(qz, z), -2.0);
    return __dadd_rn(__dadd_rn(__dadd_rn(rx, ry), rz), 6.0);
}

// a = G*S, v = v0 + a dt, p = p0 + v dt (two roundings per update like numpy, universe.py:103-104),
// squared row norm of p as sklearn's row_norms forms it for three columns: fl(fl(x^2 + z^2) + y^2)
__device__ __forceinline__ void integrate_store3(const cosmo3_state &st, const cosmo3_scratch &sc, int i,
                                                 double sx, double sy, double sz, double G, double dt,
                                                 bool store_acc) {
    const size_t cap = (size_t)st.cap;
    const double s[3] = {sx, sy, sz};
    double p[3];
    bool finite = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double a = __dmul_rn(G, s[d]);
        const double v = __dadd_rn(st.vel[d * cap + i], __dmul_rn(a, dt));
        p[d] = __dadd_rn(st.pos[d * cap + i], __dmul_rn(v, dt));
        sc.vel_new[d * cap + i] = v;
        sc.pos_new[d * cap + i] = p[d];
        if (store_acc) sc.acc[d * cap + i] = a;
        finite = finite && isfinite(a);
    }
    sc.xx_new[i] = __dadd_rn(__dadd_rn(__dmul_rn(p[0], p[0]), __dmul_rn(p[2], p[2])), __dmul_rn(p[1], p[1]));
    if (!finite)
        atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                 (unsigned long long)COSMO_ERR_NAN);
}

// symmetric flavours (state3d_sym.cu)
int configure_sym_kernels();
int launch_integrate3(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p, cudaStream_t s);
int launch_accel3_sym32(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                        cudaStream_t s);
int launch_accel3_sym64(const cosmo3_state &st, const cosmo3_scratch &sc, const cosmo3_step_params &p,
                        cudaStream_t s);

}  // namespace s3
}  // namespace cosmo
