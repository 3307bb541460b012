// Integrator epilogue shared by every stage-A kernel.
#pragma once
#include "cosmo_device.cuh"
#include "cosmo_internal.h"

namespace cosmo {

// ---------------------------------------------------------------------------------------
// epilogue shared by both flavours: a = G*S, v = v0 + a dt, p = p0 + v dt (two roundings per
// update like numpy, universe_old.py:177-178), squared row norm of p for stage B
// (sklearn row_norms: fl(fl(x^2) + fl(y^2))).  Immobile planets are integrated as well --
// the reference uses their drifted rows for the collision test but never stores them
// (universe_old.py:199-201).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void integrate_store(const cosmo_state &st, const cosmo_scratch &sc, int i,
                                                double sx, double sy, double G, double dt,
                                                bool store_acc) {
    const double2 *pos = reinterpret_cast<const double2 *>(st.pos);
    const double2 *vel = reinterpret_cast<const double2 *>(st.vel);
    double ax = __dmul_rn(G, sx), ay = __dmul_rn(G, sy);
    double2 p0 = pos[i], v0 = vel[i];
    double vx = __dadd_rn(v0.x, __dmul_rn(ax, dt));
    double vy = __dadd_rn(v0.y, __dmul_rn(ay, dt));
    double px = __dadd_rn(p0.x, __dmul_rn(vx, dt));
    double py = __dadd_rn(p0.y, __dmul_rn(vy, dt));
    reinterpret_cast<double2 *>(sc.vel_new)[i] = make_double2(vx, vy);
    reinterpret_cast<double2 *>(sc.pos_new)[i] = make_double2(px, py);
    sc.xx_new[i] = __dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py));
    if (store_acc) reinterpret_cast<double2 *>(sc.acc)[i] = make_double2(ax, ay);
    if (!(isfinite(ax) && isfinite(ay)))
        atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                 (unsigned long long)COSMO_ERR_NAN);
}

}  // namespace cosmo
