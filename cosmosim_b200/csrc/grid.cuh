// Periodic uniform grid shared by the broad phase of stage B (detect.cu) and the FP64 near-field
// pass of the FP32 fast path (nearfield.cu): cell arithmetic and position sources.
#pragma once
#include "cosmo_device.cuh"
#include "cosmo_internal.h"

namespace cosmo {

// Where a grid kernel reads positions from.
//   SRC_NEW2  scratch.pos_new as interleaved double2, scratch.xx_new     (stage B of the 2-D frame)
//   SRC_NEW3  scratch.pos_new as three PLANES double[3][cap] (the 3-D step, state3d.cu): the grid
//             itself stays two-dimensional -- it bins the (x, y) projection, and objects within a
//             flagged 3-D distance of each other are at least as close in projection, so the candidate
//             set is still a superset (for a thin disk the projection costs nothing; for a thick cloud
//             it is only less selective)
//   SRC_OLD2  state.pos, the start-of-frame positions                    (near-field pass of stage A)
//   SRC_OLD3  state.pos as three planes (near-field pass of the 3-D step; binned by (x, y) like SRC_NEW3)
enum { SRC_NEW2 = 0, SRC_NEW3 = 1, SRC_OLD2 = 2, SRC_OLD3 = 3 };

template <int SRC>
__device__ __forceinline__ double2 xy_of(const cosmo_state &st, const cosmo_scratch &sc, int k) {
    if (SRC == SRC_NEW3) return make_double2(sc.pos_new[k], sc.pos_new[(size_t)st.cap + k]);
    if (SRC == SRC_OLD2) return reinterpret_cast<const double2 *>(st.pos)[k];
    if (SRC == SRC_OLD3) return make_double2(st.pos[k], st.pos[(size_t)st.cap + k]);
    return reinterpret_cast<const double2 *>(sc.pos_new)[k];
}

template <int SRC>
__device__ __forceinline__ double xx_of(const cosmo_state &st, const cosmo_scratch &sc, int k) {
    if (SRC == SRC_OLD2 || SRC == SRC_OLD3) {
        const double2 p = xy_of<SRC>(st, sc, k);
        return p.x * p.x + p.y * p.y;
    }
    return sc.xx_new[k];
}

// Cell coordinate along one axis of the PERIODIC grid: the plane is tiled with square cells of edge
// h and the cell (ix, iy) is stored in slot (iy mod c, ix mod c).  Planets within one cell edge of
// each other lie in the same or in adjacent cells, hence in the same or in adjacent slots (mod c):
// all the broad phase needs.  Cells a multiple of c apart share a slot -- that only adds candidates,
// and every candidate goes through the exact predicate -- so the grid has no border: planets flung
// far out of the scene (a violent disk loses a quarter of its planets to a halo within 25 frames)
// land wherever they alias to, a few per slot, instead of piling up in border cells.  The period
// c * h is tuned to hold the bulk of the scene, so bulk cells do not alias each other.
__device__ __forceinline__ int cell_coord(double x, double centre, double h, int c) {
    double t = floor((x - centre) / h + 0.5 * (double)c);
    if (!(fabs(t) < 4.0e15)) t = 0.0;            // NaN / beyond exact integers: such planets are brute-forced (far2)
    long long q = (long long)t % c;
    return (int)(q < 0 ? q + c : q);
}

// Candidate ranges of the cell-sorted item array for a planet in `cell`: its 3 x 3 slot
// neighbourhood (mod c).  The three slots of a grid row are contiguous except across the seam,
// where they split in two: six ranges, the odd ones mostly empty.  Needs c >= 3.
__device__ __forceinline__ long long cand_ranges(const cosmo_scratch &sc, int cell, int c, int (&s0)[6], int (&s1)[6]) {
    const int cx = cell % c, cy = cell / c;
    long long cand = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int yy = cy + d - 1;
        yy = yy < 0 ? yy + c : (yy >= c ? yy - c : yy);
        const int *row = sc.cell_start + (size_t)yy * c;
        int a0 = cx - 1, a1 = cx + 1, b0 = 0, b1 = -1;          // [a0, a1] and, across the seam, [b0, b1]
        if (cx == 0) { a0 = 0; b0 = c - 1; b1 = c - 1; }
        else if (cx == c - 1) { a1 = c - 1; b0 = 0; b1 = 0; }
        s0[2 * d] = row[a0]; s1[2 * d] = row[a1 + 1];
        s0[2 * d + 1] = b1 >= b0 ? row[b0] : 0; s1[2 * d + 1] = b1 >= b0 ? row[b1 + 1] : 0;
        cand += (s1[2 * d] - s0[2 * d]) + (s1[2 * d + 1] - s0[2 * d + 1]);
    }
    return cand;
}

}  // namespace cosmo
