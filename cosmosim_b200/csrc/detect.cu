// Stage B: collision-pair detection on the NEW positions.  Replaces pairwise_distances +
// collision_matrix (reference cosmosim/core/universe_old.py:180,192-193).
//
// Two modes, same exact predicate (pred.cuh), same output (unsorted list of flagged ordered
// pairs; stage C orders it):
//   all-pairs   the reference's N x N test, tiled like stage A; used for small N;
//   grid        uniform-grid broad phase: a counting sort of the planets into square cells no
//               smaller than the largest possible flagged separation of two non-giant planets,
//               so every flagged pair of non-giants sits in adjacent cells; "giants" (radius
//               above giant_radius: the star) are tested against everybody.  The candidate
//               set is a superset; every candidate goes through the exact predicate, so the
//               flagged set is identical to the all-pairs one.
#include <math.h>

#include "pred.cuh"

namespace cosmo {

// All-pairs detection: same tiling as stage A (targets in registers, sources through a
// bulk-copy/mbarrier ring).  7 FP64-pipe instructions per ordered pair for the candidate
// filter d2 <= (r'_i + r'_j)^2 with radii inflated by 2^-30 (a superset of the flagged
// pairs); the rare candidates go through the exact predicate with the true radii.
template <int BLOCK, int IB, int TJ>
__global__ void __launch_bounds__(BLOCK)
k_detect_allpairs(cosmo_state st, cosmo_scratch sc, int i_begin, int i_end, int jsplit) {
    __shared__ __align__(16) double2 s_xy[2][TJ];
    __shared__ __align__(16) double s_xx[2][TJ];
    __shared__ __align__(16) double s_r[2][TJ];
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int base = i_begin + blockIdx.x * (BLOCK * IB);
    if (base >= i_hi) return;
    const int tid = threadIdx.x;
    const double2 *pn = reinterpret_cast<const double2 *>(sc.pos_new);

    double qx[IB], qy[IB], xxi[IB], ri[IB];
    int ii[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) {
        int i = base + k * BLOCK + tid;
        ii[k] = i < i_hi ? i : -1;
        int ic = min(i, i_hi - 1);
        double2 p = pn[ic];
        qx[k] = -2.0 * p.x; qy[k] = -2.0 * p.y;
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
    constexpr uint32_t kTileBytes = TJ * (sizeof(double2) + 2 * sizeof(double));
    auto issue = [&](int t, int b) {
        const size_t j0 = (size_t)t * TJ;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kTileBytes);
        bulk_g2s(&s_xy[b][0], pn + j0, TJ * sizeof(double2), &s_bar[b]);
        bulk_g2s(&s_xx[b][0], sc.xx_new + j0, TJ * sizeof(double), &s_bar[b]);
        bulk_g2s(&s_r[b][0], sc.rinf + j0, TJ * sizeof(double), &s_bar[b]);
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
            const double2 pj = s_xy[b][jj];
            const double xxj = s_xx[b][jj];
            const double rj = s_r[b][jj];
#pragma unroll
            for (int k = 0; k < IB; ++k) {
                double d2 = pred_d2(qx[k], qy[k], xxi[k], pj.x, pj.y, xxj);
                double rs = ri[k] + rj;
                if (d2 <= rs * rs) {
                    const int i = ii[k], j = j0 + jj;
                    if (i >= 0 && i != j && pred_exact(d2, st.radius[i], st.radius[j]))
                        emit_pair(st, sc, i, j);
                }
            }
        }
        __syncthreads();
    }
}


// ---------------------------------------------------------------------------------------
// Uniform-grid broad phase
// ---------------------------------------------------------------------------------------
struct GridParams {
    double h, cx, cy, rg;   // cell edge, centre of the linear box, giant radius
    int c, w;               // cells per side (total), logarithmic cells on each side
};

// Cell edge actually used.  The reference compares the COMPUTED distance (expanded form,
// absolute error E in d^2 of a few ulps of |p|^2) with r_i + r_j, so two non-giants (inflated
// radius <= rg) can be flagged while truly sqrt(4 rg^2 + E) apart; the cell edge is never
// smaller than that (plus slack for the rounding of the cell coordinate itself).
__device__ __forceinline__ GridParams grid_params(const cosmo_state &st, double h, double ox, double oy,
                                                  double rg, int c, int w) {
    GridParams g;
    const double pmax2 = __longlong_as_double(st.status[COSMO_ST_PMAX2]);
    const double e = 64.0 * 1.1102230246251565e-16 * pmax2;
    const double hmin = sqrt(4.0 * rg * rg + e) * (1.0 + 1e-6);
    g.h = fmax(h, hmin);
    const double half_in = 0.5 * (double)(c - 2 * w) * h;   // the centre is fixed by the host's (h, origin)
    g.cx = ox + half_in; g.cy = oy + half_in; g.rg = rg; g.c = c; g.w = w;
    return g;
}

// Cell coordinate along one axis.  u = (x - centre) / h is mapped by a monotone function of
// slope <= 1: identity inside the linear box |u| <= U0, U0 + log(1 + (|u| - U0)) outside, clamped
// at the last cell.  Slope <= 1 means planets within one cell edge of each other land in the same
// or in adjacent cells, which is all the broad phase needs; the logarithmic rim spreads planets
// flung far out of the scene over many cells instead of piling them into the corners.
__device__ __forceinline__ int cell_coord(double x, double centre, double h, int c, int w) {
    const double u0 = 0.5 * (double)(c - 2 * w);
    const double u = (x - centre) / h;
    double t;
    if (u > u0) t = u0 + fmin(log1p(u - u0), (double)w - 0.5);
    else if (u < -u0) t = -u0 - fmin(log1p(-u - u0), (double)w - 0.5);
    else t = u;
    t = floor(t + u0 + (double)w);
    t = fmin(fmax(t, 0.0), (double)(c - 1));   // NaN / overflow: clamp
    return (int)t;
}

__global__ void __launch_bounds__(256) k_grid_max(cosmo_state st, cosmo_scratch sc) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    double m = 0.0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        m = fmax(m, sc.xx_new[k]);
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0)  // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_PMAX2]),
                  (unsigned long long)__double_as_longlong(m));
}

// D3 = true serves the 3-D step (state3d.cu): sc.pos_new then holds three PLANES double[3][cap] and
// the exact test is the three-column predicate without the clip of d to 1 (universe.py:107-108).  The
// grid itself stays two-dimensional: it bins the (x, y) projection, and planets within a flagged
// 3-D distance of each other are at least as close in projection, so the candidate set is still a
// superset (for a thin disk the projection costs nothing; for a thick cloud it is only less selective).
template <bool D3>
__device__ __forceinline__ double2 xy_of(const cosmo_state &st, const cosmo_scratch &sc, int k) {
    if (D3) return make_double2(sc.pos_new[k], sc.pos_new[(size_t)st.cap + k]);
    return reinterpret_cast<const double2 *>(sc.pos_new)[k];
}

template <bool D3>
__global__ void __launch_bounds__(256)
k_grid_count(cosmo_state st, cosmo_scratch sc, double h, double ox, double oy, double rg, int c, int w) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const GridParams g = grid_params(st, h, ox, oy, rg, c, w);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        if (sc.rinf[k] > g.rg) {
            unsigned long long idx =
                atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_GIANTS]), 1ull);
            if ((long long)idx < sc.giant_cap)
                sc.giants[idx] = k;
            else
                atomicOr(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_ERROR]),
                         (unsigned long long)COSMO_ERR_GIANT_OVERFLOW);
            sc.cell_of[k] = -1;
        } else {
            const double2 p = xy_of<D3>(st, sc, k);
            const int cell = cell_coord(p.y, g.cy, g.h, g.c, g.w) * g.c + cell_coord(p.x, g.cx, g.h, g.c, g.w);
            sc.cell_of[k] = cell;
            atomicAdd(&sc.cell_count[cell], 1);
        }
    }
}

// counting-sort scatter; counts return to zero, so cell_count needs no per-frame memset
__global__ void __launch_bounds__(256) k_grid_fill(cosmo_state st, cosmo_scratch sc) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int cell = sc.cell_of[k];
        if (cell >= 0) {
            const int slot = sc.cell_start[cell] + atomicSub(&sc.cell_count[cell], 1) - 1;
            sc.cell_items[slot] = k;
        }
    }
}

struct GridRow {
    double qx, qy, qz, xxi, ri;
};

template <bool D3>
__device__ __forceinline__ GridRow load_row(const cosmo_state &st, const cosmo_scratch &sc, int i) {
    GridRow r;
    const double2 p = xy_of<D3>(st, sc, i);
    r.qx = -2.0 * p.x; r.qy = -2.0 * p.y;
    r.qz = D3 ? -2.0 * sc.pos_new[2 * (size_t)st.cap + i] : 0.0;
    r.xxi = sc.xx_new[i]; r.ri = sc.rinf[i];
    return r;
}

template <bool D3>
__device__ __forceinline__ void test_and_emit(const cosmo_state &st, const cosmo_scratch &sc, int i, const GridRow &r,
                                              int j) {
    const double2 pj = xy_of<D3>(st, sc, j);
    double d2;
    if (D3) {
        const double m2dot = __fma_rn(r.qz, sc.pos_new[2 * (size_t)st.cap + j], __fma_rn(r.qy, pj.y, __dmul_rn(r.qx, pj.x)));
        d2 = __dadd_rn(__dadd_rn(m2dot, r.xxi), sc.xx_new[j]);
    } else {
        d2 = pred_d2(r.qx, r.qy, r.xxi, pj.x, pj.y, sc.xx_new[j]);
    }
    const double rs = r.ri + sc.rinf[j];
    if (!(d2 <= rs * rs)) return;
    const bool hit = D3 ? sqrt(fmax(d2, 0.0)) <= __dadd_rn(st.radius[i], st.radius[j])
                        : pred_exact(d2, st.radius[i], st.radius[j]);
    if (hit) emit_pair(st, sc, i, j);
}

// The cell-sorted order is cut into chunks of GRID_CHUNK slots whose boundaries are snapped to
// CELL boundaries (the order of the planets inside a cell comes from atomics and differs from rank
// to rank, the cell populations do not), and the chunks are dealt round-robin to the ranks: the
// crowded inner cells of a disk -- most of the candidate work -- are spread over all of them.
// bounds[k] = first slot of the first cell starting at or after slot k * GRID_CHUNK.
constexpr int GRID_CHUNK = 1024;
constexpr int GRID_HEAVY = 128;   // candidates above which a row goes to the warp-per-row pass

__global__ void __launch_bounds__(128) k_grid_ranges(cosmo_state st, cosmo_scratch sc, int c) {
    const int ncell = c * c;
    const int total = sc.cell_start[ncell];
    const int nchunk = (total + GRID_CHUNK - 1) / GRID_CHUNK;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k <= nchunk; k += gridDim.x * blockDim.x) {
        const long long target = (long long)k * GRID_CHUNK;
        int bound = total;
        if (target < total) {
            int lo = 0, hi = ncell;                    // smallest cell with cell_start[cell] >= target
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (sc.cell_start[mid] >= target) hi = mid; else lo = mid + 1;
            }
            bound = sc.cell_start[lo];
        }
        sc.chunk_alive[k] = bound;                     // scratch shared with stage C's compaction (later)
    }
}

// One thread per non-giant target row, taken in CELL-SORTED order: the lanes of a warp then
// work on planets of the same or neighbouring cells, so their candidate loops have similar trip
// counts (the disk's surface density varies 20x between its inner and outer edge) and their
// gathers hit the same lines.  Every row is produced by exactly one rank.
template <bool D3>
__global__ void __launch_bounds__(128)
k_grid_detect(cosmo_state st, cosmo_scratch sc, int shard_index, int shard_count, int c) {
    const int total = sc.cell_start[c * c];
    const int nchunk = (total + GRID_CHUNK - 1) / GRID_CHUNK;
    unsigned cand_total = 0;
  for (int chunk = shard_index + blockIdx.x * shard_count; chunk < nchunk; chunk += gridDim.x * shard_count) {
   const int s_lo = sc.chunk_alive[chunk], s_hi = sc.chunk_alive[chunk + 1];
   // blockIdx.y splits the chunk's slots: a normal chunk is one slot per thread, an oversized one
   // (a crowded or border cell is never split across chunks) is still shared by gridDim.y blocks
   for (int slot = s_lo + blockIdx.y * blockDim.x + threadIdx.x; slot < s_hi; slot += gridDim.y * blockDim.x) {
    const int i = sc.cell_items[slot];
    const int cell = sc.cell_of[i];
    const GridRow row = load_row<D3>(st, sc, i);
    const int cx = cell % c, cy = cell / c;
    // candidate ranges: the three cells of a grid row are contiguous in the sorted item array
    int s0[3], s1[3];
    long long cand = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = cy + dy;
        const bool in = yy >= 0 && yy < c;
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, c - 1);
        s0[dy + 1] = in ? sc.cell_start[yy * c + x0] : 0;
        s1[dy + 1] = in ? sc.cell_start[yy * c + x1 + 1] : 0;
        cand += s1[dy + 1] - s0[dy + 1];
    }
    const int ng = (int)min((long long)st.status[COSMO_ST_N_GIANTS], (long long)sc.giant_cap);
    cand_total += (unsigned)cand;
    if (cand + ng > GRID_HEAVY) {
        // a crowded neighbourhood (dense inner disk, logarithmic rim): one thread walking thousands
        // of candidates would be the critical path of the whole frame -> warp-per-row pass
        const unsigned long long h =
            atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_HEAVY]), 1ull);
        sc.label[h] = i;       // `label` is free until stage C
        continue;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
        for (int s = s0[d]; s < s1[d]; ++s) {
            const int j = sc.cell_items[s];
            if (j != i) test_and_emit<D3>(st, sc, i, row, j);
        }
    for (int g = 0; g < ng; ++g) test_and_emit<D3>(st, sc, i, row, sc.giants[g]);
   }
  }
    // diagnostic only (all lanes are back together after the loop)
    const unsigned tot = __reduce_add_sync(0xffffffffu, cand_total);
    if ((threadIdx.x & 31) == 0 && tot)
        atomicAdd(reinterpret_cast<unsigned long long *>(&st.status[COSMO_ST_N_CAND]), (unsigned long long)tot);
}

// Warp-per-row pass for the rows k_grid_detect set aside: the lanes stride over the candidates.
template <bool D3>
__global__ void __launch_bounds__(256) k_grid_detect_heavy(cosmo_state st, cosmo_scratch sc, int c) {
    const int nh = (int)st.status[COSMO_ST_N_HEAVY];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const int ng = (int)min((long long)st.status[COSMO_ST_N_GIANTS], (long long)sc.giant_cap);
    for (int h = warp; h < nh; h += nwarp) {
        const int i = sc.label[h];
        const int cell = sc.cell_of[i];
        const GridRow row = load_row<D3>(st, sc, i);
        const int cx = cell % c, cy = cell / c;
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = cy + dy;
            if (yy < 0 || yy >= c) continue;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, c - 1);
            const int s0 = sc.cell_start[yy * c + x0], s1 = sc.cell_start[yy * c + x1 + 1];
            for (int s = s0 + lane; s < s1; s += 32) {
                const int j = sc.cell_items[s];
                if (j != i) test_and_emit<D3>(st, sc, i, row, j);
            }
        }
        for (int g = lane; g < ng; g += 32) test_and_emit<D3>(st, sc, i, row, sc.giants[g]);
    }
}

// rows of the giants: giant g against every planet (blockIdx.y strides over the giants)
template <bool D3>
__global__ void __launch_bounds__(256)
k_giant_rows(cosmo_state st, cosmo_scratch sc, int i_begin, int i_end) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int i_hi = min(i_end, n);
    const int ng = (int)min((long long)st.status[COSMO_ST_N_GIANTS], (long long)sc.giant_cap);
    for (int gi = blockIdx.y; gi < ng; gi += gridDim.y) {
        const int g = sc.giants[gi];
        if (g < i_begin || g >= i_hi) continue;
        const GridRow row = load_row<D3>(st, sc, g);
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
            if (j != g) test_and_emit<D3>(st, sc, g, row, j);
    }
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
constexpr int DET_BLOCK = 128, DET_IB = 2, DET_TJ = 256;
constexpr int DETS_BLOCK = 64, DETS_IB = 1, DETS_TJ = 64;

int detect_residency(int *rows_per_cta, int *tile) {
    int nb = 0;
    *rows_per_cta = DET_BLOCK * DET_IB; *tile = DET_TJ;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_detect_allpairs<DET_BLOCK, DET_IB, DET_TJ>, DET_BLOCK, 0) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}

template <bool D3>
static int launch_grid(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cudaStream_t s) {
    const int rows = p.i_end - p.i_begin;
    const int c = p.cells_per_side;
    if (c < 1 || (long long)c * c > sc.grid_cells || !sc.cell_count || !sc.giants) {
        set_error("grid broad phase: cells_per_side=%d does not fit grid_cells=%d", c, sc.grid_cells);
        return COSMO_E_ARG;
    }
    int nb = (p.n_upper + 255) / 256;
    if (nb < 1) nb = 1;
    if (nb > 148 * 8) nb = 148 * 8;
    k_grid_max<<<nb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_max"));
    const int w = p.grid_log_cells < 0 ? 0 : p.grid_log_cells;
    if (c - 2 * w < 1) {
        set_error("grid broad phase: %d logarithmic cells per side do not fit cells_per_side=%d", w, c);
        return COSMO_E_ARG;
    }
    k_grid_count<D3><<<nb, 256, 0, s>>>(st, sc, p.cell_size, p.grid_origin_x, p.grid_origin_y, p.giant_radius, c, w);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_count"));
    COSMO_TRY(launch_scan_exclusive(sc.cell_count, sc.cell_start, sc.scan_blk, nullptr, (long long)c * c, 0, s));
    k_grid_fill<<<nb, 256, 0, s>>>(st, sc);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_fill"));
    const int sh_n = p.shard_count > 0 ? p.shard_count : 1;
    const int sh_i = p.shard_count > 0 ? p.shard_index : 0;
    const int nchunk = (p.n_upper + GRID_CHUNK - 1) / GRID_CHUNK + 1;
    k_grid_ranges<<<(nchunk + 127) / 128, 128, 0, s>>>(st, sc, c);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_ranges"));
    const int blocks = (nchunk + sh_n - 1) / sh_n;
    k_grid_detect<D3><<<dim3(blocks > 0 ? blocks : 1, GRID_CHUNK / 128), 128, 0, s>>>(st, sc, sh_i, sh_n, c);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_detect"));
    k_grid_detect_heavy<D3><<<296, 256, 0, s>>>(st, sc, c);
    COSMO_TRY(COSMO_LAUNCH_CHECK("k_grid_detect_heavy"));
    if (rows > 0) {
        k_giant_rows<D3><<<dim3(nb > 296 ? 296 : nb, 8), 256, 0, s>>>(st, sc, p.i_begin, p.i_end);
        COSMO_TRY(COSMO_LAUNCH_CHECK("k_giant_rows"));
    }
    return COSMO_OK;
}

int launch_detect_grid(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cudaStream_t s,
                       bool planes3) {
    return planes3 ? launch_grid<true>(st, sc, p, s) : launch_grid<false>(st, sc, p, s);
}

int launch_detect(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                  cudaStream_t s) {
    if (!(p.opts & COSMO_OPT_COLLISIONS)) return COSMO_OK;
    const int rows = p.i_end - p.i_begin;
    if (rows <= 0 && p.detect_mode == 0) return COSMO_OK;
    if (p.detect_mode == 1) return launch_detect_grid(st, sc, p, s, false);
    int jsplit = p.jsplit_detect < 1 ? 1 : p.jsplit_detect;
    if (p.small_tiles) {
        const int per = DETS_BLOCK * DETS_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        k_detect_allpairs<DETS_BLOCK, DETS_IB, DETS_TJ><<<grid, DETS_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit);
    } else {
        const int per = DET_BLOCK * DET_IB;
        dim3 grid((rows + per - 1) / per, jsplit);
        k_detect_allpairs<DET_BLOCK, DET_IB, DET_TJ><<<grid, DET_BLOCK, 0, s>>>(st, sc, p.i_begin, p.i_end, jsplit);
    }
    return COSMO_LAUNCH_CHECK("k_detect_allpairs");
}

}  // namespace cosmo
