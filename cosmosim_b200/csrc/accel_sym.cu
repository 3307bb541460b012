// Stage A, FP32 fast flavour, SYMMETRIC variant: every unordered pair {i, j} is evaluated once
// and its (antisymmetric) force is accumulated on both planets -- 6 FMA-pipe instructions and
// half a MUFU per ORDERED interaction instead of 9 and 1.  Same entry point / reference lines as
// accel.cu (acc_blas, cosmosim/util/blas.py:4-41; integrator universe_old.py:177-178).
//
// Decomposition.  Planets are cut into blocks of 1024.  A CTA of 8 warps works on TWO row blocks
// at a time ("halves": warps 0-3 and 4-7) and walks a chunk of COLUMN blocks J >= I.  A warp keeps
// 256 planets of its row block in registers (8 per lane: the "row" side, accumulated in packed FP32
// registers).  A column block J, staged in shared memory by TMA bulk copies, is cut into 8 groups of
// 128; the CTA runs 8 rounds, in round `it` warp w works on column group g = (w + it) & 7, so no
// two warps touch the same column group in the same round.  Inside a round the 128 column planets
// are dealt to the lanes in chunks of 4 and ROTATE through the warp: at step s lane l reads chunk
// (l + s) & 31 from shared memory and holds that chunk's partial column accumulators, which it
// then hands to lane l-1 with 8 SHFLs -- after 32 steps every row planet of the warp has met every
// column planet of the group and the accumulators are back in their home lane.  Per step and lane:
// 8 x 4 pairs = 192 FFMA2-class instructions + 32 MUFU.RSQ + 3 LDS.128 + 8 SHFL (8 row planets per
// lane instead of 4 halve the SHFL / LDS / address instructions per pair; profiles/r02_kernel_experiments.md).
//
// A CTA serves R (4, 2; 1 on request) row blocks against its column chunk (super tile, see the kernel): the
// column sums of a column block J over those row blocks go to colbuf[super row][j] (double2: the
// eight rounds and the R row blocks are summed in FP64, exactly, so no tiling changes a bit); the row
// sums to partial[chunk][i] (double2).  k_sym_reduce adds, for every planet, its row partials and
// the column partials of all super rows below it in ascending order -- deterministic, no atomics --
// and runs the integrator epilogue.  The diagonal pair (J == I) runs the same schedule without the
// column side, which evaluates each of its ordered pairs exactly once.  HBM: colbuf is written and
// read once per frame, 16 B x N^2 / (2048 R): 2.1 GB at N = 2^20 with R = 4.
//
// What runs by default is a POST-SCHEDULED copy of this kernel's SASS (cosmosim_b200/sass_tune.py at build
// time, sym_tuned.cu at run time): an FFMA2 with three live register-pair operands costs three cycles on
// this part unless one comes from the operand-reuse cache, and the four accumulation FFMA2 of a pair group
// (a 2 x 2 outer product {s_j, s_i} x {dx, dy}) only hit that cache when they are neighbours -- which
// ptxas's latency-first schedule does not give them.  Same instructions, same bits; this kernel as
// compiled here is the fallback (COSMO_SYM_TUNED=0).
#include <math.h>

#include "cosmo_device.cuh"
#include "cosmo_internal.h"
#include "integrate.cuh"

namespace cosmo {

constexpr int SYM_BLOCK = 1024;   // planets per block
constexpr int SYM_GROUP = 128;    // column planets per warp and round (4 per lane)
constexpr int SYM_WARPS = 8;
constexpr int SYM_RL = 8;         // row planets per lane
constexpr int SYM_WPB = SYM_BLOCK / (32 * SYM_RL);   // warps per row block (4)
constexpr int SYM_NH = SYM_WARPS / SYM_WPB;          // row blocks a CTA works on at once (2)
constexpr int SYM_ROW_GROUP = 32 * SYM_RL;           // row planets per warp (256)
constexpr int SYM_THREADS = SYM_WARPS * 32;
#define COSMO_F32_TINY_SYM 8.673617379884035e-19f /* 2^-60 softening, see accel.cu */

// Row blocks are dealt to the ranks of a sharded run boustrophedon-wise (0 1 .. W-1, W-1 .. 1 0, ...):
// the work of row block I is proportional to nblk - I (the triangle), and plain round-robin would
// hand rank 0 the longest row of every round -- a 3 % imbalance at 8 ranks; this way it is O(1 / rows).
__host__ __device__ __forceinline__ int sym_row_of(int l, int rank, int world) {
    return l * world + ((l & 1) ? world - 1 - rank : rank);
}
__host__ __device__ __forceinline__ bool sym_owns(int I, int rank, int world) {
    const int l = I / world, rr = I % world;
    return ((l & 1) ? world - 1 - rr : rr) == rank;
}

// (x, x) built where it is used: ptxas then folds it into the broadcast operand form of the packed
// instruction (FADD2 d, a, -Rx.F32) and the row operands stay one register each
__device__ __forceinline__ uint64_t dup2(float x) {
    uint64_t r;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src_lane) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    lo = __shfl_sync(0xffffffffu, lo, src_lane);
    hi = __shfl_sync(0xffffffffu, hi, src_lane);
    return ((uint64_t)hi << 32) | lo;
}

// Super-tile: a CTA serves R consecutive local row blocks (rows I_0 < I_1 < ... ) against one chunk
// of column blocks.  A column tile is fetched ONCE for all R row blocks and its column sums are
// accumulated over them in shared memory before they go to colbuf -- one colbuf row per SUPER row,
// so the column-partial traffic (the only sizeable DRAM traffic of the pass) and the buffer itself
// shrink (by R / 2: the sums over several row blocks are kept in FP64, see s_c).  The row sums of the
// R blocks live in shared memory as doubles for the CTA's lifetime.
template <int R>
__global__ void __launch_bounds__(SYM_THREADS, 2)
k_accel_f32_sym(cosmo_state st, cosmo_scratch sc, int shard_index, int shard_count, int chunk) {
    constexpr int RL = SYM_RL, WPB = SYM_WPB, NH = SYM_NH, ROW_GROUP = SYM_ROW_GROUP;
    extern __shared__ __align__(16) unsigned char s_dyn[];
    float (*s_x)[SYM_BLOCK] = reinterpret_cast<float (*)[SYM_BLOCK]>(s_dyn);
    float (*s_y)[SYM_BLOCK] = s_x + 2;
    float (*s_m)[SYM_BLOCK] = s_x + 4;
    // column accumulators of the current J over the R row blocks: [x|y][1024], FP64 -- the sum of the
    // FP32 per-block sums is then exact, i.e. independent of how row blocks are grouped into super
    // rows and dealt to ranks (a sharded run reproduces the single-GPU sums bit for bit)
    double (*s_c)[SYM_BLOCK] = reinterpret_cast<double (*)[SYM_BLOCK]>(s_x + 6);
    // FP64 row sums of the R row blocks: [r][1024] (x, y)
    double2 (*s_row)[SYM_BLOCK] = reinterpret_cast<double2 (*)[SYM_BLOCK]>(s_x + 10);
    __shared__ __align__(8) uint64_t s_bar[2];

    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int nblk = (n + SYM_BLOCK - 1) / SYM_BLOCK;
    const int I_first = sym_row_of(blockIdx.x * R, shard_index, shard_count);
    if (I_first >= nblk) return;
    const int J0 = I_first + blockIdx.y * chunk;
    if (J0 >= nblk) return;
    const int J1 = min(nblk, J0 + chunk);
    const int tid = threadIdx.x, lane = tid & 31;
    const int w = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler too (no BRA.DIV around the SHFLs)
    const int half = w / WPB, wq = w % WPB;

    for (int q = tid; q < R * SYM_BLOCK; q += SYM_THREADS) s_row[0][q] = make_double2(0.0, 0.0);
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    constexpr uint32_t kBytes = SYM_BLOCK * 3 * sizeof(float);
    auto issue = [&](int J, int b) {
        const size_t j0 = (size_t)J * SYM_BLOCK;
        fence_proxy_async();
        mbar_expect_tx(&s_bar[b], kBytes);
        bulk_g2s(&s_x[b][0], sc.pk_x + j0, SYM_BLOCK * sizeof(float), &s_bar[b]);
        bulk_g2s(&s_y[b][0], sc.pk_y + j0, SYM_BLOCK * sizeof(float), &s_bar[b]);
        bulk_g2s(&s_m[b][0], sc.pk_m + j0, SYM_BLOCK * sizeof(float), &s_bar[b]);
    };
    if (tid == 0) issue(J0, 0);

    const uint64_t eps2 = pack2(COSMO_F32_TINY_SYM, COSMO_F32_TINY_SYM);
    double2 *colbuf = reinterpret_cast<double2 *>(sc.colbuf);
    const size_t col_row = (size_t)blockIdx.x * (size_t)st.cap;   // local super-row index of this rank

    float xi[RL], yi[RL], nmi[RL];  // row side: RL planets per lane, strided by 32 inside the warp's group
    auto load_rows = [&](int I) {   // coalesced, L2-resident
#pragma unroll
        for (int k = 0; k < RL; ++k) {
            const int row = I * SYM_BLOCK + wq * ROW_GROUP + k * 32 + lane;
            xi[k] = sc.pk_x[row];
            yi[k] = sc.pk_y[row];
            nmi[k] = -sc.pk_m[row];
        }
    };
    if (R == 1 && half == 0) load_rows(I_first);      // a single row block: its operands stay in registers

    for (int J = J0; J < J1; ++J) {
        const int it_j = J - J0;
        const int b = it_j & 1;
        if (tid == 0 && J + 1 < J1) issue(J + 1, b ^ 1);
        for (int q = tid; q < SYM_BLOCK; q += SYM_THREADS) { s_c[0][q] = 0.0; s_c[1][q] = 0.0; }
        mbar_wait(&s_bar[b], (it_j >> 1) & 1);
        __syncthreads();
        bool any_col = false;

#pragma unroll 1
        for (int r0 = 0; r0 < R; r0 += NH) {
            // rows are ascending: once the first block of a pass lies beyond this column block, all do
            const int Ia = sym_row_of(blockIdx.x * R + r0, shard_index, shard_count);
            if (Ia >= nblk || Ia > J) break;
            const int r = r0 + half;
            const int I = r < R ? sym_row_of(blockIdx.x * R + r, shard_index, shard_count) : nblk;
            const bool active = I < nblk && I <= J;   // warp-uniform
            const bool diag = (J == I);
#pragma unroll
            for (int h = 0; h < NH; ++h) {   // (CTA-uniform) does any half add column sums for this J?
                const int Ih = r0 + h < R ? sym_row_of(blockIdx.x * R + r0 + h, shard_index, shard_count) : nblk;
                any_col = any_col || (Ih < nblk && Ih < J);
            }
            if (R > 1 && active) load_rows(I);
            uint64_t rx2[RL], ry2[RL];
#pragma unroll
            for (int k = 0; k < RL; ++k) { rx2[k] = 0ull; ry2[k] = 0ull; }

            for (int it = 0; it < SYM_WARPS; ++it) {
                const int g = (w + it) & (SYM_WARPS - 1);
                if (active) {
                    const float *gx = &s_x[b][g * SYM_GROUP];
                    const float *gy = &s_y[b][g * SYM_GROUP];
                    const float *gm = &s_m[b][g * SYM_GROUP];
                    uint64_t tx2[2] = {0ull, 0ull}, ty2[2] = {0ull, 0ull};   // travelling column sums
#pragma unroll 2
                    for (int s = 0; s < 32; ++s) {
                        const int c = (lane + s) & 31;
                        const float4 X = *reinterpret_cast<const float4 *>(gx + 4 * c);
                        const float4 Y = *reinterpret_cast<const float4 *>(gy + 4 * c);
                        const float4 M = *reinterpret_cast<const float4 *>(gm + 4 * c);
                        const uint64_t xj2[2] = {pack2(X.x, X.y), pack2(X.z, X.w)};
                        const uint64_t yj2[2] = {pack2(Y.x, Y.y), pack2(Y.z, Y.w)};
                        const uint64_t mj2[2] = {pack2(M.x, M.y), pack2(M.z, M.w)};
                        // (column half outside, row planets inside and the products written (s, d): the
                        // schedule ptxas derives from this order has the fewest instructions and operand
                        // reads of the orders tried -- the results are the same bits for all of them)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
#pragma unroll
                            for (int k = 0; k < RL; ++k) {
                                const uint64_t xik = dup2(xi[k]), yik = dup2(yi[k]);
                                const uint64_t nmik = dup2(nmi[k]);
                                const uint64_t dx = sub2(xj2[h], xik);
                                const uint64_t dy = sub2(yj2[h], yik);
                                const uint64_t r2 = fma2(dy, dy, fma2(dx, dx, eps2));
                                float r2a, r2b;
                                unpack2(r2, r2a, r2b);
                                const uint64_t rinv = pack2(rsqrt_approx_f32(r2a), rsqrt_approx_f32(r2b));
                                const uint64_t rinv3 = mul2(mul2(rinv, rinv), rinv);
                                const uint64_t sj = mul2(mj2[h], rinv3);     // pull of the column planets on row k
                                const uint64_t nsi = mul2(rinv3, nmik);      // minus the pull of row k on them
                                rx2[k] = fma2(sj, dx, rx2[k]);
                                ry2[k] = fma2(sj, dy, ry2[k]);
                                tx2[h] = fma2(nsi, dx, tx2[h]);
                                ty2[h] = fma2(nsi, dy, ty2[h]);
                            }
                        }
                        // hand the chunk's column sums to the lane that meets this chunk next
                        const int src = (lane + 1) & 31;
                        tx2[0] = shfl64(tx2[0], src); tx2[1] = shfl64(tx2[1], src);
                        ty2[0] = shfl64(ty2[0], src); ty2[1] = shfl64(ty2[1], src);
                    }
                    if (!diag) {   // home again: lane l holds the sums of chunk l of group g
                        float a0, a1, a2, a3, b0, b1, b2, b3;
                        unpack2(tx2[0], a0, a1); unpack2(tx2[1], a2, a3);
                        unpack2(ty2[0], b0, b1); unpack2(ty2[1], b2, b3);
                        double4 *px = reinterpret_cast<double4 *>(&s_c[0][g * SYM_GROUP + 4 * lane]);
                        double4 *py = reinterpret_cast<double4 *>(&s_c[1][g * SYM_GROUP + 4 * lane]);
                        double4 cx = *px, cy = *py;
                        cx.x += (double)a0; cx.y += (double)a1; cx.z += (double)a2; cx.w += (double)a3;
                        cy.x += (double)b0; cy.y += (double)b1; cy.z += (double)b2; cy.w += (double)b3;
                        *px = cx;
                        *py = cy;
                    }
                }
                // Fold the FP32 row sums into the FP64 ones after every 4 rounds: the two halves walk the
                // column groups in orders that differ by a swap of the two 4-round segments ((w + it) & 7
                // with w = wq and w = wq + 4), so each segment's FP32 sum -- and, the FP64 fold being exact,
                // the total -- does not depend on which half served the row block (tilings, rank counts).
                if ((it & 3) == 3 && active) {
#pragma unroll
                    for (int k = 0; k < RL; ++k) {
                        float ax, ax2, ay, ay2;
                        unpack2(rx2[k], ax, ax2);
                        unpack2(ry2[k], ay, ay2);
                        double2 &acc = s_row[r][wq * ROW_GROUP + k * 32 + lane];   // each lane owns its rows
                        acc.x += (double)ax + (double)ax2;
                        acc.y += (double)ay + (double)ay2;
                        rx2[k] = 0ull; ry2[k] = 0ull;
                    }
                }
                __syncthreads();   // the next round moves every warp to another column group
            }
        }
        if (any_col) {   // column sums of this block over the R row blocks: one colbuf row per super row
            double2 *dst = colbuf + col_row + (size_t)J * SYM_BLOCK;
            for (int q = tid; q < SYM_BLOCK; q += SYM_THREADS) dst[q] = make_double2(s_c[0][q], s_c[1][q]);
        }
        __syncthreads();   // column accumulators and buffer b are free again
    }
    // row partials of every row block that met a column block of this chunk (I < J1)
    double2 *partial = reinterpret_cast<double2 *>(sc.partial);
    for (int r = 0; r < R; ++r) {
        const int I = sym_row_of(blockIdx.x * R + r, shard_index, shard_count);
        if (I >= nblk || I >= J1) break;
        for (int q = tid; q < SYM_BLOCK; q += SYM_THREADS)
            partial[(size_t)blockIdx.y * st.cap + (size_t)I * SYM_BLOCK + q] = s_row[r][q];
    }
}

// a_k = unscale * (sum of the row partials of block K + sum over I < K of colbuf[I][k]), then the
// integrator epilogue (accel.cu).  With `defer` the per-rank sum goes to scratch.acc instead
// (multi-GPU: all-reduced by the caller, then k_sym_integrate).
__global__ void __launch_bounds__(256)
k_sym_reduce(cosmo_state st, cosmo_scratch sc, int shard_index, int shard_count, int chunk, int R, double unscale,
             double G, double dt, int store_acc, int defer, int near) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int nblk = (n + SYM_BLOCK - 1) / SYM_BLOCK;
    const int K = k / SYM_BLOCK;
    double ax = 0.0, ay = 0.0;
    if (sym_owns(K, shard_index, shard_count)) {
        // the chunks of K's super row start at its first row block: K is met by the chunks from
        // (K - I_first) / chunk on
        const double2 *partial = reinterpret_cast<const double2 *>(sc.partial);
        const int I_first = sym_row_of((K / shard_count) / R * R, shard_index, shard_count);
        const int nchunk = (nblk - I_first + chunk - 1) / chunk;
        for (int c = (K - I_first) / chunk; c < nchunk; ++c) {
            const double2 v = partial[(size_t)c * st.cap + k];
            ax += v.x;
            ay += v.y;
        }
    }
    // super rows of this rank with a row block below K, in ascending order (super row r starts at
    // block sym_row_of(r * R)); each holds the (exact) column sums of its row blocks I < K
    const double2 *colbuf = reinterpret_cast<const double2 *>(sc.colbuf);
    for (int r = 0; sym_row_of(r * R, shard_index, shard_count) < K; ++r) {
        const double2 v = __ldcs(&colbuf[(size_t)r * st.cap + k]);
        ax += v.x;
        ay += v.y;
    }
    ax *= unscale;
    ay *= unscale;
    if (near) {   // FP64 near-field corrections of the rows this rank evaluated (nearfield.cu; zero elsewhere)
        const double2 c = reinterpret_cast<const double2 *>(sc.corr)[k];
        ax += c.x;
        ay += c.y;
    }
    if (defer == 2) {
        // peer exchange: this rank's partial sums go straight into slot `shard_index` of every
        // rank's exchange buffer (NVLink peer stores, coalesced 16 B per lane)
        const double2 v = make_double2(ax, ay);
        for (int r = 0; r < sc.n_peers; ++r)
            reinterpret_cast<double2 *>(sc.peer_acc[r])[(size_t)shard_index * st.cap + k] = v;
    } else if (defer) {
        reinterpret_cast<double2 *>(sc.acc)[k] = make_double2(ax, ay);
    } else {
        integrate_store(st, sc, k, ax, ay, G, dt, store_acc != 0);
    }
}

__global__ void __launch_bounds__(256)
k_sym_integrate(cosmo_state st, cosmo_scratch sc, double G, double dt, int peer, int self) {
    const int n = (int)st.status[COSMO_ST_N_ACTIVE];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double2 a;
    if (peer) {   // add the ranks' slots of THIS rank's exchange buffer in rank order: deterministic
        a = make_double2(0.0, 0.0);
        const double2 *x = reinterpret_cast<const double2 *>(sc.peer_acc[self]);
        for (int r = 0; r < sc.n_peers; ++r) {
            const double2 v = __ldcg(&x[(size_t)r * st.cap + k]);
            a.x += v.x;
            a.y += v.y;
        }
    } else {
        a = reinterpret_cast<const double2 *>(sc.acc)[k];
    }
    // sc.acc holds the all-reduced sums S; integrate_store multiplies by G and rewrites acc = G S
    integrate_store(st, sc, k, a.x, a.y, G, dt, true);
}

template <int R>
static int launch_sym_kernel(const cosmo_state &st, const cosmo_scratch &sc, dim3 grid, int sh_i, int sh_n, int chunk,
                             cudaStream_t s) {
    constexpr size_t kTma = 6 * SYM_BLOCK * sizeof(float);                 // x, y, m double-buffered
    constexpr size_t kShared = kTma + 2 * SYM_BLOCK * sizeof(double)       // + the FP64 column accumulators
                               + (size_t)R * SYM_BLOCK * sizeof(double2);  // + the FP64 row sums of R row blocks
    static bool attr_set = false;
    // the post-scheduled copy of this kernel (sym_tuned.cu, sass_tune.py) if the build produced one
    static const char *const kName = R == 4   ? "_ZN5cosmo15k_accel_f32_symILi4EEEv11cosmo_state13cosmo_scratchiii"
                                     : R == 2 ? "_ZN5cosmo15k_accel_f32_symILi2EEEv11cosmo_state13cosmo_scratchiii"
                                              : "_ZN5cosmo15k_accel_f32_symILi1EEEv11cosmo_state13cosmo_scratchiii";
    const void *tuned = tuned_kernel(kName, (int)kShared);
    if (tuned) {
        cosmo_state st_v = st;
        cosmo_scratch sc_v = sc;
        void *args[] = {&st_v, &sc_v, &sh_i, &sh_n, &chunk};
        COSMO_TRY(check_cuda(cudaLaunchKernel(tuned, grid, dim3(SYM_THREADS), args, kShared, s), "k_accel_f32_sym (tuned)"));
        return COSMO_OK;
    }
    if (!attr_set) {   // R = 4: 104 KB of dynamic shared memory (two CTAs per SM)
        COSMO_TRY(check_cuda(cudaFuncSetAttribute(k_accel_f32_sym<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)kShared), "cudaFuncSetAttribute(k_accel_f32_sym)"));
        attr_set = true;
    }
    k_accel_f32_sym<R><<<grid, SYM_THREADS, kShared, s>>>(st, sc, sh_i, sh_n, chunk);
    return COSMO_LAUNCH_CHECK("k_accel_f32_sym");
}

int launch_accel_sym(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                     cudaStream_t s) {
    const int n = p.n_upper;
    if (n <= 0) return COSMO_OK;
    const int nblk = (n + SYM_BLOCK - 1) / SYM_BLOCK;
    const int sh_n = p.shard_count > 0 ? p.shard_count : 1;
    const int sh_i = p.shard_count > 0 ? p.shard_index : 0;
    int rows_local = nblk / sh_n;                     // full boustrophedon rounds, plus a partial one
    if (sym_row_of(rows_local, sh_i, sh_n) < nblk) ++rows_local;
    int chunk = p.sym_chunk > 0 ? p.sym_chunk : 32;
    while ((nblk + chunk - 1) / chunk > sc.jsplit_max) ++chunk;
    // row blocks per CTA: p.sym_rows, raised if the column-partial buffer is too small for it
    int R = p.sym_rows >= 4 ? 4 : (p.sym_rows == 1 ? 1 : 2);   // unset (0): 2, the smallest tiling that fills the CTA
    auto rows_needed = [&](int r) { return 2LL * ((rows_local + r - 1) / r); };   // in float2 rows: a double2 row is two
    if (sc.colbuf_rows > 0)
        while (R < 4 && rows_needed(R) > sc.colbuf_rows) R *= 2;
    const int super_rows = (rows_local + R - 1) / R;
    if (!sc.colbuf || rows_needed(R) > sc.colbuf_rows) {
        set_error("symmetric accel: colbuf has %lld rows, %lld needed", (long long)sc.colbuf_rows, rows_needed(R));
        return COSMO_E_ARG;
    }
    const double unscale = scalbn(1.0, p.mass_exp - 2 * p.pos_exp);
    if (rows_local > 0) {
        dim3 grid(super_rows, (nblk + chunk - 1) / chunk);
        if (R == 4) COSMO_TRY(launch_sym_kernel<4>(st, sc, grid, sh_i, sh_n, chunk, s));
        else if (R == 2) COSMO_TRY(launch_sym_kernel<2>(st, sc, grid, sh_i, sh_n, chunk, s));
        else COSMO_TRY(launch_sym_kernel<1>(st, sc, grid, sh_i, sh_n, chunk, s));
    }
    const int defer = (p.opts & COSMO_OPT_DEFER_INTEGRATE) ? ((p.opts & COSMO_OPT_PEER_EXCHANGE) ? 2 : 1) : 0;
    k_sym_reduce<<<(n + 255) / 256, 256, 0, s>>>(st, sc, sh_i, sh_n, chunk, R, unscale, p.G, p.dt,
                                                 (p.opts & COSMO_OPT_STORE_ACC) ? 1 : 0, defer, p.near_mode != 0);
    return COSMO_LAUNCH_CHECK("k_sym_reduce");
}

int launch_sym_integrate(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                         cudaStream_t s) {
    if (p.n_upper <= 0) return COSMO_OK;
    const int peer = (p.opts & COSMO_OPT_PEER_EXCHANGE) ? 1 : 0;
    if (peer && (sc.n_peers < 1 || sc.n_peers > COSMO_MAX_PEERS || p.shard_index >= sc.n_peers || !sc.peer_acc[p.shard_index])) {
        set_error("peer exchange: n_peers=%d, rank %d", sc.n_peers, p.shard_index);
        return COSMO_E_ARG;
    }
    k_sym_integrate<<<(p.n_upper + 255) / 256, 256, 0, s>>>(st, sc, p.G, p.dt, peer, p.shard_index);
    return COSMO_LAUNCH_CHECK("k_sym_integrate");
}

}  // namespace cosmo
