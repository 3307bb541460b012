/*
 * cosmosim_b200_state3d.h -- C ABI of the B200-native step of cosmosim's NEW (3-D, batch) API.
 *
 * Replaces State.interact() of the reference (cosmosim/core/universe.py:94-129), i.e.
 *     acc_blas(p0, m, G) with (N,3) positions        universe.py:101 (cosmosim/util/blas.py:4-41)
 *     v = v0 + a dt; p = p0 + v dt                    universe.py:103-104
 *     pairwise_distances(p) <= r_i + r_j              universe.py:107-108
 *     the ordered absorb loop + Object.absorb/destroy universe.py:111-125, :39-54
 *     removal of the absorbed objects                 universe.py:126-128
 * SURVEY.md section 8f rank 3.  Conventions are those of cosmosim_b200.h (plain C, device
 * pointers owned by the caller, asynchronous launches on `stream`, 0 / negative COSMO_E_* return
 * codes, cosmo_last_error()); the status block uses the COSMO_ST_* slots of that header
 * (N_ACTIVE, STEP = State.iteration, N_EVENTS, N_PAIRS, N_DEAD, ERROR).
 *
 * Data layout in HBM: structure of arrays, one slot per object of State.objects in list order;
 * vectors are stored as three planes so that every kernel reads unit-stride doubles:
 *     pos, vel          double[3][cap]   plane k = component k
 *     mass              double[cap]      float(mass)
 *     mass_hi, mass_lo  uint64[cap]      the integer mass when COSMO3_F_MASS_INT (Python ints are
 *                                        exact; the example scenes draw them with randint)
 *     density           double[cap]      float(density); an exact integer < 2^53 when COSMO3_F_DENS_INT
 *     radius            double[cap]      Object.get_radius() of the current mass and density
 *     id                int32[cap]       index in the initial object list
 *     flags             uint8[cap]
 * `cap` is a multiple of COSMO_PAD.
 */
#ifndef COSMOSIM_B200_STATE3D_H
#define COSMOSIM_B200_STATE3D_H

#include <stdint.h>

#include "cosmosim_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define COSMO3_F_MASS_INT 1u
#define COSMO3_F_DENS_INT 2u

/* option bits */
#define COSMO3_OPT_COLLISIONS 1u /* State.interact(collisions=True) */
#define COSMO3_OPT_FP32 4u       /* FP32 fast path for the all-pairs sum (state stays FP64) */
#define COSMO3_OPT_STORE_ACC 8u  /* also write the accelerations to scratch.acc */
#define COSMO3_OPT_F64_SYM 16u   /* FP64 path: symmetric kernel (each unordered pair once); needs colbuf; row
                                    blocks are dealt round-robin to shard_count ranks */
#define COSMO3_OPT_F32_SYM 64u   /* FP32 path: symmetric kernel, same requirements as COSMO3_OPT_F64_SYM */
#define COSMO3_OPT_DEFER_INTEGRATE 32u /* symmetric kernel, multi-GPU: leave this rank's partial sums in
                                          scratch.acc (to be all-reduced), integrate with cosmo3_integrate */

#define COSMO3_ERR_INT_OVERFLOW 16 /* an integer mass x density product exceeded 128 bits (status ERROR bit) */

typedef struct cosmo3_state {
    int32_t cap;
    int32_t n_total;
    double *pos;      /* [3][cap] */
    double *vel;      /* [3][cap] */
    double *mass;
    uint64_t *mass_hi;
    uint64_t *mass_lo;
    double *density;
    double *radius;
    int32_t *id;
    uint8_t *flags;
    int64_t *status;  /* [COSMO_STATUS_LEN] */
    /* Object.absorb call stream: (iteration, absorber id, absorbed id, absorbed still existed) in
       the order of the reference's double loop */
    int64_t *events;  /* [event_cap][4] */
    int64_t event_cap;
} cosmo3_state;

typedef struct cosmo3_scratch {
    double *h;         /* [cap]     half diagonal of the d^2 chain */
    float *pk;         /* [4][cap]  FP32 source stage: scaled x, y, z, mass */
    double *acc;       /* [3][cap]  accelerations of the last step (COSMO3_OPT_STORE_ACC) */
    double *pos_new;   /* [3][cap]  integrated positions / velocities of all objects */
    double *vel_new;   /* [3][cap] */
    double *xx_new;    /* [cap]     squared row norms of pos_new */
    double *rinf;      /* [cap]     radii inflated by 2^-30 (candidate filter of stage B) */
    double *partial;   /* [jsplit_max][3][cap] partial sums of the all-pairs pass */
    int32_t jsplit_max;
    int32_t reserved0;
    int32_t *tile_ctr; /* [cap / 64 + 4] arrival counters, must start zeroed */
    uint64_t *pairs;        /* [pair_cap] flagged ordered pairs (i << 32 | j), unsorted */
    uint64_t *pairs_sorted; /* [pair_cap] */
    int64_t pair_cap;
    uint8_t *exists;   /* [cap]  Object.exists during the step */
    double *cur_pos;   /* [3][cap] merged position / velocity of absorbers */
    double *cur_vel;   /* [3][cap] */
    int32_t *touched;  /* [cap] step stamp of cur_pos/cur_vel, must start at -1 */
    int32_t *scan_in;  /* [cap + 1] */
    int32_t *scan_out; /* [cap + 1] */
    int32_t *scan_blk; /* [COSMO_SCAN_BLOCKS] */
    /* compaction staging */
    double *t_pos; double *t_vel; double *t_mass; uint64_t *t_mass_hi; uint64_t *t_mass_lo;
    double *t_density; double *t_radius; int32_t *t_id; uint8_t *t_flags;
    /* uniform-grid broad phase of stage B (detect_mode 1; same arrays as cosmo_scratch) */
    int32_t *cell_of;      /* [cap] */
    int32_t *cell_count;   /* [grid_cells + 1] must start zeroed (self-resetting) */
    int32_t *cell_start;   /* [grid_cells + 1] */
    int32_t *cell_items;   /* [cap] */
    int32_t *giants;       /* [giant_cap] */
    int32_t *chunk_bounds; /* [cap / 1024 + 2] */
    int32_t *heavy_rows;   /* [cap] */
    int32_t grid_cells;
    int32_t giant_cap;
    /* column partial sums of the symmetric kernel: [colbuf_rows][3][cap], one row per 512-object row block */
    double *colbuf;
    int64_t colbuf_rows;   /* FP64: >= ceil(ceil(n / 512) / ranks); FP32 (1024-object blocks, float planes): a quarter of that */
    cosmo_grid *grid_dev;  /* [1] broad-phase grid of the current step (detect_mode 1 and 2) */
    int32_t *tune_hist;    /* [COSMO_TUNE_HIST_LEN] detect_mode 2, must start zeroed (self-resetting) */
    /* parallel stage C (resolve_mode 1): row-major form of the flagged pairs and their connected components */
    int32_t *row_count;    /* [cap + 1] must start zeroed (self-resetting) */
    int32_t *row_start;    /* [cap + 1] */
    int32_t *label;        /* [cap] union-find parent / component root */
    int32_t *comp_root;    /* [pair_cap] component root of every sorted pair */
    int32_t *comp_items;   /* [pair_cap] sorted-pair indices grouped by component */
    int32_t *pair_flag;    /* [pair_cap] result of the absorb call of every sorted pair */
    int32_t *scan_a;       /* [pair_cap + 1] */
    double *corr;          /* [3][cap] near_mode != 0: FP64 near-field corrections of the FP32 sums */
} cosmo3_scratch;

typedef struct cosmo3_step_params {
    double G;        /* State.G */
    double dt;       /* State.dt */
    uint32_t opts;   /* COSMO3_OPT_* */
    int32_t n_upper; /* host-side upper bound of the object count (sizes the grids) */
    int32_t i_begin; /* target-row shard [i_begin, i_end) of this rank; 0, n_upper if single */
    int32_t i_end;
    int32_t pos_exp;  /* FP32 path: positions scaled by 2^-pos_exp, masses by 2^-mass_exp */
    int32_t mass_exp;
    int32_t jsplit;        /* source-range splits of the all-pairs pass */
    int32_t jsplit_detect; /* source-range splits of the detection pass */
    /* stage B mode: 0 = all-pairs predicate; 1 = uniform-grid broad phase on the (x, y) projection
       (objects close in 3-D are at least as close in projection, so the candidates stay a superset;
       every candidate goes through the exact 3-D predicate); 2 = the same grid tuned on the device
       every step.  Fields as in cosmo_step_params. */
    int32_t detect_mode;
    int32_t resolve_mode;  /* stage C: 0 = rank sort + one thread (few flagged pairs), 1 = row-major counting
                              sort + connected components replayed in parallel (crowded steps) */
    int32_t cells_per_side;
    int32_t sym_chunk;     /* symmetric kernel: column blocks per CTA (0 = default) */
    double cell_size;
    double grid_origin_x;
    double grid_origin_y;
    double giant_radius;
    int32_t shard_index;   /* this rank / number of ranks sharing the step (0 / 1 if single): the grid broad */
    int32_t shard_count;   /* phase and the symmetric kernel deal their chunks / row blocks round-robin      */
    int32_t near_mode;     /* FP32 path: != 0 re-evaluates pairs closer than 2^-9 of their coordinates in FP64
                              (cosmo_step_params.near_mode: 1 = bin with the grid of the previous step's stage B,
                              2 = tune a grid first) */
    int32_t reserved2;
} cosmo3_step_params;

/* Stage A: per-step preparation + all-pairs acceleration (FP64: the reference's rounding sequence
   of d^2 in three dimensions; FP32: direct differences) with the integrator fused in:
   scratch.pos_new / vel_new / xx_new of rows [i_begin, i_end). */
int cosmo3_accel_integrate(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p,
                           void *stream);
/* Stage B: appends every flagged ordered pair (i, j), i in [i_begin, i_end), to scratch.pairs
   (sklearn's euclidean_distances on (N,3) float64 restated bit for bit; no clip of d to 1). */
int cosmo3_detect_pairs(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p,
                        void *stream);
/* Stage C: row-major sort of the pairs, the ordered absorb loop with Python's int/float semantics
   (incl. the re-absorption of destroyed objects), commit of positions and velocities, stable
   removal of the absorbed objects, State.iteration += 1. */
int cosmo3_resolve_commit(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p,
                          void *stream);
/* Multi-GPU symmetric path: integrates ALL rows from the all-reduced sums in scratch.acc
   (see COSMO3_OPT_DEFER_INTEGRATE). */
int cosmo3_integrate(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p, void *stream);
/* Appends `n_rows` pair lists laid out as rows of (count | pairs[max_count]) int64 words -- the
   all-gathered exchange buffer of the multi-GPU step -- to scratch.pairs in one launch. */
/* cosmo_seal_pair_row for the 3-D step (count of this rank's pairs -> row[0], N_PAIRS reset). */
int cosmo3_seal_pair_row(const cosmo3_state *st, int64_t *row, void *stream);
int cosmo3_append_pair_rows(const cosmo3_state *st, const cosmo3_scratch *sc, const int64_t *rows, int32_t n_rows,
                            int64_t max_count, void *stream);
/* A + B + C on one device. */
int cosmo3_step(const cosmo3_state *st, const cosmo3_scratch *sc, const cosmo3_step_params *p, void *stream);
/* Suggested source-range split; kernel: 0 FP64 accel, 1 FP32 accel, 2 detection. */
int cosmo3_suggest_jsplit(int32_t rows, int32_t n, int kernel, int32_t jsplit_max);

/* Host-side copies of the exact scalar helpers (no GPU needed): correctly rounded int / int
   (Python's true division) and Object.get_radius() from an (int|float) mass and density. */
double cosmo3_host_truediv(uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo);
double cosmo3_host_radius(int mass_int, uint64_t mass_hi, uint64_t mass_lo, double mass_f, int dens_int,
                          double dens_f);

#ifdef __cplusplus
}
#endif
#endif /* COSMOSIM_B200_STATE3D_H */
