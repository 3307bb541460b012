/*
 * cosmosim_b200.h -- C ABI of the B200-native per-timestep physics update of cosmosim.
 *
 * The reference has no FFI: its seam is two Python call sites inside
 * Universe.interact() (reference cosmosim/core/universe_old.py:166-212):
 *     acc_blas(p0, m, G)                  universe_old.py:174  (cosmosim/util/blas.py:4-41)
 *     pairwise_distances(p, n_jobs=-1)    universe_old.py:180
 * plus the numpy integrator (:177-178) and the Python merge loop (:196-210) with
 * Planet.absorb/destroy (cosmosim/core/planet.py:74-91).  Each entry point below names
 * the reference lines it replaces.  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add at those call sites.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every array argument is a DEVICE pointer owned by
 *     the caller (the Python host keeps them alive as torch tensors) unless the name ends
 *     in _host; `stream` is a cudaStream_t passed as void*;
 *   - every function returns 0 on success, a negative COSMO_E_* code otherwise, never
 *     throws and never allocates device memory; cosmo_last_error() returns a
 *     thread-local description of the last failure;
 *   - launches are asynchronous on `stream`; nothing synchronises unless documented.
 *
 * Data layout in HBM (SoA, one slot per ACTIVE planet, slots ordered by insertion order of
 * the planets that are still alive -- the order of Universe.active_planets,
 * universe_old.py:63-64,83):
 *     pos, vel            double[cap][2]   (16-byte aligned, loaded as double2)
 *     mass                double[cap]      float(mass)
 *     mass_hi, mass_lo    uint64[cap]      128-bit integer mass when COSMO_F_MASS_INT
 *     radius, volume      double[cap]
 *     eaten               int64[cap]       Planet.planets_eaten
 *     id                  int32[cap]       insertion index in Universe.planets
 *     flags               uint8[cap]       COSMO_F_* bits
 * `cap` must be a multiple of COSMO_PAD and all arrays hold `cap` elements.
 */
#ifndef COSMOSIM_B200_H
#define COSMOSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COSMO_ABI_VERSION 2
#define COSMO_PAD 1024 /* slot capacity granularity (tile size of the all-pairs kernels) */
#define COSMO_MAX_PEERS 16   /* ranks of one NVLink domain */
#define COSMO_SCAN_BLOCKS 8192 /* device scans handle up to COSMO_SCAN_BLOCKS * 4096 items */

/* error codes */
#define COSMO_OK 0
#define COSMO_E_ARG (-1)       /* bad argument */
#define COSMO_E_CUDA (-2)      /* a CUDA runtime call failed, see cosmo_last_error() */
#define COSMO_E_NO_DEVICE (-3) /* no sm_100 device */

/* per-planet flag bits */
#define COSMO_F_IMMOBILE 1u /* Planet.immobile (planet.py:32) */
#define COSMO_F_MASS_INT 2u /* mass is a Python int held exactly in mass_hi:mass_lo */

/* step option bits */
#define COSMO_OPT_COLLISIONS 1u  /* simulate(collisions=True), universe_old.py:191,202 */
#define COSMO_OPT_BOUNDED 2u     /* Universe(bounded=True), universe_old.py:349 */
#define COSMO_OPT_FP32 4u        /* FP32 fast path for the all-pairs sum (state stays FP64) */
#define COSMO_OPT_STORE_ACC 8u   /* also write the accelerations to scratch.acc */
#define COSMO_OPT_F32_SYM 32u    /* FP32 path: symmetric kernel (each unordered pair once); needs colbuf */
#define COSMO_OPT_PEER_EXCHANGE 512u /* with DEFER_INTEGRATE: the reduce kernel stores this rank's partial sums
                                        straight into every peer's exchange buffer over NVLink
                                        (scratch.peer_acc), cosmo_integrate adds the ranks' slots in rank order */
#define COSMO_OPT_F64_SYM 256u    /* FP64 path: symmetric kernel (each unordered pair once); needs colbuf */
#define COSMO_OPT_ENERGY 128u     /* also evaluate energies / angular momentum (universe_old.py:181-189,210) */
#define COSMO_OPT_DEFER_INTEGRATE 64u /* symmetric kernel, multi-GPU: leave this rank's partial sums in
                                         scratch.acc (to be all-reduced), integrate with cosmo_integrate */

/* indices into the device-side int64 status block (cosmo_state.status, COSMO_STATUS_LEN) */
#define COSMO_ST_N_ACTIVE 0    /* active planets = slots [0, n) */
#define COSMO_ST_STEP 1        /* frames executed (Universe.iterations, universe_old.py:361) */
#define COSMO_ST_N_EVENTS 2    /* merge events logged so far (cumulative) */
#define COSMO_ST_N_PAIRS 3     /* flagged ordered pairs of the last frame */
#define COSMO_ST_N_DEAD 4      /* planets that died in the last frame (absorbed + escaped) */
#define COSMO_ST_ERROR 5       /* sticky COSMO_ERR_* bits */
#define COSMO_ST_N_ESCAPED 6   /* escaped planets destroyed so far (cumulative) */
#define COSMO_ST_N_CAND 7      /* broad-phase candidates of the last frame (diagnostic) */
#define COSMO_ST_N_GIANTS 8    /* planets brute-forced by the broad phase in the last frame */
#define COSMO_ST_PMAX2 9       /* bits of max |pos_new|^2 (double) of the last frame */
#define COSMO_ST_N_COMPLEX 10  /* flagged pairs replayed sequentially in the last frame */
#define COSMO_ST_SLOT_LO 11    /* grid broad phase: this rank's range of the cell-sorted order */
#define COSMO_ST_SLOT_HI 12
#define COSMO_ST_EV_BASE 13    /* event count at the start of the frame's log append */
#define COSMO_ST_N_HEAVY 14    /* grid broad phase: rows handed to the warp-per-row pass in the last frame */
#define COSMO_STATUS_LEN 16

#define COSMO_ERR_PAIR_OVERFLOW 1  /* more flagged pairs than pair_cap */
#define COSMO_ERR_EVENT_OVERFLOW 2 /* event log full */
#define COSMO_ERR_NAN 4            /* non-finite d^2 met in the acceleration pass */
#define COSMO_ERR_GIANT_OVERFLOW 8 /* more brute-forced planets than giant_cap */

/* Broad-phase grid of the current frame, resident on the device (scratch.grid_dev): written by the
   first kernels of cosmo_detect_pairs -- from the step parameters (detect_mode 1) or tuned on the
   device from the frame's own positions and radii (detect_mode 2) -- and read by the rest. */
typedef struct cosmo_grid {
    double h;        /* cell edge in use (>= the smallest admissible edge for giant radius rg) */
    double cx, cy;   /* centre of period 0 of the grid */
    double rg;       /* planets with inflated radius above this are tested against everybody */
    int64_t ncell;   /* c * c */
    int32_t c;       /* slots per side of the periodic grid (>= 3) */
    int32_t w;       /* 0 (reserved) */
    int32_t source;  /* 1 = from cosmo_step_params, 2 = tuned on the device */
    int32_t reserved;
    double box;      /* edge of the bulk box the grid was tuned for (diagnostic) */
    double h_tuned;  /* cell edge before the admissibility clamp (diagnostic) */
    double far2;     /* planets with |pos_new|^2 above this are brute-forced like giants */
    double rho;      /* planets per unit area a planet sees on average, as used by the tuner (detect_mode 2) */
    int64_t cand_total; /* broad-phase candidates summed over ALL non-giant planets of this frame (the same
                           on every rank): next frame's tuner derives the density from it */
    int64_t n_counted;  /* planets that sum runs over */
} cosmo_grid;
#define COSMO_TUNE_HIST_LEN (2 * 16384 + 4096 + 64) /* int32 words of scratch.tune_hist */

/* Persistent state.  All pointers are device pointers. */
typedef struct cosmo_state {
    int32_t cap;      /* slots allocated (multiple of COSMO_PAD) */
    int32_t n_total;  /* planets ever inserted = length of the rec_* arrays */
    double *pos;      /* [cap][2] */
    double *vel;      /* [cap][2] */
    double *mass;     /* [cap] */
    uint64_t *mass_hi;
    uint64_t *mass_lo;
    double *radius;
    double *volume;
    int64_t *eaten;
    int32_t *id;
    uint8_t *flags;
    int64_t *status;  /* [COSMO_STATUS_LEN] */
    double *energy;   /* [cap] Planet.energy = K_i + U_i of the last frame with COSMO_OPT_ENERGY */
    double *totals;   /* [4] total_energy, angular_momentum of the last frame (universe_old.py:187,189) */
    /* records of dead planets, indexed by insertion id (Planet keeps mass/radius after
       destroy(), planet.py:86-91) */
    double *rec_mass;      /* [n_total] */
    uint64_t *rec_mass_hi; /* [n_total] */
    uint64_t *rec_mass_lo; /* [n_total] */
    double *rec_radius;    /* [n_total] */
    double *rec_volume;    /* [n_total] */
    int64_t *rec_eaten;    /* [n_total] */
    int32_t *rec_death;    /* [n_total] frame of death, -1 while alive */
    uint8_t *rec_flags;    /* [n_total] */
    /* merge stream: (frame, absorber id, absorbed id) triples in visitation order */
    int64_t *events; /* [event_cap][3] */
    int64_t event_cap;
} cosmo_state;

/* Per-step scratch (device).  Contents are undefined between calls except where noted. */
typedef struct cosmo_scratch {
    double *h;         /* [cap]      half diagonal of the d^2 chain */
    float *pk_x;       /* [cap]      FP32 j-stage (scaled x, y, mass) */
    float *pk_y;       /* [cap] */
    float *pk_m;       /* [cap] */
    double *acc;       /* [cap][2]   accelerations of the last frame (COSMO_OPT_STORE_ACC) */
    double *pos_new;   /* [cap][2]   integrated positions of ALL active planets (A2) */
    double *vel_new;   /* [cap][2] */
    double *xx_new;    /* [cap]      squared row norms of pos_new (A3) */
    double *rinf;      /* [cap]      radii inflated by 2^-30 (candidate filter of stage B) */
    double *partial;   /* [jsplit_max][cap][2] partial sums of the all-pairs pass */
    int32_t jsplit_max;
    float *colbuf;        /* [colbuf_rows][cap][2] floats: column partial sums of the symmetric kernels, one
                             double2 (= two float2 rows) per planet and super row (FP32: sym_rows
                             1024-planet row blocks) / per 512-planet row block (FP64) */
    int64_t colbuf_rows;  /* FP32: >= 2 * ceil(ceil(ceil(n/1024)/ranks) / sym_rows); FP64: >= 2 * ceil(ceil(n/512)/ranks) */
    int32_t *tile_ctr; /* [cap / 64 + 4] arrival counters, must start zeroed */
    uint8_t *alive;    /* [cap]      1 while the slot's planet is alive in this frame */
    uint8_t *escaped;  /* [cap]      start-of-frame escape flags (A6) */
    uint64_t *pairs;   /* [pair_cap] flagged ordered pairs (i << 32 | j), unsorted */
    uint64_t *pairs_sorted; /* [pair_cap] */
    int64_t pair_cap;
    double *cur_pos;   /* [cap][2]   resolver: merged position/velocity of touched planets */
    double *cur_vel;   /* [cap][2] */
    int32_t *mod_row;  /* [cap]      resolver: row of last modification */
    int32_t *mod_step; /* [cap]      resolver: frame stamp of mod_row, must start at -1 */
    int32_t *chunk_alive; /* [cap / 1024 + 1] survivors per 1024-slot chunk */
    /* row-major (CSR) form of the flagged pairs + parallel resolver */
    int32_t *row_count;   /* [cap + 1] must start zeroed (self-resetting) */
    int32_t *row_start;   /* [cap + 1] */
    int32_t *pmin;        /* [cap] smallest partner, must start at INT32_MAX (self-resetting) */
    int32_t *pmax;        /* [cap] largest partner, must start at -1 (self-resetting) */
    int32_t *scan_blk;    /* [COSMO_SCAN_BLOCKS] block sums of the device scans */
    int32_t *pair_flag;   /* [pair_cap] bit0: merge happened, bit1: needs sequential replay */
    int32_t *cplx_list;   /* [pair_cap] sorted-pair indices of the clustered pairs */
    int32_t *scan_a;      /* [pair_cap + 1] marks / exclusive scan of the ordered compactions */
    int64_t *ev_tmp;      /* [pair_cap][2] (absorber id, absorbed id) per sorted-pair slot */
    int32_t *label;       /* [cap] connected-component label of clustered planets */
    int32_t *comp_items;  /* [pair_cap] sorted-pair indices grouped by component */
    int32_t *comp_root;   /* [pair_cap] component representative of every clustered pair */
    /* compaction staging: one array per persistent field */
    double *t_pos; double *t_vel; double *t_mass; uint64_t *t_mass_hi; uint64_t *t_mass_lo;
    double *t_radius; double *t_volume; int64_t *t_eaten; int32_t *t_id; uint8_t *t_flags;
    double *t_energy;
    /* broad phase (uniform grid) */
    int32_t *cell_of;    /* [cap] */
    int32_t *cell_count; /* [grid_cells + 1] must start zeroed (self-resetting) */
    int32_t *cell_start; /* [grid_cells + 1] */
    int32_t *cell_items; /* [cap] */
    int32_t *giants;     /* [giant_cap] */
    int32_t grid_cells;  /* capacity of the cell arrays */
    int32_t giant_cap;
    /* multi-GPU peer exchange (COSMO_OPT_PEER_EXCHANGE): peer_acc[r] = rank r's exchange buffer
       double[n_peers][cap][2] as mapped into THIS process (NVLink peer memory); slot s of every
       buffer receives rank s's partial sums */
    double *peer_acc[COSMO_MAX_PEERS];
    int32_t n_peers;
    int32_t reserved_scratch;
    cosmo_grid *grid_dev; /* [2] grid of the current frame's stage B (detect_mode 1 and 2); slot 1: the
                             near-field pass of the FP32 path */
    int32_t *tune_hist;   /* [COSMO_TUNE_HIST_LEN] detect_mode 2: coordinate / radius histograms, must
                             start zeroed (self-resetting) */
    double *corr;         /* [cap][2] near_mode != 0: FP64 near-field corrections of the FP32 sums */
} cosmo_scratch;

/* Parameters of one frame. */
typedef struct cosmo_step_params {
    double G;        /* Universe.G (universe_old.py:24-25) */
    double dt;       /* Universe.dt = speed/fps (universe_old.py:318-319) */
    double r_max;    /* Universe.r_max */
    uint32_t opts;   /* COSMO_OPT_* */
    int32_t n_upper; /* host-side upper bound of the active count (sizes the grids) */
    int32_t i_begin; /* target-row shard [i_begin, i_end) of this rank; 0, n_upper if single */
    int32_t i_end;
    int32_t pos_exp;  /* FP32 path: positions are scaled by 2^-pos_exp, masses by 2^-mass_exp */
    int32_t mass_exp;
    int32_t jsplit;   /* source-range splits of the all-pairs pass (>=1, <= jsplit_max) */
    int32_t detect_mode; /* 0 = all-pairs predicate, 1 = uniform-grid broad phase with the grid given
                            below, 2 = uniform grid tuned on the device every frame (cell_size, origin,
                            cells_per_side and giant_radius are ignored) */
    int32_t resolve_mode; /* 0 = rank sort + sequential replay (few pairs), 1 = CSR + parallel */
    int32_t small_tiles;  /* 1 = small-N kernel shape for stage A/B (more, shorter CTAs) */
    int32_t shard_index;  /* this rank / number of ranks sharing the frame (0 / 1 if single): the */
    int32_t shard_count;  /* grid broad phase splits its cell-sorted planet order evenly by these */
    int32_t sym_chunk;    /* symmetric kernel: column blocks per CTA (0 = default) */
    int32_t near_mode;    /* FP32 path: != 0 re-evaluates the near field (pairs closer than 2^-9 of their
                             coordinates, i.e. 2^14 FP32 ulps) in FP64; 1 = bin the planets with the grid
                             the previous frame's stage B left in scratch.grid_dev, 2 = tune a grid first
                             (first frame, or stage B not in grid mode) */
    int32_t sym_rows;     /* symmetric FP32 kernel: row blocks per CTA (1, 2 or 4; 0 = 1); see colbuf_rows --
                             the launcher raises sym_rows if colbuf has fewer rows */
    int32_t reserved1;
    double cell_size;     /* detect_mode 1: edge of the square cells of the PERIODIC grid: cell (ix, iy) of  */
    double grid_origin_x; /* the tiling anchored at the origin is stored in slot (iy mod c, ix mod c), c =   */
    double grid_origin_y; /* cells_per_side >= 3 (cells a multiple of c apart share a slot: a superset)      */
    int32_t cells_per_side;
    int32_t jsplit_detect; /* source-range splits of the stage-B all-pairs pass */
    double giant_radius; /* detect_mode 1: planets with radius above this are brute-forced */
} cosmo_step_params;

int cosmo_abi_version(void);
/* Kernels of this library launched by this process so far (bench.py reports the difference). */
long long cosmo_launch_count(void);
const char *cosmo_last_error(void);
/* Number of SMs / compute capability of the current device (host query). */
int cosmo_device_info(int *sm_count, int *cc_major, int *cc_minor);

/*
 * Stage A -- replaces acc_blas (blas.py:4-41; call site universe_old.py:174) and the
 * integrator (universe_old.py:177-178).  Reads state.{pos,vel,mass,flags} of slots
 * [0, n) as sources and integrates target rows [i_begin, i_end):
 *     scratch.pos_new/vel_new/xx_new[i] (and scratch.acc[i] with COSMO_OPT_STORE_ACC).
 * Also evaluates the start-of-frame escape flags (universe_old.py:75-76) and resets
 * scratch.alive.  FP64 path reproduces the reference's rounding sequence of d^2.
 */
int cosmo_accel_integrate(const cosmo_state *st, const cosmo_scratch *sc,
                          const cosmo_step_params *p, void *stream);

/*
 * Stage B -- replaces pairwise_distances + the collision matrix (universe_old.py:180,
 * 192-193): appends every flagged ordered pair (i, j), i in [i_begin, i_end), j in [0, n),
 * i != j, to scratch.pairs and counts them in status[COSMO_ST_N_PAIRS] (which the caller
 * resets through cosmo_begin_frame).  Needs pos_new/xx_new of ALL slots (all-gathered by
 * the caller when sharded).
 */
int cosmo_detect_pairs(const cosmo_state *st, const cosmo_scratch *sc,
                       const cosmo_step_params *p, void *stream);

/*
 * Stage C -- replaces the ordered merge loop (universe_old.py:196-210) and
 * Planet.absorb/destroy (planet.py:74-91): sorts the flagged pairs row-major, replays the
 * reference's visitation order, logs (frame, absorber id, absorbed id), commits positions
 * and velocities, destroys escaped planets when bounded (universe_old.py:349-350), then
 * stably compacts the survivors and bumps status[COSMO_ST_STEP].
 */
int cosmo_resolve_commit(const cosmo_state *st, const cosmo_scratch *sc,
                         const cosmo_step_params *p, void *stream);

/* Multi-GPU symmetric FP32 path: integrates ALL active rows from the all-reduced sums in
   scratch.acc (see COSMO_OPT_DEFER_INTEGRATE). */
int cosmo_integrate(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                    void *stream);

/*
 * Optional diagnostics of the frame -- replaces universe_old.py:181-189,210: on the NEW positions
 * and velocities and the pre-merge masses, U_i = m_i sum_{j != i} -G m_j / max(d_ij, 1),
 * K_i = m_i |v_i|^2 / 2, energy[i] = K_i + U_i, totals = {sum(K + U), sum m (x v_y - y v_x)}.
 * Runs between stages A and C (cosmo_step does it when COSMO_OPT_ENERGY is set); all rows.
 */
int cosmo_energies(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                   void *stream);

/* Zeroes the per-frame counters (N_PAIRS, N_DEAD, N_CAND).  First launch of a frame. */
int cosmo_begin_frame(const cosmo_state *st, const cosmo_scratch *sc,
                      const cosmo_step_params *p, void *stream);

/* Suggested source-range split for `rows` target rows against n sources on this device;
   kernel: 0 FP64 accel, 1 FP32 packed accel, 2 FP32 scalar accel, 3 all-pairs detection,
   4 / 5 the small-N shapes of the FP64 accel / detection kernels. */
int cosmo_suggest_jsplit(int32_t rows, int32_t n, int kernel, int32_t jsplit_max);

/* Convenience: begin_frame + A + B + C on one device (no sharding). */
int cosmo_step(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
               void *stream);

/* Appends min(count_ptr[0], max_count) pairs from `src` (one rank's list) to scratch.pairs;
   a count above max_count raises COSMO_ERR_PAIR_OVERFLOW in the status block. */
int cosmo_append_pairs(const cosmo_state *st, const cosmo_scratch *sc, const uint64_t *src,
                       const int64_t *count_ptr, int64_t max_count, void *stream);
/* Multi-GPU frame, before the all-gather of the pair lists: stage B wrote this rank's pairs
   straight into its row of the exchange buffer (scratch.pairs pointing at row + 1); this stores
   their count in row[0] and resets status[COSMO_ST_N_PAIRS] for cosmo_append_pair_rows.  One
   thread, no copy. */
int cosmo_seal_pair_row(const cosmo_state *st, int64_t *row, void *stream);
/* Same for `n_rows` lists laid out as rows of (count | pairs[max_count]) int64 words, i.e. the
   all-gathered exchange buffer of the multi-GPU frame, in one launch. */
int cosmo_append_pair_rows(const cosmo_state *st, const cosmo_scratch *sc, const int64_t *rows,
                           int32_t n_rows, int64_t max_count, void *stream);

/*
 * Function-level seam for callers that only want the reference's acc_blas replaced:
 * accelerations [n][2] from positions [n][2] and masses [n] (device pointers), FP64 parity
 * path or FP32 fast path (fp32 != 0; pos_exp/mass_exp as above).  workspace: at least
 * cosmo_accel_workspace_bytes(n) bytes of device memory, zero-initialised once.
 */
int64_t cosmo_accel_workspace_bytes(int32_t n);
int cosmo_accel(const double *pos, const double *mass, int32_t n, double G, int fp32,
                int32_t pos_exp, int32_t mass_exp, double *acc_out, void *workspace,
                void *stream);

/*
 * Whole frames from host memory -- what `self.interact()` (universe_old.py:347) binds to when the
 * planets live on the host.
 *
 * cosmo_workspace_bytes / cosmo_workspace_init: ONE device allocation (256-byte aligned) carved into
 * every array of cosmo_state and cosmo_scratch for up to n_total planets and initialised (zero, and
 * the idle values of rec_death / mod_step / pmin / pmax).  opts: pass COSMO_OPT_F32_SYM and/or
 * COSMO_OPT_F64_SYM to include the column-partial buffer of the symmetric kernels (8 B x cap^2 / 1024
 * for FP32, 16 B x cap^2 / 512 for FP64).
 * cosmo_params_default: launch parameters for n active planets on one device, the defaults the Python
 * host uses (kernel flavours and grid mode by size, source splits from the occupancy of this
 * device, power-of-two FP32 scales from pos_max = max |coordinate| and mass_max).  opts carries the
 * caller's COSMO_OPT_COLLISIONS / BOUNDED / FP32 / STORE_ACC / ENERGY; the symmetric-kernel bits are
 * decided here.
 */
int64_t cosmo_workspace_bytes(int32_t n_total, uint32_t opts);
int cosmo_workspace_init(void *dev_base, int64_t dev_bytes, int32_t n_total, uint32_t opts, cosmo_state *st,
                         cosmo_scratch *sc, void *stream);
int cosmo_params_default(int32_t n, double G, double dt, double r_max, uint32_t opts, double pos_max,
                         double mass_max, const cosmo_scratch *sc, cosmo_step_params *p);

/* One frame's host buffers (pinned memory for full-speed asynchronous copies). */
typedef struct cosmo_host_frame {
    int32_t n;       /* active planets handed over (== params.n_upper) */
    int32_t reserved;
    int64_t frame;   /* Universe.iterations before this frame (stamps the merge events) */
    /* in: SoA of the n active planets in insertion order.  mass_hi/mass_lo/eaten may be NULL (zeros) */
    const double *pos, *vel, *mass;
    const uint64_t *mass_hi, *mass_lo;
    const double *radius, *volume;
    const int64_t *eaten;
    const int32_t *id;
    const uint8_t *flags;
    /* out: the survivors, compacted, in insertion order; any pointer may be NULL (not wanted) except
       status_out.  Arrays must hold n elements; status_out[COSMO_ST_N_ACTIVE] says how many were written */
    double *pos_out, *vel_out, *mass_out;
    uint64_t *mass_hi_out, *mass_lo_out;
    double *radius_out, *volume_out;
    int64_t *eaten_out;
    int32_t *id_out;
    uint8_t *flags_out;
    int64_t *status_out;  /* [COSMO_STATUS_LEN] the device status block after the frame */
    int64_t *events_out;  /* [events_cap][3] this frame's merges (frame, absorber id, absorbed id), in the
                             reference's visitation order; status_out[COSMO_ST_N_EVENTS] of them */
    int64_t events_cap;
} cosmo_host_frame;

/*
 * The whole frame in one call: state H2D -> cosmo_step (stages A, B, C) -> status D2H -> the
 * survivors and the frame's merge events D2H -> stream synchronised.  One device; returns 0 or a
 * COSMO_E_* code (device-side COSMO_ERR_* bits are in status_out[COSMO_ST_ERROR]).
 */
int cosmo_step_host(const cosmo_state *st, const cosmo_scratch *sc, const cosmo_step_params *p,
                    const cosmo_host_frame *f, void *stream);

/* Host-buffer variant (the e2e call): copies pos/mass in, runs cosmo_accel, copies acc out
   and synchronises the stream.  *_host pointers are host memory (pinned for full speed). */
int cosmo_accel_host(const double *pos_host, const double *mass_host, int32_t n, double G,
                     int fp32, int32_t pos_exp, int32_t mass_exp, double *acc_out_host,
                     void *dev_workspace, int64_t dev_workspace_bytes, void *stream);

/*
 * Pipe micro-benchmarks (the roofline denominators must be measured).
 * kind: 0 FFMA, 1 FFMA2 (fma.rn.f32x2), 2 DFMA, 3 MUFU.RSQ, 4 DADD, 5 DMUL, 6 FMUL, 7 FADD, 8 MUFU.RSQ64H;
 * operand-bandwidth probes (CTAs of at most 256 threads): 9 FFMA2 with three changing operands,
 * 10 FFMA2 with two changing operands and a constant, 11 the pair chain of k_accel_f32_sym on registers
 * only (the ceiling of its instruction mix), 12 the same without the MUFUs, 13 / 14 like 9 / 10 for DFMA.
 * Runs `iters` iterations of 8 independent chains per thread on `blocks` x `threads`;
 * out (device, >= blocks*threads doubles) receives a value so nothing is optimised away.
 * ops_per_thread receives the instruction count per thread.
 */
int cosmo_microbench(int kind, int blocks, int threads, int iters, double *out,
                     int64_t *ops_per_thread, void *stream);

/* 1 if the library carries (and this device loaded) the post-scheduled copy of k_accel_f32_sym -- the
   same instructions with the accumulation FFMA2 of the inner loop regrouped for the operand-reuse cache
   (cosmosim_b200/sass_tune.py).  The environment variable COSMO_SYM_TUNED=0 selects the kernel as ptxas
   scheduled it; both produce the same bits. */
int cosmo_sym_tuned_available(void);

/* Host-side copies of the resolver's exact scalar helpers (no GPU needed): float(int) with
   round-half-even, Python's exact int/float `>=`, and pow(x, 1/3) as planet.py:82 rounds it. */
double cosmo_host_pow_third(double x);
double cosmo_host_u128_to_double(uint64_t hi, uint64_t lo);
int cosmo_host_mass_ge(int a_int, uint64_t a_hi, uint64_t a_lo, double a_f, int b_int, uint64_t b_hi,
                       uint64_t b_lo, double b_f);

#ifdef __cplusplus
}
#endif
#endif /* COSMOSIM_B200_H */
