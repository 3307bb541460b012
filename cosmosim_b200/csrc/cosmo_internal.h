// Internal declarations shared by the translation units of libcosmosim_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cosmosim_b200.h"

namespace cosmo {

// error plumbing (abi.cu)
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int check_launch(const char *what);   // counts the kernel launch, then checks cudaGetLastError()

#define COSMO_TRY(expr)                                  \
    do {                                                 \
        int _rc = (expr);                                \
        if (_rc != COSMO_OK) return _rc;                 \
    } while (0)
#define COSMO_LAUNCH_CHECK(what) ::cosmo::check_launch(what)

// accel.cu
int launch_prep(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                cudaStream_t s);
int launch_accel(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                 cudaStream_t s);

int accel_residency(int which, int *rows_per_cta, int *tile);

// accel_sym.cu
int launch_accel_sym(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                     cudaStream_t s);
int launch_accel_sym64(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                       cudaStream_t s);
int launch_sym_integrate(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                         cudaStream_t s);

// nearfield.cu: FP64 near-field corrections of the FP32 sums -> scratch.corr (before launch_accel)
int launch_nearfield(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                     cudaStream_t s);

// the same for the 3-D step, on shadow structs (state3d.cu)
int launch_nearfield3(const cosmo_state &st2, const cosmo_scratch &sc2, const cosmo_step_params &p2, const float *pk,
                      const double *mass, double *corr3, int sym, cudaStream_t s);

// sym_tuned.cu: the post-scheduled copy of a kernel by mangled name (nullptr: use the one compiled into
// the library); number of tuned kernels loaded
const void *tuned_kernel(const char *mangled, int dyn_smem);
int sym_tuned_available();

// detect.cu / resolve.cu
int detect_residency(int *rows_per_cta, int *tile);
int launch_detect(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                  cudaStream_t s);
// uniform-grid broad phase; planes3: sc.pos_new holds three planes double[3][cap] and the exact test is
// the 3-D predicate of state3d.cu (st/sc are then shadow structs filled by that file)
int launch_detect_grid(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p, cudaStream_t s,
                       bool planes3);
int launch_resolve_commit(const cosmo_state &st, const cosmo_scratch &sc,
                          const cosmo_step_params &p, cudaStream_t s);
int launch_append_pairs(const cosmo_state &st, const cosmo_scratch &sc, const uint64_t *src,
                        const int64_t *count_ptr, int64_t max_count, cudaStream_t s);

int launch_seal_pair_row(int64_t *status, int64_t *row, cudaStream_t s);
int launch_append_pair_rows(const cosmo_state &st, const cosmo_scratch &sc, const int64_t *rows, int n_rows,
                            long long max_count, cudaStream_t s);

// scan.cu: out[k] = sum(in[0..k)), out[n] = total; n = min(*n_ptr + n_add, n_max) or n_max
int launch_scan_exclusive(const int *in, int *out, int *blk, const int64_t *n_ptr, long long n_max,
                          long long n_add, cudaStream_t s);

// energy.cu
int launch_energies(const cosmo_state &st, const cosmo_scratch &sc, const cosmo_step_params &p,
                    cudaStream_t s);

// microbench.cu
int launch_microbench(int kind, int blocks, int threads, int iters, double *out,
                      int64_t *ops_per_thread, cudaStream_t s);

}  // namespace cosmo
