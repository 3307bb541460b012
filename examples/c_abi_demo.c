/*
 * c_abi_demo.c -- libcosmosim_b200.so from plain C (no Python, no torch): the function-level seam
 * that replaces acc_blas(pos, mas, G) (reference cosmosim/util/blas.py:4-41, call site
 * cosmosim/core/universe_old.py:174).
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_abi_demo.c \
 *       -L cosmosim_b200 -lcosmosim_b200 -L /usr/local/cuda/lib64 -lcudart -lm -o c_abi_demo
 *   LD_LIBRARY_PATH=cosmosim_b200:/usr/local/cuda/lib64 ./c_abi_demo [--no-gpu]
 *
 * --no-gpu only exercises what needs no device: ABI version, argument checking, error strings and
 * the exact scalar helpers.  Without it, 4097 planets around a star are sent through
 * cosmo_accel_host (host buffers in, accelerations out) and compared with a direct double-precision
 * sum in C (the reference's expanded d^2 costs ~9 digits at AU scale, hence the 1e-6 bound).
 *
 * --step FILE runs whole frames from C -- the seam of `self.interact()` (universe_old.py:347): one
 * device allocation carved by cosmo_workspace_init, launch parameters from cosmo_params_default, and
 * three frames of a crowded 2049-planet disk through cosmo_step_host (pinned host buffers in,
 * survivors and merge events out, each frame fed by the previous one's output).  The initial state,
 * the final state and the merge stream go to FILE; tests/test_c_abi_demo.py replays the same frames
 * with the CPU oracle and compares them.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "cosmosim_b200.h"
#include "cosmosim_b200_state3d.h"

static int fail(const char *what) {
    fprintf(stderr, "FAILED: %s (%s)\n", what, cosmo_last_error());
    return 1;
}

#define STEP_N 2049
#define STEP_FRAMES 3

static int put(FILE *fp, const void *p, size_t bytes) { return fwrite(p, 1, bytes, fp) == bytes ? 0 : 1; }

/* whole frames from host buffers; returns 0 on success */
static int run_frames(const char *path) {
    const int n0 = STEP_N;
    const double G = 6.674e-11, AU = 1.496e11, ME = 5.972e24, MS = 1.989e30, PI = 3.14159265358979323846;
    const double dt = 1e6 / 33, r_max = 100 * AU;
    /* pinned host buffers: two sets, ping-pong between frames */
    double *pos[2], *vel[2], *mass[2], *radius[2], *volume[2];
    uint64_t *mhi[2], *mlo[2];
    int64_t *eaten[2], *status, *events;
    int32_t *id[2];
    uint8_t *flags[2];
    for (int b = 0; b < 2; ++b) {
        if (cudaMallocHost((void **)&pos[b], 16 * n0) || cudaMallocHost((void **)&vel[b], 16 * n0) ||
            cudaMallocHost((void **)&mass[b], 8 * n0) || cudaMallocHost((void **)&radius[b], 8 * n0) ||
            cudaMallocHost((void **)&volume[b], 8 * n0) || cudaMallocHost((void **)&mhi[b], 8 * n0) ||
            cudaMallocHost((void **)&mlo[b], 8 * n0) || cudaMallocHost((void **)&eaten[b], 8 * n0) ||
            cudaMallocHost((void **)&id[b], 4 * n0) || cudaMallocHost((void **)&flags[b], n0))
            return fail("cudaMallocHost");
    }
    if (cudaMallocHost((void **)&status, 8 * COSMO_STATUS_LEN) || cudaMallocHost((void **)&events, 24 * n0))
        return fail("cudaMallocHost");
    /* star + crowded disk on a golden-angle spiral, circular orbits, radii inflated so that frames merge */
    for (int k = 0; k < n0; ++k) {
        const double r = k ? (0.05 + 0.95 * (double)k / n0) * AU : 0.0, th = 2.399963229728653 * k;
        pos[0][2 * k] = r * cos(th); pos[0][2 * k + 1] = r * sin(th);
        const double v = k ? sqrt(G * MS / r) : 0.0;
        vel[0][2 * k] = -v * sin(th); vel[0][2 * k + 1] = v * cos(th);
        mass[0][k] = k ? (0.1 + (k % 97)) * ME : MS;
        radius[0][k] = k ? 1.5e9 * (0.5 + 0.2 * (k % 5)) : 6.9634e8;
        volume[0][k] = (4.0 / 3.0) * PI * radius[0][k] * radius[0][k] * radius[0][k];
        mhi[0][k] = mlo[0][k] = 0; eaten[0][k] = 0; id[0][k] = k;
        flags[0][k] = k ? 0 : COSMO_F_IMMOBILE;
    }
    FILE *fp = fopen(path, "wb");
    if (!fp) return fail("cannot open the output file");
    const int64_t head[4] = {n0, STEP_FRAMES, 0, 0};
    int bad = put(fp, head, sizeof(head)) | put(fp, pos[0], 16 * n0) | put(fp, vel[0], 16 * n0) | put(fp, mass[0], 8 * n0) |
              put(fp, radius[0], 8 * n0) | put(fp, volume[0], 8 * n0) | put(fp, flags[0], n0);
    /* device side: one allocation, carved and initialised by the library */
    const long long bytes = cosmo_workspace_bytes(n0, 0);
    void *ws = NULL;
    cosmo_state st;
    cosmo_scratch sc;
    if (cudaMalloc(&ws, (size_t)bytes) != cudaSuccess) return fail("cudaMalloc");
    if (cosmo_workspace_init(ws, bytes, n0, 0, &st, &sc, NULL) != COSMO_OK) return fail("cosmo_workspace_init");
    int n = n0;
    long long total_events = 0;
    for (int f = 0; f < STEP_FRAMES; ++f) {
        const int a = f & 1, b = a ^ 1;
        cosmo_step_params p;
        if (cosmo_params_default(n, G, dt, r_max, COSMO_OPT_COLLISIONS | COSMO_OPT_BOUNDED, AU, MS, &sc, &p) != COSMO_OK)
            return fail("cosmo_params_default");
        cosmo_host_frame hf;
        memset(&hf, 0, sizeof(hf));
        hf.n = n; hf.frame = f;
        hf.pos = pos[a]; hf.vel = vel[a]; hf.mass = mass[a]; hf.mass_hi = mhi[a]; hf.mass_lo = mlo[a];
        hf.radius = radius[a]; hf.volume = volume[a]; hf.eaten = eaten[a]; hf.id = id[a]; hf.flags = flags[a];
        hf.pos_out = pos[b]; hf.vel_out = vel[b]; hf.mass_out = mass[b]; hf.mass_hi_out = mhi[b]; hf.mass_lo_out = mlo[b];
        hf.radius_out = radius[b]; hf.volume_out = volume[b]; hf.eaten_out = eaten[b]; hf.id_out = id[b];
        hf.flags_out = flags[b];
        hf.status_out = status; hf.events_out = events; hf.events_cap = n0;
        if (cosmo_step_host(&st, &sc, &p, &hf, NULL) != COSMO_OK) return fail("cosmo_step_host");
        if (status[COSMO_ST_ERROR]) return fail("device error bits set");
        const long long ne = status[COSMO_ST_N_EVENTS];
        printf("frame %d: %d planets in, %lld out, %lld merges, %lld flagged pairs\n", f, n,
               (long long)status[COSMO_ST_N_ACTIVE], ne, (long long)status[COSMO_ST_N_PAIRS]);
        bad |= put(fp, &ne, 8) | put(fp, events, (size_t)ne * 24);
        total_events += ne;
        n = (int)status[COSMO_ST_N_ACTIVE];
    }
    const int last = STEP_FRAMES & 1;
    const int64_t tail[2] = {n, total_events};
    bad |= put(fp, tail, sizeof(tail)) | put(fp, id[last], 4 * (size_t)n) | put(fp, pos[last], 16 * (size_t)n) |
           put(fp, vel[last], 16 * (size_t)n) | put(fp, mass[last], 8 * (size_t)n) | put(fp, radius[last], 8 * (size_t)n) |
           put(fp, eaten[last], 8 * (size_t)n);
    if (fclose(fp) || bad) return fail("writing the output file");
    cudaFree(ws);
    if (total_events < 10) return fail("the crowded disk should merge");
    printf("cosmo_step_host: %d frames, %d -> %d planets, %lld merges\nOK\n", STEP_FRAMES, n0, n, total_events);
    return 0;
}

int main(int argc, char **argv) {
    const int no_gpu = argc > 1 && strcmp(argv[1], "--no-gpu") == 0;
    if (argc > 2 && strcmp(argv[1], "--step") == 0) return run_frames(argv[2]);
    printf("abi version %d, struct sizes: state %zu scratch %zu params %zu | 3-D: %zu %zu %zu\n", cosmo_abi_version(),
           sizeof(cosmo_state), sizeof(cosmo_scratch), sizeof(cosmo_step_params), sizeof(cosmo3_state),
           sizeof(cosmo3_scratch), sizeof(cosmo3_step_params));
    if (cosmo_accel(NULL, NULL, 0, 1.0, 0, 0, 0, NULL, NULL, NULL) != COSMO_E_ARG) return fail("argument check");
    if (!strstr(cosmo_last_error(), "bad argument")) return fail("error string");
    if (cosmo3_step(NULL, NULL, NULL, NULL) != COSMO_E_ARG) return fail("3-D argument check");
    if (cosmo_host_u128_to_double(1, 0) != 18446744073709551616.0) return fail("u128_to_double");
    if (cosmo3_host_truediv(0, 10, 0, 4) != 2.5) return fail("truediv");
    if (fabs(cosmo_host_pow_third(27.0) - 3.0) > 1e-15) return fail("pow_third");
    if (no_gpu) {
        printf("OK (no device used)\n");
        return 0;
    }

    const int n = 4097;
    const double G = 6.674e-11, AU = 1.496e11, ME = 5.972e24, MS = 1.989e30;
    double *pos = (double *)malloc(sizeof(double) * 2 * n), *mass = (double *)malloc(sizeof(double) * n);
    double *acc = (double *)malloc(sizeof(double) * 2 * n);
    pos[0] = pos[1] = 0.0;
    mass[0] = MS;
    for (int k = 1; k < n; ++k) {
        const double r = (0.05 + 0.95 * (double)k / n) * AU, th = 2.399963229728653 * k;   /* golden-angle spiral */
        pos[2 * k] = r * cos(th);
        pos[2 * k + 1] = r * sin(th);
        mass[k] = (0.1 + (k % 97)) * ME;
    }
    const long long ws_bytes = cosmo_accel_workspace_bytes(n) + ((long long)(n + COSMO_PAD) / COSMO_PAD * COSMO_PAD) * 40;
    void *ws = NULL;
    if (cudaMalloc(&ws, (size_t)ws_bytes) != cudaSuccess || cudaMemset(ws, 0, (size_t)ws_bytes) != cudaSuccess) {
        fprintf(stderr, "FAILED: cudaMalloc/cudaMemset of %lld bytes\n", ws_bytes);
        return 1;
    }
    if (cosmo_accel_host(pos, mass, n, G, 0, 0, 0, acc, ws, ws_bytes, NULL) != COSMO_OK) return fail("cosmo_accel_host");

    double worst = 0.0;
    for (int i = 0; i < n; i += 64) {     /* direct sum for a sample of rows */
        double ax = 0.0, ay = 0.0;
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double dx = pos[2 * j] - pos[2 * i], dy = pos[2 * j + 1] - pos[2 * i + 1];
            const double d2 = dx * dx + dy * dy, inv = 1.0 / (d2 * sqrt(d2));
            ax += G * mass[j] * dx * inv;
            ay += G * mass[j] * dy * inv;
        }
        const double err = hypot(acc[2 * i] - ax, acc[2 * i + 1] - ay) / hypot(ax, ay);
        if (err > worst) worst = err;
    }
    printf("cosmo_accel_host: %d planets, kernels launched so far %lld, worst relative deviation from the direct sum %.2e\n",
           n, cosmo_launch_count(), worst);
    cudaFree(ws);
    free(pos); free(mass); free(acc);
    if (!(worst < 1e-6)) return fail("accuracy");
    printf("OK\n");
    return 0;
}
