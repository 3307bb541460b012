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

int main(int argc, char **argv) {
    const int no_gpu = argc > 1 && strcmp(argv[1], "--no-gpu") == 0;
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
