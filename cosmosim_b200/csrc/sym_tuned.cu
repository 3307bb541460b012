// Post-scheduled copies of the symmetric all-pairs kernels (cosmosim_b200/sass_tune.py): the same
// instructions as the kernels compiled into this library, the accumulation FFMA2 / DFMA of the inner loops
// regrouped so that they take operands from the reuse cache.  The patched cubins are embedded at build time
// (build/sym_tuned_cubin.inc; empty when the tuner declined) and loaded through the runtime's library API
// on first use.  The kernels in accel_sym.cu / accel_sym64.cu stay the fallback, and COSMO_SYM_TUNED=0
// selects them (tests compare the two bit for bit).
#include <stdlib.h>
#include <string.h>

#include "cosmo_internal.h"

namespace cosmo {

// generated: kTunedCubins[] = {data, size}, kTunedKernels[] = {mangled name, cubin index}
#include "sym_tuned_cubin.inc"

constexpr int kMaxTuned = 16;
static cudaKernel_t g_kernels[kMaxTuned];
static int g_smem[kMaxTuned];    // dynamic shared memory the kernel has been opted in for
static int g_state = 0;          // 0 = not tried, 1 = loaded, -1 = unavailable

static void load_once() {
    if (g_state != 0) return;
    g_state = -1;
    const int n_cubins = (int)(sizeof(kTunedCubins) / sizeof(kTunedCubins[0]));
    const int n_kernels = (int)(sizeof(kTunedKernels) / sizeof(kTunedKernels[0]));
    if (n_kernels > kMaxTuned) return;
    cudaLibrary_t libs[8];
    if (n_cubins > 8) return;
    for (int c = 0; c < n_cubins; ++c) {
        libs[c] = nullptr;
        if (kTunedCubins[c].size == 0) continue;
        if (cudaLibraryLoadData(&libs[c], kTunedCubins[c].data, nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
            cudaGetLastError();
            libs[c] = nullptr;
        }
    }
    int found = 0;
    for (int k = 0; k < n_kernels; ++k) {
        g_kernels[k] = nullptr;
        g_smem[k] = 0;
        cudaLibrary_t lib = libs[kTunedKernels[k].cubin];
        if (!lib) continue;
        if (cudaLibraryGetKernel(&g_kernels[k], lib, kTunedKernels[k].name) != cudaSuccess) {
            cudaGetLastError();
            g_kernels[k] = nullptr;
        } else {
            ++found;
        }
    }
    if (found > 0) g_state = 1;
}

// the tuned copy of the kernel with this mangled name, opted in for `dyn_smem` bytes of dynamic shared
// memory -- or nullptr (not built, not loadable, switched off)
const void *tuned_kernel(const char *mangled, int dyn_smem) {
    const char *e = getenv("COSMO_SYM_TUNED");
    if (e && e[0] == '0') return nullptr;
    load_once();
    if (g_state != 1) return nullptr;
    const int n_kernels = (int)(sizeof(kTunedKernels) / sizeof(kTunedKernels[0]));
    for (int k = 0; k < n_kernels; ++k) {
        if (!g_kernels[k] || strcmp(kTunedKernels[k].name, mangled) != 0) continue;
        if (dyn_smem > 48 * 1024 && dyn_smem > g_smem[k]) {
            if (cudaFuncSetAttribute((const void *)g_kernels[k], cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem) != cudaSuccess) {
                cudaGetLastError();
                g_kernels[k] = nullptr;      // cannot be configured: the built-in kernel serves
                return nullptr;
            }
            g_smem[k] = dyn_smem;
        }
        return (const void *)g_kernels[k];
    }
    return nullptr;
}

// number of tuned kernels this device loaded
int sym_tuned_available() {
    load_once();
    if (g_state != 1) return 0;
    int n = 0;
    const int n_kernels = (int)(sizeof(kTunedKernels) / sizeof(kTunedKernels[0]));
    for (int k = 0; k < n_kernels; ++k) n += g_kernels[k] != nullptr;
    return n;
}

}  // namespace cosmo
