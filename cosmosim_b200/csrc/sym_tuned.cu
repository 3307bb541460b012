// The post-scheduled copy of k_accel_f32_sym<R> (cosmosim_b200/sass_tune.py): the same instructions as the
// kernel compiled into this library, the accumulation FFMA2 of the inner loop regrouped so that they take
// operands from the reuse cache.  The patched cubin is embedded at build time (build/sym_tuned_cubin.inc;
// empty when the tuner declined) and loaded through the runtime's library API on first use.  The kernel in
// accel_sym.cu stays the fallback, and COSMO_SYM_TUNED=0 selects it (tests compare the two bit for bit).
#include <stdio.h>
#include <stdlib.h>

#include "cosmo_internal.h"

namespace cosmo {

#include "sym_tuned_cubin.inc"   // static const unsigned char kSymTunedCubin[]; static const size_t kSymTunedCubinSize;

static cudaKernel_t g_kernels[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static int g_state = 0;          // 0 = not tried, 1 = loaded, -1 = unavailable

static void load_once() {
    if (g_state != 0) return;
    g_state = -1;
    if (kSymTunedCubinSize == 0) return;
    cudaLibrary_t lib;
    if (cudaLibraryLoadData(&lib, kSymTunedCubin, nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (int r = 1; r <= 4; r *= 2) {
        char name[128];
        snprintf(name, sizeof name, "_ZN5cosmo15k_accel_f32_symILi%dEEEv11cosmo_state13cosmo_scratchiii", r);
        if (cudaLibraryGetKernel(&g_kernels[r], lib, name) != cudaSuccess) {
            cudaGetLastError();
            return;
        }
    }
    g_state = 1;
}

// the tuned kernel for R row blocks per CTA, or nullptr (not built, not loadable, or switched off)
const void *sym_tuned_kernel(int R) {
    const char *e = getenv("COSMO_SYM_TUNED");
    if (e && e[0] == '0') return nullptr;
    load_once();
    return g_state == 1 ? (const void *)g_kernels[R] : nullptr;
}

int sym_tuned_available() {
    load_once();
    return g_state == 1;
}

}  // namespace cosmo
