// capi.cu -- error plumbing and device-attribute caches of libsoftpool_b200.
#include "spk_common.cuh"
#include <string.h>
#include <stdlib.h>

namespace spk {

static thread_local char g_err[512] = {0};
char* err_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();   // clear the sticky-free error so the next call starts clean
    return (int)e;
}

// read-only attribute caches keyed by device ordinal (benign race: every writer stores the same value)
static int g_sm[64] = {0};
static int g_optin[64] = {0};

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (g_sm[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        g_sm[dev] = v;
    }
    return g_sm[dev];
}

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

int max_optin_smem() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 227 * 1024;
    if (g_optin[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) v = 227 * 1024;
        g_optin[dev] = v;
    }
    return g_optin[dev];
}

}  // namespace spk

extern "C" int spk_abi_version(void) { return SPK_ABI_VERSION; }
extern "C" const char* spk_last_error(void) { return spk::err_buf(); }
