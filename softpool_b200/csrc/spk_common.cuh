// spk_common.cuh -- shared device/host helpers of libsoftpool_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <algorithm>
#include "../../include/softpool_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsoftpool_b200 is written for sm_100a (B200) only"
#endif

namespace spk {

// ---- error plumbing (thread-local text, never printf/exit) ---------------------------------
char* err_buf();
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
#define SPK_CUDA(call)                                              \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return spk::cuda_fail(e__, #call);  \
    } while (0)
#define SPK_LAUNCH_CHECK(what)                                      \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return spk::cuda_fail(e__, what);   \
    } while (0)

int sm_count();          // SMs of the current device (cached per device)
int max_optin_smem();    // cudaDevAttrMaxSharedMemoryPerBlockOptin of the current device

// tensor-core Chamfer forward, dense form (chamfer_dense.cu): every pair on the tensor pipe
size_t chamfer_dense_workspace_bytes(int B, int n, int m);
int chamfer_dense_forward(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1,
                          float* dist2, int32_t* idx1, int32_t* idx2, float* loss /* or NULL */, void* ws,
                          size_t ws_bytes, cudaStream_t st, int B2 /* distinct xyz2 samples: B for the plain call */);
// tensor-core Chamfer forward, sorted search (chamfer_tc.cu): clouds sorted along a Hilbert curve, chunk filter on the tensor pipe
bool chamfer_tc_supported(int n, int m);      // cloud sizes the sorted tensor path handles
size_t chamfer_tc_workspace_bytes(int B, int n, int m);
int chamfer_tc_forward(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1,
                       float* dist2, int32_t* idx1, int32_t* idx2, float* loss /* or NULL */, void* ws,
                       size_t ws_bytes, cudaStream_t st);

bool pdl_enabled();      // SPK_NO_PDL=1 turns programmatic dependent launch off

// Launch with the programmatic-stream-serialization attribute (kernels call pdl_trigger()/pdl_wait()).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---- total order of float32 keys -------------------------------------------------------------
// float32 -> uint32 whose unsigned order is the order torch.sort(descending=True) uses on CPU:
// NaN greatest (all NaNs tie), -0 == +0.  Same transform as oracle/softpool_oracle.py:order_key.
__device__ __forceinline__ uint32_t order_key(float f) {
    // branch-free (selects only): keeps loads around it free to be hoisted and batched
    const uint32_t b = __float_as_uint(f);
    uint32_t m = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    m = (b == 0x80000000u) ? 0x80000000u : m;                    // -0 -> the key of +0
    m = ((b & 0x7FFFFFFFu) > 0x7F800000u) ? 0xFFFFFFFFu : m;     // NaN
    return m;
}

// ---- mbarrier + bulk async copy (TMA, 1-D form: SASS UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x4000u)        // suspend-time hint: sleep in hardware instead of spinning
        : "memory");
}
// latency-critical waits (a handful of threads on a pipeline's critical path): poll, never suspend
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk copy (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- thread-block clusters: barrier (release / acquire at cluster scope) and distributed shared memory ----------
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
__device__ __forceinline__ float ld_peer_f32(const float* own_smem, uint32_t peer_rank) {
    uint32_t ra; float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(own_smem)), "r"(peer_rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

__device__ __forceinline__ uint4 ld_peer_u4(const uint32_t* own_smem, uint32_t peer_rank) {
    uint32_t ra; uint4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(own_smem)), "r"(peer_rank));
    asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(ra) : "memory");
    return v;
}

__device__ __forceinline__ void st_peer_f4(const float4* own_smem, uint32_t peer_rank, float4 v) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(own_smem)), "r"(peer_rank));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- programmatic dependent launch (PDL): the next kernel of the stream may start its prologue while this
// one drains; pdl_wait() blocks until the previous kernel has completed and its writes are visible.
__device__ __forceinline__ void pdl_trigger() {
#ifdef SPK_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Unconditional trigger, used by sp_topk_kernel only and only AFTER its own pdl_wait() (late in the kernel, before its survivor
// sort): everything that preceded the top-k in the stream is then complete, so the next kernel (the gather) may safely read ITS
// OTHER inputs (x) in its prologue, before its own pdl_wait() -- the x tiles stream in while the top-k finishes.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Tail trigger (compile with -DSPK_PDL_TAIL; OFF by default): called by every thread once its main loop is
// done; the next kernel's CTAs then launch while this kernel drains (their pdl_wait() still waits for this
// kernel's completion and memory flush, so placement is a performance matter only).  Measured: -0.7 us per
// step before the loss was folded into the Chamfer forward, +4.3 us after (the prep kernel's trigger lets the
// tensor kernel's CTAs sit on their shared memory / TMEM early) -- so it stays off.
__device__ __forceinline__ void pdl_tail_trigger() {
#ifdef SPK_PDL_TAIL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
// Per-kernel tail triggers (SPK_TAIL_MASK, one bit per kernel): the next kernel's CTAs may launch while this kernel's last
// CTAs drain; their own pdl_wait() still waits for this kernel's completion and memory flush.
#ifndef SPK_TAIL_MASK
#define SPK_TAIL_MASK 0
#endif
template <int BIT>
__device__ __forceinline__ void pdl_tail_trigger_bit() {
    if ((SPK_TAIL_MASK >> BIT) & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// (Early griddepcontrol.launch_dependents was measured twice on B200 and is NOT used: from every kernel the
// step got slower, 107.7 vs 102.5 us -- waiting CTAs take resident slots from the persistent kernels; from
// the short single-wave kernels only (top-k, Chamfer prep, loss) 94.9 vs 90.4 us -- the tensor-core kernel's
// CTAs sit on their shared memory / TMEM while the prep pass they wait for runs slower.)

// streaming 128-bit global store / load
__device__ __forceinline__ void st_cs_f4(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace spk
