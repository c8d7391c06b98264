// tc_mma_bench.cu -- tcgen05.mma issue/execution rate alone (no drain) for the K=16 distance tile: speed-of-light skeleton of the Chamfer tensor kernel's MMA <-> TMEM-drain loop (sm_100a):
// one thread issues tcgen05.mma (M=128, N=128, K=16, kind::f16, fp16 accumulate) into NBUF accumulator buffers,
// draining warps read each accumulator (tcgen05.ld 32x32b.x64.pack::16b), hand it back and reduce it to packed
// chunk minima (VIMNMX3.U16x2) -- nothing else (operands stay in shared memory, no TMA, no exact pass).
// Answers: how many cycles per 128x128 tile can the loop sustain, by buffers, drain warps, CTAs per SM and
// whether the reduction of tile i overlaps the transfer of tile i+1.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../softpool_b200/csrc -o tc_pipe_bench tc_pipe_bench.cu
#include <cstdio>
#include <cstdlib>
#include "spk_common.cuh"
#include "tmem_ld_gen.h"

namespace spk {
char* err_buf() { static char b[8]; return b; }
int fail(int c, const char*, ...) { return c; }
int cuda_fail(cudaError_t, const char*) { return 1; }
bool pdl_enabled() { return false; }
}
using namespace spk;

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_smem_desc(const void* smem_ptr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) >> 4) & 0x3FFF);
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((256u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
template <int N, int DF32 = 0> __host__ __device__ constexpr uint32_t idesc() { return ((uint32_t)DF32 << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
template <int N, int DF32>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc<N, DF32>()), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pmin16(const uint32_t* w) {
    uint32_t m0 = __vimin3_u16x2(w[0], w[1], w[2]), m1 = __vimin3_u16x2(w[3], w[4], w[5]);
    m0 = __vimin3_u16x2(m0, w[6], w[7]); m1 = __vimin3_u16x2(m1, w[8], w[9]);
    m0 = __vimin3_u16x2(m0, w[10], w[11]); m1 = __vimin3_u16x2(m1, w[12], w[13]);
    return __vimin3_u16x2(m0, m1, __vminu2(w[14], w[15]));
}
template <int N>
__device__ __forceinline__ void pin(uint32_t* v) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(v[i]));
}



struct Smem {
    unsigned char a_tile[4096];
    unsigned char b_tile[4][8192];
    uint64_t done[8];
    uint32_t tmem_base;
};
// MMAs of N columns issued back to back into rotating accumulators (512 / N of them), one commit every CE MMAs,
// the issuer waits for the commit 2 groups back.  DF32: fp32 accumulators instead of fp16.  ROT: B tile rotates over 4 slots.
template <int N, int CE, int DF32, bool ROT>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int mmas) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& S = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (4096 + 4 * 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(S.a_tile)[i] = 0x3C003C00u;
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&S.done[i], 1); fence_mbar_init(); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const long long t0 = clock64();
    if (warp == 0 && lane == 0) {
        const uint64_t a0 = umma_smem_desc(S.a_tile);
        constexpr int NB = 512 / N;
        const int groups = mmas / CE;
        for (int g = 0; g < groups; ++g) {
            if (g >= 2) mbar_wait(&S.done[(g - 2) & 7], ((g - 2) >> 3) & 1);
#pragma unroll
            for (int j = 0; j < CE; ++j) {
                const int t = g * CE + j;
                umma_f16<N, DF32>(tmem_base + (t % NB) * N, a0, umma_smem_desc(S.b_tile[ROT ? (t & 3) : 0]));
            }
            umma_commit(&S.done[g & 7]);
        }
        mbar_wait(&S.done[(groups - 1) & 7], ((groups - 1) >> 3) & 1);
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}
template <int N, int CE, int DF32, bool ROT>
static void run(long long* cyc) {
    const int mmas = 8192;
    auto kern = k<N, CE, DF32, ROT>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        kern<<<148, 128, 120 * 1024>>>(cyc, mmas);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long h[2]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = (double)h[0] / mmas;
    printf("M=128 N=%3d K=16 %s accum, commit every %d, B %s: %6.1f cycles per MMA = %6.1f per 128 columns (ideal %d)\n", N, DF32 ? "fp32" : "fp16", CE, ROT ? "rotating" : "fixed   ",
           per, per * 128 / N, 64);
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    run<64, 4, 0, true>(cyc); run<128, 4, 0, true>(cyc); run<256, 4, 0, true>(cyc);
    run<128, 1, 0, true>(cyc); run<256, 1, 0, true>(cyc); run<128, 2, 0, true>(cyc); run<256, 2, 0, true>(cyc);
    run<128, 4, 1, true>(cyc); run<256, 4, 1, true>(cyc);
    run<128, 4, 0, false>(cyc); run<256, 4, 0, false>(cyc);
    return 0;
}
