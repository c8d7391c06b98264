// tc_pipe2_bench.cu -- second skeleton (operands streamed by TMA, wide MMAs): speed-of-light skeleton of the Chamfer tensor kernel's MMA <-> TMEM-drain loop (sm_100a):
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
template <int N> __host__ __device__ constexpr uint32_t idesc() { return (0u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
template <int N>
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc<N>()), "r"(0u)
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


constexpr int STAGES = 4;
struct Smem {
    unsigned char a_tile[2][4096];
    unsigned char b_tile[STAGES][8192];
    uint32_t cm[2][32 * 128];
    uint64_t full[4], empty[4], rfull[STAGES], rempty[STAGES];
    uint32_t tmem_base;
    long long tl[8][8];
};

// NM: accumulator columns per MMA instruction (128 or 256), NBUF accumulator buffers of NM columns,
// AP: A tiles per B tile (1, or 2 = two query tiles share every operand tile: 2 MMAs of NM=128 per handshake, accumulators paired)
// 8 drain warps: set s = (warp - 4) / 4 drains columns [128 s, 128 s + 128) of every buffer (NM = 256) or accumulator s of the pair (AP = 2);
// with NM = 128 and AP = 1 the two sets alternate buffers.
template <int NM, int NBUF, int AP, bool SPIN, bool TL, int V>
__global__ void __launch_bounds__(384, 1) k(long long* cyc, uint32_t* sink, const unsigned char* gB, int steps) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& S = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int COLS_PER_STEP = NM * AP;                 // accumulator columns produced per handshake
    constexpr int B_BYTES = NM * 32;                       // operand bytes per step
    for (int i = tid; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(S.a_tile)[i] = 0x3C003C00u;
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], ((NM == 128 && AP == 1) ? 4 : 8) * ((V & 4) ? 32 : 1)); }
        for (int i = 0; i < STAGES; ++i) { mbar_init(&S.rfull[i], 1); mbar_init(&S.rempty[i], 1); }
        fence_mbar_init();
    }
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
    uint32_t acc = 0;
    if (warp == 2) {
        if (lane == 0)
            for (int t = 0; t < steps; ++t) {
                const int s = t % STAGES;
                mbar_wait(&S.rempty[s], ((t / STAGES) & 1) ^ 1);
                mbar_expect_tx(&S.rfull[s], B_BYTES);
                bulk_g2s(S.b_tile[s], gB + (size_t)((t * 37 + blockIdx.x * 11) & 7) * 8192, B_BYTES, &S.rfull[s]);
            }
    } else if (warp == 0) {
        if (lane == 0) {
            const uint64_t a0 = umma_smem_desc(S.a_tile[0]), a1 = umma_smem_desc(S.a_tile[1]);
            for (int t = 0; t < steps; ++t) {
                const int buf = t % NBUF, s = t % STAGES;
                mbar_wait(&S.rfull[s], (t / STAGES) & 1);
                if (SPIN) mbar_wait_spin(&S.empty[buf], ((t / NBUF) & 1) ^ 1); else mbar_wait(&S.empty[buf], ((t / NBUF) & 1) ^ 1);
                if (TL && t >= 100 && t < 104) S.tl[t - 100][0] = clock64() - t0;
                tc_fence_after();
                const uint64_t bd = umma_smem_desc(S.b_tile[s]);
                umma_f16<NM>(tmem_base + buf * COLS_PER_STEP, a0, bd);
                if (AP == 2) umma_f16<NM>(tmem_base + buf * COLS_PER_STEP + NM, a1, bd);
                if (!(V & 1)) umma_commit(&S.rempty[s]);
                umma_commit(&S.full[buf]);
                if (TL && t >= 100 && t < 104) S.tl[t - 100][1] = clock64() - t0;
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, set = (warp - 4) >> 2;
        constexpr bool ALT = (NM == 128 && AP == 1);           // the sets alternate buffers
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (ALT ? 0 : set * 128);
        uint32_t* out = S.cm[set] + q * 32 + lane;
        uint32_t r[64];
        for (int t = ALT ? set : 0; t < steps; t += ALT ? 2 : 1) {
            const int buf = t % NBUF;
            if (SPIN) mbar_wait_spin(&S.full[buf], (t / NBUF) & 1); else mbar_wait(&S.full[buf], (t / NBUF) & 1);
            if (TL && warp == 4 && lane == 0 && t >= 100 && t < 104) S.tl[t - 100][2] = clock64() - t0;
            tc_fence_after();
            ld32x32_x64p(lane_addr + buf * COLS_PER_STEP, r);
            ld_wait();
            if (TL && warp == 4 && lane == 0 && t >= 100 && t < 104) S.tl[t - 100][3] = clock64() - t0;
            if (!(V & 2)) tc_fence_before();
            if (V & 4) mbar_arrive(&S.empty[buf]);
            else { __syncwarp(); if (lane == 0) mbar_arrive(&S.empty[buf]); }
            if ((V & 1) && (ALT ? (warp & 3) == 0 : warp == 4) && lane == 0) mbar_arrive(&S.rempty[t % STAGES]);      // the MMA that read this ring slot has completed
            if (TL && warp == 4 && lane == 0 && t >= 100 && t < 104) S.tl[t - 100][4] = clock64() - t0;
#pragma unroll
            for (int g = 0; g < 4; ++g) out[((t & 7) * 4 + g) * 128] = pmin16(r + 16 * g);
            if (TL && warp == 4 && lane == 0 && t >= 100 && t < 104) S.tl[t - 100][5] = clock64() - t0;
        }
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 128) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (TL && tid == 0 && blockIdx.x == 0)
        for (int i = 0; i < 4; ++i)
            printf("   step %d: issuer saw empty %lld, mma+commit issued %lld | drain warp: saw full %lld, ld done %lld, arrived %lld, reduced %lld\n", 100 + i,
                   S.tl[i][0], S.tl[i][1], S.tl[i][2], S.tl[i][3], S.tl[i][4], S.tl[i][5]);
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    }
}

template <int NM, int NBUF, int AP, bool SPIN, bool TL = false, int V = 0>
static void run(long long* cyc, uint32_t* sink, const unsigned char* gB) {
    static_assert(NM * AP * NBUF <= 512, "TMEM");
    const int tiles = 8192;                                   // 128x128 tile equivalents per CTA
    const int steps = tiles * 128 / (NM * AP);
    const int smem = 120 * 1024;
    auto kern = k<NM, NBUF, AP, SPIN, TL, V>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < (TL ? 1 : 2); ++rep) {
        kern<<<148, 384, smem>>>(cyc, sink, gB, steps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long h[2]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_tile = (double)h[0] / tiles;
    printf("V=%d MMA N=%d x %d A tile(s), %d buffers of %d columns, %s: %6.1f cycles per 128x128 tile (tensor pipe busy 64 -> %4.1f %%)\n", V, NM, AP, NBUF, NM * AP,
           SPIN ? "spin" : "susp", per_tile, 6400.0 / per_tile);
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    long long* cyc; uint32_t* sink; unsigned char* gB;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 512 * 4); cudaMalloc(&gB, 8 * 8192); cudaMemset(gB, 0x3C, 8 * 8192);
    run<256, 2, 1, false, false, 0>(cyc, sink, gB);
    run<256, 2, 1, false, false, 1>(cyc, sink, gB);
    run<256, 2, 1, false, false, 2>(cyc, sink, gB);
    run<256, 2, 1, false, false, 3>(cyc, sink, gB);
    run<256, 2, 1, false, false, 4>(cyc, sink, gB);
    run<256, 2, 1, false, false, 7>(cyc, sink, gB);
    run<256, 2, 1, true, false, 7>(cyc, sink, gB);
    run<128, 4, 1, false, false, 7>(cyc, sink, gB);
    run<128, 2, 2, false, false, 7>(cyc, sink, gB);
    run<256, 2, 1, false, true, 7>(cyc, sink, gB);
    return 0;
}
