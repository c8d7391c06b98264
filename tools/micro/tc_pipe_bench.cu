// tc_pipe_bench.cu -- speed-of-light skeleton of the Chamfer tensor kernel's MMA <-> TMEM-drain loop (sm_100a):
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
constexpr uint32_t IDESC = (0u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(0u)
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
    unsigned char b_tile[4][4096];
    uint32_t cm[2][32 * 128];
    uint64_t full[8], empty[8];
    uint32_t tmem_base;
};

// NBUF accumulators of 128 columns; DW drain warps (4 or 8: with 8, the two warps of a lane quarter alternate tiles);
// PIPE: 0 = ld, wait, hand back, reduce (the round-1 order); 1 = reduce of tile i-1 between ld of tile i and its wait
template <int NBUF, int DW, int PIPE, bool SPIN, int CPS, int X>
__global__ void __launch_bounds__(128 + 32 * DW, CPS) k(long long* cyc, uint32_t* sink, int tiles) {
    extern __shared__ __align__(128) unsigned char raw[];
    Smem& S = *reinterpret_cast<Smem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (int)(sizeof(S.a_tile) + sizeof(S.b_tile)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(S.a_tile)[i] = 0x3C003C00u;
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], 4); }
        fence_mbar_init();
    }
    constexpr int COLS = NBUF * 128;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const long long t0 = clock64();
    uint32_t acc = 0;
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t a_desc = umma_smem_desc(S.a_tile);
            long long ia = 0, ib = 0;
            for (int t = 0; t < tiles && X != 1; ++t) {
                const int buf = t % NBUF;
                const long long c0 = clock64();
                if (X == 2) { if (t >= NBUF) mbar_wait(&S.full[buf], ((t / NBUF) & 1) ^ 1); }     // own commit of the MMA NBUF tiles back
                else if (SPIN) mbar_wait_spin(&S.empty[buf], ((t / NBUF) & 1) ^ 1); else mbar_wait(&S.empty[buf], ((t / NBUF) & 1) ^ 1);
                const long long c1 = clock64();
                tc_fence_after();
                umma_f16(tmem_base + buf * 128, a_desc, umma_smem_desc(S.b_tile[t & 3]));
                umma_commit(&S.full[buf]);
                const long long c2 = clock64();
                ia += c1 - c0; ib += c2 - c1;
            }
            if (X == 5 && blockIdx.x == 0) printf("  issuer: wait empty %.1f, fence+mma+commit %.1f cycles per tile\n", (double)ia / tiles, (double)ib / tiles);
        }
    } else if (warp >= 4) {
        const int q = warp & 3, set = (warp - 4) >> 2;              // set 0/1 when DW == 8
        constexpr int NSET = DW / 4;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t* out = S.cm[0] + q * 32 + lane;
        uint32_t r[64], prev[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) prev[i] = 0;
        long long da = 0, db = 0, dc = 0, dd = 0;
        for (int t = set; t < tiles; t += NSET) {
            const int buf = t % NBUF;
            const long long c0 = clock64();
            if (X != 1 && X != 2) { if (SPIN) mbar_wait_spin(&S.full[buf], (t / NBUF) & 1); else mbar_wait(&S.full[buf], (t / NBUF) & 1); }
            const long long c1 = clock64();
            tc_fence_after();
            ld32x32_x64p(lane_addr + buf * 128, r);
            if (PIPE == 1) {
                pin<64>(prev);
#pragma unroll
                for (int g = 0; g < 4; ++g) out[((t & 7) * 4 + g) * 128] = pmin16(prev + 16 * g);
            }
            ld_wait();
            const long long c2 = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0 && X != 1 && X != 2) mbar_arrive(&S.empty[buf]);
            const long long c3 = clock64();
            da += c1 - c0; db += c2 - c1; dc += c3 - c2;
            if (X == 3) { acc ^= r[0] ^ r[21] ^ r[42] ^ r[63]; continue; }
            if (PIPE == 1) {
#pragma unroll
                for (int i = 0; i < 64; ++i) prev[i] = r[i];
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) out[((t & 7) * 4 + g) * 128] = pmin16(r + 16 * g);
            }
        }
        if (PIPE == 1)
#pragma unroll
            for (int g = 0; g < 4; ++g) acc ^= pmin16(prev + 16 * g);
        dd = clock64() - t0;
        if (X == 5 && blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 8)) printf("  drain warp %d: wait full %.1f, fence+ld+wait::ld %.1f, fence+syncwarp+arrive %.1f, rest (reduce, STS, loop) %.1f cycles per own tile\n", warp,
            (double)da * NSET / tiles, (double)db * NSET / tiles, (double)dc * NSET / tiles, (double)(dd - da - db - dc) * NSET / tiles);
    }
    const long long t1 = clock64();
    sink[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 128) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(COLS));
    }
}

template <int NBUF, int DW, int PIPE, bool SPIN, int CPS, int X = 0>
static void run(long long* cyc, uint32_t* sink) {
    const int ctas_per_sm = CPS;
    const int tiles = 4000;
    const int smem = ctas_per_sm == 1 ? 120 * 1024 : 80 * 1024;
    auto kern = k<NBUF, DW, PIPE, SPIN, CPS, X>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128 + 32 * DW, smem);
    if (occ != ctas_per_sm || NBUF * 128 * ctas_per_sm > 512) { printf("NBUF=%d DW=%d PIPE=%d SPIN=%d x %d CTA/SM: skipped (occupancy %d)\n", NBUF, DW, PIPE, (int)SPIN, ctas_per_sm, occ); return; }
    for (int rep = 0; rep < 2; ++rep) {
        kern<<<148 * ctas_per_sm, 128 + 32 * DW, smem>>>(cyc, sink, tiles);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    }
    long long h[2]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_tile = (double)h[0] / tiles;
    printf("X=%d NBUF=%d drain warps=%d %s %s x %d CTA/SM: %6.1f cycles per 128x128 tile per CTA -> %6.1f per SM  (tensor pipe busy 64 -> %4.1f %%)\n", X, NBUF, DW,
           PIPE ? "reduce overlapped" : "reduce after     ", SPIN ? "spin" : "susp", ctas_per_sm, per_tile, per_tile / ctas_per_sm, 6400.0 / (per_tile / ctas_per_sm));
}

int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 2 * 8); cudaMalloc(&sink, 148 * 2 * 512 * 4);
    run<2, 4, 0, false, 2>(cyc, sink);      // round-1 configuration
    run<2, 4, 0, false, 1>(cyc, sink);
    run<2, 4, 1, false, 2>(cyc, sink);
    run<2, 4, 0, true, 2>(cyc, sink);
    run<2, 4, 1, true, 2>(cyc, sink);
    run<4, 4, 0, false, 1>(cyc, sink);
    run<4, 4, 1, false, 1>(cyc, sink);
    run<4, 4, 0, true, 1>(cyc, sink);
    run<4, 4, 1, true, 1>(cyc, sink);
    run<4, 8, 0, false, 1>(cyc, sink);
    run<4, 8, 1, false, 1>(cyc, sink);
    run<4, 8, 1, true, 1>(cyc, sink);
    printf("-- X=1: drain chain alone (no MMA, no barrier waits); X=2: MMA and drain both free-running (no handshake); X=3: handshake, no reduction\n");
    run<4, 4, 0, false, 1, 1>(cyc, sink);
    run<4, 8, 0, false, 1, 1>(cyc, sink);
    run<4, 4, 0, false, 1, 2>(cyc, sink);
    run<4, 8, 0, false, 1, 2>(cyc, sink);
    run<4, 4, 0, false, 1, 3>(cyc, sink);
    run<4, 8, 0, false, 1, 3>(cyc, sink);
    run<2, 4, 0, false, 1, 3>(cyc, sink);
    run<4, 4, 0, false, 1, 5>(cyc, sink);
    run<4, 8, 0, false, 1, 5>(cyc, sink);
    run<2, 4, 0, false, 1, 5>(cyc, sink);
    return 0;
}
