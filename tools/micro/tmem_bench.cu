// tmem_bench.cu -- what draining a 128-lane x 128-column accumulator from TMEM costs on sm_100a, by
// tcgen05.ld shape, warps per lane quarter and CTAs per SM; and what the alternatives to a second drain cost
// (warp-shuffle butterfly that turns 64 packed registers per lane into per-16-lane column minima).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu && ./tmem_bench
// (tmem_ld_gen.h: the tcgen05.ld wrappers, generated once by a 15-line python snippet.)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tmem_ld_gen.h"

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// empty volatile asm statements keep their program order relative to the tcgen05.ld / wait statements: pin() after the
// ld makes the operands opaque there, pin() before the wait consumes the results -> the reduction stays in between
template <int N>
__device__ __forceinline__ void pin(uint32_t* v) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(v[i]));
}
__device__ __forceinline__ uint32_t pmin16(const uint32_t* w) {
    uint32_t m0 = __vimin3_u16x2(w[0], w[1], w[2]), m1 = __vimin3_u16x2(w[3], w[4], w[5]);
    m0 = __vimin3_u16x2(m0, w[6], w[7]); m1 = __vimin3_u16x2(m1, w[8], w[9]);
    m0 = __vimin3_u16x2(m0, w[10], w[11]); m1 = __vimin3_u16x2(m1, w[12], w[13]);
    return __vimin3_u16x2(m0, m1, __vminu2(w[14], w[15]));
}

// butterfly reduce-scatter over 16 lanes: 64 packed words per lane (128 columns) -> 4 words per lane
// (8 columns), each the minimum over the lane's 16-lane half.  60 SHFL + 60 VIMNMX.
template <int N>
__device__ __forceinline__ void bfly(uint32_t* r, int bit, int lane) {
    // lanes with (lane & bit) keep the upper half, the others the lower half
    const bool up = lane & bit;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const uint32_t send = up ? r[i] : r[i + N / 2];
        const uint32_t keep = up ? r[i + N / 2] : r[i];
        r[i] = __vminu2(keep, __shfl_xor_sync(0xFFFFFFFFu, send, bit));
    }
}

// MODE: 0 = 4 x (x16 pack), 1 = 2 x (x32 pack), 2 = 1 x (x64 pack), 3 = 4 x (x32 plain), 4 = 1 x (x128 plain),
//       5 = x64 pack + row minima (pmin16 x 4) after the wait, 6 = same but reduce overlapped with the next ld,
//       7 = x64 pack + shuffle butterfly (column minima per 16 lanes) + row minima, 8 = shuffle butterfly only
//       9 = 8 x (x8 pack)?? (not used)
template <int MODE>
__global__ void __launch_bounds__(256, (MODE == 3 || MODE == 4 || MODE == 6 || MODE == 7 || MODE == 9) ? 1 : 2) k(long long* cyc, uint32_t* sink, int iters, int cols_alloc) {
    __shared__ uint32_t tmem_base_s;
    extern __shared__ unsigned char dyn[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (cols_alloc == 512) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_s;
    const int q = warp & 3;
    // two warps of a quarter read different 128-column ranges
    const uint32_t ta = base + ((uint32_t)(q * 32) << 16) + (uint32_t)((warp >> 2) * 128 % cols_alloc);
    uint32_t acc = 0;
    uint32_t r[(MODE == 3 || MODE == 4) ? 128 : 64];
    uint32_t prev[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) prev[i] = lane * 77 + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int g = 0; g < 4; ++g) ld32x32_x16p(ta + 32 * g, r + 16 * g);
            ld_wait();
            acc ^= r[0] ^ r[17] ^ r[34] ^ r[63];
        } else if (MODE == 1) {
            ld32x32_x32p(ta, r); ld32x32_x32p(ta + 64, r + 32);
            ld_wait();
            acc ^= r[0] ^ r[17] ^ r[34] ^ r[63];
        } else if (MODE == 2) {
            ld32x32_x64p(ta, r);
            ld_wait();
            acc ^= r[0] ^ r[17] ^ r[34] ^ r[63];
        } else if (MODE == 3) {
#pragma unroll
            for (int g = 0; g < 4; ++g) ld32x32_x32(ta + 32 * g, r + 32 * g);
            ld_wait();
            acc ^= r[0] ^ r[37] ^ r[74] ^ r[127];
        } else if (MODE == 4) {
            ld32x32_x128(ta, r);
            ld_wait();
            acc ^= r[0] ^ r[37] ^ r[74] ^ r[127];
        } else if (MODE == 5) {
            ld32x32_x64p(ta, r);
            ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) acc ^= pmin16(r + 16 * g);
        } else if (MODE == 6) {
            // software pipeline: the transfer of accumulator it overlaps the reduction of accumulator it-1
            ld32x32_x64p(ta, r);
            pin<64>(prev);
#pragma unroll
            for (int g = 0; g < 4; ++g) acc ^= pmin16(prev + 16 * g);
            pin<1>(&acc);
            ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) prev[i] = r[i];
        } else if (MODE == 7) {
            ld32x32_x64p(ta, r);
            ld_wait();
#pragma unroll
            for (int g = 0; g < 4; ++g) acc ^= pmin16(r + 16 * g);
            bfly<64>(r, 8, lane); bfly<32>(r, 4, lane); bfly<16>(r, 2, lane); bfly<8>(r, 1, lane);
            acc ^= r[0] ^ r[1] ^ r[2] ^ r[3];
        } else if (MODE == 8) {
#pragma unroll
            for (int i = 0; i < 64; ++i) r[i] = prev[i] + it;
            bfly<64>(r, 8, lane); bfly<32>(r, 4, lane); bfly<16>(r, 2, lane); bfly<8>(r, 1, lane);
            acc ^= r[0] ^ r[1] ^ r[2] ^ r[3];
        } else if (MODE == 9) {
            // pipelined + butterfly: ld(it) in flight while (it-1) is reduced both ways
            ld32x32_x64p(ta, r);
            pin<64>(prev);
#pragma unroll
            for (int g = 0; g < 4; ++g) acc ^= pmin16(prev + 16 * g);
            bfly<64>(prev, 8, lane); bfly<32>(prev, 4, lane); bfly<16>(prev, 2, lane); bfly<8>(prev, 1, lane);
            acc ^= prev[0] ^ prev[1] ^ prev[2] ^ prev[3];
            pin<1>(&acc);
            ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) prev[i] = r[i];
        }
    }
    const long long t1 = clock64();
    if (MODE == 6 || MODE == 9) {
#pragma unroll
        for (int i = 0; i < 64; ++i) acc ^= prev[i];
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (lane == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (cols_alloc == 512) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(256));
    }
}

template <int MODE>
static void run(const char* name, int warps, int ctas_per_sm, long long* cyc, uint32_t* sink) {
    const int iters = 2000;
    const int cols = ctas_per_sm == 1 ? 512 : 256;
    // dynamic shared memory sized so that exactly ctas_per_sm CTAs fit on an SM
    const int smem = ctas_per_sm == 1 ? 120 * 1024 : 80 * 1024;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<148 * ctas_per_sm, warps * 32, smem>>>(cyc, sink, iters, cols);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
        cudaEventElapsedTime(&ms, e0, e1);
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE>, warps * 32, smem);
    if (occ != ctas_per_sm) { printf("%-58s %d warps x %d CTA/SM: skipped (occupancy %d)\n", name, warps, ctas_per_sm, occ); return; }
    long long h[8];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const double per_warp = (double)mx / iters;
    const double accs_per_sm = (double)(warps / 4) * ctas_per_sm;            // 128x128 accumulators drained per iteration and SM
    printf("%-58s %d warps x %d CTA/SM: %7.1f cyc/iter/warp  -> %6.1f cyc per 128x128 accumulator per SM  (kernel %.1f us)\n",
           name, warps, ctas_per_sm, per_warp, per_warp / accs_per_sm, ms * 1e3);
}

int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 148 * 2 * 8 * 8); cudaMalloc(&sink, 148 * 2 * 256 * 4);
    for (int cps = 1; cps <= 2; ++cps)
        for (int warps = 4; warps <= 8; warps += 4) {
            run<0>("4 x (32x32b.x16.pack16) + wait", warps, cps, cyc, sink);
            run<1>("2 x (32x32b.x32.pack16) + wait", warps, cps, cyc, sink);
            run<2>("1 x (32x32b.x64.pack16) + wait", warps, cps, cyc, sink);
            run<3>("4 x (32x32b.x32) + wait", warps, cps, cyc, sink);
            run<4>("1 x (32x32b.x128) + wait", warps, cps, cyc, sink);
            run<5>("x64.pack16 + wait + row minima", warps, cps, cyc, sink);
            run<6>("x64.pack16 || row minima of previous + wait", warps, cps, cyc, sink);
            run<7>("x64.pack16 + wait + row minima + shuffle column minima", warps, cps, cyc, sink);
            run<8>("shuffle column minima only (60 SHFL + 60 VIMNMX)", warps, cps, cyc, sink);
            run<9>("x64.pack16 || (row + shuffle column minima of previous)", warps, cps, cyc, sink);
        }
    return 0;
}
