// minbench.cu -- issue rate of the min instructions the Chamfer epilogue can use (sm_100a):
//   FMNMX3 (3-input f32 min), VIMNMX3.U16x2 (3-input packed 16-bit min, DPX), HMNMX2, FMNMX
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o minbench minbench.cu && ./minbench
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void k(const unsigned* in, unsigned* out, long long* cyc, int iters) {
    unsigned a[8], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + i];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); a[i] = __float_as_uint(r); }
            if (MODE == 1) a[i] = __vimin3_u16x2(a[i], b, c);
            if (MODE == 2) { __half2 h = __hmin2(*(__half2*)&a[i], *(__half2*)&b); a[i] = *(unsigned*)&h; }
            if (MODE == 3) a[i] = __float_as_uint(fminf(__uint_as_float(a[i]), __uint_as_float(b)));
            if (MODE == 4) a[i] = __vimin3_s32(a[i], b, c);
        }
        b += 1; c ^= 3;
    }
    const long long t1 = clock64();
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    unsigned *in, *out; long long* cyc;
    cudaMalloc(&in, 4096); cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    cudaMemset(in, 1, 4096);
    const char* names[] = {"FMNMX3 (min.f32 x3)", "VIMNMX3.U16x2", "HMNMX2", "FMNMX (2-input)", "VIMNMX3.S32"};
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int m = 0; m < 5; ++m) {
            const int iters = 4096;
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<148, warps * 32>>>(in, out, cyc, iters);
                if (m == 1) k<1><<<148, warps * 32>>>(in, out, cyc, iters);
                if (m == 2) k<2><<<148, warps * 32>>>(in, out, cyc, iters);
                if (m == 3) k<3><<<148, warps * 32>>>(in, out, cyc, iters);
                if (m == 4) k<4><<<148, warps * 32>>>(in, out, cyc, iters);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            // warp-instructions per SMSP = (warps/4) * iters * 8
            printf("%-22s %2d warps/SM: %.2f cycles per warp-instruction per SMSP\n", names[m], warps, (double)h / ((warps / 4.0) * iters * 8));
        }
    return 0;
}
