// sector_bench.cu -- is a sparse row gather (256 of 2048 floats per 8 KB row, sorted indices) cheaper than
// streaming the whole row?  Answers whether the forward gather at config A can beat its "whole rows by TMA" form.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_bench sector_bench.cu && ./sector_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int N = 2048, ROWS = 8192, SEL = 256, C = 256;   // rows of one sample share an index list

__global__ void sparse_rows(const float* __restrict__ x, const unsigned short* __restrict__ idx, float* __restrict__ out,
                            int rows_per_cta) {
    __shared__ unsigned short sidx[SEL];
    const int r0 = blockIdx.x * rows_per_cta;
    int cur = -1;
    for (int r = r0; r < min(ROWS, r0 + rows_per_cta); ++r) {
        const int b = r / C;
        if (b != cur) { __syncthreads(); sidx[threadIdx.x] = idx[b * SEL + threadIdx.x]; cur = b; __syncthreads(); }
        out[(size_t)r * SEL + threadIdx.x] = __ldg(x + (size_t)r * N + sidx[threadIdx.x]);    // 256 threads = 256 picks
    }
}
__global__ void dense_rows(const float4* __restrict__ x, float* __restrict__ out, int rows_per_cta) {
    const int r0 = blockIdx.x * rows_per_cta;
    float acc = 0.f;
    for (int r = r0; r < min(ROWS, r0 + rows_per_cta); ++r) {
        const float4 a = __ldg(x + (size_t)r * (N / 4) + threadIdx.x), b = __ldg(x + (size_t)r * (N / 4) + 256 + threadIdx.x);
        acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

int main() {
    const int NBUF = 4;
    float* x[NBUF]; float* out; unsigned short* idx;
    for (int i = 0; i < NBUF; ++i) { cudaMalloc(&x[i], (size_t)ROWS * N * 4); cudaMemset(x[i], 0, (size_t)ROWS * N * 4); }
    cudaMalloc(&out, (size_t)ROWS * SEL * 4); cudaMalloc(&idx, (ROWS / C) * SEL * 2);
    std::vector<unsigned short> h((ROWS / C) * SEL);
    srand(1);
    for (int b = 0; b < ROWS / C; ++b) {
        std::vector<int> perm(N); for (int i = 0; i < N; ++i) perm[i] = i;
        for (int i = 0; i < SEL; ++i) std::swap(perm[i], perm[i + rand() % (N - i)]);
        for (int s = 0; s < 2; ++s) {}
        std::sort(perm.begin(), perm.begin() + SEL);
        for (int i = 0; i < SEL; ++i) h[b * SEL + i] = (unsigned short)perm[i];
    }
    cudaMemcpy(idx, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rpc : {1, 2, 4, 8}) {
        for (int mode = 0; mode < 2; ++mode) {
            const int grid = (ROWS + rpc - 1) / rpc;
            for (int w = 0; w < 3; ++w) { if (mode) sparse_rows<<<grid, 256>>>(x[w % NBUF], idx, out, rpc); else dense_rows<<<grid, 256>>>((const float4*)x[w % NBUF], out, rpc); }
            cudaEventRecord(e0);
            const int reps = 40;
            for (int i = 0; i < reps; ++i) { if (mode) sparse_rows<<<grid, 256>>>(x[i % NBUF], idx, out, rpc); else dense_rows<<<grid, 256>>>((const float4*)x[i % NBUF], out, rpc); }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%-7s rows/CTA %d: %.2f us per 64 MiB tensor\n", mode ? "sparse" : "dense", rpc, ms * 1e3 / reps);
        }
    }
    return 0;
}
