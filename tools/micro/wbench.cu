// Write-path microbenchmarks: how fast can 64 MiB be written on a B200, by pattern?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void fill_gridstride(float4* p, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void fill_contig(float4* p, size_t n4) {     // every CTA its own contiguous range
    size_t lo = n4 * blockIdx.x / gridDim.x, hi = n4 * (blockIdx.x + 1) / gridDim.x;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void fill_contig_cs(float4* p, size_t n4) {
    size_t lo = n4 * blockIdx.x / gridDim.x, hi = n4 * (blockIdx.x + 1) / gridDim.x;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
}
// bulk store from shared memory: each CTA repeatedly stores a `chunk`-byte smem buffer to its range
__global__ void fill_bulk(char* p, size_t bytes, int chunk, int nbuf) {
    extern __shared__ __align__(128) char sm[];
    size_t lo = (bytes / chunk) * blockIdx.x / gridDim.x, hi = (bytes / chunk) * (blockIdx.x + 1) / gridDim.x;
    for (int i = threadIdx.x; i < chunk * nbuf / 4; i += blockDim.x) ((float*)sm)[i] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = 0;
        for (size_t c = lo; c < hi; ++c) {
            uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + (size_t)b * chunk);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + c * chunk), "r"(s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            b = (b + 1) % nbuf;
            if (nbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
__global__ void read_contig(const float4* p, size_t n4, float* out) {
    size_t lo = n4 * blockIdx.x / gridDim.x, hi = n4 * (blockIdx.x + 1) / gridDim.x;
    float s = 0.f;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) { float4 v = p[i]; s += v.x + v.y + v.z + v.w; }
    if (s == 12345.f) out[0] = s;
}

template <class F> float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(0); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f(i);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / reps;
}

int main() {
    const size_t bytes = 64ull << 20;
    const int NB = 4;                       // rotate over 4 buffers (256 MiB > L2)
    char* buf[NB]; float* out;
    for (int i = 0; i < NB; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], 0, bytes)); }
    CK(cudaMalloc(&out, 4));
    const size_t n4 = bytes / 16;
    int sms = 148;
    CK(cudaFuncSetAttribute(fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rot = 1; rot <= NB; rot += NB - 1) {
        printf("--- %s (64 MiB per launch) ---\n", rot == 1 ? "same buffer (L2-warm)" : "rotating 4 buffers");
        auto rep = [&](const char* name, float us) { printf("%-44s %7.2f us  %6.2f TB/s\n", name, us, bytes / us * 1e-6); };
        rep("fill grid-stride 148x8 CTAs x256", timeit([&](int i) { fill_gridstride<<<sms * 8, 256>>>((float4*)buf[i % rot], n4); }, 40));
        rep("fill grid-stride 16384 CTAs x256", timeit([&](int i) { fill_gridstride<<<16384, 256>>>((float4*)buf[i % rot], n4); }, 40));
        rep("fill contiguous-range 444 CTAs x256", timeit([&](int i) { fill_contig<<<444, 256>>>((float4*)buf[i % rot], n4); }, 40));
        rep("fill contiguous-range 1184 CTAs x256", timeit([&](int i) { fill_contig<<<1184, 256>>>((float4*)buf[i % rot], n4); }, 40));
        rep("fill contiguous-range .cs 444 CTAs x256", timeit([&](int i) { fill_contig_cs<<<444, 256>>>((float4*)buf[i % rot], n4); }, 40));
        rep("fill contiguous-range .cs 1184 CTAs x256", timeit([&](int i) { fill_contig_cs<<<1184, 256>>>((float4*)buf[i % rot], n4); }, 40));
        rep("bulk store 32 KB x2buf, 444 CTAs", timeit([&](int i) { fill_bulk<<<444, 256, 64 * 1024>>>(buf[i % rot], bytes, 32 * 1024, 2); }, 40));
        rep("bulk store 16 KB x2buf, 444 CTAs", timeit([&](int i) { fill_bulk<<<444, 256, 32 * 1024>>>(buf[i % rot], bytes, 16 * 1024, 2); }, 40));
        rep("bulk store 8 KB x2buf, 888 CTAs", timeit([&](int i) { fill_bulk<<<888, 256, 16 * 1024>>>(buf[i % rot], bytes, 8 * 1024, 2); }, 40));
        rep("bulk store 32 KB x1buf, 444 CTAs", timeit([&](int i) { fill_bulk<<<444, 256, 32 * 1024>>>(buf[i % rot], bytes, 32 * 1024, 1); }, 40));
        rep("read contiguous-range 444 CTAs x256", timeit([&](int i) { read_contig<<<444, 256>>>((const float4*)buf[i % rot], n4, out); }, 40));
        rep("read contiguous-range 1184 CTAs x256", timeit([&](int i) { read_contig<<<1184, 256>>>((const float4*)buf[i % rot], n4, out); }, 40));
    }
    return 0;
}
