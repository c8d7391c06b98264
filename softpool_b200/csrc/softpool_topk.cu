// softpool_topk.cu -- per-region descending top-k of the activation rows (sm_100a).
//
// Replaces the region loop's `torch.sort(..., descending=True)` + `[:, :k]` (reference
// softpool.py:139-142), the float index cube (softpool.py:136-137,146-147) and the Sorter's
// argmax (softpool.py:95).  One CTA per (b, r) row; the row lives in shared memory as unique
// 64-bit words  (order_key(key) << 32) | (0xFFFFFFFF - n)  so that ANY sorting network yields the
// stable descending order of the reference (ties -> lower n first, NaN first, -0 == +0).
//
// Algorithm: bitonic top-k.  Sort chunks of K2 = pow2 >= k with alternating direction, then
// repeatedly keep the element-wise max of chunk pairs (the K2 largest of a bitonic sequence of
// 2*K2) and re-merge, halving the live length each round.  Work ~ N*(log^2 K2 / 2 + 2 log K2)
// compare-exchanges instead of N*log^2 N / 2 for the full sort; degenerates to a full bitonic
// sort when K2 == Npad (the reference operating point k*R == N with R == 1, or k > N/2).
#include "spk_common.cuh"

namespace spk {

__device__ __forceinline__ void ce_stage(uint64_t* buf, int len, int size, int stride, int tid,
                                         int nthr) {
    for (int p = tid; p < (len >> 1); p += nthr) {
        const int i = ((p & ~(stride - 1)) << 1) | (p & (stride - 1));
        const int j = i + stride;
        const bool desc = (i & size) == 0;
        const uint64_t a = buf[i], c = buf[j];
        if ((a < c) == desc) { buf[i] = c; buf[j] = a; }
    }
}

// Strides <= 32 only move data inside the 64-element block a warp-iteration owns, so those
// stages need a __syncwarp only; a CTA barrier is needed around every stride >= 64 stage.
__device__ __forceinline__ void stage_sync(int stride, int prev_stride) {
    if (stride >= 64 || prev_stride >= 64) __syncthreads(); else __syncwarp();
}

__global__ void __launch_bounds__(1024)
sp_topk_kernel(const float* __restrict__ keys, int R, int N, int k, int Npad, int K2,
               int32_t* __restrict__ idx, float* __restrict__ sp_idx,
               int64_t* __restrict__ id_activa) {
    extern __shared__ __align__(16) uint64_t buf[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int row = blockIdx.x;
    const int b = row / R, r = row - b * R;
    const float* krow = keys + (size_t)row * N;

    // ---- load + pack ----------------------------------------------------------------------
    if ((N & 3) == 0) {
        const float4* k4 = reinterpret_cast<const float4*>(krow);
        for (int q = tid; q < (Npad >> 2); q += nthr) {
            const int i = q << 2;
            if (i < N) {
                const float4 v = __ldg(k4 + q);
                buf[i + 0] = ((uint64_t)order_key(v.x) << 32) | (uint32_t)(0xFFFFFFFFu - (i + 0));
                buf[i + 1] = ((uint64_t)order_key(v.y) << 32) | (uint32_t)(0xFFFFFFFFu - (i + 1));
                buf[i + 2] = ((uint64_t)order_key(v.z) << 32) | (uint32_t)(0xFFFFFFFFu - (i + 2));
                buf[i + 3] = ((uint64_t)order_key(v.w) << 32) | (uint32_t)(0xFFFFFFFFu - (i + 3));
            } else {
                buf[i + 0] = 0ull; buf[i + 1] = 0ull; buf[i + 2] = 0ull; buf[i + 3] = 0ull;
            }
        }
        if (Npad < 4) for (int i = tid; i < Npad; i += nthr) buf[i] = 0ull;  // N%4==0 => N>=4, unreachable
    } else {
        for (int i = tid; i < Npad; i += nthr)
            buf[i] = (i < N) ? (((uint64_t)order_key(__ldg(krow + i)) << 32) | (uint32_t)(0xFFFFFFFFu - i))
                             : 0ull;   // pads sort last: every real word is > 0
    }

    // ---- argmax over the R rows for this CTA's slice of n (softpool.py:95) --------------------
    if (id_activa != nullptr) {
        const int slice = (N + R - 1) / R;
        const int n0 = r * slice, n1 = min(N, n0 + slice);
        const float* kb = keys + (size_t)b * R * N;
        for (int n = n0 + tid; n < n1; n += nthr) {
            uint32_t bestk = order_key(__ldg(kb + n));
            int besti = 0;
            for (int rr = 1; rr < R; ++rr) {
                const uint32_t kk = order_key(__ldg(kb + (size_t)rr * N + n));
                if (kk > bestk) { bestk = kk; besti = rr; }   // strict: first maximum / first NaN
            }
            id_activa[(size_t)b * N + n] = (int64_t)besti;
        }
    }
    __syncthreads();

    // ---- sort chunks of K2, directions alternating ---------------------------------------------
    int prev = 64;
    for (int size = 2; size <= K2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            stage_sync(stride, prev);
            ce_stage(buf, Npad, size, stride, tid, nthr);
            prev = stride;
        }

    // ---- halve: keep the K2 largest of every chunk pair, re-merge ------------------------------
    const int lg = __ffs(K2) - 1;
    for (int len = Npad; len > K2;) {
        const int half = len >> 1;
        uint64_t v[8];
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int p = tid + it * nthr;
            if (p < half) {
                const int c = p >> lg, i = p & (K2 - 1);
                const uint64_t a = buf[((2 * c) << lg) + i], d = buf[((2 * c + 1) << lg) + i];
                v[it] = a > d ? a : d;
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int p = tid + it * nthr;
            if (p < half) buf[p] = v[it];
        }
        len = half;
        prev = 64;
        for (int stride = K2 >> 1; stride > 0; stride >>= 1) {
            stage_sync(stride, prev);
            ce_stage(buf, len, K2, stride, tid, nthr);
            prev = stride;
        }
    }
    __syncthreads();

    // ---- emit -----------------------------------------------------------------------------------
    int32_t* orow = idx + (size_t)row * k;
    for (int j = tid; j < k; j += nthr) orow[j] = (int32_t)(0xFFFFFFFFu - (uint32_t)buf[j]);
    if (sp_idx != nullptr) {
        const int Q = R + 3;
        for (int e = tid; e < Q * k; e += nthr) {
            const int q = e / k, j = e - q * k;
            sp_idx[(((size_t)b * Q + q) * R + r) * k + j] = (float)(0xFFFFFFFFu - (uint32_t)buf[j]);
        }
    }
}

__global__ void sp_argmax_kernel(const float* __restrict__ keys, int R, int N,
                                 int64_t* __restrict__ id_activa, long long total) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long b = t / N;
        const int n = (int)(t - b * N);
        const float* kb = keys + (size_t)b * R * N;
        uint32_t bestk = order_key(__ldg(kb + n));
        int besti = 0;
        for (int rr = 1; rr < R; ++rr) {
            const uint32_t kk = order_key(__ldg(kb + (size_t)rr * N + n));
            if (kk > bestk) { bestk = kk; besti = rr; }
        }
        id_activa[t] = (int64_t)besti;
    }
}

}  // namespace spk

extern "C" int sp_topk_f32(const float* keys, int B, int R, int N, int k, int32_t* idx,
                           float* sp_idx, int64_t* id_activa, void* stream) {
    using namespace spk;
    if (B < 0 || R < 1 || N < 1 || k < 1 || k > N) return fail(SPK_E_BADARG, "sp_topk_f32: need B>=0, R>=1, 1<=k<=N (B=%d R=%d N=%d k=%d)", B, R, N, k);
    if (B == 0) return SPK_OK;
    if (!keys || !idx) return fail(SPK_E_BADARG, "sp_topk_f32: null keys/idx");
    if (N > 16384) return fail(SPK_E_UNSUPPORTED, "sp_topk_f32: N=%d > 16384 (row must fit shared memory)", N);
    if (sp_idx && N >= (1 << 24)) return fail(SPK_E_UNSUPPORTED, "sp_topk_f32: N >= 2^24 not exact in float32");
    if ((N & 3) == 0 && ((uintptr_t)keys & 15)) return fail(SPK_E_ALIGN, "sp_topk_f32: keys must be 16-byte aligned");
    const int K2 = max(2, next_pow2(k));
    const int Npad = max(K2, next_pow2(N));
    const int nthr = min(1024, max(32, Npad / 2));
    const size_t smem = (size_t)Npad * sizeof(uint64_t);
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sp_topk_kernel<<<B * R, nthr, smem, (cudaStream_t)stream>>>(keys, R, N, k, Npad, K2, idx, sp_idx, id_activa);
    SPK_LAUNCH_CHECK("sp_topk_kernel");
    return SPK_OK;
}

extern "C" int sp_argmax_i64(const float* keys, int B, int R, int N, int64_t* id_activa, void* stream) {
    using namespace spk;
    if (B < 0 || R < 1 || N < 1) return fail(SPK_E_BADARG, "sp_argmax_i64: need B>=0, R>=1, N>=1");
    if (B == 0) return SPK_OK;
    if (!keys || !id_activa) return fail(SPK_E_BADARG, "sp_argmax_i64: null pointer");
    const long long total = (long long)B * N;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
    sp_argmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(keys, R, N, id_activa, total);
    SPK_LAUNCH_CHECK("sp_argmax_kernel");
    return SPK_OK;
}
