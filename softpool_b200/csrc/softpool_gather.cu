// softpool_gather.cu -- SoftPool gather forward (+ window max) and its atomic-free backward.
//
// Forward replaces softpool.py:142-145 + train2cabins (softpool.py:71-85); backward is what
// autograd derives for them (the reference has no hand-written backward).  Both are pure data
// movement and HBM-bound; design ("gather"):
//   * persistent grid (CTAs = resident slots); every CTA owns a contiguous, balanced range of the
//     B*C rows and walks it in TILES of T rows of one sample (one index list serves a tile);
//   * forward: a tile of x (T*N floats, contiguous) arrives in shared memory by ONE 1-D TMA bulk
//     copy (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP), double-buffered, so the random
//     index access hits shared memory, never HBM/L2, and every x byte is read exactly once; the
//     sample's index list sits in shared memory as u16; outputs leave as 128-bit streaming stores;
//   * backward: per sample the CTA orders the R*k slots by rank (number of lower regions that chose
//     the same point); slots of one rank hit different points, so a rank bucket is scattered into
//     a T x N shared-memory accumulator without conflicts or atomics, buckets in ascending rank
//     (= ascending region: deterministic order); the T finished rows leave with one TMA bulk store
//     (cp.async.bulk.global.shared::cta) while the next tile is accumulated in the other buffer.
//     grad_x is fully overwritten (zeros included): no memset.
#include "spk_common.cuh"

namespace spk {

constexpr int G_THREADS = 256;

__device__ __forceinline__ float4 ld_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// contiguous range of global rows of CTA `cta`
__device__ __forceinline__ void cta_range(long long rows, long long& lo, long long& hi) {
    lo = (long long)((unsigned long long)rows * blockIdx.x / gridDim.x);
    hi = (long long)((unsigned long long)rows * (blockIdx.x + 1) / gridDim.x);
}
// rows of the tile starting at global row g: <= T, inside the CTA range, inside one sample
__device__ __forceinline__ int tile_rows(long long g, long long g_hi, int T, int C) {
    // the host guarantees B*C < 2^31: 32-bit modulo instead of the emulated 64-bit one
    const unsigned in_sample = (unsigned)C - (unsigned)g % (unsigned)C;
    long long r = g_hi - g;
    if (r > T) r = T;
    if (r > (long long)in_sample) r = in_sample;
    return (int)r;
}
// cooperative load of one sample's index list into shared memory as u16 (4 independent loads in flight)
__device__ __forceinline__ void stage_idx_u16(const int32_t* __restrict__ src, uint16_t* dst, int n, int tid) {
    int i = tid;
    for (; i + 3 * G_THREADS < n; i += 4 * G_THREADS) {
        const int a = __ldg(src + i), b = __ldg(src + i + G_THREADS), c = __ldg(src + i + 2 * G_THREADS),
                  d = __ldg(src + i + 3 * G_THREADS);
        dst[i] = (uint16_t)a; dst[i + G_THREADS] = (uint16_t)b; dst[i + 2 * G_THREADS] = (uint16_t)c;
        dst[i + 3 * G_THREADS] = (uint16_t)d;
    }
    for (; i < n; i += G_THREADS) dst[i] = (uint16_t)__ldg(src + i);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct GatherFwdParams {
    const float* x; const int32_t* idx;
    float* sp_cube; float* cabins; uint16_t* cab_arg;
    long long rows;       // B*C
    int C, N, R, k, cab;
    int T;                // rows per tile
    int bulk_ok;          // tiles can move with cp.async.bulk (N%4==0, x 16B aligned)
    int vec4;             // k%4==0 && sp_cube 16B aligned: 4 slots per work item, 128-bit stores
    int cab_fast;         // windows are whole groups of 4 slots, (wl/4) pow2 <= 32, (R*k/4) % 32 == 0 if > 1
    int g_shift;          // log2(wl/4) when cab_fast
    int q_shift;          // log2(R*k/4) when that is a power of two, else -1
};

__global__ void __launch_bounds__(G_THREADS)
sp_gather_fwd_kernel(const GatherFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int RK = p.R * p.k, N = p.N, T = p.T, C = p.C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                            // 2
    uint16_t* idx_s = reinterpret_cast<uint16_t*>(smem_raw + 128);                     // RK
    float* tiles = reinterpret_cast<float*>(smem_raw + 128 + ((RK * 2 + 127) & ~127)); // 2 * T*N
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncthreads();

    long long g_lo, g_hi;
    cta_range(p.rows, g_lo, g_hi);
    auto issue = [&](long long g, int rows, int buf) {             // tid 0 only
        if (p.bulk_ok) {
            const uint32_t bytes = (uint32_t)rows * (uint32_t)N * 4u;
            mbar_expect_tx(&bars[buf], bytes);
            bulk_g2s(tiles + (size_t)buf * T * N, p.x + (size_t)g * N, bytes, &bars[buf]);
        }
    };
    // prologue: two tiles in flight
    long long g = g_lo, g_pref = g_lo;
    if (tid == 0) {
        for (int i = 0; i < 2 && g_pref < g_hi; ++i) { const int r = tile_rows(g_pref, g_hi, T, C); issue(g_pref, r, i); g_pref += r; }
    } else {
        for (int i = 0; i < 2 && g_pref < g_hi; ++i) g_pref += tile_rows(g_pref, g_hi, T, C);
    }

    const int wl = p.cab > 0 ? p.k / p.cab : 0;          // window length
    const int G = p.cab_fast ? (wl >> 2) : 1;            // work items (lanes) per window
    const int wins_per_row = p.R * p.cab;
    const int Q = RK >> 2;                               // 4-slot work items per row
    int cur_b = -1;

    for (int ti = 0; g < g_hi; ++ti) {
        const int buf = ti & 1;
        const int rows = tile_rows(g, g_hi, T, C);
        const int b = (int)((unsigned)g / (unsigned)C);
        if (b != cur_b) {                                // new sample: its index list (readers of the old one
            stage_idx_u16(p.idx + (size_t)b * RK, idx_s, RK, tid);      // passed the barrier ending the last tile)
            cur_b = b;
            __syncthreads();
        }
        float* tile = tiles + (size_t)buf * T * N;
        if (p.bulk_ok) {
            mbar_wait(&bars[buf], (uint32_t)((ti >> 1) & 1));
        } else {
            const float* src = p.x + (size_t)g * N;
            for (int i = tid; i < rows * N; i += G_THREADS) tile[i] = __ldg(src + i);
            __syncthreads();
        }
        if (p.vec4) {
            const int items = rows * Q;
            const int items_up = (items + G_THREADS - 1) / G_THREADS * G_THREADS;     // whole warps reach the shuffles
            for (int e = tid; e < items_up; e += G_THREADS) {
                const bool live = e < items;
                int t, q;
                if (p.q_shift >= 0) { t = e >> p.q_shift; q = e & (Q - 1); } else { t = e / Q; q = e - t * Q; }
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const float* row = tile + (size_t)t * N;
                if (live) {
                    const uint2 u = *reinterpret_cast<const uint2*>(idx_s + 4 * q);
                    v.x = row[u.x & 0xFFFFu]; v.y = row[u.x >> 16]; v.z = row[u.y & 0xFFFFu]; v.w = row[u.y >> 16];
                    st_cs_f4(reinterpret_cast<float4*>(p.sp_cube + (size_t)(g + t) * RK) + q, v);
                }
                if (p.cab_fast) {
                    // torch.max over the window: first maximum, a NaN wins (first NaN); -0 == +0.
                    // local winner of this item's 4 slots with plain float compares ...
                    float best = v.x; uint32_t off = 0;
                    if (!(v.y <= best) && best == best) { best = v.y; off = 1; }
                    if (!(v.z <= best) && best == best) { best = v.z; off = 2; }
                    if (!(v.w <= best) && best == best) { best = v.w; off = 3; }
                    uint32_t woff = off + ((uint32_t)(q & (G - 1)) << 2);         // offset inside the window
                    if (G > 1) {
                        // ... then one packed word per lane: greater key wins, equal keys -> lower offset
                        uint64_t pk = ((uint64_t)order_key(best) << 32) | (uint32_t)(0xFFFFFFFFu - woff);
                        for (int d = 1; d < G; d <<= 1) {
                            const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, pk, d);
                            pk = o > pk ? o : pk;
                        }
                        woff = 0xFFFFFFFFu - (uint32_t)pk;
                    }
                    if (live && (q & (G - 1)) == 0) {
                        const int win = q >> p.g_shift;                            // (r, w) flattened
                        const size_t o = (size_t)(g + t) * wins_per_row + win;
                        // the value itself (sign of zero, NaN payload) comes from the winning slot
                        p.cabins[o] = (G == 1) ? best : row[idx_s[win * wl + (int)woff]];
                        p.cab_arg[o] = (uint16_t)woff;
                    }
                }
            }
        } else {
            const int items = rows * RK;
            for (int e = tid; e < items; e += G_THREADS) {
                const int t = e / RK, s = e - t * RK;
                p.sp_cube[(size_t)(g + t) * RK + s] = tile[(size_t)t * N + idx_s[s]];
            }
        }
        g += rows;
        // release the buffer and refill it with the tile two ahead
        __syncthreads();
        if (g_pref < g_hi) {
            const int r = tile_rows(g_pref, g_hi, T, C);
            if (tid == 0) issue(g_pref, r, buf);
            g_pref += r;
        }
    }
}

// Generic window max over an already written sp_cube (any k, cab): one thread per window.
__global__ void sp_cabins_generic_kernel(const float* __restrict__ sp_cube, int k, int cab,
                                         long long n_rows /* B*C*R */, float* __restrict__ cabins,
                                         uint16_t* __restrict__ cab_arg) {
    const int wl = k / cab;
    const long long total = n_rows * cab;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long rowi = t / cab;
        const int w = (int)(t - rowi * cab);
        const float* src = sp_cube + rowi * k + (size_t)w * wl;
        float bv = src[0]; uint32_t bk = order_key(bv); int bo = 0;
        for (int j = 1; j < wl; ++j) {
            const float v = src[j]; const uint32_t kk = order_key(v);
            if (kk > bk) { bk = kk; bv = v; bo = j; }
        }
        cabins[t] = bv; cab_arg[t] = (uint16_t)bo;
    }
}

// Backward of the standalone window max: g_windows[row, w*wl + arg] = g_cabins[row, w], else 0.
__global__ void sp_cabins_bwd_kernel(const float* __restrict__ g_cabins, const uint16_t* __restrict__ cab_arg,
                                     int k, int cab, long long total /* rows*k */, float* __restrict__ g_windows) {
    const int wl = k / cab;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long rowi = t / k;
        const int j = (int)(t - rowi * k);
        const int w = j / wl;
        float g = 0.f;
        if (w < cab) {
            const long long o = rowi * cab + w;
            if ((int)cab_arg[o] == j - w * wl) g = g_cabins[o];
        }
        g_windows[t] = g;
    }
}

// ---------------------------------------------------------------------------------------------
// backward (default): rank-bucketed scatter into a shared-memory accumulator
//
// grad_x[b,c,n] = sum over the slots (r,j) that selected point n of g[b,c,r,j], ascending r.
// Per sample the CTA orders the R*k slots by RANK = number of lower regions that selected the same
// point.  Slots of one rank hit pairwise different points, so a rank bucket is scattered without
// conflicts and without atomics; buckets are processed in ascending rank with a barrier in between,
// which fixes the summation order (ascending region) -> deterministic.  The slot list is shared by
// all C rows of the sample; T rows are accumulated together (T independent read-modify-writes per
// list entry), then leave as ONE TMA bulk store while the next tile is accumulated in the other buffer.
// ---------------------------------------------------------------------------------------------
struct GatherBwdParams {
    const float* g_cube; const float* g_cabins; const int32_t* idx; const uint16_t* cab_arg;
    float* grad_x;
    long long rows;   // B*C
    int C, N, R, k, cab;
    int T;            // rows per tile (1, 2, 4 or 8)
    int bulk_in;      // (R*k)%4==0 and g_cube 16B aligned -> gradient tiles arrive with cp.async.bulk
    int bulk_out;     // N%4==0 and grad_x 16B aligned -> rows leave with cp.async.bulk
};

template <int T>
__device__ __forceinline__ void scatter_bucket(const uint32_t* ent, int lo, int hi, const float* gs, float* ac,
                                               int RK, int N, int tid) {
    for (int e = lo + tid; e < hi; e += G_THREADS) {
        const uint32_t en = ent[e];
        const int n = (int)(en & 0xFFFFu), s = (int)(en >> 16);
        float gv[T], av[T];
#pragma unroll
        for (int t = 0; t < T; ++t) { gv[t] = gs[t * RK + s]; av[t] = ac[t * N + n]; }
#pragma unroll
        for (int t = 0; t < T; ++t) ac[t * N + n] = av[t] + gv[t];
    }
}

__global__ void __launch_bounds__(G_THREADS)
sp_gather_bwd_kernel(const GatherBwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int roff[66];            // bucket offsets (ranks 0..63) + fill counters during the build
    __shared__ int rcount[64];
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = p.N, R = p.R, k = p.k, T = p.T, C = p.C;
    const int RK = R * k;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                        // 2
    uint32_t* ent = reinterpret_cast<uint32_t*>(smem_raw + 128);                   // RK   (slot << 16) | point, by rank
    float* gbuf = reinterpret_cast<float*>(smem_raw + 128 + (((size_t)RK * 4 + 127) & ~(size_t)127));   // 2 * T*RK
    float* acc = gbuf + 2 * (size_t)T * RK;                                        // 2 * T*N
    // build scratch, aliased onto the accumulators (no store is in flight while a sample is (re)built)
    uint16_t* idx_s = reinterpret_cast<uint16_t*>(acc);                            // RK  point of every slot
    uint16_t* rank = idx_s + RK;                                                   // RK  rank of every slot
    uint16_t* cnt = rank + RK;                                                     // N   per-point counter
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncthreads();

    long long g_lo, g_hi;
    cta_range(p.rows, g_lo, g_hi);
    const int wl = (p.g_cabins != nullptr) ? k / p.cab : 1;
    const int wins = R * p.cab;
    auto issue = [&](long long g, int rows, int buf) {             // tid 0 only
        if (p.bulk_in) {
            const uint32_t bytes = (uint32_t)rows * (uint32_t)RK * 4u;
            mbar_expect_tx(&bars[buf], bytes);
            bulk_g2s(gbuf + (size_t)buf * T * RK, p.g_cube + (size_t)g * RK, bytes, &bars[buf]);
        }
    };
    long long g = g_lo, g_pref = g_lo;
    if (tid == 0) {
        for (int i = 0; i < 2 && g_pref < g_hi; ++i) { const int r = tile_rows(g_pref, g_hi, T, C); issue(g_pref, r, i); g_pref += r; }
    } else {
        for (int i = 0; i < 2 && g_pref < g_hi; ++i) g_pref += tile_rows(g_pref, g_hi, T, C);
    }
    int cur_b = -1, n_rank = 0;

    for (int ti = 0; g < g_hi; ++ti) {
        const int buf = ti & 1;
        const int rows = tile_rows(g, g_hi, T, C);
        const int b = (int)((unsigned)g / (unsigned)C);
        // window-max gradient of this tile: fetch early, use after the gradient tile has landed
        float cab_g = 0.f; int cab_dst = -1;
        if (p.g_cabins != nullptr && tid < rows * wins) {          // rows*wins <= 8*R*cab; larger cases loop below
            const int t = tid / wins, rw = tid - t * wins;
            const int r = rw / p.cab, w = rw - r * p.cab;
            const size_t o = (size_t)(g + t) * wins + rw;
            cab_dst = t * RK + r * k + w * wl + (int)__ldg(p.cab_arg + o);
            cab_g = __ldg(p.g_cabins + o);
        }
        if (b != cur_b) {
            // ---- slot list of sample b ordered by rank: built once per sample segment of this CTA --------
            cur_b = b;
            if (p.bulk_out && tid == 0) bulk_wait_read<0>();           // scratch aliases the accumulators
            __syncthreads();
            for (int n = tid; n < N; n += G_THREADS) cnt[n] = 0;
            if (tid < 64) rcount[tid] = 0;
            stage_idx_u16(p.idx + (size_t)b * RK, idx_s, RK, tid);
            __syncthreads();
            int my_max = 0;
            for (int r = 0; r < R; ++r) {                              // points are unique inside a region
                for (int j = tid; j < k; j += G_THREADS) {
                    const int n = idx_s[r * k + j];
                    const int c = cnt[n];
                    rank[r * k + j] = (uint16_t)c;
                    cnt[n] = (uint16_t)(c + 1);
                    my_max = max(my_max, c);
                }
                __syncthreads();
            }
            my_max = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)my_max);
            if (lane == 0) atomicMax(&rcount[63], my_max);             // rcount[63] is free: R <= 64 -> ranks <= 63 ...
            __syncthreads();
            n_rank = min(rcount[63] + 1, R);                           // ... and rank 63 needs R == 64: handled below
            __syncthreads();
            if (tid == 0) rcount[63] = 0;
            __syncthreads();
            // bucket sizes (warp-aggregated), offsets, placement
            const int RK_up = (RK + 31) & ~31;
            for (int e = tid; e < RK_up; e += G_THREADS) {
                const int rk = e < RK ? (int)rank[e] : -1;
                for (int q = 0; q < n_rank; ++q) {
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, rk == q);
                    if (m && lane == 0) atomicAdd(&rcount[q], __popc(m));
                }
            }
            __syncthreads();
            if (tid == 0) {
                int run = 0;
                for (int q = 0; q < n_rank; ++q) { roff[q] = run; run += rcount[q]; rcount[q] = roff[q]; }
                roff[n_rank] = run;
            }
            __syncthreads();
            for (int e = tid; e < RK_up; e += G_THREADS) {
                const int rk = e < RK ? (int)rank[e] : -1;
                for (int q = 0; q < n_rank; ++q) {
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, rk == q);
                    if (m) {
                        int base = 0;
                        if (lane == (__ffs(m) - 1)) base = atomicAdd(&rcount[q], __popc(m));
                        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
                        if (rk == q) ent[base + __popc(m & ((1u << lane) - 1))] = ((uint32_t)e << 16) | (uint32_t)idx_s[e];
                    }
                }
            }
            __syncthreads();
        }
        float* gs = gbuf + (size_t)buf * T * RK;
        float* ac = acc + (size_t)buf * T * N;
        // acc[buf] was handed to the TMA store two tiles ago: wait until that store has READ it, then zero
        if (p.bulk_out && tid == 0) bulk_wait_read<1>();
        __syncthreads();
        {
            float4* a4 = reinterpret_cast<float4*>(ac);
            const int n4 = (rows * N) >> 2;
            for (int i = tid; i < n4; i += G_THREADS) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = (n4 << 2) + tid; i < rows * N; i += G_THREADS) ac[i] = 0.f;
        }
        if (p.bulk_in) {
            mbar_wait(&bars[buf], (uint32_t)((ti >> 1) & 1));
        } else {
            const float* src = p.g_cube + (size_t)g * RK;
            for (int i = tid; i < rows * RK; i += G_THREADS) gs[i] = __ldg(src + i);
            __syncthreads();
        }
        // fold the window-max gradient into the slot that won each window (unique slots)
        if (p.g_cabins != nullptr) {
            if (cab_dst >= 0) gs[cab_dst] += cab_g;
            for (int e = tid + G_THREADS; e < rows * wins; e += G_THREADS) {
                const int t = e / wins, rw = e - t * wins;
                const int r = rw / p.cab, w = rw - r * p.cab;
                const size_t o = (size_t)(g + t) * wins + rw;
                gs[t * RK + r * k + w * wl + (int)__ldg(p.cab_arg + o)] += __ldg(p.g_cabins + o);
            }
        }
        __syncthreads();
        for (int q = 0; q < n_rank; ++q) {
            const int lo = roff[q], hi = roff[q + 1];
            if (lo == hi) break;                                       // ranks are dense: nothing beyond
            switch (T) {
                case 1: scatter_bucket<1>(ent, lo, hi, gs, ac, RK, N, tid); break;
                case 2: scatter_bucket<2>(ent, lo, hi, gs, ac, RK, N, tid); break;
                case 4: scatter_bucket<4>(ent, lo, hi, gs, ac, RK, N, tid); break;
                default: scatter_bucket<8>(ent, lo, hi, gs, ac, RK, N, tid); break;
            }
            __syncthreads();
        }
        float* dst = p.grad_x + (size_t)g * N;
        if (p.bulk_out) {
            fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the TMA engine
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(dst, ac, (uint32_t)(rows * N) * 4u);
                bulk_commit();
            }
        } else {
            for (int i = tid; i < rows * N; i += G_THREADS) dst[i] = ac[i];
            __syncthreads();
        }
        g += rows;
        if (g_pref < g_hi) {                                  // gs[buf] is free: every thread passed a barrier after its reads
            const int r = tile_rows(g_pref, g_hi, T, C);
            if (tid == 0) issue(g_pref, r, buf);
            g_pref += r;
        }
    }
    if (p.bulk_out && tid == 0) bulk_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// backward, "push" formulation (fallback when the inverse index does not fit shared memory or
// R*k > 65535): region-ordered scatter into a shared-memory accumulator, TMA bulk store of the rows
// ---------------------------------------------------------------------------------------------
struct GatherBwdPushParams {
    const float* g_cube; const float* g_cabins; const int32_t* idx; const uint16_t* cab_arg;
    float* grad_x;
    long long rows;   // B*C
    int C, N, R, k, cab;
    int T;            // rows accumulated together in shared memory (one "group")
    int bulk_out;     // N%4==0 and grad_x 16B aligned -> rows leave with cp.async.bulk
    int bulk_in;      // (R*k)%4==0 and g_cube 16B aligned -> g rows arrive with cp.async.bulk
    int nbuf;         // 2 = double-buffered acc/g (default), 1 = single (rows too large for two)
    int g_direct;     // 1 = upstream gradient read straight from global (R*k too large to stage)
    int k_shift;      // log2(k) when k is a power of two, else -1
};

// Each CTA owns a contiguous range of global rows and walks it in groups of <= T rows that lie in
// ONE sample (so one index list serves the group).  Double-buffered: while group i is scattered
// into acc[i&1], the upstream gradient of group i+1 is in flight (bulk load) and the rows of
// group i-1 are still leaving (bulk store).
__global__ void __launch_bounds__(256)
sp_gather_bwd_push_kernel(const GatherBwdPushParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int N = p.N, R = p.R, k = p.k, T = p.T, C = p.C;
    const int RK = R * k;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                      // 2
    uint16_t* idx_s = reinterpret_cast<uint16_t*>(smem_raw + 128);               // RK
    float* gbuf = reinterpret_cast<float*>(smem_raw + 128 + ((RK * 2 + 127) & ~127));   // 2 * T*RK
    const int nbuf = p.nbuf;
    float* acc = gbuf + (p.g_direct ? 0 : (size_t)nbuf * T * RK);                // nbuf * T*N
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    __syncthreads();

    const long long g_lo = p.rows * blockIdx.x / gridDim.x;
    const long long g_hi = p.rows * (blockIdx.x + 1) / gridDim.x;
    const int wl = (p.g_cabins != nullptr) ? k / p.cab : 1;
    const int wins = R * p.cab;

    // group i covers rows [g, g + rows_i): never crosses a sample boundary
    auto group_rows = [&](long long g) -> int {
        const long long in_sample = (long long)C - (g % C);
        long long r = g_hi - g;
        if (r > T) r = T;
        if (r > in_sample) r = in_sample;
        return (int)r;
    };
    auto issue_load = [&](long long g, int rows, int buf) {        // tid 0 only
        if (p.bulk_in && !p.g_direct) {
            const uint32_t bytes = (uint32_t)rows * RK * 4u;
            mbar_expect_tx(&bars[buf], bytes);
            bulk_g2s(gbuf + (size_t)buf * T * RK, p.g_cube + (size_t)g * RK, bytes, &bars[buf]);
        }
    };

    long long g = g_lo;
    int cur_b = -1;
    if (g < g_hi && tid == 0) issue_load(g, group_rows(g), 0);
    for (int gi = 0; g < g_hi; ++gi) {
        const int buf = (nbuf == 2) ? (gi & 1) : 0;
        const int rows = group_rows(g);
        const long long g_next = g + rows;
        float* gs = gbuf + (size_t)buf * T * RK;
        float* ac = acc + (size_t)buf * T * N;
        // prefetch the next group's upstream gradient into the other buffer (its previous reader,
        // group gi-1, finished before the barrier that ended that iteration)
        if (nbuf == 2 && g_next < g_hi && tid == 0) issue_load(g_next, group_rows(g_next), buf ^ 1);
        // acc[buf] was handed to the TMA store nbuf groups ago: wait until that store has READ it
        if (p.bulk_out && tid == 0) { if (nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
        const int b = (int)(g / C);
        if (b != cur_b) {                                            // new sample: its index list
            __syncthreads();
            const int32_t* ib = p.idx + (size_t)b * RK;
            for (int i = tid; i < RK; i += nthr) idx_s[i] = (uint16_t)__ldg(ib + i);
            cur_b = b;
        }
        __syncthreads();
        {
            float4* a4 = reinterpret_cast<float4*>(ac);
            const int n4 = (rows * N) >> 2;
            for (int i = tid; i < n4; i += nthr) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = (n4 << 2) + tid; i < rows * N; i += nthr) ac[i] = 0.f;
        }
        if (p.g_direct) {
            // nothing staged
        } else if (p.bulk_in) {
            mbar_wait(&bars[buf], (uint32_t)((nbuf == 2 ? (gi >> 1) : gi) & 1));
        } else {
            const float* src = p.g_cube + (size_t)g * RK;
            for (int i = tid; i < rows * RK; i += nthr) gs[i] = __ldg(src + i);
        }
        __syncthreads();
        // fold the window-max gradient into the slot that won each window (unique slots)
        if (p.g_cabins != nullptr && !p.g_direct) {
            for (int e = tid; e < rows * wins; e += nthr) {
                const int t = e / wins, rw = e - t * wins;
                const int r = rw / p.cab, w = rw - r * p.cab;
                const size_t o = (size_t)(g + t) * wins + rw;
                gs[t * RK + r * k + w * wl + (int)__ldg(p.cab_arg + o)] += __ldg(p.g_cabins + o);
            }
            __syncthreads();
        }
        // scatter, one region per phase: indices are unique inside a region -> no conflicts; ascending
        // region order fixes the summation order
        const int per_region = rows * k;
        for (int r = 0; r < R; ++r) {
            for (int e = tid; e < per_region; e += nthr) {
                int t, j;
                if (p.k_shift >= 0) { t = e >> p.k_shift; j = e & (k - 1); } else { t = e / k; j = e - t * k; }
                float gv;
                if (!p.g_direct) {
                    gv = gs[t * RK + r * k + j];
                } else {
                    gv = __ldg(p.g_cube + (size_t)(g + t) * RK + (size_t)r * k + j);
                    if (p.g_cabins != nullptr) {
                        const int w = j / wl;
                        if (w < p.cab) {
                            const size_t o = (size_t)(g + t) * wins + r * p.cab + w;
                            if ((int)__ldg(p.cab_arg + o) == j - w * wl) gv += __ldg(p.g_cabins + o);
                        }
                    }
                }
                ac[t * N + idx_s[r * k + j]] += gv;
            }
            __syncthreads();
        }
        float* dst = p.grad_x + (size_t)g * N;
        if (p.bulk_out) {
            fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the TMA engine
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(dst, ac, (uint32_t)(rows * N) * 4u);
                bulk_commit();
            }
        } else {
            for (int i = tid; i < rows * N; i += nthr) dst[i] = ac[i];
            __syncthreads();
        }
        // single-buffered: the next group's gradient can only be fetched once this one is consumed
        if (nbuf == 1 && g_next < g_hi && tid == 0) issue_load(g_next, group_rows(g_next), 0);
        g = g_next;
    }
    if (p.bulk_out && tid == 0) bulk_wait<0>();
}

}  // namespace spk

static int occupancy_slots(const void* kernel, int threads, size_t smem, int cap_per_sm) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (cap_per_sm > 0 && per_sm > cap_per_sm) per_sm = cap_per_sm;
    return per_sm * spk::sm_count();
}
static int ilog2_exact(int v) {            // log2(v) if v is a power of two, else -1
    if (v <= 0 || (v & (v - 1))) return -1;
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

extern "C" int sp_gather_fwd_f32(const float* x, const int32_t* idx, int B, int C, int N, int R,
                                 int k, int cab, float* sp_cube, float* cabins,
                                 uint16_t* cab_arg, void* stream) {
    using namespace spk;
    if (B < 0 || C < 1 || N < 1 || R < 1 || k < 1 || k > N)
        return fail(SPK_E_BADARG, "sp_gather_fwd_f32: need B>=0, C,R>=1, 1<=k<=N (B=%d C=%d N=%d R=%d k=%d)", B, C, N, R, k);
    if (B == 0) return SPK_OK;
    if (!x || !idx || !sp_cube) return fail(SPK_E_BADARG, "sp_gather_fwd_f32: null x/idx/sp_cube");
    const bool want_cab = cabins != nullptr || cab_arg != nullptr;
    if (want_cab) {
        if (!cabins || !cab_arg) return fail(SPK_E_BADARG, "sp_gather_fwd_f32: cabins and cab_arg must both be given or both NULL");
        if (cab < 1 || k < cab) return fail(SPK_E_BADARG, "sp_gather_fwd_f32: need 1 <= cab <= k (cab=%d k=%d)", cab, k);
        if (k / cab > 65535) return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: window length k/cab=%d > 65535", k / cab);
    }
    if (N > 65536) return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: N=%d > 65536", N);
    if ((long long)B * C >= (1LL << 31)) return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: B*C >= 2^31");
    const long long RK = (long long)R * k;
    const size_t budget = (size_t)max_optin_smem();
    const size_t fixed = 128 + (((size_t)RK * 2 + 127) & ~(size_t)127);
    const size_t row_bytes = (size_t)N * 4;
    if (fixed + 2 * row_bytes > budget)
        return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: N=%d, R*k=%lld do not fit shared memory", N, RK);

    GatherFwdParams p;
    p.x = x; p.idx = idx; p.sp_cube = sp_cube; p.cabins = cabins; p.cab_arg = cab_arg;
    p.rows = (long long)B * C;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = want_cab ? cab : 0;
    p.bulk_ok = ((N & 3) == 0) && (((uintptr_t)x & 15) == 0);
    p.vec4 = ((k & 3) == 0) && (((uintptr_t)sp_cube & 15) == 0);
    const int wl = want_cab ? k / cab : 0;
    const int G = wl >> 2;
    p.cab_fast = want_cab && p.vec4 && (k % cab == 0) && (wl % 4 == 0) && G <= 32 && ilog2_exact(G) >= 0 &&
                 (G == 1 || ((RK >> 2) % 32) == 0);
    p.g_shift = p.cab_fast ? ilog2_exact(G) : 0;
    p.q_shift = ilog2_exact((int)(RK >> 2));
    // T rows per tile: two tiles of ~32 KB each -> 3 CTAs (24 warps) per SM
    int T = (int)std::max<size_t>(1, std::min<size_t>(16, (32 * 1024) / row_bytes));
    while (T > 1 && fixed + 2 * (size_t)T * row_bytes > budget) --T;
    T = std::min(T, C);
    p.T = T;
    const size_t smem = fixed + 2 * (size_t)T * row_bytes;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = occupancy_slots((const void*)sp_gather_fwd_kernel, G_THREADS, smem, 0);
    grid = std::min<long long>(grid, (p.rows + T - 1) / T);
    sp_gather_fwd_kernel<<<(int)grid, G_THREADS, smem, (cudaStream_t)stream>>>(p);
    SPK_LAUNCH_CHECK("sp_gather_fwd_kernel");
    if (want_cab && !p.cab_fast) {
        const long long n_rows = (long long)B * C * R;
        const long long total = n_rows * cab;
        const int gsz = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
        sp_cabins_generic_kernel<<<gsz, 256, 0, (cudaStream_t)stream>>>(sp_cube, k, cab, n_rows, cabins, cab_arg);
        SPK_LAUNCH_CHECK("sp_cabins_generic_kernel");
    }
    return SPK_OK;
}

static int gather_bwd_push(const float* g_cube, const float* g_cabins, const int32_t* idx,
                           const uint16_t* cab_arg, int B, int C, int N, int R, int k, int cab,
                           float* grad_x, cudaStream_t stream) {
    using namespace spk;
    const long long RK = (long long)R * k;
    GatherBwdPushParams p;
    p.g_cube = g_cube; p.g_cabins = g_cabins; p.idx = idx; p.cab_arg = cab_arg; p.grad_x = grad_x;
    p.rows = (long long)B * C;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = g_cabins ? cab : 1;
    p.bulk_out = ((N & 3) == 0) && (((uintptr_t)grad_x & 15) == 0);
    p.bulk_in = ((RK & 3) == 0) && (((uintptr_t)g_cube & 15) == 0);
    const size_t budget = (size_t)max_optin_smem();
    const size_t fixed = 128 + (((size_t)RK * 2 + 127) & ~(size_t)127);
    size_t per_T = 2 * ((size_t)N + (size_t)RK) * 4;                // double-buffered acc + g per row
    p.k_shift = ilog2_exact(k);
    p.nbuf = 2; p.g_direct = 0;
    if (fixed + per_T > budget) { per_T /= 2; p.nbuf = 1; }
    if (fixed + per_T > budget) { per_T = (size_t)N * 4; p.g_direct = 1; }
    if (fixed + per_T > budget)
        return fail(SPK_E_UNSUPPORTED, "sp_gather_bwd_f32: N=%d, R*k=%lld do not fit shared memory", N, RK);
    int T = (int)std::max<size_t>(1, std::min<size_t>(8, (72 * 1024 - std::min<size_t>(fixed, 72 * 1024)) / per_T));
    T = std::min(T, C);
    p.T = T;
    const size_t smem = fixed + (size_t)T * per_T;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_bwd_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = occupancy_slots((const void*)sp_gather_bwd_push_kernel, 256, smem, 0);
    grid = std::min<long long>(grid, (p.rows + T - 1) / T);
    sp_gather_bwd_push_kernel<<<(int)grid, 256, smem, stream>>>(p);
    SPK_LAUNCH_CHECK("sp_gather_bwd_push_kernel");
    return SPK_OK;
}

extern "C" int sp_gather_bwd_f32(const float* g_cube, const float* g_cabins, const int32_t* idx,
                                 const uint16_t* cab_arg, int B, int C, int N, int R, int k,
                                 int cab, float* grad_x, void* stream) {
    using namespace spk;
    if (B < 0 || C < 1 || N < 1 || R < 1 || k < 1 || k > N)
        return fail(SPK_E_BADARG, "sp_gather_bwd_f32: need B>=0, C,R>=1, 1<=k<=N (B=%d C=%d N=%d R=%d k=%d)", B, C, N, R, k);
    if (B == 0) return SPK_OK;
    if (!g_cube || !idx || !grad_x) return fail(SPK_E_BADARG, "sp_gather_bwd_f32: null g_cube/idx/grad_x");
    if (g_cabins != nullptr) {
        if (!cab_arg) return fail(SPK_E_BADARG, "sp_gather_bwd_f32: g_cabins given without cab_arg");
        if (cab < 1 || k < cab) return fail(SPK_E_BADARG, "sp_gather_bwd_f32: need 1 <= cab <= k");
    }
    if (N > 65536) return fail(SPK_E_UNSUPPORTED, "sp_gather_bwd_f32: N=%d > 65536", N);
    if ((long long)B * C >= (1LL << 31)) return fail(SPK_E_UNSUPPORTED, "sp_gather_bwd_f32: B*C >= 2^31");
    const long long RK = (long long)R * k;
    // rank-bucketed kernel: slot list (u32[RK]) + two gradient tiles + two accumulator tiles in shared memory
    const size_t budget = (size_t)max_optin_smem();
    const size_t fixed = 128 + (((size_t)RK * 4 + 127) & ~(size_t)127);
    const size_t per_T = 2 * ((size_t)RK + (size_t)N) * 4;
    const size_t scratch = (size_t)RK * 4 + (size_t)N * 2;            // build scratch aliased onto the accumulators
    if (RK > 65535 || R > 63 || fixed + per_T > budget || 2 * (size_t)N * 4 < scratch)
        return gather_bwd_push(g_cube, g_cabins, idx, cab_arg, B, C, N, R, k, cab, grad_x, (cudaStream_t)stream);
    GatherBwdParams p;
    p.g_cube = g_cube; p.g_cabins = g_cabins; p.idx = idx; p.cab_arg = cab_arg; p.grad_x = grad_x;
    p.rows = (long long)B * C;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = g_cabins ? cab : 1;
    p.bulk_in = ((RK & 3) == 0) && (((uintptr_t)g_cube & 15) == 0);
    p.bulk_out = ((N & 3) == 0) && (((uintptr_t)grad_x & 15) == 0);
    int T = 8;                                                       // ~72 KB per CTA -> 3 CTAs (24 warps) per SM
    while (T > 1 && fixed + (size_t)T * per_T > 74 * 1024) T >>= 1;
    while (T > 1 && T > C) T >>= 1;
    p.T = T;
    const size_t smem = fixed + (size_t)T * per_T;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = occupancy_slots((const void*)sp_gather_bwd_kernel, G_THREADS, smem, 0);
    grid = std::min<long long>(grid, (p.rows + T - 1) / T);
    sp_gather_bwd_kernel<<<(int)grid, G_THREADS, smem, (cudaStream_t)stream>>>(p);
    SPK_LAUNCH_CHECK("sp_gather_bwd_kernel");
    return SPK_OK;
}

extern "C" int sp_cabins_fwd_f32(const float* windows, long long rows, int k, int cab, float* cabins,
                                 uint16_t* cab_arg, void* stream) {
    using namespace spk;
    if (rows < 0 || cab < 1 || k < cab) return fail(SPK_E_BADARG, "sp_cabins_fwd_f32: need rows>=0, 1<=cab<=k (k=%d cab=%d)", k, cab);
    if (k / cab > 65535) return fail(SPK_E_UNSUPPORTED, "sp_cabins_fwd_f32: window length %d > 65535", k / cab);
    if (rows == 0) return SPK_OK;
    if (!windows || !cabins || !cab_arg) return fail(SPK_E_BADARG, "sp_cabins_fwd_f32: null pointer");
    const long long total = rows * cab;
    const int g = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
    sp_cabins_generic_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(windows, k, cab, rows, cabins, cab_arg);
    SPK_LAUNCH_CHECK("sp_cabins_generic_kernel");
    return SPK_OK;
}

extern "C" int sp_cabins_bwd_f32(const float* g_cabins, const uint16_t* cab_arg, long long rows, int k,
                                 int cab, float* g_windows, void* stream) {
    using namespace spk;
    if (rows < 0 || cab < 1 || k < cab) return fail(SPK_E_BADARG, "sp_cabins_bwd_f32: need rows>=0, 1<=cab<=k");
    if (rows == 0) return SPK_OK;
    if (!g_cabins || !cab_arg || !g_windows) return fail(SPK_E_BADARG, "sp_cabins_bwd_f32: null pointer");
    const long long total = rows * k;
    const int g = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
    sp_cabins_bwd_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(g_cabins, cab_arg, k, cab, total, g_windows);
    SPK_LAUNCH_CHECK("sp_cabins_bwd_kernel");
    return SPK_OK;
}
