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
//   * backward (default, "pull"): per sample an inverse table (first slot of every point + links between the
//     slots of one point) is built once in shared memory; the upstream gradient tile arrives by bulk copy;
//     every thread sums the contributions of 4 consecutive points in registers (ascending regions = fixed
//     summation order, no atomics) and writes them with one 128-bit streaming store per row.  grad_x is
//     written exactly once (zeros included): no memset, no shared-memory accumulator;
//   * backward ("push", for R*k >= 65535 or SPK_BWD=push): one region per phase is scattered into a T x N
//     shared-memory accumulator (points are unique inside a region: no conflicts), the finished rows leave
//     with one TMA bulk store (cp.async.bulk.global.shared::cta).  Bit-identical to the pull kernel.
#include "spk_common.cuh"
#include <stdlib.h>

namespace spk {


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
template <int G_THREADS>
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
    int cab_wide;         // windows of exactly 256 slots (k/cab == 256, e.g. N = 16384, k = 2048): a work item = 8 slots, a window = one warp
    int q_shift;          // log2(R*k/4) when that is a power of two, else -1
    int q_cols;           // dense case: R*k/4 is a multiple of the CTA width -> a thread keeps its 4-slot groups for all rows of a tile
};

template <int G_THREADS>
__global__ void __launch_bounds__(G_THREADS)
sp_gather_fwd_kernel(const GatherFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int RK = p.R * p.k, N = p.N, T = p.T, C = p.C;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                            // 2
    uint16_t* idx_s = reinterpret_cast<uint16_t*>(smem_raw + 128);                     // RK
    float* tiles = reinterpret_cast<float*>(smem_raw + 128 + ((RK * 2 + 127) & ~127)); // 2 * T*N
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_mbar_init(); }
    pdl_trigger();
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
    // The first two x tiles are in flight; only now wait for the kernel before this one (the top-k, which produces idx
    // and lets its dependents start early -- after ITS wait, so x was complete before either kernel began).
    pdl_wait();

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
            stage_idx_u16<G_THREADS>(p.idx + (size_t)b * RK, idx_s, RK, tid);   // passed the barrier ending the last tile)
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
        if (p.q_cols) {
            // Dense gather (R*k/4 a multiple of the CTA width, e.g. the reference's operating point R*k = N): a thread
            // owns the same 4-slot groups in every row of the tile, so the indices, the window bookkeeping and the
            // output offsets are formed once per group instead of once per (row, group).
            for (int q = tid; q < Q; q += G_THREADS) {
                const uint2 u = *reinterpret_cast<const uint2*>(idx_s + 4 * q);
                const int i0 = u.x & 0xFFFFu, i1 = u.x >> 16, i2 = u.y & 0xFFFFu, i3 = u.y >> 16;
                const uint32_t lane_off = (uint32_t)(q & (G - 1)) << 2;            // first slot of this group inside its window
                const int win = p.cab_fast ? (q >> p.g_shift) : 0;                 // (r, w) flattened
                const bool leader = (q & (G - 1)) == 0;
                float4* outp = reinterpret_cast<float4*>(p.sp_cube + (size_t)g * RK) + q;
                const float* row = tile;
                size_t cab_o = (size_t)g * wins_per_row + win;
                for (int t = 0; t < rows; ++t, row += N, outp += (RK >> 2), cab_o += wins_per_row) {
                    float4 v;
                    v.x = row[i0]; v.y = row[i1]; v.z = row[i2]; v.w = row[i3];
                    st_cs_f4(outp, v);
                    if (p.cab_fast) {
                        // torch.max over the window: first maximum, a NaN wins (first NaN); -0 == +0
                        float best = v.x; uint32_t off = 0;
                        if (!(v.y <= best) && best == best) { best = v.y; off = 1; }
                        if (!(v.z <= best) && best == best) { best = v.z; off = 2; }
                        if (!(v.w <= best) && best == best) { best = v.w; off = 3; }
                        uint32_t woff = off + lane_off;
                        if (G >= 16) {
                            // two warp reductions (REDUX) on the window's own lanes: the greatest key, then the lowest offset holding it
                            // (measured: G = 16, N = 4096: 62 -> 51 us; G = 8, k = 256: 32.5 -> 35 us, so the butterflies stay there)
                            const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((tid & 31) & ~(G - 1)));
                            const uint32_t key = order_key(best);
                            const uint32_t kmax = __reduce_max_sync(gmask, key);
                            woff = __reduce_min_sync(gmask, key == kmax ? woff : 0xFFFFu);
                        } else if (G > 1) {
                            // two 32-bit butterflies: the window's greatest key, then the lowest offset that holds it
                            const uint32_t key = order_key(best);
                            uint32_t kmax = key;
                            for (int d = 1; d < G; d <<= 1) kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
                            woff = (key == kmax) ? woff : 0xFFFFu;
                            for (int d = 1; d < G; d <<= 1) woff = min(woff, __shfl_xor_sync(0xFFFFFFFFu, woff, d));
                        }
                        if (leader) {
                            // the value itself (sign of zero, NaN payload) comes from the winning slot
                            p.cabins[cab_o] = (G == 1) ? best : row[idx_s[win * wl + (int)woff]];
                            p.cab_arg[cab_o] = (uint16_t)woff;
                        }
                    }
                }
            }
        } else if (p.cab_wide) {
            // 256-slot windows: an item = 8 consecutive slots (two 128-bit stores), the 32 lanes of a warp = one window; the
            // window max is two warp reductions as above.  (Without this the window max ran as a second kernel that re-read
            // sp_cube: N = 16384 forward 328 us.)
            const int Q8 = RK >> 3;                                                   // items per row (a multiple of 32)
            const int items = rows * Q8;                                              // whole warps: live is warp-uniform
            for (int e = tid; e < items; e += G_THREADS) {
                const int t = e / Q8, q8 = e - t * Q8;
                const float* row = tile + (size_t)t * N;
                const uint4 u = *reinterpret_cast<const uint4*>(idx_s + 8 * q8);
                float v[8];
                v[0] = row[u.x & 0xFFFFu]; v[1] = row[u.x >> 16]; v[2] = row[u.y & 0xFFFFu]; v[3] = row[u.y >> 16];
                v[4] = row[u.z & 0xFFFFu]; v[5] = row[u.z >> 16]; v[6] = row[u.w & 0xFFFFu]; v[7] = row[u.w >> 16];
                float4* outp = reinterpret_cast<float4*>(p.sp_cube + (size_t)(g + t) * RK) + 2 * q8;
                st_cs_f4(outp, make_float4(v[0], v[1], v[2], v[3]));
                st_cs_f4(outp + 1, make_float4(v[4], v[5], v[6], v[7]));
                // torch.max over the window: first maximum, a NaN wins (first NaN); -0 == +0
                float best = v[0]; uint32_t off = 0;
#pragma unroll
                for (int j = 1; j < 8; ++j)
                    if (!(v[j] <= best) && best == best) { best = v[j]; off = (uint32_t)j; }
                uint32_t woff = off + ((uint32_t)(q8 & 31) << 3);
                const uint32_t key = order_key(best);
                const uint32_t kmax = __reduce_max_sync(0xFFFFFFFFu, key);
                woff = __reduce_min_sync(0xFFFFFFFFu, key == kmax ? woff : 0xFFFFFFFFu);
                if ((q8 & 31) == 0) {
                    const int win = q8 >> 5;                                          // (r, w) flattened
                    const size_t o = (size_t)(g + t) * wins_per_row + win;
                    p.cabins[o] = row[idx_s[win * wl + (int)woff]];                   // the value itself comes from the winning slot
                    p.cab_arg[o] = (uint16_t)woff;
                }
            }
        } else if (p.vec4) {
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
                    if (G >= 16) {
                        // 16..32 lanes per window (a whole warp at N = 8192, k = 1024): two warp reductions (REDUX) -- the greatest key, then
                        // the lowest offset holding it -- instead of five 64-bit shuffle steps (the forward was issue-bound there:
                        // 109 M warp instructions, issue 73 %)
                        // (G = 16: the same on the window's own lanes -- disjoint member masks inside a warp)
                        const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << ((tid & 31) & ~(G - 1)));
                        const uint32_t key = order_key(best);
                        const uint32_t kmax = __reduce_max_sync(gmask, key);
                        woff = __reduce_min_sync(gmask, key == kmax ? woff : 0xFFFFFFFFu);
                    } else if (G > 1) {
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
    pdl_tail_trigger_bit<0>();
}

// Generic window max over an already written sp_cube (any k, cab): one thread per window.
__global__ void sp_cabins_generic_kernel(const float* __restrict__ sp_cube, int k, int cab,
                                         long long n_rows /* B*C*R */, float* __restrict__ cabins,
                                         uint16_t* __restrict__ cab_arg) {
    pdl_trigger();
    pdl_wait();
    const int wl = k / cab;
    const long long total = n_rows * cab;
    if (wl >= 32) {
        // long windows (e.g. N = 16384: 256 slots): a WARP per window -- coalesced reads, the greatest key and then the
        // lowest offset holding it by shuffles (one thread per window walked 256 strided floats: 500 us at N = 16384)
        const int lane = threadIdx.x & 31;
        const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
        for (long long t = warp0; t < total; t += nwarp) {
            const long long rowi = t / cab;
            const int w = (int)(t - rowi * cab);
            const float* src = sp_cube + rowi * k + (size_t)w * wl;
            uint32_t bk = 0; int bo = 0x7FFFFFFF;
            for (int j = lane; j < wl; j += 32) {
                const uint32_t kk = order_key(src[j]);
                if (kk > bk || bo == 0x7FFFFFFF) { bk = kk; bo = j; }     // strict: the first maximum of this lane's slots
            }
            uint32_t kmax = bk;
            for (int d = 16; d > 0; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xFFFFFFFFu, kmax, d));
            int off = (bk == kmax) ? bo : 0x7FFFFFFF;
            for (int d = 16; d > 0; d >>= 1) off = min(off, __shfl_xor_sync(0xFFFFFFFFu, off, d));
            if (lane == 0) { cabins[t] = src[off]; cab_arg[t] = (uint16_t)off; }      // the value itself (sign of zero, NaN payload) from the winning slot
        }
        return;
    }
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
    pdl_trigger();
    pdl_wait();
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
// backward: region-ordered scatter into a shared-memory accumulator, TMA bulk store of the rows.
// grad_x[b,c,n] = sum over the slots (r,j) that selected point n of g[b,c,r,j].  Points are unique
// inside a region, so one region is scattered per phase without conflicts or atomics; regions in
// ascending order fix the summation order -> deterministic.  (Tried and measured slower on B200,
// see profiles/README.md: a CSR "pull" formulation and a rank-bucketed scatter.)
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
    int ng;           // gradient ring depth (2..4) when nbuf == 2: tiles are fetched ng-1 groups ahead
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
    const int NG = (nbuf == 2) ? p.ng : 1;                                       // gradient ring depth
    const size_t gstride = ((size_t)T * RK + 3) & ~(size_t)3, astride = ((size_t)T * N + 3) & ~(size_t)3;   // 16-byte multiples
    float* acc = gbuf + (p.g_direct ? 0 : (size_t)NG * gstride);                 // nbuf * T*N
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    pdl_trigger();
    __syncthreads();
    pdl_wait();

    const long long g_lo = p.rows * blockIdx.x / gridDim.x;
    const long long g_hi = p.rows * (blockIdx.x + 1) / gridDim.x;
    const int wl = (p.g_cabins != nullptr) ? k / p.cab : 1;
    const int wins = R * p.cab;

    // group i covers rows [g, g + rows_i): never crosses a sample boundary
    auto group_rows = [&](long long g) -> int { return tile_rows(g, g_hi, T, C); };
    auto issue_load = [&](long long g, int rows, int buf) {        // tid 0 only
        if (p.bulk_in && !p.g_direct) {
            const uint32_t bytes = (uint32_t)rows * RK * 4u;
            mbar_expect_tx(&bars[buf], bytes);
            bulk_g2s(gbuf + (size_t)buf * gstride, p.g_cube + (size_t)g * RK, bytes, &bars[buf]);
        }
    };

    long long g = g_lo, g_pref = g_lo;
    int cur_b = -1, gi_pref = 0;
    // prologue: the first NG-1 gradient tiles (single-buffered: just the first)
    for (; gi_pref < (NG > 1 ? NG - 1 : 1) && g_pref < g_hi; ++gi_pref) {
        const int r = group_rows(g_pref);
        if (tid == 0) issue_load(g_pref, r, gi_pref % NG);
        g_pref += r;
    }
    for (int gi = 0; g < g_hi; ++gi) {
        const int buf = (nbuf == 2) ? (gi & 1) : 0;
        const int gslot = gi % NG;
        const int rows = group_rows(g);
        const long long g_next = g + rows;
        float* gs = gbuf + (size_t)gslot * gstride;
        float* ac = acc + (size_t)buf * astride;
        // keep NG-1 gradient tiles in flight: the slot refilled here was consumed by group gi-1, whose
        // readers finished before the barrier that ended that iteration
        if (NG > 1 && g_pref < g_hi) {
            const int r = group_rows(g_pref);
            if (tid == 0) issue_load(g_pref, r, gi_pref % NG);
            g_pref += r; ++gi_pref;
        }
        // acc[buf] was handed to the TMA store nbuf groups ago: wait until that store has READ it
        if (p.bulk_out && tid == 0) { if (nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
        const int b = (int)((unsigned)g / (unsigned)C);
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
            mbar_wait(&bars[gslot], (uint32_t)((gi / NG) & 1));
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
            // the window-max fold wrote the staged gradient tile with generic stores and the next bulk copy refills that
            // slot through the async proxy: order the two (the bulk_out branch above fences already)
            if (p.bulk_in && p.g_cabins != nullptr) fence_proxy_async_smem();
            for (int i = tid; i < rows * N; i += nthr) dst[i] = ac[i];
            __syncthreads();
        }
        // single-buffered: the next group's gradient can only be fetched once this one is consumed
        if (NG == 1 && g_pref < g_hi) {
            const int r = group_rows(g_pref);
            if (tid == 0) issue_load(g_pref, r, 0);
            g_pref += r; ++gi_pref;
        }
        g = g_next;
    }
    pdl_tail_trigger();
    if (p.bulk_out && tid == 0) bulk_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// backward, "pull" formulation: no shared-memory accumulator, no zero fill, no per-region barriers.
// Per sample the CTA builds an inverse table once (shared by all C rows of the sample):
//   first[n]   = slot (r*k+j) of the LOWEST region that selected point n, or NONE
//   link[slot] = slot of the next higher region that selected the same point, or NONE
// (built in descending region order by prepending; points are unique inside a region, so one
// region per phase needs no atomics).  Then every thread owns 4 consecutive points of T rows:
// it walks the (mostly 0- or 1-element) chains, sums the staged upstream gradient in ascending
// region order (same order as the push kernel and the oracle) and writes the finished values
// with one 128-bit streaming store per row -- grad_x is written exactly once, straight from
// registers.  The gradient tiles arrive by bulk copy through a ring, the window-max gradient
// is folded into the winning slot of the staged tile.
// ---------------------------------------------------------------------------------------------
struct GatherBwdPullParams {
    const float* g_cube; const float* g_cabins; const int32_t* idx; const uint16_t* cab_arg;
    float* grad_x;
    long long rows;   // B*C
    int C, N, R, k, cab;
    int bulk_in;      // (R*k)%4==0 and g_cube 16B aligned -> g rows arrive with cp.async.bulk
    int cab_bulk;     // window-max gradient + arg-max ride along with the tile (bulk copies)
    int vec_out;      // N%4==0 and grad_x 16B aligned -> 128-bit stores
    int ng;           // ring depth (>= 2)
    int dbg;          // timing experiments only (results wrong): 1 no table build, 2 no window-max fold, 4 no gradient tiles
};

constexpr uint32_t PULL_NONE = 0xFFFFu;

// bytes of one ring slot: gradient tile + (optionally) the group's g_cabins (f32) and cab_arg (u16)
static __host__ __device__ inline size_t pull_slot_bytes(int T, int RK, int wins, bool cab_bulk) {
    size_t b = (((size_t)T * RK + 3) & ~(size_t)3) * 4;
    if (cab_bulk) b += (size_t)T * wins * 4 + (((size_t)T * wins * 2 + 15) & ~(size_t)15);
    return b;
}

template <int T, int NT>
__global__ void __launch_bounds__(NT)
sp_gather_bwd_pull_kernel(const GatherBwdPullParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    constexpr int nthr = NT;
    const int N = p.N, R = p.R, k = p.k, C = p.C;
    const int RK = R * k;
    const int NG = p.ng;
    const bool has_cab = p.g_cabins != nullptr && !(p.dbg & 2);
    const int wl = has_cab ? k / p.cab : 1;
    const int wins = R * p.cab;
    const bool cab_bulk = has_cab && p.cab_bulk;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                              // <= 8
    uint16_t* first = reinterpret_cast<uint16_t*>(smem_raw + 128);                       // N (padded to 4)
    const size_t first_bytes = ((((size_t)N + 3) & ~(size_t)3) * 2 + 127) & ~(size_t)127;
    uint16_t* link = reinterpret_cast<uint16_t*>(smem_raw + 128 + first_bytes);          // RK
    const size_t link_bytes = ((size_t)RK * 2 + 127) & ~(size_t)127;
    unsigned char* ring = smem_raw + 128 + first_bytes + link_bytes;                     // NG slots
    const size_t slot_bytes = pull_slot_bytes(T, RK, wins, p.cab_bulk && p.g_cabins != nullptr);
    const size_t tile_bytes = (((size_t)T * RK + 3) & ~(size_t)3) * 4;
    if (tid == 0) { for (int i = 0; i < NG; ++i) mbar_init(&bars[i], 1); fence_mbar_init(); }
    pdl_trigger();
    __syncthreads();
    pdl_wait();

    const long long g_lo = p.rows * blockIdx.x / gridDim.x;
    const long long g_hi = p.rows * (blockIdx.x + 1) / gridDim.x;

    auto group_rows = [&](long long g) -> int { return tile_rows(g, g_hi, T, C); };
    auto issue_load = [&](long long g, int rows, int slot) {       // tid 0 only
        unsigned char* sl = ring + (size_t)slot * slot_bytes;
        uint32_t bytes = 0;
        const uint32_t gb = (uint32_t)rows * RK * 4u, cb = (uint32_t)rows * wins * 4u, ab = (uint32_t)rows * wins * 2u;
        if (p.bulk_in && !(p.dbg & 4)) bytes += gb;
        if (cab_bulk) bytes += cb + ab;
        if (bytes == 0) return;
        mbar_expect_tx(&bars[slot], bytes);
        if (p.bulk_in && !(p.dbg & 4)) bulk_g2s(sl, p.g_cube + (size_t)g * RK, gb, &bars[slot]);
        if (cab_bulk) {
            bulk_g2s(sl + tile_bytes, p.g_cabins + (size_t)g * wins, cb, &bars[slot]);
            bulk_g2s(sl + tile_bytes + (size_t)T * wins * 4, p.cab_arg + (size_t)g * wins, ab, &bars[slot]);
        }
    };
    const bool any_bulk = (p.bulk_in && !(p.dbg & 4)) || cab_bulk;

    // destination (inside the staged tile, before the arg-max offset) of the window-max terms this
    // thread folds: entries e = tid and tid + nthr of the group's rows*wins list -- group-invariant
    int cbase[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int e = tid + u * nthr;
        const int t = e / wins, rw = e - t * wins;
        const int r = rw / p.cab, w = rw - r * p.cab;
        cbase[u] = t * RK + r * k + w * wl;
    }

    long long g = g_lo, g_pref = g_lo;
    int cur_b = -1, gi_pref = 0;
    for (; gi_pref < NG - 1 && g_pref < g_hi; ++gi_pref) {
        const int r = group_rows(g_pref);
        if (tid == 0) issue_load(g_pref, r, gi_pref % NG);
        g_pref += r;
    }
    for (int gi = 0; g < g_hi; ++gi) {
        const int gslot = gi % NG;
        const int rows = group_rows(g);
        unsigned char* sl = ring + (size_t)gslot * slot_bytes;
        float* gs = reinterpret_cast<float*>(sl);
        // keep NG-1 groups in flight: the slot refilled here was consumed by group gi-1, whose readers
        // passed the barrier that ended that iteration
        if (g_pref < g_hi) {
            const int r = group_rows(g_pref);
            if (tid == 0) issue_load(g_pref, r, gi_pref % NG);
            g_pref += r; ++gi_pref;
        }
        // window-max gradient without the bulk path: fetched now, folded once the tile has landed
        float cg[2] = {0.f, 0.f}; int coff[2] = {-1, -1};
        if (has_cab && !cab_bulk) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int e = tid + u * nthr;
                if (e < rows * wins) {
                    const size_t o = (size_t)g * wins + e;
                    coff[u] = (int)__ldg(p.cab_arg + o);
                    cg[u] = __ldg(p.g_cabins + o);
                }
            }
        }
        const int b = (int)((unsigned)g / (unsigned)C);
        if (b != cur_b) {                                          // new sample: rebuild the inverse table
            cur_b = b;                                             // (readers of the old one passed the last barrier)
            const int32_t* ib = p.idx + (size_t)b * RK;
            for (int i = tid; i < RK; i += nthr) link[i] = (uint16_t)__ldg(ib + i);
            {
                uint2* f2 = reinterpret_cast<uint2*>(first);
                for (int i = tid; i < ((N + 3) >> 2); i += nthr) f2[i] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
            }
            __syncthreads();
            for (int r = (p.dbg & 1) ? -1 : R - 1; r >= 0; --r) {  // descending: chains come out ascending
                for (int j = tid; j < k; j += nthr) {
                    const int s = r * k + j;
                    const int n = link[s];                         // still the point index of slot s
                    link[s] = first[n];
                    first[n] = (uint16_t)s;
                }
                __syncthreads();
            }
        }
        if (any_bulk) mbar_wait(&bars[gslot], (uint32_t)((gi / NG) & 1));
        if (!p.bulk_in && !(p.dbg & 4)) {
            const float* src = p.g_cube + (size_t)g * RK;
            for (int i = tid; i < rows * RK; i += nthr) gs[i] = __ldg(src + i);
            __syncthreads();
        }
        if (has_cab) {                                             // unique destinations: plain adds
            if (cab_bulk) {
                const float* cgs = reinterpret_cast<const float*>(sl + tile_bytes);
                const uint16_t* cas = reinterpret_cast<const uint16_t*>(sl + tile_bytes + (size_t)T * wins * 4);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int e = tid + u * nthr;
                    if (e < rows * wins) gs[cbase[u] + (int)cas[e]] += cgs[e];
                }
                for (int e = tid + 2 * nthr; e < rows * wins; e += nthr) {
                    const int t = e / wins, rw = e - t * wins;
                    const int r = rw / p.cab, w = rw - r * p.cab;
                    gs[t * RK + r * k + w * wl + (int)cas[e]] += cgs[e];
                }
            } else {
                if (coff[0] >= 0) gs[cbase[0] + coff[0]] += cg[0];
                if (coff[1] >= 0) gs[cbase[1] + coff[1]] += cg[1];
                for (int e = tid + 2 * nthr; e < rows * wins; e += nthr) {
                    const int t = e / wins, rw = e - t * wins;
                    const int r = rw / p.cab, w = rw - r * p.cab;
                    const size_t o = (size_t)g * wins + e;
                    gs[t * RK + r * k + w * wl + (int)__ldg(p.cab_arg + o)] += __ldg(p.g_cabins + o);
                }
            }
            __syncthreads();
        }
        float* dst = p.grad_x + (size_t)g * N;
        if (p.vec_out) {
            const int NQ = N >> 2;
            for (int q = tid; q < NQ; q += nthr) {
                const uint2 f = *reinterpret_cast<const uint2*>(first + 4 * q);
                float acc[4][T];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int t = 0; t < T; ++t) acc[c][t] = 0.f;
                if ((f.x & f.y) != 0xFFFFFFFFu) {                  // at least one of the 4 points was selected
                    const uint32_t s4[4] = {f.x & 0xFFFFu, f.x >> 16, f.y & 0xFFFFu, f.y >> 16};
                    uint32_t nx[4];
                    // lowest region of every point: straight-line, all loads independent (slot 0 stands in
                    // for "none" so that nothing is predicated)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const bool on = s4[c] != PULL_NONE;
                        const uint32_t sc = on ? s4[c] : 0u;
                        const uint32_t l = link[sc];
                        nx[c] = on ? l : PULL_NONE;
#pragma unroll
                        for (int t = 0; t < T; ++t) {              // rows beyond `rows` read stale tile bytes: never stored
                            const float v = gs[t * RK + sc];
                            acc[c][t] = on ? 0.f + v : 0.f;        // 0 + v: the oracle's accumulator starts at +0
                        }
                    }
                    // points selected by more than one region (rare): ascending regions, in lockstep
                    while ((nx[0] & nx[1] & nx[2] & nx[3]) != PULL_NONE) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (nx[c] != PULL_NONE) {
#pragma unroll
                                for (int t = 0; t < T; ++t) acc[c][t] += gs[t * RK + nx[c]];
                                nx[c] = link[nx[c]];
                            }
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < T; ++t)
                    if (t < rows)
                        st_cs_f4(reinterpret_cast<float4*>(dst + (size_t)t * N) + q,
                                 make_float4(acc[0][t], acc[1][t], acc[2][t], acc[3][t]));
            }
        } else {
            for (int n = tid; n < N; n += nthr) {
                float acc[T];
#pragma unroll
                for (int t = 0; t < T; ++t) acc[t] = 0.f;
                uint32_t s = first[n];
                while (s != PULL_NONE) {
#pragma unroll
                    for (int t = 0; t < T; ++t) acc[t] += gs[t * RK + s];
                    s = link[s];
                }
#pragma unroll
                for (int t = 0; t < T; ++t)
                    if (t < rows) dst[(size_t)t * N + n] = acc[t];
            }
        }
        if (has_cab) fence_proxy_async_smem();                     // the fold wrote gs with generic stores; TMA refills it
        __syncthreads();                                           // the slot and the table are free again
        g += rows;
    }
    pdl_tail_trigger_bit<1>();
}

}  // namespace spk

// Tuning / timing knobs (SPK_FWD_*, SPK_PULL_*, SPK_BWD_SMEM_KB, SPK_BWD_RING): read ONLY in builds made with
// SPK_NVCC_EXTRA=-DSPK_EXPERIMENT -- a stray environment variable cannot change how the shipped library runs, and
// SPK_PULL_DBG (which skips work and so returns WRONG results, for timing only) does not exist in it at all.
static const char* tune_env(const char* name) {
#ifdef SPK_EXPERIMENT
    return getenv(name);
#else
    (void)name;
    return nullptr;
#endif
}

// resident CTAs of a persistent kernel; the occupancy query is cached per (kernel, width, shared memory, device)
static int occupancy_slots(const void* kernel, int threads, size_t smem, int cap_per_sm) {
    struct Entry { const void* k; int threads; size_t smem; int dev; int per_sm; };
    static thread_local Entry cache[16];
    static thread_local int n_cache = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    int per_sm = 0;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].k == kernel && cache[i].threads == threads && cache[i].smem == smem && cache[i].dev == dev) per_sm = cache[i].per_sm;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cache[n_cache % 16] = Entry{kernel, threads, smem, dev, per_sm};
        ++n_cache; if (n_cache > 16) n_cache = 16;
    }
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
    p.cab_wide = want_cab && !p.cab_fast && p.vec4 && (k % cab == 0) && wl == 256 && ((RK >> 3) % 32) == 0;
    p.q_shift = ilog2_exact((int)(RK >> 2));
    // tile size / CTA width: 256 threads with two ~32 KB tiles (3 CTAs per SM), or -- SPK_FWD_THREADS=128 --
    // narrow CTAs with two ~16 KB tiles (6 CTAs per SM)
    int threads = 256;
    if (const char* e = tune_env("SPK_FWD_THREADS")) threads = atoi(e) <= 128 ? 128 : 256;
    size_t tile_target = threads == 128 ? 16 * 1024 : 32 * 1024;
    if (const char* e = tune_env("SPK_FWD_TILE_KB")) tile_target = (size_t)atoi(e) * 1024;
    int T = (int)std::max<size_t>(1, std::min<size_t>(16, tile_target / row_bytes));
    while (T > 1 && fixed + 2 * (size_t)T * row_bytes > budget) --T;
    T = std::min(T, C);
    p.T = T;
    // rows so long that only ONE CTA fits an SM (two row buffers + index list > half the shared memory, e.g. N = 16384):
    // a 1024-thread CTA keeps the SM busy (no measurable difference at N = 16384 once the window max ran a warp per window; kept for the long-row case)
    if (fixed + 2 * (size_t)T * row_bytes > (size_t)110 * 1024) threads = 1024;
    else if (fixed + 2 * (size_t)T * row_bytes > (size_t)72 * 1024 && !tune_env("SPK_FWD_THREADS")) threads = 512;   // two CTAs per SM: 32 warps
    // the per-group hoisting only pays when a tile holds several rows (measured: N=2048 33 vs 34.5 us; one row per tile, N=8192: 169 vs 145 us)
    p.q_cols = T >= 2 && p.vec4 && ((RK >> 2) % threads) == 0 && (!want_cab || p.cab_fast) && !tune_env("SPK_FWD_NO_QCOLS");
    if (p.q_cols) p.cab_wide = 0;
    const size_t smem = fixed + 2 * (size_t)T * row_bytes;
    auto launch = [&](auto kern, int nt) -> int {
        if (smem > 48 * 1024)
            SPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long grid = occupancy_slots((const void*)kern, nt, smem, 0);
        grid = std::min<long long>(grid, (p.rows + T - 1) / T);
        SPK_CUDA(launch_k(kern, dim3((int)grid), dim3(nt), smem, (cudaStream_t)stream, p));
        return SPK_OK;
    };
    const int rc = threads == 128 ? launch(sp_gather_fwd_kernel<128>, 128) : threads == 1024 ? launch(sp_gather_fwd_kernel<1024>, 1024) :
                   threads == 512 ? launch(sp_gather_fwd_kernel<512>, 512) : launch(sp_gather_fwd_kernel<256>, 256);
    if (rc != SPK_OK) return rc;
    if (want_cab && !p.cab_fast && !p.cab_wide) {
        const long long n_rows = (long long)B * C * R;
        const long long total = n_rows * cab;
        const long long work = (k / cab >= 32) ? total * 32 : total;            // a warp per long window
        const int gsz = (int)std::min<long long>((work + 255) / 256, (long long)sm_count() * 16);
        SPK_CUDA(launch_k(sp_cabins_generic_kernel, dim3(gsz), dim3(256), 0, (cudaStream_t)stream, (const float*)sp_cube, k, cab, n_rows, cabins, cab_arg));
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
    // prefer two resident CTAs per SM over double buffering inside one
    if (fixed + per_T > budget || (fixed + per_T > 110 * 1024 && fixed + per_T / 2 <= 110 * 1024)) { per_T /= 2; p.nbuf = 1; }
    if (fixed + per_T > budget) { per_T = (size_t)N * 4; p.g_direct = 1; }
    if (fixed + per_T + 64 > budget)
        return fail(SPK_E_UNSUPPORTED, "sp_gather_bwd_f32: N=%d, R*k=%lld do not fit shared memory", N, RK);
    p.ng = 2;
    size_t target = 72 * 1024;
    if (const char* e = tune_env("SPK_BWD_SMEM_KB")) target = (size_t)atoi(e) * 1024;      // tuning knob
    int T = (int)std::max<size_t>(1, std::min<size_t>(8, (target - std::min<size_t>(fixed, target)) / per_T));
    T = std::min(T, C);
    p.T = T;
    size_t smem = fixed + (size_t)T * per_T + 64;                   // + padding of the buffer strides to 16 bytes
    if (p.nbuf == 2 && !p.g_direct) {                               // deeper gradient ring while it fits the same budget
        const size_t gtile = (((size_t)T * RK + 3) & ~(size_t)3) * 4;
        // measured on B200: a deeper ring (3-4 tiles in flight) is SLOWER at config A (28.0 vs 23.5 us), so the
        // default stays at 2; SPK_BWD_RING=3|4 re-enables it for experiments
        int want = 2;
        if (const char* e = tune_env("SPK_BWD_RING")) want = std::max(2, std::min(4, atoi(e)));
        while (p.ng < want && smem + gtile <= std::max(target, (size_t)74 * 1024)) { smem += gtile; ++p.ng; }
    }
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_bwd_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = occupancy_slots((const void*)sp_gather_bwd_push_kernel, 256, smem, 0);
    grid = std::min<long long>(grid, (p.rows + T - 1) / T);
    SPK_CUDA(launch_k(sp_gather_bwd_push_kernel, dim3((int)grid), dim3(256), smem, stream, p));
    return SPK_OK;
}


template <int T, int NT>
static int launch_pull(const spk::GatherBwdPullParams& p, size_t smem, cudaStream_t stream) {
    using namespace spk;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_bwd_pull_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = occupancy_slots((const void*)sp_gather_bwd_pull_kernel<T, NT>, NT, smem, 0);
    grid = std::min<long long>(grid, (p.rows + T - 1) / T);
    SPK_CUDA(launch_k(sp_gather_bwd_pull_kernel<T, NT>, dim3((int)grid), dim3(NT), smem, stream, p));
    return SPK_OK;
}

// returns 1 when the sizes do not suit the pull kernel (caller falls back to the push kernel)
static int gather_bwd_pull(const float* g_cube, const float* g_cabins, const int32_t* idx,
                           const uint16_t* cab_arg, int B, int C, int N, int R, int k, int cab,
                           float* grad_x, cudaStream_t stream) {
    using namespace spk;
    const long long RK = (long long)R * k;
    if (RK >= 65535) return 1;                                     // slots are u16, 0xFFFF = none
    GatherBwdPullParams p;
    p.g_cube = g_cube; p.g_cabins = g_cabins; p.idx = idx; p.cab_arg = cab_arg; p.grad_x = grad_x;
    p.rows = (long long)B * C;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = g_cabins ? cab : 1;
    p.bulk_in = ((RK & 3) == 0) && (((uintptr_t)g_cube & 15) == 0);
    p.vec_out = ((N & 3) == 0) && (((uintptr_t)grad_x & 15) == 0);
    const size_t budget = (size_t)max_optin_smem();
    const size_t fixed = 128 + (((((size_t)N + 3) & ~(size_t)3) * 2 + 127) & ~(size_t)127) + (((size_t)RK * 2 + 127) & ~(size_t)127);
    size_t target = 75 * 1024;                                     // three CTAs per SM
    if (const char* e = tune_env("SPK_PULL_SMEM_KB")) target = (size_t)atoi(e) * 1024;
    const int wins = R * p.cab;
    // the window-max gradient rides along as bulk copies when every group's byte ranges are 16-byte multiples
    p.cab_bulk = g_cabins != nullptr && (wins % 8) == 0 && (((uintptr_t)g_cabins & 15) == 0) && (((uintptr_t)cab_arg & 15) == 0);
    if (tune_env("SPK_PULL_NOCABBULK")) p.cab_bulk = 0;
    // T rows per group, by the size of a gradient row (measured, us per call at T / threads; B=32, C=256, R=8, k=N/8):
    //   R*k =  256 (N=2048, k=32): 8/128 14.2 (a deeper ring or wider CTAs are slower)
    //   R*k = 1024 (N=1024):       8/128 26.0, 4/128 23.3
    //   R*k = 2048 (N=2048):       2/128 39.8, 4/128 47.2, 4/256 37.9, 8/256 51.0
    //   R*k = 4096 (N=4096):       1/128 83.5, 2/128 88.1, 2/256 67.9, 4/512 69.8
    //   R*k = 8192 (N=8192):       2/512 140.9, 2/256 156.7, 3/512 150.6
    // i.e. a tile of 16-32 KB, and (below) the CTA width that keeps >= 16 warps on an SM for the CTAs that fit
    int T = RK <= 512 ? 8 : RK <= 2048 ? 4 : 2;
    if (const char* e = tune_env("SPK_PULL_T")) { T = atoi(e); if (T != 1 && T != 2 && T != 4 && T != 8) return fail(SPK_E_BADARG, "SPK_PULL_T must be 1, 2, 4 or 8"); }
    while (T > 1 && fixed + 2 * pull_slot_bytes(T, (int)RK, wins, p.cab_bulk) > budget) T >>= 1;
    T = std::max(1, std::min(8, T));
    while (T > 1 && T / 2 >= C) T >>= 1;
    const size_t tile = pull_slot_bytes(T, (int)RK, wins, p.cab_bulk);
    if (fixed + 2 * tile > budget) return 1;
    int ng = 2;
    if (const char* e = tune_env("SPK_PULL_NG")) ng = atoi(e);
    ng = std::max(2, std::min(8, ng));
    while (ng > 2 && fixed + (size_t)ng * tile > std::max(target, fixed + 2 * tile)) --ng;
    p.ng = ng;
    p.dbg = 0;
#ifdef SPK_EXPERIMENT
    if (const char* e = getenv("SPK_PULL_DBG")) p.dbg = atoi(e);      // timing only: skips work, results WRONG
#endif
    const size_t smem = fixed + (size_t)ng * tile;
    // CTA width: narrow CTAs (128 threads) put more independent CTAs on an SM, so one CTA's table build /
    // tile wait / fold overlaps another's stores (measured at config A: 14.0 us vs 18.6 us with 256);
    // with at most two resident CTAs (large tables / tiles) 512 threads keep enough warps per SM
    const int resident = (int)((228 * 1024) / (smem + 1024));
    int threads = resident <= 1 ? 512 : resident <= 3 ? 256 : 128;
    if (const char* e = tune_env("SPK_PULL_THREADS")) threads = atoi(e);
    if (T == 8 && threads > 256) threads = 256;
#define SPK_PULL_CASE(TT)                                                            \
    case TT:                                                                         \
        if (threads >= 512 && TT < 8) return launch_pull<TT, (TT < 8 ? 512 : 256)>(p, smem, stream); \
        if (threads >= 256) return launch_pull<TT, 256>(p, smem, stream);            \
        if (threads >= 128) return launch_pull<TT, 128>(p, smem, stream);            \
        return launch_pull<TT, 64>(p, smem, stream);
    switch (T) {
        SPK_PULL_CASE(8)
        SPK_PULL_CASE(4)
        SPK_PULL_CASE(2)
        default:
        SPK_PULL_CASE(1)
    }
#undef SPK_PULL_CASE
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
    // default: pull (inverse table + register accumulation, direct stores); SPK_BWD=push selects the
    // shared-memory scatter kernel, which also serves sizes the pull kernel does not take
    const char* mode = getenv("SPK_BWD");
    if (!(mode && mode[0] == 'p' && mode[1] == 'u' && mode[2] == 's')) {
        const int rc = gather_bwd_pull(g_cube, g_cabins, idx, cab_arg, B, C, N, R, k, cab, grad_x, (cudaStream_t)stream);
        if (rc != 1) return rc;
    }
    return gather_bwd_push(g_cube, g_cabins, idx, cab_arg, B, C, N, R, k, cab, grad_x, (cudaStream_t)stream);
}

extern "C" int sp_cabins_fwd_f32(const float* windows, long long rows, int k, int cab, float* cabins,
                                 uint16_t* cab_arg, void* stream) {
    using namespace spk;
    if (rows < 0 || cab < 1 || k < cab) return fail(SPK_E_BADARG, "sp_cabins_fwd_f32: need rows>=0, 1<=cab<=k (k=%d cab=%d)", k, cab);
    if (k / cab > 65535) return fail(SPK_E_UNSUPPORTED, "sp_cabins_fwd_f32: window length %d > 65535", k / cab);
    if (rows == 0) return SPK_OK;
    if (!windows || !cabins || !cab_arg) return fail(SPK_E_BADARG, "sp_cabins_fwd_f32: null pointer");
    const long long total = rows * cab;
    const long long work = (k / cab >= 32) ? total * 32 : total;                // a warp per long window
    const int g = (int)std::min<long long>((work + 255) / 256, (long long)sm_count() * 16);
    SPK_CUDA(launch_k(sp_cabins_generic_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, windows, k, cab, rows, cabins, cab_arg));
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
    SPK_CUDA(launch_k(sp_cabins_bwd_kernel, dim3(g), dim3(256), 0, (cudaStream_t)stream, g_cabins, cab_arg, k, cab, total, g_windows));
    return SPK_OK;
}
