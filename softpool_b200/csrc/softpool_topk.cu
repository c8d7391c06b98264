// softpool_topk.cu -- per-region descending top-k of the activation rows (sm_100a).
//
// Replaces the region loop's `torch.sort(..., descending=True)` + `[:, :k]` (reference
// softpool.py:139-142), the float index cube (softpool.py:136-137,146-147) and the Sorter's
// argmax (softpool.py:95).  One CTA per (b, r) row (256 threads; 512 above N = 8192).
//
// Algorithm (select, compact, sort the survivors):
//   1. keys -> order_key (u32 total order: NaN first, -0 == +0), parked in shared memory;
//   2. k <= 32: a threshold T0 from warp maxima with >= k keys at or above it, everything >= T0 survives;
//      otherwise radix select, 4 bits per round (<= 8 rounds, one barrier each, early finish when the chosen
//      bucket holds <= 32 keys): the largest V with #(key >= V) >= k is the k-th largest key;
//   3. everything > V is selected, plus the first k - #(key > V) of the keys == V in ascending
//      index order (block prefix sums) -> the selected SET equals the stable-sort prefix;
//   4. the survivors, packed as unique u64 (key << 32 | ~n), are bitonic-sorted descending (stages inside a 64- or
//      128-word block run in registers / shuffles): ties come out in ascending n, i.e. the stable descending
//      order of the reference.
// Work is O(N) + O(k log^2 k) per row instead of the full O(N log^2 N) sort the reference does.
#include "spk_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace spk {

__device__ __forceinline__ int next_pow2_dev(int v) { return v <= 1 ? 1 : 1 << (32 - __clz(v - 1)); }

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

// Strides < 32 * EPL only move data inside an aligned block of 32 * EPL elements.  A warp takes such a block into registers
// (lane l holds elements l, l + 32, ...), runs every consecutive stage with such a stride there -- strides >= 32 between its
// own registers, smaller ones by shuffle -- and writes it back once: the stages that sort a block and the last stages of
// every later merge cost one shared-memory round trip each instead of one per stage (55 -> 10 block-wide steps at 1024
// survivors with 128-element blocks).
template <int EPL>
__device__ __forceinline__ void block_stages(uint64_t* buf, int base, int lane, int size_lo, int size_hi) {
    uint64_t e[EPL];
#pragma unroll
    for (int u = 0; u < EPL; ++u) e[u] = buf[base + lane + 32 * u];
    for (int size = size_lo; size <= size_hi; size <<= 1) {
        bool desc[EPL];
#pragma unroll
        for (int u = 0; u < EPL; ++u) desc[u] = ((base + lane + 32 * u) & size) == 0;
#pragma unroll
        for (int rs = EPL / 2; rs > 0; rs >>= 1) {         // stride 32 * rs: between the lane's own registers
            if (size >= 64 * rs) {
#pragma unroll
                for (int u = 0; u < EPL; ++u)
                    if ((u & rs) == 0) {                   // (desc[u] == desc[u + rs]: size > stride)
                        const bool sw = (e[u] < e[u + rs]) == desc[u];
                        const uint64_t a = sw ? e[u + rs] : e[u], c = sw ? e[u] : e[u + rs];
                        e[u] = a; e[u + rs] = c;
                    }
            }
        }
        for (int stride = min(size >> 1, 16); stride > 0; stride >>= 1) {
            const bool lower = (lane & stride) == 0;       // this lane keeps the first of the pair
#pragma unroll
            for (int u = 0; u < EPL; ++u) {
                const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, e[u], stride);
                e[u] = (lower == desc[u]) ? (o > e[u] ? o : e[u]) : (o < e[u] ? o : e[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < EPL; ++u) buf[base + lane + 32 * u] = e[u];
}

// exclusive prefix sum of one int per thread over the NT-thread CTA; returns the grand total
template <int NT>
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot /*[NT/32]*/, int tid, int& total) {
    const int lane = tid & 31, warp = tid >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        const int t = warp_tot[w];
        if (w < warp) base += t;
        tot += t;
    }
    total = tot;
    __syncthreads();                     // warp_tot may be reused by the caller
    return base + inc - v;
}

// AMV: the Sorter arg-max reads four consecutive points per thread and row with 128-bit loads and the row's keys arrive
// eight float4 at a time (long rows, whose slices keep every thread busy: N / R >= 4 NT); otherwise one point per thread
// and four float4 at a time, which keeps the kernel at 40 registers for the short rows.
template <int NT, bool AMV>
__global__ void __launch_bounds__(NT)
sp_topk_kernel(const float* __restrict__ keys, int R, int N, int k, int E, int K2, int KS, int vec_ok,
               int32_t* __restrict__ idx, float* __restrict__ sp_idx,
               int64_t* __restrict__ id_activa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw);                 // K2 survivors
    uint32_t* sk = reinterpret_cast<uint32_t*>(sel + KS);                  // E*256 keys, [e][t]
    int* hist = reinterpret_cast<int*>(sk + (size_t)E * NT);     // 3 x NT/32 warps x 16 bins
    __shared__ int warp_tot[NT / 32];
    __shared__ uint32_t cand_list[32];
    __shared__ int cand_cnt;
    __shared__ uint32_t v_final;
    const int tid = threadIdx.x;
    const int row = blockIdx.x;
    const int b = row / R, r = row - b * R;
    const float* krow = keys + (size_t)row * N;
    pdl_trigger();
    pdl_wait();
#ifdef SPK_TIMING
    long long tq[8]; tq[0] = clock64();
#define TQ(i) tq[i] = clock64()
#else
#define TQ(i)
#endif

    // ---- arg-max operands first: this thread's first points of the slice, up to 8 of the R rows, straight into
    // registers -- issued BEFORE the row's own loads, so that both batches share one trip to L2 / DRAM instead
    // of queueing behind each other (the arg-max itself is formed after the keys are parked).  Aligned rows: four
    // consecutive points per thread and row (one 128-bit load each); otherwise one point per thread.
    constexpr int AM_PRE = 8;
    typename std::conditional<AMV, float4, float>::type am[AM_PRE];
    const int am_slice = (N + R - 1) / R;
    const int am_s0 = r * am_slice, am_s1 = min(N, am_s0 + am_slice);
    const float* kb = keys + (size_t)b * R * N;
    const bool am_mine = id_activa != nullptr && (AMV ? am_s0 + 4 * tid < am_s1 : am_s0 + tid < am_s1);
    if constexpr (AMV) {
#pragma unroll
        for (int rr = 0; rr < AM_PRE; ++rr)
            am[rr] = (am_mine && rr < R) ? __ldg(reinterpret_cast<const float4*>(kb + (size_t)rr * N + am_s0 + 4 * tid)) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
        for (int rr = 0; rr < AM_PRE; ++rr) am[rr] = (am_mine && rr < R) ? __ldg(kb + (size_t)rr * N + am_s0 + tid) : 0.f;
    }

    // ---- 1. load: thread t owns the E consecutive points n = t*E .. t*E+E-1 ------------------------
    // (all loads are issued before the first use: order_key is branch-free, nothing serialises them)
    const int n0 = tid * E;
    if (vec_ok && (E & 3) == 0 && (N & 3) == 0) {
        constexpr int LW = AMV ? 8 : 4;                             // float4 loads in flight per thread
        for (int e0 = 0; e0 < E; e0 += 4 * LW) {
            float4 v[LW];
#pragma unroll
            for (int u = 0; u < LW; ++u) {
                const int n = n0 + e0 + 4 * u;
                v[u] = (e0 + 4 * u < E && n < N) ? __ldg(reinterpret_cast<const float4*>(krow + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < LW; ++u) {
                const int e = e0 + 4 * u;
                if (e < E) {
                    const bool in = n0 + e < N;                             // N%4==0: all four or none
                    sk[(e + 0) * NT + tid] = in ? order_key(v[u].x) : 0u;   // pad 0 < every real key
                    sk[(e + 1) * NT + tid] = in ? order_key(v[u].y) : 0u;
                    sk[(e + 2) * NT + tid] = in ? order_key(v[u].z) : 0u;
                    sk[(e + 3) * NT + tid] = in ? order_key(v[u].w) : 0u;
                }
            }
        }
    } else {
        for (int e0 = 0; e0 < E; e0 += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(krow + min(n0 + e0 + u, N - 1));
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (e0 + u < E) sk[(e0 + u) * NT + tid] = (n0 + e0 + u < N) ? order_key(v[u]) : 0u;
        }
    }
    for (int i = tid; i < KS; i += NT) sel[i] = 0ull;
    for (int i = tid; i < 3 * (NT / 32) * 16; i += NT) hist[i] = 0;
    if (tid == 0) cand_cnt = 0;

    // ---- argmax over the R rows for this CTA's slice of n (softpool.py:95) --------------------------
    if constexpr (AMV) {
        if (id_activa != nullptr) {
            for (int n = am_s0 + 4 * tid, first = 1; n < am_s1; n += 4 * NT, first = 0) {
                if (!first) {
#pragma unroll
                    for (int rr = 0; rr < AM_PRE; ++rr)
                        if (rr < R) am[rr] = __ldg(reinterpret_cast<const float4*>(kb + (size_t)rr * N + n));
                }
                uint32_t bk[4] = {order_key(am[0].x), order_key(am[0].y), order_key(am[0].z), order_key(am[0].w)};
                int bi[4] = {0, 0, 0, 0};
#pragma unroll
                for (int rr = 1; rr < AM_PRE; ++rr) {
                    const uint32_t kk[4] = {order_key(am[rr].x), order_key(am[rr].y), order_key(am[rr].z), order_key(am[rr].w)};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool better = rr < R && kk[u] > bk[u];     // strict: first maximum / first NaN
                        bk[u] = better ? kk[u] : bk[u]; bi[u] = better ? rr : bi[u];
                    }
                }
                longlong2* o = reinterpret_cast<longlong2*>(id_activa + (size_t)b * N + n);
                o[0] = make_longlong2((long long)bi[0], (long long)bi[1]);
                o[1] = make_longlong2((long long)bi[2], (long long)bi[3]);
            }
        }
    } else if (id_activa != nullptr) {
        const int s0 = am_s0, s1 = am_s1;
        if (am_mine) {                                           // first point: the first AM_PRE rows are in registers
            uint32_t bestk = order_key(am[0]);
            int besti = 0;
#pragma unroll
            for (int rr = 1; rr < AM_PRE; ++rr) {
                const uint32_t kk = order_key(am[rr]);
                const bool better = rr < R && kk > bestk;        // strict: first maximum / first NaN
                bestk = better ? kk : bestk; besti = better ? rr : besti;
            }
            for (int rr = AM_PRE; rr < R; ++rr) {
                const uint32_t kk = order_key(__ldg(kb + (size_t)rr * N + s0 + tid));
                const bool better = kk > bestk;
                bestk = better ? kk : bestk; besti = better ? rr : besti;
            }
            id_activa[(size_t)b * N + s0 + tid] = (int64_t)besti;
        }
        for (int n = s0 + tid + NT; n < s1; n += NT) {
            uint32_t bestk = order_key(__ldg(kb + n));
            int besti = 0;
#pragma unroll 8
            for (int rr = 1; rr < R; ++rr) {
                const uint32_t kk = order_key(__ldg(kb + (size_t)rr * N + n));
                const bool better = kk > bestk;                  // strict: first maximum / first NaN
                bestk = better ? kk : bestk; besti = better ? rr : besti;
            }
            id_activa[(size_t)b * N + n] = (int64_t)besti;
        }
    }
    __syncthreads();
    TQ(1);

    // ---- 2a. small k (k <= 32, the BASELINE configuration): no radix rounds at all.  The J = ceil(k / 8) largest of a warp's
    // 32 thread-maxima are J distinct keys, so T0 = the smallest of the 8 warps' J-th largest has at least 8 J >= k keys at or
    // above it: the k-th largest key is >= T0.  Everything >= T0 (about 2 k keys; all ties of the k-th key included) is
    // compacted in index order and sorted; the first k words are the stable descending prefix.  Falls through to the radix
    // select when more than KS keys survive (heavy ties).
    // (For larger k the same shortcut with a sampled pivot -- an order statistic of NT keys, aimed at 1.5 k survivors, ordered
    // by a bucket / rank sort or a bitonic network over 2 K2 slots -- was built and measured: 36 us at N = 8192, k = 1024 against
    // ~30 us for the radix select below followed by the fused network over K2 slots; dropped, profiles/README.md.)
    int K2s = K2;                                        // slots the survivor sort works on
    bool done_fast = false;
#ifndef SPK_NO_TOPK_SHORTCUT
    if (k <= 32 && KS >= 256) {
        __shared__ uint32_t warp_thr[NT / 32];
        uint32_t tmax = 0;
        for (int e = 0; e < E; ++e) tmax = max(tmax, sk[e * NT + tid]);
        const int J = (k + NT / 32 - 1) / (NT / 32);
        const int lane_ = tid & 31;
        uint32_t v = tmax, jth = 0;
        for (int j = 0; j < J; ++j) {
            jth = __reduce_max_sync(0xFFFFFFFFu, v);
            const unsigned bal = __ballot_sync(0xFFFFFFFFu, v == jth);
            if (lane_ == __ffs((int)bal) - 1) v = 0u;            // remove ONE holder of the maximum
        }
        if (lane_ == 0) warp_thr[tid >> 5] = jth;
        __syncthreads();
        uint32_t T0 = warp_thr[0];
#pragma unroll
        for (int w = 1; w < NT / 32; ++w) T0 = min(T0, warp_thr[w]);
        int mine = 0;
        for (int e = 0; e < E; ++e) mine += sk[e * NT + tid] >= T0;
        int total;
        int pos = block_exclusive_scan<NT>(mine, warp_tot, tid, total);
        if (total <= KS && T0 > 0u) {                            // (T0 == 0: padding keys could be counted -> generic path)
            for (int e = 0; e < E; ++e) {
                const uint32_t key = sk[e * NT + tid];
                if (key >= T0) sel[pos++] = ((uint64_t)key << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)(n0 + e));
            }
            K2s = max(2, next_pow2_dev(total));
            done_fast = true;
        }
    }
#endif

    // ---- 2. radix select of the k-th largest key, 4 bits per round ---------------------------------------
    // Warp-private 16-bin histograms of the keys that still match the decided prefix (shared-memory
    // atomics), one barrier per round; every warp then reduces the 8 histograms itself.
    // (8-bit digits / 256-bin histograms: measured slower, 17.4 vs 12.5 us at N=2048.)
    uint32_t V = 0;
    int fin_sh = -1;                                     // >= 0: the select stopped early with bits [31:fin_sh] decided
    int want = k;                                        // rank still to be located inside the prefix bucket
    const int lane = tid & 31, warp = tid >> 5;
    for (int round = 0; round < 8 && !done_fast; ++round) {
        const int sh = 28 - 4 * round;
        int* H = hist + (round % 3) * ((NT / 32) * 16);
        const uint32_t pre = (round == 0) ? 0u : (V >> (sh + 4));
#pragma unroll 4
        for (int e = 0; e < E; ++e) {
            const uint32_t key = sk[e * NT + tid];
            const bool cand = (round == 0) || ((key >> (sh + 4)) == pre);
            if (cand) atomicAdd(&H[warp * 16 + ((key >> sh) & 15u)], 1);
        }
        __syncthreads();
        int c = 0;
        if (lane < 16) {
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) c += H[w * 16 + lane];
        }
        // suffix sums over digits: S(d) = # candidates with digit >= d   (lanes >= 16 hold 0)
        int S = c;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const int o = __shfl_down_sync(0xFFFFFFFFu, S, d);
            if (lane + d < 16) S += o;
        }
        const unsigned ok = __ballot_sync(0xFFFFFFFFu, lane < 16 && S >= want);   // digits whose suffix reaches the rank
        const int dsel = 31 - __clz((int)ok);                                      // the largest such digit (bit 0 is always set)
        const int above = __shfl_sync(0xFFFFFFFFu, S - c, dsel);                   // candidates with a larger digit
        want -= above;
        V |= (uint32_t)dsel << sh;
        // few candidates left in the chosen bucket: one warp ranks them directly (below) instead of
        // spending the remaining rounds (a barrier and a reduction each) on a handful of keys.
        // c, S, dsel are identical in every warp, so the branch is uniform.
        if (round < 7 && __shfl_sync(0xFFFFFFFFu, c, dsel) <= 32) { fin_sh = sh; break; }
        // recycle the histogram used two rounds from now (everybody is past the barrier of the previous round)
        int* Hz = hist + ((round + 2) % 3) * ((NT / 32) * 16);
        if (tid < (NT / 32) * 16) Hz[tid] = 0;
    }

    if (fin_sh >= 0 && !done_fast) {
        // the bucket's <= 32 keys -> shared list; warp 0 finds the one with exactly want-1 keys ahead of it
        const uint32_t pre = V >> fin_sh;
#pragma unroll 4
        for (int e = 0; e < E; ++e) {
            const uint32_t key = sk[e * NT + tid];
            if ((key >> fin_sh) == pre) cand_list[atomicAdd(&cand_cnt, 1)] = key;
        }
        __syncthreads();
        if (warp == 0) {
            const int nc = cand_cnt;
            const uint32_t key = lane < nc ? cand_list[lane] : 0u;
            int ahead = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t o = __shfl_sync(0xFFFFFFFFu, key, j);
                ahead += (j < nc) && (o > key || (o == key && j < lane));
            }
            if (lane < nc && ahead == want - 1) v_final = key;
        }
        __syncthreads();
        V = v_final;
    }

    TQ(2);
    // ---- 3. compaction in index order ---------------------------------------------------------------------
    if (!done_fast) {
    int my_gt = 0, my_eq = 0;
#pragma unroll 4
    for (int e = 0; e < E; ++e) {
        const uint32_t key = sk[e * NT + tid];
        my_gt += key > V; my_eq += key == V;
    }
    // one scan for both counts: (gt << 16) | eq  (each <= E*256 <= 16384)
    int total_ge;
    const int ge_before = block_exclusive_scan<NT>((my_gt << 16) | my_eq, warp_tot, tid, total_ge);
    const int eq_before = ge_before & 0xFFFF, total_gt = total_ge >> 16;
    const int need = k - total_gt;                       // >= 1 ties to take, lowest indices first
    // the ties taken are the first `need` in index order, so the ones before this thread are min(eq_before, need)
    int pos = (ge_before >> 16) + min(eq_before, need);
    int eq_rank = eq_before;
    for (int e = 0; e < E; ++e) {
        const uint32_t key = sk[e * NT + tid];
        bool take = key > V;
        if (key == V) { take = eq_rank < need; ++eq_rank; }
        if (take) { sel[pos++] = ((uint64_t)key << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)(n0 + e)); }
    }
    }

    TQ(3);
#ifndef SPK_NO_TOPK_TRIGGER
    // The gather's CTAs may start now (this kernel's own griddepcontrol.wait is long past, so everything before it in the
    // stream is complete): their first x tiles stream in during the survivor sort + emit.  Measured on the whole step at
    // config A: no trigger 74.1 us, trigger at kernel start 75.7 us (the waiting gather CTAs take the second-wave slots of
    // this kernel's own CTAs), trigger here 72.2 us.
    pdl_launch_dependents();
#endif
    // ---- 4. sort the survivors (descending; unique words => stable order) ---------------------------------
    if (K2s <= 32) {
        // one warp, one word per lane, bitonic network on shuffles: no barriers
        __syncthreads();
        if (warp == 0) {
            uint64_t v = (lane < K2s) ? sel[lane] : 0ull;
            for (int size = 2; size <= K2s; size <<= 1)
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, v, stride);
                    const bool desc = (lane & size) == 0 || size == K2s;
                    const bool lower = (lane & stride) == 0;          // this lane keeps the first of the pair
                    v = (lower == desc) ? (o > v ? o : v) : (o < v ? o : v);
                }
            if (lane < K2s) sel[lane] = v;
        }
    } else {
        // K2s is a power of two >= 64.  Blocks of 128 words when that still gives every warp one (K2s >= 1024), else of 64:
        // measured at K2s = 256, sort phase 3.9k cycles with 64-word blocks (4 warps busy) against 7.5k with 128-word ones
        __syncthreads();
        auto fused_sort = [&](auto epl_tag) {
            constexpr int EPL = decltype(epl_tag)::value, BL = 32 * EPL;
            for (int base = warp * BL; base < K2s; base += (NT / 32) * BL) block_stages<EPL>(sel, base, lane, 2, BL);
            for (int size = 2 * BL; size <= K2s; size <<= 1) {
                for (int stride = size >> 1; stride >= BL; stride >>= 1) {
                    __syncthreads();
                    ce_stage(sel, K2s, size, stride, tid, NT);
                }
                __syncthreads();
                for (int base = warp * BL; base < K2s; base += (NT / 32) * BL) block_stages<EPL>(sel, base, lane, size, size);
            }
        };
        if (K2s >= 1024) fused_sort(std::integral_constant<int, 4>()); else fused_sort(std::integral_constant<int, 2>());
    }
    __syncthreads();
    TQ(4);

    // ---- emit ---------------------------------------------------------------------------------------------
    pdl_tail_trigger();
    int32_t* orow = idx + (size_t)row * k;
    {   // one shared-memory read per survivor, then its R + 3 float copies (softpool.py:136-137,146-147), coalesced per copy
        const int Q = sp_idx != nullptr ? R + 3 : 0;
        float* cube = sp_idx != nullptr ? sp_idx + ((size_t)b * (R + 3) * R + r) * k : nullptr;
        const size_t qstride = (size_t)R * k;
        for (int j = tid; j < k; j += NT) {
            const uint32_t n = 0xFFFFFFFFu - (uint32_t)sel[j];
            orow[j] = (int32_t)n;
            const float fv = (float)n;
            for (int q = 0; q < Q; ++q) cube[(size_t)q * qstride + j] = fv;
        }
    }
#ifdef SPK_TIMING
    TQ(5);
    if (tid == 0 && (blockIdx.x == 0 || blockIdx.x == 200))
        printf("topk cta %d: load+argmax %lld  select %lld  compact %lld  sort %lld  emit %lld  (cycles)\n", blockIdx.x,
               tq[1] - tq[0], tq[2] - tq[1], tq[3] - tq[2], tq[4] - tq[3], tq[5] - tq[4]);
#endif
}

__global__ void sp_argmax_kernel(const float* __restrict__ keys, int R, int N,
                                 int64_t* __restrict__ id_activa, long long total) {
    pdl_trigger();
    pdl_wait();
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
    const int vec_ok = ((uintptr_t)keys & 15) == 0;     // misaligned keys: scalar loads (the gather and Chamfer kernels do the same)
    const int K2 = max(2, next_pow2(k));
    // CTA width: 256 threads are the fastest up to N = 8192 (measured, us at 256 / 512 / 1024 threads: N=2048 k=32 7.5 / 12.5 / 19.3,
    // k=256 14.0 / 20.6 / 30.2; N=4096 22.2 / 26.3 / 36.4; N=8192 42.6 / 43.0 / 49.6); only the longest rows gain from 512
    // (N=16384: 88.4 / 72.5 / 86.7)
    int NT = N > 8192 ? 512 : 256;
#ifdef SPK_EXPERIMENT
    if (const char* e = getenv("SPK_TOPK_THREADS")) { const int v = atoi(e); if (v == 128 || v == 256 || v == 512 || v == 1024) NT = v; }
#endif
    int E = (N + NT - 1) / NT;
    if (E > 4) E = (E + 3) & ~3;                       // float4 path wants E % 4 == 0
    const int KS = k <= 32 ? std::max(K2, 256) : K2;       // small k: room for every key at or above the shortcut's threshold
    const size_t smem = (size_t)KS * sizeof(uint64_t) + (size_t)E * NT * sizeof(uint32_t) + 3 * (NT / 32) * 16 * sizeof(int);
    auto launch = [&](auto kern) -> int {
        if (smem > 48 * 1024)
            SPK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPK_CUDA(launch_k(kern, dim3(B * R), dim3(NT), smem, (cudaStream_t)stream, keys, R, N, k, E, K2, KS, vec_ok, idx, sp_idx, id_activa));
        return SPK_OK;
    };
    // wide loads (AMV) for rows whose arg-max slices keep every thread busy with four points; needs aligned rows
    const int am_slice = (N + R - 1) / R;
    const bool amv = vec_ok && (N & 3) == 0 && (am_slice & 3) == 0 && am_slice >= 4 * NT && R <= 8 && (!id_activa || ((uintptr_t)id_activa & 15) == 0);
    if (amv) return NT == 1024 ? launch(sp_topk_kernel<1024, true>) : NT == 512 ? launch(sp_topk_kernel<512, true>) : NT == 128 ? launch(sp_topk_kernel<128, true>) : launch(sp_topk_kernel<256, true>);
    return NT == 1024 ? launch(sp_topk_kernel<1024, false>) : NT == 512 ? launch(sp_topk_kernel<512, false>) : NT == 128 ? launch(sp_topk_kernel<128, false>) : launch(sp_topk_kernel<256, false>);
}

extern "C" int sp_argmax_i64(const float* keys, int B, int R, int N, int64_t* id_activa, void* stream) {
    using namespace spk;
    if (B < 0 || R < 1 || N < 1) return fail(SPK_E_BADARG, "sp_argmax_i64: need B>=0, R>=1, N>=1");
    if (B == 0) return SPK_OK;
    if (!keys || !id_activa) return fail(SPK_E_BADARG, "sp_argmax_i64: null pointer");
    const long long total = (long long)B * N;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
    SPK_CUDA(launch_k(sp_argmax_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, keys, R, N, id_activa, total));
    return SPK_OK;
}
