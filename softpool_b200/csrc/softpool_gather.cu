// softpool_gather.cu -- SoftPool gather forward (+ window max) and its scatter-free backward.
//
// Forward replaces softpool.py:142-145 + train2cabins (softpool.py:71-85); backward is what
// autograd derives for them (the reference has no hand-written backward).  Both are pure data
// movement and HBM-bound; design ("gather"):
//   * x rows (N floats, contiguous) are streamed into shared memory with 1-D TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP), 2 stages per team, so the random
//     index access hits shared memory, never HBM/L2; every x byte is read exactly once;
//   * persistent grid (CTAs = resident slots), every team owns a contiguous, balanced range of
//     the B*C rows: one wave, no tail;
//   * outputs leave as 128-bit coalesced stores (forward) / TMA bulk stores of whole grad_x
//     rows (backward); grad_x is accumulated in shared memory in ascending region order, so it
//     is deterministic and needs neither atomics nor a pre-zeroed output.
#include "spk_common.cuh"

namespace spk {

__device__ __forceinline__ void team_sync(int team_id, int team_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(team_id + 1), "r"(team_threads) : "memory");
}

// (key, off) max with torch.max tie rule: greater key wins, equal keys -> lower off.
__device__ __forceinline__ void amax_merge(uint32_t& key, uint32_t& off, uint32_t k2, uint32_t o2) {
    if (k2 > key || (k2 == key && o2 < off)) { key = k2; off = o2; }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct GatherFwdParams {
    const float* x; const int32_t* idx;
    float* sp_cube; float* cabins; uint16_t* cab_arg;
    long long rows;       // B*C
    int C, N, R, k, cab;
    int team_threads;     // 32 (warp per row) or blockDim.x (CTA per row)
    int stages;           // ring depth per team
    int bulk_ok;          // rows can move with cp.async.bulk (N%4==0, x 16B aligned)
    int vec4;             // k%4==0 && sp_cube, idx 16B aligned
    int cab_fast;         // windows are whole groups of 4 slots, (wl/4) pow2 <= 32
};

__global__ void __launch_bounds__(256)
sp_gather_fwd_kernel(const GatherFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int team_threads = p.team_threads;
    const int n_teams = blockDim.x / team_threads;
    const int team = tid / team_threads, tt = tid - team * team_threads;
    const int RK = p.R * p.k;
    const int N = p.N;

    // smem carve-up: [mbarriers][rows: n_teams * stages * N floats]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    const int n_bars = n_teams * p.stages;
    float* rows_s = reinterpret_cast<float*>(smem_raw + ((n_bars * 8 + 127) & ~127));
    float* my_rows = rows_s + (size_t)team * p.stages * N;
    uint64_t* my_bars = bars + team * p.stages;

    if (tid == 0) {
        for (int i = 0; i < n_bars; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // this team's contiguous range of global rows g = b*C + c
    const long long total_teams = (long long)gridDim.x * n_teams;
    const long long gteam = (long long)blockIdx.x * n_teams + team;
    const long long g_lo = p.rows * gteam / total_teams;
    const long long g_hi = p.rows * (gteam + 1) / total_teams;
    const int my_count = (int)(g_hi - g_lo);

    if (p.bulk_ok && tt == 0) {
        for (int s = 0; s < p.stages && s < my_count; ++s) {
            mbar_expect_tx(&my_bars[s], (uint32_t)N * 4u);
            bulk_g2s(my_rows + (size_t)s * N, p.x + (size_t)(g_lo + s) * N, (uint32_t)N * 4u, &my_bars[s]);
        }
    }

    const int wl = p.cab > 0 ? p.k / p.cab : 0;          // window length
    const int G = p.cab_fast ? (wl >> 2) : 1;            // groups (lanes) per window
    const int wins_per_row = p.R * p.cab;

    for (int it = 0; it < my_count; ++it) {
        const long long g = g_lo + it;
        const int b = (int)(g / p.C);
        const int32_t* ib = p.idx + (size_t)b * RK;       // L1/L2-resident: shared by the C rows of b
        const int s = it % p.stages;
        float* row = my_rows + (size_t)s * N;
        if (p.bulk_ok) {
            mbar_wait(&my_bars[s], (uint32_t)((it / p.stages) & 1));
        } else {
            const float* src = p.x + (size_t)g * N;
            for (int i = tt; i < N; i += team_threads) row[i] = __ldg(src + i);
            team_sync(team, team_threads);
        }
        float* orow = p.sp_cube + (size_t)g * RK;
        if (p.vec4) {
            const int groups = RK >> 2;
            // loop bound rounded up so that every lane of a window's G-group takes part in shuffles
            const int groups_up = (groups + team_threads - 1) / team_threads * team_threads;
            for (int q = tt; q < groups_up; q += team_threads) {
                const bool live = q < groups;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                int4 iv = make_int4(0, 0, 0, 0);
                if (live) {
                    iv = __ldg(reinterpret_cast<const int4*>(ib) + q);
                    v.x = row[iv.x]; v.y = row[iv.y]; v.z = row[iv.z]; v.w = row[iv.w];
                    st_cs_f4(reinterpret_cast<float4*>(orow) + q, v);
                }
                if (p.cab_fast) {
                    // first-max over this group's 4 slots, then over the G lanes of the window
                    uint32_t key = order_key(v.x), off = 0;
                    { uint32_t k2 = order_key(v.y); if (k2 > key) { key = k2; off = 1; } }
                    { uint32_t k2 = order_key(v.z); if (k2 > key) { key = k2; off = 2; } }
                    { uint32_t k2 = order_key(v.w); if (k2 > key) { key = k2; off = 3; } }
                    float best = off == 0 ? v.x : off == 1 ? v.y : off == 2 ? v.z : v.w;
                    off += (uint32_t)(q & (G - 1)) << 2;             // offset inside the window
                    for (int d = 1; d < G; d <<= 1) {
                        const uint32_t k2 = __shfl_xor_sync(0xFFFFFFFFu, key, d);
                        const uint32_t o2 = __shfl_xor_sync(0xFFFFFFFFu, off, d);
                        const float b2 = __shfl_xor_sync(0xFFFFFFFFu, best, d);
                        if (k2 > key || (k2 == key && o2 < off)) { key = k2; off = o2; best = b2; }
                    }
                    if (live && (q & (G - 1)) == 0) {
                        const size_t o = (size_t)g * wins_per_row + q / G;       // (r, w) flattened
                        p.cabins[o] = best;
                        p.cab_arg[o] = (uint16_t)off;
                    }
                }
            }
        } else {
            for (int sidx = tt; sidx < RK; sidx += team_threads) orow[sidx] = row[__ldg(ib + sidx)];
        }
        // release the stage and refill it with the row `stages` ahead
        team_sync(team, team_threads);
        if (p.bulk_ok && tt == 0 && it + p.stages < my_count) {
            mbar_expect_tx(&my_bars[s], (uint32_t)N * 4u);
            bulk_g2s(row, p.x + (size_t)(g + p.stages) * N, (uint32_t)N * 4u, &my_bars[s]);
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
// backward
// ---------------------------------------------------------------------------------------------
struct GatherBwdParams {
    const float* g_cube; const float* g_cabins; const int32_t* idx; const uint16_t* cab_arg;
    float* grad_x;
    long long rows;   // B*C
    int C, N, R, k, cab;
    int T;            // rows accumulated together in shared memory (one "group")
    int bulk_out;     // N%4==0 and grad_x 16B aligned -> rows leave with cp.async.bulk
    int bulk_in;      // (R*k)%4==0 and g_cube 16B aligned -> g rows arrive with cp.async.bulk
    int nbuf;         // 2 = double-buffered acc/g (default), 1 = single (rows too large for two)
    int g_direct;     // 1 = upstream gradient read straight from global (R*k too large to stage)
};

// Each CTA owns a contiguous range of global rows and walks it in groups of <= T rows that lie in
// ONE sample (so one index list serves the group).  Double-buffered: while group i is scattered
// into acc[i&1], the upstream gradient of group i+1 is in flight (bulk load) and the rows of
// group i-1 are still leaving (bulk store).
__global__ void __launch_bounds__(256)
sp_gather_bwd_kernel(const GatherBwdParams p) {
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
                const int t = e / k, j = e - t * k;
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

static int occupancy_slots(const void* kernel, int threads, size_t smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * spk::sm_count();
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
    const long long RK = (long long)R * k;
    if (RK > (1LL << 30)) return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: R*k too large");

    GatherFwdParams p;
    p.x = x; p.idx = idx; p.sp_cube = sp_cube; p.cabins = cabins; p.cab_arg = cab_arg;
    p.rows = (long long)B * C;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = want_cab ? cab : 0;
    p.bulk_ok = ((N & 3) == 0) && (((uintptr_t)x & 15) == 0);
    p.vec4 = ((k & 3) == 0) && (((uintptr_t)sp_cube & 15) == 0) && (((uintptr_t)idx & 15) == 0);
    const int wl = want_cab ? k / cab : 0;
    p.cab_fast = want_cab && p.vec4 && (k % cab == 0) && (wl % 4 == 0) && ((wl >> 2) <= 32) &&
                 (((wl >> 2) & ((wl >> 2) - 1)) == 0);
    int threads;
    const size_t budget = (size_t)max_optin_smem();
    const size_t row_bytes = (size_t)N * 4;
    p.stages = 2;
    if (4 * 2 * row_bytes + 256 <= 72 * 1024) {
        p.team_threads = 32;       // 4 warps, each streaming its own rows: 3 CTAs per SM at N=2048
        threads = 128;
    } else if (2 * row_bytes + 256 <= budget) {
        p.team_threads = threads = 256;
    } else {
        return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: a row of N=%d floats does not fit shared memory twice", N);
    }
    const int n_teams = threads / p.team_threads;
    const size_t smem = (((size_t)n_teams * p.stages * 8 + 127) & ~(size_t)127) + (size_t)n_teams * p.stages * row_bytes;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // persistent: one wave of CTAs, every team a balanced contiguous range of rows
    long long grid = occupancy_slots((const void*)sp_gather_fwd_kernel, threads, smem);
    grid = std::min<long long>(grid, (p.rows + n_teams - 1) / n_teams);
    sp_gather_fwd_kernel<<<(int)grid, threads, smem, (cudaStream_t)stream>>>(p);
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
    const long long RK = (long long)R * k;
    GatherBwdParams p;
    p.g_cube = g_cube; p.g_cabins = g_cabins; p.idx = idx; p.cab_arg = cab_arg; p.grad_x = grad_x;
    p.rows = (long long)B * C;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = g_cabins ? cab : 1;
    p.bulk_out = ((N & 3) == 0) && (((uintptr_t)grad_x & 15) == 0);
    p.bulk_in = ((RK & 3) == 0) && (((uintptr_t)g_cube & 15) == 0);
    const size_t budget = (size_t)max_optin_smem();
    const size_t fixed = 128 + (((size_t)RK * 2 + 127) & ~(size_t)127);
    size_t per_T = 2 * ((size_t)N + (size_t)RK) * 4;                // double-buffered acc + g per row
    p.nbuf = 2; p.g_direct = 0;
    if (fixed + per_T > budget) { per_T /= 2; p.nbuf = 1; }
    if (fixed + per_T > budget) { per_T = (size_t)N * 4; p.g_direct = 1; }
    if (fixed + per_T > budget)
        return fail(SPK_E_UNSUPPORTED, "sp_gather_bwd_f32: N=%d, R*k=%lld do not fit shared memory", N, RK);
    // T rows per group: aim at ~72 KB per CTA (3 CTAs per SM)
    int T = (int)std::max<size_t>(1, std::min<size_t>(8, (72 * 1024 - std::min<size_t>(fixed, 72 * 1024)) / per_T));
    T = std::min(T, C);
    p.T = T;
    const size_t smem = fixed + (size_t)T * per_T;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = occupancy_slots((const void*)sp_gather_bwd_kernel, 256, smem);
    grid = std::min<long long>(grid, (p.rows + T - 1) / T);
    sp_gather_bwd_kernel<<<(int)grid, 256, smem, (cudaStream_t)stream>>>(p);
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
