// softpool_gather.cu -- SoftPool gather forward (+ window max) and its scatter-free backward.
//
// Forward replaces softpool.py:142-145 + train2cabins (softpool.py:71-85); backward is what
// autograd derives for them (reference has no hand-written backward).  Both are pure data
// movement and HBM-bound; design ("gather"):
//   * x rows (N floats, contiguous) are streamed into shared memory with 1-D TMA bulk copies
//     (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP), 2 stages per team, so the random
//     index access hits shared memory, never HBM/L2; every x byte is read exactly once;
//   * the index list of the sample (R*k entries) sits in shared memory as u16;
//   * outputs leave as 128-bit coalesced stores (forward) / TMA bulk stores of whole
//     grad_x rows (backward); grad_x is accumulated in shared memory in ascending region order,
//     so it is deterministic and needs no atomics and no pre-zeroed output.
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
    int C, N, R, k, cab;
    int rows_per_cta;     // TC
    int team_threads;     // 32 (warp per row) or blockDim.x (CTA per row)
    int stages;           // ring depth per team
    int bulk_ok;          // rows can move with cp.async.bulk (N%4==0, x 16B aligned)
    int vec4;             // k%4==0 && sp_cube 16B aligned
    int cab_fast;         // windows are whole groups of 4 slots, (wl/4) pow2 <= 32
};

__global__ void __launch_bounds__(256)
sp_gather_fwd_kernel(const GatherFwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int team_threads = p.team_threads;
    const int n_teams = blockDim.x / team_threads;
    const int team = tid / team_threads, tt = tid - team * team_threads;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * p.rows_per_cta;
    const int c1 = min(p.C, c0 + p.rows_per_cta);
    const int RK = p.R * p.k;
    const int N = p.N;

    // smem carve-up: [mbarriers][idx u16 R*k, padded to 16B][rows: n_teams * stages * N floats]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);                 // n_teams*stages
    const int n_bars = n_teams * p.stages;
    uint16_t* idx_s = reinterpret_cast<uint16_t*>(smem_raw + ((n_bars * 8 + 15) & ~15));
    float* rows_s = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(idx_s) + ((RK * 2 + 127) & ~127));
    float* my_rows = rows_s + (size_t)team * p.stages * N;
    uint64_t* my_bars = bars + team * p.stages;

    if (tid == 0) {
        for (int i = 0; i < n_bars; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // number of rows this team handles: c0+team, c0+team+n_teams, ...
    const int my_first = c0 + team;
    const int my_count = (my_first < c1) ? (c1 - my_first + n_teams - 1) / n_teams : 0;
    const float* xb = p.x + (size_t)b * p.C * N;

    // prologue: put the first `stages` rows in flight before touching the index list
    if (p.bulk_ok && tt == 0) {
        for (int s = 0; s < p.stages && s < my_count; ++s) {
            mbar_expect_tx(&my_bars[s], (uint32_t)N * 4u);
            bulk_g2s(my_rows + (size_t)s * N, xb + (size_t)(my_first + s * n_teams) * N, (uint32_t)N * 4u, &my_bars[s]);
        }
    }
    {
        const int32_t* ib = p.idx + (size_t)b * RK;
        for (int i = tid; i < RK; i += blockDim.x) idx_s[i] = (uint16_t)__ldg(ib + i);
    }
    __syncthreads();

    const int wl = p.cab > 0 ? p.k / p.cab : 0;          // window length
    const int G = p.cab_fast ? (wl >> 2) : 1;            // groups (lanes) per window
    const int wins_per_row = p.R * p.cab;

    for (int it = 0; it < my_count; ++it) {
        const int c = my_first + it * n_teams;
        const int s = it % p.stages;
        float* row = my_rows + (size_t)s * N;
        if (p.bulk_ok) {
            mbar_wait(&my_bars[s], (uint32_t)((it / p.stages) & 1));
        } else {
            const float* src = xb + (size_t)c * N;
            for (int i = tt; i < N; i += team_threads) row[i] = __ldg(src + i);
            team_sync(team, team_threads);
        }
        float* orow = p.sp_cube + ((size_t)b * p.C + c) * RK;
        if (p.vec4) {
            const int groups = RK >> 2;
            // loop bound rounded up so that every lane of a window's G-group takes part in shuffles
            const int groups_up = (groups + team_threads - 1) / team_threads * team_threads;
            for (int g = tt; g < groups_up; g += team_threads) {
                const bool live = g < groups;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    const uint2 iv = *reinterpret_cast<const uint2*>(idx_s + 4 * g);
                    v.x = row[iv.x & 0xFFFFu]; v.y = row[iv.x >> 16];
                    v.z = row[iv.y & 0xFFFFu]; v.w = row[iv.y >> 16];
                    st_cs_f4(reinterpret_cast<float4*>(orow) + g, v);
                }
                if (p.cab_fast) {
                    // first-max over this group's 4 slots, then over the G lanes of the window
                    uint32_t key = order_key(v.x), off = 0;
                    { uint32_t k2 = order_key(v.y); if (k2 > key) { key = k2; off = 1; } }
                    { uint32_t k2 = order_key(v.z); if (k2 > key) { key = k2; off = 2; } }
                    { uint32_t k2 = order_key(v.w); if (k2 > key) { key = k2; off = 3; } }
                    off += (uint32_t)(g & (G - 1)) << 2;             // offset inside the window
                    for (int d = 1; d < G; d <<= 1) {
                        const uint32_t k2 = __shfl_xor_sync(0xFFFFFFFFu, key, d);
                        const uint32_t o2 = __shfl_xor_sync(0xFFFFFFFFu, off, d);
                        amax_merge(key, off, k2, o2);
                    }
                    if (live && (g & (G - 1)) == 0) {
                        const int win = g / G;                        // (r, w) flattened
                        const int slot = win * wl + (int)off;         // slot in the row's R*k
                        const size_t o = ((size_t)b * p.C + c) * wins_per_row + win;
                        p.cabins[o] = row[idx_s[slot]];
                        p.cab_arg[o] = (uint16_t)off;
                    }
                }
            }
        } else {
            for (int sidx = tt; sidx < RK; sidx += team_threads) orow[sidx] = row[idx_s[sidx]];
        }
        // release the stage and refill it with the row `stages` ahead
        team_sync(team, team_threads);
        if (p.bulk_ok && tt == 0 && it + p.stages < my_count) {
            mbar_expect_tx(&my_bars[s], (uint32_t)N * 4u);
            bulk_g2s(row, xb + (size_t)(c + p.stages * n_teams) * N, (uint32_t)N * 4u, &my_bars[s]);
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

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct GatherBwdParams {
    const float* g_cube; const float* g_cabins; const int32_t* idx; const uint16_t* cab_arg;
    float* grad_x;
    int C, N, R, k, cab;
    int T;            // rows accumulated together in shared memory
    int groups;       // row groups per CTA
    int bulk_ok;      // N%4==0 and grad_x 16B aligned -> rows leave with cp.async.bulk
};

__global__ void __launch_bounds__(256)
sp_gather_bwd_kernel(const GatherBwdParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int b = blockIdx.y;
    const int N = p.N, R = p.R, k = p.k, T = p.T;
    const int RK = R * k;
    uint16_t* idx_s = reinterpret_cast<uint16_t*>(smem_raw);
    float* acc = reinterpret_cast<float*>(smem_raw + ((RK * 2 + 127) & ~127));     // T*N floats

    {
        const int32_t* ib = p.idx + (size_t)b * RK;
        for (int i = tid; i < RK; i += nthr) idx_s[i] = (uint16_t)__ldg(ib + i);
    }
    const int wl = (p.g_cabins != nullptr) ? k / p.cab : 1;
    const int per_region = T * k;                 // (row t, slot j) pairs per region step

    for (int gi = 0; gi < p.groups; ++gi) {
        const int c0 = (blockIdx.x * p.groups + gi) * T;
        if (c0 >= p.C) break;
        const int rows = min(T, p.C - c0);
        // the previous group's bulk store must have finished READING acc before we zero it
        if (p.bulk_ok && tid == 0) bulk_wait_read<0>();
        __syncthreads();
        {
            float4* a4 = reinterpret_cast<float4*>(acc);
            const int n4 = (rows * N) >> 2;
            for (int i = tid; i < n4; i += nthr) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = (n4 << 2) + tid; i < rows * N; i += nthr) acc[i] = 0.f;
        }
        __syncthreads();
        const size_t row0 = (size_t)b * p.C + c0;
        for (int r = 0; r < R; ++r) {
            for (int e = tid; e < per_region; e += nthr) {
                const int t = e / k, j = e - t * k;
                if (t < rows) {
                    float g = __ldg(p.g_cube + (row0 + t) * RK + (size_t)r * k + j);
                    if (p.g_cabins != nullptr) {
                        const int w = j / wl;
                        if (w < p.cab) {
                            const size_t o = ((row0 + t) * R + r) * p.cab + w;
                            if ((int)__ldg(p.cab_arg + o) == j - w * wl) g += __ldg(p.g_cabins + o);
                        }
                    }
                    acc[t * N + idx_s[r * k + j]] += g;      // indices are unique inside a region
                }
            }
            __syncthreads();
        }
        float* dst = p.grad_x + row0 * N;
        if (p.bulk_ok) {
            fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the TMA engine
            __syncthreads();
            if (tid == 0) {
                bulk_s2g(dst, acc, (uint32_t)(rows * N) * 4u);
                bulk_commit();
            }
        } else {
            for (int i = tid; i < rows * N; i += nthr) dst[i] = acc[i];
        }
    }
    if (p.bulk_ok && tid == 0) bulk_wait<0>();
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

}  // namespace spk

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
    if (B > 65535) return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: B=%d > 65535", B);
    const long long RK = (long long)R * k;

    GatherFwdParams p;
    p.x = x; p.idx = idx; p.sp_cube = sp_cube; p.cabins = cabins; p.cab_arg = cab_arg;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = want_cab ? cab : 0;
    p.bulk_ok = ((N & 3) == 0) && (((uintptr_t)x & 15) == 0);
    p.vec4 = ((k & 3) == 0) && (((uintptr_t)sp_cube & 15) == 0);
    const int wl = want_cab ? k / cab : 0;
    p.cab_fast = want_cab && p.vec4 && (k % cab == 0) && (wl % 4 == 0) && ((wl >> 2) <= 32) &&
                 (((wl >> 2) & ((wl >> 2) - 1)) == 0);
    int threads = 256;
    const size_t budget = (size_t)max_optin_smem();
    const size_t idx_bytes = ((size_t)RK * 2 + 127) & ~(size_t)127;
    // warp-per-row teams while two stages per warp fit comfortably; else the CTA is one team
    const size_t row_bytes = (size_t)N * 4;
    p.stages = 2;
    if (4 * 2 * row_bytes + idx_bytes + 256 <= 72 * 1024) {
        p.team_threads = 32;       // 4 warps, each streaming its own rows: 3 CTAs per SM at N=2048
        threads = 128;
    } else if (2 * row_bytes + idx_bytes + 256 <= budget) {
        p.team_threads = threads;
    } else {
        return fail(SPK_E_UNSUPPORTED, "sp_gather_fwd_f32: N=%d, R*k=%lld do not fit shared memory", N, RK);
    }
    const int n_teams = threads / p.team_threads;
    const size_t smem = ((size_t)(n_teams * p.stages * 8 + 15) & ~(size_t)15) + idx_bytes +
                        (size_t)n_teams * p.stages * row_bytes;
    // rows per CTA: enough CTAs for >= 2 waves when the problem allows, >= 2 rows per team
    int TC = n_teams * 4;
    while (TC > n_teams && (long long)B * ((C + TC - 1) / TC) < 2LL * sm_count()) TC -= n_teams;
    p.rows_per_cta = TC;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((C + TC - 1) / TC, B);
    sp_gather_fwd_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(p);
    SPK_LAUNCH_CHECK("sp_gather_fwd_kernel");
    if (want_cab && !p.cab_fast) {
        const long long n_rows = (long long)B * C * R;
        const long long total = n_rows * cab;
        const int g = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
        sp_cabins_generic_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(sp_cube, k, cab, n_rows, cabins, cab_arg);
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
    if (B > 65535) return fail(SPK_E_UNSUPPORTED, "sp_gather_bwd_f32: B=%d > 65535", B);
    const long long RK = (long long)R * k;
    GatherBwdParams p;
    p.g_cube = g_cube; p.g_cabins = g_cabins; p.idx = idx; p.cab_arg = cab_arg; p.grad_x = grad_x;
    p.C = C; p.N = N; p.R = R; p.k = k; p.cab = cab;
    p.bulk_ok = ((N & 3) == 0) && (((uintptr_t)grad_x & 15) == 0);
    const size_t budget = (size_t)max_optin_smem();
    const size_t idx_bytes = ((size_t)RK * 2 + 127) & ~(size_t)127;
    const size_t row_bytes = (size_t)N * 4;
    if (idx_bytes + row_bytes > budget)
        return fail(SPK_E_UNSUPPORTED, "sp_gather_bwd_f32: N=%d, R*k=%lld do not fit shared memory", N, RK);
    // T rows together: ~32 KB of accumulator keeps 4+ CTAs per SM resident
    int T = (int)std::max<size_t>(1, std::min<size_t>(8, (32 * 1024) / row_bytes));
    while (T > 1 && idx_bytes + T * row_bytes > budget) --T;
    T = min(T, C);
    p.T = T;
    const int row_groups = (C + T - 1) / T;
    int groups = 1;
    while (groups < 4 && (long long)B * ((row_groups + groups) / (groups + 1)) >= 4LL * sm_count()) ++groups;
    p.groups = groups;
    const size_t smem = idx_bytes + (size_t)T * row_bytes;
    if (smem > 48 * 1024)
        SPK_CUDA(cudaFuncSetAttribute(sp_gather_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((row_groups + groups - 1) / groups, B);
    sp_gather_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
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
