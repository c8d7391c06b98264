// chamfer_dense.cu -- Chamfer forward, DENSE form: the whole n x m pair block on the 5th-gen tensor cores (tcgen05 + TMEM),
// sm_100a.  Used for small pair blocks (chamfer.cu picks; large clouds go to the sorted search of chamfer_tc.cu).
//
// Replaces NmDistanceKernel x2 (reference distance/chamfer/chamfer.cu:12-143) and returns
// bit-identical dist/idx for finite inputs, yet evaluates the n x m pair block on the tensor pipe.
//
// Idea.  The pairwise block IS a dense contraction:  d(p,q) = |p|^2 + |q|^2 - 2 p.q.  With fp16
// hi/lo splits of the (centred, power-of-two scaled) coordinates and 3-way fp16 splits of the
// norms, ONE K=16 MMA row pair produces bias + the squared distance with absolute error
// e <= ~2^-17 (scaled units) before the accumulator's own rounding:
//     A'(p) = [-2xh,-2xh,-2xl, -2yh,-2yh,-2yl, -2zh,-2zh,-2zl,  nh,nm,nl,  1,1,1,  1   ]
//     B'(q) = [  xh,  xl,  xh,   yh,  yl,  yh,   zh,  zl,  zh,   1, 1, 1,  mh,mm,ml, bias]
// The accumulator is fp16 (one instruction forms the whole sum, so its only fp16 rounding is the last
// one) and bias (a power of two above the error bound) keeps every value a POSITIVE fp16, whose bit
// pattern orders like its value: the epilogue reads two columns per register (tcgen05.ld ...pack::16b)
// and reduces with VIMNMX3.U16x2, four new elements per instruction (tools/micro/minbench.cu: twice the
// element rate of any fp32 min).  Target rows are permuted inside groups of 32 (b_row_of) so that the two
// 16-bit lanes of a packed minimum are two contiguous chunks of 16 targets.
// The approximate block only FILTERS: per query row the minimum of every 16-target chunk is kept, then only
// chunks whose minimum is within the slack (relative 2^-8 for the fp16 rounding, absolute tau = 2e) of the
// row minimum -- or of the exact best so far -- are re-evaluated with the reference's exact float32
// expression  d = fma(dz,dz, fma(dx,dx, dy*dy))  and first-minimum tie rule.  A chunk that holds the true
// nearest neighbour always passes the filter (its approximate distance is <= d_true + e <= d_any + e <=
// approx_any + 2e), so the result equals the brute-force one.
//
// Kernel structure (persistent, 2 CTAs per SM; a job = 128 queries of one sample and direction, or -- when
// the grid would be underfilled -- a sub-range of that job's targets, merged through a 64-bit atomicMin):
//   warp 0     TMA producer: A tile (128 x 32 B) per job (double-buffered), B tiles (128 targets x 32 B) through
//              a ring, and the raw float4 coordinates of every 1024-target super-block (for the exact pass)
//   warp 1     TMEM alloc (256 columns = 2 accumulator buffers of 128) + single-thread tcgen05.mma issue
//              (M=128, N=128, K=16, kind::f16, fp16 accumulate) + tcgen05.commit -> mbarriers
//   warps 2-5  "min" warps, one per TMEM lane quarter: accumulator -> registers (buffer handed back at once)
//              -> packed chunk minima -> shared memory, double-buffered per super-block
//   warps 6-9  "exact" warps, one thread per query row: filter, exact re-evaluation from shared memory,
//              running best, final store (+ the fused mean loss); up to two super-blocks behind the min warps
// Operands are pre-formatted by chamfer_prep_kernel in the canonical no-swizzle K-major layout
// (8-row x 16-byte core matrices, LBO = 128 B, SBO = 256 B), so tiles move with 1-D bulk copies.
#include "spk_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

#ifndef SPK_SPIN
#define SPK_SPIN 0
#endif

namespace spk {
namespace dense {

constexpr int TC_TILE = 128;            // queries per CTA = targets per shared-memory B tile
#ifndef SPK_TC_N
#define SPK_TC_N 128
#endif
constexpr int TC_N = SPK_TC_N;          // targets per MMA (accumulator buffer width, TMEM columns)
constexpr int TC_NBUF = 256 / TC_N;              // accumulator buffers: hides the release -> MMA -> commit round trip
constexpr int TC_STAGES = 4;            // B-tile ring depth
constexpr int TC_THREADS = 320;         // 10 warps
constexpr int TC_SB_TILES = 8;          // tiles per super-block (1024 targets)
constexpr int TC_SB_TARGETS = TC_SB_TILES * TC_TILE;
constexpr int TC_CHUNK = 16;            // targets per filter chunk
constexpr int TC_CM_WORDS = TC_SB_TARGETS / 32;   // packed words (2 chunk minima each) per query row and super-block
constexpr int TC_MIN_WARPS = 4, TC_EXACT_WARPS = 4;
constexpr int TC_T4_BUFS = 2;           // raw-coordinate buffers: the exact warps lag the operand stream by up to two super-blocks
constexpr int TC_TILE_BYTES = TC_TILE * 32;
constexpr float TC_PAD_NORM = 30000.f;  // norm of padding targets: never the minimum

struct ChamferMeta {                    // per sample, written by the prep kernel
    float cx, cy, cz;                   // centre (bounding-box midpoint of both clouds)
    float scale;                        // power of two: |(x - c) * scale| <= 1
    float tau;                          // filter slack in scaled squared units
    float scale2;                       // scale * scale
    float bias;                         // power of two added to every approximate distance (keeps them > 0)
    float nonfinite;                    // != 0: some coordinate is NaN/inf -> every chunk is evaluated exactly
};

// ---------------------------------------------------------------------------------------------
// prep: centre/scale per sample, fp16 split operands in UMMA canonical layout
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split2(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn(v - __half2float(h));
}
__device__ __forceinline__ void split3(float v, __half& h, __half& m, __half& l) {
    h = __float2half_rn(v);
    const float r1 = v - __half2float(h);
    m = __float2half_rn(r1);
    l = __float2half_rn(r1 - __half2float(m));
}

// byte offset of row r's first 16-byte K-chunk inside an operand array
__device__ __forceinline__ size_t op_row_offset(int r) { return (size_t)(r >> 3) * 256 + (size_t)(r & 7) * 16; }

// B rows are permuted inside every aligned group of 32 targets: target u of the group sits in row
// ((u & 15) << 1) | (u >> 4), so that the EVEN accumulator columns of the group are targets 0..15 and the
// ODD columns targets 16..31 -- the two 16-bit lanes of the packed minimum then hold two contiguous chunks.
__device__ __forceinline__ int b_row_of(int r) { return (r & ~31) | ((r & 15) << 1) | ((r >> 4) & 1); }

__device__ __forceinline__ void write_rows(unsigned char* opA, unsigned char* opB, int r, bool real,
                                           float ux, float uy, float uz, float bias) {
    __align__(16) __half a[16];
    __align__(16) __half b[16];
    if (real) {
        __half xh, xl, yh, yl, zh, zl, nh, nm, nl;
        split2(ux, xh, xl); split2(uy, yh, yl); split2(uz, zh, zl);
        const float nrm = fmaf(uz, uz, fmaf(uy, uy, ux * ux));
        split3(nrm, nh, nm, nl);
        const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f), m2 = __float2half_rn(-2.f);
        a[0] = __hmul(m2, xh); a[1] = a[0]; a[2] = __hmul(m2, xl);
        a[3] = __hmul(m2, yh); a[4] = a[3]; a[5] = __hmul(m2, yl);
        a[6] = __hmul(m2, zh); a[7] = a[6]; a[8] = __hmul(m2, zl);
        a[9] = nh; a[10] = nm; a[11] = nl; a[12] = one; a[13] = one; a[14] = one; a[15] = one;
        b[0] = xh; b[1] = xl; b[2] = xh; b[3] = yh; b[4] = yl; b[5] = yh; b[6] = zh; b[7] = zl; b[8] = zh;
        b[9] = one; b[10] = one; b[11] = one; b[12] = nh; b[13] = nm; b[14] = nl; b[15] = __float2half_rn(bias);
    } else {
        const __half zero = __float2half_rn(0.f), one = __float2half_rn(1.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = zero; b[i] = zero; }
        b[9] = one; b[12] = __float2half_rn(TC_PAD_NORM);     // a padding target is "infinitely" far
    }
    const size_t o = op_row_offset(r), ob = op_row_offset(b_row_of(r));
    *reinterpret_cast<uint4*>(opA + o) = *reinterpret_cast<const uint4*>(&a[0]);
    *reinterpret_cast<uint4*>(opA + o + 128) = *reinterpret_cast<const uint4*>(&a[8]);
    *reinterpret_cast<uint4*>(opB + ob) = *reinterpret_cast<const uint4*>(&b[0]);
    *reinterpret_cast<uint4*>(opB + ob + 128) = *reinterpret_cast<const uint4*>(&b[8]);
}

struct PrepParams {
    const float* xyz1; const float* xyz2;
    int n, m, n_pad, m_pad;
    unsigned char* A1; unsigned char* B1; unsigned char* A2; unsigned char* B2;   // per sample n_pad*32 / m_pad*32 bytes
    float4* T1; float4* T2;          // raw coordinates (x,y,z,0), padded per sample to n_pad / m_pad rows
    ChamferMeta* meta;
    unsigned long long* packed1; unsigned long long* packed2;   // split jobs only: (dist bits << 32 | idx) minima, B*n / B*m
    int* counters; int n_counters;   // split jobs only: arrivals per (sample, direction, query tile)
    float* loss;                     // fused loss only: (B) accumulators, zeroed here
    int n_gt;                        // distinct xyz2 samples: sample b pairs xyz1[b] with xyz2[b % n_gt] (several predictions share one
                                     // ground truth, train.py:68-86); n_gt == B for the plain call
    int P;                           // predictions per ground truth = B / B2
};

__global__ void __launch_bounds__(256)
chamfer_prep_kernel(const PrepParams p) {
    __shared__ float red[6][8];
    __shared__ float s_meta[8];
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y, tid = threadIdx.x;
    const int b2 = b % p.n_gt;                          // the ground-truth sample; its operands are written by prediction 0 only
    const bool own_q = b < p.n_gt;
    const float* P = p.xyz1 + (size_t)b * p.n * 3;
    const float* Q = p.xyz2 + (size_t)b2 * p.m * 3;
    // this thread's own rows first (raw coordinates into registers): their trip to L2 / DRAM then overlaps the
    // bounding-box pass instead of following it
    constexpr int PRE = 2;
    float prx[PRE], pry[PRE], prz[PRE];
    {
        const int total_rows = p.n_pad + p.m_pad;
#pragma unroll
        for (int j = 0; j < PRE; ++j) {
            const int i = blockIdx.x * 256 + tid + j * (int)gridDim.x * 256;
            prx[j] = pry[j] = prz[j] = 0.f;
            if (i < total_rows) {
                const bool first = i < p.n_pad;
                const int r = first ? i : i - p.n_pad;
                if (r < (first ? p.n : p.m)) {
                    const float* sp = (first ? P : Q) + 3 * (size_t)r;
                    prx[j] = __ldg(sp); pry[j] = __ldg(sp + 1); prz[j] = __ldg(sp + 2);
                }
            }
        }
    }
    // bounding box over both clouds (every CTA of the sample recomputes it: 12*(n+m) bytes from L2).
    // 128-bit loads, three per step = four whole points, so the axis of every lane is static.
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    int bad = 0;                                        // a NaN / inf coordinate anywhere in the sample
    auto upd = [&](int a, float v) { lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); bad |= !(fabsf(v) < INFINITY); };
    // (all predictions that share the ground truth share ONE frame -- centre, scale, bias -- so that the ground truth's
    // operand rows are the same for each of them: the box runs over the ground truth and every prediction of sample b2)
    for (int c = 0; c <= p.P; ++c) {
        const float* X = c == p.P ? Q : p.xyz1 + ((size_t)c * p.n_gt + b2) * p.n * 3;
        const int cnt = c == p.P ? p.m : p.n;
        int done = 0;
        if ((((uintptr_t)X) & 15) == 0) {
            const int steps = cnt / 4;                               // 4 points = 12 floats = 3 float4
            const float4* X4 = reinterpret_cast<const float4*>(X);
            for (int st = tid; st < steps; st += 256) {
                const float4 a = __ldg(X4 + 3 * st), b4 = __ldg(X4 + 3 * st + 1), c4 = __ldg(X4 + 3 * st + 2);
                upd(0, a.x); upd(1, a.y); upd(2, a.z); upd(0, a.w);
                upd(1, b4.x); upd(2, b4.y); upd(0, b4.z); upd(1, b4.w);
                upd(2, c4.x); upd(0, c4.y); upd(1, c4.z); upd(2, c4.w);
            }
            done = steps * 4;
        }
        for (int i = done + tid; i < cnt; i += 256) { upd(0, __ldg(X + 3 * i)); upd(1, __ldg(X + 3 * i + 1)); upd(2, __ldg(X + 3 * i + 2)); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int d = 16; d > 0; d >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], d));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], d));
        }
    if ((tid & 31) == 0)
#pragma unroll
        for (int a = 0; a < 3; ++a) { red[a][tid >> 5] = lo[a]; red[3 + a][tid >> 5] = hi[a]; }
    bad = __syncthreads_or(bad);
    if (tid == 0) {
        float L[3], H[3];
        for (int a = 0; a < 3; ++a) {
            L[a] = red[a][0]; H[a] = red[3 + a][0];
            for (int w = 1; w < 8; ++w) { L[a] = fminf(L[a], red[a][w]); H[a] = fmaxf(H[a], red[3 + a][w]); }
        }
        float c[3], ext = 0.f, amax = 0.f;
        for (int a = 0; a < 3; ++a) {
            c[a] = 0.5f * L[a] + 0.5f * H[a];
            ext = fmaxf(ext, fmaxf(H[a] - c[a], c[a] - L[a]));
            amax = fmaxf(amax, fmaxf(fabsf(L[a]), fabsf(H[a])));
        }
        // scale = 2^-ceil(log2(ext)) so that |x - c| * scale <= 1; degenerate / non-finite boxes -> 1
        float scale = 1.f;
        if (ext > 0.f && ext < INFINITY) {
            int e; (void)frexpf(ext, &e);                 // ext = f * 2^e, f in [0.5, 1)
            e = max(-100, min(100, e));
            scale = ldexpf(1.f, -e);
        }
        // error budget of the approximate block, scaled units: fp16 hi/lo products + fp32 accumulate
        // (2^-17) plus the rounding of the centred coordinates themselves (|x| * 2^-23 * scale each)
        const float delta = amax * scale * 1.1920929e-7f;
        const float e_tot = 7.62939453125e-6f + 16.f * delta;
        s_meta[0] = c[0]; s_meta[1] = c[1]; s_meta[2] = c[2]; s_meta[3] = scale;
        s_meta[4] = 2.f * e_tot; s_meta[5] = scale * scale;
        // bias: a power of two above the error bound, so that every approximate distance is a POSITIVE fp16
        // (its bit pattern then orders like its value); 2^-6 unless the coordinates are badly conditioned
        float bias = 0.015625f;
        while (bias < 4.f * e_tot && bias < 1024.f) bias *= 2.f;
        s_meta[6] = bias;
        if (blockIdx.x == 0) {
            ChamferMeta mm; mm.cx = c[0]; mm.cy = c[1]; mm.cz = c[2]; mm.scale = scale; mm.tau = 2.f * e_tot;
            mm.scale2 = scale * scale; mm.bias = bias;
            mm.nonfinite = (bad || !(bias >= 4.f * e_tot)) ? 1.f : 0.f;     // hopeless conditioning counts as non-finite
            p.meta[b] = mm;
        }
    }
    __syncthreads();
    const float cx = s_meta[0], cy = s_meta[1], cz = s_meta[2], sc = s_meta[3], bias = s_meta[6];
    unsigned char* A1 = p.A1 + (size_t)b * p.n_pad * 32; unsigned char* B1 = p.B1 + (size_t)b * p.n_pad * 32;
    unsigned char* A2 = p.A2 + (size_t)b2 * p.m_pad * 32; unsigned char* B2 = p.B2 + (size_t)b2 * p.m_pad * 32;
    const int total = p.n_pad + p.m_pad;
    int it_pre = 0;
    for (int i = blockIdx.x * 256 + tid; i < total; i += gridDim.x * 256, ++it_pre) {
        const bool first = i < p.n_pad;
        const int r = first ? i : i - p.n_pad;
        const bool real = r < (first ? p.n : p.m);
        float ux = 0.f, uy = 0.f, uz = 0.f;
        float4 raw = make_float4(INFINITY, INFINITY, INFINITY, 0.f);      // padding: infinitely far in the exact pass
        if (real) {
            const float* s = (first ? P : Q) + 3 * (size_t)r;
            if (it_pre == 0) { raw.x = prx[0]; raw.y = pry[0]; raw.z = prz[0]; }
            else if (it_pre == 1) { raw.x = prx[1]; raw.y = pry[1]; raw.z = prz[1]; }
            else { raw.x = __ldg(s); raw.y = __ldg(s + 1); raw.z = __ldg(s + 2); }
            ux = (raw.x - cx) * sc; uy = (raw.y - cy) * sc; uz = (raw.z - cz) * sc;
        }
        if (first || own_q) {
            write_rows(first ? A1 : A2, first ? B1 : B2, r, real, ux, uy, uz, bias);
            (first ? p.T1 + (size_t)b * p.n_pad : p.T2 + (size_t)b2 * p.m_pad)[r] = raw;
        }
        if (real && p.packed1 != nullptr)
            (first ? p.packed1 + (size_t)b * p.n : p.packed2 + (size_t)b * p.m)[r] = ~0ull;
    }
    if (p.counters != nullptr && blockIdx.x == 0)
        for (int i = tid; i < p.n_counters; i += 256) p.counters[(size_t)b * p.n_counters + i] = 0;
    if (p.loss != nullptr && blockIdx.x == 0 && tid == 0) p.loss[b] = 0.f;
    pdl_tail_trigger_bit<3>();
}

// ---------------------------------------------------------------------------------------------
// tcgen05 helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t umma_smem_desc(const void* smem_ptr) {
    // K-major, SWIZZLE_NONE (interleaved 8x16B core matrices): LBO = 128 B between the two K chunks,
    // SBO = 256 B between 8-row groups, descriptor version 1 (sm_100)
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) >> 4) & 0x3FFF);
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((256u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16: A, B = F16 (0), D = F16 (0: one fp16 per 32-bit TMEM column), both K-major, M = 128, N = TC_N.
// A single K=16 instruction forms the whole distance, so the only fp16 rounding is the final one.
constexpr uint32_t TC_IDESC = (0u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 columns of fp16 accumulators -> 16 registers, two columns per register (even column in the low half)
__device__ __forceinline__ void tmem_ld16p(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// minimum of 16 packed words, per 16-bit lane (positive fp16 bit patterns order like their values):
// VIMNMX3.U16x2, four new elements per instruction -- twice the rate of the fp32 FMNMX3 (tools/micro/minbench.cu)
__device__ __forceinline__ uint32_t pmin16(const uint32_t* w) {
    uint32_t m0 = __vimin3_u16x2(w[0], w[1], w[2]), m1 = __vimin3_u16x2(w[3], w[4], w[5]);
    m0 = __vimin3_u16x2(m0, w[6], w[7]); m1 = __vimin3_u16x2(m1, w[8], w[9]);
    m0 = __vimin3_u16x2(m0, w[10], w[11]); m1 = __vimin3_u16x2(m1, w[12], w[13]);
    return __vimin3_u16x2(m0, m1, __vminu2(w[14], w[15]));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float min32(const float* v) {
    float m0 = min3(v[0], v[1], v[2]), m1 = min3(v[3], v[4], v[5]), m2 = min3(v[6], v[7], v[8]), m3 = min3(v[9], v[10], v[11]);
    m0 = min3(m0, v[12], v[13]); m1 = min3(m1, v[14], v[15]); m2 = min3(m2, v[16], v[17]); m3 = min3(m3, v[18], v[19]);
    m0 = min3(m0, v[20], v[21]); m1 = min3(m1, v[22], v[23]); m2 = min3(m2, v[24], v[25]); m3 = min3(m3, v[26], v[27]);
    m0 = min3(m0, v[28], v[29]); m1 = min3(m1, v[30], v[31]);
    return fminf(min3(m0, m1, m2), m3);
}
__device__ __forceinline__ float ref_sqdist_tc(float x1, float y1, float z1, float x2, float y2, float z2) {
    const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}
// NmDistanceKernel statement for statement (reference chamfer.cu:16-129): targets in batches of 512, the batch's first
// target initialises the running best (`k==0 || d<best`), the stored result is replaced only when strictly greater
// (`k2==0 || result>best`).  Used for samples with non-finite coordinates / hopeless conditioning, where the result
// depends on exactly this order (a NaN distance at a batch start hides the rest of that batch).
__device__ __forceinline__ void ref_order_nn(const float* __restrict__ T, int nt, float qx, float qy, float qz, float& res, int& res_i) {
    res = 0.f; res_i = 0;
    for (int k2 = 0; k2 < nt; k2 += 512) {
        const int end_k = min(nt, k2 + 512) - k2;
        float best = 0.f; int best_i = 0;
        for (int k = 0; k < end_k; ++k) {
            const float* t = T + 3 * (size_t)(k2 + k);
            const float d = ref_sqdist_tc(qx, qy, qz, __ldg(t), __ldg(t + 1), __ldg(t + 2));
            if (k == 0 || d < best) { best = d; best_i = k + k2; }
        }
        if (k2 == 0 || res > best) { res = best; res_i = best_i; }
    }
}
__device__ __forceinline__ void exact_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // the four exact warps

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
struct TcParams {
    const float* xyz1; const float* xyz2;
    const unsigned char* A1; const unsigned char* B1; const unsigned char* A2; const unsigned char* B2;
    const float4* T1; const float4* T2;
    const ChamferMeta* meta;
    float* dist1; float* dist2; int32_t* idx1; int32_t* idx2;
    int B, n, m, n_pad, m_pad;
    int tiles1, tiles2;      // query tiles per sample in direction 0 / 1
    int S1, S2;              // target-range splits per query tile in direction 0 / 1 (1 = whole range in one job)
    unsigned long long* packed1; unsigned long long* packed2; int* counters;    // merge of split jobs
    float* loss;             // fused loss (or NULL): (B) accumulators zeroed by the prep kernel
    int n_gt;                // distinct xyz2 samples (sample b uses xyz2[b % n_gt]); n_gt == B for the plain call
};

struct __align__(128) TcSmem {
    unsigned char a_tile[2][TC_TILE_BYTES];                    // double-buffered: the next job's queries arrive early
    unsigned char b_tile[TC_STAGES][TC_TILE_BYTES];
    float4 t4[TC_T4_BUFS][TC_SB_TARGETS];                      // raw target coordinates, per super-block
    uint32_t cm[2][TC_CM_WORDS * TC_TILE];                     // packed chunk minima of a super-block, [word][row]
    uint64_t full[TC_STAGES], empty[TC_STAGES], a_full[2], a_empty[2], tmem_full[TC_NBUF], tmem_empty[TC_NBUF], t4_full[TC_T4_BUFS], t4_empty[TC_T4_BUFS],
             cm_full[2], cm_empty[2];
    uint32_t tmem_base;
    int dbg[2];
    int last;                             // split jobs: this CTA finished the query tile's last sub-job
};

__device__ __forceinline__ float min16(const float* v) {
    float m0 = min3(v[0], v[1], v[2]), m1 = min3(v[3], v[4], v[5]);
    m0 = min3(m0, v[6], v[7]); m1 = min3(m1, v[8], v[9]);
    m0 = min3(m0, v[10], v[11]); m1 = min3(m1, v[12], v[13]);
    return min3(m0, m1, fminf(v[14], v[15]));
}

// ---- fused loss (train.py:68-69: mean(dist1,1) + mean(dist2,1)) -------------------------------------------------
// Every exact warp adds its 32 rows' distances (shuffles), scales by 1/n or 1/m and adds the result to
// loss[b] with ONE fire-and-forget float reduction (red.global.add.f32: no return value, nothing waits
// for it).  (A deterministic variant -- partial sums parked per (tile, warp), an acq_rel ticket counter,
// the last arrival summing in fixed order -- was measured 5 us slower at config A: the acquire holds the
// next job's loads back.)
__device__ __forceinline__ void loss_contribute(const TcParams& p, int b, int dir, int lane, float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if (lane == 0) atomicAdd(p.loss + b, v / (float)(dir ? p.m : p.n));
}

__global__ void __launch_bounds__(TC_THREADS, 2)
chamfer_tc_kernel(const TcParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int jobs_per_sample = p.tiles1 * p.S1 + p.tiles2 * p.S2;
    const int total_jobs = jobs_per_sample * p.B;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&S.a_full[i], 1); mbar_init(&S.a_empty[i], 1); }
        for (int i = 0; i < TC_NBUF; ++i) { mbar_init(&S.tmem_full[i], 1); mbar_init(&S.tmem_empty[i], TC_MIN_WARPS); }
        for (int i = 0; i < TC_T4_BUFS; ++i) { mbar_init(&S.t4_full[i], 1); mbar_init(&S.t4_empty[i], TC_EXACT_WARPS); }
        for (int i = 0; i < 2; ++i) { mbar_init(&S.cm_full[i], TC_MIN_WARPS); mbar_init(&S.cm_empty[i], TC_EXACT_WARPS); }
        fence_mbar_init();
#ifdef SPK_TIMING
        S.dbg[0] = 0; S.dbg[1] = 0;
#endif
    }
    if (warp == 1) {   // TMEM: 256 columns (2 x 128-column fp32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#ifdef SPK_TIMING
    const long long tk0 = clock64();
    __shared__ long long tlog_t[320]; __shared__ int tlog_a[320]; __shared__ int tlog_b[320]; __shared__ const char* tlog_s[320]; __shared__ int tlog_n;
    if (tid == 0) tlog_n = 0;
#define TCLOG(tag, a, b) do { if (blockIdx.x == 0 && (warp == 2 || warp == 6) && lane == 0) { const int ti_ = atomicAdd(&tlog_n, 1); if (ti_ < 320) { tlog_t[ti_] = clock64() - tk0; tlog_s[ti_] = tag; tlog_a[ti_] = (int)(a) + 1000 * (warp == 6); tlog_b[ti_] = (int)(b); } } } while (0)
#define TCLOGF(tag, a, b) do { if (blockIdx.x == 0 && lane == 0 && job_it == 1) { const int ti_ = atomicAdd(&tlog_n, 1); if (ti_ < 320) { tlog_t[ti_] = clock64() - tk0; tlog_s[ti_] = tag; tlog_a[ti_] = (int)(a); tlog_b[ti_] = (int)(b); } } } while (0)
#else
#define TCLOG(tag, a, b)
#define TCLOGF(tag, a, b)
#endif
    const uint32_t tmem_base = S.tmem_base;
    pdl_wait();                  // operands / metadata come from chamfer_prep_kernel

    // running counters (identical in every role): B-ring slots, accumulator buffers, super-blocks, jobs
    uint32_t ring_it = 0, acc_it = 0, sb_it = 0, job_it = 0;

    for (int job_id = blockIdx.x; job_id < total_jobs; job_id += gridDim.x, ++job_it) {
        const int b = (int)((unsigned)job_id / (unsigned)jobs_per_sample);
        int job = job_id - b * jobs_per_sample;
        const int dir = job < p.tiles1 * p.S1 ? 0 : 1;
        if (dir) job -= p.tiles1 * p.S1;
        const int NS = dir ? p.S2 : p.S1;
        const int nq = dir ? p.m : p.n, nt = dir ? p.n : p.m;
        const int nq_pad = dir ? p.m_pad : p.n_pad, nt_pad = dir ? p.n_pad : p.m_pad;
        const int n_sb_all = nt_pad / TC_SB_TARGETS;           // rows are padded to whole super-blocks
        int sb0 = 0, sb1 = n_sb_all;
        if (NS > 1) {                                          // (the common unsplit case pays no divisions)
            const int split = (int)((unsigned)job % (unsigned)NS);      // sub-jobs of one query tile are neighbours:
            job = (int)((unsigned)job / (unsigned)NS);                  // they run together and share the A tile in L2
            sb0 = split * n_sb_all / NS; sb1 = (split + 1) * n_sb_all / NS;
        }
        const int n_sb = sb1 - sb0;                            // super-blocks of this (sub-)job: [sb0, sb1)
        const int T = n_sb * TC_SB_TILES;                      // target tiles of this (sub-)job

        if (warp == 0) {
            // ===== TMA producer =====
            if (lane == 0) {
                const int b2 = b % p.n_gt, bq = dir ? b2 : b, bt = dir ? b : b2;     // cloud-2 arrays exist once per ground-truth sample
                const unsigned char* Aop = (dir ? p.A2 : p.A1) + ((size_t)bq * nq_pad + (size_t)job * TC_TILE) * 32;
                const unsigned char* Bop = (dir ? p.B1 : p.B2) + ((size_t)bt * nt_pad + (size_t)sb0 * TC_SB_TARGETS) * 32;
                const float4* T4 = (dir ? p.T1 : p.T2) + (size_t)bt * nt_pad + (size_t)sb0 * TC_SB_TARGETS;
                const uint32_t ab = job_it & 1;
                mbar_wait(&S.a_empty[ab], (uint32_t)(((job_it >> 1) & 1) ^ 1));   // the job two back has read this buffer
                mbar_expect_tx(&S.a_full[ab], TC_TILE_BYTES);
                bulk_g2s(S.a_tile[ab], Aop, TC_TILE_BYTES, &S.a_full[ab]);
                for (int sb = 0; sb < n_sb; ++sb) {
                    const uint32_t sbi = sb_it + sb, pb = sbi & 1;
                    constexpr int tiles = TC_SB_TILES;
                    for (int tt = 0; tt < tiles; ++tt) {
                        const uint32_t it = ring_it + sb * TC_SB_TILES + tt, s = it % TC_STAGES;
                        mbar_wait(&S.empty[s], (uint32_t)(((it / TC_STAGES) & 1) ^ 1));
                        mbar_expect_tx(&S.full[s], TC_TILE_BYTES);
                        bulk_g2s(S.b_tile[s], Bop + (size_t)(sb * TC_SB_TILES + tt) * TC_TILE_BYTES, TC_TILE_BYTES, &S.full[s]);
                    }
                    // raw coordinates for the exact pass: after the operand tiles, so that the exact warps (which
                    // lag up to two super-blocks behind the MMAs) never hold the operand stream back
                    const uint32_t tb = sbi % TC_T4_BUFS;
                    mbar_wait(&S.t4_empty[tb], (uint32_t)(((sbi / TC_T4_BUFS) & 1) ^ 1));
                    mbar_expect_tx(&S.t4_full[tb], (uint32_t)tiles * TC_TILE * 16u);
                    bulk_g2s(S.t4[tb], T4 + (size_t)sb * TC_SB_TARGETS, (uint32_t)tiles * TC_TILE * 16u, &S.t4_full[tb]);
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            if (lane == 0) {
                const uint32_t ab = job_it & 1;
                mbar_wait(&S.a_full[ab], (uint32_t)((job_it >> 1) & 1));
                const uint64_t a_desc = umma_smem_desc(S.a_tile[ab]);
                for (int t = 0; t < T; ++t) {
                    const uint32_t it = ring_it + t, s = it % TC_STAGES;
#if SPK_SPIN >= 1
                    mbar_wait_spin(&S.full[s], (uint32_t)((it / TC_STAGES) & 1));
#else
                    mbar_wait(&S.full[s], (uint32_t)((it / TC_STAGES) & 1));
#endif
#pragma unroll
                    for (int half = 0; half < TC_TILE / TC_N; ++half) {     // TC_TILE / TC_N MMAs per 128-target smem tile
                        const uint32_t ai = acc_it + (TC_TILE / TC_N) * t + half, buf = ai % TC_NBUF;
#if SPK_SPIN >= 1
                        mbar_wait_spin(&S.tmem_empty[buf], (uint32_t)(((ai / TC_NBUF) & 1) ^ 1));
#else
                        mbar_wait(&S.tmem_empty[buf], (uint32_t)(((ai / TC_NBUF) & 1) ^ 1));
#endif
                        TCLOGF("M empty ok", ai, 0);
                        tc_fence_after();
                        umma_f16(tmem_base + buf * TC_N, a_desc, umma_smem_desc(S.b_tile[s] + half * (TC_N * 32)), TC_IDESC);
                        umma_commit(&S.tmem_full[buf]);  // accumulator ready for the epilogue
                        TCLOGF("M issued+commit", ai, 0);
                    }
                    umma_commit(&S.empty[s]);            // smem slot free once both MMAs have read it
                }
                umma_commit(&S.a_empty[ab]);             // this A buffer may be replaced
            }
        } else if (warp < 2 + TC_MIN_WARPS) {
            // ===== min warps (4, one per TMEM lane quarter): accumulators -> packed chunk minima -> shared memory.
            // They never wait for the filter / exact pass, so the TMEM read path -- the resource that paces this
            // kernel -- stays busy; the exact warps follow up to two super-blocks behind.
            const int q = warp & 3;
            const int row = q * 32 + lane;                         // query row inside the tile = TMEM lane
            const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
            constexpr int ACC_PER_SB = TC_SB_TARGETS / TC_N;
            constexpr int WPA = TC_N / 32;                         // packed words per accumulator and row
            TCLOG("job start", job_id, n_sb);
            for (int sb = 0; sb < n_sb; ++sb) {
                const uint32_t sbi = sb_it + sb, pb = sbi & 1;
                mbar_wait(&S.cm_empty[pb], (uint32_t)(((sbi >> 1) & 1) ^ 1));   // the exact warps have read this buffer
                uint32_t* out = S.cm[pb] + row;
#pragma unroll
                for (int a = 0; a < ACC_PER_SB; ++a) {
                    const uint32_t ai = acc_it + (uint32_t)(sb * ACC_PER_SB + a), buf = ai % TC_NBUF;
#if SPK_SPIN >= 2
                    mbar_wait_spin(&S.tmem_full[buf], (ai / TC_NBUF) & 1);
#else
                    mbar_wait(&S.tmem_full[buf], (ai / TC_NBUF) & 1);
#endif
                    if (a == 0) TCLOG(" sb first acc ready", sb, 0);
                    if (warp == 2) TCLOGF("E full ok", ai, 0);
                    tc_fence_after();
                    const uint32_t ta = lane_addr + buf * TC_N;
                    uint32_t wv[WPA][16];                          // WPA x 32 columns of fp16 accumulators, two per register
#pragma unroll
                    for (int g = 0; g < WPA; ++g) tmem_ld16p(ta + 32 * g, wv[g]);
                    tmem_ld_wait();
                    if (warp == 2) TCLOGF("E ld done", ai, 0);
                    // the values are in registers: hand the accumulator back before reducing them
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&S.tmem_empty[buf]);
                    if (warp == 2) TCLOGF("E arrived", ai, 0);
#pragma unroll
                    for (int g = 0; g < WPA; ++g) out[(WPA * a + g) * TC_TILE] = pmin16(wv[g]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.cm_full[pb]);          // release: the warp's minima are visible to the exact warps
                TCLOG(" sb min pass done", sb, 0);
            }
        } else {
            // ===== exact warps (4): one thread per query row: filter the chunk minima, re-evaluate the survivors =====
            const int q = warp & 3;
            const int row = q * 32 + lane;                         // query row inside the tile
            const int gq = job * TC_TILE + row;                    // query index inside the cloud
            const bool live = gq < nq;
            const int b2 = b % p.n_gt;
            const float* Qx = (dir ? p.xyz2 + (size_t)b2 * nq * 3 : p.xyz1 + (size_t)b * nq * 3);
            const float* Tx = (dir ? p.xyz1 + (size_t)b * nt * 3 : p.xyz2 + (size_t)b2 * nt * 3);
            const ChamferMeta mt = p.meta[b];
            const float tau = mt.tau, scale2 = mt.scale2, bias = mt.bias;
            const bool eval_all = mt.nonfinite != 0.f;
            float qx = 0.f, qy = 0.f, qz = 0.f;
            if (live) { qx = __ldg(Qx + 3 * (size_t)gq); qy = __ldg(Qx + 3 * (size_t)gq + 1); qz = __ldg(Qx + 3 * (size_t)gq + 2); }
            // reference: the first target initialises the running best (`k==0 || d<best`, chamfer.cu:36)
            const int t_first = sb0 * TC_SB_TARGETS;                // first target of this (sub-)job: always a real point
            const float t0x = __ldg(Tx + 3 * (size_t)t_first), t0y = __ldg(Tx + 3 * (size_t)t_first + 1), t0z = __ldg(Tx + 3 * (size_t)t_first + 2);
            float best_d = 0.f;
            int best_i = t_first;
            // non-finite / hopelessly conditioned sample: the reference's own loop order decides (first sub-job only; the
            // super-block loop below then only keeps the pipeline's barriers moving)
            const bool ref_here = eval_all && sb0 == 0;
            if (ref_here && live) ref_order_nn(Tx, nt, qx, qy, qz, best_d, best_i);

            for (int sb = 0; sb < n_sb; ++sb) {
                const uint32_t sbi = sb_it + sb, pb = sbi & 1;
                constexpr int NW = TC_CM_WORDS;                             // 32 words = 64 chunk minima per row
                uint32_t cm[NW];
                mbar_wait(&S.cm_full[pb], (uint32_t)((sbi >> 1) & 1));
                {
                    const uint32_t* in = S.cm[pb] + row;
#pragma unroll
                    for (int i = 0; i < NW; ++i) cm[i] = in[i * TC_TILE];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.cm_empty[pb]);             // in registers: the buffer may be refilled
                // ---- filter: chunks whose minimum is within the error slack of the row minimum (or of the
                // exact best so far).  Everything is a positive fp16 pattern: 15-bit unsigned compares.
                uint32_t rm = __vimin3_u16x2(cm[0], cm[1], cm[2]);
#pragma unroll
                for (int i = 3; i + 1 < NW; i += 2) rm = __vimin3_u16x2(rm, cm[i], cm[i + 1]);
                rm = __vminu2(rm, cm[NW - 1]);
                const uint32_t r16 = min(rm & 0xFFFFu, rm >> 16);
                if (sb == 0 && !eval_all) best_d = ref_sqdist_tc(qx, qy, qz, t0x, t0y, t0z);
                // threshold: min(row minimum, exact best) widened by the fp16 rounding of the accumulator
                // (relative, 2^-8 = 4 ulp) and the error bound of the operands (absolute, tau), rounded UP to fp16
                float thr = fminf(__half2float(__ushort_as_half((unsigned short)r16)), fmaf(best_d, scale2, bias));
                thr = fmaf(thr, 1.00390625f, tau);
                uint32_t t16 = (uint32_t)__half_as_ushort(__float2half_ru(thr));
                if (!(thr == thr) || t16 > 0x7FFFu) t16 = 0x7FFFu;       // NaN / negative garbage: evaluate everything
                // mask bit i (i < 16): chunk in the low lane of word i, bit 16+i: its high lane; words 16..31 in the upper half
                uint32_t m_lo = 0, m_hi = 0;
                if (eval_all) {
                    m_lo = m_hi = 0u;                                 // (result already formed by ref_order_nn)
                } else {
                    // per lane (0x8000 | t) - c keeps bit 15 exactly when c <= t (both are 15-bit values: no borrow
                    // crosses the lanes); the bits 15 / 31 of word i go to mask bits i / 16+i
                    const uint32_t T2 = (t16 * 0x10001u) | 0x80008000u;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        m_lo |= ((T2 - cm[i]) >> (15 - i)) & (0x10001u << i);
                        m_hi |= ((T2 - cm[16 + i]) >> (15 - i)) & (0x10001u << i);
                    }
                }
                unsigned long long mask = live ? (((unsigned long long)m_hi << 32) | m_lo) : 0ull;
#ifdef SPK_TIMING
                atomicAdd(&S.dbg[0], __popcll(mask)); atomicAdd(&S.dbg[1], live ? 1 : 0);
#endif
                // ---- exact float32 re-evaluation of the surviving chunks (reference expression) ----
                const uint32_t tb = sbi % TC_T4_BUFS;
                mbar_wait(&S.t4_full[tb], (uint32_t)((sbi / TC_T4_BUFS) & 1));
                const float4* tsm = S.t4[tb];
                TCLOG(" sb filter done, t4 ready", sb, __popcll(mask));
                const int sb_base = (sb0 + sb) * TC_SB_TARGETS;
                while (mask) {
                    const int pbit = __ffsll((long long)mask) - 1;
                    mask &= mask - 1;
                    // word w -> accumulator w / 4, 32-column group w % 4; the high lane holds the odd columns = targets 16..31
                    const int w = ((pbit >> 5) << 4) | (pbit & 15), hi = (pbit >> 4) & 1;
                    const int l0 = w * 32 + hi * TC_CHUNK;                  // inside the super-block
                    // every chunk starts on a 256-byte boundary: rotate the visiting order by the lane so
                    // that the 8 lanes of a quarter-warp hit 8 different bank groups (no LDS.128 conflicts).
                    // All 16 exact distances first (FMA pipe), then ONE min tree and the lowest offset that
                    // attains it -- 2.5 ALU instructions per target instead of a compare/select chain.
                    float dv[TC_CHUNK];
                    int rj[TC_CHUNK];
#pragma unroll
                    for (int j = 0; j < TC_CHUNK; ++j) {
                        rj[j] = (j + lane) & (TC_CHUNK - 1);
                        const float4 tg = tsm[l0 + rj[j]];               // padding targets are +inf: never the minimum
                        dv[j] = ref_sqdist_tc(qx, qy, qz, tg.x, tg.y, tg.z);
                    }
                    const float dmin = min16(dv);                        // NaN distances are skipped by min
                    int rmin_off = 99;
#pragma unroll
                    for (int j = 0; j < TC_CHUNK; ++j) rmin_off = min(rmin_off, dv[j] == dmin ? rj[j] : 99);
                    const int t = sb_base + l0 + rmin_off;
                    // first minimum: smaller distance, or equal distance at a lower index
                    if (rmin_off < TC_CHUNK && (dmin < best_d || (dmin == best_d && t < best_i))) { best_d = dmin; best_i = t; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.t4_empty[tb]);
                TCLOG(" sb exact done", sb, 0);
            }
            if (NS == 1 && p.loss != nullptr)
                loss_contribute(p, b, dir, lane, live ? best_d : 0.f);
            if (live) {
                if (NS == 1 || eval_all) {
                    if (NS == 1 || ref_here) {                     // (split + non-finite: the first sub-job holds the whole answer)
                        ((dir ? p.dist2 : p.dist1) + (size_t)b * nq)[gq] = best_d;
                        ((dir ? p.idx2 : p.idx1) + (size_t)b * nq)[gq] = best_i;
                    }
                } else {
                    // distances are >= +0: their bit patterns order like the values, the index breaks ties
                    // downwards -> the 64-bit minimum over the sub-jobs IS the first minimum
                    atomicMin((dir ? p.packed2 : p.packed1) + (size_t)b * nq + gq,
                              ((unsigned long long)__float_as_uint(best_d) << 32) | (unsigned)best_i);
                }
            }
            if (NS > 1) {
                // the sub-job that arrives last at the query tile's counter unpacks the merged minima.
                // Ordering: the barrier orders the threads' atomics before the elected thread's
                // acq_rel increment (release, cumulative); the last arriver's increment acquires every
                // earlier sub-job's minima, the second barrier hands that to its other threads.
                exact_bar();
                if (warp == 2 + TC_MIN_WARPS && lane == 0) {
                    int* cnt = p.counters + (size_t)b * (p.tiles1 + p.tiles2) + (dir ? p.tiles1 : 0) + job;
                    int old;
                    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(cnt) : "memory");
                    S.last = (old == NS - 1);
                }
                exact_bar();
                float merged = 0.f;
                if (eval_all) {
                    if (ref_here && p.loss != nullptr) loss_contribute(p, b, dir, lane, live ? best_d : 0.f);
                } else if (S.last && live) {
                    unsigned long long v;
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"((dir ? p.packed2 : p.packed1) + (size_t)b * nq + gq) : "memory");
                    merged = __uint_as_float((unsigned)(v >> 32));
                    ((dir ? p.dist2 : p.dist1) + (size_t)b * nq)[gq] = merged;
                    ((dir ? p.idx2 : p.idx1) + (size_t)b * nq)[gq] = (int)(unsigned)v;
                }
                if (!eval_all && S.last && p.loss != nullptr)      // exactly one sub-job per query tile gets here
                    loss_contribute(p, b, dir, lane, merged);
                exact_bar();                                       // S.last is rewritten by the next split job
            }
        }
        if (warp >= 2) TCLOG("job end", job_id, 0);
        ring_it += (uint32_t)T; acc_it += (uint32_t)(TC_TILE / TC_N) * (uint32_t)T; sb_it += (uint32_t)n_sb;
    }

#ifdef SPK_TIMING
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) for (int i = 0; i < min(tlog_n, 320); ++i) printf("  t=%7lld %s %d %d\n", tlog_t[i], tlog_s[i], tlog_a[i], tlog_b[i]);
    if (tid == 0 && blockIdx.x < 2) printf("tc cta %d: chunks evaluated %d over %d (row, half, super-block) filters = %.3f each\n", blockIdx.x, S.dbg[0], S.dbg[1], (float)S.dbg[0] / (float)max(S.dbg[1], 1));
#endif
    pdl_tail_trigger_bit<2>();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
    }
}

}  // namespace dense

// host-side entry used by chamfer.cu -----------------------------------------------------------
using namespace dense;
static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

size_t chamfer_dense_workspace_bytes(int B, int n, int m) {
    const size_t n_pad = round_up(n, TC_SB_TARGETS), m_pad = round_up(m, TC_SB_TARGETS);
    const size_t tiles = (size_t)(n + TC_TILE - 1) / TC_TILE + (size_t)(m + TC_TILE - 1) / TC_TILE;
    return 2 * (size_t)B * (n_pad + m_pad) * 32 + (size_t)B * (n_pad + m_pad) * 16 +
           (((size_t)B * sizeof(ChamferMeta) + 255) & ~(size_t)255) + 256 +
           (size_t)B * ((size_t)n + m) * 8 + ((((size_t)B * tiles * 4) + 255) & ~(size_t)255) + 256;    // split-job merge
}

int chamfer_dense_forward(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1,
                       float* dist2, int32_t* idx1, int32_t* idx2, float* loss, void* ws, size_t ws_bytes,
                       cudaStream_t st, int B2) {
    if (B2 < 1 || B % B2 != 0) return fail(SPK_E_BADARG, "chamfer forward: %d samples do not divide into ground truths of %d", B, B2);
    if (ws_bytes < chamfer_dense_workspace_bytes(B, n, m) || ws == nullptr)
        return fail(SPK_E_WORKSPACE, "chamfer_fwd_f32: workspace of %zu bytes needed, %zu given", chamfer_dense_workspace_bytes(B, n, m), ws_bytes);
    if (((uintptr_t)ws & 15) != 0) return fail(SPK_E_ALIGN, "chamfer_fwd_f32: workspace must be 16-byte aligned");
    // rows are padded to whole super-blocks so that the target loop has no ragged tail
    const int n_pad = round_up(n, TC_SB_TARGETS), m_pad = round_up(m, TC_SB_TARGETS);
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    PrepParams pp;
    pp.xyz1 = xyz1; pp.xyz2 = xyz2; pp.n = n; pp.m = m; pp.n_pad = n_pad; pp.m_pad = m_pad;
    pp.meta = reinterpret_cast<ChamferMeta*>(base);
    unsigned char* ops = base + (((size_t)B * sizeof(ChamferMeta) + 255) & ~(size_t)255);
    pp.A1 = ops; pp.B1 = pp.A1 + (size_t)B * n_pad * 32;
    pp.A2 = pp.B1 + (size_t)B * n_pad * 32; pp.B2 = pp.A2 + (size_t)B * m_pad * 32;
    pp.T1 = reinterpret_cast<float4*>(pp.B2 + (size_t)B * m_pad * 32); pp.T2 = pp.T1 + (size_t)B * n_pad;
    // Fewer jobs than persistent CTAs (small batches: B=1 validation clouds): the target range of every
    // query tile is split into sub-jobs that merge through a 64-bit atomicMin; the last one to arrive
    // unpacks.  A sub-job costs ~1.5 us of merge on top of ~3 us per super-block (measured), so splitting
    // only pays when the grid is underfilled; the split minimising the makespan estimate wins.
    // (Splitting to even out the tail of config A -- 3.46 jobs per CTA -- was measured SLOWER: 45.8 vs 33.4 us.)
    const int tiles1 = (n + TC_TILE - 1) / TC_TILE, tiles2 = (m + TC_TILE - 1) / TC_TILE;
    const long long ctas = 2LL * sm_count();
    const long long jobs1 = (long long)(tiles1 + tiles2) * B;
    int split = 1;
    if (jobs1 < ctas) {
        const double sb_avg = 0.5 * (n_pad + m_pad) / TC_SB_TARGETS;          // super-blocks per unsplit job
        double best = 1e30;
        for (int c = 1; c <= 8; c *= 2) {
            const double waves = (double)((jobs1 * c + ctas - 1) / ctas);
            const double cost = waves * (3.0 * sb_avg / c + (c > 1 ? 1.5 : 0.0));
            if (cost < best - 1e-9) { best = cost; split = c; }
        }
    }
    if (const char* e = getenv("SPK_TC_SPLIT")) split = std::max(1, std::min(64, atoi(e)));
    const int S1 = std::max(1, std::min(split, m_pad / TC_SB_TARGETS));    // direction 0 scans xyz2
    const int S2 = std::max(1, std::min(split, n_pad / TC_SB_TARGETS));
    pp.packed1 = nullptr; pp.packed2 = nullptr; pp.counters = nullptr; pp.n_counters = tiles1 + tiles2;
    if (S1 > 1 || S2 > 1) {
        unsigned char* q = reinterpret_cast<unsigned char*>(pp.T2 + (size_t)B * m_pad);
        q = reinterpret_cast<unsigned char*>(((uintptr_t)q + 255) & ~(uintptr_t)255);
        pp.packed1 = reinterpret_cast<unsigned long long*>(q);
        pp.packed2 = pp.packed1 + (size_t)B * n;
        pp.counters = reinterpret_cast<int*>(pp.packed2 + (size_t)B * m);
    }
    pp.loss = loss; pp.n_gt = B2; pp.P = B / B2;
    const int slices = std::max(1, std::min(16, (n_pad + m_pad) / 512));
    SPK_CUDA(launch_k(chamfer_prep_kernel, dim3(slices, B), dim3(256), 0, st, pp));

    TcParams tp;
    tp.xyz1 = xyz1; tp.xyz2 = xyz2; tp.A1 = pp.A1; tp.B1 = pp.B1; tp.A2 = pp.A2; tp.B2 = pp.B2; tp.T1 = pp.T1; tp.T2 = pp.T2; tp.meta = pp.meta; tp.B = B;
    tp.dist1 = dist1; tp.dist2 = dist2; tp.idx1 = idx1; tp.idx2 = idx2;
    tp.n = n; tp.m = m; tp.n_pad = n_pad; tp.m_pad = m_pad;
    tp.tiles1 = tiles1; tp.tiles2 = tiles2; tp.S1 = S1; tp.S2 = S2;
    tp.packed1 = pp.packed1; tp.packed2 = pp.packed2; tp.counters = pp.counters;
    tp.loss = loss; tp.n_gt = B2;
    // request enough shared memory that at most 2 CTAs share an SM (each owns 256 of the 512 TMEM columns)
    const size_t smem = std::max(sizeof(TcSmem) + 128, (size_t)80 * 1024);
    SPK_CUDA(cudaFuncSetAttribute(chamfer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long jobs = ((long long)tp.tiles1 * S1 + (long long)tp.tiles2 * S2) * B;
    int grid = (int)std::min<long long>(jobs, 2LL * sm_count());         // persistent: 2 CTAs per SM
#ifdef SPK_EXPERIMENT
    if (const char* e = getenv("SPK_TC_GRID")) grid = std::max(1, std::min(grid, atoi(e)));
#endif
    SPK_CUDA(launch_k(chamfer_tc_kernel, dim3(grid), dim3(TC_THREADS), smem, st, tp));
    return SPK_OK;
}

}  // namespace spk
