// chamfer_tc.cu -- Chamfer forward: cell-sorted clouds, a tcgen05 chunk filter, exact fp32 refinement (sm_100a).
//
// Replaces NmDistanceKernel x2 (reference distance/chamfer/chamfer.cu:12-143) and returns bit-identical
// dist/idx (same float32 expression, first-minimum rule; non-finite samples follow the reference's 512-target
// batching statement for statement), without evaluating the n x m pair block.
//
// Round 1 evaluated every pair on the tensor pipe and was bound by TMEM: an accumulator column is occupied for
// the whole MMA -> commit -> tcgen05.ld -> release round trip (~600 cycles, tools/micro/tc_pipe2_bench.cu), so
// 512 columns x 128 lanes cap an SM at ~110 pair distances per cycle whatever the kernel does.  This version
// spends the tensor pipe on 16x fewer outputs:
//   prep   chamfer_sort_reg_kernel / chamfer_sort_kernel: one thread-block CLUSTER per sample, 1, 2 or 4 CTAs per
//          cloud.  Counting sort of the points by Hilbert cell (16^3 grid on the cloud's bounding box), so that
//          CHUNK = 16 consecutive points are neighbours.  Per chunk: bounding box, centre c, radius r.  Emits the
//          sorted points (x, y, z, original index) and the fp16 operand rows B'(c) of every CHUNK CENTRE as a
//          target (a query's row A'(p) is formatted by the search kernel).  Slices of <= 4096 points per CTA take the
//          register kernel (points in registers, scatter through distributed shared memory, one bulk copy out);
//          larger ones the generic kernel (scatter through global memory).
//   search chamfer_search_kernel: a job = 128 queries of one sample and direction; per pass of 128 chunks
//          (2048 targets) ONE M=128 x N=128 x K=16 tcgen05.mma gives, for every query, 128 values
//              V_j = bias + |p - c_j|^2 - 1.5 r_j^2      (scaled units, fp16 accumulators)
//          A chunk can only hold a point nearer than the best so far if |p - c_j| <= sqrt(best) + r_j, and
//          (s + r)^2 <= 3 s^2 + 1.5 r^2 (AM-GM), so "V_j <= bias + 3 best" is a NECESSARY condition -- one
//          packed 16-bit compare per chunk (3 integer instructions per two chunks).  Survivors are tested against
//          the chunk's bounding box in float32, and only then are the chunk's 16 points evaluated with the
//          reference's exact expression  d = fma(dz,dz, fma(dx,dx, dy*dy))  and first-minimum tie rule on the
//          ORIGINAL indices.  Every approximation errs on the side of evaluating more (slack: relative 2^-8 for the
//          fp16 accumulator, absolute tau for the fp16 operand splits), so the result equals brute force.
// ~100 exact distances per query instead of m (tools/chunk_search_estimate.py), independent of m for uniform data.
#include "spk_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

namespace spk {

constexpr int TC_TILE = 128;             // queries per job = TMEM lanes
constexpr int TC_CHUNK = 16;             // targets per chunk
constexpr int TC_NC = 128;               // chunks per pass = accumulator columns of one MMA
constexpr int TC_THREADS = 128;          // search kernel: one thread per query row
constexpr int TC_TILE_BYTES = TC_TILE * 32;
constexpr int SORT_THREADS = 1024;
constexpr int SORT_BITS = 4;             // Morton grid: 2^SORT_BITS cells per axis
constexpr int SORT_CELLS = 1 << (3 * SORT_BITS);
constexpr int TC_MAX_POINTS = 32768;     // per cloud: the sort keeps 6 bytes per point in shared memory
constexpr float TC_PAD_NORM = 30000.f;   // norm term of padding chunks: never pass the filter
constexpr float TC_RCAP_MAX = 1.f;        // largest 1.5 r^2 term folded into a chunk's operand (scaled units); the cap adapts per cloud

struct ChamferMeta {                     // per sample, written by the sort kernel
    float cx, cy, cz;                    // centre (bounding-box midpoint of both clouds)
    float scale;                         // power of two: |(x - c) * scale| <= 1
    float tau;                           // filter slack in scaled squared units
    float scale2;                        // scale * scale
    float bias;                          // (unused: the bias is per target cloud, ChamferGrid)
    float nonfinite;                     // != 0: NaN/inf coordinate or hopeless conditioning -> reference-order brute force
};

struct ChamferGrid {                     // per (sample, cloud): the Morton grid the cloud was sorted on
    float lo[3], inv[3];                 // cell = clamp((x - lo) * inv, 0, 2^SORT_BITS - 1) per axis
    float bias;                          // power of two added to every V of this cloud's chunks (keeps them positive fp16)
    float pad;
};

// ---------------------------------------------------------------------------------------------
// operand formatting
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
// byte offset of row r's first 16-byte K-chunk inside an operand array (UMMA canonical no-swizzle K-major layout:
// 8-row x 16-byte core matrices, 128 B between the two K chunks, 256 B between 8-row groups)
__device__ __forceinline__ size_t op_row_offset(int r) { return (size_t)(r >> 3) * 256 + (size_t)(r & 7) * 16; }
// Target rows are permuted inside every aligned group of 32: chunk u of the group sits in row
// ((u & 15) << 1) | (u >> 4), so that the EVEN accumulator columns of the group are chunks 0..15 and the ODD
// columns chunks 16..31 -- bit i / 16+i of a packed-compare mask then IS chunk i / 16+i.
__device__ __forceinline__ int b_row_of(int r) { return (r & ~31) | ((r & 15) << 1) | ((r >> 4) & 1); }

//  A'(p) = [-2xh,-2xh,-2xl, -2yh,-2yh,-2yl, -2zh,-2zh,-2zl,  nh,nm,nl,  1,1,1,  1   ]     n = |p|^2
//  B'(c) = [  xh,  xl,  xh,   yh,  yl,  yh,   zh,  zl,  zh,   1, 1, 1,  mh,mm,ml, bias]    m = |c|^2 - sub
__device__ __forceinline__ void write_a_row(unsigned char* opA, int r, bool real, float ux, float uy, float uz) {
    __align__(16) __half a[16];
    if (real) {
        __half xh, xl, yh, yl, zh, zl, nh, nm, nl;
        split2(ux, xh, xl); split2(uy, yh, yl); split2(uz, zh, zl);
        split3(fmaf(uz, uz, fmaf(uy, uy, ux * ux)), nh, nm, nl);
        const __half one = __float2half_rn(1.f), m2 = __float2half_rn(-2.f);
        a[0] = __hmul(m2, xh); a[1] = a[0]; a[2] = __hmul(m2, xl);
        a[3] = __hmul(m2, yh); a[4] = a[3]; a[5] = __hmul(m2, yl);
        a[6] = __hmul(m2, zh); a[7] = a[6]; a[8] = __hmul(m2, zl);
        a[9] = nh; a[10] = nm; a[11] = nl; a[12] = one; a[13] = one; a[14] = one; a[15] = one;
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __float2half_rn(0.f);
    }
    const size_t o = op_row_offset(r);
    *reinterpret_cast<uint4*>(opA + o) = *reinterpret_cast<const uint4*>(&a[0]);
    *reinterpret_cast<uint4*>(opA + o + 128) = *reinterpret_cast<const uint4*>(&a[8]);
}
__device__ __forceinline__ void write_b_row(unsigned char* opB, int chunk, bool real, float ux, float uy, float uz, float sub, float bias) {
    __align__(16) __half b[16];
    const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
    if (real) {
        __half xh, xl, yh, yl, zh, zl, nh, nm, nl;
        split2(ux, xh, xl); split2(uy, yh, yl); split2(uz, zh, zl);
        split3(fmaf(uz, uz, fmaf(uy, uy, ux * ux)) - sub, nh, nm, nl);
        b[0] = xh; b[1] = xl; b[2] = xh; b[3] = yh; b[4] = yl; b[5] = yh; b[6] = zh; b[7] = zl; b[8] = zh;
        b[9] = one; b[10] = one; b[11] = one; b[12] = nh; b[13] = nm; b[14] = nl; b[15] = __float2half_rn(bias);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) b[i] = zero;
        b[9] = one; b[12] = __float2half_rn(TC_PAD_NORM);          // a padding chunk is "infinitely" far
    }
    const size_t o = op_row_offset(b_row_of(chunk));
    *reinterpret_cast<uint4*>(opB + o) = *reinterpret_cast<const uint4*>(&b[0]);
    *reinterpret_cast<uint4*>(opB + o + 128) = *reinterpret_cast<const uint4*>(&b[8]);
}

// ---------------------------------------------------------------------------------------------
// prep: bounding boxes, Morton-cell counting sort, sorted points, operand rows, chunk boxes
// ---------------------------------------------------------------------------------------------
struct SortParams {
    const float* xyz[2];                 // (B, n[c], 3)
    int n[2], n_pad[2], nc[2], nc_pad[2];
    float4* S[2];                        // sorted points (x, y, z, original index bits), nc_pad * 16 per sample
    unsigned char* Bc[2];                // chunk-centre operand rows, nc_pad * 32 bytes per sample
    float4* box[2];                      // per chunk {lo.xyz, capped flag}, {hi.xyz, 0}; nc_pad * 2 per sample
    uint16_t* cellstart[2];              // first sorted position of every Morton cell, SORT_CELLS per sample
    ChamferGrid* grid[2];                // per sample
    ChamferMeta* meta;
    float* loss;                         // fused loss only: (B) accumulators, zeroed here
    int cl_ctas;                         // CTAs per cloud (the cluster holds 2 * cl_ctas)
    int pps_max;                         // points per slice of the larger cloud (shared-memory layout)
    int cps_max;                         // chunks per slice of the larger cloud
};

constexpr uint32_t spread3(uint32_t v) {           // 4 bits -> every third bit
    v = (v | (v << 8)) & 0x0000F00Fu;
    v = (v | (v << 4)) & 0x000C30C3u;
    v = (v | (v << 2)) & 0x00249249u;
    return v;
}
// Hilbert index of a cell (Skilling's axes-to-transpose, 3 axes x SORT_BITS bits): consecutive indices are face-adjacent
// cells, so a run of consecutive sorted points never straddles a jump of the curve (a Morton run does: its chunks came
// out with twice the radius and twice the candidates, tools/chunk_search_estimate.py).  Evaluated at COMPILE time into a
// SORT_CELLS-entry table: computing it per point was a third of the sort kernel's instructions.
constexpr uint32_t hilbert_of_cell(uint32_t x0, uint32_t x1, uint32_t x2) {
    constexpr uint32_t M = 1u << (SORT_BITS - 1);
    for (uint32_t Q = M; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
        if (x0 & Q) x0 ^= P;
        if (x1 & Q) x0 ^= P; else { const uint32_t t = (x0 ^ x1) & P; x0 ^= t; x1 ^= t; }
        if (x2 & Q) x0 ^= P; else { const uint32_t t = (x0 ^ x2) & P; x0 ^= t; x2 ^= t; }
    }
    x1 ^= x0; x2 ^= x1;
    uint32_t t = 0;
    for (uint32_t Q = M; Q > 1; Q >>= 1) if (x2 & Q) t ^= Q - 1;
    x0 ^= t; x1 ^= t; x2 ^= t;
    return (spread3(x0) << 2) | (spread3(x1) << 1) | spread3(x2);
}
struct HilbertTable { uint16_t v[SORT_CELLS]; };
constexpr HilbertTable make_hilbert_table() {
    HilbertTable t{};
    for (uint32_t i = 0; i < (uint32_t)SORT_CELLS; ++i)
        t.v[i] = (uint16_t)hilbert_of_cell(i >> (2 * SORT_BITS), (i >> SORT_BITS) & ((1u << SORT_BITS) - 1), i & ((1u << SORT_BITS) - 1));
    return t;
}
__device__ const HilbertTable g_hilbert = make_hilbert_table();
__device__ __forceinline__ uint32_t hilbert_code(uint32_t x0, uint32_t x1, uint32_t x2) {
    return __ldg(&g_hilbert.v[(x0 << (2 * SORT_BITS)) | (x1 << SORT_BITS) | x2]);
}
__device__ __forceinline__ int cell_of(float v, float lo, float inv) {
    const float t = (v - lo) * inv;
    int c = (t >= 0.f) ? (int)fminf(t, (float)((1 << SORT_BITS) - 1)) : 0;      // NaN -> 0
    return c;
}

// Thread 0 of every CTA of a sample's cluster: the common centre / scale / error bound of the two clouds (from their
// boxes in s_box: own lo, own hi, other lo, other hi) -> s_meta; the first CTA also writes the sample's ChamferMeta.
__device__ __forceinline__ void sort_write_meta(const SortParams& p, const float* s_box, float* s_meta, int* s_bad, bool writer, int b) {
    float c[3], ext = 0.f, amax = 0.f;
    for (int a = 0; a < 3; ++a) {
        const float L = fminf(s_box[a], s_box[6 + a]), H = fmaxf(s_box[3 + a], s_box[9 + a]);
        c[a] = 0.5f * L + 0.5f * H;
        ext = fmaxf(ext, fmaxf(H - c[a], c[a] - L));
        amax = fmaxf(amax, fmaxf(fabsf(L), fabsf(H)));
    }
    // scale = 2^-ceil(log2(ext)) so that |x - c| * scale <= 1; degenerate / non-finite boxes -> 1
    float scale = 1.f;
    if (ext > 0.f && ext < INFINITY) {
        int e = ((__float_as_int(ext) >> 23) & 255) - 126;       // ext = f * 2^e, f in [0.5, 1) (denormals end up at the clamp)
        e = max(-100, min(100, e));
        scale = __int_as_float((127 - e) << 23);
    }
    // error budget of V in scaled units: fp16 hi/lo products + fp32-grade accumulation (2^-17) plus the rounding
    // of the centred coordinates themselves (|x| * 2^-23 * scale each)
    const float delta = amax * scale * 1.1920929e-7f;
    const float e_tot = 7.62939453125e-6f + 16.f * delta;
    // the bias of the chunk rows (a power of two <= 1024) must stay above the radius cap + error bound
    const bool hopeless = !(TC_RCAP_MAX + 4.f * e_tot <= 1024.f) || !(scale * scale > 0.f) || !(scale * scale < INFINITY);
    s_meta[0] = c[0]; s_meta[1] = c[1]; s_meta[2] = c[2]; s_meta[3] = scale; s_meta[4] = 2.f * e_tot; s_meta[5] = scale * scale;
    const int sb = (*s_bad || hopeless) ? 1 : 0;
    *s_bad = sb;
    if (writer) {
        ChamferMeta mm; mm.cx = c[0]; mm.cy = c[1]; mm.cz = c[2]; mm.scale = scale; mm.tau = 2.f * e_tot;
        mm.scale2 = scale * scale; mm.bias = 0.f; mm.nonfinite = sb ? 1.f : 0.f;
        p.meta[b] = mm;
        if (p.loss != nullptr) p.loss[b] = 0.f;
    }
}

// One thread-block CLUSTER per sample: 2 clouds x CL CTAs (CL = p.cl_ctas = 1, 2 or 4, chosen so that the grid covers the
// machine once).  A cloud's CTA h owns the h-th slice of its POINTS for the bounding box, the cell histogram and the
// scatter, and the h-th slice of its CHUNKS for the chunk statistics and operand rows; partial boxes, cell counters and
// radius sums travel through distributed shared memory, the sorted points through global memory (visible cluster-wide
// after the release/acquire cluster barrier).
__global__ void __launch_bounds__(SORT_THREADS)
chamfer_sort_kernel(const SortParams p) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    __shared__ float red[6][32];
    __shared__ float s_part[8];                  // read by the peers: slice lo[3], hi[3], non-finite flag, chunk radius sum
    __shared__ float s_box[12];                  // own cloud lo/hi, other cloud lo/hi
    __shared__ float s_meta[8];
    __shared__ int s_bad;
    __shared__ float s_sum;
    __shared__ uint32_t warp_tot[32];
    pdl_trigger();
    pdl_wait();
#ifdef SPK_TIMING
    long long tph[8]; int nph = 0;
#define SORT_PHASE() do { __syncthreads(); tph[nph++] = clock64(); } while (0)
#else
#define SORT_PHASE()
#endif
    const int CL = p.cl_ctas;
    const int crank = blockIdx.x;                // == %cluster_ctarank: the cluster spans the grid's x extent
    const int cl = crank / CL, h = crank - cl * CL, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SORT_PHASE();
    const int n = p.n[cl], n_pad = p.n_pad[cl], nc = p.nc[cl], nc_pad = p.nc_pad[cl];
    const int pps = (((n + CL - 1) / CL) + 3) & ~3;                 // points per slice (multiple of 4: keeps the float4 path aligned)
    const int i0 = min(n, h * pps), i1 = min(n, i0 + pps);
    const int nch = n_pad / TC_CHUNK;                               // chunks that hold points or tile padding (multiple of 8)
    const int cps = (((nch + CL - 1) / CL) + 7) & ~7;               // chunks per slice (multiple of 8: whole warps below)
    const int c0 = min(nch, h * cps), c1 = min(nch, c0 + cps);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sort_smem);        // SORT_CELLS: this slice's points per cell
    uint32_t* base = hist + SORT_CELLS;                             // SORT_CELLS: first sorted position of this slice's points per cell
    uint32_t* cellrank = base + SORT_CELLS;                         // pps: (cell << 16) | rank inside (slice, cell)
    float4* cst = reinterpret_cast<float4*>(cellrank + p.pps_max);  // cps: per chunk (centre, 1.5 r^2 scaled)
    const float* X = p.xyz[cl] + (size_t)b * n * 3;
    for (int i = tid; i < SORT_CELLS; i += SORT_THREADS) hist[i] = 0;

    // ---- bounding box of the slice + non-finite detection -------------------------------------------------------------
    int bad = 0;
    {
        const float* Xs = X + (size_t)3 * i0;
        const int cnt = i1 - i0;
        float l[3] = {INFINITY, INFINITY, INFINITY}, hh[3] = {-INFINITY, -INFINITY, -INFINITY};
        auto upd = [&](int a, float v) { l[a] = fminf(l[a], v); hh[a] = fmaxf(hh[a], v); bad |= !(fabsf(v) < INFINITY); };
        int done = 0;
        if ((((uintptr_t)Xs) & 15) == 0) {                      // 4 points = 12 floats = 3 float4: the axis of every lane is static
            const int steps = cnt / 4;
            const float4* X4 = reinterpret_cast<const float4*>(Xs);
#pragma unroll 2
            for (int st = tid; st < steps; st += SORT_THREADS) {       // (two steps' loads in flight: the phase is latency-bound)
                const float4 a = __ldg(X4 + 3 * st), b4 = __ldg(X4 + 3 * st + 1), c4 = __ldg(X4 + 3 * st + 2);
                upd(0, a.x); upd(1, a.y); upd(2, a.z); upd(0, a.w);
                upd(1, b4.x); upd(2, b4.y); upd(0, b4.z); upd(1, b4.w);
                upd(2, c4.x); upd(0, c4.y); upd(1, c4.z); upd(2, c4.w);
            }
            done = steps * 4;
        }
        for (int i = done + tid; i < cnt; i += SORT_THREADS) {
            upd(0, __ldg(Xs + 3 * i)); upd(1, __ldg(Xs + 3 * i + 1)); upd(2, __ldg(Xs + 3 * i + 2));
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            for (int d = 16; d > 0; d >>= 1) {
                l[a] = fminf(l[a], __shfl_xor_sync(0xFFFFFFFFu, l[a], d));
                hh[a] = fmaxf(hh[a], __shfl_xor_sync(0xFFFFFFFFu, hh[a], d));
            }
            if (lane == 0) { red[a][warp] = l[a]; red[3 + a][warp] = hh[a]; }
        }
    }
    bad = __syncthreads_or(bad);
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            float v = red[k][lane];
            for (int d = 16; d > 0; d >>= 1) {
                const float o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
                v = k < 3 ? fminf(v, o) : fmaxf(v, o);
            }
            if (lane == 0) s_part[k] = v;
        }
        if (lane == 0) s_part[6] = bad ? 1.f : 0.f;
    }
    cluster_sync_all();                                          // #1 (also a CTA barrier): every slice's box is published
    if (tid < 12) {                                              // 0..5 own cloud, 6..11 other cloud
        const int which = tid < 6 ? cl : 1 - cl, k = tid < 6 ? tid : tid - 6;
        float v = k < 3 ? INFINITY : -INFINITY;
        for (int r = 0; r < CL; ++r) {
            const float o = ld_peer_f32(&s_part[k], (uint32_t)(which * CL + r));
            v = k < 3 ? fminf(v, o) : fmaxf(v, o);
        }
        s_box[tid] = v;
    } else if (tid == 32) {
        int any = 0;
        for (int r = 0; r < 2 * CL; ++r) any |= ld_peer_f32(&s_part[6], (uint32_t)r) != 0.f;
        s_bad = any;
    }
    __syncthreads();
    if (tid == 0) sort_write_meta(p, s_box, s_meta, &s_bad, crank == 0, b);
    __syncthreads();
    SORT_PHASE();                 // 1: bounding boxes + meta
    // every CTA of the cluster takes the same decision (same inputs); the peers may still be reading s_part
    if (s_bad) { cluster_sync_all(); return; }   // the search kernel walks the original arrays in reference order: nothing to sort
    const float cx = s_meta[0], cy = s_meta[1], cz = s_meta[2], sc = s_meta[3], sc2 = s_meta[5];

    // ---- counting sort by Hilbert cell of the cloud's own bounding box: this slice's counters -------------------------
    float inv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float e = s_box[3 + a] - s_box[a];
        inv[a] = e > 0.f ? (float)(1 << SORT_BITS) / e : 0.f;
        if (!(inv[a] < INFINITY)) inv[a] = 0.f;
    }
    for (int i = i0 + tid; i < i1; i += SORT_THREADS) {
        const float x = __ldg(X + 3 * i), y = __ldg(X + 3 * i + 1), z = __ldg(X + 3 * i + 2);
        const uint32_t code = hilbert_code((uint32_t)cell_of(x, s_box[0], inv[0]), (uint32_t)cell_of(y, s_box[1], inv[1]), (uint32_t)cell_of(z, s_box[2], inv[2]));
        const uint32_t rank = atomicAdd(&hist[code], 1u);
        cellrank[i - i0] = (code << 16) | rank;
    }
    cluster_sync_all();                                          // #2: every slice's counters are final
    SORT_PHASE();                 // 2: cells + histogram
    {   // exclusive scan of the cloud's SORT_CELLS counters (sum over its slices): SORT_CELLS / SORT_THREADS consecutive
        // cells per thread.  Inside a cell the slices' points follow each other in slice order.
        constexpr int PER = SORT_CELLS / SORT_THREADS;
        static_assert(PER == 4, "one 16-byte read of the peers' counters per thread");
        uint32_t v[PER] = {0u, 0u, 0u, 0u}, lower[PER] = {0u, 0u, 0u, 0u}, sum = 0;
        for (int r = 0; r < CL; ++r) {
            const uint4 c = (r == h) ? *reinterpret_cast<const uint4*>(hist + tid * PER) : ld_peer_u4(hist + tid * PER, (uint32_t)(cl * CL + r));
            v[0] += c.x; v[1] += c.y; v[2] += c.z; v[3] += c.w;
            if (r < h) { lower[0] += c.x; lower[1] += c.y; lower[2] += c.z; lower[3] += c.w; }
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) sum += v[k];
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t t = warp_tot[lane], it = t;
            for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, it, d); if (lane >= d) it += o; }
            warp_tot[lane] = it - t;
        }
        __syncthreads();
        uint32_t run = warp_tot[warp] + incl - sum;
        uint32_t st4[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) { st4[k] = run; base[tid * PER + k] = run + lower[k]; run += v[k]; }
        if (h == 0) {   // the search kernel starts every query at the chunk of its own cell ("home" chunk)
            uint16_t* cs = p.cellstart[cl] + (size_t)b * SORT_CELLS;
            *reinterpret_cast<uint2*>(cs + tid * PER) = make_uint2(st4[0] | (st4[1] << 16), st4[2] | (st4[3] << 16));
        }
    }
    __syncthreads();
    SORT_PHASE();                 // 3: scan + cell starts
    float4* S = p.S[cl] + (size_t)b * nc_pad * TC_CHUNK;
    for (int i = i0 + tid; i < i1; i += SORT_THREADS) {          // the slice's points go straight to their sorted rows
        const uint32_t cr = cellrank[i - i0];
        const float x = __ldg(X + 3 * i), y = __ldg(X + 3 * i + 1), z = __ldg(X + 3 * i + 2);
        S[base[cr >> 16] + (cr & 0xFFFFu)] = make_float4(x, y, z, __int_as_float(i));
    }
    if (tid == 0) s_sum = 0.f;
    cluster_sync_all();                                          // #3: the cloud's sorted rows are complete (global memory, cluster scope)
    SORT_PHASE();                 // 4: scatter
    // ---- this CTA's chunks: tile padding rows, per-chunk box / centre / radius -----------------------------------------
    unsigned char* Bc = p.Bc[cl] + (size_t)b * nc_pad * 32;
    float4* box = p.box[cl] + (size_t)b * nc_pad * 2;
    float sub_sum = 0.f;
    // four consecutive sorted positions per thread, a chunk = 4 consecutive lanes: 14 shuffles per 128 points (one point per
    // thread took 28 per 32).  4 (c1 - c0) is a multiple of 32: whole warps.
    for (int q = 4 * c0 + tid; q < 4 * c1; q += SORT_THREADS) {
        const int s0 = q * 4;
        float x[4], y[4], z[4];
        float l[3] = {INFINITY, INFINITY, INFINITY}, hh[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j] = INFINITY; y[j] = INFINITY; z[j] = INFINITY;
            if (s0 + j < n) {
                const float4 t = __ldcg(S + s0 + j);
                x[j] = t.x; y[j] = t.y; z[j] = t.z;
                l[0] = fminf(l[0], t.x); l[1] = fminf(l[1], t.y); l[2] = fminf(l[2], t.z);
                hh[0] = fmaxf(hh[0], t.x); hh[1] = fmaxf(hh[1], t.y); hh[2] = fmaxf(hh[2], t.z);
            } else {
                S[s0 + j] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7FFFFFFF));
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
            for (int d = 2; d > 0; d >>= 1) {
                l[a] = fminf(l[a], __shfl_xor_sync(0xFFFFFFFFu, l[a], d));
                hh[a] = fmaxf(hh[a], __shfl_xor_sync(0xFFFFFFFFu, hh[a], d));
            }
        const float ccx = 0.5f * l[0] + 0.5f * hh[0], ccy = 0.5f * l[1] + 0.5f * hh[1], ccz = 0.5f * l[2] + 0.5f * hh[2];
        float r2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (s0 + j < n) { const float dx = x[j] - ccx, dy = y[j] - ccy, dz = z[j] - ccz; r2 = fmaxf(r2, fmaf(dz, dz, fmaf(dx, dx, dy * dy))); }
        for (int d = 2; d > 0; d >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xFFFFFFFFu, r2, d));
        const int chunk = q >> 2;
        if ((lane & 3) == 0 && chunk < nc) {
            const float sub = 1.5f * r2 * sc2 * 1.001f;          // 1.5 r^2 in scaled units, inflated for the rounding of r2 itself
            cst[chunk - c0] = make_float4(ccx, ccy, ccz, sub);
            box[2 * chunk] = make_float4(l[0], l[1], l[2], 0.f);
            box[2 * chunk + 1] = make_float4(hh[0], hh[1], hh[2], 0.f);
            sub_sum += fminf(sub, 4.f);
        }
    }
    for (int d = 16; d > 0; d >>= 1) sub_sum += __shfl_xor_sync(0xFFFFFFFFu, sub_sum, d);
    if (lane == 0 && sub_sum != 0.f) atomicAdd(&s_sum, sub_sum);
    __syncthreads();
    if (tid == 0) s_part[7] = s_sum;
    cluster_sync_all();                                          // #4: every slice's radius sum is published
    if (tid == 0) {
        float t = 0.f;
        for (int r = 0; r < CL; ++r) t += ld_peer_f32(&s_part[7], (uint32_t)(cl * CL + r));    // same order in every CTA of the cloud
        s_sum = t;
    }
    __syncthreads();
    cluster_arrive();             // (the matching wait is the kernel's last statement: peers may still be reading s_part)
    SORT_PHASE();                 // 5: chunk stats
    // ---- chunk operand rows.  The radius term folded into a row is capped at 3x the cloud's mean (outliers: chunks that
    // straddle a sparse region) -- capped chunks are flagged and always go to the box test -- and the bias, a power of two
    // above the cap + error bound, keeps every V a POSITIVE fp16 (its bit pattern then orders like its value).
    const float cap = fminf(fmaxf(3.f * s_sum / (float)nc, 1.f / 1024.f), TC_RCAP_MAX);
    const float e4 = 2.f * s_meta[4];                             // 4 e_tot
    float bias = 1.f / 64.f;
    while (bias < cap + e4 && bias < 1024.f) bias *= 2.f;
    if (tid == 0 && h == 0) {
        ChamferGrid gi;
        for (int a = 0; a < 3; ++a) { gi.lo[a] = s_box[a]; gi.inv[a] = inv[a]; }
        gi.bias = bias; gi.pad = 0.f;
        p.grid[cl][b] = gi;
    }
    auto pad_chunk = [&](int c) {
        write_b_row(Bc, c, false, 0.f, 0.f, 0.f, 0.f, 0.f);
        box[2 * c] = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
        box[2 * c + 1] = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    };
    for (int c = c0 + tid; c < c1; c += SORT_THREADS) {
        if (c < nc) {
            const float4 ci = cst[c - c0];
            const bool capped = !(ci.w <= cap);
            write_b_row(Bc, c, true, (ci.x - cx) * sc, (ci.y - cy) * sc, (ci.z - cz) * sc, capped ? cap : ci.w, bias);
            if (capped) box[2 * c].w = 1.f;
        } else {
            pad_chunk(c);
        }
    }
    for (int c = nch + h * SORT_THREADS + tid; c < nc_pad; c += CL * SORT_THREADS) pad_chunk(c);   // whole-pass padding
    SORT_PHASE();                 // 6: chunk rows
#ifdef SPK_TIMING
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("sort kernel phases (cycles, %d CTAs per cloud): bbox+meta %lld, cells+hist %lld, scan %lld, scatter %lld, chunk stats %lld, chunk rows %lld\n",
               CL, tph[1] - tph[0], tph[2] - tph[1], tph[3] - tph[2], tph[4] - tph[3], tph[5] - tph[4], tph[6] - tph[5]);
#endif
    pdl_tail_trigger_bit<5>();
    cluster_wait();
}

// The same sort for slices of at most SORT_THREADS * 4 points (clouds of up to 16384 points with 4 CTAs each): every
// thread keeps its four points in REGISTERS from the one global read to the scatter, the Hilbert table sits in shared
// memory, and the scatter goes through distributed shared memory straight into the CTA that owns the sorted position --
// which then holds its slice of the sorted cloud in shared memory, writes it out with one bulk copy and computes the
// chunk statistics from it.  (In chamfer_sort_kernel the phases were bound by the latency of re-reading the points and
// of the round trip of the sorted rows through global memory, not by their work.)
__global__ void __launch_bounds__(SORT_THREADS)
chamfer_sort_reg_kernel(const SortParams p) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    __shared__ float red[6][32];
    __shared__ float s_part[8];                  // read by the peers: slice lo[3], hi[3], non-finite flag, chunk radius sum
    __shared__ float s_box[12];                  // own cloud lo/hi, other cloud lo/hi
    __shared__ float s_meta[8];
    __shared__ int s_bad;
    __shared__ float s_sum;
    __shared__ uint32_t warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CL = p.cl_ctas;
    uint32_t* hist = reinterpret_cast<uint32_t*>(sort_smem);        // SORT_CELLS: this slice's points per cell
    uint32_t* base = hist + SORT_CELLS;                             // SORT_CELLS: first sorted position of this slice's points per cell
    uint16_t* tab = reinterpret_cast<uint16_t*>(base + SORT_CELLS); // SORT_CELLS: Hilbert index of a cell
    float4* srt = reinterpret_cast<float4*>(tab + SORT_CELLS);      // cps_max * 16: this CTA's slice of the sorted cloud
    float4* cst = srt + (size_t)p.cps_max * TC_CHUNK;               // cps_max: per chunk (centre, 1.5 r^2 scaled)
    pdl_trigger();
    {   // constants and zeroing: before the wait on the previous kernel
        static_assert(SORT_CELLS == SORT_THREADS * 4, "one 8-byte table read and four counters per thread");
        reinterpret_cast<uint2*>(tab)[tid] = __ldg(reinterpret_cast<const uint2*>(g_hilbert.v) + tid);
        reinterpret_cast<uint4*>(hist)[tid] = make_uint4(0u, 0u, 0u, 0u);
    }
    pdl_wait();
#ifdef SPK_TIMING
    long long tph[8]; int nph = 0;
    long long tf[8]; int nf = 0;
#define RT() do { tf[nf++] = clock64(); } while (0)
#else
#define RT()
#endif
    const int crank = blockIdx.x;                // == %cluster_ctarank: the cluster spans the grid's x extent
    const int cl = crank / CL, h = crank - cl * CL, b = blockIdx.y;
    SORT_PHASE();
    const int n = p.n[cl], n_pad = p.n_pad[cl], nc = p.nc[cl], nc_pad = p.nc_pad[cl];
    const int pps = (((n + CL - 1) / CL) + 3) & ~3;                 // points per slice (<= 4 SORT_THREADS, host-checked)
    const int i0 = min(n, h * pps), i1 = min(n, i0 + pps);
    const int nch = n_pad / TC_CHUNK;                               // chunks that hold points or tile padding (multiple of 8)
    const int cps = (((nch + CL - 1) / CL) + 7) & ~7;               // chunks per slice (multiple of 8: whole warps below)
    const int c0 = min(nch, h * cps), c1 = min(nch, c0 + cps);
    const float* X = p.xyz[cl] + (size_t)b * n * 3;

    // ---- the thread's four points, bounding box of the slice, non-finite detection ---------------------------------------
    float px[4], py[4], pz[4];
    const int k0 = 4 * tid;                                         // first of them inside the slice
    const int nv = max(0, min(4, i1 - i0 - k0));
    {
        const float* Xs = X + (size_t)3 * (i0 + k0);
        if (nv == 4 && (((uintptr_t)Xs) & 15) == 0) {
            const float4* X4 = reinterpret_cast<const float4*>(Xs);
            const float4 a = __ldg(X4), b4 = __ldg(X4 + 1), c4 = __ldg(X4 + 2);
            px[0] = a.x; py[0] = a.y; pz[0] = a.z; px[1] = a.w; py[1] = b4.x; pz[1] = b4.y;
            px[2] = b4.z; py[2] = b4.w; pz[2] = c4.x; px[3] = c4.y; py[3] = c4.z; pz[3] = c4.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nv) { px[j] = __ldg(Xs + 3 * j); py[j] = __ldg(Xs + 3 * j + 1); pz[j] = __ldg(Xs + 3 * j + 2); }
                else { px[j] = 0.f; py[j] = 0.f; pz[j] = 0.f; }
        }
    }
    int bad = 0;
    {
        float l[3] = {INFINITY, INFINITY, INFINITY}, hh[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nv) {
                l[0] = fminf(l[0], px[j]); l[1] = fminf(l[1], py[j]); l[2] = fminf(l[2], pz[j]);
                hh[0] = fmaxf(hh[0], px[j]); hh[1] = fmaxf(hh[1], py[j]); hh[2] = fmaxf(hh[2], pz[j]);
                bad |= !(fabsf(px[j]) < INFINITY) | !(fabsf(py[j]) < INFINITY) | !(fabsf(pz[j]) < INFINITY);
            }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            for (int d = 16; d > 0; d >>= 1) {
                l[a] = fminf(l[a], __shfl_xor_sync(0xFFFFFFFFu, l[a], d));
                hh[a] = fmaxf(hh[a], __shfl_xor_sync(0xFFFFFFFFu, hh[a], d));
            }
            if (lane == 0) { red[a][warp] = l[a]; red[3 + a][warp] = hh[a]; }
        }
    }
    bad = __syncthreads_or(bad);
    if (warp < 6) {                                              // one warp per component
        float v = red[warp][lane];
        for (int d = 16; d > 0; d >>= 1) {
            const float o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
            v = warp < 3 ? fminf(v, o) : fmaxf(v, o);
        }
        if (lane == 0) s_part[warp] = v;
    } else if (tid == 6 * 32) {
        s_part[6] = bad ? 1.f : 0.f;
    }
    cluster_sync_all();                                          // #1 (also a CTA barrier): every slice's box is published
    if (tid < 12) {                                              // 0..5 own cloud, 6..11 other cloud
        const int which = tid < 6 ? cl : 1 - cl, k = tid < 6 ? tid : tid - 6;
        float v = k < 3 ? INFINITY : -INFINITY;
        for (int r = 0; r < CL; ++r) {
            const float o = ld_peer_f32(&s_part[k], (uint32_t)(which * CL + r));
            v = k < 3 ? fminf(v, o) : fmaxf(v, o);
        }
        s_box[tid] = v;
    } else if (tid == 32) {
        int any = 0;
        for (int r = 0; r < 2 * CL; ++r) any |= ld_peer_f32(&s_part[6], (uint32_t)r) != 0.f;
        s_bad = any;
    }
    __syncthreads();
    if (tid == 0) sort_write_meta(p, s_box, s_meta, &s_bad, crank == 0, b);
    __syncthreads();
    SORT_PHASE();                 // 1: bounding boxes + meta
    // every CTA of the cluster takes the same decision (same inputs); the peers may still be reading s_part
    if (s_bad) { cluster_sync_all(); return; }   // the search kernel walks the original arrays in reference order: nothing to sort
    const float cx = s_meta[0], cy = s_meta[1], cz = s_meta[2], sc = s_meta[3], sc2 = s_meta[5];

    // ---- counting sort by Hilbert cell of the cloud's own bounding box: this slice's counters -------------------------
    float inv[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float e = s_box[3 + a] - s_box[a];
        inv[a] = e > 0.f ? (float)(1 << SORT_BITS) / e : 0.f;
        if (!(inv[a] < INFINITY)) inv[a] = 0.f;
    }
    uint32_t cr[4];                                              // (cell << 16) | rank inside (slice, cell)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        cr[j] = 0;
        if (j < nv) {
            const uint32_t cell = ((uint32_t)cell_of(px[j], s_box[0], inv[0]) << (2 * SORT_BITS)) | ((uint32_t)cell_of(py[j], s_box[1], inv[1]) << SORT_BITS) |
                                  (uint32_t)cell_of(pz[j], s_box[2], inv[2]);
            const uint32_t code = tab[cell];
            cr[j] = (code << 16) | atomicAdd(&hist[code], 1u);
        }
    }
    cluster_sync_all();                                          // #2: every slice's counters are final
    SORT_PHASE();                 // 2: cells + histogram
    {   // exclusive scan of the cloud's SORT_CELLS counters (sum over its slices), four consecutive cells per thread.
        // Inside a cell the slices' points follow each other in slice order.
        uint32_t v[4] = {0u, 0u, 0u, 0u}, lower[4] = {0u, 0u, 0u, 0u}, sum = 0;
        for (int r = 0; r < CL; ++r) {
            const uint4 c = (r == h) ? *reinterpret_cast<const uint4*>(hist + tid * 4) : ld_peer_u4(hist + tid * 4, (uint32_t)(cl * CL + r));
            v[0] += c.x; v[1] += c.y; v[2] += c.z; v[3] += c.w;
            if (r < h) { lower[0] += c.x; lower[1] += c.y; lower[2] += c.z; lower[3] += c.w; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) sum += v[k];
        uint32_t incl = sum;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        uint32_t wt = warp_tot[lane], wincl = wt;                  // every warp scans the 32 warp totals itself (no second barrier)
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, wincl, d); if (lane >= d) wincl += o; }
        const uint32_t wexcl = __shfl_sync(0xFFFFFFFFu, wincl - wt, warp);
        uint32_t run = wexcl + incl - sum;
        uint32_t st4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { st4[k] = run; run += v[k]; }
        *reinterpret_cast<uint4*>(base + tid * 4) = make_uint4(st4[0] + lower[0], st4[1] + lower[1], st4[2] + lower[2], st4[3] + lower[3]);
        if (h == 0) {   // the search kernel starts every query at the chunk of its own cell ("home" chunk)
            uint16_t* cs = p.cellstart[cl] + (size_t)b * SORT_CELLS;
            *reinterpret_cast<uint2*>(cs + tid * 4) = make_uint2(st4[0] | (st4[1] << 16), st4[2] | (st4[3] << 16));
        }
    }
    __syncthreads();
    SORT_PHASE();                 // 3: scan + cell starts
    {   // every point goes to the shared memory of the CTA whose chunk slice holds its sorted position
        const uint32_t span = (uint32_t)cps * TC_CHUNK;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nv) {
                const uint32_t pos = base[cr[j] >> 16] + (cr[j] & 0xFFFFu);
                const uint32_t owner = pos / span;
                st_peer_f4(srt + (pos - owner * span), (uint32_t)(cl * CL) + owner, make_float4(px[j], py[j], pz[j], __int_as_float(i0 + k0 + j)));
            }
    }
    cluster_sync_all();                                          // #3: this CTA's slice of the sorted cloud is complete
    SORT_PHASE();                 // 4: scatter
    // ---- this CTA's chunks: sorted rows out (one bulk copy), tile padding rows, per-chunk box / centre / radius -------------
    float4* S = p.S[cl] + (size_t)b * nc_pad * TC_CHUNK;
    unsigned char* Bc = p.Bc[cl] + (size_t)b * nc_pad * 32;
    float4* box = p.box[cl] + (size_t)b * nc_pad * 2;
    const int real_rows = max(0, min(n, c1 * TC_CHUNK) - c0 * TC_CHUNK);     // rows of this slice that hold points
    if (tid == 0 && real_rows > 0) {
        fence_proxy_async_smem();                                // the peers' / own generic stores -> visible to the bulk engine
        bulk_s2g(S + (size_t)c0 * TC_CHUNK, srt, (uint32_t)real_rows * 16u);
        bulk_commit();
    }
    RT();
    float sub_sum = 0.f;
    // a chunk = 4 consecutive lanes, lane li takes its rows li, li + 4, li + 8, li + 12 (shared-memory reads of a warp then
    // cover whole 64-byte runs).  4 (c1 - c0) is a multiple of 32: whole warps.
    for (int q = tid; q < 4 * (c1 - c0); q += SORT_THREADS) {
        const int lc = q >> 2, li = q & 3, chunk = c0 + lc;
        float x[4], y[4], z[4];
        float l[3] = {INFINITY, INFINITY, INFINITY}, hh[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = lc * TC_CHUNK + li + 4 * j;
            x[j] = INFINITY; y[j] = INFINITY; z[j] = INFINITY;
            if (row < real_rows) {
                const float4 t = srt[row];
                x[j] = t.x; y[j] = t.y; z[j] = t.z;
                l[0] = fminf(l[0], t.x); l[1] = fminf(l[1], t.y); l[2] = fminf(l[2], t.z);
                hh[0] = fmaxf(hh[0], t.x); hh[1] = fmaxf(hh[1], t.y); hh[2] = fmaxf(hh[2], t.z);
            } else {
                S[(size_t)c0 * TC_CHUNK + row] = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7FFFFFFF));
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
            for (int d = 2; d > 0; d >>= 1) {
                l[a] = fminf(l[a], __shfl_xor_sync(0xFFFFFFFFu, l[a], d));
                hh[a] = fmaxf(hh[a], __shfl_xor_sync(0xFFFFFFFFu, hh[a], d));
            }
        const float ccx = 0.5f * l[0] + 0.5f * hh[0], ccy = 0.5f * l[1] + 0.5f * hh[1], ccz = 0.5f * l[2] + 0.5f * hh[2];
        float r2 = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (lc * TC_CHUNK + li + 4 * j < real_rows) { const float dx = x[j] - ccx, dy = y[j] - ccy, dz = z[j] - ccz; r2 = fmaxf(r2, fmaf(dz, dz, fmaf(dx, dx, dy * dy))); }
        for (int d = 2; d > 0; d >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xFFFFFFFFu, r2, d));
        if (li == 0 && chunk < nc) {
            const float sub = 1.5f * r2 * sc2 * 1.001f;          // 1.5 r^2 in scaled units, inflated for the rounding of r2 itself
            cst[lc] = make_float4(ccx, ccy, ccz, sub);
            box[2 * chunk] = make_float4(l[0], l[1], l[2], 0.f);
            box[2 * chunk + 1] = make_float4(hh[0], hh[1], hh[2], 0.f);
            sub_sum += fminf(sub, 4.f);
        }
    }
    RT();
    for (int d = 16; d > 0; d >>= 1) sub_sum += __shfl_xor_sync(0xFFFFFFFFu, sub_sum, d);
    if (lane == 0) red[0][warp] = sub_sum;                       // (a shared-memory float atomicAdd is a contended CAS loop)
    __syncthreads();
    RT();
    if (warp == 0) {
        float t = red[0][lane];
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, d);
        if (lane == 0) s_part[7] = t;
    }
    cluster_sync_all();                                          // #4: every slice's radius sum is published
    RT();
    if (tid == 0) {
        float t = 0.f;
        for (int r = 0; r < CL; ++r) t += ld_peer_f32(&s_part[7], (uint32_t)(cl * CL + r));    // same order in every CTA of the cloud
        s_sum = t;
    }
    __syncthreads();
    cluster_arrive();             // (the matching wait is the kernel's last statement: peers may still be reading s_part)
    SORT_PHASE();                 // 5: chunk stats
    // ---- chunk operand rows (cap and bias as in chamfer_sort_kernel) ------------------------------------------------------
    const float cap = fminf(fmaxf(3.f * s_sum / (float)nc, 1.f / 1024.f), TC_RCAP_MAX);
    const float e4 = 2.f * s_meta[4];                             // 4 e_tot
    float bias = 1.f / 64.f;
    while (bias < cap + e4 && bias < 1024.f) bias *= 2.f;
    if (tid == 0 && h == 0) {
        ChamferGrid gi;
        for (int a = 0; a < 3; ++a) { gi.lo[a] = s_box[a]; gi.inv[a] = inv[a]; }
        gi.bias = bias; gi.pad = 0.f;
        p.grid[cl][b] = gi;
    }
    auto pad_chunk = [&](int c) {
        write_b_row(Bc, c, false, 0.f, 0.f, 0.f, 0.f, 0.f);
        box[2 * c] = make_float4(INFINITY, INFINITY, INFINITY, 0.f);
        box[2 * c + 1] = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
    };
    for (int c = c0 + tid; c < c1; c += SORT_THREADS) {
        if (c < nc) {
            const float4 ci = cst[c - c0];
            const bool capped = !(ci.w <= cap);
            write_b_row(Bc, c, true, (ci.x - cx) * sc, (ci.y - cy) * sc, (ci.z - cz) * sc, capped ? cap : ci.w, bias);
            if (capped) box[2 * c].w = 1.f;
        } else {
            pad_chunk(c);
        }
    }
    for (int c = nch + h * SORT_THREADS + tid; c < nc_pad; c += CL * SORT_THREADS) pad_chunk(c);   // whole-pass padding
    SORT_PHASE();                 // 6: chunk rows
#ifdef SPK_TIMING
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("sort (register) kernel phases (cycles, %d CTAs per cloud): bbox+meta %lld, cells+hist %lld, scan %lld, scatter %lld, chunk stats %lld, chunk rows %lld\n",
               CL, tph[1] - tph[0], tph[2] - tph[1], tph[3] - tph[2], tph[4] - tph[3], tph[5] - tph[4], tph[6] - tph[5]);
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("  stats: bulk issue %lld, loop %lld, ->barrier %lld, ->csync4 %lld, ->end %lld\n", tf[0] - tph[4], tf[1] - tf[0], tf[2] - tf[1], tf[3] - tf[2], tph[5] - tf[3]);
#endif
    pdl_tail_trigger_bit<5>();
    if (tid == 0) bulk_wait<0>();                                // the sorted rows' bulk copy reads this CTA's shared memory
    cluster_wait();
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
// kind::f16: A, B = F16 (0), D = F16 (0: one fp16 per 32-bit TMEM column), both K-major, M = 128, N = TC_NC.
// A single K=16 instruction forms the whole value, so the only fp16 rounding is the final one.
constexpr uint32_t TC_IDESC = (0u << 4) | ((uint32_t)(TC_NC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float min16(const float* v) {
    float m0 = min3(v[0], v[1], v[2]), m1 = min3(v[3], v[4], v[5]);
    m0 = min3(m0, v[6], v[7]); m1 = min3(m1, v[8], v[9]);
    m0 = min3(m0, v[10], v[11]); m1 = min3(m1, v[12], v[13]);
    return min3(m0, m1, fminf(v[14], v[15]));
}
__device__ __forceinline__ float ref_sqdist_tc(float x1, float y1, float z1, float x2, float y2, float z2) {
    const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// NmDistanceKernel statement for statement (reference chamfer.cu:16-129): targets in batches of 512, the batch's
// first target initialises the running best (`k==0 || d<best`), the stored result is replaced only when strictly
// greater (`k2==0 || result>best`).  Used for samples with non-finite coordinates / hopeless conditioning, where the
// result depends on exactly this order (a NaN distance at a batch start hides the rest of that batch).
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

// ---------------------------------------------------------------------------------------------
// search kernel
// ---------------------------------------------------------------------------------------------
struct SearchParams {
    const float* xyz[2];
    int n[2], n_pad[2], nc[2], nc_pad[2];
    const float4* S[2]; const unsigned char* Bc[2]; const float4* box[2];
    const uint16_t* cellstart[2]; const ChamferGrid* grid[2];
    const ChamferMeta* meta;
    float* dist[2]; int32_t* idx[2];
    float* loss;                 // fused loss (or NULL): (B) accumulators zeroed by the sort kernel
    int B, tiles[2];
};

#ifdef SPK_TIMING
__device__ unsigned long long g_tc_dbg[8];      // 0 queries x passes, 1 step-1 mask bits, 2 step-2 mask bits, 3 box tests, 4 chunk evaluations, 5 warp-max evaluations
#define TC_COUNT(i, v) atomicAdd(&g_tc_dbg[i], (unsigned long long)(v))
#else
#define TC_COUNT(i, v)
#endif

constexpr int TC_PASS_TARGETS = TC_NC * TC_CHUNK;          // 2048 sorted targets per pass
constexpr int TC_QCAP = 1536;                              // candidate queue slots per pass (overflow: the owner evaluates at once)

// Shared memory of the search kernel.  The tail region holds, for a one-pass target cloud (<= 2048 points, STAGED), the
// cloud's chunk boxes and its sorted points -- every box test and exact distance then reads shared memory; for larger
// clouds only the chunk boxes of the current and the next pass (double-buffered): ~1 chunk per query and pass is
// evaluated there, straight from global memory / L2.
struct __align__(128) SearchSmem {
    unsigned char a_tile[TC_TILE_BYTES];         // query operand tile of the job whose MMAs are being issued (written by the CTA itself)
    unsigned char b_tile[2][TC_TILE_BYTES];      // chunk-centre operand tile, per step parity
    float4 q4[TC_TILE];                          // this job's queries (x, y, z, -)
    unsigned long long best[TC_TILE];            // per query (distance bits << 32) | original target index: its 64-bit minimum IS the first minimum
    uint16_t queue[TC_QCAP];                     // (query row << 7) | chunk of the pass: candidates that passed the box test
    uint64_t b_full[2], mma_done, t_full[2];
    uint32_t tmem_base;
    uint32_t big[4];                             // this pass's capped chunks (always box-tested)
    uint32_t qn;                                 // queue fill
    float4 tail[1];                              // (aligned up to 256 bytes at run time) STAGED: box[256] + tgt[2048];  streamed: box[2][256]
};
constexpr size_t TC_SMEM_STAGED = sizeof(SearchSmem) + 256 + (size_t)(TC_NC * 2 + TC_PASS_TARGETS) * 16;
constexpr size_t TC_SMEM_STREAM = sizeof(SearchSmem) + 256 + (size_t)(2 * TC_NC * 2) * 16;

// per lane (0x8000 | t) - v keeps bit 15 exactly when v <= t (both 15-bit values: no borrow crosses the lanes);
// bits 15 / 31 of word i go to mask bits i / 16+i  -> bit u of m[g] <=> chunk 32 g + u of the pass
__device__ __forceinline__ void build_mask(const uint32_t* w, uint32_t t16, uint32_t* m) {
    const uint32_t T2 = (t16 * 0x10001u) | 0x80008000u;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t mm = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) mm |= ((T2 - w[16 * g + i]) >> (15 - i)) & (0x10001u << i);
        m[g] = mm;
    }
}

// a query's operand row, straight into the shared-memory tile (canonical K-major core-matrix layout)
__device__ __forceinline__ void write_a_row_smem(unsigned char* tile, int r, float ux, float uy, float uz) {
    write_a_row(tile, r, true, ux, uy, uz);
}

template <bool STAGED>
__global__ void __launch_bounds__(TC_THREADS, 4)
chamfer_search_kernel(const SearchParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SearchSmem& S = *reinterpret_cast<SearchSmem*>(smem_raw);
    // (256-byte aligned: the staged points' XOR-swizzled addressing needs chunk bases with eight zero low bits)
    float4* const box_s = reinterpret_cast<float4*>((reinterpret_cast<uintptr_t>(S.tail) + 255) & ~(uintptr_t)255);   // STAGED: one buffer; streamed: [2][256]
    float4* const tgt_s = box_s + TC_NC * 2;                        // STAGED only
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int jobs_per_sample = p.tiles[0] + p.tiles[1];
    const int total_jobs = jobs_per_sample * p.B;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&S.b_full[i], 1); mbar_init(&S.t_full[i], 1); }
        mbar_init(&S.mma_done, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(TC_NC));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = S.tmem_base;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    pdl_wait();                  // operands / metadata come from chamfer_sort_kernel

    auto decode = [&](int job_id, int& b, int& dir, int& tile) {
        b = (int)((unsigned)job_id / (unsigned)jobs_per_sample);
        tile = job_id - b * jobs_per_sample;
        dir = tile < p.tiles[0] ? 0 : 1;
        if (dir) tile -= p.tiles[0];
    };
    // thread 0: the chunk-centre operand tile of step (job, pass)
    auto issue_b = [&](int job_id, int pass, uint32_t step) {
        int b, dir, tile; decode(job_id, b, dir, tile);
        const int tgt = 1 - dir;
        mbar_expect_tx(&S.b_full[step & 1], TC_TILE_BYTES);
        bulk_g2s(S.b_tile[step & 1], p.Bc[tgt] + ((size_t)b * p.nc_pad[tgt] + (size_t)pass * TC_NC) * 32, TC_TILE_BYTES, &S.b_full[step & 1]);
    };
    // thread 0: this pass's chunk boxes (+ the whole sorted cloud when STAGED)
    auto issue_t = [&](int job_id, int pass, uint32_t step) {
        int b, dir, tile; decode(job_id, b, dir, tile);
        const int tgt = 1 - dir;
        const int sl = STAGED ? 0 : (int)(step & 1);
        const uint32_t box_bytes = TC_NC * 2 * 16, tgt_bytes = STAGED ? TC_PASS_TARGETS * 16 : 0;
        mbar_expect_tx(&S.t_full[sl], box_bytes + tgt_bytes);
        bulk_g2s(box_s + sl * TC_NC * 2, p.box[tgt] + ((size_t)b * p.nc_pad[tgt] + (size_t)pass * TC_NC) * 2, box_bytes, &S.t_full[sl]);
        if (STAGED) bulk_g2s(tgt_s, p.S[tgt] + (size_t)b * p.nc_pad[tgt] * TC_CHUNK + (size_t)pass * TC_PASS_TARGETS, tgt_bytes, &S.t_full[sl]);
    };
    auto issue_mma = [&](uint32_t step) {
        mbar_wait(&S.b_full[step & 1], (step >> 1) & 1);
        tc_fence_after();
        umma_f16(tmem_base, umma_smem_desc(S.a_tile), umma_smem_desc(S.b_tile[step & 1]), TC_IDESC);
        umma_commit(&S.mma_done);
    };
    // the step after (job, pass) of this CTA's static job list; false when there is none
    auto next_step = [&](int& job_id, int& pass) -> bool {
        int b, dir, tile; decode(job_id, b, dir, tile);
        const int passes = p.nc_pad[1 - dir] / TC_NC;
        if (pass + 1 < passes) { ++pass; return true; }
        job_id += gridDim.x; pass = 0;
        return job_id < total_jobs;
    };
    // a job's query of this thread, in the SORTED order of its cloud: the 128 queries of a tile are neighbours, so they test and
    // evaluate the same few chunks (measured with queries in original order: 46 vs 36 us at 32 x 2048^2, 297 vs 164 us at
    // 8192^2 -- scattered box / chunk reads).  Non-finite samples are not sorted: original order.
    auto load_query = [&](int job_id, float& x, float& y, float& z, int& orig) {
        x = y = z = 0.f; orig = 0;
        if (job_id < total_jobs) {
            int b, dir, tile; decode(job_id, b, dir, tile);
            const int row = tile * TC_TILE + tid;
            if (row < p.n[dir]) {
                if (__ldg(&p.meta[b].nonfinite) != 0.f) {
                    const float* q = p.xyz[dir] + ((size_t)b * p.n[dir] + row) * 3;
                    x = __ldg(q); y = __ldg(q + 1); z = __ldg(q + 2); orig = row;
                } else {
                    const float4 q = __ldg(p.S[dir] + (size_t)b * p.nc_pad[dir] * TC_CHUNK + row);
                    x = q.x; y = q.y; z = q.z; orig = __float_as_int(q.w);
                }
            }
        }
    };
    // every thread: its query's operand row of job `job_id` into the (free) query tile
    auto format_a = [&](int job_id, float x, float y, float z) {
        int b, dir, tile; decode(job_id, b, dir, tile);
        const ChamferMeta mt = p.meta[b];
        const bool live = tile * TC_TILE + tid < p.n[dir];
        write_a_row(S.a_tile, tid, live, (x - mt.cx) * mt.scale, (y - mt.cy) * mt.scale, (z - mt.cz) * mt.scale);
        fence_proxy_async_smem();                                    // generic-proxy writes -> visible to the MMA's operand reads
    };

    // producer state (thread 0): the next step whose chunk-centre tile is loaded / whose boxes are loaded
    int ld_job = blockIdx.x, ld_pass = 0; uint32_t ld_step = 0; bool ld_valid = ld_job < total_jobs;
    int tl_job = blockIdx.x, tl_pass = 0; uint32_t tl_step = 0; bool tl_valid = tl_job < total_jobs;
    float nqx, nqy, nqz; int nqo;                                    // this thread's query of the NEXT job (prefetched)
    load_query(blockIdx.x, nqx, nqy, nqz, nqo);
    if (ld_valid) {
        if (tid == 0) {
            issue_t(tl_job, tl_pass, tl_step); tl_valid = next_step(tl_job, tl_pass); ++tl_step;
            if (!STAGED && tl_valid) { issue_t(tl_job, tl_pass, tl_step); tl_valid = next_step(tl_job, tl_pass); ++tl_step; }
            for (int i = 0; i < 2 && ld_valid; ++i) { issue_b(ld_job, ld_pass, ld_step); ld_valid = next_step(ld_job, ld_pass); ++ld_step; }
        }
        format_a(blockIdx.x, nqx, nqy, nqz);
        __syncthreads();
        if (tid == 0) issue_mma(0);
    }

    uint32_t step = 0;
    for (int job_id = blockIdx.x; job_id < total_jobs; job_id += gridDim.x) {
        int b, dir, tile; decode(job_id, b, dir, tile);
        const int tgt = 1 - dir;
        const int nq = p.n[dir], nt = p.n[tgt];
        const int passes = p.nc_pad[tgt] / TC_NC, nc_t = p.nc[tgt];
        const ChamferMeta mt = p.meta[b];
        const bool eval_all = mt.nonfinite != 0.f;
        const float* Tx = p.xyz[tgt] + (size_t)b * nt * 3;
        const float4* Sg = p.S[tgt] + (size_t)b * p.nc_pad[tgt] * TC_CHUNK;     // sorted targets (global)
        const int row = tile * TC_TILE + tid;                       // position in the sorted query cloud
        const bool live = row < nq;
        const float qx = nqx, qy = nqy, qz = nqz;
        const int qorig = nqo;
        load_query(job_id + gridDim.x, nqx, nqy, nqz, nqo);         // the next job's query: in registers long before it is needed
        // reference: the first target initialises the running best (`k==0 || d<best`, chamfer.cu:36)
        float best_d = ref_sqdist_tc(qx, qy, qz, __ldg(Tx), __ldg(Tx + 1), __ldg(Tx + 2));
        int best_i = 0;
        if (eval_all && live) ref_order_nn(Tx, nt, qx, qy, qz, best_d, best_i);
        // first minimum: smaller distance, or equal distance at a lower original index
        auto take = [&](float dmin, int imin) { if (dmin < best_d || (dmin == best_d && imin < best_i)) { best_d = dmin; best_i = imin; } };
        // the 16 points of a chunk against query (x, y, z): minimum distance and the lowest original index that attains it.
        // Shared memory (STAGED): every chunk starts on a 256-byte boundary, so the visiting order is XOR-swizzled by the lane
        // and the lanes of a quarter-warp hit different banks (one LOP3 per address).  Global memory otherwise.
        auto chunk_min = [&](bool in_smem, int jl, int jg, float x, float y, float z, float bound, float& dmin, int& imin) {
            float dv[TC_CHUNK];
            imin = 0x7FFFFFFF;
            if (in_smem) {
                const uint32_t base = smem_u32(tgt_s + jl * TC_CHUNK) | ((uint32_t)(lane & 15) << 4);
#pragma unroll
                for (int k = 0; k < TC_CHUNK; ++k) {
                    float tx, ty, tz, tw;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(tx), "=f"(ty), "=f"(tz), "=f"(tw) : "r"(base ^ (uint32_t)(k << 4)));
                    (void)tw;
                    dv[k] = ref_sqdist_tc(x, y, z, tx, ty, tz);      // padding points are +inf: never the minimum
                }
                dmin = min16(dv);
                if (dmin <= bound) {
                    // which points attain the minimum: almost always one -> one more shared-memory read for its original index
                    uint32_t eq = 0;
#pragma unroll
                    for (int k = 0; k < TC_CHUNK; ++k) eq |= (dv[k] == dmin) ? (1u << k) : 0u;
                    while (eq) {
                        const int k = __ffs((int)eq) - 1;
                        eq &= eq - 1;
                        int iw;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(iw) : "r"((base ^ (uint32_t)(k << 4)) + 12u));
                        imin = min(imin, iw);
                    }
                }
            } else {
                const float4* tg = Sg + (size_t)jg * TC_CHUNK;
#pragma unroll
                for (int k = 0; k < TC_CHUNK; ++k) { const float4 t = __ldg(tg + k); dv[k] = ref_sqdist_tc(x, y, z, t.x, t.y, t.z); }
                dmin = min16(dv);
                if (dmin <= bound) {
                    uint32_t eq = 0;
#pragma unroll
                    for (int k = 0; k < TC_CHUNK; ++k) eq |= (dv[k] == dmin) ? (1u << k) : 0u;
                    while (eq) { const int k = __ffs((int)eq) - 1; eq &= eq - 1; imin = min(imin, __float_as_int(__ldg(&tg[k].w))); }
                }
            }
            TC_COUNT(4, 1);
        };
        float tbias = 0.f;                                          // bias of the target cloud's chunk rows
        int home = 0;                                               // chunk of the query's own cell in the target cloud's grid
        if (!eval_all && live) {
            const ChamferGrid gi = p.grid[tgt][b];
            const uint32_t code = hilbert_code((uint32_t)cell_of(qx, gi.lo[0], gi.inv[0]), (uint32_t)cell_of(qy, gi.lo[1], gi.inv[1]), (uint32_t)cell_of(qz, gi.lo[2], gi.inv[2]));
            tbias = gi.bias;
            home = min((int)__ldg(p.cellstart[tgt] + (size_t)b * SORT_CELLS + code), nt - 1) >> 4;
            if (!STAGED || passes > 1) {   // a good first bound before the first pass: the points around the query's own cell (global memory)
                float dmin; int imin;
                chunk_min(false, 0, home, qx, qy, qz, best_d, dmin, imin);
                if (dmin <= best_d) take(dmin, imin);
            }
        }

        for (int pass = 0; pass < passes; ++pass, ++step) {
            const bool last_pass = pass == passes - 1;
            const int sl = STAGED ? 0 : (int)(step & 1);
            const float4* box_p = box_s + sl * TC_NC * 2;
            uint32_t w[64];                                        // 128 fp16 values V_j, two per register
            mbar_wait(&S.mma_done, step & 1);
            tc_fence_after();
#pragma unroll
            for (int g = 0; g < 4; ++g) tmem_ld16p(lane_addr + 32 * g, w + 16 * g);
            tmem_ld_wait();
            tc_fence_before();
            mbar_wait(&S.t_full[sl], STAGED ? (step & 1) : ((step >> 1) & 1));     // this pass's boxes (and targets) have landed
            {   // this pass's capped chunks (one flag per thread -> four ballot words)
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, box_p[2 * tid].w != 0.f);
                if (lane == 0) S.big[warp] = bal;
            }
            if (tid == 0) S.qn = 0;
            __syncthreads();                                       // accumulator in registers everywhere: TMEM and the operand stages are free
            if (tid == 0) {
                // the next pass of this job can go at once (same query tile); the next JOB's first MMA waits for its query tile,
                // which the CTA writes at the end of this pass
                if (!last_pass) issue_mma(step + 1);
                while (ld_valid && ld_step <= step + 2) { issue_b(ld_job, ld_pass, ld_step); ld_valid = next_step(ld_job, ld_pass); ++ld_step; }
            }
            const bool act = !eval_all && live;
            uint32_t over[4] = {0u, 0u, 0u, 0u};
            if (act) {
                if (STAGED && passes == 1) {                         // a good first bound: the points around the query's own cell
                    float dmin; int imin;
                    chunk_min(true, home, home, qx, qy, qz, best_d, dmin, imin);
                    if (dmin <= best_d) take(dmin, imin);
                }
                S.q4[tid] = make_float4(qx, qy, qz, 0.f);
                S.best[tid] = ((unsigned long long)__float_as_uint(best_d) << 32) | (unsigned)best_i;
                // ---- filter: V_j <= bias + 3 * scale2 * best (necessary for chunk j to hold a nearer point), plus slack ----
                float t = fmaf(3.f * mt.scale2, best_d, tbias);
                t = fmaf(t, 1.00390625f, mt.tau);
                uint32_t t16 = (uint32_t)__half_as_ushort(__float2half_ru(t));
                if (!(t == t) || t16 > 0x7C00u) t16 = 0x7C00u;       // NaN / garbage: everything passes
                uint32_t m[4] = {0u, 0u, 0u, 0u};
                bool any = true;
                if (passes > 1) {
                    // most passes of a large cloud hold no candidate at all for a query (its neighbours sit in one or two of
                    // them): the row minimum (22 packed min instructions) decides before the mask is built (192)
                    uint32_t rm = __vimin3_u16x2(w[0], w[1], w[2]);
#pragma unroll
                    for (int i = 3; i + 1 < 64; i += 2) rm = __vimin3_u16x2(rm, w[i], w[i + 1]);
                    rm = __vminu2(rm, w[63]);
                    any = min(rm & 0xFFFFu, rm >> 16) <= t16;
                }
                if (any) build_mask(w, t16, m);
                TC_COUNT(0, 1); TC_COUNT(2, __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]));
                // ---- box test (float32, this query's best so far); survivors go to the queue -------------------------------
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t mm = m[g] | S.big[g];
                    if (last_pass) {                                 // chunks past the cloud's end (only reachable when best is +inf)
                        const int lim = nc_t - pass * TC_NC - g * 32;
                        mm &= lim >= 32 ? 0xFFFFFFFFu : (lim <= 0 ? 0u : ((1u << lim) - 1u));
                    }
                    uint32_t pass_bits = 0;
                    while (mm) {
                        const int bit = __ffs((int)mm) - 1;
                        mm &= mm - 1;
                        const int jl = g * 32 + bit;
                        const float4 lo = box_p[2 * jl], hi = box_p[2 * jl + 1];
                        const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f), dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f),
                                    dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
                        const float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                        TC_COUNT(3, 1);
                        if (d2 * 0.999999f <= best_d) pass_bits |= 1u << bit;
                    }
                    const int cnt = __popc(pass_bits);
                    uint32_t pos = cnt ? atomicAdd(&S.qn, (uint32_t)cnt) : 0u;
                    while (pass_bits) {
                        const int bit = __ffs((int)pass_bits) - 1;
                        if (pos >= (uint32_t)TC_QCAP) { over[g] = pass_bits; break; }
                        pass_bits &= pass_bits - 1;
                        S.queue[pos++] = (uint16_t)((tid << 7) | (g * 32 + bit));
                    }
                }
            }
            __syncthreads();
            if (!eval_all) {
                // ---- every thread takes queued candidates round-robin: balanced whatever the per-query counts are.  The result
                // merges through a 64-bit shared-memory minimum: distances are >= +0, so their bit patterns order like the values
                // and the original index breaks ties downwards.
                auto run_item = [&](uint32_t item) {
                    const int r = (int)(item >> 7), jl = (int)(item & 127u);
                    const float4 q = S.q4[r];
                    const float bd = __uint_as_float((uint32_t)(S.best[r] >> 32));
                    float dmin; int imin;
                    chunk_min(STAGED, jl, pass * TC_NC + jl, q.x, q.y, q.z, bd, dmin, imin);
                    if (dmin <= bd) atomicMin(&S.best[r], ((unsigned long long)__float_as_uint(dmin) << 32) | (unsigned)imin);
                };
                const uint32_t total = min(S.qn, (uint32_t)TC_QCAP);
                for (uint32_t it = tid; it < total; it += TC_THREADS) run_item(S.queue[it]);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t mm = over[g];
                    while (mm) { const int bit = __ffs((int)mm) - 1; mm &= mm - 1; run_item((uint32_t)((tid << 7) | (g * 32 + bit))); }
                }
            }
            // the next job's query tile: this job's last MMA is long done, the tile is free
            if (last_pass && job_id + (int)gridDim.x < total_jobs) format_a(job_id + gridDim.x, nqx, nqy, nqz);
            __syncthreads();                                       // every thread is done with this pass's boxes / targets / queue
            if (act) {
                const unsigned long long v = S.best[tid];
                best_d = __uint_as_float((uint32_t)(v >> 32)); best_i = (int)(uint32_t)v;
            }
            if (tid == 0) {
                if (last_pass && job_id + (int)gridDim.x < total_jobs) issue_mma(step + 1);
                // boxes (+ targets): STAGED has one buffer, refilled now; streamed boxes run one pass ahead in the other buffer
                while (tl_valid && tl_step <= step + (STAGED ? 1u : 2u)) { issue_t(tl_job, tl_pass, tl_step); tl_valid = next_step(tl_job, tl_pass); ++tl_step; }
            }
        }
        if (p.loss != nullptr) {
            float v = live ? best_d : 0.f;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
            if (lane == 0) atomicAdd(p.loss + b, v / (float)nq);
        }
        if (live) {
            p.dist[dir][(size_t)b * nq + qorig] = best_d;
            p.idx[dir][(size_t)b * nq + qorig] = best_i;
        }
    }
    pdl_tail_trigger_bit<6>();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_NC));
    }
}

// host-side entry used by chamfer.cu -----------------------------------------------------------
static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct TcLayout {
    int n_pad[2], nc[2], nc_pad[2];
    size_t off_meta, off_S[2], off_Bc[2], off_box[2], off_cs[2], off_grid[2], total;
};
static TcLayout tc_layout(int B, int n, int m) {
    TcLayout L;
    const int cnt[2] = {n, m};
    size_t o = 0;
    L.off_meta = o; o += round_up((size_t)B * sizeof(ChamferMeta), 256);
    for (int c = 0; c < 2; ++c) {
        L.n_pad[c] = (int)round_up((size_t)cnt[c], TC_TILE);
        L.nc[c] = (cnt[c] + TC_CHUNK - 1) / TC_CHUNK;
        L.nc_pad[c] = (int)round_up((size_t)L.nc[c], TC_NC);
        L.off_S[c] = o; o += round_up((size_t)B * L.nc_pad[c] * TC_CHUNK * 16, 256);     // whole passes: a pass is one bulk copy
        L.off_Bc[c] = o; o += round_up((size_t)B * L.nc_pad[c] * 32, 256);
        L.off_box[c] = o; o += round_up((size_t)B * L.nc_pad[c] * 32, 256);
        L.off_cs[c] = o; o += round_up((size_t)B * SORT_CELLS * 2, 256);
        L.off_grid[c] = o; o += round_up((size_t)B * sizeof(ChamferGrid), 256);
    }
    L.total = o + 256;                    // + alignment of the caller's pointer
    return L;
}

bool chamfer_tc_supported(int n, int m) { return n >= 1 && m >= 1 && n <= TC_MAX_POINTS && m <= TC_MAX_POINTS; }

size_t chamfer_tc_workspace_bytes(int B, int n, int m) { return tc_layout(B, n, m).total; }

int chamfer_tc_forward(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1,
                       float* dist2, int32_t* idx1, int32_t* idx2, float* loss, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
    const TcLayout L = tc_layout(B, n, m);
    if (ws_bytes < L.total || ws == nullptr)
        return fail(SPK_E_WORKSPACE, "chamfer_fwd_f32: workspace of %zu bytes needed, %zu given", L.total, ws_bytes);
    if (((uintptr_t)ws & 15) != 0) return fail(SPK_E_ALIGN, "chamfer_fwd_f32: workspace must be 16-byte aligned");
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    SortParams sp;
    SearchParams qp;
    const float* xyz[2] = {xyz1, xyz2};
    const int cnt[2] = {n, m};
    for (int c = 0; c < 2; ++c) {
        sp.xyz[c] = xyz[c]; sp.n[c] = cnt[c]; sp.n_pad[c] = L.n_pad[c]; sp.nc[c] = L.nc[c]; sp.nc_pad[c] = L.nc_pad[c];
        sp.S[c] = reinterpret_cast<float4*>(base + L.off_S[c]);
        sp.Bc[c] = base + L.off_Bc[c];
        sp.box[c] = reinterpret_cast<float4*>(base + L.off_box[c]);
        sp.cellstart[c] = reinterpret_cast<uint16_t*>(base + L.off_cs[c]);
        sp.grid[c] = reinterpret_cast<ChamferGrid*>(base + L.off_grid[c]);
        qp.xyz[c] = xyz[c]; qp.n[c] = cnt[c]; qp.n_pad[c] = L.n_pad[c]; qp.nc[c] = L.nc[c]; qp.nc_pad[c] = L.nc_pad[c];
        qp.S[c] = sp.S[c]; qp.Bc[c] = sp.Bc[c]; qp.box[c] = sp.box[c];
        qp.cellstart[c] = sp.cellstart[c]; qp.grid[c] = sp.grid[c];
        qp.tiles[c] = (cnt[c] + TC_TILE - 1) / TC_TILE;
    }
    sp.meta = reinterpret_cast<ChamferMeta*>(base + L.off_meta); sp.loss = loss;
    qp.meta = sp.meta; qp.loss = loss; qp.B = B;
    qp.dist[0] = dist1; qp.dist[1] = dist2; qp.idx[0] = idx1; qp.idx[1] = idx2;

    const int nmax = std::max(n, m);
    // CTAs per cloud (1, 2, 4): enough for the register kernel (a slice of at most 4 points per thread) when the clouds
    // allow it, then as many as keep the whole grid within one wave of 1024-thread CTAs and a slice worth a CTA
    int CL = 1;
    while (CL < 4 && nmax > CL * SORT_THREADS * 4) CL *= 2;
    while (CL < 4 && 2 * B * CL * 2 <= 160 && nmax / (CL * 2) >= 2048) CL *= 2;
    if (const char* e = getenv("SPK_SORT_CTAS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) CL = v; }
    sp.cl_ctas = CL;
    sp.pps_max = (((nmax + CL - 1) / CL) + 3) & ~3;
    sp.cps_max = (((int)(round_up((size_t)nmax, TC_TILE) / TC_CHUNK) + CL - 1) / CL + 7) & ~7;
    bool reg_kernel = sp.pps_max <= SORT_THREADS * 4;
    if (const char* e = getenv("SPK_SORT_KERNEL")) { if (strcmp(e, "generic") == 0) reg_kernel = false; }
    const size_t sort_smem = reg_kernel ? (size_t)SORT_CELLS * 10 + (size_t)sp.cps_max * (TC_CHUNK + 1) * 16
                                        : (size_t)SORT_CELLS * 8 + (size_t)sp.pps_max * 4 + (size_t)sp.cps_max * 16;
    // function attributes are per device; set once per (device, size) -- benign race: every writer stores the same value
    int dev = 0;
    SPK_CUDA(cudaGetDevice(&dev));
    static size_t sort_smem_set[64] = {0}, sort_reg_smem_set[64] = {0};
    static bool search_attr[64] = {false};
    const int dslot = (dev >= 0 && dev < 64) ? dev : 0;
    if (!reg_kernel && (sort_smem > sort_smem_set[dslot] || dev != dslot)) {
        SPK_CUDA(cudaFuncSetAttribute(chamfer_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
        sort_smem_set[dslot] = sort_smem;
    }
    if (reg_kernel && (sort_smem > sort_reg_smem_set[dslot] || dev != dslot)) {
        SPK_CUDA(cudaFuncSetAttribute(chamfer_sort_reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_smem));
        sort_reg_smem_set[dslot] = sort_smem;
    }
    {   // one cluster per sample (2 clouds x CL CTAs), programmatic dependent launch like every kernel of the library
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * CL, B); cfg.blockDim = dim3(SORT_THREADS); cfg.dynamicSmemBytes = sort_smem; cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)(2 * CL); at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        SPK_CUDA(cudaLaunchKernelEx(&cfg, reg_kernel ? chamfer_sort_reg_kernel : chamfer_sort_kernel, sp));
    }

    // dynamic shared memory of at least 50 KB, so that at most 4 CTAs share an SM (each owns 128 of the 512 TMEM columns)
    // every pass's boxes + sorted points are staged in shared memory (one buffer).  The streamed variant (boxes only, points
    // from global memory / L2 when evaluated: chamfer_search_kernel<false>) was measured SLOWER: 192 vs 164 us at 32 x 8192^2.
    const bool staged = true;
    const size_t smem = std::max(staged ? TC_SMEM_STAGED : TC_SMEM_STREAM, (size_t)50 * 1024);
    if (!search_attr[dslot] || dev != dslot) {
        SPK_CUDA(cudaFuncSetAttribute(chamfer_search_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(TC_SMEM_STAGED, (size_t)50 * 1024)));
        SPK_CUDA(cudaFuncSetAttribute(chamfer_search_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(TC_SMEM_STREAM, (size_t)50 * 1024)));
        search_attr[dslot] = true;
    }
    const long long jobs = (long long)(qp.tiles[0] + qp.tiles[1]) * B;
    const int grid = (int)std::min<long long>(jobs, 4LL * sm_count());
    if (staged) SPK_CUDA(launch_k(chamfer_search_kernel<true>, dim3(grid), dim3(TC_THREADS), smem, st, qp));
    else SPK_CUDA(launch_k(chamfer_search_kernel<false>, dim3(grid), dim3(TC_THREADS), smem, st, qp));
    return SPK_OK;
}

}  // namespace spk

#ifdef SPK_TIMING
extern "C" void spk_debug_tc_counters(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, spk::g_tc_dbg, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(spk::g_tc_dbg, z, sizeof(z)); }
}
#endif
