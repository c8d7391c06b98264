/*
 * softpool_b200.h -- C ABI of libsoftpool_b200.so (sm_100a)
 *
 * The drop-in boundary for the SoftPool sort/top-k + gather hot path and the Chamfer
 * distance of wangyida/softpool (reference @ 31a2d18).  Plain `extern "C"`, raw device
 * pointers and sizes, no torch / ATen / pybind types.  Each entry point names the
 * reference interface it replaces (file:line, relative to the reference root).
 *
 * Conventions (all entry points)
 *   - every pointer is a DEVICE pointer to a contiguous, naturally aligned buffer owned by
 *     the caller (PyTorch's caching allocator in the Python host layer); the library
 *     allocates nothing and keeps no mutable global state (thread-/device-re-entrant:
 *     the reference is driven from one Python thread per GPU under nn.DataParallel,
 *     train.py:228);
 *   - `stream` is the caller's CUDA stream (a cudaStream_t passed as void*); work is
 *     enqueued on it and the call returns without synchronising.  (The reference launches on
 *     the legacy default stream with no device guard, distance/chamfer/chamfer.cu:142.)
 *     The caller has made the buffers' device current;
 *   - return value: SPK_OK (0) or a negative SPK_E_* / positive cudaError_t code; never
 *     printf/exit (the reference printf()s and returns 0/1 that Python ignores,
 *     chamfer.cu:145-151, dist_chamfer.py:30).  spk_last_error() gives a thread-local text;
 *   - inputs are never modified; outputs are fully overwritten (no "must arrive zeroed"
 *     precondition, unlike chamfer.cu:166-171).
 */
#ifndef SOFTPOOL_B200_H_
#define SOFTPOOL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPK_ABI_VERSION 3   /* 2: + chamfer_fwd_loss_f32; 3: + chamfer_fwd_multi_f32 */

enum {
    SPK_OK = 0,
    SPK_E_BADARG = -1,      /* null pointer / non-positive size / k > N / k < cab ...      */
    SPK_E_UNSUPPORTED = -2, /* size outside what the sm_100a kernels are built for        */
    SPK_E_WORKSPACE = -3,   /* workspace too small (see the *_workspace_bytes functions)  */
    SPK_E_ALIGN = -4,       /* a pointer is not aligned as documented                     */
    SPK_E_INTERNAL = -5     /* SPK_CHAMFER_SELFCHECK=1 found a result that differs from the plain kernel */
};

int spk_abi_version(void);
/* Thread-local description of the last non-zero return on this thread ("" if none). */
const char* spk_last_error(void);

/* ------------------------------------------------------------------------------------
 * SoftPool
 * ---------------------------------------------------------------------------------- */

/* Per-region descending top-k of the activation rows, plus every per-sample index product of
 * the forward, in ONE launch.
 * Replaces the R x `torch.sort(val_activa[:, region, :], dim=1, descending=True)` +
 * `x_idx[:, :pnt_per_sort]` of softpool.py:139-142, the float index cube of
 * softpool.py:136-137,146-147 and `torch.argmax(val_activa, dim=1)` of softpool.py:95.
 *   keys      (B,R,N) f32       val_activa, output of Sorter.conv1d (softpool.py:94)
 *   idx       (B,R,k) i32       first k entries of the STABLE descending argsort of every row:
 *                               ties keep ascending original index, NaN sorts first, -0 == +0
 *                               (the order torch's CPU sort produces -- the parity contract)
 *   sp_idx    (B,R+3,R,k) f32   the index cube, float32 and (R+3)-fold replicated exactly as
 *                               the reference builds it; NULL to skip
 *   id_activa (B,N) i64         first maximum over R wins, a NaN wins outright; NULL to skip
 * Limits: 1 <= k <= N <= 16384 (one row lives in shared memory), N < 2^24 (float index).  */
int sp_topk_f32(const float* keys, int B, int R, int N, int k, int32_t* idx, float* sp_idx,
                int64_t* id_activa, void* stream);

/* `id_activa = torch.argmax(val_activa, dim=1)` of Sorter.forward (softpool.py:95) alone.  */
int sp_argmax_i64(const float* keys, int B, int R, int N, int64_t* id_activa, void* stream);

/* Gather of all C feature channels by the top-k indices, fused with the window max.
 * Replaces softpool.py:142-145 (index repeat + torch.gather + slice-assign into sp_cube) and
 * train2cabins (softpool.py:71-85, called :151).
 *   x        (B,C,N) f32
 *   idx      (B,R,k) i32         from sp_topk_f32 (every entry in [0,N))
 *   sp_cube  (B,C,R,k) f32       sp_cube[b,c,r,j] = x[b,c,idx[b,r,j]]
 *   cabins   (B,C,R,cab) f32     max over the cab windows of k/cab consecutive slots (the
 *                                trailing k % cab slots are ignored); torch.max semantics:
 *                                first maximum, a NaN wins
 *   cab_arg  (B,C,R,cab) u16     slot offset INSIDE its window of each window's arg-max, saved
 *                                for the backward (autograd keeps the same from torch.max)
 * Requires cab >= 1, k >= cab, k/cab <= 65535; `cabins` and `cab_arg` may both be NULL (then
 * cab is ignored).                                                                       */
int sp_gather_fwd_f32(const float* x, const int32_t* idx, int B, int C, int N, int R, int k,
                      int cab, float* sp_cube, float* cabins, uint16_t* cab_arg, void* stream);

/* Backward of the gather + window max (the reference has no hand-written backward; this is
 * what autograd derives from softpool.py:139-151: GatherBackward -> scatter_add, CopySlices,
 * MaxBackward).
 *   g_cube    (B,C,R,k) f32      upstream gradient on sp_cube
 *   g_cabins  (B,C,R,cab) f32    upstream gradient on cabins, or NULL
 *   cab_arg   (B,C,R,cab) u16    from the forward (ignored when g_cabins is NULL)
 *   grad_x    (B,C,N) f32        fully overwritten (dense, zeros where nothing was selected);
 *                                contributions are summed in ascending region order -> the
 *                                result is deterministic (no atomics)                     */
int sp_gather_bwd_f32(const float* g_cube, const float* g_cabins, const int32_t* idx,
                      const uint16_t* cab_arg, int B, int C, int N, int R, int k, int cab,
                      float* grad_x, void* stream);

/* Standalone train2cabins (softpool.py:71-85) on any (rows, k) window tensor, rows = B*C*R.
 *   cabins (rows,cab) f32, cab_arg (rows,cab) u16 as above.                                 */
int sp_cabins_fwd_f32(const float* windows, long long rows, int k, int cab, float* cabins,
                      uint16_t* cab_arg, void* stream);
/* Its backward (MaxBackward + CopySlices): g_windows (rows,k) fully overwritten.            */
int sp_cabins_bwd_f32(const float* g_cabins, const uint16_t* cab_arg, long long rows, int k,
                      int cab, float* g_windows, void* stream);

/* ------------------------------------------------------------------------------------
 * Chamfer distance
 * ---------------------------------------------------------------------------------- */

/* Bytes of scratch chamfer_fwd_f32 needs for these sizes (0 is possible).               */
size_t chamfer_fwd_workspace_bytes(int B, int n, int m);

/* Both directions of the squared nearest-neighbour distance.  Replaces
 * `chamfer_cuda_forward` = 2 x NmDistanceKernel (distance/chamfer/chamfer.cu:12-152; same
 * kernel in GRNet/extensions/chamfer_dist/chamfer.cu:15-171).
 *   xyz1 (B,n,3) f32, xyz2 (B,m,3) f32
 *   dist1 (B,n) f32 = min_k |xyz1[b,j]-xyz2[b,k]|^2, idx1 (B,n) i32 = arg min (FIRST minimum);
 *   dist2 (B,m), idx2 (B,m): the mirror.  The distance is evaluated with the reference's
 *   float32 expression d = fma(dz,dz, fma(dx,dx, dy*dy)) (its SASS under nvcc's default
 *   -fmad=true), so dist/idx are bit-identical to the reference kernel for finite inputs.
 *   n == 0 or m == 0 leaves zeros in the outputs, like the reference's zero-initialised
 *   buffers (dist_chamfer.py:19-23).  Samples with a NaN / inf coordinate are evaluated in the
 *   reference's own loop order (targets in batches of 512, `k==0 || d<best` inside a batch,
 *   `k2==0 || result>best` across batches), so they match the reference kernel too.
 *   ws / ws_bytes: scratch of at least chamfer_fwd_workspace_bytes(B,n,m), 16-byte aligned. */
int chamfer_fwd_f32(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1,
                    float* dist2, int32_t* idx1, int32_t* idx2, void* ws, size_t ws_bytes,
                    void* stream);

/* chamfer_fwd_f32 with the loss reduction of every call site of the reference folded into the same
 * launch (train.py:68-69,82-86: `torch.mean(dist1,1) + torch.mean(dist2,1)`; val.py:302-303):
 *   loss (B) f32 = mean_j dist1[b,j] + mean_k dist2[b,k].  Per-warp partial sums are added with float
 *   reductions, so the last bits of the loss depend on arrival order (dist/idx stay bit-exact).
 *   Requires n, m >= 1.  Everything else as chamfer_fwd_f32.                                         */
int chamfer_fwd_loss_f32(const float* xyz1, const float* xyz2, int B, int n, int m, float* dist1,
                         float* dist2, int32_t* idx1, int32_t* idx2, float* loss, void* ws,
                         size_t ws_bytes, void* stream);

/* P predictions against ONE ground truth in one call -- the training step's `self.CD(output1[i], gt)` for i = 1..3 and
 * the `output2` pair (train.py:68-86), which the reference runs as P separate forward calls.
 *   xyz1 (P*B, n, 3): the predictions stacked along the batch (prediction p, sample b at row p*B + b);
 *   xyz2 (B, m, 3): the shared ground truth.  Outputs as chamfer_fwd_loss_f32 with batch P*B:
 *   dist1/idx1 (P*B, n), dist2/idx2 (P*B, m), loss (P*B) or NULL.  Results are bit-identical to P separate calls.
 *   On the dense tensor path this is ONE prep launch + ONE tensor launch, and the ground truth's operand rows are
 *   formatted once per sample (all P predictions of a sample share one frame); other paths loop over p internally.
 *   ws: chamfer_fwd_workspace_bytes(P*B, n, m).                                                                  */
int chamfer_fwd_multi_f32(const float* xyz1, const float* xyz2, int P, int B, int n, int m, float* dist1,
                          float* dist2, int32_t* idx1, int32_t* idx2, float* loss, void* ws,
                          size_t ws_bytes, void* stream);

/* Replaces `chamfer_cuda_backward` = 2 x NmDistanceGradKernel (chamfer.cu:155-196); one launch
 * (a thread-block cluster per sample: direct terms, cluster barrier, scatter terms).
 *   grad_xyz1 (B,n,3), grad_xyz2 (B,m,3): fully overwritten (the reference accumulates into
 *   caller-zeroed buffers with atomicAdd):
 *     grad_xyz1[j]      = 2 g1[j] (p_j - q_idx1[j])  -  sum_{k: idx2[k]==j} 2 g2[k] (q_k - p_j)
 *     grad_xyz2[k]      = 2 g2[k] (q_k - p_idx2[k])  -  sum_{j: idx1[j]==k} 2 g1[j] (p_j - q_k) */
int chamfer_bwd_f32(const float* xyz1, const float* xyz2, const float* g1, const float* g2,
                    const int32_t* idx1, const int32_t* idx2, int B, int n, int m,
                    float* grad_xyz1, float* grad_xyz2, void* stream);

/* Fused loss epilogue used at every call site of the reference
 * (train.py:68-69,82-86: `torch.mean(dist1,1) + torch.mean(dist2,1)`):
 *   loss (B) f32 = mean_j dist1[b,j] + mean_k dist2[b,k]                                 */
int chamfer_loss_f32(const float* dist1, const float* dist2, int B, int n, int m, float* loss,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SOFTPOOL_B200_H_ */
