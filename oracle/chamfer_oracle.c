/* chamfer_oracle.c -- CPU restatement of the reference Chamfer kernels.  TEST INFRASTRUCTURE ONLY
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference); never linked or
 * loaded by softpool_b200/.
 *
 * Restates, statement for statement in its arithmetic and tie rules:
 *   NmDistanceKernel      /root/reference/distance/chamfer/chamfer.cu:12-134
 *   NmDistanceGradKernel  /root/reference/distance/chamfer/chamfer.cu:155-174
 * (the GRNet copy, GRNet/extensions/chamfer_dist/chamfer.cu:15-201, is the same code).
 *
 *  - targets are visited in chunks of 512 (chamfer.cu:12,16); inside a chunk the first element
 *    initialises the running best (`k==0 || d<best`, :36) and later ones replace it only when
 *    strictly smaller; across chunks the stored result is replaced only when strictly greater
 *    (`k2==0 || result>best`, :126)  ->  FIRST minimum wins; NaN behaviour follows from the same
 *    comparisons;
 *  - the squared distance is float32 with the contraction nvcc's default -fmad=true produces for
 *    `x2*x2+y2*y2+z2*z2` (checked in the SASS of oracle/_ref: FMUL y*y, FFMA x*x+., FFMA z*z+.):
 *        d = fmaf(dz, dz, fmaf(dx, dx, dy*dy)),  dx = x2 - x1 (target minus query, :31-33);
 *  - backward (:155-174): g = grad_dist*2; grad_xyz1[j] += g*(p-q); grad_xyz2[idx] += -(g*(p-q)),
 *    accumulated here sequentially (the reference uses atomicAdd, order unspecified).
 *
 * PINNED on the GPU box against oracle/_ref (the reference .cu compiled unmodified for sm_100):
 * tests/test_chamfer_gpu.py::test_c_oracle_matches_reference_kernel, and against the committed
 * outputs of that kernel in tests/golden/chamfer_ref_*.npz (CPU test).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -pthread; this image's gcc has no OpenMP)
 * Threads: work items (sample x direction) are handed out over `nthreads` pthreads.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#define CHUNK 512

static inline float sqdist(float x1, float y1, float z1, float x2, float y2, float z2) {
    const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* one direction: for every query j of xyz (n points) the nearest of xyz2 (m points) */
static void nm_distance(int n, const float* xyz, int m, const float* xyz2, float* result, int32_t* result_i) {
    for (int k2 = 0; k2 < m; k2 += CHUNK) {
        const int end_k = (m < k2 + CHUNK ? m : k2 + CHUNK) - k2;
        const float* buf = xyz2 + (size_t)k2 * 3;
        for (int j = 0; j < n; ++j) {
            const float x1 = xyz[j * 3 + 0], y1 = xyz[j * 3 + 1], z1 = xyz[j * 3 + 2];
            int best_i = 0;
            float best = 0.f;
            for (int k = 0; k < end_k; ++k) {
                const float d = sqdist(x1, y1, z1, buf[k * 3 + 0], buf[k * 3 + 1], buf[k * 3 + 2]);
                if (k == 0 || d < best) { best = d; best_i = k + k2; }
            }
            if (k2 == 0 || result[j] > best) { result[j] = best; result_i[j] = best_i; }
        }
    }
}

typedef struct {
    int B, n, m, mode;                       /* mode 0 = forward, 1 = backward */
    const float *xyz1, *xyz2, *g1, *g2;
    float *dist1, *dist2, *gx1, *gx2;
    int32_t *idx1, *idx2;
    int next;                                /* next work item, taken with an atomic add */
} job_t;

static void nm_grad(int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                    const int32_t* idx1, float* grad_xyz1, float* grad_xyz2);

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    const int n = J->n, m = J->m;
    const int items = J->mode == 0 ? 2 * J->B : J->B;
    for (;;) {
        const int w = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
        if (w >= items) break;
        if (J->mode == 0) {
            const int b = w >> 1;
            if ((w & 1) == 0)
                nm_distance(n, J->xyz1 + (size_t)b * n * 3, m, J->xyz2 + (size_t)b * m * 3, J->dist1 + (size_t)b * n, J->idx1 + (size_t)b * n);
            else
                nm_distance(m, J->xyz2 + (size_t)b * m * 3, n, J->xyz1 + (size_t)b * n * 3, J->dist2 + (size_t)b * m, J->idx2 + (size_t)b * m);
        } else {
            const int b = w;
            const float* p = J->xyz1 + (size_t)b * n * 3;
            const float* q = J->xyz2 + (size_t)b * m * 3;
            float* gp = J->gx1 + (size_t)b * n * 3;
            float* gq = J->gx2 + (size_t)b * m * 3;
            nm_grad(n, p, m, q, J->g1 + (size_t)b * n, J->idx1 + (size_t)b * n, gp, gq);     /* chamfer.cu:184 */
            nm_grad(m, q, n, p, J->g2 + (size_t)b * m, J->idx2 + (size_t)b * m, gq, gp);     /* chamfer.cu:185 */
        }
    }
    return 0;
}

static void run(job_t* J, int nthreads) {
    pthread_t th[256];
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    J->next = 0;
    for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], 0, worker, J);
    worker(J);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], 0);
}

/* dist1/idx1 (B,n), dist2/idx2 (B,m); outputs must arrive zeroed when n or m is 0 (dist_chamfer.py:19-23) */
void chamfer_oracle_forward(int B, int n, int m, const float* xyz1, const float* xyz2, float* dist1,
                            float* dist2, int32_t* idx1, int32_t* idx2, int nthreads) {
    job_t J;
    memset(&J, 0, sizeof(J));
    J.B = B; J.n = n; J.m = m; J.mode = 0; J.xyz1 = xyz1; J.xyz2 = xyz2;
    J.dist1 = dist1; J.dist2 = dist2; J.idx1 = idx1; J.idx2 = idx2;
    run(&J, nthreads);
}

static void nm_grad(int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                    const int32_t* idx1, float* grad_xyz1, float* grad_xyz2) {
    (void)m;
    for (int j = 0; j < n; ++j) {
        const float x1 = xyz1[j * 3 + 0], y1 = xyz1[j * 3 + 1], z1 = xyz1[j * 3 + 2];
        const int j2 = idx1[j];
        const float x2 = xyz2[j2 * 3 + 0], y2 = xyz2[j2 * 3 + 1], z2 = xyz2[j2 * 3 + 2];
        const float g = grad_dist1[j] * 2;
        grad_xyz1[j * 3 + 0] += g * (x1 - x2);
        grad_xyz1[j * 3 + 1] += g * (y1 - y2);
        grad_xyz1[j * 3 + 2] += g * (z1 - z2);
        grad_xyz2[j2 * 3 + 0] += -(g * (x1 - x2));
        grad_xyz2[j2 * 3 + 1] += -(g * (y1 - y2));
        grad_xyz2[j2 * 3 + 2] += -(g * (z1 - z2));
    }
}

/* grad_xyz1 (B,n,3), grad_xyz2 (B,m,3) are zeroed here (dist_chamfer.py:40-41) */
void chamfer_oracle_backward(int B, int n, int m, const float* xyz1, const float* xyz2, const float* g1,
                             const float* g2, const int32_t* idx1, const int32_t* idx2, float* grad_xyz1,
                             float* grad_xyz2, int nthreads) {
    job_t J;
    memset(grad_xyz1, 0, (size_t)B * n * 3 * sizeof(float));
    memset(grad_xyz2, 0, (size_t)B * m * 3 * sizeof(float));
    memset(&J, 0, sizeof(J));
    J.B = B; J.n = n; J.m = m; J.mode = 1; J.xyz1 = xyz1; J.xyz2 = xyz2; J.g1 = g1; J.g2 = g2;
    J.idx1 = (int32_t*)idx1; J.idx2 = (int32_t*)idx2; J.gx1 = grad_xyz1; J.gx2 = grad_xyz2;
    run(&J, nthreads);
}
