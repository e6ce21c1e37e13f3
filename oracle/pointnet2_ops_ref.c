/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
 *
 * Plain-C CPU restatement of the nine `pointnet2_ops._ext` entry points that the
 * reference calls at pointnet2/utils/pointnet2_utils.py:56 (furthest_point_sampling),
 * :92/:98 (gather_points / _grad), :125 (three_nn), :162/:184 (three_interpolate / _grad),
 * :217/:237 (group_points / _grad) and :268 (ball_query).
 *
 * The arithmetic itself lives in the third-party dependency `pointnet2_ops`
 * (erikwijmans/Pointnet2_PyTorch, sub-directory pointnet2_ops_lib, version string 3.0.0,
 * un-pinned in requirement.txt:5 and NOT vendored under /root/reference).  What follows
 * restates that package's published CUDA algorithm thread-for-thread (block size rule,
 * strided scans, shared-memory tree reduction, strict comparisons, first-hit padding).
 *
 * PARITY UNPINNED: neither the reference nor that dependency ships golden vectors or tests
 * for these ops, and the dependency's binary cannot be built here (no source, no network).
 * The known-answer tests in tests/test_oracle_ops.py pin the *semantics* listed in
 * SURVEY.md §2.3; the composition above these ops is pinned against the reference's own
 * Python (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Floating point: upstream is compiled by nvcc with its default --fmad=true, so
 *   (a-b)*(a-b) + (c-d)*(c-d) + (e-f)*(e-f)
 * contracts to  fma(dz,dz, fma(dy,dy, dx*dx)).  `fma_mode`=1 (default everywhere)
 * reproduces that; `fma_mode`=0 is the un-contracted left-to-right evaluation.
 * This file must be compiled with -ffp-contract=off so only the explicit fmaf() fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define O3D_REF_TOTAL_THREADS 512

/* cuda_utils.h: opt_n_threads(work) = clamp(2^floor(log2(work)), 1, TOTAL_THREADS) */
int o3d_ref_opt_n_threads(int work_size) {
    if (work_size < 1) return 1;
    int pow2 = 0;
    while ((1 << (pow2 + 1)) <= work_size) ++pow2;
    int t = 1 << pow2;
    if (t > O3D_REF_TOTAL_THREADS) t = O3D_REF_TOTAL_THREADS;
    if (t < 1) t = 1;
    return t;
}

static inline float sq3(float dx, float dy, float dz, int fma_mode) {
    if (fma_mode) return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float a = dx * dx, b = dy * dy, c = dz * dz;
    return (a + b) + c;
}

/* ------------------------------------------------------------------ FPS ---- */
/* furthest_point_sampling_kernel<block>: one block per cloud, thread t scans
 * k = t, t+block, ... with strict '>' ; then a shared-memory tree (stride
 * block/2 ... 1) where the left operand survives ties.  temp[] starts at 1e10.  */

/* The `mag <= 1e-3` test upstream compares a float against a double literal.  (double)mag <= 1e-3
 * differs from mag <= 1e-3f only if mag lies strictly between 1e-3 (double) and (float)1e-3;
 * (float)1e-3 = 0.001000000047..., so a float mag equal to (float)1e-3 is > 1e-3 (double) and is
 * NOT skipped upstream.  Handle that exactly:                                                   */
static inline int skip_origin(float mag) { return (double)mag <= 1e-3; }

void o3d_ref_fps(const float *xyz, int B, int N, int npoint, int32_t *idx, int fma_mode) {
    if (npoint <= 0) return;
    const int block = o3d_ref_opt_n_threads(N);
    float *temp = (float *)malloc(sizeof(float) * (size_t)N);
    float *dists = (float *)malloc(sizeof(float) * (size_t)block);
    int *dists_i = (int *)malloc(sizeof(int) * (size_t)block);
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        int32_t *out = idx + (size_t)b * npoint;
        for (int k = 0; k < N; ++k) temp[k] = 1e10f;
        int old = 0;
        out[0] = 0;
        for (int j = 1; j < npoint; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < block; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < N; k += block) {
                    const float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
                    if (skip_origin(sq3(x2, y2, z2, fma_mode))) continue;
                    const float d = sq3(x2 - x1, y2 - y1, z2 - z1, fma_mode);
                    const float d2 = fminf(d, temp[k]);
                    temp[k] = d2;
                    if (d2 > best) { besti = k; best = d2; }
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int s = block / 2; s >= 1; s >>= 1)
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2;
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(temp); free(dists); free(dists_i);
}

/* ----------------------------------------------------------- ball query ---- */
/* query_ball_point_kernel: ascending-k scan, strict d2 < r*r (r*r in float),
 * first hit fills all nsample slots, stop at nsample hits, no hit -> zeros.   */
void o3d_ref_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M,
                        float radius, int nsample, int32_t *idx, int fma_mode) {
    const float radius2 = radius * radius;
    memset(idx, 0, sizeof(int32_t) * (size_t)B * M * nsample);
    for (int b = 0; b < B; ++b) {
        const float *p = xyz + (size_t)b * N * 3;
        const float *q = new_xyz + (size_t)b * M * 3;
        int32_t *o = idx + (size_t)b * M * nsample;
        for (int j = 0; j < M; ++j) {
            const float nx = q[j * 3 + 0], ny = q[j * 3 + 1], nz = q[j * 3 + 2];
            int cnt = 0;
            for (int k = 0; k < N && cnt < nsample; ++k) {
                const float d2 = sq3(nx - p[k * 3 + 0], ny - p[k * 3 + 1], nz - p[k * 3 + 2], fma_mode);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[j * nsample + l] = k;
                    o[j * nsample + cnt] = k;
                    ++cnt;
                }
            }
        }
    }
}

/* ------------------------------------------------------- gather / group ---- */
void o3d_ref_gather(const float *feat, const int32_t *idx, int B, int C, int N, int M, float *out) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                out[((size_t)b * C + c) * M + j] = feat[((size_t)b * C + c) * N + idx[(size_t)b * M + j]];
}

void o3d_ref_gather_grad(const float *grad_out, const int32_t *idx, int B, int C, int N, int M,
                         float *grad_feat /* zero-filled by caller */) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                grad_feat[((size_t)b * C + c) * N + idx[(size_t)b * M + j]] +=
                    grad_out[((size_t)b * C + c) * M + j];
}

void o3d_ref_group(const float *feat, const int32_t *idx, int B, int C, int N, int M, int S, float *out) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                for (int k = 0; k < S; ++k)
                    out[(((size_t)b * C + c) * M + j) * S + k] =
                        feat[((size_t)b * C + c) * N + idx[((size_t)b * M + j) * S + k]];
}

void o3d_ref_group_grad(const float *grad_out, const int32_t *idx, int B, int C, int N, int M, int S,
                        float *grad_feat /* zero-filled */) {
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c)
            for (int j = 0; j < M; ++j)
                for (int k = 0; k < S; ++k)
                    grad_feat[((size_t)b * C + c) * N + idx[((size_t)b * M + j) * S + k]] +=
                        grad_out[(((size_t)b * C + c) * M + j) * S + k];
}

/* ------------------------------------------------------------ three_nn ---- */
/* three_nn_kernel: bests kept in double, initialised 1e40, strict '<' insertion
 * (ties keep the lower index first); outputs squared distances cast to float.  */
void o3d_ref_three_nn(const float *unknown, const float *known, int B, int n, int m,
                      float *dist2, int32_t *idx, int fma_mode) {
    for (int b = 0; b < B; ++b) {
        const float *u = unknown + (size_t)b * n * 3;
        const float *kn = known + (size_t)b * m * 3;
        for (int j = 0; j < n; ++j) {
            const float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int b1 = 0, b2 = 0, b3 = 0;
            for (int k = 0; k < m; ++k) {
                const float d = sq3(ux - kn[k * 3 + 0], uy - kn[k * 3 + 1], uz - kn[k * 3 + 2], fma_mode);
                if (d < best1) {
                    best3 = best2; b3 = b2; best2 = best1; b2 = b1; best1 = d; b1 = k;
                } else if (d < best2) {
                    best3 = best2; b3 = b2; best2 = d; b2 = k;
                } else if (d < best3) {
                    best3 = d; b3 = k;
                }
            }
            float *od = dist2 + ((size_t)b * n + j) * 3;
            int32_t *oi = idx + ((size_t)b * n + j) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = b1; oi[1] = b2; oi[2] = b3;
        }
    }
}

/* ----------------------------------------------------- three_interpolate ---- */
void o3d_ref_three_interpolate(const float *feat, const int32_t *idx, const float *w,
                               int B, int c, int m, int n, float *out, int fma_mode) {
    for (int b = 0; b < B; ++b)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < n; ++j) {
                const int32_t *i3 = idx + ((size_t)b * n + j) * 3;
                const float *w3 = w + ((size_t)b * n + j) * 3;
                const float *f = feat + ((size_t)b * c + l) * m;
                float r;
                if (fma_mode) r = fmaf(f[i3[2]], w3[2], fmaf(f[i3[1]], w3[1], f[i3[0]] * w3[0]));
                else r = (f[i3[0]] * w3[0] + f[i3[1]] * w3[1]) + f[i3[2]] * w3[2];
                out[((size_t)b * c + l) * n + j] = r;
            }
}

void o3d_ref_three_interpolate_grad(const float *grad_out, const int32_t *idx, const float *w,
                                    int B, int c, int n, int m, float *grad_feat /* zero-filled */) {
    for (int b = 0; b < B; ++b)
        for (int l = 0; l < c; ++l)
            for (int j = 0; j < n; ++j) {
                const int32_t *i3 = idx + ((size_t)b * n + j) * 3;
                const float *w3 = w + ((size_t)b * n + j) * 3;
                const float g = grad_out[((size_t)b * c + l) * n + j];
                float *gf = grad_feat + ((size_t)b * c + l) * m;
                gf[i3[0]] += g * w3[0];
                gf[i3[1]] += g * w3[1];
                gf[i3[2]] += g * w3[2];
            }
}
