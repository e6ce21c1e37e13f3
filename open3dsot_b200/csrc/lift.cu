// Lifted first layer (include/o3d_b200.h: o3d_lift_t) — the two gather / scatter passes around the GEMMs.
//
// The first 1x1 convolution of a grouped stack (QueryAndGroup -> SharedMLP, pointnet2_utils.py:317-329 +
// pointnet2_modules.py:64-69; BoxAwareXCorr, xcorr.py:87-98; P2B_XCorr, xcorr.py:39-47) is linear in the grouped row, so
//     Y0[p] = W0 . [x(idx_p) - centre, f(idx_p)] = Z[idx_p] - cc[centre(p)],   Z = W0 . [x, f] per SOURCE point
// and the (B, 3 + C, npoint, nsample) tensor, the layer's GEMM over it and their gradients never exist.
//
//   lift_stats_kernel    one pass over the positions: global row index gidx[p], per-channel sum / sum of squares of Y0 for the
//                        BatchNorm that follows (train mode), optionally Y0 itself (small problems that stay on the CUDA-core
//                        GEMMs read it as an ordinary activation matrix)
//   lift_scatter_kernel  backward: dY0 = a*g + b + c*Y0 (BatchNorm backward, Y0 re-gathered) accumulated into dZ[gidx[p]] (vector
//                        REDs), -sum over the group into dcc (plain stores: one thread owns a group), ds[p] = dY0 . u,
//                        du += s[p] * dY0
// Thread layout of both: a row of C0 channels = C0/4 threads (one float4 each); a block holds `rl` row lanes; a lane walks
// whole groups of `grp` consecutive positions, so the group reduction for dcc stays in registers.  The source rows (Z) are a
// few MB and L2-resident; the traffic to DRAM is gidx, g and the REDs.
#include "common.cuh"
#include "lift.cuh"
#include "../../include/o3d_b200.h"

namespace {

constexpr int LIFT_THREADS = 256;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct LiftGeom { const int32_t* ridx; int ridx_mod, rows_per_cloud, pos_per_cloud; };

__device__ __forceinline__ int lift_row(const LiftGeom& g, int p) {
    const int cloud = p / g.pos_per_cloud;
    const int local = g.ridx ? __ldg(g.ridx + p) : (p % g.ridx_mod);
    return cloud * g.rows_per_cloud + local;
}

template <bool STORE>
__global__ void __launch_bounds__(LIFT_THREADS)
    lift_stats_kernel(LiftView lv, LiftGeom geo, int P, int C0, int grp, int32_t* __restrict__ gidx, float* __restrict__ y0,
                      double* __restrict__ sum, double* __restrict__ sumsq) {
    extern __shared__ double red[];                 // [2][rl][C0]
    const int tpr = C0 >> 2, rl = blockDim.x / tpr;
    const int t = threadIdx.x % tpr, lane = threadIdx.x / tpr;
    const int c = t * 4;
    const int n_groups = P / grp;
    const float4 u4 = lv.u ? ldg4(lv.u + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    double d1[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
    {
        for (int g = blockIdx.x * rl + lane; g < n_groups; g += gridDim.x * rl) {
            const float4 c4 = lv.cc ? ldg4(lv.cc + (size_t)g * lv.ldz + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
            const int pb = g * grp;
#pragma unroll 4
            for (int s = 0; s < grp; ++s) {
                const int p = pb + s;
                const int row = lift_row(geo, p);
                if (t == 0) gidx[p] = row;
                const float4 z4 = ldg4(lv.z + (size_t)row * lv.ldz + c);
                const float sv = lv.s ? __ldg(lv.s + p) : 0.f;
                const float4 v = lift_val4(z4, c4, sv, u4);
                if (STORE) *reinterpret_cast<float4*>(y0 + (size_t)p * lv.ldz + c) = v;
                a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
                a2.x = fmaf(v.x, v.x, a2.x); a2.y = fmaf(v.y, v.y, a2.y); a2.z = fmaf(v.z, v.z, a2.z); a2.w = fmaf(v.w, v.w, a2.w);
                if ((s & 31) == 31 || s == grp - 1) {      // fp32 inside 32 positions, fp64 across (same rule as the GEMM epilogues)
                    d1[0] += a1.x; d1[1] += a1.y; d1[2] += a1.z; d1[3] += a1.w;
                    d2[0] += a2.x; d2[1] += a2.y; d2[2] += a2.z; d2[3] += a2.w;
                    a1 = a2 = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    if (!sum) return;
    {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[(size_t)lane * C0 + c + j] = d1[j];
            red[(size_t)(rl + lane) * C0 + c + j] = d2[j];
        }
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C0; ch += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int l = 0; l < rl; ++l) { s1 += red[(size_t)l * C0 + ch]; s2 += red[(size_t)(rl + l) * C0 + ch]; }
        atomicAdd(sum + ch, s1);
        atomicAdd(sumsq + ch, s2);
    }
}

__global__ void __launch_bounds__(LIFT_THREADS)
    lift_scatter_kernel(LiftView lv, int P, int C0, int grp, const float* __restrict__ y0, const float* __restrict__ g, int ldg,
                        const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ ccf,
                        float* __restrict__ dz, float* __restrict__ dcc, float* __restrict__ ds, float* __restrict__ du) {
    extern __shared__ float redf[];                 // [rl][C0] (du only)
    const int tpr = C0 >> 2, rl = blockDim.x / tpr;
    const int t = threadIdx.x % tpr, lane = threadIdx.x / tpr;
    const int c = t * 4;
    const int n_groups = P / grp;
    const float4 u4 = lv.u ? ldg4(lv.u + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a4 = ca ? ldg4(ca + c) : one, b4 = ca ? ldg4(cb + c) : zero, k4 = ca ? ldg4(ccf + c) : zero;
    float4 du4 = zero;
    // shuffle width for the per-row dot product ds[p] = dY0[p] . u: the row's threads inside one warp (only when a warp
    // holds whole rows or a row holds whole warps; otherwise one RED per thread)
    const int w = tpr >= 32 ? 32 : tpr;
    const bool shfl = (tpr <= 32 && (tpr & (tpr - 1)) == 0) || (tpr % 32 == 0);
    {
        // uniform trip count per block (the shuffles below need every lane of a warp): lanes past the last group idle
        for (int gb = blockIdx.x * rl; gb < n_groups; gb += gridDim.x * rl) {
            const bool valid = gb + lane < n_groups;
            const int gi = valid ? gb + lane : n_groups - 1;
            const float4 c4 = lv.cc ? ldg4(lv.cc + (size_t)gi * lv.ldz + c) : zero;
            float4 acc = zero, acc0 = zero;
            const int pb = gi * grp;
            // ball-query padding repeats a group's FIRST neighbour in every unused slot (pointnet2_ops ball_query), so a
            // sparse ball sends most of its positions to one row: those are summed in registers and leave as ONE RED
            const int row0 = __ldg(lv.gidx + pb);
#pragma unroll 4
            for (int s = 0; s < grp; ++s) {
                const int p = pb + s;
                const int row = __ldg(lv.gidx + p);
                const float4 g4 = ldg4(g + (size_t)p * ldg + c);
                const float sv = lv.s ? __ldg(lv.s + p) : 0.f;
                float4 dy = g4;
                if (ca) {
                    float4 v;
                    if (y0) v = ldg4(y0 + (size_t)p * lv.ldz + c);
                    else v = lift_val4(ldg4(lv.z + (size_t)row * lv.ldz + c), c4, sv, u4);
                    dy.x = fmaf(a4.x, g4.x, fmaf(k4.x, v.x, b4.x)); dy.y = fmaf(a4.y, g4.y, fmaf(k4.y, v.y, b4.y));
                    dy.z = fmaf(a4.z, g4.z, fmaf(k4.z, v.z, b4.z)); dy.w = fmaf(a4.w, g4.w, fmaf(k4.w, v.w, b4.w));
                }
                if (row == row0) { acc0.x += dy.x; acc0.y += dy.y; acc0.z += dy.z; acc0.w += dy.w; }
                else if (dz && valid) atomicAdd(reinterpret_cast<float4*>(dz + (size_t)row * lv.ldz + c), dy);   // sm_90+: one vector RED
                acc.x += dy.x; acc.y += dy.y; acc.z += dy.z; acc.w += dy.w;
                if (du && valid) {
                    du4.x = fmaf(sv, dy.x, du4.x); du4.y = fmaf(sv, dy.y, du4.y);
                    du4.z = fmaf(sv, dy.z, du4.z); du4.w = fmaf(sv, dy.w, du4.w);
                }
                if (ds) {
                    float d = fmaf(dy.x, u4.x, fmaf(dy.y, u4.y, fmaf(dy.z, u4.z, dy.w * u4.w)));
                    if (shfl) {
                        for (int o = w >> 1; o >= 1; o >>= 1) d += __shfl_xor_sync(0xFFFFFFFFu, d, o, 32);
                        if ((t & (w - 1)) == 0 && valid) atomicAdd(ds + p, d);
                    } else if (valid) {
                        atomicAdd(ds + p, d);
                    }
                }
            }
            if (dz && valid) atomicAdd(reinterpret_cast<float4*>(dz + (size_t)row0 * lv.ldz + c), acc0);
            if (dcc && valid) *reinterpret_cast<float4*>(dcc + (size_t)gi * lv.ldz + c) = make_float4(-acc.x, -acc.y, -acc.z, -acc.w);
        }
    }
    if (!du) return;
    *reinterpret_cast<float4*>(redf + (size_t)lane * C0 + c) = du4;
    __syncthreads();
    for (int ch = threadIdx.x; ch < C0; ch += blockDim.x) {
        float s1 = 0.f;
        for (int l = 0; l < rl; ++l) s1 += redf[(size_t)l * C0 + ch];
        atomicAdd(du + ch, s1);
    }
}

inline int ilog2_floor(int v) { int l = 0; while ((2 << l) <= v) ++l; return l; }

int lift_check(const o3d_lift_t* lf, int P, int C0, const char* who) {
    O3D_REQUIRE(lf && lf->z, O3D_ERR_ARG, "%s: null lift descriptor", who);
    O3D_REQUIRE(C0 >= 4 && (C0 & 3) == 0 && C0 <= 1024 && lf->ldz == C0, O3D_ERR_ARG, "%s: C0=%d ldz=%d (need C0 %% 4 == 0, ldz == C0)",
                who, C0, lf->ldz);
    O3D_REQUIRE(lf->grp >= 1 && (lf->grp & (lf->grp - 1)) == 0 && P % lf->grp == 0, O3D_ERR_ARG,
                "%s: grp=%d must be a power of two dividing P=%d", who, lf->grp, P);
    O3D_REQUIRE(lf->pos_per_cloud >= 1 && lf->rows_per_cloud >= 1 && (lf->ridx || lf->ridx_mod >= 1), O3D_ERR_ARG,
                "%s: bad cloud geometry", who);
    O3D_REQUIRE((lf->s == nullptr) == (lf->u == nullptr), O3D_ERR_ARG, "%s: s and u come together", who);
    return O3D_OK;
}

inline LiftView make_view(const o3d_lift_t* lf, const int32_t* gidx) {
    return LiftView{lf->z, lf->ldz, gidx, lf->cc, ilog2_floor(lf->grp), lf->s, lf->u};
}

inline void lift_launch_shape(int P, int C0, int grp, int& threads, int& blocks, int& rl) {
    const int tpr = C0 / 4;
    rl = LIFT_THREADS / tpr;
    if (rl < 1) rl = 1;
    threads = tpr * rl;
    const int n_groups = P / grp;
    blocks = (n_groups + rl - 1) / rl;
    const int cap = 8 * o3d_num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
}

}  // namespace

extern "C" int o3d_lift_stats(const o3d_lift_t* lf, int P, int C0, int32_t* gidx, float* y0, double* sum, double* sumsq,
                              void* stream) {
    if (int e = lift_check(lf, P, C0, "o3d_lift_stats")) return e;
    O3D_REQUIRE(gidx, O3D_ERR_ARG, "o3d_lift_stats: gidx workspace missing");
    if (P == 0) return O3D_OK;
    int threads, blocks, rl;
    lift_launch_shape(P, C0, lf->grp, threads, blocks, rl);
    const LiftView lv = make_view(lf, gidx);
    const LiftGeom geo{lf->ridx, lf->ridx_mod, lf->rows_per_cloud, lf->pos_per_cloud};
    const size_t smem = sum ? sizeof(double) * 2 * rl * C0 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (y0) lift_stats_kernel<true><<<blocks, threads, smem, st>>>(lv, geo, P, C0, lf->grp, gidx, y0, sum, sumsq);
    else lift_stats_kernel<false><<<blocks, threads, smem, st>>>(lv, geo, P, C0, lf->grp, gidx, y0, sum, sumsq);
    O3D_CHECK_LAUNCH("o3d_lift_stats");
    return O3D_OK;
}

extern "C" int o3d_lift_scatter(const o3d_lift_t* lf, int P, int C0, const int32_t* gidx, const float* y0, const float* g,
                                int ldg, const float* a, const float* b, const float* cc, void* stream) {
    if (int e = lift_check(lf, P, C0, "o3d_lift_scatter")) return e;
    O3D_REQUIRE(gidx && g && (ldg & 3) == 0, O3D_ERR_ARG, "o3d_lift_scatter: null pointer / ldg");
    O3D_REQUIRE(!lf->d_s || lf->u, O3D_ERR_ARG, "o3d_lift_scatter: d_s without u");
    if (P == 0) return O3D_OK;
    int threads, blocks, rl;
    lift_launch_shape(P, C0, lf->grp, threads, blocks, rl);
    const LiftView lv = make_view(lf, gidx);
    const size_t smem = lf->d_u ? sizeof(float) * rl * C0 : 0;
    lift_scatter_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>(lv, P, C0, lf->grp, y0, g, ldg, a, b, cc, lf->d_z,
                                                                        lf->cc ? lf->d_cc : nullptr, lf->s ? lf->d_s : nullptr,
                                                                        lf->s ? lf->d_u : nullptr);
    O3D_CHECK_LAUNCH("o3d_lift_scatter");
    return O3D_OK;
}
