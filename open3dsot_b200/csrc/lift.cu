// Lifted first layer (include/o3d_b200.h: o3d_lift_t) — the two gather / scatter passes around the GEMMs.
//
// The first 1x1 convolution of a grouped stack (QueryAndGroup -> SharedMLP, pointnet2_utils.py:317-329 +
// pointnet2_modules.py:64-69; BoxAwareXCorr, xcorr.py:87-98; P2B_XCorr, xcorr.py:39-47) is linear in the grouped row
// [x(idx_p) - centre, f(idx_p)] (or [sim, xyz_i, f_i]), so
//     Y0[p] = W0_f . f(idx_p) + W0_x . (x(idx_p) - centre)  =  Z[idx_p] + sum_j s[p][j] * u[j]
// with Z = W0_f . f computed ONCE per source point (an ordinary one-layer stack) and the few per-position scalars s
// (relative coordinates dx dy dz, or the cosine similarity) applied directly — in the same difference-then-multiply form
// as the reference, so no precision is lost to cancellation.  The (B, 3 + C, npoint, nsample) tensor, the layer's GEMM over
// it and their gradients never exist.
//
//   lift_stats_kernel    one pass over the positions: global row index gidx[p], per-channel sum / sum of squares of Y0 for the
//                        BatchNorm that follows (train mode), optionally Y0 itself (small problems that stay on the CUDA-core
//                        GEMMs read it as an ordinary activation matrix)
//   lift_scatter_kernel  backward: dY0 = a*g + b + c*Y0 (BatchNorm backward, Y0 re-evaluated) accumulated into dZ[gidx[p]]
//                        (vector REDs; a group's first-hit padding duplicates are summed in registers first),
//                        ds[p][j] = dY0[p] . u[j], du[j] += s[p][j] * dY0[p]
// Thread layout of both: a row of C0 channels = C0/4 threads (one float4 each); a block holds `rl` row lanes; a lane walks
// whole groups of `grp` consecutive positions.  The source rows (Z) are a few MB and L2-resident; the traffic to DRAM is
// gidx, s, g and the REDs.
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

struct U4 { float4 u0, u1, u2, u3; };
__device__ __forceinline__ U4 load_u(const LiftView& lv, int c) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!lv.u) return U4{z, z, z, z};
    return U4{ldg4(lv.u + c), ldg4(lv.u + lv.ldz + c), ldg4(lv.u + 2 * lv.ldz + c), ldg4(lv.u + 3 * lv.ldz + c)};
}

template <bool STORE>
__global__ void __launch_bounds__(LIFT_THREADS, 4)      // <= 64 registers: 4 blocks per SM (ncu: 3 blocks at 78 registers left the warps 34 % active)
    lift_stats_kernel(LiftView lv, LiftGeom geo, int P, int C0, int grp, int32_t* __restrict__ gidx, float* __restrict__ y0,
                      double* __restrict__ sum, double* __restrict__ sumsq) {
    extern __shared__ double red[];                 // [2][rl][C0]
    const int tpr = C0 >> 2, rl = blockDim.x / tpr;
    const int t = threadIdx.x % tpr, lane = threadIdx.x / tpr;
    const int c = t * 4;
    const int n_groups = P / grp;
    const U4 u = load_u(lv, c);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!STORE && !sum) {      // inference on the tensor-core path: only the row indices are needed
        if (lv.z && t == 0)
            for (int g = blockIdx.x * rl + lane; g < n_groups; g += gridDim.x * rl)
                for (int s = 0; s < grp; ++s) gidx[g * grp + s] = lift_row(geo, g * grp + s);
        return;
    }
    double d1[4] = {0.0, 0.0, 0.0, 0.0}, d2[4] = {0.0, 0.0, 0.0, 0.0};
    for (int g = blockIdx.x * rl + lane; g < n_groups; g += gridDim.x * rl) {
        float4 a1 = zero, a2 = zero;
        const int pb = g * grp;
#pragma unroll 4
        for (int s = 0; s < grp; ++s) {
            const int p = pb + s;
            float4 z4 = zero;
            if (lv.z) {
                const int row = lift_row(geo, p);
                if (t == 0) gidx[p] = row;
                z4 = ldg4(lv.z + (size_t)row * lv.ldz + c);
            }
            const float4 sv = lv.s ? ldg4(lv.s + (size_t)p * 4) : zero;
            const float4 v = lift_val4(z4, sv, u.u0, u.u1, u.u2, u.u3);
            if (STORE) *reinterpret_cast<float4*>(y0 + (size_t)p * lv.ldz + c) = v;
            a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w;
            a2.x = fmaf(v.x, v.x, a2.x); a2.y = fmaf(v.y, v.y, a2.y); a2.z = fmaf(v.z, v.z, a2.z); a2.w = fmaf(v.w, v.w, a2.w);
            if ((s & 31) == 31 || s == grp - 1) {      // fp32 inside 32 positions, fp64 across (same rule as the GEMM epilogues)
                d1[0] += a1.x; d1[1] += a1.y; d1[2] += a1.z; d1[3] += a1.w;
                d2[0] += a2.x; d2[1] += a2.y; d2[2] += a2.z; d2[3] += a2.w;
                a1 = a2 = zero;
            }
        }
    }
    if (!sum) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[(size_t)lane * C0 + c + j] = d1[j];
        red[(size_t)(rl + lane) * C0 + c + j] = d2[j];
    }
    __syncthreads();
    for (int ch = threadIdx.x; ch < C0; ch += blockDim.x) {
        double s1 = 0.0, s2 = 0.0;
        for (int l = 0; l < rl; ++l) { s1 += red[(size_t)l * C0 + ch]; s2 += red[(size_t)(rl + l) * C0 + ch]; }
        atomicAdd(sum + ch, s1);
        atomicAdd(sumsq + ch, s2);
    }
}

__global__ void __launch_bounds__(LIFT_THREADS, 4)
    lift_scatter_kernel(LiftView lv, int P, int C0, int grp, const float* __restrict__ y0, const float* __restrict__ g, int ldg,
                        const float* __restrict__ ca, const float* __restrict__ cb, const float* __restrict__ ccf,
                        float* __restrict__ dz, float* __restrict__ ds, float* __restrict__ du) {
    extern __shared__ float redf[];                 // [rl][4][C0] (du only)
    const int tpr = C0 >> 2, rl = blockDim.x / tpr;
    const int t = threadIdx.x % tpr, lane = threadIdx.x / tpr;
    const int c = t * 4;
    const int n_groups = P / grp;
    const U4 u = load_u(lv, c);
    const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a4 = ca ? ldg4(ca + c) : one, b4 = ca ? ldg4(cb + c) : zero, k4 = ca ? ldg4(ccf + c) : zero;
    float4 du0 = zero, du1 = zero, du2 = zero, du3 = zero;
    // shuffle width for the per-row dot products ds[p][j] = dY0[p] . u[j]: the row's threads inside one warp (only when a warp
    // holds whole rows or a row holds whole warps; otherwise one RED per thread)
    const int w = tpr >= 32 ? 32 : tpr;
    const bool shfl = (tpr <= 32 && (tpr & (tpr - 1)) == 0) || (tpr % 32 == 0);
    // uniform trip count per block (the shuffles below need every lane of a warp): lanes past the last group idle
    for (int gb = blockIdx.x * rl; gb < n_groups; gb += gridDim.x * rl) {
        const bool valid = gb + lane < n_groups;
        const int gi = valid ? gb + lane : n_groups - 1;
        float4 acc0 = zero;
        const int pb = gi * grp;
        // ball-query padding repeats a group's FIRST neighbour in every unused slot (pointnet2_ops ball_query), so a sparse
        // ball sends most of its positions to one row: those are summed in registers and leave as ONE RED
        const int row0 = lv.z ? __ldg(lv.gidx + pb) : 0;
#pragma unroll 4
        for (int s = 0; s < grp; ++s) {
            const int p = pb + s;
            const int row = lv.z ? __ldg(lv.gidx + p) : 0;
            const float4 g4 = ldg4(g + (size_t)p * ldg + c);
            const float4 sv = lv.s ? ldg4(lv.s + (size_t)p * 4) : zero;
            float4 dy = g4;
            if (ca) {
                float4 v;
                if (y0) v = ldg4(y0 + (size_t)p * lv.ldz + c);
                else v = lift_val4(lv.z ? ldg4(lv.z + (size_t)row * lv.ldz + c) : zero, sv, u.u0, u.u1, u.u2, u.u3);
                dy.x = fmaf(a4.x, g4.x, fmaf(k4.x, v.x, b4.x)); dy.y = fmaf(a4.y, g4.y, fmaf(k4.y, v.y, b4.y));
                dy.z = fmaf(a4.z, g4.z, fmaf(k4.z, v.z, b4.z)); dy.w = fmaf(a4.w, g4.w, fmaf(k4.w, v.w, b4.w));
            }
            if (dz) {
                if (row == row0) { acc0.x += dy.x; acc0.y += dy.y; acc0.z += dy.z; acc0.w += dy.w; }
                else if (valid) atomicAdd(reinterpret_cast<float4*>(dz + (size_t)row * lv.ldz + c), dy);   // sm_90+: one vector RED
            }
            if (du && valid) {
                du0.x = fmaf(sv.x, dy.x, du0.x); du0.y = fmaf(sv.x, dy.y, du0.y); du0.z = fmaf(sv.x, dy.z, du0.z); du0.w = fmaf(sv.x, dy.w, du0.w);
                du1.x = fmaf(sv.y, dy.x, du1.x); du1.y = fmaf(sv.y, dy.y, du1.y); du1.z = fmaf(sv.y, dy.z, du1.z); du1.w = fmaf(sv.y, dy.w, du1.w);
                du2.x = fmaf(sv.z, dy.x, du2.x); du2.y = fmaf(sv.z, dy.y, du2.y); du2.z = fmaf(sv.z, dy.z, du2.z); du2.w = fmaf(sv.z, dy.w, du2.w);
                du3.x = fmaf(sv.w, dy.x, du3.x); du3.y = fmaf(sv.w, dy.y, du3.y); du3.z = fmaf(sv.w, dy.z, du3.z); du3.w = fmaf(sv.w, dy.w, du3.w);
            }
            if (ds) {
                float4 d;
                d.x = fmaf(dy.x, u.u0.x, fmaf(dy.y, u.u0.y, fmaf(dy.z, u.u0.z, dy.w * u.u0.w)));
                d.y = fmaf(dy.x, u.u1.x, fmaf(dy.y, u.u1.y, fmaf(dy.z, u.u1.z, dy.w * u.u1.w)));
                d.z = fmaf(dy.x, u.u2.x, fmaf(dy.y, u.u2.y, fmaf(dy.z, u.u2.z, dy.w * u.u2.w)));
                d.w = fmaf(dy.x, u.u3.x, fmaf(dy.y, u.u3.y, fmaf(dy.z, u.u3.z, dy.w * u.u3.w)));
                if (shfl) {
                    for (int o = w >> 1; o >= 1; o >>= 1) {
                        d.x += __shfl_xor_sync(0xFFFFFFFFu, d.x, o, 32); d.y += __shfl_xor_sync(0xFFFFFFFFu, d.y, o, 32);
                        d.z += __shfl_xor_sync(0xFFFFFFFFu, d.z, o, 32); d.w += __shfl_xor_sync(0xFFFFFFFFu, d.w, o, 32);
                    }
                    if ((t & (w - 1)) == 0 && valid) atomicAdd(reinterpret_cast<float4*>(ds + (size_t)p * 4), d);
                } else if (valid) {
                    atomicAdd(reinterpret_cast<float4*>(ds + (size_t)p * 4), d);
                }
            }
        }
        if (dz && valid) atomicAdd(reinterpret_cast<float4*>(dz + (size_t)row0 * lv.ldz + c), acc0);
    }
    if (!du) return;
    float* r = redf + (size_t)lane * 4 * C0 + c;
    *reinterpret_cast<float4*>(r) = du0;
    *reinterpret_cast<float4*>(r + C0) = du1;
    *reinterpret_cast<float4*>(r + 2 * C0) = du2;
    *reinterpret_cast<float4*>(r + 3 * C0) = du3;
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * C0; e += blockDim.x) {
        float s1 = 0.f;
        for (int l = 0; l < rl; ++l) s1 += redf[(size_t)l * 4 * C0 + e];
        atomicAdd(du + e, s1);        // du is [4, C0] contiguous: e = j * C0 + channel
    }
}

int lift_check(const o3d_lift_t* lf, int P, int C0, const char* who) {
    O3D_REQUIRE(lf && (lf->z || lf->s), O3D_ERR_ARG, "%s: null lift descriptor (need z and / or s)", who);
    O3D_REQUIRE(C0 >= 4 && (C0 & 3) == 0 && C0 <= 1024 && lf->ldz == C0, O3D_ERR_ARG, "%s: C0=%d ldz=%d (need C0 %% 4 == 0, ldz == C0)",
                who, C0, lf->ldz);
    O3D_REQUIRE(lf->grp >= 1 && (lf->grp & (lf->grp - 1)) == 0 && P % lf->grp == 0, O3D_ERR_ARG,
                "%s: grp=%d must be a power of two dividing P=%d", who, lf->grp, P);
    O3D_REQUIRE(!lf->z || (lf->pos_per_cloud >= 1 && lf->rows_per_cloud >= 1 && (lf->ridx || lf->ridx_mod >= 1)), O3D_ERR_ARG,
                "%s: bad cloud geometry", who);
    O3D_REQUIRE(((uintptr_t)lf->s & 15) == 0 && ((uintptr_t)lf->u & 15) == 0 && ((uintptr_t)lf->z & 15) == 0, O3D_ERR_ALIGN,
                "%s: z / s / u must be 16-byte aligned", who);
    O3D_REQUIRE((lf->s == nullptr) == (lf->u == nullptr), O3D_ERR_ARG, "%s: s and u come together", who);
    return O3D_OK;
}

inline LiftView make_view(const o3d_lift_t* lf, const int32_t* gidx) {
    return LiftView{lf->z, lf->ldz, lf->z ? gidx : nullptr, lf->s, lf->u};
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
    O3D_REQUIRE(gidx || !lf->z, O3D_ERR_ARG, "o3d_lift_stats: gidx workspace missing");
    if (P == 0 || (!lf->z && !y0 && !sum)) return O3D_OK;      // nothing to gather, store or count
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
    O3D_REQUIRE((gidx || !lf->z) && g && (ldg & 3) == 0, O3D_ERR_ARG, "o3d_lift_scatter: null pointer / ldg");
    O3D_REQUIRE(!lf->d_s || lf->u, O3D_ERR_ARG, "o3d_lift_scatter: d_s without u");
    O3D_REQUIRE(((uintptr_t)lf->d_s & 15) == 0 && ((uintptr_t)lf->d_z & 15) == 0, O3D_ERR_ALIGN, "o3d_lift_scatter: alignment");
    if (P == 0) return O3D_OK;
    int threads, blocks, rl;
    lift_launch_shape(P, C0, lf->grp, threads, blocks, rl);
    const LiftView lv = make_view(lf, gidx);
    const size_t smem = (lf->s && lf->d_u) ? sizeof(float) * rl * 4 * C0 : 0;
    lift_scatter_kernel<<<blocks, threads, smem, (cudaStream_t)stream>>>(lv, P, C0, lf->grp, y0, g, ldg, a, b, cc,
                                                                        lf->z ? lf->d_z : nullptr, lf->s ? lf->d_s : nullptr,
                                                                        lf->s ? lf->d_u : nullptr);
    O3D_CHECK_LAUNCH("o3d_lift_scatter");
    return O3D_OK;
}
