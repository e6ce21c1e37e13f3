// Device-side view of a "lifted" first layer (include/o3d_b200.h: o3d_lift_t):
//     Y0[p, c] = Z[gidx[p], c] + s[p].x * u[0][c] + s[p].y * u[1][c] + s[p].z * u[2][c] + s[p].w * u[3][c]
// (Z optional, s/u optional).  Every consumer (statistics pass, operand loaders, dgrad epilogue, scatter) evaluates it through
// lift_val() — one fixed fma chain — so the value, and with it every ReLU-mask decision derived from it, is bit-identical
// everywhere.
#pragma once
#include <stdint.h>

struct LiftView {
    const float* z; int ldz;      // [zrows, ldz] or nullptr
    const int32_t* gidx;          // [P] global Z row per position (written by the forward statistics pass); nullptr iff z == nullptr
    const float* s;               // [P, 4] per-position scalars or nullptr
    const float* u;               // [4, ldz] their weight rows (nullptr iff s == nullptr)
};

#ifdef __CUDACC__
__device__ __forceinline__ float lift_val(float z, const float4& s, float u0, float u1, float u2, float u3) {
    return fmaf(s.w, u3, fmaf(s.z, u2, fmaf(s.y, u1, fmaf(s.x, u0, z))));
}
__device__ __forceinline__ float4 lift_val4(const float4& z, const float4& s, const float4& u0, const float4& u1, const float4& u2,
                                            const float4& u3) {
    return make_float4(lift_val(z.x, s, u0.x, u1.x, u2.x, u3.x), lift_val(z.y, s, u0.y, u1.y, u2.y, u3.y),
                       lift_val(z.z, s, u0.z, u1.z, u2.z, u3.z), lift_val(z.w, s, u0.w, u1.w, u2.w, u3.w));
}
#endif
