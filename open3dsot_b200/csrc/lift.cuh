// Device-side view of a "lifted" first layer (include/o3d_b200.h: o3d_lift_t):
//     Y0[p, c] = Z[gidx[p], c] - cc[p >> gsh, c] + s[p] * u[c]
// Every consumer (statistics pass, operand loaders, dgrad epilogue, scatter) evaluates it through lift_val() so the value —
// and with it every ReLU-mask decision derived from it — is bit-identical everywhere.
#pragma once
#include <stdint.h>

struct LiftView {
    const float* z; int ldz;
    const int32_t* gidx;     // [P] global Z row per position (written by the forward statistics pass)
    const float* cc; int gsh;   // [P >> gsh, ldz] or nullptr
    const float* s; const float* u;   // [P], [ldz] or nullptr
};

#ifdef __CUDACC__
__device__ __forceinline__ float lift_val(float z, float c, float s, float u) { return fmaf(s, u, z - c); }
__device__ __forceinline__ float4 lift_val4(const float4& z, const float4& c, float s, const float4& u) {
    return make_float4(lift_val(z.x, c.x, s, u.x), lift_val(z.y, c.y, s, u.y), lift_val(z.z, c.z, s, u.z),
                       lift_val(z.w, c.w, s, u.w));
}
#endif
