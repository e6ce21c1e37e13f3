// Warp-cooperative ball query shared by ball_query.cu (stand-alone / fused with grouping) and sa_fused.cu (the single-kernel
// inference SA layer).  Exact upstream semantics: see the header of ball_query.cu.
#pragma once
#include "common.cuh"

// Warp-cooperative ball query for one centre; writes the nsample indices to `o` (global or shared).
__device__ __forceinline__ void warp_ball_query(const float* s_xyz, int N, float nx, float ny, float nz, float radius2,
                                                int nsample, int32_t* o, int lane) {
    int cnt = 0, first = 0;
    const uint32_t lt = o3d_lanemask_lt();
    for (int k0 = 0; k0 < N && cnt < nsample; k0 += 32) {
        const int k = k0 + lane;
        bool hit = false;
        if (k < N) {
            const float d2 = o3d_dist2(nx, ny, nz, s_xyz[k * 3 + 0], s_xyz[k * 3 + 1], s_xyz[k * 3 + 2]);
            hit = d2 < radius2;
        }
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
        if (mask) {
            if (cnt == 0) first = k0 + __ffs(mask) - 1;
            const int pos = cnt + __popc(mask & lt);
            if (hit && pos < nsample) o[pos] = k;
            cnt += __popc(mask);
        }
    }
    if (cnt > nsample) cnt = nsample;
    const int pad = cnt == 0 ? 0 : first;
    for (int l = cnt + lane; l < nsample; l += 32) o[l] = pad;
}

