// Box-frame crop of LiDAR scans — the per-frame / per-pair geometry of the tracking loop and of the training sampler
// (reference: datasets/points_utils.py generate_subwindow :223-254, cropAndCenterPC :102-124, crop_pc_axis_aligned :147-173).
//
//   local[b, i, :] = R[b]^T (scan[frame[b], i, :] - c[b])            points of sample b's scan in the frame of its box
//   keep[b, i]     = i < count[frame[b]]  &&  |local| < half[b]      strictly inside the scaled + padded box, per axis
//
// One pass over the scans: the frame gather, the rigid transform, the three comparisons and the padding mask that the
// tensor formulation spreads over a batched 3x3 GEMM and a dozen elementwise kernels.  HBM-bound: 12 B read, 13 B written
// per point; thread = point, a warp reads 384 contiguous bytes.
#include "common.cuh"
#include "../../include/o3d_b200.h"

namespace {

__global__ void __launch_bounds__(256)
    crop_box_frame_kernel(const float* __restrict__ scans, const long long* __restrict__ count, const long long* __restrict__ frame,
                          const float* __restrict__ center, const float* __restrict__ rot, const float* __restrict__ half, int N,
                          float* __restrict__ local, uint8_t* __restrict__ keep) {
    const int b = blockIdx.y;
    const long long f = frame ? frame[b] : b;
    const int n_valid = count ? (int)min((long long)N, count[f]) : N;
    const float cx = center[b * 3 + 0], cy = center[b * 3 + 1], cz = center[b * 3 + 2];
    const float* R = rot + b * 9;
    const float r00 = R[0], r01 = R[1], r02 = R[2], r10 = R[3], r11 = R[4], r12 = R[5], r20 = R[6], r21 = R[7], r22 = R[8];
    const float hx = half[b * 3 + 0], hy = half[b * 3 + 1], hz = half[b * 3 + 2];
    const float* __restrict__ src = scans + (size_t)f * N * 3;
    float* __restrict__ dst = local + (size_t)b * N * 3;
    uint8_t* __restrict__ k = keep + (size_t)b * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float dx = src[i * 3 + 0] - cx, dy = src[i * 3 + 1] - cy, dz = src[i * 3 + 2] - cz;
        // row vector times R: the columns of R are the box axes
        const float x = fmaf(dz, r20, fmaf(dy, r10, dx * r00));
        const float y = fmaf(dz, r21, fmaf(dy, r11, dx * r01));
        const float z = fmaf(dz, r22, fmaf(dy, r12, dx * r02));
        dst[i * 3 + 0] = x;
        dst[i * 3 + 1] = y;
        dst[i * 3 + 2] = z;
        k[i] = (uint8_t)(i < n_valid && fabsf(x) < hx && fabsf(y) < hy && fabsf(z) < hz);
    }
}

}  // namespace

extern "C" int o3d_crop_box_frame(const float* scans, const long long* count, const long long* frame, const float* center,
                                  const float* rot, const float* half, int B, int N, float* local, unsigned char* keep,
                                  void* stream) {
    O3D_REQUIRE(scans && center && rot && half && local && keep, O3D_ERR_ARG, "o3d_crop_box_frame: null pointer");
    O3D_REQUIRE(B >= 0 && N >= 0 && B <= 65535, O3D_ERR_ARG, "o3d_crop_box_frame: bad sizes B=%d N=%d", B, N);
    if (B == 0 || N == 0) return O3D_OK;
    int gx = (N + 255) / 256;
    const int cap = (8 * o3d_num_sms() + B - 1) / B;          // ~8 blocks per SM over the whole batch
    if (gx > cap) gx = cap < 1 ? 1 : cap;
    crop_box_frame_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(scans, count, frame, center, rot, half, N, local, keep);
    O3D_CHECK_LAUNCH("o3d_crop_box_frame");
    return O3D_OK;
}
