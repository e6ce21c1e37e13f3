// Adam over the flat parameter bucket (one launch per step), CUDA-graph friendly: the step count and the learning
// rate live in device memory, so a captured training step can be replayed without re-recording.
//
// Same update rule as the reference's optimizer, `torch.optim.Adam(lr, betas=(0.5, 0.999), eps=1e-6, weight_decay=wd)`
// (models/base_model.py:28-36):  g += wd*p;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//                                p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "common.cuh"
#include "../../include/o3d_b200.h"

namespace {

__global__ void adam_tick_kernel(float* __restrict__ state) { state[0] += 1.0f; }  // state = [step, lr]

__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                long long n, const float* __restrict__ state, float beta1, float beta2, float eps, float wd) {
    const float step = state[0], lr = state[1];
    const float bc1 = 1.0f - powf(beta1, step);
    const float bc2s = sqrtf(1.0f - powf(beta2, step));
    const float step_size = lr / bc1;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
#define O3D_ADAM1(c)                                           \
    {                                                          \
        float gr = gg.c + wd * pp.c;                           \
        mm.c = beta1 * mm.c + (1.0f - beta1) * gr;             \
        vv.c = beta2 * vv.c + (1.0f - beta2) * gr * gr;        \
        pp.c -= step_size * mm.c / (sqrtf(vv.c) / bc2s + eps); \
    }
        O3D_ADAM1(x) O3D_ADAM1(y) O3D_ADAM1(z) O3D_ADAM1(w)
#undef O3D_ADAM1
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float gr = g[i] + wd * p[i];
        m[i] = beta1 * m[i] + (1.0f - beta1) * gr;
        v[i] = beta2 * v[i] + (1.0f - beta2) * gr * gr;
        p[i] -= step_size * m[i] / (sqrtf(v[i]) / bc2s + eps);
    }
}

}  // namespace

extern "C" int o3d_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float* state,
                             float beta1, float beta2, float eps, float weight_decay, void* stream) {
    O3D_REQUIRE(param && grad && exp_avg && exp_avg_sq && state, O3D_ERR_ARG, "o3d_adam_step: null pointer");
    O3D_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, O3D_ERR_ALIGN,
                "o3d_adam_step: buffers must be 16-byte aligned");
    if (n <= 0) return O3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 1, 0, st>>>(state);
    long long blocks = ((n >> 2) + 255) / 256;
    const long long cap = (long long)o3d_num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps, weight_decay);
    O3D_CHECK_LAUNCH("o3d_adam_step");
    return O3D_OK;
}
