// Tensor-core (tcgen05 / TMEM) implementation of the point-wise MLP forward and data-gradient GEMMs, 3xTF32.
//
// Same contract as pw_fwd_kernel / pw_dgrad_kernel in pwmlp.cu (which remain the exact-fp32 ground truth and
// serve the shapes this kernel does not take: K < 32, fewer than 128 output channels, ragged channel counts).
//
//   D[ch, pos] = sum_k  Wmat[ch, k] * Act[pos, k]          ch tile = 128 (UMMA M), pos tile = 128 (UMMA N)
//
// with Act produced on the fly from global memory (forward: relu(bn(Y_prev)); dgrad: dY = a*g + b + c*Y) and split
// into a TF32 "hi" part (the fp32 word with its 13 low mantissa bits cleared) and a "lo" part
// (x - hi, exact), so that   Whi*Xhi + Wlo*Xhi + Whi*Xlo   carries ~21 mantissa bits — fp32-grade accuracy, which
// the 1e-4 parity bar needs and a single TF32 pass (10 bits) cannot give.
//
// Roles (384 threads, one persistent CTA per SM, all roles walk the same static tile sequence):
//   warp 0      allocates TMEM (2 x 128 fp32 columns: double-buffered accumulator) and, one elected lane, issues
//               tcgen05.mma.cta_group::1.kind::tf32 (M128 N128 K8), 12 per 32-channel k-block, committing each
//               stage back to the producers and each finished tile to the epilogue through mbarriers
//   warp 1      one lane streams the pre-tiled, pre-swizzled weight images (hi|lo, 32 KB per k-block) with
//               cp.async.bulk (UBLKCP) onto the stage's "full" barrier
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns; lane = output channel, columns = positions, so the
//               batch statistics, the group max/min/arg and the ReLU-mask sums are plain per-thread loops and every
//               global store of a warp is one coalesced 128-byte line
//   warps 8-11  operand producers: coalesced 16-byte loads, transform, hi/lo split, 128B-swizzled st.shared,
//               fence.proxy.async, arrive
// Shared memory: 3 stages x (W_hi 16K | W_lo 16K | X_hi 16K | X_lo 16K) = 192 KB, K-major SWIZZLE_128B tiles.
#include "common.cuh"
#include "../../include/o3d_b200.h"

namespace {

constexpr int TC_M = 128;       // channels per tile
constexpr int TC_N = 128;       // positions per tile
constexpr int TC_K = 32;        // tf32 elements per k-block = 128 bytes per row
constexpr int TC_STAGES = 3;
constexpr int TILE_BYTES = TC_M * TC_K * 4;            // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;            // Whi | Wlo | Xhi | Xlo
constexpr int TC_THREADS = 384;
constexpr int TC_SMEM = TC_STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 256;

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(o3d_smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(o3d_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 (ignored for
// swizzled K-major) | SBO = 1024 B between 8-row groups | version 1 | layout 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D=F32 (bits 4-5 = 1), A=B=TF32 (2 at bits 7-9 / 10-12), K-major both, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of the 16-byte chunk `c` (0..7) of row `r` inside a [rows x 32 tf32] SWIZZLE_128B K-major tile
__device__ __host__ __forceinline__ uint32_t sw128(int r, int c) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// hi = x with the 13 low mantissa bits cleared (exactly a TF32 value, so the tensor core's own fp32->tf32 conversion,
// truncating or rounding, leaves it unchanged); lo = x - hi (exact in fp32).
__device__ __forceinline__ float hi1(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float4 hi_part(const float4& v) { return make_float4(hi1(v.x), hi1(v.y), hi1(v.z), hi1(v.w)); }
__device__ __forceinline__ float4 lo_part(const float4& v) {
    return make_float4(v.x - hi1(v.x), v.y - hi1(v.y), v.z - hi1(v.z), v.w - hi1(v.w));
}
__device__ __forceinline__ float4 ld4g(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- operand descriptions (same semantics as ActIn / DyIn in pwmlp.cu) --------------------------------------
struct TcAct {
    const float* x; int ld; const float* scale; const float* shift; int relu;
    __device__ __forceinline__ float4 load(int p, int P, int k, int K) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < P && k < K) {
            v = ld4g(x + (size_t)p * ld + k);
            if (scale) {
                const float4 s = ld4g(scale + k), t = ld4g(shift + k);
                v.x = fmaf(v.x, s.x, t.x); v.y = fmaf(v.y, s.y, t.y); v.z = fmaf(v.z, s.z, t.z); v.w = fmaf(v.w, s.w, t.w);
            }
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        }
        return v;
    }
};

struct TcDy {
    const float* g; int ldg; const float* y; int ldy; const float* a; const float* b; const float* cc;
    const float* dpool; const int32_t* sel; int S; int ldp;
    __device__ __forceinline__ float4 load(int p, int P, int c, int C) const {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < P && c < C) {
            if (dpool) {
                const int grp = p / S, s = p - grp * S;
                const int4 sl = __ldg(reinterpret_cast<const int4*>(sel + (size_t)grp * ldp + c));
                const float4 d = ld4g(dpool + (size_t)grp * ldp + c);
                v.x = sl.x == s ? d.x : 0.f; v.y = sl.y == s ? d.y : 0.f; v.z = sl.z == s ? d.z : 0.f; v.w = sl.w == s ? d.w : 0.f;
            } else {
                v = ld4g(g + (size_t)p * ldg + c);
            }
            if (a) {
                const float4 aa = ld4g(a + c), bb = ld4g(b + c), c2 = ld4g(cc + c);
                const float4 yy = ld4g(y + (size_t)p * ldy + c);
                v.x = fmaf(aa.x, v.x, fmaf(c2.x, yy.x, bb.x)); v.y = fmaf(aa.y, v.y, fmaf(c2.y, yy.y, bb.y));
                v.z = fmaf(aa.z, v.z, fmaf(c2.z, yy.z, bb.z)); v.w = fmaf(aa.w, v.w, fmaf(c2.w, yy.w, bb.w));
            }
        }
        return v;
    }
};

// ---- epilogues: thread = one output channel `ch`, called once per 32-position column group ------------------
struct TcFwdEpi {
    float* y; int ldy; const float* bias; double* sum; double* sumsq;
    int S; float* ymax; float* ymin; int32_t* arg; int ldp;
    // per-thread running state (fp32 inside a 32-position group, fp64 across groups and tiles)
    float bv, mx, mn; int ax, an; double d1, d2;
    __device__ __forceinline__ void begin(int ch, int Nw) {
        d1 = d2 = 0.0;
        bv = (bias && ch < Nw) ? bias[ch] : 0.f;
        mx = -INFINITY; mn = INFINITY; ax = an = 0;
    }
    __device__ __forceinline__ void group(const uint32_t (&r)[32], int ch, int Nw, int pbase, int P) {
        if (ch >= Nw) return;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int p = pbase + j;
            if (p >= P) break;
            const float v = __uint_as_float(r[j]) + bv;
            if (y) y[(size_t)p * ldy + ch] = v;
            s1 += v;
            s2 = fmaf(v, v, s2);
            if (S > 0) {
                const int s = p % S;
                if (s == 0) { mx = -INFINITY; mn = INFINITY; ax = an = 0; }
                if (v > mx) { mx = v; ax = s; }
                if (v < mn) { mn = v; an = s; }
                if (s == S - 1) {
                    const size_t o = (size_t)(p / S) * ldp + ch;
                    ymax[o] = mx; ymin[o] = mn; arg[o] = ax | (an << 16);
                }
            }
        }
        d1 += (double)s1;
        d2 += (double)s2;
    }
    __device__ __forceinline__ void end(int ch, int Nw) {
        if (sum && ch < Nw) {
            atomicAdd(sum + ch, d1);
            atomicAdd(sumsq + ch, d2);
        }
    }
};

struct TcDgradEpi {
    float* out; int ldo; const float* yprev; int ldyp; const float* scale; const float* shift; int relu;
    double* s1g; double* s2y;
    float sc, sh; double d1, d2;
    __device__ __forceinline__ void begin(int ch, int Nw) {
        d1 = d2 = 0.0;
        sc = (scale && ch < Nw) ? scale[ch] : 1.f;
        sh = (shift && ch < Nw) ? shift[ch] : 0.f;
    }
    __device__ __forceinline__ void group(const uint32_t (&r)[32], int ch, int Nw, int pbase, int P) {
        if (ch >= Nw) return;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int p = pbase + j;
            if (p >= P) break;
            float v = __uint_as_float(r[j]);
            if (yprev) {
                const float yv = __ldg(yprev + (size_t)p * ldyp + ch);
                if (relu && !(fmaf(yv, sc, sh) > 0.f)) v = 0.f;
                s2 = fmaf(v, yv, s2);
            }
            s1 += v;
            out[(size_t)p * ldo + ch] = v;
        }
        d1 += (double)s1;
        d2 += (double)s2;
    }
    __device__ __forceinline__ void end(int ch, int Nw) {
        if (s1g && ch < Nw) {
            atomicAdd(s1g + ch, d1);
            atomicAdd(s2y + ch, d2);
        }
    }
};

// ------------------------------------------------------------------------------------------------------------
template <class BLoad, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
    pw_tc_kernel(BLoad bl, const uint8_t* __restrict__ wtiles, int P, int K, int Nw, int nkb, Epi epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * STAGE_BYTES);
    uint64_t* full = bars;                      // [TC_STAGES]  producers + weight copy -> MMA
    uint64_t* empty = bars + TC_STAGES;         // [TC_STAGES]  MMA (tcgen05.commit) -> producers
    uint64_t* tfull = bars + 2 * TC_STAGES;     // [2]          MMA -> epilogue
    uint64_t* tempty = bars + 2 * TC_STAGES + 2;  // [2]        epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.y;
    const int n_ptiles = (P + TC_N - 1) / TC_N;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            o3d_mbar_init(full + s, 128 + 1);
            o3d_mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            o3d_mbar_init(tfull + a, 1);
            o3d_mbar_init(tempty + a, 128);
        }
        o3d_fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== MMA issuer
        const uint32_t idesc = make_idesc(TC_M, TC_N);
        int stage = 0, phase = 0, acc = 0, aphase = 0;
        for (int t = blockIdx.x; t < n_ptiles; t += gridDim.x) {
            o3d_mbar_wait(tempty + acc, aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_N);
            for (int kb = 0; kb < nkb; ++kb) {
                o3d_mbar_wait(full + stage, phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sb = o3d_smem_u32(smem + stage * STAGE_BYTES);
                    const uint64_t whi = make_desc(sb), wlo = make_desc(sb + TILE_BYTES);
                    const uint64_t xhi = make_desc(sb + 2 * TILE_BYTES), xlo = make_desc(sb + 3 * TILE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < TC_K / 8; ++ks) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);   // +32 bytes along K inside the 128B swizzle row
                        umma_tf32(d_tmem, wlo + adv, xhi + adv, idesc, (kb | ks) != 0);
                        umma_tf32(d_tmem, whi + adv, xlo + adv, idesc, 1u);
                        umma_tf32(d_tmem, whi + adv, xhi + adv, idesc, 1u);
                    }
                    umma_commit(empty + stage);                           // frees the stage when these MMAs retire
                    if (kb == nkb - 1) umma_commit(tfull + acc);          // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================================================== weight-tile streamer (bulk copy engine)
        if (lane == 0) {
            int stage = 0, phase = 0;
            for (int t = blockIdx.x; t < n_ptiles; t += gridDim.x) {
                for (int kb = 0; kb < nkb; ++kb) {
                    o3d_mbar_wait(empty + stage, phase ^ 1);
                    o3d_mbar_expect_tx(full + stage, 2 * TILE_BYTES);
                    o3d_bulk_g2s(smem + stage * STAGE_BYTES, wtiles + ((size_t)m_tile * nkb + kb) * (2 * TILE_BYTES),
                                 2 * TILE_BYTES, full + stage);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================================================== epilogue
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int ch = m_tile * TC_M + q * 32 + lane;
        epi.begin(ch, Nw);
        int acc = 0, aphase = 0;
        for (int t = blockIdx.x; t < n_ptiles; t += gridDim.x) {
            o3d_mbar_wait(tfull + acc, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_N);
#pragma unroll 1
            for (int cg = 0; cg < TC_N / 32; ++cg) {
                uint32_t r[32];
                tmem_ld32(taddr + cg * 32, r);
                epi.group(r, ch, Nw, t * TC_N + cg * 32, P);
            }
            tc_fence_before();
            o3d_mbar_arrive(tempty + acc);
            if (++acc == 2) { acc = 0; aphase ^= 1; }
        }
        epi.end(ch, Nw);
    } else if (warp >= 8) {
        // ===================================================== activation-operand producers (128 threads)
        const int pt = threadIdx.x - 256;             // 0..127
        const int chunk = pt & 7;                     // 16-byte chunk (4 channels) inside the 128-byte row
        const int row0 = pt >> 3;                     // rows row0 + 16*i
        int stage = 0, phase = 0;
        for (int t = blockIdx.x; t < n_ptiles; t += gridDim.x) {
            const int p0 = t * TC_N;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = bl.load(p0 + row0 + 16 * i, P, chunk * 4, K);
            for (int kb = 0; kb < nkb; ++kb) {
                float4 nx[8];
                const bool more = kb + 1 < nkb;
                if (more) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) nx[i] = bl.load(p0 + row0 + 16 * i, P, (kb + 1) * TC_K + chunk * 4, K);
                }
                o3d_mbar_wait(empty + stage, phase ^ 1);
                uint8_t* xhi = smem + stage * STAGE_BYTES + 2 * TILE_BYTES;
                uint8_t* xlo = xhi + TILE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t off = sw128(row0 + 16 * i, chunk);
                    *reinterpret_cast<float4*>(xhi + off) = hi_part(v[i]);
                    *reinterpret_cast<float4*>(xlo + off) = lo_part(v[i]);
                }
                o3d_fence_proxy_async();              // generic-proxy stores -> visible to the tensor core (async proxy)
                o3d_mbar_arrive(full + stage);
                if (more) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = nx[i];
                }
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// Pre-tile a weight matrix W[rows, ld] (rows = UMMA M channels, k contiguous) into the per-(m_tile, k-block) shared-memory
// images the kernel bulk-copies: [hi 16 KB | lo 16 KB], K-major SWIZZLE_128B, zero padded.
__global__ void w_pretile_kernel(const float* __restrict__ W, int ld, int rows, int K, int nkb, uint8_t* __restrict__ out) {
    const int m_tile = blockIdx.y, kb = blockIdx.x;
    uint8_t* dst = out + ((size_t)m_tile * nkb + kb) * (2 * TILE_BYTES);
    for (int id = threadIdx.x; id < TC_M * 8; id += blockDim.x) {
        const int r = id >> 3, c = id & 7;
        const int row = m_tile * TC_M + r, k = kb * TC_K + c * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows && k < K) v = *reinterpret_cast<const float4*>(W + (size_t)row * ld + k);
        const uint32_t off = sw128(r, c);
        *reinterpret_cast<float4*>(dst + off) = hi_part(v);
        *reinterpret_cast<float4*>(dst + TILE_BYTES + off) = lo_part(v);
    }
}

template <class BLoad, class Epi>
int launch_tc(BLoad bl, const uint8_t* wtiles, int P, int K, int Nw, Epi epi, cudaStream_t st, const char* name) {
    auto kern = pw_tc_kernel<BLoad, Epi>;
    O3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM), name);
    const int mt = (Nw + TC_M - 1) / TC_M;
    const int nkb = (K + TC_K - 1) / TC_K;
    const int n_ptiles = (P + TC_N - 1) / TC_N;
    int gx = o3d_num_sms() / mt;
    if (gx < 1) gx = 1;
    if (gx > n_ptiles) gx = n_ptiles;
    kern<<<dim3(gx, mt), TC_THREADS, TC_SMEM, st>>>(bl, wtiles, P, K, Nw, nkb, epi);
    O3D_CHECK_LAUNCH(name);
    return O3D_OK;
}

}  // namespace

extern "C" long long o3d_pw_tc_wtile_bytes(int rows, int K) {
    const long long mt = (rows + TC_M - 1) / TC_M, nkb = (K + TC_K - 1) / TC_K;
    return mt * nkb * 2 * TILE_BYTES;
}

extern "C" int o3d_pw_tc_pretile(const float* w, int ldw, int rows, int K, void* wtiles, void* stream) {
    O3D_REQUIRE(w && wtiles, O3D_ERR_ARG, "o3d_pw_tc_pretile: null pointer");
    O3D_REQUIRE((K & 3) == 0 && (ldw & 3) == 0, O3D_ERR_ARG, "o3d_pw_tc_pretile: K and ldw must be multiples of 4");
    O3D_REQUIRE(((uintptr_t)wtiles & 15) == 0 && ((uintptr_t)w & 15) == 0, O3D_ERR_ALIGN, "o3d_pw_tc_pretile: alignment");
    const int mt = (rows + TC_M - 1) / TC_M, nkb = (K + TC_K - 1) / TC_K;
    w_pretile_kernel<<<dim3(nkb, mt), 256, 0, (cudaStream_t)stream>>>(w, ldw, rows, K, nkb, (uint8_t*)wtiles);
    O3D_CHECK_LAUNCH("o3d_pw_tc_pretile");
    return O3D_OK;
}

extern "C" int o3d_pw_fwd_tc(const float* x, int ldx, const float* in_scale, const float* in_shift, int in_relu,
                             const void* wtiles, const float* bias, int P, int K, int N, float* y, int ldy, double* sum,
                             double* sumsq, int S, float* ymax, float* ymin, int32_t* arg, int ldp, void* stream) {
    O3D_REQUIRE(x && wtiles, O3D_ERR_ARG, "o3d_pw_fwd_tc: null pointer");
    O3D_REQUIRE(P >= 0 && K >= 4 && N >= 1 && (K & 3) == 0 && (ldx & 3) == 0, O3D_ERR_ARG, "o3d_pw_fwd_tc: bad sizes");
    O3D_REQUIRE(S == 0 || (P % S == 0 && ymax && ymin && arg), O3D_ERR_ARG, "o3d_pw_fwd_tc: bad pooling arguments");
    if (P == 0) return O3D_OK;
    const int Nw = (N + 3) & ~3;
    TcAct bl{x, ldx, in_scale, in_shift, in_relu};
    TcFwdEpi ep{};
    ep.y = y; ep.ldy = ldy; ep.bias = bias; ep.sum = sum; ep.sumsq = sumsq;
    ep.S = S; ep.ymax = ymax; ep.ymin = ymin; ep.arg = arg; ep.ldp = ldp;
    return launch_tc(bl, (const uint8_t*)wtiles, P, K, Nw, ep, (cudaStream_t)stream, "o3d_pw_fwd_tc");
}

extern "C" int o3d_pw_dgrad_tc(const float* g, int ldg, const float* y, int ldy, const float* a, const float* b,
                               const float* cc, const float* dpool, const int32_t* sel, int S, int ldp,
                               const void* wtiles_t, int P, int Cout, int Cin, float* out, int ldo, const float* yprev,
                               int ldyp, const float* pscale, const float* pshift, int prelu, double* s1, double* s2y,
                               void* stream) {
    O3D_REQUIRE((g || dpool) && wtiles_t && out, O3D_ERR_ARG, "o3d_pw_dgrad_tc: null pointer");
    O3D_REQUIRE((Cout & 3) == 0 && (Cin & 3) == 0, O3D_ERR_ARG, "o3d_pw_dgrad_tc: channel counts must be multiples of 4");
    if (P == 0) return O3D_OK;
    TcDy bl{g, ldg, y, ldy, a, b, cc, dpool, sel, S > 0 ? S : 1, ldp};
    TcDgradEpi ep{};
    ep.out = out; ep.ldo = ldo; ep.yprev = yprev; ep.ldyp = ldyp; ep.scale = pscale; ep.shift = pshift; ep.relu = prelu;
    ep.s1g = s1; ep.s2y = s2y;
    // GEMM: D[cin, pos] = sum_cout Wt[cin, cout] * dY[pos, cout]  ->  "K" = Cout, "Nw" = Cin
    return launch_tc(bl, (const uint8_t*)wtiles_t, P, Cout, Cin, ep, (cudaStream_t)stream, "o3d_pw_dgrad_tc");
}
