// K5 — GroupNorm(+SiLU) and LayerNorm, forward and input-gradient, on channels-last bf16 tensors.  HBM-bound:
// every pass reads/writes 16-byte vectors, fully coalesced along C; statistics in fp32 (fp64 for the cross-CTA
// GroupNorm combine).  Replaces ATen native_group_norm / layer_norm (+ F.silu) called by diffusers' ResnetBlock2D,
// Transformer2DModel, BasicTransformerBlock and by models.py:515-543 (ConvBlock2D).
#include <stdio.h>

#include "common.cuh"
#include "host_common.h"
#include "../../include/controllora_b200.h"

namespace clb {

__device__ __forceinline__ float silu_f(float z) { return z / (1.f + __expf(-z)); }
__device__ __forceinline__ float silu_grad_f(float z) {
    const float s = 1.f / (1.f + __expf(-z));
    return s * (1.f + z * (1.f - s));
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------------- GroupNorm
// Layout: x [n, HW, C]; group g = channels [g*cpg, (g+1)*cpg).  A CTA owns a slab of rows of one image; thread t
// always handles the same 8-channel chunk (blockDim.x is a multiple of C/8), so per-channel partial sums stay in
// registers; the per-group combine goes through shared memory and one fp64 atomic per (group, CTA).

static constexpr int GN_MAX_CHUNKS = 512;  // C <= 4096

// Per-thread channel constants: thread t owns channels [c0, c0+8) for its whole slab, so everything that depends only
// on (image, channel) is computed once and the row loop is load -> fma -> (store), unrolled for memory-level
// parallelism (one 16-byte load in flight per thread caps a 1024-thread SM at ~20 GB/s; four reach the HBM limit).
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ float fast_sigmoid(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }

// mode 0: forward statistics      -> ws[(n*G+g)*2 + {0,1}] += {sum x, sum x^2}
// mode 1: backward reductions     -> ws[(n*G+g)*2 + {0,1}] += {sum dz*gamma, sum dz*gamma*xhat},
//                                    dgamma[c] += sum dz*xhat, dbeta[c] += sum dz   (when dgamma != nullptr)
// In mode 1 the group sums are gamma-weighted channel sums of the dgamma/dbeta partials, so only those accumulate.
template <int MODE>
__global__ void __launch_bounds__(512)
gn_reduce_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                 const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
                 double* __restrict__ ws, float* __restrict__ dgamma, float* __restrict__ dbeta, int HW, int C, int G,
                 int rows_per_cta, int silu) {
    const int n = blockIdx.y;
    const int chunks = C / 8;
    const int cpg = C / G;
    const int rows_par = blockDim.x / chunks;      // rows processed per iteration
    const int chunk = threadIdx.x % chunks;
    const int rsub = threadIdx.x / chunks;
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(HW, row0 + rows_per_cta);
    const bool active = rsub < rows_par;
    const int c0 = chunk * 8;

    float a0[8], a1[8];  // MODE 0: sum, sumsq.  MODE 1: sum dz*xhat, sum dz
    float ga[8], gb[8], rs[8], nm[8];   // MODE 1: z = xh*ga + gb, xh = x*rs + nm
#pragma unroll
    for (int j = 0; j < 8; ++j) { a0[j] = a1[j] = 0.f; }
    if (MODE == 1 && active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ga[j] = gamma[c0 + j];
            gb[j] = beta[c0 + j];
            const int g = (c0 + j) / cpg;
            const float mu = stats[(n * G + g) * 2];
            rs[j] = stats[(n * G + g) * 2 + 1];
            nm[j] = -mu * rs[j];
        }
    }
    if (active) {
        const __nv_bfloat16* xp = x + (long long)n * HW * C + c0;
        const __nv_bfloat16* dp = dy + (long long)n * HW * C + c0;
        const long long rstep = (long long)rows_par * C;
        int r = row0 + rsub;
        if (MODE == 0) {
            for (; r + 3 * rows_par < row1; r += 4 * rows_par) {
                uint4 u[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) u[k] = ldg16(xp + (long long)r * C + k * rstep);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float xv[8];
                    unpack8(u[k], xv);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { a0[j] += xv[j]; a1[j] = fmaf(xv[j], xv[j], a1[j]); }
                }
            }
            for (; r < row1; r += rows_par) {
                float xv[8];
                unpack8(ldg16(xp + (long long)r * C), xv);
#pragma unroll
                for (int j = 0; j < 8; ++j) { a0[j] += xv[j]; a1[j] = fmaf(xv[j], xv[j], a1[j]); }
            }
        } else {
            auto acc = [&](const uint4& ux, const uint4& ud) {
                float xv[8], dv[8];
                unpack8(ux, xv);
                unpack8(ud, dv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = fmaf(xv[j], rs[j], nm[j]);
                    float dz = dv[j];
                    if (silu) {
                        const float z = fmaf(xh, ga[j], gb[j]);
                        const float s = fast_sigmoid(z);
                        dz *= s * (1.f + z * (1.f - s));
                    }
                    a0[j] = fmaf(dz, xh, a0[j]);
                    a1[j] += dz;
                }
            };
            for (; r + 3 * rows_par < row1; r += 4 * rows_par) {
                uint4 ux[4], ud[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ux[k] = ldg16(xp + (long long)r * C + k * rstep);
                    ud[k] = ldg16(dp + (long long)r * C + k * rstep);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) acc(ux[k], ud[k]);
            }
            for (; r < row1; r += rows_par) acc(ldg16(xp + (long long)r * C), ldg16(dp + (long long)r * C));
        }
    }
    // combine without shared-memory atomics (ATOMS costs ~2 cycles per lane): every row-group stores its per-channel
    // partials, then the CTA sums them per channel, per group (fp64 atomics to global) and per parameter.
    extern __shared__ float gsh[];                        // [2][rows_par + 1][C]
    float* part0 = gsh;                                   // [rows_par][C]
    float* part1 = gsh + (size_t)rows_par * C;            // [rows_par][C]
    float* tot0 = gsh + (size_t)2 * rows_par * C;         // [C]
    float* tot1 = tot0 + C;
    if (active) {
        float* d0 = part0 + (size_t)rsub * C + c0;
        float* d1 = part1 + (size_t)rsub * C + c0;
        *reinterpret_cast<float4*>(d0) = make_float4(a0[0], a0[1], a0[2], a0[3]);
        *reinterpret_cast<float4*>(d0 + 4) = make_float4(a0[4], a0[5], a0[6], a0[7]);
        *reinterpret_cast<float4*>(d1) = make_float4(a1[0], a1[1], a1[2], a1[3]);
        *reinterpret_cast<float4*>(d1 + 4) = make_float4(a1[4], a1[5], a1[6], a1[7]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s0 = 0.f, s1 = 0.f;
        for (int g = 0; g < rows_par; ++g) { s0 += part0[(size_t)g * C + c]; s1 += part1[(size_t)g * C + c]; }
        tot0[c] = s0;
        tot1[c] = s1;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s0 = 0.0, s1 = 0.0;
        if (MODE == 0) {
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) { s0 += tot0[c]; s1 += tot1[c]; }
        } else {
            for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
                const float gm = gamma[c];
                s0 += (double)(gm * tot1[c]);   // sum dz*gamma
                s1 += (double)(gm * tot0[c]);   // sum dz*gamma*xhat
            }
        }
        atomicAdd(&ws[(n * G + g) * 2], s0);
        atomicAdd(&ws[(n * G + g) * 2 + 1], s1);
    }
    if (MODE == 1 && dgamma != nullptr) {
        for (int i = threadIdx.x; i < C; i += blockDim.x) {
            atomicAdd(&dgamma[i], tot0[i]);
            atomicAdd(&dbeta[i], tot1[i]);
        }
    }
}

// ws (fp64 sums) -> stats {mean, rstd} in fp32, one thread per (image, group)
__global__ void gn_finalize_kernel(const double* __restrict__ ws, float* __restrict__ stats, int total, double inv_m, float eps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double mu = ws[2 * i] * inv_m;
    double var = ws[2 * i + 1] * inv_m - mu * mu;
    if (var < 0.0) var = 0.0;
    stats[2 * i] = (float)mu;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
}

// forward apply: y = act((x - mean) * rstd * gamma + beta) with the fp32 stats {mean, rstd} of gn_finalize_kernel.
// Same slab geometry as the reduction: the affine pair (scale, shift) per owned channel lives in registers.
__global__ void __launch_bounds__(512)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ stats, __nv_bfloat16* __restrict__ y, int HW, int C, int G, int silu,
                int rows_per_cta) {
    const int n = blockIdx.y;
    const int chunks = C / 8;
    const int cpg = C / G;
    const int rows_par = blockDim.x / chunks;
    const int chunk = threadIdx.x % chunks;
    const int rsub = threadIdx.x / chunks;
    if (rsub >= rows_par) return;
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(HW, row0 + rows_per_cta);
    const int c0 = chunk * 8;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c0 + j) / cpg;
        const float mu = stats[(n * G + g) * 2], rs = stats[(n * G + g) * 2 + 1];
        sc[j] = rs * gamma[c0 + j];
        sh[j] = fmaf(-mu, sc[j], beta[c0 + j]);
    }
    const __nv_bfloat16* xp = x + (long long)n * HW * C + c0;
    __nv_bfloat16* yp = y + (long long)n * HW * C + c0;
    const long long rstep = (long long)rows_par * C;
    auto emit = [&](const uint4& u, __nv_bfloat16* dst) {
        float xv[8], o[8];
        unpack8(u, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float z = fmaf(xv[j], sc[j], sh[j]);
            o[j] = silu ? z * fast_sigmoid(z) : z;
        }
        store8(dst, o);
    };
    int r = row0 + rsub;
    for (; r + 3 * rows_par < row1; r += 4 * rows_par) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = ldg16(xp + (long long)r * C + k * rstep);
#pragma unroll
        for (int k = 0; k < 4; ++k) emit(u[k], yp + (long long)r * C + k * rstep);
    }
    for (; r < row1; r += rows_par) emit(ldg16(xp + (long long)r * C), yp + (long long)r * C);
}

// backward apply: dx (+)= rstd * (dz*gamma - s1/m - xhat * s2/m)
__global__ void __launch_bounds__(512)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                    const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ stats,
                    const double* __restrict__ ws, __nv_bfloat16* __restrict__ dx, int HW, int C, int G, int silu,
                    int accumulate, int rows_per_cta) {
    const int n = blockIdx.y;
    const int chunks = C / 8;
    const int cpg = C / G;
    const int rows_par = blockDim.x / chunks;
    const int chunk = threadIdx.x % chunks;
    const int rsub = threadIdx.x / chunks;
    if (rsub >= rows_par) return;
    const int row0 = blockIdx.x * rows_per_cta;
    const int row1 = min(HW, row0 + rows_per_cta);
    const int c0 = chunk * 8;
    const float inv_m = 1.f / ((float)HW * cpg);
    // dx = dz*A - K1 - x*K2 with A = rs*gamma, K2 = rs^2*s2/m, K1 = rs*s1/m - mu*K2;  z = x*A + Bz, Bz = beta - mu*A
    float A[8], Bz[8], K1[8], K2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int g = (c0 + j) / cpg;
        const float mu = stats[(n * G + g) * 2];
        const float rs = stats[(n * G + g) * 2 + 1];
        A[j] = rs * gamma[c0 + j];
        Bz[j] = fmaf(-mu, A[j], beta[c0 + j]);
        K2[j] = rs * rs * (float)ws[(n * G + g) * 2 + 1] * inv_m;
        K1[j] = fmaf(-mu, K2[j], rs * (float)ws[(n * G + g) * 2] * inv_m);
    }
    const __nv_bfloat16* xp = x + (long long)n * HW * C + c0;
    const __nv_bfloat16* dp = dy + (long long)n * HW * C + c0;
    __nv_bfloat16* op = dx + (long long)n * HW * C + c0;
    const long long rstep = (long long)rows_par * C;
    auto emit = [&](const uint4& ux, const uint4& ud, __nv_bfloat16* dst) {
        float xv[8], dv[8], o[8];
        unpack8(ux, xv);
        unpack8(ud, dv);
        if (accumulate) load8(dst, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float dz = dv[j];
            if (silu) {
                const float z = fmaf(xv[j], A[j], Bz[j]);
                const float s = fast_sigmoid(z);
                dz *= s * (1.f + z * (1.f - s));
            }
            const float d = fmaf(dz, A[j], -fmaf(xv[j], K2[j], K1[j]));
            o[j] = accumulate ? o[j] + d : d;
        }
        store8(dst, o);
    };
    int r = row0 + rsub;
    for (; r + 3 * rows_par < row1; r += 4 * rows_par) {
        uint4 ux[4], ud[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ux[k] = ldg16(xp + (long long)r * C + k * rstep);
            ud[k] = ldg16(dp + (long long)r * C + k * rstep);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) emit(ux[k], ud[k], op + (long long)r * C + k * rstep);
    }
    for (; r < row1; r += rows_par)
        emit(ldg16(xp + (long long)r * C), ldg16(dp + (long long)r * C), op + (long long)r * C);
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row (C <= 2560) lives in registers.
static constexpr int LN_MAX_IT = 10;  // 10 * 32 lanes * 8 = 2560 channels

template <int LN_IT>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              __nv_bfloat16* __restrict__ y, float* __restrict__ stats, int T, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= T) return;
    const int chunks = C / 8;
    float v[LN_IT][8];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < LN_IT; ++it) {
        const int ch = it * 32 + lane;
        if (ch < chunks) {
            load8(x + (long long)row * C + ch * 8, v[it]);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[it][j];
        }
    }
    const float mu = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < LN_IT; ++it) {
        const int ch = it * 32 + lane;
        if (ch < chunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = v[it][j] - mu; q += d * d; }
        }
    }
    const float rs = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
    for (int it = 0; it < LN_IT; ++it) {
        const int ch = it * 32 + lane;
        if (ch < chunks) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[it][j] - mu) * rs * gamma[ch * 8 + j] + beta[ch * 8 + j];
            store8(y + (long long)row * C + ch * 8, o);
        }
    }
    if (lane == 0 && stats != nullptr) { stats[row * 2] = mu; stats[row * 2 + 1] = rs; }
}

template <int LN_IT>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
              const float* __restrict__ gamma, const float* __restrict__ stats, __nv_bfloat16* __restrict__ dx, int T,
              int C, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= T) return;
    const int chunks = C / 8;
    const float mu = stats[row * 2], rs = stats[row * 2 + 1];
    float xh[LN_IT][8], dg[LN_IT][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < LN_IT; ++it) {
        const int ch = it * 32 + lane;
        if (ch < chunks) {
            float xv[8], dv[8];
            load8(x + (long long)row * C + ch * 8, xv);
            load8(dy + (long long)row * C + ch * 8, dv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xh[it][j] = (xv[j] - mu) * rs;
                dg[it][j] = dv[j] * gamma[ch * 8 + j];
                s1 += dg[it][j];
                s2 += dg[it][j] * xh[it][j];
            }
        }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
#pragma unroll
    for (int it = 0; it < LN_IT; ++it) {
        const int ch = it * 32 + lane;
        if (ch < chunks) {
            float o[8];
            if (accumulate) load8(dx + (long long)row * C + ch * 8, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = rs * (dg[it][j] - s1 - xh[it][j] * s2);
                o[j] = accumulate ? o[j] + d : d;
            }
            store8(dx + (long long)row * C + ch * 8, o);
        }
    }
}

}  // namespace clb

using namespace clb;

static int gn_launch_geometry(int HW, int C, int n, int& threads, int& rows_per_cta, int& grid_x, int ctas_per_sm = 2) {
    const int chunks = C / 8;
    if (C % 8 != 0 || chunks > GN_MAX_CHUNKS) return set_error(CL_ERR_UNSUPPORTED, "groupnorm: C must be a multiple of 8 and <= 4096");
    int rows_par = 512 / chunks;
    if (rows_par < 1) rows_par = 1;
    threads = ((rows_par * chunks + 31) / 32) * 32;
    if (threads > 512) { rows_par = 1; threads = ((chunks + 31) / 32) * 32; }
    // one wave of ctas_per_sm CTAs per SM across the batch
    int target_ctas = (num_sms() * ctas_per_sm + n - 1) / n;
    rows_per_cta = (HW + target_ctas - 1) / target_ctas;
    if (rows_per_cta < rows_par) rows_per_cta = rows_par;
    grid_x = (HW + rows_per_cta - 1) / rows_per_cta;
    return CL_OK;
}

extern "C" int cl_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats,
                                double* ws, int n, int HW, int C, int G, float eps, int silu, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!x || !gamma || !beta || !y || !ws) return set_error(CL_ERR_INVALID, "cl_groupnorm_fwd: null pointer");
    if (G <= 0 || C % G != 0) return set_error(CL_ERR_INVALID, "cl_groupnorm_fwd: C %% G != 0");
    int threads, rows_per_cta, grid_x;
    CL_CHECK(gn_launch_geometry(HW, C, n, threads, rows_per_cta, grid_x));
    CL_CUDA_CHECK(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * n * G, stream));
    const size_t rsmem = (size_t)2 * (threads / (C / 8) + 1) * C * sizeof(float);
    {
        static bool done = false;
        if (!done) {
            CL_CUDA_CHECK(cudaFuncSetAttribute(gn_reduce_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            done = true;
        }
    }
    gn_reduce_kernel<0><<<dim3(grid_x, n), threads, rsmem, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), nullptr, nullptr, nullptr, nullptr, ws, nullptr, nullptr, HW, C, G,
        rows_per_cta, 0);
    if (stats == nullptr) return set_error(CL_ERR_INVALID, "cl_groupnorm_fwd: stats buffer is required");
    gn_finalize_kernel<<<(n * G + 127) / 128, 128, 0, stream>>>(ws, stats, n * G, 1.0 / ((double)HW * (C / G)), eps);
    int a_threads, a_rows, a_grid;
    CL_CHECK(gn_launch_geometry(HW, C, n, a_threads, a_rows, a_grid, 8));
    gn_apply_kernel<<<dim3(a_grid, n), a_threads, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta,
                                                               stats, reinterpret_cast<__nv_bfloat16*>(y), HW, C, G,
                                                               silu, a_rows);
    count_launch(3);
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}

extern "C" int cl_groupnorm_bwd(const void* x, const void* dy, const float* gamma, const float* beta,
                                const float* stats, void* dx, float* dgamma, float* dbeta, double* ws, int n, int HW,
                                int C, int G, int silu, int accumulate, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!x || !dy || !gamma || !beta || !stats || !dx || !ws) return set_error(CL_ERR_INVALID, "cl_groupnorm_bwd: null pointer");
    if (G <= 0 || C % G != 0) return set_error(CL_ERR_INVALID, "cl_groupnorm_bwd: C %% G != 0");
    int threads, rows_per_cta, grid_x;
    CL_CHECK(gn_launch_geometry(HW, C, n, threads, rows_per_cta, grid_x));
    CL_CUDA_CHECK(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * n * G, stream));
    const size_t rsmem = (size_t)2 * (threads / (C / 8) + 1) * C * sizeof(float);
    {
        static bool done = false;
        if (!done) {
            CL_CUDA_CHECK(cudaFuncSetAttribute(gn_reduce_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            done = true;
        }
    }
    gn_reduce_kernel<1><<<dim3(grid_x, n), threads, rsmem, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), gamma, beta, stats, ws,
        dgamma, dbeta, HW, C, G, rows_per_cta, silu);
    int a_threads, a_rows, a_grid;
    CL_CHECK(gn_launch_geometry(HW, C, n, a_threads, a_rows, a_grid, 8));
    gn_bwd_apply_kernel<<<dim3(a_grid, n), a_threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), gamma, beta, stats, ws,
        reinterpret_cast<__nv_bfloat16*>(dx), HW, C, G, silu, accumulate, a_rows);
    count_launch(2);
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}

extern "C" int cl_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, int T,
                                int C, float eps, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!x || !gamma || !beta || !y) return set_error(CL_ERR_INVALID, "cl_layernorm_fwd: null pointer");
    if (C % 8 != 0 || C > LN_MAX_IT * 256) return set_error(CL_ERR_UNSUPPORTED, "cl_layernorm_fwd: C must be a multiple of 8, <= 2560");
#define LN_FWD(IT) ln_fwd_kernel<IT><<<(T + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), gamma, beta, reinterpret_cast<__nv_bfloat16*>(y), stats, T, C, eps)
    if (C <= 512) LN_FWD(2); else if (C <= 768) LN_FWD(3); else if (C <= 1280) LN_FWD(5); else LN_FWD(10);
#undef LN_FWD
    count_launch();
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}

extern "C" int cl_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* stats, void* dx, int T,
                                int C, int accumulate, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!x || !dy || !gamma || !stats || !dx) return set_error(CL_ERR_INVALID, "cl_layernorm_bwd: null pointer");
    if (C % 8 != 0 || C > LN_MAX_IT * 256) return set_error(CL_ERR_UNSUPPORTED, "cl_layernorm_bwd: C must be a multiple of 8, <= 2560");
#define LN_BWD(IT) ln_bwd_kernel<IT><<<(T + 7) / 8, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), gamma, stats, reinterpret_cast<__nv_bfloat16*>(dx), T, C, accumulate)
    if (C <= 512) LN_BWD(2); else if (C <= 768) LN_BWD(3); else if (C <= 1280) LN_BWD(5); else LN_BWD(10);
#undef LN_BWD
    count_launch();
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}
