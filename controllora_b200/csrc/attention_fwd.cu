// K2 (forward) — fused scaled-dot-product attention on tcgen05 / TMEM for sm_100a.
//
//   O = softmax(Q K^T * scale) V,   per (batch, head);  Q/K/V/O are bf16 token matrices [B, N, H*d] (row stride ld*),
//   head h = columns [h*d, (h+1)*d).  The N x N probability matrix is never written to HBM (the reference
//   materialises it: attn.get_attention_scores + torch.bmm, models.py:140-141, 270-271, 407-408).
//
// One CTA = one (batch, head, 128-query block); 8 warps:
//   warp 0     TMA producer: Q once, then a ring of K/V tiles (4-D tensor maps {d, H, N, B}; the head dim is padded
//              to a multiple of 64 purely by TMA out-of-bounds zero fill — nothing is padded in HBM)
//   warp 1     MMA issuer:   S = Q K^T      (A, B K-major from smem, D in TMEM columns [0, BLOCK_N))
//                            O += P V       (A = P K-major from smem, B = V *MN-major* from smem, D in TMEM)
//   warp 2     TMEM allocator
//   warps 4-7  softmax: thread r owns query row r == TMEM lane r (no cross-thread reductions): online max / sum in
//              the exp2 domain, rescales O in TMEM, writes P (bf16) into the 128B-swizzled K-major smem tile.
// Per-CTA work is serialised S -> softmax -> PV; the kernel is sized so that two CTAs are resident per SM and one
// CTA's softmax (MUFU-bound) overlaps the other's tensor-core work.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#define CLB_FAMILY 2      // CLB_PDL_MASK bit of this file's kernels
#include "host_common.h"
#include "../../include/controllora_b200.h"

namespace clb {

struct AttnFwdParams {
    int B, H, Nq, Nk, d;
    __nv_bfloat16* o;
    long long ldo;
    float* lse;
    float scale_log2;
    int num_q_blocks;
};

template <int DP, int BLOCK_N, int STAGES>
struct AttnFwdCfg {
    static constexpr int BLOCK_M = 128;
    static constexpr int DCH = DP / 64;                       // 64-column chunks of the (padded) head dim
    static constexpr int NCH = BLOCK_N / 64;                  // 64-key chunks of the P tile
    static constexpr int Q_BYTES = DCH * BLOCK_M * 128;
    static constexpr int KV_TILE_BYTES = DCH * BLOCK_N * 128; // one K (or V) tile
    static constexpr int STAGE_BYTES = 2 * KV_TILE_BYTES;
    static constexpr int P_BYTES = NCH * BLOCK_M * 128;
    static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + P_BYTES + 128;
    static constexpr int TMEM_NEED = BLOCK_N + DP;
    static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
    static_assert(DP % 64 == 0 && BLOCK_N % 64 == 0, "tile dims");
    static_assert(DP <= 256 && BLOCK_N <= 256, "UMMA N limit");
};

template <int DP, int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(256, (DP <= 64) ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
    pdl_launch_dependents();
    using Cfg = AttnFwdCfg<DP, BLOCK_N, STAGES>;
    constexpr int BLOCK_M = Cfg::BLOCK_M;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // no static shared memory in this kernel: the dynamic segment starts 1024-byte aligned (checked below)
    uint8_t* smem = smem_raw;
    uint8_t* smem_q = smem;
    uint8_t* smem_kv = smem_q + Cfg::Q_BYTES;
    uint8_t* smem_p = smem_kv + STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + Cfg::P_BYTES);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // STAGES
    uint64_t* kv_empty = kv_full + STAGES;   // STAGES
    uint64_t* s_full = kv_empty + STAGES;    // 1
    uint64_t* p_full = s_full + 1;           // 1
    uint64_t* o_full = p_full + 1;           // 1
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 1);

    const int warp_idx = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const int qb = blockIdx.x % p.num_q_blocks;
    const int bh = blockIdx.x / p.num_q_blocks;
    const int h = bh % p.H;
    const int b = bh / p.H;
    const int q0 = qb * BLOCK_M;
    const int num_kv = (p.Nk + BLOCK_N - 1) / BLOCK_N;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) {
        printf("attn_fwd: dynamic smem base not 1024-aligned\n");
        __trap();
    }
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) {
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    pdl_wait();   // prologue above touched only smem / TMEM / kernel params
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + BLOCK_N;

    if (warp_idx == 0) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            mbar_arrive_expect_tx_e(q_full, Cfg::Q_BYTES);
            for (int c = 0; c < Cfg::DCH; ++c)
                tma_load_4d_e(&tmQ, q_full, smem_q + c * BLOCK_M * 128, c * 64, h, q0, b);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < num_kv; ++i) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx_e(&kv_full[stage], Cfg::STAGE_BYTES);
                uint8_t* sk = smem_kv + stage * Cfg::STAGE_BYTES;
                uint8_t* sv = sk + Cfg::KV_TILE_BYTES;
                for (int c = 0; c < Cfg::DCH; ++c) {
                    tma_load_4d_e(&tmK, &kv_full[stage], sk + c * BLOCK_N * 128, c * 64, h, i * BLOCK_N, b);
                    tma_load_4d_e(&tmV, &kv_full[stage], sv + c * BLOCK_N * 128, c * 64, h, i * BLOCK_N, b);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            constexpr uint32_t idesc_s = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(BLOCK_M, DP, 0, 1);  // B = V is MN-major
            mbar_wait(q_full, 0);
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t sq = smem_u32(smem_q);
            const uint32_t sp = smem_u32(smem_p);
            for (int i = 0; i < num_kv; ++i) {
                mbar_wait(&kv_full[stage], phase);
                tc_fence_after();
                const uint32_t sk = smem_u32(smem_kv + stage * Cfg::STAGE_BYTES);
                const uint32_t sv = sk + Cfg::KV_TILE_BYTES;
                // S = Q K^T  (K loop over the padded head dim)
#pragma unroll
                for (int kk = 0; kk < DP / 16; ++kk) {
                    const uint32_t off = (kk / 4) * (BLOCK_M * 128) + (kk % 4) * 32;
                    const uint32_t offk = (kk / 4) * (BLOCK_N * 128) + (kk % 4) * 32;
                    tc_mma_ss_e(tmem_s, make_smem_desc(sq + off, 16, 1024, 2), make_smem_desc(sk + offk, 16, 1024, 2),
                              idesc_s, kk != 0 ? 1u : 0u);
                }
                tc_commit_e(s_full);
                // O += P V  (K loop over the keys of this block)
                mbar_wait(p_full, i & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < BLOCK_N / 16; ++kk) {
                    const uint32_t offp = (kk / 4) * (BLOCK_M * 128) + (kk % 4) * 32;
                    const uint64_t adesc = make_smem_desc(sp + offp, 16, 1024, 2);
                    // V tile: [keys][64 d] rows of 128 B; MN-major: LBO = next 64-d chunk, SBO = next 8 keys
                    const uint64_t bdesc = make_smem_desc(sv + kk * 2048, BLOCK_N * 128, 1024, 2);
                    tc_mma_ss_e(tmem_o, adesc, bdesc, idesc_o, (i | kk) != 0 ? 1u : 0u);
                }
                tc_commit_e(&kv_empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit_e(o_full);
        }
    } else if (warp_idx >= 4) {
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane;                 // query row within the block == TMEM lane
        const uint32_t lane_off = uint32_t(quad * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        for (int i = 0; i < num_kv; ++i) {
            mbar_wait(s_full, i & 1);
            tc_fence_after();
            const int kbase = i * BLOCK_N;
            // ---- pass 1: row max
            float mloc = -INFINITY;
            const bool tail = kbase + BLOCK_N > p.Nk;   // only the last key block needs column masking
#pragma unroll
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_s + lane_off + c * 32, v);
                tc_wait_ld();
                if (!tail) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mloc = fmaxf(mloc, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float s = (kbase + c * 32 + j < p.Nk) ? __uint_as_float(v[j]) : -INFINITY;
                        mloc = fmaxf(mloc, s);
                    }
                }
            }
            const float m_new = fmaxf(m_run, mloc * p.scale_log2);
            const float alpha = fast_exp2(m_run - m_new);   // first block: exp2(-inf) = 0
            l_run *= alpha;
            m_run = m_new;
            // ---- rescale the O accumulator (S_i was issued after PV_{i-1}: s_full implies PV_{i-1} retired);
            //      skipped when no row of this warp raised its max (warp-uniform: tcgen05.ld/st are .aligned)
            if (i > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll
                for (int c = 0; c < DP / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_o + lane_off + c * 32, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(
                            tmem_o + lane_off + c * 32),
                        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
                        "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
                        "r"(v[30]), "r"(v[31])
                        : "memory");
                }
                tc_wait_st();
            }
            // ---- pass 2: P = exp2(S*scale - m), row sum, bf16 -> swizzled K-major smem tile
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_s + lane_off + c * 32, v);
                tc_wait_ld();
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(v[j]), p.scale_log2, -m_new));
                    float p1 = fast_exp2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -m_new));
                    if (tail) {
                        if (kbase + c * 32 + j >= p.Nk) p0 = 0.f;
                        if (kbase + c * 32 + j + 1 >= p.Nk) p1 = 0.f;
                    }
                    lsum += p0 + p1;
                    pk[j >> 1] = pack_bf16x2(p0, p1);
                }
                // 32 keys = 64 B = four 16-byte chunks of this row
                uint8_t* rowp = smem_p + (c / 2) * (BLOCK_M * 128) + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (c & 1) * 4 + q;
                    *reinterpret_cast<uint4*>(rowp + ((chunk ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l_run += lsum;
            tc_fence_before();
            fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // ---- epilogue: O / l -> bf16 -> global; LSE (log2 domain) for the backward pass
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int q = q0 + r;
        const float inv_l = 1.f / l_run;
        __nv_bfloat16* orow = p.o + ((long long)b * p.Nq + q) * p.ldo + h * p.d;
#pragma unroll
        for (int c = 0; c < DP / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_o + lane_off + c * 32, v);
            tc_wait_ld();
            if (q < p.Nq) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const int col = c * 32 + j;
                    if (col < p.d) {   // d % 8 == 0
                        uint4 o4;
                        o4.x = pack_bf16x2(__uint_as_float(v[j]) * inv_l, __uint_as_float(v[j + 1]) * inv_l);
                        o4.y = pack_bf16x2(__uint_as_float(v[j + 2]) * inv_l, __uint_as_float(v[j + 3]) * inv_l);
                        o4.z = pack_bf16x2(__uint_as_float(v[j + 4]) * inv_l, __uint_as_float(v[j + 5]) * inv_l);
                        o4.w = pack_bf16x2(__uint_as_float(v[j + 6]) * inv_l, __uint_as_float(v[j + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + col) = o4;
                    }
                }
            }
        }
        if (q < p.Nq && p.lse != nullptr) p.lse[((long long)b * p.H + h) * p.Nq + q] = m_run + log2f(l_run);
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Paired-tile variant for head dims <= 64 (every SD-1.5 attention: d = 40 / 80 / 160 -> this one serves d = 40).
//
// The single-tile kernel above reads every S tile from TMEM twice (row max, then exp) and TMEM reads run at
// ~16 B/clk per SM sub-partition: 2 x 128 x 128 fp32 = 2048 clk per tile, which is what it measures (~2.4k clk/tile).
// Here one CTA owns TWO 128-query tiles (FlashAttention-4 style ping-pong), one CTA per SM with a large register
// budget: a softmax thread keeps its whole 128-key S row in registers, so S is read from TMEM exactly once, and the
// tensor core works on tile 1 (PV, next S) while tile 0 is in its exp phase and vice versa.  The O accumulator is only
// rescaled when a row max grows by more than 2^8 (the stale max is used otherwise; the final O / l normalisation is
// exact either way), which removes nearly all O read-modify-write traffic from TMEM.
//   warp 0 TMA, warp 1 MMA issue, warp 2 TMEM alloc, warps 4-7 softmax of tile 0, warps 8-11 softmax of tile 1.
//   TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384).
struct AttnFwd2Cfg {
    static constexpr int BLOCK_M = 128, BLOCK_N = 128, DP = 64, STAGES = 3, THREADS = 384;
    static constexpr int Q_BYTES = 2 * BLOCK_M * 128;          // two query tiles
    static constexpr int KV_TILE_BYTES = BLOCK_N * 128;
    static constexpr int STAGE_BYTES = 2 * KV_TILE_BYTES;
    static constexpr int P_TILE_BYTES = 2 * BLOCK_M * 128;     // 128 keys = two 64-key chunks
    static constexpr int SMEM_BYTES = Q_BYTES + STAGES * STAGE_BYTES + 2 * P_TILE_BYTES + 128;
    static constexpr int TMEM_COLS = 512;
    static constexpr float RESCALE_THRESHOLD = 8.f;            // log2 domain
};

__global__ void __launch_bounds__(AttnFwd2Cfg::THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
    pdl_launch_dependents();
    using Cfg = AttnFwd2Cfg;
    constexpr int BLOCK_M = Cfg::BLOCK_M, BLOCK_N = Cfg::BLOCK_N, DP = Cfg::DP, STAGES = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem_q = smem_raw;
    uint8_t* smem_kv = smem_q + Cfg::Q_BYTES;
    uint8_t* smem_p = smem_kv + STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + 2 * Cfg::P_TILE_BYTES);
    uint64_t* q_full = bars;                 // 1
    uint64_t* kv_full = bars + 1;            // STAGES
    uint64_t* kv_empty = kv_full + STAGES;   // STAGES
    uint64_t* s_full = kv_empty + STAGES;    // 2
    uint64_t* p_full = s_full + 2;           // 2
    uint64_t* o_full = p_full + 2;           // 1
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 1);

    const int warp_idx = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const int qb = blockIdx.x % p.num_q_blocks;   // 256-query blocks
    const int bh = blockIdx.x / p.num_q_blocks;
    const int h = bh % p.H;
    const int b = bh / p.H;
    const int q0 = qb * 2 * BLOCK_M;
    const int num_kv = (p.Nk + BLOCK_N - 1) / BLOCK_N;

    if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) {
        printf("attn_fwd2: dynamic smem base not 1024-aligned\n");
        __trap();
    }
    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[t], 4);
        }
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) {
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    pdl_wait();   // prologue above touched only smem / TMEM / kernel params
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    // 384 threads x 168 registers at launch; the softmax warpgroups hold a 128-column S row per thread
    if (warp_idx == 0) {
        setmaxnreg_dec<88>();
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            mbar_arrive_expect_tx_e(q_full, Cfg::Q_BYTES);
            tma_load_4d_e(&tmQ, q_full, smem_q, 0, h, q0, b);
            tma_load_4d_e(&tmQ, q_full, smem_q + BLOCK_M * 128, 0, h, q0 + BLOCK_M, b);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < num_kv; ++i) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx_e(&kv_full[stage], Cfg::STAGE_BYTES);
                uint8_t* sk = smem_kv + stage * Cfg::STAGE_BYTES;
                tma_load_4d_e(&tmK, &kv_full[stage], sk, 0, h, i * BLOCK_N, b);
                tma_load_4d_e(&tmV, &kv_full[stage], sk + Cfg::KV_TILE_BYTES, 0, h, i * BLOCK_N, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        setmaxnreg_dec<88>();
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            constexpr uint32_t idesc_s = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(BLOCK_M, DP, 0, 1);  // B = V is MN-major
            const uint32_t sq = smem_u32(smem_q);
            const uint32_t sp = smem_u32(smem_p);
            auto issue_s = [&](int t, uint32_t sk) {
#pragma unroll
                for (int kk = 0; kk < DP / 16; ++kk)
                    tc_mma_ss_e(tmem_base + t * BLOCK_N, make_smem_desc(sq + t * (BLOCK_M * 128) + kk * 32, 16, 1024, 2),
                              make_smem_desc(sk + kk * 32, 16, 1024, 2), idesc_s, kk != 0 ? 1u : 0u);
                tc_commit_e(&s_full[t]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0, smem_u32(smem_kv));
            issue_s(1, smem_u32(smem_kv));
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < num_kv; ++i) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == STAGES) { nstage = 0; nphase ^= 1; }
                const bool more = i + 1 < num_kv;
                if (more) {
                    mbar_wait(&kv_full[nstage], nphase);
                    tc_fence_after();
                }
                const uint32_t sv = smem_u32(smem_kv + stage * Cfg::STAGE_BYTES) + Cfg::KV_TILE_BYTES;
                const uint32_t sk_next = smem_u32(smem_kv + nstage * Cfg::STAGE_BYTES);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    mbar_wait(&p_full[t], i & 1);
                    tc_fence_after();
                    // O_t += P_t V_i
#pragma unroll
                    for (int kk = 0; kk < BLOCK_N / 16; ++kk) {
                        const uint32_t offp = t * Cfg::P_TILE_BYTES + (kk / 4) * (BLOCK_M * 128) + (kk % 4) * 32;
                        tc_mma_ss_e(tmem_base + 2 * BLOCK_N + t * DP, make_smem_desc(sp + offp, 16, 1024, 2),
                                  make_smem_desc(sv + kk * 2048, BLOCK_N * 128, 1024, 2), idesc_o,
                                  (i | kk) != 0 ? 1u : 0u);
                    }
                    // S_t of the next key block: its softmax starts while the other tile's PV runs
                    if (more) issue_s(t, sk_next);
                }
                tc_commit_e(&kv_empty[stage]);
                stage = nstage;
                phase = nphase;
            }
            tc_commit_e(o_full);
        }
    } else if (warp_idx < 4) {
        setmaxnreg_dec<88>();
    } else {
        setmaxnreg_inc<208>();
        const int t = (warp_idx - 4) >> 2;              // query tile of this warp
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane;                 // query row within the tile == TMEM lane
        const uint32_t lane_off = uint32_t(quad * 32) << 16;
        const uint32_t tmem_s = tmem_base + t * BLOCK_N + lane_off;
        const uint32_t tmem_o = tmem_base + 2 * BLOCK_N + t * DP + lane_off;
        uint8_t* my_p = smem_p + t * Cfg::P_TILE_BYTES;
        float m_use = -INFINITY, l_run = 0.f;           // m_use: the (possibly stale) max the exponentials refer to
        for (int i = 0; i < num_kv; ++i) {
            mbar_wait(&s_full[t], i & 1);
            tc_fence_after();
            const int kbase = i * BLOCK_N;
            const bool tail = kbase + BLOCK_N > p.Nk;   // only the last key block needs column masking
            uint32_t s[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_s + c * 32, s[c]);
            tc_wait_ld();
            float mloc = -INFINITY;
            if (!tail) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) mloc = fmaxf(mloc, __uint_as_float(s[c][j]));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (kbase + c * 32 + j < p.Nk) mloc = fmaxf(mloc, __uint_as_float(s[c][j]));
            }
            const float m_new = fmaxf(m_use, mloc * p.scale_log2);
            const bool grow = m_new > m_use + Cfg::RESCALE_THRESHOLD;   // first block: m_use = -inf -> true
            float alpha = 1.f;
            if (grow) {
                alpha = fast_exp2(m_use - m_new);       // first block: exp2(-inf) = 0
                l_run *= alpha;
                m_use = m_new;
            }
            // rescale O (s_full(i) implies PV(i-1) retired); warp-uniform decision: tcgen05.ld/st are .aligned
            if (i > 0 && __any_sync(0xffffffffu, grow)) {
#pragma unroll
                for (int c = 0; c < DP / 16; ++c) {   // 16-column pieces: the S row (128 registers) is live here
                    uint32_t v[16];
                    tmem_ld_32x16(tmem_o + c * 16, v);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
                    tmem_st_32x16(tmem_o + c * 16, v);
                }
                tc_wait_st();
            }
            // P = exp2(S*scale - m_use) from registers, row sum, bf16 -> swizzled K-major smem tile
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(s[c][j]), p.scale_log2, -m_use));
                    float p1 = fast_exp2(fmaf(__uint_as_float(s[c][j + 1]), p.scale_log2, -m_use));
                    if (tail) {
                        if (kbase + c * 32 + j >= p.Nk) p0 = 0.f;
                        if (kbase + c * 32 + j + 1 >= p.Nk) p1 = 0.f;
                    }
                    lsum += p0 + p1;
                    pk[j >> 1] = pack_bf16x2(p0, p1);
                }
                uint8_t* rowp = my_p + (c / 2) * (BLOCK_M * 128) + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (c & 1) * 4 + q;
                    *reinterpret_cast<uint4*>(rowp + ((chunk ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l_run += lsum;
            tc_fence_before();
            fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
        }
        // ---- epilogue: O / l -> bf16 -> global; LSE (log2 domain) for the backward pass
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int q = q0 + t * BLOCK_M + r;
        const float inv_l = 1.f / l_run;
        __nv_bfloat16* orow = p.o + ((long long)b * p.Nq + q) * p.ldo + h * p.d;
#pragma unroll
        for (int c = 0; c < DP / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_o + c * 32, v);
            tc_wait_ld();
            if (q < p.Nq) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const int col = c * 32 + j;
                    if (col < p.d) {   // d % 8 == 0
                        uint4 o4;
                        o4.x = pack_bf16x2(__uint_as_float(v[j]) * inv_l, __uint_as_float(v[j + 1]) * inv_l);
                        o4.y = pack_bf16x2(__uint_as_float(v[j + 2]) * inv_l, __uint_as_float(v[j + 3]) * inv_l);
                        o4.z = pack_bf16x2(__uint_as_float(v[j + 4]) * inv_l, __uint_as_float(v[j + 5]) * inv_l);
                        o4.w = pack_bf16x2(__uint_as_float(v[j + 6]) * inv_l, __uint_as_float(v[j + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + col) = o4;
                    }
                }
            }
        }
        if (q < p.Nq && p.lse != nullptr) p.lse[((long long)b * p.H + h) * p.Nq + q] = m_use + log2f(l_run);
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

int make_head_map(CUtensorMap* tm, const void* ptr, int B, int H, int N, int d, long long ld, int box_rows) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)H, (uint64_t)N, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)N * ld * 2};
    uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    return get_tensor_map(tm, ptr, 4, dims, strides, box, 128);
}

template <int DP, int BLOCK_N, int STAGES>
static int launch_attn_fwd(const cl_attn_fwd_args* a, cudaStream_t stream) {
    using Cfg = AttnFwdCfg<DP, BLOCK_N, STAGES>;
    CUtensorMap tq, tk, tv;
    CL_CHECK(make_head_map(&tq, a->q, a->B, a->H, a->Nq, a->d, a->ldq, 128));
    CL_CHECK(make_head_map(&tk, a->k, a->B, a->H, a->Nk, a->d, a->ldk, BLOCK_N));
    CL_CHECK(make_head_map(&tv, a->v, a->B, a->H, a->Nk, a->d, a->ldv, BLOCK_N));
    AttnFwdParams p;
    p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.d = a->d;
    p.o = reinterpret_cast<__nv_bfloat16*>(a->o); p.ldo = a->ldo; p.lse = a->lse;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    p.num_q_blocks = (a->Nq + 127) / 128;
    static bool attr_done = false;
    if (!attr_done) {
        CL_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<DP, BLOCK_N, STAGES>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_done = true;
    }
    const int grid = a->B * a->H * p.num_q_blocks;
    launch_k(attn_fwd_kernel<DP, BLOCK_N, STAGES>, grid, 256, Cfg::SMEM_BYTES, stream, tq, tk, tv, p);
    count_launch();
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}

static int launch_attn_fwd2(const cl_attn_fwd_args* a, cudaStream_t stream) {
    using Cfg = AttnFwd2Cfg;
    CUtensorMap tq, tk, tv;
    CL_CHECK(make_head_map(&tq, a->q, a->B, a->H, a->Nq, a->d, a->ldq, 128));
    CL_CHECK(make_head_map(&tk, a->k, a->B, a->H, a->Nk, a->d, a->ldk, Cfg::BLOCK_N));
    CL_CHECK(make_head_map(&tv, a->v, a->B, a->H, a->Nk, a->d, a->ldv, Cfg::BLOCK_N));
    AttnFwdParams p;
    p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.d = a->d;
    p.o = reinterpret_cast<__nv_bfloat16*>(a->o); p.ldo = a->ldo; p.lse = a->lse;
    p.scale_log2 = a->scale * 1.4426950408889634f;
    p.num_q_blocks = (a->Nq + 255) / 256;
    static bool attr_done = false;
    if (!attr_done) {
        CL_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_done = true;
    }
    const int grid = a->B * a->H * p.num_q_blocks;
    launch_k(attn_fwd2_kernel, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, tq, tk, tv, p);
    count_launch();
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}

}  // namespace clb

using namespace clb;

extern "C" int cl_attn_fwd(const cl_attn_fwd_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->q || !a->k || !a->v || !a->o) return set_error(CL_ERR_INVALID, "cl_attn_fwd: null pointer");
    if (a->d % 8 != 0 || a->d <= 0 || a->d > 192) return set_error(CL_ERR_UNSUPPORTED, "cl_attn_fwd: head dim must be a multiple of 8, <= 192");
    if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 8)) return set_error(CL_ERR_INVALID, "cl_attn_fwd: row strides must be multiples of 8");
    if (a->B <= 0 || a->H <= 0 || a->Nq <= 0 || a->Nk <= 0) return set_error(CL_ERR_INVALID, "cl_attn_fwd: dims");
    if (a->d <= 64 && a->Nq > 128) return launch_attn_fwd2(a, stream);   // paired query tiles, S read from TMEM once
    if (a->d <= 64) return launch_attn_fwd<64, 128, 2>(a, stream);
    if (a->d <= 128) return launch_attn_fwd<128, 128, 2>(a, stream);
    return launch_attn_fwd<192, 64, 2>(a, stream);
}
