// K2 (backward) — attention gradients on tcgen05 / TMEM, recomputing the probabilities from Q, K and the saved
// log-sum-exp (the reference's autograd instead keeps the materialised N x N probabilities of every layer).
//
//   P = exp2(S*scale*log2e - LSE),  dP = dO V^T,  dS = P o (dP - delta) * scale,  delta = rowsum(dO o O)
//   dV = P^T dO,   dK = dS^T Q,   dQ = dS K
//
// Two kernels, no atomics:
//   attn_bwd_dkv: CTA = (batch, head, 128-key block), loops over query blocks.  Works on the transposed problem so
//       that TMEM lane == key row:   S^T = K Q^T,  dP^T = V dO^T  (both operands K-major smem tiles);  the softmax
//       warps turn them into bf16 P^T and dS^T and write them back INTO TENSOR MEMORY (tcgen05.st, packed two bf16 per
//       32-bit column, over the S^T / dP^T columns each warp has already consumed);  dV += P^T dO and dK += dS^T Q then
//       take their A operand from TMEM (tcgen05.mma with a TMEM A operand) and the *same* dO / Q smem tiles again as
//       MN-major B operands.
//   attn_bwd_dq:  CTA = (batch, head, 128-query block), loops over key blocks:  S = Q K^T, dP = dO V^T,
//       dS -> TMEM,  dQ += dS K (dS as TMEM A operand, K tile as MN-major B operand).
// Why TMEM operands: ncu (profiles/r02_attn_bwd_ncu.md) showed both kernels bound by SHARED-MEMORY bandwidth - the
// P^T / dS^T tiles cost 16-byte st.shared wavefronts on the way in and 16 KB of operand fetch per MMA group on the way
// out (dkv: 116 M LSU wavefronts + 96 KB of MMA operand reads per query block against 128 B/clk/SM).  Through TMEM they
// cost neither.
// Warp roles as in the forward kernel: warp 0 TMA, warp 1 MMA issue, warp 2 TMEM alloc, warps 4-7 softmax/epilogue.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#define CLB_FAMILY 2      // CLB_PDL_MASK bit of this file's kernels
#include "host_common.h"
#include "../../include/controllora_b200.h"

namespace clb {

int make_head_map(CUtensorMap* tm, const void* ptr, int B, int H, int N, int d, long long ld, int box_rows);

struct AttnBwdParams {
    int B, H, Nq, Nk, d;
    const float* lse;     // [B, H, Nq] log2-domain
    const float* delta;   // [B, H, Nq]
    __nv_bfloat16* dq; long long lddq;
    __nv_bfloat16* dk; long long lddk;
    __nv_bfloat16* dv; long long lddv;
    float scale, scale_log2;
    int num_blocks;       // key blocks (dkv) or query blocks (dq) per (b, h)
    int stats_all;        // dkv: LSE / delta of ALL query blocks are staged in shared memory once (no per-block CTA barrier)
};

// write 32 consecutive bf16 values (columns [c32*32, c32*32+32) of row r) into a K-major 128B-swizzled tile whose
// 64-column chunks are `chunk_stride` bytes apart
__device__ __forceinline__ void store_row32_swz(uint8_t* tile, int chunk_stride, int r, int c32, const uint32_t (&pk)[16]) {
    uint8_t* rowp = tile + (c32 >> 1) * chunk_stride + r * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int chunk = (c32 & 1) * 4 + q;
        *reinterpret_cast<uint4*>(rowp + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
}

// same for 16 consecutive columns [c16*16, c16*16+16)
__device__ __forceinline__ void store_row16_swz(uint8_t* tile, int chunk_stride, int r, int c16, const uint32_t (&pk)[8]) {
    uint8_t* rowp = tile + (c16 >> 2) * chunk_stride + r * 128;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int chunk = (c16 & 3) * 2 + q;
        *reinterpret_cast<uint4*>(rowp + ((chunk ^ (r & 7)) << 4)) =
            make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
}

// TMEM accumulator rows -> bf16 global rows (thread = row), columns [0, d)
template <int DP>
__device__ __forceinline__ void store_acc_rows(uint32_t taddr, __nv_bfloat16* rowptr, bool row_ok, int d, float mul,
                                               int c_begin = 0, int c_step = 1) {
#pragma unroll
    for (int c = 0; c < DP / 32; ++c) {
        if (c < c_begin || ((c - c_begin) % c_step) != 0) continue;   // (warp-uniform) this warp's share of the columns
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tc_wait_ld();
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const int col = c * 32 + j;
                if (col < d) {
                    uint4 o4;
                    o4.x = pack_bf16x2(__uint_as_float(v[j]) * mul, __uint_as_float(v[j + 1]) * mul);
                    o4.y = pack_bf16x2(__uint_as_float(v[j + 2]) * mul, __uint_as_float(v[j + 3]) * mul);
                    o4.z = pack_bf16x2(__uint_as_float(v[j + 4]) * mul, __uint_as_float(v[j + 5]) * mul);
                    o4.w = pack_bf16x2(__uint_as_float(v[j + 6]) * mul, __uint_as_float(v[j + 7]) * mul);
                    *reinterpret_cast<uint4*>(rowptr + col) = o4;
                }
            }
        }
    }
}

// ===================================================================================================== dK, dV
template <int DP, int BQ, int STAGES>
struct DkvCfg {
    static constexpr int BK = 128;                              // keys per CTA == TMEM lanes
    static constexpr int DCH = DP / 64;
    static constexpr int KV_BYTES = DCH * BK * 128;             // K tile (and V tile)
    static constexpr int QD_TILE = DCH * BQ * 128;              // Q_j tile (and dO_j tile)
    static constexpr int STAGE_BYTES = 2 * QD_TILE;
    static constexpr int SMEM_BYTES = 1024 + 2 * KV_BYTES + STAGES * STAGE_BYTES + 2 * 2 * BQ * 4 + 256;
    // P^T / dS^T live in TMEM over the S^T / dP^T columns: the softmax warp of column half h owns S^T columns
    // [h*BQ/2, (h+1)*BQ/2) and packs the bf16 result of its chunk c (16 queries -> 8 columns) at the start of that range
    static constexpr int CH_PER_HALF = BQ / 32;                 // 16-query chunks per softmax warp
    __host__ __device__ static constexpr int a_col(int chunk) { return (chunk / CH_PER_HALF) * (BQ / 2) + (chunk % CH_PER_HALF) * 8; }
    static constexpr int TM_ST = 0, TM_DPT = BQ, TM_DK = 2 * BQ, TM_DV = 2 * BQ + DP;
    static constexpr int TMEM_NEED = 2 * BQ + 2 * DP;
    static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
    static constexpr int MIN_CTAS = (TMEM_COLS <= 256 && SMEM_BYTES <= 110 * 1024) ? 2 : 1;
    static_assert(2 * BQ + 2 * DP <= 512, "TMEM");
    static_assert(BQ % 64 == 0, "BQ");
};

// 12 warps: 0 TMA, 1 MMA, 2 TMEM alloc, 4-11 softmax (two warps per TMEM lane quadrant, each takes half of the columns:
// with one softmax warp per SM sub-partition the dependent tcgen05.ld -> exp -> st.shared chains left the issue slots idle)
static constexpr int BWD_THREADS = 384;

template <int DP, int BQ, int STAGES>
__global__ void __launch_bounds__(BWD_THREADS, (DkvCfg<DP, BQ, STAGES>::MIN_CTAS))
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                    const AttnBwdParams p) {
    pdl_launch_dependents();
    using Cfg = DkvCfg<DP, BQ, STAGES>;
    constexpr int BK = Cfg::BK;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_k = smem;
    uint8_t* smem_v = smem_k + Cfg::KV_BYTES;
    uint8_t* smem_st = smem_v + Cfg::KV_BYTES;                       // stages: [Q_j | dO_j]
    float* smem_lse = reinterpret_cast<float*>(smem_st + STAGES * Cfg::STAGE_BYTES);   // [2][BQ]
    float* smem_delta = smem_lse + 2 * BQ;                                  // [2][BQ]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_delta + 2 * BQ);
    uint64_t* kv_full = bars;
    uint64_t* st_full = bars + 1;             // STAGES
    uint64_t* st_empty = st_full + STAGES;    // STAGES
    uint64_t* s_full = st_empty + STAGES;
    uint64_t* p_full = s_full + 1;
    uint64_t* acc_full = p_full + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp_idx = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int kb = blockIdx.x % p.num_blocks;
    const int bh = blockIdx.x / p.num_blocks;
    const int h = bh % p.H, b = bh / p.H;
    const int k0 = kb * BK;
    const int num_q = (p.Nq + BQ - 1) / BQ;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_full, 8);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) { tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish(); }
    pdl_wait();   // prologue above touched only smem / TMEM / kernel params
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            mbar_arrive_expect_tx_e(kv_full, 2 * Cfg::KV_BYTES);
            for (int c = 0; c < Cfg::DCH; ++c) {
                tma_load_4d_e(&tmK, kv_full, smem_k + c * BK * 128, c * 64, h, k0, b);
                tma_load_4d_e(&tmV, kv_full, smem_v + c * BK * 128, c * 64, h, k0, b);
            }
            int stage = 0; uint32_t phase = 0;
            for (int j = 0; j < num_q; ++j) {
                mbar_wait(&st_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx_e(&st_full[stage], Cfg::STAGE_BYTES);
                uint8_t* sq = smem_st + stage * Cfg::STAGE_BYTES;
                uint8_t* sdo = sq + Cfg::QD_TILE;
                for (int c = 0; c < Cfg::DCH; ++c) {
                    tma_load_4d_e(&tmQ, &st_full[stage], sq + c * BQ * 128, c * 64, h, j * BQ, b);
                    tma_load_4d_e(&tmDO, &st_full[stage], sdo + c * BQ * 128, c * 64, h, j * BQ, b);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BQ, 0, 0);    // [keys x queries], both K-major
            constexpr uint32_t idesc_g = make_idesc_bf16(128, DP, 0, 1);    // [keys x d], B MN-major
            mbar_wait(kv_full, 0);
            const uint32_t sk = smem_u32(smem_k), sv = smem_u32(smem_v);
            int stage = 0; uint32_t phase = 0;
            for (int j = 0; j < num_q; ++j) {
                mbar_wait(&st_full[stage], phase);
                tc_fence_after();
                const uint32_t sq = smem_u32(smem_st + stage * Cfg::STAGE_BYTES);
                const uint32_t sdo = sq + Cfg::QD_TILE;
#pragma unroll
                for (int kk = 0; kk < DP / 16; ++kk) {
                    const uint32_t offa = (kk / 4) * (BK * 128) + (kk % 4) * 32;
                    const uint32_t offb = (kk / 4) * (BQ * 128) + (kk % 4) * 32;
                    tc_mma_ss_e(tmem_base + Cfg::TM_ST, make_smem_desc(sk + offa, 16, 1024, 2),
                              make_smem_desc(sq + offb, 16, 1024, 2), idesc_s, kk != 0);
                }
#pragma unroll
                for (int kk = 0; kk < DP / 16; ++kk) {
                    const uint32_t offa = (kk / 4) * (BK * 128) + (kk % 4) * 32;
                    const uint32_t offb = (kk / 4) * (BQ * 128) + (kk % 4) * 32;
                    tc_mma_ss_e(tmem_base + Cfg::TM_DPT, make_smem_desc(sv + offa, 16, 1024, 2),
                              make_smem_desc(sdo + offb, 16, 1024, 2), idesc_s, kk != 0);
                }
                tc_commit_e(s_full);
                mbar_wait(p_full, j & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < BQ / 16; ++kk) {
                    // dV += P^T dO_j ;  dK += dS^T Q_j   (A: bf16 P^T / dS^T of query chunk kk in TMEM; B tiles [queries][64 d]:
                    // MN-major, LBO = chunk stride)
                    tc_mma_ts_e(tmem_base + Cfg::TM_DV, tmem_base + Cfg::TM_ST + Cfg::a_col(kk),
                              make_smem_desc(sdo + kk * 2048, BQ * 128, 1024, 2), idesc_g, (j | kk) != 0);
                    tc_mma_ts_e(tmem_base + Cfg::TM_DK, tmem_base + Cfg::TM_DPT + Cfg::a_col(kk),
                              make_smem_desc(sq + kk * 2048, BQ * 128, 1024, 2), idesc_g, (j | kk) != 0);
                }
                tc_commit_e(&st_empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit_e(acc_full);
        }
    } else if (warp_idx >= 4) {
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane;                // key row within the block
        const int half = (warp_idx - 4) >> 2;          // which half of the query columns this warp converts
        const int st = threadIdx.x - 128;              // 0..255
        const uint32_t lane_off = uint32_t(quad * 32) << 16;
        const bool key_ok = (k0 + r) < p.Nk;
        const long long stat_base = ((long long)b * p.H + h) * p.Nq;
        // LSE and delta * scale of every query of this (b, h): staged ONCE when they fit (Nq <= 4096: 32 KB) - the per-block
        // staging + CTA-wide named barrier of the fallback path was the top stall of this kernel (ncu, round 2)
        float* all_lse = reinterpret_cast<float*>(tmem_ptr_smem + 4);
        float* all_del = all_lse + num_q * BQ;
        if (p.stats_all) {
            for (int i = st; i < num_q * BQ; i += 256) {
                all_lse[i] = (i < p.Nq) ? p.lse[stat_base + i] : INFINITY;     // +inf -> P = 0 for padded queries
                all_del[i] = (i < p.Nq) ? p.delta[stat_base + i] * p.scale : 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        for (int j = 0; j < num_q; ++j) {
            float* ls = p.stats_all ? all_lse + j * BQ : smem_lse + (j & 1) * BQ;
            float* de = p.stats_all ? all_del + j * BQ : smem_delta + (j & 1) * BQ;
            if (!p.stats_all) {
                // fallback: stage this query block's LSE / delta (double-buffered; the named barrier orders it w.r.t. the reads)
                if (st < BQ) {
                    const int q = j * BQ + st;
                    ls[st] = (q < p.Nq) ? p.lse[stat_base + q] : INFINITY;
                    de[st] = (q < p.Nq) ? p.delta[stat_base + q] * p.scale : 0.f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            mbar_wait(s_full, j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = half * (BQ / 32); c < (half + 1) * (BQ / 32); ++c) {
                uint32_t s[16], g[16];
                tmem_ld_32x16(tmem_base + Cfg::TM_ST + lane_off + c * 16, s);
                tmem_ld_32x16(tmem_base + Cfg::TM_DPT + lane_off + c * 16, g);
                uint32_t pk[8], dk_[8];
                float lsv[16], dev[16];     // this chunk's per-query LSE / delta*scale: broadcast 16-byte shared loads
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 l4 = *reinterpret_cast<const float4*>(ls + c * 16 + i);
                    const float4 d4 = *reinterpret_cast<const float4*>(de + c * 16 + i);
                    lsv[i] = l4.x; lsv[i + 1] = l4.y; lsv[i + 2] = l4.z; lsv[i + 3] = l4.w;
                    dev[i] = d4.x; dev[i + 1] = d4.y; dev[i + 2] = d4.z; dev[i + 3] = d4.w;
                }
                tc_wait_ld();
                // no key mask: row r of P^T / dS^T only feeds row r of dV / dK, and rows beyond Nk are never stored
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float p0 = fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -lsv[i]));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -lsv[i + 1]));
                    const float d0 = p0 * fmaf(__uint_as_float(g[i]), p.scale, -dev[i]);
                    const float d1 = p1 * fmaf(__uint_as_float(g[i + 1]), p.scale, -dev[i + 1]);
                    pk[i >> 1] = pack_bf16x2(p0, p1);
                    dk_[i >> 1] = pack_bf16x2(d0, d1);
                }
                // bf16 P^T / dS^T of this chunk -> TMEM, over S^T / dP^T columns this warp has already read
                tmem_st_32x8(tmem_base + Cfg::TM_ST + lane_off + Cfg::a_col(c), pk);
                tmem_st_32x8(tmem_base + Cfg::TM_DPT + lane_off + Cfg::a_col(c), dk_);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const long long row = (long long)b * p.Nk + k0 + r;
        if (half == 0) store_acc_rows<DP>(tmem_base + Cfg::TM_DK + lane_off, p.dk + row * p.lddk + h * p.d, key_ok, p.d, 1.f);
        else store_acc_rows<DP>(tmem_base + Cfg::TM_DV + lane_off, p.dv + row * p.lddv + h * p.d, key_ok, p.d, 1.f);
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ===================================================================================================== dQ
template <int DP, int BKB, int STAGES>
struct DqCfg {
    static constexpr int BM = 128;
    static constexpr int DCH = DP / 64;
    static constexpr int QD_BYTES = DCH * BM * 128;             // Q tile (and dO tile)
    static constexpr int KV_TILE = DCH * BKB * 128;
    static constexpr int STAGE_BYTES = 2 * KV_TILE;
    static constexpr int SMEM_BYTES = 1024 + 2 * QD_BYTES + STAGES * STAGE_BYTES + 256;
    static constexpr int CH_PER_HALF = BKB / 32;                // dS (bf16, TMEM) over the S columns: see DkvCfg::a_col
    __host__ __device__ static constexpr int a_col(int chunk) { return (chunk / CH_PER_HALF) * (BKB / 2) + (chunk % CH_PER_HALF) * 8; }
    static constexpr int TM_S = 0, TM_DP = BKB, TM_DQ = 2 * BKB;
    static constexpr int TMEM_NEED = 2 * BKB + DP;
    static constexpr int TMEM_COLS = TMEM_NEED <= 256 ? 256 : 512;
    static constexpr int MIN_CTAS = (TMEM_COLS <= 256 && SMEM_BYTES <= 110 * 1024) ? 2 : 1;
    static_assert(2 * BKB + DP <= 512, "TMEM");
};

template <int DP, int BKB, int STAGES>
__global__ void __launch_bounds__(BWD_THREADS, (DqCfg<DP, BKB, STAGES>::MIN_CTAS))
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                   const AttnBwdParams p) {
    pdl_launch_dependents();
    using Cfg = DqCfg<DP, BKB, STAGES>;
    constexpr int BM = Cfg::BM;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;
    uint8_t* smem_do = smem_q + Cfg::QD_BYTES;
    uint8_t* smem_kv = smem_do + Cfg::QD_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + STAGES * Cfg::STAGE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;
    uint64_t* p_full = s_full + 1;
    uint64_t* acc_full = p_full + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp_idx = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int qb = blockIdx.x % p.num_blocks;
    const int bh = blockIdx.x / p.num_blocks;
    const int h = bh % p.H, b = bh / p.H;
    const int q0 = qb * BM;
    const int num_kv = (p.Nk + BKB - 1) / BKB;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_full, 8);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) { tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish(); }
    pdl_wait();   // prologue above touched only smem / TMEM / kernel params
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            mbar_arrive_expect_tx_e(q_full, 2 * Cfg::QD_BYTES);
            for (int c = 0; c < Cfg::DCH; ++c) {
                tma_load_4d_e(&tmQ, q_full, smem_q + c * BM * 128, c * 64, h, q0, b);
                tma_load_4d_e(&tmDO, q_full, smem_do + c * BM * 128, c * 64, h, q0, b);
            }
            int stage = 0; uint32_t phase = 0;
            for (int i = 0; i < num_kv; ++i) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx_e(&kv_full[stage], Cfg::STAGE_BYTES);
                uint8_t* sk = smem_kv + stage * Cfg::STAGE_BYTES;
                uint8_t* sv = sk + Cfg::KV_TILE;
                for (int c = 0; c < Cfg::DCH; ++c) {
                    tma_load_4d_e(&tmK, &kv_full[stage], sk + c * BKB * 128, c * 64, h, i * BKB, b);
                    tma_load_4d_e(&tmV, &kv_full[stage], sv + c * BKB * 128, c * 64, h, i * BKB, b);
                }
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BKB, 0, 0);
            constexpr uint32_t idesc_g = make_idesc_bf16(128, DP, 0, 1);
            mbar_wait(q_full, 0);
            const uint32_t sq = smem_u32(smem_q), sdo = smem_u32(smem_do);
            int stage = 0; uint32_t phase = 0;
            for (int i = 0; i < num_kv; ++i) {
                mbar_wait(&kv_full[stage], phase);
                tc_fence_after();
                const uint32_t sk = smem_u32(smem_kv + stage * Cfg::STAGE_BYTES);
                const uint32_t sv = sk + Cfg::KV_TILE;
#pragma unroll
                for (int kk = 0; kk < DP / 16; ++kk) {
                    const uint32_t offa = (kk / 4) * (BM * 128) + (kk % 4) * 32;
                    const uint32_t offb = (kk / 4) * (BKB * 128) + (kk % 4) * 32;
                    tc_mma_ss_e(tmem_base + Cfg::TM_S, make_smem_desc(sq + offa, 16, 1024, 2),
                              make_smem_desc(sk + offb, 16, 1024, 2), idesc_s, kk != 0);
                }
#pragma unroll
                for (int kk = 0; kk < DP / 16; ++kk) {
                    const uint32_t offa = (kk / 4) * (BM * 128) + (kk % 4) * 32;
                    const uint32_t offb = (kk / 4) * (BKB * 128) + (kk % 4) * 32;
                    tc_mma_ss_e(tmem_base + Cfg::TM_DP, make_smem_desc(sdo + offa, 16, 1024, 2),
                              make_smem_desc(sv + offb, 16, 1024, 2), idesc_s, kk != 0);
                }
                tc_commit_e(s_full);
                mbar_wait(p_full, i & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < BKB / 16; ++kk) {
                    tc_mma_ts_e(tmem_base + Cfg::TM_DQ, tmem_base + Cfg::TM_S + Cfg::a_col(kk),
                              make_smem_desc(sk + kk * 2048, BKB * 128, 1024, 2), idesc_g, (i | kk) != 0);
                }
                tc_commit_e(&kv_empty[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            tc_commit_e(acc_full);
        }
    } else if (warp_idx >= 4) {
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_off = uint32_t(quad * 32) << 16;
        const int half = (warp_idx - 4) >> 2;
        const int q = q0 + r;
        const bool q_ok = q < p.Nq;
        const long long stat = ((long long)b * p.H + h) * p.Nq + q;
        const float lse = q_ok ? p.lse[stat] : INFINITY;
        const float delta = q_ok ? p.delta[stat] : 0.f;
        for (int i = 0; i < num_kv; ++i) {
            mbar_wait(s_full, i & 1);
            tc_fence_after();
            const int kbase = i * BKB;
            const bool tail = kbase + BKB > p.Nk;
#pragma unroll
            for (int c = half * (BKB / 32); c < (half + 1) * (BKB / 32); ++c) {
                uint32_t s[16], g[16];
                tmem_ld_32x16(tmem_base + Cfg::TM_S + lane_off + c * 16, s);
                tmem_ld_32x16(tmem_base + Cfg::TM_DP + lane_off + c * 16, g);
                tc_wait_ld();
                uint32_t dk_[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float p0 = fast_exp2(fmaf(__uint_as_float(s[j]), p.scale_log2, -lse));
                    float p1 = fast_exp2(fmaf(__uint_as_float(s[j + 1]), p.scale_log2, -lse));
                    if (tail) {
                        if (kbase + c * 16 + j >= p.Nk) p0 = 0.f;
                        if (kbase + c * 16 + j + 1 >= p.Nk) p1 = 0.f;
                    }
                    const float d0 = p0 * (__uint_as_float(g[j]) - delta) * p.scale;
                    const float d1 = p1 * (__uint_as_float(g[j + 1]) - delta) * p.scale;
                    dk_[j >> 1] = pack_bf16x2(d0, d1);
                }
                tmem_st_32x8(tmem_base + Cfg::TM_S + lane_off + Cfg::a_col(c), dk_);    // bf16 dS chunk -> TMEM (A operand of dQ)
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const long long row = (long long)b * p.Nq + q;
        store_acc_rows<DP>(tmem_base + Cfg::TM_DQ + lane_off, p.dq + row * p.lddq + h * p.d, q_ok, p.d, 1.f, half, 2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ===================================================================================================== d <= 64: pipelined kernels
// The head dim of the 4096- and 1024-token... (d = 40 at 64x64 latents) layers pads to 64.  For that case both gradients run
// as ONE CTA per SM with software pipelining instead of two CTAs that each serialise MMA -> softmax -> MMA:
//   * TMEM holds TWO {S, dP} buffers (dkv: {S^T, dP^T}), so the MMA warp issues the score MMAs of block i+1 BEFORE it
//     waits for the softmax of block i; the tensor core then works on block i+1 while the 16 softmax warps convert block i
//     (ncu, round 2: the softmax warps of the serial kernels spent 36 % of their samples spinning on `s_full`);
//   * bf16 P^T / dS^T / dS go to dedicated TMEM columns (no aliasing with live accumulators) and are consumed as TMEM A operands;
//   * 16 softmax warps (4 per TMEM lane quadrant, one 16-column chunk each) instead of 8: four warps per scheduler hide the
//     tcgen05.ld -> MUFU -> tcgen05.st latency of one another;
//   * dkv: the per-block LSE / delta vectors arrive with the Q / dO stage through 1-D bulk copies (no per-iteration CTA barrier).
// TMEM map (512 columns):  dq : S0 0-63 | dP0 64-127 | S1 128-191 | dP1 192-255 | dS0 256-287 | dS1 288-319 | dQ 320-383
//                          dkv: ST0 | dPT0 | ST1 | dPT1 (0-255) | PT0 256-287 | dST0 288-319 | PT1 320-351 | dST1 352-383 | dK 384-447 | dV 448-511
static constexpr int BWD64_SOFTMAX_WARPS = 16;
static constexpr int BWD64_THREADS = 128 + BWD64_SOFTMAX_WARPS * 32;

template <int STAGES>
struct Dq64Cfg {
    static constexpr int BM = 128, BKB = 64, DP = 64;
    static constexpr int QD_BYTES = BM * 128;                  // Q tile (and dO tile): 128 rows x 64 (padded) dims
    static constexpr int KV_TILE = BKB * 128;
    static constexpr int STAGE_BYTES = 2 * KV_TILE;
    // >= 120 KB on purpose: the kernel owns all 512 TMEM columns, a second resident CTA would only wait for them
    static constexpr int SMEM_NEED = 1024 + 2 * QD_BYTES + STAGES * STAGE_BYTES + 256;
    static constexpr int SMEM_BYTES = SMEM_NEED > 120 * 1024 ? SMEM_NEED : 120 * 1024;
    __host__ __device__ static constexpr int tm_s(int b) { return b * 128; }
    __host__ __device__ static constexpr int tm_dp(int b) { return b * 128 + 64; }
    __host__ __device__ static constexpr int tm_ds(int b) { return 256 + b * 32; }
    static constexpr int TM_DQ = 320;
};

template <int STAGES>
__global__ void __launch_bounds__(BWD64_THREADS, 1)
attn_bwd_dq64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                     const AttnBwdParams p) {
    pdl_launch_dependents();
    using Cfg = Dq64Cfg<STAGES>;
    constexpr int BM = Cfg::BM, BKB = Cfg::BKB;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_q = smem;
    uint8_t* smem_do = smem_q + Cfg::QD_BYTES;
    uint8_t* smem_kv = smem_do + Cfg::QD_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + STAGES * Cfg::STAGE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + STAGES;
    uint64_t* s_full = kv_empty + STAGES;     // [2]
    uint64_t* p_full = s_full + 2;            // [2]
    uint64_t* acc_full = p_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp_idx = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int qb = blockIdx.x % p.num_blocks;
    const int bh = blockIdx.x / p.num_blocks;
    const int h = bh % p.H, b = bh / p.H;
    const int q0 = qb * BM;
    const int num_kv = (p.Nk + BKB - 1) / BKB;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], BWD64_SOFTMAX_WARPS); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            mbar_arrive_expect_tx_e(q_full, 2 * Cfg::QD_BYTES);
            tma_load_4d_e(&tmQ, q_full, smem_q, 0, h, q0, b);
            tma_load_4d_e(&tmDO, q_full, smem_do, 0, h, q0, b);
            int stage = 0; uint32_t phase = 0;
            for (int i = 0; i < num_kv; ++i) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx_e(&kv_full[stage], Cfg::STAGE_BYTES);
                uint8_t* sk = smem_kv + stage * Cfg::STAGE_BYTES;
                tma_load_4d_e(&tmK, &kv_full[stage], sk, 0, h, i * BKB, b);
                tma_load_4d_e(&tmV, &kv_full[stage], sk + Cfg::KV_TILE, 0, h, i * BKB, b);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BKB, 0, 0);
            constexpr uint32_t idesc_g = make_idesc_bf16(128, 64, 0, 1);
            mbar_wait(q_full, 0);
            const uint32_t sq = smem_u32(smem_q), sdo = smem_u32(smem_do);
            int stage_s = 0; uint32_t phase_s = 0;      // ring position of the NEXT score block to issue
            int stage_q = 0;                            // ring position of the block whose dQ is issued next
            auto issue_scores = [&](int i) {
                mbar_wait(&kv_full[stage_s], phase_s);
                tc_fence_after();
                const uint32_t sk = smem_u32(smem_kv + stage_s * Cfg::STAGE_BYTES);
                const uint32_t sv = sk + Cfg::KV_TILE;
                const int bb = i & 1;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tc_mma_ss_e(tmem_base + Cfg::tm_s(bb), make_smem_desc(sq + kk * 32, 16, 1024, 2),
                              make_smem_desc(sk + kk * 32, 16, 1024, 2), idesc_s, kk != 0);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tc_mma_ss_e(tmem_base + Cfg::tm_dp(bb), make_smem_desc(sdo + kk * 32, 16, 1024, 2),
                              make_smem_desc(sv + kk * 32, 16, 1024, 2), idesc_s, kk != 0);
                tc_commit_e(&s_full[bb]);
                if (++stage_s == STAGES) { stage_s = 0; phase_s ^= 1; }
            };
            issue_scores(0);
            for (int i = 0; i < num_kv; ++i) {
                if (i + 1 < num_kv) issue_scores(i + 1);           // block i+1 runs on the tensor core while block i is in its softmax
                mbar_wait(&p_full[i & 1], (i >> 1) & 1);
                tc_fence_after();
                const uint32_t sk = smem_u32(smem_kv + stage_q * Cfg::STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BKB / 16; ++kk)
                    tc_mma_ts_e(tmem_base + Cfg::TM_DQ, tmem_base + Cfg::tm_ds(i & 1) + kk * 8,
                              make_smem_desc(sk + kk * 2048, BKB * 128, 1024, 2), idesc_g, (i | kk) != 0);
                tc_commit_e(&kv_empty[stage_q]);
                if (++stage_q == STAGES) stage_q = 0;
            }
            tc_commit_e(acc_full);
        }
    } else if (warp_idx >= 4) {
        const int quad = warp_idx & 3;
        const int sub = (warp_idx - 4) >> 2;          // 0..3: which 16-key chunk of the block this warp converts
        const int r = quad * 32 + lane;
        const uint32_t lane_off = uint32_t(quad * 32) << 16;
        const int q = q0 + r;
        const bool q_ok = q < p.Nq;
        const long long stat = ((long long)b * p.H + h) * p.Nq + q;
        const float lse = q_ok ? p.lse[stat] : INFINITY;
        const float delta_s = q_ok ? p.delta[stat] * p.scale : 0.f;
        for (int i = 0; i < num_kv; ++i) {
            const int bb = i & 1;
            mbar_wait(&s_full[bb], (i >> 1) & 1);
            tc_fence_after();
            const int kbase = i * BKB + sub * 16;
            uint32_t s[16], g[16];
            tmem_ld_32x16(tmem_base + Cfg::tm_s(bb) + lane_off + sub * 16, s);
            tmem_ld_32x16(tmem_base + Cfg::tm_dp(bb) + lane_off + sub * 16, g);
            tc_wait_ld();
            uint32_t dk_[8];
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float p0 = fast_exp2(fmaf(__uint_as_float(s[j]), p.scale_log2, -lse));
                float p1 = fast_exp2(fmaf(__uint_as_float(s[j + 1]), p.scale_log2, -lse));
                if (kbase + j >= p.Nk) p0 = 0.f;
                if (kbase + j + 1 >= p.Nk) p1 = 0.f;
                const float d0 = p0 * fmaf(__uint_as_float(g[j]), p.scale, -delta_s);
                const float d1 = p1 * fmaf(__uint_as_float(g[j + 1]), p.scale, -delta_s);
                dk_[j >> 1] = pack_bf16x2(d0, d1);
            }
            tmem_st_32x8(tmem_base + Cfg::tm_ds(bb) + lane_off + sub * 8, dk_);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[bb]);
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const long long row = (long long)b * p.Nq + q;
        if (sub < 2) store_acc_rows<64>(tmem_base + Cfg::TM_DQ + lane_off, p.dq + row * p.lddq + h * p.d, q_ok, p.d, 1.f, sub, 2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <int STAGES>
struct Dkv64Cfg {
    static constexpr int BK = 128, BQ = 64, DP = 64;
    static constexpr int KV_BYTES = BK * 128;                  // K tile (and V tile)
    static constexpr int QD_TILE = BQ * 128;                   // Q_j tile (and dO_j tile)
    static constexpr int STAT_BYTES = 2 * BQ * 4;              // lse | delta of the query block
    static constexpr int STAGE_BYTES = 2 * QD_TILE + STAT_BYTES;
    static constexpr int SMEM_NEED = 1024 + 2 * KV_BYTES + STAGES * STAGE_BYTES + 256;
    static constexpr int SMEM_BYTES = SMEM_NEED > 120 * 1024 ? SMEM_NEED : 120 * 1024;
    __host__ __device__ static constexpr int tm_st(int b) { return b * 128; }
    __host__ __device__ static constexpr int tm_dpt(int b) { return b * 128 + 64; }
    __host__ __device__ static constexpr int tm_pt(int b) { return 256 + b * 64; }
    __host__ __device__ static constexpr int tm_dst(int b) { return 256 + b * 64 + 32; }
    static constexpr int TM_DK = 384, TM_DV = 448;
};

// Requires Nq % 64 == 0 and 16-byte aligned lse / delta rows (checked on the host; otherwise the serial kernel runs).
template <int STAGES>
__global__ void __launch_bounds__(BWD64_THREADS, 1)
attn_bwd_dkv64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                      const AttnBwdParams p) {
    pdl_launch_dependents();
    using Cfg = Dkv64Cfg<STAGES>;
    constexpr int BK = Cfg::BK, BQ = Cfg::BQ;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_k = smem;
    uint8_t* smem_v = smem_k + Cfg::KV_BYTES;
    uint8_t* smem_st = smem_v + Cfg::KV_BYTES;                       // stages: [Q_j | dO_j | lse_j | delta_j]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_st + STAGES * Cfg::STAGE_BYTES);
    uint64_t* kv_full = bars;
    uint64_t* st_full = bars + 1;             // STAGES
    uint64_t* st_empty = st_full + STAGES;    // STAGES
    uint64_t* s_full = st_empty + STAGES;     // [2]
    uint64_t* p_full = s_full + 2;            // [2]
    uint64_t* acc_full = p_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp_idx = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int kb = blockIdx.x % p.num_blocks;
    const int bh = blockIdx.x / p.num_blocks;
    const int h = bh % p.H, b = bh / p.H;
    const int k0 = kb * BK;
    const int num_q = p.Nq / BQ;
    const long long stat_base = ((long long)b * p.H + h) * p.Nq;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    }
    if (warp_idx == 1 && lane == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&st_full[i], 1); mbar_init(&st_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], BWD64_SOFTMAX_WARPS); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) { tmem_alloc(tmem_ptr_smem, 512); tmem_relinquish(); }
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            mbar_arrive_expect_tx_e(kv_full, 2 * Cfg::KV_BYTES);
            tma_load_4d_e(&tmK, kv_full, smem_k, 0, h, k0, b);
            tma_load_4d_e(&tmV, kv_full, smem_v, 0, h, k0, b);
            int stage = 0; uint32_t phase = 0;
            for (int j = 0; j < num_q; ++j) {
                mbar_wait(&st_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx_e(&st_full[stage], Cfg::STAGE_BYTES);
                uint8_t* sq = smem_st + stage * Cfg::STAGE_BYTES;
                tma_load_4d_e(&tmQ, &st_full[stage], sq, 0, h, j * BQ, b);
                tma_load_4d_e(&tmDO, &st_full[stage], sq + Cfg::QD_TILE, 0, h, j * BQ, b);
                bulk_copy_g2s_e(sq + 2 * Cfg::QD_TILE, p.lse + stat_base + j * BQ, BQ * 4, &st_full[stage]);
                bulk_copy_g2s_e(sq + 2 * Cfg::QD_TILE + BQ * 4, p.delta + stat_base + j * BQ, BQ * 4, &st_full[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        {   // whole warp converged; the *_e calls elect one lane for the instruction itself
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BQ, 0, 0);    // [keys x queries], both K-major
            constexpr uint32_t idesc_g = make_idesc_bf16(128, 64, 0, 1);    // [keys x d], B MN-major
            mbar_wait(kv_full, 0);
            const uint32_t sk = smem_u32(smem_k), sv = smem_u32(smem_v);
            int stage_s = 0; uint32_t phase_s = 0;
            int stage_g = 0;
            auto issue_scores = [&](int j) {
                mbar_wait(&st_full[stage_s], phase_s);
                tc_fence_after();
                const uint32_t sq = smem_u32(smem_st + stage_s * Cfg::STAGE_BYTES);
                const uint32_t sdo = sq + Cfg::QD_TILE;
                const int bb = j & 1;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tc_mma_ss_e(tmem_base + Cfg::tm_st(bb), make_smem_desc(sk + kk * 32, 16, 1024, 2),
                              make_smem_desc(sq + kk * 32, 16, 1024, 2), idesc_s, kk != 0);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tc_mma_ss_e(tmem_base + Cfg::tm_dpt(bb), make_smem_desc(sv + kk * 32, 16, 1024, 2),
                              make_smem_desc(sdo + kk * 32, 16, 1024, 2), idesc_s, kk != 0);
                tc_commit_e(&s_full[bb]);
                if (++stage_s == STAGES) { stage_s = 0; phase_s ^= 1; }
            };
            issue_scores(0);
            for (int j = 0; j < num_q; ++j) {
                if (j + 1 < num_q) issue_scores(j + 1);
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t sq = smem_u32(smem_st + stage_g * Cfg::STAGE_BYTES);
                const uint32_t sdo = sq + Cfg::QD_TILE;
#pragma unroll
                for (int kk = 0; kk < BQ / 16; ++kk) {
                    tc_mma_ts_e(tmem_base + Cfg::TM_DV, tmem_base + Cfg::tm_pt(j & 1) + kk * 8,
                              make_smem_desc(sdo + kk * 2048, BQ * 128, 1024, 2), idesc_g, (j | kk) != 0);
                    tc_mma_ts_e(tmem_base + Cfg::TM_DK, tmem_base + Cfg::tm_dst(j & 1) + kk * 8,
                              make_smem_desc(sq + kk * 2048, BQ * 128, 1024, 2), idesc_g, (j | kk) != 0);
                }
                tc_commit_e(&st_empty[stage_g]);
                if (++stage_g == STAGES) stage_g = 0;
            }
            tc_commit_e(acc_full);
        }
    } else if (warp_idx >= 4) {
        const int quad = warp_idx & 3;
        const int sub = (warp_idx - 4) >> 2;          // 0..3: 16-query chunk of the block
        const int r = quad * 32 + lane;               // key row within the block == TMEM lane
        const uint32_t lane_off = uint32_t(quad * 32) << 16;
        const bool key_ok = (k0 + r) < p.Nk;
        int stage = 0;
        for (int j = 0; j < num_q; ++j) {
            const int bb = j & 1;
            mbar_wait(&s_full[bb], (j >> 1) & 1);     // S^T / dP^T of block j are complete => its stage (with lse / delta) has landed
            tc_fence_after();
            const float* ls = reinterpret_cast<const float*>(smem_st + stage * Cfg::STAGE_BYTES + 2 * Cfg::QD_TILE) + sub * 16;
            const float* de = ls + BQ;
            uint32_t s[16], g[16];
            tmem_ld_32x16(tmem_base + Cfg::tm_st(bb) + lane_off + sub * 16, s);
            tmem_ld_32x16(tmem_base + Cfg::tm_dpt(bb) + lane_off + sub * 16, g);
            float lsv[16], dev[16];                   // per-query LSE / delta of this chunk: broadcast 16-byte shared loads
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(ls + i);
                const float4 d4 = *reinterpret_cast<const float4*>(de + i);
                lsv[i] = l4.x; lsv[i + 1] = l4.y; lsv[i + 2] = l4.z; lsv[i + 3] = l4.w;
                dev[i] = d4.x; dev[i + 1] = d4.y; dev[i + 2] = d4.z; dev[i + 3] = d4.w;
            }
            tc_wait_ld();
            uint32_t pk[8], dk_[8];
            // no key mask: a row of P^T / dS^T only feeds the same row of dV / dK, and rows beyond Nk are never stored
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const float p0 = fast_exp2(fmaf(__uint_as_float(s[i]), p.scale_log2, -lsv[i]));
                const float p1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]), p.scale_log2, -lsv[i + 1]));
                const float d0 = p0 * (__uint_as_float(g[i]) - dev[i]) * p.scale;
                const float d1 = p1 * (__uint_as_float(g[i + 1]) - dev[i + 1]) * p.scale;
                pk[i >> 1] = pack_bf16x2(p0, p1);
                dk_[i >> 1] = pack_bf16x2(d0, d1);
            }
            tmem_st_32x8(tmem_base + Cfg::tm_pt(bb) + lane_off + sub * 8, pk);
            tmem_st_32x8(tmem_base + Cfg::tm_dst(bb) + lane_off + sub * 8, dk_);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[bb]);
            if (++stage == STAGES) stage = 0;
        }
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const long long row = (long long)b * p.Nk + k0 + r;
        // 16 warps, 4 x 32-column granules (dK 0/1, dV 0/1): one per warp group `sub`
        if (sub < 2) store_acc_rows<64>(tmem_base + Cfg::TM_DK + lane_off, p.dk + row * p.lddk + h * p.d, key_ok, p.d, 1.f, sub, 2);
        else store_acc_rows<64>(tmem_base + Cfg::TM_DV + lane_off, p.dv + row * p.lddv + h * p.d, key_ok, p.d, 1.f, sub - 2, 2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ===================================================================================================== delta
// delta[b, h, q] = sum_d dO[b, q, h*d + :] * O[b, q, h*d + :]
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, long long ldo, const __nv_bfloat16* __restrict__ d_o,
                                  long long lddo, float* __restrict__ delta, int B, int H, int Nq, int d) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)B * Nq * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int h = (int)(i % H);
        const long long bq = i / H;
        const int q = (int)(bq % Nq);
        const int b = (int)(bq / Nq);
        const __nv_bfloat16* po = o + bq * ldo + h * d;
        const __nv_bfloat16* pd = d_o + bq * lddo + h * d;
        float s = 0.f;
        for (int c = 0; c < d; c += 8) {
            const uint4 a = *reinterpret_cast<const uint4*>(po + c);
            const uint4 g = *reinterpret_cast<const uint4*>(pd + c);
            const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
            const float2 g0 = unpack_bf16x2(g.x), g1 = unpack_bf16x2(g.y), g2 = unpack_bf16x2(g.z), g3 = unpack_bf16x2(g.w);
            s += a0.x * g0.x + a0.y * g0.y + a1.x * g1.x + a1.y * g1.y + a2.x * g2.x + a2.y * g2.y + a3.x * g3.x + a3.y * g3.y;
        }
        delta[((long long)b * H + h) * Nq + q] = s;
    }
}

template <int DP, int BQ, int STAGES_KV, int BKB, int STAGES_Q>
static int launch_attn_bwd(const cl_attn_bwd_args* a, cudaStream_t stream) {
    CUtensorMap tq, tk, tv, tdo;
    AttnBwdParams p;
    p.B = a->B; p.H = a->H; p.Nq = a->Nq; p.Nk = a->Nk; p.d = a->d;
    p.lse = a->lse; p.delta = a->delta;
    p.dq = reinterpret_cast<__nv_bfloat16*>(a->dq); p.lddq = a->lddq;
    p.dk = reinterpret_cast<__nv_bfloat16*>(a->dk); p.lddk = a->lddk;
    p.dv = reinterpret_cast<__nv_bfloat16*>(a->dv); p.lddv = a->lddv;
    p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
    p.num_blocks = 0; p.stats_all = 0;
    {
        const long long total = (long long)a->B * a->Nq * a->H;
        int blocks = (int)((total + 255) / 256);
        if (blocks > num_sms() * 8) blocks = num_sms() * 8;
        launch_k(attn_delta_kernel, blocks, 256, 0, stream, reinterpret_cast<const __nv_bfloat16*>(a->o), a->ldo,
                                                      reinterpret_cast<const __nv_bfloat16*>(a->d_o), a->lddo, a->delta,
                                                      a->B, a->H, a->Nq, a->d);
        count_launch();
    }
    // opt-in (CLB_ATTN_BWD_V2=1): measured on B200 (round 2) the one-CTA-per-SM pipelined kernels run the 4096-token d = 40
    // backward in 1.94 ms against 1.56 ms for the two-CTAs-per-SM kernels below - one MMA-issuing thread per SM paces the
    // tensor pipe at ~150 clk per tcgen05.mma in this kind of loop, two resident CTAs give it two issuers.
    static const int v2_enabled = [] { const char* e = getenv("CLB_ATTN_BWD_V2"); return (e && e[0] == '1') ? 1 : 0; }();
    constexpr int ST64 = 4;
    const bool stats_aligned = (a->Nq % 64 == 0) && ((reinterpret_cast<uintptr_t>(a->lse) & 15) == 0) &&
                               ((reinterpret_cast<uintptr_t>(a->delta) & 15) == 0);
    if (DP == 64 && v2_enabled && a->dk != nullptr && stats_aligned) {
        using Cfg = Dkv64Cfg<ST64>;
        CL_CHECK(make_head_map(&tq, a->q, a->B, a->H, a->Nq, a->d, a->ldq, Cfg::BQ));
        CL_CHECK(make_head_map(&tdo, a->d_o, a->B, a->H, a->Nq, a->d, a->lddo, Cfg::BQ));
        CL_CHECK(make_head_map(&tk, a->k, a->B, a->H, a->Nk, a->d, a->ldk, 128));
        CL_CHECK(make_head_map(&tv, a->v, a->B, a->H, a->Nk, a->d, a->ldv, 128));
        static bool done = false;
        if (!done) {
            CL_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv64_kernel<ST64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
            done = true;
        }
        p.num_blocks = (a->Nk + 127) / 128;
        launch_k(attn_bwd_dkv64_kernel<ST64>, a->B * a->H * p.num_blocks, BWD64_THREADS, Cfg::SMEM_BYTES, stream, tq, tk, tv, tdo, p);
        count_launch();
    } else if (a->dk != nullptr) {
        using Cfg = DkvCfg<DP, BQ, STAGES_KV>;
        CL_CHECK(make_head_map(&tq, a->q, a->B, a->H, a->Nq, a->d, a->ldq, BQ));
        CL_CHECK(make_head_map(&tdo, a->d_o, a->B, a->H, a->Nq, a->d, a->lddo, BQ));
        CL_CHECK(make_head_map(&tk, a->k, a->B, a->H, a->Nk, a->d, a->ldk, 128));
        CL_CHECK(make_head_map(&tv, a->v, a->B, a->H, a->Nk, a->d, a->ldv, 128));
        static bool done = false;
        if (!done) {
            CL_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_kernel<DP, BQ, STAGES_KV>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            done = true;
        }
        p.num_blocks = (a->Nk + 127) / 128;
        const int nq_pad = (a->Nq + BQ - 1) / BQ * BQ;
        const int stats_bytes = nq_pad * 8;
        p.stats_all = (Cfg::SMEM_BYTES + stats_bytes <= (Cfg::MIN_CTAS == 2 ? 110 * 1024 : 200 * 1024)) ? 1 : 0;
        launch_k(attn_bwd_dkv_kernel<DP, BQ, STAGES_KV>, a->B * a->H * p.num_blocks, BWD_THREADS,
                 Cfg::SMEM_BYTES + (p.stats_all ? stats_bytes : 0), stream, tq, tk, tv, tdo, p);
        count_launch();
    }
    if (DP == 64 && v2_enabled && a->dq != nullptr) {
        using Cfg = Dq64Cfg<ST64>;
        CL_CHECK(make_head_map(&tq, a->q, a->B, a->H, a->Nq, a->d, a->ldq, 128));
        CL_CHECK(make_head_map(&tdo, a->d_o, a->B, a->H, a->Nq, a->d, a->lddo, 128));
        CL_CHECK(make_head_map(&tk, a->k, a->B, a->H, a->Nk, a->d, a->ldk, Cfg::BKB));
        CL_CHECK(make_head_map(&tv, a->v, a->B, a->H, a->Nk, a->d, a->ldv, Cfg::BKB));
        static bool done = false;
        if (!done) {
            CL_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq64_kernel<ST64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
            done = true;
        }
        p.num_blocks = (a->Nq + 127) / 128;
        launch_k(attn_bwd_dq64_kernel<ST64>, a->B * a->H * p.num_blocks, BWD64_THREADS, Cfg::SMEM_BYTES, stream, tq, tk, tv, tdo, p);
        count_launch();
    } else if (a->dq != nullptr) {
        using Cfg = DqCfg<DP, BKB, STAGES_Q>;
        CL_CHECK(make_head_map(&tq, a->q, a->B, a->H, a->Nq, a->d, a->ldq, 128));
        CL_CHECK(make_head_map(&tdo, a->d_o, a->B, a->H, a->Nq, a->d, a->lddo, 128));
        CL_CHECK(make_head_map(&tk, a->k, a->B, a->H, a->Nk, a->d, a->ldk, BKB));
        CL_CHECK(make_head_map(&tv, a->v, a->B, a->H, a->Nk, a->d, a->ldv, BKB));
        static bool done = false;
        if (!done) {
            CL_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_kernel<DP, BKB, STAGES_Q>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
            done = true;
        }
        p.num_blocks = (a->Nq + 127) / 128;
        launch_k(attn_bwd_dq_kernel<DP, BKB, STAGES_Q>, a->B * a->H * p.num_blocks, BWD_THREADS, Cfg::SMEM_BYTES, stream, tq, tk, tv, tdo, p);
        count_launch();
    }
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}

}  // namespace clb

using namespace clb;

extern "C" int cl_attn_bwd(const cl_attn_bwd_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->q || !a->k || !a->v || !a->o || !a->d_o || !a->lse || !a->delta)
        return set_error(CL_ERR_INVALID, "cl_attn_bwd: null pointer");
    if ((a->dk == nullptr) != (a->dv == nullptr)) return set_error(CL_ERR_INVALID, "cl_attn_bwd: dk and dv go together");
    if (a->d % 8 != 0 || a->d <= 0 || a->d > 192) return set_error(CL_ERR_UNSUPPORTED, "cl_attn_bwd: head dim must be a multiple of 8, <= 192");
    if ((a->ldq % 8) || (a->ldk % 8) || (a->ldv % 8) || (a->ldo % 8) || (a->lddo % 8) || (a->dq && a->lddq % 8) ||
        (a->dk && ((a->lddk % 8) || (a->lddv % 8))))
        return set_error(CL_ERR_INVALID, "cl_attn_bwd: row strides must be multiples of 8");
    if (a->d <= 64) return launch_attn_bwd<64, 64, 2, 64, 2>(a, stream);   // 256 TMEM columns, ~97 KB smem: two CTAs per SM
    if (a->d <= 128) return launch_attn_bwd<128, 128, 1, 128, 1>(a, stream);
    return launch_attn_bwd<192, 64, 1, 64, 2>(a, stream);
}
