// K1/K4 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One CTA per SM, 12 warps:
//   warp 0      TMA producer   (A tile 128x64, B tile BNx64 [+16 LoRA-down rows], 128B swizzle, mbarrier ring)
//   warp 1      MMA issuer     (one elected thread, tcgen05.mma cta_group::1, M=128, N=BN[+16], K=16)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue       (tcgen05.ld -> smem transpose / TMA store -> bias/residual -> coalesced global stores); two
//                               TMEM accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1.
// LoRA (EXT = 16): the 16 extra B rows (bf16 hi / lo halves of the rank-r down matrix) make the main loop produce
// t = x A^T in 16 extra accumulator columns; four epilogue warps read them (+ t_add, -> t_out, * scale), split t into bf16
// hi / lo and write it as a 128 x 32 K-major operand into shared memory, next to the same split of the up matrix, and the MMA
// warp adds  t B^T  to the accumulator with one (rank <= 4) or two (rank <= 8) more tcgen05.mma (K = 16): the rank-r update
// costs the tensor pipe ~100 clk per tile instead of 128 FMAs + 32 broadcast shared loads per thread and granule.
// The A operand is either a plain row-major matrix or an NHWC image read as 3x3 windows (implicit GEMM: the
// "im2col" is done by TMA box loads with shifted coordinates, out-of-bounds zero fill == conv zero padding).
//
// Replaces the cuBLAS/cuDNN calls behind models.py:124-147,231-282,373-423 (attention projections + LoRA side
// path) and diffusers' FeedForward / ResnetBlock2D / Transformer2DModel projections.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <unordered_map>

#include "common.cuh"
#define CLB_FAMILY 1      // CLB_PDL_MASK bit of this file's kernels
#include "host_common.h"
#include "../../include/controllora_b200.h"

namespace clb {

#ifndef CLB_REGS_CTRL
#define CLB_REGS_CTRL 56     // setmaxnreg of the producer / MMA / allocator warpgroup
#define CLB_REGS_EPI 224     // ... and of the two epilogue warpgroups (128 * 56 + 256 * 224 = 64512 = 384 * 168, the launch allocation)
#endif
static constexpr int BLOCK_M = 128;
// BLOCK_K (template BK): 64 bf16 = 128 B rows with the 128B swizzle, or 32 bf16 = 64 B rows with the 64B swizzle
// (used for the 32-channel layers of the hint encoder, where a 3x3 tap only offers 32 contiguous K elements).
static constexpr int STRIP_ROWS = 130;                    // 128 output pixels + one halo pixel either side
static constexpr int STRIP_SLOT = 8704;                   // 130 x 64 B rounded up to the 512-byte period of the 64B swizzle
static constexpr int STRIP_STAGE = 26624;                 // three halo rows, 1024-aligned
static constexpr int NUM_EPI_WARPS = 8;
static constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;
static constexpr int STAGE_ROW_BYTES = 128;              // epilogue staging: 32 fp32 columns per row
static constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * 32 * STAGE_ROW_BYTES;

struct GemmParams {
    int M, N, K;
    int num_k_blocks;
    int a_mode;
    // conv geometry (output space)
    int n_img, Ho, Wo, C, cblocks, pad_lo;
    int bw, bh, bn, tiles_w, tiles_h, tiles_n;
    int num_m_blocks, num_n_blocks;
    // epilogue
    const float* bias;
    const float* row_bias;
    int rows_per_group;
    long long ld_rb;
    const __nv_bfloat16* residual;
    long long ldr;
    const float* lora_up;
    int lora_rp;
    float lora_scale;
    const float* t_add;
    float* t_out;
    void* out;
    long long ldd;
    int out_fp32;
    int tma_store;    // bf16 output through per-warp TMA stores of 32x32 sub-tiles (new epilogue); 0 = per-lane global stores
    int ew, eh, en;   // conv: the 32 rows of a TMEM lane quadrant as a box of the output image (ew * eh * en == 32)
    int strip;        // 32-channel stride-1 conv, 128-pixel row tiles: a stage = the three 130-pixel halo rows of the tile, the nine taps
                      // are shared-memory descriptors shifted by 0 / 1 / 2 pixels (3 TMA loads of 130 rows instead of 9 of 128)
    int strip_bo;     // debug: how the descriptor base offset of a shifted tap is formed (0: none, 1: (addr >> 7) & 3, 2: (addr >> 7) & 7)
    int epi_bf16;     // per-lane store epilogue: nothing is added to the accumulator -> transpose in bf16 (host-checked alignment)
    int b_resident;   // B (weights) tile of this CTA's n-block stays in smem for the CTA lifetime (small K)
    // split-K (streaming mode only): tile space is (m, n, split); split s covers k-blocks [s*kb_per_split, ...) and stores
    // its fp32 partial tile to split_ws[s][M][N]; splitk_finish_kernel sums the partials and applies the epilogue.
    int splits, kb_per_split;
    float* split_ws;
    int dbg_reps;     // CLB_TIMELINE builds only: every k-step's MMAs are issued 1 + dbg_reps times (tensor-pipe rate probe)
    int dbg_id;       // CLB_TIMELINE builds only: launch ordinal (mod 64) for the globaltimer entry / exit record
};

#ifdef CLB_TIMELINE
// debug: per-SM event log  [sm][slot] = (clock64, tag)  -- tools/gemm_timeline.py prints it
// three roles per SM (producer / MMA / first epilogue warp), 80 slots each; the slot counter is a register of the calling
// thread, so an event costs one clock read and two plain stores (the first version used a global atomic per event, ~1000
// cycles each, which distorted the very thing it measured)
__device__ unsigned long long g_tl[160 * 256 * 2];
__device__ unsigned int g_tl_n[160];
// globaltimer (ns) per launch ordinal: [id][0] = first CTA entry, [1] = last CTA entry, [2] = first CTA exit, [3] = last CTA exit
__device__ unsigned long long g_gt[64 * 4];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_rec(int tag, int& slot) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int role = tag < 20 ? 0 : (tag < 30 ? 1 : 2);
    if (slot < 80) {
        const unsigned int i = role * 80 + slot;
        g_tl[(smid * 256 + i) * 2] = clock64();
        g_tl[(smid * 256 + i) * 2 + 1] = (unsigned long long)tag;
        ++slot;
    }
}
#define TL(tag) tl_rec(tag, tl_slot)
#if CLB_TIMELINE >= 2
#define TLK(tag) tl_rec(tag, tl_slot)      // per-k-block events (perturb the producer / MMA cadence: level 2 only)
#else
#define TLK(tag)
#endif
#else
#define TL(tag)
#define TLK(tag)
#endif

// tile -> (m_blk, n_blk).  Streaming mode: n fastest (CTAs that run concurrently share the A rows through L2).
// B-resident mode: a CTA keeps ONE n-block for its lifetime and walks the m-blocks.
struct TileIter {
    int m_blk, n_blk, split, step, num_m;
    bool resident;
    int tile, num_tiles, num_n, splits, kbps, nkb;
    // cta / num_ctas: index and count of the scheduling units (CTAs, or CTA pairs); num_m_units: m-blocks per unit row
    __device__ TileIter(const GemmParams& p, int cta, int num_ctas, int num_m_units) {
        resident = p.b_resident != 0;
        num_m = num_m_units; num_n = p.num_n_blocks; splits = p.splits; kbps = p.kb_per_split; nkb = p.num_k_blocks;
        num_tiles = num_m * num_n * splits;
        split = 0; tile = 0;
        if (resident) { n_blk = cta % num_n; m_blk = cta / num_n; step = num_ctas / num_n; }
        else { tile = cta; step = num_ctas; decode(); }
    }
    __device__ void decode() {
        const int t2 = tile / splits;
        split = tile - t2 * splits;
        m_blk = t2 / num_n;
        n_blk = t2 - m_blk * num_n;
    }
    __device__ bool valid() const { return resident ? (m_blk < num_m) : (tile < num_tiles); }
    __device__ void next() {
        if (resident) m_blk += step;
        else { tile += step; decode(); }
    }
    __device__ int kb_begin() const { return split * kbps; }
    __device__ int kb_end() const { return min(nkb, (split + 1) * kbps); }
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile: each
// CTA stages its own 128 A rows and HALF of the B rows, the leader's single MMA thread drives both tensor cores, and
// each CTA ends up with its 128 x BN accumulator rows in its own TMEM -> 1.5x less L2->SM traffic per flop at BN = 256.
template <int BN, int EXT, int BK, int CG = 1>
struct GemmCfg {
    // BN > 256 (the 256 x 320 CTA-pair tile): the tile's columns are produced by NSPLIT MMAs of MMA_N columns each that share
    // the A stage - per k-block a CTA then moves 128 A rows + BN/2 B rows for BN columns of work, 1.45x fewer L2->SM bytes per
    // flop than the 256 x 160 tile (the deep-K N = 320 / 640 GEMMs sat at ~37 B/clk/SM of operand traffic = the L2 limit)
    static constexpr int NSPLIT = (BN > 256) ? 2 : 1;
    static constexpr int MMA_N = BN / NSPLIT + EXT;           // N of one tcgen05.mma
    static constexpr int UMMA_N = MMA_N;
    static constexpr int ROW_BYTES = BK * 2;
    // k-blocks per pipeline stage: the 32-channel convs (64-byte rows) move only 8 KB of A per k-block, so the per-k-block costs
    // of the two single-thread loops (barrier round trip, coordinate math, commit: ~400 clk) dominated; three k-blocks (the
    // three horizontal taps of one filter row when C = 32) share a stage, a barrier and a commit.  9 * C / 32 k-blocks: always % 3.
    static constexpr int KPS = (BK == 32) ? 3 : 1;
    static constexpr int A_SUB_BYTES = BLOCK_M * ROW_BYTES;
    static constexpr int A_STAGE_BYTES = KPS * A_SUB_BYTES;
    static constexpr int MMA_B_ROWS = MMA_N / CG;             // B rows one CTA stages for one MMA
    static constexpr int B_ROWS = NSPLIT * MMA_B_ROWS;        // B rows staged by one CTA per k-block
    static constexpr int B_SUB_BYTES = B_ROWS * ROW_BYTES;
    static constexpr int B_STAGE_BYTES = KPS * B_SUB_BYTES;
    static_assert(CG == 1 || (CG == 2 && EXT == 0 && (BN / NSPLIT / 2) % 8 == 0), "CTA-pair variant: no LoRA rows, B half % 8 == 0");
    static_assert(NSPLIT == 1 || (CG == 2 && EXT == 0), "split-N tiles are CTA-pair, no-LoRA only");
    static constexpr int SBO = 8 * ROW_BYTES;                  // 8-row swizzle atom
    static constexpr int LAYOUT = (BK == 64) ? 2 : 4;          // UMMA layout type: 128B / 64B swizzle
    static_assert(BK == 64 || BK == 32, "BK");
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int BUF_COLS = (BN + EXT + 31) / 32 * 32;
    static constexpr int NBUF = (2 * BUF_COLS <= 512) ? 2 : 1; // accumulator buffers (epilogue of tile i overlaps tile i+1 when 2)
    static constexpr int TMEM_COLS = (NBUF * BUF_COLS <= 32) ? 32 : (NBUF * BUF_COLS <= 64) ? 64 : (NBUF * BUF_COLS <= 128) ? 128
                                   : (NBUF * BUF_COLS <= 256) ? 256 : 512;
    static_assert(NBUF * BUF_COLS <= 512, "TMEM overflow");
    static_assert(MMA_N % 16 == 0 && MMA_N >= 16 && MMA_N <= 256, "invalid UMMA N");
    static_assert(BN % 32 == 0, "BN must be a multiple of the 32-column epilogue granule");
};

// Decode tile-local row r (0..127) into the global output row (or -1 if masked) and its row_bias group.
__device__ __forceinline__ void decode_row(const GemmParams& p, int m_blk, int r, int& m, int& grp) {
    if (p.a_mode == 0) {
        m = m_blk * BLOCK_M + r;
        if (m >= p.M) m = -1;
        grp = (m >= 0 && p.rows_per_group > 0) ? m / p.rows_per_group : 0;
    } else {
        int tw = m_blk % p.tiles_w;
        int th = (m_blk / p.tiles_w) % p.tiles_h;
        int tn = m_blk / (p.tiles_w * p.tiles_h);
        int dw = r % p.bw, dh = (r / p.bw) % p.bh, dn = r / (p.bw * p.bh);
        int n = tn * p.bn + dn, h = th * p.bh + dh, w = tw * p.bw + dw;
        if (n < p.n_img && h < p.Ho && w < p.Wo) {
            m = (n * p.Ho + h) * p.Wo + w;
            grp = (p.rows_per_group > 0) ? m / p.rows_per_group : 0;
        } else {
            m = -1;
            grp = 0;
        }
    }
}

template <int BN, int EXT, int BK, int CG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmD, const GemmParams p,
               const int num_stages) {
    pdl_launch_dependents();
    using Cfg = GemmCfg<BN, EXT, BK, CG>;
    const int cta_rank = (CG == 2) ? (int)cluster_ctarank() : 0;      // 0 = leader of the pair
    const int sched_cta = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int sched_n = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int sched_m = (CG == 2) ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks;
    const bool strip = (BK == 32) && p.strip != 0;
    const int A_STAGE_BYTES = strip ? STRIP_STAGE : Cfg::A_STAGE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve (base is 1024-aligned by the runtime for dynamic smem declared __align__(1024); re-align anyway)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + num_stages * A_STAGE_BYTES;
    const int b_bytes = p.b_resident ? p.num_k_blocks * Cfg::B_SUB_BYTES : num_stages * Cfg::B_STAGE_BYTES;
    uint8_t* smem_stage = smem_b + ((b_bytes + 1023) & ~1023);               // epilogue staging
    // LoRA: t (128 x 32 bf16, 64-byte rows, 64B swizzle) and the up matrix of the tile's n-block (BN x 32 bf16, same layout)
    uint8_t* smem_text = smem_stage + EPI_STAGING_BYTES;
    uint8_t* smem_bext = smem_text + (EXT ? BLOCK_M * 64 : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bext + (EXT ? BN * 64 : 0));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + num_stages;
    uint64_t* tmem_full = bars + 2 * num_stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* b_full = tmem_empty + 2;
    uint64_t* ext_ready = b_full + 1;       // [2] t operand of the buffer's tile is in shared memory (4 converting warps)
    uint64_t* tmem_full2 = ext_ready + 2;   // [2] accumulator incl. the rank-r update is complete
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full2 + 2);

    // warp index made provably warp-uniform (shuffle from lane 0): the producer / MMA warps below run their loops with
    // all 32 lanes converged and elect one lane only around the TMA / tcgen05 instructions, so that ptxas keeps addresses,
    // descriptors and coordinates in UNIFORM registers.  (With the whole loop under `if (lane == 0)` every tcgen05.mma was
    // preceded by an ELECT + 5 x R2UR.BROADCAST "waterfall" that paced the tensor pipe at ~150 clk per MMA - round 2 probe,
    // profiles/r02_mma_rate.log - against a 72-128 clk floor.)
    const int warp_idx = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
#ifdef CLB_TIMELINE
    int tl_slot = 0;
#endif

    // Prologue.  The producer warp initialises the barriers and (single-CTA tiles) does not wait for the rest of the CTA: it
    // announces itself on a named barrier and starts streaming operands while warp 2 allocates TMEM, so the first TMA round trip (~1 k clk) overlaps the ~1.5 k clk of set-up instead of following it.
    if (warp_idx == 0) {
        if (lane == 0) {
#ifdef CLB_TIMELINE
            { const unsigned long long t = gtime(); atomicMin(&g_gt[p.dbg_id * 4], t); atomicMax(&g_gt[p.dbg_id * 4 + 1], t); }
#endif
            TL(1);   // kernel entry
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            if (EXT) tma_prefetch_desc(&tmE);
            if (p.tma_store) tma_prefetch_desc(&tmD);
            for (int i = 0; i < num_stages; ++i) {
                mbar_init(&full_bar[i], 1);
                mbar_init(&empty_bar[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&tmem_full[i], 1);
                mbar_init(&tmem_empty[i], CG * NUM_EPI_WARPS);   // pair: both CTAs' epilogues release the leader's barrier
                mbar_init(&ext_ready[i], 4);
                mbar_init(&tmem_full2[i], 1);
            }
            mbar_init(b_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
    }
    if (warp_idx == 2) {
        if (CG == 2) { tmem_alloc_cg2(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish_cg2(); }
        else { tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    pdl_wait();   // prologue above touched only smem / TMEM / kernel params
#ifdef CLB_NO_EARLY_PRODUCER
    constexpr bool kEarlyProducer = false;
#else
    constexpr bool kEarlyProducer = true;
#endif
    if (CG == 2 || !kEarlyProducer) {
        tc_fence_before();
        if (CG == 2) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrive / TMA credit
        else __syncthreads();
        tc_fence_after();
    } else if (warp_idx == 0) {
        __threadfence_block();
        named_bar_arrive(1, NUM_THREADS);   // barriers are initialised; the producer runs ahead
    } else {
        tc_fence_before();
        named_bar_sync(1, NUM_THREADS);
        tc_fence_after();
    }
    const uint32_t tmem_base = *tmem_ptr_smem;
    // register re-balancing between the warpgroups: the producer / MMA / allocator warps need few registers, the epilogue
    // warps hold two 32-column granules plus their row operands (128 x 56 + 256 x 224 <= 64 K registers)
    if (warp_idx < 4) {
#ifndef CLB_NO_SETMAXNREG
    setmaxnreg_dec<CLB_REGS_CTRL>();
#endif
    if (warp_idx == 0) {
        // ===================================================== TMA producer (whole warp converged, one elected lane issues)
        {
            int stage = 0;
            uint32_t phase = 0;
            TileIter ti(p, sched_cta, sched_n, sched_m);
            if (p.b_resident && ti.valid()) {
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(b_full, p.num_k_blocks * Cfg::B_SUB_BYTES);
                    for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                        uint8_t* sb = smem_b + kb * Cfg::B_SUB_BYTES;
                        tma_load_2d(&tmB, b_full, sb, kb * BK, ti.n_blk * BN);
                        if (EXT) tma_load_2d(&tmE, b_full, sb + BN * Cfg::ROW_BYTES, kb * BK, 0);
                    }
                }
                __syncwarp();
            }
            for (; ti.valid(); ti.next()) {
                const int m_blk = (CG == 2) ? ti.m_blk * 2 + cta_rank : ti.m_blk;
                const int n_blk = ti.n_blk;
                int tw = 0, th = 0, tn = 0;
                if (p.a_mode != 0) {
                    tw = m_blk % p.tiles_w;
                    th = (m_blk / p.tiles_w) % p.tiles_h;
                    tn = m_blk / (p.tiles_w * p.tiles_h);
                }
                if (lane == 0) TL(10);  // producer: tile start
                const int kb1 = ti.kb_end();
                if (strip) {
                    // one stage per tile: the halo rows h-1, h, h+1 of the tile's 128-pixel run (out-of-bounds pixels / rows read zero)
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&full_bar[stage], 3 * STRIP_ROWS * 64);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky)
                            tma_load_4d(&tmA, &full_bar[stage], smem_a + stage * A_STAGE_BYTES + ky * STRIP_SLOT, 0, tw * p.bw - 1, th * p.bh + ky - 1,
                                        tn * p.bn);
                    }
                    __syncwarp();
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                    continue;
                }
                for (int kb = ti.kb_begin(); kb < kb1; kb += Cfg::KPS) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (lane == 0 && kb < 10) TLK(11);   // producer: ring slot acquired
                    if (elect_one_sync()) {
                        if (CG == 2) {
                            // both CTAs' bytes are credited to the LEADER's full barrier (one arrival: the leader's producer)
                            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], p.b_resident ? A_STAGE_BYTES : Cfg::STAGE_BYTES);
                        }
                        // conv: 3x3 tap and channel block of the stage's first k-block (one division per stage, then counted up)
                        int tap = 0, cb = 0;
                        if (p.a_mode != 0) { tap = kb / p.cblocks; cb = kb - tap * p.cblocks; }
#pragma unroll
                        for (int sub = 0; sub < Cfg::KPS; ++sub) {
                            const int kbs = kb + sub;
                            uint8_t* sa = smem_a + stage * A_STAGE_BYTES + sub * Cfg::A_SUB_BYTES;
                            uint8_t* sb = smem_b + stage * Cfg::B_STAGE_BYTES + sub * Cfg::B_SUB_BYTES;
                            // A-operand coordinates of this k-block
                            int ca0 = kbs * BK, ca1 = m_blk * BLOCK_M, ca2 = 0, ca3 = 0, ca4 = 0;
                            if (p.a_mode != 0) {
                                const int ky = tap / 3, kx = tap - ky * 3;
                                if (p.a_mode == 1) {
                                    ca0 = cb * BK; ca1 = tw * p.bw + kx - 1; ca2 = th * p.bh + ky - 1; ca3 = tn * p.bn;
                                } else {
                                    const int iy = ky - p.pad_lo, ix = kx - p.pad_lo;
                                    ca0 = (ix & 1) * p.C + cb * BK; ca1 = tw * p.bw + (ix >> 1); ca2 = iy & 1; ca3 = th * p.bh + (iy >> 1);
                                    ca4 = tn * p.bn;
                                }
                                if (++cb == p.cblocks) { cb = 0; ++tap; }
                            }
                            if (CG == 2) {
                                const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
                                if (p.a_mode == 0) tma_load_2d_cg2(&tmA, fb, sa, ca0, ca1);
                                else if (p.a_mode == 1) tma_load_4d_cg2(&tmA, fb, sa, ca0, ca1, ca2, ca3);
                                else tma_load_5d_cg2(&tmA, fb, sa, ca0, ca1, ca2, ca3, ca4);
#pragma unroll
                                for (int hh = 0; hh < Cfg::NSPLIT; ++hh)
                                    tma_load_2d_cg2(&tmB, fb, sb + hh * Cfg::MMA_B_ROWS * Cfg::ROW_BYTES, kbs * BK,
                                                    n_blk * BN + hh * (BN / Cfg::NSPLIT) + cta_rank * Cfg::MMA_B_ROWS);
                            } else {
                                if (p.a_mode == 0) tma_load_2d(&tmA, &full_bar[stage], sa, ca0, ca1);
                                else if (p.a_mode == 1) tma_load_4d(&tmA, &full_bar[stage], sa, ca0, ca1, ca2, ca3);
                                else tma_load_5d(&tmA, &full_bar[stage], sa, ca0, ca1, ca2, ca3, ca4);
                                if (!p.b_resident) {
                                    tma_load_2d(&tmB, &full_bar[stage], sb, kbs * BK, n_blk * BN);
                                    if (EXT) tma_load_2d(&tmE, &full_bar[stage], sb + BN * Cfg::ROW_BYTES, kbs * BK, 0);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0 && kb < 10) TLK(12);   // producer: loads of this k-block issued
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================================== MMA issuer
        if (cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M * CG, Cfg::UMMA_N, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            TileIter ti(p, sched_cta, sched_n, sched_m);
            if (p.b_resident && ti.valid()) mbar_wait(b_full, 0);
            for (; ti.valid(); ti.next(), ++it) {
                const int buf = it % Cfg::NBUF;
                const uint32_t buf_phase = (it / Cfg::NBUF) & 1;
                if (lane == 0) TL(20);  // mma: waiting for a free accumulator
                mbar_wait(&tmem_empty[buf], buf_phase ^ 1);
                tc_fence_after();
                if (lane == 0) TL(21);  // mma: accumulator free
                const uint32_t d_tmem = tmem_base + buf * Cfg::BUF_COLS;
                const int kb0 = ti.kb_begin(), kb1 = ti.kb_end();
                if (strip) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) TL(22);
                    const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                    const uint32_t sb = smem_u32(smem_b);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            // tap (ky, kx): rows kx .. kx + 127 of halo row ky - the operand starts kx pixels (64-byte rows) into the strip
                            const uint32_t a0 = sa + (tap / 3) * STRIP_SLOT + (tap % 3) * 64;
                            const uint32_t bo = (p.strip_bo == 0) ? 0u : ((a0 >> 7) & (p.strip_bo == 1 ? 3u : 7u));
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                tc_mma_ss(d_tmem, make_smem_desc(a0 + k * 32, 16, Cfg::SBO, Cfg::LAYOUT, bo),
                                          make_smem_desc(sb + tap * Cfg::B_SUB_BYTES + k * 32, 16, Cfg::SBO, Cfg::LAYOUT), idesc, (tap | k) ? 1u : 0u);
                        }
                        tc_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                } else
                for (int kb = kb0; kb < kb1; kb += Cfg::KPS) {
                    mbar_wait(&full_bar[stage], phase);      // all 32 lanes poll: keeps the loop convergent
                    tc_fence_after();
                    if (lane == 0) { if (kb == kb0) TL(22); else if (kb < kb0 + 10) TLK(24); }
                    const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                    const uint32_t sb = smem_u32(smem_b) + (p.b_resident ? kb * Cfg::B_SUB_BYTES : stage * Cfg::B_STAGE_BYTES);
                    const uint32_t first = (kb != kb0) ? 1u : 0u;
                    if (elect_one_sync()) {
#ifdef CLB_TIMELINE
                        for (int rep = 0; rep < p.dbg_reps; ++rep) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t adesc = make_smem_desc(sa + k * 32, 16, Cfg::SBO, Cfg::LAYOUT);
#pragma unroll
                                for (int hh = 0; hh < Cfg::NSPLIT; ++hh) {
                                    const uint64_t bdesc = make_smem_desc(sb + hh * Cfg::MMA_B_ROWS * Cfg::ROW_BYTES + k * 32, 16, Cfg::SBO, Cfg::LAYOUT);
                                    if (CG == 2) tc_mma_ss_cg2(d_tmem + hh * Cfg::MMA_N, adesc, bdesc, idesc, 1u);
                                    else tc_mma_ss(d_tmem + hh * Cfg::MMA_N, adesc, bdesc, idesc, 1u);
                                }
                            }
                        }
#endif
#pragma unroll
                        for (int sub = 0; sub < Cfg::KPS; ++sub) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                const uint64_t adesc = make_smem_desc(sa + sub * Cfg::A_SUB_BYTES + k * 32, 16, Cfg::SBO, Cfg::LAYOUT);
#pragma unroll
                                for (int hh = 0; hh < Cfg::NSPLIT; ++hh) {
                                    const uint64_t bdesc = make_smem_desc(sb + sub * Cfg::B_SUB_BYTES + hh * Cfg::MMA_B_ROWS * Cfg::ROW_BYTES + k * 32, 16,
                                                                          Cfg::SBO, Cfg::LAYOUT);
                                    const uint32_t acc = (sub != 0 || k != 0) ? 1u : first;
                                    if (CG == 2) tc_mma_ss_cg2(d_tmem + hh * Cfg::MMA_N, adesc, bdesc, idesc, acc);
                                    else tc_mma_ss(d_tmem + hh * Cfg::MMA_N, adesc, bdesc, idesc, acc);
                                }
                            }
                        }
                        // smem slot is free once these MMAs retire (pair: in both CTAs)
                        if (CG == 2) tc_commit_cg2(&empty_bar[stage], 3); else tc_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (lane == 0 && kb < kb0 + 10) TLK(25);   // mma: k-block issued + committed
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                // accumulator complete -> epilogue (pair: each CTA drains its own 128 rows)
                if (elect_one_sync()) {
                    if (CG == 2) tc_commit_cg2(&tmem_full[buf], 3); else tc_commit(&tmem_full[buf]);
                }
                __syncwarp();
                if (EXT) {
                    // rank-r update on the tensor core: the epilogue turned the 16 extra accumulator columns into the K-major
                    // operand t (bf16 hi | lo); D[:, 0:BN] += t * up^T with K = 16 per instruction
                    mbar_wait(&ext_ready[buf], buf_phase);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        constexpr uint32_t idesc_x = make_idesc_bf16(BLOCK_M, BN, 0, 0);
                        const uint32_t ta = smem_u32(smem_text), tb = smem_u32(smem_bext);
                        const int nk = (p.lora_rp > 4) ? 2 : 1;
                        for (int k = 0; k < nk; ++k)
                            tc_mma_ss(d_tmem, make_smem_desc(ta + k * 32, 16, 512, 4), make_smem_desc(tb + k * 32, 16, 512, 4), idesc_x, 1u);
                        tc_commit(&tmem_full2[buf]);
                    }
                    __syncwarp();
                }
                if (lane == 0) TL(23);  // mma: all MMAs of the tile issued
            }
        }
    }
    } else {
#ifndef CLB_NO_SETMAXNREG
        setmaxnreg_inc<CLB_REGS_EPI>();
#endif
        // ===================================================== epilogue
        const int ew = warp_idx - 4;              // 0..7
        const int quad = warp_idx & 3;            // TMEM lane quadrant this warp may access
        const int half = ew >> 2;                 // the two warps of a quadrant split the column granules
        const uint32_t stg_addr = smem_u32(smem_stage + ew * 32 * STAGE_ROW_BYTES);
        const int rp = p.lora_rp;
        int it = 0;
        int epi_gran = 0;                          // granules this warp has stored through TMA (staging buffer = parity)
        int bext_n_blk = -1;                       // n-block whose up matrix is in smem_bext
        for (TileIter ti(p, sched_cta, sched_n, sched_m); ti.valid(); ti.next(), ++it) {
            const int m_blk = (CG == 2) ? ti.m_blk * 2 + cta_rank : ti.m_blk;
            const int n_blk = ti.n_blk;
            const int buf = it % Cfg::NBUF;
            const uint32_t buf_phase = (it / Cfg::NBUF) & 1;
            // rows: phase-1 thread owns row (quad*32 + lane); phase-2 lane handles rows (lane>>3) + 4*i
            int my_m, my_grp;
            decode_row(p, m_blk, quad * 32 + lane, my_m, my_grp);
            const uint32_t t_base = tmem_base + (uint32_t(quad * 32) << 16) + buf * Cfg::BUF_COLS;

            if (EXT && half == 0) {
                // ---- LoRA: the four half-0 warps (one per TMEM lane quadrant, 128 threads) prepare the operands of the rank-r MMA
                // up matrix of this n-block -> smem_bext (only when the n-block changed; the previous tile's rank-r MMA has
                // retired: its commit preceded this tile's main-loop commit).  Row n: 64 bytes = 4 chunks of 8 bf16,
                // rank <= 4: [hi hi | lo lo | 0 | 0], rank <= 8: [hi | hi | lo | lo]; chunk c of row n sits at c ^ ((n >> 1) & 3).
                if (n_blk != bext_n_blk) {
                    bext_n_blk = n_blk;
                    for (int n = quad * 32 + lane; n < BN; n += 128) {
                        const int ng = n_blk * BN + n;
                        float u[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) u[j] = 0.f;
                        if (ng < p.N) {
                            const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.lora_up + (long long)ng * rp));
                            u[0] = a4.x; u[1] = a4.y; u[2] = a4.z; u[3] = a4.w;
                            if (rp > 4) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.lora_up + (long long)ng * rp + 4));
                                u[4] = b4.x; u[5] = b4.y; u[6] = b4.z; u[7] = b4.w;
                            }
                        }
                        uint32_t hi[4], lo[4];     // bf16x2 pairs (j, j+1)
#pragma unroll
                        for (int j = 0; j < 4; ++j) split_bf16x2(u[2 * j], u[2 * j + 1], hi[j], lo[j]);
                        uint4 c0, c1, c2, c3;
                        if (rp <= 4) {
                            c0 = make_uint4(hi[0], hi[1], hi[0], hi[1]); c1 = make_uint4(lo[0], lo[1], lo[0], lo[1]);
                            c2 = make_uint4(0u, 0u, 0u, 0u); c3 = c2;
                        } else {
                            c0 = make_uint4(hi[0], hi[1], hi[2], hi[3]); c1 = c0;
                            c2 = make_uint4(lo[0], lo[1], lo[2], lo[3]); c3 = c2;
                        }
                        const uint32_t row = smem_u32(smem_bext) + n * 64;
                        const int sw = (n >> 1) & 3;
                        st_shared_u4(row + ((0 ^ sw) << 4), c0); st_shared_u4(row + ((1 ^ sw) << 4), c1);
                        st_shared_u4(row + ((2 ^ sw) << 4), c2); st_shared_u4(row + ((3 ^ sw) << 4), c3);
                    }
                }
                if (ew == 0 && lane == 0) TL(30);  // epilogue: waiting for the accumulator
                mbar_wait(&tmem_full[buf], buf_phase);
                tc_fence_after();
                if (ew == 0 && lane == 0) TL(31);  // epilogue: accumulator ready
                // t = x A^T from the 16 extra columns (hi + lo halves of the down matrix), + t_add, -> t_out, * scale
                float tl[8];
                {
                    uint32_t e[16];
                    tmem_ld_32x16(t_base + BN, e);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 8; ++j) tl[j] = __uint_as_float(e[j]) + __uint_as_float(e[j + 8]);
                }
                if (my_m >= 0) {
                    if (p.t_add != nullptr) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < rp) tl[j] += p.t_add[(long long)my_m * rp + j];
                    }
                    if (p.t_out != nullptr && n_blk == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < rp) p.t_out[(long long)my_m * rp + j] = tl[j];
                    }
                }
                {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) split_bf16x2(tl[2 * j] * p.lora_scale, tl[2 * j + 1] * p.lora_scale, hi[j], lo[j]);
                    // row r = quad*32 + lane of the K-major t operand: rank <= 4: [hi lo | hi lo | 0 | 0]  (k 0-3 t_hi, 4-7 t_lo, 8-11 t_hi,
                    // 12-15 t_lo against up [hi hi | lo lo]), rank <= 8: [hi | lo | hi | lo] against up [hi | hi | lo | lo]
                    uint4 c0, c1, c2, c3;
                    if (rp <= 4) {
                        c0 = make_uint4(hi[0], hi[1], lo[0], lo[1]); c1 = c0;
                        c2 = make_uint4(0u, 0u, 0u, 0u); c3 = c2;
                    } else {
                        c0 = make_uint4(hi[0], hi[1], hi[2], hi[3]); c1 = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        c2 = c0; c3 = c1;
                    }
                    const int r = quad * 32 + lane;
                    const uint32_t row = smem_u32(smem_text) + r * 64;
                    const int sw = (r >> 1) & 3;
                    st_shared_u4(row + ((0 ^ sw) << 4), c0); st_shared_u4(row + ((1 ^ sw) << 4), c1);
                    st_shared_u4(row + ((2 ^ sw) << 4), c2); st_shared_u4(row + ((3 ^ sw) << 4), c3);
                }
                fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor core's async-proxy reads
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ext_ready[buf]);
                if (ew == 0 && lane == 0) TL(38);  // epilogue: LoRA operands handed to the MMA warp
            }
            if (!EXT) {
                if (ew == 0 && lane == 0) TL(30);  // epilogue: waiting for the accumulator
                mbar_wait(&tmem_full[buf], buf_phase);
                tc_fence_after();
                if (ew == 0 && lane == 0) TL(31);  // epilogue: accumulator ready
            } else {
                mbar_wait(&tmem_full2[buf], buf_phase);
                tc_fence_after();
                if (ew == 0 && lane == 0) TL(39);  // epilogue: accumulator incl. rank-r update ready
            }

            if (p.tma_store == 4) {
                // timing experiment only (CLB_GEMM_EPI_DEBUG=4): no epilogue work at all, the accumulator is dropped
            } else if (p.tma_store) {
                // ---- bf16 output, thread == output row: accumulator (+ LoRA + bias + row bias + residual) in registers,
                //      packed to bf16, staged as a 32 x 64 B sub-tile (64B-swizzled, 4 conflict-free 16-byte stores per lane)
                //      and written by ONE TMA store per granule; two staging buffers per warp keep a store in flight while
                //      the next granule is produced.  Row-wise operands (residual) are fetched before the TMEM load is
                //      waited on, so their latency overlaps it.
                int sc1 = 0, sc2 = 0, sc3 = 0;            // store coordinates of this warp's 32 rows
                if (p.a_mode == 0) {
                    sc1 = m_blk * BLOCK_M + quad * 32;
                } else {
                    const int r0 = quad * 32;
                    const int tw = m_blk % p.tiles_w, th = (m_blk / p.tiles_w) % p.tiles_h, tn = m_blk / (p.tiles_w * p.tiles_h);
                    sc1 = tw * p.bw + r0 % p.bw;
                    sc2 = th * p.bh + (r0 / p.bw) % p.bh;
                    sc3 = tn * p.bn + r0 / (p.bw * p.bh);
                }
                const bool row_ok = my_m >= 0;
                const __nv_bfloat16* res_row = (p.residual != nullptr && row_ok) ? p.residual + (long long)my_m * p.ldr : nullptr;
                const float* rb_row = (p.row_bias != nullptr && row_ok) ? p.row_bias + (long long)my_grp * p.ld_rb : nullptr;
                for (int g = half; g < BN / 32; g += 2) {
                    const int col0 = n_blk * BN + g * 32;
                    if (col0 >= p.N) break;                   // (warp-uniform)
                    uint4 rr[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        rr[j] = make_uint4(0u, 0u, 0u, 0u);
                        if (res_row != nullptr && col0 + 8 * j < p.N) rr[j] = *reinterpret_cast<const uint4*>(res_row + col0 + 8 * j);
                    }
                    uint32_t v[32];
                    tmem_ld_32x32(t_base + g * 32, v);
                    float f[32];
                    // column-wise operands while the TMEM load is in flight
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (col0 + 4 * j < p.N) {
                            if (p.bias != nullptr) bz = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 4 * j));
                            if (rb_row != nullptr) {
                                const float4 rb = *reinterpret_cast<const float4*>(rb_row + col0 + 4 * j);
                                bz.x += rb.x; bz.y += rb.y; bz.z += rb.z; bz.w += rb.w;
                            }
                        }
                        f[4 * j] = bz.x; f[4 * j + 1] = bz.y; f[4 * j + 2] = bz.z; f[4 * j + 3] = bz.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 a0 = unpack_bf16x2(rr[j].x), a1 = unpack_bf16x2(rr[j].y), a2 = unpack_bf16x2(rr[j].z),
                                     a3 = unpack_bf16x2(rr[j].w);
                        f[8 * j] += a0.x; f[8 * j + 1] += a0.y; f[8 * j + 2] += a1.x; f[8 * j + 3] += a1.y;
                        f[8 * j + 4] += a2.x; f[8 * j + 5] += a2.y; f[8 * j + 6] += a3.x; f[8 * j + 7] += a3.y;
                    }
                    tc_wait_ld();
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] += __uint_as_float(v[c]);
                    if (p.tma_store >= 2) {
                        // thread == row, four 16-byte global stores per granule straight from registers
                        if (row_ok && p.tma_store == 2) {
                            __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)my_m * p.ldd + col0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (col0 + 8 * j < p.N) {
                                    uint4 o;
                                    o.x = pack_bf16x2(f[8 * j], f[8 * j + 1]); o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                                    o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]); o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                                    *reinterpret_cast<uint4*>(orow + 8 * j) = o;
                                }
                            }
                        } else if (p.tma_store == 3 && f[0] == 12345.678f) {
                            *reinterpret_cast<float*>(p.out) = f[1] + f[31];      // timing experiment: compute, never store
                        }
                        ++epi_gran;
                        continue;
                    }
                    // the store issued two granules ago has finished reading this staging buffer
                    const uint32_t sbuf = stg_addr + (uint32_t)(epi_gran & 1) * 2048u;
                    if (lane == 0) bulk_wait_group_read<1>();
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_bf16x2(f[8 * j], f[8 * j + 1]); o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                        o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]); o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                        st_shared_u4(sbuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), o);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.a_mode == 0) tma_store_2d(&tmD, sbuf, col0, sc1);
                        else tma_store_4d(&tmD, sbuf, col0, sc1, sc2, sc3);
                        bulk_commit_group();
                    }
                    ++epi_gran;
                }
            } else {
                // ---- per-lane store epilogue (fp32 outputs, split-K partials, smem-resident short-K tiles).  The warp's granules
                //      (g = half, half + 2, ...) are software-pipelined: the TMEM load of the next granule and the global loads of
                //      this one (residual, row bias) are in flight while the LoRA update / transpose / stores of this one run -
                //      with two epilogue warps per scheduler the old load -> wait -> use chain ran at ~0.3 IPC.
                auto gran_ok = [&](int g) { return g < BN / 32 && n_blk * BN + g * 32 < p.N; };   // warp-uniform
                const bool bf16_stage = p.epi_bf16 != 0;
                auto process = [&](uint32_t (&v)[32], const int g) {
                    const int col0 = n_blk * BN + g * 32;   // first global column of this granule
                    // phase-2 coordinates and operands first (lane -> row (lane>>3) + 4*i, chunk lane&7): loads issued before use
                    const int ch = lane & 7;
                    const int n0 = col0 + ch * 4;
                    const bool col_ok = n0 < p.N;  // N % 4 == 0 is enforced on the host
                    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (col_ok && p.bias != nullptr && p.splits == 1) bz = *reinterpret_cast<const float4*>(p.bias + n0);
                    int mrow[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int m_r = __shfl_sync(0xffffffffu, my_m, (lane >> 3) + 4 * i);   // every lane takes part in the shuffle
                        mrow[i] = col_ok ? m_r : -1;
                    }
                    uint2 rr[8];
                    const bool has_rb = p.row_bias != nullptr, has_res = p.residual != nullptr;
                    if (has_res) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            rr[i] = make_uint2(0u, 0u);
                            if (mrow[i] >= 0) rr[i] = *reinterpret_cast<const uint2*>(p.residual + (long long)mrow[i] * p.ldr + n0);
                        }
                    }
                    if (ew == 0 && lane == 0) TL(33);  // epilogue: granule in registers, phase-2 loads issued
                    float f[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) f[c] = __uint_as_float(v[c]);
                    // ---- phase 1 -> smem (row = lane, 8 chunks of 16 B, XOR swizzle keeps both phases conflict-free)
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        st_shared_f4(stg_addr + lane * STAGE_ROW_BYTES + ((j ^ (lane & 7)) << 4),
                                     make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                    __syncwarp();
                    if (ew == 0 && lane == 0) TL(35);  // epilogue: staged
                    // ---- phase 2: coalesced bias / residual / store
                    float4 q[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = (lane >> 3) + 4 * i;
                        q[i] = ld_shared_f4(stg_addr + r * STAGE_ROW_BYTES + ((ch ^ (r & 7)) << 4));
                    }
                    if (has_rb) {      // (conv time-embedding bias: rare, loaded late to keep the register budget of the pipelined loop)
                        float4 rb[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int grp = __shfl_sync(0xffffffffu, my_grp, (lane >> 3) + 4 * i);
                            rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (mrow[i] >= 0) rb[i] = *reinterpret_cast<const float4*>(p.row_bias + (long long)grp * p.ld_rb + n0);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) { q[i].x += rb[i].x; q[i].y += rb[i].y; q[i].z += rb[i].z; q[i].w += rb[i].w; }
                    }
                    if (has_res) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 a = unpack_bf16x2(rr[i].x), b = unpack_bf16x2(rr[i].y);
                            q[i].x += a.x; q[i].y += a.y; q[i].z += b.x; q[i].w += b.y;
                        }
                    }
                    if (ew == 0 && lane == 0) TL(36);  // epilogue: phase-2 operands applied
                    if (p.splits > 1) {
                        float* part = p.split_ws + (long long)ti.split * p.M * p.N;   // host cleared bias/row_bias/residual
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (mrow[i] >= 0) *reinterpret_cast<float4*>(part + (long long)mrow[i] * p.N + n0) = q[i];
                    } else if (p.out_fp32) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (mrow[i] >= 0) {
                                float4 o = make_float4(q[i].x + bz.x, q[i].y + bz.y, q[i].z + bz.z, q[i].w + bz.w);
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)mrow[i] * p.ldd + n0) = o;
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (mrow[i] >= 0) {
                                uint2 o;
                                o.x = pack_bf16x2(q[i].x + bz.x, q[i].y + bz.y);
                                o.y = pack_bf16x2(q[i].z + bz.z, q[i].w + bz.w);
                                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)mrow[i] * p.ldd + n0) = o;
                            }
                        }
                    }
                };
                uint32_t va[32], vb[32];
                int g = half;
                if (bf16_stage) {
                    // ---- nothing to add after the accumulator (no bias / residual / LoRA): round to bf16 first, transpose 64-byte
                    //      rows through the staging buffer (half the shared-memory traffic of the fp32 transpose) and store 16 bytes
                    //      per lane (8 rows x 64 B per instruction)
                    auto process16 = [&](uint32_t (&v)[32], const int gg) {
                        const int col0 = n_blk * BN + gg * 32;
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 o;
                            o.x = pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]));
                            o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                            o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                            o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                            st_shared_u4(stg_addr + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), o);
                        }
                        __syncwarp();
                        const int cq = lane & 3;
                        const int n0 = col0 + cq * 8;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = (lane >> 2) + 8 * i;
                            const uint4 o = ld_shared_u4(stg_addr + r * 64 + ((cq ^ ((r >> 1) & 3)) << 4));
                            const int m_r = __shfl_sync(0xffffffffu, my_m, r);
                            if (m_r >= 0 && n0 < p.N)
                                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)m_r * p.ldd + n0) = o;
                        }
                    };
                    if (gran_ok(g)) tmem_ld_32x32(t_base + g * 32, va);
                    while (gran_ok(g)) {
                        tc_wait_ld();
                        if (gran_ok(g + 2)) tmem_ld_32x32(t_base + (g + 2) * 32, vb);
                        process16(va, g);
                        g += 2;
                        if (!gran_ok(g)) break;
                        tc_wait_ld();
                        if (gran_ok(g + 2)) tmem_ld_32x32(t_base + (g + 2) * 32, va);
                        process16(vb, g);
                        g += 2;
                    }
                } else {
                if (gran_ok(g)) tmem_ld_32x32(t_base + g * 32, va);
                while (gran_ok(g)) {
                    tc_wait_ld();
                    if (gran_ok(g + 2)) tmem_ld_32x32(t_base + (g + 2) * 32, vb);
                    process(va, g);
                    g += 2;
                    if (!gran_ok(g)) break;
                    tc_wait_ld();
                    if (gran_ok(g + 2)) tmem_ld_32x32(t_base + (g + 2) * 32, va);
                    process(vb, g);
                    g += 2;
                }
                }
            }
            if (ew == 0 && lane == 0) TL(37);  // epilogue: all stores issued
            // all TMEM reads of this buffer are complete (tc_wait_ld above) -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[buf]), 0));
                else mbar_arrive(&tmem_empty[buf]);
            }
            if (ew == 0 && lane == 0) TL(32);  // epilogue: tile done
        }
        if (p.tma_store && lane == 0) bulk_wait_group_read<0>();   // staging smem must outlive the last TMA stores
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();   // the leader's MMAs read the peer's smem: nobody leaves before both epilogues end
    else __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
        else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
#ifdef CLB_TIMELINE
        if (lane == 0) { const unsigned long long t = gtime(); atomicMin(&g_gt[p.dbg_id * 4 + 2], t); atomicMax(&g_gt[p.dbg_id * 4 + 3], t); }
#endif
    }
}

// out = epilogue(sum_s ws[s]) for the split-K path: one thread per 4 output columns.
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ ws, int splits, int M, int N, const float* __restrict__ bias,
                     const float* __restrict__ row_bias, int rows_per_group, long long ld_rb,
                     const __nv_bfloat16* __restrict__ residual, long long ldr, void* __restrict__ out, long long ldd,
                     int out_fp32) {
    pdl_launch_dependents();
    pdl_wait();
    const int nq = N >> 2;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)M * nq) return;
    const int m = (int)(idx / nq);
    const int n0 = (int)(idx - (long long)m * nq) * 4;
    const long long MN = (long long)M * N;
    const float* src = ws + (long long)m * N + n0;
    float4 acc = *reinterpret_cast<const float4*>(src);
    for (int s = 1; s < splits; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(src + s * MN);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (row_bias != nullptr) {
        const float4 v = *reinterpret_cast<const float4*>(row_bias + (long long)(m / rows_per_group) * ld_rb + n0);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (residual != nullptr) {
        const uint2 rr = *reinterpret_cast<const uint2*>(residual + (long long)m * ldr + n0);
        const float2 a = unpack_bf16x2(rr.x), b = unpack_bf16x2(rr.y);
        acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
    }
    if (bias != nullptr) {
        const float4 v = *reinterpret_cast<const float4*>(bias + n0);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (out_fp32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (long long)m * ldd + n0) = acc;
    } else {
        uint2 o;
        o.x = pack_bf16x2(acc.x, acc.y);
        o.y = pack_bf16x2(acc.z, acc.w);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (long long)m * ldd + n0) = o;
    }
}

// =====================================================================================================
// host side
// =====================================================================================================

template <int BN, int EXT, int BK = 64, int CG = 1>
static int launch_gemm(const CUtensorMap& tA_in, const CUtensorMap& tB, const CUtensorMap& tE, const CUtensorMap& tD,
                       const GemmParams& p_in, cudaStream_t stream, const CUtensorMap* tA_strip = nullptr) {
    const CUtensorMap* tA_sel = &tA_in;
    using Cfg = GemmCfg<BN, EXT, BK, CG>;
    GemmParams p = p_in;
    const int ext_bytes = EXT ? (BLOCK_M + BN) * 64 : 0;      // LoRA: t and up operands of the rank-r MMA
    const int fixed = 1024 /*align slack*/ + EPI_STAGING_BYTES + ext_bytes + 512 /*barriers: 2 * stages + 9 words*/;
    int stages, smem_bytes;
    int grid = num_sms();
    // B-resident mode: the weight tile of one n-block fits next to >= 3 A stages and every CTA re-uses it >= 3 times
    const int b_res_bytes = (p.num_k_blocks * Cfg::B_SUB_BYTES + 1023) & ~1023;
    const int a_room = 232448 - fixed - b_res_bytes;
    const int grid_res = (grid / p.num_n_blocks) * p.num_n_blocks;
    // (the 32-channel convs of the hint encoder qualify too: all nine taps of a 32 x 32 filter are 18 KB)
    if (CG == 1 && p.splits == 1 && (p.a_mode == 0 || BK == 32) && a_room >= 3 * Cfg::A_STAGE_BYTES && grid_res > 0 && grid_res * 10 >= grid * 9 &&
        p.num_m_blocks >= 3 * (grid_res / p.num_n_blocks)) {
        p.b_resident = 1;
        // short-K resident tiles hand a finished accumulator to the epilogue every ~2 k cycles: measured (round 2) the
        // per-lane store epilogue keeps up with that better than two TMA stores in flight per warp (18.5 vs 20.1 us at
        // 32768 x 320 x 320), while the TMA epilogue wins wherever the epilogue is exposed (convs, 256 x 320 tiles)
        if (p.tma_store == 1) p.tma_store = 0;
        int a_stage = Cfg::A_STAGE_BYTES;
        if (BK == 32 && tA_strip != nullptr && p.num_n_blocks == 1 && a_room >= 3 * STRIP_STAGE) {
            p.strip = 1; a_stage = STRIP_STAGE; tA_sel = tA_strip;
        }
        stages = a_room / a_stage;
        if (stages > 8) stages = 8;
        smem_bytes = fixed + b_res_bytes + stages * a_stage;
        grid = grid_res;
    } else {
        p.b_resident = 0;
        stages = (232448 - fixed) / Cfg::STAGE_BYTES;
        if (stages > 24) stages = 24;   // small stages (32-channel convs: 10 KB) need a deep ring to keep ~190 KB in flight
        if (stages < 2) return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: not enough shared memory for 2 stages");
        smem_bytes = fixed + stages * Cfg::STAGE_BYTES;
        const int m_units = (CG == 2) ? (p.num_m_blocks + 1) / 2 : p.num_m_blocks;
        const int num_tiles = m_units * p.num_n_blocks * p.splits;
        if (CG == 2) grid /= 2;                       // scheduling units are CTA pairs
        if (grid > num_tiles) grid = num_tiles;
        grid *= CG;
    }
    const GemmParams full = p;
    if (p.splits > 1) { p.bias = nullptr; p.row_bias = nullptr; p.residual = nullptr; }
    const CUtensorMap& tA = *tA_sel;
    static bool attr_done = false;
    if (!attr_done) {
        CL_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EXT, BK, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
        attr_done = true;
    }
    if (CG == 2) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled(CLB_FAMILY) ? 2 : 1;
        CL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EXT, BK, CG>, tA, tB, tE, tD, p, stages));
    } else {
        launch_k(gemm_tc_kernel<BN, EXT, BK, CG>, grid, NUM_THREADS, smem_bytes, stream, tA, tB, tE, tD, p, stages);
    }
    count_launch();
    if (p.splits > 1) {
        const long long quads = (long long)p.M * (p.N / 4);
        launch_k(splitk_finish_kernel, (unsigned)((quads + 255) / 256), 256, 0, stream, 
            p.split_ws, p.splits, p.M, p.N, full.bias, full.row_bias, full.rows_per_group, full.ld_rb, full.residual,
            full.ldr, full.out, full.ldd, full.out_fp32);
        count_launch();
    }
    CL_CUDA_CHECK(cudaGetLastError());
    return CL_OK;
}

}  // namespace clb

using namespace clb;

#ifdef CLB_TIMELINE
extern "C" int cl_debug_timeline(unsigned long long* host_buf, unsigned int* host_n, int reset) {
    if (reset) {
        static unsigned long long z[160 * 256 * 2];
        memset(z, 0, sizeof(z));
        cudaMemcpyToSymbol(clb::g_tl, z, sizeof(z));
        unsigned long long g[64 * 4];
        for (int i = 0; i < 64; ++i) { g[i * 4] = ~0ull; g[i * 4 + 1] = 0; g[i * 4 + 2] = ~0ull; g[i * 4 + 3] = 0; }
        cudaMemcpyToSymbol(clb::g_gt, g, sizeof(g));
        return 0;
    }
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_buf, clb::g_tl, sizeof(unsigned long long) * 160 * 256 * 2);
    for (int i = 0; i < 160; ++i) host_n[i] = 240;       // fixed slots; unused ones have tag 0
    return 0;
}
extern "C" int cl_debug_gtimes(unsigned long long* host_buf) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_buf, clb::g_gt, sizeof(unsigned long long) * 64 * 4);
    return 0;
}
#endif

// Output-tile plan: block_n and the number of K splits.  Cost model (the 1-CTA mainloop is bound by the L2->SM stream of
// A and B rows): a CTA moves (128 + bn) rows per k-block and the kernel runs `waves` rounds of tiles, so
//   cost = waves(tiles * splits) * (128 + bn + 32) * ceil(nkb / splits)          (+32: per-tile epilogue / ramp)
// Splitting K is only considered when the (m, n) tiles leave more than half of the SMs idle and every split keeps
// >= 8 k-blocks.  The caller provides split_ws = splits * M * N floats (cl_gemm_split_hint tells how many).
struct TilePlan { int bn, splits; };

static inline int gemm_block_k(const cl_gemm_args* a) { return (a->a_mode != 0 && a->C % 64 != 0) ? 32 : 64; }

// CTA pairs (cta_group::2) serve everything except the LoRA epilogue, the 32-channel convs and the smem-resident small-K
// case; CLB_GEMM_2CTA=0 falls back to single-CTA tiles.
static int gemm_cta_group(const cl_gemm_args* a, int BK, int num_m_blocks) {
    static const int pair_enabled = [] { const char* e = getenv("CLB_GEMM_2CTA"); return (e && e[0] == '0') ? 0 : 1; }();
    // smallest K that runs as CTA pairs.  Round-2 measurement: a CTA ingests ~64 B/clk through TMA (about one 128-byte row per
    // 2 clk), so a 1-CTA 128 x BN tile needs (128 + BN) rows per k-block against 2*BN clk of MMA work and is load-bound by 1.5x or
    // more, while a pair stages 128 + BN/2 rows per CTA - balanced from BN = 256 up.  CLB_GEMM_PAIR_MINK overrides.
    static const int pair_min_k = [] { const char* e = getenv("CLB_GEMM_PAIR_MINK"); return e ? atoi(e) : 2048; }();
    const bool small_k_resident = a->a_mode == 0 && a->K <= 320 && a->N % 160 == 0 && a->M >= 8192;
    return (pair_enabled && a->lora_up == nullptr && BK == 64 && !small_k_resident && num_m_blocks >= 2 && a->K >= pair_min_k) ? 2 : 1;
}

static TilePlan plan_tiles(const cl_gemm_args* a, int BK, int num_m_blocks, bool allow_split) {
    const bool lora = a->lora_up != nullptr;
    const int nkb = (a->K + BK - 1) / BK;
    const int cg = gemm_cta_group(a, BK, num_m_blocks);
    const int sms = num_sms() / cg;                        // scheduling units: CTAs or CTA pairs
    TilePlan best = {0, 1};
#ifdef CLB_TIMELINE
    if (a->block_n != 0) return TilePlan{a->block_n, 1};     // probe builds: any instantiated width, as given
#endif
    int fixed_bn = 0;                                         // tile width imposed by the caller / the resident rule
    if (a->block_n != 0) fixed_bn = a->block_n;
    else if (!lora && a->a_mode == 0 && a->K <= 320 && a->N % 160 == 0 && a->M >= 8192) fixed_bn = 160;  // weights stay smem-resident
    long long best_cost = -1;
    static const int wide_enabled = [] { const char* e = getenv("CLB_GEMM_BN320"); return (e && e[0] == '0') ? 0 : 1; }();
    static const int cands[6] = {320, 256, 160, 128, 64, 32};
    for (int ci = 0; ci < 6; ++ci) {
        const int bn = cands[ci];
        if (fixed_bn != 0 && bn != fixed_bn) continue;
        if (bn == 320 && (cg != 2 || lora || !wide_enabled || a->N % 320 != 0)) continue;   // CTA-pair, no-LoRA tile
        if (fixed_bn == 0) {
            if (a->N % bn != 0) continue;
            if (lora && bn > 160) continue;
            if (BK == 32 ? (bn > 128) : (bn < 64)) continue;   // instantiated variants
            if (lora && bn < 64) continue;
        }
        const int tiles = ((num_m_blocks + cg - 1) / cg) * ((a->N + bn - 1) / bn);
        int splits = 1;
        if (allow_split && !lora && tiles * 2 <= sms) {
            splits = sms / tiles;
            if (splits > nkb / 8) splits = nkb / 8;
            if (splits > 16) splits = 16;
            if (splits < 2) splits = 1;
        }
        int kbps = (nkb + splits - 1) / splits;
        if (BK == 32) kbps = (kbps + 2) / 3 * 3;       // three k-blocks share a pipeline stage in the 32-channel variant
        splits = (nkb + kbps - 1) / kbps;
        const long long waves = ((long long)tiles * splits + sms - 1) / sms;
        // Cost in clocks (round-2 measurements): a CTA ingests ~one 128-byte operand row per 2 clk through TMA, so a k-block costs
        // max(2 * rows staged by the CTA, MMA time = 2 * bn) + ~100; a tile additionally pays its exposed epilogue (the 320-column
        // tile has ONE accumulator buffer: ~3000 clk, otherwise ~300), and a split-K plan its finishing launch (~12000 clk).
        // (Round 1 compared only the staged rows - and, through a bookkeeping bug, never looked past the first valid width.)
        const long long t_load = 2LL * (128 + bn / cg), t_mma = 2LL * bn;
        const long long t_kb = (t_load > t_mma ? t_load : t_mma) + 100;
        const long long cost = waves * (t_kb * kbps + (bn == 320 ? 3000 : 300)) + (splits > 1 ? 12000 : 0);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = {bn, splits}; }
    }
    if (best_cost < 0) {   // no candidate divides N: tail tiles
        if (lora) best.bn = (a->N % 128 == 0) ? 128 : 160;
        else if (a->N <= 32 && BK == 32) best.bn = 32;
        else if (a->N <= 64) best.bn = 64;
        else if (a->N <= 128 || BK == 32) best.bn = 128;
        else best.bn = 160;
        best.splits = 1;
    }
    return best;
}

extern "C" int cl_gemm_split_hint(const cl_gemm_args* a) {
    if (a == nullptr || a->lora_up != nullptr || a->M <= 0 || a->N <= 0 || a->K <= 0) return 1;
    // conv tiles may pad M when the image does not tile evenly; cl_gemm re-plans with the exact count and clamps
    return plan_tiles(a, gemm_block_k(a), (a->M + BLOCK_M - 1) / BLOCK_M, true).splits;
}

extern "C" int cl_gemm(const cl_gemm_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (a == nullptr || a->a == nullptr || a->b == nullptr || a->out == nullptr)
        return set_error(CL_ERR_INVALID, "cl_gemm: null pointer");
    if (a->M <= 0 || a->N <= 0 || a->K <= 0) return set_error(CL_ERR_INVALID, "cl_gemm: non-positive dims");
    if (a->N % 4 != 0) return set_error(CL_ERR_INVALID, "cl_gemm: N must be a multiple of 4");
    if (a->K % 8 != 0 || a->ldb % 8 != 0) return set_error(CL_ERR_INVALID, "cl_gemm: K/ldb must be multiples of 8");
    const bool lora = a->lora_up != nullptr;
    if (lora && (a->ext == nullptr || (a->lora_rp != 4 && a->lora_rp != 8)))
        return set_error(CL_ERR_INVALID, "cl_gemm: LoRA epilogue needs ext and lora_rp in {4,8}");
    if ((a->t_add || a->t_out) && !lora) return set_error(CL_ERR_INVALID, "cl_gemm: t_add/t_out need the LoRA epilogue");

    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = a->M; p.N = a->N; p.K = a->K;
    p.a_mode = a->a_mode;
    // K-block: 64 (128B swizzle) unless this is a conv whose channel count only allows 32-element K runs
    const int BK = gemm_block_k(a);
    const int swz = (BK == 64) ? 128 : 64;
    p.num_k_blocks = (a->K + BK - 1) / BK;
    CUtensorMap tA, tB, tE;
    memset(&tE, 0, sizeof(tE));

    if (a->a_mode == 0) {
        if (a->lda % 8 != 0) return set_error(CL_ERR_INVALID, "cl_gemm: lda must be a multiple of 8");
        p.num_m_blocks = (a->M + BLOCK_M - 1) / BLOCK_M;
        uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
        uint64_t strides[1] = {(uint64_t)a->lda * 2};
        uint32_t box[2] = {(uint32_t)BK, BLOCK_M};
        CL_CHECK(get_tensor_map(&tA, a->a, 2, dims, strides, box, swz));
    } else if (a->a_mode == 1 || a->a_mode == 2) {
        if (a->C % 32 != 0) return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: conv C must be a multiple of 32");
        if (a->K != 9 * a->C) return set_error(CL_ERR_INVALID, "cl_gemm: conv K must be 9*C");
        const int s = (a->a_mode == 2) ? 2 : 1;
        if (a->H % s || a->W % s) return set_error(CL_ERR_INVALID, "cl_gemm: stride-2 conv needs even H, W");
        p.n_img = a->n_img; p.Ho = a->H / s; p.Wo = a->W / s; p.C = a->C; p.cblocks = a->C / BK;
        p.pad_lo = a->pad_lo;
        if (a->M != p.n_img * p.Ho * p.Wo) return set_error(CL_ERR_INVALID, "cl_gemm: conv M != n*Ho*Wo");
        int bw = 128;
        while (bw > 1 && (p.Wo % bw) != 0) bw >>= 1;
        int bh = 128 / bw;
        while (bh > 1 && (p.Ho % bh) != 0) bh >>= 1;
        int bn = 128 / (bw * bh);
        if (bw > 256 || bh > 256 || bn > 256) return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: conv tile");
        p.bw = bw; p.bh = bh; p.bn = bn;
        p.tiles_w = p.Wo / bw; p.tiles_h = p.Ho / bh; p.tiles_n = (p.n_img + bn - 1) / bn;
        p.num_m_blocks = p.tiles_w * p.tiles_h * p.tiles_n;
        const uint64_t C = a->C, W = a->W, H = a->H, NI = a->n_img;
        if (a->a_mode == 1) {
            uint64_t dims[4] = {C, W, H, NI};
            uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
            uint32_t box[4] = {(uint32_t)BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
            CL_CHECK(get_tensor_map(&tA, a->a, 4, dims, strides, box, swz));
        } else {
            uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, NI};
            uint64_t strides[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
            uint32_t box[5] = {(uint32_t)BK, (uint32_t)bw, 1, (uint32_t)bh, (uint32_t)bn};
            CL_CHECK(get_tensor_map(&tA, a->a, 5, dims, strides, box, swz));
        }
    } else {
        return set_error(CL_ERR_INVALID, "cl_gemm: bad a_mode");
    }

    // block_n / split choice: LoRA needs UMMA_N = BN + 16 <= 256.
    if (a->split_k > 1 && lora) return set_error(CL_ERR_INVALID, "cl_gemm: split_k cannot be combined with the LoRA epilogue");
    if (a->split_k > 1 && a->split_ws == nullptr) return set_error(CL_ERR_INVALID, "cl_gemm: split_k needs split_ws");
    const TilePlan plan = plan_tiles(a, BK, p.num_m_blocks, a->split_k > 1);
    const int bn_sel = plan.bn;
    const int cg = (bn_sel >= 64) ? gemm_cta_group(a, BK, p.num_m_blocks) : 1;
    {
        static const int dbg = [] { const char* e = getenv("CLB_GEMM_DEBUG"); return (e && e[0] == '1') ? 1 : 0; }();
        if (dbg) fprintf(stderr, "cl_gemm M=%d N=%d K=%d mode=%d lora=%d -> bn=%d cg=%d splits=%d (asked bn=%d split_k=%d)\n", a->M, a->N, a->K,
                         a->a_mode, lora ? 1 : 0, bn_sel, cg, plan.splits, a->block_n, a->split_k);
    }
    p.num_n_blocks = (a->N + bn_sel - 1) / bn_sel;
    p.splits = 1;
    p.kb_per_split = p.num_k_blocks;
    if (plan.splits > 1) {
        const int want = plan.splits < a->split_k ? plan.splits : a->split_k;   // split_ws holds a->split_k partials
        p.kb_per_split = (p.num_k_blocks + want - 1) / want;
        if (BK == 32) p.kb_per_split = (p.kb_per_split + 2) / 3 * 3;
        p.splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;     // no empty split
        p.split_ws = reinterpret_cast<float*>(a->split_ws);
    }
    {
        uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
        uint64_t strides[1] = {(uint64_t)a->ldb * 2};
        // a CTA of a pair stages half of the B rows of each MMA (the 320-column tile issues two 160-column MMAs per k-step)
        uint32_t box[2] = {(uint32_t)BK, (uint32_t)((bn_sel > 256 ? bn_sel / 2 : bn_sel) / cg)};
        CL_CHECK(get_tensor_map(&tB, a->b, 2, dims, strides, box, swz));
    }
    if (lora) {
        if (a->ldb_ext % 8 != 0) return set_error(CL_ERR_INVALID, "cl_gemm: ldb_ext must be a multiple of 8");
        uint64_t dims[2] = {(uint64_t)a->K, 16};
        uint64_t strides[1] = {(uint64_t)a->ldb_ext * 2};
        uint32_t box[2] = {(uint32_t)BK, 16};
        CL_CHECK(get_tensor_map(&tE, a->ext, 2, dims, strides, box, swz));
    }

#ifdef CLB_TIMELINE
    { const char* e = getenv("CLB_TL_MMA_REPS"); p.dbg_reps = e ? atoi(e) : 0; }
    { static int ordinal = 0; p.dbg_id = (ordinal++) & 63; }
#endif
    p.bias = a->bias; p.row_bias = a->row_bias; p.rows_per_group = a->rows_per_group;
    p.ld_rb = a->ld_row_bias > 0 ? a->ld_row_bias : a->N;
    if (p.row_bias && (p.ld_rb % 4)) return set_error(CL_ERR_INVALID, "cl_gemm: ld_row_bias must be a multiple of 4");
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual); p.ldr = a->ldr;
    p.lora_up = a->lora_up; p.lora_rp = lora ? a->lora_rp : 4; p.lora_scale = a->lora_scale;
    p.t_add = a->t_add; p.t_out = a->t_out;
    p.out = a->out; p.ldd = a->ldd; p.out_fp32 = a->out_fp32;
    if (p.row_bias && p.rows_per_group <= 0) return set_error(CL_ERR_INVALID, "cl_gemm: rows_per_group");
    if ((p.ldd % 4) || (p.residual && (p.ldr % 4))) return set_error(CL_ERR_INVALID, "cl_gemm: ldd/ldr % 4");

    // bf16 outputs whose rows are 16-byte addressable leave through per-warp TMA stores (thread == row epilogue); fp32
    // outputs, split-K partial tiles and odd strides keep the per-lane store path.  CLB_GEMM_TMA_STORE=0 disables.
    CUtensorMap tD;
    memset(&tD, 0, sizeof(tD));
    static const int tma_store_enabled = [] { const char* e = getenv("CLB_GEMM_TMA_STORE"); return (e && e[0] == '0') ? 0 : 1; }();
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(a->out) & 15) == 0) && (p.ldd % 8 == 0) && (a->N % 8 == 0) &&
                           (!p.residual || (((reinterpret_cast<uintptr_t>(a->residual) & 15) == 0) && (p.ldr % 8 == 0)));
    // CLB_GEMM_EPI: 1 = TMA stores (default), 2 = per-thread 16-byte global stores; 3 / 4 = timing experiments that produce
    // WRONG results (no stores / no epilogue) and exist only to attribute the kernel's time
    static const int epi_mode = [] { const char* e = getenv("CLB_GEMM_EPI"); return (e && e[0] >= '1' && e[0] <= '4') ? e[0] - '0' : 1; }();
    if (tma_store_enabled && !p.out_fp32 && p.splits == 1 && aligned16 && epi_mode >= 2) {
        p.tma_store = epi_mode;
    } else if (tma_store_enabled && !p.out_fp32 && p.splits == 1 && aligned16) {
        if (a->a_mode == 0) {
            uint64_t dims[2] = {(uint64_t)a->N, (uint64_t)a->M};
            uint64_t strides[1] = {(uint64_t)p.ldd * 2};
            uint32_t box[2] = {32, 32};
            CL_CHECK(get_tensor_map(&tD, a->out, 2, dims, strides, box, 64));
            p.tma_store = 1;
        } else if (p.ldd == a->N) {
            // rows quad*32 .. quad*32+31 of a (bn x bh x bw) output tile, w fastest, form the box (en x eh x ew)
            p.ew = p.bw < 32 ? p.bw : 32;
            p.eh = (32 / p.ew) < p.bh ? (32 / p.ew) : p.bh;
            p.en = 32 / (p.ew * p.eh);
            if (p.en <= p.bn && p.ew * p.eh * p.en == 32) {
                uint64_t dims[4] = {(uint64_t)a->N, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.n_img};
                uint64_t strides[3] = {(uint64_t)a->N * 2, (uint64_t)p.Wo * a->N * 2, (uint64_t)p.Ho * p.Wo * a->N * 2};
                uint32_t box[4] = {32, (uint32_t)p.ew, (uint32_t)p.eh, (uint32_t)p.en};
                CL_CHECK(get_tensor_map(&tD, a->out, 4, dims, strides, box, 64));
                p.tma_store = 1;
            }
        }
    }

    p.epi_bf16 = (!p.out_fp32 && p.splits == 1 && aligned16 && p.bias == nullptr && p.row_bias == nullptr && p.residual == nullptr) ? 1 : 0;
    if (BK == 32) {
        if (lora) return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: LoRA epilogue is not instantiated for 32-channel convs");
        // halo-strip operand (stride-1 conv, C = 32, tiles = 128-pixel runs of one image row): box = 130 pixels of one row
        CUtensorMap tS;
        const CUtensorMap* tSp = nullptr;
        static const int strip_mode = [] { const char* e = getenv("CLB_GEMM_STRIP"); return e ? atoi(e) : 1; }();   // 0 off, 1..3: base-offset rule + 1
        if (strip_mode > 0 && a->a_mode == 1 && a->C == 32 && p.bw == 128 && p.bh == 1 && p.bn == 1 && p.splits == 1) {
            const uint64_t C = a->C, W = a->W, H = a->H, NI = a->n_img;
            uint64_t dims[4] = {C, W, H, NI};
            uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
            uint32_t box[4] = {32, STRIP_ROWS, 1, 1};
            CL_CHECK(get_tensor_map(&tS, a->a, 4, dims, strides, box, 64));
            tSp = &tS;
            p.strip_bo = strip_mode - 1;
        }
        switch (bn_sel) {
            case 32: return launch_gemm<32, 0, 32>(tA, tB, tE, tD, p, stream, tSp);
            case 64: return launch_gemm<64, 0, 32>(tA, tB, tE, tD, p, stream, tSp);
            case 128: return launch_gemm<128, 0, 32>(tA, tB, tE, tD, p, stream, tSp);
            default: return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: block_n for 32-channel convs must be 32/64/128");
        }
    }
    if (lora) {
        switch (bn_sel) {
            case 64: return launch_gemm<64, 16>(tA, tB, tE, tD, p, stream);
            case 128: return launch_gemm<128, 16>(tA, tB, tE, tD, p, stream);
            case 160: return launch_gemm<160, 16>(tA, tB, tE, tD, p, stream);
            default: return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: block_n for LoRA must be 64/128/160");
        }
    }
    if (cg == 2) {
        switch (bn_sel) {
            case 64: return launch_gemm<64, 0, 64, 2>(tA, tB, tE, tD, p, stream);
            case 128: return launch_gemm<128, 0, 64, 2>(tA, tB, tE, tD, p, stream);
            case 160: return launch_gemm<160, 0, 64, 2>(tA, tB, tE, tD, p, stream);
            case 256: return launch_gemm<256, 0, 64, 2>(tA, tB, tE, tD, p, stream);
            case 320: return launch_gemm<320, 0, 64, 2>(tA, tB, tE, tD, p, stream);
            default: return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: block_n must be 64/128/160/256/320");
        }
    }
    switch (bn_sel) {
#ifdef CLB_TIMELINE
        case 96: return launch_gemm<96, 0>(tA, tB, tE, tD, p, stream);
        case 192: return launch_gemm<192, 0>(tA, tB, tE, tD, p, stream);
        case 224: return launch_gemm<224, 0>(tA, tB, tE, tD, p, stream);
#endif
        case 64: return launch_gemm<64, 0>(tA, tB, tE, tD, p, stream);
        case 128: return launch_gemm<128, 0>(tA, tB, tE, tD, p, stream);
        case 160: return launch_gemm<160, 0>(tA, tB, tE, tD, p, stream);
        case 256: return launch_gemm<256, 0>(tA, tB, tE, tD, p, stream);
        default: return set_error(CL_ERR_UNSUPPORTED, "cl_gemm: block_n must be 64/128/160/256");
    }
}
