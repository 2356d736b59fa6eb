/* controllora_b200 — C ABI of the B200-native ControlLoRA UNet hot path.
 *
 * The reference (HighCWu/ControlLoRA) is pure Python: its "FFI" for this path is the set of torch/ATen calls
 * issued by diffusers' UNet2DConditionModel and by models.py's attention processors.  Each entry point below
 * replaces one family of those calls; the reference call site it stands in for is cited (paths relative to
 * /root/reference, diffusers = the un-vendored diffusers 0.13/0.14 dependency).
 *
 * Conventions
 *   - every function is stream-ordered, takes raw device pointers + sizes, never allocates, never syncs,
 *     never throws; it returns 0 on success or a negative cl_status, with cl_last_error() giving the text.
 *   - activations are bf16, channels-last: an (N,H,W,C) image tensor and a (N*H*W, C) token matrix are the
 *     same memory.  Frozen weights are bf16, row-major [out, in] (conv: [out, ky, kx, in]).
 *   - trainable (LoRA / hint-encoder) parameters and all gradients of trainable parameters are fp32.
 *   - `stream` is a cudaStream_t passed as void* so that the header has no CUDA dependency.
 */
#ifndef CONTROLLORA_B200_H
#define CONTROLLORA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CL_OK = 0,
    CL_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    CL_ERR_CUDA = -2,      /* a CUDA runtime/driver call failed */
    CL_ERR_UNSUPPORTED = -3
} cl_status;

const char* cl_last_error(void);
int cl_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t cl_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * K1/K4: tcgen05 GEMM  D[M,N] = epilogue( A[M,K] * B[N,K]^T )   (bf16 x bf16 -> fp32 in TMEM)
 *
 * Replaces: attn.to_q/to_k/to_v/to_out[0] + LoRALinearLayer side path  (models.py:124-147, 231-282, 373-423),
 *           diffusers FeedForward linears, Transformer2DModel proj_in/proj_out (1x1 conv == GEMM on NHWC),
 *           ResnetBlock2D conv1/conv2/conv_shortcut, Downsample2D/Upsample2D convs (3x3 as implicit GEMM),
 *           and their dX backward (same kernel on the transposed / flipped weight copies).
 *
 * a_mode 0: A is a row-major [M, K] matrix (lda elements between rows).
 * a_mode 1: A is an NHWC image (n_img, H, W, C); the GEMM row m = (n*H + h)*W + w reads the 3x3 window around
 *           (h, w) with zero padding 1; K = 9*C, k = (ky*3 + kx)*C + c.  B is [N, 9*C].   C % 64 == 0.
 * a_mode 2: as 1 but stride 2: output (n, ho, wo) of size (H/2, W/2) reads input (2ho+ky-pad_lo, 2wo+kx-pad_lo);
 *           pad_lo = 1 is diffusers' UNet Downsample2D(padding=1), pad_lo = 0 is the hint encoder's
 *           F.pad(0,1,0,1) + conv(padding=0) (models.py:591-598).
 *
 * LoRA epilogue (lora_up != NULL): `ext` is a bf16 [16, K] matrix appended below B's rows, so the tensor core also
 * produces e = A*ext^T (16 extra accumulator columns); t[j] = e[j] + e[j+8] (+ t_add[m, j]) for j < lora_rp and
 * D[m, n] += lora_scale * sum_j t[j] * lora_up[n*lora_rp + j].   (ext rows j / j+8 hold the bf16 hi / lo split of
 * LoRA-down row j, which keeps the rank-r path at ~fp32 accuracy like the reference's fp32 LoRALinearLayer.)
 * If t_out != NULL, t (fp32 [M, lora_rp]) is stored for the backward pass.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t M, N, K;
    int32_t a_mode;
    const void* a;      /* bf16 */
    int64_t lda;        /* a_mode 0 only */
    int32_t n_img, H, W, C; /* a_mode 1/2: INPUT image geometry */
    int32_t pad_lo;     /* a_mode 2 */
    const void* b;      /* bf16 [N, K], row pitch ldb */
    int64_t ldb;
    const void* ext;    /* bf16 [16, K] row pitch ldb_ext, or NULL */
    int64_t ldb_ext;
    /* epilogue */
    const float* bias;       /* [N] or NULL */
    const float* row_bias;   /* [n_groups, N] or NULL; group of row m = m / rows_per_group */
    int32_t rows_per_group;
    const void* residual;    /* bf16 [M, N] pitch ldr, or NULL; may alias out */
    int64_t ldr;
    const float* lora_up;    /* fp32 [N, lora_rp] or NULL */
    int32_t lora_rp;         /* 4 or 8 */
    float lora_scale;
    const float* t_add;      /* fp32 [M, lora_rp] or NULL */
    float* t_out;            /* fp32 [M, lora_rp] or NULL */
    void* out;               /* bf16 (out_fp32 = 0) or fp32 [M, N], pitch ldd */
    int64_t ldd;
    int32_t out_fp32;
    int32_t block_n;         /* 0 = auto */
} cl_gemm_args;

int cl_gemm(const cl_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
