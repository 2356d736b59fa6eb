// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting, TMA descriptor cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/controllora_b200.h"

namespace clb {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

// Encode (or fetch from the cache) a tiled bf16 tensor map.  dims/box are innermost-first, strides are the byte
// strides of dims 1..rank-1.  Returns CL_OK or a negative status.
int get_tensor_map(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes /* 0, 64 or 128 */);

// Programmatic dependent launch (default: GEMM family only; CLB_PDL=0 disables): see pdl_launch_dependents / pdl_wait in common.cuh.
// family: the CLB_FAMILY bit of the launching file (gemm 1, attention 2, norm 4, lora 8, elementwise 16, smallops 32, wgrad 64,
// noise / optim 128); CLB_PDL_MASK (default: 1 = gemm, see host_common.cu) selects which families' kernels may start early.
bool pdl_enabled(int family);
#ifndef CLB_FAMILY
#define CLB_FAMILY 128
#endif

// Launch `kernel` with the programmatic-stream-serialization attribute: inside a stream (or a captured CUDA graph) the
// grid may start while the previous kernel drains; every kernel of this library executes griddepcontrol.wait before its
// first global-memory access.  A launch error is left in cudaGetLastError() for the caller's check.
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled(CLB_FAMILY) ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace clb

#define CL_CHECK(expr)                   \
    do {                                 \
        int _st = (expr);                \
        if (_st != CL_OK) return _st;    \
    } while (0)

#define CL_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return clb::set_error(CL_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
    } while (0)
