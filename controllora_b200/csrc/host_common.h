// Host-side helpers shared by the C-ABI translation units: error reporting, launch counting, TMA descriptor cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/controllora_b200.h"

namespace clb {

int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();

// Encode (or fetch from the cache) a tiled bf16 tensor map.  dims/box are innermost-first, strides are the byte
// strides of dims 1..rank-1.  Returns CL_OK or a negative status.
int get_tensor_map(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, int swizzle_bytes /* 0, 64 or 128 */);

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
