// common.cuh -- shared definitions for the cuGP B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace cugp {

// Status codes live in include/cugp.h (CUGP_OK ...); only capi.cu returns them.
void set_last_error(const char* fmt, ...);

struct CudaError {
    cudaError_t code;
    const char* file;
    int line;
};

#define CUGP_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) throw ::cugp::CudaError{_e, __FILE__, __LINE__};           \
    } while (0)

constexpr int kDiag = 128;  // diagonal-block size of the blocked factorisation (POTRF/TRTRI unit)

__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Leading dimension of every n x n device matrix: rows are 128-byte aligned so 16-byte async copies and
// vector stores never straddle a row start (n = 1500 or 10000 are not multiples of the tile sizes).
inline int64_t padded_ld(int64_t n) { return round_up(n, 16); }

}  // namespace cugp
