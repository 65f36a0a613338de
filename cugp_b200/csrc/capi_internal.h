// capi_internal.h -- helpers shared by the translation units that implement include/cugp.h.
#pragma once
#include "../../include/cugp.h"
#include "common.cuh"

#include <new>

namespace cugp {
struct GpBatch;
void set_last_error(const char* fmt, ...);
}  // namespace cugp
void track(cugp::GpBatch* g);    // launch counting (cugp_launch_count) covers every live batch
void untrack(cugp::GpBatch* g);
int require_device();      // CUGP_OK or CUGP_ERR_NODEVICE with the last-error message set

#define CUGP_TRY try {
#define CUGP_CATCH                                                                                          \
    }                                                                                                       \
    catch (const CudaError& e) {                                                                            \
        set_last_error("CUDA error %d (%s) at %s:%d", (int)e.code, cudaGetErrorString(e.code), e.file, e.line); \
        cudaGetLastError();                                                                                 \
        return e.code == cudaErrorMemoryAllocation ? CUGP_ERR_NOMEM : CUGP_ERR_CUDA;                        \
    }                                                                                                       \
    catch (const std::bad_alloc&) {                                                                         \
        set_last_error("host allocation failed");                                                           \
        return CUGP_ERR_NOMEM;                                                                              \
    }                                                                                                       \
    catch (...) {                                                                                           \
        set_last_error("unexpected exception");                                                             \
        return CUGP_ERR_CUDA;                                                                               \
    }

