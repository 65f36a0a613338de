// probe.cu -- measured FP64 roofline denominators (MEASURED_PEAKS.json carries only HBM and bf16):
// register-resident DMMA and DFMA loops, the DMMA GEMM on resident operands, and a device copy.
#include "common.cuh"
#include "gemm_dmma.cuh"

#include <algorithm>
#include <vector>

namespace cugp {

namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, long long* cycles = nullptr) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double c[16][2];
#pragma unroll
    for (int j = 0; j < 16; j++) c[j][0] = c[j][1] = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1])
                         : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += c[j][0] + c[j][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (cycles && blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
    double c[16];
#pragma unroll
    for (int j = 0; j < 16; j++) c[j] = j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) c[j] = fma(c[j], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += c[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void fill_kernel(double* p, size_t count, unsigned seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        unsigned h = (unsigned)(i * 2654435761u) ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = (double)(h & 0xffff) / 65536.0 - 0.5;
    }
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = src[i];
}

template <typename F>
float time_ms(F&& f, cudaStream_t st) {
    cudaEvent_t e0, e1;
    CUGP_CUDA(cudaEventCreate(&e0));
    CUGP_CUDA(cudaEventCreate(&e1));
    CUGP_CUDA(cudaEventRecord(e0, st));
    f();
    CUGP_CUDA(cudaEventRecord(e1, st));
    CUGP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUGP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return ms;
}

}  // namespace

void probe_fp64_peak(float target_ms, double* dmma_tflops, double* dfma_tflops) {
    int dev = 0, sms = 148;
    CUGP_CUDA(cudaGetDevice(&dev));
    CUGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 4, threads = 256;
    double* out = nullptr;
    CUGP_CUDA(cudaMalloc(&out, (size_t)grid * threads * sizeof(double)));
    for (int which = 0; which < 2; which++) {
        auto run = [&](int iters) {
            if (which == 0) dmma_peak_kernel<<<grid, threads>>>(out, iters);
            else dfma_peak_kernel<<<grid, threads>>>(out, iters);
        };
        run(1000);  // warm-up
        CUGP_CUDA(cudaDeviceSynchronize());
        int iters = 20000;
        float ms = time_ms([&] { run(iters); }, 0);
        // scale to the requested duration so clocks settle under load, then measure
        iters = (int)std::min(2.0e9, std::max(20000.0, iters * (double)target_ms / std::max(ms, 1e-3f)));
        ms = time_ms([&] { run(iters); }, 0);
        double flops_per_thread_iter = which == 0 ? 16.0 * 512.0 / 32.0 : 16.0 * 2.0;
        double tf = (double)grid * threads * iters * flops_per_thread_iter / (ms * 1e-3) / 1e12;
        if (which == 0) *dmma_tflops = tf;
        else *dfma_tflops = tf;
    }
    CUGP_CUDA(cudaGetLastError());
    cudaFree(out);
}

// DMMA throughput over one launch of about `target_ms` milliseconds, one 256-thread CTA per SM (8 warps: the GEMM's
// occupancy), and the SM clock that launch actually ran at (clock64 ticks of one CTA / event time): a short launch
// gives the burst peak, a long one what the power limit sustains.
void probe_dmma(float target_ms, double* tflops, double* sm_mhz) {
    int dev = 0, sms = 148;
    CUGP_CUDA(cudaGetDevice(&dev));
    CUGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms, threads = 256;
    double* out = nullptr;
    long long* cyc = nullptr;
    CUGP_CUDA(cudaMalloc(&out, (size_t)grid * threads * sizeof(double)));
    CUGP_CUDA(cudaMalloc(&cyc, sizeof(long long)));
    dmma_peak_kernel<<<grid, threads>>>(out, 1000, cyc);
    CUGP_CUDA(cudaDeviceSynchronize());
    int iters = 20000;
    float ms = time_ms([&] { dmma_peak_kernel<<<grid, threads>>>(out, iters, cyc); }, 0);
    iters = (int)std::min(2.0e9, std::max(2000.0, iters * (double)target_ms / std::max(ms, 1e-3f)));
    ms = time_ms([&] { dmma_peak_kernel<<<grid, threads>>>(out, iters, cyc); }, 0);
    long long c = 0;
    CUGP_CUDA(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    *tflops = (double)grid * threads * iters * (16.0 * 512.0 / 32.0) / (ms * 1e-3) / 1e12;
    *sm_mhz = (double)c / (ms * 1e-3) / 1e6;
    cudaFree(out);
    cudaFree(cyc);
}

void probe_gemm(int M, int N, int K, int iters, double* tflops) {
    double *A = nullptr, *B = nullptr, *C = nullptr;
    const int64_t lda = padded_ld(K), ldc = padded_ld(N);
    CUGP_CUDA(cudaMalloc(&A, (size_t)M * lda * 8));
    CUGP_CUDA(cudaMalloc(&B, (size_t)N * lda * 8));
    CUGP_CUDA(cudaMalloc(&C, (size_t)M * ldc * 8));
    fill_kernel<<<1024, 256>>>(A, (size_t)M * lda, 1u);
    fill_kernel<<<1024, 256>>>(B, (size_t)N * lda, 2u);
    fill_kernel<<<1024, 256>>>(C, (size_t)M * ldc, 3u);
    GemmParams p{};
    p.A = A; p.lda = lda; p.B = B; p.ldb = lda; p.C = C; p.ldc = ldc;
    p.M = M; p.N = N; p.K = K; p.alpha = -1.0; p.beta = 1.0; p.batch = 1;
    launch_gemm(p, true, true, GEMM_BIG, 0);
    CUGP_CUDA(cudaDeviceSynchronize());
    float ms = time_ms([&] { for (int i = 0; i < iters; i++) launch_gemm(p, true, true, GEMM_BIG, 0); }, 0);
    *tflops = 2.0 * M * N * (double)K * iters / (ms * 1e-3) / 1e12;
    cudaFree(A); cudaFree(B); cudaFree(C);
}

void probe_copy(size_t bytes, int iters, double* gbs) {
    double2 *a = nullptr, *b = nullptr;
    bytes = bytes / 16 * 16;
    CUGP_CUDA(cudaMalloc(&a, bytes));
    CUGP_CUDA(cudaMalloc(&b, bytes));
    CUGP_CUDA(cudaMemset(a, 1, bytes));
    int dev = 0, sms = 148;
    CUGP_CUDA(cudaGetDevice(&dev));
    CUGP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    copy_kernel<<<sms * 8, 256>>>(a, b, bytes / 16);
    CUGP_CUDA(cudaDeviceSynchronize());
    float ms = time_ms([&] { for (int i = 0; i < iters; i++) copy_kernel<<<sms * 8, 256>>>(a, b, bytes / 16); }, 0);
    *gbs = 2.0 * (double)bytes * iters / (ms * 1e-3) / 1e9;
    cudaFree(a); cudaFree(b);
}

}  // namespace cugp
