// fp64_microbench.cu -- what the B200 FP64 pipes sustain: DMMA (mma.sync.m8n8k4.f64) alone, DFMA alone, both
// interleaved, and DMMA fed from shared memory.  Prints TFLOP/s, the SM clock seen by clock64 and flop/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_microbench tools/fp64_microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NM DMMA accumulators + NF DFMA accumulators per thread, interleaved per iteration.
template <int NM, int NF>
__global__ void mix_kernel(double* out, long long* cyc, int iters) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double c[NM > 0 ? NM : 1][2];
    double f[NF > 0 ? NF : 1];
#pragma unroll
    for (int j = 0; j < NM; j++) c[j][0] = c[j][1] = 0.0;
#pragma unroll
    for (int j = 0; j < NF; j++) f[j] = j;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < (NM > NF ? NM : NF); j++) {
            if (j < NM) dmma(c[j][0], c[j][1], a, b);
            if (j < NF) asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(f[j]) : "d"(a), "d"(b));
        }
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NM; j++) s += c[j][0] + c[j][1];
#pragma unroll
    for (int j = 0; j < NF; j++) s += f[j];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

// DMMA with fragments loaded from shared memory each k-step (MI x NI register tile per warp: the GEMM inner loop).
template <int MI, int NI>
__global__ void lds_kernel(double* out, long long* cyc, int iters) {
    __shared__ double sa[64 * 20], sb[64 * 20];
    for (int i = threadIdx.x; i < 64 * 20; i += blockDim.x) { sa[i] = 1e-3 * i; sb[i] = 1.0 - 1e-3 * i; }
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 16; kk += 4) {
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; i++) af[i] = sa[((i * 8 + g) & 63) * 20 + kk + q];
#pragma unroll
            for (int j = 0; j < NI; j++) bf[j] = sb[((j * 8 + g) & 63) * 20 + kk + q];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) s += acc[i][j][0] + acc[i][j][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
}

static double* g_out;
static long long* g_cyc;
static int g_sms;

template <typename K>
void run(const char* name, K kern, int threads, int ctas_per_sm, double flop_per_thread_iter, int iters) {
    int grid = g_sms * ctas_per_sm;
    kern<<<grid, threads>>>(g_out, g_cyc, 200);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    long long cyc = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        kern<<<grid, threads>>>(g_out, g_cyc, iters);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; CK(cudaMemcpy(&cyc, g_cyc, 8, cudaMemcpyDeviceToHost)); }
    }
    double flop = (double)grid * threads * iters * flop_per_thread_iter;
    double tf = flop / (best * 1e-3) / 1e12;
    double mhz = cyc / (best * 1e-3) / 1e6;
    printf("%-44s thr=%4d cta/sm=%d  %8.2f ms  %6.2f TF  clk~%5.0f MHz  %6.1f flop/clk/SM\n", name, threads, ctas_per_sm, best, tf,
           mhz, flop / g_sms / (double)cyc);
    fflush(stdout);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    g_sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, g_sms);
    CK(cudaMalloc(&g_out, (size_t)g_sms * 8 * 1024 * 8));
    CK(cudaMalloc(&g_cyc, 8));
    const int IT = 40000;
    const double DM = 512.0 / 32.0;  // flop per thread per DMMA
    run("dmma x4", mix_kernel<4, 0>, 256, 2, 4 * DM, IT);
    run("dmma x8", mix_kernel<8, 0>, 256, 2, 8 * DM, IT);
    run("dmma x16", mix_kernel<16, 0>, 256, 2, 16 * DM, IT / 2);
    run("dmma x16", mix_kernel<16, 0>, 128, 1, 16 * DM, IT / 2);
    run("dmma x16", mix_kernel<16, 0>, 256, 1, 16 * DM, IT / 2);
    run("dmma x16", mix_kernel<16, 0>, 512, 1, 16 * DM, IT / 2);
    run("dmma x16", mix_kernel<16, 0>, 1024, 1, 16 * DM, IT / 2);
    run("dmma x32", mix_kernel<32, 0>, 256, 1, 32 * DM, IT / 4);
    run("dmma x32", mix_kernel<32, 0>, 256, 2, 32 * DM, IT / 4);
    run("dfma x8", mix_kernel<0, 8>, 256, 2, 8 * 2.0, IT * 2);
    run("dfma x16", mix_kernel<0, 16>, 256, 2, 16 * 2.0, IT);
    run("dfma x16", mix_kernel<0, 16>, 1024, 1, 16 * 2.0, IT);
    run("dfma x32", mix_kernel<0, 32>, 256, 2, 32 * 2.0, IT / 2);
    run("mix dmma x16 + dfma x2", mix_kernel<16, 2>, 256, 2, 16 * DM + 2 * 2.0, IT / 2);
    run("mix dmma x16 + dfma x4", mix_kernel<16, 4>, 256, 2, 16 * DM + 4 * 2.0, IT / 2);
    run("mix dmma x16 + dfma x8", mix_kernel<16, 8>, 256, 2, 16 * DM + 8 * 2.0, IT / 2);
    run("mix dmma x16 + dfma x16", mix_kernel<16, 16>, 256, 2, 16 * DM + 16 * 2.0, IT / 2);
    run("mix dmma x8 + dfma x16", mix_kernel<8, 16>, 256, 2, 8 * DM + 16 * 2.0, IT / 2);
    run("lds dmma 4x4 tile (per 16-k: 16 lds, 64 dmma)", lds_kernel<4, 4>, 256, 1, 4 * 16 * DM, IT / 8);
    run("lds dmma 8x4 tile", lds_kernel<8, 4>, 256, 1, 4 * 32 * DM, IT / 16);
    run("lds dmma 8x4 tile", lds_kernel<8, 4>, 256, 2, 4 * 32 * DM, IT / 16);
    run("lds dmma 8x8 tile", lds_kernel<8, 8>, 128, 1, 4 * 64 * DM, IT / 32);
    run("lds dmma 8x8 tile", lds_kernel<8, 8>, 256, 1, 4 * 64 * DM, IT / 32);
    run("lds dmma 4x8 tile", lds_kernel<4, 8>, 256, 1, 4 * 32 * DM, IT / 16);
    run("lds dmma 4x8 tile", lds_kernel<4, 8>, 512, 1, 4 * 32 * DM, IT / 16);
    return 0;
}
