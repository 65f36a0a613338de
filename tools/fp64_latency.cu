// Single-warp FP64 latency / issue probe on B200: dependent vs independent DFMA, DMMA, MUFU.RSQ64H (via rsqrt), SHFL.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
    const int lane = threadIdx.x & 31;
    double x = seed + lane * 1e-3, y = 1.000001, z = 0.5;
    long long t0, t1;
    // 1. dependent DFMA chain, 256 long
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = fma(x, y, z);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // 2. 8 independent chains x 32
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = x + i;
    t0 = clock64();
#pragma unroll
    for (int r = 0; r < 32; r++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], y, z);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    for (int i = 0; i < 8; i++) x += a[i];
    // 3. dependent DMUL chain
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) x = x * y;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // 4. dependent rsqrt chain (64)
    double w = fabs(x) + 2.0;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) w = rsqrt(w) + 1.5;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // 5. dependent SHFL chain of a double (64)
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) w = __shfl_sync(0xffffffffu, w, (i * 7 + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // 6. dependent DMMA chain (64)
    double c0 = w, c1 = x;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(y), "d"(z));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // 7. 8 independent DMMA chains x 16
    double m0[8], m1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { m0[i] = c0 + i; m1[i] = c1 - i; }
    t0 = clock64();
#pragma unroll
    for (int r = 0; r < 16; r++)
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(m0[i]), "+d"(m1[i]) : "d"(y), "d"(z));
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    // 8. dependent FFMA chain (256) for reference
    float f = (float)x, fy = 1.0001f, fz = 0.25f;
    t0 = clock64();
#pragma unroll
    for (int i = 0; i < 256; i++) f = fmaf(f, fy, fz);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = t1 - t0;
    double s = x + w + c0 + c1 + f;
    for (int i = 0; i < 8; i++) s += m0[i] + m1[i];
    out[threadIdx.x] = s;
}
int main() {
    double* d; long long* c;
    cudaMalloc(&d, 1024 * 8); cudaMalloc(&c, 64 * 8);
    for (int threads : {32, 128, 512}) {
        for (int rep = 0; rep < 2; rep++) k<<<1, threads>>>(d, c, 1.25);
        cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads=%3d: DFMA dep %.1f cyc/op | DFMA 8 indep %.1f cyc/op | DMUL dep %.1f | rsqrt(double)+add dep %.1f | SHFL.64 dep %.1f | DMMA dep %.1f | DMMA 8 indep %.1f cyc/op | FFMA dep %.1f\n",
               threads, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 128.0, h[7] / 256.0);
    }
    return 0;
}
