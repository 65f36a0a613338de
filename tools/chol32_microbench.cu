// Microbenchmark of the 32x32 warp-level Cholesky(+inverse) column loop: which part of a column costs what.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/chol32_microbench tools/chol32_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int SBW = 32, LDS_ = 132;

// VARIANT 0: full (chol + inverse, masked zeroing)   1: chol only (no inverse, no masks)
//         2: chain only (no column update)           3: update only (rd = const, no rsqrt / shfl chain)
//         4: full, but broadcast through SHFL instead of shared memory
__device__ long long g_loop_cycles[4];
__device__ long long g_ls_cycles[2];
__device__ __forceinline__ long long clk() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
    return t;
}
// branch-free 1/sqrt(x): MUFU.RSQ64H seed + the one cubic Newton step CUDA's rsqrt() uses, without its special-case branch
__device__ __forceinline__ double rsqrt_nb(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = y * y;
    const double e = fma(-t, x, 1.0);
    const double p2 = fma(e, 0.375, 0.5);
    const double s = y * e;
    return fma(p2, s, y);
}
template <int VARIANT>
__device__ __forceinline__ void chol32(double* S, int c0, double* cb, int lane) {
    double u[SBW];
    double* row = S + (c0 + lane) * LDS_ + c0;
    const long long ta = clk();
#pragma unroll
    for (int c = 0; c < SBW; c++) u[c] = (c < lane) ? row[c] : 0.0;
    double adiag = row[lane];
    double ldiag = 0.0, tdiag = 0.0;
    double ajj = __shfl_sync(0xffffffffu, adiag, 0);
    const long long tl0 = clk();
    if (lane == 0) g_ls_cycles[0] = tl0 - ta;
#pragma unroll
    for (int j = 0; j < SBW; j++) {
        const double rd = VARIANT == 3 ? 0.9 : (VARIANT == 7 ? ajj * 0.01 : (VARIANT == 8 || VARIANT == 9 ? rsqrt_nb(ajj) : rsqrt(ajj)));
        const double d = ajj * rd;
        const double val = u[j] * rd;
        const double l = (lane > j) ? val : 0.0;
        const double m = (VARIANT == 1 || VARIANT == 9) ? l : ((lane == j) ? rd : val);
        if (lane == j) { ldiag = d; tdiag = rd; }
        u[j] = val;
        if (j + 1 < SBW) {
            adiag = fma(-l, l, adiag);
            if (VARIANT == 6) ajj = adiag + 100.0;
            else if (VARIANT != 3) ajj = __shfl_sync(0xffffffffu, adiag, j + 1);
            if (VARIANT == 2 || VARIANT == 6 || VARIANT == 7) continue;
            if (VARIANT == 4) {
                const long long keep = (lane == j) ? 0ll : ~0ll;
#pragma unroll
                for (int c = j + 1; c < SBW; c++) {
                    const double bc = __shfl_sync(0xffffffffu, l, c);
                    u[c] = fma(-bc, m, __longlong_as_double(__double_as_longlong(u[c]) & keep));
                }
                continue;
            }
            double* buf = cb + (j & 1) * SBW;
            buf[lane] = l;
            __syncwarp();
            if (VARIANT == 1 || VARIANT == 9) {
#pragma unroll
                for (int c = j + 1; c < SBW; c++) u[c] = fma(-buf[c], m, u[c]);
            } else {
                const long long keep = (lane == j) ? 0ll : ~0ll;
#pragma unroll
                for (int c = j + 1; c < SBW; c++)
                    u[c] = fma(-buf[c], m, __longlong_as_double(__double_as_longlong(u[c]) & keep));
            }
        }
    }
    const long long tl1 = clk();
    if (lane == 0) g_loop_cycles[c0 / 32] = tl1 - tl0;
#pragma unroll
    for (int c = 0; c < SBW; c++)
        if (c < lane) row[c] = u[c];
    row[lane] = ldiag;
    row[lane + 1] = tdiag;
#pragma unroll
    for (int c = 1; c < SBW; c++)
        if (c > lane) row[c + 1] = u[c];
    const long long tb = clk();
    if (lane == 0) g_ls_cycles[1] = tb - tl1;
}

// VARIANT 5: round 1's loop (row in registers, pivot through the broadcast, reciprocal diagonal kept)
__device__ __forceinline__ void chol32_r1(double* S, int c0, double* colbuf, double* rdiag, int lane) {
    double a[SBW];
#pragma unroll
    for (int c = 0; c < SBW; c++) a[c] = (c <= lane) ? S[(c0 + lane) * LDS_ + c0 + c] : 0.0;
#pragma unroll
    for (int j = 0; j < SBW; j++) {
        const double ajj = __shfl_sync(0xffffffffu, a[j], j);
        const double rd = rsqrt(ajj);
        const double d = ajj * rd;
        const double l = (lane == j) ? d : a[j] * rd;
        a[j] = l;
        if (lane == j) rdiag[c0 + j] = rd;
        if (j + 1 < SBW) {
            if (lane == j + 1) a[j + 1] -= l * l;
            double* cb = colbuf + (j & 1) * SBW;
            cb[lane] = l;
            __syncwarp();
#pragma unroll
            for (int c = j + 1; c < SBW; c++) {
                const double lc = cb[c];
                if (c == j + 1) {
                    if (lane != j + 1) a[c] -= l * lc;
                } else {
                    a[c] -= l * lc;
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < SBW; c++)
        if (c <= lane) S[(c0 + lane) * LDS_ + c0 + c] = a[c];
}

template <int VARIANT>
__global__ void __launch_bounds__(512, 1) bench(const double* A, double* out, long long* cyc, int cb_off) {
    extern __shared__ double sm[];
    double* S = sm;
    double* cb = sm + cb_off;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < 128 * 128; e += 512) S[(e >> 7) * LDS_ + (e & 127)] = A[e];
    __syncthreads();
    for (int rep = 0; rep < 3; rep++) {
        for (int c = 0; c < 4; c++) {
            if (warp == 0) {
                long long t0 = clock64();
                if (VARIANT == 5) chol32_r1(S, 32 * c, cb, cb + 64, lane);
                else chol32<VARIANT>(S, 32 * c, cb, lane);
                long long t1 = clock64();
                if (lane == 0) cyc[rep * 4 + c] = t1 - t0;
            }
            __syncthreads();
        }
        if (rep < 2) {  // restore the input
            for (int e = tid; e < 128 * 128; e += 512) S[(e >> 7) * LDS_ + (e & 127)] = A[e];
            __syncthreads();
        }
    }
    for (int e = tid; e < 128 * 128; e += 512) out[e] = S[(e >> 7) * LDS_ + (e & 127)];
}

static size_t g_smem = (128 * LDS_ + 64 + 128) * sizeof(double);
static int g_cb_off = 128 * LDS_;
template <int V>
void run(const char* name, const double* dA, double* dOut, long long* dC) {
    size_t smem = g_smem;
    cudaFuncSetAttribute(bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<V><<<1, 512, smem>>>(dA, dOut, dC, g_cb_off);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[12];
    cudaMemcpy(h, dC, sizeof(h), cudaMemcpyDeviceToHost);
    static double hout[128 * 128];
    cudaMemcpy(hout, dOut, sizeof(hout), cudaMemcpyDeviceToHost);
    double cs = 0; for (int i = 0; i < 32; i++) for (int j = 0; j <= 32; j++) cs += hout[i * 128 + j] * (1 + 0.01 * i + 0.001 * j);
    printf("[cs %.15g] ", cs);
    printf("%-46s err=%d cycles per 32x32 (3 reps x 4 blocks):", name, (int)e);
    for (int i = 0; i < 12; i++) printf(" %lld", h[i]);
    long long lc[4] = {0, 0, 0, 0};
    cudaMemcpyFromSymbol(lc, g_loop_cycles, sizeof(lc));
    long long ls[2] = {0, 0};
    cudaMemcpyFromSymbol(ls, g_ls_cycles, sizeof(ls));
    printf(" load %lld store %lld", ls[0], ls[1]);
    printf("  -> %.0f per column (last rep); column loop alone %lld %lld %lld %lld\n", (h[8] + h[9] + h[10] + h[11]) / 128.0, lc[0], lc[1], lc[2], lc[3]);
}

int main() {
    static double A[128 * 128];
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < 128; j++) A[i * 128 + j] = (i == j) ? 130.0 : 1.0 / (1.0 + (i > j ? i - j : j - i));
    double *dA, *dOut;
    long long* dC;
    cudaMalloc(&dA, sizeof(A)); cudaMalloc(&dOut, sizeof(A)); cudaMalloc(&dC, 64 * 8);
    cudaMemcpy(dA, A, sizeof(A), cudaMemcpyHostToDevice);
    run<0>("full: chol + inverse, masked", dA, dOut, dC);
    g_smem = (size_t)(128 * LDS_ + 2 * 32 * LDS_ + 64 + 128) * sizeof(double) + 16;   // the step kernel's 204 KB
    g_cb_off = 128 * LDS_ + 2 * 32 * LDS_;
    run<0>("full, with the step kernel's smem size/layout", dA, dOut, dC);
    g_smem = (128 * LDS_ + 64 + 128) * sizeof(double);
    g_cb_off = 128 * LDS_;
    run<1>("chol only (no inverse, no masks)", dA, dOut, dC);
    run<2>("chain only (rsqrt/shfl, no column update)", dA, dOut, dC);
    run<3>("update only (no rsqrt/shfl chain)", dA, dOut, dC);
    run<4>("full, broadcast by SHFL instead of smem", dA, dOut, dC);
    run<8>("full, branch-free rsqrt", dA, dOut, dC);
    run<9>("chol only, branch-free rsqrt", dA, dOut, dC);
    run<5>("round 1 loop (chol only, pivot via broadcast)", dA, dOut, dC);
    run<6>("chain only, no SHFL (lane-local pivot)", dA, dOut, dC);
    run<7>("chain only, no rsqrt (rd = a*0.01), with SHFL", dA, dOut, dC);
    return 0;
}
