// kernels.cu -- hand-written sm_100a kernels of the GP hot path other than the DMMA GEMM:
//   K1  covariance build (TMA bulk staging of X tiles, strict reference arithmetic order)
//   K5a cross covariance fused with the predictive-mean partials
//   K4  fused gradient trace (K and D rebuilt on the fly, Kinv streamed once)
//   K2a 128x128 diagonal-block Cholesky + triangular inverse (one CTA, shared memory)
//   K3  backward sweep alpha = L^-T z: thread-block-cluster panel chain (DSMEM) + single-owner panel updates,
//       alpha = T^T z streaming pass, log-likelihood finalisation
// Every reduction is a fixed-shape tree over per-CTA partials: results are run-to-run deterministic.
#include "kernels.cuh"

#include <cooperative_groups.h>

namespace cugp {

namespace {

// ------------------------------------------------------------------------------------------------
// TMA (1-D bulk async copy) + mbarrier primitives.  SASS: UBLKCP / SYNCS.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
                 "l"(gsrc), "r"(bytes), "r"(a)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}

// (ti, tj) with ti >= tj from a linear index over the lower-triangular tile set.
__device__ __forceinline__ void lower_tile(int x, int& ti, int& tj) {
    ti = (int)((sqrt(8.0 * (double)x + 1.0) - 1.0) * 0.5);
    while ((int64_t)(ti + 1) * (ti + 2) / 2 <= x) ti++;
    while ((int64_t)ti * (ti + 1) / 2 > x) ti--;
    tj = x - (int)((int64_t)ti * (ti + 1) / 2);
}

// ------------------------------------------------------------------------------------------------
// Covariance tiles.  64x64 outputs per CTA, 256 threads, each thread a 4x4 micro-tile:
//   rows  i0 + ty*4 + a            (a < 4)
//   cols  j0 + tx*2 + 32*b + e     (b < 2, e < 2)  -> 16-byte vector stores, 256 B contiguous per half-warp
// ------------------------------------------------------------------------------------------------
constexpr int CT = kCovTile;
constexpr int COV_THREADS = 256;
enum CovMode { COV_LOWER = 0, COV_FULL = 1, COV_CROSS = 2, COV_TRACE = 3 };

struct CovArgs {
    const double* Xi; int64_t sXi; int ni;
    const double* Xj; int64_t sXj; int nj;
    int dp;
    Hyper h;
    double* out; int64_t ld, sOut;
    const double* alpha; int64_t sAlpha;
    double* part; int64_t sPart;
    const double* Kinv; int64_t sKinv;
    int tiles_j;
};

// Stage the two X tiles: TMA bulk copy of the contiguous row blocks, then an in-smem transpose to
// [k][row] so the micro-tile operand loads are 16-byte, conflict-free and broadcast friendly.
__device__ __forceinline__ void stage_x_tiles(double* smem, const double* Xi, int i0, int ni, const double* Xj, int j0,
                                              int nj, int dp, unsigned long long* bar, double*& xt_i, double*& xt_j) {
    const int tid = threadIdx.x;
    double* raw_i = smem;
    double* raw_j = smem + CT * dp;
    xt_i = smem + 2 * CT * dp;
    xt_j = smem + 3 * CT * dp;
    const int rows_i = min(CT, ni - i0), rows_j = min(CT, nj - j0);
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        unsigned bi = (unsigned)(rows_i * dp * 8), bj = (unsigned)(rows_j * dp * 8);
        mbar_expect_tx(bar, bi + bj);
        bulk_g2s(raw_i, Xi + (int64_t)i0 * dp, bi, bar);
        bulk_g2s(raw_j, Xj + (int64_t)j0 * dp, bj, bar);
    }
    mbar_wait(bar, 0);
    for (int e = tid; e < CT * dp; e += COV_THREADS) {
        int r = e % CT, k = e / CT;
        xt_i[k * CT + r] = r < rows_i ? raw_i[r * dp + k] : 0.0;
        xt_j[k * CT + r] = r < rows_j ? raw_j[r * dp + k] : 0.0;
    }
    __syncthreads();
}

// Squared distances of the 4x4 micro-tile, accumulated in dimension order.  FAST = false: separately rounded
// subtract / multiply / add -- the reference's subtract_vec + dotproduct_vec (matrixops.cpp:216-229) compiled without FMA
// contraction -- so d2 is bit-identical to the CPU path.  FAST = true: the multiply-add is fused (20 instead of 30 FP64
// instructions per pair at d = 10).
template <bool FAST>
__device__ __forceinline__ void micro_d2(const double* xt_i, const double* xt_j, int dp, int ty, int tx,
                                         double d2[4][4]) {
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int c = 0; c < 4; c++) d2[a][c] = 0.0;
    for (int k = 0; k < dp; k++) {
        const double2 ia = *reinterpret_cast<const double2*>(xt_i + k * CT + ty * 4);
        const double2 ib = *reinterpret_cast<const double2*>(xt_i + k * CT + ty * 4 + 2);
        const double2 ja = *reinterpret_cast<const double2*>(xt_j + k * CT + tx * 2);
        const double2 jb = *reinterpret_cast<const double2*>(xt_j + k * CT + tx * 2 + 32);
        const double xi[4] = {ia.x, ia.y, ib.x, ib.y};
        const double xj[4] = {ja.x, ja.y, jb.x, jb.y};
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double t = __dsub_rn(xi[a], xj[c]);
                d2[a][c] = FAST ? fma(t, t, d2[a][c]) : __dadd_rn(d2[a][c], __dmul_rn(t, t));
            }
    }
}

// sf2 * exp(-d2 * 0.5 / ell_sq).  FAST = false: same operation order as covkernel.cpp:89 (division kept as a division);
// FAST = true: one multiply by the precomputed -0.5 / ell_sq.
template <bool FAST>
__device__ __forceinline__ double se_kernel(double d2, const Hyper& h) {
    return h.sf2 * exp(FAST ? d2 * h.neg_half_inv_ell_sq : __ddiv_rn(__dmul_rn(-d2, 0.5), h.ell_sq));
}

// (68 registers: three CTAs per SM, 36 % occupancy, FP64 pipe 70 % and issue slots 67 % busy, profiles/r2_ncu_summaries.txt;
// forcing four CTAs per SM -- 64 registers, a few spills -- changed nothing: 2.34 vs 2.36 ms at n = 40 000)
template <int MODE, bool FAST>
__global__ void __launch_bounds__(COV_THREADS) cov_tile_kernel(const CovArgs p) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ double red[3 * (COV_THREADS / 32)];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t b = blockIdx.y;
    int ti, tj;
    if (MODE == COV_CROSS) {
        ti = blockIdx.x / p.tiles_j;
        tj = blockIdx.x % p.tiles_j;
    } else {
        lower_tile(blockIdx.x, ti, tj);
    }
    const int i0 = ti * CT, j0 = tj * CT;
    const double* Xi = p.Xi + b * p.sXi;
    const double* Xj = p.Xj + b * p.sXj;
    double *xt_i, *xt_j;
    stage_x_tiles(smem, Xi, i0, p.ni, Xj, j0, p.nj, p.dp, &bar, xt_i, xt_j);

    double d2[4][4];
    micro_d2<FAST>(xt_i, xt_j, p.dp, ty, tx, d2);

    if (MODE == COV_LOWER || MODE == COV_FULL) {
        double* K = p.out + b * p.sOut;
        double v[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            int gi = i0 + ty * 4 + a;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int gj = j0 + tx * 2 + 32 * (c >> 1) + (c & 1);
                double val = se_kernel<FAST>(d2[a][c], p.h);
                if (gi == gj) val += p.h.sn2;  // covkernel.cpp:93-94
                v[a][c] = val;
            }
            if (gi < p.ni) {
#pragma unroll
                for (int bb = 0; bb < 2; bb++) {
                    int gj = j0 + tx * 2 + 32 * bb;
                    double* dst = K + (int64_t)gi * p.ld + gj;
                    if (gj + 1 < p.nj) *reinterpret_cast<double2*>(dst) = make_double2(v[a][2 * bb], v[a][2 * bb + 1]);
                    else if (gj < p.nj) dst[0] = v[a][2 * bb];
                }
            }
        }
        if (MODE == COV_FULL && ti > tj) {  // mirror: covkernel.cpp:90-91 fills both triangles
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int gj = j0 + tx * 2 + 32 * (c >> 1) + (c & 1);
                if (gj >= p.nj) continue;
                int gi = i0 + ty * 4;
                double* dst = K + (int64_t)gj * p.ld + gi;
                if (gi + 3 < p.ni) {
                    *reinterpret_cast<double2*>(dst) = make_double2(v[0][c], v[1][c]);
                    *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2][c], v[3][c]);
                } else {
#pragma unroll
                    for (int a = 0; a < 4; a++)
                        if (gi + a < p.ni) dst[a] = v[a][c];
                }
            }
        }
    } else if (MODE == COV_CROSS) {
        double* Ks = p.out + b * p.sOut;
        const double* alpha = p.alpha + b * p.sAlpha;
        double al[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int gj = j0 + tx * 2 + 32 * (c >> 1) + (c & 1);
            al[c] = gj < p.nj ? alpha[gj] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            int gi = i0 + ty * 4 + a;
            double v[4], s = 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                v[c] = se_kernel<FAST>(d2[a][c], p.h);
                s += v[c] * al[c];
            }
            if (gi < p.ni) {
#pragma unroll
                for (int bb = 0; bb < 2; bb++) {
                    int gj = j0 + tx * 2 + 32 * bb;
                    double* dst = Ks + (int64_t)gi * p.ld + gj;
                    if (gj + 1 < p.nj) *reinterpret_cast<double2*>(dst) = make_double2(v[2 * bb], v[2 * bb + 1]);
                    else if (gj < p.nj) dst[0] = v[2 * bb];
                }
            }
#pragma unroll
            for (int off = 1; off < 16; off <<= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (tx == 0 && gi < p.ni) p.part[b * p.sPart + (int64_t)tj * p.ni + gi] = s;
        }
    } else {  // COV_TRACE
        const double* Kinv = p.Kinv + b * p.sKinv;
        const double* alpha = p.alpha + b * p.sAlpha;
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        double aj[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            int gj = j0 + tx * 2 + 32 * (c >> 1) + (c & 1);
            aj[c] = gj < p.nj ? alpha[gj] : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            int gi = i0 + ty * 4 + a;
            if (gi >= p.ni) continue;
            double ai = alpha[gi];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int gj = j0 + tx * 2 + 32 * (c >> 1) + (c & 1);
                if (gj > gi || gj >= p.nj) continue;
                double w = Kinv[(int64_t)gi * p.ld + gj] - ai * aj[c];  // W = Kinv - alpha alpha^T (covkernel.cpp:221)
                if (gi == gj) {
                    s2 += w * p.h.sf2;  // K_ii - sn2 = sf2 * exp(0)
                    s3 += w;
                } else {
                    double kv = se_kernel<FAST>(d2[a][c], p.h);
                    s1 += 2.0 * (w * (kv * (FAST ? d2[a][c] * p.h.inv_ell_sq : d2[a][c] / p.h.ell_sq)));  // covkernel.cpp:148,186,247
                    s2 += 2.0 * (w * kv);
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
            s3 += __shfl_xor_sync(0xffffffffu, s3, off);
        }
        const int warp = tid >> 5, lane = tid & 31;
        if (lane == 0) {
            red[warp] = s1;
            red[8 + warp] = s2;
            red[16 + warp] = s3;
        }
        __syncthreads();
        if (tid < 3) {
            double s = 0.0;
            for (int w = 0; w < COV_THREADS / 32; w++) s += red[tid * 8 + w];
            p.part[b * p.sPart + (int64_t)blockIdx.x * 3 + tid] = s;
        }
    }
}

// Fixed-order reduction of the per-tile trace partials -> gradient of the NEGATIVE log-likelihood.
__global__ void __launch_bounds__(256) grad_reduce_kernel(const double* part, int64_t sPart, int ntiles, Hyper h,
                                                          double* out) {
    __shared__ double red[3][256];
    const int64_t b = blockIdx.x;
    const double* pp = part + b * sPart;
    double s[3] = {0.0, 0.0, 0.0};
    for (int t = threadIdx.x; t < ntiles; t += 256)
        for (int c = 0; c < 3; c++) s[c] += pp[(int64_t)t * 3 + c];
    for (int c = 0; c < 3; c++) red[c][threadIdx.x] = s[c];
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off)
            for (int c = 0; c < 3; c++) red[c][threadIdx.x] += red[c][threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[b * 3 + 0] = red[0][0] / 2.0;        // psum1/2 (covkernel.cpp:259)
        out[b * 3 + 1] = red[1][0];              // psum2/2 = sum W.*(K - sn2 I)
        out[b * 3 + 2] = h.sn2 * red[2][0];      // psum3/2 = sn2 * tr W
    }
}

// Matrix-free residual r = (K(X,X) + sn2 I) v - y for the correctness check of a factorisation whose K has already been
// overwritten by L (bench.py `c5_residual`; the reference printed a residual after every factorisation,
// cuda_src/cuda_gp.cu:1126-1139).  One CTA owns 64 rows and walks all column tiles, rebuilding K tile by tile with the
// same arithmetic as K1; per-row sums are reduced in a fixed order.
__global__ void __launch_bounds__(COV_THREADS) cov_matvec_kernel(const double* __restrict__ X, int n, int dp, Hyper h,
                                                                 const double* __restrict__ v, const double* __restrict__ y,
                                                                 double* __restrict__ r) {
    extern __shared__ __align__(16) double smem[];
    double* xt_i = smem;                 // [dp][64]
    double* xt_j = smem + CT * dp;       // [dp][64]
    double* vj = xt_j + CT * dp;         // [64]
    double* red = vj + CT;               // [16][64]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * CT;
    for (int e = tid; e < CT * dp; e += COV_THREADS) {
        const int rr = e % CT, k = e / CT;
        xt_i[k * CT + rr] = (i0 + rr < n) ? X[(int64_t)(i0 + rr) * dp + k] : 0.0;
    }
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int j0 = 0; j0 < n; j0 += CT) {
        __syncthreads();
        for (int e = tid; e < CT * dp; e += COV_THREADS) {
            const int rr = e % CT, k = e / CT;
            xt_j[k * CT + rr] = (j0 + rr < n) ? X[(int64_t)(j0 + rr) * dp + k] : 0.0;
        }
        if (tid < CT) vj[tid] = (j0 + tid < n) ? v[j0 + tid] : 0.0;
        __syncthreads();
        double d2[4][4];
        micro_d2<false>(xt_i, xt_j, dp, ty, tx, d2);   // the CHECK keeps the reference's arithmetic
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int gi = i0 + ty * 4 + a;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int cj = tx * 2 + 32 * (c >> 1) + (c & 1), gj = j0 + cj;
                if (gj < n) {
                    double kv = se_kernel<false>(d2[a][c], h);
                    if (gi == gj) kv += h.sn2;
                    acc[a] += kv * vj[cj];
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; a++) red[tx * CT + ty * 4 + a] = acc[a];
    __syncthreads();
    if (tid < CT && i0 + tid < n) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < 16; t++) s += red[t * CT + tid];
        r[i0 + tid] = s - y[i0 + tid];
    }
}

template <int MODE, bool FAST>
void launch_cov_impl(const CovArgs& a, int64_t tiles, int batch, cudaStream_t st) {
    size_t smem = (size_t)4 * CT * a.dp * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        CUGP_CUDA(cudaFuncSetAttribute(cov_tile_kernel<MODE, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (tiles <= 0 || batch <= 0) return;
    cov_tile_kernel<MODE, FAST><<<dim3((unsigned)tiles, (unsigned)batch), COV_THREADS, smem, st>>>(a);
    CUGP_CUDA(cudaGetLastError());
}
template <int MODE>
void launch_cov(const CovArgs& a, int64_t tiles, int batch, cudaStream_t st) {
    if (a.h.fast) launch_cov_impl<MODE, true>(a, tiles, batch, st);
    else launch_cov_impl<MODE, false>(a, tiles, batch, st);
}

// ------------------------------------------------------------------------------------------------
// K2a: Cholesky + triangular inverse of a 128x128 diagonal block in shared memory (one CTA per matrix).
// S is 128 x 129: the lower triangle holds A -> L; the inverse is built in the strictly upper part,
// T(i,j) = S[j][i+1] for i >= j, so one 132 KB array serves both (odd row stride: column walks are
// conflict-free).
// ------------------------------------------------------------------------------------------------
constexpr int DB = kDiag;
constexpr int DLD = DB + 1;
constexpr int DIAG_THREADS = 512;

// The block is processed in four 32-column panels so the serial chain is 128 short warp-level column steps
// (registers + one shared broadcast per column; a column-at-a-time CTA-wide version with 256 barriers took
// 211 us, this one 40 us), and everything quadratic in the panel (SYRK inside the block, the off-diagonal
// blocks of the inverse) runs on DMMA over 8x8 output blocks spread across the 16 warps.
//   L(i,j) = S[i][j] (j <= i);   T(i,j) = inv(L)(i,j) = S[j][i+1] (j <= i).
constexpr int SB = 32;                 // panel width inside the diagonal block
constexpr int TLD = 65;                // row stride of the 64x64 scratch of the inverse recursion

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// C(i,j) = sum_k A(i,k) B(k,j) over 8x8 output blocks, one block per warp at a time (mb, nbk: counts of 8-blocks,
// both even).  The 8 rows (columns) of a block are an interleaved half of a 16-row group,
//     idx(b, g) = 16*(b/2) + 2*(b&1) + 4*(g&3) + (g>>2),
// so that with the odd row strides used here (129, 65) the 8-byte fragment loads of a half-warp fall in 16
// distinct bank pairs; with 8 consecutive rows they collide 4 ways and shared-memory bandwidth, not the DMMA
// pipe, bounds these small products.  `lower`: only blocks whose 16-group satisfies gi >= gj; the store sees
// real coordinates and masks j <= i itself.  krange(i16, j16, klo, khi) gives the k interval (multiples of 4)
// for the 16-groups starting at i16 / j16.  k is split over 8 independent accumulator chains: a dependent DMMA
// costs a few hundred clocks.
__device__ __forceinline__ int frag_idx(int b, int g) { return 16 * (b >> 1) + 2 * (b & 1) + 4 * (g & 3) + (g >> 2); }

template <class FA, class FB, class FK, class FS>
__device__ __forceinline__ void cta_gemm8(int mb, int nbk, bool lower, FA A, FB B, FK krange, FS store, int warp,
                                          int nwarps, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const int mg = mb >> 1, ng = nbk >> 1;  // 16-groups
    const int ngroups = lower ? mg * (mg + 1) / 2 : mg * ng;
    for (int x = warp; x < ngroups * 4; x += nwarps) {
        const int xg = x >> 2;
        int gi, gj;
        if (lower) {
            gi = (int)((sqrtf(8.f * (float)xg + 1.f) - 1.f) * 0.5f);
            while ((gi + 1) * (gi + 2) / 2 <= xg) gi++;
            while (gi * (gi + 1) / 2 > xg) gi--;
            gj = xg - gi * (gi + 1) / 2;
        } else {
            gi = xg / ng;
            gj = xg % ng;
        }
        const int bi = 2 * gi + ((x >> 1) & 1), bj = 2 * gj + (x & 1);
        const int ia = frag_idx(bi, g);                                  // A-fragment row of this lane
        const int jb = frag_idx(bj, g);                                  // B-fragment column of this lane
        int klo, khi;
        krange(16 * gi, 16 * gj, klo, khi);
        double c[8][2];
#pragma unroll
        for (int u = 0; u < 8; u++) c[u][0] = c[u][1] = 0.0;
        for (int k = klo; k < khi; k += 32) {
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (k + 4 * u < khi) dmma884(c[u][0], c[u][1], A(ia, k + 4 * u + q), B(k + 4 * u + q, jb));
        }
        const double c0 = ((c[0][0] + c[1][0]) + (c[2][0] + c[3][0])) + ((c[4][0] + c[5][0]) + (c[6][0] + c[7][0]));
        const double c1 = ((c[0][1] + c[1][1]) + (c[2][1] + c[3][1])) + ((c[4][1] + c[5][1]) + (c[6][1] + c[7][1]));
        // accumulator (g, 2q) and (g, 2q+1): row idx(bi, g), columns idx(bj, 2q) and idx(bj, 2q+1)
        store(ia, frag_idx(bj, 2 * q), frag_idx(bj, 2 * q + 1), c0, c1);
    }
}

__global__ void __launch_bounds__(DIAG_THREADS, 1)
    potrf_diag_blocked_kernel(double* A, int64_t ld, int64_t sA, int n, int j0, double* invd, int64_t sInvd,
                              double* logdet_part, int nblk, int blk, int factor, long long* prof) {
    if (blk < 0) {   // one launch over a range of diagonal blocks: blockIdx.y selects it, -1 - blk is the first one
        blk = blockIdx.y + (-1 - blk);
        j0 = blk * kDiag;
        invd += (int64_t)blk * kDiag * kDiag;
    }
    extern __shared__ __align__(16) double S[];
    double* tmp = S + DB * DLD;                // [64][TLD]
    double* rdiag = tmp + 64 * TLD;            // [128] reciprocals of L's diagonal
    double* colbuf = rdiag + DB;               // [2][32] broadcast buffer of the warp-level factorisation
    double* red = colbuf + 2 * SB;             // [128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = DIAG_THREADS / 32;
    const int64_t b = blockIdx.x;
    double* Ab = A + b * sA + (int64_t)j0 * ld + j0;
    const int nb = min(DB, n - j0);
    int stamp = 0;
#define DIAG_STAMP()                                                \
    do {                                                            \
        if (prof && tid == 0 && blockIdx.x == 0) prof[stamp] = clock64(); \
        stamp++;                                                    \
    } while (0)
    DIAG_STAMP();

    {
        // warp w owns rows w, w+16, ...; lane owns columns lane + 32c: all 32 loads of a thread are in flight together
        double v[DB / NW][4];
#pragma unroll
        for (int rr = 0; rr < DB / NW; rr++) {
            const int i = warp + NW * rr;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = lane + 32 * c;
                v[rr][c] = (j <= i && i < nb) ? Ab[(int64_t)i * ld + j] : 0.0;
            }
        }
#pragma unroll
        for (int rr = 0; rr < DB / NW; rr++) {
            const int i = warp + NW * rr;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int j = lane + 32 * c;
                S[i * DLD + j] = (i >= nb && i == j) ? 1.0 : v[rr][c];  // lower triangle of A, identity padded
            }
            if (lane == 0) S[i * DLD + DB] = 0.0;
        }
    }
    __syncthreads();
    DIAG_STAMP();

    if (!factor) {  // the block already holds a triangular factor: only its inverse is wanted
        if (tid < DB) rdiag[tid] = 1.0 / S[tid * DLD + tid];
        __syncthreads();
    }
    // ---------------- Cholesky, panel by panel ----------------
    for (int c0 = 0; factor && c0 < DB; c0 += SB) {
        if (warp == 0) {
            // 32x32 diagonal sub-block: lane i owns row i in registers (matrixops.cpp:74-98 on the sub-block).
            double a[SB];
#pragma unroll
            for (int c = 0; c < SB; c++) a[c] = (c <= lane) ? S[(c0 + lane) * DLD + c0 + c] : 0.0;
#pragma unroll
            for (int j = 0; j < SB; j++) {
                const double ajj = __shfl_sync(0xffffffffu, a[j], j);
                const double rd = rsqrt(ajj);        // negative pivot -> NaN, propagates (matrixops.cpp:77)
                const double d = ajj * rd;           // sqrt(ajj) to 2 ulp without the sqrt -> divide latency chain
                const double l = (lane == j) ? d : a[j] * rd;   // lanes < j hold 0
                a[j] = l;
                if (lane == j) rdiag[c0 + j] = rd;
                if (j + 1 < SB) {
                    if (lane == j + 1) a[j + 1] -= l * l;   // next pivot first: it does not wait for the broadcast
                    double* cb = colbuf + (j & 1) * SB;
                    cb[lane] = l;
                    __syncwarp();
#pragma unroll
                    for (int c = j + 1; c < SB; c++) {
                        const double lc = cb[c];     // L(c, j), broadcast
                        if (c == j + 1) {
                            if (lane != j + 1) a[c] -= l * lc;   // lane j+1 already applied l*l
                        } else {
                            a[c] -= l * lc;
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < SB; c++)
                if (c <= lane) S[(c0 + lane) * DLD + c0 + c] = a[c];
        }
        __syncthreads();
        DIAG_STAMP();
        const int r0 = c0 + SB, m = DB - r0;
        if (m <= 0) break;
        // rows below: X L11^T = A21 by forward substitution, one thread per row (matrixops.cpp:330-340 transposed)
        if (tid < m) {
            const int r = r0 + tid;
            double a[SB];
#pragma unroll
            for (int c = 0; c < SB; c++) a[c] = S[r * DLD + c0 + c];
#pragma unroll
            for (int c = 0; c < SB; c++) {
                const double x = a[c] * rdiag[c0 + c];
                a[c] = x;
#pragma unroll
                for (int k = c + 1; k < SB; k++) a[k] -= x * S[(c0 + k) * DLD + c0 + c];
            }
#pragma unroll
            for (int c = 0; c < SB; c++) S[r * DLD + c0 + c] = a[c];
        }
        __syncthreads();
        DIAG_STAMP();
        // trailing block of the diagonal block: A22 -= L21 L21^T (lower 8x8 blocks), K = 32
        cta_gemm8(
            m / 8, m / 8, true, [&](int i, int k) { return S[(r0 + i) * DLD + c0 + k]; },
            [&](int k, int j) { return S[(r0 + j) * DLD + c0 + k]; },
            [&](int, int, int& klo, int& khi) { klo = 0; khi = SB; },
            [&](int i, int ja, int jb, double v0, double v1) {
                double* row = S + (r0 + i) * DLD + r0;
                if (ja <= i) row[ja] -= v0;
                if (jb <= i) row[jb] -= v1;
            },
            warp, NW, lane);
        __syncthreads();
        DIAG_STAMP();
    }

    // ---------------- inverse ----------------
    stamp = 11;
    DIAG_STAMP();
    // diagonal 32x32 blocks: lane c solves L x = e_c (matrixops.cpp:330-340), 4 warps = 4 blocks
    if (warp < DB / SB) {
        const int c0 = warp * SB;
        double r[SB];
#pragma unroll
        for (int i = 0; i < SB; i++) r[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < SB; k++) {
            const double x = r[k] * rdiag[c0 + k];
            r[k] = x;
#pragma unroll
            for (int i = k + 1; i < SB; i++) r[i] -= S[(c0 + i) * DLD + c0 + k] * x;
        }
#pragma unroll
        for (int k = 0; k < SB; k++)
            if (k >= lane) S[(c0 + lane) * DLD + c0 + k + 1] = r[k];  // T(c0+k, c0+lane)
    }
    __syncthreads();
    DIAG_STAMP();
    // off-diagonal blocks by doubling: T21 = -T22 (L21 T11)
    for (int h = SB; h < DB; h *= 2) {
        const int npairs = DB / (2 * h);
        const int hb = h / 8;
        for (int pr = 0; pr < npairs; pr++) {   // pairs are processed one after the other (tmp is one buffer)
            const int r1 = pr * 2 * h, r2 = r1 + h;
            // tmp = L21 T11 : k in [j0 rounded down to 4, h)  (T11 lower triangular)
            cta_gemm8(
                hb, hb, false, [&](int i, int k) { return S[(r2 + i) * DLD + r1 + k]; },
                [&](int k, int j) { return k >= j ? S[(r1 + j) * DLD + r1 + k + 1] : 0.0; },
                [&](int, int jj, int& klo, int& khi) { klo = jj; khi = h; },
                [&](int i, int ja, int jb, double v0, double v1) {
                    tmp[i * TLD + ja] = v0;
                    tmp[i * TLD + jb] = v1;
                },
                warp, NW, lane);
            __syncthreads();
            // T21 = -T22 tmp : k in [0, i0 + 8)  (T22 lower triangular)
            cta_gemm8(
                hb, hb, false, [&](int i, int k) { return k <= i ? S[(r2 + k) * DLD + r2 + i + 1] : 0.0; },
                [&](int k, int j) { return tmp[k * TLD + j]; },
                [&](int ii, int, int& klo, int& khi) { klo = 0; khi = ii + 16; },
                [&](int i, int ja, int jb, double v0, double v1) {
                    S[(r1 + ja) * DLD + r2 + i + 1] = -v0;       // T(r2+i, r1+j)
                    S[(r1 + jb) * DLD + r2 + i + 1] = -v1;
                },
                warp, NW, lane);
            __syncthreads();
        }
        DIAG_STAMP();
    }

    // write back: L11 with zeroed upper triangle, inv(L11) dense 128x128 (zero upper)
    double* inv = invd + b * sInvd;
#pragma unroll
    for (int rr = 0; rr < DB / NW; rr++) {
        const int i = warp + NW * rr;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int j = lane + 32 * c;
            if (j <= i) {  // upper triangles: invd is zeroed at allocation, L's is never read (exports mask it)
                if (factor && i < nb) Ab[(int64_t)i * ld + j] = S[i * DLD + j];
                inv[i * DB + j] = S[j * DLD + i + 1];
            }
        }
    }
    DIAG_STAMP();
#undef DIAG_STAMP
    if (!logdet_part) return;
    if (tid < DB) red[tid] = log(S[tid * DLD + tid]);
    __syncthreads();
    for (int off = DB / 2; off > 0; off >>= 1) {
        if (tid < off) red[tid] += red[tid + off];
        __syncthreads();
    }
    if (tid == 0) logdet_part[b * nblk + blk] = red[0];
}

// ------------------------------------------------------------------------------------------------
// K3: triangular solves.  The forward substitution L z = y is fused into the Cholesky (y^T is row n of the
// matrix, see gp.cu).  alpha = L^-T z is either one streaming pass over T = L^-1 (gemv_t_kernel) or the
// two-level backward sweep of launch_trsv_backward: per 1024-row panel one cluster launch for the chain of 128-row
// block solves (trsv_bwd_chain_kernel) and single-owner streaming updates of everything left of the panel
// (trsv_bwd_update_kernel), so the sweep streams L exactly once.  trsv_bwd_step_kernel is the older one-launch-per-block
// chain (tuning "bwd_cluster" = 0), kept as the reference the cluster kernel is tested against.
// ------------------------------------------------------------------------------------------------
constexpr int TRSV_THREADS = 256;
constexpr int BWD_STEP_COLS = 128;  // columns per CTA of the step kernel (4 row groups x 64 column pairs)

// One 128-row block of the backward sweep.  The launch is latency bound (a chain of dependent global round
// trips), so every load batch is issued in full before its first use: 32 independent loads per thread.
__global__ void __launch_bounds__(TRSV_THREADS)
    trsv_bwd_step_kernel(const double* __restrict__ L, int64_t ld, int64_t sL, int n, int j0, int c_lo,
                         const double* __restrict__ invd, int64_t sInvd, double* work, double* alpha, int64_t sVec) {
    __shared__ double sj[DB];
    __shared__ double aj[DB];
    __shared__ double part[4][BWD_STEP_COLS];
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    L += b * sL;
    invd += b * sInvd + (int64_t)(j0 / DB) * DB * DB;
    work += b * sVec;
    alpha += b * sVec;
    const int nb = min(DB, n - j0);
    if (tid < DB) sj[tid] = tid < nb ? work[j0 + tid] : 0.0;
    // alpha_j = inv(L_jj)^T s_j : thread (c, half) sums 64 rows of column c; the loads do not depend on s_j, so they
    // are in flight while it arrives (the zero upper triangle of the stored inverse makes the full loop exact)
    {
        const int c = tid & (DB - 1), hlf = tid >> 7;
        const double* col = invd + (hlf * (DB / 2)) * DB + c;
        double s = 0.0;
#pragma unroll
        for (int r0 = 0; r0 < DB / 2; r0 += 32) {
            double v[32];
#pragma unroll
            for (int u = 0; u < 32; u++) v[u] = col[(r0 + u) * DB];
            if (r0 == 0) __syncthreads();  // s_j is in shared memory
#pragma unroll
            for (int u = 0; u < 32; u++) s += v[u] * sj[hlf * (DB / 2) + r0 + u];
        }
        part[hlf][c] = s;
    }
    __syncthreads();
    if (tid < DB) aj[tid] = part[0][tid] + part[1][tid];
    __syncthreads();
    if (blockIdx.x == 0 && tid < nb) alpha[j0 + tid] = aj[tid];
    // s[c] -= sum_r L[j0+r][c] alpha_j[r] for this CTA's columns in [c_lo, j0) (c_lo: start of the outer panel):
    // thread (column pair, row group) covers 32 rows of 2 columns
    const int cp = tid & 63, rg = tid >> 6;
    const int c = c_lo + blockIdx.x * BWD_STEP_COLS + cp * 2;
    double a0 = 0.0, a1 = 0.0;
    if (c < j0) {
        const double* src = L + (int64_t)(j0 + rg * 32) * ld + c;
        double2 v[32];
#pragma unroll
        for (int u = 0; u < 32; u++)
            v[u] = (rg * 32 + u < nb) ? *reinterpret_cast<const double2*>(src + (int64_t)u * ld) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 32; u++) {
            a0 += v[u].x * aj[rg * 32 + u];
            a1 += v[u].y * aj[rg * 32 + u];
        }
    }
    __syncthreads();  // part[] is reused
    part[rg][cp * 2] = a0;
    part[rg][cp * 2 + 1] = a1;
    __syncthreads();
    if (tid < BWD_STEP_COLS) {
        const int cc = c_lo + blockIdx.x * BWD_STEP_COLS + tid;
        if (cc < j0) work[cc] -= (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    }
}

// The whole chain of one 1024-row panel in ONE launch: a thread-block cluster of up to 8 CTAs, CTA c owning the
// panel's c-th block of 128 columns -- its slice s_c of the right-hand side and the inverse of its diagonal block
// live in that CTA's shared memory for the whole launch.  Step j (from the bottom block up):
//   CTA j:    alpha_j = inv(L_jj)^T s_j      (128x128 mat-vec out of shared memory)
//             -> global alpha, and into the landing buffer of every CTA c < j through distributed shared memory
//   cluster barrier (release/acquire)
//   CTA c<j:  s_c -= L[block j, block c]^T alpha_j, from registers: the 128x128 block of L was requested one step
//             earlier, before the barrier wait (its address does not depend on alpha), so the HBM/L2 round trip of
//             every step overlaps the previous step's dependent work.
// This replaces 8 dependent launches (each a chain of global round trips: right-hand side, inverse block, L block,
// write-back) by one; the per-step cost becomes a shared-memory mat-vec plus one cluster barrier.
namespace cgs = cooperative_groups;
constexpr int PC_THREADS = 512;
constexpr size_t PC_SMEM = (size_t)(DB * DB + DB + 2 * DB + 8 * DB) * sizeof(double);

__global__ void __launch_bounds__(PC_THREADS, 1)
    trsv_bwd_chain_kernel(const double* __restrict__ L, int64_t ld, int64_t sL, int n, int P0, int nbk,
                          const double* __restrict__ invd, int64_t sInvd, const double* __restrict__ work, double* alpha,
                          int64_t sVec) {
    cgs::cluster_group cluster = cgs::this_cluster();
    extern __shared__ __align__(16) double pcs[];
    double* inv = pcs;                  // [128][128] inverse of this CTA's diagonal block (zero upper triangle)
    double* s = inv + DB * DB;          // [128] this CTA's slice of the right-hand side
    double* abuf = s + DB;              // [2][128] landing buffers for alpha_j (alternating by step parity)
    double* part = abuf + 2 * DB;       // [8][128] partial sums
    const int tid = threadIdx.x;
    const int c = (int)cluster.block_rank();   // == blockIdx.x: grid.x equals the cluster width
    const int64_t b = blockIdx.y;
    L += b * sL;
    work += b * sVec;
    alpha += b * sVec;
    const int jc = P0 + c * DB;                // first row / column of this CTA's block
    const int nbc = min(DB, n - jc);
    {
        const double2* src = reinterpret_cast<const double2*>(invd + b * sInvd + (int64_t)(jc / DB) * DB * DB);
        double2 v[DB * DB / 2 / PC_THREADS];
#pragma unroll
        for (int i = 0; i < DB * DB / 2 / PC_THREADS; i++) v[i] = src[tid + PC_THREADS * i];
#pragma unroll
        for (int i = 0; i < DB * DB / 2 / PC_THREADS; i++) reinterpret_cast<double2*>(inv)[tid + PC_THREADS * i] = v[i];
        if (tid < DB) s[tid] = tid < nbc ? work[jc + tid] : 0.0;
    }
    // thread (column pair cp, row group rg): 16 rows x 2 columns of the 128x128 block of L below this CTA's block
    const int cp = tid & 63, rg = tid >> 6;
    double2 lv[16];
    auto request = [&](int j) {   // block (j, c): rows P0 + 128 j + ..., columns jc + ...
        const int r0 = P0 + j * DB + rg * 16;
        const double* src = L + (int64_t)r0 * ld + jc + 2 * cp;
#pragma unroll
        for (int u = 0; u < 16; u++)
            lv[u] = (r0 + u < n) ? *reinterpret_cast<const double2*>(src + (int64_t)u * ld) : make_double2(0.0, 0.0);
    };
    if (c < nbk - 1) request(nbk - 1);
    cluster.sync();   // every CTA's shared memory is initialised before the first remote store
    for (int j = nbk - 1; j >= 0; j--) {
        double* ab = abuf + (j & 1) * DB;
        if (c == j) {
            // alpha_j[col] = sum_r inv[r][col] s[r]: thread (col, q) takes rows q, q+4, ...
            const int col = tid & (DB - 1), q = tid >> 7;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
            for (int r = q; r < DB; r += 8) {
                a0 += inv[r * DB + col] * s[r];
                a1 += inv[(r + 4) * DB + col] * s[r + 4];
            }
            part[q * DB + col] = a0 + a1;
            __syncthreads();
            if (tid < DB) {
                const double a = (part[tid] + part[DB + tid]) + (part[2 * DB + tid] + part[3 * DB + tid]);
                if (tid < nbc) alpha[jc + tid] = a;
                for (int dst = 0; dst < j; dst++) cluster.map_shared_rank(ab, dst)[tid] = a;
            }
        }
        cluster.sync();
        if (c < j) {
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const double av = ab[rg * 16 + u];
                a0 += lv[u].x * av;
                a1 += lv[u].y * av;
            }
            if (j - 1 > c) request(j - 1);   // next step's block: in flight across the next barrier
            part[rg * DB + 2 * cp] = a0;
            part[rg * DB + 2 * cp + 1] = a1;
            __syncthreads();
            if (tid < DB) {
                double t = 0.0;
#pragma unroll
                for (int g = 0; g < 8; g++) t += part[g * DB + tid];
                s[tid] -= t;
            }
            __syncthreads();
        }
    }
}

// Outer step of the two-level backward sweep: the columns [c_lo, c_hi) left of a finished panel of rows
// [P0, P0 + rows) receive s[c] -= sum_r L[P0+r][c] alpha[P0+r].  One CTA owns 64 columns for ALL rows of the panel
// (up to 8 sub-blocks of 128 rows, 16 independent 16-byte loads per thread each, two CTAs per SM), so every column
// is reduced by exactly one CTA in a fixed order and written once: no partial sums, no second kernel, no atomics.
// (A 2-D grid with per-chunk partials and a reduction launch streamed the slab at 4.7 TB/s; the extra launch
// boundary per panel drained the GPU twice as often.)  The per-128-row steps inside the panel only touch the panel's
// own columns and stay in L2.
constexpr int BWD_PANEL = 1024;
constexpr int BWD_PCHUNKS = BWD_PANEL / DB;
constexpr int BWD_UCOLS = 64;
__global__ void __launch_bounds__(TRSV_THREADS, 2)
    trsv_bwd_update_kernel(const double* __restrict__ L, int64_t ld, int64_t sL, int P0, int rows, int c_lo, int c_hi,
                           const double* __restrict__ alpha, double* work, int64_t sVec) {
    __shared__ double ap[BWD_PANEL];
    __shared__ double red[8][BWD_UCOLS];
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.y;
    L += b * sL;
    alpha += b * sVec;
    work += b * sVec;
    for (int i = tid; i < BWD_PANEL; i += TRSV_THREADS) ap[i] = i < rows ? alpha[P0 + i] : 0.0;
    const int cp = tid & 31, rg = tid >> 5;                // column pair, row group (16 rows of every 128)
    const int c = c_lo + blockIdx.x * BWD_UCOLS + cp * 2;  // c_lo, c_hi are multiples of 64: whole CTAs are valid
    __syncthreads();
    double a0 = 0.0, a1 = 0.0;
    const int nsb = (rows + DB - 1) / DB;
    const double* src = L + (int64_t)(P0 + rg * 16) * ld + c;
    for (int sb = 0; sb < nsb; sb++, src += (int64_t)DB * ld) {
        double2 v[16];
#pragma unroll
        for (int u = 0; u < 16; u++)
            v[u] = (sb * DB + rg * 16 + u < rows) ? *reinterpret_cast<const double2*>(src + (int64_t)u * ld)
                                                   : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const double av = ap[sb * DB + rg * 16 + u];
            a0 += v[u].x * av;
            a1 += v[u].y * av;
        }
    }
    red[rg][cp * 2] = a0;
    red[rg][cp * 2 + 1] = a1;
    __syncthreads();
    if (tid < BWD_UCOLS) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 8; g++) t += red[g][tid];
        work[c_lo + blockIdx.x * BWD_UCOLS + tid] -= t;
    }
}

// alpha = T^T z for lower-triangular T = L^-1: alpha[c] = sum_{r >= c} T[r][c] z[r].  T is streamed once by a 2-D
// grid: CTA (column block cb of 128, row chunk k) covers rows [(cb + k*rpc) * 128, +rpc*128) of its 128 columns in
// 128-row sub-blocks, 32 independent 16-byte loads per thread each (a column-per-CTA version was bound by its longest
// CTA: n/8 dependent round trips at 4 loads in flight, 2.5 TB/s at n = 16 000); the chunks leave partial sums that a
// second small kernel adds in a fixed order.  chunks <= 16 so the partials fit the backward sweep's scratch.
constexpr int GT_CHUNKS = 2 * BWD_PCHUNKS;
__global__ void __launch_bounds__(TRSV_THREADS)
    gemv_t_kernel(const double* __restrict__ T, int64_t ld, int64_t sT, int n, int rpc, const double* __restrict__ z,
                  int64_t sZ, double* partial, int64_t sPart) {
    __shared__ double zs[DB];
    __shared__ double red[4][DB];
    const int tid = threadIdx.x;
    const int64_t b = blockIdx.z;
    const int cb = blockIdx.x, k = blockIdx.y;
    const int r_begin = (cb + k * rpc) * DB;
    if (r_begin >= n) return;                       // uniform per CTA: this chunk lies below the matrix
    T += b * sT;
    z += b * sZ;
    const int cp = tid & 63, rg = tid >> 6;         // column pair, row group (32 rows each)
    const int c = cb * DB + 2 * cp;
    double a0 = 0.0, a1 = 0.0;
    for (int sb = 0; sb < rpc; sb++) {
        const int r0 = r_begin + sb * DB;
        if (r0 >= n) break;
        __syncthreads();                            // zs is reused
        if (tid < DB) zs[tid] = (r0 + tid < n) ? z[r0 + tid] : 0.0;
        const double* src = T + (int64_t)(r0 + rg * 32) * ld + c;
        double2 v[32];
#pragma unroll
        for (int u = 0; u < 32; u++)
            v[u] = (r0 + rg * 32 + u < n) ? *reinterpret_cast<const double2*>(src + (int64_t)u * ld) : make_double2(0.0, 0.0);
        __syncthreads();
        if (r0 == cb * DB) {                        // diagonal block: keep r >= c only
#pragma unroll
            for (int u = 0; u < 32; u++) {
                const int r = r0 + rg * 32 + u;
                a0 += (r >= c ? v[u].x : 0.0) * zs[rg * 32 + u];
                a1 += (r >= c + 1 ? v[u].y : 0.0) * zs[rg * 32 + u];
            }
        } else {
#pragma unroll
            for (int u = 0; u < 32; u++) {
                a0 += v[u].x * zs[rg * 32 + u];
                a1 += v[u].y * zs[rg * 32 + u];
            }
        }
    }
    red[rg][2 * cp] = a0;
    red[rg][2 * cp + 1] = a1;
    __syncthreads();
    if (tid < DB) {
        const int cc = cb * DB + tid;
        if (cc < n) partial[b * sPart + (int64_t)k * n + cc] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    }
}

__global__ void gemv_t_reduce_kernel(const double* partial, int64_t sPart, int n, int rpc, double* alpha, int64_t sAlpha) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int nblk = (n + DB - 1) / DB, cb = c / DB;
    const int chunks = (nblk - cb + rpc - 1) / rpc;   // chunks that exist for this column block
    const double* p = partial + (int64_t)blockIdx.y * sPart + c;
    double s = 0.0;
    for (int k = 0; k < chunks; k++) s += p[(int64_t)k * n];
    alpha[(int64_t)blockIdx.y * sAlpha + c] = s;
}

__global__ void init_identity_kernel(double* R, int64_t ld, int64_t sR, int n) {
    const int64_t b = blockIdx.z;
    const int i = blockIdx.y;
    double2* row = reinterpret_cast<double2*>(R + b * sR + (int64_t)i * ld);
    for (int c2 = blockIdx.x * blockDim.x + threadIdx.x; c2 < ld / 2; c2 += gridDim.x * blockDim.x)
        row[c2] = make_double2(2 * c2 == i ? 1.0 : 0.0, 2 * c2 + 1 == i ? 1.0 : 0.0);
}

// Identity rows + K^-1 accumulator in one pass, and only what is ever read (gp.cu: potrf_with_rhs): row i of the identity
// block from the start of the block column LEFT of its own (the step's prologue reads one block column back; everything
// further left is never touched: the rows join at step i / 128), row i of the accumulator up to the end of its own
// 128-column block (lower tiles of the rank-128 updates).
__global__ void init_idrows_kernel(double* R, int64_t sR, double* W, int64_t sW, int64_t ld, int n) {
    const int64_t b = blockIdx.z;
    const int i = blockIdx.y;
    const int cs = max(0, (i / kDiag - 1) * kDiag), ce = min((int)ld, (i / kDiag + 1) * kDiag);
    double2* row = reinterpret_cast<double2*>(R + b * sR + (int64_t)i * ld);
    double2* wrow = reinterpret_cast<double2*>(W + b * sW + (int64_t)i * ld);
    for (int c2 = blockIdx.x * blockDim.x + threadIdx.x; c2 < ld / 2; c2 += gridDim.x * blockDim.x) {
        if (2 * c2 >= cs) row[c2] = make_double2(2 * c2 == i ? 1.0 : 0.0, 2 * c2 + 1 == i ? 1.0 : 0.0);
        if (2 * c2 < ce) wrow[c2] = make_double2(0.0, 0.0);
    }
}

__global__ void __launch_bounds__(256) gemv_upper_kernel(const double* __restrict__ U, int64_t ld, int64_t sU, int n,
                                                         const double* __restrict__ z, int64_t sZ, double* alpha, int64_t sAlpha) {
    const int64_t b = blockIdx.y;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const double* row = U + b * sU + (int64_t)i * ld;
    const double* zz = z + b * sZ;
    double s = 0.0;
    for (int k = (i & ~1) + 2 * lane; k < n; k += 64) {   // 16-byte loads from the even column at or before the diagonal
        const double2 u = *reinterpret_cast<const double2*>(row + k);
        if (k >= i) s += u.x * zz[k];
        if (k + 1 < n) s += u.y * zz[k + 1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) alpha[b * sAlpha + i] = s;
}

// dst[b][0..n) = src[b][0..n) with independent batch strides (y -> row n of the matrix, z -> work vector)
__global__ void copy_rows_kernel(const double* src, int64_t sSrc, double* dst, int64_t sDst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[(int64_t)blockIdx.y * sDst + i] = src[(int64_t)blockIdx.y * sSrc + i];
}

__global__ void __launch_bounds__(256)
    ll_finalize_kernel(const double* y, const double* alpha, int64_t sVec, int n, const double* logdet_part, int nblk,
                       double* scal) {
    __shared__ double red[256];
    const int64_t b = blockIdx.x;
    y += b * sVec;
    alpha += b * sVec;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += y[i] * alpha[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double ld = 0.0;
        for (int k = 0; k < nblk; k++) ld += logdet_part[b * nblk + k];
        ld = 2 * ld;  // matrixops.cpp:133-136
        double quad = red[0];
        scal[b * 4 + 0] = quad;
        scal[b * 4 + 1] = ld;
        // covkernel.cpp:127: -0.5 * (first + second + n * 1.83787), truncated log(2 pi); no FMA contraction so the
        // scalar arithmetic is the reference's to the last bit
        scal[b * 4 + 2] = -0.5 * __dadd_rn(__dadd_rn(quad, ld), __dmul_rn((double)n, 1.83787));
        scal[b * 4 + 3] = 0.0;
    }
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void copy_vec_kernel(const double* src, double* dst, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[i];
}

__global__ void scatter_invdiag_kernel(const double* invd, int64_t sInvd, double* T, int64_t ld, int64_t sT, int n) {
    const int blk = blockIdx.x;
    const int64_t b = blockIdx.y;
    const int j0 = blk * DB, nb = min(DB, n - j0);
    const double* src = invd + b * sInvd + (int64_t)blk * DB * DB;
    double* dst = T + b * sT + (int64_t)j0 * ld + j0;
    for (int e = threadIdx.x; e < DB * DB; e += blockDim.x) {
        int i = e / DB, j = e % DB;
        if (i < nb && j < nb) dst[(int64_t)i * ld + j] = src[e];
    }
}

// mode 0: zero the upper triangle; 1: mirror the lower triangle; 2: plain copy
__global__ void export_kernel(const double* A, int64_t ld, int n, double* out, int mode) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int i = blockIdx.y; i < n; i += gridDim.y) {
        double v;
        if (j <= i || mode == 2) v = A[(int64_t)i * ld + j];
        else v = mode == 1 ? A[(int64_t)j * ld + i] : 0.0;
        out[(int64_t)i * n + j] = v;
    }
}
__global__ void import_kernel(const double* in, int n, double* A, int64_t ld) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (int i = blockIdx.y; i < n; i += gridDim.y) A[(int64_t)i * ld + j] = in[(int64_t)i * n + j];
}

__global__ void predict_finalize_kernel(const double* meanpart, int ntile_mean, const double* css, int ntile_css, int m,
                                        Hyper h, double* mean, double* var, int64_t sOut, int64_t sMp, int64_t sCss) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.y;
    if (t >= m) return;
    double mu = 0.0, q = 0.0;
    for (int k = 0; k < ntile_mean; k++) mu += meanpart[b * sMp + (int64_t)k * m + t];
    for (int k = 0; k < ntile_css; k++) q += css[b * sCss + (int64_t)k * m + t];
    mean[b * sOut + t] = mu;
    var[b * sOut + t] = (h.sf2 + h.sn2) - q;  // covkernel.cpp:299-302: k** includes the noise term
}

__global__ void poe_accumulate_kernel(const double* mean, const double* var, int64_t sOut, int nexp, int m, double* Pp,
                                      double* Qp, int accumulate) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    double P = accumulate ? Pp[t] : 0.0, Q = accumulate ? Qp[t] : 0.0;
    for (int e = 0; e < nexp; e++) {  // BCM.cpp:51-55
        double invvar = 1.0 / var[e * sOut + t];
        P += invvar;
        Q += invvar * mean[e * sOut + t];
    }
    Pp[t] = P;
    Qp[t] = Q;
}
__global__ void poe_finalize_kernel(const double* PQ, int m, double* mean, double* var) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    double tempvar = 1.0 / PQ[t];  // BCM.cpp:56-57
    mean[t] = tempvar * PQ[m + t];
    var[t] = tempvar;
}

}  // namespace

static int g_cov_fast = 1;
int cov_fast_default() { return g_cov_fast; }
void set_cov_fast(int v) { g_cov_fast = v; }

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
void launch_cov_train(const double* X, int64_t sX, int n, int dp, Hyper h, double* K, int64_t ld, int64_t sK, int batch,
                      int full, cudaStream_t st) {
    CovArgs a{};
    a.Xi = X; a.sXi = sX; a.ni = n;
    a.Xj = X; a.sXj = sX; a.nj = n;
    a.dp = dp; a.h = h;
    a.out = K; a.ld = ld; a.sOut = sK;
    int64_t t = cdiv(n, CT);
    if (full) launch_cov<COV_FULL>(a, t * (t + 1) / 2, batch, st);
    else launch_cov<COV_LOWER>(a, t * (t + 1) / 2, batch, st);
}

void launch_cov_residual(const double* X, int n, int dp, Hyper h, const double* v, const double* y, double* r, cudaStream_t st) {
    const size_t smem = (size_t)(2 * CT * dp + CT + 16 * CT) * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        CUGP_CUDA(cudaFuncSetAttribute(cov_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cov_matvec_kernel<<<cdiv(n, CT), COV_THREADS, smem, st>>>(X, n, dp, h, v, y, r);
    CUGP_CUDA(cudaGetLastError());
}

void launch_cov_cross(const double* Xt, int m, const double* X, int64_t sX, int n, int dp, Hyper h, const double* alpha,
                      int64_t sAlpha, double* Kstar, int64_t ldk, int64_t sKs, double* meanpart, int64_t sMp, int batch,
                      cudaStream_t st) {
    CovArgs a{};
    a.Xi = Xt; a.sXi = 0; a.ni = m;
    a.Xj = X; a.sXj = sX; a.nj = n;
    a.dp = dp; a.h = h;
    a.out = Kstar; a.ld = ldk; a.sOut = sKs;
    a.alpha = alpha; a.sAlpha = sAlpha;
    a.part = meanpart; a.sPart = sMp;
    a.tiles_j = cdiv(n, CT);
    launch_cov<COV_CROSS>(a, (int64_t)cdiv(m, CT) * a.tiles_j, batch, st);
}

size_t grad_trace_partials(int n, int batch) {
    int64_t t = cdiv(n, CT);
    return (size_t)(t * (t + 1) / 2) * 3 * (size_t)batch;
}

void launch_grad_trace(const double* X, int64_t sX, int n, int dp, Hyper h, const double* Kinv, int64_t ld, int64_t sKinv,
                       const double* alpha, int64_t sAlpha, double* partials, double* out, int batch, cudaStream_t st) {
    CovArgs a{};
    a.Xi = X; a.sXi = sX; a.ni = n;
    a.Xj = X; a.sXj = sX; a.nj = n;
    a.dp = dp; a.h = h; a.ld = ld;
    a.alpha = alpha; a.sAlpha = sAlpha;
    a.Kinv = Kinv; a.sKinv = sKinv;
    int64_t t = cdiv(n, CT), tiles = t * (t + 1) / 2;
    a.part = partials; a.sPart = tiles * 3;
    launch_cov<COV_TRACE>(a, tiles, batch, st);
    grad_reduce_kernel<<<batch, 256, 0, st>>>(partials, tiles * 3, (int)tiles, h, out);
    CUGP_CUDA(cudaGetLastError());
}

void launch_potrf_diag(double* A, int64_t ld, int64_t sA, int n, int j0, double* invd, int64_t sInvd, double* logdet_part,
                       int nblk, int blk, int batch, cudaStream_t st) {
    constexpr size_t smem = (size_t)(DB * DLD + 64 * TLD + DB + 2 * SB + DB) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        CUGP_CUDA(cudaFuncSetAttribute(potrf_diag_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    potrf_diag_blocked_kernel<<<batch, DIAG_THREADS, smem, st>>>(A, ld, sA, n, j0, invd + (int64_t)blk * DB * DB, sInvd,
                                                                 logdet_part, nblk, blk, 1, nullptr);
    CUGP_CUDA(cudaGetLastError());
}

// Phase timestamps (clock64 of thread 0) of one diagonal-block factorisation: tuning aid, see tools/diag_phases.py.
void debug_diag_phases(double* A, int64_t ld, int n, double* invd, double* logdet, long long* stamps_dev, cudaStream_t st) {
    constexpr size_t smem = (size_t)(DB * DLD + 64 * TLD + DB + 2 * SB + DB) * sizeof(double);
    CUGP_CUDA(cudaFuncSetAttribute(potrf_diag_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    potrf_diag_blocked_kernel<<<1, DIAG_THREADS, smem, st>>>(A, ld, 0, n, 0, invd, 0, logdet, 1, 0, 1, stamps_dev);
    CUGP_CUDA(cudaGetLastError());
}

void launch_trtri_diag(const double* L, int64_t ld, int64_t sL, int n, double* invd, int64_t sInvd, int batch,
                       cudaStream_t st, int blk0, int nblocks) {
    constexpr size_t smem = (size_t)(DB * DLD + 64 * TLD + DB + 2 * SB + DB) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        CUGP_CUDA(cudaFuncSetAttribute(potrf_diag_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    // one launch, blockIdx.y = diagonal block (blk0 ..)
    if (nblocks <= 0) nblocks = cdiv(n, DB) - blk0;
    if (nblocks <= 0) return;
    potrf_diag_blocked_kernel<<<dim3(batch, nblocks), DIAG_THREADS, smem, st>>>(const_cast<double*>(L), ld, sL, n, 0, invd,
                                                                                 sInvd, nullptr, 0, -1 - blk0, 0, nullptr);
    CUGP_CUDA(cudaGetLastError());
}

// partial sums of gemv_t (the sweep itself needs none)
size_t trsv_backward_scratch(int n, int batch) { return (size_t)2 * BWD_PCHUNKS * n * batch; }

int trsv_backward_events(int n) { return 2 * cdiv(n, BWD_PANEL) + 2; }

static int g_bwd_cluster = 1;   // 1: one cluster launch per panel chain; 0: one launch per 128-row block (the older path)
void set_bwd_cluster(int v) { g_bwd_cluster = v; }

static void launch_bwd_chain(const double* L, int64_t ld, int64_t sL, int n, int P0, int nbk, const double* invd, int64_t sInvd,
                             const double* work, double* alpha, int64_t sVec, int batch, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        CUGP_CUDA(cudaFuncSetAttribute(trsv_bwd_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PC_SMEM));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)nbk, (unsigned)batch);
    cfg.blockDim = dim3(PC_THREADS);
    cfg.dynamicSmemBytes = PC_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nbk;   // <= 8: portable cluster size
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUGP_CUDA(cudaLaunchKernelEx(&cfg, trsv_bwd_chain_kernel, L, ld, sL, n, P0, nbk, invd, sInvd, work, alpha, sVec));
}

namespace {
// work[c_lo:c_hi) -= L[P0:P0+rows, c_lo:c_hi)^T alpha[P0:P0+rows)
void bwd_panel_update(const double* L, int64_t ld, int64_t sL, int P0, int rows, int c_lo, int c_hi, double* work,
                      const double* alpha, int64_t sVec, int batch, cudaStream_t st) {
    if (c_hi <= c_lo) return;
    trsv_bwd_update_kernel<<<dim3(cdiv(c_hi - c_lo, BWD_UCOLS), batch), TRSV_THREADS, 0, st>>>(L, ld, sL, P0, rows, c_lo, c_hi,
                                                                                              alpha, work, sVec);
}
}  // namespace

// Two-level backward sweep with look-ahead.  Panel p = rows [P0, P1) (1024 rows):
//   steps(p): the chain of 128-row block solves inside the panel (latency bound, one launch per block)
//   U1(p):    the panel's update of the NEXT panel's columns [P0-1024, P0)      -- on the chain's stream
//   U2(p):    its update of everything left of that, [0, P0-1024): the bulk of the HBM traffic -- stays on `st`,
//             concurrent with steps(p-1).
// work[c] receives its updates in the same order as in the single-stream sweep (U2(p+1) before U1(p), U2's in
// panel order), so the result does not depend on the overlap.  st_chain == nullptr: everything on `st`.
void launch_trsv_backward(const double* L, int64_t ld, int64_t sL, int n, const double* invd, int64_t sInvd, double* work,
                          double* alpha, int64_t sVec, double* scratch, int batch, cudaStream_t st, cudaStream_t st_chain,
                          cudaEvent_t* ev, int nev) {
    const int npanels = cdiv(n, BWD_PANEL);
    const bool overlap = st_chain != nullptr && npanels >= 3 && nev >= 2 * npanels + 2;
    (void)scratch;
    // the caller's work vector is ready on `st`; with overlap the chain runs on `st_chain` (the high-priority
    // look-ahead stream) and the bulk updates stay on `st`
    cudaStream_t chain = overlap ? st_chain : st;
    cudaStream_t bulk = st;
    if (overlap) {
        CUGP_CUDA(cudaEventRecord(ev[2 * npanels], st));
        CUGP_CUDA(cudaStreamWaitEvent(chain, ev[2 * npanels], 0));
    }
    int p = npanels - 1;
    for (int P1 = n; P1 > 0; p--) {
        const int P0 = ((P1 - 1) / BWD_PANEL) * BWD_PANEL;
        const int last = P0 + ((P1 - P0 - 1) / DB) * DB;
        if (g_bwd_cluster) {
            launch_bwd_chain(L, ld, sL, n, P0, cdiv(P1 - P0, DB), invd, sInvd, work, alpha, sVec, batch, chain);
        } else {
            for (int j0 = last; j0 >= P0; j0 -= DB) {
                const int ctas = j0 > P0 ? cdiv(j0 - P0, BWD_STEP_COLS) : 1;
                trsv_bwd_step_kernel<<<dim3(ctas, batch), TRSV_THREADS, 0, chain>>>(L, ld, sL, n, j0, P0, invd, sInvd, work,
                                                                                   alpha, sVec);
            }
        }
        if (P0 > 0) {
            if (!overlap) {
                bwd_panel_update(L, ld, sL, P0, P1 - P0, 0, P0, work, alpha, sVec, batch, st);
            } else {
                const int Pm = P0 - BWD_PANEL;                      // P0 is a multiple of the panel height
                CUGP_CUDA(cudaEventRecord(ev[2 * p], chain));       // alpha[P0:P1) is final
                if (Pm > 0) {
                    CUGP_CUDA(cudaStreamWaitEvent(bulk, ev[2 * p], 0));
                    bwd_panel_update(L, ld, sL, P0, P1 - P0, 0, Pm, work, alpha, sVec, batch, bulk);   // U2(p)
                    CUGP_CUDA(cudaEventRecord(ev[2 * p + 1], bulk));
                }
                // U1(p) writes columns U2(p+1) also wrote (and U2(p+1) exists because P0 > 0): keep that order
                if (p + 1 < npanels) CUGP_CUDA(cudaStreamWaitEvent(chain, ev[2 * (p + 1) + 1], 0));
                bwd_panel_update(L, ld, sL, P0, P1 - P0, Pm, P0, work, alpha, sVec, batch, chain);     // U1(p)
            }
        }
        P1 = P0;
    }
    if (overlap) {
        CUGP_CUDA(cudaEventRecord(ev[2 * npanels + 1], chain));
        CUGP_CUDA(cudaStreamWaitEvent(st, ev[2 * npanels + 1], 0));
    }
    CUGP_CUDA(cudaGetLastError());
}

void launch_gemv_t(const double* T, int64_t ld, int64_t sT, int n, const double* z, int64_t sZ, double* alpha,
                   int64_t sAlpha, double* scratch, int batch, cudaStream_t st) {
    const int nblk = cdiv(n, DB);
    const int rpc = cdiv(nblk, GT_CHUNKS);            // 128-row sub-blocks per CTA; at most GT_CHUNKS chunks per column block
    const int chunks = cdiv(nblk, rpc);
    const int64_t sPart = (int64_t)GT_CHUNKS * n;     // == trsv_backward_scratch(n, 1)
    gemv_t_kernel<<<dim3(nblk, chunks, batch), TRSV_THREADS, 0, st>>>(T, ld, sT, n, rpc, z, sZ, scratch, sPart);
    gemv_t_reduce_kernel<<<dim3(cdiv(n, 256), batch), 256, 0, st>>>(scratch, sPart, n, rpc, alpha, sAlpha);
    CUGP_CUDA(cudaGetLastError());
}

void launch_init_identity(double* R, int64_t ld, int64_t sR, int n, int batch, cudaStream_t st) {
    if (n <= 0 || batch <= 0) return;
    init_identity_kernel<<<dim3(cdiv(ld / 2, 256), n, batch), 256, 0, st>>>(R, ld, sR, n);
    CUGP_CUDA(cudaGetLastError());
}

void launch_init_idrows(double* R, int64_t sR, double* W, int64_t sW, int64_t ld, int n, int batch, cudaStream_t st) {
    if (n <= 0 || batch <= 0) return;
    init_idrows_kernel<<<dim3(cdiv(ld / 2, 256), n, batch), 256, 0, st>>>(R, sR, W, sW, ld, n);
    CUGP_CUDA(cudaGetLastError());
}

void launch_gemv_upper(const double* U, int64_t ld, int64_t sU, int n, const double* z, int64_t sZ, double* alpha, int64_t sAlpha,
                       int batch, cudaStream_t st) {
    if (n <= 0 || batch <= 0) return;
    gemv_upper_kernel<<<dim3(cdiv(n, 8), batch), 256, 0, st>>>(U, ld, sU, n, z, sZ, alpha, sAlpha);
    CUGP_CUDA(cudaGetLastError());
}

void launch_copy_rows(const double* src, int64_t sSrc, double* dst, int64_t sDst, int n, int batch, cudaStream_t st) {
    if (n <= 0 || batch <= 0) return;
    copy_rows_kernel<<<dim3(cdiv(n, 256), batch), 256, 0, st>>>(src, sSrc, dst, sDst, n);
    CUGP_CUDA(cudaGetLastError());
}

void launch_ll_finalize(const double* y, const double* alpha, int64_t sVec, int n, const double* logdet_part, int nblk,
                        double* scal, int batch, cudaStream_t st) {
    ll_finalize_kernel<<<batch, 256, 0, st>>>(y, alpha, sVec, n, logdet_part, nblk, scal);
    CUGP_CUDA(cudaGetLastError());
}

void launch_copy_vec(const double* src, double* dst, int64_t count, cudaStream_t st) {
    if (count <= 0) return;
    copy_vec_kernel<<<(unsigned)cdiv(count, 256), 256, 0, st>>>(src, dst, count);
    CUGP_CUDA(cudaGetLastError());
}

void launch_scatter_invdiag(const double* invd, int64_t sInvd, double* T, int64_t ld, int64_t sT, int n, int batch,
                            cudaStream_t st) {
    scatter_invdiag_kernel<<<dim3(cdiv(n, DB), batch), 256, 0, st>>>(invd, sInvd, T, ld, sT, n);
    CUGP_CUDA(cudaGetLastError());
}

static void launch_export(const double* A, int64_t ld, int n, double* out, int mode, cudaStream_t st) {
    if (n <= 0) return;
    export_kernel<<<dim3(cdiv(n, 256), n < 32768 ? n : 32768), 256, 0, st>>>(A, ld, n, out, mode);
    CUGP_CUDA(cudaGetLastError());
}
void launch_export_lower(const double* A, int64_t ld, int n, double* out, cudaStream_t st) { launch_export(A, ld, n, out, 0, st); }
void launch_export_symmetric(const double* A, int64_t ld, int n, double* out, cudaStream_t st) { launch_export(A, ld, n, out, 1, st); }
void launch_export_full(const double* A, int64_t ld, int n, double* out, cudaStream_t st) { launch_export(A, ld, n, out, 2, st); }
void launch_import_full(const double* in, int n, double* A, int64_t ld, cudaStream_t st) {
    if (n <= 0) return;
    import_kernel<<<dim3(cdiv(n, 256), n < 32768 ? n : 32768), 256, 0, st>>>(in, n, A, ld);
    CUGP_CUDA(cudaGetLastError());
}

void launch_predict_finalize(const double* meanpart, int ntile_mean, const double* css, int ntile_css, int m, Hyper h,
                             double* mean, double* var, int64_t sOut, int64_t sMp, int64_t sCss, int batch,
                             cudaStream_t st) {
    if (m <= 0) return;
    predict_finalize_kernel<<<dim3(cdiv(m, 256), batch), 256, 0, st>>>(meanpart, ntile_mean, css, ntile_css, m, h, mean,
                                                                        var, sOut, sMp, sCss);
    CUGP_CUDA(cudaGetLastError());
}

void launch_poe_accumulate(const double* mean, const double* var, int64_t sOut, int nexp, int m, double* P, double* Q,
                           int accumulate, cudaStream_t st) {
    if (m <= 0) return;
    poe_accumulate_kernel<<<cdiv(m, 256), 256, 0, st>>>(mean, var, sOut, nexp, m, P, Q, accumulate);
    CUGP_CUDA(cudaGetLastError());
}
void launch_poe_finalize(const double* PQ, int m, double* mean, double* var, cudaStream_t st) {
    if (m <= 0) return;
    poe_finalize_kernel<<<cdiv(m, 256), 256, 0, st>>>(PQ, m, mean, var);
    CUGP_CUDA(cudaGetLastError());
}

}  // namespace cugp
