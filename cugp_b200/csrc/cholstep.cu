// cholstep.cu -- one 128-column block step of the right-looking Cholesky as ONE launch (round 2).
//
// Replaces, for the latency-bound regime (outer width 128: every factorisation below n ~ 5000 and every BCM expert), the
// chain  diag kernel -> TRSM GEMM -> next-block update GEMM  (three dependent launches, ~70 us per 128 columns) of
// get_cholesky (common/matrixops.cpp:68-108).  The CTAs of the launch take ROLES in the order they start (an atomic
// ticket per matrix, so a CTA only ever waits for CTAs that are already running):
//
//   SYRKD (10 CTAs, only with `prologue`): the 32x32 lower blocks of this step's 128x128 diagonal block receive the
//          previous block column's contribution  A_jj -= L[j, j-1] L[j, j-1]^T  (K = 128), then signal a counter.
//   DIAG  (1 CTA): waits for that counter, factors the 128x128 block in shared memory -- four 32-column panels; one warp
//          factors the 32x32 diagonal sub-block AND inverts it in the same register-resident column loop (row i of L and
//          column i of inv(L) share one 32-entry array), the other 15 warps apply the previous panel's TRSM / SYRK
//          remainder on DMMA meanwhile (look-ahead inside the CTA) -- writes L11 and the four 32x32 diagonal inverses,
//          then releases a flag.
//   ROWS  (one CTA per 32 rows below the block, the appended y^T row included): while DIAG works they apply the
//          previous block column's contribution to their 32x128 tile in shared memory (the `prologue`, K = 128 on DMMA;
//          the tile never returns to HBM in between), then wait for the flag and run the TRSM as a 4-step blocked
//          substitution with the 32x32 inverses:  X_c = (C_c - sum_{k<c} X_k L_ck^T) inv(L_cc)^T.
//
// The pivot chain of the 32x32 factorisation is kept free of the shared-memory broadcast round trip (round 1's kernel
// fed the SHFL of the next pivot from a register that also waited on the LDS of the column broadcast: 238 clocks per
// column, see profiles/r2_diag_chain.txt).
#include "cholstep.cuh"

#include <algorithm>

namespace cugp {

namespace {

constexpr int DB = kDiag;           // 128
constexpr int SBW = 32;             // panel width inside the diagonal block / row-tile height
constexpr int LDS_ = DB + 4;        // shared row stride: (132 % 16 == 4) -> conflict-free 8-byte DMMA fragments both ways
constexpr int NT = 512;
constexpr int NWARP = NT / 32;
constexpr int NSYRKD = 10;          // 32x32 lower blocks of a 128x128 block

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// rows x 128 tile of a row-major matrix into shared memory (stride LDS_), zero filled outside [0, rows_valid) x
// [0, cols_valid).  16-byte cp.async.cg: L2-coherent, so data another CTA of this launch released is seen.
__device__ __forceinline__ void load_tile(double* dst, const double* src, int64_t ld, int rows, int rows_valid, int cols_valid,
                                          int tid, const double* safe) {
    for (int e = tid; e < rows * (DB / 2); e += NT) {
        const int r = e >> 6, c = (e & 63) * 2;
        int bytes = (cols_valid - c) * 8;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        if (r >= rows_valid) bytes = 0;
        cp_async16(dst + r * LDS_ + c, bytes ? src + (int64_t)r * ld + c : safe, bytes);   // 0 bytes: nothing is read
    }
}

// One warp: acc[MI][NI] += A(i, k) B(k, j) over k in [klo, khi) (multiples of 4); i, j relative to the warp tile.
template <int MI, int NI, class FA, class FB>
__device__ __forceinline__ void warp_mma(double (&acc)[MI][NI][2], int klo, int khi, FA A, FB B, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll 4
    for (int k = klo; k < khi; k += 4) {
        double af[MI], bf[NI];
#pragma unroll
        for (int i = 0; i < MI; i++) af[i] = A(i * 8 + g, k + q);
#pragma unroll
        for (int j = 0; j < NI; j++) bf[j] = B(k + q, j * 8 + g);
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NI; j++) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
}

// 8x8 output blocks x = first, first + stride, ... < count, each a K-deep product on two accumulator chains; the
// `blk` functor maps the linear index to (row0, col0, klo, khi); `out(row, col0, col1, v0, v1)` consumes the result.
template <class FBLK, class FA, class FB, class FO>
__device__ __forceinline__ void blocks_mma(int first, int stride, int count, FBLK blk, FA A, FB B, FO out, int lane) {
    const int g = lane >> 2, q = lane & 3;
    for (int x = first; x < count; x += stride) {
        int r0, c0, klo, khi;
        blk(x, r0, c0, klo, khi);
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
        for (int k = klo; k < khi; k += 8) {
            dmma(a0, a1, A(r0 + g, k + q), B(k + q, c0 + g));
            if (k + 4 < khi) dmma(b0, b1, A(r0 + g, k + 4 + q), B(k + 4 + q, c0 + g));
        }
        out(r0 + g, c0 + 2 * q, a0 + b0, a1 + b1);
    }
}

// ------------------------------------------------------------------------------------------------
// 32x32 Cholesky + inverse by ONE warp (matrixops.cpp:74-98 and :330-340 on the sub-block).
// Lane i owns row i of A -> L (entries c < i) and column i of T = inv(L) (entries c > i) in ONE array u[32]; the
// diagonal entries live in scalars.  Column j:
//     rd = rsqrt(a_jj),  L(i,j) = a_ij rd (i > j),  T(j,i) = r_j rd (i < j),  d = a_jj rd = L(j,j),  T(j,j) = rd
//     cb[] <- column j of L (shared-memory broadcast)
//     u[c] -= cb[c] * m   for c > j,  m = L(i,j) for lanes below the pivot, T(j,i) for lanes up to it
//     a_ii -= L(i,j)^2    kept in a scalar: the NEXT pivot leaves through SHFL without waiting for the broadcast.
// Results go to S: row (c0+i): columns c0..c0+i = L(i, :), columns c0+i+1..c0+32 = T(:, i) (T(r,i) at column c0+r+1),
// the layout every consumer below reads (T(r, c) = S[c][r + 1]).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void chol32_inv(double* S, int c0, double* cb, int lane) {
    double u[SBW];
    double* row = S + (c0 + lane) * LDS_ + c0;
#pragma unroll
    for (int c = 0; c < SBW; c++) u[c] = (c < lane) ? row[c] : 0.0;
    double adiag = row[lane];
    double ldiag = 0.0, tdiag = 0.0;
    double ajj = __shfl_sync(0xffffffffu, adiag, 0);
#pragma unroll
    for (int j = 0; j < SBW; j++) {
        const double rd = rsqrt(ajj);          // negative pivot -> NaN, propagates (matrixops.cpp:77)
        const double d = ajj * rd;             // sqrt(a_jj) without the sqrt -> divide chain
        const double val = u[j] * rd;
        const double l = (lane > j) ? val : 0.0;                    // L(lane, j)
        const double m = (lane == j) ? rd : val;                    // multiplier of the column update
        if (lane == j) {
            ldiag = d;
            tdiag = rd;
        }
        u[j] = val;
        if (j + 1 < SBW) {
            adiag = fma(-l, l, adiag);                               // a_ii -= L(i,j)^2
            ajj = __shfl_sync(0xffffffffu, adiag, j + 1);            // next pivot: on its way before the broadcast
            double* buf = cb + (j & 1) * SBW;
            buf[lane] = (lane > j) ? val : 0.0;
            __syncwarp();
            if (lane == j) {
#pragma unroll
                for (int c = j + 1; c < SBW; c++) u[c] = 0.0;        // column `lane` of T starts from e_lane
            }
#pragma unroll
            for (int c = j + 1; c < SBW; c++) u[c] = fma(-buf[c], m, u[c]);
        }
    }
#pragma unroll
    for (int p = 0; p <= SBW; p++) {
        double v;
        if (p < lane) v = u[p < SBW ? p : 0];
        else if (p == lane) v = ldiag;
        else if (p == lane + 1) v = tdiag;
        else v = u[p > 0 ? p - 1 : 0];
        row[p] = v;
    }
}

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
}
// optional phase stamps (tuning aid, tools/r2_step_phases.py): stamps[blk][role 0..2][16], globaltimer ns, batch 0 only
#define STEP_STAMP(role, idx)                                                                      \
    do {                                                                                           \
        if (p.stamps && tid == 0 && blockIdx.y == 0) p.stamps[(p.blk * 3 + (role)) * 16 + (idx)] = gtime(); \
    } while (0)

struct StepArgs {
    double* A;
    int64_t ld, sA;
    int n, nrows, j0;
    double* invd;          // [batch][nblk][128][128]
    int64_t sInvd;
    double* logdet_part;   // [batch][nblk]
    int nblk, blk;
    int* sync;             // [batch][nblk][4]: ticket, SYRKD counter, DIAG flag
    int prologue;
    int nrow_tiles;
    long long* stamps;
};

// Shared memory: Lb [128][132] | As [32][132] | Cs [32][132] | cb [2][32] | red [128] | role
constexpr size_t STEP_SMEM = (size_t)(DB * LDS_ + 2 * SBW * LDS_ + 2 * SBW + DB) * sizeof(double) + 16;

__global__ void __launch_bounds__(NT, 1) chol_step_kernel(const StepArgs p) {
    extern __shared__ __align__(16) double sm[];
    double* Lb = sm;
    double* As = Lb + DB * LDS_;
    double* Cs = As + SBW * LDS_;
    double* cb = Cs + SBW * LDS_;
    double* red = cb + 2 * SBW;
    int* role_s = reinterpret_cast<int*>(red + DB);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int64_t b = blockIdx.y;
    double* A = p.A + b * p.sA;
    int* sync = p.sync + (b * p.nblk + p.blk) * 4;
    const int j0 = p.j0, n = p.n;
    const int nb = min(DB, n - j0);          // columns of this block (< 128 only for the last one)
    if (tid == 0) role_s[0] = atomicAdd(&sync[0], 1);
    __syncthreads();
    const int ticket = role_s[0];
    const int nsy = p.prologue ? NSYRKD : 0;

    if (ticket < nsy) {
        // ------------------------------------------------------------------ SYRKD
        int bi = 0, bj = ticket;              // ticket -> (bi, bj), bi >= bj, row-major over the lower 4x4 block set
        while (bj > bi) {
            bj -= bi + 1;
            bi++;
        }
        const int ri = j0 + SBW * bi, rj = j0 + SBW * bj;
        const int pj = j0 - DB;               // previous block column
        if (ticket == 0) STEP_STAMP(0, 0);
        load_tile(As, A + (int64_t)ri * p.ld + pj, p.ld, SBW, n - ri, DB, tid, A);
        load_tile(Cs, A + (int64_t)rj * p.ld + pj, p.ld, SBW, n - rj, DB, tid, A);
        cp_async_wait_all();
        __syncthreads();
        if (ticket == 0) STEP_STAMP(0, 1);
        // warp w: 8x8 block (w / 4, w % 4) of the 32x32 output, K = 128
        const int r8 = (warp >> 2) * 8, c8 = (warp & 3) * 8;
        double acc[1][1][2] = {{{0.0, 0.0}}};
        double acc2[1][1][2] = {{{0.0, 0.0}}};
        warp_mma<1, 1>(acc, 0, DB / 2, [&](int i, int k) { return As[(r8 + i) * LDS_ + k]; },
                       [&](int k, int j) { return Cs[(c8 + j) * LDS_ + k]; }, lane);
        warp_mma<1, 1>(acc2, DB / 2, DB, [&](int i, int k) { return As[(r8 + i) * LDS_ + k]; },
                       [&](int k, int j) { return Cs[(c8 + j) * LDS_ + k]; }, lane);
        const int gi = ri + r8 + g, gj = rj + c8 + 2 * q;
        if (gi < n) {
            double* dst = A + (int64_t)gi * p.ld + gj;
            if (gj <= gi && gj < n) dst[0] -= acc[0][0][0] + acc2[0][0][0];
            if (gj + 1 <= gi && gj + 1 < n) dst[1] -= acc[0][0][1] + acc2[0][0][1];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) red_release_add(&sync[1], 1);
        if (ticket == 0) STEP_STAMP(0, 2);
        return;
    }

    if (ticket == nsy) {
        // ------------------------------------------------------------------ DIAG
        double* S = Lb;
        STEP_STAMP(1, 0);
        if (nsy) {
            if (tid == 0)
                while (ld_acquire(&sync[1]) < NSYRKD) __nanosleep(40);
            __syncthreads();
        }
        STEP_STAMP(1, 1);
        load_tile(S, A + (int64_t)j0 * p.ld + j0, p.ld, DB, nb, nb, tid, A);
        cp_async_wait_all();
        __syncthreads();
        STEP_STAMP(1, 2);
        if (tid < DB && tid >= nb) S[tid * LDS_ + tid] = 1.0;   // identity padding of the last block
        __syncthreads();

        for (int c = 0; c < DB / SBW; c++) {
            const int c0 = c * SBW;
            if (warp == 0) {
                chol32_inv(S, c0, cb, lane);
            } else if (c > 0) {
                // remainder of panel c-1 (rows below the next diagonal sub-block), warps 1..15:
                //   X = A[R, p0:p0+32] T_pp^T  for R = [c0 + 32, 128), then  A[R, c0:] -= X X[c0:, :]^T (lower part)
                const int p0 = c0 - SBW, R0 = c0 + SBW, nr = DB - R0;
                if (nr > 0) {
                    const int w = warp - 1;
                    // TRSM blocks: (nr/8) x 4, results to registers first (in place: every block reads the whole strip)
                    double v0[3], v1[3];
                    int rr[3], cc[3], cnt = 0;
                    blocks_mma(
                        w, NWARP - 1, (nr / 8) * 4,
                        [&](int x, int& r0, int& cc0, int& klo, int& khi) {
                            r0 = R0 + (x >> 2) * 8;
                            cc0 = (x & 3) * 8;
                            klo = 0;
                            khi = cc0 + 8;   // T_pp lower triangular: k <= j
                        },
                        [&](int i, int k) { return S[i * LDS_ + p0 + k]; },
                        [&](int k, int j) { return k <= j ? S[(p0 + k) * LDS_ + p0 + j + 1] : 0.0; },   // T(j, k)
                        [&](int i, int j, double a, double bb) {
                            if (cnt < 3) { rr[cnt] = i; cc[cnt] = j; v0[cnt] = a; v1[cnt] = bb; }
                            cnt++;
                        },
                        lane);
                    named_bar(1, NT - 32);
                    for (int t = 0; t < cnt && t < 3; t++) {
                        S[rr[t] * LDS_ + p0 + cc[t]] = v0[t];
                        S[rr[t] * LDS_ + p0 + cc[t] + 1] = v1[t];
                    }
                    named_bar(1, NT - 32);
                    // SYRK blocks: rows R (8-row blocks rb), columns c0 .. row block's diagonal: lower 8x8 blocks
                    const int nrb = nr / 8, cb0 = SBW / 8;   // column blocks left of R0: cb0 (the c0..R0 strip), then rb + 1
                    const int total = nrb * cb0 + nrb * (nrb + 1) / 2;
                    blocks_mma(
                        w, NWARP - 1, total,
                        [&](int x, int& r0, int& cc0, int& klo, int& khi) {
                            klo = 0;
                            khi = SBW;
                            if (x < nrb * cb0) {
                                r0 = R0 + (x / cb0) * 8;
                                cc0 = c0 + (x % cb0) * 8;
                            } else {
                                int y = x - nrb * cb0, rb = 0;
                                while (y > rb) {
                                    y -= rb + 1;
                                    rb++;
                                }
                                r0 = R0 + rb * 8;
                                cc0 = R0 + y * 8;
                            }
                        },
                        [&](int i, int k) { return S[i * LDS_ + p0 + k]; },
                        [&](int k, int j) { return S[j * LDS_ + p0 + k]; },
                        [&](int i, int j, double a, double bb) {
                            S[i * LDS_ + j] -= a;
                            S[i * LDS_ + j + 1] -= bb;
                        },
                        lane);
                }
            }
            __syncthreads();
            STEP_STAMP(1, 3 + 2 * c);
            if (c0 + SBW >= DB) break;
            // critical part of panel c: the next 32 rows.  X = A[c0+32 : c0+64, c0 : c0+32] T_cc^T, one 8x8 block per warp
            {
                const int R0 = c0 + SBW;
                const int r8 = R0 + (warp >> 2) * 8, c8 = (warp & 3) * 8;
                double a0 = 0.0, a1 = 0.0;
                for (int k = 0; k < c8 + 8; k += 4)   // T_cc(j, k) = S[c0 + k][c0 + j + 1] for k <= j
                    dmma(a0, a1, S[(r8 + g) * LDS_ + c0 + k + q], k + q <= c8 + g ? S[(c0 + k + q) * LDS_ + c0 + c8 + g + 1] : 0.0);
                __syncthreads();
                S[(r8 + g) * LDS_ + c0 + c8 + 2 * q] = a0;
                S[(r8 + g) * LDS_ + c0 + c8 + 2 * q + 1] = a1;
                __syncthreads();
                // next diagonal sub-block: A[R0:R0+32, R0:R0+32] -= X X^T, 10 lower 8x8 blocks
                if (warp < 10) {
                    int bi = 0, bj = warp;
                    while (bj > bi) {
                        bj -= bi + 1;
                        bi++;
                    }
                    double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
                    for (int k = 0; k < SBW; k += 8) {
                        dmma(s0, s1, S[(R0 + bi * 8 + g) * LDS_ + c0 + k + q], S[(R0 + bj * 8 + g) * LDS_ + c0 + k + q]);
                        dmma(t0, t1, S[(R0 + bi * 8 + g) * LDS_ + c0 + k + 4 + q], S[(R0 + bj * 8 + g) * LDS_ + c0 + k + 4 + q]);
                    }
                    S[(R0 + bi * 8 + g) * LDS_ + R0 + bj * 8 + 2 * q] -= s0 + t0;
                    S[(R0 + bi * 8 + g) * LDS_ + R0 + bj * 8 + 2 * q + 1] -= s1 + t1;
                }
                __syncthreads();
                STEP_STAMP(1, 4 + 2 * c);
            }
        }

        // write back: L11 (lower) into A, the four 32x32 diagonal inverses into invd; then release the flag
        double* inv = p.invd + b * p.sInvd + (int64_t)p.blk * DB * DB;
        for (int e = tid; e < DB * DB; e += NT) {
            const int i = e >> 7, j = e & (DB - 1);
            if (j <= i) {
                if (i < nb) A[(int64_t)(j0 + i) * p.ld + j0 + j] = S[i * LDS_ + j];
                if ((i >> 5) == (j >> 5)) inv[i * DB + j] = S[j * LDS_ + i + 1];
            }
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&sync[2], 1);
        STEP_STAMP(1, 11);
        if (p.logdet_part) {
            if (tid < DB) red[tid] = log(S[tid * LDS_ + tid]);
            __syncthreads();
            for (int off = DB / 2; off > 0; off >>= 1) {
                if (tid < off) red[tid] += red[tid + off];
                __syncthreads();
            }
            if (tid == 0) p.logdet_part[b * p.nblk + p.blk] = red[0];
        }
        STEP_STAMP(1, 12);
        return;
    }

    // ---------------------------------------------------------------------- ROWS
    const int tile = ticket - nsy - 1;
    if (tile >= p.nrow_tiles) return;
    const int r0 = j0 + nb + tile * SBW;          // first row of the tile (rows below the diagonal block)
    const int rows_valid = p.nrows - r0;
    const bool st0 = tile == 0;
    if (st0) STEP_STAMP(2, 0);
    load_tile(Cs, A + (int64_t)r0 * p.ld + j0, p.ld, SBW, rows_valid, nb, tid, A);
    if (p.prologue) {
        const int pj = j0 - DB;
        load_tile(As, A + (int64_t)r0 * p.ld + pj, p.ld, SBW, rows_valid, DB, tid, A);
        load_tile(Lb, A + (int64_t)j0 * p.ld + pj, p.ld, DB, nb, DB, tid, A);
        cp_async_wait_all();
        __syncthreads();
        if (st0) STEP_STAMP(2, 1);
        // C[32 x 128] -= L[R, prev] L[block rows, prev]^T : warp tile 16 x 16, K = 128
        const int wr = (warp >> 3) * 16, wc = (warp & 7) * 16;
        double acc[2][2][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
        warp_mma<2, 2>(acc, 0, DB, [&](int i, int k) { return As[(wr + i) * LDS_ + k]; },
                       [&](int k, int j) { return Lb[(wc + j) * LDS_ + k]; }, lane);
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double* dst = Cs + (wr + i * 8 + g) * LDS_ + wc + j * 8 + 2 * q;
                dst[0] -= acc[i][j][0];
                dst[1] -= acc[i][j][1];
            }
    } else {
        cp_async_wait_all();
    }
    __syncthreads();   // Lb is free again, Cs holds the updated tile
    if (st0) STEP_STAMP(2, 2);
    if (tid == 0)
        while (ld_acquire(&sync[2]) == 0) __nanosleep(40);
    __syncthreads();
    if (st0) STEP_STAMP(2, 3);
    load_tile(Lb, A + (int64_t)j0 * p.ld + j0, p.ld, DB, nb, nb, tid, A);
    cp_async_wait_all();
    __syncthreads();
    {
        // T_cc blocks into the transposed-upper positions: T(i, j) at Lb[j][i + 1] (same layout as DIAG's S)
        const double* inv = p.invd + b * p.sInvd + (int64_t)p.blk * DB * DB;
        for (int e = tid; e < 4 * SBW * SBW; e += NT) {
            const int c = e >> 10, i = (e >> 5) & 31, j = e & 31;
            if (j <= i) Lb[(c * SBW + j) * LDS_ + c * SBW + i + 1] = __ldcg(inv + (c * SBW + i) * DB + c * SBW + j);
        }
    }
    __syncthreads();
    if (st0) STEP_STAMP(2, 4);
    // blocked substitution, one 8x8 block of the 32x32 step per warp
    const int r8 = (warp >> 2) * 8, c8 = (warp & 3) * 8;
    for (int c = 0; c < DB / SBW; c++) {
        const int c0 = c * SBW;
        if (c0 >= nb) break;
        if (c > 0) {
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            for (int k = 0; k < c0; k += 8) {
                dmma(a0, a1, Cs[(r8 + g) * LDS_ + k + q], Lb[(c0 + c8 + g) * LDS_ + k + q]);
                dmma(b0, b1, Cs[(r8 + g) * LDS_ + k + 4 + q], Lb[(c0 + c8 + g) * LDS_ + k + 4 + q]);
            }
            double* dst = Cs + (r8 + g) * LDS_ + c0 + c8 + 2 * q;   // this warp's own block: no other reader yet
            dst[0] -= a0 + b0;
            dst[1] -= a1 + b1;
            __syncthreads();
        }
        double x0 = 0.0, x1 = 0.0;
        for (int k = 0; k < c8 + 8; k += 4)
            dmma(x0, x1, Cs[(r8 + g) * LDS_ + c0 + k + q], k + q <= c8 + g ? Lb[(c0 + k + q) * LDS_ + c0 + c8 + g + 1] : 0.0);
        __syncthreads();
        Cs[(r8 + g) * LDS_ + c0 + c8 + 2 * q] = x0;
        Cs[(r8 + g) * LDS_ + c0 + c8 + 2 * q + 1] = x1;
        __syncthreads();
    }
    if (st0) STEP_STAMP(2, 5);
    // write X back (rows < nrows, columns < nb)
    for (int e = tid; e < SBW * (DB / 2); e += NT) {
        const int r = e >> 6, c = (e & 63) * 2;
        if (r < rows_valid) {
            double* dst = A + (int64_t)(r0 + r) * p.ld + j0 + c;
            if (c + 1 < nb) *reinterpret_cast<double2*>(dst) = make_double2(Cs[r * LDS_ + c], Cs[r * LDS_ + c + 1]);
            else if (c < nb) dst[0] = Cs[r * LDS_ + c];
        }
    }
    if (st0) STEP_STAMP(2, 6);
}

static long long* g_step_stamps = nullptr;

}  // namespace

void set_step_stamps(long long* dev) { g_step_stamps = dev; }

void launch_chol_step(double* A, int64_t ld, int64_t sA, int n, int nrows, int j0, double* invd, int64_t sInvd,
                      double* logdet_part, int nblk, int* sync, int prologue, int batch, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        CUGP_CUDA(cudaFuncSetAttribute(chol_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM));
        configured = true;
    }
    StepArgs p{};
    p.A = A; p.ld = ld; p.sA = sA; p.n = n; p.nrows = nrows; p.j0 = j0;
    p.invd = invd; p.sInvd = sInvd; p.logdet_part = logdet_part; p.nblk = nblk; p.blk = j0 / DB;
    p.sync = sync; p.prologue = prologue; p.stamps = g_step_stamps;
    const int nb = std::min(DB, n - j0);
    const int below = nrows - (j0 + nb);
    p.nrow_tiles = below > 0 ? cdiv(below, SBW) : 0;
    const int ctas = (prologue ? NSYRKD : 0) + 1 + p.nrow_tiles;
    chol_step_kernel<<<dim3((unsigned)ctas, (unsigned)batch), NT, STEP_SMEM, st>>>(p);
    CUGP_CUDA(cudaGetLastError());
}

}  // namespace cugp
