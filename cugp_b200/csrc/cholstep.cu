// cholstep.cu -- one 128-column block step of the right-looking Cholesky as ONE launch (round 2).
//
// Replaces, wherever a factorisation is bound by its chain of block steps (outer width 128: everything below n ~ 5000 and
// every BCM expert; with `prologue` = 0 also the 128-column blocks inside wider outer panels), the chain
//     diag kernel (40 us) -> TRSM GEMM -> next-block update GEMM        (three dependent launches, ~70 us per 128 columns)
// of get_cholesky (common/matrixops.cpp:68-108).  The CTAs of the launch take ROLES in the order they start (an atomic
// ticket per matrix, so a CTA only ever waits for CTAs that are already running -- no co-residency assumption beyond
// "a started CTA makes progress"):
//
//   SYRKD (10 CTAs, only with `prologue`): the 32x32 lower blocks of this step's 128x128 diagonal block receive the
//          previous block column's contribution  A_jj -= L[j, j-1] L[j, j-1]^T  (K = 128, DMMA), then bump a counter.
//   DIAG  (1 CTA, 16 warps): waits for that counter, loads the block, and factors it slab by slab (4 slabs of 32 columns):
//            pivot warp      8-column panels in registers (SHFL broadcasts, branch-free rsqrt), rank-8 DMMA updates
//            follower warps  the rows below the 32x32 sub-block follow panel by panel (named barriers: the pivot warp
//                            only ever ARRIVES), so the in-block TRSM needs no inverse and no extra phase
//            all warps       the slab's SYRK on the next slab's columns (critical), 12 helper warps the rest of it, the
//                            32x32 inverse T_cc of the finished sub-block and the PUBLICATION of the finished column slab
//                            (L(:, slab) into the matrix, L and T_cc into a 128x132 image in global memory) with a
//                            release flag -- all in the shadow of the next slab's pivot chain
//   ROWS  (one CTA per 32 rows below the block; appended right-hand-side / identity rows included): while DIAG works
//          they apply the previous block column's contribution to their 32x128 tile in shared memory (the `prologue`,
//          K = 128 on DMMA; the tile never returns to HBM in between), then follow DIAG slab by slab (acquire flag c,
//          fetch slab c, X_c = C_c T_cc^T, C_c' -= X_c L(c', c)^T for c' > c): when DIAG finishes its last slab only one
//          32x32x32 product is left.
//
// Measured (profiles/r2_step_phases_*.txt, n = 1500, one B200): 70 us -> 35 us per block step; what is left is the pivot
// chain (4 x 3.7 us), the SYRKD hand-over (3.6 us), the tile loads and the launch gap.
// The pivot chain itself: tools/chol32_microbench.cu (profiles/r2_chol32_microbench.txt) -- a whole 32x32 factorisation
// unrolled in one warp is bound by instruction issue and by how ptxas orders the 31-j column updates around the rsqrt
// (240-390 clocks per column); the 8-column panel form keeps <= 7 updates per column next to the chain.
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
// The 32-column slab of the diagonal block (columns c0..c0+31, all rows from c0 down to 127), matrixops.cpp:74-98 on it.
// Measured (tools/chol32_microbench.cu, profiles/r2_chol32_microbench.txt): a whole 32x32 factorisation unrolled in one
// warp is bound by instruction issue, not by its pivot chain (chain alone 120 clocks per column, with the 31-j column
// updates 200-390 depending on how ptxas orders them).  So the pivot warp keeps only an 8-column panel in registers
// (<= 7 updates per column), the rest of the slab follows on DMMA:
//   pivot warp (rows c0..c0+31): per panel p -- 8 pivot columns (SHFL broadcasts, branch-free rsqrt), store, signal
//                                barrier 1+p, then the rank-8 update of its own rows' remaining slab columns (DMMA)
//   follower warps w = 1.. (rows c0+32w..): wait for barrier 1+p, solve their rows against the panel (8 columns),
//                                store, rank-8 update of their rows (DMMA).  They never hold the pivot warp up.
// 1/L(k,k) is left in rdg[] for the followers and for the inverse of the diagonal sub-block.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double rsqrt_nb(double x) {
    // MUFU.RSQ64H seed + the one cubic Newton step of CUDA's rsqrt(), without its special-case branch (which pins the
    // whole column update between the Newton step and its use).  x <= 0 -> NaN / Inf, which propagates (matrixops.cpp:77).
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = y * y;
    const double e = fma(-t, x, 1.0);
    const double p2 = fma(e, 0.375, 0.5);
    const double s = y * e;
    return fma(p2, s, y);
}
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// rank-8 update of rows [r0, r0 + 8 nrb) x slab columns [pc + 8, c0 + 32) by one warp:  S[R][C] -= sum_k S[R][pc+k] S[C][pc+k].
// `tri`: the rows are the diagonal block's own (row block rb only needs column blocks up to its own).
__device__ __forceinline__ void slab_rank8(double* S, int r0, int nrb, int c0, int p, bool tri, int lane) {
    const int g = lane >> 2, q = lane & 3;
    const int pc = c0 + 8 * p;
    for (int cbk = p + 1; cbk < 4; cbk++) {
        const int cc = c0 + 8 * cbk;
        const double b0 = S[(cc + g) * LDS_ + pc + q], b1 = S[(cc + g) * LDS_ + pc + 4 + q];
        for (int rb = 0; rb < nrb; rb++) {
            const int rr = r0 + 8 * rb;
            if (tri && rr < cc) continue;
            double x0 = 0.0, x1 = 0.0;
            dmma(x0, x1, S[(rr + g) * LDS_ + pc + q], b0);
            dmma(x0, x1, S[(rr + g) * LDS_ + pc + 4 + q], b1);
            double* dst = S + (rr + g) * LDS_ + cc + 2 * q;
            dst[0] -= x0;
            dst[1] -= x1;
        }
    }
}

// pivot warp: rows c0 + lane.  nfollow: follower warps to signal (0: none).
__device__ __forceinline__ void slab_pivot(double* S, double* rdg, int c0, int nfollow, int lane) {
    double* row = S + (c0 + lane) * LDS_ + c0;
#pragma unroll 1
    for (int p = 0; p < 4; p++) {
        const int pl0 = 8 * p;
        double a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = (lane >= pl0) ? row[pl0 + c] : 0.0;
        double ajj = __shfl_sync(0xffffffffu, a[0], pl0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int pl = pl0 + j;                       // pivot lane
            const double rd = rsqrt_nb(ajj);
            const double d = ajj * rd;                    // sqrt(a_jj) without the sqrt -> divide chain
            const double la = (lane > pl) ? a[j] * rd : (lane == pl ? d : 0.0);
            a[j] = la;
            if (lane == pl) rdg[c0 + pl] = rd;
            if (j + 1 < 8) {
                const double pvn = fma(-la, la, a[j + 1]);               // next pivot, valid in lane pl + 1
                ajj = __shfl_sync(0xffffffffu, pvn, pl + 1);
#pragma unroll
                for (int c = j + 1; c < 8; c++) {
                    const double lc = __shfl_sync(0xffffffffu, la, pl0 + c);   // L(pl0 + c, pl)
                    a[c] = fma(-la, lc, a[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 8; c++)
            if (lane >= pl0 + c) row[pl0 + c] = a[c];      // lower part only: the upper part of S holds the inverses
        __syncwarp();
        if (nfollow > 0) named_arrive(1 + p, 32 * (1 + nfollow));
        if (p < 3) {
            slab_rank8(S, c0 + 8 * (p + 1), 3 - p, c0, p, true, lane);
            __syncwarp();
        }
    }
}

// follower warp w >= 1: rows c0 + 32 w + lane.
__device__ __forceinline__ void slab_follow(double* S, const double* rdg, int c0, int w, int nfollow, int lane) {
    double* row = S + (c0 + 32 * w + lane) * LDS_ + c0;
#pragma unroll 1
    for (int p = 0; p < 4; p++) {
        const int pl0 = 8 * p;
        double b[8];
#pragma unroll
        for (int c = 0; c < 8; c++) b[c] = row[pl0 + c];
        named_bar(1 + p, 32 * (1 + nfollow));             // the pivot warp has stored panel p and its 1/L(k,k)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double lb = b[j] * rdg[c0 + pl0 + j];
            b[j] = lb;
#pragma unroll
            for (int c = j + 1; c < 8; c++) b[c] = fma(-lb, S[(c0 + pl0 + c) * LDS_ + c0 + pl0 + j], b[c]);
        }
#pragma unroll
        for (int c = 0; c < 8; c++) row[pl0 + c] = b[c];
        __syncwarp();
        if (p < 3) {
            slab_rank8(S, c0 + 32 * w, 4, c0, p, false, lane);
            __syncwarp();
        }
    }
}

// inverse of the 32x32 diagonal sub-block at c0 by one warp: lane c solves L x = e_c (matrixops.cpp:330-340);
// T(k, c) goes to S[c0 + c][c0 + k + 1] (the layout every consumer reads: T(r, c) = S[c][r + 1]).
__device__ __forceinline__ void inv32_warp(double* S, const double* rdg, int c0, int lane) {
    double r[SBW];
#pragma unroll
    for (int i = 0; i < SBW; i++) r[i] = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < SBW; k++) {
        const double x = r[k] * rdg[c0 + k];
        r[k] = x;
#pragma unroll
        for (int i = k + 1; i < SBW; i++) r[i] = fma(-S[(c0 + i) * LDS_ + c0 + k], x, r[i]);
    }
    double* row = S + (c0 + lane) * LDS_ + c0;
#pragma unroll
    for (int k = 0; k < SBW; k++)
        if (k >= lane) row[k + 1] = r[k];
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
    double* pub;           // [batch][128][132]: image of DIAG's shared tile (L11 lower, T_cc at [c][r + 1]), column slab by slab
    double* logdet_part;   // [batch][nblk]
    int nblk, blk;
    int* sync;             // [batch][nblk][4]: ticket, SYRKD counter, DIAG column slabs published (0..4)
    int prologue;
    int nrow_tiles;
    int rt;                // rows per ROWS tile: 32 or 64
    int roles;             // 0: every role in this launch; 1: SYRKD + DIAG only; 2: ROWS only (DIAG's launch is complete)
    long long* stamps;
};

// Shared memory: Lb [128][132] | As [32][132] | Cs [32][132] | cb [2][32] | red [128] | role
constexpr size_t STEP_SMEM = (size_t)(DB * LDS_ + 2 * SBW * LDS_ + 2 * SBW + DB) * sizeof(double) + 16;

// DIAG: column slab `sl` of the shared tile is final -- L(i, 32 sl .. 32 sl + 31) for every row i >= 32 sl, and T_sl in the
// shifted upper positions of the diagonal sub-block's rows (columns up to 32 sl + 32).  Copy it (34 columns: 17 chunks
// of 16 bytes per row) to the published image and its L part into the matrix.  Threads t = first, first + count, ...
constexpr int SLAB_CH = SBW / 2 + 1;
// Rows [32 sl + row_lo, 32 sl + row_hi) of the slab (row_hi = 0: down to row 127).
__device__ __forceinline__ void publish_slab(const double* S, double* pub, double* A, int64_t ld, int j0, int nb, int sl,
                                             int t, int count, int row_lo = 0, int row_hi = 0) {
    const int s0 = SBW * sl, r_lo = s0 + row_lo, rows = (row_hi ? s0 + row_hi : DB) - r_lo;
    for (int e = t; e < rows * SLAB_CH; e += count) {
        const int i = r_lo + e / SLAB_CH, ch = e % SLAB_CH, c = s0 + 2 * ch;
        const double2 v = *reinterpret_cast<const double2*>(S + i * LDS_ + c);
        *reinterpret_cast<double2*>(pub + i * LDS_ + c) = v;
        if (i < nb && ch < SBW / 2) {
            double* dst = A + (int64_t)(j0 + i) * ld + j0 + c;
            if (c + 1 <= i) *reinterpret_cast<double2*>(dst) = v;
            else if (c == i) dst[0] = v.x;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ROWS with 64-row tiles (StepArgs::rt == 64): half as many CTAs sit on SMs waiting for DIAG's flags, so the trailing
// updates of the previous step (other stream) keep more of the chip.  The 64x128 tile takes the As+Cs region; the
// prologue's operands go through the (still unused) Lb region in two K-halves of 64 (stride 68: 68 % 16 == 4).
// ------------------------------------------------------------------------------------------------
constexpr int RT64 = 64;
constexpr int LDH_ = DB / 2 + 4;    // 68
__device__ __forceinline__ void load_half(double* dst, const double* src, int64_t ld, int rows, int rows_valid, int tid,
                                          const double* safe) {
    for (int e = tid; e < rows * (DB / 4); e += NT) {
        const int r = e >> 5, c = (e & 31) * 2;
        const int bytes = r < rows_valid ? 16 : 0;
        cp_async16(dst + r * LDH_ + c, bytes ? src + (int64_t)r * ld + c : safe, bytes);
    }
}

template <class STAMP>
__device__ __forceinline__ void rows64_role(const StepArgs& p, double* A, const double* pub, int* sync, double* Lb, double* Ct,
                                            int tile, int j0, int nb, int tid, STAMP stamp) {
    const int lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int r0 = j0 + nb + tile * RT64;
    const int rows_valid = p.nrows - r0;
    load_tile(Ct, A + (int64_t)r0 * p.ld + j0, p.ld, RT64, rows_valid, nb, tid, A);
    if (p.prologue) {
        double* Pa = Lb;                    // [64][68]  L[R, prev half]
        double* Pb = Lb + RT64 * LDH_;      // [128][68] L[block rows, prev half]
        const int pj = j0 - DB;
        // C[64 x 128] -= L[R, prev] L[block rows, prev]^T : warp tile 16 x 32, K = 2 x 64
        const int wr = (warp >> 2) * 16, wc = (warp & 3) * 32;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll 1
        for (int h = 0; h < 2; h++) {
            if (h) __syncthreads();   // every warp is done with the first half
            load_half(Pa, A + (int64_t)r0 * p.ld + pj + h * (DB / 2), p.ld, RT64, rows_valid, tid, A);
            load_half(Pb, A + (int64_t)j0 * p.ld + pj + h * (DB / 2), p.ld, DB, nb, tid, A);
            cp_async_wait_all();
            __syncthreads();
            if (h == 0) stamp(1);
            warp_mma<2, 4>(acc, 0, DB / 2, [&](int i, int k) { return Pa[(wr + i) * LDH_ + k]; },
                           [&](int k, int j) { return Pb[(wc + j) * LDH_ + k]; }, lane);
        }
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double* dst = Ct + (wr + i * 8 + g) * LDS_ + wc + j * 8 + 2 * q;
                dst[0] -= acc[i][j][0];
                dst[1] -= acc[i][j][1];
            }
    } else {
        cp_async_wait_all();
    }
    __syncthreads();   // Lb is free again, Ct holds the updated tile
    stamp(2);
    // as the 32-row form below, two 8x8 blocks (rows r8 and r8 + 32) per warp
    const int r8 = (warp >> 2) * 8, c8 = (warp & 3) * 8;
    for (int c = 0; c < DB / SBW; c++) {
        const int c0 = c * SBW;
        if (c0 >= nb) break;
        if (tid == 0)
            while (ld_acquire(&sync[2]) <= c) __nanosleep(20);
        __syncthreads();
        stamp(3 + c);
        for (int e = tid; e < (DB - c0) * SLAB_CH; e += NT) {
            const int i = c0 + e / SLAB_CH, cc = c0 + 2 * (e % SLAB_CH);
            cp_async16(Lb + i * LDS_ + cc, pub + i * LDS_ + cc, 16);
        }
        cp_async_wait_all();
        __syncthreads();
        double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int k = 0; k < c8 + 8; k += 4) {
            const double t = k + q <= c8 + g ? Lb[(c0 + k + q) * LDS_ + c0 + c8 + g + 1] : 0.0;
            dmma(x[0][0], x[0][1], Ct[(r8 + g) * LDS_ + c0 + k + q], t);
            dmma(x[1][0], x[1][1], Ct[(r8 + 32 + g) * LDS_ + c0 + k + q], t);
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 2; h++) {
            Ct[(r8 + 32 * h + g) * LDS_ + c0 + c8 + 2 * q] = x[h][0];
            Ct[(r8 + 32 * h + g) * LDS_ + c0 + c8 + 2 * q + 1] = x[h][1];
        }
        __syncthreads();
        // columns c0 .. c0+31 of these rows are final: they go home now, in the shadow of DIAG's next slab
        for (int e = tid; e < RT64 * (SBW / 2); e += NT) {
            const int r = e >> 4, cc = c0 + (e & 15) * 2;
            if (r < rows_valid) {
                double* dst = A + (int64_t)(r0 + r) * p.ld + j0 + cc;
                if (cc + 1 < nb) *reinterpret_cast<double2*>(dst) = make_double2(Ct[r * LDS_ + cc], Ct[r * LDS_ + cc + 1]);
                else if (cc < nb) dst[0] = Ct[r * LDS_ + cc];
            }
        }
        for (int c1 = c0 + SBW; c1 < nb; c1 += SBW) {
            double a[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, bb[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int k = 0; k < SBW; k += 8) {
                const double l0 = Lb[(c1 + c8 + g) * LDS_ + c0 + k + q], l1 = Lb[(c1 + c8 + g) * LDS_ + c0 + k + 4 + q];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    dmma(a[h][0], a[h][1], Ct[(r8 + 32 * h + g) * LDS_ + c0 + k + q], l0);
                    dmma(bb[h][0], bb[h][1], Ct[(r8 + 32 * h + g) * LDS_ + c0 + k + 4 + q], l1);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double* dst = Ct + (r8 + 32 * h + g) * LDS_ + c1 + c8 + 2 * q;   // this warp's own blocks
                dst[0] -= a[h][0] + bb[h][0];
                dst[1] -= a[h][1] + bb[h][1];
            }
        }
    }
    stamp(7);
    stamp(8);
}

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
    double* pub = p.pub + b * (DB * LDS_);
    int* sync = p.sync + (b * p.nblk + p.blk) * 4;
    const int j0 = p.j0, n = p.n;
    const int nb = min(DB, n - j0);          // columns of this block (< 128 only for the last one)
    const int nsy = p.prologue ? NSYRKD : 0;
    if (tid == 0) role_s[0] = p.roles == 2 ? nsy + 1 + (int)blockIdx.x : atomicAdd(&sync[0], 1);
    __syncthreads();
    const int ticket = role_s[0];

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
        {   // the 32x32 block itself (one 16-byte chunk per thread), so the update is not a dependent global round trip
            const int r = tid >> 4, c = (tid & 15) * 2;
            int bytes = (n - (rj + c)) * 8;
            bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
            if (ri + r >= n) bytes = 0;
            cp_async16(Lb + r * LDS_ + c, bytes ? A + (int64_t)(ri + r) * p.ld + rj + c : A, bytes);
        }
        cp_async_wait_all();
        __syncthreads();
        if (ticket == 0) STEP_STAMP(0, 1);
        // warp w: 8x8 block (w / 4, w % 4) of the 32x32 output, K = 128 on two accumulator chains
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
            const double* old = Lb + (r8 + g) * LDS_ + c8 + 2 * q;
            if (gj <= gi && gj < n) dst[0] = old[0] - (acc[0][0][0] + acc2[0][0][0]);
            if (gj + 1 <= gi && gj + 1 < n) dst[1] = old[1] - (acc[0][0][1] + acc2[0][0][1]);
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
                while (ld_acquire(&sync[1]) < NSYRKD) __nanosleep(20);
            __syncthreads();
        }
        STEP_STAMP(1, 1);
        load_tile(S, A + (int64_t)j0 * p.ld + j0, p.ld, DB, nb, nb, tid, A);
        cp_async_wait_all();
        __syncthreads();
        STEP_STAMP(1, 2);
        if (tid < DB && tid >= nb) S[tid * LDS_ + tid] = 1.0;   // identity padding of the last block
        __syncthreads();

        double* rdg = red;   // 1 / L(k,k); `red` is only needed for the log-determinant at the very end
        // The SMSP arbiter serves the highest warp id first: the pivot warp is warp 15, the followers 14, 13, 12 (one per
        // SMSP), the helpers warps 0..11 -- so the latency-critical chain never queues behind helper work.
        constexpr int NH = NWARP - 4, HT = NH * 32;    // helper warps 0..11
        constexpr int PW = NWARP - 1;                  // pivot warp
        for (int c = 0; c < DB / SBW; c++) {
            const int c0 = c * SBW;
            const int nfollow = DB / SBW - 1 - c;        // 32-row groups below the diagonal sub-block
            if (warp == PW) {
                if (p.stamps && blockIdx.y == 0 && c == 0) {
                    const long long t = clock64();
                    if (lane == 0) p.stamps[(p.blk * 3 + 1) * 16 + 13] = t;
                }
                slab_pivot(S, rdg, c0, nfollow, lane);
                if (p.stamps && blockIdx.y == 0 && c == 0) {
                    const long long t = clock64();
                    if (lane == 0) p.stamps[(p.blk * 3 + 1) * 16 + 14] = t;
                }
            } else if (warp >= PW - nfollow && warp < PW) {
                slab_follow(S, rdg, c0, PW - warp, nfollow, lane);
            } else if (warp < NH && c > 0) {
                // helpers, in the shadow of slab c: what slab c-1 left behind
                const int p0 = c0 - SBW;                 // previous slab's columns
                // one helper warp inverts the diagonal sub-block (T_(c-1)) while the others publish the slab's rows below it;
                // then its own 32 rows (L and T), the flag, and only then the rest of SYRK(c-1), which no one waits for
                if (warp == NH - 1) inv32_warp(S, rdg, p0, lane);
                else publish_slab(S, pub, A, p.ld, j0, nb, c - 1, tid, HT - 32, SBW);
                named_bar(5, HT);
                publish_slab(S, pub, A, p.ld, j0, nb, c - 1, tid, HT, 0, SBW);
                __threadfence();
                named_bar(6, HT);
                if (tid == 0) st_release(&sync[2], c);
                {
                    // rest of SYRK(c-1): rows >= c0 + 32, columns c0 + 32 .. row (slab c only touches columns < c0 + 32)
                    const int R0 = c0 + SBW, nrb = (DB - R0) / 8;
                    blocks_mma(
                        warp, NH, nrb * (nrb + 1) / 2,
                        [&](int x, int& r0, int& cc0, int& klo, int& khi) {
                            klo = 0;
                            khi = SBW;
                            int y = x, rb = 0;
                            while (y > rb) {
                                y -= rb + 1;
                                rb++;
                            }
                            r0 = R0 + rb * 8;
                            cc0 = R0 + y * 8;
                        },
                        [&](int i, int k) { return S[i * LDS_ + p0 + k]; },
                        [&](int k, int j) { return S[j * LDS_ + p0 + k]; },
                        [&](int i, int j, double a, double bb) {
                            if (j <= i) S[i * LDS_ + j] -= a;
                            if (j + 1 <= i) S[i * LDS_ + j + 1] -= bb;
                        },
                        lane);
                }
            }
            __syncthreads();
            STEP_STAMP(1, 3 + 2 * c);
            if (c0 + SBW >= DB) break;
            // critical part of SYRK(c): columns c0+32 .. c0+63 of every row below (lower part), K = 32 -- what slab c+1 reads
            {
                const int R0 = c0 + SBW, nrb = (DB - R0) / 8;
                blocks_mma(
                    warp, NWARP, nrb * 4,
                    [&](int x, int& r0, int& cc0, int& klo, int& khi) {
                        r0 = R0 + (x >> 2) * 8;
                        cc0 = R0 + (x & 3) * 8;
                        klo = 0;
                        khi = cc0 <= r0 ? SBW : 0;      // blocks above the diagonal: nothing to do
                    },
                    [&](int i, int k) { return S[i * LDS_ + c0 + k]; },
                    [&](int k, int j) { return S[j * LDS_ + c0 + k]; },
                    [&](int i, int j, double a, double bb) {
                        if (j <= i) S[i * LDS_ + j] -= a;
                        if (j + 1 <= i) S[i * LDS_ + j + 1] -= bb;
                    },
                    lane);
                __syncthreads();
                STEP_STAMP(1, 4 + 2 * c);
            }
        }
        if (warp == PW) inv32_warp(S, rdg, DB - SBW, lane);
        __syncthreads();

        // last slab (32 rows), then the flag every row tile's final sub-step waits for
        publish_slab(S, pub, A, p.ld, j0, nb, DB / SBW - 1, tid, NT);
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(&sync[2], DB / SBW);
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
    if (p.rt == RT64) {
        if (tile == 0) STEP_STAMP(2, 0);
        rows64_role(p, A, pub, sync, Lb, As, tile, j0, nb, tid, [&](int idx) {
            if (tile == 0) STEP_STAMP(2, idx);
        });
        return;
    }
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
    // Blocked substitution, right-looking: sub-step c starts as soon as DIAG has published column slab c (T_cc and
    // L(c' > c, c)):  X_c = C_c T_cc^T, then C_c' -= X_c L(c', c)^T for the later column blocks.  One 8x8 block of every
    // 32x32 product per warp.  After DIAG's last slab only one 32x32x32 product is left.
    const int r8 = (warp >> 2) * 8, c8 = (warp & 3) * 8;
    for (int c = 0; c < DB / SBW; c++) {
        const int c0 = c * SBW;
        if (c0 >= nb) break;
        if (tid == 0)
            while (ld_acquire(&sync[2]) <= c) __nanosleep(20);
        __syncthreads();
        if (st0) STEP_STAMP(2, 3 + c);
        for (int e = tid; e < (DB - c0) * SLAB_CH; e += NT) {
            const int i = c0 + e / SLAB_CH, cc = c0 + 2 * (e % SLAB_CH);
            cp_async16(Lb + i * LDS_ + cc, pub + i * LDS_ + cc, 16);
        }
        cp_async_wait_all();
        __syncthreads();
        double x0 = 0.0, x1 = 0.0;
        for (int k = 0; k < c8 + 8; k += 4)
            dmma(x0, x1, Cs[(r8 + g) * LDS_ + c0 + k + q], k + q <= c8 + g ? Lb[(c0 + k + q) * LDS_ + c0 + c8 + g + 1] : 0.0);
        __syncthreads();
        Cs[(r8 + g) * LDS_ + c0 + c8 + 2 * q] = x0;
        Cs[(r8 + g) * LDS_ + c0 + c8 + 2 * q + 1] = x1;
        __syncthreads();
        for (int c1 = c0 + SBW; c1 < nb; c1 += SBW) {
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
            for (int k = 0; k < SBW; k += 8) {
                dmma(a0, a1, Cs[(r8 + g) * LDS_ + c0 + k + q], Lb[(c1 + c8 + g) * LDS_ + c0 + k + q]);
                dmma(b0, b1, Cs[(r8 + g) * LDS_ + c0 + k + 4 + q], Lb[(c1 + c8 + g) * LDS_ + c0 + k + 4 + q]);
            }
            double* dst = Cs + (r8 + g) * LDS_ + c1 + c8 + 2 * q;   // this warp's own block
            dst[0] -= a0 + b0;
            dst[1] -= a1 + b1;
        }
    }
    if (st0) STEP_STAMP(2, 7);
    // write X back (rows < nrows, columns < nb)
    for (int e = tid; e < SBW * (DB / 2); e += NT) {
        const int r = e >> 6, c = (e & 63) * 2;
        if (r < rows_valid) {
            double* dst = A + (int64_t)(r0 + r) * p.ld + j0 + c;
            if (c + 1 < nb) *reinterpret_cast<double2*>(dst) = make_double2(Cs[r * LDS_ + c], Cs[r * LDS_ + c + 1]);
            else if (c < nb) dst[0] = Cs[r * LDS_ + c];
        }
    }
    if (st0) STEP_STAMP(2, 8);
}

static long long* g_step_stamps = nullptr;

}  // namespace

void set_step_stamps(long long* dev) { g_step_stamps = dev; }
size_t chol_step_pub_doubles(int batch) { return (size_t)batch * DB * LDS_; }
static int g_rows_tile = 0;   // 0: by CTA count (below), else forced 32 / 64
void set_step_rows_tile(int rows) { g_rows_tile = rows == 32 ? 32 : rows == 64 ? 64 : 0; }
int step_rows_tile() { return g_rows_tile; }
// 64-row tiles once the step's CTAs would take more than half the chip (the tiles wait on their SMs for DIAG while the
// previous step's trailing update runs on the other stream); below that the 32-row form ends ~3 us earlier per step.
// profiles/r2_rows_tile.txt: n = 2048 LL+gradient 0.80 -> 0.73 ms, 4096 3.78 -> 3.65, two 1500-row experts 0.74 -> 0.66.
static int pick_rows_tile(int below, int batch) {
    if (g_rows_tile) return g_rows_tile;
    static const int sms = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    return (int64_t)batch * (NSYRKD + 1 + cdiv(std::max(below, 0), SBW)) > sms / 2 ? 64 : 32;
}
int chol_step_ctas(int n, int nrows, int j0, int batch) {
    const int nb = std::min(DB, n - j0), below = nrows - (j0 + nb);
    return NSYRKD + 1 + (below > 0 ? cdiv(below, pick_rows_tile(below, batch)) : 0);
}

void launch_chol_step(double* A, int64_t ld, int64_t sA, int n, int nrows, int j0, double* pub, double* logdet_part, int nblk,
                      int* sync, int prologue, int batch, cudaStream_t st, int roles) {
    static bool configured = false;
    if (!configured) {
        CUGP_CUDA(cudaFuncSetAttribute(chol_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM));
        configured = true;
    }
    StepArgs p{};
    p.A = A; p.ld = ld; p.sA = sA; p.n = n; p.nrows = nrows; p.j0 = j0;
    p.pub = pub; p.logdet_part = logdet_part; p.nblk = nblk; p.blk = j0 / DB;
    p.sync = sync; p.prologue = prologue; p.stamps = g_step_stamps;
    const int nb = std::min(DB, n - j0);
    const int below = nrows - (j0 + nb);
    p.rt = pick_rows_tile(below, batch);
    p.nrow_tiles = below > 0 ? cdiv(below, p.rt) : 0;
    p.roles = roles;
    const int head = (prologue ? NSYRKD : 0) + 1;
    const int ctas = roles == 1 ? head : roles == 2 ? p.nrow_tiles : head + p.nrow_tiles;
    if (ctas <= 0) return;
    chol_step_kernel<<<dim3((unsigned)ctas, (unsigned)batch), NT, STEP_SMEM, st>>>(p);
    CUGP_CUDA(cudaGetLastError());
}

}  // namespace cugp
