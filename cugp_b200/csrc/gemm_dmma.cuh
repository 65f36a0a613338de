// gemm_dmma.cuh -- FP64 tensor-core (DMMA) GEMM building block for the blocked factorisations.
//
// Blackwell's tcgen05 has no f64 kind; the FP64 tensor pipe is reached with warp-level
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA.8x8x4), accumulators in registers.  One kernel template
// serves every dense contraction on the hot path:
//   SYRK trailing update   A22 -= L21 L21^T          (A k-contig, B k-contig, lower tiles only)
//   TRSM via inverse       L21  = A21 invL11^T       (in place, CTA owns whole rows)
//   TRTRI recursion        tmp  = L21 T11 ; T21 = -T22 tmp   (B row-contig; triangular k-ranges)
//   LAUUM                  Kinv = T^T T              (A and B row-contig, k >= max(i,j))
//   predictive variance    V    = T Kstar^T          (epilogue reduces column sums of squares)
// Operands are staged global->shared by a producer warp group (cp.async + mbarrier, 3-4 stages of k-depth 32);
// shared rows are padded by 4 doubles so the 8-byte fragment loads of a half-warp fall in distinct bank pairs.
#pragma once
#include "common.cuh"

namespace cugp {

struct GemmParams {
    const double* A;
    const double* B;
    double* C;
    int64_t lda, ldb, ldc;
    int M, N, K;
    double alpha, beta;      // C = alpha * op(A) op(B) + beta * C   (beta == 0: C is not read)
    int64_t sA, sB, sC;      // batch strides in elements (blockIdx.y = batch index)
    int batch;               // total batch count = batch_inner * outer
    int batch_inner;         // 0/1: single level.  >1: index b -> (b % inner) * s? + (b / inner) * s?2
    int64_t sA2, sB2, sC2;   // outer batch strides (used when batch_inner > 1)
    int lower_tiles;         // 1: only output tiles with ti >= tj are computed (needs BM == BN and M >= N: lower trapezoid)
    int klo_ti, klo_tj;      // restrict k >= ti*BM / k >= tj*BN   (triangular operand structure)
    int khi_ti, khi_tj;      // restrict k <  (ti+1)*BM / (tj+1)*BN
    double* colsumsq;        // non-null: write per-row-tile column sums of squares [tiles_m][N] instead of C
    int64_t sCss;            // batch stride of colsumsq
    int max_ctas;            // > 0: the launch uses at most this many CTAs (= SMs), each walking more tiles -- background
                             // work that must leave SMs free for a latency-critical stream
    int64_t tiles_per_mat;   // filled by the launcher: output tiles of one matrix (raster order)
    int tiles_per_cta;       // filled by the launcher: work items (tile, batch) one CTA walks
    int cta_stride;          // filled by the launcher: distance between a CTA's successive tiles (SM count; 1 if tpc == 1)
    int raster_w;            // filled by the launcher: tile columns per raster strip
};

enum GemmConfig { GEMM_BIG = 0, GEMM_TALL = 1, GEMM_SMALL = 2 };

// a_kc / b_kc: operand stored with k contiguous ([rows][K]) -- otherwise stored [K][rows].
void launch_gemm(const GemmParams& p, bool a_kc, bool b_kc, GemmConfig cfg, cudaStream_t stream);
// Picks BIG when the grid fills the chip, SMALL otherwise.
GemmConfig pick_config(int M, int N, int batch, bool lower_tiles);
int gemm_tile_m(GemmConfig cfg);
// Tuning: tiles walked per CTA (0 = by grid size).
void set_gemm_tiles_per_cta(int v);
void set_gemm_small_two(int v);
void set_gemm_big_min_tiles(int v);      // 64x64 configuration: 1 (default) three stages and two CTAs per SM, 0 four stages and one
void set_gemm_raster_width(int v);   // tile columns per raster strip (0 = default)

}  // namespace cugp
