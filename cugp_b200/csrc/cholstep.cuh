// cholstep.cuh -- fused 128-column block step of the blocked Cholesky (see cholstep.cu).
#pragma once
#include "common.cuh"

namespace cugp {

// One launch = SYRK prologue of the diagonal block (if `prologue`) + 128x128 POTRF with its four 32x32 diagonal
// inverses + (prologue and) TRSM of every 32-row tile below, rows [.., nrows) (nrows > n: appended right-hand sides).
// `sync`: int[batch][nblk][4], zeroed before the first step of a factorisation.  `pub`: chol_step_pub_doubles(batch)
// doubles of scratch (the diagonal CTA publishes its tile there, row block by row block, for the row tiles).  With
// `prologue` the block column j0 - 128 must be final and must NOT have been applied to block column j0 yet (the launch
// applies it).  The 128x128 inverses of the diagonal blocks are NOT produced: launch_trtri_diag() afterwards.
void launch_chol_step(double* A, int64_t ld, int64_t sA, int n, int nrows, int j0, double* pub, double* logdet_part, int nblk,
                      int* sync, int prologue, int batch, cudaStream_t st, int roles = 0);
// roles = 0: one launch (the row tiles wait for the diagonal CTA on their SMs: only when every CTA of the launch is
// resident at once); roles = 1 then roles = 2: diagonal part, then the row tiles, as two launches (wide batches).
int chol_step_ctas(int n, int nrows, int j0, int batch);   // CTAs per matrix of a roles = 0 launch of `batch` matrices
size_t chol_step_pub_doubles(int batch);
// Rows per ROWS tile: 0 (default) = by CTA count, or forced 32 / 64.  Fewer, taller tiles = fewer CTAs waiting on SMs.
void set_step_rows_tile(int rows);
int step_rows_tile();

// Tuning aid: device buffer [nblk][3][16] of globaltimer stamps written by every step (nullptr: off).
void set_step_stamps(long long* dev);

}  // namespace cugp
