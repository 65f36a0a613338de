// gp.cu -- orchestration of the GP hot path on one GPU (see gp.cuh).
#include "gp.cuh"
#include "cholstep.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace cugp {

namespace {
template <typename T>
void dalloc(T*& p, size_t count) {
    if (p) return;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) throw CudaError{e, __FILE__, __LINE__};
    p = static_cast<T*>(q);
}
template <typename T>
void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// blocked factorisations built from the DMMA GEMM and the diagonal-block kernel
// ------------------------------------------------------------------------------------------------
static cudaEvent_t prof_event(GpBatch::Prof* prof) {
    if (prof->used == prof->ev.size()) {
        cudaEvent_t e;
        CUGP_CUDA(cudaEventCreate(&e));
        prof->ev.push_back(e);
    }
    return prof->ev[prof->used++];
}

// Outer block width of the two-level factorisation: the trailing update runs with K = outer width, so its
// C tiles are read and written once per outer step instead of once per 128 columns (the K = 128 update is
// bound by that traffic: 8 flop/B).  Small matrices keep narrow outer blocks so the update still fills the GPU.
static long g_epoch = 0;
long tuning_epoch() { return g_epoch; }
void bump_tuning_epoch() { g_epoch++; }
// measured (profiles/r1_graph_replay.txt): n = 1500 LL+grad 1.28 -> 1.18 ms, but 4.50 -> 4.80 ms at n = 4096 and
// 36.1 -> 37.4 ms at n = 10 000 -- once the bulk updates matter, graph branches lose the stream-priority scheduling
// the look-ahead relies on
static int g_graph_max_n = 2048;
void set_graph_max_n(int n) { g_graph_max_n = n; bump_tuning_epoch(); }
// Inverse overlapped with the factorisation: only where the factorisation is bound by its chain of block steps and
// leaves SMs idle.  Measured (profiles/r1_overlap_inverse.txt, LL + gradient per evaluation): n = 3000 3.02 -> 2.68 ms,
// 4096 4.54 -> 4.22, 6000 10.45 -> 10.06, 8192 21.0 -> 22.0 (slower), 16 x 1500 3.75 -> 3.60.  Capping the background
// GEMMs to fewer SMs (32 / 64 / 96) only made the background the bottleneck, so the default cap is the whole chip.
// Round 2 (fused steps: the chain is twice as fast and leaves less idle time): n = 2500 1.66 -> 1.54 ms, 3000 2.19 -> 2.22,
// 4096 3.78 -> 4.12, 6000 9.40 -> 10.01 (profiles/r2_overlap_inverse.txt): the bound moved down to 2800.
static int g_overlap_max_n = 2800, g_overlap_cap = 148;
void set_overlap_inverse(int max_n, int cap) {
    if (max_n >= 0) g_overlap_max_n = max_n;
    if (cap > 0) g_overlap_cap = cap;
    bump_tuning_epoch();
}
static int g_pred_chunk = 0;  // > 0: at most this many test points per prediction chunk (tests: several chunks at small m)
void set_pred_chunk(int v) { g_pred_chunk = v; }
static int g_lookahead = 1;
void set_lookahead(int v) { g_lookahead = v; bump_tuning_epoch(); }
bool lookahead_enabled() { return g_lookahead != 0; }
// Fused block step (cholstep.cu): one launch per 128 columns instead of diag kernel + TRSM GEMM + next-block update GEMM.
// Used whenever the outer width is 128 (the latency-bound regime).  0 restores round 1's launch chain (A/B runs, tests).
static int g_fused_step = 1, g_fused_max_batch = 10;
void set_fused_step(int v) { g_fused_step = v; bump_tuning_epoch(); }
void set_fused_max_batch(int v) { g_fused_max_batch = v; bump_tuning_epoch(); }
// Identity rows: for small n the factorisation carries n more appended rows that start as I and end as L^-T -- the inverse
// of the factor costs no launch chain of its own (the TRTRI recursion is 12 dependent launches, 310 us at n = 1500 against
// 420 us for the factorisation itself).  The row tiles of the fused step do the extra work in the diagonal CTA's shadow.
// Bound (profiles/r2_idrows_bound.txt, LL + gradient): n = 2500 1.54 -> 1.21 ms, 3000 2.15 -> 1.87, 3500 2.87 -> 2.78, 4096 3.66 -> 3.96 --
// from there the rank-128 updates of the (2n+1) x n array stream it through HBM once per block column.
static int g_idrows_max_n = 3500;
void set_idrows_max_n(int n) { g_idrows_max_n = n; bump_tuning_epoch(); }
int idrows_max_n() { return g_idrows_max_n; }
static int g_potrf_nb = 0;  // 0: by size; otherwise forced (cugp_set_tuning("potrf_nb", v) or CUGP_POTRF_NB)
void set_potrf_outer_width(int nb) { g_potrf_nb = nb; bump_tuning_epoch(); }
int potrf_outer_width(int n) {
    static const int env = [] {
        const char* e = std::getenv("CUGP_POTRF_NB");
        return e ? std::atoi(e) : 0;
    }();
    const int forced = g_potrf_nb ? g_potrf_nb : env;
    if (forced >= kDiag) return forced / kDiag * kDiag;
    // measured with look-ahead (profiles/r1_cholesky_nb_sweep.txt): wider panels only pay once the trailing update
    // dominates the chain of diagonal blocks
    if (n >= 24576) return 1024;
    if (n >= 6000) return 512;   // (round 2, look-ahead inside the panel: 512 also wins at 6000 and 8192, profiles/r2_panel_lookahead.txt)
    if (n >= 5000) return 256;
    return kDiag;
}

bool fused_step_applies(int n, int batch) {
    return g_fused_step && potrf_outer_width(n) == kDiag && batch <= g_fused_max_batch;
}

// One outer panel [J0, Jend): right-looking over its 128-column blocks, full height, updates confined to the panel.
// `nrows` >= n: rows n.. of A are appended right-hand sides (y^T): they ride through the TRSM and the updates like
// any other row below the diagonal, which performs the forward substitution L z = y for free (z^T ends up in row n).
static int g_kinv_stream = 0;   // 1: identity-row path accumulates K^-1 on the third stream -- SLOWER (profiles/r2_kinv_stream.txt:
// two experts 0.66 -> 0.74 ms, n = 2048 0.72 -> 0.93: its CTAs take the SMs the next step's launch needs)
void set_kinv_stream(int v) { g_kinv_stream = v; bump_tuning_epoch(); }
static int g_fused_gemm_cap = 0;   // (measured slower, profiles/r2_kinv_stream.txt) > 0: the fused path's main-stream GEMMs leave the step launch's SMs alone (at least this many SMs stay theirs)
void set_fused_gemm_cap(int v) { g_fused_gemm_cap = v; bump_tuning_epoch(); }
static int g_kinv_group = 1;   // identity-row path: K^-1 accumulated every this many block columns; > 1 measured slower
// (profiles/r2_kinv_stream.txt: the bigger bursts delay U2 and the last one cannot hide behind the chain)
void set_kinv_group(int v) { g_kinv_group = v; bump_tuning_epoch(); }
static int g_split_ctas = 100;   // the step's row tiles as a second launch (after DIAG, no waiting CTAs) above this many CTAs; 0: the SM
// count (profiles/r2_panel_lookahead.txt: n = 10000 13.26 -> 13.02 ms, 8192 8.24 -> 8.10; chain-bound sizes stay below it)
void set_step_split_ctas(int v) { g_split_ctas = v; bump_tuning_epoch(); }
static int g_panel_lookahead = 1;   // wide outer panels: look-ahead inside the panel too (profiles/r2_panel_lookahead.txt)
void set_panel_lookahead(int v) { g_panel_lookahead = v; bump_tuning_epoch(); }
static int g_id_init_sparse = 1;   // identity rows / K^-1 accumulator: only the parts that are read, one launch
void set_id_init_sparse(int v) { g_id_init_sparse = v; bump_tuning_epoch(); }
static int g_fused_panel = 1;   // outer width > 128: the diagonal block + TRSM of every 128-column block as the fused step
void set_fused_panel(int v) { g_fused_panel = v; bump_tuning_epoch(); }

static void potrf_panel(double* A, int64_t ld, int64_t sA, int n, int nrows, int J0, int Jend, double* invd, int64_t sInvd,
                        double* logdet_part, int nblk, int batch, cudaStream_t st, long* launches, FusedCtx* fx = nullptr) {
    static const int sms = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    const bool fused = fx && fx->sync && fx->pub && g_fused_step && g_fused_panel && batch <= g_fused_max_batch;
    for (int j0 = J0; j0 < Jend; j0 += kDiag) {
        const int blk = j0 / kDiag;
        const int nbk = std::min(kDiag, n - j0);  // < 128 only for the last block
        const int r0 = j0 + nbk;
        const int m = nrows - r0;
        double* A21 = A + (int64_t)r0 * ld + j0;
        if (fused) {
            // diagonal block (slab factorisation, 32x32 inverses) and the TRSM of every row below in the fused step kernel
            // (cholstep.cu) without its prologue: one launch while its CTAs fit the chip together, else two
            const bool split = (int64_t)batch * chol_step_ctas(n, nrows, j0, batch) > (g_split_ctas > 0 && batch == 1 ? g_split_ctas : sms);
            if (!split) {
                launch_chol_step(A, ld, sA, n, nrows, j0, fx->pub, logdet_part, nblk, fx->sync, 0, batch, st, 0);
                if (launches) ++*launches;
            } else {
                launch_chol_step(A, ld, sA, n, nrows, j0, fx->pub, logdet_part, nblk, fx->sync, 0, batch, st, 1);
                launch_chol_step(A, ld, sA, n, nrows, j0, fx->pub, logdet_part, nblk, fx->sync, 0, batch, st, 2);
                if (launches) *launches += 2;
            }
            if (m <= 0) break;
        } else {
            launch_potrf_diag(A, ld, sA, n, j0, invd, sInvd, logdet_part, nblk, blk, batch, st);
            if (launches) ++*launches;
            if (m <= 0) break;
            // TRSM panel: L21 = A21 * inv(L11)^T, in place (each CTA owns whole rows of the panel).
            GemmParams p{};
            p.A = A21; p.lda = ld; p.sA = sA;
            p.B = invd + (int64_t)blk * kDiag * kDiag; p.ldb = kDiag; p.sB = sInvd;
            p.C = A21; p.ldc = ld; p.sC = sA;
            p.M = m; p.N = nbk; p.K = nbk;
            p.alpha = 1.0; p.beta = 0.0;
            p.batch = batch;
            launch_gemm(p, true, true, GEMM_TALL, st);
            if (launches) ++*launches;
        }
        // remaining columns of the panel: A[r0:, r0:Jend] -= L21 L21[0:w]^T, lower trapezoid
        const int w = Jend - r0;
        if (w <= 0) continue;
        GemmParams q{};
        q.A = A21; q.lda = ld; q.sA = sA;
        q.B = A21; q.ldb = ld; q.sB = sA;
        q.C = A + (int64_t)r0 * (ld + 1); q.ldc = ld; q.sC = sA;
        q.M = m; q.N = w; q.K = kDiag;
        q.alpha = -1.0; q.beta = 1.0;
        q.batch = batch;
        q.lower_tiles = 1;
        launch_gemm(q, true, true, pick_config(m, w, batch, true), st);
        if (launches) ++*launches;
    }
}

// A[c0:, c0:c1] -= P[c0:, :] P[c0:c1, :]^T with P = L[:, J0:Jend] (lower trapezoid; c1 == n: the whole trailing block)
static void potrf_trailing(double* A, int64_t ld, int64_t sA, int nrows, int J0, int Jend, int c0, int c1, int batch,
                           cudaStream_t st, long* launches, GpBatch::Prof* prof, int cap_sms = 0) {
    const int m = nrows - c0, w = c1 - c0;
    if (m <= 0 || w <= 0) return;
    GemmParams q{};
    q.A = A + (int64_t)c0 * ld + J0; q.lda = ld; q.sA = sA;
    q.B = q.A; q.ldb = ld; q.sB = sA;
    q.C = A + (int64_t)c0 * (ld + 1); q.ldc = ld; q.sC = sA;
    q.M = m; q.N = w; q.K = Jend - J0;
    q.alpha = -1.0; q.beta = 1.0;
    q.batch = batch;
    q.lower_tiles = 1;
    if (prof) CUGP_CUDA(cudaEventRecord(prof_event(prof), st));
    const GemmConfig cfg = pick_config(m, w, batch, true);
    if (cap_sms > 0) q.max_ctas = cfg == GEMM_BIG ? cap_sms : 2 * cap_sms;
    launch_gemm(q, true, true, cfg, st);
    if (prof) {
        CUGP_CUDA(cudaEventRecord(prof_event(prof), st));
        // lower trapezoid, 2 flop per MAC: w(w+1)/2 + (m-w)w outputs
        prof->flops += (double)batch * 2.0 * ((double)w * (w + 1.0) / 2.0 + (double)(m - w) * w) * (double)(Jend - J0);
        prof->count++;
    }
    if (launches) ++*launches;
}

// Outer width 128 with the fused step kernel.  Block column J is final after step(J); step(J) itself applies block
// column J-1 to block column J (its prologue), so the trailing update of panel J only covers the columns right of
// block J+1 -- U2(J) -- and, with look-ahead, runs on the main stream while the panel stream is already at step(J+1).
// Both step(J+2) and U2(J) touch block column J+2: step(J+2) waits for U2(J).
// `nblk_stride`: blocks per matrix of logdet_part / sync (the caller may pass both advanced to the first block of a
// trailing sub-matrix); `nblk`: block columns of THIS n x n factorisation.
static void potrf_fused(double* A, int64_t ld, int64_t sA, int n, int nrows, double* invd, int64_t sInvd, double* logdet_part,
                        int nblk_stride, int nblk, int batch, cudaStream_t st, long* launches, GpBatch::Prof* prof,
                        PotrfLookahead* la, FusedCtx* fx) {
    int* stepsync = fx->sync;
    double* steppub = fx->pub;
    const int id_rows = fx->id_rows;
    fx->invd_done = false;   // the 128x128 inverses of the diagonal blocks are not a by-product here: GpBatch::ensure_invd
    // With the identity rows, column block J of U = L^-T (rows < Jend) is final after step(J): K^-1 = U U^T is accumulated
    // block column by block column on the main stream, in the chain's shadow, instead of one LAUUM launch after it.
    int cap_sms = 0;   // set below (look-ahead only): SMs the main stream's GEMMs may hold while the chain runs
    // ... in groups of g_kinv_group block columns: nothing reads K^-1 before the end, and a K = 256 / 384 update reads and
    // writes the accumulator half / a third as often (rows >= a block column's own Jend of U are still zero)
    const int kgroup = std::max(1, g_kinv_group);
    auto kinv_update = [&](int J0, int Jend, cudaStream_t s) {
        if (!id_rows || !fx->kinv) return;
        const int J = J0 / kDiag;
        if ((J + 1) % kgroup != 0 && Jend < n) return;    // not the last block column of its group
        J0 = J / kgroup * kgroup * kDiag;
        GemmParams q{};
        const double* U = A + (int64_t)(n + 1) * ld + J0;     // identity rows start below the right-hand-side row
        q.A = U; q.lda = ld; q.sA = sA;
        q.B = U; q.ldb = ld; q.sB = sA;
        q.C = fx->kinv; q.ldc = ld; q.sC = sA;
        q.M = Jend; q.N = Jend; q.K = Jend - J0;
        q.alpha = 1.0; q.beta = 1.0;
        q.batch = batch;
        q.lower_tiles = 1;
        const GemmConfig cfg = pick_config(Jend, Jend, batch, true);
        if (cap_sms > 0) q.max_ctas = cfg == GEMM_BIG ? cap_sms : 2 * cap_sms;
        launch_gemm(q, true, true, cfg, s);
        if (launches) ++*launches;
    };
    if (nblk == nblk_stride)   // (a trailing sub-matrix: the caller has zeroed the whole array)
        CUGP_CUDA(cudaMemsetAsync(stepsync, 0, (size_t)batch * nblk * 4 * sizeof(int), st));
    const bool ahead = la && la->st2 && nblk >= 3 && lookahead_enabled();
    // one launch per step only while all its CTAs are resident together (the row tiles wait on their SMs for the
    // diagonal CTA); a wider batch takes the diagonal part and the row tiles as two launches
    static const int sms = [] {
        int dev = 0, v = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : 148;
    }();
    // The last id_rows appended rows are the identity: row i of it is zero left of column i, so at the step of block
    // column J only its first Jend rows take part -- in the step itself and in the trailing update of panel J.
    const int base_rows = nrows - id_rows;
    auto rows_at = [&](int Jend) { return base_rows + std::min(id_rows, Jend); };
    int widest = 0;
    for (int J0 = 0; J0 < n; J0 += kDiag) widest = std::max(widest, chol_step_ctas(n, rows_at(std::min(n, J0 + kDiag)), J0, batch));
    // (the lower bound only for a single matrix: three or four 1500-row experts per GPU lose 10-20 % with it,
    // profiles/r2_panel_lookahead.txt)
    const bool split = (int64_t)batch * widest > (g_split_ctas > 0 && batch == 1 ? g_split_ctas : sms);
    auto step = [&](int J0, int prologue, cudaStream_t s) {
        const int nr = rows_at(std::min(n, J0 + kDiag));
        if (!split) {
            launch_chol_step(A, ld, sA, n, nr, J0, steppub, logdet_part, nblk_stride, stepsync, prologue, batch, s, 0);
            if (launches) ++*launches;
        } else {
            launch_chol_step(A, ld, sA, n, nr, J0, steppub, logdet_part, nblk_stride, stepsync, prologue, batch, s, 1);
            launch_chol_step(A, ld, sA, n, nr, J0, steppub, logdet_part, nblk_stride, stepsync, prologue, batch, s, 2);
            if (launches) *launches += 2;
        }
    };
    if (!ahead) {
        for (int J = 0; J < nblk; J++) {
            const int J0 = J * kDiag, Jend = std::min(n, J0 + kDiag), Jend2 = std::min(n, Jend + kDiag);
            step(J0, J > 0, st);
            potrf_trailing(A, ld, sA, rows_at(Jend), J0, Jend, Jend2, n, batch, st, launches, prof);
            kinv_update(J0, Jend, st);
        }
        return;
    }
    if (g_fused_gemm_cap > 0 && !split) cap_sms = std::max(g_fused_gemm_cap, sms - batch * widest);
    cudaStream_t s2 = la->st2;
    while ((int)la->ev.size() < 2 * nblk + 2) {
        cudaEvent_t e;
        CUGP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        la->ev.push_back(e);
    }
    auto evP = [&](int J) { return la->ev[2 * J]; };
    auto evU = [&](int J) { return la->ev[2 * J + 1]; };
    cudaEvent_t ev_start = la->ev[2 * nblk], ev_end = la->ev[2 * nblk + 1];
    CUGP_CUDA(cudaEventRecord(ev_start, st));
    CUGP_CUDA(cudaStreamWaitEvent(s2, ev_start, 0));
    for (int J = 0; J < nblk; J++) {
        const int J0 = J * kDiag, Jend = std::min(n, J0 + kDiag), Jend2 = std::min(n, Jend + kDiag);
        if (J >= 2) CUGP_CUDA(cudaStreamWaitEvent(s2, evU(J - 2), 0));
        step(J0, J > 0, s2);
        CUGP_CUDA(cudaEventRecord(evP(J), s2));                // "block column J is final" (also for the overlapped inverse)
        // K^-1 += U[:, J] U[:, J]^T has no reader before the end: on its own stream it fills the gaps U2 leaves instead of
        // delaying U2(J+1) (two experts per GPU: the main stream's GEMMs, not the chain, set the pace)
        cudaStream_t sk = id_rows && fx->kinv && fx->kinv_stream && fx->kinv_done ? fx->kinv_stream : nullptr;
        if (sk) {
            CUGP_CUDA(cudaStreamWaitEvent(sk, evP(J), 0));
            kinv_update(J0, Jend, sk);
        }
        if (Jend2 >= n && (sk || !(id_rows && fx->kinv) || ((J + 1) % kgroup != 0 && Jend < n))) continue;   // nothing to do on the main stream
        CUGP_CUDA(cudaStreamWaitEvent(st, evP(J), 0));
        if (Jend2 < n) {
            potrf_trailing(A, ld, sA, rows_at(Jend), J0, Jend, Jend2, n, batch, st, launches, prof, cap_sms);   // U2(J)
            CUGP_CUDA(cudaEventRecord(evU(J), st));
        }
        if (!sk) kinv_update(J0, Jend, st);
    }
    CUGP_CUDA(cudaEventRecord(ev_end, s2));
    CUGP_CUDA(cudaStreamWaitEvent(st, ev_end, 0));
    if (id_rows && fx->kinv && fx->kinv_stream && fx->kinv_done) {
        CUGP_CUDA(cudaEventRecord(fx->kinv_done, fx->kinv_stream));
        CUGP_CUDA(cudaStreamWaitEvent(st, fx->kinv_done, 0));
    }
    if (nblk == nblk_stride) la->panel_events = true;   // ev[2J] = "block column J is final" (outer width 128)
}

static int g_adaptive_nb = 0;   // (measured slower, profiles/r2_adaptive_nb.txt: the last wide panel loses its overlap) 1: the outer width follows the REMAINING size; its last part takes the fused 128 path
void set_adaptive_nb(int v) { g_adaptive_nb = v; bump_tuning_epoch(); }
static bool nb_is_forced() { return g_potrf_nb != 0 || std::getenv("CUGP_POTRF_NB") != nullptr; }

void potrf_blocked(double* A, int64_t ld, int64_t sA, int n, double* invd, int64_t sInvd, double* logdet_part, int batch,
                   cudaStream_t st, long* launches, GpBatch::Prof* prof, PotrfLookahead* la, int rhs_rows, FusedCtx* fx) {
    if (prof && !prof->on) prof = nullptr;
    const int nrows = n + rhs_rows;
    const int nblk = cdiv(n, kDiag);
    const int NB = potrf_outer_width(n);
    if (la) la->panel_events = false;
    // (a batch wider than ~10 matrices is throughput bound: there the batched GEMM chain of round 1 is as fast)
    const bool fused_ok = fx && fx->sync && fx->pub && g_fused_step && batch <= g_fused_max_batch;
    if (fused_ok && NB == kDiag) {
        potrf_fused(A, ld, sA, n, nrows, invd, sInvd, logdet_part, nblk, nblk, batch, st, launches, prof, la, fx);
        return;
    }
    const bool fpanel = fused_ok && g_fused_panel;
    if (fx) {
        if (fx->id_rows) throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};   // identity rows: outer width 128 only
        fx->invd_done = !fpanel;
    }
    if (fpanel) CUGP_CUDA(cudaMemsetAsync(fx->sync, 0, (size_t)batch * nblk * 4 * sizeof(int), st));
    // Panel boundaries.  With the fused step available the outer width follows the size that is LEFT (a wide panel only
    // pays while its trailing update dominates the chain of block steps), and once that calls for 128 the rest -- a
    // Cholesky of the fully updated Schur complement -- takes the fused 128-column path with its own look-ahead.
    std::vector<int> P;
    int tail = n;
    const bool adaptive = fpanel && g_adaptive_nb && !nb_is_forced();
    for (int s0 = 0; s0 < n;) {
        const int w = adaptive ? potrf_outer_width(n - s0) : NB;
        if (adaptive && w == kDiag) {
            tail = s0;
            break;
        }
        P.push_back(s0);
        s0 += w;
    }
    const int npanels = (int)P.size();
    auto pend = [&](int J) { return J + 1 < npanels ? P[J + 1] : tail; };   // right edge of panel J (tail == n without a tail)
    // One outer panel.  With the fused step: the panel is itself a (tall) factorisation of outer width 128 -- the chain of
    // block steps (each applying the previous block column to its own, `prologue`) on the inner stream, the in-panel
    // trailing updates on the panel's stream: the same look-ahead one level down, ~35 instead of ~70 us per 128 columns.
    auto panel = [&](int J0, int Jend, cudaStream_t s) {
        if (fpanel && g_panel_lookahead && la && la->st_inner) {   // (look-ahead off: potrf_fused runs it on one stream)
            FusedCtx sub = *fx;
            sub.sync = fx->sync + (int64_t)(J0 / kDiag) * 4;
            PotrfLookahead in;
            in.st2 = la->st_inner;
            in.ev.swap(la->ev_inner);
            potrf_fused(A + (int64_t)J0 * (ld + 1), ld, sA, Jend - J0, nrows - J0, invd, sInvd, logdet_part + J0 / kDiag, nblk,
                        cdiv(Jend - J0, kDiag), batch, s, launches, nullptr, &in, &sub);
            in.ev.swap(la->ev_inner);
            return;
        }
        potrf_panel(A, ld, sA, n, nrows, J0, Jend, invd, sInvd, logdet_part, nblk, batch, s, launches, fx);
    };
    auto run_tail = [&] {
        if (tail >= n) return;
        FusedCtx sub = *fx;
        sub.sync = fx->sync + (int64_t)(tail / kDiag) * 4;
        potrf_fused(A + (int64_t)tail * (ld + 1), ld, sA, n - tail, nrows - tail, invd, sInvd, logdet_part + tail / kDiag, nblk,
                    cdiv(n - tail, kDiag), batch, st, launches, prof, la, &sub);
    };
    if (!la || !la->st2 || npanels < 3 || !lookahead_enabled()) {
        for (int J = 0; J < npanels; J++) {
            const int J0 = P[J], Jend = std::min(n, pend(J));
            panel(J0, Jend, st);
            // trailing update: A[Jend:, Jend:] -= P P^T with P = L[Jend:, J0:Jend], lower tiles, K = outer width
            potrf_trailing(A, ld, sA, nrows, J0, Jend, Jend, n, batch, st, launches, prof);
        }
        run_tail();
        return;
    }
    // Look-ahead of one panel: the panel stream (high priority) factors panel J+1 while the main stream applies the
    // bulk of panel J's trailing update.  The update of panel J is split at the next panel's right edge:
    //   U1(J) = columns of panel J+1 (panel stream, on the critical path),  U2(J) = everything right of it (main stream).
    // Both U1(J) and U2(J-1) write the columns of panel J+1, so U1(J) waits for U2(J-1).
    cudaStream_t s2 = la->st2;
    while ((int)la->ev.size() < 2 * npanels + 2) {
        cudaEvent_t e;
        CUGP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        la->ev.push_back(e);
    }
    auto evP = [&](int J) { return la->ev[2 * J]; };
    auto evU = [&](int J) { return la->ev[2 * J + 1]; };
    cudaEvent_t ev_start = la->ev[2 * npanels], ev_end = la->ev[2 * npanels + 1];
    CUGP_CUDA(cudaEventRecord(ev_start, st));          // K is complete on the main stream
    CUGP_CUDA(cudaStreamWaitEvent(s2, ev_start, 0));
    for (int J = 0; J < npanels; J++) {
        const int J0 = P[J], Jend = std::min(n, pend(J));
        // the last two-level panel before a fused tail updates ALL remaining columns on the panel stream
        const int Jend2 = J + 1 < npanels ? std::min(n, pend(J + 1)) : n;
        panel(J0, Jend, s2);
        CUGP_CUDA(cudaEventRecord(evP(J), s2));
        if (Jend >= n) break;
        if (J >= 1) CUGP_CUDA(cudaStreamWaitEvent(s2, evU(J - 1), 0));
        potrf_trailing(A, ld, sA, nrows, J0, Jend, Jend, Jend2, batch, s2, launches, nullptr);   // U1(J)
        if (Jend2 >= n) continue;
        CUGP_CUDA(cudaStreamWaitEvent(st, evP(J), 0));
        potrf_trailing(A, ld, sA, nrows, J0, Jend, Jend2, n, batch, st, launches, prof);          // U2(J)
        CUGP_CUDA(cudaEventRecord(evU(J), st));
    }
    CUGP_CUDA(cudaEventRecord(ev_end, s2));
    CUGP_CUDA(cudaStreamWaitEvent(st, ev_end, 0));
    la->panel_events = tail >= n;
    run_tail();
}

void trtri_recursive(const double* L, double* T, double* W, int64_t ld, int64_t sM, int n, const double* invd,
                     int64_t sInvd, int batch, cudaStream_t st, long* launches, int max_ctas) {
    launch_scatter_invdiag(invd, sInvd, T, ld, sM, n, batch, st);
    if (launches) ++*launches;
    for (int64_t h = kDiag; h < n; h *= 2) {
        // pairs (r1 = 2ph, r2 = r1 + h): T21 = -T22 (L21 T11).  `full` pairs have an h x h lower-right block.
        const int full = (int)(n / (2 * h));
        const int64_t pair_stride = 2 * h * (ld + 1);
        for (int pass = 0; pass < 2; pass++) {
            int npairs, n2;
            int64_t base;
            if (pass == 0) {
                npairs = full; n2 = (int)h; base = 0;
            } else {
                const int64_t r2 = (int64_t)full * 2 * h + h;
                if (r2 >= n) break;
                npairs = 1; n2 = (int)(n - r2); base = (int64_t)full * pair_stride;
            }
            if (npairs == 0) continue;
            const int64_t o21 = base + h * ld;        // block (2,1): rows r2.., cols r1..
            const int64_t o11 = base;                 // block (1,1)
            const int64_t o22 = base + h * (ld + 1);  // block (2,2)
            const int tot = npairs * batch;
            GemmParams p{};  // tmp = L21 * T11   (T11 lower: k >= column tile start)
            p.A = L + o21; p.lda = ld;
            p.B = T + o11; p.ldb = ld;
            p.C = W + o21; p.ldc = ld;
            p.M = n2; p.N = (int)h; p.K = (int)h;
            p.alpha = 1.0; p.beta = 0.0;
            p.batch = tot; p.batch_inner = npairs;
            p.sA = p.sB = p.sC = pair_stride;
            p.sA2 = p.sB2 = p.sC2 = sM;
            p.klo_tj = 1;
            p.max_ctas = max_ctas;
            if (npairs == 1) { p.batch_inner = 0; p.sA = p.sB = p.sC = sM; }
            GemmConfig cfg = pick_config(n2, (int)h, tot, false);
            launch_gemm(p, true, false, cfg, st);
            GemmParams q{};  // T21 = -T22 * tmp   (T22 lower: k < row tile end)
            q.A = T + o22; q.lda = ld;
            q.B = W + o21; q.ldb = ld;
            q.C = T + o21; q.ldc = ld;
            q.M = n2; q.N = (int)h; q.K = n2;
            q.alpha = -1.0; q.beta = 0.0;
            q.batch = tot; q.batch_inner = npairs;
            q.sA = q.sB = q.sC = pair_stride;
            q.sA2 = q.sB2 = q.sC2 = sM;
            q.khi_ti = 1;
            q.max_ctas = max_ctas;
            if (npairs == 1) { q.batch_inner = 0; q.sA = q.sB = q.sC = sM; }
            launch_gemm(q, true, false, cfg, st);
            if (launches) *launches += 2;
        }
    }
}

// In-place variants for sizes where three n x n buffers do not fit (n = 100 000: 80 GB each).
// TRTRI: the same recursive doubling, but T overwrites L block by block (T21 = -T22 (L21 T11) only needs L21 until the
// product tmp = L21 T11 exists) and tmp lives in a COMPACT scratch: level h needs n h / 2 doubles at most, n^2 / 4 at the
// top level.  Single matrix.
size_t trtri_inplace_scratch(int n) {
    size_t need = 0;
    for (int64_t h = kDiag; h < n; h *= 2) {
        const int64_t full = n / (2 * h);
        const int64_t r2 = full * 2 * h + h;
        const int64_t rem = r2 < n ? (n - r2) : 0;
        need = std::max<size_t>(need, (size_t)(full * h * h + rem * h));
    }
    return need;
}
void trtri_inplace(double* A, int64_t ld, int n, const double* invd, double* Wc, cudaStream_t st, long* launches) {
    launch_scatter_invdiag(invd, 0, A, ld, 0, n, 1, st);   // diagonal blocks: L_jj -> inv(L_jj) (lower; upper never read)
    if (launches) ++*launches;
    for (int64_t h = kDiag; h < n; h *= 2) {
        const int full = (int)(n / (2 * h));
        const int64_t pair_stride = 2 * h * (ld + 1);
        for (int pass = 0; pass < 2; pass++) {
            int npairs, n2;
            int64_t base;
            double* W;
            if (pass == 0) {
                npairs = full; n2 = (int)h; base = 0; W = Wc;
            } else {
                const int64_t r2 = (int64_t)full * 2 * h + h;
                if (r2 >= n) break;
                npairs = 1; n2 = (int)(n - r2); base = (int64_t)full * pair_stride; W = Wc + (int64_t)full * h * h;
            }
            if (npairs == 0) continue;
            const int64_t o21 = base + h * ld, o11 = base, o22 = base + h * (ld + 1);
            GemmParams p{};  // tmp = L21 * T11   (T11 lower: k >= column tile start)
            p.A = A + o21; p.lda = ld;
            p.B = A + o11; p.ldb = ld;
            p.C = W; p.ldc = h;
            p.M = n2; p.N = (int)h; p.K = (int)h;
            p.alpha = 1.0; p.beta = 0.0;
            p.batch = npairs;
            p.sA = p.sB = pair_stride;
            p.sC = h * h;
            p.klo_tj = 1;
            const GemmConfig cfg = pick_config(n2, (int)h, npairs, false);
            launch_gemm(p, true, false, cfg, st);
            GemmParams q{};  // T21 = -T22 * tmp   (T22 lower: k < row tile end), written over L21
            q.A = A + o22; q.lda = ld;
            q.B = W; q.ldb = h;
            q.C = A + o21; q.ldc = ld;
            q.M = n2; q.N = (int)h; q.K = n2;
            q.alpha = -1.0; q.beta = 0.0;
            q.batch = npairs;
            q.sA = q.sC = pair_stride;
            q.sB = h * h;
            q.khi_ti = 1;
            launch_gemm(q, true, false, cfg, st);
            if (launches) *launches += 2;
        }
    }
}

// LAUUM in place: K^-1 = T^T T (lower) over T, row block by row block from the top.  Row block R = [i0, i0 + ib) of the
// result needs rows >= i0 of T only: the block's own rows are copied to a scratch of ib x ld first (with the strict
// upper part of its diagonal block zeroed: it is the k range's only guard there), rows below are still T.
//   C[R, 0:i0+ib] = S[:, R]^T S[:, 0:i0+ib]  +  T[i0+ib:, R]^T T[i0+ib:, 0:i0+ib]
__global__ void zero_upper_kernel(double* S, int64_t ld, int c0, int ib) {
    const int r = blockIdx.x, c = threadIdx.x + blockIdx.y * blockDim.x;
    if (r < ib && c < ib && c > r) S[(int64_t)r * ld + c0 + c] = 0.0;
}
void lauum_inplace(double* A, int64_t ld, int n, double* S, int ib, cudaStream_t st, long* launches) {
    for (int i0 = 0; i0 < n; i0 += ib) {
        const int rows = std::min(ib, n - i0), cols = i0 + rows;
        CUGP_CUDA(cudaMemcpy2DAsync(S, (size_t)ld * 8, A + (int64_t)i0 * ld, (size_t)ld * 8, (size_t)cols * 8, (size_t)rows,
                                    cudaMemcpyDeviceToDevice, st));
        zero_upper_kernel<<<dim3(rows, cdiv(rows, 256)), 256, 0, st>>>(S, ld, i0, rows);
        CUGP_CUDA(cudaGetLastError());
        GemmParams p{};   // own rows: K = rows
        p.A = S + i0; p.lda = ld;
        p.B = S; p.ldb = ld;
        p.C = A + (int64_t)i0 * ld; p.ldc = ld;
        p.M = rows; p.N = cols; p.K = rows;
        p.alpha = 1.0; p.beta = 0.0;
        p.batch = 1;
        launch_gemm(p, false, false, pick_config(rows, cols, 1, false), st);
        if (launches) *launches += 2;
        const int below = n - (i0 + rows);
        if (below > 0) {
            GemmParams q{};   // rows below: still T
            q.A = A + (int64_t)(i0 + rows) * ld + i0; q.lda = ld;
            q.B = A + (int64_t)(i0 + rows) * ld; q.ldb = ld;
            q.C = A + (int64_t)i0 * ld; q.ldc = ld;
            q.M = rows; q.N = cols; q.K = below;
            q.alpha = 1.0; q.beta = 1.0;
            q.batch = 1;
            launch_gemm(q, false, false, pick_config(rows, cols, 1, false), st);
            if (launches) ++*launches;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GpBatch
// ------------------------------------------------------------------------------------------------
GpBatch::GpBatch(int B_, int n_, int d_, cudaStream_t stream) : B(B_), n(n_), d(d_), Bcap(B_) {
    dp = (int)round_up(d, 2);
    h = make_hyper(theta);
    ld = padded_ld(n);
    nblk = cdiv(n, kDiag);
    if (stream) {
        st = stream;
    } else {
        CUGP_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        own_stream = true;
    }
    {
        int lo = 0, hi = 0;  // the panel stream of the look-ahead Cholesky gets the highest priority
        CUGP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUGP_CUDA(cudaStreamCreateWithPriority(&la.st2, cudaStreamNonBlocking, hi));
        CUGP_CUDA(cudaStreamCreateWithPriority(&la.st_inner, cudaStreamNonBlocking, hi));
    }
    const size_t bn = (size_t)B * n;
    dalloc(X, bn * dp);
    dalloc(y, bn);
    rows_alloc = n <= g_idrows_max_n ? 2 * n + 1 : n + 1;
    dalloc(Kb, (size_t)B * rows_alloc * ld);  // row n of every matrix: y^T, then z^T = (L^-1 y)^T; rows n+1..: I -> L^-T
    dalloc(invd, (size_t)B * nblk * kDiag * kDiag);
    CUGP_CUDA(cudaMemsetAsync(invd, 0, (size_t)B * nblk * kDiag * kDiag * sizeof(double), st));  // upper triangles stay zero
    dalloc(logdet_part, (size_t)B * nblk);
    dalloc(work, bn);
    dalloc(alpha, bn);
    dalloc(scal, (size_t)B * 4);
    dalloc(tpart, trsv_backward_scratch(n, B));
    dalloc(gradout, (size_t)B * 3);
    dalloc(stepsync, (size_t)B * nblk * 4);
    dalloc(steppub, chol_step_pub_doubles(B));
}

GpBatch::~GpBatch() {
    if (st) cudaStreamSynchronize(st);
    dfree(X); dfree(y); dfree(Kb); dfree(invd); dfree(logdet_part); dfree(work); dfree(alpha); dfree(scal);
    dfree(Tb); dfree(Wb); dfree(gradpart); dfree(gradout); dfree(tpart); dfree(stepsync); dfree(steppub); dfree(Wc);
    dfree(Xt); dfree(Ks); dfree(meanpart); dfree(css); dfree(pmean); dfree(pvar); dfree(Xt_all);
    if (hstage) cudaFreeHost(hstage);
    if (hres) cudaFreeHost(hres);
    for (cudaEvent_t e : prof.ev) cudaEventDestroy(e);
    for (cudaEvent_t e : la.ev) cudaEventDestroy(e);
    for (cudaEvent_t e : bwd_ev) cudaEventDestroy(e);
    for (auto* cache : {&graph_potrf, &graph_potrf_rhs, &graph_potrf_id, &graph_inv, &graph_inv_id})
        for (auto& kv : *cache)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (la.st2) {
        cudaStreamSynchronize(la.st2);
        cudaStreamDestroy(la.st2);
    }
    if (la.st_inner) {
        cudaStreamSynchronize(la.st_inner);
        cudaStreamDestroy(la.st_inner);
    }
    for (cudaEvent_t e : la.ev_inner) cudaEventDestroy(e);
    if (st3) {
        cudaStreamSynchronize(st3);
        cudaStreamDestroy(st3);
    }
    if (ev_T) cudaEventDestroy(ev_T);
    if (ev_kinv) cudaEventDestroy(ev_kinv);
    if (own_stream && st) cudaStreamDestroy(st);
}

void* GpBatch::stage(size_t bytes) {
    if (bytes > hstage_bytes) {
        if (hstage) {
            CUGP_CUDA(cudaStreamSynchronize(st));
            cudaFreeHost(hstage);
            hstage = nullptr;
        }
        CUGP_CUDA(cudaMallocHost((void**)&hstage, bytes));
        hstage_bytes = bytes;
    }
    return hstage;
}

void GpBatch::set_data(const double* Xh, const double* yh) {
    const size_t rows = (size_t)B * n;
    CUGP_CUDA(cudaStreamSynchronize(st));  // staging buffer may still feed an earlier copy
    double* s = static_cast<double*>(stage((rows * dp + rows) * sizeof(double)));
    if (dp == d) {
        std::memcpy(s, Xh, rows * d * sizeof(double));
    } else {
        for (size_t r = 0; r < rows; r++) {
            std::memcpy(s + r * dp, Xh + r * d, d * sizeof(double));
            for (int k = d; k < dp; k++) s[r * dp + k] = 0.0;  // a zero extra coordinate adds exactly +0 to |xi-xj|^2
        }
    }
    std::memcpy(s + rows * dp, yh, rows * sizeof(double));
    CUGP_CUDA(cudaMemcpyAsync(X, s, rows * dp * sizeof(double), cudaMemcpyHostToDevice, st));
    CUGP_CUDA(cudaMemcpyAsync(y, s + rows * dp, rows * sizeof(double), cudaMemcpyHostToDevice, st));
    have_data = true;
    invalidate();
}

void GpBatch::set_active(int b) {
    if (b < 1 || b > Bcap) throw CudaError{cudaErrorInvalidValue, __FILE__, __LINE__};
    if (b != B) {
        B = b;
        invalidate();
    }
}

void GpBatch::adopt_device_data(const double* Xd, const double* yd, int b, cudaEvent_t ready) {
    set_active(b);
    if (ready) CUGP_CUDA(cudaStreamWaitEvent(st, ready, 0));
    const size_t rows = (size_t)b * n;
    CUGP_CUDA(cudaMemcpyAsync(X, Xd, rows * dp * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CUGP_CUDA(cudaMemcpyAsync(y, yd, rows * sizeof(double), cudaMemcpyDeviceToDevice, st));
    have_data = true;
    invalidate();
}

void GpBatch::eval_enqueue(bool want_grad) {
    factorize();
    if (want_grad) gradient_launch();
}

void GpBatch::eval_launch(bool want_grad) {
    if (!hres) CUGP_CUDA(cudaMallocHost((void**)&hres, (size_t)Bcap * 8 * sizeof(double)));
    factorize();
    CUGP_CUDA(cudaMemcpyAsync(hres, scal, (size_t)B * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (want_grad) {
        gradient_launch();
        CUGP_CUDA(cudaMemcpyAsync(hres + (size_t)Bcap * 4, gradout, (size_t)B * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    eval_grad = want_grad;
}

void GpBatch::eval_collect(double* ll_out, double* g_out) {
    sync();
    for (int b = 0; b < B; b++) {
        if (ll_out) ll_out[b] = hres[(size_t)b * 4 + 2];
        if (g_out)
            for (int k = 0; k < 3; k++) g_out[(size_t)b * 3 + k] = eval_grad ? hres[(size_t)Bcap * 4 + (size_t)b * 3 + k] : 0.0;
    }
}

void GpBatch::set_theta(const double th[3]) {
    if (th[0] == theta[0] && th[1] == theta[1] && th[2] == theta[2] && have_L) return;  // cached factor stays valid
    theta[0] = th[0]; theta[1] = th[1]; theta[2] = th[2];
    h = make_hyper(theta);
    invalidate();
}

// Replay `body` (a fixed sequence of launches on `st` and the look-ahead stream, independent of theta) as a CUDA graph.
// First use: direct run (kernel attributes get configured, events created).  Second use: stream capture, instantiate,
// launch.  Later: launch.  A failed capture disables the graph for this batch and the caller launches directly.
template <class F>
bool GpBatch::run_graphed(std::map<int, GraphEntry>& cache, F&& body) {
    // a wide batch is throughput bound again (16 x 1500: 3.75 ms direct, 3.84 ms replayed; 2 x 1500: 1.41 -> 1.30 ms)
    if (n > g_graph_max_n || (int64_t)B * n > 4 * (int64_t)g_graph_max_n || prof.on) return false;
    GraphEntry& e = cache[B];
    if (e.failed) return false;
    if (e.exec && e.epoch != tuning_epoch()) {
        cudaGraphExecDestroy(e.exec);
        e.exec = nullptr;
        e.uses = 0;
    }
    if (!e.exec) {
        if (e.uses++ == 0) return false;
        const long before = launches;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            try {
                body();
            } catch (const CudaError&) {
                ok = false;
            }
            ok = (cudaStreamEndCapture(st, &graph) == cudaSuccess) && ok && graph != nullptr;
        }
        e.launches = launches - before;
        launches = before;
        if (ok) ok = cudaGraphInstantiate(&e.exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            cudaGetLastError();
            e.exec = nullptr;
            e.failed = true;
            return false;
        }
        e.epoch = tuning_epoch();
    }
    CUGP_CUDA(cudaGraphLaunch(e.exec, st));
    launches += e.launches;
    la.panel_events = false;   // the per-panel events were graph nodes, not records another stream can wait on
    return true;
}

void GpBatch::build_K(int full) {
    launch_cov_train(X, (int64_t)n * dp, n, dp, h, Kb, ld, mat_stride(), B, full, st);
    launches++;
}

void GpBatch::potrf(bool with_rhs) {
    auto body = [&] {
        FusedCtx fx{stepsync, steppub, 0, nullptr, false};
        potrf_blocked(Kb, ld, mat_stride(), n, invd, (int64_t)nblk * kDiag * kDiag, logdet_part, B, st, &launches, &prof, &la,
                      with_rhs ? 1 : 0, &fx);
        have_invd = fx.invd_done;
    };
    have_invd = !g_fused_step || B > g_fused_max_batch;     // what a replayed graph leaves behind
    if (with_rhs || !run_graphed(graph_potrf, body)) body();   // graph_potrf holds the rhs-free sequence only
}

// Cholesky of the matrix in Kb with y appended as row n: L in place, z = L^-1 y in row n, then
// (quad = z'z = y'K^-1 y, logdet, LL) -- matrixops.cpp:113-185 + covkernel.cpp:127 without a separate forward sweep.
void GpBatch::potrf_with_rhs() {
    // A batch that has needed the inverse of its factor before gets it from the factorisation itself: n more appended
    // rows that start as I and leave as L^-T (small n, fused step path only).
    const bool id = wants_inverse && n <= g_idrows_max_n && rows_alloc >= 2 * n + 1 && fused_step_applies(n, B) && !prof.on;
    auto body = [&] {
        launch_copy_rows(y, n, Kb + (int64_t)n * ld, mat_stride(), n, B, st);
        if (id) {
            if (g_id_init_sparse) {   // one launch, half the bytes (profiles/r2_kinv_stream.txt)
                launch_init_idrows(Tt(), mat_stride(), Wb, mat_stride(), ld, n, B, st);
                launches++;
            } else {
                launch_init_identity(Tt(), ld, mat_stride(), n, B, st);
                launches++;
                for (int b = 0; b < B; b++)   // K^-1 is accumulated block column by block column during the factorisation
                    CUGP_CUDA(cudaMemsetAsync(Wb + (int64_t)b * mat_stride(), 0, (size_t)n * ld * sizeof(double), st));
            }
        }
        FusedCtx fx{stepsync, steppub, id ? n : 0, id ? Wb : nullptr, false};
        if (id && g_kinv_stream) {
            fx.kinv_stream = st3;
            fx.kinv_done = ev_kinv;
        }
        potrf_blocked(Kb, ld, mat_stride(), n, invd, (int64_t)nblk * kDiag * kDiag, logdet_part, B, st, &launches, &prof, &la,
                      id ? 1 + n : 1, &fx);
        have_invd = fx.invd_done;
        const double* zrow = Kb + (int64_t)n * ld;
        launch_ll_finalize(zrow, zrow, mat_stride(), n, logdet_part, nblk, scal, B, st);
        launches += 2;
    };
    if (id) ensure_TW();
    if (id && g_kinv_stream) ensure_st3();
    have_invd = !g_fused_step || B > g_fused_max_batch;     // what a replayed graph leaves behind
    if (!run_graphed(id ? graph_potrf_id : graph_potrf_rhs, body)) body();
    have_Tt = id;
    have_Kinv = id;   // accumulated alongside (lower triangle of Wb)
}

// 128x128 inverses of L's diagonal blocks: a by-product of the round-1 launch chain, one extra launch after the fused
// steps -- paid only by the paths that need them (backward sweep, TRTRI recursion).
void GpBatch::ensure_invd() {
    if (invd_on_st3) join_T();   // the overlapped inverse is producing them on the third stream
    if (have_invd) return;
    launch_trtri_diag(Kb, ld, mat_stride(), n, invd, (int64_t)nblk * kDiag * kDiag, B, st);
    launches++;
    have_invd = true;
}

void GpBatch::prof_begin() {
    prof.used = 0;
    prof.flops = 0.0;
    prof.count = 0;
}

void GpBatch::prof_collect(double* ms, double* flops, long* count) {
    sync();
    double total = 0.0;
    for (size_t i = 0; i + 1 < prof.used; i += 2) {
        float t = 0.f;
        CUGP_CUDA(cudaEventElapsedTime(&t, prof.ev[i], prof.ev[i + 1]));
        total += t;
    }
    if (ms) *ms = total;
    if (flops) *flops = prof.flops;
    if (count) *count = prof.count;
}

void GpBatch::join_T() {
    if (!t_inflight) return;
    CUGP_CUDA(cudaStreamWaitEvent(st, ev_T, 0));
    t_inflight = false;
    if (invd_on_st3) {
        have_invd = true;
        invd_on_st3 = false;
    }
}

// T = L^-1 in row groups of 512, each enqueued on a third stream behind the event "the panel holding the group's last
// column is final":  T_gg by recursive doubling inside the group, then  T[g, 0:r0] = -T_gg (L[g, 0:r0] T[0:r0, 0:r0])
// (a left-looking block row: it needs only rows 0..r1 of L and the groups above).  While the factorisation is a chain
// of 128-column block steps (n <~ 6000: 2.5 ms for 0.65 ms worth of flops at n = 4096) the other SMs do this work; the
// GEMMs are capped to g_overlap_cap CTAs so the chain's high-priority launches always find free SMs.  The last group
// can only start when the chain has ended and runs uncapped.
void GpBatch::ensure_st3() {
    if (st3) return;
    CUGP_CUDA(cudaStreamCreateWithFlags(&st3, cudaStreamNonBlocking));
    CUGP_CUDA(cudaEventCreateWithFlags(&ev_T, cudaEventDisableTiming));
    CUGP_CUDA(cudaEventCreateWithFlags(&ev_kinv, cudaEventDisableTiming));
}

void GpBatch::enqueue_trtri_overlapped() {
    const int NB = potrf_outer_width(n);
    constexpr int G = 512;
    ensure_st3();
    const int64_t sI = (int64_t)nblk * kDiag * kDiag;
    // (Accumulating K^-1 = sum_g T[g, :]^T T[g, :] behind every finished group as well was measured and dropped: n = 4096
    // 4.22 -> 4.48 ms, n = 6000 10.06 -> 10.76 ms -- more background CTAs delay the chain more than the LAUUM they hide.)
    for (int r0 = 0; r0 < n; r0 += G) {
        const int r1 = std::min(n, r0 + G), rows = r1 - r0;
        CUGP_CUDA(cudaStreamWaitEvent(st3, la.ev[2 * ((r1 - 1) / NB)], 0));
        if (!have_invd) {   // fused steps leave no 128x128 inverses behind: this group's, from its finished diagonal blocks
            launch_trtri_diag(Kb, ld, mat_stride(), n, invd, sI, B, st3, r0 / kDiag, cdiv(r1 - r0, kDiag));
            launches++;
        }
        const int cap = r1 == n ? 0 : g_overlap_cap;
        const int64_t o = (int64_t)r0 * (ld + 1);
        trtri_recursive(Kb + o, Tb + o, Wb + o, ld, mat_stride(), rows, invd + (int64_t)(r0 / kDiag) * kDiag * kDiag, sI, B, st3,
                        &launches, cap);
        if (r0 > 0) {
            GemmParams p{};  // tmp = L[g, 0:r0] T[0:r0, 0:r0]   (T lower: k >= column tile start)
            p.A = Kb + (int64_t)r0 * ld; p.lda = ld; p.sA = mat_stride();
            p.B = Tb; p.ldb = ld; p.sB = mat_stride();
            p.C = Wb + (int64_t)r0 * ld; p.ldc = ld; p.sC = mat_stride();
            p.M = rows; p.N = r0; p.K = r0;
            p.alpha = 1.0; p.beta = 0.0;
            p.batch = B;
            p.klo_tj = 1;
            p.max_ctas = cap;
            const GemmConfig cfg = pick_config(rows, r0, B, false);
            launch_gemm(p, true, false, cfg, st3);
            GemmParams q{};  // T[g, 0:r0] = -T_gg tmp   (T_gg lower: k < row tile end)
            q.A = Tb + o; q.lda = ld; q.sA = mat_stride();
            q.B = Wb + (int64_t)r0 * ld; q.ldb = ld; q.sB = mat_stride();
            q.C = Tb + (int64_t)r0 * ld; q.ldc = ld; q.sC = mat_stride();
            q.M = rows; q.N = r0; q.K = rows;
            q.alpha = -1.0; q.beta = 0.0;
            q.batch = B;
            q.khi_ti = 1;
            q.max_ctas = cap;
            launch_gemm(q, true, false, cfg, st3);
            launches += 2;
        }
    }
    CUGP_CUDA(cudaEventRecord(ev_T, st3));
    t_inflight = true;
    t_valid = true;
    invd_on_st3 = !have_invd;   // complete once the third stream is joined
}

void GpBatch::factorize() {
    if (have_L) return;
    join_T();       // an earlier inverse may still be reading L on the third stream
    build_K(0);
    potrf_with_rhs();
    have_L = true;
    // a batch that has computed gradients / predictions before (T exists) will want T = L^-1 again: start it now
    if (Tb && Wb && la.panel_events && !have_Tt && n <= g_overlap_max_n) enqueue_trtri_overlapped();
}

// alpha = L^-T z.  With T = L^-1 at hand (gradient / prediction paths) it is one streaming pass alpha = T^T z;
// otherwise the blocked backward sweep over L (matrixops.cpp:156-164).
void GpBatch::solve() {
    if (have_alpha) return;
    factorize();
    const double* zrow = Kb + (int64_t)n * ld;
    if (have_Tt) {
        launch_gemv_upper(Tt(), ld, mat_stride(), n, zrow, mat_stride(), alpha, n, B, st);   // alpha = L^-T z
        launches++;
    } else if (have_T) {
        launch_gemv_t(Tb, ld, mat_stride(), n, zrow, mat_stride(), alpha, n, tpart, B, st);
        launches += 2;
    } else {
        const int64_t sI = (int64_t)nblk * kDiag * kDiag;
        ensure_invd();
        launch_copy_rows(zrow, mat_stride(), work, n, n, B, st);  // the sweep consumes its right-hand side
        const int nev = trsv_backward_events(n);
        while ((int)bwd_ev.size() < nev) {
            cudaEvent_t e;
            CUGP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            bwd_ev.push_back(e);
        }
        launch_trsv_backward(Kb, ld, mat_stride(), n, invd, sI, work, alpha, n, tpart, B, st,
                             lookahead_enabled() ? la.st2 : nullptr, bwd_ev.data(), nev);
        launches += 1 + 3 * cdiv(n, 1024);  // per panel: chain, near update, far update
    }
    have_alpha = true;
}

void GpBatch::ensure_TW() {
    dalloc(Tb, (size_t)Bcap * rows_alloc * ld);  // same batch stride as Kb
    dalloc(Wb, (size_t)Bcap * rows_alloc * ld);
}

void GpBatch::trtri() {
    wants_inverse = true;
    if (have_T) return;
    factorize();
    if (t_valid) {   // computed alongside the factorisation
        join_T();
        t_valid = false;
        have_T = true;
        have_Kinv = false;
        return;
    }
    join_T();
    ensure_TW();
    ensure_invd();
    trtri_recursive(Kb, Tb, Wb, ld, mat_stride(), n, invd, (int64_t)nblk * kDiag * kDiag, B, st, &launches);
    have_T = true;
    have_Kinv = false;
}

void GpBatch::lauum() {
    if (have_Kinv) return;
    factorize();
    if (have_Tt) {
        ensure_TW();
        GemmParams p{};  // Kinv = U U^T with U = L^-T upper triangular, row-major: k >= max(i, j) = row-tile start (lower tiles)
        p.A = Tt(); p.lda = ld; p.sA = mat_stride();
        p.B = Tt(); p.ldb = ld; p.sB = mat_stride();
        p.C = Wb; p.ldc = ld; p.sC = mat_stride();
        p.M = n; p.N = n; p.K = n;
        p.alpha = 1.0; p.beta = 0.0;
        p.batch = B;
        p.lower_tiles = 1;
        p.klo_ti = 1;
        launch_gemm(p, true, true, pick_config(n, n, B, true), st);
        launches++;
        have_Kinv = true;
        return;
    }
    trtri();
    GemmParams p{};  // Kinv = T^T T: k >= max(i,j) = row-tile start on the lower tile set
    p.A = Tb; p.lda = ld; p.sA = mat_stride();
    p.B = Tb; p.ldb = ld; p.sB = mat_stride();
    p.C = Wb; p.ldc = ld; p.sC = mat_stride();
    p.M = n; p.N = n; p.K = n;
    p.alpha = 1.0; p.beta = 0.0;
    p.batch = B;
    p.lower_tiles = 1;
    p.klo_ti = 1;
    launch_gemm(p, false, false, pick_config(n, n, B, true), st);
    launches++;
    have_Kinv = true;
}

void GpBatch::scalars(double* out4) {
    factorize();
    CUGP_CUDA(cudaMemcpyAsync(out4, scal, (size_t)B * 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync();
}

void GpBatch::loglik(double* ll_out) {
    std::vector<double> s((size_t)B * 4);
    scalars(s.data());
    for (int b = 0; b < B; b++) ll_out[b] = s[(size_t)b * 4 + 2];
}

// Largest n whose gradient keeps L, L^-1 and K^-1 in three buffers; above it (or when they do not fit the device) the
// inverse is formed IN PLACE over L (one n x n buffer + n^2 / 4 of scratch): the factor is gone afterwards.
static int64_t g_inplace_min_n = 60000;
void set_inplace_inverse_min_n(int64_t v) { g_inplace_min_n = v; }

bool GpBatch::use_inplace_inverse() {
    if (B != 1 || Bcap != 1) return false;
    if (n >= g_inplace_min_n) return true;
    if (Tb && Wb) return false;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return false;
    return (double)free_b < 2.1 * (double)rows_alloc * (double)ld * 8.0;   // Tb + Wb would not fit
}

void GpBatch::gradient_inplace() {
    // alpha first (backward sweep over L), then T over L, then K^-1 over T; the trace streams K^-1 from Kb
    solve();
    ensure_invd();
    const size_t need = trtri_inplace_scratch(n), srows = 1024;
    if (need + srows * (size_t)ld > wc_cap) {
        sync();
        dfree(Wc);
        dalloc(Wc, need + srows * (size_t)ld);
        wc_cap = need + srows * (size_t)ld;
    }
    trtri_inplace(Kb, ld, n, invd, Wc, st, &launches);
    lauum_inplace(Kb, ld, n, Wc + need, (int)srows, st, &launches);
    dalloc(gradpart, grad_trace_partials(n, Bcap));
    launch_grad_trace(X, (int64_t)n * dp, n, dp, h, Kb, ld, mat_stride(), alpha, n, gradpart, gradout, B, st);
    launches += 2;
    // Kb no longer holds L: the next request for the factor rebuilds it (alpha and the scalars stay valid)
    have_L = have_T = have_Kinv = have_Tt = have_invd = false;
    kinv_in_kb = true;
}

void GpBatch::gradient_launch() {
    wants_inverse = true;
    factorize();
    if (!have_Tt && use_inplace_inverse()) {
        gradient_inplace();
        return;
    }
    if (have_Tt) {
        // L^-T and K^-1 = L^-T L^-1 came out of the factorisation: only alpha = L^-T z is left
        solve();
        lauum();
    } else {
        if (!have_T && !have_alpha && !have_Kinv) {
            // the whole inverse chain (TRTRI recursion, alpha = T^T z, LAUUM) is theta independent: one graph replay
            ensure_TW();
            ensure_invd();
            auto body = [&] {
                have_T = have_alpha = have_Kinv = false;
                trtri();   // before solve(): alpha then is a single pass over T
                solve();
                lauum();
            };
            if (run_graphed(graph_inv, body)) have_T = have_alpha = have_Kinv = true;
        }
        trtri();
        solve();
        lauum();
    }
    dalloc(gradpart, grad_trace_partials(n, Bcap));
    launch_grad_trace(X, (int64_t)n * dp, n, dp, h, Wb, ld, mat_stride(), alpha, n, gradpart, gradout, B, st);
    launches += 2;
}

void GpBatch::gradient(double* g_out) {
    gradient_launch();
    CUGP_CUDA(cudaMemcpyAsync(g_out, gradout, (size_t)B * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync();
}

void GpBatch::get_alpha(double* out) {
    solve();
    CUGP_CUDA(cudaMemcpyAsync(out, alpha, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToHost, st));
    sync();
}

void GpBatch::ensure_pred(int mc) {
    if (mc <= pred_cap) return;
    sync();
    dfree(Xt); dfree(Ks); dfree(meanpart); dfree(css); dfree(pmean); dfree(pvar);
    const size_t tiles64 = (size_t)cdiv(n, 64);
    dalloc(Xt, (size_t)mc * dp);
    dalloc(Ks, (size_t)Bcap * mc * ld);
    dalloc(meanpart, (size_t)Bcap * tiles64 * mc);
    dalloc(css, (size_t)Bcap * tiles64 * mc);
    dalloc(pmean, (size_t)Bcap * mc);
    dalloc(pvar, (size_t)Bcap * mc);
    pred_cap = mc;
}

// Predictive moments of every GP in the batch at m test points (covkernel.cpp:277-306):
//   mean = k*' alpha ;  var = sf2 + sn2 - |L^-1 k*|^2  ( = sf2 + sn2 - k*' K^-1 k* ).
// Xt_dev: [m][dp] device, zero padded.  mean_h/var_h: [B][m] host (may be null).  PQ_dev: [2][m] device
// product-of-experts moments (may be null).  Nothing here waits for the device unless host outputs were asked for.
void GpBatch::predict_dev(const double* Xt_dev, int m, double* mean_h, double* var_h, double* PQ_dev, int accumulate) {
    if (m <= 0) return;
    wants_inverse = true;
    factorize();
    if (!have_Tt) trtri();
    solve();
    // chunk the test set so Kstar stays near 2 GB
    int64_t cap = (int64_t)(2.0e9 / ((double)Bcap * ld * 8.0));
    int mc = (int)std::min<int64_t>(m, std::max<int64_t>(64, cap / 64 * 64));
    if (g_pred_chunk > 0) mc = std::min(mc, (int)round_up(g_pred_chunk, 64));
    ensure_pred(mc);
    const int tiles_j = cdiv(n, kCovTile);
    for (int t0 = 0; t0 < m; t0 += mc) {
        const int cur = std::min(mc, m - t0);
        launch_cov_cross(Xt_dev + (size_t)t0 * dp, cur, X, (int64_t)n * dp, n, dp, h, alpha, n, Ks, ld, (int64_t)mc * ld, meanpart,
                         (int64_t)tiles_j * mc, B, st);
        GemmParams p{};  // V = T Kstar^T ; only the column sums of squares are kept
        p.A = have_Tt ? Tt() : Tb; p.lda = ld; p.sA = mat_stride();   // T (rows) or T^T = L^-T (rows): same product
        p.B = Ks; p.ldb = ld; p.sB = (int64_t)mc * ld;
        p.M = n; p.N = cur; p.K = n;
        p.alpha = 1.0; p.beta = 0.0;
        p.batch = B;
        p.khi_ti = 1;
        p.colsumsq = css;
        const GemmConfig cfg = pick_config(n, cur, B, false);
        const int tiles_m = cdiv(n, gemm_tile_m(cfg));
        p.sCss = (int64_t)tiles_m * cur;
        launch_gemm(p, !have_Tt, true, cfg, st);
        // meanpart was written with row length `cur` (ni) and batch stride tiles_j*mc
        launch_predict_finalize(meanpart, tiles_j, css, tiles_m, cur, h, pmean, pvar, mc, (int64_t)tiles_j * mc, p.sCss,
                                B, st);
        launches += 3;
        if (PQ_dev) {
            launch_poe_accumulate(pmean, pvar, mc, B, cur, PQ_dev + t0, PQ_dev + m + t0, accumulate, st);
            launches++;
        }
        if (mean_h && var_h) {
            for (int b = 0; b < B; b++) {
                CUGP_CUDA(cudaMemcpyAsync(mean_h + (size_t)b * m + t0, pmean + (size_t)b * mc, cur * sizeof(double),
                                          cudaMemcpyDeviceToHost, st));
                CUGP_CUDA(cudaMemcpyAsync(var_h + (size_t)b * m + t0, pvar + (size_t)b * mc, cur * sizeof(double),
                                          cudaMemcpyDeviceToHost, st));
            }
        }
    }
    if (mean_h && var_h) sync();
}

// Host test points: packed to [m][dp] and uploaded ONCE (not per chunk), then predict_dev.
void GpBatch::predict(const double* Xt_h, int m, double* mean_h, double* var_h, double* PQ_dev, int accumulate) {
    if (m <= 0) return;
    if (m > xt_all_cap) {
        sync();
        dfree(Xt_all);
        dalloc(Xt_all, (size_t)m * dp);
        xt_all_cap = m;
    }
    sync();   // the staging buffer may still feed an earlier copy
    double* s = static_cast<double*>(stage((size_t)m * dp * sizeof(double)));
    if (dp == d) {
        std::memcpy(s, Xt_h, (size_t)m * d * sizeof(double));
    } else {
        for (int r = 0; r < m; r++) {
            std::memcpy(s + (size_t)r * dp, Xt_h + (size_t)r * d, d * sizeof(double));
            for (int k = d; k < dp; k++) s[(size_t)r * dp + k] = 0.0;
        }
    }
    CUGP_CUDA(cudaMemcpyAsync(Xt_all, s, (size_t)m * dp * sizeof(double), cudaMemcpyHostToDevice, st));
    predict_dev(Xt_all, m, mean_h, var_h, PQ_dev, accumulate);
    if (!(mean_h && var_h)) sync();   // callers of the host-pointer variant expect the staging buffer to be free again
}

}  // namespace cugp
