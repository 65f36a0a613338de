// cugp_shim/matrixops.h -- the free functions of common/matrixops.h:5-25 with the reference's signatures.
// The O(n^3) ones (get_cholesky, compute_chol_and_det, vector_Kinvy_using_cholesky, compute_K_inverse, the matrix
// substitutions) run on the GPU through libcugp.so; the BLAS-1/2 helpers and printers are the trivial host
// loops they are in the reference (they are fused into kernels on the hot path and kept only for callers).
#ifndef CUGP_SHIM_MATRIXOPS_H
#define CUGP_SHIM_MATRIXOPS_H
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <utility>
#include <vector>

#include "../cugp.h"

namespace cugp_shim {
inline void pack_square(double** M, int n, std::vector<double>& out) {
    out.resize((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[(size_t)i * n + j] = M[i][j];
}
inline void unpack_square(const std::vector<double>& in, int n, double** M) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) M[i][j] = in[(size_t)i * n + j];
}
inline bool mok(int rc, const char* what) {
    if (rc == CUGP_OK) return true;
    std::fprintf(stderr, "cugp: %s failed (%d): %s\n", what, rc, cugp_last_error());
    return false;
}
}  // namespace cugp_shim

// ---- host utilities (matrixops.cpp:26-56, 216-260, 319-328, 437-448) ----
inline void vector_matrix_multiply(double* v, double** M, int n, double* out) {
    for (int j = 0; j < n; j++) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += v[i] * M[i][j];
        out[j] = s;
    }
}
inline void matrix_vector_multiply(double** M, double* v, int n, double* out) {
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int j = 0; j < n; j++) s += M[i][j] * v[j];
        out[i] = s;
    }
}
inline double vector_vector_multiply(double* a, double* b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}
inline void print_matrix(double** M, int r, int c) {
    for (int i = 0; i < r; i++) {
        for (int j = 0; j < c; j++) std::printf("%lf ", M[i][j]);
        std::printf("\n");
    }
}
inline void print_vector(double* v, int n) {
    for (int i = 0; i < n; i++) std::printf("%lf ", v[i]);
    std::printf("\n");
}
inline void subtract_vec(double* a, double* b, double* out, int n) {
    for (int i = 0; i < n; i++) out[i] = a[i] - b[i];
}
inline double dotproduct_vec(double* a, double* b, int n) { return vector_vector_multiply(a, b, n); }
inline void subtract_matrices(double** A, double** B, double** out, int r, int c) {
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) out[i][j] = A[i][j] - B[i][j];
}
inline void get_outer_product(double* a, double* b, double** out, int n) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[i][j] = a[i] * b[j];
}
inline void make_identity(double** M, int n) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) M[i][j] = (i == j) ? 1.0 : 0.0;
}
inline void elementwise_matrixmultiply(double** A, double** B, double** out, int r, int c) {
    for (int i = 0; i < r; i++)
        for (int j = 0; j < c; j++) out[i][j] = A[i][j] * B[i][j];
}

// ---- GPU-backed (matrixops.cpp:68-108, 232-234, 264-316, 330-372, 383-435) ----
inline void get_cholesky(double** in, double** out, int n) {
    std::vector<double> A, L((size_t)n * n);
    cugp_shim::pack_square(in, n, A);
    if (cugp_shim::mok(cugp_cholesky(A.data(), L.data(), n), "cugp_cholesky")) cugp_shim::unpack_square(L, n, out);
}
inline std::pair<double, double> compute_chol_and_det(double** K, double* y, int n) {
    std::vector<double> A;
    cugp_shim::pack_square(K, n, A);
    double quad = std::numeric_limits<double>::quiet_NaN(), logdet = quad;
    cugp_shim::mok(cugp_chol_and_det(A.data(), y, n, &quad, &logdet), "cugp_chol_and_det");
    return std::make_pair(quad, logdet);
}
inline void vector_Kinvy_using_cholesky(double** K, double* y, double* ans, int n) {
    std::vector<double> A;
    cugp_shim::pack_square(K, n, A);
    cugp_shim::mok(cugp_kinv_y(A.data(), y, ans, n), "cugp_kinv_y");
}
inline void compute_K_inverse(double** K, double** out, int n) {
    std::vector<double> A, Ki((size_t)n * n);
    cugp_shim::pack_square(K, n, A);
    if (cugp_shim::mok(cugp_k_inverse(A.data(), Ki.data(), n), "cugp_k_inverse")) cugp_shim::unpack_square(Ki, n, out);
}
// L T = B  (matrixops.cpp:330-340) and U T = B with U upper triangular (matrixops.cpp:361-372): n right-hand sides
inline void matrix_forward_substitution(double** L, double** B, double** T, int n) {
    std::vector<double> l, b, t((size_t)n * n);
    cugp_shim::pack_square(L, n, l);
    cugp_shim::pack_square(B, n, b);
    if (cugp_shim::mok(cugp_tri_solve_matrix(l.data(), b.data(), t.data(), n, 0), "cugp_tri_solve_matrix")) cugp_shim::unpack_square(t, n, T);
}
inline void matrix_backward_substitution(double** U, double** B, double** T, int n) {
    std::vector<double> u, b, t((size_t)n * n);
    cugp_shim::pack_square(U, n, u);
    cugp_shim::pack_square(B, n, b);
    if (cugp_shim::mok(cugp_tri_solve_matrix(u.data(), b.data(), t.data(), n, 1), "cugp_tri_solve_matrix")) cugp_shim::unpack_square(t, n, T);
}
#endif
