/*
 * cugp_oracle.c -- CPU restatement of the cuGP reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see cugp_oracle.h for the rules: never linked or called by the product path).
 *
 * The loops below restate, in plain C and in the same order, the arithmetic of
 *   /root/reference/common/matrixops.cpp, distributed_gp/covkernel.cpp, distributed_gp/BCM.cpp.
 * Matrices are held as arrays of separately allocated rows (the reference's `double**`) so cache
 * behaviour, and therefore timing, matches the reference.  Build with -O3 -ffp-contract=off
 * (the reference Makefiles use plain `-O3` on x86-64, where g++ emits no fused multiply-adds).
 */
#include "cugp_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers: row-pointer matrices */
static double **mat_new(int r, int c) {
    double **M = (double **)malloc(sizeof(double *) * (size_t)(r > 0 ? r : 1));
    for (int i = 0; i < r; i++) M[i] = (double *)malloc(sizeof(double) * (size_t)(c > 0 ? c : 1));
    return M;
}
static void mat_free(double **M, int r) {
    for (int i = 0; i < r; i++) free(M[i]);
    free(M);
}
static double **mat_from_flat(const double *A, int r, int c) {
    double **M = mat_new(r, c);
    for (int i = 0; i < r; i++) memcpy(M[i], A + (size_t)i * c, sizeof(double) * (size_t)c);
    return M;
}
static void mat_to_flat(double **M, int r, int c, double *A) {
    for (int i = 0; i < r; i++) memcpy(A + (size_t)i * c, M[i], sizeof(double) * (size_t)c);
}

/* ------------------------------------------------------------------ matrixops.cpp */
/* matrixops.cpp:216-220 */
static void subtract_vec(const double *a, const double *b, double *c, int dim) {
    for (int i = 0; i < dim; i++) c[i] = a[i] - b[i];
}
/* matrixops.cpp:223-229 */
static double dotproduct_vec(const double *a, const double *b, int dim) {
    double ans = 0.0;
    for (int i = 0; i < dim; i++) ans += a[i] * b[i];
    return ans;
}
/* matrixops.cpp:26-37 */
static void vector_matrix_multiply(const double *v, double **M, int n, double *out) {
    for (int k = 0; k < n; k++) {
        double sum = 0.0;
        for (int i = 0; i < n; i++) sum += v[i] * M[i][k];
        out[k] = sum;
    }
}
/* matrixops.cpp:50-56 */
static double vector_vector_multiply(const double *a, const double *b, int n) {
    double ret = 0.0;
    for (int i = 0; i < n; i++) ret += a[i] * b[i];
    return ret;
}
/* matrixops.cpp:58-63 */
static void matrix_transpose(double **in, double **out, int n) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) out[j][i] = in[i][j];
}
/* matrixops.cpp:68-108 */
static void get_cholesky(double **in, double **out, int dim) {
    for (int i = 0; i < dim; i++)
        for (int j = 0; j < dim; j++) out[i][j] = in[i][j];
    for (int col = 0; col < dim; col++) {
        out[col][col] = sqrt(out[col][col]); /* no PD check: NaN propagates (matrixops.cpp:77) */
        for (int row = col + 1; row < dim; row++) out[row][col] = out[row][col] / out[col][col];
        for (int col2 = col + 1; col2 < dim; col2++)
            for (int row2 = col2; row2 < dim; row2++)
                out[row2][col2] = out[row2][col2] - out[row2][col] * out[col2][col];
    }
    for (int row = 0; row < dim; row++)
        for (int col = row + 1; col < dim; col++) out[row][col] = 0.0;
}
/* matrixops.cpp:113-185 */
static void multiply_and_get_logdeterminant(const double *yt, double **X, const double *y, int n,
                                            double *product_out, double *det_out) {
    double product = 0.0, det = 0.0;
    double **L = mat_new(n, n), **U = mat_new(n, n);
    get_cholesky(X, L, n);
    for (int i = 0; i < n; i++) det += log(L[i][i]);
    det = 2 * det;
    matrix_transpose(L, U, n);
    double *x = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *temp = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) { /* forward solve, matrixops.cpp:145-153 */
        temp[i] = y[i];
        for (int j = 0; j < i; j++) temp[i] -= L[i][j] * temp[j];
        temp[i] /= L[i][i];
    }
    for (int i = n - 1; i >= 0; i--) { /* backward solve, matrixops.cpp:156-164 */
        x[i] = temp[i];
        for (int j = i + 1; j < n; j++) x[i] -= U[i][j] * x[j];
        x[i] /= U[i][i];
    }
    for (int i = 0; i < n; i++) product += yt[i] * x[i];
    *product_out = product;
    *det_out = det;
    free(x);
    free(temp);
    mat_free(L, n);
    mat_free(U, n);
}
/* matrixops.cpp:264-316 */
static void vector_Kinvy_using_cholesky(double **K, const double *y, double *ans, int n) {
    double **L = mat_new(n, n), **U = mat_new(n, n);
    get_cholesky(K, L, n);
    matrix_transpose(L, U, n);
    double *temp = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        temp[i] = y[i];
        for (int j = 0; j < i; j++) temp[i] -= L[i][j] * temp[j];
        temp[i] /= L[i][i];
    }
    for (int i = n - 1; i >= 0; i--) {
        ans[i] = temp[i];
        for (int j = i + 1; j < n; j++) ans[i] -= U[i][j] * ans[j];
        ans[i] /= U[i][i];
    }
    free(temp);
    mat_free(L, n);
    mat_free(U, n);
}
/* matrixops.cpp:319-326 */
static void make_identity(double **M, int n) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) M[i][j] = (i == j) ? 1.0 : 0.0;
}
/* matrixops.cpp:330-340 */
static void matrix_forward_substitution(double **A, double **B, double **out, int dim) {
    for (int k = 0; k < dim; k++)
        for (int i = 0; i < dim; i++) {
            out[i][k] = B[i][k];
            for (int j = 0; j < i; j++) out[i][k] = out[i][k] - A[i][j] * out[j][k];
            out[i][k] = out[i][k] / A[i][i];
        }
}
/* matrixops.cpp:361-372 */
static void matrix_backward_substitution(double **A, double **B, double **out, int dim) {
    for (int k = 0; k < dim; k++)
        for (int i = dim - 1; i >= 0; i--) {
            out[i][k] = B[i][k];
            for (int j = i + 1; j < dim; j++) out[i][k] = out[i][k] - A[i][j] * out[j][k];
            out[i][k] = out[i][k] / A[i][i];
        }
}
/* matrixops.cpp:383-435 */
static void compute_K_inverse(double **K, double **outK, int n) {
    double **temp1 = mat_new(n, n), **T = mat_new(n, n), **I = mat_new(n, n), **L = mat_new(n, n);
    make_identity(I, n);
    get_cholesky(K, L, n);
    matrix_forward_substitution(L, I, T, n);
    matrix_transpose(L, temp1, n);
    matrix_backward_substitution(temp1, T, outK, n);
    mat_free(temp1, n);
    mat_free(T, n);
    mat_free(I, n);
    mat_free(L, n);
}

/* ------------------------------------------------------------------ covkernel.cpp (class Covsum) */
typedef struct {
    int n, d;
    double theta[3];
    double **tempK, **tempKinv, **tempW, **tempAlpha, **temp2, **temp3;
    double *temp1dvec, *covtempvec;
} covsum_t;

/* covkernel.cpp:13-36 (tempmatrix4 is allocated but never used by the reference; omitted) */
static covsum_t *covsum_new(int n, int d) {
    covsum_t *c = (covsum_t *)calloc(1, sizeof(covsum_t));
    c->n = n;
    c->d = d;
    c->tempK = mat_new(n, n);
    c->tempKinv = mat_new(n, n);
    c->tempW = mat_new(n, n);
    c->tempAlpha = mat_new(n, n);
    c->temp2 = mat_new(n, n);
    c->temp3 = mat_new(n, n);
    c->temp1dvec = (double *)malloc(sizeof(double) * (size_t)(n > d ? n : d) + 8);
    c->covtempvec = (double *)malloc(sizeof(double) * (size_t)(n > d ? n : d) + 8);
    return c;
}
static void covsum_free(covsum_t *c) {
    int n = c->n;
    mat_free(c->tempK, n);
    mat_free(c->tempKinv, n);
    mat_free(c->tempW, n);
    mat_free(c->tempAlpha, n);
    mat_free(c->temp2, n);
    mat_free(c->temp3, n);
    free(c->temp1dvec);
    free(c->covtempvec);
    free(c);
}
/* covkernel.cpp:64-102 */
static void covsum_K_train(covsum_t *c, double **X, double **out) {
    double ell_sq = exp(c->theta[0] * 2);
    double signal_var = exp(c->theta[1] * 2);
    double noise_var = exp(c->theta[2] * 2);
    int n = c->n;
    for (int i = 0; i < n; i++)
        for (int j = i; j < n; j++) {
            subtract_vec(X[i], X[j], c->temp1dvec, c->d);
            double val = dotproduct_vec(c->temp1dvec, c->temp1dvec, c->d);
            val = signal_var * exp(-val * 0.5 / ell_sq);
            out[i][j] = val;
            out[j][i] = val;
            if (i == j) out[i][j] += noise_var;
        }
}
/* covkernel.cpp:105-116 */
static void covsum_k_test(covsum_t *c, double **X, const double *xtest, double *out) {
    double ell_sq = exp(c->theta[0] * 2);
    double signal_var = exp(c->theta[1] * 2);
    for (int i = 0; i < c->n; i++) {
        subtract_vec(X[i], xtest, c->covtempvec, c->d);
        double val = dotproduct_vec(c->covtempvec, c->covtempvec, c->d);
        out[i] = signal_var * exp(-val * 0.5 / ell_sq);
    }
}
/* covkernel.cpp:118-129 */
static double covsum_loglik(covsum_t *c, double **X, const double *y) {
    int n = c->n;
    double quad, logdet;
    covsum_K_train(c, X, c->tempK);
    multiply_and_get_logdeterminant(y, c->tempK, y, n, &quad, &logdet); /* compute_chol_and_det */
    return -0.5 * (quad + logdet + n * 1.83787); /* truncated log(2*pi), covkernel.cpp:127 */
}
/* covkernel.cpp:130-157 */
static void covsum_squared_dist(covsum_t *c, double **X, double cc) {
    int n = c->n, d = c->d;
    for (int i = 0; i < n; i++)
        for (int j = i; j < n; j++) {
            if (i == j) {
                c->temp2[i][j] = 0.0;
                continue;
            }
            subtract_vec(X[i], X[j], c->temp1dvec, d);
            double val = dotproduct_vec(c->temp1dvec, c->temp1dvec, d) / cc;
            c->temp2[i][j] = val;
            c->temp2[j][i] = val;
        }
}
/* covkernel.cpp:162-263 */
static void covsum_grad(covsum_t *c, double **X, const double *y, double *ans) {
    int n = c->n;
    double ell_sq = exp(c->theta[0] * 2);
    double noise_var = exp(c->theta[2] * 2);
    covsum_K_train(c, X, c->tempK);
    covsum_squared_dist(c, X, ell_sq);
    for (int i = 0; i < n; i++) /* elementwise_matrixmultiply, matrixops.cpp:437-442 */
        for (int j = 0; j < n; j++) c->temp3[i][j] = c->tempK[i][j] * c->temp2[i][j];
    compute_K_inverse(c->tempK, c->tempKinv, n);
    vector_Kinvy_using_cholesky(c->tempK, y, c->temp1dvec, n);
    for (int i = 0; i < n; i++) /* get_outer_product, matrixops.cpp:250-255 */
        for (int j = 0; j < n; j++) c->tempAlpha[i][j] = c->temp1dvec[i] * c->temp1dvec[j];
    for (int i = 0; i < n; i++) /* subtract_matrices, matrixops.cpp:237-242 */
        for (int j = 0; j < n; j++) c->tempW[i][j] = c->tempKinv[i][j] - c->tempAlpha[i][j];
    double psum1 = 0.0, psum2 = 0.0, psum3 = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double curele = c->tempW[i][j];
            psum1 += curele * c->temp3[i][j];
            psum2 += curele * 2.0 * c->tempK[i][j];
            if (i == j) {
                psum3 += curele * noise_var * 2;
                psum2 -= curele * 2.0 * noise_var;
            }
        }
    ans[0] = psum1 / 2.0;
    ans[1] = psum2 / 2.0;
    ans[2] = psum3 / 2.0;
}
/* covkernel.cpp:277-306 */
static void covsum_predict(covsum_t *c, double **X, const double *y, double **Xtest, double *tmean,
                           double *tvar, int numtest) {
    int n = c->n;
    double signal_var = exp(c->theta[1] * 2);
    double noise_var = exp(c->theta[2] * 2);
    double *testK = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    double *singlevec = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    covsum_K_train(c, X, c->tempK);
    vector_Kinvy_using_cholesky(c->tempK, y, c->temp1dvec, n);
    compute_K_inverse(c->tempK, c->tempKinv, n);
    for (int i = 0; i < numtest; i++) {
        covsum_k_test(c, X, Xtest[i], testK);
        tmean[i] = vector_vector_multiply(testK, c->temp1dvec, n);
        tvar[i] = signal_var + noise_var;
        vector_matrix_multiply(testK, c->tempKinv, n, singlevec);
        double tempans = vector_vector_multiply(singlevec, testK, n);
        tvar[i] -= tempans;
    }
    free(testK);
    free(singlevec);
}

/* ------------------------------------------------------------------ BCM.cpp (class BCM) */
typedef struct {
    int K, N, D;
    int *offset, *size;
    covsum_t **experts;
    double **X; /* row pointers into the caller's data (BCM.cpp:87-88 stores, does not copy) */
    const double *y;
    double theta[3];
} bcm_t;

/* BCM.cpp:85-110: K contiguous chunks of floor(N/K) rows, the last takes the remainder. */
static bcm_t *bcm_new(double **X, const double *y, int N, int D, int K) {
    bcm_t *b = (bcm_t *)calloc(1, sizeof(bcm_t));
    b->K = K;
    b->N = N;
    b->D = D;
    b->X = X;
    b->y = y;
    b->offset = (int *)malloc(sizeof(int) * (size_t)K);
    b->size = (int *)malloc(sizeof(int) * (size_t)K);
    b->experts = (covsum_t **)malloc(sizeof(covsum_t *) * (size_t)K);
    int start = 0, partition = N / K, cursize = partition;
    for (int i = 0; i < K; i++) {
        if (i == K - 1) cursize = N - start;
        b->offset[i] = start;
        b->size[i] = cursize;
        b->experts[i] = covsum_new(cursize, D);
        start += partition;
    }
    return b;
}
static void bcm_free(bcm_t *b) {
    for (int i = 0; i < b->K; i++) covsum_free(b->experts[i]);
    free(b->experts);
    free(b->offset);
    free(b->size);
    free(b);
}
/* BCM.cpp:123-130 */
static void bcm_set_theta(bcm_t *b, const double *th) {
    for (int i = 0; i < 3; i++) b->theta[i] = th[i];
    for (int k = 0; k < b->K; k++)
        for (int i = 0; i < 3; i++) b->experts[k]->theta[i] = b->theta[i];
}
/* BCM.cpp:182-198 */
static double bcm_loglik(bcm_t *b) {
    double ans = 0.0;
    for (int k = 0; k < b->K; k++) {
        double val = covsum_loglik(b->experts[k], b->X + b->offset[k], b->y + b->offset[k]);
        ans = ans + val;
    }
    return ans;
}
/* BCM.cpp:153-180 */
static void bcm_grad(bcm_t *b, double *reqd) {
    double md[3], acc[3];
    covsum_grad(b->experts[0], b->X + b->offset[0], b->y + b->offset[0], md);
    for (int i = 0; i < 3; i++) acc[i] = md[i];
    for (int k = 1; k < b->K; k++) {
        covsum_grad(b->experts[k], b->X + b->offset[k], b->y + b->offset[k], md);
        for (int i = 0; i < 3; i++) acc[i] += md[i];
    }
    for (int i = 0; i < 3; i++) reqd[i] = acc[i];
}
/* BCM.cpp:64-83 with product_of_experts BCM.cpp:45-62 */
static void bcm_predict(bcm_t *b, double **Xtest, double *tmean, double *tvar, int size) {
    double **im = mat_new(b->K, size), **iv = mat_new(b->K, size);
    for (int k = 0; k < b->K; k++)
        covsum_predict(b->experts[k], b->X + b->offset[k], b->y + b->offset[k], Xtest, im[k], iv[k], size);
    for (int i = 0; i < size; i++) {
        double tempvar = 0.0, tempmean = 0.0;
        for (int E = 0; E < b->K; E++) {
            double invvar = 1.0 / iv[E][i];
            tempvar += invvar;
            tempmean += invvar * im[E][i];
        }
        tempvar = 1.0 / tempvar;
        tempmean = tempvar * tempmean;
        tmean[i] = tempmean;
        tvar[i] = tempvar;
    }
    mat_free(im, b->K);
    mat_free(iv, b->K);
}

/* ------------------------------------------------------------------ optimisers (covkernel.cpp:320-627) */
typedef struct {
    covsum_t *c; /* K == 0 */
    bcm_t *b;    /* K >= 1 */
    double **X;
    const double *y;
} objective_t;

static void obj_set(objective_t *o, const double *th) {
    if (o->b) bcm_set_theta(o->b, th);
    else for (int i = 0; i < 3; i++) o->c->theta[i] = th[i];
}
static double obj_ll(objective_t *o) { return o->b ? bcm_loglik(o->b) : covsum_loglik(o->c, o->X, o->y); }
static void obj_grad(objective_t *o, double *g) {
    if (o->b) bcm_grad(o->b, g);
    else covsum_grad(o->c, o->X, o->y, g);
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void axpy3(double *out, const double *x, const double *s, double a) { /* out = x + s*a */
    for (int i = 0; i < 3; i++) out[i] = x[i] + s[i] * a;
}
static double dmin(double a, double b) { return b < a ? b : a; } /* std::min */
static double dmax(double a, double b) { return a < b ? b : a; } /* std::max */

/* covkernel.cpp:388-627; identical control flow in distributed_ver1.cpp:13-232. */
static int cg_minimize(objective_t *o, double *theta, double *f_trace, int trace_cap) {
    const double INT = 0.1, EXT = 3.0, RATIO = 10, SIG = 0.1, RHO = SIG / 2;
    const int MAX = 20;
    int n = 100, nevals = 0, ntrace = 0;
    int ls_failed = 0;
    double tmp[3];

    obj_set(o, theta);
    double f0 = -1.0 * obj_ll(o);
    double df0[3], Xv[3], s[3], df3[3];
    obj_grad(o, df0);
    for (int i = 0; i < 3; i++) Xv[i] = theta[i];
    for (int i = 0; i < 3; i++) s[i] = -df0[i];
    double d0 = -dot3(s, s);
    double x3 = 1 / (1 - d0);
    double f3 = 0, d3 = 0;
    for (int i = 0; i < 3; i++) df3[i] = df0[i];
    double x2 = 0, x4 = 0, f2 = 0, f4 = 0, d2 = 0, d4 = 0;

    for (int i = 0; i < n; ++i) {
        double X0[3], dF0[3], F0 = f0;
        for (int k = 0; k < 3; k++) { X0[k] = Xv[k]; dF0[k] = df0[k]; }
        unsigned int M = (unsigned int)(MAX < (n - i) ? MAX : (n - i));
        while (1) {
            x2 = 0; f2 = f0; d2 = d0; f3 = f0;
            for (int k = 0; k < 3; k++) df3[k] = df0[k];
            int success = 0;
            while (!success && M > 0) {
                M--; i++;
                axpy3(tmp, Xv, s, x3);
                obj_set(o, tmp);
                f3 = -1.0 * obj_ll(o);
                obj_grad(o, df3);
                nevals++;
                if (f_trace && ntrace < trace_cap) f_trace[ntrace++] = f3;
                int nanFound = 0;
                for (int j = 0; j < 3; ++j) if (isnan(df3[j])) { nanFound = 1; break; }
                if (!isnan(f3) && !isinf(f3) && !nanFound) success = 1;
                else x3 = (x2 + x3) / 2;
            }
            if (f3 < F0) { axpy3(X0, Xv, s, x3); F0 = f3; for (int k = 0; k < 3; k++) dF0[k] = df3[k]; }
            d3 = dot3(df3, s);
            if ((d3 > SIG * d0) || (f3 > f0 + x3 * RHO * d0) || M == 0) break;
            double x1 = x2, f1 = f2, d1 = d2;
            x2 = x3; f2 = f3; d2 = d3;
            double A = 6 * (f1 - f2) + 3 * (d2 + d1) * (x2 - x1);
            double B = 3 * (f2 - f1) - (2 * d1 + d2) * (x2 - x1);
            x3 = x1 - d1 * (x2 - x1) * (x2 - x1) / (B + sqrt(B * B - A * d1 * (x2 - x1)));
            if (isnan(x3) || x3 < 0 || x3 > x2 * EXT) x3 = EXT * x2;
            else if (x3 < x2 + INT * (x2 - x1)) x3 = x2 + INT * (x2 - x1);
        }
        while (((fabs(d3) > -SIG * d0) || (f3 > f0 + x3 * RHO * d0)) && (M > 0)) {
            if ((d3 > 0) || (f3 > f0 + x3 * RHO * d0)) { x4 = x3; f4 = f3; d4 = d3; }
            else { x2 = x3; f2 = f3; d2 = d3; }
            if (f4 > f0) x3 = x2 - (0.5 * d2 * (x4 - x2) * (x4 - x2)) / (f4 - f2 - d2 * (x4 - x2));
            else {
                double A = 6 * (f2 - f4) / (x4 - x2) + 3 * (d4 + d2);
                double B = 3 * (f4 - f2) - (2 * d2 + d4) * (x4 - x2);
                x3 = x2 + sqrt(B * B - A * d2 * (x4 - x2) * (x4 - x2) - B) / A;
            }
            if (isnan(x3) || isinf(x3)) x3 = (x2 + x4) / 2;
            x3 = dmax(dmin(x3, x4 - INT * (x4 - x2)), x2 + INT * (x4 - x2));
            axpy3(tmp, Xv, s, x3);
            obj_set(o, tmp);
            f3 = -1.0 * obj_ll(o);
            obj_grad(o, df3);
            nevals++;
            if (f_trace && ntrace < trace_cap) f_trace[ntrace++] = f3;
            if (f3 < F0) { axpy3(X0, Xv, s, x3); F0 = f3; for (int k = 0; k < 3; k++) dF0[k] = df3[k]; }
            M--; i++;
            d3 = dot3(df3, s);
        }
        if ((fabs(d3) < -SIG * d0) && (f3 < f0 + x3 * RHO * d0)) {
            axpy3(Xv, Xv, s, x3);
            f0 = f3;
            double coef = (dot3(df3, df3) - dot3(df0, df3)) / (dot3(df0, df0));
            for (int k = 0; k < 3; k++) s[k] = coef * s[k] - df3[k];
            for (int k = 0; k < 3; k++) df0[k] = df3[k];
            d3 = d0; d0 = dot3(df0, s);
            if (d0 > 0) { for (int k = 0; k < 3; k++) s[k] = -df0[k]; d0 = -dot3(s, s); }
            x3 = x3 * dmin(RATIO, d3 / (d0 - DBL_MIN));
            ls_failed = 0;
        } else {
            for (int k = 0; k < 3; k++) { Xv[k] = X0[k]; df0[k] = dF0[k]; }
            f0 = F0;
            if (ls_failed || i >= n) break;
            for (int k = 0; k < 3; k++) s[k] = -df0[k];
            d0 = -dot3(s, s);
            x3 = 1 / (1 - d0);
            ls_failed = 1;
        }
    }
    obj_set(o, Xv);
    for (int k = 0; k < 3; k++) theta[k] = Xv[k];
    return nevals;
}

static double sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : 0.0); }

/* covkernel.cpp:320-385 */
static int rprop_minimize(objective_t *o, double *theta) {
    const double eps_stop = 0.0, Delta0 = 0.1, Deltamin = 1e-6, Deltamax = 50, etaminus = 0.5, etaplus = 1.2;
    const int n = 100;
    double Delta[3] = {Delta0, Delta0, Delta0}, grad_old[3] = {0, 0, 0}, params[3], best_params[3], grad[3];
    for (int k = 0; k < 3; k++) params[k] = best_params[k] = theta[k];
    obj_set(o, params);
    double best = -INFINITY; /* log(0) */
    int it = 0;
    for (int i = 0; i < n; ++i) {
        obj_grad(o, grad);
        it++;
        for (int j = 0; j < 3; ++j) grad_old[j] = grad_old[j] * grad[j];
        for (int j = 0; j < 3; ++j) {
            if (grad_old[j] > 0) Delta[j] = dmin(Delta[j] * etaplus, Deltamax);
            else if (grad_old[j] < 0) { Delta[j] = dmax(Delta[j] * etaminus, Deltamin); grad[j] = 0; }
            params[j] += -sign(grad[j]) * Delta[j];
        }
        for (int j = 0; j < 3; ++j) grad_old[j] = grad[j];
        if (sqrt(dot3(grad_old, grad_old)) < eps_stop) break;
        obj_set(o, params);
        double lik = obj_ll(o);
        if (lik > best) { best = lik; for (int k = 0; k < 3; k++) best_params[k] = params[k]; }
    }
    obj_set(o, best_params);
    for (int k = 0; k < 3; k++) theta[k] = best_params[k];
    return it;
}

/* ------------------------------------------------------------------ flat-array entry points */
void oracle_K_train(const double *X, int n, int d, const double *theta, double *K) {
    covsum_t *c = covsum_new(n, d);
    memcpy(c->theta, theta, sizeof(double) * 3);
    double **Xr = mat_from_flat(X, n, d);
    covsum_K_train(c, Xr, c->tempK);
    mat_to_flat(c->tempK, n, n, K);
    mat_free(Xr, n);
    covsum_free(c);
}
void oracle_k_test(const double *X, int n, int d, const double *theta, const double *xtest, double *out) {
    covsum_t *c = covsum_new(n, d);
    memcpy(c->theta, theta, sizeof(double) * 3);
    double **Xr = mat_from_flat(X, n, d);
    covsum_k_test(c, Xr, xtest, out);
    mat_free(Xr, n);
    covsum_free(c);
}
void oracle_cholesky(const double *A, int n, double *L) {
    double **Ar = mat_from_flat(A, n, n), **Lr = mat_new(n, n);
    get_cholesky(Ar, Lr, n);
    mat_to_flat(Lr, n, n, L);
    mat_free(Ar, n);
    mat_free(Lr, n);
}
void oracle_chol_and_det(const double *K, const double *y, int n, double *quad, double *logdet) {
    double **Kr = mat_from_flat(K, n, n);
    multiply_and_get_logdeterminant(y, Kr, y, n, quad, logdet);
    mat_free(Kr, n);
}
void oracle_kinv_y(const double *K, const double *y, int n, double *alpha) {
    double **Kr = mat_from_flat(K, n, n);
    vector_Kinvy_using_cholesky(Kr, y, alpha, n);
    mat_free(Kr, n);
}
void oracle_k_inverse(const double *K, int n, double *Kinv) {
    double **Kr = mat_from_flat(K, n, n), **Or = mat_new(n, n);
    compute_K_inverse(Kr, Or, n);
    mat_to_flat(Or, n, n, Kinv);
    mat_free(Kr, n);
    mat_free(Or, n);
}
double oracle_loglik(const double *X, const double *y, int n, int d, const double *theta) {
    covsum_t *c = covsum_new(n, d);
    memcpy(c->theta, theta, sizeof(double) * 3);
    double **Xr = mat_from_flat(X, n, d);
    double ll = covsum_loglik(c, Xr, y);
    mat_free(Xr, n);
    covsum_free(c);
    return ll;
}
void oracle_grad(const double *X, const double *y, int n, int d, const double *theta, double *g3) {
    covsum_t *c = covsum_new(n, d);
    memcpy(c->theta, theta, sizeof(double) * 3);
    double **Xr = mat_from_flat(X, n, d);
    covsum_grad(c, Xr, y, g3);
    mat_free(Xr, n);
    covsum_free(c);
}
void oracle_predict(const double *X, const double *y, int n, int d, const double *theta, const double *Xtest,
                    int m, double *mean, double *var) {
    covsum_t *c = covsum_new(n, d);
    memcpy(c->theta, theta, sizeof(double) * 3);
    double **Xr = mat_from_flat(X, n, d), **Xt = mat_from_flat(Xtest, m, d);
    covsum_predict(c, Xr, y, Xt, mean, var, m);
    mat_free(Xr, n);
    mat_free(Xt, m);
    covsum_free(c);
}
double oracle_nlpp(const double *actual, const double *mean, const double *var, int m) {
    double ans = 0.0;
    for (int i = 0; i < m; i++) {
        double val = 0.5 * log(6.283185 * var[i]) + pow((mean[i] - actual[i]), 2) / (2 * var[i]);
        ans += val;
    }
    return ans / m;
}
double oracle_bcm_loglik(const double *X, const double *y, int N, int D, int K, const double *theta) {
    double **Xr = mat_from_flat(X, N, D);
    bcm_t *b = bcm_new(Xr, y, N, D, K);
    bcm_set_theta(b, theta);
    double ll = bcm_loglik(b);
    bcm_free(b);
    mat_free(Xr, N);
    return ll;
}
void oracle_bcm_grad(const double *X, const double *y, int N, int D, int K, const double *theta, double *g3) {
    double **Xr = mat_from_flat(X, N, D);
    bcm_t *b = bcm_new(Xr, y, N, D, K);
    bcm_set_theta(b, theta);
    bcm_grad(b, g3);
    bcm_free(b);
    mat_free(Xr, N);
}
void oracle_bcm_predict(const double *X, const double *y, int N, int D, int K, const double *theta,
                        const double *Xtest, int m, double *mean, double *var) {
    double **Xr = mat_from_flat(X, N, D), **Xt = mat_from_flat(Xtest, m, D);
    bcm_t *b = bcm_new(Xr, y, N, D, K);
    bcm_set_theta(b, theta);
    bcm_predict(b, Xt, mean, var, m);
    bcm_free(b);
    mat_free(Xr, N);
    mat_free(Xt, m);
}
int oracle_cg_solve(const double *X, const double *y, int N, int D, int K, double *theta, double *f_trace,
                    int trace_cap) {
    double **Xr = mat_from_flat(X, N, D);
    objective_t o;
    memset(&o, 0, sizeof(o));
    o.X = Xr;
    o.y = y;
    if (K >= 1) o.b = bcm_new(Xr, y, N, D, K);
    else o.c = covsum_new(N, D);
    int ne = cg_minimize(&o, theta, f_trace, trace_cap);
    if (o.b) bcm_free(o.b);
    else covsum_free(o.c);
    mat_free(Xr, N);
    return ne;
}
int oracle_rprop_solve(const double *X, const double *y, int n, int d, double *theta) {
    double **Xr = mat_from_flat(X, n, d);
    objective_t o;
    memset(&o, 0, sizeof(o));
    o.X = Xr;
    o.y = y;
    o.c = covsum_new(n, d);
    int it = rprop_minimize(&o, theta);
    covsum_free(o.c);
    mat_free(Xr, n);
    return it;
}
