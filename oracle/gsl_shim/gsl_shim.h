/*
 * gsl_shim.h -- minimal GSL-compatible API subset (TEST INFRASTRUCTURE ONLY).
 *
 * Purpose: GNU GSL is an external dependency of the reference (timflutre/eqtlbma
 * v1.3.3) that is NOT present in this image and cannot be installed (no network).
 * This shim re-implements, from the published algorithms, exactly the symbol
 * subset the reference's eqtlbma_bf uses, so that the
 * UNMODIFIED reference sources under /root/reference/src compile into
 * oracle/_ref/eqtlbma_bf_ref (see oracle/Makefile).  Integer paths (MT19937,
 * gsl_rng_uniform_int, gsl_ran_shuffle, gsl_combination_next, gsl_sort_index)
 * follow the documented GSL algorithms bit-for-bit; floating paths
 * (multifit_linear, SV_decomp, LU, cdfs) are accurate to ~1e-14 relative.
 *
 * Nothing in the product path (eqtlbma_b200/) includes or links this file.
 */
#ifndef EQTLBMA_ORACLE_GSL_SHIM_H
#define EQTLBMA_ORACLE_GSL_SHIM_H

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <float.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- version (reference requires <=1.16 or >=2.3: utils_math.cpp:185-196) ---- */
#define GSL_VERSION "2.7-shim"
#define GSL_MAJOR_VERSION 2
#define GSL_MINOR_VERSION 7
extern const char *gsl_version;

#define GSL_SUCCESS 0
#define GSL_FAILURE (-1)
#define GSL_DBL_EPSILON 2.2204460492503131e-16
#define GSL_DBL_MIN 2.2250738585072014e-308
#define GSL_DBL_MAX 1.7976931348623157e+308
#define GSL_POSINF (HUGE_VAL)
#define GSL_NEGINF (-HUGE_VAL)
#define GSL_NAN (NAN)

/* ---- containers ---- */
typedef struct {
  size_t size;
  size_t stride;
  double *data;
  void *block;
  int owner;
} gsl_vector;

typedef struct {
  size_t size1;
  size_t size2;
  size_t tda;
  double *data;
  void *block;
  int owner;
} gsl_matrix;

typedef struct { gsl_vector vector; } _gsl_vector_view;
typedef _gsl_vector_view gsl_vector_view;
typedef struct { gsl_vector vector; } _gsl_vector_const_view;
typedef const _gsl_vector_const_view gsl_vector_const_view;

gsl_vector *gsl_vector_alloc(size_t n);
gsl_vector *gsl_vector_calloc(size_t n);
void gsl_vector_free(gsl_vector *v);
static inline double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector *v, size_t i, double x) { v->data[i * v->stride] = x; }
void gsl_vector_set_all(gsl_vector *v, double x);
int gsl_vector_memcpy(gsl_vector *dst, const gsl_vector *src);
int gsl_vector_sub(gsl_vector *a, const gsl_vector *b);
int gsl_vector_add(gsl_vector *a, const gsl_vector *b);
int gsl_vector_scale(gsl_vector *a, double x);
int gsl_vector_fprintf(FILE *stream, const gsl_vector *v, const char *format);

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2);
gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix *m);
static inline double gsl_matrix_get(const gsl_matrix *m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
static inline void gsl_matrix_set(gsl_matrix *m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
void gsl_matrix_set_all(gsl_matrix *m, double x);
void gsl_matrix_set_identity(gsl_matrix *m);
int gsl_matrix_memcpy(gsl_matrix *dst, const gsl_matrix *src);
int gsl_matrix_add(gsl_matrix *a, const gsl_matrix *b);
int gsl_matrix_sub(gsl_matrix *a, const gsl_matrix *b);
int gsl_matrix_scale(gsl_matrix *a, double x);
int gsl_matrix_mul_elements(gsl_matrix *a, const gsl_matrix *b);
int gsl_matrix_get_col(gsl_vector *v, const gsl_matrix *m, size_t j);
int gsl_matrix_set_col(gsl_matrix *m, size_t j, const gsl_vector *v);
_gsl_vector_view gsl_matrix_diagonal(gsl_matrix *m);
_gsl_vector_const_view gsl_matrix_const_diagonal(const gsl_matrix *m);

/* ---- BLAS ---- */
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE_t;
int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                   const gsl_matrix *A, const gsl_matrix *B, double beta, gsl_matrix *C);
int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A,
                   const gsl_vector *X, double beta, gsl_vector *Y);
int gsl_blas_ddot(const gsl_vector *X, const gsl_vector *Y, double *result);

/* ---- permutation / combination ---- */
typedef struct { size_t size; size_t *data; } gsl_permutation;
gsl_permutation *gsl_permutation_alloc(size_t n);
gsl_permutation *gsl_permutation_calloc(size_t n);
void gsl_permutation_free(gsl_permutation *p);
static inline size_t gsl_permutation_get(const gsl_permutation *p, size_t i) { return p->data[i]; }

typedef struct { size_t n; size_t k; size_t *data; } gsl_combination;
gsl_combination *gsl_combination_calloc(size_t n, size_t k);
void gsl_combination_free(gsl_combination *c);
int gsl_combination_next(gsl_combination *c);
static inline size_t gsl_combination_get(const gsl_combination *c, size_t i) { return c->data[i]; }

/* ---- linalg ---- */
int gsl_linalg_SV_decomp(gsl_matrix *A, gsl_matrix *V, gsl_vector *S, gsl_vector *work);
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum);
int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse);
double gsl_linalg_LU_lndet(gsl_matrix *LU);

/* ---- multifit ---- */
typedef struct {
  size_t nmax, pmax, n, p;
  gsl_matrix *A;   /* balanced X, then U (n x p) */
  gsl_matrix *Q;   /* V (p x p) */
  gsl_matrix *QSI; /* V S^-1 */
  gsl_vector *S, *t, *xt, *D;
  double rcond;
} gsl_multifit_linear_workspace;
gsl_multifit_linear_workspace *gsl_multifit_linear_alloc(size_t n, size_t p);
void gsl_multifit_linear_free(gsl_multifit_linear_workspace *w);
int gsl_multifit_linear(const gsl_matrix *X, const gsl_vector *y, gsl_vector *c,
                        gsl_matrix *cov, double *chisq, gsl_multifit_linear_workspace *work);
int gsl_multifit_wlinear(const gsl_matrix *X, const gsl_vector *w, const gsl_vector *y,
                         gsl_vector *c, gsl_matrix *cov, double *chisq,
                         gsl_multifit_linear_workspace *work);
size_t gsl_multifit_linear_rank(double tol, const gsl_multifit_linear_workspace *work);

/* ---- stats / sort ---- */
double gsl_stats_mean(const double data[], size_t stride, size_t n);
double gsl_stats_tss(const double data[], size_t stride, size_t n);
void gsl_sort_index(size_t *p, const double *data, size_t stride, size_t n);

/* ---- cdf / sf ---- */
double gsl_cdf_ugaussian_P(double x);
double gsl_cdf_ugaussian_Q(double x);
double gsl_cdf_ugaussian_Pinv(double P);
double gsl_cdf_gaussian_P(double x, double sigma);
double gsl_cdf_gaussian_Pinv(double P, double sigma);
double gsl_cdf_tdist_P(double x, double nu);
double gsl_cdf_tdist_Q(double x, double nu);
double gsl_cdf_fdist_Q(double x, double nu1, double nu2);
double gsl_cdf_chisq_Q(double x, double nu);
double gsl_cdf_chisq_Qinv(double Q, double nu);
double gsl_sf_choose(unsigned int n, unsigned int m);
/* exposed for the shim's own accuracy tests */
double gsl_shim_beta_inc(double a, double b, double x);

/* ---- rng ---- */
typedef struct {
  const char *name;
  unsigned long int max;
  unsigned long int min;
  size_t size;
} gsl_rng_type;
typedef struct {
  const gsl_rng_type *type;
  void *state;
} gsl_rng;
extern const gsl_rng_type *gsl_rng_mt19937;
extern const gsl_rng_type *gsl_rng_default;
extern unsigned long int gsl_rng_default_seed;
const gsl_rng_type *gsl_rng_env_setup(void);
gsl_rng *gsl_rng_alloc(const gsl_rng_type *T);
void gsl_rng_free(gsl_rng *r);
void gsl_rng_set(const gsl_rng *r, unsigned long int seed);
unsigned long int gsl_rng_get(const gsl_rng *r);
double gsl_rng_uniform(const gsl_rng *r);
unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n);
void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t nmembm, size_t size);
double gsl_ran_flat(const gsl_rng *r, double a, double b);
double gsl_ran_exponential(const gsl_rng *r, double mu);

#ifdef __cplusplus
}
#endif

#endif
