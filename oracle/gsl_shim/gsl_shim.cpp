/*
 * gsl_shim.cpp -- implementation of the minimal GSL-compatible subset declared in
 * gsl_shim.h.  TEST INFRASTRUCTURE ONLY: it exists so that the unmodified reference
 * (timflutre/eqtlbma v1.3.3) can be compiled here without GNU GSL (absent from this
 * image).  Written from the published algorithms (Matsumoto & Nishimura MT19937,
 * Fisher-Yates, Hestenes one-sided Jacobi SVD, Lentz continued fractions, Acklam +
 * Halley for the normal quantile), not from GSL sources.
 */
#include "gsl_shim.h"

#include <string.h>
#include <vector>
#include <algorithm>

extern "C" {

const char *gsl_version = GSL_VERSION;

/* ------------------------------------------------------------------ containers */

gsl_vector *gsl_vector_alloc(size_t n)
{
  gsl_vector *v = (gsl_vector *)malloc(sizeof(gsl_vector));
  v->size = n;
  v->stride = 1;
  v->data = (double *)malloc((n ? n : 1) * sizeof(double));
  v->block = v->data;
  v->owner = 1;
  return v;
}

gsl_vector *gsl_vector_calloc(size_t n)
{
  gsl_vector *v = gsl_vector_alloc(n);
  for (size_t i = 0; i < n; ++i) v->data[i] = 0.0;
  return v;
}

void gsl_vector_free(gsl_vector *v)
{
  if (!v) return;
  if (v->owner) free(v->data);
  free(v);
}

void gsl_vector_set_all(gsl_vector *v, double x)
{
  for (size_t i = 0; i < v->size; ++i) v->data[i * v->stride] = x;
}

int gsl_vector_memcpy(gsl_vector *dst, const gsl_vector *src)
{
  for (size_t i = 0; i < src->size; ++i) dst->data[i * dst->stride] = src->data[i * src->stride];
  return GSL_SUCCESS;
}

int gsl_vector_sub(gsl_vector *a, const gsl_vector *b)
{
  for (size_t i = 0; i < a->size; ++i) a->data[i * a->stride] -= b->data[i * b->stride];
  return GSL_SUCCESS;
}

int gsl_vector_add(gsl_vector *a, const gsl_vector *b)
{
  for (size_t i = 0; i < a->size; ++i) a->data[i * a->stride] += b->data[i * b->stride];
  return GSL_SUCCESS;
}

int gsl_vector_scale(gsl_vector *a, double x)
{
  for (size_t i = 0; i < a->size; ++i) a->data[i * a->stride] *= x;
  return GSL_SUCCESS;
}

int gsl_vector_fprintf(FILE *stream, const gsl_vector *v, const char *format)
{
  for (size_t i = 0; i < v->size; ++i) {
    fprintf(stream, format, v->data[i * v->stride]);
    fputc('\n', stream);
  }
  return GSL_SUCCESS;
}

gsl_matrix *gsl_matrix_alloc(size_t n1, size_t n2)
{
  gsl_matrix *m = (gsl_matrix *)malloc(sizeof(gsl_matrix));
  m->size1 = n1;
  m->size2 = n2;
  m->tda = n2;
  size_t n = n1 * n2;
  m->data = (double *)malloc((n ? n : 1) * sizeof(double));
  m->block = m->data;
  m->owner = 1;
  return m;
}

gsl_matrix *gsl_matrix_calloc(size_t n1, size_t n2)
{
  gsl_matrix *m = gsl_matrix_alloc(n1, n2);
  for (size_t i = 0; i < n1 * n2; ++i) m->data[i] = 0.0;
  return m;
}

void gsl_matrix_free(gsl_matrix *m)
{
  if (!m) return;
  if (m->owner) free(m->data);
  free(m);
}

void gsl_matrix_set_all(gsl_matrix *m, double x)
{
  for (size_t i = 0; i < m->size1; ++i)
    for (size_t j = 0; j < m->size2; ++j) m->data[i * m->tda + j] = x;
}

void gsl_matrix_set_identity(gsl_matrix *m)
{
  for (size_t i = 0; i < m->size1; ++i)
    for (size_t j = 0; j < m->size2; ++j) m->data[i * m->tda + j] = (i == j) ? 1.0 : 0.0;
}

int gsl_matrix_memcpy(gsl_matrix *dst, const gsl_matrix *src)
{
  for (size_t i = 0; i < src->size1; ++i)
    for (size_t j = 0; j < src->size2; ++j) dst->data[i * dst->tda + j] = src->data[i * src->tda + j];
  return GSL_SUCCESS;
}

int gsl_matrix_add(gsl_matrix *a, const gsl_matrix *b)
{
  for (size_t i = 0; i < a->size1; ++i)
    for (size_t j = 0; j < a->size2; ++j) a->data[i * a->tda + j] += b->data[i * b->tda + j];
  return GSL_SUCCESS;
}

int gsl_matrix_sub(gsl_matrix *a, const gsl_matrix *b)
{
  for (size_t i = 0; i < a->size1; ++i)
    for (size_t j = 0; j < a->size2; ++j) a->data[i * a->tda + j] -= b->data[i * b->tda + j];
  return GSL_SUCCESS;
}

int gsl_matrix_scale(gsl_matrix *a, double x)
{
  for (size_t i = 0; i < a->size1; ++i)
    for (size_t j = 0; j < a->size2; ++j) a->data[i * a->tda + j] *= x;
  return GSL_SUCCESS;
}

int gsl_matrix_mul_elements(gsl_matrix *a, const gsl_matrix *b)
{
  for (size_t i = 0; i < a->size1; ++i)
    for (size_t j = 0; j < a->size2; ++j) a->data[i * a->tda + j] *= b->data[i * b->tda + j];
  return GSL_SUCCESS;
}

int gsl_matrix_get_col(gsl_vector *v, const gsl_matrix *m, size_t j)
{
  for (size_t i = 0; i < m->size1; ++i) v->data[i * v->stride] = m->data[i * m->tda + j];
  return GSL_SUCCESS;
}

int gsl_matrix_set_col(gsl_matrix *m, size_t j, const gsl_vector *v)
{
  for (size_t i = 0; i < m->size1; ++i) m->data[i * m->tda + j] = v->data[i * v->stride];
  return GSL_SUCCESS;
}

_gsl_vector_view gsl_matrix_diagonal(gsl_matrix *m)
{
  _gsl_vector_view view;
  view.vector.size = std::min(m->size1, m->size2);
  view.vector.stride = m->tda + 1;
  view.vector.data = m->data;
  view.vector.block = m->block;
  view.vector.owner = 0;
  return view;
}

_gsl_vector_const_view gsl_matrix_const_diagonal(const gsl_matrix *m)
{
  _gsl_vector_const_view view;
  view.vector.size = std::min(m->size1, m->size2);
  view.vector.stride = m->tda + 1;
  view.vector.data = m->data;
  view.vector.block = m->block;
  view.vector.owner = 0;
  return view;
}

/* ------------------------------------------------------------------ BLAS */

int gsl_blas_dgemm(CBLAS_TRANSPOSE_t TransA, CBLAS_TRANSPOSE_t TransB, double alpha,
                   const gsl_matrix *A, const gsl_matrix *B, double beta, gsl_matrix *C)
{
  const size_t M = C->size1, N = C->size2;
  const bool ta = (TransA != CblasNoTrans), tb = (TransB != CblasNoTrans);
  const size_t K = ta ? A->size1 : A->size2;
  const size_t MA = ta ? A->size2 : A->size1;
  const size_t KB = tb ? B->size2 : B->size1;
  const size_t NB = tb ? B->size1 : B->size2;
  if (MA != M || KB != K || NB != N) {
    fprintf(stderr, "gsl_shim: dgemm size mismatch\n");
    abort();
  }
  for (size_t i = 0; i < M; ++i) {
    for (size_t j = 0; j < N; ++j) {
      double acc = 0.0;
      for (size_t k = 0; k < K; ++k) {
        const double a = ta ? A->data[k * A->tda + i] : A->data[i * A->tda + k];
        const double b = tb ? B->data[j * B->tda + k] : B->data[k * B->tda + j];
        acc += a * b;
      }
      double *c = &C->data[i * C->tda + j];
      *c = (beta == 0.0 ? 0.0 : beta * (*c)) + alpha * acc;
    }
  }
  return GSL_SUCCESS;
}

int gsl_blas_dgemv(CBLAS_TRANSPOSE_t TransA, double alpha, const gsl_matrix *A,
                   const gsl_vector *X, double beta, gsl_vector *Y)
{
  const bool ta = (TransA != CblasNoTrans);
  const size_t M = ta ? A->size2 : A->size1, K = ta ? A->size1 : A->size2;
  for (size_t i = 0; i < M; ++i) {
    double acc = 0.0;
    for (size_t k = 0; k < K; ++k) {
      const double a = ta ? A->data[k * A->tda + i] : A->data[i * A->tda + k];
      acc += a * X->data[k * X->stride];
    }
    double *y = &Y->data[i * Y->stride];
    *y = (beta == 0.0 ? 0.0 : beta * (*y)) + alpha * acc;
  }
  return GSL_SUCCESS;
}

int gsl_blas_ddot(const gsl_vector *X, const gsl_vector *Y, double *result)
{
  double acc = 0.0;
  for (size_t i = 0; i < X->size; ++i) acc += X->data[i * X->stride] * Y->data[i * Y->stride];
  *result = acc;
  return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ permutation / combination */

gsl_permutation *gsl_permutation_alloc(size_t n)
{
  gsl_permutation *p = (gsl_permutation *)malloc(sizeof(gsl_permutation));
  p->size = n;
  p->data = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
  return p;
}

gsl_permutation *gsl_permutation_calloc(size_t n)
{
  gsl_permutation *p = gsl_permutation_alloc(n);
  for (size_t i = 0; i < n; ++i) p->data[i] = i; /* identity */
  return p;
}

void gsl_permutation_free(gsl_permutation *p)
{
  if (!p) return;
  free(p->data);
  free(p);
}

gsl_combination *gsl_combination_calloc(size_t n, size_t k)
{
  if (k > n) return NULL;
  gsl_combination *c = (gsl_combination *)malloc(sizeof(gsl_combination));
  c->n = n;
  c->k = k;
  c->data = (size_t *)malloc((k ? k : 1) * sizeof(size_t));
  for (size_t i = 0; i < k; ++i) c->data[i] = i; /* lexicographically first */
  return c;
}

void gsl_combination_free(gsl_combination *c)
{
  if (!c) return;
  free(c->data);
  free(c);
}

/* lexicographic successor of a k-subset of {0..n-1} */
int gsl_combination_next(gsl_combination *c)
{
  const size_t n = c->n, k = c->k;
  size_t *d = c->data;
  if (k == 0) return GSL_FAILURE;
  size_t i = k - 1;
  while (i > 0 && d[i] == n - k + i) --i;
  if (i == 0 && d[i] == n - k) return GSL_FAILURE;
  ++d[i];
  for (; i < k - 1; ++i) d[i + 1] = d[i] + 1;
  return GSL_SUCCESS;
}

/* ------------------------------------------------------------------ linalg */

/* One-sided (Hestenes) Jacobi SVD: A (M x N, M >= N) -> U in A, V (N x N), S sorted
 * in decreasing order.  High relative accuracy; column signs are arbitrary (every
 * call site in the reference is sign-invariant). */
int gsl_linalg_SV_decomp(gsl_matrix *A, gsl_matrix *V, gsl_vector *S, gsl_vector *work)
{
  (void)work;
  const size_t M = A->size1, N = A->size2;
  gsl_matrix_set_identity(V);
  const double tol = 1e-15;
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (size_t p = 0; p + 1 < N; ++p) {
      for (size_t q = p + 1; q < N; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (size_t i = 0; i < M; ++i) {
          const double ap = A->data[i * A->tda + p], aq = A->data[i * A->tda + q];
          alpha += ap * ap;
          beta += aq * aq;
          gamma += ap * aq;
        }
        if (gamma == 0.0 || fabs(gamma) <= tol * sqrt(alpha * beta)) continue;
        rotated = 1;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (size_t i = 0; i < M; ++i) {
          const double ap = A->data[i * A->tda + p], aq = A->data[i * A->tda + q];
          A->data[i * A->tda + p] = c * ap - s * aq;
          A->data[i * A->tda + q] = s * ap + c * aq;
        }
        for (size_t i = 0; i < N; ++i) {
          const double vp = V->data[i * V->tda + p], vq = V->data[i * V->tda + q];
          V->data[i * V->tda + p] = c * vp - s * vq;
          V->data[i * V->tda + q] = s * vp + c * vq;
        }
      }
    }
    if (!rotated) break;
  }
  std::vector<double> sv(N);
  for (size_t j = 0; j < N; ++j) {
    double nrm = 0.0;
    for (size_t i = 0; i < M; ++i) nrm += A->data[i * A->tda + j] * A->data[i * A->tda + j];
    sv[j] = sqrt(nrm);
  }
  /* a column whose norm is at rounding level relative to the largest is a null direction */
  double smax = 0.0;
  for (size_t j = 0; j < N; ++j) smax = std::max(smax, sv[j]);
  std::vector<size_t> ord(N);
  for (size_t j = 0; j < N; ++j) ord[j] = j;
  std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return sv[a] > sv[b]; });
  std::vector<double> Ucopy(M * N), Vcopy(N * N);
  for (size_t i = 0; i < M; ++i)
    for (size_t j = 0; j < N; ++j) Ucopy[i * N + j] = A->data[i * A->tda + j];
  for (size_t i = 0; i < N; ++i)
    for (size_t j = 0; j < N; ++j) Vcopy[i * N + j] = V->data[i * V->tda + j];
  for (size_t jj = 0; jj < N; ++jj) {
    const size_t j = ord[jj];
    const double s = sv[j];
    S->data[jj * S->stride] = s;
    for (size_t i = 0; i < M; ++i)
      A->data[i * A->tda + jj] = (s > 0.0) ? Ucopy[i * N + j] / s : 0.0;
    for (size_t i = 0; i < N; ++i) V->data[i * V->tda + jj] = Vcopy[i * N + j];
  }
  return GSL_SUCCESS;
}

/* LU with partial pivoting (row interchanges), PA = LU, unit lower L stored below the diagonal */
int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum)
{
  const size_t N = A->size1;
  *signum = 1;
  for (size_t i = 0; i < N; ++i) p->data[i] = i;
  for (size_t j = 0; j + 1 < N; ++j) {
    double maxv = fabs(A->data[j * A->tda + j]);
    size_t ipiv = j;
    for (size_t i = j + 1; i < N; ++i) {
      const double a = fabs(A->data[i * A->tda + j]);
      if (a > maxv) {
        maxv = a;
        ipiv = i;
      }
    }
    if (ipiv != j) {
      for (size_t k = 0; k < N; ++k) std::swap(A->data[j * A->tda + k], A->data[ipiv * A->tda + k]);
      std::swap(p->data[j], p->data[ipiv]);
      *signum = -(*signum);
    }
    const double ajj = A->data[j * A->tda + j];
    if (ajj != 0.0) {
      for (size_t i = j + 1; i < N; ++i) {
        const double aij = A->data[i * A->tda + j] / ajj;
        A->data[i * A->tda + j] = aij;
        for (size_t k = j + 1; k < N; ++k) A->data[i * A->tda + k] -= aij * A->data[j * A->tda + k];
      }
    }
  }
  return GSL_SUCCESS;
}

int gsl_linalg_LU_invert(const gsl_matrix *LU, const gsl_permutation *p, gsl_matrix *inverse)
{
  const size_t N = LU->size1;
  std::vector<double> x(N);
  for (size_t col = 0; col < N; ++col) {
    /* solve L U x = P e_col */
    for (size_t i = 0; i < N; ++i) x[i] = (p->data[i] == col) ? 1.0 : 0.0;
    for (size_t i = 0; i < N; ++i) {
      double acc = x[i];
      for (size_t k = 0; k < i; ++k) acc -= LU->data[i * LU->tda + k] * x[k];
      x[i] = acc;
    }
    for (size_t ii = N; ii-- > 0;) {
      double acc = x[ii];
      for (size_t k = ii + 1; k < N; ++k) acc -= LU->data[ii * LU->tda + k] * x[k];
      x[ii] = acc / LU->data[ii * LU->tda + ii];
    }
    for (size_t i = 0; i < N; ++i) inverse->data[i * inverse->tda + col] = x[i];
  }
  return GSL_SUCCESS;
}

double gsl_linalg_LU_lndet(gsl_matrix *LU)
{
  double lndet = 0.0;
  for (size_t i = 0; i < LU->size1; ++i) lndet += log(fabs(LU->data[i * LU->tda + i]));
  return lndet;
}

/* ------------------------------------------------------------------ multifit (GSL >= 2.3 semantics:
 * column balancing by powers of two, SVD, components with s_j <= DBL_EPSILON*s_0 dropped,
 * rss = ||y - U U^T y||^2, cov = rss/(n-rank) * (V S^-1)(V S^-1)^T / (D_i D_j)) */

gsl_multifit_linear_workspace *gsl_multifit_linear_alloc(size_t n, size_t p)
{
  gsl_multifit_linear_workspace *w =
      (gsl_multifit_linear_workspace *)malloc(sizeof(gsl_multifit_linear_workspace));
  w->nmax = w->n = n;
  w->pmax = w->p = p;
  w->A = gsl_matrix_alloc(n, p);
  w->Q = gsl_matrix_alloc(p, p);
  w->QSI = gsl_matrix_alloc(p, p);
  w->S = gsl_vector_alloc(p);
  w->t = gsl_vector_alloc(n);
  w->xt = gsl_vector_calloc(p);
  w->D = gsl_vector_calloc(p);
  w->rcond = 0.0;
  return w;
}

void gsl_multifit_linear_free(gsl_multifit_linear_workspace *w)
{
  if (!w) return;
  gsl_matrix_free(w->A);
  gsl_matrix_free(w->Q);
  gsl_matrix_free(w->QSI);
  gsl_vector_free(w->S);
  gsl_vector_free(w->t);
  gsl_vector_free(w->xt);
  gsl_vector_free(w->D);
  free(w);
}

static void balance_columns(gsl_matrix *A, gsl_vector *D)
{
  const size_t M = A->size1, N = A->size2;
  for (size_t j = 0; j < N; ++j) {
    double s = 0.0;
    for (size_t i = 0; i < M; ++i) s += fabs(A->data[i * A->tda + j]);
    double f = 1.0;
    if (s == 0.0 || !(s <= DBL_MAX)) {
      D->data[j] = f;
      continue;
    }
    while (s > 1.0) {
      s /= 2.0;
      f *= 2.0;
    }
    while (s < 0.5) {
      s *= 2.0;
      f /= 2.0;
    }
    D->data[j] = f;
    if (f != 1.0)
      for (size_t i = 0; i < M; ++i) A->data[i * A->tda + j] /= f;
  }
}

static int multifit_core(const gsl_matrix *X, const gsl_vector *w, const gsl_vector *y,
                         gsl_vector *c, gsl_matrix *cov, double *chisq,
                         gsl_multifit_linear_workspace *work)
{
  const size_t n = X->size1, p = X->size2;
  work->n = n;
  work->p = p;
  gsl_matrix *A = work->A;
  std::vector<double> yw(n);
  for (size_t i = 0; i < n; ++i) {
    const double sw = w ? sqrt(std::max(0.0, w->data[i * w->stride])) : 1.0;
    for (size_t j = 0; j < p; ++j) A->data[i * A->tda + j] = sw * X->data[i * X->tda + j];
    yw[i] = sw * y->data[i * y->stride];
  }
  balance_columns(A, work->D);
  gsl_linalg_SV_decomp(A, work->Q, work->S, NULL);
  const double s0 = work->S->data[0];
  work->rcond = (s0 > 0.0) ? work->S->data[p - 1] / s0 : 0.0;
  /* xt = U^T y */
  for (size_t j = 0; j < p; ++j) {
    double acc = 0.0;
    for (size_t i = 0; i < n; ++i) acc += A->data[i * A->tda + j] * yw[i];
    work->xt->data[j] = acc;
  }
  double rho2 = 0.0;
  if (n > p) {
    for (size_t i = 0; i < n; ++i) {
      double fit = 0.0;
      for (size_t j = 0; j < p; ++j) fit += A->data[i * A->tda + j] * work->xt->data[j];
      const double r = yw[i] - fit;
      rho2 += r * r;
    }
  }
  size_t rank = 0;
  for (size_t j = 0; j < p; ++j) {
    const double sj = work->S->data[j];
    double alpha = 0.0;
    if (!(sj <= GSL_DBL_EPSILON * s0)) {
      alpha = 1.0 / sj;
      ++rank;
    }
    for (size_t i = 0; i < p; ++i)
      work->QSI->data[i * p + j] = work->Q->data[i * work->Q->tda + j] * alpha;
  }
  for (size_t i = 0; i < p; ++i) {
    double acc = 0.0;
    for (size_t j = 0; j < p; ++j) acc += work->QSI->data[i * p + j] * work->xt->data[j];
    c->data[i * c->stride] = acc / work->D->data[i];
  }
  *chisq = rho2;
  const double s2 = rho2 / (double)(n - rank);
  for (size_t i = 0; i < p; ++i) {
    for (size_t j = i; j < p; ++j) {
      double s = 0.0;
      for (size_t k = 0; k < p; ++k) s += work->QSI->data[i * p + k] * work->QSI->data[j * p + k];
      const double v = s * s2 / (work->D->data[i] * work->D->data[j]);
      cov->data[i * cov->tda + j] = v;
      cov->data[j * cov->tda + i] = v;
    }
  }
  return GSL_SUCCESS;
}

int gsl_multifit_linear(const gsl_matrix *X, const gsl_vector *y, gsl_vector *c,
                        gsl_matrix *cov, double *chisq, gsl_multifit_linear_workspace *work)
{
  return multifit_core(X, NULL, y, c, cov, chisq, work);
}

int gsl_multifit_wlinear(const gsl_matrix *X, const gsl_vector *w, const gsl_vector *y,
                         gsl_vector *c, gsl_matrix *cov, double *chisq,
                         gsl_multifit_linear_workspace *work)
{
  /* weighted fit: the reference's IRLS path (--lik poisson, out of scope) only links it */
  int st = multifit_core(X, w, y, c, cov, chisq, work);
  /* GSL's wlinear covariance is (X^T W X)^-1 without the s2 factor */
  const size_t n = X->size1, p = X->size2;
  size_t rank = gsl_multifit_linear_rank(GSL_DBL_EPSILON, work);
  const double s2 = *chisq / (double)(n - rank);
  if (s2 > 0.0)
    for (size_t i = 0; i < p; ++i)
      for (size_t j = 0; j < p; ++j) cov->data[i * cov->tda + j] /= s2;
  return st;
}

size_t gsl_multifit_linear_rank(double tol, const gsl_multifit_linear_workspace *work)
{
  const double s0 = work->S->data[0];
  size_t rank = 0;
  for (size_t j = 0; j < work->p; ++j)
    if (work->S->data[j] > tol * s0) ++rank;
  return rank;
}

/* ------------------------------------------------------------------ stats / sort */

double gsl_stats_mean(const double data[], size_t stride, size_t n)
{
  /* running-mean recurrence */
  long double mean = 0;
  for (size_t i = 0; i < n; ++i) mean += (data[i * stride] - mean) / (i + 1);
  return (double)mean;
}

double gsl_stats_tss(const double data[], size_t stride, size_t n)
{
  const double mean = gsl_stats_mean(data, stride, n);
  long double tss = 0;
  for (size_t i = 0; i < n; ++i) {
    const long double delta = data[i * stride] - mean;
    tss += delta * delta;
  }
  return (double)tss;
}

static inline void index_downheap(size_t *p, const double *data, size_t stride, size_t N, size_t k)
{
  const size_t pki = p[k];
  while (k <= N / 2) {
    size_t j = 2 * k;
    if (j < N && data[p[j] * stride] < data[p[j + 1] * stride]) j++;
    if (!(data[pki * stride] < data[p[j] * stride])) break;
    p[k] = p[j];
    k = j;
  }
  p[k] = pki;
}

/* index heapsort (not stable): the order of tied values is the heap's, as in GSL */
void gsl_sort_index(size_t *p, const double *data, size_t stride, size_t n)
{
  if (n == 0) return;
  for (size_t i = 0; i < n; ++i) p[i] = i;
  size_t N = n - 1;
  size_t k = N / 2;
  k++;
  do {
    k--;
    index_downheap(p, data, stride, N, k);
  } while (k > 0);
  while (N > 0) {
    size_t tmp = p[0];
    p[0] = p[N];
    p[N] = tmp;
    N--;
    index_downheap(p, data, stride, N, 0);
  }
}

/* ------------------------------------------------------------------ cdf / sf */

double gsl_cdf_ugaussian_P(double x) { return 0.5 * erfc(-x * M_SQRT1_2); }
double gsl_cdf_ugaussian_Q(double x) { return 0.5 * erfc(x * M_SQRT1_2); }
double gsl_cdf_gaussian_P(double x, double sigma) { return gsl_cdf_ugaussian_P(x / sigma); }

/* Lower-tail standard normal quantile: Acklam's rational start + Halley refinement on
 * the relative residual (stable down to the smallest normal double). */
double gsl_cdf_ugaussian_Pinv(double P)
{
  if (P != P) return NAN;
  if (P <= 0.0) return (P == 0.0) ? -HUGE_VAL : NAN;
  if (P >= 1.0) return (P == 1.0) ? HUGE_VAL : NAN;
  if (P > 0.5) {
    /* use symmetry on the complementary probability when it is exactly representable */
    const double q = 1.0 - P;
    if (q > 0.0 && (1.0 - q) == P) return -gsl_cdf_ugaussian_Pinv(q);
  }
  static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                              1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                              6.680131188771972e+01,  -1.328068155288572e+01};
  static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                              -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
  static const double d[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                              3.754408661907416e+00};
  const double plow = 0.02425, phigh = 1.0 - plow;
  double x;
  if (P < plow) {
    const double q = sqrt(-2.0 * log(P));
    x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
  } else if (P <= phigh) {
    const double q = P - 0.5, r = q * q;
    x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
        (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
  } else {
    const double q = sqrt(-2.0 * log1p(-P));
    x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
        ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
  }
  for (int it = 0; it < 4; ++it) {
    double u; /* u = (Phi(x) - P) / phi(x) */
    if (x < -5.0) {
      /* work with the relative residual and the Mills ratio to avoid overflow of exp(x^2/2) */
      const double Phi = 0.5 * erfc(-x * M_SQRT1_2);
      const double rel = (Phi - P) / Phi;
      const double x2 = x * x;
      /* asymptotic Mills ratio Phi(x)/phi(x) for x << 0 */
      const double mills = (-1.0 / x) * (1.0 - 1.0 / x2 + 3.0 / (x2 * x2) - 15.0 / (x2 * x2 * x2) +
                                         105.0 / (x2 * x2 * x2 * x2));
      u = rel * mills;
    } else {
      const double e = 0.5 * erfc(-x * M_SQRT1_2) - P;
      u = e * sqrt(2.0 * M_PI) * exp(0.5 * x * x);
    }
    const double dx = u / (1.0 + 0.5 * x * u);
    x -= dx;
    if (fabs(dx) <= 1e-16 * fabs(x)) break;
  }
  return x;
}

double gsl_cdf_gaussian_Pinv(double P, double sigma) { return sigma * gsl_cdf_ugaussian_Pinv(P); }

/* continued fraction of the incomplete beta function (modified Lentz) */
static double beta_cf(double a, double b, double x)
{
  const double tiny = 1e-300, eps = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 100000; ++m) {
    const int m2 = 2 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

/* Stirling correction lnGamma(z) - [(z-1/2) ln z - z + ln(2 pi)/2], z >= 10 */
static double lgam_corr(double z)
{
  const double z2 = z * z;
  return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0) / z2) / z2) / z2) / z2) / z;
}

/* ln[Gamma(a+b) / (Gamma(a) Gamma(b))] without the cancellation of three large lgamma values */
static double ln_inv_beta(double a, double b)
{
  if (a < b) std::swap(a, b);
  if (b >= 10.0) {
    return a * log1p(b / a) + b * log1p(a / b) + 0.5 * (log(a) + log(b) - log(a + b)) -
           0.5 * log(2.0 * M_PI) + lgam_corr(a + b) - lgam_corr(a) - lgam_corr(b);
  }
  if (a >= 10.0) {
    return (a - 0.5) * log1p(b / a) + b * log(a + b) - b + lgam_corr(a + b) - lgam_corr(a) - lgamma(b);
  }
  return lgamma(a + b) - lgamma(a) - lgamma(b);
}

/* regularized incomplete beta I_x(a,b) and its complement, given x and y = 1-x separately
 * (so neither tail suffers cancellation); logx/logy may be supplied for extra accuracy */
static void beta_inc_pair(double a, double b, double x, double y, double logx, double logy,
                          double *I, double *Ic)
{
  if (x <= 0.0) {
    *I = 0.0;
    *Ic = 1.0;
    return;
  }
  if (y <= 0.0) {
    *I = 1.0;
    *Ic = 0.0;
    return;
  }
  const double lnpre = ln_inv_beta(a, b) + a * logx + b * logy;
  const double bt = exp(lnpre);
  if (x < (a + 1.0) / (a + b + 2.0)) {
    *I = bt * beta_cf(a, b, x) / a;
    *Ic = 1.0 - *I;
  } else {
    *Ic = bt * beta_cf(b, a, y) / b;
    *I = 1.0 - *Ic;
  }
}

double gsl_shim_beta_inc(double a, double b, double x)
{
  double I, Ic;
  beta_inc_pair(a, b, x, 1.0 - x, log(x), log1p(-x), &I, &Ic);
  return I;
}

/* two-sided tail of Student's t: Pr(|T| > |t|) = I_{nu/(nu+t^2)}(nu/2, 1/2), and its complement */
static void tdist_tails(double t, double nu, double *tail, double *central)
{
  const double t2 = t * t;
  if (t2 == HUGE_VAL) {
    *tail = 0.0;
    *central = 1.0;
    return;
  }
  const double x = nu / (nu + t2), y = t2 / (nu + t2);
  const double logx = -log1p(t2 / nu);
  const double logy = (t2 > 0.0) ? -log1p(nu / t2) : -HUGE_VAL;
  if (t2 == 0.0) {
    *tail = 1.0;
    *central = 0.0;
    return;
  }
  beta_inc_pair(0.5 * nu, 0.5, x, y, logx, logy, tail, central);
}

double gsl_cdf_tdist_P(double x, double nu)
{
  if (x != x || nu != nu) return NAN;
  double tail, central;
  tdist_tails(x, nu, &tail, &central);
  return (x < 0.0) ? 0.5 * tail : 0.5 + 0.5 * central;
}

double gsl_cdf_tdist_Q(double x, double nu)
{
  if (x != x || nu != nu) return NAN;
  double tail, central;
  tdist_tails(x, nu, &tail, &central);
  return (x > 0.0) ? 0.5 * tail : 0.5 + 0.5 * central;
}

/* upper tail of F(nu1, nu2): I_{nu2/(nu2+nu1 x)}(nu2/2, nu1/2) */
double gsl_cdf_fdist_Q(double x, double nu1, double nu2)
{
  if (x != x) return NAN;
  if (x <= 0.0) return 1.0;
  const double r = nu1 * x / nu2;
  const double bx = 1.0 / (1.0 + r), by = r / (1.0 + r);
  double I, Ic;
  beta_inc_pair(0.5 * nu2, 0.5 * nu1, bx, by, -log1p(r), -log1p(1.0 / r), &I, &Ic);
  return I;
}

/* regularized upper incomplete gamma Q(a, x); if lnQ != NULL also its logarithm */
static double gamma_inc_Q_ln(double a, double x, double *lnQ);
static double gamma_inc_Q(double a, double x) { return gamma_inc_Q_ln(a, x, NULL); }
static double gamma_inc_Q_ln(double a, double x, double *lnQ)
{
  if (x <= 0.0) {
    if (lnQ) *lnQ = 0.0;
    return 1.0;
  }
  const double lg = lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    const double q = 1.0 - sum * exp(-x + a * log(x) - lg);
    if (lnQ) *lnQ = log(q);
    return q;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; ++i) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < 1e-16) break;
  }
  if (lnQ) *lnQ = -x + a * log(x) - lg + log(h);
  return exp(-x + a * log(x) - lg) * h;
}

double gsl_cdf_chisq_Q(double x, double nu) { return gamma_inc_Q(0.5 * nu, 0.5 * x); }

double gsl_cdf_chisq_Qinv(double Q, double nu)
{
  if (Q != Q) return NAN;
  if (Q >= 1.0) return 0.0;
  if (Q <= 0.0) return HUGE_VAL;
  if (nu == 1.0) {
    /* chi2_1 = Z^2 : upper tail Q  <=>  |Z| > z with Phi(-z) = Q/2 */
    const double z = gsl_cdf_ugaussian_Pinv(0.5 * Q);
    return z * z;
  }
  /* Wilson-Hilferty start, then safeguarded Newton on log Q */
  const double z = -gsl_cdf_ugaussian_Pinv(Q);
  const double h = 2.0 / (9.0 * nu);
  double x = nu * pow(std::max(1e-3, 1.0 - h + z * sqrt(h)), 3.0);
  const double a = 0.5 * nu, lg = lgamma(a);
  for (int it = 0; it < 200; ++it) {
    double lnq;
    gamma_inc_Q_ln(a, 0.5 * x, &lnq);
    const double lnpdf = log(0.5) - 0.5 * x + (a - 1.0) * log(0.5 * x) - lg;
    /* Newton on f(x) = log q(x) - log Q, f' = -pdf/q */
    double dx = (lnq - log(Q)) * exp(lnq - lnpdf);
    if (x + dx <= 0.0) dx = -0.5 * x;
    x += dx;
    if (fabs(dx) <= 1e-15 * x) break;
  }
  return x;
}

double gsl_sf_choose(unsigned int n, unsigned int m)
{
  if (m > n) return NAN;
  if (m == n || m == 0) return 1.0;
  if (2 * m > n) m = n - m;
  /* exact in double while the result fits in 2^53 (covers every S the path can enumerate) */
  long double r = 1.0L;
  for (unsigned int i = 1; i <= m; ++i) r = r * (long double)(n - m + i) / (long double)i;
  return (double)floorl(r + 0.5L);
}

/* ------------------------------------------------------------------ rng: MT19937 (2002 seeding) */

typedef struct {
  unsigned long mt[624];
  int mti;
} mt_state_t;

static const gsl_rng_type mt19937_type = {"mt19937", 0xffffffffUL, 0, sizeof(mt_state_t)};
const gsl_rng_type *gsl_rng_mt19937 = &mt19937_type;
const gsl_rng_type *gsl_rng_default = &mt19937_type;
unsigned long int gsl_rng_default_seed = 0;

const gsl_rng_type *gsl_rng_env_setup(void)
{
  const char *p = getenv("GSL_RNG_TYPE");
  if (p) {
    if (strcmp(p, "mt19937") != 0) {
      fprintf(stderr, "gsl_shim: only GSL_RNG_TYPE=mt19937 is supported (got %s)\n", p);
      exit(EXIT_FAILURE);
    }
    fprintf(stderr, "GSL_RNG_TYPE=%s\n", p);
  }
  gsl_rng_default = &mt19937_type;
  unsigned long seed = 0;
  p = getenv("GSL_RNG_SEED");
  if (p) {
    seed = strtoul(p, 0, 0);
    fprintf(stderr, "GSL_RNG_SEED=%lu\n", seed);
  }
  gsl_rng_default_seed = seed;
  return gsl_rng_default;
}

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
  gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
  r->type = T;
  r->state = calloc(1, T->size);
  gsl_rng_set(r, gsl_rng_default_seed);
  return r;
}

void gsl_rng_free(gsl_rng *r)
{
  if (!r) return;
  free(r->state);
  free(r);
}

void gsl_rng_set(const gsl_rng *r, unsigned long int s)
{
  mt_state_t *st = (mt_state_t *)r->state;
  if (s == 0) s = 4357; /* the default seed of the reference implementation */
  st->mt[0] = s & 0xffffffffUL;
  for (int i = 1; i < 624; ++i)
    st->mt[i] = (1812433253UL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
  st->mti = 624;
}

unsigned long int gsl_rng_get(const gsl_rng *r)
{
  mt_state_t *st = (mt_state_t *)r->state;
  unsigned long *mt = st->mt;
  const unsigned long UPPER = 0x80000000UL, LOWER = 0x7fffffffUL;
  if (st->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; ++kk) {
      unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    for (; kk < 623; ++kk) {
      unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    {
      unsigned long y = (mt[623] & UPPER) | (mt[0] & LOWER);
      mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
    }
    st->mti = 0;
  }
  unsigned long k = mt[st->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680UL;
  k ^= (k << 15) & 0xefc60000UL;
  k ^= (k >> 18);
  return k & 0xffffffffUL;
}

double gsl_rng_uniform(const gsl_rng *r) { return gsl_rng_get(r) / 4294967296.0; }

unsigned long int gsl_rng_uniform_int(const gsl_rng *r, unsigned long int n)
{
  const unsigned long offset = r->type->min;
  const unsigned long range = r->type->max - offset;
  if (n > range || n == 0) {
    fprintf(stderr, "gsl_shim: invalid n for gsl_rng_uniform_int\n");
    abort();
  }
  const unsigned long scale = range / n;
  unsigned long k;
  do {
    k = (gsl_rng_get(r) - offset) / scale;
  } while (k >= n);
  return k;
}

void gsl_ran_shuffle(const gsl_rng *r, void *base, size_t n, size_t size)
{
  char *b = (char *)base;
  std::vector<char> tmp(size);
  for (size_t i = n - 1; i > 0 && n > 0; --i) {
    const size_t j = gsl_rng_uniform_int(r, i + 1);
    if (i != j) {
      memcpy(tmp.data(), b + i * size, size);
      memcpy(b + i * size, b + j * size, size);
      memcpy(b + j * size, tmp.data(), size);
    }
  }
}

double gsl_ran_flat(const gsl_rng *r, double a, double b)
{
  const double u = gsl_rng_uniform(r);
  return a * (1.0 - u) + b * u;
}

/* GSL 2.x randist/exponential.c: -mu * log1p(-u), u from gsl_rng_uniform (eqtlbma_hm --rand, eqtlbma_hm.cpp:545-568) */
double gsl_ran_exponential(const gsl_rng *r, double mu)
{
  const double u = gsl_rng_uniform(r);
  return -mu * log1p(-u);
}

} /* extern "C" */
