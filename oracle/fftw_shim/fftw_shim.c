/*
 * TEST INFRASTRUCTURE ONLY — float64 DFT behind the FFTW3f entry points the reference calls.
 * See fftw3.h in this directory for why this exists.  Not part of the product library.
 */
#include "fftw3.h"

#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double _Complex cd_t;

struct fftwf_plan_s {
  int            n;
  int            is, os;           /* element strides inside one transform */
  int            howmany, idist, odist;
  int            sign;             /* -1 forward, +1 backward */
  int            is_r2r;
  fftwf_r2r_kind kind;
  fftwf_complex* in;
  fftwf_complex* out;
  float*         rin;
  float*         rout;
  cd_t*          tw;               /* exp(sign*2*pi*i*k/n), k<n */
  cd_t*          a;
  cd_t*          b;
};

void* fftwf_malloc(size_t n)
{
  void* p = NULL;
  if (posix_memalign(&p, 64, n ? n : 64)) {
    return NULL;
  }
  return p;
}

void fftwf_free(void* p)
{
  free(p);
}

static int smallest_factor(int n)
{
  if (n % 4 == 0) return 4;
  if (n % 2 == 0) return 2;
  for (int f = 3; f * f <= n; f += 2) {
    if (n % f == 0) return f;
  }
  return n;
}

/* Recursive decimation-in-time mixed radix: x has stride xs, result contiguous in y (length n).
 * tw is the root table of the TOP-level size ntop; tstep = ntop / n. */
static void dft_rec(const cd_t* x, int xs, cd_t* y, cd_t* scratch, int n, const cd_t* tw, int tstep, int ntop)
{
  if (n == 1) {
    y[0] = x[0];
    return;
  }
  int r = smallest_factor(n);
  int m = n / r;
  /* r sub-transforms of length m over x[q + r*j] */
  for (int q = 0; q < r; q++) {
    dft_rec(x + (size_t)q * xs, xs * r, scratch + (size_t)q * m, y, m, tw, tstep * r, ntop);
  }
  /* butterflies: y[k + m*p] = sum_q W_n^{q(k+mp)} S_q[k] */
  for (int k = 0; k < m; k++) {
    for (int p = 0; p < r; p++) {
      int  kk  = k + m * p;
      cd_t acc = 0;
      for (int q = 0; q < r; q++) {
        long idx = ((long)q * kk * tstep) % ntop;
        acc += scratch[(size_t)q * m + k] * tw[idx];
      }
      y[kk] = acc;
    }
  }
}

static fftwf_plan new_plan(int n, int sign)
{
  fftwf_plan p = calloc(1, sizeof(*p));
  p->n         = n;
  p->sign      = sign;
  p->is = p->os = 1;
  p->howmany   = 1;
  p->tw        = malloc(sizeof(cd_t) * (size_t)n);
  p->a         = malloc(sizeof(cd_t) * (size_t)n);
  p->b         = malloc(sizeof(cd_t) * (size_t)n);
  for (int k = 0; k < n; k++) {
    double ang = (double)sign * 2.0 * M_PI * (double)k / (double)n;
    p->tw[k]   = cos(ang) + I * sin(ang);
  }
  return p;
}

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned flags)
{
  (void)flags;
  fftwf_plan p = new_plan(n, sign);
  p->in        = in;
  p->out       = out;
  return p;
}

fftwf_plan fftwf_plan_guru_dft(int                rank,
                               const fftwf_iodim* dims,
                               int                howmany_rank,
                               const fftwf_iodim* howmany_dims,
                               fftwf_complex*     in,
                               fftwf_complex*     out,
                               int                sign,
                               unsigned           flags)
{
  (void)flags;
  if (rank != 1 || howmany_rank != 1) {
    return NULL;
  }
  fftwf_plan p = new_plan(dims[0].n, sign);
  p->is        = dims[0].is;
  p->os        = dims[0].os;
  p->howmany   = howmany_dims[0].n;
  p->idist     = howmany_dims[0].is;
  p->odist     = howmany_dims[0].os;
  p->in        = in;
  p->out       = out;
  return p;
}

fftwf_plan fftwf_plan_r2r_1d(int n, float* in, float* out, fftwf_r2r_kind kind, unsigned flags)
{
  (void)flags;
  fftwf_plan p = new_plan(n, kind == FFTW_R2HC ? -1 : +1);
  p->is_r2r    = 1;
  p->kind      = kind;
  p->rin       = in;
  p->rout      = out;
  return p;
}

static void run_c(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out)
{
  cd_t* x = malloc(sizeof(cd_t) * (size_t)p->n);
  for (int h = 0; h < p->howmany; h++) {
    fftwf_complex* src = in + (size_t)h * p->idist;
    fftwf_complex* dst = out + (size_t)h * p->odist;
    for (int i = 0; i < p->n; i++) {
      x[i] = (cd_t)src[(size_t)i * p->is];
    }
    dft_rec(x, 1, p->a, p->b, p->n, p->tw, 1, p->n);
    for (int i = 0; i < p->n; i++) {
      dst[(size_t)i * p->os] = (fftwf_complex)p->a[i];
    }
  }
  free(x);
}

static void run_r2r(const fftwf_plan p)
{
  int   n = p->n;
  cd_t* x = malloc(sizeof(cd_t) * (size_t)n);
  if (p->kind == FFTW_R2HC) {
    for (int i = 0; i < n; i++) x[i] = p->rin[i];
    dft_rec(x, 1, p->a, p->b, n, p->tw, 1, n);
    /* halfcomplex: r0 r1 ... r(n/2) i((n+1)/2-1) ... i1 */
    for (int k = 0; k <= n / 2; k++) p->rout[k] = (float)creal(p->a[k]);
    for (int k = 1; k < (n + 1) / 2; k++) p->rout[n - k] = (float)cimag(p->a[k]);
  } else {
    x[0] = p->rin[0];
    for (int k = 1; k < (n + 1) / 2; k++) {
      x[k]     = p->rin[k] + I * p->rin[n - k];
      x[n - k] = p->rin[k] - I * p->rin[n - k];
    }
    if (n % 2 == 0) x[n / 2] = p->rin[n / 2];
    dft_rec(x, 1, p->a, p->b, n, p->tw, 1, n);
    for (int i = 0; i < n; i++) p->rout[i] = (float)creal(p->a[i]);
  }
  free(x);
}

void fftwf_execute(const fftwf_plan p)
{
  if (!p) return;
  if (p->is_r2r) {
    run_r2r(p);
  } else {
    run_c(p, p->in, p->out);
  }
}

void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out)
{
  if (!p) return;
  run_c(p, in, out);
}

void fftwf_destroy_plan(fftwf_plan p)
{
  if (!p) return;
  free(p->tw);
  free(p->a);
  free(p->b);
  free(p);
}

int fftwf_import_wisdom_from_filename(const char* filename)
{
  (void)filename;
  return 0;
}

int fftwf_export_wisdom_to_filename(const char* filename)
{
  (void)filename;
  return 0;
}

void fftwf_cleanup(void) {}
