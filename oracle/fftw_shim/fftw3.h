/*
 * TEST INFRASTRUCTURE ONLY — minimal stand-in for <fftw3.h>.
 *
 * The reference's OFDM object (lib/src/phy/dft/ofdm.c, dft_fftw.c) calls the un-vendored FFTW3f
 * library (no version pin; found via cmake/modules/FindFFTW3F.cmake). libfftw3f is absent from this
 * image and cannot be installed, so the oracle build (oracle/Makefile) compiles the reference's
 * dft_fftw.c against this header; fftw_shim.c implements the handful of entry points it calls
 * (dft_fftw.c:62,74,76,109-110,131-133,160,188,215,245,261,333,343,359,373,396-401) with a
 * float64 mixed-radix DFT, i.e. the published definition X[k] = sum_n x[n] exp(-/+ 2 pi i k n / N),
 * unnormalised, which is what FFTW computes.  Every CP/offset/shift/copy rule stays the reference's
 * own code; only the transform is replaced.
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float _Complex fftwf_complex;

typedef struct {
  int n;
  int is;
  int os;
} fftwf_iodim;

typedef struct fftwf_plan_s* fftwf_plan;

typedef enum { FFTW_R2HC = 0, FFTW_HC2R = 1 } fftwf_r2r_kind;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

void* fftwf_malloc(size_t n);
void  fftwf_free(void* p);

fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned flags);
fftwf_plan fftwf_plan_guru_dft(int                rank,
                               const fftwf_iodim* dims,
                               int                howmany_rank,
                               const fftwf_iodim* howmany_dims,
                               fftwf_complex*     in,
                               fftwf_complex*     out,
                               int                sign,
                               unsigned           flags);
fftwf_plan fftwf_plan_r2r_1d(int n, float* in, float* out, fftwf_r2r_kind kind, unsigned flags);

void fftwf_execute(const fftwf_plan p);
void fftwf_execute_dft(const fftwf_plan p, fftwf_complex* in, fftwf_complex* out);
void fftwf_destroy_plan(fftwf_plan p);

int  fftwf_import_wisdom_from_filename(const char* filename);
int  fftwf_export_wisdom_to_filename(const char* filename);
void fftwf_cleanup(void);

#ifdef __cplusplus
}
#endif

#endif
