/*
 * srslte_b200 — the reference's own C API for the receive-side hot path, served by the GPU library.
 *
 * libsrslte_b200.so exports these symbols with the reference's names, argument meaning and return conventions so that
 * code written against lib/include/srsran/phy/{dft/dft.h, dft/ofdm.h, fec/turbo/turbodecoder.h, fec/turbo/rm_turbo.h,
 * fec/turbo/tc_interl.h, fec/cbsegm.h, fec/crc.h} links against it unchanged.  Struct layouts repeat the reference's
 * field order (callers embed these structs by value and read fields such as q->fft_plan.norm directly, ofdm.c:397-414);
 * the GPU state hangs off the opaque pointer slots the reference already has (dec16_hdlr[0], fft_plan.p).
 * Each declaration cites the reference interface it replaces.  This header may be included INSTEAD of the reference's
 * headers; when both are needed include the reference's and skip this one (the symbols are the same).
 *
 * Differences from the reference, all deliberate:
 *   - srsran_tdec_autoimp_get_subblocks() returns 0 for every size: the GPU decoder consumes the NATURAL 3K+12 layout, so
 *     srsran_rm_turbo_rx_lut() always produces that layout (the reference's sub-block SIMD layout is a CPU artefact).
 *   - the decoder computes what the reference's GENERIC int16 implementation computes (bit-exact), whatever dec_type is
 *     passed to srsran_tdec_init_manual.
 *   - srsran_ofdm_rx_sf() does not write the frequency-shifted samples back into the caller's input buffer.
 *   - 8-bit LLR entries and MBSFN subframes return SRSRAN_ERROR (out of scope of this path).
 */
#ifndef SRSLTE_B200_SRSRAN_API_H
#define SRSLTE_B200_SRSRAN_API_H

#include <stdbool.h>
#include <stdint.h>

#include "srslte_b200.h"

#ifdef __cplusplus
#include <complex>
extern "C" {
#endif

#ifndef SRSRAN_SUCCESS /* lib/include/srsran/config.h:57-59 */
#define SRSRAN_SUCCESS 0
#define SRSRAN_ERROR -1
#define SRSRAN_ERROR_INVALID_INPUTS -2
#endif

#ifndef SRSRAN_CONFIG_H /* cf_t, lib/include/srsran/config.h:67 */
#ifdef __cplusplus
typedef std::complex<float> cf_t;
#else
typedef _Complex float cf_t;
#endif
#endif

/* ---- code block sizes: lib/include/srsran/phy/fec/cbsegm.h ------------------------------------------------------ */
#ifndef SRSRAN_CBSEGM_H
#define SRSRAN_NOF_TC_CB_SIZES 188

typedef struct {
  uint32_t F, C, K1, K2, K1_idx, K2_idx, C1, C2, tbs, L_tb, L_cb, Z; /* cbsegm.h:32-45 */
} srsran_cbsegm_t;
#endif

SRSRAN_B200_API int  srsran_cbsegm(srsran_cbsegm_t* s, uint32_t tbs); /* cbsegm.c:62 */
SRSRAN_B200_API int  srsran_cbsegm_cbsize(uint32_t index);            /* cbsegm.c:132 */
SRSRAN_B200_API bool srsran_cbsegm_cbsize_isvalid(uint32_t size);     /* cbsegm.c:141 */
SRSRAN_B200_API int  srsran_cbsegm_cbindex(uint32_t long_cb);         /* cbsegm.c:119 */

/* ---- CRC: lib/include/srsran/phy/fec/crc.h ------------------------------------------------------------------------ */
#ifndef SRSRAN_CRC_H
typedef struct {
  uint64_t table[256];
  int      polynom;
  int      order;
  uint64_t crcinit;
  uint64_t crcmask;
  uint64_t crchighbit;
  uint32_t srsran_crc_out;
} srsran_crc_t; /* crc.h:37-45 */
#endif

SRSRAN_B200_API int      srsran_crc_init(srsran_crc_t* h, uint32_t crc_poly, int crc_order);        /* crc.c:69 */
SRSRAN_B200_API uint32_t srsran_crc_checksum_byte(srsran_crc_t* h, const uint8_t* data, int len);    /* crc.c:147 */

/* ---- QPP interleaver tables: lib/include/srsran/phy/fec/turbo/tc_interl.h ------------------------------------------ */
#ifndef SRSRAN_TC_INTERL_H
typedef struct {
  uint16_t* forward;
  uint16_t* reverse;
  uint32_t  max_long_cb;
} srsran_tc_interl_t; /* tc_interl.h:36-40 */
#endif

SRSRAN_B200_API int  srsran_tc_interl_init(srsran_tc_interl_t* h, uint32_t max_long_cb);   /* tc_interl_umts.c:51 */
SRSRAN_B200_API void srsran_tc_interl_free(srsran_tc_interl_t* h);                         /* tc_interl_umts.c:70 */
SRSRAN_B200_API int  srsran_tc_interl_LTE_gen(srsran_tc_interl_t* h, uint32_t long_cb);    /* tc_interl_lte.c:61 */
SRSRAN_B200_API int  srsran_tc_interl_LTE_gen_interl(srsran_tc_interl_t* h, uint32_t long_cb, uint32_t interl_win); /* :69 */

/* ---- rate de-matching: lib/include/srsran/phy/fec/turbo/rm_turbo.h ---------------------------------------------------- */
SRSRAN_B200_API void srsran_rm_turbo_gentables(void);   /* rm_turbo.c:276 */
SRSRAN_B200_API void srsran_rm_turbo_free_tables(void); /* rm_turbo.c:319 */
/* output[deinter[i % (3K+12)]] += input[i], natural layout (rm_turbo.c:390-445) */
SRSRAN_B200_API int srsran_rm_turbo_rx_lut(int16_t* input, int16_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx);
SRSRAN_B200_API int srsran_rm_turbo_rx_lut_(int16_t* input,
                                            int16_t* output,
                                            uint32_t in_len,
                                            uint32_t cb_idx,
                                            uint32_t rv_idx,
                                            bool     enable_input_tdec);
SRSRAN_B200_API int srsran_rm_turbo_rx_lut_8bit(int8_t* input, int8_t* output, uint32_t in_len, uint32_t cb_idx, uint32_t rv_idx);

/* ---- turbo decoder: lib/include/srsran/phy/fec/turbo/turbodecoder.h ---------------------------------------------------- */
#ifndef SRSRAN_TURBODECODER_H
#define SRSRAN_TCOD_RATE 3
#define SRSRAN_TCOD_TOTALTAIL 12
#define SRSRAN_TCOD_MAX_LEN_CB 6144

typedef enum { /* turbodecoder_impl.h:27-37 */
  SRSRAN_TDEC_AUTO = 0,
  SRSRAN_TDEC_GENERIC,
  SRSRAN_TDEC_SSE,
  SRSRAN_TDEC_SSE_WINDOW,
  SRSRAN_TDEC_NEON_WINDOW,
  SRSRAN_TDEC_AVX_WINDOW,
  SRSRAN_TDEC_SSE8_WINDOW,
  SRSRAN_TDEC_AVX8_WINDOW,
  SRSRAN_TDEC_NOF_IMP
} srsran_tdec_impl_type_t;

typedef enum { SRSRAN_TDEC_8, SRSRAN_TDEC_16 } srsran_tdec_llr_type_t;

typedef struct { /* turbodecoder.h:63-95, same field order */
  uint32_t                max_long_cb;
  void*                   dec8_hdlr[2];
  void*                   dec16_hdlr[3]; /* [0] holds the GPU decoder object */
  void*                   dec8[2];
  void*                   dec16[3];
  int                     nof_blocks8[2];
  int                     nof_blocks16[3];
  void*                   app1;
  void*                   app2;
  void*                   ext1;
  void*                   ext2;
  void*                   syst0;
  void*                   parity0;
  void*                   parity1;
  void*                   input_conv;
  bool                    force_not_sb;
  srsran_tdec_impl_type_t dec_type;
  srsran_tdec_llr_type_t  current_llr_type;
  uint32_t                current_dec;
  uint32_t                current_long_cb;
  uint32_t                current_inter_idx;
  int                     current_cbidx;
  srsran_tc_interl_t      interleaver[4][SRSRAN_NOF_TC_CB_SIZES];
  int                     n_iter;
} srsran_tdec_t;
#endif

SRSRAN_B200_API int  srsran_tdec_init(srsran_tdec_t* h, uint32_t max_long_cb);                                       /* turbodecoder.c:129 */
SRSRAN_B200_API int  srsran_tdec_init_manual(srsran_tdec_t* h, uint32_t max_long_cb, srsran_tdec_impl_type_t dec_type); /* :151 */
SRSRAN_B200_API void srsran_tdec_free(srsran_tdec_t* h);                                                             /* :319 */
SRSRAN_B200_API void srsran_tdec_force_not_sb(srsran_tdec_t* h);                                                     /* :365 */
SRSRAN_B200_API int  srsran_tdec_new_cb(srsran_tdec_t* h, uint32_t long_cb);                                         /* :510 */
SRSRAN_B200_API int  srsran_tdec_get_nof_iterations(srsran_tdec_t* h);                                               /* :579 */
SRSRAN_B200_API uint32_t srsran_tdec_autoimp_get_subblocks(uint32_t long_cb);                                        /* :381 */
SRSRAN_B200_API uint32_t srsran_tdec_autoimp_get_subblocks_8bit(uint32_t long_cb);                                   /* :410 */
/* one SISO pass + hard decision (turbodecoder.c:527-533); silently does nothing before srsran_tdec_new_cb */
SRSRAN_B200_API void srsran_tdec_iteration(srsran_tdec_t* h, int16_t* input, uint8_t* output);
/* nof_iterations passes (at least one, turbodecoder.c:536-549) + hard decision */
SRSRAN_B200_API int srsran_tdec_run_all(srsran_tdec_t* h, int16_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb);
SRSRAN_B200_API void srsran_tdec_iteration_8bit(srsran_tdec_t* h, int8_t* input, uint8_t* output); /* widened to int16, see below */
SRSRAN_B200_API int
srsran_tdec_run_all_8bit(srsran_tdec_t* h, int8_t* input, uint8_t* output, uint32_t nof_iterations, uint32_t long_cb);
/* 8-bit entries: the int8 values are widened and decoded with the generic int16 arithmetic -- the route the reference itself takes
 * for the lengths its 8-bit window decoders cannot handle (convert_8_to_16, turbodecoder.c:441-470).  Bit-exact with
 * srsran_tdec_run_all on the widened values; not with the reference's approximate saturating 8-bit decoders. */

/* ---- DFT plans: lib/include/srsran/phy/dft/dft.h ------------------------------------------------------------------------ */
#ifndef SRSRAN_DFT_H
typedef enum { SRSRAN_DFT_COMPLEX, SRSRAN_REAL } srsran_dft_mode_t;
typedef enum { SRSRAN_DFT_FORWARD, SRSRAN_DFT_BACKWARD } srsran_dft_dir_t;

typedef struct { /* dft.h:54-68, same field order */
  int               init_size;
  int               size;
  void*             in;
  void*             out;
  void*             p; /* GPU plan object */
  bool              is_guru;
  bool              forward;
  bool              mirror;
  bool              db;
  bool              norm;
  bool              dc;
  srsran_dft_dir_t  dir;
  srsran_dft_mode_t mode;
} srsran_dft_plan_t;
#endif

SRSRAN_B200_API int srsran_dft_plan_c(srsran_dft_plan_t* plan, int dft_points, srsran_dft_dir_t dir); /* dft_fftw.c:208 */
SRSRAN_B200_API int srsran_dft_plan_guru_c(srsran_dft_plan_t* plan,                                  /* dft_fftw.c:170 */
                                           int                dft_points,
                                           srsran_dft_dir_t   dir,
                                           cf_t*              in_buffer,
                                           cf_t*              out_buffer,
                                           int                istride,
                                           int                ostride,
                                           int                how_many,
                                           int                idist,
                                           int                odist);
SRSRAN_B200_API int  srsran_dft_replan(srsran_dft_plan_t* plan, const int new_dft_points);   /* dft_fftw.c:90 */
SRSRAN_B200_API int  srsran_dft_replan_c(srsran_dft_plan_t* plan, int new_dft_points);       /* dft_fftw.c:148 */
SRSRAN_B200_API void srsran_dft_plan_free(srsran_dft_plan_t* plan);                          /* dft_fftw.c:387 */
SRSRAN_B200_API void srsran_dft_plan_set_mirror(srsran_dft_plan_t* plan, bool val);          /* dft_fftw.c:280 */
SRSRAN_B200_API void srsran_dft_plan_set_db(srsran_dft_plan_t* plan, bool val);
SRSRAN_B200_API void srsran_dft_plan_set_norm(srsran_dft_plan_t* plan, bool val);
SRSRAN_B200_API void srsran_dft_plan_set_dc(srsran_dft_plan_t* plan, bool val);
SRSRAN_B200_API void srsran_dft_run_c(srsran_dft_plan_t* plan, const cf_t* in, cf_t* out);   /* dft_fftw.c:336 */
SRSRAN_B200_API void srsran_dft_run_c_zerocopy(srsran_dft_plan_t* plan, const cf_t* in, cf_t* out); /* :331 */
SRSRAN_B200_API void srsran_dft_run_guru_c(srsran_dft_plan_t* plan);                         /* dft_fftw.c:356 */

/* ---- SC-FDMA transform (de-)precoding: lib/include/srsran/phy/dft/dft_precoding.h ----------------------------------------- */
#ifndef SRSRAN_MAX_PRB
#define SRSRAN_MAX_PRB 110 /* phy_common.h */
#endif
typedef struct { /* dft_precoding.h:39-44 */
  uint32_t          max_prb;
  srsran_dft_plan_t dft_plan[SRSRAN_MAX_PRB + 1];
} srsran_dft_precoding_t;

SRSRAN_B200_API int      srsran_dft_precoding_init(srsran_dft_precoding_t* q, uint32_t max_prb, bool is_tx); /* dft_precoding.c:39 */
SRSRAN_B200_API int      srsran_dft_precoding_init_tx(srsran_dft_precoding_t* q, uint32_t max_prb);
SRSRAN_B200_API int      srsran_dft_precoding_init_rx(srsran_dft_precoding_t* q, uint32_t max_prb);
SRSRAN_B200_API void     srsran_dft_precoding_free(srsran_dft_precoding_t* q);
SRSRAN_B200_API bool     srsran_dft_precoding_valid_prb(uint32_t nof_prb);     /* 12*nof_prb = 2^a 3^b 5^c, dft_precoding.c:88-104 */
SRSRAN_B200_API uint32_t srsran_dft_precoding_get_valid_prb(uint32_t nof_prb);
/* nof_symbols transforms of 12*nof_prb points in ONE kernel launch (the reference loops srsran_dft_run_c, :120-123) */
SRSRAN_B200_API int
srsran_dft_precoding(srsran_dft_precoding_t* q, cf_t* input, cf_t* output, uint32_t nof_prb, uint32_t nof_symbols);

/* ---- OFDM receive: lib/include/srsran/phy/dft/ofdm.h ------------------------------------------------------------------------ */
#ifndef SRSRAN_OFDM_H
#ifndef SRSRAN_PHY_COMMON_H
typedef enum { SRSRAN_CP_NORM = 0, SRSRAN_CP_EXT } srsran_cp_t;   /* phy_common.h:83 */
typedef enum { SRSRAN_SF_NORM = 0, SRSRAN_SF_MBSFN } srsran_sf_t; /* phy_common.h:84 */
#endif

typedef struct { /* ofdm.h:48-63 */
  uint32_t    nof_prb;
  cf_t*       in_buffer;
  cf_t*       out_buffer;
  srsran_cp_t cp;
  srsran_sf_t sf_type;
  bool        normalize;
  float       freq_shift_f;
  float       rx_window_offset;
  uint32_t    symbol_sz;
  bool        keep_dc;
} srsran_ofdm_cfg_t;

typedef struct { /* ofdm.h:69-86 */
  srsran_ofdm_cfg_t cfg;
  srsran_dft_plan_t fft_plan; /* .p holds the GPU OFDM object; .size/.norm/.dc mirror the reference's bookkeeping */
  srsran_dft_plan_t fft_plan_sf[2];
  uint32_t          max_prb;
  uint32_t          nof_symbols;
  uint32_t          nof_guards;
  uint32_t          nof_re;
  uint32_t          slot_sz;
  uint32_t          sf_sz;
  cf_t*             tmp;
  bool              mbsfn_subframe;
  uint32_t          mbsfn_guard_len;
  uint32_t          nof_symbols_mbsfn;
  uint8_t           non_mbsfn_region;
  uint32_t          window_offset_n;
  cf_t*             shift_buffer;
  cf_t*             window_offset_buffer;
} srsran_ofdm_t;
#endif

SRSRAN_B200_API void srsran_use_standard_symbol_size(bool enabled); /* phy_common.c:322 */
SRSRAN_B200_API int  srsran_symbol_sz(uint32_t nof_prb);            /* phy_common.c:361 */

SRSRAN_B200_API int  srsran_ofdm_rx_init_cfg(srsran_ofdm_t* q, srsran_ofdm_cfg_t* cfg);                                             /* ofdm.c:290 */
SRSRAN_B200_API int  srsran_ofdm_rx_init(srsran_ofdm_t* q, srsran_cp_t cp, cf_t* in_buffer, cf_t* out_buffer, uint32_t max_prb);  /* ofdm.c:243 */
SRSRAN_B200_API int  srsran_ofdm_rx_init_mbsfn(srsran_ofdm_t* q, srsran_cp_t cp, cf_t* in_buffer, cf_t* out_buffer, uint32_t max_prb); /* SRSRAN_ERROR */
SRSRAN_B200_API int  srsran_ofdm_rx_set_prb(srsran_ofdm_t* q, srsran_cp_t cp, uint32_t nof_prb);                                    /* ofdm.c:309 */
SRSRAN_B200_API void srsran_ofdm_rx_free(srsran_ofdm_t* q);                                                                         /* ofdm.c:325 */
SRSRAN_B200_API void srsran_ofdm_rx_sf(srsran_ofdm_t* q);                                                                           /* ofdm.c:453 */
SRSRAN_B200_API void srsran_ofdm_rx_sf_ng(srsran_ofdm_t* q, cf_t* input, cf_t* output);                                             /* ofdm.c:468 */
SRSRAN_B200_API int  srsran_ofdm_set_freq_shift(srsran_ofdm_t* q, float freq_shift);                                                /* ofdm.c:334 */
SRSRAN_B200_API void srsran_ofdm_set_normalize(srsran_ofdm_t* q, bool normalize_enable);                                            /* ofdm.c:482 */
SRSRAN_B200_API void srsran_ofdm_set_non_mbsfn_region(srsran_ofdm_t* q, uint8_t non_mbsfn_region);                                  /* ofdm.c:214 */

#ifdef __cplusplus
}
#endif

#endif /* SRSLTE_B200_SRSRAN_API_H */
