/*
 * srslte_b200 — C ABI of the B200-native receive-side PHY hot path.
 *
 * Two groups of entry points live in libsrslte_b200.so:
 *
 *  1. The reference's own per-object API for this path (srsran_ofdm_*, srsran_dft_*, srsran_rm_turbo_*,
 *     srsran_tdec_*, ...), declared in srslte_b200_srsran_api.h with the reference's names, argument meaning and
 *     error behaviour, so existing callers (sch.c, enb_ul.c, ue_dl.c and the reference's unit tests) link unchanged.
 *
 *  2. The batched entries declared HERE.  They are what a maintainer binds to move the hot loops of
 *     lib/src/phy/phch/sch.c:370-492 (decode_tb_cb) and lib/src/phy/enb/enb_ul.c:151-154 (srsran_enb_ul_fft) onto the
 *     GPU a whole batch of code blocks / subframes at a time.  See INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns SRSRAN_SUCCESS (0), SRSRAN_ERROR (-1) or
 * SRSRAN_ERROR_INVALID_INPUTS (-2) like lib/include/srsran/config.h:57-64 and prints a one-line reason on stderr.
 * There is NO CPU fallback: without a usable CUDA device every entry fails with SRSRAN_ERROR.
 * "pass" below means one SISO half-iteration, exactly what the reference calls an iteration
 * (turbodecoder_iter.h:104-140): expert.pusch_max_its = 8 is 8 passes = 4 full turbo iterations.
 */
#ifndef SRSLTE_B200_H
#define SRSLTE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRSRAN_B200_API __attribute__((visibility("default")))

/* flags */
#define SRSRAN_B200_FLAG_DEVICE_PTRS 0x1u    /* data pointers are device memory on the object's GPU (else host memory) */
#define SRSRAN_B200_FLAG_SOFT_ON_DEVICE 0x2u /* only the HARQ soft-buffer pool is device memory (stays resident) */
#define SRSRAN_B200_FLAG_LLR_INT8 0x4u       /* srsran_b200_tdec_run: `llr` points to int8_t values (the reference's 8-bit soft-bit
                                                container, rm_turbo.h:80 / demod_soft.h srsran_demod_soft_demodulate_b); they are
                                                widened to int16 on the device and decoded with the SAME int16 arithmetic, so the
                                                result equals the int16 entry on the widened values.  Halves the PCIe bytes. */

/* int16 values per code block in a soft-buffer pool: SOFTBUFFER_SIZE of lib/include/srsran/phy/fec/softbuffer.h:56 */
#define SRSRAN_B200_SOFTBUFFER_SIZE 18600

/* CRC the per-pass early-stop check uses (sch.c:437-444) */
#define SRSRAN_B200_CRC_NONE 0
#define SRSRAN_B200_CRC24A 1 /* single-code-block transport block: CRC24A over the K = tbs+24 bits */
#define SRSRAN_B200_CRC24B 2 /* segmented transport block: CRC24B over the K bits of each code block */

#define SRSRAN_B200_FLAG_IQ_INT16 0x8u       /* srsran_b200_ofdm_rx_sf_batch: `in` holds int16 I/Q pairs (the radio's wire format, what
                                                the RF front ends convert with srsran_vec_convert_if(.., 32768, ..) on the host): the
                                                first FFT pass converts (x / 32768, exact in float) so the result equals the float
                                                entry on the converted samples.  Halves the PCIe bytes per subframe. */

SRSRAN_B200_API int srsran_b200_device_count(void);

/* Number of CUDA kernels this library has launched so far in this process (all objects, all devices). */
SRSRAN_B200_API uint64_t srsran_b200_kernel_launches(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched turbo decoding.  Replaces the loop  srsran_tdec_new_cb(); do { srsran_tdec_iteration(); crc } while (..)
 * of sch.c:420-454 for many code blocks at once, bit-exact with the reference's generic int16 decoder
 * (lib/src/phy/fec/turbo/turbodecoder_gen.c).
 */
typedef struct srsran_b200_tdec srsran_b200_tdec_t; /* opaque; one per host thread, not thread-safe (like srsran_tdec_t) */

/* max_cb_hint: expected batch size, used to pre-size device memory (it still grows on demand). */
SRSRAN_B200_API int  srsran_b200_tdec_init(srsran_b200_tdec_t** h, int device, uint32_t max_cb_hint);
SRSRAN_B200_API void srsran_b200_tdec_free(srsran_b200_tdec_t* h);

/*
 * Decode ncb code blocks of equal length K.
 *   llr        [ncb][3K+12] int16, the natural decoder input layout of turbodecoder_gen.c:238-258
 *              (what srsran_rm_turbo_rx_lut_(.., enable_input_tdec=false) leaves in the soft buffer)
 *   max_passes upper bound on SISO passes per block (q->max_iterations, sch.c:454); >= 1
 *   crc_kind   SRSRAN_B200_CRC_*; the syndrome is evaluated after every pass
 *   early_stop non-zero: a block stops at its first CRC match (sch.c:446-449); zero: every block runs max_passes
 *   out        [ncb][K/8] decided bits of each block's LAST pass, MSB first (tdec_decision_byte)
 *   crc_ok     [ncb] 1 if the CRC matched after some pass (may be NULL)
 *   npass      [ncb] pass count at the first CRC match, else the passes spent — cb_noi of sch.c:431 (may be NULL)
 *   stream     cudaStream_t to enqueue on when SRSRAN_B200_FLAG_DEVICE_PTRS is set (NULL = default stream); the
 *              call then returns without synchronising.  With host pointers the call is synchronous.
 */
SRSRAN_B200_API int srsran_b200_tdec_run(srsran_b200_tdec_t* h,
                                         const int16_t*      llr,
                                         uint32_t            ncb,
                                         uint32_t            K,
                                         uint32_t            max_passes,
                                         int                 crc_kind,
                                         int                 early_stop,
                                         uint8_t*            out,
                                         uint8_t*            crc_ok,
                                         uint8_t*            npass,
                                         uint32_t            flags,
                                         void*               stream);

/*
 * BASELINE config 3: code blocks of SEVERAL lengths in one batch (turbodecoder.c:510-525 re-arms the reference's decoder per
 * block length; here every tile of 64 blocks carries its own K and one launch per pass covers all of them, longest first).
 *   llr     group after group: ncb[0] vectors of 3*K[0]+12 int16, then ncb[1] vectors of 3*K[1]+12, ...
 *   out     group after group: K[g]/8 bytes per block;  crc_ok, npass: one entry per block in the same order
 *   K, ncb  host arrays of n_groups entries; every K must be one of the 188 sizes (else SRSRAN_ERROR like srsran_tdec_new_cb)
 * Other arguments and the host / device pointer rule as srsran_b200_tdec_run (no int8 container here).
 */
SRSRAN_B200_API int srsran_b200_tdec_run_mixed(srsran_b200_tdec_t* h,
                                               const int16_t*      llr,
                                               uint32_t            n_groups,
                                               const uint32_t*     K,
                                               const uint32_t*     ncb,
                                               uint32_t            max_passes,
                                               int                 crc_kind,
                                               int                 early_stop,
                                               uint8_t*            out,
                                               uint8_t*            crc_ok,
                                               uint8_t*            npass,
                                               uint32_t            flags,
                                               void*               stream);

/*
 * Optional kernel timing: with enable != 0 every kernel the object launches is bracketed by CUDA events on the
 * launching stream; profile_get synchronises the device and returns accumulated milliseconds and launch counts for
 * the three kernel classes [0] layout load, [1] SISO pass, [2] decision/pack.  Reset clears the history.
 */
SRSRAN_B200_API void srsran_b200_tdec_profile_reset(srsran_b200_tdec_t* h, int enable);
SRSRAN_B200_API int  srsran_b200_tdec_profile_get(srsran_b200_tdec_t* h, double* ms_by_class, uint64_t* launches_by_class);
/* Same with nclasses entries: [3] = re-packing of the still-running blocks between the passes of an early-stop decode. */
/* Every timed span since the last reset in launch order: ms[i] and its class cls[i]; returns the number written (<= max_spans). */
SRSRAN_B200_API int  srsran_b200_tdec_profile_spans(srsran_b200_tdec_t* h, float* ms, int* cls, int max_spans);
SRSRAN_B200_API int  srsran_b200_tdec_profile_get_ex(srsran_b200_tdec_t* h, double* ms_by_class, uint64_t* launches_by_class, int nclasses);
/* Resident CTAs (tiles of 64 code blocks) per SM of the SISO pass kernel on the current device, as the CUDA occupancy
 * calculator reports it; -1 on error.  Diagnostic: a 65,536-block batch is one wave when this is >= 7 on 148 SMs. */
SRSRAN_B200_API int  srsran_b200_tdec_resident_tiles_per_sm(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Shared-channel receive processing: batched rate de-matching and the transport-block decode loop.
 *
 * srsran_b200_sch_t plays the role of srsran_sch_t (lib/include/srsran/phy/phch/sch.h:60-85) for the receive side:
 * it owns the decoder and the de-matching tables.  One per host thread.
 */
typedef struct srsran_b200_sch srsran_b200_sch_t;

SRSRAN_B200_API int  srsran_b200_sch_init(srsran_b200_sch_t** q, int device);
SRSRAN_B200_API void srsran_b200_sch_free(srsran_b200_sch_t* q);
/* srsran_sch_set_max_noi (sch.c:222-229): maximum SISO passes per code block, 0 selects the default of 10 */
SRSRAN_B200_API void srsran_b200_sch_set_max_noi(srsran_b200_sch_t* q, uint32_t max_iterations);
/* The NEXT srsran_b200_sch_decode_batch call reads device buffers that are still being produced on `producer_stream`
 * (e.g. the soft bits of srsran_b200_pusch_rx_batch): its kernels are ordered after everything queued on that stream at
 * the time of the call, so the caller need not synchronise first and the call's host-side bookkeeping overlaps the producer. */
SRSRAN_B200_API void srsran_b200_sch_decode_after(srsran_b200_sch_t* q, void* producer_stream);
/* The same with a cudaEvent_t the caller has recorded behind the producing work: the next call waits for that event only,
 * not for what was queued on the producer stream after it (a later chunk of a pipelined batch). */
SRSRAN_B200_API void srsran_b200_sch_decode_after_event(srsran_b200_sch_t* q, void* event);

/* One srsran_rm_turbo_rx_lut_(input, output, in_len, cb_idx, rv_idx, enable_input_tdec=false) call (rm_turbo.c:403) */
typedef struct {
  uint32_t cb_idx;      /* index into the 188 code block sizes (srsran_cbsegm_cbindex) */
  uint32_t rv;          /* redundancy version 0..3 */
  uint32_t E;           /* in_len: received soft bits of this code block */
  uint32_t new_data;    /* non-zero: the soft buffer is taken as all-zero before combining (fresh HARQ process) */
  uint64_t in_offset;   /* first soft bit inside e_bits, int16 units */
  uint64_t soft_offset; /* this block's 3K+12 soft buffer inside soft_pool, int16 units, natural layout */
} srsran_b200_rm_cb_t;

/*
 * soft_pool[soft_offset + table[i mod (3K+12)]] += e_bits[in_offset + i]  for i < E, int16 wrap-around, for n code
 * blocks at once.  With SRSRAN_B200_FLAG_DEVICE_PTRS both buffers are device memory and the work is enqueued on
 * `stream`; otherwise host memory, synchronous.  Invalid rv / cb_idx returns SRSRAN_ERROR_INVALID_INPUTS (rm_turbo.c:442).
 */
SRSRAN_B200_API int srsran_b200_rm_turbo_rx_batch(srsran_b200_sch_t*         q,
                                                  const int16_t*             e_bits,
                                                  uint64_t                   e_len,
                                                  int16_t*                   soft_pool,
                                                  uint64_t                   soft_len,
                                                  const srsran_b200_rm_cb_t* cbs,
                                                  uint32_t                   n,
                                                  uint32_t                   flags,
                                                  void*                      stream);

/* One decode_tb() call (sch.c:507-572) */
typedef struct {
  /* in */
  uint32_t tbs;         /* transport block size in bits (cb_segm->tbs) */
  uint32_t Qm;          /* bits per modulation symbol (x layers) */
  uint32_t rv;
  uint32_t nof_e_bits;  /* G: soft bits of this transport block */
  uint64_t e_offset;    /* first soft bit inside e_bits, int16 units */
  uint64_t soft_offset; /* soft buffers of this HARQ process inside soft_pool: code block c at
                           soft_offset + c * SRSRAN_B200_SOFTBUFFER_SIZE (softbuffer->buffer_f[c]) */
  uint64_t data_offset; /* first output byte inside data; tbs/8 + 3 bytes are produced, keep 768 bytes of slack */
  uint32_t new_data;    /* non-zero: srsran_softbuffer_rx_reset_tbs semantics, soft buffers start from zero */
  /* in/out */
  uint32_t cb_crc_mask; /* bit c: code block c is already decoded (in: skipped, its bytes must still be in data,
                           sch.c:390,466-471; out: updated with the blocks that passed now) */
  /* out */
  int32_t  result;         /* SRSRAN_SUCCESS: every code block and the TB CRC24A matched; SRSRAN_ERROR: CRC failure;
                              SRSRAN_ERROR_INVALID_INPUTS: filler bits / too many blocks / buffer overrun */
  uint32_t nof_cb;         /* C */
  float    avg_iterations; /* q->avg_iterations of sch.c:490 */
} srsran_b200_tb_t;

/*
 * Decodes n_tb transport blocks: per code block rate de-matching into the soft pool, up to max_noi SISO passes with
 * a CRC check after each (CRC24B per block, CRC24A when C == 1), payload assembly, TB CRC.  Synchronous.
 * flags: 0 = e_bits/soft_pool/data are host memory; SRSRAN_B200_FLAG_SOFT_ON_DEVICE = the soft pool is device
 * memory; SRSRAN_B200_FLAG_DEVICE_PTRS = all three are device memory.
 */
SRSRAN_B200_API int srsran_b200_sch_decode_batch(srsran_b200_sch_t* q,
                                                 const int16_t*     e_bits,
                                                 uint64_t           e_len,
                                                 int16_t*           soft_pool,
                                                 uint64_t           soft_len,
                                                 uint8_t*           data,
                                                 uint64_t           data_len,
                                                 srsran_b200_tb_t*  tbs,
                                                 uint32_t           n_tb,
                                                 uint32_t           flags);

/* The same call in two halves, for a caller thread that keeps several batches in flight (one object per batch): _begin plans
 * and queues the whole batch on the object's stream and returns; _finish waits for it and fills tbs[] (which, like the
 * buffers, must stay valid and untouched in between).  Device buffers only (SRSRAN_B200_FLAG_DEVICE_PTRS).  One batch per
 * object at a time; srsran_b200_sch_decode_batch is _begin followed by _finish. */
SRSRAN_B200_API int srsran_b200_sch_decode_begin(srsran_b200_sch_t* q, const int16_t* e_bits, uint64_t e_len, int16_t* soft_pool,
                                                 uint64_t soft_len, uint8_t* data, uint64_t data_len, srsran_b200_tb_t* tbs, uint32_t n_tb,
                                                 uint32_t flags);
SRSRAN_B200_API int srsran_b200_sch_decode_finish(srsran_b200_sch_t* q);

/* ---------------------------------------------------------------------------------------------------------------
 * Batched OFDM demodulation: srsran_ofdm_rx_sf (lib/src/phy/dft/ofdm.c:453-466) for many subframes per call.
 * The configuration mirrors srsran_ofdm_cfg_t (lib/include/srsran/phy/dft/ofdm.h:48-63) minus the buffer bindings.
 */
typedef struct srsran_b200_ofdm srsran_b200_ofdm_t;

typedef struct {
  uint32_t nof_prb;          /* number of resource blocks: 12*nof_prb elements are kept per symbol */
  int      cp_ext;           /* 0: normal cyclic prefix (7 symbols/slot), 1: extended (6) */
  uint32_t symbol_sz;        /* DFT size; 0 derives it from nof_prb like srsran_symbol_sz (phy_common.c:361-385) */
  float    freq_shift_f;     /* frequency shift in subcarriers applied before the DFT (UL: -0.5), 0 = none */
  float    rx_window_offset; /* fraction (0..1) of the CP the DFT window is advanced by, 0 = none */
  int      normalize;        /* non-zero: scale the output by 1/sqrt(symbol_sz) */
  int      keep_dc;          /* non-zero: keep the DC bin (it is dropped only when there is no frequency shift) */
} srsran_b200_ofdm_cfg_t;

/* srsran_use_standard_symbol_size / srsran_symbol_sz (phy_common.c:322,361): process-wide choice of the size table */
SRSRAN_B200_API void srsran_b200_use_standard_symbol_size(int enabled);
SRSRAN_B200_API int  srsran_b200_symbol_sz(uint32_t nof_prb);

SRSRAN_B200_API int  srsran_b200_ofdm_rx_init(srsran_b200_ofdm_t** q, int device, const srsran_b200_ofdm_cfg_t* cfg);
SRSRAN_B200_API int  srsran_b200_ofdm_rx_reconfigure(srsran_b200_ofdm_t* q, const srsran_b200_ofdm_cfg_t* cfg);
SRSRAN_B200_API void srsran_b200_ofdm_rx_free(srsran_b200_ofdm_t* q);
SRSRAN_B200_API int  srsran_b200_ofdm_rx_geometry(const srsran_b200_ofdm_t* q,
                                                  uint32_t*                 symbol_sz,
                                                  uint32_t*                 sf_sz,
                                                  uint32_t*                 nof_symbols,
                                                  uint32_t*                 nof_re);
/*
 * in : nsf subframes of sf_sz = 15*symbol_sz complex float samples (cf_t, interleaved re/im); NOT modified (the
 *      reference multiplies the frequency shift into its input buffer in place, ofdm.c:455-457)
 * out: nsf * nof_symbols * nof_re complex floats, symbol major, FFT-shifted, guards and (when applicable) DC removed
 * Host pointers: synchronous.  SRSRAN_B200_FLAG_DEVICE_PTRS: enqueued on `stream`, returns immediately.
 */
SRSRAN_B200_API int srsran_b200_ofdm_rx_sf_batch(srsran_b200_ofdm_t* q, const void* in, void* out, uint32_t nsf, uint32_t flags,
                                                 void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * int16 soft demapping: srsran_demod_soft_demodulate_s (lib/src/phy/modem/demod_soft.c:871-894) for QPSK (1), 16QAM (2)
 * and 64QAM (3) -- srsran_mod_t numbering -- bit-exact with the reference's x86 (SSE/AVX2) build, whose vector body
 * rounds to nearest and whose scalar tail truncates.  symbols_per_call is the nsymbols of each reference call this
 * batch stands for (the body/tail split is per call); 0 means one call of nsymbols.
 * symbols: nsymbols complex floats;  llr: nsymbols * {2,4,6} int16.
 */
SRSRAN_B200_API int srsran_b200_demod_soft_demodulate_s(int         device,
                                                        int         modulation,
                                                        const void* symbols,
                                                        int16_t*    llr,
                                                        uint32_t    nsymbols,
                                                        uint32_t    symbols_per_call,
                                                        uint32_t    flags,
                                                        void*       stream);

/*
 * PUSCH glue between srsran_b200_ofdm_rx_sf_batch and srsran_b200_sch_decode_batch (device buffers only, flags must carry
 * SRSRAN_B200_FLAG_DEVICE_PTRS): per subframe, the OFDM symbols whose bit is set in sym_mask (bit l = symbol l; clear
 * the DMRS symbols 3 and 10 of a normal-CP PUSCH subframe: 0x3BF7) are soft-demapped in order as one
 * srsran_demod_soft_demodulate_s(modulation, d, q, nof_re_total) call (lib/src/phy/phch/pusch.c:449) and every soft
 * bit is then shifted right arithmetically by llr_shift bits (0 = the reference's scale; 4 brings 64QAM's 700x scale
 * into the overflow-free envelope of the generic int16 decoder).  Channel estimation, equalisation, transform
 * de-precoding, descrambling and UL-SCH de-interleaving are NOT part of this entry (identity channel pipeline).
 * grid: nsf x nof_symbols x nof_re cf_t;  llr: nsf x popcount(sym_mask) x nof_re x {2,4,6} int16.
 */
SRSRAN_B200_API int srsran_b200_pusch_demap_batch(int         device,
                                                  int         modulation,
                                                  const void* grid,
                                                  int16_t*    llr,
                                                  uint32_t    nsf,
                                                  uint32_t    nof_symbols,
                                                  uint32_t    nof_re,
                                                  uint32_t    sym_mask,
                                                  uint32_t    llr_shift,
                                                  uint32_t    flags,
                                                  void*       stream);

/* ---------------------------------------------------------------------------------------------------------------
 * PUSCH receive chain between the OFDM demodulator and the rate de-matcher, for a batch of subframes that share one
 * allocation: what srsran_chest_ul_estimate_pusch (lib/src/phy/ch_estimation/chest_ul.c:370) and the front half of
 * srsran_pusch_decode (lib/src/phy/phch/pusch.c:392-443) + ulsch_deinterleave (sch.c:993, called from
 * srsran_ulsch_decode sch.c:1150) do per subframe:
 *   DMRS least-squares estimate, 3-tap smoothing, noise / SNR / CFO  ->  MMSE equaliser (srsran_predecoding_single)
 *   ->  transform de-precoding (srsran_dft_precoding, backward DFT of 12*L_prb points / sqrt(N))  ->  int16 soft
 *   demapping  ->  descrambling (srsran_sequence_pusch_apply_s)  ->  UL-SCH de-interleaving.
 * Scope: one receive antenna (as srsran_pusch_decode, pusch.c:413); control information through the _uci_ entries below;
 * no intra-subframe hopping; L_prb >= 1 with
 * 12*L_prb = 2^a 3^b 5^c (srsran_dft_precoding_valid_prb; 1 and 2 PRB use the phi(n) tables of TS 36.211 5.5.1.2).  All data pointers are DEVICE memory (flags must carry
 * SRSRAN_B200_FLAG_DEVICE_PTRS), the per-subframe parameter arrays (rnti, tti, n_dmrs) are HOST memory; every call is
 * enqueued on `stream` and returns without synchronising.
 */
typedef struct srsran_b200_pusch srsran_b200_pusch_t; /* opaque; one per host thread */

typedef struct {
  uint32_t cell_id;             /* physical cell identity: scrambling seed (sequences.c:120) and DMRS sequences */
  uint32_t cell_nof_prb;        /* the resource grid holds 12*cell_nof_prb elements per OFDM symbol */
  int      cp_ext;              /* 0: normal CP (14 symbols, DMRS in symbols 3 and 10), 1: extended (12; 2 and 8) */
  uint32_t L_prb;               /* allocation width in PRB (srsran_pusch_grant_t.L_prb) */
  uint32_t n_prb;               /* first allocated PRB, the same in both slots (grant.n_prb[0] == n_prb[1]) */
  int      modulation;          /* srsran_mod_t: 1 QPSK, 2 16QAM, 3 64QAM */
  uint32_t llr_shift;           /* arithmetic right shift applied to every soft bit after the demapper (0 = reference scale) */
  uint32_t dmrs_cyclic_shift;   /* srsran_refsignal_dmrs_pusch_cfg_t (refsignal_ul.h:46-51) */
  uint32_t dmrs_delta_ss;
  int      group_hopping_en;
  int      sequence_hopping_en;
  int      shortened;           /* srsran_ul_sf_cfg_t.shortened: the subframe's last symbol carries the SRS, the PUSCH has one data
                                   symbol less (srsran_ra_ul_compute_nof_re with N_srs = 1, ra_ul.c:232; pusch.c:63-72) */
} srsran_b200_pusch_cfg_t;

SRSRAN_B200_API int  srsran_b200_pusch_init(srsran_b200_pusch_t** q, int device, const srsran_b200_pusch_cfg_t* cfg);
SRSRAN_B200_API void srsran_b200_pusch_free(srsran_b200_pusch_t* q);
/* nof_re = data symbols * 12 * L_prb (grant.nof_re), nof_bits = nof_re * Qm (grant.tb.nof_bits) */
SRSRAN_B200_API int srsran_b200_pusch_geometry(const srsran_b200_pusch_t* q, uint32_t* nof_re, uint32_t* nof_bits,
                                               uint32_t* nof_data_symbols);

/* srsran_refsignal_dmrs_pusch_gen (refsignal_ul.c:337): the known DMRS of subframe index sf_idx (0..9) for the DCI's cyclic
 * shift n_dmrs (0..7), r = 2 slots x 12*L_prb cf_t written to HOST memory (the object pre-generates all 80 like
 * srsran_refsignal_dmrs_pusch_pregen). */
SRSRAN_B200_API int srsran_b200_refsignal_dmrs_pusch_gen(const srsran_b200_pusch_t* q, uint32_t sf_idx, uint32_t n_dmrs, void* r);

/* srsran_chest_ul_estimate_pusch for nsf subframes.  grid: nsf x nof_symbols x 12*cell_nof_prb cf_t (output of
 * srsran_b200_ofdm_rx_sf_batch); tti[nsf], n_dmrs[nsf] (NULL = all zero): host arrays;
 * ce: nsf x 2 slots x 12*L_prb cf_t -- the reference copies a slot's estimate to every symbol of the slot
 * (chest_ul.c:246-259), the equaliser below reads it per slot instead;
 * meas: nsf x 4 floats {noise_estimate, snr (linear), cfo_hz, ta_us = 0} (srsran_chest_ul_res_t). */
SRSRAN_B200_API int srsran_b200_chest_ul_pusch_batch(srsran_b200_pusch_t* q, const void* grid, uint32_t nsf, const uint32_t* tti,
                                                     const uint32_t* n_dmrs, void* ce, float* meas, uint32_t flags, void* stream);

/* srsran_predecoding_single (noise_estimate = meas[4*sf], NULL = zero forcing) fused into the first pass of
 * srsran_dft_precoding.  d: nsf x nof_re cf_t, data-symbol major like pusch.c's q->d. */
SRSRAN_B200_API int srsran_b200_pusch_equalize_deprecode_batch(srsran_b200_pusch_t* q, const void* grid, const void* ce,
                                                               const float* meas, void* d, uint32_t nsf, uint32_t flags, void* stream);

/* srsran_demod_soft_demodulate_s (one call per subframe over nof_re symbols), >> llr_shift, srsran_sequence_pusch_apply_s
 * with nslot = 2*(tti % 10), ulsch_deinterleave.  g: nsf x nof_bits int16, the e_bits of srsran_b200_sch_decode_batch.
 * rnti[nsf], tti[nsf]: host arrays. */
SRSRAN_B200_API int srsran_b200_pusch_demod_descramble_batch(srsran_b200_pusch_t* q, const void* d, int16_t* g, uint32_t nsf,
                                                             const uint32_t* rnti, const uint32_t* tti, uint32_t flags, void* stream);

/* The three stages above back to back on the object's scratch buffers.  meas may be NULL. */
SRSRAN_B200_API int srsran_b200_pusch_rx_batch(srsran_b200_pusch_t* q, const void* grid, int16_t* g, float* meas, uint32_t nsf,
                                               const uint32_t* rnti, const uint32_t* tti, const uint32_t* n_dmrs, uint32_t flags,
                                               void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Uplink control information multiplexed into the PUSCH (TS 36.212 5.2.2.6-5.2.2.8).  Replaces, per subframe, what
 * srsran_ulsch_decode does before decode_tb (lib/src/phy/phch/sch.c:1121-1190): uci_decode_ri_ack (sch.c:1022-1119,
 * srsran_uci_decode_ack_ri uci.c:637-713), ulsch_deinterleave with the RI positions left out (sch.c:993-1020) and
 * srsran_uci_decode_cqi_pusch (uci.c:289-330) on the front of the de-interleaved stream.  The kernel that demaps, descrambles
 * and de-interleaves also finds the three fields' soft bits (reference scale, before llr_shift) and zeroes the HARQ-ACK
 * positions; the decisions (1 bit, 2 bits, (32,O) block code, CRC-8 + tail-biting convolutional code) run on the host on those
 * few soft bits in srsran_b200_pusch_uci_collect.  g keeps the reference's layout and its one quirk: element 0 of the stream
 * holds the soft bit of the LAST RI position whenever RI is present (sch.c:672-674 maps every RI position to index 0 and
 * srsran_vec_lut_sis scatters in order).  The UL-SCH bits of subframe i start at g + i*nof_bits + e_offset and are nof_e_bits long.
 */
#define SRSRAN_B200_UCI_MAX_ACK_BITS 10 /* SRSRAN_UCI_MAX_ACK_BITS (uci_cfg.h:27) */
#define SRSRAN_B200_UCI_MAX_CQI_BITS 64 /* SRSRAN_CQI_MAX_BITS (cqi.h:38) */

typedef struct {
  uint32_t nof_ack;      /* srsran_uci_cfg_total_ack(&cfg->uci_cfg), 0 = no HARQ-ACK field */
  uint32_t ri_len;       /* cfg->uci_cfg.cqi.ri_len: 0 or 1 (the reference carries a 1-bit RI, uci.c:635) */
  uint32_t cqi_len;      /* srsran_cqi_size(&cfg->uci_cfg.cqi) when cqi.data_enable, else 0 */
  uint32_t I_offset_ack; /* srsran_uci_offset_cfg_t: indices into TS 36.213 Tables 8.6.3-1..3 (sch.c:41-95) */
  uint32_t I_offset_ri;
  uint32_t I_offset_cqi;
} srsran_b200_uci_cfg_t;

typedef struct {
  uint8_t  ack_value[SRSRAN_B200_UCI_MAX_ACK_BITS]; /* srsran_uci_value_t.ack.ack_value; bits not carried stay 2 (srsran_uci_data_reset) */
  uint8_t  ack_valid;                               /* ack.valid: correlation above the reference's threshold (uci.c:694-710) */
  uint8_t  ri;
  uint8_t  cqi_crc;                                 /* cqi.data_crc: 1 for the block-coded form, the CRC-8 verdict above 11 bits */
  uint8_t  reserved;
  uint8_t  cqi_bits[SRSRAN_B200_UCI_MAX_CQI_BITS];  /* the cqi_len payload bits, one per byte: input of srsran_cqi_value_unpack */
  uint32_t Q_prime_ack, Q_prime_ri, Q_prime_cqi;    /* coded modulation symbols of each field */
  uint32_t e_offset, nof_e_bits;                    /* the UL-SCH part of the subframe's g: arguments of the transport-block decode */
} srsran_b200_uci_value_t;

/* Host only: Q' of the three fields and the UL-SCH span for a grant of transport block size tbs on this object's allocation
 * (Q_prime_ri_ack / Q_prime_cqi, uci.c:172-190,395-418, with K_segm from srsran_cbsegm).  Fills the last five members of out. */
SRSRAN_B200_API int srsran_b200_pusch_uci_geometry(const srsran_b200_pusch_t* q, uint32_t tbs, const srsran_b200_uci_cfg_t* uci,
                                                   srsran_b200_uci_value_t* out);

/* srsran_b200_pusch_rx_batch for subframes that carry control information: tbs[nsf] and uci[nsf] are host arrays (an all-zero
 * uci[i] is a subframe without control information).  Enqueues the kernels and the copy of the fields' soft bits on `stream`. */
SRSRAN_B200_API int srsran_b200_pusch_rx_uci_batch(srsran_b200_pusch_t* q, const void* grid, int16_t* g, float* meas, uint32_t nsf,
                                                   const uint32_t* rnti, const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* tbs,
                                                   const srsran_b200_uci_cfg_t* uci, uint32_t flags, void* stream);

/* Host only, no GPU involved: the decisions of srsran_b200_pusch_uci_collect on soft bits the caller already has, in the order the
 * reference walks them (field symbol by field symbol, Qm soft bits each; for the 1-bit forms with the repeated bit's scrambling
 * already undone).  ack_llr: Q_prime_ack*Qm values, ri_llr: Q_prime_ri*Qm, cqi_llr: Q_prime_cqi*Qm (NULL where the field is absent).
 * Fills the value members of out.  srsran_uci_decode_ack_ri / srsran_uci_decode_cqi_pusch (uci.c:289-330,637-713) without the
 * position search. */
SRSRAN_B200_API int srsran_b200_uci_decide(const srsran_b200_uci_cfg_t* uci, uint32_t Qm, uint32_t Q_prime_ack, uint32_t Q_prime_ri,
                                           uint32_t Q_prime_cqi, const int16_t* ack_llr, const int16_t* ri_llr, const int16_t* cqi_llr,
                                           srsran_b200_uci_value_t* out);

/* Waits for the soft bits of every srsran_b200_pusch_rx_uci_batch call since the last collect and decides them, in call order:
 * out[0 .. sum of those calls' nsf).  nof_out must equal that sum; out = NULL discards what is pending. */
SRSRAN_B200_API int srsran_b200_pusch_uci_collect(srsran_b200_pusch_t* q, srsran_b200_uci_value_t* out, uint32_t nof_out);

/* ---------------------------------------------------------------------------------------------------------------
 * One cell's PUSCH receiver for a batch of subframes in ONE call: time samples in, transport-block bytes out.  Replaces, per
 * subframe, srsran_enb_ul_fft (lib/src/phy/enb/enb_ul.c:151-154) and get_pusch (enb_ul.c:262-290: srsran_chest_ul_estimate_pusch +
 * srsran_pusch_decode) for subframes that all carry the configured allocation.  The object owns the three engines above, the
 * device buffers between them and the HARQ soft buffers (slot i = subframe index i of the batch; the caller maps
 * (UE, HARQ process) to slots and passes new_data[i] = 0, rv[i] for a retransmission into slot i).  With host samples the
 * host->device copies are chunked on a second stream and overlap the front-end kernels.
 */
typedef struct srsran_b200_enb_ul srsran_b200_enb_ul_t; /* opaque; one per host thread */

typedef struct {
  uint32_t cell_id;
  uint32_t cell_nof_prb;
  int      cp_ext;
  uint32_t symbol_sz;           /* 0 = srsran_symbol_sz(cell_nof_prb) */
  uint32_t dmrs_cyclic_shift;   /* srsran_refsignal_dmrs_pusch_cfg_t */
  uint32_t dmrs_delta_ss;
  int      group_hopping_en;
  int      sequence_hopping_en;
  uint32_t L_prb;               /* the allocation of every subframe of a batch */
  uint32_t n_prb;
  int      modulation;          /* 1 QPSK, 2 16QAM, 3 64QAM */
  uint32_t tbs;                 /* transport block size in bits (a standard size: no filler bits) */
  uint32_t llr_shift;           /* see srsran_b200_pusch_cfg_t */
  uint32_t max_iterations;      /* decoder passes per code block, 0 = 8 */
  int      shortened;           /* see srsran_b200_pusch_cfg_t: every subframe of the batch is an SRS subframe */
} srsran_b200_enb_ul_cfg_t;

typedef struct {
  int32_t crc_ok;          /* srsran_pusch_res_t.crc */
  float   avg_iterations;  /* srsran_pusch_res_t.avg_iterations_block */
  float   noise_estimate;  /* srsran_chest_ul_res_t */
  float   snr;
  float   cfo_hz;
} srsran_b200_pusch_res_t;

SRSRAN_B200_API int  srsran_b200_enb_ul_init(srsran_b200_enb_ul_t** q, int device, const srsran_b200_enb_ul_cfg_t* cfg);
SRSRAN_B200_API void srsran_b200_enb_ul_free(srsran_b200_enb_ul_t* q);
SRSRAN_B200_API int  srsran_b200_enb_ul_geometry(const srsran_b200_enb_ul_t* q, uint32_t* sf_sz, uint32_t* tb_bytes);
/*
 * samples  nsf x sf_sz I/Q samples: cf_t, or int16 pairs with SRSRAN_B200_FLAG_IQ_INT16; host memory (page-locked for full PCIe
 *          speed) or, with SRSRAN_B200_FLAG_DEVICE_PTRS, device memory (then `data` is device memory too)
 * rnti, tti, n_dmrs, rv, new_data   per-subframe host arrays; NULL = 0 (new_data: NULL = all new)
 * data     nsf x tb_bytes (= tbs/8 + 3: payload + CRC24A), rows tightly packed
 * res      nsf results (host)
 * Synchronous: returns when data and res are complete.
 */
SRSRAN_B200_API int srsran_b200_enb_ul_pusch_batch(srsran_b200_enb_ul_t* q, const void* samples, uint32_t nsf, const uint32_t* rnti,
                                                   const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv, const uint32_t* new_data,
                                                   uint8_t* data, srsran_b200_pusch_res_t* res, uint32_t flags);

/* The same call for subframes that carry control information (get_pusch with cfg->uci_cfg set, enb_ul.c:262-290): uci[nsf] in,
 * uci_out[nsf] out; the UL-SCH bits are de-matched from what is left of each subframe after the RI and CQI symbols. */
SRSRAN_B200_API int srsran_b200_enb_ul_pusch_uci_batch(srsran_b200_enb_ul_t* q, const void* samples, uint32_t nsf, const uint32_t* rnti,
                                                       const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv,
                                                       const uint32_t* new_data, const srsran_b200_uci_cfg_t* uci, uint8_t* data,
                                                       srsran_b200_pusch_res_t* res, srsran_b200_uci_value_t* uci_out, uint32_t flags);

/* The call in two halves, for ONE caller thread that serves several cells (one object each) or keeps two batches of a cell in
 * flight: _begin queues the sample copies, the front end and every decode group on the object's own streams and returns; _finish
 * waits and fills data, res and uci_out (which must stay valid in between; uci and uci_out may be NULL for subframes without control
 * information).  One batch per object at a time.  srsran_b200_enb_ul_pusch_batch is _begin followed by _finish. */
SRSRAN_B200_API int srsran_b200_enb_ul_pusch_batch_begin(srsran_b200_enb_ul_t* q, const void* samples, uint32_t nsf, const uint32_t* rnti,
                                                         const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv,
                                                         const uint32_t* new_data, const srsran_b200_uci_cfg_t* uci, uint8_t* data,
                                                         srsran_b200_pusch_res_t* res, srsran_b200_uci_value_t* uci_out, uint32_t flags);
SRSRAN_B200_API int srsran_b200_enb_ul_pusch_batch_finish(srsran_b200_enb_ul_t* q);

#ifdef __cplusplus
}
#endif

#endif /* SRSLTE_B200_H */
