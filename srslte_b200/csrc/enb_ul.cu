// Native PUSCH receive pipeline of one cell for a batch of subframes: what srsran_enb_ul_fft (lib/src/phy/enb/enb_ul.c:151-154)
// and get_pusch (enb_ul.c:262-290: srsran_chest_ul_estimate_pusch + srsran_pusch_decode) do per subframe, as ONE call with time
// samples in and transport-block bytes out.  It only orchestrates the batched entries of this library
//     srsran_b200_ofdm_rx_sf_batch -> srsran_b200_pusch_rx_batch -> srsran_b200_sch_decode_batch
// on device buffers it owns: with host samples the batch is cut into chunks whose host->device copies (second stream) overlap the
// front-end kernels of the previous chunk, the decode loop then runs over the whole batch (it needs the batch to fill the GPU)
// while its host-side bookkeeping overlaps the tail of the front end.  HARQ soft buffers live in the object, one slot per
// subframe index of the batch (the caller maps (UE, HARQ process) to slots).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <new>
#include <vector>

#include "../../include/srslte_b200.h"
#include "b200_runtime.h"
#include "tdec_engine.h"

namespace b200 {

struct EnbUl {
  int                       device = 0;
  srsran_b200_enb_ul_cfg_t  cfg{};
  srsran_b200_ofdm_t*       ofdm  = nullptr;
  srsran_b200_pusch_t*      pusch = nullptr;
  srsran_b200_sch_t*        sch   = nullptr;
  // With host samples a large batch is decoded in groups, each as soon as its front end is done, so that the decoder works
  // while the later groups' samples are still crossing PCIe.  One decode object per group: each keeps the cached plan of its
  // own slice of the transport-block list.
  static constexpr uint32_t MAX_GROUPS = 8, MIN_GROUP_SF = 1024;
  srsran_b200_sch_t*        sch_g[MAX_GROUPS] = {};
  cudaEvent_t               ev_g[MAX_GROUPS]  = {};
  cudaEvent_t               ev_t0 = nullptr; // SRSLTE_B200_ENB_UL_TIMING
  cudaStream_t              compute = nullptr, copy = nullptr, out_st = nullptr;
  std::vector<cudaEvent_t>  ev;
  uint32_t sf_sz = 0, nsym = 0, nre = 0, nbits = 0, ncb = 0, data_stride = 0;
  // device buffers, grow-only
  uint32_t cap_sf = 0, cap_iq = 0;
  void*    d_iq   = nullptr;
  float2*  d_grid = nullptr;
  int16_t* d_llr  = nullptr;
  int16_t* d_soft = nullptr;
  uint8_t* d_data = nullptr;
  float*   d_meas = nullptr;
  std::vector<srsran_b200_tb_t> tbs;
  std::vector<uint32_t>         tbs_vec; // the configured size once per subframe: argument of the UCI entry
  std::vector<uint32_t>         crc_mask; // per slot, carried across retransmissions (sch.c:474-488)
  std::vector<float>            h_meas;

  ~EnbUl()
  {
    cudaSetDevice(device);
    if (pend.active) finish();
    if (ofdm) srsran_b200_ofdm_rx_free(ofdm);
    if (pusch) srsran_b200_pusch_free(pusch);
    if (sch) srsran_b200_sch_free(sch);
    for (uint32_t g = 1; g < MAX_GROUPS; g++) {
      if (sch_g[g]) srsran_b200_sch_free(sch_g[g]);
    }
    for (cudaEvent_t e : ev_g) {
      if (e) cudaEventDestroy(e);
    }
    if (out_st) cudaStreamDestroy(out_st);
    for (void* p : {d_iq, (void*)d_grid, (void*)d_llr, (void*)d_soft, (void*)d_data, (void*)d_meas}) {
      if (p) cudaFree(p);
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (compute) cudaStreamDestroy(compute);
    if (copy) cudaStreamDestroy(copy);
  }

  int init(int dev, const srsran_b200_enb_ul_cfg_t& c)
  {
    device = dev;
    cfg    = c;
    if (c.modulation < 1 || c.modulation > 3 || c.tbs == 0) return B200_ERROR_INVALID_INPUTS;
    B200_CUDA_TRY(cudaSetDevice(dev));
    // the eNB's uplink OFDM configuration (enb_ul.c:50-58): half-subcarrier shift, window advanced by half a CP, no normalisation
    srsran_b200_ofdm_cfg_t oc;
    memset(&oc, 0, sizeof(oc));
    oc.nof_prb          = c.cell_nof_prb;
    oc.cp_ext           = c.cp_ext;
    oc.symbol_sz        = c.symbol_sz;
    oc.freq_shift_f     = -0.5f;
    oc.rx_window_offset = 0.5f;
    int rc              = srsran_b200_ofdm_rx_init(&ofdm, dev, &oc);
    if (rc != B200_SUCCESS) return rc;
    srsran_b200_pusch_cfg_t pc;
    memset(&pc, 0, sizeof(pc));
    pc.cell_id             = c.cell_id;
    pc.cell_nof_prb        = c.cell_nof_prb;
    pc.cp_ext              = c.cp_ext;
    pc.L_prb               = c.L_prb;
    pc.n_prb               = c.n_prb;
    pc.modulation          = c.modulation;
    pc.llr_shift           = c.llr_shift;
    pc.dmrs_cyclic_shift   = c.dmrs_cyclic_shift;
    pc.dmrs_delta_ss       = c.dmrs_delta_ss;
    pc.group_hopping_en    = c.group_hopping_en;
    pc.sequence_hopping_en = c.sequence_hopping_en;
    pc.shortened           = c.shortened;
    if ((rc = srsran_b200_pusch_init(&pusch, dev, &pc)) != B200_SUCCESS) return rc;
    if ((rc = srsran_b200_sch_init(&sch, dev)) != B200_SUCCESS) return rc;
    srsran_b200_sch_set_max_noi(sch, c.max_iterations ? c.max_iterations : 8);
    uint32_t symsz = 0;
    srsran_b200_ofdm_rx_geometry(ofdm, &symsz, &sf_sz, &nsym, &nre);
    uint32_t nof_re = 0, nd = 0;
    srsran_b200_pusch_geometry(pusch, &nof_re, &nbits, &nd);
    // code blocks of the transport block (36.212 5.1.2): B = tbs + 24, C = ceil(B / (6144 - 24)) when B > 6144
    const uint32_t B = c.tbs + 24;
    ncb              = B <= 6144 ? 1 : (B + 6119) / 6120;
    data_stride      = (c.tbs / 8 + 3 + 768 + 15) / 16 * 16;
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
    B200_CUDA_TRY(cudaStreamCreateWithFlags(&out_st, cudaStreamNonBlocking));
    sch_g[0] = sch;
    return B200_SUCCESS;
  }

  // Grows the per-slot buffers.  Slots are HARQ processes: the soft buffers, the payload bytes of code blocks decoded in an
  // earlier transmission and the per-slot CRC masks of the slots that already exist survive the growth (a retransmission
  // into slot i after a larger batch still combines with what slot i received before).
  int reserve(uint32_t nsf)
  {
    if (nsf > cap_sf) {
      const size_t soft_slot = (size_t)ncb * SRSRAN_B200_SOFTBUFFER_SIZE * sizeof(int16_t);
      int16_t*     n_soft = nullptr;
      uint8_t*     n_data = nullptr;
      B200_CUDA_TRY(cudaMalloc(&n_soft, (size_t)nsf * soft_slot));
      B200_CUDA_TRY(cudaMalloc(&n_data, (size_t)nsf * data_stride));
      B200_CUDA_TRY(cudaMemset(n_soft, 0, (size_t)nsf * soft_slot));
      B200_CUDA_TRY(cudaMemset(n_data, 0, (size_t)nsf * data_stride));
      if (cap_sf) { // (cudaMemcpy: synchronous with respect to the host, after whatever the object's streams still run)
        B200_CUDA_TRY(cudaDeviceSynchronize());
        B200_CUDA_TRY(cudaMemcpy(n_soft, d_soft, (size_t)cap_sf * soft_slot, cudaMemcpyDeviceToDevice));
        B200_CUDA_TRY(cudaMemcpy(n_data, d_data, (size_t)cap_sf * data_stride, cudaMemcpyDeviceToDevice));
      }
      for (void* p : {(void*)d_grid, (void*)d_llr, (void*)d_soft, (void*)d_data, (void*)d_meas, d_iq}) {
        if (p) cudaFree(p);
      }
      d_soft = n_soft;
      d_data = n_data;
      d_grid = nullptr; d_llr = nullptr; d_meas = nullptr; d_iq = nullptr;
      const uint32_t old_sf = cap_sf;
      cap_sf = 0;
      B200_CUDA_TRY(cudaMalloc(&d_grid, (size_t)nsf * nsym * nre * sizeof(float2)));
      B200_CUDA_TRY(cudaMalloc(&d_llr, (size_t)nsf * nbits * sizeof(int16_t)));
      B200_CUDA_TRY(cudaMalloc(&d_meas, (size_t)nsf * 4 * sizeof(float)));
      cap_iq = 0; // the staging buffer of host samples is allocated when host samples arrive (run)
      cap_sf = nsf;
      tbs.resize(nsf, srsran_b200_tb_t{});
      crc_mask.resize(nsf, 0u); // existing slots keep their masks
      h_meas.assign((size_t)nsf * 4, 0.f);
      for (uint32_t i = old_sf; i < nsf; i++) {
        tbs[i].tbs         = cfg.tbs;
        tbs[i].Qm          = 2u * (uint32_t)cfg.modulation;
        tbs[i].nof_e_bits  = nbits;
        tbs[i].e_offset    = (uint64_t)i * nbits;
        tbs[i].soft_offset = (uint64_t)i * ncb * SRSRAN_B200_SOFTBUFFER_SIZE;
        tbs[i].data_offset = (uint64_t)i * data_stride;
      }
    }
    return B200_SUCCESS;
  }

  // a batch between begin() and finish(): run() is the two back to back
  struct Pending {
    bool                     active = false, dev_ptrs = false, uci = false;
    uint32_t                 nsf = 0, chunk = 0, nchunks = 0, cpg = 0, ngroups = 0, begun = 0;
    uint8_t*                 data    = nullptr;
    srsran_b200_pusch_res_t* res     = nullptr;
    srsran_b200_uci_value_t* uci_out = nullptr;
    std::chrono::steady_clock::time_point h0, h1, h2;
  } pend;

  int run(const void* samples, uint32_t nsf, const uint32_t* rnti, const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv,
          const uint32_t* new_data, uint8_t* data, srsran_b200_pusch_res_t* res, uint32_t flags, const srsran_b200_uci_cfg_t* uci = nullptr,
          srsran_b200_uci_value_t* uci_out = nullptr)
  {
    const int rc = begin(samples, nsf, rnti, tti, n_dmrs, rv, new_data, data, res, flags, uci, uci_out);
    if (rc != B200_SUCCESS) return rc;
    return finish();
  }

  // Queues the whole batch (sample copies, front end, every decode group) and returns; data / res / uci_out are written by finish()
  int begin(const void* samples, uint32_t nsf, const uint32_t* rnti, const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv,
            const uint32_t* new_data, uint8_t* data, srsran_b200_pusch_res_t* res, uint32_t flags, const srsran_b200_uci_cfg_t* uci = nullptr,
            srsran_b200_uci_value_t* uci_out = nullptr)
  {
    if (!samples || !data || !res || (uci && !uci_out)) return B200_ERROR_INVALID_INPUTS;
    if (pend.active) {
      B200_LOG_ERROR("srsran_b200_enb_ul_pusch_batch_begin: the previous batch has not been finished");
      return B200_ERROR_INVALID_INPUTS;
    }
    if (nsf == 0) return B200_SUCCESS;
    B200_CUDA_TRY(cudaSetDevice(device));
    const bool   dev_ptrs = (flags & SRSRAN_B200_FLAG_DEVICE_PTRS) != 0;
    const bool   iq16     = (flags & SRSRAN_B200_FLAG_IQ_INT16) != 0;
    const size_t ssz      = iq16 ? 2 * sizeof(int16_t) : sizeof(float2);
    int          rc       = reserve(nsf);
    if (rc != B200_SUCCESS) return rc;
    if (!dev_ptrs && cap_iq < cap_sf) {
      if (d_iq) cudaFree(d_iq);
      d_iq = nullptr;
      B200_CUDA_TRY(cudaMalloc(&d_iq, (size_t)cap_sf * sf_sz * sizeof(float2))); // sized for float samples
      cap_iq = cap_sf;
    }
    const uint32_t fe_flags = SRSRAN_B200_FLAG_DEVICE_PTRS | (iq16 ? SRSRAN_B200_FLAG_IQ_INT16 : 0u);
    // where the UL-SCH bits of each subframe lie: everything, or what the RI and CQI symbols leave (sch.c:1186-1190)
    if (uci) tbs_vec.assign(nsf, cfg.tbs);
    for (uint32_t i = 0; i < nsf; i++) {
      uint32_t e_off = 0, e_bits = nbits;
      if (uci) {
        srsran_b200_uci_value_t span;
        if ((rc = srsran_b200_pusch_uci_geometry(pusch, cfg.tbs, &uci[i], &span)) != B200_SUCCESS) return rc;
        e_off  = span.e_offset;
        e_bits = span.nof_e_bits;
      }
      tbs[i].e_offset   = (uint64_t)i * nbits + e_off;
      tbs[i].nof_e_bits = e_bits;
    }

    // ---- front end, chunk by chunk -----------------------------------------------------------------------------------------
    const uint32_t chunk   = dev_ptrs ? nsf : (nsf > 1024 ? 512u : (nsf + 1) / 2);
    const uint32_t nchunks = (nsf + chunk - 1) / chunk;
    // decode groups: whole chunks, at least MIN_GROUP_SF subframes each (a smaller group leaves most SMs without a tile)
    static const bool no_groups = getenv("SRSLTE_B200_ENB_UL_NO_GROUPS") != nullptr;
    static const bool timing    = getenv("SRSLTE_B200_ENB_UL_TIMING") != nullptr; // where a call's time goes: host stamps + event times
    const unsigned    ev_flags  = timing ? cudaEventDefault : cudaEventDisableTiming;
    auto              now       = [] { return std::chrono::steady_clock::now(); };
    const auto        h0        = now();
    if (timing) {
      if (!ev_t0) B200_CUDA_TRY(cudaEventCreate(&ev_t0));
      B200_CUDA_TRY(cudaEventRecord(ev_t0, copy));
    }
    static const char* gsf_env  = getenv("SRSLTE_B200_ENB_UL_GROUP_SF");
    const uint32_t    group_sf  = gsf_env && atoi(gsf_env) >= 64 ? (uint32_t)atoi(gsf_env) : MIN_GROUP_SF;
    uint32_t          ngroups   = (dev_ptrs || no_groups) ? 1u : nsf / group_sf;
    ngroups                     = ngroups < 1 ? 1 : ngroups > MAX_GROUPS ? MAX_GROUPS : ngroups;
    const uint32_t cpg          = (nchunks + ngroups - 1) / ngroups; // chunks per group
    ngroups                     = (nchunks + cpg - 1) / cpg;
    for (uint32_t g = 0; g < ngroups; g++) {
      if (!sch_g[g]) {
        if ((rc = srsran_b200_sch_init(&sch_g[g], device)) != B200_SUCCESS) return rc;
        srsran_b200_sch_set_max_noi(sch_g[g], cfg.max_iterations ? cfg.max_iterations : 8);
      }
      if (!ev_g[g]) B200_CUDA_TRY(cudaEventCreateWithFlags(&ev_g[g], ev_flags));
    }
    while (ev.size() < nchunks) {
      cudaEvent_t e;
      B200_CUDA_TRY(cudaEventCreateWithFlags(&e, ev_flags));
      ev.push_back(e);
    }
    for (uint32_t c = 0; c < nchunks; c++) {
      const uint32_t first = c * chunk, n = (nsf - first) < chunk ? (nsf - first) : chunk;
      const char*    in    = (const char*)samples + (size_t)first * sf_sz * ssz;
      const void*    d_in  = in;
      if (!dev_ptrs) {
        char* dst = (char*)d_iq + (size_t)first * sf_sz * ssz;
        B200_CUDA_TRY(cudaMemcpyAsync(dst, in, (size_t)n * sf_sz * ssz, cudaMemcpyHostToDevice, copy));
        B200_CUDA_TRY(cudaEventRecord(ev[c], copy));
        B200_CUDA_TRY(cudaStreamWaitEvent(compute, ev[c], 0));
        d_in = dst;
      }
      float2* grid_c = d_grid + (size_t)first * nsym * nre;
      if ((rc = srsran_b200_ofdm_rx_sf_batch(ofdm, d_in, grid_c, n, fe_flags, compute)) != B200_SUCCESS) return rc;
      if (uci) {
        rc = srsran_b200_pusch_rx_uci_batch(pusch, grid_c, d_llr + (size_t)first * nbits, d_meas + (size_t)first * 4, n, rnti ? rnti + first : nullptr,
                                            tti ? tti + first : nullptr, n_dmrs ? n_dmrs + first : nullptr, tbs_vec.data() + first, uci + first,
                                            SRSRAN_B200_FLAG_DEVICE_PTRS, compute);
      } else {
        rc = srsran_b200_pusch_rx_batch(pusch, grid_c, d_llr + (size_t)first * nbits, d_meas + (size_t)first * 4, n, rnti ? rnti + first : nullptr,
                                        tti ? tti + first : nullptr, n_dmrs ? n_dmrs + first : nullptr, SRSRAN_B200_FLAG_DEVICE_PTRS, compute);
      }
      if (rc != B200_SUCCESS) {
        if (uci) srsran_b200_pusch_uci_collect(pusch, nullptr, 0); // drop the chunks already queued
        return rc;
      }
      if ((c + 1) % cpg == 0 || c + 1 == nchunks) B200_CUDA_TRY(cudaEventRecord(ev_g[c / cpg], compute));
    }

    // ---- decode, group by group; the bytes of a group travel back while the next one is decoded ---------------------------------
    for (uint32_t i = 0; i < nsf; i++) {
      const bool fresh   = new_data ? new_data[i] != 0 : true;
      tbs[i].rv          = rv ? rv[i] : 0u;
      tbs[i].new_data    = fresh ? 1u : 0u;
      tbs[i].cb_crc_mask = fresh ? 0u : crc_mask[i];
    }
    const auto h1 = now();
    // every group is queued behind its own front end right away (the groups' kernels share the SMs as their inputs arrive) ...
    auto span = [&](uint32_t g, uint32_t* first, uint32_t* last) {
      *first = g * cpg * chunk;
      *last  = (g + 1) * cpg * chunk < nsf ? (g + 1) * cpg * chunk : nsf;
    };
    uint32_t begun = 0;
    for (uint32_t g = 0; g < ngroups && rc == B200_SUCCESS; g++) {
      uint32_t first, last;
      span(g, &first, &last);
      srsran_b200_sch_decode_after_event(sch_g[g], ev_g[g]);
      rc = srsran_b200_sch_decode_begin(sch_g[g], d_llr, (uint64_t)nsf * nbits, d_soft, (uint64_t)cap_sf * ncb * SRSRAN_B200_SOFTBUFFER_SIZE, d_data,
                                        (uint64_t)cap_sf * data_stride, tbs.data() + first, last - first, SRSRAN_B200_FLAG_DEVICE_PTRS);
      if (rc == B200_SUCCESS) begun++;
    }
    pend          = Pending();
    pend.dev_ptrs = dev_ptrs;
    pend.uci      = uci != nullptr;
    pend.nsf = nsf; pend.chunk = chunk; pend.nchunks = nchunks; pend.cpg = cpg; pend.ngroups = ngroups; pend.begun = begun;
    pend.data = data; pend.res = res; pend.uci_out = uci_out;
    pend.h0 = h0; pend.h1 = h1; pend.h2 = now();
    pend.active = true;
    if (rc != B200_SUCCESS) { // some group could not be queued: wait for the others and give up
      finish();
      return rc;
    }
    return B200_SUCCESS;
  }

  // Waits for the batch of begin(): the groups are finished in order, the bytes of a group travel back while the next ones decode
  int finish()
  {
    if (!pend.active) return B200_SUCCESS;
    cudaSetDevice(device);
    static const bool timing = getenv("SRSLTE_B200_ENB_UL_TIMING") != nullptr;
    auto              now    = [] { return std::chrono::steady_clock::now(); };
    auto              us     = [](auto a, auto b) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() / 1e3; };
    const uint32_t nsf = pend.nsf, chunk = pend.chunk, nchunks = pend.nchunks, cpg = pend.cpg, ngroups = pend.ngroups, begun = pend.begun;
    const bool     dev_ptrs = pend.dev_ptrs, uci = pend.uci;
    uint8_t*       data = pend.data;
    srsran_b200_pusch_res_t* res     = pend.res;
    srsran_b200_uci_value_t* uci_out = pend.uci_out;
    const auto     h0 = pend.h0, h1 = pend.h1, h2 = pend.h2;
    const size_t   out_b = (size_t)cfg.tbs / 8 + 3;
    int            rc    = begun == ngroups ? B200_SUCCESS : B200_ERROR;
    pend.active          = false;
    auto span = [&](uint32_t g, uint32_t* first, uint32_t* last) {
      *first = g * cpg * chunk;
      *last  = (g + 1) * cpg * chunk < nsf ? (g + 1) * cpg * chunk : nsf;
    };
    double fin_us[MAX_GROUPS] = {};
    for (uint32_t g = 0; g < begun; g++) {
      uint32_t first, last;
      span(g, &first, &last);
      const int r = srsran_b200_sch_decode_finish(sch_g[g]);
      fin_us[g]   = us(h0, now());
      if (r != B200_SUCCESS && rc == B200_SUCCESS) rc = r;
      if (rc == B200_SUCCESS) {
        const cudaError_t ce = cudaMemcpy2DAsync(data + (size_t)first * out_b, out_b, d_data + (size_t)first * data_stride, data_stride, out_b,
                                                 last - first, dev_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, out_st);
        if (ce != cudaSuccess) rc = B200_ERROR;
      }
    }
    if (rc != B200_SUCCESS) {
      if (uci) srsran_b200_pusch_uci_collect(pusch, nullptr, 0);
      cudaStreamSynchronize(out_st);
      return rc;
    }
    B200_CUDA_TRY(cudaMemcpyAsync(h_meas.data(), d_meas, (size_t)nsf * 4 * sizeof(float), cudaMemcpyDeviceToHost, compute));
    B200_CUDA_TRY(cudaStreamSynchronize(out_st));
    B200_CUDA_TRY(cudaStreamSynchronize(compute));
    if (timing && !dev_ptrs) {
      fprintf(stderr, "[enb_ul timing] nsf %u chunks %u groups %u: host enqueue front end %.0f us, begin %.0f; copies done at", nsf, nchunks, ngroups,
              us(h0, h1), us(h0, h2));
      for (uint32_t c = 0; c < nchunks; c++) {
        float t = 0;
        cudaEventElapsedTime(&t, ev_t0, ev[c]);
        fprintf(stderr, " %.2f", t);
      }
      fprintf(stderr, " ms; front end of group done at");
      for (uint32_t g = 0; g < ngroups; g++) {
        float t = 0;
        cudaEventElapsedTime(&t, ev_t0, ev_g[g]);
        fprintf(stderr, " %.2f", t);
      }
      fprintf(stderr, " ms; decode finished (host) at");
      for (uint32_t g = 0; g < ngroups; g++) fprintf(stderr, " %.2f", fin_us[g] / 1e3);
      fprintf(stderr, " ms; end %.2f ms\n", us(h0, now()) / 1e3);
    }
    if (uci && (rc = srsran_b200_pusch_uci_collect(pusch, uci_out, nsf)) != B200_SUCCESS) return rc;
    for (uint32_t i = 0; i < nsf; i++) {
      crc_mask[i]           = tbs[i].cb_crc_mask;
      res[i].crc_ok         = tbs[i].result == B200_SUCCESS ? 1 : 0;
      res[i].avg_iterations = tbs[i].avg_iterations;
      res[i].noise_estimate = h_meas[4 * (size_t)i + 0];
      res[i].snr            = h_meas[4 * (size_t)i + 1];
      res[i].cfo_hz         = h_meas[4 * (size_t)i + 2];
    }
    return B200_SUCCESS;
  }
};

} // namespace b200

using namespace b200;

struct srsran_b200_enb_ul {
  EnbUl e;
};

extern "C" SRSRAN_B200_API int srsran_b200_enb_ul_init(srsran_b200_enb_ul_t** q, int device, const srsran_b200_enb_ul_cfg_t* cfg)
{
  if (!q || !cfg) return B200_ERROR_INVALID_INPUTS;
  *q                      = nullptr;
  srsran_b200_enb_ul_t* h = new (std::nothrow) srsran_b200_enb_ul_t();
  if (!h) return B200_ERROR;
  const int rc = h->e.init(device, *cfg);
  if (rc != B200_SUCCESS) {
    delete h;
    return rc;
  }
  *q = h;
  return B200_SUCCESS;
}

extern "C" SRSRAN_B200_API void srsran_b200_enb_ul_free(srsran_b200_enb_ul_t* q)
{
  delete q;
}

extern "C" SRSRAN_B200_API int srsran_b200_enb_ul_geometry(const srsran_b200_enb_ul_t* q, uint32_t* sf_sz, uint32_t* tb_bytes)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  if (sf_sz) *sf_sz = q->e.sf_sz;
  if (tb_bytes) *tb_bytes = q->e.cfg.tbs / 8 + 3;
  return B200_SUCCESS;
}

extern "C" SRSRAN_B200_API int srsran_b200_enb_ul_pusch_batch_begin(srsran_b200_enb_ul_t* q, const void* samples, uint32_t nsf, const uint32_t* rnti,
                                                                   const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv,
                                                                   const uint32_t* new_data, const srsran_b200_uci_cfg_t* uci, uint8_t* data,
                                                                   srsran_b200_pusch_res_t* res, srsran_b200_uci_value_t* uci_out, uint32_t flags)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  return q->e.begin(samples, nsf, rnti, tti, n_dmrs, rv, new_data, data, res, flags, uci, uci_out);
}

extern "C" SRSRAN_B200_API int srsran_b200_enb_ul_pusch_batch_finish(srsran_b200_enb_ul_t* q)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  return q->e.finish();
}

extern "C" SRSRAN_B200_API int srsran_b200_enb_ul_pusch_uci_batch(srsran_b200_enb_ul_t* q, const void* samples, uint32_t nsf, const uint32_t* rnti,
                                                                 const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv,
                                                                 const uint32_t* new_data, const srsran_b200_uci_cfg_t* uci, uint8_t* data,
                                                                 srsran_b200_pusch_res_t* res, srsran_b200_uci_value_t* uci_out, uint32_t flags)
{
  if (!q || !uci || !uci_out) return B200_ERROR_INVALID_INPUTS;
  return q->e.run(samples, nsf, rnti, tti, n_dmrs, rv, new_data, data, res, flags, uci, uci_out);
}

extern "C" SRSRAN_B200_API int srsran_b200_enb_ul_pusch_batch(srsran_b200_enb_ul_t* q, const void* samples, uint32_t nsf, const uint32_t* rnti,
                                                             const uint32_t* tti, const uint32_t* n_dmrs, const uint32_t* rv,
                                                             const uint32_t* new_data, uint8_t* data, srsran_b200_pusch_res_t* res,
                                                             uint32_t flags)
{
  if (!q) return B200_ERROR_INVALID_INPUTS;
  return q->e.run(samples, nsf, rnti, tti, n_dmrs, rv, new_data, data, res, flags);
}
