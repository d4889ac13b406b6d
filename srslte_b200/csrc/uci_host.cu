// Host half of the control information multiplexed into the PUSCH: see uci_host.h for the reference map.
#include "uci_host.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "b200_runtime.h"

namespace b200 {

// TS 36.213 Tables 8.6.3-1 / -2 / -3 (sch.c:41-95).  A reserved or out-of-range index logs an error and falls back to the
// first valid entry, like the reference's accessors.
static float beta_harq(uint32_t idx)
{
  static const float t[15] = {2.0f, 2.5f, 3.125f, 4.0f, 5.0f, 6.25f, 8.0f, 10.0f, 12.625f, 15.875f, 20.0f, 31.0f, 50.0f, 80.0f, 126.0f};
  if (idx < 15) return t[idx];
  B200_LOG_ERROR("I_offset_ack %u out of range (0..14)", idx);
  return t[0];
}
static float beta_ri(uint32_t idx)
{
  static const float t[13] = {1.25f, 1.625f, 2.0f, 2.5f, 3.125f, 4.0f, 5.0f, 6.25f, 8.0f, 10.0f, 12.625f, 15.875f, 20.0f};
  if (idx < 13) return t[idx];
  B200_LOG_ERROR("I_offset_ri %u out of range (0..12)", idx);
  return t[0];
}
static float beta_cqi(uint32_t idx)
{
  static const float t[16] = {-1.0f, -1.0f, 1.125f, 1.25f, 1.375f, 1.625f, 1.75f, 2.0f, 2.25f, 2.5f, 2.875f, 3.125f, 3.5f, 4.0f, 5.0f, 6.25f};
  if (idx > 1 && idx < 16) return t[idx];
  B200_LOG_ERROR("I_offset_cqi %u out of range (2..15)", idx);
  return t[2];
}

int uci_geometry(const srsran_b200_uci_cfg_t& c, uint32_t K_segm, uint32_t M_sc, uint32_t nsymb, UciGeometry* g)
{
  *g = UciGeometry();
  if (c.nof_ack == 0 && c.ri_len == 0 && c.cqi_len == 0) return B200_SUCCESS;
  if (c.nof_ack > SRSRAN_B200_UCI_MAX_ACK_BITS || c.ri_len > 1 || c.cqi_len > SRSRAN_B200_UCI_MAX_CQI_BITS - 8 || K_segm == 0) {
    B200_LOG_ERROR("unsupported control information (nof_ack=%u ri_len=%u cqi_len=%u K=%u)", c.nof_ack, c.ri_len, c.cqi_len, K_segm);
    return B200_ERROR_INVALID_INPUTS;
  }
  // Q_prime_ri_ack (uci.c:395-418): float arithmetic, left to right as written there
  if (c.nof_ack) {
    const uint32_t x = (uint32_t)ceilf((float)c.nof_ack * M_sc * nsymb * beta_harq(c.I_offset_ack) / K_segm);
    g->Q_ack         = x < 4 * M_sc ? x : 4 * M_sc;
    g->ack_one_bit   = c.nof_ack == 1;
  }
  if (c.ri_len) {
    const uint32_t x = (uint32_t)ceilf((float)c.ri_len * M_sc * nsymb * beta_ri(c.I_offset_ri) / K_segm);
    g->Q_ri          = x < 4 * M_sc ? x : 4 * M_sc;
    g->ri_one_bit    = c.ri_len == 1;
  }
  // Q_prime_cqi (uci.c:172-190): the CRC length is added from 11 bits on, one bit before the coding scheme changes (uci.c:309)
  if (c.cqi_len) {
    const uint32_t L = c.cqi_len < 11 ? 0u : 8u;
    const uint32_t x = (uint32_t)ceilf((float)(c.cqi_len + L) * M_sc * nsymb * beta_cqi(c.I_offset_cqi) / K_segm);
    const uint32_t m = M_sc * nsymb - g->Q_ri;
    g->Q_cqi         = x < m ? x : m;
  }
  return B200_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------
// (32, O) block code, TS 36.212 Table 5.2.2.6.4-1: bit b of row i is M(i, b)
static const uint16_t kBasis[32] = {0x403, 0x607, 0x749, 0x50d, 0x48f, 0x5d3, 0x755, 0x599, 0x69b, 0x65d, 0x6e5, 0x567, 0x7a9, 0x6ab, 0x4b1, 0x6f3,
                                    0x277, 0x139, 0x0fb, 0x061, 0x445, 0x60b, 0x591, 0x717, 0x3df, 0x4e3, 0x32d, 0x3af, 0x175, 0x1fd, 0x7ff, 0x001};

// block.c:146-180: exhaustive correlation with all 2^O code words, first maximum wins, a correlation of 0 decodes to word 0
static int32_t block_decode32(const int16_t llr[32], uint8_t* data, uint32_t nbits)
{
  if (nbits > 11) nbits = 11;
  int32_t  best = 0;
  uint32_t word = 0;
  for (uint32_t guess = 0; guess < (1u << nbits); guess++) {
    int32_t corr = 0;
    for (int i = 0; i < 32; i++) corr += __builtin_parity(guess & kBasis[i]) ? (int32_t)llr[i] : -(int32_t)llr[i];
    if (corr > best) {
      best = corr;
      word = guess;
    }
  }
  for (uint32_t i = 0; i < nbits; i++) data[i] = (uint8_t)((word >> i) & 1u);
  return best;
}

// srsran_block_decode_i16 (block.c:198-215): the copies of the 32-bit code word are summed in wrap-around int16 first
static int32_t block_decode_i16(const int16_t* llr, uint32_t n, uint8_t* data, uint32_t nbits)
{
  int16_t acc[32] = {};
  for (uint32_t i = 0; i < n; i++) acc[i % 32] = (int16_t)(acc[i % 32] + llr[i]);
  return block_decode32(acc, data, nbits);
}

// srsran_uci_decode_ack_ri (uci.c:657-712) on the field's soft bits in the order the reference walks them (symbol by symbol,
// Qm bits each).  The kernel has already undone the scrambling of the repeated bit of the 1-bit form (uci.c:678-682).
static bool ack_ri_decide(const int16_t* llr, uint32_t Qprime, uint32_t Qm, uint32_t nbits, uint8_t* data)
{
  int16_t        acc[32] = {};
  const uint32_t nacc    = nbits == 1 ? Qm : nbits == 2 ? 3 * Qm : 32u;
  const uint32_t count   = Qprime * Qm;
  for (uint32_t n = 0; n < count; n++) {
    int16_t& a = acc[n % nacc];
    a          = (int16_t)(a + llr[n]);
    if (a > INT16_MAX / 2) a = INT16_MAX / 2;
    if (a < -INT16_MAX / 2) a = -INT16_MAX / 2;
  }
  const int32_t thr  = (int32_t)((count * (Qm < 4 ? 100u : Qm < 6 ? 200u : Qm < 8 ? 700u : 1000u)) / Qm);
  int32_t       corr = 0;
  if (nbits == 1) { // uci.c:505-514
    const int32_t sum = (int32_t)acc[0] + (int32_t)acc[1];
    data[0]           = sum > 0 ? 1 : 0;
    corr              = abs(sum);
  } else if (nbits == 2) { // uci.c:516-541: the three sums are int16 like the reference's locals
    const int16_t s1 = (int16_t)(acc[0] + acc[Qm + 1]), s2 = (int16_t)(acc[1] + acc[2 * Qm]), s3 = (int16_t)(acc[Qm] + acc[2 * Qm + 1]);
    data[0]          = s1 > 0 ? 1 : 0;
    data[1]          = s2 > 0 ? 1 : 0;
    const bool par   = (s3 > 0) == ((data[0] ^ data[1]) != 0);
    corr             = par ? abs((int)s1) + abs((int)s2) + abs((int)s3) : 0;
  } else {
    corr = block_decode32(acc, data, nbits);
  }
  return corr > thr;
}

// ---------------------------------------------------------------------------------------------------------------
// CQI above 11 bits: CRC-8 attached, rate-1/3 tail-biting convolutional code, rate matching of TS 36.212 5.1.4.2 (uci.c:258-287).

// srsran_rm_conv_rx_s (rm_conv.c:186-245) including its sentinel: a buffer entry equal to SRSRAN_RX_NULL counts as empty
static void conv_rm_rx(const int16_t* in, uint32_t in_len, int16_t* out, uint32_t out_len)
{
  static const uint8_t perm[32] = {1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31, 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30};
  uint8_t              inv[32];
  for (int i = 0; i < 32; i++) inv[perm[i]] = (uint8_t)i;
  const int16_t NUL   = 10000;
  const int     nrows = (int)((out_len / 3 - 1) / 32 + 1), Kp = nrows * 32;
  int           ndummy = Kp - (int)(out_len / 3);
  if (ndummy < 0) ndummy = 0;
  std::vector<int16_t> tmp((size_t)3 * Kp, NUL);
  uint32_t             k = 0;
  int                  j = 0;
  while (k < in_len) {
    const int d_i = (j % Kp) / nrows, d_j = (j % Kp) % nrows;
    if (d_j * 32 + perm[d_i] >= ndummy) {
      if (tmp[j] == NUL) tmp[j] = in[k];
      else if (in[k] != NUL) tmp[j] = (int16_t)(tmp[j] + in[k]);
      k++;
    }
    if (++j == 3 * Kp) j = 0;
  }
  for (uint32_t i = 0; i < out_len / 3; i++) {
    const int d_i = ((int)i + ndummy) / 32, d_j = ((int)i + ndummy) % 32;
    for (int s = 0; s < 3; s++) {
      const int16_t o = tmp[(size_t)Kp * s + inv[d_j] * nrows + d_i];
      out[i * 3 + s]  = o != NUL ? o : 0;
    }
  }
}

// Tail-biting Viterbi decoder for the K = 7 code with generators 133, 171, 165 (octal): a scalar restatement of the decoder the
// reference's x86 build selects, decode37_avx2_16bit (viterbi.c:129-157) over viterbi37_avx2_16bit.c:
//  * soft bits -> unsigned 16-bit symbols, x + 32767 clipped to [0, 65535] (viterbi.c:599-601, vector.c:801-814);
//  * the block is decoded three times back to back from all-zero path metrics and the middle copy is kept;
//  * branch metric = avg(avg(s0 ^ b0, s1 ^ b1), s2 ^ b2) >> 3 with the rounding of _mm256_avg_epu16, its complement is
//    8191 - metric (viterbi37_avx2_16bit.c:228-239);
//  * path metrics are uint16 with wrap-around, compared through the sign of their 16-bit difference (:243-252);
//  * the renormalisation never subtracts anything: its horizontal minimum shifts a 128-bit lane by 16 bytes, which yields zero (:279-303);
//  * the traceback reads the decision of step n + 6 for bit n ("look past tail", :140-152) although no tail was sent: the last six
//    bits of the third copy come out as zeros and the traceback enters the real decisions in state 0, whatever the best state was.
// With all of that the decided bits are the reference's on every input, not only where the CRC passes.
static void viterbi_tb(const int16_t* llr, uint32_t F, uint8_t* bits)
{
  static const uint32_t poly[3] = {0x6D, 0x4F, 0x57}; // uci.c:158
  const uint32_t        N = 3 * F;
  std::vector<uint16_t> sym((size_t)3 * F);
  for (uint32_t i = 0; i < 3 * F; i++) {
    int32_t v = (int32_t)(32767.0f + (float)llr[i]);
    v         = v < 0 ? 0 : v > 65535 ? 65535 : v;
    sym[i]    = (uint16_t)v;
  }
  uint16_t tab[3][32];
  for (uint32_t st = 0; st < 32; st++) {
    for (int k = 0; k < 3; k++) tab[k][st] = __builtin_parity((2u * st) & poly[k]) ? 65535 : 0;
  }
  std::vector<uint8_t> dec((size_t)N * 64);
  uint16_t             a[64] = {}, b[64];
  uint16_t *           old = a, *nw = b;
  auto                 avg = [](uint32_t x, uint32_t y) { return (uint16_t)((x + y + 1u) >> 1); };
  for (uint32_t t = 0; t < N; t++) {
    const uint16_t* y = &sym[(size_t)(t % F) * 3];
    for (uint32_t st = 0; st < 32; st++) {
      const uint16_t metric = (uint16_t)(avg((uint16_t)(tab[2][st] ^ y[2]), avg((uint16_t)(tab[0][st] ^ y[0]), (uint16_t)(tab[1][st] ^ y[1]))) >> 3);
      const uint16_t mm     = (uint16_t)(8191 - metric);
      const uint16_t m0 = (uint16_t)(old[st] + metric), m1 = (uint16_t)(old[st + 32] + mm);
      const uint16_t m2 = (uint16_t)(old[st] + mm), m3 = (uint16_t)(old[st + 32] + metric);
      const bool     d0 = (int16_t)(uint16_t)(m0 - m1) > 0, d1 = (int16_t)(uint16_t)(m2 - m3) > 0;
      nw[2 * st]      = d0 ? m1 : m0;
      nw[2 * st + 1]  = d1 ? m3 : m2;
      dec[(size_t)t * 64 + 2 * st]     = d0;
      dec[(size_t)t * 64 + 2 * st + 1] = d1;
    }
    uint16_t* tmp = old;
    old           = nw;
    nw            = tmp;
  }
  uint32_t state = 0; // (the six zero decisions beyond the last step have shifted the best state out)
  for (uint32_t n = N - 6; n-- > 0;) {
    const uint32_t k = dec[(size_t)(n + 6) * 64 + state];
    state            = (state >> 1) | (k << 5);
    if (n >= F && n < 2 * F) bits[n - F] = (uint8_t)k;
  }
}

// srsran_crc_checksum with SRSRAN_LTE_CRC8 (crc.h: polynomial 0x19B) over unpacked bits: remainder of the whole word
static uint32_t crc8_bits(const uint8_t* bits, uint32_t n)
{
  uint32_t r = 0;
  for (uint32_t i = 0; i < n; i++) {
    r = (r << 1) | (bits[i] & 1u);
    if (r & 0x100u) r ^= 0x19Bu;
  }
  return r & 0xFFu;
}

void uci_decide(const srsran_b200_uci_cfg_t& c, const UciGeometry& g, uint32_t Qm, const int16_t* ack_llr, const int16_t* ri_llr,
                const int16_t* cqi_llr, srsran_b200_uci_value_t* out)
{
  memset(out->ack_value, 2, sizeof(out->ack_value));
  out->ack_valid = 0;
  out->ri        = 0;
  out->cqi_crc   = 0;
  out->reserved  = 0;
  memset(out->cqi_bits, 0, sizeof(out->cqi_bits));
  if (c.nof_ack) out->ack_valid = ack_ri_decide(ack_llr, g.Q_ack, Qm, c.nof_ack, out->ack_value) ? 1 : 0;
  if (c.ri_len) {
    uint8_t ri[11] = {};
    ack_ri_decide(ri_llr, g.Q_ri, Qm, c.ri_len, ri);
    out->ri = ri[0];
  }
  if (c.cqi_len) {
    const uint32_t Q = g.Q_cqi * Qm;
    if (c.cqi_len <= 11) { // uci.c:204-216, 309-313
      block_decode_i16(cqi_llr, Q, out->cqi_bits, c.cqi_len);
      out->cqi_crc = 1;
    } else { // uci.c:258-287
      const uint32_t       F = c.cqi_len + 8;
      std::vector<int16_t> coded((size_t)3 * F);
      uint8_t              bits[SRSRAN_B200_UCI_MAX_CQI_BITS + 8];
      conv_rm_rx(cqi_llr, Q, coded.data(), 3 * F);
      viterbi_tb(coded.data(), F, bits);
      if (crc8_bits(bits, F) == 0) {
        memcpy(out->cqi_bits, bits, c.cqi_len);
        out->cqi_crc = 1;
      }
    }
  }
}

} // namespace b200

extern "C" SRSRAN_B200_API int srsran_b200_uci_decide(const srsran_b200_uci_cfg_t* uci, uint32_t Qm, uint32_t Q_prime_ack, uint32_t Q_prime_ri,
                                                     uint32_t Q_prime_cqi, const int16_t* ack_llr, const int16_t* ri_llr, const int16_t* cqi_llr,
                                                     srsran_b200_uci_value_t* out)
{
  if (!uci || !out || (Qm != 2 && Qm != 4 && Qm != 6)) return B200_ERROR_INVALID_INPUTS;
  if (uci->nof_ack > SRSRAN_B200_UCI_MAX_ACK_BITS || uci->ri_len > 1 || uci->cqi_len > SRSRAN_B200_UCI_MAX_CQI_BITS - 8) return B200_ERROR_INVALID_INPUTS;
  if ((uci->nof_ack && !ack_llr) || (uci->ri_len && !ri_llr) || (uci->cqi_len && !cqi_llr)) return B200_ERROR_INVALID_INPUTS;
  b200::UciGeometry g;
  g.Q_ack = Q_prime_ack;
  g.Q_ri  = Q_prime_ri;
  g.Q_cqi = Q_prime_cqi;
  b200::uci_decide(*uci, g, Qm, ack_llr, ri_llr, cqi_llr, out);
  out->Q_prime_ack = Q_prime_ack;
  out->Q_prime_ri  = Q_prime_ri;
  out->Q_prime_cqi = Q_prime_cqi;
  out->e_offset    = Q_prime_cqi * Qm;
  out->nof_e_bits  = 0;
  return B200_SUCCESS;
}
