// Uplink control information multiplexed into the PUSCH (TS 36.212 5.2.2.6-5.2.2.8): the host half.
//
//   coded symbols per field    Q_prime_ri_ack / Q_prime_cqi            lib/src/phy/phch/uci.c:172-190,395-418
//   beta offsets               get_beta_harq/ri/cqi_offset              lib/src/phy/phch/sch.c:55-110 (TS 36.213 Tables 8.6.3-1..3)
//   positions in the matrix    uci_ulsch_interleave_ack_gen / _ri_gen   lib/src/phy/phch/uci.c:346-393
//   HARQ-ACK / RI decision     srsran_uci_decode_ack_ri                 lib/src/phy/phch/uci.c:505-541,637-713
//   (32, O) block code         srsran_block_decode_i16                  lib/src/phy/fec/block/block.c:146-215
//   CQI up to 11 bits          decode_cqi_short                         lib/src/phy/phch/uci.c:204-216
//   order of the steps         uci_decode_ri_ack, srsran_ulsch_decode   lib/src/phy/phch/sch.c:1022-1195
//
// The per-bit work (finding the field's soft bits in the interleaver matrix, zeroing the HARQ-ACK positions, leaving the RI
// positions out of the de-interleaved stream) is done by pusch_demod_descramble_kernel<.., UCI = true>; it hands the few
// soft bits of the three fields to the host in one small array, and the functions here turn them into values.
#ifndef SRSLTE_B200_UCI_HOST_H
#define SRSLTE_B200_UCI_HOST_H

#include <stdint.h>

#include "../../include/srslte_b200.h"

namespace b200 {

struct UciGeometry {
  uint32_t Q_ack = 0, Q_ri = 0, Q_cqi = 0; // coded modulation symbols of the three fields (Q')
  bool     ack_one_bit = false, ri_one_bit = false;
};

// K_segm = C1 K1 + C2 K2 of the transport block (sch.c:1136); M_sc = 12 L_prb; nsymb = data symbols of the subframe.
// Returns B200_ERROR_INVALID_INPUTS for a reserved offset index or a field the reference does not carry.
int uci_geometry(const srsran_b200_uci_cfg_t& c, uint32_t K_segm, uint32_t M_sc, uint32_t nsymb, UciGeometry* g);

// llr: the sub-frame's compact soft-bit array written by the kernel: [Q_ack*Qm][Q_ri*Qm][Q_cqi*Qm] at the given offsets
void uci_decide(const srsran_b200_uci_cfg_t& c, const UciGeometry& g, uint32_t Qm, const int16_t* ack_llr, const int16_t* ri_llr,
                const int16_t* cqi_llr, srsran_b200_uci_value_t* out);

} // namespace b200

#endif
