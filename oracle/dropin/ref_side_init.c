/* TEST INFRASTRUCTURE ONLY (drop-in proof, see oracle/Makefile target `dropin`).
 *
 * The reference's own test programs are linked against libsrslte_b200.so FIRST and against the reference's remaining
 * sources (oracle/_ref/libsrsref.so) second, so every symbol this repo's library exports replaces the reference's
 * implementation -- also for calls made from inside the reference's sch.c / pusch.c / ofdm.c / turbocoder.c.
 * One reference function keeps private state that the TRANSMIT side of the test programs needs:
 * srsran_rm_turbo_gentables() (lib/src/phy/fec/turbo/rm_turbo.c:276-318) also fills the rate-MATCHING tables used by
 * srsran_rm_turbo_tx_lut, which is not part of the receive path this repo replaces.  Because the replacement
 * srsran_rm_turbo_gentables wins the symbol lookup, the reference's is called here once, by address, before main().
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>

__attribute__((constructor)) static void ref_side_init(void)
{
  void* h = dlopen("libsrsref.so", RTLD_NOW | RTLD_NOLOAD);
  if (!h) {
    h = dlopen("libsrsref.so", RTLD_NOW);
  }
  if (!h) {
    fprintf(stderr, "ref_side_init: %s\n", dlerror());
    exit(3);
  }
  void (*gen)(void) = (void (*)(void))dlsym(h, "srsran_rm_turbo_gentables");
  if (!gen) {
    fprintf(stderr, "ref_side_init: reference srsran_rm_turbo_gentables not found\n");
    exit(3);
  }
  gen();
}
