/* abi_consumer.c — a plain-C client of include/sfron_b200.h (test infrastructure, built and run by
 * tests/test_host_logic.py::test_c_abi_from_plain_c).  Proves the header is valid C99 (no C++ in the boundary), that the
 * shared library links from C, and that on a machine without a CUDA device every compute entry point FAILS LOUDLY with
 * SFR_ERR_NO_DEVICE instead of falling back to host code.  With a device present it only reports what it found. */
#include <stdio.h>
#include <string.h>
#include "sfron_b200.h"

int main(void) {
  int sms = -1, major = -1, minor = -1;
  float buf[8] = {0};
  unsigned char mask[8] = {0};
  unsigned long long zero_count = 0;
  int rc_info, rc_k1, rc_k2;

  if (sfr_abi_version() <= 0) return 10;
  if (sfr_error_string(SFR_OK) == NULL || sfr_error_string(SFR_ERR_NO_DEVICE) == NULL) return 11;
  printf("abi %d\n", sfr_abi_version());
  printf("sizeof sfr_update_args %zu sfr_select_state %zu sfr_peer_buf %zu sfr_peer_geom %zu\n", sizeof(sfr_update_args),
         sizeof(sfr_select_state), sizeof(sfr_peer_buf), sizeof(sfr_peer_geom));

  rc_info = sfr_device_info(&sms, &major, &minor);
  if (rc_info == SFR_OK) {                     /* a GPU box: nothing more to prove here, the -m gpu tests do the rest */
    printf("device sm_count %d cc %d.%d\n", sms, major, minor);
    return 0;
  }
  /* no device: host pointers would be wrong arguments on a GPU box, but here the call must stop at the device check */
  rc_k1 = sfr_fisher_accum(buf, buf, SFR_F32, 1, 8, 8, 1.0f, NULL, 0.0f, NULL);
  rc_k2 = sfr_ratio_mask(buf, buf, 8, 1.0f, 1e-15f, mask, &zero_count, NULL);
  printf("no device: info %d k1 %d k2a %d (%s)\n", rc_info, rc_k1, rc_k2, sfr_error_string(rc_k1));
  if (rc_info != SFR_ERR_NO_DEVICE || rc_k1 != SFR_ERR_NO_DEVICE || rc_k2 != SFR_ERR_NO_DEVICE) return 12;
  if (strstr(sfr_error_string(SFR_ERR_NO_DEVICE), "no CPU fallback") == NULL) return 13;
  if (buf[0] != 0.0f || mask[0] != 0) return 14; /* nothing was computed on the host */
  return 0;
}
