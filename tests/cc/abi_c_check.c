/* The C ABI from plain C (gcc -std=c99, no C++, no torch): both headers parse as C, every declared entry point
 * links, and without a CUDA device the constructors fail with SVI_ERR_CUDA and a message instead of falling
 * back to anything.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <string.h>

#include "svi_fa2.h"
#include "svi_ls.h"

int main(void) {
  uint32_t links[4] = {0, 1, 1, 2};
  svi_ls_config cfg;
  svi_ls *h = NULL;
  svi_fa2_config fc;
  svi_fa2 *f = NULL;
  int rc;
  if (svi_ls_abi_version() != SVI_LS_ABI_VERSION) return 10;
  memset(&cfg, 0, sizeof cfg);
  cfg.n = 3; cfg.k = 4; cfg.nlinks = 2; cfg.alpha = 0.25; cfg.eta0 = 1; cfg.eta1 = 1; cfg.ones = 2; cfg.device = -1;
  cfg.node_begin = 0; cfg.node_end = 3;
  rc = svi_ls_create(&cfg, links, NULL, &h);
  printf("svi_ls_create rc=%d msg=%s\n", rc, rc ? svi_ls_last_error() : "");
  if (rc == SVI_OK) {
    double gamma[12], lambda[8];
    int i;
    for (i = 0; i < 12; ++i) gamma[i] = 0.3 + 0.1 * i;
    for (i = 0; i < 8; ++i) lambda[i] = 1.0;
    if (svi_ls_set_state(h, gamma, lambda) || svi_ls_step(h, 0, 1, 1) || svi_ls_get_state(h, gamma, lambda)) return 11;
    svi_ls_destroy(h);
  } else if (rc != SVI_ERR_CUDA || !strlen(svi_ls_last_error())) {
    return 12;
  }
  svi_fa2_default_config(&fc, 10, 4);
  if (fc.m_sets != 10 || fc.online_iterations != 50 || fc.eager_blend != 0) return 13;
  rc = svi_fa2_create(&fc, &f);
  printf("svi_fa2_create rc=%d msg=%s\n", rc, rc ? svi_ls_last_error() : "");
  if (rc == SVI_OK) svi_fa2_destroy(f);
  else if (rc != SVI_ERR_CUDA) return 14;
  /* argument validation does not need a device */
  if (svi_ls_step(NULL, 0, 0, 0) != SVI_ERR_INVALID) return 15;
  if (svi_fa2_step(NULL, 0, 0, 0, 0, NULL) != SVI_ERR_INVALID) return 16;
  return 0;
}
