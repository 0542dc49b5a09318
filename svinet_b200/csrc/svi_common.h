// svi_common.h -- host-side helpers shared by the translation units of libsvi_ls.so.
#pragma once
#include <cuda_runtime.h>

#include "../../include/svi_ls.h"

namespace svi {
// records the message returned by svi_ls_last_error() (thread-local) and returns `code`
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
}  // namespace svi

#define SVI_CK(call)                                                                            \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess)                                                                      \
      return svi::fail(SVI_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
