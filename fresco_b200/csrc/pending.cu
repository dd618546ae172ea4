// Entry points declared in include/fresco_b200.h whose kernels are not written yet.
#include "fresco_internal.h"
using namespace fresco;

extern "C" int fresco_gram_sign(const void*, const float*, void*, float*, int, int, int, float, void*) {
  return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_gram_sign: not implemented yet");
}
extern "C" int fresco_gram_grad(const void*, const void*, const float*, float*, int, int, int, float, void*, size_t,
                                void*) {
  return set_error(FRESCO_ERR_UNSUPPORTED, "fresco_gram_grad: not implemented yet");
}
extern "C" size_t fresco_gram_grad_workspace_bytes(int, int, int) { return 0; }
extern "C" int gmflow_global_corr_softmax(const float*, const float*, float*, int, int, int, int, int, void*, size_t,
                                          void*) {
  return set_error(FRESCO_ERR_UNSUPPORTED, "gmflow_global_corr_softmax: not implemented yet");
}
extern "C" size_t fresco_gmflow_corr_workspace_bytes(int, int, int, int) { return 0; }
