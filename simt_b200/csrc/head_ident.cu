// Plain cross entropy (T = NULL -> identity) instantiations of the fused head kernel: tools/trainV2_simt.py:394-395,
// tools/trainV1_warmup.py:222-224 and CrossEntropy2d(is_softmax=True) (utils/loss.py:35-36).  A separate translation
// unit so that the two sets of instantiations compile in parallel.
#include "head_kernel.cuh"

namespace simt {

int dispatch_modes_ident(int mode, int label_bytes, const HeadArgs& A, const Plan& P, cudaStream_t st, int* grid_out) {
  if (mode == MODE_PLACE) return SIMT_EINVAL;
  return dispatch_modes<true>(mode, label_bytes, A, P, st, grid_out);
}

}  // namespace simt
