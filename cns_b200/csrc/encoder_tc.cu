// tcgen05 (5th-gen tensor core) version of the encoder edge kernel - placeholder until the
// UMMA pipeline lands; selecting CNS_PREC_TF32X3 / CNS_PREC_BF16 reports UNSUPPORTED loudly.
#include "common.cuh"
#include "kernels.cuh"

namespace cns {
int enc_edge_tc(const EncEdgeArgs&, int, const void*, int, cudaStream_t) {
  set_error("tensor-core encoder not built into this library");
  return CNS_ERR_UNSUPPORTED;
}
}  // namespace cns
