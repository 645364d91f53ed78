// Workspace / saved-activation layout shared by forward.cu and backward.cu.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

// Bump allocator over a caller-provided buffer; with base == nullptr it only measures.
struct Arena {
  char* base; size_t off;
  explicit Arena(void* b) : base((char*)b), off(0) {}
  template <typename T> T* take(size_t n) {
    size_t o = off;
    off += align_up(n * sizeof(T));
    return base ? reinterpret_cast<T*>(base + o) : nullptr;
  }
};

// Everything the backward pass needs from the forward pass (cluster-level only: per-edge encoder
// tensors are recomputed).  Lives in the caller's `save` buffer (cns_forward_save_bytes).
struct SaveBuffers {
  int* clu_rowptr; int* graph_ptr; int* cur_rowptr; int* cur_src; int* tar_rowptr; int* tar_src;
  int* arg0; int* argg; int* argk; int* arg1;        // arg-max source per (row, channel) of the 4 edge convs
  float* pos_clu; float* clu_feat; float* x_clu;
  float* conv0; float* gn0; float* stats0; float* x1;
  float* g; float* u; float* hr; float* ht;
  float* conv1; float* gn1; float* stats1; float* x2;
  float* gate; float* x_scene; float* hid;
};

inline void carve_save(Arena& ar, const cns_batch* b, SaveBuffers& v) {
  const int64_t C = b->n_clusters, B = b->n_graphs;
  v.clu_rowptr = ar.take<int>(C + 1);
  v.graph_ptr = ar.take<int>(B + 1);
  v.cur_rowptr = ar.take<int>(C + 1);
  v.cur_src = ar.take<int>(b->n_e11_cur + 1);
  v.tar_rowptr = ar.take<int>(C + 1);
  v.tar_src = ar.take<int>(b->n_e11_tar + 1);
  v.arg0 = ar.take<int>(C * H); v.argg = ar.take<int>(C * 2 * H); v.argk = ar.take<int>(C * H); v.arg1 = ar.take<int>(C * H);
  v.pos_clu = ar.take<float>(C * 2 + 2);
  v.clu_feat = ar.take<float>(2 * C * H);
  v.x_clu = ar.take<float>(C * H);
  v.conv0 = ar.take<float>(C * H); v.gn0 = ar.take<float>(C * H); v.stats0 = ar.take<float>(B * 2 * H); v.x1 = ar.take<float>(C * H);
  v.g = ar.take<float>(C * 2 * H); v.u = ar.take<float>(C * H); v.hr = ar.take<float>(C * H); v.ht = ar.take<float>(C * H);
  v.conv1 = ar.take<float>(C * H); v.gn1 = ar.take<float>(C * H); v.stats1 = ar.take<float>(B * 2 * H); v.x2 = ar.take<float>(C * H);
  v.gate = ar.take<float>(C + 1); v.x_scene = ar.take<float>(B * H); v.hid = ar.take<float>(B * 2 * H);
}

}  // namespace cns
