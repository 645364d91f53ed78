// Graph construction on device (cns/midend/graph_gen.py).
//
//  cns_graph_mask   : per-frame sub-graph masking of GraphGenerator._get_graph
//                     (graph_gen.py:199-228) for a whole batch of scenes at once: order-preserving
//                     stream compaction of the keypoint->cluster edges and regeneration of the dense
//                     cluster graph over the surviving clusters.  Integer work, bit-exact.
//  cns_label_nearest: 1-NN keypoint -> exemplar assignment, the only nearest-neighbour search in the
//                     reference's clustering (final step of sklearn's AffinityPropagation).
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

__global__ void i64_copy_i32_kernel(const int64_t* __restrict__ in, int64_t n, int* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}

// flag[k] = visible[member_order[k]]  (member_order = static l0_to_l1 j, cluster-major)
__global__ void member_flag_kernel(const int64_t* __restrict__ member_order, const uint8_t* __restrict__ visible,
                                   int64_t n, int* __restrict__ flag) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) flag[k] = visible[member_order[k]] ? 1 : 0;
}

// cluster survives iff at least one member is visible: pos[mptr[c+1]] - pos[mptr[c]] > 0
__global__ void cluster_alive_kernel(const int* __restrict__ mptr, const int* __restrict__ pos, int64_t n_clusters,
                                     uint8_t* __restrict__ cluster_mask, int* __restrict__ alive) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clusters) return;
  int a = (pos[mptr[c + 1]] - pos[mptr[c]]) > 0;
  cluster_mask[c] = (uint8_t)a;
  alive[c] = a;
}

__global__ void e01_scatter_kernel(const int64_t* __restrict__ member_order, const int64_t* __restrict__ belong,
                                   const int* __restrict__ flag, const int* __restrict__ pos, int64_t n,
                                   int64_t* __restrict__ e01_j, int64_t* __restrict__ e01_i,
                                   int64_t* __restrict__ counts_out) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  if (k == n) { counts_out[0] = pos[n]; return; }
  if (flag[k]) {
    int64_t j = member_order[k];
    e01_j[pos[k]] = j;
    e01_i[pos[k]] = belong[j];
  }
}

// m2[g] = (#alive clusters of graph g)^2
__global__ void graph_alive_sq_kernel(const int64_t* __restrict__ gptr, const int* __restrict__ arank, int64_t n_graphs,
                                      int* __restrict__ m2) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_graphs) return;
  int m = arank[gptr[g + 1]] - arank[gptr[g]];
  m2[g] = m * m;
}

// compact list of alive cluster ids (global), in order
__global__ void alive_list_kernel(const int* __restrict__ alive, const int* __restrict__ arank, int64_t n_clusters,
                                  int* __restrict__ list) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_clusters && alive[c]) list[arank[c]] = (int)c;
}

// one CTA per graph: edges (j = v_a, i = v_b) at off + a*m + b : source-major like
// gen_dense_connect_edges (graph_gen.py:7-21) followed by the boolean filter (graph_gen.py:62).
__global__ void e11_emit_kernel(const int64_t* __restrict__ gptr, const int* __restrict__ arank,
                                const int* __restrict__ list, const int* __restrict__ eoff, int64_t n_graphs,
                                int64_t* __restrict__ e11, int64_t stride, int64_t* __restrict__ counts_out) {
  int g = blockIdx.x;
  int a0 = arank[gptr[g]];
  int m = arank[gptr[g + 1]] - a0;
  int off = eoff[g];
  for (int t = threadIdx.x; t < m * m; t += blockDim.x) {
    int a = t / m, b = t - a * m;
    e11[off + t] = list[a0 + a];
    e11[stride + off + t] = list[a0 + b];
  }
  if (g == 0 && threadIdx.x == 0) counts_out[1] = eoff[n_graphs];
}

// labels[p] = argmax_k -(|x|^2 - 2 x.e_k + |e_k|^2) in the operation order of
// sklearn.metrics.euclidean_distances (d = -2*dot; d += xx; d += yy; max(d,0)); first max wins.
__global__ void label_nearest_kernel(const double* __restrict__ pts, const int64_t* __restrict__ point_graph,
                                     int64_t n_points, const double* __restrict__ ex,
                                     const int64_t* __restrict__ xptr, int* __restrict__ labels) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_points) return;
  int64_t g = point_graph ? point_graph[p] : 0;
  int64_t k0 = xptr[g], k1 = xptr[g + 1];
  double x0 = pts[2 * p], x1 = pts[2 * p + 1];
  double xx = __dadd_rn(__dmul_rn(x0, x0), __dmul_rn(x1, x1));
  double best = -INFINITY;
  int arg = 0;
  for (int64_t k = k0; k < k1; ++k) {
    double y0 = ex[2 * k], y1 = ex[2 * k + 1];
    double yy = __dadd_rn(__dmul_rn(y0, y0), __dmul_rn(y1, y1));
    double dot = __dadd_rn(__dmul_rn(x0, y0), __dmul_rn(x1, y1));
    double d = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), xx), yy);
    d = fmax(d, 0.0);
    double sim = -d;
    if (sim > best) { best = sim; arg = (int)(k - k0); }
  }
  labels[p] = arg;
}

}  // namespace cns

using namespace cns;

extern "C" {

size_t cns_graph_mask_workspace_bytes(int64_t n_nodes, int64_t n_clusters, int64_t n_graphs) {
  size_t b = 0;
  b += align_up((size_t)(n_clusters + 2) * 4) * 5;   // csize, mptr, alive, arank, list
  b += align_up((size_t)(n_nodes + 2) * 4) * 2;      // flag, pos
  b += align_up((size_t)(n_graphs + 2) * 4) * 2;     // m2, eoff
  b += scan_scratch_bytes(n_nodes + n_clusters + n_graphs + 3);
  return b + 256;
}

int cns_graph_mask(const int64_t* belong, const int64_t* member_order, const int64_t* cluster_size,
                   const int64_t* cluster_graph_ptr, const uint8_t* visible, int64_t n_nodes, int64_t n_clusters,
                   int64_t n_graphs, int64_t* e01_j, int64_t* e01_i, int64_t* e11_cur, int64_t e11_capacity,
                   uint8_t* cluster_mask, int64_t* counts_out, void* workspace, size_t workspace_bytes,
                   void* stream) {
  CNS_REQUIRE(n_nodes >= 0 && n_clusters >= 0 && n_graphs >= 0, "negative size");
  CNS_REQUIRE(n_nodes < (1ll << 31) && n_clusters < (1ll << 31), "sizes must fit int32");
  CNS_REQUIRE(belong && member_order && cluster_size && cluster_graph_ptr && visible && e01_j && e01_i &&
              cluster_mask && counts_out && workspace, "NULL pointer");
  if (workspace_bytes < cns_graph_mask_workspace_bytes(n_nodes, n_clusters, n_graphs)) {
    set_error("cns_graph_mask: workspace too small");
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* w = (char*)workspace;
  auto take = [&](size_t n) { int* r = (int*)w; w += align_up(n * 4); return r; };
  int* csize = take(n_clusters + 2); int* mptr = take(n_clusters + 2); int* alive = take(n_clusters + 2);
  int* arank = take(n_clusters + 2); int* list = take(n_clusters + 2);
  int* flag = take(n_nodes + 2); int* pos = take(n_nodes + 2);
  int* m2 = take(n_graphs + 2); int* eoff = take(n_graphs + 2);
  int* scratch = (int*)w;
  int rc;
  const int T = 256;
  if (n_clusters > 0) {
    i64_copy_i32_kernel<<<ceil_div(n_clusters, T), T, 0, s>>>(cluster_size, n_clusters, csize);
    CNS_LAUNCH_CHECK();
  }
  if ((rc = exclusive_scan_i32(csize, n_clusters, mptr, scratch, s)) != CNS_OK) return rc;
  if (n_nodes > 0) {
    member_flag_kernel<<<ceil_div(n_nodes, T), T, 0, s>>>(member_order, visible, n_nodes, flag);
    CNS_LAUNCH_CHECK();
  }
  if ((rc = exclusive_scan_i32(flag, n_nodes, pos, scratch, s)) != CNS_OK) return rc;
  if (n_clusters > 0) {
    cluster_alive_kernel<<<ceil_div(n_clusters, T), T, 0, s>>>(mptr, pos, n_clusters, cluster_mask, alive);
    CNS_LAUNCH_CHECK();
  }
  e01_scatter_kernel<<<ceil_div(n_nodes + 1, T), T, 0, s>>>(member_order, belong, flag, pos, n_nodes, e01_j, e01_i,
                                                            counts_out);
  CNS_LAUNCH_CHECK();
  if (e11_cur == nullptr) {   // the caller only needs cluster_mask: the cluster graph is implied (CNS_BATCH_L1_CUR_COMPLETE)
    CNS_CUDA_CHECK(cudaMemsetAsync(counts_out + 1, 0, sizeof(int64_t), s));
    return CNS_OK;
  }
  if ((rc = exclusive_scan_i32(alive, n_clusters, arank, scratch, s)) != CNS_OK) return rc;
  if (n_clusters > 0) {
    alive_list_kernel<<<ceil_div(n_clusters, T), T, 0, s>>>(alive, arank, n_clusters, list);
    CNS_LAUNCH_CHECK();
  }
  if (n_graphs > 0) {
    graph_alive_sq_kernel<<<ceil_div(n_graphs, T), T, 0, s>>>(cluster_graph_ptr, arank, n_graphs, m2);
    CNS_LAUNCH_CHECK();
  }
  if ((rc = exclusive_scan_i32(m2, n_graphs, eoff, scratch, s)) != CNS_OK) return rc;
  if (n_graphs > 0) {
    e11_emit_kernel<<<(unsigned)n_graphs, 128, 0, s>>>(cluster_graph_ptr, arank, list, eoff, n_graphs, e11_cur,
                                                       e11_capacity, counts_out);
    CNS_LAUNCH_CHECK();
  } else {
    CNS_CUDA_CHECK(cudaMemsetAsync(counts_out + 1, 0, sizeof(int64_t), s));
  }
  return CNS_OK;
}

int cns_frame_mask(const int64_t* member_order, const int32_t* member_ptr, const uint8_t* visible, int64_t n_nodes,
                   int64_t n_clusters, int64_t* e01_j, int64_t* e01_i, uint8_t* cluster_mask, int32_t* clu_rowptr,
                   void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(n_nodes >= 0 && n_clusters >= 0 && n_nodes < (1ll << 31) && n_clusters < (1ll << 31), "bad size");
  CNS_REQUIRE(member_order && member_ptr && visible && e01_j && e01_i && cluster_mask && clu_rowptr && workspace, "NULL pointer");
  CNS_REQUIRE(workspace_bytes >= (size_t)(n_clusters + 1) * sizeof(int), "cns_frame_mask: workspace too small");
  return frame_mask(member_order, member_ptr, visible, n_clusters, e01_j, e01_i, cluster_mask, clu_rowptr, (int*)workspace,
                    (cudaStream_t)stream);
}

int cns_label_nearest(const double* points, const int64_t* point_graph, int64_t n_points,
                      const double* exemplars, const int64_t* exemplar_ptr, int32_t* labels, void* stream) {
  CNS_REQUIRE(points && exemplars && exemplar_ptr && labels, "NULL pointer");
  if (n_points == 0) return CNS_OK;
  label_nearest_kernel<<<ceil_div(n_points, 128), 128, 0, (cudaStream_t)stream>>>(points, point_graph, n_points,
                                                                               exemplars, exemplar_ptr, labels);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // extern "C"
