// Device side of the correspondence-graph construction (cns/midend/graph_gen.py:104-196) around the Affinity
// Propagation loop of affinity.cu, for a batch of G target point sets at once:
//
//   cns_ap_similarity   S = -||x_i - x_j||^2 as scikit-learn's euclidean_distances(X, squared=True) rounds it
//                       (fl(fl(-2 x_i.x_j + |x_i|^2) + |x_j|^2), the dot product accumulated k = 0 then k = 1 with one
//                       FMA like the BLAS micro-kernel, |x|^2 without FMA like numpy's einsum, clamped at 0, zero diagonal)
//   cns_segment_median  exact np.median of fp64 segments (radix select of the two middle order statistics, 8 passes of
//                       8 bits over order-preserving keys) - the AP preference is median(S)
//   cns_ap_prepare      preference on the diagonal, then sklearn's degeneracy noise
//                       S += (eps * S + tiny * 100) * noise   with the noise drawn by the HOST numpy stream
//   cns_ap_labels       tail of sklearn's _affinity_propagation: exemplar list, assignment argmax_k S[i, I_k],
//                       exemplar refinement (argmax of the member-submatrix column sums, rows added in ascending order
//                       like numpy's axis-0 sum), re-assignment, labels = rank of the exemplar's point index
//   cns_graph_layout    GraphGenerator.__init__ after clustering (graph_gen.py:146-196): cluster-major member order
//                       (= l0_to_l1_edge_index), member counts, centre = member nearest to the per-cluster coordinate
//                       median (np.median: mean of the two middle values; first minimum wins)
//
// All fp64 arithmetic uses explicit round-to-nearest intrinsics in numpy's operation order (no contraction), so the
// outputs are bit-identical with the host path; integer outputs are exact by construction.
#include <float.h>
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

namespace {

// ------------------------------------------------------------------------------------------- similarity
__global__ void __launch_bounds__(256)
gb_similarity_kernel(const double* __restrict__ pts, const int64_t* __restrict__ pt_off, const int64_t* __restrict__ mat_off,
                     const int* __restrict__ row_graph, int64_t n_rows, double* __restrict__ S) {
  const int64_t row = blockIdx.x;
  if (row >= n_rows) return;
  const int g = row_graph[row];
  const int64_t p0 = pt_off[g], n = pt_off[g + 1] - p0, i = row - p0;
  const double xi0 = pts[row * 2], xi1 = pts[row * 2 + 1];
  const double xxi = __dadd_rn(__dmul_rn(xi0, xi0), __dmul_rn(xi1, xi1));
  double* out = S + mat_off[g] + i * n;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double xj0 = pts[(p0 + j) * 2], xj1 = pts[(p0 + j) * 2 + 1];
    const double xxj = __dadd_rn(__dmul_rn(xj0, xj0), __dmul_rn(xj1, xj1));
    const double dot = __fma_rn(xi1, xj1, __dmul_rn(xi0, xj0));
    double d = __dadd_rn(__dadd_rn(__dmul_rn(-2.0, dot), xxi), xxj);
    d = fmax(d, 0.0);
    if (j == i) d = 0.0;
    out[j] = -d;
  }
}

// ------------------------------------------------------------------------------------------- radix select
__device__ __forceinline__ unsigned long long f64_key(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

struct SelState { unsigned long long prefix; long long k; };

// query q = 2 * segment + {0: lower middle, 1: upper middle}
__global__ void sel_init_kernel(const int64_t* __restrict__ seg_off, int64_t n_seg, SelState* __restrict__ st,
                                unsigned long long* __restrict__ hist) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < 2 * n_seg) {
    const int64_t m = seg_off[(q >> 1) + 1] - seg_off[q >> 1];
    st[q].prefix = 0ull;
    st[q].k = (q & 1) ? m / 2 : (m - 1) / 2;
  }
  for (int64_t i = q; i < 2 * n_seg * 256; i += (int64_t)gridDim.x * blockDim.x) hist[i] = 0ull;
}

__global__ void __launch_bounds__(256)
sel_hist_kernel(const double* __restrict__ v, const int64_t* __restrict__ seg_off, const SelState* __restrict__ st,
                int pass, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int h[2][256];
  const int seg = blockIdx.y;
  h[0][threadIdx.x] = 0u; h[1][threadIdx.x] = 0u;
  __syncthreads();
  const int64_t beg = seg_off[seg], end = seg_off[seg + 1];
  const unsigned long long p0 = st[2 * seg].prefix, p1 = st[2 * seg + 1].prefix;
  const int hi_shift = 64 - 8 * pass, bin_shift = 56 - 8 * pass;
  // the block's grid-stride slices flush their shared histogram before 2^32 elements could have been counted
  for (int64_t i = beg + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long key = f64_key(v[i]);
    const unsigned long long top = pass ? (key >> hi_shift) : 0ull;
    const unsigned int bin = (unsigned int)(key >> bin_shift) & 255u;
    if (top == p0) atomicAdd(&h[0][bin], 1u);
    if (top == p1) atomicAdd(&h[1][bin], 1u);
  }
  __syncthreads();
  if (h[0][threadIdx.x]) atomicAdd(&hist[(size_t)(2 * seg) * 256 + threadIdx.x], (unsigned long long)h[0][threadIdx.x]);
  if (h[1][threadIdx.x]) atomicAdd(&hist[(size_t)(2 * seg + 1) * 256 + threadIdx.x], (unsigned long long)h[1][threadIdx.x]);
}

__global__ void sel_pick_kernel(int64_t n_q, SelState* __restrict__ st, unsigned long long* __restrict__ hist) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_q) return;
  unsigned long long* h = hist + (size_t)q * 256;
  long long k = st[q].k;
  int bin = 255;
  for (int b = 0; b < 256; ++b) {
    const long long c = (long long)h[b];
    if (k < c) { bin = b; break; }
    k -= c;
  }
  for (int b = 0; b < 256; ++b) h[b] = 0ull;
  st[q].prefix = (st[q].prefix << 8) | (unsigned long long)bin;
  st[q].k = k;
}

__global__ void sel_finish_kernel(const int64_t* __restrict__ seg_off, int64_t n_seg, const SelState* __restrict__ st,
                                  double* __restrict__ out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg) return;
  const int64_t m = seg_off[s + 1] - seg_off[s];
  if (m <= 0) { out[s] = __longlong_as_double(0x7ff8000000000000ll); return; }
  const double a = key_f64(st[2 * s].prefix), b = key_f64(st[2 * s + 1].prefix);
  out[s] = (m & 1) ? a : __ddiv_rn(__dadd_rn(a, b), 2.0);            // np.median: mean of the two middle values
}

// ------------------------------------------------------------------------------------------- preference + noise
__global__ void __launch_bounds__(256)
gb_prepare_kernel(double* __restrict__ S, const double* __restrict__ noise, const double* __restrict__ pref,
                  const int64_t* __restrict__ pt_off, const int64_t* __restrict__ mat_off, const int* __restrict__ row_graph,
                  int64_t n_rows) {
  const int64_t row = blockIdx.x;
  if (row >= n_rows) return;
  const int g = row_graph[row];
  const int64_t n = pt_off[g + 1] - pt_off[g], i = row - pt_off[g];
  double* s = S + mat_off[g] + i * n;
  const double* z = noise + mat_off[g] + i * n;
  const double pf = pref[g];
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = (j == i) ? pf : s[j];
    // S += (np.finfo(float64).eps * S + np.finfo(float64).tiny * 100) * random_state.standard_normal((n, n))
    const double t = __dadd_rn(__dmul_rn(DBL_EPSILON, v), DBL_MIN * 100.0);
    s[j] = __dadd_rn(v, __dmul_rn(t, z[j]));
  }
}

// ------------------------------------------------------------------------------------------- labels from exemplars
// One CTA per problem.  ex (n) int32 scratch: exemplar point indices (first K entries); lab (n) int32 scratch.
__global__ void __launch_bounds__(256)
gb_labels_kernel(const double* __restrict__ S, const unsigned char* __restrict__ E, const int64_t* __restrict__ pt_off,
                 const int64_t* __restrict__ mat_off, int* __restrict__ ex_all, int* __restrict__ lab_all,
                 int* __restrict__ labels_out, int* __restrict__ exemplars_out, int* __restrict__ n_clusters_out) {
  __shared__ int s_k;
  const int g = blockIdx.x;
  const int64_t p0 = pt_off[g], n = pt_off[g + 1] - p0;
  const double* Sg = S + mat_off[g];
  const unsigned char* Eg = E + p0;
  int* ex = ex_all + p0;
  int* lab = lab_all + p0;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  // I = flatnonzero(E), ascending: warp 0 compacts with ballots
  if (warp == 0) {
    int k = 0;
    for (int64_t b = 0; b < n; b += 32) {
      const int64_t i = b + lane;
      const bool e = i < n && Eg[i];
      const unsigned m = __ballot_sync(0xffffffffu, e);
      if (e) ex[k + __popc(m & ((1u << lane) - 1u))] = (int)i;
      k += __popc(m);
    }
    if (lane == 0) s_k = k;
  }
  __syncthreads();
  const int K = s_k;
  if (K == 0) {      // sklearn: labels = [-1] * n, no centres
    for (int64_t i = tid; i < n; i += nt) labels_out[p0 + i] = -1;
    if (tid == 0) n_clusters_out[g] = 0;
    return;
  }
  auto assign = [&]() {
    // c = argmax(S[:, I], axis=1) (first maximum); c[I] = arange(K)
    for (int64_t i = tid; i < n; i += nt) {
      const double* row = Sg + i * n;
      double best = row[ex[0]];
      int arg = 0;
      for (int k = 1; k < K; ++k) {
        const double v = row[ex[k]];
        if (v > best) { best = v; arg = k; }
      }
      lab[i] = arg;
    }
    __syncthreads();
    for (int k = tid; k < K; k += nt) lab[ex[k]] = k;
    __syncthreads();
  };
  assign();
  // refine: for cluster k with members ii (ascending): j = argmax_j sum_r S[ii[r], ii[j]] (rows added in order); I[k] = ii[j]
  // one warp per cluster; lanes own member columns
  for (int k = warp; k < K; k += nw) {
    double best = -INFINITY;
    long long best_pos = 0x7fffffffffffffffll;     // position among the members (ties: first)
    int best_pt = -1;
    long long pos_base = 0;
    for (int64_t b = 0; b < n; b += 32) {
      const int64_t j = b + lane;
      const bool mine = j < n && lab[j] == k;
      const unsigned m = __ballot_sync(0xffffffffu, mine);
      if (mine) {
        double s = 0.0;
        bool first = true;
        for (int64_t r = 0; r < n; ++r) {
          if (lab[r] != k) continue;
          const double v = Sg[r * n + j];
          s = first ? v : __dadd_rn(s, v);
          first = false;
        }
        const long long pos = pos_base + __popc(m & ((1u << lane) - 1u));
        if (s > best || (s == best && pos < best_pos)) { best = s; best_pos = pos; best_pt = (int)j; }
      }
      pos_base += __popc(m);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, off);
      const long long op = __shfl_xor_sync(0xffffffffu, best_pos, off);
      const int ot = __shfl_xor_sync(0xffffffffu, best_pt, off);
      if (ov > best || (ov == best && op < best_pos)) { best = ov; best_pos = op; best_pt = ot; }
    }
    __syncwarp();
    if (lane == 0) exemplars_out[p0 + k] = best_pt;       // staged: ex is still being read by other warps' members
  }
  __syncthreads();
  for (int k = tid; k < K; k += nt) ex[k] = exemplars_out[p0 + k];
  __syncthreads();
  assign();
  // labels = I[c]; centres = unique(labels) (ascending point index); final label = rank of the exemplar's index.
  // K is small (~0.8 sqrt(n)): rank by counting
  for (int64_t i = tid; i < n; i += nt) {
    const int e = ex[lab[i]];
    int r = 0;
    for (int k = 0; k < K; ++k) r += ex[k] < e;
    labels_out[p0 + i] = r;
  }
  __syncthreads();
  for (int k = tid; k < K; k += nt) {
    const int e = ex[k];
    int r = 0;
    for (int q = 0; q < K; ++q) r += ex[q] < e;
    exemplars_out[p0 + r] = e;
  }
  if (tid == 0) n_clusters_out[g] = K;
}

// ------------------------------------------------------------------------------------------- layout
// One CTA per problem; labels are gapless ranks 0..K-1.  clu_off[g] = first cluster slot of problem g in the outputs.
__global__ void __launch_bounds__(256)
gb_layout_kernel(const double* __restrict__ pts, const int* __restrict__ labels, const int64_t* __restrict__ pt_off,
                 const int64_t* __restrict__ clu_off, int* __restrict__ member_order, int* __restrict__ cluster_size,
                 int* __restrict__ centers) {
  const int g = blockIdx.x;
  const int64_t p0 = pt_off[g], n = pt_off[g + 1] - p0;
  const int64_t c0 = clu_off[g];
  const int K = (int)(clu_off[g + 1] - c0);
  const int* lab = labels + p0;
  int* order = member_order + p0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  // sizes (needed for the member offsets): one warp per cluster counts with ballots
  for (int k = warp; k < K; k += nw) {
    int cnt = 0;
    for (int64_t b = 0; b < n; b += 32) cnt += __popc(__ballot_sync(0xffffffffu, b + lane < n && lab[b + lane] == k));
    if (lane == 0) cluster_size[c0 + k] = cnt;
  }
  __syncthreads();
  for (int k = warp; k < K; k += nw) {
    int start = 0;
    for (int q = 0; q < k; ++q) start += cluster_size[c0 + q];
    const int m = cluster_size[c0 + k];
    // members in ascending point order
    int w = 0;
    for (int64_t b = 0; b < n; b += 32) {
      const int64_t i = b + lane;
      const bool mine = i < n && lab[i] == k;
      const unsigned bal = __ballot_sync(0xffffffffu, mine);
      if (mine) order[start + w + __popc(bal & ((1u << lane) - 1u))] = (int)i;
      w += __popc(bal);
    }
    __syncwarp();
    // coordinate-wise median by rank counting (ties ordered by member position): the two middle order statistics
    double med[2];
    const int r_lo = (m - 1) / 2, r_hi = m / 2;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      double v_lo = 0.0, v_hi = 0.0;
      for (int a = lane; a < m; a += 32) {
        const double va = pts[(p0 + order[start + a]) * 2 + d];
        int rank = 0;
        for (int b2 = 0; b2 < m; ++b2) {
          const double vb = pts[(p0 + order[start + b2]) * 2 + d];
          rank += (vb < va) || (vb == va && b2 < a);
        }
        if (rank == r_lo) v_lo = va;
        if (rank == r_hi) v_hi = va;
      }
      // exactly one lane holds each of them
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        v_lo = __dadd_rn(v_lo, __shfl_xor_sync(0xffffffffu, v_lo, off));      // x + 0.0 is exact
        v_hi = __dadd_rn(v_hi, __shfl_xor_sync(0xffffffffu, v_hi, off));
      }
      med[d] = (r_lo == r_hi) ? v_lo : __ddiv_rn(__dadd_rn(v_lo, v_hi), 2.0);
    }
    // centre: argmin_i || p_i - median ||, first minimum in member order (np.linalg.norm: sqrt(dx^2 + dy^2))
    double best = INFINITY;
    int best_a = 0x7fffffff;
    for (int a = lane; a < m; a += 32) {
      const int64_t pi = p0 + order[start + a];
      const double dx = __dsub_rn(pts[pi * 2], med[0]), dy = __dsub_rn(pts[pi * 2 + 1], med[1]);
      const double dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      if (dist < best) { best = dist; best_a = a; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, off);
      const int oa = __shfl_xor_sync(0xffffffffu, best_a, off);
      if (ov < best || (ov == best && oa < best_a)) { best = ov; best_a = oa; }
    }
    if (lane == 0) centers[c0 + k] = (m > 0) ? order[start + best_a] : -1;
  }
}

}  // namespace

}  // namespace cns

using namespace cns;

extern "C" {

int cns_ap_similarity(const double* points, const int64_t* pt_off, const int64_t* mat_off, const int32_t* row_graph,
                      int64_t n_rows_total, double* S_out, void* stream) {
  CNS_REQUIRE(points && pt_off && mat_off && row_graph && S_out, "NULL pointer");
  if (n_rows_total <= 0) return CNS_OK;
  gb_similarity_kernel<<<(unsigned)n_rows_total, 256, 0, (cudaStream_t)stream>>>(points, pt_off, mat_off, row_graph, n_rows_total, S_out);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

size_t cns_segment_median_workspace_bytes(int64_t n_segments) {
  return align_up((size_t)2 * n_segments * sizeof(SelState)) + align_up((size_t)2 * n_segments * 256 * 8) + 256;
}

int cns_segment_median(const double* values, const int64_t* seg_off, int64_t n_segments, int64_t max_segment,
                       double* median_out, void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(values && seg_off && median_out && workspace, "NULL pointer");
  CNS_REQUIRE(n_segments > 0 && n_segments < 65536, "1..65535 segments per call");
  if (workspace_bytes < cns_segment_median_workspace_bytes(n_segments)) {
    set_error("cns_segment_median: workspace too small");
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  SelState* st = (SelState*)workspace;
  unsigned long long* hist = (unsigned long long*)((char*)workspace + align_up((size_t)2 * n_segments * sizeof(SelState)));
  sel_init_kernel<<<ceil_div(2 * n_segments * 256, 256), 256, 0, s>>>(seg_off, n_segments, st, hist);
  CNS_LAUNCH_CHECK();
  int blocks = ceil_div(max_segment > 0 ? max_segment : 1, 256 * 16);
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  for (int pass = 0; pass < 8; ++pass) {
    sel_hist_kernel<<<dim3(blocks, (unsigned)n_segments), 256, 0, s>>>(values, seg_off, st, pass, hist);
    CNS_LAUNCH_CHECK();
    sel_pick_kernel<<<ceil_div(2 * n_segments, 128), 128, 0, s>>>(2 * n_segments, st, hist);
    CNS_LAUNCH_CHECK();
  }
  sel_finish_kernel<<<ceil_div(n_segments, 128), 128, 0, s>>>(seg_off, n_segments, st, median_out);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int cns_ap_prepare(double* S, const double* noise, const double* preference, const int64_t* pt_off, const int64_t* mat_off,
                   const int32_t* row_graph, int64_t n_rows_total, void* stream) {
  CNS_REQUIRE(S && noise && preference && pt_off && mat_off && row_graph, "NULL pointer");
  if (n_rows_total <= 0) return CNS_OK;
  gb_prepare_kernel<<<(unsigned)n_rows_total, 256, 0, (cudaStream_t)stream>>>(S, noise, preference, pt_off, mat_off, row_graph, n_rows_total);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int cns_ap_labels(const double* S, const unsigned char* E, const int64_t* pt_off, const int64_t* mat_off, int64_t n_graphs,
                  int64_t n_rows_total, int32_t* labels_out, int32_t* exemplars_out, int32_t* n_clusters_out,
                  void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(S && E && pt_off && mat_off && labels_out && exemplars_out && n_clusters_out && workspace, "NULL pointer");
  CNS_REQUIRE(n_graphs > 0, "empty batch");
  const size_t need = 2 * align_up((size_t)n_rows_total * sizeof(int));
  if (workspace_bytes < need) { set_error("cns_ap_labels: workspace %zu < %zu bytes", workspace_bytes, need); return CNS_ERR_WORKSPACE_TOO_SMALL; }
  int* ex = (int*)workspace;
  int* lab = (int*)((char*)workspace + align_up((size_t)n_rows_total * sizeof(int)));
  gb_labels_kernel<<<(unsigned)n_graphs, 256, 0, (cudaStream_t)stream>>>(S, E, pt_off, mat_off, ex, lab, labels_out, exemplars_out, n_clusters_out);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int cns_graph_layout(const double* points, const int32_t* labels, const int64_t* pt_off, const int64_t* clu_off,
                     int64_t n_graphs, int32_t* member_order, int32_t* cluster_size, int32_t* centers, void* stream) {
  CNS_REQUIRE(points && labels && pt_off && clu_off && member_order && cluster_size && centers, "NULL pointer");
  CNS_REQUIRE(n_graphs > 0, "empty batch");
  gb_layout_kernel<<<(unsigned)n_graphs, 256, 0, (cudaStream_t)stream>>>(points, labels, pt_off, clu_off, member_order, cluster_size, centers);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // extern "C"
