// Affinity Propagation message passing on the device, bit-exact with scikit-learn's
// `_affinity_propagation` loop (sklearn/cluster/_affinity_propagation.py, the clustering the reference
// calls in cns/midend/graph_gen.py:104-106 with damping 0.9).
//
// The similarity matrix S (with the diagonal preference and the degeneracy-removing noise already
// applied on the host, exactly as sklearn builds it) comes in; the exemplar indicator E, the iteration
// count and a converged flag come out.  Everything is IEEE fp64 with the operation order of the numpy
// code and no FMA contraction (__dadd_rn / __dmul_rn), column sums run sequentially over rows like
// numpy's axis-0 reduction - so responsibilities / availabilities are bit-identical and so are the
// exemplars.  Batched: G independent problems of different sizes per launch; a problem that has met
// sklearn's convergence test is frozen at that iteration.
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

struct ApArgs {
  int64_t n_graphs, n_rows_total;
  const int64_t* pt_off;     // (G+1) prefix of n_g
  const int64_t* mat_off;    // (G+1) prefix of n_g^2
  const int* row_graph;      // (sum n) graph of every row
  const double* S; double* A; double* R; double* colsum;   // colsum: (sum n)
  unsigned char* hist;       // (sum n, conv_iter) ring of E
  unsigned char* E;          // (sum n) exemplar indicator of the last evaluated iteration
  int* done; int* n_iter;    // (G)
  int* n_active;             // (1) problems still iterating
  double damping, one_minus_damping;
  int conv_iter, max_iter, it;
};

// ---- responsibilities: one warp per row -----------------------------------------------------------
__global__ void __launch_bounds__(256)
ap_row_kernel(ApArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.n_rows_total) return;
  const int g = p.row_graph[row];
  if (p.done[g]) return;
  const int64_t n = p.pt_off[g + 1] - p.pt_off[g];
  const int64_t i = row - p.pt_off[g];
  const double* S = p.S + p.mat_off[g] + i * n;
  const double* A = p.A + p.mat_off[g] + i * n;
  double* R = p.R + p.mat_off[g] + i * n;
  // pass 1: Y = max(A + S), I = first argmax
  double best = -INFINITY; int64_t arg = 0x7fffffffffffffffll;
  for (int64_t k = lane; k < n; k += 32) {
    const double v = __dadd_rn(A[k], S[k]);
    if (v > best) { best = v; arg = k; }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, off);
    const int64_t oa = __shfl_xor_sync(0xffffffffu, arg, off);
    if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
  }
  // pass 2: Y2 = max over k != I
  double second = -INFINITY;
  for (int64_t k = lane; k < n; k += 32) {
    if (k == arg) continue;
    const double v = __dadd_rn(A[k], S[k]);
    second = fmax(second, v);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) second = fmax(second, __shfl_xor_sync(0xffffffffu, second, off));
  // pass 3: R = R * damping + (S - Y [or Y2 at I]) * (1 - damping)
  for (int64_t k = lane; k < n; k += 32) {
    const double rn = __dsub_rn(S[k], (k == arg) ? second : best);
    R[k] = __dadd_rn(__dmul_rn(R[k], p.damping), __dmul_rn(rn, p.one_minus_damping));
  }
}

// ---- column sums of Rp = max(R, 0) with the diagonal kept: sequential over rows (numpy axis-0 order) --
__global__ void __launch_bounds__(128)
ap_colsum_kernel(ApArgs p) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= p.n_rows_total) return;
  const int g = p.row_graph[col];
  if (p.done[g]) return;
  const int64_t n = p.pt_off[g + 1] - p.pt_off[g];
  const int64_t j = col - p.pt_off[g];
  const double* R = p.R + p.mat_off[g] + j;
  double s = 0.0;
  int64_t i = 0;
  {
    const double r = R[0];
    s = (j == 0) ? r : fmax(r, 0.0);
    i = 1;
  }
  for (; i + 4 <= n; i += 4) {
    double r0 = R[i * n], r1 = R[(i + 1) * n], r2 = R[(i + 2) * n], r3 = R[(i + 3) * n];
    r0 = (i == j) ? r0 : fmax(r0, 0.0);
    r1 = (i + 1 == j) ? r1 : fmax(r1, 0.0);
    r2 = (i + 2 == j) ? r2 : fmax(r2, 0.0);
    r3 = (i + 3 == j) ? r3 : fmax(r3, 0.0);
    s = __dadd_rn(s, r0); s = __dadd_rn(s, r1); s = __dadd_rn(s, r2); s = __dadd_rn(s, r3);
  }
  for (; i < n; ++i) {
    double r = R[i * n];
    r = (i == j) ? r : fmax(r, 0.0);
    s = __dadd_rn(s, r);
  }
  p.colsum[col] = s;
}

// ---- availabilities: element-wise given the column sums ----------------------------------------------
__global__ void __launch_bounds__(256)
ap_avail_kernel(ApArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.n_rows_total) return;
  const int g = p.row_graph[row];
  if (p.done[g]) return;
  const int64_t n = p.pt_off[g + 1] - p.pt_off[g];
  const int64_t i = row - p.pt_off[g];
  const double* R = p.R + p.mat_off[g] + i * n;
  double* A = p.A + p.mat_off[g] + i * n;
  const double* cs = p.colsum + p.pt_off[g];
  for (int64_t k = lane; k < n; k += 32) {
    const double r = R[k];
    const double rp = (k == i) ? r : fmax(r, 0.0);
    double t = __dsub_rn(rp, cs[k]);
    if (k != i) t = fmax(t, 0.0);
    A[k] = __dsub_rn(__dmul_rn(A[k], p.damping), __dmul_rn(t, p.one_minus_damping));
  }
}

// ---- convergence test of sklearn: exemplar set unchanged for conv_iter iterations and K > 0 -----------
__global__ void __launch_bounds__(256)
ap_converge_kernel(ApArgs p) {
  __shared__ int s_k, s_stable;
  const int g = blockIdx.x;
  if (p.done[g]) return;
  const int64_t n = p.pt_off[g + 1] - p.pt_off[g];
  const double* A = p.A + p.mat_off[g];
  const double* R = p.R + p.mat_off[g];
  unsigned char* hist = p.hist + p.pt_off[g] * p.conv_iter;
  unsigned char* E = p.E + p.pt_off[g];
  if (threadIdx.x == 0) { s_k = 0; s_stable = 0; }
  __syncthreads();
  int k_loc = 0, st_loc = 0;
  const int slot = p.it % p.conv_iter;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned char e = __dadd_rn(A[i * n + i], R[i * n + i]) > 0.0 ? 1 : 0;
    hist[i * p.conv_iter + slot] = e;
    E[i] = e;
    k_loc += e;
    int se = 0;
    for (int c = 0; c < p.conv_iter; ++c) se += hist[i * p.conv_iter + c];
    st_loc += (se == p.conv_iter) + (se == 0);
  }
  atomicAdd(&s_k, k_loc);
  atomicAdd(&s_stable, st_loc);
  __syncthreads();
  if (threadIdx.x == 0) {
    bool stop = false;
    if (p.it >= p.conv_iter && s_stable == (int)n && s_k > 0) stop = true;     // converged
    if (stop || p.it == p.max_iter - 1) {
      p.done[g] = stop ? 1 : 2;                                               // 2: ran out of iterations
      p.n_iter[g] = p.it + 1;
      atomicSub(p.n_active, 1);
    }
  }
}

}  // namespace cns

using namespace cns;

extern "C" {

size_t cns_ap_workspace_bytes(int64_t n_rows_total, int64_t n_mat_total, int64_t n_graphs, int conv_iter) {
  size_t b = 0;
  b += align_up((size_t)n_mat_total * 8) * 2;            // A, R
  b += align_up((size_t)n_rows_total * 8);               // colsum
  b += align_up((size_t)n_rows_total * conv_iter);       // hist
  b += align_up((size_t)n_graphs * 4) * 1 + 256;         // done, n_active
  return b + 256;
}

// Runs sklearn's Affinity Propagation iterations for a batch of problems.
//   S        (sum n_g^2) fp64 device, prepared like sklearn (preference on the diagonal, noise added)
//   pt_off   (G+1) int64 device, mat_off (G+1) int64 device, row_graph (sum n) int32 device
//   E_out    (sum n) u8 device: exemplar indicator;  n_iter_out (G) int32 device;
//   converged_host (G) int32 HOST: 1 converged, 2 hit max_iter
// Polls the device every `poll` iterations (one 4-byte copy) to stop once every problem is done.
int cns_affinity_propagation(const double* S, const int64_t* pt_off, const int64_t* mat_off, const int32_t* row_graph,
                             int64_t n_graphs, int64_t n_rows_total, int64_t n_mat_total, double damping,
                             int max_iter, int conv_iter, unsigned char* E_out, int32_t* n_iter_out,
                             int32_t* converged_host, void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(S && pt_off && mat_off && row_graph && E_out && n_iter_out && workspace, "NULL pointer");
  CNS_REQUIRE(n_graphs > 0 && conv_iter > 0 && conv_iter <= 64 && max_iter > 0, "bad argument");
  if (workspace_bytes < cns_ap_workspace_bytes(n_rows_total, n_mat_total, n_graphs, conv_iter)) {
    set_error("cns_affinity_propagation: workspace too small");
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* w = (char*)workspace;
  ApArgs a{};
  a.n_graphs = n_graphs; a.n_rows_total = n_rows_total; a.pt_off = pt_off; a.mat_off = mat_off; a.row_graph = row_graph;
  a.S = S;
  a.A = (double*)w; w += align_up((size_t)n_mat_total * 8);
  a.R = (double*)w; w += align_up((size_t)n_mat_total * 8);
  a.colsum = (double*)w; w += align_up((size_t)n_rows_total * 8);
  a.hist = (unsigned char*)w; w += align_up((size_t)n_rows_total * conv_iter);
  a.done = (int*)w; w += align_up((size_t)n_graphs * 4);
  a.n_active = (int*)w;
  a.E = E_out; a.n_iter = n_iter_out;
  a.damping = damping; a.one_minus_damping = 1.0 - damping;      // same fp64 expression as `1 - damping` in numpy
  a.conv_iter = conv_iter; a.max_iter = max_iter;
  CNS_CUDA_CHECK(cudaMemsetAsync(a.A, 0, (size_t)n_mat_total * 8, s));
  CNS_CUDA_CHECK(cudaMemsetAsync(a.R, 0, (size_t)n_mat_total * 8, s));
  CNS_CUDA_CHECK(cudaMemsetAsync(a.hist, 0, (size_t)n_rows_total * conv_iter, s));
  CNS_CUDA_CHECK(cudaMemsetAsync(a.done, 0, (size_t)n_graphs * 4, s));
  int active = (int)n_graphs;
  CNS_CUDA_CHECK(cudaMemcpyAsync(a.n_active, &active, sizeof(int), cudaMemcpyHostToDevice, s));
  const int row_blocks = ceil_div(n_rows_total, 8);
  const int poll = 8;
  for (int it = 0; it < max_iter; ++it) {
    a.it = it;
    ap_row_kernel<<<row_blocks, 256, 0, s>>>(a);
    CNS_LAUNCH_CHECK();
    ap_colsum_kernel<<<ceil_div(n_rows_total, 128), 128, 0, s>>>(a);
    CNS_LAUNCH_CHECK();
    ap_avail_kernel<<<row_blocks, 256, 0, s>>>(a);
    CNS_LAUNCH_CHECK();
    ap_converge_kernel<<<(unsigned)n_graphs, 256, 0, s>>>(a);
    CNS_LAUNCH_CHECK();
    if (it >= conv_iter && (it % poll == poll - 1 || it == max_iter - 1)) {
      CNS_CUDA_CHECK(cudaMemcpyAsync(&active, a.n_active, sizeof(int), cudaMemcpyDeviceToHost, s));
      CNS_CUDA_CHECK(cudaStreamSynchronize(s));
      if (active <= 0) break;
    }
  }
  if (converged_host) {
    CNS_CUDA_CHECK(cudaMemcpyAsync(converged_host, a.done, (size_t)n_graphs * 4, cudaMemcpyDeviceToHost, s));
    CNS_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return CNS_OK;
}

}  // extern "C"
