// Index preparation: segment pointers and destination-sorted CSR from the reference's int64 edge
// lists (PyG convention edge_index[0] = source j, edge_index[1] = destination i,
// cns/models/graph_vs.py:41-43).  Integer work only; int32 on device.
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

// ---------------------------------------------------------------------------------------------
// exclusive scan of int32 (n up to 2^31), three phases, 1024 elements per block
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 1024;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = warp_sums[lane];
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    warp_sums[lane] = winc - w;  // exclusive
    if (lane == 31 && total) *total = winc;
  }
  __syncthreads();
  int res = inc - v + warp_sums[wid];
  __syncthreads();
  return res;
}

__global__ void scan_block_sums_kernel(const int* __restrict__ in, int64_t n, int* __restrict__ sums) {
  __shared__ int total;
  int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  int v = (i < n) ? in[i] : 0;
  block_exclusive_scan(v, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of up to 1024*1024 block sums, in place
__global__ void scan_sums_kernel(int* sums, int n_blocks) {
  __shared__ int total;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += SCAN_BLOCK) {
    int i = base + threadIdx.x;
    int v = (i < n_blocks) ? sums[i] : 0;
    int ex = block_exclusive_scan(v, &total);
    if (i < n_blocks) sums[i] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
}

// out[i] = exclusive prefix; out has n+1 entries when `write_total`
__global__ void scan_apply_kernel(const int* __restrict__ in, int64_t n, const int* __restrict__ sums,
                                  int* __restrict__ out) {
  __shared__ int total;
  int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  int v = (i < n) ? in[i] : 0;
  int ex = block_exclusive_scan(v, &total) + sums[blockIdx.x];
  if (i < n) out[i] = ex;
  if (i == n - 1) out[n] = ex + v;
}

// in-place safe (in may alias out).  scratch: ceil(n/1024) ints.
int exclusive_scan_i32(const int* in, int64_t n, int* out, int* scratch, cudaStream_t s) {
  if (n == 0) {
    CNS_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(int), s));
    return CNS_OK;
  }
  int nb = ceil_div(n, SCAN_BLOCK);
  CNS_REQUIRE(nb <= SCAN_BLOCK * SCAN_BLOCK, "scan too large");
  scan_block_sums_kernel<<<nb, SCAN_BLOCK, 0, s>>>(in, n, scratch);
  CNS_LAUNCH_CHECK();
  scan_sums_kernel<<<1, SCAN_BLOCK, 0, s>>>(scratch, nb);
  CNS_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, SCAN_BLOCK, 0, s>>>(in, n, scratch, out);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

size_t scan_scratch_bytes(int64_t n) { return align_up((size_t)(ceil_div(n, SCAN_BLOCK) + 1) * sizeof(int)); }

// ---------------------------------------------------------------------------------------------
// rowptr of a sorted destination list (l0_to_l1_edge_index_i): boundary detection, no atomics
// ---------------------------------------------------------------------------------------------
__global__ void sorted_rowptr_kernel(const int64_t* __restrict__ dst, int64_t n_edges, int64_t n_rows,
                                     int* __restrict__ rowptr, int* __restrict__ err) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > n_edges) return;
  int64_t prev = (e == 0) ? -1 : dst[e - 1];
  int64_t cur = (e == n_edges) ? n_rows : dst[e];
  if (cur < prev || cur < 0 || cur > n_rows || (e < n_edges && cur >= n_rows)) {
    atomicExch(err, 1);
    return;
  }
  for (int64_t r = prev + 1; r <= cur; ++r) rowptr[r] = (int)e;
}

int sorted_rowptr(const int64_t* dst, int64_t n_edges, int64_t n_rows, int* rowptr, int* err_flag,
                  cudaStream_t s) {
  sorted_rowptr_kernel<<<ceil_div(n_edges + 1, 256), 256, 0, s>>>(dst, n_edges, n_rows, rowptr, err_flag);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// ---------------------------------------------------------------------------------------------
// generic CSR by destination: count (int atomics) -> scan -> slot fill
// ---------------------------------------------------------------------------------------------
__global__ void csr_count_kernel(const int64_t* __restrict__ dst, int64_t n_edges, int64_t n_rows,
                                 int* __restrict__ deg, int* __restrict__ err) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int64_t i = dst[e];
  if (i < 0 || i >= n_rows) { atomicExch(err, 2); return; }
  atomicAdd(&deg[i], 1);
}

// Ascending in-place sort of one CSR row.  The fill kernels place edges with an atomic cursor, so the order inside a
// row depends on scheduling; sorting makes the CSR - and with it every sum over a row in the backward pass and the
// tie-breaking of the arg-max - identical from run to run.  Rows are short (cluster graphs: <= a few dozen entries).
__device__ __forceinline__ void sort_row_i32(int* __restrict__ a, int n) {
  if (n <= 48) {
    for (int i = 1; i < n; ++i) {
      const int v = a[i];
      int j = i - 1;
      while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
      a[j + 1] = v;
    }
    return;
  }
  auto sift = [&](int root, int end) {        // heap sort for the rare long row
    for (;;) {
      int child = 2 * root + 1;
      if (child >= end) break;
      if (child + 1 < end && a[child] < a[child + 1]) ++child;
      if (a[root] >= a[child]) break;
      const int t = a[root]; a[root] = a[child]; a[child] = t;
      root = child;
    }
  };
  for (int i = n / 2 - 1; i >= 0; --i) sift(i, n);
  for (int end = n - 1; end > 0; --end) {
    const int t = a[0]; a[0] = a[end]; a[end] = t;
    sift(0, end);
  }
}

// One warp per row: rows of up to 128 entries (cluster graphs: a few dozen) are ranked out of shared memory - entry k goes
// to position #(smaller entries) + #(equal entries in front of it) - which replaces a serial insertion sort in global
// memory by one thread (44 us for 4.8 k rows of ~19 entries).  Longer rows fall back to that serial sort.
__global__ void __launch_bounds__(256)
csr_sort_rows_kernel(const int* __restrict__ rowptr, int64_t n_rows, int* __restrict__ out_src) {
  __shared__ int buf[8][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + w;
  if (i >= n_rows) return;
  const int beg = rowptr[i], n = rowptr[i + 1] - beg;
  int* a = out_src + beg;
  if (n > 128) {
    if (lane == 0) sort_row_i32(a, n);
    return;
  }
  for (int k = lane; k < n; k += 32) buf[w][k] = a[k];
  __syncwarp();
  for (int k = lane; k < n; k += 32) {
    const int v = buf[w][k];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const int u = buf[w][j];
      rank += (u < v || (u == v && j < k)) ? 1 : 0;
    }
    a[rank] = v;
  }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                int64_t n_edges, int64_t n_rows, const int* __restrict__ rowptr,
                                int* __restrict__ cursor, int* __restrict__ out_src) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int64_t i = dst[e];
  if (i < 0 || i >= n_rows) return;
  int slot = rowptr[i] + atomicAdd(&cursor[i], 1);
  out_src[slot] = (int)src[e];
}

// workspace: cursor (n_rows+1 ints) + scan scratch
size_t csr_workspace_bytes(int64_t n_rows) {
  return align_up((size_t)(n_rows + 1) * sizeof(int)) + scan_scratch_bytes(n_rows + 1);
}

int csr_by_dst(const int64_t* ej, const int64_t* ei, int64_t n_edges, int64_t n_rows, int* rowptr, int* src,
               void* workspace, int* err_flag, cudaStream_t s) {
  int* cursor = (int*)workspace;
  int* scratch = (int*)((char*)workspace + align_up((size_t)(n_rows + 1) * sizeof(int)));
  CNS_CUDA_CHECK(cudaMemsetAsync(cursor, 0, (size_t)(n_rows + 1) * sizeof(int), s));
  if (n_edges > 0) {
    csr_count_kernel<<<ceil_div(n_edges, 256), 256, 0, s>>>(ei, n_edges, n_rows, cursor, err_flag);
    CNS_LAUNCH_CHECK();
  }
  // scan over n_rows degrees, writes rowptr[0..n_rows]
  int rc = exclusive_scan_i32(cursor, n_rows, rowptr, scratch, s);
  if (rc != CNS_OK) return rc;
  if (n_rows == 0) return CNS_OK;
  CNS_CUDA_CHECK(cudaMemsetAsync(cursor, 0, (size_t)(n_rows + 1) * sizeof(int), s));
  if (n_edges > 0) {
    csr_fill_kernel<<<ceil_div(n_edges, 256), 256, 0, s>>>(ej, ei, n_edges, n_rows, rowptr, cursor, src);
    CNS_LAUNCH_CHECK();
    csr_sort_rows_kernel<<<ceil_div(n_rows, 8), 256, 0, s>>>(rowptr, n_rows, src);
    CNS_LAUNCH_CHECK();
  }
  return CNS_OK;
}

// ---------------------------------------------------------------------------------------------
// per-cluster metadata: centre position, graph id; graph_ptr from num_clusters
// ---------------------------------------------------------------------------------------------
__global__ void i64_to_i32_kernel(const int64_t* __restrict__ in, int64_t n, int* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)in[i];
}

int graph_ptr_from_counts(const int64_t* num_clusters, int64_t n_graphs, int* graph_ptr, int* tmp,
                          int* scratch, cudaStream_t s) {
  if (n_graphs > 0) {
    i64_to_i32_kernel<<<ceil_div(n_graphs, 256), 256, 0, s>>>(num_clusters, n_graphs, tmp);
    CNS_LAUNCH_CHECK();
  }
  return exclusive_scan_i32(tmp, n_graphs, graph_ptr, scratch, s);
}

// ---------------------------------------------------------------------------------------------
// Small-batch fast path (servo loop, batch 1): every index structure of one forward in ONE
// single-CTA launch instead of 13 - graph_ptr, cluster rowptr and both destination CSRs.
// Requires n_clusters <= PREP_SMALL_MAX (degrees / cursors live in shared memory).
// ---------------------------------------------------------------------------------------------
__device__ void block_scan_inplace(int* v, int n, int* out /* n+1 */) {
  // exclusive scan of v[0..n) into out[0..n], n <= PREP_SMALL_MAX, blockDim = 1024
  __shared__ int carry_s;
  __shared__ int tot_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int x = (i < n) ? v[i] : 0;
    const int ex = block_exclusive_scan(x, &tot_s);
    if (i < n) out[i] = ex + carry_s;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += tot_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = carry_s;
  __syncthreads();
}

// graph_ptr (exclusive scan of num_clusters) and the whole-graph row tiles of the fused backbone in one
// single-CTA launch: tile_graph[t] = first graph of tile t, greedily as many consecutive graphs as fit
// into 128 rows.  n_graphs <= GRAPH_TILES_MAX.
__global__ void __launch_bounds__(1024)
graph_tiles_kernel(const int64_t* __restrict__ num_clusters, int n_graphs, int* __restrict__ graph_ptr,
                   int* __restrict__ tile_graph, int* __restrict__ n_tiles, int* __restrict__ err) {
  __shared__ int cnt[GRAPH_TILES_MAX];
  __shared__ int gp[GRAPH_TILES_MAX + 1];
  const int tid = threadIdx.x;
  for (int i = tid; i < n_graphs; i += 1024) cnt[i] = (int)num_clusters[i];
  __syncthreads();
  block_scan_inplace(cnt, n_graphs, gp);
  for (int i = tid; i <= n_graphs; i += 1024) graph_ptr[i] = gp[i];
  // cnt[g] := first graph that does not fit into a tile starting at g (binary search on the prefix sums)
  for (int g = tid; g < n_graphs; g += 1024) {
    const int lim = gp[g] + 128;
    if (gp[g + 1] > lim) atomicExch(err, 4);            // a graph with more than 128 clusters
    int lo = g + 1, hi = n_graphs;                      // largest q in [g+1, B] with gp[q] <= lim
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (gp[mid] <= lim) lo = mid; else hi = mid - 1;
    }
    cnt[g] = lo;
  }
  __syncthreads();
  if (tid == 0) {
    int t = 0, g = 0;
    while (g < n_graphs) { tile_graph[t++] = g; const int q = cnt[g]; g = q > g ? q : g + 1; }
    tile_graph[t] = n_graphs;
    *n_tiles = t;
  }
}

// ---------------------------------------------------------------------------------------------
// Per-frame masking of B resident target graphs in cluster-major form (GraphGenerator._get_graph,
// graph_gen.py:199-228, for every scene at once): the keypoint -> cluster edge list keeps the visible
// members of every cluster in their static order, so it is a per-cluster compaction:
// count (warp per cluster) -> scan (1024 clusters per CTA) -> emit (warp per cluster).  The scan IS the cluster rowptr the
// encoder needs, so the forward does not have to re-derive it from the edge list.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
frame_count_kernel(const int64_t* __restrict__ member_order, const int* __restrict__ member_ptr,
                   const uint8_t* __restrict__ visible, int n_clusters, int* __restrict__ cnt, uint8_t* __restrict__ cluster_mask) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= n_clusters) return;
  const int beg = member_ptr[c], end = member_ptr[c + 1];
  int n = 0;
  for (int k = beg + lane; k < end; k += 32) n += visible[member_order[k]] ? 1 : 0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
  if (lane == 0) { cnt[c] = n; cluster_mask[c] = n > 0 ? 1 : 0; }
}

// exclusive scan of the counts, 1024 per CTA: a CTA first sums everything in front of its chunk (coalesced, L2-resident,
// at most n / 1024 loads per thread), then scans its own 1024 -- no serial single-CTA pass over the whole array
__global__ void __launch_bounds__(1024)
frame_scan_kernel(const int* __restrict__ cnt, int n, int* __restrict__ rowptr) {
  __shared__ int front, total;
  const int base = blockIdx.x * SCAN_BLOCK, i = base + threadIdx.x;
  int pre = 0;
  for (int k = threadIdx.x; k < base; k += SCAN_BLOCK) pre += cnt[k];
  block_exclusive_scan(pre, &front);
  __syncthreads();
  const int v = i < n ? cnt[i] : 0;
  const int ex = block_exclusive_scan(v, &total) + front;
  if (i < n) rowptr[i] = ex;
  if (i == n - 1) rowptr[n] = ex + v;
}

__global__ void __launch_bounds__(256)
frame_emit_kernel(const int64_t* __restrict__ member_order, const int* __restrict__ member_ptr,
                  const uint8_t* __restrict__ visible, int n_clusters, const int* __restrict__ rowptr,
                  int64_t* __restrict__ e01_j, int64_t* __restrict__ e01_i) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= n_clusters) return;
  const int beg = member_ptr[c], end = member_ptr[c + 1];
  int out = rowptr[c];
  for (int k0 = beg; k0 < end; k0 += 32) {
    const int k = k0 + lane;
    const int64_t node = k < end ? member_order[k] : 0;
    const bool v = k < end && visible[node];
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (v) {
      const int pos = out + __popc(bal & ((1u << lane) - 1u));
      e01_j[pos] = node;
      e01_i[pos] = c;
    }
    out += __popc(bal);
  }
}

int frame_mask(const int64_t* member_order, const int* member_ptr, const uint8_t* visible, int64_t n_clusters,
               int64_t* e01_j, int64_t* e01_i, uint8_t* cluster_mask, int* clu_rowptr, int* cnt_scratch, cudaStream_t s) {
  if (n_clusters == 0) return CNS_OK;
  const int C = (int)n_clusters;
  frame_count_kernel<<<ceil_div(C, 8), 256, 0, s>>>(member_order, member_ptr, visible, C, cnt_scratch, cluster_mask);
  CNS_LAUNCH_CHECK();
  frame_scan_kernel<<<ceil_div(C, SCAN_BLOCK), SCAN_BLOCK, 0, s>>>(cnt_scratch, C, clu_rowptr);
  CNS_LAUNCH_CHECK();
  frame_emit_kernel<<<ceil_div(C, 8), 256, 0, s>>>(member_order, member_ptr, visible, C, clu_rowptr, e01_j, e01_i);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int graph_tiles(const int64_t* num_clusters, int64_t n_graphs, int* graph_ptr, int* tile_graph, int* n_tiles, int* err,
                cudaStream_t s) {
  CNS_REQUIRE(n_graphs <= GRAPH_TILES_MAX, "graph_tiles: too many graphs");
  graph_tiles_kernel<<<1, 1024, 0, s>>>(num_clusters, (int)n_graphs, graph_ptr, tile_graph, n_tiles, err);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

__global__ void __launch_bounds__(1024)
prep_small_kernel(PrepSmallArgs p) {
  __shared__ int deg[PREP_SMALL_MAX + 1];
  __shared__ int cur[PREP_SMALL_MAX + 1];
  const int tid = threadIdx.x;
  const int C = (int)p.n_clusters, B = (int)p.n_graphs;
  // graph_ptr = exclusive scan of num_clusters
  for (int i = tid; i < B; i += 1024) deg[i] = (int)p.num_clusters[i];
  __syncthreads();
  block_scan_inplace(deg, B, p.graph_ptr);
  // cluster rowptr from the sorted destination list
  for (int64_t e = tid; e <= p.n_e01; e += 1024) {
    const int64_t prev = (e == 0) ? -1 : p.ei[e - 1];
    const int64_t cu = (e == p.n_e01) ? C : p.ei[e];
    if (cu < prev || cu < 0 || cu > C || (e < p.n_e01 && cu >= C)) { atomicExch(p.err, 1); continue; }
    for (int64_t r = prev + 1; r <= cu; ++r) p.clu_rowptr[r] = (int)e;
  }
  // the two destination CSRs
  for (int which = 0; which < 2; ++which) {
    const int64_t* ej = which ? p.tar_j : p.cur_j;
    const int64_t* ei = which ? p.tar_i : p.cur_i;
    const int64_t ne = which ? p.n_e11_tar : p.n_e11_cur;
    int* rowptr = which ? p.tar_rowptr : p.cur_rowptr;
    int* src = which ? p.tar_src : p.cur_src;
    __syncthreads();
    for (int i = tid; i <= C; i += 1024) { deg[i] = 0; cur[i] = 0; }
    __syncthreads();
    for (int64_t e = tid; e < ne; e += 1024) {
      const int64_t i = ei[e];
      if (i < 0 || i >= C) { atomicExch(p.err, 2); continue; }
      atomicAdd(&deg[i], 1);
    }
    __syncthreads();
    block_scan_inplace(deg, C, rowptr);
    for (int64_t e = tid; e < ne; e += 1024) {
      const int64_t i = ei[e];
      if (i < 0 || i >= C) continue;
      const int slot = rowptr[i] + atomicAdd(&cur[i], 1);
      src[slot] = (int)ej[e];
    }
    __syncthreads();
    for (int i = tid; i < C; i += 1024) sort_row_i32(src + rowptr[i], rowptr[i + 1] - rowptr[i]);   // canonical order
  }
}

int prep_small(const PrepSmallArgs& a, cudaStream_t s) {
  prep_small_kernel<<<1, 1024, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns

using namespace cns;

extern "C" int cns_csr_by_dst(const int64_t* edge_index, int64_t n_edges, int64_t n_dst, int32_t* rowptr,
                              int32_t* src, void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(n_edges >= 0 && n_dst >= 0, "negative size");
  CNS_REQUIRE(rowptr && (src || n_edges == 0) && workspace, "NULL pointer");
  size_t need = csr_workspace_bytes(n_dst) + 256;
  if (workspace_bytes < need) {
    set_error("cns_csr_by_dst: workspace %zu < %zu", workspace_bytes, need);
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int* err = (int*)workspace;
  CNS_CUDA_CHECK(cudaMemsetAsync(err, 0, sizeof(int), s));
  return csr_by_dst(edge_index, edge_index + n_edges, n_edges, n_dst, rowptr, src, (char*)workspace + 256, err, s);
}
