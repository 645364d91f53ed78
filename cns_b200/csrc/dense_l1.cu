// Cluster-level graphs that are complete digraphs (CNS_BATCH_L1_*_COMPLETE).
//
// GraphGenerator builds l1_dense_edge_index as ALL ordered pairs, self loops included, over a node set
// (graph_gen.py:31-36): every cluster of the graph for the target-side list, the clusters that still have
// a visible keypoint for the current-side list (graph_gen.py:199-228).  For such a graph the max
// aggregation of PointEdgeConv (graph_vs.py:73,94-111) over the neighbours of cluster i,
//     conv_i = P_i + max_{j in N(i)} Q_j ,          N(i) = node set S of i's graph,
// is the same column max for every i in S: one pass over the graph's rows instead of |S| gathers per
// row, and no CSR at all.  Rows outside S have no incoming edge and aggregate to 0 like PyG does.
//
//   l1_alive / l1_verify : prove the claim on the device (positional check of both edge lists against
//                          the complete digraph in source-major order); violations set the status word
//   edgeconv_dense       : one CTA per graph, thread = channel; optional fused GraphNorm + ReLU
//                          (PERConv, graph_vs.py:141-143) or the GRU gate / candidate epilogues
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

// ---- alive list per graph -------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
l1_alive_kernel(const int* __restrict__ graph_ptr, const uint8_t* __restrict__ mask, int* __restrict__ alive_idx,
                int* __restrict__ alive_cnt) {
  __shared__ int wsum[4];
  __shared__ int base_s;
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int beg = graph_ptr[g], end = graph_ptr[g + 1];
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int r0 = beg; r0 < end; r0 += 128) {
    const int r = r0 + tid;
    const bool a = r < end && mask[r] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, a);
    if (lane == 0) wsum[wid] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < wid; ++w) off += wsum[w];
    if (a) alive_idx[beg + off + __popc(bal & ((1u << lane) - 1u))] = r;
    __syncthreads();
    if (tid == 0) base_s += wsum[0] + wsum[1] + wsum[2] + wsum[3];
    __syncthreads();
  }
  if (tid == 0) alive_cnt[g] = base_s;
}

// ---- positional verification: edges of graph g are (S[k / n], S[k % n]), k = 0 .. n^2-1 ---------------
struct L1VerifySet {
  const int64_t* ej; const int64_t* ei; int64_t n_edges;
  const int* alive_idx; const int* alive_cnt;     // NULL / NULL: every cluster is a node (n = rows of the graph)
};

__global__ void __launch_bounds__(256)
l1_verify_kernel(L1VerifySet s0, L1VerifySet s1, const int* __restrict__ graph_ptr, int n_graphs, int* __restrict__ err) {
  __shared__ long long red[8];
  const L1VerifySet& s = blockIdx.y ? s1 : s0;
  if (s.ej == nullptr) return;
  const int g = blockIdx.x, tid = threadIdx.x;
  auto count = [&](int q) { return s.alive_cnt ? s.alive_cnt[q] : graph_ptr[q + 1] - graph_ptr[q]; };
  long long part = 0;
  for (int q = tid; q < g; q += 256) { const long long n = count(q); part += n * n; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  long long off0 = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) off0 += red[w];
  const int beg = graph_ptr[g];
  const long long n = count(g), nn = n * n;
  bool bad = off0 + nn > s.n_edges;
  if (!bad) {
    for (long long k = tid; k < nn; k += 256) {
      const int a = (int)(k / n), b = (int)(k - (long long)a * n);
      const int sj = s.alive_idx ? s.alive_idx[beg + a] : beg + a;
      const int si = s.alive_idx ? s.alive_idx[beg + b] : beg + b;
      if (s.ej[off0 + k] != sj || s.ei[off0 + k] != si) bad = true;
    }
  }
  if (g == n_graphs - 1 && off0 + nn != s.n_edges) bad = true;
  if (bad) atomicExch(err, 3);
}

int l1_complete_verify(const L1CompleteArgs& a, cudaStream_t s) {
  if (a.n_graphs == 0) return CNS_OK;
  if (a.cur_mask && a.cur_j) {
    l1_alive_kernel<<<(unsigned)a.n_graphs, 128, 0, s>>>(a.graph_ptr, a.cur_mask, a.alive_idx, a.alive_cnt);
    CNS_LAUNCH_CHECK();
  }
  if (!a.cur_j && !a.tar_j) return CNS_OK;
  L1VerifySet s0{a.cur_j, a.cur_i, a.n_e11_cur, a.alive_idx, a.alive_cnt};
  L1VerifySet s1{a.tar_j, a.tar_i, a.n_e11_tar, nullptr, nullptr};
  if (!a.cur_mask) s0.ej = nullptr;
  l1_verify_kernel<<<dim3((unsigned)a.n_graphs, 2), 256, 0, s>>>(s0, s1, a.graph_ptr, (int)a.n_graphs, a.err);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// ---- max aggregation over a complete graph ---------------------------------------------------------------
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }

constexpr int ECD_REG_ROWS = 32;   // graphs up to this many cluster rows are held in registers (one pass over [P|Q])

template <int F, int MODE, bool GN>
__global__ void __launch_bounds__(F)
edgeconv_dense_kernel(EdgeConvDenseArgs p) {
  const int g = blockIdx.x, c = threadIdx.x;
  const int beg = p.graph_ptr[g], end = p.graph_ptr[g + 1];
  const int n = end - beg;
  if (n <= 0) return;
  const uint8_t* __restrict__ mask = p.mask;
  const float* __restrict__ pq = p.pq;
  auto alive = [&](int r) { return mask == nullptr || mask[r] != 0; };

  // the candidate (four operand streams per row) was measured faster streaming at full occupancy (22 vs 27 us)
  if (MODE != EC_CANDIDATE && n <= ECD_REG_ROWS) {
    // the usual case (C ~ 0.8 sqrt(N) clusters per graph): every [P|Q] value of the graph is loaded exactly once,
    // all loads in flight together, and max / GraphNorm / gates run on registers
    float pv[ECD_REG_ROWS], qv[ECD_REG_ROWS];
    unsigned am = 0;
#pragma unroll
    for (int i = 0; i < ECD_REG_ROWS; ++i) {
      const bool in = i < n;
      if (in && alive(beg + i)) am |= 1u << i;
      pv[i] = in ? pq[(size_t)(beg + i) * 2 * F + c] : 0.f;
      qv[i] = in ? pq[(size_t)(beg + i) * 2 * F + F + c] : 0.f;
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < ECD_REG_ROWS; ++i)
      if (am & (1u << i)) m = fmaxf(m, qv[i]);
#pragma unroll
    for (int i = 0; i < ECD_REG_ROWS; ++i) pv[i] = (am & (1u << i)) ? pv[i] + m : 0.f;      // conv
    if (MODE == EC_PLAIN) {
      if (!GN) {
#pragma unroll
        for (int i = 0; i < ECD_REG_ROWS; ++i)
          if (i < n) p.out[(size_t)(beg + i) * F + c] = pv[i];
        return;
      }
      const float inv_n = 1.f / (float)n;
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < ECD_REG_ROWS; ++i)
        if (i < n) sum += pv[i];
      const float shift = sum * inv_n * p.gn_mean_scale[c];
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < ECD_REG_ROWS; ++i)
        if (i < n) { const float o = pv[i] - shift; sq = fmaf(o, o, sq); }
      const float rstd = 1.f / sqrtf(sq * inv_n + 1e-5f);
      const float w = p.gn_weight[c], b = p.gn_bias[c];
#pragma unroll
      for (int i = 0; i < ECD_REG_ROWS; ++i)
        if (i < n) p.gn_out[(size_t)(beg + i) * F + c] = fmaxf(fmaf(w * (pv[i] - shift), rstd, b), 0.f);
    } else if (MODE == EC_GATES) {
#pragma unroll
      for (int i = 0; i < ECD_REG_ROWS; ++i) {
        if (i >= n) break;
        const int row = beg + i;
        const float gt = sigmoid_(pv[i]);
        if (c < 128) p.hr[(size_t)row * 128 + c] = (p.h ? p.h[(size_t)row * 128 + c] : 0.f) * gt;
        else p.u[(size_t)row * 128 + c - 128] = gt;
      }
    }
    return;
  }

  // larger graphs: stream the rows (L2-resident after the first pass)
  // column max of Q over the node set
  float m = -INFINITY;
  int r = beg;
  for (; r + 4 <= end; r += 4) {
    float q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = alive(r + i) ? pq[(size_t)(r + i) * 2 * F + F + c] : -INFINITY;
    m = fmaxf(fmaxf(fmaxf(m, q[0]), fmaxf(q[1], q[2])), q[3]);
  }
  for (; r < end; ++r)
    if (alive(r)) m = fmaxf(m, pq[(size_t)r * 2 * F + F + c]);
  auto conv = [&](int row) { return alive(row) ? pq[(size_t)row * 2 * F + c] + m : 0.f; };

  if (MODE == EC_PLAIN) {
    if (!GN) {
      for (int row = beg; row < end; ++row) p.out[(size_t)row * F + c] = conv(row);
      return;
    }
    // GraphNorm over every cluster row of the graph (masked rows included, like the reference) + ReLU
    const float inv_n = 1.f / (float)n;
    float sum = 0.f;
    for (int row = beg; row < end; ++row) sum += conv(row);
    const float shift = sum * inv_n * p.gn_mean_scale[c];
    float sq = 0.f;
    for (int row = beg; row < end; ++row) {
      const float o = conv(row) - shift;
      sq = fmaf(o, o, sq);
    }
    const float rstd = 1.f / sqrtf(sq * inv_n + 1e-5f);
    const float w = p.gn_weight[c], b = p.gn_bias[c];
    for (int row = beg; row < end; ++row) {
      const float o = conv(row) - shift;
      p.gn_out[(size_t)row * F + c] = fmaxf(fmaf(w * o, rstd, b), 0.f);
    }
  } else if (MODE == EC_GATES) {
    // gates = sigmoid(conv) ; r = first 128 channels, u = last 128 (torch.chunk, graph_vs.py:180-181)
    for (int row = beg; row < end; ++row) {
      const float gt = sigmoid_(conv(row));
      if (c < 128) {
        const float hv = p.h ? p.h[(size_t)row * 128 + c] : 0.f;
        p.hr[(size_t)row * 128 + c] = hv * gt;
      } else {
        p.u[(size_t)row * 128 + c - 128] = gt;
      }
    }
  } else {
    // h' = (1 - u) h + u tanh(conv)     (graph_vs.py:184-186)
    for (int row = beg; row < end; ++row) {
      const float ht = tanhf(conv(row));
      const float hv = p.h ? p.h[(size_t)row * 128 + c] : 0.f;
      const float uv = p.u_in[(size_t)row * 128 + c];
      p.h_out[(size_t)row * 128 + c] = (1.f - uv) * hv + uv * ht;
    }
  }
}

int edgeconv_dense(const EdgeConvDenseArgs& a, cudaStream_t s) {
  if (a.n_graphs == 0) return CNS_OK;
  const unsigned grid = (unsigned)a.n_graphs;
  if (a.mode == EC_PLAIN) {
    CNS_REQUIRE(a.f_out == 128, "edgeconv_dense: plain mode needs f_out = 128");
    if (a.gn_weight) edgeconv_dense_kernel<128, EC_PLAIN, true><<<grid, 128, 0, s>>>(a);
    else edgeconv_dense_kernel<128, EC_PLAIN, false><<<grid, 128, 0, s>>>(a);
  } else if (a.mode == EC_GATES) {
    CNS_REQUIRE(a.f_out == 256, "gates need f_out = 256");
    edgeconv_dense_kernel<256, EC_GATES, false><<<grid, 256, 0, s>>>(a);
  } else {
    CNS_REQUIRE(a.f_out == 128, "candidate needs f_out = 128");
    edgeconv_dense_kernel<128, EC_CANDIDATE, false><<<grid, 128, 0, s>>>(a);
  }
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns
