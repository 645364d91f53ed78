// Encoder (cns/models/graph_vs.py:190-231): keypoint -> cluster attention
// (PyG PointTransformerConv on the bipartite l0_to_l1 graph) fused per 128-edge tile:
//   in_enc -> pos_nn -> lin / lin_src -> attn_nn -> per-channel segment softmax -> weighted sum.
// No [E01,128] intermediate ever reaches HBM; the only per-edge global traffic is the 2-D
// keypoint coordinates and the two int64 indices.  Edges arrive sorted by destination
// cluster (the reference emits them cluster-major, graph_gen.py:167-182), so the softmax
// is a segmented scan with no atomics; segments cut by a 64-edge piece boundary leave
// (max, sum, acc) partials that enc_finish_kernel merges.
//
// This file holds the fp32 FFMA version (CNS_PREC_FP32): register-tiled 8x8 micro-kernel,
// activations kept k-major in shared memory, weights streamed from L2 with cp.async.
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

constexpr int TE = ENC_TILE_E;     // 128 edges per tile
constexpr int LDT = TE + 4;        // k-major activation buffers: [128 k][LDT]
constexpr int LDF = TE + 1;        // final-phase buffers: [128 o][LDF], scalar access
constexpr int WCH = 32;            // weight rows (k) per cp.async chunk
constexpr int N_CHUNK = 5 * (H / WCH);

struct EncSmem {
  float bufA[H * LDT];
  float bufB[H * LDT];
  float wbuf[2][WCH * H];
  float w_in[H * 2];
  float b_in[H];
  float w_p0[H * 4];
  float b_p0[H];
  float dp[TE * 4];
  float feat[TE * 2];
  int ci[TE];
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[a][b]: edge e = (a<4 ? eg*4+a : 64+eg*4+a-4), channel o = (b<4 ? og*4+b : 64+og*4+b-4)
__device__ __forceinline__ int e_of(int eg, int a) { return (a < 4) ? eg * 4 + a : 64 + eg * 4 + (a - 4); }
__device__ __forceinline__ int o_of(int og, int b) { return (b < 4) ? og * 4 + b : 64 + og * 4 + (b - 4); }

__device__ __forceinline__ void chunk_fma(float (&acc)[8][8], const float* __restrict__ As,
                                          const float* __restrict__ Ws, int eg, int og) {
#pragma unroll 8
  for (int kk = 0; kk < WCH; ++kk) {
    float4 a0 = *reinterpret_cast<const float4*>(As + kk * LDT + eg * 4);
    float4 a1 = *reinterpret_cast<const float4*>(As + kk * LDT + 64 + eg * 4);
    float4 w0 = *reinterpret_cast<const float4*>(Ws + kk * H + og * 4);
    float4 w1 = *reinterpret_cast<const float4*>(Ws + kk * H + 64 + og * 4);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
  }
}

// store an 8x8 register tile k-major ([o][e], stride LDT) for use as the next GEMM's A operand
__device__ __forceinline__ void store_kmajor(float* buf, const float (&v)[8][8], int eg, int og) {
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    int o = o_of(og, b);
    *reinterpret_cast<float4*>(buf + o * LDT + eg * 4) = make_float4(v[0][b], v[1][b], v[2][b], v[3][b]);
    *reinterpret_cast<float4*>(buf + o * LDT + 64 + eg * 4) = make_float4(v[4][b], v[5][b], v[6][b], v[7][b]);
  }
}
// store for the final phase: [o][e] with odd stride LDF (bank-conflict-free both ways)
__device__ __forceinline__ void store_final(float* buf, const float (&v)[8][8], int eg, int og) {
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    int o = o_of(og, b);
#pragma unroll
    for (int a = 0; a < 8; ++a) buf[o * LDF + e_of(eg, a)] = v[a][b];
  }
}

__global__ void __launch_bounds__(256, 1)
enc_edge_fp32_kernel(EncEdgeArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  EncSmem& sm = *reinterpret_cast<EncSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int side = blockIdx.y;  // 0: current frame, 1: target frame
  const int64_t e0 = (int64_t)tile * TE;
  const int n_valid = (int)min((int64_t)TE, p.n_e01 - e0);
  const float* __restrict__ feat_src = side == 0 ? p.x_cur : p.x_tar;
  const float* __restrict__ pos_src = side == 0 ? p.pos_cur : p.pos_tar;
  const int eg = tid & 15, og = tid >> 4;

  // weight chunk stream: enc_chain is [5*128 k][128 o]; chunk q = rows [32q, 32q+32)
  auto prefetch = [&](int q) {
    const float* src = p.enc_chain + (size_t)q * WCH * H;
    float* dst = sm.wbuf[q & 1];
#pragma unroll
    for (int i = 0; i < (WCH * H / 4) / 256; ++i) {
      int idx = tid + i * 256;
      cp_async16(dst + idx * 4, src + idx * 4);
    }
    cp_async_commit();
  };
  prefetch(0);

  // ---- stage 0: small weights + per-edge metadata ---------------------------------------------
  for (int i = tid; i < H * 2; i += 256) sm.w_in[i] = p.w_in[i];
  for (int i = tid; i < H * 4; i += 256) sm.w_p0[i] = p.w_p0[i];
  if (tid < H) { sm.b_in[tid] = p.b_in[tid]; sm.b_p0[tid] = p.b_p0[tid]; }
  if (tid < TE) {
    int e = tid;
    int ci = -1;
    float f0 = 0.f, f1 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (e < n_valid) {
      int64_t j = p.ej[e0 + e];
      ci = (int)p.ei[e0 + e];
      float2 f = reinterpret_cast<const float2*>(feat_src)[j];
      float2 ps = reinterpret_cast<const float2*>(pos_src)[j];
      float2 pt = reinterpret_cast<const float2*>(p.pos_tar)[j];
      float2 pc = reinterpret_cast<const float2*>(p.pos_clu)[ci];
      f0 = f.x; f1 = f.y;
      d0 = pc.x - ps.x; d1 = pc.y - ps.y; d2 = pc.x - pt.x; d3 = pc.y - pt.y;   // pos_i - pos_j
    }
    sm.ci[e] = ci;
    sm.feat[e * 2] = f0; sm.feat[e * 2 + 1] = f1;
    sm.dp[e * 4] = d0; sm.dp[e * 4 + 1] = d1; sm.dp[e * 4 + 2] = d2; sm.dp[e * 4 + 3] = d3;
  }
  __syncthreads();

  // ---- stage 1: x^T = relu(in_enc(feat)) -> bufA ; hp^T = relu(pos_nn.0(dp)) -> bufB -----------
  {
    int e = tid & (TE - 1);
    int k0 = (tid >> 7) * 64;
    float f0 = sm.feat[e * 2], f1 = sm.feat[e * 2 + 1];
    float d0 = sm.dp[e * 4], d1 = sm.dp[e * 4 + 1], d2 = sm.dp[e * 4 + 2], d3 = sm.dp[e * 4 + 3];
#pragma unroll 4
    for (int k = k0; k < k0 + 64; ++k) {
      float x = fmaf(sm.w_in[k * 2 + 1], f1, fmaf(sm.w_in[k * 2], f0, sm.b_in[k]));
      float hp = sm.b_p0[k];
      hp = fmaf(sm.w_p0[k * 4], d0, hp);
      hp = fmaf(sm.w_p0[k * 4 + 1], d1, hp);
      hp = fmaf(sm.w_p0[k * 4 + 2], d2, hp);
      hp = fmaf(sm.w_p0[k * 4 + 3], d3, hp);
      sm.bufA[k * LDT + e] = fmaxf(x, 0.f);
      sm.bufB[k * LDT + e] = fmaxf(hp, 0.f);
    }
  }

  float acc[8][8];
  float dlt[8][8];
  int q = 0;
  auto gemm = [&](const float* As) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < H / WCH; ++c, ++q) {
      if (q + 1 < N_CHUNK) { prefetch(q + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
      __syncthreads();
      chunk_fma(acc, As + c * WCH * LDT, sm.wbuf[q & 1], eg, og);
      __syncthreads();
    }
  };

  // ---- G0: delta = hp @ pos_nn.1^T + b ----------------------------------------------------------
  gemm(sm.bufB);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float bias = p.b_p1[o_of(og, b)];
#pragma unroll
    for (int a = 0; a < 8; ++a) dlt[a][b] = acc[a][b] + bias;
  }
  // ---- G1: t = a_dst[i] - x @ lin_src^T + delta -> bufB (hp is dead) ----------------------------
  gemm(sm.bufA);
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    int ci = sm.ci[e_of(eg, a)];
    if (ci >= 0) {
      float4 d0 = *reinterpret_cast<const float4*>(p.a_dst + (size_t)ci * H + og * 4);
      float4 d1 = *reinterpret_cast<const float4*>(p.a_dst + (size_t)ci * H + 64 + og * 4);
      float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = d[b] + acc[a][b] + dlt[a][b];
    } else {
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
    }
  }
  store_kmajor(sm.bufB, acc, eg, og);
  // ---- G2: val = x @ lin^T + delta -> bufA (final layout; x is dead after this GEMM) ------------
  gemm(sm.bufA);
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] += dlt[a][b];
  store_final(sm.bufA, acc, eg, og);
  // ---- G3: u = relu(t @ attn_nn.0^T + b) -> bufB -------------------------------------------------
  gemm(sm.bufB);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float bias = p.b_a0[o_of(og, b)];
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a][b] = fmaxf(acc[a][b] + bias, 0.f);
  }
  store_kmajor(sm.bufB, acc, eg, og);
  // ---- G4: logits = u @ attn_nn.1^T + b -> bufB (final layout) -----------------------------------
  gemm(sm.bufB);
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float bias = p.b_a1[o_of(og, b)];
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a][b] += bias;
  }
  store_final(sm.bufB, acc, eg, og);
  __syncthreads();

  // ---- per-channel segment softmax + weighted sum over each 64-edge piece -------------------------
  {
    const int c = tid & (H - 1);
    const int half = tid >> 7;
    const int e_beg = half * ENC_PIECE_E;
    const int e_end = min(e_beg + ENC_PIECE_E, n_valid);
    const float* L = sm.bufB + c * LDF;
    const float* V = sm.bufA + c * LDF;
    float* out = p.clu_feat + (size_t)side * p.n_clusters * H;
    float* part = p.partial + (size_t)side * p.n_slots * 3 * H;
    const int64_t piece = (int64_t)tile * 2 + half;
    int e = e_beg;
    while (e < e_end) {
      const int seg = sm.ci[e];
      int e2 = e;
      float m = -INFINITY;
      for (; e2 < e_end && sm.ci[e2] == seg; ++e2) m = fmaxf(m, L[e2]);
      float s = 0.f, a = 0.f;
      for (int k = e; k < e2; ++k) {
        float pr = expf(L[k] - m);
        s += pr;
        a = fmaf(pr, V[k], a);
      }
      const bool complete = (p.clu_rowptr[seg] == (int)(e0 + e)) && (p.clu_rowptr[seg + 1] == (int)(e0 + e2));
      if (complete) {
        out[(size_t)seg * H + c] = fmaxf(a / (s + 1e-16f), 0.f);
      } else {
        float* slot = part + ((size_t)seg + piece) * 3 * H;
        slot[c] = m; slot[H + c] = s; slot[2 * H + c] = a;
      }
      e = e2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// per-cluster pre-pass: x_ctr = relu(in_enc(x_tar[centre])), a_dst = x_ctr @ lin_dst^T,
// pos_clu = pos_tar[centre], cluster -> graph id.   (graph_vs.py:207-208,312-313)
// ---------------------------------------------------------------------------------------------
constexpr int CLU_ROWS = 8;
constexpr int PRE_ROWS = 16;     // clusters per CTA of the pre-pass

__global__ void __launch_bounds__(128)
enc_cluster_pre_kernel(EncPreArgs p) {
  __shared__ __align__(16) float xs[PRE_ROWS][H];
  const int o = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * PRE_ROWS;
  float w0 = p.w_in[o * 2], w1 = p.w_in[o * 2 + 1], b = p.b_in[o];
#pragma unroll
  for (int r = 0; r < PRE_ROWS; ++r) {
    int64_t c = c0 + r;
    float v = 0.f;
    if (c < p.n_clusters) {
      int64_t ctr = p.centers[c];
      float2 f = reinterpret_cast<const float2*>(p.x_tar)[ctr];
      v = fmaxf(fmaf(w1, f.y, fmaf(w0, f.x, b)), 0.f);
      if (o == 0) {
        reinterpret_cast<float2*>(p.pos_clu)[c] = reinterpret_cast<const float2*>(p.pos_tar)[ctr];
        p.clu_graph[c] = p.node_graph ? (int)p.node_graph[ctr] : 0;
      }
    }
    xs[r][o] = v;
  }
  __syncthreads();
  float acc[PRE_ROWS];
  const float ab = p.a_bias ? p.a_bias[o] : 0.f;
#pragma unroll
  for (int r = 0; r < PRE_ROWS; ++r) acc[r] = ab;
  // the row values are the same for every thread: one 16-byte broadcast load feeds four FMAs (a scalar
  // load per FMA makes the kernel shared-memory bound)
#pragma unroll 2
  for (int k = 0; k < H; k += 4) {
    const float wa = p.lin_dst_t[k * H + o], wb = p.lin_dst_t[(k + 1) * H + o];
    const float wc = p.lin_dst_t[(k + 2) * H + o], wd = p.lin_dst_t[(k + 3) * H + o];
#pragma unroll
    for (int r = 0; r < PRE_ROWS; ++r) {
      const float4 x = *reinterpret_cast<const float4*>(&xs[r][k]);
      acc[r] = fmaf(x.w, wd, fmaf(x.z, wc, fmaf(x.y, wb, fmaf(x.x, wa, acc[r]))));
    }
  }
#pragma unroll
  for (int r = 0; r < PRE_ROWS; ++r)
    if (c0 + r < p.n_clusters) p.a_dst[(c0 + r) * H + o] = acc[r];
}

// ---------------------------------------------------------------------------------------------
// finish: merge softmax partials of clusters cut by piece boundaries, then
// x_clu = relu(out_enc([cur, tar - cur])) * cluster_mask        (graph_vs.py:229-230)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
enc_finish_kernel(EncFinishArgs p) {
  __shared__ float cat[CLU_ROWS][2 * H];
  const int o = threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.x * CLU_ROWS;
#pragma unroll 1
  for (int r = 0; r < CLU_ROWS; ++r) {
    int64_t c = c0 + r;
    float f[2] = {0.f, 0.f};
    if (c < p.n_clusters) {
      int beg = p.clu_rowptr[c], end = p.clu_rowptr[c + 1];
      if (end > beg) {
        int p0 = beg / p.piece_e, p1 = (end - 1) / p.piece_e;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          float* feat = p.clu_feat + ((size_t)side * p.n_clusters + c) * H;
          if (p0 == p1 || (p.cur_only && side == 1)) {      // (target side from a cache: already final)
            if (!p.fix_only) f[side] = feat[o];
          } else {
            const float* part = p.partial + (size_t)side * p.n_slots * 3 * H;
            float m = -INFINITY;
            for (int pc = p0; pc <= p1; ++pc) m = fmaxf(m, part[((size_t)c + pc) * 3 * H + o]);
            float s = 0.f, a = 0.f;
            for (int pc = p0; pc <= p1; ++pc) {
              const float* slot = part + ((size_t)c + pc) * 3 * H;
              float w = expf(slot[o] - m);
              s = fmaf(slot[H + o], w, s);
              a = fmaf(slot[2 * H + o], w, a);
            }
            f[side] = fmaxf(a / (s + 1e-16f), 0.f);
            feat[o] = f[side];
          }
        }
      } else {
        p.clu_feat[((size_t)0 * p.n_clusters + c) * H + o] = 0.f;
        p.clu_feat[((size_t)1 * p.n_clusters + c) * H + o] = 0.f;
      }
    }
    if (p.fix_only) continue;
    cat[r][o] = f[0];
    cat[r][H + o] = f[1] - f[0];
    if (p.cat_out && c < p.n_clusters) {
      p.cat_out[c * 2 * H + o] = f[0];
      p.cat_out[c * 2 * H + H + o] = f[1] - f[0];
    }
  }
  if (p.cat_out || p.fix_only) return;
  __syncthreads();
  float acc[CLU_ROWS];
  float b = p.b_out[o];
#pragma unroll
  for (int r = 0; r < CLU_ROWS; ++r) acc[r] = b;
#pragma unroll 4
  for (int k = 0; k < 2 * H; ++k) {
    float w = p.out_t[k * H + o];
#pragma unroll
    for (int r = 0; r < CLU_ROWS; ++r) acc[r] = fmaf(cat[r][k], w, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < CLU_ROWS; ++r) {
    int64_t c = c0 + r;
    if (c < p.n_clusters) p.x_clu[c * H + o] = p.cluster_mask[c] ? fmaxf(acc[r], 0.f) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------
int enc_cluster_pre(const EncPreArgs& a, cudaStream_t s) {
  if (a.n_clusters == 0) return CNS_OK;
  enc_cluster_pre_kernel<<<ceil_div(a.n_clusters, PRE_ROWS), 128, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int enc_edge_fp32(const EncEdgeArgs& a, cudaStream_t s) {
  if (a.n_e01 == 0) return CNS_OK;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(enc_edge_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(EncSmem)));
  }
  dim3 grid(ceil_div(a.n_e01, TE), 2);
  enc_edge_fp32_kernel<<<grid, 256, sizeof(EncSmem), s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// Repair only (the fused backbone applies out_enc itself): one thread per cluster finds the clusters that need
// work - no edges (feature 0) or cut by a softmax-piece boundary (merge the pieces) - and the CTA's 256 threads
// (side, channel) then handle those few.
constexpr int FIX_CLUSTERS = 64;     // clusters per CTA: the flagged ones are a serial chain per warp, so spread them over many CTAs
__global__ void __launch_bounds__(256)
enc_fix_kernel(EncFinishArgs p) {
  __shared__ int todo[FIX_CLUSTERS];
  __shared__ int n_todo;
  const int tid = threadIdx.x;
  if (tid == 0) n_todo = 0;
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * FIX_CLUSTERS + tid;
  if (tid < FIX_CLUSTERS && c < p.n_clusters) {
    const int beg = p.clu_rowptr[c], end = p.clu_rowptr[c + 1];
    const bool empty = end == beg;
    const bool cut = !empty && (beg / p.piece_e != (end - 1) / p.piece_e);
    if (empty || cut) todo[atomicAdd(&n_todo, 1)] = tid;
  }
  __syncthreads();
  // one warp per flagged (cluster, side): a lane owns 4 channels
  const int warp = tid >> 5, lane = tid & 31;
  for (int item = warp; item < 2 * n_todo; item += 8) {
    const int side = item & 1;
    const int64_t cc = (int64_t)blockIdx.x * FIX_CLUSTERS + todo[item >> 1];
    const int beg = p.clu_rowptr[cc], end = p.clu_rowptr[cc + 1];
    float4* feat = reinterpret_cast<float4*>(p.clu_feat + ((size_t)side * p.n_clusters + cc) * H) + lane;
    if (end == beg) { *feat = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
    if (p.cur_only && side == 1) continue;               // target side from a cache: already final
    const int p0 = beg / p.piece_e, p1 = (end - 1) / p.piece_e;
    const float* part = p.partial + (size_t)side * p.n_slots * 3 * H;
    auto slot4 = [&](int pc, int which) { return reinterpret_cast<const float4*>(part + ((size_t)cc + pc) * 3 * H + which * H)[lane]; };
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int pc = p0; pc <= p1; ++pc) {
      const float4 v = slot4(pc, 0);
      m = make_float4(fmaxf(m.x, v.x), fmaxf(m.y, v.y), fmaxf(m.z, v.z), fmaxf(m.w, v.w));
    }
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), a = s;
    for (int pc = p0; pc <= p1; ++pc) {
      const float4 lm = slot4(pc, 0), ls = slot4(pc, 1), la = slot4(pc, 2);
      const float4 w = make_float4(expf(lm.x - m.x), expf(lm.y - m.y), expf(lm.z - m.z), expf(lm.w - m.w));
      s = make_float4(fmaf(ls.x, w.x, s.x), fmaf(ls.y, w.y, s.y), fmaf(ls.z, w.z, s.z), fmaf(ls.w, w.w, s.w));
      a = make_float4(fmaf(la.x, w.x, a.x), fmaf(la.y, w.y, a.y), fmaf(la.z, w.z, a.z), fmaf(la.w, w.w, a.w));
    }
    *feat = make_float4(fmaxf(a.x / (s.x + 1e-16f), 0.f), fmaxf(a.y / (s.y + 1e-16f), 0.f),
                        fmaxf(a.z / (s.z + 1e-16f), 0.f), fmaxf(a.w / (s.w + 1e-16f), 0.f));
  }
}

int enc_finish(const EncFinishArgs& a, cudaStream_t s) {
  if (a.n_clusters == 0) return CNS_OK;
  if (a.fix_only) {
    enc_fix_kernel<<<ceil_div(a.n_clusters, FIX_CLUSTERS), 256, 0, s>>>(a);
    CNS_LAUNCH_CHECK();
    return CNS_OK;
  }
  enc_finish_kernel<<<ceil_div(a.n_clusters, CLU_ROWS), 128, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns
