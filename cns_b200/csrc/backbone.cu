// Cluster-level backbone + decoder kernels (cns/models/graph_vs.py:46-187,234-270).
//
//  linear        : batched node MLP  Y = act([X1|X2] W^T + pos Wp + b + res)      (fp32 FFMA tile GEMM)
//  edgeconv_max  : PointEdgeConv aggregation in factorised form, CSR by destination, one warp per
//                  destination row, float4 lanes, no atomics; GRU gate / candidate math fused in
//  graphnorm_relu: PyG GraphNorm + ReLU, one CTA per graph
//  pool_head     : AttentionalAggregation + both velocity heads + postprocess, one CTA per graph
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

// =============================================================================================
// linear: 64x64 output tile, BK=16, 256 threads, 4x4 register tile
// =============================================================================================
constexpr int LBM = 64, LBN = 64, LBK = 16;

__global__ void __launch_bounds__(256)
linear_kernel(LinearArgs p) {
  __shared__ __align__(16) float Xs[LBK][LBM + 4];
  __shared__ __align__(16) float Ws[LBK][LBN];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * LBM;
  const int n0 = blockIdx.x * LBN;
  const int tx = tid & 15, ty = tid >> 4;
  const int K = p.k1 + p.k2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lrow = tid >> 2, lk = (tid & 3) * 4;   // X tile loader: 64 rows x 4 float4
  const int wk = tid >> 4, wn = (tid & 15) * 4;    // W tile loader: 16 rows x 16 float4
  for (int k0 = 0; k0 < K; k0 += LBK) {
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t m = m0 + lrow;
    if (m < p.m) {
      int k = k0 + lk;
      const float* src = (k < p.k1) ? p.x1 + m * p.k1 + k : p.x2 + m * p.k2 + (k - p.k1);
      xv = *reinterpret_cast<const float4*>(src);
    }
    float4 wv = *reinterpret_cast<const float4*>(p.wt + (size_t)(k0 + wk) * p.n_out + n0 + wn);
    __syncthreads();
    Xs[lk][lrow] = xv.x; Xs[lk + 1][lrow] = xv.y; Xs[lk + 2][lrow] = xv.z; Xs[lk + 3][lrow] = xv.w;
    *reinterpret_cast<float4*>(&Ws[wk][wn]) = wv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LBK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  const int n = n0 + tx * 4;
  float4 bias = p.bias ? *reinterpret_cast<const float4*>(p.bias + n) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 wp0 = make_float4(0.f, 0.f, 0.f, 0.f), wp1 = wp0;
  if (p.pos) {
    wp0 = *reinterpret_cast<const float4*>(p.wp + n);
    wp1 = *reinterpret_cast<const float4*>(p.wp + p.n_out + n);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= p.m) continue;
    float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
    if (p.pos) {
      float2 ps = reinterpret_cast<const float2*>(p.pos)[m];
      v.x = fmaf(ps.y, wp1.x, fmaf(ps.x, wp0.x, v.x));
      v.y = fmaf(ps.y, wp1.y, fmaf(ps.x, wp0.y, v.y));
      v.z = fmaf(ps.y, wp1.z, fmaf(ps.x, wp0.z, v.z));
      v.w = fmaf(ps.y, wp1.w, fmaf(ps.x, wp0.w, v.w));
    }
    if (p.residual) {
      float4 r = *reinterpret_cast<const float4*>(p.residual + m * p.n_out + n);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (p.relu_gate) {
      const float4 g = *reinterpret_cast<const float4*>(p.relu_gate + m * p.n_out + n);
      v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f; v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
    }
    *reinterpret_cast<float4*>(p.y + m * p.n_out + n) = v;
  }
}

int linear(const LinearArgs& a, cudaStream_t s) {
  if (a.m == 0) return CNS_OK;
  CNS_REQUIRE(a.n_out % LBN == 0, "linear: n_out must be a multiple of 64");
  CNS_REQUIRE(a.k1 % LBK == 0 && a.k2 % LBK == 0 && a.k1 > 0, "linear: k must be a multiple of 16");
  CNS_REQUIRE(a.k2 == 0 || a.x2 != nullptr, "linear: x2 missing");
  dim3 grid(a.n_out / LBN, ceil_div(a.m, LBM));
  linear_kernel<<<grid, 256, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// edgeconv_max: out_i = deg(i) ? P_i + max_{j in N(i)} Q_j : 0     (PQ row = [P (f) | Q (f)])
// one warp per destination row; VPL float4 per lane (f = 128 -> 1, f = 256 -> 2)
// =============================================================================================
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int VPL>
__global__ void __launch_bounds__(256)
edgeconv_max_kernel(EdgeConvArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.n_rows) return;
  constexpr int F = VPL * 128;
  const int beg = p.rowptr[row], end = p.rowptr[row + 1];
  float4 best[VPL];
  int4 arg[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    best[v] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    arg[v] = make_int4(-1, -1, -1, -1);
  }
  for (int base = beg; base < end; base += 32) {
    int my = (base + lane < end) ? p.src[base + lane] : -1;
    int cnt = min(32, end - base);
    for (int t = 0; t < cnt; ++t) {
      int j = __shfl_sync(0xffffffffu, my, t);
      const float4* q = reinterpret_cast<const float4*>(p.pq + (size_t)j * 2 * F + F);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float4 x = q[v * 32 + lane];
        if (x.x > best[v].x) { best[v].x = x.x; arg[v].x = j; }
        if (x.y > best[v].y) { best[v].y = x.y; arg[v].y = j; }
        if (x.z > best[v].z) { best[v].z = x.z; arg[v].z = j; }
        if (x.w > best[v].w) { best[v].w = x.w; arg[v].w = j; }
      }
    }
  }
  const bool has = end > beg;
  const float4* pr = reinterpret_cast<const float4*>(p.pq + (size_t)row * 2 * F);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has) {
      float4 pv = pr[v * 32 + lane];
      o = make_float4(pv.x + best[v].x, pv.y + best[v].y, pv.z + best[v].z, pv.w + best[v].w);
    }
    const int c4 = v * 32 + lane;          // float4 column index in [0, F/4)
    if (p.argmax) reinterpret_cast<int4*>(p.argmax + (size_t)row * F)[c4] = arg[v];
    if (p.mode == EC_PLAIN) {
      reinterpret_cast<float4*>(p.out + (size_t)row * F)[c4] = o;
    } else if (p.mode == EC_GATES) {
      // gates = sigmoid(conv) ; r = first 128 channels, u = last 128 (torch.chunk, graph_vs.py:180-181)
      float4 g = make_float4(sigmoidf_(o.x), sigmoidf_(o.y), sigmoidf_(o.z), sigmoidf_(o.w));
      if (p.act) reinterpret_cast<float4*>(p.act + (size_t)row * F)[c4] = g;
      if (c4 < 32) {
        float4 hv = p.h ? reinterpret_cast<const float4*>(p.h + (size_t)row * 128)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(p.hr + (size_t)row * 128)[c4] = make_float4(hv.x * g.x, hv.y * g.y, hv.z * g.z, hv.w * g.w);
      } else {
        reinterpret_cast<float4*>(p.u + (size_t)row * 128)[c4 - 32] = g;
      }
    } else {
      // h' = (1 - u) h + u tanh(conv)     (graph_vs.py:184-186)
      float4 ht = make_float4(tanhf(o.x), tanhf(o.y), tanhf(o.z), tanhf(o.w));
      if (p.act) reinterpret_cast<float4*>(p.act + (size_t)row * F)[c4] = ht;
      float4 hv = p.h ? reinterpret_cast<const float4*>(p.h + (size_t)row * 128)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 uv = reinterpret_cast<const float4*>(p.u_in + (size_t)row * 128)[c4];
      float4 hn;
      hn.x = (1.f - uv.x) * hv.x + uv.x * ht.x;
      hn.y = (1.f - uv.y) * hv.y + uv.y * ht.y;
      hn.z = (1.f - uv.z) * hv.z + uv.z * ht.z;
      hn.w = (1.f - uv.w) * hv.w + uv.w * ht.w;
      reinterpret_cast<float4*>(p.h_out + (size_t)row * 128)[c4] = hn;
    }
  }
}

int edgeconv_max(const EdgeConvArgs& a, cudaStream_t s) {
  if (a.n_rows == 0) return CNS_OK;
  CNS_REQUIRE(a.f_out == 128 || a.f_out == 256, "edgeconv_max: f_out must be 128 or 256");
  CNS_REQUIRE(a.mode != EC_GATES || a.f_out == 256, "gates need f_out = 256");
  CNS_REQUIRE(a.mode != EC_CANDIDATE || a.f_out == 128, "candidate needs f_out = 128");
  const int warps = 8;
  int grid = ceil_div(a.n_rows, warps);
  if (a.f_out == 128) edgeconv_max_kernel<1><<<grid, warps * 32, 0, s>>>(a);
  else edgeconv_max_kernel<2><<<grid, warps * 32, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// GraphNorm + ReLU, one CTA (128 threads = channels) per graph.  Statistics run over every
// cluster row of the graph, masked (all-zero) rows included, like the reference.
// =============================================================================================
__global__ void __launch_bounds__(128)
graphnorm_relu_kernel(GraphNormArgs p) {
  const int g = blockIdx.x, c = threadIdx.x;
  const int beg = p.graph_ptr[g], end = p.graph_ptr[g + 1];
  const int n = end - beg;
  if (n <= 0) return;
  const float inv_n = 1.f / (float)n;
  float sum = 0.f;
  for (int r = beg; r < end; ++r) sum += p.x[(size_t)r * H + c];
  const float shift = sum * inv_n * p.mean_scale[c];
  float sq = 0.f;
  for (int r = beg; r < end; ++r) {
    float o = p.x[(size_t)r * H + c] - shift;
    sq = fmaf(o, o, sq);
  }
  const float rstd = 1.f / sqrtf(sq * inv_n + 1e-5f);
  const float w = p.weight[c], b = p.bias[c];
  for (int r = beg; r < end; ++r) {
    float o = p.x[(size_t)r * H + c] - shift;
    p.y[(size_t)r * H + c] = fmaxf(fmaf(w * o, rstd, b), 0.f);
  }
  if (p.stats) {
    p.stats[((size_t)g * 2) * H + c] = sum * inv_n;
    p.stats[((size_t)g * 2 + 1) * H + c] = rstd;
  }
}

int graphnorm_relu(const GraphNormArgs& a, cudaStream_t s) {
  if (a.n_graphs == 0) return CNS_OK;
  graphnorm_relu_kernel<<<(unsigned)a.n_graphs, 128, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// pool_head: softmax-gated sum over the graph's un-masked clusters, two 2-layer heads, and the
// velocity post-processing  vel = vec/(|vec|+1e-7) * (1+elu(norm)), vel[:3] *= tPo_norm
// (graph_vs.py:254-270, functions.py:26-38).
// A CTA of 8 warps owns 8 graphs: warp w pools graph w (a lane owns 4 channels, rows are read as whole 512-byte
// lines), then the 256 threads run the first layer of both heads for all 8 graphs at once - one weight load (from L2)
// feeds 8 FMAs (one CTA per graph re-read the two 64 KB head matrices from L2 for every graph: 134 MB per 1024 graphs).
// A graph whose clusters are all masked yields x_scene = 0 (PyG would drop the row).
// =============================================================================================
constexpr int PH_G = 8;          // graphs per CTA

// 72 registers x 256 threads + 13.5 KB: fits beside a resident encoder CTA (640 threads x 72 registers, 207 KB)
__global__ void __maxnreg__(72)
pool_head_kernel(PoolHeadArgs p) {
  __shared__ __align__(16) float xs[PH_G][H];
  __shared__ __align__(16) float hid[2][PH_G][H];
  __shared__ float outv[PH_G][8];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t g0 = (int64_t)blockIdx.x * PH_G;

  // ---- pooling: warp = graph ----------------------------------------------------------------------------
  {
    const int64_t g = g0 + wid;
    float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
    float den = 0.f;
    if (g < p.n_graphs) {
      const int beg = p.graph_ptr[g], end = p.graph_ptr[g + 1];
      const float4 gw = reinterpret_cast<const float4*>(p.pool_w)[lane];
      const float gb = p.pool_b[0];
      // online softmax over the un-masked rows (one pass: the row is in registers when its gate is known); the mask
      // bytes of 32 rows arrive with one load and four rows are in flight at a time - a load per row behind a mask
      // byte behind the previous row made this loop a chain of L2 round trips
      float m = -INFINITY;
      auto take = [&](const float4 x, int r) {
        float d = x.x * gw.x + x.y * gw.y + x.z * gw.z + x.w * gw.w;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
        d += gb;
        if (lane == 0 && p.gate) p.gate[r] = d;
        const float mn = fmaxf(m, d);
        const float sc = expf(m - mn), e = expf(d - mn);      // m = -inf on the first row: sc = 0
        den = fmaf(den, sc, e);
        num.x = fmaf(num.x, sc, e * x.x); num.y = fmaf(num.y, sc, e * x.y);
        num.z = fmaf(num.z, sc, e * x.z); num.w = fmaf(num.w, sc, e * x.w);
        m = mn;
      };
      for (int base = beg; base < end; base += 32) {
        const int rr = base + lane;
        uint32_t bits = __ballot_sync(0xffffffffu, rr < end && p.cluster_mask[rr]);
        while (bits) {
          int r[4];
          float4 x[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            r[i] = -1;
            if (bits) { r[i] = base + __ffs(bits) - 1; bits &= bits - 1; }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (r[i] >= 0) x[i] = reinterpret_cast<const float4*>(p.x + (size_t)r[i] * H)[lane];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (r[i] >= 0) take(x[i], r[i]);
        }
      }
    }
    const float inv = (den > 0.f) ? 1.f / (den + 1e-16f) : 0.f;
    const float4 xsc = make_float4(num.x * inv, num.y * inv, num.z * inv, num.w * inv);
    reinterpret_cast<float4*>(xs[wid])[lane] = xsc;
    if (p.x_scene && g < p.n_graphs) reinterpret_cast<float4*>(p.x_scene + (size_t)g * H)[lane] = xsc;
  }
  __syncthreads();

  // ---- first layers of both heads: thread = (head, output channel), all 8 graphs per weight load ---------
  {
    const int o = tid & (H - 1), hd = tid >> 7;
    const float* __restrict__ wt = hd ? p.nrm0_t : p.vec0_t;
    const float b = hd ? p.nrm0_b[o] : p.vec0_b[o];
    float acc[PH_G];
#pragma unroll
    for (int g = 0; g < PH_G; ++g) acc[g] = b;
    // the weights come straight from L2 (a warp reads 128 consecutive bytes per k), 16 loads in flight per thread: the
    // kernel needs no large shared-memory block, so its CTAs fit beside the resident encoder / backbone CTAs of a
    // neighbouring forward instead of waiting for a free SM (staged through 128 KB of shared memory it held 128 SMs
    // for 17 us per step)
    float wn[16];                               // the next round's weights, loaded while this round's FMAs run
#pragma unroll
    for (int i = 0; i < 16; ++i) wn[i] = __ldg(wt + i * H + o);
#pragma unroll 1
    for (int k0 = 0; k0 < H; k0 += 16) {
      float w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = wn[i];
      if (k0 + 16 < H) {
#pragma unroll
        for (int i = 0; i < 16; ++i) wn[i] = __ldg(wt + (k0 + 16 + i) * H + o);
      }
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
#pragma unroll
        for (int g = 0; g < PH_G; ++g) {
          const float4 x = *reinterpret_cast<const float4*>(&xs[g][k0 + i]);      // broadcast read
          acc[g] = fmaf(x.w, w[i + 3], fmaf(x.z, w[i + 2], fmaf(x.y, w[i + 1], fmaf(x.x, w[i], acc[g]))));
        }
      }
    }
#pragma unroll
    for (int g = 0; g < PH_G; ++g) {
      const float v = fmaxf(acc[g], 0.f);
      hid[hd][g][o] = v;
      if (p.hid && g0 + g < p.n_graphs) p.hid[((size_t)(g0 + g) * 2 + hd) * H + o] = v;
    }
  }
  __syncthreads();

  // ---- second layers + post-processing: warp = graph, 7 dot products of length 128 -----------------------
  {
    const int64_t g = g0 + wid;
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const float* w = (q < 6) ? p.vec1_w + q * H : p.nrm1_w;
      const float* hv = (q < 6) ? hid[0][wid] : hid[1][wid];
      const float4 a = reinterpret_cast<const float4*>(w)[lane];
      const float4 b = reinterpret_cast<const float4*>(hv)[lane];
      float d = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      if (lane == 0) outv[wid][q] = d + ((q < 6) ? p.vec1_b[q] : p.nrm1_b[0]);
    }
    __syncwarp();
    if (g < p.n_graphs) {
      const float* ov = outv[wid];
      if (lane < 6) {
        p.vel_si_vec[(size_t)g * 6 + lane] = ov[lane];
        if (p.vel) {
          float nn = 0.f;
#pragma unroll
          for (int q = 0; q < 6; ++q) nn = fmaf(ov[q], ov[q], nn);
          const float nr = ov[6];
          const float scale = 1.f + (nr > 0.f ? nr : expm1f(nr));
          float v = ov[lane] / (sqrtf(nn) + 1e-7f) * scale;
          if (lane < 3) v *= (p.tPo_norm ? p.tPo_norm[g] : 1.f);
          p.vel[(size_t)g * 6 + lane] = v;
        }
      }
      if (lane == 6) p.vel_si_norm[g] = ov[6];
    }
  }
}

int pool_head(const PoolHeadArgs& a, cudaStream_t s) {
  if (a.n_graphs == 0) return CNS_OK;
  pool_head_kernel<<<(unsigned)ceil_div(a.n_graphs, PH_G), 256, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns

// ---- per-op C ABI -------------------------------------------------------------------------------
using namespace cns;

extern "C" int cns_edgeconv_max(const float* pq, const int32_t* rowptr, const int32_t* src, int64_t n_rows,
                                int f_out, float* out, void* stream) {
  CNS_REQUIRE(pq && rowptr && out, "NULL pointer");
  EdgeConvArgs a{};
  a.n_rows = n_rows; a.f_out = f_out; a.mode = EC_PLAIN;
  a.pq = pq; a.rowptr = rowptr; a.src = src; a.out = out;
  return edgeconv_max(a, (cudaStream_t)stream);
}

extern "C" int cns_graphnorm_relu(const float* x, const int32_t* graph_ptr, int64_t n_graphs,
                                  const float* weight, const float* bias, const float* mean_scale, float* y,
                                  void* stream) {
  CNS_REQUIRE(x && graph_ptr && weight && bias && mean_scale && y, "NULL pointer");
  GraphNormArgs a{};
  a.n_graphs = n_graphs; a.graph_ptr = graph_ptr; a.x = x;
  a.weight = weight; a.bias = bias; a.mean_scale = mean_scale; a.y = y; a.stats = nullptr;
  return graphnorm_relu(a, (cudaStream_t)stream);
}

extern "C" int cns_linear(const float* x1, int k1, const float* x2, int k2, const float* wt, const float* bias,
                          const float* residual, int relu, int64_t m, int n_out, float* y, void* stream) {
  CNS_REQUIRE(x1 && wt && y, "NULL pointer");
  LinearArgs a{};
  a.m = m; a.k1 = k1; a.k2 = k2; a.n_out = n_out;
  a.x1 = x1; a.x2 = x2; a.wt = wt; a.bias = bias; a.pos = nullptr; a.wp = nullptr;
  a.residual = residual; a.relu = relu; a.y = y;
  return linear(a, (cudaStream_t)stream);
}
