// Backward-pass building blocks (fp32, deterministic: every reduction is a fixed-order two-stage sum,
// no floating-point atomics).  Used by backward.cu to differentiate GraphVS.forward
// (cns/models/graph_vs.py:284-322) for the training loops of cns/train_gvs_{short,long}_seq.py.
#include "common.cuh"
#include "kernels.cuh"
#include "bwd.cuh"

namespace cns {

// =============================================================================================
// dw_gemm: out[na][nb] (+)= scale * sum_m A[m][na] * B[m][nb]        (weight gradients dY^T X)
// stage 1: one CTA per (128x128 output tile, row slab) -> partial; stage 2: fixed-order reduce.
// =============================================================================================
constexpr int DW_CH = 16;

__global__ void __launch_bounds__(256)
dw_gemm_kernel(DwGemmArgs p, int rows_per_slab) {
  __shared__ __align__(16) float As[DW_CH][128];
  __shared__ __align__(16) float Bs[DW_CH][128];
  const int tid = threadIdx.x;
  const int tiles_b = p.nb / 128;
  const int ta = blockIdx.x / tiles_b, tb = blockIdx.x % tiles_b;
  const int slab = blockIdx.y;
  const int64_t r0 = (int64_t)slab * rows_per_slab;
  const int64_t r1 = min(p.m, r0 + rows_per_slab);
  const float* __restrict__ A = p.a + ta * 128;
  const float* __restrict__ Bp = (tb < p.b_tiles0) ? p.b0 + tb * 128 : p.b1 + (tb - p.b_tiles0) * 128;
  const int ldb = (tb < p.b_tiles0) ? p.ldb0 : p.ldb1;
  const int ag = tid >> 4, bg = tid & 15;   // 16 x 16 threads, 8 x 8 outputs each (4 + 4 split)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int lr = tid >> 4, lc = (tid & 15) * 8;   // loader: 16 rows x 16 threads x 8 floats
  const bool want_cs = p.colsum_out != nullptr && tb == 0;
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
  for (int64_t base = r0; base < r1; base += DW_CH) {
    const int64_t m = base + lr;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
    if (m < r1) {
      a0 = *reinterpret_cast<const float4*>(A + m * p.lda + lc);
      a1 = *reinterpret_cast<const float4*>(A + m * p.lda + lc + 4);
      b0 = *reinterpret_cast<const float4*>(Bp + m * ldb + lc);
      b1 = *reinterpret_cast<const float4*>(Bp + m * ldb + lc + 4);
    }
    if (want_cs) {     // column sums of A from the values this thread loads anyway (rows lr, lr+16, ... of the slab)
      cs[0] += a0.x; cs[1] += a0.y; cs[2] += a0.z; cs[3] += a0.w; cs[4] += a1.x; cs[5] += a1.y; cs[6] += a1.z; cs[7] += a1.w;
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lr][lc]) = a0; *reinterpret_cast<float4*>(&As[lr][lc + 4]) = a1;
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = b0; *reinterpret_cast<float4*>(&Bs[lr][lc + 4]) = b1;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < DW_CH; ++kk) {
      float4 x0 = *reinterpret_cast<const float4*>(&As[kk][ag * 4]);
      float4 x1 = *reinterpret_cast<const float4*>(&As[kk][64 + ag * 4]);
      float4 y0 = *reinterpret_cast<const float4*>(&Bs[kk][bg * 4]);
      float4 y1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + bg * 4]);
      float a[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      float b[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* part = p.partial + ((size_t)slab * p.na + ta * 128) * p.nb + tb * 128;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ra = (i < 4) ? ag * 4 + i : 64 + ag * 4 + (i - 4);
    *reinterpret_cast<float4*>(part + (size_t)ra * p.nb + bg * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(part + (size_t)ra * p.nb + 64 + bg * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
  if (want_cs) {       // fixed-order sum over the 16 loader rows
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(cs[0], cs[1], cs[2], cs[3]);
    *reinterpret_cast<float4*>(&As[lr][lc + 4]) = make_float4(cs[4], cs[5], cs[6], cs[7]);
    __syncthreads();
    if (tid < 128) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < DW_CH; ++r) t += As[r][tid];
      p.colsum_partial[(size_t)slab * p.na + ta * 128 + tid] = t;
    }
  }
}

// out[r*ld_out + c] (+)= scale * sum_s partial[s][r][c]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n_slabs, int rows, int cols,
                                       float* __restrict__ out, int ld_out, float scale, int accumulate,
                                       const float* __restrict__ cs_partial, float* __restrict__ cs_out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) {
    const int r = idx - rows * cols;               // trailing threads: column sums of A (one per output row)
    if (cs_out == nullptr || r >= rows) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 4 <= n_slabs; k += 4) {
      s0 += cs_partial[(size_t)k * rows + r]; s1 += cs_partial[(size_t)(k + 1) * rows + r];
      s2 += cs_partial[(size_t)(k + 2) * rows + r]; s3 += cs_partial[(size_t)(k + 3) * rows + r];
    }
    for (; k < n_slabs; ++k) s0 += cs_partial[(size_t)k * rows + r];
    const float t = (s0 + s1) + (s2 + s3);
    cs_out[r] = accumulate ? cs_out[r] + t : t;
    return;
  }
  int r = idx / cols, c = idx % cols;
  // four independent chains (fixed order): the loop is pure load latency otherwise
  const size_t stride = (size_t)rows * cols;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = 0;
  for (; k + 4 <= n_slabs; k += 4) {
    s0 += partial[(size_t)k * stride + idx];
    s1 += partial[(size_t)(k + 1) * stride + idx];
    s2 += partial[(size_t)(k + 2) * stride + idx];
    s3 += partial[(size_t)(k + 3) * stride + idx];
  }
  for (; k < n_slabs; ++k) s0 += partial[(size_t)k * stride + idx];
  const float s = (s0 + s1) + (s2 + s3);
  float* o = out + (size_t)r * ld_out + c;
  *o = accumulate ? *o + scale * s : scale * s;
}

int dw_gemm(const DwGemmArgs& a, cudaStream_t s) {
  CNS_REQUIRE(a.na % 128 == 0 && a.nb % 128 == 0, "dw_gemm: dims must be multiples of 128");
  const int tiles = (a.na / 128) * (a.nb / 128);
  const int64_t m1 = a.m > 0 ? a.m : 1;
  int slabs = ceil_div(m1, 64);          // (256-row slabs left 19 CTAs per tile at 4.8 k cluster rows)
  if (a.allow_split && tiles == 1 && a.b_tiles0 >= 1 && a.m >= 4096 && a.lda % 4 == 0 && a.ldb0 % 4 == 0) {
    // long reductions of a single 128 x 128 tile: tensor cores with split operands, one partial per persistent CTA
    int rc = dw_split(a.m, a.a, a.lda, a.b0, a.ldb0, a.partial, a.colsum_out ? a.colsum_partial : nullptr, a.max_slabs, &slabs, s);
    if (rc != CNS_OK) return rc;
  } else {
  if (slabs > 592 / tiles) slabs = 592 / tiles > 0 ? 592 / tiles : 1;
  if (slabs > a.max_slabs) slabs = a.max_slabs;
  const int rows_per_slab = (int)((m1 + slabs - 1) / slabs);
  dim3 grid(tiles, slabs);
  dw_gemm_kernel<<<grid, 256, 0, s>>>(a, rows_per_slab);
  CNS_LAUNCH_CHECK();
  }
  CNS_REQUIRE(a.colsum_out == nullptr || a.colsum_partial != nullptr, "dw_gemm: colsum needs its partial buffer");
  const int total = a.na * a.nb + (a.colsum_out ? a.na : 0);
  reduce_partials_kernel<<<ceil_div(total, 256), 256, 0, s>>>(a.partial, slabs, a.na, a.nb, a.out, a.ld_out, a.scale,
                                                             a.accumulate, a.colsum_partial, a.colsum_out);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// dw_small: out (+)= scale * sum_m A[m][n] * S[m][q]   for q < Q <= 8 (S == NULL: S = 1, Q = 1)
// layout: transposed == 0 -> out[n*ld_out + q] (weight [out][in] with tiny `in`)
//         transposed == 1 -> out[q*ld_out + n] (weight [tiny out][in])
// one thread per column n, a block per row slab; stage 2 as above.
// =============================================================================================
constexpr int DWS_ROWS = 128;   // rows per slab: many small CTAs so that long reductions fill the GPU

__global__ void __launch_bounds__(128)
dw_small_kernel(DwSmallArgs p, int rows_per_slab) {
  __shared__ float ss[DWS_ROWS * 8];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int slab = blockIdx.y;
  const int64_t r0 = (int64_t)slab * rows_per_slab, r1 = min(p.m, r0 + rows_per_slab);
  const int rows = (int)(r1 - r0);
  if (p.s) {
    for (int i = threadIdx.x; i < rows * p.q; i += blockDim.x) ss[i] = p.s[(r0 + i / p.q) * p.lds + i % p.q];
  }
  __syncthreads();
  if (n >= p.n) return;
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  const float* __restrict__ a = p.a + r0 * p.lda + n;
  int r = 0;
  for (; r + 4 <= rows; r += 4) {        // 4 independent loads in flight
    const float a0 = a[(int64_t)r * p.lda], a1 = a[(int64_t)(r + 1) * p.lda];
    const float a2 = a[(int64_t)(r + 2) * p.lda], a3 = a[(int64_t)(r + 3) * p.lda];
    if (p.s) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < p.q) {
          acc[q] = fmaf(a0, ss[r * p.q + q], acc[q]);
          acc[q] = fmaf(a1, ss[(r + 1) * p.q + q], acc[q]);
          acc[q] = fmaf(a2, ss[(r + 2) * p.q + q], acc[q]);
          acc[q] = fmaf(a3, ss[(r + 3) * p.q + q], acc[q]);
        } else if (q == p.q && p.ones_out) {
          acc[q] += (a0 + a1) + (a2 + a3);
        }
    } else {
      acc[0] += (a0 + a1) + (a2 + a3);
    }
  }
  for (; r < rows; ++r) {
    const float a0 = a[(int64_t)r * p.lda];
    if (p.s) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < p.q) acc[q] = fmaf(a0, ss[r * p.q + q], acc[q]);
        else if (q == p.q && p.ones_out) acc[q] += a0;
    } else {
      acc[0] += a0;
    }
  }
  const int qq = p.q + (p.ones_out ? 1 : 0);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    if (q < qq) p.partial[((size_t)slab * p.n + n) * qq + q] = acc[q];
}

// fixed-order reduction over the slabs: a CTA owns 32 consecutive outputs, its 8 warps take the slabs k = w mod 8
// (coalesced 128-byte reads), then the 8 partial sums are added in warp order - deterministic, and ~60x more
// loads in flight than one thread per output walking all slabs.
__global__ void __launch_bounds__(256)
reduce_small_kernel(DwSmallArgs p, int n_slabs) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + lane;
  const int qq = p.q + (p.ones_out ? 1 : 0);
  const int total = p.n * qq;
  float s0 = 0.f, s1 = 0.f;
  if (idx < total) {
    int k = w;
    for (; k + 8 < n_slabs; k += 16) {
      s0 += p.partial[(size_t)k * total + idx];
      s1 += p.partial[(size_t)(k + 8) * total + idx];
    }
    if (k < n_slabs) s0 += p.partial[(size_t)k * total + idx];
  }
  part[w][lane] = s0 + s1;
  __syncthreads();
  if (w == 0 && idx < total) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += part[g][lane];
    const int n = idx / qq, q = idx % qq;
    float* o = q == p.q ? p.ones_out + n
                        : (p.transposed ? p.out + (size_t)q * p.ld_out + n : p.out + (size_t)n * p.ld_out + q);
    *o = p.accumulate ? *o + p.scale * s : p.scale * s;
  }
}

int dw_small(const DwSmallArgs& a, cudaStream_t s) {
  CNS_REQUIRE(a.q >= 1 && a.q <= 8, "dw_small: q must be 1..8");
  CNS_REQUIRE(a.ones_out == nullptr || (a.s != nullptr && a.q <= 7), "dw_small: the fused column sum needs s and q <= 7");
  const int64_t m1 = a.m > 0 ? a.m : 1;
  // short reductions (cluster / graph rows): 32-row slabs - a thread's serial walk over its rows is the whole cost there
  // (128 rows = 32 dependent L2 round trips for 38 CTAs); long ones (one row per keypoint->cluster edge) keep 128
  int slabs = ceil_div(m1, m1 <= 16384 ? 32 : DWS_ROWS);
  if (slabs > a.max_slabs) slabs = a.max_slabs;
  int rows_per_slab = (int)((m1 + slabs - 1) / slabs);
  if (rows_per_slab > DWS_ROWS) {          // more rows than the smem staging holds: use more, smaller slabs
    set_error("dw_small: %lld rows need more than %d slabs", (long long)a.m, a.max_slabs);
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  dim3 grid(ceil_div(a.n, 128), slabs);
  dw_small_kernel<<<grid, 128, 0, s>>>(a, rows_per_slab);
  CNS_LAUNCH_CHECK();
  reduce_small_kernel<<<ceil_div(a.n * (a.q + (a.ones_out ? 1 : 0)), 32), 256, 0, s>>>(a, slabs);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// edgeconv_max backward: dPQ = [dP | dQ];  dP_i = deg_in(i) ? dconv_i : 0,
// dQ_j[c] = sum over out-edges (j -> i) with argmax[i][c] == j of dconv_i[c].  One warp per source
// row j walking its out-neighbourhood (CSR by source): a gather, so no atomics.
// =============================================================================================
template <int VPL>
__global__ void __launch_bounds__(256)
edgeconv_max_bwd_kernel(EdgeConvBwdArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= p.n_rows) return;
  constexpr int F = VPL * 128;
  float4 dq[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) dq[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int beg = p.out_rowptr[j], end = p.out_rowptr[j + 1];
  for (int e = beg; e < end; ++e) {
    const int i = p.out_dst[e];
    const int4* ar = reinterpret_cast<const int4*>(p.argmax + (size_t)i * F);
    const float4* dr = reinterpret_cast<const float4*>(p.dconv + (size_t)i * F);
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int4 a = ar[v * 32 + lane];
      const float4 d = dr[v * 32 + lane];
      if (a.x == (int)j) dq[v].x += d.x;
      if (a.y == (int)j) dq[v].y += d.y;
      if (a.z == (int)j) dq[v].z += d.z;
      if (a.w == (int)j) dq[v].w += d.w;
    }
  }
  const bool has_in = p.in_rowptr[j + 1] > p.in_rowptr[j];
  float4* out = reinterpret_cast<float4*>(p.dpq + (size_t)j * 2 * F);
  const float4* dme = reinterpret_cast<const float4*>(p.dconv + (size_t)j * F);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    out[v * 32 + lane] = has_in ? dme[v * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    out[F / 4 + v * 32 + lane] = dq[v];
  }
}

int edgeconv_max_bwd(const EdgeConvBwdArgs& a, cudaStream_t s) {
  if (a.n_rows == 0) return CNS_OK;
  const int warps = 8;
  const int grid = ceil_div(a.n_rows, warps);
  if (a.f_out == 128) edgeconv_max_bwd_kernel<1><<<grid, warps * 32, 0, s>>>(a);
  else edgeconv_max_bwd_kernel<2><<<grid, warps * 32, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// GraphNorm + ReLU backward (one CTA per graph; thread = channel).
//   n = (x - mean*ms) * rstd ; y = relu(w n + b)
// =============================================================================================
__global__ void __launch_bounds__(128)
graphnorm_relu_bwd_kernel(GraphNormBwdArgs p) {
  const int g = blockIdx.x, c = threadIdx.x;
  const int beg = p.graph_ptr[g], end = p.graph_ptr[g + 1];
  const int n = end - beg;
  float dw = 0.f, db = 0.f, dms = 0.f;
  if (n > 0) {
    const float mean = p.stats[((size_t)g * 2) * H + c], rstd = p.stats[((size_t)g * 2 + 1) * H + c];
    const float w = p.weight[c], ms = p.mean_scale[c];
    const float shift = mean * ms, inv_n = 1.f / (float)n;
    float s_dn_o = 0.f;
    for (int r = beg; r < end; ++r) {
      const size_t k = (size_t)r * H + c;
      const float dz = p.y[k] > 0.f ? p.dy[k] : 0.f;
      const float o = p.x[k] - shift;
      dw = fmaf(dz, o * rstd, dw);
      db += dz;
      s_dn_o = fmaf(dz * w, o, s_dn_o);
    }
    const float dvar = -0.5f * rstd * rstd * rstd * s_dn_o;
    float s_do = 0.f;
    for (int r = beg; r < end; ++r) {
      const size_t k = (size_t)r * H + c;
      const float dz = p.y[k] > 0.f ? p.dy[k] : 0.f;
      const float o = p.x[k] - shift;
      s_do += dz * w * rstd + dvar * 2.f * o * inv_n;
    }
    dms = -mean * s_do;
    const float corr = ms * inv_n * s_do;
    for (int r = beg; r < end; ++r) {
      const size_t k = (size_t)r * H + c;
      const float dz = p.y[k] > 0.f ? p.dy[k] : 0.f;
      const float o = p.x[k] - shift;
      p.dx[k] = dz * w * rstd + dvar * 2.f * o * inv_n - corr;
    }
  }
  p.param_partial[((size_t)g * 3 + 0) * H + c] = dw;
  p.param_partial[((size_t)g * 3 + 1) * H + c] = db;
  p.param_partial[((size_t)g * 3 + 2) * H + c] = dms;
}

int graphnorm_relu_bwd(const GraphNormBwdArgs& a, cudaStream_t s) {
  if (a.n_graphs == 0) return CNS_OK;
  graphnorm_relu_bwd_kernel<<<(unsigned)a.n_graphs, 128, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// pool + heads backward (one CTA per graph)
// =============================================================================================
__global__ void __launch_bounds__(128)
pool_head_bwd_kernel(PoolHeadBwdArgs p) {
  __shared__ float dh[2][H];
  __shared__ float dxs[H];
  __shared__ float red[4];
  __shared__ float red2[4];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int beg = p.graph_ptr[g], end = p.graph_ptr[g + 1];
  // heads, second layers
  float dv = 0.f;
#pragma unroll
  for (int q = 0; q < 6; ++q) dv = fmaf(p.dvec ? p.dvec[(size_t)g * 6 + q] : 0.f, p.vec1_w[q * H + tid], dv);
  const float dn = (p.dnrm ? p.dnrm[g] : 0.f) * p.nrm1_w[tid];
  const float h0 = p.hid[((size_t)g * 2) * H + tid], h1 = p.hid[((size_t)g * 2 + 1) * H + tid];
  const float d0 = h0 > 0.f ? dv : 0.f, d1 = h1 > 0.f ? dn : 0.f;
  dh[0][tid] = d0; dh[1][tid] = d1;
  p.dhid[((size_t)g * 2) * H + tid] = d0;
  p.dhid[((size_t)g * 2 + 1) * H + tid] = d1;
  __syncthreads();
  // first layers: dxs[k] = sum_o dh0[o] W0v[o][k] + dh1[o] W0n[o][k]   (W in [out][in] layout)
  float acc = 0.f;
#pragma unroll 4
  for (int o = 0; o < H; ++o) acc = fmaf(dh[0][o], p.vec0_w[o * H + tid], fmaf(dh[1][o], p.nrm0_w[o * H + tid], acc));
  dxs[tid] = acc;
  __syncthreads();
  // softmax statistics of the gate over un-masked rows
  float wmax = -INFINITY;
  for (int r = beg + tid; r < end; r += 128)
    if (p.cluster_mask[r]) wmax = fmaxf(wmax, p.gate[r]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  if (lane == 0) red[wid] = wmax;
  __syncthreads();
  const float m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  // dg_r = dxs . x_r (one warp per row), den and S = sum g_r dg_r
  const float4 dx4 = reinterpret_cast<const float4*>(dxs)[lane];
  float den = 0.f, sgd = 0.f;
  for (int r = beg + wid; r < end; r += 4) {
    if (!p.cluster_mask[r]) continue;
    const float4 x = reinterpret_cast<const float4*>(p.x + (size_t)r * H)[lane];
    float d = x.x * dx4.x + x.y * dx4.y + x.z * dx4.z + x.w * dx4.w;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
    const float e = expf(p.gate[r] - m);
    den += e;
    sgd = fmaf(e, d, sgd);
    if (lane == 0) p.dgate[r] = d;        // temporarily dg_r
  }
  if (lane == 0) { red[wid] = den; red2[wid] = sgd; }
  __syncthreads();
  den = red[0] + red[1] + red[2] + red[3];
  sgd = red2[0] + red2[1] + red2[2] + red2[3];
  const float inv = den > 0.f ? 1.f / (den + 1e-16f) : 0.f;
  const float S = sgd * inv;
  const float pw = p.pool_w[tid], dxk = dxs[tid];
  for (int r = beg; r < end; ++r) {
    float out = 0.f;
    if (p.cluster_mask[r]) {
      const float gr = expf(p.gate[r] - m) * inv;
      const float dgate = gr * (p.dgate[r] - S);
      out = fmaf(gr, dxk, dgate * pw);
      __syncthreads();                    // every thread has read dg_r before it is overwritten
      if (tid == 0) p.dgate[r] = dgate;
    } else {
      __syncthreads();
      if (tid == 0) p.dgate[r] = 0.f;
    }
    p.dx[(size_t)r * H + tid] = out;
  }
}

int pool_head_bwd(const PoolHeadBwdArgs& a, cudaStream_t s) {
  if (a.n_graphs == 0) return CNS_OK;
  pool_head_bwd_kernel<<<(unsigned)a.n_graphs, 128, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// element-wise helpers
// =============================================================================================
__global__ void ew_kernel(EwArgs p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  switch (p.op) {
    case EW_RELU_BWD:        // out = (y > 0) ? a : 0
      p.out[i] = p.b[i] > 0.f ? p.a[i] : 0.f; break;
    case EW_RELU_BWD_ADD:    // out = ((y > 0) ? a : 0) ; out2 += out   (residual branch)
      { float v = p.b[i] > 0.f ? p.a[i] : 0.f; p.out[i] = v; p.out2[i] = (p.c ? p.c[i] : 0.f) + v; } break;
    case EW_ADD:
      p.out[i] = p.a[i] + p.b[i]; break;
    case EW_GRU_A: {         // a = dh', b = u, c = ht, d = h  -> out = dconv_k, out2 = du, out3 = dh (direct)
      const float dhn = p.a[i], u = p.b[i], ht = p.c[i], h = p.d ? p.d[i] : 0.f;
      p.out[i] = dhn * u * (1.f - ht * ht);
      p.out2[i] = dhn * (ht - h);
      p.out3[i] = dhn * (1.f - u);
    } break;
    default: break;
  }
}

// GRU part B: rows of 256 gate channels.  dhr (n,256 view: cols 128..255 of dXk), h, g=(r|u), du -> dconv_g (n,256),
// dh_acc += dhr * r
__global__ void gru_b_kernel(const float* __restrict__ dxk, const float* __restrict__ h, const float* __restrict__ g,
                             const float* __restrict__ du, float* __restrict__ dconvg, float* __restrict__ dh_acc,
                             int64_t n_rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * H) return;
  const int64_t r = i / H; const int c = (int)(i % H);
  const float dhr = dxk[r * 2 * H + H + c];
  const float hv = h ? h[i] : 0.f;
  const float rr = g[r * 2 * H + c], uu = g[r * 2 * H + H + c];
  dconvg[r * 2 * H + c] = dhr * hv * rr * (1.f - rr);
  dconvg[r * 2 * H + H + c] = du[i] * uu * (1.f - uu);
  dh_acc[i] += dhr * rr;
}

// out[r][c] = a[r*lda + c] + b[r*ldb + c]   (column-slice adds, 128 columns)
__global__ void add_slices_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                  float* __restrict__ out, int64_t n_rows, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * H) return;
  const int64_t r = i / H; const int c = (int)(i % H);
  float v = a[r * lda + c] + (b ? b[r * ldb + c] : 0.f);
  out[i] = accumulate ? out[i] + v : v;
}

int ew(const EwArgs& a, cudaStream_t s) {
  if (a.n == 0) return CNS_OK;
  ew_kernel<<<ceil_div(a.n, 256), 256, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}
int gru_b(const float* dxk, const float* h, const float* g, const float* du, float* dconvg, float* dh_acc,
          int64_t n_rows, cudaStream_t s) {
  if (n_rows == 0) return CNS_OK;
  gru_b_kernel<<<ceil_div(n_rows * H, 256), 256, 0, s>>>(dxk, h, g, du, dconvg, dh_acc, n_rows);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}
int add_slices(const float* a, int lda, const float* b, int ldb, float* out, int64_t n_rows, int accumulate,
               cudaStream_t s) {
  if (n_rows == 0) return CNS_OK;
  add_slices_kernel<<<ceil_div(n_rows * H, 256), 256, 0, s>>>(a, lda, b, ldb, out, n_rows, accumulate);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// PointEdgeConv weight-gradient unfolding: from G (2F x Fin) = dPQ^T X, Gp (2F x 2) = dPQ^T pos,
// gb (2F) = colsum(dPQ) to dW (F x (2 Fin + 2)) and db (F), columns [x_i | x_j - x_i | pos_j - pos_i]:
//   dW1 = G_P, dW2 = G_Q - G_P, dW3 = Gp_Q - Gp_P, db = gb_P
// =============================================================================================
__global__ void unfold_edgeconv_kernel(const float* __restrict__ G, const float* __restrict__ Gp,
                                       const float* __restrict__ gb, int f_in, int f_out, float* __restrict__ dW,
                                       float* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int ld = 2 * f_in + 2;
  if (idx >= f_out * (f_in + 3)) return;
  const int o = idx / (f_in + 3), k = idx % (f_in + 3);
  if (k < f_in) {
    const float gp = G[(size_t)o * f_in + k], gq = G[(size_t)(f_out + o) * f_in + k];
    dW[(size_t)o * ld + k] += gp;
    dW[(size_t)o * ld + f_in + k] += gq - gp;
  } else if (k < f_in + 2) {
    const int d = k - f_in;
    dW[(size_t)o * ld + 2 * f_in + d] += Gp[(f_out + o) * 2 + d] - Gp[o * 2 + d];
  } else {
    db[o] += gb[o];
  }
}

int unfold_edgeconv(const float* G, const float* Gp, const float* gb, int f_in, int f_out, float* dW, float* db,
                    cudaStream_t s) {
  const int total = f_out * (f_in + 3);
  unfold_edgeconv_kernel<<<ceil_div(total, 256), 256, 0, s>>>(G, Gp, gb, f_in, f_out, dW, db);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// =============================================================================================
// encoder backward helpers (per-edge tensors are materialised in the backward workspace)
// =============================================================================================
// X = relu(in_enc(feat[j])), HP = relu(pos_nn.0(dpos)), FE = feat[j] (E,2), DP = dpos (E,4)
__global__ void __launch_bounds__(128)
enc_gather_kernel(EncGatherArgs p) {
  __shared__ float fe[8][2];
  __shared__ float dp[8][4];
  const int64_t e0 = (int64_t)blockIdx.x * 8;
  const int k = threadIdx.x;
  if (k < 8) {
    const int64_t e = e0 + k;
    float f0 = 0.f, f1 = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (e < p.n_rows) {
      const int64_t j = p.ej ? p.ej[e] : p.centers[e];
      const float2 f = reinterpret_cast<const float2*>(p.feat)[j];
      f0 = f.x; f1 = f.y;
      if (p.ei) {
        const int64_t i = p.ei[e];
        const float2 ps = reinterpret_cast<const float2*>(p.pos_src)[j];
        const float2 pt = reinterpret_cast<const float2*>(p.pos_tar)[j];
        const float2 pc = reinterpret_cast<const float2*>(p.pos_clu)[i];
        d0 = pc.x - ps.x; d1 = pc.y - ps.y; d2 = pc.x - pt.x; d3 = pc.y - pt.y;
      }
      p.FE[e * 2] = f0; p.FE[e * 2 + 1] = f1;
      if (p.DP) { p.DP[e * 4] = d0; p.DP[e * 4 + 1] = d1; p.DP[e * 4 + 2] = d2; p.DP[e * 4 + 3] = d3; }
    }
    fe[k][0] = f0; fe[k][1] = f1; dp[k][0] = d0; dp[k][1] = d1; dp[k][2] = d2; dp[k][3] = d3;
  }
  __syncthreads();
  const float w0 = p.w_in[k * 2], w1 = p.w_in[k * 2 + 1], b = p.b_in[k];
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, qb = 0.f;
  if (p.HP) { q0 = p.w_p0[k * 4]; q1 = p.w_p0[k * 4 + 1]; q2 = p.w_p0[k * 4 + 2]; q3 = p.w_p0[k * 4 + 3]; qb = p.b_p0[k]; }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int64_t e = e0 + r;
    if (e >= p.n_rows) break;
    p.X[e * H + k] = fmaxf(fmaf(w1, fe[r][1], fmaf(w0, fe[r][0], b)), 0.f);
    if (p.HP) {
      float hp = qb;
      hp = fmaf(q0, dp[r][0], hp); hp = fmaf(q1, dp[r][1], hp); hp = fmaf(q2, dp[r][2], hp); hp = fmaf(q3, dp[r][3], hp);
      p.HP[e * H + k] = fmaxf(hp, 0.f);
    }
  }
}

int enc_gather(const EncGatherArgs& a, cudaStream_t s) {
  if (a.n_rows == 0) return CNS_OK;
  enc_gather_kernel<<<ceil_div(a.n_rows, 8), 128, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// T[e] += a_dst[ei[e]]
__global__ void add_rows_gather_kernel(float* __restrict__ t, const float* __restrict__ rows, const int64_t* __restrict__ ei,
                                       int64_t n_edges) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges * (H / 4)) return;
  const int64_t e = i / (H / 4); const int c4 = (int)(i % (H / 4));
  float4 v = reinterpret_cast<float4*>(t)[i];
  const float4 r = reinterpret_cast<const float4*>(rows + ei[e] * H)[c4];
  v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
  reinterpret_cast<float4*>(t)[i] = v;
}
int add_rows_gather(float* t, const float* rows, const int64_t* ei, int64_t n_edges, cudaStream_t s) {
  if (n_edges == 0) return CNS_OK;
  add_rows_gather_kernel<<<ceil_div(n_edges * (H / 4), 256), 256, 0, s>>>(t, rows, ei, n_edges);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// per-cluster segment softmax backward.  One warp per cluster (float4 lanes over 128 channels):
//   pass 1: m, s over the cluster's edges of L;  pass 2: alpha = exp(L - m)/(s + 1e-16),
//   dV = alpha * dout, dL = alpha * (V - out) * dout   with out = relu-output y (dout == 0 where y == 0)
// also accumulates nothing across clusters -> deterministic.
__global__ void __launch_bounds__(256)
enc_softmax_bwd_kernel(EncSoftmaxBwdArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= p.n_clusters) return;
  const int beg = p.clu_rowptr[c], end = p.clu_rowptr[c + 1];
  if (end <= beg) return;
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int e = beg; e < end; ++e) {
    const float4 l = reinterpret_cast<const float4*>(p.L + (size_t)e * H)[lane];
    m.x = fmaxf(m.x, l.x); m.y = fmaxf(m.y, l.y); m.z = fmaxf(m.z, l.z); m.w = fmaxf(m.w, l.w);
  }
  float4 sden = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = beg; e < end; ++e) {
    const float4 l = reinterpret_cast<const float4*>(p.L + (size_t)e * H)[lane];
    sden.x += expf(l.x - m.x); sden.y += expf(l.y - m.y); sden.z += expf(l.z - m.z); sden.w += expf(l.w - m.w);
  }
  const float4 y = reinterpret_cast<const float4*>(p.y + (size_t)c * H)[lane];
  float4 dy = reinterpret_cast<const float4*>(p.dy + (size_t)c * H)[lane];
  dy.x = y.x > 0.f ? dy.x : 0.f; dy.y = y.y > 0.f ? dy.y : 0.f; dy.z = y.z > 0.f ? dy.z : 0.f; dy.w = y.w > 0.f ? dy.w : 0.f;
  const float4 inv = make_float4(1.f / (sden.x + 1e-16f), 1.f / (sden.y + 1e-16f), 1.f / (sden.z + 1e-16f), 1.f / (sden.w + 1e-16f));
  for (int e = beg; e < end; ++e) {
    const float4 l = reinterpret_cast<const float4*>(p.L + (size_t)e * H)[lane];
    const float4 v = reinterpret_cast<const float4*>(p.V + (size_t)e * H)[lane];
    float4 a = make_float4(expf(l.x - m.x) * inv.x, expf(l.y - m.y) * inv.y, expf(l.z - m.z) * inv.z, expf(l.w - m.w) * inv.w);
    reinterpret_cast<float4*>(p.dV + (size_t)e * H)[lane] = make_float4(a.x * dy.x, a.y * dy.y, a.z * dy.z, a.w * dy.w);
    reinterpret_cast<float4*>(p.dL + (size_t)e * H)[lane] =
        make_float4(a.x * (v.x - y.x) * dy.x, a.y * (v.y - y.y) * dy.y, a.z * (v.z - y.z) * dy.z, a.w * (v.w - y.w) * dy.w);
  }
}
int enc_softmax_bwd(const EncSoftmaxBwdArgs& a, cudaStream_t s) {
  if (a.n_clusters == 0) return CNS_OK;
  enc_softmax_bwd_kernel<<<ceil_div(a.n_clusters, 8), 256, 0, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// out[c] (+)= sum_{e in segment c} rows[e]     (warp per cluster, fixed order)
__global__ void __launch_bounds__(256)
seg_sum_rows_kernel(const float* __restrict__ rows, const int* __restrict__ rowptr, int64_t n_clusters,
                    float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n_clusters) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = rowptr[c]; e < rowptr[c + 1]; ++e) {
    const float4 v = reinterpret_cast<const float4*>(rows + (size_t)e * H)[lane];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  float4* o = reinterpret_cast<float4*>(out + (size_t)c * H) + lane;
  if (accumulate) { float4 t = *o; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
  *o = s;
}
int seg_sum_rows(const float* rows, const int* rowptr, int64_t n_clusters, float* out, int accumulate, cudaStream_t s) {
  if (n_clusters == 0) return CNS_OK;
  seg_sum_rows_kernel<<<ceil_div(n_clusters, 8), 256, 0, s>>>(rows, rowptr, n_clusters, out, accumulate);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// out_enc backward prologue / epilogue on (C,*) rows:
//  mode 0: cat = [cur | tar - cur] from clu_feat (2,C,128);  dpre = dx * mask * (x_clu > 0)
//  mode 1: dy_cur = dcat[:, :128] - dcat[:, 128:], dy_tar = dcat[:, 128:]   -> dclu (2,C,128)
__global__ void out_enc_bwd_kernel(int mode, int64_t n_clusters, const float* __restrict__ clu_feat,
                                   const float* __restrict__ dx, const float* __restrict__ x_clu,
                                   const uint8_t* __restrict__ mask, float* __restrict__ cat, float* __restrict__ dpre,
                                   const float* __restrict__ dcat, float* __restrict__ dclu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_clusters * H) return;
  const int64_t r = i / H; const int c = (int)(i % H);
  if (mode == 0) {
    const float cur = clu_feat[i], tar = clu_feat[n_clusters * H + i];
    cat[r * 2 * H + c] = cur;
    cat[r * 2 * H + H + c] = tar - cur;
    dpre[i] = (mask[r] && x_clu[i] > 0.f) ? dx[i] : 0.f;
  } else {
    const float a = dcat[r * 2 * H + c], b = dcat[r * 2 * H + H + c];
    dclu[i] = a - b;
    dclu[n_clusters * H + i] = b;
  }
}
int out_enc_bwd(int mode, int64_t n_clusters, const float* clu_feat, const float* dx, const float* x_clu,
                const uint8_t* mask, float* cat, float* dpre, const float* dcat, float* dclu, cudaStream_t s) {
  if (n_clusters == 0) return CNS_OK;
  out_enc_bwd_kernel<<<ceil_div(n_clusters * H, 256), 256, 0, s>>>(mode, n_clusters, clu_feat, dx, x_clu, mask, cat,
                                                                 dpre, dcat, dclu);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// dst[k][o] = sign * src[o][k]  (generic transpose for backward operands)
__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float sign, float* __restrict__ dst) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx % cols;
  dst[(size_t)c * rows + r] = sign * src[idx];
}
int transpose(const float* src, int rows, int cols, float sign, float* dst, cudaStream_t s) {
  transpose_kernel<<<ceil_div(rows * cols, 256), 256, 0, s>>>(src, rows, cols, sign, dst);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns

using namespace cns;

extern "C" {

size_t cns_weight_grad_workspace_bytes(void) { return ((size_t)592 * 128 * 128 + (size_t)592 * 128) * sizeof(float); }

int cns_weight_grad(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, int64_t m, float* dw, float* dbias,
                    int accumulate, int tensor_cores, void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(dy && x && dw && workspace && m >= 0, "NULL pointer / bad size");
  CNS_REQUIRE(workspace_bytes >= cns_weight_grad_workspace_bytes(), "cns_weight_grad: workspace too small");
  DwGemmArgs d{};
  d.m = m; d.na = 128; d.nb = 128; d.a = dy; d.lda = (int)ld_dy; d.b0 = x; d.ldb0 = (int)ld_x; d.b_tiles0 = 1;
  d.out = dw; d.ld_out = 128; d.scale = 1.f; d.accumulate = accumulate;
  d.partial = (float*)workspace; d.max_slabs = 592; d.allow_split = tensor_cores;
  d.colsum_out = dbias; d.colsum_partial = (float*)workspace + (size_t)592 * 128 * 128;
  return dw_gemm(d, (cudaStream_t)stream);
}

}  // extern "C"
