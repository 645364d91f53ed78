// tcgen05 GEMM for the batched cluster-level MLPs (16-bit operands, fp32 accumulate in TMEM):
//   Y = act([X1 | X2] W^T + pos Wp + bias + residual) * row_mask
// used for the PointEdgeConv [P|Q] projections, the PERConv fc layers and out_enc
// (cns/models/graph_vs.py:16-31,108,145,229) when the precision mode is bf16 / f16.
//
// One CTA per 128-row tile of X (rows = clusters, TMEM lanes).  X is converted fp32 -> 16-bit into a
// K-major SWIZZLE_128B A tile once; the weight image streams through two 64 KB shared-memory buffers
// in chunks of 128 output columns (1-D TMA bulk copies); each chunk is K/16 tcgen05.mma
// (M128 x N128 x K16) into its own 128 TMEM columns, so the epilogue of chunk c overlaps the MMAs of
// chunk c+1.  Warps 0-7: conversion + epilogue (warp%4 = TMEM lane quarter, warp/4 = column half);
// warp 8: TMA + MMA issue.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

constexpr int GT_THREADS = 288;

struct GemmTcSmem {
  static constexpr int A_BYTES = 128 * 256 * 2;          // K <= 256
  static constexpr int W_BYTES = 128 * 256 * 2;          // one 128-column chunk
  static constexpr int VEC_FLOATS = 3 * 512;             // bias, wp[0], wp[1]
  static constexpr int TOTAL = A_BYTES + 2 * W_BYTES + VEC_FLOATS * 4 + 256 + 1024;
};

// weight image: wt (k, n) fp32 row-major  ->  [n/128][k/64][128 rows][128 B] swizzled 16-bit (B operand)
template <bool BF16>
__global__ void gemm_tc_pack_kernel(const float* __restrict__ wt, int k_dim, int n_out, unsigned char* __restrict__ img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k_dim * n_out) return;
  int k = idx / n_out, n = idx % n_out;
  float v = wt[idx];
  unsigned short bits = BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
  size_t chunk_bytes = (size_t)128 * k_dim * 2;
  *reinterpret_cast<unsigned short*>(img + (size_t)(n / 128) * chunk_bytes + sw128_off(128, n % 128, k)) = bits;
}

template <bool BF16>
__global__ void __launch_bounds__(GT_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ GemmTcArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // buffers are sized by the actual K (a K = 128 product needs half the bytes: two CTAs then share an SM,
  // which matters because ceil(M / 128) tiles rarely divide the SM count)
  const uint32_t a_bytes = 128u * (uint32_t)(p.k1 + p.k2) * 2u, w_bytes = a_bytes;
  unsigned char* a_s = smem;
  unsigned char* w_s = smem + a_bytes;
  float* vec_s = reinterpret_cast<float*>(w_s + 2 * w_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec_s + GemmTcSmem::VEC_FLOATS);   // [0..1] wfull, [2..5] done
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.k1 + p.k2;
  const int n_chunks = p.n_out / 128;
  const uint32_t chunk_bytes = (uint32_t)(128 * K * 2);                                   // bytes consumed per chunk
  const size_t chunk_stride = (size_t)128 * (p.k_img ? p.k_img : K) * 2;                 // bytes between chunks
  const uint32_t tmem_cols = p.n_out <= 128 ? 128u : (p.n_out <= 256 ? 256u : 512u);
  const int64_t m0 = (int64_t)blockIdx.x * 128;

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), tmem_cols);
    if (lane == 0) {
      for (int c = 0; c < 2 && c < n_chunks; ++c) {
        mbar_expect_tx(smem_u32(&bars[c]), chunk_bytes);
        bulk_g2s(smem_u32(w_s + c * w_bytes), p.w_img + (size_t)c * chunk_stride, chunk_bytes, smem_u32(&bars[c]));
      }
    }
  } else {
    // ---- A tile: fp32 rows -> 16-bit SWIZZLE_128B, 16 bytes (8 elements) per store ----------------
    const int chunks_per_row = K / 8;
    if (p.centers) {
      // centre-feature producer (K = 128): this thread owns 8 channels (k .. k+7) of rows tid / 16 + 16 j, j < 8: the
      // in_enc weights are loaded once, the eight centre indices and then the eight coordinate pairs in one batch each
      // (a load per item behind the previous item's shared-memory store was a chain of 16 L2 round trips)
      const int k = (tid & 15) * 8, r0 = tid >> 4;
      float2 cw[8]; float cb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { cw[i] = __ldg(reinterpret_cast<const float2*>(p.w_in) + k + i); cb[i] = __ldg(p.b_in + k + i); }
      int64_t ctr[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int64_t m = m0 + r0 + 16 * j; ctr[j] = m < p.m ? p.centers[m] : -1; }
      float2 f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = ctr[j] >= 0 ? reinterpret_cast<const float2*>(p.ctr_xy)[ctr[j]] : make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = r0 + 16 * j;
        const int64_t m = m0 + row;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = ctr[j] >= 0 ? fmaxf(fmaf(cw[i].y, f[j].y, fmaf(cw[i].x, f[j].x, cb[i])), 0.f) : 0.f;
        *reinterpret_cast<uint4*>(a_s + sw128_off(128, row, k)) =
            make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7]));
        if (k == 0 && ctr[j] >= 0) {
          reinterpret_cast<float2*>(p.pos_out)[m] = reinterpret_cast<const float2*>(p.ctr_pos)[ctr[j]];
          p.clu_graph[m] = p.node_graph ? (int)p.node_graph[ctr[j]] : 0;
        }
      }
    } else
    for (int q = tid; q < 128 * chunks_per_row; q += 256) {
      const int row = q / chunks_per_row, kc8 = q % chunks_per_row;
      const int k = kc8 * 8;
      const int64_t m = m0 + row;
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
      if (m < p.m) {
        const float* src = (k < p.k1) ? p.x1 + m * p.k1 + k : p.x2 + m * p.k2 + (k - p.k1);
        lo = *reinterpret_cast<const float4*>(src);
        hi = *reinterpret_cast<const float4*>(src + 4);
      }
      uint4 v = make_uint4(pack2<BF16>(lo.x, lo.y), pack2<BF16>(lo.z, lo.w), pack2<BF16>(hi.x, hi.y), pack2<BF16>(hi.z, hi.w));
      *reinterpret_cast<uint4*>(a_s + sw128_off(128, row, k)) = v;
    }
    for (int i = tid; i < p.n_out; i += 256) {
      vec_s[i] = p.bias ? p.bias[i] : 0.f;
      vec_s[512 + i] = p.pos ? p.wp[i] : 0.f;
      vec_s[1024 + i] = p.pos ? p.wp[p.n_out + i] : 0.f;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 8) {
    if (lane == 0) {
      const uint32_t fmt = BF16 ? 1u : 0u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_base = smem_u32(a_s);
      for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        mbar_wait(smem_u32(&bars[b]), (uint32_t)((c >> 1) & 1));
        tc_fence_after();
        const uint32_t w_base = smem_u32(w_s + b * w_bytes);
        const uint32_t d = tmem_base + (uint32_t)(c * 128);
        for (int kc = 0; kc < K / 64; ++kc)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(d, sw128_desc(a_base + kc * (128 * 128) + ks * 32), sw128_desc(w_base + kc * (128 * 128) + ks * 32),
                     idesc, (kc | ks) ? 1u : 0u);
        umma_commit(smem_u32(&bars[2 + c]));
        if (c + 2 < n_chunks) {                      // refill this buffer once its MMAs have drained
          mbar_wait(smem_u32(&bars[2 + c]), 0);
          mbar_expect_tx(smem_u32(&bars[b]), chunk_bytes);
          bulk_g2s(smem_u32(w_s + b * w_bytes), p.w_img + (size_t)(c + 2) * chunk_stride, chunk_bytes,
                   smem_u32(&bars[b]));
        }
      }
    }
  } else {
    // ---- epilogue: thread = row (TMEM lane), 64 columns of every 128-column chunk -------------------
    const int wq = warp & 3, half = warp >> 2;
    const int64_t m = m0 + wq * 32 + lane;
    const bool live = m < p.m;
    float2 ps = make_float2(0.f, 0.f);
    if (live && p.pos) ps = reinterpret_cast<const float2*>(p.pos)[m];
    const bool keep = live && (!p.row_mask || p.row_mask[m]);
    for (int c = 0; c < n_chunks; ++c) {
      mbar_wait(smem_u32(&bars[2 + c]), 0);
      tc_fence_after();
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int n0 = c * 128 + half * 64 + sub * 32;
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)n0, acc);
        tmem_ld_wait();
        if (live) {
          float* yrow = p.y + m * p.n_out + n0;
          const float* rrow = p.residual ? p.residual + m * p.n_out + n0 : nullptr;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(vec_s + n0 + q);
            const float4 w0 = *reinterpret_cast<const float4*>(vec_s + 512 + n0 + q);
            const float4 w1 = *reinterpret_cast<const float4*>(vec_s + 1024 + n0 + q);
            float4 v;
            v.x = fmaf(ps.y, w1.x, fmaf(ps.x, w0.x, acc[q] + b4.x));
            v.y = fmaf(ps.y, w1.y, fmaf(ps.x, w0.y, acc[q + 1] + b4.y));
            v.z = fmaf(ps.y, w1.z, fmaf(ps.x, w0.z, acc[q + 2] + b4.z));
            v.w = fmaf(ps.y, w1.w, fmaf(ps.x, w0.w, acc[q + 3] + b4.w));
            if (rrow) {
              const float4 r = *reinterpret_cast<const float4*>(rrow + q);
              v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (!keep) v = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(yrow + q) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tmem_base, tmem_cols);
}

int gemm_tc(const GemmTcArgs& a, bool bf16, cudaStream_t s) {
  if (a.m == 0) return CNS_OK;
  CNS_REQUIRE(a.n_out % 128 == 0 && a.n_out <= 512, "gemm_tc: n_out must be 128..512, multiple of 128");
  CNS_REQUIRE((a.k1 == 128 || a.k1 == 256) && (a.k2 == 0 || a.k2 == 128) && a.k1 + a.k2 <= 256, "gemm_tc: K must be 128 or 256");
  CNS_REQUIRE(a.w_img != nullptr, "gemm_tc: weight image missing");
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmTcSmem::TOTAL));
    CNS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmTcSmem::TOTAL));
  }
  const int grid = ceil_div(a.m, 128);
  const size_t smem = (size_t)3 * 128 * (a.k1 + a.k2) * 2 + GemmTcSmem::VEC_FLOATS * 4 + 256 + 1024;
  if (bf16) gemm_tc_kernel<true><<<grid, GT_THREADS, smem, s>>>(a);
  else gemm_tc_kernel<false><<<grid, GT_THREADS, smem, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int gemm_tc_pack(const float* wt, int k_dim, int n_out, unsigned char* img_bf16, unsigned char* img_f16, cudaStream_t s) {
  const int total = k_dim * n_out;
  gemm_tc_pack_kernel<true><<<ceil_div(total, 256), 256, 0, s>>>(wt, k_dim, n_out, img_bf16);
  CNS_LAUNCH_CHECK();
  gemm_tc_pack_kernel<false><<<ceil_div(total, 256), 256, 0, s>>>(wt, k_dim, n_out, img_f16);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns
