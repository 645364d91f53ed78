// fp32-accurate tcgen05 GEMM for the batched cluster-level MLPs of the fp32 mode (PointEdgeConv [P|Q] projections,
// PERConv fc layers, out_enc; cns/models/graph_vs.py:16-31,108,145,229):
//   Y = act([X1 | X2] W^T + pos Wp + bias + residual) * row_mask
// Both operands are split into two 16-bit slices (bf16: hi + lo = 16 mantissa bits at fp32's exponent range; f16: 22 bits at
// f16's range - template parameter F16S, tc_common.cuh split_pair) and the three
// products hi*hi, lo*hi, hi*lo accumulate in one fp32 TMEM tile - the same scheme as the encoder edge kernel of this
// mode (encoder_split.cu); measured against the fp64 oracle the whole forward stays at ~1e-5 (budget 1e-4).
//
// Grid (ceil(M / 128), N / 256 or 1): a CTA owns 128 rows (TMEM lanes) x up to 256 output columns, i.e. at most 256
// TMEM columns and ~100 KB of shared memory, so two CTAs share an SM - the 150 row tiles of the benchmark batch fit
// in one wave of 148 x 2.  K is walked in halves of 128: the A half is split once into (hi, lo) K-major SWIZZLE_128B
// images (64 KB); the weight image streams in 32 KB units (one slice of a 128 x 128 block, 1-D TMA bulk copy) through
// a single buffer.  Warps 0-7: operand split + epilogue; warp 8: TMA + MMA issue.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

namespace {

constexpr int G2_THREADS = 288;
constexpr uint32_t G2_UNIT = 128 * 128 * 2;                       // one slice of a 128 x 128 block
constexpr size_t G2_SMEM = 3 * G2_UNIT + 3 * 256 * 4 + 256 + 1024;  // A hi, A lo, W unit, bias / wp rows, barriers, slack

// image: [n / 128][k / 128][slice][128 rows (n) x 128 k, K-major SWIZZLE_128B]
template <bool F16S>
__global__ void g2_pack_kernel(const float* __restrict__ wt, int k_dim, int n_out, unsigned char* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k_dim * n_out) return;
  const int k = idx / n_out, n = idx % n_out;
  const float v = wt[idx];
  unsigned short hi, lo;
  split_scalar<F16S>(v, hi, lo);
  const size_t block = ((size_t)(n / 128) * (k_dim / 128) + (k / 128)) * 2;
  const uint32_t off = sw128_off(128, n % 128, k % 128);
  *reinterpret_cast<unsigned short*>(img + block * G2_UNIT + off) = hi;
  *reinterpret_cast<unsigned short*>(img + (block + 1) * G2_UNIT + off) = lo;
}

template <bool F16S>
__global__ void __launch_bounds__(G2_THREADS, 2)
gemm_tc_split_kernel(const __grid_constant__ GemmTcArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + G2_UNIT;
  unsigned char* w_s = smem + 2 * G2_UNIT;
  float* vec_s = reinterpret_cast<float*>(w_s + G2_UNIT);                 // bias[256], wp0[256], wp1[256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec_s + 3 * 256);          // [0] W unit landed, [1] MMAs of the unit done, [2] half done
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.k1 + p.k2, KH = K / 128;
  const int col0 = (int)blockIdx.y * 256;                                 // first output column of this CTA
  const int n_cols = min(256, p.n_out - col0);
  const int NC = n_cols / 128;
  const uint32_t tmem_cols = n_cols <= 128 ? 128u : 256u;
  const int64_t m0 = (int64_t)blockIdx.x * 128;

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bars[i]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), tmem_cols);
  } else {
    for (int i = tid; i < n_cols; i += 256) {
      vec_s[i] = p.bias ? p.bias[col0 + i] : 0.f;
      vec_s[256 + i] = p.pos ? p.wp[col0 + i] : 0.f;
      vec_s[512 + i] = p.pos ? p.wp[p.n_out + col0 + i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  for (int h = 0; h < KH; ++h) {
    if (warp < 8) {
      // ---- A half h: fp32 rows -> (hi, lo) SWIZZLE_128B images.  The previous half's MMAs have drained. ----
      if (h > 0) { mbar_wait(smem_u32(&bars[2]), (uint32_t)((h - 1) & 1)); tc_fence_after(); }
      for (int q = tid; q < 128 * 16; q += 256) {
        const int row = q >> 4, k = (q & 15) * 8;
        const int kk = h * 128 + k;
        const int64_t m = m0 + row;
        float4 lo4 = make_float4(0.f, 0.f, 0.f, 0.f), hi4 = lo4;
        if (m < p.m) {
          const float* src = (kk < p.k1) ? p.x1 + m * p.k1 + kk : p.x2 + m * p.k2 + (kk - p.k1);
          lo4 = *reinterpret_cast<const float4*>(src);
          hi4 = *reinterpret_cast<const float4*>(src + 4);
        }
        uint32_t hh[4], ll[4];
        split_pair<F16S>(lo4.x, lo4.y, hh[0], ll[0]); split_pair<F16S>(lo4.z, lo4.w, hh[1], ll[1]);
        split_pair<F16S>(hi4.x, hi4.y, hh[2], ll[2]); split_pair<F16S>(hi4.z, hi4.w, hh[3], ll[3]);
        const uint32_t off = sw128_off(128, row, k);
        *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 8 && lane == 0) {
      const uint32_t idesc = (1u << 4) | split_fmt_bits<F16S>() | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), wb = smem_u32(w_s);
      for (int c = 0; c < NC; ++c) {
        const uint32_t d = tmem_base + (uint32_t)(c * 128);
        for (int sl = 0; sl < 2; ++sl) {
          const int u = (h * NC + c) * 2 + sl;                       // running unit index of this CTA
          if (u > 0) { mbar_wait(smem_u32(&bars[1]), (uint32_t)((u - 1) & 1)); tc_fence_after(); }   // buffer drained
          const size_t block = ((size_t)((col0 >> 7) + c) * (p.k_img ? p.k_img / 128 : KH) + h) * 2 + sl;
          mbar_expect_tx(smem_u32(&bars[0]), G2_UNIT);
          bulk_g2s(wb, p.w_img + block * G2_UNIT, G2_UNIT, smem_u32(&bars[0]));
          mbar_wait(smem_u32(&bars[0]), (uint32_t)(u & 1));
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t o = (uint32_t)((kk >> 2) * (128 * 128) + (kk & 3) * 32);
            const uint32_t first = (h == 0 && sl == 0 && kk == 0) ? 0u : 1u;
            umma_f16(d, sw128_desc(ah + o), sw128_desc(wb + o), idesc, first);                  // hi x hi | hi x lo
            if (sl == 0) umma_f16(d, sw128_desc(al + o), sw128_desc(wb + o), idesc, 1u);         // lo x hi
          }
          umma_commit(smem_u32(&bars[1]));
        }
      }
      umma_commit(smem_u32(&bars[2]));                               // every MMA of this half (A images free / results ready)
    }
  }

  if (warp < 8) {
    // ---- epilogue: thread = row (TMEM lane), 64 columns of every 128-column chunk -------------------
    mbar_wait(smem_u32(&bars[2]), (uint32_t)((KH - 1) & 1));
    tc_fence_after();
    const int wq = warp & 3, half = warp >> 2;
    const int64_t m = m0 + wq * 32 + lane;
    const bool live = m < p.m;
    float2 ps = make_float2(0.f, 0.f);
    if (live && p.pos) ps = reinterpret_cast<const float2*>(p.pos)[m];
    const bool keep = live && (!p.row_mask || p.row_mask[m]);
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int n0 = c * 128 + half * 64 + sub * 32;               // column inside this CTA's group
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)n0, acc);
        tmem_ld_wait();
        if (live) {
          float* yrow = p.y + m * p.n_out + col0 + n0;
          const float* rrow = p.residual ? p.residual + m * p.n_out + col0 + n0 : nullptr;
#pragma unroll
          for (int q = 0; q < 32; q += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(vec_s + n0 + q);
            const float4 w0 = *reinterpret_cast<const float4*>(vec_s + 256 + n0 + q);
            const float4 w1 = *reinterpret_cast<const float4*>(vec_s + 512 + n0 + q);
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

}  // namespace

int gemm_tc_split(const GemmTcArgs& a, bool f16_slices, cudaStream_t s) {
  if (a.m == 0) return CNS_OK;
  CNS_REQUIRE(a.n_out % 128 == 0 && a.n_out <= 512, "gemm_tc_split: n_out must be 128..512, multiple of 128");
  CNS_REQUIRE((a.k1 == 128 || a.k1 == 256) && (a.k2 == 0 || a.k2 == 128) && a.k1 + a.k2 <= 256, "gemm_tc_split: K must be 128 or 256");
  CNS_REQUIRE(a.w_img != nullptr, "gemm_tc_split: weight image missing");
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM));
    CNS_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM));
  }
  const dim3 grid((unsigned)ceil_div(a.m, 128), (unsigned)ceil_div(a.n_out, 256));
  if (f16_slices) gemm_tc_split_kernel<true><<<grid, G2_THREADS, G2_SMEM, s>>>(a);
  else gemm_tc_split_kernel<false><<<grid, G2_THREADS, G2_SMEM, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int gemm_tc_split_pack(const float* wt, int k_dim, int n_out, unsigned char* img, bool f16_slices, cudaStream_t s) {
  const int total = k_dim * n_out;
  if (f16_slices) g2_pack_kernel<true><<<ceil_div(total, 256), 256, 0, s>>>(wt, k_dim, n_out, img);
  else g2_pack_kernel<false><<<ceil_div(total, 256), 256, 0, s>>>(wt, k_dim, n_out, img);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns
