// fp32-accurate linear layer on the tensor cores, for the edge-level GEMMs of the backward pass
// (recomputed encoder chain + dX = dY W, cns/models/graph_vs.py:195-214 under torch autograd):
//   Y = gate(act(X Wt + bias + residual)),   X (m,128) fp32, Wt (128,128) fp32 row-major (k, n)
// Both operands are split into three bf16 slices (x = hi + mid + lo carries 24 mantissa bits, bf16 keeps fp32's
// exponent range, so no scaling is needed even for tiny gradients); the six slice products of weight >= 2^-18
// (hh, hm, mh, hl, lh, mm) accumulate in one fp32 TMEM tile: 48 tcgen05.mma (M128 x N128 x K16) per 128-row tile,
// error ~2^-24 like an fp32 FMA chain, at 6x fewer tensor FLOPs than the fp32 FFMA kernel has CUDA-core FLOPs.
// One CTA per 128-row tile: warps 0-7 split the A rows into the three SWIZZLE_128B images and run the epilogue,
// warp 8 loads the pre-split weight image (96 KB, 1-D TMA bulk copies) and issues the MMAs.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

namespace {

constexpr int GS_THREADS = 288;
constexpr uint32_t GS_SLICE = 128 * 128 * 2;                    // one bf16 slice of a 128 x 128 operand
constexpr size_t GS_SMEM = 6 * GS_SLICE + 128 * 4 + 64 + 1024;  // A slices, W slices, bias, barriers, alignment slack

__device__ __forceinline__ void split3(float v, unsigned short& h, unsigned short& m, unsigned short& l) {
  const __nv_bfloat16 bh = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(bh);                    // exact
  const __nv_bfloat16 bm = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(bm);                   // exact
  h = __bfloat16_as_ushort(bh); m = __bfloat16_as_ushort(bm); l = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
}

__global__ void split_pack_kernel(const float* __restrict__ wt, unsigned char* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 128 * 128) return;
  const int k = idx / 128, n = idx % 128;
  unsigned short h, m, l;
  split3(wt[idx], h, m, l);
  const uint32_t off = sw128_off(128, n, k);
  *reinterpret_cast<unsigned short*>(img + off) = h;
  *reinterpret_cast<unsigned short*>(img + GS_SLICE + off) = m;
  *reinterpret_cast<unsigned short*>(img + 2 * GS_SLICE + off) = l;
}

struct SplitArgs {
  int64_t m;
  const float* x; const unsigned char* w_img; const float* bias; const float* residual; const float* relu_gate;
  int relu;
  float* y;
};

__global__ void __launch_bounds__(GS_THREADS, 1)
gemm_split_kernel(const __grid_constant__ SplitArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_s = smem;
  unsigned char* w_s = smem + 3 * GS_SLICE;
  float* bias_s = reinterpret_cast<float*>(w_s + 3 * GS_SLICE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 128);   // [0] weights landed, [1] MMAs done
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * 128;

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars[0]), 1);
      mbar_init(smem_u32(&bars[1]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), 128);
    if (lane == 0) {
      mbar_expect_tx(smem_u32(&bars[0]), 3 * GS_SLICE);
      for (int sl = 0; sl < 3; ++sl)
        bulk_g2s(smem_u32(w_s + sl * GS_SLICE), p.w_img + (size_t)sl * GS_SLICE, GS_SLICE, smem_u32(&bars[0]));
    }
  } else {
    // A tile: 8 consecutive k of one row per step -> one 16-byte store into each of the three slice images
    for (int q = tid; q < 128 * 16; q += 256) {
      const int row = q >> 4, k = (q & 15) * 8;
      const int64_t m = m0 + row;
      float v[8];
      if (m < p.m) {
        const float4 lo = *reinterpret_cast<const float4*>(p.x + m * 128 + k);
        const float4 hi = *reinterpret_cast<const float4*>(p.x + m * 128 + k + 4);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      unsigned short h[8], md[8], l[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) split3(v[i], h[i], md[i], l[i]);
      auto pk4 = [](const unsigned short* s) {
        return make_uint4((uint32_t)s[0] | ((uint32_t)s[1] << 16), (uint32_t)s[2] | ((uint32_t)s[3] << 16),
                          (uint32_t)s[4] | ((uint32_t)s[5] << 16), (uint32_t)s[6] | ((uint32_t)s[7] << 16));
      };
      const uint32_t off = sw128_off(128, row, k);
      *reinterpret_cast<uint4*>(a_s + off) = pk4(h);
      *reinterpret_cast<uint4*>(a_s + GS_SLICE + off) = pk4(md);
      *reinterpret_cast<uint4*>(a_s + 2 * GS_SLICE + off) = pk4(l);
    }
    if (tid < 128) bias_s[tid] = p.bias ? p.bias[tid] : 0.f;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 8) {
    if (lane == 0) {
      // bf16 x bf16 -> fp32, M = 128, N = 128, both operands K-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      mbar_wait(smem_u32(&bars[0]), 0);
      tc_fence_after();
      // (A slice, W slice): smallest products first
      const int pa[6] = {1, 2, 0, 1, 0, 0}, pw[6] = {1, 0, 2, 0, 1, 0};
      bool first = true;
      for (int t = 0; t < 6; ++t)
        for (int kc = 0; kc < 2; ++kc)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_f16(tmem_base, sw128_desc(a_base + pa[t] * GS_SLICE + kc * (128 * 128) + ks * 32),
                     sw128_desc(w_base + pw[t] * GS_SLICE + kc * (128 * 128) + ks * 32), idesc, first ? 0u : 1u);
            first = false;
          }
      umma_commit(smem_u32(&bars[1]));
    }
  } else {
    // epilogue: thread = row (TMEM lane), 64 of the 128 columns
    const int wq = warp & 3, half = warp >> 2;
    const int64_t m = m0 + wq * 32 + lane;
    const bool live = m < p.m;
    mbar_wait(smem_u32(&bars[1]), 0);
    tc_fence_after();
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int n0 = half * 64 + sub * 32;
      float acc[32];
      tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)n0, acc);
      tmem_ld_wait();
      if (live) {
        float* yrow = p.y + m * 128 + n0;
        const float* rrow = p.residual ? p.residual + m * 128 + n0 : nullptr;
        const float* grow = p.relu_gate ? p.relu_gate + m * 128 + n0 : nullptr;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n0 + q);
          float4 v = make_float4(acc[q] + b4.x, acc[q + 1] + b4.y, acc[q + 2] + b4.z, acc[q + 3] + b4.w);
          if (rrow) {
            const float4 r = *reinterpret_cast<const float4*>(rrow + q);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
          }
          if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          if (grow) {
            const float4 g = *reinterpret_cast<const float4*>(grow + q);
            v.x = g.x > 0.f ? v.x : 0.f; v.y = g.y > 0.f ? v.y : 0.f; v.z = g.z > 0.f ? v.z : 0.f; v.w = g.w > 0.f ? v.w : 0.f;
          }
          *reinterpret_cast<float4*>(yrow + q) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tmem_base, 128);
}

}  // namespace

int gemm_split_pack(const float* wt, unsigned char* img, cudaStream_t s) {
  split_pack_kernel<<<64, 256, 0, s>>>(wt, img);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// same contract as `linear` for k1 = 128, k2 = 0, n_out = 128, no pos term; w_img from gemm_split_pack
int gemm_split(const LinearArgs& a, const unsigned char* w_img, cudaStream_t s) {
  if (a.m == 0) return CNS_OK;
  CNS_REQUIRE(a.k1 == 128 && a.k2 == 0 && a.n_out == 128 && a.pos == nullptr && w_img != nullptr,
              "gemm_split: 128 x 128 layers without a position term only");
  static bool configured = false;
  if (!configured) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(gemm_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_SMEM));
    configured = true;
  }
  SplitArgs p;
  p.m = a.m; p.x = a.x1; p.w_img = w_img; p.bias = a.bias; p.residual = a.residual; p.relu_gate = a.relu_gate;
  p.relu = a.relu; p.y = a.y;
  gemm_split_kernel<<<ceil_div(a.m, 128), GS_THREADS, GS_SMEM, s>>>(p);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns

using namespace cns;

extern "C" {

size_t cns_linear_split_workspace_bytes(void) { return GEMM_SPLIT_IMG_BYTES; }

int cns_linear_split(const float* x, int64_t m, const float* wt, const float* bias, const float* residual,
                     const float* relu_gate, int relu, float* y, void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(m >= 0 && x && wt && y && workspace, "NULL pointer / bad size");
  CNS_REQUIRE(workspace_bytes >= GEMM_SPLIT_IMG_BYTES, "cns_linear_split: workspace too small");
  int rc = gemm_split_pack(wt, (unsigned char*)workspace, (cudaStream_t)stream);
  if (rc != CNS_OK) return rc;
  LinearArgs a{};
  a.m = m; a.k1 = 128; a.n_out = 128; a.x1 = x; a.bias = bias; a.residual = residual; a.relu = relu; a.relu_gate = relu_gate;
  a.y = y;
  return gemm_split(a, (const unsigned char*)workspace, (cudaStream_t)stream);
}

}  // extern "C"
