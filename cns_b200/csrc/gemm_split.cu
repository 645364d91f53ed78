// fp32-accurate linear layer on the tensor cores, for the edge-level GEMMs of the backward pass
// (recomputed encoder chain + dX = dY W, cns/models/graph_vs.py:195-214 under torch autograd):
//   Y = gate(act(X Wt + bias + residual)),   X (m,128) fp32, Wt (128,128) fp32 row-major (k, n)
// Both operands are split into three bf16 slices (x = hi + mid + lo carries 24 mantissa bits, bf16 keeps fp32's
// exponent range, so no scaling is needed even for tiny gradients); the six slice products of weight >= 2^-18
// (hh, hm, mh, hl, lh, mm) accumulate in one fp32 TMEM tile: 48 tcgen05.mma (M128 x N128 x K16) per 128-row tile,
// error ~2^-24 like an fp32 FMA chain, at 6x fewer tensor FLOPs than the fp32 FFMA kernel has CUDA-core FLOPs.
// Persistent CTAs (one per SM) keep the pre-split weight image (96 KB, 1-D TMA bulk copies) in shared memory and
// walk the 128-row tiles with three roles: split the A rows into the three SWIZZLE_128B images / issue the MMAs /
// epilogue, the accumulator double-buffered in TMEM.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

namespace {

constexpr int GS_THREADS = 17 * 32;        // gemm_split: warps 0-7 epilogue, 8-15 operand split, 16 MMA issue
constexpr int DW_THREADS = 17 * 32;        // dw_split: warps 0-7 split A (+ epilogue), 8-15 split B, 16 MMA issue
constexpr uint32_t GS_SLICE = 128 * 128 * 2;                    // one bf16 slice of a 128 x 128 operand
constexpr int GS_STAGE_LD = 20;                                 // floats per staged 16-column row (16-byte aligned, conflict-free)
constexpr size_t GS_SMEM = 6 * GS_SLICE + 128 * 4 + 128 + 8 * 32 * GS_STAGE_LD * 4 + 1024;   // A, W slices, bias, barriers, epilogue tiles, slack
constexpr size_t GS_DW_SMEM = 6 * GS_SLICE + 16 * 128 * 4 + 128 + 1024;

__device__ __forceinline__ void split3(float v, unsigned short& h, unsigned short& m, unsigned short& l) {
  const __nv_bfloat16 bh = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(bh);                    // exact
  const __nv_bfloat16 bm = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(bm);                   // exact
  h = __bfloat16_as_ushort(bh); m = __bfloat16_as_ushort(bm); l = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
}

// the same for two neighbouring elements with packed conversions (cvt.rn.bf16x2.f32 instead of three F2F each):
// returns the packed (a | b << 16) words of the three slices
__device__ __forceinline__ void split3x2(float a, float b, uint32_t& h, uint32_t& m, uint32_t& l) {
  auto pack = [](float x, float y) {
    __nv_bfloat162 t = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&t);
  };
  h = pack(a, b);
  const float ra = a - __uint_as_float(h << 16), rb = b - __uint_as_float(h & 0xffff0000u);
  m = pack(ra, rb);
  l = pack(ra - __uint_as_float(m << 16), rb - __uint_as_float(m & 0xffff0000u));
}

// A thread's share of a 128-row block: NB x (row, 8 consecutive columns) chunks.  Loading and splitting are separate
// steps so the loads of the NEXT block are in flight (32 NB bytes per thread) while the tensor core still reads the
// current operand images - the loop is a pure stream and would otherwise expose the HBM latency once per tile.
template <int NB>
struct RowChunks { float4 lo[NB], hi[NB]; };

template <int NB>
__device__ __forceinline__ void load_rows(RowChunks<NB>& r, const float* __restrict__ src, int64_t ld, int64_t m0, int64_t m_end,
                                          int q0, int q_stride) {
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int q = q0 + j * q_stride;
    const int64_t m = m0 + (q >> 4);
    const int c = (q & 15) * 8;
    if (m < m_end) {
      r.lo[j] = *reinterpret_cast<const float4*>(src + m * ld + c);
      r.hi[j] = *reinterpret_cast<const float4*>(src + m * ld + c + 4);
    } else {
      r.lo[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      r.hi[j] = r.lo[j];
    }
  }
}

template <int NB>
__device__ __forceinline__ void split_store_rows(const RowChunks<NB>& r, int q0, int q_stride, unsigned char* dst, float* cs) {
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const int q = q0 + j * q_stride;
    const float v[8] = {r.lo[j].x, r.lo[j].y, r.lo[j].z, r.lo[j].w, r.hi[j].x, r.hi[j].y, r.hi[j].z, r.hi[j].w};
    if (cs) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cs[i] += v[i];
    }
    uint32_t h[4], md[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split3x2(v[2 * i], v[2 * i + 1], h[i], md[i], l[i]);
    const uint32_t off = sw128_off(128, q >> 4, (q & 15) * 8);
    *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst + GS_SLICE + off) = make_uint4(md[0], md[1], md[2], md[3]);
    *reinterpret_cast<uint4*>(dst + 2 * GS_SLICE + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
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
  // persistent: one CTA per SM walks the 128-row tiles.  Roles: warps 0-7 epilogue, warps 8-15 split the
  // next tile's rows into the three operand images, warp 16 issues the MMAs.  The fp32
  // accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the split and the MMAs of tile i+1.
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_s = smem;
  unsigned char* w_s = smem + 3 * GS_SLICE;
  float* bias_s = reinterpret_cast<float*>(w_s + 3 * GS_SLICE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 128);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 8);
  float* stage_s = reinterpret_cast<float*>(bars + 16);           // [8 warps][32 rows][GS_STAGE_LD]
  const uint32_t w_full = smem_u32(&bars[0]), a_full = smem_u32(&bars[1]);
  const uint32_t mma_done0 = smem_u32(&bars[2]), d_free0 = smem_u32(&bars[4]);      // [2], [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = (int)((p.m + 127) / 128);

  if (warp == 16) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      mbar_init(a_full, 256);
      mbar_init(mma_done0, 1); mbar_init(mma_done0 + 8, 1);
      mbar_init(d_free0, 256); mbar_init(d_free0 + 8, 256);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), 256);
    if (lane == 0) {
      mbar_expect_tx(w_full, 3 * GS_SLICE);
      for (int sl = 0; sl < 3; ++sl) bulk_g2s(smem_u32(w_s + sl * GS_SLICE), p.w_img + (size_t)sl * GS_SLICE, GS_SLICE, w_full);
    }
  } else if (tid < 128) {
    bias_s[tid] = p.bias ? p.bias[tid] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 16) {
    if (lane == 0) {
      // bf16 x bf16 -> fp32, M = 128, N = 128, both operands K-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_base = smem_u32(a_s), w_base = smem_u32(w_s);
      const int pa[6] = {1, 2, 0, 1, 0, 0}, pw[6] = {1, 0, 2, 0, 1, 0};     // (A slice, W slice): smallest products first
      mbar_wait(w_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it >= 2) mbar_wait(d_free0 + 8 * buf, (uint32_t)(((it - 2) >> 1) & 1));   // epilogue drained this accumulator
        mbar_wait(a_full, (uint32_t)(it & 1));
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(buf * 128);
        bool first = true;
        for (int t = 0; t < 6; ++t)
          for (int kc = 0; kc < 2; ++kc)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              umma_f16(d, sw128_desc(a_base + pa[t] * GS_SLICE + kc * (128 * 128) + ks * 32),
                       sw128_desc(w_base + pw[t] * GS_SLICE + kc * (128 * 128) + ks * 32), idesc, first ? 0u : 1u);
              first = false;
            }
        umma_commit(mma_done0 + 8 * buf);
      }
    }
  } else if (warp >= 8) {
    // ---- operand split: 8 consecutive k of one row per step -> one 16-byte store into each slice image ----
    const int ct = tid - 256;
    RowChunks<8> rows;
    if ((int)blockIdx.x < n_tiles) load_rows<8>(rows, p.x, 128, (int64_t)blockIdx.x * 128, p.m, ct, 256);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      if (it > 0) mbar_wait(mma_done0 + 8 * ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));   // previous tile's MMAs read A
      split_store_rows<8>(rows, ct, 256, a_s, nullptr);
      fence_proxy_async();
      mbar_arrive(a_full);
      const int next = tile + gridDim.x;
      if (next < n_tiles) load_rows<8>(rows, p.x, 128, (int64_t)next * 128, p.m, ct, 256);      // in flight during the MMAs
    }
  } else {
    // ---- epilogue (warps 0-7: TMEM lane quarter = warp & 3, column half = warp >> 2).  A TMEM read hands every lane
    // one ROW; going to global memory like that costs one L1 sector cycle per 16 bytes (32 rows per instruction) and
    // was measured to double the kernel time.  Each warp therefore turns 32 x 16 sub-chunks through a private
    // shared-memory tile and moves residual / gate / output with 4 lanes per row (64 contiguous bytes).  All loads of
    // a sub-chunk are issued before its first store: y may alias the residual, so the compiler would otherwise
    // serialise load -> store -> load.
    float* stage = stage_s + warp * (32 * GS_STAGE_LD);
    const int wq = warp & 3, ch = warp >> 2;
    const int er = lane >> 2, ec = (lane & 3) * 4;          // this lane's row (mod 8) and column within the sub-chunk
    const bool has_res = p.residual != nullptr, has_gate = p.relu_gate != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int64_t mw = (int64_t)tile * 128 + wq * 32;      // first row of this warp
      mbar_wait(mma_done0 + 8 * buf, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int n0 = ch * 64 + c * 32;
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * 128 + n0), acc);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int nb = n0 + 16 * hf;
          float4 res[4], gate[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t m = mw + 8 * j + er;
            const bool ok = m < p.m;
            res[j] = (ok && has_res) ? *reinterpret_cast<const float4*>(p.residual + m * 128 + nb + ec) : make_float4(0.f, 0.f, 0.f, 0.f);
            gate[j] = (ok && has_gate) ? *reinterpret_cast<const float4*>(p.relu_gate + m * 128 + nb + ec) : make_float4(1.f, 1.f, 1.f, 1.f);
          }
          if (hf == 0) tmem_ld_wait();
          __syncwarp();                                      // the previous sub-chunk has been read out of `stage`
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(stage + lane * GS_STAGE_LD + 4 * q) =
                make_float4(acc[16 * hf + 4 * q], acc[16 * hf + 4 * q + 1], acc[16 * hf + 4 * q + 2], acc[16 * hf + 4 * q + 3]);
          __syncwarp();
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + nb + ec);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t m = mw + 8 * j + er;
            const float4 a4 = *reinterpret_cast<const float4*>(stage + (8 * j + er) * GS_STAGE_LD + ec);
            float4 v = make_float4(a4.x + b4.x + res[j].x, a4.y + b4.y + res[j].y, a4.z + b4.z + res[j].z, a4.w + b4.w + res[j].w);
            if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            v.x = gate[j].x > 0.f ? v.x : 0.f; v.y = gate[j].y > 0.f ? v.y : 0.f;
            v.z = gate[j].z > 0.f ? v.z : 0.f; v.w = gate[j].w > 0.f ? v.w : 0.f;
            if (m < p.m) *reinterpret_cast<float4*>(p.y + m * 128 + nb + ec) = v;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(d_free0 + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 16) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// dW = A^T B over m rows (weight gradients of the same layers): D[i][j] = sum_m A[m][i] B[m][j], i, j < 128.
// The reduction index is the row index, so both operands are MN-major: a 128-row block of A (or B), split into
// three bf16 slices and stored row by row with the 128-byte swizzle, is exactly the byte image the K-major tiles
// above use - only the descriptors differ (MN-major, LBO = the second 64-column block, SBO = the next 8 rows).
// Persistent CTAs each own a contiguous range of 128-row blocks and accumulate all of them into ONE TMEM tile
// (48 MMAs per block); the per-CTA partial products go through the fixed-order reduce_partials_kernel.
// ---------------------------------------------------------------------------------------------------------------
struct DwSplitArgs {
  int64_t m; const float* a; int lda; const float* b; int ldb;
  float* partial; float* colsum_partial; int blocks_per_cta;
};

__device__ __forceinline__ uint64_t mn128_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((128 * 128) >> 4) << 16;             // LBO: next block of 64 MN elements (128 rows x 128 B further)
  d |= (uint64_t)(1024 >> 4) << 32;                    // SBO: next group of 8 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(DW_THREADS, 1)
dw_split_kernel(const __grid_constant__ DwSplitArgs p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_s = smem;                           // [3 slices][2 column blocks][128 rows][128 B]
  unsigned char* b_s = smem + 3 * GS_SLICE;
  float* cs_s = reinterpret_cast<float*>(b_s + 3 * GS_SLICE);          // [16][128] column-sum staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(cs_s + 16 * 128);
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 4);
  const uint32_t full = smem_u32(&bars[0]), mma_done = smem_u32(&bars[1]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_blocks = (int)((p.m + 127) / 128);
  const int blk0 = blockIdx.x * p.blocks_per_cta;
  const int blk1 = min(n_blocks, blk0 + p.blocks_per_cta);

  if (warp == 16) {
    if (lane == 0) {
      mbar_init(full, 512);
      mbar_init(mma_done, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 16) {
    if (lane == 0) {
      // bf16 x bf16 -> fp32, M = 128, N = 128, A and B MN-major (bits 15 / 16)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t a_base = smem_u32(a_s), b_base = smem_u32(b_s);
      const int pa[6] = {1, 2, 0, 1, 0, 0}, pb[6] = {1, 0, 2, 0, 1, 0};
      bool first = true;
      for (int blk = blk0, it = 0; blk < blk1; ++blk, ++it) {
        mbar_wait(full, (uint32_t)(it & 1));
        tc_fence_after();
        for (int t = 0; t < 6; ++t)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {             // 16 rows (two 8-row groups) per MMA
            umma_f16(tmem_base, mn128_desc(a_base + pa[t] * GS_SLICE + ks * 2048),
                     mn128_desc(b_base + pb[t] * GS_SLICE + ks * 2048), idesc, first ? 0u : 1u);
            first = false;
          }
        umma_commit(mma_done);
      }
    }
  } else {
    // warps 0-7 split the A rows, warps 8-15 the B rows of the block
    const bool is_b = warp >= 8;
    const int ct = tid & 255;
    const float* __restrict__ src = is_b ? p.b : p.a;
    const int ld = is_b ? p.ldb : p.lda;
    unsigned char* dst = is_b ? b_s : a_s;
    float cs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = 0.f;
    RowChunks<8> rows;
    if (blk0 < blk1) load_rows<8>(rows, src, ld, (int64_t)blk0 * 128, p.m, ct, 256);
    for (int blk = blk0, it = 0; blk < blk1; ++blk, ++it) {
      if (it > 0) mbar_wait(mma_done, (uint32_t)((it - 1) & 1));
      split_store_rows<8>(rows, ct, 256, dst, is_b ? nullptr : cs);
      fence_proxy_async();
      mbar_arrive(full);
      if (blk + 1 < blk1) load_rows<8>(rows, src, ld, (int64_t)(blk + 1) * 128, p.m, ct, 256);   // in flight during the MMAs
    }
    // column sums: thread ct owns columns (ct & 15) * 8 .. + 7 for the rows (ct >> 4) + 16 j
    if (!is_b && p.colsum_partial) {
#pragma unroll
      for (int i = 0; i < 8; ++i) cs_s[(ct >> 4) * 128 + (ct & 15) * 8 + i] = cs[i];
    }
    named_bar(1, 512);
    if (tid < 128 && p.colsum_partial) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) t += cs_s[r * 128 + tid];
      p.colsum_partial[(size_t)blockIdx.x * 128 + tid] = t;
    }
    // epilogue (warps 0-7): the CTA's partial product, thread = row i, 64 columns per warp group
    const int n_it = blk1 - blk0;
    if (warp < 8) {
    const int wq = warp & 3, half = warp >> 2;
    float* prow = p.partial + ((size_t)blockIdx.x * 128 + wq * 32 + lane) * 128 + half * 64;
    if (n_it > 0) {
      mbar_wait(mma_done, (uint32_t)((n_it - 1) & 1));
      tc_fence_after();
    }
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      float acc[32];
      if (n_it > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(half * 64 + sub * 32), acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) acc[q] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 32; q += 4)
        *reinterpret_cast<float4*>(prow + sub * 32 + q) = make_float4(acc[q], acc[q + 1], acc[q + 2], acc[q + 3]);
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 16) tmem_dealloc(tmem_base, 128);
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
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(gemm_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_SMEM));
  }
  SplitArgs p;
  p.m = a.m; p.x = a.x1; p.w_img = w_img; p.bias = a.bias; p.residual = a.residual; p.relu_gate = a.relu_gate;
  p.relu = a.relu; p.y = a.y;
  static int n_sm = 0;      // one process per GPU model: every device of a box has the same SM count
  if (!n_sm) {
    int dev = 0;
    CNS_CUDA_CHECK(cudaGetDevice(&dev));
    CNS_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles = ceil_div(a.m, 128);
  gemm_split_kernel<<<tiles < n_sm ? tiles : n_sm, GS_THREADS, GS_SMEM, s>>>(p);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// stage 1 of dw_gemm for a single 128 x 128 output tile on the tensor cores; returns the number of partials
int dw_split(int64_t m, const float* a, int lda, const float* b, int ldb, float* partial, float* colsum_partial, int max_slabs,
             int* n_slabs, cudaStream_t s) {
  static unsigned long long configured = 0;
  static int n_sm = 0;
  if (first_use_on_device(configured)) {
    int dev = 0;
    CNS_CUDA_CHECK(cudaGetDevice(&dev));
    CNS_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    CNS_CUDA_CHECK(cudaFuncSetAttribute(dw_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_DW_SMEM));
  }
  const int n_blocks = ceil_div(m, 128);
  int ctas = n_blocks < n_sm ? n_blocks : n_sm;
  if (ctas > max_slabs) ctas = max_slabs;
  DwSplitArgs p;
  p.m = m; p.a = a; p.lda = lda; p.b = b; p.ldb = ldb; p.partial = partial; p.colsum_partial = colsum_partial;
  p.blocks_per_cta = (n_blocks + ctas - 1) / ctas;
  ctas = (n_blocks + p.blocks_per_cta - 1) / p.blocks_per_cta;
  dw_split_kernel<<<ctas, DW_THREADS, GS_DW_SMEM, s>>>(p);
  CNS_LAUNCH_CHECK();
  *n_slabs = ctas;
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
