// fp32-accurate encoder edge kernel on tcgen05 (CNS_PREC_FP32): the keypoint->cluster attention chain of
// cns/models/graph_vs.py:195-214 with BOTH operands of every 128x128 GEMM split into two 16-bit slices
// (F16S = true: f16, 22 mantissa bits at f16's range - the default; F16S = false: bf16, v = hi + lo carries 16 mantissa
//  bits at fp32's exponent range; tc_common.cuh split_pair) and the three
// slice products hi*hi, hi*lo, lo*hi accumulated in one fp32 TMEM tile.  Measured against the fp64 oracle the
// cluster features come out at ~1e-5 with bf16 slices (the dropped lo*lo term is 2^-18 relative), an order of magnitude inside
// the 1e-4 budget of the fp32 mode, at 3 tensor-core passes instead of one FFMA pass per weight.
//
// Same algebra, tile walk and online segment softmax as encoder_tc.cu (see the notes there):
//   round A :  U = x (-(A0 lin_src))^T + hp (A0 pos_nn.1)^T      V = x lin^T + hp pos_nn.1^T    (K = 256 each)
//   epi   A :  u = relu(U + a'[cluster])          -> (hi, lo) -> smem, MN-major ([channel][edge])
//   round C :  L = u attn_nn.1^T                  final : online segment softmax of L weighting (V + b_p1)
// x = relu(in_enc(feat)), hp = relu(pos_nn.0(dpos)) are fp32 FFMA results split on the way into the operand tiles.
//
// Resources per CTA (one per SM): ten weight slices of 32 KB - hi slices of all five matrices and the lo slice
// of attn_nn.1 live in tensor memory (A operand of TS-mode MMAs: used by 2 of the 3 products, no shared-memory
// fetch), the other four lo slices stay in shared memory (128 KB).  TMEM: 128 accumulator columns (2 slots x
// {U/L, V} x 32 edges) + 6 x 64 weight columns.  Activation tiles: 2 slots x {x_hi, x_lo, hp_hi, hp_lo} x 8 KB.
//
// That budget leaves room for only two MMA slots of N = 32, i.e. 8 warps if a slot were one 32-edge tile - too few to
// hide the latencies of the SIMT phases.  So a slot's 32 MMA columns are TWO independent 16-edge sub-tiles, each
// walked by its own 4 warps through its own contiguous chunk of the edge list (own running softmax state): 16 warps
// per SM, the per-thread work between two MMA round trips halves, the MMA work per edge stays the same.
#include <cuda_bf16.h>
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

namespace {

constexpr int SP_TN = 32;                          // MMA N: edges per slot tile
constexpr int SP_SUB = 2;                          // independent sub-tiles (edge streams) per slot
constexpr int SP_TS = SP_TN / SP_SUB;              // 16 edges per sub-tile
constexpr int SP_SLOTS = 2;                        // MMA slots per CTA
constexpr int SP_W_BYTES = 128 * 128 * 2;          // one 16-bit slice of a weight matrix
constexpr int SP_NW = 5;
constexpr int SP_TMEM_SLICES = 6;                  // hi0..hi4, lo4
constexpr int SP_SMEM_SLICES = 4;                  // lo0..lo3
constexpr int SP_ACT_TILE = SP_TN * 256;           // one [32 x 128] 16-bit operand tile
constexpr int SP_ACT_BYTES = SP_SLOTS * 4 * SP_ACT_TILE;
constexpr int SP_ACC_COLS = SP_SLOTS * 2 * SP_TN;  // 128
constexpr size_t SP_SMEM_BYTES = SP_SMEM_SLICES * SP_W_BYTES + SP_ACT_BYTES + 256 + SP_SLOTS * 2 * SP_TN * sizeof(int) + 4096 + 1024;
static_assert(SP_TMEM_SLICES * SP_W_BYTES <= SP_SMEM_SLICES * SP_W_BYTES + SP_ACT_BYTES, "staging area for the TMEM-bound slices");
static_assert(SP_ACC_COLS + SP_TMEM_SLICES * 64 <= 512, "tensor memory budget");

// image order in global memory: [hi0 hi1 hi2 hi3 hi4 lo4 | lo0 lo1 lo2 lo3]
__host__ __device__ __forceinline__ int sp_image_slot(int w, int lo) { return lo ? (w == 4 ? 5 : 6 + w) : w; }

template <bool F16S>
__global__ void sp_pack_weights_kernel(const float* __restrict__ chain, unsigned char* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= SP_NW * 128 * 128) return;
  const int w = idx / 16384, rem = idx % 16384;
  const int k = rem / 128, o = rem % 128;          // chain[w][k][o] = W_w[o][k]
  float v = chain[idx];
  if (w == 4) v *= 1.4426950408889634f;            // logits in log2 units (softmax with ex2)
  unsigned short hi, lo;
  split_scalar<F16S>(v, hi, lo);
  const uint32_t off = sw128_off(128, o, k);
  *reinterpret_cast<unsigned short*>(img + (size_t)sp_image_slot(w, 0) * SP_W_BYTES + off) = hi;
  *reinterpret_cast<unsigned short*>(img + (size_t)sp_image_slot(w, 1) * SP_W_BYTES + off) = lo;
}

struct SpSmallW {          // kernel parameter space: per channel k {in_enc.w[k][0..1], in_enc.b[k], pos_nn.0.b[k]}, pos_nn.0.w[k][0..3]
  float4 w[128][2];
};

struct SpArgs {
  EncEdgeArgs e;
  const unsigned char* w_img;
  int n_tiles;
  int tiles_per_chunk;
};

// F16S: slice format of both operands (tc_common.cuh, split_pair)
template <bool F16S>
__global__ void __launch_bounds__(128 * SP_SUB * SP_SLOTS, 1)
enc_edge_split_kernel(const __grid_constant__ SpArgs args, const __grid_constant__ SpSmallW sw) {
  constexpr int TN = SP_TN, SLOTS = SP_SLOTS, TS = SP_TS;
  constexpr int U_ROW = TN * 2;                      // bytes per channel row of an MN-major u tile
  constexpr int NPRE = 4;
  constexpr uint32_t W_COL0 = SP_ACC_COLS;           // first weight column in tensor memory
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* w_s = smem;                                           // lo0..lo3 (after the staging pass)
  unsigned char* act_s = smem + SP_SMEM_SLICES * SP_W_BYTES;
  unsigned char* tail = act_s + SP_ACT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);                  // [0] weights landed, [1+s] MMAs of slot s done
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tail + 128);
  int* ci_s = reinterpret_cast<int*>(tail + 256);                      // [SLOTS][SUB][2][TS] cluster id per edge, double-buffered
  float4* sw_s = reinterpret_cast<float4*>(tail + 256 + SLOTS * 2 * TN * sizeof(int));   // first-layer weights [128][2]

  const EncEdgeArgs& p = args.e;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars[0]), 1);
      for (int s = 0; s < SLOTS; ++s) mbar_init(smem_u32(&bars[1 + s]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), 512);
    if (lane == 0) {
      // staging pass: the six TMEM-bound slices land in [w_s, w_s + 192 KB) (weights + activation area)
      mbar_expect_tx(smem_u32(&bars[0]), SP_TMEM_SLICES * SP_W_BYTES);
      for (int j = 0; j < SP_TMEM_SLICES; ++j)
        bulk_g2s(smem_u32(w_s + j * SP_W_BYTES), args.w_img + (size_t)j * SP_W_BYTES, SP_W_BYTES, smem_u32(&bars[0]));
    }
  }
  if (tid < 256) sw_s[tid] = sw.w[tid >> 1][tid & 1];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);
  {
    // row o of a slice = TMEM lane o, k and k+1 share a 32-bit column.  A warp reaches the lane quarter warp % 4:
    // warp group g = warp / 4 moves slices g and g + 4.
    mbar_wait(smem_u32(&bars[0]), 0);
    const int o = tid & 127;
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + W_COL0;
#pragma unroll 1
    for (int j = (warp >> 2); j < SP_TMEM_SLICES; j += 4) {
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {
        const uint4 lo = *reinterpret_cast<const uint4*>(w_s + j * SP_W_BYTES + sw128_off(128, o, g * 16));
        const uint4 hi = *reinterpret_cast<const uint4*>(w_s + j * SP_W_BYTES + sw128_off(128, o, g * 16 + 8));
        float v[8];
        uint32_t* u = reinterpret_cast<uint32_t*>(v);
        u[0] = lo.x; u[1] = lo.y; u[2] = lo.z; u[3] = lo.w; u[4] = hi.x; u[5] = hi.y; u[6] = hi.z; u[7] = hi.w;
        tmem_st8(t_row + (uint32_t)(j * 64 + g * 8), v);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  fence_proxy_async();          // generic-proxy reads of the staging area are ordered before the TMA writes that follow
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), SP_SMEM_SLICES * SP_W_BYTES);
    for (int j = 0; j < SP_SMEM_SLICES; ++j)
      bulk_g2s(smem_u32(w_s + j * SP_W_BYTES), args.w_img + (size_t)(SP_TMEM_SLICES + j) * SP_W_BYTES, SP_W_BYTES,
               smem_u32(&bars[0]));
  }

  // instruction descriptor: D = f32 (bit 4), A / B = bf16 (1 at bits 7 / 10), N >> 3 at bit 17, M >> 4 at bit 24;
  // bit 16 = B is MN-major
  const uint32_t idesc = (1u << 4) | split_fmt_bits<F16S>() | ((uint32_t)(TN >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_mn = idesc | (1u << 16);

  const int s = warp >> 3;                       // MMA slot
  const int sub = (warp >> 2) & 1;               // sub-tile (edge stream) of the slot
  const int t = tid & 127;                       // thread in sub-slot == output channel == TMEM lane
  const int wq = warp & 3;                       // TMEM lane quarter
  unsigned char* const xh = act_s + (s * 4 + 0) * SP_ACT_TILE;      // x hi  (round C: u hi)
  unsigned char* const xl = act_s + (s * 4 + 1) * SP_ACT_TILE;      // x lo  (round C: u lo)
  unsigned char* const ph = act_s + (s * 4 + 2) * SP_ACT_TILE;      // hp hi
  unsigned char* const pl = act_s + (s * 4 + 3) * SP_ACT_TILE;      // hp lo
  int* const ci_base = ci_s + (s * SP_SUB + sub) * 2 * TS;
  const uint32_t done = smem_u32(&bars[1 + s]);
  const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((s * 2 + 0) * TN + sub * TS);
  const uint32_t t1 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((s * 2 + 1) * TN + sub * TS);
  const float b_p1 = p.b_p1[t];
  uint32_t done_parity = 0;
  constexpr float LN2 = 0.6931471805599453f;

  const uint64_t hi_k = sw128_desc(0) & 0xffffffff00000000ull, hi_mn = mn_major_desc(0, U_ROW) & 0xffffffff00000000ull;
  const uint32_t lo_mn = (uint32_t)mn_major_desc(0, U_ROW);
  const uint32_t w_lo = smem_u32(w_s) >> 4;
  const uint32_t xh_lo = smem_u32(xh) >> 4, xl_lo = smem_u32(xl) >> 4, ph_lo = smem_u32(ph) >> 4, pl_lo = smem_u32(pl) >> 4;
  // one weight matrix times one activation operand (hi and lo tiles): W_hi B_hi + W_hi B_lo + W_lo B_hi
  auto gemm3_kmajor = [&](uint32_t d, int w, uint32_t bh_lo, uint32_t bl_lo, uint32_t acc_first) {
    const uint32_t a_hi = tmem_base + W_COL0 + (uint32_t)(w * 64);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t boff = (uint32_t)(((kk >> 2) * (TN * 128) + (kk & 3) * 32) >> 4);
      const uint64_t bh = hi_k | (uint64_t)(bh_lo + boff), bl = hi_k | (uint64_t)(bl_lo + boff);
      umma_f16_ts(d, a_hi + (uint32_t)(kk * 8), bh, idesc, kk ? 1u : acc_first);
      umma_f16_ts(d, a_hi + (uint32_t)(kk * 8), bl, idesc, 1u);
      const uint64_t ad = hi_k | (uint64_t)(w_lo + (uint32_t)((w * SP_W_BYTES + (kk >> 2) * (128 * 128) + (kk & 3) * 32) >> 4));
      umma_f16(d, ad, bh, idesc, 1u);                   // lo slices of w = 0..3 are read from shared memory
    }
  };
  auto gemm3_mnmajor = [&](uint32_t d, uint32_t bh_lo, uint32_t bl_lo) {       // attn_nn.1: both slices in tensor memory
    const uint32_t a_hi = tmem_base + W_COL0 + 4u * 64u, a_lo = tmem_base + W_COL0 + 5u * 64u;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t boff = (uint32_t)((kk * (16 * U_ROW)) >> 4);                  // 16 k-rows per MMA
      const uint64_t bh = hi_mn | (uint64_t)(lo_mn + bh_lo + boff), bl = hi_mn | (uint64_t)(lo_mn + bl_lo + boff);
      umma_f16_ts(d, a_hi + (uint32_t)(kk * 8), bh, idesc_mn, kk ? 1u : 0u);
      umma_f16_ts(d, a_hi + (uint32_t)(kk * 8), bl, idesc_mn, 1u);
      umma_f16_ts(d, a_lo + (uint32_t)(kk * 8), bh, idesc_mn, 1u);
    }
  };
  auto publish_and_issue = [&](int round) {
    tc_fence_before();
    fence_proxy_async();
    named_bar(1 + s, 128 * SP_SUB);
    if (wq == 0 && sub == 0) {
      if (elect_one()) {
        tc_fence_after();
        const uint32_t d0 = tmem_base + (uint32_t)((s * 2 + 0) * TN);
        const uint32_t d1 = tmem_base + (uint32_t)((s * 2 + 1) * TN);
        if (round == 0) {
          // everything that reads the x tiles first, commit after U: the epilogue may then overwrite the x tiles with u
          // while V's hp half (hp tiles only) still runs; the round-C commit covers it
          gemm3_kmajor(d1, 2, xh_lo, xl_lo, 0);   // V  = x  lin^T
          gemm3_kmajor(d0, 1, xh_lo, xl_lo, 0);   // U  = x  (-(A0 lin_src))^T
          gemm3_kmajor(d0, 3, ph_lo, pl_lo, 1);   //    + hp (A0 pos_nn.1)^T
          umma_commit(done);
          gemm3_kmajor(d1, 0, ph_lo, pl_lo, 1);   // V += hp pos_nn.1^T
        } else {
          gemm3_mnmajor(d0, xh_lo, xl_lo);        // L = u attn_nn.1^T   (u lives in the x tiles, [channel][edge])
          umma_commit(done);
        }
      }
      __syncwarp();
    }
  };
  if (wq == 0 && sub == 0) mbar_wait(smem_u32(&bars[0]), 1);     // the shared-memory lo slices landed (second phase of the barrier)

  const int pe = lane & (TS - 1);                // edge of this thread in the producer
  const int kp = wq * 2 + (lane >> 4);           // channel block [kp*16, kp*16 + 16)
  const int row = sub * TS + pe;                 // row of that edge in the slot's operand tiles
  const int n_chunks_side = ((int)gridDim.x * SLOTS * SP_SUB) / 2;
  const int chunk_q = ((int)blockIdx.x * SLOTS + s) * SP_SUB + sub;
  const int side = chunk_q / n_chunks_side;
  const int chunk = chunk_q - side * n_chunks_side;
  const int tpc = args.tiles_per_chunk;
  const int t_begin = chunk * tpc, t_end = min(t_begin + tpc, args.n_tiles);
  int idx_j = 0, idx_c = -1;
  float2 c_f = make_float2(0.f, 0.f), c_ps = c_f, c_pt = c_f, c_pc = c_f;
  bool c_live = false;
  float2 n_f = c_f, n_ps = c_f, n_pt = c_f, n_pc = c_f;
  int n_cid = -1;
  auto fetch_index = [&](int g) {
    idx_c = -1; idx_j = 0;
    if (g < t_end) {
      const int64_t e = (int64_t)g * TS + pe;
      if (e < p.n_e01) { idx_j = (int)p.ej[e]; idx_c = (int)p.ei[e]; }
    }
  };
  auto fetch_coords = [&]() {
    n_cid = idx_c;
    if (n_cid >= 0) {
      const float* __restrict__ fsrc = side ? p.x_tar : p.x_cur;
      const float* __restrict__ psrc = side ? p.pos_tar : p.pos_cur;
      n_f = reinterpret_cast<const float2*>(fsrc)[idx_j];
      n_ps = reinterpret_cast<const float2*>(psrc)[idx_j];
      n_pt = reinterpret_cast<const float2*>(p.pos_tar)[idx_j];
      n_pc = reinterpret_cast<const float2*>(p.pos_clu)[n_cid];
    }
  };
  auto advance_coords = [&]() { c_f = n_f; c_ps = n_ps; c_pt = n_pt; c_pc = n_pc; c_live = n_cid >= 0; };
  fetch_index(t_begin);
  if (kp == 0) ci_base[pe] = idx_c;
  fetch_coords();
  advance_coords();
  fetch_index(t_begin + 1);
  fetch_coords();
  fetch_index(t_begin + 2);
  named_bar(1 + s, 128 * SP_SUB);

  float run_m = 0.f, run_s = 0.f, run_a = 0.f;
  int run_seg = -1;
  float* const out_side = p.clu_feat + (size_t)side * p.n_clusters * H;
  float* const part_side = p.partial + (size_t)side * p.n_slots * 3 * H;
  bool run_whole, chunk_whole_end;
  {
    const int64_t ce0 = (int64_t)t_begin * TS, ce1 = min((int64_t)max(t_end, t_begin) * TS, p.n_e01);
    run_whole = (ce0 == 0) || (ce0 < p.n_e01 && p.ei[ce0 - 1] != p.ei[ce0]);
    chunk_whole_end = (ce1 >= p.n_e01) || (ce1 > 0 && p.ei[ce1 - 1] != p.ei[ce1]);
  }
  auto flush_run = [&](bool whole) {
    const float ab = fmaf(b_p1, run_s, run_a);
    if (whole) {
      out_side[(size_t)run_seg * H + t] = fmaxf(ab / (run_s + 1e-16f), 0.f);
    } else {
      float* slot = part_side + ((size_t)run_seg + chunk) * 3 * H;
      slot[t] = run_m * LN2; slot[H + t] = run_s; slot[2 * H + t] = ab;
    }
  };
  // u tiles, MN-major: row = channel t (U_ROW bytes of edges), 16-byte chunk index XOR-swizzled by the row
  const uint32_t u_off = (uint32_t)((t >> 3) * (8 * U_ROW) + (t & 7) * U_ROW);
  const uint32_t u_xor = (uint32_t)(((t & 7) * U_ROW) >> 7) & (uint32_t)(U_ROW / 16 - 1);

  uint32_t seg_mask, seg_next;
  float aq[NPRE], aq_next[NPRE];
  auto seg_and_prefetch = [&](const int* ci, uint32_t& sm, float (&a)[NPRE]) {
    const int cprev = (pe == 0) ? -2 : ci[pe - 1];
    sm = __ballot_sync(0xffffffffu, ci[pe] != cprev) & ((1u << TS) - 1u);
    uint32_t m0 = sm;
#pragma unroll
    for (int i = 0; i < NPRE; ++i) {
      a[i] = 0.f;
      if (m0) {
        const int q = __ffs(m0) - 1;
        m0 &= m0 - 1;
        const int c = ci[q];
        if (c >= 0) a[i] = p.a_dst[(size_t)c * H + t];
      }
    }
  };
  seg_and_prefetch(ci_base, seg_mask, aq);

  // both sub-tiles of a slot share its barriers and MMAs: every stream runs the same number of tiles; tiles past the
  // end of a stream are padding (cluster -1 everywhere: zero operands, a run that is never flushed)
  for (int it = 0; it < tpc; ++it) {
    const int g = t_begin + it;
    const int* const ci = ci_base + (it & 1) * TS;

    // ---- producer: x -> (xh, xl), hp -> (ph, pl): fp32 first layers, split on the way out -------------------
    {
      const bool live = c_live;
      const float f0 = live ? c_f.x : 0.f, f1 = live ? c_f.y : 0.f;
      const float d0 = live ? c_pc.x - c_ps.x : 0.f, d1 = live ? c_pc.y - c_ps.y : 0.f;   // pos_i - pos_j
      const float d2 = live ? c_pc.x - c_pt.x : 0.f, d3 = live ? c_pc.y - c_pt.y : 0.f;
#pragma unroll 1
      for (int gq = 0; gq < TS / 8; ++gq) {
        const int k0 = kp * TS + gq * 8;
        uint32_t xo[4], xt[4], ho[4], ht[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float xv[2], hv[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float4 wa = sw_s[(k0 + q * 2 + h2) * 2], wb = sw_s[(k0 + q * 2 + h2) * 2 + 1];
            const float x = fmaf(wa.y, f1, fmaf(wa.x, f0, wa.z));
            float hp = fmaf(wb.x, d0, wa.w);
            hp = fmaf(wb.y, d1, hp);
            hp = fmaf(wb.z, d2, hp);
            hp = fmaf(wb.w, d3, hp);
            xv[h2] = fmaxf(x, 0.f);
            hv[h2] = fmaxf(hp, 0.f);
          }
          split_pair<F16S>(xv[0], xv[1], xo[q], xt[q]);
          split_pair<F16S>(hv[0], hv[1], ho[q], ht[q]);
        }
        const uint32_t off = sw128_off(TN, row, k0);
        *reinterpret_cast<uint4*>(xh + off) = make_uint4(xo[0], xo[1], xo[2], xo[3]);
        *reinterpret_cast<uint4*>(xl + off) = make_uint4(xt[0], xt[1], xt[2], xt[3]);
        *reinterpret_cast<uint4*>(ph + off) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
        *reinterpret_cast<uint4*>(pl + off) = make_uint4(ht[0], ht[1], ht[2], ht[3]);
      }
    }
    publish_and_issue(0);
    if (kp == 0) ci_base[((it + 1) & 1) * TS + pe] = n_cid;

    // ---- epilogue: u = relu(U + a'[cluster]) -> (u hi, u lo), MN-major rows of this thread's channel ---------
    mbar_wait(done, done_parity); done_parity ^= 1;
    tc_fence_after();
    advance_coords();
    fetch_coords();
    fetch_index(t_begin + it + 3);
    {
      float adb = 0.f;
      int nseg = 0;
#pragma unroll 1
      for (int cg = 0; cg < TS / 8; ++cg) {
        float acc[8];
        tmem_ld8(t0 + cg * 8, acc);
        tmem_ld_wait();
        const uint32_t m8 = (seg_mask >> (cg * 8)) & 0xffu;
        uint32_t pk[4], pt[4];
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          float v[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            if (m8 & (1u << (q + h2))) {
              if (nseg < NPRE) {
                adb = aq[0];
#pragma unroll
                for (int i = 0; i + 1 < NPRE; ++i) aq[i] = aq[i + 1];
              } else {
                const int c = ci[cg * 8 + q + h2];
                adb = (c >= 0) ? p.a_dst[(size_t)c * H + t] : 0.f;
              }
              ++nseg;
            }
            v[h2] = fmaxf(acc[q + h2] + adb, 0.f);
          }
          split_pair<F16S>(v[0], v[1], pk[q / 2], pt[q / 2]);
        }
        const uint32_t off = u_off + (((uint32_t)(sub * (TS / 8) + cg) ^ u_xor) << 4);
        *reinterpret_cast<uint4*>(xh + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(xl + off) = make_uint4(pt[0], pt[1], pt[2], pt[3]);
      }
    }
    publish_and_issue(1);
    seg_and_prefetch(ci_base + ((it + 1) & 1) * TS, seg_next, aq_next);

    // ---- final: per-channel online segment softmax of L weighting V (see encoder_tc.cu) ---------------------
    mbar_wait(done, done_parity); done_parity ^= 1;
    tc_fence_after();
    {
      auto update = [&](float l, float v) {
        const float d = l - run_m;
        const float ex = ex2_approx(-fabsf(d));
        const bool up = d > 0.f;
        const float sc = up ? ex : 1.f, pr = up ? 1.f : ex;
        run_s = fmaf(run_s, sc, pr);
        run_a = fmaf(run_a, sc, pr * v);
        run_m = fmaxf(run_m, l);
      };
      const uint32_t first_bit = (ci[0] != run_seg) ? 1u : 0u;
#pragma unroll 1
      for (int cg = 0; cg < TS / 8; ++cg) {
        float lg[8], vv[8];
        tmem_ld8(t0 + cg * 8, lg);
        tmem_ld8(t1 + cg * 8, vv);
        tmem_ld_wait();
        uint32_t m8 = (seg_mask >> (cg * 8)) & 0xffu;
        if (cg == 0) m8 = (m8 & ~1u) | first_bit;
        if (__builtin_expect(m8 == 0, 1)) {
#pragma unroll
          for (int q = 0; q < 8; ++q) update(lg[q], vv[q]);
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (m8 & (1u << q)) {
              if (run_seg >= 0) { flush_run(run_whole); run_whole = true; }
              run_seg = ci[cg * 8 + q]; run_m = lg[q]; run_s = 0.f; run_a = 0.f;
            }
            update(lg[q], vv[q]);
          }
        }
      }
      if (g + 1 >= t_end && run_seg >= 0) { flush_run(run_whole && chunk_whole_end); run_seg = -1; }   // (also true on padding tiles: nothing open)
    }
    seg_mask = seg_next;
#pragma unroll
    for (int i = 0; i < NPRE; ++i) aq[i] = aq_next[i];
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// grid size and 16-edge tiles per stream: grid * SLOTS * SUB streams, half of them per side; a stream (= softmax piece)
// keeps at least two tiles (32 edges) whenever the input has that many, so the partial-slot table sized by the
// forward (clusters + edges / 32 + 2 slots per side) always suffices
void sp_chunking(int64_t n_e01, int sm_count, int* grid, int* tiles_per_chunk) {
  const int n_tiles = ceil_div(n_e01, SP_TS);
  const int streams_per_cta = SP_SLOTS * SP_SUB;
  int g = sm_count;
  const int want = (int)((2 * (int64_t)n_tiles) / (2 * streams_per_cta));      // CTAs that still get >= 2 tiles per stream
  if (g > want) g = want > 0 ? want : 1;
  const int chunks_side = (g * streams_per_cta) / 2;
  *grid = g;
  *tiles_per_chunk = ceil_div(n_tiles > 0 ? n_tiles : 1, chunks_side);
}

}  // namespace

int split_piece_edges(int64_t n_e01, const cns_handle* h) {
  int grid, tpc;
  sp_chunking(n_e01, h->sm_count, &grid, &tpc);
  return tpc * SP_TS;
}

int split_pack_weights(cns_handle* h, cudaStream_t s) {
  if (!h->tc_w_split) CNS_CUDA_CHECK(cudaMalloc(&h->tc_w_split, (size_t)2 * SP_NW * SP_W_BYTES));
  const int total = SP_NW * 128 * 128;
  if (h->split_f16) sp_pack_weights_kernel<true><<<ceil_div(total, 256), 256, 0, s>>>(h->pk.tc_chain, h->tc_w_split);
  else sp_pack_weights_kernel<false><<<ceil_div(total, 256), 256, 0, s>>>(h->pk.tc_chain, h->tc_w_split);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int enc_edge_split(const EncEdgeArgs& a, const cns_handle* h, cudaStream_t s) {
  if (a.n_e01 == 0) return CNS_OK;
  CNS_REQUIRE(h->tc_w_split != nullptr, "split weight image missing");
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(enc_edge_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM_BYTES));
    CNS_CUDA_CHECK(cudaFuncSetAttribute(enc_edge_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM_BYTES));
  }
  SpArgs ta;
  ta.e = a;
  ta.w_img = h->tc_w_split;
  ta.n_tiles = ceil_div(a.n_e01, SP_TS);
  int grid, tpc;
  sp_chunking(a.n_e01, h->sm_count, &grid, &tpc);
  ta.tiles_per_chunk = tpc;
  static_assert(sizeof(SpSmallW) == sizeof(h->tc_small_host), "first-layer weight table");
  const SpSmallW* sw = reinterpret_cast<const SpSmallW*>(h->tc_small_host);
  if (h->split_f16) enc_edge_split_kernel<true><<<grid, 128 * SP_SUB * SP_SLOTS, SP_SMEM_BYTES, s>>>(ta, *sw);
  else enc_edge_split_kernel<false><<<grid, 128 * SP_SUB * SP_SLOTS, SP_SMEM_BYTES, s>>>(ta, *sw);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns
