// tcgen05 (5th-gen tensor core) version of the encoder edge kernel (CNS_PREC_BF16 / CNS_PREC_F16).
//
// One persistent CTA per SM.  The five 128x128 weight matrices of the keypoint->cluster attention
// chain (cns/models/graph_vs.py:195-214) stay resident for the whole kernel, fetched once per CTA with 1-D TMA
// bulk copies: the three in front of the round-A commit in TENSOR memory (A operand of TS-mode MMAs, no
// shared-memory fetch per instruction), the other two in shared memory (64 KB).  Edges are processed
// in tiles of TN = 32 edges; each tile is a "slot" owned by 4 warps (thread == output channel == TMEM
// lane) and SLOTS = 5 tiles are in flight per CTA (20 warps; 320 accumulator + 192 weight TMEM columns) so that
// the SIMT epilogues of one tile overlap the tensor-core work of the others - round 1 kept four matrices in TMEM
// plus all five in shared memory and had room for 4 slots only (0.264 ms); 5 slots: 0.227 ms, 6 slots (two matrices
// in TMEM, more shared-memory A fetches): 0.227 ms.  Accumulators are transposed (D[channel][edge]): the
// per-channel segment softmax over the edges of a cluster is then a sequential walk along TMEM
// columns inside one thread, with warp-uniform segment boundaries - no shuffles, no atomics.
//
// attn_nn.fcs.0 (A0) is linear in its input a_dst[i] - lin_src x_j + delta_ij, so it is composed into
// the producing weights once per weight load (handle.cu) and a whole GEMM round disappears:
//
//   round A :  U = x (-(A0 lin_src))^T + hp (A0 pos_nn.1)^T      V = x lin^T + hp pos_nn.1^T    (K = 256 each)
//   epi   A :  u = relu(U + a'[cluster]),  a' = A0 (a_dst + b_pos1) + b_attn0 (per cluster, enc_cluster_pre)
//              -> 16-bit -> smem, MN-major ([channel][edge]): every thread stores its own row, no transposition
//   round C :  L = u attn_nn.1^T                final : online segment softmax of (L + b_a1) weighting (V + b_p1)
//
// x = relu(in_enc(feat)) and hp = relu(pos_nn.0(dpos)) have K = 2 / 4 and are produced by FFMA straight
// into the swizzled B-operand layout.  Shared-memory operand tiles use the canonical K-major
// SWIZZLE_128B layout ([k/64][row][128 B], 16-byte chunk index XOR (row & 7)).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

constexpr int TC_W_BYTES = 128 * 128 * 2;         // one weight matrix, 16-bit
constexpr int TC_NW = 5;                          // pos_nn.1, -(A0 lin_src), lin, A0 pos_nn.1, attn_nn.1

// ---- weight image: fp32 (k-major "chain" [5][k][o]) -> 16-bit swizzled A-operand tiles ----------
template <bool BF16>
__global__ void tc_pack_weights_kernel(const float* __restrict__ enc_chain, unsigned char* __restrict__ img) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= TC_NW * 128 * 128) return;
  int w = idx / 16384, rem = idx % 16384;
  int k = rem / 128, o = rem % 128;               // enc_chain[w][k][o] = W_w[o][k]
  float v = enc_chain[idx];
  if (w == 4) v *= 1.4426950408889634f;           // attn_nn.1 emits logits in log2 units (softmax with ex2)
  unsigned short bits;
  if (BF16) bits = __bfloat16_as_ushort(__float2bfloat16_rn(v));
  else bits = __half_as_ushort(__float2half_rn(v));
  *reinterpret_cast<unsigned short*>(img + (size_t)w * TC_W_BYTES + sw128_off(128, o, k)) = bits;
}

struct TcSmallW {          // passed by value in the kernel parameter space (constant bank)
  float4 w[128][2];        // per channel k: {in_enc.w[k][0], in_enc.w[k][1], in_enc.b[k], pos_nn.0.b[k]}, pos_nn.0.w[k][0..3]
};

// the same first-layer weights as fp16 pairs (channel k in the low half, k+1 in the high half) for the packed
// HFMA2 producer of the f16 mode: per channel pair {in_enc.w[.][0], in_enc.w[.][1], in_enc.b, pos_nn.0.b}, pos_nn.0.w[.][0..3]
struct TcSmallH {
  uint4 w[64][2];
};

struct TcArgs {
  EncEdgeArgs e;
  const unsigned char* w_img;   // 5 swizzled 16-bit weight tiles
  const unsigned char* w_first; // f16 mode: first-layer A tile (tc_mma_first)
  int n_tiles;                  // edge tiles of TN edges
  int tiles_per_chunk;          // consecutive tiles handled by one slot (per side)
  int side_mode;                // 0: both sides (first half of the slot chunks = current frame, second half = target),
                                // 1: current side only, 2: target side only
  float* store_l; __half* store_v;   // STORE kernels: per-keypoint logits (log2 units, fp32) and values (f16) instead of the softmax
};

// =================================================================================================
// Which weight matrices live where, by the number of tiles in flight.  TMEM has 512 columns: SLOTS * 2 * TN for the
// accumulators, 64 per resident matrix.  4 slots: matrices 0..3 in tensor memory, attn_nn.1 (4) in shared memory.
// 5 slots: the three matrices in front of the round-A commit (1, 2, 3: U and V's x half) in tensor memory, 0 and 4 in
// shared memory.  6 slots: U's two matrices (1, 3) in tensor memory.  -1 = not there.
__host__ __device__ constexpr int tc_tmem_slot(int slots, int w) {
  return slots <= 4 ? (w < 4 ? w : -1) : slots == 5 ? (w == 1 ? 0 : w == 2 ? 1 : w == 3 ? 2 : -1) : (w == 1 ? 0 : w == 3 ? 1 : -1);
}
__host__ __device__ constexpr int tc_smem_slot(int slots, int w) {
  return slots <= 4 ? (w == 4 ? 0 : -1) : slots == 5 ? (w == 0 ? 0 : w == 4 ? 1 : -1) : (w == 0 ? 0 : w == 2 ? 1 : w == 4 ? 2 : -1);
}
__host__ __device__ constexpr int tc_n_tmem(int slots) { return slots <= 4 ? 4 : slots == 5 ? 3 : 2; }
// resident shared-memory matrices + activation tiles; the same area first stages the TMEM-bound matrices
__host__ __device__ constexpr size_t tc_main_bytes(int slots, int tn) {
  const size_t run = (size_t)(TC_NW - tc_n_tmem(slots)) * TC_W_BYTES + (size_t)slots * 2 * tn * 256;   // x (later u) and hp tiles per slot
  const size_t stage = (size_t)tc_n_tmem(slots) * TC_W_BYTES;
  return run > stage ? run : stage;
}
// f16 mode, 32-edge tiles, <= 5 slots: the two first layers (K = 2 / 4 inputs) run on the tensor core as one K = 16
// step: a 16 KB A tile [channel][16 k of in_enc | 16 k of pos_nn.0] and one [edge][16 k] input tile (4 KB) per slot
__host__ __device__ constexpr bool tc_mma_first(int slots, int tn, bool bf16) { return !bf16 && tn == 32 && slots <= 5; }
constexpr int TC_FIRST_W_BYTES = 128 * 128;       // one SWIZZLE_128B tile of 128 rows; 64 of the 128 bytes per row are used
__host__ __device__ constexpr size_t tc_first_bytes(int slots, int tn) {
  return tc_mma_first(slots, tn, false) ? (size_t)TC_FIRST_W_BYTES + (size_t)slots * tn * 128 : 0;
}
constexpr int TC_NSTG = 8;                        // a' rows staged in shared memory per tile (its first 8 segment starts)
__host__ __device__ constexpr size_t tc_smem_bytes(int slots, int tn) {
  return tc_main_bytes(slots, tn) + tc_first_bytes(slots, tn) + (size_t)slots * TC_NSTG * 128 * sizeof(float) + 256 +
         (size_t)slots * 4 * tn * sizeof(int) + 1024;
}

template <int TN, bool BF16, int SLOTS_, bool STORE = false>
__global__ void __launch_bounds__(128 * SLOTS_, 1)
enc_edge_tc_kernel(const __grid_constant__ TcArgs args, const __grid_constant__ TcSmallW sw, const __grid_constant__ TcSmallH sh) {
  constexpr int SLOTS = SLOTS_;
  constexpr int ACT_TILE = TN * 256;                 // one [TN x 128] 16-bit operand tile
  constexpr int TMEM_COLS = 512;                     // [0, ACC): SLOTS * 2 regions * TN accumulator columns; then 64 per resident matrix
  constexpr int TMEM_W = tc_n_tmem(SLOTS);           // weight matrices kept in tensor memory (A operand without a smem fetch per MMA)
  constexpr int SMEM_W = TC_NW - TMEM_W;             // ... and in shared memory
  constexpr int TMEM_W_COL0 = SLOTS * 2 * TN;
  static_assert(TMEM_W_COL0 + TMEM_W * 64 <= 512, "tensor memory budget");
  constexpr int U_ROW = TN * 2;                      // bytes per channel row of the MN-major u tile
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment: align by hand (1 KB of slack is allocated)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* w_s = smem;                                     // SMEM_W x 32 KB (staging area of the TMEM-bound matrices first)
  unsigned char* act_s = smem + SMEM_W * TC_W_BYTES;             // SLOTS x 2 x ACT_TILE
  constexpr bool MMA_FIRST = tc_mma_first(SLOTS, TN, BF16);
  unsigned char* wfirst_s = smem + tc_main_bytes(SLOTS, TN);     // first-layer A tile, then SLOTS input tiles [TN][128 B]
  unsigned char* in16_s = wfirst_s + TC_FIRST_W_BYTES;
  float* astg_s = reinterpret_cast<float*>(smem + tc_main_bytes(SLOTS, TN) + tc_first_bytes(SLOTS, TN));   // [SLOTS][TC_NSTG][128]
  unsigned char* tail = reinterpret_cast<unsigned char*>(astg_s + SLOTS * TC_NSTG * 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);            // [0] = weights landed, [1+s] = MMAs of slot s done
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tail + 128);
  int* ci_s = reinterpret_cast<int*>(tail + 256);                // [SLOTS][2][TN] cluster id per edge (-1: padding), double-buffered
  int* ej_s = ci_s + SLOTS * 2 * TN;                             // STORE: [SLOTS][2][TN] keypoint id per edge

  const EncEdgeArgs& p = args.e;
  const int tid = threadIdx.x, lane = tid & 31;
  // warp-uniform by construction: lets the compiler keep slot addresses / descriptors in uniform registers
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(smem_u32(&bars[0]), 1);
      for (int s = 0; s < SLOTS; ++s) mbar_init(smem_u32(&bars[1 + s]), 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), TMEM_COLS);
    if (lane == 0) {
      // staging pass: the TMEM-bound matrices land in [w_s, ...) (weight + activation area), slot order
      mbar_expect_tx(smem_u32(&bars[0]), TMEM_W * TC_W_BYTES);
      for (int w = 0; w < TC_NW; ++w)
        if (tc_tmem_slot(SLOTS, w) >= 0)
          bulk_g2s(smem_u32(w_s + tc_tmem_slot(SLOTS, w) * TC_W_BYTES), args.w_img + (size_t)w * TC_W_BYTES, TC_W_BYTES,
                   smem_u32(&bars[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);
  // Those matrices move on into tensor memory: with N = 32 an MMA whose A operand sits in shared memory spends most
  // of its time fetching the 4 KB A slice (measured ~80 cycles per instruction against 16 of math).  Row o of a
  // matrix = TMEM lane o, k and k+1 share a 32-bit column.  Warp group g = warp / 4 moves matrices g, g + SLOTS, ...
  {
    mbar_wait(smem_u32(&bars[0]), 0);
    const int o = tid & 127;                         // TMEM lane
    const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + TMEM_W_COL0;
#pragma unroll 1
    for (int w = warp >> 2; w < TMEM_W; w += SLOTS) {
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {                  // 16 k per trip
        const uint4 lo = *reinterpret_cast<const uint4*>(w_s + w * TC_W_BYTES + sw128_off(128, o, g * 16));
        const uint4 hi = *reinterpret_cast<const uint4*>(w_s + w * TC_W_BYTES + sw128_off(128, o, g * 16 + 8));
        float v[8];
        uint32_t* u = reinterpret_cast<uint32_t*>(v);
        u[0] = lo.x; u[1] = lo.y; u[2] = lo.z; u[3] = lo.w; u[4] = hi.x; u[5] = hi.y; u[6] = hi.z; u[7] = hi.w;
        tmem_st8(t_row + (uint32_t)(w * 64 + g * 8), v);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  fence_proxy_async();          // generic-proxy reads of the staging area before the TMA writes that follow
  __syncthreads();
  tc_fence_after();
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(smem_u32(&bars[0]), SMEM_W * TC_W_BYTES + (MMA_FIRST ? TC_FIRST_W_BYTES : 0));
    for (int w = 0; w < TC_NW; ++w)
      if (tc_smem_slot(SLOTS, w) >= 0)
        bulk_g2s(smem_u32(w_s + tc_smem_slot(SLOTS, w) * TC_W_BYTES), args.w_img + (size_t)w * TC_W_BYTES, TC_W_BYTES,
                 smem_u32(&bars[0]));
    if (MMA_FIRST) bulk_g2s(smem_u32(wfirst_s), args.w_first, TC_FIRST_W_BYTES, smem_u32(&bars[0]));
  }

  // instruction descriptor: D=f32 (bit 4), A/B format at bits 7/10 (0 f16, 1 bf16), A K-major (bit 15 = 0),
  // B K-major or MN-major (bit 16), N>>3 at bit 17, M>>4 at bit 24
  const uint32_t fmt = BF16 ? 1u : 0u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(TN >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_mn = idesc | (1u << 16);
  const uint32_t w_base = smem_u32(w_s);

  {
    // ============================== producer / epilogue warps ===============================
    const int s = warp >> 2;                       // slot
    const int t = tid & 127;                       // thread in slot == output channel == TMEM lane
    const int wq = warp & 3;                       // TMEM lane quarter of this warp
    unsigned char* buf0 = act_s + (s * 2 + 0) * ACT_TILE;      // x, then u
    unsigned char* buf1 = act_s + (s * 2 + 1) * ACT_TILE;      // hp
    int* const ci_base = ci_s + s * 2 * TN;
    const uint32_t done = smem_u32(&bars[1 + s]);
    const uint32_t t0 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((s * 2 + 0) * TN);
    const uint32_t t1 = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)((s * 2 + 1) * TN);
    const float b_p1 = p.b_p1[t];
    uint32_t done_parity = 0;
    constexpr float LN2 = 0.6931471805599453f;

    // One elected thread of the slot's first warp issues the slot's tensor-core work (tcgen05.mma is a
    // single-thread instruction): after the slot barrier every operand byte written by the 128 threads is
    // visible to the async proxy, so no dedicated MMA warp / "ready" mbarrier is needed.
    // One K=128 product = 8 MMAs of K = 16.
    // Descriptors differ only in their start-address field ((addr >> 4) in the low word; shared memory is
    // < 256 KB, so adding a 16-byte-unit offset never carries out of the field): the high words and the
    // low-word bases are loop invariants, one uniform add per operand per MMA.
    const uint64_t hi_k = sw128_desc(0) & 0xffffffff00000000ull, hi_mn = mn_major_desc(0, U_ROW) & 0xffffffff00000000ull;
    const uint32_t lo_mn = (uint32_t)mn_major_desc(0, U_ROW);          // LBO placeholder bit
    const uint32_t w_lo = w_base >> 4, b0_lo = smem_u32(buf0) >> 4, b1_lo = smem_u32(buf1) >> 4;
    auto gemm_kmajor = [&](uint32_t d, int w, uint32_t b_lo, uint32_t acc_first) {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = hi_k | (uint64_t)(b_lo + (uint32_t)((kc * (TN * 128) + ks * 32) >> 4));
          if (tc_tmem_slot(SLOTS, w) >= 0) {
            umma_f16_ts(d, tmem_base + (uint32_t)(TMEM_W_COL0 + tc_tmem_slot(SLOTS, w) * 64 + (kc * 4 + ks) * 8), bd, idesc,
                        (kc | ks) ? 1u : acc_first);
          } else {
            const uint64_t ad = hi_k | (uint64_t)(w_lo + (uint32_t)((tc_smem_slot(SLOTS, w) * TC_W_BYTES + kc * (128 * 128) + ks * 32) >> 4));
            umma_f16(d, ad, bd, idesc, (kc | ks) ? 1u : acc_first);
          }
        }
    };
    // the same with the B tile stored MN-major ([k][edge], like the u tile): x / hp tiles written by the first-layer epilogue
    auto gemm_mn_any = [&](uint32_t d, int w, uint32_t b_lo, uint32_t acc_first) {
#pragma unroll
      for (int kc = 0; kc < 2; ++kc)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bd = hi_mn | (uint64_t)(lo_mn + b_lo + (uint32_t)(((kc * 4 + ks) * (16 * U_ROW)) >> 4));
          if (tc_tmem_slot(SLOTS, w) >= 0) {
            umma_f16_ts(d, tmem_base + (uint32_t)(TMEM_W_COL0 + tc_tmem_slot(SLOTS, w) * 64 + (kc * 4 + ks) * 8), bd, idesc_mn,
                        (kc | ks) ? 1u : acc_first);
          } else {
            const uint64_t ad = hi_k | (uint64_t)(w_lo + (uint32_t)((tc_smem_slot(SLOTS, w) * TC_W_BYTES + kc * (128 * 128) + ks * 32) >> 4));
            umma_f16(d, ad, bd, idesc_mn, (kc | ks) ? 1u : acc_first);
          }
        }
    };
    auto gemm_mnmajor = [&](uint32_t d, int w, uint32_t b_lo) {     // B tile stored [k][edge]
#pragma unroll
      for (int kc = 0; kc < 2; ++kc)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ad = hi_k | (uint64_t)(w_lo + (uint32_t)((tc_smem_slot(SLOTS, w) * TC_W_BYTES + kc * (128 * 128) + ks * 32) >> 4));
          const uint64_t bd = hi_mn | (uint64_t)(lo_mn + b_lo + (uint32_t)(((kc * 4 + ks) * (16 * U_ROW)) >> 4));   // 16 k-rows per MMA
          umma_f16(d, ad, bd, idesc_mn, (kc | ks) ? 1u : 0u);
        }
    };
    auto publish_and_issue = [&](int round) {
      tc_fence_before();          // this thread's earlier TMEM loads are ordered before the next MMAs
      fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core
      named_bar(1 + s, 128);
      if (wq == 0) {
        if (elect_one()) {
          tc_fence_after();
          const uint32_t b0 = b0_lo, b1 = b1_lo;
          const uint32_t d0 = tmem_base + (uint32_t)((s * 2 + 0) * TN);
          const uint32_t d1 = tmem_base + (uint32_t)((s * 2 + 1) * TN);
          if (MMA_FIRST && round == 2) {
            // first layers: X = W_in [f | 1], HP = W_p0 [dpos | 1] (K = 16: fp16 head + tail of every input, bias column)
            const uint32_t wf_lo = smem_u32(wfirst_s) >> 4, in_lo = smem_u32(in16_s + s * (TN * 128)) >> 4;
            umma_f16(d0, hi_k | (uint64_t)wf_lo, hi_k | (uint64_t)in_lo, idesc, 0u);
            umma_f16(d1, hi_k | (uint64_t)(wf_lo + 2), hi_k | (uint64_t)in_lo, idesc, 0u);
            umma_commit(done);
          } else if (MMA_FIRST && round == 0) {
            gemm_mn_any(d1, 2, b0, 0);   // V  = x  lin^T               (x, hp tiles are MN-major here)
            gemm_mn_any(d0, 1, b0, 0);   // U  = x  (-(A0 lin_src))^T
            gemm_mn_any(d0, 3, b1, 1);   //    + hp (A0 pos_nn.1)^T
            umma_commit(done);
            gemm_mn_any(d1, 0, b1, 1);   // V += hp pos_nn.1^T
          } else if (round == 0) {
            // The epilogue needs U complete and buffer 0 (x) free: everything that reads buffer 0 goes first and the
            // commit follows U.  V's hp half only reads buffer 1: it runs behind the commit, under the epilogue, and
            // is covered by the round-C commit (a commit tracks every earlier MMA of the issuing thread).
            // (Measured: giving u its own tile so that the commit can follow U's 16 MMAs alone changes nothing.)
            gemm_kmajor(d1, 2, b0, 0);   // V  = x  lin^T
            gemm_kmajor(d0, 1, b0, 0);   // U  = x  (-(A0 lin_src))^T
            gemm_kmajor(d0, 3, b1, 1);   //    + hp (A0 pos_nn.1)^T
            umma_commit(done);
            gemm_kmajor(d1, 0, b1, 1);   // V += hp pos_nn.1^T
          } else {
            gemm_mnmajor(d0, 4, b0);     // L = u attn_nn.1^T     (u lives in buffer 0, [channel][edge])
            umma_commit(done);
          }
        }
        __syncwarp();
      }
    };
    if (wq == 0) mbar_wait(smem_u32(&bars[0]), 1);     // shared-memory matrices landed (second phase; only the issuing warp cares)

    // Per-edge inputs of the producer (thread = edge pe, channel block kp) are fetched ahead: while tile i
    // waits for its round A, the coordinate gathers of tile i+1 (their indices arrived a tile ago) and the
    // index loads of tile i+2 are issued, so both latencies hide behind a whole tile of work.
    const int pe = (TN == 32) ? lane : ((wq & 1) * 32 + lane);   // == t % TN
    const int kp = (TN == 32) ? wq : (wq >> 1);                  // == t / TN (warp-uniform): channels [kp*TN, kp*TN + TN)
    // Work assignment: the 2 * n_tiles (side, edge tile) items are cut into gridDim.x * SLOTS contiguous
    // chunks, one per slot, the first half of the chunks on the current-frame side, the second half on
    // the target side.  Consecutive tiles in one slot let a cluster cut by a tile boundary carry its
    // running (max, sum, acc) in registers; only chunk boundaries leave partials for enc_finish.
    // (side_mode 1 / 2: every chunk belongs to that one side - the other side comes from a target cache, or this is the
    // kernel that fills one)
    const int n_chunks_side = args.side_mode ? (int)gridDim.x * SLOTS : ((int)gridDim.x * SLOTS) / 2;
    const int chunk_q = (int)blockIdx.x * SLOTS + s;
    const int side = args.side_mode ? args.side_mode - 1 : min(chunk_q / n_chunks_side, 1);
    const int chunk = chunk_q - (args.side_mode ? 0 : side * n_chunks_side);   // >= n_chunks_side: the odd stream out, no tiles
    const int tpc = args.tiles_per_chunk;
    const int t_begin = min(chunk * tpc, args.n_tiles), t_end = min(t_begin + tpc, args.n_tiles);
    auto tile_of = [&](int it) { return t_begin + it; };
    int idx_j = 0, idx_c = -1;                     // prefetched tile: source keypoint / cluster of edge pe
    float2 c_f = make_float2(0.f, 0.f), c_ps = c_f, c_pt = c_f, c_pc = c_f;   // coordinates of the tile produced next
    bool c_live = false;
    // ... and of the tile after it (gathers in flight for a whole tile)
    float2 n_f = c_f, n_ps = c_f, n_pt = c_f, n_pc = c_f;
    int n_cid = -1, n_j = 0;
    auto fetch_index = [&](int g) {
      idx_c = -1; idx_j = 0;
      if (g < t_end) {
        const int64_t e = (int64_t)g * TN + pe;
        if (e < p.n_e01) { idx_j = (int)p.ej[e]; idx_c = (int)p.ei[e]; }
      }
    };
    auto fetch_coords = [&]() {                    // gathers of the tile whose indices sit in idx_* -> n_*
      n_cid = idx_c; n_j = idx_j;
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
    // cluster ids travel through a double-buffered smem row: ci(i+1) is written while tile i waits for its
    // round A (everyone has left tile i-1 by then) and is published by tile i's second slot barrier
    // (with the first layers on the tensor core only the slot's first warp consumes the coordinates; letting the other
    // three warps skip the gathers was measured SLOWER, 0.202 against 0.193 ms - all four keep them)
    int* const ej_base = ej_s + s * 2 * TN;
    fetch_index(tile_of(0));
    if (kp == 0) { ci_base[pe] = idx_c; if (STORE) ej_base[pe] = idx_j; }
    fetch_coords();
    advance_coords();                              // c_* = tile 0
    fetch_index(tile_of(1));
    fetch_coords();                                // n_* = tile 1
    fetch_index(tile_of(2));
    named_bar(1 + s, 128);

    // running softmax state of the cluster that is open at the end of a tile (carried to the next tile)
    float run_m = 0.f, run_s = 0.f, run_a = 0.f;
    int run_seg = -1;
    float* const out_side = p.clu_feat + (size_t)side * p.n_clusters * H;
    float* const part_side = p.partial + (size_t)side * p.n_slots * 3 * H;
    // A run is a whole cluster unless it is the chunk's first run and the chunk starts inside a cluster, or
    // it is cut by the chunk's end: only those leave (max, sum, acc) partials for enc_finish.
    bool run_whole;                                // the open run started at a true cluster start
    bool chunk_whole_end;                          // the chunk's last edge is the last edge of its cluster
    {
      const int64_t ce0 = (int64_t)t_begin * TN, ce1 = min((int64_t)t_end * TN, p.n_e01);
      run_whole = (ce0 == 0) || (ce0 < p.n_e01 && p.ei[ce0 - 1] != p.ei[ce0]);
      chunk_whole_end = (ce1 >= p.n_e01) || (ce1 > 0 && p.ei[ce1 - 1] != p.ei[ce1]);
    }
    auto flush_run = [&](bool whole) {
      const float ab = fmaf(b_p1, run_s, run_a);
      if (whole) {
        out_side[(size_t)run_seg * H + t] = fmaxf(__fdividef(ab, run_s + 1e-16f), 0.f);
      } else {
        float* slot = part_side + ((size_t)run_seg + chunk) * 3 * H;
        slot[t] = run_m * LN2; slot[H + t] = run_s; slot[2 * H + t] = ab;     // natural-log units, like the fp32 kernel
      }
    };
    // u tile, MN-major: row = channel t (U_ROW bytes of edges), 16-byte chunk index XOR-swizzled by the row
    unsigned char* const u_row = buf0 + (t >> 3) * (8 * U_ROW) + (t & 7) * U_ROW;
    const uint32_t u_xor = (uint32_t)(((t & 7) * U_ROW) >> 7) & (uint32_t)(U_ROW / 16 - 1);

    // segment-start bits of a tile (bit e: ci[e] != ci[e-1]; warp-uniform registers, no per-element LDS) and the
    // a'[cluster][t] = (A0 (a_dst + b_pos1) + b_attn0)[t] values of its first TC_NSTG segment starts: staged in shared
    // memory by cp.async a tile ahead (no registers held, no scoreboard wait when the epilogue reaches a segment start;
    // register prefetch of 2 .. 8 rows was measured: the more rows, the slower)
    uint32_t seg_mask[TN / 32], seg_next[TN / 32];
    float* const astg = astg_s + s * (TC_NSTG * 128) + t;          // this thread's column: astg[k * 128]
    auto seg_and_prefetch = [&](const int* ci, uint32_t (&sm)[TN / 32]) {
#pragma unroll
      for (int c0 = 0; c0 < TN; c0 += 32) {
        const int e = c0 + lane;
        const int cprev = (e == 0) ? -2 : ci[e - 1];
        sm[c0 / 32] = __ballot_sync(0xffffffffu, ci[e] != cprev);
      }
      uint32_t m0 = sm[0];
      uint32_t m1 = (TN > 32) ? sm[(TN > 32) ? 1 : 0] : 0u;
#pragma unroll 1
      for (int k = 0; k < TC_NSTG; ++k) {
        int q;
        if (m0) { q = __ffs(m0) - 1; m0 &= m0 - 1; }
        else if (TN > 32 && m1) { q = 32 + __ffs(m1) - 1; m1 &= m1 - 1; }
        else break;
        const int c = ci[q];
        if (c >= 0) cp_async4(smem_u32(astg + k * 128), p.a_dst + (size_t)c * H + t);
        else astg[k * 128] = 0.f;
      }
      cp_async_commit();
    };
    seg_and_prefetch(ci_base, seg_mask);           // tile 0

    for (int it = 0;; ++it) {
      const int g = tile_of(it);
      if (g >= t_end) break;
      const int* const ci = ci_base + (it & 1) * TN;

      // ---- producer: x -> buf0, hp -> buf1 ---------------------------------------------------------
      {
        const bool live = c_live;
        const float f0 = live ? c_f.x : 0.f, f1 = live ? c_f.y : 0.f;
        const float d0 = live ? c_pc.x - c_ps.x : 0.f, d1 = live ? c_pc.y - c_ps.y : 0.f;   // pos_i - pos_j
        const float d2 = live ? c_pc.x - c_pt.x : 0.f, d3 = live ? c_pc.y - c_pt.y : 0.f;
        // kp and gq are warp-uniform: the first-layer weights are uniform constant-bank loads
        if constexpr (MMA_FIRST) {
          // f16 mode: the first layers run on the tensor core.  One warp writes the [edge][16 k] input tile: every input
          // as an fp16 head and an fp16 tail (v = hi + lo to ~22 bits: the geometry lives in small coordinate
          // differences), then a 1 for the bias column; products are exact and accumulate in fp32.
          if (kp == 0) {
            auto split = [](float v) {
              const __half hi = __float2half_rn(v);
              const __half lo = __float2half_rn(v - __half2float(hi));
              return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
            };
            unsigned char* in16 = in16_s + s * (TN * 128);
            *reinterpret_cast<uint4*>(in16 + sw128_off(TN, pe, 0)) = make_uint4(split(f0), split(f1), split(d0), split(d1));
            *reinterpret_cast<uint4*>(in16 + sw128_off(TN, pe, 8)) = make_uint4(split(d2), split(d3), 0x00003c00u, 0u);   // 1.0, 0
          }
        } else if constexpr (!BF16) {
          // f16 mode: packed half2 math, two channels per instruction, results land directly in operand format.
          // (inputs and weights rounded to fp16, fp16 accumulation of 3 / 5 terms: the result is rounded to
          // fp16 for the tensor core anyway)
          // every input is split into an fp16 head and an fp16 tail (v = hi + lo to ~22 bits): the geometry lives
          // in small coordinate differences, which a single fp16 would quantise to 2^-11
          auto split = [](float v, __half2& hi, __half2& lo) {
            hi = __float2half2_rn(v);
            lo = __float2half2_rn(v - __low2float(hi));
          };
          __half2 h_f0, l_f0, h_f1, l_f1, h_d0, l_d0, h_d1, l_d1, h_d2, l_d2, h_d3, l_d3;
          split(f0, h_f0, l_f0); split(f1, h_f1, l_f1);
          split(d0, h_d0, l_d0); split(d1, h_d1, l_d1); split(d2, h_d2, l_d2); split(d3, h_d3, l_d3);
          const __half2 zero = __float2half2_rn(0.f);
#pragma unroll 1
          for (int gq = 0; gq < TN / 8; ++gq) {
            const int k0 = kp * TN + gq * 8;
            uint32_t xo[4], ho[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 wa = sh.w[(k0 >> 1) + q][0], wb = sh.w[(k0 >> 1) + q][1];
              const __half2 w_i0 = *reinterpret_cast<const __half2*>(&wa.x), w_i1 = *reinterpret_cast<const __half2*>(&wa.y);
              const __half2 b_i = *reinterpret_cast<const __half2*>(&wa.z), b_p = *reinterpret_cast<const __half2*>(&wa.w);
              const __half2 w_p0 = *reinterpret_cast<const __half2*>(&wb.x), w_p1 = *reinterpret_cast<const __half2*>(&wb.y);
              const __half2 w_p2 = *reinterpret_cast<const __half2*>(&wb.z), w_p3 = *reinterpret_cast<const __half2*>(&wb.w);
              // tails first (small terms), heads last
              __half2 x2 = __hfma2(w_i1, l_f1, __hmul2(w_i0, l_f0));
              x2 = __hfma2(w_i1, h_f1, __hfma2(w_i0, h_f0, __hadd2(x2, b_i)));
              __half2 p2 = __hfma2(w_p3, l_d3, __hfma2(w_p2, l_d2, __hfma2(w_p1, l_d1, __hmul2(w_p0, l_d0))));
              p2 = __hfma2(w_p3, h_d3, __hfma2(w_p2, h_d2, __hfma2(w_p1, h_d1, __hfma2(w_p0, h_d0, __hadd2(p2, b_p)))));
              x2 = __hmax2(x2, zero);
              p2 = __hmax2(p2, zero);
              xo[q] = *reinterpret_cast<const uint32_t*>(&x2);
              ho[q] = *reinterpret_cast<const uint32_t*>(&p2);
            }
            const uint32_t off = sw128_off(TN, pe, k0);
            *reinterpret_cast<uint4*>(buf0 + off) = make_uint4(xo[0], xo[1], xo[2], xo[3]);
            *reinterpret_cast<uint4*>(buf1 + off) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
          }
        } else
#pragma unroll 1
        for (int gq = 0; gq < TN / 8; ++gq) {
          const int k0 = kp * TN + gq * 8;
          uint32_t xo[4], ho[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float xv[2], hv[2];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const float4 wa = sw.w[k0 + q * 2 + h2][0], wb = sw.w[k0 + q * 2 + h2][1];
              const float x = fmaf(wa.y, f1, fmaf(wa.x, f0, wa.z));
              float hp = fmaf(wb.x, d0, wa.w);
              hp = fmaf(wb.y, d1, hp);
              hp = fmaf(wb.z, d2, hp);
              hp = fmaf(wb.w, d3, hp);
              xv[h2] = fmaxf(x, 0.f);
              hv[h2] = fmaxf(hp, 0.f);
            }
            xo[q] = pack2<BF16>(xv[0], xv[1]);
            ho[q] = pack2<BF16>(hv[0], hv[1]);
          }
          const uint32_t off = sw128_off(TN, pe, k0);
          *reinterpret_cast<uint4*>(buf0 + off) = make_uint4(xo[0], xo[1], xo[2], xo[3]);
          *reinterpret_cast<uint4*>(buf1 + off) = make_uint4(ho[0], ho[1], ho[2], ho[3]);
        }
      }
      if constexpr (MMA_FIRST) {
        publish_and_issue(2);
        mbar_wait(done, done_parity); done_parity ^= 1;
        tc_fence_after();
        // x = relu(X) -> buf0, hp = relu(HP) -> buf1, both MN-major ([channel][edge]: this thread's rows)
        unsigned char* const hp_row = buf1 + (t >> 3) * (8 * U_ROW) + (t & 7) * U_ROW;
#pragma unroll 1
        for (int cg = 0; cg < TN / 8; ++cg) {
          float xa[8], ha[8];
          tmem_ld8(t0 + cg * 8, xa);
          tmem_ld8(t1 + cg * 8, ha);
          tmem_ld_wait();
          const uint32_t o16 = ((uint32_t)cg ^ u_xor) << 4;
          *reinterpret_cast<uint4*>(u_row + o16) = make_uint4(pack2_relu<BF16>(xa[0], xa[1]), pack2_relu<BF16>(xa[2], xa[3]),
                                                              pack2_relu<BF16>(xa[4], xa[5]), pack2_relu<BF16>(xa[6], xa[7]));
          *reinterpret_cast<uint4*>(hp_row + o16) = make_uint4(pack2_relu<BF16>(ha[0], ha[1]), pack2_relu<BF16>(ha[2], ha[3]),
                                                               pack2_relu<BF16>(ha[4], ha[5]), pack2_relu<BF16>(ha[6], ha[7]));
        }
      }
      publish_and_issue(0);
      if (kp == 0) {                                              // ci(i+1): its buffer was last read in tile i-1
        ci_base[((it + 1) & 1) * TN + pe] = n_cid;
        if (STORE) ej_base[((it + 1) & 1) * TN + pe] = n_j;
      }

      // ---- epilogue: u = relu(U + a'[cluster]) -> buf0, MN-major ([channel][edge]: this thread's row) ---
      // (padding columns of the last tile keep finite garbage; they are never read back)
      mbar_wait(done, done_parity); done_parity ^= 1;
      tc_fence_after();
      advance_coords();                // tile i+1: gathered a whole tile ago
      fetch_coords();                  // tile i+2: dependent gathers fly for a whole tile
      fetch_index(tile_of(it + 3));    // tile i+3: consumed a whole tile later
      {
        float adb = 0.f;
        int nseg = 0;
        // 8 edges per group = one 16-byte chunk of this thread's row.  The TMEM load of group g+1 is issued before
        // group g is processed (two register buffers), so only the first load of a phase is exposed.
        auto epi_group = [&](const float (&acc)[8], int cg) {
          const uint32_t m8 = (seg_mask[(TN > 32) ? (cg >> 2) : 0] >> ((cg & 3) * 8)) & 0xffu;
          uint32_t pk[4];
          if (__builtin_expect(m8 == 0, 1)) {
#pragma unroll
            for (int q = 0; q < 8; q += 2)
              pk[q / 2] = pack2_relu<BF16>(acc[q] + adb, acc[q + 1] + adb);
          } else {
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
              float v[2];
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                if (m8 & (1u << (q + h2))) {
                  if (nseg < TC_NSTG) {
                    adb = astg[nseg * 128];
                  } else {
                    const int c = ci[cg * 8 + q + h2];
                    adb = (c >= 0) ? p.a_dst[(size_t)c * H + t] : 0.f;
                  }
                  ++nseg;
                }
                v[h2] = acc[q + h2] + adb;
              }
              pk[q / 2] = pack2_relu<BF16>(v[0], v[1]);
            }
          }
          *reinterpret_cast<uint4*>(u_row + (((uint32_t)cg ^ u_xor) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        };
        cp_async_wait_all();             // this thread's staged a' values (issued a tile ago)
        // (16 columns per TMEM load - two exposed load latencies per tile instead of four - measured slower, 0.200 against
        // 0.190 ms, like the two-buffer software prefetch before it: the larger loop body costs more in instruction-cache
        // misses than the load latency it hides)
#pragma unroll 1
        for (int cg = 0; cg < TN / 8; ++cg) {
          float acc[8];
          tmem_ld8(t0 + cg * 8, acc);
          tmem_ld_wait();
          epi_group(acc, cg);
        }
      }
      publish_and_issue(1);
      seg_and_prefetch(ci_base + ((it + 1) & 1) * TN, seg_next);   // tile i+1: its a' values land during the softmax

      // ---- final: per-channel online segment softmax of L weighting V ---------------------------
      // softmax is shift-invariant per channel, so the attn_nn.1 bias drops out; the value bias b_p1
      // is added once per segment (sum p (v + b) = sum p v + b sum p).  One ex2 per element: of the
      // two rescale factors exp(m_old - m_new), exp(l - m_new) one is always 1.  A segment start only
      // resets the state (m = l, s = a = 0): the common update then yields s = 1, a = v.  Padding edges
      // (cluster -1) open a run that is never flushed.
      mbar_wait(done, done_parity); done_parity ^= 1;
      tc_fence_after();
      if constexpr (STORE) {
        // target-cache fill: logit (+ attn_nn.1 bias: shift-invariant, dropped) and value of every (keypoint, channel)
        const int* const ej = ej_base + (it & 1) * TN;
#pragma unroll 1
        for (int cg = 0; cg < TN / 8; ++cg) {
          float lg[8], vv[8];
          tmem_ld8(t0 + cg * 8, lg);
          tmem_ld8(t1 + cg * 8, vv);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (ci[cg * 8 + q] >= 0) {
              const size_t o = (size_t)ej[cg * 8 + q] * H + t;
              args.store_l[o] = lg[q];
              args.store_v[o] = __float2half_rn(vv[q]);
            }
          }
        }
      } else
      {
        // logits arrive in log2 units (log2(e) is folded into the attn_nn.1 weight image): one ex2, no multiply
        auto update = [&](float l, float v) {
          const float d = l - run_m;
          const float ex = ex2_approx(-fabsf(d));
          const bool up = d > 0.f;
          const float sc = up ? ex : 1.f, pr = up ? 1.f : ex;
          run_s = fmaf(run_s, sc, pr);
          run_a = fmaf(run_a, sc, pr * v);
          run_m = fmaxf(run_m, l);
        };
        // edge 0 of a tile opens a new run only if its cluster differs from the carried one
        const uint32_t first_bit = (ci[0] != run_seg) ? 1u : 0u;
        auto sm_group = [&](const float (&lg)[8], const float (&vv)[8], int cg) {
          uint32_t m8 = (seg_mask[(TN > 32) ? (cg >> 2) : 0] >> ((cg & 3) * 8)) & 0xffu;
          if (cg == 0) m8 = (m8 & ~1u) | first_bit;
          if (__builtin_expect(m8 == 0, 1)) {
            // no segment start among these 8 edges: one rescale for the group (to the group maximum), then
            // sub / ex2 / add / fma per element
            float gm = fmaxf(fmaxf(fmaxf(lg[0], lg[1]), fmaxf(lg[2], lg[3])), fmaxf(fmaxf(lg[4], lg[5]), fmaxf(lg[6], lg[7])));
            gm = fmaxf(gm, run_m);
            const float sc = ex2_approx(run_m - gm);
            run_s *= sc; run_a *= sc; run_m = gm;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float e = ex2_approx(lg[q] - gm);
              run_s += e;
              run_a = fmaf(e, vv[q], run_a);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (m8 & (1u << q)) {
                if (run_seg >= 0) { flush_run(run_whole); run_whole = true; }   // later runs start at a cluster start
                run_seg = ci[cg * 8 + q]; run_m = lg[q]; run_s = 0.f; run_a = 0.f;
              }
              update(lg[q], vv[q]);
            }
          }
        };
#pragma unroll 1
        for (int cg = 0; cg < TN / 8; ++cg) {
          float lg[8], vv[8];
          tmem_ld8(t0 + cg * 8, lg);
          tmem_ld8(t1 + cg * 8, vv);
          tmem_ld_wait();
          sm_group(lg, vv, cg);
        }
        if (g + 1 >= t_end && run_seg >= 0) { flush_run(run_whole && chunk_whole_end); run_seg = -1; }   // chunk end
      }
#pragma unroll
      for (int i = 0; i < TN / 32; ++i) seg_mask[i] = seg_next[i];
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =================================================================================================
template <int TN, bool BF16, int SLOTS, bool STORE = false>
static int launch_tc(const EncEdgeArgs& a, const unsigned char* w_img, const cns_handle* h, cudaStream_t s) {
  static unsigned long long configured = 0;
  auto kern = enc_edge_tc_kernel<TN, BF16, SLOTS, STORE>;
  constexpr size_t TC_SMEM_BYTES = tc_smem_bytes(SLOTS, TN);
  if (first_use_on_device(configured)) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
  }
  TcArgs ta;
  ta.e = a;
  ta.w_img = w_img;
  ta.w_first = h->tc_w_img + 2 * TC_NW * TC_W_BYTES;
  ta.n_tiles = ceil_div(a.n_e01, TN);
  int grid, tpc;
  tc_chunking(a.n_e01, TN, SLOTS, h->sm_count, a.side_mode ? 1 : 2, &grid, &tpc);
  ta.tiles_per_chunk = tpc;
  ta.side_mode = a.side_mode;
  ta.store_l = a.store_l; ta.store_v = reinterpret_cast<__half*>(a.store_v);
  CNS_REQUIRE(!STORE || (a.store_l && a.store_v && a.side_mode == 2), "enc_edge_tc: the store kernel fills a target cache");
  const int slots = SLOTS;
  // fp16 pair table of the first-layer weights (f16 mode), derived from the fp32 table on every launch: 2 KB
  const TcSmallW* sw = reinterpret_cast<const TcSmallW*>(h->tc_small_host);
  TcSmallH sh;
  for (int q = 0; q < 64; ++q) {
    const float4 a0 = sw->w[2 * q][0], a1 = sw->w[2 * q + 1][0], b0 = sw->w[2 * q][1], b1 = sw->w[2 * q + 1][1];
    auto pk = [](float lo, float hi) {
      const __half2 v = __floats2half2_rn(lo, hi);
      return *reinterpret_cast<const uint32_t*>(&v);
    };
    sh.w[q][0] = make_uint4(pk(a0.x, a1.x), pk(a0.y, a1.y), pk(a0.z, a1.z), pk(a0.w, a1.w));
    sh.w[q][1] = make_uint4(pk(b0.x, b1.x), pk(b0.y, b1.y), pk(b0.z, b1.z), pk(b0.w, b1.w));
  }
  kern<<<grid, 128 * slots, TC_SMEM_BYTES, s>>>(ta, *sw, sh);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

// grid size and tiles per slot-chunk for n_e01 edges: (grid * slots) / sides contiguous chunks per side (sides = 2: both
// sides in one launch; 1: a single side, the other one comes from / goes to a target cache)
void tc_chunking(int64_t n_e01, int tile_n, int slots, int sm_count, int sides, int* grid, int* tiles_per_chunk) {
  const int n_tiles = ceil_div(n_e01, tile_n);
  int g = sm_count;
  const int want = ceil_div(sides * (int64_t)n_tiles, slots);
  if (g > want) g = want > 0 ? want : 1;
  int chunks_side = (g * slots) / sides;
  if (chunks_side < 1) { chunks_side = 1; if (g * slots < sides) g = (sides + slots - 1) / slots; }
  *grid = g;
  *tiles_per_chunk = ceil_div(n_tiles > 0 ? n_tiles : 1, chunks_side);
}

static int tc_slots(const cns_handle* h) { return h->tc_tile_n == 64 ? 2 : (h->tc_slots >= 4 && h->tc_slots <= 6 ? h->tc_slots : 4); }

int tc_piece_edges(int64_t n_e01, const cns_handle* h, int sides) {
  int grid, tpc;
  tc_chunking(n_e01, h->tc_tile_n, tc_slots(h), h->sm_count, sides, &grid, &tpc);
  return tpc * h->tc_tile_n;
}

// =================================================================================================
// Target side from a cns_target_cache (forward.cu): the per-keypoint logits L (log2 units) and values V of the target side
// depend on the target alone, so a servo loop computes them once per target (store kernel above) and every frame only
// runs the per-channel segment softmax over the keypoints that are visible in that frame - a streaming pass over
// 768 bytes per visible keypoint instead of the five 128 x 128 GEMMs.  One CTA per cluster, thread = channel.
// =================================================================================================
struct TarCachedArgs {
  const int64_t* ej; const int* clu_rowptr; int64_t n_clusters;      // this frame: visible keypoints, cluster-major
  const float* lc; const __half* vc; const float* b_p1;
  // per target (cache): every member of every cluster (cluster-major) and the softmax statistics over ALL of them
  const int* all_src; const int* all_rowptr; const float* all_stats;  // (N,), (C+1,), (C,3,128): max | sum | weighted sum
  const uint8_t* visible;                                             // (N,) 1 = keypoint visible in this frame
  float* out;                                                         // target-side clu_feat (C,128); NULL: write statistics
  float* stats_out;
};

// One WARP per cluster: a warp reads a keypoint's whole row per instruction (lane = 4 channels: 16 bytes of logits,
// 8 bytes of values) and keeps four rows in flight.  No shared memory: this kernel shares the SMs with the encoder edge
// kernel (207 KB of shared memory, 45 k registers per SM), so only what fits beside it runs concurrently.
// A cluster that keeps most of its keypoints is not re-summed: its statistics over ALL members are cached, and only the
// rows of the MISSING members are read and subtracted (sum_visible = sum_all - sum_missing, both relative to the cached
// maximum) - 10 % of the rows at the usual 10 % of unmatched keypoints.  If the visible mass of any channel falls below
// 2^-8 of the total (a dominant keypoint went missing: the subtraction would cancel), the warp re-sums the visible rows.
__global__ void __launch_bounds__(128)
enc_tar_cached_kernel(const TarCachedArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (c >= p.n_clusters) return;
  const float* __restrict__ lc = p.lc;
  const __half* __restrict__ vc = p.vc;
  float m[4], sa[4], aa[4];
  auto row = [&](int64_t j, float4& l, uint2& v) {
    l = reinterpret_cast<const float4*>(lc + j * H)[lane];
    v = reinterpret_cast<const uint2*>(vc + j * H)[lane];
  };
  auto unpack = [](const float4 l, const uint2 vraw, float (&lv)[4], float (&vv)[4]) {
    const __half2 v01 = *reinterpret_cast<const __half2*>(&vraw.x), v23 = *reinterpret_cast<const __half2*>(&vraw.y);
    lv[0] = l.x; lv[1] = l.y; lv[2] = l.z; lv[3] = l.w;
    vv[0] = __low2float(v01); vv[1] = __high2float(v01); vv[2] = __low2float(v23); vv[3] = __high2float(v23);
  };
  auto take = [&](const float4 l, const uint2 vraw) {               // online softmax update
    float lv[4], vv[4];
    unpack(l, vraw, lv, vv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float gm = fmaxf(lv[i], m[i]);
      const float sc = ex2_approx(m[i] - gm), ee = ex2_approx(lv[i] - gm);     // m = -inf at the start: sc = 0
      sa[i] = fmaf(sa[i], sc, ee);
      aa[i] = fmaf(aa[i], sc, ee * vv[i]);
      m[i] = gm;
    }
  };
  // direct pass over a list of keypoint ids (int64 `ej` or int32 `all_src`): four rows in flight, next ids prefetched
  auto direct = [&](auto ids, int rb, int re) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; sa[i] = 0.f; aa[i] = 0.f; }
    int e = rb;
    int64_t jw = (e + (lane & 7) < re) ? (int64_t)ids[e + (lane & 7)] : 0;   // ids of rows e .. e+7 in lanes 0..7 (replicated)
    for (; e < re; e += 4) {
      const int64_t j0 = __shfl_sync(0xffffffffu, jw, 0), j1 = __shfl_sync(0xffffffffu, jw, 1);
      const int64_t j2 = __shfl_sync(0xffffffffu, jw, 2), j3 = __shfl_sync(0xffffffffu, jw, 3);
      const int n = re - e;
      float4 l0, l1 = make_float4(0.f, 0.f, 0.f, 0.f), l2 = l1, l3 = l1;
      uint2 v0, v1 = make_uint2(0u, 0u), v2 = v1, v3 = v1;
      row(j0, l0, v0);
      if (n > 1) row(j1, l1, v1);
      if (n > 2) row(j2, l2, v2);
      if (n > 3) row(j3, l3, v3);
      const int64_t jn_hi = (e + 8 + (lane & 3) < re) ? (int64_t)ids[e + 8 + (lane & 3)] : 0;
      const int64_t keep = __shfl_sync(0xffffffffu, jw, 4 + (lane & 3));
      jw = (lane & 7) < 4 ? keep : jn_hi;
      take(l0, v0);
      if (n > 1) take(l1, v1);
      if (n > 2) take(l2, v2);
      if (n > 3) take(l3, v3);
    }
  };
  if (!p.out) {                      // cache fill: statistics over all members of the cluster
    direct(p.all_src, p.all_rowptr[c], p.all_rowptr[c + 1]);
    float* st = p.stats_out + c * 3 * H;
    reinterpret_cast<float4*>(st)[lane] = make_float4(m[0], m[1], m[2], m[3]);
    reinterpret_cast<float4*>(st + H)[lane] = make_float4(sa[0], sa[1], sa[2], sa[3]);
    reinterpret_cast<float4*>(st + 2 * H)[lane] = make_float4(aa[0], aa[1], aa[2], aa[3]);
    return;
  }
  const int rb = p.clu_rowptr[c], re = p.clu_rowptr[c + 1];
  bool done = false;
  if (p.all_stats && re > rb) {
    const int ab0 = p.all_rowptr[c], ae0 = p.all_rowptr[c + 1];
    const int n_missing = (ae0 - ab0) - (re - rb);
    if (n_missing == 0 || 2 * n_missing < re - rb) {                 // subtract the missing members from the totals
      const float* st = p.all_stats + c * 3 * H;
      const float4 m4 = reinterpret_cast<const float4*>(st)[lane], s4 = reinterpret_cast<const float4*>(st + H)[lane];
      const float4 a4 = reinterpret_cast<const float4*>(st + 2 * H)[lane];
      m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w;
      sa[0] = s4.x; sa[1] = s4.y; sa[2] = s4.z; sa[3] = s4.w;
      aa[0] = a4.x; aa[1] = a4.y; aa[2] = a4.z; aa[3] = a4.w;
      const float tot[4] = {s4.x, s4.y, s4.z, s4.w};
      for (int base = ab0; base < ae0 && n_missing > 0; base += 32) {
        const int a = base + lane;
        const int j = a < ae0 ? p.all_src[a] : 0;
        uint32_t miss = __ballot_sync(0xffffffffu, a < ae0 && !p.visible[j]);
        while (miss) {                                               // up to four missing rows in flight
          int src[4]; int n = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) { src[k] = 0; if (miss) { src[k] = __ffs(miss) - 1; miss &= miss - 1; n = k + 1; } }
          float4 l[4]; uint2 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int jk = __shfl_sync(0xffffffffu, j, src[k]);
            if (k < n) row(jk, l[k], v[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < n) {
              float lv[4], vv[4];
              unpack(l[k], v[k], lv, vv);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float pr = ex2_approx(lv[i] - m[i]);
                sa[i] -= pr;
                aa[i] = fmaf(-pr, vv[i], aa[i]);
              }
            }
          }
        }
      }
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 4; ++i) ok = ok && sa[i] > tot[i] * 0.00390625f;
      done = __all_sync(0xffffffffu, ok);
    }
  }
  if (!done) direct(p.ej, rb, re);
  // sum p (v + b) = sum p v + b sum p ; relu ; no visible keypoint -> 0 (like the scatter of an empty segment)
  const float4 b4 = reinterpret_cast<const float4*>(p.b_p1)[lane];
  const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
  float o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = re > rb ? fmaxf(__fdividef(fmaf(bb[i], sa[i], aa[i]), sa[i] + 1e-16f), 0.f) : 0.f;
  reinterpret_cast<float4*>(p.out + c * H)[lane] = make_float4(o[0], o[1], o[2], o[3]);
}

__global__ void mark_visible_kernel(const int64_t* __restrict__ ej, int64_t n, uint8_t* __restrict__ flag) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) flag[ej[e]] = 1;
}

int enc_tar_cached(const EncTarCachedArgs& a, cudaStream_t s) {
  if (a.n_clusters == 0) return CNS_OK;
  TarCachedArgs p{};
  p.ej = a.ej; p.clu_rowptr = a.clu_rowptr; p.n_clusters = a.n_clusters; p.lc = a.lc; p.vc = reinterpret_cast<const __half*>(a.vc);
  p.b_p1 = a.b_p1; p.all_src = a.all_src; p.all_rowptr = a.all_rowptr; p.all_stats = a.all_stats; p.visible = a.visible;
  p.out = a.out; p.stats_out = a.stats_out;
  if (a.out && a.all_stats) {        // which keypoints this frame shows (the subtract path walks a cluster's full member list)
    CNS_CUDA_CHECK(cudaMemsetAsync(a.visible, 0, (size_t)a.n_nodes, s));
    if (a.n_e01 > 0) {
      mark_visible_kernel<<<ceil_div(a.n_e01, 256), 256, 0, s>>>(a.ej, a.n_e01, a.visible);
      CNS_LAUNCH_CHECK();
    }
  }
  enc_tar_cached_kernel<<<(unsigned)ceil_div(a.n_clusters, 4), 128, 0, s>>>(p);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int tc_pack_weights(cns_handle* h, cudaStream_t s) {
  if (!h->tc_w_img) CNS_CUDA_CHECK(cudaMalloc(&h->tc_w_img, 2 * TC_NW * TC_W_BYTES + TC_FIRST_W_BYTES));
  const int total = TC_NW * 128 * 128;
  tc_pack_weights_kernel<true><<<ceil_div(total, 256), 256, 0, s>>>(h->pk.tc_chain, h->tc_w_img);
  CNS_LAUNCH_CHECK();
  tc_pack_weights_kernel<false><<<ceil_div(total, 256), 256, 0, s>>>(h->pk.tc_chain, h->tc_w_img + TC_NW * TC_W_BYTES);
  CNS_LAUNCH_CHECK();
  // small first-layer weights travel in the kernel parameter space: keep a host copy
  static_assert(sizeof(TcSmallW) == sizeof(h->tc_small_host), "TcSmallW size");
  TcSmallW* sw = reinterpret_cast<TcSmallW*>(h->tc_small_host);
  float w_in[128 * 2], b_in[128], w_p0[128 * 4], b_p0[128];
  CNS_CUDA_CHECK(cudaMemcpyAsync(w_in, h->P(P_IN_W), sizeof(w_in), cudaMemcpyDeviceToHost, s));
  CNS_CUDA_CHECK(cudaMemcpyAsync(b_in, h->P(P_IN_B), sizeof(b_in), cudaMemcpyDeviceToHost, s));
  CNS_CUDA_CHECK(cudaMemcpyAsync(w_p0, h->P(P_POS0_W), sizeof(w_p0), cudaMemcpyDeviceToHost, s));
  CNS_CUDA_CHECK(cudaMemcpyAsync(b_p0, h->P(P_POS0_B), sizeof(b_p0), cudaMemcpyDeviceToHost, s));
  CNS_CUDA_CHECK(cudaStreamSynchronize(s));
  for (int k = 0; k < 128; ++k) {
    sw->w[k][0] = make_float4(w_in[2 * k], w_in[2 * k + 1], b_in[k], b_p0[k]);
    sw->w[k][1] = make_float4(w_p0[4 * k], w_p0[4 * k + 1], w_p0[4 * k + 2], w_p0[4 * k + 3]);
  }
  // first-layer A tile of the f16 tensor-core path (tc_mma_first): row = channel, k 0..15 = in_enc against the input
  // columns [f0 hi, f0 lo, f1 hi, f1 lo, dpos 0..3 hi/lo, 1, 0, 0, 0], k 16..31 = pos_nn.0 against the same columns
  static unsigned short first[TC_FIRST_W_BYTES / 2];
  memset(first, 0, sizeof(first));
  auto put = [&](int o, int k, float v) {
    first[sw128_off(128, o, k) / 2] = __half_as_ushort(__float2half_rn(v));
  };
  for (int o = 0; o < 128; ++o) {
    put(o, 0, w_in[2 * o]); put(o, 1, w_in[2 * o]); put(o, 2, w_in[2 * o + 1]); put(o, 3, w_in[2 * o + 1]);
    put(o, 12, b_in[o]);
    for (int i = 0; i < 4; ++i) { put(o, 16 + 4 + 2 * i, w_p0[4 * o + i]); put(o, 16 + 5 + 2 * i, w_p0[4 * o + i]); }
    put(o, 16 + 12, b_p0[o]);
  }
  CNS_CUDA_CHECK(cudaMemcpyAsync(h->tc_w_img + 2 * TC_NW * TC_W_BYTES, first, sizeof(first), cudaMemcpyHostToDevice, s));
  CNS_CUDA_CHECK(cudaStreamSynchronize(s));
  return CNS_OK;
}

int enc_edge_tc(const EncEdgeArgs& a, int precision, const cns_handle* h, cudaStream_t s) {
  if (a.n_e01 == 0) return CNS_OK;
  CNS_REQUIRE(h->tc_w_img != nullptr, "tensor-core weight image missing");
  const int tn = h->tc_tile_n, slots = tc_slots(h);
  if (a.store_l) {          // target-cache fill: one configuration per precision
    if (precision == CNS_PREC_BF16) return launch_tc<32, true, 5, true>(a, h->tc_w_img, h, s);
    return launch_tc<32, false, 5, true>(a, h->tc_w_img + TC_NW * TC_W_BYTES, h, s);
  }
  if (precision == CNS_PREC_BF16) {
    if (tn == 64) return launch_tc<64, true, 2>(a, h->tc_w_img, h, s);
    if (slots == 5) return launch_tc<32, true, 5>(a, h->tc_w_img, h, s);
    if (slots == 6) return launch_tc<32, true, 6>(a, h->tc_w_img, h, s);
    return launch_tc<32, true, 4>(a, h->tc_w_img, h, s);
  }
  if (precision == CNS_PREC_F16) {
    const unsigned char* img = h->tc_w_img + TC_NW * TC_W_BYTES;
    if (tn == 64) return launch_tc<64, false, 2>(a, img, h, s);
    if (slots == 5) return launch_tc<32, false, 5>(a, img, h, s);
    if (slots == 6) return launch_tc<32, false, 6>(a, img, h, s);
    return launch_tc<32, false, 4>(a, img, h, s);
  }
  set_error("enc_edge_tc: unsupported precision %d", precision);
  return CNS_ERR_UNSUPPORTED;
}

}  // namespace cns
