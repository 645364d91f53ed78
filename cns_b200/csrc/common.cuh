// Shared declarations for libcns_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/cns_b200.h"

namespace cns {

constexpr int H = CNS_HIDDEN_DIM;  // 128

void set_error(const char* fmt, ...);

#define CNS_CUDA_CHECK(expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      cns::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CNS_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

extern long long g_launch_count;   // kernels launched by this library (host-side counter)
#define CNS_LAUNCH_CHECK()        \
  do {                            \
    ++cns::g_launch_count;        \
    CNS_CUDA_CHECK(cudaGetLastError()); \
  } while (0)

#define CNS_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      cns::set_error("%s:%d %s", __FILE__, __LINE__, msg);       \
      return CNS_ERR_INVALID_ARGUMENT;                           \
    }                                                            \
  } while (0)

// Function attributes (dynamic shared memory opt-in) are per device: true the first time `seen` meets the current one.
inline bool first_use_on_device(unsigned long long& seen) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (d & 63);
  if (seen & bit) return false;
  seen |= bit;
  return true;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- canonical parameter table (state-dict order of checkpoints/cns_state_dict.pth) ----
enum ParamId {
  P_IN_W = 0, P_IN_B, P_POS0_W, P_POS0_B, P_POS1_W, P_POS1_B, P_ATT0_W, P_ATT0_B, P_ATT1_W, P_ATT1_B,
  P_LIN_W, P_LINSRC_W, P_LINDST_W, P_OUT_W, P_OUT_B,
  P_C0_MSG_W, P_C0_MSG_B, P_C0_BN_W, P_C0_BN_B, P_C0_BN_MS, P_C0_FC_W, P_C0_FC_B,
  P_G_MSG_W, P_G_MSG_B, P_K_MSG_W, P_K_MSG_B,
  P_C1_MSG_W, P_C1_MSG_B, P_C1_BN_W, P_C1_BN_B, P_C1_BN_MS, P_C1_FC_W, P_C1_FC_B,
  P_POOL_W, P_POOL_B, P_VEC0_W, P_VEC0_B, P_VEC1_W, P_VEC1_B, P_NRM0_W, P_NRM0_B, P_NRM1_W, P_NRM1_B,
  P_COUNT
};
static_assert(P_COUNT == CNS_NUM_PARAM_TENSORS, "param table");

struct ParamInfo { const char* name; int rows; int cols; };
extern const ParamInfo kParams[P_COUNT];

// Packed, device-resident GEMM operands derived from the raw parameters.
struct Packed {
  // encoder edge chain: 5 (k=128, o=128) matrices, k-major rows ("transposed" weights):
  //   0: pos_nn.fcs.1^T   1: -lin_src^T   2: lin^T   3: attn_nn.fcs.0^T   4: attn_nn.fcs.1^T
  float* enc_chain;      // [5][128][128]
  float* lin_dst_t;      // [128][128]
  // tensor-core encoder: attn_nn.fcs.0 is linear in (a_dst - a_src + delta), so it is composed into its
  // producers once per weight load (A0 = attn_nn.fcs.0):
  //   0: pos_nn.fcs.1^T   1: -(A0 lin_src)^T   2: lin^T   3: (A0 pos_nn.fcs.1)^T   4: attn_nn.fcs.1^T
  float* tc_chain;       // [5][128][128]
  float* tc_lin_dst_t;   // [128][128]  (A0 lin_dst)^T
  float* tc_adst_bias;   // [128]       A0 b_pos1 + b_attn0
  float* out_t;          // [256][128]
  // PointEdgeConv factorisation  PQ = X @ wt + pos @ wp + b ; columns [P (f) | Q (f)]
  float* c0_wt; float* c0_wp; float* c0_b;      // [128][256], [2][256], [256]
  float* g_wt;  float* g_wp;  float* g_b;       // [256][512], [2][512], [512]
  float* k_wt;  float* k_wp;  float* k_b;       // [256][256], [2][256], [256]
  float* c1_wt; float* c1_wp; float* c1_b;      // [128][256], [2][256], [256]
  float* c0_fc_t; float* c1_fc_t;               // [128][128]
  float* vec0_t; float* nrm0_t;                 // [128][128]
  // 16-bit swizzled B-operand images for gemm_tc: index 0 = bf16, 1 = f16
  unsigned char* tc_c0[2]; unsigned char* tc_c0_fc[2]; unsigned char* tc_g[2]; unsigned char* tc_k[2];
  unsigned char* tc_c1[2]; unsigned char* tc_c1_fc[2]; unsigned char* tc_out[2];
  unsigned char* tc_adst[2];                    // (A0 lin_dst): the per-cluster attention term a' as a gemm_tc product (f16 mode)
  unsigned char* tc_node_base;
  // (hi, lo) split images of the same seven matrices for the fp32 mode (gemm_tc_split.cu; slice format = cns_handle::split_f16)
  unsigned char* tc2_c0; unsigned char* tc2_c0_fc; unsigned char* tc2_g; unsigned char* tc2_k;
  unsigned char* tc2_c1; unsigned char* tc2_c1_fc; unsigned char* tc2_out;
  unsigned char* tc2_base;
  float* bf_vec;         // [3][1664] per-column epilogue vectors of the fused backbone kernel (bias | wp row 0 | wp row 1)
};

}  // namespace cns

struct cns_handle {
  int device;
  int precision;
  int sm_count;
  float* raw;                       // all 43 tensors, concatenated, fp32 (device)
  size_t raw_off[cns::P_COUNT + 1]; // element offsets into raw
  float* packed_base;
  cns::Packed pk;
  bool weights_loaded;
  long long weights_version;        // bumped by every (re)pack of the weights
  void* bwd_pack; long long bwd_pack_version;   // transposed operands of the backward pass (backward.cu)
  // tensor-core (tcgen05) encoder operands: bf16 image then f16 image, 5 swizzled 32 KB tiles each
  unsigned char* tc_w_img;
  unsigned char* tc_w_split;                            // fp32 mode: (hi, lo) slices of the same five matrices (encoder_split.cu)
  int no_split_encoder;                                 // CNS_NO_SPLIT_ENCODER=1: fp32 mode keeps the FFMA edge kernel (A-B runs)
  alignas(16) float tc_small_host[128 * 2 + 128 + 128 * 4 + 128];   // in_enc / pos_nn.0 weights passed as kernel params
  int tc_tile_n;                                        // edges per tensor-core tile (32 or 64)
  int tc_slots;                                         // tiles in flight per SM of the 32-edge kernel: 4, 5 or 6 (CNS_TC_SLOTS)
  int split_f16;                                        // slices of the fp32-tolerance kernels: 1 = two f16 (default), 0 = two bf16 (fp32's exponent range, 16 mantissa bits)
  int no_split_gemm;                                    // CNS_NO_SPLIT_GEMM=1: backward GEMMs stay on the fp32 FFMA kernel (A-B runs)
  int no_edge_subtract;                                 // CNS_NO_EDGE_SUBTRACT=1: cached target side always re-sums the visible keypoints (A-B runs)
  int no_edge_cache;                                    // CNS_NO_EDGE_CACHE=1: target caches keep no per-keypoint logits / values (A-B runs)
  int no_tc_cluster_pre;                                // CNS_NO_TC_CLUSTER_PRE=1: per-cluster encoder terms stay on the FFMA kernel (A-B runs)
  int no_fused_backbone;                                // CNS_NO_FUSED_BACKBONE=1: keep the per-layer kernels (debugging / A-B runs)
  // optional CUDA-event timing of the dominant kernel (encoder edge kernel), see cns_profile_*
  int timing; int ev_n; cudaEvent_t ev[2 * 256];
  // pinned / device staging for the *_host entry point
  void* stage_host; size_t stage_host_bytes;
  void* stage_dev;  size_t stage_dev_bytes;
  // recurrent hidden state of cns_servo_step_host: two ping-pong device buffers in their OWN allocation (the layout of
  // stage_dev changes from frame to frame with the number of visible keypoints / alive clusters)
  int servo_valid; int servo_cur; long long servo_nc;
  void* servo_hid; size_t servo_hid_bytes;
  // cns_servo_step_host keeps the target-side results while consecutive frames show the same target (host copy of the
  // target arrays = the key)
  struct cns_target_cache* servo_tcache; void* servo_tkey; size_t servo_tkey_bytes;
  // fork/join pair for work that is independent of the index preparation (per-cluster encoder terms), created lazily
  cudaStream_t side_stream; cudaEvent_t side_fork, side_join, side_join2;
  cudaStream_t low_stream; cudaEvent_t low_fork, low_join;      // lowest-priority stream: cached target side beside the edge kernel
  // CUDA-graph cache of cns_graphvs_forward (forward.cu): a call whose whole argument set (sizes, pointers, mode) was seen
  // before is replayed as one graph launch instead of ~10 kernel launches + 3 event fork/joins
  int use_graphs; void* graph_cache; cudaStream_t cap_stream;
  long long graph_replays, graph_captures, edge_cache_forwards;
  const float* P(int id) const { return raw + raw_off[id]; }
};
