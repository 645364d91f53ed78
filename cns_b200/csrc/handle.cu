// Handle lifetime, parameter table and weight packing.
// Replaces GraphVS.__init__ / load_state_dict of the reference (cns/models/graph_vs.py:273-279).
#include <stdarg.h>
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const ParamInfo kParams[P_COUNT] = {
    {"encoder.in_enc.fcs.0.weight", 128, 2},
    {"encoder.in_enc.fcs.0.bias", 128, 1},
    {"encoder.att_clu_conv.pos_nn.fcs.0.weight", 128, 4},
    {"encoder.att_clu_conv.pos_nn.fcs.0.bias", 128, 1},
    {"encoder.att_clu_conv.pos_nn.fcs.1.weight", 128, 128},
    {"encoder.att_clu_conv.pos_nn.fcs.1.bias", 128, 1},
    {"encoder.att_clu_conv.attn_nn.fcs.0.weight", 128, 128},
    {"encoder.att_clu_conv.attn_nn.fcs.0.bias", 128, 1},
    {"encoder.att_clu_conv.attn_nn.fcs.1.weight", 128, 128},
    {"encoder.att_clu_conv.attn_nn.fcs.1.bias", 128, 1},
    {"encoder.att_clu_conv.lin.weight", 128, 128},
    {"encoder.att_clu_conv.lin_src.weight", 128, 128},
    {"encoder.att_clu_conv.lin_dst.weight", 128, 128},
    {"encoder.out_enc.fcs.0.weight", 128, 256},
    {"encoder.out_enc.fcs.0.bias", 128, 1},
    {"backbone.l1_conv0.conv.msg_nn.fcs.0.weight", 128, 258},
    {"backbone.l1_conv0.conv.msg_nn.fcs.0.bias", 128, 1},
    {"backbone.l1_conv0.bn.weight", 128, 1},
    {"backbone.l1_conv0.bn.bias", 128, 1},
    {"backbone.l1_conv0.bn.mean_scale", 128, 1},
    {"backbone.l1_conv0.fc.weight", 128, 128},
    {"backbone.l1_conv0.fc.bias", 128, 1},
    {"backbone.temporal_aggr.conv_gates.msg_nn.fcs.0.weight", 256, 514},
    {"backbone.temporal_aggr.conv_gates.msg_nn.fcs.0.bias", 256, 1},
    {"backbone.temporal_aggr.conv_candi.msg_nn.fcs.0.weight", 128, 514},
    {"backbone.temporal_aggr.conv_candi.msg_nn.fcs.0.bias", 128, 1},
    {"backbone.l1_conv1.conv.msg_nn.fcs.0.weight", 128, 258},
    {"backbone.l1_conv1.conv.msg_nn.fcs.0.bias", 128, 1},
    {"backbone.l1_conv1.bn.weight", 128, 1},
    {"backbone.l1_conv1.bn.bias", 128, 1},
    {"backbone.l1_conv1.bn.mean_scale", 128, 1},
    {"backbone.l1_conv1.fc.weight", 128, 128},
    {"backbone.l1_conv1.fc.bias", 128, 1},
    {"decoder.pool.gate_nn.weight", 1, 128},
    {"decoder.pool.gate_nn.bias", 1, 1},
    {"decoder.vel_si_vec.fcs.0.weight", 128, 128},
    {"decoder.vel_si_vec.fcs.0.bias", 128, 1},
    {"decoder.vel_si_vec.fcs.1.weight", 6, 128},
    {"decoder.vel_si_vec.fcs.1.bias", 6, 1},
    {"decoder.vel_si_norm.fcs.0.weight", 128, 128},
    {"decoder.vel_si_norm.fcs.0.bias", 128, 1},
    {"decoder.vel_si_norm.fcs.1.weight", 1, 128},
    {"decoder.vel_si_norm.fcs.1.bias", 1, 1},
};

// dst[k][o] = sign * src[o][col0 + k]   (src has `ld` columns) : a (k x n_out) k-major operand
__global__ void pack_transpose_kernel(const float* __restrict__ src, int ld, int col0, int k_dim,
                                      int n_out, float sign, float* __restrict__ dst, int dst_ld,
                                      int dst_col0) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k_dim * n_out) return;
  int k = idx / n_out, o = idx % n_out;
  dst[(size_t)k * dst_ld + dst_col0 + o] = sign * src[(size_t)o * ld + col0 + k];
}

// PointEdgeConv message Linear W (f_out x (2*f_in + 2)), columns [x_i | x_j - x_i | pos_j - pos_i]
// (graph_vs.py:108).  Folded into  P = x_i (W1 - W2)^T - pos_i W3^T + b ,  Q = x_j W2^T + pos_j W3^T:
//   wt (f_in x 2 f_out) = [(W1-W2)^T | W2^T],  wp (2 x 2 f_out) = [-W3^T | W3^T],  b2 = [b | 0]
__global__ void pack_edgeconv_kernel(const float* __restrict__ w, const float* __restrict__ b, int f_in,
                                     int f_out, float* __restrict__ wt, float* __restrict__ wp,
                                     float* __restrict__ b2) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int ld = 2 * f_in + 2;
  int total = (f_in + 2 + 1) * f_out;
  if (idx >= total) return;
  int row = idx / f_out, o = idx % f_out;
  const float* wo = w + (size_t)o * ld;
  if (row < f_in) {
    float w1 = wo[row], w2 = wo[f_in + row];
    wt[(size_t)row * 2 * f_out + o] = w1 - w2;
    wt[(size_t)row * 2 * f_out + f_out + o] = w2;
  } else if (row < f_in + 2) {
    int d = row - f_in;
    float w3 = wo[2 * f_in + d];
    wp[(size_t)d * 2 * f_out + o] = -w3;
    wp[(size_t)d * 2 * f_out + f_out + o] = w3;
  } else {
    b2[o] = b[o];
    b2[f_out + o] = 0.f;
  }
}

static size_t packed_floats() {
  size_t n = 0;
  n += 5 * 128 * 128;                 // enc_chain
  n += 128 * 128;                     // lin_dst_t
  n += 5 * 128 * 128 + 128 * 128 + 128;   // tc_chain, tc_lin_dst_t, tc_adst_bias
  n += 256 * 128;                     // out_t
  n += 128 * 256 + 2 * 256 + 256;     // c0
  n += 256 * 512 + 2 * 512 + 512;     // gates
  n += 256 * 256 + 2 * 256 + 256;     // candi
  n += 128 * 256 + 2 * 256 + 256;     // c1
  n += 4 * 128 * 128;                 // c0_fc, c1_fc, vec0, nrm0
  return n;
}

// C[r][o] = sum_m S[r][m] A[m][o] (+ bias[o]); one block per row r, 128 columns
__global__ void compose_kernel(const float* __restrict__ S, const float* __restrict__ A, const float* __restrict__ bias,
                               float* __restrict__ C, int rows) {
  const int r = blockIdx.x, o = threadIdx.x;
  if (r >= rows) return;
  float acc = bias ? bias[o] : 0.f;
  for (int m = 0; m < 128; ++m) acc = fmaf(S[r * 128 + m], A[m * 128 + o], acc);
  C[r * 128 + o] = acc;
}

static int launch_transpose(const float* src, int ld, int col0, int k_dim, int n_out, float sign,
                            float* dst, int dst_ld, int dst_col0, cudaStream_t s) {
  int total = k_dim * n_out;
  pack_transpose_kernel<<<ceil_div(total, 256), 256, 0, s>>>(src, ld, col0, k_dim, n_out, sign, dst,
                                                             dst_ld, dst_col0);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

static int launch_edgeconv_pack(const float* w, const float* b, int f_in, int f_out, float* wt,
                                float* wp, float* b2, cudaStream_t s) {
  int total = (f_in + 3) * f_out;
  pack_edgeconv_kernel<<<ceil_div(total, 256), 256, 0, s>>>(w, b, f_in, f_out, wt, wp, b2);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

void free_bwd_pack(cns_handle* h);

int repack(cns_handle* h, cudaStream_t s) {
  Packed& pk = h->pk;
  int rc;
#define TR(pid, ld, col0, k, n, sign, dst, dld, dcol) \
  if ((rc = launch_transpose(h->P(pid), ld, col0, k, n, sign, dst, dld, dcol, s)) != CNS_OK) return rc;
  TR(P_POS1_W, 128, 0, 128, 128, 1.f, pk.enc_chain + 0 * 16384, 128, 0);
  TR(P_LINSRC_W, 128, 0, 128, 128, -1.f, pk.enc_chain + 1 * 16384, 128, 0);
  TR(P_LIN_W, 128, 0, 128, 128, 1.f, pk.enc_chain + 2 * 16384, 128, 0);
  TR(P_ATT0_W, 128, 0, 128, 128, 1.f, pk.enc_chain + 3 * 16384, 128, 0);
  TR(P_ATT1_W, 128, 0, 128, 128, 1.f, pk.enc_chain + 4 * 16384, 128, 0);
  TR(P_LINDST_W, 128, 0, 128, 128, 1.f, pk.lin_dst_t, 128, 0);
  TR(P_OUT_W, 256, 0, 256, 128, 1.f, pk.out_t, 128, 0);
  TR(P_C0_FC_W, 128, 0, 128, 128, 1.f, pk.c0_fc_t, 128, 0);
  TR(P_C1_FC_W, 128, 0, 128, 128, 1.f, pk.c1_fc_t, 128, 0);
  TR(P_VEC0_W, 128, 0, 128, 128, 1.f, pk.vec0_t, 128, 0);
  TR(P_NRM0_W, 128, 0, 128, 128, 1.f, pk.nrm0_t, 128, 0);
#undef TR
  if ((rc = launch_edgeconv_pack(h->P(P_C0_MSG_W), h->P(P_C0_MSG_B), 128, 128, pk.c0_wt, pk.c0_wp, pk.c0_b, s))) return rc;
  if ((rc = launch_edgeconv_pack(h->P(P_G_MSG_W), h->P(P_G_MSG_B), 256, 256, pk.g_wt, pk.g_wp, pk.g_b, s))) return rc;
  if ((rc = launch_edgeconv_pack(h->P(P_K_MSG_W), h->P(P_K_MSG_B), 256, 128, pk.k_wt, pk.k_wp, pk.k_b, s))) return rc;
  if ((rc = launch_edgeconv_pack(h->P(P_C1_MSG_W), h->P(P_C1_MSG_B), 128, 128, pk.c1_wt, pk.c1_wp, pk.c1_b, s))) return rc;
  // composed operands of the tensor-core encoder (k-major: C[k][o] = sum_m S[k][m] A0t[m][o])
  {
    const float* a0t = pk.enc_chain + 3 * 16384;
    CNS_CUDA_CHECK(cudaMemcpyAsync(pk.tc_chain + 0 * 16384, pk.enc_chain + 0 * 16384, 16384 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CNS_CUDA_CHECK(cudaMemcpyAsync(pk.tc_chain + 2 * 16384, pk.enc_chain + 2 * 16384, 16384 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CNS_CUDA_CHECK(cudaMemcpyAsync(pk.tc_chain + 4 * 16384, pk.enc_chain + 4 * 16384, 16384 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    compose_kernel<<<128, 128, 0, s>>>(pk.enc_chain + 1 * 16384, a0t, nullptr, pk.tc_chain + 1 * 16384, 128);
    CNS_LAUNCH_CHECK();
    compose_kernel<<<128, 128, 0, s>>>(pk.enc_chain + 0 * 16384, a0t, nullptr, pk.tc_chain + 3 * 16384, 128);
    CNS_LAUNCH_CHECK();
    compose_kernel<<<128, 128, 0, s>>>(pk.lin_dst_t, a0t, nullptr, pk.tc_lin_dst_t, 128);
    CNS_LAUNCH_CHECK();
    compose_kernel<<<1, 128, 0, s>>>(h->P(P_POS1_B), a0t, h->P(P_ATT0_B), pk.tc_adst_bias, 1);
    CNS_LAUNCH_CHECK();
  }
  if ((rc = tc_pack_weights(h, s)) != CNS_OK) return rc;
  if ((rc = split_pack_weights(h, s)) != CNS_OK) return rc;
  {
    // node-GEMM weight images (bf16 + f16), 2 bytes per element
    struct Item { const float* wt; int k, n; unsigned char** dst; };
    Item items[] = {{pk.c0_wt, 128, 256, pk.tc_c0}, {pk.c0_fc_t, 128, 128, pk.tc_c0_fc}, {pk.g_wt, 256, 512, pk.tc_g},
                    {pk.k_wt, 256, 256, pk.tc_k},   {pk.c1_wt, 128, 256, pk.tc_c1},   {pk.c1_fc_t, 128, 128, pk.tc_c1_fc},
                    {pk.out_t, 256, 128, pk.tc_out}, {pk.tc_lin_dst_t, 128, 128, pk.tc_adst}};
    if (!pk.tc_node_base) {
      size_t total = 0;
      for (auto& it : items) total += (size_t)it.k * it.n * 2 * 2;
      CNS_CUDA_CHECK(cudaMalloc(&pk.tc_node_base, total));
      unsigned char* q = pk.tc_node_base;
      for (auto& it : items) { it.dst[0] = q; q += (size_t)it.k * it.n * 2; it.dst[1] = q; q += (size_t)it.k * it.n * 2; }
    }
    for (auto& it : items)
      if ((rc = gemm_tc_pack(it.wt, it.k, it.n, it.dst[0], it.dst[1], s)) != CNS_OK) return rc;
    // (hi, lo) split images for the fp32 mode
    unsigned char** split_dst[] = {&pk.tc2_c0, &pk.tc2_c0_fc, &pk.tc2_g, &pk.tc2_k, &pk.tc2_c1, &pk.tc2_c1_fc, &pk.tc2_out};
    if (!pk.tc2_base) {
      size_t total = 0;
      for (int i = 0; i < 7; ++i) total += (size_t)items[i].k * items[i].n * 2 * 2;
      CNS_CUDA_CHECK(cudaMalloc(&pk.tc2_base, total));
      unsigned char* q = pk.tc2_base;
      for (int i = 0; i < 7; ++i) { *split_dst[i] = q; q += (size_t)items[i].k * items[i].n * 2 * 2; }
    }
    for (int i = 0; i < 7; ++i)
      if ((rc = gemm_tc_split_pack(items[i].wt, items[i].k, items[i].n, *split_dst[i], h->split_f16 != 0, s)) != CNS_OK) return rc;
  }
  if ((rc = bf_pack_vec(h, s)) != CNS_OK) return rc;
  h->weights_loaded = true;
  ++h->weights_version;
  return CNS_OK;
}

}  // namespace cns

using namespace cns;

extern "C" {

const char* cns_version(void) { return "cns_b200 0.1 (sm_100a)"; }
const char* cns_last_error_string(void) { return g_err; }

int cns_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* cns_param_name(int k) { return (k >= 0 && k < P_COUNT) ? kParams[k].name : ""; }
int64_t cns_param_numel(int k) {
  return (k >= 0 && k < P_COUNT) ? (int64_t)kParams[k].rows * kParams[k].cols : -1;
}

int cns_create(cns_handle** out, int device) {
  CNS_REQUIRE(out != nullptr, "out is NULL");
  int n = cns_device_count();
  if (n <= 0) { set_error("no CUDA device visible"); return CNS_ERR_NO_DEVICE; }
  CNS_REQUIRE(device >= 0 && device < n, "bad device ordinal");
  CNS_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CNS_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return CNS_ERR_UNSUPPORTED;
  }
  cns_handle* h = new cns_handle();
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->precision = CNS_PREC_FP32;
  h->tc_tile_n = 32;
  h->no_fused_backbone = getenv("CNS_NO_FUSED_BACKBONE") ? 1 : 0;
  h->no_tc_cluster_pre = getenv("CNS_NO_TC_CLUSTER_PRE") ? 1 : 0;
  h->no_edge_cache = getenv("CNS_NO_EDGE_CACHE") ? 1 : 0;
  h->no_edge_subtract = getenv("CNS_NO_EDGE_SUBTRACT") ? 1 : 0;
  h->no_split_gemm = getenv("CNS_NO_SPLIT_GEMM") ? 1 : 0;
  h->split_f16 = getenv("CNS_SPLIT_BF16") ? 0 : 1;
  h->no_split_encoder = getenv("CNS_NO_SPLIT_ENCODER") ? 1 : 0;
  h->tc_slots = getenv("CNS_TC_SLOTS") ? atoi(getenv("CNS_TC_SLOTS")) : 5;
  h->use_graphs = getenv("CNS_NO_GRAPHS") ? 0 : 1;
  h->sm_count = prop.multiProcessorCount;
  size_t off = 0;
  for (int k = 0; k < P_COUNT; ++k) {
    h->raw_off[k] = off;
    off += (size_t)kParams[k].rows * kParams[k].cols;
    off = (off + 3) & ~(size_t)3;  // keep every tensor 16-byte aligned
  }
  h->raw_off[P_COUNT] = off;
  CNS_CUDA_CHECK(cudaMalloc(&h->raw, off * sizeof(float)));
  CNS_CUDA_CHECK(cudaMemset(h->raw, 0, off * sizeof(float)));
  CNS_CUDA_CHECK(cudaMalloc(&h->packed_base, packed_floats() * sizeof(float)));
  float* p = h->packed_base;
  Packed& pk = h->pk;
  auto take = [&](size_t nfl) { float* r = p; p += nfl; return r; };
  pk.enc_chain = take(5 * 16384);
  pk.lin_dst_t = take(16384);
  pk.tc_chain = take(5 * 16384); pk.tc_lin_dst_t = take(16384); pk.tc_adst_bias = take(128);
  pk.out_t = take(256 * 128);
  pk.c0_wt = take(128 * 256); pk.c0_wp = take(2 * 256); pk.c0_b = take(256);
  pk.g_wt = take(256 * 512);  pk.g_wp = take(2 * 512);  pk.g_b = take(512);
  pk.k_wt = take(256 * 256);  pk.k_wp = take(2 * 256);  pk.k_b = take(256);
  pk.c1_wt = take(128 * 256); pk.c1_wp = take(2 * 256); pk.c1_b = take(256);
  pk.c0_fc_t = take(16384); pk.c1_fc_t = take(16384); pk.vec0_t = take(16384); pk.nrm0_t = take(16384);
  *out = h;
  return CNS_OK;
}

int cns_destroy(cns_handle* h) {
  if (!h) return CNS_OK;
  int prev_device = -1;
  cudaGetDevice(&prev_device);                // the caller's current device is restored on the way out
  cudaSetDevice(h->device);
  if (h->ev[0]) for (int i = 0; i < 2 * 256; ++i) cudaEventDestroy(h->ev[i]);
  cudaFree(h->raw);
  cudaFree(h->packed_base);
  free_bwd_pack(h);
  if (h->tc_w_img) cudaFree(h->tc_w_img);
  if (h->tc_w_split) cudaFree(h->tc_w_split);
  if (h->pk.tc_node_base) cudaFree(h->pk.tc_node_base);
  if (h->pk.tc2_base) cudaFree(h->pk.tc2_base);
  if (h->pk.bf_vec) cudaFree(h->pk.bf_vec);
  graph_cache_clear(h);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->low_stream) { cudaStreamDestroy(h->low_stream); cudaEventDestroy(h->low_fork); cudaEventDestroy(h->low_join); }
  if (h->servo_tcache) cns_target_cache_destroy(h->servo_tcache);
  free(h->servo_tkey);
  if (h->side_stream) { cudaStreamDestroy(h->side_stream); cudaEventDestroy(h->side_fork); cudaEventDestroy(h->side_join); cudaEventDestroy(h->side_join2); }
  if (h->stage_host) cudaFreeHost(h->stage_host);
  if (h->stage_dev) cudaFree(h->stage_dev);
  if (h->servo_hid) cudaFree(h->servo_hid);
  delete h;
  if (prev_device >= 0) cudaSetDevice(prev_device);
  return CNS_OK;
}

static int load_common(cns_handle* h, const float* const* params, const int64_t* numels, int n_tensors,
                       void* stream, cudaMemcpyKind kind) {
  CNS_REQUIRE(h && params && numels, "NULL argument");
  CNS_REQUIRE(n_tensors == P_COUNT, "expected 43 parameter tensors");
  cudaStream_t s = (cudaStream_t)stream;
  CNS_CUDA_CHECK(cudaSetDevice(h->device));
  for (int k = 0; k < P_COUNT; ++k) {
    if (numels[k] != cns_param_numel(k)) {
      set_error("parameter %d (%s): numel %lld, expected %lld", k, kParams[k].name, (long long)numels[k],
                (long long)cns_param_numel(k));
      return CNS_ERR_INVALID_ARGUMENT;
    }
    CNS_CUDA_CHECK(cudaMemcpyAsync(h->raw + h->raw_off[k], params[k], numels[k] * sizeof(float), kind, s));
  }
  return repack(h, s);
}

int cns_load_weights(cns_handle* h, const float* const* params_host, const int64_t* numels, int n_tensors,
                     void* stream) {
  int rc = load_common(h, params_host, numels, n_tensors, stream, cudaMemcpyHostToDevice);
  if (rc == CNS_OK) CNS_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));  // host buffers may go away
  return rc;
}

int cns_load_weights_device(cns_handle* h, const float* const* params_dev, const int64_t* numels,
                            int n_tensors, void* stream) {
  return load_common(h, params_dev, numels, n_tensors, stream, cudaMemcpyDeviceToDevice);
}

// CUDA-event timing of the encoder edge kernel on the launching stream + kernel-launch counter.
int cns_profile_begin(cns_handle* h) {
  CNS_REQUIRE(h != nullptr, "NULL handle");
  CNS_CUDA_CHECK(cudaSetDevice(h->device));
  if (!h->ev[0]) for (int i = 0; i < 2 * 256; ++i) CNS_CUDA_CHECK(cudaEventCreate(&h->ev[i]));
  h->ev_n = 0;
  h->timing = 1;
  graph_cache_begin_timing(h);
  g_launch_count = 0;
  return CNS_OK;
}

int cns_profile_end(cns_handle* h, float* enc_edge_ms_total, int* n_timed, long long* launches) {
  CNS_REQUIRE(h != nullptr, "NULL handle");
  h->timing = 0;
  float total = 0.f;
  for (int i = 0; i < h->ev_n; ++i) {
    float ms = 0.f;
    CNS_CUDA_CHECK(cudaEventSynchronize(h->ev[2 * i + 1]));
    CNS_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev[2 * i], h->ev[2 * i + 1]));
    total += ms;
  }
  int n = h->ev_n;
  {
    float gms = 0.f; int gn = 0;
    int rc = graph_cache_collect_timing(h, &gms, &gn);
    if (rc != CNS_OK) return rc;
    total += gms; n += gn;
  }
  if (enc_edge_ms_total) *enc_edge_ms_total = total;
  if (n_timed) *n_timed = n;
  if (launches) *launches = g_launch_count;
  return CNS_OK;
}

// Tuning knob of the tensor-core encoder kernel: edges per tile (32 -> 4 tiles in flight per SM,
// 64 -> 2).  Also the granularity of the softmax partials.
int cns_set_tc_tile(cns_handle* h, int tile_n) {
  CNS_REQUIRE(h != nullptr, "NULL handle");
  CNS_REQUIRE(tile_n == 32 || tile_n == 64, "tile_n must be 32 or 64");
  h->tc_tile_n = tile_n;
  return CNS_OK;
}

int cns_set_option(cns_handle* h, const char* name, long long value) {
  CNS_REQUIRE(h != nullptr && name != nullptr, "NULL argument");
  if (!strcmp(name, "cuda_graphs")) { h->use_graphs = value != 0; if (!value) graph_cache_clear(h); return CNS_OK; }
  if (!strcmp(name, "target_edge_cache")) { h->no_edge_cache = value == 0; return CNS_OK; }   // read by cns_target_cache_create
  if (!strcmp(name, "split_slices_f16")) {
    CNS_REQUIRE(!h->weights_loaded || (value != 0) == (h->split_f16 != 0), "split_slices_f16: set it before the weights are loaded");
    h->split_f16 = value != 0;
    return CNS_OK;
  }
  if (!strcmp(name, "tc_slots")) { CNS_REQUIRE(value >= 4 && value <= 6, "tc_slots must be 4..6"); h->tc_slots = (int)value; return CNS_OK; }
  set_error("cns_set_option: unknown option '%s'", name);
  return CNS_ERR_INVALID_ARGUMENT;
}

long long cns_get_counter(cns_handle* h, const char* name) {
  if (!h || !name) return -1;
  if (!strcmp(name, "graph_replays")) return h->graph_replays;
  if (!strcmp(name, "graph_captures")) return h->graph_captures;
  if (!strcmp(name, "edge_cache_forwards")) return h->edge_cache_forwards;
  if (!strcmp(name, "kernel_launches")) return g_launch_count;
  return -1;
}

int cns_set_precision(cns_handle* h, int precision) {
  CNS_REQUIRE(h != nullptr, "NULL handle");
  CNS_REQUIRE(precision >= CNS_PREC_FP32 && precision <= CNS_PREC_BF16, "bad precision");
  h->precision = precision;
  return CNS_OK;
}

}  // extern "C"
