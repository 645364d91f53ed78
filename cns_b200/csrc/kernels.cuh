// Internal launch interfaces shared between the translation units of libcns_sm100.so.
#pragma once
#include "common.cuh"

namespace cns {

constexpr int ENC_TILE_E = 128;   // edges per encoder tile
constexpr int ENC_PIECE_E = 64;   // softmax piece granularity (partials are indexed per piece)

// ---- forward.cu: CUDA-graph cache of the forward ------------------------------------------
void graph_cache_clear(cns_handle* h);
void graph_cache_begin_timing(cns_handle* h);
int graph_cache_collect_timing(cns_handle* h, float* ms_total, int* n);

// ---- prep.cu ------------------------------------------------------------------------------
int exclusive_scan_i32(const int* in, int64_t n, int* out, int* scratch, cudaStream_t s);
size_t scan_scratch_bytes(int64_t n);
int sorted_rowptr(const int64_t* dst, int64_t n_edges, int64_t n_rows, int* rowptr, int* err_flag,
                  cudaStream_t s);
size_t csr_workspace_bytes(int64_t n_rows);
int csr_by_dst(const int64_t* ej, const int64_t* ei, int64_t n_edges, int64_t n_rows, int* rowptr, int* src,
               void* workspace, int* err_flag, cudaStream_t s);
int graph_ptr_from_counts(const int64_t* num_clusters, int64_t n_graphs, int* graph_ptr, int* tmp,
                          int* scratch, cudaStream_t s);

int frame_mask(const int64_t* member_order, const int* member_ptr, const uint8_t* visible, int64_t n_clusters,
               int64_t* e01_j, int64_t* e01_i, uint8_t* cluster_mask, int* clu_rowptr, int* cnt_scratch, cudaStream_t s);

constexpr int GRAPH_TILES_MAX = 4096;
int graph_tiles(const int64_t* num_clusters, int64_t n_graphs, int* graph_ptr, int* tile_graph, int* n_tiles, int* err,
                cudaStream_t s);

constexpr int PREP_SMALL_MAX = 4096;
struct PrepSmallArgs {
  int64_t n_clusters, n_graphs, n_e01, n_e11_cur, n_e11_tar;
  const int64_t* num_clusters; const int64_t* ei;
  const int64_t* cur_j; const int64_t* cur_i; const int64_t* tar_j; const int64_t* tar_i;
  int* graph_ptr; int* clu_rowptr; int* cur_rowptr; int* cur_src; int* tar_rowptr; int* tar_src; int* err;
};
int prep_small(const PrepSmallArgs& a, cudaStream_t s);

// ---- encoder.cu ---------------------------------------------------------------------------
struct EncPreArgs {
  int64_t n_clusters;
  const float* x_tar; const float* pos_tar; const int64_t* centers; const int64_t* node_graph;
  const float* w_in; const float* b_in; const float* lin_dst_t;
  const float* a_bias;   // optional (128,): added to a_dst (tensor-core encoder: composed attn_nn.0 biases)
  float* a_dst;      // (C,128)
  float* pos_clu;    // (C,2)
  int* clu_graph;    // (C,)
};
int enc_cluster_pre(const EncPreArgs& a, cudaStream_t s);

struct EncEdgeArgs {
  int64_t n_e01, n_clusters, n_slots;
  const float* x_cur; const float* x_tar; const float* pos_cur; const float* pos_tar;
  const int64_t* ej; const int64_t* ei;
  const float* pos_clu; const float* a_dst; const int* clu_rowptr;
  const float* w_in; const float* b_in; const float* w_p0; const float* b_p0;
  const float* enc_chain; const float* b_p1; const float* b_a0; const float* b_a1;
  float* clu_feat;   // (2, C, 128): relu(ptconv) for cur / tar
  float* partial;    // (2, n_slots, 3, 128)
  // enc_edge_tc only: 0 = both sides, 1 = current side only (the target side comes from a cns_target_cache), 2 = target side
  // only; store_l / store_v (with side_mode 2): fill a target cache - per-keypoint logits (fp32, log2 units) and values (f16)
  // instead of the segment softmax
  int side_mode; float* store_l; void* store_v;
};
int enc_edge_fp32(const EncEdgeArgs& a, cudaStream_t s);
int enc_edge_tc(const EncEdgeArgs& a, int precision, const cns_handle* h, cudaStream_t s);
int tc_pack_weights(cns_handle* h, cudaStream_t s);
int enc_edge_split(const EncEdgeArgs& a, const cns_handle* h, cudaStream_t s);   // fp32-accurate: (hi, lo) operand split on tcgen05 (slices: h->split_f16)
int split_pack_weights(cns_handle* h, cudaStream_t s);
int split_piece_edges(int64_t n_e01, const cns_handle* h);
void tc_chunking(int64_t n_e01, int tile_n, int slots, int sm_count, int sides, int* grid, int* tiles_per_chunk);
int tc_piece_edges(int64_t n_e01, const cns_handle* h, int sides);   // edges per softmax piece (= slot chunk) of enc_edge_tc
// target side from a target cache's per-keypoint logits / values (enc_tar_cached_kernel); out == NULL: fill `stats_out`
// with the per-cluster softmax statistics over all members (cache creation)
struct EncTarCachedArgs {
  const int64_t* ej; int64_t n_e01; const int* clu_rowptr; int64_t n_clusters, n_nodes;
  const float* lc; const void* vc; const float* b_p1;
  const int* all_src; const int* all_rowptr; const float* all_stats;
  uint8_t* visible;          // (n_nodes,) scratch
  float* out; float* stats_out;
};
int enc_tar_cached(const EncTarCachedArgs& a, cudaStream_t s);

struct EncFinishArgs {
  int64_t n_clusters, n_slots;
  int piece_e;       // edges per softmax piece used by the edge kernel that ran
  const int* clu_rowptr; const float* partial; float* clu_feat;
  const float* out_t; const float* b_out; const uint8_t* cluster_mask;
  float* x_clu;      // (C,128)
  float* cat_out;    // optional (C,256): write [cur | tar - cur] and skip out_enc (done by gemm_tc)
  int fix_only;      // only repair clu_feat (merge cut clusters, zero empty ones): out_enc runs in the fused backbone
  int cur_only;      // the target-side features are final (enc_tar_cached): merge / zero the current side only
};
int enc_finish(const EncFinishArgs& a, cudaStream_t s);

// ---- backbone.cu --------------------------------------------------------------------------
// Y = act([X1 | X2] @ wt + pos @ wp + bias + residual);  wt (k1+k2, n_out) row-major
struct LinearArgs {
  int64_t m; int k1, k2, n_out;
  const float* x1; const float* x2; const float* wt; const float* bias;
  const float* pos; const float* wp;       // optional (m,2) / (2,n_out)
  const float* residual;                   // optional (m,n_out)
  int relu;
  const float* relu_gate;                  // optional (m,n_out): y = 0 where relu_gate <= 0 (ReLU backward fused into dX)
  float* y;
};
int linear(const LinearArgs& a, cudaStream_t s);

// ---- gemm_split.cu: the same contract for 128 x 128 layers on tcgen05 with fp32 accuracy (3-way bf16 operand split)
constexpr size_t GEMM_SPLIT_IMG_BYTES = 3 * 128 * 128 * 2;
int gemm_split_pack(const float* wt, unsigned char* img, cudaStream_t s);
int gemm_split(const LinearArgs& a, const unsigned char* w_img, cudaStream_t s);
int dw_split(int64_t m, const float* a, int lda, const float* b, int ldb, float* partial, float* colsum_partial, int max_slabs,
             int* n_slabs, cudaStream_t s);

// ---- gemm_tc.cu: same contract as `linear` on tcgen05 (16-bit operands), plus an optional row mask
struct GemmTcArgs {
  int64_t m; int k1, k2, n_out;
  const float* x1; const float* x2;
  const unsigned char* w_img;              // swizzled 16-bit weight image (gemm_tc_pack)
  int k_img;                               // K the image was packed with (>= k1+k2; 0 => k1+k2): a GEMM over
                                           // the first k1+k2 input columns only reads a prefix of each chunk
  const float* bias; const float* pos; const float* wp; const float* residual;
  const uint8_t* row_mask;                 // optional (m,): rows with 0 are written as zeros
  int relu;
  float* y;
  // optional A-row producer instead of x1 (k1 = 128): row m = relu(in_enc(ctr_xy[centers[m]])), the per-cluster centre
  // feature of the encoder (graph_vs.py:207-208); also emits pos_out[m] = ctr_pos[centers[m]] and the cluster's graph id
  const int64_t* centers; const float* ctr_xy; const float* ctr_pos; const float* w_in; const float* b_in;
  const int64_t* node_graph; float* pos_out; int* clu_graph;
};
int gemm_tc(const GemmTcArgs& a, bool bf16, cudaStream_t s);
int gemm_tc_pack(const float* wt, int k_dim, int n_out, unsigned char* img_bf16, unsigned char* img_f16, cudaStream_t s);
// the same GEMM at fp32 accuracy: (hi, lo) operand split, three products (gemm_tc_split.cu); w_img from gemm_tc_split_pack
// with the same slice format (f16_slices: two f16 slices instead of two bf16 ones, tc_common.cuh split_pair)
int gemm_tc_split(const GemmTcArgs& a, bool f16_slices, cudaStream_t s);
int gemm_tc_split_pack(const float* wt, int k_dim, int n_out, unsigned char* img, bool f16_slices, cudaStream_t s);

enum EdgeConvMode { EC_PLAIN = 0, EC_GATES = 1, EC_CANDIDATE = 2 };
struct EdgeConvArgs {
  int64_t n_rows; int f_out; int mode;
  const float* pq; const int* rowptr; const int* src;
  float* out;              // EC_PLAIN: (n, f)
  const float* h;          // EC_GATES / EC_CANDIDATE: previous hidden (n,128) or NULL (zeros)
  float* hr; float* u;     // EC_GATES: h * r (n,128), update gate (n,128)
  const float* u_in;       // EC_CANDIDATE
  float* h_out;            // EC_CANDIDATE: (1-u) h + u tanh(.)
  int* argmax;             // optional (n, f): winning source per channel (training)
  float* act;              // optional: post-activation values kept for backward
};
int edgeconv_max(const EdgeConvArgs& a, cudaStream_t s);

// ---- dense_l1.cu: cluster graphs that are complete digraphs (CNS_BATCH_L1_*_COMPLETE) -----------------
struct L1CompleteArgs {
  int64_t n_graphs; const int* graph_ptr;
  const uint8_t* cur_mask;                        // non-NULL: the current-side list is claimed complete over mask != 0
  const int64_t* cur_j; const int64_t* cur_i; int64_t n_e11_cur;   // NULL: nothing to verify (graph implied by the mask)
  const int64_t* tar_j; const int64_t* tar_i; int64_t n_e11_tar;   // NULL: target side not claimed / implied
  int* alive_idx; int* alive_cnt;                 // scratch (C,), (B,)
  int* err;
};
int l1_complete_verify(const L1CompleteArgs& a, cudaStream_t s);

struct EdgeConvDenseArgs {
  int64_t n_graphs; int f_out; int mode;
  const float* pq; const int* graph_ptr;
  const uint8_t* mask;     // node set of every graph: rows with mask != 0, or every row when NULL
  float* out;              // EC_PLAIN without GraphNorm: (n, f)
  const float* gn_weight; const float* gn_bias; const float* gn_mean_scale; float* gn_out;   // EC_PLAIN + GraphNorm + ReLU
  const float* h; float* hr; float* u; const float* u_in; float* h_out;                      // as EdgeConvArgs
};
int edgeconv_dense(const EdgeConvDenseArgs& a, cudaStream_t s);

// ---- backbone_fused.cu: PERConv -> GRU -> PERConv over whole-graph row tiles (complete graphs, <= 128 clusters each)
struct BackboneFusedArgs {
  int64_t n_clusters; int n_graphs;
  const int* graph_ptr;
  const int* tile_graph; const int* n_tiles;      // scratch: (B + 2,), (1,) filled by the tile packer
  const float* clu_feat;                          // (2, C, 128) encoder features: current side, target side
  const float* hidden_in; const float* pos_clu; const uint8_t* mask;
  const unsigned char* w_out; const float* out_b;
  const unsigned char* w_c0; const unsigned char* w_c0fc; const unsigned char* w_g; const unsigned char* w_k;
  const unsigned char* w_c1; const unsigned char* w_c1fc;      // gemm_tc weight images
  const float* c0_b; const float* c0_wp; const float* g_b; const float* g_wp; const float* k_b; const float* k_wp;
  const float* c1_b; const float* c1_wp; const float* c0fc_b; const float* c1fc_b;
  const float* gn0_w; const float* gn0_b; const float* gn0_ms; const float* gn1_w; const float* gn1_b; const float* gn1_ms;
  const float* vec_tab;                           // (3, 1664) per-column epilogue vectors (bf_pack_vec)
  float* hidden_out; float* x2;
};
int bf_pack_vec(cns_handle* h, cudaStream_t s);
int backbone_fused(const BackboneFusedArgs& a, bool bf16, cudaStream_t s);

struct GraphNormArgs {
  int64_t n_graphs; const int* graph_ptr; const float* x;
  const float* weight; const float* bias; const float* mean_scale;
  float* y; float* stats;  // stats optional (B, 2, 128): mean, rstd (training)
};
int graphnorm_relu(const GraphNormArgs& a, cudaStream_t s);

struct PoolHeadArgs {
  int64_t n_graphs; const int* graph_ptr; const float* x; const uint8_t* cluster_mask;
  const float* pool_w; const float* pool_b;
  const float* vec0_t; const float* vec0_b; const float* vec1_w; const float* vec1_b;
  const float* nrm0_t; const float* nrm0_b; const float* nrm1_w; const float* nrm1_b;
  const float* tPo_norm;
  float* gate;      // scratch (C,)
  float* vel_si_vec; float* vel_si_norm; float* vel;  // vel optional
  float* x_scene; float* hid;  // optional saves (B,128), (B,2,128)
};
int pool_head(const PoolHeadArgs& a, cudaStream_t s);

}  // namespace cns
