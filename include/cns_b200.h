/*
 * cns_b200.h - C ABI of libcns_sm100.so: the B200 (sm_100a) implementation of the
 * CNS hot path (hhcaz/CNS: cns/midend graph construction + cns/models GraphVS).
 *
 * The reference is 100 % Python and has no FFI of its own, so
 * every entry point below names the reference *Python* interface it replaces
 * (file:line relative to hhcaz/CNS).  A reference maintainer binds these with ctypes;
 * the stub is shown in INTEGRATION.md and implemented in cns_b200/_lib.py.
 *
 * Conventions
 *  - plain C: pointers + sizes only, no torch / C++ types.
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in
 *    `_host`.  fp32 tensors are row-major and contiguous; index tensors are int64
 *    exactly as the reference's GraphData stores them; masks are 1 byte per element.
 *  - every function returns 0 (CNS_OK) or a negative cns_status.  Nothing throws,
 *    exits or allocates device memory behind the caller's back: scratch comes from a
 *    caller-provided workspace whose size a `*_workspace_bytes` query returns.
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work
 *    is enqueued asynchronously on it; a handle is used by one host thread at a time.
 *  - kernels are deterministic: segment-sorted aggregation, no floating-point atomics.
 */
#ifndef CNS_B200_H
#define CNS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum cns_status {
  CNS_OK = 0,
  CNS_ERR_INVALID_ARGUMENT = -1,
  CNS_ERR_WORKSPACE_TOO_SMALL = -2,
  CNS_ERR_CUDA = -3,
  CNS_ERR_UNSORTED_EDGES = -4,   /* l0_to_l1 edges must be sorted by cluster (reference order) */
  CNS_ERR_NO_DEVICE = -5,
  CNS_ERR_UNSUPPORTED = -6
} cns_status;

#define CNS_HIDDEN_DIM 128        /* hidden_dim of checkpoints/cns.pth (graph_vs.py:274) */
#define CNS_FEAT_DIM 2            /* input_dim  (train_cns.py:56-61)                      */
#define CNS_POS_DIM 2             /* pos_dim                                              */
#define CNS_NUM_PARAM_TENSORS 43  /* state-dict entries, SURVEY.md section 8b             */

/* arithmetic mode of the batched MLP (GEMM) kernels */
typedef enum cns_precision {
  CNS_PREC_FP32 = 0,     /* fp32 FFMA everywhere (rel 1e-4 vs the fp64 oracle)         */
  CNS_PREC_F16 = 1,      /* tcgen05 kind::f16 fp16 inputs, fp32 accumulate (rel ~2e-3)   */
  CNS_PREC_BF16 = 2      /* tcgen05 kind::f16 bf16 inputs, fp32 accumulate (rel 2e-2)   */
} cns_precision;

typedef struct cns_handle cns_handle;   /* packed weights + per-device constants */

/* ---- library / handle ------------------------------------------------------------ */
const char* cns_version(void);
const char* cns_last_error_string(void);            /* thread-local, never NULL */
int cns_device_count(void);

/* Replaces GraphVS.__init__ + load_state_dict (graph_vs.py:273-279, controller.py:18-19).
 * `params_host` : 43 host pointers to fp32 tensors in the canonical state-dict order
 *                 returned by cns_param_name(k); `numels[k]` must equal cns_param_numel(k).
 * Packs (transposes / folds) the weights into device-resident GEMM operands. */
int cns_create(cns_handle** out, int device);
int cns_destroy(cns_handle* h);
int cns_load_weights(cns_handle* h, const float* const* params_host, const int64_t* numels,
                     int n_tensors, void* stream);
/* Same, from device-resident parameters (training: re-pack after every optimizer step). */
int cns_load_weights_device(cns_handle* h, const float* const* params_dev, const int64_t* numels,
                            int n_tensors, void* stream);
const char* cns_param_name(int k);      /* e.g. "encoder.in_enc.fcs.0.weight" */
int64_t cns_param_numel(int k);
int cns_set_precision(cns_handle* h, int precision /* cns_precision */);
int cns_set_tc_tile(cns_handle* h, int tile_n /* 32 or 64: edges per tensor-core tile */);
/* Tuning switches (no reference counterpart): "cuda_graphs" 0/1 - replay repeated cns_graphvs_forward calls (same sizes,
 * pointers and mode) as one CUDA-graph launch, default 1 (CNS_NO_GRAPHS=1 in the environment turns it off);
 * "tc_slots" 4..6 - edge tiles in flight per SM; "target_edge_cache" 0/1 - whether cns_target_cache_create keeps the
 * per-keypoint target-side logits / values (default 1; caches created while it is 0 recompute the target side per frame);
 * "split_slices_f16" 0/1 - operand slices of the fp32-tolerance kernels (CNS_PREC_FP32 and the training forward): 1 (default)
 * = two f16 slices per operand (22 mantissa bits; activations and weights must stay below 65504 in magnitude or the outputs
 * turn NaN), 0 = two bf16 slices (16 mantissa bits at fp32's exponent range: ~1e-5 instead of ~1e-6 of fp64 on the outputs;
 * CNS_SPLIT_BF16=1 in the environment); set it before the weights are loaded. */
int cns_set_option(cns_handle* h, const char* name, long long value);
/* "graph_replays", "graph_captures", "edge_cache_forwards" (forwards enqueued with the target side taken from a
 * cns_target_cache's per-keypoint cache), "kernel_launches" (library-wide host-side count); -1 for an unknown name */
long long cns_get_counter(cns_handle* h, const char* name);

/* Measurement hooks: between begin and end every cns_graphvs_forward brackets its dominant
 * kernel (the encoder edge kernel) with CUDA events on the launching stream (first 256
 * forwards) and the library counts the kernels it launches.  end() waits for the events. */
int cns_profile_begin(cns_handle* h);
int cns_profile_end(cns_handle* h, float* enc_edge_ms_total, int* n_timed, long long* kernel_launches);

/* ---- one batch of servo graphs ------------------------------------------------------
 * Field-for-field the tensors GraphVS.forward unpacks from a GraphData/Batch
 * (graph_vs.py:285-307; field list graph_gen.py:257-275). */
typedef struct cns_batch {
  int64_t n_nodes;       /* sum N  : keypoints                                    */
  int64_t n_clusters;    /* sum C  : clusters (rows of hidden)                    */
  int64_t n_graphs;      /* B                                                     */
  int64_t n_e01;         /* len(l0_to_l1_edge_index_*_cur)                        */
  int64_t n_e11_cur;     /* l1_dense_edge_index_cur.shape[1]                      */
  int64_t n_e11_tar;     /* l1_dense_edge_index_tar.shape[1]                      */
  const float* x_cur;    /* (N,2) */
  const float* x_tar;    /* (N,2) */
  const float* pos_cur;  /* (N,2) */
  const float* pos_tar;  /* (N,2) */
  const int64_t* l0_to_l1_j;      /* (E01,) source keypoint                       */
  const int64_t* l0_to_l1_i;      /* (E01,) destination cluster, non-decreasing   */
  const int64_t* l1_edge_cur;     /* (2,E11c) row 0 = source j, row 1 = dest i    */
  const int64_t* l1_edge_tar;     /* (2,E11t)                                     */
  int64_t l1_cur_stride;          /* elements between the two rows; 0 => n_e11_cur */
  int64_t l1_tar_stride;          /* 0 => n_e11_tar                                */
  const uint8_t* cluster_mask;    /* (C,)  bool                                   */
  const int64_t* cluster_centers; /* (C,)  keypoint index of each cluster centre  */
  const int64_t* node_graph;      /* (N,)  `batch` vector, or NULL => all zero    */
  const int64_t* num_clusters;    /* (B,)  clusters per graph                     */
  const float* tPo_norm;          /* (B,)  distance scale, or NULL => 1           */
  int64_t flags;                  /* CNS_BATCH_* hints, 0 = none                   */
  int64_t max_clusters_per_graph; /* upper bound on num_clusters[g] known to the host, 0 = unknown.  With both
                                   * COMPLETE hints, a bound <= 128 and a 16-bit precision mode the cluster-level
                                   * backbone runs as one fused kernel over whole-graph row tiles (verified on the
                                   * device like the hints)                                                   */
  const struct cns_target_cache* target_cache;   /* optional, see cns_target_cache_create; NULL = recompute */
  const int32_t* clu_rowptr;      /* optional (C+1,): segment pointers of l0_to_l1_i (cns_frame_mask); NULL = derived */
} cns_batch;

/* Structure hints about the cluster-level edge lists.  What GraphGenerator builds (graph_gen.py:31-36 via
 * :199-228) is, per graph, the complete digraph WITH self loops over a node set, in source-major order:
 * all clusters for l1_dense_edge_index_tar, the clusters with cluster_mask != 0 for ..._cur.  With the
 * hint set the forward skips the CSR build and replaces every max-aggregation over neighbours by one
 * column max per graph.  The claim is verified on the device on every call; a violation is reported by
 * cns_forward_status (CNS_ERR_INVALID_ARGUMENT) and the outputs of that call are undefined.  Ignored when
 * `save` != NULL (training keeps the general path). */
#define CNS_BATCH_L1_CUR_COMPLETE 1
#define CNS_BATCH_L1_TAR_COMPLETE 2

/* Everything the forward derives from the TARGET side of a batch alone - the reference builds the target graph
 * once per target image and only re-masks it per frame (Midend.get_graph_data, corr2graph.py:11-22;
 * GraphGenerator.__init__, graph_gen.py:133-190): the per-cluster attention term of the encoder (in_enc + lin_dst
 * of the cluster centres, graph_vs.py:207-208), cluster positions, graph pointers and the row tiles of the fused
 * backbone.  Built from a batch's x_tar / pos_tar / cluster_centers / num_clusters; valid for every later batch
 * with the same target tensors while the handle keeps its weights and precision (checked: a stale cache is
 * rejected with CNS_ERR_INVALID_ARGUMENT).
 * When `b` also lists EVERY target keypoint with its cluster (n_e01 == n_nodes; l0_to_l1_j = keypoint index, l0_to_l1_i =
 * cluster index, any order - belong_cluster_index of graph_gen.py:266) and the precision mode is f16, the cache also keeps
 * the target side of the keypoint -> cluster attention per keypoint (logits fp32 + values f16, 768 bytes per keypoint): a
 * forward with such a cache runs the attention chain on the current side only and a streaming segment softmax over the
 * keypoints that are visible in that frame for the target side (graph_vs.py:195-214: the target-side messages do not
 * depend on the current frame, only which of them take part in the softmax). */
typedef struct cns_target_cache cns_target_cache;
int cns_target_cache_create(cns_handle* h, const cns_batch* b, void* stream, cns_target_cache** out);
int cns_target_cache_destroy(cns_target_cache* c);

/* Replaces GraphVS.forward (graph_vs.py:284-322) + GraphVS.postprocess
 * (functions.py:26-38).  hidden_in may be NULL (== init_hidden zeros, graph_vs.py:315).
 * Outputs: vel_si_vec (B,6), vel_si_norm (B,1), hidden_out (C,128), vel (B,6) or NULL.
 * `save` != NULL (training) keeps the activations the backward pass needs in
 * `save` (size cns_forward_save_bytes). */
size_t cns_forward_workspace_bytes(const cns_batch* b);
size_t cns_forward_save_bytes(const cns_batch* b);
int cns_graphvs_forward(cns_handle* h, const cns_batch* b, const float* hidden_in,
                        float* vel_si_vec, float* vel_si_norm, float* hidden_out, float* vel,
                        void* workspace, size_t workspace_bytes,
                        void* save, size_t save_bytes, void* stream);

/* Backward of cns_graphvs_forward (training: trainer.py:75-100, train_gvs_short_seq.py:24-53).
 * `save` is the buffer the forward filled; hidden_in / hidden_out are the forward's input / output
 * hidden states; d_* are upstream gradients (NULL == zero).  Parameter gradients are ACCUMULATED
 * (+=) into grad_flat, a flat fp32 buffer of cns_param_offset(h, 43) floats where parameter k
 * starts at cns_param_offset(h, k).  d_hidden_in (C,128) is overwritten (may be NULL).
 * fp32 arithmetic; deterministic (fixed-order reductions, no floating-point atomics). */
size_t cns_backward_workspace_bytes(const cns_batch* b);
int64_t cns_param_offset(const cns_handle* h, int k);
int cns_graphvs_backward(cns_handle* h, const cns_batch* b, const float* hidden_in, const float* hidden_out,
                         const float* d_vel_si_vec, const float* d_vel_si_norm, const float* d_hidden_out,
                         float* d_hidden_in, float* grad_flat, const void* save, size_t save_bytes,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Loss of GraphVS.objectives (functions.py:41-65, gt_si=True, weight=1) and, when d_vec / d_norm are given,
 * grad_scale * d loss / d (vel_si_vec, vel_si_norm), in one launch.  loss_out3 (device): dir, norm, total. */
int cns_objectives(const float* vel_si_vec, const float* vel_si_norm, const float* vel_si_gt, int64_t n_graphs,
                   float grad_scale, float* loss_out3, float* d_vec, float* d_norm, void* stream);

/* GraphVS.postprocess (functions.py:26-38, post_scale=True) on device predictions: vel (B,6) = vel_si_vec /
 * (|vel_si_vec| + 1e-7) * (1 + ELU(vel_si_norm)), translational part x tPo_norm.  vel_si_norm NULL: vel = vel_si_vec
 * (regress_norm=False); tPo_norm NULL: no distance scaling (post_scale=False). */
int cns_postprocess(const float* vel_si_vec, const float* vel_si_norm, const float* tPo_norm, int64_t n_graphs,
                    float* vel, void* stream);

/* clip_grad_norm_(max_norm) + AdamW over flat parameter / gradient / moment buffers (trainer.py:96-100,
 * train_cns.py:65-69); decay_mask (n) u8 selects the decayed group.  `grad_div` pre-divides the gradients
 * (world size after a summing all-reduce).  grad_norm_out: optional device float (pre-clip norm). */
size_t cns_clip_adamw_workspace_bytes(void);
int cns_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                   const unsigned char* decay_mask, int64_t n, float max_norm, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int64_t step, float grad_div, float* grad_norm_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Error word left in `workspace` by the index-preparation kernels of the last forward on
 * `stream` (unsorted l0_to_l1 edges, out-of-range indices).  Synchronises `stream`. */
int cns_forward_status(const void* workspace, void* stream);

/* Same call with every cns_batch pointer (and hidden_in / outputs) in HOST memory: the
 * library stages inputs through its own pinned buffers, runs the forward and copies the
 * results back before returning (synchronous).  This is what a caller without device
 * tensors binds - GraphVSController.__call__ incl. data.to(device) and .cpu()
 * (controller.py:23-38). */
int cns_graphvs_forward_host(cns_handle* h, const cns_batch* b_host, const float* hidden_in_host,
                             float* vel_si_vec_host, float* vel_si_norm_host,
                             float* hidden_out_host, float* vel_host);

/* One servo frame from HOST memory with the recurrent hidden state kept on the device inside the
 * handle - GraphVSController.__call__ (controller.py:23-38) including its self.hidden bookkeeping:
 * reset_hidden != 0 (new scene) or a changed cluster count start from the zero state.  Inputs are
 * packed into one pinned block (one H2D copy), outputs come back in one D2H copy.  Synchronous.
 * vel_host (B,6) is the post-processed velocity; the raw heads are optional. */
int cns_servo_step_host(cns_handle* h, const cns_batch* b_host, int reset_hidden, float* vel_host,
                        float* vel_si_vec_host, float* vel_si_norm_host);

/* ---- per-op entry points (each is one stage of the forward; parity-tested alone) ---- */

/* CSR by destination of an int64 (2,E) edge list (PyG flow source->target,
 * graph_vs.py:41-43): rowptr (n_dst+1) int32, src (E) int32.  Needs (n_dst+1)*4 bytes
 * of workspace.  Row order inside a row is unspecified (max/sum of a set). */
int cns_csr_by_dst(const int64_t* edge_index, int64_t n_edges, int64_t n_dst,
                   int32_t* rowptr, int32_t* src, void* workspace, size_t workspace_bytes,
                   void* stream);

/* PointEdgeConv aggregation in factorised form (graph_vs.py:46-111):
 * out[i,:] = deg(i) ? P[i,:] + max_{j in N(i)} Q[j,:] : 0, PQ = (n, 2*f) rows [P | Q]. */
int cns_edgeconv_max(const float* pq, const int32_t* rowptr, const int32_t* src,
                     int64_t n_rows, int f_out, float* out, void* stream);

/* PyG GraphNorm + ReLU (graph_vs.py:132,142-143): per graph g over rows
 * [graph_ptr[g], graph_ptr[g+1]) ; y = relu(w * (x - mean*ms) / sqrt(var + 1e-5) + b). */
int cns_graphnorm_relu(const float* x, const int32_t* graph_ptr, int64_t n_graphs,
                       const float* weight, const float* bias, const float* mean_scale,
                       float* y, void* stream);

/* Y = act(X1 @ W[:, :k1].T + X2 @ W[:, k1:].T + bias (+ residual)); W given TRANSPOSED
 * (k, n_out) row-major.  Batched node MLP (graph_vs.py:16-31). */
int cns_linear(const float* x1, int k1, const float* x2, int k2, const float* wt,
               const float* bias, const float* residual, int relu, int64_t m, int n_out,
               float* y, void* stream);

/* cns_linear for 128 -> 128 layers on the tensor cores at fp32 accuracy: both operands are split into three bf16
 * slices and the six significant slice products accumulate in fp32 (gemm_split.cu).  Used by the backward pass for the
 * edge-level GEMMs (the recomputed PointTransformerConv chain and dX = dY W, graph_vs.py:195-214 under autograd).
 * y = gate(act(x wt + bias + residual)): x (m,128), wt (128,128) row-major (k, n), bias (128) / residual (m,128) /
 * relu_gate (m,128: y = 0 where relu_gate <= 0) optional.  y may be the residual buffer (in place); it must not
 * overlap x or relu_gate.  workspace: cns_linear_split_workspace_bytes(). */
size_t cns_linear_split_workspace_bytes(void);
int cns_linear_split(const float* x, int64_t m, const float* wt, const float* bias, const float* residual,
                     const float* relu_gate, int relu, float* y, void* workspace, size_t workspace_bytes, void* stream);

/* ---- graph construction (cns/midend/graph_gen.py) ----------------------------------- */

/* Per-frame sub-graph masking for a batch of scenes (graph_gen.py:199-228): given the
 * static two-level graph of every scene and a per-keypoint visibility mask, emits the
 * filtered, order-preserving edge lists and masks of GraphGenerator._get_graph.
 *   in : belong (N,) int64 global cluster id of each keypoint (batched offsets applied),
 *        member_order (N,) int64 keypoints in cluster-major order (= static l0_to_l1 j),
 *        cluster_size (C,) int64, cluster_graph_ptr (B+1,) int64, visible (N,) u8
 *   out: e01_j, e01_i (<=N) ; e11_cur (2, <= sum C_g^2) written as two rows of stride
 *        e11_capacity (pass it as cns_batch.l1_cur_stride) ; cluster_mask (C,) ;
 *        counts_out[0]=E01, [1]=E11cur (device int64).
 *        e11_cur may be NULL: the current-side cluster graph is then left implied by cluster_mask
 *        (complete digraph over the alive clusters, what CNS_BATCH_L1_CUR_COMPLETE describes) and
 *        counts_out[1] is 0.  E01 equals the number of visible keypoints, which the host already knows.
 */
size_t cns_graph_mask_workspace_bytes(int64_t n_nodes, int64_t n_clusters, int64_t n_graphs);
int cns_graph_mask(const int64_t* belong, const int64_t* member_order, const int64_t* cluster_size,
                   const int64_t* cluster_graph_ptr, const uint8_t* visible,
                   int64_t n_nodes, int64_t n_clusters, int64_t n_graphs,
                   int64_t* e01_j, int64_t* e01_i, int64_t* e11_cur, int64_t e11_capacity,
                   uint8_t* cluster_mask, int64_t* counts_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* The same masking in cluster-major form, for callers that keep the static per-target arrays around
 * (BatchedServo): member_ptr (C+1,) int32 = prefix sums of cluster_size.  Three launches; additionally returns the
 * cluster rowptr of the filtered edge list (C+1,) int32, which can be handed to the forward as
 * cns_batch.clu_rowptr.  workspace: (C+1) int32. */
int cns_frame_mask(const int64_t* member_order, const int32_t* member_ptr, const uint8_t* visible, int64_t n_nodes,
                   int64_t n_clusters, int64_t* e01_j, int64_t* e01_i, uint8_t* cluster_mask, int32_t* clu_rowptr,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Weight / bias gradient of one 128 -> 128 Linear layer (what autograd computes for nn.Linear inside GraphVS,
 * graph_vs.py:16-31): dw[o][i] (+)= sum_m dy[m][o] x[m][i], dbias[o] (+)= sum_m dy[m][o] (dbias may be NULL), rows
 * with strides ld_dy / ld_x floats.  Split-M partial products + a fixed-order reduction (deterministic).  With
 * tensor_cores != 0 and m >= 4096 the partial products run on tcgen05 with both operands split into three bf16
 * slices (fp32 accuracy, gemm_split.cu); otherwise on the fp32 FFMA kernel.  workspace:
 * cns_weight_grad_workspace_bytes(). */
size_t cns_weight_grad_workspace_bytes(void);
int cns_weight_grad(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, int64_t m, float* dw, float* dbias,
                    int accumulate, int tensor_cores, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the steps either side of the graph path in the servo loop --------------------------------
 *
 * cns_align_matches replaces Classic._align_cur_to_tar (cns/frontend/classic.py:142-148) followed by
 * CameraIntrinsic.pixel_to_norm_camera_plane (cns/utils/perception.py:66-68, called by Midend.get_graph_data,
 * cns/midend/corr2graph.py:11-22) and the float32 cast of GraphGenerator.get_data (graph_gen.py:257-263):
 *   cur_pos (n_cur,2) f32 pixel coordinates of the detected keypoints; matches (n_matches,2) int64 rows
 *   [target index, current index] (a later row overrides an earlier one with the same target index, like numpy's
 *   fancy assignment) -> cur_xy (n_tar,2) f32 normalized coordinates aligned with the target keypoints (unmatched
 *   rows: pixel (0,0) normalized, as the reference) and valid (n_tar,) u8.  Bit-exact.  Out-of-range indices are
 *   skipped and reported by the int at workspace[n_tar] != 0 after the call.  workspace: (n_tar + 1) int32. */
int cns_align_matches(const float* cur_pos, int64_t n_cur, const int64_t* matches, int64_t n_matches, int64_t n_tar,
                      double fx, double fy, double cx, double cy, float* cur_xy, uint8_t* valid,
                      void* workspace, size_t workspace_bytes, void* stream);

/* cns_pixel_stop_error replaces PixelStopPolicy.calculate_error (cns/benchmark/stop_policy.py:86-126) for
 * n_graphs scenes at once (one CTA per scene; node_ptr (n_graphs+1,) int32 keypoint offsets, NULL = one scene):
 * witness counts / decayed weights (device state, (n_nodes,) int32 / f32, zero-initialised by the caller) are
 * updated, the weighted keypoint errors are cut at their 90th percentile and the weighted mean of the rest is
 * written to err_out (n_graphs,) f32.  Every float32 operation follows numpy's order (np.linalg.norm,
 * np.percentile "linear" with a float32 virtual index, pairwise np.sum), so the result is bit-identical with the
 * reference running on numpy 2.x.  reset_flags (n_graphs,) u8 or NULL: scenes whose state is cleared before this
 * frame (PixelStopPolicy.reset).  A scene without a visible keypoint gets NaN (the reference raises).
 * n_valid_out (n_graphs,) int32 or NULL.  workspace: cns_pixel_stop_workspace_bytes(n_nodes). */
size_t cns_pixel_stop_workspace_bytes(int64_t n_nodes);
int cns_pixel_stop_error(const float* pos_cur, const float* pos_tar, const uint8_t* node_mask, const int32_t* node_ptr,
                         const uint8_t* reset_flags, int64_t n_nodes, int64_t n_graphs, float decay, float maximum,
                         float* weight, int32_t* witness, float* err_out, int32_t* n_valid_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Affinity Propagation message passing (the clustering behind cluster_points, graph_gen.py:104-106 ->
 * sklearn AffinityPropagation(damping=0.9)) for a batch of G problems, bit-exact in fp64 with
 * sklearn's loop.  S (sum n_g^2) is the similarity matrix as sklearn prepares it (median preference on
 * the diagonal, noise added) - device; pt_off / mat_off (G+1) int64 prefix sums of n_g / n_g^2 - device;
 * row_graph (sum n) int32 - device.  Outputs: E_out (sum n) u8 exemplar indicator and n_iter_out (G)
 * on the device, converged_host (G) on the HOST (1 converged, 2 hit max_iter).  Synchronises. */
size_t cns_ap_workspace_bytes(int64_t n_rows_total, int64_t n_mat_total, int64_t n_graphs, int conv_iter);
int cns_affinity_propagation(const double* S, const int64_t* pt_off, const int64_t* mat_off,
                             const int32_t* row_graph, int64_t n_graphs, int64_t n_rows_total,
                             int64_t n_mat_total, double damping, int max_iter, int conv_iter,
                             unsigned char* E_out, int32_t* n_iter_out, int32_t* converged_host,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- device graph builder around the Affinity Propagation loop (cns/midend/graph_gen.py:104-196) -----------------
 * Batched over G target point sets: points (sum n, 2) f64, pt_off / mat_off (G+1) int64 prefix sums of n_g / n_g^2,
 * row_graph (sum n) int32 - all on the device.  Every fp64 result is bit-identical with the numpy / scikit-learn
 * expression it replaces (same operation order, no contraction).
 *
 * cns_ap_similarity   : S = -euclidean_distances(X, squared=True) as sklearn.cluster.AffinityPropagation.fit builds it
 *                       (_affinity_propagation.py:517).
 * cns_segment_median  : np.median of each fp64 segment values[seg_off[s]:seg_off[s+1]] (exact radix select; the AP
 *                       preference is median(S), _affinity_propagation.py:526).  max_segment = longest segment.
 * cns_ap_prepare      : S.flat[::n+1] = preference; S += (eps * S + tiny * 100) * noise (_affinity_propagation.py:68-78);
 *                       noise (sum n^2) f64 = the HOST RandomState's standard_normal((n, n)) draws, in order.
 * cns_ap_labels       : exemplar indicator E (from cns_affinity_propagation) -> labels: assignment, exemplar refinement,
 *                       re-assignment, gapless relabelling (_affinity_propagation.py:147-166).  labels_out (sum n) int32
 *                       (-1 everywhere when no exemplar emerged), exemplars_out (sum n) int32: the K_g cluster-centre
 *                       point indices of problem g ascending at [pt_off[g], pt_off[g] + K_g), n_clusters_out (G).
 *                       workspace: 2 * align256(4 * sum n) bytes.
 * cns_graph_layout    : GraphGenerator.__init__ after clustering (graph_gen.py:146-196) from gapless labels:
 *                       member_order (sum n) int32 = l0_to_l1_edge_index[0] (cluster-major, ascending point index, local
 *                       to the problem), cluster_size / centers (sum K) int32 at clu_off (G+1) int64 prefix of K_g:
 *                       centre = member nearest to the per-cluster coordinate median, first minimum. */
int cns_ap_similarity(const double* points, const int64_t* pt_off, const int64_t* mat_off, const int32_t* row_graph,
                      int64_t n_rows_total, double* S_out, void* stream);
size_t cns_segment_median_workspace_bytes(int64_t n_segments);
int cns_segment_median(const double* values, const int64_t* seg_off, int64_t n_segments, int64_t max_segment,
                       double* median_out, void* workspace, size_t workspace_bytes, void* stream);
int cns_ap_prepare(double* S, const double* noise, const double* preference, const int64_t* pt_off,
                   const int64_t* mat_off, const int32_t* row_graph, int64_t n_rows_total, void* stream);
int cns_ap_labels(const double* S, const unsigned char* E, const int64_t* pt_off, const int64_t* mat_off,
                  int64_t n_graphs, int64_t n_rows_total, int32_t* labels_out, int32_t* exemplars_out,
                  int32_t* n_clusters_out, void* workspace, size_t workspace_bytes, void* stream);
int cns_graph_layout(const double* points, const int32_t* labels, const int64_t* pt_off, const int64_t* clu_off,
                     int64_t n_graphs, int32_t* member_order, int32_t* cluster_size, int32_t* centers, void* stream);

/* 1-NN keypoint -> exemplar labelling, the only nearest-neighbour search of the
 * reference's clustering (sklearn AffinityPropagation final assignment,
 * argmax_k -||x - e_k||^2 ).  points (N,2) f64 ; point_graph (N,) graph id of
 * each point or NULL (one graph) ; exemplars (sum K,2) f64 with per-graph ranges
 * exemplar_ptr (B+1,) ; labels (N,) int32 = rank of the nearest exemplar inside its graph. */
int cns_label_nearest(const double* points, const int64_t* point_graph, int64_t n_points,
                      const double* exemplars, const int64_t* exemplar_ptr, int32_t* labels,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNS_B200_H */
