// GraphVS.forward + postprocess as one enqueue of kernels on the caller's stream
// (cns/models/graph_vs.py:284-322, cns/models/functions.py:26-38).
#include "common.cuh"
#include "kernels.cuh"
#include "save.cuh"

namespace cns {

struct FwdBuffers {
  int* err; uint8_t* vis_flag; int* clu_rowptr; int* clu_graph; int* graph_ptr; int* gp_tmp; int* scan_scratch;
  int* cur_rowptr; int* cur_src; int* tar_rowptr; int* tar_src; void* csr_ws;
  int* alive_idx; int* alive_cnt; int* tile_graph; int* tile_next; int* n_tiles;
  float* a_dst; float* pos_clu; float* clu_feat; float* partial; float* x_clu;
  float* pq; float* conv0; float* conv1; float* gn0; float* gn1; float* stats0; float* stats1;
  float* x1; float* u; float* hr; float* x2; float* gate;
  float* g; float* ht; float* x_scene; float* hid;          // training only
  int* arg0; int* argg; int* argk; int* arg1;               // training only
  int64_t n_slots;
};

// Scratch always comes from the workspace; activations the backward pass needs come from the
// caller's save buffer when one is given (training), else they alias workspace scratch.
static void carve(Arena& ar, const cns_batch* b, FwdBuffers& f, const SaveBuffers* sv) {
  const int64_t C = b->n_clusters, B = b->n_graphs;
  memset(&f, 0, sizeof(f));
  f.n_slots = C + (b->n_e01 + 31) / 32 + 2;   // one partial slot per (cluster, softmax piece) pair; pieces >= 32 edges
  f.err = ar.take<int>(64);
  f.vis_flag = ar.take<uint8_t>(b->n_nodes + 1);      // keypoints visible in this frame (target side from a cache)
  f.clu_graph = ar.take<int>(C + 1);
  f.gp_tmp = ar.take<int>(B + 1);
  f.scan_scratch = (int*)ar.take<char>(scan_scratch_bytes(B + 1));
  f.csr_ws = ar.take<char>(csr_workspace_bytes(C));
  f.alive_idx = ar.take<int>(C + 1);
  f.alive_cnt = ar.take<int>(B + 1);
  f.tile_graph = ar.take<int>(B + 2); f.tile_next = ar.take<int>(B + 1); f.n_tiles = ar.take<int>(4);
  f.a_dst = ar.take<float>(C * H);
  f.partial = ar.take<float>((size_t)2 * f.n_slots * 3 * H);
  f.pq = ar.take<float>(C * 4 * H);
  if (sv) {
    f.clu_rowptr = sv->clu_rowptr; f.graph_ptr = sv->graph_ptr;
    f.cur_rowptr = sv->cur_rowptr; f.cur_src = sv->cur_src; f.tar_rowptr = sv->tar_rowptr; f.tar_src = sv->tar_src;
    f.pos_clu = sv->pos_clu; f.clu_feat = sv->clu_feat; f.x_clu = sv->x_clu;
    f.conv0 = sv->conv0; f.conv1 = sv->conv1; f.gn0 = sv->gn0; f.gn1 = sv->gn1; f.stats0 = sv->stats0; f.stats1 = sv->stats1;
    f.x1 = sv->x1; f.u = sv->u; f.hr = sv->hr; f.x2 = sv->x2; f.gate = sv->gate;
    f.g = sv->g; f.ht = sv->ht; f.x_scene = sv->x_scene; f.hid = sv->hid;
    f.arg0 = sv->arg0; f.argg = sv->argg; f.argk = sv->argk; f.arg1 = sv->arg1;
  } else {
    f.clu_rowptr = ar.take<int>(C + 1);
    f.graph_ptr = ar.take<int>(B + 1);
    f.cur_rowptr = ar.take<int>(C + 1);
    f.cur_src = ar.take<int>(b->n_e11_cur + 1);
    f.tar_rowptr = ar.take<int>(C + 1);
    f.tar_src = ar.take<int>(b->n_e11_tar + 1);
    f.pos_clu = ar.take<float>(C * 2 + 2);
    f.clu_feat = ar.take<float>(2 * C * H);
    f.x_clu = ar.take<float>(C * H);
    f.conv0 = f.conv1 = ar.take<float>(C * H);
    f.gn0 = f.gn1 = ar.take<float>(C * H);
    f.x1 = ar.take<float>(C * H);
    f.u = ar.take<float>(C * H);
    f.hr = ar.take<float>(C * H);
    f.x2 = ar.take<float>(C * H);
    f.gate = ar.take<float>(C + 1);
  }
}

}  // namespace cns

// target-side results kept across frames (cns_target_cache_create)
struct cns_target_cache {
  long long uid;                    // unique per created cache (a destroyed cache's address may come back: the graph cache keys on this)
  int device; int precision; long long weights_version;
  int64_t n_clusters, n_graphs;
  int has_tiles;
  float* a_dst; float* pos_clu; int* clu_graph; int* graph_ptr; int* tile_graph; int* n_tiles;
  // per-keypoint target-side logits (fp32, log2 units) and values (f16) of the keypoint -> cluster attention (f16 mode,
  // when the caller passed the keypoint -> cluster assignment of ALL target keypoints): the per-frame target side is then a
  // streaming segment softmax over the visible keypoints (enc_tar_cached) instead of half of the encoder edge kernel
  int64_t n_nodes; float* edge_l; void* edge_v;
  // ... plus every cluster's full member list (cluster-major) and its softmax statistics over ALL members: a frame that
  // keeps most keypoints of a cluster only reads the rows of the missing ones and subtracts them from the totals
  int* all_rowptr; int* all_src; float* all_stats;
  void* base;
};

namespace cns {

static int validate(const cns_batch* b) {
  CNS_REQUIRE(b != nullptr, "batch is NULL");
  CNS_REQUIRE(b->n_nodes >= 0 && b->n_clusters >= 0 && b->n_graphs >= 0 && b->n_e01 >= 0 &&
              b->n_e11_cur >= 0 && b->n_e11_tar >= 0, "negative size in cns_batch");
  CNS_REQUIRE(b->n_nodes < (1ll << 31) && b->n_clusters < (1ll << 31) && b->n_e01 < (1ll << 31) &&
              b->n_e11_cur < (1ll << 31) && b->n_e11_tar < (1ll << 31), "sizes must fit int32");
  if (b->n_nodes > 0) CNS_REQUIRE(b->x_cur && b->x_tar && b->pos_cur && b->pos_tar, "NULL keypoint tensor");
  if (b->n_e01 > 0) CNS_REQUIRE(b->l0_to_l1_j && b->l0_to_l1_i, "NULL l0_to_l1 edge list");
  if (b->n_e11_cur > 0 && !(b->flags & CNS_BATCH_L1_CUR_COMPLETE)) CNS_REQUIRE(b->l1_edge_cur, "NULL l1_dense_edge_index_cur");
  if (b->n_e11_tar > 0 && !(b->flags & CNS_BATCH_L1_TAR_COMPLETE)) CNS_REQUIRE(b->l1_edge_tar, "NULL l1_dense_edge_index_tar");
  if (b->n_clusters > 0) CNS_REQUIRE(b->cluster_mask && b->cluster_centers, "NULL cluster tensor");
  if (b->n_graphs > 0) CNS_REQUIRE(b->num_clusters, "NULL num_clusters");
  return CNS_OK;
}

// per-cluster encoder terms a' = A0 (lin_dst x_c + b_pos1) + b_attn0, pos_clu, clu_graph: in f16 mode one gemm_tc product
// whose A rows are produced on the fly from the cluster centres (the FFMA kernel took 36 us at 19 k clusters)
static int cluster_pre(cns_handle* h, const EncPreArgs& pre, cudaStream_t s) {
  if (h->precision == CNS_PREC_F16 && pre.a_bias && h->pk.tc_adst[1] && !h->no_tc_cluster_pre) {
    GemmTcArgs ga{};
    ga.m = pre.n_clusters; ga.k1 = H; ga.k2 = 0; ga.n_out = H; ga.w_img = h->pk.tc_adst[1];
    ga.bias = pre.a_bias; ga.relu = 0; ga.y = pre.a_dst;
    ga.centers = pre.centers; ga.ctr_xy = pre.x_tar; ga.ctr_pos = pre.pos_tar; ga.w_in = pre.w_in; ga.b_in = pre.b_in;
    ga.node_graph = pre.node_graph; ga.pos_out = pre.pos_clu; ga.clu_graph = pre.clu_graph;
    return gemm_tc(ga, false, s);
  }
  return enc_cluster_pre(pre, s);
}

// `cap_ev`: the call is being captured into a CUDA graph with kernel timing on - the encoder edge kernel is bracketed by
// these two events (recorded as external event nodes, readable after every replay) instead of the handle's event ring
int forward_impl(cns_handle* h, const cns_batch* b, const float* hidden_in, float* vel_si_vec,
                 float* vel_si_norm, float* hidden_out, float* vel, void* workspace, void* save, cudaStream_t s,
                 cudaEvent_t* cap_ev = nullptr) {
  Arena ar(workspace);
  FwdBuffers f;
  SaveBuffers sv_store;
  const SaveBuffers* sv = nullptr;
  if (save) {
    Arena sa(save);
    carve_save(sa, b, sv_store);
    sv = &sv_store;
  }
  carve(ar, b, f, sv);
  const int64_t C = b->n_clusters, B = b->n_graphs;
  const cns_target_cache* tcache = sv ? nullptr : b->target_cache;
  if (tcache) {
    CNS_REQUIRE(tcache->device == h->device && tcache->precision == h->precision && tcache->weights_version == h->weights_version &&
                tcache->n_clusters == C && tcache->n_graphs == B, "stale or foreign cns_target_cache");
    f.a_dst = tcache->a_dst; f.pos_clu = tcache->pos_clu; f.clu_graph = tcache->clu_graph;
  }
  const Packed& pk = h->pk;
  int rc;
#define RUN(expr) if ((rc = (expr)) != CNS_OK) return rc;

  CNS_CUDA_CHECK(cudaMemsetAsync(f.err, 0, 64 * sizeof(int), s));
  // ---- per-cluster encoder terms (target side only) ------------------------------------------
  // independent of the index preparation below: on batches large enough to fill the GPU they run on a forked
  // stream beside the (single-CTA, latency-bound) scans and rejoin before the edge kernel
  EncPreArgs pre{};
  pre.n_clusters = C; pre.x_tar = b->x_tar; pre.pos_tar = b->pos_tar; pre.centers = b->cluster_centers;
  pre.node_graph = b->node_graph; pre.w_in = h->P(P_IN_W); pre.b_in = h->P(P_IN_B); pre.lin_dst_t = pk.lin_dst_t;
  // every tensor-core edge kernel (16-bit modes and the operand-split fp32 mode) consumes the composed form
  const bool composed = h->precision != CNS_PREC_FP32 || !h->no_split_encoder;
  if (composed) { pre.lin_dst_t = pk.tc_lin_dst_t; pre.a_bias = pk.tc_adst_bias; }   // a_dst := A0 (a_dst + b_pos1) + b_attn0
  pre.a_dst = f.a_dst; pre.pos_clu = f.pos_clu; pre.clu_graph = f.clu_graph;
  const bool fork_pre = !tcache && (C > PREP_SMALL_MAX || B > PREP_SMALL_MAX);
  if (fork_pre) {
    if (!h->side_stream) {
      CNS_CUDA_CHECK(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
      CNS_CUDA_CHECK(cudaEventCreateWithFlags(&h->side_fork, cudaEventDisableTiming));
      CNS_CUDA_CHECK(cudaEventCreateWithFlags(&h->side_join, cudaEventDisableTiming));
      CNS_CUDA_CHECK(cudaEventCreateWithFlags(&h->side_join2, cudaEventDisableTiming));
    }
    CNS_CUDA_CHECK(cudaEventRecord(h->side_fork, s));
    CNS_CUDA_CHECK(cudaStreamWaitEvent(h->side_stream, h->side_fork, 0));
    RUN(cluster_pre(h, pre, h->side_stream));
    CNS_CUDA_CHECK(cudaEventRecord(h->side_join, h->side_stream));
  }
  // graph pointers, whole-graph tiles and the hint verification feed the cluster-level backbone only: on the forked
  // stream they run beside the edge kernel
  cudaStream_t s2 = fork_pre ? h->side_stream : s;
  // ---- index preparation ------------------------------------------------------------------
  const int64_t cur_stride = b->l1_cur_stride ? b->l1_cur_stride : b->n_e11_cur;
  const int64_t tar_stride = b->l1_tar_stride ? b->l1_tar_stride : b->n_e11_tar;
  // complete cluster graphs (hint, verified below): no CSR, one column max per graph instead of a gather per row.
  // Training (save != NULL) keeps the general path: the backward pass walks the CSRs.
  const bool dense_cur = (b->flags & CNS_BATCH_L1_CUR_COMPLETE) && !sv && B <= 8192;
  const bool dense_tar = (b->flags & CNS_BATCH_L1_TAR_COMPLETE) && !sv && B <= 8192;
  if (!dense_cur && b->n_e11_cur > 0) CNS_REQUIRE(b->l1_edge_cur, "NULL l1_dense_edge_index_cur");
  if (!dense_tar && b->n_e11_tar > 0) CNS_REQUIRE(b->l1_edge_tar, "NULL l1_dense_edge_index_tar");
  const int64_t ne_cur = dense_cur ? 0 : b->n_e11_cur, ne_tar = dense_tar ? 0 : b->n_e11_tar;   // edges that need a CSR
  // Cluster-level GEMMs (7 % of the FLOPs) go to the tensor cores only in f16 mode: with bf16's 8-bit
  // mantissa on BOTH the edge-level and the recurrent cluster-level layers the hidden state drifts past
  // the 2e-2 budget after a few frames, so bf16 mode keeps them in fp32 (measured).
  const bool tc = h->precision == CNS_PREC_F16;
  // fp32 mode: the same GEMMs with (hi, lo) operand splits (three tensor-core products, ~1e-5 of fp64); bf16 mode takes
  // them too (its cluster level has to stay at fp32 accuracy, see above - on the tensor cores instead of the FFMA kernel)
  const bool tc2 = (h->precision == CNS_PREC_FP32 && !h->no_split_encoder) || (h->precision == CNS_PREC_BF16 && !h->no_split_gemm);
  const int tci = (h->precision == CNS_PREC_BF16) ? 0 : 1;
  // whole-graph row tiles: the three backbone blocks as one kernel (backbone_fused.cu)
  const bool fused = tc && dense_cur && dense_tar && b->max_clusters_per_graph > 0 && b->max_clusters_per_graph <= 128 &&
                     B <= GRAPH_TILES_MAX && !h->no_fused_backbone;
  if (C <= PREP_SMALL_MAX && B <= PREP_SMALL_MAX && b->n_e01 <= (1 << 20) && ne_cur + ne_tar <= (1 << 20)) {
    // servo-loop sizes: one single-CTA launch instead of thirteen
    PrepSmallArgs pa{};
    pa.n_clusters = C; pa.n_graphs = B; pa.n_e01 = b->n_e01; pa.n_e11_cur = ne_cur; pa.n_e11_tar = ne_tar;
    pa.num_clusters = b->num_clusters; pa.ei = b->l0_to_l1_i;
    pa.cur_j = b->l1_edge_cur; pa.cur_i = b->l1_edge_cur + cur_stride; pa.tar_j = b->l1_edge_tar; pa.tar_i = b->l1_edge_tar + tar_stride;
    pa.graph_ptr = f.graph_ptr; pa.clu_rowptr = f.clu_rowptr; pa.cur_rowptr = f.cur_rowptr; pa.cur_src = f.cur_src;
    pa.tar_rowptr = f.tar_rowptr; pa.tar_src = f.tar_src; pa.err = f.err;
    RUN(prep_small(pa, s));
    if (fused && tcache && tcache->has_tiles) { f.tile_graph = tcache->tile_graph; f.n_tiles = tcache->n_tiles; }
    else if (fused) RUN(graph_tiles(b->num_clusters, B, f.graph_ptr, f.tile_graph, f.n_tiles, f.err, s));
  } else {
    if (!b->clu_rowptr || sv) RUN(sorted_rowptr(b->l0_to_l1_i, b->n_e01, C, f.clu_rowptr, f.err, s));
    if (tcache && (!fused || tcache->has_tiles)) {
      f.graph_ptr = tcache->graph_ptr; f.tile_graph = tcache->tile_graph; f.n_tiles = tcache->n_tiles;
    } else if (fused) { RUN(graph_tiles(b->num_clusters, B, f.graph_ptr, f.tile_graph, f.n_tiles, f.err, s2)); }
    else { RUN(graph_ptr_from_counts(b->num_clusters, B, f.graph_ptr, f.gp_tmp, f.scan_scratch, s2)); }
    if (!dense_cur) RUN(csr_by_dst(b->l1_edge_cur, b->l1_edge_cur + cur_stride, b->n_e11_cur, C, f.cur_rowptr, f.cur_src, f.csr_ws, f.err, s));
    if (!dense_tar) RUN(csr_by_dst(b->l1_edge_tar, b->l1_edge_tar + tar_stride, b->n_e11_tar, C, f.tar_rowptr, f.tar_src, f.csr_ws, f.err, s));
  }
  if (b->clu_rowptr && !sv) f.clu_rowptr = const_cast<int*>(b->clu_rowptr);   // the caller's segment pointers (cns_frame_mask)
  if (dense_cur || dense_tar) {
    L1CompleteArgs la{};
    la.n_graphs = B; la.graph_ptr = f.graph_ptr; la.alive_idx = f.alive_idx; la.alive_cnt = f.alive_cnt; la.err = f.err;
    if (dense_cur) {
      la.cur_mask = b->cluster_mask;
      if (b->l1_edge_cur) { la.cur_j = b->l1_edge_cur; la.cur_i = b->l1_edge_cur + cur_stride; la.n_e11_cur = b->n_e11_cur; }
    }
    if (dense_tar && b->l1_edge_tar) { la.tar_j = b->l1_edge_tar; la.tar_i = b->l1_edge_tar + tar_stride; la.n_e11_tar = b->n_e11_tar; }
    RUN(l1_complete_verify(la, s2));
  }
  if (fork_pre) CNS_CUDA_CHECK(cudaEventRecord(h->side_join2, h->side_stream));

  // ---- encoder ----------------------------------------------------------------------------
  if (fork_pre) { CNS_CUDA_CHECK(cudaStreamWaitEvent(s, h->side_join, 0)); }
  else if (!tcache) RUN(cluster_pre(h, pre, s));

  // target side from the cache: a streaming softmax on the forked stream beside the (then current-side only) edge kernel
  const bool tar_cached = tcache && tcache->n_nodes > 0 && tcache->n_nodes == b->n_nodes && h->precision == CNS_PREC_F16;
  const bool fork_tar = tar_cached && !fork_pre && (C > PREP_SMALL_MAX || B > PREP_SMALL_MAX) && !getenv("CNS_TAR_SERIAL");
  if (fork_tar && !h->low_stream) {
    // the streaming kernel runs beside the edge kernel on a LOW-priority stream and is enqueued AFTER it: its 19 k small
    // CTAs would otherwise fill every SM first and the edge kernel (one 207 KB CTA per SM) would wait for them to drain
    int lo = 0, hi = 0;
    CNS_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CNS_CUDA_CHECK(cudaStreamCreateWithPriority(&h->low_stream, cudaStreamNonBlocking, lo));
    CNS_CUDA_CHECK(cudaEventCreateWithFlags(&h->low_fork, cudaEventDisableTiming));
    CNS_CUDA_CHECK(cudaEventCreateWithFlags(&h->low_join, cudaEventDisableTiming));
  }
  if (fork_tar) CNS_CUDA_CHECK(cudaEventRecord(h->low_fork, s));
  auto launch_tar = [&]() -> int {
    ++h->edge_cache_forwards;          // (enqueues, also while a graph is being captured; replays are not counted)
    cudaStream_t st = s;
    if (fork_tar) {
      CNS_CUDA_CHECK(cudaStreamWaitEvent(h->low_stream, h->low_fork, 0));
      st = h->low_stream;
    }
    EncTarCachedArgs ta{};
    ta.ej = b->l0_to_l1_j; ta.n_e01 = b->n_e01; ta.clu_rowptr = f.clu_rowptr; ta.n_clusters = C; ta.n_nodes = b->n_nodes;
    ta.lc = tcache->edge_l; ta.vc = tcache->edge_v; ta.b_p1 = h->P(P_POS1_B);
    ta.all_src = tcache->all_src; ta.all_rowptr = tcache->all_rowptr; ta.all_stats = h->no_edge_subtract ? nullptr : tcache->all_stats;
    ta.visible = f.vis_flag; ta.out = f.clu_feat + (size_t)C * H;
    int r = enc_tar_cached(ta, st);
    if (r != CNS_OK) return r;
    if (fork_tar) CNS_CUDA_CHECK(cudaEventRecord(h->low_join, h->low_stream));
    return CNS_OK;
  };
  EncEdgeArgs ee{};
  ee.side_mode = tar_cached ? 1 : 0;
  ee.n_e01 = b->n_e01; ee.n_clusters = C; ee.n_slots = f.n_slots;
  ee.x_cur = b->x_cur; ee.x_tar = b->x_tar; ee.pos_cur = b->pos_cur; ee.pos_tar = b->pos_tar;
  ee.ej = b->l0_to_l1_j; ee.ei = b->l0_to_l1_i;
  ee.pos_clu = f.pos_clu; ee.a_dst = f.a_dst; ee.clu_rowptr = f.clu_rowptr;
  ee.w_in = h->P(P_IN_W); ee.b_in = h->P(P_IN_B); ee.w_p0 = h->P(P_POS0_W); ee.b_p0 = h->P(P_POS0_B);
  ee.enc_chain = pk.enc_chain; ee.b_p1 = h->P(P_POS1_B); ee.b_a0 = h->P(P_ATT0_B); ee.b_a1 = h->P(P_ATT1_B);
  ee.clu_feat = f.clu_feat; ee.partial = f.partial;
  const bool timed = !cap_ev && h->timing && h->ev_n < 256;
  if (timed) CNS_CUDA_CHECK(cudaEventRecord(h->ev[2 * h->ev_n], s));
  if (cap_ev) CNS_CUDA_CHECK(cudaEventRecordWithFlags(cap_ev[0], s, cudaEventRecordExternal));
  if (h->precision == CNS_PREC_FP32) {
    if (h->no_split_encoder) { RUN(enc_edge_fp32(ee, s)); }
    else { RUN(enc_edge_split(ee, h, s)); }
  } else {
    RUN(enc_edge_tc(ee, h->precision, h, s));
  }
  if (timed) { CNS_CUDA_CHECK(cudaEventRecord(h->ev[2 * h->ev_n + 1], s)); ++h->ev_n; }
  if (cap_ev) CNS_CUDA_CHECK(cudaEventRecordWithFlags(cap_ev[1], s, cudaEventRecordExternal));
  if (tar_cached) RUN(launch_tar());

  EncFinishArgs ef{};
  ef.n_clusters = C; ef.n_slots = f.n_slots;
  ef.piece_e = (h->precision == CNS_PREC_FP32) ? (h->no_split_encoder ? ENC_PIECE_E : split_piece_edges(b->n_e01, h)) : tc_piece_edges(b->n_e01, h, tar_cached ? 1 : 2); ef.clu_rowptr = f.clu_rowptr; ef.partial = f.partial;
  ef.cur_only = tar_cached ? 1 : 0;
  if (fork_tar) CNS_CUDA_CHECK(cudaStreamWaitEvent(s, h->low_join, 0));      // the cached target side has landed
  ef.clu_feat = f.clu_feat; ef.out_t = pk.out_t; ef.b_out = h->P(P_OUT_B); ef.cluster_mask = b->cluster_mask;
  ef.x_clu = f.x_clu;
  ef.cat_out = ((tc && !fused) || tc2) ? f.pq : nullptr;       // (C,256) scratch; pq is free until the first [P|Q] projection
  ef.fix_only = fused ? 1 : 0;
  RUN(enc_finish(ef, s));
  if ((tc && !fused) || tc2) {
    GemmTcArgs ga{};
    ga.m = C; ga.k1 = 2 * H; ga.k2 = 0; ga.n_out = H; ga.x1 = f.pq; ga.w_img = tc2 ? pk.tc2_out : pk.tc_out[tci];
    ga.bias = h->P(P_OUT_B); ga.row_mask = b->cluster_mask; ga.relu = 1; ga.y = f.x_clu;
    if (tc2) { RUN(gemm_tc_split(ga, h->split_f16 != 0, s)); } else { RUN(gemm_tc(ga, tci == 0, s)); }
  }

  // ---- backbone ---------------------------------------------------------------------------
  if (fork_pre) CNS_CUDA_CHECK(cudaStreamWaitEvent(s, h->side_join2, 0));
  auto pq_linear = [&](const float* x1, const float* x2, const float* wt, const float* wp, const float* bias,
                       int n_out, unsigned char* const* img, int k_img, const unsigned char* img2) {
    if (tc || tc2) {
      GemmTcArgs ga{};
      ga.m = C; ga.k1 = H; ga.k2 = x2 ? H : 0; ga.n_out = n_out; ga.x1 = x1; ga.x2 = x2; ga.w_img = tc2 ? img2 : img[tci];
      ga.k_img = k_img;
      ga.bias = bias; ga.pos = f.pos_clu; ga.wp = wp; ga.relu = 0; ga.y = f.pq;
      return tc2 ? gemm_tc_split(ga, h->split_f16 != 0, s) : gemm_tc(ga, tci == 0, s);
    }
    LinearArgs la{};
    la.m = C; la.k1 = H; la.k2 = x2 ? H : 0; la.n_out = n_out;
    la.x1 = x1; la.x2 = x2; la.wt = wt; la.bias = bias; la.pos = f.pos_clu; la.wp = wp;
    la.residual = nullptr; la.relu = 0; la.y = f.pq;
    return linear(la, s);
  };
  auto perconv = [&](const float* x, const float* wt, const float* wp, const float* bias, int bn_w, int bn_b,
                     int bn_ms, const float* fc_t, int fc_b, float* y, unsigned char* const* img,
                     unsigned char* const* fc_img, float* conv, float* gnb, float* stats, int* arg,
                     const unsigned char* img2, const unsigned char* fc_img2) {
    int r;
    if ((r = pq_linear(x, nullptr, wt, wp, bias, 2 * H, img, H, img2)) != CNS_OK) return r;
    if (dense_cur) {
      EdgeConvDenseArgs ed{};
      ed.n_graphs = B; ed.f_out = H; ed.mode = EC_PLAIN; ed.pq = f.pq; ed.graph_ptr = f.graph_ptr; ed.mask = b->cluster_mask;
      ed.gn_weight = h->P(bn_w); ed.gn_bias = h->P(bn_b); ed.gn_mean_scale = h->P(bn_ms); ed.gn_out = gnb;
      if ((r = edgeconv_dense(ed, s)) != CNS_OK) return r;
    } else {
    EdgeConvArgs ec{};
    ec.n_rows = C; ec.f_out = H; ec.mode = EC_PLAIN; ec.pq = f.pq; ec.rowptr = f.cur_rowptr; ec.src = f.cur_src;
    ec.out = conv; ec.argmax = arg;
    if ((r = edgeconv_max(ec, s)) != CNS_OK) return r;
    GraphNormArgs gn{};
    gn.n_graphs = B; gn.graph_ptr = f.graph_ptr; gn.x = conv; gn.weight = h->P(bn_w); gn.bias = h->P(bn_b);
    gn.mean_scale = h->P(bn_ms); gn.y = gnb; gn.stats = stats;
    if ((r = graphnorm_relu(gn, s)) != CNS_OK) return r;
    }
    if (tc || tc2) {
      GemmTcArgs ga{};
      ga.m = C; ga.k1 = H; ga.k2 = 0; ga.n_out = H; ga.x1 = gnb; ga.w_img = tc2 ? fc_img2 : fc_img[tci]; ga.bias = h->P(fc_b);
      ga.residual = x; ga.relu = 1; ga.y = y;
      return tc2 ? gemm_tc_split(ga, h->split_f16 != 0, s) : gemm_tc(ga, tci == 0, s);
    }
    LinearArgs la{};
    la.m = C; la.k1 = H; la.k2 = 0; la.n_out = H; la.x1 = gnb; la.wt = fc_t; la.bias = h->P(fc_b);
    la.residual = x; la.relu = 1; la.y = y;   // F.relu(l1_conv(x)) of Backbone.forward folded in
    return linear(la, s);
  };

  if (fused) {
    BackboneFusedArgs fa{};
    fa.n_clusters = C; fa.n_graphs = (int)B; fa.graph_ptr = f.graph_ptr; fa.tile_graph = f.tile_graph; fa.n_tiles = f.n_tiles;
    fa.clu_feat = f.clu_feat; fa.hidden_in = hidden_in; fa.pos_clu = f.pos_clu; fa.mask = b->cluster_mask;
    fa.w_out = pk.tc_out[tci]; fa.out_b = h->P(P_OUT_B);
    fa.w_c0 = pk.tc_c0[tci]; fa.w_c0fc = pk.tc_c0_fc[tci]; fa.w_g = pk.tc_g[tci]; fa.w_k = pk.tc_k[tci];
    fa.w_c1 = pk.tc_c1[tci]; fa.w_c1fc = pk.tc_c1_fc[tci];
    fa.c0_b = pk.c0_b; fa.c0_wp = pk.c0_wp; fa.g_b = pk.g_b; fa.g_wp = pk.g_wp; fa.k_b = pk.k_b; fa.k_wp = pk.k_wp;
    fa.c1_b = pk.c1_b; fa.c1_wp = pk.c1_wp; fa.c0fc_b = h->P(P_C0_FC_B); fa.c1fc_b = h->P(P_C1_FC_B);
    fa.gn0_w = h->P(P_C0_BN_W); fa.gn0_b = h->P(P_C0_BN_B); fa.gn0_ms = h->P(P_C0_BN_MS);
    fa.gn1_w = h->P(P_C1_BN_W); fa.gn1_b = h->P(P_C1_BN_B); fa.gn1_ms = h->P(P_C1_BN_MS);
    fa.vec_tab = pk.bf_vec; fa.hidden_out = hidden_out; fa.x2 = f.x2;
    RUN(backbone_fused(fa, tci == 0, s));
  } else {
  RUN(perconv(f.x_clu, pk.c0_wt, pk.c0_wp, pk.c0_b, P_C0_BN_W, P_C0_BN_B, P_C0_BN_MS, pk.c0_fc_t, P_C0_FC_B, f.x1, pk.tc_c0, pk.tc_c0_fc, f.conv0, f.gn0, f.stats0, f.arg0, pk.tc2_c0, pk.tc2_c0_fc));

  // GRU gates on the TARGET cluster graph, candidate on the CURRENT one (graph_vs.py:246-249)
  RUN(pq_linear(f.x1, hidden_in, pk.g_wt, pk.g_wp, pk.g_b, 4 * H, pk.tc_g, 2 * H, pk.tc2_g));
  if (dense_tar) {
    EdgeConvDenseArgs ed{};
    ed.n_graphs = B; ed.f_out = 2 * H; ed.mode = EC_GATES; ed.pq = f.pq; ed.graph_ptr = f.graph_ptr; ed.mask = nullptr;
    ed.h = hidden_in; ed.hr = f.hr; ed.u = f.u;
    RUN(edgeconv_dense(ed, s));
  } else {
    EdgeConvArgs ec{};
    ec.n_rows = C; ec.f_out = 2 * H; ec.mode = EC_GATES; ec.pq = f.pq; ec.rowptr = f.tar_rowptr; ec.src = f.tar_src;
    ec.h = hidden_in; ec.hr = f.hr; ec.u = f.u; ec.argmax = f.argg; ec.act = f.g;
    RUN(edgeconv_max(ec, s));
  }
  RUN(pq_linear(f.x1, hidden_in ? f.hr : nullptr, pk.k_wt, pk.k_wp, pk.k_b, 2 * H, pk.tc_k, 2 * H, pk.tc2_k));
  if (dense_cur) {
    EdgeConvDenseArgs ed{};
    ed.n_graphs = B; ed.f_out = H; ed.mode = EC_CANDIDATE; ed.pq = f.pq; ed.graph_ptr = f.graph_ptr; ed.mask = b->cluster_mask;
    ed.h = hidden_in; ed.u_in = f.u; ed.h_out = hidden_out;
    RUN(edgeconv_dense(ed, s));
  } else {
    EdgeConvArgs ec{};
    ec.n_rows = C; ec.f_out = H; ec.mode = EC_CANDIDATE; ec.pq = f.pq; ec.rowptr = f.cur_rowptr; ec.src = f.cur_src;
    ec.h = hidden_in; ec.u_in = f.u; ec.h_out = hidden_out; ec.argmax = f.argk; ec.act = f.ht;
    RUN(edgeconv_max(ec, s));
  }
  RUN(perconv(hidden_out, pk.c1_wt, pk.c1_wp, pk.c1_b, P_C1_BN_W, P_C1_BN_B, P_C1_BN_MS, pk.c1_fc_t, P_C1_FC_B, f.x2, pk.tc_c1, pk.tc_c1_fc, f.conv1, f.gn1, f.stats1, f.arg1, pk.tc2_c1, pk.tc2_c1_fc));
  }

  // ---- decoder ----------------------------------------------------------------------------
  PoolHeadArgs ph{};
  ph.n_graphs = B; ph.graph_ptr = f.graph_ptr; ph.x = f.x2; ph.cluster_mask = b->cluster_mask;
  ph.pool_w = h->P(P_POOL_W); ph.pool_b = h->P(P_POOL_B);
  ph.vec0_t = pk.vec0_t; ph.vec0_b = h->P(P_VEC0_B); ph.vec1_w = h->P(P_VEC1_W); ph.vec1_b = h->P(P_VEC1_B);
  ph.nrm0_t = pk.nrm0_t; ph.nrm0_b = h->P(P_NRM0_B); ph.nrm1_w = h->P(P_NRM1_W); ph.nrm1_b = h->P(P_NRM1_B);
  ph.tPo_norm = b->tPo_norm; ph.gate = f.gate;
  ph.vel_si_vec = vel_si_vec; ph.vel_si_norm = vel_si_norm; ph.vel = vel;
  ph.x_scene = f.x_scene; ph.hid = f.hid;
  RUN(pool_head(ph, s));
#undef RUN
  return CNS_OK;
}


// ---- CUDA-graph cache ---------------------------------------------------------------------------------------------
// forward_impl is a pure enqueue (no host read of device data), so a call whose whole argument set repeats - a servo
// loop over persistent buffers, a throughput loop over resident batches - can be replayed as ONE graph launch.  An
// argument set is captured the second time it is seen (one-off calls never pay for a capture); entries are evicted
// least-recently-used.  Everything that shapes the enqueue is part of the key.
struct GraphEntry {
  uint64_t key; int seen; cudaGraphExec_t exec; long long launches; unsigned long long last_use;
  cudaEvent_t ev[2]; int timed, replayed;
};
constexpr int GRAPH_CACHE_N = 48;
struct GraphCache { GraphEntry e[GRAPH_CACHE_N]; unsigned long long clock; };

static uint64_t fnv(uint64_t hsh, const void* p, size_t n) {
  const unsigned char* c = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { hsh ^= c[i]; hsh *= 1099511628211ull; }
  return hsh;
}

static void graph_entry_drop(GraphEntry& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  if (g.ev[0]) { cudaEventDestroy(g.ev[0]); cudaEventDestroy(g.ev[1]); }
  memset(&g, 0, sizeof(g));
}

void graph_cache_clear(cns_handle* h) {
  GraphCache* gc = (GraphCache*)h->graph_cache;
  if (!gc) return;
  for (int i = 0; i < GRAPH_CACHE_N; ++i) graph_entry_drop(gc->e[i]);
  delete gc;
  h->graph_cache = nullptr;
}

void graph_cache_begin_timing(cns_handle* h) {
  GraphCache* gc = (GraphCache*)h->graph_cache;
  if (gc) for (int i = 0; i < GRAPH_CACHE_N; ++i) gc->e[i].replayed = 0;
}

// the encoder-edge-kernel time of the LAST replay of every timed graph replayed since cns_profile_begin
int graph_cache_collect_timing(cns_handle* h, float* ms_total, int* n) {
  *ms_total = 0.f; *n = 0;
  GraphCache* gc = (GraphCache*)h->graph_cache;
  if (!gc) return CNS_OK;
  for (int i = 0; i < GRAPH_CACHE_N; ++i) {
    GraphEntry& g = gc->e[i];
    if (!g.exec || !g.timed || !g.replayed) continue;
    float ms = 0.f;
    CNS_CUDA_CHECK(cudaEventSynchronize(g.ev[1]));
    CNS_CUDA_CHECK(cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]));
    *ms_total += ms; ++*n;
    g.replayed = 0;
  }
  return CNS_OK;
}

static int forward_graphed(cns_handle* h, const cns_batch* b, const float* hidden_in, float* vel_si_vec,
                           float* vel_si_norm, float* hidden_out, float* vel, void* workspace, cudaStream_t s) {
  if (!h->graph_cache) { h->graph_cache = new GraphCache(); memset(h->graph_cache, 0, sizeof(GraphCache)); }
  GraphCache* gc = (GraphCache*)h->graph_cache;
  uint64_t key = 1469598103934665603ull;
  key = fnv(key, b, sizeof(*b));
  const void* ptrs[6] = {hidden_in, vel_si_vec, vel_si_norm, hidden_out, vel, workspace};
  key = fnv(key, ptrs, sizeof(ptrs));
  const long long mode[8] = {h->precision, h->weights_version, h->timing, h->tc_tile_n, h->tc_slots, h->split_f16,
                             h->no_fused_backbone * 2 + h->no_split_encoder + 4 * h->no_tc_cluster_pre, b->target_cache ? b->target_cache->uid : 0};
  key = fnv(key, mode, sizeof(mode));
  if (key == 0) key = 1;
  GraphEntry* hit = nullptr; GraphEntry* victim = &gc->e[0];
  for (int i = 0; i < GRAPH_CACHE_N; ++i) {
    if (gc->e[i].key == key) { hit = &gc->e[i]; break; }
    if (gc->e[i].last_use < victim->last_use) victim = &gc->e[i];
  }
  ++gc->clock;
  if (hit && hit->exec) {
    hit->last_use = gc->clock; hit->replayed = 1;
    CNS_CUDA_CHECK(cudaGraphLaunch(hit->exec, s));
    g_launch_count += hit->launches;
    ++h->graph_replays;
    return CNS_OK;
  }
  if (!hit || hit->seen < 0) {          // first sight (or not capturable): plain enqueue
    if (!hit) { graph_entry_drop(*victim); victim->key = key; victim->seen = 1; victim->last_use = gc->clock; }
    else hit->last_use = gc->clock;
    return forward_impl(h, b, hidden_in, vel_si_vec, vel_si_norm, hidden_out, vel, workspace, nullptr, s);
  }
  // second sight: capture, instantiate, launch
  hit->last_use = gc->clock;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {   // the caller is capturing already
    cudaGetLastError();
    return forward_impl(h, b, hidden_in, vel_si_vec, vel_si_norm, hidden_out, vel, workspace, nullptr, s);
  }
  // the legacy default stream cannot be captured: record on the handle's side stream instead (a graph can be launched
  // into any stream)
  cudaStream_t cs = s;
  if (s == nullptr || s == cudaStreamLegacy) {
    if (!h->cap_stream) CNS_CUDA_CHECK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    cs = h->cap_stream;
  }
  hit->timed = h->timing;
  if (hit->timed && !hit->ev[0]) { CNS_CUDA_CHECK(cudaEventCreate(&hit->ev[0])); CNS_CUDA_CHECK(cudaEventCreate(&hit->ev[1])); }
  const long long l0 = g_launch_count;
  CNS_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  int rc = forward_impl(h, b, hidden_in, vel_si_vec, vel_si_norm, hidden_out, vel, workspace, nullptr, cs, hit->timed ? hit->ev : nullptr);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(cs, &graph);
  hit->launches = g_launch_count - l0;
  g_launch_count = l0;
  if (rc == CNS_OK && ce == cudaSuccess && graph) {
    ce = cudaGraphInstantiate(&hit->exec, graph, 0);
    if (ce != cudaSuccess) hit->exec = nullptr;
  }
  if (graph) cudaGraphDestroy(graph);
  if (rc != CNS_OK || ce != cudaSuccess || !hit->exec) {
    cudaGetLastError();
    hit->seen = -1;                     // never try again for this argument set
    if (rc != CNS_OK) return rc;        // an argument error: the plain path reports the same
    return forward_impl(h, b, hidden_in, vel_si_vec, vel_si_norm, hidden_out, vel, workspace, nullptr, s);
  }
  ++h->graph_captures;
  hit->replayed = 1;
  CNS_CUDA_CHECK(cudaGraphLaunch(hit->exec, s));
  g_launch_count += hit->launches;
  ++h->graph_replays;
  return CNS_OK;
}

}  // namespace cns

using namespace cns;

extern "C" {

size_t cns_forward_workspace_bytes(const cns_batch* b) {
  if (!b) return 0;
  Arena ar(nullptr);
  FwdBuffers f;
  carve(ar, b, f, nullptr);
  return ar.off + 256;
}

size_t cns_forward_save_bytes(const cns_batch* b) {
  if (!b) return 0;
  Arena ar(nullptr);
  SaveBuffers sv;
  carve_save(ar, b, sv);
  return ar.off + 256;
}

int cns_graphvs_forward(cns_handle* h, const cns_batch* b, const float* hidden_in, float* vel_si_vec,
                        float* vel_si_norm, float* hidden_out, float* vel, void* workspace,
                        size_t workspace_bytes, void* save, size_t save_bytes, void* stream) {
  CNS_REQUIRE(h != nullptr, "NULL handle");
  if (save && save_bytes < cns_forward_save_bytes(b)) {
    set_error("cns_graphvs_forward: save buffer %zu < %zu bytes", save_bytes, cns_forward_save_bytes(b));
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  CNS_REQUIRE(h->weights_loaded, "weights not loaded (cns_load_weights)");
  int rc = validate(b);
  if (rc != CNS_OK) return rc;
  if (b->n_graphs == 0) return CNS_OK;
  CNS_REQUIRE(vel_si_vec && vel_si_norm && (hidden_out || b->n_clusters == 0), "NULL output");
  CNS_REQUIRE(workspace != nullptr, "NULL workspace");
  size_t need = cns_forward_workspace_bytes(b);
  if (workspace_bytes < need) {
    set_error("cns_graphvs_forward: workspace %zu < %zu bytes", workspace_bytes, need);
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  CNS_CUDA_CHECK(cudaSetDevice(h->device));
  if (h->use_graphs && !save)
    return forward_graphed(h, b, hidden_in, vel_si_vec, vel_si_norm, hidden_out, vel, workspace, (cudaStream_t)stream);
  return forward_impl(h, b, hidden_in, vel_si_vec, vel_si_norm, hidden_out, vel, workspace, save, (cudaStream_t)stream);
}

int cns_target_cache_create(cns_handle* h, const cns_batch* b, void* stream, cns_target_cache** out) {
  CNS_REQUIRE(h != nullptr && b != nullptr && out != nullptr, "NULL argument");
  CNS_REQUIRE(h->weights_loaded, "weights not loaded (cns_load_weights)");
  CNS_REQUIRE(b->n_clusters > 0 && b->n_graphs > 0 && b->x_tar && b->pos_tar && b->cluster_centers && b->num_clusters,
              "cns_target_cache_create: target tensors missing");
  CNS_CUDA_CHECK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t C = b->n_clusters, B = b->n_graphs;
  cns_target_cache* c = new cns_target_cache();
  static long long next_uid = 0;
  c->uid = ++next_uid;
  c->device = h->device; c->precision = h->precision; c->weights_version = h->weights_version;
  c->n_clusters = C; c->n_graphs = B;
  int* tmp = nullptr; int* scratch = nullptr; int* err = nullptr; void* csr_ws = nullptr;
  // the edge cache needs: f16 mode (bf16's worst case sits at 1.85e-2 of its 2e-2 budget: no room for the cached values'
  // extra rounding), every target keypoint's cluster (n_e01 == n_nodes edges, any order) and the
  // tensor-core kernels' default tile configuration
  const bool want_edges = h->precision == CNS_PREC_F16 && b->n_e01 > 0 && b->n_e01 == b->n_nodes && b->l0_to_l1_j && b->l0_to_l1_i &&
                          h->tc_tile_n == 32 && !h->no_edge_cache;
  const int64_t N = b->n_nodes;
  c->n_nodes = want_edges ? N : 0;
  auto lay = [&](Arena& a) {
    c->a_dst = a.take<float>(C * H); c->pos_clu = a.take<float>(C * 2 + 2); c->clu_graph = a.take<int>(C + 1);
    c->graph_ptr = a.take<int>(B + 1); c->tile_graph = a.take<int>(B + 2); c->n_tiles = a.take<int>(4);
    tmp = a.take<int>(B + 1); scratch = (int*)a.take<char>(scan_scratch_bytes(B + 1)); err = a.take<int>(4);
    if (want_edges) {
      c->edge_l = a.take<float>(N * H); c->edge_v = a.take<char>(N * H * 2);
      c->all_rowptr = a.take<int>(C + 1); c->all_src = a.take<int>(N + 1); c->all_stats = a.take<float>(C * 3 * H);
      csr_ws = a.take<char>(csr_workspace_bytes(C));
    }
  };
  Arena count(nullptr);
  lay(count);
  if (cudaMalloc(&c->base, count.off + 256) != cudaSuccess) { delete c; set_error("cns_target_cache_create: out of memory"); return CNS_ERR_CUDA; }
  Arena real(c->base);
  lay(real);
  int rc = CNS_OK;
  do {
    if (cudaMemsetAsync(err, 0, 4 * sizeof(int), s) != cudaSuccess) { rc = CNS_ERR_CUDA; break; }
    c->has_tiles = B <= GRAPH_TILES_MAX && b->max_clusters_per_graph > 0 && b->max_clusters_per_graph <= 128;
    if (c->has_tiles) rc = graph_tiles(b->num_clusters, B, c->graph_ptr, c->tile_graph, c->n_tiles, err, s);
    else rc = graph_ptr_from_counts(b->num_clusters, B, c->graph_ptr, tmp, scratch, s);
    if (rc != CNS_OK) break;
    EncPreArgs pre{};
    pre.n_clusters = C; pre.x_tar = b->x_tar; pre.pos_tar = b->pos_tar; pre.centers = b->cluster_centers;
    pre.node_graph = b->node_graph; pre.w_in = h->P(P_IN_W); pre.b_in = h->P(P_IN_B); pre.lin_dst_t = h->pk.lin_dst_t;
    if (h->precision != CNS_PREC_FP32 || !h->no_split_encoder) { pre.lin_dst_t = h->pk.tc_lin_dst_t; pre.a_bias = h->pk.tc_adst_bias; }
    pre.a_dst = c->a_dst; pre.pos_clu = c->pos_clu; pre.clu_graph = c->clu_graph;
    rc = cluster_pre(h, pre, s);
    if (rc != CNS_OK || !want_edges) break;
    // fill the edge cache: the target side of the encoder edge kernel over ALL target keypoints, storing logits and values
    EncEdgeArgs ee{};
    ee.n_e01 = N; ee.n_clusters = C; ee.n_slots = 0;
    ee.x_cur = b->x_tar; ee.x_tar = b->x_tar; ee.pos_cur = b->pos_tar; ee.pos_tar = b->pos_tar;
    ee.ej = b->l0_to_l1_j; ee.ei = b->l0_to_l1_i;
    ee.pos_clu = c->pos_clu; ee.a_dst = c->a_dst; ee.clu_rowptr = nullptr;
    ee.w_in = h->P(P_IN_W); ee.b_in = h->P(P_IN_B); ee.w_p0 = h->P(P_POS0_W); ee.b_p0 = h->P(P_POS0_B);
    ee.enc_chain = h->pk.enc_chain; ee.b_p1 = h->P(P_POS1_B); ee.b_a0 = h->P(P_ATT0_B); ee.b_a1 = h->P(P_ATT1_B);
    ee.clu_feat = nullptr; ee.partial = nullptr;
    ee.side_mode = 2; ee.store_l = c->edge_l; ee.store_v = c->edge_v;
    rc = enc_edge_tc(ee, h->precision, h, s);
    if (rc != CNS_OK) break;
    // cluster-major member lists and the per-cluster statistics over all members
    rc = csr_by_dst(b->l0_to_l1_j, b->l0_to_l1_i, N, C, c->all_rowptr, c->all_src, csr_ws, err, s);
    if (rc != CNS_OK) break;
    EncTarCachedArgs ta{};
    ta.n_clusters = C; ta.n_nodes = N; ta.lc = c->edge_l; ta.vc = c->edge_v; ta.b_p1 = h->P(P_POS1_B);
    ta.all_src = c->all_src; ta.all_rowptr = c->all_rowptr; ta.stats_out = c->all_stats;
    rc = enc_tar_cached(ta, s);
  } while (false);
  if (rc != CNS_OK) { cudaFree(c->base); delete c; return rc; }
  *out = c;
  return CNS_OK;
}

int cns_target_cache_destroy(cns_target_cache* c) {
  if (!c) return CNS_OK;
  cudaSetDevice(c->device);
  cudaFree(c->base);
  delete c;
  return CNS_OK;
}

// Error word written by the index-preparation kernels (0 = fine).  Synchronises `stream`.
int cns_forward_status(const void* workspace, void* stream) {
  CNS_REQUIRE(workspace != nullptr, "NULL workspace");
  int flag = 0;
  CNS_CUDA_CHECK(cudaMemcpyAsync(&flag, workspace, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CNS_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  if (flag == 1) { set_error("l0_to_l1_edge_index_i is not sorted by cluster / out of range"); return CNS_ERR_UNSORTED_EDGES; }
  if (flag == 4) { set_error("a graph has more clusters than cns_batch.max_clusters_per_graph / 128"); return CNS_ERR_INVALID_ARGUMENT; }
  if (flag == 3) { set_error("cluster-level edge lists are not the complete graphs claimed by cns_batch.flags"); return CNS_ERR_INVALID_ARGUMENT; }
  if (flag != 0) { set_error("edge index out of range"); return CNS_ERR_INVALID_ARGUMENT; }
  return CNS_OK;
}

// One servo frame from host memory with the recurrent hidden state kept ON THE DEVICE inside the handle
// (what GraphVSController does with self.hidden, controller.py:23-38): every input array is packed into
// one pinned staging block -> ONE host->device copy, the forward, ONE device->host copy of
// (vel_si_vec | vel | vel_si_norm).  Meant for small batches (the staging memcpy is on the CPU).
int cns_servo_step_host(cns_handle* h, const cns_batch* bh, int reset_hidden, float* vel_host,
                        float* vel_si_vec_host, float* vel_si_norm_host) {
  CNS_REQUIRE(h != nullptr && vel_host != nullptr, "NULL argument");
  int rc = validate(bh);
  if (rc != CNS_OK) return rc;
  if (bh->n_graphs == 0) return CNS_OK;
  CNS_CUDA_CHECK(cudaSetDevice(h->device));
  const int64_t N = bh->n_nodes, C = bh->n_clusters, B = bh->n_graphs;
  cns_batch d = *bh;
  d.l1_cur_stride = 0; d.l1_tar_stride = 0;
  const int64_t cs = bh->l1_cur_stride ? bh->l1_cur_stride : bh->n_e11_cur;
  const int64_t ts = bh->l1_tar_stride ? bh->l1_tar_stride : bh->n_e11_tar;
  // input block layout (offsets shared by the pinned host copy and the device copy)
  size_t off[16]; size_t bytes[16]; const void* src[16]; int n_items = 0; size_t in_bytes = 0;
  auto add = [&](const void* p, size_t nb) {
    off[n_items] = in_bytes; bytes[n_items] = nb; src[n_items] = p; ++n_items;
    in_bytes += align_up(nb);
    return off[n_items - 1];
  };
  const size_t o_xc = add(bh->x_cur, N * 2 * sizeof(float)), o_xt = add(bh->x_tar, N * 2 * sizeof(float));
  const size_t o_pc = add(bh->pos_cur, N * 2 * sizeof(float)), o_pt = add(bh->pos_tar, N * 2 * sizeof(float));
  const size_t o_ej = add(bh->l0_to_l1_j, bh->n_e01 * sizeof(int64_t)), o_ei = add(bh->l0_to_l1_i, bh->n_e01 * sizeof(int64_t));
  const size_t o_cj = add(bh->l1_edge_cur, bh->n_e11_cur * sizeof(int64_t));
  add(bh->l1_edge_cur ? bh->l1_edge_cur + cs : nullptr, bh->n_e11_cur * sizeof(int64_t));   // row 1 right after row 0
  const size_t o_tj = add(bh->l1_edge_tar, bh->n_e11_tar * sizeof(int64_t));
  add(bh->l1_edge_tar ? bh->l1_edge_tar + ts : nullptr, bh->n_e11_tar * sizeof(int64_t));
  const size_t o_cm = add(bh->cluster_mask, C), o_cc = add(bh->cluster_centers, C * sizeof(int64_t));
  const size_t o_nc = add(bh->num_clusters, B * sizeof(int64_t));
  const size_t o_tp = bh->tPo_norm ? add(bh->tPo_norm, B * sizeof(float)) : 0;
  // rows of a (2,E) edge list must be contiguous on the device: row 1 sits align_up(row 0) later
  d.l1_cur_stride = (int64_t)(align_up(bh->n_e11_cur * sizeof(int64_t)) / sizeof(int64_t));
  d.l1_tar_stride = (int64_t)(align_up(bh->n_e11_tar * sizeof(int64_t)) / sizeof(int64_t));
  const size_t out_bytes = align_up((size_t)B * 13 * sizeof(float));
  const size_t hid_bytes = align_up((size_t)C * H * sizeof(float));
  const size_t ws_bytes = cns_forward_workspace_bytes(bh);
  const size_t dev_need = in_bytes + out_bytes + ws_bytes;
  const size_t host_need = in_bytes + out_bytes;
  if (dev_need > h->stage_dev_bytes) {
    if (h->stage_dev) CNS_CUDA_CHECK(cudaFree(h->stage_dev));
    h->stage_dev = nullptr; h->stage_dev_bytes = 0;
    CNS_CUDA_CHECK(cudaMalloc(&h->stage_dev, dev_need + dev_need / 4));
    h->stage_dev_bytes = dev_need + dev_need / 4;
  }
  if (2 * hid_bytes > h->servo_hid_bytes) {
    // a different cluster count: the carried state could not be used anyway (`carry` below needs servo_nc == C)
    if (h->servo_hid) CNS_CUDA_CHECK(cudaFree(h->servo_hid));
    h->servo_hid = nullptr; h->servo_hid_bytes = 0; h->servo_valid = 0;
    CNS_CUDA_CHECK(cudaMalloc(&h->servo_hid, 2 * hid_bytes));
    h->servo_hid_bytes = 2 * hid_bytes;
  }
  if (host_need > h->stage_host_bytes) {
    if (h->stage_host) CNS_CUDA_CHECK(cudaFreeHost(h->stage_host));
    h->stage_host = nullptr; h->stage_host_bytes = 0;
    CNS_CUDA_CHECK(cudaMallocHost(&h->stage_host, host_need + host_need / 4));
    h->stage_host_bytes = host_need + host_need / 4;
  }
  char* hb = (char*)h->stage_host;
  char* db = (char*)h->stage_dev;
  for (int i = 0; i < n_items; ++i)
    if (bytes[i] && src[i]) memcpy(hb + off[i], src[i], bytes[i]);
  cudaStream_t s = 0;
  CNS_CUDA_CHECK(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, s));
  d.x_cur = (const float*)(db + o_xc); d.x_tar = (const float*)(db + o_xt);
  d.pos_cur = (const float*)(db + o_pc); d.pos_tar = (const float*)(db + o_pt);
  d.l0_to_l1_j = (const int64_t*)(db + o_ej); d.l0_to_l1_i = (const int64_t*)(db + o_ei);
  d.l1_edge_cur = (const int64_t*)(db + o_cj); d.l1_edge_tar = (const int64_t*)(db + o_tj);
  d.cluster_mask = (const uint8_t*)(db + o_cm); d.cluster_centers = (const int64_t*)(db + o_cc);
  d.node_graph = nullptr;
  d.num_clusters = (const int64_t*)(db + o_nc);
  d.tPo_norm = bh->tPo_norm ? (const float*)(db + o_tp) : nullptr;
  float* out_d = (float*)(db + in_bytes);                 // vec (B,6) | vel (B,6) | nrm (B,1)
  // fixed addresses: slot k of the ping-pong pair never moves while the cluster count stays the same
  float* hid[2] = {(float*)h->servo_hid, (float*)((char*)h->servo_hid + h->servo_hid_bytes / 2)};
  void* ws = db + in_bytes + out_bytes;
  {
    // same target as the previous frame (the usual case in a servo loop)?  then what depends on the target alone is reused
    const size_t kb[4] = {(size_t)N * 2 * sizeof(float), (size_t)N * 2 * sizeof(float), (size_t)C * sizeof(int64_t),
                          (size_t)B * sizeof(int64_t)};
    const void* kp[4] = {bh->x_tar, bh->pos_tar, bh->cluster_centers, bh->num_clusters};
    const size_t key_bytes = kb[0] + kb[1] + kb[2] + kb[3];
    bool same = h->servo_tcache != nullptr && h->servo_tkey_bytes == key_bytes &&
                h->servo_tcache->precision == h->precision && h->servo_tcache->weights_version == h->weights_version &&
                h->servo_tcache->n_clusters == C && h->servo_tcache->n_graphs == B;
    size_t o = 0;
    for (int i = 0; i < 4 && same; ++i) { same = memcmp((char*)h->servo_tkey + o, kp[i], kb[i]) == 0; o += kb[i]; }
    if (!same) {
      if (h->servo_tcache) { cns_target_cache_destroy(h->servo_tcache); h->servo_tcache = nullptr; }
      free(h->servo_tkey);
      h->servo_tkey = malloc(key_bytes > 0 ? key_bytes : 1);
      h->servo_tkey_bytes = key_bytes;
      CNS_REQUIRE(h->servo_tkey != nullptr, "out of host memory");
      o = 0;
      for (int i = 0; i < 4; ++i) { memcpy((char*)h->servo_tkey + o, kp[i], kb[i]); o += kb[i]; }
      rc = cns_target_cache_create(h, &d, s, &h->servo_tcache);
      if (rc != CNS_OK) { h->servo_tkey_bytes = 0; return rc; }
    }
    d.target_cache = h->servo_tcache;
  }
  const bool carry = h->servo_valid && !reset_hidden && h->servo_nc == C;
  const float* hin = carry ? hid[h->servo_cur] : nullptr;
  float* hout = hid[h->servo_cur ^ 1];
  rc = cns_graphvs_forward(h, &d, hin, out_d, out_d + B * 12, hout, out_d + B * 6, ws, ws_bytes, nullptr, 0, s);
  if (rc != CNS_OK) return rc;
  h->servo_cur ^= 1; h->servo_valid = 1; h->servo_nc = C;
  float* out_h = (float*)(hb + in_bytes);
  CNS_CUDA_CHECK(cudaMemcpyAsync(out_h, out_d, (size_t)B * 13 * sizeof(float), cudaMemcpyDeviceToHost, s));
  rc = cns_forward_status(ws, s);                         // synchronises the stream
  if (rc != CNS_OK) return rc;
  if (vel_si_vec_host) memcpy(vel_si_vec_host, out_h, (size_t)B * 6 * sizeof(float));
  memcpy(vel_host, out_h + B * 6, (size_t)B * 6 * sizeof(float));
  if (vel_si_norm_host) memcpy(vel_si_norm_host, out_h + B * 12, (size_t)B * sizeof(float));
  return CNS_OK;
}

int cns_graphvs_forward_host(cns_handle* h, const cns_batch* bh, const float* hidden_in_host,
                             float* vel_si_vec_host, float* vel_si_norm_host, float* hidden_out_host,
                             float* vel_host) {
  CNS_REQUIRE(h != nullptr, "NULL handle");
  int rc = validate(bh);
  if (rc != CNS_OK) return rc;
  if (bh->n_graphs == 0) return CNS_OK;
  CNS_CUDA_CHECK(cudaSetDevice(h->device));
  const int64_t N = bh->n_nodes, C = bh->n_clusters, B = bh->n_graphs;
  // device staging layout: inputs | hidden_in | outputs | workspace
  Arena ar(nullptr);
  auto plan = [&](Arena& a, cns_batch& d, float*& hin, float*& vec, float*& nrm, float*& hout, float*& vel, void*& ws) {
    d = *bh;
    d.l1_cur_stride = 0; d.l1_tar_stride = 0;
    d.x_cur = a.take<float>(N * 2); d.x_tar = a.take<float>(N * 2);
    d.pos_cur = a.take<float>(N * 2); d.pos_tar = a.take<float>(N * 2);
    d.l0_to_l1_j = a.take<int64_t>(bh->n_e01); d.l0_to_l1_i = a.take<int64_t>(bh->n_e01);
    d.l1_edge_cur = a.take<int64_t>(2 * bh->n_e11_cur); d.l1_edge_tar = a.take<int64_t>(2 * bh->n_e11_tar);
    d.cluster_mask = a.take<uint8_t>(C); d.cluster_centers = a.take<int64_t>(C);
    d.node_graph = bh->node_graph ? a.take<int64_t>(N) : nullptr;
    d.num_clusters = a.take<int64_t>(B);
    d.tPo_norm = bh->tPo_norm ? a.take<float>(B) : nullptr;
    hin = hidden_in_host ? a.take<float>(C * H) : nullptr;
    vec = a.take<float>(B * 6); nrm = a.take<float>(B); hout = a.take<float>(C * H); vel = a.take<float>(B * 6);
    ws = a.take<char>(cns_forward_workspace_bytes(bh));
  };
  cns_batch d; float *hin, *vec, *nrm, *hout, *vel; void* ws;
  plan(ar, d, hin, vec, nrm, hout, vel, ws);
  if (ar.off > h->stage_dev_bytes) {
    if (h->stage_dev) CNS_CUDA_CHECK(cudaFree(h->stage_dev));
    h->stage_dev = nullptr; h->stage_dev_bytes = 0;
    size_t want = ar.off + ar.off / 4;
    CNS_CUDA_CHECK(cudaMalloc(&h->stage_dev, want));
    h->stage_dev_bytes = want;
  }
  Arena ad(h->stage_dev);
  plan(ad, d, hin, vec, nrm, hout, vel, ws);
  cudaStream_t s = 0;
#define H2D(dst, src, bytes) \
  if ((bytes) > 0) CNS_CUDA_CHECK(cudaMemcpyAsync((void*)(dst), (src), (bytes), cudaMemcpyHostToDevice, s));
  H2D(d.x_cur, bh->x_cur, N * 2 * sizeof(float)); H2D(d.x_tar, bh->x_tar, N * 2 * sizeof(float));
  H2D(d.pos_cur, bh->pos_cur, N * 2 * sizeof(float)); H2D(d.pos_tar, bh->pos_tar, N * 2 * sizeof(float));
  H2D(d.l0_to_l1_j, bh->l0_to_l1_j, bh->n_e01 * sizeof(int64_t)); H2D(d.l0_to_l1_i, bh->l0_to_l1_i, bh->n_e01 * sizeof(int64_t));
  {
    const int64_t cs = bh->l1_cur_stride ? bh->l1_cur_stride : bh->n_e11_cur;
    const int64_t ts = bh->l1_tar_stride ? bh->l1_tar_stride : bh->n_e11_tar;
    H2D(d.l1_edge_cur, bh->l1_edge_cur, bh->n_e11_cur * sizeof(int64_t));
    H2D(d.l1_edge_cur + bh->n_e11_cur, bh->l1_edge_cur + cs, bh->n_e11_cur * sizeof(int64_t));
    H2D(d.l1_edge_tar, bh->l1_edge_tar, bh->n_e11_tar * sizeof(int64_t));
    H2D(d.l1_edge_tar + bh->n_e11_tar, bh->l1_edge_tar + ts, bh->n_e11_tar * sizeof(int64_t));
  }
  H2D(d.cluster_mask, bh->cluster_mask, C); H2D(d.cluster_centers, bh->cluster_centers, C * sizeof(int64_t));
  if (bh->node_graph) H2D(d.node_graph, bh->node_graph, N * sizeof(int64_t));
  H2D(d.num_clusters, bh->num_clusters, B * sizeof(int64_t));
  if (bh->tPo_norm) H2D(d.tPo_norm, bh->tPo_norm, B * sizeof(float));
  if (hidden_in_host) H2D(hin, hidden_in_host, C * H * sizeof(float));
#undef H2D
  rc = cns_graphvs_forward(h, &d, hin, vec, nrm, hout, vel, ws, cns_forward_workspace_bytes(bh), nullptr, 0, s);
  if (rc != CNS_OK) return rc;
#define D2H(dst, src, bytes) \
  if ((dst) && (bytes) > 0) CNS_CUDA_CHECK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, s));
  D2H(vel_si_vec_host, vec, B * 6 * sizeof(float)); D2H(vel_si_norm_host, nrm, B * sizeof(float));
  D2H(hidden_out_host, hout, C * H * sizeof(float)); D2H(vel_host, vel, B * 6 * sizeof(float));
#undef D2H
  return cns_forward_status(ws, s);
}

}  // extern "C"
