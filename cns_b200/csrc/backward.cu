// Backward of GraphVS.forward (cns/models/graph_vs.py:284-322): gradients of the 43 parameters and of
// the incoming hidden state, for the training loops of cns/utils/trainer.py:75-100 and
// cns/train_gvs_{short,long}_seq.py.  fp32, deterministic.  Cluster-level activations come from the
// forward pass's `save` buffer; per-edge encoder tensors are recomputed (unfused) into the workspace.
#include "common.cuh"
#include "kernels.cuh"
#include "save.cuh"
#include "bwd.cuh"
#include <algorithm>

namespace cns {

struct BwdPack {           // transposed / negated operands for dX = dY W
  float* base;
  float* c0_bwd; float* g_bwd; float* k_bwd; float* c1_bwd;   // (2F, K) = transpose of *_wt
  float* neg_lin_src;                                          // -lin_src.weight, [out][in]
  // bf16-split images (gemm_split.cu) of the ten 128 x 128 matrices the edge-level GEMMs multiply by:
  // 0..4 the recomputed chain (enc_chain 0..4), 5 attn_nn.1, 6 attn_nn.0, 7 lin, 8 -lin_src, 9 pos_nn.1 (all [out][in] = (k, n) for dX)
  unsigned char* split_img;
  const unsigned char* img(int i) const { return split_img + (size_t)i * GEMM_SPLIT_IMG_BYTES; }
};

static int ensure_bwd_pack(cns_handle* h, cudaStream_t s) {
  BwdPack* bp = reinterpret_cast<BwdPack*>(h->bwd_pack);
  if (!bp) {
    bp = new BwdPack();
    const size_t n = 256 * 128 + 512 * 256 + 256 * 256 + 256 * 128 + 128 * 128;
    CNS_CUDA_CHECK(cudaMalloc(&bp->base, n * sizeof(float)));
    CNS_CUDA_CHECK(cudaMalloc(&bp->split_img, 10 * GEMM_SPLIT_IMG_BYTES));
    float* q = bp->base;
    bp->c0_bwd = q; q += 256 * 128; bp->g_bwd = q; q += 512 * 256; bp->k_bwd = q; q += 256 * 256;
    bp->c1_bwd = q; q += 256 * 128; bp->neg_lin_src = q;
    h->bwd_pack = bp;
    h->bwd_pack_version = -1;
  }
  if (h->bwd_pack_version != h->weights_version) {
    int rc;
    if ((rc = transpose(h->pk.c0_wt, 128, 256, 1.f, bp->c0_bwd, s))) return rc;
    if ((rc = transpose(h->pk.g_wt, 256, 512, 1.f, bp->g_bwd, s))) return rc;
    if ((rc = transpose(h->pk.k_wt, 256, 256, 1.f, bp->k_bwd, s))) return rc;
    if ((rc = transpose(h->pk.c1_wt, 128, 256, 1.f, bp->c1_bwd, s))) return rc;
    // enc_chain[1] holds -lin_src^T (k-major); transposing it back gives -lin_src in [out][in] layout
    if ((rc = transpose(h->pk.enc_chain + 1 * 16384, 128, 128, 1.f, bp->neg_lin_src, s))) return rc;
    const float* mats[10] = {h->pk.enc_chain, h->pk.enc_chain + 16384, h->pk.enc_chain + 2 * 16384, h->pk.enc_chain + 3 * 16384,
                             h->pk.enc_chain + 4 * 16384, h->P(P_ATT1_W), h->P(P_ATT0_W), h->P(P_LIN_W), bp->neg_lin_src,
                             h->P(P_POS1_W)};
    for (int i = 0; i < 10; ++i)
      if ((rc = gemm_split_pack(mats[i], bp->split_img + (size_t)i * GEMM_SPLIT_IMG_BYTES, s))) return rc;
    h->bwd_pack_version = h->weights_version;
  }
  return CNS_OK;
}

void free_bwd_pack(cns_handle* h) {
  BwdPack* bp = reinterpret_cast<BwdPack*>(h->bwd_pack);
  if (bp) { cudaFree(bp->base); cudaFree(bp->split_img); delete bp; h->bwd_pack = nullptr; }
}

constexpr int N_EBUF = 12;
constexpr int64_t SPLIT_MIN_ROWS = 4096;   // below this the 96 KB weight image per CTA costs more than the FFMA kernel
struct BwdBuffers {
  int* err; int* curT_rowptr; int* curT_dst; int* tarT_rowptr; int* tarT_dst; void* csr_ws;
  float *dx2, *dz, *dgn, *dconv, *dpq, *dxk, *dxg, *dh_acc, *du, *dconvk, *dconvg, *dx1, *dxclu, *dpre, *cat, *dcat,
      *dclu, *d_adst, *a_dst, *d_xtc, *xtc, *fe_c, *zeros_h, *dhid, *dgate, *gn_partial, *G, *Gp, *gb, *dw_partial, *cs_partial,
      *small_partial;
  float* E[N_EBUF]; float* FE; float* DP;
  int small_slabs;
};

constexpr size_t DW_PARTIAL_FLOATS = (size_t)10 * 1024 * 1024;   // 40 MB of split-M partials
// dw_small splits rows into slabs of 128; sized per batch in carve_bwd

static void carve_bwd(Arena& ar, const cns_batch* b, BwdBuffers& w) {
  const int64_t C = b->n_clusters, B = b->n_graphs, E = b->n_e01;
  w.err = ar.take<int>(64);
  w.curT_rowptr = ar.take<int>(C + 1); w.curT_dst = ar.take<int>(b->n_e11_cur + 1);
  w.tarT_rowptr = ar.take<int>(C + 1); w.tarT_dst = ar.take<int>(b->n_e11_tar + 1);
  w.csr_ws = ar.take<char>(csr_workspace_bytes(C));
  w.dx2 = ar.take<float>(C * H); w.dz = ar.take<float>(C * H); w.dgn = ar.take<float>(C * H); w.dconv = ar.take<float>(C * H);
  w.dpq = ar.take<float>(C * 4 * H); w.dxk = ar.take<float>(C * 2 * H); w.dxg = ar.take<float>(C * 2 * H);
  w.dh_acc = ar.take<float>(C * H); w.du = ar.take<float>(C * H); w.dconvk = ar.take<float>(C * H);
  w.dconvg = ar.take<float>(C * 2 * H); w.dx1 = ar.take<float>(C * H); w.dxclu = ar.take<float>(C * H);
  w.dpre = ar.take<float>(C * H); w.cat = ar.take<float>(C * 2 * H); w.dcat = ar.take<float>(C * 2 * H);
  w.dclu = ar.take<float>(2 * C * H); w.d_adst = ar.take<float>(C * H); w.a_dst = ar.take<float>(C * H);
  w.d_xtc = ar.take<float>(C * H); w.xtc = ar.take<float>(C * H); w.fe_c = ar.take<float>(C * 2 + 2);
  w.zeros_h = ar.take<float>(C * H); w.dhid = ar.take<float>(B * 2 * H); w.dgate = ar.take<float>(C + 1);
  w.gn_partial = ar.take<float>(B * 3 * H);
  w.G = ar.take<float>(512 * 256); w.Gp = ar.take<float>(512 * 2); w.gb = ar.take<float>(512);
  w.dw_partial = ar.take<float>(DW_PARTIAL_FLOATS);
  w.cs_partial = ar.take<float>(592 * 128);       // dw_gemm: slabs * na <= 592 * 128
  w.small_slabs = (int)((std::max<int64_t>(std::max<int64_t>(E, C), B) + 127) / 128) + 1;
  w.small_partial = ar.take<float>((size_t)w.small_slabs * 512 * 4);   // n*q <= 512*2 or 128*8
  for (int i = 0; i < N_EBUF; ++i) w.E[i] = ar.take<float>((size_t)E * H + 4);
  w.FE = ar.take<float>(E * 2 + 2); w.DP = ar.take<float>(E * 4 + 4);
}

int backward_impl(cns_handle* h, const cns_batch* b, const float* hidden_in, const float* hidden_out,
                  const float* d_vec, const float* d_nrm, const float* d_hidden_out, float* d_hidden_in,
                  float* grad, const void* save, void* workspace, cudaStream_t s) {
  int rc;
#define RUN(expr) if ((rc = (expr)) != CNS_OK) return rc;
  RUN(ensure_bwd_pack(h, s));
  const BwdPack& bp = *reinterpret_cast<BwdPack*>(h->bwd_pack);
  const Packed& pk = h->pk;
  const int64_t C = b->n_clusters, B = b->n_graphs, E = b->n_e01;
  SaveBuffers sv;
  { Arena sa(const_cast<void*>(save)); carve_save(sa, b, sv); }
  Arena ar(workspace);
  BwdBuffers w;
  carve_bwd(ar, b, w);
  auto G = [&](int pid) { return grad + h->raw_off[pid]; };

  // ---- helpers ---------------------------------------------------------------------------------
  auto lin = [&](const float* x, int k, const float* wt, int n_out, const float* bias, const float* residual, int relu,
                 int64_t m, float* y, const float* relu_gate = nullptr, int split_id = -1) {
    LinearArgs la{};
    la.m = m; la.k1 = k; la.k2 = 0; la.n_out = n_out; la.x1 = x; la.wt = wt; la.bias = bias; la.residual = residual;
    la.relu = relu; la.relu_gate = relu_gate; la.y = y;
    // long edge-level products: tensor cores with split operands (fp32 accuracy); short ones stay on the FFMA kernel
    if (split_id >= 0 && m >= SPLIT_MIN_ROWS && !h->no_split_gemm) return gemm_split(la, bp.img(split_id), s);
    return linear(la, s);
  };
  auto dwg = [&](const float* a, int lda, int na, const float* b0, int ldb0, int tiles0, const float* b1, int ldb1, int nb,
                 int64_t m, float* out, int ld_out, float scale, int accumulate, float* colsum_out = nullptr) {
    DwGemmArgs d{};
    d.m = m; d.na = na; d.nb = nb; d.a = a; d.lda = lda; d.b0 = b0; d.ldb0 = ldb0; d.b_tiles0 = tiles0; d.b1 = b1; d.ldb1 = ldb1;
    d.out = out; d.ld_out = ld_out; d.scale = scale; d.accumulate = accumulate; d.partial = w.dw_partial;
    d.max_slabs = (int)(DW_PARTIAL_FLOATS / ((size_t)na * nb));
    d.allow_split = h->no_split_gemm ? 0 : 1;
    d.colsum_out = colsum_out; d.colsum_partial = w.cs_partial;     // bias gradient from the same pass over `a`
    return dw_gemm(d, s);
  };
  auto dws = [&](const float* a, int lda, int n, const float* sm, int lds, int q, int64_t m, float* out, int ld_out,
                 int transposed, int accumulate, float* ones_out = nullptr) {
    DwSmallArgs d{};
    d.m = m; d.n = n; d.q = q; d.a = a; d.lda = lda; d.s = sm; d.lds = lds; d.out = out; d.ld_out = ld_out;
    d.transposed = transposed; d.scale = 1.f; d.accumulate = accumulate; d.partial = w.small_partial; d.max_slabs = w.small_slabs;
    d.ones_out = ones_out;
    return dw_small(d, s);
  };
  auto colsum = [&](const float* a, int lda, int n, int64_t m, float* out, int accumulate) {
    return dws(a, lda, n, nullptr, 0, 1, m, out, 1, 0, accumulate);
  };
  auto ew1 = [&](int op, int64_t n, const float* a, const float* bb, const float* c, const float* d, float* o, float* o2,
                 float* o3) {
    EwArgs e{};
    e.op = op; e.n = n; e.a = a; e.b = bb; e.c = c; e.d = d; e.out = o; e.out2 = o2; e.out3 = o3;
    return ew(e, s);
  };

  CNS_CUDA_CHECK(cudaMemsetAsync(w.err, 0, 64 * sizeof(int), s));
  const float* hin = hidden_in;
  if (!hin) { CNS_CUDA_CHECK(cudaMemsetAsync(w.zeros_h, 0, (size_t)C * H * sizeof(float), s)); hin = w.zeros_h; }

  // CSR by SOURCE of both cluster graphs (out-neighbourhoods): swap the two rows of the edge lists
  {
    const int64_t cs = b->l1_cur_stride ? b->l1_cur_stride : b->n_e11_cur;
    const int64_t ts = b->l1_tar_stride ? b->l1_tar_stride : b->n_e11_tar;
    RUN(csr_by_dst(b->l1_edge_cur + cs, b->l1_edge_cur, b->n_e11_cur, C, w.curT_rowptr, w.curT_dst, w.csr_ws, w.err, s));
    RUN(csr_by_dst(b->l1_edge_tar + ts, b->l1_edge_tar, b->n_e11_tar, C, w.tarT_rowptr, w.tarT_dst, w.csr_ws, w.err, s));
  }

  // ---- decoder (graph_vs.py:254-270) -------------------------------------------------------------
  {
    PoolHeadBwdArgs a{};
    a.n_graphs = B; a.graph_ptr = sv.graph_ptr; a.x = sv.x2; a.cluster_mask = b->cluster_mask; a.gate = sv.gate; a.hid = sv.hid;
    a.dvec = d_vec; a.dnrm = d_nrm; a.pool_w = h->P(P_POOL_W); a.vec0_w = h->P(P_VEC0_W); a.vec1_w = h->P(P_VEC1_W);
    a.nrm0_w = h->P(P_NRM0_W); a.nrm1_w = h->P(P_NRM1_W); a.dx = w.dx2; a.dgate = w.dgate; a.dhid = w.dhid;
    RUN(pool_head_bwd(a, s));
    const float* hid_v = sv.hid; const float* hid_n = sv.hid + H;          // (B, 2, 128): row stride 256
    const float* dhid_v = w.dhid; const float* dhid_n = w.dhid + H;
    if (d_vec) {
      RUN(dws(hid_v, 2 * H, H, d_vec, 6, 6, B, G(P_VEC1_W), H, 1, 1));
      RUN(colsum(d_vec, 6, 6, B, G(P_VEC1_B), 1));
    }
    if (d_nrm) {
      RUN(dws(hid_n, 2 * H, H, d_nrm, 1, 1, B, G(P_NRM1_W), H, 1, 1));
      RUN(colsum(d_nrm, 1, 1, B, G(P_NRM1_B), 1));
    }
    RUN(dwg(dhid_v, 2 * H, H, sv.x_scene, H, 1, nullptr, 0, H, B, G(P_VEC0_W), H, 1.f, 1, G(P_VEC0_B)));
    RUN(dwg(dhid_n, 2 * H, H, sv.x_scene, H, 1, nullptr, 0, H, B, G(P_NRM0_W), H, 1.f, 1, G(P_NRM0_B)));
    RUN(dws(sv.x2, H, H, w.dgate, 1, 1, C, G(P_POOL_W), H, 1, 1));
    RUN(colsum(w.dgate, 1, 1, C, G(P_POOL_B), 1));
  }

  // ---- PERConv backward (graph_vs.py:140-152 + the outer ReLU of Backbone.forward) -----------------
  // y = relu(fc(relu(GN(conv(x)))) + x).  dy -> dx (written to dx_out).
  auto perconv_bwd = [&](const float* dy, const float* y, const float* x, const float* conv, const float* gnb,
                         const float* stats, const int* arg, int msg_w, int msg_b, int bn_w, int bn_b, int bn_ms, int fc_w,
                         int fc_b, const float* pq_bwd, float* dx_out) {
    int r;
    if ((r = ew1(EW_RELU_BWD, C * H, dy, y, nullptr, nullptr, w.dz, nullptr, nullptr))) return r;
    if ((r = lin(w.dz, H, h->P(fc_w), H, nullptr, nullptr, 0, C, w.dgn))) return r;                 // dgn = dz W_fc
    if ((r = dwg(w.dz, H, H, gnb, H, 1, nullptr, 0, H, C, G(fc_w), H, 1.f, 1, G(fc_b)))) return r;
    GraphNormBwdArgs ga{};
    ga.n_graphs = B; ga.graph_ptr = sv.graph_ptr; ga.x = conv; ga.y = gnb; ga.dy = w.dgn; ga.stats = stats;
    ga.weight = h->P(bn_w); ga.mean_scale = h->P(bn_ms); ga.dx = w.dconv; ga.param_partial = w.gn_partial;
    if ((r = graphnorm_relu_bwd(ga, s))) return r;
    if ((r = colsum(w.gn_partial, 3 * H, H, B, G(bn_w), 1))) return r;
    if ((r = colsum(w.gn_partial + H, 3 * H, H, B, G(bn_b), 1))) return r;
    if ((r = colsum(w.gn_partial + 2 * H, 3 * H, H, B, G(bn_ms), 1))) return r;
    EdgeConvBwdArgs ea{};
    ea.n_rows = C; ea.f_out = H; ea.dconv = w.dconv; ea.argmax = arg; ea.out_rowptr = w.curT_rowptr; ea.out_dst = w.curT_dst;
    ea.in_rowptr = sv.cur_rowptr; ea.dpq = w.dpq;
    if ((r = edgeconv_max_bwd(ea, s))) return r;
    if ((r = lin(w.dpq, 2 * H, pq_bwd, H, nullptr, w.dz, 0, C, dx_out))) return r;                  // dx = dPQ W_pq + dz (skip)
    if ((r = dwg(w.dpq, 2 * H, 2 * H, x, H, 1, nullptr, 0, H, C, w.G, H, 1.f, 0, w.gb))) return r;
    if ((r = dws(w.dpq, 2 * H, 2 * H, sv.pos_clu, 2, 2, C, w.Gp, 2, 0, 0))) return r;
    return unfold_edgeconv(w.G, w.Gp, w.gb, H, H, G(msg_w), G(msg_b), s);
  };

  // l1_conv1: input hidden_out, output x2
  RUN(perconv_bwd(w.dx2, sv.x2, hidden_out, sv.conv1, sv.gn1, sv.stats1, sv.arg1, P_C1_MSG_W, P_C1_MSG_B, P_C1_BN_W,
                  P_C1_BN_B, P_C1_BN_MS, P_C1_FC_W, P_C1_FC_B, bp.c1_bwd, w.dh_acc));
  if (d_hidden_out) { RUN(ew1(EW_ADD, C * H, w.dh_acc, d_hidden_out, nullptr, nullptr, w.dh_acc, nullptr, nullptr)); }

  // ---- GRU cell backward (graph_vs.py:178-187) ---------------------------------------------------
  // h' = (1-u) h + u ht ; ht = tanh(conv_k([x1 | h r])) on edge_cur ; (r|u) = sigmoid(conv_g([x1 | h])) on edge_tar
  RUN(ew1(EW_GRU_A, C * H, w.dh_acc, sv.u, sv.ht, hin, w.dconvk, w.du, w.dx2 /* dh direct -> reuse dx2 */));
  float* dh_in = w.dx2;
  {
    EdgeConvBwdArgs ea{};
    ea.n_rows = C; ea.f_out = H; ea.dconv = w.dconvk; ea.argmax = sv.argk; ea.out_rowptr = w.curT_rowptr; ea.out_dst = w.curT_dst;
    ea.in_rowptr = sv.cur_rowptr; ea.dpq = w.dpq;
    RUN(edgeconv_max_bwd(ea, s));
    RUN(lin(w.dpq, 2 * H, bp.k_bwd, 2 * H, nullptr, nullptr, 0, C, w.dxk));                        // (dx1_k | dhr)
    RUN(dwg(w.dpq, 2 * H, 2 * H, sv.x1, H, 1, sv.hr, H, 2 * H, C, w.G, 2 * H, 1.f, 0, w.gb));
    RUN(dws(w.dpq, 2 * H, 2 * H, sv.pos_clu, 2, 2, C, w.Gp, 2, 0, 0));
    RUN(unfold_edgeconv(w.G, w.Gp, w.gb, 2 * H, H, G(P_K_MSG_W), G(P_K_MSG_B), s));
  }
  RUN(gru_b(w.dxk, hin, sv.g, w.du, w.dconvg, dh_in, C, s));
  {
    EdgeConvBwdArgs ea{};
    ea.n_rows = C; ea.f_out = 2 * H; ea.dconv = w.dconvg; ea.argmax = sv.argg; ea.out_rowptr = w.tarT_rowptr; ea.out_dst = w.tarT_dst;
    ea.in_rowptr = sv.tar_rowptr; ea.dpq = w.dpq;
    RUN(edgeconv_max_bwd(ea, s));
    RUN(lin(w.dpq, 4 * H, bp.g_bwd, 2 * H, nullptr, nullptr, 0, C, w.dxg));                        // (dx1_g | dh_g)
    RUN(dwg(w.dpq, 4 * H, 4 * H, sv.x1, H, 1, hin, H, 2 * H, C, w.G, 2 * H, 1.f, 0, w.gb));
    RUN(dws(w.dpq, 4 * H, 4 * H, sv.pos_clu, 2, 2, C, w.Gp, 2, 0, 0));
    RUN(unfold_edgeconv(w.G, w.Gp, w.gb, 2 * H, 2 * H, G(P_G_MSG_W), G(P_G_MSG_B), s));
  }
  if (d_hidden_in) {
    RUN(add_slices(dh_in, H, w.dxg + H, 2 * H, d_hidden_in, C, 0, s));
  }
  RUN(add_slices(w.dxk, 2 * H, w.dxg, 2 * H, w.dx1, C, 0, s));                                     // dx1 = dx1_k + dx1_g

  // l1_conv0: input x_clu, output x1
  RUN(perconv_bwd(w.dx1, sv.x1, sv.x_clu, sv.conv0, sv.gn0, sv.stats0, sv.arg0, P_C0_MSG_W, P_C0_MSG_B, P_C0_BN_W,
                  P_C0_BN_B, P_C0_BN_MS, P_C0_FC_W, P_C0_FC_B, bp.c0_bwd, w.dxclu));

  // ---- encoder tail: x_clu = relu(out_enc([cur | tar - cur])) * mask (graph_vs.py:229-230) --------
  RUN(out_enc_bwd(0, C, sv.clu_feat, w.dxclu, sv.x_clu, b->cluster_mask, w.cat, w.dpre, nullptr, nullptr, s));
  RUN(lin(w.dpre, H, h->P(P_OUT_W), 2 * H, nullptr, nullptr, 0, C, w.dcat));
  RUN(dwg(w.dpre, H, H, w.cat, 2 * H, 2, nullptr, 0, 2 * H, C, G(P_OUT_W), 2 * H, 1.f, 1, G(P_OUT_B)));
  RUN(out_enc_bwd(1, C, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, w.dcat, w.dclu, s));

  // ---- keypoint -> cluster attention, both sides (graph_vs.py:204-214) -----------------------------
  // cluster-side operands: xt_c = relu(in_enc(x_tar[centre])), a_dst = xt_c lin_dst^T
  {
    EncGatherArgs ga{};
    ga.n_rows = C; ga.centers = b->cluster_centers; ga.feat = b->x_tar; ga.w_in = h->P(P_IN_W); ga.b_in = h->P(P_IN_B);
    ga.X = w.xtc; ga.FE = w.fe_c;
    RUN(enc_gather(ga, s));
    RUN(lin(w.xtc, H, pk.lin_dst_t, H, nullptr, nullptr, 0, C, w.a_dst));
  }
  float* X = w.E[0]; float* HP = w.E[1]; float* D = w.E[2]; float* T = w.E[3]; float* U = w.E[4]; float* L = w.E[5];
  float* V = w.E[6]; float* dL = w.E[7]; float* dV = w.E[8]; float* dT = w.E[9]; float* dX = w.E[10]; float* tmp = w.E[11];
  for (int side = 0; side < 2; ++side) {
    const float* feat = side == 0 ? b->x_cur : b->x_tar;
    const float* pos_src = side == 0 ? b->pos_cur : b->pos_tar;
    EncGatherArgs ga{};
    ga.n_rows = E; ga.ej = b->l0_to_l1_j; ga.ei = b->l0_to_l1_i; ga.feat = feat; ga.pos_src = pos_src; ga.pos_tar = b->pos_tar;
    ga.pos_clu = sv.pos_clu; ga.w_in = h->P(P_IN_W); ga.b_in = h->P(P_IN_B); ga.w_p0 = h->P(P_POS0_W); ga.b_p0 = h->P(P_POS0_B);
    ga.X = X; ga.HP = HP; ga.FE = w.FE; ga.DP = w.DP;
    RUN(enc_gather(ga, s));
    // forward recompute
    RUN(lin(HP, H, pk.enc_chain + 0 * 16384, H, h->P(P_POS1_B), nullptr, 0, E, D, nullptr, 0));                 // delta
    RUN(lin(X, H, pk.enc_chain + 1 * 16384, H, nullptr, D, 0, E, T, nullptr, 1));                              // -a_src + delta
    RUN(add_rows_gather(T, w.a_dst, b->l0_to_l1_i, E, s));                                          // + a_dst[i]
    RUN(lin(T, H, pk.enc_chain + 3 * 16384, H, h->P(P_ATT0_B), nullptr, 1, E, U, nullptr, 3));
    RUN(lin(U, H, pk.enc_chain + 4 * 16384, H, h->P(P_ATT1_B), nullptr, 0, E, L, nullptr, 4));
    RUN(lin(X, H, pk.enc_chain + 2 * 16384, H, nullptr, D, 0, E, V, nullptr, 2));                              // v + delta
    // softmax / aggregation / relu backward
    EncSoftmaxBwdArgs sa{};
    sa.n_clusters = C; sa.clu_rowptr = sv.clu_rowptr; sa.L = L; sa.V = V; sa.y = sv.clu_feat + (size_t)side * C * H;
    sa.dy = w.dclu + (size_t)side * C * H; sa.dL = dL; sa.dV = dV;
    RUN(enc_softmax_bwd(sa, s));
    // attn_nn
    RUN(dwg(dL, H, H, U, H, 1, nullptr, 0, H, E, G(P_ATT1_W), H, 1.f, 1, G(P_ATT1_B)));
    RUN(lin(dL, H, h->P(P_ATT1_W), H, nullptr, nullptr, 0, E, tmp, U, 5));                            // dU_pre = (U > 0) dL A1
    RUN(dwg(tmp, H, H, T, H, 1, nullptr, 0, H, E, G(P_ATT0_W), H, 1.f, 1, G(P_ATT0_B)));
    RUN(lin(tmp, H, h->P(P_ATT0_W), H, nullptr, nullptr, 0, E, dT, nullptr, 6));                               // dT
    // t = a_dst[i] - a_src + delta
    RUN(seg_sum_rows(dT, sv.clu_rowptr, C, w.d_adst, side, s));                                     // both sides accumulate
    RUN(dwg(dT, H, H, X, H, 1, nullptr, 0, H, E, G(P_LINSRC_W), H, -1.f, 1));                      // a_src = x lin_src^T
    RUN(dwg(dV, H, H, X, H, 1, nullptr, 0, H, E, G(P_LIN_W), H, 1.f, 1));                          // v = x lin^T
    RUN(lin(dV, H, h->P(P_LIN_W), H, nullptr, nullptr, 0, E, dX, nullptr, 7));
    RUN(lin(dT, H, bp.neg_lin_src, H, nullptr, dX, 0, E, dX, X, 8));                                  // dX = (X > 0)(dV lin - dT lin_src)
    RUN(ew1(EW_ADD, E * H, dT, dV, nullptr, nullptr, tmp, nullptr, nullptr));                      // dDelta
    RUN(dwg(tmp, H, H, HP, H, 1, nullptr, 0, H, E, G(P_POS1_W), H, 1.f, 1, G(P_POS1_B)));
    RUN(lin(tmp, H, h->P(P_POS1_W), H, nullptr, nullptr, 0, E, dL, HP, 9));                           // dHP_pre (reuse dL)
    RUN(dws(dL, H, H, w.DP, 4, 4, E, G(P_POS0_W), 4, 0, 1, G(P_POS0_B)));
    RUN(dws(dX, H, H, w.FE, 2, 2, E, G(P_IN_W), 2, 0, 1, G(P_IN_B)));
  }
  // cluster side: a_dst = xt_c lin_dst^T
  RUN(dwg(w.d_adst, H, H, w.xtc, H, 1, nullptr, 0, H, C, G(P_LINDST_W), H, 1.f, 1));
  RUN(lin(w.d_adst, H, h->P(P_LINDST_W), H, nullptr, nullptr, 0, C, w.d_xtc, w.xtc));
  RUN(dws(w.d_xtc, H, H, w.fe_c, 2, 2, C, G(P_IN_W), 2, 0, 1, G(P_IN_B)));
#undef RUN
  return CNS_OK;
}

}  // namespace cns

using namespace cns;

extern "C" {

size_t cns_backward_workspace_bytes(const cns_batch* b) {
  if (!b) return 0;
  Arena ar(nullptr);
  BwdBuffers w;
  carve_bwd(ar, b, w);
  return ar.off + 256;
}

int64_t cns_param_offset(const cns_handle* h, int k) {
  if (!h || k < 0 || k > P_COUNT) return -1;
  return (int64_t)h->raw_off[k];
}

int cns_graphvs_backward(cns_handle* h, const cns_batch* b, const float* hidden_in, const float* hidden_out,
                         const float* d_vel_si_vec, const float* d_vel_si_norm, const float* d_hidden_out,
                         float* d_hidden_in, float* grad_flat, const void* save, size_t save_bytes, void* workspace,
                         size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(h && b && grad_flat && save && workspace && hidden_out, "NULL argument");
  CNS_REQUIRE(h->weights_loaded, "weights not loaded");
  if (b->n_graphs == 0) return CNS_OK;
  if (save_bytes < cns_forward_save_bytes(b) || workspace_bytes < cns_backward_workspace_bytes(b)) {
    set_error("cns_graphvs_backward: save/workspace buffer too small");
    return CNS_ERR_WORKSPACE_TOO_SMALL;
  }
  CNS_CUDA_CHECK(cudaSetDevice(h->device));
  return backward_impl(h, b, hidden_in, hidden_out, d_vel_si_vec, d_vel_si_norm, d_hidden_out, d_hidden_in, grad_flat,
                       save, workspace, (cudaStream_t)stream);
}

}  // extern "C"
