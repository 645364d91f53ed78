// Launch interfaces of the backward-pass kernels (backward_kernels.cu), used by backward.cu.
#pragma once
#include "common.cuh"

namespace cns {

// out[na][nb] (+)= scale * A^T B over m rows;  B columns may come from two matrices (concat inputs)
struct DwGemmArgs {
  int64_t m; int na, nb;
  const float* a; int lda;
  const float* b0; int ldb0; int b_tiles0;   // first b_tiles0 128-column tiles come from b0, the rest from b1
  const float* b1; int ldb1;
  float* out; int ld_out; float scale; int accumulate;
  float* partial; int max_slabs;
  int allow_split;                            // 128 x 128 products over many rows may run on the tensor cores (gemm_split.cu)
  float* colsum_out; float* colsum_partial;   // optional: colsum_out[i] (+)= sum_m a[m][i] from the same pass (bias gradient)
};
int dw_gemm(const DwGemmArgs& a, cudaStream_t s);

struct DwSmallArgs {
  int64_t m; int n, q;
  const float* a; int lda;
  const float* s; int lds;       // (m, q) or NULL (ones)
  float* out; int ld_out; int transposed; float scale; int accumulate;
  float* partial; int max_slabs;
  float* ones_out;               // optional (needs s, q <= 7): ones_out[n] (+)= sum_m a[m][n] from the same pass (bias gradient)
};
int dw_small(const DwSmallArgs& a, cudaStream_t s);

struct EdgeConvBwdArgs {
  int64_t n_rows; int f_out;
  const float* dconv; const int* argmax;
  const int* out_rowptr; const int* out_dst;   // CSR by SOURCE
  const int* in_rowptr;                        // CSR by destination (deg_in)
  float* dpq;                                  // (n, 2 f)
};
int edgeconv_max_bwd(const EdgeConvBwdArgs& a, cudaStream_t s);

struct GraphNormBwdArgs {
  int64_t n_graphs; const int* graph_ptr;
  const float* x; const float* y; const float* dy; const float* stats;
  const float* weight; const float* mean_scale;
  float* dx; float* param_partial;             // (B, 3, 128): dweight, dbias, dmean_scale per graph
};
int graphnorm_relu_bwd(const GraphNormBwdArgs& a, cudaStream_t s);

struct PoolHeadBwdArgs {
  int64_t n_graphs; const int* graph_ptr; const float* x; const uint8_t* cluster_mask;
  const float* gate; const float* hid;
  const float* dvec; const float* dnrm;
  const float* pool_w; const float* vec0_w; const float* vec1_w; const float* nrm0_w; const float* nrm1_w;
  float* dx; float* dgate; float* dhid;
};
int pool_head_bwd(const PoolHeadBwdArgs& a, cudaStream_t s);

enum EwOp { EW_RELU_BWD = 0, EW_RELU_BWD_ADD = 1, EW_ADD = 2, EW_GRU_A = 3 };
struct EwArgs {
  int op; int64_t n;
  const float* a; const float* b; const float* c; const float* d;
  float* out; float* out2; float* out3;
};
int ew(const EwArgs& a, cudaStream_t s);
int gru_b(const float* dxk, const float* h, const float* g, const float* du, float* dconvg, float* dh_acc,
          int64_t n_rows, cudaStream_t s);
int add_slices(const float* a, int lda, const float* b, int ldb, float* out, int64_t n_rows, int accumulate,
               cudaStream_t s);
int unfold_edgeconv(const float* G, const float* Gp, const float* gb, int f_in, int f_out, float* dW, float* db,
                    cudaStream_t s);

struct EncGatherArgs {
  int64_t n_rows;
  const int64_t* ej; const int64_t* ei;     // edge mode; or
  const int64_t* centers;                   // cluster mode (ej == NULL): rows = clusters, feat[centers[c]]
  const float* feat; const float* pos_src; const float* pos_tar; const float* pos_clu;
  const float* w_in; const float* b_in; const float* w_p0; const float* b_p0;
  float* X; float* HP; float* FE; float* DP;   // HP / DP optional
};
int enc_gather(const EncGatherArgs& a, cudaStream_t s);
int add_rows_gather(float* t, const float* rows, const int64_t* ei, int64_t n_edges, cudaStream_t s);

struct EncSoftmaxBwdArgs {
  int64_t n_clusters; const int* clu_rowptr;
  const float* L; const float* V; const float* y; const float* dy;
  float* dL; float* dV;
};
int enc_softmax_bwd(const EncSoftmaxBwdArgs& a, cudaStream_t s);
int seg_sum_rows(const float* rows, const int* rowptr, int64_t n_clusters, float* out, int accumulate, cudaStream_t s);
int out_enc_bwd(int mode, int64_t n_clusters, const float* clu_feat, const float* dx, const float* x_clu,
                const uint8_t* mask, float* cat, float* dpre, const float* dcat, float* dclu, cudaStream_t s);
int transpose(const float* src, int rows, int cols, float sign, float* dst, cudaStream_t s);

}  // namespace cns
