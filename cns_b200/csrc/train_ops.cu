// Training-side fused ops:
//   cns_objectives      : loss of cns/models/functions.py:41-65 (1 - cos direction loss + MSE on the inverse-ELU
//                         of |gt|) forward AND its gradient w.r.t. (vel_si_vec, vel_si_norm) in one launch,
//                         one device->host-free result triple (the reference does three .item() syncs).
//   cns_clip_adamw      : global-norm clipping (trainer.py:97, max_norm 10) + both AdamW groups
//                         (train_cns.py:65-69: decayed Linear weights / un-decayed biases + GraphNorm) over the
//                         flat parameter / gradient buffers, after the data-parallel all-reduce.
// Deterministic: fixed-order block reductions, no floating-point atomics.
#include "common.cuh"
#include "kernels.cuh"

namespace cns {

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// single block; out[0] = dir_loss, out[1] = norm_loss, out[2] = total
__global__ void __launch_bounds__(256)
objectives_kernel(const float* __restrict__ vec, const float* __restrict__ nrm, const float* __restrict__ gt, int64_t nb,
                  float grad_scale, float* __restrict__ out, float* __restrict__ dvec, float* __restrict__ dnrm) {
  __shared__ float red[8];
  float dir = 0.f, nl = 0.f;
  const float inv_b = 1.f / (float)nb;
  for (int64_t g = threadIdx.x; g < nb; g += blockDim.x) {
    float v[6], t[6];
    float vv = 0.f, tt = 0.f, vt = 0.f;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      v[q] = vec[g * 6 + q]; t[q] = gt[g * 6 + q];
      vv = fmaf(v[q], v[q], vv); tt = fmaf(t[q], t[q], tt); vt = fmaf(v[q], t[q], vt);
    }
    // F.cosine_similarity: x.y / (max(|x|, eps) * max(|y|, eps)), eps = 1e-8
    const float nv = fmaxf(sqrtf(vv), 1e-8f), nt = fmaxf(sqrtf(tt), 1e-8f);
    const float cosv = vt / (nv * nt);
    dir += 1.f - cosv;
    // reverse_alpha_p_elu(|gt|, eps = 1e-5): y < 1 ? log(y + eps) : y - 1
    const float gn = sqrtf(tt);
    const float target = gn < 1.f ? logf(gn + 1e-5f) : gn - 1.f;
    const float diff = nrm[g] - target;
    nl = fmaf(diff, diff, nl);
    if (dvec) {
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        // d(1 - cos)/dv = -(t / (|v||t|) - cos * v / |v|^2)
        const float d = -(t[q] / (nv * nt) - cosv * v[q] / (nv * nv));
        dvec[g * 6 + q] = grad_scale * inv_b * d;
      }
      dnrm[g] = grad_scale * inv_b * 2.f * diff;
    }
  }
  dir = block_sum(dir, red);
  nl = block_sum(nl, red);
  if (threadIdx.x == 0) {
    out[0] = dir * inv_b;
    out[1] = nl * inv_b;
    out[2] = (dir + nl) * inv_b;
  }
}

// GraphVS.postprocess on (B,6) predictions (functions.py:26-38): unit direction x (1 + ELU(norm)), translation x tPo_norm.
// (the inference forward emits this fused from the pool/head kernel; this is the stand-alone form for training loops)
__global__ void __launch_bounds__(256)
postprocess_kernel(const float* __restrict__ vec, const float* __restrict__ nrm, const float* __restrict__ tpo, int64_t nb,
                   float* __restrict__ vel) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nb) return;
  float v[6], vv = 0.f;
#pragma unroll
  for (int q = 0; q < 6; ++q) { v[q] = vec[g * 6 + q]; vv = fmaf(v[q], v[q], vv); }
  float scale = 1.f;
  if (nrm) {
    const float n = nrm[g];
    scale = (1.f + (n > 0.f ? n : expm1f(n))) / (sqrtf(vv) + 1e-7f);
  }
  const float t = tpo ? tpo[g] : 1.f;
#pragma unroll
  for (int q = 0; q < 6; ++q) vel[g * 6 + q] = v[q] * scale * (q < 3 ? t : 1.f);
}

// ---- clip + AdamW -------------------------------------------------------------------------------
constexpr int NORM_BLOCKS = 256;

__global__ void __launch_bounds__(256)
sq_norm_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fmaf(g[i], g[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// one launch: every block first reduces the NORM_BLOCKS partials (fixed order) to the clip coefficient, then
// applies torch.optim.AdamW's update (decoupled weight decay) to its slice.
__global__ void __launch_bounds__(256)
clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  const unsigned char* __restrict__ decay_mask, int64_t n, const float* __restrict__ partial,
                  float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2,
                  float grad_div, float* __restrict__ norm_out) {
  __shared__ float coef_s;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < NORM_BLOCKS; ++k) s += partial[k];
    const float total = sqrtf(s) / grad_div;
    const float c = max_norm / (total + 1e-6f);            // torch.nn.utils.clip_grad_norm_
    coef_s = (c < 1.f ? c : 1.f) / grad_div;
    if (blockIdx.x == 0 && norm_out) *norm_out = total;
  }
  __syncthreads();
  const float coef = coef_s;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i];
    if (decay_mask[i]) pi *= 1.f - lr * weight_decay;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * mi / denom;
  }
}

}  // namespace cns

using namespace cns;

extern "C" {

int cns_objectives(const float* vel_si_vec, const float* vel_si_norm, const float* vel_si_gt, int64_t n_graphs,
                   float grad_scale, float* loss_out3, float* d_vec, float* d_norm, void* stream) {
  CNS_REQUIRE(vel_si_vec && vel_si_norm && vel_si_gt && loss_out3, "NULL pointer");
  CNS_REQUIRE((d_vec == nullptr) == (d_norm == nullptr), "give both gradient buffers or neither");
  CNS_REQUIRE(n_graphs > 0, "empty batch");
  objectives_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(vel_si_vec, vel_si_norm, vel_si_gt, n_graphs, grad_scale, loss_out3,
                                                         d_vec, d_norm);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int cns_postprocess(const float* vel_si_vec, const float* vel_si_norm, const float* tPo_norm, int64_t n_graphs,
                    float* vel, void* stream) {
  CNS_REQUIRE(vel_si_vec && vel, "NULL pointer");
  if (n_graphs <= 0) return CNS_OK;
  postprocess_kernel<<<ceil_div(n_graphs, 256), 256, 0, (cudaStream_t)stream>>>(vel_si_vec, vel_si_norm, tPo_norm, n_graphs, vel);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

size_t cns_clip_adamw_workspace_bytes(void) { return NORM_BLOCKS * sizeof(float) + 256; }

// params / grads / exp_avg / exp_avg_sq: flat fp32 buffers of n elements; decay_mask (n) u8 = 1 where weight decay
// applies.  `step` is the 1-based update count; `grad_div` divides the gradients first (world size after a summing
// all-reduce, 1 otherwise).  grad_norm_out (device, optional) receives the pre-clip global norm.
int cns_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const unsigned char* decay_mask,
                   int64_t n, float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                   int64_t step, float grad_div, float* grad_norm_out, void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(params && grads && exp_avg && exp_avg_sq && decay_mask && workspace, "NULL pointer");
  CNS_REQUIRE(n > 0 && step >= 1 && grad_div > 0.f, "bad argument");
  if (workspace_bytes < cns_clip_adamw_workspace_bytes()) { set_error("cns_clip_adamw: workspace too small"); return CNS_ERR_WORKSPACE_TOO_SMALL; }
  cudaStream_t s = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  sq_norm_partial_kernel<<<NORM_BLOCKS, 256, 0, s>>>(grads, n, partial);
  CNS_LAUNCH_CHECK();
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  clip_adamw_kernel<<<296, 256, 0, s>>>(params, grads, exp_avg, exp_avg_sq, decay_mask, n, partial, max_norm, lr, beta1, beta2,
                                        eps, weight_decay, bc1, bc2, grad_div, grad_norm_out);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // extern "C"
