// The two steps either side of the graph path in the servo loop, on the device:
//   cns_align_matches    : Classic._align_cur_to_tar (cns/frontend/classic.py:142-148) + pixel -> normalized camera
//                          plane (cns/utils/perception.py:66-68) + the float32 cast of graph_gen.py:257-263
//   cns_pixel_stop_error : PixelStopPolicy.calculate_error (cns/benchmark/stop_policy.py:86-126) for B scenes,
//                          weights / witness counts resident on the device
// Everything is written with explicitly rounded float32 operations in numpy's order (percentile "linear" method
// with a float32 virtual index, pairwise summation), so the error is bit-identical with the reference under
// numpy 2.x, not merely close: the stop decision compares it with thresholds.
#include "common.cuh"

namespace cns {
namespace {

// ------------------------------------------------------------------------------------------------------------
// match alignment
// ------------------------------------------------------------------------------------------------------------
__global__ void align_last_kernel(const int64_t* __restrict__ matches, int n_match, int n_tar, int n_cur,
                                  int* __restrict__ last, int* __restrict__ err) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_match) return;
  const int64_t t = matches[2 * m], c = matches[2 * m + 1];
  if (t < 0 || t >= n_tar || c < 0 || c >= n_cur) { *err = 1; return; }
  atomicMax(&last[t], m);                    // numpy fancy assignment: the last occurrence of an index wins
}

__global__ void align_emit_kernel(const float2* __restrict__ cur_pos, const int64_t* __restrict__ matches,
                                  const int* __restrict__ last, int n_tar, double fx, double fy, double cx, double cy,
                                  float2* __restrict__ xy, uint8_t* __restrict__ valid) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tar) return;
  const int m = last[t];
  float2 px = make_float2(0.f, 0.f);         // unmatched rows stay at pixel (0, 0) like np.zeros_like(tar_pos)
  if (m >= 0) px = cur_pos[matches[2 * m + 1]];
  xy[t] = make_float2(__double2float_rn(__ddiv_rn(__dsub_rn((double)px.x, cx), fx)),
                      __double2float_rn(__ddiv_rn(__dsub_rn((double)px.y, cy), fy)));
  valid[t] = m >= 0;
}

// ------------------------------------------------------------------------------------------------------------
// pixel stop error
// ------------------------------------------------------------------------------------------------------------
constexpr int PS_THREADS = 512;
constexpr int PS_SMEM_ROWS = 4096;           // scenes up to this many keypoints keep the kept (w, delta) pairs in smem
constexpr unsigned PS_INVALID = 0xFFFFFFFFu; // key of a masked keypoint: sorts after every finite error

// numpy's float32 add.reduce of a contiguous vector (pairwise_sum, numpy/core/src/umath/loops_utils.h): blocks of
// <= 128 elements with 8 running sums, halves split at a multiple of 8.  One thread, explicit stack.
__device__ float np_sum_leaf(const float* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

__device__ float np_sum_f32(const float* a, int n) {
  struct Frame { int beg, n, stage; float left; };
  Frame st[32];
  int sp = 0;
  float ret = 0.f;
  st[sp++] = Frame{0, n, 0, 0.f};
  while (sp > 0) {
    Frame& f = st[sp - 1];
    if (f.n <= 128) { ret = np_sum_leaf(a + f.beg, f.n); --sp; continue; }
    int n2 = f.n / 2;
    n2 -= n2 % 8;
    if (f.stage == 0) { f.stage = 1; st[sp++] = Frame{f.beg, n2, 0, 0.f}; }
    else if (f.stage == 1) { f.left = ret; f.stage = 2; st[sp++] = Frame{f.beg + n2, f.n - n2, 0, 0.f}; }
    else { ret = __fadd_rn(f.left, ret); --sp; }
  }
  return ret;
}

// k-th smallest (0-based) of the keys of one scene: 4 passes of an 8-bit radix histogram in shared memory
__device__ unsigned select_kth(const unsigned* __restrict__ keys, int n, int k, int* hist, int* bcast) {
  unsigned prefix = 0, known = 0;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned key = keys[i];
      if ((key & known) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, b = 0;
      for (; b < 255; ++b) {
        if (acc + hist[b] > k) break;
        acc += hist[b];
      }
      bcast[0] = b;
      bcast[1] = k - acc;
    }
    __syncthreads();
    prefix |= (unsigned)bcast[0] << shift;
    known |= 255u << shift;
    k = bcast[1];
    __syncthreads();
  }
  return prefix;
}

struct PixelStopArgs {
  const float2* pos_cur; const float2* pos_tar; const uint8_t* mask; const int32_t* node_ptr; const uint8_t* reset;
  int n_nodes; float decay, maximum;
  float* weight; int32_t* witness;
  unsigned* keys; float* kept_w; float* kept_d;      // workspace, (n_nodes,) each
  float* err_out; int32_t* n_valid_out;
};

__global__ void __launch_bounds__(PS_THREADS) pixel_stop_kernel(PixelStopArgs p) {
  __shared__ int hist[256];
  __shared__ int bcast[2];
  __shared__ int warp_tot[PS_THREADS / 32];
  __shared__ int carry;
  __shared__ float sh_scalar;
  __shared__ float sm_w[PS_SMEM_ROWS];
  __shared__ float sm_d[PS_SMEM_ROWS];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int beg = p.node_ptr ? p.node_ptr[g] : 0;
  const int end = p.node_ptr ? p.node_ptr[g + 1] : p.n_nodes;
  const int n = end - beg;
  unsigned* keys = p.keys + beg;

  // 1. witness counts and weights (stop_policy.py:103-110), weighted error per visible keypoint (:112-115)
  const bool fresh = p.reset && p.reset[g];           // PixelStopPolicy.reset (:96-100) folded into the next frame
  int mine = 0;
  for (int i = tid; i < n; i += PS_THREADS) {
    const int r = beg + i;
    const bool m = p.mask[r] != 0;
    const int wc = fresh ? 0 : p.witness[r];
    p.witness[r] = wc + (m ? 1 : 0);
    const float w_old = fresh ? 0.f : p.weight[r];
    const float w = (m && wc == 0) ? p.maximum : __fadd_rn(__fmul_rn(p.decay, w_old), m ? 1.f : 0.f);
    p.weight[r] = w;
    unsigned key = PS_INVALID;
    if (m) {
      const float2 a = p.pos_cur[r], b = p.pos_tar[r];
      const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      key = __float_as_uint(__fmul_rn(w, d));        // w, d >= 0: the bit pattern orders like the value
      p.kept_d[r] = d;
      ++mine;
    }
    keys[i] = key;
  }
  if (tid == 0) carry = 0;
  __syncthreads();
  atomicAdd(&carry, mine);                            // integer count: order does not matter
  __syncthreads();
  const int n_valid = carry;
  __syncthreads();
  if (tid == 0 && p.n_valid_out) p.n_valid_out[g] = n_valid;
  if (n_valid == 0) {                                 // the reference cannot take a percentile of nothing
    if (tid == 0) p.err_out[g] = __int_as_float(0x7fc00000);
    return;
  }

  // 2. 90th percentile, numpy "linear" method with a float32 virtual index (np.percentile of a float32 array)
  const float vidx = __fmul_rn((float)(n_valid - 1), 0.9f);
  float pct;
  if (vidx >= (float)(n_valid - 1)) {
    pct = __uint_as_float(select_kth(keys, n, n_valid - 1, hist, bcast));
  } else {
    const int lo = (int)floorf(vidx);
    const float gam = __fsub_rn(vidx, (float)lo);
    const float a = __uint_as_float(select_kth(keys, n, lo, hist, bcast));
    const float b = __uint_as_float(select_kth(keys, n, lo + 1, hist, bcast));
    const float d = __fsub_rn(b, a);
    pct = gam >= 0.5f ? __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.f, gam))) : __fadd_rn(a, __fmul_rn(d, gam));
  }

  // 3. keep the keypoints below it, in keypoint order (boolean indexing)
  const bool in_smem = n <= PS_SMEM_ROWS;
  float* kw = in_smem ? sm_w : p.kept_w + beg;
  float* kd = in_smem ? sm_d : p.kept_d + beg;       // compaction moves entries to lower or equal positions only
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += PS_THREADS) {
    const int i = base + tid;
    bool keep = false;
    float w = 0.f, d = 0.f;
    if (i < n) {
      const unsigned key = keys[i];
      keep = key != PS_INVALID && __uint_as_float(key) < pct;
      if (keep) { w = p.weight[beg + i]; d = p.kept_d[beg + i]; }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();                                  // also orders the reads of kept_d above before the writes below
    int off = carry;
    for (int wq = 0; wq < warp; ++wq) off += warp_tot[wq];
    if (keep) {
      const int dst = off + __popc(bal & ((1u << lane) - 1));
      kw[dst] = w;
      kd[dst] = d;
    }
    __syncthreads();
    if (tid == PS_THREADS - 1) carry = off + __popc(bal);
    __syncthreads();
  }
  const int n_keep = carry;

  // 4. weights normalised by their float32 sum + 1e-7, error = float32 sum of w * delta (:117-121)
  if (tid == 0) sh_scalar = __fadd_rn(np_sum_f32(kw, n_keep), 1e-7f);
  __syncthreads();
  const float denom = sh_scalar;
  for (int i = tid; i < n_keep; i += PS_THREADS) kw[i] = __fmul_rn(__fdiv_rn(kw[i], denom), kd[i]);
  __syncthreads();
  if (tid == 0) p.err_out[g] = np_sum_f32(kw, n_keep);
}

}  // namespace
}  // namespace cns

using namespace cns;

extern "C" {

int cns_align_matches(const float* cur_pos, int64_t n_cur, const int64_t* matches, int64_t n_matches, int64_t n_tar,
                      double fx, double fy, double cx, double cy, float* cur_xy, uint8_t* valid,
                      void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(n_cur >= 0 && n_matches >= 0 && n_tar >= 0 && n_tar < (1ll << 31) && n_matches < (1ll << 31) &&
              n_cur < (1ll << 31), "bad size");
  CNS_REQUIRE(cur_xy && valid && workspace && (n_matches == 0 || (cur_pos && matches)), "NULL pointer");
  CNS_REQUIRE(workspace_bytes >= (size_t)(n_tar + 1) * sizeof(int), "cns_align_matches: workspace too small");
  CNS_REQUIRE(fx != 0.0 && fy != 0.0, "zero focal length");
  if (n_tar == 0) return CNS_OK;
  cudaStream_t s = (cudaStream_t)stream;
  int* last = (int*)workspace;
  int* err = last + n_tar;
  CNS_CUDA_CHECK(cudaMemsetAsync(last, 0xFF, (size_t)n_tar * sizeof(int), s));     // -1
  CNS_CUDA_CHECK(cudaMemsetAsync(err, 0, sizeof(int), s));
  if (n_matches > 0) {
    align_last_kernel<<<ceil_div(n_matches, 256), 256, 0, s>>>(matches, (int)n_matches, (int)n_tar, (int)n_cur, last, err);
    CNS_LAUNCH_CHECK();
  }
  align_emit_kernel<<<ceil_div(n_tar, 256), 256, 0, s>>>((const float2*)cur_pos, matches, last, (int)n_tar, fx, fy, cx, cy,
                                                         (float2*)cur_xy, valid);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

size_t cns_pixel_stop_workspace_bytes(int64_t n_nodes) { return align_up((size_t)(n_nodes > 0 ? n_nodes : 1) * 12); }

int cns_pixel_stop_error(const float* pos_cur, const float* pos_tar, const uint8_t* node_mask, const int32_t* node_ptr,
                         const uint8_t* reset_flags, int64_t n_nodes, int64_t n_graphs, float decay, float maximum,
                         float* weight, int32_t* witness,
                         float* err_out, int32_t* n_valid_out, void* workspace, size_t workspace_bytes, void* stream) {
  CNS_REQUIRE(n_nodes >= 0 && n_graphs >= 0 && n_nodes < (1ll << 31) && n_graphs < (1ll << 31), "bad size");
  if (n_graphs == 0) return CNS_OK;
  CNS_REQUIRE(pos_cur && pos_tar && node_mask && weight && witness && err_out && workspace, "NULL pointer");
  CNS_REQUIRE(node_ptr || n_graphs == 1, "node_ptr is required for more than one graph");
  CNS_REQUIRE(workspace_bytes >= cns_pixel_stop_workspace_bytes(n_nodes), "cns_pixel_stop_error: workspace too small");
  PixelStopArgs a;
  a.pos_cur = (const float2*)pos_cur; a.pos_tar = (const float2*)pos_tar; a.mask = node_mask; a.node_ptr = node_ptr; a.reset = reset_flags;
  a.n_nodes = (int)n_nodes; a.decay = decay; a.maximum = maximum; a.weight = weight; a.witness = witness;
  a.keys = (unsigned*)workspace;
  a.kept_w = (float*)workspace + n_nodes;
  a.kept_d = (float*)workspace + 2 * n_nodes;
  a.err_out = err_out; a.n_valid_out = n_valid_out;
  pixel_stop_kernel<<<(int)n_graphs, PS_THREADS, 0, (cudaStream_t)stream>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // extern "C"
