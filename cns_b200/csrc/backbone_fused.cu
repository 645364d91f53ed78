// Fused cluster-level backbone (tcgen05): out_enc -> PERConv -> PEConvGRUCell -> PERConv in ONE kernel
// (cns/models/graph_vs.py:115-190,229-251), for batches whose cluster graphs are complete digraphs
// (CNS_BATCH_L1_*_COMPLETE) with at most 128 clusters per graph.
//
// With complete graphs the only coupling between cluster rows is per graph: the column max of the
// PointEdgeConv messages and the GraphNorm statistics.  A CTA therefore owns a tile of WHOLE graphs
// (<= 128 cluster rows = the 128 TMEM lanes) and walks the GEMMs of the chain back to back:
//
//   [cur | tar - cur] --out_enc--> x = relu(.) * cluster_mask
//   x --c0 [P|Q]--> max/GraphNorm/ReLU --c0.fc + x--> y0
//   [y0|h] --gates [P|Q]--> max/sigmoid -> r, u ; [y0|h*r] --candidate [P|Q]--> max/tanh -> h' = (1-u) h + u c
//   h' --c1 [P|Q]--> max/GraphNorm/ReLU --c1.fc + h'--> x2
//
// Activations never leave the SM between the layers: the A operand (16-bit, K-major SWIZZLE_128B) is
// rebuilt in shared memory by every epilogue, accumulators live in TMEM, the per-graph reductions go
// through one fp32 staging tile (thread = row writes, thread = column reduces, conflict-free with a
// 132-float row pitch).  Weights stream from L2 as 16 KB atoms (128 outputs x 64 k) through a 4-deep
// TMA ring that runs ahead across GEMM boundaries.  Warps 0-7: epilogues (warp % 4 = TMEM lane quarter,
// warp / 4 = column slice of each 128-column chunk); the last warp: TMA + MMA issue.
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace cns {

constexpr int BF_WORKER_WARPS = 16;                      // epilogue warps: warp % 4 = TMEM lane quarter, warp / 4 = column slice
constexpr int BF_WORKERS = BF_WORKER_WARPS * 32;
constexpr int BF_THREADS = BF_WORKERS + 32;             // + one TMA / MMA-issue warp
constexpr int BF_SLICES = BF_WORKER_WARPS / 4;          // column slices of a 128-column chunk (and graph slots of the column passes)
constexpr int BF_SLICE_COLS = 128 / BF_SLICES;
constexpr int BF_GROUPS = BF_SLICE_COLS / 8;            // 8-column trips per thread
constexpr int BF_ATOM = 128 * 128;        // bytes: 128 output rows x 64 k x 2 B
constexpr int BF_RING = 4;
constexpr int BF_LD = 132;                // staging row pitch (floats): 16-byte aligned rows; float4 accesses are conflict-free by row AND by column
constexpr int BF_NVEC = 1664;             // per-column epilogue vectors of the GEMMs, concatenated
constexpr int BF_MAX_ATOMS = 48;
// vector bases of the six GEMMs inside the concatenation
constexpr int VB_C0 = 0, VB_C0FC = 256, VB_G = 384, VB_K = 896, VB_C1 = 1152, VB_C1FC = 1408, VB_OUT = 1536;

constexpr size_t BF_SMEM = 65536 /*A*/ + BF_RING * BF_ATOM /*ring*/ + 128 * BF_LD * 4 /*staging*/ + 3 * BF_NVEC * 4 +
                           BF_MAX_ATOMS * 16 + 1024 /*bars, tables*/ + 1024 /*alignment slack*/;

#ifdef BF_TIMELINE
// A-B / debugging builds only (scripts/build_variant.sh ... -DBF_TIMELINE): clock64 stamps of tile 0's worker thread 0
__device__ long long g_bf_timeline[64];
#define BF_MARK(i) do { if (blockIdx.x == 0 && tid == 0) g_bf_timeline[i] = clock64(); } while (0)
#else
#define BF_MARK(i) do { } while (0)
#endif

struct BfAtom { const unsigned char* src; int meta; int pad; };   // meta: g | nc << 4 | kc << 8 | first << 12 | last << 13

template <bool BF16>
__global__ void __launch_bounds__(BF_THREADS, 1)
backbone_fused_kernel(const __grid_constant__ BackboneFusedArgs p) {
  const int tile = blockIdx.x;
  if (tile >= *p.n_tiles) return;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_s = smem;                                   // [4 k-chunks][128 rows][128 B]
  unsigned char* ring = a_s + 65536;
  float* stg = reinterpret_cast<float*>(ring + BF_RING * BF_ATOM);
  float* vec = stg + 128 * BF_LD;                              // [3][BF_NVEC]: bias, wp row 0, wp row 1
  BfAtom* atoms = reinterpret_cast<BfAtom*>(vec + 3 * BF_NVEC);
  unsigned char* tail = reinterpret_cast<unsigned char*>(atoms + BF_MAX_ATOMS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);          // full[4], empty[4], done[8]
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tail + 128);
  int* n_atoms_s = reinterpret_cast<int*>(tail + 136);
  int* g_beg = reinterpret_cast<int*>(tail + 144);             // [<=129] local first row of the tile's graphs
  unsigned char* alive_s = tail + 144 + 132 * 4;               // [128]
  unsigned char* g_first = alive_s + 128;                      // [128] local first row of the graph a row belongs to

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g0 = p.tile_graph[tile], g1 = p.tile_graph[tile + 1];
  const int G = g1 - g0;
  const int row0 = p.graph_ptr[g0];
  const int R = p.graph_ptr[g1] - row0;
  if (R > 128) return;                                         // a graph above the advertised bound (status word set by graph_tiles)
  const bool has_h = p.hidden_in != nullptr;
  const int kh = has_h ? 4 : 2;                                // k-chunks of the two GRU products

  BF_MARK(0);
#ifdef BF_TIMELINE
  if (blockIdx.x == 0 && tid == 0) { g_bf_timeline[60] = G; g_bf_timeline[61] = R; g_bf_timeline[50] = g_bf_timeline[51] = g_bf_timeline[52] = 0; }
#endif
  // ---- per-tile tables ---------------------------------------------------------------------------------
  for (int k = tid; k <= G; k += BF_THREADS) g_beg[k] = p.graph_ptr[g0 + k] - row0;
  for (int r = tid; r < 128; r += BF_THREADS) alive_s[r] = (r < R && p.mask[row0 + r]) ? 1 : 0;
  for (int k = tid; k < G; k += BF_THREADS) {
    const int rb = p.graph_ptr[g0 + k] - row0, re = p.graph_ptr[g0 + k + 1] - row0;
    for (int r = rb; r < re; ++r) g_first[r] = (unsigned char)rb;
  }
  // per-column epilogue vectors [bias | wp row 0 | wp row 1] of all GEMMs: packed once per weight load (bf_pack_vec)
  for (int i = tid; i < 3 * BF_NVEC / 4; i += BF_THREADS)
    reinterpret_cast<float4*>(vec)[i] = reinterpret_cast<const float4*>(p.vec_tab)[i];
  if (warp == BF_WORKER_WARPS) {
    if (lane == 0) {
      for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bars[i]), 1);
      fence_mbar_init();
      // weight atoms in issue order.  The gate product runs as two N = 256 halves, update gate first: every
      // accumulator then fits TMEM columns [0,256) and [256,512) stays free for fp32 row stashes.
      const unsigned char* img[8] = {p.w_out, p.w_c0, p.w_c0fc, p.w_g, p.w_g, p.w_k, p.w_c1, p.w_c1fc};
      const int ncs[8] = {1, 2, 1, 2, 2, 2, 2, 1};
      const int nc0[8] = {0, 0, 0, 1, 0, 0, 0, 0}, ncstep[8] = {1, 1, 1, 2, 2, 1, 1, 1};    // image n-chunk = nc0 + j * ncstep
      const int kimg[8] = {4, 2, 2, 4, 4, 4, 2, 2};
      const int kuse[8] = {4, 2, 2, kh, kh, kh, 2, 2};
      int n = 0;
      for (int g = 0; g < 8; ++g)
        for (int j = 0; j < ncs[g]; ++j)
          for (int kc = 0; kc < kuse[g]; ++kc) {
            const bool first = j == 0 && kc == 0, last = j == ncs[g] - 1 && kc == kuse[g] - 1;
            atoms[n].src = img[g] + (size_t)((nc0[g] + j * ncstep[g]) * kimg[g] + kc) * BF_ATOM;
            atoms[n].meta = g | (j << 4) | (kc << 8) | ((int)first << 12) | ((int)last << 13);
            ++n;
          }
      *n_atoms_s = n;
      for (int i = 0; i < BF_RING; ++i) {
        mbar_expect_tx(smem_u32(&bars[i]), BF_ATOM);
        bulk_g2s(smem_u32(ring + i * BF_ATOM), atoms[i].src, BF_ATOM, smem_u32(&bars[i]));
      }
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_ptr_s), 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_s, 0);

  if (warp == BF_WORKER_WARPS) {
    // =============================== TMA + MMA issue ====================================================
    const uint32_t fmt = BF16 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(a_s), r_base = smem_u32(ring);
    const int n_atoms = *n_atoms_s;
    int li = BF_RING;
    for (int mi = 0; mi < n_atoms; ++mi) {
      const int meta = atoms[mi].meta;
      if (meta & (1 << 12)) {                   // first atom of a GEMM: wait until its A operand is published
        __syncwarp();
        named_bar(2, BF_THREADS);
        tc_fence_after();
      }
      if (lane == 0) {
        const int slot = mi % BF_RING;
        mbar_wait(smem_u32(&bars[slot]), (uint32_t)((mi / BF_RING) & 1));
        tc_fence_after();
        const int g = meta & 15, j = (meta >> 4) & 15, kc = (meta >> 8) & 15;
        const uint32_t d = tmem_base + (uint32_t)(j * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(d, sw128_desc(a_base + kc * BF_ATOM + ks * 32), sw128_desc(r_base + slot * BF_ATOM + ks * 32), idesc,
                   (kc | ks) ? 1u : 0u);
        umma_commit(smem_u32(&bars[4 + slot]));
        if (meta & (1 << 13)) umma_commit(smem_u32(&bars[8 + g]));
        // refill the slot of the PREVIOUS atom once its MMAs have drained - after this atom's MMAs are queued, so the
        // tensor pipe keeps working while the completion of the previous group is awaited (waiting first serialised
        // issue -> completion -> issue: ~760 cycles per atom against ~270 of MMA work)
        if (mi >= 1 && li < n_atoms) {
          const int ps = (mi - 1) % BF_RING;
          mbar_wait(smem_u32(&bars[4 + ps]), (uint32_t)(((mi - 1) / BF_RING) & 1));
          mbar_expect_tx(smem_u32(&bars[ps]), BF_ATOM);
          bulk_g2s(smem_u32(ring + ps * BF_ATOM), atoms[li].src, BF_ATOM, smem_u32(&bars[ps]));
          ++li;
        }
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogues ==========================================================
    const int wq = warp & 3, half = warp >> 2;
    const int row = wq * 32 + lane;                  // TMEM lane == local cluster row
    const bool live = row < R;
    const bool alive = live && alive_s[row];
    const int64_t grow = (int64_t)row0 + (live ? row : 0);
    const float2 ps = live ? reinterpret_cast<const float2*>(p.pos_clu)[grow] : make_float2(0.f, 0.f);
    const uint32_t t_lane = tmem_base + ((uint32_t)(wq * 32) << 16);
    const float* vb = vec; const float* vw0 = vec + BF_NVEC; const float* vw1 = vec + 2 * BF_NVEC;
    float* srow = stg + row * BF_LD;
    const float* mrow = stg + (live ? (int)g_first[row] : 0) * BF_LD;   // first row of this row's graph: where the column maxima land
    constexpr int S1 = 256, S2 = 384;                // TMEM stashes (fp32 rows): x then u ; h then h'

    auto load8 = [&](const float* src, float (&v)[8]) {
      const float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    };
    auto store8 = [&](float* dst, const float (&v)[8]) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
    };
    auto store_a8 = [&](int acol, const float (&v)[8]) {     // 8 consecutive k of this thread's A row
      *reinterpret_cast<uint4*>(a_s + sw128_off(128, row, acol)) =
          make_uint4(pack2<BF16>(v[0], v[1]), pack2<BF16>(v[2], v[3]), pack2<BF16>(v[4], v[5]), pack2<BF16>(v[6], v[7]));
    };
    auto rows_to_global = [&](float* out) {                // staging rows [0, R) -> out rows [row0, row0 + R)
      for (int r = warp; r < R; r += BF_WORKER_WARPS)
        *reinterpret_cast<float4*>(out + (size_t)(row0 + r) * 128 + lane * 4) = *reinterpret_cast<const float4*>(stg + r * BF_LD + lane * 4);
    };
    // ---- inputs -> A = [cur | tar - cur] (out_enc, graph_vs.py:229-230) ; h -> staging (fp32) -----------------------
    // Coalesced: a warp reads one 512-byte row per instruction (lane = 4 columns), 8 rows per warp; thread = row loads
    // (32 sectors per instruction) kept the load/store unit saturated for ~14 us per tile.
    {
      const float* tar = p.clu_feat + (size_t)p.n_clusters * 128;
      constexpr int RPW = 128 / BF_WORKER_WARPS;       // rows per warp
      float4 xv[RPW], dv[RPW], hv[RPW];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int r = warp + i * BF_WORKER_WARPS;
        xv[i] = dv[i] = hv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < R) {
          const size_t o = (size_t)(row0 + r) * 128 + lane * 4;
          xv[i] = *reinterpret_cast<const float4*>(p.clu_feat + o);
          dv[i] = *reinterpret_cast<const float4*>(tar + o);
          if (has_h) hv[i] = *reinterpret_cast<const float4*>(p.hidden_in + o);
        }
      }
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int r = warp + i * BF_WORKER_WARPS;
        *reinterpret_cast<uint2*>(a_s + sw128_off(128, r, lane * 4)) =
            make_uint2(pack2<BF16>(xv[i].x, xv[i].y), pack2<BF16>(xv[i].z, xv[i].w));
        *reinterpret_cast<uint2*>(a_s + sw128_off(128, r, 128 + lane * 4)) =
            make_uint2(pack2<BF16>(dv[i].x - xv[i].x, dv[i].y - xv[i].y), pack2<BF16>(dv[i].z - xv[i].z, dv[i].w - xv[i].w));
        if (has_h) {
          *reinterpret_cast<float4*>(stg + r * BF_LD + lane * 4) = hv[i];
        }
      }
    }
    BF_MARK(1);      // inputs loaded, A tile of out_enc written
    auto wait_done = [&](int g) { mbar_wait(smem_u32(&bars[8 + g]), 0); tc_fence_after(); BF_MARK(2 + 2 * g); };
    int bf_stage = 0;
    auto publish_a = [&]() { fence_proxy_async(); tc_fence_before(); BF_MARK(3 + 2 * (bf_stage - 1) + (bf_stage == 0 ? 30 : 0)); ++bf_stage; named_bar(2, BF_THREADS); };
    // The epilogues walk this thread's 64 columns of a 128-column TMEM chunk in 8 trips of 8 columns (runtime
    // loops: the whole chain is executed once per tile, so a small instruction footprint matters more than
    // unrolling).  v[i] = accumulator + bias + pos . wp for vector index vi + i.
    // Table reads go through a non-volatile ld.shared (the tables are read-only after the prologue): the compiler may then
    // hoist them above the staging / A-tile stores of the previous column group, which ordinary loads cannot pass
    // (possible aliasing) - every group used to wait for its own six table loads behind the previous group's stores.
    auto lds4 = [](const float* ptr) {
      float4 v;
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(ptr)));
      return v;
    };
    auto lin8 = [&](const float (&acc)[8], int vi, float (&o)[8]) {
      const float4 b0 = lds4(vb + vi), b1 = lds4(vb + vi + 4);
      const float4 x0 = lds4(vw0 + vi), x1 = lds4(vw0 + vi + 4);
      const float4 y0 = lds4(vw1 + vi), y1 = lds4(vw1 + vi + 4);
      o[0] = fmaf(ps.y, y0.x, fmaf(ps.x, x0.x, acc[0] + b0.x)); o[1] = fmaf(ps.y, y0.y, fmaf(ps.x, x0.y, acc[1] + b0.y));
      o[2] = fmaf(ps.y, y0.z, fmaf(ps.x, x0.z, acc[2] + b0.z)); o[3] = fmaf(ps.y, y0.w, fmaf(ps.x, x0.w, acc[3] + b0.w));
      o[4] = fmaf(ps.y, y1.x, fmaf(ps.x, x1.x, acc[4] + b1.x)); o[5] = fmaf(ps.y, y1.y, fmaf(ps.x, x1.y, acc[5] + b1.y));
      o[6] = fmaf(ps.y, y1.z, fmaf(ps.x, x1.z, acc[6] + b1.z)); o[7] = fmaf(ps.y, y1.w, fmaf(ps.x, x1.w, acc[7] + b1.w));
    };
    auto sigmoid_fast = [](float x) { return __fdividef(1.f, 1.f + __expf(-x)); };
    auto tanh_fast = [](float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); };
    // Q part of a [P|Q] product (TMEM columns dcol..dcol+127) -> staging
    auto q_to_staging = [&](int dcol, int vbase) {
#pragma unroll 1
      for (int cp = 0; cp < BF_GROUPS; cp += 2) {          // two TMEM loads per wait
        float acc[2][8];
#pragma unroll
        for (int j = 0; j < 2; ++j) tmem_ld8(t_lane + (uint32_t)(dcol + half * BF_SLICE_COLS + (cp + j) * 8), acc[j]);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int c = half * BF_SLICE_COLS + (cp + j) * 8;
          float o[8];
          lin8(acc[j], vbase + c, o);
          if (live) store8(srow + c, o);
        }
      }
    };
    // Column passes: a warp owns a graph at a time (graphs warp, warp + 16, ...), a lane four adjacent columns, so a warp
    // reads one 512-byte staging row per instruction.
    // per graph: max over the node set (all rows, or rows with mask), left in the graph's FIRST row (row threads read it
    // through `mrow`) or written to every row of the graph (epilogues that go on to overwrite their own row)
    // A unit of column work = (graph, half of the 128 columns): units warp, warp + 16, ... ; a lane owns two adjacent
    // columns, so a typical tile (6 graphs) keeps 12 of the 16 warps busy and a warp's serial walk over the rows is short.
    auto max2 = [](float2& m, const float2 v) { m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); };
    auto colmax = [&](bool use_mask, bool broadcast) {
      for (int u = warp; u < 2 * G; u += BF_WORKER_WARPS) {
        const int k = u >> 1;
        const int rb = g_beg[k], re = g_beg[k + 1];
        float* base = stg + (u & 1) * 64 + lane * 2;
        float2 m0 = make_float2(-INFINITY, -INFINITY), m1 = m0, m2 = m0, m3 = m0;
        int r = rb;
        for (; r + 4 <= re; r += 4) {           // four independent loads in flight
          const float2 a0 = *reinterpret_cast<const float2*>(base + r * BF_LD), a1 = *reinterpret_cast<const float2*>(base + (r + 1) * BF_LD);
          const float2 a2 = *reinterpret_cast<const float2*>(base + (r + 2) * BF_LD), a3 = *reinterpret_cast<const float2*>(base + (r + 3) * BF_LD);
          if (!use_mask || alive_s[r]) max2(m0, a0);
          if (!use_mask || alive_s[r + 1]) max2(m1, a1);
          if (!use_mask || alive_s[r + 2]) max2(m2, a2);
          if (!use_mask || alive_s[r + 3]) max2(m3, a3);
        }
        for (; r < re; ++r)
          if (!use_mask || alive_s[r]) max2(m0, *reinterpret_cast<const float2*>(base + r * BF_LD));
        max2(m0, m1); max2(m2, m3); max2(m0, m2);
        *reinterpret_cast<float2*>(base + rb * BF_LD) = m0;
        if (broadcast) {
#pragma unroll 4
          for (r = rb + 1; r < re; ++r) *reinterpret_cast<float2*>(base + r * BF_LD) = m0;
        }
      }
    };
    auto colmax_first = [&](bool use_mask) { colmax(use_mask, false); };
    auto colmax_broadcast = [&](bool use_mask) { colmax(use_mask, true); };
    // PERConv tail: conv = P + max (0 for rows outside the node set) -> GraphNorm -> ReLU -> A k-chunks 0,1
    auto perconv_epilogue = [&](int vbase, const float* gw, const float* gb, const float* gms) {
      BF_MARK(40);
      q_to_staging(128, vbase + 128);
      BF_MARK(41);
      named_bar(1, BF_WORKERS);
      BF_MARK(42);
      colmax_broadcast(true);
      BF_MARK(43);
      named_bar(1, BF_WORKERS);
      BF_MARK(44);
#pragma unroll 1
      for (int cp = 0; cp < BF_GROUPS; cp += 2) {
      float accp[2][8];
#pragma unroll
      for (int j = 0; j < 2; ++j) tmem_ld8(t_lane + (uint32_t)(half * BF_SLICE_COLS + (cp + j) * 8), accp[j]);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = half * BF_SLICE_COLS + (cp + j) * 8;
        float o[8];
        lin8(accp[j], vbase + c, o);
        if (live) {
          float mx[8];
          load8(srow + c, mx);
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = alive ? o[i] + mx[i] : 0.f;
          store8(srow + c, o);
        }
      }
      }
      BF_MARK(45);
      named_bar(1, BF_WORKERS);
      BF_MARK(46);
      for (int u = warp; u < 2 * G; u += BF_WORKER_WARPS) {
        const int k = u >> 1, c0 = (u & 1) * 64 + lane * 2;
        const int rb = g_beg[k], re = g_beg[k + 1];
        const float inv_n = __fdividef(1.f, (float)(re - rb));
        const float* base = stg + c0;
        const float2 w = *reinterpret_cast<const float2*>(gw + c0), b = *reinterpret_cast<const float2*>(gb + c0);
        const float2 ms = *reinterpret_cast<const float2*>(gms + c0);
        // one pass: sum and sum of squares (two independent chains); mean((x - s)^2) = E[x^2] - 2 s E[x] + s^2 with
        // s = mean * mean_scale
        float2 su0 = make_float2(0.f, 0.f), su1 = su0, sq0 = su0, sq1 = su0;
        int r = rb;
        for (; r + 2 <= re; r += 2) {
          const float2 v0 = *reinterpret_cast<const float2*>(base + r * BF_LD), v1 = *reinterpret_cast<const float2*>(base + (r + 1) * BF_LD);
          su0.x += v0.x; su0.y += v0.y; su1.x += v1.x; su1.y += v1.y;
          sq0.x = fmaf(v0.x, v0.x, sq0.x); sq0.y = fmaf(v0.y, v0.y, sq0.y); sq1.x = fmaf(v1.x, v1.x, sq1.x); sq1.y = fmaf(v1.y, v1.y, sq1.y);
        }
        if (r < re) {
          const float2 v0 = *reinterpret_cast<const float2*>(base + r * BF_LD);
          su0.x += v0.x; su0.y += v0.y; sq0.x = fmaf(v0.x, v0.x, sq0.x); sq0.y = fmaf(v0.y, v0.y, sq0.y);
        }
        BF_MARK(47 + 3 * (u >= BF_WORKER_WARPS));
        auto stat = [&](float su, float s2, float msc, float wc, float& shift, float& scale) {
          const float mean = su * inv_n;
          shift = mean * msc;
          const float var = fmaxf(fmaf(shift, shift - 2.f * mean, s2 * inv_n), 0.f);
          scale = wc * rsqrtf(var + 1e-5f);
        };
        float2 sh, sc;
        stat(su0.x + su1.x, sq0.x + sq1.x, ms.x, w.x, sh.x, sc.x);
        stat(su0.y + su1.y, sq0.y + sq1.y, ms.y, w.y, sh.y, sc.y);
        BF_MARK(48 + 3 * (u >= BF_WORKER_WARPS));
        auto norm_store = [&](int rr, const float2 v) {
          *reinterpret_cast<uint32_t*>(a_s + sw128_off(128, rr, c0)) =
              pack2_relu<BF16>(fmaf(v.x - sh.x, sc.x, b.x), fmaf(v.y - sh.y, sc.y, b.y));
        };
        for (r = rb; r + 4 <= re; r += 4) {     // loads first: a load cannot pass the (possibly aliasing) store in front of it
          const float2 v0 = *reinterpret_cast<const float2*>(base + r * BF_LD), v1 = *reinterpret_cast<const float2*>(base + (r + 1) * BF_LD);
          const float2 v2 = *reinterpret_cast<const float2*>(base + (r + 2) * BF_LD), v3 = *reinterpret_cast<const float2*>(base + (r + 3) * BF_LD);
          norm_store(r, v0); norm_store(r + 1, v1); norm_store(r + 2, v2); norm_store(r + 3, v3);
        }
        for (; r < re; ++r) norm_store(r, *reinterpret_cast<const float2*>(base + r * BF_LD));
        BF_MARK(49 + 3 * (u >= BF_WORKER_WARPS));
      }
    };
    // fc tail: y = relu(acc + b + residual), residual = fp32 row stash in TMEM ; to_a: 16-bit -> A k-chunks 0,1 ;
    // out: fp32 rows to global
    auto fc_epilogue = [&](int vbase, int stash, bool to_a, float* out) {
#pragma unroll 2
      for (int cg = 0; cg < BF_GROUPS; ++cg) {
        const int c = half * BF_SLICE_COLS + cg * 8;
        float acc[8], r8[8], y[8];
        tmem_ld8(t_lane + (uint32_t)c, acc);
        tmem_ld8(t_lane + (uint32_t)(stash + c), r8);
        tmem_ld_wait();
        const float4 b0 = *reinterpret_cast<const float4*>(vb + vbase + c), b1 = *reinterpret_cast<const float4*>(vb + vbase + c + 4);
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = fmaxf(acc[i] + bb[i] + r8[i], 0.f);
        if (live) {
          if (to_a) store_a8(c, y);
          if (out) store8(srow + c, y);
        }
      }
      if (out) { named_bar(1, BF_WORKERS); rows_to_global(out); }
    };
    // one half of the gate product: sigmoid(P + column max of Q over the whole graph) for this thread's columns
    auto gate_half = [&](int vp, int vq, auto&& sink) {
      q_to_staging(128, vq);
      named_bar(1, BF_WORKERS);
      colmax_first(false);
      named_bar(1, BF_WORKERS);
#pragma unroll 2
      for (int cg = 0; cg < BF_GROUPS; ++cg) {
        const int c = half * BF_SLICE_COLS + cg * 8;
        float acc[8], o[8];
        tmem_ld8(t_lane + (uint32_t)c, acc);
        tmem_ld_wait();
        lin8(acc, vp + c, o);
        float mx[8];
        load8(mrow + c, mx);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = sigmoid_fast(o[i] + (live ? mx[i] : 0.f));
        sink(c, o);
      }
      named_bar(1, BF_WORKERS);      // staging is reused by the next epilogue
    };

    // ---- GEMM 0: out_enc: x = relu(.) * cluster_mask -> A k-chunks 0,1 + stash S1 ; h -> A k-chunks 2,3 --------
    publish_a();
    wait_done(0);
#pragma unroll 2
    for (int cg = 0; cg < BF_GROUPS; ++cg) {
      const int c = half * BF_SLICE_COLS + cg * 8;
      float acc[8], y[8];
      tmem_ld8(t_lane + (uint32_t)c, acc);
      tmem_ld_wait();
      const float4 b0 = *reinterpret_cast<const float4*>(vb + VB_OUT + c), b1 = *reinterpret_cast<const float4*>(vb + VB_OUT + c + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = alive ? fmaxf(acc[i] + bb[i], 0.f) : 0.f;
      store_a8(c, y);
      tmem_st8(t_lane + (uint32_t)(S1 + c), y);
      if (has_h) {               // h: staging (prologue) -> fp32 stash S2 + A k-chunks 2,3
        float hv[8];
        load8(srow + c, hv);
        store_a8(128 + c, hv);
        tmem_st8(t_lane + (uint32_t)(S2 + c), hv);
      }
    }
    tmem_st_wait();
    // ---- GEMM 1: c0 [P|Q] --------------------------------------------------------------------------------
    publish_a();
    wait_done(1);
    perconv_epilogue(VB_C0, p.gn0_w, p.gn0_b, p.gn0_ms);
    // ---- GEMM 2: c0.fc + x -> y0 (A k-chunks 0,1) ----------------------------------------------------------
    publish_a();
    wait_done(2);
    fc_epilogue(VB_C0FC, S1, true, nullptr);
    // ---- GEMM 3: update gate u on the target graph (every cluster is a node) -> stash S1 (x is dead) -----------
    publish_a();
    wait_done(3);
    gate_half(VB_G + 128, VB_G + 384, [&](int c, float (&o)[8]) { tmem_st8(t_lane + (uint32_t)(S1 + c), o); });
    tmem_st_wait();
    // ---- GEMM 4: reset gate r -> h * r replaces h in A k-chunks 2,3 (the product has already read them) --------
    publish_a();               // nothing new in A: the barrier only orders this epilogue's TMEM reads before the MMAs
    wait_done(4);
    gate_half(VB_G, VB_G + 256, [&](int c, float (&o)[8]) {
      if (has_h) {
        float hv[8];
        tmem_ld8(t_lane + (uint32_t)(S2 + c), hv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] *= hv[i];
        if (live) store_a8(128 + c, o);
      }
    });
    // ---- GEMM 5: candidate on the current graph: h' = (1 - u) h + u tanh(conv) -> hidden_out, A k-chunks 0,1, S2 --
    publish_a();
    wait_done(5);
    q_to_staging(128, VB_K + 128);
    named_bar(1, BF_WORKERS);
    colmax_broadcast(true);
    named_bar(1, BF_WORKERS);
#pragma unroll 2
    for (int cg = 0; cg < BF_GROUPS; ++cg) {
      const int c = half * BF_SLICE_COLS + cg * 8;
      float acc[8], o[8], hv[8], uu[8];
      tmem_ld8(t_lane + (uint32_t)c, acc);
      tmem_ld8(t_lane + (uint32_t)(S1 + c), uu);
      if (has_h) tmem_ld8(t_lane + (uint32_t)(S2 + c), hv);
      tmem_ld_wait();
      lin8(acc, VB_K + c, o);
      float mx[8];
      load8(srow + c, mx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float cv = alive ? o[i] + mx[i] : 0.f;
        o[i] = (1.f - uu[i]) * (has_h ? hv[i] : 0.f) + uu[i] * tanh_fast(cv);
      }
      tmem_st8(t_lane + (uint32_t)(S2 + c), o);
      if (live) {
        store8(srow + c, o);           // in place (this thread's own row and columns)
        store_a8(c, o);
      }
    }
    tmem_st_wait();
    named_bar(1, BF_WORKERS);
    rows_to_global(p.hidden_out);      // coalesced: a warp writes one 512-byte row per instruction
    named_bar(1, BF_WORKERS);          // staging is reused by the next epilogue
    // ---- GEMM 6: c1 [P|Q] on h' ----------------------------------------------------------------------------
    publish_a();
    wait_done(6);
    perconv_epilogue(VB_C1, p.gn1_w, p.gn1_b, p.gn1_ms);
    // ---- GEMM 7: c1.fc + h' -> x2 ----------------------------------------------------------------------------
    publish_a();
    wait_done(7);
    fc_epilogue(VB_C1FC, S2, false, p.x2);
    BF_MARK(17);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == BF_WORKER_WARPS) tmem_dealloc(tmem_base, 512);
}

#ifdef BF_TIMELINE
extern "C" int cns_bf_timeline(long long* out) {
  return cudaMemcpyFromSymbol(out, g_bf_timeline, sizeof(g_bf_timeline)) == cudaSuccess ? 0 : -1;
}
#endif

__global__ void bf_pack_vec_kernel(BackboneFusedArgs p, float* __restrict__ vec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BF_NVEC) return;
  const float* b; const float* wp; int j, n;
  if (i < VB_C0FC) { b = p.c0_b; wp = p.c0_wp; j = i; n = 256; }
  else if (i < VB_G) { b = p.c0fc_b; wp = nullptr; j = i - VB_C0FC; n = 128; }
  else if (i < VB_K) { b = p.g_b; wp = p.g_wp; j = i - VB_G; n = 512; }
  else if (i < VB_C1) { b = p.k_b; wp = p.k_wp; j = i - VB_K; n = 256; }
  else if (i < VB_C1FC) { b = p.c1_b; wp = p.c1_wp; j = i - VB_C1; n = 256; }
  else if (i < VB_OUT) { b = p.c1fc_b; wp = nullptr; j = i - VB_C1FC; n = 128; }
  else { b = p.out_b; wp = nullptr; j = i - VB_OUT; n = 128; }
  vec[i] = b[j];
  vec[BF_NVEC + i] = wp ? wp[j] : 0.f;
  vec[2 * BF_NVEC + i] = wp ? wp[n + j] : 0.f;
}

// per-column epilogue vectors of the fused kernel, rebuilt with every weight (re)pack
int bf_pack_vec(cns_handle* h, cudaStream_t s) {
  if (!h->pk.bf_vec) CNS_CUDA_CHECK(cudaMalloc(&h->pk.bf_vec, 3 * BF_NVEC * sizeof(float)));
  BackboneFusedArgs a{};
  const Packed& pk = h->pk;
  a.c0_b = pk.c0_b; a.c0_wp = pk.c0_wp; a.g_b = pk.g_b; a.g_wp = pk.g_wp; a.k_b = pk.k_b; a.k_wp = pk.k_wp;
  a.c1_b = pk.c1_b; a.c1_wp = pk.c1_wp; a.c0fc_b = h->P(P_C0_FC_B); a.c1fc_b = h->P(P_C1_FC_B); a.out_b = h->P(P_OUT_B);
  bf_pack_vec_kernel<<<ceil_div(BF_NVEC, 256), 256, 0, s>>>(a, h->pk.bf_vec);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

int backbone_fused(const BackboneFusedArgs& a, bool bf16, cudaStream_t s) {
  if (a.n_graphs == 0 || a.n_clusters == 0) return CNS_OK;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured)) {
    CNS_CUDA_CHECK(cudaFuncSetAttribute(backbone_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BF_SMEM));
    CNS_CUDA_CHECK(cudaFuncSetAttribute(backbone_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BF_SMEM));
  }
  // greedy whole-graph tiles: two consecutive tiles always hold more than 128 rows together
  int64_t grid = 2 * ceil_div(a.n_clusters, 128) + 1;
  if (grid > a.n_graphs) grid = a.n_graphs;
  if (bf16) backbone_fused_kernel<true><<<(unsigned)grid, BF_THREADS, BF_SMEM, s>>>(a);
  else backbone_fused_kernel<false><<<(unsigned)grid, BF_THREADS, BF_SMEM, s>>>(a);
  CNS_LAUNCH_CHECK();
  return CNS_OK;
}

}  // namespace cns
