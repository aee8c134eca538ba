// mp_step_tc.cu -- the edge-row message-passing step on Blackwell tensor cores (tcgen05 + TMEM).
//
// Same math as k_mp_edge<64> in mp_step.cu (reference models/layers.py:84-116, msg_type 'diff',
// + heads of models/track_mpnn.py:73-75), for batches large enough to fill 128-row tiles:
//   gi = x . W_ih^T,  gh = h . W_hh^T  with x = h[src] - h[dst]
// run as tcgen05.mma kind::f16 with a 3-term fp16 split of both operands
//   a = a_hi + a_lo (a_hi = fp16(a), a_lo = fp16(a - a_hi)),  a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo
// (products of fp16 pairs are exact in the fp32 accumulator; the dropped a_lo.b_lo term is 2^-22
// relative), which keeps the result at fp32-reordering level (measured ~2e-6 max-abs vs the
// reference, tolerance 1e-4) at 1.5x the tensor time of a single TF32 pass.
//
// One persistent CTA per SM, 512 threads (128 registers each), warp specialised:
//   warps 0-7   epilogue: TMEM -> registers (tcgen05.ld 32x32b, thread == row; warp w owns TMEM
//               lane quadrant w & 3 and the 32-column half w >> 2 of every gate), previous state
//               from the stage's fp16 hi/lo images, gates, h', head; h' goes through a swizzled
//               per-warp transpose buffer (the stage's h images, dead once read) so
//               that global stores are full 128 B lines
//   warps 8-15  producers: 16 lanes per row, float4 each (full-line loads): gather h[src], h[dst],
//               h[row], subtract, split to fp16 hi/lo, store into the 128B-swizzled K-major UMMA
//               layout.  Software pipelined in registers: the tile's src/dst and its own rows are
//               loaded one tile ahead (HBM latency), the endpoint gathers one round ahead (L2 hits)
//               The 36 tcgen05.mma of tile k are issued by lane 0 of producer warp k mod 8 once
//               all producers have arrived (rotating the issuer keeps the producer warps balanced
//               and the CTA at 16 warps, i.e. 128 registers per thread for the register pipeline)
// Shared memory (227 KB): packed weights as fp16 hi/lo UMMA images (96 KB, resident for the whole
// kernel) + two A stages of [128 x (64 x | 64 h)] fp16 hi/lo (2 x 64 KB).  TMEM: two accumulator
// stages of 256 columns (r | z | i_n | h_n).  mbarriers: full[2] (producers -> MMA),
// done[2] (tcgen05.commit -> epilogue), xfree[2] (tcgen05.commit -> producers: the stage's x images
// are reusable), hfree[2] (epilogue -> producers: the stage's h images, which the epilogue reads the
// previous state from and then re-uses as its transpose buffer, are reusable),
// tfree[2] (epilogue -> MMA: accumulator stage drained).
#include "tc_common.cuh"

// Debug build only (-DTMPNN_TC_TRACE, profiles/trace_tc.py): CTA 0 stamps clock64() at the hand-over points of
// its three roles into g_tc_trace[iteration][16]; compiled out of the shipped library.
#ifdef TMPNN_TC_TRACE
__device__ long long* g_tc_trace = nullptr;
__device__ int g_tc_trace_cap = 0;
#define TC_TRACE(it_, slot_, cond_)                                                                 \
  do {                                                                                              \
    if (blockIdx.x == 0 && (cond_) && g_tc_trace && (it_) < g_tc_trace_cap) g_tc_trace[(it_) * 16 + (slot_)] = clock64(); \
  } while (0)
extern "C" int tmpnn_debug_set_tc_trace(long long* buf, int cap) {
  cudaMemcpyToSymbol(g_tc_trace, &buf, sizeof(buf));
  cudaMemcpyToSymbol(g_tc_trace_cap, &cap, sizeof(cap));
  return 0;
}
#else
#define TC_TRACE(it_, slot_, cond_) do { } while (0)
#endif

// -log2e 2^-k and 2 log2e 2^-k of the launch's weight image (k_pack_gru_tc), refreshed by a device-to-symbol copy in front of
// every launch: constant-bank operands cost the epilogue no registers
__constant__ float c_tc_expo[4];

namespace {

constexpr int EPI_WARPS = 8, PROD_WARPS = 8;
constexpr int TC_THREADS = 32 * (EPI_WARPS + PROD_WARPS);  // 512 -> 128 registers per thread

// ---- weight image ----------------------------------------------------------------------------
__global__ void k_pack_gru_tc(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                              const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                              const float* __restrict__ head_w, const float* __restrict__ head_b,
                              int ldw, int xcol0, unsigned char* __restrict__ img) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // Power-of-two pre-scale of the tensor-core weights (exact in fp32 and fp16): with the stock N(0, 0.01) init the residuals
  // w - fp16(w) are ~5e-6, i.e. fp16 SUBNORMALS (6e-8 steps, ~6 significant bits): the split would carry ~17 bits instead of
  // 22.  Scaling by 2^k with max |w| 2^k in [8192, 16384) keeps every residual of a weight within 2^17 of the largest one
  // a normal number; the epilogues multiply the accumulators by 2^-k (folded into constants they apply anyway).  Every
  // block computes the same maximum (37 k loads, once per weight change).
  __shared__ float s_max[32];
  __shared__ float s_scale;
  {
    float m = 0.f;
    for (int i = threadIdx.x; i < 2 * 192 * 64; i += blockDim.x) {
      const int mm = i / (192 * 64), n = (i / 64) % 192, k = i % 64;
      m = fmaxf(m, fabsf(mm == 0 ? w_ih[n * ldw + xcol0 + k] : w_hh[n * 64 + k]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float mx = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, s_max[w]);
      int k = 0;
      if (mx > 0.f && isfinite(mx)) {
        int e;
        frexpf(mx, &e);            // mx = f 2^e, f in [0.5, 1)
        k = min(max(14 - e, 0), 40);  // mx 2^k in [8192, 16384); never scale down (large weights keep k = 0)
      }
      s_scale = ldexpf(1.0f, k);
    }
    __syncthreads();
  }
  const float wscale = s_scale;
  for (int i = tid; i < 2 * 192 * 64; i += nth) {  // element (matrix m, row n, col k)
    const int m = i / (192 * 64), n = (i / 64) % 192, k = i % 64;
    // msg_type 'concat': the tensor cores take the far-endpoint half W_ih[:, 64:128] (xcol0 = 64)
    const float w = wscale * (m == 0 ? w_ih[n * ldw + xcol0 + k] : w_hh[n * 64 + k]);
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const int nr = m == 0 ? n : (n + 64) % 192;  // W_hh rows in the order n | r | z
    const uint32_t off = sw128(nr, k >> 3) + (k & 7) * 2;
    *reinterpret_cast<__half*>(img + (m == 0 ? OFF_BX_HI : OFF_BH_HI) + off) = hi;
    *reinterpret_cast<__half*>(img + (m == 0 ? OFF_BX_LO : OFF_BH_LO) + off) = lo;
  }
  float* bias = reinterpret_cast<float*>(img + OFF_BIAS);
  for (int i = tid; i < 4 * H; i += nth) {
    const int g = i / H, j = i % H;
    // the n-gate biases live in the accumulators' scale (x 2^k), see the epilogues
    bias[i] = g < 2 ? -LOG2E * (b_ih[g * H + j] + b_hh[g * H + j]) : wscale * (g == 2 ? b_ih[2 * H + j] : b_hh[2 * H + j]);
  }
  float* hw = reinterpret_cast<float*>(img + OFF_HEADW);
  for (int i = tid; i < H; i += nth) hw[i] = head_w[i];
  // k_det_prepare's operands: the source-side half of W_ih transposed ([c][n]: lanes read consecutive n) and
  // b_ih (+ b_hh for r, z)
  float* wt = reinterpret_cast<float*>(img + OFF_WT);
  for (int i = tid; i < 64 * 192; i += nth) {
    const int c = i / 192, n = i % 192;
    wt[i] = w_ih[n * ldw + c];
  }
  float* bs = reinterpret_cast<float*>(img + OFF_BS);
  for (int n = tid; n < 192; n += nth) bs[n] = b_ih[n] + (n < 2 * H ? b_hh[n] : 0.f);
  if (tid == 0) {
    float* hb = reinterpret_cast<float*>(img + OFF_HEADB);
    // hb[1], hb[3]: the epilogues' two exponent constants with 2^-k folded in (copied into constant memory per launch, so
    // that they stay instruction operands); hb[2]: 2^k for k_det_prepare
    hb[0] = head_b[0]; hb[1] = -LOG2E / wscale; hb[2] = wscale; hb[3] = 2.0f * LOG2E / wscale;
  }
}

// ---- the kernel ------------------------------------------------------------------------------
// The producers gather h[src], h[dst], subtract and split per association row (msg_type 'diff' only); the form with
// the endpoints prepared once per detection row lives in mp_step_tc3.cu.
// TRAIN: additionally stores r | z | n | (W_hn h + b_hn) of every association row into gates[row][4][64] (what the backward
// pass needs from torch.nn.GRUCell, models/layers.py:97): the training forward on the tensor cores.
template <bool TRAIN>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mp_edge_tc(const float* __restrict__ h_in, float* __restrict__ h_out, int ldh, int col,
             const int32_t* __restrict__ n_rows, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
             int cap_rows, int num_seqs, const int32_t* __restrict__ tile_ptr, const unsigned char* __restrict__ image,
             float* __restrict__ logit, float* __restrict__ score, int first_group, int last_group,
             int32_t* __restrict__ status, const int32_t* __restrict__ phys, const int32_t* __restrict__ psrc,
             const int32_t* __restrict__ pdst, float* __restrict__ gates) {
  extern __shared__ unsigned char smem_dyn[];
  const int total = tile_ptr[num_seqs];
  if ((int)blockIdx.x >= total) return;  // uniform: whole CTA leaves before touching TMEM / barriers
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = sm_u + OFF_BAR, bar_done = bar_full + 16, bar_xfree = bar_full + 32, bar_hfree = bar_full + 48,
                 bar_tfree = bar_full + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 80);

  // resident weight image (generic-proxy stores, made visible to the async proxy below)
  {
    const uint4* gsrc = reinterpret_cast<const uint4*>(image);
    uint4* sdst = reinterpret_cast<uint4*>(sm);
    for (int i = threadIdx.x; i < IMAGE_BYTES / 16; i += TC_THREADS) sdst[i] = __ldg(gsrc + i);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, PROD_WARPS);  // one arrive per producer warp
      mbar_init(bar_done + 8 * s, 1);           // tcgen05.commit
      mbar_init(bar_xfree + 8 * s, 1);          // tcgen05.commit: the x images are dead once the MMAs retired
      mbar_init(bar_hfree + 8 * s, EPI_WARPS);  // one arrive per epilogue warp: h images / transpose buffer released
      mbar_init(bar_tfree + 8 * s, EPI_WARPS);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int stride = gridDim.x;

  if (warp >= EPI_WARPS) {
    // ================= producers =================
    const int pt = threadIdx.x - 32 * EPI_WARPS;
    const int g = pt >> 4, l = pt & 15, gl0 = lane & 16;  // row group (rows g + 16 p), float4 within the row
    const uint32_t FULL = 0xffffffffu;
    // all addressing in units of float4 from h_in: (global row) * ldh4 + col4 + l fits 32 bits
    // (S * cap_rows * ldh / 4 < 2^32 is checked on the host)
    const float4* __restrict__ h4p = reinterpret_cast<const float4*>(h_in);
    const uint32_t ldh4 = (uint32_t)ldh >> 2, cl4 = ((uint32_t)col >> 2) + (uint32_t)l;
    // lane l < 8 of a group keeps src of row g + 16 l, lane l >= 8 keeps dst of row g + 16 (l - 8)
    // deferred compaction (phys != NULL): the input state sits at GLOBAL physical rows -- endpoints at
    // psrc / pdst, the row itself at phys -- instead of slab base + logical row
    const bool dfr = phys != nullptr;
    const int32_t* __restrict__ idx_arr = dfr ? (l < 8 ? psrc : pdst) : (l < 8 ? src : dst);
    const int idx_row = g + 16 * (l & 7);
    int seq = 0;
    seek_seq(tile_ptr, num_seqs, blockIdx.x, seq);
    uint32_t base = (uint32_t)seq * (uint32_t)cap_rows;                 // first global row of the slab
    int r0 = (blockIdx.x - __ldg(tile_ptr + seq)) * TCM;                 // first slab row of the tile
    int nleft = __ldg(n_rows + seq) - r0;                                // rows of the slab from r0 on
    int idxv = idx_row < nleft ? __ldg(idx_arr + base + r0 + idx_row) : -1;
    // physical row of this thread group's row (l & 7) of the tile (deferred compaction only)
    int pownv = dfr ? __ldg(phys + base + r0 + min(idx_row, nleft - 1)) : 0;
    // Every load below is unconditional with a clamped (always valid) address: a predicated load
    // is compiled as load + predicated move, and the move would wait for the load right away,
    // which defeats the register pipeline.
    float4 own[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int sp = __shfl_sync(FULL, pownv, gl0 + p);
      const uint32_t orow = dfr ? (uint32_t)sp : base + r0 + min(g + 16 * p, nleft - 1);
      own[p] = __ldg(h4p + orow * ldh4 + cl4);
    }
    float4 gs[2][2], gd[2][2];
    auto gather = [&](int buf, int round, uint32_t gbase, int iv) {
      const uint32_t gb = dfr ? 0u : gbase;  // physical endpoints are global rows already
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int p = 2 * round + u;
        const int a = __shfl_sync(FULL, iv, gl0 + p), b = __shfl_sync(FULL, iv, gl0 + 8 + p);
        gs[buf][u] = __ldg(h4p + (gb + (uint32_t)max(a, 0)) * ldh4 + cl4);
        gd[buf][u] = __ldg(h4p + (gb + (uint32_t)max(b, 0)) * ldh4 + cl4);
      }
    };
    gather(0, 0, base, idxv);
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (uint32_t)(it >> 1) & 1u;
      unsigned char* a_stage = sm + OFF_A + stage * A_STAGE;
      // next tile (or this one again when it is the last: the loads are then simply unused)
      uint32_t nbase = base;
      int nr0 = r0, nnleft = nleft, nidx = -1, npown = pownv;
      if (tile + stride < total) {
        seek_seq(tile_ptr, num_seqs, tile + stride, seq);
        nbase = (uint32_t)seq * (uint32_t)cap_rows;
        nr0 = (tile + stride - __ldg(tile_ptr + seq)) * TCM;
        nnleft = __ldg(n_rows + seq) - nr0;
        // in flight for the whole tile; rows past the end of the slab repeat its last row (masked by the epilogue)
        nidx = __ldg(idx_arr + nbase + nr0 + min(idx_row, nnleft - 1));
        if (dfr) npown = __ldg(phys + nbase + nr0 + min(idx_row, nnleft - 1));
      }
      float amax = 0.f;
      TC_TRACE(it, 0, threadIdx.x == 32 * EPI_WARPS);
      // x images first: they are free as soon as the previous tile of this stage left the tensor core
      mbar_wait(bar_xfree + 8 * stage, phase ^ 1u, status);
      TC_TRACE(it, 1, threadIdx.x == 32 * EPI_WARPS);
#pragma unroll
      for (int round = 0; round < 4; ++round) {
        // endpoint rows of the following round (or of round 0 of the next tile): L2 hits, one round ahead
        if (round < 3) gather((round + 1) & 1, round + 1, base, idxv);
        else gather(0, 0, nbase, nidx);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int p = 2 * round + u;
          // rows that are not edge rows (detections inside the tile, rows past the end of the slab) carry
          // real, finite h values from the clamped addresses; their results are masked by the epilogue
          const float4 s4 = gs[round & 1][u], d4 = gd[round & 1][u];
          const float4 x4 = make_float4(s4.x - d4.x, s4.y - d4.y, s4.z - d4.z, s4.w - d4.w);
          uint2 xh, xl;
          split4(x4, xh, xl, amax);
          const uint32_t off = sw128(g + 16 * p, l >> 1) + ((l & 1) << 3);
          *reinterpret_cast<uint2*>(a_stage + off) = xh;
          *reinterpret_cast<uint2*>(a_stage + A_PART + off) = xl;
        }
      }
      TC_TRACE(it, 2, threadIdx.x == 32 * EPI_WARPS);
      // h images: released by the epilogue of the previous tile of this stage
      mbar_wait(bar_hfree + 8 * stage, phase ^ 1u, status);
      TC_TRACE(it, 3, threadIdx.x == 32 * EPI_WARPS);
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const float4 h4 = own[p];
        // this row's slot of the next tile: HBM latency, one tile ahead
        const int sp = __shfl_sync(FULL, npown, gl0 + p);
        const uint32_t orow = dfr ? (uint32_t)sp : nbase + nr0 + min(g + 16 * p, nnleft - 1);
        own[p] = __ldg(h4p + orow * ldh4 + cl4);
        uint2 hh, hl;
        split4(h4, hh, hl, amax);
        const uint32_t off = sw128(g + 16 * p, l >> 1) + ((l & 1) << 3);
        *reinterpret_cast<uint2*>(a_stage + 2 * A_PART + off) = hh;
        *reinterpret_cast<uint2*>(a_stage + 3 * A_PART + off) = hl;
      }
      if (amax > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);  // fp16 split would overflow: use the FMA path
      fence_proxy_async();
      __syncwarp();
      TC_TRACE(it, 4, threadIdx.x == 32 * EPI_WARPS);
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
      __syncwarp();
      if (warp - EPI_WARPS == (it & (PROD_WARPS - 1))) {  // this tile's MMA issuer (warp-uniform test, then one elected lane:
        if (elect_one()) {                                //  descriptors stay in uniform registers, the MMAs issue back to back)
          mbar_wait(bar_tfree + 8 * stage, phase ^ 1u, status);  // accumulator stage drained
          TC_TRACE(it, 5, true);
          mbar_wait(bar_full + 8 * stage, phase, status);        // every producer warp has landed its rows
          TC_TRACE(it, 6, true);
          tc_fence_after();
          issue_tile_mma(sm_u, tmem_base, stage, 0u, bar_xfree + 8 * stage);
          umma_commit(bar_done + 8 * stage);  // accumulators ready (implies tcgen05.fence::before_thread_sync)
          TC_TRACE(it, 7, true);
        }
      }
      __syncwarp();
      base = nbase; r0 = nr0; nleft = nnleft; idxv = nidx; pownv = npown;
    }
  } else {
    // ================= epilogue =================
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;  // row of the tile == TMEM lane
    const int c0 = 32 * half;        // this warp's columns of every gate: [c0, c0 + 32)
    const float* bias = reinterpret_cast<const float*>(sm + OFF_BIAS);
    const float* headw = reinterpret_cast<const float*>(sm + OFF_HEADW);
    const float headb = *reinterpret_cast<const float*>(sm + OFF_HEADB);
    float* dot_part = reinterpret_cast<float*>(sm + OFF_DOT);
    const f32x2 NLOG2E2 = pk2(c_tc_expo[1], c_tc_expo[1]), TWOLOG2E2 = pk2(c_tc_expo[3], c_tc_expo[3]), ONE2 = pk2(1.0f, 1.0f);
    const f32x2 NTWO2 = pk2(-2.0f, -2.0f), NONE2 = pk2(-1.0f, -1.0f);
    const float inv_s = 1.0f / *reinterpret_cast<const float*>(sm + OFF_HEADB + 8);  // 2^-k of the weight pre-scale (exact)
    const f32x2 INVS2 = pk2(inv_s, inv_s);
    // this row's coordinates and its source (< 0: not an edge row) are fetched two tiles ahead, the source's
    int seq = 0, it = 0;
    auto coords = [&](int tile, size_t& rw, int& nr, int& lrr) {
      seek_seq(tile_ptr, num_seqs, tile, seq);
      lrr = (tile - __ldg(tile_ptr + seq)) * TCM + r;
      rw = (size_t)seq * cap_rows + lrr;
      nr = __ldg(n_rows + seq) - lrr;                          // > 0 iff the row exists
    };
    auto ld_src = [&](size_t rw, int nr, int lrr) { return __ldg(src + (nr > 0 ? rw : rw - lrr)); };  // clamped to the slab's first row
    size_t row, row1;
    int nrem, nrem1, lr, lr1, srcv, srcv1;
    coords(blockIdx.x, row, nrem, lr);
    srcv = ld_src(row, nrem, lr);
    row1 = row; nrem1 = nrem; lr1 = lr; srcv1 = srcv;
    if ((int)blockIdx.x + stride < total) {
      coords(blockIdx.x + stride, row1, nrem1, lr1);
      srcv1 = ld_src(row1, nrem1, lr1);
    }
    for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (uint32_t)(it >> 1) & 1u;
      unsigned char* a_stage = sm + OFF_A + stage * A_STAGE;
      const size_t row_cur = row;
      const int nrem_cur = nrem, src_cur = srcv;
      size_t row2 = row1;
      int nrem2 = nrem1, lr2 = lr1, srcv2 = srcv1;
      if (tile + 2 * stride < total) {
        coords(tile + 2 * stride, row2, nrem2, lr2);
        srcv2 = ld_src(row2, nrem2, lr2);
      }
      TC_TRACE(it, 8, threadIdx.x == 0);
      mbar_wait(bar_done + 8 * stage, phase, status);
      tc_fence_after();
      TC_TRACE(it, 9, threadIdx.x == 0);
      // previous state of this row's 32 columns = hi + lo of the stage's h images; once every epilogue
      // warp holds its part in registers the A stage goes back to the producers (before the gate math)
      f32x2 hp[16];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const uint32_t off = sw128(r, 4 * half + ch);
        const uint4 vh = *reinterpret_cast<const uint4*>(a_stage + 2 * A_PART + off);
        const uint4 vl = *reinterpret_cast<const uint4*>(a_stage + 3 * A_PART + off);
        const __half2* ph = reinterpret_cast<const __half2*>(&vh);
        const __half2* pl = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 fh = __half22float2(ph[i]), fl = __half22float2(pl[i]);
          hp[4 * ch + i] = add2(pk2(fh.x, fh.y), pk2(fl.x, fl.y));
        }
      }
      // the pair of warps sharing this row quadrant has read both h images of its rows: from here on
      // they are this warp's transpose buffer ([32 rows x 32 floats], 16 B chunks XOR-swizzled by row)
      named_bar_sync(1 + quad, 64);
      TC_TRACE(it, 10, threadIdx.x == 0);
      const bool valid = nrem_cur > 0 && src_cur >= 0;
      unsigned char* tbuf = a_stage + 2 * A_PART + warp * 4096;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(stage * 256 + c0);
      f32x2 dot2 = 0ull;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t ar[8], az[8], an[8], ahn[8];
        tmem_ld8u(t0 + ch * 8, ar);
        tmem_ld8u(t0 + 64 + ch * 8, az);
        tmem_ld8u(t0 + 128 + ch * 8, an);
        tmem_ld8u(t0 + 192 + ch * 8, ahn);
        const int j0 = c0 + ch * 8;
        tmem_ld_wait();
        f32x2 gr[4], gz[4], gn[4], gh[4];   // TRAIN: the chunk's 8 columns of every gate, stored as one 256-bit word each
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const ulonglong2 br = *reinterpret_cast<const ulonglong2*>(bias + j0 + 4 * v);
          const ulonglong2 bz = *reinterpret_cast<const ulonglong2*>(bias + H + j0 + 4 * v);
          const ulonglong2 bi = *reinterpret_cast<const ulonglong2*>(bias + 2 * H + j0 + 4 * v);
          const ulonglong2 bh = *reinterpret_cast<const ulonglong2*>(bias + 3 * H + j0 + 4 * v);
          const ulonglong2 hw = *reinterpret_cast<const ulonglong2*>(headw + j0 + 4 * v);
          f32x2 o[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 4 * v + 2 * e;  // columns j0 + i, j0 + i + 1
            // r, z = 1 / (1 + 2^(-log2e (acc + b)))   (2^x -> inf gives exactly 0, no clamp needed)
            const f32x2 rg = rcp_2(add2(ex2_2(fma2(pk2u(ar[i], ar[i + 1]), NLOG2E2, e ? br.y : br.x)), ONE2));
            const f32x2 zg = rcp_2(add2(ex2_2(fma2(pk2u(az[i], az[i + 1]), NLOG2E2, e ? bz.y : bz.x)), ONE2));
            // n = tanh(u) = 1 - 2 / (1 + 2^(2 log2e u)),  u = i_n + b_in + r (h_n + b_hn)
            // u is kept in the accumulators' scale (x 2^k: b_in, b_hn and P_n arrive pre-multiplied), TWOLOG2E2 undoes it
            const f32x2 u = fma2(rg, add2(pk2u(ahn[i], ahn[i + 1]), e ? bh.y : bh.x), add2(pk2u(an[i], an[i + 1]), e ? bi.y : bi.x));
            const f32x2 ng = fma2(rcp_2(add2(ex2_2(mul2(u, TWOLOG2E2)), ONE2)), NTWO2, ONE2);
            const f32x2 ov = fma2(zg, fma2(ng, NONE2, hp[4 * ch + 2 * v + e]), ng);  // n + z (h - n) = (1 - z) n + z h
            o[e] = ov;
            dot2 = fma2(ov, e ? hw.y : hw.x, dot2);
            if (TRAIN) {
              gr[2 * v + e] = rg; gz[2 * v + e] = zg; gn[2 * v + e] = ng;
              gh[2 * v + e] = mul2(add2(pk2u(ahn[i], ahn[i + 1]), e ? bh.y : bh.x), INVS2);
            }
          }
          if (TRAIN && valid && v == 1) {
            // columns j0 .. j0 + 7 of the four gates of this row: one 256-bit store (a full 32-byte sector) per gate -- the
            // rows of a warp are 1 KB apart, so 16-byte stores wrote half sectors with twice the requests
            float* gp = gates + row_cur * (4 * H) + j0;
            st_global_256(gp, gr);
            st_global_256(gp + H, gz);
            st_global_256(gp + 2 * H, gn);
            st_global_256(gp + 3 * H, gh);
          }
          *reinterpret_cast<ulonglong2*>(tbuf + lane * 128 + (((2 * ch + v) ^ (lane & 7)) << 4)) = make_ulonglong2(o[0], o[1]);
        }
      }
      float dot;
      {
        float d0, d1;
        up2(dot2, d0, d1);
        dot = d0 + d1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tfree + 8 * stage);  // accumulator stage drained
      TC_TRACE(it, 11, threadIdx.x == 0);
      // transposed read-back: each store instruction writes 4 rows x 128 B (full lines)
      {
        float* out0 = h_out + (row_cur - lane) * ldh + col + c0;  // first row of this warp's quadrant
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rr = 4 * k + (lane >> 3), cc = lane & 7;
          const float4 v = *reinterpret_cast<const float4*>(tbuf + rr * 128 + ((cc ^ (rr & 7)) << 4));
          if ((vmask >> rr) & 1u) *reinterpret_cast<float4*>(out0 + (size_t)rr * ldh + 4 * cc) = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfree + 8 * stage);  // the h images may be refilled
      TC_TRACE(it, 12, threadIdx.x == 0);
      // head: the two column halves of a row live in warps quad and quad + 4
      if (half == 1) dot_part[r] = dot;
      named_bar_sync(1 + quad, 64);
      if (half == 0 && valid) {
        const float lg = dot + dot_part[r] + (first_group ? headb : logit[row_cur]);
        logit[row_cur] = lg;
        if (last_group) score[row_cur] = tmpnn_sigmoid(lg);
      }
      named_bar_sync(1 + quad, 64);
      TC_TRACE(it, 13, threadIdx.x == 0);
      row = row1; nrem = nrem1; lr = lr1; srcv = srcv1;
      row1 = row2; nrem1 = nrem2; lr1 = lr2; srcv1 = srcv2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- per-detection preparation -------------------------------------------------------------------------
// One warp per detection row: (a) its fp16 hi/lo image, written at the row's LOGICAL position of det_img (geometry of
// h: 64 hi halves then 64 lo halves in the row's 256 B), which is what the producers copy for the far endpoint of
// an association row; (b) P'[k] = the row's source-side contribution to the input gates, fp32 FMA:
//   P = h W_ih[:, 0:64]^T,  P'[0:128) = -log2e (P + b_ih + b_hh),  P'[128:192) = P + b_ih.
#ifndef TMPNN_PREP_D
#define TMPNN_PREP_D 8
#endif
constexpr int PREP_D = TMPNN_PREP_D;  // detections per warp pass: every weight read from shared memory feeds PREP_D FMAs
constexpr int PREP_SMEM = (64 * 192 + 8 * PREP_D * 64 + 192) * 4;
__global__ void __launch_bounds__(256)
k_det_prepare(const float* __restrict__ h_in, int ldh, int col, const int32_t* __restrict__ n_dets,
              const int32_t* __restrict__ det_rows, const int32_t* __restrict__ phys, const unsigned char* __restrict__ image,
              float* __restrict__ det_img, float* __restrict__ det_p, int32_t* __restrict__ status) {
  extern __shared__ float prep_sm[];
  float* wt = prep_sm;             // [64][192]: W_ih^T (source half)
  float* hr = prep_sm + 64 * 192;  // [8 warps][PREP_D][64]
  float* bs = hr + 8 * PREP_D * 64;  // [192]
  const int nd = *n_dets;
  if ((int)blockIdx.x * 8 * PREP_D >= nd) return;
  {
    const float4* gw = reinterpret_cast<const float4*>(image + OFF_WT);
    for (int i = threadIdx.x; i < 64 * 192 / 4; i += blockDim.x) reinterpret_cast<float4*>(wt)[i] = __ldg(gw + i);
    const float* gb = reinterpret_cast<const float*>(image + OFF_BS);
    for (int n = threadIdx.x; n < 192; n += blockDim.x) bs[n] = gb[n];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float wscale = *reinterpret_cast<const float*>(image + OFF_HEADB + 8);  // 2^k of the weight pre-scale: P_n joins i_n scaled
  float* hw = hr + w * PREP_D * 64;
  for (int k0 = (blockIdx.x * 8 + w) * PREP_D; k0 < nd; k0 += gridDim.x * 8 * PREP_D) {
    const int cnt = min(PREP_D, nd - k0);
#pragma unroll
    for (int d = 0; d < PREP_D; ++d) {
      float2 v = make_float2(0.f, 0.f);
      if (d < cnt) {
        const int row = det_rows[k0 + d];
        const size_t pr = phys ? (size_t)phys[row] : (size_t)row;  // deferred compaction: the state sits at the physical row
        v = *reinterpret_cast<const float2*>(h_in + pr * ldh + col + 2 * lane);
        const __half2 hi = __floats2half2_rn(v.x, v.y);
        const float2 f = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(v.x - f.x, v.y - f.y);
        uint32_t* ib = reinterpret_cast<uint32_t*>(det_img + (size_t)row * ldh + col);
        ib[lane] = *reinterpret_cast<const uint32_t*>(&hi);
        ib[32 + lane] = *reinterpret_cast<const uint32_t*>(&lo);
        if (fmaxf(fabsf(v.x), fabsf(v.y)) > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);
      }
      hw[d * 64 + 2 * lane] = v.x;
      hw[d * 64 + 2 * lane + 1] = v.y;
    }
    __syncwarp();
    float acc[PREP_D][6];
#pragma unroll
    for (int d = 0; d < PREP_D; ++d)
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[d][q] = 0.f;
#pragma unroll 4
    for (int c = 0; c < 64; ++c) {
      float wv[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) wv[q] = wt[c * 192 + lane + 32 * q];
#pragma unroll
      for (int d = 0; d < PREP_D; ++d) {
        const float hv = hw[d * 64 + c];
#pragma unroll
        for (int q = 0; q < 6; ++q) acc[d][q] = fmaf(hv, wv[q], acc[d][q]);
      }
    }
#pragma unroll
    for (int d = 0; d < PREP_D; ++d) {
      if (d < cnt) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int n = lane + 32 * q;
          const float val = acc[d][q] + bs[n];
          det_p[(size_t)(k0 + d) * 192 + n] = q < 4 ? -LOG2E * val : wscale * val;
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace

int tmpnn_init_tc() {
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_edge_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_edge_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_det_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, PREP_SMEM));
  return TMPNN_OK;
}

extern "C" size_t tmpnn_gru_tc_pack_bytes(void) { return (size_t)IMAGE_TOTAL_BYTES; }

extern "C" int tmpnn_pack_gru_tc(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                 const float* head_w, const float* head_b, int concat, void* packed, void* stream) {
  TMPNN_REQUIRE(w_ih && w_hh && b_ih && b_hh && head_w && head_b && packed, "null argument");
  TMPNN_REQUIRE(((uintptr_t)packed & 15) == 0, "packed image must be 16-byte aligned");
  k_pack_gru_tc<<<48, 256, 0, (cudaStream_t)stream>>>(w_ih, w_hh, b_ih, b_hh, head_w, head_b, concat ? 128 : 64,
                                                      concat ? 64 : 0, (unsigned char*)packed);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

static int edge_tc_launch(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh, int group,
                          int num_groups, const void* edge_image, float* gates, void* stream) {
  TMPNN_REQUIRE(g && ix && h_in && h_out && edge_image && ix->tile128_ptr, "null argument");
  TMPNN_REQUIRE(h_in != h_out, "h_in and h_out must be distinct buffers (Jacobi update)");
  TMPNN_REQUIRE(ldh % 4 == 0 && group >= 0 && group < num_groups && ldh >= num_groups * H, "bad ldh / group");
  int rc = tmpnn_init();
  if (rc) return rc;
  TMPNN_CUDA_TRY(cudaMemcpyToSymbolAsync(c_tc_expo, (const unsigned char*)edge_image + OFF_HEADB, 16, 0, cudaMemcpyDeviceToDevice,
                                         (cudaStream_t)stream));
  if (gates)
    k_mp_edge_tc<true><<<TMPNN_SM_COUNT, TC_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(
        h_in, h_out, ldh, group * H, g->n_rows, g->src, g->dst, g->cap_rows, g->num_seqs, ix->tile128_ptr,
        (const unsigned char*)edge_image, g->logit, g->score, group == 0, group == num_groups - 1, g->status, g->phys,
        g->psrc, g->pdst, gates);
  else
    k_mp_edge_tc<false><<<TMPNN_SM_COUNT, TC_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(
        h_in, h_out, ldh, group * H, g->n_rows, g->src, g->dst, g->cap_rows, g->num_seqs, ix->tile128_ptr,
        (const unsigned char*)edge_image, g->logit, g->score, group == 0, group == num_groups - 1, g->status, g->phys,
        g->psrc, g->pdst, nullptr);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_mp_edge_fwd_tc(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                    int group, int num_groups, const void* edge_image, void* stream) {
  return edge_tc_launch(g, ix, h_in, h_out, ldh, group, num_groups, edge_image, nullptr, stream);
}

extern "C" int tmpnn_mp_edge_fwd_tc_train(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                          int group, int num_groups, const void* edge_image, float* gates, void* stream) {
  TMPNN_REQUIRE(gates, "null argument");
  return edge_tc_launch(g, ix, h_in, h_out, ldh, group, num_groups, edge_image, gates, stream);
}

extern "C" int tmpnn_mp_edge_fwd_tc_pre(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                        int group, int num_groups, int concat, const void* edge_image, float* det_img,
                                        float* det_p, void* tile_table, void* stream) {
  TMPNN_REQUIRE(g && ix && h_in && h_out && edge_image && ix->tile128_ptr && ix->det_of_row && ix->det_rows, "null argument");
  TMPNN_REQUIRE(det_img && det_p && tile_table, "null argument");
  TMPNN_REQUIRE(h_in != h_out && det_img != h_in && det_img != h_out, "h_in, h_out and det_img must be distinct buffers");
  TMPNN_REQUIRE(ldh % 4 == 0 && group >= 0 && group < num_groups && ldh >= num_groups * H, "bad ldh / group");
  int rc = tmpnn_init();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  k_det_prepare<<<TMPNN_SM_COUNT * 4, 256, PREP_SMEM, st>>>(h_in, ldh, group * H, ix->n_dets, ix->det_rows, g->phys,
                                                       (const unsigned char*)edge_image, det_img, det_p, g->status);
  TMPNN_LAUNCH_CHECK();
  TMPNN_REQUIRE(((uintptr_t)tile_table & 15) == 0, "tile_table must be 16-byte aligned");
  return tmpnn_edge_tc3_launch(g, ix, h_in, h_out, ldh, group, num_groups, concat, edge_image, det_img, det_p, tile_table, st);
}
