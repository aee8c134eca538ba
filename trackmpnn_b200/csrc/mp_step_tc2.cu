// mp_step_tc2.cu -- the edge-row message-passing step on tcgen05, second generation.
//
// Same math and the same shared-memory / TMEM layout as k_mp_edge_tc<true> in mp_step_tc.cu (reference
// models/layers.py:84-116 + heads of models/track_mpnn.py:73-75; endpoints prepared once per detection row by
// k_det_prepare below), re-balanced after the profile of that kernel (profiles/r01_tc_pre_v1_*): with the
// producers reduced to copies, the eight epilogue warps were the critical path -- two warps per scheduler running
// long dependent chains (tcgen05.ld -> FFMA2 -> MUFU -> ...) at ~8 cycles per instruction, plus three dependent
// global loads per tile to find the tile's slab.  Here:
//   * 16 epilogue warps (4 per scheduler; warp w owns TMEM lane quadrant w & 3 and 16 of the 64 columns) and
//     4 producer warps (one per scheduler): 640 threads, 96 registers each;
//   * a tile table {first global row of the slab, first slab row of the tile, rows left} written by k_tile_table
//     replaces the tile_ptr search, so finding a tile is one 16-byte load issued three tiles ahead;
//   * the head's partial sums are combined in a fixed order ((c0 + c1) + (c2 + c3)) through two 512-byte arrays.
#include "tc_common.cuh"

namespace {

constexpr int EPI2 = 16, PROD2 = 8;
constexpr int TC2_THREADS = 32 * (EPI2 + PROD2);  // 768 -> 80 registers per thread at launch; setmaxnreg moves the
                                                  // budget to 88 for the epilogue warps, 56 for the producers (8 x 32 x 24 freed >= 16 x 32 x 8 taken)
constexpr int NG2 = 2 * PROD2;                    // producer row groups (16 lanes each)
constexpr int RPT2 = TCM / NG2;                   // tile rows per producer thread: g + NG2 p
constexpr uint32_t PSTEP = (uint32_t)(NG2 / 8) * 1024u;  // swizzled-image distance between rows g + NG2 p and g + NG2 (p + 1)

// ---- tile table ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tile_table(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ tile128_ptr, int cap_rows,
             int4* __restrict__ tab) {
  const int s = blockIdx.y;
  const int t0 = tile128_ptr[s], nt = tile128_ptr[s + 1] - t0, n = n_rows[s];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nt; j += gridDim.x * blockDim.x)
    tab[t0 + j] = make_int4(s * cap_rows, j * TCM, n - j * TCM, 0);
}

// ---- per-detection preparation -------------------------------------------------------------------------
// One warp per detection row: (a) its fp16 hi/lo image, written at the row's LOGICAL position of det_img (geometry of
// h: 64 hi halves then 64 lo halves in the row's 256 B), which is what the producers copy for the far endpoint of
// an association row; (b) P'[k] = the row's source-side contribution to the input gates, fp32 FMA:
//   P = h W_ih[:, 0:64]^T,  P'[0:128) = -log2e (P + b_ih + b_hh),  P'[128:192) = P + b_ih.
constexpr int PREP_SMEM = (64 * 192 + 8 * 64 + 192) * 4;
__global__ void __launch_bounds__(256)
k_det_prepare(const float* __restrict__ h_in, int ldh, int col, const int32_t* __restrict__ n_dets,
              const int32_t* __restrict__ det_rows, const int32_t* __restrict__ phys, const float* __restrict__ w_ih, int ldw,
              const float* __restrict__ b_ih, const float* __restrict__ b_hh, float* __restrict__ det_img,
              float* __restrict__ det_p, int32_t* __restrict__ status) {
  extern __shared__ float prep_sm[];
  float* wt = prep_sm;             // [64][192]: W_ih^T (source half)
  float* hr = prep_sm + 64 * 192;  // [8][64]
  float* bs = hr + 8 * 64;         // [192]
  const int nd = *n_dets;
  if ((int)blockIdx.x * 8 >= nd) return;
  for (int i = threadIdx.x; i < 192 * 64; i += blockDim.x) {
    const int n = i % 192, c = i / 192;
    wt[c * 192 + n] = w_ih[n * ldw + c];
  }
  for (int n = threadIdx.x; n < 192; n += blockDim.x) bs[n] = b_ih[n] + (n < 2 * H ? b_hh[n] : 0.f);
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* hw = hr + w * 64;
  for (int k = blockIdx.x * 8 + w; k < nd; k += gridDim.x * 8) {
    const int row = det_rows[k];
    const size_t pr = phys ? (size_t)phys[row] : (size_t)row;  // deferred compaction: the state sits at the physical row
    const float2 v = *reinterpret_cast<const float2*>(h_in + pr * ldh + col + 2 * lane);
    hw[2 * lane] = v.x;
    hw[2 * lane + 1] = v.y;
    const __half2 hi = __floats2half2_rn(v.x, v.y);
    const float2 f = __half22float2(hi);
    const __half2 lo = __floats2half2_rn(v.x - f.x, v.y - f.y);
    uint32_t* ib = reinterpret_cast<uint32_t*>(det_img + (size_t)row * ldh + col);
    ib[lane] = *reinterpret_cast<const uint32_t*>(&hi);
    ib[32 + lane] = *reinterpret_cast<const uint32_t*>(&lo);
    if (fmaxf(fabsf(v.x), fabsf(v.y)) > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);
    __syncwarp();
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int c = 0; c < 64; ++c) {
      const float hv = hw[c];
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[q] = fmaf(hv, wt[c * 192 + lane + 32 * q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const int n = lane + 32 * q;
      const float val = acc[q] + bs[n];
      det_p[(size_t)k * 192 + n] = q < 4 ? -LOG2E * val : val;
    }
    __syncwarp();
  }
}

// ---- the kernel ------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC2_THREADS, 1)
k_mp_edge_tc2(const float* __restrict__ h_in, float* __restrict__ h_out, int ldh, int col,
              const int32_t* __restrict__ src, const int32_t* __restrict__ dst, const int32_t* __restrict__ n_tiles,
              const int4* __restrict__ tab, const unsigned char* __restrict__ image, float* __restrict__ logit,
              float* __restrict__ score, int first_group, int last_group, int32_t* __restrict__ status,
              const int32_t* __restrict__ phys, const float* __restrict__ det_img, const float* __restrict__ det_p,
              const int32_t* __restrict__ det_of_row, uint32_t xflags) {
  extern __shared__ unsigned char smem_dyn[];
  const int total = *n_tiles;
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
    for (int i = threadIdx.x; i < IMAGE_BYTES / 16; i += TC2_THREADS) sdst[i] = __ldg(gsrc + i);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, PROD2);   // one arrive per producer warp
      mbar_init(bar_done + 8 * s, 1);       // tcgen05.commit
      mbar_init(bar_xfree + 8 * s, 1);      // tcgen05.commit: the x images are dead once the x MMAs retired
      mbar_init(bar_hfree + 8 * s, EPI2);   // one arrive per epilogue warp: h images / transpose buffer released
      mbar_init(bar_tfree + 8 * s, EPI2);   // one arrive per epilogue warp: accumulator stage drained
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
  // tile table entry {slab's first global row, tile's first slab row, rows of the slab from there on}; tiles past
  // the end repeat the last one (their loads are simply unused)
  auto ldtab = [&](int tile) { return __ldg(tab + min(tile, total - 1)); };

  if (warp >= EPI2) {
    // ================= producers: 16 lanes per row, rows g + NG2 p =================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int pt = threadIdx.x - 32 * EPI2;
    const int g = pt >> 4, l = pt & 15, gl0 = lane & 16;
    const uint32_t FULL = 0xffffffffu;
    // all addressing in units of float4 from h_in: (global row) * ldh4 + col4 + l fits 32 bits (checked on the host)
    const float4* __restrict__ h4p = reinterpret_cast<const float4*>(h_in);
    const uint32_t ldh4 = (uint32_t)ldh >> 2, cl4 = ((uint32_t)col >> 2) + (uint32_t)l;
    const bool dfr = phys != nullptr;  // deferred compaction: own rows at GLOBAL physical rows
    const int idx_row = g + NG2 * (l % RPT2);  // lane l of a group keeps the far endpoint / physical row of tile row g + NG2 (l % RPT2)
    // detection images have the geometry of h: chunk l of the 256 B image of row R sits at
    // (R * ldh + col) * 4 + 16 l; chunks 0-7 are the hi halves (K order), 8-15 the lo halves
    const unsigned char* __restrict__ imgb = reinterpret_cast<const unsigned char*>(det_img) + (size_t)col * 4 + 16 * l;
    const size_t row_bytes = (size_t)ldh * 4;
    // swizzled offsets of tile row g + NG2 p (same row & 7 for every p): + PSTEP p
    const uint32_t x_dst0 = sm_u + OFF_A + (uint32_t)(l >> 3) * A_PART + sw128(g, l & 7);
    const uint32_t h_off0 = sw128(g, l >> 1) + ((l & 1) << 3);
    // rows past the end of the slab repeat its last row; everything they produce is masked by the epilogue
    auto ld_idx = [&](const int4 T) { return __ldg(dst + T.x + T.y + min(idx_row, T.z - 1)); };
    auto ld_phys = [&](const int4 T) { return dfr ? __ldg(phys + T.x + T.y + min(idx_row, T.z - 1)) : 0; };
    // far-endpoint images of one tile -> x images of a stage: 16 x 16 B per thread, no registers, no ALU
    auto issue_x = [&](int st, uint32_t b, int iv) {
      const uint32_t s0 = x_dst0 + (uint32_t)st * A_STAGE;
#pragma unroll
      for (int p = 0; p < RPT2; ++p) {
        const int d = __shfl_sync(FULL, iv, gl0 + p);  // -1 for detection rows inside the tile: any valid row will do
        const unsigned char* sp = imgb + (size_t)(b + (uint32_t)max(d, 0)) * row_bytes;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + PSTEP * p), "l"(sp) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int4 T0 = ldtab(blockIdx.x), T1 = ldtab(blockIdx.x + stride), T2 = ldtab(blockIdx.x + 2 * stride);
    const int i0 = ld_idx(T0), pw0 = ld_phys(T0);
    int i1 = ld_idx(T1), pw1 = ld_phys(T1);
    float4 own[RPT2];
#pragma unroll
    for (int p = 0; p < RPT2; ++p) {
      const int sp = __shfl_sync(FULL, pw0, gl0 + p);
      const uint32_t orow = dfr ? (uint32_t)sp : (uint32_t)(T0.x + T0.y + min(g + NG2 * p, T0.z - 1));
      own[p] = __ldg(h4p + orow * ldh4 + cl4);
    }
    issue_x(0, (uint32_t)T0.x, i0);
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (uint32_t)(it >> 1) & 1u;
      unsigned char* a_stage = sm + OFF_A + stage * A_STAGE;
      const int4 T3 = ldtab(tile + 3 * stride);            // in flight for a whole tile
      const int i2 = ld_idx(T2), pw2 = ld_phys(T2);        // T2 landed a tile ago; these are consumed a tile from now
      float amax = 0.f;
      // h images: released by the epilogue of the previous tile of this stage
      mbar_wait(bar_hfree + 8 * stage, phase ^ 1u, status);
#pragma unroll
      for (int p = 0; p < RPT2; ++p) {
        const float4 h4 = own[p];
        // this row's slot of the next tile: HBM latency, one tile ahead (unconditional, clamped address)
        const int sp = __shfl_sync(FULL, pw1, gl0 + p);
        const uint32_t orow = dfr ? (uint32_t)sp : (uint32_t)(T1.x + T1.y + min(g + NG2 * p, T1.z - 1));
        own[p] = __ldg(h4p + orow * ldh4 + cl4);
        uint2 hh, hl;
        split4(h4, hh, hl, amax);
        const uint32_t off = h_off0 + PSTEP * p;
        *reinterpret_cast<uint2*>(a_stage + 2 * A_PART + off) = hh;
        *reinterpret_cast<uint2*>(a_stage + 3 * A_PART + off) = hl;
      }
      if (amax > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);  // fp16 split would overflow: use the FMA path
      asm volatile("cp.async.wait_group 0;" ::: "memory");  // this tile's x images (issued one tile ago) have landed
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar_full + 8 * stage);
        if (warp - EPI2 == (it & (PROD2 - 1))) {  // this tile's MMA issuer
          mbar_wait(bar_tfree + 8 * stage, phase ^ 1u, status);  // accumulator stage drained
          mbar_wait(bar_full + 8 * stage, phase, status);        // every producer warp has landed its rows
          tc_fence_after();
          issue_tile_mma(sm_u, tmem_base, stage, xflags, bar_xfree + 8 * stage);
          umma_commit(bar_done + 8 * stage);  // accumulators ready (implies tcgen05.fence::before_thread_sync)
        }
      }
      __syncwarp();
      if (tile + stride < total) {
        // the other stage's x images are free once the previous tile's x MMAs retired (issued a tile ago)
        mbar_wait(bar_xfree + 8 * (stage ^ 1), ((uint32_t)((it + 1) >> 1) & 1u) ^ 1u, status);
        issue_x(stage ^ 1, (uint32_t)T1.x, i1);
      }
      T0 = T1; T1 = T2; T2 = T3; i1 = i2; pw1 = pw2;
    }
  } else {
    // ================= epilogue: 16 warps =================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int quad = warp & 3, cq = warp >> 2;
    const int r = quad * 32 + lane;  // row of the tile == TMEM lane
    const int c0 = 16 * cq;          // this warp's columns of every gate: [c0, c0 + 16)
    const float* bias = reinterpret_cast<const float*>(sm + OFF_BIAS);
    const float* headw = reinterpret_cast<const float*>(sm + OFF_HEADW);
    const float headb = *reinterpret_cast<const float*>(sm + OFF_HEADB);
    // head partial sums: the r | z | i_n bias slots of the image are unused here (folded into P')
    float* part_a = reinterpret_cast<float*>(sm + OFF_BIAS);  // column quarter 1
    float* part_b = reinterpret_cast<float*>(sm + OFF_DOT);   // column quarters 3, then 2 + 3
    const f32x2 NLOG2E2 = pk2(-LOG2E, -LOG2E), TWOLOG2E2 = pk2(2.0f * LOG2E, 2.0f * LOG2E), ONE2 = pk2(1.0f, 1.0f);
    const f32x2 NTWO2 = pk2(-2.0f, -2.0f), NONE2 = pk2(-1.0f, -1.0f);
    // transpose buffer: the slices that alias the h images of this quadrant's rows belong to this quadrant's warps
    const uint32_t tb_off = (uint32_t)(2 * A_PART + (cq >> 1) * A_PART + quad * 4096 + (cq & 1) * 2048);
    // this row's coordinates and source (< 0: not an edge row) are fetched two tiles ahead, the source's position
    // in the detection list (row of P') one tile ahead
    auto ld_src = [&](const int4 T) { return __ldg(src + T.x + (T.z - r > 0 ? T.y + r : 0)); };  // clamped to the slab's first row
    int4 T0 = ldtab(blockIdx.x), T1 = ldtab(blockIdx.x + stride), T2 = ldtab(blockIdx.x + 2 * stride);
    int srcv = ld_src(T0), srcv1 = ld_src(T1);
    int ks = __ldg(det_of_row + T0.x + max(srcv, 0));
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
      const int stage = it & 1;
      const uint32_t phase = (uint32_t)(it >> 1) & 1u;
      unsigned char* a_stage = sm + OFF_A + stage * A_STAGE;
      const size_t row_cur = (size_t)T0.x + T0.y + r;
      const bool valid = T0.z - r > 0 && srcv >= 0;
      const float* __restrict__ pp = det_p + (size_t)max(ks, 0) * 192 + c0;  // this row's source contribution
      const int4 T3 = ldtab(tile + 3 * stride);
      const int ks1 = __ldg(det_of_row + T1.x + max(srcv1, 0));  // srcv1 landed during the previous tile
      const int srcv2 = ld_src(T2);
      mbar_wait(bar_done + 8 * stage, phase, status);
      tc_fence_after();
      // previous state of this row's 16 columns = hi + lo of the stage's h images
      f32x2 hp[8];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const uint32_t off = sw128(r, 2 * cq + ch);
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
      // the four warps sharing this row quadrant have read the h images of its rows: from here on they are
      // these warps' transpose buffers ([32 rows x 16 floats] each, 16 B chunks XOR-swizzled by row)
      named_bar_sync(1 + quad, 128);
      unsigned char* tbuf = a_stage + tb_off;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(stage * 256 + c0);
      f32x2 dot2 = 0ull;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t ar[8], az[8], an[8], ahn[8];
        tmem_ld8u(t0 + ch * 8, ar);
        tmem_ld8u(t0 + 64 + ch * 8, az);
        tmem_ld8u(t0 + 128 + ch * 8, an);
        tmem_ld8u(t0 + 192 + ch * 8, ahn);
        const int j0 = c0 + ch * 8;
        // additive terms of the three input gates: the source's P' row, which already holds
        // -log2e (P_r + b_ir + b_hr) | -log2e (P_z + b_iz + b_hz) | P_n + b_in
        ulonglong2 brv[2], bzv[2], biv[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          brv[v] = __ldg(reinterpret_cast<const ulonglong2*>(pp + ch * 8 + 4 * v));
          bzv[v] = __ldg(reinterpret_cast<const ulonglong2*>(pp + H + ch * 8 + 4 * v));
          biv[v] = __ldg(reinterpret_cast<const ulonglong2*>(pp + 2 * H + ch * 8 + 4 * v));
        }
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const ulonglong2 br = brv[v], bz = bzv[v], bi = biv[v];
          const ulonglong2 bh = *reinterpret_cast<const ulonglong2*>(bias + 3 * H + j0 + 4 * v);
          const ulonglong2 hw = *reinterpret_cast<const ulonglong2*>(headw + j0 + 4 * v);
          f32x2 o[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 4 * v + 2 * e;  // columns j0 + i, j0 + i + 1
            // r, z = 1 / (1 + 2^(-log2e (acc + P + b)))   (2^x -> inf gives exactly 0, no clamp needed)
            const f32x2 rg = rcp_2(add2(ex2_2(fma2(pk2u(ar[i], ar[i + 1]), NLOG2E2, e ? br.y : br.x)), ONE2));
            const f32x2 zg = rcp_2(add2(ex2_2(fma2(pk2u(az[i], az[i + 1]), NLOG2E2, e ? bz.y : bz.x)), ONE2));
            // n = tanh(u) = 1 - 2 / (1 + 2^(2 log2e u)),  u = i_n + P_n + b_in + r (h_n + b_hn)
            const f32x2 u = fma2(rg, add2(pk2u(ahn[i], ahn[i + 1]), e ? bh.y : bh.x), add2(pk2u(an[i], an[i + 1]), e ? bi.y : bi.x));
            const f32x2 ng = fma2(rcp_2(add2(ex2_2(mul2(u, TWOLOG2E2)), ONE2)), NTWO2, ONE2);
            const f32x2 ov = fma2(zg, fma2(ng, NONE2, hp[4 * ch + 2 * v + e]), ng);  // n + z (h - n) = (1 - z) n + z h
            o[e] = ov;
            dot2 = fma2(ov, e ? hw.y : hw.x, dot2);
          }
          *reinterpret_cast<ulonglong2*>(tbuf + lane * 64 + (((2 * ch + v) ^ ((lane >> 1) & 3)) << 4)) = make_ulonglong2(o[0], o[1]);
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
      // transposed read-back: each store instruction writes 8 rows x 64 B
      {
        float* out0 = h_out + (row_cur - lane) * ldh + col + c0;  // first row of this warp's quadrant
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rr = 8 * k + (lane >> 2), cc = lane & 3;
          const float4 v = *reinterpret_cast<const float4*>(tbuf + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
          if ((vmask >> rr) & 1u) *reinterpret_cast<float4*>(out0 + (size_t)rr * ldh + 4 * cc) = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfree + 8 * stage);  // the h images may be refilled
      // head: the four column quarters of a row live in warps quad + 4 cq; fixed order (c0 + c1) + (c2 + c3)
      if (cq == 1) part_a[r] = dot;
      if (cq == 3) part_b[r] = dot;
      named_bar_sync(1 + quad, 128);
      if (cq == 2) part_b[r] = dot + part_b[r];
      named_bar_sync(1 + quad, 128);
      if (cq == 0 && valid) {
        const float lg = ((dot + part_a[r]) + part_b[r]) + (first_group ? headb : logit[row_cur]);
        logit[row_cur] = lg;
        if (last_group) score[row_cur] = tmpnn_sigmoid(lg);
      }
      named_bar_sync(1 + quad, 128);
      T0 = T1; T1 = T2; T2 = T3; srcv = srcv1; srcv1 = srcv2; ks = ks1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

int tmpnn_init_tc2() {
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_edge_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_det_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, PREP_SMEM));
  return TMPNN_OK;
}

int tmpnn_det_prepare_launch(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, int ldh, int group, int concat,
                             const float* w_ih, const float* b_ih, const float* b_hh, float* det_img, float* det_p,
                             cudaStream_t st) {
  k_det_prepare<<<TMPNN_SM_COUNT * 2, 256, PREP_SMEM, st>>>(h_in, ldh, group * H, ix->n_dets, ix->det_rows, g->phys, w_ih,
                                                           concat ? 128 : 64, b_ih, b_hh, det_img, det_p, g->status);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" size_t tmpnn_tc_tile_table_bytes(int num_seqs, int cap_rows) {
  return (size_t)num_seqs * (size_t)tmpnn_div_up(cap_rows, TCM) * sizeof(int4) + sizeof(int4);
}

extern "C" int tmpnn_mp_edge_fwd_tc2(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                     int group, int num_groups, int concat, const void* edge_image, const float* w_ih,
                                     const float* b_ih, const float* b_hh, float* det_img, float* det_p, void* tile_table,
                                     void* stream) {
  TMPNN_REQUIRE(g && ix && h_in && h_out && edge_image && ix->tile128_ptr && ix->det_of_row && ix->det_rows, "null argument");
  TMPNN_REQUIRE(w_ih && b_ih && b_hh && det_img && det_p && tile_table, "null argument");
  TMPNN_REQUIRE(h_in != h_out && det_img != h_in && det_img != h_out, "h_in, h_out and det_img must be distinct buffers");
  TMPNN_REQUIRE(ldh % 4 == 0 && group >= 0 && group < num_groups && ldh >= num_groups * H, "bad ldh / group");
  TMPNN_REQUIRE(((uintptr_t)tile_table & 15) == 0, "tile_table must be 16-byte aligned");
  int rc = tmpnn_init();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = tmpnn_det_prepare_launch(g, ix, h_in, ldh, group, concat, w_ih, b_ih, b_hh, det_img, det_p, st);
  if (rc) return rc;
  if (group == 0) {
    dim3 grid(max(1, min(tmpnn_div_up(tmpnn_div_up(g->cap_rows, TCM), 256), 8)), g->num_seqs);
    k_tile_table<<<grid, 256, 0, st>>>(g->n_rows, ix->tile128_ptr, g->cap_rows, (int4*)tile_table);
    TMPNN_LAUNCH_CHECK();
  }
  // 'diff': x = h[src] - h[dst]  ->  the far endpoint enters negated (instruction descriptor bit 13: negate A)
  k_mp_edge_tc2<<<TMPNN_SM_COUNT, TC2_THREADS, SMEM_BYTES, st>>>(
      h_in, h_out, ldh, group * H, g->src, g->dst, ix->tile128_ptr + g->num_seqs, (const int4*)tile_table,
      (const unsigned char*)edge_image, g->logit, g->score, group == 0, group == num_groups - 1, g->status, g->phys, det_img,
      det_p, ix->det_of_row, concat ? 0u : (1u << 13));
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
