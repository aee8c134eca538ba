// mp_step_tc3.cu -- the edge-row message-passing step on tcgen05, re-staged after the role timeline of
// mp_step_tc.cu (profiles/r01_tc_trace_gather_v9.txt: 8.6 k cycles per 128-row tile, of which the epilogue is
// busy 6.2 k, the MMA issue blocks a producer warp 2.8 k, and the MUFU pipe -- the real floor of the gate math,
// 3.1 k -- idles while all eight epilogue warps store / exchange / read in lock step).
//
// Same math (reference models/layers.py:84-116, heads of models/track_mpnn.py:73-75), same endpoint preparation
// as tmpnn_mp_edge_fwd_tc_pre (k_det_prepare: fp16 hi/lo image + source-side gate contribution P' per detection
// row), same shared-memory images and TMEM accumulators.  Roles, 25 warps / 800 threads / 80 registers:
//   warps 0-15   two epilogue TEAMS of 8 warps; team t owns accumulator stage t and A stage t, i.e. every other
//                tile, so one team's loads / stores / head overlap the other team's MUFU-bound gate math.  The
//                previous state is read from the h images chunk by chunk (8 registers instead of 32) because the
//                transpose buffer now lives in the stage's x images, dead since the MMAs retired
//   warps 16-23  producers: far-endpoint images by cp.async (no registers), own rows loaded one tile ahead,
//                split to fp16 hi/lo and stored into the swizzled h images
//   warp 24      MMA issuer: one thread waits for the stage, issues the 36 tcgen05.mma and commits; producers
//                never block behind the tensor pipe any more
// mbarriers per stage: full (8 producer warps), done (tcgen05.commit), gfree (8 team warps: gates done ->
// accumulators drained and h images read), xfree (8 team warps: stores done -> x images / transpose buffer free).
// Tiles are located through a table {slab's first global row, tile's first slab row, rows left} (k_tile_table).
#include <cuda.h>
#include <cstring>

#include <type_traits>

#include "tc_common.cuh"

#ifdef TMPNN_TC_TRACE
__device__ long long* g_tc3_trace = nullptr;
__device__ int g_tc3_trace_cap = 0;
#define TC3_TRACE(it_, slot_, cond_)                                                                \
  do {                                                                                              \
    if (blockIdx.x == 0 && (cond_) && g_tc3_trace && (it_) < g_tc3_trace_cap) g_tc3_trace[(it_) * 16 + (slot_)] = clock64(); \
  } while (0)
extern "C" int tmpnn_debug_set_tc3_trace(long long* buf, int cap) {
  cudaMemcpyToSymbol(g_tc3_trace, &buf, sizeof(buf));
  cudaMemcpyToSymbol(g_tc3_trace_cap, &cap, sizeof(cap));
  return 0;
}
#else
#define TC3_TRACE(it_, slot_, cond_) do { } while (0)
#endif

// Per-launch constants of the weight image, refreshed by ONE device-to-symbol copy in front of every launch (the three pieces
// are adjacent in the packed image): [0, 64) b_hn (x 2^k) | [64, 128) head weights | [128, 132) head bias, -log2e 2^-k, 2^k,
// 2 log2e 2^-k (k_pack_gru_tc's power-of-two pre-scale).  Constant-bank operands cost the epilogue no registers (it sits at
// the 72-register ceiling; the same constants read from shared memory added 28 bytes of spills), and because the epilogue's
// chunk loop is instantiated per column half every offset into the table is a compile-time constant: no shared-memory
// broadcast loads (sixteen LDS.128 per warp and tile), no registers held across the math
__constant__ float c_tc3_const[132];
#define c_tc3_tail c_tc3_const
#define c_tc3_expo (c_tc3_const + 128)

// Two measured experiments, kept behind flags (profiles/r02_ab_tc3_split.txt, r02_tc_trace_split.txt; parity green for both):
// TC3_SPLIT: the accumulator stage is released to the MMA issuer in two halves (hidden units 0-15 | 32-47 after the first
// four gate steps, the rest after the last; weight images group-major, every MMA N = 96), so that the next tile's MMAs on
// the stage overlap the second half of the gate phase.  14 % SLOWER (2.71 -> 3.10 ms): the single pair of far-endpoint
// images is now held from the first group's MMAs until the second group's retire, the copies of the next tile start that
// much later, and the issuer waits for them (xfull 0.5 k -> 1.3 k cycles, producers' xfree wait 1.6 k -> 2.3 k); a second
// pair of x images does not fit beside three own-row buffers.
// TC3_HEAD_EARLY: the upper column half of a row publishes its head partial sum right after the gate phase (before its
// stores) through an mbarrier pair, so the lower half never waits for its partner when it writes the logit: 2 % slower.
#ifndef TC3_SPLIT
#define TC3_SPLIT 0
#endif
#ifndef TC3_HEAD_EARLY
#define TC3_HEAD_EARLY 0
#endif
// TC3_ST256: the transposed read-back writes h' with sm_100's 256-bit stores (8 rows x 128 B per instruction instead of 4):
// 1.4 % slower (2.80 -> 2.84 ms, profiles/r02_ab_tc3_st256.txt) -- the stores were full lines already; where the 256-bit
// store pays is the training kernel's row-strided gate stores (mp_step_tc.cu: 150 -> 103 us)
#ifndef TC3_ST256
#define TC3_ST256 0
#endif
// TC3_RCP5: one reciprocal for the update gate and the tanh together (5 MUFU per element instead of 6, see the gate loop): parity
// green, 1 % slower (2.80 -> 2.82 ms, profiles/r02_ab_tc3_rcp5.txt; 136 instead of 88 bytes of spills) -- 17 % fewer MUFU
// instructions buy nothing, i.e. the MUFU pipe is not what bounds the gate phase either
#ifndef TC3_RCP5
#define TC3_RCP5 0
#endif
// TC3_TMA: the far-endpoint images are fetched by the TMA engine (cp.async.bulk.tensor tile::gather4: four image rows per
// instruction, written by the async proxy in the 128B-swizzled layout) instead of 2 048 16-byte cp.async per tile
#ifndef TC3_TMA
#define TC3_TMA 1
#endif
#ifndef TC3_XFENCE
#define TC3_XFENCE 0
#endif
// TC3_XSPLIT (with TC3_TMA): the lo and the hi image of the far endpoints are handed over separately -- the x_lo . B_hi term
// goes first, so the lo image is released after 4 MMAs, its gathers for the next tile start early, and the next tile's first
// four MMAs run while its hi gathers are still landing
#ifndef TC3_XSPLIT
#define TC3_XSPLIT 1
#endif
#if TC3_XSPLIT && (!TC3_TMA || TC3_SPLIT || TC3_HFIRST)
#error "TC3_XSPLIT needs TC3_TMA and excludes TC3_SPLIT / TC3_HFIRST"
#endif
// TC3_HFIRST: own-row MMAs in front of the far-endpoint MMAs (see the issuer): parity green, 6-7 % slower with the TMA gathers too
// (2.78 -> 2.97 ms, profiles/r02_ab_tc3_hfirst.txt) -- the x images are released 12 MMAs later, the next gathers start later
#ifndef TC3_HFIRST
#define TC3_HFIRST 0
#endif
// TC3_LD1: single-set accumulator drain (see the gate loop): the next step's TMEM loads issued under the current step's MUFU
// chains, P' / previous-state loads in front of tcgen05.wait::ld -- no difference (2.76 vs 2.76 ms,
// profiles/r02_ab_tc3_ld1.txt): neither the TMEM-load nor the L1 latency at the top of a step is what bounds the gate phase
#ifndef TC3_LD1
#define TC3_LD1 0
#endif
#if TC3_LD1 && TC3_SPLIT
#error "TC3_LD1 and TC3_SPLIT are separate experiments"
#endif

namespace {

constexpr int EPI3 = 16, PROD3 = 8;
constexpr int ISSUER3 = EPI3 + PROD3;              // warp 24
constexpr int TC3_THREADS = 32 * (EPI3 + PROD3 + 1);  // 800 -> 80 registers per thread
constexpr int H_BUF = 2 * A_PART;                      // one own-row buffer: [h_hi | h_lo]
// barriers (8 bytes each, two stages): full, done, gfree, xfree; then the TMEM pointer
// head partial sums: team 0 in the r | z bias slots of the image (unused here: folded into P'), team 1 at OFF_DOT

__global__ void __launch_bounds__(256)
k_tile_table(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ tile128_ptr, int cap_rows,
             int4* __restrict__ tab) {
  const int s = blockIdx.y;
  const int t0 = tile128_ptr[s], nt = tile128_ptr[s + 1] - t0, n = n_rows[s];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nt; j += gridDim.x * blockDim.x)
    tab[t0 + j] = make_int4(s * cap_rows, j * TCM, n - j * TCM, 0);
}

// 25 warps -> 7 on one scheduler, whose 16 K registers allow 72 per thread (ptxas derives this from the bounds)
__global__ void __launch_bounds__(TC3_THREADS, 1)
k_mp_edge_tc3(const float* __restrict__ h_in, float* __restrict__ h_out, int ldh, int col,
              const int32_t* __restrict__ src, const int32_t* __restrict__ dst, const int32_t* __restrict__ n_tiles,
              const int4* __restrict__ tab, const unsigned char* __restrict__ image, float* __restrict__ logit,
              float* __restrict__ score, int first_group, int last_group, int32_t* __restrict__ status,
              const int32_t* __restrict__ phys, const float* __restrict__ det_img, const float* __restrict__ det_p,
              const int32_t* __restrict__ det_of_row, uint32_t xflags, const __grid_constant__ CUtensorMap tmx) {
  extern __shared__ unsigned char smem_dyn[];
  const int total = *n_tiles;
  if ((int)blockIdx.x >= total) return;  // uniform: whole CTA leaves before touching TMEM / barriers
  // detection mode (tmpnn_mp_det_fwd_tc): the tiles cover the detection segments, the "far endpoint" of a row is the row
  // itself (det_img then holds the images of the detections' aggregates), P' is one bias row and every row in range counts
  const bool detm = (xflags & 1u) != 0;
  xflags &= ~1u;
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // shared-memory images behind the weights: ONE pair of far-endpoint images and THREE pairs of own-row images
  const uint32_t x_u = sm_u + OFF_A;
  unsigned char* const h_img = sm + OFF_A + 2 * A_PART;  // + hb * H_BUF: [h_hi | h_lo] of buffer hb
  const uint32_t bar_hfull = sm_u + OFF_BAR, bar_hfree = bar_hfull + 24, bar_xfull = bar_hfull + 48, bar_xfree = bar_hfull + 56,
                 bar_done = bar_hfull + 64, bar_gfree = bar_hfull + 80, bar_doneB = bar_hfull + 96, bar_gfreeB = bar_hfull + 112;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 128);
#if TC3_XSPLIT
  const uint32_t bar_xfull_hi = bar_doneB, bar_xfree_hi = bar_doneB + 8;   // the slots of the (excluded) TC3_SPLIT experiment
#endif

  // resident weight image (generic-proxy stores, made visible to the async proxy below)
  {
    const uint4* gsrc = reinterpret_cast<const uint4*>(image);
    uint4* sdst = reinterpret_cast<uint4*>(sm);
#if TC3_HEAD_EARLY
    constexpr int IMG_CHUNKS = 4 * B_BYTES / 16;  // the bias / head slots are not read from shared memory (constant bank): they
                                                  // hold the head partial sums and their barriers
#else
    constexpr int IMG_CHUNKS = IMAGE_BYTES / 16;
#endif
#if TC3_SPLIT
    // weight images group-major: packed row 64 t + 32 h + 16 g + q (gate t, hidden unit 32 h + 16 g + q) goes to row
    // 96 g + 32 t + 16 h + q.  Both rows have the same (row & 7), i.e. the same 128B-swizzle phase: a plain row move
    for (int i = threadIdx.x; i < IMG_CHUNKS; i += TC3_THREADS) {
      int o = i;
      if (i < 4 * B_BYTES / 16) {
        const int img = i / (B_BYTES / 16), rem = i - img * (B_BYTES / 16);
        const int row = rem >> 3, c = rem & 7;
        const int t = row >> 6, h = (row >> 5) & 1, g = (row >> 4) & 1, q = row & 15;
        o = img * (B_BYTES / 16) + ((96 * g + 32 * t + 16 * h + q) << 3) + c;
      }
      sdst[o] = __ldg(gsrc + i);
    }
#else
    for (int i = threadIdx.x; i < IMG_CHUNKS; i += TC3_THREADS) sdst[i] = __ldg(gsrc + i);
#endif
  }
  if (threadIdx.x == 0) {
    for (int b = 0; b < 3; ++b) {
      mbar_init(bar_hfull + 8 * b, PROD3);  // one arrive per producer warp: own rows written
      mbar_init(bar_hfree + 8 * b, 8);      // one arrive per warp of the tile's team: previous state + transpose buffer read back
    }
#if TC3_XSPLIT
    mbar_init(bar_xfull, 4 * PROD3);        // lo images: four issuing lanes per producer warp, 512 B each
    mbar_init(bar_xfull_hi, 4 * PROD3);     // hi images
    mbar_init(bar_xfree_hi, 1);
#elif TC3_TMA
    mbar_init(bar_xfull, 4 * PROD3);        // four issuing lanes per producer warp, each expecting the 1 KB of its two gathers
#else
    mbar_init(bar_xfull, 32 * PROD3);       // one arrive per producer THREAD: its far-endpoint copies landed
#endif
    mbar_init(bar_xfree, 1);                // tcgen05.commit behind the far-endpoint MMAs: x images reusable
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_done + 8 * s, 1);       // tcgen05.commit behind the own-row MMAs: accumulators complete
      mbar_init(bar_gfree + 8 * s, 8);      // one arrive per warp of the stage's team: accumulators drained
#if TC3_SPLIT
      mbar_init(bar_doneB + 8 * s, 1);      // TC3_SPLIT: the same pair for the second hidden group of the stage
      mbar_init(bar_gfreeB + 8 * s, 8);
#endif
    }
#if TC3_HEAD_EARLY
    for (int q = 0; q < 16; ++q) mbar_init(sm_u + OFF_HEADW + 8 * q, 1);  // head partials published / consumed, per (team, quadrant)
#endif
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
  // the own-row MMAs only ever accumulate: the h_n columns of both accumulator stages start at zero (each epilogue
  // warp owns the 32 lanes x 32 columns it will later drain and re-zero)
#if TC3_SPLIT
  if (warp < EPI3) {
    const uint32_t tz = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 3) * 256 + 16 * ((warp & 7) >> 2));
    tmem_zero16(tz);
    tmem_zero16(tz + 128);
  }
#else
  if (warp < EPI3) tmem_zero32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 3) * 256 + 32 * ((warp & 7) >> 2) + (TC3_HFIRST ? 192 : 0)));
#endif
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // tiles past the end repeat the last one (their loads are simply unused)
  auto ldtab = [&](int tile) { return __ldg(tab + min(tile, total - 1)); };

  if (warp == ISSUER3) {
    // ================= MMA issuer =================
    if (elect_one()) {
      int it = 0, hb = 0;
      uint32_t hphase = 0;
      for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
        const int stage = it & 1;
        const uint32_t phase = (uint32_t)(it >> 1) & 1u;
        const uint32_t d0 = tmem_base + (uint32_t)(stage * 256);
#if TC3_SPLIT
        // first hidden group: its columns were drained by the team half way through the gate phase of the tile two back
        mbar_wait(bar_gfree + 8 * stage, phase ^ 1u, status);
        TC3_TRACE(it, 5, true);
        mbar_wait(bar_xfull, (uint32_t)it & 1u, status);       // far-endpoint images landed (issued a tile ago)
        TC3_TRACE(it, 6, true);
        fence_proxy_async();  // the copies were generic-proxy writes of other threads, observed through the barrier
        tc_fence_after();
        issue_group_mma_x(sm_u, d0, x_u, xflags, 0);
        mbar_wait(bar_hfull + 8 * hb, hphase, status);         // own rows: written up to two tiles ahead
        TC3_TRACE(it, 10, true);
        tc_fence_after();
        issue_group_mma_h(sm_u, d0, x_u + 2 * A_PART + (uint32_t)hb * H_BUF, 0);
        umma_commit(bar_done + 8 * stage);  // first group ready (implies tcgen05.fence::before_thread_sync)
        // second hidden group: drained at the end of that gate phase
        mbar_wait(bar_gfreeB + 8 * stage, phase ^ 1u, status);
        tc_fence_after();
        issue_group_mma_x(sm_u, d0 + 128, x_u, xflags, 1);
        umma_commit(bar_xfree);  // x images reusable once these retire: the copies of tile it + 1 start here
        issue_group_mma_h(sm_u, d0 + 128, x_u + 2 * A_PART + (uint32_t)hb * H_BUF, 1);
        umma_commit(bar_doneB + 8 * stage);
        TC3_TRACE(it, 7, true);
#elif TC3_HFIRST
        // own-row part first (its images are written up to two tiles ahead: no wait), far-endpoint part behind it: the wait for
        // the gathered images hides under the first 12 MMAs.  The own-row part initialises h_n | r | z, the far-endpoint part
        // accumulates into r | z | i_n, whose i_n columns the epilogue re-zeroed.
        mbar_wait(bar_gfree + 8 * stage, phase ^ 1u, status);
        TC3_TRACE(it, 5, true);
        mbar_wait(bar_hfull + 8 * hb, hphase, status);
        tc_fence_after();
        {
          const uint32_t h_u = x_u + 2 * A_PART + (uint32_t)hb * H_BUF;
          const uint32_t ah[3] = {h_u, h_u + A_PART, h_u};
          const uint32_t bh[3] = {sm_u + OFF_BH_HI, sm_u + OFF_BH_HI, sm_u + OFF_BH_LO};
          uint32_t acc = 0;
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              umma_f16(d0, umma_desc(ah[t] + 32 * j), umma_desc(bh[t] + 32 * j), umma_idesc(192), acc);
              acc = 1;
            }
        }
        TC3_TRACE(it, 10, true);
        mbar_wait(bar_xfull, (uint32_t)it & 1u, status);
        TC3_TRACE(it, 6, true);
        tc_fence_after();
        {
          const uint32_t ax[3] = {x_u, x_u + A_PART, x_u};
          const uint32_t bx[3] = {sm_u + OFF_BX_HI, sm_u + OFF_BX_HI, sm_u + OFF_BX_LO};
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_f16(d0 + 64, umma_desc(ax[t] + 32 * j), umma_desc(bx[t] + 32 * j), umma_idesc(192) | xflags, 1);
        }
        umma_commit(bar_xfree);
        umma_commit(bar_done + 8 * stage);
        TC3_TRACE(it, 7, true);
#else
        mbar_wait(bar_gfree + 8 * stage, phase ^ 1u, status);  // accumulator stage drained by the tile two back
        TC3_TRACE(it, 5, true);
#if TC3_XSPLIT
        {
          // x_lo . B_hi first (initialises r | z | i_n), lo image released; then x_hi . B_hi and x_hi . B_lo, hi image released
          mbar_wait(bar_xfull, (uint32_t)it & 1u, status);
          TC3_TRACE(it, 6, true);
          tc_fence_after();
          uint32_t acc = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_f16(d0 + 64, umma_desc(x_u + A_PART + 32 * j), umma_desc(sm_u + OFF_BX_HI + 32 * j), umma_idesc(192) | xflags, acc);
            acc = 1;
          }
          umma_commit(bar_xfree);
          mbar_wait(bar_xfull_hi, (uint32_t)it & 1u, status);
          tc_fence_after();
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_f16(d0 + 64, umma_desc(x_u + 32 * j), umma_desc(sm_u + (t ? OFF_BX_LO : OFF_BX_HI) + 32 * j), umma_idesc(192) | xflags, 1);
          umma_commit(bar_xfree_hi);
        }
        if (false) {
#else
        {
#endif
        mbar_wait(bar_xfull, (uint32_t)it & 1u, status);       // far-endpoint images landed (issued a tile ago)
        TC3_TRACE(it, 6, true);
#if !TC3_TMA || TC3_XFENCE
        fence_proxy_async();  // cp.async copies are generic-proxy writes of other threads, observed through the barrier (the
                              // TMA gathers write through the async proxy themselves: no proxy fence in front of the MMAs)
#endif
        tc_fence_after();
        issue_tile_mma_x_first(sm_u, d0, x_u, xflags);
        umma_commit(bar_xfree);  // x images reusable once these retire: the copies of tile it + 1 start here
        }
        // own rows: written up to two tiles ahead into the third buffer, so this wait is normally over already and
        // a team's hand-over from one tile to its next is the 36 MMAs only
        mbar_wait(bar_hfull + 8 * hb, hphase, status);
        TC3_TRACE(it, 10, true);
        tc_fence_after();
        issue_tile_mma_h_second(sm_u, d0, x_u + 2 * A_PART + (uint32_t)hb * H_BUF);
        umma_commit(bar_done + 8 * stage);  // accumulators ready (implies tcgen05.fence::before_thread_sync)
        TC3_TRACE(it, 7, true);
#endif
        if (++hb == 3) { hb = 0; hphase ^= 1u; }
      }
    }
  } else if (warp >= EPI3) {
    // ================= producers: 8 warps, 16 lanes per row, rows g + 16 p =================
    const int pt = threadIdx.x - 32 * EPI3;
    const int g = pt >> 4, l = pt & 15, gl0 = lane & 16;
    const uint32_t FULL = 0xffffffffu;
    const bool tr = pt == 0;
    // all addressing in units of float4 from h_in: (global row) * ldh4 + col4 + l fits 32 bits (checked on the host)
    const float4* __restrict__ h4p = reinterpret_cast<const float4*>(h_in);
    const uint32_t ldh4 = (uint32_t)ldh >> 2, cl4 = ((uint32_t)col >> 2) + (uint32_t)l;
    const bool dfr = phys != nullptr;      // deferred compaction: own rows at GLOBAL physical rows
    const int idx_row = g + 16 * (l & 7);  // lanes l and l + 8 of a group keep the far endpoint / physical row of tile row g + 16 (l & 7)
    // detection images have the geometry of h: chunk l of the 256 B image of row R sits at
    // (R * ldh + col) * 4 + 16 l; chunks 0-7 are the hi halves (K order), 8-15 the lo halves
    const unsigned char* __restrict__ imgb = reinterpret_cast<const unsigned char*>(det_img) + (size_t)col * 4 + 16 * l;
    const size_t row_bytes = (size_t)ldh * 4;
    const uint32_t x_dst0 = sm_u + OFF_A + (uint32_t)(l >> 3) * A_PART + sw128(g, l & 7);  // + 2048 p: row g + 16 p
    const uint32_t h_off0 = sw128(g, l >> 1) + ((l & 1) << 3);
    // rows past the end of the slab repeat its last row; everything they produce is masked by the epilogue
#if TC3_TMA
    // lane k (mod 16) of producer warp wq keeps the GLOBAL image row of the far endpoint of tile row 16 wq + k
    const int wq = warp - EPI3;
    const int colh = 2 * col;   // first fp16 column of this group's hi half in an image row (the lo half follows 64 further)
    auto ld_idx = [&](const int4 T) {
      const int rr = T.y + min(16 * wq + (lane & 15), T.z - 1);
      return T.x + (detm ? rr : max(__ldg(dst + T.x + rr), 0));   // -1 (a detection row inside the tile): any valid row will do
    };
#else
    auto ld_idx = [&](const int4 T) {
      const int rr = T.y + min(idx_row, T.z - 1);
      return detm ? rr : __ldg(dst + T.x + rr);
    };
#endif
    auto ld_phys = [&](const int4 T) { return dfr ? __ldg(phys + T.x + T.y + min(idx_row, T.z - 1)) : 0; };
#if TC3_TMA
    auto issue_x = [&](uint32_t, int iv, bool xwait = false, uint32_t xpar = 0) {
      (void)xwait; (void)xpar;
      const int q = 4 * (lane & 3);
      const int r0 = __shfl_sync(FULL, iv, q), r1 = __shfl_sync(FULL, iv, q + 1), r2 = __shfl_sync(FULL, iv, q + 2),
                r3 = __shfl_sync(FULL, iv, q + 3);
#if TC3_XSPLIT
      // lo image first (released first by the issuer), then the hi image: each behind its own xfree barrier
      const uint32_t d = x_u + (uint32_t)(16 * wq + 4 * lane) * 128u;
      if (xwait) mbar_wait(bar_xfree, xpar, status);
      if (lane < 4) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_xfull), "r"(512u) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(d + A_PART), "l"(&tmx), "r"(bar_xfull), "r"(colh + 64), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
      }
      if (xwait) mbar_wait(bar_xfree_hi, xpar, status);
      if (lane < 4) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_xfull_hi), "r"(512u) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(d), "l"(&tmx), "r"(bar_xfull_hi), "r"(colh), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
      }
#else
      if (lane < 4) {   // rows 16 wq + 4 lane .. + 3: rows are 128 B apart in the swizzled image (8-row groups of 1024 B)
        const uint32_t d = x_u + (uint32_t)(16 * wq + 4 * lane) * 128u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_xfull), "r"(1024u) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(d), "l"(&tmx), "r"(bar_xfull), "r"(colh), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
            ::"r"(d + A_PART), "l"(&tmx), "r"(bar_xfull), "r"(colh + 64), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
      }
#endif
    };
    auto issue_x_cp = [&](uint32_t b, int iv) {
#else
    auto issue_x = [&](uint32_t b, int iv) {
#endif
      const uint32_t s0 = x_dst0;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int d = __shfl_sync(FULL, iv, gl0 + p);  // -1 for detection rows inside the tile: any valid row will do
        const unsigned char* sp = imgb + (size_t)(b + (uint32_t)max(d, 0)) * row_bytes;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 2048u * p), "l"(sp) : "memory");
      }
      // completion is signalled by the copy engine itself (one arrival per thread once its copies have landed)
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_xfull) : "memory");
    };
    int4 T0 = ldtab(blockIdx.x), T1 = ldtab(blockIdx.x + stride), T2 = ldtab(blockIdx.x + 2 * stride);
    int pw1 = ld_phys(T1), i1 = ld_idx(T1);
    uint2 hh[8], hl[8];
    {
      // far-endpoint images and own rows of the first tile
      const int i0 = ld_idx(T0), pw0 = ld_phys(T0);
      issue_x((uint32_t)T0.x, i0);
      float amax = 0.f;
      float4 own[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int sp = __shfl_sync(FULL, pw0, gl0 + p);
        const uint32_t orow = dfr ? (uint32_t)sp : (uint32_t)(T0.x + T0.y + min(g + 16 * p, T0.z - 1));
        own[p] = __ldg(h4p + orow * ldh4 + cl4);
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) split4(own[p], hh[p], hl[p], amax);
      if (amax > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);
    }
    int it = 0, hb = 0;
    uint32_t hphase = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
      unsigned char* hbuf = h_img + hb * H_BUF;
      const int4 T3 = ldtab(tile + 3 * stride);      // in flight for a whole tile
      const int i2 = ld_idx(T2), pw2 = ld_phys(T2);  // T2 landed a tile ago; consumed a tile from now
      TC3_TRACE(it, 2, tr);
      // own-row images, buffer it % 3: previous state and transpose buffer of the tile three back -- free long ago, so
      // the rows (split while waiting) go in one or two tiles before their MMAs can start
      mbar_wait(bar_hfree + 8 * hb, hphase ^ 1u, status);
      TC3_TRACE(it, 3, tr);
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const uint32_t off = h_off0 + 2048u * p;
        *reinterpret_cast<uint2*>(hbuf + off) = hh[p];
        *reinterpret_cast<uint2*>(hbuf + A_PART + off) = hl[p];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfull + 8 * hb);
      TC3_TRACE(it, 4, tr);
      // own rows of the next tile: HBM latency overlaps the copy issue below (unconditional, clamped addresses)
      float4 own[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const int sp = __shfl_sync(FULL, pw1, gl0 + p);
        const uint32_t orow = dfr ? (uint32_t)sp : (uint32_t)(T1.x + T1.y + min(g + 16 * p, T1.z - 1));
        own[p] = __ldg(h4p + orow * ldh4 + cl4);
      }
      // far-endpoint images of the next tile: the single pair of x images is free as soon as this tile's MMAs on it retire
      TC3_TRACE(it, 0, tr);
#if TC3_XSPLIT
      issue_x((uint32_t)T1.x, i1, true, (uint32_t)it & 1u);   // waits for the lo / hi image in turn
      TC3_TRACE(it, 1, tr);
#else
      mbar_wait(bar_xfree, (uint32_t)it & 1u, status);
      TC3_TRACE(it, 1, tr);
      issue_x((uint32_t)T1.x, i1);
#endif
      float amax = 0.f;
#pragma unroll
      for (int p = 0; p < 8; ++p) split4(own[p], hh[p], hl[p], amax);
      if (amax > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);  // fp16 split would overflow: use the FMA path
      T0 = T1; T1 = T2; T2 = T3; pw1 = pw2; i1 = i2;
      if (++hb == 3) { hb = 0; hphase ^= 1u; }
    }
#if TC3_TMA
    mbar_wait(bar_xfull, (uint32_t)it & 1u, status);  // the gathers issued for the tile past the end have landed
#if TC3_XSPLIT
    mbar_wait(bar_xfull_hi, (uint32_t)it & 1u, status);
#endif
    (void)issue_x_cp;
#else
    asm volatile("cp.async.wait_all;" ::: "memory");  // copies issued for tiles past the end
#endif
  } else {
    // ================= epilogue: two teams of 8 warps, team t takes tiles it = t, t + 2, ... (stage t) =================
    const int team = warp >> 3, w8 = warp & 7;
    const int quad = w8 & 3, half = w8 >> 2;  // quad == warp % 4: the TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;           // row of the tile == TMEM lane
    const int c0 = 32 * half;                 // this warp's columns of every gate: [c0, c0 + 32)
    const int stage = team;
    const bool tr = threadIdx.x == 0 || threadIdx.x == 256;
    const float headb = c_tc3_expo[0];
#if TC3_HEAD_EARLY
    float* dot_part = reinterpret_cast<float*>(sm + OFF_BIAS + 512 * team);
    const uint32_t bar_dfull = sm_u + OFF_HEADW + 8 * (team * 4 + quad), bar_dfree = bar_dfull + 64;
#else
    float* dot_part = reinterpret_cast<float*>(sm + (team ? OFF_DOT : OFF_BIAS));
#endif
    const f32x2 NLOG2E2 = pk2(c_tc3_expo[1], c_tc3_expo[1]), TWOLOG2E2 = pk2(c_tc3_expo[3], c_tc3_expo[3]), ONE2 = pk2(1.0f, 1.0f);
    const f32x2 NTWO2 = pk2(-2.0f, -2.0f), NONE2 = pk2(-1.0f, -1.0f);
#if TC3_SPLIT
    // this warp's 16 columns inside each gate block of either hidden group: group g at + 128 g, gates at + 0 / 32 / 64 / 96
    const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(stage * 256 + 16 * half);
#else
    const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(stage * 256 + c0);
#endif
    const int bar_id = 1 + team * 4 + quad;
    auto ld_src = [&](const int4 T) { return detm ? 0 : __ldg(src + T.x + (T.z - r > 0 ? T.y + r : 0)); };  // clamped to the slab's first row
    const int first = blockIdx.x + team * stride, step2 = 2 * stride;
    int4 T0 = ldtab(first), T1 = ldtab(first + step2);
    int srcv = ld_src(T0);
    uint32_t n = 0;  // tiles this team has done
    int hb = team;   // own-row buffer of tile it = 2 n + team: it % 3
    for (int tile = first; tile < total; tile += step2, ++n) {
      const uint32_t phase = n & 1u;
      unsigned char* const h_hi = h_img + hb * H_BUF;
      unsigned char* const h_lo = h_hi + A_PART;
      const int it = 2 * (int)n + team;
      const size_t row_cur = (size_t)T0.x + T0.y + r;
      const bool valid = T0.z - r > 0 && srcv >= 0;
      const int ks = detm ? 0 : __ldg(det_of_row + T0.x + max(srcv, 0));  // srcv landed during the team's previous tile
      const int4 T2 = ldtab(tile + 2 * step2);
      const int srcv1 = ld_src(T1);  // T1 landed a tile ago; consumed by the team's next tile
      // the first step's P' terms are fetched before the wait (cold lines: an L2 round trip that would otherwise sit
      // in front of the first gate of the tile); the later steps hit the same three lines in L1
      const float* __restrict__ pp_pre = det_p + (size_t)max(ks, 0) * 192 + c0;
      const ulonglong2 br0 = __ldg(reinterpret_cast<const ulonglong2*>(pp_pre));
      const ulonglong2 bz0 = __ldg(reinterpret_cast<const ulonglong2*>(pp_pre + H));
      const ulonglong2 bi0 = __ldg(reinterpret_cast<const ulonglong2*>(pp_pre + 2 * H));
      TC3_TRACE(it, 8, tr);
      mbar_wait(bar_done + 8 * stage, phase, status);
#if TC3_SPLIT
      // both hidden groups: the gate loop overwrites the own-row images (its transpose buffer) from its first step on, so
      // the second group's own-row MMAs must have retired too (issued a whole store / head phase ago: no wait in practice)
      mbar_wait(bar_doneB + 8 * stage, phase, status);
#endif
      tc_fence_after();
      TC3_TRACE(it, 9, tr);
      const float* __restrict__ pp = det_p + (size_t)max(ks, 0) * 192 + c0;  // this row's source contribution
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      float* out0 = h_out + (row_cur - lane) * ldh + col + c0;  // first row of this warp's quadrant
      f32x2 dot2 = 0ull;
      auto gate_chunks = [&](auto half_c) {
      constexpr int HALF = decltype(half_c)::value;
#if TC3_LD1
      // Accumulator drain in steps of 4 columns x 4 gates out of ONE set of 16 registers: a step first folds its accumulators
      // into the four pre-activations (8 packed registers), then issues the TMEM loads of the NEXT step into the registers it
      // has just freed, so they land under the step's MUFU chains; everything a step reads from memory (P', previous
      // state) is requested in front of tcgen05.wait::ld.  (wait::ld waits for ALL outstanding loads: the two-set form,
      // which issued the loads of step s + 2 at the end of step s, waited for them at the top of step s + 1.)
      uint32_t A[16];
      auto ldstep = [&](int s, uint32_t* a) {
        const uint32_t cb = t0 + (uint32_t)((s >> 1) * 8 + (s & 1) * 4);
        tmem_ld4u(cb + 64, a);        // r
        tmem_ld4u(cb + 128, a + 4);   // z
        tmem_ld4u(cb + 192, a + 8);   // i_n
        tmem_ld4u(cb, a + 12);        // h_n
      };
      ldstep(0, A);
      f32x2 hp[4];
      uint32_t off = 0;
#if TC3_LD1 == 2
      // P' one step ahead: the three 16-byte loads of step s + 1 are issued at the top of step s (the 28 KB L1 beside 227 KB of
      // shared memory rarely holds the lines: an L2 round trip in front of every step's first FMA otherwise)
      ulonglong2 nr = br0, nz = bz0, ni = bi0;
#endif
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int ch = s >> 1, v = s & 1;
#if TC3_LD1 == 2
        const ulonglong2 br = nr, bz = nz, bi = ni;
        if (s + 1 < 8) {
          const int o1 = ((s + 1) >> 1) * 8 + 4 * ((s + 1) & 1);
          nr = __ldg(reinterpret_cast<const ulonglong2*>(pp + o1));
          nz = __ldg(reinterpret_cast<const ulonglong2*>(pp + H + o1));
          ni = __ldg(reinterpret_cast<const ulonglong2*>(pp + 2 * H + o1));
        }
#else
        const ulonglong2 br = s == 0 ? br0 : __ldg(reinterpret_cast<const ulonglong2*>(pp + ch * 8 + 4 * v));
        const ulonglong2 bz = s == 0 ? bz0 : __ldg(reinterpret_cast<const ulonglong2*>(pp + H + ch * 8 + 4 * v));
        const ulonglong2 bi = s == 0 ? bi0 : __ldg(reinterpret_cast<const ulonglong2*>(pp + 2 * H + ch * 8 + 4 * v));
#endif
        if (v == 0) {
          off = sw128(r, 4 * half + ch);
          const uint4 vh = *reinterpret_cast<const uint4*>(h_hi + off);
          const uint4 vl = *reinterpret_cast<const uint4*>(h_lo + off);
          const __half2* ph = reinterpret_cast<const __half2*>(&vh);
          const __half2* pl = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 fh = __half22float2(ph[i]), fl = __half22float2(pl[i]);
            hp[i] = add2(pk2(fh.x, fh.y), pk2(fl.x, fl.y));
          }
          __syncwarp();
        }
        const int jc = 32 * HALF + 8 * ch + 4 * v;  // compile-time after unrolling
        const ulonglong2 bh = make_ulonglong2(pk2(c_tc3_tail[jc], c_tc3_tail[jc + 1]), pk2(c_tc3_tail[jc + 2], c_tc3_tail[jc + 3]));
        const ulonglong2 hw = make_ulonglong2(pk2(c_tc3_tail[64 + jc], c_tc3_tail[64 + jc + 1]), pk2(c_tc3_tail[64 + jc + 2], c_tc3_tail[64 + jc + 3]));
        tmem_ld_wait();
        f32x2 xr[2], xz[2], hb[2], ib[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = 2 * e;
          xr[e] = fma2(pk2u(A[i], A[i + 1]), NLOG2E2, e ? br.y : br.x);
          xz[e] = fma2(pk2u(A[4 + i], A[5 + i]), NLOG2E2, e ? bz.y : bz.x);
          hb[e] = add2(pk2u(A[12 + i], A[13 + i]), e ? bh.y : bh.x);
          ib[e] = add2(pk2u(A[8 + i], A[9 + i]), e ? bi.y : bi.x);
        }
        if (s + 1 < 8) ldstep(s + 1, A);
        f32x2 o[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const f32x2 rg = rcp_2(add2(ex2_2(xr[e]), ONE2));
          const f32x2 zg = rcp_2(add2(ex2_2(xz[e]), ONE2));
          const f32x2 u = fma2(rg, hb[e], ib[e]);
          const f32x2 ng = fma2(rcp_2(add2(ex2_2(mul2(u, TWOLOG2E2)), ONE2)), NTWO2, ONE2);
          const f32x2 ov = fma2(zg, fma2(ng, NONE2, hp[2 * v + e]), ng);
          o[e] = ov;
          dot2 = fma2(ov, e ? hw.y : hw.x, dot2);
        }
        if (v == 0) *reinterpret_cast<ulonglong2*>(h_hi + off) = make_ulonglong2(o[0], o[1]);
        else        *reinterpret_cast<ulonglong2*>(h_lo + sw128(r ^ 4, 4 * half + ch)) = make_ulonglong2(o[0], o[1]);
      }
      };
#else
      // software-pipelined accumulator drain: two sets of 4 columns x 4 gates; the TMEM loads of step s + 2 are issued
      // as soon as step s has consumed its set, so they land during step s + 1 (same 32 accumulator registers)
      uint32_t A[2][16];
      auto ldstep = [&](int s, uint32_t* a) {
#if TC3_SPLIT
        const uint32_t cb = t0 + (uint32_t)((s >> 2) * 128 + (s & 3) * 4);   // steps 0-3: first hidden group, 4-7: second
        tmem_ld4u(cb + 32, a);        // r
        tmem_ld4u(cb + 64, a + 4);    // z
        tmem_ld4u(cb + 96, a + 8);    // i_n
        tmem_ld4u(cb, a + 12);        // h_n
#else
        const uint32_t cb = t0 + (uint32_t)((s >> 1) * 8 + (s & 1) * 4);
        tmem_ld4u(cb + 64, a);        // r
        tmem_ld4u(cb + 128, a + 4);   // z
        tmem_ld4u(cb + 192, a + 8);   // i_n
        tmem_ld4u(cb, a + 12);        // h_n
#endif
      };
      ldstep(0, A[0]);
      ldstep(1, A[1]);
      f32x2 hp[4];
      uint32_t off = 0;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        const int ch = s >> 1, v = s & 1;
        uint32_t* a = A[s & 1];
        if (v == 0) {
          off = sw128(r, 4 * half + ch);
          const uint4 vh = *reinterpret_cast<const uint4*>(h_hi + off);
          const uint4 vl = *reinterpret_cast<const uint4*>(h_lo + off);
          const __half2* ph = reinterpret_cast<const __half2*>(&vh);
          const __half2* pl = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 fh = __half22float2(ph[i]), fl = __half22float2(pl[i]);
            hp[i] = add2(pk2(fh.x, fh.y), pk2(fl.x, fl.y));
          }
          __syncwarp();
        }
        tmem_ld_wait();
        const ulonglong2 br = s == 0 ? br0 : __ldg(reinterpret_cast<const ulonglong2*>(pp + ch * 8 + 4 * v));
        const ulonglong2 bz = s == 0 ? bz0 : __ldg(reinterpret_cast<const ulonglong2*>(pp + H + ch * 8 + 4 * v));
        const ulonglong2 bi = s == 0 ? bi0 : __ldg(reinterpret_cast<const ulonglong2*>(pp + 2 * H + ch * 8 + 4 * v));
        const int jc = 32 * HALF + 8 * ch + 4 * v;  // compile-time after unrolling
        const ulonglong2 bh = make_ulonglong2(pk2(c_tc3_tail[jc], c_tc3_tail[jc + 1]), pk2(c_tc3_tail[jc + 2], c_tc3_tail[jc + 3]));
        const ulonglong2 hw = make_ulonglong2(pk2(c_tc3_tail[64 + jc], c_tc3_tail[64 + jc + 1]), pk2(c_tc3_tail[64 + jc + 2], c_tc3_tail[64 + jc + 3]));
        f32x2 o[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = 2 * e;
          const f32x2 rg = rcp_2(add2(ex2_2(fma2(pk2u(a[i], a[i + 1]), NLOG2E2, e ? br.y : br.x)), ONE2));
#if TC3_RCP5
          // five MUFU per element instead of six: z = 1 / (1 + B) and n = (E - 1) / (E + 1) share ONE reciprocal,
          //   h' = n + z (h - n) = ((E - 1) B + h (E + 1)) / ((E + 1) (1 + B)),   B = 2^(-log2e x_z), E = 2^(2 log2e u),
          // both exponents clamped at 57 (2^57 = 1.4e17: z < 1e-17, 1 - n < 2e-17, the product stays finite)
          const f32x2 B = ex2_2(min2c(fma2(pk2u(a[4 + i], a[5 + i]), NLOG2E2, e ? bz.y : bz.x), 57.0f));
          const f32x2 u = fma2(rg, add2(pk2u(a[12 + i], a[13 + i]), e ? bh.y : bh.x), add2(pk2u(a[8 + i], a[9 + i]), e ? bi.y : bi.x));
          const f32x2 E = ex2_2(min2c(mul2(u, TWOLOG2E2), 57.0f));
          const f32x2 t2 = add2(E, ONE2);
          const f32x2 num = fma2(add2(E, NONE2), B, mul2(hp[2 * v + e], t2));
          const f32x2 ov = mul2(num, rcp_2(mul2(t2, add2(B, ONE2))));
#else
          const f32x2 zg = rcp_2(add2(ex2_2(fma2(pk2u(a[4 + i], a[5 + i]), NLOG2E2, e ? bz.y : bz.x)), ONE2));
          const f32x2 u = fma2(rg, add2(pk2u(a[12 + i], a[13 + i]), e ? bh.y : bh.x), add2(pk2u(a[8 + i], a[9 + i]), e ? bi.y : bi.x));
          const f32x2 ng = fma2(rcp_2(add2(ex2_2(mul2(u, TWOLOG2E2)), ONE2)), NTWO2, ONE2);
          const f32x2 ov = fma2(zg, fma2(ng, NONE2, hp[2 * v + e]), ng);
#endif
          o[e] = ov;
          dot2 = fma2(ov, e ? hw.y : hw.x, dot2);
        }
        if (s + 2 < 8) ldstep(s + 2, a);
        if (v == 0) *reinterpret_cast<ulonglong2*>(h_hi + off) = make_ulonglong2(o[0], o[1]);
        else        *reinterpret_cast<ulonglong2*>(h_lo + sw128(r ^ 4, 4 * half + ch)) = make_ulonglong2(o[0], o[1]);
#if TC3_SPLIT
        if (s == 3) {
          // first hidden group drained (the loads of steps 4 and 5 in flight read the other group): re-zero this warp's
          // h_n columns and hand the group to the issuer -- its MMAs for the team's next tile run under steps 4-7
          tmem_zero16(t0);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_gfree + 8 * stage);
        }
#endif
      }
      };
#endif
      if (half == 0) gate_chunks(std::integral_constant<int, 0>{});
      else gate_chunks(std::integral_constant<int, 1>{});
      float dot;
      {
        float d0, d1;
        up2(dot2, d0, d1);
        dot = d0 + d1;
      }
      // accumulator stage drained: re-zero this warp's h_n columns for the accumulate-only own-row MMAs
#if TC3_SPLIT
      tmem_zero16(t0 + 128);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfreeB + 8 * stage);  // second hidden group drained
#else
      tmem_zero32(t0 + (TC3_HFIRST ? 192u : 0u));   // the columns the accumulate-only half of the next tile's MMAs adds into
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_gfree + 8 * stage);  // accumulator stage drained: the next far-endpoint MMAs may start
#endif
      TC3_TRACE(it, 11, tr);
#if TC3_HEAD_EARLY
      if (half == 1) {
        mbar_wait(bar_dfree, phase ^ 1u, status);  // the previous tile's partials have been read (normally long ago)
        dot_part[r] = dot;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dfull);
      }
#endif
      // transposed read-back: each store instruction writes 4 rows x 128 B (full lines).  Lane (rr, cc): float4 cc of
      // row rr; even cc from the hi image at the row's slot, odd cc from the lo image at row rr ^ 4
#if TC3_ST256
      {
        // 256-bit stores: lane (rgrp, rsel, j) takes columns 8 j .. 8 j + 7 of row 8 k + rgrp + 4 rsel -- the hi-image chunk of
        // the row and the lo-image chunk of row ^ 4 (rows r and r + 4 share a quarter warp: their swizzled slots are disjoint,
        // so the shared-memory reads stay conflict-free); four store instructions of 8 rows x 128 B instead of eight of 4 rows
        const int j = lane & 3, rl0 = (lane >> 3) + 4 * ((lane >> 2) & 1);
        const unsigned char* a_hi = h_hi + sw128(quad * 32 + rl0, 4 * half + j);
        const unsigned char* a_lo = h_lo + sw128(quad * 32 + (rl0 ^ 4), 4 * half + j);
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[2 * k] = *reinterpret_cast<const float4*>(a_hi + k * 1024);
          v[2 * k + 1] = *reinterpret_cast<const float4*>(a_lo + k * 1024);
        }
        float* op = out0 + (size_t)rl0 * ldh + 8 * j;
        const size_t step = (size_t)ldh * 8;
        const uint32_t vm = vmask >> rl0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.u32 p, %0, 0;\n\t"
              "@p st.global.v8.f32 [%1], {%2, %3, %4, %5, %6, %7, %8, %9};\n\t}"
              ::"r"((vm >> (8 * k)) & 1u), "l"(op), "f"(v[2 * k].x), "f"(v[2 * k].y), "f"(v[2 * k].z), "f"(v[2 * k].w),
                "f"(v[2 * k + 1].x), "f"(v[2 * k + 1].y), "f"(v[2 * k + 1].z), "f"(v[2 * k + 1].w)
              : "memory");
          op += step;
        }
      }
#else
      {
        // all eight shared-memory reads first, then eight predicated stores off one running pointer (no branches, no
        // 64-bit multiply per row)
        const int cc = lane & 7, wh = cc & 1;
        // row 4 k + (lane >> 3) of the quadrant: even k and odd k differ in bit 2 of the row (and of the swizzled slot)
        const unsigned char* img = wh ? h_lo : h_hi;
        const unsigned char* a0 = img + sw128(quad * 32 + (lane >> 3) + 4 * wh, 4 * half + (cc >> 1));
        const unsigned char* a1 = img + sw128(quad * 32 + (lane >> 3) + 4 * (wh ^ 1), 4 * half + (cc >> 1));
        float4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const float4*>(((k & 1) ? a1 : a0) + (k >> 1) * 1024);
        float* op = out0 + (size_t)(lane >> 3) * ldh + 4 * cc;
        const size_t step = (size_t)ldh * 4;
        const uint32_t vm = vmask >> (lane >> 3);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          asm volatile(
              "{\n\t.reg .pred p;\n\t"
              "setp.ne.u32 p, %0, 0;\n\t"
              "@p st.global.v4.f32 [%1], {%2, %3, %4, %5};\n\t}"
              ::"r"((vm >> (4 * k)) & 1u), "l"(op), "f"(v[k].x), "f"(v[k].y), "f"(v[k].z), "f"(v[k].w)
              : "memory");
          op += step;
        }
      }
#endif
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfree + 8 * hb);  // h images read back: the producers may write the next own rows
      TC3_TRACE(it, 12, tr);
      // head: the two column halves of a row live in warps (quad, 0) and (quad, 1) of the team
#if TC3_HEAD_EARLY
      if (half == 0) {
        mbar_wait(bar_dfull, phase, status);
        if (valid) {
          const float lg = dot + dot_part[r] + (first_group ? headb : logit[row_cur]);
          logit[row_cur] = lg;
          if (last_group) score[row_cur] = tmpnn_sigmoid(lg);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dfree);
      }
#else
      if (half == 1) dot_part[r] = dot;
      named_bar_sync(bar_id, 64);
      if (half == 0 && valid) {
        const float lg = dot + dot_part[r] + (first_group ? headb : logit[row_cur]);
        logit[row_cur] = lg;
        if (last_group) score[row_cur] = tmpnn_sigmoid(lg);
      }
      named_bar_sync(bar_id, 64);
#endif
      TC3_TRACE(it, 13, tr);
      T0 = T1; T1 = T2; srcv = srcv1;
      hb = hb == 0 ? 2 : hb - 1;  // (hb + 2) % 3
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

int tmpnn_init_tc3() {
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_edge_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return TMPNN_OK;
}

extern "C" size_t tmpnn_tc_tile_table_bytes(int num_seqs, int cap_rows) {
  return (size_t)num_seqs * (size_t)tmpnn_div_up(cap_rows, TCM) * sizeof(int4) + sizeof(int4);
}

// Tensor map of the endpoint images for the TMA gathers: det_img as a 2-D fp16 tensor [S cap_rows rows] x [2 ldh columns]
// (a row = per feature group 64 hi halves then 64 lo halves), box = 64 columns x 1 row (tile::gather4 takes four rows per
// instruction), 128-byte swizzle = the UMMA layout of the x images.  Host-side encode (driver entry point fetched once).
static int make_image_map(const float* det_img, long long rows, int ldh, CUtensorMap* tm) {
#if TC3_TMA
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn enc = nullptr;
  if (!enc) {
    cudaDriverEntryPointQueryResult qr;
    void* fn = nullptr;
    TMPNN_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn || qr != cudaDriverEntryPointSuccess) return tmpnn_set_error(TMPNN_E_CUDA, "cuTensorMapEncodeTiled is not available");
    enc = (EncodeFn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)ldh * 2, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ldh * 4};
  const cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
  const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)det_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return tmpnn_set_error(TMPNN_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
#else
  (void)det_img; (void)rows; (void)ldh;
  memset(tm, 0, sizeof(*tm));
#endif
  return TMPNN_OK;
}

// ---- detection rows on the same kernel (tmpnn_mp_det_fwd_tc) ---------------------------------------------------------------
// Tiles over the detection segments of every slab (contiguous rows: the kernel writes its output at consecutive logical rows).
// tab[0].x = number of tiles, tiles from tab[1] on.
__global__ void __launch_bounds__(1024) k_det_tile_table(const SlabSegs* __restrict__ segs, int num_seqs, int cap_rows, int cap_tiles,
                                                         int4* __restrict__ tab, int32_t* __restrict__ status) {
  __shared__ int sm[33];
  int carry = 0;
  for (int s0 = 0; s0 < num_seqs; s0 += 1024) {
    const int s = s0 + threadIdx.x;
    int cnt = 0;
    if (s < num_seqs) {
      const SlabSegs& sg = segs[s];
      for (int q = 0; q < sg.nseg; ++q)
        if (sg.eord[q] < 0) cnt += (sg.start[q + 1] - sg.start[q] + TCM - 1) / TCM;
    }
    int total;
    int o = carry + block_exclusive_scan(cnt, sm, &total);
    if (s < num_seqs && carry + total <= cap_tiles) {
      const SlabSegs& sg = segs[s];
      for (int q = 0; q < sg.nseg; ++q) {
        if (sg.eord[q] >= 0) continue;
        const int r0 = sg.start[q], len = sg.start[q + 1] - r0;
        for (int j = 0; j * TCM < len; ++j) tab[1 + o++] = make_int4(s * cap_rows, r0 + j * TCM, len - j * TCM, 0);
      }
    }
    carry += total;
  }
  if (threadIdx.x == 0) {
    if (carry > cap_tiles) { atomicOr(status, TMPNN_FLAG_DET_CAPACITY); carry = 0; }
    tab[0] = make_int4(carry, 0, 0, 0);
  }
}
// Half-warp per detection: the fp16 hi / lo image of its aggregate at the detection's logical row of det_img (dead since the
// edge step finished with the far-endpoint images); thread 0 of the grid writes the bias-only P' row the kernel adds.
__global__ void __launch_bounds__(256)
k_agg_image(const float* __restrict__ agg, const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows,
            const unsigned char* __restrict__ image, float* __restrict__ det_img, int ldh, int col, float* __restrict__ det_p,
            int32_t* __restrict__ status) {
  if (blockIdx.x == 0 && threadIdx.x < 192) {
    const float b = reinterpret_cast<const float*>(image + OFF_BS)[threadIdx.x];
    const float wscale = *reinterpret_cast<const float*>(image + OFF_HEADB + 8);
    det_p[threadIdx.x] = threadIdx.x < 2 * H ? -LOG2E * b : wscale * b;   // k_det_prepare's P' of an all-zero source
  }
  const int k = (blockIdx.x * 256 + threadIdx.x) >> 4, l = threadIdx.x & 15;
  if (k >= *n_dets) return;
  const float4 v = ldg4(agg + (size_t)k * H + 4 * l);
  uint2 hi, lo;
  float amax = 0.f;
  split4(v, hi, lo, amax);
  unsigned char* ib = reinterpret_cast<unsigned char*>(det_img + (size_t)det_rows[k] * ldh + col);
  *reinterpret_cast<uint2*>(ib + 8 * l) = hi;
  *reinterpret_cast<uint2*>(ib + 128 + 8 * l) = lo;
  if (amax > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);
}

extern "C" size_t tmpnn_tc_det_tile_table_bytes(int num_seqs, int cap_dets) {
  return ((size_t)num_seqs * MAXSEG + (size_t)tmpnn_div_up(cap_dets, TCM) + 2) * sizeof(int4);
}

extern "C" int tmpnn_mp_det_fwd_tc(const tmpnn_graph* g, const tmpnn_index* ix, const void* index_scratch2, const float* h_in,
                                   float* h_out, int ldh, int group, int num_groups, const void* node_image, const float* agg,
                                   float* det_img, float* det_p, void* det_tile_table, void* stream) {
  TMPNN_REQUIRE(g && ix && index_scratch2 && h_in && h_out && node_image && agg && det_img && det_p && det_tile_table, "null argument");
  TMPNN_REQUIRE(h_in != h_out && det_img != h_in && det_img != h_out, "h_in, h_out and det_img must be distinct buffers");
  TMPNN_REQUIRE(ldh % 4 == 0 && group >= 0 && group < num_groups && ldh >= num_groups * H, "bad ldh / group");
  TMPNN_REQUIRE(((uintptr_t)det_tile_table & 15) == 0, "det_tile_table must be 16-byte aligned");
  int rc = tmpnn_init();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int4* tab = (int4*)det_tile_table;
  if (group == 0) {
    const int cap_tiles = (int)(tmpnn_tc_det_tile_table_bytes(g->num_seqs, ix->cap_dets) / sizeof(int4)) - 2;
    k_det_tile_table<<<1, 1024, 0, st>>>((const SlabSegs*)index_scratch2, g->num_seqs, g->cap_rows, cap_tiles, tab, g->status);
    TMPNN_LAUNCH_CHECK();
  }
  k_agg_image<<<tmpnn_div_up(ix->cap_dets, 16), 256, 0, st>>>(agg, ix->n_dets, ix->det_rows, (const unsigned char*)node_image, det_img,
                                                            ldh, group * H, det_p, g->status);
  TMPNN_LAUNCH_CHECK();
  TMPNN_CUDA_TRY(cudaMemcpyToSymbolAsync(c_tc3_const, (const unsigned char*)node_image + OFF_BIAS + 3 * H * 4, 528, 0, cudaMemcpyDeviceToDevice, st));
  CUtensorMap tmx;
  rc = make_image_map(det_img, (long long)g->num_seqs * g->cap_rows, ldh, &tmx);
  if (rc) return rc;
  // x = the aggregate enters with its own sign (no negation), bit 0: detection mode
  k_mp_edge_tc3<<<TMPNN_SM_COUNT, TC3_THREADS, SMEM_BYTES, st>>>(
      h_in, h_out, ldh, group * H, g->src, g->dst, reinterpret_cast<const int32_t*>(tab), tab + 1, (const unsigned char*)node_image,
      g->logit, g->score, group == 0, group == num_groups - 1, g->status, g->phys, det_img, det_p, ix->det_of_row, 1u, tmx);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

int tmpnn_edge_tc3_launch(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh, int group,
                          int num_groups, int concat, const void* edge_image, const float* det_img, const float* det_p,
                          void* tile_table, cudaStream_t st) {
  if (group == 0) {
    dim3 grid(max(1, min(tmpnn_div_up(tmpnn_div_up(g->cap_rows, TCM), 256), 8)), g->num_seqs);
    k_tile_table<<<grid, 256, 0, st>>>(g->n_rows, ix->tile128_ptr, g->cap_rows, (int4*)tile_table);
    TMPNN_LAUNCH_CHECK();
  }
  TMPNN_CUDA_TRY(cudaMemcpyToSymbolAsync(c_tc3_const, (const unsigned char*)edge_image + OFF_BIAS + 3 * H * 4, 528, 0, cudaMemcpyDeviceToDevice, st));
  CUtensorMap tmx;
  { const int rc = make_image_map(det_img, (long long)g->num_seqs * g->cap_rows, ldh, &tmx); if (rc) return rc; }
  // 'diff': x = h[src] - h[dst]  ->  the far endpoint enters negated (instruction descriptor bit 13: negate A)
  k_mp_edge_tc3<<<TMPNN_SM_COUNT, TC3_THREADS, SMEM_BYTES, st>>>(
      h_in, h_out, ldh, group * H, g->src, g->dst, ix->tile128_ptr + g->num_seqs, (const int4*)tile_table,
      (const unsigned char*)edge_image, g->logit, g->score, group == 0, group == num_groups - 1, g->status, g->phys, det_img,
      det_p, ix->det_of_row, concat ? 0u : (1u << 13), tmx);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
