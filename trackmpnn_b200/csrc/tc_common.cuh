// tc_common.cuh -- shared pieces of the tcgen05 edge-step kernels (mp_step_tc.cu):
// shared-memory layout, PTX wrappers (mbarrier, tcgen05.mma / ld / commit, descriptors), the fp16 hi/lo
// split and the MMA issue sequence of one 128-row tile.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int H = TMPNN_HIDDEN;
constexpr int TCM = 128;            // rows per tile == UMMA M
constexpr int B_BYTES = 192 * 128;  // one [192 x 64] fp16 weight image
constexpr int OFF_BX_HI = 0, OFF_BX_LO = B_BYTES, OFF_BH_HI = 2 * B_BYTES, OFF_BH_LO = 3 * B_BYTES;
// row order of the images: W_ih r | z | n, W_hh n | r | z (so that one N = 192 instruction covers h_n | r | z, see mp_step_tc3.cu)
constexpr int OFF_BIAS = 4 * B_BYTES;            // 4 x 64 floats: -log2e (b_ir+b_hr), -log2e (b_iz+b_hz), b_in, b_hn
constexpr int OFF_HEADW = OFF_BIAS + 1024;       // 64 floats
constexpr int OFF_HEADB = OFF_HEADW + 256;       // 1 float (+ pad)
constexpr int IMAGE_BYTES = OFF_HEADB + 16;      // the part of the packed image the step kernels keep in shared memory
// behind it, for k_det_prepare: the source-side W_ih^T as fp32 [64][192] and the folded bias [192]
constexpr int OFF_WT = IMAGE_BYTES;
constexpr int OFF_BS = OFF_WT + 64 * 192 * 4;
constexpr int IMAGE_TOTAL_BYTES = OFF_BS + 192 * 4;
constexpr int OFF_BAR = IMAGE_BYTES;             // up to 16 mbarriers + tmem pointer, inside the alignment gap
constexpr int OFF_DOT = OFF_BAR + 144;           // 128 floats: head partial sums of the upper column half
constexpr int OFF_A = 98 * 1024;                 // first A stage (1024-aligned)
constexpr int A_PART = TCM * 128;                // [128 rows x 64 fp16] = 16 KB
constexpr int A_STAGE = 4 * A_PART;              // x_hi, x_lo, h_hi, h_lo
constexpr int SMEM_BYTES = OFF_A + 2 * A_STAGE + 1024;  // + slack to 1024-align the base
static_assert(OFF_DOT + 512 <= OFF_A, "barriers and head partials must fit in the gap");
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");
static_assert(OFF_HEADW == OFF_BIAS + 1024 && OFF_HEADB == OFF_HEADW + 256, "b_hn | head weights | head constants must be adjacent (one constant copy)");
constexpr float LOG2E = 1.4426950408889634f;

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// TMPNN_TC_WAIT_HINT_NS > 0: suspend-time hint of the hardware wait (the thread sleeps until the phase completes or the hint
// expires, instead of the short system default): fewer spin iterations of idle warps without delaying their wake-up.
// Measured with 2 us and 20 us (profiles/r02_ab_tc3_waithint.txt): no difference -- the spinning warps do not cost the epilogue
// issue slots
#ifndef TMPNN_TC_WAIT_HINT_NS
#define TMPNN_TC_WAIT_HINT_NS 0
#endif
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#if TMPNN_TC_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"((uint32_t)TMPNN_TC_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
#endif
  return ok;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must never hang the GPU -- after 2 s the kernel flags an error
// and runs to completion with garbage instead.  mbarrier.try_wait is itself a suspending wait (the
// thread sleeps in hardware until the phase completes or a time limit expires), so the loop simply
// re-issues it; TMPNN_TC_SLEEP_NS > 0 adds a nanosleep back-off between polls.
#ifndef TMPNN_TC_SLEEP_NS
#define TMPNN_TC_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* status) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spin = 0;
  unsigned long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (TMPNN_TC_SLEEP_NS > 0) __nanosleep(TMPNN_TC_SLEEP_NS);
    if ((++spin & 1023u) == 0) {
      const unsigned long long t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ull) {
        atomicOr(status, TMPNN_FLAG_TC_TIMEOUT);
        return;
      }
    }
  }
}
// one elected lane of a converged warp (the compiler knows a region guarded by elect.sync has a single active thread, so
// descriptors computed inside it go to uniform registers without a per-lane broadcast loop)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset       bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major, M = 128
__device__ __forceinline__ constexpr uint32_t umma_idesc(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of 16-byte chunk c (8 fp16 along K) of row r inside a [rows x 128 B] swizzled image
__device__ __forceinline__ uint32_t sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// 4 floats -> 4 fp16 "hi" (round to nearest) + 4 fp16 "lo" (the residual)
__device__ __forceinline__ void split4(const float4 a, uint2& hi, uint2& lo, float& amax) {
  amax = fmaxf(amax, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
  const __half2 h01 = __floats2half2_rn(a.x, a.y), h23 = __floats2half2_rn(a.z, a.w);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(a.x - f01.x, a.y - f01.y), l23 = __floats2half2_rn(a.z - f23.x, a.w - f23.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2): halves the issue slots of the gate math
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t a, uint32_t b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void up2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 32 bytes (four packed pairs) in one st.global.v4.b64: sm_100's 256-bit store (STG.E.256), address 32-byte aligned
__device__ __forceinline__ void st_global_256(float* p, const f32x2* v) {
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
}
// element-wise min(v, c) of a packed pair
__device__ __forceinline__ f32x2 min2c(f32x2 v, float c) {
  float a, b;
  up2(v, a, b);
  return pk2(fminf(a, c), fminf(b, c));
}
__device__ __forceinline__ f32x2 ex2_2(f32x2 v) {
  float a, b;
  up2(v, a, b);
  return pk2(ex2_approx(a), ex2_approx(b));
}
__device__ __forceinline__ f32x2 rcp_2(f32x2 v) {
  float a, b;
  up2(v, a, b);
  return pk2(rcp_approx(a), rcp_approx(b));
}
__device__ __forceinline__ void tmem_ld8u(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4u(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// first sequence whose tile range contains `tile`, advancing a cursor (tiles are visited in
// increasing order, so this is 0-1 steps per call after the first)
__device__ __forceinline__ void seek_seq(const int32_t* __restrict__ tile_ptr, int num_seqs, int tile, int& seq) {
  while (seq + 1 < num_seqs && __ldg(tile_ptr + seq + 1) <= tile) ++seq;
}

// ---- the two halves of the MMA issue, far-endpoint part FIRST (mp_step_tc3.cu) ---------------------------
// The x images are complete long before the tile's h images may be written (those double as the epilogue's
// transpose buffer), so the far-endpoint MMAs initialise r | z | i_n as soon as the accumulators are drained;
// only the own-row half waits for the h images.  x_u / h_u: shared-memory addresses of [x_hi | x_lo] and
// [h_hi | h_lo]; d0: first TMEM column of the accumulator stage.
__device__ __forceinline__ void issue_tile_mma_x_first(uint32_t sm_u, uint32_t d0, uint32_t x_u, uint32_t xflags) {
  const uint32_t ax[3] = {x_u, x_u + A_PART, x_u};
  const uint32_t bx[3] = {sm_u + OFF_BX_HI, sm_u + OFF_BX_HI, sm_u + OFF_BX_LO};
  uint32_t acc = 0;
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      umma_f16(d0 + 64, umma_desc(ax[t] + 32 * j), umma_desc(bx[t] + 32 * j), umma_idesc(192) | xflags, acc);
      acc = 1;
    }
}
// own-row part as ONE N = 192 instruction per K step: the W_hh image holds its rows in the order n | r | z (k_pack_gru_tc),
// the accumulator stage is laid out h_n | r | z | i_n, so [0,192) takes the own-row product and [64,256) the far-endpoint
// product.  Always accumulates: r | z continue the far-endpoint sums, the h_n columns were zeroed by the epilogue when it
// drained the stage.  12 instructions and 120 KB of operand reads per tile instead of 24 and 168 KB.
__device__ __forceinline__ void issue_tile_mma_h_second(uint32_t sm_u, uint32_t d0, uint32_t h_u) {
  const uint32_t ah[3] = {h_u, h_u + A_PART, h_u};
  const uint32_t bh[3] = {sm_u + OFF_BH_HI, sm_u + OFF_BH_HI, sm_u + OFF_BH_LO};
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) umma_f16(d0, umma_desc(ah[t] + 32 * j), umma_desc(bh[t] + 32 * j), umma_idesc(192), 1);
}
// the same two halves for ONE hidden group (mp_step_tc3.cu, TC3_SPLIT): the kernel keeps its weight images group-major --
// rows [96 g, 96 g + 96) = the three gates of hidden group g -- and the accumulator stage as two 128-column groups
// h_n | r | z | i_n of 32 hidden units each, so every instruction has N = 96: the far-endpoint product lands in columns
// [32, 128) of the group, the own-row product in [0, 96).  dg: first TMEM column of the group.
__device__ __forceinline__ void issue_group_mma_x(uint32_t sm_u, uint32_t dg, uint32_t x_u, uint32_t xflags, int g) {
  const uint32_t ax[3] = {x_u, x_u + A_PART, x_u};
  const uint32_t bx[3] = {sm_u + OFF_BX_HI + 96 * 128 * g, sm_u + OFF_BX_HI + 96 * 128 * g, sm_u + OFF_BX_LO + 96 * 128 * g};
  uint32_t acc = 0;
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      umma_f16(dg + 32, umma_desc(ax[t] + 32 * j), umma_desc(bx[t] + 32 * j), umma_idesc(96) | xflags, acc);
      acc = 1;
    }
}
__device__ __forceinline__ void issue_group_mma_h(uint32_t sm_u, uint32_t dg, uint32_t h_u, int g) {
  const uint32_t ah[3] = {h_u, h_u + A_PART, h_u};
  const uint32_t bh[3] = {sm_u + OFF_BH_HI + 96 * 128 * g, sm_u + OFF_BH_HI + 96 * 128 * g, sm_u + OFF_BH_LO + 96 * 128 * g};
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) umma_f16(dg, umma_desc(ah[t] + 32 * j), umma_desc(bh[t] + 32 * j), umma_idesc(96), 1);
}
// 16 zero columns for this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
      "r"(0u)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 32 zero columns for this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- MMA issue for one 128-row tile (one thread) ------------------------------------------------
__device__ __forceinline__ void issue_tile_mma(uint32_t sm_u, uint32_t tmem_base, int stage, uint32_t xflags,
                                               uint32_t bar_xfree) {
  const uint32_t a_u = sm_u + OFF_A + stage * A_STAGE;
  const uint32_t d0 = tmem_base + (uint32_t)(stage * 256);
  // x part: columns [0,192) = r | z | i_n
  const uint32_t ax[3] = {a_u, a_u + A_PART, a_u};
  const uint32_t bx[3] = {sm_u + OFF_BX_HI, sm_u + OFF_BX_HI, sm_u + OFF_BX_LO};
  uint32_t acc = 0;
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      umma_f16(d0, umma_desc(ax[t] + 32 * j), umma_desc(bx[t] + 32 * j), umma_idesc(192) | xflags, acc);
      acc = 1;
    }
  if (bar_xfree) umma_commit(bar_xfree);  // the x images are reusable once these MMAs retire (0: released by the epilogue)
  // h part: columns [0,128) += r | z, columns [192,256) = h_n
  const uint32_t ah[3] = {a_u + 2 * A_PART, a_u + 3 * A_PART, a_u + 2 * A_PART};
  const uint32_t bh[3] = {sm_u + OFF_BH_HI, sm_u + OFF_BH_HI, sm_u + OFF_BH_LO};
  uint32_t acc_n = 0;
#pragma unroll
  for (int t = 0; t < 3; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      umma_f16(d0, umma_desc(ah[t] + 32 * j), umma_desc(bh[t] + 64 * 128 + 32 * j), umma_idesc(128), 1);  // rows r | z
      umma_f16(d0 + 192, umma_desc(ah[t] + 32 * j), umma_desc(bh[t] + 32 * j), umma_idesc(64), acc_n);        // rows n
      acc_n = 1;
    }
}


}  // namespace
