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
// One persistent CTA per SM, 288 threads, warp specialised:
//   warps 0-3  epilogue: TMEM -> registers (tcgen05.ld 32x32b, thread == row), gates, h', head
//   warps 4-7  producers: gather h[src], h[dst], h[row] (128-bit loads), subtract, split to fp16
//              hi/lo, store into the 128B-swizzled K-major UMMA layout
//   warp  8    TMEM allocation + single-thread MMA issue (36 tcgen05.mma per 128-row tile)
// Shared memory (227 KB): packed weights as fp16 hi/lo UMMA images (96 KB, resident for the whole
// kernel) + two A stages of [128 x (64 x | 64 h)] fp16 hi/lo (2 x 64 KB).  TMEM: two accumulator
// stages of 256 columns (r | z | i_n | h_n).  mbarriers: full[2] (producers -> MMA),
// done[2] (tcgen05.commit -> epilogue), free[2] (epilogue -> producers / MMA).
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int H = TMPNN_HIDDEN;
constexpr int TCM = 128;            // rows per tile == UMMA M
constexpr int TC_THREADS = 288;
constexpr int B_BYTES = 192 * 128;  // one [192 x 64] fp16 weight image
constexpr int OFF_BX_HI = 0, OFF_BX_LO = B_BYTES, OFF_BH_HI = 2 * B_BYTES, OFF_BH_LO = 3 * B_BYTES;
constexpr int OFF_BIAS = 4 * B_BYTES;            // 4 x 64 floats
constexpr int OFF_HEADW = OFF_BIAS + 1024;       // 64 floats
constexpr int OFF_HEADB = OFF_HEADW + 256;       // 1 float (+ pad)
constexpr int IMAGE_BYTES = OFF_HEADB + 16;      // what tmpnn_pack_gru_tc writes
constexpr int OFF_BAR = IMAGE_BYTES;             // 6 mbarriers + tmem pointer, inside the alignment gap
constexpr int OFF_A = 98 * 1024;                 // first A stage (1024-aligned)
constexpr int A_PART = TCM * 128;                // [128 rows x 64 fp16] = 16 KB
constexpr int A_STAGE = 4 * A_PART;              // x_hi, x_lo, h_hi, h_lo
constexpr int SMEM_BYTES = OFF_A + 2 * A_STAGE + 1024;  // + slack to 1024-align the base
static_assert(OFF_BAR + 64 <= OFF_A, "barriers must fit in the gap");
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB");

// ---- PTX helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must never hang the GPU -- after 2 s the kernel flags an error
// and runs to completion with garbage instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* status) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = globaltimer_ns();
  uint32_t spin = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spin & 255u) == 0 && globaltimer_ns() - t0 > 2000000000ull) {
      atomicOr(status, TMPNN_FLAG_TC_TIMEOUT);
      return;
    }
  }
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

__device__ __forceinline__ void split8(const float* a, uint4& hi, uint4& lo, float& amax) {
  __half2 h2[4], l2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a0 = a[2 * i], a1 = a[2 * i + 1];
    amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
    const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
    h2[i] = __halves2half2(h0, h1);
    l2[i] = __halves2half2(__float2half_rn(a0 - __half2float(h0)), __float2half_rn(a1 - __half2float(h1)));
  }
  hi = *reinterpret_cast<uint4*>(h2);
  lo = *reinterpret_cast<uint4*>(l2);
}

// sigmoid / tanh from ex2.approx + rcp.approx: ~1e-6 relative, far inside the 1e-4 tolerance
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// ---- weight image ----------------------------------------------------------------------------
__global__ void k_pack_gru_tc(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                              const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                              const float* __restrict__ head_w, const float* __restrict__ head_b,
                              unsigned char* __restrict__ img) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < 2 * 192 * 64; i += nth) {  // element (matrix m, row n, col k)
    const int m = i / (192 * 64), n = (i / 64) % 192, k = i % 64;
    const float w = m == 0 ? w_ih[n * 64 + k] : w_hh[n * 64 + k];
    const __half hi = __float2half_rn(w);
    const __half lo = __float2half_rn(w - __half2float(hi));
    const uint32_t off = sw128(n, k >> 3) + (k & 7) * 2;
    *reinterpret_cast<__half*>(img + (m == 0 ? OFF_BX_HI : OFF_BH_HI) + off) = hi;
    *reinterpret_cast<__half*>(img + (m == 0 ? OFF_BX_LO : OFF_BH_LO) + off) = lo;
  }
  float* bias = reinterpret_cast<float*>(img + OFF_BIAS);
  for (int i = tid; i < 4 * H; i += nth) {
    const int g = i / H, j = i % H;
    bias[i] = g < 2 ? b_ih[g * H + j] + b_hh[g * H + j] : g == 2 ? b_ih[2 * H + j] : b_hh[2 * H + j];
  }
  float* hw = reinterpret_cast<float*>(img + OFF_HEADW);
  for (int i = tid; i < H; i += nth) hw[i] = head_w[i];
  if (tid == 0) {
    float* hb = reinterpret_cast<float*>(img + OFF_HEADB);
    hb[0] = head_b[0]; hb[1] = hb[2] = hb[3] = 0.f;
  }
}

// ---- the kernel ------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
k_mp_edge_tc(const float* __restrict__ h_in, float* __restrict__ h_out, int ldh, int col,
             const int32_t* __restrict__ n_rows, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
             int cap_rows, int num_seqs, const int32_t* __restrict__ tile_ptr, const unsigned char* __restrict__ image,
             float* __restrict__ logit, float* __restrict__ score, int first_group, int last_group,
             int32_t* __restrict__ status) {
  extern __shared__ unsigned char smem_dyn[];
  const int total = tile_ptr[num_seqs];
  if ((int)blockIdx.x >= total) return;  // uniform: whole CTA leaves before touching TMEM / barriers
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = sm_u + OFF_BAR, bar_done = bar_full + 16, bar_free = bar_full + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + OFF_BAR + 48);

  // resident weight image (generic-proxy stores, made visible to the async proxy below)
  {
    const uint4* gsrc = reinterpret_cast<const uint4*>(image);
    uint4* sdst = reinterpret_cast<uint4*>(sm);
    for (int i = threadIdx.x; i < IMAGE_BYTES / 16; i += TC_THREADS) sdst[i] = __ldg(gsrc + i);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_full + 8 * s, 4);  // one arrive per producer warp
      mbar_init(bar_done + 8 * s, 1);  // tcgen05.commit
      mbar_init(bar_free + 8 * s, 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const float* bias = reinterpret_cast<const float*>(sm + OFF_BIAS);
  const float* headw = reinterpret_cast<const float*>(sm + OFF_HEADW);
  const float headb = *reinterpret_cast<const float*>(sm + OFF_HEADB);

  int it = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    const uint32_t phase = (uint32_t)(it >> 1) & 1u;
    unsigned char* a_stage = sm + OFF_A + stage * A_STAGE;

    if (warp == 8) {
      // ================= MMA issuer =================
      if (lane == 0) {
        mbar_wait(bar_free + 8 * stage, phase ^ 1u, status);  // accumulator stage drained
        mbar_wait(bar_full + 8 * stage, phase, status);       // operands landed
        tc_fence_after();
        const uint32_t a_u = smem_u32(a_stage);
        const uint32_t d0 = tmem_base + (uint32_t)(stage * 256);
        // x part: columns [0,192) = r | z | i_n
        const uint32_t ax[3] = {a_u, a_u + A_PART, a_u};
        const uint32_t bx[3] = {sm_u + OFF_BX_HI, sm_u + OFF_BX_HI, sm_u + OFF_BX_LO};
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_f16(d0, umma_desc(ax[t] + 32 * j), umma_desc(bx[t] + 32 * j), umma_idesc(192), acc);
            acc = 1;
          }
        // h part: columns [0,128) += r | z, columns [192,256) = h_n
        const uint32_t ah[3] = {a_u + 2 * A_PART, a_u + 3 * A_PART, a_u + 2 * A_PART};
        const uint32_t bh[3] = {sm_u + OFF_BH_HI, sm_u + OFF_BH_HI, sm_u + OFF_BH_LO};
        uint32_t acc_n = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            umma_f16(d0, umma_desc(ah[t] + 32 * j), umma_desc(bh[t] + 32 * j), umma_idesc(128), 1);
            umma_f16(d0 + 192, umma_desc(ah[t] + 32 * j), umma_desc(bh[t] + 128 * 128 + 32 * j), umma_idesc(64), acc_n);
            acc_n = 1;
          }
        umma_commit(bar_done + 8 * stage);  // implies tcgen05.fence::before_thread_sync
      }
      __syncwarp();
      continue;
    }

    // tile -> (sequence, first row); tile_ptr counts 128-row tiles per sequence
    int lo = 0, hi = num_seqs;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tile_ptr[mid] <= tile) lo = mid; else hi = mid;
    }
    const int seq = lo;
    const int r0 = (tile - tile_ptr[seq]) * TCM;
    const int n = n_rows[seq];
    const size_t base = (size_t)seq * cap_rows;

    if (warp >= 4) {
      // ================= producers =================
      mbar_wait(bar_free + 8 * stage, phase ^ 1u, status);  // epilogue is done with this stage (h_prev lives here)
      const int pt = threadIdx.x - 128, q = pt & 7, rsub = pt >> 3;
      float amax = 0.f;
#pragma unroll 2
      for (int p = 0; p < TCM / 16; ++p) {
        const int r = rsub + 16 * p;
        const int lr = r0 + r;
        int a = -1, b = -1;
        if (lr < n) { a = src[base + lr]; b = dst[base + lr]; }
        float x[8], hp[8];
        if (a >= 0) {
          const float* ps = h_in + (base + a) * ldh + col + 8 * q;
          const float* pd = h_in + (base + b) * ldh + col + 8 * q;
          const float* pr = h_in + (base + lr) * ldh + col + 8 * q;
          const float4 s0 = ldg4(ps), s1 = ldg4(ps + 4), d0 = ldg4(pd), d1 = ldg4(pd + 4);
          const float4 h0 = ldg4(pr), h1 = ldg4(pr + 4);
          x[0] = s0.x - d0.x; x[1] = s0.y - d0.y; x[2] = s0.z - d0.z; x[3] = s0.w - d0.w;
          x[4] = s1.x - d1.x; x[5] = s1.y - d1.y; x[6] = s1.z - d1.z; x[7] = s1.w - d1.w;
          hp[0] = h0.x; hp[1] = h0.y; hp[2] = h0.z; hp[3] = h0.w;
          hp[4] = h1.x; hp[5] = h1.y; hp[6] = h1.z; hp[7] = h1.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) { x[i] = 0.f; hp[i] = 0.f; }
        }
        uint4 xh, xl, hh, hl;
        split8(x, xh, xl, amax);
        split8(hp, hh, hl, amax);
        const uint32_t off = sw128(r, q);
        *reinterpret_cast<uint4*>(a_stage + off) = xh;
        *reinterpret_cast<uint4*>(a_stage + A_PART + off) = xl;
        *reinterpret_cast<uint4*>(a_stage + 2 * A_PART + off) = hh;
        *reinterpret_cast<uint4*>(a_stage + 3 * A_PART + off) = hl;
      }
      if (amax > 60000.f) atomicOr(status, TMPNN_FLAG_TC_RANGE);  // fp16 split would overflow: use the FMA path
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * stage);
    } else {
      // ================= epilogue =================
      const int r = threadIdx.x;  // row of the tile == TMEM lane
      const int lr = r0 + r;
      const bool valid = lr < n && src[base + lr] >= 0;
      const size_t row = base + lr;
      mbar_wait(bar_done + 8 * stage, phase, status);
      tc_fence_after();
      const uint32_t t0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(stage * 256);
      float dot = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        float ar[16], az[16], an[16], ahn[16];
        tmem_ld16(t0 + ch * 16, ar);
        tmem_ld16(t0 + 64 + ch * 16, az);
        tmem_ld16(t0 + 128 + ch * 16, an);
        tmem_ld16(t0 + 192 + ch * 16, ahn);
        // h_prev = h_hi + h_lo from the stage's h images
        float hp[16];
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const uint32_t off = sw128(r, 2 * ch + c2);
          const uint4 vh = *reinterpret_cast<const uint4*>(a_stage + 2 * A_PART + off);
          const uint4 vl = *reinterpret_cast<const uint4*>(a_stage + 3 * A_PART + off);
          const __half2* ph = reinterpret_cast<const __half2*>(&vh);
          const __half2* pl = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 fh = __half22float2(ph[i]), fl = __half22float2(pl[i]);
            hp[8 * c2 + 2 * i] = fh.x + fl.x;
            hp[8 * c2 + 2 * i + 1] = fh.y + fl.y;
          }
        }
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = ch * 16 + i;
          const float rg = fast_sigmoid(ar[i] + bias[j]);
          const float zg = fast_sigmoid(az[i] + bias[H + j]);
          const float ng = fast_tanh(an[i] + bias[2 * H + j] + rg * (ahn[i] + bias[3 * H + j]));
          o[i] = (1.0f - zg) * ng + zg * hp[i];
          dot = fmaf(o[i], headw[j], dot);
        }
        if (valid) {
          float4* po = reinterpret_cast<float4*>(h_out + row * ldh + col + ch * 16);
          po[0] = make_float4(o[0], o[1], o[2], o[3]);
          po[1] = make_float4(o[4], o[5], o[6], o[7]);
          po[2] = make_float4(o[8], o[9], o[10], o[11]);
          po[3] = make_float4(o[12], o[13], o[14], o[15]);
        }
      }
      if (valid) {
        const float lg = dot + (first_group ? headb : logit[row]);
        logit[row] = lg;
        if (last_group) score[row] = tmpnn_sigmoid(lg);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_free + 8 * stage);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

int tmpnn_init_tc() {
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_edge_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  return TMPNN_OK;
}

extern "C" size_t tmpnn_gru_tc_pack_bytes(void) { return (size_t)IMAGE_BYTES; }

extern "C" int tmpnn_pack_gru_tc(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                 const float* head_w, const float* head_b, void* packed, void* stream) {
  TMPNN_REQUIRE(w_ih && w_hh && b_ih && b_hh && head_w && head_b && packed, "null argument");
  TMPNN_REQUIRE(((uintptr_t)packed & 15) == 0, "packed image must be 16-byte aligned");
  k_pack_gru_tc<<<48, 256, 0, (cudaStream_t)stream>>>(w_ih, w_hh, b_ih, b_hh, head_w, head_b, (unsigned char*)packed);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_mp_edge_fwd_tc(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                    int group, int num_groups, const void* edge_image, void* stream) {
  TMPNN_REQUIRE(g && ix && h_in && h_out && edge_image && ix->tile128_ptr, "null argument");
  TMPNN_REQUIRE(h_in != h_out, "h_in and h_out must be distinct buffers (Jacobi update)");
  TMPNN_REQUIRE(ldh % 4 == 0 && group >= 0 && group < num_groups && ldh >= num_groups * H, "bad ldh / group");
  int rc = tmpnn_init();
  if (rc) return rc;
  k_mp_edge_tc<<<TMPNN_SM_COUNT, TC_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(
      h_in, h_out, ldh, group * H, g->n_rows, g->src, g->dst, g->cap_rows, g->num_seqs, ix->tile128_ptr,
      (const unsigned char*)edge_image, g->logit, g->score, group == 0, group == num_groups - 1, g->status);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
