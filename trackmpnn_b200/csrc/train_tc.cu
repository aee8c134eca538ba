// train_tc.cu -- the two contractions of the step's backward pass on tcgen05 tensor cores, one kernel for both:
//
//   C[rc(i)][0:64]  (+)= A[ra(i)][0:192] . W[192][0:64]                (dx = dgi . W_ih,  dh_self += dgh . W_hh)
//   G[192][0:64]     +=  sum_i A[ra(i)][0:192]^T  X[rx(i)][0:64]        (dW_ih += dgi^T x,  dW_hh += dgh^T h)
//
// i.e. what loss.backward() does for the two torch.nn.GRUCell of models/layers.py:97,114 (reference train.py:127-131),
// over the rows of a (block-diagonal, single-slab) training graph.  They replace k_rows_times_w / k_rows_outer of
// train.cu (fp32 FMA, 44 % of a batched training step) and share ONE pass over A:
//
//   * per 128-row tile the producers load A [128 x 192] and X [128 x 64] (fp32, 16 lanes x float4 per row), split both into
//     bf16 hi / lo pairs (a = hi + lo; bf16 keeps the fp32 exponent range, which gradients need -- fp16 residuals of
//     1e-6-sized values would be subnormal) and store them as [128 rows x 64 columns] images of 128-byte rows with the
//     128-byte swizzle: 3 column blocks of A and 1 of X, hi and lo = 8 images of 16 KB;
//   * the SAME image is both a K-major operand (row = M index, 64 columns = K) for C = A W and an MN-major operand
//     (64 columns = M or N index, row = K) for G = A^T X: the canonical MN-major SWIZZLE_128B atom is 64 elements along
//     M/N (one 128-byte row) by 8 along K (8 consecutive rows, 1024 bytes) -- exactly the K-major tile read the other way
//     round; only the descriptor (leading / stride offsets) and two bits of the instruction descriptor differ.  No transposed
//     copy of A or X is ever made;
//   * one elected thread issues, per tile, 36 tcgen05.mma (M 128, N 64, K 16; terms hi.hi + lo.hi + hi.lo over 3 column
//     blocks x 4 K steps) into a double-buffered C accumulator and 48 (2 M blocks x 8 K steps x 3 terms, A and B MN-major)
//     into the G accumulators, which live in TMEM for the whole kernel (the contraction index is the row: every tile
//     accumulates);
//   * four epilogue warps drain C (tcgen05.ld 32x32b, thread = row) into global memory while the producers fill the next
//     tile; at the end every CTA writes its G partial and a second kernel adds the partials in a fixed order -- the weight
//     gradients are run-to-run deterministic (the FMA version combined its partials with float atomics).
//
// Precision: 3-term bf16 split = 16 significant bits per factor; measured against the FMA kernels in tests/test_train_tc_gpu.py
// (gradient bar of the golden fixtures: 2e-3 of the largest entry).
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace {

constexpr int BGK = 192;                      // gate columns of A
constexpr int BW_ATOM = 64 * 128;             // one [64 n x 64 k] bf16 block of the weight image
constexpr int BW_OFF_W_HI = 0, BW_OFF_W_LO = 3 * BW_ATOM;
constexpr int BW_IMAGE_BYTES = 6 * BW_ATOM;   // 48 KB: what tmpnn_pack_w_tc writes
constexpr int BW_OFF_BAR = BW_IMAGE_BYTES;    // 8 mbarriers + the TMEM pointer
constexpr int BW_OFF_A = 50 * 1024;           // images (1024-aligned)
constexpr int BW_IMG = TCM * 128;             // [128 rows x 64 columns] bf16 = 16 KB
// image slots: A hi 0..2, A lo 3..5 (column block a sits in slot (a + 1) % 3, so that both M blocks of the G product --
// columns [0,128) and [128,192) + [0,64) -- are two ADJACENT slots), X hi 6, X lo 7
constexpr int BW_OFF_T = BW_OFF_A + 8 * BW_IMG;          // epilogue transpose buffer: 4 warps x [32 rows x 64 floats]
constexpr int BW_SMEM = BW_OFF_T + 4 * 8192 + 1024;
constexpr int BW_EPI = 4, BW_PROD = 8;
constexpr int BW_THREADS = 32 * (BW_EPI + BW_PROD + 1);
constexpr int BW_TMEM_COLS = 256;             // C stage 0 | C stage 1 | G rows 0-127 | G rows 128-191 (+ a duplicate block)
static_assert(BW_SMEM <= 232448, "exceeds 227 KB");

__device__ __forceinline__ int slot_of(int a) { return a == 2 ? 0 : a + 1; }

// MN-major, SWIZZLE_128B descriptor: 64 elements (128 B) contiguous along M/N, further M/N blocks `lbo` bytes apart, 8 K
// indices 128 B apart inside a 1024-byte group, groups `sbo` bytes apart (CUTLASS make_umma_desc<Major::MN>, B128)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 with bf16 operands, fp32 accumulation; mn != 0: both operands MN-major
__device__ __forceinline__ constexpr uint32_t bw_idesc(int n, int mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn ? (1u << 15) | (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TCM >> 4) << 24);
}
__device__ __forceinline__ void split4_bf16(const float4 a, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(a.x, a.y), h23 = __floats2bfloat162_rn(a.z, a.w);
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(a.x - f01.x, a.y - f01.y), l23 = __floats2bfloat162_rn(a.z - f23.x, a.w - f23.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// W [192][ldw] fp32 (torch.nn.GRUCell weight, gate rows x input columns), columns [col0, col0 + 64) -> the B operand of
// C = A W: [64 n][192 k] K-major, three 64-k blocks of 128-byte rows, 128-byte swizzle, bf16 hi / lo
__global__ void k_pack_w_bf16(const float* __restrict__ W, int ldw, int col0, unsigned char* __restrict__ img) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 64 * BGK; i += gridDim.x * blockDim.x) {
    const int n = i / BGK, k = i % BGK;
    const float w = W[(size_t)k * ldw + col0 + n];
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
    const uint32_t off = (uint32_t)(k >> 6) * BW_ATOM + sw128(n, (k & 63) >> 3) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(img + BW_OFF_W_HI + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(img + BW_OFF_W_LO + off) = lo;
  }
}

__global__ void __launch_bounds__(BW_THREADS, 1)
k_rows_gemm_tc(const int32_t* __restrict__ r_dev, int r_host, const int32_t* __restrict__ a_rows,
               const int32_t* __restrict__ c_rows, const int32_t* __restrict__ x_rows, const int32_t* __restrict__ mask,
               const float* __restrict__ A, const unsigned char* __restrict__ w_image, float* __restrict__ C, int ldc,
               int accumulate, const float* __restrict__ X, int ldx, float* __restrict__ partials, int32_t* __restrict__ status) {
  extern __shared__ unsigned char smem_dyn[];
  const int R = r_dev ? *r_dev : r_host;
  const int total = (R + TCM - 1) / TCM;
  float* my_part = partials + (size_t)blockIdx.x * BGK * 64;
  if ((int)blockIdx.x >= total) return;  // no tile for this CTA (uniform exit before any barrier / TMEM use); k_reduce_partials
                                         // only reads the partials of the CTAs that had one
  unsigned char* sm = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sm_u = smem_u32(sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = sm_u + BW_OFF_BAR, bar_afree = bar_full + 8, bar_cdone = bar_full + 16, bar_cfree = bar_full + 32,
                 bar_gdone = bar_full + 48;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + BW_OFF_BAR + 64);
  {
    const uint4* gsrc = reinterpret_cast<const uint4*>(w_image);
    uint4* sdst = reinterpret_cast<uint4*>(sm);
    for (int i = threadIdx.x; i < BW_IMAGE_BYTES / 16; i += BW_THREADS) sdst[i] = __ldg(gsrc + i);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar_full, BW_PROD);   // one arrive per producer warp: images of the tile written
    mbar_init(bar_afree, 1);        // tcgen05.commit: the tile's MMAs retired, images reusable
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_cdone + 8 * s, 1);      // tcgen05.commit: C accumulator stage s complete
      mbar_init(bar_cfree + 8 * s, BW_EPI); // one arrive per epilogue warp: stage s drained
    }
    mbar_init(bar_gdone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BW_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int stride = gridDim.x;
  const uint32_t img_u = sm_u + BW_OFF_A;

  if (warp == BW_EPI + BW_PROD) {
    // ================= MMA issuer =================
    if (elect_one()) {  // one elected lane: descriptors stay in uniform registers, the MMAs issue back to back
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
        const int stage = it & 1;
        mbar_wait(bar_full, (uint32_t)it & 1u, status);
        mbar_wait(bar_cfree + 8 * stage, ((uint32_t)(it >> 1) & 1u) ^ 1u, status);
        tc_fence_after();
        // C = A W: K-major A images (row = M) x K-major weight blocks
        const uint32_t dc = tmem_base + (uint32_t)(64 * stage);
        uint32_t acc = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t a0 = img_u + (t == 1 ? 3 * BW_IMG : 0);
          const uint32_t w0 = sm_u + (t == 2 ? BW_OFF_W_LO : BW_OFF_W_HI);
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              umma_f16(dc, umma_desc(a0 + slot_of(a) * BW_IMG + 32 * j), umma_desc(w0 + a * BW_ATOM + 32 * j), bw_idesc(64, 0), acc);
              acc = 1;
            }
        }
        // G += A^T X: the same images as MN-major operands (64 columns = M / N, rows = K; 16 rows per instruction)
        const uint32_t first = it == 0 ? 0u : 1u;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const uint32_t a0 = img_u + (t == 1 ? 3 * BW_IMG : 0);
          const uint32_t x0 = img_u + (t == 2 ? 7 * BW_IMG : 6 * BW_IMG);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint64_t xd = umma_desc_mn(x0 + 2048 * j, BW_IMG, 1024);
            const uint32_t accg = (t == 0 && j == 0) ? first : 1u;
            // M block 0: columns [0,128) = slots 1, 2;  M block 1: columns [128,192) + [0,64) = slots 0, 1 (upper half unused)
            umma_f16(tmem_base + 128, umma_desc_mn(a0 + 1 * BW_IMG + 2048 * j, BW_IMG, 1024), xd, bw_idesc(64, 1), accg);
            umma_f16(tmem_base + 192, umma_desc_mn(a0 + 0 * BW_IMG + 2048 * j, BW_IMG, 1024), xd, bw_idesc(64, 1), accg);
          }
        }
        umma_commit(bar_afree);
        umma_commit(bar_cdone + 8 * stage);
      }
      umma_commit(bar_gdone);
    }
  } else if (warp >= BW_EPI) {
    // ================= producers: 16 lanes per row, rows g + 16 p =================
    // Software-pipelined in registers over "rounds" of two rows per thread (4 rounds per tile): the eight 128-bit loads of
    // round r + 1 are issued before round r is converted and stored, and the wait for the previous tile's MMAs sits between
    // the loads of a tile's first round and its first store -- the sweep is bound by HBM latency, so loads must stay in flight
    // across tile boundaries and across the tensor-core phase.
    const int pt = threadIdx.x - 32 * BW_EPI;
    const int g = pt >> 4, l = pt & 15;
    const uint32_t off0 = sw128(g, l >> 1) + ((l & 1) << 3);
    const int my_tiles = (total - (int)blockIdx.x + stride - 1) / stride;
    const int rounds = 4 * my_tiles;
    struct RoundRegs { float4 va[2][3]; float4 vx[2]; int mk[2]; };
    auto issue = [&](RoundRegs& rr, int r) {
      const int tile = blockIdx.x + (r >> 2) * stride;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int i0 = tile * TCM + g + 16 * (2 * (r & 3) + q);
        const int i = min(i0, R - 1);
        const int ra = a_rows ? __ldg(a_rows + i) : i;
        const int rx = x_rows ? __ldg(x_rows + i) : i;
        const float* ap = A + (size_t)ra * BGK + 4 * l;
        rr.va[q][0] = ldg4(ap); rr.va[q][1] = ldg4(ap + 64); rr.va[q][2] = ldg4(ap + 128);
        rr.vx[q] = ldg4(X + (size_t)rx * ldx + 4 * l);
        // validity, resolved when the round is processed: rows past the end and masked rows read a valid row and are zeroed
        rr.mk[q] = i0 < R ? (mask ? __ldg(mask + ra) : 0) : -1;
      }
    };
    auto process = [&](const RoundRegs& rr, int r) {
      const int sub = r & 3;
      if (sub == 0) mbar_wait(bar_afree, ((uint32_t)(r >> 2) & 1u) ^ 1u, status);  // the previous tile's MMAs have read the images
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const bool ok = rr.mk[q] >= 0;
        const uint32_t off = off0 + 2048u * (2 * sub + q);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          uint2 hi, lo;
          split4_bf16(ok ? rr.va[q][a] : z4, hi, lo);
          *reinterpret_cast<uint2*>(sm + BW_OFF_A + slot_of(a) * BW_IMG + off) = hi;
          *reinterpret_cast<uint2*>(sm + BW_OFF_A + (3 + slot_of(a)) * BW_IMG + off) = lo;
        }
        uint2 hi, lo;
        split4_bf16(ok ? rr.vx[q] : z4, hi, lo);
        *reinterpret_cast<uint2*>(sm + BW_OFF_A + 6 * BW_IMG + off) = hi;
        *reinterpret_cast<uint2*>(sm + BW_OFF_A + 7 * BW_IMG + off) = lo;
      }
      if (sub == 3) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full);
      }
    };
    RoundRegs ra0, ra1;
    issue(ra0, 0);
    for (int r = 0; r < rounds; r += 2) {   // rounds is a multiple of 4
      issue(ra1, r + 1);
      process(ra0, r);
      if (r + 2 < rounds) issue(ra0, r + 2);
      process(ra1, r + 1);
    }
  } else {
    // ================= epilogue: warp w owns TMEM lanes [32 w, 32 w + 32) =================
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += stride, ++it) {
      const int stage = it & 1;
      const int i = tile * TCM + warp * 32 + lane;
      int rc = -1;
      if (i < R) {
        const int ra = a_rows ? __ldg(a_rows + i) : i;
        if (!mask || __ldg(mask + ra) >= 0) rc = c_rows ? __ldg(c_rows + i) : i;
      }
      mbar_wait(bar_cdone + 8 * stage, (uint32_t)(it >> 1) & 1u, status);
      tc_fence_after();
      // TMEM -> registers (thread = row) -> this warp's [32 rows x 64 floats] buffer (16-byte chunks XOR-swizzled by row), then
      // the accumulator stage is free; the rows go out 2 per instruction as full 256-byte rows
      unsigned char* tb = sm + BW_OFF_T + warp * 8192;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float v[16];
        tmem_ld16(lane_base + (uint32_t)(64 * stage + 16 * ch), v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(tb + lane * 256 + (((4 * ch + q) ^ (lane & 15)) << 4)) =
              make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_cfree + 8 * stage);
#pragma unroll 4
      for (int k = 0; k < 16; ++k) {
        const int rr = 2 * k + (lane >> 4), cc = lane & 15;
        const int rcr = __shfl_sync(0xffffffffu, rc, rr);
        if (rcr >= 0) {
          float4 o = *reinterpret_cast<const float4*>(tb + rr * 256 + ((cc ^ (rr & 15)) << 4));
          float4* dst = reinterpret_cast<float4*>(C + (size_t)rcr * ldc + 4 * cc);
          if (accumulate) {
            const float4 pv = *dst;
            o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w;
          }
          *dst = o;
        }
      }
      __syncwarp();   // the buffer is rewritten by the next tile
    }
    // the weight-gradient partial of this CTA: rows [0,128) from the first M block, [128,192) from the lower half of the second
    mbar_wait(bar_gdone, 0u, status);
    tc_fence_after();
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
      const int m = blk * 128 + warp * 32 + lane;
      if (blk == 1 && warp >= 2) break;  // warp-uniform: only 64 rows of the second block are real
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        float v[16];
        tmem_ld16(lane_base + (uint32_t)(128 + 64 * blk + 16 * ch), v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(my_part + (size_t)m * 64 + 16 * ch + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BW_TMEM_COLS) : "memory");
  }
}

// G[m][0:64] += sum over the CTA partials in a fixed order, so the weight gradients are reproducible bit for bit.  64 outputs
// per CTA x 16 groups of partials (group p takes CTAs p, p + 16, ...: <= 10 independent loads per thread, all in flight at
// once; one thread per output walking all 148 partials was latency bound at ~11 us per call, 32 calls per training step)
__global__ void __launch_bounds__(1024) k_reduce_partials(const float* __restrict__ partials, int n_sm, const int32_t* __restrict__ r_dev,
                                                          int r_host, float* __restrict__ G, int ldg) {
  __shared__ float part[16][64];
  const int il = threadIdx.x & 63, pg = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + il;
  // only the CTAs that had a tile hold a partial (the others did not even write zeros)
  const int R = r_dev ? *r_dev : r_host;
  const int n_part = min(n_sm, (R + TCM - 1) / TCM);
  float v[10];
#pragma unroll
  for (int q = 0; q < 10; ++q) {
    const int c = pg + 16 * q;
    v[q] = c < n_part ? partials[(size_t)c * BGK * 64 + i] : 0.f;
  }
  float s = v[0];
#pragma unroll
  for (int q = 1; q < 10; ++q) s += v[q];
  part[pg][il] = s;
  __syncthreads();
  if (pg == 0) {
    float t = part[0][il];
#pragma unroll
    for (int p = 1; p < 16; ++p) t += part[p][il];
    G[(size_t)(i / 64) * ldg + (i % 64)] += t;
  }
}

bool g_bw_init = false;

}  // namespace

int tmpnn_init_train_tc() {
  if (!g_bw_init) {
    TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_rows_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
    g_bw_init = true;
  }
  return TMPNN_OK;
}

extern "C" size_t tmpnn_bwd_tc_image_bytes(void) { return (size_t)BW_IMAGE_BYTES; }
extern "C" size_t tmpnn_bwd_tc_partial_floats(void) { return (size_t)TMPNN_SM_COUNT * BGK * 64; }

extern "C" int tmpnn_pack_w_tc(const float* W, int ldw, int col0, void* image, void* stream) {
  TMPNN_REQUIRE(W && image && ldw >= 64 && col0 >= 0 && col0 + 64 <= ldw, "bad argument");
  TMPNN_REQUIRE(((uintptr_t)image & 15) == 0, "image must be 16-byte aligned");
  k_pack_w_bf16<<<48, 256, 0, (cudaStream_t)stream>>>(W, ldw, col0, (unsigned char*)image);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_rows_gemm_tc(const int32_t* r_dev, int r_host, const int32_t* a_rows, const int32_t* c_rows,
                                  const int32_t* x_rows, const int32_t* mask, const float* A, const void* w_image, float* C,
                                  int ldc, int accumulate, const float* X, int ldx, float* partials, float* G, int ldg,
                                  int32_t* status, void* stream) {
  TMPNN_REQUIRE(A && w_image && C && X && partials && G && status, "null argument");
  TMPNN_REQUIRE(ldc % 4 == 0 && ldx % 4 == 0 && ldg >= 64, "rows must be 16-byte aligned");
  TMPNN_REQUIRE((((uintptr_t)C | (uintptr_t)X | (uintptr_t)A | (uintptr_t)partials) & 15) == 0, "buffers must be 16-byte aligned");
  if (!r_dev && r_host <= 0) return TMPNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  { int rc0 = tmpnn_init_train_tc(); if (rc0) return rc0; }
  k_rows_gemm_tc<<<TMPNN_SM_COUNT, BW_THREADS, BW_SMEM, st>>>(r_dev, r_host, a_rows, c_rows, x_rows, mask, A,
                                                             (const unsigned char*)w_image, C, ldc, accumulate, X, ldx, partials,
                                                             status);
  TMPNN_LAUNCH_CHECK();
  static_assert(TMPNN_SM_COUNT <= 160, "k_reduce_partials: 16 groups x 10 partials");
  k_reduce_partials<<<BGK, 1024, 0, st>>>(partials, TMPNN_SM_COUNT, r_dev, r_host, G, ldg);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
