// mp_step.cu -- K0 input transform, K1 aggregation, K2+K3 fused GRU step + heads (fp32 FMA path).
//
// Reference behaviour: models/track_mpnn.py:54-75 and models/layers.py:84-116 of
// arangesh/TrackMPNN, restated on the edge list:
//   edge row e: h'[e] = GRUCell_edge(x = h[src]-h[dst] (diff) | [h[src] | h[dst]] (concat), h[e])
//   det  row d: h'[d] = GRUCell_node(x = sum_{e: src=d} h[e] - sum_{e: dst=d} h[e],         h[d])
//   logit = w_type . h' + b_type, score = sigmoid(logit)
// All reads use the old h (Jacobi), as layers.py:97 and :114 both consume h.
//
// Kernel shape: one persistent CTA per SM (grid = 148) keeps one packed GRU cell (96 KB for
// KX=64, 144 KB for KX=128) resident in shared memory and walks 64-row tiles.  A tile's
// GEMMs [64 x KX].[KX x 192] + [64 x 64].[64 x 192] are register tiled: each thread owns
// 8 rows x 2 hidden units x 4 gate accumulators (r, z, i_n, h_n); x rows come from shared
// memory as broadcast 128-bit loads, weights as conflict-free 64-bit loads.  The gate
// non-linearities, the head dot product (warp shuffle reduction over the 64 hidden units)
// and the sigmoid are the epilogue, so h' / logit / score are written exactly once.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------
// error slot
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int tmpnn_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char* tmpnn_last_error(void) { return g_err; }
extern "C" int tmpnn_version(void) { return 100; }

// ------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------
constexpr int H = TMPNN_HIDDEN;
constexpr int TM = TMPNN_TILE_ROWS;
constexpr int NT = 256;

__host__ __device__ constexpr int pack_off_whh(int kx) { return kx * 3 * H; }
__host__ __device__ constexpr int pack_off_bias(int kx) { return pack_off_whh(kx) + H * 3 * H; }
__host__ __device__ constexpr int pack_off_headw(int kx) { return pack_off_bias(kx) + 4 * H; }
__host__ __device__ constexpr int pack_off_headb(int kx) { return pack_off_headw(kx) + H; }
__host__ __device__ constexpr int pack_floats(int kx) { return pack_off_headb(kx) + 4; }

extern "C" size_t tmpnn_gru_pack_floats(int kx) { return (size_t)pack_floats(kx); }

__global__ void k_pack_gru(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                           const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                           const float* __restrict__ head_w, const float* __restrict__ head_b, int kx,
                           float* __restrict__ out) {
  const int n_ih = kx * 3 * H, n_hh = H * 3 * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_ih + n_hh + 4 * H + H + 4; i += gridDim.x * blockDim.x) {
    float v;
    if (i < n_ih) {  // out[k][g][j] = w_ih[g*64 + j][k]
      int k = i / (3 * H), gj = i % (3 * H);
      v = w_ih[(size_t)gj * kx + k];
    } else if (i < n_ih + n_hh) {
      int q = i - n_ih;
      int k = q / (3 * H), gj = q % (3 * H);
      v = w_hh[(size_t)gj * H + k];
    } else if (i < n_ih + n_hh + 4 * H) {
      int q = i - n_ih - n_hh;
      int g = q / H, j = q % H;
      if (g < 2) v = b_ih[g * H + j] + b_hh[g * H + j];
      else if (g == 2) v = b_ih[2 * H + j];
      else v = b_hh[2 * H + j];
    } else if (i < n_ih + n_hh + 4 * H + H) {
      v = head_w[i - (n_ih + n_hh + 4 * H)];
    } else {
      v = (i == n_ih + n_hh + 4 * H + H) ? head_b[0] : 0.f;
    }
    out[i] = v;
  }
}

extern "C" int tmpnn_pack_gru(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                              const float* head_w, const float* head_b, int kx, float* packed, void* stream) {
  TMPNN_REQUIRE(kx == 64 || kx == 128, "kx must be 64 or 128");
  k_pack_gru<<<64, 256, 0, (cudaStream_t)stream>>>(w_ih, w_hh, b_ih, b_hh, head_w, head_b, kx, packed);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// ------------------------------------------------------------------------------------------
// K0: input transform
// ------------------------------------------------------------------------------------------
// One warp per detection row, lane owns hidden units (lane, lane+32).
__global__ void k_input_linear1(const float* __restrict__ x, int ldx, int col0, int f_in,
                                const int32_t* __restrict__ x_idx, const float* __restrict__ w1,
                                const float* __restrict__ b1, float* __restrict__ a,
                                const int32_t* __restrict__ n_dev, int n_host) {
  extern __shared__ float sm[];  // w1 transposed [f_in][64]
  const int n = n_dev ? *n_dev : n_host;
  for (int i = threadIdx.x; i < f_in * H; i += blockDim.x) {
    int k = i / H, j = i % H;
    sm[i] = w1[j * f_in + k];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
    const float* xr = x + (size_t)(x_idx ? x_idx[r] : r) * ldx + col0;
    float a0 = b1[lane], a1 = b1[lane + 32];
    for (int k = 0; k < f_in; ++k) {
      float xv = __ldg(xr + k);
      a0 = fmaf(xv, sm[k * H + lane], a0);
      a1 = fmaf(xv, sm[k * H + lane + 32], a1);
    }
    a[(size_t)r * H + lane] = a0;
    a[(size_t)r * H + lane + 32] = a1;
  }
}

extern "C" int tmpnn_input_linear1(const float* x, int ldx, int col0, int f_in, const int32_t* x_idx,
                                   const float* w1, const float* b1, float* a, const int32_t* n_dev, int n_host,
                                   void* stream) {
  TMPNN_REQUIRE(f_in > 0 && f_in <= 512, "f_in out of range");
  int n_max = n_dev ? TMPNN_SM_COUNT * 8 * 4 : n_host;
  if (n_max <= 0) return TMPNN_OK;
  int blocks = min(tmpnn_div_up(n_max, 4), TMPNN_SM_COUNT * 8);
  size_t smem = (size_t)f_in * H * sizeof(float);
  if (smem > 48 * 1024) TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_input_linear1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_input_linear1<<<blocks, 128, smem, (cudaStream_t)stream>>>(x, ldx, col0, f_in, x_idx, w1, b1, a, n_dev, n_host);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// Deterministic two-pass batch statistics in one CTA: thread j owns channel j % 64, part j / 64.
__global__ void k_input_bn_stats(const float* __restrict__ a, int n, int n_edge, const float* __restrict__ b1,
                                 float* __restrict__ stats, float* __restrict__ rmean, float* __restrict__ rvar) {
  __shared__ double red[4][H];
  const int j = threadIdx.x & 63, part = threadIdx.x >> 6;  // 256 threads = 4 parts
  const double ntot = (double)n + (double)n_edge;
  double s = 0;
  for (int r = part; r < n; r += 4) s += a[(size_t)r * H + j];
  red[part][j] = s;
  __syncthreads();
  const double mu = (red[0][j] + red[1][j] + red[2][j] + red[3][j] + (double)n_edge * b1[j]) / ntot;
  __syncthreads();
  double q = 0;
  for (int r = part; r < n; r += 4) {
    double d = a[(size_t)r * H + j] - mu;
    q += d * d;
  }
  red[part][j] = q;
  __syncthreads();
  if (part == 0) {
    double db = (double)b1[j] - mu;
    double var = (red[0][j] + red[1][j] + red[2][j] + red[3][j] + (double)n_edge * db * db) / ntot;
    stats[j] = (float)mu;
    stats[H + j] = (float)var;
    if (rmean) rmean[j] = 0.9f * rmean[j] + 0.1f * (float)mu;
    if (rvar) rvar[j] = 0.9f * rvar[j] + 0.1f * (float)(var * ntot / (ntot - 1.0));
  }
}

extern "C" int tmpnn_input_bn_stats(const float* a, int n, int n_edge_rows, const float* b1, float* stats,
                                    float* running_mean, float* running_var, void* stream) {
  TMPNN_REQUIRE(n + n_edge_rows > 1, "Expected more than 1 value per channel when training");
  k_input_bn_stats<<<1, 256, 0, (cudaStream_t)stream>>>(a, n, n_edge_rows, b1, stats, running_mean, running_var);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

__global__ void k_input_bn_relu_linear2(const float* __restrict__ a, const float* __restrict__ mean,
                                        const float* __restrict__ var, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ w2,
                                        const float* __restrict__ b2, float* __restrict__ h, int ldh, int col,
                                        const int32_t* __restrict__ out_rows, const int32_t* __restrict__ n_dev,
                                        int n_host) {
  __shared__ float w2t[H * H];   // [k][j]
  __shared__ float act[4][H];
  const int n = n_dev ? *n_dev : n_host;
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) {
    int k = i / H, j = i % H;
    w2t[i] = w2[j * H + k];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int r = blockIdx.x * 4 + w; r < n; r += gridDim.x * 4) {
    // (a - mean) / sqrt(var + eps) * gamma + beta, written like torch's batch_norm
    float v0 = (a[(size_t)r * H + lane] - mean[lane]) * (1.0f / sqrtf(var[lane] + 1e-5f)) * gamma[lane] + beta[lane];
    float v1 = (a[(size_t)r * H + lane + 32] - mean[lane + 32]) * (1.0f / sqrtf(var[lane + 32] + 1e-5f)) * gamma[lane + 32] + beta[lane + 32];
    act[w][lane] = fmaxf(v0, 0.f);
    act[w][lane + 32] = fmaxf(v1, 0.f);
    __syncwarp();
    float o0 = b2[lane], o1 = b2[lane + 32];
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      float av = act[w][k];
      o0 = fmaf(av, w2t[k * H + lane], o0);
      o1 = fmaf(av, w2t[k * H + lane + 32], o1);
    }
    __syncwarp();
    float* hr = h + (size_t)out_rows[r] * ldh + col;
    hr[lane] = o0;
    hr[lane + 32] = o1;
  }
}

extern "C" int tmpnn_input_bn_relu_linear2(const float* a, const float* mean, const float* var, const float* gamma,
                                           const float* beta, const float* w2, const float* b2, float* h, int ldh,
                                           int col, const int32_t* out_rows, const int32_t* n_dev, int n_host,
                                           void* stream) {
  int n_max = n_dev ? TMPNN_SM_COUNT * 8 * 4 : n_host;
  if (n_max <= 0) return TMPNN_OK;
  int blocks = min(tmpnn_div_up(n_max, 4), TMPNN_SM_COUNT * 8);
  k_input_bn_relu_linear2<<<blocks, 128, 0, (cudaStream_t)stream>>>(a, mean, var, gamma, beta, w2, b2, h, ldh, col,
                                                                   out_rows, n_dev, n_host);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// ------------------------------------------------------------------------------------------
// K1: aggregation
// ------------------------------------------------------------------------------------------
// One CTA per detection: its 16 half-warps stream the incident rows round-robin (16 lanes x float4 = one 256 B
// row per half-warp per load, 4 rows in flight per half-warp), past edges subtract, future edges add; the 16
// partial sums are combined in a fixed order -> bit-reproducible.  A CTA (not a warp) per detection keeps the
// detections in flight at any time within ~3 sequences (~50 MB of state), so the second read of every
// association row -- each is incident to two detections -- hits the 126 MB L2 instead of HBM.
__global__ void __launch_bounds__(256) k_aggregate_dets(const float* __restrict__ h, int ldh, int col,
                                                        const int32_t* __restrict__ n_dets,
                                                        const int32_t* __restrict__ seg_ptr,
                                                        const int32_t* __restrict__ inc, float* __restrict__ agg,
                                                        const int32_t* __restrict__ phys) {
  __shared__ float4 part[16][16];
  const int nd = *n_dets;
  const int q = threadIdx.x >> 4, l16 = threadIdx.x & 15;  // half-warp, float4 within the row
  for (int k = blockIdx.x; k < nd; k += gridDim.x) {
    const int s0 = seg_ptr[2 * k], s1 = seg_ptr[2 * k + 1], s2 = seg_ptr[2 * k + 2];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = s0 + q;
    for (; i + 48 < s2; i += 64) {
      int e0 = inc[i], e1 = inc[i + 16], e2 = inc[i + 32], e3 = inc[i + 48];
      if (phys) { e0 = phys[e0]; e1 = phys[e1]; e2 = phys[e2]; e3 = phys[e3]; }  // deferred compaction: physical rows
      float4 v0 = ldg4(h + (size_t)e0 * ldh + col + 4 * l16);
      float4 v1 = ldg4(h + (size_t)e1 * ldh + col + 4 * l16);
      float4 v2 = ldg4(h + (size_t)e2 * ldh + col + 4 * l16);
      float4 v3 = ldg4(h + (size_t)e3 * ldh + col + 4 * l16);
      float g0 = (i < s1) ? -1.f : 1.f, g1 = (i + 16 < s1) ? -1.f : 1.f;
      float g2 = (i + 32 < s1) ? -1.f : 1.f, g3 = (i + 48 < s1) ? -1.f : 1.f;
      acc.x = fmaf(g0, v0.x, acc.x); acc.y = fmaf(g0, v0.y, acc.y); acc.z = fmaf(g0, v0.z, acc.z); acc.w = fmaf(g0, v0.w, acc.w);
      acc.x = fmaf(g1, v1.x, acc.x); acc.y = fmaf(g1, v1.y, acc.y); acc.z = fmaf(g1, v1.z, acc.z); acc.w = fmaf(g1, v1.w, acc.w);
      acc.x = fmaf(g2, v2.x, acc.x); acc.y = fmaf(g2, v2.y, acc.y); acc.z = fmaf(g2, v2.z, acc.z); acc.w = fmaf(g2, v2.w, acc.w);
      acc.x = fmaf(g3, v3.x, acc.x); acc.y = fmaf(g3, v3.y, acc.y); acc.z = fmaf(g3, v3.z, acc.z); acc.w = fmaf(g3, v3.w, acc.w);
    }
    for (; i < s2; i += 16) {
      const int e = phys ? phys[inc[i]] : inc[i];
      float4 v = ldg4(h + (size_t)e * ldh + col + 4 * l16);
      float g = (i < s1) ? -1.f : 1.f;
      acc.x = fmaf(g, v.x, acc.x); acc.y = fmaf(g, v.y, acc.y); acc.z = fmaf(g, v.z, acc.z); acc.w = fmaf(g, v.w, acc.w);
    }
    part[q][l16] = acc;
    __syncthreads();
    if (q == 0) {
      float4 t = part[0][l16];
#pragma unroll
      for (int w = 1; w < 16; ++w) {
        const float4 u = part[w][l16];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      *reinterpret_cast<float4*>(agg + (size_t)k * H + 4 * l16) = t;
    }
    __syncthreads();
  }
}

extern "C" int tmpnn_aggregate_dets(const tmpnn_graph* g, const tmpnn_index* ix, const float* h, int ldh, int col,
                                    float* agg, void* stream) {
  TMPNN_REQUIRE(g && ix && h && agg, "null argument");
  TMPNN_REQUIRE(ldh % 4 == 0 && col % 4 == 0, "h rows must be 16-byte aligned");
  k_aggregate_dets<<<TMPNN_SM_COUNT * 8, 256, 0, (cudaStream_t)stream>>>(h, ldh, col, ix->n_dets, ix->seg_ptr, ix->inc, agg,
                                                                         g->phys);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// ---- block-structured K1-det: every association row read ONCE ------------------------------------------------
// The engine's window graphs are chains of dense edge blocks (graph_index.cu, tmpnn_index_build_structured): block b of a
// slab holds rows e0 + a nt + j joining source a to detection j of the following detection segment.  k_aggregate_dets
// walks the incidence lists, so every association row is requested twice (once by each endpoint; the second request only
// hits L2 while the detections in flight stay within a few sequences) plus 8 B of incidence entries per row.  Here a CTA
// takes a stripe of AB_SA sources x all nt columns of one block and forms, in one pass, the stripe's complete row sums
// (the future-run sum of each source) and its column partials (the past sum of each detection over the stripe's
// sources); k_aggregate_blocks_combine then adds, per detection and in a fixed order, its column partials (-) and its run
// sums (+).  A block appended this frame is all zero (its rows alias the slab's zero row) and is skipped outright.
// Same sums as k_aggregate_dets up to fp32 re-association; bit-reproducible (no atomics, fixed orders, the partition into
// stripes depends on the slab's own structure only).
constexpr int AB_SA = 32;  // sources per stripe
struct AggBlock { int32_t e0, nt, A, rbase, cbase, tile0, fresh, nst; };
struct AggSlab {
  int32_t nblk, ntiles, pad[6];
  AggBlock b[MAXE];  // by edge-segment ordinal
};

__global__ void __launch_bounds__(128) k_agg_block_table(const SlabSegs* __restrict__ segs, int num_seqs, int cap_rows,
                                                         const int32_t* __restrict__ phys, int cap_runs, int cap_cpart,
                                                         AggSlab* __restrict__ tab, int32_t* __restrict__ status) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= num_seqs) return;
  const SlabSegs& sg = segs[s];
  AggSlab& o = tab[s];
  const size_t base = (size_t)s * cap_rows;
  const int32_t zrow = (int32_t)(base + cap_rows - 1);
  int runs = 0, cp = 0, tiles = 0;
  for (int q = 0; q < sg.nseg; ++q) {
    const int b = sg.eord[q];
    if (b < 0) continue;
    const int e0 = sg.start[q], d0 = sg.start[q + 1], d1 = q + 2 <= sg.nseg ? sg.start[q + 2] : d0;
    const int nt = d1 - d0;
    const bool ok = nt > 0 && q + 1 < sg.nseg && sg.eord[q + 1] < 0 && (d0 - e0) % nt == 0;  // else TMPNN_FLAG_UNSTRUCTURED is set
    int A = ok ? (d0 - e0) / nt : 0;
    const int fresh = (phys && A > 0 && phys[base + e0] == zrow) ? 1 : 0;
    int nst = fresh ? 0 : (A + AB_SA - 1) / AB_SA;
    if (!fresh && ((long long)runs + A > cap_runs || (long long)cp + (long long)nst * nt > cap_cpart)) {
      atomicOr(status, TMPNN_FLAG_INC_CAPACITY);
      A = 0; nst = 0;
    }
    AggBlock B;
    B.e0 = e0; B.nt = max(nt, 1); B.A = A; B.rbase = runs; B.cbase = cp; B.tile0 = tiles; B.fresh = fresh; B.nst = nst;
    o.b[b] = B;
    if (!fresh) { runs += A; cp += nst * nt; }
    tiles += nst;
  }
  o.nblk = sg.nedge;
  o.ntiles = tiles;
}

__global__ void __launch_bounds__(128) k_aggregate_blocks(const AggSlab* __restrict__ tab, const float* __restrict__ h, int ldh,
                                                          int col, int cap_rows, const int32_t* __restrict__ phys,
                                                          float* __restrict__ rsum, int cap_runs, float* __restrict__ cpart,
                                                          int cap_cpart) {
  __shared__ float4 rowpart[AB_SA][4][16];  // [source of the stripe][warp][float4 of the row]: 32 KB
  __shared__ int32_t idx[AB_SA][64];        // physical rows of the stripe x column chunk: 8 KB
  const int s = blockIdx.y;
  const AggSlab& sl = tab[s];
  const int ntiles = sl.ntiles;
  if ((int)blockIdx.x >= ntiles) return;
  const size_t base = (size_t)s * cap_rows;
  const int hw = threadIdx.x >> 4, l16 = threadIdx.x & 15, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* __restrict__ rs = rsum + (size_t)s * cap_runs * H;
  float* __restrict__ cpb = cpart + (size_t)s * cap_cpart * H;
  const float* __restrict__ hc = h + col + 4 * l16;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    int b = 0;
    while (b + 1 < sl.nblk && t >= sl.b[b].tile0 + sl.b[b].nst) ++b;
    const AggBlock B = sl.b[b];
    const int i = t - B.tile0, a0 = i * AB_SA, na = min(AB_SA, B.A - a0), nt = B.nt;
    // column chunks of equal width (<= 64, a multiple of 8): half-warp hw owns columns c0 + hw + 8 u, u < cw / 8, so
    // consecutive half-warps read consecutive rows and every chunk keeps all half-warps equally busy
    const int nch = (nt + 63) >> 6;
    const int cw = 8 * ((nt + 8 * nch - 1) / (8 * nch));
    for (int c0 = 0; c0 < nt; c0 += cw) {
      // the chunk's physical rows first (coalesced; deferred compaction scatters them), so that the state loads below
      // are independent of each other
      __syncthreads();
      for (int o = threadIdx.x; o < na * cw; o += 128) {
        const int al = o / cw, jj = o - al * cw, j = c0 + jj;
        const size_t r = base + B.e0 + (size_t)(a0 + al) * nt + j;
        idx[al][jj] = j < nt ? (phys ? __ldg(phys + r) : (int32_t)r) : -1;
      }
      __syncthreads();
      float4 ca[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) ca[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int nu = cw >> 3;
      for (int al = 0; al < na; al += 2) {
        // two sources per pass: up to 16 independent 16-byte loads in flight per thread
        float4 v[2][8];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int pr = (u < nu && al + w < na) ? idx[al + w][hw + 8 * u] : -1;
            v[w][u] = pr >= 0 ? ldg4(hc + (size_t)pr * ldh) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          float4 rp = v[w][0];
          ca[0].x += v[w][0].x; ca[0].y += v[w][0].y; ca[0].z += v[w][0].z; ca[0].w += v[w][0].w;
#pragma unroll
          for (int u = 1; u < 8; ++u) {
            rp.x += v[w][u].x; rp.y += v[w][u].y; rp.z += v[w][u].z; rp.w += v[w][u].w;
            ca[u].x += v[w][u].x; ca[u].y += v[w][u].y; ca[u].z += v[w][u].z; ca[u].w += v[w][u].w;
          }
          // the warp's two half-warps hold different columns of the same source row
          rp.x += __shfl_xor_sync(0xffffffffu, rp.x, 16);
          rp.y += __shfl_xor_sync(0xffffffffu, rp.y, 16);
          rp.z += __shfl_xor_sync(0xffffffffu, rp.z, 16);
          rp.w += __shfl_xor_sync(0xffffffffu, rp.w, 16);
          if (lane < 16 && al + w < na) {
            float4* p = &rowpart[al + w][warp][l16];
            if (c0 != 0) {
              const float4 o = *p;
              rp.x += o.x; rp.y += o.y; rp.z += o.z; rp.w += o.w;
            }
            *p = rp;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = c0 + hw + 8 * u;
        if (u < nu && j < nt) *reinterpret_cast<float4*>(cpb + ((size_t)B.cbase + (size_t)i * nt + j) * H + 4 * l16) = ca[u];
      }
    }
    __syncthreads();
    for (int o = threadIdx.x; o < na * 16; o += 128) {
      const int al = o >> 4, l = o & 15;
      float4 t4 = rowpart[al][0][l];
#pragma unroll
      for (int w = 1; w < 4; ++w) {
        const float4 u4 = rowpart[al][w][l];
        t4.x += u4.x; t4.y += u4.y; t4.z += u4.z; t4.w += u4.w;
      }
      *reinterpret_cast<float4*>(rs + ((size_t)B.rbase + a0 + al) * H + 4 * l) = t4;
    }
  }
}

// half-warp per detection: - column partials of the block in front of its segment, + run sums of its future runs
__global__ void __launch_bounds__(256) k_aggregate_blocks_combine(const SlabSegs* __restrict__ segs, const AggSlab* __restrict__ tab,
                                                                  const int32_t* __restrict__ n_dets,
                                                                  const int32_t* __restrict__ det_rows,
                                                                  const int32_t* __restrict__ seg_ptr,
                                                                  const int32_t* __restrict__ inc,
                                                                  const int32_t* __restrict__ futoff, int cap_rows,
                                                                  const float* __restrict__ rsum, int cap_runs,
                                                                  const float* __restrict__ cpart, int cap_cpart,
                                                                  float* __restrict__ agg) {
  const int k = (blockIdx.x * 256 + threadIdx.x) >> 4, l16 = threadIdx.x & 15;
  if (k >= *n_dets) return;
  const int row = det_rows[k];
  const int s = row / cap_rows, dl = row - s * cap_rows;
  const SlabSegs& sg = segs[s];
  const AggSlab& sl = tab[s];
  int qd = 0;
  while (qd + 1 < sg.nseg && sg.start[qd + 1] <= dl) ++qd;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (qd >= 1 && sg.eord[qd - 1] >= 0) {
    const AggBlock B = sl.b[sg.eord[qd - 1]];
    if (B.A > 0 && !B.fresh) {
      const float* p = cpart + ((size_t)s * cap_cpart + B.cbase + (dl - sg.start[qd])) * H + 4 * l16;
      for (int i = 0; i < B.nst; ++i) {
        const float4 v = ldg4(p + (size_t)i * B.nt * H);
        acc.x -= v.x; acc.y -= v.y; acc.z -= v.z; acc.w -= v.w;
      }
    }
  }
  const int f0 = seg_ptr[2 * k + 1], flen = seg_ptr[2 * k + 2] - f0;
  if (flen > 0) {
    const int32_t* f = futoff + (size_t)k * MAXE;
    for (int b = 0; b < sl.nblk; ++b) {
      const int o0 = f[b], o1 = b + 1 < sl.nblk ? f[b + 1] : flen;
      if (o1 <= o0) continue;
      const AggBlock B = sl.b[b];
      if (B.fresh || B.A == 0) continue;
      const int e = inc[f0 + o0] - s * cap_rows;  // first row of the run
      const int a = (e - B.e0) / B.nt;
      const float4 v = ldg4(rsum + ((size_t)s * cap_runs + B.rbase + a) * H + 4 * l16);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  *reinterpret_cast<float4*>(agg + (size_t)k * H + 4 * l16) = acc;
}

extern "C" size_t tmpnn_aggregate_blocks_scratch_bytes(int num_seqs, int cap_runs, int cap_cpart) {
  return (size_t)num_seqs * sizeof(AggSlab) + (size_t)num_seqs * ((size_t)cap_runs + (size_t)cap_cpart) * H * sizeof(float) + 256;
}

extern "C" int tmpnn_aggregate_dets_blocks(const tmpnn_graph* g, const tmpnn_index* ix, const void* index_scratch2,
                                           const float* h, int ldh, int col, float* agg, void* scratch, int cap_runs,
                                           int cap_cpart, void* stream) {
  TMPNN_REQUIRE(g && ix && index_scratch2 && h && agg && scratch, "null argument");
  TMPNN_REQUIRE(ldh % 4 == 0 && col % 4 == 0, "h rows must be 16-byte aligned");
  TMPNN_REQUIRE(cap_runs > 0 && cap_cpart > 0 && ((uintptr_t)scratch & 15) == 0, "bad scratch");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = g->num_seqs;
  const SlabSegs* segs = (const SlabSegs*)index_scratch2;
  // the run offsets of k_block_futoff sit behind the segment tables (tmpnn_index_build_structured's layout)
  const int32_t* futoff = (const int32_t*)((const unsigned char*)index_scratch2 + (size_t)S * sizeof(SlabSegs));
  AggSlab* tab = (AggSlab*)scratch;
  float* rsum = (float*)((unsigned char*)scratch + (((size_t)S * sizeof(AggSlab) + 255) & ~(size_t)255));
  float* cpart = rsum + (size_t)S * cap_runs * H;
  k_agg_block_table<<<tmpnn_div_up(S, 128), 128, 0, st>>>(segs, S, g->cap_rows, g->phys, cap_runs, cap_cpart, tab, g->status);
  TMPNN_LAUNCH_CHECK();
  // CTAs per slab: enough for the largest windows (a stripe is AB_SA x nt rows), and at least ~6 per SM in total
  const int gx = max(1, min(tmpnn_div_up(g->cap_rows, 64 * AB_SA), max(64, tmpnn_div_up(TMPNN_SM_COUNT * 6, S))));
  k_aggregate_blocks<<<dim3(gx, S), 128, 0, st>>>(tab, h, ldh, col, g->cap_rows, g->phys, rsum, cap_runs, cpart, cap_cpart);
  TMPNN_LAUNCH_CHECK();
  k_aggregate_blocks_combine<<<tmpnn_div_up(ix->cap_dets, 16), 256, 0, st>>>(segs, tab, ix->n_dets, ix->det_rows, ix->seg_ptr, ix->inc,
                                                                            futoff, g->cap_rows, rsum, cap_runs, cpart, cap_cpart, agg);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// Stand-alone node_support: half-warp per edge row.
__global__ void __launch_bounds__(256) k_aggregate_edges(const float* __restrict__ h, int ldh, int col, int concat,
                                                         const int32_t* __restrict__ n_rows,
                                                         const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                         int cap_rows, int num_seqs, float* __restrict__ support) {
  const int l16 = threadIdx.x & 15;
  const int hw_per_block = blockDim.x >> 4;
  const int s = blockIdx.y;
  const int n = n_rows[s];
  const int width = concat ? 2 * H : H;
  for (int r = blockIdx.x * hw_per_block + (threadIdx.x >> 4); r < n; r += gridDim.x * hw_per_block) {
    const size_t row = (size_t)s * cap_rows + r;
    const int a = src[row], b = dst[row];
    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
    if (a >= 0) {
      float4 va = ldg4(h + ((size_t)s * cap_rows + a) * ldh + col + 4 * l16);
      float4 vb = ldg4(h + ((size_t)s * cap_rows + b) * ldh + col + 4 * l16);
      if (concat) { o0 = va; o1 = vb; }
      else o0 = make_float4(va.x - vb.x, va.y - vb.y, va.z - vb.z, va.w - vb.w);
    }
    *reinterpret_cast<float4*>(support + row * width + 4 * l16) = o0;
    if (concat) *reinterpret_cast<float4*>(support + row * width + H + 4 * l16) = o1;
  }
}

extern "C" int tmpnn_aggregate_edges(const tmpnn_graph* g, const tmpnn_index* ix, const float* h, int ldh, int col,
                                     int concat, float* support, void* stream) {
  TMPNN_REQUIRE(g && h && support, "null argument");
  (void)ix;
  dim3 grid(min(tmpnn_div_up(g->cap_rows, 16), TMPNN_SM_COUNT * 8 / max(1, min(g->num_seqs, 8))), g->num_seqs);
  if (grid.x < 1) grid.x = 1;
  k_aggregate_edges<<<grid, 256, 0, (cudaStream_t)stream>>>(h, ldh, col, concat, g->n_rows, g->src, g->dst,
                                                           g->cap_rows, g->num_seqs, support);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// ------------------------------------------------------------------------------------------
// K2 + K3: fused GRU tile
// ------------------------------------------------------------------------------------------
template <int KX>
struct __align__(16) StepSmem {
  float wih[KX * 3 * H];
  float whh[H * 3 * H];
  float bias[4 * H];
  float headw[H];
  float headb[4];
  float X[TM][KX + 4];
  float Hp[TM][H + 4];
  int32_t rowid[TM];
};

template <int KX>
__device__ __forceinline__ void load_pack(StepSmem<KX>& s, const float* __restrict__ pack) {
  // the packed cell and the head of StepSmem share one layout -> one linear 128-bit copy
  float4* d = reinterpret_cast<float4*>(s.wih);
  const float4* p = reinterpret_cast<const float4*>(pack);
  for (int i = threadIdx.x; i < pack_floats(KX) / 4; i += NT) d[i] = __ldg(p + i);
}

// GEMMs + gates + head for the tile staged in s.X / s.Hp / s.rowid.
template <int KX>
__device__ __forceinline__ void gru_tile(StepSmem<KX>& s, float* __restrict__ h_out, int ldh, int col,
                                         float* __restrict__ logit, float* __restrict__ score, bool first_group,
                                         bool last_group, float* __restrict__ gates) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float ar[8][2], az[8][2], an[8][2], ahn[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ar[i][0] = ar[i][1] = az[i][0] = az[i][1] = an[i][0] = an[i][1] = ahn[i][0] = ahn[i][1] = 0.f;
  }
  // input part: r, z, i_n += x . W_ih^T
#pragma unroll 2
  for (int k = 0; k < KX; k += 4) {
    float4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(&s.X[ty * 8 + i][k]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float2 wr = *reinterpret_cast<const float2*>(&s.wih[(k + kk) * 3 * H + 2 * tx]);
      const float2 wz = *reinterpret_cast<const float2*>(&s.wih[(k + kk) * 3 * H + H + 2 * tx]);
      const float2 wn = *reinterpret_cast<const float2*>(&s.wih[(k + kk) * 3 * H + 2 * H + 2 * tx]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
        ar[i][0] = fmaf(x, wr.x, ar[i][0]); ar[i][1] = fmaf(x, wr.y, ar[i][1]);
        az[i][0] = fmaf(x, wz.x, az[i][0]); az[i][1] = fmaf(x, wz.y, az[i][1]);
        an[i][0] = fmaf(x, wn.x, an[i][0]); an[i][1] = fmaf(x, wn.y, an[i][1]);
      }
    }
  }
  // hidden part: r, z, h_n += h . W_hh^T
#pragma unroll 2
  for (int k = 0; k < H; k += 4) {
    float4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(&s.Hp[ty * 8 + i][k]);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float2 wr = *reinterpret_cast<const float2*>(&s.whh[(k + kk) * 3 * H + 2 * tx]);
      const float2 wz = *reinterpret_cast<const float2*>(&s.whh[(k + kk) * 3 * H + H + 2 * tx]);
      const float2 wn = *reinterpret_cast<const float2*>(&s.whh[(k + kk) * 3 * H + 2 * H + 2 * tx]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = kk == 0 ? xv[i].x : kk == 1 ? xv[i].y : kk == 2 ? xv[i].z : xv[i].w;
        ar[i][0] = fmaf(x, wr.x, ar[i][0]); ar[i][1] = fmaf(x, wr.y, ar[i][1]);
        az[i][0] = fmaf(x, wz.x, az[i][0]); az[i][1] = fmaf(x, wz.y, az[i][1]);
        ahn[i][0] = fmaf(x, wn.x, ahn[i][0]); ahn[i][1] = fmaf(x, wn.y, ahn[i][1]);
      }
    }
  }
  // epilogue: gates, h', head
  const float2 br = *reinterpret_cast<const float2*>(&s.bias[2 * tx]);
  const float2 bz = *reinterpret_cast<const float2*>(&s.bias[H + 2 * tx]);
  const float2 bin = *reinterpret_cast<const float2*>(&s.bias[2 * H + 2 * tx]);
  const float2 bhn = *reinterpret_cast<const float2*>(&s.bias[3 * H + 2 * tx]);
  const float2 hw = *reinterpret_cast<const float2*>(&s.headw[2 * tx]);
  const float hb = s.headb[0];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ty * 8 + i;
    const int row = s.rowid[r];  // warp-uniform
    const float2 hp = *reinterpret_cast<const float2*>(&s.Hp[r][2 * tx]);
    const float r0 = tmpnn_sigmoid(ar[i][0] + br.x), r1 = tmpnn_sigmoid(ar[i][1] + br.y);
    const float z0 = tmpnn_sigmoid(az[i][0] + bz.x), z1 = tmpnn_sigmoid(az[i][1] + bz.y);
    const float n0 = tanhf(an[i][0] + bin.x + r0 * (ahn[i][0] + bhn.x));
    const float n1 = tanhf(an[i][1] + bin.y + r1 * (ahn[i][1] + bhn.y));
    const float o0 = (1.0f - z0) * n0 + z0 * hp.x;
    const float o1 = (1.0f - z1) * n1 + z1 * hp.y;
    float dot = warp_sum_f(fmaf(o0, hw.x, o1 * hw.y));
    if (row >= 0) {
      *reinterpret_cast<float2*>(h_out + (size_t)row * ldh + col + 2 * tx) = make_float2(o0, o1);
      if (gates) {  // training: keep r | z | n | (W_hn h + b_hn) for the backward pass
        float* gr = gates + (size_t)row * 4 * H + 2 * tx;
        *reinterpret_cast<float2*>(gr) = make_float2(r0, r1);
        *reinterpret_cast<float2*>(gr + H) = make_float2(z0, z1);
        *reinterpret_cast<float2*>(gr + 2 * H) = make_float2(n0, n1);
        *reinterpret_cast<float2*>(gr + 3 * H) = make_float2(ahn[i][0] + bhn.x, ahn[i][1] + bhn.y);
      }
      if (tx == 0) {
        const float lg = dot + (first_group ? hb : logit[row]);
        logit[row] = lg;
        if (last_group) score[row] = tmpnn_sigmoid(lg);
      }
    }
  }
}

// Edge rows: tiles of 64 consecutive slab rows; detection rows inside a tile are masked.
template <int KX>
__global__ void __launch_bounds__(NT, 1)
k_mp_edge(const float* __restrict__ h_in, float* __restrict__ h_out, int ldh, int col,
          const int32_t* __restrict__ n_rows, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
          int cap_rows, int num_seqs, const int32_t* __restrict__ tile_ptr, const float* __restrict__ pack,
          float* __restrict__ logit, float* __restrict__ score, int first_group, int last_group,
          float* __restrict__ gates, const int32_t* __restrict__ phys, const int32_t* __restrict__ psrc,
          const int32_t* __restrict__ pdst, const int32_t* __restrict__ run_if_status, int run_if_mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<KX>& s = *reinterpret_cast<StepSmem<KX>*>(smem_raw);
  if (run_if_status && !(*run_if_status & run_if_mask)) return;  // conditional re-run (tmpnn_mp_edge_fwd_on_flag)
  const int total = tile_ptr[num_seqs];
  if ((int)blockIdx.x >= total) return;
  load_pack<KX>(s, pack);
  const int l16 = threadIdx.x & 15, rsub = threadIdx.x >> 4;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    // sequence of this tile: last s with tile_ptr[s] <= tile
    int lo = 0, hi = num_seqs;
    while (hi - lo > 1) {
      int mid = (lo + hi) >> 1;
      if (tile_ptr[mid] <= tile) lo = mid; else hi = mid;
    }
    const int seq = lo;
    const int r0 = (tile - tile_ptr[seq]) * TM;
    const int n = n_rows[seq];
    const size_t base = (size_t)seq * cap_rows;
    __syncthreads();  // previous tile's epilogue is done with X / Hp / rowid
#pragma unroll
    for (int p = 0; p < TM / 16; ++p) {
      const int r = rsub + 16 * p;
      const int lr = r0 + r;
      int a = -1, b = -1;
      if (lr < n) { a = src[base + lr]; b = dst[base + lr]; }
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va, vh = va;
      if (a >= 0) {
        // deferred compaction: the input state sits at physical (global) rows
        const size_t ra = phys ? (size_t)psrc[base + lr] : base + a, rb = phys ? (size_t)pdst[base + lr] : base + b;
        const size_t rh = phys ? (size_t)phys[base + lr] : base + lr;
        va = ldg4(h_in + ra * ldh + col + 4 * l16);
        vb = ldg4(h_in + rb * ldh + col + 4 * l16);
        vh = ldg4(h_in + rh * ldh + col + 4 * l16);
      }
      if (KX == 2 * H) {
        *reinterpret_cast<float4*>(&s.X[r][4 * l16]) = va;
        *reinterpret_cast<float4*>(&s.X[r][H + 4 * l16]) = vb;
      } else {
        *reinterpret_cast<float4*>(&s.X[r][4 * l16]) = make_float4(va.x - vb.x, va.y - vb.y, va.z - vb.z, va.w - vb.w);
      }
      *reinterpret_cast<float4*>(&s.Hp[r][4 * l16]) = vh;
      if (l16 == 0) s.rowid[r] = a >= 0 ? (int)(base + lr) : -1;
    }
    __syncthreads();
    gru_tile<KX>(s, h_out, ldh, col, logit, score, first_group != 0, last_group != 0, gates);
  }
}

// Detection rows: tiles of 64 entries of the detection list, x = agg.
__global__ void __launch_bounds__(NT, 1)
k_mp_det(const float* __restrict__ h_in, float* __restrict__ h_out, int ldh, int col,
         const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows, const float* __restrict__ agg,
         const float* __restrict__ pack, float* __restrict__ logit, float* __restrict__ score, int first_group,
         int last_group, float* __restrict__ gates, const int32_t* __restrict__ phys,
         const int32_t* __restrict__ run_if_status, int run_if_mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<H>& s = *reinterpret_cast<StepSmem<H>*>(smem_raw);
  if (run_if_status && !(*run_if_status & run_if_mask)) return;  // conditional re-run (tmpnn_mp_det_fwd_on_flag)
  const int nd = *n_dets;
  const int total = (nd + TM - 1) / TM;
  if ((int)blockIdx.x >= total) return;
  load_pack<H>(s, pack);
  const int l16 = threadIdx.x & 15, rsub = threadIdx.x >> 4;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int p = 0; p < TM / 16; ++p) {
      const int r = rsub + 16 * p;
      const int k = tile * TM + r;
      float4 vx = make_float4(0.f, 0.f, 0.f, 0.f), vh = vx;
      int row = -1;
      if (k < nd) {
        row = det_rows[k];
        vx = ldg4(agg + (size_t)k * H + 4 * l16);
        vh = ldg4(h_in + (size_t)(phys ? phys[row] : row) * ldh + col + 4 * l16);
      }
      *reinterpret_cast<float4*>(&s.X[r][4 * l16]) = vx;
      *reinterpret_cast<float4*>(&s.Hp[r][4 * l16]) = vh;
      if (l16 == 0) s.rowid[r] = row;
    }
    __syncthreads();
    gru_tile<H>(s, h_out, ldh, col, logit, score, first_group != 0, last_group != 0, gates);
  }
}

int tmpnn_init_graph_ops();  // graph_ops.cu
int tmpnn_init_tc();         // mp_step_tc.cu
int tmpnn_init_tc3();        // mp_step_tc3.cu
int tmpnn_init_train_tc();   // train_tc.cu
static bool g_init_done = false;

// Opts the big-shared-memory kernels in (once per process / device).  Called lazily by the
// entry points that need it; call it explicitly before capturing a CUDA graph.
extern "C" int tmpnn_init(void) {
  if (g_init_done) return TMPNN_OK;
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_edge<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem<H>)));
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_edge<2 * H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem<2 * H>)));
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_mp_det, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem<H>)));
  int rc = tmpnn_init_graph_ops();
  if (rc) return rc;
  rc = tmpnn_init_tc();
  if (rc) return rc;
  rc = tmpnn_init_tc3();
  if (rc) return rc;
  rc = tmpnn_init_train_tc();
  if (rc) return rc;
  g_init_done = true;
  return TMPNN_OK;
}

static int mp_edge_launch(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh, int group,
                          int num_groups, int concat, const float* edge_pack, float* gates, void* stream,
                          int run_if_mask = 0) {
  TMPNN_REQUIRE(g && ix && h_in && h_out && edge_pack, "null argument");
  TMPNN_REQUIRE(h_in != h_out, "h_in and h_out must be distinct buffers (Jacobi update)");
  TMPNN_REQUIRE(ldh % 4 == 0 && group >= 0 && group < num_groups && ldh >= num_groups * H, "bad ldh / group");
  if (!g_init_done) { int rc0 = tmpnn_init(); if (rc0) return rc0; }
  cudaStream_t st = (cudaStream_t)stream;
  const int col = group * H;
  const int first = group == 0, last = group == num_groups - 1;
  if (concat)
    k_mp_edge<2 * H><<<TMPNN_SM_COUNT, NT, sizeof(StepSmem<2 * H>), st>>>(
        h_in, h_out, ldh, col, g->n_rows, g->src, g->dst, g->cap_rows, g->num_seqs, ix->tile_ptr, edge_pack, g->logit,
        g->score, first, last, gates, g->phys, g->psrc, g->pdst, run_if_mask ? g->status : nullptr, run_if_mask);
  else
    k_mp_edge<H><<<TMPNN_SM_COUNT, NT, sizeof(StepSmem<H>), st>>>(
        h_in, h_out, ldh, col, g->n_rows, g->src, g->dst, g->cap_rows, g->num_seqs, ix->tile_ptr, edge_pack, g->logit,
        g->score, first, last, gates, g->phys, g->psrc, g->pdst, run_if_mask ? g->status : nullptr, run_if_mask);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

static int mp_det_launch(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh, int group,
                         int num_groups, const float* node_pack, const float* agg, float* gates, void* stream,
                         int run_if_mask = 0) {
  TMPNN_REQUIRE(g && ix && h_in && h_out && node_pack && agg, "null argument");
  TMPNN_REQUIRE(h_in != h_out, "h_in and h_out must be distinct buffers (Jacobi update)");
  TMPNN_REQUIRE(ldh % 4 == 0 && group >= 0 && group < num_groups && ldh >= num_groups * H, "bad ldh / group");
  if (!g_init_done) { int rc0 = tmpnn_init(); if (rc0) return rc0; }
  k_mp_det<<<TMPNN_SM_COUNT, NT, sizeof(StepSmem<H>), (cudaStream_t)stream>>>(
      h_in, h_out, ldh, group * H, ix->n_dets, ix->det_rows, agg, node_pack, g->logit, g->score, group == 0,
      group == num_groups - 1, gates, g->phys, run_if_mask ? g->status : nullptr, run_if_mask);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_mp_edge_fwd(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                 int group, int num_groups, int concat, const float* edge_pack, void* stream) {
  return mp_edge_launch(g, ix, h_in, h_out, ldh, group, num_groups, concat, edge_pack, nullptr, stream);
}

extern "C" int tmpnn_mp_edge_fwd_on_flag(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                         int group, int num_groups, int concat, const float* edge_pack, int flag_mask,
                                         void* stream) {
  TMPNN_REQUIRE(flag_mask != 0, "flag_mask must name at least one status bit");
  return mp_edge_launch(g, ix, h_in, h_out, ldh, group, num_groups, concat, edge_pack, nullptr, stream, flag_mask);
}

extern "C" int tmpnn_mp_det_fwd(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                int group, int num_groups, const float* node_pack, const float* agg, void* stream) {
  return mp_det_launch(g, ix, h_in, h_out, ldh, group, num_groups, node_pack, agg, nullptr, stream);
}

extern "C" int tmpnn_mp_det_fwd_on_flag(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                        int group, int num_groups, const float* node_pack, const float* agg, int flag_mask,
                                        void* stream) {
  TMPNN_REQUIRE(flag_mask != 0, "flag_mask must name at least one status bit");
  return mp_det_launch(g, ix, h_in, h_out, ldh, group, num_groups, node_pack, agg, nullptr, stream, flag_mask);
}

extern "C" int tmpnn_mp_det_fwd_train(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                      int group, int num_groups, const float* node_pack, const float* agg, float* gates,
                                      void* stream) {
  TMPNN_REQUIRE(gates, "null argument");
  return mp_det_launch(g, ix, h_in, h_out, ldh, group, num_groups, node_pack, agg, gates, stream);
}

extern "C" int tmpnn_mp_step_fwd_train(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                       int group, int num_groups, int concat, const float* edge_pack,
                                       const float* node_pack, float* agg, float* gates, void* stream) {
  TMPNN_REQUIRE(gates, "null argument");
  int rc = tmpnn_aggregate_dets(g, ix, h_in, ldh, group * H, agg, stream);
  if (rc) return rc;
  rc = mp_edge_launch(g, ix, h_in, h_out, ldh, group, num_groups, concat, edge_pack, gates, stream);
  if (rc) return rc;
  return mp_det_launch(g, ix, h_in, h_out, ldh, group, num_groups, node_pack, agg, gates, stream);
}

extern "C" int tmpnn_mp_step_fwd_train_agg(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                           int group, int num_groups, int concat, const float* edge_pack,
                                           const float* node_pack, const float* agg, float* gates, void* stream) {
  TMPNN_REQUIRE(gates && agg, "null argument");
  int rc = mp_edge_launch(g, ix, h_in, h_out, ldh, group, num_groups, concat, edge_pack, gates, stream);
  if (rc) return rc;
  return mp_det_launch(g, ix, h_in, h_out, ldh, group, num_groups, node_pack, agg, gates, stream);
}

extern "C" int tmpnn_mp_step_fwd(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh,
                                 int group, int num_groups, int concat, const float* edge_pack, const float* node_pack,
                                 float* agg, void* stream) {
  TMPNN_REQUIRE(g && ix && h_in && h_out && edge_pack && node_pack && agg, "null argument");
  int rc = tmpnn_aggregate_dets(g, ix, h_in, ldh, group * H, agg, stream);
  if (rc) return rc;
  rc = tmpnn_mp_edge_fwd(g, ix, h_in, h_out, ldh, group, num_groups, concat, edge_pack, stream);
  if (rc) return rc;
  return tmpnn_mp_det_fwd(g, ix, h_in, h_out, ldh, group, num_groups, node_pack, agg, stream);
}
