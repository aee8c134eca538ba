// train.cu -- backward pass of the message-passing step and the training losses (fp32).
//
// Reference behaviour: the autograd graph torch builds for models/track_mpnn.py:54-75 +
// models/layers.py:84-116 when train.py:65-134 calls loss.backward(), and models/loss.py:8-115
// (create_targets, CELoss, FocalLoss with gamma = 0).  Restated on the edge list:
//
//   forward (per feature group, Jacobi):   x_e = h[src]-h[dst] | [h[src] | h[dst]],  x_d = agg[d]
//                                          h' = GRUCell_type(x, h),  logit = w_type . h' + b_type
//   backward, given dL/dh', dL/dlogit, dL/dscore:
//     dh'   += dlogit_tot * w_type                         dlogit_tot = dlogit + dscore * p (1 - p)
//     dn = dh' (1-z), dz = dh' (h - n), dh_self = dh' z
//     dpn = dn (1 - n^2), dpz = dz z (1-z), dpr = dpn * hn_pre * r (1-r)
//     dgi = [dpr, dpz, dpn], dgh = [dpr, dpz, dpn r]
//     dx = dgi . W_ih,  dh_self += dgh . W_hh,  dW_ih += dgi^T x,  dW_hh += dgh^T h,  db_* += sum dg*
//   and the transpose of the gather / segmented sum in GATHER form (no float atomics on the state):
//     edge row e: dh_in[e] = dh_self[e] + dagg[src(e)] - dagg[dst(e)]
//     det  row d: dh_in[d] = dh_self[d] + sum_{e: src=d} dx_e[0:64] -/+ sum_{e: dst=d} dx_e[0:64 | 64:128]
//
// The training graphs of the reference are small (<= 10^4 rows per step, BPTT over <= 10 steps), so
// these kernels favour simple, checkable structure; weight-gradient partial sums are combined with
// fp32 atomics (order-dependent in the last bits, well inside the gradient tolerance).
#include "common.cuh"

namespace {

constexpr int H = TMPNN_HIDDEN;

// ------------------------------------------------------------------------------------------
// gate gradients: 16 lanes x float4 per row (128-bit loads / stores: the kernel moves ~3.5 KB per row and nothing else),
// block = 16 rows
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_acc(float4& a, const float4 v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }

constexpr int GB_PART = 5 * H + 4;   // per CTA and row type: five 64-wide sums + the head-bias sum (+ pad)
__global__ void __launch_bounds__(256)
k_gate_bwd(int n_rows, const int32_t* __restrict__ src, const float* __restrict__ gates,
           const float* __restrict__ h_prev, const float* __restrict__ h_new, int ldh, int col,
           const float* __restrict__ dh_out, const float* __restrict__ dlogits, const float* __restrict__ dscores,
           const float* __restrict__ score, const float* __restrict__ hw_e, const float* __restrict__ hw_d,
           float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dhself,
           float* __restrict__ gb_e, float* __restrict__ gb_d,   // [2][192]: d bias_ih | d bias_hh
           float* __restrict__ ghw_e, float* __restrict__ ghw_d, // [64] head weight slices
           float* __restrict__ ghb_e, float* __restrict__ ghb_d, // [1] head biases (group 0 only, else null)
           float* __restrict__ partials)  // null: the CTAs add atomically; else [grid][2][GB_PART]: summed by k_gate_bwd_reduce
{
  __shared__ float4 red[16][16];
  __shared__ float redb[16];
  const int l = threadIdx.x & 15, rl = threadIdx.x >> 4;
  float4 acc[2][5];  // per row type: dpr, dpz, dpn, dpnr, dl * h'
#pragma unroll
  for (int t = 0; t < 2; ++t)
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[t][q] = f4_zero();
  float accb[2] = {0.f, 0.f};
  const float4 we = ldg4(hw_e + 4 * l), wd = ldg4(hw_d + 4 * l);
  for (int row = blockIdx.x * 16 + rl; row < n_rows; row += gridDim.x * 16) {
    const int t = __ldg(src + row) >= 0 ? 0 : 1;  // 0 = edge row, 1 = detection row
    const float p = __ldg(score + row);
    const float dl = (dlogits ? __ldg(dlogits + row) : 0.f) + (dscores ? __ldg(dscores + row) * p * (1.0f - p) : 0.f);
    const float* gr = gates + (size_t)row * 4 * H + 4 * l;
    const float4 r = ldg4(gr), z = ldg4(gr + H), n = ldg4(gr + 2 * H), hn = ldg4(gr + 3 * H);
    const size_t o = (size_t)row * ldh + col + 4 * l;
    const float4 h = ldg4(h_prev + o), hnew = ldg4(h_new + o);
    const float4 dho = dh_out ? ldg4(dh_out + o) : f4_zero();
    const float4 w = t ? wd : we;
    float4 dpr, dpz, dpn, dpnr, dhz;
#define TMPNN_GATE1(c)                                                             \
    {                                                                              \
      const float dh = dho.c + dl * w.c;                                           \
      const float dn = dh * (1.0f - z.c), dz = dh * (h.c - n.c);                   \
      dpn.c = dn * (1.0f - n.c * n.c);                                             \
      dpz.c = dz * z.c * (1.0f - z.c);                                             \
      dpr.c = dpn.c * hn.c * r.c * (1.0f - r.c);                                   \
      dpnr.c = dpn.c * r.c;                                                        \
      dhz.c = dh * z.c;                                                            \
    }
    TMPNN_GATE1(x) TMPNN_GATE1(y) TMPNN_GATE1(z) TMPNN_GATE1(w)
#undef TMPNN_GATE1
    float* gi = dgi + (size_t)row * 3 * H + 4 * l;
    float* gh = dgh + (size_t)row * 3 * H + 4 * l;
    *reinterpret_cast<float4*>(gi) = dpr; *reinterpret_cast<float4*>(gi + H) = dpz; *reinterpret_cast<float4*>(gi + 2 * H) = dpn;
    *reinterpret_cast<float4*>(gh) = dpr; *reinterpret_cast<float4*>(gh + H) = dpz; *reinterpret_cast<float4*>(gh + 2 * H) = dpnr;
    *reinterpret_cast<float4*>(dhself + (size_t)row * H + 4 * l) = dhz;
    f4_acc(acc[t][0], dpr); f4_acc(acc[t][1], dpz); f4_acc(acc[t][2], dpn); f4_acc(acc[t][3], dpnr);
    f4_acc(acc[t][4], make_float4(dl * hnew.x, dl * hnew.y, dl * hnew.z, dl * hnew.w));
    if (l == 0) accb[t] += dl;
  }
  for (int t = 0; t < 2; ++t) {
    float* gb = t ? gb_d : gb_e;
    float* ghw = t ? ghw_d : ghw_e;
    for (int q = 0; q < 5; ++q) {
      __syncthreads();
      red[rl][l] = acc[t][q];
      __syncthreads();
      if (threadIdx.x < 64) {   // thread j sums hidden unit j over the 16 row lanes (fixed order)
        const int j = threadIdx.x;
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) v += reinterpret_cast<const float*>(&red[k][j >> 2])[j & 3];
        if (partials) {
          partials[((size_t)blockIdx.x * 2 + t) * GB_PART + q * H + j] = v;
        } else if (v != 0.f) {
          if (q < 3) atomicAdd(&gb[q * H + j], v);                  // d bias_ih
          if (q < 2) atomicAdd(&gb[3 * H + q * H + j], v);          // d bias_hh (r, z)
          if (q == 3) atomicAdd(&gb[3 * H + 2 * H + j], v);         // d bias_hh (n)
          if (q == 4) atomicAdd(&ghw[j], v);
        }
      }
    }
    __syncthreads();
    if (l == 0) redb[rl] = accb[t];
    __syncthreads();
    float* ghb = t ? ghb_d : ghb_e;
    if (threadIdx.x == 0 && (ghb || partials)) {
      float v = 0.f;
      for (int k = 0; k < 16; ++k) v += redb[k];
      if (partials) partials[((size_t)blockIdx.x * 2 + t) * GB_PART + 5 * H] = v;
      else if (v != 0.f) atomicAdd(ghb, v);
    }
  }
}
// The CTAs' partial sums of k_gate_bwd added in CTA order (16 groups of CTAs per output, combined in a fixed order): the bias and
// head-weight gradients come out bit-reproducible, which the float atomics could not give.  Outputs o < 640: (row type t, sum q,
// hidden unit j); o = 640, 641: the two head biases.
__global__ void __launch_bounds__(1024)
k_gate_bwd_reduce(const float* __restrict__ partials, int n_blocks, float* __restrict__ gb_e, float* __restrict__ gb_d,
                  float* __restrict__ ghw_e, float* __restrict__ ghw_d, float* __restrict__ ghb_e, float* __restrict__ ghb_d) {
  __shared__ float part[16][64];
  const int il = threadIdx.x & 63, pg = threadIdx.x >> 6;
  const int o = blockIdx.x * 64 + il;
  const bool bias = o >= 2 * 5 * H;
  const int t = bias ? o - 2 * 5 * H : o / (5 * H), rem = bias ? 5 * H : o % (5 * H);
  float s = 0.f;
  if (t < 2) {
    const float* p = partials + (size_t)t * GB_PART + rem;
    for (int b0 = pg; b0 < n_blocks; b0 += 16 * 8) {   // eight loads in flight, added in CTA order
      float v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int b = b0 + 16 * q;
        v[q] = b < n_blocks ? p[(size_t)b * 2 * GB_PART] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) s += v[q];
    }
  }
  part[pg][il] = s;
  __syncthreads();
  if (pg != 0 || t >= 2) return;
  float v = part[0][il];
#pragma unroll
  for (int k = 1; k < 16; ++k) v += part[k][il];
  if (bias) {
    float* ghb = t ? ghb_d : ghb_e;
    if (ghb) *ghb += v;
    return;
  }
  const int q = rem / H, j = rem % H;
  float* gb = t ? gb_d : gb_e;
  float* ghw = t ? ghw_d : ghw_e;
  if (q < 3) gb[q * H + j] += v;                  // d bias_ih
  if (q < 2) gb[3 * H + q * H + j] += v;          // d bias_hh (r, z)
  if (q == 3) gb[3 * H + 2 * H + j] += v;         // d bias_hh (n)
  if (q == 4) ghw[j] += v;
}

// ------------------------------------------------------------------------------------------
// C[rc(i)][0:N] (+)= A[ra(i)][0:K] . W[K][N]      rows i < R with mask[ra(i)] >= 0
// ------------------------------------------------------------------------------------------
constexpr int GK = 192;
// Register tiled: 32 rows per pass, thread (tr, tn) owns rows 2 tr, 2 tr + 1 and N / 16 consecutive columns; per k it
// reads 2 staged A values (warp-uniform per row pair) and one 128-bit slice of W (L1 resident) for 2 N / 16 FMAs
// (the first version: one column per thread, 4 shared loads + 1 global load per 4 FMAs).
template <int N>
__global__ void __launch_bounds__(256)
k_rows_times_w(const int32_t* __restrict__ r_dev, int r_host, const int32_t* __restrict__ a_rows,
               const int32_t* __restrict__ c_rows, const int32_t* __restrict__ mask, const float* __restrict__ A,
               const float* __restrict__ W, float* __restrict__ C, int ldc, int accumulate) {
  constexpr int RPB = 32;     // rows per block pass
  constexpr int RT = 2;       // rows per thread
  constexpr int TN = N / 16;  // 4 or 8 consecutive columns per thread
  __shared__ __align__(16) float As[RPB][GK + 4];
  __shared__ int32_t rc_s[RPB];
  const int R = r_dev ? *r_dev : r_host;
  const int tn = threadIdx.x & 15, tr = threadIdx.x >> 4;
  for (int i0 = blockIdx.x * RPB; i0 < R; i0 += gridDim.x * RPB) {
    __syncthreads();
    for (int q = threadIdx.x; q < RPB * GK / 4; q += 256) {
      const int ri = q / (GK / 4), k4 = q % (GK / 4), i = i0 + ri;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < R) {
        const int ra = a_rows ? a_rows[i] : i;
        if (!mask || mask[ra] >= 0) v = ldg4(A + (size_t)ra * GK + 4 * k4);
      }
      *reinterpret_cast<float4*>(&As[ri][4 * k4]) = v;
    }
    if (threadIdx.x < RPB) {
      const int i = i0 + threadIdx.x;
      int rc = -1;
      if (i < R) {
        const int ra = a_rows ? a_rows[i] : i;
        if (!mask || mask[ra] >= 0) rc = c_rows ? c_rows[i] : i;
      }
      rc_s[threadIdx.x] = rc;
    }
    __syncthreads();
    float acc[RT][TN];
#pragma unroll
    for (int u = 0; u < RT; ++u)
#pragma unroll
      for (int v = 0; v < TN; ++v) acc[u][v] = 0.f;
#pragma unroll 8
    for (int k = 0; k < GK; ++k) {
      float wv[TN];
#pragma unroll
      for (int v = 0; v < TN / 4; ++v) *reinterpret_cast<float4*>(&wv[4 * v]) = ldg4(W + (size_t)k * N + TN * tn + 4 * v);
#pragma unroll
      for (int u = 0; u < RT; ++u) {
        const float a = As[RT * tr + u][k];
#pragma unroll
        for (int v = 0; v < TN; ++v) acc[u][v] = fmaf(a, wv[v], acc[u][v]);
      }
    }
#pragma unroll
    for (int u = 0; u < RT; ++u) {
      const int rc = rc_s[RT * tr + u];
      if (rc >= 0) {
        float4* c = reinterpret_cast<float4*>(C + (size_t)rc * ldc + TN * tn);
#pragma unroll
        for (int v = 0; v < TN / 4; ++v) {
          float4 o = make_float4(acc[u][4 * v], acc[u][4 * v + 1], acc[u][4 * v + 2], acc[u][4 * v + 3]);
          if (accumulate) {
            const float4 p = c[v];
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          c[v] = o;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// G[m][n] += sum_i A[ra(i)][m] * B[rb(i)][n]      m < 192, n < N, rows i < R with mask[ra(i)] >= 0
// ------------------------------------------------------------------------------------------
// Register tiled: the 192 x N outputs are cut into 16 x 16 thread tiles of 12 x (N / 16) outputs; per staged row a
// thread reads 12 + N / 16 values as 128-bit shared loads (warp-uniform or contiguous -> conflict free) for
// 12 N / 16 FMAs, instead of two shared loads per FMA (the first version: 7.6 TFLOP/s, 35 % of a batched step).
template <int N>
__global__ void __launch_bounds__(256)
k_rows_outer(const int32_t* __restrict__ r_dev, int r_host, const int32_t* __restrict__ a_rows,
             const int32_t* __restrict__ b_rows, const int32_t* __restrict__ mask, const float* __restrict__ A,
             const float* __restrict__ B, int ldb, float* __restrict__ G) {
  constexpr int RPB = 32;
  constexpr int TN = N / 16;  // 4 or 8 consecutive n per thread
  __shared__ __align__(16) float As[RPB][GK];
  __shared__ __align__(16) float Bs[RPB][N];
  const int R = r_dev ? *r_dev : r_host;
  const int tn = threadIdx.x & 15, tm = threadIdx.x >> 4;  // outputs (12 tm + u, TN tn + v)
  float acc[12][TN];
#pragma unroll
  for (int u = 0; u < 12; ++u)
#pragma unroll
    for (int v = 0; v < TN; ++v) acc[u][v] = 0.f;
  for (int i0 = blockIdx.x * RPB; i0 < R; i0 += gridDim.x * RPB) {
    __syncthreads();
    for (int q = threadIdx.x; q < RPB * GK / 4; q += 256) {
      const int ri = q / (GK / 4), k4 = q % (GK / 4), i = i0 + ri;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < R) {
        const int ra = a_rows ? a_rows[i] : i;
        if (!mask || mask[ra] >= 0) v = ldg4(A + (size_t)ra * GK + 4 * k4);
      }
      *reinterpret_cast<float4*>(&As[ri][4 * k4]) = v;
    }
    for (int q = threadIdx.x; q < RPB * N / 4; q += 256) {
      const int ri = q / (N / 4), k4 = q % (N / 4), i = i0 + ri;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < R) v = ldg4(B + (size_t)(b_rows ? b_rows[i] : i) * ldb + 4 * k4);
      *reinterpret_cast<float4*>(&Bs[ri][4 * k4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int ri = 0; ri < RPB; ++ri) {
      float av[12], bv[TN];
#pragma unroll
      for (int u = 0; u < 3; ++u) *reinterpret_cast<float4*>(&av[4 * u]) = *reinterpret_cast<const float4*>(&As[ri][12 * tm + 4 * u]);
#pragma unroll
      for (int v = 0; v < TN / 4; ++v) *reinterpret_cast<float4*>(&bv[4 * v]) = *reinterpret_cast<const float4*>(&Bs[ri][TN * tn + 4 * v]);
#pragma unroll
      for (int u = 0; u < 12; ++u)
#pragma unroll
        for (int v = 0; v < TN; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 12; ++u)
#pragma unroll
    for (int v = 0; v < TN; ++v)
      if (acc[u][v] != 0.f) atomicAdd(&G[(12 * tm + u) * N + TN * tn + v], acc[u][v]);
}

// ------------------------------------------------------------------------------------------
// transpose of the gather / segmented sum
// ------------------------------------------------------------------------------------------
// edge rows: half-warp per row
__global__ void __launch_bounds__(256)
k_scatter_bwd_edges(int n_rows, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                    const int32_t* __restrict__ det_of_row, const float* __restrict__ dhself,
                    const float* __restrict__ dagg, float* __restrict__ dh_in, int ldh, int col) {
  const int l16 = threadIdx.x & 15;
  for (int row = blockIdx.x * 16 + (threadIdx.x >> 4); row < n_rows; row += gridDim.x * 16) {
    const int a = src[row];
    if (a < 0) continue;
    const int ka = det_of_row[a], kb = det_of_row[dst[row]];
    const float4 s = ldg4(dhself + (size_t)row * H + 4 * l16);
    const float4 pa = ldg4(dagg + (size_t)ka * H + 4 * l16), pb = ldg4(dagg + (size_t)kb * H + 4 * l16);
    *reinterpret_cast<float4*>(dh_in + (size_t)row * ldh + col + 4 * l16) =
        make_float4(s.x + pa.x - pb.x, s.y + pa.y - pb.y, s.z + pa.z - pb.z, s.w + pa.w - pb.w);
  }
}

// detection rows: warp per detection over its incidence segments (ascending -> reproducible)
__global__ void __launch_bounds__(256)
k_scatter_bwd_dets(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows,
                   const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ inc,
                   const float* __restrict__ dhself, const float* __restrict__ dx, int kx,
                   float* __restrict__ dh_in, int ldh, int col) {
  const int nd = *n_dets;
  const int lane = threadIdx.x & 31;
  const int concat = kx == 2 * H;
  for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < nd; k += gridDim.x * 8) {
    const int s0 = seg_ptr[2 * k], s1 = seg_ptr[2 * k + 1], s2 = seg_ptr[2 * k + 2];
    const int row = det_rows[k];
    float a0 = dhself[(size_t)row * H + lane], a1 = dhself[(size_t)row * H + lane + 32];
    // four rows of dx in flight per step (the loop is bound by load latency); the additions keep the list order, so the sum
    // is the same bit for bit as one row at a time
    const int po = concat ? H : 0;
    const float ps = concat ? 1.f : -1.f;
    int i = s0;
    for (; i + 3 < s1; i += 4) {  // past edges: this detection is their dst
      const float* d0 = dx + (size_t)inc[i] * kx + po; const float* d1 = dx + (size_t)inc[i + 1] * kx + po;
      const float* d2 = dx + (size_t)inc[i + 2] * kx + po; const float* d3 = dx + (size_t)inc[i + 3] * kx + po;
      const float u0 = d0[lane], v0 = d0[lane + 32], u1 = d1[lane], v1 = d1[lane + 32];
      const float u2 = d2[lane], v2 = d2[lane + 32], u3 = d3[lane], v3 = d3[lane + 32];
      a0 = fmaf(ps, u0, a0); a1 = fmaf(ps, v0, a1); a0 = fmaf(ps, u1, a0); a1 = fmaf(ps, v1, a1);
      a0 = fmaf(ps, u2, a0); a1 = fmaf(ps, v2, a1); a0 = fmaf(ps, u3, a0); a1 = fmaf(ps, v3, a1);
    }
    for (; i < s1; ++i) {
      const float* d = dx + (size_t)inc[i] * kx + po;
      a0 = fmaf(ps, d[lane], a0); a1 = fmaf(ps, d[lane + 32], a1);
    }
    for (; i + 3 < s2; i += 4) {  // future edges: this detection is their src
      const float* d0 = dx + (size_t)inc[i] * kx; const float* d1 = dx + (size_t)inc[i + 1] * kx;
      const float* d2 = dx + (size_t)inc[i + 2] * kx; const float* d3 = dx + (size_t)inc[i + 3] * kx;
      const float u0 = d0[lane], v0 = d0[lane + 32], u1 = d1[lane], v1 = d1[lane + 32];
      const float u2 = d2[lane], v2 = d2[lane + 32], u3 = d3[lane], v3 = d3[lane + 32];
      a0 += u0; a1 += v0; a0 += u1; a1 += v1; a0 += u2; a1 += v2; a0 += u3; a1 += v3;
    }
    for (; i < s2; ++i) {
      const float* d = dx + (size_t)inc[i] * kx;
      a0 += d[lane]; a1 += d[lane + 32];
    }
    dh_in[(size_t)row * ldh + col + lane] = a0;
    dh_in[(size_t)row * ldh + col + lane + 32] = a1;
  }
}

// ------------------------------------------------------------------------------------------
// input transform backward (Linear -> BatchNorm -> ReLU -> Linear), one CTA
// ------------------------------------------------------------------------------------------
constexpr int GROUPS_PER_LAUNCH = 32;  // descriptors travel by value in the kernel parameters (32 x 56 B)
struct InputGroupBatch { tmpnn_input_group g[GROUPS_PER_LAUNCH]; };
// ATOMIC 0: one CTA owns the gradient buffers (+=); 1: several CTAs (one per group of rows = one BatchNorm batch) add into
// the same buffers atomically; 2: every CTA stores into its own partial buffer (k_input_bwd_reduce adds them in group order)
template <int ATOMIC>
__device__ __forceinline__ void grad_add(float* p, float v) {
  if (ATOMIC == 1) atomicAdd(p, v);
  else if (ATOMIC == 2) *p = v;
  else *p += v;
}
template <int ATOMIC>
__device__ __forceinline__ void
input_bwd_body(const float* __restrict__ x, int ldx, int col0, int f_in, const int32_t* __restrict__ x_idx,
               const float* __restrict__ a, const float* __restrict__ mean, const float* __restrict__ var,
               const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ b1,
               const float* __restrict__ w2, const float* __restrict__ dh, int ldh, int col,
               const int32_t* __restrict__ out_rows, int n, int n_edge, int training, float* __restrict__ scratch,
               float* __restrict__ gw1, float* __restrict__ gb1, float* __restrict__ ggamma, float* __restrict__ gbeta,
               float* __restrict__ gw2, float* __restrict__ gb2) {
  __shared__ float red[2][4][H];
  __shared__ float m_s[2][H];
  const int j = threadIdx.x & 63, rl = threadIdx.x >> 6;
  float* dbn = scratch;                   // [n][64], later da
  float* act = scratch + (size_t)n * H;   // [n][64]
  const float inv = 1.0f / sqrtf(var[j] + 1e-5f), mu = mean[j], gm = gamma[j], bt = beta[j];
  const float ntot = (float)(n + n_edge);
  // pass 1: d act = dh . W2, through the ReLU; channel sums
  float s1 = 0.f, s2 = 0.f, sb2 = 0.f;
  for (int i = rl; i < n; i += 4) {
    const float* dr = dh + (size_t)out_rows[i] * ldh + col;
    float d = 0.f;
    for (int o = 0; o < H; ++o) d = fmaf(dr[o], w2[o * H + j], d);
    const float xh = (a[(size_t)i * H + j] - mu) * inv;
    const float bn = xh * gm + bt;
    const float g = bn > 0.f ? d : 0.f;
    dbn[(size_t)i * H + j] = g;
    act[(size_t)i * H + j] = fmaxf(bn, 0.f);
    s1 += g; s2 += g * xh; sb2 += dr[j];
  }
  red[0][rl][j] = s1; red[1][rl][j] = s2;
  __syncthreads();
  if (rl == 0) {
    const float t1 = red[0][0][j] + red[0][1][j] + red[0][2][j] + red[0][3][j];
    const float t2 = red[1][0][j] + red[1][1][j] + red[1][2][j] + red[1][3][j];
    m_s[0][j] = t1; m_s[1][j] = t2;
    grad_add<ATOMIC>(&gbeta[j], t1);
    grad_add<ATOMIC>(&ggamma[j], t2);
  }
  __syncthreads();
  red[0][rl][j] = sb2;
  __syncthreads();
  if (rl == 0) grad_add<ATOMIC>(&gb2[j], red[0][0][j] + red[0][1][j] + red[0][2][j] + red[0][3][j]);
  const float m1 = training ? m_s[0][j] / ntot : 0.f, m2 = training ? m_s[1][j] / ntot : 0.f;
  __syncthreads();
  // d W2[o][j] += sum_i dh[i][o] act[i][j]: thread (rl, j) owns o = rl + 4 m; the rows are staged 32 at a time in shared
  // memory (the global-memory form spent most of the kernel in this loop: 2 dependent-address loads per FMA); same
  // ascending order over i as before
  {
    __shared__ float sdh[32][H], sact[32][H];
    float acc[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) acc[m] = 0.f;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int nb = min(32, n - i0);
      __syncthreads();
      for (int e = threadIdx.x; e < nb * H; e += 256) {
        const int r = e >> 6, c = e & 63;
        sdh[r][c] = dh[(size_t)out_rows[i0 + r] * ldh + col + c];
        sact[r][c] = act[(size_t)(i0 + r) * H + c];
      }
      __syncthreads();
      for (int i = 0; i < nb; ++i) {
        const float av = sact[i][j];
#pragma unroll
        for (int m = 0; m < 16; ++m) acc[m] = fmaf(sdh[i][rl + 4 * m], av, acc[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) grad_add<ATOMIC>(&gw2[(rl + 4 * m) * H + j], acc[m]);
  }
  // pass 2: through the BatchNorm
  float sda = 0.f;
  for (int i = rl; i < n; i += 4) {
    const float xh = (a[(size_t)i * H + j] - mu) * inv;
    const float da = gm * inv * (dbn[(size_t)i * H + j] - m1 - xh * m2);
    dbn[(size_t)i * H + j] = da;
    sda += da;
  }
  __syncthreads();
  red[0][rl][j] = sda;
  __syncthreads();
  if (rl == 0) {
    float t = red[0][0][j] + red[0][1][j] + red[0][2][j] + red[0][3][j];
    if (training) {  // the n_edge all-zero rows (value b1 after Linear1) sit in the batch statistics too
      const float xe = (b1[j] - mu) * inv;
      t += (float)n_edge * gm * inv * (-m1 - xe * m2);
    }
    grad_add<ATOMIC>(&gb1[j], t);
  }
  __syncthreads();
  // d W1[j][k] += sum_i da[i][j] x[i][k]
  for (int q = threadIdx.x; q < H * f_in; q += 256) {
    const int jj = q / f_in, k = q % f_in;
    float s = 0.f;
    for (int i = 0; i < n; ++i) s = fmaf(dbn[(size_t)i * H + jj], x[(size_t)(x_idx ? x_idx[i] : i) * ldx + col0 + k], s);
    grad_add<ATOMIC>(&gw1[q], s);
  }
}

__global__ void __launch_bounds__(256)
k_input_bwd(const float* __restrict__ x, int ldx, int col0, int f_in, const int32_t* __restrict__ x_idx,
            const float* __restrict__ a, const float* __restrict__ mean, const float* __restrict__ var,
            const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ b1,
            const float* __restrict__ w2, const float* __restrict__ dh, int ldh, int col,
            const int32_t* __restrict__ out_rows, int n, int n_edge, int training, float* __restrict__ scratch,
            float* __restrict__ gw1, float* __restrict__ gb1, float* __restrict__ ggamma, float* __restrict__ gbeta,
            float* __restrict__ gw2, float* __restrict__ gb2) {
  input_bwd_body<0>(x, ldx, col0, f_in, x_idx, a, mean, var, gamma, beta, b1, w2, dh, ldh, col, out_rows, n, n_edge,
                        training, scratch, gw1, gb1, ggamma, gbeta, gw2, gb2);
}
// one CTA per group (tmpnn_input_group): the batched trainer's chunks, each its own BatchNorm batch
__global__ void __launch_bounds__(256)
k_input_bwd_groups(const float* __restrict__ x, int ldx, int col0, int f_in, const InputGroupBatch groups,
                   const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ b1,
                   const float* __restrict__ w2, const float* __restrict__ dh, int ldh, int col, int training,
                   float* __restrict__ gw1, float* __restrict__ gb1, float* __restrict__ ggamma, float* __restrict__ gbeta,
                   float* __restrict__ gw2, float* __restrict__ gb2) {
  const tmpnn_input_group& g = groups.g[blockIdx.x];
  if (g.n <= 0) return;
  input_bwd_body<1>(x, ldx, col0, f_in, g.x_idx, g.a, g.mean, g.var, gamma, beta, b1, w2, dh, ldh, col, g.out_rows, g.n,
                    g.n_edge_rows, training, g.scratch, gw1, gb1, ggamma, gbeta, gw2, gb2);
}
// The same with per-group partial gradients instead of atomics: 32 CTAs adding 4 k values each into the same addresses
// spent most of the kernel in the L2 atomic units, and the sums came out in arbitrary order.  Partial of group k (floats):
// [0, 4096) d W2 | [4096, 8192) d W1 (64 x f_in) | then 64 each of d b1, d gamma, d beta, d b2.
constexpr int IBW_PART = 2 * H * H + 4 * H;
__global__ void __launch_bounds__(256)
k_input_bwd_groups_part(const float* __restrict__ x, int ldx, int col0, int f_in, const InputGroupBatch groups,
                        const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ b1,
                        const float* __restrict__ w2, const float* __restrict__ dh, int ldh, int col, int training,
                        float* __restrict__ partials) {
  const tmpnn_input_group& g = groups.g[blockIdx.x];
  if (g.n <= 0) return;
  float* p = partials + (size_t)blockIdx.x * IBW_PART;
  input_bwd_body<2>(x, ldx, col0, f_in, g.x_idx, g.a, g.mean, g.var, gamma, beta, b1, w2, dh, ldh, col, g.out_rows, g.n,
                    g.n_edge_rows, training, g.scratch, p + H * H, p + 2 * H * H, p + 2 * H * H + H, p + 2 * H * H + 2 * H, p,
                    p + 2 * H * H + 3 * H);
}
__global__ void __launch_bounds__(256)
k_input_bwd_reduce(const InputGroupBatch groups, int n_groups, int f_in, const float* __restrict__ partials,
                   float* __restrict__ gw1, float* __restrict__ gb1, float* __restrict__ ggamma, float* __restrict__ gbeta,
                   float* __restrict__ gw2, float* __restrict__ gb2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= IBW_PART) return;
  float* dst;
  if (i < H * H) dst = gw2 + i;
  else if (i < 2 * H * H) { if (i - H * H >= H * f_in) return; dst = gw1 + (i - H * H); }
  else {
    const int q = (i - 2 * H * H) / H, j = (i - 2 * H * H) % H;
    dst = (q == 0 ? gb1 : q == 1 ? ggamma : q == 2 ? gbeta : gb2) + j;
  }
  float v[GROUPS_PER_LAUNCH];
#pragma unroll
  for (int k = 0; k < GROUPS_PER_LAUNCH; ++k) v[k] = (k < n_groups && groups.g[k].n > 0) ? partials[(size_t)k * IBW_PART + i] : 0.f;
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < GROUPS_PER_LAUNCH; ++k) t += v[k];   // group order: reproducible bit for bit
  *dst += t;
}

// ------------------------------------------------------------------------------------------
// losses (models/loss.py)
// ------------------------------------------------------------------------------------------
__global__ void k_targets_init(int n, const int32_t* __restrict__ ts, const int32_t* __restrict__ label,
                               int32_t* __restrict__ targets) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) targets[i] = ts[i] >= 0 ? label[i] : 0;
}

// per detection: latest positive past edge, earliest positive future edge (models/loss.py:26-43)
__global__ void k_targets_mark(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ seg_ptr,
                               const int32_t* __restrict__ inc, const int32_t* __restrict__ label,
                               int32_t* __restrict__ targets) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= *n_dets) return;
  const int s0 = seg_ptr[2 * k], s1 = seg_ptr[2 * k + 1], s2 = seg_ptr[2 * k + 2];
  for (int i = s1 - 1; i >= s0; --i)
    if (label[inc[i]]) { targets[inc[i]] = 1; break; }
  for (int i = s1; i < s2; ++i)
    if (label[inc[i]]) { targets[inc[i]] = 1; break; }
}

// one warp per segment (2 per detection): log-sum-exp, the chosen positive, loss / len
__global__ void __launch_bounds__(256)
k_ce_fwd(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ inc,
         const int32_t* __restrict__ targets, const float* __restrict__ logit, float* __restrict__ seg_lse,
         int32_t* __restrict__ seg_pos, float* __restrict__ seg_loss) {
  const int lane = threadIdx.x & 31;
  const int nseg = 2 * (*n_dets);
  for (int x = blockIdx.x * 8 + (threadIdx.x >> 5); x < nseg; x += gridDim.x * 8) {
    const int s0 = seg_ptr[x], s1 = seg_ptr[x + 1];
    const bool past = (x & 1) == 0;
    // position of the chosen positive: last one for past segments, first one for future segments
    int pos = past ? -1 : 0x7fffffff;
    float mx = -INFINITY;
    for (int i = s0 + lane; i < s1; i += 32) {
      const int e = inc[i];
      mx = fmaxf(mx, logit[e]);
      if (targets[e]) pos = past ? max(pos, i) : min(pos, i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const int p2 = __shfl_xor_sync(0xffffffffu, pos, o);
      pos = past ? max(pos, p2) : min(pos, p2);
    }
    const bool has = past ? pos >= 0 : pos != 0x7fffffff;
    float lse = 0.f, loss = 0.f;
    int pe = -1;
    if (has && s1 > s0) {
      float sum = 0.f;
      for (int i = s0 + lane; i < s1; i += 32) sum += expf(logit[inc[i]] - mx);
      sum = warp_sum_f(sum);
      lse = mx + logf(sum);
      pe = inc[pos];
      loss = (lse - logit[pe]) / (float)(s1 - s0);
    }
    if (lane == 0) { seg_lse[x] = lse; seg_pos[x] = pe; seg_loss[x] = loss; }
  }
}

// deterministic sum of a float array with one CTA (double accumulation), out[0] (+)= scale * sum
__global__ void __launch_bounds__(256) k_sum(const float* __restrict__ v, const int32_t* __restrict__ n_dev, int n_mul,
                                             int n_host, float scale_by_n, float* __restrict__ out) {
  __shared__ double red[256];
  const int n = n_dev ? n_mul * (*n_dev) : n_host;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)v[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(scale_by_n != 0.f ? red[0] / (double)max(n, 1) : red[0]);
}

// d logit of every edge row from its two segments (past segment of dst, future segment of src)
__global__ void k_ce_bwd(int n, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                         const int32_t* __restrict__ det_of_row, const int32_t* __restrict__ seg_ptr,
                         const float* __restrict__ seg_lse, const int32_t* __restrict__ seg_pos,
                         const float* __restrict__ logit, const float* __restrict__ gout, float* __restrict__ dlogit) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float d = 0.f;
  const int a = src[e];
  if (a >= 0) {
    const float g = gout[0], lg = logit[e];
    const int xs[2] = {2 * det_of_row[dst[e]], 2 * det_of_row[a] + 1};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int x = xs[q];
      if (seg_pos[x] >= 0) {
        const float len = (float)(seg_ptr[x + 1] - seg_ptr[x]);
        d += g * (expf(lg - seg_lse[x]) - (seg_pos[x] == e ? 1.0f : 0.0f)) / len;
      }
    }
  }
  dlogit[e] = d;
}

// FocalLoss(gamma = 0): -log(p_t + 1e-10) per element (models/loss.py:57-74)
__global__ void k_focal_fwd(int n, const float* __restrict__ p, const int64_t* __restrict__ t, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = -logf((t[i] == 1 ? p[i] : 1.0f - p[i]) + 1e-10f);
}
__global__ void k_focal_bwd(int n, const float* __restrict__ p, const int64_t* __restrict__ t,
                            const float* __restrict__ gout, float inv_n, float* __restrict__ dp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const bool pos = t[i] == 1;
    const float pt = (pos ? p[i] : 1.0f - p[i]) + 1e-10f;
    dp[i] = gout[0] * inv_n * (pos ? -1.0f : 1.0f) / pt;
  }
}

// weighted form for the batched trainer: every row carries its own weight (1 / rows of its kind in its chunk = the
// per-chunk means of train.py:76-85; 0 = the row does not count)
// out[blockIdx.x] = the block's sum (fixed order): the final k_sum adds n / 256 partials instead of n terms
__global__ void __launch_bounds__(256) k_wbce_fwd(int n, const float* __restrict__ p, const int64_t* __restrict__ t,
                                                  const float* __restrict__ w, float* __restrict__ out) {
  __shared__ float red[8];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.f;
  if (i < n && w[i] != 0.f) v = -w[i] * logf((t[i] == 1 ? p[i] : 1.0f - p[i]) + 1e-10f);
  v = warp_sum_f(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = ((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]));
}
__global__ void k_wbce_bwd(int n, const float* __restrict__ p, const int64_t* __restrict__ t, const float* __restrict__ w,
                           const float* __restrict__ gout, float* __restrict__ dp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const bool pos = t[i] == 1;
    const float pt = (pos ? p[i] : 1.0f - p[i]) + 1e-10f;
    dp[i] = w[i] != 0.f ? gout[0] * w[i] * (pos ? -1.0f : 1.0f) / pt : 0.f;
  }
}

// dst[seg_dst[s] + r][:] = src[seg_src[s] + r][:] for r < seg_len[s]: the rows a chunk carries from one step's
// block-diagonal layout into the next (float4 units; grid.y = segment)
__global__ void __launch_bounds__(256) k_rows_move(const float4* __restrict__ src, float4* __restrict__ dst,
                                                   const int32_t* __restrict__ seg_src, const int32_t* __restrict__ seg_dst,
                                                   const int32_t* __restrict__ seg_len, int ld4) {
  const int s = blockIdx.y;
  const size_t n4 = (size_t)seg_len[s] * ld4;
  const float4* a = src + (size_t)seg_src[s] * ld4;
  float4* b = dst + (size_t)seg_dst[s] * ld4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) b[i] = __ldg(a + i);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int tmpnn_loss_wbce_fwd(int n, const float* p, const int64_t* targets, const float* w, float* per_elem, float* loss,
                                   void* stream) {
  TMPNN_REQUIRE(p && targets && w && per_elem && loss && n > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  k_wbce_fwd<<<tmpnn_div_up(n, 256), 256, 0, st>>>(n, p, targets, w, per_elem);   // per_elem[0 .. n / 256]: block partials
  TMPNN_LAUNCH_CHECK();
  k_sum<<<1, 256, 0, st>>>(per_elem, nullptr, 0, tmpnn_div_up(n, 256), 0.f, loss);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_loss_wbce_bwd(int n, const float* p, const int64_t* targets, const float* w, const float* grad_out,
                                   float* dp, void* stream) {
  TMPNN_REQUIRE(p && targets && w && grad_out && dp && n > 0, "bad argument");
  k_wbce_bwd<<<tmpnn_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, p, targets, w, grad_out, dp);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_rows_move(const float* src, float* dst, const int32_t* seg_src, const int32_t* seg_dst,
                               const int32_t* seg_len, int n_seg, int ld, int max_len, void* stream) {
  TMPNN_REQUIRE(src && dst && seg_src && seg_dst && seg_len && ld % 4 == 0, "bad argument");
  TMPNN_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "rows must be 16-byte aligned");
  if (n_seg <= 0 || max_len <= 0) return TMPNN_OK;
  dim3 grid(max(1, min(tmpnn_div_up((long long)max_len * (ld / 4), 256 * 4), 64)), n_seg);
  k_rows_move<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (float4*)dst, seg_src, seg_dst, seg_len, ld / 4);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" size_t tmpnn_gate_bwd_partial_floats(void) { return (size_t)TMPNN_SM_COUNT * 8 * 2 * GB_PART; }
extern "C" int tmpnn_gate_bwd(int n_rows, const int32_t* src, const float* gates, const float* h_prev, const float* h_new,
                              int ldh, int col, const float* dh_out, const float* dlogits, const float* dscores,
                              const float* score, const float* head_w_edge, const float* head_w_node, float* dgi,
                              float* dgh, float* dhself, float* gbias_edge, float* gbias_node, float* ghw_edge,
                              float* ghw_node, float* ghb_edge, float* ghb_node, float* partials, void* stream) {
  TMPNN_REQUIRE(src && gates && h_prev && h_new && score && dgi && dgh && dhself, "null argument");
  if (n_rows <= 0) return TMPNN_OK;
  const int blocks = min(tmpnn_div_up(n_rows, 16), TMPNN_SM_COUNT * 8);
  k_gate_bwd<<<blocks, 256, 0, (cudaStream_t)stream>>>(n_rows, src, gates, h_prev, h_new, ldh, col, dh_out, dlogits, dscores,
                                                      score, head_w_edge, head_w_node, dgi, dgh, dhself, gbias_edge,
                                                      gbias_node, ghw_edge, ghw_node, ghb_edge, ghb_node, partials);
  TMPNN_LAUNCH_CHECK();
  if (partials) {
    k_gate_bwd_reduce<<<11, 1024, 0, (cudaStream_t)stream>>>(partials, blocks, gbias_edge, gbias_node, ghw_edge, ghw_node, ghb_edge,
                                                            ghb_node);
    TMPNN_LAUNCH_CHECK();
  }
  return TMPNN_OK;
}

extern "C" int tmpnn_rows_times_w(const int32_t* r_dev, int r_host, const int32_t* a_rows, const int32_t* c_rows,
                                  const int32_t* mask, const float* A, const float* W, int n, float* C, int ldc,
                                  int accumulate, void* stream) {
  TMPNN_REQUIRE(A && W && C && (n == 64 || n == 128), "bad argument");
  TMPNN_REQUIRE(ldc % 4 == 0 && ((uintptr_t)C & 15) == 0, "C rows must be 16-byte aligned");
  const int r_max = r_dev ? TMPNN_SM_COUNT * 32 * 4 : r_host;
  if (r_max <= 0) return TMPNN_OK;
  const int blocks = min(tmpnn_div_up(r_max, 32), TMPNN_SM_COUNT * 4);
  if (n == 64)
    k_rows_times_w<64><<<blocks, 256, 0, (cudaStream_t)stream>>>(r_dev, r_host, a_rows, c_rows, mask, A, W, C, ldc, accumulate);
  else
    k_rows_times_w<128><<<blocks, 256, 0, (cudaStream_t)stream>>>(r_dev, r_host, a_rows, c_rows, mask, A, W, C, ldc, accumulate);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_rows_outer(const int32_t* r_dev, int r_host, const int32_t* a_rows, const int32_t* b_rows,
                                const int32_t* mask, const float* A, const float* B, int ldb, int n, float* G,
                                void* stream) {
  TMPNN_REQUIRE(A && B && G && (n == 64 || n == 128), "bad argument");
  const int r_max = r_dev ? TMPNN_SM_COUNT * 32 : r_host;
  if (r_max <= 0) return TMPNN_OK;
  const int blocks = min(tmpnn_div_up(r_max, 32), TMPNN_SM_COUNT * 2);
  if (n == 64)
    k_rows_outer<64><<<blocks, 256, 0, (cudaStream_t)stream>>>(r_dev, r_host, a_rows, b_rows, mask, A, B, ldb, G);
  else
    k_rows_outer<128><<<blocks, 256, 0, (cudaStream_t)stream>>>(r_dev, r_host, a_rows, b_rows, mask, A, B, ldb, G);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_scatter_bwd(const tmpnn_graph* g, const tmpnn_index* ix, int n_rows, const float* dhself,
                                 const float* dx, int kx, const float* dagg, float* dh_in, int ldh, int col, void* stream) {
  TMPNN_REQUIRE(g && ix && dhself && dx && dagg && dh_in && g->num_seqs == 1, "bad argument (single-slab graphs only)");
  TMPNN_REQUIRE(kx == H || kx == 2 * H, "kx must be 64 or 128");
  if (n_rows <= 0) return TMPNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  k_scatter_bwd_edges<<<min(tmpnn_div_up(n_rows, 16), TMPNN_SM_COUNT * 8), 256, 0, st>>>(n_rows, g->src, g->dst, ix->det_of_row,
                                                                                        dhself, dagg, dh_in, ldh, col);
  TMPNN_LAUNCH_CHECK();
  k_scatter_bwd_dets<<<min(tmpnn_div_up(n_rows, 8), TMPNN_SM_COUNT * 4), 256, 0, st>>>(ix->n_dets, ix->det_rows, ix->seg_ptr,
                                                                                      ix->inc, dhself, dx, kx, dh_in, ldh, col);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_input_bwd(const float* x, int ldx, int col0, int f_in, const int32_t* x_idx, const float* a,
                               const float* mean, const float* var, const float* gamma, const float* beta,
                               const float* b1, const float* w2, const float* dh, int ldh, int col,
                               const int32_t* out_rows, int n, int n_edge_rows, int training, float* scratch,
                               float* gw1, float* gb1, float* ggamma, float* gbeta, float* gw2, float* gb2, void* stream) {
  TMPNN_REQUIRE(x && a && mean && var && gamma && beta && b1 && w2 && dh && out_rows && scratch, "null argument");
  TMPNN_REQUIRE(gw1 && gb1 && ggamma && gbeta && gw2 && gb2, "null gradient buffer");
  if (n <= 0) return TMPNN_OK;
  k_input_bwd<<<1, 256, 0, (cudaStream_t)stream>>>(x, ldx, col0, f_in, x_idx, a, mean, var, gamma, beta, b1, w2, dh, ldh, col,
                                                  out_rows, n, n_edge_rows, training, scratch, gw1, gb1, ggamma, gbeta, gw2,
                                                  gb2);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// ---- forward of the input transform for several groups (BatchNorm batches) in one launch each ----------------
// batch statistics of every group (same two-pass arithmetic as k_input_bn_stats in mp_step.cu), one CTA per group
__global__ void __launch_bounds__(256) k_input_bn_stats_groups(const InputGroupBatch groups, const float* __restrict__ b1) {
  __shared__ double red[4][H];
  const tmpnn_input_group& g = groups.g[blockIdx.x];
  const int n = g.n, n_edge = g.n_edge_rows;
  if (n <= 0) return;
  const float* a = g.a;
  const int j = threadIdx.x & 63, part = threadIdx.x >> 6;
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
    const_cast<float*>(g.mean)[j] = (float)mu;
    const_cast<float*>(g.var)[j] = (float)var;
  }
}
// running statistics: the exponential averages depend on the order, so one thread per channel walks the groups in order
__global__ void k_bn_running_groups(const InputGroupBatch groups, int n_groups, float* __restrict__ rmean,
                                    float* __restrict__ rvar) {
  const int j = threadIdx.x;
  float m = rmean[j], v = rvar[j];
  // the groups' statistics are fetched up front (independent loads), the recurrence itself stays in group order
  float gm[GROUPS_PER_LAUNCH], gv[GROUPS_PER_LAUNCH];
#pragma unroll
  for (int k = 0; k < GROUPS_PER_LAUNCH; ++k) {
    const bool on = k < n_groups && groups.g[k].n > 0;
    gm[k] = on ? groups.g[k].mean[j] : 0.f;
    gv[k] = on ? groups.g[k].var[j] : 0.f;
  }
  // unbiased variance: var n / (n - 1) in double like before, but all the (slow) fp64 divisions issue before the recurrence
  float gu[GROUPS_PER_LAUNCH];
#pragma unroll
  for (int k = 0; k < GROUPS_PER_LAUNCH; ++k) {
    const double ntot = k < n_groups ? (double)groups.g[k].n + (double)groups.g[k].n_edge_rows : 2.0;
    gu[k] = (float)((double)gv[k] * ntot / (ntot - 1.0));
  }
#pragma unroll
  for (int k = 0; k < GROUPS_PER_LAUNCH; ++k) {
    if (k >= n_groups || groups.g[k].n <= 0) continue;
    m = 0.9f * m + 0.1f * gm[k];
    v = 0.9f * v + 0.1f * gu[k];
  }
  rmean[j] = m;
  rvar[j] = v;
}
// BatchNorm -> ReLU -> Linear2 of every group with its own statistics, one CTA per group (8 warps, two rows per warp pass;
// the per-channel constants are fetched once)
__global__ void __launch_bounds__(256)
k_input_bn_relu_linear2_groups(const InputGroupBatch groups, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ h, int ldh,
                               int col) {
  __shared__ float w2t[H * (H + 1)];  // [j][k], rows padded to 65 floats: coalesced global reads, conflict-free reads below
  __shared__ float act[8][2][H];
  const tmpnn_input_group& g = groups.g[blockIdx.x];
  const int n = g.n;
  if (n <= 0) return;
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) w2t[(i >> 6) * (H + 1) + (i & 63)] = w2[i];
  const float* a = g.a;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // (a - mean) / sqrt(var + eps) * gamma + beta, written like torch's batch_norm
  const float mu0 = g.mean[lane], mu1 = g.mean[lane + 32];
  const float is0 = 1.0f / sqrtf(g.var[lane] + 1e-5f), is1 = 1.0f / sqrtf(g.var[lane + 32] + 1e-5f);
  const float ga0 = gamma[lane], ga1 = gamma[lane + 32], be0 = beta[lane], be1 = beta[lane + 32];
  const float bo0 = b2[lane], bo1 = b2[lane + 32];
  __syncthreads();
  for (int r = 2 * w; r < n; r += 16) {
    const bool two = r + 1 < n;
    const float x00 = a[(size_t)r * H + lane], x01 = a[(size_t)r * H + lane + 32];
    const float x10 = two ? a[(size_t)(r + 1) * H + lane] : 0.f, x11 = two ? a[(size_t)(r + 1) * H + lane + 32] : 0.f;
    const int orow0 = g.out_rows[r], orow1 = two ? g.out_rows[r + 1] : 0;
    act[w][0][lane] = fmaxf((x00 - mu0) * is0 * ga0 + be0, 0.f);
    act[w][0][lane + 32] = fmaxf((x01 - mu1) * is1 * ga1 + be1, 0.f);
    act[w][1][lane] = fmaxf((x10 - mu0) * is0 * ga0 + be0, 0.f);
    act[w][1][lane + 32] = fmaxf((x11 - mu1) * is1 * ga1 + be1, 0.f);
    __syncwarp();
    float o00 = bo0, o01 = bo1, o10 = bo0, o11 = bo1;
#pragma unroll 8
    for (int k = 0; k < H; ++k) {
      const float w0 = w2t[lane * (H + 1) + k], w1 = w2t[(lane + 32) * (H + 1) + k];
      const float a0 = act[w][0][k], a1 = act[w][1][k];
      o00 = fmaf(a0, w0, o00); o01 = fmaf(a0, w1, o01);
      o10 = fmaf(a1, w0, o10); o11 = fmaf(a1, w1, o11);
    }
    __syncwarp();
    float* hr = h + (size_t)orow0 * ldh + col;
    hr[lane] = o00;
    hr[lane + 32] = o01;
    if (two) {
      hr = h + (size_t)orow1 * ldh + col;
      hr[lane] = o10;
      hr[lane + 32] = o11;
    }
  }
}

extern "C" int tmpnn_input_bn_groups_fwd(const tmpnn_input_group* groups, int n_groups, const float* b1, const float* gamma,
                                         const float* beta, const float* w2, const float* b2, float* running_mean,
                                         float* running_var, float* h, int ldh, int col, void* stream) {
  TMPNN_REQUIRE(groups && b1 && gamma && beta && w2 && b2 && h, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int g0 = 0; g0 < n_groups; g0 += GROUPS_PER_LAUNCH) {
    InputGroupBatch batch;
    const int nb = min(GROUPS_PER_LAUNCH, n_groups - g0);
    for (int k = 0; k < nb; ++k) {
      batch.g[k] = groups[g0 + k];
      TMPNN_REQUIRE(batch.g[k].n <= 0 || batch.g[k].n + batch.g[k].n_edge_rows > 1,
                    "Expected more than 1 value per channel when training");
    }
    k_input_bn_stats_groups<<<nb, 256, 0, st>>>(batch, b1);
    TMPNN_LAUNCH_CHECK();
    if (running_mean && running_var) {
      k_bn_running_groups<<<1, H, 0, st>>>(batch, nb, running_mean, running_var);
      TMPNN_LAUNCH_CHECK();
    }
    k_input_bn_relu_linear2_groups<<<nb, 256, 0, st>>>(batch, gamma, beta, w2, b2, h, ldh, col);
    TMPNN_LAUNCH_CHECK();
  }
  return TMPNN_OK;
}

extern "C" int tmpnn_input_bwd_groups(const float* x, int ldx, int col0, int f_in, const tmpnn_input_group* groups,
                                      int n_groups, const float* gamma, const float* beta, const float* b1,
                                      const float* w2, const float* dh, int ldh, int col, int training, float* gw1,
                                      float* gb1, float* ggamma, float* gbeta, float* gw2, float* gb2, float* partials,
                                      void* stream) {
  TMPNN_REQUIRE(x && groups && gamma && beta && b1 && w2 && dh, "null argument");
  TMPNN_REQUIRE(gw1 && gb1 && ggamma && gbeta && gw2 && gb2, "null gradient buffer");
  TMPNN_REQUIRE(!partials || f_in <= H, "partial buffers hold at most 64 input features per group");
  cudaStream_t st = (cudaStream_t)stream;
  for (int g0 = 0; g0 < n_groups; g0 += GROUPS_PER_LAUNCH) {
    InputGroupBatch batch;
    const int nb = min(GROUPS_PER_LAUNCH, n_groups - g0);
    for (int k = 0; k < nb; ++k) batch.g[k] = groups[g0 + k];
    if (partials) {
      k_input_bwd_groups_part<<<nb, 256, 0, st>>>(x, ldx, col0, f_in, batch, gamma, beta, b1, w2, dh, ldh, col, training, partials);
      TMPNN_LAUNCH_CHECK();
      k_input_bwd_reduce<<<tmpnn_div_up(IBW_PART, 256), 256, 0, st>>>(batch, nb, f_in, partials, gw1, gb1, ggamma, gbeta, gw2, gb2);
    } else {
      k_input_bwd_groups<<<nb, 256, 0, st>>>(x, ldx, col0, f_in, batch, gamma, beta, b1, w2, dh, ldh, col, training, gw1, gb1,
                                             ggamma, gbeta, gw2, gb2);
    }
    TMPNN_LAUNCH_CHECK();
  }
  return TMPNN_OK;
}
extern "C" size_t tmpnn_input_bwd_partial_floats(int n_groups) { return (size_t)min(max(n_groups, 1), GROUPS_PER_LAUNCH) * IBW_PART; }

extern "C" int tmpnn_loss_targets(const tmpnn_graph* g, const tmpnn_index* ix, int n_rows, int32_t* targets, void* stream) {
  TMPNN_REQUIRE(g && ix && targets && g->label && g->num_seqs == 1, "bad argument (single-slab graph with labels)");
  if (n_rows <= 0) return TMPNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  k_targets_init<<<tmpnn_div_up(n_rows, 256), 256, 0, st>>>(n_rows, g->ts, g->label, targets);
  TMPNN_LAUNCH_CHECK();
  k_targets_mark<<<tmpnn_div_up(n_rows, 128), 128, 0, st>>>(ix->n_dets, ix->seg_ptr, ix->inc, g->label, targets);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_loss_ce_fwd(const tmpnn_index* ix, int n_rows, const int32_t* targets, const float* logit,
                                 float* seg_lse, int32_t* seg_pos, float* seg_loss, float* loss, void* stream) {
  TMPNN_REQUIRE(ix && targets && logit && seg_lse && seg_pos && seg_loss && loss, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  k_ce_fwd<<<max(1, min(tmpnn_div_up(n_rows, 8), TMPNN_SM_COUNT * 4)), 256, 0, st>>>(ix->n_dets, ix->seg_ptr, ix->inc, targets, logit,
                                                                                    seg_lse, seg_pos, seg_loss);
  TMPNN_LAUNCH_CHECK();
  k_sum<<<1, 256, 0, st>>>(seg_loss, ix->n_dets, 2, 0, 0.f, loss);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_loss_ce_bwd(const tmpnn_graph* g, const tmpnn_index* ix, int n_rows, const float* seg_lse,
                                 const int32_t* seg_pos, const float* logit, const float* grad_out, float* dlogit,
                                 void* stream) {
  TMPNN_REQUIRE(g && ix && seg_lse && seg_pos && logit && grad_out && dlogit, "null argument");
  if (n_rows <= 0) return TMPNN_OK;
  k_ce_bwd<<<tmpnn_div_up(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(n_rows, g->src, g->dst, ix->det_of_row, ix->seg_ptr, seg_lse,
                                                                       seg_pos, logit, grad_out, dlogit);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_loss_focal_fwd(int n, const float* p, const int64_t* targets, float* per_elem, float* loss, void* stream) {
  TMPNN_REQUIRE(p && targets && per_elem && loss && n > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  k_focal_fwd<<<tmpnn_div_up(n, 256), 256, 0, st>>>(n, p, targets, per_elem);
  TMPNN_LAUNCH_CHECK();
  k_sum<<<1, 256, 0, st>>>(per_elem, nullptr, 0, n, 1.0f, loss);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_loss_focal_bwd(int n, const float* p, const int64_t* targets, const float* grad_out, float* dp, void* stream) {
  TMPNN_REQUIRE(p && targets && grad_out && dp && n > 0, "bad argument");
  k_focal_bwd<<<tmpnn_div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(n, p, targets, grad_out, 1.0f / (float)n, dp);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
