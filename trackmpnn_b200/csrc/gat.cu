// gat.cu -- attention-weighted aggregation for the detection rows (--num-att-heads > 0): eval-mode forward,
// train-mode forward (dropout on the attention through a caller-supplied keep mask) and the backward.
//
// Reference models/layers.py:7-46 (GraphAttentionLayer) and :105-112 (heads averaged), restated on the
// edge list: for one head with parameters W_att [64, 64], a [64]
//   e_j     = LeakyReLU_0.2( a . | W_att^T h[src_j] - W_att^T h[dst_j] | )        per edge row j
//   alpha_dj = softmax over the edges j incident to detection d of e_j
//   agg[d] += (1 / heads) * sum_j alpha_dj * (+1 if d == src_j else -1) * h[j]
// which replaces the plain signed sum of tmpnn_aggregate_dets as the node GRU's input.  The dense
// N x N attention matrix of the reference is never formed; alpha is kept per incidence entry.
// Training (tmpnn_gat_aggregate_dets_train / tmpnn_gat_bwd): nn.Dropout(0.5) acts on the softmax output
// (layers.py:37), i.e. on every incidence entry independently: alpha~_dj = keep_dj ? alpha_dj / (1 - p) : 0 with the
// keep mask drawn by the caller (one byte per incidence entry).  The backward walks the same lists:
//   d alpha~_dj = (1/heads) s_dj (dagg[d] . h[j]);  softmax: d e_dj = alpha_dj (d alpha_dj - sum_j' alpha_dj' d alpha_dj');
//   e_j is shared by the edge's two detections: d pre_j = (d e_{src,j} + d e_{dst,j}) LeakyReLU'(pre_j);
//   d a += d pre_j |p - q|;  d hatt[src] += d pre_j a * sign(p - q), d hatt[dst] -= the same (p, q = hatt of src, dst);
//   d W_att += h_det^T d hatt;  d h_det += d hatt W_att^T;  d h[j] += (1/heads)(alpha~_{src,j} dagg[src] - alpha~_{dst,j} dagg[dst]).
// No atomics on the state gradient: per-edge quantities are stored per (edge, side) and summed by the edge pass.
#include "common.cuh"

namespace {

constexpr int H = TMPNN_HIDDEN;

// h_att[k] = h[det_rows[k]] . W_att  (only detection rows are ever endpoints); one warp per detection
__global__ void __launch_bounds__(256) k_gat_project(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows,
                                                     const float* __restrict__ h, int ldh, int col,
                                                     const float* __restrict__ w_att, float* __restrict__ hatt,
                                                     const int32_t* __restrict__ phys) {
  __shared__ float w[H * H];
  __shared__ float hr[8][H];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) w[i] = w_att[i];
  __syncthreads();
  const int nd = *n_dets;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  for (int k = blockIdx.x * 8 + wp; k < nd; k += gridDim.x * 8) {
    const int row = det_rows[k];
    const float* r = h + (size_t)(phys ? phys[row] : row) * ldh + col;
    hr[wp][lane] = r[lane];
    hr[wp][lane + 32] = r[lane + 32];
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
    for (int c = 0; c < H; ++c) {
      const float v = hr[wp][c];
      a0 = fmaf(v, w[c * H + lane], a0);
      a1 = fmaf(v, w[c * H + lane + 32], a1);
    }
    __syncwarp();
    hatt[(size_t)k * H + lane] = a0;
    hatt[(size_t)k * H + lane + 32] = a1;
  }
}

// e_j per edge row: half-warp per row
__global__ void __launch_bounds__(256) k_gat_edge_score(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ src,
                                                        const int32_t* __restrict__ dst, int cap_rows,
                                                        const int32_t* __restrict__ det_of_row,
                                                        const float* __restrict__ hatt, const float* __restrict__ a,
                                                        float* __restrict__ escore) {
  const int s = blockIdx.y, n = n_rows[s];
  const size_t base = (size_t)s * cap_rows;
  const int l16 = threadIdx.x & 15;
  const float4 av = ldg4(a + 4 * l16);
  for (int r = blockIdx.x * 16 + (threadIdx.x >> 4); r < n; r += gridDim.x * 16) {
    const int sa = src[base + r];
    float v = 0.f;
    if (sa >= 0) {
      const float4 p = ldg4(hatt + (size_t)det_of_row[base + sa] * H + 4 * l16);
      const float4 q = ldg4(hatt + (size_t)det_of_row[base + dst[base + r]] * H + 4 * l16);
      v = av.x * fabsf(p.x - q.x) + av.y * fabsf(p.y - q.y) + av.z * fabsf(p.z - q.z) + av.w * fabsf(p.w - q.w);
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l16 == 0) escore[base + r] = sa >= 0 ? (v > 0.f ? v : 0.2f * v) : 0.f;
  }
}

// softmax over each detection's incidence list and the weighted signed sum; one warp per detection
__global__ void __launch_bounds__(256) k_gat_aggregate(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ seg_ptr,
                                                       const int32_t* __restrict__ inc, const float* __restrict__ escore,
                                                       const float* __restrict__ h, int ldh, int col, float scale,
                                                       int accumulate, float* __restrict__ agg, float* __restrict__ alpha,
                                                       const int32_t* __restrict__ phys, const uint8_t* __restrict__ keep,
                                                       float keep_scale, float* __restrict__ att_edge) {
  const int nd = *n_dets;
  const int lane = threadIdx.x & 31;
  for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < nd; k += gridDim.x * 8) {
    const int s0 = seg_ptr[2 * k], s1 = seg_ptr[2 * k + 1], s2 = seg_ptr[2 * k + 2];
    float mx = -INFINITY;
    for (int i = s0 + lane; i < s2; i += 32) mx = fmaxf(mx, escore[inc[i]]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int i = s0 + lane; i < s2; i += 32) sum += expf(escore[inc[i]] - mx);
    sum = warp_sum_f(sum);
    float a0 = 0.f, a1 = 0.f;
    for (int i = s0; i < s2; ++i) {  // ascending rows: past edges (-), then future edges (+)
      const int e = inc[i];
      const float w = expf(escore[e] - mx) / sum;
      const float wd = keep ? (keep[i] ? w * keep_scale : 0.f) : w;  // dropout on the attention (training only)
      const float sw = i < s1 ? -wd : wd;
      const float* r = h + (size_t)(phys ? phys[e] : e) * ldh + col;
      a0 = fmaf(sw, r[lane], a0);
      a1 = fmaf(sw, r[lane + 32], a1);
      if (alpha && lane == 0) alpha[i] = w;
      if (att_edge && lane == 0) att_edge[2 * (size_t)e + (i < s1)] = wd;  // side 0: d is the edge's src, 1: its dst
    }
    float* o = agg + (size_t)k * H;
    if (accumulate) { o[lane] += scale * a0; o[lane + 32] += scale * a1; }
    else { o[lane] = scale * a0; o[lane + 32] = scale * a1; }
  }
}

// ---- backward -------------------------------------------------------------------------------------------------
// softmax backward per detection (one warp each): de_side[2 e + side] = d e_dj
__global__ void __launch_bounds__(256) k_gat_bwd_softmax(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ seg_ptr,
                                                         const int32_t* __restrict__ inc, const float* __restrict__ h, int ldh,
                                                         int col, const float* __restrict__ dagg, const float* __restrict__ alpha,
                                                         const uint8_t* __restrict__ keep, float keep_scale, float scale,
                                                         float* __restrict__ dal, float* __restrict__ de_side) {
  const int nd = *n_dets;
  const int lane = threadIdx.x & 31;
  for (int k = blockIdx.x * 8 + (threadIdx.x >> 5); k < nd; k += gridDim.x * 8) {
    const int s0 = seg_ptr[2 * k], s1 = seg_ptr[2 * k + 1], s2 = seg_ptr[2 * k + 2];
    const float g0 = dagg[(size_t)k * H + lane], g1 = dagg[(size_t)k * H + lane + 32];
    float dot_sum = 0.f;  // sum_j alpha_dj d alpha_dj
    for (int i = s0; i < s2; ++i) {
      const float* r = h + (size_t)inc[i] * ldh + col;
      const float dot = warp_sum_f(g0 * r[lane] + g1 * r[lane + 32]);
      const float m = keep ? (keep[i] ? keep_scale : 0.f) : 1.f;
      const float d = scale * (i < s1 ? -dot : dot) * m;
      dot_sum = fmaf(alpha[i], d, dot_sum);
      if (lane == 0) dal[i] = d;
    }
    __syncwarp();
    for (int i = s0 + lane; i < s2; i += 32) de_side[2 * (size_t)inc[i] + (i < s1)] = alpha[i] * (dal[i] - dot_sum);
  }
}

// per edge row (half-warp each): d pre_j, the state gradient of the edge row, the partial d a
__global__ void __launch_bounds__(256) k_gat_bwd_edges(int n_rows, const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                       const int32_t* __restrict__ det_of_row, const float* __restrict__ hatt,
                                                       const float* __restrict__ a, const float* __restrict__ escore,
                                                       const float* __restrict__ de_side, const float* __restrict__ att_edge,
                                                       const float* __restrict__ dagg, float scale, float* __restrict__ dpre,
                                                       float* __restrict__ dh_in, int ldh, int col, float* __restrict__ da) {
  __shared__ float red[16][H];
  const int l16 = threadIdx.x & 15, hw = threadIdx.x >> 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = blockIdx.x * 16 + hw; r < n_rows; r += gridDim.x * 16) {
    const int sa = src[r];
    if (sa < 0) continue;
    const int ks = det_of_row[sa], kd = det_of_row[dst[r]];
    const float de = de_side[2 * (size_t)r] + de_side[2 * (size_t)r + 1];
    const float dp = escore[r] > 0.f ? de : 0.2f * de;
    const float4 p = ldg4(hatt + (size_t)ks * H + 4 * l16), q = ldg4(hatt + (size_t)kd * H + 4 * l16);
    acc.x = fmaf(dp, fabsf(p.x - q.x), acc.x);
    acc.y = fmaf(dp, fabsf(p.y - q.y), acc.y);
    acc.z = fmaf(dp, fabsf(p.z - q.z), acc.z);
    acc.w = fmaf(dp, fabsf(p.w - q.w), acc.w);
    if (l16 == 0) dpre[r] = dp;
    const float w0 = scale * att_edge[2 * (size_t)r], w1 = scale * att_edge[2 * (size_t)r + 1];
    const float4 ga = ldg4(dagg + (size_t)ks * H + 4 * l16), gb = ldg4(dagg + (size_t)kd * H + 4 * l16);
    float4* o = reinterpret_cast<float4*>(dh_in + (size_t)r * ldh + col + 4 * l16);
    float4 v = *o;
    v.x += w0 * ga.x - w1 * gb.x;
    v.y += w0 * ga.y - w1 * gb.y;
    v.z += w0 * ga.z - w1 * gb.z;
    v.w += w0 * ga.w - w1 * gb.w;
    *o = v;
  }
  *reinterpret_cast<float4*>(&red[hw][4 * l16]) = acc;
  __syncthreads();
  if (threadIdx.x < H) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += red[i][threadIdx.x];
    if (s != 0.f) atomicAdd(da + threadIdx.x, s);
  }
}

// per detection (one warp each): d hatt[k] from its incident edges, then d h_det += d hatt . W_att^T
__global__ void __launch_bounds__(256) k_gat_bwd_dets(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows,
                                                      const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ inc,
                                                      const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                                                      const int32_t* __restrict__ det_of_row, const float* __restrict__ hatt,
                                                      const float* __restrict__ a, const float* __restrict__ dpre,
                                                      const float* __restrict__ w_att, float* __restrict__ dhatt,
                                                      float* __restrict__ dh_in, int ldh, int col) {
  __shared__ float wt[H * H];  // wt[o][c] = W_att[c][o]
  __shared__ float dv[8][H];
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) wt[(i % H) * H + i / H] = w_att[i];
  __syncthreads();
  const int nd = *n_dets;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const float av0 = a[lane], av1 = a[lane + 32];
  for (int k = blockIdx.x * 8 + wp; k < nd; k += gridDim.x * 8) {
    const int s0 = seg_ptr[2 * k], s1 = seg_ptr[2 * k + 1], s2 = seg_ptr[2 * k + 2];
    float a0 = 0.f, a1 = 0.f;
    for (int i = s0; i < s2; ++i) {
      const int e = inc[i];
      const float* p = hatt + (size_t)det_of_row[src[e]] * H;
      const float* q = hatt + (size_t)det_of_row[dst[e]] * H;
      const float dp = i < s1 ? -dpre[e] : dpre[e];
      const float d0 = p[lane] - q[lane], d1 = p[lane + 32] - q[lane + 32];
      a0 = fmaf(dp * av0, (float)((d0 > 0.f) - (d0 < 0.f)), a0);
      a1 = fmaf(dp * av1, (float)((d1 > 0.f) - (d1 < 0.f)), a1);
    }
    dhatt[(size_t)k * H + lane] = a0;
    dhatt[(size_t)k * H + lane + 32] = a1;
    dv[wp][lane] = a0;
    dv[wp][lane + 32] = a1;
    __syncwarp();
    float b0 = 0.f, b1 = 0.f;
#pragma unroll 8
    for (int o = 0; o < H; ++o) {
      const float v = dv[wp][o];
      b0 = fmaf(v, wt[o * H + lane], b0);
      b1 = fmaf(v, wt[o * H + lane + 32], b1);
    }
    __syncwarp();
    float* out = dh_in + (size_t)det_rows[k] * ldh + col;
    out[lane] += b0;
    out[lane + 32] += b1;
  }
}

// d W_att[c][o] += sum_k h[det k][c] d hatt[k][o]; CTA (c, part) over a slice of the detections
__global__ void __launch_bounds__(H) k_gat_bwd_w(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows,
                                                 const float* __restrict__ h, int ldh, int col, const float* __restrict__ dhatt,
                                                 float* __restrict__ dw) {
  const int nd = *n_dets, c = blockIdx.x, o = threadIdx.x;
  float acc = 0.f;
  for (int k = blockIdx.y; k < nd; k += gridDim.y) acc = fmaf(h[(size_t)det_rows[k] * ldh + col + c], dhatt[(size_t)k * H + o], acc);
  if (acc != 0.f) atomicAdd(dw + c * H + o, acc);
}

}  // namespace

static int gat_forward(const tmpnn_graph* g, const tmpnn_index* ix, const float* h, int ldh, int col, const float* w_att,
                       const float* a, int head, int num_heads, float* hatt, float* escore, float* agg, float* alpha,
                       const uint8_t* keep, float keep_scale, float* att_edge, cudaStream_t st) {
  k_gat_project<<<TMPNN_SM_COUNT * 2, 256, 0, st>>>(ix->n_dets, ix->det_rows, h, ldh, col, w_att, hatt, g->phys);
  TMPNN_LAUNCH_CHECK();
  dim3 grid(max(1, min(tmpnn_div_up(g->cap_rows, 16), TMPNN_SM_COUNT * 8 / max(1, min(g->num_seqs, 8)))), g->num_seqs);
  k_gat_edge_score<<<grid, 256, 0, st>>>(g->n_rows, g->src, g->dst, g->cap_rows, ix->det_of_row, hatt, a, escore);
  TMPNN_LAUNCH_CHECK();
  k_gat_aggregate<<<TMPNN_SM_COUNT * 8, 256, 0, st>>>(ix->n_dets, ix->seg_ptr, ix->inc, escore, h, ldh, col,
                                                     1.0f / (float)num_heads, head > 0, agg, alpha, g->phys, keep, keep_scale,
                                                     att_edge);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_gat_aggregate_dets_train(const tmpnn_graph* g, const tmpnn_index* ix, const float* h, int ldh, int col,
                                              const float* w_att, const float* a, int head, int num_heads,
                                              const uint8_t* keep, float keep_scale, float* hatt, float* escore, float* agg,
                                              float* alpha, float* att_edge, void* stream) {
  TMPNN_REQUIRE(g && ix && h && w_att && a && hatt && escore && agg && alpha && att_edge, "null argument");
  TMPNN_REQUIRE(ldh % 4 == 0 && col % 4 == 0 && head >= 0 && head < num_heads, "bad argument");
  TMPNN_REQUIRE(g->num_seqs == 1 && g->phys == nullptr, "training graphs are single, compacted slabs");
  cudaStream_t st = (cudaStream_t)stream;
  TMPNN_CUDA_TRY(cudaMemsetAsync(att_edge, 0, sizeof(float) * 2 * (size_t)g->cap_rows, st));
  return gat_forward(g, ix, h, ldh, col, w_att, a, head, num_heads, hatt, escore, agg, alpha, keep, keep_scale, att_edge, st);
}

extern "C" int tmpnn_gat_bwd(const tmpnn_graph* g, const tmpnn_index* ix, int n_rows, const float* h, int ldh, int col,
                             const float* w_att, const float* a, int num_heads, const uint8_t* keep, float keep_scale,
                             const float* hatt, const float* escore, const float* alpha, const float* att_edge,
                             const float* dagg, float* dal, float* de_side, float* dpre, float* dhatt, float* dh_in,
                             float* dw_att, float* da, void* stream) {
  TMPNN_REQUIRE(g && ix && h && w_att && a && hatt && escore && alpha && att_edge && dagg, "null argument");
  TMPNN_REQUIRE(dal && de_side && dpre && dhatt && dh_in && dw_att && da, "null output / scratch");
  TMPNN_REQUIRE(ldh % 4 == 0 && col % 4 == 0 && num_heads > 0 && n_rows >= 0 && n_rows <= g->cap_rows, "bad argument");
  TMPNN_REQUIRE(g->num_seqs == 1 && g->phys == nullptr, "training graphs are single, compacted slabs");
  if (n_rows == 0) return TMPNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = 1.0f / (float)num_heads;
  TMPNN_CUDA_TRY(cudaMemsetAsync(de_side, 0, sizeof(float) * 2 * (size_t)n_rows, st));
  k_gat_bwd_softmax<<<min(tmpnn_div_up(n_rows, 8), TMPNN_SM_COUNT * 8), 256, 0, st>>>(ix->n_dets, ix->seg_ptr, ix->inc, h, ldh, col,
                                                                                     dagg, alpha, keep, keep_scale, scale, dal,
                                                                                     de_side);
  TMPNN_LAUNCH_CHECK();
  k_gat_bwd_edges<<<min(tmpnn_div_up(n_rows, 16), TMPNN_SM_COUNT * 4), 256, 0, st>>>(n_rows, g->src, g->dst, ix->det_of_row, hatt, a,
                                                                                    escore, de_side, att_edge, dagg, scale, dpre,
                                                                                    dh_in, ldh, col, da);
  TMPNN_LAUNCH_CHECK();
  k_gat_bwd_dets<<<min(tmpnn_div_up(n_rows, 8), TMPNN_SM_COUNT * 2), 256, 0, st>>>(ix->n_dets, ix->det_rows, ix->seg_ptr, ix->inc,
                                                                                  g->src, g->dst, ix->det_of_row, hatt, a, dpre,
                                                                                  w_att, dhatt, dh_in, ldh, col);
  TMPNN_LAUNCH_CHECK();
  k_gat_bwd_w<<<dim3(H, 16), H, 0, st>>>(ix->n_dets, ix->det_rows, h, ldh, col, dhatt, dw_att);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_gat_aggregate_dets(const tmpnn_graph* g, const tmpnn_index* ix, const float* h, int ldh, int col,
                                        const float* w_att, const float* a, int head, int num_heads, float* hatt,
                                        float* escore, float* agg, float* alpha, void* stream) {
  TMPNN_REQUIRE(g && ix && h && w_att && a && hatt && escore && agg, "null argument");
  TMPNN_REQUIRE(ldh % 4 == 0 && col % 4 == 0 && head >= 0 && head < num_heads, "bad argument");
  return gat_forward(g, ix, h, ldh, col, w_att, a, head, num_heads, hatt, escore, agg, alpha, nullptr, 1.0f, nullptr,
                     (cudaStream_t)stream);
}
