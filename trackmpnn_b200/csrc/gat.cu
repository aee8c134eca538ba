// gat.cu -- attention-weighted aggregation for the detection rows (--num-att-heads > 0), eval mode.
//
// Reference models/layers.py:7-46 (GraphAttentionLayer) and :105-112 (heads averaged), restated on the
// edge list: for one head with parameters W_att [64, 64], a [64]
//   e_j     = LeakyReLU_0.2( a . | W_att^T h[src_j] - W_att^T h[dst_j] | )        per edge row j
//   alpha_dj = softmax over the edges j incident to detection d of e_j
//   agg[d] += (1 / heads) * sum_j alpha_dj * (+1 if d == src_j else -1) * h[j]
// which replaces the plain signed sum of tmpnn_aggregate_dets as the node GRU's input.  The dense
// N x N attention matrix of the reference is never formed; alpha is kept per incidence entry.
// Dropout(0.5) on the attention only exists in training mode, which this path does not cover.
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
                                                       const int32_t* __restrict__ phys) {
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
      const float sw = i < s1 ? -w : w;
      const float* r = h + (size_t)(phys ? phys[e] : e) * ldh + col;
      a0 = fmaf(sw, r[lane], a0);
      a1 = fmaf(sw, r[lane + 32], a1);
      if (alpha && lane == 0) alpha[i] = w;
    }
    float* o = agg + (size_t)k * H;
    if (accumulate) { o[lane] += scale * a0; o[lane + 32] += scale * a1; }
    else { o[lane] = scale * a0; o[lane + 32] = scale * a1; }
  }
}

}  // namespace

extern "C" int tmpnn_gat_aggregate_dets(const tmpnn_graph* g, const tmpnn_index* ix, const float* h, int ldh, int col,
                                        const float* w_att, const float* a, int head, int num_heads, float* hatt,
                                        float* escore, float* agg, float* alpha, void* stream) {
  TMPNN_REQUIRE(g && ix && h && w_att && a && hatt && escore && agg, "null argument");
  TMPNN_REQUIRE(ldh % 4 == 0 && col % 4 == 0 && head >= 0 && head < num_heads, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  k_gat_project<<<TMPNN_SM_COUNT * 2, 256, 0, st>>>(ix->n_dets, ix->det_rows, h, ldh, col, w_att, hatt, g->phys);
  TMPNN_LAUNCH_CHECK();
  dim3 grid(max(1, min(tmpnn_div_up(g->cap_rows, 16), TMPNN_SM_COUNT * 8 / max(1, min(g->num_seqs, 8)))), g->num_seqs);
  k_gat_edge_score<<<grid, 256, 0, st>>>(g->n_rows, g->src, g->dst, g->cap_rows, ix->det_of_row, hatt, a, escore);
  TMPNN_LAUNCH_CHECK();
  k_gat_aggregate<<<TMPNN_SM_COUNT * 8, 256, 0, st>>>(ix->n_dets, ix->seg_ptr, ix->inc, escore, h, ldh, col,
                                                     1.0f / (float)num_heads, head > 0, agg, alpha, g->phys);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
