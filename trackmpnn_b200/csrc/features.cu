// features.cu -- feature construction from raw detections: the step in front of the hot path.
//
// Reference dataset/kitti_mot.py:545-566 / dataset/bdd100k_mot.py:530-551 without the visual block:
//   [one-hot category | score, xc, yc, w, h | sin, cos(pi (frame mod fr_range) / fr_range)], then (x - mean) / std.
// One thread per detection row; the 16-float detection record is read as four 128-bit loads.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_build_features(const float* __restrict__ bbox, int nd, int ncat, int use_2d, int use_temp, float fr_range,
                 const float* __restrict__ mean, const float* __restrict__ stdv, float* __restrict__ x, int ldx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += gridDim.x * blockDim.x) {
    const float4* r = reinterpret_cast<const float4*>(bbox + (size_t)i * 16);
    const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r3 = __ldg(r + 3);
    // [fr, trk, cat_id, alpha | x1, y1, x2, y2 | h, w, l, x | y, z, rotation_y, score]
    float* o = x + (size_t)i * ldx;
    const int cat = (int)r0.z - 1;
    int c = 0;
    for (; c < ncat; ++c) o[c] = ((c == cat ? 1.0f : 0.0f) - mean[c]) / stdv[c];
    if (use_2d) {
      const float f[5] = {r3.w, (r1.x + r1.z) / 2.0f, (r1.y + r1.w) / 2.0f, r1.z - r1.x, r1.w - r1.y};
#pragma unroll
      for (int k = 0; k < 5; ++k, ++c) o[c] = (f[k] - mean[c]) / stdv[c];
    }
    if (use_temp) {
      // numpy: np.mod(frames, fr_range) * np.pi / fr_range in float32 (the result of mod takes the divisor's sign)
      float m = fmodf(r0.x, fr_range);
      if (m != 0.0f && (m < 0.0f) != (fr_range < 0.0f)) m += fr_range;
      const float a = m * 3.14159265358979323846f / fr_range;
      o[c] = (sinf(a) - mean[c]) / stdv[c];
      ++c;
      o[c] = (cosf(a) - mean[c]) / stdv[c];
    }
  }
}

}  // namespace

extern "C" int tmpnn_build_features(const float* bbox_pred, int n_dets, int ncat, int use_2d, int use_temp, int fr_range,
                                    const float* mean, const float* std, float* x, int ldx, void* stream) {
  TMPNN_REQUIRE(bbox_pred && mean && std && x, "null argument");
  TMPNN_REQUIRE(ncat > 0 && fr_range != 0 && ldx >= ncat + (use_2d ? 5 : 0) + (use_temp ? 2 : 0), "bad shape");
  TMPNN_REQUIRE(((uintptr_t)bbox_pred & 15) == 0, "bbox_pred must be 16-byte aligned");
  if (n_dets <= 0) return TMPNN_OK;
  const int blocks = min(tmpnn_div_up(n_dets, 256), TMPNN_SM_COUNT * 8);
  k_build_features<<<blocks, 256, 0, (cudaStream_t)stream>>>(bbox_pred, n_dets, ncat, use_2d, use_temp, (float)fr_range, mean,
                                                            std, x, ldx);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
