// Shared helpers for libtmpnn_sm100a.so (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tmpnn.h"

#define TMPNN_SM_COUNT 148  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of it

int tmpnn_set_error(int code, const char* fmt, ...);
// mp_step_tc3.cu: tile table + the re-staged tcgen05 edge kernel (endpoints already prepared by k_det_prepare)
int tmpnn_edge_tc3_launch(const tmpnn_graph* g, const tmpnn_index* ix, const float* h_in, float* h_out, int ldh, int group,
                          int num_groups, int concat, const void* edge_image, const float* det_img, const float* det_p,
                          void* tile_table, cudaStream_t st);

#define TMPNN_CUDA_TRY(expr)                                                                   \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return tmpnn_set_error(TMPNN_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
  } while (0)

#define TMPNN_LAUNCH_CHECK() TMPNN_CUDA_TRY(cudaGetLastError())

#define TMPNN_REQUIRE(cond, msg)                                          \
  do {                                                                    \
    if (!(cond)) return tmpnn_set_error(TMPNN_E_BADARG, "%s: %s", __func__, msg); \
  } while (0)

// Segment table of one slab, written by tmpnn_index_build_structured at the start of its scratch2 (graph_index.cu) and
// read by the block-structured aggregation (mp_step.cu): the window graph as a chain [dets][edges][dets]...
constexpr int MAXSEG = 128;  // segments per slab (2 per frame in the window)
constexpr int MAXE = 64;     // edge segments per slab
struct SlabSegs {            // per slab, in global scratch
  int32_t nedge;             // number of edge segments (<= MAXE)
  int32_t nseg;              // number of segments
  int32_t start[MAXSEG + 1]; // slab-local first row of segment q; start[nseg] = n_rows
  int32_t eord[MAXSEG];      // ordinal among the edge segments, -1 for detection segments
};

static inline int tmpnn_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float tmpnn_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// 128-bit read-only load (h rows are 256 B = 16 lanes x float4)
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exclusive scan of one int per thread across a CTA (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix; *total gets the CTA sum.  smem must hold 33 ints.
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // smem may still be read from a previous call
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nw ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  *total = smem[32];
  return smem[warp] + inc - v;
}
