// Generic exclusive prefix sum over a device-sized int32 array: reduce per 2048-item chunk,
// scan the chunk sums in one CTA, then rescan each chunk with its offset (3 launches).
#pragma once
#include "common.cuh"

namespace {
constexpr int SCAN_CHUNK = 2048;  // 512 threads x 4 items

// ---- generic exclusive scan over a device-sized int array (3 launches) --------------------
__global__ void __launch_bounds__(512) k_scan_reduce(const int32_t* __restrict__ in, const int32_t* __restrict__ n_dev,
                                                     int n_mul, int n_add, int32_t* __restrict__ sums) {
  __shared__ int sm[33];
  const int n = n_dev ? (*n_dev) * n_mul + n_add : n_add;
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
  int v = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) v += (base + q < n) ? in[base + q] : 0;
  int total;
  block_exclusive_scan(v, sm, &total);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_sums(int32_t* __restrict__ sums, int nblocks) {
  __shared__ int sm[33];
  int carry = 0;
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = i < nblocks ? sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, sm, &total);
    if (i < nblocks) sums[i] = carry + ex;
    carry += total;
  }
}

__global__ void __launch_bounds__(512) k_scan_down(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                   const int32_t* __restrict__ n_dev, int n_mul, int n_add,
                                                   const int32_t* __restrict__ sums) {
  __shared__ int sm[33];
  const int n = n_dev ? (*n_dev) * n_mul + n_add : n_add;
  const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
  int x[4];
  int v = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    x[q] = (base + q < n) ? in[base + q] : 0;
    v += x[q];
  }
  int total;
  int ex = block_exclusive_scan(v, sm, &total) + sums[blockIdx.x];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (base + q <= n) out[base + q] = ex;  // out[n] = grand total
    ex += x[q];
  }
}


// n = (*n_dev) * n_mul + n_add when n_dev != NULL, else n_add; n_max bounds it for grid sizing.
// out[0..n] (n+1 entries, out[n] = total).  sums needs tmpnn_div_up(n_max + 1, SCAN_CHUNK) ints.
inline cudaError_t scan_exclusive(const int32_t* in, int32_t* out, const int32_t* n_dev, int n_mul, int n_add, long long n_max,
                          int32_t* sums, cudaStream_t st) {
  const int nchunks = tmpnn_div_up(n_max + 1, SCAN_CHUNK);
  k_scan_reduce<<<nchunks, 512, 0, st>>>(in, n_dev, n_mul, n_add, sums);
  k_scan_sums<<<1, 1024, 0, st>>>(sums, nchunks);
  k_scan_down<<<nchunks, 512, 0, st>>>(in, out, n_dev, n_mul, n_add, sums);
  return cudaGetLastError();
}
}  // namespace
