// Does cp.async.bulk.tensor.2d ... tile::gather4 place four arbitrary 128-byte rows of a [rows x 128] fp16 tensor into shared
// memory in the SWIZZLE_128B K-major layout the UMMA descriptors of the edge kernel read (chunk c of row r at
// (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4))?   nvcc -arch=sm_100a gather4_test.cu -o gather4_test
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, const int* rows, uint16_t* out, int col0) {
  __shared__ __align__(1024) unsigned char sm[2048];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t sm_u = (uint32_t)__cvta_generic_to_shared(sm), bar_u = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = 0xEE;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u), "r"(1024u) : "memory");
    // rows 0-3 of the tile, then rows 4-7 (second gather4 lands 512 B further)
    for (int q = 0; q < 2; ++q)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                   ::"r"(sm_u + 512 * q), "l"(&tm), "r"(bar_u), "r"(col0), "r"(rows[4 * q]), "r"(rows[4 * q + 1]), "r"(rows[4 * q + 2]), "r"(rows[4 * q + 3])
                   : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar_u) : "memory");
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = ((uint16_t*)sm)[i];
}
int main() {
  const int R = 1000, C = 128;   // 128 fp16 per row = 256 B: [hi 64 | lo 64]
  std::vector<uint16_t> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 128 + c);   // unique (mod 65536) per (row, col)
  uint16_t* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  for (int box1 = 1; box1 <= 4; box1 += 3) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}, strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box1}, es[2] = {1, 1};
    CUresult rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {64, %d}: encode rc=%d\n", box1, (int)rc);
    if (rc) continue;
    int hr[8] = {17, 900, 3, 512, 77, 78, 999, 0};
    int* dr; cudaMalloc(&dr, 32); cudaMemcpy(dr, hr, 32, cudaMemcpyHostToDevice);
    uint16_t* dout; cudaMalloc(&dout, 2048); cudaMemset(dout, 0, 2048);
    for (int col0 = 0; col0 <= 64; col0 += 64) {
      k<<<1, 128>>>(tm, dr, dout, col0);
      cudaError_t e = cudaDeviceSynchronize();
      printf("  col0 %d: kernel %s\n", col0, cudaGetErrorString(e));
      if (e) break;
      std::vector<uint16_t> o(1024); cudaMemcpy(o.data(), dout, 2048, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) for (int e2 = 0; e2 < 8; ++e2) {
        const int off = ((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) / 2 + e2;
        const uint16_t want = (uint16_t)(hr[r] * 128 + col0 + c * 8 + e2);
        if (o[off] != want) { if (bad < 4) printf("    row %d chunk %d elem %d: got %u want %u\n", r, c, e2, o[off], want); ++bad; }
      }
      printf("  col0 %d: %s (%d mismatches)\n", col0, bad ? "MISMATCH" : "layout matches sw128", bad);
    }
  }
  return 0;
}
