// hungarian.cu -- Hungarian association on the device (reference utils/graph.py:33-93, driven by
// :247-249 and :433-435).
//
// The reference builds, for every timestep t of the window, a small dense cost matrix
// C[|prev|, |cur|] (cur = detections at t, prev = still unassociated sources of the edges into cur,
// C = scores[e, 0] = 1 - p for an existing edge, 100.0 otherwise) and hands it to
// scipy.optimize.linear_sum_assignment; pairs with C <= 0.5 become associations, and associations
// made at an earlier t of the same call remove rows from later matrices.  scipy is an un-vendored
// third-party dependency of the reference (Pipfile:14, unpinned; 1.18 in this image).  Its solver is
// the rectangular shortest-augmenting-path algorithm of D. F. Crouse, "On implementing 2D rectangular
// assignment algorithms", IEEE T-AES 52(4), 2016, which is restated here step by step -- including the
// transposition of tall matrices, the reverse-ordered `remaining` list with swap-removal, and the
// tie-break "among equal shortest path costs prefer an unassigned column, the last one seen" -- so
// that equal-cost optima (the 100.0 fillers, saturated scores) resolve to the same assignment.
// All path arithmetic is fp64 like scipy's (the fp32 costs convert exactly).
//
// One warp per sequence (timesteps are sequential, matrices are <= a few hundred on a side); the inner
// scan over the remaining columns is spread over the 32 lanes with an order-aware reduction.
#include <math.h>

#include "common.cuh"

namespace {

struct LsapWork {   // per-solve scratch (global memory, L1/L2 resident), sized for n_max x n_max
  double* u;        // [n_max]
  double* v;        // [n_max]
  double* spc;      // [n_max]  shortest path costs
  int32_t* path;    // [n_max]
  int32_t* col4row; // [n_max]
  int32_t* row4col; // [n_max]
  int32_t* remaining; // [n_max]
  unsigned char* SR; // [n_max]
  unsigned char* SC; // [n_max]
};

__host__ __device__ inline size_t lsap_work_bytes(int n_max) {
  return ((size_t)n_max * (3 * sizeof(double) + 4 * sizeof(int32_t) + 2) + 127) / 64 * 64;
}

__device__ inline LsapWork lsap_carve(unsigned char* p, int n_max) {
  LsapWork w;
  w.u = reinterpret_cast<double*>(p); p += sizeof(double) * n_max;
  w.v = reinterpret_cast<double*>(p); p += sizeof(double) * n_max;
  w.spc = reinterpret_cast<double*>(p); p += sizeof(double) * n_max;
  w.path = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * n_max;
  w.col4row = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * n_max;
  w.row4col = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * n_max;
  w.remaining = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * n_max;
  w.SR = p; p += n_max;
  w.SC = p;
  return w;
}

// Solves min sum M[i][col4row[i]] for an nr x nc matrix with nr <= nc (row-major fp32, ld = nc).
// Executed by one full warp; on return w.col4row[i] is the column of row i.
__device__ void lsap_warp(const float* __restrict__ M, int nr, int nc, LsapWork w) {
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  for (int i = lane; i < nr; i += 32) { w.u[i] = 0.0; w.col4row[i] = -1; }
  for (int j = lane; j < nc; j += 32) { w.v[j] = 0.0; w.path[j] = -1; w.row4col[j] = -1; }
  __syncwarp();
  for (int cur_row = 0; cur_row < nr; ++cur_row) {
    // ---- augmenting path from cur_row ----
    double min_val = 0.0;
    int num_remaining = nc;
    for (int it = lane; it < nc; it += 32) { w.remaining[it] = nc - it - 1; w.SC[it] = 0; w.spc[it] = INFINITY; }
    for (int i = lane; i < nr; i += 32) w.SR[i] = 0;
    __syncwarp();
    int sink = -1, i = cur_row;
    while (sink == -1) {
      if (lane == 0) w.SR[i] = 1;
      const double ui = w.u[i];
      const float* mrow = M + (size_t)i * nc;
      double lmin = INFINITY;
      int lF = 0x7fffffff, lU = -1;
      for (int it = lane; it < num_remaining; it += 32) {
        const int j = w.remaining[it];
        const double r = min_val + (double)mrow[j] - ui - w.v[j];
        double s = w.spc[j];
        if (r < s) { w.path[j] = i; w.spc[j] = r; s = r; }
        const bool unassigned = w.row4col[j] == -1;
        if (s < lmin) { lmin = s; lF = it; lU = unassigned ? it : -1; }
        else if (s == lmin && unassigned) lU = it;
      }
      double m = lmin;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_xor_sync(FULL, m, o));
      int F = lmin == m ? lF : 0x7fffffff, U = lmin == m ? lU : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        F = min(F, __shfl_xor_sync(FULL, F, o));
        U = max(U, __shfl_xor_sync(FULL, U, o));
      }
      // the sequential scan keeps the first column reaching the minimum and then replaces it by every
      // later unassigned column of equal cost: the last unassigned one wins, else the first one
      const int index = U >= 0 ? U : F;
      min_val = m;
      if (!(m < INFINITY)) { sink = -2; break; }  // infeasible: cannot happen with finite costs
      __syncwarp();
      const int j = w.remaining[index];
      if (w.row4col[j] == -1) sink = j; else i = w.row4col[j];
      __syncwarp();
      if (lane == 0) {
        w.SC[j] = 1;
        w.remaining[index] = w.remaining[num_remaining - 1];
      }
      --num_remaining;
      __syncwarp();
    }
    if (sink < 0) return;
    // ---- dual variables ----
    if (lane == 0) w.u[cur_row] += min_val;
    for (int r = lane; r < nr; r += 32)
      if (w.SR[r] && r != cur_row) w.u[r] += min_val - w.spc[w.col4row[r]];
    for (int j = lane; j < nc; j += 32)
      if (w.SC[j]) w.v[j] -= min_val - w.spc[j];
    __syncwarp();
    // ---- augment ----
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = w.path[j];
        w.row4col[j] = r;
        const int t = w.col4row[r];
        w.col4row[r] = j;
        j = t;
        if (r == cur_row) break;
      }
    }
    __syncwarp();
  }
}

// ---- stand-alone batched solver (tests, and anything else that wants scipy's answer on the device) ----
__global__ void __launch_bounds__(32) k_lsap_batch(const float* __restrict__ C, int nr, int nc, int batch,
                                                   int32_t* __restrict__ col_of_row, float* __restrict__ mt,
                                                   unsigned char* __restrict__ scratch, int n_max) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  const int lane = threadIdx.x;
  const float* Cb = C + (size_t)b * nr * nc;
  LsapWork w = lsap_carve(scratch + (size_t)b * lsap_work_bytes(n_max), n_max);
  int32_t* out = col_of_row + (size_t)b * nr;
  if (nr == 0 || nc == 0) return;
  if (nc < nr) {  // tall: solve the transpose (scipy does the same), rows of M = columns of C
    float* M = mt + (size_t)b * nr * nc;
    for (int q = lane; q < nr * nc; q += 32) M[(size_t)(q % nc) * nr + q / nc] = Cb[q];
    __syncwarp();
    lsap_warp(M, nc, nr, w);
    for (int i = lane; i < nr; i += 32) out[i] = -1;
    __syncwarp();
    for (int j = lane; j < nc; j += 32) out[w.col4row[j]] = j;
  } else {
    lsap_warp(Cb, nr, nc, w);
    for (int i = lane; i < nr; i += 32) out[i] = w.col4row[i];
  }
}

// ---- the association pass: one warp per sequence ----
__global__ void __launch_bounds__(32)
k_associate_hungarian(tmpnn_graph g, tmpnn_index ix, const float* __restrict__ cost, const int32_t* __restrict__ active,
                      int n_max, unsigned char* __restrict__ scratch, size_t per_seq, int only_t, int t_only,
                      float threshold) {
  const int s = blockIdx.x, lane = threadIdx.x;
  const unsigned FULL = 0xffffffffu;
  if (active && !active[s]) return;
  const int n = g.n_rows[s];
  if (n == 0) return;
  const size_t base = (size_t)s * g.cap_rows;
  const int k0 = ix.seq_det_ptr[s], k1 = ix.seq_det_ptr[s + 1], nd = k1 - k0;
  if (nd > n_max) {
    if (lane == 0) atomicOr(g.status, TMPNN_FLAG_WALK_CAPACITY);
    return;
  }
  unsigned char* p = scratch + (size_t)s * per_seq;
  LsapWork w = lsap_carve(p, n_max); p += lsap_work_bytes(n_max);
  int32_t* cur = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * n_max;      // detection-list positions at t
  int32_t* prev = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * n_max;     // ... of the candidate sources
  int32_t* prevpos = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * n_max;  // position in prev, -1 if not a candidate
  float* M = reinterpret_cast<float*>(p);                                          // [min side][max side]
  const int t_first = only_t ? t_only : g.ts[base], t_last = only_t ? t_only : g.ts[base + n - 1];
  for (int t = t_first; t <= t_last; ++t) {
    // cur = detections at t, ascending (ordered compaction with ballots)
    int ncur = 0, has_edges = 0;
    for (int q0 = 0; q0 < nd; q0 += 32) {
      const int q = q0 + lane;
      const bool is = q < nd && g.ts[ix.det_rows[k0 + q]] == t;
      const unsigned bal = __ballot_sync(FULL, is);
      if (is) cur[ncur + __popc(bal & ((1u << lane) - 1u))] = q;
      ncur += __popc(bal);
    }
    if (ncur == 0) continue;
    for (int q = lane; q < nd; q += 32) prevpos[q] = -1;
    __syncwarp();
    // candidate sources: still unassociated endpoints of the edges into cur
    for (int c = 0; c < ncur; ++c) {
      const int k = k0 + cur[c];
      const int e0 = ix.seg_ptr[2 * k], e1 = ix.seg_ptr[2 * k + 1];
      if (e1 > e0) has_edges = 1;
      for (int i = e0 + lane; i < e1; i += 32) {
        const int a = g.src[ix.inc[i]];
        if (g.ass[base + a] == -1) prevpos[ix.det_of_row[base + a] - k0] = 0;
      }
    }
    if (!has_edges) continue;
    __syncwarp();
    int nprev = 0;
    for (int q0 = 0; q0 < nd; q0 += 32) {
      const int q = q0 + lane;
      const bool is = q < nd && prevpos[q] == 0;
      const unsigned bal = __ballot_sync(FULL, is);
      if (is) {
        const int pos = nprev + __popc(bal & ((1u << lane) - 1u));
        prev[pos] = q;
        prevpos[q] = pos;
      }
      nprev += __popc(bal);
    }
    __syncwarp();
    if (nprev == 0) continue;
    // cost matrix in the orientation scipy solves: rows = the smaller side (prev unless it is the taller one)
    const bool transposed = ncur < nprev;
    const int nr = transposed ? ncur : nprev, nc = transposed ? nprev : ncur;
    for (int q = lane; q < nr * nc; q += 32) M[q] = 100.0f;
    __syncwarp();
    for (int c = 0; c < ncur; ++c) {
      const int k = k0 + cur[c];
      const int e0 = ix.seg_ptr[2 * k], e1 = ix.seg_ptr[2 * k + 1];
      for (int i = e0 + lane; i < e1; i += 32) {
        const int e = ix.inc[i];
        const int pi = prevpos[ix.det_of_row[base + g.src[e]] - k0];
        if (pi >= 0) {
          const float cv = cost ? cost[e] : 1.0f - g.score[e];
          if (transposed) M[(size_t)c * nc + pi] = cv; else M[(size_t)pi * nc + c] = cv;
        }
      }
    }
    __syncwarp();
    lsap_warp(M, nr, nc, w);
    // accept pairs with C <= 0.5 (utils/graph.py:87-91)
    for (int r = lane; r < nr; r += 32) {
      const int c = w.col4row[r];
      if (c < 0) continue;
      if (M[(size_t)r * nc + c] > threshold) continue;
      const int pi = transposed ? c : r, ci = transposed ? r : c;
      g.ass[ix.det_rows[k0 + prev[pi]]] = g.det[ix.det_rows[k0 + cur[ci]]];
    }
    __syncwarp();
  }
}

__global__ void k_reset_ass_all(tmpnn_graph g, const int32_t* __restrict__ active) {
  const int s = blockIdx.y;
  const int n = (active && !active[s]) ? 0 : g.n_rows[s];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) g.ass[(size_t)s * g.cap_rows + r] = -1;
}

}  // namespace

extern "C" size_t tmpnn_lsap_scratch_bytes(int batch, int nr, int nc) {
  const int n_max = nr > nc ? nr : nc;
  return (size_t)batch * lsap_work_bytes(n_max) + (size_t)batch * nr * nc * sizeof(float);
}

extern "C" int tmpnn_lsap_solve(const float* cost, int nr, int nc, int batch, int32_t* col_of_row, void* scratch, void* stream) {
  TMPNN_REQUIRE(cost && col_of_row && scratch && nr >= 0 && nc >= 0 && batch >= 0, "bad argument");
  if (batch == 0 || nr == 0 || nc == 0) return TMPNN_OK;
  const int n_max = nr > nc ? nr : nc;
  unsigned char* sc = (unsigned char*)scratch;
  float* mt = (float*)(sc + (size_t)batch * lsap_work_bytes(n_max));
  k_lsap_batch<<<batch, 32, 0, (cudaStream_t)stream>>>(cost, nr, nc, batch, col_of_row, mt, sc, n_max);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

static size_t hungarian_per_seq(int max_dets) {
  size_t b = lsap_work_bytes(max_dets) + 3 * sizeof(int32_t) * (size_t)max_dets + sizeof(float) * (size_t)max_dets * max_dets;
  return (b + 63) / 64 * 64;
}

extern "C" size_t tmpnn_hungarian_scratch_bytes(int num_seqs, int max_dets) {
  return (size_t)num_seqs * hungarian_per_seq(max_dets) + 64;
}

extern "C" int tmpnn_graph_associate_hungarian(const tmpnn_graph* g, const tmpnn_index* ix, const float* cost,
                                               const int32_t* active, int max_dets, int only_t, int t, float threshold,
                                               void* scratch, void* stream) {
  TMPNN_REQUIRE(g && ix && scratch && max_dets > 0, "bad argument");
  TMPNN_REQUIRE(((uintptr_t)scratch & 7) == 0, "scratch must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (!only_t) {
    dim3 grid(max(1, min(tmpnn_div_up(g->cap_rows, 1024), 64)), g->num_seqs);
    k_reset_ass_all<<<grid, 256, 0, st>>>(*g, active);
    TMPNN_LAUNCH_CHECK();
  }
  k_associate_hungarian<<<g->num_seqs, 32, 0, st>>>(*g, *ix, cost, active, max_dets, (unsigned char*)scratch,
                                                   hungarian_per_seq(max_dets), only_t, t, threshold);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
