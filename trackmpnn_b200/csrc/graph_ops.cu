// graph_ops.cu -- device-side sliding-window graph bookkeeping (reference utils/graph.py).
//
//   tmpnn_graph_associate   greedy association / teacher forcing   utils/graph.py:229-268, 437-454
//   tmpnn_graph_append      update_graph + initialize_graph        utils/graph.py:96-186, 270-327
//   tmpnn_graph_decode      track-id walk + deletion mask          utils/graph.py:456-512
//   tmpnn_graph_prune_mask  prune_graph                            utils/graph.py:361-377
//   tmpnn_graph_compact     order-preserving row deletion          utils/graph.py:379-387, 514-520
//   tmpnn_ypred_*, tmpnn_coo_*  boundary converters to the reference's tensors
//
// Everything is integer / index work on the structure-of-arrays slabs of tmpnn.h; results
// must be bit-exact with the reference.  The graph never returns to the host between frames:
// row counts and detection counts stay in device memory.
#include <limits.h>

#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int ROWS_PER_BLOCK = 1024;
// The decode walk keeps six ints per detection row of the window.  They live in shared memory up to
// WALK_SMEM_MAX detection rows (196 KB of the 227 KB an sm_100a CTA can have; the launch sizes the buffer from the
// caller's bound max_dets, so ordinary windows keep several CTAs per SM) and in a global-memory scratch beyond that --
// same sort, same sequential walk (utils/graph.py:456-490), no capacity limit.
constexpr int WALK_SMEM_DEFAULT = 4096;
constexpr int WALK_SMEM_MAX = 8192;

__device__ __forceinline__ bool seq_off(const int32_t* active, int s) { return active && !active[s]; }

// ---- converters ------------------------------------------------------------------------------
__global__ void k_ypred_unpack(const int64_t* __restrict__ y, int n, int32_t* ts, int32_t* det, int32_t* ass) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    ts[i] = (int32_t)y[3 * (size_t)i];
    det[i] = (int32_t)y[3 * (size_t)i + 1];
    ass[i] = (int32_t)y[3 * (size_t)i + 2];
  }
}
__global__ void k_ypred_pack(const int32_t* __restrict__ ts, const int32_t* __restrict__ det,
                             const int32_t* __restrict__ ass, int n, int64_t* y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    y[3 * (size_t)i] = ts[i];
    y[3 * (size_t)i + 1] = det[i];
    y[3 * (size_t)i + 2] = ass[i];
  }
}

__global__ void k_coo_count(const int32_t* __restrict__ ts, int n, int transpose, int32_t* cnt) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    cnt[i] = transpose ? (ts[i] < 0 ? 3 : 0) : (ts[i] < 0 ? 2 : 1);
}
__global__ void k_coo_write(const int32_t* __restrict__ ts, const int32_t* __restrict__ src,
                            const int32_t* __restrict__ dst, int n, int transpose, const int32_t* __restrict__ off,
                            int64_t* __restrict__ idx, float* __restrict__ val, int64_t nnz) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int64_t o = off[i];
    if (ts[i] < 0) {
      if (!transpose) {  // node_adj row e: +1 at src, -1 at dst (columns ascending: src < dst)
        if (o + 1 < nnz) {
          idx[o] = i; idx[nnz + o] = src[i]; val[o] = 1.f;
          idx[o + 1] = i; idx[nnz + o + 1] = dst[i]; val[o + 1] = -1.f;
        }
      } else {  // edge_adj = node_adj^T off the diagonal, plus I_edge
        if (o + 2 < nnz) {
          idx[o] = src[i]; idx[nnz + o] = i; val[o] = 1.f;
          idx[o + 1] = dst[i]; idx[nnz + o + 1] = i; val[o + 1] = -1.f;
          idx[o + 2] = i; idx[nnz + o + 2] = i; val[o + 2] = 1.f;
        }
      }
    } else if (!transpose && o < nnz) {
      idx[o] = i; idx[nnz + o] = i; val[o] = 1.f;  // I_node
    }
  }
}
__global__ void k_edges_from_coo(const int64_t* __restrict__ idx, const float* __restrict__ val, int64_t nnz, int n,
                                 int32_t* src, int32_t* dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx[i], c = idx[nnz + i];
    if (r == c || r < 0 || r >= n) continue;
    const float v = val[i];
    if (v > 0.f) src[r] = (int32_t)c;
    else if (v < 0.f) dst[r] = (int32_t)c;
  }
}

// infer.py:54-57, 77-80 (--no-tp-classifier): detection rows count as true positives, scores (0, 1)
__global__ void k_force_det_scores(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows,
                                   float* __restrict__ score) {
  const int nd = *n_dets;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nd; k += gridDim.x * blockDim.x) score[det_rows[k]] = 1.0f;
}

// moves the bits `from` of the sticky status word to `to` (a handled condition becomes a note)
__global__ void k_status_ack(int32_t* status, int from, int to) {
  const int v = *status;
  if (v & from) *status = (v & ~from) | to;
}

// ---- association -----------------------------------------------------------------------------
__global__ void k_reset_ass(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active, int cap_rows,
                            int32_t* __restrict__ ass) {
  const int s = blockIdx.y, n = seq_off(active, s) ? 0 : n_rows[s];
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    ass[(size_t)s * cap_rows + r] = -1;
}

// One warp per detection row.  Greedy: lexicographic best over the positive future edges of
// (timestamp of the far end ascending == nearest edge block, p descending, row ascending ==
// np.argmax's first maximum).
__global__ void __launch_bounds__(256) k_associate(const int32_t* __restrict__ n_dets, const int32_t* __restrict__ det_rows,
                                                   const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ inc,
                                                   const int32_t* __restrict__ ts, const int32_t* __restrict__ det,
                                                   const int32_t* __restrict__ dst, const int32_t* __restrict__ label,
                                                   const float* __restrict__ score, int cap_rows, int mode,
                                                   int32_t* __restrict__ ass, int32_t* __restrict__ status) {
  const int nd = *n_dets;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int k = blockIdx.x * wpb + (threadIdx.x >> 5); k < nd; k += gridDim.x * wpb) {
    const int row = det_rows[k];
    const int base = (row / cap_rows) * cap_rows;
    const int f0 = seg_ptr[2 * k + 1], f1 = seg_ptr[2 * k + 2];
    int result = -1;
    if (mode == 1) {  // teacher forcing
      if (label[row] == 1) {
        int cnt = 0, e_pos = INT_MAX;
        for (int i = f0 + lane; i < f1; i += 32) {
          const int e = inc[i];
          if (label[e] != 0) { ++cnt; e_pos = min(e_pos, e); }
        }
        cnt = warp_sum_i(cnt);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e_pos = min(e_pos, __shfl_xor_sync(0xffffffffu, e_pos, o));
        if (cnt == 1) result = det[base + dst[e_pos]];
        else if (cnt > 1 && lane == 0) atomicOr(status, TMPNN_FLAG_MULTI_GT_EDGE);
      } else {
        result = det[row];  // self-assignment keeps a false positive inactive
      }
    } else if (score[row] >= 0.5f) {
      int b_ts = INT_MAX, b_e = INT_MAX;
      float b_p = -1.f;
      // four strides of the list in flight (the walk is a chain of dependent loads inc -> score; one stride at a time left
      // the kernel latency bound); the best candidate is a minimum of a total order, so the visiting order is free
      for (int i0 = f0 + lane; i0 < f1; i0 += 128) {
        int ev[4];
        float pv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) ev[q] = i0 + 32 * q < f1 ? inc[i0 + 32 * q] : -1;
#pragma unroll
        for (int q = 0; q < 4; ++q) pv[q] = ev[q] >= 0 ? score[ev[q]] : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = ev[q];
          const float pe = pv[q];
          if (!(pe >= 0.5f)) continue;
          const int d = base + dst[e];
          if (!(score[d] >= 0.5f)) continue;
          const int t = ts[d];
          if (t < b_ts || (t == b_ts && (pe > b_p || (pe == b_p && e < b_e)))) { b_ts = t; b_p = pe; b_e = e; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const int t = __shfl_xor_sync(0xffffffffu, b_ts, o);
        const float p = __shfl_xor_sync(0xffffffffu, b_p, o);
        const int e = __shfl_xor_sync(0xffffffffu, b_e, o);
        if (t < b_ts || (t == b_ts && (p > b_p || (p == b_p && e < b_e)))) { b_ts = t; b_p = p; b_e = e; }
      }
      if (b_e != INT_MAX) result = det[base + dst[b_e]];
    }
    if (lane == 0) ass[row] = result;
  }
}

// ---- append / initialise ----------------------------------------------------------------------
// desc[s*DESC + ..]: 0 n_old, 1 A (or N0), 2 Nt (or N1), 3 offset of the new frame's ids in frame_dets,
//   4 first slot in new_det_rows, 5 kind (0 none, 1 append, 2 init), 6 offset of t0's ids, 7 t (or t1), 8 t0
constexpr int DESC = 12;
// Inference-mode active set (detection, unassociated, p >= 0.5) over a grid: per-block counts, scanned by the plan
// kernel, then written in row order.  (A 20-frame x 200-detection window has millions of rows per slab; the plan
// kernel's own one-CTA scan took 14 % of a frame there.  Training graphs are small and keep the in-CTA scan.)
__device__ __forceinline__ int active_flag(const tmpnn_graph& g, size_t row) {
  return g.ts[row] >= 0 && g.ass[row] == -1 && g.score[row] >= 0.5f;
}
__global__ void __launch_bounds__(256) k_active_count(tmpnn_graph g, int nblk, int32_t* __restrict__ blk_cnt) {
  __shared__ int sm[33];
  const int s = blockIdx.y, b = blockIdx.x, n = g.n_rows[s];
  if (b * ROWS_PER_BLOCK >= n) { if (threadIdx.x == 0) blk_cnt[s * nblk + b] = 0; return; }
  const size_t base = (size_t)s * g.cap_rows;
  const int r0 = b * ROWS_PER_BLOCK + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (r0 + q < n) c += active_flag(g, base + r0 + q);
  int total;
  block_exclusive_scan(c, sm, &total);
  if (threadIdx.x == 0) blk_cnt[s * nblk + b] = total;
}
__global__ void __launch_bounds__(256) k_active_write(tmpnn_graph g, int nblk, const int32_t* __restrict__ blk_off,
                                                      const int32_t* __restrict__ desc, int desc_stride,
                                                      int32_t* __restrict__ act) {
  __shared__ int sm[33];
  const int s = blockIdx.y, b = blockIdx.x;
  if (desc[desc_stride * s + 5] != 1) return;   // this sequence appends nothing from an active list
  const int n = desc[desc_stride * s + 0];       // rows before the append
  if (b * ROWS_PER_BLOCK >= n) return;
  const size_t base = (size_t)s * g.cap_rows;
  const int r0 = b * ROWS_PER_BLOCK + threadIdx.x * 4;
  int f[4], c = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[q] = (r0 + q < n) ? active_flag(g, base + r0 + q) : 0;
    c += f[q];
  }
  int total;
  int pos = blk_off[s * nblk + b] + block_exclusive_scan(c, sm, &total);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (f[q]) act[base + pos++] = r0 + q;
}
__global__ void __launch_bounds__(1024)
k_append_plan(tmpnn_graph g, tmpnn_frames fr, tmpnn_seq_state st, int has_state, const int32_t* __restrict__ t_dev,
              int start, int cur_win, int mode, int32_t* __restrict__ act, int32_t* __restrict__ desc,
              int32_t* __restrict__ n_new, int cap_new, int32_t* __restrict__ n_appended, int32_t* __restrict__ blk_cnt,
              int nblk) {
  __shared__ int sm[33];
  __shared__ int s_kind, s_t, s_t1, s_tprev;
  const int s = blockIdx.x;
  const size_t base = (size_t)s * g.cap_rows;
  const int32_t* fp = fr.frame_ptr + (size_t)s * (fr.t_max + 2);
  const int n = g.n_rows[s];
  if (threadIdx.x == 0) {
    int t = t_dev ? *t_dev : 0;
    int kind = 1, t1 = -1;
    if (has_state) {
      st.fresh[s] = 0;
      if (start) {
        kind = 2;
      } else if (st.phase[s] != 1 || t < st.skip_until[s] || t >= st.t_end[s]) {
        kind = 0;
        st.active[s] = 0;
      } else {
        // infer.py:64-69: "feats.size()[0] == 0 and states.size()[0] == 0" -- the rows the PREVIOUS executed iteration
        // added and the graph its decode left behind; the current frame's size does not enter
        kind = (n == 0 && st.last_new[s] == 0) ? 2 : 1;
        st.active[s] = 1;
        st.t_upto[s] = (t == st.t_end[s] - 1) ? st.t_end[s] : t - cur_win + 2;  // infer.py:82-87
      }
      if (kind == 2) {  // first two non-empty timesteps at or after t (utils/graph.py:120-133)
        int t0 = -1;
        for (int u = t; u <= fr.t_max; ++u) {
          if (fp[u + 1] > fp[u]) {
            if (t0 < 0) t0 = u; else { t1 = u; break; }
          }
        }
        if (t1 < 0) {  // "return None": nothing left to track
          kind = 0;
          st.phase[s] = 2;
          st.active[s] = 0;
        } else {
          int tn = fr.t_max;
          while (tn > 0 && fp[tn + 1] == fp[tn]) --tn;
          if (start) { st.t_end[s] = tn + 1; st.t_upto[s] = INT_MIN; }
          st.phase[s] = 1;
          st.skip_until[s] = t1 + 1;
          st.active[s] = 1;
          st.fresh[s] = 1;
          t = t0;
        }
      }
    }
    s_kind = kind; s_t = t; s_t1 = t1;
    s_tprev = -1;
  }
  __syncthreads();
  const int kind = s_kind, t = s_t;
  int32_t* d = desc + DESC * s;
  if (kind == 0) {
    if (threadIdx.x == 0) { d[5] = 0; n_appended[s] = 0; }
    return;
  }
  if (kind == 2) {
    if (threadIdx.x == 0) {
      const int t1 = s_t1;
      const int n0 = fp[t + 1] - fp[t], n1 = fp[t1 + 1] - fp[t1];
      const long long total = (long long)n0 + (long long)n0 * n1 + n1;
      int slot = -1;
      // deferred compaction keeps the slab's last row as the shared all-zero row
      bool ok = total <= g.cap_rows - (g.phys ? 1 : 0);
      if (ok) {
        slot = atomicAdd(&n_new[0], n0 + n1);
        if (slot + n0 + n1 > cap_new) ok = false;
      }
      if (!ok) {
        atomicOr(g.status, TMPNN_FLAG_ROW_CAPACITY);
        d[5] = 0; n_appended[s] = 0; g.n_rows[s] = 0;
      } else {
        atomicAdd(&n_new[1], n0 * n1);
        d[0] = 0; d[1] = n0; d[2] = n1; d[3] = fp[t1]; d[4] = slot; d[5] = 2; d[6] = fp[t]; d[7] = t1; d[8] = t;
        n_appended[s] = (int)total;
        g.n_rows[s] = (int)total;
        if (has_state) st.last_new[s] = (int)total;
      }
    }
    return;
  }
  // kind 1: active list
  if (mode == 1) {  // t_prev = latest detection timestamp before t (utils/graph.py:273)
    int m = -1;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const int v = g.ts[base + r];
      if (v < t) m = max(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(&s_tprev, m);
    __syncthreads();
  }
  const int tprev = s_tprev;
  int carry = 0;
  if (mode != 1 && blk_cnt) {
    // inference: the per-block counts of k_active_count become offsets for k_active_write
    for (int b0 = 0; b0 < nblk; b0 += blockDim.x) {
      const int b = b0 + threadIdx.x;
      const int v = b < nblk ? blk_cnt[s * nblk + b] : 0;
      int total;
      const int ex = block_exclusive_scan(v, sm, &total);
      if (b < nblk) blk_cnt[s * nblk + b] = carry + ex;
      carry += total;
    }
  } else
  for (int r0 = 0; r0 < n; r0 += 4 * 1024) {
    const int r = r0 + threadIdx.x * 4;
    int f[4], c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      f[q] = 0;
      if (r + q < n) {
        const int tsv = g.ts[base + r + q];
        if (tsv >= 0) {
          const bool un = g.ass[base + r + q] == -1;
          f[q] = mode == 1 ? (un || tsv == tprev) : (un && g.score[base + r + q] >= 0.5f);
        }
      }
      c += f[q];
    }
    int total;
    int pos = carry + block_exclusive_scan(c, sm, &total);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (f[q]) act[base + pos++] = r + q;
    carry += total;
  }
  if (threadIdx.x == 0) {
    const int A = carry;
    const int nt = (t >= 0 && t <= fr.t_max) ? fp[t + 1] - fp[t] : 0;
    const long long add = nt ? (long long)A * nt + nt : 0;
    int slot = -1;
    bool ok = n + add <= g.cap_rows;
    // deferred compaction: the new detections' state goes behind the dense part of the step's input buffer,
    // and the slab's last row stays the shared all-zero row
    if (g.phys) ok = n + add <= g.cap_rows - 1 && (long long)g.phys_end[s] + nt <= g.cap_rows - 1;
    if (ok && nt) {
      slot = atomicAdd(&n_new[0], nt);
      if (slot + nt > cap_new) ok = false;
    }
    if (!ok) {
      atomicOr(g.status, TMPNN_FLAG_ROW_CAPACITY);
      d[5] = 0; n_appended[s] = 0;
    } else if (nt == 0) {  // graph unchanged (utils/graph.py:284,295)
      d[5] = 0; n_appended[s] = 0;
      if (has_state) st.last_new[s] = 0;
    } else {
      if (has_state) st.last_new[s] = (int)add;
      atomicAdd(&n_new[1], A * nt);
      d[0] = n; d[1] = A; d[2] = nt; d[3] = fp[t]; d[4] = slot; d[5] = 1; d[6] = 0; d[7] = t;
      n_appended[s] = (int)add;
      g.n_rows[s] = n + (int)add;
    }
  }
}

__global__ void __launch_bounds__(256)
k_append_fill(tmpnn_graph g, tmpnn_frames fr, const int32_t* __restrict__ act, const int32_t* __restrict__ desc,
              float* __restrict__ h, int ldh, int32_t* __restrict__ new_det_rows, int32_t* __restrict__ new_det_x) {
  const int s = blockIdx.y;
  const int32_t* d = desc + DESC * s;
  const int kind = d[5];
  if (kind == 0) return;
  const int n_old = d[0], A = d[1], nt = d[2], f_off = d[3], slot = d[4], f0_off = d[6], tnew = d[7], t0 = d[8];
  const size_t base = (size_t)s * g.cap_rows;
  const int dp = fr.det_ptr[s];
  const int lead = kind == 2 ? A : 0;  // init: A = N0 detection rows come first
  const int n_edge = A * nt;
  const int total = lead + n_edge + nt;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // deferred compaction: new detections take the physical rows behind the dense part of the step's input
  // buffer, new edge rows all alias the slab's zero row
  const bool deferred = g.phys != nullptr;
  const int p_det0 = deferred ? (kind == 2 ? lead : g.phys_end[s]) : 0;  // physical slab row of detection j of frame t
  const int32_t zrow = (int32_t)(base + g.cap_rows - 1);
  for (int k = tid; k < total; k += nth) {
    const size_t row = base + n_old + k;
    int ts = -1, det = -1, sr = -1, ds = -1, lab = 0;
    int32_t ph = (int32_t)row, ps = -1, pd = -1;
    if (k < lead) {  // detections of t0 (init only)
      det = fr.frame_dets[f0_off + k];
      ts = t0;
      if (deferred) ph = (int32_t)(base + k);
      new_det_rows[slot + k] = ph;
      new_det_x[slot + k] = dp + det;
      if (fr.det_track) lab = fr.det_track[dp + det] >= 0;
    } else if (k < lead + n_edge) {
      const int q = k - lead, a = q / nt, j = q % nt;
      sr = kind == 2 ? a : act[base + a];
      ds = n_old + lead + n_edge + j;
      if (deferred) {
        ph = zrow;
        ps = kind == 2 ? (int32_t)(base + a) : g.phys[base + sr];
        pd = (int32_t)(base + p_det0 + j);
      }
      if (fr.det_track && g.label) {
        const int da = kind == 2 ? fr.frame_dets[f0_off + a] : g.det[base + sr];
        const int tr_a = fr.det_track[dp + da];
        const int tr_j = fr.det_track[dp + fr.frame_dets[f_off + j]];
        lab = (tr_j != -1 && tr_a == tr_j) ? 1 : 0;
      }
    } else {
      const int j = k - lead - n_edge;
      det = fr.frame_dets[f_off + j];
      ts = tnew;
      if (deferred) ph = (int32_t)(base + p_det0 + j);
      new_det_rows[slot + lead + j] = ph;
      new_det_x[slot + lead + j] = dp + det;
      if (fr.det_track) lab = fr.det_track[dp + det] >= 0;
    }
    g.ts[row] = ts;
    g.det[row] = det;
    g.ass[row] = -1;
    g.src[row] = sr;
    g.dst[row] = ds;
    if (g.label) g.label[row] = lab;
    g.score[row] = 0.f;
    g.logit[row] = 0.f;
    if (deferred) { g.phys[row] = ph; g.psrc[row] = ps; g.pdst[row] = pd; }
  }
  // h = 0 for the new rows: one flat, fully coalesced range per sequence
  if (h && !deferred) {
    float4* p = reinterpret_cast<float4*>(h + (base + n_old) * ldh);
    const size_t n4 = (size_t)total * ldh / 4;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = tid; i < n4; i += nth) p[i] = z;
  }
}

// ---- decode ------------------------------------------------------------------------------------
struct WalkArrays {
  int32_t* key;   // detection id (sort key)
  int32_t* idx;   // position in det_rows
  int32_t* ts;
  int32_t* nxt;
  int32_t* yo;
  int32_t* flag;  // bit0: p >= 0.5, bit1: visited
};
__device__ __forceinline__ WalkArrays walk_arrays(int32_t* base, int cap) {
  return WalkArrays{base, base + cap, base + 2 * cap, base + 3 * cap, base + 4 * cap, base + 5 * cap};
}
__host__ __device__ inline int pow2_ceil(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

__global__ void __launch_bounds__(256)
k_decode(tmpnn_graph g, const int32_t* __restrict__ seq_det_ptr, const int32_t* __restrict__ det_rows,
         const int32_t* __restrict__ det_ptr, int32_t* __restrict__ y_out_track, int32_t* __restrict__ next_track_id,
         const int32_t* __restrict__ t_upto_seq, int t_upto_host, const int32_t* __restrict__ active, int ret_win,
         uint8_t* __restrict__ keep, int32_t* __restrict__ max_id_out, int smem_cap, int32_t* __restrict__ gwalk, int gwalk_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_unsorted, s_maxid;
  const int s = blockIdx.x;
  if (seq_off(active, s)) {
    if (threadIdx.x == 0) max_id_out[s] = -1;  // "keep everything"
    return;
  }
  const int k0 = seq_det_ptr[s], nd = seq_det_ptr[s + 1] - k0;
  const int t_upto = t_upto_seq ? t_upto_seq[s] : t_upto_host;
  const size_t base = (size_t)s * g.cap_rows;
  if (threadIdx.x == 0) { s_unsorted = 0; s_maxid = 0; }
  __syncthreads();
  const int p2 = pow2_ceil(nd);
  WalkArrays w;
  if (p2 <= smem_cap) {
    w = walk_arrays(reinterpret_cast<int32_t*>(smem_raw), smem_cap);
  } else if (gwalk && p2 <= gwalk_cap) {  // a window with more detection rows than shared memory holds: walk out of global memory
    w = walk_arrays(gwalk + (size_t)s * 6 * gwalk_cap, gwalk_cap);
  } else {
    if (threadIdx.x == 0) { atomicOr(g.status, TMPNN_FLAG_WALK_CAPACITY); max_id_out[s] = -1; }
    return;
  }
  for (int i = threadIdx.x; i < p2; i += blockDim.x) {
    if (i < nd) {
      const int v = g.det[det_rows[k0 + i]];
      w.key[i] = v;
      w.idx[i] = i;
      if (i + 1 < nd && g.det[det_rows[k0 + i + 1]] < v) s_unsorted = 1;
    } else {
      w.key[i] = INT_MAX;
      w.idx[i] = -1;
    }
  }
  __syncthreads();
  if (s_unsorted) {  // detection ids not in row order (time-reversed training streams): sort by id
    for (int k = 2; k <= p2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < p2; i += blockDim.x) {
          const int l = i ^ j;
          if (l > i) {
            const bool up = (i & k) == 0;
            if ((w.key[i] > w.key[l]) == up) {
              int a = w.key[i]; w.key[i] = w.key[l]; w.key[l] = a;
              a = w.idx[i]; w.idx[i] = w.idx[l]; w.idx[l] = a;
            }
          }
        }
        __syncthreads();
      }
  }
  const int dp = det_ptr[s];
  int local_max = 0;
  for (int q = threadIdx.x; q < nd; q += blockDim.x) {
    const int row = det_rows[k0 + w.idx[q]];
    const int tsv = g.ts[row];
    w.ts[q] = tsv;
    w.flag[q] = g.score[row] >= 0.5f ? 1 : 0;
    w.yo[q] = y_out_track[dp + w.key[q]];
    const int a = g.ass[row];
    int nx = -1;
    if (a >= 0) {  // binary search of the associated detection id
      int lo = 0, hi = nd - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int kv = w.key[mid];
        if (kv == a) { nx = mid; break; }
        if (kv < a) lo = mid + 1; else hi = mid - 1;
      }
    }
    w.nxt[q] = nx;
    if (tsv < t_upto) local_max = max(local_max, (int)(row - base) + 1);
  }
  atomicMax(&s_maxid, local_max);
  __syncthreads();
  if (threadIdx.x == 0) {  // the reference's sequential walk, utils/graph.py:459-490
    int next_id = next_track_id[s];
    for (int q = 0; q < nd; ++q) {
      const int f = w.flag[q];
      if (w.ts[q] >= t_upto || !(f & 1)) { w.flag[q] = f | 2; continue; }
      if (f & 2) continue;
      int cur = w.yo[q];
      if (cur == -1) cur = next_id++;
      int c = q;
      for (int guard = 0; guard <= nd; ++guard) {
        w.flag[c] |= 2;
        w.yo[c] = cur;
        const int nx = w.nxt[c];
        if (nx < 0) break;
        if (w.ts[c] >= t_upto && w.ts[nx] >= t_upto) break;
        c = nx;
      }
    }
    next_track_id[s] = next_id;
    max_id_out[s] = s_maxid;
  }
  __syncthreads();
  const int max_id = s_maxid;
  for (int q = threadIdx.x; q < nd; q += blockDim.x) {
    y_out_track[dp + w.key[q]] = w.yo[q];
    const int row = det_rows[k0 + w.idx[q]];
    const int lr = (int)(row - base);
    // retain unassociated true positives inside the retention window (utils/graph.py:505)
    const bool retained = g.ass[row] == -1 && (w.flag[q] & 1) && w.ts[q] >= t_upto - ret_win;
    keep[row] = (lr >= max_id || retained) ? 1 : 0;
  }
}

// keep flags of the edge rows: gone if before max_id or hanging off a deleted detection
__global__ void __launch_bounds__(256) k_keep_edges(tmpnn_graph g, const int32_t* __restrict__ max_id_arr,
                                                    uint8_t* __restrict__ keep) {
  const int s = blockIdx.y, n = g.n_rows[s];
  const int max_id = max_id_arr[s];
  const size_t base = (size_t)s * g.cap_rows;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    if (max_id < 0) { keep[base + r] = 1; continue; }
    const int a = g.src[base + r];
    if (a < 0) continue;  // detection rows were written by k_decode
    keep[base + r] = (r >= max_id && keep[base + a]) ? 1 : 0;
  }
}

__global__ void __launch_bounds__(256) k_prune_bounds(tmpnn_graph g, const int32_t* __restrict__ seq_det_ptr,
                                                      const int32_t* __restrict__ det_rows, int t_st, int t_ed,
                                                      int32_t* __restrict__ bounds) {
  __shared__ int lo, hi;
  const int s = blockIdx.x;
  if (threadIdx.x == 0) { lo = INT_MAX; hi = -1; }
  __syncthreads();
  const size_t base = (size_t)s * g.cap_rows;
  for (int k = seq_det_ptr[s] + threadIdx.x; k < seq_det_ptr[s + 1]; k += blockDim.x) {
    const int row = det_rows[k], t = g.ts[row];
    if (t >= t_st && t <= t_ed) { atomicMin(&lo, (int)(row - base)); atomicMax(&hi, (int)(row - base)); }
  }
  __syncthreads();
  if (threadIdx.x == 0) { bounds[2 * s] = lo; bounds[2 * s + 1] = hi; }
}
__global__ void __launch_bounds__(256) k_prune_mask(tmpnn_graph g, const int32_t* __restrict__ bounds, float thr,
                                                    uint8_t* __restrict__ keep) {
  const int s = blockIdx.y, n = g.n_rows[s];
  const int lo = bounds[2 * s], hi = bounds[2 * s + 1];
  const size_t base = (size_t)s * g.cap_rows;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    keep[base + r] = (hi < 0 || g.score[base + r] >= thr || g.ts[base + r] != -1 || r < lo || r > hi) ? 1 : 0;
}

// ---- compaction --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_compact_count(const int32_t* __restrict__ n_rows, int cap_rows, int nblk,
                                                       const uint8_t* __restrict__ keep, int32_t* __restrict__ blk_cnt) {
  __shared__ int sm[33];
  const int s = blockIdx.y, b = blockIdx.x, n = n_rows[s];
  const int r0 = b * ROWS_PER_BLOCK + threadIdx.x * 4;
  int c = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (r0 + q < n && keep[(size_t)s * cap_rows + r0 + q]) ++c;
  int total;
  block_exclusive_scan(c, sm, &total);
  if (threadIdx.x == 0) blk_cnt[s * nblk + b] = total;
}
// one CTA per sequence: exclusive scan of its block counts, new row count
__global__ void __launch_bounds__(128) k_compact_scan(int nblk, int32_t* __restrict__ blk_cnt, int32_t* __restrict__ n_rows_out) {
  __shared__ int sm[33];
  const int s = blockIdx.x;
  int carry = 0;
  for (int b0 = 0; b0 < nblk; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    const int v = b < nblk ? blk_cnt[s * nblk + b] : 0;
    int total;
    const int ex = block_exclusive_scan(v, sm, &total);
    if (b < nblk) blk_cnt[s * nblk + b] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) n_rows_out[s] = carry;
}
__global__ void __launch_bounds__(256) k_compact_map(const int32_t* __restrict__ n_rows, int cap_rows, int nblk,
                                                     const uint8_t* __restrict__ keep, const int32_t* __restrict__ blk_off,
                                                     int32_t* __restrict__ new_of_old) {
  __shared__ int sm[33];
  const int s = blockIdx.y, b = blockIdx.x, n = n_rows[s];
  if (b * ROWS_PER_BLOCK >= n) return;
  const int r0 = b * ROWS_PER_BLOCK + threadIdx.x * 4;
  int f[4], c = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[q] = (r0 + q < n && keep[(size_t)s * cap_rows + r0 + q]) ? 1 : 0;
    c += f[q];
  }
  int total;
  int pos = blk_off[s * nblk + b] + block_exclusive_scan(c, sm, &total);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (r0 + q < n) new_of_old[(size_t)s * cap_rows + r0 + q] = f[q] ? pos++ : -1;
}
// moves the surviving rows: metadata (one thread per row) and h (16 lanes x float4 per 256 B)
__global__ void __launch_bounds__(256)
k_compact_move(tmpnn_graph gi, tmpnn_graph go, const int32_t* __restrict__ new_of_old, const float* __restrict__ h_src,
               const float* __restrict__ h_src_inactive, const int32_t* __restrict__ active, float* __restrict__ h_dst,
               int ldh) {
  const int s = blockIdx.y, n = gi.n_rows[s];
  const size_t base = (size_t)s * gi.cap_rows;
  const bool off = seq_off(active, s);
  const float* hs = (off && h_src_inactive) ? h_src_inactive : h_src;
  const int r_begin = blockIdx.x * ROWS_PER_BLOCK, r_end = min(n, r_begin + ROWS_PER_BLOCK);
  // deferred compaction: the step left the state of an active sequence dense at the OLD logical rows of
  // h_src; nothing is moved, the maps say where the survivors are.  A sequence that did not step this frame
  // still lives in the previous buffer (through its own maps) and is copied over densely.
  const bool deferred = go.phys != nullptr;
  if (deferred && blockIdx.x == 0 && threadIdx.x == 0) go.phys_end[s] = off ? go.n_rows[s] : n;
  for (int r = r_begin + threadIdx.x; r < r_end; r += blockDim.x) {
    const int nr = new_of_old[base + r];
    if (nr < 0) continue;
    const size_t o = base + nr, i = base + r;
    if (deferred) {
      const int a0 = gi.src[i];
      if (off) {
        go.phys[o] = (int32_t)o;
        go.psrc[o] = a0 < 0 ? -1 : (int32_t)(base + new_of_old[base + a0]);
        go.pdst[o] = a0 < 0 ? -1 : (int32_t)(base + new_of_old[base + gi.dst[i]]);
      } else {
        go.phys[o] = (int32_t)i;
        go.psrc[o] = a0 < 0 ? -1 : (int32_t)(base + a0);
        go.pdst[o] = a0 < 0 ? -1 : (int32_t)(base + gi.dst[i]);
      }
    }
    go.ts[o] = gi.ts[i];
    go.det[o] = gi.det[i];
    go.ass[o] = gi.ass[i];
    const int a = gi.src[i];
    go.src[o] = a < 0 ? -1 : new_of_old[base + a];
    go.dst[o] = a < 0 ? -1 : new_of_old[base + gi.dst[i]];
    if (gi.label && go.label) go.label[o] = gi.label[i];
    go.score[o] = gi.score[i];
    go.logit[o] = gi.logit[i];
  }
  if (h_dst && (!deferred || off)) {
    const int v4 = ldh / 4;
    const long long items = (long long)(r_end - r_begin) * v4;
    for (long long it = threadIdx.x; it < items; it += blockDim.x) {
      const int r = r_begin + (int)(it / v4), c = (int)(it % v4);
      const int nr = new_of_old[base + r];
      if (nr < 0) continue;
      const size_t from = deferred ? (size_t)gi.phys[base + r] : base + r;
      reinterpret_cast<float4*>(h_dst + (base + nr) * ldh)[c] = __ldg(reinterpret_cast<const float4*>(hs + from * ldh) + c);
    }
  }
}

__global__ void __launch_bounds__(256) k_phys_identity(tmpnn_graph g) {
  const int s = blockIdx.y, n = g.n_rows[s];
  const size_t base = (size_t)s * g.cap_rows;
  if (blockIdx.x == 0 && threadIdx.x == 0) g.phys_end[s] = n;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int a = g.src[base + r];
    g.phys[base + r] = (int32_t)(base + r);
    g.psrc[base + r] = a < 0 ? -1 : (int32_t)(base + a);
    g.pdst[base + r] = a < 0 ? -1 : (int32_t)(base + g.dst[base + r]);
  }
}

inline dim3 row_grid(const tmpnn_graph* g) { return dim3(tmpnn_div_up(g->cap_rows, ROWS_PER_BLOCK), g->num_seqs); }
inline dim3 stride_grid(const tmpnn_graph* g) {
  return dim3(max(1, min(tmpnn_div_up(g->cap_rows, 256 * 8), 64)), g->num_seqs);
}

}  // namespace

static bool g_graph_init_done = false;
int tmpnn_init_graph_ops() {
  TMPNN_CUDA_TRY(cudaFuncSetAttribute(k_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 4 * WALK_SMEM_MAX));
  g_graph_init_done = true;
  return TMPNN_OK;
}

extern "C" int tmpnn_ypred_unpack(const int64_t* y_pred, int n, int32_t* ts, int32_t* det, int32_t* ass, void* stream) {
  if (n <= 0) return TMPNN_OK;
  k_ypred_unpack<<<min(tmpnn_div_up(n, 256), 1184), 256, 0, (cudaStream_t)stream>>>(y_pred, n, ts, det, ass);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
extern "C" int tmpnn_ypred_pack(const int32_t* ts, const int32_t* det, const int32_t* ass, int n, int64_t* y_pred, void* stream) {
  if (n <= 0) return TMPNN_OK;
  k_ypred_pack<<<min(tmpnn_div_up(n, 256), 1184), 256, 0, (cudaStream_t)stream>>>(ts, det, ass, n, y_pred);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_coo_from_edges(const int32_t* ts, const int32_t* src, const int32_t* dst, int n, int transpose,
                                    int64_t* idx, float* val, int64_t nnz, int32_t* scratch, void* stream) {
  if (n <= 0) return TMPNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* cnt = scratch;            // [n]
  int32_t* off = scratch + n;        // [n+1]
  int32_t* sums = off + n + 1;       // [div_up(n+1, 2048)]
  const int blocks = min(tmpnn_div_up(n, 256), 1184);
  k_coo_count<<<blocks, 256, 0, st>>>(ts, n, transpose, cnt);
  TMPNN_LAUNCH_CHECK();
  TMPNN_CUDA_TRY(scan_exclusive(cnt, off, nullptr, 0, n, n, sums, st));
  k_coo_write<<<blocks, 256, 0, st>>>(ts, src, dst, n, transpose, off, idx, val, nnz);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_edges_from_coo(const int64_t* idx, const float* val, int64_t nnz, int n, int32_t* src, int32_t* dst,
                                    void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return TMPNN_OK;
  TMPNN_CUDA_TRY(cudaMemsetAsync(src, 0xff, sizeof(int32_t) * (size_t)n, st));
  TMPNN_CUDA_TRY(cudaMemsetAsync(dst, 0xff, sizeof(int32_t) * (size_t)n, st));
  if (nnz <= 0) return TMPNN_OK;
  k_edges_from_coo<<<(int)((nnz + 255) / 256 < 1184 ? (nnz + 255) / 256 : 1184), 256, 0, st>>>(idx, val, nnz, n, src, dst);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// The engine's work counters and clock in one launch (they were five tiny PyTorch kernels per frame): every pair is optional.
__global__ void __launch_bounds__(256) k_counters(int64_t* __restrict__ edge_updates, const int32_t* __restrict__ n_edges,
                                                  int64_t* __restrict__ det_updates, const int32_t* __restrict__ n_dets,
                                                  int64_t* __restrict__ frames_done, const int32_t* __restrict__ active,
                                                  int num_seqs, int32_t* __restrict__ t_dev) {
  __shared__ int sm[8];
  int c = 0;
  if (frames_done && active)
    for (int s = threadIdx.x; s < num_seqs; s += blockDim.x) c += active[s] != 0 ? active[s] : 0;
  c = warp_sum_i(c);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < 8; ++w) tot += sm[w];
    if (frames_done && active) *frames_done += tot;
    if (edge_updates && n_edges) *edge_updates += *n_edges;
    if (det_updates && n_dets) *det_updates += *n_dets;
    if (t_dev) *t_dev += 1;
  }
}
extern "C" int tmpnn_graph_counters(int64_t* edge_updates, const int32_t* n_edges, int64_t* det_updates, const int32_t* n_dets,
                                    int64_t* frames_done, const int32_t* active, int num_seqs, int32_t* t_dev, void* stream) {
  k_counters<<<1, 256, 0, (cudaStream_t)stream>>>(edge_updates, n_edges, det_updates, n_dets, frames_done, active, num_seqs, t_dev);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_graph_associate(const tmpnn_graph* g, const tmpnn_index* ix, int mode, const int32_t* active,
                                     void* stream) {
  TMPNN_REQUIRE(g && ix, "null argument");
  TMPNN_REQUIRE(mode == 0 || (mode == 1 && g->label), "teacher forcing needs labels");
  cudaStream_t st = (cudaStream_t)stream;
  k_reset_ass<<<stride_grid(g), 256, 0, st>>>(g->n_rows, active, g->cap_rows, g->ass);
  TMPNN_LAUNCH_CHECK();
  k_associate<<<TMPNN_SM_COUNT * 8, 256, 0, st>>>(ix->n_dets, ix->det_rows, ix->seg_ptr, ix->inc, g->ts, g->det, g->dst,
                                                 g->label, g->score, g->cap_rows, mode, g->ass, g->status);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" size_t tmpnn_graph_append_scratch_ints(int num_seqs, int cap_rows) {
  return (size_t)num_seqs * cap_rows + DESC * (size_t)num_seqs + (size_t)num_seqs * tmpnn_div_up(cap_rows, ROWS_PER_BLOCK);
}

extern "C" int tmpnn_graph_append(const tmpnn_graph* g, const tmpnn_frames* fr, const tmpnn_seq_state* st_,
                                  const int32_t* t_dev, int start, int cur_win_size, int mode, float* h, int ldh,
                                  int32_t* new_det_rows, int32_t* new_det_x, int32_t* n_new, int cap_new,
                                  int32_t* n_appended, int32_t* scratch, void* stream) {
  TMPNN_REQUIRE(g && fr && n_new && n_appended && scratch && new_det_rows && new_det_x, "null argument");
  TMPNN_REQUIRE(start || t_dev, "t_dev is required unless start != 0 (which defaults to t = 0)");
  TMPNN_REQUIRE(!h || ldh % 4 == 0, "ldh must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* act = scratch;
  int32_t* desc = scratch + (size_t)g->num_seqs * g->cap_rows;
  tmpnn_seq_state zero = {};
  TMPNN_CUDA_TRY(cudaMemsetAsync(n_new, 0, 2 * sizeof(int32_t), st));
  const int nblk = tmpnn_div_up(g->cap_rows, ROWS_PER_BLOCK);
  int32_t* blk_cnt = desc + DESC * (size_t)g->num_seqs;
  const bool grid_scan = mode != 1 && !start;
  dim3 grid_rows(nblk, g->num_seqs);
  if (grid_scan) {
    k_active_count<<<grid_rows, 256, 0, st>>>(*g, nblk, blk_cnt);
    TMPNN_LAUNCH_CHECK();
  }
  k_append_plan<<<g->num_seqs, 1024, 0, st>>>(*g, *fr, st_ ? *st_ : zero, st_ ? 1 : 0, t_dev, start, cur_win_size, mode,
                                              act, desc, n_new, cap_new, n_appended, grid_scan ? blk_cnt : nullptr, nblk);
  TMPNN_LAUNCH_CHECK();
  if (grid_scan) {
    k_active_write<<<grid_rows, 256, 0, st>>>(*g, nblk, blk_cnt, desc, DESC, act);
    TMPNN_LAUNCH_CHECK();
  }
  dim3 grid(max(1, min(tmpnn_div_up(g->cap_rows, 256 * 4), 128)), g->num_seqs);
  k_append_fill<<<grid, 256, 0, st>>>(*g, *fr, act, desc, h, ldh, new_det_rows, new_det_x);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

// shared-memory capacity (detection rows, a power of two) the walk is launched with for windows of up to max_dets
// detection rows (max_dets <= 0: unknown)
static int walk_smem_cap(int max_dets) {
  if (max_dets <= 0 || max_dets > WALK_SMEM_MAX) return WALK_SMEM_DEFAULT;
  return max(256, pow2_ceil(max_dets));
}
extern "C" size_t tmpnn_graph_decode_scratch_ints(int num_seqs, int max_dets) {
  size_t n = (size_t)max(num_seqs, 1) + 4;
  if (max_dets > WALK_SMEM_MAX) n += (size_t)6 * pow2_ceil(max_dets) * max(num_seqs, 1);
  return n;
}

extern "C" int tmpnn_graph_decode(const tmpnn_graph* g, const tmpnn_index* ix, const tmpnn_frames* fr,
                                  int32_t* y_out_track, int32_t* next_track_id, const int32_t* t_upto_seq,
                                  int t_upto_host, const int32_t* active, int ret_win_size, uint8_t* keep, int max_dets,
                                  int32_t* scratch, void* stream) {
  TMPNN_REQUIRE(g && ix && fr && y_out_track && next_track_id && keep && scratch, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_graph_init_done) { int rc0 = tmpnn_init_graph_ops(); if (rc0) return rc0; }
  const int smem_cap = walk_smem_cap(max_dets);
  const bool spill = max_dets > WALK_SMEM_MAX;
  int32_t* gwalk = spill ? scratch + g->num_seqs + 4 : nullptr;
  k_decode<<<g->num_seqs, 256, (size_t)6 * 4 * smem_cap, st>>>(*g, ix->seq_det_ptr, ix->det_rows, fr->det_ptr, y_out_track,
                                                             next_track_id, t_upto_seq, t_upto_host, active, ret_win_size,
                                                             keep, scratch, smem_cap, gwalk, spill ? pow2_ceil(max_dets) : 0);
  TMPNN_LAUNCH_CHECK();
  k_keep_edges<<<stride_grid(g), 256, 0, st>>>(*g, scratch, keep);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_graph_force_det_scores(const tmpnn_graph* g, const tmpnn_index* ix, void* stream) {
  TMPNN_REQUIRE(g && ix, "null argument");
  k_force_det_scores<<<max(1, min(tmpnn_div_up(ix->cap_dets, 256), 148)), 256, 0, (cudaStream_t)stream>>>(ix->n_dets, ix->det_rows,
                                                                                                        g->score);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_status_ack(const tmpnn_graph* g, int from_bits, int to_bits, void* stream) {
  TMPNN_REQUIRE(g && g->status, "null argument");
  k_status_ack<<<1, 1, 0, (cudaStream_t)stream>>>(g->status, from_bits, to_bits);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_graph_prune_mask(const tmpnn_graph* g, const tmpnn_index* ix, int t_st, int t_ed, float threshold,
                                      uint8_t* keep, int32_t* scratch, void* stream) {
  TMPNN_REQUIRE(g && ix && keep && scratch, "null argument");
  if (t_st > t_ed) return tmpnn_set_error(TMPNN_E_BADARG, "t_st must be lesser than or equal to t_ed!");
  cudaStream_t st = (cudaStream_t)stream;
  k_prune_bounds<<<g->num_seqs, 256, 0, st>>>(*g, ix->seq_det_ptr, ix->det_rows, t_st, t_ed, scratch);
  TMPNN_LAUNCH_CHECK();
  k_prune_mask<<<stride_grid(g), 256, 0, st>>>(*g, scratch, threshold, keep);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" int tmpnn_graph_phys_identity(const tmpnn_graph* g, void* stream) {
  TMPNN_REQUIRE(g && g->phys && g->psrc && g->pdst && g->phys_end, "the graph has no deferred-compaction maps");
  k_phys_identity<<<stride_grid(g), 256, 0, (cudaStream_t)stream>>>(*g);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" size_t tmpnn_graph_compact_scratch_ints(int num_seqs, int cap_rows) {
  return (size_t)num_seqs * tmpnn_div_up(cap_rows, ROWS_PER_BLOCK) + 16;
}

extern "C" int tmpnn_graph_compact(const tmpnn_graph* g_in, const tmpnn_graph* g_out, const uint8_t* keep,
                                   const float* h_src, const float* h_src_inactive, const int32_t* active, float* h_dst,
                                   int ldh, int32_t* new_of_old, int32_t* scratch, void* stream) {
  TMPNN_REQUIRE(g_in && g_out && keep && new_of_old && scratch, "null argument");
  TMPNN_REQUIRE(g_in->ts != g_out->ts && (!h_dst || h_src != h_dst || g_out->phys), "compaction is out of place");
  TMPNN_REQUIRE(!g_out->phys || (g_in->phys && h_dst && h_src_inactive && h_src_inactive != h_dst),
                "deferred compaction needs maps on both graphs and the two state buffers");
  TMPNN_REQUIRE(g_in->num_seqs == g_out->num_seqs && g_in->cap_rows == g_out->cap_rows, "slab shapes differ");
  TMPNN_REQUIRE(!h_dst || ldh % 4 == 0, "ldh must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = tmpnn_div_up(g_in->cap_rows, ROWS_PER_BLOCK);
  k_compact_count<<<row_grid(g_in), 256, 0, st>>>(g_in->n_rows, g_in->cap_rows, nblk, keep, scratch);
  TMPNN_LAUNCH_CHECK();
  k_compact_scan<<<g_in->num_seqs, 128, 0, st>>>(nblk, scratch, g_out->n_rows);
  TMPNN_LAUNCH_CHECK();
  k_compact_map<<<row_grid(g_in), 256, 0, st>>>(g_in->n_rows, g_in->cap_rows, nblk, keep, scratch, new_of_old);
  TMPNN_LAUNCH_CHECK();
  k_compact_move<<<row_grid(g_in), 256, 0, st>>>(*g_in, *g_out, new_of_old, h_src, h_src_inactive, active, h_dst, ldh);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
