// graph_index.cu -- per-step index over the slabs: detection-row list, per-detection
// incidence CSR (past edges, then future edges, each ascending) and the tile table.
//
// Replaces the reference's dense edge_adj = node_adj^T (utils/graph.py:158, 300) and every
// np.where(node_adj[:, i]) column scan (utils/graph.py:56, 232, 256, 508; models/loss.py:20, 34).
// Pure integer work, HBM bound: ~4 B/row read for the detection list, ~40 B/edge row for the
// CSR (src/dst read twice, two counters, one incidence written and sorted).
#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int ROWS_PER_BLOCK = 1024;  // 256 threads x 4 rows
constexpr int SORT_CAP = 4096;        // longest incidence segment one CTA sorts in shared memory

// ---- detection list ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_dets(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active,
                                                    const int32_t* __restrict__ ts, int cap_rows, int nblk,
                                                    int32_t* __restrict__ blk_cnt) {
  __shared__ int sm[33];
  const int s = blockIdx.y, b = blockIdx.x;
  const int n = (active && !active[s]) ? 0 : n_rows[s];
  int c = 0;
  const int r0 = b * ROWS_PER_BLOCK + threadIdx.x * 4;
  if (r0 < n) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (r0 + q < n && ts[(size_t)s * cap_rows + r0 + q] >= 0) ++c;
  }
  int total;
  block_exclusive_scan(c, sm, &total);
  if (threadIdx.x == 0) blk_cnt[s * nblk + b] = total;
}

// single CTA: scan of the block counts, per-sequence ranges, tile table, totals
__global__ void __launch_bounds__(1024) k_scan_det_blocks(const int32_t* __restrict__ n_rows,
                                                          const int32_t* __restrict__ active, int num_seqs, int nblk,
                                                          int32_t* __restrict__ blk_cnt, int32_t* __restrict__ seq_det_ptr,
                                                          int32_t* __restrict__ tile_ptr, int32_t* __restrict__ tile128_ptr,
                                                          int32_t* __restrict__ n_dets,
                                                          int32_t* __restrict__ n_edges, int cap_dets, int cap_inc,
                                                          int32_t* __restrict__ status) {
  __shared__ int sm[33];
  const int nb = num_seqs * nblk;
  int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = i < nb ? blk_cnt[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, sm, &total);
    if (i < nb) {
      blk_cnt[i] = carry + ex;
      if (i % nblk == 0) seq_det_ptr[i / nblk] = carry + ex;
    }
    carry += total;
  }
  const int nd = carry;
  int tcarry = 0, rcarry = 0, t128carry = 0;
  for (int s0 = 0; s0 < num_seqs; s0 += 1024) {
    const int s = s0 + threadIdx.x;
    const int n = (s < num_seqs && !(active && !active[s])) ? n_rows[s] : 0;
    int total;
    const int ex = block_exclusive_scan((n + TMPNN_TILE_ROWS - 1) / TMPNN_TILE_ROWS, sm, &total);
    if (s < num_seqs) tile_ptr[s] = tcarry + ex;
    tcarry += total;
    if (tile128_ptr) {
      int t128;
      const int ex128 = block_exclusive_scan((n + 127) / 128, sm, &t128);
      if (s < num_seqs) tile128_ptr[s] = t128carry + ex128;
      t128carry += t128;
    }
    int rt;
    block_exclusive_scan(n, sm, &rt);
    rcarry += rt;
  }
  if (threadIdx.x == 0) {
    seq_det_ptr[num_seqs] = nd;
    tile_ptr[num_seqs] = tcarry;
    if (tile128_ptr) tile128_ptr[num_seqs] = t128carry;
    const int ne = rcarry - nd;
    int flags = 0;
    if (nd > cap_dets) flags |= TMPNN_FLAG_DET_CAPACITY;
    if (2 * ne > cap_inc) flags |= TMPNN_FLAG_INC_CAPACITY;
    if (flags) atomicOr(status, flags);
    // on overflow the consumers see an empty index instead of writing out of bounds
    *n_dets = flags ? 0 : nd;
    *n_edges = ne;
  }
}

__global__ void __launch_bounds__(256) k_write_dets(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active,
                                                    const int32_t* __restrict__ ts, int cap_rows, int nblk,
                                                    const int32_t* __restrict__ blk_off,
                                                    const int32_t* __restrict__ n_dets, int32_t* __restrict__ det_rows,
                                                    int32_t* __restrict__ det_of_row, int32_t* __restrict__ cnt) {
  __shared__ int sm[33];
  const int s = blockIdx.y, b = blockIdx.x;
  const int n = (active && !active[s]) ? 0 : n_rows[s];
  const int r0 = b * ROWS_PER_BLOCK + threadIdx.x * 4;
  if (b * ROWS_PER_BLOCK >= n) return;  // whole block past the end (uniform)
  int f[4];
  int c = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[q] = (r0 + q < n && ts[(size_t)s * cap_rows + r0 + q] >= 0) ? 1 : 0;
    c += f[q];
  }
  int total;
  int k = block_exclusive_scan(c, sm, &total) + blk_off[s * nblk + b];
  const bool overflow = (*n_dets == 0);  // set when capacity was exceeded (or the graph has no detections)
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (r0 + q < n) {
      const size_t row = (size_t)s * cap_rows + r0 + q;
      if (f[q] && !overflow) {
        det_rows[k] = (int32_t)row;
        det_of_row[row] = k;
        cnt[2 * k] = 0;
        cnt[2 * k + 1] = 0;
        ++k;
      } else {
        det_of_row[row] = -1;
      }
    }
  }
}

// ---- incidence CSR -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_degree(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active,
                                                const int32_t* __restrict__ src,
                                                const int32_t* __restrict__ dst, int cap_rows,
                                                const int32_t* __restrict__ det_of_row, const int32_t* __restrict__ n_dets,
                                                int32_t* __restrict__ cnt) {
  const int s = blockIdx.y;
  const int n = (active && !active[s]) ? 0 : n_rows[s];
  if (*n_dets == 0) return;
  const size_t base = (size_t)s * cap_rows;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int a = src[base + r];
    if (a < 0) continue;
    const int b = dst[base + r];
    atomicAdd(&cnt[2 * det_of_row[base + a] + 1], 1);  // future edge of src
    atomicAdd(&cnt[2 * det_of_row[base + b]], 1);      // past edge of dst
  }
}

__global__ void __launch_bounds__(256) k_fill(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active,
                                                const int32_t* __restrict__ src,
                                              const int32_t* __restrict__ dst, int cap_rows,
                                              const int32_t* __restrict__ det_of_row, const int32_t* __restrict__ n_dets,
                                              int32_t* __restrict__ cnt, const int32_t* __restrict__ seg_ptr,
                                              int32_t* __restrict__ inc) {
  const int s = blockIdx.y;
  const int n = (active && !active[s]) ? 0 : n_rows[s];
  if (*n_dets == 0) return;
  const size_t base = (size_t)s * cap_rows;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    const int a = src[base + r];
    if (a < 0) continue;
    const int b = dst[base + r];
    const int ka = 2 * det_of_row[base + a] + 1, kb = 2 * det_of_row[base + b];
    inc[seg_ptr[ka] + atomicSub(&cnt[ka], 1) - 1] = (int32_t)(base + r);
    inc[seg_ptr[kb] + atomicSub(&cnt[kb], 1) - 1] = (int32_t)(base + r);
  }
}

// Ascending order inside every segment (the atomic fill order is arbitrary): one CTA per
// segment, bitonic network in shared memory.  Gives bit-reproducible aggregation sums and the
// "first / last positive edge" semantics of models/loss.py:26-43.
__global__ void __launch_bounds__(128) k_sort_segments(const int32_t* __restrict__ n_dets,
                                                       const int32_t* __restrict__ seg_ptr, int32_t* __restrict__ inc,
                                                       int32_t* __restrict__ status) {
  __shared__ int32_t key[SORT_CAP];
  const int nseg = 2 * (*n_dets);
  for (int x = blockIdx.x; x < nseg; x += gridDim.x) {
    const int s0 = seg_ptr[x], len = seg_ptr[x + 1] - s0;
    if (len <= 1) continue;
    if (len > SORT_CAP) {
      if (threadIdx.x == 0) atomicOr(status, TMPNN_FLAG_SEG_CAPACITY);
      continue;
    }
    int p = 2;
    while (p < len) p <<= 1;
    __syncthreads();
    int unsorted = 0;
    for (int i = threadIdx.x; i < p; i += blockDim.x) {
      const int v = i < len ? inc[s0 + i] : 0x7fffffff;
      key[i] = v;
      if (i + 1 < len && inc[s0 + i + 1] < v) unsorted = 1;
    }
    if (!__syncthreads_or(unsorted)) continue;
    for (int k = 2; k <= p; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) {
          const int l = i ^ j;
          if (l > i) {
            const int a = key[i], b = key[l];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { key[i] = b; key[l] = a; }
          }
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < len; i += blockDim.x) inc[s0 + i] = key[i];
  }
}

// ---- structured build: incidence CSR from the block structure, no atomics, no sort ----------
// A window graph that only ever grew by tmpnn_graph_append and shrank by tmpnn_graph_decode's deletion
// (whole prefixes, plus whole source runs of later edge blocks) is a chain of segments
//   [dets f0] [edges -> f1] [dets f1] [edges -> f2] [dets f2] ...
// where the edge segment in front of a detection segment of Nt rows is a dense [A x Nt] matrix:
// row e0 + a Nt + j joins source a (constant along the run) to detection d0 + j (utils/graph.py:283-301;
// Appendix A.4/A.5 of SURVEY.md).  Every incidence list is then a formula: the past edges of detection
// d0 + j are e0 + a Nt + j (a ascending), the future edges of a source are its runs in ascending
// segment order.  The builder finds the segment boundaries from ts, derives the counts, scans them and
// writes inc with coalesced stores; it verifies the structure as it goes (TMPNN_FLAG_UNSTRUCTURED).
// MAXSEG, MAXE and SlabSegs (the per-slab segment table this builder leaves in scratch2) live in common.cuh: the
// block-structured aggregation (mp_step.cu) reads the same table

// Segment boundaries: a boundary sits in front of row i when i is the first row or the timestamp changes (edge rows all
// carry -1).  Found by a grid over the rows (a window of W = 20 frames x 200 detections has millions of rows per
// slab; one CTA per slab took 22 % of a frame there), collected per slab, then sorted by one thread (<= 128 of them).
constexpr int BND_STRIDE = MAXSEG + 4;  // [0] = count, [1 ..] = positions
__global__ void __launch_bounds__(256) k_segment_bounds(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active,
                                                        const int32_t* __restrict__ ts, int cap_rows, int32_t* __restrict__ bnd) {
  const int s = blockIdx.y;
  const int n = (active && !active[s]) ? 0 : n_rows[s];
  const int32_t* t = ts + (size_t)s * cap_rows;
  int32_t* o = bnd + (size_t)s * BND_STRIDE;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i == 0 || t[i] != t[i - 1]) {
      const int k = atomicAdd(&o[0], 1);
      if (k < MAXSEG + 1) o[1 + k] = i;
    }
  }
}
__global__ void __launch_bounds__(32) k_block_segments(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active,
                                                       const int32_t* __restrict__ ts, int cap_rows, int32_t* __restrict__ bnd_all,
                                                       SlabSegs* __restrict__ segs, int32_t* __restrict__ slab_nd,
                                                       int32_t* __restrict__ status) {
  const int s = blockIdx.x;
  if (threadIdx.x != 0) return;
  const int n = (active && !active[s]) ? 0 : n_rows[s];
  const int32_t* t = ts + (size_t)s * cap_rows;
  int32_t* bnd = bnd_all + (size_t)s * BND_STRIDE + 1;
  SlabSegs& o = segs[s];
  int m = bnd[-1];
  bnd[-1] = 0;  // ready for the next build
  if (m > MAXSEG) { atomicOr(status, TMPNN_FLAG_UNSTRUCTURED); m = 0; }
  for (int a = 1; a < m; ++a) {  // insertion sort of <= 128 positions
    const int v = bnd[a];
    int b = a - 1;
    while (b >= 0 && bnd[b] > v) { bnd[b + 1] = bnd[b]; --b; }
    bnd[b + 1] = v;
  }
  int ne = 0;
  for (int q = 0; q < m; ++q) {
    o.start[q] = bnd[q];
    const bool edge = t[bnd[q]] < 0;
    o.eord[q] = edge ? ne : -1;
    if (edge) ++ne;
  }
  if (ne > MAXE) { atomicOr(status, TMPNN_FLAG_UNSTRUCTURED); m = 0; ne = 0; }
  o.nedge = ne;
  o.nseg = m;
  o.start[m] = n;
  // detection rows of the slab = total length of its detection segments (the structured builder derives the detection
  // list from the segments instead of scanning every row's timestamp twice)
  int nd = 0;
  for (int q = 0; q < m; ++q)
    if (o.eord[q] < 0) nd += o.start[q + 1] - o.start[q];
  slab_nd[s] = nd;
}

// single CTA: per-slab detection ranges, tile tables, totals (the structured counterpart of k_scan_det_blocks)
__global__ void __launch_bounds__(1024) k_scan_slabs(const int32_t* __restrict__ n_rows, const int32_t* __restrict__ active,
                                                     int num_seqs, const int32_t* __restrict__ slab_nd,
                                                     int32_t* __restrict__ seq_det_ptr, int32_t* __restrict__ tile_ptr,
                                                     int32_t* __restrict__ tile128_ptr, int32_t* __restrict__ n_dets,
                                                     int32_t* __restrict__ n_edges, int cap_dets, int cap_inc,
                                                     int32_t* __restrict__ status) {
  __shared__ int sm[33];
  int dcarry = 0, tcarry = 0, rcarry = 0, t128carry = 0;
  for (int s0 = 0; s0 < num_seqs; s0 += 1024) {
    const int s = s0 + threadIdx.x;
    const int n = (s < num_seqs && !(active && !active[s])) ? n_rows[s] : 0;
    int total;
    int ex = block_exclusive_scan(s < num_seqs ? slab_nd[s] : 0, sm, &total);
    if (s < num_seqs) seq_det_ptr[s] = dcarry + ex;
    dcarry += total;
    ex = block_exclusive_scan((n + TMPNN_TILE_ROWS - 1) / TMPNN_TILE_ROWS, sm, &total);
    if (s < num_seqs) tile_ptr[s] = tcarry + ex;
    tcarry += total;
    if (tile128_ptr) {
      ex = block_exclusive_scan((n + 127) / 128, sm, &total);
      if (s < num_seqs) tile128_ptr[s] = t128carry + ex;
      t128carry += total;
    }
    block_exclusive_scan(n, sm, &total);
    rcarry += total;
  }
  if (threadIdx.x == 0) {
    const int nd = dcarry;
    seq_det_ptr[num_seqs] = nd;
    tile_ptr[num_seqs] = tcarry;
    if (tile128_ptr) tile128_ptr[num_seqs] = t128carry;
    const int ne = rcarry - nd;
    int flags = 0;
    if (nd > cap_dets) flags |= TMPNN_FLAG_DET_CAPACITY;
    if (2 * ne > cap_inc) flags |= TMPNN_FLAG_INC_CAPACITY;
    if (flags) atomicOr(status, flags);
    *n_dets = flags ? 0 : nd;   // on overflow the consumers see an empty index instead of writing out of bounds
    *n_edges = ne;
  }
}

// detection list from the detection segments: only the detection rows are touched (det_of_row of association rows is
// never read: every consumer indexes it with an endpoint row)
__global__ void __launch_bounds__(256) k_write_dets_segs(const SlabSegs* __restrict__ segs, int cap_rows,
                                                         const int32_t* __restrict__ seq_det_ptr,
                                                         const int32_t* __restrict__ n_dets, int32_t* __restrict__ det_rows,
                                                         int32_t* __restrict__ det_of_row, int32_t* __restrict__ cnt) {
  if (*n_dets == 0) return;  // capacity exceeded (or no detections at all)
  const int s = blockIdx.y;
  const SlabSegs& sg = segs[s];
  const size_t base = (size_t)s * cap_rows;
  int k0 = seq_det_ptr[s];
  for (int q = 0; q < sg.nseg; ++q) {
    if (sg.eord[q] >= 0) continue;
    const int r0 = sg.start[q], len = sg.start[q + 1] - r0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x) {
      const int k = k0 + i;
      det_rows[k] = (int32_t)(base + r0 + i);
      det_of_row[base + r0 + i] = k;
      cnt[2 * k] = 0;
      cnt[2 * k + 1] = 0;
    }
    k0 += len;
  }
}

// counts: cnt[2k] = past edges, futlen[k][eord] = run length of detection k in edge segment eord
__global__ void __launch_bounds__(256) k_block_degree(const SlabSegs* __restrict__ segs, const int32_t* __restrict__ src,
                                                      int cap_rows, const int32_t* __restrict__ det_of_row,
                                                      const int32_t* __restrict__ n_dets, int32_t* __restrict__ cnt,
                                                      int32_t* __restrict__ futlen, int32_t* __restrict__ status) {
  if (*n_dets == 0) return;
  const int s = blockIdx.y;
  const SlabSegs& sg = segs[s];
  const size_t base = (size_t)s * cap_rows;
  for (int q = 0; q < sg.nseg; ++q) {
    if (sg.eord[q] < 0) continue;
    const int e0 = sg.start[q], d0 = sg.start[q + 1], d1 = q + 2 <= sg.nseg ? sg.start[q + 2] : d0;
    const int nt = d1 - d0;
    if (nt <= 0 || q + 1 >= sg.nseg || sg.eord[q + 1] >= 0 || (d0 - e0) % nt != 0) {
      if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(status, TMPNN_FLAG_UNSTRUCTURED);
      continue;
    }
    const int A = (d0 - e0) / nt;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < nt; j += gridDim.x * blockDim.x)
      cnt[2 * det_of_row[base + d0 + j]] = A;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < A; a += gridDim.x * blockDim.x) {
      const int d = src[base + e0 + (size_t)a * nt];
      futlen[(size_t)det_of_row[base + d] * MAXE + sg.eord[q]] = nt;
    }
  }
}

// per detection: run lengths -> offsets inside its future list, total -> cnt[2k+1]
__global__ void __launch_bounds__(256) k_block_futoff(const int32_t* __restrict__ n_dets, int32_t* __restrict__ futlen,
                                                      int32_t* __restrict__ cnt, const SlabSegs* __restrict__ segs,
                                                      const int32_t* __restrict__ det_rows, int cap_rows) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= *n_dets) return;
  int32_t* f = futlen + (size_t)k * MAXE;
  // only the slab's own edge segments can hold a run (2 W - 1 at most, usually W): the other MAXE slots are never touched
  const int ne = segs[det_rows[k] / cap_rows].nedge;
  int acc = 0;
  for (int b = 0; b < ne; ++b) {
    const int v = f[b];
    f[b] = acc;
    acc += v;
  }
  cnt[2 * k + 1] = acc;
}

// Incidence lists by formula, one warp per list piece: a source's run in an edge block (future list: nt consecutive rows,
// checked against src / dst on the way) or a detection's column of the block in front of it (past list: A rows nt apart, no
// loads at all).  List offsets are fetched once per piece, every store is coalesced.  Measured per frame on configs[2]
// (16 M rows): row-parallel form with two integer divisions and eight dependent loads per row 144 us, this form 136 us, a
// warp walking 32 pieces after fetching their headers at once 201 us (too few warps in flight for the checking loads).
__global__ void __launch_bounds__(256) k_block_fill(const SlabSegs* __restrict__ segs, const int32_t* __restrict__ src,
                                                    const int32_t* __restrict__ dst, int cap_rows,
                                                    const int32_t* __restrict__ det_of_row, const int32_t* __restrict__ n_dets,
                                                    const int32_t* __restrict__ futoff, const int32_t* __restrict__ seg_ptr,
                                                    int32_t* __restrict__ inc, int32_t* __restrict__ status) {
  if (*n_dets == 0) return;
  const int s = blockIdx.y;
  const SlabSegs& sg = segs[s];
  const size_t base = (size_t)s * cap_rows;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nwarps = gridDim.x * (blockDim.x >> 5);
  int bad = 0;
  int item0 = 0;   // pieces of the blocks in front of this one
  for (int q = 0; q + 1 < sg.nseg; ++q) {
    if (sg.eord[q] < 0 || sg.eord[q + 1] >= 0) continue;
    const int e0 = sg.start[q], d0 = sg.start[q + 1], nt = sg.start[q + 2] - d0;
    if (nt <= 0 || (d0 - e0) % nt != 0) continue;
    const int A = (d0 - e0) / nt, eo = sg.eord[q];
    // pieces [item0, item0 + A): runs, [item0 + A, item0 + A + nt): columns; dealt round-robin to the slab's warps
    int first = warp - item0 % nwarps;
    if (first < 0) first += nwarps;
    for (int it = first; it < A + nt; it += nwarps) {
      if (it < A) {
        const int a = it;
        const size_t r0 = base + e0 + (size_t)a * nt;
        const int sd = src[r0];
        const int ks = det_of_row[base + sd];
        int32_t* out = inc + seg_ptr[2 * ks + 1] + futoff[(size_t)ks * MAXE + eo];
        for (int j = lane; j < nt; j += 32) {
          if (src[r0 + j] != sd || dst[r0 + j] != d0 + j) bad = 1;
          out[j] = (int32_t)(r0 + j);
        }
      } else {
        const int j = it - A;
        int32_t* out = inc + seg_ptr[2 * det_of_row[base + d0 + j]];
        const int32_t r0 = (int32_t)(base + e0 + j);
        for (int a = lane; a < A; a += 32) out[a] = r0 + a * nt;
      }
    }
    item0 += A + nt;
  }
  if (bad) atomicOr(status, TMPNN_FLAG_UNSTRUCTURED);
}

}  // namespace

extern "C" size_t tmpnn_index_scratch_ints(int num_seqs, int cap_rows, int cap_dets) {
  const size_t nblk = (size_t)tmpnn_div_up(cap_rows, ROWS_PER_BLOCK);
  const size_t nchunks = (size_t)tmpnn_div_up(2 * (size_t)cap_dets + 2, SCAN_CHUNK);
  return (size_t)num_seqs * nblk + (2 * (size_t)cap_dets + 4) + nchunks + 64;
}

extern "C" int tmpnn_index_build(const tmpnn_graph* g, const tmpnn_index* ix, const int32_t* active, void* stream) {
  TMPNN_REQUIRE(g && ix && ix->scratch, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = g->num_seqs;
  const int nblk = tmpnn_div_up(g->cap_rows, ROWS_PER_BLOCK);
  int32_t* blk = ix->scratch;
  int32_t* cnt = blk + (size_t)S * nblk;
  int32_t* sums = cnt + (2 * (size_t)ix->cap_dets + 4);

  dim3 grid_rows(nblk, S);
  k_count_dets<<<grid_rows, 256, 0, st>>>(g->n_rows, active, g->ts, g->cap_rows, nblk, blk);
  TMPNN_LAUNCH_CHECK();
  k_scan_det_blocks<<<1, 1024, 0, st>>>(g->n_rows, active, S, nblk, blk, ix->seq_det_ptr, ix->tile_ptr, ix->tile128_ptr, ix->n_dets,
                                        ix->n_edges,
                                        ix->cap_dets, ix->cap_inc, g->status);
  TMPNN_LAUNCH_CHECK();
  k_write_dets<<<grid_rows, 256, 0, st>>>(g->n_rows, active, g->ts, g->cap_rows, nblk, blk, ix->n_dets, ix->det_rows,
                                          ix->det_of_row, cnt);
  TMPNN_LAUNCH_CHECK();
  // one grid column per ~8K rows keeps the atomics kernels short without thousands of empty CTAs
  dim3 grid_e(max(1, min(tmpnn_div_up(g->cap_rows, 256 * 8), 64)), S);
  k_degree<<<grid_e, 256, 0, st>>>(g->n_rows, active, g->src, g->dst, g->cap_rows, ix->det_of_row, ix->n_dets, cnt);
  TMPNN_LAUNCH_CHECK();
  TMPNN_CUDA_TRY(scan_exclusive(cnt, ix->seg_ptr, ix->n_dets, 2, 0, 2 * (long long)ix->cap_dets, sums, st));
  k_fill<<<grid_e, 256, 0, st>>>(g->n_rows, active, g->src, g->dst, g->cap_rows, ix->det_of_row, ix->n_dets, cnt, ix->seg_ptr,
                                 ix->inc);
  TMPNN_LAUNCH_CHECK();
  k_sort_segments<<<TMPNN_SM_COUNT * 8, 128, 0, st>>>(ix->n_dets, ix->seg_ptr, ix->inc, g->status);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}

extern "C" size_t tmpnn_index_structured_scratch_bytes(int num_seqs, int cap_dets) {
  return (size_t)num_seqs * sizeof(SlabSegs) + (size_t)cap_dets * MAXE * sizeof(int32_t) +
         (size_t)num_seqs * BND_STRIDE * sizeof(int32_t) + 64;
}

extern "C" int tmpnn_index_build_structured(const tmpnn_graph* g, const tmpnn_index* ix, const int32_t* active,
                                            void* scratch2, void* stream) {
  TMPNN_REQUIRE(g && ix && ix->scratch && scratch2, "null argument");
  TMPNN_REQUIRE(((uintptr_t)scratch2 & 3) == 0, "scratch must be 4-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int S = g->num_seqs;
  const int nblk = tmpnn_div_up(g->cap_rows, ROWS_PER_BLOCK);
  int32_t* blk = ix->scratch;
  int32_t* cnt = blk + (size_t)S * nblk;
  int32_t* sums = cnt + (2 * (size_t)ix->cap_dets + 4);
  SlabSegs* segs = (SlabSegs*)scratch2;
  int32_t* futlen = (int32_t*)((unsigned char*)scratch2 + (size_t)S * sizeof(SlabSegs));
  int32_t* bnd = futlen + (size_t)ix->cap_dets * MAXE;  // per-slab boundary lists; the counts start at zero and are
                                                        // re-zeroed by k_block_segments (SlabIndex allocates zeros)

  // segments first; the detection list, the per-slab ranges and the tile tables follow from them (the general builder scans
  // every row's timestamp twice for the same list: ~100 us per frame on the 16 M rows of configs[2])
  dim3 grid_b(max(1, min(tmpnn_div_up(g->cap_rows, 256 * 16), 256)), S);
  k_segment_bounds<<<grid_b, 256, 0, st>>>(g->n_rows, active, g->ts, g->cap_rows, bnd);
  TMPNN_LAUNCH_CHECK();
  k_block_segments<<<S, 32, 0, st>>>(g->n_rows, active, g->ts, g->cap_rows, bnd, segs, blk, g->status);
  TMPNN_LAUNCH_CHECK();
  k_scan_slabs<<<1, 1024, 0, st>>>(g->n_rows, active, S, blk, ix->seq_det_ptr, ix->tile_ptr, ix->tile128_ptr, ix->n_dets,
                                   ix->n_edges, ix->cap_dets, ix->cap_inc, g->status);
  TMPNN_LAUNCH_CHECK();
  k_write_dets_segs<<<dim3(max(1, min(tmpnn_div_up(ix->cap_dets, 256 * max(S, 1)), 16)), S), 256, 0, st>>>(
      segs, g->cap_rows, ix->seq_det_ptr, ix->n_dets, ix->det_rows, ix->det_of_row, cnt);
  TMPNN_LAUNCH_CHECK();
  TMPNN_CUDA_TRY(cudaMemsetAsync(futlen, 0, (size_t)ix->cap_dets * MAXE * sizeof(int32_t), st));
  dim3 grid_d(4, S);
  k_block_degree<<<grid_d, 256, 0, st>>>(segs, g->src, g->cap_rows, ix->det_of_row, ix->n_dets, cnt, futlen, g->status);
  TMPNN_LAUNCH_CHECK();
  k_block_futoff<<<tmpnn_div_up(ix->cap_dets, 256), 256, 0, st>>>(ix->n_dets, futlen, cnt, segs, ix->det_rows, g->cap_rows);
  TMPNN_LAUNCH_CHECK();
  TMPNN_CUDA_TRY(scan_exclusive(cnt, ix->seg_ptr, ix->n_dets, 2, 0, 2 * (long long)ix->cap_dets, sums, st));
  // 8 warps per CTA, one list piece per warp at a time: enough CTAs per slab to fill the machine when the slabs are few
  dim3 grid_e(max(1, min(tmpnn_div_up(g->cap_rows, 256 * 8), max(64, tmpnn_div_up(TMPNN_SM_COUNT * 8, S)))), S);
  k_block_fill<<<grid_e, 256, 0, st>>>(segs, g->src, g->dst, g->cap_rows, ix->det_of_row, ix->n_dets, futlen, ix->seg_ptr, ix->inc,
                                       g->status);
  TMPNN_LAUNCH_CHECK();
  return TMPNN_OK;
}
