/* tmpnn.h -- C ABI of libtmpnn_sm100a.so: the TrackMPNN message-passing hot path on B200.
 *
 * The reference (arangesh/TrackMPNN) is pure Python/PyTorch and has no FFI of its own; the
 * entry points below are what a binding for its hot path would call.  Each one names the
 * reference code it replaces (paths relative to the reference repo).  INTEGRATION.md shows
 * the ctypes stub and the sys.modules overlay that put them behind the reference's own
 * module names (models.track_mpnn, models.layers, models.loss, utils.graph).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  Every pointer is a DEVICE pointer unless
 *    the parameter name ends in _host.  The caller owns all memory (in practice PyTorch's
 *    caching allocator); the library allocates nothing that outlives a call.
 *  - Every function returns 0 on success or a negative tmpnn_status; tmpnn_last_error()
 *    returns a thread-local message.  Nothing throws.
 *  - The last argument is the cudaStream_t (as void*) the work is enqueued on.  No function
 *    synchronises the device unless its name ends in _sync.
 *  - Re-entrant; one stream per call; safe for one host thread per GPU.
 *  - Quantities that change every frame (row counts, detection counts) live in device
 *    memory so a whole frame can be enqueued -- or captured in a CUDA graph -- without the
 *    host ever reading them.  Capacity overflows raise a sticky flag in tmpnn_graph.status.
 *
 * Data layout (device resident, structure-of-arrays, never a dense N x N matrix)
 *  The window graphs of S independent sequences live in S slabs of cap_rows rows each;
 *  sequence s owns global rows [s*cap_rows, s*cap_rows + n_rows[s]).  A row is either a
 *  detection row (ts >= 0) or an association ("edge") row (ts == -1) joining an earlier
 *  detection row src to a later one dst, with src < row < dst (slab-local indices).  This
 *  is the reference's y_pred[N,3] = [ts, det_id, ass_id] plus the two non-zeros of the
 *  row of node_adj (utils/graph.py:137-163).  The hidden state h is [S*cap_rows, ldh] fp32
 *  row-major with feature group g in columns [64 g, 64 g + 64) (= the reference's
 *  states[N, G*H], models/track_mpnn.py:72).
 */
#ifndef TMPNN_H_
#define TMPNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMPNN_HIDDEN 64 /* --num-hidden-feats (utils/training_options.py:22); kernels are specialised for it */

typedef enum tmpnn_status {
  TMPNN_OK = 0,
  TMPNN_E_BADARG = -1,
  TMPNN_E_CAPACITY = -2,
  TMPNN_E_CUDA = -3,
  TMPNN_E_UNSUPPORTED = -4
} tmpnn_status;

/* bits of the sticky device-side status word */
#define TMPNN_FLAG_ROW_CAPACITY 1  /* a slab ran out of rows in tmpnn_graph_append        */
#define TMPNN_FLAG_DET_CAPACITY 2  /* more detection rows than tmpnn_index.cap_dets        */
#define TMPNN_FLAG_INC_CAPACITY 4  /* more incidences than tmpnn_index.cap_inc             */
#define TMPNN_FLAG_SEG_CAPACITY 8  /* a detection has more incident edges than a CTA sorts */
#define TMPNN_FLAG_MULTI_GT_EDGE 16 /* "More than one GT edge from same node!" (utils/graph.py:243) */
#define TMPNN_FLAG_WALK_CAPACITY 32 /* more detection rows in one window than the decode walk holds */
#define TMPNN_FLAG_TC_TIMEOUT 64    /* tensor-core kernel: an mbarrier wait timed out (results invalid) */
#define TMPNN_FLAG_UNSTRUCTURED 256 /* tmpnn_index_build_structured: the window graph is not a chain of dense edge blocks */
#define TMPNN_FLAG_TC_RANGE 128     /* tensor-core kernel: |value| > 6e4 would overflow the fp16 split; use the FMA path */
#define TMPNN_NOTE_TC_RANGE_RERUN 512 /* not an error: a step hit TMPNN_FLAG_TC_RANGE and was re-run by the fp32 FMA kernel */

typedef struct tmpnn_graph {
  int32_t num_seqs;   /* S */
  int32_t cap_rows;   /* rows per slab */
  int32_t *n_rows;    /* [S]   rows in use per sequence */
  int32_t *ts;        /* [S*cap_rows]  timestamp, -1 for edge rows          (y_pred[:,0]) */
  int32_t *det;       /* [S*cap_rows]  detection id, -1 for edge rows        (y_pred[:,1]) */
  int32_t *ass;       /* [S*cap_rows]  associated detection id or -1         (y_pred[:,2]) */
  int32_t *src;       /* [S*cap_rows]  slab-local row of the earlier endpoint, -1 for detections */
  int32_t *dst;       /* [S*cap_rows]  slab-local row of the later endpoint,   -1 for detections */
  int32_t *label;     /* [S*cap_rows]  ground-truth class, may be NULL        (labels[N]) */
  float *score;       /* [S*cap_rows]  p = scores[:,1] */
  float *logit;       /* [S*cap_rows]  */
  int32_t *status;    /* [1] sticky TMPNN_FLAG_* bits */
  /* Deferred compaction of h (all NULL = rows of h are the logical rows).  When set, the state the next
   * message-passing step consumes is NOT stored at the logical rows: row i lives at GLOBAL row phys[i] of
   * the step's input buffer, its endpoints at psrc[i] / pdst[i] (-1 for detection rows).  The step writes
   * its output densely at the logical rows of the other buffer, so tmpnn_graph_compact only has to emit
   * these maps (the positions before the deletion) instead of moving 512 B per surviving row, and
   * tmpnn_graph_append points new edge rows at an all-zero row (slab-local row cap_rows-1, never written)
   * instead of zero-filling them.  phys_end[s]: slab-local end of the dense part of the input buffer. */
  int32_t *phys;      /* [S*cap_rows] */
  int32_t *psrc;      /* [S*cap_rows] */
  int32_t *pdst;      /* [S*cap_rows] */
  int32_t *phys_end;  /* [S] */
} tmpnn_graph;

/* Per-step index over the slabs: detection-row list, per-detection incidence lists (CSR,
 * "past" edges with dst == d first, then "future" edges with src == d, each ascending), and
 * the tile table of the row-tiled kernels.  Rebuilt by tmpnn_index_build whenever rows were
 * appended or deleted.  It replaces the reference's edge_adj = node_adj^T
 * (utils/graph.py:158,300) and every np.where(node_adj[:, i]) column scan. */
typedef struct tmpnn_index {
  int32_t cap_dets;      /* capacity of det_rows (all sequences together) */
  int32_t cap_inc;       /* capacity of inc (2 x edge rows) */
  int32_t *n_dets;       /* [1]  */
  int32_t *n_edges;      /* [1]  total edge rows (for throughput accounting) */
  int32_t *det_rows;     /* [cap_dets] global row ids, ascending */
  int32_t *det_of_row;   /* [S*cap_rows] position in det_rows, -1 for edge rows */
  int32_t *seq_det_ptr;  /* [S+1] range of det_rows per sequence */
  int32_t *seg_ptr;      /* [2*cap_dets+1] past segment of detection k = [seg_ptr[2k], seg_ptr[2k+1]), future = [seg_ptr[2k+1], seg_ptr[2k+2]) */
  int32_t *inc;          /* [cap_inc] global edge-row ids */
  int32_t *tile_ptr;     /* [S+1] prefix sum of ceil(n_rows[s] / TMPNN_TILE_ROWS) */
  int32_t *tile128_ptr;  /* [S+1] prefix sum of ceil(n_rows[s] / 128): tiles of the tensor-core kernel */
  int32_t *scratch;      /* [tmpnn_index_scratch_ints(...)] */
} tmpnn_index;

#define TMPNN_TILE_ROWS 64

const char *tmpnn_last_error(void);
int tmpnn_version(void);
/* One-time per process: opts the kernels that need > 48 KB of shared memory in.  Entry points
 * call it lazily; call it yourself before capturing a CUDA graph. */
int tmpnn_init(void);

/* ---- parameters ---------------------------------------------------------------------- */

/* Floats in one packed GRU cell (kx = input width: 64, or 128 for the edge cell with
 * --msg-type concat).  Layout: w_ih as [kx][3][64], w_hh as [64][3][64], bias [4][64]
 * (b_ir+b_hr, b_iz+b_hz, b_in, b_hn), head weight slice [64], head bias [1] (+3 pad). */
size_t tmpnn_gru_pack_floats(int kx);

/* Re-lays one torch.nn.GRUCell (models/layers.py:59-68; gate order r,z,n) plus the 64-wide
 * slice of the output head that scores this row type (models/track_mpnn.py:35-41,73) into
 * the layout above. */
int tmpnn_pack_gru(const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh,
                   const float *head_w, const float *head_b, int kx, float *packed, void *stream);

/* ---- K0: input transform (models/track_mpnn.py:45-52,59-61) ---------------------------- */

/* Linear(f_in,64) on n detection feature rows: a[i,:] = W1 x[x_idx ? x_idx[i] : i, col0:col0+f_in] + b1.
 * n = *n_dev when n_dev != NULL, else n_host. */
int tmpnn_input_linear1(const float *x, int ldx, int col0, int f_in, const int32_t *x_idx,
                        const float *w1, const float *b1, float *a, const int32_t *n_dev, int n_host,
                        void *stream);

/* Train-mode BatchNorm1d statistics over n detection rows plus n_edge_rows identical rows
 * of value b1 (the all-zero edge rows the reference also feeds, track_mpnn.py:59).  Writes
 * mean/biased var to stats[2][64] and updates running_mean/var (momentum 0.1, unbiased). */
int tmpnn_input_bn_stats(const float *a, int n, int n_edge_rows, const float *b1, float *stats,
                         float *running_mean, float *running_var, void *stream);

/* BatchNorm (given mean/var) -> ReLU -> Linear(64,64); writes h[out_rows[i], col:col+64]. */
int tmpnn_input_bn_relu_linear2(const float *a, const float *mean, const float *var, const float *gamma,
                                const float *beta, const float *w2, const float *b2, float *h, int ldh,
                                int col, const int32_t *out_rows, const int32_t *n_dev, int n_host,
                                void *stream);

/* ---- index --------------------------------------------------------------------------- */
size_t tmpnn_index_scratch_ints(int num_seqs, int cap_rows, int cap_dets);
/* active (may be NULL): sequences with active[s] == 0 are indexed as empty (they sit out this tick). */
int tmpnn_index_build(const tmpnn_graph *g, const tmpnn_index *ix, const int32_t *active, void *stream);

/* Same index for window graphs that were only ever changed by tmpnn_graph_append and the deletion of
 * tmpnn_graph_decode (the TrackEngine loop; never tmpnn_graph_prune_mask): every edge block is then a
 * dense [sources x detections-of-the-next-frame] matrix and the incidence lists follow from the block
 * boundaries -- no atomics, no sort, coalesced writes only (csrc/graph_index.cu).  A graph that breaks
 * the structure raises TMPNN_FLAG_UNSTRUCTURED.  scratch2: tmpnn_index_structured_scratch_bytes(), zero-filled
 * before its first use (the per-slab boundary counters live there and are left at zero by every call). */
size_t tmpnn_index_structured_scratch_bytes(int num_seqs, int cap_dets);
int tmpnn_index_build_structured(const tmpnn_graph *g, const tmpnn_index *ix, const int32_t *active, void *scratch2,
                                 void *stream);

/* ---- K1: aggregation (models/layers.py:90-95,103) -------------------------------------- */

/* edge_support for every detection row k of the index: agg[k,:] = sum_{future e} h[e] - sum_{past e} h[e]
 * over columns [col, col+64) of h (the sparse product edge_adj_norm . h, layers.py:103). */
int tmpnn_aggregate_dets(const tmpnn_graph *g, const tmpnn_index *ix, const float *h, int ldh, int col,
                         float *agg, void *stream);

/* tmpnn_aggregate_dets for a graph indexed by tmpnn_index_build_structured (index_scratch2 = that call's scratch2, which
 * holds the slab's segment tables and run offsets): the window graph is a chain of dense [sources x detections] edge
 * blocks, so one pass over each block yields the row sums (every source's future run) and per-stripe column partials
 * (every detection's past edges), and every association row is read exactly once -- the incidence-list form requests each
 * row twice.  Blocks appended this frame (deferred compaction: rows aliasing the slab's zero row) are skipped.  Same sums up
 * to fp32 re-association, bit-reproducible.  Per slab the scratch holds cap_runs run sums (>= detections x blocks a
 * detection can be a source of, i.e. window size) and cap_cpart column partials (>= cap_rows / 32 + detections) of 64
 * floats; exceeding either sets TMPNN_FLAG_INC_CAPACITY.  Replaces the sparse product edge_adj_norm . h of
 * models/layers.py:103 like tmpnn_aggregate_dets. */
size_t tmpnn_aggregate_blocks_scratch_bytes(int num_seqs, int cap_runs, int cap_cpart);
int tmpnn_aggregate_dets_blocks(const tmpnn_graph *g, const tmpnn_index *ix, const void *index_scratch2, const float *h,
                                int ldh, int col, float *agg, void *scratch, int cap_runs, int cap_cpart, void *stream);

/* node_support for every edge row (layers.py:91-95): diff -> h[src]-h[dst] (64 wide),
 * concat -> [h[src] | h[dst]] (128 wide); rows of support that belong to detections are zero.
 * Stand-alone form of what tmpnn_mp_step_fwd fuses; kept for the roofline study. */
int tmpnn_aggregate_edges(const tmpnn_graph *g, const tmpnn_index *ix, const float *h, int ldh, int col,
                          int concat, float *support, void *stream);

/* Attention heads (models/layers.py:7-46, 105-112; --num-att-heads > 0), eval mode: for head `head` of
 * `num_heads`, e_j = LeakyReLU_0.2(a . |W_att^T h[src_j] - W_att^T h[dst_j]|) per edge row, softmax over the
 * edges incident to each detection, agg[k] (+)= (1/num_heads) sum_j alpha_j (+-1) h[j] -- the node GRU's input
 * in place of tmpnn_aggregate_dets.  Call once per head with head = 0 first.  hatt: cap_dets*64 floats,
 * escore: S*cap_rows floats (scratch); alpha (nullable): cap_inc floats, the attention weight of every
 * incidence entry of this head (what the reference returns as a dense N x N matrix). */
int tmpnn_gat_aggregate_dets(const tmpnn_graph *g, const tmpnn_index *ix, const float *h, int ldh, int col,
                             const float *w_att, const float *a, int head, int num_heads, float *hatt,
                             float *escore, float *agg, float *alpha, void *stream);

/* Train-mode form (single-slab training graph): nn.Dropout(0.5) on the attention (layers.py:24,37) through the
 * caller's keep mask -- keep[i] != 0 keeps incidence entry i with weight alpha_i * keep_scale (= 1 / (1 - p)),
 * keep == NULL is eval mode.  Stores what tmpnn_gat_bwd needs: hatt (cap_dets*64, per head), escore (cap_rows, per
 * head), alpha (cap_inc, before dropout) and att_edge[2 e + side] (2*cap_rows): the weight after dropout that edge
 * row e has in the list of its src (side 0) / dst (side 1) detection. */
int tmpnn_gat_aggregate_dets_train(const tmpnn_graph *g, const tmpnn_index *ix, const float *h, int ldh, int col,
                                   const float *w_att, const float *a, int head, int num_heads, const uint8_t *keep,
                                   float keep_scale, float *hatt, float *escore, float *agg, float *alpha,
                                   float *att_edge, void *stream);

/* Backward of one head (autograd of models/layers.py:26-43 on the edge list): given dagg [cap_dets][64] (gradient of
 * the node GRU's input) ADDS this head's share to dh_in (columns [col, col+64): edge rows through the weighted sum,
 * detection rows through W_att), dw_att [64][64] and da [64].  Scratch: dal cap_inc, de_side 2*n_rows, dpre n_rows,
 * dhatt cap_dets*64 floats.  h is the state the step consumed. */
int tmpnn_gat_bwd(const tmpnn_graph *g, const tmpnn_index *ix, int n_rows, const float *h, int ldh, int col,
                  const float *w_att, const float *a, int num_heads, const uint8_t *keep, float keep_scale,
                  const float *hatt, const float *escore, const float *alpha, const float *att_edge,
                  const float *dagg, float *dal, float *de_side, float *dpre, float *dhatt, float *dh_in,
                  float *dw_att, float *da, void *stream);

/* ---- K2+K3: one message-passing step (models/layers.py:84-116 + track_mpnn.py:73-75) ---- */

/* For feature group `group` (columns [64 group, 64 group + 64) of h): every edge row runs the
 * edge GRUCell on h[src]-h[dst] (or the concat), every detection row runs the node GRUCell on
 * agg (from tmpnn_aggregate_dets); both read h_in and write h_out (Jacobi update).  The output
 * heads are fused: logit (+)= w . h' (+ b when group == 0) and, for the last group,
 * score = sigmoid(logit).  agg is scratch of cap_dets*64 floats. */
int tmpnn_mp_step_fwd(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                      int group, int num_groups, int concat, const float *edge_pack, const float *node_pack,
                      float *agg, void *stream);

/* The two row-type halves of tmpnn_mp_step_fwd, separately launchable (bench.py times the
 * dominant edge-row kernel alone; tmpnn_mp_det_fwd needs agg from tmpnn_aggregate_dets). */
int tmpnn_mp_edge_fwd(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                      int group, int num_groups, int concat, const float *edge_pack, void *stream);
int tmpnn_mp_det_fwd(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                     int group, int num_groups, const float *node_pack, const float *agg, void *stream);

/* The detection rows' half on the tensor cores, through the edge-step kernel of tmpnn_mp_edge_fwd_tc_pre in "detection mode":
 * tiles over the detection segments of every slab (index_scratch2 = the scratch2 of tmpnn_index_build_structured), x = agg
 * (from tmpnn_aggregate_dets[_blocks]) enters as fp16 hi / lo images written into det_img at the detections' rows, the
 * weights are the NODE cell's image (tmpnn_pack_gru_tc of the node GRU + detection head, concat = 0).  Call it after the
 * edge step of the same group: det_img / det_p (the edge step's scratch) are dead by then and reused.  det_tile_table:
 * tmpnn_tc_det_tile_table_bytes() of scratch.  Sets TMPNN_FLAG_TC_RANGE like the edge step; tmpnn_mp_det_fwd_on_flag is the
 * fp32 FMA re-run (tmpnn_mp_det_fwd that only runs when the status word holds one of flag_mask). */
size_t tmpnn_tc_det_tile_table_bytes(int num_seqs, int cap_dets);
int tmpnn_mp_det_fwd_tc(const tmpnn_graph *g, const tmpnn_index *ix, const void *index_scratch2, const float *h_in,
                        float *h_out, int ldh, int group, int num_groups, const void *node_image, const float *agg,
                        float *det_img, float *det_p, void *det_tile_table, void *stream);
int tmpnn_mp_det_fwd_on_flag(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                             int group, int num_groups, const float *node_pack, const float *agg, int flag_mask,
                             void *stream);

/* The batched engine's device-side work counters and frame clock in one launch (no reference equivalent; bench.py's
 * edge-updates/s and frames/s are read from them): *edge_updates += *n_edges, *det_updates += *n_dets, *frames_done += sum of
 * active[0..num_seqs), *t_dev += 1.  Every (destination, source) pair may be NULL. */
int tmpnn_graph_counters(int64_t *edge_updates, const int32_t *n_edges, int64_t *det_updates, const int32_t *n_dets,
                         int64_t *frames_done, const int32_t *active, int num_seqs, int32_t *t_dev, void *stream);

/* tmpnn_mp_edge_fwd that only runs when the sticky status word holds one of flag_mask at launch time (the test is made on
 * the device, so the call can sit in a CUDA graph): the engine enqueues it behind the tensor-core step with
 * flag_mask = TMPNN_FLAG_TC_RANGE, which re-runs a step whose activations left the fp16 split's range on the fp32 FMA
 * kernel (both are this library's kernels; for num_groups > 1 re-run every group, group 0 first: the logit accumulates
 * over the groups), followed by tmpnn_status_ack(g, TMPNN_FLAG_TC_RANGE, TMPNN_NOTE_TC_RANGE_RERUN). */
int tmpnn_mp_edge_fwd_on_flag(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                              int group, int num_groups, int concat, const float *edge_pack, int flag_mask, void *stream);

/* Tensor-core form of tmpnn_mp_edge_fwd for msg_type 'diff' (tcgen05.mma kind::f16, 3-term fp16
 * split of activations and weights, accumulators in TMEM; csrc/mp_step_tc.cu).  edge_image is the
 * tmpnn_gru_tc_pack_bytes() byte image written by tmpnn_pack_gru_tc (fp16 hi/lo weights in the
 * 128B-swizzled K-major UMMA layout + fused biases + head slice).  Same outputs as the FMA
 * kernel to ~2e-6; sets TMPNN_FLAG_TC_RANGE if an activation exceeds the fp16 range. */
size_t tmpnn_gru_tc_pack_bytes(void);
int tmpnn_pack_gru_tc(const float *w_ih, const float *w_hh, const float *b_ih, const float *b_hh,
                      const float *head_w, const float *head_b, int concat, void *packed, void *stream);
int tmpnn_mp_edge_fwd_tc(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                         int group, int num_groups, const void *edge_image, void *stream);
/* Same step with the endpoints prepared once per DETECTION row instead of once per association row (both
 * msg_types; edge_image packed with the same `concat`).  gi = h[src] W_s^T -+ h[dst] W_d^T is linear in the two
 * endpoints: a first kernel writes every detection's fp16 hi/lo image into det_img (same geometry as h: S*cap_rows
 * rows of ldh floats, indexed by logical row; scratch, any content) and its source-side gate contribution with the
 * biases folded in into det_p [index cap_dets][192] (fp32 FMA); the step kernel then copies the far endpoint's
 * image (cp.async), multiplies it on the tensor cores (negated for 'diff') and adds det_p[src] in the epilogue.
 * The source-side half of weight_ih (transposed, fp32) and the folded biases travel in the tail of edge_image.
 * The step kernel (csrc/mp_step_tc3.cu) has a dedicated MMA issuer warp and two epilogue teams on alternating tiles;
 * it locates tiles through a table the call rebuilds (when group == 0) in tile_table, a 16-byte aligned scratch of
 * tmpnn_tc_tile_table_bytes(). */
size_t tmpnn_tc_tile_table_bytes(int num_seqs, int cap_rows);
int tmpnn_mp_edge_fwd_tc_pre(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                             int group, int num_groups, int concat, const void *edge_image, float *det_img,
                             float *det_p, void *tile_table, void *stream);

/* ---- training: backward of the step and the losses (train.py:65-134, models/loss.py) ------- */

/* tmpnn_mp_step_fwd that also stores, per row, r | z | n | (W_hn h + b_hn) of feature group `group`
 * into gates[row][4][64] -- what loss.backward() needs from torch.nn.GRUCell (models/layers.py:97,114). */
int tmpnn_mp_step_fwd_train(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                            int group, int num_groups, int concat, const float *edge_pack, const float *node_pack,
                            float *agg, float *gates, void *stream);

/* The same with the detection aggregates supplied by the caller (attention heads: tmpnn_gat_aggregate_dets_train). */
int tmpnn_mp_step_fwd_train_agg(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                                int group, int num_groups, int concat, const float *edge_pack, const float *node_pack,
                                const float *agg, float *gates, void *stream);

/* The training forward on the tensor cores (msg_type 'diff'): tmpnn_mp_edge_fwd_tc that also stores the gates of every
 * association row, and the detection-row half of tmpnn_mp_step_fwd_train alone (agg from tmpnn_aggregate_dets).  The three
 * calls tmpnn_aggregate_dets, tmpnn_mp_edge_fwd_tc_train, tmpnn_mp_det_fwd_train replace tmpnn_mp_step_fwd_train on large
 * graphs (states to ~2e-6 of the FMA kernel's). */
int tmpnn_mp_edge_fwd_tc_train(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                               int group, int num_groups, const void *edge_image, float *gates, void *stream);
int tmpnn_mp_det_fwd_train(const tmpnn_graph *g, const tmpnn_index *ix, const float *h_in, float *h_out, int ldh,
                           int group, int num_groups, const float *node_pack, const float *agg, float *gates, void *stream);

/* Gate gradients of one feature group for every row of a single-slab graph (row type from src):
 * dh' = dh_out (nullable) + (dlogits + dscores p (1-p)) w_type;  dgi = [dpr,dpz,dpn], dgh = [dpr,dpz,dpn r]
 * ([n][192] each), dhself = dh' z ([n][64]).  Accumulates (atomically) the bias gradients of both cells
 * (gbias_*: [2][192] = d bias_ih | d bias_hh), the 64-wide head weight slices and, when the pointers are
 * non-null (pass them for group 0 only), the head biases. */
/* partials: NULL = the CTAs add their bias / head-weight sums atomically (order not reproducible); else
 * tmpnn_gate_bwd_partial_floats() floats of scratch: per-CTA sums, added in CTA order by a second kernel (bit-reproducible). */
size_t tmpnn_gate_bwd_partial_floats(void);
int tmpnn_gate_bwd(int n_rows, const int32_t *src, const float *gates, const float *h_prev, const float *h_new,
                   int ldh, int col, const float *dh_out, const float *dlogits, const float *dscores,
                   const float *score, const float *head_w_edge, const float *head_w_node, float *dgi, float *dgh,
                   float *dhself, float *gbias_edge, float *gbias_node, float *ghw_edge, float *ghw_node,
                   float *ghb_edge, float *ghb_node, float *partials, void *stream);

/* C[rc(i)][0:n] (+)= A[ra(i)][0:192] . W[192][n] for rows i < R (R = *r_dev or r_host) whose
 * mask[ra(i)] >= 0; ra = a_rows ? a_rows[i] : i, rc likewise.  n = 64 or 128.  (dx = dgi . W_ih,
 * dh_self += dgh . W_hh with the torch.nn.GRUCell weight layouts [192][n].) */
int tmpnn_rows_times_w(const int32_t *r_dev, int r_host, const int32_t *a_rows, const int32_t *c_rows,
                       const int32_t *mask, const float *A, const float *W, int n, float *C, int ldc,
                       int accumulate, void *stream);

/* G[192][n] += sum_i A[ra(i)][0:192]^T B[rb(i)][0:n]  (weight gradients dW_ih += dgi^T x, dW_hh += dgh^T h). */
int tmpnn_rows_outer(const int32_t *r_dev, int r_host, const int32_t *a_rows, const int32_t *b_rows,
                     const int32_t *mask, const float *A, const float *B, int ldb, int n, float *G, void *stream);

/* Both contractions above for one operand A in ONE pass on the tensor cores (csrc/train_tc.cu: tcgen05.mma kind::f16 on a
 * 3-term bf16 split, accumulators in TMEM; the [128 rows x 64] shared-memory images of A and X serve as K-major operands of
 * C = A W and, read as MN-major, as the operands of G = A^T X, so nothing is transposed):
 *   C[rc(i)][0:64] (+)= A[ra(i)][0:192] . W[192][col0:col0+64],   G[192][0:64] (leading dimension ldg) += sum_i A[ra(i)]^T X[rx(i)][0:64]
 * for i < R (= *r_dev or r_host) with mask[ra(i)] >= 0; ra / rc / rx = a_rows / c_rows / x_rows[i] or i.  w_image: W packed
 * by tmpnn_pack_w_tc (tmpnn_bwd_tc_image_bytes() bytes); partials: scratch of tmpnn_bwd_tc_partial_floats() floats (one G
 * partial per SM, summed in a fixed order: the weight gradients are reproducible bit for bit, unlike tmpnn_rows_outer's float
 * atomics).  A 128-wide W (msg_type concat) takes two calls, col0 = 0 and 64.  status: a device word for the bounded
 * barrier waits (TMPNN_FLAG_TC_TIMEOUT). */
size_t tmpnn_bwd_tc_image_bytes(void);
size_t tmpnn_bwd_tc_partial_floats(void);
int tmpnn_pack_w_tc(const float *W, int ldw, int col0, void *image, void *stream);
int tmpnn_rows_gemm_tc(const int32_t *r_dev, int r_host, const int32_t *a_rows, const int32_t *c_rows, const int32_t *x_rows,
                       const int32_t *mask, const float *A, const void *w_image, float *C, int ldc, int accumulate,
                       const float *X, int ldx, float *partials, float *G, int ldg, int32_t *status, void *stream);

/* Transpose of the gather / segmented sum (models/layers.py:90-95,103) in gather form:
 * edge row e: dh_in[e] = dhself[e] + dagg[det(src)] - dagg[det(dst)];
 * detection d: dh_in[d] = dhself[d] + sum over its future edges of dx[e][0:64] -/+ sum over its past edges of
 * dx[e][0:64] (diff) / dx[e][64:128] (concat).  dagg is indexed by detection-list position. */
int tmpnn_scatter_bwd(const tmpnn_graph *g, const tmpnn_index *ix, int n_rows, const float *dhself, const float *dx,
                      int kx, const float *dagg, float *dh_in, int ldh, int col, void *stream);

/* Backward of Linear -> BatchNorm1d -> ReLU -> Linear (models/track_mpnn.py:45-52) on the n new detection
 * rows of a step; `a`, mean, var are what the forward produced (batch statistics over n + n_edge_rows rows
 * when training).  dh rows are dh[out_rows[i]][col:col+64].  Gradients are added
 * to gw1[64][f_in], gb1, ggamma, gbeta, gw2[64][64], gb2.  scratch: 2 n 64 floats. */
int tmpnn_input_bwd(const float *x, int ldx, int col0, int f_in, const int32_t *x_idx, const float *a,
                    const float *mean, const float *var, const float *gamma, const float *beta, const float *b1,
                    const float *w2, const float *dh, int ldh, int col, const int32_t *out_rows, int n,
                    int n_edge_rows, int training, float *scratch, float *gw1, float *gb1, float *ggamma,
                    float *gbeta, float *gw2, float *gb2, void *stream);

/* The same for several groups of rows in one launch, one CTA per group -- the chunks of the batched trainer, each its own
 * BatchNorm batch (trackmpnn_b200/train_engine.py).  groups: HOST array of n_groups descriptors (device pointers inside;
 * they travel in the kernel parameters, 32 per launch); the gradient buffers are shared and added to atomically. */
typedef struct tmpnn_input_group {
  const float *a, *mean, *var;      /* Linear1 outputs [n][64] and the batch statistics [64] of this group */
  const int32_t *x_idx, *out_rows;  /* feature rows / state rows of its n detections */
  float *scratch;                   /* 2 n 64 floats */
  int32_t n, n_edge_rows;
} tmpnn_input_group;
/* Forward counterpart for train-mode BatchNorm: per group the batch statistics of its Linear1 outputs g.a (written to
 * g.mean / g.var), the running statistics updated group by group in order, then BatchNorm -> ReLU -> Linear2 into
 * h[g.out_rows][col:col+64].  (Linear1 itself is row-wise: one tmpnn_input_linear1 call over all groups' rows.) */
int tmpnn_input_bn_groups_fwd(const tmpnn_input_group *groups, int n_groups, const float *b1, const float *gamma,
                              const float *beta, const float *w2, const float *b2, float *running_mean,
                              float *running_var, float *h, int ldh, int col, void *stream);
/* partials: NULL = the groups add into the gradient buffers atomically; else tmpnn_input_bwd_partial_floats(n_groups)
 * floats of scratch: every group stores its own partial gradients and a second kernel adds them in group order (no
 * contention on the 4 k shared addresses, and the weight gradients of the input transform come out reproducible bit for
 * bit). */
size_t tmpnn_input_bwd_partial_floats(int n_groups);
int tmpnn_input_bwd_groups(const float *x, int ldx, int col0, int f_in, const tmpnn_input_group *groups, int n_groups,
                           const float *gamma, const float *beta, const float *b1, const float *w2, const float *dh,
                           int ldh, int col, int training, float *gw1, float *gb1, float *ggamma, float *gbeta,
                           float *gw2, float *gb2, float *partials, void *stream);

/* create_targets (models/loss.py:8-44) on a single-slab graph with labels: targets[N] int32. */
int tmpnn_loss_targets(const tmpnn_graph *g, const tmpnn_index *ix, int n_rows, int32_t *targets, void *stream);

/* CELoss (models/loss.py:81-115): per detection and per past / future incidence segment holding a positive
 * target, (logsumexp(logit[segment]) - logit[chosen positive]) / len.  seg_* have 2 cap_dets entries and
 * are what tmpnn_loss_ce_bwd needs; loss[0] = the sum (deterministic order). */
int tmpnn_loss_ce_fwd(const tmpnn_index *ix, int n_rows, const int32_t *targets, const float *logit, float *seg_lse,
                      int32_t *seg_pos, float *seg_loss, float *loss, void *stream);
int tmpnn_loss_ce_bwd(const tmpnn_graph *g, const tmpnn_index *ix, int n_rows, const float *seg_lse,
                      const int32_t *seg_pos, const float *logit, const float *grad_out, float *dlogit, void *stream);

/* FocalLoss(gamma=0, alpha=None, size_average=True) (models/loss.py:57-74): mean(-log(p_t + 1e-10)). */
int tmpnn_loss_focal_fwd(int n, const float *p, const int64_t *targets, float *per_elem, float *loss, void *stream);
int tmpnn_loss_focal_bwd(int n, const float *p, const int64_t *targets, const float *grad_out, float *dp, void *stream);

/* Batched trainer (trackmpnn_b200/train_engine.py): sum_i w[i] (-log(p_t[i] + 1e-10)) -- the BCE terms of train.py:76-85 for
 * B chunks at once, w[i] = 1 / (rows of row i's kind in its chunk), i.e. the sum over chunks of their per-chunk means.
 * per_elem: scratch of n floats (holds block partials afterwards, not the per-row terms). */
int tmpnn_loss_wbce_fwd(int n, const float *p, const int64_t *targets, const float *w, float *per_elem, float *loss, void *stream);
int tmpnn_loss_wbce_bwd(int n, const float *p, const int64_t *targets, const float *w, const float *grad_out, float *dp, void *stream);

/* dst[seg_dst[s] + r][0:ld] = src[seg_src[s] + r][0:ld] for r < seg_len[s], s < n_seg (max_len = the longest segment, for
 * the grid): the rows every chunk carries from one step's block-diagonal layout into the next. */
int tmpnn_rows_move(const float *src, float *dst, const int32_t *seg_src, const int32_t *seg_dst, const int32_t *seg_len,
                    int n_seg, int ld, int max_len, void *stream);

/* ---- feature construction: the step in front of the path (dataset/kitti_mot.py:545-566) ---- */

/* x[i] = ([one-hot(cat_id - 1, ncat) | score, xc, yc, w, h (use_2d) | sin, cos(pi (frame mod fr_range) / fr_range)
 * (use_temp)] - mean) / std for the n_dets detection records bbox_pred[i][16] =
 * (frame, track, cat_id, alpha, x1, y1, x2, y2, h, w, l, x, y, z, rotation_y, score); the visual block of the
 * reference ('vis' features, the embedding CNN) is out of scope. */
int tmpnn_build_features(const float *bbox_pred, int n_dets, int ncat, int use_2d, int use_temp, int fr_range,
                         const float *mean, const float *std, float *x, int ldx, void *stream);

/* ---- graph bookkeeping (utils/graph.py) ------------------------------------------------ */

/* y_pred[N,3] int64 <-> ts/det/ass, scores[N,2] -> p, labels int64 -> int32 for ONE slab. */
int tmpnn_ypred_unpack(const int64_t *y_pred, int n, int32_t *ts, int32_t *det, int32_t *ass, void *stream);
int tmpnn_ypred_pack(const int32_t *ts, const int32_t *det, const int32_t *ass, int n, int64_t *y_pred, void *stream);

/* Sparse COO of node_adj (with the I_node diagonal) from the edge list of one slab and back
 * (utils/graph.py:152-163, 299-308).  idx is [2, nnz] int64, val [nnz] fp32,
 * nnz = 2*n_edges + n_dets; entries are emitted row-major (coalesced COO order). */
int tmpnn_coo_from_edges(const int32_t *ts, const int32_t *src, const int32_t *dst, int n, int transpose,
                         int64_t *idx, float *val, int64_t nnz, int32_t *nnz_prefix_scratch, void *stream);
int tmpnn_edges_from_coo(const int64_t *idx, const float *val, int64_t nnz, int n, int32_t *src, int32_t *dst,
                         void *stream);

/* Greedy association (utils/graph.py:251-268 == 437-454): for each detection row with p >= 0.5
 * take its positive future edges whose far end is positive, keep the nearest timestep, pick the
 * first arg-max; ass = det id of that far end.  mode 0 = greedy from scores, 1 = teacher forcing
 * from labels (utils/graph.py:229-245).
 * active must be the mask the index was built with (sequences sitting out keep their ass). */
int tmpnn_graph_associate(const tmpnn_graph *g, const tmpnn_index *ix, int mode, const int32_t *active,
                          void *stream);

/* Detections of the sequences, grouped by timestamp: frame_ptr[s*(t_max+2) + t] .. [+1] indexes
 * frame_dets (ascending detection ids of sequence s at time t). */
typedef struct tmpnn_frames {
  int32_t t_max;            /* timestamps 0..t_max */
  int32_t ldx;              /* floats per feature row */
  const int32_t *frame_ptr; /* [S*(t_max+2)] */
  const int32_t *frame_dets;/* [sum ND] detection ids (sequence-local) */
  const int32_t *det_ptr;   /* [S+1] offset of sequence s in the per-detection arrays */
  const int32_t *det_track; /* [sum ND] ground-truth track id (labels), may be NULL */
} tmpnn_frames;

/* Per-sequence driver state of the batched engine: the loop variables of infer.py:48-87
 * (t_skip, t_end, "re-initialise when the graph and the frame are both empty") kept on the
 * device so that many sequences advance in lock step without host round trips. */
typedef struct tmpnn_seq_state {
  int32_t *phase;      /* [S] 0 = not started, 1 = running, 2 = finished */
  int32_t *skip_until; /* [S] infer.py's t_skip */
  int32_t *t_end;      /* [S] last timestamp + 1 (initialize_graph's tN+1) */
  int32_t *active;     /* [S] out: 1 if the sequence takes part in this tick */
  int32_t *t_upto;     /* [S] out: decode_tracks' t_upto for this tick */
  int32_t *fresh;      /* [S] out: 1 if the graph was (re-)initialised this tick (states = None) */
  int32_t *last_new;   /* [S] rows the sequence's last executed iteration added (infer.py's feats.size()[0]): the loop
                          re-initialises when the graph AND that are both empty (infer.py:64-69) */
} tmpnn_seq_state;

/* Hungarian association (utils/graph.py:33-93 driven by :247-249 / :433-435): resets ass, then for every
 * timestep of each window solves the reference's cost matrix (scores[e,0] for existing edges, 100.0
 * otherwise; cost == NULL means 1 - score) with scipy.optimize.linear_sum_assignment's algorithm
 * (rectangular shortest augmenting path, same transposition and tie-breaking; csrc/hungarian.cu) and
 * keeps pairs with cost <= threshold (0.5 in the drivers).  only_t != 0 runs the single timestep t without
 * resetting ass (the stand-alone hungarian() of utils/graph.py:33).  max_dets bounds the detection rows
 * of one window; scratch holds tmpnn_hungarian_scratch_bytes(num_seqs, max_dets) bytes. */
size_t tmpnn_hungarian_scratch_bytes(int num_seqs, int max_dets);
int tmpnn_graph_associate_hungarian(const tmpnn_graph *g, const tmpnn_index *ix, const float *cost,
                                    const int32_t *active, int max_dets, int only_t, int t, float threshold,
                                    void *scratch, void *stream);

/* The assignment solver alone: `batch` independent nr x nc fp32 cost matrices -> col_of_row[batch][nr]
 * (-1 for unassigned rows of a tall matrix); equals scipy.optimize.linear_sum_assignment(C). */
size_t tmpnn_lsap_scratch_bytes(int batch, int nr, int nc);
int tmpnn_lsap_solve(const float *cost, int nr, int nc, int batch, int32_t *col_of_row, void *scratch, void *stream);

/* update_graph steps 2-4 (utils/graph.py:270-327) and initialize_graph (utils/graph.py:96-186)
 * for every sequence at timestep t = *t_dev.
 *  Append: active set (mode 0: detection && ass == -1 && p >= 0.5; mode 1 (train):
 *  (detection && ass == -1) || ts == t_prev), then A*Nt edge rows and Nt detection rows,
 *  edge (a,j) at n + a*Nt + j with src = active[a], dst = n + A*Nt + j; labels from det_track.
 *  Initialise: the first two non-empty timesteps t0 < t1 at or after t give rows
 *  [N0 dets][N0*N1 edges][N1 dets], edge (i,j) at N0 + i*N1 + j.
 * st == NULL: plain append for every sequence (the drop-in update_graph).  With st: start != 0
 * initialises every sequence from t (t = 0 when t_dev == NULL; infer.py:48); start == 0 runs one iteration of
 * infer.py:60-74 per sequence (skip / re-initialise / append) and sets st->active, st->t_upto
 * (cur_win_size as in infer.py:86), st->fresh.
 * New rows get h = 0 (ldh floats per row; fresh sequences are zeroed completely).
 * new_det_rows / new_det_x (capacity cap_new) receive the global row and the feature-row index
 * (det_ptr[s] + detection id) of every new detection row, n_new[0] their count, n_new[1] the
 * number of new edge rows; n_appended[s] the rows added to sequence s. */
int tmpnn_graph_append(const tmpnn_graph *g, const tmpnn_frames *fr, const tmpnn_seq_state *st,
                       const int32_t *t_dev, int start, int cur_win_size, int mode, float *h, int ldh,
                       int32_t *new_det_rows, int32_t *new_det_x, int32_t *n_new, int cap_new,
                       int32_t *n_appended, int32_t *scratch, void *stream);
size_t tmpnn_graph_append_scratch_ints(int num_seqs, int cap_rows);

/* decode_tracks walk (utils/graph.py:456-490) + deletion mask (:492-512) for every sequence:
 * y_out_track[det_ptr[s] + d] is the reference's y_out[d,1]; next_track_id[s] its running
 * max+1.  t_upto = t_upto_seq[s] when given, else t_upto_host.  Sequences with active[s] == 0
 * (when active != NULL) keep every row.  Writes keep[row] in {0,1} for all rows in use.
 * Needs a fresh index and tmpnn_graph_associate first.  max_dets: the caller's bound of the detection rows one
 * window can hold (<= 0: unknown, windows of up to 4096 detection rows); the walk's per-detection arrays live in shared
 * memory up to 8192 detection rows and in `scratch` beyond (no limit; stress windows of BASELINE configs[3] and larger).
 * scratch: tmpnn_graph_decode_scratch_ints(num_seqs, max_dets) ints.  A window that exceeds max_dets raises
 * TMPNN_FLAG_WALK_CAPACITY. */
size_t tmpnn_graph_decode_scratch_ints(int num_seqs, int max_dets);
int tmpnn_graph_decode(const tmpnn_graph *g, const tmpnn_index *ix, const tmpnn_frames *fr,
                       int32_t *y_out_track, int32_t *next_track_id, const int32_t *t_upto_seq,
                       int t_upto_host, const int32_t *active, int ret_win_size, uint8_t *keep, int max_dets,
                       int32_t *scratch, void *stream);

/* --no-tp-classifier (infer.py:54-57, 77-80): every detection row of the index gets score p = 1 (scores row (0, 1))
 * before association and decoding; logits are left alone. */
int tmpnn_graph_force_det_scores(const tmpnn_graph *g, const tmpnn_index *ix, void *stream);

/* Sticky status word: if any of from_bits is set, clears them and sets to_bits (a condition that was handled on the
 * device -- e.g. TMPNN_FLAG_TC_RANGE after the step was re-run by tmpnn_mp_edge_fwd_on_flag -- becomes a note). */
int tmpnn_status_ack(const tmpnn_graph *g, int from_bits, int to_bits, void *stream);

/* prune_graph mask (utils/graph.py:361-377): keep = p >= thr || detection || row < first || row > last
 * detection row with t_st <= ts <= t_ed. */
int tmpnn_graph_prune_mask(const tmpnn_graph *g, const tmpnn_index *ix, int t_st, int t_ed, float threshold,
                           uint8_t *keep, int32_t *scratch /* 2*S ints */, void *stream);

/* Order-preserving deletion of the rows with keep == 0 in every slab (the np.delete /
 * advanced-indexing of utils/graph.py:379-387, 514-520): prefix-sum stream compaction of
 * ts/det/ass/label/score/logit and of h (h_src -> h_dst, ldh floats per row), src/dst
 * re-mapped to the new row numbers.  Out of place: reads g_in / h_src, writes g_out / h_dst
 * (distinct arrays; the caller ping-pongs two sets) and g_out->n_rows;
 * new_of_old[row] = new slab-local row or -1.  Sequences with active[s] == 0 (they sat out the
 * message-passing step, so their current state is still in the step's input buffer) take h from
 * h_src_inactive instead of h_src. */
/* Deferred-compaction graphs only: declares the state dense at the logical rows (phys = identity,
 * psrc/pdst = src/dst, phys_end = n_rows), e.g. after the first step on a freshly initialised graph. */
int tmpnn_graph_phys_identity(const tmpnn_graph *g, void *stream);

int tmpnn_graph_compact(const tmpnn_graph *g_in, const tmpnn_graph *g_out, const uint8_t *keep,
                        const float *h_src, const float *h_src_inactive, const int32_t *active, float *h_dst,
                        int ldh, int32_t *new_of_old, int32_t *scratch, void *stream);
size_t tmpnn_graph_compact_scratch_ints(int num_seqs, int cap_rows);

#ifdef __cplusplus
}
#endif
#endif /* TMPNN_H_ */
