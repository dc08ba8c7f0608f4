/*
 * segnn_b200.h -- C ABI of the B200-native SEGNN message-passing hot path.
 *
 * The reference (a pure-Python repository) has no FFI for this path; its boundary is the Python
 * module contract of models/segnn (SEGNN, SEGNNLayer, O3TensorProduct, O3Transform) and
 * utils/build_fully_connected_graph.py.  Every entry point below names the reference code it replaces
 * (paths are into the reference repository).  INTEGRATION.md shows the ctypes stub a maintainer adds.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless stated otherwise; tensors are dense, row-major, float32
 *     unless the name says otherwise;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing is allocated, nothing
 *     synchronises, every launch goes to `stream` => all calls are CUDA-graph capturable and re-entrant;
 *   - return value 0 = ok, negative = error (SEGNN_E_*); segnn_last_error() returns a thread-local message;
 *   - "planar" hidden features: [nodes][4][n] = (scalars, v_x, v_y, v_z) per node, n = multiplicity of the
 *     hidden irreps n x0e + n x1o.  e3nn's mul-major layout (models/segnn: [s(n) | v(u,k) at n+3u+k]) is
 *     converted at the module boundary only;
 *   - B graphs of N nodes each, fully connected without self loops: nodes = B*N, E = B*N*(N-1).  The edge
 *     list is never materialised on the hot path; enumeration order (when materialised) is the reference's:
 *     graph-major, then source ascending, then target ascending (utils/build_fully_connected_graph.py:4-20).
 */
#ifndef SEGNN_B200_H_
#define SEGNN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* segnn_stream_t; /* cudaStream_t */

enum {
  SEGNN_OK = 0,
  SEGNN_E_INVALID = -1,   /* bad argument (null pointer, unsupported size) */
  SEGNN_E_UNSUPPORTED = -2, /* configuration not built (e.g. lmax_h != 1) */
  SEGNN_E_CUDA = -3       /* CUDA runtime error at launch */
};

/* compute mode of the per-edge contraction (message_layer_2) */
enum {
  SEGNN_MODE_FP32 = 0,   /* FFMA, fp32 everywhere: the 1e-5 parity mode */
  SEGNN_MODE_BF16_TC = 1, /* tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM): the throughput mode */
  SEGNN_MODE_FP16_TC = 2, /* same kernels with fp16 operands (11-bit mantissa): 8x smaller operand rounding at the same
                             speed; operands must stay below 65504 (true for normalised features) */
  SEGNN_MODE_FP16_PACKED = 3 /* fp16 operands AND the message_layer_1 combine in packed fp16 (HFMA2 over sender pairs)
                                on fp16 projections written by segnn_node_gemm_tc_pair16: p, q are then
                                [nodes / 2][4][3n][2] fp16 arrays (passed through the float pointers), the column
                                order is that of the tensor-core modes and the (0s, 0g) parts carry a factor 1/2
                                (folded into the node-GEMM weight image).  Needs an even graph size N. */
};

/* 16-bit operand format of the tensor-core kernels and of their packed weight images */
enum { SEGNN_OPERAND_BF16 = 0, SEGNN_OPERAND_FP16 = 1 };

int segnn_version(void);
const char* segnn_last_error(void);

/* ---- graph / geometry materialisation (tests and legacy consumers only) -------------------------- */

/* utils/build_fully_connected_graph.py:4-20 _build_fully_connected_edge_index.
 * edge_index: int64 [2, B*N*(N-1)], row 0 = source (sender), row 1 = target (receiver). */
int segnn_edge_index(int B, int N, int64_t* edge_index, segnn_stream_t stream);

/* models/segnn/o3_building_blocks.py:237-245,277 (O3Transform edge part), lmax_attr = 1.
 * edge_attr [E,4] = Y_0..1(pos[src]-pos[tgt]) ('integral', normalised); add [E,2] = (|r|, m_src*m_tgt). */
int segnn_edge_attr(const float* pos, const float* mass, int B, int N, float* edge_attr, float* add,
                    segnn_stream_t stream);

/* ---- K1: per-step node geometry ------------------------------------------------------------------ */

/* models/segnn/o3_building_blocks.py:253-276 (node_attr = mean_j Y(r_j - r_i) + Y(v_i); x = [pos -
 * mean_xyz(pos), vel, |vel|]) plus models/segnn/segnn.py:148 (node_attr[:,0] = 1).
 * pos, vel [nodes,3] -> x_in [nodes,7], node_attr [nodes,4]. */
int segnn_prep_fwd(const float* pos, const float* vel, int B, int N, float* x_in, float* node_attr,
                   segnn_stream_t stream);

/* models/segnn/o3_building_blocks.py:230-278 (O3Transform) for lmax_attr in 0..2, feeding the generic-irreps kernels
 * (segnn_generic_*): edge_attr [E,(lmax_attr+1)^2] = Y_0..lmax(pos[src]-pos[tgt]) ('integral', normalised, e3nn
 * component order), add [E,2] = (|r|, m_src*m_tgt); x_in [nodes,7] as segnn_prep_fwd, node_attr
 * [nodes,(lmax_attr+1)^2] = mean_j Y(r_j - r_i) + Y(v_i) with node_attr[:,0] = 1 (models/segnn/segnn.py:148). */
int segnn_edge_attr_lmax(const float* pos, const float* mass, int B, int N, int lmax_attr, float* edge_attr, float* add,
                         segnn_stream_t stream);
int segnn_prep_fwd_lmax(const float* pos, const float* vel, int B, int N, int lmax_attr, float* x_in, float* node_attr,
                        segnn_stream_t stream);

/* Explicit edge lists (kNN graphs of utils/build_fully_connected_graph.py:42-80, num_neighbors < N - 1), generic-irreps
 * path, inference.  edge_index int64 [2,E]: row 0 = source j, row 1 = target i (PyG source_to_target).
 *   segnn_edge_attr_list              o3_building_blocks.py:237-245,277 on an edge list
 *   segnn_generic_message_input_list  segnn.py:264-279: out [E, 2D + d_add] = cat(x[target], x[source], add)
 *   segnn_segment_reduce              segnn.py:205 (sum) / o3_building_blocks.py:257-263 (mean): out[node] = reduction of
 *                                     values[order[k]] over k in [ptr[node], ptr[node+1]), fixed order, no atomics
 *   segnn_prep_fwd_list               o3_building_blocks.py:253-276 + segnn.py:148 with the scatter-mean given */
int segnn_edge_attr_list(const float* pos, const float* mass, const int64_t* edge_index, int64_t E, int lmax_attr,
                         float* edge_attr, float* add, segnn_stream_t stream);
int segnn_generic_message_input_list(const float* x, const float* add, const int64_t* edge_index, int64_t E, int D,
                                     int d_add, float* out, segnn_stream_t stream);
int segnn_segment_reduce(const float* values, const int64_t* order, const int64_t* ptr, int64_t nodes, int D, int mean,
                         float* out, segnn_stream_t stream);
int segnn_prep_fwd_list(const float* pos, const float* vel, const float* mean_attr, int64_t nodes, int lmax_attr,
                        float* x_in, float* node_attr, segnn_stream_t stream);

/* ---- K2: embedding tensor product ---------------------------------------------------------------- */

/* models/segnn/segnn.py:170 embedding_layer = O3TensorProduct(2x1o+1x0e -> h, node_attr).
 * w_embed: packed [6][n] = (W_vec0->1o, W_vec1->1o, W_vec0->0e /sqrt3, W_vec1->0e /sqrt3, W_sc->0e, W_sc->1o),
 * bias [n].  h_out planar [nodes][4][n]. */
int segnn_embed_fwd(const float* x_in, const float* node_attr, const float* w_embed, const float* bias,
                    int nodes, int n, float* h_out, segnn_stream_t stream);

/* Same kernel, also writing h16_out: an fp16 copy [nodes][4][n] of h_out (round to nearest even), the operand rows of
 * segnn_node_gemm_tc_x16.  The tensor-core node GEMMs round their fp32 inputs to the operand format on load; a
 * producer that stores the rounded copy itself halves the bytes those GEMMs read and leaves their results bit-identical
 * (models/segnn/segnn.py:170 feeds segnn.py:264-304 of the first layer). */
int segnn_embed_fwd_x16(const float* x_in, const float* node_attr, const float* w_embed, const float* bias,
                        int nodes, int n, float* h_out, void* h16_out, segnn_stream_t stream);

/* ---- node-level tensor products: plain GEMM + attribute combine ------------------------------------ */

/* The weight contraction of an O3TensorProduct over node rows (o3_building_blocks.py:150-162) hoisted out
 * of the attribute coupling:  y[node][c][:] = concat_K(x0[node][c], x1[node][c]) @ (c == 0 ? w_s : w_v),
 * c = 0 scalar plane, c = 1..3 vector planes.  x0, x1 planar [nodes][4][n_in] (x1 may be NULL);
 * w_s, w_v [K][n_out] row-major with K = n_in * (x1 ? 2 : 1); bias (may be NULL) [n_bias] is added to
 * columns [0, n_bias) of the c = 0 plane.  Output columns [0, split) go to y0 [nodes][4][split] and columns
 * [split, n_out) to y1 [nodes][4][n_out - split]; y1 == NULL => everything goes to y0 [nodes][4][n_out].
 * fp32 FFMA kernel (the 1e-5 parity mode). */
int segnn_node_gemm(const float* x0, const float* x1, int nodes, int n_in, const float* w_s, const float* w_v,
                    const float* bias, int n_bias, int n_out, float* y0, float* y1, int split,
                    segnn_stream_t stream);

/* Same contraction on tcgen05 (bf16 or fp16 operands, fp32 accumulate in TMEM, fp32 output): wt_s, wt_v are the
 * weights pre-transposed to 16-bit [n_out][K] by segnn_pack_node_weight_tc with the same `operand` format.  Needs
 * n_in % 16 == 0, K <= 192, n_out / split / n_bias multiples of 32, 32-byte aligned tensors. */
int segnn_node_gemm_tc(const float* x0, const float* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                       const float* bias, int n_bias, int n_out, float* y0, float* y1, int split, int operand,
                       segnn_stream_t stream);

/* Same GEMM with 16-bit output, nodes interleaved in pairs: y [nodes / 2][4][cols][2] fp16 (element (node, plane, col)
 * at ((node / 2) * 4 + plane) * cols * 2 + col * 2 + (node & 1)); cols = split for y0, n_out - split for y1.  This is
 * the layout SEGNN_MODE_FP16_PACKED reads: one 32-bit word holds the same projection of two consecutive senders.
 * `nodes` must be even (graphs of even size, so pairs never straddle two graphs). */
int segnn_node_gemm_tc_pair16(const float* x0, const float* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                              const float* bias, int n_bias, int n_out, void* y0, void* y1, int split, int operand,
                              segnn_stream_t stream);

/* Same GEMM with plain fp16 rows y [nodes][4][n_out] (no split, no bias): for outputs that only feed an
 * attribute-combine pass (segnn_tp_combine_y16), which halves the bytes of both kernels. */
int segnn_node_gemm_tc_out16(const float* x0, const float* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                             int n_out, void* y, int operand, segnn_stream_t stream);

/* Same GEMM reading 16-bit rows: x0, x1 planar [nodes][4][n_in] already in the `operand` format (fp16 copies written
 * by segnn_embed_fwd_x16 / segnn_tp_combine_y16_x16 / segnn_edge_layer_fwd_out16), copied to the swizzled A tile
 * without conversion, two batches of loads in flight per loader thread across tile boundaries.  out_mode 1: the
 * pair-interleaved fp16 output of segnn_node_gemm_tc_pair16 (y0 / y1 / split / bias as there); out_mode 2: the fp16
 * rows of segnn_node_gemm_tc_out16 (y1 NULL, split = n_out, bias NULL).  Results are bit-identical to the fp32-input
 * entry points on inputs whose fp16 copy was rounded to nearest even (o3_building_blocks.py:150-162 over node rows). */
int segnn_node_gemm_tc_x16(const void* x0, const void* x1, int nodes, int n_in, const void* wt_s, const void* wt_v,
                           const float* bias, int n_bias, int n_out, void* y0, void* y1, int split, int operand,
                           int out_mode, segnn_stream_t stream);

/* w [K][n_out] fp32 -> wt [n_out][K] bf16 / fp16 (the B operand image of segnn_node_gemm_tc). */
int segnn_pack_node_weight_tc(const float* w, int K, int n_out, int operand, void* wt, segnn_stream_t stream);

/* Attribute coupling + epilogue of a node-level tensor product.  With a = node_attr[node] = (a0, a1[3]):
 *   z0[w]    = a0 * y[0][w]      + sum_k a1[k] * y[1+k][w] + bias[w]   w < n0 (l=0 outputs; bias may be NULL)
 *   z1[w][k] = a1[k] * y[0][n0+w] + a0 * y[1+k][n0+w]             w < n    (l=1 outputs)
 * gate != 0 (O3TensorProductSwishGate, o3_building_blocks.py:197-203 + e3nn Gate): n0 = 2n,
 *   out = (c_silu*silu(z0[w]), c_sig*sigmoid(z0[n+w]) * z1[w][k]);
 * gate == 0 (O3TensorProduct): n0 = n, out = (z0, z1).
 * residual (may be NULL) planar [nodes][4][n] is added (models/segnn/segnn.py:303 x += update).
 * bn_mul [2n] / bn_add [n] (may be NULL): eval-mode e3nn BatchNorm folded to out_s*mul[w]+add[w],
 * out_v*mul[n+w] (segnn.py:257-261).  out planar [nodes][4][n]. */
int segnn_tp_combine(const float* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                     const float* residual, const float* bn_mul, const float* bn_add, float* out,
                     segnn_stream_t stream);

/* Same pass reading y as fp16 rows [nodes][4][n0+n] (written by segnn_node_gemm_tc_out16). */
int segnn_tp_combine_y16(const void* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                         const float* residual, const float* bn_mul, const float* bn_add, float* out,
                         segnn_stream_t stream);

/* Same pass, also (or only: out may be NULL) writing out16, the fp16 copy [nodes][4][n] of the result that the next
 * tensor-core node GEMM reads through segnn_node_gemm_tc_x16 (update_layer_1 -> update_layer_2, segnn.py:286-304; the
 * layer output -> the next layer's message_layer_1 / update_layer_1).  Needs an even n. */
int segnn_tp_combine_y16_x16(const void* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                             const float* residual, const float* bn_mul, const float* bn_add, float* out,
                             void* out16, segnn_stream_t stream);

/* ---- K3: fused edge layer ---------------------------------------------------------------------------- */

/* SEGNNLayer.message + aggregate (models/segnn/segnn.py:264-284 + PyG scatter-add, segnn.py:205) over the
 * implicit fully-connected graph, eval-mode message BatchNorm folded into the aggregate:
 *   agg[i] = bn_mul * sum_{j != i} gate(msg2(gate(msg1(x_i, x_j, |r_ij|, m_i m_j; Y(r_ij))); Y(r_ij))) + bn_add
 * message_layer_1 is linear in (x_i, x_j), so its weight contraction is hoisted to the node GEMM:
 *   p, q [nodes][4][3n] = per plane (X0[2n], X1[n]) -- receiver (P) and sender (Q) projections with the
 *   constant Y_0 and the bias folded in (see pack_msg1 in the host package; the node GEMM writes both with
 *   split = 3n).  Column order inside each plane row: SEGNN_MODE_FP32 reads (0s [n] | 0g [n] | 1o [n]); the
 *   tensor-core modes read (0s [n] | (0g, 1o) pairs [n][2]) so that one 64-bit load fetches both parts of a channel
 *   (the host package permutes the columns of the node-GEMM weight image accordingly).
 * w_edge1 [6n] = (d->0e [2n], mm->0e [2n], d->1o [n], mm->1o [n]) (Y_0 folded into the 0e parts).
 * msg2 weights (fp32 mode): w2_ss [n][2n] (Y_0 folded), w2_vs [n][2n] (1/sqrt3 folded), w2_sv [n][n],
 *   w2_vv [n][n] (Y_0 folded), b2 [2n].
 * bn_mul [2n], bn_add [n]: folded eval BatchNorm with the degree (N-1) applied to the mean/bias terms;
 *   both NULL => raw sum.
 * agg_out planar [nodes][4][n].
 * moments (may be NULL): [nodes][2n] per-receiver sums needed for train-mode batch statistics:
 *   (sum_j m_s[w]^2 , sum_j |m_v[w]|^2) of the raw messages; the plain sums are agg itself when bn_mul == NULL.
 *   Emitted by SEGNN_MODE_FP32 and SEGNN_MODE_FP16_PACKED (the reference evaluates rollouts with train-mode BatchNorm,
 *   trainer.py:373,929-942); the other tensor-core modes refuse a non-NULL pointer. */
int segnn_edge_layer_fwd(int mode, const float* pos, const float* mass, int B, int N, int n, const float* p,
                         const float* q, const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                         const float* w2_vv, const float* b2, const void* w2_tc, const float* bn_mul,
                         const float* bn_add, float* agg_out, float* moments, segnn_stream_t stream);

/* SEGNN_MODE_FP16_PACKED with the aggregate written as fp16 rows agg16_out [nodes][4][n] (round to nearest even of the
 * same fp32 values, eval-mode BatchNorm folded as above): the x1 operand of update_layer_1's tensor-core GEMM
 * (segnn_node_gemm_tc_x16; segnn.py:205 aggregation -> segnn.py:286-299 update).  p, q, w2_tc as in that mode. */
int segnn_edge_layer_fwd_out16(const float* pos, const float* mass, int B, int N, int n, const void* p, const void* q,
                               const float* w_edge1, const float* b2, const void* w2_tc, const float* bn_mul,
                               const float* bn_add, void* agg16_out, segnn_stream_t stream);

/* Packs the message_layer_2 weights for SEGNN_MODE_BF16_TC / SEGNN_MODE_FP16_TC (operand = SEGNN_OPERAND_*) into the
 * [128 lanes][3n] image of 16-bit pairs the tensor-core kernel copies verbatim into TMEM (gate constants folded).
 * Returns the image size in bytes (also when out == NULL). */
int64_t segnn_pack_w2_tc(const float* w2_ss, const float* w2_vs, const float* w2_sv, const float* w2_vv, int n,
                         int operand, void* out, segnn_stream_t stream);

/* ---- K5/K6: head + self-feed integration ------------------------------------------------------------- */

/* models/segnn/segnn.py:180 pre_pool2 = O3TensorProduct(h -> 2x1o, node_attr), no bias.
 * w_head [2][n][2]: (W_s->1o [n][2], W_v->1o [n][2]).  pred [nodes][6] = (out_0 xyz, out_1 xyz). */
int segnn_head_fwd(const float* h, const float* node_attr, const float* w_head, int nodes, int n, float* pred,
                   segnn_stream_t stream);

/* helper_scripts/infer_self_feed.py:182-194 with target 'pos_dt+vel': pos += pred[:, :3]; vel = pred[:, 3:].
 * traj_pos / traj_vel (may be NULL): trajectory buffers [frames][nodes][3]; the new state is also written to
 * frame slot *frame (device int; NULL => slot 0), which replaces predicted_loc.append(...) (:188-189) and keeps
 * the step replayable as a CUDA graph.  The trajectory write is skipped when *frame >= max_frames (a replayed graph
 * can never write past the buffers); max_frames <= 0 disables the check. */
int segnn_integrate(const float* pred, float* pos, float* vel, int nodes, float* traj_pos, float* traj_vel,
                    const int* frame, int max_frames, segnn_stream_t stream);

/* *counter += delta on the device (the rollout's frame cursor; `for step in range(...)`, infer_self_feed.py:99). */
int segnn_counter_add(int* counter, int delta, segnn_stream_t stream);


/* ---- training: train-mode BatchNorm statistics and the backward pass (fp32) ---------------------------- */

/* Deterministic column reduction out[c] = sum_r f(x[r][c], y[r][c]); mode 0: x, 1: x*x, 2: x*y.  Used for e3nn
 * BatchNorm batch statistics (models/segnn/segnn.py:233-235) and for bias / BatchNorm-parameter gradients.
 * workspace: segnn_colsum_workspace(rows, cols) bytes. */
int64_t segnn_colsum_workspace(int64_t rows, int cols);
int segnn_colsum(const float* x, const float* y, int64_t rows, int cols, int mode, float* workspace, float* out,
                 segnn_stream_t stream);
/* Two column sums in one call: (xa, ya, mode_a) -> out_a [cols_a] and (xb, yb, mode_b) -> out_b [cols_b]; up to 2048
 * rows each they are ONE launch (train-mode BatchNorm statistics come in pairs: sum / sum of squares in the forward,
 * sum g / sum g*x in the backward; e3nn BatchNorm, segnn.py:233-235), otherwise two segnn_colsum calls that share
 * `workspace` (max of the two segnn_colsum_workspace sizes).  Same sums bit for bit as segnn_colsum. */
int segnn_colsum2(const float* xa, const float* ya, int64_t rows_a, int cols_a, int mode_a, float* out_a,
                  const float* xb, const float* yb, int64_t rows_b, int cols_b, int mode_b, float* out_b,
                  float* workspace, segnn_stream_t stream);

/* out[r][c] = A[c]*dy[r][c] + B[c]*x[r][c] + C[c] (x/B and C may be NULL): train-mode BatchNorm forward
 * (dy := pre-norm features) and backward with the batch-statistics terms folded into per-column coefficients. */
int segnn_lincomb(const float* dy, const float* x, const float* A, const float* B, const float* C, int64_t rows,
                  int cols, float* out, segnn_stream_t stream);

/* e3nn BatchNorm (models/segnn/segnn.py:233-235) for hidden irreps n x0e + n x1o from column sums over `rows` rows
 * (rows = E for the message norm with deg = N-1 messages per receiver folded through the sender sum, rows = nodes and
 * deg = 1 for the feature norm).  sums [>= n]: sum of the scalar channels; sq [n + v_planes*n]: sums of squares of the
 * scalar channels, then of the vector planes (v_planes = 1 when already summed over xyz).  weight [2n], bias [n],
 * running_mean [n], running_var [2n] (updated in place with `momentum` when training && update, e3nn semantics).
 * Outputs: cols [2][4n] = per planar column (mul, add) of the folded affine; stats [5][n] = (mean, var_s, var_v,
 * rsqrt(var_s + eps), rsqrt(var_v + eps)) kept for the backward pass.  training == 0 uses the running statistics. */
int segnn_bn_coeffs_fwd(const float* sums, const float* sq, int v_planes, int n, double rows, double deg,
                        const float* weight, const float* bias, float* running_mean, float* running_var, double eps,
                        double momentum, int training, int update, float* cols, float* stats, segnn_stream_t stream);

/* Backward of the same: sum_g [4n] = sum over nodes of dL/d(output) per planar column, sum_gx [4n] = sum of
 * dL/d(output) * (raw aggregate or pre-norm feature).  cols [3][4n] = (A, B, C) with dL/dx = A*g + B*x + C per planar
 * column; edge [5n] = (A_s, A_v, B_s, B_v, C_s) for segnn_edge_layer_bwd; dparam [3n] = (dweight [2n], dbias [n]). */
int segnn_bn_coeffs_bwd(const float* sum_g, const float* sum_gx, int n, double rows, double deg, const float* weight,
                        const float* stats, int training, float* cols, float* edge, float* dparam,
                        segnn_stream_t stream);

/* out = a + b + c elementwise (c may be NULL): sums the gradient streams meeting at a layer input (residual,
 * update_layer_1 and message_layer_1 branches of models/segnn/segnn.py:249-304). */
int segnn_add3(const float* a, const float* b, const float* c, int64_t count, float* out, segnn_stream_t stream);

/* Backward of segnn_tp_combine without residual / BatchNorm: dout [nodes][4][n] -> dy [nodes][4][n0+n] and
 * dz0 [nodes][n0] (rows of the bias gradient; may be NULL). */
int segnn_tp_combine_bwd(const float* y, const float* node_attr, int nodes, int n, int gate, const float* bias,
                         const float* dout, float* dy, float* dz0, segnn_stream_t stream);

/* Weight gradient of the node GEMM: dw_s[k][c] = sum over scalar-plane rows x*dy, dw_v likewise over the three
 * vector planes; x = cat_K(x0, x1), dy = cat_cols(dy0 [split], dy1).  Deterministic (slab partials + reduction).
 * (The input gradient is segnn_node_gemm itself with transposed weights.) */
int64_t segnn_node_gemm_wgrad_workspace(int nodes, int K, int n_out);
int segnn_node_gemm_wgrad(const float* x0, const float* x1, const float* dy0, const float* dy1, int split, int nodes,
                          int n_in, int n_out, float* workspace, float* dw_s, float* dw_v, segnn_stream_t stream);

/* Backward of the fused edge layer (autograd of models/segnn/segnn.py:264-284 + the scatter-add): messages are
 * recomputed, nothing per-edge is stored.  pass 0: dout = dP [nodes][4][3n], message_layer_2 weight/bias gradients
 * (written to dw2_*, db2 by a fixed-order reduction of per-thread-group slabs in `workspace`) and per-receiver w_edge1 gradient rows dwe_partial [nodes][6n];
 * pass 1: dout = dQ; pass 2: dP and dwe_partial only, pass 3: the weight/bias gradients only (dout may be NULL) --
 * training-size graphs run 2 on the main stream and 3 beside pass 1 on side streams, large graphs run 0 (one recompute).
 * The gradient reaching every message of receiver i is bn_a*dagg_i + bn_b*m + bn_c
 * (bn_a, bn_b [2n], bn_c [n]): plain sum, eval BatchNorm or train-mode BatchNorm.  w2t_* are the transposes
 * ([out][in]) of the w2_* blocks. */
int segnn_edge_layer_bwd(int pass, const float* pos, const float* mass, int B, int N, int n, const float* p,
                         const float* q, const float* w_edge1, const float* w2_ss, const float* w2_vs,
                         const float* w2_sv, const float* w2_vv, const float* b2, const float* w2t_ss,
                         const float* w2t_vs, const float* w2t_sv, const float* w2t_vv, const float* bn_a,
                         const float* bn_b, const float* bn_c, const float* dagg, float* dout, float* dw2_ss,
                         float* dw2_vs, float* dw2_sv, float* dw2_vv, float* db2, float* dwe_partial,
                         float* workspace, segnn_stream_t stream);

/* Bytes of `workspace` for pass 0 of segnn_edge_layer_bwd: one private slab of partial message_layer_2 weight
 * gradients per thread group of the persistent grid, reduced in a fixed order (no atomics: the gradients are
 * bit-identical from run to run).  The workspace needs no initialisation; pass 1 ignores it. */
int64_t segnn_edge_layer_bwd_workspace(int B, int N, int n);

/* Backward of the embedding (inputs carry no gradient): per-node contribution rows [nodes][7][n] to
 * (w_embed [6][n], bias [n]); reduce with segnn_colsum. */
int segnn_embed_bwd(const float* x_in, const float* node_attr, const float* dh, int nodes, int n, float* contrib,
                    segnn_stream_t stream);

/* Backward of the head: dh [nodes][4][n] and per-node contribution rows [nodes][4][n] to w_head. */
int segnn_head_bwd(const float* h, const float* node_attr, const float* w_head, const float* dpred, int nodes, int n,
                   float* dh, float* contrib, segnn_stream_t stream);

/* ---- generic-irreps path (fp32 inference for hidden irreps the fused kernels are not specialised for, e.g. lmax_h = 2) */

/* O3TensorProduct.forward_tp_rescale_bias (models/segnn/o3_building_blocks.py:150-162) for arbitrary irreps:
 * out[row] = FullyConnectedTensorProduct(x1[row], x2[row]) with e3nn path normalisation x the reference's
 * sqrt_k_correction, + bias.  instr [n_instr][9] = (offset, mul, 2l+1 of in1; offset, 2l+1 of in2; offset, mul, 2l+1
 * of out; flat weight offset) in e3nn instruction order; cg [n_instr][5][3][5] = real Wigner 3j scaled by the net path
 * coefficient; bias: dense [dout] (zeros on l > 0 columns) or NULL. */
int segnn_generic_tp(const float* x1, int d1, const float* x2, int d2, int64_t rows, const float* weights,
                     const int* instr, int n_instr, const float* cg, const float* bias, int dout, float* out,
                     segnn_stream_t stream);

/* "Expand + GEMM" form of segnn_generic_tp for large row counts, one output irrep block (mulo x dimo) at a time:
 * segnn_generic_tp_expand writes A [rows * dimo][K] with the coupling to x2 applied
 *   A[row * dimo + k][koff_p + u] = sum_{i,j} C_p[i][j][k] x1[row][off1_p + u * dim1_p + i] x2[row][off2_p + j]
 * for the paths p of the block (paths [n_paths][6] = off1, mul1, dim1, off2, dim2, koff; cg [n_paths][75]); the caller
 * multiplies A by the stacked path weights [K][mulo] with a library SGEMM; segnn_generic_tp_scatter stores
 * Y [rows * dimo][mulo] into out[row][offo + w * dimo + k] (+ bias [dout], may be NULL). */
int segnn_generic_tp_expand(const float* x1, int d1, const float* x2, int d2, int64_t rows, const int* paths,
                            int n_paths, const float* cg, int dimo, int K, float* A, segnn_stream_t stream);
int segnn_generic_tp_scatter(const float* Y, int64_t rows, int dimo, int mulo, int offo, int dout, const float* bias,
                             float* out, segnn_stream_t stream);
/* The same two with explicit leading dimensions (lda >= K floats per row of A, ldy >= mulo floats per row of Y): rows
 * padded to a multiple of four floats let segnn_gemm_tf32x3 use 128-bit accesses. */
int segnn_generic_tp_expand_ld(const float* x1, int d1, const float* x2, int d2, int64_t rows, const int* paths,
                               int n_paths, const float* cg, int dimo, int K, int64_t lda, float* A,
                               segnn_stream_t stream);
int segnn_generic_tp_scatter_ld(const float* Y, int64_t ldy, int64_t rows, int dimo, int mulo, int offo, int dout,
                                const float* bias, float* out, segnn_stream_t stream);

/* fp32-accurate GEMM on tcgen05 ("3xTF32"): C[M][N] = A[M][K] * B[K][N], row-major with leading dimensions lda, ldb,
 * ldc (floats).  Every operand value is split into tf32 hi + lo parts and hi*hi + hi*lo + lo*hi is accumulated in fp32
 * in TMEM (kind::tf32 MMAs), <= 2e-7 relative to a float64 product.  It is the weight contraction of the generic-irreps
 * tensor products (o3_building_blocks.py:150-162 for arbitrary irreps, the lmax_h = 2 configuration), which round 1 ran
 * as a library SGEMM.  workspace: segnn_gemm_tf32x3_workspace(K, N) bytes, 16-byte aligned (the swizzled hi / lo images
 * of B, rebuilt by every call). */
int64_t segnn_gemm_tf32x3_workspace(int K, int N);
int segnn_gemm_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int K, int N, float* C,
                      int64_t ldc, float* workspace, segnn_stream_t stream);

/* segnn_node_gemm (same arguments and results: the weight contraction of O3TensorProduct.forward_tp_rescale_bias on
 * node rows, o3_building_blocks.py:150-162) as a 3xTF32 GEMM on tcgen05: both row classes (scalar planes x W_s, vector
 * planes x W_v) in one launch, K optionally the concatenation x0 | x1, bias on the scalar planes, split outputs.  It is
 * the fp32-mode / training node GEMM (the FFMA kernel stays as the checker).  workspace:
 * segnn_node_gemm_tf32x3_workspace(K, n_out) bytes (K = n_in or 2 n_in), 16-byte aligned. */
int64_t segnn_node_gemm_tf32x3_workspace(int K, int n_out);
int segnn_node_gemm_tf32x3(const float* x0, const float* x1, int nodes, int n_in, const float* w_s, const float* w_v,
                           const float* bias, int n_bias, int n_out, float* y0, float* y1, int split, float* workspace,
                           segnn_stream_t stream);

/* The same accuracy for C[M][N] (+)= sum_k A[k][m] * B[k][n]: both operands row-major over K ("TN"), K split over the
 * CTAs of the grid with the partial sums in TMEM and one fixed-order reduction (bit-identical from run to run, no
 * atomics).  It is the weight-gradient contraction of message_layer_2 over the edge rows (autograd of
 * o3_building_blocks.py:150-162 inside models/segnn/segnn.py:264-284).  M, N, lda, ldb multiples of 4 floats, N <= 512;
 * accumulate != 0 adds to C.  workspace: segnn_gemm_tn_tf32x3_workspace(K, M, N) bytes, 16-byte aligned. */
int64_t segnn_gemm_tn_tf32x3_workspace(int64_t K, int M, int N);
int segnn_gemm_tn_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t K, int M, int N, float* C,
                         int64_t ldc, int accumulate, float* workspace, segnn_stream_t stream);
/* The same with `groups` blocks side by side in every row: C[M][N] (+)= sum_g A[:, g M : (g + 1) M]^T B[:, g N : (g + 1) N]
 * (lda >= groups * M, ldb >= groups * N; groups > 1: M, N multiples of 32, M <= 128, N <= 256).  The W_vv weight
 * gradient sums over the three vector components, which sit side by side in a row of the edge tensors: one pass over
 * rows of 3n floats instead of three times as many rows of n.  Same workspace as segnn_gemm_tn_tf32x3. */
int segnn_gemm_tn_grouped_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t K, int M, int N,
                                 int groups, float* C, int64_t ldc, int accumulate, float* workspace,
                                 segnn_stream_t stream);

/* Large-graph form of segnn_edge_layer_fwd (SEGNN_MODE_FP32 semantics, same arguments and outputs) and of
 * segnn_edge_layer_bwd (passes 0 and 1 in one call): SEGNNLayer.message + the scatter-add (models/segnn/segnn.py:205,
 * 264-284) and their autograd over graphs with many nodes (BASELINE configuration 4: N = 1000, ~1 M edges).  The
 * message_layer_2 contraction and its data / weight gradients run as 3xTF32 GEMMs on tcgen05 over the edge rows of one
 * chunk of graphs at a time (rows = graphs * N * N, 11 n floats per row forward, 16 n backward, in `workspace`);
 * gradients are bit-identical from run to run.  n must be a multiple of 4, <= 96.  workspace: 256-byte aligned,
 * segnn_edge_layer_gemm_workspace(B, N, n, backward, budget_bytes) bytes = the fixed part plus as many whole graphs per
 * chunk as fit into budget_bytes (at least one; budget_bytes <= 0: all B graphs in one chunk); the calls derive the
 * chunking from workspace_bytes.  fwd_workspace (may be NULL): the untouched workspace of the forward call of the same
 * layer, when that call ran as ONE chunk -- the backward call then reads the rows the forward left there instead of
 * recomputing them (11 n floats per edge row kept alive between the two calls; it does not modify them). */
int64_t segnn_edge_layer_gemm_workspace(int B, int N, int n, int backward, int64_t budget_bytes);
int segnn_edge_layer_gemm_fwd(const float* pos, const float* mass, int B, int N, int n, const float* p, const float* q,
                              const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                              const float* w2_vv, const float* b2, const float* bn_mul, const float* bn_add,
                              float* agg_out, float* moments, float* workspace, int64_t workspace_bytes,
                              segnn_stream_t stream);
int segnn_edge_layer_gemm_bwd(const float* pos, const float* mass, int B, int N, int n, const float* p, const float* q,
                              const float* w_edge1, const float* w2_ss, const float* w2_vs, const float* w2_sv,
                              const float* w2_vv, const float* b2, const float* w2t_ss, const float* w2t_vs,
                              const float* w2t_sv, const float* w2t_vv, const float* bn_a, const float* bn_b,
                              const float* bn_c, const float* dagg, float* dP, float* dQ, float* dw2_ss, float* dw2_vs,
                              float* dw2_sv, float* dw2_vv, float* db2, float* dwe_partial, float* workspace,
                              int64_t workspace_bytes, const float* fwd_workspace, int64_t fwd_workspace_bytes,
                              segnn_stream_t stream);
/* The same call in separately launchable phases (bit mask; 7 = segnn_edge_layer_gemm_bwd): bit 0 = rows (recompute
 * unless fwd_workspace holds them) + gate backward + bias-gradient sums, bit 1 = the message_layer_2 weight gradients
 * (split-K TN GEMMs over the edge rows, dw2_* / db2), bit 2 = the data gradients down to dP / dQ / dwe_partial.  Phases
 * 1 and 2 only read what phase 0 left in `workspace`, and write disjoint parts of it, so a caller may launch the weight
 * gradients on a second stream beside the data-gradient chain (trainer.py:309 loss.backward(): the weight gradients of
 * a layer are off the critical path of the layers below it).  Needs a workspace that holds every graph in ONE chunk
 * when phases != 7.  Same arguments, same results bit for bit. */
int segnn_edge_layer_gemm_bwd_phases(const float* pos, const float* mass, int B, int N, int n, const float* p,
                                     const float* q, const float* w_edge1, const float* w2_ss, const float* w2_vs,
                                     const float* w2_sv, const float* w2_vv, const float* b2, const float* w2t_ss,
                                     const float* w2t_vs, const float* w2t_sv, const float* w2t_vv, const float* bn_a,
                                     const float* bn_b, const float* bn_c, const float* dagg, float* dP, float* dQ,
                                     float* dw2_ss, float* dw2_vs, float* dw2_sv, float* dw2_vv, float* db2,
                                     float* dwe_partial, float* workspace, int64_t workspace_bytes,
                                     const float* fwd_workspace, int64_t fwd_workspace_bytes, int phases,
                                     segnn_stream_t stream);

/* lmax_h = 2 (hidden irreps n x 0e + n x 1o + n x 2e, attribute 0e + 1o, two additional scalars: BASELINE
 * configuration 3) edge layer in GEMM form, inference: the per-edge work of SEGNNLayer.message + aggregation
 * (models/segnn/segnn.py:205,264-284; o3_building_blocks.py:150-203) around the weight contraction of message_layer_2.
 * Edge rows are (graph, receiver, sender) of `graphs` graphs starting at node `node0`, diagonal kept and masked.
 * Instruction types t = 0..6 = (l1, l2, lo) in the order (0,0,0) (0,1,1) (1,0,1) (1,1,0) (1,1,2) (2,0,2) (2,1,1).
 * segnn_l2_planarize: hidden features x [nodes][9n] (e3nn layout) -> x_l [(node (2l + 1) + i)][ld], channel contiguous:
 *   the A operands of the node-level products of message_layer_1, Y_l = x_l * (stacked weights of the instructions
 *   with input degree l, both roles), computed with segnn_gemm_tf32x3.
 * segnn_l2_msg_rows: Y_l [(node (2l + 1) + i)][ldy_l] with instruction (t, role) in columns yoff[t][role] .. + its
 *   output multiplicity (role 0 = x_i, 1 = x_j), cg [7][5][3][5] net couplings, w_add0 [2][3n] / w_add1 [2][n] weights
 *   of the additional scalars, bias1 [3n]; writes the three operands of message_layer_2's contraction, A0 [rows][lda0],
 *   A1 [3 rows][lda1], A2 [5 rows][lda2], type t in columns koff[t] .. koff[t] + n.
 * segnn_l2_gate_aggregate: Y_b = A_b * stacked weights (segnn_gemm_tf32x3) -> bias2 [3n], e3nn Gate, sum over senders,
 *   folded eval BatchNorm (bn_mul, bn_add [9n] per hidden column, add applied N - 1 times; both NULL = none) ->
 *   agg [nodes][9n] in e3nn layout. */
int segnn_l2_planarize(const float* x, int64_t nodes, int n, int64_t ld, float* x0, float* x1, float* x2,
                       segnn_stream_t stream);
int segnn_l2_msg_rows(const float* pos, const float* mass, int graphs, int N, int n, int64_t node0, const float* Y0,
                      int64_t ldy0, const float* Y1, int64_t ldy1, const float* Y2, int64_t ldy2, const int* yoff,
                      const float* cg, const float* w_add0, const float* w_add1, const float* bias1, const int* koff,
                      int64_t lda0, int64_t lda1, int64_t lda2, float* A0, float* A1, float* A2,
                      segnn_stream_t stream);
int segnn_l2_gate_aggregate(int graphs, int N, int n, int64_t node0, const float* Y0, int64_t ld0, const float* Y1,
                            int64_t ld1, const float* Y2, int64_t ld2, const float* bias2, const float* bn_mul,
                            const float* bn_add, float* agg, segnn_stream_t stream);

/* message_layer_1 (models/segnn/segnn.py:264-279) with its weight contraction hoisted to node level, for any hidden
 * irreps: Y [nodes][ydim] holds, for every (x_i or x_j) instruction, sum_u W[u][w] x[node][u, i] at yoff + w * dim1 + i
 * (computed with segnn_generic_tp and an identity coupling); this kernel applies the coupling with the edge attribute
 * per edge of the reference enumeration, adds the `additional_message_features` instructions (weights = the flat
 * tp.weight) and the bias.  pairs [n_pairs][8] = offo, mulo, dimo, dim1, off2, dim2, yoff_i, yoff_j; adds [n_adds][7] =
 * offo, mulo, dimo, off2, dim2, woff, mul1; blocks [n_blocks][3] = offo, mulo, dimo of every output irrep block,
 * n_items = sum of their mulo; cg [n_pairs + n_adds][75]. */
int segnn_generic_hoisted_msg1(const float* Y, int ydim, const float* attr, int d2, const float* add, int d_add, int B,
                               int N, const int* pairs, int n_pairs, const int* adds, int n_adds, const int* blocks,
                               int n_blocks, int n_items, const float* cg, const float* weights, const float* bias,
                               int dout, float* out, segnn_stream_t stream);

/* e3nn Gate as used by O3TensorProductSwishGate (:186-203): x [rows][n_scalars + n_gates + d_gated] ->
 * out [rows][n_scalars + d_gated]; gate_index [d_gated] = gate of every gated column. */
int segnn_generic_gate(const float* x, int64_t rows, int n_scalars, int n_gates, int d_gated, const int* gate_index,
                       float* out, segnn_stream_t stream);

/* cat(x_i, x_j, additional_message_features) in the reference edge order (models/segnn/segnn.py:264-277):
 * x [nodes][D], add [E][d_add] -> out [E][2D + d_add]. */
int segnn_generic_message_input(const float* x, const float* add, int B, int N, int D, int d_add, float* out,
                                segnn_stream_t stream);

/* Deterministic scatter-sum over the implicit complete graph (segnn.py:205): m [E][D] -> agg [nodes][D]. */
int segnn_generic_aggregate(const float* m, int B, int N, int D, float* agg, segnn_stream_t stream);

/* ---- rollout macros (evaluation of rollouts) ------------------------------------------- */

/* trainer.py:888-927 `_compute_nbody_energies` and datasets/nbody/visualization_utils.py:959-960 (momentum), per
 * (frame, simulation) instead of per Python loop iteration: traj_pos, traj_vel [frames][B*N][3] (the rollout's
 * trajectory buffers) -> out [frames][B][3] = (kinetic = 0.5 sum v^2, potential = -G sum_{i<j} (r_ij^2 +
 * softening^2)^-1/2, |sum_i v_i|), unit masses as in the reference. */
int segnn_macros_energy_momentum(const float* traj_pos, const float* traj_vel, int frames, int B, int N, float G,
                                 float softening, float* out, segnn_stream_t stream);

/* datasets/nbody/visualization_utils.py:1093-1124 (count_stickings_and_collisions), :1145-1167
 * (count_balls_leaving_defined_area), :1170-1187 (get_max_distance_of_com_from_starting_position), :1201-1222
 * (count_sharp_turns) on the device trajectory buffers [frames][B*N][3]: out_counts [B][4] int32 = (stickings,
 * collisions, bodies_left, sharp_turns), out_com [B] = max distance of the centre of mass from its start.
 * Reference defaults: time_threshold 3, contact_distance 0.5, leave_distance 15, turn_angle_degrees 30. */
int segnn_macros_counters(const float* traj_pos, const float* traj_vel, int frames, int B, int N, int time_threshold,
                          float contact_distance, float leave_distance, float turn_angle_degrees, int* out_counts,
                          float* out_com, segnn_stream_t stream);

/* ---- ground-truth simulator ------------------------------------------------------------- */

/* GravitySim.sample_trajectory (datasets/nbody/dataset/synthetic_sim.py:305-420; the OTF dataset's generator,
 * datasets/nbody/dataset_gravity_otf.py:91-107): softened gravity, kick-drift-kick leapfrog in float64, one CTA per
 * simulation, the whole trajectory in one launch.  pos, vel [B*N][3] (in: initial state, out: final state), mass [B*N];
 * traj_* [steps / sample_freq][B*N][3]: frame k = state after k * sample_freq steps (frame 0 = initial state);
 * traj_force may be NULL.  Observation noise (noise_var, 0 by default) is left to the caller. */
int segnn_sim_gravity(double* pos, double* vel, const double* mass, int B, int N, double G, double softening, double dt,
                      int steps, int sample_freq, double* traj_pos, double* traj_vel, double* traj_force,
                      segnn_stream_t stream);

/* Charged-particle system with isolated bodies (datasets/nbody_offline/datagen/system.py:78-123 System.compute_F /
 * simulate_one_step; physical_objects.py:49-57 Isolated.update): F_i = sum_j k q_i q_j (x_i - x_j) / |x_i - x_j|^3,
 * every component clamped to +-max_force (the reference uses 0.1 / dt), then v += F dt, x += v dt, in float64.
 * pos, vel [B*N][3] (in: initial, out: final state), charge [B*N]; traj_* [steps / sample_freq][B*N][3], frame f =
 * state after (f + 1) * sample_freq steps. */
int segnn_sim_charged(double* pos, double* vel, const double* charge, int B, int N, double interaction_strength,
                      double dt, double max_force, int steps, int sample_freq, double* traj_pos, double* traj_vel,
                      segnn_stream_t stream);

/* datasets/nbody/visualization_utils.py:1455-1610, the counting part of plot_group_collision_distribution_multiplot:
 * per simulation the number of (stuck pair interval, stuck interval of a disjoint triplet) combinations with
 * overlapping lifetimes whose two groups touch at some step >= the start of the overlap.  traj_pos
 * [frames][B*N][3]; workspace: segnn_macros_group_collisions_workspace(frames, B, N) bytes; out_counts [B] int32.
 * Reference defaults: time_threshold 2, distance_threshold 2. */
int64_t segnn_macros_group_collisions_workspace(int frames, int B, int N);
int segnn_macros_group_collisions(const float* traj_pos, int frames, int B, int N, int time_threshold,
                                  float distance_threshold, void* workspace, int* out_counts, segnn_stream_t stream);

/* utils/build_fully_connected_graph.py:42-80, the k-nearest-neighbour branch of build_graph_with_knn: loc
 * [B*N][dim] float64 -> edge_index int64 [2, B*N*k]; edge (g N + i) k + r: row 0 = g N + i, row 1 = g N + (r-th
 * nearest other node of i in graph g, ascending distance). */
int segnn_knn_edge_index(const double* loc, int B, int N, int dim, int k, int64_t* edge_index, segnn_stream_t stream);

/* models/segnn/instance_norm.py:53-129 InstanceNorm.forward (reduce='mean', normalization='component'): x, out
 * [rows][dim]; graph_ptr int64 [graphs + 1] = row range of each graph (rows sorted by graph, as in a PyG batch);
 * blocks int32 [n_blocks][6] = (column offset, multiplicity, 2l+1, l, weight offset, bias offset) per irrep block;
 * weight / bias may be NULL (affine=False). */
int segnn_instance_norm(const float* x, const int64_t* graph_ptr, int graphs, int dim, const int* blocks, int n_blocks,
                        const float* weight, const float* bias, float eps, float* out, segnn_stream_t stream);

/* ---- weight packing: reference parameters -> kernel operand blocks ------------------------------------------ */

/* Which O3TensorProduct of the SEGNN a flat `tp.weight` / `biases` pair belongs to (models/segnn/segnn.py:63-65,
 * 104-106, 212-223).  Output layouts (fp32, concatenated in `out`):
 *   MSG1     message_layer_1   w_s [n][6n] | w_v [n][6n] | bias [2n] | w_edge [6n]     (inputs of segnn_node_gemm*
 *                              producing the hoisted projections P | Q, and of segnn_edge_layer_fwd)
 *   MSG2     message_layer_2   ss [n][2n] | vs [n][2n] | sv [n][n] | vv [n][n] | b [2n]  (segnn_edge_layer_fwd / bwd,
 *                              segnn_pack_w2_tc)
 *   UPDATE1  update_layer_1    w_s [2n][3n] | w_v [2n][3n] | bias [2n]
 *   UPDATE2  update_layer_2    w_s [n][2n]  | w_v [n][2n]  | bias [n]
 *   POOL1    pre_pool1         w_s [n][3n]  | w_v [n][3n]  | bias [2n]                 (segnn_node_gemm* + segnn_tp_combine)
 *   EMBED    embedding_layer   w [6][n] | bias [n]                                      (segnn_embed_fwd)
 *   HEAD     pre_pool2         w_head [2][n][2] (no bias: pass NULL)                    (segnn_head_fwd)
 * The constants folded in: Y_0 of the edge attribute, 1/sqrt(3) of the 1o x 1o -> 0e coupling (the net of e3nn's path
 * weights and the reference's sqrt_k_correction, o3_building_blocks.py:150-162). */
enum {
  SEGNN_PACK_MSG1 = 0,
  SEGNN_PACK_MSG2 = 1,
  SEGNN_PACK_UPDATE1 = 2,
  SEGNN_PACK_UPDATE2 = 3,
  SEGNN_PACK_POOL1 = 4,
  SEGNN_PACK_EMBED = 5,
  SEGNN_PACK_HEAD = 6
};
int64_t segnn_pack_weights_size(int kind, int n); /* number of floats of `out`; -1 for an unknown kind */
int segnn_pack_weights(int kind, int n, const float* tp_weight, const float* biases, float* out, segnn_stream_t stream);

/* Eval-mode e3nn BatchNorm (models/segnn/segnn.py:233-235) folded to per-channel (mul [2n], add [n]): out_s = x_s *
 * mul[:n] + add, out_v = x_v * mul[n:].  degree = N - 1 folds the BatchNorm of every message through the sum over the
 * N - 1 senders (segnn_edge_layer_fwd's bn_mul / bn_add); degree = 1 is the feature norm (segnn_tp_combine). */
int segnn_fold_batchnorm(const float* weight, const float* bias, const float* running_mean, const float* running_var,
                         int n, float eps, float degree, float* mul, float* add, segnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGNN_B200_H_ */
