/*
 * b2me.h — C-ABI of libb2me.so: the B200 (sm_100a) back end behind the MinkowskiEngine-compatible
 * Python package and the batched pose pipeline of this repo.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - the caller owns every buffer (inputs, outputs, hash tables, workspaces); the library never
 *     allocates, frees or synchronises; every entry point enqueues on the caller's stream and acts on the
 *     CURRENT device (the only process-wide state is a cache of per-device constants: SM count, a driver
 *     function pointer). One host thread per stream; entry points are not re-entrant on one stream;
 *   - return value 0 = success, negative B2ME_E* otherwise (b2me_strerror gives the text);
 *   - data-dependent counts (number of voxels V, ...) are written to device scalars; the caller reads
 *     them back when it needs a host-side shape.
 *
 * Each entry point cites the reference interface it replaces (file:line in
 * bcsefercik/markerless-robot-camera-calibration, or the third-party call made there).
 */
#ifndef B2ME_H_
#define B2ME_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b2me_stream_t; /* cudaStream_t */

/* error codes */
#define B2ME_OK 0
#define B2ME_EINVAL (-1)   /* bad argument */
#define B2ME_EWORKSPACE (-2) /* workspace / table too small */
#define B2ME_ELAUNCH (-3)  /* CUDA launch error (cudaGetLastError) */
#define B2ME_EUNSUPPORTED (-4) /* shape/dtype combination not built */

/* dtypes */
#define B2ME_F32 0
#define B2ME_BF16 1
#define B2ME_TF32 2 /* fp32 storage holding tf32-representable values (10-bit mantissa, rounded to nearest): the
                       operand / output type of the tf32 tensor-core mode */

/* activations fused in epilogues */
#define B2ME_ACT_NONE 0
#define B2ME_ACT_RELU 1
#define B2ME_ACT_LEAKY 2

int b2me_version(void);
const char* b2me_strerror(int code);

/* ------------------------------------------------------------------------------------------------
 * Coordinate hashing (K1-K3).  A "table" is an open-addressing hash of 16-byte slots
 * {uint64 key, uint32 val, pad}; key packs (batch:10 | x:18 | y:18 | z:18, biased), val = voxel row.
 * Home slot: the eight voxels of a 2x2x2 group (lowest bit of x, y, z) share one 128-byte line of 8 slots (the group
 * is hashed, the three low bits pick the slot), linear probing after that; a table has a power-of-two number of slots,
 * at least 8 (b2me_table_slots returns at least 1024).  The layout is private to the library: tables are only ever
 * read by the entry points below.
 * ---------------------------------------------------------------------------------------------- */

/* slots needed for n keys (power of two >= 2n) and the byte size of such a table */
int64_t b2me_table_slots(int64_t n);
size_t b2me_table_bytes(int64_t n);
/* scratch needed by b2me_quantize_unique / b2me_stride_map for n input rows with C feature channels */
size_t b2me_unique_workspace_bytes(int64_t n, int C);

/*
 * K1 voxelise: replaces ME.TensorField(...).sparse() (app/inference_engine.py:405-415,
 * test_segmentation.py:62-70) and ME.SparseTensor(...)/ME.utils.sparse_quantize coordinate handling
 * (data/alivev2.py:290-296).
 *   coords_f  [N,4] f32 (b,x,y,z), already multiplied by `scale` by the caller; floorf() is applied here;
 *             OR coords_i [N,4] i32 (exactly one of the two non-null).
 *   feats     [N,C] f32 (may be null when C == 0)
 *   mode      0 = first point of each voxel wins (ME RANDOM_SUBSAMPLE / sparse_quantize),
 *             1 = UNWEIGHTED_AVERAGE (exact fixed-point sum)
 *   out_coords[N,4] i32 capacity, rows 0..V-1 valid, FIRST-OCCURRENCE order
 *   out_feats [N,C] f32 capacity
 *   inverse   [N] i32  point -> voxel row
 *   first_idx [N] i32 capacity: voxel row -> index of its first point (may be null)
 *   counts    [2] i32: counts[0] = V, counts[1] = error flag (non-zero: coordinate out of key range)
 *   table     caller-owned, b2me_table_bytes(N); on return maps key -> voxel row (kept for K3)
 */
int b2me_quantize_unique(const float* coords_f, const int32_t* coords_i, int64_t N,
                         const float* feats, int C, int mode,
                         int32_t* out_coords, float* out_feats, int32_t* inverse, int32_t* first_idx,
                         int32_t* counts, void* table, size_t table_bytes,
                         void* ws, size_t ws_bytes, b2me_stream_t stream);

/* per-voxel label for sparse_quantize(labels=..., ignore_label=...): the common label of the member
 * points, ignore_label when they disagree (data/alivev2.py:290-296). labels [N] i32 -> out [V] i32 */
int b2me_quantize_labels(const int32_t* labels, const int32_t* inverse, const int32_t* first_idx,
                         int64_t N, int64_t V, int32_t ignore_label, int32_t* out_labels,
                         b2me_stream_t stream);

/*
 * K2 stride map: output coordinates of a stride-2 MinkowskiConvolution (model/backbone/minkunet.py:60-82):
 * unique(floor(c / ts_out) * ts_out) in first-occurrence order.
 *   in_coords [V_in,4] i32 at tensor stride ts_out/2;  out_coords [V_in,4] capacity
 *   in2out [V_in] i32 parent row; koff [V_in] u8 child offset index (x fastest, 0..7)
 *   counts [2] i32 (V_out, error flag); table_out: b2me_table_bytes(V_in)
 */
int b2me_stride_map(const int32_t* in_coords, int64_t V_in, int ts_out,
                    int32_t* out_coords, int32_t* in2out, uint8_t* koff, int32_t* counts,
                    void* table_out, size_t table_bytes, void* ws, size_t ws_bytes,
                    b2me_stream_t stream);

/* kernel maps of the k=2,s=2 convolution and of its transpose from (in2out, koff):
 *   nbr_down [V_out,8] i32: fine row feeding coarse row o at offset k, or -1   (MinkowskiConvolution k2 s2)
 *   nbr_up   [V_in,8]  i32: coarse row feeding fine row f at offset k, or -1   (MinkowskiConvolutionTranspose,
 *                            model/backbone/minkunet.py:87-109: exactly one valid entry per row) */
int b2me_stride_kernel_maps(const int32_t* in2out, const uint8_t* koff, int64_t V_in, int64_t V_out,
                            int32_t* nbr_down, int32_t* nbr_up, b2me_stream_t stream);

/*
 * K3 kernel map of a k=3, stride-1 convolution at tensor stride ts (every BasicBlock conv and the stem):
 *   nbr [V,27] i32, offset index = (dx+1) + 3(dy+1) + 9(dz+1), entry = row of voxel c + d*ts, or -1.
 *   tile_mask [ceil(V/128)] u32 (may be null): bit k set when any row of the 128-row tile has neighbour k.
 */
int b2me_kernel_map_k3(const int32_t* coords, int64_t V, int ts, const void* table, size_t table_bytes,
                       int32_t* nbr, uint32_t* tile_mask, b2me_stream_t stream);

/* K3 for large maps through 4 x 4 x 4 blocks: the coordinate map at tensor stride 4 ts (two stride-2 levels above
 * this one) lists the occupied blocks and its table maps a block origin to its row.
 *   b2me_block_rows: brows [V_blocks,64] i32 = voxel row in every cell of every block (-1 = empty), from the two parent
 *                    maps in2out1 [V] (ts -> 2 ts) and in2out2 [V_2ts] (2 ts -> 4 ts) of b2me_stride_map
 *   b2me_kernel_map_k3_blocks: same nbr [V,27] as b2me_kernel_map_k3 (1..8 probes of the small block table per voxel
 *                    instead of 26 of the voxel table; results staged in shared memory, contiguous stores), plus
 *                    row_masks [V] u32 (bit k = neighbour k present) and offset_counts [32] u32 (neighbours per offset)
 *   b2me_row_masks:  row_masks / offset_counts of an existing nbr [V,K] table (maps built by b2me_kernel_map_k3) */
int b2me_block_rows(const int32_t* coords, int64_t V, int ts, const int32_t* in2out1, const int32_t* in2out2,
                    int64_t V_blocks, int32_t* brows, b2me_stream_t stream);
int b2me_kernel_map_k3_blocks(const int32_t* coords, int64_t V, int ts, const void* block_table,
                              size_t block_table_bytes, const int32_t* brows, int32_t* nbr, uint32_t* row_masks,
                              uint32_t* offset_counts, b2me_stream_t stream);
int b2me_row_masks(const int32_t* nbr, int64_t V, int K, uint32_t* row_masks, uint32_t* offset_counts,
                   b2me_stream_t stream);
/* K3b from row masks: the keys of b2me_mask_sort_keys and the tile masks of b2me_tc_tile_masks without reading the
 * nbr table again (4 bytes per row instead of 4 K) */
int b2me_mask_sort_keys_rows(const uint32_t* row_masks, const uint32_t* offset_counts, int64_t V, int K, int32_t* keys,
                             b2me_stream_t stream);
int b2me_tile_masks_rows(const uint32_t* row_masks, const int32_t* perm, int64_t V, uint32_t* masks,
                         b2me_stream_t stream);

/*
 * K3b: sort keys for the row permutation the tcgen05 convolution takes (`perm`): key[row] = the row's K-bit
 * neighbour-occupancy mask of `nbr` [V,K], bits ordered rarest offset first. Rows sorted by key make 128-row
 * tiles need few kernel offsets. The caller sorts (any sort: the convolution result does not depend on the
 * order of equal keys) and passes the resulting row order as `perm`. ws: >= 128 bytes.
 */
int b2me_mask_sort_keys(const int32_t* nbr, int64_t V, int K, int32_t* keys, void* ws, size_t ws_bytes,
                        b2me_stream_t stream);
/* Mask keys with a spatial tie-break: key[row] = mask key << 37 | frame << 30 | Morton code of coords[row] / ts
 * (coords [V,4] i32 = b,x,y,z; ts a power of two). Rows with the same neighbour pattern follow a space-filling curve,
 * so the neighbours a tile gathers for its different kernel offsets overlap and are re-read from L2. ws: >= 128 bytes. */
int b2me_mask_sort_keys_morton(const int32_t* nbr, const int32_t* coords, int64_t V, int K, int ts, int64_t* keys,
                               void* ws, size_t ws_bytes, b2me_stream_t stream);
/* Two-level keys (K = 27 maps): the 10 rarest offsets of the map form the most significant part as above, the other
 * offsets are ordered per segment by their frequency among the rows of that segment. Fewer (tile, offset) passes than
 * the one-level keys; same contract otherwise. ws: >= b2me_mask_sort_keys2_ws_bytes(V). */
size_t b2me_mask_sort_keys2_ws_bytes(int64_t V);
int b2me_mask_sort_keys2(const int32_t* nbr, int64_t V, int K, int32_t* keys, void* ws, size_t ws_bytes,
                         b2me_stream_t stream);
/* 64-bit keys with a locality prefix: key[row] = (row / block_rows) << 32 | mask key (block_rows = 0: no prefix).
 * Rows of one block of consecutive first-occurrence rows stay together, so concurrently running tiles gather from
 * one L2-sized window of the input. */
int b2me_mask_sort_keys64(const int32_t* nbr, int64_t V, int K, int block_rows, int64_t* keys, void* ws,
                          size_t ws_bytes, b2me_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse convolution (K4) = gather - GEMM - (no scatter: output-stationary), replaces
 * ME.MinkowskiConvolution / MinkowskiConvolutionTranspose / MinkowskiLinear forward
 * (model/backbone/minkunet.py:126-181, model/robotnet_segmentation.py:43-49) with the following
 * MinkowskiBatchNorm (eval), residual add, ME.cat of two sources and ReLU/LeakyReLU folded in:
 *     out[o, :] = act( (sum_k [in1|in2][nbr[o,k], :] @ W[k]) * scale + shift + residual[o, :] )
 * nbr == NULL means identity map (K must be 1).
 * ---------------------------------------------------------------------------------------------- */

/* fp32-accumulate SIMT path; in/out/residual may be f32 or bf16; W is the module parameter
 * [K, Cin1+Cin2, Cout] f32 */
int b2me_spconv_fwd_simt(const void* in1, int Cin1, const void* in2, int Cin2, int in_dtype,
                         const float* W, const int32_t* nbr, int K, int64_t V_out, int Cout,
                         const float* scale, const float* shift,
                         const void* residual, int res_dtype, int act, float slope,
                         void* out, int out_dtype, b2me_stream_t stream);

/* tcgen05 path: bf16 (B2ME_BF16) or tf32 (B2ME_TF32: fp32 rows holding tf32-representable values) operands, fp32
 * accumulation in TMEM.  Weights must be pre-packed for the same operand type. */
size_t b2me_tc_packed_bytes(int K, int Cin1, int Cin2, int Cout, int op_dtype);
int b2me_tc_supported(int K, int Cin1, int Cin2, int Cout);
int b2me_tc_pack_weights(const float* W, int K, int Cin1, int Cin2, int Cout, int op_dtype, void* packed,
                         b2me_stream_t stream);
/* perm [V_out] i32 (may be null = identity): tile t computes the output rows perm[128 t .. 128 t + 127]; any
 * permutation gives bit-identical results (absent neighbours contribute exact zeros), a mask-sorted one
 * (b2me_mask_sort_keys) lets tiles skip the kernel offsets none of their rows has.
 * tile_masks [ceil(V_out / 256)] u32 (required when nbr is given): bit k set when any of the rows
 * perm[256 t .. 256 t + 255] has neighbour k; computed once per (nbr, perm) by b2me_tc_tile_masks and shared by
 * every convolution on that kernel map. */
int b2me_tc_tile_masks(const int32_t* nbr, const int32_t* perm, int64_t V_out, int K, uint32_t* masks,
                       b2me_stream_t stream);

/* flags of the tcgen05 entry points */
#define B2ME_TC_FLAG_TMA 1        /* operands through the TMA unit (cp.async.bulk.tensor tile::gather4 rows + 2-D weight
                                     boxes) instead of cp.async gathers + cp.async.bulk; same results bit for bit */
#define B2ME_TC_FLAG_NO_ROT128 2  /* 384-column tiles: one 384-column accumulator (round-1 layout) instead of the early
                                     release of the 256-column part + alternating 128-column regions; same results */

#define B2ME_TC_FLAG_PF_BULK 4    /* L2 prefetch of the next offset's gathered rows by one cp.async.bulk.prefetch.L2 per row */
#define B2ME_TC_FLAG_PF_NONE 8    /* no L2 prefetch: the default (kept so that callers can say so explicitly) */
#define B2ME_TC_FLAG_PF_NEAR 16   /* L2 prefetch of the next offset's rows by one prefetch.global.L2 per 128-byte chunk (the
                                     round-1 default; measured no faster than none) */
/* bits 8-10: B-ring depth override (weight stages), 0 = library default (3) */

/* V_in: rows of in1 (and of in2, which lies on the same coordinate map; with B2ME_TC_FLAG_TMA row indices outside
 * [0, V_in) read as zeros). op_dtype: type of in1 / in2 / packed_w (B2ME_BF16 | B2ME_TF32); residual rows are bf16
 * for bf16 operands and f32 for tf32 operands. out_dtype: B2ME_BF16 | B2ME_F32 (bf16 operands), B2ME_TF32 | B2ME_F32
 * (tf32 operands; B2ME_TF32 rounds the stored f32 values to tf32 so that they can feed the next tf32 layer). */
int b2me_spconv_fwd_tc(const void* in1, int Cin1, const void* in2, int Cin2, int64_t V_in, int op_dtype,
                       const void* packed_w, const int32_t* nbr, const int32_t* perm,
                       const uint32_t* tile_masks, int K,
                       int64_t V_out, int Cout, const float* scale, const float* shift,
                       const void* residual, int act, float slope,
                       void* out, int out_dtype, int flags, b2me_stream_t stream);

/* Fused head of RobotNetSegmentation / RobotNetVote (model/robotnet_segmentation.py:43-49,
 * model/robotnet_vote.py:50-56): MinkowskiLinear(Cin, Chid) -> activation -> MinkowskiLinear(Chid, C2) in ONE launch;
 * the [V, Chid] hidden activation stays on chip.
 *     hid = act1((in @ W1) * scale1 + shift1)          (tcgen05, W1 packed with b2me_tc_pack_weights(K = 1))
 *     out_logits[v, :] = hid[v, :] @ W2 + b2           (epilogue FMAs, fp32; bf16 operands: hid is rounded to bf16
 *                                                       first, exactly like the stored tensor of the unfused path)
 *     out_argmax[v] = lowest index of the row maximum (utils/output.py:67-73), may be null
 * W2p [Chid, C2p] f32 with C2p = 4 ceil(C2 / 4), columns >= C2 zero; b2 [C2] or null; C2 <= 16; the output tile of
 * Chid must be a multiple of 64 columns (Chid = 1024 -> 256). */
int b2me_head_fused_tc(const void* in, int Cin, int64_t V, int op_dtype, const void* packed_w1, int Chid,
                       const float* scale1, const float* shift1, int act1, float slope1,
                       const float* W2p, const float* b2, int C2, float* out_logits, uint8_t* out_argmax,
                       int flags, b2me_stream_t stream);

/* stand-alone per-channel affine + residual + activation (BatchNorm/ReLU that could not be folded) */
int b2me_affine_act(const void* in, int in_dtype, int64_t V, int C, const float* scale,
                    const float* shift, const void* residual, int res_dtype, int act, float slope,
                    void* out, int out_dtype, b2me_stream_t stream);

/* dtype conversion f32 <-> bf16 of n elements */
int b2me_convert(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, b2me_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Heads (K5, K6, K8)
 * ---------------------------------------------------------------------------------------------- */

/* small-N linear: out[V,Cout] f32 = in[V,Cin] @ Wt^T + bias, Wt [Cout,Cin] f32 (nn.Linear layout),
 * Cout <= 16; optional per-row argmax (lowest index wins ties, torch semantics of utils/output.py:67-73) */
int b2me_linear_small(const void* in, int in_dtype, int64_t V, int Cin, const float* Wt,
                      const float* bias, int Cout, float* out_logits, uint8_t* out_argmax,
                      b2me_stream_t stream);

/* SparseTensor.slice (app/inference_engine.py:417): out[i,:] = in[inverse[i],:] (f32 or bf16 rows) */
int b2me_gather_rows(const void* in, int dtype, int C, const int32_t* index, int64_t N, void* out,
                     b2me_stream_t stream);
/* per-point labels straight from voxel labels (slice + argmax fused): out[i] = lab[inverse[i]] */
int b2me_gather_labels(const uint8_t* voxel_labels, const int32_t* inverse, int64_t N,
                       uint8_t* out, b2me_stream_t stream);

/* K6: MinkowskiGlobalAvgPooling / MinkowskiGlobalMaxPooling (model/robotnet_encode.py:41,102;
 * model/robotnet.py:43): per batch index reduce over voxel rows. coords [V,4] i32 (col 0 = batch),
 * out [B,C] f32. mode 0 = mean, 1 = max. Deterministic. */
int b2me_global_pool(const void* in, int dtype, const int32_t* coords, int64_t V, int C, int B,
                     int mode, float* out, b2me_stream_t stream);

/* K8a: get_key_point_predictions (utils/output.py:81-87), batched over segments.
 * logits [n,K] f32, seg_offsets [S+1] i32 rows of each segment; for every (segment, class):
 * best_prob = max_i softmax(logits[i])[class], best_idx = lowest such i (global row index). */
int b2me_keypoint_reduce(const float* logits, int K, const int32_t* seg_offsets, int S,
                         float* best_prob, int32_t* best_idx, b2me_stream_t stream);

/* K8b: get_pred_center (utils/output.py:45-64): mean coordinate of the `topk` rows with the largest
 * logits[:,col] in every segment (ties: lowest index first). out_center [S,3] f32 */
int b2me_vote_center(const float* logits, int K, int col, const float* points_xyz,
                     const int32_t* seg_offsets, int S, int topk, float* out_center,
                     b2me_stream_t stream);

/* predict_translation "magic" (app/inference_engine.py:459-489): per segment, p' = R(q)^T p,
 * o = (max+min)/2, pos = R ([-0.015, 0, min_z(p'-o)] + o).  quat [S,4] wxyz f32, out [S,3] f32 */
int b2me_translation_magic(const float* points_xyz, const int32_t* seg_offsets, int S,
                           const float* quat_wxyz, float x_offset, double* out_pos,
                           b2me_stream_t stream);

/* InferenceEngine.check_sanity (app/inference_engine.py:246-279) with get_6_key_points (utils/data.py:255-335) and
 * compute_kp_error (utils/metrics.py:130-136), batched over EE crops: confident[s] = 0 when the crop has fewer than
 * min_points points, when an EE corner is not within 0.04 m of where the pose puts it, or when the mean distance of
 * the predicted key points (class k of crop s counts when kp_prob[s,k] > kp_threshold; kp_xyz[s,k] = its coordinates)
 * to the corners / gripper tips found on the crop exceeds kp_margin. points_xyz [n,3] f32, seg_offsets [S+1] i32,
 * ee_pose [S,7] f64 (x y z qw qx qy qz, the pose BEFORE the ICP refinement), kp_prob [S,K] f32, kp_xyz [S,K,3] f32
 * (K <= 6; both may be null with K = 0). */
int b2me_sanity_check(const float* points_xyz, const int32_t* seg_offsets, int S, const double* ee_pose,
                      const float* kp_prob, const float* kp_xyz, int K, float kp_threshold, int min_points,
                      double kp_margin, uint8_t* out_confident, b2me_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K7: ClusterUtil.get_largest_cluster (utils/output.py:13-28; sklearn single linkage, 0.06 m) batched:
 * connected components of the graph {d(i,j) < dist} inside every segment; mask[i] = 1 for the points
 * of the largest component of their segment (ties: the component holding the lowest point index).
 * ---------------------------------------------------------------------------------------------- */
size_t b2me_cluster_workspace_bytes(int64_t n, int S);
int b2me_largest_cluster(const float* points_xyz, const int32_t* seg_offsets, int S, int64_t n,
                         double dist, uint8_t* out_mask,
                         int32_t* out_sizes /* [S] largest size; -1 everywhere when a point was non-finite or outside
                                               the cell key range (the masks are then not meaningful) */,
                         void* ws, size_t ws_bytes, b2me_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K9: get_rigid_transform_3D (utils/transformation.py:178-222) batched. ref/tgt [P,kmax,3] f64,
 * npairs [P] i32 (<= kmax). out_R [P,9] f64 row-major, out_t [P,3] f64.
 * ---------------------------------------------------------------------------------------------- */
int b2me_kabsch_batched(const double* ref, const double* tgt, const int32_t* npairs, int P, int kmax,
                        double* out_R, double* out_t, b2me_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K10: Open3D registration_icp(source=CAD, target=EE, max_corr, init, PointToPoint) as used by
 * utils/icp.py:50-81, batched over frames. ONE launch for all iterations: a thread-block cluster per frame keeps the
 * frame's targets in shared memory; exact nearest neighbour within max_corr through a uniform grid over them.
 *   source_xyz [S,3] f32 (shared CAD cloud); target_xyz [T_total,3] f32, tgt_offsets [F+1] i32
 *   init_T [F,16] f64 row-major 4x4; out_T [F,16] f64; out_stats [F,4] f64 = fitness, inlier_rmse,
 *   iterations run, #correspondences
 * ---------------------------------------------------------------------------------------------- */
size_t b2me_icp_workspace_bytes(int64_t T_total, int F, int S);
int b2me_icp_p2p_batched(const float* source_xyz, int S, const float* target_xyz,
                         const int32_t* tgt_offsets, int F, int64_t T_total, const double* init_T,
                         double max_corr, int max_iter, double rel_fitness, double rel_rmse,
                         double* out_T, double* out_stats, void* ws, size_t ws_bytes,
                         int cluster_size /* CTAs per frame: 1, 2, 4 or 8; 0 = as many as the source cloud fills */,
                         b2me_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Ingest: organised PointCloud2 / PCD records of a batch of frames -> the tensors K1 takes.
 * Replaces utils/ros_utils.py:142-167 (drop non-finite points, split the PCL-packed rgb), rgb / 255
 * (app/freenect_data_engine.py:81), utils/preprocess.py:20-37 (rgb - 0.5) and utils/data.py:58-75 (ROI mask).
 *   xyzrgb [n,4] f32: x, y, z, rgb bits 0x00RRGGBB;  frame_offsets [F+1] i32 (device) records of every frame
 *   roi6 HOST float[6] = min_x, max_x, min_y, max_y, min_z, max_z (strict inequalities), null = +-500
 *   out_xyz [n,3] f32, out_rgb [n,3] f32 in [-0.5, 0.5], out_bidx [n] f32 frame index, out_src [n] i32 (may be
 *   null) input record of every kept point, out_offsets [F+1] i32 compacted frame offsets (out_offsets[F] = kept)
 * Order preserving. ---------------------------------------------------------------------------------------------- */
size_t b2me_ingest_workspace_bytes(int64_t n);
int b2me_ingest_clouds(const float* xyzrgb, int64_t n, const int32_t* frame_offsets, int F, const float* roi6,
                       float* out_xyz, float* out_rgb, float* out_bidx, int32_t* out_src, int32_t* out_offsets,
                       void* ws, size_t ws_bytes, b2me_stream_t stream);

/* Order-preserving selection of the rows with key[i] == want (want < 0: key[i] != 0), e.g. the EE points
 * np.where(seg == 2) of every frame of a batch (app/inference_engine.py:422-433). key [n] u8; src_rows [n] i32 maps
 * row i to the row id that is written (null = i); seg_offsets [S+1] i32 (device) rows of every segment.
 * out_rows [n] i32 capacity, out_offsets [S+1] i32: selected rows before segment s (out_offsets[S] = selected). */
size_t b2me_select_workspace_bytes(int64_t n);
int b2me_select_rows(const uint8_t* key, int want, const int32_t* src_rows, int64_t n, const int32_t* seg_offsets,
                     int S, int32_t* out_rows, int32_t* out_offsets, void* ws, size_t ws_bytes, b2me_stream_t stream);
/* crop compaction: out_xyz[i] = xyz[rows[i]], out_rgb[i] = rgb[rows[i]] (may be null), out_seg[i] = segment of row i
 * of the selection as f32 (may be null; seg_offsets [S+1] i32 offsets INTO the selection) */
int b2me_gather_crops(const float* xyz, const float* rgb, const int32_t* rows, int64_t m, const int32_t* seg_offsets,
                      int S, float* out_xyz, float* out_rgb, float* out_seg, b2me_stream_t stream);
/* center_at_origin (utils/preprocess.py:8-11) per segment: out_center [S,3] = (max + min) / 2 (float32),
 * out_centered [n,3] = points - centre of their segment (may be null) */
int b2me_center_segments(const float* points_xyz, const int32_t* seg_offsets, int S, float* out_center,
                         float* out_centered, b2me_stream_t stream);

/* normalize_colors (utils/preprocess.py:20-37) applied per frame of a batch on the device: rgb [n,3] f32 -> out [n,3]
 * (may alias rgb): / 255 when the frame's maximum exceeds 2, per-channel min-max rescale when the frame has a negative
 * value, - 0.5 when the result lies in [0, 1]. bidx [n] f32 frame index of every point, frame_offsets [F+1] i32 (device)
 * rows of every frame, ws: F * 24 bytes. */
int b2me_normalize_colors(const float* rgb, const float* bidx, int64_t n, const int32_t* frame_offsets, int F,
                          float* out, void* ws, size_t ws_bytes, b2me_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * PointNet++ primitives of the default key-point network (model/pointnet2.py:9-43 through
 * model/pointnet2_utils.py). All clouds of a batch have N points.
 * ---------------------------------------------------------------------------------------------- */

/* farthest_point_sample (model/pointnet2_utils.py:65-86): xyz [B,N,3] f32, start [B] i32 first sample of every cloud
 * (the reference draws it with torch.randint; null = 0), out_idx [B,npoint] i32. N <= 8192. Ties: lowest index. */
int b2me_fps(const float* xyz, int B, int N, int npoint, const int32_t* start, int32_t* out_idx,
             b2me_stream_t stream);

/* query_ball_point (model/pointnet2_utils.py:89-110): out_idx [B,S,nsample] i32 = the first nsample indices
 * (ascending) with squared distance <= radius^2 to new_xyz [B,S,3], padded with the first; N when none. */
int b2me_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float radius, int nsample,
                    int32_t* out_idx, b2me_stream_t stream);

/* 3-NN of PointNetFeaturePropagation (model/pointnet2_utils.py:283-293): for every point of xyz1 [B,N,3] the three
 * nearest points of xyz2 [B,S,3] (out_idx [B,N,3] i32) and the weights (1/(d+1e-8)) / sum (out_w [B,N,3] f32). */
int b2me_three_nn(const float* xyz1, const float* xyz2, int B, int N, int S, int32_t* out_idx, float* out_w,
                  b2me_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2ME_H_ */
