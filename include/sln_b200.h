/*
 * sln_b200.h — C ABI of the B200-native 3D_SLN hot path (libsln_b200.so).
 *
 * The reference (aluo-x/3D_SLN) has no FFI/plugin layer: its hot path is reached through Python nn.Module calls.
 * The drop-in boundary is therefore the Python class surface (sln_b200/models/*.py mirrors the reference's
 * models/graph.py, models/Sg2ScVAE_model.py, ...) and THIS header is what those Python classes bind with ctypes.
 * Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain device pointers + extents; no torch types; every buffer (inputs, outputs, workspace) is caller-owned
 *   - return 0 on success, negative SLN_E* / on failure; message via sln_last_error() (thread-local)
 *   - all launches go to `stream` (a cudaStream_t passed as void*); no host synchronisation, no allocation,
 *     so every call is CUDA-graph capturable
 *   - all floating point is fp32, row-major; index tensors are int64 exactly as suncg_collate_fn produces them
 *     (reference data/suncg_dataset.py:295-337)
 */
#ifndef SLN_B200_H
#define SLN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLN_ABI_VERSION 2

int sln_version(void);
const char* sln_last_error(void);

/* Launch accounting and per-kernel-class timing (used by bench.py; the reference only has the unused wall-clock helper
 * utils.py:127-137 `timeit`).  sln_launch_count(): kernels launched by this library since load.  With profiling enabled
 * (never during CUDA-graph capture) every launch is bracketed by CUDA events on its stream; sln_prof_read() sums, for one
 * class, the device milliseconds, the algorithmic work (FLOPs for contractions, bytes for the rest) and the launch count
 * since the last sln_prof_enable() call (which also clears the records).
 * classes: 0 gemm_fwd 1 gemm_bwd_x 2 gemm_bwd_w 3 pool 4 prep 5 misc 6 raster_fwd 7 raster_bwd 8 spade_conv 9 spade_misc */
int64_t sln_launch_count(void);
int sln_prof_enable(int on);
int sln_prof_read(int cls, double* ms, double* work, int64_t* launches);

/* ------------------------------------------------------------------------------------------------ VAE-graph model
 * Model description: mirrors the ctor kwargs of Sg2ScVAEModel (reference models/Sg2ScVAE_model.py:7-113,
 * build_dataset_model.py:40-52). */
typedef struct sln_vae_desc {
  int32_t embedding_dim;  /* E (default 64): gconv dim D = 2E, hidden H = 4E, latent Z = E */
  int32_t n_layers;       /* gconv_num_layers (>= 1) */
  int32_t recurrent;      /* gconv_mode == 'recurrent': one weight set reused by every layer */
  int32_t norm;           /* 0 = mlp_normalization 'none', 1 = 'batch' */
  int32_t training;       /* module.training: BN uses batch statistics and updates running stats */
  int32_t box_dim;        /* 6 (train_3d) or 4 */
  int32_t n_angle;        /* Nangle = 24 */
  int32_t num_objs;       /* rows of obj embedding tables = len(vocab.object_idx_to_name) + 1 */
  int32_t num_preds;
  int32_t num_attrs;
  float bn_eps;           /* 1e-5 */
  float bn_momentum;      /* 0.1 */
  int32_t gconv_dim_override;    /* 0, or Din = Dout of a standalone GraphTripleConv (sln_gconv_layer_*) */
  int32_t gconv_hidden_override; /* 0, or hidden_dim of a standalone GraphTripleConv */
  const void* packed_weights;    /* NULL, or the buffer sln_vae_pack_weights() filled FROM THE CURRENT PARAMETER VALUES: pre-split
                                    (TF32 hi | lo), pre-tiled images of every Linear weight, pulled into shared memory by cp.async.bulk
                                    instead of being split on the fly (forward and backward-data contractions) */
  const void* bn_sync;           /* NULL (per-rank BatchNorm statistics), or a DEVICE pointer to an sln_bn_sync table: training-mode
                                    BatchNorm statistics are then summed over all ranks INSIDE the finalising kernels (SyncBatchNorm) */
  const void* graph_ws;          /* decoder calls only (the encoder ignores it): NULL, or the workspace of the sln_vae_encoder_fwd call
                                    of THIS step that ran on the same objs / triples / attributes (same O, T, still intact): the decoder
                                    reads the encoder's CSR and int32 index arrays instead of rebuilding them (the reference's model
                                    passes the same graph to both halves, Sg2ScVAE_model.py:118-172) */
} sln_vae_desc;

/* Cross-rank BatchNorm statistics (SURVEY 8e policy P1: the N-GPU step equals the 1-GPU step at the global batch; reference
 * models/graph.py:14-15 nn.BatchNorm1d over ALL rows of the batch).  The kernel that finalises a BatchNorm layer's column sums (the
 * last CTA of a column block in the contraction epilogue / k_prep) stores its (sum, sum of squares | sum g, sum g*yhat, rows) to
 * EVERY rank's receive buffer with peer stores over NVLink, bumps every rank's flag for that slot with a system-scope atomic, spins
 * on its own flag until `world` arrivals, and reduces the `world` contributions in rank order (fixed order: bit-identical statistics
 * on every rank, run-to-run deterministic).  No NCCL call, no host round trip, capturable in a CUDA graph.
 * The table lives in device memory; all pointers are device addresses valid ON THIS RANK (peer-mapped, e.g. from
 * torch.distributed._symmetric_memory).  recv / flag / use must be zero-initialised once and never reset.
 *   slot  = direction * 2 * SLN_BN_SYNC_SLOTS + which * SLN_BN_SYNC_SLOTS + (layer's first counter + column block)
 *   recv[r] + ((slot * world + src_rank) * SLN_BN_SYNC_COLS + col) * 3   (doubles: a, b, rows)
 * Cross-step reuse of a slot is ordered by the per-step gradient all-reduce. */
#define SLN_BN_SYNC_MAX_WORLD 8
#define SLN_BN_SYNC_COLS 128
#define SLN_BN_SYNC_SLOTS 2304        /* counters of one encoder / decoder workspace */
typedef struct sln_bn_sync {
  int32_t world, rank;
  double* recv[SLN_BN_SYNC_MAX_WORLD];      /* recv[r]: rank r's receive buffer: 4 * SLN_BN_SYNC_SLOTS * world * SLN_BN_SYNC_COLS * 3 doubles */
  uint32_t* flag[SLN_BN_SYNC_MAX_WORLD];    /* flag[r]: rank r's arrival counters: 4 * SLN_BN_SYNC_SLOTS */
  uint32_t* use;                            /* this rank's use counters: 4 * SLN_BN_SYNC_SLOTS */
} sln_bn_sync;
size_t sln_bn_sync_recv_bytes(int32_t world);     /* bytes of one rank's receive buffer */
size_t sln_bn_sync_flag_bytes(void);              /* bytes of one rank's flag array (and of `use`) */

/* Parameter table.  `params[i]` / `grads[i]` are device pointers in this canonical order (grads may be NULL
 * for forward-only calls; individual entries may be NULL to skip a gradient):
 *   0 obj_embeddings_ec.weight  1 attr_embedding_ec.weight  2 angle_embeddings.weight  3 pred_embeddings_ec.weight
 *   4 obj_embeddings_dc.weight  5 attr_embedding_dc.weight  6 pred_embeddings_dc.weight
 *   then, per Linear block in the order below: weight, bias, and (norm == 1 and the block has a BatchNorm) bn.weight, bn.bias
 *     box_embeddings
 *     gconv_net_ec.gconvs[l].{net1.0, net1.<2nd Linear>, net2.0, net2.<2nd Linear>}   l = 0..(recurrent ? 0 : n_layers-1)
 *     gconv_net_dc.gconvs[l]....                                                        (all four have BN when norm == 1)
 *     box_mean_var.{0, 2nd}  angle_mean_var.{0, 2nd}          (BN)
 *     box_mean.0  box_var.0  angle_mean.0  angle_var.0         (no BN: make_mlp(norelu=True), graph.py:22-26)
 *     box_net.0 (BN)  box_net.<2nd> (no BN)  angle_net.0 (BN)  angle_net.<2nd> (no BN)
 * `bn_bufs`: for every BatchNorm in the same order: running_mean (float*), running_var (float*),
 *            num_batches_tracked (int64_t*).  NULL when norm == 0. */
int sln_vae_num_params(const sln_vae_desc* d);  /* entries of params[] / grads[] */
/* Weight images for sln_vae_desc::packed_weights: size of the buffer, and the (graph-capturable) kernel that fills it from params[].
 * Re-run after every optimizer step; passing a stale buffer computes with stale weights. */
size_t sln_vae_packed_bytes(const sln_vae_desc* d);
int sln_vae_pack_weights(const sln_vae_desc* d, const void* const* params, void* packed, size_t packed_bytes, void* stream);
int sln_vae_num_bn(const sln_vae_desc* d);      /* BatchNorm count; bn_bufs has 3x this many entries */

/* Workspace holding graph CSR, saved activations, BN statistics and backward scratch.  which: 0 = encoder, 1 = decoder,
 * 2 = one standalone GraphTripleConv layer. */
size_t sln_vae_workspace_bytes(const sln_vae_desc* d, int64_t O, int64_t T, int which);

/* Index validation.  The reference raises IndexError for an id outside its embedding table / node range (nn.Embedding,
 * Sg2ScVAE_model.py:121-129; obj_vecs[s_idx], graph.py:78-79).  The kernels never read or write out of bounds: a bad id is
 * remapped to row 0 and one of the bits below is OR-ed into an int32 flag that lives in the call's workspace at byte offset
 * sln_vae_index_flag_offset(d, O, T, which) (which: 0 encoder, 1 decoder, 2 standalone layer; -1 on bad arguments).  The flag is
 * rewritten by every *_fwd call; the CALLER reads it back (one 4-byte D2H copy) and raises — sln_b200's Python wrappers do so on
 * the first call of every (O, T) shape and on every call when `check_indices=True`, VAETrainStep.check_indices() on demand. */
#define SLN_IDX_OBJS 1      /* objs[i]       outside [0, num_objs)  */
#define SLN_IDX_ATTRS 2     /* attributes[i] outside [0, num_attrs) */
#define SLN_IDX_ANGLES 4    /* angles[i]     outside [0, n_angle)   */
#define SLN_IDX_NODES 8     /* triples[t,0|2] / edges outside [0, O) */
#define SLN_IDX_PREDS 16    /* triples[t,1]  outside [0, num_preds) */
int64_t sln_vae_index_flag_offset(const sln_vae_desc* d, int64_t O, int64_t T, int which);

/* Sg2ScVAEModel.encoder (reference Sg2ScVAE_model.py:115-143): objs[O] triples[T,3] boxes[O,box_dim] angles[O]
 * attributes[O] -> mu[O,E], logvar[O,E].  Saves what encoder_bwd needs in `ws`. */
int sln_vae_encoder_fwd(const sln_vae_desc* d, const void* const* params, void* const* bn_bufs,
                        const int64_t* objs, const int64_t* triples, const float* boxes, const int64_t* angles,
                        const int64_t* attributes, int64_t O, int64_t T, float* mu, float* logvar,
                        void* ws, size_t ws_bytes, void* stream);
/* Backward of the encoder: d_mu, d_logvar [O,E] -> parameter gradients ACCUMULATED into grads[] (caller zeroes). */
int sln_vae_encoder_bwd(const sln_vae_desc* d, const void* const* params, void* const* grads,
                        const float* boxes, const float* d_mu, const float* d_logvar, int64_t O, int64_t T,
                        void* ws, size_t ws_bytes, void* stream);

/* Sg2ScVAEModel.decoder (reference Sg2ScVAE_model.py:145-172, decoder_cat=True): z[O,E] objs triples attributes ->
 * boxes_pred[O,box_dim], angles_pred[O,n_angle] (log-probabilities). */
int sln_vae_decoder_fwd(const sln_vae_desc* d, const void* const* params, void* const* bn_bufs,
                        const float* z, const int64_t* objs, const int64_t* triples, const int64_t* attributes,
                        int64_t O, int64_t T, float* boxes_pred, float* angles_pred,
                        void* ws, size_t ws_bytes, void* stream);
/* d_angles is the gradient w.r.t. angles_pred (log-probs) when angles_are_logits == 0, else w.r.t. the pre-softmax
 * logits (what sln_vae_loss emits).  d_z [O,E] is written (may be NULL). */
int sln_vae_decoder_bwd(const sln_vae_desc* d, const void* const* params, void* const* grads,
                        const float* d_boxes, const float* d_angles, int angles_are_logits, float* d_z,
                        int64_t O, int64_t T, void* ws, size_t ws_bytes, void* stream);

/* GraphTripleConv.forward (reference models/graph.py:57-111), one layer, standalone.
 * layer_params: W1a,b1a,[g,b] W1b,b1b,[g,b] W2a,b2a,[g,b] W2b,b2b,[g,b]; layer_bn_bufs: 4 x (rm, rv, nbt).
 * obj_vecs[O,Din] pred_vecs[T,Din] edges[T,2] -> new_obj[O,Dout], new_pred[T,Dout].  Din = Dout = 2E of `d`, H = 4E. */
int sln_gconv_layer_fwd(const sln_vae_desc* d, const void* const* layer_params, void* const* layer_bn_bufs,
                        const float* obj_vecs, const float* pred_vecs, const int64_t* edges, int64_t O, int64_t T,
                        float* new_obj, float* new_pred, void* ws, size_t ws_bytes, void* stream);
int sln_gconv_layer_bwd(const sln_vae_desc* d, const void* const* layer_params, void* const* layer_grads,
                        const float* obj_vecs, const float* pred_vecs, const float* d_new_obj, const float* d_new_pred,
                        int64_t O, int64_t T, float* d_obj, float* d_pred, void* ws, size_t ws_bytes, void* stream);

/* The pooling stage alone (reference graph.py:92-108): edges -> CSR, then
 * pooled[o] = (sum_{s_t=o} new_s[t] + sum_{o_t=o} new_o[t]) / max(deg(o),1), new_t_vecs[T, 2H+Dout] laid out s|p|o.
 * Used by the roofline benchmark and the parity tests of the north-star "scatter" kernel. */
size_t sln_gconv_pool_workspace_bytes(int64_t O, int64_t T);
int sln_csr_build(const int64_t* edges, int64_t edge_stride, int64_t O, int64_t T, void* ws, size_t ws_bytes, void* stream);
int sln_gconv_pool_fwd(const float* new_t_vecs, int64_t O, int64_t T, int32_t H, int32_t Dout, float* pooled,
                       const void* ws, size_t ws_bytes, void* stream);
/* CSR read-back helpers for tests: copies row_ptr[O+1] / ent[2T] (device pointers inside ws) */
int sln_csr_pointers(void* ws, int64_t O, int64_t T, const int32_t** row_ptr, const int32_t** ent);

/* Scene assembly (SURVEY §8 a10 i-iii / §8f N3).  Replaces the per-object Python loop of mesh_render_func (reference
 * models/diff_render.py:76-159: scale = min(box size / model size), Ry(-angle * 2pi/24), trans = centre - scale * R * model centre,
 * vertices = scale * R * v + trans) and the near-plane face cull (:344-356) for meshes that stay resident on the device.
 *   boxes [n_rows,6] (objects normalised to the room), angles [n_rows] fp32, kept [n_kept] int32 = layout rows that own a mesh,
 *   room3_host = HOST array {x,y,z} of the room size, model_verts [n_obj_verts,3] (objects contiguous, in `kept` order),
 *   vert_obj [n_obj_verts] int32 = kept-object of each vertex, shell_verts [n_shell,3] = wall/floor/ceiling (copied through),
 *   model_size / model_center [n_kept,3], faces [F,3] int32 into the concatenated vertex array, R [3,3], t [3] = camera.
 *   -> vertices [n_obj_verts + n_shell, 3], sizes [n_kept,3] (box sizes, for the size loss :98), faces_out [F,3]: faces with a
 *   vertex at camera depth < cull_eps become the zero-area triangle (0,0,0) (static shapes; the rasterizer never draws them).
 * bwd: grad_vertices [V,3], grad_sizes [n_kept,3] or NULL, row_to_kept [n_rows] int32 (-1 = no mesh), vert_start [n_kept+1] int32
 *   -> d_boxes [n_rows,6], d_angles [n_rows], fully overwritten; fixed-order reductions (bit-reproducible).  ws: the buffer the
 *   forward call filled (sln_scene_assemble_workspace_bytes).  fix_corners != 0 / angle_grad_scale apply the refinement loop's gradient
 *   hooks in place (testing/test_render_refine.py:220-230: fix_grad averages the min- and max-corner gradients, quad_grad = x4; pass 0 / 1
 *   for plain gradients). */
size_t sln_scene_assemble_workspace_bytes(int64_t n_kept);
int sln_scene_assemble_fwd(const float* boxes, const float* angles, int64_t n_rows, const int32_t* kept, int64_t n_kept, const float* room3_host,
                           const float* model_verts, const int32_t* vert_obj, int64_t n_obj_verts, const float* shell_verts, int64_t n_shell,
                           const float* model_size, const float* model_center, const int32_t* faces, int64_t F, const float* R, const float* t,
                           float cull_eps, float* vertices, float* sizes, int32_t* faces_out, void* ws, size_t ws_bytes, void* stream);
int sln_scene_assemble_bwd(const float* grad_vertices, const float* grad_sizes, int64_t n_rows, const int32_t* row_to_kept, int64_t n_kept,
                           const float* room3_host, const float* model_verts, const int32_t* vert_start, const float* model_size,
                           const float* model_center, const void* ws, size_t ws_bytes, int32_t fix_corners, float angle_grad_scale,
                           float* d_boxes, float* d_angles, void* stream);

/* Compositing of the 70-channel render (reference models/diff_render.py:366-434, the per-class loop vectorised):
 *   depth [P] (depth render), images [C,P] (class masks, class order of the caller) ->
 *   out [1 + (n_onehot-1) + n_keep, P]: out[0] = d = depth > 15 ? -1 : depth (:367); out[ch], 1 <= ch < n_onehot = images[inv_index[ch]]
 *   (0 where inv_index[ch] < 0; :429-431); out[n_onehot + k] = images[keep[k]] > 0.1 ? d / wall_max : mean_c / wall_max (:401-421) with
 *   mean_c = mean of d over the class mask (wall_max when empty) and wall_max = max of d over the mask of class `wall` (10 when empty).
 *   stats [2C+1] = cnt | fill | wall_max is kept for the backward call.  index [C] / inv_index [n_onehot] / keep [n_keep]: device int32.
 * bwd: grad_out (same shape as out) -> grad_depth [P], grad_images [C,P], fully overwritten; the mask threshold and wall_max carry no
 * gradient, as in the reference (.detach()).  Two-level fixed-order reductions: bit-reproducible. */
size_t sln_composite_workspace_bytes(int32_t C);
int sln_composite_fwd(const float* depth, const float* images, int32_t C, int64_t P, int32_t wall, const int32_t* inv_index, int32_t n_onehot,
                      const int32_t* keep, int32_t n_keep, float* out, float* stats, void* ws, size_t ws_bytes, void* stream);
int sln_composite_bwd(const float* depth, const float* images, const float* grad_out, int32_t C, int64_t P, const int32_t* index, int32_t n_onehot,
                      const int32_t* keep, int32_t n_keep, const float* stats, float* grad_depth, float* grad_images, void* ws, size_t ws_bytes,
                      void* stream);

/* Fused multi-scale refinement loss.  Replaces testing/test_render_refine.py:332-352 + PSP_pool_new :192-215:
 *   image[1 + n_sem + n_dep, S, S] (the [1,70,256,256] render of mesh_render_func: channel 0 depth, 1..n_sem class masks, then the
 *   normalised per-class depth planes) -> null-fill of the last plane where the depth planes sum to < 0.5 (:333), every plane resized
 *   S -> sizes4[i] (bilinear, align_corners=True) -> sizes4[3] (bilinear, align_corners=False),
 *   loss3[0] = 100 * 0.5 * mean|pooled depth - t_depth| + 100 * sum_i CrossEntropy_i(pooled classes, t_labels[i]; ignore < 0) / 800,
 *   loss3[1], loss3[2] = the depth and semantic terms before the factor 100.
 * t_depth [4 * n_dep, top, top] (level-major, as torch.cat(priors, 1)), t_labels [4, top, top] int64, label_counts4 = HOST array with
 * the number of labels >= 0 per level (CrossEntropyLoss averages over them).  d_image [1 + n_sem + n_dep, S, S] (may be NULL):
 * d loss3[0] / d image, fully overwritten, computed without atomics (bit-reproducible). */
size_t sln_refine_loss_workspace_bytes(int32_t image_size, const int32_t* sizes4, int32_t n_sem, int32_t n_dep);
int sln_refine_loss(const float* image, int32_t image_size, const int32_t* sizes4, int32_t n_sem, int32_t n_dep, const float* t_depth,
                    const int64_t* t_labels, const float* label_counts4, float* loss3, float* d_image, void* ws, size_t ws_bytes,
                    void* stream);

/* GPU-side batch assembly.  Replaces suncg_collate_fn (reference data/suncg_dataset.py:295-337) + the eight
 * .cuda() copies of tensor_aug (utils.py:114-124).  The host packs the kept scenes of a batch into one pinned wire buffer:
 *   scene_index[B] i64 (position of the scene in the DataLoader batch, :323-324) | obj_off[B+1] i64 | tri_off[B+1] i64 (prefix
 *   sums) | ids[B] i64 | objs[O] i64 | angles[O] i64 | attributes[O] i64 | triples[T,3] i64 with SCENE-LOCAL ids | boxes[O,box_dim] f32
 * at the byte offsets sln_collate_layout() returns (offsets10 = those nine sections in this order, then the total size), copies
 * it to the device once, and sln_collate_finish() writes triples[T,3] with global ids (:318-320, may alias the wire section),
 * obj_to_img[O] and triple_to_img[T] (either may be NULL).  objs / angles / attributes / boxes / ids are used in place.
 * err_count (may be NULL; caller zeroes it): number of triples whose local ids fall outside their scene. */
int sln_collate_layout(int64_t B, int64_t O, int64_t T, int32_t box_dim, int64_t* offsets10);
int sln_collate_finish(const void* wire, size_t wire_bytes, int64_t B, int64_t O, int64_t T, int32_t box_dim, int64_t* triples,
                       int64_t* obj_to_img, int64_t* triple_to_img, int32_t* err_count, void* stream);

/* The contraction primitive under every Linear of the MLP stage (reference graph.py:12 nn.Linear -> aten::addmm / mm):
 *   C[i,j] (+)= sum_k A(i,k) * B(j,k),   i < M, j < N, k < K
 * a_rc != 0: A stored [M][K] row-major with leading dimension lda, else stored [K][M] (the transposed reads of the backward
 * passes); same for B.  accumulate != 0: C += (split-K, RED.ADD) else C = .  engine 1 = tcgen05 3xTF32 tiles (fp32-level
 * accuracy: hi/lo TF32 split, fp32 accumulation in TMEM), engine 0 = FP32 SIMT tiles.  sln_set_engine selects the engine
 * used inside the sln_vae_* and sln_gconv_* entry points (default 1) and returns the previous setting. */
int sln_set_engine(int engine);
int sln_contract(const float* A, int64_t lda, int32_t a_rc, const float* B, int64_t ldb, int32_t b_rc, float* C, int64_t ldc,
                 int64_t M, int64_t N, int64_t K, int32_t accumulate, int32_t engine, void* stream);

/* z = eps*exp(0.5*logvar)+mu and its backward (reference Sg2ScVAE_model.py:180-183); n = O*E elements. */
int sln_reparam_fwd(const float* mu, const float* logvar, const float* eps, int64_t n, float* z, void* stream);
int sln_reparam_bwd(const float* d_z, const float* logvar, const float* eps, int64_t n, float* d_mu, float* d_logvar, void* stream);

/* calculate_model_losses (reference utils.py:12-33): losses[4] = {bbox, angle, KL_weight*KLD, total} on the device and
 * the gradient seeds of total w.r.t. boxes_pred, angles (the pre-softmax logits when angles_grad_is_logits != 0, else the
 * log-probabilities angles_pred), mu and logvar (the d_* may be NULL for a loss-only call).
 * mu == NULL selects use_AE.  scratch: >= 16 + 12*ceil(O/64) bytes, zero-initialised once by the caller. */
int sln_vae_loss(const float* boxes_pred, const float* boxes_gt, int32_t box_dim, const float* angles_pred,
                 const int64_t* angles_gt, int32_t n_angle, const float* mu, const float* logvar, int32_t Z,
                 float kl_weight, int64_t O, float* losses, float* d_boxes, float* d_angles, int32_t angles_grad_is_logits,
                 float* d_mu, float* d_logvar, void* scratch, size_t scratch_bytes, void* stream);

/* Same, with the KL weight read from device memory at launch time: a captured CUDA graph follows the reference's KL_linear_decay
 * schedule (train.py:73-74) without re-capture. */
int sln_vae_loss_dyn(const float* boxes_pred, const float* boxes_gt, int32_t box_dim, const float* angles_pred,
                     const int64_t* angles_gt, int32_t n_angle, const float* mu, const float* logvar, int32_t Z,
                     const float* kl_weight_dev, int64_t O, float* losses, float* d_boxes, float* d_angles,
                     int32_t angles_grad_is_logits, float* d_mu, float* d_logvar, void* scratch, size_t scratch_bytes, void* stream);

/* torch.optim.Adam step (reference train.py:15,82-84) over one flat fp32 arena.  *step is a device int64; when
 * advance_step != 0 it is incremented first, so the launch is identical every iteration (graph-safe); pass 0 for the
 * 2nd..nth arena of the same optimizer step.  grad_scale multiplies g (1/world_size for gradient averaging). */
int sln_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                  float beta2, float eps, float weight_decay, float grad_scale, int64_t* step, int32_t advance_step,
                  void* stream);
/* Same, graph-friendly: the learning rate is read from device memory (*lr_dev) at launch time, and when guard_loss != NULL the
 * whole update (moments, parameters, step counter) is skipped if *guard_loss is not finite — the reference's "not backpropping"
 * guard of train.py:78-80 without a host round trip. */
int sln_adam_step_dyn(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr_dev,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale, int64_t* step,
                      int32_t advance_step, const float* guard_loss, void* stream);

/* ------------------------------------------------------------------------------------------------ mesh rasterizer
 * Replaces what the reference reaches through the un-vendored `neural_renderer` package: nr.Renderer(camera_mode=
 * 'projection', image_size, K, R, t, anti_aliasing=False, orig_size, near, ...) called with mode='depth' and mode='rgb'
 * (reference models/diff_render.py:359-366,398; package semantics in SURVEY.md App. C / oracle/raster_oracle.c).
 * All maps are in the renderer's INTERNAL orientation (row yi = bottom-up); the Python Renderer flips rows on output as
 * upstream's rasterize.py does.  fill_back doubles the face list: face F+f is face f with reversed winding.
 *
 * sln_raster_setup       vertices [V,3] (world), faces [F,3] int32, K [9], R [9], t [3] (device) -> workspace:
 *                        projected vertices [V,3] (nr.projection, distortion-free per README.md:13-18), per-face vertex
 *                        array [F2,3,3] (nr.vertices_to_faces), pixel-space inverse [F2,9], pixel bounding boxes.
 * sln_raster_forward     z-buffer: face_index_map [is,is] int32 (-1 = background; ties keep the lower face index),
 *                        weight_map [is,is,3], depth_map [is,is] (far on empty pixels); near/far per call because upstream's
 *                        depth pass uses the rasterizer defaults (0.1, 100) while the rgb pass uses the constructor's near.
 * sln_raster_texture_sample  trilinear face-texture sampling, textures [F,ts,ts,ts,3] -> rgb_map [is,is,3] (background 0).
 * sln_raster_backward_rgb    Kato's gradient of the rgb image w.r.t. the face vertices' x,y: grad_faces [F2,9] += (caller zeroes).
 *                            (both rgb backward calls first mark, in a scratch array inside ws, the faces that own a pixel of
 *                            face_index_map: a face that shows nowhere contributes to neither sweep and is skipped)
 * sln_raster_backward_depth  gradient of depth_map w.r.t. face vertices: grad_faces [F2,9] +=.
 * sln_raster_vertex_grad     grad_faces [F2,9] -> grad w.r.t. the world vertices [V,3] (transpose of vertices_to_faces, then
 *                            the projection's Jacobian); grad_proj_scratch is caller-owned scratch of 24 V BYTES (64-bit fixed-point accumulators:
 *                            order-independent, hence bit-reproducible, scatter of the face gradients onto shared vertices).
 * sln_scene_classes_fwd/bwd  the 32 per-class mask renders of mesh_render_func (diff_render.py:381-431) from ONE rasterization:
 *                            class image c = what a 0/1 texture of class c renders to (torch.sum(images,1)/3), written in OUTPUT
 *                            orientation [n_cls,is,is]; the backward applies Kato's per-render clamp per class. */
size_t sln_raster_workspace_bytes(int64_t V, int64_t F, int32_t fill_back);
int sln_raster_setup(const float* vertices, int64_t V, const int32_t* faces, int64_t F, int32_t fill_back, const float* K,
                     const float* R, const float* t, float orig_size, int32_t image_size, void* ws, size_t ws_bytes, void* stream);
int sln_raster_face_arrays(void* ws, int64_t V, int64_t F, int32_t fill_back, const float** proj_vertices,
                           const float** face_vertices, const float** face_inv);
int sln_raster_forward(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float near, float far,
                       int32_t* face_index_map, float* weight_map, float* depth_map, void* stream);
/* Two z-buffers with different near planes from ONE pass over the faces (maps *_a clip at near_a, maps *_b at near_b): what
 * mesh_render_func needs — depth render at the rasterizer default near, class renders at the constructor's near (diff_render.py:366,398). */
int sln_raster_forward2(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float near_a, float near_b, float far,
                        int32_t* face_index_map_a, float* weight_map_a, float* depth_map_a, int32_t* face_index_map_b, float* weight_map_b,
                        float* depth_map_b, void* stream);
int sln_raster_texture_sample(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, const float* textures,
                              int32_t texture_size, float eps, const int32_t* face_index_map, const float* weight_map,
                              const float* depth_map, float* rgb_map, void* stream);
int sln_raster_backward_rgb(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float eps,
                            const int32_t* face_index_map, const float* rgb_map, const float* grad_rgb_map, float* grad_faces,
                            void* stream);
int sln_raster_backward_depth(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size,
                              const int32_t* face_index_map, const float* weight_map, const float* depth_map,
                              const float* grad_depth_map, float* grad_faces, void* stream);
int sln_raster_vertex_grad(const void* ws, const float* vertices, int64_t V, const int32_t* faces, int64_t F, int32_t fill_back,
                           const float* K, const float* R, const float* t, float orig_size, const float* grad_faces,
                           float* grad_proj_scratch, float* grad_vertices, void* stream);
int sln_scene_classes_fwd(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, int32_t texture_size, float eps,
                          const int32_t* face_index_map, const float* weight_map, const float* depth_map, const int32_t* face_cls,
                          int32_t n_cls, float* sval, float* class_images, void* stream);
int sln_scene_classes_bwd(const void* ws, int64_t V, int64_t F, int32_t fill_back, int32_t image_size, float eps,
                          const int32_t* face_index_map, const int32_t* face_cls, int32_t n_cls, const float* sval,
                          const float* grad_class_images_internal, float* grad_faces, void* stream);

/* ------------------------------------------------------------------------------------------------ SPADE generator
 * Inference path of SPADEGenerator4 (reference models/SPADE_related.py:1507-1605, instantiated at testing/test_SPADE_shade.py:9).
 * Activations are NHWC fp32; weights are packed by the caller ONCE per weight version: convolution kernels as
 * [Cout][ky][kx][Cin], spectral norm folded (W / (u^T W v), eval mode).
 *
 * sln_spade_conv        ReflectionPad2d(ks/2) + Conv2d(Cin, Cout, ks, padding=0) (ks = 3; SPADE_related.py:1429-1436,1472-1474) or a
 *                       1x1 convolution / nn.Linear (ks = 1; conv_s :1476, fc :1521) as an implicit GEMM; relu_in != 0 applies
 *                       ReLU to the input on load; bias may be NULL.  x [B,H,W,Cin] -> out [B,H,W,Cout].
 * sln_spade_modulate    the gamma AND beta convolutions of one SPADE4 (:1448-1449) as ONE contraction plus the modulation
 *                       out = act((x - mean[b]) * inv[b] * (1 + gamma) + beta) (:1451; act = leaky_relu(slope), slope 1 = none):
 *                       Wgb [2C][9*Ca] holds, per tile of `pair` rows, the gamma rows of pair/2 channels followed by the beta
 *                       rows of the same channels; gamma / beta are never written to memory.
 * sln_spade_ln_stats    LayerNorm2D(affine=False) statistics (:139-144): per sample mean and 1/(unbiased std + eps) over
 *                       n_per_sample = C*H*W values; scratch >= 16*B bytes.
 * sln_spade_seg_features  SPADE4 part 2 up to the concat (:1444-1447): resize the NCHW map seg [B,nc,S,S] to (h,w) (mode 0 =
 *                       bilinear align_corners=False, mode 1 = the nearest-neighbour map head_0 receives, :1579), depth channel
 *                       -> reflect-padded 3x3 conv to nd channels + LeakyReLU(0.01), concatenated with the nc-1 label channels
 *                       -> NHWC [B,h,w,nd+nc-1].
 * sln_spade_upsample2x  nn.Upsample(scale_factor=2, nearest | bilinear) (:1544-1545) on NHWC.
 * sln_spade_se_residual SEBlock2 (:81-85; W1 [Ch][C], W2 [C][Ch], no biases) on dx, then out = xs + dx * s (:1493);
 *                       scratch >= 4*(B*min(HW,64)*C + B*C) bytes, 16-byte aligned.
 * sln_spade_to_rgb      leaky_relu(slope) -> Conv2d(Cin, Cout <= 4, ks, zero padding ks/2) -> tanh (:1602-1603); NHWC in, NCHW
 *                       out; `pre` (optional) receives the pre-tanh values. */
/* sln_pack_weights: optional one-off transform of a weight matrix W [N][K] (row-major) into the pre-split (TF32 hi | lo), pre-tiled
 * image the contraction kernel can pull into shared memory with cp.async.bulk; pass the result as `Wpacked` / `Wgb_packed` (or NULL:
 * the weights are then split on the fly).  out: sln_packed_weights_bytes(N, K) bytes, 16-byte aligned. */
size_t sln_packed_weights_bytes(int64_t N, int64_t K);
int sln_pack_weights(const float* W, int64_t N, int64_t K, float* out, void* stream);
int sln_spade_conv(const float* x, int64_t B, int64_t H, int64_t W, int64_t Cin, int32_t ks, int32_t relu_in, const float* Wp, const float* Wpacked,
                   const float* bias, int64_t Cout, float* out, void* stream);
/* The same implicit-GEMM convolution with the padding selectable: pad_mode 0 = ReflectionPad2d(1) (SPADE4 / SPADEResnetBlock4,
 * reference SPADE_related.py:9-14,1429-1436), 1 = zeros = nn.Conv2d(kernel_size=3, padding=1) of the plain SPADE / SPADEResnetBlock
 * (:261-263,322-326). */
int sln_conv2d_nhwc(const float* x, int64_t B, int64_t H, int64_t W, int64_t Cin, int32_t ks, int32_t relu_in, int32_t pad_mode,
                    const float* Wp, const float* Wpacked, const float* bias, int64_t Cout, float* out, void* stream);

int sln_spade_modulate(const float* actv, int64_t B, int64_t H, int64_t W, int64_t Ca, const float* Wgb, const float* Wgb_packed, const float* bias_g,
                       const float* bias_b, int64_t C, int32_t pair, const float* x, const float* mean, const float* inv, float slope, float* out,
                       void* stream);
/* sln_spade_modulate with the parameter-free normalisation's statistics per (sample, channel): the value of (b, c) is read at
 * [b * stat_stride_b + c * stat_stride_c] — (1, 0) LayerNorm2D (SPADE4), (C, 1) nn.InstanceNorm2d(affine=False), (0, 1) eval-mode
 * nn.BatchNorm2d(affine=False) (plain SPADE, SPADE_related.py:308-316,328,337) — and the convolution padding selectable (pad_mode as
 * in sln_conv2d_nhwc).  inv is the multiplier: 1 / (std + eps) or 1 / sqrt(var + eps). */
int sln_spade_modulate_ex(const float* actv, int64_t B, int64_t H, int64_t W, int64_t Ca, const float* Wgb, const float* Wgb_packed,
                          const float* bias_g, const float* bias_b, int64_t C, int32_t pair, const float* x, const float* mean,
                          const float* inv, int32_t stat_stride_b, int32_t stat_stride_c, int32_t pad_mode, float slope, float* out,
                          void* stream);
/* nn.InstanceNorm2d(affine=False, track_running_stats=False) statistics of an NHWC tensor: mean[B*C], inv[B*C] = 1/sqrt(biased var + eps). */
int sln_instnorm_stats(const float* x, int64_t B, int64_t HW, int64_t C, float eps, float* mean, float* inv, void* stream);
/* out = act((x - mean[b,c]) * inv[b,c]), act 0 none / 1 ReLU: the norm + activation of a Conv2dBlock (SPADE_related.py:58-63). */
int sln_norm_act(const float* x, int64_t B, int64_t HW, int64_t C, const float* mean, const float* inv, int32_t stat_stride_b,
                 int32_t stat_stride_c, int32_t act, float* out, void* stream);
/* F.interpolate of the NCHW label map [B, nc, S, S] to (h, w), bilinear (align_corners=False, SPADE_related.py:330) or nearest (:226),
 * written NHWC with the channels zero-padded to cpad (a multiple of 4). */
int sln_seg_resize_nhwc(const float* seg, int64_t B, int32_t nc, int32_t S, int32_t nearest, int64_t h, int64_t w, int32_t cpad, float* out,
                        void* stream);

int sln_spade_ln_stats(const float* x, int64_t B, int64_t n_per_sample, float eps, void* scratch, float* mean, float* inv, void* stream);
int sln_spade_seg_features(const float* seg, int64_t B, int32_t nc, int32_t S, int32_t mode, int64_t h, int64_t w, const float* dw, const float* db,
                           int32_t nd, float* out, void* stream);
int sln_spade_upsample2x(const float* x, int64_t B, int64_t H, int64_t W, int64_t C, int32_t bilinear, float* out, void* stream);
int sln_spade_se_residual(const float* dx, const float* xs, int64_t B, int64_t H, int64_t W, int64_t C, const float* W1, const float* W2, int32_t Ch,
                          void* scratch, size_t scratch_bytes, float* out, void* stream);
int sln_spade_to_rgb(const float* x, int64_t B, int64_t H, int64_t W, int64_t Cin, const float* Wt, const float* bias, int32_t Cout, int32_t ks,
                     float slope, float* pre, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SLN_B200_H */
