/*
 * dcap.h -- C ABI of the B200-native per-RoI captioning path (libdcap.so, sm_100a).
 *
 * The reference (frosinastojanovska/image-captioning) exposes NO FFI for this path: its boundary
 * is the Keras layer `PyramidROIAlign` and the Keras models returned by `build_lstm_model` /
 * `build_model`.  Each entry point below cites the reference interface it replaces (paths relative
 * to /root/reference); the Python shims in image-captioning_b200/ mirror the Keras surface on top
 * of these calls, and INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types;
 *   - unless a function name ends in `_host`, all data pointers are DEVICE pointers, row-major,
 *     caller-owned; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous with respect to the host;
 *   - `_host` variants take HOST pointers, perform the host<->device copies themselves and return
 *     after the result is in the host buffer (this is what a Keras `predict` caller sees);
 *   - return value: 0 = OK, negative = error (DC_ERR_*); dc_last_error() returns a thread-local
 *     human-readable message;
 *   - re-entrant: no global mutable state besides per-handle buffers; a handle must not be used
 *     from two threads at once.
 */
#ifndef DCAP_H_
#define DCAP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DC_OK               0
#define DC_ERR_INVALID     -1   /* bad argument (shape, null pointer, unsupported size)   */
#define DC_ERR_CUDA        -2   /* CUDA runtime error; message holds cudaGetErrorString   */
#define DC_ERR_UNSUPPORTED -3   /* valid request this build cannot serve                  */
#define DC_ERR_STATE       -4   /* handle used before weights were set, etc.              */

/* Thread-local message describing the last error returned on this thread. */
const char *dc_last_error(void);

/* Library / device introspection: writes the SM count and compute capability (major*10+minor)
 * of the current device.  Returns DC_ERR_CUDA when no device is usable. */
int dc_device_info(int *sm_count, int *compute_capability);

/* ------------------------------------------------------------------------------------------
 * RoI feature stage
 * ---------------------------------------------------------------------------------------- */

/*
 * FPN level of every box: min(5, max(2, 4 + int32(round_half_even(log2(sqrt(h*w) /
 * (224/sqrt(img_h*img_w))))))).  Replaces the level-assignment prologue of
 * PyramidROIAlign.call, evaluate_models/modified_dense_model.py:351-363 (log2_graph :313-315).
 * boxes: [n_boxes,4] normalised (y1,x1,y2,x2) fp32; levels: [n_boxes] int32.
 */
int dc_fpn_levels_f32(const float *boxes, int64_t n_boxes, int img_h, int img_w, int32_t *levels,
                      void *stream);

/*
 * PyramidROIAlign.call, evaluate_models/modified_dense_model.py:343-416 (identical copies:
 * mask_rcnn/mask_rcnn_model.py:317-423, dense_img_cap/dense_model.py:313-419, ...):
 * level assignment, per-level tf.image.crop_and_resize(bilinear, one sample per bin, end points
 * inclusive, out-of-range sample -> 0) and the re-ordering back to (image, box) order, in ONE
 * pass that writes every RoI directly to its final slot.
 *
 *   boxes   [n_images, n_boxes, 4] fp32 normalised (y1,x1,y2,x2)
 *   fmaps   HOST array of 4 DEVICE pointers: level-l map [n_images, fm_h[l], fm_w[l], channels]
 *           fp32 NHWC (P2..P5)
 *   out     [n_images*n_boxes, pool_h, pool_w, channels] fp32  (the reference's literal output
 *           is this with a leading 1: [1, n_images*n_boxes, ...])
 *   levels  optional [n_images*n_boxes] int32 (NULL to skip)
 *
 * channels must be a multiple of 4 and every map pointer 16-byte aligned; n_boxes <= 100000
 * (the reference's sort key batch*100000+box, :408, collides beyond that -> DC_ERR_INVALID).
 */
int dc_pyramid_roi_align_f32(const float *boxes, const float *const fmaps[4], const int fm_h[4],
                             const int fm_w[4], int n_images, int n_boxes, int channels,
                             int pool_h, int pool_w, int img_h, int img_w, float *out,
                             int32_t *levels, void *stream);

/* Same contract, output written as bf16 (round-to-nearest-even of the fp32 result): the form the
 * bf16 RoI head consumes directly, halving the output traffic.  out is uint16_t bf16 bits. */
int dc_pyramid_roi_align_bf16out(const float *boxes, const float *const fmaps[4],
                                 const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                 int channels, int pool_h, int pool_w, int img_h, int img_w,
                                 uint16_t *out, int32_t *levels, void *stream);

/*
 * Gradient of PyramidROIAlign with respect to the four feature maps (TF CropAndResizeGradImage summed over the
 * per-level crops; the boxes receive no gradient: tf.stop_gradient, evaluate_models/modified_dense_model.py:379-380).
 * This is what the joint model of dense_img_cap/dense_model.py:738-755 back-propagates through the layer.
 *   grad_out  [n_images*n_boxes, pool_h, pool_w, channels] fp32 (device), (image, box) order
 *   d_fmaps   HOST array of 4 DEVICE pointers [n_images, fm_h[l], fm_w[l], channels] fp32, ACCUMULATED into
 *             (the caller zeroes them or passes running sums); fp32 atomic accumulation, order not deterministic
 */
int dc_pyramid_roi_align_backward_f32(const float *boxes, const float *grad_out, float *const d_fmaps[4],
                                      const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                      int channels, int pool_h, int pool_w, int img_h, int img_w, void *stream);

/*
 * Host-buffer form of dc_pyramid_roi_align_f32 (what `generate_features` /
 * `keras_model.predict` callers see, evaluate_models/generate_one_roi_features.py:69-76):
 * all pointers are HOST memory (pinned memory makes the copies asynchronous); the call stages the
 * pyramid image by image so copies overlap the kernel, and returns when `out` is complete.
 */
int dc_pyramid_roi_align_host_f32(const float *boxes, const float *const fmaps[4],
                                  const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                  int channels, int pool_h, int pool_w, int img_h, int img_w,
                                  float *out, int32_t *levels);

/* ------------------------------------------------------------------------------------------
 * RoI head + "inject" LSTM caption decoder
 * ---------------------------------------------------------------------------------------- */

#define DC_ARCH_V1         1   /* build_lstm_model: text_generation_model.py:235-283          */
#define DC_ARCH_V2_INJECT  2   /* build_model(inject=True): text_generation_model_v2.py:140-166 */
#define DC_DTYPE_F32       0   /* FFMA GEMMs: greedy ids equal the fp32 model's               */
#define DC_DTYPE_BF16      1   /* tcgen05 GEMMs, bf16 operands, fp32 accumulate and state     */

/* What `feats` points at in the decoder calls. */
#define DC_FEATS_ROI_F32   0   /* [B, pool, pool, channels] fp32 (ROIAlign output)            */
#define DC_FEATS_HEAD_F32  1   /* [B, feat] fp32: output of the RoI head (`features_new`,
                                  text_generation_model.py:261)                               */
#define DC_FEATS_ROI_BF16  2   /* [B, pool, pool, channels] bf16 (dc_pyramid_roi_align_bf16out) */

typedef struct DcDecoder DcDecoder;

/* Mirrors the reference's DenseCapConfig fields the text models read
 * (text_generation_model.py:23-49: PADDING_SIZE, VOCABULARY_SIZE, EMBEDDING_SIZE, POOL_SIZE)
 * plus the `units` argument of build_lstm_model / build_model. */
typedef struct DcDecoderConfig {
    int arch;        /* DC_ARCH_*                                                              */
    int dtype;       /* DC_DTYPE_*                                                             */
    int vocab;       /* VOCABULARY_SIZE                                                        */
    int embed;       /* EMBEDDING_SIZE                                                         */
    int feat;        /* RoI head width (1024)                                                  */
    int units;       /* LSTM units (v1: both LSTMs; v2: the `imgcap_lstm` layer)               */
    int word_units;  /* v2 only: width of the word LSTM (1024 in the reference)                */
    int pool;        /* POOL_SIZE (7)                                                          */
    int channels;    /* feature-map channels (256)                                             */
    int padding;     /* PADDING_SIZE (caption length P)                                        */
} DcDecoderConfig;

/* Creates a decoder on the CURRENT device.  Replaces the Keras model objects returned by
 * build_lstm_model(features_input, config, units, mode) (text_generation_model.py:235) and
 * build_model(features_shape, word_shape, config, units, inject) (text_generation_model_v2.py:140). */
int dc_decoder_create(const DcDecoderConfig *cfg, DcDecoder **out);
int dc_decoder_destroy(DcDecoder *dec);

/* Weights travel as fp32 HOST arrays in the Keras layout under the Keras weight name
 * ("<layer>/<kernel|recurrent_kernel|bias|gamma|beta|moving_mean|moving_variance|embeddings>"),
 * i.e. what model.get_weights()/set_weights()/load_weights(by_name=True) exchange
 * (text_generation_model.py:468,484).  LSTM kernels are [in,4u] / [u,4u] / [4u] with gate blocks
 * i|f|c|o; conv kernels are HWIO.  dc_decoder_weight_* enumerate the expected tensors. */
int dc_decoder_weight_count(const DcDecoder *dec);
const char *dc_decoder_weight_name(const DcDecoder *dec, int index);
int64_t dc_decoder_weight_numel(const DcDecoder *dec, int index);
int dc_decoder_set_weight(DcDecoder *dec, const char *name, const float *host, int64_t numel);
int dc_decoder_get_weight(DcDecoder *dec, const char *name, float *host, int64_t numel);
/* Builds the derived device buffers (folded BatchNorm, stacked / hoisted matrices, bf16 copies).
 * Must be called after the weights change and before any forward call. */
int dc_decoder_finalize(DcDecoder *dec, void *stream);

/* RoI head: relu(bn2(relu(bn1(conv7x7_valid(X)))*W2+b2)), text_generation_model.py:249-262.
 * feats per DC_FEATS_ROI_*; out [B, feat] fp32 (device). */
int dc_head_forward(DcDecoder *dec, const void *feats, int feats_kind, int B, float *out,
                    void *stream);

/* v1 greedy captioning = model.predict(features) of build_lstm_model(mode='inference'):
 * ROICaptionInferenceLayer, text_generation_model.py:192-232 -- start id 1, P steps, argmax
 * (first index on ties) fed back, no early stop -- as ONE incremental masked scan.
 *   tokens [B, P] int32 (device); probs optional [B, P, V] fp32 (device; the Keras output). */
int dc_decoder_greedy(DcDecoder *dec, const void *feats, int feats_kind, int B, int32_t *tokens,
                      float *probs, void *stream);

/* Greedy captioning that also returns the caption score of refine_generations
 * (evaluate_models/test_score_dense_captions.py:256-258): scores[b] = sum over the P steps of
 * log(max_v p) -- without materialising the [B,P,V] probabilities.  scores [B] fp32 (device). */
int dc_decoder_greedy_scored(DcDecoder *dec, const void *feats, int feats_kind, int B, int32_t *tokens,
                             float *scores, void *stream);

/* Beam search with the semantics of gen_captions (image captioning/test.py:23-64) applied to the
 * v1 decoder: width k, scores are SUMS OF PROBABILITIES (fp64 accumulation of fp32), children
 * pooled in generation order, stable ascending sort, keep the last k, no end-token handling.
 *   tokens [B, k, P] int32 (beams in ascending score order, best last); scores [B, k] fp64. */
int dc_decoder_beam(DcDecoder *dec, const void *feats, int feats_kind, int B, int k,
                    int32_t *tokens, double *scores, void *stream);

/* v2 inject model.predict([features, words]) (text_generation_model_v2.py:140-166; caller
 * evaluate_models/test_score_dense_captions.py:221): words [B, L] int32 pre-padded ids
 * (0 = masked).  probs [B, V] fp32. */
int dc_decoder_v2_predict(DcDecoder *dec, const void *feats, int feats_kind, const int32_t *words,
                          int B, int L, float *probs, void *stream);

/* v2 greedy loop (test_score_dense_captions.py:216-225): start from [0], P-1 predictions.
 *   tokens [B, P-1] int32; probs optional [B, P-1, V]. */
int dc_decoder_v2_greedy(DcDecoder *dec, const void *feats, int feats_kind, int B,
                         int32_t *tokens, float *probs, void *stream);

/* Same loop started from a given first word per RoI (evaluate_models/eval_text_generation_model_v2.py:176-186:
 * prev = [gt_caption[0]]): start [B] int32 ids (device), NULL = all zeros (the loop above).  scores (optional,
 * [B] fp32): the caption score refine_generations ranks by, sum over the P-1 steps of log max p
 * (test_score_dense_captions.py:256-258), without materialising the probabilities. */
int dc_decoder_v2_greedy_from(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *start,
                              int32_t *tokens, float *probs, float *scores, void *stream);

/* Host-buffer forms (what Keras predict callers see): HOST pointers in and out, copies inside. */
int dc_decoder_greedy_host(DcDecoder *dec, const float *feats, int feats_kind, int B,
                           int32_t *tokens, float *probs);

/* ------------------------------------------------------------------------------------------
 * Training step of the v1 model (bf16 decoder only): replaces model.compile(Adam(amsgrad=True),
 * roi_caption_loss) + fit_generator / train_on_batch of build_lstm_model(mode='training'),
 * text_generation_model.py:159-189, 264-294, 424-426, 470-472.
 * ---------------------------------------------------------------------------------------- */

/* Forward (teacher-forced masked scan over gt, equal to the reference's per-prefix evaluation) +
 * roi_caption_loss + backward.  Gradients of every trainable tensor are left in the handle's flat
 * gradient buffer (dc_decoder_grad_buffer), Keras layout, same offsets as the parameters.
 *   feats    per feats_kind; DC_FEATS_HEAD_F32 trains the word model only (head frozen / absent)
 *   gt       [B, P] int32 caption ids as the data generator yields them (1, w.., 2, 0-pad; :111-114)
 *   targets  [B, P] int32 class id per position, or NULL = shift-left(gt) ++ [0] (:352-358);
 *            a negative id marks a position with an all-zero target row (excluded by the loss, :287)
 *   inv_count  1 / (number of positions the loss averages over, GLOBALLY when the batch is sharded
 *            over ranks); <= 0 means 1/(B*P)
 *   loss     device float: sum over this call's positions of -log(clip(p_y,1e-7,1-1e-7)) * inv_count
 * VOCABULARY_SIZE must be a multiple of 8.  recurrent_dropout off (the parity mode); see dc_decoder_train_step_ex. */
int dc_decoder_train_step(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *gt,
                          const int32_t *targets, float inv_count, float *loss, void *stream);

/* Same step with the optional parts of the reference's training graph:
 *   d_feats            NULL, or [B, pool, pool, C] fp32 (device): dL/d(RoI features), the gradient the JOINT model
 *                      (dense_img_cap/dense_model.py:738-755: extract_roi_features -> caption head) sends back into
 *                      PyramidROIAlign -- feed it to dc_pyramid_roi_align_backward_f32.  Needs RoI-feature input
 *                      (DC_FEATS_ROI_F32 / DC_FEATS_ROI_BF16).
 *   recurrent_dropout  rate of KL.LSTM(..., recurrent_dropout=0.2) (text_generation_model.py:141-142): four
 *                      time-invariant masks per LSTM (one per gate i,f,c,o), values {0, 1/(1-rate)}, applied to
 *                      h_{t-1} before its recurrent product.  0 = off.  Masks are counter-based (Philox-4x32-10 keyed
 *                      by dropout_seed, counter = (global row, unit, layer, dropout_step)): no state, and a batch
 *                      sharded over any number of ranks draws the same masks as the unsharded batch when
 *                      row_offset = global index of this call's first row.  TF's own RNG stream cannot be matched;
 *                      the oracle (oracle/decoder.py: philox_masks) draws the identical masks.
 * opts == NULL behaves as dc_decoder_train_step. */
typedef struct DcTrainOptions {
    float *d_feats;
    float recurrent_dropout;
    uint64_t dropout_seed;
    int64_t dropout_step;
    int64_t row_offset;
} DcTrainOptions;
int dc_decoder_train_step_ex(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *gt,
                             const int32_t *targets, float inv_count, float *loss, const DcTrainOptions *opts,
                             void *stream);

/* One training step of the v2 inject model (text_generation_model_v2.py:140-166 build_model(inject=True); :263-267
 * compile(Adam(amsgrad=True), keras.losses.categorical_crossentropy); :312 fit_generator on (prefix, next word) batches).
 *   words    [B, L] int32 pre-padded prefix ids (0 = masked)      targets  [B] int32 next-word class id (< 0: ignored)
 *   inv_count  1 / (GLOBAL batch size); <= 0 means 1/B            loss     device float: sum_b -log(clip(p_y)) * inv_count
 * Gradients of lstm_1, imgcap_lstm and imgcap_d1 land in the flat gradient buffer; the RoI head (trainable=False in
 * the reference) and the embedding are frozen: their slots stay zero.  bf16 v2 handles only. */
int dc_decoder_v2_train_step(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *words, int L,
                             const int32_t *targets, float inv_count, float *loss, void *stream);

/* model.predict([features, gt_captions]) of the TRAINING graph (text_generation_model.py:264-277):
 * teacher-forced word probabilities, probs [B, P, V] fp32 (device). */
int dc_decoder_teacher_forced(DcDecoder *dec, const void *feats, int feats_kind, int B, const int32_t *gt,
                              float *probs, void *stream);

/* keras.optimizers.Adam(lr, beta_1, beta_2, epsilon, amsgrad).get_updates on the flat parameter
 * buffer (text_generation_model.py:425): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m, v, vhat = max(vhat, v);
 * p -= lr_t*m/(sqrt(vhat)+epsilon).  `iteration` t starts at 1.  Gradients are multiplied by
 * grad_scale first (1/world_size after a sum all-reduce).  Re-derives the operand copies. */
int dc_adam_step(DcDecoder *dec, float lr, float beta1, float beta2, float epsilon, int amsgrad,
                 int64_t iteration, float grad_scale, void *stream);

/* Sharded optimiser (one process per GPU, parallel.DataParallelTrainer(shard_optimizer=True); replaces the replicated
 * update behind parallel_model.py:22-102): the same update on the range [offset, offset + numel) of the flat buffers
 * only (both multiples of 4 floats) -- each rank steps the range whose summed gradient it received from the
 * reduce-scatter; the operand copies are NOT re-derived.  After the ranks have all-gathered their updated ranges into
 * dc_decoder_param_buffer, dc_decoder_params_updated refreshes the bf16 mirror and every derived operand copy. */
int dc_adam_step_range(DcDecoder *dec, float lr, float beta1, float beta2, float epsilon, int amsgrad,
                       int64_t iteration, float grad_scale, int64_t offset, int64_t numel, void *stream);
int dc_decoder_params_updated(DcDecoder *dec, void *stream);

/* Flat fp32 device buffers over all TRAINABLE tensors (BatchNorm moving statistics and the frozen
 * embedding excluded): the gradient buffer is what a data-parallel host all-reduces (NCCL) between
 * dc_decoder_train_step and dc_adam_step; dc_decoder_weight_offset gives each tensor's offset
 * (in floats; -1 for frozen tensors). */
int dc_decoder_grad_buffer(DcDecoder *dec, float **ptr, int64_t *numel);
int dc_decoder_param_buffer(DcDecoder *dec, float **ptr, int64_t *numel);
/* The backward pass finishes the gradients in reverse layer order; bucket `index` in [0,4) = {vocabulary
 * projection, dense1, both LSTMs, RoI head} is one contiguous [offset, offset+numel) range of the
 * gradient buffer.  dc_decoder_wait_grad_bucket makes `waiting_stream` wait (cudaStreamWaitEvent) until
 * the LAST dc_decoder_train_step has produced that bucket, so a communication stream can all-reduce
 * bucket i while the step's stream is still computing bucket i+1. */
#define DC_GRAD_BUCKETS 4
int dc_decoder_grad_bucket(DcDecoder *dec, int index, int64_t *offset, int64_t *numel);
int dc_decoder_wait_grad_bucket(DcDecoder *dec, int index, void *waiting_stream);
int64_t dc_decoder_weight_offset(const DcDecoder *dec, int index);
int dc_decoder_get_grad(DcDecoder *dec, const char *name, float *host, int64_t numel);

/* ------------------------------------------------------------------------------------------
 * End-to-end: PyramidROIAlign -> RoI head -> greedy decoding in one call
 * ---------------------------------------------------------------------------------------- */

/* generate_features(...) followed by model.predict(features) as the reference's evaluation loop
 * runs them per image (evaluate_models/eval_text_generation_model.py:138-148): boxes
 * [n_images, n_boxes, 4] normalised, 4 NHWC fp32 maps (channels = the decoder's `channels`,
 * pool = its `pool`), tokens [n_images*n_boxes, P] int32.  The RoI features stay in HBM (bf16
 * for a bf16 decoder).  Device pointers; asynchronous on `stream`. */
int dc_caption_rois(DcDecoder *dec, const float *boxes, const float *const fmaps[4],
                    const int fm_h[4], const int fm_w[4], int n_images, int n_boxes, int img_h,
                    int img_w, int32_t *tokens, void *stream);

/* Host-buffer form: HOST pointers (pinned memory makes the copies asynchronous); the pyramid is
 * uploaded image by image while the previous image is aligned and decoded; returns when `tokens`
 * is complete.  Streams, events and device staging are kept in the handle and reused by every call.  No caller
 * stream: work enqueued on the handle from another stream (dc_adam_step ...) must have completed. */
int dc_caption_rois_host(DcDecoder *dec, const float *boxes, const float *const fmaps[4],
                         const int fm_h[4], const int fm_w[4], int n_images, int n_boxes, int img_h,
                         int img_w, int32_t *tokens);
/* The same pipeline split in two so that consecutive calls overlap: submit() enqueues the uploads, kernels and the
 * token download of one call and returns; wait() blocks until the OLDEST outstanding call's tokens are on the
 * host.  At most two calls are outstanding (a third submit first retires the oldest); host buffers of a submitted
 * call must stay valid and unmodified until its wait() returns.  With two in flight, call k+1's upload runs under
 * call k's last-image decode.  `boxes` and `tokens` may be pageable (numpy arrays): boxes are copied into pinned
 * staging inside submit(), token ids land in pinned staging and wait() copies them into `tokens` -- a device -> host
 * copy straight into pageable memory would block submit() until the whole call had run.  The feature maps are not
 * staged (89 MB per image): pin them, or the uploads serialise with the host. */
int dc_caption_rois_host_submit(DcDecoder *dec, const float *boxes, const float *const fmaps[4],
                         const int fm_h[4], const int fm_w[4], int n_images, int n_boxes, int img_h,
                         int img_w, int32_t *tokens);
int dc_caption_rois_host_wait(DcDecoder *dec);


/* ------------------------------------------------------------------------------------------
 * Caption post-processing (the step after the decoder in the dense-captioning evaluation path)
 * ---------------------------------------------------------------------------------------- */

/* refine_generations (evaluate_models/test_score_dense_captions.py:245-283) with non_max_suppression
 * (evaluate_models/utils.py:69-104) and that copy's overlap measure 2*I/(A+B) (utils.py:30-48), per image:
 * RoIs are visited in descending caption score (equal scores: larger index first), a visited RoI suppresses
 * every later one whose overlap exceeds nms_threshold, and the first max_keep survivors are returned.
 *   boxes [n_images, n_boxes, 4] fp32 (y1,x1,y2,x2), scores [n_images, n_boxes] fp32 (device)
 *   keep [n_images, max_keep] int32 indices into the image's boxes (-1 padded), n_keep [n_images] int32 */
int dc_refine_generations(const float *boxes, const float *scores, int n_images, int n_boxes,
                          float nms_threshold, int max_keep, int32_t *keep, int32_t *n_keep, void *stream);

/* ------------------------------------------------------------------------------------------
 * Box front-end: what produces the boxes PyramidROIAlign consumes.
 * ---------------------------------------------------------------------------------------- */

/* Replaces ProposalLayer.call (modified_dense_model.py:247-303): per image, the pre_nms_limit best anchors
 * by foreground score (tf.nn.top_k: descending, lower index first among equals), deltas * bbox_std_dev applied
 * to them (apply_box_deltas_graph :179-200), clipped to [0,h]x[0,w] (clip_boxes_graph :203-218), divided by
 * [h,w,h,w], greedy NMS with tf.image.non_max_suppression's IoU rule (suppress when IoU > nms_threshold),
 * at most proposal_count survivors in score order, zero padded.
 *   rpn_probs [n_images, n_anchors, 2] fp32 (bg, fg), rpn_bbox [n_images, n_anchors, 4] fp32,
 *   anchors [n_anchors, 4] fp32 pixels (y1,x1,y2,x2), bbox_std_dev: 4 HOST floats (config.RPN_BBOX_STD_DEV)
 *   proposals [n_images, proposal_count, 4] fp32; optional n_valid [n_images] int32 and
 *   anchor_index [n_images, proposal_count] int32 (-1 padded; which anchor each proposal came from)
 *   workspace: device scratch of at least dc_proposal_workspace_bytes(...) bytes, 16-byte aligned.
 * pre_nms_limit (the reference's constant 6000) may not exceed 8192. */
size_t dc_proposal_workspace_bytes(int n_images, int n_anchors, int pre_nms_limit, int proposal_count);
int dc_proposal_layer(const float *rpn_probs, const float *rpn_bbox, const float *anchors, int n_images,
                      int n_anchors, const float *bbox_std_dev, float image_h, float image_w, int pre_nms_limit,
                      int proposal_count, float nms_threshold, float *proposals, int32_t *n_valid,
                      int32_t *anchor_index, void *workspace, size_t workspace_bytes, void *stream);

/* Replaces the Lambda `x / image_scale` (modified_dense_model.py:1523-1526): pixel boxes (y1,x1,y2,x2) ->
 * normalised coordinates, fp32 division by [h,w,h,w].  boxes / out: [n_boxes, 4] fp32 device (may alias). */
int dc_normalize_boxes(const float *boxes, int64_t n_boxes, float image_h, float image_w, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * Dense contraction primitives (exported so that tests can pin the GEMM kernels in isolation;
 * the decoder calls the same code).  They replace the MatMul ops behind KL.Dense / KL.LSTM /
 * KL.Conv2D(valid, full-window) on this path (text_generation_model.py:141-154, 251-262).
 * ---------------------------------------------------------------------------------------- */

/* fp32 FFMA GEMM: C[M,N] = relu?(op(A)*op(B) + bias[n] + addend[m,n]); op = identity or transpose
 * (trans_a: A stored [K,M]; trans_b: B stored [N,K]); accumulate != 0 adds to the existing C. */
int dc_gemm_f32(const float *A, int64_t lda, int trans_a, const float *B, int64_t ldb, int trans_b,
                int M, int N, int K, const float *bias, const float *addend, int64_t ld_addend,
                int relu, int accumulate, float *C, int64_t ldc, void *stream);

/* bf16 tcgen05 GEMM: D[M,N] = relu?(A[M,K] * Bt[N,K]^T + bias[n] + addend[m,n]) with fp32
 * accumulation in TMEM.  A and Bt are bf16 bit patterns, K contiguous, lda/ldb multiples of 8.
 * Writes fp32 and/or bf16 outputs (either may be NULL, not both). */
int dc_gemm_bf16(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M, int N, int K,
                 const float *bias, const float *addend, int64_t ld_addend, int relu,
                 float *out_f32, int64_t ld_f32, uint16_t *out_bf16, int64_t ld_bf16, void *stream);

/* General form used by the training step (all of dc_gemm_bf16 plus):
 *   a_mn / b_mn   operand is MN-major: stored [K, rows] row-major (rows contiguous), ld = stride
 *                 between consecutive k.  Weight gradients X^T*dY read X [R,in] and dY [R,out] as they
 *                 lie in memory (both MN-major, K = R) -- nothing is transposed in HBM.
 *   atomic        fp32 output is ACCUMULATED (red.global.add); with split_k != 1 the K range is
 *                 split over CTAs (0 = automatic).  Output must be zeroed / hold the running sum.
 *   addend_mod    > 0: addend row index is m % addend_mod (per-RoI term broadcast over time steps)
 *   deint_units   > 0: N == 4*units gate-interleaved columns (4u+g) are written to the Keras
 *                 block layout (g*units+u)
 *   mask_src      bf16 [M, ld_mask]: result is zeroed where mask_src <= 0 (ReLU backward)        */
int dc_gemm_bf16_ex(const uint16_t *A, int64_t lda, int a_mn, const uint16_t *B, int64_t ldb, int b_mn,
                    int M, int N, int K, const float *bias, const float *addend, int64_t ld_addend,
                    int addend_mod, int relu, const uint16_t *mask_src, int64_t ld_mask, int deint_units,
                    int atomic, int split_k, float *out_f32, int64_t ld_f32, uint16_t *out_bf16,
                    int64_t ld_bf16, void *stream);

/* bf16 tcgen05 GEMM with the fused arg-max epilogue (Dense(V) + softmax + tf.argmax of the greedy
 * loop, text_generation_model.py:144,222-225): tokens[m] = first arg-max over n of
 * (A*Bt^T + bias)[m,n]; maxprob[m] (optional) = its softmax probability.  The [M,N] logits are
 * never written. */
int dc_gemm_bf16_argmax(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M, int N,
                        int K, const float *bias, int32_t *tokens, float *maxprob, void *stream);

/* bf16 tcgen05 GEMM with the fused top-k epilogue (Dense(V) + softmax + np.argsort(p)[-k:] of the beam
 * step, image captioning/test.py:41-48): per row the k largest softmax probabilities of
 * (A*Bt^T + bias) in ASCENDING order, ties ranked like a stable ascending argsort (the larger index is
 * the better candidate).  idx [M,k] int32, prob [M,k] fp32; 1 <= k <= 8.  The [M,N] logits are never written. */
int dc_gemm_bf16_topk(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M, int N, int K,
                      const float *bias, int k, int32_t *idx, float *prob, void *stream);

/* bf16 tcgen05 GEMM with the fused Keras LSTM cell epilogue (KL.LSTM step,
 * text_generation_model.py:141-142): z = A*Bt^T + addend + bias with GATE-INTERLEAVED columns
 * (column 4*u+g holds gate g in {i,f,c,o} of unit u, i.e. Bt row 4*u+g is column g*units+u of the
 * Keras kernel); i,f,o = hard_sigmoid, g = tanh; c' = f*c + i*g; h' = o*tanh(c').  Rows whose
 * tok[m] == 0 carry (h, c) (K.rnn mask).  c [M,units] fp32 is updated in place; h' is written as
 * bf16 to h_out (and h_out2 if not NULL); h_prev supplies the carried h of masked rows. */
int dc_gemm_bf16_lstm_cell(const uint16_t *A, int64_t lda, const uint16_t *Bt, int64_t ldb, int M,
                           int units, int K, const float *addend, int64_t ld_addend,
                           const float *bias, const int32_t *tok, float *c,
                           const uint16_t *h_prev, int64_t ld_h_prev, uint16_t *h_out,
                           int64_t ld_h_out, uint16_t *h_out2, int64_t ld_h_out2, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DCAP_H_ */
