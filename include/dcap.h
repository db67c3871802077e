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
 * Host-buffer form of dc_pyramid_roi_align_f32 (what `generate_features` /
 * `keras_model.predict` callers see, evaluate_models/generate_one_roi_features.py:69-76):
 * all pointers are HOST memory (pinned memory makes the copies asynchronous); the call stages the
 * pyramid image by image so copies overlap the kernel, and returns when `out` is complete.
 */
int dc_pyramid_roi_align_host_f32(const float *boxes, const float *const fmaps[4],
                                  const int fm_h[4], const int fm_w[4], int n_images, int n_boxes,
                                  int channels, int pool_h, int pool_w, int img_h, int img_w,
                                  float *out, int32_t *levels);

#ifdef __cplusplus
}
#endif
#endif /* DCAP_H_ */
