/* disyolo.h -- C ABI of libdisyolo_b200.so: the DIS-YOLO hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (ZHANGKEON/DIS-YOLO) has no FFI: its hot path is a Python class that builds a
 * TensorFlow-1.x graph (yolo/yolo3_net_pos.py) driven through tf.Session.run.  This header is the
 * boundary a maintainer binds with ctypes instead; each entry point cites the
 * reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns 0 on success, a negative dy_status otherwise; dy_last_error() returns
 *     a thread-local human readable message.  Nothing aborts, nothing throws across this ABI.
 *   - "dev" pointers are CUDA device pointers on the net's device, "host" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all work is enqueued on
 *     it asynchronously; the caller owns inputs/outputs, the library owns weights and workspaces.
 *   - one dy_net per GPU; calls on one net must be serialised by the caller.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef DISYOLO_H_
#define DISYOLO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dy_net dy_net;

enum dy_status {
  DY_STATUS_OK = 0,
  DY_STATUS_INVALID = -1,
  DY_STATUS_CUDA = -2,
  DY_STATUS_STATE = -3,
  DY_STATUS_NOTFOUND = -4,
  DY_STATUS_UNSUPPORTED = -5
};

enum dy_precision {
  DY_PRECISION_BF16 = 0, /* tcgen05 bf16 x bf16 -> fp32 accumulate, bf16 activations        */
  DY_PRECISION_FP32 = 1  /* verification mode: fp32 CUDA-core convolutions, fp32 activations */
};

/* What dy_forward_host_begin* brings back besides boxes and counts. */
enum dy_mask_mode {
  DY_MASKS_NONE = 0,    /* boxes / counts only                                                            */
  DY_MASKS_FULL = 1,    /* the reference's layout: det_count[b] maps of [S/2,S/2] fp32 per image          */
  DY_MASKS_CROPPED = 2  /* only what the reference's consumer reads: det_mask[y1:y2, x1:x2] of every
                           detection, (y1,x1,y2,x2) = round(box * S/2) (calculate_test_map.py:247-252),
                           packed back to back -- same floats, a fraction of the PCIe traffic          */
};

/* Constructor arguments = the constants YOLONet.__init__ reads from yolo/config.py
 * (yolo3_net_pos.py:13-37; config.py:21-22,38,41-46,60-72) plus the per-layer lock flags that are
 * literals inside build_network (yolo3_net_pos.py:155-156). */
typedef struct dy_config {
  int32_t num_classes;     /* len(cfg.CLASSES) = 3                                  */
  float anchors[18];       /* cfg.ANCHORS, 9 x (w,h) in pixels, smallest first      */
  int32_t image_size;      /* cfg.IMAGE_SIZE / TEST_SIZE, multiple of 32            */
  int32_t k_map;           /* cfg.K_MAP (3, 5 or 7)                                 */
  float alpha;             /* cfg.ALPHA leaky slope                                 */
  float bn_eps;            /* 1e-5  (yolo3_net_pos.py:75)                           */
  float iou_threshold;     /* cfg.IOU_THRESHOLD                                     */
  int32_t max_detection;   /* cfg.MAX_DETECTION                                     */
  int32_t max_batch;       /* cfg.BATCH_SIZE upper bound; buffers are sized for it  */
  int32_t precision;       /* enum dy_precision                                     */
  int32_t device;          /* CUDA device ordinal                                   */
  uint8_t lock[82];        /* lock flag of convolutional1..82 (1 = frozen)          */
  uint8_t reserved[2];
} dy_config;

/* Library / build information: "disyolo_b200 <version> sm_100a". */
const char* dy_version(void);
/* Message of the last failing call on this thread. */
const char* dy_last_error(void);
/* Number of CUDA devices visible (0 when there is no driver / GPU). */
int dy_device_count(void);

/* Replaces: YOLONet(training) graph construction (yolo3_net_pos.py:13-65, 153-463).
 * Allocates every activation buffer / workspace for max_batch images. */
int dy_create(const dy_config* cfg, dy_net** out);
int dy_destroy(dy_net* net);

/* Replaces: tf.train.Saver.restore / slim.assign_from_checkpoint_fn by variable name
 * (train_yolo3_mask.py:85-107, calculate_test_map.py:184-185).  `tf_name` is the reference's
 * variable name, e.g. "yolo/convolutional7/weights" (HWIO fp32), ".../BatchNorm/gamma|beta|
 * moving_mean|moving_variance", ".../biases".  `host` holds prod(shape) floats. */
int dy_load_weights(dy_net* net, const char* tf_name, const float* host, const int64_t* shape, int32_t ndim);
/* Folds BN, packs bf16 operands, builds TMA descriptors.  Must follow the last dy_load_weights and
 * precede dy_forward.  After dy_train_init the device master copies are the truth: variables loaded since
 * the previous finalize replace their masters (and restart Adam, whose slots the reference's Saver does
 * not store either), all others are read back from the masters, so a re-finalize never reverts training. */
int dy_finalize_weights(dy_net* net);

/* Replaces: sess.run(net.evaluation, {images, clip_window, det_thresh, is_training: False})
 * (calculate_test_map.py:214-218, train_yolo3_mask.py:171-174; graph: yolo3_net_pos.py:153-463,
 * 465-628, 862-938).
 *   images_dev   [B,S,S,3] fp32 NHWC RGB/255          windows_dev [B,4] fp32 (y1,x1,y2,x2)
 *   det_raw_dev  [B,max_det,6] fp32  filter_detections output, zero padded (may be NULL)
 *   det_box_dev  [B,max_det,6] fp32  rows kept by val_test's w>0,h>0 filter, zero padded
 *   det_count_dev[B] int32           number of valid rows of det_box
 *   masks_dev    [B,max_det,S/2,S/2] fp32; only the first det_count[b] maps of image b are written
 */
int dy_forward(dy_net* net, const float* images_dev, int32_t B, const float* windows_dev, float det_thresh,
               float* det_raw_dev, float* det_box_dev, int32_t* det_count_dev, float* masks_dev, void* stream);

/* The same call with HOST buffers: pinned staging, H2D of images/windows, the forward, D2H of the
 * boxes, the counts and exactly det_count[b] masks per image; synchronises before returning.
 * This is the call the Python facade's Session.run makes. */
int dy_forward_host(dy_net* net, const float* images_host, int32_t B, const float* windows_host, float det_thresh,
                    float* det_raw_host, float* det_box_host, int32_t* det_count_host, float* masks_host);

/* The pipelined form of dy_forward_host for back-to-back batches: _begin enqueues H2D + forward
 * for one batch and returns a ticket immediately; _end performs the D2H of that batch and blocks
 * until the host buffers are filled.  Up to three tickets may be in flight (three streams, three
 * device slots), so
 *     t0 = begin(b0); t1 = begin(b1); loop { t2 = begin(next); end(t0); t0 = t1; t1 = t2; }
 * overlaps the H2D of batch k+2 and the D2H of batch k with the convolutions of batch k+1 without a
 * bubble (with two slots the next H2D could only start after the previous D2H had finished).
 * dy_forward_host == begin + end. */
int dy_forward_host_begin(dy_net* net, const float* images_host, int32_t B, const float* windows_host,
                          float det_thresh, int32_t mask_mode /* enum dy_mask_mode */, int32_t* ticket);
int dy_forward_host_end(dy_net* net, int32_t ticket, float* det_raw_host, float* det_box_host,
                        int32_t* det_count_host, float* masks_host);
/* The same with the image as image_read holds it BEFORE its `/ 255.` (calculate_test_map.py:160-175):
 * [B,S,S,3] uint8 RGB, letterboxed.  The division happens on the device and yields bit-identical fp32
 * inputs ((float)(v / 255.0), 256 possible values), with a quarter of the host->device bytes. */
int dy_forward_host_begin_u8(dy_net* net, const uint8_t* images_host, int32_t B, const float* windows_host,
                             float det_thresh, int32_t mask_mode, int32_t* ticket);
/* Completion of a DY_MASKS_CROPPED ticket.  crop_offsets_host [B*max_det + 1] int64: crop (b,d) starts at
 * float offset crop_offsets_host[b*max_det + d] of crops_host and is (y2-y1) x (x2-x1) row-major with
 * (y1,x1,y2,x2) = round(det_box[b,d,0:4] * S/2) clamped to the map; the last entry is the total.  All results
 * of the ticket leave the device in two copies (small block, crops).  crops_capacity in floats;
 * B*max_det*(S/2)^2 always suffices. */
int dy_forward_host_end_cropped(dy_net* net, int32_t ticket, float* det_raw_host, float* det_box_host,
                                int32_t* det_count_host, int64_t* crop_offsets_host, float* crops_host,
                                int64_t crops_capacity);

/* Network only (conv1..82), no decode / NMS / masks: fills the head and score-map buffers and materialises
 * EVERY layer's output (the parity-tap path; dy_forward fuses convolutional82 into convolutional81's
 * epilogue on the bf16 engine and never writes convolutional81's activation). */
int dy_forward_network(dy_net* net, const float* images_dev, int32_t B, void* stream);

/* Measurement aid for bench.py: the same launches as dy_forward_network with a CUDA event between
 * consecutive layers; layer_ms_host[1..82] receives each layer's device time in milliseconds
 * ([0] = 0).  Synchronises the stream.  bf16 engine only. */
int dy_forward_profile(dy_net* net, const float* images_dev, int32_t B, float* layer_ms_host, void* stream);

/* Parity taps (replace sess.run on intermediate tensors of build_network).
 * dy_get_activation: output of convolutional<layer> (after BN/leaky/residual) as fp32 NHWC
 *   [B,h,w,c]; dims are reported by dy_layer_shape.
 * dy_get_yolo: scale 0/1/2 = stride 8/16/32 map [B,g,g,3*(5+C)] fp32 (yolo3_net_pos.py:353).
 * dy_get_mask_pos: [B,S/2,S/2,k*k] fp32 NHWC (yolo3_net_pos.py:410-412). */
int dy_layer_shape(dy_net* net, int32_t layer, int32_t* h, int32_t* w, int32_t* c);
int dy_get_activation(dy_net* net, int32_t layer, int32_t B, float* out_dev, void* stream);
int dy_get_yolo(dy_net* net, int32_t scale, int32_t B, float* out_dev, void* stream);
int dy_get_mask_pos(dy_net* net, int32_t B, float* out_dev, void* stream);

/* Stand-alone stages, fed identical inputs in the parity tests.
 * dy_decode   replaces interpret_output + the candidate part of filter_detections
 *             (yolo3_net_pos.py:465-514, 523-561).  yolo8/16/32_dev: [B,g,g,3*(5+C)] fp32.
 *             Outputs (dense, candidate order of :527-542): box [B,N0,4] (y1,x1,y2,x2 clipped),
 *             cls [B,N0] int32, score [B,N0] fp32.
 * dy_detect   replaces filter_detections + val_test's box filter (:517-628, 876-880) from head
 *             maps: decode -> threshold -> per-class NMS -> top-k.
 * dy_nms      the same selection from caller-supplied boxes/classes/scores [B,N,...] (identical
 *             decoded boxes in -> keep set must be bit-exact); sel_idx [B,max_det] int32 receives
 *             the selected candidate indices in output order (-1 padded), sel_count [B].
 * dy_assemble_masks replaces val_test's mask assembly (:881-933).  score maps either NHWC
 *             [B,S,S,k*k] (layout=0, the reference's) or planar [B,k*k,S,S] (layout=1);
 *             det_box [B,max_det,6] / det_count [B] as produced by dy_forward. */
int dy_decode(dy_net* net, const float* yolo8_dev, const float* yolo16_dev, const float* yolo32_dev, int32_t B,
              const float* windows_dev, float* box_dev, int32_t* cls_dev, float* score_dev, void* stream);
int dy_detect(dy_net* net, const float* yolo8_dev, const float* yolo16_dev, const float* yolo32_dev, int32_t B,
              const float* windows_dev, float det_thresh, float* det_raw_dev, float* det_box_dev,
              int32_t* det_count_dev, void* stream);
int dy_nms(dy_net* net, const float* box_dev, const int32_t* cls_dev, const float* score_dev, int32_t B, int32_t N,
           float det_thresh, int32_t* sel_idx_dev, int32_t* sel_count_dev, float* det_raw_dev, void* stream);
int dy_assemble_masks(dy_net* net, const float* score_dev, int32_t layout, int32_t B, const float* det_box_dev,
                      const int32_t* det_count_dev, float* masks_dev, void* stream);

/* ---- either side of the hot path ----------------------------
 * Replaces image_read (calculate_test_map.py:149-176, utils/val_data.py:36-63): the RGB uint8 image
 * [h,w,3] (device) is resized with cv2.INTER_LINEAR semantics to fit image_size, embedded in a
 * 127-filled square and divided by 255 -> out_dev [image_size,image_size,3] fp32; window_host[4]
 * (may be NULL) receives the clip window (top, left, bottom, right) / image_size. */
int dy_letterbox(const uint8_t* rgb_dev, int32_t h, int32_t w, int32_t image_size, float* out_dev, float* window_host,
                 void* stream);
/* The same for B images of one shape in ONE launch: image b starts at rgb_dev + b*image_stride (bytes),
 * out_dev [B,image_size,image_size,3], windows_host [B,4] (may be NULL). */
int dy_letterbox_batch(const uint8_t* rgb_dev, int64_t image_stride, int32_t B, int32_t h, int32_t w,
                       int32_t image_size, float* out_dev, float* windows_host, void* stream);
/* Replaces the per-detection loop of calculate_test_map.py:233-269 (utils/validation_map.py:137-166)
 * for ONE image: correct_yolo_boxes (:121-138), crop of each [S,S] mask to its box, cv2.resize
 * INTER_LINEAR to the box size in the original image, > 0.5, paste.  det_box_dev [max_det,6] and
 * det_count_dev [1] and masks_dev [max_det,S,S] are one image's slice of dy_forward's outputs.
 * boxes_out_dev [max_det,4] int32 = (x1,y1,x2,y2) in original pixels; valid_out_dev [max_det] = 0 for
 * detections the reference skips ((y2-y1)*(x2-x1) <= 0); full_masks_dev [max_det,image_h,image_w]
 * bool bytes (may be NULL; planes at or beyond det_count are left unwritten); merged_dev [image_h,image_w] = class+1 of the last detection covering
 * each pixel, 0 elsewhere (may be NULL). */
int dy_postprocess(const float* det_box_dev, const int32_t* det_count_dev, int32_t max_det, const float* masks_dev,
                   int32_t S, int32_t image_h, int32_t image_w, int32_t net_size, int32_t* boxes_out_dev,
                   uint8_t* valid_out_dev, uint8_t* full_masks_dev, uint8_t* merged_dev, void* stream);
/* The same for B images that share one original shape, in ONE launch pair: every array is the [B, ...] stack of
 * the one-image form (det_box [B,max_det,6], det_count [B], masks [B,max_det,S,S] = dy_forward's outputs as they
 * stand; boxes_out [B,max_det,4], valid_out [B,max_det], full_masks [B,max_det,h,w], merged [B,h,w]). */
int dy_postprocess_batch(const float* det_box_dev, const int32_t* det_count_dev, int32_t B, int32_t max_det,
                         const float* masks_dev, int32_t S, int32_t image_h, int32_t image_w, int32_t net_size,
                         int32_t* boxes_out_dev, uint8_t* valid_out_dev, uint8_t* full_masks_dev, uint8_t* merged_dev,
                         void* stream);

/* Replaces compute_overlaps_masks (utils/voc_eval_mask.py:38-56), the inner operation of the mask-level
 * mAP (voc_eval, :58-134): IoU of every mask of set 1 with every mask of set 2.
 * Masks are instance-major [n, pixels] bytes (non-zero = inside; dy_postprocess's full_masks layout);
 * overlaps_dev [n1, n2] fp32 = inter / (area1 + area2 - inter), NaN for two empty masks like NumPy. */
int dy_mask_overlaps(const uint8_t* masks1_dev, int32_t n1, const uint8_t* masks2_dev, int32_t n2, int64_t pixels,
                     float* overlaps_dev, void* stream);

/* Replaces the label assignment of defect_train.get (utils/train_data.py:134-178; flips :189-228;
 * normalisation :258-262): ground-truth boxes (x1,y1,x2,y2,class) in original image
 * pixels [B,max_box,5] with nbox_dev[b] valid rows, the placement place_dev [B,4] = (sx, sy, dx, dy) of the
 * (augmented) image inside the net square and flip_dev [B] (1 none / 2 horizontal / 3 vertical, may be
 * NULL) -> the feed tensors of the training step: yolo3/yolo2/yolo1 [B,g,g,3,5+C] (best-IoU anchor, first
 * box wins a cell, coordinates / image_size) and true_boxes [B,max_box,5].  Outputs are zeroed first. */
int dy_assign_labels(dy_net* net, const float* boxes_dev, const int32_t* nbox_dev, const float* place_dev,
                     const int32_t* flip_dev, int32_t B, int32_t max_box, float* yolo3_dev, float* yolo2_dev,
                     float* yolo1_dev, float* true_boxes_dev, void* stream);

/* ---- the training data pipeline on the device (utils/train_data.py) -----------------
 * Every random draw of the reference (np.random in get(), add_salt_pepper_noise, change_light,
 * linearmotion_blur3C) is an ARGUMENT: the host draws, the device executes.
 * dy_polygon_masks  load_mask (:321-338): the VIA polygons of one image -> one byte mask per instance.
 *     verts [nv,2] float64 (x, y); poly [np,3] int32 = (first vertex, vertex count, type: 1 = 'out', 0 = an inner
 *     background region); inst [n_inst+1] int32 = polygon range of each instance.  A polygon fills the pixels
 *     skimage.draw.polygon returns (interior + boundary, O'Rourke's test) with 1 ('out') or 0, then sets its
 *     vertices to 1, in annotation order.
 * dy_mask_boxes     extract_bboxes (:358-374): (x1, y1, x2, y2) int32 per mask, x2 / y2 one past the last set pixel;
 *     (0,0,0,0) for an empty mask (load_box skips those).
 * dy_augment_image  apply_random_scale_and_crop (:423-450, mode 'image') + the flip of image_read (:388-393):
 *     cv2.resize(uint8, (new_w,new_h), INTER_LINEAR) bit for bit, placed at (dx,dy) in a 127-filled
 *     image_size square (negative offsets crop), flip 1 none / 2 columns reversed / 3 rows reversed.
 * dy_augment_masks  the same for n byte masks as resize_mask does it (:403-421): float32 bilinear, pad 0,
 *     flip, np.around -> bool.
 * dy_salt_pepper    add_salt_pepper_noise (:494-509): all channels of the (row, col) pairs set to 1, then to 0.
 * dy_change_light   change_light (:511-521): cv2 RGB->HLS, L = min(L*coeff, 255) truncated, HLS->RGB; in place.
 * dy_motion_blur3   linearmotion_blur3C (:452-481) for the 3x3 line kernel pyblur builds (9 floats, host):
 *     per channel scipy.signal.convolve2d(float32, kernel, 'same', fillvalue=255).astype(uint8).
 * dy_u8_to_unit_float  image.astype(np.float32) / 255.0 (:399-400). */
int dy_polygon_masks(const double* verts_dev, const int32_t* poly_dev, const int32_t* inst_dev, int32_t n_inst,
                     int32_t h, int32_t w, uint8_t* masks_dev, void* stream);
int dy_mask_boxes(const uint8_t* masks_dev, int32_t n, int32_t h, int32_t w, int32_t* boxes_dev, void* stream);
int dy_augment_image(const uint8_t* rgb_dev, int32_t h, int32_t w, int32_t image_size, int32_t new_w, int32_t new_h,
                     int32_t dx, int32_t dy, int32_t flip, uint8_t* out_dev, void* stream);
int dy_augment_masks(const uint8_t* masks_dev, int32_t n, int32_t h, int32_t w, int32_t image_size, int32_t new_w,
                     int32_t new_h, int32_t dx, int32_t dy, int32_t flip, uint8_t* out_dev, void* stream);
int dy_salt_pepper(uint8_t* img_dev, int32_t image_size, const int32_t* salt_rc_dev, int32_t n_salt,
                   const int32_t* pepper_rc_dev, int32_t n_pepper, void* stream);
int dy_change_light(uint8_t* img_dev, int64_t npix, double coeff, void* stream);
int dy_motion_blur3(const uint8_t* img_dev, int32_t image_size, const float* kernel9_host, uint8_t* out_dev, void* stream);
int dy_u8_to_unit_float(const uint8_t* src_dev, float* dst_dev, int64_t n, void* stream);

/* Measurement aid for bench.py: device milliseconds of each post-processing kernel (decode+threshold,
 * per-class NMS, top-k/finalize, mask assembly), each launched `reps` times back to back between two
 * CUDA events on `stream`; ms_host[4] receives the per-launch averages.  Same inputs as dy_detect +
 * dy_assemble_masks (score maps: layout 0 = NHWC, 1 = planar). */
int dy_postproc_profile(dy_net* net, const float* yolo8_dev, const float* yolo16_dev, const float* yolo32_dev,
                        const float* score_dev, int32_t layout, int32_t B, const float* windows_dev, float det_thresh,
                        float* masks_dev, int32_t reps, float* ms_host, void* stream);

/* One convolution of the reference's builders through the same engine the network uses
 * (conv / conv_bn / res_conv_bn, yolo3_net_pos.py:109-151): x [B,H,W,cin] fp32 NHWC (device),
 * w HWIO [k,k,cin,cout] fp32 (host), per-channel scale/shift (host; folded BN or 1/bias),
 * act = leaky flag, residual [B,Ho,Wo,cout] fp32 NHWC (device) or NULL, out [B,Ho,Wo,cout] fp32.
 * stride in {1,2} with TensorFlow 'SAME' padding.  precision selects the tcgen05 bf16 engine or
 * the fp32 verification kernel.  Used by the per-layer parity tests. */
int dy_conv_layer(int32_t precision, const float* x_dev, int32_t B, int32_t H, int32_t W, int32_t cin,
                  const float* w_host, int32_t k, int32_t stride, int32_t cout, const float* scale_host,
                  const float* shift_host, int32_t act, float alpha, const float* residual_dev, float* out_dev,
                  void* stream);

/* Backward of one convolution through the tensor-core training engine (what TF's autodiff emits for
 * tf.nn.conv2d, yolo3_net_pos.py:125,142: Conv2DBackpropInput / Conv2DBackpropFilter), stride 1 or 2 with
 * TensorFlow 'SAME' padding (stride 2: 3x3 kernel, even extents -- the five down-sampling convs):
 * x [B,H,W,cin], dz [B,H/stride,W/stride,cout] fp32 NHWC (device), w HWIO (host) ->
 * dx [B,H,W,cin] fp32 NHWC (device, may be NULL), dw [k,k,cin,cout] fp32 (device, may be NULL).
 * bf16 operands, fp32 accumulation.  Used by the per-layer gradient parity tests. */
int dy_conv_backward(const float* x_dev, const float* dz_dev, int32_t B, int32_t H, int32_t W, int32_t cin,
                     const float* w_host, int32_t k, int32_t stride, int32_t cout, float* dx_dev, float* dw_dev,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training step.  Replaces sess.run([net.total_loss, optimizer], feed_dict) of
 * train_yolo3_mask.py:146-149,216 = training-mode forward (batch-statistics BatchNorm for the
 * unlocked layers, yolo3_net_pos.py:88-98) + loss_yolo (:631-747) + loss_mask (:750-860) + L2
 * (:38) + backward + tf.train.AdamOptimizer (train_yolo3_mask.py:55).  fp32 engine.
 *   yolo3/2/1_dev   label maps [B,g,g,3,8] for stride 8/16/32 (net.yolo3/2/1, :52-55)
 *   true_boxes_dev  [B,20,5] (xc,yc,w,h,class) normalised (net.true_boxes, :56)
 *   true_masks_dev  [B,20,S,S] bool bytes (net.true_masks, :57)
 *   perm_prop_dev   [B,max_det] int32, perm_gt_dev [B,20] int32: the permutations that stand in
 *                   for the reference's unseeded tf.random_shuffle (:781-782)
 *   losses_host[8]  total (incl. L2), object, noobject, class, xy, wh, mask, l2
 * dy_train_backward runs the backward pass of layers layer_hi..layer_lo (descending) and writes
 * their parameter gradients into grad_flat_dev (layout: dy_train_layer_span; per layer weights
 * HWIO, then gamma, beta or biases).  Calling it in descending ranges lets the caller all-reduce
 * finished gradient buckets on another stream while earlier layers are still in backward.
 * dy_train_apply: g*grad_scale (+L2) -> Adam -> parameters; updates the BN moving averages
 * (decay 0.997, :74,92-95) and refreshes the inference-mode folded weights. */
int dy_train_init(dy_net* net);
/* The loss constants YOLONet.__init__ reads from yolo/config.py:49-57 (OBJECT_SCALE, NOOBJECT_SCALE,
 * CLASS_SCALE, COORD_SCALE, MASK_SCALE, IGNORE_THRESH; yolo3_net_pos.py:30-35).  Defaults = the reference's
 * (2, 1, 1, 1, 5, 0.5).  MAX_BOX_PER_IMAGE is fixed at 20 (the shape of true_boxes / true_masks). */
int dy_set_loss_params(dy_net* net, float object_scale, float noobject_scale, float class_scale, float coord_scale,
                       float mask_scale, float ignore_thresh);
int64_t dy_train_param_count(dy_net* net);
int dy_train_layer_span(dy_net* net, int32_t layer, int64_t* offset, int64_t* count);
int dy_train_forward(dy_net* net, const float* images_dev, int32_t B, const float* yolo3_dev, const float* yolo2_dev,
                     const float* yolo1_dev, const float* true_boxes_dev, const uint8_t* true_masks_dev,
                     const int32_t* perm_prop_dev, const int32_t* perm_gt_dev, float det_thresh, float* losses_host,
                     void* stream);
int dy_train_backward(dy_net* net, int32_t B, int32_t layer_hi, int32_t layer_lo, float* grad_flat_dev, void* stream);
int dy_train_apply(dy_net* net, const float* grad_flat_dev, float lr, float grad_scale, void* stream);
/* Parity tap of the training step: which = 0 -> pre-BatchNorm conv output z of `layer`,
 * which = 1 -> d total_loss / d (layer output); fp32 NHWC [B,h,w,c]. */
int dy_train_get_tensor(dy_net* net, int32_t layer, int32_t which, int32_t B, float* out_dev, void* stream);
/* Current value of a variable by its TF name (Saver.save counterpart, train_yolo3_mask.py:221-226). */
int dy_get_weights(dy_net* net, const char* tf_name, float* host, int64_t capacity);

/* CRC-32C (Castagnoli) of `n` bytes continuing from `crc` (0 to start): the checksum of TensorFlow
 * checkpoint-V2 bundles (Saver.save / Saver.restore, train_yolo3_mask.py:104-111,221-226), used by
 * dis-yolo_b200/tf_checkpoint.py.  Host-only. */
uint32_t dy_crc32c(const void* data, uint64_t n, uint32_t crc);

/* Tuning / test / measurement overrides (A/B runs in scripts/, forced-mode parity tests); value -1 =
 * automatic for the planning switches.  Planning switches affect plans built afterwards
 * (dy_finalize_weights, dy_conv_layer, dy_train_init).
 *   "tc_resident"   0 never / 1 whenever it fits -- keep a CTA's weight slab resident in smem
 *   "tc_halo"       0 never / 1 whenever legal   -- one halo'd activation box feeds all 3 horizontal taps
 *   "tc_staged", "tc_tma_epi"   epilogue through shared memory / through TMA stores + TMA residual loads
 *   "tc_dual_issue" two MMA-issuer threads with split stage rings (thin tiles)
 *   "tc_pdl"           0: conv launches in plain stream order; 1 (default): programmatic dependent launch
 *   "train_pdl"        the same switch for the BN / elementwise / wgrad kernels of the training step
 *   "tc_dual_producer" with dual issue: 0 one TMA producer thread feeds both half rings, 1 (default) one each
 *   "tc_max_stages"    cap of the shared-memory pipeline depth (2..16)
 *   "tc_skip_epilogue"          measurement only: epilogues drain the accumulator and do nothing else
 *   "conv1_tc"      0: the stem on CUDA cores instead of the tcgen05 im2col kernel
 *   "tc_cta2"       0: never plan CTA pairs (2-CTA clusters, tcgen05.mma.cta_group::2 with M = 256, each CTA staging half
 *                   of the weight tile); 1: wherever legal; -1 (default): every 3x3 / plain 1x1 layer with >= 128 output
 *                   channels and at least one wave of tiles
 *   "tc_split_n"    0: the two epilogue warpgroups take alternate tiles; 1: both drain every tile, half of its columns
 *                   each, wherever legal; -1 (default): planner's choice (tiles of 128+ columns)
 *   "tc_epi_wg"     2 / 3 epilogue warpgroups (3 = the 512-thread instantiation, thin single-CTA tiles only; -1 planner)
 *   "tc_kchunk"     32: 32-wide K chunks everywhere (A/B);  "tc_slab"  32: 32-column TMA-epilogue staging slabs (A/B)
 *   "tc_fuse_tail"  0: convolutional82 as its own launch in dy_forward (default: fused into convolutional81)
 *   "wgrad_fuse_kw" 0 one CTA per tap / 1 (default) fused kernel rows where the N tile is kept / 2 always
 *   "wgrad_lbo_a", "wgrad_sbo_a", "wgrad_lbo_b", "wgrad_sbo_b"  bring-up: UMMA descriptor field overrides
 *   "mask_streaming_stores"     1 (default) st.global.cs in the mask-assembly kernel, 0 plain stores
 *   "mask_work_list"            1 mask assembly walks the existing detections with a fixed grid, 0 (default) one CTA per
 *                               possible (slab, detection, image)
 * Unknown names return DY_STATUS_NOTFOUND. */
int dy_set_option(const char* name, int32_t value);

/* Kernel launches issued by this library since the last call (bench.py's gpu_launches). */
int64_t dy_launch_count(int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* DISYOLO_H_ */
