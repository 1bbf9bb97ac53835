// Pre- / post-processing on either side of the hot path: interface.
//   letterbox   = image_read (calculate_test_map.py:149-176, utils/val_data.py:36-63): bilinear resize of the
//                 RGB image to fit image_size, 127-padding, /255
//   postprocess = the per-detection loop of calculate_test_map.py:233-269 / utils/validation_map.py:137-166:
//                 correct_yolo_boxes (:121-138), crop of the [S,S] sigmoid map, cv2.resize INTER_LINEAR to
//                 the box size in the original image, > 0.5, paste; merged semantic mask
#pragma once
#include "common.cuh"

namespace dy {

// geometry image_read computes on the host (integer arithmetic of the reference, :152-158, :167-168)
struct LetterboxGeom {
  int src_h, src_w;   // original image
  int new_h, new_w;   // resized extent
  int top, left;      // (size - new) // 2
  int size;
};
LetterboxGeom letterbox_geom(int src_h, int src_w, int size);

// n uint8 values -> fp32 v/255 (the float64 quotient rounded to float, like image_read's `/ 255.`)
int launch_u8_to_f32(const unsigned char* src, float* dst, long long n, cudaStream_t st);

// B same-shape images: rgb + b*rgb_stride [src_h, src_w, 3] uint8 (device) -> out [B, size, size, 3] fp32 (device)
int launch_letterbox(const unsigned char* rgb, long long rgb_stride, int B, const LetterboxGeom& g, float* out,
                     cudaStream_t st);

struct PostDet {        // one detection in original-image coordinates
  int x1, y1, x2, y2;   // corrected box (correct_yolo_boxes)
  int cx1, cy1, cx2, cy2;   // crop of the [S,S] map: np.around(norm * S)
  int cls;
  int valid;            // 0: (y2-y1)*(x2-x1) <= 0 or empty crop -> skipped like the reference's `continue`
  double scale_x, scale_y;   // cv2.resize source step per destination pixel
};
// det_box [n_max,6] = (y1,x1,y2,x2 normalised, class, score), count on the device
// B images of the same original shape in one launch pair (every array is the [B, ...] stack of the one-image form)
int launch_postprocess(const float* det_box, const int* count, int B, int n_max, const float* masks, int S,
                       int image_h, int image_w, int net_size, PostDet* ws, int* boxes_out, unsigned char* valid_out,
                       unsigned char* full_masks, unsigned char* merged, cudaStream_t st);

// training labels (utils/train_data.py:134-178, flips :189-228, normalisation :258-262)
struct LabelArgs {
  const float* boxes;     // [B,max_box,5] (x1,y1,x2,y2,class) in ORIGINAL image pixels, first nbox[b] rows valid
  const int* nbox;        // [B]
  const float* place;     // [B,4] (sx, sy, dx, dy): letterbox / scale-crop placement into the net square
  const int* flip;        // [B] 1 none, 2 horizontal, 3 vertical (null = none)
  float* yolo[3];         // yolo3 (stride 8), yolo2 (stride 16), yolo1 (stride 32): [B,g,g,3,5+C], zeroed here
  float* true_boxes;      // [B,max_box,5] (xc,yc,w,h)/net, class; zeroed here
  int grid[3];
  float anchors[18];
  int B, max_box, num_class, net;
};
int launch_assign_labels(const LabelArgs& a, cudaStream_t st);

// compute_overlaps_masks (utils/voc_eval_mask.py:38-56): IoU of every mask of set 1 with every mask of set 2;
// masks are [n, P] bytes (non-zero = inside), ws = n1*n2 + n1 + n2 ints of scratch, out [n1, n2] fp32
int launch_mask_overlaps(const unsigned char* m1, int n1, const unsigned char* m2, int n2, long long P, int* ws,
                         float* out, cudaStream_t st);

}  // namespace dy
