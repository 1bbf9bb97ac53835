// YOLO head decode, score threshold, per-class NMS, top-k and position-sensitive mask assembly:
// interface (see postproc.cu).  Reference: yolo/yolo3_net_pos.py:465-628 and :862-952.
#pragma once
#include "common.cuh"

namespace dy {

constexpr int kMaxClasses = 8;
constexpr int kMaxK = 7;   // k x k position-sensitive bins, reference supports 3/5/7 (:889-907)

struct Cand {       // one above-threshold candidate (32 bytes)
  float y1, x1, y2, x2;
  float score;
  int idx;          // candidate index: scale offset + (y*g + x)*3 + anchor  (:527-542)
  int cls;
  int pad;
};

struct DecodeArgs {
  const float* yolo[3];   // stride 8, 16, 32 maps, fp32 [B,g,g,3*(5+C)]   (list order of :353)
  int g[3];               // grid sizes
  int B, num_class, net;  // net = 32 * g[2]                                 (:474-476)
  float anchors[18];      // 9 x (w,h) pixels, 3 per scale                   (:495-496)
  const float* windows;   // [B,4] y1,x1,y2,x2 clip windows                  (:554-555)
  float thresh;           // strict >                                        (:558)
  float* dense_box;       // optional [B,N0,4]
  int* dense_cls;         // optional [B,N0]
  float* dense_score;     // optional [B,N0]
  Cand* cand;             // [B,cap] compacted survivors (any order)
  int* cand_count;        // [B], must be zero on entry
  int cap;
};

struct NmsArgs {
  const Cand* cand;
  const int* cand_count;
  int cap, B, num_class, max_det;
  float iou_thr;
  int* sel;       // [B,num_class,max_det] positions into cand
  int* sel_cnt;   // [B,num_class]
  // scratch for classes with more candidates than the kernel keeps in shared memory: [B,num_class,cap] each
  float4* ovf_box;
  unsigned long long* ovf_key;
  int* ovf_pos;
};

struct FinalizeArgs {
  const Cand* cand;
  int cap, B, num_class, max_det;
  const int* sel;
  const int* sel_cnt;
  int S, k;            // score-map size and bins per side
  float* det_raw;      // [B,max_det,6] zero padded: filter_detections output      (:615-628)
  int* raw_count;      // [B]
  float* det_box;      // [B,max_det,6] rows surviving val_test's w>0,h>0 filter   (:876-880)
  int* det_count;      // [B]
  int* edges;          // [B,max_det,2*(kMaxK+1)] gx[0..k], gy[0..k]               (:889-897)
};

struct MaskArgs {
  const float* score;   // score maps
  long long s_img, s_ch, s_row, s_pix;   // element strides (planar or NHWC)
  const int* det_count;
  const int* edges;
  int B, max_det, S, k;
  float* out;           // [B,max_det,S,S]
};

int launch_decode(const DecodeArgs& a, cudaStream_t st);
int launch_nms(const NmsArgs& a, cudaStream_t st);
int launch_finalize(const FinalizeArgs& a, cudaStream_t st);
int launch_masks(const MaskArgs& a, cudaStream_t st);
// box-cropped maps: off[B*max_det+1] = exclusive scan of the crop areas (floats), then the packed crops
int launch_crop_offsets(const MaskArgs& a, long long* off, cudaStream_t st);
int launch_masks_cropped(const MaskArgs& a, const long long* off, float* out, cudaStream_t st);
void masks_set_work_list(int on);   // 1: fixed grid walking the existing (slab, detection) items; 0 (default): one CTA per possible item
void masks_set_streaming(int on);   // 1 (default): st.global.cs streaming stores, 0: plain stores

}  // namespace dy
