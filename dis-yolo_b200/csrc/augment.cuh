// Training data pipeline on the device: interface.
// Reference: utils/train_data.py -- load_mask (:321-338, polygon -> mask), extract_bboxes (:358-374),
// apply_random_scale_and_crop (:423-450), the flips of image_read / resize_mask (:376-421),
// add_salt_pepper_noise (:494-509), change_light (:511-521), linearmotion_blur3C (:452-481).
// Every random draw of the reference (np.random) is an INPUT here: the host decides, the device executes.
#pragma once
#include "common.cuh"

namespace dy {

// polygons of ONE image: verts [nv,2] double (x, y); poly [np,3] = (first vertex, vertex count, type: 1 = 'out',
// 0 = inner background region); inst [ni+1] = polygon range of each instance -> masks [ni,h,w] bytes (0 / 1)
int launch_polygon_masks(const double* verts, const int* poly, const int* inst, int ni, int h, int w,
                         unsigned char* masks, cudaStream_t st);
// extract_bboxes: boxes [n,4] int = (x1, y1, x2, y2) with x2 / y2 one past the last set pixel; (0,0,0,0) if empty
int launch_mask_boxes(const unsigned char* masks, int n, int h, int w, int* boxes, cudaStream_t st);

struct PlaceGeom {      // apply_random_scale_and_crop + flip
  int src_h, src_w;     // original image
  int new_w, new_h;     // cv2.resize target
  int dx, dy;           // placement of the resized image inside the net square (may be negative: cropped)
  int size;             // net_w = net_h
  int flip;             // 1 none, 2 horizontal (columns reversed), 3 vertical
};
// image uint8 [src_h,src_w,3] -> [size,size,3] uint8: cv2.resize INTER_LINEAR in its 8-bit fixed-point arithmetic,
// pad 127, flip
int launch_place_image_u8(const unsigned char* rgb, const PlaceGeom& g, unsigned char* out, cudaStream_t st);
// masks [n,src_h,src_w] bytes (0/1, resized as float32 like the reference) -> [n,size,size] bytes = around(v) != 0
int launch_place_masks(const unsigned char* masks, int n, const PlaceGeom& g, unsigned char* out, cudaStream_t st);
// im[r, c, :] = 1 at the salt coordinates, 0 at the pepper coordinates (salt first, then pepper, like the reference)
int launch_salt_pepper(unsigned char* img, int size, const int* salt_rc, int n_salt, const int* pepper_rc, int n_pepper,
                       cudaStream_t st);
// RGB -> HLS (cv2 8-bit), L = min(L * coeff, 255) truncated to uint8, HLS -> RGB; in place on npix pixels
int launch_change_light(unsigned char* img, long long npix, double coeff, cudaStream_t st);
// scipy.signal.convolve2d(channel.astype(float32), kernel3x3, mode='same', fillvalue=255).astype(uint8) per channel
int launch_motion_blur3(const unsigned char* img, int size, const float* kernel9, unsigned char* out, cudaStream_t st);
// image.astype(np.float32) / 255.0 (float32 division)
int launch_u8_div255_f32(const unsigned char* src, float* dst, long long n, cudaStream_t st);

}  // namespace dy
