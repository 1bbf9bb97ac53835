// conv1, layout converters and the fp32 verification conv: interface (see conv_misc.cu).
#pragma once
#include "common.cuh"

namespace dy {

enum Form : int { FORM_SAME = 0, FORM_S2D = 1, FORM_UP2 = 2 };

struct RefConvArgs {
  const float* src0;   // [B,Hi,Wi,c0]
  const float* src1;   // [B,Hi/2,Wi/2,c1] (nearest-upsampled + concatenated after src0) or null
  const float* w;      // HWIO [k,k,c0+c1,cout]
  const float* scale;  // [cout]
  const float* shift;  // [cout]
  const float* residual;  // [B,Ho,Wo,cout] or null (added after the activation)
  float* out;             // [B,Ho,Wo,cout]
  int Hi, Wi, Ho, Wo, c0, c1, cout, k, s, pad_t, pad_l, act;
  float alpha;
};

int launch_conv1(const float* img, const float* w_hwio, const float* scale, const float* shift, float alpha,
                 int B, int H, int W, __nv_bfloat16* out_s2d, __nv_bfloat16* out_same, int use_tc, int num_sms,
                 cudaStream_t st);
int launch_nhwc_to_p1(const float* src, __nv_bfloat16* dst, int N, int H, int W, int C, int form, cudaStream_t st);
int launch_p1_to_nhwc(const __nv_bfloat16* src, float* dst, int N, int H, int W, int C, int form, cudaStream_t st);
int launch_planar_to_nhwc(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st);
int launch_conv_ref(const RefConvArgs& a, int B, cudaStream_t st);

}  // namespace dy
