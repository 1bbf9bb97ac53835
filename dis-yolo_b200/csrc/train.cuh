// Training-step kernels (fp32): batch-statistics BatchNorm forward/backward, convolution dgrad and
// wgrad, the YOLO detection loss and the position-sensitive mask loss with their closed-form
// gradients, L2 regularisation, Adam.  Reference: yolo/yolo3_net_pos.py:71-107 (BN), :631-860
// (losses), :38/:61 (L2, total loss), train_yolo3_mask.py:55 (Adam).  Interface; see train.cu.
#pragma once
#include "common.cuh"

namespace dy {

// ---- BatchNorm with batch moments over (N,H,W) ------------------------------------------------
int launch_bn_stats(const float* z, long long M, int C, double* sum, double* sumsq, cudaStream_t st);
// a = gamma*rsqrt(var+eps), b = beta - mean*a; also writes mean / biased var / invstd
int launch_bn_finalize(const double* sum, const double* sumsq, long long M, int C, const float* gamma,
                       const float* beta, float eps, float* a, float* b, float* mean, float* var, float* invstd,
                       cudaStream_t st);
// y = act(z*a+b) (+residual)
int launch_bn_act(const float* z, const float* a, const float* b, int C, long long total, float alpha, int act,
                  const float* residual, float* y, cudaStream_t st);
// g = dy * leaky'(z*a+b);  s1 = sum g, s2 = sum g*xhat   (xhat = (z-mean)*invstd)
int launch_bn_bwd_reduce(const float* dy, const float* z, const float* a, const float* b, const float* mean,
                         const float* invstd, float alpha, int act, long long M, int C, double* s1, double* s2,
                         cudaStream_t st);
// mode 0: batch-stat BN  dz = gamma*invstd*(g - s1/M - xhat*s2/M);  mode 1: frozen affine  dz = g*a;
// mode 2: no BN (biased linear conv)  dz = dy
int launch_bn_bwd_apply(const float* dy, const float* z, const float* a, const float* b, const float* mean,
                        const float* invstd, const float* gamma, const double* s1, const double* s2, float alpha,
                        int act, int mode, long long M, int C, float* dz, cudaStream_t st);
int launch_copy_stats_to_grads(const double* s1, const double* s2, int C, float* dgamma, float* dbeta,
                               cudaStream_t st);

// ---- convolution backward ----------------------------------------------------------------------
struct ConvGeom {
  int B, Hi, Wi, Ho, Wo, cin, cout, k, s, pad_t, pad_l;
};
// w [k*k][cin][cout] -> wt [k*k][cout][cin]
int launch_weight_transpose(const float* w, float* wt, int kk, int cin, int cout, cudaStream_t st);
// dx [B,Hi,Wi,cin] (=/+=) sum dz[...] * w
int launch_conv_dgrad(const float* dz, const float* wt, float* dx, const ConvGeom& g, int accumulate, cudaStream_t st);
// dw [k*k][cin][cout] += x^T dz; x = concat(x0 [B,Hi,Wi,c0], up2(x1 [B,Hi/2,Wi/2,c1])); dw must be zeroed
int launch_conv_wgrad(const float* x0, int c0, const float* x1, int c1, const float* dz, float* dw, const ConvGeom& g,
                      int num_sms, cudaStream_t st);
// dst0 [.., c0] += src[.., :c0];  dst1 [B,H/2,W/2,c1] += 2x2-pooled src[.., c0:]   (either may be null)
int launch_split_accumulate(const float* src, int B, int H, int W, int c0, int c1, float* dst0, float* dst1,
                            cudaStream_t st);
int launch_add(float* dst, const float* src, long long n, cudaStream_t st);

// ---- losses ------------------------------------------------------------------------------------
struct YoloLossArgs {
  const float* pred[3];    // stride 8/16/32 head maps [B,g,g,3,8]
  const float* label[3];   // yolo3, yolo2, yolo1 labels, same shapes
  float* dpred[3];         // gradients (written)
  int g[3];
  int B, net;
  float anchors[18];
  const float* true_boxes; // [B,20,5]
  float ignore_thresh, object_scale, noobject_scale, class_scale, coord_scale;
  double* loss;            // [5] obj, noobj, cls, xy, wh  (accumulated; zero on entry)
};
int launch_yolo_loss(const YoloLossArgs& a, cudaStream_t st);

struct MaskLossArgs {
  const float* det;         // [B,max_det,6] detections of the training-mode forward
  const float* true_boxes;  // [B,20,5]
  const unsigned char* true_masks;   // [B,20,H,W] bool
  const int* perm_prop;     // [B,max_det] permutation standing in for tf.random_shuffle (:782)
  const int* perm_gt;       // [B,20]                                                 (:781)
  const float* mask_pos;    // [B,S,S,kk] logits (NHWC), or planar [B,kk,S,S] when mp_planar
  int mp_planar;
  float* dmask;             // [B,S,S,kk] gradient (accumulated with atomics; zero on entry)
  int B, max_det, S, H, k;
  float mask_scale, iou_thresh;
  float* rois;              // workspace [B,10,4]
  int* assign;              // workspace [B,10]
  int* npos;                // workspace [B]
  double* loss;             // [1]
};
int launch_mask_loss(const MaskLossArgs& a, cudaStream_t st);

int launch_sumsq(const float* p, long long n, double scale, double* out, cudaStream_t st);

// ---- optimizer -----------------------------------------------------------------------------------
// g' = g*grad_scale + l2*p ; m,v update ; p -= lr_t*m/(sqrt(v)+eps)      (tf.train.AdamOptimizer)
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
                float l2, float grad_scale, cudaStream_t st);
// ---- multi-tensor forms: ONE launch over a device table of segments (blockIdx.y = segment) ------------
struct ParamSeg {      // one trainable tensor
  float* p;            // fp32 master parameter
  long long off;       // offset into the flat gradient / Adam m / Adam v vectors
  long long n;
  float l2;            // 1e-4 for conv weights and biases, 0 for BatchNorm gamma / beta (:38,83-84)
  int pad_;
};
struct BnSeg {         // one unlocked BatchNorm layer: moving averages (:92-95) + inference fold
  float *mean, *var;
  const float *bmean, *bvar, *gamma, *beta;
  float *scale, *shift;
  int C, pad_;
};
int launch_adam_multi(const ParamSeg* segs_dev, int nseg, const float* g, float* m, float* v, float lr_t, float b1,
                      float b2, float eps, float grad_scale, cudaStream_t st);
// out += sum over segments with l2 > 0 of 0.5 * l2 * sum p^2
int launch_sumsq_multi(const ParamSeg* segs_dev, int nseg, double* out, cudaStream_t st);
int launch_bn_post_multi(const BnSeg* segs_dev, int nseg, float decay, float eps, cudaStream_t st);

int launch_moving_update(float* mov_mean, float* mov_var, const float* bmean, const float* bvar, int C, float decay,
                         cudaStream_t st);
int launch_refold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int C,
                  float* scale, float* shift, cudaStream_t st);

}  // namespace dy
