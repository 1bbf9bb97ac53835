// Training step on the tensor-core engine (bf16 operands, fp32 accumulate, fp32 master weights):
// interface.  Reference: the graph TF differentiates for sess.run([total_loss, optimizer])
// (train_yolo3_mask.py:146-149,216): tf.nn.conv2d backward (yolo3_net_pos.py:125,142), batch-stat
// BatchNorm (:88-98), leaky (:69), resize_nearest_neighbor + concat backward (:290-291 ...).
//
//   dgrad   = the forward tcgen05 conv kernel (conv_tc.cu) run on dz with the 180-degree rotated,
//             in/out-transposed weights (pack_dgrad), accumulating into the producer's gradient
//             through the kernel's residual input
//   wgrad   = wgrad_tc_kernel: D[ci, co] = sum over pixels X[p + tap shift, ci] * dz[p, co]; the
//             contraction runs over P1 pixel rows, so both operands are "MN-major" UMMA operands --
//             the same [64 pixel rows x 64 channels] TMA boxes the forward uses, no transposes
//   BN / leaky / residual / upsample: vectorised bf16 kernels over the P1 layout (HBM bound)
#pragma once
#include "common.cuh"

namespace dy {

// ---- elementwise over P1 bf16 [rows, C] (C % 8 == 0) --------------------------------------------
// per-channel sum / sum of squares over all rows (pad rows are zero and contribute nothing)
int launch_bn_stats_p1(const __nv_bfloat16* z, long long rows, int C, double* sum, double* sumsq, cudaStream_t st);
// y = leaky(z*a+b) (+residual) at valid pixels -> out_same (P1, same geometry) and / or out_up
// (P1 of the 2x nearest-upsampled tensor); pad pixels are never written (they stay zero)
int launch_bn_act_p1(const __nv_bfloat16* z, const float* a, const float* b, const __nv_bfloat16* residual, int B,
                     int H, int W, int C, float alpha, int act, __nv_bfloat16* out_same, __nv_bfloat16* out_up,
                     cudaStream_t st, __nv_bfloat16* out_s2d = nullptr);
// (out_s2d: the space-to-depth copy [N, H/2+1, W/2+1, 4C] a stride-2 consumer reads)
// the same with bn_finalize fused in: a, b, mean, biased var, invstd are derived from the per-channel sums
// (and written out for the backward pass by block 0)
int launch_bn_finalize_act_p1(const __nv_bfloat16* z, const double* sum, const double* sumsq, long long M,
                              const float* gamma, const float* beta, float eps, float* a, float* b, float* mean,
                              float* var, float* invstd, const __nv_bfloat16* residual, int B, int H, int W, int C,
                              float alpha, int act, __nv_bfloat16* out_same, __nv_bfloat16* out_up, cudaStream_t st,
                              __nv_bfloat16* out_s2d = nullptr);
// g = dy * leaky'(z*a+b);  s1 = sum g, s2 = sum g*xhat
int launch_bn_bwd_reduce_p1(const __nv_bfloat16* dy, const __nv_bfloat16* z, const float* a, const float* b,
                            const float* mean, const float* invstd, float alpha, int act, long long rows, int C,
                            double* s1, double* s2, cudaStream_t st);
// mode 0: dz = gamma*invstd*(g - s1/M - xhat*s2/M); mode 1 (frozen affine): dz = g*a.  Writes EVERY
// row of [0, round_up(rows, 64)): zeros at pad pixels and in the tail (dz is a shared scratch)
int launch_bn_bwd_apply_p1(const __nv_bfloat16* dy, const __nv_bfloat16* z, const float* a, const float* b,
                           const float* mean, const float* invstd, const float* gamma, const double* s1,
                           const double* s2, float alpha, int act, int mode, int B, int H, int W, int C,
                           __nv_bfloat16* dz, float* dgamma, float* dbeta, cudaStream_t st);
// (dgamma / dbeta non-null: block 0 also writes d gamma = s2, d beta = s1 into the flat gradient vector)
// fp32 [B,H,W,C] -> bf16 P1 [B,H+1,W+1,Cg] (channels >= C and pad pixels zero; tail rows zeroed)
// colsum (may be null): [C] fp32 receives the per-channel sums of src (zeroed here first)
int launch_f32_to_p1(const float* src, int B, int H, int W, int C, __nv_bfloat16* dst, int Cg, float* colsum,
                     cudaStream_t st);
// 2x2 sum pooling: P1 [B,2h+1,2w+1,C] -> P1 [B,h+1,w+1,C]   (backward of nearest-neighbour upsampling)
int launch_pool2x2_p1(const __nv_bfloat16* src, int B, int h, int w, int C, __nv_bfloat16* dst, cudaStream_t st);
int launch_add_p1(__nv_bfloat16* dst, const __nv_bfloat16* src, long long n, cudaStream_t st);
int launch_copy_p1(__nv_bfloat16* dst, const __nv_bfloat16* src, long long n, cudaStream_t st);

// ---- weight repacking (fp32 master HWIO -> bf16 GEMM operands) ------------------------------------
// [K][cout] -> [cout_pad][K]
int launch_pack_fwd_bf16(const float* w, int K, int cout, int cout_pad, __nv_bfloat16* out, cudaStream_t st);
// dgrad operand for input channels [ci0, ci0+cin_sel): out[ci][(kh'*k+kw')*Cg + co] =
// w[k-1-kh'][k-1-kw'][ci0+ci][co], zero for co >= cout
int launch_pack_dgrad_bf16(const float* w, int k, int cin_total, int ci0, int cin_sel, int cout, int Cg,
                           __nv_bfloat16* out, cudaStream_t st);

// dgrad operands of a 3x3 stride-2 conv: four regions (parity blocks 0..3 with 4 / 2 / 2 / 1 taps), region b =
// [cin][ntap_b * Cg]; starts at cin*Cg*{0, 4, 6, 8}
int launch_pack_dgrad_s2_bf16(const float* w, int cin, int cout, int Cg, __nv_bfloat16* out, cudaStream_t st);
// convolutional1's weight gradient [27][32] from the fp32 image and dz (P1 bf16, 32 channels); zeroes dw first
int launch_conv1_wgrad(const float* images, const __nv_bfloat16* dz, int B, int H, int W, float* dw, cudaStream_t st);

// all trained layers in ONE launch each (blockIdx.y = layer): forward operand and dgrad operands
struct PackSeg {
  const float* w;          // fp32 master HWIO
  __nv_bfloat16* wpk;      // [cout_pad][K]
  __nv_bfloat16* wdg0;     // dgrad operand towards src0 or null
  __nv_bfloat16* wdg1;     // dgrad operand towards src1 or null
  int K, cout, cout_pad, k, cin0, cin1, Cg, pad_;
};
int launch_pack_multi(const PackSeg* segs_dev, int nseg, int max_tiles, cudaStream_t st);

// ---- wgrad on tcgen05 -------------------------------------------------------------------------
struct WgradParams {
  int M;               // pixel rows to contract over = B*(H+1)*(W+1)
  int nsrc;            // 1, or 2 for a 1x1 conv over concat[skip, up]
  int src_c[2];        // channels of each source
  int src_aw[2];       // TMA box width of each source: 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
  int src_koff[2];     // first dW row of each source inside a tap
  int src_blk[2];      // ceil(c / 128)
  int fuse_kw;         // 3x3 only: 1 = one CTA per kernel ROW, three accumulators (kw = 0,1,2) share the dz boxes
  int ntap;            // 1 or 9
  int tap_shift[9];    // pixel-row shift of X for each tap: (kh-1)*(W+1)+(kw-1)
  int tap_col[9];      // first channel of X for each tap: 0, or the parity block ((kh&1)*2+(kw&1))*cin of the
                       // space-to-depth source of a stride-2 conv (row shift (kh>>1)*(W+1)+(kw>>1) then)
  int cin_total;       // dW rows per tap
  int cout;            // real output channels (dW row pitch)
  int zc, z_aw;        // dz channels (multiple of 32) and its box width
  int block_n, n_tiles_n;
  int ksplit, chunks_per_split, total_chunks;
  int num_stages;
  float* dw;           // [ntap*cin_total][cout] fp32, accumulated with atomics (zero on entry)
};
struct WgradPlan {
  CUtensorMap x[2], z;
  WgradParams p;
};
// x0/x1: P1 activations (x1 = the materialised 2x-upsampled tensor of the concat branch or null),
// dz: P1 gradient w.r.t. the conv output with zc channels; rows_max = max_batch*(H+1)*(W+1)
// stride 2 (3x3, no concat): x0 is the space-to-depth copy [rows, 4*c0] the forward consumes; H, W = OUTPUT extent
int build_wgrad_plan(const __nv_bfloat16* x0, int c0, const __nv_bfloat16* x1, int c1, const __nv_bfloat16* dz,
                     int zc, int cout, int k, int H, int W, long long rows_max, float* dw, WgradPlan* plan,
                     int stride = 1);
int run_wgrad_plan(WgradPlan& plan, int B, int H, int W, int num_sms, cudaStream_t st);
// measurement / bring-up aid: descriptor field overrides (0 = computed value)
void wgrad_set_debug(int lbo_a, int sbo_a, int lbo_b, int sbo_b);
void train_set_pdl(int on);    // A/B aid: 0 = plain stream order for the BN / elementwise / wgrad kernels, 1 (default) = PDL
void wgrad_set_fuse(int on);   // A/B aid: 0 = one CTA per tap (no sharing), 1 (default) = fused kernel rows

}  // namespace dy
