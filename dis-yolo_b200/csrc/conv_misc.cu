// Non-tensor-core kernels around the conv engine:
//   * conv1 (3->32, 3x3, stride 1; yolo3_net_pos.py:159-161): Cin=3 is not a tensor-core shape;
//     direct fp32 conv that writes the bf16 space-to-depth P1 tensor conv2 (stride 2) consumes.
//   * layout converters between compact NHWC fp32 (the reference's layout) and the bf16 P1 forms.
//   * conv_ref: plain fp32 direct convolution = the FP32 verification mode of the network
//     (north_star: "within 1e-4 in a TF32/FP32-accumulate verification mode").
#include "conv_misc.cuh"

namespace dy {

namespace {

// ------------------------------------------------------------------------------------------
// conv1: one thread = one output pixel x 32 channels.  Weights [27][32] + scale/shift in smem.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
conv1_kernel(const float* __restrict__ img, const float* __restrict__ w_hwio, const float* __restrict__ scale,
             const float* __restrict__ shift, float alpha, int B, int H, int W,
             __nv_bfloat16* __restrict__ out_s2d, __nv_bfloat16* __restrict__ out_same) {
  __shared__ __align__(16) float sw[27 * 32];
  __shared__ __align__(16) float ssc[32];
  __shared__ __align__(16) float ssh[32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = w_hwio[i];
  if (threadIdx.x < 32) {
    ssc[threadIdx.x] = scale[threadIdx.x];
    ssh[threadIdx.x] = shift[threadIdx.x];
  }
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  if (x >= W) return;

  float in[27];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int yy = y + kh - 1;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int xx = x + kw - 1;
      const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
      const float* p = img + (((long long)n * H + (ok ? yy : 0)) * W + (ok ? xx : 0)) * 3;
      in[(kh * 3 + kw) * 3 + 0] = ok ? __ldg(p + 0) : 0.f;
      in[(kh * 3 + kw) * 3 + 1] = ok ? __ldg(p + 1) : 0.f;
      in[(kh * 3 + kw) * 3 + 2] = ok ? __ldg(p + 2) : 0.f;
    }
  }
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const float a = in[k];
    const float4* wr = reinterpret_cast<const float4*>(sw + k * 32);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 w4 = wr[c4];
      acc[4 * c4 + 0] = fmaf(a, w4.x, acc[4 * c4 + 0]);
      acc[4 * c4 + 1] = fmaf(a, w4.y, acc[4 * c4 + 1]);
      acc[4 * c4 + 2] = fmaf(a, w4.z, acc[4 * c4 + 2]);
      acc[4 * c4 + 3] = fmaf(a, w4.w, acc[4 * c4 + 3]);
    }
  }
  uint32_t packed[16];
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    float v0 = fmaf(acc[c], ssc[c], ssh[c]);
    float v1 = fmaf(acc[c + 1], ssc[c + 1], ssh[c + 1]);
    v0 = fmaxf(alpha * v0, v0);
    v1 = fmaxf(alpha * v1, v1);
    __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
    packed[c >> 1] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (out_s2d != nullptr) {
    const int Hq = H / 2 + 1, Wq = W / 2 + 1;
    const long long r = ((long long)n * Hq + (y >> 1)) * Wq + (x >> 1);
    const int cb = (((y & 1) << 1) | (x & 1)) * 32;
    uint4* d = reinterpret_cast<uint4*>(out_s2d + r * 128 + cb);
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
  }
  if (out_same != nullptr) {
    const long long r = ((long long)n * (H + 1) + y) * (W + 1) + x;
    uint4* d = reinterpret_cast<uint4*>(out_same + r * 32);
#pragma unroll
    for (int i = 0; i < 4; ++i) d[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
  }
}

// ------------------------------------------------------------------------------------------
// layout converters (one thread per element; test / parity-tap paths, not the hot path)
// ------------------------------------------------------------------------------------------
__global__ void nhwc_to_p1_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int N, int H,
                                  int W, int C, int form) {
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const __nv_bfloat16 v = __float2bfloat16_rn(src[i]);
    if (form == FORM_SAME) {
      dst[(((long long)n * (H + 1) + y) * (W + 1) + x) * C + c] = v;
    } else if (form == FORM_S2D) {
      const int Hq = H / 2 + 1, Wq = W / 2 + 1;
      dst[(((long long)n * Hq + (y >> 1)) * Wq + (x >> 1)) * (4 * C) + (((y & 1) << 1) | (x & 1)) * C + c] = v;
    } else {  // FORM_UP2
      const int Hu = 2 * H + 1, Wu = 2 * W + 1;
      const long long r = ((long long)n * Hu + 2 * y) * Wu + 2 * x;
      dst[r * C + c] = v;
      dst[(r + 1) * C + c] = v;
      dst[(r + Wu) * C + c] = v;
      dst[(r + Wu + 1) * C + c] = v;
    }
  }
}

__global__ void p1_to_nhwc_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int N, int H,
                                  int W, int C, int form) {
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    __nv_bfloat16 v;
    if (form == FORM_SAME) {
      v = src[(((long long)n * (H + 1) + y) * (W + 1) + x) * C + c];
    } else if (form == FORM_S2D) {
      const int Hq = H / 2 + 1, Wq = W / 2 + 1;
      v = src[(((long long)n * Hq + (y >> 1)) * Wq + (x >> 1)) * (4 * C) + (((y & 1) << 1) | (x & 1)) * C + c];
    } else {
      const int Hu = 2 * H + 1, Wu = 2 * W + 1;
      v = src[(((long long)n * Hu + 2 * y) * Wu + 2 * x) * C + c];
    }
    dst[i] = __bfloat162float(v);
  }
}

__global__ void planar_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int H, int W,
                                      int C) {
  const long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long t = i / C;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    dst[i] = src[(((long long)n * C + c) * H + y) * W + x];
  }
}

// ------------------------------------------------------------------------------------------
// conv_ref: fp32 direct conv, TF 'SAME' padding, optional [src0, up2(src1)] channel concat.
// thread (tx, ty): tx -> output channel, ty -> group of 4 consecutive output x.
// ------------------------------------------------------------------------------------------
constexpr int kRefPix = 4;

__global__ void __launch_bounds__(256)
conv_ref_kernel(RefConvArgs a) {
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  const int xg = (blockIdx.y * blockDim.y + threadIdx.y) * kRefPix;
  const int y = blockIdx.z % a.Ho;
  const int n = blockIdx.z / a.Ho;
  if (co >= a.cout || xg >= a.Wo) return;
  float acc[kRefPix];
#pragma unroll
  for (int i = 0; i < kRefPix; ++i) acc[i] = 0.f;
  const int cin = a.c0 + a.c1;
  for (int kh = 0; kh < a.k; ++kh) {
    const int yy = y * a.s + kh - a.pad_t;
    if (yy < 0 || yy >= a.Hi) continue;
    for (int kw = 0; kw < a.k; ++kw) {
      const float* wp = a.w + ((long long)(kh * a.k + kw) * cin) * a.cout + co;
      int xx[kRefPix];
      bool ok[kRefPix];
#pragma unroll
      for (int i = 0; i < kRefPix; ++i) {
        xx[i] = (xg + i) * a.s + kw - a.pad_l;
        ok[i] = (xg + i < a.Wo) && xx[i] >= 0 && xx[i] < a.Wi;
      }
      for (int ci = 0; ci < a.c0; ++ci) {
        const float wv = __ldg(wp + (long long)ci * a.cout);
#pragma unroll
        for (int i = 0; i < kRefPix; ++i) {
          if (ok[i]) acc[i] = fmaf(__ldg(a.src0 + (((long long)n * a.Hi + yy) * a.Wi + xx[i]) * a.c0 + ci), wv, acc[i]);
        }
      }
      if (a.c1 > 0) {
        const int Hh = a.Hi / 2, Wh = a.Wi / 2;
        for (int ci = 0; ci < a.c1; ++ci) {
          const float wv = __ldg(wp + (long long)(a.c0 + ci) * a.cout);
#pragma unroll
          for (int i = 0; i < kRefPix; ++i) {
            if (ok[i])
              acc[i] = fmaf(__ldg(a.src1 + (((long long)n * Hh + (yy >> 1)) * Wh + (xx[i] >> 1)) * a.c1 + ci), wv,
                            acc[i]);
          }
        }
      }
    }
  }
  const float sc = a.scale[co], sh = a.shift[co];
#pragma unroll
  for (int i = 0; i < kRefPix; ++i) {
    if (xg + i >= a.Wo) break;
    float v = fmaf(acc[i], sc, sh);
    if (a.act) v = fmaxf(a.alpha * v, v);
    const long long o = (((long long)n * a.Ho + y) * a.Wo + xg + i) * a.cout + co;
    if (a.residual) v += a.residual[o];
    a.out[o] = v;
  }
}

}  // namespace

int launch_conv1(const float* img, const float* w_hwio, const float* scale, const float* shift, float alpha,
                 int B, int H, int W, __nv_bfloat16* out_s2d, __nv_bfloat16* out_same, cudaStream_t st) {
  dim3 grid((W + 127) / 128, H, B);
  conv1_kernel<<<grid, 128, 0, st>>>(img, w_hwio, scale, shift, alpha, B, H, W, out_s2d, out_same);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

static int grid_for(long long total) {
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  return (int)g;
}

int launch_nhwc_to_p1(const float* src, __nv_bfloat16* dst, int N, int H, int W, int C, int form, cudaStream_t st) {
  nhwc_to_p1_kernel<<<grid_for((long long)N * H * W * C), 256, 0, st>>>(src, dst, N, H, W, C, form);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_p1_to_nhwc(const __nv_bfloat16* src, float* dst, int N, int H, int W, int C, int form, cudaStream_t st) {
  p1_to_nhwc_kernel<<<grid_for((long long)N * H * W * C), 256, 0, st>>>(src, dst, N, H, W, C, form);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_planar_to_nhwc(const float* src, float* dst, int N, int H, int W, int C, cudaStream_t st) {
  planar_to_nhwc_kernel<<<grid_for((long long)N * H * W * C), 256, 0, st>>>(src, dst, N, H, W, C);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_conv_ref(const RefConvArgs& a, int B, cudaStream_t st) {
  int bx = a.cout >= 64 ? 64 : 32;
  dim3 block(bx, 256 / bx);
  const int xgroups = (a.Wo + kRefPix - 1) / kRefPix;
  dim3 grid((a.cout + bx - 1) / bx, (xgroups + block.y - 1) / block.y, (unsigned)(B * a.Ho));
  conv_ref_kernel<<<grid, block, 0, st>>>(a);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

}  // namespace dy
