// Training data pipeline kernels (see augment.cuh).  All HBM-bound byte / integer work: one thread per output
// pixel, coalesced rows, no tensor cores.  Arithmetic restates what the reference's dependencies compute:
//   * cv2.resize(uint8, INTER_LINEAR): 11-bit fixed-point coefficients cvRound(w * 2048) from a FLOAT source
//     coordinate; horizontal pass in int32, vertical ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; horizontal
//     taps clamp with the weight zeroed, vertical taps keep the weights and clip the row indices (recovered from
//     cv2 4.13 with random images: bit-exact, tests/golden/augment_kat.npz)
//   * cv2.cvtColor RGB2HLS / HLS2RGB on uint8: float32 with v * (1/255), fused multiply-add on the +120 / +240
//     hue branches, 2 - (max + min) in the saturation denominator, round-half-even to uint8 (checked against cv2 on
//     all 2^24 RGB triples: 3 differ by one hue unit; HLS2RGB exact on all 181 * 65536 inputs)
//   * scipy.signal.convolve2d(float32, mode='same', fillvalue=255): taps accumulated in kernel row-major order
//   * skimage.draw.polygon: O'Rourke's crossing test with vertex / edge inclusion (skimage._shared.geometry)
#include "augment.cuh"

namespace dy {

namespace {

// ------------------------------------------------------------------------------------------
// polygon -> mask (load_mask, utils/train_data.py:321-338)
// ------------------------------------------------------------------------------------------
// skimage._shared.geometry.point_in_polygon: 0 outside, non-zero inside / on a vertex / on an edge
__device__ __forceinline__ int point_in_polygon(const double* v, int n, double x, double y) {
  const double eps = 1e-12;
  int l_cross = 0, r_cross = 0;
  double x1 = v[2 * (n - 1)] - x, y1 = v[2 * (n - 1) + 1] - y;
  for (int i = 0; i < n; ++i) {
    const double x0 = v[2 * i] - x, y0 = v[2 * i + 1] - y;
    if (-eps < x0 && x0 < eps && -eps < y0 && y0 < eps) return 2;     // vertex
    if ((y0 > 0) != (y1 > 0)) {
      if ((x0 * y1 - x1 * y0) / (y1 - y0) > 0) ++r_cross;
    }
    if ((y0 < 0) != (y1 < 0)) {
      if ((x0 * y1 - x1 * y0) / (y1 - y0) < 0) ++l_cross;
    }
    x1 = x0;
    y1 = y0;
  }
  if ((r_cross & 1) != (l_cross & 1)) return 3;                       // edge
  return r_cross & 1;
}

__global__ void __launch_bounds__(128) polygon_mask_kernel(const double* __restrict__ verts, const int* __restrict__ poly,
                                                           const int* __restrict__ inst, int h, int w,
                                                           unsigned char* __restrict__ masks) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, i = blockIdx.z;
  if (x >= w) return;
  unsigned char state = 0;
  for (int p = inst[i]; p < inst[i + 1]; ++p) {
    const int v0 = poly[3 * p], nv = poly[3 * p + 1], type = poly[3 * p + 2];
    const double* v = verts + 2 * (size_t)v0;
    // skimage.draw.polygon only visits the polygon's bounding rows / columns [int(min), ceil(max)]
    if (nv > 0 && point_in_polygon(v, nv, (double)x, (double)y)) state = type ? 1 : 0;
    for (int k = 0; k < nv; ++k)                                      // each_mask[y_points, x_points] = True
      if ((double)x == v[2 * k] && (double)y == v[2 * k + 1]) state = 1;
  }
  masks[((size_t)i * h + y) * w + x] = state;
}

// ------------------------------------------------------------------------------------------
// bounding boxes of masks (extract_bboxes, :358-374)
// ------------------------------------------------------------------------------------------
__global__ void mask_boxes_init_kernel(int* boxes, int n, int h, int w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { boxes[4 * i] = w; boxes[4 * i + 1] = h; boxes[4 * i + 2] = -1; boxes[4 * i + 3] = -1; }
}
__global__ void __launch_bounds__(256) mask_boxes_kernel(const unsigned char* __restrict__ masks, int h, int w,
                                                         int* __restrict__ boxes) {
  const int i = blockIdx.y;
  const unsigned char* m = masks + (size_t)i * h * w;
  int x1 = w, y1 = h, x2 = -1, y2 = -1;
  const long long total = (long long)h * w;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    if (m[p]) {
      const int y = (int)(p / w), x = (int)(p - (long long)y * w);
      x1 = min(x1, x); y1 = min(y1, y); x2 = max(x2, x); y2 = max(y2, y);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    x1 = min(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = min(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    x2 = max(x2, __shfl_xor_sync(0xffffffffu, x2, o)); y2 = max(y2, __shfl_xor_sync(0xffffffffu, y2, o));
  }
  if ((threadIdx.x & 31) == 0 && x2 >= 0) {
    atomicMin(boxes + 4 * i, x1); atomicMin(boxes + 4 * i + 1, y1);
    atomicMax(boxes + 4 * i + 2, x2); atomicMax(boxes + 4 * i + 3, y2);
  }
}
__global__ void mask_boxes_final_kernel(int* boxes, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (boxes[4 * i + 2] < 0) { boxes[4 * i] = boxes[4 * i + 1] = boxes[4 * i + 2] = boxes[4 * i + 3] = 0; }
  else { boxes[4 * i + 2] += 1; boxes[4 * i + 3] += 1; }              // x2, y2 one past the last pixel
}

// ------------------------------------------------------------------------------------------
// scale / crop / place / flip
// ------------------------------------------------------------------------------------------
struct FixTap {
  int i0, i1;
  int a0, a1;     // cvRound(w * 2048)
};
// cv::resize, 8-bit fixed point: fx = (float)((d + 0.5) * scale - 0.5); s = cvFloor(fx); fx -= s
__device__ __forceinline__ FixTap fix_tap(int d, double scale, int n, bool horizontal) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  FixTap t;
  if (horizontal) {               // clamped taps lose their weight
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n - 1) { s = n - 1; f = 0.f; }
    t.i0 = s;
    t.i1 = min(s + 1, n - 1);
  } else {                        // vertical: weights kept, row indices clipped
    t.i0 = min(max(s, 0), n - 1);
    t.i1 = min(max(s + 1, 0), n - 1);
  }
  t.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  t.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
  return t;
}

// destination pixel (y, x) of the net square -> pixel of the placed canvas before the flip
__device__ __forceinline__ void unflip(const PlaceGeom& g, int y, int x, int* cy, int* cx) {
  *cy = g.flip == 3 ? g.size - 1 - y : y;
  *cx = g.flip == 2 ? g.size - 1 - x : x;
}

__global__ void __launch_bounds__(128) place_image_u8_kernel(const unsigned char* __restrict__ rgb, PlaceGeom g,
                                                             double scale_x, double scale_y,
                                                             unsigned char* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= g.size) return;
  int cy, cx;
  unflip(g, y, x, &cy, &cx);
  unsigned char* o = out + ((size_t)y * g.size + x) * 3;
  const int ry = cy - g.dy, rx = cx - g.dx;
  if (ry < 0 || ry >= g.new_h || rx < 0 || rx >= g.new_w) {
    o[0] = 127; o[1] = 127; o[2] = 127;
    return;
  }
  if (g.new_w == g.src_w && g.new_h == g.src_h) {        // cv::resize copies when the sizes agree
    const unsigned char* s = rgb + ((size_t)ry * g.src_w + rx) * 3;
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
    return;
  }
  const FixTap tx = fix_tap(rx, scale_x, g.src_w, true), ty = fix_tap(ry, scale_y, g.src_h, false);
  const unsigned char* r0 = rgb + (size_t)ty.i0 * g.src_w * 3;
  const unsigned char* r1 = rgb + (size_t)ty.i1 * g.src_w * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int s0 = (int)r0[tx.i0 * 3 + c] * tx.a0 + (int)r0[tx.i1 * 3 + c] * tx.a1;
    const int s1 = (int)r1[tx.i0 * 3 + c] * tx.a0 + (int)r1[tx.i1 * 3 + c] * tx.a1;
    const int v = (((ty.a0 * (s0 >> 4)) >> 16) + ((ty.a1 * (s1 >> 4)) >> 16) + 2) >> 2;
    o[c] = (unsigned char)min(max(v, 0), 255);
  }
}

// float32 path of cv::resize (what the reference's float masks go through): coordinate in double, weight rounded
// to float, clamped taps lose their weight (dis-yolo_b200/csrc/imgproc.cu linear_tap)
struct FTap {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ FTap f_tap(int d, double scale, int n) {
  const double fd = ((double)d + 0.5) * scale - 0.5;
  int s = (int)floor(fd);
  float f = (float)(fd - (double)s);
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= n - 1) { s = n - 1; f = 0.f; }
  FTap t;
  t.i0 = s; t.i1 = s + 1 < n ? s + 1 : n - 1; t.w0 = 1.f - f; t.w1 = f;
  return t;
}

__global__ void __launch_bounds__(128) place_masks_kernel(const unsigned char* __restrict__ masks, PlaceGeom g,
                                                          double scale_x, double scale_y,
                                                          unsigned char* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, i = blockIdx.z;
  if (x >= g.size) return;
  int cy, cx;
  unflip(g, y, x, &cy, &cx);
  unsigned char* o = out + ((size_t)i * g.size + y) * g.size + x;
  const int ry = cy - g.dy, rx = cx - g.dx;
  if (ry < 0 || ry >= g.new_h || rx < 0 || rx >= g.new_w) { *o = 0; return; }
  const unsigned char* m = masks + (size_t)i * g.src_h * g.src_w;
  if (g.new_w == g.src_w && g.new_h == g.src_h) { *o = m[(size_t)ry * g.src_w + rx] ? 1 : 0; return; }
  const FTap tx = f_tap(rx, scale_x, g.src_w), ty = f_tap(ry, scale_y, g.src_h);
  const unsigned char* r0 = m + (size_t)ty.i0 * g.src_w;
  const unsigned char* r1 = m + (size_t)ty.i1 * g.src_w;
  const float a = r0[tx.i0] ? 1.f : 0.f, b = r0[tx.i1] ? 1.f : 0.f, c = r1[tx.i0] ? 1.f : 0.f, d = r1[tx.i1] ? 1.f : 0.f;
  const float h0 = __fadd_rn(__fmul_rn(a, tx.w0), __fmul_rn(b, tx.w1));
  const float h1 = __fadd_rn(__fmul_rn(c, tx.w0), __fmul_rn(d, tx.w1));
  const float v = __fadd_rn(__fmul_rn(h0, ty.w0), __fmul_rn(h1, ty.w1));
  *o = rintf(v) != 0.f ? 1 : 0;                         // np.around(...).astype(bool): half to even, 0.5 -> 0
}

// ------------------------------------------------------------------------------------------
// salt & pepper, lighting, motion blur
// ------------------------------------------------------------------------------------------
__global__ void salt_pepper_kernel(unsigned char* img, int size, const int* rc, int n, unsigned char value) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = rc[2 * i], c = rc[2 * i + 1];
  if (r < 0 || r >= size || c < 0 || c >= size) return;
  unsigned char* p = img + ((size_t)r * size + c) * 3;
  p[0] = value; p[1] = value; p[2] = value;
}

__device__ __forceinline__ unsigned char sat_u8(float v) {          // saturate_cast<uchar>(float) = cvRound, clamped
  const int r = __float2int_rn(v);
  return (unsigned char)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256) change_light_kernel(unsigned char* __restrict__ img, long long npix, double coeff) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= npix) return;
  unsigned char* p = img + 3 * i;
  const float k255 = 1.0f / 255.0f;
  // ---- RGB2HLS, 8-bit ----
  const float r = __fmul_rn((float)p[0], k255), g = __fmul_rn((float)p[1], k255), b = __fmul_rn((float)p[2], k255);
  const float vmax = fmaxf(fmaxf(r, g), b), vmin = fminf(fminf(r, g), b);
  const float diff = __fsub_rn(vmax, vmin), sum = __fadd_rn(vmax, vmin);
  const float l = __fmul_rn(sum, 0.5f);
  float h = 0.f, s = 0.f;
  if (diff > 1.1920929e-07f) {
    s = __fdiv_rn(diff, l < 0.5f ? sum : __fsub_rn(2.f, sum));
    const float d60 = __fdiv_rn(60.f, diff);
    if (vmax == r) h = __fmul_rn(__fsub_rn(g, b), d60);
    else if (vmax == g) h = __fmaf_rn(__fsub_rn(b, r), d60, 120.f);
    else h = __fmaf_rn(__fsub_rn(r, g), d60, 240.f);
    if (h < 0.f) h = __fadd_rn(h, 360.f);
  }
  const unsigned char H = sat_u8(__fmul_rn(h, 0.5f)), L0 = sat_u8(__fmul_rn(l, 255.f)), S = sat_u8(__fmul_rn(s, 255.f));
  // ---- the reference's edit: float64 L * coeff, clipped at 255, truncated by the uint8 cast (:513-518) ----
  double Ld = (double)L0 * coeff;
  if (Ld > 255.0) Ld = 255.0;
  const unsigned char L = (unsigned char)Ld;
  // ---- HLS2RGB, 8-bit ----
  const float lf = __fmul_rn((float)L, k255), sf = __fmul_rn((float)S, k255);
  float R, G, B;
  if (S == 0) {
    R = G = B = lf;
  } else {
    const float p2 = lf <= 0.5f ? __fmul_rn(lf, __fadd_rn(1.f, sf)) : __fsub_rn(__fadd_rn(lf, sf), __fmul_rn(lf, sf));
    const float p1 = __fsub_rn(__fmul_rn(2.f, lf), p2);
    float hh = __fmul_rn((float)H, (float)(6.0 / 180.0));
    if (hh >= 6.f) hh = __fsub_rn(hh, 6.f);
    int sector = (int)floorf(hh);
    const float fr = __fsub_rn(hh, (float)sector);
    sector = min(max(sector, 0), 5);
    const float dp = __fsub_rn(p2, p1);
    const float tab[4] = {p2, p1, __fadd_rn(p1, __fmul_rn(dp, __fsub_rn(1.f, fr))), __fadd_rn(p1, __fmul_rn(dp, fr))};
    const int sd[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};   // b, g, r
    B = tab[sd[sector][0]]; G = tab[sd[sector][1]]; R = tab[sd[sector][2]];
  }
  p[0] = sat_u8(__fmul_rn(R, 255.f)); p[1] = sat_u8(__fmul_rn(G, 255.f)); p[2] = sat_u8(__fmul_rn(B, 255.f));
}

struct Ker9 {
  float k[9];
};
__global__ void __launch_bounds__(128) motion_blur3_kernel(const unsigned char* __restrict__ img, int size, Ker9 ker,
                                                           unsigned char* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= size) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int yy = y + 1 - i, xx = x + 1 - j;                     // true convolution: flipped kernel
        const float v = (yy >= 0 && yy < size && xx >= 0 && xx < size) ? (float)img[((size_t)yy * size + xx) * 3 + c] : 255.f;
        acc = __fadd_rn(acc, __fmul_rn(ker.k[i * 3 + j], v));
      }
    out[((size_t)y * size + x) * 3 + c] = (unsigned char)(int)acc;    // astype(uint8): truncation
  }
}

__global__ void __launch_bounds__(256) u8_div255_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst,
                                                        long long n) {
  __shared__ float lut[256];
  lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.f);
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = lut[src[i]];
}

}  // namespace

int launch_polygon_masks(const double* verts, const int* poly, const int* inst, int ni, int h, int w,
                         unsigned char* masks, cudaStream_t st) {
  DY_CHECK(ni >= 1 && ni <= 65535 && h >= 1 && h <= 65535 && w >= 1, "geometry");
  dim3 grid((w + 127) / 128, h, ni);
  polygon_mask_kernel<<<grid, 128, 0, st>>>(verts, poly, inst, h, w, masks);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_mask_boxes(const unsigned char* masks, int n, int h, int w, int* boxes, cudaStream_t st) {
  DY_CHECK(n >= 1 && n <= 65535 && h >= 1 && w >= 1, "geometry");
  mask_boxes_init_kernel<<<(n + 63) / 64, 64, 0, st>>>(boxes, n, h, w);
  long long blocks = ((long long)h * w + 256 * 8 - 1) / (256 * 8);
  if (blocks > 148) blocks = 148;
  mask_boxes_kernel<<<dim3((unsigned)blocks, n), 256, 0, st>>>(masks, h, w, boxes);
  mask_boxes_final_kernel<<<(n + 63) / 64, 64, 0, st>>>(boxes, n);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

static int check_geom(const PlaceGeom& g) {
  DY_CHECK(g.src_h >= 1 && g.src_w >= 1 && g.new_w >= 1 && g.new_h >= 1 && g.size >= 1 && g.size <= 65535, "geometry");
  DY_CHECK(g.flip >= 1 && g.flip <= 3, "flip must be 1 (none), 2 (horizontal) or 3 (vertical)");
  return DY_OK;
}

int launch_place_image_u8(const unsigned char* rgb, const PlaceGeom& g, unsigned char* out, cudaStream_t st) {
  DY_TRY(check_geom(g));
  const double sx = 1.0 / ((double)g.new_w / (double)g.src_w), sy = 1.0 / ((double)g.new_h / (double)g.src_h);
  place_image_u8_kernel<<<dim3((g.size + 127) / 128, g.size), 128, 0, st>>>(rgb, g, sx, sy, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_place_masks(const unsigned char* masks, int n, const PlaceGeom& g, unsigned char* out, cudaStream_t st) {
  DY_TRY(check_geom(g));
  DY_CHECK(n >= 1 && n <= 65535, "mask count");
  const double sx = 1.0 / ((double)g.new_w / (double)g.src_w), sy = 1.0 / ((double)g.new_h / (double)g.src_h);
  place_masks_kernel<<<dim3((g.size + 127) / 128, g.size, n), 128, 0, st>>>(masks, g, sx, sy, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_salt_pepper(unsigned char* img, int size, const int* salt_rc, int n_salt, const int* pepper_rc, int n_pepper,
                       cudaStream_t st) {
  if (n_salt > 0) salt_pepper_kernel<<<(n_salt + 127) / 128, 128, 0, st>>>(img, size, salt_rc, n_salt, 1);
  if (n_pepper > 0) salt_pepper_kernel<<<(n_pepper + 127) / 128, 128, 0, st>>>(img, size, pepper_rc, n_pepper, 0);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_change_light(unsigned char* img, long long npix, double coeff, cudaStream_t st) {
  DY_CHECK(npix >= 1, "npix");
  change_light_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, st>>>(img, npix, coeff);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_motion_blur3(const unsigned char* img, int size, const float* kernel9, unsigned char* out, cudaStream_t st) {
  DY_CHECK(size >= 1 && size <= 65535 && kernel9 != nullptr, "geometry");
  Ker9 k;
  for (int i = 0; i < 9; ++i) k.k[i] = kernel9[i];
  motion_blur3_kernel<<<dim3((size + 127) / 128, size), 128, 0, st>>>(img, size, k, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_u8_div255_f32(const unsigned char* src, float* dst, long long n, cudaStream_t st) {
  long long blocks = (n + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  u8_div255_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

}  // namespace dy
