// Letterbox pre-processing and detection post-processing kernels (see imgproc.cuh).  Both restate
// cv2.resize(..., INTER_LINEAR) on float32 data: destination pixel d samples the source at
// f = (d + 0.5) * scale - 0.5 (double), s = floor(f), weight (float)(f - s), with weight 0 when s is clamped at
// either end; horizontal pass first, then vertical, in fp32.  Measured against cv2 4.13: boolean masks
// identical; letterboxed images within 3 ulp of the 0..255 value (cv2's SIMD kernels order / contract the
// two-tap sums differently).
#include "imgproc.cuh"

namespace dy {

namespace {

struct Tap {
  int i0, i1;
  float w0, w1;
};

// OpenCV 4.x: the source coordinate and its fractional part are formed in double, only the weight is
// rounded to float (recovered from cv2.resize with impulse inputs; forming the coordinate in float first
// costs 2^-14 of weight precision at x ~ 500)
__device__ __forceinline__ Tap linear_tap(int d, double scale, int n) {
  const double fd = ((double)d + 0.5) * scale - 0.5;
  int s = (int)floor(fd);
  float f = (float)(fd - (double)s);
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= n - 1) { s = n - 1; f = 0.f; }
  Tap t;
  t.i0 = s;
  t.i1 = s + 1 < n ? s + 1 : n - 1;
  t.w0 = 1.f - f;
  t.w1 = f;
  return t;
}

__device__ __forceinline__ float lerp2(float a, float b, float w0, float w1) {
  return __fadd_rn(__fmul_rn(a, w0), __fmul_rn(b, w1));
}

__global__ void letterbox_kernel(const unsigned char* __restrict__ rgb, long long rgb_stride, LetterboxGeom g,
                                 double scale_x, double scale_y, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= g.size) return;
  rgb += (size_t)blockIdx.z * rgb_stride;                        // blockIdx.z = image of a same-shape batch
  float* o = out + (((size_t)blockIdx.z * g.size + y) * g.size + x) * 3;
  const int dy_ = y - g.top, dx_ = x - g.left;
  if (dy_ < 0 || dy_ >= g.new_h || dx_ < 0 || dx_ >= g.new_w) {
    const float pad = (float)(127.0 / 255.0);
    o[0] = pad; o[1] = pad; o[2] = pad;
    return;
  }
  const Tap ty = linear_tap(dy_, scale_y, g.src_h), tx = linear_tap(dx_, scale_x, g.src_w);
  const unsigned char* r0 = rgb + (size_t)ty.i0 * g.src_w * 3;
  const unsigned char* r1 = rgb + (size_t)ty.i1 * g.src_w * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float h0 = lerp2((float)r0[tx.i0 * 3 + c], (float)r0[tx.i1 * 3 + c], tx.w0, tx.w1);
    const float h1 = lerp2((float)r1[tx.i0 * 3 + c], (float)r1[tx.i1 * 3 + c], tx.w0, tx.w1);
    const float v = lerp2(h0, h1, ty.w0, ty.w1);
    o[c] = (float)((double)v / 255.0);      // the reference divides its float64 canvas by 255.0 (:174)
  }
}

// correct_yolo_boxes (calculate_test_map.py:121-138) + crop indices (:246-251), one thread per detection
__global__ void post_prepare_kernel(const float* __restrict__ det_box, const int* __restrict__ count, int n_max, int S,
                                    int image_h, int image_w, int net, PostDet* __restrict__ ws,
                                    int* __restrict__ boxes_out, unsigned char* __restrict__ valid_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_max) return;
  {                                                              // blockIdx.y = image of a same-shape batch
    const size_t b = blockIdx.y;
    det_box += b * n_max * 6; count += b; ws += b * n_max; boxes_out += b * n_max * 4; valid_out += b * n_max;
  }
  PostDet d;
  memset(&d, 0, sizeof(d));
  if (k < count[0]) {
    int new_w, new_h;
    if ((double)net / image_w < (double)net / image_h) {
      new_w = net;
      new_h = (image_h * net) / image_w;
    } else {
      new_h = net;
      new_w = (image_w * net) / image_h;
    }
    const double x_off = (double)((net - new_w) / 2) / net, x_scale = (double)new_w / net;
    const double y_off = (double)((net - new_h) / 2) / net, y_scale = (double)new_h / net;
    const float y1n = det_box[k * 6 + 0], x1n = det_box[k * 6 + 1], y2n = det_box[k * 6 + 2], x2n = det_box[k * 6 + 3];
    // np.around = round half to even = rint
    auto corr = [](double v, double off, double sc, int extent) {
      int r = (int)rint((v - off) / sc * extent);
      r = r < extent ? r : extent;
      return r > 0 ? r : 0;
    };
    d.x1 = corr(x1n, x_off, x_scale, image_w);
    d.x2 = corr(x2n, x_off, x_scale, image_w);
    d.y1 = corr(y1n, y_off, y_scale, image_h);
    d.y2 = corr(y2n, y_off, y_scale, image_h);
    d.cls = (int)det_box[k * 6 + 4];
    const float Sf = (float)S;
    d.cy1 = (int)rintf(__fmul_rn(y1n, Sf)); d.cx1 = (int)rintf(__fmul_rn(x1n, Sf));
    d.cy2 = (int)rintf(__fmul_rn(y2n, Sf)); d.cx2 = (int)rintf(__fmul_rn(x2n, Sf));
    // numpy slicing clips to the array
    d.cy1 = max(0, min(d.cy1, S)); d.cy2 = max(0, min(d.cy2, S));
    d.cx1 = max(0, min(d.cx1, S)); d.cx2 = max(0, min(d.cx2, S));
    const int bw = d.x2 - d.x1, bh = d.y2 - d.y1, cw = d.cx2 - d.cx1, ch = d.cy2 - d.cy1;
    d.valid = ((long long)bw * bh > 0 && cw > 0 && ch > 0) ? 1 : 0;
    if (d.valid) {
      // cv::resize: inv_scale = dsize / ssize; scale = 1 / inv_scale
      d.scale_x = 1.0 / ((double)bw / (double)cw);
      d.scale_y = 1.0 / ((double)bh / (double)ch);
    }
  }
  ws[k] = d;
  boxes_out[k * 4 + 0] = d.x1; boxes_out[k * 4 + 1] = d.y1; boxes_out[k * 4 + 2] = d.x2; boxes_out[k * 4 + 3] = d.y2;
  valid_out[k] = (unsigned char)d.valid;
}

// one thread per pixel of the original image: every detection's boolean mask at that pixel, and the
// merged semantic mask (class + 1 of the LAST detection covering it: the reference overwrites in order).
// A block covers 128 pixels of one row: the detections are staged in shared memory and warp 0 compacts
// the ones whose box meets this row segment, so a pixel only visits boxes that can contain it.
constexpr int kPostMaxSmemDet = 64;

__global__ void __launch_bounds__(128)
post_pixel_kernel(const PostDet* __restrict__ ws, const int* __restrict__ count, int n_max,
                  const float* __restrict__ masks, int S, int image_h, int image_w,
                  unsigned char* __restrict__ full_masks, unsigned char* __restrict__ merged) {
  __shared__ PostDet sd[kPostMaxSmemDet];
  __shared__ int list[kPostMaxSmemDet];
  __shared__ int nlist;
  {                                                              // blockIdx.z = image of a same-shape batch
    const size_t b = blockIdx.z, plane_b = (size_t)image_h * image_w;
    ws += b * n_max; count += b; masks += b * n_max * S * S;
    if (full_masks) full_masks += b * n_max * plane_b;
    if (merged) merged += b * plane_b;
  }
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int xb0 = blockIdx.x * blockDim.x, xb1 = min(xb0 + (int)blockDim.x, image_w);
  const int n = min(count[0], n_max);
  const bool staged = n <= kPostMaxSmemDet;
  if (staged) {
    if (threadIdx.x < 32) {
      int base = 0;
      for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + threadIdx.x;
        bool hit = false;
        if (k < n) {
          const PostDet d = ws[k];
          hit = d.valid && y >= d.y1 && y < d.y2 && d.x1 < xb1 && d.x2 > xb0;
          if (hit) sd[k] = d;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) list[base + __popc(m & ((1u << threadIdx.x) - 1u))] = k;     // ascending k: "last wins" order kept
        base += __popc(m);
      }
      if (threadIdx.x == 0) nlist = base;
    }
    __syncthreads();
  }
  if (x >= image_w) return;
  const size_t pix = (size_t)y * image_w + x, plane = (size_t)image_h * image_w;
  unsigned char m = 0;
  auto sample = [&](const PostDet& d, int k) -> unsigned char {
    const int cw = d.cx2 - d.cx1, ch = d.cy2 - d.cy1;
    const Tap ty = linear_tap(y - d.y1, d.scale_y, ch), tx = linear_tap(x - d.x1, d.scale_x, cw);
    const float* base = masks + (size_t)k * S * S;
    const float* r0 = base + (size_t)(d.cy1 + ty.i0) * S + d.cx1;
    const float* r1 = base + (size_t)(d.cy1 + ty.i1) * S + d.cx1;
    const float h0 = lerp2(__ldg(r0 + tx.i0), __ldg(r0 + tx.i1), tx.w0, tx.w1);
    const float h1 = lerp2(__ldg(r1 + tx.i0), __ldg(r1 + tx.i1), tx.w0, tx.w1);
    return lerp2(h0, h1, ty.w0, ty.w1) > 0.5f ? 1 : 0;
  };
  if (staged) {
    if (full_masks)
      for (int k = 0; k < n; ++k) full_masks[(size_t)k * plane + pix] = 0;
    for (int i = 0; i < nlist; ++i) {
      const int k = list[i];
      const PostDet& d = sd[k];
      if (x >= d.x1 && x < d.x2) {
        const unsigned char v = sample(d, k);
        if (v) {
          m = (unsigned char)(d.cls + 1);
          if (full_masks) full_masks[(size_t)k * plane + pix] = 1;
        }
      }
    }
  } else {
    for (int k = 0; k < n; ++k) {
      const PostDet d = ws[k];
      unsigned char v = 0;
      if (d.valid && x >= d.x1 && x < d.x2 && y >= d.y1 && y < d.y2) {
        v = sample(d, k);
        if (v) m = (unsigned char)(d.cls + 1);
      }
      if (full_masks) full_masks[(size_t)k * plane + pix] = v;
    }
  }
  if (merged) merged[pix] = m;
}

// ---- mask IoU matrix: compute_overlaps_masks (utils/voc_eval_mask.py:38-56) ----------------------------
// A block owns 4096 pixels: every mask's bytes are turned into bits with warp ballots (one coalesced
// 32-byte read per ballot, no alignment requirement), then each (i, j) pair is 128 AND + POPC.
constexpr int kOvChunk = 4096, kOvWords = kOvChunk / 32;

__global__ void __launch_bounds__(256)
overlaps_count_kernel(const unsigned char* __restrict__ m1, int n1, const unsigned char* __restrict__ m2, int n2,
                      long long P, int* __restrict__ inter, int* __restrict__ area) {
  extern __shared__ uint32_t bits[];                       // [(n1+n2)][kOvWords]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p0 = (long long)blockIdx.x * kOvChunk;
  const int nm = n1 + n2;
  for (int t = warp; t < nm * kOvWords; t += 8) {
    const int r = t / kOvWords, w = t - r * kOvWords;
    const unsigned char* src = r < n1 ? m1 + (long long)r * P : m2 + (long long)(r - n1) * P;
    const long long px = p0 + w * 32 + lane;
    const bool on = px < P && __ldg(src + px) != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) bits[t] = word;
  }
  __syncthreads();
  for (int pr = threadIdx.x; pr < n1 * n2; pr += 256) {
    const int i = pr / n2, j = pr - i * n2;
    const uint32_t* a = bits + i * kOvWords;
    const uint32_t* b = bits + (n1 + j) * kOvWords;
    int s = 0;
#pragma unroll 8
    for (int w = 0; w < kOvWords; ++w) s += __popc(a[w] & b[w]);
    if (s) atomicAdd(inter + pr, s);
  }
  for (int r = threadIdx.x; r < nm; r += 256) {
    int s = 0;
    for (int w = 0; w < kOvWords; ++w) s += __popc(bits[r * kOvWords + w]);
    if (s) atomicAdd(area + r, s);
  }
}

__global__ void overlaps_finalize_kernel(const int* __restrict__ inter, const int* __restrict__ area, int n1, int n2,
                                         float* __restrict__ out) {
  const int pr = blockIdx.x * blockDim.x + threadIdx.x;
  if (pr >= n1 * n2) return;
  const int i = pr / n2, j = pr - i * n2;
  const float it = (float)inter[pr], a1 = (float)area[i], a2 = (float)area[n1 + j];
  out[pr] = __fdiv_rn(it, __fsub_rn(__fadd_rn(a1, a2), it));      // 0/0 = NaN, like NumPy
}

// ---- training label assignment: utils/train_data.py:134-178 (+ flips :189-228, normalisation :258-262) -------
// One thread per image (at most 20 boxes, and "the first box wins a cell" makes the loop sequential).
__global__ void assign_labels_kernel(LabelArgs a) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int depth = 5 + a.num_class, net = a.net;
  const double sx = a.place[b * 4 + 0], sy = a.place[b * 4 + 1], dx = a.place[b * 4 + 2], dy = a.place[b * 4 + 3];
  const int flip = a.flip ? a.flip[b] : 1;
  const int nb = min(a.nbox[b], a.max_box);
  const float netm1 = (float)(net - 1), size_f = (float)net;
  for (int i = 0; i < nb; ++i) {
    const float* bx = a.boxes + ((long long)b * a.max_box + i) * 5;
    const int cls = (int)bx[4];
    auto clampd = [&](double v) { return fmax(fmin(v, (double)(net - 1)), 0.0); };
    const double x1 = clampd((double)bx[0] * sx + dx), y1 = clampd((double)bx[1] * sy + dy);
    const double x2 = clampd((double)bx[2] * sx + dx), y2 = clampd((double)bx[3] * sy + dy);
    const double xc = (x2 + x1) / 2.0, yc = (y2 + y1) / 2.0, w = x2 - x1, h = y2 - y1;
    float fx = (float)xc, fy = (float)yc;
    const float fw = (float)w, fh = (float)h;
    // best anchor by IoU of the centred (w, h) boxes, fp32 like the reference's np.float32 arrays
    const float hw = (float)(w / 2.0), hh = (float)(h / 2.0);
    const float box_area = __fmul_rn(__fmul_rn(hw, hh), 4.f);
    float best = -1.f;
    int besti = 0;
    for (int k = 0; k < 9; ++k) {
      const float aw = __fdiv_rn(a.anchors[2 * k], 2.f), ah = __fdiv_rn(a.anchors[2 * k + 1], 2.f);
      const float anc_area = __fmul_rn(__fmul_rn(ah, aw), 4.f);
      const float iw = fmaxf(__fsub_rn(fminf(hw, aw), fmaxf(-hw, -aw)), 0.f);
      const float ih = fmaxf(__fsub_rn(fminf(hh, ah), fmaxf(-hh, -ah)), 0.f);
      const float inter = __fmul_rn(iw, ih);
      const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(box_area, anc_area), inter));
      if (iou > best) { best = iou; besti = k; }        // first maximum, like np.argmax
    }
    if (best > 0.f) {
      const int s = besti / 3, an = besti % 3;
      const int g = a.grid[s];
      int x_ind = (int)(xc * g / net), y_ind = (int)(yc * g / net);
      float lx = fx, ly = fy;
      if (flip == 2) { x_ind = g - 1 - x_ind; lx = __fsub_rn(netm1, fx); }
      if (flip == 3) { y_ind = g - 1 - y_ind; ly = __fsub_rn(netm1, fy); }
      float* cell = a.yolo[s] + ((((long long)b * g + y_ind) * g + x_ind) * 3 + an) * depth;
      if (cell[4] != 1.f) {                              // an earlier box owns this cell: skipped (:166-167)
        cell[0] = __fdiv_rn(lx, size_f); cell[1] = __fdiv_rn(ly, size_f);
        cell[2] = __fdiv_rn(fw, size_f); cell[3] = __fdiv_rn(fh, size_f);
        cell[4] = 1.f;
        cell[5 + cls] = 1.f;
      }
    }
    if (flip == 2) fx = __fsub_rn(netm1, fx);
    if (flip == 3) fy = __fsub_rn(netm1, fy);
    float* tb = a.true_boxes + ((long long)b * a.max_box + i) * 5;
    tb[0] = __fdiv_rn(fx, size_f); tb[1] = __fdiv_rn(fy, size_f);
    tb[2] = __fdiv_rn(fw, size_f); tb[3] = __fdiv_rn(fh, size_f);
    tb[4] = (float)cls;
  }
}

// ------------------------------------------------------------------------------------------
// uint8 RGB -> the network's fp32 input: v / 255 exactly as image_read forms it (float64 division of
// the 0..255 value, then the float32 cast of the feed; calculate_test_map.py:175).  Only 256 inputs
// exist, so the quotients are formed once in double (g_u8_lut) and looked up from shared memory.
// ------------------------------------------------------------------------------------------
__device__ float g_u8_lut[256];

__global__ void u8_lut_init_kernel() { g_u8_lut[threadIdx.x] = (float)((double)threadIdx.x / 255.0); }

__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint4* __restrict__ src, float4* __restrict__ dst,
                                                        long long n16, const unsigned char* __restrict__ tail_src,
                                                        float* __restrict__ tail_dst, int ntail) {
  __shared__ float lut[256];
  lut[threadIdx.x] = g_u8_lut[threadIdx.x];
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(src + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      __stcs(dst + 4 * i + j, make_float4(lut[w[j] & 0xFF], lut[(w[j] >> 8) & 0xFF], lut[(w[j] >> 16) & 0xFF],
                                          lut[w[j] >> 24]));
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail_dst[threadIdx.x] = lut[tail_src[threadIdx.x]];
}

}  // namespace

int launch_assign_labels(const LabelArgs& a, cudaStream_t st) {
  DY_CHECK(a.B >= 1 && a.max_box >= 1 && a.num_class >= 1 && a.net >= 32, "label geometry");
  const int depth = 5 + a.num_class;
  for (int s = 0; s < 3; ++s)
    DY_CUDA(cudaMemsetAsync(a.yolo[s], 0, (size_t)a.B * a.grid[s] * a.grid[s] * 3 * depth * 4, st));
  DY_CUDA(cudaMemsetAsync(a.true_boxes, 0, (size_t)a.B * a.max_box * 5 * 4, st));
  assign_labels_kernel<<<(a.B + 63) / 64, 64, 0, st>>>(a);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_mask_overlaps(const unsigned char* m1, int n1, const unsigned char* m2, int n2, long long P, int* ws,
                         float* out, cudaStream_t st) {
  DY_CHECK(n1 >= 1 && n2 >= 1 && P >= 1, "empty mask set");
  DY_CHECK(n1 + n2 <= 384, "at most 384 masks per call");
  const size_t smem = (size_t)(n1 + n2) * kOvWords * 4;
  static bool attr = false;
  if (!attr) {
    DY_CUDA(cudaFuncSetAttribute(overlaps_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  DY_CUDA(cudaMemsetAsync(ws, 0, (size_t)(n1 * n2 + n1 + n2) * 4, st));
  const long long blocks = (P + kOvChunk - 1) / kOvChunk;
  DY_CHECK(blocks < (1ll << 31), "mask too large");
  overlaps_count_kernel<<<(int)blocks, 256, smem, st>>>(m1, n1, m2, n2, P, ws, ws + n1 * n2);
  DY_CUDA(cudaGetLastError());
  overlaps_finalize_kernel<<<(n1 * n2 + 127) / 128, 128, 0, st>>>(ws, ws + n1 * n2, n1, n2, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

LetterboxGeom letterbox_geom(int src_h, int src_w, int size) {
  LetterboxGeom g;
  g.src_h = src_h; g.src_w = src_w; g.size = size;
  int h = src_h, w = src_w;
  if ((double)size / w < (double)size / h) {     // :152-157
    h = (int)(((long long)h * size) / w);
    w = size;
  } else {
    w = (int)(((long long)w * size) / h);
    h = size;
  }
  g.new_h = h; g.new_w = w;
  g.top = (size - h) / 2;
  g.left = (size - w) / 2;
  return g;
}

int launch_u8_to_f32(const unsigned char* src, float* dst, long long n, cudaStream_t st) {
  DY_CHECK((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "alignment");
  static bool lut_ready[64] = {false};
  int dev = 0;
  DY_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !lut_ready[dev]) {
    u8_lut_init_kernel<<<1, 256, 0, st>>>();
    DY_CUDA(cudaGetLastError());
    lut_ready[dev] = true;
  } else if (dev >= 64) {
    u8_lut_init_kernel<<<1, 256, 0, st>>>();
  }
  const long long n16 = n / 16;
  const int ntail = (int)(n - n16 * 16);
  long long blocks = (n16 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  u8_to_f32_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<float4*>(dst),
                                               n16, src + n16 * 16, dst + n16 * 16, ntail);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_letterbox(const unsigned char* rgb, long long rgb_stride, int B, const LetterboxGeom& g, float* out,
                     cudaStream_t st) {
  DY_CHECK(g.src_h > 0 && g.src_w > 0 && g.new_h > 0 && g.new_w > 0 && g.size > 0, "image geometry");
  DY_CHECK(B >= 1 && B <= 65535 && g.size <= 65535, "batch / grid limits");
  const double scale_x = 1.0 / ((double)g.new_w / (double)g.src_w), scale_y = 1.0 / ((double)g.new_h / (double)g.src_h);
  dim3 grid((g.size + 127) / 128, g.size, B);
  letterbox_kernel<<<grid, 128, 0, st>>>(rgb, rgb_stride, g, scale_x, scale_y, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_postprocess(const float* det_box, const int* count, int B, int n_max, const float* masks, int S,
                       int image_h, int image_w, int net_size, PostDet* ws, int* boxes_out, unsigned char* valid_out,
                       unsigned char* full_masks, unsigned char* merged, cudaStream_t st) {
  DY_CHECK(n_max >= 1 && S >= 1 && image_h >= 1 && image_w >= 1 && net_size >= 1, "geometry");
  DY_CHECK(B >= 1 && B <= 65535 && image_h <= 65535, "batch / grid limits");
  post_prepare_kernel<<<dim3((n_max + 63) / 64, B), 64, 0, st>>>(det_box, count, n_max, S, image_h, image_w, net_size,
                                                                 ws, boxes_out, valid_out);
  DY_CUDA(cudaGetLastError());
  if (full_masks || merged) {
    dim3 grid((image_w + 127) / 128, image_h, B);
    post_pixel_kernel<<<grid, 128, 0, st>>>(ws, count, n_max, masks, S, image_h, image_w, full_masks, merged);
    DY_CUDA(cudaGetLastError());
  }
  return DY_OK;
}

}  // namespace dy
