// Training-step kernels, fp32 (see train.cuh): the fused loss + gradient kernels, Adam, the multi-tensor
// helpers (shared by both engines) and the CUDA-core conv backward of the fp32 VERIFICATION engine.  The
// tensor-core dgrad / wgrad / BN kernels of the bf16 engine live in train_tc.cu.
#include "train.cuh"

namespace dy {

namespace {

constexpr int kT = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static int blocks_for(long long total, int per_block = kT, int cap = 148 * 8) {
  long long g = (total + per_block - 1) / per_block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------------------------
// BatchNorm
// ---------------------------------------------------------------------------------------------
// thread (tx, ty): tx walks channels, ty walks rows; double accumulation, one atomic per (thread, channel)
template <bool BWD>
__global__ void __launch_bounds__(kT)
bn_reduce_kernel(const float* __restrict__ u, const float* __restrict__ z, const float* __restrict__ a,
                 const float* __restrict__ b, const float* __restrict__ mean, const float* __restrict__ invstd,
                 float alpha, int act, long long M, int C, int rows_per_block, double* __restrict__ o1,
                 double* __restrict__ o2) {
  const int cw = C < kT ? C : kT;            // channels covered per pass (C is a multiple of 32 or small)
  const int tx = threadIdx.x % cw, ty = threadIdx.x / cw, R = kT / cw;
  if (ty >= R) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  for (int c = tx; c < C; c += cw) {
    double s1 = 0.0, s2 = 0.0;
    if (BWD) {
      const float ac = a[c], bc = b[c], mc = mean[c], ic = invstd[c];
      for (long long r = r0 + ty; r < r1; r += R) {
        const float zz = z[r * C + c];
        float g = u[r * C + c];
        if (act && !(fmaf(zz, ac, bc) > 0.f)) g *= alpha;
        s1 += (double)g;
        s2 += (double)g * (double)((zz - mc) * ic);
      }
    } else {
      for (long long r = r0 + ty; r < r1; r += R) {
        const double v = (double)u[r * C + c];
        s1 += v;
        s2 += v * v;
      }
    }
    atomicAdd(o1 + c, s1);
    if (o2) atomicAdd(o2 + c, s2);
  }
}

__global__ void bn_finalize_kernel(const double* sum, const double* sumsq, long long M, int C, const float* gamma,
                                   const float* beta, float eps, float* a, float* b, float* mean, float* var,
                                   float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sum[c] / (double)M;
  double v = sumsq[c] / (double)M - m * m;
  if (v < 0.0) v = 0.0;
  const float is = (float)(1.0 / sqrt(v + (double)eps));
  mean[c] = (float)m;
  var[c] = (float)v;
  invstd[c] = is;
  const float aa = gamma[c] * is;
  a[c] = aa;
  b[c] = beta[c] - (float)m * aa;
}

__global__ void bn_act_kernel(const float* __restrict__ z, const float* __restrict__ a, const float* __restrict__ b,
                              int C, long long total, float alpha, int act, const float* __restrict__ residual,
                              float* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float v = fmaf(z[i], a[c], b[c]);
    if (act) v = fmaxf(alpha * v, v);
    if (residual) v += residual[i];
    y[i] = v;
  }
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                    const float* __restrict__ a, const float* __restrict__ b,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const double* __restrict__ s1,
                                    const double* __restrict__ s2, float alpha, int act, int mode, long long M, int C,
                                    float* __restrict__ dz) {
  const long long total = M * C;
  const float invM = 1.f / (float)M;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (mode == 2) {
      dz[i] = dy[i];
      continue;
    }
    const int c = (int)(i % C);
    const float zz = z[i];
    float g = dy[i];
    if (act && !(fmaf(zz, a[c], b[c]) > 0.f)) g *= alpha;
    if (mode == 1) {
      dz[i] = g * a[c];
    } else {
      const float xh = (zz - mean[c]) * invstd[c];
      dz[i] = gamma[c] * invstd[c] * (g - (float)s1[c] * invM - xh * (float)s2[c] * invM);
    }
  }
}

__global__ void copy_stats_kernel(const double* s1, const double* s2, int C, float* dgamma, float* dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = (float)s1[c];
  if (dgamma) dgamma[c] = (float)s2[c];
}

// ---------------------------------------------------------------------------------------------
// conv backward
// ---------------------------------------------------------------------------------------------
__global__ void weight_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int kk, int cin,
                                        int cout) {
  const long long total = (long long)kk * cin * cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout);
    const long long t = i / cout;
    const int ci = (int)(t % cin);
    const int tap = (int)(t / cin);
    wt[((long long)tap * cout + co) * cin + ci] = w[i];
  }
}

constexpr int kDgPix = 4;
// dx[n,iy,ix,ci] = sum_{kh,kw,co} dz[n,oy,ox,co] * w[kh,kw,ci,co],  iy = oy*s + kh - pad_t
__global__ void __launch_bounds__(256) conv_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ wt,
                                                         float* __restrict__ dx, ConvGeom g, int accumulate) {
  const int ci = blockIdx.x * blockDim.x + threadIdx.x;
  const int xg = (blockIdx.y * blockDim.y + threadIdx.y) * kDgPix;
  const int iy = blockIdx.z % g.Hi, n = blockIdx.z / g.Hi;
  if (ci >= g.cin || xg >= g.Wi) return;
  float acc[kDgPix];
#pragma unroll
  for (int i = 0; i < kDgPix; ++i) acc[i] = 0.f;
  for (int kh = 0; kh < g.k; ++kh) {
    const int oyn = iy + g.pad_t - kh;
    if (oyn < 0 || oyn % g.s != 0) continue;
    const int oy = oyn / g.s;
    if (oy >= g.Ho) continue;
    for (int kw = 0; kw < g.k; ++kw) {
      int ox[kDgPix];
      bool ok[kDgPix];
      bool any = false;
#pragma unroll
      for (int i = 0; i < kDgPix; ++i) {
        const int oxn = xg + i + g.pad_l - kw;
        ok[i] = (xg + i < g.Wi) && oxn >= 0 && (oxn % g.s == 0) && (oxn / g.s < g.Wo);
        ox[i] = ok[i] ? oxn / g.s : 0;
        any |= ok[i];
      }
      if (!any) continue;
      const float* wp = wt + ((long long)(kh * g.k + kw) * g.cout) * g.cin + ci;
      const float* zr = dz + ((long long)n * g.Ho + oy) * g.Wo * g.cout;
      for (int co = 0; co < g.cout; ++co) {
        const float wv = __ldg(wp + (long long)co * g.cin);
#pragma unroll
        for (int i = 0; i < kDgPix; ++i)
          if (ok[i]) acc[i] = fmaf(__ldg(zr + (long long)ox[i] * g.cout + co), wv, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kDgPix; ++i) {
    if (xg + i >= g.Wi) break;
    const long long o = (((long long)n * g.Hi + iy) * g.Wi + xg + i) * g.cin + ci;
    dx[o] = accumulate ? dx[o] + acc[i] : acc[i];
  }
}

// dw[tap][ci][co] += sum_pix x[pix@tap][ci] * dz[pix][co]; 64x64 tile per block, split over pixels
constexpr int kWgT = 64, kWgP = 16;
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ x0, int c0, const float* __restrict__ x1, int c1,
                  const float* __restrict__ dz, float* __restrict__ dw, ConvGeom g, long long pix_per_split) {
  __shared__ float xs[kWgP][kWgT + 1];
  __shared__ float zs[kWgP][kWgT + 1];
  const int cin = c0 + c1;
  const int ci_tiles = (cin + kWgT - 1) / kWgT;
  const int tap = blockIdx.y / ci_tiles, cit = blockIdx.y % ci_tiles;
  const int kh = tap / g.k, kw = tap % g.k;
  const int co0 = blockIdx.x * kWgT, ci0 = cit * kWgT;
  const long long npix = (long long)g.B * g.Ho * g.Wo;
  const long long p0 = (long long)blockIdx.z * pix_per_split;
  long long p1 = p0 + pix_per_split;
  if (p1 > npix) p1 = npix;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;     // thread computes ci = ci0+ty*4.., co = co0+tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lp = threadIdx.x / kWgT;          // 0..3: pixel sub-row this thread loads
  const int lc = threadIdx.x % kWgT;          // channel this thread loads
  for (long long pb = p0; pb < p1; pb += kWgP) {
#pragma unroll
    for (int r = 0; r < kWgP / 4; ++r) {
      const int pr = lp + 4 * r;
      const long long p = pb + pr;
      float xv = 0.f, zv = 0.f;
      if (p < p1) {
        const int ox = (int)(p % g.Wo);
        const long long t = p / g.Wo;
        const int oy = (int)(t % g.Ho), n = (int)(t / g.Ho);
        if (co0 + lc < g.cout) zv = __ldg(dz + p * g.cout + co0 + lc);
        const int iy = oy * g.s + kh - g.pad_t, ix = ox * g.s + kw - g.pad_l;
        const int ci = ci0 + lc;
        if (ci < cin && iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi) {
          if (ci < c0) xv = __ldg(x0 + (((long long)n * g.Hi + iy) * g.Wi + ix) * c0 + ci);
          else xv = __ldg(x1 + (((long long)n * (g.Hi / 2) + (iy >> 1)) * (g.Wi / 2) + (ix >> 1)) * c1 + (ci - c0));
        }
      }
      xs[pr][lc] = xv;
      zs[pr][lc] = zv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < kWgP; ++pp) {
      float xv[4], zv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xv[i] = xs[pp][ty * 4 + i];
        zv[i] = zs[pp][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], zv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= cin) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < g.cout) atomicAdd(dw + ((long long)tap * cin + ci) * g.cout + co, acc[i][j]);
    }
  }
}

__global__ void split_accumulate_kernel(const float* __restrict__ src, int B, int H, int W, int c0, int c1,
                                        float* __restrict__ dst0, float* __restrict__ dst1) {
  const int ct = c0 + c1;
  if (dst0) {
    const long long total = (long long)B * H * W * c0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(i % c0);
      const long long p = i / c0;
      dst0[i] += src[p * ct + c];
    }
  }
  if (dst1) {
    const int Hh = H / 2, Wh = W / 2;
    const long long total = (long long)B * Hh * Wh * c1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int c = (int)(i % c1);
      long long t = i / c1;
      const int x = (int)(t % Wh);
      t /= Wh;
      const int y = (int)(t % Hh);
      const int n = (int)(t / Hh);
      const long long b00 = (((long long)n * H + 2 * y) * W + 2 * x) * ct + c0 + c;
      dst1[i] += src[b00] + src[b00 + ct] + src[b00 + (long long)W * ct] + src[b00 + (long long)W * ct + ct];
    }
  }
}

__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] += src[i];
}

// ---------------------------------------------------------------------------------------------
// YOLO loss (:631-747): one thread per (image, scale, cell, anchor); loss terms and d loss / d logits
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) yolo_loss_kernel(YoloLossArgs a) {
  __shared__ double red[5][kT / 32];
  const int n0 = 3 * (a.g[0] * a.g[0] + a.g[1] * a.g[1] + a.g[2] * a.g[2]);
  const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double part[5] = {0, 0, 0, 0, 0};
  if (gid < (long long)a.B * n0) {
    const int b = (int)(gid / n0), idx = (int)(gid % n0);
    const int off1 = 3 * a.g[0] * a.g[0], off2 = off1 + 3 * a.g[1] * a.g[1];
    const int j = idx < off1 ? 0 : (idx < off2 ? 1 : 2);
    const int local = idx - (j == 0 ? 0 : (j == 1 ? off1 : off2));
    const int g = a.g[j], an = local % 3, cell = local / 3, cy = cell / g, cx = cell % g;
    const long long e = ((((long long)b * g + cy) * g + cx) * 3 + an) * 8;
    const float* p = a.pred[j] + e;
    const float* lab = a.label[j] + e;
    float* d = a.dpred[j] + e;
    const float invB = 1.f / (float)a.B;
    const float aw = a.anchors[(3 * j + an) * 2], ah = a.anchors[(3 * j + an) * 2 + 1];
    const float sx = 1.f / (1.f + expf(-p[0])), sy = 1.f / (1.f + expf(-p[1]));
    const float bx = ((float)cx + sx) / (float)g, by = ((float)cy + sy) / (float)g;
    const float bw = expf(p[2]) * aw / (float)a.net, bh = expf(p[3]) * ah / (float)a.net;
    // best IoU with the true boxes -> ignore mask (:657-680)
    float best = 0.f;
    const float* tb = a.true_boxes + (long long)b * 20 * 5;
    for (int t = 0; t < 20; ++t) {
      const float tx = tb[t * 5], ty = tb[t * 5 + 1], tw = tb[t * 5 + 2], th = tb[t * 5 + 3];
      const float iw = fmaxf(fminf(bx + bw / 2.f, tx + tw / 2.f) - fmaxf(bx - bw / 2.f, tx - tw / 2.f), 0.f);
      const float ih = fmaxf(fminf(by + bh / 2.f, ty + th / 2.f) - fmaxf(by - bh / 2.f, ty - th / 2.f), 0.f);
      const float inter = iw * ih;
      const float uni = fmaxf(bw * bh + tw * th - inter, 1e-10f);
      best = fmaxf(best, fminf(fmaxf(inter / uni, 0.f), 1.f));
    }
    const float ignore = best < a.ignore_thresh ? 1.f : 0.f;
    const float obj = lab[4], noobj = 1.f - obj;
    // confidence: sigmoid cross entropy with label = object mask (:686-695)
    const float l = p[4];
    const float bce = fmaxf(l, 0.f) - l * obj + log1pf(expf(-fabsf(l)));
    const float sig = 1.f / (1.f + expf(-l));
    const float wconf = obj * a.object_scale + ignore * noobj * a.noobject_scale;
    part[0] = (double)(obj * bce * a.object_scale);
    part[1] = (double)(ignore * noobj * bce * a.noobject_scale);
    d[4] = wconf * (sig - obj) * invB;
    // class: sparse softmax cross entropy on object cells (:698-703)
    float mx = fmaxf(p[5], fmaxf(p[6], p[7]));
    const float e0 = expf(p[5] - mx), e1 = expf(p[6] - mx), e2 = expf(p[7] - mx), es = e0 + e1 + e2;
    int tc = 0;
    if (lab[6] > lab[5]) tc = 1;
    if (lab[7] > lab[5 + tc]) tc = 2;
    const float ce = logf(es) - (p[5 + tc] - mx);
    part[2] = (double)(obj * ce * a.class_scale);
    const float wc = obj * a.class_scale * invB;
    d[5] = wc * (e0 / es - (tc == 0 ? 1.f : 0.f));
    d[6] = wc * (e1 / es - (tc == 1 ? 1.f : 0.f));
    d[7] = wc * (e2 / es - (tc == 2 ? 1.f : 0.f));
    // coordinates (:706-727)
    const float tcx = lab[0] * (float)g - (float)cx, tcy = lab[1] * (float)g - (float)cy;
    const float ttw = fminf(fmaxf(logf(lab[2] * (float)a.net / aw), -100.f), 100.f);
    const float tth = fminf(fmaxf(logf(lab[3] * (float)a.net / ah), -100.f), 100.f);
    const float ws = 2.f - lab[2] * lab[3], ws2 = ws * ws * a.coord_scale;
    if (obj != 0.f) {
      const float dx_ = obj * (sx - tcx), dy_ = obj * (sy - tcy), dw_ = obj * (p[2] - ttw), dh_ = obj * (p[3] - tth);
      part[3] = (double)((dx_ * dx_ + dy_ * dy_) * ws2);
      part[4] = (double)((dw_ * dw_ + dh_ * dh_) * ws2);
      d[0] = 2.f * dx_ * obj * ws2 * sx * (1.f - sx) * invB;
      d[1] = 2.f * dy_ * obj * ws2 * sy * (1.f - sy) * invB;
      d[2] = 2.f * dw_ * obj * ws2 * invB;
      d[3] = 2.f * dh_ * obj * ws2 * invB;
    } else {
      d[0] = d[1] = d[2] = d[3] = 0.f;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double s = warp_sum(part[k]);
    if (lane == 0) red[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double s = 0;
    for (int w = 0; w < kT / 32; ++w) s += red[threadIdx.x][w];
    atomicAdd(a.loss + threadIdx.x, s / (double)a.B);      // mean over the batch of per-image sums
  }
}

// ---------------------------------------------------------------------------------------------
// mask loss (:750-860)
// ---------------------------------------------------------------------------------------------
// steps 1-4: RoIs = 7 shuffled proposals + 3 shuffled GT boxes, positives = IoU >= thr with some GT
__global__ void mask_roi_kernel(MaskLossArgs a) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  const float* det = a.det + (long long)b * a.max_det * 6;
  const float* tb = a.true_boxes + (long long)b * 20 * 5;
  // trimmed lists (rows with a non-zero box)
  int gt_ids[20], ngt = 0;
  float gtb[20][4];
  for (int t = 0; t < 20; ++t) {
    const float xc = tb[t * 5], yc = tb[t * 5 + 1], w = tb[t * 5 + 2], h = tb[t * 5 + 3];
    if (fabsf(xc) + fabsf(yc) + fabsf(w) + fabsf(h) > 0.f) {
      gtb[ngt][0] = yc - h / 2.f; gtb[ngt][1] = xc - w / 2.f; gtb[ngt][2] = yc + h / 2.f; gtb[ngt][3] = xc + w / 2.f;
      gt_ids[ngt++] = t;
    }
  }
  int nprop = 0;
  for (int r = 0; r < a.max_det; ++r) {
    const float* q = det + r * 6;
    if (fabsf(q[0]) + fabsf(q[1]) + fabsf(q[2]) + fabsf(q[3]) > 0.f) ++nprop;   // zero padding is at the end
  }
  float rois[10][4];
  int nroi = 0, taken = 0;
  for (int i = 0; i < a.max_det && taken < 7; ++i) {
    const int pi = a.perm_prop[b * a.max_det + i];
    if (pi < nprop) {
      // the pi-th non-zero proposal == row pi (non-zero rows form a prefix)
      for (int c = 0; c < 4; ++c) rois[nroi][c] = det[pi * 6 + c];
      ++nroi; ++taken;
    }
  }
  taken = 0;
  for (int i = 0; i < 20 && taken < 3; ++i) {
    const int gi = a.perm_gt[b * 20 + i];
    if (gi < ngt) {
      for (int c = 0; c < 4; ++c) rois[nroi][c] = gtb[gi][c];
      ++nroi; ++taken;
    }
  }
  int np = 0;
  if (ngt > 0) {
    for (int r = 0; r < nroi; ++r) {
      float best = -1.f;
      int arg = 0;
      for (int t = 0; t < ngt; ++t) {
        const float y1 = fmaxf(rois[r][0], gtb[t][0]), x1 = fmaxf(rois[r][1], gtb[t][1]);
        const float y2 = fminf(rois[r][2], gtb[t][2]), x2 = fminf(rois[r][3], gtb[t][3]);
        const float inter = fmaxf(x2 - x1, 0.f) * fmaxf(y2 - y1, 0.f);
        const float a1 = (rois[r][2] - rois[r][0]) * (rois[r][3] - rois[r][1]);
        const float a2 = (gtb[t][2] - gtb[t][0]) * (gtb[t][3] - gtb[t][1]);
        float iou = inter / (a1 + a2 - inter);
        if (!(iou == iou)) iou = -1.f;              // 0/0
        if (iou > best) { best = iou; arg = t; }
      }
      if (best >= a.iou_thresh) {
        for (int c = 0; c < 4; ++c) a.rois[((long long)b * 10 + np) * 4 + c] = rois[r][c];
        a.assign[b * 10 + np] = gt_ids[arg];
        ++np;
      }
    }
  }
  a.npos[b] = np;
}

// one CTA per (roi, image): BCE over the box pixels of the assembled mask, gradient into dmask
__global__ void __launch_bounds__(kT) mask_loss_kernel(MaskLossArgs a) {
  __shared__ double red[kT / 32];
  const int r = blockIdx.x, b = blockIdx.y;
  const int np = a.npos[b];
  if (r >= np) return;
  const float* roi = a.rois + ((long long)b * 10 + r) * 4;
  const float Sf = (float)a.S;
  float pb[4];
  for (int i = 0; i < 4; ++i) pb[i] = rintf(__fmul_rn(roi[i], Sf));
  int gx[8], gy[8];
  const float sub_w = __fdiv_rn(__fsub_rn(pb[3], pb[1]), (float)a.k), sub_h = __fdiv_rn(__fsub_rn(pb[2], pb[0]), (float)a.k);
  gx[0] = (int)pb[1]; gy[0] = (int)pb[0];
  for (int j = 1; j < a.k; ++j) {
    gx[j] = (int)rintf(__fadd_rn(pb[1], __fmul_rn((float)j, sub_w)));
    gy[j] = (int)rintf(__fadd_rn(pb[0], __fmul_rn((float)j, sub_h)));
  }
  gx[a.k] = (int)pb[3]; gy[a.k] = (int)pb[2];
  const int bw = gx[a.k] - gx[0], bh = gy[a.k] - gy[0];
  const long long area = (long long)bw * bh;
  double part = 0.0;
  if (area > 0) {
    const int kk = a.k * a.k, f = a.H / a.S;
    const unsigned char* gm = a.true_masks + ((long long)b * 20 + a.assign[b * 10 + r]) * a.H * a.H;
    const float wgt = a.mask_scale / ((float)area * (float)np * (float)a.B);
    for (long long i = threadIdx.x; i < area; i += blockDim.x) {
      const int y = gy[0] + (int)(i / bw), x = gx[0] + (int)(i % bw);
      if (y < 0 || y >= a.S || x < 0 || x >= a.S) continue;
      int by = 0, bx = 0;
      for (int j = 1; j < a.k; ++j) {
        if (y >= gy[j]) by = j;
        if (x >= gx[j]) bx = j;
      }
      const long long e = (((long long)b * a.S + y) * a.S + x) * kk + by * a.k + bx;
      const float l = a.mp_planar
                          ? a.mask_pos[(((long long)b * kk + by * a.k + bx) * a.S + y) * a.S + x]
                          : a.mask_pos[e];
      const float t = gm[(long long)(y * f) * a.H + x * f] ? 1.f : 0.f;
      part += (double)(fmaxf(l, 0.f) - l * t + log1pf(expf(-fabsf(l))));
      atomicAdd(a.dmask + e, wgt * (1.f / (1.f + expf(-l)) - t));
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double s = warp_sum(part);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0 && area > 0) {
    double tot = 0;
    for (int w = 0; w < kT / 32; ++w) tot += red[w];
    atomicAdd(a.loss, (double)a.mask_scale * tot / (double)area / (double)np / (double)a.B);
  }
}

__global__ void sumsq_kernel(const float* __restrict__ p, long long n, double scale, double* out) {
  __shared__ double red[kT / 32];
  double s = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += (double)p[i] * (double)p[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < kT / 32; ++w) t += red[w];
    atomicAdd(out, t * scale);
  }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps, float l2,
                            float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gg = g[i] * grad_scale + l2 * pv;
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm;
    v[i] = vv;
    p[i] = pv - lr_t * mm / (sqrtf(vv) + eps);
  }
}

__global__ void __launch_bounds__(kT)
adam_multi_kernel(const ParamSeg* __restrict__ segs, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, float lr_t, float b1, float b2, float eps, float grad_scale) {
  const ParamSeg sg = segs[blockIdx.y];
  float* __restrict__ p = sg.p;
  const float* __restrict__ gs = g + sg.off;
  float* __restrict__ ms = m + sg.off;
  float* __restrict__ vs = v + sg.off;
  for (long long i = blockIdx.x * (long long)kT + threadIdx.x; i < sg.n; i += (long long)gridDim.x * kT) {
    const float pv = p[i];
    const float gg = gs[i] * grad_scale + sg.l2 * pv;
    const float mm = b1 * ms[i] + (1.f - b1) * gg;
    const float vv = b2 * vs[i] + (1.f - b2) * gg * gg;
    ms[i] = mm;
    vs[i] = vv;
    p[i] = pv - lr_t * mm / (sqrtf(vv) + eps);
  }
}

__global__ void __launch_bounds__(kT) sumsq_multi_kernel(const ParamSeg* __restrict__ segs, double* out) {
  __shared__ double red[kT / 32];
  const ParamSeg sg = segs[blockIdx.y];
  if (!(sg.l2 > 0.f)) return;
  double s = 0;
  for (long long i = blockIdx.x * (long long)kT + threadIdx.x; i < sg.n; i += (long long)gridDim.x * kT)
    s += (double)sg.p[i] * (double)sg.p[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < kT / 32; ++w) t += red[w];
    if (t != 0.0) atomicAdd(out, t * 0.5 * (double)sg.l2);
  }
}

__global__ void bn_post_multi_kernel(const BnSeg* __restrict__ segs, float decay, float eps) {
  const BnSeg sg = segs[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= sg.C) return;
  const float mm = sg.mean[c] * decay + sg.bmean[c] * (1.f - decay);
  const float mv = sg.var[c] * decay + sg.bvar[c] * (1.f - decay);
  sg.mean[c] = mm;
  sg.var[c] = mv;
  const float inv = sg.gamma[c] / sqrtf(mv + eps);
  sg.scale[c] = inv;
  sg.shift[c] = sg.beta[c] - mm * inv;
}

__global__ void moving_update_kernel(float* mm, float* mv, const float* bm, const float* bv, int C, float decay) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mm[c] = mm[c] * decay + bm[c] * (1.f - decay);
  mv[c] = mv[c] * decay + bv[c] * (1.f - decay);
}

__global__ void refold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                              int C, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv = gamma[c] / sqrtf(var[c] + eps);
  scale[c] = inv;
  shift[c] = beta[c] - mean[c] * inv;
}

}  // namespace

#define DY_LAUNCH_OK()          \
  DY_CUDA(cudaGetLastError());  \
  return DY_OK

static int reduce_rows_per_block(long long M) {
  long long r = (M + 148 * 8 - 1) / (148 * 8);
  if (r < 32) r = 32;
  return (int)r;
}

int launch_bn_stats(const float* z, long long M, int C, double* sum, double* sumsq, cudaStream_t st) {
  const int rpb = reduce_rows_per_block(M);
  bn_reduce_kernel<false><<<(int)((M + rpb - 1) / rpb), kT, 0, st>>>(z, nullptr, nullptr, nullptr, nullptr, nullptr,
                                                                      0.f, 0, M, C, rpb, sum, sumsq);
  DY_LAUNCH_OK();
}
int launch_bn_finalize(const double* sum, const double* sumsq, long long M, int C, const float* gamma,
                       const float* beta, float eps, float* a, float* b, float* mean, float* var, float* invstd,
                       cudaStream_t st) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sum, sumsq, M, C, gamma, beta, eps, a, b, mean, var, invstd);
  DY_LAUNCH_OK();
}
int launch_bn_act(const float* z, const float* a, const float* b, int C, long long total, float alpha, int act,
                  const float* residual, float* y, cudaStream_t st) {
  bn_act_kernel<<<blocks_for(total), kT, 0, st>>>(z, a, b, C, total, alpha, act, residual, y);
  DY_LAUNCH_OK();
}
int launch_bn_bwd_reduce(const float* dy, const float* z, const float* a, const float* b, const float* mean,
                         const float* invstd, float alpha, int act, long long M, int C, double* s1, double* s2,
                         cudaStream_t st) {
  const int rpb = reduce_rows_per_block(M);
  bn_reduce_kernel<true><<<(int)((M + rpb - 1) / rpb), kT, 0, st>>>(dy, z, a, b, mean, invstd, alpha, act, M, C, rpb,
                                                                     s1, s2);
  DY_LAUNCH_OK();
}
int launch_bn_bwd_apply(const float* dy, const float* z, const float* a, const float* b, const float* mean,
                        const float* invstd, const float* gamma, const double* s1, const double* s2, float alpha,
                        int act, int mode, long long M, int C, float* dz, cudaStream_t st) {
  bn_bwd_apply_kernel<<<blocks_for(M * C), kT, 0, st>>>(dy, z, a, b, mean, invstd, gamma, s1, s2, alpha, act, mode, M,
                                                         C, dz);
  DY_LAUNCH_OK();
}
int launch_copy_stats_to_grads(const double* s1, const double* s2, int C, float* dgamma, float* dbeta,
                               cudaStream_t st) {
  copy_stats_kernel<<<(C + 127) / 128, 128, 0, st>>>(s1, s2, C, dgamma, dbeta);
  DY_LAUNCH_OK();
}
int launch_weight_transpose(const float* w, float* wt, int kk, int cin, int cout, cudaStream_t st) {
  weight_transpose_kernel<<<blocks_for((long long)kk * cin * cout), kT, 0, st>>>(w, wt, kk, cin, cout);
  DY_LAUNCH_OK();
}
int launch_conv_dgrad(const float* dz, const float* wt, float* dx, const ConvGeom& g, int accumulate,
                      cudaStream_t st) {
  const int bx = g.cin >= 64 ? 64 : 32;
  dim3 block(bx, 256 / bx);
  const int xgroups = (g.Wi + kDgPix - 1) / kDgPix;
  dim3 grid((g.cin + bx - 1) / bx, (xgroups + block.y - 1) / block.y, (unsigned)(g.B * g.Hi));
  conv_dgrad_kernel<<<grid, block, 0, st>>>(dz, wt, dx, g, accumulate);
  DY_LAUNCH_OK();
}
int launch_conv_wgrad(const float* x0, int c0, const float* x1, int c1, const float* dz, float* dw, const ConvGeom& g,
                      int num_sms, cudaStream_t st) {
  const int cin = c0 + c1;
  const int co_tiles = (g.cout + kWgT - 1) / kWgT, ci_tiles = (cin + kWgT - 1) / kWgT;
  const long long npix = (long long)g.B * g.Ho * g.Wo;
  const long long base = (long long)co_tiles * ci_tiles * g.k * g.k;
  long long splits = (8LL * num_sms + base - 1) / base;
  const long long max_splits = (npix + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long pps = (npix + splits - 1) / splits;
  pps = (pps + kWgP - 1) / kWgP * kWgP;
  splits = (npix + pps - 1) / pps;
  dim3 grid(co_tiles, ci_tiles * g.k * g.k, (unsigned)splits);
  conv_wgrad_kernel<<<grid, 256, 0, st>>>(x0, c0, x1, c1, dz, dw, g, pps);
  DY_LAUNCH_OK();
}
int launch_split_accumulate(const float* src, int B, int H, int W, int c0, int c1, float* dst0, float* dst1,
                            cudaStream_t st) {
  split_accumulate_kernel<<<blocks_for((long long)B * H * W * (c0 + c1)), kT, 0, st>>>(src, B, H, W, c0, c1, dst0,
                                                                                      dst1);
  DY_LAUNCH_OK();
}
int launch_add(float* dst, const float* src, long long n, cudaStream_t st) {
  add_kernel<<<blocks_for(n), kT, 0, st>>>(dst, src, n);
  DY_LAUNCH_OK();
}
int launch_yolo_loss(const YoloLossArgs& a, cudaStream_t st) {
  const long long n0 = 3LL * (a.g[0] * a.g[0] + a.g[1] * a.g[1] + a.g[2] * a.g[2]);
  yolo_loss_kernel<<<(int)((a.B * n0 + kT - 1) / kT), kT, 0, st>>>(a);
  DY_LAUNCH_OK();
}
int launch_mask_loss(const MaskLossArgs& a, cudaStream_t st) {
  DY_CHECK(a.k >= 1 && a.k <= 7, "k_map");
  mask_roi_kernel<<<a.B, 32, 0, st>>>(a);
  DY_CUDA(cudaGetLastError());
  mask_loss_kernel<<<dim3(10, a.B), kT, 0, st>>>(a);
  DY_LAUNCH_OK();
}
int launch_adam_multi(const ParamSeg* segs_dev, int nseg, const float* g, float* m, float* v, float lr_t, float b1,
                      float b2, float eps, float grad_scale, cudaStream_t st) {
  if (nseg <= 0) return DY_OK;
  adam_multi_kernel<<<dim3(148, nseg), kT, 0, st>>>(segs_dev, g, m, v, lr_t, b1, b2, eps, grad_scale);
  DY_LAUNCH_OK();
}
int launch_sumsq_multi(const ParamSeg* segs_dev, int nseg, double* out, cudaStream_t st) {
  if (nseg <= 0) return DY_OK;
  sumsq_multi_kernel<<<dim3(32, nseg), kT, 0, st>>>(segs_dev, out);
  DY_LAUNCH_OK();
}
int launch_bn_post_multi(const BnSeg* segs_dev, int nseg, float decay, float eps, cudaStream_t st) {
  if (nseg <= 0) return DY_OK;
  bn_post_multi_kernel<<<dim3(4, nseg), 256, 0, st>>>(segs_dev, decay, eps);
  DY_LAUNCH_OK();
}
int launch_sumsq(const float* p, long long n, double scale, double* out, cudaStream_t st) {
  sumsq_kernel<<<blocks_for(n, kT, 148), kT, 0, st>>>(p, n, scale, out);
  DY_LAUNCH_OK();
}
int launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float b1, float b2, float eps,
                float l2, float grad_scale, cudaStream_t st) {
  adam_kernel<<<blocks_for(n), kT, 0, st>>>(p, g, m, v, n, lr_t, b1, b2, eps, l2, grad_scale);
  DY_LAUNCH_OK();
}
int launch_moving_update(float* mov_mean, float* mov_var, const float* bmean, const float* bvar, int C, float decay,
                         cudaStream_t st) {
  moving_update_kernel<<<(C + 127) / 128, 128, 0, st>>>(mov_mean, mov_var, bmean, bvar, C, decay);
  DY_LAUNCH_OK();
}
int launch_refold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int C,
                  float* scale, float* shift, cudaStream_t st) {
  refold_kernel<<<(C + 127) / 128, 128, 0, st>>>(gamma, beta, mean, var, eps, C, scale, shift);
  DY_LAUNCH_OK();
}

}  // namespace dy
