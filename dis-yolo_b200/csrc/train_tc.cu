// Training step on the tensor-core engine: wgrad tcgen05 kernel, bf16 P1 BatchNorm / activation /
// pooling kernels, weight repacking (see train_tc.cuh).
#include "train_tc.cuh"

#include "conv_tc.cuh"

namespace dy {

namespace {

constexpr int kT = 256;
int g_train_pdl = 1;      // 1: the BN / elementwise / wgrad kernels are launched with programmatic stream serialization

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  __nv_bfloat162 h;
  h = __floats2bfloat162_rn(f[0], f[1]); q.x = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[2], f[3]); q.y = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[4], f[5]); q.z = *reinterpret_cast<uint32_t*>(&h);
  h = __floats2bfloat162_rn(f[6], f[7]); q.w = *reinterpret_cast<uint32_t*>(&h);
  return q;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
  const float4 u = __ldg(reinterpret_cast<const float4*>(p));
  const float4 v = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w; f[4] = v.x; f[5] = v.y; f[6] = v.z; f[7] = v.w;
}

static int grid_for(long long total, int cap = 148 * 8) {
  long long g = (total + kT - 1) / kT;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// -------------------------------------------------------------------------------------------
// Elementwise / reduction kernels over a P1 matrix [rows, C] of bf16.  Thread t of a block owns the
// 16-byte vector column v = t % (C/8) and the rows (t / (C/8)) + k*R of the block's row range: no
// per-element index division, 32-bit row arithmetic (rows < 2^31), and kUnroll independent 16-byte
// loads per operand in flight per thread -- these kernels are HBM-bound.
// -------------------------------------------------------------------------------------------
constexpr int kUnroll = 4;

__device__ __forceinline__ bool p1_pixel(uint32_t r, uint32_t Hp, uint32_t Wp, int H, int W, int* n, int* y, int* x) {
  const uint32_t t = r / Wp;
  *x = (int)(r - t * Wp);
  const uint32_t nn = t / Hp;
  *n = (int)nn;
  *y = (int)(t - nn * Hp);
  return *x < W && *y < H;
}

// per-channel reductions: fp32 partial sums per thread are flushed to double every 8 row groups.
// BWD: g = dy * leaky'(z*a+b);  s1 = sum g;  s2 = sum g*xhat = invstd * (sum g*z - mean * sum g) -- the loop only
// accumulates sum g and sum g*z (two per-channel constants fewer in registers), the affine part is applied once
// per block in double.  Pad pixels carry dy = 0 and z = 0: the flat row space needs no pixel arithmetic.
template <bool BWD>
__global__ void __launch_bounds__(kT, 2)
p1_reduce_kernel(const __nv_bfloat16* __restrict__ u, const __nv_bfloat16* __restrict__ z,
                 const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mean,
                 const float* __restrict__ invstd, float alpha, int act, uint32_t rows, int C,
                 double* __restrict__ o1, double* __restrict__ o2) {
  pdl_launch_dependents();      // the next kernel of the step may be scheduled while this one drains ...
  pdl_wait();                   // ... and this one starts its work only when its predecessor has completed
  __shared__ double sm[16 * kT];
  const int cv = C >> 3;
  const int v = threadIdx.x % cv, rl = threadIdx.x / cv, R = kT / cv;
  const uint32_t G = gridDim.x * (uint32_t)R;
  double d1[8], d2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) d1[j] = d2[j] = 0.0;
  float fa[8], fb[8];
  if (BWD) {
    load8f(a + v * 8, fa);
    load8f(b + v * 8, fb);
  }
  uint32_t r = blockIdx.x * (uint32_t)R + rl;
  while (r < rows) {
    float p1[8], p2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p1[j] = p2[j] = 0.f;
    for (int it = 0; it < 8 && r < rows; ++it, r += kUnroll * G) {
      uint4 qu[kUnroll], qz[kUnroll];
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        const uint32_t rr = r + k * G;
        qu[k] = make_uint4(0, 0, 0, 0);
        qz[k] = make_uint4(0, 0, 0, 0);
        if (rr < rows) {
          qu[k] = __ldg(reinterpret_cast<const uint4*>(u + (size_t)rr * C) + v);
          if (BWD) qz[k] = __ldg(reinterpret_cast<const uint4*>(z + (size_t)rr * C) + v);
        }
      }
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        float fu[8];
        unpack8(qu[k], fu);
        if (BWD) {
          float fz[8];
          unpack8(qz[k], fz);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g = fu[j];               // rows beyond the end and pad pixels carry dy = 0
            if (act && !(fmaf(fz[j], fa[j], fb[j]) > 0.f)) g *= alpha;
            p1[j] += g;
            p2[j] = fmaf(g, fz[j], p2[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            p1[j] += fu[j];
            p2[j] = fmaf(fu[j], fu[j], p2[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d1[j] += (double)p1[j];
      d2[j] += (double)p2[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sm[j * kT + threadIdx.x] = d1[j];
    sm[(8 + j) * kT + threadIdx.x] = d2[j];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 8 * cv; o += kT) {
    const int j = o / cv, x = o - j * cv;
    double t1 = 0.0, t2 = 0.0;
    for (int q = 0; q < R; ++q) {
      t1 += sm[j * kT + q * cv + x];
      t2 += sm[(8 + j) * kT + q * cv + x];
    }
    const int ch = x * 8 + j;
    if (BWD) t2 = (double)__ldg(invstd + ch) * (t2 - (double)__ldg(mean + ch) * t1);
    atomicAdd(o1 + ch, t1);
    if (o2) atomicAdd(o2 + ch, t2);
  }
}

// fin.sum != nullptr: the batch-moment finalize (bn_finalize_kernel: a = gamma*invstd, b = beta - mean*a, in
// double from the per-channel sums) is done here by every thread for its own 8 channels, and block 0 also
// publishes a / b / mean / var / invstd for the backward pass -- one launch less per trained layer
struct BnFinalize {
  const double* sum;
  const double* sumsq;
  const float* gamma;
  const float* beta;
  float* a_out;
  float* b_out;
  float* mean_out;
  float* var_out;
  float* invstd_out;
  long long M;
  float eps;
};

// Both elementwise kernels walk the tensor by IMAGE ROW: a row (n, y) of the P1 layout is W*C contiguous valid
// elements, so a block streams it with thread t owning the 16-byte channel vector t % (C/8) of the pixels
// t / (C/8) + k*(256 / (C/8)) -- no per-vector pixel arithmetic (one division per image row), no validity tests
// beyond x < W.  (The first version walked flat pixel rows and divided twice per 16-byte vector: 2 TB/s.)
__global__ void __launch_bounds__(kT, 3)
bn_act_p1_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ a, const float* __restrict__ b,
                 BnFinalize fin, const __nv_bfloat16* __restrict__ residual, int B, int H, int W, int C,
                 float alpha, int act, __nv_bfloat16* __restrict__ out_same, __nv_bfloat16* __restrict__ out_up,
                 __nv_bfloat16* __restrict__ out_s2d) {
  pdl_launch_dependents();      // the next kernel of the step may be scheduled while this one drains ...
  pdl_wait();                   // ... and this one starts its work only when its predecessor has completed
  const int cv = C >> 3;
  const int Hp = H + 1, Wp = W + 1;
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv, P = kT / cv;
  float fa[8], fb[8];
  if (fin.sum != nullptr) {
    // batch-moment finalize in double, ONE channel per thread (FP64 division / sqrt are long dependent chains:
    // eight channels per thread on eight threads cost every block ~10 us), shared through smem
    __shared__ float s_a[8 * kT], s_b[8 * kT];
    const double invM = 1.0 / (double)fin.M;
    for (int c = threadIdx.x; c < C; c += kT) {
      const double m = fin.sum[c] * invM;
      double var = fin.sumsq[c] * invM - m * m;
      if (var < 0.0) var = 0.0;
      const float is = (float)(1.0 / sqrt(var + (double)fin.eps));
      const float aa = fin.gamma[c] * is;
      const float bb = fin.beta[c] - (float)m * aa;
      s_a[c] = aa;
      s_b[c] = bb;
      if (blockIdx.x == 0) {
        fin.a_out[c] = aa;
        fin.b_out[c] = bb;
        fin.mean_out[c] = (float)m;
        fin.var_out[c] = (float)var;
        fin.invstd_out[c] = is;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      fa[j] = s_a[v * 8 + j];
      fb[j] = s_b[v * 8 + j];
    }
  } else {
    load8f(a + v * 8, fa);
    load8f(b + v * 8, fb);
  }
  const int NH = B * H;
  for (int ry = blockIdx.x; ry < NH; ry += gridDim.x) {
    const int n = ry / H, y = ry - n * H;
    const size_t rowpix = ((size_t)n * Hp + y) * Wp;
    for (int x0 = pl; x0 < W; x0 += kUnroll * P) {
      uint4 qz[kUnroll], qr[kUnroll];
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        const int x = x0 + k * P;
        qr[k] = make_uint4(0, 0, 0, 0);
        if (x < W) {
          qz[k] = __ldg(reinterpret_cast<const uint4*>(z + (rowpix + x) * C) + v);
          if (residual) qr[k] = __ldg(reinterpret_cast<const uint4*>(residual + (rowpix + x) * C) + v);
        }
      }
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        const int x = x0 + k * P;
        if (x >= W) continue;
        float fz[8], fr[8];
        unpack8(qz[k], fz);
        unpack8(qr[k], fr);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = fmaf(fz[j], fa[j], fb[j]);
          if (act) t = fmaxf(alpha * t, t);
          fz[j] = t + fr[j];
        }
        const uint4 q = pack8(fz);
        if (out_same) reinterpret_cast<uint4*>(out_same + (rowpix + x) * C)[v] = q;
        if (out_up) {
          const size_t Wu = 2 * W + 1, Hu = 2 * H + 1;
          const size_t ru = ((size_t)n * Hu + 2 * y) * Wu + 2 * x;
          reinterpret_cast<uint4*>(out_up + ru * C)[v] = q;
          reinterpret_cast<uint4*>(out_up + (ru + 1) * C)[v] = q;
          reinterpret_cast<uint4*>(out_up + (ru + Wu) * C)[v] = q;
          reinterpret_cast<uint4*>(out_up + (ru + Wu + 1) * C)[v] = q;
        }
        if (out_s2d) {      // space-to-depth copy for a stride-2 consumer: [N, H/2+1, W/2+1, 4C], block (y&1)*2+(x&1)
          const size_t Hq = H / 2 + 1, Wq = W / 2 + 1;
          const size_t rs = ((size_t)n * Hq + (y >> 1)) * Wq + (x >> 1);
          reinterpret_cast<uint4*>(out_s2d + rs * 4 * C + (((y & 1) << 1) | (x & 1)) * C)[v] = q;
        }
      }
    }
  }
}

// dz = g*k0 - k1 - xhat*k2 (mode 0) with xhat = (z - mean)*invstd, folded to  dz = g*k0 + c1 + z*c2
// (c2 = -invstd*k2, c1 = mean*invstd*k2 - k1: three per-channel constants instead of five), or g*a (mode 1).
// Writes EVERY pixel of the flat row space [0, rows_out): zeros at pad pixels and in the tail (dz is a scratch
// shared by layers of different geometry).
__global__ void __launch_bounds__(kT, 3)
bn_bwd_apply_p1_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ z,
                       const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mean,
                       const float* __restrict__ invstd, const float* __restrict__ gamma,
                       const double* __restrict__ s1, const double* __restrict__ s2, float alpha, int act, int mode,
                       uint32_t rows, uint32_t rows_out, long long mvalid, int B, int H, int W, int C,
                       __nv_bfloat16* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_launch_dependents();      // the next kernel of the step may be scheduled while this one drains ...
  pdl_wait();                   // ... and this one starts its work only when its predecessor has completed
  const int cv = C >> 3;
  const int Hp = H + 1, Wp = W + 1;
  const int v = threadIdx.x % cv, pl = threadIdx.x / cv, P = kT / cv;
  const float invM = 1.f / (float)mvalid;
  if (dgamma != nullptr && blockIdx.x == 0 && pl == 0) {      // d gamma = sum g*xhat, d beta = sum g (copy_stats)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dgamma[v * 8 + j] = (float)s2[v * 8 + j];
      dbeta[v * 8 + j] = (float)s1[v * 8 + j];
    }
  }
  float fa[8], fb[8], k0[8], c1[8], c2[8];
  load8f(a + v * 8, fa);
  load8f(b + v * 8, fb);
  if (mode == 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { k0[j] = fa[j]; c1[j] = c2[j] = 0.f; }
  } else {
    float fg[8], fm[8], fi[8];
    load8f(mean + v * 8, fm);
    load8f(invstd + v * 8, fi);
    load8f(gamma + v * 8, fg);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k0[j] = fg[j] * fi[j];
      const float k1 = k0[j] * ((float)s1[v * 8 + j] * invM);
      const float k2 = k0[j] * ((float)s2[v * 8 + j] * invM);
      c2[j] = -fi[j] * k2;
      c1[j] = fm[j] * fi[j] * k2 - k1;
    }
  }
  const int NHp = B * Hp;
  for (int fr = blockIdx.x; fr < NHp; fr += gridDim.x) {
    const int n = fr / Hp, y = fr - n * Hp;
    const size_t rowpix = (size_t)fr * Wp;
    const bool vrow = y < H;
    for (int x0 = pl; x0 < Wp; x0 += kUnroll * P) {
      uint4 qg[kUnroll], qz[kUnroll];
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        const int x = x0 + k * P;
        if (vrow && x < W) {
          qg[k] = __ldg(reinterpret_cast<const uint4*>(dy + (rowpix + x) * C) + v);
          qz[k] = __ldg(reinterpret_cast<const uint4*>(z + (rowpix + x) * C) + v);
        }
      }
#pragma unroll
      for (int k = 0; k < kUnroll; ++k) {
        const int x = x0 + k * P;
        if (x >= Wp) continue;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (vrow && x < W) {
          float fg[8], fz[8];
          unpack8(qg[k], fg);
          unpack8(qz[k], fz);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g = fg[j];
            if (act && !(fmaf(fz[j], fa[j], fb[j]) > 0.f)) g *= alpha;
            fg[j] = fmaf(fz[j], c2[j], fmaf(g, k0[j], c1[j]));
          }
          q = pack8(fg);
        }
        reinterpret_cast<uint4*>(dz + (rowpix + x) * C)[v] = q;
      }
    }
  }
  if (blockIdx.x == 0) {      // tail rows [rows, rows_out)
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (uint32_t i = threadIdx.x; i < (rows_out - rows) * (uint32_t)cv; i += kT)
      reinterpret_cast<uint4*>(dz + (size_t)rows * C)[i] = zero;
  }
}

// colsum != nullptr: per-channel sums of src (= the bias gradient of a biased linear conv, dz = dy) are
// accumulated on the way: shared-memory float partials per block, one float atomic per channel per block
__global__ void __launch_bounds__(kT)
f32_to_p1_kernel(const float* __restrict__ src, long long rows, long long rows_out, int H, int W, int C, int Cg,
                 __nv_bfloat16* __restrict__ dst, float* __restrict__ colsum) {
  pdl_launch_dependents();      // the next kernel of the step may be scheduled while this one drains ...
  pdl_wait();                   // ... and this one starts its work only when its predecessor has completed
  __shared__ float part[64];
  const int cv = Cg >> 3, Hp = H + 1, Wp = W + 1;
  const long long total = rows_out * cv;
  if (colsum) {
    if (threadIdx.x < 64) part[threadIdx.x] = 0.f;
    __syncthreads();
  }
  for (long long i = blockIdx.x * (long long)kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const long long r = i / cv;
    const int v = (int)(i - r * cv);
    int n, y, x;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (r < rows && p1_pixel((uint32_t)r, (uint32_t)Hp, (uint32_t)Wp, H, W, &n, &y, &x)) {
      const float* s = src + (((long long)n * H + y) * W + x) * C;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (v * 8 + j < C) ? __ldg(s + v * 8 + j) : 0.f;
      q = pack8(f);
      if (colsum) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v * 8 + j < C && f[j] != 0.f) atomicAdd(&part[v * 8 + j], f[j]);
      }
    }
    reinterpret_cast<uint4*>(dst + r * Cg)[v] = q;
  }
  if (colsum) {
    __syncthreads();
    if (threadIdx.x < C && part[threadIdx.x] != 0.f) atomicAdd(colsum + threadIdx.x, part[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(kT)
pool2x2_p1_kernel(const __nv_bfloat16* __restrict__ src, long long rows, long long rows_out, int h, int w, int C,
                  __nv_bfloat16* __restrict__ dst) {
  pdl_launch_dependents();      // the next kernel of the step may be scheduled while this one drains ...
  pdl_wait();                   // ... and this one starts its work only when its predecessor has completed
  const int cv = C >> 3, Hp = h + 1, Wp = w + 1;
  const long long total = rows_out * cv;
  const long long Wu = 2 * w + 1, Hu = 2 * h + 1;
  for (long long i = blockIdx.x * (long long)kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    const long long r = i / cv;
    const int v = (int)(i - r * cv);
    int n, y, x;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (r < rows && p1_pixel((uint32_t)r, (uint32_t)Hp, (uint32_t)Wp, h, w, &n, &y, &x)) {
      const long long ru = ((long long)n * Hu + 2 * y) * Wu + 2 * x;
      float f0[8], f1[8], f2[8], f3[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + ru * C) + v), f0);
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (ru + 1) * C) + v), f1);
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (ru + Wu) * C) + v), f2);
      unpack8(__ldg(reinterpret_cast<const uint4*>(src + (ru + Wu + 1) * C) + v), f3);
#pragma unroll
      for (int j = 0; j < 8; ++j) f0[j] = (f0[j] + f1[j]) + (f2[j] + f3[j]);
      q = pack8(f0);
    }
    reinterpret_cast<uint4*>(dst + r * C)[v] = q;
  }
}

__global__ void add_p1_kernel(__nv_bfloat16* __restrict__ dst, const __nv_bfloat16* __restrict__ src, long long nvec,
                              int copy) {
  pdl_launch_dependents();      // the next kernel of the step may be scheduled while this one drains ...
  pdl_wait();                   // ... and this one starts its work only when its predecessor has completed
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 s = __ldg(reinterpret_cast<const uint4*>(src) + i);
    if (copy) {
      reinterpret_cast<uint4*>(dst)[i] = s;
    } else {
      float fd[8], fs[8];
      unpack8(reinterpret_cast<const uint4*>(dst)[i], fd);
      unpack8(s, fs);
#pragma unroll
      for (int j = 0; j < 8; ++j) fd[j] += fs[j];
      reinterpret_cast<uint4*>(dst)[i] = pack8(fd);
    }
  }
}

// [K][cout] fp32 -> [cout_pad][K] bf16 through a 32x32 smem tile (both sides coalesced)
__global__ void pack_fwd_kernel(const float* __restrict__ w, int K, int cout, int cout_pad,
                                __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int kk = k0 + i, co = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (kk < K && co < cout) ? w[(size_t)kk * cout + co] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int co = c0 + i, kk = k0 + threadIdx.x;
    if (co < cout_pad && kk < K) out[(size_t)co * K + kk] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

__global__ void pack_dgrad_kernel(const float* __restrict__ w, int k, int cin_total, int ci0, int cin_sel, int cout,
                                  int Cg, __nv_bfloat16* __restrict__ out) {
  const int kk = k * k;
  const long long total = (long long)cin_sel * kk * Cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cg);
    const long long t = i / Cg;
    const int tp = (int)(t % kk);
    const int ci = (int)(t / kk);
    float v = 0.f;
    if (co < cout) v = w[((size_t)(kk - 1 - tp) * cin_total + ci0 + ci) * cout + co];
    out[i] = __float2bfloat16(v);
  }
}

// dgrad operands of a 3x3 stride-2 conv (TF 'SAME': input pixel (2oy+kh, 2ox+kw)).  In the space-to-depth view of
// the input, parity block (by,bx) receives the taps with (kh&1, kw&1) == (by,bx): 4 / 2 / 2 / 1 taps for blocks
// 0..3, listed kh-major.  Region of block b: [cin][ntap_b * Cg], element (ci, t*Cg + co) = w[kh_t][kw_t][ci][co]
// (no rotation: dX[q, b] = sum_t dz[q - (kh_t>>1, kw_t>>1)] . W[tap_t]^T).  The four regions are concatenated.
__global__ void pack_dgrad_s2_kernel(const float* __restrict__ w, int cin, int cout, int Cg,
                                     __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)cin * 9 * Cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // region starts (in units of cin*Cg): block 0 -> 0 (4 taps), 1 -> 4 (2), 2 -> 6 (2), 3 -> 8 (1)
    const long long unit = (long long)cin * Cg;
    const int u = (int)(i / unit);
    const int b = u < 4 ? 0 : (u < 6 ? 1 : (u < 8 ? 2 : 3));
    const int ubase = b == 0 ? 0 : (b == 1 ? 4 : (b == 2 ? 6 : 8));
    const int nt = b == 0 ? 4 : (b == 3 ? 1 : 2);
    const long long j = i - (long long)ubase * unit;           // offset inside the region [cin][nt*Cg]
    const int co = (int)(j % Cg);
    const int t = (int)((j / Cg) % nt);
    const int ci = (int)(j / ((long long)nt * Cg));
    const int by = b >> 1, bx = b & 1;
    const int nkw = bx == 0 ? 2 : 1;
    const int kh = by == 0 ? 2 * (t / nkw) : 1;
    const int kw = bx == 0 ? 2 * (t % nkw) : 1;
    float v = 0.f;
    if (co < cout) v = w[((size_t)(kh * 3 + kw) * cin + ci) * cout + co];
    out[i] = __float2bfloat16(v);
  }
}

// convolutional1 weight gradient (3 -> 32, 3x3 stride 1, TF 'SAME' pad 1): dW[k][c] = sum_p patch_k(p) * dz[p][c],
// k = (kh*3+kw)*3 + ci.  K = 27 does not make a tensor-core operand; 9 GFLOP per 16 images on CUDA cores.
// The image taps are rounded to bf16 first: that is the operand the forward's tcgen05 kernel multiplied with.
// A block works on segments of kC1wPix pixels of one image row: the three image rows it needs (with a one-pixel
// halo, zero outside the image) and the dz segment are staged once, coalesced, and every patch element is then a
// direct shared-memory read -- no per-element index arithmetic.  216 threads = 27 taps x 8 groups of 4 channels,
// fp32 partial sums in registers across all of the block's segments, one atomic per (tap, channel) per block.
constexpr int kC1wPix = 64;
__global__ void __launch_bounds__(216)
conv1_wgrad_kernel(const float* __restrict__ img, const __nv_bfloat16* __restrict__ dz, int B, int H, int W,
                   float* __restrict__ dw) {
  __shared__ float s_img[3][(kC1wPix + 2) * 3];
  __shared__ __align__(16) float s_dz[kC1wPix][32];
  const int k = threadIdx.x / 8, cg = threadIdx.x % 8;
  const int kh = k / 9, koff = k % 9;                    // patch_k(x) = s_img[kh][x*3 + (kw*3 + ci)], kw*3+ci = k % 9
  const int segs_per_row = (W + kC1wPix - 1) / kC1wPix;
  const long long nseg = (long long)B * H * segs_per_row;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long sg = blockIdx.x; sg < nseg; sg += gridDim.x) {
    const int x0 = (int)(sg % segs_per_row) * kC1wPix;
    const int y = (int)((sg / segs_per_row) % H), n = (int)(sg / ((long long)segs_per_row * H));
    const int npx = min(kC1wPix, W - x0);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * (kC1wPix + 2) * 3; i += 216) {
      const int r = i / ((kC1wPix + 2) * 3), e = i % ((kC1wPix + 2) * 3);
      const int yy = y + r - 1, xx = x0 - 1 + e / 3;
      float v = 0.f;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = __bfloat162float(__float2bfloat16(__ldg(img + (((long long)n * H + yy) * W + xx) * 3 + e % 3)));
      s_img[r][e] = v;
    }
    const __nv_bfloat16* zrow = dz + (((long long)n * (H + 1) + y) * (W + 1) + x0) * 32;
    for (int i = threadIdx.x; i < kC1wPix * 4; i += 216) {      // 16-byte vectors: 8 channels each
      const int pp = i >> 2, v8 = i & 3;
      float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (pp < npx) unpack8(__ldg(reinterpret_cast<const uint4*>(zrow + (size_t)pp * 32) + v8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s_dz[pp][v8 * 8 + j] = f[j];
    }
    __syncthreads();
    const float* prow = &s_img[kh][koff];
#pragma unroll 16
    for (int pp = 0; pp < kC1wPix; ++pp) {
      const float a = prow[pp * 3];
      const float4 g = *reinterpret_cast<const float4*>(&s_dz[pp][cg * 4]);
      acc[0] = fmaf(a, g.x, acc[0]); acc[1] = fmaf(a, g.y, acc[1]);
      acc[2] = fmaf(a, g.z, acc[2]); acc[3] = fmaf(a, g.w, acc[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) atomicAdd(dw + k * 32 + cg * 4 + j, acc[j]);
}

__global__ void pack_fwd_multi_kernel(const PackSeg* __restrict__ segs) {
  __shared__ float tile[32][33];
  const PackSeg sg = segs[blockIdx.y];
  const int tiles_k = (sg.K + 31) / 32, tiles_c = (sg.cout_pad + 31) / 32;
  if ((int)blockIdx.x >= tiles_k * tiles_c) return;
  const int k0 = ((int)blockIdx.x % tiles_k) * 32, c0 = ((int)blockIdx.x / tiles_k) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int kk = k0 + i, co = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (kk < sg.K && co < sg.cout) ? sg.w[(size_t)kk * sg.cout + co] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int co = c0 + i, kk = k0 + threadIdx.x;
    if (co < sg.cout_pad && kk < sg.K) sg.wpk[(size_t)co * sg.K + kk] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}

__global__ void pack_dgrad_multi_kernel(const PackSeg* __restrict__ segs) {
  const PackSeg sg = segs[blockIdx.y >> 1];
  const int which = blockIdx.y & 1;
  __nv_bfloat16* out = which ? sg.wdg1 : sg.wdg0;
  if (out == nullptr) return;
  const int kk = sg.k * sg.k, cin_total = sg.cin0 + sg.cin1;
  const int ci0 = which ? sg.cin0 : 0, cin_sel = which ? sg.cin1 : sg.cin0;
  const long long total = (long long)cin_sel * kk * sg.Cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % sg.Cg);
    const long long t = i / sg.Cg;
    const int tp = (int)(t % kk);
    const int ci = (int)(t / kk);
    float v = 0.f;
    if (co < sg.cout) v = sg.w[((size_t)(kk - 1 - tp) * cin_total + ci0 + ci) * sg.cout + co];
    out[i] = __float2bfloat16(v);
  }
}

// -------------------------------------------------------------------------------------------
// wgrad: one CTA per (tap, 128-input-channel block, N tile, K split).  256 threads:
//   warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..7 epilogue.
// Stage = A [128 channels x 64 pixel rows] + B [block_n channels x 64 pixel rows], both exactly as
// TMA delivers the forward's activation boxes: channel-contiguous rows, 128B/64B swizzle.  For the
// MMA these are MN-major operands: "leading" offset = distance between two 64-(32-)channel boxes,
// "stride" offset = 8 pixel rows; one K=16 step advances the start address by 16 rows.
// -------------------------------------------------------------------------------------------
constexpr int kWgradRows = 64;     // pixel rows (K) per pipeline stage
constexpr int kWgradHaloRows = 72; // fused 3x3 mode: 64 + 2 halo rows, rounded up to the 8-row swizzle group
constexpr int kWgradMaxStages = 6;

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(kT, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX0, const __grid_constant__ CUtensorMap mapX1,
                const __grid_constant__ CUtensorMap mapZ, const __grid_constant__ WgradParams p, int dbg_lbo_a,
                int dbg_sbo_a, int dbg_lbo_b, int dbg_sbo_b) {
  pdl_launch_dependents();      // the next kernel of the step may be scheduled while this one drains ...
  pdl_wait();                   // ... and this one starts its work only when its predecessor has completed
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kWgradMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kWgradMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- which unit of work ----
  int u = blockIdx.x;
  const int ks = u % p.ksplit; u /= p.ksplit;
  const int nt = u % p.n_tiles_n; u /= p.n_tiles_n;
  const int blk_total = p.src_blk[0] + (p.nsrc > 1 ? p.src_blk[1] : 0);
  const int cb = u % blk_total;
  const int tap = u / blk_total;                 // fused mode: the kernel ROW kh (taps kh*3 .. kh*3+2)
  const int nacc = p.fuse_kw ? 3 : 1;            // accumulators = taps sharing this CTA's dz boxes
  const int a_rows = p.fuse_kw ? kWgradHaloRows : kWgradRows;
  const int src = cb >= p.src_blk[0] ? 1 : 0;
  const int ci0 = (src ? cb - p.src_blk[0] : cb) * 128;
  const int c_src = p.src_c[src], aw = p.src_aw[src];
  const int a_valid = ((c_src - ci0 < 128) ? (c_src - ci0) : 128) / aw;     // boxes holding real channels
  const int b_boxes = p.block_n / p.z_aw;
  const int n0 = nt * p.block_n;
  const int c_begin = ks * p.chunks_per_split;
  int c_end = c_begin + p.chunks_per_split;
  if (c_end > p.total_chunks) c_end = p.total_chunks;
  const int nchunks = c_end - c_begin;

  const uint32_t a_box = (uint32_t)a_rows * aw * 2;              // bytes of one A box
  const uint32_t b_box = (uint32_t)kWgradRows * p.z_aw * 2;
  const uint32_t a_bytes = 128u * (uint32_t)a_rows * 2;          // 16 KB (18 KB with the halo): all 128 channels
  const uint32_t stage_bytes = a_bytes + (uint32_t)p.block_n * kWgradRows * 2;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));

  // channels beyond the source's extent are never loaded: their A rows must read as zero
  if (a_valid * aw < 128) {
    uint4* q = reinterpret_cast<uint4*>(smem_gen);
    const int nvec = (int)((uint32_t)p.num_stages * stage_bytes / 16);
    for (int i = threadIdx.x; i < nvec; i += kT) q[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(src ? &mapX1 : &mapX0);
    tma_prefetch_desc(&mapZ);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, p.fuse_kw ? 512u : 256u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (elect_one() && nchunks > 0) {
      const CUtensorMap* xm = src ? &mapX1 : &mapX0;
      const uint32_t tx = (uint32_t)a_valid * a_box + (uint32_t)b_boxes * b_box;
      // fused: one (64+8)-row X box starting one pixel left of the kernel row's centre tap feeds kw = 0, 1, 2
      const int shift = p.fuse_kw ? p.tap_shift[tap * 3] : p.tap_shift[tap];
      uint32_t stage = 0, phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_expect_tx(&full_bar[stage], tx);
        uint8_t* sa = smem_gen + (size_t)stage * stage_bytes;
        const int row = c * kWgradRows;
        const int col = p.tap_col[p.fuse_kw ? tap * 3 : tap] + ci0;
        for (int j = 0; j < a_valid; ++j) tma_load_2d(sa + (size_t)j * a_box, xm, &full_bar[stage], col + j * aw, row + shift);
        for (int j = 0; j < b_boxes; ++j)
          tma_load_2d(sa + a_bytes + (size_t)j * b_box, &mapZ, &full_bar[stage], n0 + j * p.z_aw, row);
        if (++stage == (uint32_t)p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one() && nchunks > 0) {
      // instruction descriptor: bf16 x bf16 -> fp32, M=128, N=block_n, A and B MN-major (bits 15, 16)
      const uint32_t idesc = umma_idesc_bf16(p.block_n) | (1u << 15) | (1u << 16);
      const uint32_t lay_a = aw == 64 ? 2u : 4u, lay_b = p.z_aw == 64 ? 2u : 4u;
      const uint32_t lbo_a = dbg_lbo_a ? (uint32_t)dbg_lbo_a : a_box, sbo_a = dbg_sbo_a ? (uint32_t)dbg_sbo_a : 8u * aw * 2;
      const uint32_t lbo_b = dbg_lbo_b ? (uint32_t)dbg_lbo_b : b_box, sbo_b = dbg_sbo_b ? (uint32_t)dbg_sbo_b : 8u * p.z_aw * 2;
      const uint32_t kstep_a = 16u * aw * 2, kstep_b = 16u * p.z_aw * 2;   // 16 pixel rows
      uint32_t stage = 0, phase = 0, acc = 0u;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
#pragma unroll
        for (int k = 0; k < kWgradRows / 16; ++k) {
          const uint64_t bdesc = mn_desc(sa + a_bytes + k * kstep_b, lbo_b, sbo_b, lay_b);
          // tap kw = the same box read kw pixel rows further down: the K (row) offset of an MN-major operand is
          // linear in the start address (SBO = 8 rows), and the swizzle is a function of the smem address
          for (int t = 0; t < nacc; ++t) {
            const uint64_t adesc = mn_desc(sa + k * kstep_a + (uint32_t)t * aw * 2, lbo_a, sbo_a, lay_a);
            umma_bf16(tmem_base + (uint32_t)(t * p.block_n), adesc, bdesc, idesc, acc);
          }
          acc = 1u;
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == (uint32_t)p.num_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(&tfull_bar);
    }
  } else if (warp >= 4 && nchunks > 0) {
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // input channel inside the block == TMEM lane
    const bool valid = ci0 + row < c_src;
    mbar_wait(&tfull_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool vec4 = (p.cout & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dw) & 15) == 0;
    uint32_t r[16];
    for (int t = 0; t < nacc; ++t) {
      const int tp = p.fuse_kw ? tap * 3 + t : tap;
      float* drow = p.dw + ((size_t)tp * p.cin_total + p.src_koff[src] + ci0 + row) * p.cout;
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        tmem_ld16(taddr + (uint32_t)(t * p.block_n + c0), r);
        tmem_ld_wait();
        if (valid) {
          const int co0 = n0 + c0;
          if (vec4 && co0 + 16 <= p.cout) {
            // 16 consecutive fp32 of one dW row: four 16-byte reductions / stores instead of sixteen scalar ones
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float* d = drow + co0 + j;
              if (p.ksplit > 1) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(__uint_as_float(r[j])),
                             "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])),
                             "f"(__uint_as_float(r[j + 3]))
                             : "memory");
              } else {
                *reinterpret_cast<float4*>(d) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                            __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = co0 + j;
              if (co < p.cout) {
                if (p.ksplit > 1) atomicAdd(drow + co, __uint_as_float(r[j]));
                else drow[co] = __uint_as_float(r[j]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.fuse_kw ? 512u : 256u);
  }
}

int g_wdbg[4] = {0, 0, 0, 0};
int g_wgrad_fuse = 1;

}  // namespace

void wgrad_set_fuse(int on) { g_wgrad_fuse = on; }
void train_set_pdl(int on) { g_train_pdl = on; }

void wgrad_set_debug(int lbo_a, int sbo_a, int lbo_b, int sbo_b) {
  g_wdbg[0] = lbo_a; g_wdbg[1] = sbo_a; g_wdbg[2] = lbo_b; g_wdbg[3] = sbo_b;
}

// grid for the row-strided kernels: enough 256-thread blocks to keep ~8 independent 16-byte loads per
// thread in flight on every SM, never more blocks than row groups
static int row_grid(long long rows, int C) {
  const int R = kT / (C / 8);
  long long g = (rows + (long long)R * kUnroll - 1) / ((long long)R * kUnroll);
  if (g > 148 * 6) g = 148 * 6;
  if (g < 1) g = 1;
  return (int)g;
}

// image-row kernels: one block per image row, at most 8 resident waves
static int image_row_grid(long long image_rows) {
  long long g = image_rows;
  if (g > 148 * 8) g = 148 * 8;
  return g < 1 ? 1 : (int)g;
}

// reductions end with 2*C double atomics per block: cap the grid so that a launch issues at most ~256K of
// them (a wide, low-resolution layer would otherwise spend its time serialising atomics in L2)
static int reduce_grid(long long rows, int C) {
  int g = row_grid(rows, C);
  const int cap = (256 * 1024) / (2 * C);
  if (g > cap) g = cap;
  return g < 1 ? 1 : g;
}

int launch_bn_stats_p1(const __nv_bfloat16* z, long long rows, int C, double* sum, double* sumsq, cudaStream_t st) {
  DY_CHECK(C % 8 == 0 && kT % (C / 8) == 0 && C <= 8 * kT && rows < (1ll << 31), "channel count / rows");
  DY_CUDA(launch_kernel_pdl(p1_reduce_kernel<false>, dim3(reduce_grid(rows, C)), dim3(kT), 0, st, g_train_pdl != 0, z, nullptr, nullptr,
                            nullptr, nullptr, nullptr, 0.f, 0, (uint32_t)rows, C, sum, sumsq));
  return DY_OK;
}

int launch_bn_bwd_reduce_p1(const __nv_bfloat16* dy, const __nv_bfloat16* z, const float* a, const float* b,
                            const float* mean, const float* invstd, float alpha, int act, long long rows, int C,
                            double* s1, double* s2, cudaStream_t st) {
  DY_CHECK(C % 8 == 0 && kT % (C / 8) == 0 && C <= 8 * kT && rows < (1ll << 31), "channel count / rows");
  DY_CUDA(launch_kernel_pdl(p1_reduce_kernel<true>, dim3(reduce_grid(rows, C)), dim3(kT), 0, st, g_train_pdl != 0, dy, z, a, b, mean,
                            invstd, alpha, act, (uint32_t)rows, C, s1, s2));
  return DY_OK;
}

int launch_bn_act_p1(const __nv_bfloat16* z, const float* a, const float* b, const __nv_bfloat16* residual, int B,
                     int H, int W, int C, float alpha, int act, __nv_bfloat16* out_same, __nv_bfloat16* out_up,
                     cudaStream_t st, __nv_bfloat16* out_s2d) {
  const long long rows = (long long)B * (H + 1) * (W + 1);
  DY_CHECK(C % 8 == 0 && kT % (C / 8) == 0 && C <= 8 * kT && rows < (1ll << 31), "channel count / rows");
  BnFinalize fin;
  memset(&fin, 0, sizeof(fin));
  DY_CUDA(launch_kernel_pdl(bn_act_p1_kernel, dim3(image_row_grid((long long)B * H)), dim3(kT), 0, st, g_train_pdl != 0, z, a, b, fin,
                            residual, B, H, W, C, alpha, act, out_same, out_up, out_s2d));
  return DY_OK;
}

int launch_bn_finalize_act_p1(const __nv_bfloat16* z, const double* sum, const double* sumsq, long long M,
                              const float* gamma, const float* beta, float eps, float* a, float* b, float* mean,
                              float* var, float* invstd, const __nv_bfloat16* residual, int B, int H, int W, int C,
                              float alpha, int act, __nv_bfloat16* out_same, __nv_bfloat16* out_up, cudaStream_t st,
                              __nv_bfloat16* out_s2d) {
  const long long rows = (long long)B * (H + 1) * (W + 1);
  DY_CHECK(C % 8 == 0 && kT % (C / 8) == 0 && C <= 8 * kT && rows < (1ll << 31), "channel count / rows");
  BnFinalize fin{sum, sumsq, gamma, beta, a, b, mean, var, invstd, M, eps};
  DY_CUDA(launch_kernel_pdl(bn_act_p1_kernel, dim3(image_row_grid((long long)B * H)), dim3(kT), 0, st, g_train_pdl != 0, z, nullptr,
                            nullptr, fin, residual, B, H, W, C, alpha, act, out_same, out_up, out_s2d));
  return DY_OK;
}

static long long round_up64(long long r) { return (r + 63) / 64 * 64; }

int launch_bn_bwd_apply_p1(const __nv_bfloat16* dy, const __nv_bfloat16* z, const float* a, const float* b,
                           const float* mean, const float* invstd, const float* gamma, const double* s1,
                           const double* s2, float alpha, int act, int mode, int B, int H, int W, int C,
                           __nv_bfloat16* dz, float* dgamma, float* dbeta, cudaStream_t st) {
  DY_CHECK(C % 8 == 0, "channel count");
  const long long rows = (long long)B * (H + 1) * (W + 1);
  const long long ro = round_up64(rows);
  DY_CHECK(kT % (C / 8) == 0 && C <= 8 * kT && ro < (1ll << 31), "channel count / rows");
  DY_CUDA(launch_kernel_pdl(bn_bwd_apply_p1_kernel, dim3(image_row_grid((long long)B * (H + 1))), dim3(kT), 0, st, g_train_pdl != 0, dy,
                            z, a, b, mean, invstd, gamma, s1, s2, alpha, act, mode, (uint32_t)rows, (uint32_t)ro,
                            (long long)B * H * W, B, H, W, C, dz, dgamma, dbeta));
  return DY_OK;
}

int launch_f32_to_p1(const float* src, int B, int H, int W, int C, __nv_bfloat16* dst, int Cg, float* colsum,
                     cudaStream_t st) {
  DY_CHECK(Cg % 8 == 0 && Cg >= C && (colsum == nullptr || C <= 64), "channel count");
  const long long rows = (long long)B * (H + 1) * (W + 1);
  const long long ro = round_up64(rows);
  if (colsum) DY_CUDA(cudaMemsetAsync(colsum, 0, (size_t)C * 4, st));
  DY_CUDA(launch_kernel_pdl(f32_to_p1_kernel, dim3(grid_for(ro * (Cg / 8))), dim3(kT), 0, st, g_train_pdl != 0, src, rows, ro, H, W, C,
                            Cg, dst, colsum));
  return DY_OK;
}

int launch_pool2x2_p1(const __nv_bfloat16* src, int B, int h, int w, int C, __nv_bfloat16* dst, cudaStream_t st) {
  DY_CHECK(C % 8 == 0, "channel count");
  const long long rows = (long long)B * (h + 1) * (w + 1);
  const long long ro = round_up64(rows);
  DY_CUDA(launch_kernel_pdl(pool2x2_p1_kernel, dim3(grid_for(ro * (C / 8))), dim3(kT), 0, st, g_train_pdl != 0, src, rows, ro, h, w, C,
                            dst));
  return DY_OK;
}

int launch_add_p1(__nv_bfloat16* dst, const __nv_bfloat16* src, long long n, cudaStream_t st) {
  DY_CHECK(n % 8 == 0, "element count");
  DY_CUDA(launch_kernel_pdl(add_p1_kernel, dim3(grid_for(n / 8)), dim3(kT), 0, st, g_train_pdl != 0, dst, src, n / 8, 0));
  return DY_OK;
}

int launch_copy_p1(__nv_bfloat16* dst, const __nv_bfloat16* src, long long n, cudaStream_t st) {
  DY_CHECK(n % 8 == 0, "element count");
  DY_CUDA(launch_kernel_pdl(add_p1_kernel, dim3(grid_for(n / 8)), dim3(kT), 0, st, g_train_pdl != 0, dst, src, n / 8, 1));
  return DY_OK;
}

int launch_pack_fwd_bf16(const float* w, int K, int cout, int cout_pad, __nv_bfloat16* out, cudaStream_t st) {
  dim3 grid((K + 31) / 32, (cout_pad + 31) / 32), block(32, 8);
  pack_fwd_kernel<<<grid, block, 0, st>>>(w, K, cout, cout_pad, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_pack_dgrad_bf16(const float* w, int k, int cin_total, int ci0, int cin_sel, int cout, int Cg,
                           __nv_bfloat16* out, cudaStream_t st) {
  pack_dgrad_kernel<<<grid_for((long long)cin_sel * k * k * Cg), kT, 0, st>>>(w, k, cin_total, ci0, cin_sel, cout, Cg,
                                                                               out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_pack_dgrad_s2_bf16(const float* w, int cin, int cout, int Cg, __nv_bfloat16* out, cudaStream_t st) {
  pack_dgrad_s2_kernel<<<grid_for((long long)cin * 9 * Cg), kT, 0, st>>>(w, cin, cout, Cg, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_conv1_wgrad(const float* images, const __nv_bfloat16* dz, int B, int H, int W, float* dw, cudaStream_t st) {
  DY_CUDA(cudaMemsetAsync(dw, 0, 27 * 32 * 4, st));
  conv1_wgrad_kernel<<<148 * 8, 216, 0, st>>>(images, dz, B, H, W, dw);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_pack_multi(const PackSeg* segs_dev, int nseg, int max_tiles, cudaStream_t st) {
  if (nseg <= 0) return DY_OK;
  pack_fwd_multi_kernel<<<dim3(max_tiles, nseg), dim3(32, 8), 0, st>>>(segs_dev);
  DY_CUDA(cudaGetLastError());
  pack_dgrad_multi_kernel<<<dim3(148, 2 * nseg), kT, 0, st>>>(segs_dev);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int build_wgrad_plan(const __nv_bfloat16* x0, int c0, const __nv_bfloat16* x1, int c1, const __nv_bfloat16* dz,
                     int zc, int cout, int k, int H, int W, long long rows_max, float* dw, WgradPlan* plan, int stride) {
  DY_CHECK(k == 1 || k == 3, "kernel size");
  DY_CHECK(stride == 1 || (stride == 2 && k == 3 && c1 == 0), "stride 2 only for 3x3 without concat");
  DY_CHECK(c0 % 32 == 0 && c1 % 32 == 0 && zc % 32 == 0 && c0 > 0, "channel counts must be multiples of 32");
  DY_CHECK(c1 == 0 || k == 1, "concat only feeds 1x1 convs");
  DY_CHECK(cout <= zc, "dz must hold at least cout channels");
  WgradParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.nsrc = c1 > 0 ? 2 : 1;
  p.src_c[0] = c0; p.src_c[1] = c1;
  p.src_aw[0] = c0 % 64 == 0 ? 64 : 32;
  p.src_aw[1] = (c1 > 0 && c1 % 64 == 0) ? 64 : 32;
  p.src_koff[0] = 0; p.src_koff[1] = c0;
  p.src_blk[0] = (c0 + 127) / 128;
  p.src_blk[1] = (c1 + 127) / 128;
  p.ntap = k * k;
  const int Wp = W + 1;
  for (int kh = 0; kh < k; ++kh)
    for (int kw = 0; kw < k; ++kw) {
      if (stride == 2) {     // TF 'SAME', even input: in (2oy+kh, 2ox+kw) = s2d pixel (oy+(kh>>1), ox+(kw>>1)), block (kh&1, kw&1)
        p.tap_shift[kh * k + kw] = (kh >> 1) * Wp + (kw >> 1);
        p.tap_col[kh * k + kw] = (((kh & 1) << 1) | (kw & 1)) * c0;
      } else {
        p.tap_shift[kh * k + kw] = k == 3 ? (kh - 1) * Wp + (kw - 1) : 0;
      }
    }
  p.cin_total = c0 + c1;
  p.cout = cout;
  p.zc = zc;
  p.z_aw = zc % 64 == 0 ? 64 : 32;
  // 3x3: the three horizontal taps of a kernel row share one dz box and one halo'd X box (three
  // accumulators, 3 * block_n <= 512 TMEM columns) -- X and dz are read from L2 3 times instead of 9
  // Measured (ncu, batch 16): fusing wins where it keeps the N tile (dz <= 128 channels: conv81 912 -> 567 us,
  // conv78 248 -> 154 us) and loses where it would halve it (dz >= 256 channels); option value 2 forces it.
  p.fuse_kw = (k == 3 && stride == 1 && (g_wgrad_fuse == 2 || (g_wgrad_fuse == 1 && zc <= 128))) ? 1 : 0;
  const int bn_max = p.fuse_kw ? 128 : 256;
  p.block_n = zc >= bn_max ? bn_max : zc;    // zc in {32, 64, 128, 256, 512, 1024}
  DY_CHECK(zc % p.block_n == 0 && p.block_n % p.z_aw == 0 && p.block_n % 16 == 0, "dz channel tiling");
  p.n_tiles_n = zc / p.block_n;
  const int a_rows = p.fuse_kw ? kWgradHaloRows : kWgradRows;
  const size_t stage = 128 * (size_t)a_rows * 2 + (size_t)p.block_n * kWgradRows * 2;
  int ns = (int)((200 * 1024) / stage);
  if (ns > kWgradMaxStages) ns = kWgradMaxStages;
  p.num_stages = ns;
  p.dw = dw;
  const int x0_cols = stride == 2 ? 4 * c0 : c0;
  DY_TRY(make_tmap_2d(&plan->x[0], x0, rows_max, x0_cols, x0_cols, p.src_aw[0], a_rows));
  if (c1 > 0) DY_TRY(make_tmap_2d(&plan->x[1], x1, rows_max, c1, c1, p.src_aw[1], a_rows));
  else plan->x[1] = plan->x[0];
  DY_TRY(make_tmap_2d(&plan->z, dz, rows_max, zc, zc, p.z_aw, kWgradRows));
  return DY_OK;
}

int run_wgrad_plan(WgradPlan& plan, int B, int H, int W, int num_sms, cudaStream_t st) {
  WgradParams& p = plan.p;
  const long long M = (long long)B * (H + 1) * (W + 1);
  DY_CHECK(M < (1ll << 31) - 256, "too many rows");
  p.M = (int)M;
  p.total_chunks = (int)((M + kWgradRows - 1) / kWgradRows);
  const int units = (p.fuse_kw ? 3 : p.ntap) * (p.src_blk[0] + (p.nsrc > 1 ? p.src_blk[1] : 0)) * p.n_tiles_n;
  // split K only when the (tap, channel block, N tile) units alone leave SMs idle: every split costs a
  // pipeline fill and turns the epilogue's stores into atomics
  // (floor: units * ksplit must not exceed the SM count, or two CTAs would queue for a second wave)
  int ksplit = units * 4 >= num_sms * 3 ? 1 : num_sms / units;
  if (ksplit < 1) ksplit = 1;
  // keep at least 8 chunks (512 pixel rows) per CTA so that the pipeline fill is amortised
  const int max_split = (p.total_chunks + 7) / 8;
  if (ksplit > max_split) ksplit = max_split;
  if (ksplit < 1) ksplit = 1;
  p.chunks_per_split = (p.total_chunks + ksplit - 1) / ksplit;
  p.ksplit = (p.total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;     // <= ksplit
  const size_t stage = 128 * (size_t)(p.fuse_kw ? kWgradHaloRows : kWgradRows) * 2 + (size_t)p.block_n * kWgradRows * 2;
  const size_t smem = (size_t)p.num_stages * stage + 1024;
  static bool attr = false;
  if (!attr) {
    DY_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    attr = true;
  }
  DY_CUDA(launch_kernel_pdl(wgrad_tc_kernel, dim3(units * p.ksplit), dim3(kT), smem, st, g_train_pdl != 0, plan.x[0], plan.x[1], plan.z,
                            p, g_wdbg[0], g_wdbg[1], g_wdbg[2], g_wdbg[3]));
  return DY_OK;
}

}  // namespace dy
