// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel (see conv_tc.cuh for the scheme).
//
// CTA = 256 threads, 1 CTA per SM, persistent over output tiles (128 pixels x block_n channels):
//   warp 0      TMA producer   (one elected lane): A tile [128 x KCHUNK] + B tile [block_n x KCHUNK]
//                              per pipeline stage, 128B/64B hardware swizzle
//   warp 1      MMA issuer     (one elected lane): tcgen05.mma kind::f16, M=128, N=block_n, K=16,
//                              fp32 accumulators in TMEM, double buffered across tiles
//   warp 2      TMEM allocator
//   warps 4..7  epilogue: tcgen05.ld -> folded BN scale/shift -> leaky -> (+residual) -> bf16/fp32
//                              stores in up to two destination layouts (same / space-to-depth /
//                              2x-upsampled / fp32 head layouts); pad pixels are never written.
#include "conv_tc.cuh"

namespace dy {

namespace {

struct PixelInfo {
  int n, y, x;
  bool valid;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const float (&v)[16]) {
  uint4 a, b;
  a.x = pack_bf16(v[0], v[1]);
  a.y = pack_bf16(v[2], v[3]);
  a.z = pack_bf16(v[4], v[5]);
  a.w = pack_bf16(v[6], v[7]);
  b.x = pack_bf16(v[8], v[9]);
  b.y = pack_bf16(v[10], v[11]);
  b.z = pack_bf16(v[12], v[13]);
  b.w = pack_bf16(v[14], v[15]);
  reinterpret_cast<uint4*>(dst)[0] = a;
  reinterpret_cast<uint4*>(dst)[1] = b;
}

__device__ __forceinline__ void write_out16(const ConvParams& p, const OutDesc& o, const PixelInfo& px,
                                            long long m, int gcol, const float (&v)[16]) {
  switch (o.mode) {
    case OUT_SAME: {
      store16_bf16(reinterpret_cast<__nv_bfloat16*>(o.ptr) + m * o.ld + gcol, v);
      break;
    }
    case OUT_S2D: {
      const int Hq = p.H / 2 + 1, Wq = p.W / 2 + 1;
      long long r = ((long long)px.n * Hq + (px.y >> 1)) * Wq + (px.x >> 1);
      int cb = (((px.y & 1) << 1) | (px.x & 1)) * p.cout;
      store16_bf16(reinterpret_cast<__nv_bfloat16*>(o.ptr) + r * o.ld + cb + gcol, v);
      break;
    }
    case OUT_UP2: {
      const int Hu = 2 * p.H + 1, Wu = 2 * p.W + 1;
      long long r = ((long long)px.n * Hu + 2 * px.y) * Wu + 2 * px.x;
      __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(o.ptr) + gcol;
      store16_bf16(b + r * o.ld, v);
      store16_bf16(b + (r + 1) * o.ld, v);
      store16_bf16(b + (r + Wu) * o.ld, v);
      store16_bf16(b + (r + Wu + 1) * o.ld, v);
      break;
    }
    case OUT_F32_COMPACT: {
      float* d = reinterpret_cast<float*>(o.ptr) + (((long long)px.n * p.H + px.y) * p.W + px.x) * o.ld;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (gcol + j < p.cout) d[gcol + j] = v[j];
      break;
    }
    case OUT_F32_PLANAR: {
      float* d = reinterpret_cast<float*>(o.ptr);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (gcol + j < p.cout)
          d[(((long long)px.n * p.cout + gcol + j) * p.H + px.y) * p.W + px.x] = v[j];
      break;
    }
    default:
      break;
  }
}

template <int KCHUNK>
__global__ void __launch_bounds__(256, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ ConvParams p) {
  static_assert(KCHUNK == 64 || KCHUNK == 32, "K chunk is one 128B or 64B swizzle row");
  constexpr uint32_t kLayout = (KCHUNK == 64) ? 2u : 4u;          // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t kSBO = (KCHUNK == 64) ? 1024u : 512u;        // 8 rows of the swizzle atom
  constexpr uint32_t kABytes = kBlockM * KCHUNK * 2;

  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b_bytes = (uint32_t)p.block_n * KCHUNK * 2;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // swizzle atoms need 1024B alignment
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const int num_tiles = p.n_tiles_m * p.n_tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);   // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles_n) * kBlockM;
        const int n0 = (tile % p.n_tiles_n) * p.block_n;
        int kb = 0;
        for (int s = 0; s < p.num_seg; ++s) {
          const ConvSeg sg = p.seg[s];
          const CUtensorMap* am = sg.map ? &mapA1 : &mapA0;
          for (int c = 0; c < sg.nchunk; ++c, ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_expect_tx(&full_bar[stage], stage_bytes);
            uint8_t* sa = smem_gen + (size_t)stage * stage_bytes;
            tma_load_2d(sa, am, &full_bar[stage], sg.col0 + c * KCHUNK, m0 + sg.shift);
            tma_load_2d(sa + kABytes, &mapB, &full_bar[stage], kb * KCHUNK, n0);
            if (++stage == p.num_stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);    // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
        for (int kb = 0; kb < p.num_chunks; ++kb) {
          mbar_wait(&full_bar[stage], phase);           // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          const uint64_t adesc = umma_desc(sa, kSBO, kLayout);
          const uint64_t bdesc = umma_desc(sa + kABytes, kSBO, kLayout);
#pragma unroll
          for (int k = 0; k < KCHUNK / 16; ++k) {
            // +32 bytes (16 bf16) along K inside the swizzle row: +2 in the (addr>>4) field
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                      (uint32_t)((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);               // frees the smem slot when the MMAs retire
          if (++stage == p.num_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[acc]);                   // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int q = warp - 4;                             // TMEM lane quarter == warp % 4
    const int Hp = p.H + 1, Wp = p.W + 1;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int m0 = (tile / p.n_tiles_n) * kBlockM;
      const int n0 = (tile % p.n_tiles_n) * p.block_n;
      const long long m = (long long)m0 + q * 32 + lane;
      PixelInfo px;
      {
        int x = (int)(m % Wp);
        long long t = m / Wp;
        px.x = x;
        px.y = (int)(t % Hp);
        px.n = (int)(t / Hp);
        px.valid = (m < p.M) && (x < p.W) && (px.y < p.H);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (px.valid) {
          const int gcol = n0 + c0;
          float v[16];
          const float4* sc4 = reinterpret_cast<const float4*>(p.scale + gcol);
          const float4* sh4 = reinterpret_cast<const float4*>(p.shift + gcol);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 sc = __ldg(sc4 + j4);
            const float4 sh = __ldg(sh4 + j4);
            v[4 * j4 + 0] = fmaf(__uint_as_float(r[4 * j4 + 0]), sc.x, sh.x);
            v[4 * j4 + 1] = fmaf(__uint_as_float(r[4 * j4 + 1]), sc.y, sh.y);
            v[4 * j4 + 2] = fmaf(__uint_as_float(r[4 * j4 + 2]), sc.z, sh.z);
            v[4 * j4 + 3] = fmaf(__uint_as_float(r[4 * j4 + 3]), sc.w, sh.w);
          }
          if (p.act) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(p.alpha * v[j], v[j]);
          }
          if (p.residual != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + m * p.res_ld + gcol);
            const uint4 ra = __ldg(rp), rb = __ldg(rp + 1);
            const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
              const float2 f = __bfloat1622float2(h);
              v[2 * j] += f.x;
              v[2 * j + 1] += f.y;
            }
          }
          write_out16(p, p.out[0], px, m, gcol, v);
          if (p.out[1].mode != OUT_NONE) write_out16(p, p.out[1], px, m, gcol, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

}  // namespace

int make_tmap_2d(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld_elems,
                 int box_cols, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return DY_ERR_CUDA;
  }
  DY_CHECK(box_cols == 64 || box_cols == 32, "TMA box inner extent must be 64 or 32 bf16");
  DY_CHECK(box_rows >= 1 && box_rows <= 256, "TMA box rows");
  DY_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  DY_CHECK((ld_elems * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = (box_cols == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return DY_ERR_CUDA;
  }
  return DY_OK;
}

// dynamic smem ceiling: 227 KB per CTA minus the kernel's static shared memory, rounded down
static constexpr int kConvTcMaxSmem = 224 * 1024;

size_t conv_tc_smem_bytes(int kchunk, int block_n, int stages) {
  return (size_t)stages * (size_t)(kBlockM + block_n) * kchunk * 2 + 1024;
}

int conv_tc_pick_stages(int kchunk, int block_n) {
  const size_t budget = 200 * 1024;
  size_t per = (size_t)(kBlockM + block_n) * kchunk * 2;
  int s = (int)(budget / per);
  if (s > kMaxStages) s = kMaxStages;
  if (s < 2) s = 2;
  return s;
}

int launch_conv_tc(int kchunk, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                   const ConvParams& p, int num_sms, cudaStream_t stream) {
  DY_CHECK(kchunk == 64 || kchunk == 32, "kchunk");
  DY_CHECK(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "block_n");
  DY_CHECK(p.num_stages >= 2 && p.num_stages <= kMaxStages, "stages");
  DY_CHECK(p.tmem_cols >= 2 * p.block_n && p.tmem_cols <= 512, "tmem_cols");
  const size_t smem = conv_tc_smem_bytes(kchunk, p.block_n, p.num_stages);
  DY_CHECK(smem <= (size_t)kConvTcMaxSmem, "pipeline does not fit in shared memory");
  const int tiles = p.n_tiles_m * p.n_tiles_n;
  const int grid = tiles < num_sms ? tiles : num_sms;
  if (kchunk == 64) {
    static bool attr64 = false;
    if (!attr64) {
      DY_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvTcMaxSmem));
      attr64 = true;
    }
    conv_tc_kernel<64><<<grid, 256, smem, stream>>>(a0, a1, b, p);
  } else {
    static bool attr32 = false;
    if (!attr32) {
      DY_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvTcMaxSmem));
      attr32 = true;
    }
    conv_tc_kernel<32><<<grid, 256, smem, stream>>>(a0, a1, b, p);
  }
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

}  // namespace dy
