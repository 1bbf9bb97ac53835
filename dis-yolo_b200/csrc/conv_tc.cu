// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel (see conv_tc.cuh for the scheme).
//
// CTA = 384 threads (512 in the three-warpgroup instantiation), 1 CTA per SM, persistent over output tiles
// (128 pixels x block_n channels; a CTA PAIR works on 256 pixels x block_n channels, see conv_tc_kernel):
//   warp 0      TMA producer   (one elected lane): A box [128 or 136 x KCHUNK] + the weight chunk(s) [block_n x KCHUNK]
//                              of a pipeline stage, 128B/64B hardware swizzle
//   warp 1      MMA issuer     (one elected lane): tcgen05.mma kind::f16, M=128 (cta_group::2: 256), N=block_n, K=16,
//                              fp32 accumulators in TMEM, 2-6 stages across tiles
//   warp 2      TMEM allocator; second TMA producer when two MMA issuers split the stage ring (thin tiles)
//   warp 3      second MMA issuer (thin tiles)
//   warps 4..   epilogue warpgroups: tcgen05.ld -> folded BN scale/shift -> leaky -> (+residual) -> bf16/fp32
//                              through swizzled staging buffers and TMA stores (or direct / cooperative stores) in up
//                              to two destination layouts (same / space-to-depth / 2x-upsampled / fp32 head
//                              layouts); pad pixels are never written.
#include "conv_tc.cuh"

#include <type_traits>

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

// v[0..15] += 16 bf16 residual values (two 16-byte vectors); bf16 -> fp32 is a 16-bit shift / mask
__device__ __forceinline__ void add_res16(float (&v)[16], const uint4& ra, const uint4& rb) {
  const uint32_t rw[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 f = make_float2(__uint_as_float(rw[j] << 16), __uint_as_float(rw[j] & 0xFFFF0000u));
#ifdef DY_SCALAR_EPILOGUE
    v[2 * j] += f.x;
    v[2 * j + 1] += f.y;
#else
    const float2 o = __fadd2_rn(make_float2(v[2 * j], v[2 * j + 1]), f);
    v[2 * j] = o.x;
    v[2 * j + 1] = o.y;
#endif
  }
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
    case OUT_UNS2D:
    case OUT_UNS2D_ACC: {
      const int Hu = 2 * p.H + 1, Wu = 2 * p.W + 1;
      const long long r = ((long long)px.n * Hu + 2 * px.y + (o.blk >> 1)) * Wu + 2 * px.x + (o.blk & 1);
      __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(o.ptr) + r * o.ld + gcol;
      if (o.mode == OUT_UNS2D_ACC) {
        float w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = v[j];
        add_res16(w, reinterpret_cast<const uint4*>(d)[0], reinterpret_cast<const uint4*>(d)[1]);
        store16_bf16(d, w);
      } else {
        store16_bf16(d, v);
      }
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

// 4 role warps (producer, MMA, TMEM-alloc / second producer, second MMA issuer) + NWG epilogue warpgroups of 4 warps
constexpr int conv_threads(int nwg) { return 128 * (1 + nwg); }

// folded BN scale/shift (+ leaky) on one 16-column accumulator chunk.  Packed fp32x2 arithmetic
// (FFMA2 / FMUL2, sm_100): each lane of a pair is an ordinary IEEE round-to-nearest fma / mul, so the
// results are bit-identical to the scalar form at half the issue slots -- the epilogue warps (2 per SM
// sub-partition) are issue/latency bound on the thin layers.
// 16-byte shared-memory read through a 32-bit shared-window address (a generic pointer to a __shared__ array costs a
// window-base computation per use once the register allocator starts rematerialising it)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void bn_act16(const ConvParams& p, const uint32_t (&r)[16], uint32_t s_scale,
                                         uint32_t s_shift, int c0, float (&v)[16]) {
#ifdef DY_SCALAR_EPILOGUE      // A/B build (scripts/build_variant.sh): the scalar form
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 sc = lds128(s_scale + (uint32_t)(c0 + 4 * j4) * 4u);
    const float4 sh = lds128(s_shift + (uint32_t)(c0 + 4 * j4) * 4u);
    v[4 * j4 + 0] = fmaf(__uint_as_float(r[4 * j4 + 0]), sc.x, sh.x);
    v[4 * j4 + 1] = fmaf(__uint_as_float(r[4 * j4 + 1]), sc.y, sh.y);
    v[4 * j4 + 2] = fmaf(__uint_as_float(r[4 * j4 + 2]), sc.z, sh.z);
    v[4 * j4 + 3] = fmaf(__uint_as_float(r[4 * j4 + 3]), sc.w, sh.w);
  }
  if (p.act) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(p.alpha * v[j], v[j]);
  }
  return;
#endif
  // leaky: max(alpha*x, x); a linear layer uses alpha' = 1 (max(x, x) = x) -- no branch in the per-chunk code (the
  // ncu source page showed ~100 issued instructions per 16-column chunk, a third of them control / address overhead)
  const float alf = p.act ? p.alpha : 1.f;
  const float2 al = make_float2(alf, alf);
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 sc = lds128(s_scale + (uint32_t)(c0 + 4 * j4) * 4u);   // smem broadcast
    const float4 sh = lds128(s_shift + (uint32_t)(c0 + 4 * j4) * 4u);
    float2 a = __ffma2_rn(make_float2(__uint_as_float(r[4 * j4 + 0]), __uint_as_float(r[4 * j4 + 1])),
                          make_float2(sc.x, sc.y), make_float2(sh.x, sh.y));
    float2 b = __ffma2_rn(make_float2(__uint_as_float(r[4 * j4 + 2]), __uint_as_float(r[4 * j4 + 3])),
                          make_float2(sc.z, sc.w), make_float2(sh.z, sh.w));
    {
      const float2 ta = __fmul2_rn(a, al), tb = __fmul2_rn(b, al);
      a.x = fmaxf(ta.x, a.x); a.y = fmaxf(ta.y, a.y);
      b.x = fmaxf(tb.x, b.x); b.y = fmaxf(tb.y, b.y);
    }
    v[4 * j4 + 0] = a.x; v[4 * j4 + 1] = a.y; v[4 * j4 + 2] = b.x; v[4 * j4 + 3] = b.y;
  }
}

// residual of this chunk (16 bf16 = 2 x 16B) -> registers; issued one chunk ahead of use
__device__ __forceinline__ void load_res16(const __nv_bfloat16* rp, uint4& a, uint4& b) {
  const uint4* q = reinterpret_cast<const uint4*>(rp);
  a = __ldg(q);
  b = __ldg(q + 1);
}

// direct path: the row's owner thread adds its residual and writes its own pixels
__device__ __forceinline__ void epilogue_chunk_direct(const ConvParams& p, const PixelInfo& px, long long m,
                                                      int gcol, const uint32_t (&r)[16], uint32_t s_scale,
                                                      uint32_t s_shift, int c0, bool has_res, const uint4& ra,
                                                      const uint4& rb) {
  float v[16];
  bn_act16(p, r, s_scale, s_shift, c0, v);
  if (has_res) {
    add_res16(v, ra, rb);
  }
  write_out16(p, p.out[0], px, m, gcol, v);
  if (p.out[1].mode != OUT_NONE) write_out16(p, p.out[1], px, m, gcol, v);
}

// staged path, phase 1: (+ residual already sitting in the staging row) -> bf16 -> staging row
__device__ __forceinline__ void epilogue_chunk_staged(const ConvParams& p, const uint32_t (&r)[16],
                                                      uint32_t s_scale, uint32_t s_shift, int c0,
                                                      bool has_res, uint8_t* srow /* row base + 2*col */) {
  float v[16];
  bn_act16(p, r, s_scale, s_shift, c0, v);
  uint4* q = reinterpret_cast<uint4*>(srow);
  if (has_res) {
    const uint4 ra = q[0], rb = q[1];
    add_res16(v, ra, rb);
  }
  uint4 a, b;
  a.x = pack_bf16(v[0], v[1]);   a.y = pack_bf16(v[2], v[3]);
  a.z = pack_bf16(v[4], v[5]);   a.w = pack_bf16(v[6], v[7]);
  b.x = pack_bf16(v[8], v[9]);   b.y = pack_bf16(v[10], v[11]);
  b.z = pack_bf16(v[12], v[13]); b.w = pack_bf16(v[14], v[15]);
  q[0] = a;
  q[1] = b;
}

// TMA staged path, phase 1: 16 columns = two 16-byte chunks (index ch0, ch0+1) of a swizzled row;
// (+ residual already in place) -> bf16 in place; pad pixels / rows beyond M become zeros because
// the TMA store writes every row of the box
template <bool RES>
__device__ __forceinline__ void epilogue_chunk_swz(const ConvParams& p, const uint32_t (&r)[16],
                                                   uint32_t s_scale, uint32_t s_shift, int c0,
                                                   bool valid, uint32_t srow_a, int ch0, int rsw) {
  float v[16];
  bn_act16(p, r, s_scale, s_shift, c0, v);
  // (32-bit shared-window addresses; ch0 is even, so chunk ch0+1 of the swizzled row is the neighbour 16 bytes: ^ 16)
  const uint32_t qa = srow_a + (uint32_t)((ch0 ^ rsw) << 4), qb = qa ^ 16u;
  if constexpr (RES) {
    const uint4 ra = lds128u(qa), rb = lds128u(qb);
    add_res16(v, ra, rb);
  }
  uint4 a = make_uint4(0, 0, 0, 0), b = a;
  if (valid) {
    a.x = pack_bf16(v[0], v[1]);   a.y = pack_bf16(v[2], v[3]);
    a.z = pack_bf16(v[4], v[5]);   a.w = pack_bf16(v[6], v[7]);
    b.x = pack_bf16(v[8], v[9]);   b.y = pack_bf16(v[10], v[11]);
    b.z = pack_bf16(v[12], v[13]); b.w = pack_bf16(v[14], v[15]);
  }
  sts128u(qa, a);
  sts128u(qb, b);
}

// destination element offset of a pixel's row in an output form (without the channel), or -1
__device__ __forceinline__ long long dest_offset(const ConvParams& p, const OutDesc& o, const PixelInfo& px,
                                                 long long m) {
  if (!px.valid) return -1;
  switch (o.mode) {
    case OUT_SAME:
      return m * o.ld;
    case OUT_S2D: {
      const int Hq = p.H / 2 + 1, Wq = p.W / 2 + 1;
      const long long r = ((long long)px.n * Hq + (px.y >> 1)) * Wq + (px.x >> 1);
      return r * o.ld + (((px.y & 1) << 1) | (px.x & 1)) * p.cout;
    }
    case OUT_UP2: {
      const int Hu = 2 * p.H + 1, Wu = 2 * p.W + 1;
      return (((long long)px.n * Hu + 2 * px.y) * Wu + 2 * px.x) * o.ld;
    }
    default:
      return -1;
  }
}

// staged path, phase 2: one 16-byte vector (8 channels) of one row -> every destination form
__device__ __forceinline__ void store_vec(const ConvParams& p, const OutDesc& o, long long off, int gcol,
                                          const uint4& v) {
  __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(o.ptr) + off + gcol;
  *reinterpret_cast<uint4*>(d) = v;
  if (o.mode == OUT_UP2) {
    const long long Wu = 2 * p.W + 1;
    *reinterpret_cast<uint4*>(d + o.ld) = v;
    *reinterpret_cast<uint4*>(d + Wu * o.ld) = v;
    *reinterpret_cast<uint4*>(d + (Wu + 1) * o.ld) = v;
  }
}

// CTA2: the grid is made of 2-CTA clusters (one TPC each).  A cluster works on a PAIR tile of 256 GEMM rows x
// block_n columns: CTA r stages rows [128 r, 128 r + 128) of A and rows [r block_n/2, (r+1) block_n/2) of the weight
// tile; the even CTA's MMA thread issues tcgen05.mma.cta_group::2 (M = 256) for both, each CTA's epilogue drains its
// own 128 accumulator rows from its own TMEM.  Per SM this halves the weight bytes that cross L2 -> shared memory
// and that the tensor core reads back, and doubles the pipeline depth a given amount of shared memory buys.
// NWG: epilogue warpgroups.  2 (384 threads, 160 registers) everywhere except the thin tiles (<= 64 columns): their
// MMAs take 300-800 cycles per tile, which two warpgroups of latency-bound tcgen05.ld -> BN -> pack -> st.shared
// chains cannot keep up with; there a THIRD warpgroup (512 threads, 128 registers) takes every third tile.
template <int KCHUNK, bool FUSE, bool CTA2, int NWG>
__global__ void __launch_bounds__(conv_threads(NWG), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapR,
               const __grid_constant__ CUtensorMap mapO, const __grid_constant__ ConvParams p) {
  static_assert(KCHUNK == 64 || KCHUNK == 32, "K chunk is one 128B or 64B swizzle row");
  constexpr uint32_t kLayout = (KCHUNK == 64) ? 2u : 4u;          // SWIZZLE_128B : SWIZZLE_64B
  constexpr uint32_t kSBO = (KCHUNK == 64) ? 1024u : 512u;        // 8 rows of the swizzle atom

  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[6];
  __shared__ __align__(8) uint64_t tempty_bar[6];
  __shared__ __align__(8) uint64_t bres_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float s_scale[NWG][256];               // per epilogue warpgroup
  __shared__ __align__(16) float s_shift[NWG][256];
  __shared__ long long s_dst[NWG][2][kBlockM];                    // [warpgroup][output][row] element offset / -1
  __shared__ __align__(8) uint64_t res_full[NWG][2];              // [warpgroup][staging buffer] residual landed
  __shared__ __align__(8) uint64_t fuse_bar[NWG][2];              // [warpgroup][accumulator] fused-tail MMA complete
  __shared__ __align__(16) float s_fbias[16];
  __shared__ uint32_t tap_a16[kMaxSeg][3];                        // A start offset inside the stage, >>4
  __shared__ uint32_t tap_b16[kMaxSeg][3];                        // B start (absolute if resident, else in-stage), >>4

  pdl_launch_dependents();            // the next layer's CTAs may take over SMs as this grid's CTAs retire
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // first (pair) tile of this CTA (cluster)
  const int tstride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;    // (pair) tiles between two iterations
  constexpr int kTileM = CTA2 ? 2 * kBlockM : kBlockM;                  // GEMM rows of one (pair) tile
  const int m_off = (int)cta_rank * kBlockM;                            // this CTA's rows inside the pair tile
  const int b_rows = CTA2 ? p.block_n >> 1 : p.block_n;                 // weight rows this CTA stages
  const uint32_t b_bytes = (uint32_t)b_rows * KCHUNK * 2;               // one weight K chunk (this CTA's part)
  const uint32_t a_tx = (uint32_t)p.a_rows * KCHUNK * 2;                // bytes one A box delivers
  const uint32_t a_bytes = (a_tx + 1023u) & ~1023u;
  const uint32_t stage_bytes = a_bytes + (p.b_resident ? 0u : (uint32_t)p.max_ntap * b_bytes);
  const uint32_t bres_bytes = p.b_resident ? (uint32_t)p.num_chunks * b_bytes : 0u;
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;   // swizzle atoms need 1024B alignment
  uint8_t* smem_gen = smem_dyn + (smem_base - smem_u32(smem_dyn));     // [resident weights][stages...]
  const uint32_t stages_off = bres_bytes;
  constexpr uint32_t kRowBytes = KCHUNK * 2;
  // staging row pitch: padded for the cooperative path, dense (hardware-swizzled) for the TMA path
  const uint32_t epi_pitch = p.tma_epi ? (uint32_t)p.slab * 2 : (uint32_t)p.slab * 2 + 16;
  const uint32_t epi_off = stages_off + (uint32_t)p.num_stages * stage_bytes;
  const int num_tiles = (CTA2 ? (p.n_tiles_m + 1) >> 1 : p.n_tiles_m) * p.n_tiles_n;
  const bool has_res = p.residual != nullptr;
  // fused tail: its weights [fuse_n x 64] bf16 (SWIZZLE_128B rows) sit behind the epilogue staging buffers
  const uint32_t fuse_off = epi_off + (uint32_t)NWG * 2u * (uint32_t)kBlockM * 64u * 2u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    if (has_res) tma_prefetch_desc(&mapR);
    if (p.tma_epi) tma_prefetch_desc(&mapO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 6; ++i) {
      mbar_init(&tfull_bar[i], 1);
      // one arrive per epilogue warp that drains the accumulator (a CTA pair: the warps of BOTH CTAs arrive on the
      // leader's barrier, where the MMA thread waits)
      mbar_init(&tempty_bar[i], (p.split_n ? 8 : 4) * (CTA2 ? 2 : 1));
    }
    mbar_init(&bres_bar, 1);
    for (int i = 0; i < 2 * NWG; ++i) mbar_init(&res_full[i >> 1][i & 1], 1);
    for (int i = 0; i < 2 * NWG; ++i) mbar_init(&fuse_bar[i >> 1][i & 1], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CTA2) {
      tmem_alloc_cta2(&tmem_base_smem, (uint32_t)p.tmem_cols);
      tmem_relinquish_cta2();
    } else {
      tmem_alloc(&tmem_base_smem, (uint32_t)p.tmem_cols);
      tmem_relinquish();
    }
  }
  if (warp == 3 && lane < kMaxSeg * 3) {
    // per-(segment, tap) descriptor offsets for the MMA issuers (tile-invariant)
    const int sgi = lane / 3, t = lane % 3;
    if (sgi < p.num_seg && t < p.seg[sgi].ntap) {
      tap_a16[sgi][t] = ((uint32_t)p.seg[sgi].tap_row[t] * kRowBytes) >> 4;
      tap_b16[sgi][t] = p.b_resident ? (smem_base + (uint32_t)p.seg[sgi].tap_b0[t] * b_bytes) >> 4
                                     : (a_bytes + (uint32_t)t * b_bytes) >> 4;
    }
  }
  if (FUSE && warp >= 4 && warp < 8) {
    // fused-tail weights -> shared memory in the canonical K-major SWIZZLE_128B layout (row = output
    // channel, 128 B = 64 input channels; 16-byte chunk c of row r lands at chunk c ^ (r & 7))
    const int t = threadIdx.x - 128;
    for (int i = t; i < p.fuse_n * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.fuse_w + (size_t)r * 64) + c);
      *reinterpret_cast<uint4*>(smem_gen + fuse_off + (size_t)r * 128 + ((c ^ (r & 7)) << 4)) = v;
    }
    if (t < 16) s_fbias[t] = t < p.fuse_n ? __ldg(p.fuse_bias + t) : 0.f;
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();      // the peer's mbarriers are initialised before anything arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous layer's tail;
  // from here on this grid reads what that layer wrote
  pdl_wait();

  if (warp == 0 || (warp == 2 && p.dual_issue && p.dual_producer)) {
    // ================================ TMA producers ================================
    // With dual issue the ring is split in two halves, one per MMA issuer (= per tile parity), so that
    // every mbarrier still has exactly one producer and one consumer.  With dual_producer each half also
    // has its OWN producer thread (warp 0: even tiles, warp 2: odd tiles): a producer waiting for a free
    // slot of its ring never holds back the loads of the other ring.
    if (elect_one()) {
      const bool split = p.dual_issue && p.dual_producer;
      const int my_ring = (warp == 2) ? 1 : 0;
      if (my_ring == 0 && p.b_resident) {
        // this CTA only ever sees one N tile (grid is a multiple of n_tiles_n): park its weights
        const int n0c = (int)(blockIdx.x % p.n_tiles_n) * p.block_n;
        mbar_expect_tx(&bres_bar, bres_bytes);
        for (int kb = 0; kb < p.num_chunks; ++kb)
          tma_load_2d(smem_gen + (size_t)kb * b_bytes, &mapB, &bres_bar, kb * KCHUNK, n0c);
      }
      const int ring_sz = p.dual_issue ? p.num_stages / 2 : p.num_stages;
      const int it_step = split ? 2 : 1;
      // 32-bit shared-window addresses of the barrier arrays and of the stage ring (see common.cuh)
      const uint32_t full_a = opaque_u32(smem_u32(full_bar)), empty_a = opaque_u32(smem_u32(empty_bar));
      const uint32_t stage_a = opaque_u32(smem_base + stages_off);
      int rstage[2] = {0, 0};
      uint32_t rphase[2] = {0u, 0u};
      int it = my_ring;
      for (int tile = tile0 + my_ring * tstride; tile < num_tiles; tile += it_step * tstride, it += it_step) {
        const int ring = p.dual_issue ? (it & 1) : 0;
        int stage = ring ? rstage[1] : rstage[0];
        uint32_t phase = ring ? rphase[1] : rphase[0];
        const int mt = p.n_tiles_n == 1 ? tile : tile / p.n_tiles_n;      // (integer divisions cost ~20 issue slots each)
        const int m0 = mt * kTileM + m_off;
        const int n0 = (tile - mt * p.n_tiles_n) * p.block_n;
        const int nb0 = n0 + (CTA2 ? (int)cta_rank * b_rows : 0);   // first weight row this CTA stages
        if (has_res) {
          // pull the residual tile towards L2 now; the epilogue reads it a few microseconds later
          const int step = p.slab ? p.slab : 64;
          for (int c = 0; c < p.block_n; c += step) tma_prefetch_2d(&mapR, n0 + c, m0);
        }
        for (int s = 0; s < p.num_seg; ++s) {
          const ConvSeg& sg = p.seg[s];   // stays in constant param space (dynamic tap index)
          const CUtensorMap* am = sg.map ? &mapA1 : &mapA0;
          const uint32_t tx = a_tx + (p.b_resident ? 0u : (uint32_t)sg.ntap * b_bytes);
          for (int c = 0; c < sg.nchunk; ++c) {
            const int gs = ring * ring_sz + stage;                   // global stage slot
            mbar_wait_a(empty_a + 8u * (uint32_t)gs, phase ^ 1u);
            const uint32_t sa = stage_a + (uint32_t)gs * stage_bytes, fb = full_a + 8u * (uint32_t)gs;
            if constexpr (CTA2) {
              // the leader's barrier counts the bytes of both CTAs (complete_tx may precede expect_tx within a phase)
              if (cta_rank == 0) mbar_expect_tx_a(fb, 2u * tx);
              tma_load_2d_cta2_a(sa, am, fb, sg.col0 + c * KCHUNK, m0 + sg.shift);
              for (int t = 0; t < sg.ntap; ++t)
                tma_load_2d_cta2_a(sa + a_bytes + (uint32_t)t * b_bytes, &mapB, fb, (sg.tap_b0[t] + c) * KCHUNK, nb0);
            } else {
              mbar_expect_tx_a(fb, tx);
              tma_load_2d_a(sa, am, fb, sg.col0 + c * KCHUNK, m0 + sg.shift);
              if (!p.b_resident) {
                for (int t = 0; t < sg.ntap; ++t)
                  tma_load_2d_a(sa + a_bytes + (uint32_t)t * b_bytes, &mapB, fb, (sg.tap_b0[t] + c) * KCHUNK, nb0);
              }
            }
            if (++stage == ring_sz) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        if (ring) { rstage[1] = stage; rphase[1] = phase; } else { rstage[0] = stage; rphase[0] = phase; }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ================================ MMA issuers ================================
    // Two issuing threads: issuer i owns accumulator stage i, i.e. every other tile of this CTA.
    // For thin layers (N <= 64: 32-cycle MMAs) the per-stage wait/commit and per-MMA descriptor work
    // of a single thread is the bottleneck; the smem ring is consumed in tile order either way.
    if ((warp == 1 || p.dual_issue) && (!CTA2 || cta_rank == 0) && elect_one()) {
      const int issuer = (warp == 3) ? 1 : 0;
      const int it_step = p.dual_issue ? 2 : 1;
      const uint32_t idesc = umma_idesc_bf16(p.block_n, kTileM);
      // descriptor bits that never change: LBO=1, SBO, version, layout
      const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46) | ((uint64_t)kLayout << 61);
      const uint32_t S = (uint32_t)(p.dual_issue ? p.num_stages / 2 : p.num_stages);   // this issuer's ring
      const uint32_t full_a = opaque_u32(smem_u32(full_bar)), empty_a = opaque_u32(smem_u32(empty_bar));
      const uint32_t ring_base = (uint32_t)issuer * S;
      const uint32_t b16 = b_bytes >> 4;
      if (p.b_resident) {
        mbar_wait(&bres_bar, 0);                        // resident weights have landed
        tc_fence_after();
      }
      uint32_t stage = 0, phase = 0;
      int it = issuer;
      int acc = issuer;                                  // = it % num_acc, used for the (it / num_acc)-th time (aph = its parity)
      uint32_t aph = 0u;
      for (int tile = tile0 + issuer * tstride; tile < num_tiles; tile += it_step * tstride, it += it_step) {
        // accumulator stage of tile `it`: it % num_acc, used for the (it / num_acc)-th time.  num_acc is even with two
        // issuers (issuer = parity of it = parity of acc) and a multiple of the number of epilogue warpgroups when
        // they take alternate tiles, so every accumulator always meets the same issuer and the same warpgroup (an
        // mbarrier only tells consecutive phases apart).
        // (acc / aph are advanced at the end of the iteration: no division in the loop)
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
        mbar_wait(&tempty_bar[acc], aph ^ 1u);          // epilogue has drained this accumulator
        tc_fence_after();
        uint32_t acc_flag = 0u;
        for (int s = 0; s < p.num_seg; ++s) {
          const int nchunk = p.seg[s].nchunk, ntap = p.seg[s].ntap;
          for (int c = 0; c < nchunk; ++c) {
            const uint32_t gs = ring_base + stage;
            mbar_wait_a(full_a + 8u * gs, phase);       // TMA bytes have landed
            tc_fence_after();
            // (a clustered CTA's shared-window addresses carry its rank above bit 18: keep the descriptor's 14 bits)
            const uint32_t sa16 = ((smem_base + stages_off + gs * stage_bytes) & 0x3FFFFu) >> 4;
            for (int t = 0; t < ntap; ++t) {
              // same box, start address advanced by whole rows: tap t of the shared halo'd segment
              const uint64_t adesc = desc_hi | (uint64_t)(sa16 + tap_a16[s][t]);
              const uint64_t bdesc = desc_hi | (uint64_t)(p.b_resident ? tap_b16[s][t] + (uint32_t)c * b16
                                                                       : sa16 + tap_b16[s][t]);
#pragma unroll
              for (int k = 0; k < KCHUNK / 16; ++k) {
                // +32 bytes (16 bf16) along K inside the swizzle row: +2 in the (addr>>4) field
                if constexpr (CTA2) umma_bf16_cta2(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc_flag);
                else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, acc_flag);
                acc_flag = 1u;
              }
            }
            if constexpr (CTA2) umma_commit_cta2_a(empty_a + 8u * gs);   // (both CTAs' slots)
            else umma_commit_a(empty_a + 8u * gs);      // frees the smem slot when the MMAs retire
            if (++stage == S) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        if constexpr (CTA2) umma_commit_cta2(&tfull_bar[acc]);    // (both CTAs' epilogues)
        else umma_commit(&tfull_bar[acc]);              // accumulator complete -> epilogue
        acc += it_step;                                 // it_step <= num_acc: one wrap at most
        if (acc >= p.num_acc) {
          acc -= p.num_acc;
          aph ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    // Two warpgroups; warpgroup g drains accumulator stage g, i.e. every other tile of this CTA.
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;                             // TMEM lane quarter == warp % 4
    const int wg_tid = (warp - 4 - 4 * wg) * 32 + lane; // 0..127 inside the warpgroup
    const int Hp = p.H + 1, Wp = p.W + 1;
    float* my_scale = s_scale[wg];
    float* my_shift = s_shift[wg];
    const uint32_t scale_a = smem_u32(my_scale), shift_a = smem_u32(my_shift);      // 32-bit shared-window addresses
    int cached_n0 = -1;
    uint32_t res_par = 0u;                              // bit b = parity of res_full[wg][b]; persists across tiles
    uint32_t sc = 0u;                                   // slabs processed by this warpgroup so far (buffer = sc & 1)
    // fused tail, software pipelined: the 1x1 MMA of tile i is issued at the end of tile i and its result
    // is read (and stored) while tile i+1 of this warpgroup is processed -- two 16-column accumulators
    uint32_t f_cnt = 0u;                                // fused MMAs issued by this warpgroup
    PixelInfo f_px = {0, 0, 0, false};                  // this thread's pixel of the pending fused tile
    bool res_primed = false;
    auto fuse_drain = [&]() {                           // result of fused MMA number f_cnt-1 -> global memory
      const uint32_t b = (f_cnt - 1u) & 1u;
      uint32_t fr[16];
      mbar_wait(&fuse_bar[wg][b], ((f_cnt - 1u) >> 1) & 1u);
      tc_fence_after();
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.num_acc * p.block_n + (wg * 2 + (int)b) * 16), fr);
      tmem_ld_wait();
      tc_fence_before();
      if (f_px.valid) {
        // (one pointer walking the channel planes, biases read as float4: the rolled form recomputed a 64-bit
        // address, compared and loaded a bias per channel -- ~140 issued instructions per tile and warp)
        float* d = p.fuse_out + (((long long)f_px.n * p.fuse_cout) * p.H + f_px.y) * p.W + f_px.x;
        const long long plane = (long long)p.H * p.W;
        const int nout = p.fuse_cout;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (4 * j4 >= nout) break;
          const float4 bq = *reinterpret_cast<const float4*>(&s_fbias[4 * j4]);
          const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (4 * j4 + e < nout) __stcs(d, __uint_as_float(fr[4 * j4 + e]) + bb[e]);
            d += plane;
          }
        }
      }
    };
    // split_n: BOTH warpgroups drain EVERY tile, half of its columns each -- the accumulator is handed back to
    // the MMA issuer after half the tcgen05.ld / residual round trips (a 256-column tile only has two TMEM
    // stages, so the drain time of one tile bounds the start of the tile after next)
    const int t_first = p.split_n ? 0 : wg, t_step = p.split_n ? 1 : NWG;     // (split_n: NWG == 2, checked on the host)
    const int cbase = p.split_n ? wg * (p.block_n >> 1) : 0;       // first column of this warpgroup inside the tile
    const int ncols = p.split_n ? (p.block_n >> 1) : p.block_n;    // columns this warpgroup drains
    int it = t_first;
    uint32_t my_tiles = 0u;                             // tiles this warpgroup has processed (parity of its tables)
    // The per-tile code of a thin layer is instruction-issue bound (ncu: five integer divisions of ~20 issue slots
    // each per tile and warp): accumulator index / phase and -- when this warpgroup's row offset advances by a
    // constant, i.e. with one N tile -- the pixel coordinates of the thread's row are advanced incrementally.
    int acc = t_first % p.num_acc;                      // = it % num_acc (no split: the accumulators with it % NWG == wg)
    uint32_t acc_phase = 0u;                            // = (it / num_acc) & 1
    const bool inc_px = p.n_tiles_n == 1;
    int cx = 0, cy = 0, cn = 0, dx = 0, dy = 0, dn = 0;  // coordinates of this thread's row / their per-iteration step
    if (inc_px) {
      const uint32_t mu = (uint32_t)((tile0 + t_first * tstride) * kTileM + m_off + q * 32 + lane);
      uint32_t t = mu / (uint32_t)Wp;
      cx = (int)(mu - t * (uint32_t)Wp);
      cn = (int)(t / (uint32_t)Hp);
      cy = (int)(t - (uint32_t)cn * (uint32_t)Hp);
      const uint32_t du = (uint32_t)(t_step * tstride * kTileM);
      t = du / (uint32_t)Wp;
      dx = (int)(du - t * (uint32_t)Wp);
      dn = (int)(t / (uint32_t)Hp);
      dy = (int)(t - (uint32_t)dn * (uint32_t)Hp);
    }
    for (int tile = tile0 + t_first * tstride; tile < num_tiles; tile += t_step * tstride, it += t_step, ++my_tiles) {
      const int mt = inc_px ? tile : tile / p.n_tiles_n;
      const int m0 = mt * kTileM + m_off;
      const int n0 = (tile - mt * p.n_tiles_n) * p.block_n;
      const long long m = (long long)m0 + q * 32 + lane;
      PixelInfo px;
      if (inc_px) {
        px.x = cx; px.y = cy; px.n = cn;
        cx += dx;
        if (cx >= Wp) { cx -= Wp; ++cy; }
        cy += dy;
        if (cy >= Hp) { cy -= Hp; ++cn; }
        cn += dn;
      } else {
        // 32-bit unsigned arithmetic (M < 2^31 is checked on the host)
        const uint32_t mu = (uint32_t)m;
        const uint32_t t = mu / (uint32_t)Wp;
        px.x = (int)(mu - t * (uint32_t)Wp);
        px.n = (int)(t / (uint32_t)Hp);
        px.y = (int)(t - (uint32_t)px.n * (uint32_t)Hp);
      }
      px.valid = (m < p.M) && (px.x < p.W) && (px.y < p.H);
      if (n0 != cached_n0) {
        // (re)stage this N tile's folded BN scale/shift; 128 threads of the warpgroup only
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        for (int i = wg_tid; i < p.block_n; i += 128) {
          my_scale[i] = __ldg(p.scale + n0 + i);
          my_shift[i] = __ldg(p.shift + n0 + i);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        cached_n0 = n0;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n + cbase);
      uint32_t r0[16], r1[16];
      if (p.debug_skip == 1) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
      } else if (p.slab == 0) {
        // ---------------- direct path: the row's owner thread reads/writes global memory ----------------
        const __nv_bfloat16* res_row = has_res ? p.residual + m * p.res_ld + n0 : nullptr;
        const bool do_res = has_res && px.valid;
        uint4 ra0 = make_uint4(0, 0, 0, 0), rb0 = ra0, ra1 = ra0, rb1 = ra0;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        tmem_ld16(taddr, r0);
        if (do_res) load_res16(res_row, ra0, rb0);
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          tmem_ld_wait();
          const bool more1 = c0 + 16 < p.block_n;
          if (more1) {
            tmem_ld16(taddr + (uint32_t)(c0 + 16), r1);
            if (do_res) load_res16(res_row + c0 + 16, ra1, rb1);
          }
          if (px.valid) epilogue_chunk_direct(p, px, m, n0 + c0, r0, scale_a, shift_a, c0, has_res, ra0, rb0);
          if (more1) {
            tmem_ld_wait();
            if (c0 + 32 < p.block_n) {
              tmem_ld16(taddr + (uint32_t)(c0 + 32), r0);
              if (do_res) load_res16(res_row + c0 + 32, ra0, rb0);
            }
            if (px.valid)
              epilogue_chunk_direct(p, px, m, n0 + c0 + 16, r1, scale_a, shift_a, c0 + 16, has_res, ra1, rb1);
          }
        }
      } else if (p.tma_epi) {
        // ---------------- TMA staged path ----------------
        // Two swizzled staging buffers per warpgroup used round-robin over ALL slabs of ALL tiles of
        // this warpgroup (slab counter `sc`).  The residual slab is TMA-loaded INTO its buffer one
        // slab ahead (the next tile's first slab is requested at the end of the current tile, so it
        // lands while that tile's MMAs are still running); phase 1 updates the buffer in place
        // thread-per-row (swizzle => conflict-free) and one elected thread TMA-stores it.  A buffer
        // is reused two slabs later, so only `wait_group.read 1` is ever needed: no store latency is
        // exposed.  No global load/store instruction is executed for out[0].
        const uint32_t buf_bytes = (uint32_t)kBlockM * p.slab * 2;
        uint8_t* stg0 = smem_gen + epi_off + (size_t)wg * 2 * buf_bytes;
        const int row = q * 32 + lane;
        const int rsw = (p.slab == 64) ? (row & 7) : ((row >> 1) & 3);      // this row's swizzle XOR
        const int vpr = p.slab >> 3, vsh = (p.slab == 64) ? 3 : 2;
        const int nslab = ncols >> (p.slab == 64 ? 6 : 5);   // slab is 64 or 32
        const int nb = n0 + cbase;                           // first global column of this warpgroup
        const bool elected = wg_tid == 0;
        const bool dual = p.out[1].mode != OUT_NONE;
        // Double buffered by tile parity: a warp that has finished this tile's cooperative stores writes
        // the NEXT tile's offsets while slower warps of the warpgroup are still reading this tile's (there
        // is no warpgroup barrier between two tiles; within a tile every slab has one, so nobody can be
        // two tiles ahead).
        // (the TMA path needs no table for out[0], so s_dst[wg][0..1] serve as the two parities)
        long long* dst2 = s_dst[wg][my_tiles & 1];
        if (dual) dst2[row] = dest_offset(p, p.out[1], px, m);
        if (elected && has_res && !res_primed) {
          // very first slab of this warpgroup: nothing has used the buffers yet
          mbar_expect_tx(&res_full[wg][sc & 1], buf_bytes);
          tma_load_2d(stg0 + (size_t)(sc & 1) * buf_bytes, &mapR, &res_full[wg][sc & 1], nb, m0);
        }
        res_primed = true;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        for (int sidx = 0; sidx < nslab; ++sidx, ++sc) {
          const int slab0 = sidx * p.slab;
          const int bsel = sc & 1;
          uint8_t* stg = stg0 + (size_t)bsel * buf_bytes;
          if (has_res) {
            mbar_wait(&res_full[wg][bsel], (res_par >> bsel) & 1u);
            res_par ^= 1u << bsel;
          } else {
            if (elected) bulk_wait_read<1>();                // the store two slabs ago has read this buffer
            asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
          }
          const uint32_t srow = smem_u32(stg) + (uint32_t)row * epi_pitch;
          // the chunk sequence of one slab, unswitched on (residual?, 2 or 4 chunks): no loop control, no per-chunk
          // tests of tile-invariant conditions in the issue-bound part of the epilogue
          auto run_slab = [&](auto res_tag, auto nch_tag) {
            constexpr bool RES = decltype(res_tag)::value;
            constexpr int NCH = decltype(nch_tag)::value;
            tmem_ld16(taddr + (uint32_t)slab0, r0);
#pragma unroll
            for (int ch = 0; ch < NCH; ch += 2) {
              tmem_ld_wait();
              tmem_ld16(taddr + (uint32_t)(slab0 + 16 * (ch + 1)), r1);
              epilogue_chunk_swz<RES>(p, r0, scale_a, shift_a, cbase + slab0 + 16 * ch, px.valid, srow, 2 * ch, rsw);
              tmem_ld_wait();
              if (ch + 2 < NCH) tmem_ld16(taddr + (uint32_t)(slab0 + 16 * (ch + 2)), r0);
              epilogue_chunk_swz<RES>(p, r1, scale_a, shift_a, cbase + slab0 + 16 * (ch + 1), px.valid, srow, 2 * ch + 2, rsw);
            }
          };
#ifdef DY_DEBUG_SKIP
          if (p.debug_skip == 2) {
            for (int c0 = 0; c0 < p.slab; c0 += 16) { tmem_ld16(taddr + (uint32_t)(slab0 + c0), r0); tmem_ld_wait(); }
          } else
#endif
          if (p.slab == 64) {
            if (has_res) run_slab(std::true_type{}, std::integral_constant<int, 4>{});
            else run_slab(std::false_type{}, std::integral_constant<int, 4>{});
          } else {
            if (has_res) run_slab(std::true_type{}, std::integral_constant<int, 2>{});
            else run_slab(std::false_type{}, std::integral_constant<int, 2>{});
          }
          if (sidx + 1 == nslab) {
            // every tcgen05.ld of this tile has completed: hand the accumulator back to the MMA issuer
            // now, before the store phases
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CTA2) mbar_arrive_leader(&tempty_bar[acc]);
              else mbar_arrive(&tempty_bar[acc]);
            }
          }
          fence_proxy_async_smem();                          // generic-proxy smem writes -> async proxy (TMA)
          asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
          if (elected) {
            if (p.debug_skip == 0 && (!FUSE || p.fuse_store)) {
              tma_store_2d(&mapO, stg, nb + slab0, m0);
              bulk_commit();
            }
            if constexpr (FUSE) {
              // the finished bf16 tile [128 x 64] is a K-major SWIZZLE_128B A operand as it stands
              tc_fence_after();
              const uint64_t dhi = (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
              const uint64_t fa = dhi | (uint64_t)((smem_u32(stg) & 0x3FFFFu) >> 4);
              const uint64_t fb = dhi | (uint64_t)(((smem_base + fuse_off) & 0x3FFFFu) >> 4);
              const uint32_t fd = tmem_base + (uint32_t)(p.num_acc * p.block_n + (wg * 2 + (int)(f_cnt & 1u)) * 16);
              const uint32_t fidesc = umma_idesc_bf16(p.fuse_n);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(fd, fa + (uint64_t)(2 * k), fb + (uint64_t)(2 * k), fidesc, k ? 1u : 0u);
              umma_commit(&fuse_bar[wg][f_cnt & 1u]);
            }
            if (has_res) {
              // request the residual of the NEXT slab (this tile's, or the first one of this
              // warpgroup's next tile) into the other buffer, whose last store is one slab old
              int nm0 = m0, nn0 = nb + slab0 + p.slab;
              bool have = sidx + 1 < nslab;
              if (!have) {
                const int ntile = tile + t_step * tstride;
                if (ntile < num_tiles) {
                  have = true;
                  const int nmt = inc_px ? ntile : ntile / p.n_tiles_n;
                  nm0 = nmt * kTileM + m_off;
                  nn0 = (ntile - nmt * p.n_tiles_n) * p.block_n + cbase;
                }
              }
              if (have) {
                bulk_wait_read<1>();
                mbar_expect_tx(&res_full[wg][bsel ^ 1], buf_bytes);
                tma_load_2d(stg0 + (size_t)(bsel ^ 1) * buf_bytes, &mapR, &res_full[wg][bsel ^ 1], nn0, nm0);
              }
            }
          }
          if constexpr (FUSE) {
            if (f_cnt > 0u) fuse_drain();                    // the PREVIOUS tile's 1x1 result (long complete)
            f_px = px;
            ++f_cnt;
          }
          if (dual) {
            // second destination form (space-to-depth / upsampled): cooperative vector stores
            // (vpr passes of 128 vectors; four passes at a time with all shared-memory loads ahead of the stores:
            // the dependent LDS -> LDS -> STG chain of a rolled loop was the hottest spot of the dual-output layers)
            for (int i0 = 0; i0 < vpr; i0 += 4) {
              long long d1[4];
              uint4 v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) d1[i] = dst2[((i0 + i) * 128 + wg_tid) >> vsh];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int idx = (i0 + i) * 128 + wg_tid;
                const int rr = idx >> vsh, cj = idx & (vpr - 1);
                const int sw = (p.slab == 64) ? (rr & 7) : ((rr >> 1) & 3);
                v[i] = *reinterpret_cast<const uint4*>(stg + (size_t)rr * epi_pitch + ((cj ^ sw) << 4));
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int cj = ((i0 + i) * 128 + wg_tid) & (vpr - 1);
                if (d1[i] >= 0) store_vec(p, p.out[1], d1[i], nb + slab0 + cj * 8, v[i]);
              }
            }
          }
        }
      } else {
        // ---------------- staged path: coalesced residual reads and output stores ----------------
        // Phase 1 works thread-per-row (that is how TMEM is read); global memory is touched only in
        // the cooperative phases 0/2 where consecutive lanes cover consecutive 16-byte vectors of a
        // row, so one instruction touches 4-8 cache lines instead of 32.
        uint8_t* stg = smem_gen + epi_off + (size_t)wg * kBlockM * epi_pitch;
        const int row = q * 32 + lane;
        const int vpr = p.slab >> 3;                      // 16-byte vectors per staged row (4 or 8)
        const int vsh = (p.slab == 64) ? 3 : 2;           // log2(vpr)
        const int nvec = kBlockM * vpr;                   // vectors per slab, nvec / 128 per thread
        s_dst[wg][0][row] = dest_offset(p, p.out[0], px, m);
        s_dst[wg][1][row] = (p.out[1].mode != OUT_NONE) ? dest_offset(p, p.out[1], px, m) : -1;
        uint4 rres[8];
        auto fetch_res = [&](int slab0) {                 // coalesced residual slab -> registers
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 128 + wg_tid;
            if (i < vpr) {
              const int rr = idx >> vsh, cj = idx & (vpr - 1);
              const long long mm = (long long)m0 + rr;
              rres[i] = (mm < p.M) ? __ldg(reinterpret_cast<const uint4*>(p.residual + mm * p.res_ld + n0 + cbase + slab0 +
                                                                          cj * 8))
                                   : make_uint4(0, 0, 0, 0);
            }
          }
        };
        if (has_res) fetch_res(0);
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        for (int slab0 = 0; slab0 < ncols; slab0 += p.slab) {
          if (has_res) {
            // phase 0: park this slab's residual in the staging rows, start fetching the next slab's
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int idx = i * 128 + wg_tid;
              if (i < vpr) {
                const int rr = idx >> vsh, cj = idx & (vpr - 1);
                *reinterpret_cast<uint4*>(stg + (size_t)rr * epi_pitch + cj * 16) = rres[i];
              }
            }
            if (slab0 + p.slab < ncols) fetch_res(slab0 + p.slab);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
          // phase 1: accumulator -> BN/leaky (+ residual) -> bf16, thread per row
          uint8_t* srow = stg + (size_t)row * epi_pitch;
          tmem_ld16(taddr + (uint32_t)slab0, r0);
          for (int c0 = 0; c0 < p.slab; c0 += 32) {
            tmem_ld_wait();
            const bool more1 = c0 + 16 < p.slab;
            if (more1) tmem_ld16(taddr + (uint32_t)(slab0 + c0 + 16), r1);
            epilogue_chunk_staged(p, r0, scale_a, shift_a, cbase + slab0 + c0, has_res, srow + c0 * 2);
            if (more1) {
              tmem_ld_wait();
              if (c0 + 32 < p.slab) tmem_ld16(taddr + (uint32_t)(slab0 + c0 + 32), r0);
              epilogue_chunk_staged(p, r1, scale_a, shift_a, cbase + slab0 + c0 + 16, has_res, srow + (c0 + 16) * 2);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
          // phase 2: cooperative, coalesced stores to every destination form
          for (int idx = wg_tid; idx < nvec; idx += 128) {
            const int rr = idx >> vsh, cj = idx & (vpr - 1);
            const long long d0 = s_dst[wg][0][rr];
            if (d0 < 0) continue;                         // pad pixel / beyond M: never written
            const uint4 v = *reinterpret_cast<const uint4*>(stg + (size_t)rr * epi_pitch + cj * 16);
            const int gcol = n0 + cbase + slab0 + cj * 8;
            store_vec(p, p.out[0], d0, gcol, v);
            const long long d1 = s_dst[wg][1][rr];
            if (d1 >= 0) store_vec(p, p.out[1], d1, gcol, v);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        }
      }
      if (!(p.tma_epi && p.slab != 0 && p.debug_skip != 1)) {   // (the TMA path released it after its last tcgen05.ld)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CTA2) mbar_arrive_leader(&tempty_bar[acc]);
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
      acc += t_step;                                    // t_step <= num_acc: one wrap at most
      if (acc >= p.num_acc) {
        acc -= p.num_acc;
        acc_phase ^= 1u;
      }
    }
    if constexpr (FUSE) {
      if (f_cnt > 0u) fuse_drain();
    }
    if (p.tma_epi && wg_tid == 0) bulk_wait_all<0>();      // stores complete before the CTA retires
  }

  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();      // neither CTA frees TMEM / retires while the pair's MMAs or arrivals are in flight
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_cta2(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
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
static constexpr int kConvTcMaxSmem = 216 * 1024;

static size_t a_stage_bytes(int kchunk, const ConvParams& p) {
  return ((size_t)p.a_rows * kchunk * 2 + 1023) & ~(size_t)1023;
}
static size_t stage_bytes_of(int kchunk, const ConvParams& p) {
  const size_t b_rows = p.cta2 ? p.block_n / 2 : p.block_n;     // a CTA pair: each CTA stages half of the weight tile
  return a_stage_bytes(kchunk, p) + (p.b_resident ? 0 : (size_t)p.max_ntap * b_rows * kchunk * 2);
}
static size_t resident_bytes_of(int kchunk, const ConvParams& p) {
  return p.b_resident ? (size_t)p.num_chunks * p.block_n * kchunk * 2 : 0;
}

static size_t epilogue_bytes_of(const ConvParams& p) {
  if (!p.slab) return 0;
  const size_t nwg = p.num_epi_wg == 3 ? 3 : 2;
  if (p.tma_epi)                                                        // warpgroups x 2 swizzled buffers
    return nwg * 2 * (size_t)kBlockM * p.slab * 2 + (p.fuse_n ? (size_t)p.fuse_n * 128 : 0);   // (+ fused-tail weights)
  return nwg * (size_t)kBlockM * ((size_t)p.slab * 2 + 16);
}

int make_tmap_image_f32(CUtensorMap* out, const float* base, int N, int H, int W, int box_w_elems, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return DY_ERR_CUDA;
  }
  DY_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "image base must be 16B aligned");
  DY_CHECK(((long long)W * 3 * 4) % 16 == 0 && (box_w_elems * 4) % 16 == 0 && box_w_elems <= 256, "image row geometry");
  cuuint64_t gdim[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[2] = {(cuuint64_t)W * 3 * 4, (cuuint64_t)H * W * 3 * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_w_elems, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (image) failed with CUresult " + std::to_string((int)r));
    return DY_ERR_CUDA;
  }
  return DY_OK;
}

size_t conv_tc_smem_bytes(int kchunk, const ConvParams& p) {
  return resident_bytes_of(kchunk, p) + (size_t)p.num_stages * stage_bytes_of(kchunk, p) + epilogue_bytes_of(p) +
         1024;
}

int conv_tc_pick_stages(int kchunk, const ConvParams& p) {
  // (the three-warpgroup instantiation has 4 KB more static shared memory: per-warpgroup BN / offset tables)
  const size_t budget = 214 * 1024 - epilogue_bytes_of(p) - (p.num_epi_wg == 3 ? 4096 : 0);
  const size_t res = resident_bytes_of(kchunk, p);
  if (res + 2 * stage_bytes_of(kchunk, p) > budget) return 0;
  int s = (int)((budget - res) / stage_bytes_of(kchunk, p));
  if (s > kMaxStages) s = kMaxStages;
  return s;
}

static int g_conv_pdl = 1;
void conv_tc_set_pdl(int on) { g_conv_pdl = on; }

int launch_conv_tc(int kchunk, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                   const CUtensorMap& r, const CUtensorMap& o, const ConvParams& p, int num_sms,
                   cudaStream_t stream) {
  const bool pdl = g_conv_pdl != 0;
  DY_CHECK(kchunk == 64 || kchunk == 32, "kchunk");
  DY_CHECK(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "block_n");
  DY_CHECK(p.num_stages >= 2 && p.num_stages <= kMaxStages, "stages");
  DY_CHECK(p.num_acc == 2 || p.num_acc == 3 || p.num_acc == 4 || p.num_acc == 6, "num_acc");
  const int nwg = p.num_epi_wg == 3 ? 3 : 2;
  DY_CHECK(p.tmem_cols >= p.num_acc * p.block_n + (p.fuse_n ? nwg * 32 : 0) && p.tmem_cols <= 512, "tmem_cols");
  DY_CHECK(nwg == 2 || (!p.cta2 && !p.split_n && p.block_n <= 64), "three epilogue warpgroups: thin single-CTA tiles only");
  DY_CHECK(p.split_n || p.num_acc % nwg == 0, "every accumulator must always meet the same epilogue warpgroup");
  DY_CHECK(!p.dual_issue || p.num_acc % 2 == 0, "every accumulator must always meet the same MMA issuer");
  DY_CHECK(!p.fuse_n || (p.fuse_n == 16 && p.tma_epi && p.slab == 64 && p.block_n == 64 && p.cout == 64 &&
                         p.n_tiles_n == 1 && p.residual == nullptr && p.out[1].mode == OUT_NONE &&
                         p.fuse_cout >= 1 && p.fuse_cout <= 16 && p.fuse_w && p.fuse_bias && p.fuse_out),
           "fused tail needs the TMA staged epilogue of a 64-channel layer");
  DY_CHECK(p.a_rows == kBlockM || p.a_rows == kHaloRows, "a_rows");
  DY_CHECK(p.max_ntap >= 1 && p.max_ntap <= 3, "max_ntap");
  DY_CHECK(p.slab == 0 || ((p.slab == 32 || p.slab == 64) && p.block_n % p.slab == 0), "slab");
  DY_CHECK(!p.split_n || (p.slab != 0 && !p.dual_issue && !p.fuse_n && p.num_acc == 2 && (p.block_n / 2) % p.slab == 0),
           "split-N epilogue needs a staged epilogue, one MMA issuer and a whole number of slabs per half tile");
  const size_t smem = conv_tc_smem_bytes(kchunk, p);
  DY_CHECK(smem <= (size_t)kConvTcMaxSmem - (p.num_epi_wg == 3 ? 4096 : 0), "pipeline does not fit in shared memory");
  const int tiles = p.n_tiles_m * p.n_tiles_n;
  int grid = tiles < num_sms ? tiles : num_sms;
  if (p.b_resident) {
    // every CTA must keep seeing the same N tile: tile = blockIdx.x + i*grid, n = tile % n_tiles_n
    grid = grid / p.n_tiles_n * p.n_tiles_n;
    DY_CHECK(grid >= p.n_tiles_n, "grid too small for resident weights");
  }
  DY_CHECK(!p.fuse_n || kchunk == 32, "the fused tail is instantiated for 32-wide K chunks (convolutional81)");
  if (p.cta2) {
    DY_CHECK(kchunk == 64 && !p.b_resident && !p.dual_issue && !p.fuse_n && p.slab != 0 && p.tma_epi && p.num_acc == 2 &&
                 p.block_n % 32 == 0 && p.block_n >= 32,
             "CTA-pair plan: 64-wide K chunks, streamed weights, one issuer, TMA staged epilogue");
    static bool attr2 = false;
    if (!attr2) {
      DY_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvTcMaxSmem));
      attr2 = true;
    }
    const int pair_tiles = ((p.n_tiles_m + 1) / 2) * p.n_tiles_n;
    int g2 = 2 * pair_tiles < num_sms ? 2 * pair_tiles : (num_sms & ~1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(g2);
    cfg.blockDim = dim3(conv_threads(2));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 2 : 1;
    DY_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, false, true, 2>, a0, a1, b, r, o, p));
    DY_CUDA(cudaGetLastError());
    return DY_OK;
  }
  // instantiations: (K chunk, fused tail, CTA pair, epilogue warpgroups)
#define DY_LAUNCH_CONV(KC, FU, NW)                                                                                     \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      DY_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KC, FU, false, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                   kConvTcMaxSmem - (NW == 3 ? 4096 : 0)));                                            \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    DY_CUDA(launch_kernel_pdl(conv_tc_kernel<KC, FU, false, NW>, dim3(grid), dim3(conv_threads(NW)), smem, stream, pdl, \
                              a0, a1, b, r, o, p));                                                                    \
  } while (0)
  if (p.fuse_n) {
    if (nwg == 3) DY_LAUNCH_CONV(32, true, 3);
    else DY_LAUNCH_CONV(32, true, 2);
  } else if (kchunk == 64) {
    if (nwg == 3) DY_LAUNCH_CONV(64, false, 3);
    else DY_LAUNCH_CONV(64, false, 2);
  } else {
    if (nwg == 3) DY_LAUNCH_CONV(32, false, 3);
    else DY_LAUNCH_CONV(32, false, 2);
  }
#undef DY_LAUNCH_CONV
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

}  // namespace dy
