// Experiment: can a K-major swizzled A operand be read starting `dx` rows into a TMA-loaded tile by
// moving the UMMA descriptor start address (and setting base_offset)?  D[128x16] = A[dx..dx+128, K] * B^T
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../../dis-yolo_b200/csrc/common.cuh"
#include "../../dis-yolo_b200/csrc/conv_tc.cuh"
namespace dy { void set_error(const std::string& m) { fprintf(stderr, "err: %s\n", m.c_str()); } }
using namespace dy;

template <int KC>
__global__ void k(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* out,
                  int dx, int mode) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t bar, done;
  __shared__ uint32_t tbase;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* gen = smem_dyn + (base - smem_u32(smem_dyn));
  constexpr uint32_t ABYTES = 136 * KC * 2;
  constexpr uint32_t AB_ALIGNED = (ABYTES + 1023) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&tbase, 32); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, ABYTES + 16 * KC * 2);
    tma_load_2d(gen, &mapA, &bar, 0, 0);
    tma_load_2d(gen + AB_ALIGNED, &mapB, &bar, 0, 0);
    mbar_wait(&bar, 0);
    tc_fence_after();
    constexpr uint32_t layout = KC == 64 ? 2u : 4u, sbo = KC == 64 ? 1024u : 512u, rowb = KC * 2;
    const uint32_t astart = base + dx * rowb;
    uint64_t ad = umma_desc(astart, sbo, layout);
    if (mode == 1) ad |= (uint64_t)((astart >> 7) & 7) << 49;
    if (mode == 2) ad |= (uint64_t)(dx & 7) << 49;
    const uint64_t bd = umma_desc(base + AB_ALIGNED, sbo, layout);
    for (int kk = 0; kk < KC / 16; ++kk) umma_bf16(tbase, ad + 2 * kk, bd + 2 * kk, umma_idesc_bf16(16), kk != 0);
    umma_commit(&done);
  }
  __syncwarp();
  if (warp < 4) {
    mbar_wait(&done, 0);
    tc_fence_after();
    uint32_t r[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tbase, 32); }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int KC> int run() {
  const int R = 160;
  std::vector<__nv_bfloat16> A(R * KC), B(16 * KC);
  std::vector<float> Af(R * KC), Bf(16 * KC);
  srand(1);
  for (int i = 0; i < R * KC; ++i) { Af[i] = bf((rand() % 200 - 100) / 50.f); A[i] = __float2bfloat16(Af[i]); }
  for (int i = 0; i < 16 * KC; ++i) { Bf[i] = bf((rand() % 200 - 100) / 50.f); B[i] = __float2bfloat16(Bf[i]); }
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 16 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  if (make_tmap_2d(&mA, dA, R, KC, KC, KC, 136) || make_tmap_2d(&mB, dB, 16, KC, KC, KC, 16)) return 1;
  cudaFuncSetAttribute(k<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 3; ++mode)
    for (int dx = 0; dx < 9; ++dx) {
      cudaMemset(dO, 0, 128 * 16 * 4);
      k<KC><<<1, 128, 48 * 1024>>>(mA, mB, dO, dx, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("KC=%d mode=%d dx=%d CUDA error %s\n", KC, mode, dx, cudaGetErrorString(e)); return 2; }
      std::vector<float> O(128 * 16);
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
          double ref = 0;
          for (int kk = 0; kk < KC; ++kk) ref += (double)Af[(m + dx) * KC + kk] * Bf[n * KC + kk];
          maxerr = fmax(maxerr, fabs(ref - O[m * 16 + n]));
        }
      printf("KC=%d mode=%d dx=%d maxerr=%.4g %s\n", KC, mode, dx, maxerr, maxerr < 1e-2 ? "OK" : "BAD");
    }
  return 0;
}

int main() { int r = run<64>(); if (r) return r; return run<32>(); }
