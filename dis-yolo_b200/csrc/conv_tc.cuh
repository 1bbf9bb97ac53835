// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a: interface.
//
// Every conv of the network (reference: yolo/yolo3_net_pos.py:109-151, tf.nn.conv2d + BN + leaky
// [+ residual]) except conv1 (Cin=3) runs through ONE kernel: a persistent, warp-specialised GEMM
// whose A operand rows are pixels of a P1-layout activation (common.cuh).  The K loop is a list of
// SEGMENTS; a segment is (A tensor map, constant row shift, first channel, #K-chunks):
//   1x1 conv            1 segment   shift 0
//   3x3 stride-1 conv   9 segments  shift (kh-1)*(W+1)+(kw-1)
//   3x3 stride-2 conv   9 segments over the space-to-depth copy of the input (TF 'SAME' pads 0
//                       before / 1 after): tap (kh,kw) -> shift (kh>>1)*(Wo+1)+(kw>>1),
//                       channel block ((kh&1)*2+(kw&1))*Cin
//   1x1 over concat     2 segments  (skip tensor, upsampled tensor)   [skip, up] order (:291)
// so A is always a plain 2D TMA tile load with out-of-bounds zero fill -- no im2col buffer.
#pragma once
#include "common.cuh"

namespace dy {

constexpr int kBlockM = 128;
constexpr int kMaxSeg = 9;
constexpr int kMaxStages = 16;

enum OutMode : int {
  OUT_NONE = 0,
  OUT_SAME = 1,         // bf16 P1, same geometry as the GEMM row space
  OUT_S2D = 2,          // bf16 P1 of the space-to-depth tensor [N, H/2+1, W/2+1, 4*C]
  OUT_UP2 = 3,          // bf16 P1 of the 2x nearest-neighbour upsampled tensor [N, 2H+1, 2W+1, C]
  OUT_F32_COMPACT = 4,  // fp32 [N,H,W,cout]          (detection heads)
  OUT_F32_PLANAR = 5,   // fp32 [N,cout,H,W]          (position-sensitive score maps)
  OUT_UNS2D = 6,        // bf16 P1 [N, 2H+1, 2W+1, cout]: GEMM row (n,y,x) -> pixel (2y + blk/2, 2x + blk%2); the
                        // inverse of OUT_S2D for ONE parity block (dgrad of a stride-2 conv, train_tc.cuh)
  OUT_UNS2D_ACC = 7     // the same, accumulating into the destination (a second consumer's gradient)
};

struct OutDesc {
  void* ptr;
  int mode;
  int ld;   // destination row pitch in elements
  int blk;  // OUT_UNS2D*: parity block (y&1)*2 + (x&1) this launch produces
};

constexpr int kHaloRows = 136;   // 128 + 2 halo rows, rounded up to the 8-row swizzle atom

// One A-operand segment of the K loop.  Every K chunk of the segment is ONE TMA box of a_rows rows
// that feeds `ntap` MMA groups: group t reads the box starting tap_row[t] rows in (the UMMA
// descriptor start address is simply advanced by whole rows -- the hardware swizzle is a pure
// function of the shared-memory address, verified on B200 by scripts/experiments/umma_shift_test.cu)
// against weight chunk tap_b0[t] + c.  A 3x3 stride-1 conv is therefore 3 segments (one per kernel
// row) of 3 taps each instead of 9 separate loads.
struct ConvSeg {
  int map;         // 0 -> mapA0, 1 -> mapA1
  int shift;       // row shift added to the tile's first row
  int col0;        // first channel (element index along the A tensor's inner dimension)
  int nchunk;      // number of K chunks in this segment
  int ntap;        // 1..3 MMA groups per chunk
  int tap_row[3];  // row offset of each group inside the loaded box
  int tap_b0[3];   // weight K-chunk index of each group for c = 0
};

struct ConvParams {
  int M;           // GEMM rows = N*(H+1)*(W+1)
  int H, W;        // valid output extent (row space is (H+1)x(W+1) per image)
  int cout;        // real output channels
  int block_n;     // N tile: multiple of 16, <= 256
  int n_tiles_m, n_tiles_n;
  int num_seg;
  ConvSeg seg[kMaxSeg];
  int num_chunks;  // number of weight K chunks = sum of seg[].nchunk * seg[].ntap
  int num_stages;  // smem pipeline depth
  int tile_stages; // pipeline stages consumed per tile = sum of seg[].nchunk
  int num_acc;     // TMEM accumulator stages: 2, or 4 (two per MMA issuer) for thin dual-issue tiles
  int dual_issue;  // 1: two MMA-issuer threads, each with its own half of the stage ring (thin layers)
  int dual_producer;  // with dual_issue: 1 = one TMA producer thread per half ring, 0 = one thread feeds both in tile order
  int a_rows;      // rows per A box: 128, or kHaloRows when taps share a halo'd box
  int max_ntap;    // max seg[].ntap (sizes the per-stage weight slots when weights are streamed)
  int b_resident;  // 1: the CTA's whole [block_n x K] weight slab is loaded to smem once
  int slab;        // epilogue staging width in columns (64 or 32); 0 = direct per-thread stores
  int debug_skip;  // measurement aid: 1 = epilogue only drains the accumulator barrier (no math, no stores)
  int cta2;        // 1: CTA-pair plan -- 2-CTA clusters, tcgen05.mma.cta_group::2 with M = 256, each CTA stages half of the
                   // weight tile (64-wide K chunks, streamed weights, TMA staged epilogue)
  int num_epi_wg;  // epilogue warpgroups: 2, or 3 for thin single-CTA tiles (<= 64 columns; 512-thread instantiation)
  int split_n;     // staged epilogue only: 1 = both epilogue warpgroups drain every tile, half of its columns each
                   // (0: they take alternate tiles)
  int tma_epi;     // staged epilogue only: out[0] (P1 layout) leaves through TMA stores, the residual
                   // arrives through TMA loads into the same swizzled staging buffers
  int tmem_cols;   // power of two >= 2*block_n, >= 32
  const float* scale;  // [n_tiles_n*block_n] folded BN scale (1 for biased convs)
  const float* shift;  // [n_tiles_n*block_n] folded BN shift / bias
  float alpha;
  int act;                          // 1 -> leaky relu max(alpha*x, x)
  const __nv_bfloat16* residual;    // P1, same geometry, added AFTER the activation (:148-151)
  int res_ld;
  OutDesc out[2];
  // Fused tail: a biased linear 1x1 conv (convolutional82, 64 -> 9, yolo3_net_pos.py:410-412) evaluated on
  // this layer's bf16 output tile while it still sits in the swizzled staging buffer: a second
  // tcgen05.mma (M=128, N=fuse_n, K=64) issued by the epilogue warpgroup itself.  Needs the TMA staged
  // epilogue with block_n == cout == slab == 64 and no residual.
  int fuse_n;                       // 0 = off, else the fused conv's padded N (16)
  int fuse_cout;                    // its real output channels
  int fuse_store;                   // 1: out[0] is still stored (parity taps / training), 0: never materialised
  const __nv_bfloat16* fuse_w;      // [fuse_n][64] bf16 K-major (= the fused layer's packed forward weights)
  const float* fuse_bias;           // [fuse_n]
  float* fuse_out;                  // fp32 planar [N, fuse_cout, H, W]
};

// Build a 2D bf16 tensor map over a row-major [rows, cols] matrix with a (box_cols x box_rows) box.
// box_cols*2 bytes must be 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B).
int make_tmap_2d(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld_elems,
                 int box_cols, int box_rows);

// 3-D fp32 tensor map over an NHWC RGB image batch viewed as [N][H][W*3]; out-of-bounds reads give 0
int make_tmap_image_f32(CUtensorMap* out, const float* base, int N, int H, int W, int box_w_elems, int box_rows);

size_t conv_tc_smem_bytes(int kchunk, const ConvParams& p);
int conv_tc_pick_stages(int kchunk, const ConvParams& p);

// 1 (default): conv launches carry the programmatic-stream-serialization attribute (their prologue overlaps the
// previous layer's tail); 0: ordinary stream order (A/B)
void conv_tc_set_pdl(int on);

// kchunk in {32, 64}
int launch_conv_tc(int kchunk, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b,
                   const CUtensorMap& r, const CUtensorMap& o, const ConvParams& p, int num_sms,
                   cudaStream_t stream);

}  // namespace dy
