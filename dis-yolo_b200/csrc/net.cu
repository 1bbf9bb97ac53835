// libdisyolo_b200: the network object, the layer plan and the C ABI (include/disyolo.h).
//
// Reference being replaced: class YOLONet (yolo/yolo3_net_pos.py:12-975) driven through
// tf.Session.run by train_yolo3_mask.py / calculate_test_map.py.
#include <map>
#include <string>
#include <vector>
#include <atomic>
#include <cmath>
#include <cstring>

#include "../../include/disyolo.h"
#include "common.cuh"
#include "conv_misc.cuh"
#include "conv_tc.cuh"
#include "postproc.cuh"
#include "train.cuh"
#include "train_tc.cuh"
#include "imgproc.cuh"
#include "augment.cuh"

#include <nvtx3/nvToolsExt.h>

// NVTX ranges around the phases of the hot path (host-side markers: free unless a profiler is attached)
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

namespace dy {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<long long> g_launches{0};
static inline void note_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// the 82-layer topology (yolo3_net_pos.py:159-412).  src 0 = the input image.
// ---------------------------------------------------------------------------------------------
struct LayerDef {
  int id, src0, src1, res;   // src1: layer whose output is 2x-upsampled and concatenated AFTER src0
  int cin0, cin1, cout, k, s;
  bool bn;
  int H;                     // output spatial size (square)
};

static std::vector<LayerDef> build_defs(int S) {
  std::vector<LayerDef> d(83);
  auto add = [&](int id, int src0, int cin0, int cout, int k, int s, int H, int res = 0, int src1 = 0, int cin1 = 0,
                 bool bn = true) {
    d[id] = LayerDef{id, src0, src1, res, cin0, cin1, cout, k, s, bn, H};
  };
  add(1, 0, 3, 32, 3, 1, S);                                  // :159-161
  add(2, 1, 32, 64, 3, 2, S / 2);                             // :165-167
  add(3, 2, 64, 32, 1, 1, S / 2);                             // :169-172
  add(4, 3, 32, 64, 3, 1, S / 2, 2);                          // :174-176 (+shortcut)
  add(5, 4, 64, 128, 3, 2, S / 4);                            // :180-182
  add(6, 5, 128, 64, 1, 1, S / 4);
  add(7, 6, 64, 128, 3, 1, S / 4, 5);
  add(8, 7, 128, 64, 1, 1, S / 4);
  add(9, 8, 64, 128, 3, 1, S / 4, 7);                         // skip3 :202
  add(10, 9, 128, 256, 3, 2, S / 8);                          // :204-206
  for (int i = 0; i < 8; ++i) {                               // :208-218
    add(11 + 2 * i, 10 + 2 * i, 256, 128, 1, 1, S / 8);
    add(12 + 2 * i, 11 + 2 * i, 128, 256, 3, 1, S / 8, 10 + 2 * i);
  }
  add(27, 26, 256, 512, 3, 2, S / 16);                        // :222-224
  for (int i = 0; i < 8; ++i) {                               // :226-236
    add(28 + 2 * i, 27 + 2 * i, 512, 256, 1, 1, S / 16);
    add(29 + 2 * i, 28 + 2 * i, 256, 512, 3, 1, S / 16, 27 + 2 * i);
  }
  add(44, 43, 512, 1024, 3, 2, S / 32);                       // :240-242
  for (int i = 0; i < 4; ++i) {                               // :244-254
    add(45 + 2 * i, 44 + 2 * i, 1024, 512, 1, 1, S / 32);
    add(46 + 2 * i, 45 + 2 * i, 512, 1024, 3, 1, S / 32, 44 + 2 * i);
  }
  add(53, 52, 1024, 512, 1, 1, S / 32);                       // :258-272
  add(54, 53, 512, 1024, 3, 1, S / 32);
  add(55, 54, 1024, 512, 1, 1, S / 32);
  add(56, 55, 512, 1024, 3, 1, S / 32);
  add(57, 56, 1024, 512, 1, 1, S / 32);
  add(58, 57, 512, 1024, 3, 1, S / 32);                       // :274-276
  add(59, 58, 1024, 24, 1, 1, S / 32, 0, 0, 0, false);        // :277-279 biased, linear
  add(60, 57, 512, 256, 1, 1, S / 32);                        // :285-287
  add(61, 43, 512, 256, 1, 1, S / 16, 0, 60, 256);            // concat[skip5, up] :291-295
  add(62, 61, 256, 512, 3, 1, S / 16);
  add(63, 62, 512, 256, 1, 1, S / 16);
  add(64, 63, 256, 512, 3, 1, S / 16);
  add(65, 64, 512, 256, 1, 1, S / 16);
  add(66, 65, 256, 512, 3, 1, S / 16);                        // :309-311
  add(67, 66, 512, 24, 1, 1, S / 16, 0, 0, 0, false);         // :312-314
  add(68, 65, 256, 128, 1, 1, S / 16);                        // :320-322
  add(69, 26, 256, 128, 1, 1, S / 8, 0, 68, 128);             // concat[skip4, up] :326-330
  add(70, 69, 128, 256, 3, 1, S / 8);
  add(71, 70, 256, 128, 1, 1, S / 8);
  add(72, 71, 128, 256, 3, 1, S / 8);
  add(73, 72, 256, 128, 1, 1, S / 8);
  add(74, 73, 128, 256, 3, 1, S / 8);                         // :344-346
  add(75, 74, 256, 24, 1, 1, S / 8, 0, 0, 0, false);          // :347-349
  add(76, 73, 128, 64, 1, 1, S / 8);                          // :381-383
  add(77, 9, 128, 64, 1, 1, S / 4, 0, 76, 64);                // concat[skip3, up] :387-391
  add(78, 77, 64, 128, 3, 1, S / 4);
  add(79, 78, 128, 32, 1, 1, S / 4);                          // :396-398
  add(80, 4, 64, 32, 1, 1, S / 2, 0, 79, 32);                 // concat[skip2, up] :402-406
  add(81, 80, 32, 64, 3, 1, S / 2);
  add(82, 81, 64, 9, 1, 1, S / 2, 0, 0, 0, false);            // :410-412
  return d;
}

static void tf_same_pad(int in, int k, int s, int* before) {
  const int out = (in + s - 1) / s;
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  *before = total / 2;
}

// ---------------------------------------------------------------------------------------------
// one tensor-core conv: descriptors + launch parameters
// ---------------------------------------------------------------------------------------------
struct TcPlan {
  int kchunk = 64;
  CUtensorMap a0, a1, b, r, o;
  ConvParams p;
};

struct TcConvDesc {
  const __nv_bfloat16* a0 = nullptr;   // src0 operand: P1 SAME (s=1) or S2D (s=2) tensor
  const __nv_bfloat16* a1 = nullptr;   // UP2 tensor of src1 or null
  int cin0 = 0, cin1 = 0, cout = 0, k = 1, s = 1;
  int H = 0, W = 0;                    // output extent
  int max_batch = 1;
  const __nv_bfloat16* wpk = nullptr;  // [cout_pad][K] packed weights
  int cout_pad = 0;
  const float* scale = nullptr;
  const float* shift = nullptr;
  int act = 0;
  float alpha = 0.1f;
  const __nv_bfloat16* residual = nullptr;
  OutDesc out[2];
  // explicit tap list (k must be 1): K = ntaps_custom * cin0, tap t reads the A rows shifted by tap_shift_custom[t]
  // against weight K-chunks [t*cin0, (t+1)*cin0) -- the per-parity-block dgrad of a stride-2 conv
  int ntaps_custom = 0;
  int tap_shift_custom[4] = {0, 0, 0, 0};
  // fused 1x1 linear tail (ConvParams::fuse_*)
  int fuse_n = 0, fuse_cout = 0, fuse_store = 0;
  const __nv_bfloat16* fuse_w = nullptr;
  const float* fuse_bias = nullptr;
  float* fuse_out = nullptr;
};

// tuning overrides (dy_set_option): -1 = automatic
static int g_opt_resident = -1;     // 0: never keep weights resident, 1: whenever the slab fits
static int g_opt_halo = -1;         // 0: never share a halo'd A box between taps, 1: whenever legal
static int g_opt_staged = -1;       // 0: per-thread row stores, 1: smem-staged cooperative stores
static int g_opt_tma_epi = -1;      // 0: cooperative staged epilogue, 1: TMA store / TMA residual staged epilogue
static int g_opt_dual = -1;         // 0: one MMA issuer, 1: two issuers whenever the ring has >= 4 stages
static int g_opt_skip_epi = 0;      // measurement only: conv epilogues do nothing
static int g_opt_conv1_tc = -1;     // 0: conv1 on CUDA cores, otherwise the tcgen05 im2col stem kernel
static int g_opt_dual_producer = -1; // with dual issue: 0 one TMA producer thread for both half rings, 1 one per half ring
static int g_opt_max_stages = -1;   // cap of the shared-memory pipeline depth (default kMaxStages)
static int g_opt_split_n = -1;      // 0: epilogue warpgroups take alternate tiles, 1: both drain every tile (half the columns each) where legal
static int g_opt_kchunk = -1;       // 32: 32-wide K chunks (SWIZZLE_64B) even where 64 divides the channel counts (A/B)
static int g_opt_slab = -1;         // 32: 32-column staging slabs in the TMA epilogue (A/B)
static int g_opt_cta2 = -1;         // 0: never plan CTA pairs (cta_group::2), 1: wherever legal, -1: planner's choice
static int g_opt_epi_wg = -1;       // 2 / 3 epilogue warpgroups (3: thin single-CTA tiles only); -1: planner's choice
static int g_opt_fuse_tail = -1;    // 0: convolutional82 as its own launch, otherwise fused into convolutional81's epilogue
static int g_wg_dbg[4] = {0, 0, 0, 0};   // bring-up aid: wgrad UMMA descriptor overrides (0 = computed)

static int pick_block_n(int cout_pad, long long rows, int num_sms) {
  static const int cands[] = {256, 128, 64, 32, 16};
  int best = 16;
  for (int bn : cands) {
    if (cout_pad % bn != 0) continue;
    best = bn;
    const long long tiles = ((rows + kBlockM - 1) / kBlockM) * (cout_pad / bn);
    if (tiles >= num_sms || bn <= 64) break;     // enough CTAs, or tiles would get too thin
    // otherwise try the next smaller tile to spread over more SMs
  }
  return best;
}

static const size_t kResidentSlabMax = 80 * 1024;

static int build_tc_plan(const TcConvDesc& d, int num_sms, TcPlan* plan) {
  DY_CHECK(d.k == 1 || d.k == 3, "kernel size must be 1 or 3");
  DY_CHECK(d.s == 1 || (d.s == 2 && d.k == 3 && d.cin1 == 0), "stride 2 only for 3x3 without concat");
  DY_CHECK(d.cin0 % 32 == 0 && d.cin1 % 32 == 0, "input channels must be multiples of 32");
  DY_CHECK(d.cin1 == 0 || d.k == 1, "concat only feeds 1x1 convs");
  const int kchunk = (d.cin0 % 64 == 0 && d.cin1 % 64 == 0 && g_opt_kchunk != 32) ? 64 : 32;
  const int Hp = d.H + 1, Wp = d.W + 1;
  const long long rows_max = (long long)d.max_batch * Hp * Wp;
  const long long m_tiles = (rows_max + kBlockM - 1) / kBlockM;
  DY_CHECK(d.ntaps_custom == 0 || (d.k == 1 && d.s == 1 && d.cin1 == 0 && d.ntaps_custom <= 4), "custom taps");
  const int K = d.ntaps_custom ? d.ntaps_custom * d.cin0 : d.k * d.k * d.cin0 + d.cin1;
  ConvParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  plan->kchunk = kchunk;
  p.H = d.H;
  p.W = d.W;
  p.cout = d.cout;

  // ---- planning mode.  Rules distilled from a per-layer A/B of every mode at batch 64 on B200
  // (scripts/ab_layers.py, profiles/r1_ab_layers.txt):
  //   resident : the CTA's whole weight slab lives in smem (only when all of N fits: <= 80 KB)
  //   halo     : 3x3 -- one 136-row activation box per kernel row feeds the 3 horizontal taps
  //   staged   : epilogue goes through smem for coalesced residual reads / output stores
  bool resident = false, halo = false, staged = false, tma = false;
  int block_n = pick_block_n(d.cout_pad, rows_max, num_sms);
  const bool enough_work = m_tiles >= num_sms;
  const bool fits = d.cout_pad <= 256 && (size_t)d.cout_pad * K * 2 <= kResidentSlabMax;
  const bool has_res = d.residual != nullptr;
  const bool up2 = d.out[0].mode == OUT_UP2 || d.out[1].mode == OUT_UP2;
  if (enough_work) {
    if (d.k == 1) {
      resident = fits;
      if (d.cin1 == 0) staged = tma = d.cout_pad >= 64;        // bottleneck 1x1s: TMA-store epilogue
      else staged = tma = d.cout_pad == 64;                    // 1x1 over concat: only the thin one gains
      if (up2) staged = true;
    } else if (d.s == 1) {
      if (fits) {
        resident = halo = staged = tma = true;
      } else if (d.cout_pad <= 128) {
        halo = staged = tma = true;
      } else if (d.cout_pad == 256) {
        staged = tma = true;                                    // N=256 tile, 3 stages + 64 KB staging
      } else {
        staged = true;                                          // MMA bound: keep 4 stages, small staging
        tma = has_res && d.cout_pad == 512;
      }
    } else {
      if (fits) resident = halo = true;
      else if (d.cout_pad <= 128) halo = staged = tma = true;
    }
  } else {
    staged = up2;
  }
  // CTA pairs (cta_group::2, M = 256): the wide 3x3 layers, whose 128 x 256 single-CTA tiles are bound by the bytes
  // that cross L2 -> shared memory (48 KB per 512 tensor-pipe cycles); a pair halves the weight part of that
  const bool cta2_ok = kchunk == 64 && d.ntaps_custom == 0 && d.out[0].mode == OUT_SAME && d.cout_pad % 128 == 0 &&
                       !d.fuse_n && m_tiles >= 2;
  // (measured per layer, scripts/ab_opts.py: every 3x3 and plain 1x1 layer with >= 128 output channels gains 4-18 %;
  // the 1x1s over a concat and the ones that also write an upsampled copy do not)
  bool cta2 = cta2_ok && enough_work && d.cin1 == 0 && !up2 && d.cout_pad >= 128;
  if (g_opt_cta2 == 0) cta2 = false;
  if (g_opt_cta2 == 1) cta2 = cta2_ok;
  if (cta2) {
    resident = false;
    staged = tma = true;
    block_n = d.cout_pad % 256 == 0 ? 256 : 128;
    if (d.k == 3 && d.s == 1) halo = true;     // one 136-row box feeds the three horizontal taps (measured: -4..-8 %)
  }
  if (g_opt_resident == 0) resident = false;
  if (g_opt_resident == 1 && fits && !cta2) resident = true;
  if (g_opt_halo == 0) halo = false;
  if (g_opt_halo == 1 && d.k == 3) halo = true;
  if (g_opt_staged >= 0) staged = g_opt_staged != 0;
  if (g_opt_tma_epi >= 0) tma = g_opt_tma_epi != 0;
  if (resident) block_n = d.cout_pad;
  if (halo && !resident) {
    // three streamed weight chunks per stage: keep a stage <= ~64 KB so that >= 3 stages fit
    while (block_n > 64 && 3 * (size_t)(cta2 ? block_n / 2 : block_n) * kchunk * 2 > 48 * 1024) block_n /= 2;
  }
  p.block_n = block_n;
  p.n_tiles_n = d.cout_pad / p.block_n;
  p.b_resident = resident ? 1 : 0;
  p.cta2 = cta2 ? 1 : 0;
  p.a_rows = halo ? kHaloRows : kBlockM;
  int tc = 32;
  while (tc < 2 * p.block_n) tc <<= 1;
  p.tmem_cols = tc;
  p.scale = d.scale;
  p.shift = d.shift;
  p.alpha = d.alpha;
  p.act = d.act;
  p.residual = d.residual;
  p.res_ld = d.cout;
  p.out[0] = d.out[0];
  p.out[1] = d.out[1];

  // ---- K loop segments.  Weight K-chunk index of tap (kh,kw), chunk c: ((kh*3+kw)*cpt + c).
  const int cpt = d.cin0 / kchunk;
  int ns = 0;
  auto seg1 = [&](int map, int shift, int col0, int nchunk, int b0) {
    ConvSeg g;
    memset(&g, 0, sizeof(g));
    g.map = map; g.shift = shift; g.col0 = col0; g.nchunk = nchunk; g.ntap = 1;
    g.tap_row[0] = 0; g.tap_b0[0] = b0;
    p.seg[ns++] = g;
  };
  if (d.ntaps_custom) {
    for (int t = 0; t < d.ntaps_custom; ++t) seg1(0, d.tap_shift_custom[t], 0, cpt, t * cpt);
  } else if (d.k == 1) {
    seg1(0, 0, 0, cpt, 0);
    if (d.cin1 > 0) seg1(1, 0, 0, d.cin1 / kchunk, cpt);
  } else if (d.s == 1) {
    if (halo) {
      for (int kh = 0; kh < 3; ++kh) {      // one box per kernel row, three row-shifted taps
        ConvSeg g;
        memset(&g, 0, sizeof(g));
        g.map = 0; g.shift = (kh - 1) * Wp - 1; g.col0 = 0; g.nchunk = cpt; g.ntap = 3;
        for (int kw = 0; kw < 3; ++kw) { g.tap_row[kw] = kw; g.tap_b0[kw] = (kh * 3 + kw) * cpt; }
        p.seg[ns++] = g;
      }
    } else {
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) seg1(0, (kh - 1) * Wp + (kw - 1), 0, cpt, (kh * 3 + kw) * cpt);
    }
  } else {
    // TF 'SAME', stride 2, even input: pad 0 before / 1 after -> input pixel (2oy+kh, 2ox+kw) =
    // space-to-depth pixel (oy+(kh>>1), ox+(kw>>1)), channel block ((kh&1)*2 + (kw&1))
    if (halo) {
      for (int kh = 0; kh < 3; ++kh) {
        ConvSeg g;                           // column parity 0: taps kw=0 (row 0) and kw=2 (row 1)
        memset(&g, 0, sizeof(g));
        g.map = 0; g.shift = (kh >> 1) * Wp; g.col0 = ((kh & 1) << 1) * d.cin0; g.nchunk = cpt; g.ntap = 2;
        g.tap_row[0] = 0; g.tap_b0[0] = (kh * 3 + 0) * cpt;
        g.tap_row[1] = 1; g.tap_b0[1] = (kh * 3 + 2) * cpt;
        p.seg[ns++] = g;
        seg1(0, (kh >> 1) * Wp, (((kh & 1) << 1) | 1) * d.cin0, cpt, (kh * 3 + 1) * cpt);   // parity 1: kw=1
      }
    } else {
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          seg1(0, (kh >> 1) * Wp + (kw >> 1), (((kh & 1) << 1) | (kw & 1)) * d.cin0, cpt, (kh * 3 + kw) * cpt);
    }
  }
  int nchunks = 0, max_ntap = 1;
  p.tile_stages = 0;
  for (int i = 0; i < ns; ++i) {
    p.tile_stages += p.seg[i].nchunk;
    nchunks += p.seg[i].nchunk * p.seg[i].ntap;
    if (p.seg[i].ntap > max_ntap) max_ntap = p.seg[i].ntap;
  }
  p.num_seg = ns;
  p.num_chunks = nchunks;
  p.max_ntap = max_ntap;
  // epilogue staging: bf16 outputs go through shared memory for coalesced stores
  const bool bf16_out = d.out[0].mode == OUT_SAME || d.out[0].mode == OUT_S2D || d.out[0].mode == OUT_UP2;
  if (!bf16_out || p.block_n % 32 != 0) staged = false;
  p.slab = !staged ? 0 : (p.block_n % 64 == 0 && p.block_n <= 128 ? 64 : 32);
  p.tma_epi = (staged && tma && d.out[0].mode == OUT_SAME) ? 1 : 0;
  p.debug_skip = g_opt_skip_epi > 0 ? g_opt_skip_epi : 0;
  if (p.tma_epi && p.block_n % 64 == 0) p.slab = 64;
  if (p.tma_epi && g_opt_slab == 32 && !d.fuse_n) p.slab = 32;
  // CTA-pair plans without a shared halo box (the stride-2 convs: nine 32 KB stages per K chunk group): the smaller
  // staging area buys a fifth pipeline stage (measured -6..-9 %)
  if (p.tma_epi && cta2 && !halo && d.k == 3 && g_opt_slab < 0) p.slab = 32;
  DY_CHECK(nchunks * kchunk == K, "K chunking mismatch");
  p.fuse_n = d.fuse_n;      // (sizes the epilogue staging area)
  // thin tiles: a third epilogue warpgroup (the 512-thread instantiation); decided before the stage count because
  // it adds a third pair of staging buffers (tiles of <= 64 columns never use the column-split epilogue)
  const bool wg3_ok = !cta2 && p.block_n <= 64;
  bool wg3 = false;      // (measured, scripts/ab_opts.py tc_epi_wg=3: only conv3 / 6 / 8 gain 5 %, the 3x3 thin layers lose 5-15 %)
  if (g_opt_epi_wg == 2) wg3 = false;
  if (g_opt_epi_wg == 3) wg3 = wg3_ok;
  p.num_epi_wg = wg3 ? 3 : 2;
  p.num_stages = conv_tc_pick_stages(kchunk, p);
  // deeper rings than 8 stages measured slower on every layer (profiles/r2_ab_opts.txt): default cap 8
  const int stage_cap = g_opt_max_stages >= 2 ? g_opt_max_stages : 8;
  if (p.num_stages > stage_cap) p.num_stages = stage_cap;
  DY_CHECK(p.num_stages >= 2, "no room for a shared-memory pipeline");
  // two MMA issuers for thin tiles (32/64-cycle MMAs: the issue thread, not the tensor pipe, is the
  // bottleneck); each issuer needs its own half ring with at least one tile's worth of stages in flight
  bool dual_issue = enough_work && p.block_n <= 64 && p.num_stages >= 6;
  if (g_opt_dual == 0) dual_issue = false;
  if (g_opt_dual == 1) dual_issue = p.num_stages >= 4;
  if (cta2) dual_issue = false;
  if (dual_issue) p.num_stages &= ~1;
  p.dual_issue = dual_issue ? 1 : 0;
  p.dual_producer = (dual_issue && g_opt_dual_producer != 0) ? 1 : 0;
  // four TMEM accumulator stages (two per issuer) when they fit: otherwise each issuer would be
  // serialised with its own epilogue warpgroup
  p.num_acc = (dual_issue && 4 * p.block_n <= 512) ? 4 : 2;
  if (p.num_epi_wg == 3) p.num_acc = dual_issue ? 6 : 3;      // (block_n <= 64: at most 384 columns)
  // wide tiles (two TMEM stages only): both epilogue warpgroups drain every tile, half of its columns each, so the
  // accumulator returns to the MMA issuer after half the tcgen05.ld / residual round trips
  const bool split_ok = !dual_issue && p.slab != 0 && !d.fuse_n && p.num_acc == 2 && (p.block_n / 2) % p.slab == 0;
  bool split_n = split_ok && enough_work && p.block_n == 256;     // (measured: 128-column tiles lose 4 %)
  if (g_opt_split_n == 0) split_n = false;
  if (g_opt_split_n == 1) split_n = split_ok;
  p.split_n = split_n ? 1 : 0;
  if (d.fuse_n) {
    DY_CHECK(p.tma_epi && p.slab == 64 && p.block_n == 64 && d.cout == 64 && !has_res && d.out[1].mode == OUT_NONE,
             "fused tail: the producer must be a 64-channel layer on the TMA staged epilogue");
    p.fuse_n = d.fuse_n; p.fuse_cout = d.fuse_cout; p.fuse_store = d.fuse_store;
    p.fuse_w = d.fuse_w; p.fuse_bias = d.fuse_bias; p.fuse_out = d.fuse_out;
  }
  tc = 32;
  while (tc < p.num_acc * p.block_n + (p.fuse_n ? p.num_epi_wg * 32 : 0)) tc <<= 1;
  DY_CHECK(tc <= 512, "TMEM columns");
  p.tmem_cols = tc;
  const int a0_cols = (d.s == 2) ? 4 * d.cin0 : d.cin0;
  DY_TRY(make_tmap_2d(&plan->a0, d.a0, rows_max, a0_cols, a0_cols, kchunk, p.a_rows));
  if (d.cin1 > 0) {
    DY_TRY(make_tmap_2d(&plan->a1, d.a1, rows_max, d.cin1, d.cin1, kchunk, p.a_rows));
  } else {
    plan->a1 = plan->a0;
  }
  DY_TRY(make_tmap_2d(&plan->b, d.wpk, d.cout_pad, K, K, kchunk, cta2 ? p.block_n / 2 : p.block_n));
  const int rbox = p.slab ? p.slab : 64;
  if (d.residual != nullptr) {
    DY_CHECK(d.cout % 64 == 0 && p.block_n % 64 == 0, "residual layers need cout % 64 == 0");
    DY_TRY(make_tmap_2d(&plan->r, d.residual, rows_max, d.cout, d.cout, rbox, kBlockM));  // L2 prefetch / TMA load
  } else {
    plan->r = plan->a0;
  }
  if (p.tma_epi) {
    DY_TRY(make_tmap_2d(&plan->o, d.out[0].ptr, rows_max, d.cout, d.out[0].ld, p.slab, kBlockM));
  } else {
    plan->o = plan->a0;
  }
  return DY_OK;
}

static int run_tc_plan(TcPlan& plan, int B, int num_sms, cudaStream_t st) {
  ConvParams& p = plan.p;
  const long long M = (long long)B * (p.H + 1) * (p.W + 1);
  DY_CHECK(M < (1ll << 31) - 256, "too many rows");
  p.M = (int)M;
  p.n_tiles_m = (int)((M + kBlockM - 1) / kBlockM);
  note_launch();
  return launch_conv_tc(plan.kchunk, plan.a0, plan.a1, plan.b, plan.r, plan.o, p, num_sms, st);
}

// pack HWIO fp32 weights -> [cout_pad][K] bf16 (K index = (kh*k+kw)*cin + ci, zero rows beyond cout)
static void pack_weights_bf16(const float* w_hwio, int K, int cout, int cout_pad, std::vector<__nv_bfloat16>* out) {
  out->assign((size_t)cout_pad * K, __float2bfloat16(0.f));
  for (int kk = 0; kk < K; ++kk)
    for (int co = 0; co < cout; ++co) (*out)[(size_t)co * K + kk] = __float2bfloat16(w_hwio[(size_t)kk * cout + co]);
}

// ---------------------------------------------------------------------------------------------
// the net
// ---------------------------------------------------------------------------------------------
struct LayerState {
  LayerDef def;
  std::vector<float> w, gamma, beta, mean, var, bias;   // host copies as loaded
  unsigned host_new = 0;     // bit v set: variable v (0 w, 1 gamma, 2 beta, 3 mean, 4 var, 5 bias) was loaded
                             // by dy_load_weights since the last dy_finalize_weights
  int cout_pad = 0, K = 0;
  float* d_w_f32 = nullptr;
  __nv_bfloat16* d_wpk = nullptr;
  float* d_scale = nullptr;
  float* d_shift = nullptr;
  // outputs
  bool need_same = false, need_s2d = false, need_up = false;
  __nv_bfloat16* same = nullptr;
  __nv_bfloat16* s2d = nullptr;
  __nv_bfloat16* up = nullptr;
  float* f32 = nullptr;     // bf16 mode: heads (compact) / score maps (planar); fp32 mode: every layer, NHWC
  TcPlan plan;
  TcPlan plan_fused;         // convolutional81 with convolutional82 evaluated in its epilogue (inference)
  bool has_fused = false;
  bool planned = false;
  // ---- training state (fp32 engine) ----
  bool unlocked = false, in_bwd = false, need_in_grad = false;
  float *d_gamma = nullptr, *d_beta = nullptr, *d_mean = nullptr, *d_var = nullptr, *d_bias = nullptr;
  float* z = nullptr;          // pre-BN conv output of this step
  float* dy = nullptr;         // gradient w.r.t. this layer's output
  float* wt = nullptr;         // [k*k][cout][cin] weights for dgrad
  float *bn_a = nullptr, *bn_b = nullptr, *bmean = nullptr, *bvar = nullptr, *binvstd = nullptr;
  double* stat = nullptr;      // [4*cout]: sum, sumsq | s1, s2
  long long off_w = -1, off_g = -1, off_b = -1;   // offsets into the flat trainable vector
  // ---- training state (bf16 tensor-core engine) ----
  int Cg = 0;                       // channels of dyb: cout rounded up to 32
  __nv_bfloat16* zb = nullptr;      // pre-BN conv output of this step, P1
  __nv_bfloat16* dyb = nullptr;     // gradient w.r.t. this layer's output, P1, Cg channels
  float* dyf = nullptr;             // biased linear convs: fp32 NHWC gradient written by the loss kernels
  __nv_bfloat16* wdg0 = nullptr;    // dgrad operand towards src0: [cin0][k*k*Cg], taps rotated by 180 degrees
  __nv_bfloat16* wdg1 = nullptr;    // dgrad operand towards src1 (concat branch, 1x1): [cin1][Cg]
  TcPlan plan_z, plan_dg0, plan_dg1;
  TcPlan plan_dg_s2[4];             // stride-2 conv: dgrad of the four parity blocks of the space-to-depth input
  WgradPlan plan_wg;
  bool dg0 = false, dg1 = false;    // which input gradients this layer produces
  bool dg_s2 = false;
  bool res_acc = false;             // shortcut gradient: accumulate (true) or first write (copy)
};

}  // namespace dy

using namespace dy;

struct dy_net {
  dy_config cfg;
  int S = 0;                       // image size
  int num_sms = 148;
  std::vector<LayerState> L;       // 1..82
  bool finalized = false;
  std::vector<void*> allocs;
  // post-processing workspace
  int n0 = 0, cap = 0;
  Cand* cand = nullptr;
  int* cand_count = nullptr;
  int* sel = nullptr;
  int* sel_cnt = nullptr;
  float4* nms_box = nullptr;            // NMS overflow scratch [B,num_classes,cap] (postproc.cuh NmsArgs)
  unsigned long long* nms_key = nullptr;
  int* nms_pos = nullptr;
  int* edges = nullptr;
  int* raw_count = nullptr;
  float* det_raw_ws = nullptr;
  float* det_box_ws = nullptr;
  int* det_count_ws = nullptr;
  // dy_forward_host*: two device-side slots so that the copies of one step overlap the compute of
  // the next (H2D, compute and D2H each on their own stream)
  static constexpr int kHostSlots = 3;
  struct HostSlot {
    float* images = nullptr;
    uint8_t* images_u8 = nullptr;    // dy_forward_host_begin_u8 staging
    float* windows = nullptr;
    void* small = nullptr;           // [det_count | det_box | det_raw | crop offsets] in one block
    uint8_t* small_host = nullptr;   // pinned mirror of `small`
    float* det_raw = nullptr;
    float* det_box = nullptr;
    int* det_count = nullptr;
    long long* crop_off = nullptr;
    float* masks = nullptr;          // [B,max_det,S,S] maps, or the packed box crops
    cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr;
    int B = 0;
    bool busy = false;
    int mask_mode = 0;
  } slot[kHostSlots];
  int next_slot = 0;
  cudaStream_t h2d_stream = nullptr, comp_stream = nullptr, d2h_stream = nullptr;
  // The host-buffer pipeline runs the network on comp_stream, every other entry point on the caller's
  // stream, and both use the same activation / weight buffers: ev_user is recorded after the last
  // caller-stream work (comp_stream waits for it), ev_host after the last comp_stream work (caller
  // streams wait for it).
  cudaEvent_t ev_user = nullptr, ev_host = nullptr;
  bool user_work = false, host_work = false;
  // loss hyper-parameters (yolo/config.py:49-57), dy_set_loss_params
  float object_scale = 2.f, noobject_scale = 1.f, class_scale = 1.f, coord_scale = 1.f, mask_scale = 5.f;
  float ignore_thresh = 0.5f;
  bool last_fused = false;         // the last forward skipped materialising convolutional81 (fused tail)
  // ---- training ----
  bool train_ready = false;
  long long n_train = 0;           // number of trainable scalars
  float* adam_m = nullptr;
  float* adam_v = nullptr;
  float* dz_scratch = nullptr;     // largest dz
  float* dx_scratch = nullptr;     // largest concat dgrad result
  double* loss_acc = nullptr;      // [8] obj, noobj, cls, xy, wh, mask, l2, (unused)
  float* mask_rois = nullptr;
  int* mask_assign = nullptr;
  int* mask_npos = nullptr;
  float* train_windows = nullptr;
  const float* train_images = nullptr;   // images of the current training step (convolutional1's weight gradient)
  float* ones_dev = nullptr;       // [1024] identity scale for "conv only" passes
  float* zeros_dev = nullptr;
  __nv_bfloat16* dzb_scratch = nullptr;   // bf16 engine: dz of the layer being differentiated (P1)
  __nv_bfloat16* pool_scratch = nullptr;  // bf16 engine: 2x2-pooled dz of a concat consumer (P1, low resolution)
  long long adam_step = 0;
  // device tables for the one-launch multi-tensor kernels (Adam / L2, BN moving averages + fold, repacking)
  ParamSeg* pseg_dev = nullptr;
  BnSeg* bseg_dev = nullptr;
  PackSeg* kseg_dev = nullptr;
  int n_pseg = 0, n_bseg = 0, n_kseg = 0, max_pack_tiles = 0;
};

namespace dy {

static int dev_alloc(dy_net* net, void** p, size_t bytes, bool zero = true) {
  if (bytes == 0) bytes = 16;
  DY_CUDA(cudaMalloc(p, bytes));
  net->allocs.push_back(*p);
  if (zero) DY_CUDA(cudaMemset(*p, 0, bytes));
  return DY_OK;
}

static size_t p1_elems(int B, int H, int W, int C) { return (size_t)B * (H + 1) * (W + 1) * C; }

static bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return cs != cudaStreamCaptureStatusNone;
}

// Entry points that run on a caller stream bracket their work with these two: the caller stream first
// waits for the host-buffer pipeline's last network pass (same activation buffers), and the pipeline's
// compute stream will wait for what the caller enqueued (train-then-evaluate, train_yolo3_mask.py:146-176).
static int user_begin(dy_net* net, cudaStream_t st) {
  if (net->host_work && !stream_capturing(st)) DY_CUDA(cudaStreamWaitEvent(st, net->ev_host, 0));
  return DY_OK;
}
static int user_end(dy_net* net, cudaStream_t st) {
  if (stream_capturing(st)) return DY_OK;
  if (!net->ev_user) DY_CUDA(cudaEventCreateWithFlags(&net->ev_user, cudaEventDisableTiming));
  DY_CUDA(cudaEventRecord(net->ev_user, st));
  net->user_work = true;
  return DY_OK;
}

static int allocate_buffers(dy_net* net) {
  const int B = net->cfg.max_batch;
  auto& L = net->L;
  const bool bf16 = net->cfg.precision == DY_PRECISION_BF16;
  for (int n = 1; n <= 82; ++n) {
    const LayerDef& d = L[n].def;
    L[n].cout_pad = (d.cout + 15) / 16 * 16;
    L[n].K = d.k * d.k * d.cin0 + d.cin1;
  }
  if (bf16) {
    for (int n = 1; n <= 82; ++n) {
      const LayerDef& d = L[n].def;
      if (d.src0 > 0) {
        if (d.s == 2) L[d.src0].need_s2d = true; else L[d.src0].need_same = true;
      }
      if (d.src1 > 0) L[d.src1].need_up = true;
      if (d.res > 0) L[d.res].need_same = true;
    }
    for (int n = 1; n <= 82; ++n) {
      const LayerDef& d = L[n].def;
      LayerState& s = L[n];
      if (!d.bn) {   // heads + score maps: fp32
        DY_TRY(dev_alloc(net, (void**)&s.f32, (size_t)B * d.H * d.H * d.cout * 4));
        continue;
      }
      if (s.need_same) DY_TRY(dev_alloc(net, (void**)&s.same, p1_elems(B, d.H, d.H, d.cout) * 2));
      if (s.need_s2d) DY_TRY(dev_alloc(net, (void**)&s.s2d, p1_elems(B, d.H / 2, d.H / 2, 4 * d.cout) * 2));
      if (s.need_up) DY_TRY(dev_alloc(net, (void**)&s.up, p1_elems(B, 2 * d.H, 2 * d.H, d.cout) * 2));
    }
  } else {
    for (int n = 1; n <= 82; ++n) {
      const LayerDef& d = L[n].def;
      DY_TRY(dev_alloc(net, (void**)&L[n].f32, (size_t)B * d.H * d.H * d.cout * 4));
    }
  }
  // post-processing workspace
  const int g0 = net->S / 8, g1 = net->S / 16, g2 = net->S / 32;
  net->n0 = 3 * (g0 * g0 + g1 * g1 + g2 * g2);
  net->cap = net->n0;
  const int md = net->cfg.max_detection, C = net->cfg.num_classes;
  DY_TRY(dev_alloc(net, (void**)&net->cand, (size_t)B * net->cap * sizeof(Cand)));
  DY_TRY(dev_alloc(net, (void**)&net->cand_count, (size_t)B * 4));
  DY_TRY(dev_alloc(net, (void**)&net->sel, (size_t)B * C * md * 4));
  DY_TRY(dev_alloc(net, (void**)&net->sel_cnt, (size_t)B * C * 4));
  DY_TRY(dev_alloc(net, (void**)&net->nms_box, (size_t)B * C * net->cap * 16, false));
  DY_TRY(dev_alloc(net, (void**)&net->nms_key, (size_t)B * C * net->cap * 8, false));
  DY_TRY(dev_alloc(net, (void**)&net->nms_pos, (size_t)B * C * net->cap * 4, false));
  DY_TRY(dev_alloc(net, (void**)&net->edges, (size_t)B * md * 2 * (kMaxK + 1) * 4));
  DY_TRY(dev_alloc(net, (void**)&net->raw_count, (size_t)B * 4));
  DY_TRY(dev_alloc(net, (void**)&net->det_raw_ws, (size_t)B * md * 6 * 4));
  DY_TRY(dev_alloc(net, (void**)&net->det_box_ws, (size_t)B * md * 6 * 4));
  DY_TRY(dev_alloc(net, (void**)&net->det_count_ws, (size_t)B * 4));
  return DY_OK;
}

static int fold_and_upload(dy_net* net, int n) {
  LayerState& s = net->L[n];
  const LayerDef& d = s.def;
  const size_t wn = (size_t)s.K * d.cout;
  if (s.w.size() != wn) {
    set_error("missing or mis-shaped weights for convolutional" + std::to_string(n));
    return DY_ERR_STATE;
  }
  std::vector<float> scale(s.cout_pad, 0.f), shift(s.cout_pad, 0.f);
  if (d.bn) {
    if ((int)s.gamma.size() != d.cout || (int)s.beta.size() != d.cout || (int)s.mean.size() != d.cout ||
        (int)s.var.size() != d.cout) {
      set_error("missing BatchNorm variables for convolutional" + std::to_string(n));
      return DY_ERR_STATE;
    }
    for (int c = 0; c < d.cout; ++c) {
      // inference-mode BN with moving statistics (yolo3_net_pos.py:81,101), folded
      const float inv = s.gamma[c] / sqrtf(s.var[c] + net->cfg.bn_eps);
      scale[c] = inv;
      shift[c] = s.beta[c] - s.mean[c] * inv;
    }
  } else {
    if ((int)s.bias.size() != d.cout) {
      set_error("missing biases for convolutional" + std::to_string(n));
      return DY_ERR_STATE;
    }
    for (int c = 0; c < d.cout; ++c) {
      scale[c] = 1.f;
      shift[c] = s.bias[c];
    }
  }
  if (!s.d_scale) {
    DY_TRY(dev_alloc(net, (void**)&s.d_scale, (size_t)s.cout_pad * 4));
    DY_TRY(dev_alloc(net, (void**)&s.d_shift, (size_t)s.cout_pad * 4));
  }
  DY_CUDA(cudaMemcpy(s.d_scale, scale.data(), (size_t)s.cout_pad * 4, cudaMemcpyHostToDevice));
  DY_CUDA(cudaMemcpy(s.d_shift, shift.data(), (size_t)s.cout_pad * 4, cudaMemcpyHostToDevice));
  const bool bf16 = net->cfg.precision == DY_PRECISION_BF16;
  if (!bf16 || n == 1) {
    if (!s.d_w_f32) DY_TRY(dev_alloc(net, (void**)&s.d_w_f32, wn * 4));
    DY_CUDA(cudaMemcpy(s.d_w_f32, s.w.data(), wn * 4, cudaMemcpyHostToDevice));
  }
  if (bf16 && n > 1) {
    std::vector<__nv_bfloat16> pk;
    pack_weights_bf16(s.w.data(), s.K, d.cout, s.cout_pad, &pk);
    if (!s.d_wpk) DY_TRY(dev_alloc(net, (void**)&s.d_wpk, pk.size() * 2));
    DY_CUDA(cudaMemcpy(s.d_wpk, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
  }
  return DY_OK;
}

static int plan_layer(dy_net* net, int n) {
  LayerState& s = net->L[n];
  const LayerDef& d = s.def;
  TcConvDesc c;
  c.a0 = (d.s == 2) ? net->L[d.src0].s2d : net->L[d.src0].same;
  c.a1 = d.src1 > 0 ? net->L[d.src1].up : nullptr;
  DY_CHECK(c.a0 != nullptr && (d.src1 == 0 || c.a1 != nullptr), "layer input buffer missing");
  c.cin0 = d.cin0;
  c.cin1 = d.cin1;
  c.cout = d.cout;
  c.k = d.k;
  c.s = d.s;
  c.H = d.H;
  c.W = d.H;
  c.max_batch = net->cfg.max_batch;
  c.wpk = s.d_wpk;
  c.cout_pad = s.cout_pad;
  c.scale = s.d_scale;
  c.shift = s.d_shift;
  c.act = d.bn ? 1 : 0;
  c.alpha = net->cfg.alpha;
  c.residual = d.res > 0 ? net->L[d.res].same : nullptr;
  int no = 0;
  c.out[0] = OutDesc{nullptr, OUT_NONE, 0};
  c.out[1] = OutDesc{nullptr, OUT_NONE, 0};
  if (!d.bn) {
    c.out[no++] = OutDesc{s.f32, n == 82 ? OUT_F32_PLANAR : OUT_F32_COMPACT, d.cout};
  } else {
    if (s.need_same) c.out[no++] = OutDesc{s.same, OUT_SAME, d.cout};
    if (s.need_s2d) c.out[no++] = OutDesc{s.s2d, OUT_S2D, 4 * d.cout};
    if (s.need_up) {
      DY_CHECK(no < 2, "too many output forms");
      c.out[no++] = OutDesc{s.up, OUT_UP2, d.cout};
    }
  }
  DY_CHECK(no >= 1 && no <= 2, "layer has no consumer");
  DY_TRY(build_tc_plan(c, net->num_sms, &s.plan));
  s.planned = true;
  // convolutional81 -> convolutional82 (64 -> k*k linear 1x1, :408-412): in inference nothing else reads
  // convolutional81's output, so the 1x1 runs on the finished tile inside the epilogue and the 64-channel
  // 288x288 activation is never written (dy_forward); dy_forward_network keeps the two launches.
  s.has_fused = false;
  if (n == 81 && g_opt_fuse_tail != 0 && s.plan.p.tma_epi && s.plan.p.slab == 64 && s.plan.p.block_n == 64 &&
      d.cout == 64 && no == 1 && net->L[82].def.src0 == 81 && net->L[82].def.k == 1 && net->L[82].cout_pad == 16 &&
      net->L[82].d_wpk != nullptr) {
    c.fuse_n = 16;
    c.fuse_cout = net->L[82].def.cout;
    c.fuse_store = 0;
    c.fuse_w = net->L[82].d_wpk;
    c.fuse_bias = net->L[82].d_shift;
    c.fuse_out = net->L[82].f32;
    DY_TRY(build_tc_plan(c, net->num_sms, &s.plan_fused));
    s.has_fused = true;
  }
  return DY_OK;
}

static int run_network_bf16(dy_net* net, const float* images, int B, cudaStream_t st, bool fused) {
  NvtxRange nvtx_("dy:network(82 convs)");
  auto& L = net->L;
  note_launch();
  DY_TRY(launch_conv1(images, L[1].d_w_f32, L[1].d_scale, L[1].d_shift, net->cfg.alpha, B, net->S, net->S, L[1].s2d,
                      L[1].same, g_opt_conv1_tc != 0, net->num_sms, st));
  fused = fused && L[81].has_fused;
  for (int n = 2; n <= 82; ++n) {
    if (fused && n == 81) DY_TRY(run_tc_plan(L[81].plan_fused, B, net->num_sms, st));
    else if (fused && n == 82) continue;
    else DY_TRY(run_tc_plan(L[n].plan, B, net->num_sms, st));
  }
  net->last_fused = fused;
  return DY_OK;
}

static int run_network_fp32(dy_net* net, const float* images, int B, cudaStream_t st) {
  auto& L = net->L;
  for (int n = 1; n <= 82; ++n) {
    const LayerDef& d = L[n].def;
    RefConvArgs a;
    memset(&a, 0, sizeof(a));
    a.src0 = d.src0 == 0 ? images : L[d.src0].f32;
    a.src1 = d.src1 > 0 ? L[d.src1].f32 : nullptr;
    a.w = L[n].d_w_f32;
    a.scale = L[n].d_scale;
    a.shift = L[n].d_shift;
    a.residual = d.res > 0 ? L[d.res].f32 : nullptr;
    a.out = L[n].f32;
    a.Ho = a.Wo = d.H;
    a.Hi = a.Wi = d.H * d.s;
    a.c0 = d.cin0;
    a.c1 = d.cin1;
    a.cout = d.cout;
    a.k = d.k;
    a.s = d.s;
    tf_same_pad(a.Hi, d.k, d.s, &a.pad_t);
    a.pad_l = a.pad_t;
    a.act = d.bn ? 1 : 0;
    a.alpha = net->cfg.alpha;
    note_launch();
    DY_TRY(launch_conv_ref(a, B, st));
  }
  return DY_OK;
}

static int run_network(dy_net* net, const float* images, int B, cudaStream_t st, bool fused) {
  DY_CHECK(net->finalized, "dy_finalize_weights has not been called");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch exceeds max_batch");
  if (net->cfg.precision == DY_PRECISION_BF16) return run_network_bf16(net, images, B, st, fused);
  net->last_fused = false;
  return run_network_fp32(net, images, B, st);
}

// decode -> NMS -> top-k (-> masks) on the given head / score maps
static int run_detect(dy_net* net, const float* y8, const float* y16, const float* y32, int B, const float* windows,
                      float thresh, float* dense_box, int* dense_cls, float* dense_score, float* det_raw,
                      float* det_box, int* det_count, cudaStream_t st) {
  NvtxRange nvtx_("dy:decode+nms+topk");
  DecodeArgs da;
  memset(&da, 0, sizeof(da));
  da.yolo[0] = y8; da.yolo[1] = y16; da.yolo[2] = y32;
  da.g[0] = net->S / 8; da.g[1] = net->S / 16; da.g[2] = net->S / 32;
  da.B = B;
  da.num_class = net->cfg.num_classes;
  da.net = 32 * da.g[2];
  memcpy(da.anchors, net->cfg.anchors, sizeof(da.anchors));
  da.windows = windows;
  da.thresh = thresh;
  da.dense_box = dense_box; da.dense_cls = dense_cls; da.dense_score = dense_score;
  da.cand = net->cand; da.cand_count = net->cand_count; da.cap = net->cap;
  DY_CUDA(cudaMemsetAsync(net->cand_count, 0, (size_t)B * 4, st));
  note_launch();
  DY_TRY(launch_decode(da, st));
  if (det_box == nullptr && det_raw == nullptr) return DY_OK;
  NmsArgs na;
  na.cand = net->cand; na.cand_count = net->cand_count; na.cap = net->cap;
  na.B = B; na.num_class = net->cfg.num_classes; na.max_det = net->cfg.max_detection;
  na.iou_thr = net->cfg.iou_threshold;
  na.sel = net->sel; na.sel_cnt = net->sel_cnt;
  na.ovf_box = net->nms_box; na.ovf_key = net->nms_key; na.ovf_pos = net->nms_pos;
  note_launch();
  DY_TRY(launch_nms(na, st));
  FinalizeArgs fa;
  fa.cand = net->cand; fa.cap = net->cap; fa.B = B; fa.num_class = net->cfg.num_classes;
  fa.max_det = net->cfg.max_detection; fa.sel = net->sel; fa.sel_cnt = net->sel_cnt;
  fa.S = net->S / 2; fa.k = net->cfg.k_map;
  fa.det_raw = det_raw ? det_raw : net->det_raw_ws;
  fa.raw_count = net->raw_count;
  fa.det_box = det_box ? det_box : net->det_box_ws;
  fa.det_count = det_count ? det_count : net->det_count_ws;
  fa.edges = net->edges;
  note_launch();
  DY_TRY(launch_finalize(fa, st));
  return DY_OK;
}

static MaskArgs mask_args(dy_net* net, const float* score, int layout, int B, const int* det_count, float* masks) {
  MaskArgs ma;
  const int Sm = net->S / 2, kk = net->cfg.k_map * net->cfg.k_map;
  ma.score = score;
  if (layout == 1) {   // planar [B,kk,S,S]
    ma.s_img = (long long)kk * Sm * Sm; ma.s_ch = (long long)Sm * Sm; ma.s_row = Sm; ma.s_pix = 1;
  } else {             // NHWC [B,S,S,kk]
    ma.s_img = (long long)kk * Sm * Sm; ma.s_ch = 1; ma.s_row = (long long)Sm * kk; ma.s_pix = kk;
  }
  ma.det_count = det_count; ma.edges = net->edges;
  ma.B = B; ma.max_det = net->cfg.max_detection; ma.S = Sm; ma.k = net->cfg.k_map;
  ma.out = masks;
  return ma;
}

static int run_masks(dy_net* net, const float* score, int layout, int B, const int* det_count, float* masks,
                     cudaStream_t st) {
  NvtxRange nvtx_("dy:mask_assembly");
  const MaskArgs ma = mask_args(net, score, layout, B, det_count, masks);
  note_launch();
  return launch_masks(ma, st);
}

}  // namespace dy

namespace dy {
__global__ void pack_cands_kernel(const float* box, const int* cls, const float* score, int N, float thresh,
                                  Cand* cand, int* cand_count, int cap) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float s = score[(long long)b * N + i];
  if (s > thresh) {
    const int slot = atomicAdd(cand_count + b, 1);
    if (slot < cap) {
      const float4 bx = reinterpret_cast<const float4*>(box)[(long long)b * N + i];
      Cand c;
      c.y1 = bx.x; c.x1 = bx.y; c.y2 = bx.z; c.x2 = bx.w;
      c.score = s; c.idx = i; c.cls = cls[(long long)b * N + i]; c.pad = 0;
      cand[(long long)b * cap + slot] = c;
    }
  }
}
__global__ void sel_to_idx_kernel(const Cand* cand, int cap, const float* det_raw, const int* raw_count, int max_det,
                                  const int* sel, const int* sel_cnt, int num_class, int* sel_idx, int* sel_count) {
  // recover the candidate index of every output row by matching (score, box) is fragile; instead
  // re-rank here exactly like finalize_kernel: rows are ordered by (score desc, idx asc).
  const int b = blockIdx.x;
  const Cand* cd = cand + (long long)b * cap;
  if (threadIdx.x != 0) return;
  int E = 0;
  for (int c = 0; c < num_class; ++c) E += sel_cnt[b * num_class + c];
  int written = 0;
  for (int c = 0; c < num_class; ++c) {
    const int* s = sel + ((long long)b * num_class + c) * max_det;
    for (int i = 0; i < sel_cnt[b * num_class + c]; ++i) {
      const Cand me = cd[s[i]];
      int r = 0;
      for (int c2 = 0; c2 < num_class; ++c2) {
        const int* s2 = sel + ((long long)b * num_class + c2) * max_det;
        for (int j = 0; j < sel_cnt[b * num_class + c2]; ++j) {
          const Cand o = cd[s2[j]];
          if (o.score > me.score || (o.score == me.score && o.idx < me.idx)) ++r;
        }
      }
      if (r < max_det) {
        sel_idx[b * max_det + r] = me.idx;
        ++written;
      }
    }
  }
  for (int r = written; r < max_det; ++r) sel_idx[b * max_det + r] = -1;
  sel_count[b] = written;
}
}  // namespace dy

namespace dy {
__global__ void edges_from_boxes_kernel(const float* det_box, const int* det_count, int max_det, int S, int k,
                                        int* edges) {
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < det_count[b] && d < max_det; d += blockDim.x) {
    const float* row = det_box + ((long long)b * max_det + d) * 6;
    const float Sf = (float)S;
    float pb[4];
    for (int i = 0; i < 4; ++i) pb[i] = rintf(__fmul_rn(row[i], Sf));
    const float sub_w = __fdiv_rn(__fsub_rn(pb[3], pb[1]), (float)k);
    const float sub_h = __fdiv_rn(__fsub_rn(pb[2], pb[0]), (float)k);
    int* ed = edges + ((long long)b * max_det + d) * (2 * (kMaxK + 1));
    ed[0] = (int)pb[1];
    ed[kMaxK + 1] = (int)pb[0];
    for (int j = 1; j < k; ++j) {
      ed[j] = (int)rintf(__fadd_rn(pb[1], __fmul_rn((float)j, sub_w)));
      ed[kMaxK + 1 + j] = (int)rintf(__fadd_rn(pb[0], __fmul_rn((float)j, sub_h)));
    }
    ed[k] = (int)pb[3];
    ed[kMaxK + 1 + k] = (int)pb[2];
  }
}
}  // namespace dy

namespace dy {
static int sync_masters(dy_net* net, bool* any_new);
static int retrain_refresh(dy_net* net, bool any_new);
}  // namespace dy

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* dy_version(void) { return "disyolo_b200 0.1 sm_100a"; }
const char* dy_last_error(void) { return g_last_error.c_str(); }

int dy_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int dy_set_option(const char* name, int32_t value) {
  DY_CHECK(name != nullptr, "null option name");
  const std::string n(name);
  if (n == "tc_resident") g_opt_resident = value;
  else if (n == "tc_halo") g_opt_halo = value;
  else if (n == "tc_staged") g_opt_staged = value;
  else if (n == "tc_tma_epi") g_opt_tma_epi = value;
  else if (n == "conv1_tc") g_opt_conv1_tc = value;
  else if (n == "tc_fuse_tail") g_opt_fuse_tail = value;
  else if (n == "tc_split_n") g_opt_split_n = value;
  else if (n == "tc_cta2") g_opt_cta2 = value;
  else if (n == "tc_epi_wg") g_opt_epi_wg = value;
  else if (n == "tc_kchunk") g_opt_kchunk = value;
  else if (n == "tc_slab") g_opt_slab = value;
  else if (n == "tc_pdl") conv_tc_set_pdl(value < 0 ? 1 : value);
  else if (n == "train_pdl") train_set_pdl(value < 0 ? 1 : value);
  else if (n == "tc_dual_producer") g_opt_dual_producer = value;
  else if (n == "tc_max_stages") g_opt_max_stages = value;
  else if (n == "tc_skip_epilogue") g_opt_skip_epi = value;
  else if (n == "tc_dual_issue") g_opt_dual = value;
  else if (n == "mask_streaming_stores") masks_set_streaming(value);
  else if (n == "mask_work_list") masks_set_work_list(value < 0 ? 0 : value);
  else if (n == "wgrad_fuse_kw") wgrad_set_fuse(value);
  else if (n == "wgrad_lbo_a") { g_wg_dbg[0] = value; wgrad_set_debug(g_wg_dbg[0], g_wg_dbg[1], g_wg_dbg[2], g_wg_dbg[3]); }
  else if (n == "wgrad_sbo_a") { g_wg_dbg[1] = value; wgrad_set_debug(g_wg_dbg[0], g_wg_dbg[1], g_wg_dbg[2], g_wg_dbg[3]); }
  else if (n == "wgrad_lbo_b") { g_wg_dbg[2] = value; wgrad_set_debug(g_wg_dbg[0], g_wg_dbg[1], g_wg_dbg[2], g_wg_dbg[3]); }
  else if (n == "wgrad_sbo_b") { g_wg_dbg[3] = value; wgrad_set_debug(g_wg_dbg[0], g_wg_dbg[1], g_wg_dbg[2], g_wg_dbg[3]); }
  else {
    set_error("unknown option " + n);
    return DY_ERR_NOTFOUND;
  }
  return DY_OK;
}

int64_t dy_launch_count(int32_t reset) {
  long long v = g_launches.load();
  if (reset) g_launches.store(0);
  return v;
}

int dy_create(const dy_config* cfg, dy_net** out) {
  DY_CHECK(cfg != nullptr && out != nullptr, "null argument");
  DY_CHECK(cfg->image_size >= 32 && cfg->image_size % 32 == 0, "image_size must be a positive multiple of 32");
  DY_CHECK(cfg->num_classes == 3, "the 24-channel heads of convolutional59/67/75 imply 3 classes");
  DY_CHECK(cfg->k_map == 3, "convolutional82 emits 9 = 3x3 score maps");
  DY_CHECK(cfg->max_batch >= 1, "max_batch");
  DY_CHECK(cfg->max_detection >= 1 && cfg->max_detection <= 4096, "max_detection");
  DY_CHECK(cfg->precision == DY_PRECISION_BF16 || cfg->precision == DY_PRECISION_FP32, "precision");
  int ndev = dy_device_count();
  if (ndev <= 0) {
    set_error("no CUDA device: libdisyolo_b200 has no CPU fallback");
    return DY_ERR_CUDA;
  }
  DY_CHECK(cfg->device >= 0 && cfg->device < ndev, "device ordinal");
  DY_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  DY_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (cfg->precision == DY_PRECISION_BF16 && prop.major != 10) {
    set_error(std::string("the bf16 engine needs an sm_100 GPU (tcgen05/TMEM); found ") + prop.name);
    return DY_ERR_UNSUPPORTED;
  }
  dy_net* net = new dy_net();
  net->cfg = *cfg;
  net->S = cfg->image_size;
  net->num_sms = prop.multiProcessorCount;
  std::vector<LayerDef> defs = build_defs(net->S);
  net->L.resize(83);
  for (int n = 1; n <= 82; ++n) net->L[n].def = defs[n];
  int rc = allocate_buffers(net);
  if (rc != DY_OK) {
    dy_destroy(net);
    return rc;
  }
  *out = net;
  return DY_OK;
}

int dy_destroy(dy_net* net) {
  if (!net) return DY_OK;
  cudaSetDevice(net->cfg.device);
  cudaDeviceSynchronize();
  for (void* p : net->allocs) cudaFree(p);
  for (auto st : {net->h2d_stream, net->comp_stream, net->d2h_stream})
    if (st) cudaStreamDestroy(st);
  for (auto& sl : net->slot) {
    if (sl.ev_h2d) cudaEventDestroy(sl.ev_h2d);
    if (sl.ev_comp) cudaEventDestroy(sl.ev_comp);
    if (sl.small_host) cudaFreeHost(sl.small_host);
  }
  if (net->ev_user) cudaEventDestroy(net->ev_user);
  if (net->ev_host) cudaEventDestroy(net->ev_host);
  delete net;
  return DY_OK;
}

int dy_load_weights(dy_net* net, const char* tf_name, const float* host, const int64_t* shape, int32_t ndim) {
  DY_CHECK(net && tf_name && host && shape, "null argument");
  std::string name(tf_name);
  const std::string prefix = "yolo/convolutional";
  if (name.compare(0, prefix.size(), prefix) != 0) {
    set_error("unknown variable " + name);
    return DY_ERR_NOTFOUND;
  }
  size_t p = prefix.size();
  int n = 0;
  while (p < name.size() && name[p] >= '0' && name[p] <= '9') n = n * 10 + (name[p++] - '0');
  if (n < 1 || n > 82 || p >= name.size() || name[p] != '/') {
    set_error("unknown variable " + name);
    return DY_ERR_NOTFOUND;
  }
  const std::string what = name.substr(p + 1);
  LayerState& s = net->L[n];
  const LayerDef& d = s.def;
  size_t count = 1;
  for (int i = 0; i < ndim; ++i) count *= (size_t)shape[i];
  std::vector<float>* dst = nullptr;
  unsigned bit = 0;
  if (what == "weights") {
    DY_CHECK(ndim == 4 && shape[0] == d.k && shape[1] == d.k && shape[2] == d.cin0 + d.cin1 && shape[3] == d.cout,
             "weights must be HWIO [k,k,cin,cout]");
    dst = &s.w; bit = 1u << 0;
  } else {
    DY_CHECK(ndim == 1 && shape[0] == d.cout, "per-channel variable must be [cout]");
    if (what == "biases") { dst = &s.bias; bit = 1u << 5; }
    else if (what == "BatchNorm/gamma") { dst = &s.gamma; bit = 1u << 1; }
    else if (what == "BatchNorm/beta") { dst = &s.beta; bit = 1u << 2; }
    else if (what == "BatchNorm/moving_mean") { dst = &s.mean; bit = 1u << 3; }
    else if (what == "BatchNorm/moving_variance") { dst = &s.var; bit = 1u << 4; }
  }
  if (!dst) {
    set_error("unknown variable " + name);
    return DY_ERR_NOTFOUND;
  }
  dst->assign(host, host + count);
  s.host_new |= bit;
  net->finalized = false;
  return DY_OK;
}

int dy_finalize_weights(dy_net* net) {
  DY_CHECK(net, "null net");
  DY_CUDA(cudaSetDevice(net->cfg.device));
  DY_CUDA(cudaDeviceSynchronize());                  // nothing in flight may still read the old operands
  // Once dy_train_init has run the truth lives in the device master copies: variables loaded since the
  // last finalize (Saver.restore, train_yolo3_mask.py:104-111) overwrite their masters, every other
  // variable's host copy is refreshed FROM its master, so that the fold below never reverts training.
  bool any_new = false;
  if (net->train_ready) DY_TRY(sync_masters(net, &any_new));
  for (int n = 1; n <= 82; ++n) DY_TRY(fold_and_upload(net, n));
  if (net->cfg.precision == DY_PRECISION_BF16)
    for (int n = 2; n <= 82; ++n) DY_TRY(plan_layer(net, n));
  if (net->train_ready) DY_TRY(retrain_refresh(net, any_new));
  for (int n = 1; n <= 82; ++n) net->L[n].host_new = 0;
  DY_CUDA(cudaDeviceSynchronize());
  net->finalized = true;
  return DY_OK;
}

int dy_set_loss_params(dy_net* net, float object_scale, float noobject_scale, float class_scale, float coord_scale,
                       float mask_scale, float ignore_thresh) {
  DY_CHECK(net, "null net");
  net->object_scale = object_scale; net->noobject_scale = noobject_scale; net->class_scale = class_scale;
  net->coord_scale = coord_scale; net->mask_scale = mask_scale; net->ignore_thresh = ignore_thresh;
  return DY_OK;
}

int dy_forward_profile(dy_net* net, const float* images_dev, int32_t B, float* layer_ms_host, void* stream) {
  DY_CHECK(net && images_dev && layer_ms_host, "null argument");
  DY_CHECK(net->finalized, "dy_finalize_weights has not been called");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch exceeds max_batch");
  DY_CHECK(net->cfg.precision == DY_PRECISION_BF16, "per-layer profile is for the bf16 engine");
  cudaStream_t st = (cudaStream_t)stream;
  DY_TRY(user_begin(net, st));
  std::vector<cudaEvent_t> ev(83);
  for (auto& e : ev) DY_CUDA(cudaEventCreate(&e));
  auto& L = net->L;
  int rc = DY_OK;
  cudaEventRecord(ev[0], st);
  note_launch();
  rc = launch_conv1(images_dev, L[1].d_w_f32, L[1].d_scale, L[1].d_shift, net->cfg.alpha, B, net->S, net->S,
                    L[1].s2d, L[1].same, g_opt_conv1_tc != 0, net->num_sms, st);
  cudaEventRecord(ev[1], st);
  // the launches dy_forward issues: with the fused tail convolutional82 has no launch of its own (0 ms)
  const bool fused = L[81].has_fused;
  for (int n = 2; n <= 82 && rc == DY_OK; ++n) {
    if (fused && n == 81) rc = run_tc_plan(L[81].plan_fused, B, net->num_sms, st);
    else if (!(fused && n == 82)) rc = run_tc_plan(L[n].plan, B, net->num_sms, st);
    cudaEventRecord(ev[n], st);
  }
  net->last_fused = fused;
  if (rc == DY_OK && cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("stream synchronize failed in dy_forward_profile");
    rc = DY_ERR_CUDA;
  }
  if (rc == DY_OK) {
    layer_ms_host[0] = 0.f;
    for (int n = 1; n <= 82; ++n) cudaEventElapsedTime(&layer_ms_host[n], ev[n - 1], ev[n]);
  }
  for (auto& e : ev) cudaEventDestroy(e);
  if (rc == DY_OK) rc = user_end(net, st);
  return rc;
}

int dy_forward_network(dy_net* net, const float* images_dev, int32_t B, void* stream) {
  DY_CHECK(net && images_dev, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DY_TRY(user_begin(net, st));
  DY_TRY(run_network(net, images_dev, B, st, /*fused=*/false));     // parity taps: every layer materialised
  return user_end(net, st);
}

static void head_ptrs(dy_net* net, const float** y8, const float** y16, const float** y32, const float** mp,
                      int* layout) {
  *y8 = net->L[75].f32;
  *y16 = net->L[67].f32;
  *y32 = net->L[59].f32;
  *mp = net->L[82].f32;
  *layout = net->cfg.precision == DY_PRECISION_BF16 ? 1 : 0;
}

// network + decode + NMS + top-k (+ masks) on `st`; crop_* non-null selects the box-cropped mask form
static int forward_impl(dy_net* net, const float* images_dev, int32_t B, const float* windows_dev, float det_thresh,
                        float* det_raw_dev, float* det_box_dev, int32_t* det_count_dev, float* masks_dev,
                        long long* crop_off_dev, float* crops_dev, cudaStream_t st) {
  DY_TRY(run_network(net, images_dev, B, st, /*fused=*/true));
  const float *y8, *y16, *y32, *mp;
  int layout;
  head_ptrs(net, &y8, &y16, &y32, &mp, &layout);
  float* box = det_box_dev ? det_box_dev : net->det_box_ws;
  int* cnt = det_count_dev ? det_count_dev : net->det_count_ws;
  DY_TRY(run_detect(net, y8, y16, y32, B, windows_dev, det_thresh, nullptr, nullptr, nullptr, det_raw_dev, box, cnt,
                    st));
  if (masks_dev) DY_TRY(run_masks(net, mp, layout, B, cnt, masks_dev, st));
  if (crops_dev) {
    MaskArgs ma = mask_args(net, mp, layout, B, cnt, nullptr);
    note_launch(2);
    DY_TRY(launch_crop_offsets(ma, crop_off_dev, st));
    DY_TRY(launch_masks_cropped(ma, crop_off_dev, crops_dev, st));
  }
  return DY_OK;
}

int dy_forward(dy_net* net, const float* images_dev, int32_t B, const float* windows_dev, float det_thresh,
               float* det_raw_dev, float* det_box_dev, int32_t* det_count_dev, float* masks_dev, void* stream) {
  NvtxRange nvtx_("dy_forward");
  DY_CHECK(net && images_dev && windows_dev, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DY_TRY(user_begin(net, st));
  DY_TRY(forward_impl(net, images_dev, B, windows_dev, det_thresh, det_raw_dev, det_box_dev, det_count_dev, masks_dev,
                      nullptr, nullptr, st));
  return user_end(net, st);
}

// small results of one slot, contiguous so that they leave in ONE device -> host copy:
// [det_count B*4 | det_box B*md*24 | det_raw B*md*24 | crop offsets (B*md+1)*8], each part 16-byte aligned
static size_t small_part(size_t bytes) { return (bytes + 15) & ~(size_t)15; }
static size_t small_bytes(const dy_net* net) {
  const size_t MB = net->cfg.max_batch, md = net->cfg.max_detection;
  return small_part(MB * 4) + 2 * small_part(MB * md * 24) + small_part((MB * md + 1) * 8);
}

static int host_slots_init(dy_net* net) {
  if (net->h2d_stream) return DY_OK;
  const int S = net->S, MB = net->cfg.max_batch;
  DY_CUDA(cudaStreamCreateWithFlags(&net->h2d_stream, cudaStreamNonBlocking));
  DY_CUDA(cudaStreamCreateWithFlags(&net->comp_stream, cudaStreamNonBlocking));
  DY_CUDA(cudaStreamCreateWithFlags(&net->d2h_stream, cudaStreamNonBlocking));
  DY_CUDA(cudaEventCreateWithFlags(&net->ev_host, cudaEventDisableTiming));
  const size_t md = net->cfg.max_detection;
  for (auto& sl : net->slot) {
    DY_TRY(dev_alloc(net, (void**)&sl.images, (size_t)MB * S * S * 3 * 4, false));
    DY_TRY(dev_alloc(net, (void**)&sl.windows, (size_t)MB * 16, false));
    DY_TRY(dev_alloc(net, (void**)&sl.small, small_bytes(net), false));
    // carve the small-result block
    uint8_t* q = reinterpret_cast<uint8_t*>(sl.small);
    sl.det_count = reinterpret_cast<int*>(q);          q += small_part((size_t)MB * 4);
    sl.det_box = reinterpret_cast<float*>(q);          q += small_part((size_t)MB * md * 24);
    sl.det_raw = reinterpret_cast<float*>(q);          q += small_part((size_t)MB * md * 24);
    sl.crop_off = reinterpret_cast<long long*>(q);
    // pinned host mirror (allocated by the calling thread: NUMA-local under the caller's CPU affinity)
    DY_CUDA(cudaHostAlloc((void**)&sl.small_host, small_bytes(net), cudaHostAllocDefault));
    DY_CUDA(cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming));
    DY_CUDA(cudaEventCreateWithFlags(&sl.ev_comp, cudaEventDisableTiming));
  }
  return DY_OK;
}

// the [B,max_det,S,S] mask block (full or cropped form: the cropped maps never exceed it) and the uint8
// image staging are allocated on first use
static int slot_need(dy_net* net, dy_net::HostSlot& sl, bool masks, bool u8) {
  const int S = net->S, Sm = S / 2, md = net->cfg.max_detection, MB = net->cfg.max_batch;
  if (masks && !sl.masks) DY_TRY(dev_alloc(net, (void**)&sl.masks, (size_t)MB * md * Sm * Sm * 4, false));
  if (u8 && !sl.images_u8) DY_TRY(dev_alloc(net, (void**)&sl.images_u8, (size_t)MB * S * S * 3, false));
  return DY_OK;
}

static int host_begin(dy_net* net, const void* images_host, bool u8, int32_t B, const float* windows_host,
                      float det_thresh, int32_t mask_mode, int32_t* ticket) {
  DY_CHECK(net && images_host && windows_host && ticket, "null argument");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch exceeds max_batch");
  DY_CHECK(mask_mode >= DY_MASKS_NONE && mask_mode <= DY_MASKS_CROPPED, "mask_mode");
  DY_CUDA(cudaSetDevice(net->cfg.device));
  DY_TRY(host_slots_init(net));
  const int id = net->next_slot;
  auto& sl = net->slot[id];
  DY_CHECK(!sl.busy, "all three pipeline slots are in flight: call dy_forward_host_end first");
  DY_TRY(slot_need(net, sl, mask_mode != DY_MASKS_NONE, u8));
  const int S = net->S;
  const size_t npix = (size_t)B * S * S * 3;
  if (u8) DY_CUDA(cudaMemcpyAsync(sl.images_u8, images_host, npix, cudaMemcpyHostToDevice, net->h2d_stream));
  else DY_CUDA(cudaMemcpyAsync(sl.images, images_host, npix * 4, cudaMemcpyHostToDevice, net->h2d_stream));
  DY_CUDA(cudaMemcpyAsync(sl.windows, windows_host, (size_t)B * 16, cudaMemcpyHostToDevice, net->h2d_stream));
  DY_CUDA(cudaEventRecord(sl.ev_h2d, net->h2d_stream));
  DY_CUDA(cudaStreamWaitEvent(net->comp_stream, sl.ev_h2d, 0));
  // the caller's own streams may still be using the activation / weight buffers (training step, dy_forward)
  if (net->user_work) DY_CUDA(cudaStreamWaitEvent(net->comp_stream, net->ev_user, 0));
  if (u8) {
    note_launch();
    DY_TRY(launch_u8_to_f32(sl.images_u8, sl.images, (long long)npix, net->comp_stream));
  }
  const bool crop = mask_mode == DY_MASKS_CROPPED;
  DY_TRY(forward_impl(net, sl.images, B, sl.windows, det_thresh, sl.det_raw, sl.det_box, sl.det_count,
                      mask_mode == DY_MASKS_FULL ? sl.masks : nullptr, crop ? sl.crop_off : nullptr,
                      crop ? sl.masks : nullptr, net->comp_stream));
  DY_CUDA(cudaEventRecord(sl.ev_comp, net->comp_stream));
  DY_CUDA(cudaEventRecord(net->ev_host, net->comp_stream));
  net->host_work = true;
  sl.B = B;
  sl.busy = true;
  sl.mask_mode = mask_mode;
  net->next_slot = (id + 1) % dy_net::kHostSlots;
  *ticket = id;
  return DY_OK;
}

int dy_forward_host_begin(dy_net* net, const float* images_host, int32_t B, const float* windows_host,
                          float det_thresh, int32_t mask_mode, int32_t* ticket) {
  return host_begin(net, images_host, false, B, windows_host, det_thresh, mask_mode, ticket);
}

int dy_forward_host_begin_u8(dy_net* net, const uint8_t* images_host, int32_t B, const float* windows_host,
                             float det_thresh, int32_t mask_mode, int32_t* ticket) {
  NvtxRange nvtx_("dy_forward_host_begin");
  return host_begin(net, images_host, true, B, windows_host, det_thresh, mask_mode, ticket);
}

// one D2H copy of the slot's small results into the pinned mirror, then plain memcpy to the caller
static int host_end_small(dy_net* net, dy_net::HostSlot& sl, float* det_raw_host, float* det_box_host,
                          int32_t* det_count_host, long long* crop_off_host) {
  const size_t MB = net->cfg.max_batch, md = net->cfg.max_detection, B = sl.B;
  cudaStream_t st = net->d2h_stream;
  DY_CUDA(cudaStreamWaitEvent(st, sl.ev_comp, 0));
  const size_t used = (sl.mask_mode == DY_MASKS_CROPPED || det_raw_host)
                          ? small_bytes(net) - (sl.mask_mode == DY_MASKS_CROPPED ? 0 : small_part((MB * md + 1) * 8))
                          : small_part(MB * 4) + small_part(MB * md * 24);
  DY_CUDA(cudaMemcpyAsync(sl.small_host, sl.small, used, cudaMemcpyDeviceToHost, st));
  DY_CUDA(cudaStreamSynchronize(st));
  const uint8_t* q = sl.small_host;
  memcpy(det_count_host, q, B * 4);                      q += small_part(MB * 4);
  memcpy(det_box_host, q, B * md * 24);                  q += small_part(MB * md * 24);
  if (det_raw_host) memcpy(det_raw_host, q, B * md * 24);
  q += small_part(MB * md * 24);
  if (crop_off_host) memcpy(crop_off_host, q, (B * md + 1) * 8);
  return DY_OK;
}

int dy_forward_host_end(dy_net* net, int32_t ticket, float* det_raw_host, float* det_box_host,
                        int32_t* det_count_host, float* masks_host) {
  NvtxRange nvtx_("dy_forward_host_end");
  DY_CHECK(net && det_box_host && det_count_host, "null argument");
  DY_CHECK(ticket >= 0 && ticket < dy_net::kHostSlots, "bad ticket");
  auto& sl = net->slot[ticket];
  DY_CHECK(sl.busy, "ticket is not in flight");
  DY_CHECK(sl.mask_mode != DY_MASKS_CROPPED, "this ticket carries cropped masks: call dy_forward_host_end_cropped");
  DY_CHECK(!masks_host || sl.mask_mode == DY_MASKS_FULL, "masks were not requested at dy_forward_host_begin");
  const int Sm = net->S / 2, md = net->cfg.max_detection, B = sl.B;
  DY_TRY(host_end_small(net, sl, det_raw_host, det_box_host, det_count_host, nullptr));
  if (masks_host) {
    // reference layout [B,max_det,S,S]: exactly det_count[b] maps per image; runs of images whose maps are
    // all valid (n == max_det) or empty are merged so that the copy count stays small
    cudaStream_t st = net->d2h_stream;
    const size_t per = (size_t)Sm * Sm;
    for (int b = 0; b < B; ++b) {
      const int n = det_count_host[b];
      if (n > 0)
        DY_CUDA(cudaMemcpyAsync(masks_host + (size_t)b * md * per, sl.masks + (size_t)b * md * per,
                                (size_t)n * per * 4, cudaMemcpyDeviceToHost, st));
    }
    DY_CUDA(cudaStreamSynchronize(st));
  }
  sl.busy = false;
  return DY_OK;
}

int dy_forward_host_end_cropped(dy_net* net, int32_t ticket, float* det_raw_host, float* det_box_host,
                                int32_t* det_count_host, int64_t* crop_offsets_host, float* crops_host,
                                int64_t crops_capacity) {
  NvtxRange nvtx_("dy_forward_host_end_cropped");
  DY_CHECK(net && det_box_host && det_count_host && crop_offsets_host && crops_host, "null argument");
  DY_CHECK(ticket >= 0 && ticket < dy_net::kHostSlots, "bad ticket");
  auto& sl = net->slot[ticket];
  DY_CHECK(sl.busy, "ticket is not in flight");
  DY_CHECK(sl.mask_mode == DY_MASKS_CROPPED, "cropped masks were not requested at dy_forward_host_begin");
  const int md = net->cfg.max_detection, B = sl.B;
  DY_TRY(host_end_small(net, sl, det_raw_host, det_box_host, det_count_host,
                        reinterpret_cast<long long*>(crop_offsets_host)));
  const long long total = crop_offsets_host[(size_t)B * md];
  if (total > crops_capacity) {
    sl.busy = false;
    set_error("crops_host is too small: " + std::to_string(total) + " floats needed");
    return DY_ERR_INVALID;
  }
  if (total > 0) {
    DY_CUDA(cudaMemcpyAsync(crops_host, sl.masks, (size_t)total * 4, cudaMemcpyDeviceToHost, net->d2h_stream));
    DY_CUDA(cudaStreamSynchronize(net->d2h_stream));
  }
  sl.busy = false;
  return DY_OK;
}

int dy_forward_host(dy_net* net, const float* images_host, int32_t B, const float* windows_host, float det_thresh,
                    float* det_raw_host, float* det_box_host, int32_t* det_count_host, float* masks_host) {
  DY_CHECK(net && images_host && windows_host && det_box_host && det_count_host, "null argument");
  int32_t ticket = -1;
  DY_TRY(dy_forward_host_begin(net, images_host, B, windows_host, det_thresh,
                               masks_host != nullptr ? DY_MASKS_FULL : DY_MASKS_NONE, &ticket));
  return dy_forward_host_end(net, ticket, det_raw_host, det_box_host, det_count_host, masks_host);
}

int dy_layer_shape(dy_net* net, int32_t layer, int32_t* h, int32_t* w, int32_t* c) {
  DY_CHECK(net && layer >= 1 && layer <= 82, "layer must be 1..82");
  const LayerDef& d = net->L[layer].def;
  if (h) *h = d.H;
  if (w) *w = d.H;
  if (c) *c = d.cout;
  return DY_OK;
}

int dy_get_activation(dy_net* net, int32_t layer, int32_t B, float* out_dev, void* stream) {
  DY_CHECK(net && out_dev && layer >= 1 && layer <= 82, "bad argument");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch");
  cudaStream_t st = (cudaStream_t)stream;
  const LayerState& s = net->L[layer];
  const LayerDef& d = s.def;
  const size_t n = (size_t)B * d.H * d.H * d.cout;
  if (layer == 81 && net->last_fused) {
    set_error("convolutional81 was not materialised by the last forward (fused tail): use dy_forward_network");
    return DY_ERR_STATE;
  }
  if (net->cfg.precision == DY_PRECISION_FP32) {
    DY_CUDA(cudaMemcpyAsync(out_dev, s.f32, n * 4, cudaMemcpyDeviceToDevice, st));
    return DY_OK;
  }
  note_launch();
  if (!d.bn) {
    if (layer == 82) return launch_planar_to_nhwc(s.f32, out_dev, B, d.H, d.H, d.cout, st);
    DY_CUDA(cudaMemcpyAsync(out_dev, s.f32, n * 4, cudaMemcpyDeviceToDevice, st));
    return DY_OK;
  }
  if (s.same) return launch_p1_to_nhwc(s.same, out_dev, B, d.H, d.H, d.cout, FORM_SAME, st);
  if (s.s2d) return launch_p1_to_nhwc(s.s2d, out_dev, B, d.H, d.H, d.cout, FORM_S2D, st);
  if (s.up) return launch_p1_to_nhwc(s.up, out_dev, B, d.H, d.H, d.cout, FORM_UP2, st);
  set_error("layer has no buffer");
  return DY_ERR_STATE;
}

int dy_get_yolo(dy_net* net, int32_t scale, int32_t B, float* out_dev, void* stream) {
  DY_CHECK(net && out_dev && scale >= 0 && scale <= 2, "bad argument");
  static const int layer_of[3] = {75, 67, 59};
  return dy_get_activation(net, layer_of[scale], B, out_dev, stream);
}

int dy_get_mask_pos(dy_net* net, int32_t B, float* out_dev, void* stream) {
  return dy_get_activation(net, 82, B, out_dev, stream);
}

int dy_decode(dy_net* net, const float* yolo8_dev, const float* yolo16_dev, const float* yolo32_dev, int32_t B,
              const float* windows_dev, float* box_dev, int32_t* cls_dev, float* score_dev, void* stream) {
  DY_CHECK(net && yolo8_dev && yolo16_dev && yolo32_dev && windows_dev && box_dev && cls_dev && score_dev,
           "null argument");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch");
  return run_detect(net, yolo8_dev, yolo16_dev, yolo32_dev, B, windows_dev, 3.0e38f, box_dev, cls_dev, score_dev,
                    nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int dy_detect(dy_net* net, const float* yolo8_dev, const float* yolo16_dev, const float* yolo32_dev, int32_t B,
              const float* windows_dev, float det_thresh, float* det_raw_dev, float* det_box_dev,
              int32_t* det_count_dev, void* stream) {
  DY_CHECK(net && yolo8_dev && yolo16_dev && yolo32_dev && windows_dev && det_box_dev && det_count_dev,
           "null argument");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch");
  return run_detect(net, yolo8_dev, yolo16_dev, yolo32_dev, B, windows_dev, det_thresh, nullptr, nullptr, nullptr,
                    det_raw_dev, det_box_dev, det_count_dev, (cudaStream_t)stream);
}

int dy_nms(dy_net* net, const float* box_dev, const int32_t* cls_dev, const float* score_dev, int32_t B, int32_t N,
           float det_thresh, int32_t* sel_idx_dev, int32_t* sel_count_dev, float* det_raw_dev, void* stream) {
  DY_CHECK(net && box_dev && cls_dev && score_dev && sel_idx_dev && sel_count_dev, "null argument");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch");
  DY_CHECK(N >= 1 && N <= net->cap, "N exceeds the net's candidate capacity");
  cudaStream_t st = (cudaStream_t)stream;
  DY_CUDA(cudaMemsetAsync(net->cand_count, 0, (size_t)B * 4, st));
  dim3 grid((N + 255) / 256, B);
  note_launch(2);
  pack_cands_kernel<<<grid, 256, 0, st>>>(box_dev, cls_dev, score_dev, N, det_thresh, net->cand, net->cand_count,
                                          net->cap);
  DY_CUDA(cudaGetLastError());
  NmsArgs na;
  na.cand = net->cand; na.cand_count = net->cand_count; na.cap = net->cap;
  na.B = B; na.num_class = net->cfg.num_classes; na.max_det = net->cfg.max_detection;
  na.iou_thr = net->cfg.iou_threshold;
  na.sel = net->sel; na.sel_cnt = net->sel_cnt;
  na.ovf_box = net->nms_box; na.ovf_key = net->nms_key; na.ovf_pos = net->nms_pos;
  DY_TRY(launch_nms(na, st));
  FinalizeArgs fa;
  fa.cand = net->cand; fa.cap = net->cap; fa.B = B; fa.num_class = net->cfg.num_classes;
  fa.max_det = net->cfg.max_detection; fa.sel = net->sel; fa.sel_cnt = net->sel_cnt;
  fa.S = net->S / 2; fa.k = net->cfg.k_map;
  fa.det_raw = det_raw_dev ? det_raw_dev : net->det_raw_ws;
  fa.raw_count = net->raw_count;
  fa.det_box = net->det_box_ws;
  fa.det_count = net->det_count_ws;
  fa.edges = net->edges;
  note_launch(2);
  DY_TRY(launch_finalize(fa, st));
  sel_to_idx_kernel<<<B, 32, 0, st>>>(net->cand, net->cap, fa.det_raw, net->raw_count, fa.max_det, net->sel,
                                      net->sel_cnt, fa.num_class, sel_idx_dev, sel_count_dev);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int dy_assemble_masks(dy_net* net, const float* score_dev, int32_t layout, int32_t B, const float* det_box_dev,
                      const int32_t* det_count_dev, float* masks_dev, void* stream) {
  DY_CHECK(net && score_dev && det_box_dev && det_count_dev && masks_dev, "null argument");
  DY_CHECK(layout == 0 || layout == 1, "layout: 0 = NHWC, 1 = planar");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch");
  cudaStream_t st = (cudaStream_t)stream;
  note_launch();
  edges_from_boxes_kernel<<<B, 64, 0, st>>>(det_box_dev, det_count_dev, net->cfg.max_detection, net->S / 2,
                                            net->cfg.k_map, net->edges);
  DY_CUDA(cudaGetLastError());
  return run_masks(net, score_dev, layout, B, det_count_dev, masks_dev, st);
}

int dy_postproc_profile(dy_net* net, const float* yolo8_dev, const float* yolo16_dev, const float* yolo32_dev,
                        const float* score_dev, int32_t layout, int32_t B, const float* windows_dev, float det_thresh,
                        float* masks_dev, int32_t reps, float* ms_host, void* stream) {
  DY_CHECK(net && yolo8_dev && yolo16_dev && yolo32_dev && score_dev && windows_dev && masks_dev && ms_host,
           "null argument");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch && reps >= 1, "batch / reps");
  DY_CHECK(layout == 0 || layout == 1, "layout: 0 = NHWC, 1 = planar");
  cudaStream_t st = (cudaStream_t)stream;
  DecodeArgs da;
  memset(&da, 0, sizeof(da));
  da.yolo[0] = yolo8_dev; da.yolo[1] = yolo16_dev; da.yolo[2] = yolo32_dev;
  da.g[0] = net->S / 8; da.g[1] = net->S / 16; da.g[2] = net->S / 32;
  da.B = B; da.num_class = net->cfg.num_classes; da.net = 32 * da.g[2];
  memcpy(da.anchors, net->cfg.anchors, sizeof(da.anchors));
  da.windows = windows_dev; da.thresh = det_thresh;
  da.cand = net->cand; da.cand_count = net->cand_count; da.cap = net->cap;
  NmsArgs na;
  na.cand = net->cand; na.cand_count = net->cand_count; na.cap = net->cap;
  na.B = B; na.num_class = net->cfg.num_classes; na.max_det = net->cfg.max_detection;
  na.iou_thr = net->cfg.iou_threshold; na.sel = net->sel; na.sel_cnt = net->sel_cnt;
  na.ovf_box = net->nms_box; na.ovf_key = net->nms_key; na.ovf_pos = net->nms_pos;
  FinalizeArgs fa;
  fa.cand = net->cand; fa.cap = net->cap; fa.B = B; fa.num_class = net->cfg.num_classes;
  fa.max_det = net->cfg.max_detection; fa.sel = net->sel; fa.sel_cnt = net->sel_cnt;
  fa.S = net->S / 2; fa.k = net->cfg.k_map;
  fa.det_raw = net->det_raw_ws; fa.raw_count = net->raw_count; fa.det_box = net->det_box_ws;
  fa.det_count = net->det_count_ws; fa.edges = net->edges;
  cudaEvent_t ev[5];
  for (int i = 0; i < 5; ++i) DY_CUDA(cudaEventCreate(&ev[i]));
  int rc = DY_OK;
  // each stage `reps` times back to back between two events; stages see the previous stage's output
  DY_CUDA(cudaEventRecord(ev[0], st));
  for (int r = 0; r < reps && rc == DY_OK; ++r) {
    cudaMemsetAsync(net->cand_count, 0, (size_t)B * 4, st);
    note_launch();
    rc = launch_decode(da, st);
  }
  DY_CUDA(cudaEventRecord(ev[1], st));
  for (int r = 0; r < reps && rc == DY_OK; ++r) { note_launch(); rc = launch_nms(na, st); }
  DY_CUDA(cudaEventRecord(ev[2], st));
  for (int r = 0; r < reps && rc == DY_OK; ++r) { note_launch(); rc = launch_finalize(fa, st); }
  DY_CUDA(cudaEventRecord(ev[3], st));
  for (int r = 0; r < reps && rc == DY_OK; ++r) rc = run_masks(net, score_dev, layout, B, net->det_count_ws, masks_dev, st);
  DY_CUDA(cudaEventRecord(ev[4], st));
  DY_CUDA(cudaEventSynchronize(ev[4]));
  for (int i = 0; i < 4; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
    ms_host[i] = ms / (float)reps;
  }
  for (int i = 0; i < 5; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

static void letterbox_window(const LetterboxGeom& g, int image_size, float* window) {
  // clip window for filter_detections (calculate_test_map.py:163-168)
  window[0] = (float)((double)g.top / image_size);
  window[1] = (float)((double)g.left / image_size);
  window[2] = (float)((double)(g.new_h + g.top) / image_size);
  window[3] = (float)((double)(g.new_w + g.left) / image_size);
}

int dy_letterbox(const uint8_t* rgb_dev, int32_t h, int32_t w, int32_t image_size, float* out_dev, float* window_host,
                 void* stream) {
  return dy_letterbox_batch(rgb_dev, 0, 1, h, w, image_size, out_dev, window_host, stream);
}

int dy_letterbox_batch(const uint8_t* rgb_dev, int64_t image_stride, int32_t B, int32_t h, int32_t w,
                       int32_t image_size, float* out_dev, float* windows_host, void* stream) {
  DY_CHECK(rgb_dev && out_dev, "null argument");
  DY_CHECK(B >= 1 && h >= 1 && w >= 1 && image_size >= 1, "image geometry");
  DY_CHECK(B == 1 || image_stride >= (int64_t)h * w * 3, "image_stride smaller than one image");
  const LetterboxGeom g = letterbox_geom(h, w, image_size);
  if (windows_host)
    for (int b = 0; b < B; ++b) letterbox_window(g, image_size, windows_host + 4 * b);
  note_launch();
  return launch_letterbox(rgb_dev, image_stride, B, g, out_dev, (cudaStream_t)stream);
}

int dy_postprocess(const float* det_box_dev, const int32_t* det_count_dev, int32_t max_det, const float* masks_dev,
                   int32_t S, int32_t image_h, int32_t image_w, int32_t net_size, int32_t* boxes_out_dev,
                   uint8_t* valid_out_dev, uint8_t* full_masks_dev, uint8_t* merged_dev, void* stream) {
  return dy_postprocess_batch(det_box_dev, det_count_dev, 1, max_det, masks_dev, S, image_h, image_w, net_size,
                              boxes_out_dev, valid_out_dev, full_masks_dev, merged_dev, stream);
}

int dy_postprocess_batch(const float* det_box_dev, const int32_t* det_count_dev, int32_t B, int32_t max_det,
                         const float* masks_dev, int32_t S, int32_t image_h, int32_t image_w, int32_t net_size,
                         int32_t* boxes_out_dev, uint8_t* valid_out_dev, uint8_t* full_masks_dev, uint8_t* merged_dev,
                         void* stream) {
  DY_CHECK(det_box_dev && det_count_dev && masks_dev && boxes_out_dev && valid_out_dev, "null argument");
  DY_CHECK(max_det >= 1 && max_det <= 65535 && B >= 1, "max_det / batch");
  cudaStream_t st = (cudaStream_t)stream;
  PostDet* ws = nullptr;
  DY_CUDA(cudaMallocAsync((void**)&ws, (size_t)B * max_det * sizeof(PostDet), st));
  note_launch(2);
  const int rc = launch_postprocess(det_box_dev, det_count_dev, B, max_det, masks_dev, S, image_h, image_w, net_size, ws,
                                    boxes_out_dev, valid_out_dev, full_masks_dev, merged_dev, st);
  cudaFreeAsync(ws, st);
  return rc;
}

int dy_mask_overlaps(const uint8_t* masks1_dev, int32_t n1, const uint8_t* masks2_dev, int32_t n2, int64_t pixels,
                     float* overlaps_dev, void* stream) {
  DY_CHECK(masks1_dev && masks2_dev && overlaps_dev, "null argument");
  DY_CHECK(n1 >= 1 && n2 >= 1 && pixels >= 1, "empty mask set");
  cudaStream_t st = (cudaStream_t)stream;
  int* ws = nullptr;
  DY_CUDA(cudaMallocAsync((void**)&ws, (size_t)(n1 * n2 + n1 + n2) * 4, st));
  note_launch(2);
  const int rc = launch_mask_overlaps(masks1_dev, n1, masks2_dev, n2, pixels, ws, overlaps_dev, st);
  cudaFreeAsync(ws, st);
  return rc;
}

int dy_assign_labels(dy_net* net, const float* boxes_dev, const int32_t* nbox_dev, const float* place_dev,
                     const int32_t* flip_dev, int32_t B, int32_t max_box, float* yolo3_dev, float* yolo2_dev,
                     float* yolo1_dev, float* true_boxes_dev, void* stream) {
  DY_CHECK(net && boxes_dev && nbox_dev && place_dev && yolo3_dev && yolo2_dev && yolo1_dev && true_boxes_dev,
           "null argument");
  DY_CHECK(B >= 1 && max_box >= 1, "batch / max_box");
  LabelArgs a;
  memset(&a, 0, sizeof(a));
  a.boxes = boxes_dev; a.nbox = nbox_dev; a.place = place_dev; a.flip = flip_dev;
  a.yolo[0] = yolo3_dev; a.yolo[1] = yolo2_dev; a.yolo[2] = yolo1_dev;
  a.true_boxes = true_boxes_dev;
  a.grid[0] = net->S / 8; a.grid[1] = net->S / 16; a.grid[2] = net->S / 32;
  memcpy(a.anchors, net->cfg.anchors, sizeof(a.anchors));
  a.B = B; a.max_box = max_box; a.num_class = net->cfg.num_classes; a.net = net->S;
  note_launch();
  return launch_assign_labels(a, (cudaStream_t)stream);
}

// ---- training data pipeline on the device (utils/train_data.py) ----------------------
int dy_polygon_masks(const double* verts_dev, const int32_t* poly_dev, const int32_t* inst_dev, int32_t n_inst,
                     int32_t h, int32_t w, uint8_t* masks_dev, void* stream) {
  DY_CHECK(verts_dev && poly_dev && inst_dev && masks_dev, "null argument");
  note_launch();
  return launch_polygon_masks(verts_dev, poly_dev, inst_dev, n_inst, h, w, masks_dev, (cudaStream_t)stream);
}

int dy_mask_boxes(const uint8_t* masks_dev, int32_t n, int32_t h, int32_t w, int32_t* boxes_dev, void* stream) {
  DY_CHECK(masks_dev && boxes_dev, "null argument");
  note_launch(3);
  return launch_mask_boxes(masks_dev, n, h, w, boxes_dev, (cudaStream_t)stream);
}

static PlaceGeom place_geom(int32_t h, int32_t w, int32_t image_size, int32_t new_w, int32_t new_h, int32_t dx, int32_t dy_,
                            int32_t flip) {
  PlaceGeom g;
  g.src_h = h; g.src_w = w; g.new_w = new_w; g.new_h = new_h; g.dx = dx; g.dy = dy_; g.size = image_size; g.flip = flip;
  return g;
}

int dy_augment_image(const uint8_t* rgb_dev, int32_t h, int32_t w, int32_t image_size, int32_t new_w, int32_t new_h,
                     int32_t dx, int32_t dy_, int32_t flip, uint8_t* out_dev, void* stream) {
  DY_CHECK(rgb_dev && out_dev, "null argument");
  note_launch();
  return launch_place_image_u8(rgb_dev, place_geom(h, w, image_size, new_w, new_h, dx, dy_, flip), out_dev,
                               (cudaStream_t)stream);
}

int dy_augment_masks(const uint8_t* masks_dev, int32_t n, int32_t h, int32_t w, int32_t image_size, int32_t new_w,
                     int32_t new_h, int32_t dx, int32_t dy_, int32_t flip, uint8_t* out_dev, void* stream) {
  DY_CHECK(masks_dev && out_dev, "null argument");
  note_launch();
  return launch_place_masks(masks_dev, n, place_geom(h, w, image_size, new_w, new_h, dx, dy_, flip), out_dev,
                            (cudaStream_t)stream);
}

int dy_salt_pepper(uint8_t* img_dev, int32_t image_size, const int32_t* salt_rc_dev, int32_t n_salt,
                   const int32_t* pepper_rc_dev, int32_t n_pepper, void* stream) {
  DY_CHECK(img_dev && (n_salt == 0 || salt_rc_dev) && (n_pepper == 0 || pepper_rc_dev), "null argument");
  DY_CHECK(n_salt >= 0 && n_pepper >= 0 && image_size >= 1, "counts");
  note_launch(2);
  return launch_salt_pepper(img_dev, image_size, salt_rc_dev, n_salt, pepper_rc_dev, n_pepper, (cudaStream_t)stream);
}

int dy_change_light(uint8_t* img_dev, int64_t npix, double coeff, void* stream) {
  DY_CHECK(img_dev, "null argument");
  note_launch();
  return launch_change_light(img_dev, npix, coeff, (cudaStream_t)stream);
}

int dy_motion_blur3(const uint8_t* img_dev, int32_t image_size, const float* kernel9_host, uint8_t* out_dev, void* stream) {
  DY_CHECK(img_dev && kernel9_host && out_dev && img_dev != out_dev, "null / aliased argument");
  note_launch();
  return launch_motion_blur3(img_dev, image_size, kernel9_host, out_dev, (cudaStream_t)stream);
}

int dy_u8_to_unit_float(const uint8_t* src_dev, float* dst_dev, int64_t n, void* stream) {
  DY_CHECK(src_dev && dst_dev && n >= 1, "null argument");
  note_launch();
  return launch_u8_div255_f32(src_dev, dst_dev, n, (cudaStream_t)stream);
}

// CRC-32C (Castagnoli), slicing-by-8: checksums of TensorFlow checkpoint-V2 bundles (tf_checkpoint.py).
// Host-only utility: no device is touched.
uint32_t dy_crc32c(const void* data, uint64_t n, uint32_t crc) {
  static uint32_t T[8][256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      T[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) T[t][i] = (T[t - 1][i] >> 8) ^ T[0][T[t - 1][i] & 0xFFu];
    init = true;
  }
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= (uint64_t)c;
    c = T[7][w & 0xFF] ^ T[6][(w >> 8) & 0xFF] ^ T[5][(w >> 16) & 0xFF] ^ T[4][(w >> 24) & 0xFF] ^
        T[3][(w >> 32) & 0xFF] ^ T[2][(w >> 40) & 0xFF] ^ T[1][(w >> 48) & 0xFF] ^ T[0][(w >> 56) & 0xFF];
    p += 8;
    n -= 8;
  }
  while (n--) c = T[0][(c ^ *p++) & 0xFFu] ^ (c >> 8);
  return ~c;
}

int dy_conv_layer(int32_t precision, const float* x_dev, int32_t B, int32_t H, int32_t W, int32_t cin,
                  const float* w_host, int32_t k, int32_t stride, int32_t cout, const float* scale_host,
                  const float* shift_host, int32_t act, float alpha, const float* residual_dev, float* out_dev,
                  void* stream) {
  DY_CHECK(x_dev && w_host && scale_host && shift_host && out_dev, "null argument");
  DY_CHECK(k == 1 || k == 3, "k");
  DY_CHECK(stride == 1 || stride == 2, "stride");
  DY_CHECK(H % stride == 0 && W % stride == 0, "extent must divide the stride");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = H / stride, Wo = W / stride;
  const int K = k * k * cin;
  int rc = DY_OK;
  std::vector<void*> tmp;
  auto talloc = [&](void** p, size_t bytes) -> int {
    DY_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    tmp.push_back(*p);
    DY_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, st));
    return DY_OK;
  };
  auto cleanup = [&]() {
    cudaStreamSynchronize(st);
    for (void* p : tmp) cudaFree(p);
  };
  const int cout_pad = (cout + 15) / 16 * 16;
  std::vector<float> sc(cout_pad, 0.f), sh(cout_pad, 0.f);
  memcpy(sc.data(), scale_host, (size_t)cout * 4);
  memcpy(sh.data(), shift_host, (size_t)cout * 4);
  float *d_sc = nullptr, *d_sh = nullptr;
  if ((rc = talloc((void**)&d_sc, (size_t)cout_pad * 4)) || (rc = talloc((void**)&d_sh, (size_t)cout_pad * 4))) {
    cleanup();
    return rc;
  }
  cudaMemcpyAsync(d_sc, sc.data(), (size_t)cout_pad * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_sh, sh.data(), (size_t)cout_pad * 4, cudaMemcpyHostToDevice, st);

  if (precision == DY_PRECISION_FP32) {
    float* d_w = nullptr;
    if ((rc = talloc((void**)&d_w, (size_t)K * cout * 4))) { cleanup(); return rc; }
    cudaMemcpyAsync(d_w, w_host, (size_t)K * cout * 4, cudaMemcpyHostToDevice, st);
    RefConvArgs a;
    memset(&a, 0, sizeof(a));
    a.src0 = x_dev; a.w = d_w; a.scale = d_sc; a.shift = d_sh; a.residual = residual_dev; a.out = out_dev;
    a.Hi = H; a.Wi = W; a.Ho = Ho; a.Wo = Wo; a.c0 = cin; a.c1 = 0; a.cout = cout; a.k = k; a.s = stride;
    tf_same_pad(H, k, stride, &a.pad_t);
    tf_same_pad(W, k, stride, &a.pad_l);
    a.act = act; a.alpha = alpha;
    note_launch();
    rc = launch_conv_ref(a, B, st);
    cleanup();
    return rc;
  }

  // ---- bf16 tensor-core engine ----
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (cin == 3) {
    if (!(k == 3 && stride == 1 && cout == 32)) {
      set_error("cin=3 is only supported as convolutional1 (3->32, 3x3, stride 1)");
      cleanup();
      return DY_ERR_UNSUPPORTED;
    }
    float* d_w = nullptr;
    __nv_bfloat16* d_out = nullptr;
    if ((rc = talloc((void**)&d_w, (size_t)K * cout * 4)) ||
        (rc = talloc((void**)&d_out, p1_elems(B, H, W, cout) * 2))) { cleanup(); return rc; }
    cudaMemcpyAsync(d_w, w_host, (size_t)K * cout * 4, cudaMemcpyHostToDevice, st);
    note_launch(2);
    rc = launch_conv1(x_dev, d_w, d_sc, d_sh, alpha, B, H, W, nullptr, d_out, g_opt_conv1_tc != 0, num_sms, st);
    if (rc == DY_OK) rc = launch_p1_to_nhwc(d_out, out_dev, B, H, W, cout, FORM_SAME, st);
    cleanup();
    return rc;
  }
  if (cin % 32 != 0) {
    set_error("the bf16 engine needs cin % 32 == 0");
    cleanup();
    return DY_ERR_UNSUPPORTED;
  }
  __nv_bfloat16 *d_a = nullptr, *d_wpk = nullptr, *d_res = nullptr, *d_out = nullptr;
  const size_t a_elems = stride == 2 ? p1_elems(B, Ho, Wo, 4 * cin) : p1_elems(B, H, W, cin);
  std::vector<__nv_bfloat16> pk;
  pack_weights_bf16(w_host, K, cout, cout_pad, &pk);
  if ((rc = talloc((void**)&d_a, a_elems * 2)) || (rc = talloc((void**)&d_wpk, pk.size() * 2)) ||
      (rc = talloc((void**)&d_out, p1_elems(B, Ho, Wo, cout_pad) * 2))) { cleanup(); return rc; }
  cudaMemcpyAsync(d_wpk, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice, st);
  note_launch(3);
  rc = launch_nhwc_to_p1(x_dev, d_a, B, H, W, cin, stride == 2 ? FORM_S2D : FORM_SAME, st);
  if (rc == DY_OK && residual_dev) {
    if ((rc = talloc((void**)&d_res, p1_elems(B, Ho, Wo, cout) * 2)) == DY_OK)
      rc = launch_nhwc_to_p1(residual_dev, d_res, B, Ho, Wo, cout, FORM_SAME, st);
  }
  if (rc == DY_OK && cout % 16 != 0) {
    set_error("dy_conv_layer (bf16) needs cout % 16 == 0");
    rc = DY_ERR_UNSUPPORTED;
  }
  if (rc == DY_OK) {
    TcConvDesc c;
    c.a0 = d_a; c.cin0 = cin; c.cout = cout; c.k = k; c.s = stride; c.H = Ho; c.W = Wo; c.max_batch = B;
    c.wpk = d_wpk; c.cout_pad = cout_pad; c.scale = d_sc; c.shift = d_sh; c.act = act; c.alpha = alpha;
    c.residual = d_res;
    c.out[0] = OutDesc{d_out, OUT_SAME, cout};
    c.out[1] = OutDesc{nullptr, OUT_NONE, 0};
    TcPlan plan;
    rc = build_tc_plan(c, num_sms, &plan);
    if (rc == DY_OK) rc = run_tc_plan(plan, B, num_sms, st);
    if (rc == DY_OK) rc = launch_p1_to_nhwc(d_out, out_dev, B, Ho, Wo, cout, FORM_SAME, st);
  }
  cleanup();
  return rc;
}

int dy_conv_backward(const float* x_dev, const float* dz_dev, int32_t B, int32_t H, int32_t W, int32_t cin,
                     const float* w_host, int32_t k, int32_t stride, int32_t cout, float* dx_dev, float* dw_dev,
                     void* stream) {
  DY_CHECK(x_dev && dz_dev && w_host && (dx_dev || dw_dev), "null argument");
  DY_CHECK(k == 1 || k == 3, "k");
  DY_CHECK(stride == 1 || (stride == 2 && k == 3 && H % 2 == 0 && W % 2 == 0), "stride 2 needs a 3x3 kernel and even extents");
  DY_CHECK(cin % 32 == 0 && cin >= 32, "the bf16 engine needs cin % 32 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = DY_OK;
  std::vector<void*> tmp;
  auto talloc = [&](void** p, size_t bytes) -> int {
    DY_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    tmp.push_back(*p);
    DY_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, st));
    return DY_OK;
  };
  auto cleanup = [&]() {
    cudaStreamSynchronize(st);
    for (void* p : tmp) cudaFree(p);
  };
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  const int Ho = H / stride, Wo = W / stride;            // dz / GEMM row space
  const int Cg = (cout + 31) / 32 * 32, K = k * k * cin;
  const size_t rows_in = (size_t)B * (H + 1) * (W + 1) + 64, rows_out = (size_t)B * (Ho + 1) * (Wo + 1) + 64;
  __nv_bfloat16 *d_x = nullptr, *d_dz = nullptr, *d_dx = nullptr, *d_wdg = nullptr;
  float *d_w = nullptr, *d_one = nullptr, *d_zero = nullptr;
  const int npad = cin > 1024 ? cin : 1024;
  std::vector<float> ones(npad, 1.f);
  // x: P1 (stride 1) or the space-to-depth copy [B, Ho+1, Wo+1, 4*cin] (stride 2): what the forward consumed
  const size_t x_elems = stride == 2 ? rows_out * 4 * cin : rows_in * cin;
  if ((rc = talloc((void**)&d_x, x_elems * 2)) || (rc = talloc((void**)&d_dz, rows_out * Cg * 2)) ||
      (rc = talloc((void**)&d_dx, rows_in * cin * 2)) || (rc = talloc((void**)&d_wdg, (size_t)cin * k * k * Cg * 2)) ||
      (rc = talloc((void**)&d_w, (size_t)K * cout * 4)) || (rc = talloc((void**)&d_one, (size_t)npad * 4)) ||
      (rc = talloc((void**)&d_zero, (size_t)npad * 4))) {
    cleanup();
    return rc;
  }
  cudaMemcpyAsync(d_w, w_host, (size_t)K * cout * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(d_one, ones.data(), (size_t)npad * 4, cudaMemcpyHostToDevice, st);
  note_launch(2);
  rc = launch_nhwc_to_p1(x_dev, d_x, B, H, W, cin, stride == 2 ? FORM_S2D : FORM_SAME, st);
  if (rc == DY_OK) rc = launch_f32_to_p1(dz_dev, B, Ho, Wo, cout, d_dz, Cg, nullptr, st);
  if (rc == DY_OK && dx_dev && stride == 1) {
    note_launch(3);
    rc = launch_pack_dgrad_bf16(d_w, k, cin, 0, cin, cout, Cg, d_wdg, st);
    TcConvDesc c;
    c.a0 = d_dz; c.cin0 = Cg; c.cout = cin; c.k = k; c.s = 1; c.H = H; c.W = W; c.max_batch = B;
    c.wpk = d_wdg; c.cout_pad = cin; c.scale = d_one; c.shift = d_zero; c.act = 0; c.alpha = 0.1f;
    c.out[0] = OutDesc{d_dx, OUT_SAME, cin};
    c.out[1] = OutDesc{nullptr, OUT_NONE, 0};
    TcPlan plan;
    if (rc == DY_OK) rc = build_tc_plan(c, num_sms, &plan);
    if (rc == DY_OK) rc = run_tc_plan(plan, B, num_sms, st);
    if (rc == DY_OK) rc = launch_p1_to_nhwc(d_dx, dx_dev, B, H, W, cin, FORM_SAME, st);
  }
  if (rc == DY_OK && dx_dev && stride == 2) {
    // four parity-block GEMMs over the output row space, scattered into the P1 input gradient (train_init_tc)
    note_launch(6);
    rc = launch_pack_dgrad_s2_bf16(d_w, cin, cout, Cg, d_wdg, st);
    static const int kRegion[4] = {0, 4, 6, 8};
    for (int b = 0; b < 4 && rc == DY_OK; ++b) {
      const int by = b >> 1, bx = b & 1, nkw = bx == 0 ? 2 : 1, nkh = by == 0 ? 2 : 1;
      TcConvDesc c;
      c.a0 = d_dz; c.cin0 = Cg; c.cout = cin; c.k = 1; c.s = 1; c.H = Ho; c.W = Wo; c.max_batch = B;
      c.ntaps_custom = nkh * nkw;
      for (int tt = 0; tt < c.ntaps_custom; ++tt) {
        const int kh = by == 0 ? 2 * (tt / nkw) : 1, kw = bx == 0 ? 2 * (tt % nkw) : 1;
        c.tap_shift_custom[tt] = -((kh >> 1) * (Wo + 1) + (kw >> 1));
      }
      c.wpk = d_wdg + (size_t)kRegion[b] * cin * Cg;
      c.cout_pad = cin; c.scale = d_one; c.shift = d_zero; c.act = 0; c.alpha = 0.1f;
      c.out[0] = OutDesc{d_dx, OUT_UNS2D, cin, b};
      c.out[1] = OutDesc{nullptr, OUT_NONE, 0, 0};
      TcPlan plan;
      rc = build_tc_plan(c, num_sms, &plan);
      if (rc == DY_OK) rc = run_tc_plan(plan, B, num_sms, st);
    }
    if (rc == DY_OK) rc = launch_p1_to_nhwc(d_dx, dx_dev, B, H, W, cin, FORM_SAME, st);
  }
  if (rc == DY_OK && dw_dev) {
    note_launch();
    cudaMemsetAsync(dw_dev, 0, (size_t)K * cout * 4, st);
    WgradPlan wp;
    rc = build_wgrad_plan(d_x, cin, nullptr, 0, d_dz, Cg, cout, k, Ho, Wo, (long long)B * (Ho + 1) * (Wo + 1), dw_dev,
                          &wp, stride);
    if (rc == DY_OK) rc = run_wgrad_plan(wp, B, Ho, Wo, num_sms, st);
  }
  cleanup();
  return rc;
}

}  // extern "C"

// =============================================================================================
// Training step (fp32 engine).  Reference: sess.run([total_loss, optimizer]) --
// train_yolo3_mask.py:146-149,216; graph yolo3_net_pos.py:52-61 (placeholders, losses), :71-107 (BN),
// :631-860 (loss_yolo, loss_mask), :38/:61 (L2 + total), train_yolo3_mask.py:55 (Adam).
// =============================================================================================
namespace dy {

static const float kBnDecay = 0.997f;      // yolo3_net_pos.py:74
static const float kL2 = 1e-4f;            // yolo3_net_pos.py:38
static const float kAdamB1 = 0.9f, kAdamB2 = 0.999f, kAdamEps = 1e-8f;

static size_t out_elems(const LayerDef& d, int B) { return (size_t)B * d.H * d.H * d.cout; }
static int train_init_tc(dy_net* net);

static int train_init(dy_net* net) {
  if (net->train_ready) return DY_OK;
  DY_CHECK(net->finalized, "load weights before dy_train_init");
  const bool bf16 = net->cfg.precision == DY_PRECISION_BF16;
  auto& L = net->L;
  const int B = net->cfg.max_batch;
  // which layers are trainable, which take part in the backward pass
  for (int n = 1; n <= 82; ++n) L[n].unlocked = net->cfg.lock[n - 1] == 0;
  // a layer's output needs a gradient iff the layer or one of its ancestors is unlocked
  for (int n = 1; n <= 82; ++n) {
    const LayerDef& d = L[n].def;
    bool anc = false;
    if (d.src0 > 0) anc |= L[d.src0].in_bwd;
    if (d.src1 > 0) anc |= L[d.src1].in_bwd;
    if (d.res > 0) anc |= L[d.res].in_bwd;
    L[n].in_bwd = L[n].unlocked || anc;
    L[n].need_in_grad = anc;
  }
  long long off = 0;
  size_t max_out = 0, max_in = 0;
  for (int n = 1; n <= 82; ++n) {
    LayerState& s = L[n];
    const LayerDef& d = s.def;
    const int C = d.cout;
    // master copies of the per-channel variables on the device
    if (d.bn) {
      DY_TRY(dev_alloc(net, (void**)&s.d_gamma, C * 4)); DY_TRY(dev_alloc(net, (void**)&s.d_beta, C * 4));
      DY_TRY(dev_alloc(net, (void**)&s.d_mean, C * 4));  DY_TRY(dev_alloc(net, (void**)&s.d_var, C * 4));
      DY_CUDA(cudaMemcpy(s.d_gamma, s.gamma.data(), C * 4, cudaMemcpyHostToDevice));
      DY_CUDA(cudaMemcpy(s.d_beta, s.beta.data(), C * 4, cudaMemcpyHostToDevice));
      DY_CUDA(cudaMemcpy(s.d_mean, s.mean.data(), C * 4, cudaMemcpyHostToDevice));
      DY_CUDA(cudaMemcpy(s.d_var, s.var.data(), C * 4, cudaMemcpyHostToDevice));
    } else {
      DY_TRY(dev_alloc(net, (void**)&s.d_bias, C * 4));
      DY_CUDA(cudaMemcpy(s.d_bias, s.bias.data(), C * 4, cudaMemcpyHostToDevice));
    }
    if (!s.in_bwd) continue;
    if (bf16) {
      // tensor-core engine: any lock pattern -- stride-1 convs over P1 activations, stride-2 convs over the
      // space-to-depth copies, convolutional1 (K = 27) with a CUDA-core weight gradient
      s.Cg = (C + 31) / 32 * 32;
      const size_t rows_pad = (size_t)B * (d.H + 1) * (d.H + 1) + 64;
      if (d.bn) DY_TRY(dev_alloc(net, (void**)&s.zb, rows_pad * C * 2));
      else DY_TRY(dev_alloc(net, (void**)&s.dyf, out_elems(d, B) * 4));
      DY_TRY(dev_alloc(net, (void**)&s.dyb, rows_pad * s.Cg * 2));
      if (!s.d_w_f32) {                      // fp32 master copy of the weights
        DY_TRY(dev_alloc(net, (void**)&s.d_w_f32, (size_t)s.K * C * 4));
        DY_CUDA(cudaMemcpy(s.d_w_f32, s.w.data(), (size_t)s.K * C * 4, cudaMemcpyHostToDevice));
      }
    } else {
      DY_TRY(dev_alloc(net, (void**)&s.z, out_elems(d, B) * 4));
      DY_TRY(dev_alloc(net, (void**)&s.dy, out_elems(d, B) * 4));
    }
    DY_TRY(dev_alloc(net, (void**)&s.stat, (size_t)4 * C * 8));
    DY_TRY(dev_alloc(net, (void**)&s.bn_a, C * 4)); DY_TRY(dev_alloc(net, (void**)&s.bn_b, C * 4));
    DY_TRY(dev_alloc(net, (void**)&s.bmean, C * 4)); DY_TRY(dev_alloc(net, (void**)&s.bvar, C * 4));
    DY_TRY(dev_alloc(net, (void**)&s.binvstd, C * 4));
    if (s.need_in_grad && !bf16) DY_TRY(dev_alloc(net, (void**)&s.wt, (size_t)s.K * C * 4));
    if (out_elems(d, B) > max_out) max_out = out_elems(d, B);
    const size_t in_e = (size_t)B * (d.H * d.s) * (d.H * d.s) * (d.cin0 + d.cin1);
    if (s.need_in_grad && in_e > max_in) max_in = in_e;
    if (s.unlocked) {
      s.off_w = off; off += (long long)s.K * C;
      if (d.bn) { s.off_g = off; off += C; s.off_b = off; off += C; }
      else { s.off_b = off; off += C; }
    }
  }
  net->n_train = off;
  DY_TRY(dev_alloc(net, (void**)&net->adam_m, (size_t)off * 4));
  DY_TRY(dev_alloc(net, (void**)&net->adam_v, (size_t)off * 4));
  if (!bf16) {
    DY_TRY(dev_alloc(net, (void**)&net->dz_scratch, max_out * 4));
    DY_TRY(dev_alloc(net, (void**)&net->dx_scratch, max_in * 4));
  }
  DY_TRY(dev_alloc(net, (void**)&net->loss_acc, 8 * 8));
  DY_TRY(dev_alloc(net, (void**)&net->mask_rois, (size_t)B * 10 * 4 * 4));
  DY_TRY(dev_alloc(net, (void**)&net->mask_assign, (size_t)B * 10 * 4));
  DY_TRY(dev_alloc(net, (void**)&net->mask_npos, (size_t)B * 4));
  DY_TRY(dev_alloc(net, (void**)&net->train_windows, (size_t)B * 16));
  DY_TRY(dev_alloc(net, (void**)&net->ones_dev, 1024 * 4));
  DY_TRY(dev_alloc(net, (void**)&net->zeros_dev, 1024 * 4));
  {
    std::vector<float> ones(1024, 1.f);
    DY_CUDA(cudaMemcpy(net->ones_dev, ones.data(), 1024 * 4, cudaMemcpyHostToDevice));
  }
  std::vector<float> w(B * 4);
  for (int b = 0; b < B; ++b) { w[4 * b] = 0; w[4 * b + 1] = 0; w[4 * b + 2] = 1; w[4 * b + 3] = 1; }
  DY_CUDA(cudaMemcpy(net->train_windows, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  if (bf16) DY_TRY(train_init_tc(net));
  // segment tables of the multi-tensor kernels
  std::vector<ParamSeg> ps;
  std::vector<BnSeg> bs;
  std::vector<PackSeg> ks;
  for (int n = 1; n <= 82; ++n) {
    LayerState& s = L[n];
    const LayerDef& d = s.def;
    if (!s.unlocked) continue;
    const int C = d.cout;
    ps.push_back(ParamSeg{s.d_w_f32, s.off_w, (long long)s.K * C, kL2, 0});
    if (d.bn) {
      ps.push_back(ParamSeg{s.d_gamma, s.off_g, C, 0.f, 0});
      ps.push_back(ParamSeg{s.d_beta, s.off_b, C, 0.f, 0});
      bs.push_back(BnSeg{s.d_mean, s.d_var, s.bmean, s.bvar, s.d_gamma, s.d_beta, s.d_scale, s.d_shift, C, 0});
    } else {
      ps.push_back(ParamSeg{s.d_bias, s.off_b, C, kL2, 0});
    }
    if (bf16) {
      if (n == 1) continue;               // the stem kernel packs its fp32 weights itself
      // (a stride-2 conv's dgrad operands have their own layout: re-packed by repack_s2 after every step)
      ks.push_back(PackSeg{s.d_w_f32, s.d_wpk, d.s == 2 ? nullptr : s.wdg0, s.wdg1, s.K, C, s.cout_pad, d.k, d.cin0,
                           d.cin1, s.Cg, 0});
      const int tiles = ((s.K + 31) / 32) * ((s.cout_pad + 31) / 32);
      if (tiles > net->max_pack_tiles) net->max_pack_tiles = tiles;
    }
  }
  net->n_pseg = (int)ps.size(); net->n_bseg = (int)bs.size(); net->n_kseg = (int)ks.size();
  DY_TRY(dev_alloc(net, (void**)&net->pseg_dev, ps.size() * sizeof(ParamSeg)));
  DY_TRY(dev_alloc(net, (void**)&net->bseg_dev, bs.size() * sizeof(BnSeg)));
  DY_TRY(dev_alloc(net, (void**)&net->kseg_dev, ks.size() * sizeof(PackSeg)));
  if (!ps.empty()) DY_CUDA(cudaMemcpy(net->pseg_dev, ps.data(), ps.size() * sizeof(ParamSeg), cudaMemcpyHostToDevice));
  if (!bs.empty()) DY_CUDA(cudaMemcpy(net->bseg_dev, bs.data(), bs.size() * sizeof(BnSeg), cudaMemcpyHostToDevice));
  if (!ks.empty()) DY_CUDA(cudaMemcpy(net->kseg_dev, ks.data(), ks.size() * sizeof(PackSeg), cudaMemcpyHostToDevice));
  net->train_ready = true;
  return DY_OK;
}

static ConvGeom geom_of(const LayerDef& d, int B) {
  ConvGeom g;
  g.B = B; g.Ho = g.Wo = d.H; g.Hi = g.Wi = d.H * d.s; g.cin = d.cin0 + d.cin1; g.cout = d.cout; g.k = d.k; g.s = d.s;
  tf_same_pad(g.Hi, d.k, d.s, &g.pad_t);
  g.pad_l = g.pad_t;
  return g;
}

// training-mode forward of one layer: batch-stat BN for unlocked layers (yolo3_net_pos.py:88-98),
// moving statistics for locked ones (:76-81)
static int train_forward_layer(dy_net* net, int n, const float* images, int B, cudaStream_t st) {
  auto& L = net->L;
  LayerState& s = L[n];
  const LayerDef& d = s.def;
  RefConvArgs a;
  memset(&a, 0, sizeof(a));
  a.src0 = d.src0 == 0 ? images : L[d.src0].f32;
  a.src1 = d.src1 > 0 ? L[d.src1].f32 : nullptr;
  a.w = s.d_w_f32;
  a.Ho = a.Wo = d.H; a.Hi = a.Wi = d.H * d.s;
  a.c0 = d.cin0; a.c1 = d.cin1; a.cout = d.cout; a.k = d.k; a.s = d.s;
  tf_same_pad(a.Hi, d.k, d.s, &a.pad_t);
  a.pad_l = a.pad_t;
  a.alpha = net->cfg.alpha;
  const float* res = d.res > 0 ? L[d.res].f32 : nullptr;
  const long long M = (long long)B * d.H * d.H;
  if (!s.in_bwd) {            // frozen prefix: identical to inference
    a.scale = s.d_scale; a.shift = s.d_shift; a.residual = res; a.out = s.f32; a.act = d.bn ? 1 : 0;
    note_launch();
    return launch_conv_ref(a, B, st);
  }
  if (!d.bn) {                // biased linear conv: y = conv + bias (scale = 1)
    a.scale = net->ones_dev; a.shift = s.d_bias; a.residual = nullptr; a.out = s.f32; a.act = 0;
    note_launch();
    return launch_conv_ref(a, B, st);
  }
  // z = conv(x) with an identity epilogue (scale 1, shift 0, no activation)
  a.scale = net->ones_dev; a.shift = net->zeros_dev; a.residual = nullptr; a.out = s.z; a.act = 0;
  note_launch(3);
  DY_TRY(launch_conv_ref(a, B, st));
  if (s.unlocked) {
    DY_CUDA(cudaMemsetAsync(s.stat, 0, (size_t)4 * d.cout * 8, st));
    DY_TRY(launch_bn_stats(s.z, M, d.cout, s.stat, s.stat + d.cout, st));
    DY_TRY(launch_bn_finalize(s.stat, s.stat + d.cout, M, d.cout, s.d_gamma, s.d_beta, net->cfg.bn_eps, s.bn_a,
                              s.bn_b, s.bmean, s.bvar, s.binvstd, st));
  } else {
    DY_TRY(launch_refold(s.d_gamma, s.d_beta, s.d_mean, s.d_var, net->cfg.bn_eps, d.cout, s.bn_a, s.bn_b, st));
  }
  return launch_bn_act(s.z, s.bn_a, s.bn_b, d.cout, M * d.cout, net->cfg.alpha, 1, res, s.f32, st);
}

static int train_backward_layer(dy_net* net, int n, int B, float* grad_flat, cudaStream_t st) {
  auto& L = net->L;
  LayerState& s = L[n];
  const LayerDef& d = s.def;
  if (!s.in_bwd) return DY_OK;
  const long long M = (long long)B * d.H * d.H;
  const int C = d.cout;
  float* dz = net->dz_scratch;
  // residual shortcut: y = leaky(bn(conv)) + shortcut  ->  d shortcut += dy
  if (d.res > 0 && L[d.res].in_bwd) { note_launch(); DY_TRY(launch_add(L[d.res].dy, s.dy, M * C, st)); }
  double* s1 = s.stat + 2 * C;
  double* s2 = s.stat + 3 * C;
  if (d.bn) {
    if (s.unlocked) {
      DY_CUDA(cudaMemsetAsync(s1, 0, (size_t)2 * C * 8, st));
      DY_TRY(launch_bn_bwd_reduce(s.dy, s.z, s.bn_a, s.bn_b, s.bmean, s.binvstd, net->cfg.alpha, 1, M, C, s1, s2, st));
      DY_TRY(launch_copy_stats_to_grads(s1, s2, C, grad_flat + s.off_g, grad_flat + s.off_b, st));
      DY_TRY(launch_bn_bwd_apply(s.dy, s.z, s.bn_a, s.bn_b, s.bmean, s.binvstd, s.d_gamma, s1, s2, net->cfg.alpha, 1,
                                 0, M, C, dz, st));
    } else {
      DY_TRY(launch_bn_bwd_apply(s.dy, s.z, s.bn_a, s.bn_b, nullptr, nullptr, nullptr, nullptr, nullptr,
                                 net->cfg.alpha, 1, 1, M, C, dz, st));
    }
    note_launch(3);
  } else {
    dz = s.dy;       // dz = dy
    if (s.unlocked) {
      DY_CUDA(cudaMemsetAsync(s1, 0, (size_t)C * 8, st));
      DY_TRY(launch_bn_stats(s.dy, M, C, s1, nullptr, st));                       // d bias = column sums
      DY_TRY(launch_copy_stats_to_grads(s1, s1, C, nullptr, grad_flat + s.off_b, st));
      note_launch(2);
    }
  }
  const ConvGeom g = geom_of(d, B);
  if (s.unlocked) {
    DY_CUDA(cudaMemsetAsync(grad_flat + s.off_w, 0, (size_t)s.K * C * 4, st));
    const float* x0 = d.src0 == 0 ? nullptr : L[d.src0].f32;
    DY_CHECK(x0 != nullptr, "convolutional1 cannot be trained from here (image gradient path)");
    note_launch();
    DY_TRY(launch_conv_wgrad(x0, d.cin0, d.src1 > 0 ? L[d.src1].f32 : nullptr, d.cin1, dz, grad_flat + s.off_w, g,
                             net->num_sms, st));
  }
  if (s.need_in_grad) {
    note_launch(3);
    DY_TRY(launch_weight_transpose(s.d_w_f32, s.wt, d.k * d.k, g.cin, C, st));
    float* d0 = (d.src0 > 0 && L[d.src0].in_bwd) ? L[d.src0].dy : nullptr;
    float* d1 = (d.src1 > 0 && L[d.src1].in_bwd) ? L[d.src1].dy : nullptr;
    if (d.src1 == 0 && d0) {
      DY_TRY(launch_conv_dgrad(dz, s.wt, d0, g, 1, st));                         // accumulate in place
    } else if (d0 || d1) {
      DY_TRY(launch_conv_dgrad(dz, s.wt, net->dx_scratch, g, 0, st));
      DY_TRY(launch_split_accumulate(net->dx_scratch, B, g.Hi, g.Wi, d.cin0, d.cin1, d0, d1, st));
    }
  }
  return DY_OK;
}


// ---------------------------------------------------------------------------------------------
// Training step on the tensor-core engine (precision = bf16): bf16 operands and activations, fp32
// accumulation, fp32 master weights / BN statistics / gradients / Adam.  See train_tc.cuh.
// ---------------------------------------------------------------------------------------------
static int repack_tc(dy_net* net, int n, cudaStream_t st) {
  LayerState& s = net->L[n];
  const LayerDef& d = s.def;
  if (s.unlocked && n != 1) {
    note_launch();
    DY_TRY(launch_pack_fwd_bf16(s.d_w_f32, s.K, d.cout, s.cout_pad, s.d_wpk, st));
  }
  const float* w = s.d_w_f32;
  if (s.wdg0 && d.s == 2) {
    note_launch();
    DY_TRY(launch_pack_dgrad_s2_bf16(w, d.cin0, d.cout, s.Cg, s.wdg0, st));
  } else if (s.wdg0) {
    note_launch();
    DY_TRY(launch_pack_dgrad_bf16(w, d.k, d.cin0 + d.cin1, 0, d.cin0, d.cout, s.Cg, s.wdg0, st));
  }
  if (s.wdg1) {
    note_launch();
    DY_TRY(launch_pack_dgrad_bf16(w, d.k, d.cin0 + d.cin1, d.cin0, d.cin1, d.cout, s.Cg, s.wdg1, st));
  }
  return DY_OK;
}

static int train_init_tc(dy_net* net) {
  auto& L = net->L;
  const int B = net->cfg.max_batch;
  // scratch: dz of the layer being differentiated, and its 2x2-pooled copy for concat consumers
  size_t max_dz = 0, max_pool = 0;
  for (int n = 1; n <= 82; ++n) {
    const LayerState& s = L[n];
    const LayerDef& d = s.def;
    if (!s.in_bwd) continue;
    const size_t rows_pad = (size_t)B * (d.H + 1) * (d.H + 1) + 64;
    if (d.bn && rows_pad * d.cout > max_dz) max_dz = rows_pad * d.cout;
    if (d.src1 > 0 && L[d.src1].in_bwd) {
      const size_t rp = (size_t)B * (d.H / 2 + 1) * (d.H / 2 + 1) + 64;
      if (rp * s.Cg > max_pool) max_pool = rp * s.Cg;
    }
  }
  DY_TRY(dev_alloc(net, (void**)&net->dzb_scratch, max_dz * 2));
  DY_TRY(dev_alloc(net, (void**)&net->pool_scratch, max_pool * 2));
  // Backward visits layers 82 -> 1; inside a layer the order is shortcut, src0, src1.  The first
  // writer of a gradient buffer overwrites it, every later one accumulates through the conv
  // kernel's residual input (in place: each tile reads its residual before it stores).
  std::vector<char> written(83, 0);
  for (int n = 82; n >= 1; --n) {
    LayerState& s = L[n];
    const LayerDef& d = s.def;
    if (!s.in_bwd) continue;
    const long long rows_max = (long long)B * (d.H + 1) * (d.H + 1);
    const __nv_bfloat16* dz = d.bn ? net->dzb_scratch : s.dyb;
    if (d.res > 0 && L[d.res].in_bwd) {
      s.res_acc = written[d.res] != 0;
      written[d.res] = 1;
    }
    if (d.bn && n != 1) {      // forward: z = conv(x), identity epilogue (convolutional1: the stem kernel)
      TcConvDesc c;
      c.a0 = d.s == 2 ? L[d.src0].s2d : L[d.src0].same;
      c.a1 = d.src1 > 0 ? L[d.src1].up : nullptr;
      DY_CHECK(c.a0 != nullptr && (d.src1 == 0 || c.a1 != nullptr), "layer input buffer missing");
      c.cin0 = d.cin0; c.cin1 = d.cin1; c.cout = d.cout; c.k = d.k; c.s = d.s; c.H = c.W = d.H; c.max_batch = B;
      c.wpk = s.d_wpk; c.cout_pad = s.cout_pad; c.scale = net->ones_dev; c.shift = net->zeros_dev; c.act = 0;
      c.alpha = net->cfg.alpha;
      c.out[0] = OutDesc{s.zb, OUT_SAME, d.cout};
      c.out[1] = OutDesc{nullptr, OUT_NONE, 0};
      DY_TRY(build_tc_plan(c, net->num_sms, &s.plan_z));
    }
    if (s.unlocked && n != 1) {
      DY_TRY(build_wgrad_plan(d.s == 2 ? L[d.src0].s2d : L[d.src0].same, d.cin0, d.src1 > 0 ? L[d.src1].up : nullptr,
                              d.cin1, dz, s.Cg, d.cout, d.k, d.H, d.H, rows_max, nullptr, &s.plan_wg, d.s));
    }
    if (d.src0 > 0 && L[d.src0].in_bwd && d.s == 2) {
      // dgrad of a stride-2 conv: one GEMM per parity block of the space-to-depth input over the OUTPUT row space,
      // A = dz shifted back by the tap's (kh>>1, kw>>1), output scattered to pixel (2y + by, 2x + bx) of the
      // producer's P1 gradient (OUT_UNS2D); a producer with an earlier consumer accumulates
      LayerState& t = L[d.src0];
      DY_CHECK(t.Cg == d.cin0, "producer gradient width");
      s.dg_s2 = true;
      DY_TRY(dev_alloc(net, (void**)&s.wdg0, (size_t)d.cin0 * 9 * s.Cg * 2));
      const int Wp = d.H + 1;
      static const int kRegion[4] = {0, 4, 6, 8};
      for (int b = 0; b < 4; ++b) {
        const int by = b >> 1, bx = b & 1, nkw = bx == 0 ? 2 : 1, nkh = by == 0 ? 2 : 1;
        TcConvDesc c;
        c.a0 = dz; c.cin0 = s.Cg; c.cout = d.cin0; c.k = 1; c.s = 1; c.H = c.W = d.H; c.max_batch = B;
        c.ntaps_custom = nkh * nkw;
        for (int tt = 0; tt < c.ntaps_custom; ++tt) {
          const int kh = by == 0 ? 2 * (tt / nkw) : 1, kw = bx == 0 ? 2 * (tt % nkw) : 1;
          c.tap_shift_custom[tt] = -((kh >> 1) * Wp + (kw >> 1));
        }
        c.wpk = s.wdg0 + (size_t)kRegion[b] * d.cin0 * s.Cg;
        c.cout_pad = d.cin0; c.scale = net->ones_dev; c.shift = net->zeros_dev; c.act = 0; c.alpha = net->cfg.alpha;
        c.out[0] = OutDesc{t.dyb, written[d.src0] ? OUT_UNS2D_ACC : OUT_UNS2D, d.cin0, b};
        c.out[1] = OutDesc{nullptr, OUT_NONE, 0, 0};
        DY_TRY(build_tc_plan(c, net->num_sms, &s.plan_dg_s2[b]));
      }
      written[d.src0] = 1;
    } else if (d.src0 > 0 && L[d.src0].in_bwd) {
      LayerState& t = L[d.src0];
      DY_CHECK(t.Cg == d.cin0, "producer gradient width");
      s.dg0 = true;
      DY_TRY(dev_alloc(net, (void**)&s.wdg0, (size_t)d.cin0 * d.k * d.k * s.Cg * 2));
      TcConvDesc c;
      c.a0 = dz; c.cin0 = s.Cg; c.cout = d.cin0; c.k = d.k; c.s = 1; c.H = c.W = d.H; c.max_batch = B;
      c.wpk = s.wdg0; c.cout_pad = d.cin0; c.scale = net->ones_dev; c.shift = net->zeros_dev; c.act = 0;
      c.alpha = net->cfg.alpha;
      c.residual = written[d.src0] ? t.dyb : nullptr;
      c.out[0] = OutDesc{t.dyb, OUT_SAME, d.cin0};
      c.out[1] = OutDesc{nullptr, OUT_NONE, 0};
      DY_TRY(build_tc_plan(c, net->num_sms, &s.plan_dg0));
      written[d.src0] = 1;
    }
    if (d.src1 > 0 && L[d.src1].in_bwd) {
      LayerState& t = L[d.src1];
      DY_CHECK(t.Cg == d.cin1 && d.k == 1, "concat branch gradient width");
      s.dg1 = true;
      DY_TRY(dev_alloc(net, (void**)&s.wdg1, (size_t)d.cin1 * s.Cg * 2));
      TcConvDesc c;      // 1x1 dgrad at the LOW resolution on the 2x2-pooled dz (pooling commutes with a 1x1 conv)
      c.a0 = net->pool_scratch; c.cin0 = s.Cg; c.cout = d.cin1; c.k = 1; c.s = 1; c.H = c.W = d.H / 2; c.max_batch = B;
      c.wpk = s.wdg1; c.cout_pad = d.cin1; c.scale = net->ones_dev; c.shift = net->zeros_dev; c.act = 0;
      c.alpha = net->cfg.alpha;
      c.residual = written[d.src1] ? t.dyb : nullptr;
      c.out[0] = OutDesc{t.dyb, OUT_SAME, d.cin1};
      c.out[1] = OutDesc{nullptr, OUT_NONE, 0};
      DY_TRY(build_tc_plan(c, net->num_sms, &s.plan_dg1));
      written[d.src1] = 1;
    }
    DY_TRY(repack_tc(net, n, 0));
  }
  DY_CUDA(cudaDeviceSynchronize());
  return DY_OK;
}

// dy_finalize_weights after dy_train_init: host copies <-> device masters (see dy_finalize_weights)
static int sync_masters(dy_net* net, bool* any_new) {
  *any_new = false;
  for (int n = 1; n <= 82; ++n) {
    LayerState& s = net->L[n];
    const size_t C = (size_t)s.def.cout, wn = (size_t)s.K * C;
    struct Var { unsigned bit; std::vector<float>* host; float* dev; size_t cnt; };
    const Var vars[6] = {{1u << 0, &s.w, s.d_w_f32, wn},     {1u << 1, &s.gamma, s.d_gamma, C}, {1u << 2, &s.beta, s.d_beta, C},
                         {1u << 3, &s.mean, s.d_mean, C},    {1u << 4, &s.var, s.d_var, C},     {1u << 5, &s.bias, s.d_bias, C}};
    for (const Var& v : vars) {
      if (v.dev == nullptr || v.host->size() != v.cnt) continue;
      if (s.host_new & v.bit) {
        DY_CUDA(cudaMemcpy(v.dev, v.host->data(), v.cnt * 4, cudaMemcpyHostToDevice));
        *any_new = true;
      } else if (s.unlocked) {
        DY_CUDA(cudaMemcpy(v.host->data(), v.dev, v.cnt * 4, cudaMemcpyDeviceToHost));
      }
    }
  }
  return DY_OK;
}

// ... and afterwards: the tensor-core engine's dgrad operands follow the (possibly restored) masters; a restore
// starts the optimizer afresh (the reference's Saver does not save Adam slots, train_yolo3_mask.py:47-58)
static int retrain_refresh(dy_net* net, bool any_new) {
  if (net->cfg.precision == DY_PRECISION_BF16)
    for (int n = 1; n <= 82; ++n)
      if (net->L[n].in_bwd) DY_TRY(repack_tc(net, n, 0));
  if (any_new && net->n_train > 0) {
    DY_CUDA(cudaMemset(net->adam_m, 0, (size_t)net->n_train * 4));
    DY_CUDA(cudaMemset(net->adam_v, 0, (size_t)net->n_train * 4));
    net->adam_step = 0;
  }
  return DY_OK;
}

static int train_forward_layer_tc(dy_net* net, int n, const float* images, int B, cudaStream_t st) {
  auto& L = net->L;
  LayerState& s = L[n];
  const LayerDef& d = s.def;
  if (n == 1 && !s.in_bwd) {
    note_launch();
    return launch_conv1(images, s.d_w_f32, s.d_scale, s.d_shift, net->cfg.alpha, B, net->S, net->S, s.s2d, s.same,
                        g_opt_conv1_tc != 0, net->num_sms, st);
  }
  // frozen prefix and the biased linear convs (bias lives in d_shift): the inference plan
  if (n != 1 && (!s.in_bwd || !d.bn)) return run_tc_plan(s.plan, B, net->num_sms, st);
  if (n == 1) {
    // trained stem: z = conv(x) through the same kernel with an identity epilogue (scale 1, shift 0, slope 1)
    note_launch();
    DY_TRY(launch_conv1(images, s.d_w_f32, net->ones_dev, net->zeros_dev, 1.0f, B, net->S, net->S, nullptr, s.zb,
                        g_opt_conv1_tc != 0, net->num_sms, st));
  } else {
    DY_TRY(run_tc_plan(s.plan_z, B, net->num_sms, st));
  }
  const long long rows = (long long)B * (d.H + 1) * (d.H + 1);
  const long long M = (long long)B * d.H * d.H;
  note_launch(2);
  const __nv_bfloat16* res = d.res > 0 ? L[d.res].same : nullptr;
  if (s.unlocked) {
    DY_CUDA(cudaMemsetAsync(s.stat, 0, (size_t)4 * d.cout * 8, st));
    DY_TRY(launch_bn_stats_p1(s.zb, rows, d.cout, s.stat, s.stat + d.cout, st));
    return launch_bn_finalize_act_p1(s.zb, s.stat, s.stat + d.cout, M, s.d_gamma, s.d_beta, net->cfg.bn_eps, s.bn_a,
                                     s.bn_b, s.bmean, s.bvar, s.binvstd, res, B, d.H, d.H, d.cout, net->cfg.alpha, 1,
                                     s.same, s.up, st, s.s2d);
  }
  DY_TRY(launch_refold(s.d_gamma, s.d_beta, s.d_mean, s.d_var, net->cfg.bn_eps, d.cout, s.bn_a, s.bn_b, st));
  return launch_bn_act_p1(s.zb, s.bn_a, s.bn_b, res, B, d.H, d.H, d.cout, net->cfg.alpha, 1, s.same, s.up, st, s.s2d);
}

static int train_backward_layer_tc(dy_net* net, int n, int B, float* grad_flat, cudaStream_t st) {
  auto& L = net->L;
  LayerState& s = L[n];
  const LayerDef& d = s.def;
  if (!s.in_bwd) return DY_OK;
  const long long rows = (long long)B * (d.H + 1) * (d.H + 1);
  const int C = d.cout;
  double* s1 = s.stat + 2 * C;
  double* s2 = s.stat + 3 * C;
  if (d.res > 0 && L[d.res].in_bwd) {      // y = leaky(bn(conv)) + shortcut  ->  d shortcut (+)= dy
    note_launch();
    if (s.res_acc) DY_TRY(launch_add_p1(L[d.res].dyb, s.dyb, rows * C, st));
    else DY_TRY(launch_copy_p1(L[d.res].dyb, s.dyb, rows * C, st));
  }
  if (d.bn) {
    if (s.unlocked) {
      DY_CUDA(cudaMemsetAsync(s1, 0, (size_t)2 * C * 8, st));
      DY_TRY(launch_bn_bwd_reduce_p1(s.dyb, s.zb, s.bn_a, s.bn_b, s.bmean, s.binvstd, net->cfg.alpha, 1, rows, C, s1,
                                     s2, st));
      DY_TRY(launch_bn_bwd_apply_p1(s.dyb, s.zb, s.bn_a, s.bn_b, s.bmean, s.binvstd, s.d_gamma, s1, s2,
                                    net->cfg.alpha, 1, 0, B, d.H, d.H, C, net->dzb_scratch, grad_flat + s.off_g,
                                    grad_flat + s.off_b, st));
    } else {
      DY_TRY(launch_bn_bwd_apply_p1(s.dyb, s.zb, s.bn_a, s.bn_b, nullptr, nullptr, nullptr, nullptr, nullptr,
                                    net->cfg.alpha, 1, 1, B, d.H, d.H, C, net->dzb_scratch, nullptr, nullptr, st));
    }
    note_launch(2);
  } else {
    // biased linear conv: dz = dy (the loss kernels wrote it in fp32 NHWC); d bias = column sums
    DY_TRY(launch_f32_to_p1(s.dyf, B, d.H, d.H, C, s.dyb, s.Cg, s.unlocked ? grad_flat + s.off_b : nullptr, st));
    note_launch();
  }
  if (s.unlocked && n == 1) {
    DY_CHECK(net->train_images != nullptr, "dy_train_forward has not run");
    note_launch();
    DY_TRY(launch_conv1_wgrad(net->train_images, net->dzb_scratch, B, d.H, d.H, grad_flat + s.off_w, st));
  } else if (s.unlocked) {
    DY_CUDA(cudaMemsetAsync(grad_flat + s.off_w, 0, (size_t)s.K * C * 4, st));
    s.plan_wg.p.dw = grad_flat + s.off_w;
    note_launch();
    DY_TRY(run_wgrad_plan(s.plan_wg, B, d.H, d.H, net->num_sms, st));
  }
  if (s.dg_s2)
    for (int b = 0; b < 4; ++b) DY_TRY(run_tc_plan(s.plan_dg_s2[b], B, net->num_sms, st));
  if (s.dg0) DY_TRY(run_tc_plan(s.plan_dg0, B, net->num_sms, st));
  if (s.dg1) {
    note_launch();
    DY_TRY(launch_pool2x2_p1(d.bn ? net->dzb_scratch : s.dyb, B, d.H / 2, d.H / 2, s.Cg, net->pool_scratch, st));
    DY_TRY(run_tc_plan(s.plan_dg1, B, net->num_sms, st));
  }
  return DY_OK;
}

}  // namespace dy

extern "C" {

int dy_train_init(dy_net* net) {
  DY_CHECK(net != nullptr, "null net");
  DY_CUDA(cudaSetDevice(net->cfg.device));
  return train_init(net);
}

int64_t dy_train_param_count(dy_net* net) { return (net && net->train_ready) ? net->n_train : -1; }

int dy_train_layer_span(dy_net* net, int32_t layer, int64_t* offset, int64_t* count) {
  DY_CHECK(net && net->train_ready && layer >= 1 && layer <= 82 && offset && count, "bad argument");
  const LayerState& s = net->L[layer];
  if (!s.unlocked) { *offset = -1; *count = 0; return DY_OK; }
  *offset = s.off_w;
  *count = (long long)s.K * s.def.cout + (s.def.bn ? 2 : 1) * s.def.cout;
  return DY_OK;
}

int dy_train_forward(dy_net* net, const float* images_dev, int32_t B, const float* yolo3_dev, const float* yolo2_dev,
                     const float* yolo1_dev, const float* true_boxes_dev, const uint8_t* true_masks_dev,
                     const int32_t* perm_prop_dev, const int32_t* perm_gt_dev, float det_thresh, float* losses_host,
                     void* stream) {
  NvtxRange nvtx_("dy_train_forward");
  DY_CHECK(net && images_dev && yolo3_dev && yolo2_dev && yolo1_dev && true_boxes_dev && true_masks_dev &&
               perm_prop_dev && perm_gt_dev && losses_host, "null argument");
  DY_CHECK(net->train_ready, "dy_train_init has not been called");
  DY_CHECK(B >= 1 && B <= net->cfg.max_batch, "batch exceeds max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  DY_TRY(user_begin(net, st));
  net->last_fused = false;
  auto& L = net->L;
  const bool tc = net->cfg.precision == DY_PRECISION_BF16;
  net->train_images = images_dev;
  if (tc) {
    for (int n = 1; n <= 82; ++n) DY_TRY(train_forward_layer_tc(net, n, images_dev, B, st));
    // only the loss kernels' fp32 targets need clearing: every bf16 gradient buffer is fully rewritten
    for (int n = 1; n <= 82; ++n)
      if (L[n].dyf) DY_CUDA(cudaMemsetAsync(L[n].dyf, 0, out_elems(L[n].def, B) * 4, st));
  } else {
    for (int n = 1; n <= 82; ++n) DY_TRY(train_forward_layer(net, n, images_dev, B, st));
    for (int n = 1; n <= 82; ++n)
      if (L[n].in_bwd) DY_CUDA(cudaMemsetAsync(L[n].dy, 0, out_elems(L[n].def, B) * 4, st));
  }
  DY_CUDA(cudaMemsetAsync(net->loss_acc, 0, 8 * 8, st));
  // detection loss + gradient into the three head maps
  YoloLossArgs ya;
  memset(&ya, 0, sizeof(ya));
  const int heads[3] = {75, 67, 59};
  const float* labs[3] = {yolo3_dev, yolo2_dev, yolo1_dev};
  for (int j = 0; j < 3; ++j) {
    DY_CHECK(L[heads[j]].in_bwd, "detection heads must be trainable");
    ya.pred[j] = L[heads[j]].f32; ya.label[j] = labs[j]; ya.dpred[j] = tc ? L[heads[j]].dyf : L[heads[j]].dy;
    ya.g[j] = L[heads[j]].def.H;
  }
  ya.B = B; ya.net = 32 * ya.g[2];
  memcpy(ya.anchors, net->cfg.anchors, sizeof(ya.anchors));
  ya.true_boxes = true_boxes_dev;
  ya.ignore_thresh = net->ignore_thresh; ya.object_scale = net->object_scale; ya.noobject_scale = net->noobject_scale;
  ya.class_scale = net->class_scale; ya.coord_scale = net->coord_scale;
  ya.loss = net->loss_acc;
  note_launch();
  DY_TRY(launch_yolo_loss(ya, st));
  // proposals for the mask loss: filter_detections on the training-mode predictions (:357-359)
  DY_TRY(run_detect(net, L[75].f32, L[67].f32, L[59].f32, B, net->train_windows, det_thresh, nullptr, nullptr, nullptr,
                    net->det_raw_ws, net->det_box_ws, net->det_count_ws, st));
  MaskLossArgs ma;
  memset(&ma, 0, sizeof(ma));
  DY_CHECK(L[82].in_bwd, "the mask subnet must be trainable");
  ma.det = net->det_raw_ws; ma.true_boxes = true_boxes_dev; ma.true_masks = true_masks_dev;
  ma.perm_prop = perm_prop_dev; ma.perm_gt = perm_gt_dev; ma.mask_pos = L[82].f32;
  ma.dmask = tc ? L[82].dyf : L[82].dy;
  ma.mp_planar = tc ? 1 : 0;           // the bf16 engine keeps the score maps planar [B,kk,S,S]
  ma.B = B; ma.max_det = net->cfg.max_detection; ma.S = net->S / 2; ma.H = net->S; ma.k = net->cfg.k_map;
  ma.mask_scale = net->mask_scale; ma.iou_thresh = 0.5f;      // 0.5: literal of loss_mask (:784-791)
  ma.rois = net->mask_rois; ma.assign = net->mask_assign; ma.npos = net->mask_npos;
  ma.loss = net->loss_acc + 5;
  note_launch(2);
  DY_TRY(launch_mask_loss(ma, st));
  // L2 regulariser over unlocked weights and biases (:38,118-123,138-140)
  note_launch();
  DY_TRY(launch_sumsq_multi(net->pseg_dev, net->n_pseg, net->loss_acc + 6, st));
  double acc[8];
  DY_CUDA(cudaMemcpyAsync(acc, net->loss_acc, 8 * 8, cudaMemcpyDeviceToHost, st));
  DY_CUDA(cudaStreamSynchronize(st));
  // losses_host: total, obj, noobj, cls, xy, wh, mask, l2
  double total = 0;
  for (int i = 0; i < 7; ++i) total += acc[i];
  losses_host[0] = (float)total;
  for (int i = 0; i < 7; ++i) losses_host[1 + i] = (float)acc[i];
  return user_end(net, st);
}

int dy_train_backward(dy_net* net, int32_t B, int32_t layer_hi, int32_t layer_lo, float* grad_flat_dev, void* stream) {
  NvtxRange nvtx_("dy_train_backward");
  DY_CHECK(net && net->train_ready && grad_flat_dev, "bad argument");
  DY_CHECK(layer_lo >= 1 && layer_hi <= 82 && layer_lo <= layer_hi, "layer range");
  DY_TRY(user_begin(net, (cudaStream_t)stream));
  const bool tc = net->cfg.precision == DY_PRECISION_BF16;
  for (int n = layer_hi; n >= layer_lo; --n) {
    if (tc) DY_TRY(train_backward_layer_tc(net, n, B, grad_flat_dev, (cudaStream_t)stream));
    else DY_TRY(train_backward_layer(net, n, B, grad_flat_dev, (cudaStream_t)stream));
  }
  return user_end(net, (cudaStream_t)stream);
}

int dy_train_apply(dy_net* net, const float* grad_flat_dev, float lr, float grad_scale, void* stream) {
  NvtxRange nvtx_("dy_train_apply");
  DY_CHECK(net && net->train_ready && grad_flat_dev, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  DY_TRY(user_begin(net, st));
  auto& L = net->L;
  net->adam_step += 1;
  const double t = (double)net->adam_step;
  const float lr_t = (float)(lr * sqrt(1.0 - pow((double)kAdamB2, t)) / (1.0 - pow((double)kAdamB1, t)));
  // one launch each: Adam over every trainable tensor, moving averages + inference fold of every trained
  // BatchNorm, bf16 re-packing of the forward / dgrad operands
  note_launch(2);
  DY_TRY(launch_adam_multi(net->pseg_dev, net->n_pseg, grad_flat_dev, net->adam_m, net->adam_v, lr_t, kAdamB1, kAdamB2,
                           kAdamEps, grad_scale, st));
  DY_TRY(launch_bn_post_multi(net->bseg_dev, net->n_bseg, kBnDecay, net->cfg.bn_eps, st));
  for (int n = 1; n <= 82; ++n) {
    LayerState& s = L[n];
    if (s.unlocked && !s.def.bn)
      DY_CUDA(cudaMemcpyAsync(s.d_shift, s.d_bias, s.def.cout * 4, cudaMemcpyDeviceToDevice, st));
  }
  if (net->cfg.precision == DY_PRECISION_BF16) {
    note_launch(2);
    DY_TRY(launch_pack_multi(net->kseg_dev, net->n_kseg, net->max_pack_tiles, st));
    for (int n = 2; n <= 82; ++n) {      // stride-2 convs: their per-parity-block dgrad operands
      LayerState& s = L[n];
      if (s.unlocked && s.dg_s2) {
        note_launch();
        DY_TRY(launch_pack_dgrad_s2_bf16(s.d_w_f32, s.def.cin0, s.def.cout, s.Cg, s.wdg0, st));
      }
    }
  }
  return user_end(net, st);      // the host-buffer pipeline must not read weights Adam is still rewriting
}

int dy_train_get_tensor(dy_net* net, int32_t layer, int32_t which, int32_t B, float* out_dev, void* stream) {
  DY_CHECK(net && net->train_ready && out_dev && layer >= 1 && layer <= 82, "bad argument");
  const LayerState& s = net->L[layer];
  if (net->cfg.precision == DY_PRECISION_BF16) {
    const LayerDef& d = s.def;
    if (which != 0 && s.dyf) {
      DY_CUDA(cudaMemcpyAsync(out_dev, s.dyf, out_elems(d, B) * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
      return DY_OK;
    }
    const __nv_bfloat16* src16 = which == 0 ? s.zb : s.dyb;
    DY_CHECK(src16 != nullptr && s.Cg == d.cout, "layer has no such training tensor");
    return launch_p1_to_nhwc(src16, out_dev, B, d.H, d.H, d.cout, FORM_SAME, (cudaStream_t)stream);
  }
  const float* src = which == 0 ? s.z : s.dy;
  DY_CHECK(src != nullptr, "layer does not take part in the backward pass");
  DY_CUDA(cudaMemcpyAsync(out_dev, src, out_elems(s.def, B) * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return DY_OK;
}

int dy_get_weights(dy_net* net, const char* tf_name, float* host, int64_t capacity) {
  DY_CHECK(net && tf_name && host, "null argument");
  std::string name(tf_name);
  const std::string prefix = "yolo/convolutional";
  DY_CHECK(name.compare(0, prefix.size(), prefix) == 0, "unknown variable");
  size_t p = prefix.size();
  int n = 0;
  while (p < name.size() && name[p] >= '0' && name[p] <= '9') n = n * 10 + (name[p++] - '0');
  DY_CHECK(n >= 1 && n <= 82 && p < name.size() && name[p] == '/', "unknown variable");
  const std::string what = name.substr(p + 1);
  LayerState& s = net->L[n];
  const int C = s.def.cout;
  const float* src = nullptr;
  long long cnt = C;
  std::vector<float>* hostcopy = nullptr;
  if (what == "weights") { src = s.d_w_f32; cnt = (long long)s.K * C; hostcopy = &s.w; }
  else if (what == "biases") { src = s.d_bias; hostcopy = &s.bias; }
  else if (what == "BatchNorm/gamma") { src = s.d_gamma; hostcopy = &s.gamma; }
  else if (what == "BatchNorm/beta") { src = s.d_beta; hostcopy = &s.beta; }
  else if (what == "BatchNorm/moving_mean") { src = s.d_mean; hostcopy = &s.mean; }
  else if (what == "BatchNorm/moving_variance") { src = s.d_var; hostcopy = &s.var; }
  DY_CHECK(hostcopy != nullptr, "unknown variable");
  DY_CHECK(capacity >= cnt, "output buffer too small");
  if (src != nullptr && net->train_ready) {
    DY_CUDA(cudaMemcpy(host, src, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  } else {
    DY_CHECK((long long)hostcopy->size() == cnt, "variable was never loaded");
    memcpy(host, hostcopy->data(), (size_t)cnt * 4);
  }
  return DY_OK;
}

}  // extern "C"
