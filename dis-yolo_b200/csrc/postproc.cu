// YOLO head decode + threshold, per-class greedy NMS on a warp-ballot alive bitmask, top-k and
// position-sensitive mask assembly.  Reference: yolo/yolo3_net_pos.py:465-628, :862-952.
// All arithmetic is fp32 with explicit round-to-nearest intrinsics where the reference (TensorFlow)
// evaluates separate multiply / add ops, so that no FMA contraction changes a comparison.
#include "postproc.cuh"

namespace dy {

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return __fdiv_rn(1.f, __fadd_rn(1.f, expf(-x))); }
// mask assembly only (float maps compared at 1e-5 / thresholded at 0.5 by the consumer, never ordered or
// compared for equality): MUFU.EX2 + MUFU.RCP, ~5 instructions instead of ~40; absolute error < 1e-6
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

// ------------------------------------------------------------------------------------------
// decode: one thread per candidate (interpret_output :465-514 + filter_detections :523-561)
// ------------------------------------------------------------------------------------------
// one candidate: class-specific confidence, threshold, box, clip, append (any order: NMS is order independent)
__device__ __forceinline__ void decode_candidate(const DecodeArgs& a, const float (&t)[5 + kMaxClasses], int b, int j,
                                                 int g, int cy, int cx, int anchor, int idx, int n0) {
  // class-specific confidence = sigmoid(obj) * max softmax(cls)   (:528-548, softmax not sigmoid)
  float mx = t[5];
  int cls = 0;
  for (int c = 1; c < a.num_class; ++c)
    if (t[5 + c] > mx) { mx = t[5 + c]; cls = c; }     // first maximum, like tf.argmax
  float sum = 0.f;
  for (int c = 0; c < a.num_class; ++c) sum = __fadd_rn(sum, expf(__fsub_rn(t[5 + c], mx)));
  const float pmax = __fdiv_rn(1.f, sum);              // exp(0)/sum
  const float score = __fmul_rn(sigmoidf_(t[4]), pmax);
  // below the threshold and no dense output requested: the box is never looked at (keep = score >
  // thresh, :558) -- skip its two sigmoids, two exps and eight IEEE divisions
  if (!a.dense_box && !(score > a.thresh)) return;
  // box (:487-505): xy = (cell + sigmoid(t_xy)) / grid ; wh = exp(t_wh) * anchor / net
  const float gf = (float)g, nf = (float)a.net;
  const float xc = __fdiv_rn(__fadd_rn((float)cx, sigmoidf_(t[0])), gf);
  const float yc = __fdiv_rn(__fadd_rn((float)cy, sigmoidf_(t[1])), gf);
  const float w = __fdiv_rn(__fmul_rn(expf(t[2]), a.anchors[(3 * j + anchor) * 2 + 0]), nf);
  const float h = __fdiv_rn(__fmul_rn(expf(t[3]), a.anchors[(3 * j + anchor) * 2 + 1]), nf);
  const float hh = __fdiv_rn(h, 2.f), hw = __fdiv_rn(w, 2.f);
  float y1 = __fsub_rn(yc, hh), x1 = __fsub_rn(xc, hw), y2 = __fadd_rn(yc, hh), x2 = __fadd_rn(xc, hw);
  // clip_boxes_graph (:940-952)
  const float wy1 = __ldg(a.windows + b * 4 + 0), wx1 = __ldg(a.windows + b * 4 + 1);
  const float wy2 = __ldg(a.windows + b * 4 + 2), wx2 = __ldg(a.windows + b * 4 + 3);
  y1 = fmaxf(fminf(y1, wy2), wy1);
  x1 = fmaxf(fminf(x1, wx2), wx1);
  y2 = fmaxf(fminf(y2, wy2), wy1);
  x2 = fmaxf(fminf(x2, wx2), wx1);

  if (a.dense_box) {
    reinterpret_cast<float4*>(a.dense_box)[(long long)b * n0 + idx] = make_float4(y1, x1, y2, x2);
    a.dense_cls[(long long)b * n0 + idx] = cls;
    a.dense_score[(long long)b * n0 + idx] = score;
  }
  if (score > a.thresh) {
    const int slot = atomicAdd(a.cand_count + b, 1);
    if (slot < a.cap) {
      Cand c;
      c.y1 = y1; c.x1 = x1; c.y2 = y2; c.x2 = x2;
      c.score = score; c.idx = idx; c.cls = cls; c.pad = 0;
      a.cand[(long long)b * a.cap + slot] = c;
    }
  }
}

// One thread per grid CELL (its 3 anchors): for 3 classes a cell is 24 consecutive floats = six 16-byte
// loads, consecutive threads read consecutive cells (96 B apart), and the cell -> (y, x) division is paid
// once per 3 candidates.  Candidate index = scale offset + cell*3 + anchor (the reference's flattening order,
// :527-542).
__global__ void __launch_bounds__(256) decode_kernel(DecodeArgs a) {
  const int b = blockIdx.y;
  const int c0 = a.g[0] * a.g[0], c1 = c0 + a.g[1] * a.g[1], ncell = c1 + a.g[2] * a.g[2];
  const int n0 = 3 * ncell;
  const int cell_all = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell_all >= ncell) return;
  const int j = cell_all < c0 ? 0 : (cell_all < c1 ? 1 : 2);
  const int cell = cell_all - (j == 0 ? 0 : (j == 1 ? c0 : c1));
  const int g = a.g[j];
  const int cy = cell / g, cx = cell - cy * g;
  const int depth = 5 + a.num_class;
  const float* p = a.yolo[j] + ((long long)b * g * g + cell) * (3 * depth);
  const int idx0 = 3 * (cell_all);                     // = 3*(cells of the earlier scales) + cell*3
  float t[5 + kMaxClasses];
  if (depth == 8) {
    float4 q[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) q[i] = __ldg(reinterpret_cast<const float4*>(p) + i);
#pragma unroll
    for (int an = 0; an < 3; ++an) {
      const float4 u = q[2 * an], v = q[2 * an + 1];
      t[0] = u.x; t[1] = u.y; t[2] = u.z; t[3] = u.w;
      t[4] = v.x; t[5] = v.y; t[6] = v.z; t[7] = v.w;
      decode_candidate(a, t, b, j, g, cy, cx, an, idx0 + an, n0);
    }
  } else {
    for (int an = 0; an < 3; ++an) {
      for (int i = 0; i < depth; ++i) t[i] = __ldg(p + an * depth + i);
      decode_candidate(a, t, b, j, g, cy, cx, an, idx0 + an, n0);
    }
  }
}

// ------------------------------------------------------------------------------------------
// NMS
// ------------------------------------------------------------------------------------------
// ordering key: higher score first, ties -> lower candidate index (our definition; TF1's
// priority_queue leaves ties unspecified)
__device__ __forceinline__ unsigned long long cand_key(const Cand& c) {
  return ((unsigned long long)__float_as_uint(c.score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)c.idx);
}

// IoU exactly as tf.image.non_max_suppression evaluates it (fp32, separate mul/add/div)
__device__ __forceinline__ float iou_tf(const float4 a, const float4 b) {  // (y1,x1,y2,x2) as (x,y,z,w)
  const float ay1 = fminf(a.x, a.z), ay2 = fmaxf(a.x, a.z), ax1 = fminf(a.y, a.w), ax2 = fmaxf(a.y, a.w);
  const float by1 = fminf(b.x, b.z), by2 = fmaxf(b.x, b.z), bx1 = fminf(b.y, b.w), bx2 = fmaxf(b.y, b.w);
  const float aa = __fmul_rn(__fsub_rn(ay2, ay1), __fsub_rn(ax2, ax1));
  const float ab = __fmul_rn(__fsub_rn(by2, by1), __fsub_rn(bx2, bx1));
  if (aa <= 0.f || ab <= 0.f) return 0.f;
  const float iy1 = fmaxf(ay1, by1), ix1 = fmaxf(ax1, bx1), iy2 = fminf(ay2, by2), ix2 = fminf(ax2, bx2);
  const float inter = __fmul_rn(fmaxf(__fsub_rn(iy2, iy1), 0.f), fmaxf(__fsub_rn(ix2, ix1), 0.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
}

// One CTA (1024 threads) per (class, image): exact greedy NMS (= tf.image.non_max_suppression, :566-573) in
// BATCHES of the best kNmsBatch remaining candidates.
//   phase 0  the class's candidates are compacted (warp ballots) into structure-of-arrays form -- box, ordering
//            key, position in the candidate list; up to kNmsBatch of them live in shared memory, a larger class
//            (stress configuration: tens of thousands per class) goes to a global scratch area;
//   batch    a multi-level 12-bit radix select over the keys finds the threshold T such that the not yet
//            examined candidates with key >= T number at most kNmsBatch (whole histogram bins; a bin that is too
//            large on its own is refined by the next 12 key bits -- keys are unique, so this terminates); they are
//            loaded into shared memory, each one tested against the boxes kept so far (a candidate suppressed by
//            an earlier batch's selection never enters), and sorted by key (bitonic, in shared memory);
//   rounds   the greedy scan of the sorted batch: every warp finds the first alive bit of a ping-pong bitmask by
//            itself (two words per lane, ballot + ffs -- no block-wide arg-max), the box is appended to the keep
//            list, every thread clears its own candidates with IoU > threshold against it: ONE __syncthreads per
//            selected box.
// A candidate of a later batch has a smaller key than everything in the current one, so it can neither be
// selected before nor suppress anything in it: the batches reproduce the sequential algorithm exactly, ties
// included (the key carries the candidate index).  Stops at max_det selections (max_output_size).
constexpr int kNmsThreads = 1024;
constexpr int kNmsBatch = 2048;
constexpr int kNmsBins = 4096;

__global__ void __launch_bounds__(kNmsThreads) nms_kernel(NmsArgs a) {
  extern __shared__ __align__(16) uint8_t nms_smem[];
  float4* bbox = reinterpret_cast<float4*>(nms_smem);                                    // [kNmsBatch] arrival order
  float4* sbox = bbox + kNmsBatch;                                                       // [kNmsBatch] sorted order
  float4* kbox = sbox + kNmsBatch;                                                       // [max_det] kept boxes
  unsigned long long* bkey = reinterpret_cast<unsigned long long*>(kbox + ((a.max_det + 3) & ~3));   // [kNmsBatch]
  int* bpos = reinterpret_cast<int*>(bkey + kNmsBatch);                                  // [kNmsBatch] by arrival slot
  int* bslot = bpos + kNmsBatch;                                                         // [kNmsBatch] slot of sorted pos
  uint32_t* hist = reinterpret_cast<uint32_t*>(bslot + kNmsBatch);                       // [kNmsBins]
  __shared__ uint32_t abits[2][kNmsBatch / 32];
  __shared__ uint32_t wsum[32];
  __shared__ int count_s;
  __shared__ int sel_bin_s;
  __shared__ uint32_t sel_above_s;

  const int c = blockIdx.x, b = blockIdx.y;
  int n = a.cand_count[b];
  if (n > a.cap) n = a.cap;
  const Cand* cd = a.cand + (long long)b * a.cap;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = kNmsThreads >> 5;
  int* sel = a.sel + ((long long)b * a.num_class + c) * a.max_det;
  const long long ovf = ((long long)b * a.num_class + c) * a.cap;
  float4* gbox = a.ovf_box + ovf;
  unsigned long long* gkey = a.ovf_key + ovf;
  int* gpos = a.ovf_pos + ovf;

  // ---- phase 0: compaction (any order: the ordering key carries the candidate index) ----
  if (tid == 0) count_s = 0;
  __syncthreads();
  for (int i0 = warp * 32; i0 < n; i0 += nwarps * 32) {
    const int i = i0 + lane;
    Cand ci;
    bool m = false;
    if (i < n) {
      ci = cd[i];
      m = ci.cls == c;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, m);
    if (bal == 0u) continue;
    int base = 0;
    if (lane == 0) base = atomicAdd(&count_s, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (m) {
      const int e = base + __popc(bal & ((1u << lane) - 1u));
      const float4 bx = make_float4(ci.y1, ci.x1, ci.y2, ci.x2);
      if (e < kNmsBatch) { bbox[e] = bx; bkey[e] = cand_key(ci); bpos[e] = i; }
      else { gbox[e - kNmsBatch] = bx; gkey[e - kNmsBatch] = cand_key(ci); gpos[e - kNmsBatch] = i; }
    }
  }
  __syncthreads();
  const int nc = count_s;
  const bool big = nc > kNmsBatch;
  if (big) {
    // more than one batch: everything lives in the global scratch; the shared-memory part goes behind the overflow
    const int tail = nc - kNmsBatch;
    for (int e = tid; e < kNmsBatch; e += kNmsThreads) {
      gbox[tail + e] = bbox[e];
      gkey[tail + e] = bkey[e];
      gpos[tail + e] = bpos[e];
    }
    __syncthreads();
  }

  int nsel = 0;
  unsigned long long upper = ~0ull;          // keys >= upper have been examined already
  bool last = !big;
  int nb = big ? 0 : nc;                      // candidates of the current batch (arrival slots 0..nb-1)
  while (true) {
    if (big) {
      // ---- threshold T of the next batch: multi-level radix select over keys < upper ----
      unsigned long long T = 0ull;
      unsigned long long prefix = 0ull;      // value of the key bits fixed so far
      int fixed = 0;                         // number of high key bits fixed
      last = false;
      while (true) {
        const int bits = (64 - fixed) < 12 ? (64 - fixed) : 12;
        const int shift = 64 - fixed - bits;
        const uint32_t dmask = (1u << bits) - 1u;
        for (int i = tid; i < kNmsBins; i += kNmsThreads) hist[i] = 0u;
        __syncthreads();
        for (int e = tid; e < nc; e += kNmsThreads) {
          const unsigned long long k = gkey[e];
          if (k < upper && (fixed == 0 || (k >> (64 - fixed)) == prefix))
            atomicAdd(&hist[(uint32_t)(k >> shift) & dmask], 1u);
        }
        __syncthreads();
        // scan the bins from the top: thread t owns reversed bins 4t..4t+3 (bin = kNmsBins-1 - r)
        uint32_t h[4], mine = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          h[j] = hist[kNmsBins - 1 - (4 * tid + j)];
          mine += h[j];
        }
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        if (tid == 0) sel_bin_s = -1;
        __syncthreads();
        if (warp == 0) {
          uint32_t v = wsum[lane], w = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
          }
          wsum[lane] = w - v;                // exclusive warp offsets
        }
        __syncthreads();
        uint32_t above = wsum[warp] + incl - mine;     // candidates in bins above this thread's first bin
        if (above <= (uint32_t)kNmsBatch && above + mine > (uint32_t)kNmsBatch) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (above <= (uint32_t)kNmsBatch && above + h[j] > (uint32_t)kNmsBatch) {
              sel_bin_s = kNmsBins - 1 - (4 * tid + j);   // the first bin (from the top) that no longer fits
              sel_above_s = above;
            }
            above += h[j];
          }
        }
        __syncthreads();
        const int d = sel_bin_s;
        if (d < 0) {                           // everything that is left fits (only possible at level 0)
          T = 0ull;
          last = true;
          break;
        }
        if (sel_above_s > 0u) {                // whole bins above d form the batch
          T = ((prefix << bits) | (unsigned long long)(d + 1)) << shift;
          break;
        }
        prefix = (prefix << bits) | (unsigned long long)d;   // the top bin alone is too large: refine inside it
        fixed += bits;
        __syncthreads();
      }
      // ---- load the batch: keys in [T, upper), minus everything an earlier selection suppresses ----
      if (tid == 0) count_s = 0;
      __syncthreads();
      for (int e0 = warp * 32; e0 < nc; e0 += nwarps * 32) {
        const int e = e0 + lane;
        bool m = false;
        unsigned long long k = 0ull;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < nc) {
          k = gkey[e];
          m = k >= T && k < upper;
        }
        if (m) {
          bx = gbox[e];
          for (int i = 0; i < nsel; ++i)
            if (iou_tf(bx, kbox[i]) > a.iou_thr) { m = false; break; }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, m);
        if (bal == 0u) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&count_s, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (m) {
          const int slot = base + __popc(bal & ((1u << lane) - 1u));
          bbox[slot] = bx; bkey[slot] = k; bpos[slot] = gpos[e];
        }
      }
      __syncthreads();
      nb = count_s;
      upper = T;
    }
    // ---- sort the batch by key, descending (bitonic network over the next power of two) ----
    int ns = 32;
    while (ns < nb) ns <<= 1;
    for (int i = tid; i < ns; i += kNmsThreads) {
      bslot[i] = i;
      if (i >= nb) bkey[i] = 0ull;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= ns; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int t = tid; t < (ns >> 1); t += kNmsThreads) {
          const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ixj = i | j;
          const unsigned long long ka = bkey[i], kb = bkey[ixj];
          const bool desc = (i & k2) == 0;
          if (desc ? (ka < kb) : (ka > kb)) {
            bkey[i] = kb; bkey[ixj] = ka;
            const int sa = bslot[i];
            bslot[i] = bslot[ixj]; bslot[ixj] = sa;
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < nb; i += kNmsThreads) sbox[i] = bbox[bslot[i]];
    if (tid < kNmsBatch / 32) {
      const int lo = tid * 32;
      abits[0][tid] = nb >= lo + 32 ? 0xffffffffu : (nb > lo ? ((1u << (nb - lo)) - 1u) : 0u);
    }
    __syncthreads();
    // ---- greedy rounds over the sorted batch ----
    int cur = 0;
    const int nhalf = nb > 1024 ? 2 : 1;
    while (true) {
      const uint32_t w0 = abits[cur][lane], w1 = nhalf == 2 ? abits[cur][lane + 32] : 0u;
      const uint32_t nz0 = __ballot_sync(0xffffffffu, w0 != 0u), nz1 = __ballot_sync(0xffffffffu, w1 != 0u);
      if ((nz0 | nz1) == 0u) break;
      int pp;
      if (nz0) {
        const int l = __ffs(nz0) - 1;
        pp = l * 32 + __ffs(__shfl_sync(0xffffffffu, w0, l)) - 1;
      } else {
        const int l = __ffs(nz1) - 1;
        pp = (32 + l) * 32 + __ffs(__shfl_sync(0xffffffffu, w1, l)) - 1;
      }
      const float4 sb = sbox[pp];
      if (tid == 0) {
        sel[nsel] = bpos[bslot[pp]];
        kbox[nsel] = sb;
      }
      ++nsel;
      if (nsel >= a.max_det) break;
      // thread t owns sorted positions t (word `warp`) and 1024 + t (word 32 + `warp`)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        if (hh < nhalf) {
          const uint32_t word = __shfl_sync(0xffffffffu, hh ? w1 : w0, warp);
          uint32_t m = 0u;
          if (word != 0u) {                       // warp-uniform
            const int e = (hh * 32 + warp) * 32 + lane;
            bool al = (word >> lane) & 1u;
            if (al && (e == pp || iou_tf(sbox[e], sb) > a.iou_thr)) al = false;
            m = __ballot_sync(0xffffffffu, al);
          }
          if (lane == 0) abits[cur ^ 1][hh * 32 + warp] = m;
        }
      }
      __syncthreads();
      cur ^= 1;
    }
    if (last || nsel >= a.max_det) break;
    __syncthreads();                              // (kbox / batch arrays are rewritten by the next batch)
  }
  if (tid == 0) a.sel_cnt[b * a.num_class + c] = nsel;
}

// ------------------------------------------------------------------------------------------
// finalize: union of per-class keep sets -> top-k by score (ties: lower index) -> zero padded
// [max_det,6] rows (:590-628); then val_test's rounding / positive-size filter and the integer
// bin edges of assemble_kmask_from_box (:876-897).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void box_edges(const Cand& c, int S, int k, float (&pb)[4], int* gx, int* gy) {
  const float Sf = (float)S;
  pb[0] = rintf(__fmul_rn(c.y1, Sf));    // tf.round = half-to-even
  pb[1] = rintf(__fmul_rn(c.x1, Sf));
  pb[2] = rintf(__fmul_rn(c.y2, Sf));
  pb[3] = rintf(__fmul_rn(c.x2, Sf));
  const float sub_w = __fdiv_rn(__fsub_rn(pb[3], pb[1]), (float)k);
  const float sub_h = __fdiv_rn(__fsub_rn(pb[2], pb[0]), (float)k);
  gx[0] = (int)pb[1];
  gy[0] = (int)pb[0];
  for (int j = 1; j < k; ++j) {
    gx[j] = (int)rintf(__fadd_rn(pb[1], __fmul_rn((float)j, sub_w)));
    gy[j] = (int)rintf(__fadd_rn(pb[0], __fmul_rn((float)j, sub_h)));
  }
  gx[k] = (int)pb[3];
  gy[k] = (int)pb[2];
}

__global__ void __launch_bounds__(1024) finalize_kernel(FinalizeArgs a) {
  extern __shared__ unsigned long long fin_smem[];
  const int Emax = a.num_class * a.max_det;
  unsigned long long* keys = fin_smem;                       // [Emax]
  int* pos = reinterpret_cast<int*>(keys + Emax);            // [Emax]
  int* rank_pos = pos + Emax;                                // [max_det] cand position by rank
  int* offs = rank_pos + a.max_det;                          // [max_det] output slot or -1
  __shared__ int E_s, nvalid_s;

  const int b = blockIdx.x;
  const Cand* cd = a.cand + (long long)b * a.cap;
  if (threadIdx.x == 0) {
    int e = 0;
    for (int c = 0; c < a.num_class; ++c) {
      const int cnt = a.sel_cnt[b * a.num_class + c];
      const int* s = a.sel + ((long long)b * a.num_class + c) * a.max_det;
      for (int i = 0; i < cnt; ++i) pos[e++] = s[i];
    }
    E_s = e;
  }
  __syncthreads();
  const int E = E_s;
  for (int e = threadIdx.x; e < E; e += blockDim.x) keys[e] = cand_key(cd[pos[e]]);
  __syncthreads();
  const int nraw = E < a.max_det ? E : a.max_det;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const unsigned long long k = keys[e];
    int r = 0;
    for (int j = 0; j < E; ++j) r += (keys[j] > k) ? 1 : 0;
    if (r < a.max_det) rank_pos[r] = pos[e];
  }
  __syncthreads();
  // raw rows + keep flags
  for (int r = threadIdx.x; r < a.max_det; r += blockDim.x) {
    float* row = a.det_raw + ((long long)b * a.max_det + r) * 6;
    if (r < nraw) {
      const Cand c = cd[rank_pos[r]];
      row[0] = c.y1; row[1] = c.x1; row[2] = c.y2; row[3] = c.x2;
      row[4] = (float)c.cls; row[5] = c.score;
      float pb[4];
      int gx[kMaxK + 1], gy[kMaxK + 1];
      box_edges(c, a.S, a.k, pb, gx, gy);
      offs[r] = (__fsub_rn(pb[2], pb[0]) > 0.f && __fsub_rn(pb[3], pb[1]) > 0.f) ? 1 : 0;
    } else {
      for (int i = 0; i < 6; ++i) row[i] = 0.f;
      offs[r] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nv = 0;
    for (int r = 0; r < nraw; ++r) {
      const int keep = offs[r];
      offs[r] = keep ? nv : -1;
      nv += keep;
    }
    for (int r = nraw; r < a.max_det; ++r) offs[r] = -1;
    nvalid_s = nv;
    a.raw_count[b] = nraw;
    a.det_count[b] = nv;
  }
  __syncthreads();
  const int nv = nvalid_s;
  for (int r = threadIdx.x; r < a.max_det; r += blockDim.x) {
    if (r < nraw && offs[r] >= 0) {
      const int o = offs[r];
      const Cand c = cd[rank_pos[r]];
      float* row = a.det_box + ((long long)b * a.max_det + o) * 6;
      row[0] = c.y1; row[1] = c.x1; row[2] = c.y2; row[3] = c.x2;
      row[4] = (float)c.cls; row[5] = c.score;
      float pb[4];
      int gx[kMaxK + 1], gy[kMaxK + 1];
      box_edges(c, a.S, a.k, pb, gx, gy);
      int* ed = a.edges + ((long long)b * a.max_det + o) * (2 * (kMaxK + 1));
      for (int j = 0; j <= a.k; ++j) {
        ed[j] = gx[j];
        ed[kMaxK + 1 + j] = gy[j];
      }
    }
    if (r >= nv) {
      float* row = a.det_box + ((long long)b * a.max_det + r) * 6;
      for (int i = 0; i < 6; ++i) row[i] = 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------
// position-sensitive mask assembly (:862-933): out[d,y,x] = sigmoid(score[y,x, by*k+bx]) inside
// bin (by,bx) of box d, sigmoid(0)=0.5 outside the box.  One thread = 4 consecutive x (float4
// store); planar score maps make the gather a coalesced row read.
// ------------------------------------------------------------------------------------------
constexpr int kMaskRowsPerCta = 96;    // fat CTAs: the two dependent global loads at CTA start (count, edges) are
                                         // amortised over ~110 KB of stores

// One CTA per (32-row slab, detection, image); blockDim = (S/4) * rpi threads: thread t owns the float4
// column q = t % (S/4) of rows (t / (S/4)) + i*rpi of the slab.  Everything that depends on x (inside the
// box?  which horizontal bin?  gather offsets) is computed once per thread; a row outside the box or a
// column quad outside it costs one streaming store of 0.5 -- the kernel is a 436 MB fill at batch 64
// with a gather inside the boxes, and every warp store instruction writes 512 contiguous bytes.
// one (96-row slab, detection, image) work item
template <bool kStream>
__device__ __forceinline__ void mask_slab(const MaskArgs& a, int rpi, int slab, int d, int b) {
  const int* ed = a.edges + ((long long)b * a.max_det + d) * (2 * (kMaxK + 1));
  int gx[kMaxK + 1], gy[kMaxK + 1];
#pragma unroll
  for (int j = 0; j <= kMaxK; ++j) {
    gx[j] = (j <= a.k) ? __ldg(ed + j) : 0x7fffffff;
    gy[j] = (j <= a.k) ? __ldg(ed + kMaxK + 1 + j) : 0x7fffffff;
  }
  const int x_lo = gx[0], x_hi = gx[a.k], y_lo = gy[0], y_hi = gy[a.k];
  const int qpr = a.S >> 2;
  const int q = threadIdx.x % qpr, rl = threadIdx.x / qpr;
  const int x0 = q << 2;
  // per-thread column state
  bool inx[4];
  long long xoff[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int x = x0 + t;
    int bx = 0;
#pragma unroll
    for (int j = 1; j < kMaxK; ++j)
      if (j < a.k && x >= gx[j]) bx = j;
    inx[t] = x >= x_lo && x < x_hi;
    xoff[t] = (long long)bx * a.s_ch + (long long)x * a.s_pix;
  }
  const bool any_x = inx[0] || inx[1] || inx[2] || inx[3];
  const int y0 = slab * kMaskRowsPerCta;
  const int y1 = min(a.S, y0 + kMaskRowsPerCta);
  const float* sbase = a.score + b * a.s_img;
  float4* op = reinterpret_cast<float4*>(a.out + (((long long)b * a.max_det + d) * a.S + y0 + rl) * a.S) + q;
  const long long ostep = (long long)rpi * qpr;
  const float4 half4 = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
  // four rows per pass, every gather load of the pass issued before the first sigmoid: a thread inside the box
  // has up to 16 independent L2 requests in flight instead of waiting out one round trip per row
  for (int y = y0 + rl; y < y1; y += 4 * rpi, op += 4 * ostep) {
    float v[4][4];
    bool in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int yy = y + u * rpi;
      in[u] = any_x && yy < y1 && yy >= y_lo && yy < y_hi;
      if (in[u]) {
        int by = 0;
#pragma unroll
        for (int j = 1; j < kMaxK; ++j)
          if (j < a.k && yy >= gy[j]) by = j;
        const float* srow = sbase + (long long)(by * a.k) * a.s_ch + (long long)yy * a.s_row;
#pragma unroll
        for (int t = 0; t < 4; ++t) v[u][t] = inx[t] ? __ldg(srow + xoff[t]) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (y + u * rpi >= y1) break;
      float4 o = half4;
      if (in[u]) {
        if (inx[0]) o.x = sigmoid_fast(v[u][0]);
        if (inx[1]) o.y = sigmoid_fast(v[u][1]);
        if (inx[2]) o.z = sigmoid_fast(v[u][2]);
        if (inx[3]) o.w = sigmoid_fast(v[u][3]);
      }
      if (kStream) __stcs(op + u * ostep, o);
      else op[u * ostep] = o;
    }
  }
}

// grid = (slabs, max_det, B): one CTA per possible work item (large batches: B > kMaskMaxB)
template <bool kStream>
__global__ void __launch_bounds__(512) mask_kernel(MaskArgs a, int rpi) {
  const int d = blockIdx.y, b = blockIdx.z;
  if (d >= __ldg(a.det_count + b)) return;
  mask_slab<kStream>(a, rpi, blockIdx.x, d, b);
}

// Work-list form (option mask_work_list, off by default): a fixed grid of CTAs walks the EXISTING (slab,
// detection) items only -- at batch 64 the (slabs, max_det, B) grid above is 19,200 CTAs of which ~4,000 have work.
// Measured: the empty CTAs are NOT what the kernel waits for (90 us with them, 104 us with the work list: fewer,
// longer-lived CTAs hide the gather latency worse).  Every CTA
// scans det_count once (warp shuffles, B <= kMaskMaxB), then maps item w -> (image, detection) by a binary
// search of the prefix in shared memory; consecutive CTAs write consecutive slabs of the same map.
constexpr int kMaskMaxB = 2048;
template <bool kStream>
__global__ void __launch_bounds__(512) mask_list_kernel(MaskArgs a, int rpi, int nslab) {
  __shared__ int pre[kMaskMaxB + 1];
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int run = 0;
    if (lane == 0) pre[0] = 0;
    for (int b0 = 0; b0 < a.B; b0 += 32) {
      int c = b0 + lane < a.B ? min(__ldg(a.det_count + b0 + lane), a.max_det) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (b0 + lane < a.B) pre[b0 + lane + 1] = run + incl;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  __syncthreads();
  const int total = pre[a.B] * nslab;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int slab = w % nslab, i = w / nslab;
    int lo = 0, hi = a.B;                       // largest b with pre[b] <= i
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (pre[mid] <= i) lo = mid;
      else hi = mid;
    }
    mask_slab<kStream>(a, rpi, slab, i - pre[lo], lo);
  }
}

int g_mask_stream = 1;
int g_mask_list = 0;      // 1: work-list kernel, 0 (default; measured 13 % faster at batch 64): one CTA per possible (slab, detection, image)

// ------------------------------------------------------------------------------------------
// Box-cropped form of the same maps: the reference's consumer only ever reads
// det_mask[y1:y2, x1:x2] with (y1,x1,y2,x2) = round(box * S) (calculate_test_map.py:247-252), i.e. the
// box region whose bin edges finalize_kernel already holds.  crop_offsets_kernel: exclusive scan of the
// crop areas over (image, detection) -> off[B*max_det + 1] (off[last] = total floats).
// crop_mask_kernel: crop d of image b = rows [y1,y2) x cols [x1,x2) of its [S,S] map, row-major, at off.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void crop_rect(const int* ed, int k, int S, int& x1, int& y1, int& w, int& h) {
  x1 = max(ed[0], 0);
  y1 = max(ed[kMaxK + 1], 0);
  w = max(min(ed[k], S) - x1, 0);
  h = max(min(ed[kMaxK + 1 + k], S) - y1, 0);
}

__global__ void __launch_bounds__(1024) crop_offsets_kernel(MaskArgs a, long long* off) {
  __shared__ long long part[1024];
  const int n = a.B * a.max_det;
  const int per = (n + 1023) / 1024;
  const int i0 = threadIdx.x * per;
  long long sum = 0;
  for (int i = i0; i < i0 + per && i < n; ++i) {
    const int b = i / a.max_det, d = i - b * a.max_det;
    if (d < a.det_count[b]) {
      int x1, y1, w, h;
      crop_rect(a.edges + (long long)i * (2 * (kMaxK + 1)), a.k, a.S, x1, y1, w, h);
      sum += (long long)w * h;
    }
  }
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {          // Hillis-Steele inclusive scan
    const long long v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  long long run = part[threadIdx.x] - sum;      // exclusive prefix of this thread's chunk
  for (int i = i0; i < i0 + per && i < n; ++i) {
    off[i] = run;
    const int b = i / a.max_det, d = i - b * a.max_det;
    if (d < a.det_count[b]) {
      int x1, y1, w, h;
      crop_rect(a.edges + (long long)i * (2 * (kMaxK + 1)), a.k, a.S, x1, y1, w, h);
      run += (long long)w * h;
    }
  }
  if (threadIdx.x == 1023) off[n] = part[1023];
}

constexpr int kCropPerCta = 4096;   // crop elements per CTA (256 threads x 16)

__global__ void __launch_bounds__(256) crop_mask_kernel(MaskArgs a, const long long* __restrict__ off,
                                                        float* __restrict__ out) {
  const int d = blockIdx.y, b = blockIdx.z;
  if (d >= __ldg(a.det_count + b)) return;
  const int* ed = a.edges + ((long long)b * a.max_det + d) * (2 * (kMaxK + 1));
  int gx[kMaxK + 1], gy[kMaxK + 1];
#pragma unroll
  for (int j = 0; j <= kMaxK; ++j) {
    gx[j] = (j <= a.k) ? __ldg(ed + j) : 0x7fffffff;
    gy[j] = (j <= a.k) ? __ldg(ed + kMaxK + 1 + j) : 0x7fffffff;
  }
  int x1, y1, w, h;
  crop_rect(ed, a.k, a.S, x1, y1, w, h);
  const int area = w * h;
  const int e0 = blockIdx.x * kCropPerCta;
  if (e0 >= area) return;
  const float* sbase = a.score + b * a.s_img;
  float* o = out + __ldg(off + (long long)b * a.max_det + d);
  const int e1 = min(area, e0 + kCropPerCta);
  for (int e = e0 + threadIdx.x; e < e1; e += 256) {
    const int r = e / w;
    const int y = y1 + r, x = x1 + (e - r * w);
    int bx = 0, by = 0;
#pragma unroll
    for (int j = 1; j < kMaxK; ++j) {
      if (j < a.k && x >= gx[j]) bx = j;
      if (j < a.k && y >= gy[j]) by = j;
    }
    const float v = __ldg(sbase + (long long)(by * a.k + bx) * a.s_ch + (long long)y * a.s_row + (long long)x * a.s_pix);
    __stcs(o + e, sigmoid_fast(v));
  }
}

}  // namespace

void masks_set_streaming(int on) { g_mask_stream = on; }
void masks_set_work_list(int on) { g_mask_list = on; }

int launch_decode(const DecodeArgs& a, cudaStream_t st) {
  DY_CHECK(a.num_class >= 1 && a.num_class <= kMaxClasses, "num_class");
  const int n0 = 3 * (a.g[0] * a.g[0] + a.g[1] * a.g[1] + a.g[2] * a.g[2]);
  dim3 grid((n0 / 3 + 255) / 256, a.B);              // one thread per cell
  decode_kernel<<<grid, 256, 0, st>>>(a);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_nms(const NmsArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)kNmsBatch * (16 + 16 + 8 + 4 + 4) + (size_t)((a.max_det + 3) & ~3) * 16 + (size_t)kNmsBins * 4 + 16;
  DY_CHECK(smem <= 200 * 1024, "max_detection too large for the NMS keep list");
  DY_CHECK(a.ovf_box && a.ovf_key && a.ovf_pos, "NMS overflow scratch missing");
  static bool attr = false;
  if (!attr) {
    DY_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid(a.num_class, a.B);
  nms_kernel<<<grid, kNmsThreads, smem, st>>>(a);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_finalize(const FinalizeArgs& a, cudaStream_t st) {
  DY_CHECK(a.k >= 1 && a.k <= kMaxK, "k_map");
  const int Emax = a.num_class * a.max_det;
  const size_t smem = (size_t)Emax * 12 + (size_t)a.max_det * 8 + 16;
  DY_CHECK(smem <= 96 * 1024, "max_detection too large");
  static bool attr = false;
  if (!attr) {
    DY_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  // (the rank pass is O(E^2 / threads): 1024 threads for the stress configuration's 3 x 1000 entries)
  finalize_kernel<<<a.B, Emax > 512 ? 1024 : 256, smem, st>>>(a);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}


int launch_masks(const MaskArgs& a, cudaStream_t st) {
  DY_CHECK(a.S % 4 == 0, "score map size must be a multiple of 4");
  DY_CHECK(a.max_det <= 65535 && a.B <= 65535, "grid limits");
  const int qpr = a.S / 4;
  DY_CHECK(qpr <= 512, "score map too wide (S <= 2048)");
  int rpi = (256 + qpr / 2) / qpr;       // ~256 threads per CTA, a whole number of rows per pass
  if (rpi < 1) rpi = 1;
  const int nslab = (a.S + kMaskRowsPerCta - 1) / kMaskRowsPerCta;
  if (a.B <= kMaskMaxB && g_mask_list) {
    static int sms = 0;
    if (!sms) {
      int dev = 0;
      DY_CUDA(cudaGetDevice(&dev));
      DY_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long items = (long long)nslab * a.max_det * a.B;
    const int per_sm = 2048 / (qpr * rpi) > 8 ? 8 : 2048 / (qpr * rpi);      // resident CTAs per SM
    const int grid = (int)(items < (long long)sms * per_sm ? items : (long long)sms * per_sm);
    if (g_mask_stream) mask_list_kernel<true><<<grid, qpr * rpi, 0, st>>>(a, rpi, nslab);
    else mask_list_kernel<false><<<grid, qpr * rpi, 0, st>>>(a, rpi, nslab);
  } else {
    dim3 grid(nslab, a.max_det, a.B);
    if (g_mask_stream) mask_kernel<true><<<grid, qpr * rpi, 0, st>>>(a, rpi);
    else mask_kernel<false><<<grid, qpr * rpi, 0, st>>>(a, rpi);
  }
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_crop_offsets(const MaskArgs& a, long long* off, cudaStream_t st) {
  DY_CHECK(off != nullptr, "null offsets");
  crop_offsets_kernel<<<1, 1024, 0, st>>>(a, off);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

int launch_masks_cropped(const MaskArgs& a, const long long* off, float* out, cudaStream_t st) {
  DY_CHECK(a.max_det <= 65535 && a.B <= 65535, "grid limits");
  dim3 grid((a.S * a.S + kCropPerCta - 1) / kCropPerCta, a.max_det, a.B);
  crop_mask_kernel<<<grid, 256, 0, st>>>(a, off, out);
  DY_CUDA(cudaGetLastError());
  return DY_OK;
}

}  // namespace dy
