// kab_warp.cuh -- one WARP per lattice, for the silence-split short segments of BASELINE
// config 2 (S = 2L+1 <= 256, window never clips: lo_i = 0, hi_i = S for every frame,
// max_move = 4, labels in 1..V-1, V <= 128).
//
//   * state row lives in registers: lane l owns states K*l .. K*l+K-1 (K = 2,4,6,8 even, so
//     even register index == blank state); the three halo scores come from lane l-1 by
//     warp shuffle -- no shared-memory round trip and no block barrier on the recurrence.
//   * emission rows (V floats per frame) are staged a few frames ahead into a per-warp
//     shared-memory ring with 1-D bulk copies (cp.async.bulk + mbarrier complete_tx); the
//     per-state emission is a conflict-free LDS gather row[col[v]] (V = 39 spans 1.2 banks rows).
//   * backpointers: 2 bits per cell, packed per lane into 32-bit words (32/BPF frames per
//     word) and stored as one coalesced 128-byte row per warp: word (i / FPW) * 32 + lane.
//   * the same warp then backtracks: 16 word-rows are loaded coalesced into registers and the
//     walk fetches the owner lane's word with a shuffle (no dependent global load per frame);
//     best_path / best_labels / best_scores leave as coalesced 32-frame stores.
//   * warps pull lattices from a global queue sorted by decreasing cost (LPT order).
#pragma once
#include "kab_common.cuh"

#define KAB_WARP_STAGES 3
#define KAB_WARPS_PER_CTA 4

template <int K>
struct KabWarpCfg {
  static constexpr int BPF = K <= 2 ? 4 : (K <= 4 ? 8 : 16);  // backpointer bits per lane per frame
  static constexpr int FPW = 32 / BPF;                        // frames per 32-bit word
};

// One frame of the recurrence for this lane's K states.  s[] holds frame i-1 on entry and
// frame i on return; returns the 2K backpointer bits of the frame.
template <int K>
__device__ __forceinline__ uint32_t kab_warp_frame(float (&s)[K], const float eb, const float (&el)[K / 2],
                                                   const int lane) {
  const float ninf = kab_neg_inf();
  float h1, h2, h3;  // scores of states K*lane-1, -2, -3 (previous frame)
  if (K >= 4) {
    h1 = __shfl_up_sync(KAB_FULL_MASK, s[K - 1], 1);
    h2 = __shfl_up_sync(KAB_FULL_MASK, s[K - 2], 1);
    h3 = __shfl_up_sync(KAB_FULL_MASK, s[K >= 4 ? K - 3 : 0], 1);
    if (lane == 0) { h1 = ninf; h2 = ninf; h3 = ninf; }
  } else {
    h1 = __shfl_up_sync(KAB_FULL_MASK, s[1], 1);
    h2 = __shfl_up_sync(KAB_FULL_MASK, s[0], 1);
    h3 = __shfl_up_sync(KAB_FULL_MASK, s[1], 2);
    if (lane == 0) { h1 = ninf; h2 = ninf; }
    if (lane < 2) h3 = ninf;
  }
  float n[K];
  uint32_t bits = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    // previous-frame score of state (K*lane + k - d)
    const float s0 = s[k];
    const float s1 = k >= 1 ? s[k >= 1 ? k - 1 : 0] : h1;
    const float s2 = k >= 2 ? s[k >= 2 ? k - 2 : 0] : (k == 1 ? h1 : h2);
    const float s3 = k >= 3 ? s[k >= 3 ? k - 3 : 0] : (k == 2 ? h1 : (k == 1 ? h2 : h3));
    uint32_t mv;
    if ((k & 1) == 0) n[k] = kab_cell_blank(s0, s1, s3, eb, mv);
    else n[k] = kab_cell_label(s0, s1, s2, s3, el[k >> 1], mv);
    bits |= mv << (2 * k);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = n[k];
  return bits;
}

template <int K>
__device__ void kab_warp_align(const KabLattice &lat, const KabParams &p, float *stage_base, uint64_t *bars,
                               uint32_t &chunk_counter, const int lane) {
  using Cfg = KabWarpCfg<K>;
  constexpr int BPF = Cfg::BPF, FPW = Cfg::FPW;
  const int T = lat.T, S = 2 * lat.L + 1, V = p.V;
  const int F = p.stage_frames;
  const uint32_t stage_words = p.stage_bytes >> 2;
  const int n_chunks = (T + F - 1) / F;
  const uint16_t *col16 = p.col16 + lat.col_off;
  uint32_t *bpw = reinterpret_cast<uint32_t *>(p.bp + lat.bp_off);

  // byte offsets (col * 4) of this lane's K/2 label states inside an emission row
  uint32_t coff[K / 2];
#pragma unroll
  for (int q = 0; q < K / 2; ++q) {
    const int v = K * lane + 2 * q + 1;
    coff[q] = v < S ? 4u * col16[(v - 1) >> 1] : 0u;
  }

  float s[K];
#pragma unroll
  for (int k = 0; k < K; ++k) s[k] = kab_neg_inf();
  if (lane == 0) s[0] = 0.0f;  // virtual start state 0, score 0 (align.py:57-58)

  // -- emission pipeline: chunk c uses stage (chunk_counter + c) % STAGES
  const uint32_t cc0 = chunk_counter;
  auto issue = [&](int c) {
    const uint32_t g = cc0 + c, st = g % KAB_WARP_STAGES;
    const int f0 = c * F, nf = min(F, T - f0);
    const KabStageDesc d = kab_stage_desc(p, lat.t_off, f0, nf);
    float *dst = stage_base + st * stage_words;
    if (lane == 0) {
      kab_mbar_expect_tx(&bars[st], d.bytes);
      if (d.bytes) kab_bulk_g2s(dst, d.src, d.bytes, &bars[st]);
    }
    if (lane < (int)d.tail_n)  // last (< 16 B) words of the whole log_probs buffer
      dst[d.tail_word + lane] = __ldg(reinterpret_cast<const float *>(d.src) + d.tail_word + lane);
  };
  const int pre = min(n_chunks, KAB_WARP_STAGES - 1);
  for (int c = 0; c < pre; ++c) issue(c);

  bool bad = false;
  uint32_t word = 0;
  int i = 0;
  for (int c = 0; c < n_chunks; ++c) {
    // the stage consumed in iteration c-1 is free again: refill it with chunk c + STAGES - 1
    __syncwarp();
    if (c + KAB_WARP_STAGES - 1 < n_chunks) {
      kab_fence_proxy_async();
      issue(c + KAB_WARP_STAGES - 1);
    }
    const uint32_t g = cc0 + c, st = g % KAB_WARP_STAGES;
    kab_mbar_wait(&bars[st], (g / KAB_WARP_STAGES) & 1u);
    __syncwarp();
    const int f0 = c * F, nf = min(F, T - f0);
    const uint32_t skew = (uint32_t)((((lat.t_off + f0) * (int64_t)V * 4) & 15) >> 2);
    const char *rowb = reinterpret_cast<const char *>(stage_base + st * stage_words + skew);
    for (int f = 0; f < nf; ++f, ++i, rowb += V * 4) {
      const float *row = reinterpret_cast<const float *>(rowb);
      for (int cidx = lane; cidx < V; cidx += 32) bad |= !kab_finite(row[cidx]);
      const float eb = row[0];
      float el[K / 2];
#pragma unroll
      for (int q = 0; q < K / 2; ++q) el[q] = *reinterpret_cast<const float *>(rowb + coff[q]);
      const uint32_t bits = kab_warp_frame<K>(s, eb, el, lane);
      const int sub = i % FPW;
      word |= bits << (sub * BPF);
      if (sub == FPW - 1) {
        bpw[(size_t)(i / FPW) * 32 + lane] = word;
        word = 0;
      }
    }
  }
  chunk_counter = cc0 + n_chunks;
  if (T % FPW) bpw[(size_t)(T / FPW) * 32 + lane] = word;
  __syncwarp();

  // -- forced end state: highest active state of frame T-1 (align.py:99-101)
  int cand = -1;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const int v = K * lane + k;
    if (v < S && s[k] > kab_neg_inf()) cand = v;
  }
  int v = __reduce_max_sync(KAB_FULL_MASK, cand);
  const bool any_bad = __any_sync(KAB_FULL_MASK, bad);
  const int status = any_bad ? 3 : (v < 0 ? 1 : 0);
  {
    float fs = __int_as_float(0x7fc00000);
    if (status == 0) {
      float mine = s[0];
#pragma unroll
      for (int k = 1; k < K; ++k) if (k == v % K) mine = s[k];
      fs = __shfl_sync(KAB_FULL_MASK, mine, v / K);
    }
    if (lane == 0) {
      p.status[lat.index] = status;
      if (p.final_score) p.final_score[lat.index] = fs;
    }
  }
  if (status != 0) return;

  // -- backtrack (== flush_determined_path, align.py:21-40), 16 word-rows per block
  const int n_rows = (T + FPW - 1) / FPW;
  int32_t *out_path = p.best_path + lat.t_off;
  int32_t *out_lab = p.best_labels + lat.t_off;
  float *out_sc = p.best_scores + lat.t_off;
  const float *lp = p.lp + lat.t_off * (int64_t)V;
  int myv = 0;
  for (int rb = ((n_rows - 1) / 16) * 16; rb >= 0; rb -= 16) {
    uint32_t wr[16];
#pragma unroll
    for (int r = 0; r < 16; ++r)
      wr[r] = (rb + r) < n_rows ? __ldcg(&bpw[(size_t)(rb + r) * 32 + lane]) : 0u;
#pragma unroll
    for (int r = 15; r >= 0; --r) {
#pragma unroll
      for (int f = FPW - 1; f >= 0; --f) {
        const int fi = (rb + r) * FPW + f;  // frame index (warp-uniform)
        if (fi < T) {
          const int owner = v / K, k = v - owner * K;
          const uint32_t w = __shfl_sync(KAB_FULL_MASK, wr[r], owner);
          const uint32_t mv = (w >> (f * BPF + 2 * k)) & 3u;
          if (lane == (fi & 31)) myv = v;
          v -= (int)mv;
        }
        if ((fi & 31) == 0) {  // lanes now hold frames fi .. fi+31
          const int t = fi + lane;
          if (t < T) {
            const int lab = (myv & 1) ? (int)col16[(myv - 1) >> 1] : 0;
            out_path[t] = myv;
            out_lab[t] = lab;                              // align.py:106
            out_sc[t] = __ldg(&lp[(int64_t)t * V + lab]);  // align.py:107
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(KAB_WARPS_PER_CTA * 32)
    kab_warp_kernel(const KabLattice *__restrict__ lats, int n_lat, KabParams p) {
  extern __shared__ __align__(128) unsigned char kab_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t *bars = reinterpret_cast<uint64_t *>(kab_smem) + warp * KAB_WARP_STAGES;
  float *stage_base = reinterpret_cast<float *>(kab_smem + 128 +
                                                (size_t)warp * KAB_WARP_STAGES * p.stage_bytes);
  if (lane == 0) {
    for (int s = 0; s < KAB_WARP_STAGES; ++s) kab_mbar_init(&bars[s], 1);
    kab_fence_mbar_init();
  }
  __syncwarp();
  uint32_t chunk_counter = 0;
  for (;;) {
    unsigned int item = 0;
    if (lane == 0) item = atomicAdd(p.queue, 1u);
    item = __shfl_sync(KAB_FULL_MASK, item, 0);
    if (item >= (unsigned int)n_lat) break;
    const KabLattice lat = lats[item];
    switch (lat.k) {
      case 2: kab_warp_align<2>(lat, p, stage_base, bars, chunk_counter, lane); break;
      case 4: kab_warp_align<4>(lat, p, stage_base, bars, chunk_counter, lane); break;
      case 6: kab_warp_align<6>(lat, p, stage_base, bars, chunk_counter, lane); break;
      default: kab_warp_align<8>(lat, p, stage_base, bars, chunk_counter, lane); break;
    }
    __syncwarp();
  }
}
