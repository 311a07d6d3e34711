// kab_band.cuh -- one CTA per lattice, for chapter-length lattices with the reference's
// diagonal band (align.py:64-65): max_move = 4, labels in 1..V-1, V <= 128, S <= 3T,
// min(beam_size, S) + 12 <= R where R = 4 * NT ring slots.
//
//   * state v lives in ring slot v mod R of a double-buffered shared-memory score row; thread
//     t owns slots 4t..4t+3 (even slot == blank state).  As the window [lo_i, hi_i) slides up,
//     a chunk that has fallen more than 6 states below lo_i is recycled to its next alias
//     (v + R); slots outside the window always hold -inf, so recycled slots start inactive
//     without any clearing pass (needs lo_i - lo_{i-1} <= 3, i.e. S <= 3T).
//   * frames run in groups of 8.  A thread whose four cells stay inside the window for the whole
//     group (the common case) runs a group body with no window arithmetic at all; threads near
//     a window edge, outside the window, or due for recycling run the same body plus the exact
//     incremental window (S*i = q*T + r, no divisions) and -inf masks.
//   * per frame and thread: two LDS.128 (own slots + the 3-state halo below), six packed fp32x2
//     adds, four cell updates, one STS.128, one STS.U8 of the four 2-bit backpointers, one
//     block barrier.
//   * emission rows are staged ahead with 1-D bulk copies (cp.async.bulk + mbarrier, thread 0
//     waits and the frame barrier publishes); the gather for the next frame is issued before
//     the current frame's barrier.
//   * backpointers: byte (frame i, thread t) at i * NT + t.  They are staged in shared memory
//     and leave as one bulk store (cp.async.bulk shared -> global) per FB frames.
//   * backtrack in the same CTA: backpointer blocks come back with bulk copies
//     (double-buffered) and one thread walks them at shared-memory latency; the block's outputs
//     are then written coalesced by all threads.
#pragma once
#include "kab_common.cuh"

struct KabTrue { static constexpr bool value = true; };
struct KabFalse { static constexpr bool value = false; };

#define KAB_BAND_STAGES 3
#define KAB_BAND_BPBLOCK_BYTES 16384

template <int NT>
struct KabBandCfg {
  static constexpr int R = 4 * NT;                          // ring slots
  static constexpr int FB = KAB_BAND_BPBLOCK_BYTES / NT;    // frames per backpointer block
};

// dynamic shared memory layout (bytes):
//   [0, 128)                      mbarriers: STAGES emission + 2 backpointer blocks
//   [128, 128 + 2*R*4)            score ring, two buffers
//   then 2 * BPBLOCK_BYTES        backpointer blocks (forward staging / backtrack fetch)
//   then FB * 4                   path staging
//   then STAGES * stage_bytes     emission stages
template <int NT>
__host__ __device__ constexpr size_t kab_band_smem_fixed() {
  return 128 + 2 * (size_t)KabBandCfg<NT>::R * 4 + 2 * (size_t)KAB_BAND_BPBLOCK_BYTES +
         (size_t)KabBandCfg<NT>::FB * 4;
}

__device__ __forceinline__ void kab_bulk_s2g(void *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(kab_smem_u32(src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void kab_bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void kab_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// All NT threads, arriving warp by warp (possibly from different group bodies).
template <int NT>
__device__ __forceinline__ void kab_frame_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

template <int NT>
__global__ void __launch_bounds__(NT) kab_band_kernel(const KabLattice *__restrict__ lats, int n_lat,
                                                      KabParams p) {
  using Cfg = KabBandCfg<NT>;
  constexpr int R = Cfg::R, FB = Cfg::FB;
  extern __shared__ __align__(128) unsigned char kab_smem[];
  uint64_t *ebars = reinterpret_cast<uint64_t *>(kab_smem);
  uint64_t *bbars = ebars + KAB_BAND_STAGES;
  float *ring = reinterpret_cast<float *>(kab_smem + 128);
  unsigned char *bpblk = kab_smem + 128 + 2 * (size_t)R * 4;
  int *pathbuf = reinterpret_cast<int *>(bpblk + 2 * (size_t)KAB_BAND_BPBLOCK_BYTES);
  float *stage_base = reinterpret_cast<float *>(pathbuf + FB);
  __shared__ unsigned int s_item;
  __shared__ int s_vmax;

  const int tid = threadIdx.x;
  const float ninf = kab_neg_inf();
  if (tid == 0) {
    for (int s = 0; s < KAB_BAND_STAGES; ++s) kab_mbar_init(&ebars[s], 1);
    kab_mbar_init(&bbars[0], 1);
    kab_mbar_init(&bbars[1], 1);
    kab_fence_mbar_init();
  }
  __syncthreads();
  uint32_t echunks = 0;  // emission chunks issued so far by this CTA (mbarrier phase tracking)
  uint32_t bblocks = 0;  // backpointer blocks fetched so far

  for (;;) {
    if (tid == 0) {
      s_item = atomicAdd(p.queue, 1u);
      s_vmax = -1;
    }
    __syncthreads();
    const unsigned int item = s_item;
    if (item >= (unsigned int)n_lat) break;
    const KabLattice lat = lats[item];
    const int T = lat.T, S = 2 * lat.L + 1, V = p.V, W = p.W;
    const int F = p.stage_frames;
    const uint32_t stage_words = p.stage_bytes >> 2;
    const int n_chunks = (T + F - 1) / F;
    const uint16_t *col16 = p.col16 + lat.col_off;
    unsigned char *bp = p.bp + lat.bp_off;
    const uint32_t one = p.one;  // runtime 1 (see kab_blank_sel)

    // ---- emission pipeline.  Chunk g (counted over the CTA's lifetime) lives in stage
    // g % STAGES and completes phase (g / STAGES) & 1 of that stage's mbarrier.
    const uint32_t ec0 = echunks;
    // F*V*4 is a multiple of 16, so every chunk of this lattice has the same 16-byte skew
    const uint32_t skew = (uint32_t)(((lat.t_off * (int64_t)V * 4) & 15) >> 2);
    auto issue = [&](int c, uint32_t stg) {
      const int f0 = c * F, nf = min(F, T - f0);
      const KabStageDesc d = kab_stage_desc(p, lat.t_off, f0, nf);
      float *dst = stage_base + stg * stage_words;
      if (tid == 0) {
        kab_mbar_expect_tx(&ebars[stg], d.bytes);
        if (d.bytes) kab_bulk_g2s(dst, d.src, d.bytes, &ebars[stg]);
      }
      if (tid < (int)d.tail_n)
        dst[d.tail_word + tid] = __ldg(reinterpret_cast<const float *>(d.src) + d.tail_word + tid);
    };
    uint32_t st = ec0 % KAB_BAND_STAGES, ph = (ec0 / KAB_BAND_STAGES) & 1u;  // of the chunk being read
    for (int c = 0; c < min(n_chunks, KAB_BAND_STAGES); ++c) issue(c, (st + c) % KAB_BAND_STAGES);

    // ---- ring init: everything inactive except the virtual start state 0 (align.py:57-58)
    for (int sl = tid; sl < R; sl += NT) ring[sl] = sl == 0 ? 0.0f : ninf;

    int vb = 4 * tid;  // first state of this thread's chunk (alias level 0)
    auto load_cols = [&](int base, uint32_t &ca, uint32_t &cb) {
      ca = base + 1 < S ? 4u * col16[base >> 1] : 0u;
      cb = base + 3 < S ? 4u * col16[(base >> 1) + 1] : 0u;
    };
    uint32_t c1, c3, nc1, nc3;  // byte offsets of the two label columns; next alias prefetched
    load_cols(vb, c1, c3);
    load_cols(vb + R, nc1, nc3);
    const int half = W / 2;

    bool bad = false;
    if (tid == 0) {
      kab_mbar_wait(&ebars[st], ph);
      if (F < 16 && n_chunks > 1) {  // F == G: chunk 1 opens at group 1, before the look-ahead wait
        const uint32_t s1 = st + 1 == KAB_BAND_STAGES ? 0 : st + 1;
        kab_mbar_wait(&ebars[s1], s1 == 0 ? ph ^ 1u : ph);
      }
    }
    __syncthreads();  // ring init, tail words and chunk 0 visible to everybody
    // rowc: staged row whose emissions are in (eb, e1, e3); cn: chunk being read
    const char *rowc = reinterpret_cast<const char *>(stage_base + st * stage_words + skew);
    int cn = 0;
    auto check_chunk = [&](int c) {  // finiteness of chunk c's words (all threads, strided)
      const float *w = stage_base + st * stage_words + skew;
      const int nw = min(F, T - c * F) * V;
      for (int j = tid; j < nw; j += NT) bad |= !kab_finite(w[j]);
    };
    check_chunk(0);
    float eb = *reinterpret_cast<const float *>(rowc);
    float e1 = *reinterpret_cast<const float *>(rowc + c1);
    float e3 = *reinterpret_cast<const float *>(rowc + c3);

    // ---- window state.  S*i = q*T + r is tracked per GROUP of G frames (step S*G) and, inside
    // a group that needs it, per frame (step S): exact floor(S*i/T) without divisions.
    constexpr int G = 8;
    const int qd = S / T, rd = S % T;                                    // per-frame step
    const int qdg = (int)(((int64_t)S * G) / T), rdg = (int)(((int64_t)S * G) % T);  // per-group step
    int qg = 0, rg = 0;  // q, r at the first frame of the current group

    // ---- the frame loop, in groups of G frames.  F (frames per emission chunk) and FB (frames
    // per backpointer block) are multiples of G, so chunk crossings, stage refills and block
    // stores happen only between groups; inside a group a frame is: event check, two LDS.128,
    // the four cell updates, STS.128 + STS.U8, the emission gather for the next frame, barrier.
    float *bufA = ring, *bufB = ring + R;  // frame i reads buf[i & 1], writes buf[~i & 1]
    unsigned char *bpst = bpblk + tid;     // this thread's byte of the group's first staging row
    // per-frame window of a SLOW group (threads near a window edge or outside the window)
    int q = 0, r = 0;
    auto frame = [&](auto slow_tag, const float *__restrict__ pv, float *__restrict__ cu, unsigned char *bpdst,
                     const char *rownext, const bool has_next) {
      constexpr bool SLOW = decltype(slow_tag)::value;
      int lo = 0, hi = 0;
      if (SLOW) {
        lo = max(0, q - half);   // align.py:64
        hi = min(lo + W, S);     // align.py:65
        // recycle a chunk that lies entirely more than 3 states below the window
        while (vb + 3 < lo - 3) {
          vb += R;
          c1 = nc1; c3 = nc3;
          load_cols(vb + R, nc1, nc3);
          e1 = *reinterpret_cast<const float *>(rowc + c1);  // the prefetched emissions belonged
          e3 = *reinterpret_cast<const float *>(rowc + c3);  // to the old alias
        }
        q += qd; r += rd;
        if (r >= T) { r -= T; ++q; }
      }
      const float4 P = *reinterpret_cast<const float4 *>(pv + 4 * tid);
      const float4 H = *reinterpret_cast<const float4 *>(pv + ((4 * tid + R - 4) & (R - 1)));
      // candidates: (even, odd) state pairs share one packed add
      float t0, t1, t2, t3;
      kab_add2(P.x, P.y, eb, t0, t1);
      kab_add2(P.z, P.w, eb, t2, t3);
      const float th1 = __fadd_rn(H.w, eb), th3 = __fadd_rn(H.y, eb);
      float a0, a1, a2, a3, b0, b1, b2, b3;
      kab_add2(P.x, P.y, e1, a1, a0);
      kab_add2(H.z, H.w, e1, a3, a2);
      kab_add2(P.z, P.w, e3, b1, b0);
      kab_add2(P.x, P.y, e3, b3, b2);
      (void)t3;
      uint32_t m = 0;
      float4 N;
      N.x = kab_blank_sel(t0, th1, th3, m, 1u << 0, 2u << 0, one);
      N.y = kab_label_sel(a0, a1, a2, a3, m, 1u << 2, 2u << 2, one);
      N.z = kab_blank_sel(t2, t1, th1, m, 1u << 4, 2u << 4, one);
      N.w = kab_label_sel(b0, b1, b2, b3, m, 1u << 6, 2u << 6, one);
      if (SLOW) {  // cells outside [lo, hi) are inactive
        const int nlo = lo - vb, nhi = hi - vb;
        if (0 < nlo || 0 >= nhi) N.x = ninf;
        if (1 < nlo || 1 >= nhi) N.y = ninf;
        if (2 < nlo || 2 >= nhi) N.z = ninf;
        if (3 < nlo || 3 >= nhi) N.w = ninf;
      }
      *reinterpret_cast<float4 *>(cu + 4 * tid) = N;
      *bpdst = (unsigned char)m;
      if (has_next) {  // emissions of frame i+1 (its chunk was published by an earlier barrier)
        rowc = rownext;
        eb = *reinterpret_cast<const float *>(rownext);
        e1 = *reinterpret_cast<const float *>(rownext + c1);
        e3 = *reinterpret_cast<const float *>(rownext + c3);
      }
      kab_frame_barrier<NT>();
    };
    using SlowTag = KabTrue;
    using FastTag = KabFalse;

    const int n_groups = (T + G - 1) / G;
    const int VB = V * 4;  // bytes per emission row
    for (int g = 0; g < n_groups; ++g) {
      const int i0 = g * G, nfr = min(G, T - i0);
      // row of the first frame of the NEXT group: same chunk, or the start of the next stage
      const bool next_crosses = ((i0 + G) % F) == 0;
      const uint32_t nst = st + 1 == KAB_BAND_STAGES ? 0 : st + 1;
      const char *row0 = rowc;  // frame i0 (its emissions are already in eb/e1/e3)
      const char *rowng = next_crosses ? reinterpret_cast<const char *>(stage_base + nst * stage_words + skew)
                                       : row0 + G * VB;
      // window at the first frame of this group and (conservatively) at its last frame
      const int lo0 = max(0, qg - half), hi0 = min(lo0 + W, S);
      int qn = qg + qdg, rn = rg + rdg;
      if (rn >= T) { rn -= T; ++qn; }
      const int lo1 = max(0, qn - half);  // lo of the next group's first frame >= lo of every frame here
      // (warp-uniform choice: the frame barrier sits inside the group bodies, so a warp must not
      // split between them; whole warps arriving at bar.sync from different bodies is fine)
      const bool safe = __all_sync(KAB_FULL_MASK, nfr == G && vb >= lo1 && vb + 4 <= hi0);  // all cells inside for all G frames
      if (safe) {
#pragma unroll
        for (int f = 0; f < G; ++f)
          frame(FastTag{}, (f & 1) ? bufB : bufA, (f & 1) ? bufA : bufB, bpst + f * NT,
                f + 1 < G ? row0 + (f + 1) * VB : rowng, f + 1 < G || i0 + G < T);
      } else if (nfr == G) {
        q = qg; r = rg;
#pragma unroll
        for (int f = 0; f < G; ++f)
          frame(SlowTag{}, (f & 1) ? bufB : bufA, (f & 1) ? bufA : bufB, bpst + f * NT,
                f + 1 < G ? row0 + (f + 1) * VB : rowng, f + 1 < G || i0 + G < T);
      } else {
        q = qg; r = rg;
        for (int f = 0; f < nfr; ++f)
          frame(SlowTag{}, (f & 1) ? bufB : bufA, (f & 1) ? bufA : bufB, bpst + f * NT, row0 + (f + 1) * VB,
                f + 1 < nfr);
      }
      qg = qn; rg = rn;
      bpst += G * NT;
      // ---- between groups (uniform bookkeeping)
      if (next_crosses && i0 + G < T) {  // the next group opens chunk cn+1
        ++cn;
        st = nst;
        if (st == 0) ph ^= 1u;
        check_chunk(cn);
        // everybody has finished with chunk cn-1: refill its stage with chunk cn + STAGES - 1
        if (cn + KAB_BAND_STAGES - 1 < n_chunks)
          issue(cn + KAB_BAND_STAGES - 1, (st + KAB_BAND_STAGES - 1) % KAB_BAND_STAGES);
      }
      // thread 0 waits for the chunk that opens two groups ahead; any later barrier publishes it
      if (tid == 0 && ((i0 + 2 * G) % F) == 0 && i0 + 2 * G < T) {
        const int ahead = (i0 + 2 * G) / F - cn;  // 1 or 2 chunks ahead of the one being read
        uint32_t s2 = st, p2 = ph;
        for (int k = 0; k < ahead; ++k) {
          if (++s2 == KAB_BAND_STAGES) { s2 = 0; p2 ^= 1u; }
        }
        kab_mbar_wait(&ebars[s2], p2);
      }
      const bool block_done = ((i0 + G) & (FB - 1)) == 0 || i0 + G >= T;  // backpointer block complete
      if (block_done) {
        kab_fence_proxy_async();  // staged backpointer bytes -> visible to the bulk store
        if (tid == 0) kab_bulk_wait_read0();  // the previous bulk store no longer reads its buffer
        __syncthreads();
        const int blk = i0 / FB, nb = min(FB, T - blk * FB);
        if (tid == 0)
          kab_bulk_s2g(bp + (size_t)blk * FB * NT, bpblk + (size_t)(blk & 1) * KAB_BAND_BPBLOCK_BYTES,
                       (uint32_t)nb * NT);
        bpst = bpblk + (size_t)((blk + 1) & 1) * KAB_BAND_BPBLOCK_BYTES + tid;
      }
    }
    float *prev = (T & 1) ? bufB : bufA;  // buffer written by frame T-1
    echunks = ec0 + n_chunks;

    // ---- forced end state: highest active state of frame T-1 (align.py:99-101)
    {
      const float4 P = *reinterpret_cast<const float4 *>(prev + 4 * tid);
      int cand = -1;
      if (vb + 0 < S && P.x > ninf) cand = vb + 0;
      if (vb + 1 < S && P.y > ninf) cand = vb + 1;
      if (vb + 2 < S && P.z > ninf) cand = vb + 2;
      if (vb + 3 < S && P.w > ninf) cand = vb + 3;
      cand = __reduce_max_sync(KAB_FULL_MASK, cand);
      if ((tid & 31) == 0 && cand >= 0) atomicMax(&s_vmax, cand);
    }
    if (tid == 0) kab_bulk_wait0();  // all backpointer blocks are in global memory
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    int v = s_vmax;
    const int status = any_bad ? 3 : (v < 0 ? 1 : 0);
    if (tid == 0) {
      p.status[lat.index] = status;
      if (p.final_score)
        p.final_score[lat.index] = status == 0 ? prev[v & (R - 1)] : __int_as_float(0x7fc00000);
    }
    if (status == 0) {
      // ---- backtrack (== flush_determined_path, align.py:21-40)
      const int n_blocks = (T + FB - 1) / FB;
      const uint32_t bb0 = bblocks;
      auto fetch = [&](int blk) {  // thread 0 only
        const uint32_t g = bb0 + (uint32_t)(n_blocks - 1 - blk), bs = g & 1u;
        const int nfr = min(FB, T - blk * FB);
        const uint32_t bytes = (uint32_t)nfr * NT;
        kab_mbar_expect_tx(&bbars[bs], bytes);
        kab_bulk_g2s(bpblk + (size_t)bs * KAB_BAND_BPBLOCK_BYTES, bp + (size_t)blk * FB * NT, bytes, &bbars[bs]);
      };
      if (tid == 0) fetch(n_blocks - 1);
      int32_t *out_path = p.best_path + lat.t_off;
      int32_t *out_lab = p.best_labels + lat.t_off;
      float *out_sc = p.best_scores + lat.t_off;
      const float *lp = p.lp + lat.t_off * (int64_t)V;
      for (int blk = n_blocks - 1; blk >= 0; --blk) {
        const uint32_t g = bb0 + (uint32_t)(n_blocks - 1 - blk), bs = g & 1u;
        const int i0 = blk * FB, i1 = min(T, i0 + FB);
        if (tid == 0) {
          if (blk > 0) fetch(blk - 1);  // other buffer: its previous contents were consumed
          kab_mbar_wait(&bbars[bs], (g >> 1) & 1u);
          const unsigned char *blkp = bpblk + (size_t)bs * KAB_BAND_BPBLOCK_BYTES;
          for (int i = i1 - 1; i >= i0; --i) {
            const int slot = v & (R - 1);
            const unsigned char byte = blkp[(i - i0) * NT + (slot >> 2)];
            pathbuf[i - i0] = v;
            v -= kab_decode_move((byte >> (2 * (slot & 3))) & 3u, v);
          }
        }
        __syncthreads();
        for (int i = i0 + tid; i < i1; i += NT) {
          const int pv = pathbuf[i - i0];
          const int lab = (pv & 1) ? (int)col16[(pv - 1) >> 1] : 0;
          out_path[i] = pv;
          out_lab[i] = lab;                              // align.py:106
          out_sc[i] = __ldg(&lp[(int64_t)i * V + lab]);  // align.py:107
        }
        __syncthreads();
      }
      bblocks = bb0 + n_blocks;
    }
    __syncthreads();
  }
}
