// kab_band.cuh -- one CTA per lattice, for chapter-length lattices with the reference's
// diagonal band (align.py:64-65): max_move = 4, labels in 1..V-1, V <= 128, S <= 3T,
// min(beam_size, S) + 12 <= R where R = 4 * NT ring slots.
//
//   * state v lives in ring slot v mod R of a double-buffered shared-memory score row; thread
//     t owns slots 4t..4t+3 (even slot == blank state).  As the window [lo_i, hi_i) slides up,
//     a chunk that has fallen more than 6 states below lo_i is recycled to its next alias
//     (v + R); slots outside the window always hold -inf, so recycled slots start inactive
//     without any clearing pass (needs lo_i - lo_{i-1} <= 3, i.e. S <= 3T).
//   * per frame and thread: two LDS.128 (own slots + the 3-state halo below), four cell
//     updates, one STS.128, one block barrier.  Cells outside the window are computed and then
//     overwritten with -inf by the (few) edge threads only.
//   * emission rows are staged ahead with 1-D bulk copies (cp.async.bulk + mbarrier); the
//     gather for the next frame is issued before the current frame's barrier.
//   * backpointers: one byte (4 cells x 2 bits) per thread and frame, four frames per 32-bit
//     word, stored coalesced: word (i / 4) * NT + t.  R/4 bytes per frame (256 B at R = 1024).
//   * backtrack in the same CTA: backpointer blocks come back through shared memory with bulk
//     copies (double-buffered) and one thread walks them at shared-memory latency; the block's
//     outputs are then written coalesced by all threads.
#pragma once
#include "kab_common.cuh"

#define KAB_BAND_STAGES 3
#define KAB_BAND_BPBLOCK_BYTES 16384

template <int NT>
struct KabBandCfg {
  static constexpr int R = 4 * NT;                                  // ring slots
  static constexpr int ROWS_PER_BLOCK = KAB_BAND_BPBLOCK_BYTES / (4 * NT);  // word-rows per backtrack block
  static constexpr int FRAMES_PER_BLOCK = 4 * ROWS_PER_BLOCK;
};

// dynamic shared memory layout (bytes):
//   [0, 128)                      mbarriers: STAGES emission + 2 backpointer blocks
//   [128, 128 + 2*R*4)            score ring, two buffers
//   then 2 * BPBLOCK_BYTES        backpointer blocks
//   then FRAMES_PER_BLOCK * 4     path staging
//   then STAGES * stage_bytes     emission stages
template <int NT>
__host__ __device__ constexpr size_t kab_band_smem_fixed() {
  return 128 + 2 * (size_t)KabBandCfg<NT>::R * 4 + 2 * (size_t)KAB_BAND_BPBLOCK_BYTES +
         (size_t)KabBandCfg<NT>::FRAMES_PER_BLOCK * 4;
}

template <int NT>
__global__ void __launch_bounds__(NT) kab_band_kernel(const KabLattice *__restrict__ lats, int n_lat,
                                                      KabParams p) {
  using Cfg = KabBandCfg<NT>;
  constexpr int R = Cfg::R, RPB = Cfg::ROWS_PER_BLOCK, FPB = Cfg::FRAMES_PER_BLOCK;
  extern __shared__ __align__(128) unsigned char kab_smem[];
  uint64_t *ebars = reinterpret_cast<uint64_t *>(kab_smem);
  uint64_t *bbars = ebars + KAB_BAND_STAGES;
  float *ring = reinterpret_cast<float *>(kab_smem + 128);
  unsigned char *bpblk = kab_smem + 128 + 2 * (size_t)R * 4;
  int *pathbuf = reinterpret_cast<int *>(bpblk + 2 * (size_t)KAB_BAND_BPBLOCK_BYTES);
  float *stage_base = reinterpret_cast<float *>(pathbuf + FPB);
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
    uint32_t *bpw = reinterpret_cast<uint32_t *>(p.bp + lat.bp_off);

    // ---- emission pipeline (thread 0 issues, everybody waits on the stage's mbarrier).
    // Chunk g (counted over the CTA's lifetime) lives in stage g % STAGES and completes phase
    // (g / STAGES) & 1 of that stage's mbarrier; both are tracked incrementally.
    const uint32_t ec0 = echunks;
    // F*V*4 is a multiple of 16, so every chunk of this lattice has the same 16-byte skew
    const uint32_t skew = (uint32_t)(((lat.t_off * (int64_t)V * 4) & 15) >> 2);
    auto issue = [&](int c, uint32_t st) {
      const int f0 = c * F, nf = min(F, T - f0);
      const KabStageDesc d = kab_stage_desc(p, lat.t_off, f0, nf);
      float *dst = stage_base + st * stage_words;
      if (tid == 0) {
        kab_fence_proxy_async();
        kab_mbar_expect_tx(&ebars[st], d.bytes);
        if (d.bytes) kab_bulk_g2s(dst, d.src, d.bytes, &ebars[st]);
      }
      if (tid < (int)d.tail_n)
        dst[d.tail_word + tid] = __ldg(reinterpret_cast<const float *>(d.src) + d.tail_word + tid);
    };
    uint32_t st = ec0 % KAB_BAND_STAGES, ph = (ec0 / KAB_BAND_STAGES) & 1u;  // stage / phase of chunk `cn`
    for (int c = 0; c < min(n_chunks, KAB_BAND_STAGES); ++c) issue(c, (st + c) % KAB_BAND_STAGES);

    // ---- ring init: everything inactive except the virtual start state 0 (align.py:57-58)
    float *prev = ring, *cur = ring + R;
    for (int sl = tid; sl < R; sl += NT) prev[sl] = sl == 0 ? 0.0f : ninf;

    int vb = 4 * tid;  // first state of this thread's chunk (alias level 0)
    auto load_cols = [&](int base, uint32_t &c1, uint32_t &c3) {
      c1 = base + 1 < S ? 4u * col16[base >> 1] : 0u;
      c3 = base + 3 < S ? 4u * col16[(base >> 1) + 1] : 0u;
    };
    uint32_t c1, c3, nc1, nc3;  // byte offsets of the two label columns; next alias prefetched
    load_cols(vb, c1, c3);
    load_cols(vb + R, nc1, nc3);

    // window arithmetic: q_i = floor(S*i/T) kept incrementally (exact, no 64-bit division)
    const int qd = S / T, rd = S % T;
    int q = 0, acc = 0;
    const int half = W / 2;

    bool bad = false;
    const uint32_t one = p.one;  // runtime 1 (see kab_blank_sel)
    kab_mbar_wait(&ebars[st], ph);
    __syncthreads();  // ring init + tail words visible
    // rowc: staged row whose emissions are in (eb, e1, e3): frame `fin` of chunk `cn`
    const char *rowc = reinterpret_cast<const char *>(stage_base + st * stage_words + skew);
    int cn = 0, fin = 0;
    float eb, e1, e3;
    eb = *reinterpret_cast<const float *>(rowc);
    e1 = *reinterpret_cast<const float *>(rowc + c1);
    e3 = *reinterpret_cast<const float *>(rowc + c3);
    for (int c = tid; c < V; c += NT) bad |= !kab_finite(reinterpret_cast<const float *>(rowc)[c]);
    uint32_t word = 0;
    int wsh = 0;  // bit position of this frame's byte inside `word`
    uint32_t *bprow = bpw + tid;

    for (int i = 0; i < T; ++i) {
      const int lo = max(0, q - half);  // align.py:64
      const int hi = min(lo + W, S);    // align.py:65
      // recycle a chunk that lies entirely more than 3 states below the window
      while (vb + 3 < lo - 3) {
        vb += R;
        c1 = nc1; c3 = nc3;
        load_cols(vb + R, nc1, nc3);
        e1 = *reinterpret_cast<const float *>(rowc + c1);  // the prefetched emissions belonged
        e3 = *reinterpret_cast<const float *>(rowc + c3);  // to the old alias
      }
      const float4 P = *reinterpret_cast<const float4 *>(prev + 4 * tid);
      const float4 H = *reinterpret_cast<const float4 *>(prev + ((4 * tid + R - 4) & (R - 1)));
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
      uint32_t m = 0;
      float4 N;
      N.x = kab_blank_sel(t0, th1, th3, m, 1u << 0, 2u << 0, one);
      N.y = kab_label_sel(a0, a1, a2, a3, m, 1u << 2, 2u << 2, one);
      N.z = kab_blank_sel(t2, t1, th1, m, 1u << 4, 2u << 4, one);
      N.w = kab_label_sel(b0, b1, b2, b3, m, 1u << 6, 2u << 6, one);
      (void)t3;
      const int nlo = lo - vb, nhi = hi - vb;  // window in chunk coordinates
      if (nlo > 0 || nhi < 4) {                // edge / outside threads only
        if (0 < nlo || 0 >= nhi) N.x = ninf;
        if (1 < nlo || 1 >= nhi) N.y = ninf;
        if (2 < nlo || 2 >= nhi) N.z = ninf;
        if (3 < nlo || 3 >= nhi) N.w = ninf;
      }
      *reinterpret_cast<float4 *>(cur + 4 * tid) = N;
      word |= m << wsh;
      wsh += 8;
      if (wsh == 32) {
        *bprow = word;
        bprow += NT;
        word = 0; wsh = 0;
      }
      // next frame's window and emissions (off the barrier's critical path)
      q += qd; acc += rd;
      if (acc >= T) { acc -= T; ++q; }
      bool crossed = false;
      if (i + 1 < T) {
        if (++fin == F) {  // frame i+1 opens chunk cn+1
          fin = 0; ++cn; crossed = true;
          if (++st == KAB_BAND_STAGES) { st = 0; ph ^= 1u; }
          kab_mbar_wait(&ebars[st], ph);
          rowc = reinterpret_cast<const char *>(stage_base + st * stage_words + skew);
        } else {
          rowc += V * 4;
        }
        eb = *reinterpret_cast<const float *>(rowc);
        e1 = *reinterpret_cast<const float *>(rowc + c1);
        e3 = *reinterpret_cast<const float *>(rowc + c3);
        for (int c = tid; c < V; c += NT) bad |= !kab_finite(reinterpret_cast<const float *>(rowc)[c]);
      }
      __syncthreads();
      if (crossed && cn + KAB_BAND_STAGES - 1 < n_chunks) {
        // every thread has finished with chunk cn-1 (its last frame was frame i): refill its
        // stage (the one before `st`) with chunk cn + STAGES - 1
        issue(cn + KAB_BAND_STAGES - 1, (st + KAB_BAND_STAGES - 1) % KAB_BAND_STAGES);
      }
      float *tswap = prev; prev = cur; cur = tswap;
    }
    echunks = ec0 + n_chunks;
    if (T & 3) *bprow = word;

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
    __threadfence();          // backpointer words -> visible device-wide ...
    kab_fence_proxy_async();  // ... and to the async proxy that copies them back
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
      const int n_rows = (T + 3) >> 2;
      const int n_blocks = (n_rows + RPB - 1) / RPB;
      const uint32_t bb0 = bblocks;
      auto fetch = [&](int blk) {  // thread 0 only
        const uint32_t g = bb0 + (uint32_t)(n_blocks - 1 - blk), st = g & 1u;
        const int r0 = blk * RPB, nr = min(RPB, n_rows - r0);
        const uint32_t bytes = (uint32_t)nr * NT * 4;
        kab_fence_proxy_async();
        kab_mbar_expect_tx(&bbars[st], bytes);
        kab_bulk_g2s(bpblk + (size_t)st * KAB_BAND_BPBLOCK_BYTES, bpw + (size_t)r0 * NT, bytes, &bbars[st]);
      };
      if (tid == 0) fetch(n_blocks - 1);
      int32_t *out_path = p.best_path + lat.t_off;
      int32_t *out_lab = p.best_labels + lat.t_off;
      float *out_sc = p.best_scores + lat.t_off;
      const float *lp = p.lp + lat.t_off * (int64_t)V;
      for (int blk = n_blocks - 1; blk >= 0; --blk) {
        const uint32_t g = bb0 + (uint32_t)(n_blocks - 1 - blk), st = g & 1u;
        const int i0 = blk * FPB, i1 = min(T, i0 + FPB);
        if (tid == 0) {
          if (blk > 0) fetch(blk - 1);  // other buffer: its previous contents were consumed
          kab_mbar_wait(&bbars[st], (g >> 1) & 1u);
          const unsigned char *blkp = bpblk + (size_t)st * KAB_BAND_BPBLOCK_BYTES;
          for (int i = i1 - 1; i >= i0; --i) {
            const int slot = v & (R - 1);
            const unsigned char byte = blkp[(((i - i0) >> 2) * NT + (slot >> 2)) * 4 + (i & 3)];
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
