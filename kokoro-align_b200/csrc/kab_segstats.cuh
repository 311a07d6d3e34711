// kab_segstats.cuh -- what the reference's align() (kokoro_align/align.py:127-169) takes from the
// three T-length arrays, computed where they are: per silence-split segment [audio_start, audio_end)
//     text_start       = best_path[audio_start] // 2                                  align.py:131,151
//     text_end         = best_path[audio_end] // 2  (or "end of transcript")           align.py:152
//     non_blanks       = np.sum(best_labels[a:b] != 0)                                 align.py:160
//     non_blanks_score = np.sum(best_scores[a:b][best_labels[a:b] != 0])               align.py:161
//     all_score        = np.sum(best_scores[a:b])                                      align.py:162
// so that a caller who only wants the *.align.txt numbers reads back 24 bytes per SEGMENT instead
// of 12 bytes per FRAME.
//
// The two sums are float32 np.sum over a contiguous 1-D array, i.e. numpy's pairwise summation
// (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum): n < 8 sequential; n <= 128 eight strided
// accumulators combined as a balanced tree, then the tail one by one; larger n split at
// n2 = n/2 - (n/2) % 8 and the halves added.  Reproduced operation for operation (every add is
// one __fadd_rn), so the printed numbers are the reference's, digit for digit.  The voiced scores
// are first compacted into a scratch array (np's boolean indexing makes that copy too): the
// pairwise tree depends on the COMPACTED length.
//
// One warp per segment: a ballot pass counts / compacts the voiced frames; the <= 128-element
// leaves of the summation tree go to the lanes round-robin (kab_np_rowsum), lane 0 combines them
// in the tree's post-order.  Segments with more leaves than KAB_SEG_LEAVES (> 14 000 frames) are
// summed by lane 0 alone -- correct, slower, and not a shape the reference's 3-15 s segments reach.
#pragma once
#include "kab_common.cuh"
#include "kab_softmax.cuh"  // kab_np_rowsum: numpy's pairwise sum of one block of <= 128 values

#define KAB_SEG_WARPS 4      // warps per CTA
#define KAB_SEG_LEAVES 256   // leaf sums kept per warp
#define KAB_SEG_BUF 2048     // frames of a segment staged in shared memory (per warp: scores + compacted voiced scores)
#define KAB_SEG_SMEM (KAB_SEG_WARPS * (2 * KAB_SEG_BUF + KAB_SEG_LEAVES) * 4)

// mirrors kab_segment_record (include/kokoro_align_b200.h)
struct KabSegmentRecord {
  int32_t text_start, text_end, non_blanks;
  float non_blanks_score, all_score;
  int32_t status;
};

// Leaves of numpy's pairwise tree over n elements are contiguous, in order, each <= 128 long.
// Calls leaf(k, offset, length) for every leaf k = 0, 1, ... ; returns their number.
template <int DEPTH = 48, class F>
__device__ __forceinline__ int kab_np_leaves(int64_t n, F leaf) {
  int64_t off_stack[DEPTH], len_stack[DEPTH];
  int sp = 0, k = 0;
  off_stack[0] = 0; len_stack[0] = n; sp = 1;
  while (sp > 0) {
    --sp;
    int64_t off = off_stack[sp], len = len_stack[sp];
    while (len > 128) {  // split; the right half waits on the stack
      int64_t n2 = len / 2;
      n2 -= n2 % 8;
      off_stack[sp] = off + n2; len_stack[sp] = len - n2; ++sp;
      len = n2;
    }
    leaf(k++, off, len);
  }
  return k;
}

// Post-order combination of the leaf sums (leafsum(k) = sum of leaf k), same tree as above.
template <int DEPTH = 48, class F>
__device__ __forceinline__ float kab_np_combine(int64_t n, F leafsum) {
  int64_t len_stack[DEPTH];
  float val_stack[DEPTH];
  bool have_left[DEPTH];
  int sp = 0, k = 0;
  int64_t cur = n;
  for (;;) {
    while (cur > 128) {
      int64_t n2 = cur / 2;
      n2 -= n2 % 8;
      len_stack[sp] = cur - n2; have_left[sp] = false; ++sp;
      cur = n2;
    }
    float v = leafsum(k++);
    for (;;) {
      if (sp == 0) return v;
      if (!have_left[sp - 1]) {  // v is the left half: evaluate the right one
        val_stack[sp - 1] = v; have_left[sp - 1] = true;
        cur = len_stack[sp - 1];
        break;
      }
      v = __fadd_rn(val_stack[sp - 1], v);
      --sp;
    }
  }
}

// The leaves of the tree one at a time, left to right (the order kab_np_combine consumes them in).
struct KabNpLeafWalk {
  int64_t off_stack[48], len_stack[48];
  int sp;
  __device__ __forceinline__ void init(int64_t n) { off_stack[0] = 0; len_stack[0] = n; sp = 1; }
  __device__ __forceinline__ void next(int64_t &off, int64_t &len) {
    --sp;
    off = off_stack[sp]; len = len_stack[sp];
    while (len > 128) {
      int64_t n2 = len / 2;
      n2 -= n2 % 8;
      off_stack[sp] = off + n2; len_stack[sp] = len - n2; ++sp;
      len = n2;
    }
  }
};

// np.sum(a[0:n]) of a segment staged in SHARED memory (n <= KAB_SEG_BUF: at most 37 leaves, a tree of
// depth <= 6), by one warp; every lane returns the result.
__device__ __forceinline__ float kab_np_sum_warp_staged(const float *a, int n, float *leafbuf, int lane) {
  float res = 0.0f;
  if (n <= 0) return res;
  kab_np_leaves<8>(n, [&](int k, int64_t off, int64_t len) {
    if ((k & 31) == lane) leafbuf[k] = kab_np_rowsum((int)len, [&](int i) { return a[off + i]; });
  });
  __syncwarp();
  if (lane == 0) res = kab_np_combine<8>(n, [&](int k) { return leafbuf[k]; });
  return __shfl_sync(KAB_FULL_MASK, res, 0);
}

// np.sum(a[0:n]) in float32, by one warp; every lane returns the result.  leafbuf: KAB_SEG_LEAVES floats.
__device__ __forceinline__ float kab_np_sum_warp(const float *a, int64_t n, float *leafbuf, int lane) {
  float res = 0.0f;
  if (n <= 0) return res;
  // a leaf of a split array holds at least 57 elements (n2 >= 64 - 7)
  if (n / 56 + 1 <= KAB_SEG_LEAVES) {
    kab_np_leaves(n, [&](int k, int64_t off, int64_t len) {
      if ((k & 31) == lane) leafbuf[k] = kab_np_rowsum((int)len, [&](int i) { return __ldcg(a + off + i); });
    });
    __syncwarp();
    if (lane == 0) res = kab_np_combine(n, [&](int k) { return leafbuf[k]; });
  } else if (lane == 0) {  // very long segment: one lane, leaves evaluated on demand in tree order
    KabNpLeafWalk w;
    w.init(n);
    res = kab_np_combine(n, [&](int) {
      int64_t off, len;
      w.next(off, len);
      return kab_np_rowsum((int)len, [&](int i) { return __ldcg(a + off + i); });
    });
  }
  return __shfl_sync(KAB_FULL_MASK, res, 0);
}

// seg_lat_off[B+1]: segments of lattice b are seg_lat_off[b] .. seg_lat_off[b+1]-1; seg_end[s]: the
// reference's `indices` (cumulative segment ends, frames relative to the lattice), concatenated.
__global__ void __launch_bounds__(KAB_SEG_WARPS * 32)
kab_segment_stats_kernel(int64_t n_seg, int64_t B, const int64_t *__restrict__ seg_lat_off,
                         const int64_t *__restrict__ seg_end, const int64_t *__restrict__ t_off,
                         const int32_t *__restrict__ path, const int32_t *__restrict__ labs,
                         const float *__restrict__ scores, const int32_t *__restrict__ status,
                         float *__restrict__ scratch, KabSegmentRecord *__restrict__ rec) {
  // per warp: the segment's scores, its voiced scores compacted, the leaf sums
  extern __shared__ __align__(16) float kab_seg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *s_all = kab_seg_smem + (size_t)warp * (2 * KAB_SEG_BUF + KAB_SEG_LEAVES);
  float *s_comp = s_all + KAB_SEG_BUF;
  float *leafbuf = s_comp + KAB_SEG_BUF;
  const int64_t n_warps = (int64_t)gridDim.x * KAB_SEG_WARPS;
  for (int64_t s = (int64_t)blockIdx.x * KAB_SEG_WARPS + warp; s < n_seg; s += n_warps) {
    // lattice of segment s: the last b with seg_lat_off[b] <= s
    int64_t lo = 0, hi = B;
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (seg_lat_off[mid] <= s) lo = mid; else hi = mid;
    }
    const int64_t b = lo, r0 = t_off[b], T = t_off[b + 1] - r0;
    const int64_t a = s == seg_lat_off[b] ? 0 : seg_end[s - 1], e = seg_end[s];
    KabSegmentRecord r;
    r.status = status[b];
    r.text_start = 0; r.text_end = -1; r.non_blanks = 0; r.non_blanks_score = 0.0f; r.all_score = 0.0f;
    if (r.status == 0 && a >= 0 && a < T) {
      r.text_start = path[r0 + a] >> 1;
      r.text_end = (e >= 0 && e < T) ? path[r0 + e] >> 1 : -1;  // -1: "len(aligner)", align.py:152
      const int64_t ec = e < a ? a : (e > T ? T : e);           // python slice [a:e]
      const int64_t n = ec - a;
      if (n <= KAB_SEG_BUF) {
        // the usual case (the reference's segments are 3-15 s, 300-1500 frames): one coalesced pass over
        // the frames into shared memory -- every leaf of both summation trees then reads shared memory
        // instead of walking global memory one dependent L2 round trip per eight elements (189 us for
        // the 10 000 segments of config 2 before)
        int cnt = 0;
#pragma unroll 4
        for (int i = 0; i < (int)n; i += 32) {
          const bool in = i + lane < (int)n;
          const float sc = in ? scores[r0 + a + i + lane] : 0.0f;
          const bool voiced = in && labs[r0 + a + i + lane] != 0;
          const unsigned m = __ballot_sync(KAB_FULL_MASK, voiced);
          if (in) s_all[i + lane] = sc;
          if (voiced) s_comp[cnt + __popc(m & ((1u << lane) - 1u))] = sc;
          cnt += __popc(m);
        }
        __syncwarp();
        r.non_blanks = cnt;
        r.all_score = kab_np_sum_warp_staged(s_all, (int)n, leafbuf, lane);
        __syncwarp();
        r.non_blanks_score = kab_np_sum_warp_staged(s_comp, cnt, leafbuf, lane);
        __syncwarp();
        if (lane == 0) rec[s] = r;
        continue;
      }
      // voiced frames compacted into scratch[r0 + a ...] (this warp's own rows)
      float *comp = scratch + r0 + a;
      int64_t cnt = 0;
      for (int64_t i = 0; i < n; i += 32) {
        const bool in = i + lane < n;
        const bool voiced = in && labs[r0 + a + i + lane] != 0;
        const unsigned m = __ballot_sync(KAB_FULL_MASK, voiced);
        if (voiced) comp[cnt + __popc(m & ((1u << lane) - 1u))] = scores[r0 + a + i + lane];
        cnt += __popc(m);
      }
      __syncwarp();
      r.non_blanks = (int32_t)cnt;
      r.all_score = kab_np_sum_warp(scores + r0 + a, n, leafbuf, lane);
      __syncwarp();
      r.non_blanks_score = kab_np_sum_warp(comp, cnt, leafbuf, lane);
      __syncwarp();
    } else if (r.status == 0) {
      r.status = -1;  // audio_start outside the lattice: the reference raises IndexError (align.py:151)
    }
    if (lane == 0) rec[s] = r;
  }
}

// best_labels as bytes (the `decoded` column of align() needs the per-frame labels; V <= 256)
__global__ void kab_labels_u8_kernel(const int32_t *__restrict__ labs, uint8_t *__restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n && (reinterpret_cast<uintptr_t>(labs + i) & 15) == 0 && (reinterpret_cast<uintptr_t>(out + i) & 3) == 0) {
      const int4 v = *reinterpret_cast<const int4 *>(labs + i);
      *reinterpret_cast<uchar4 *>(out + i) = make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z, (unsigned char)v.w);
    } else {
      for (int64_t k = i; k < n && k < i + 4; ++k) out[k] = (uint8_t)labs[k];
    }
  }
}
