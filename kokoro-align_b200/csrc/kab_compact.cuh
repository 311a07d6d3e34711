// kab_compact.cuh -- wide vocabularies (V > 512, BASELINE config 5's V = 4096) for the staged
// kernels.  A lattice only ever reads the columns of its own labels and the blank
// (align.py:77: log_probs[i, ext[v]]), so when no lattice of the plan uses more than 511 distinct
// labels the plan works on a COMPACT copy of the log-probs: row t of lattice b keeps
// log_probs[t, gather[b][j]], j < Vc, with gather[b][0] = 0 (blank) and gather[b][1..] the
// lattice's distinct label columns in ascending order.  The staged kernels then run unchanged
// with V = Vc and labels renumbered to compact indices; best_labels is mapped back afterwards
// (a label value IS its column), best_scores are the same floats.  Non-finite values are
// therefore only rejected in columns the lattice uses -- like the reference, which never looks
// at the others.
#pragma once
#include "kab_common.cuh"

#define KAB_COMPACT_FRAMES 64  // frames per CTA of the two helper kernels

// grid (lattices of one work list, frame chunks); block 256 = 8 warps, a warp per frame
__global__ void kab_compact_kernel(const KabLattice *__restrict__ lats, const float *__restrict__ lp,
                                   float *__restrict__ lpc, const int32_t *__restrict__ gather, int V, int Vc) {
  const KabLattice lat = lats[blockIdx.x];
  const int f0 = blockIdx.y * KAB_COMPACT_FRAMES;
  if (f0 >= lat.T) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t *g = gather + (size_t)lat.index * Vc;
  const int f1 = min(lat.T, f0 + KAB_COMPACT_FRAMES);
  for (int f = f0 + warp; f < f1; f += 8) {
    const float *src = lp + (size_t)(lat.t_off + f) * V;
    float *dst = lpc + (size_t)(lat.t_off + f) * Vc;
    for (int j = lane; j < Vc; j += 32) dst[j] = __ldg(src + g[j]);
  }
}

// best_labels: compact index -> label value, for the lattices that were aligned (status 0)
__global__ void kab_expand_labels_kernel(const KabLattice *__restrict__ lats, int32_t *__restrict__ best_labels,
                                         const int32_t *__restrict__ status, const int32_t *__restrict__ gather,
                                         int Vc) {
  const KabLattice lat = lats[blockIdx.x];
  const int f0 = blockIdx.y * KAB_COMPACT_FRAMES;
  if (f0 >= lat.T || status[lat.index] != 0) return;
  const int32_t *g = gather + (size_t)lat.index * Vc;
  const int f1 = min(lat.T, f0 + KAB_COMPACT_FRAMES);
  for (int f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
    int32_t *p = best_labels + lat.t_off + f;
    const int32_t c = *p;
    if ((unsigned)c < (unsigned)Vc) *p = g[c];
  }
}
