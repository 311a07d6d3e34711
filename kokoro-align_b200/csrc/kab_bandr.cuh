// kab_bandr.cuh -- warp-specialised cluster band kernel: the recurrence warps do NOTHING but the
// recurrence.  Same shapes, ring, lanes and backpointer words as kab_bandq.cuh (two states per
// lane, 12 ghost lanes, 40 owned ring slots per warp, 8 compute warps per CTA, clusters of
// NC <= 8 CTAs; KabBtLayoutQ for the traceback).
//
// Why (measured with the KAB_BANDQ_TIMING build and tools/ubench/chain.cu on B200): the frame of
// the two-states-per-lane update is a dependent chain SHFL.UP (24.9 cycles) -> FADD2 (4.7) ->
// FMNMX3 (4.3); the whole value update runs at 43-46 cycles per frame, alone or with two warps
// per scheduler.  kab_bandq.cuh spends ~2 230 cycles per 8-frame group and warp, i.e. 280 per
// frame: the warps of a lattice form a chain in time whose pace is the SERIAL work of its slowest
// member per group, and only ~400 of those cycles are frames -- the rest is the warp's own
// bookkeeping on the same instruction stream: emission gathers for the next group (16 LDS +
// addresses), the exact window turned into masks (edge warps: +270), recycling, mbarrier waits,
// neighbour messages through L2 (loads / polls of 300-700 cycles), credit checks, backpointer
// staging and bulk stores.  Two warps per scheduler do not help a chain: its links are serial.
//
// Here every compute warp has a PREP warp that runs up to KAB_BR_TD groups ahead and hands it
// finished emission tiles in shared memory -- gathered, with the window of align.py:64-65
// already applied as -inf -- so the compute warp's group is: check the neighbour's message (a
// shared-memory word), 8 x (one LDS.64 + the 19-instruction frame), publish its top 12 lanes to
// the warp above (shared memory, or distributed shared memory across the CTA boundary: one
// st.shared::cluster per (score, seq) word, no mbarrier, no L2 round trip), one STS of the
// backpointer word.  The prep warp also owns the ring bookkeeping (aliases, recycling, label
// columns), releases the emission stages, and issues the bulk stores of its compute warp's
// backpointer blocks.  A producer warp stages emission rows by bulk copies and checks finiteness.
#pragma once
#include "kab_band.cuh"
#include "kab_bandp.cuh"
#include "kab_bandq.cuh"
#include "kab_common.cuh"

#define KAB_BR_CW KAB_BQ_CW   // compute warps per CTA (and as many prep warps)
#define KAB_BR_GH KAB_BQ_GH   // ghost lanes
#define KAB_BR_OW KAB_BQ_OW   // owned ring slots per warp
#ifndef KAB_BR_NS
#define KAB_BR_NS 40          // emission stages per CTA (head and tail of the chain can sit in one CTA)
#endif
#define KAB_BR_TD 4           // emission tiles per compute warp (groups the prep warp may run ahead)
#define KAB_BR_MD 16          // mailbox depth (messages)
#define KAB_BR_LAG 2          // a warp joining the chain lets its lower neighbour get this many groups ahead
#define KAB_BR_BG 16          // groups per backpointer block (2 KB)
#define KAB_BR_THREADS ((2 * KAB_BR_CW + 1) * 32)

struct KabBandrGeom {
  size_t ctrl_off, tile_off, mbox_off, bp_off, vbf_off, stage_off, smem_bytes;
};
__host__ __device__ inline KabBandrGeom kab_bandr_geom(int stage_bytes) {
  KabBandrGeom g;
  g.ctrl_off = ((size_t)2 * KAB_BR_NS * 8 + 64 + 127) & ~(size_t)127;      // after the mbarriers and CTA scalars
  g.tile_off = g.ctrl_off + (size_t)KAB_BR_CW * 128;                       // one 128-byte control block per compute warp
  g.mbox_off = g.tile_off + (size_t)KAB_BR_CW * KAB_BR_TD * 8 * 32 * 8;    // tiles [w][TD][8 frames][32 lanes] float2
  g.bp_off = g.mbox_off + (size_t)KAB_BR_CW * KAB_BR_MD * KAB_BR_GH * 16;  // mailboxes [w][MD][12 lanes][2] (score, seq)
  g.vbf_off = g.bp_off + (size_t)KAB_BR_CW * 2 * KAB_BR_BG * 128;          // backpointer staging [w][2][BG][32] u32
  g.stage_off = g.vbf_off + (size_t)KAB_BR_CW * 32 * 4;                    // final alias of every lane
  g.smem_bytes = g.stage_off + (size_t)KAB_BR_NS * stage_bytes;
  return g;
}

// control block of compute warp w (u32 words, shared memory)
#define KAB_BR_C_TILESEQ 0    // [TD] tile t holds group g  <=>  word == g + 1          (prep -> compute)
#define KAB_BR_C_TILEFLG 4    // [TD] bit 0: the ghost lanes need the neighbour's message  (prep -> compute)
#define KAB_BR_C_COMPDONE 8   // groups whose tile the compute warp has finished reading   (compute -> prep)
#define KAB_BR_C_MBOXDONE 9   // messages 0 .. n-1 are consumed                            (compute -> lower neighbour)
#define KAB_BR_C_BPREADY 10   // backpointer blocks staged                                 (compute -> prep)
#define KAB_BR_C_BPFREE 11    // backpointer blocks whose staging buffer is free again     (prep -> compute)

__device__ __forceinline__ uint32_t kab_lds_relaxed_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.relaxed.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void kab_sts_relaxed_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.relaxed.cluster.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 kab_lds_relaxed_b64(uint32_t addr) {
  uint2 v;
  asm volatile(
      "{\n\t.reg .b64 q;\n\t"
      "ld.relaxed.cluster.shared::cta.b64 q, [%2];\n\t"
      "mov.b64 {%0, %1}, q;\n\t}"
      : "=r"(v.x), "=r"(v.y)
      : "r"(addr)
      : "memory");
  return v;
}
// (score, seq) as ONE 64-bit store into the shared memory of any CTA of the cluster (addr from mapa)
__device__ __forceinline__ void kab_st_cluster_b64(uint32_t addr, uint32_t lo, uint32_t hi) {
  asm volatile(
      "{\n\t.reg .b64 q;\n\t"
      "mov.b64 q, {%1, %2};\n\t"
      "st.relaxed.cluster.shared::cluster.b64 [%0], q;\n\t}" ::"r"(addr), "r"(lo), "r"(hi)
      : "memory");
}
__device__ __forceinline__ uint32_t kab_ld_cluster_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.relaxed.cluster.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void kab_fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

__global__ void __launch_bounds__(KAB_BR_THREADS, 1)
    kab_bandr_kernel(const KabLattice *__restrict__ lats, int n_lat, KabParams p) {
  constexpr int G = KAB_BAND_G, GH = KAB_BR_GH, OW = KAB_BR_OW;
  constexpr int CW = KAB_BR_CW, NS = KAB_BR_NS, TD = KAB_BR_TD, MD = KAB_BR_MD, BG = KAB_BR_BG;
  static_assert(G == 8, "a group of 8 frames is one 32-bit backpointer word per lane");
  const KabBandrGeom geo = kab_bandr_geom(p.stage_bytes);
  extern __shared__ __align__(128) unsigned char kab_smem[];
  uint64_t *efull = reinterpret_cast<uint64_t *>(kab_smem);  // [NS]
  uint64_t *eempty = efull + NS;                             // [NS]
  unsigned int *s_item = reinterpret_cast<unsigned int *>(eempty + NS);
  int *s_vmax = reinterpret_cast<int *>(s_item + 1);
  unsigned int *s_bad = s_item + 2;
  float *stage_base = reinterpret_cast<float *>(kab_smem + geo.stage_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = kab_cluster_rank(), NC = kab_cluster_size();
  const int NWT = CW * (int)NC, R = OW * NWT;
  const bool is_prod = warp == 2 * CW, is_prep = warp >= CW && warp < 2 * CW;
  const int cw = is_prep ? warp - CW : warp;  // the compute warp this warp is / serves
  const int gw = (int)rank * CW + cw;         // its global index in the ring
  const bool owned = lane >= GH;
  // ring slot of this lane's blank state: owned lanes tile the warp's 40 slots, ghost lanes mirror
  // the previous warp's lanes 20..31
  const int slot0 = owned ? OW * gw + 2 * (lane - GH) : (OW * gw - 2 * GH + 2 * lane + R) % R;
  const float ninf = kab_neg_inf();
  const uint32_t smem0 = kab_smem_u32(kab_smem);
  const uint32_t ctrl = smem0 + (uint32_t)geo.ctrl_off + (uint32_t)cw * 128u;       // this pair's control block
  const uint32_t mbox = smem0 + (uint32_t)geo.mbox_off + (uint32_t)cw * (MD * GH * 16u);
  const uint32_t bpst = smem0 + (uint32_t)geo.bp_off + (uint32_t)cw * (2u * BG * 128u);
  int *vbf = reinterpret_cast<int *>(kab_smem + geo.vbf_off) + cw * 32;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      kab_mbar_init(&efull[s], 1);
      kab_mbar_init(&eempty[s], CW);  // the CW prep warps release a stage
    }
    kab_fence_mbar_init();
  }
  __syncthreads();
  uint32_t echunks = 0;  // emission chunks staged so far by this CTA (same count in every warp)

  for (;;) {
    // ---- reset the per-lattice words (sequence numbers restart at 1), then the cluster agrees on
    // the next work item; the cluster barrier also keeps remote mailbox stores of the new lattice
    // behind every CTA's reset
    for (uint32_t o = geo.ctrl_off + tid * 4; o < geo.tile_off; o += KAB_BR_THREADS * 4)
      *reinterpret_cast<uint32_t *>(kab_smem + o) = 0u;
    for (uint32_t o = geo.mbox_off + tid * 4; o < geo.bp_off; o += KAB_BR_THREADS * 4)
      *reinterpret_cast<uint32_t *>(kab_smem + o) = 0u;
    if (tid == 0) {
      *s_vmax = -1;
      *s_bad = 0u;
    }
    if (rank == 0 && tid == 0) {
      const unsigned int it = atomicAdd(p.queue, 1u);
      for (uint32_t r = 0; r < NC; ++r) kab_st_cluster_u32(kab_mapa(kab_smem_u32(s_item), r), it);
    }
    __syncthreads();
    kab_cluster_sync();
    const unsigned int item = *s_item;
    if (item >= (unsigned int)n_lat) break;
    const KabLattice lat = lats[item];
    const int T = lat.T, S = 2 * lat.L + 1, V = p.V, W = p.W;
    const int F = p.stage_frames;
    const uint32_t stage_words = p.stage_bytes >> 2;
    const int n_chunks = (T + F - 1) / F;
    const int n_groups = (T + G - 1) / G;
    const int n_blocks = (n_groups + BG - 1) / BG;
    const uint32_t ec0 = echunks;
    const uint32_t skew = (uint32_t)(((lat.t_off * (int64_t)V * 4) & 15) >> 2);
    float s0 = ninf, s1 = ninf;  // compute warps: blank state vb, label state vb + 1
    int vb = slot0;              // prep warps track the alias; compute warps read the final one from vbf

    if (is_prod) {
      // ================= producer warp: emission ring + finiteness of the staged rows
      const char *lp_base = reinterpret_cast<const char *>(p.lp) + ((lat.t_off * (int64_t)V * 4) & ~(int64_t)15);
      const uint32_t chunk_stride = (uint32_t)(F * V * 4);
      const uint32_t full_bytes = (chunk_stride + skew * 4 + 15) & ~15u;
      float poison = 0.0f;
      auto check_chunk = [&](int c) {  // waits for chunk c, then scans it (the prep warps may be reading it too)
        const uint32_t gc = ec0 + (uint32_t)c, stg = gc % NS;
        kab_mbar_wait(&efull[stg], (gc / NS) & 1u);
        const float *w = stage_base + stg * stage_words + skew;
        const int nw = min(F, T - c * F) * V;
        for (int j = lane; j < nw; j += 32) poison = kab_poison(poison, w[j]);
      };
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t gc = ec0 + (uint32_t)c, stg = gc % NS, use = gc / NS;
        if (c >= NS) check_chunk(c - NS);  // the chunk that used this stage is scanned before the stage is given away
        if (use > 0) kab_mbar_wait(&eempty[stg], (use - 1u) & 1u);  // all prep warps released it
        float *dst = stage_base + stg * stage_words;
        if (c + 1 < n_chunks) {
          if (lane == 0) {
            kab_mbar_expect_tx(&efull[stg], full_bytes);
            kab_bulk_g2s(dst, lp_base + (size_t)c * chunk_stride, full_bytes, &efull[stg]);
          }
        } else {
          const int f0 = c * F, nf = T - f0;
          const KabStageDesc d = kab_stage_desc(p, lat.t_off, f0, nf);
          if (lane < (int)d.tail_n)
            dst[d.tail_word + lane] = __ldg(reinterpret_cast<const float *>(d.src) + d.tail_word + lane);
          __syncwarp();
          if (lane == 0) {
            kab_mbar_expect_tx(&efull[stg], d.bytes);  // (release: the tail words above are ordered before it)
            if (d.bytes) kab_bulk_g2s(dst, d.src, d.bytes, &efull[stg]);
          }
        }
        __syncwarp();
      }
      for (int c = max(0, n_chunks - NS); c < n_chunks; ++c) check_chunk(c);
      if (__any_sync(KAB_FULL_MASK, poison != poison) && lane == 0)
        for (uint32_t rr = 0; rr < NC; ++rr) kab_red_or_cluster_u32(kab_mapa(kab_smem_u32(s_bad), rr), 1u);
    } else if (is_prep) {
      // ================= prep warp of compute warp cw: emission tiles, ring bookkeeping, backpointer stores
      const uint16_t *col16 = p.col16 + lat.col_off;
      auto load_col = [&](int base) -> uint32_t { return base + 1 < S ? 4u * col16[base >> 1] : 0u; };
      uint32_t c1 = load_col(vb), nc1 = load_col(vb + R);  // byte offset of the label column; next alias prefetched
      const int half = W / 2;
      const int VB = V * 4;
      uint32_t st = ec0 % NS, ph = (ec0 / NS) & 1u;  // stage / phase of the chunk being read
      auto chunk_ptr = [&](uint32_t stg) { return reinterpret_cast<const char *>(stage_base + stg * stage_words + skew); };
      const int qd = S / T, rd = S % T;
      const int qdg = (int)(((int64_t)S * G) / T), rdg = (int)(((int64_t)S * G) % T);
      int qg = 0, rg = 0;
      int fic = 0;      // frame offset of the current group inside its emission chunk
      int lo_prev = 0;  // lo of the first frame of the previous group (<= lo of every later frame)
      unsigned char *bpg = p.bp + lat.bp_off + (size_t)gw * n_groups * 128;  // the compute warp's region of the workspace
      int bp_issued = 0;  // backpointer blocks handed to the bulk-copy engine
      // backpointer blocks the compute warp has staged -> bulk stores (lane 0 issues, the warp follows)
      auto service_bp = [&]() {  // (warp-uniform: lane 0 looks, everybody follows)
        int ready = lane == 0 ? (int)kab_lds_relaxed_u32(ctrl + 4 * KAB_BR_C_BPREADY) : 0;
        ready = __shfl_sync(KAB_FULL_MASK, ready, 0);
        if (ready > bp_issued) {  // (at most one new block per visit: a block is 16 groups long)
          kab_fence_cta();
          const int b = bp_issued;
          const int ng = min(BG, n_groups - b * BG);
          if (lane == 0) {
            kab_bulk_s2g(bpg + (size_t)b * BG * 128, kab_smem + geo.bp_off + (size_t)cw * (2 * BG * 128) + (size_t)(b & 1) * BG * 128,
                         (uint32_t)ng * 128u);
            kab_bulk_wait_read1();  // the block before this one has left its buffer
            kab_sts_relaxed_u32(ctrl + 4 * KAB_BR_C_BPFREE, (uint32_t)b);  // blocks 0 .. b-1 are free
          }
          bp_issued = b + 1;
          __syncwarp();
        }
      };
      kab_mbar_spin(&efull[st], ph);
      const char *rowc = chunk_ptr(st);  // first row of the current group
      for (int g = 0; g < n_groups; ++g) {
        const int i0 = g * G, nfr = min(G, T - i0);
        const bool more = i0 + G < T;
        const int t = g % TD;
        // ---- the tile slot is free once the compute warp has finished group g - TD
        if (g >= TD) {
          for (;;) {
            int done = lane == 0 ? (int)kab_lds_relaxed_u32(ctrl + 4 * KAB_BR_C_COMPDONE) : 0;
            done = __shfl_sync(KAB_FULL_MASK, done, 0);
            if (done >= g - TD + 1) break;
            service_bp();
          }
        }
        kab_fence_cta();
        // ---- does the compute warp need its neighbour's message for this group?  (same test as
        // kab_bandq.cuh: not if all 24 ghost states are outside the window for the whole group)
        uint32_t flags = 0;
        if (g > 0) {
          int qn2 = qg + qdg;
          if (rg + rdg >= T) ++qn2;
          const int hi1g = min(max(0, qn2 - half) + W, S);  // >= hi of every frame of this group
          const bool outside = owned || vb + 1 < lo_prev || vb >= hi1g;
          flags = __all_sync(KAB_FULL_MASK, outside) ? 0u : 1u;
        }
        lo_prev = max(0, qg - half);
        const int lo0 = lo_prev, hi0 = min(lo0 + W, S);
        int qn = qg + qdg, rn = rg + rdg;
        if (rn >= T) { rn -= T; ++qn; }
        const int lo1 = max(0, qn - half);
        while (vb + 1 < lo0 - 3) {  // recycle a chunk that fell below the window (between groups only)
          vb += R;
          c1 = nc1;
          nc1 = load_col(vb + R);
        }
        const bool safe = __all_sync(KAB_FULL_MASK, nfr == G && vb >= lo1 && vb + 2 <= hi0);
        float2 *tile = reinterpret_cast<float2 *>(kab_smem + geo.tile_off + (size_t)cw * (TD * 2048) + (size_t)t * 2048) + lane;
        if (safe) {
#pragma unroll
          for (int f = 0; f < G; ++f)
            tile[f * 32] = make_float2(*reinterpret_cast<const float *>(rowc + f * VB),
                                       *reinterpret_cast<const float *>(rowc + f * VB + c1));
        } else {
          // edge warp (or the last, partial group): the exact per-frame window (S*i = q*T + r, no
          // divisions) turned into masked emissions
          int q = qg, r = rg;
#pragma unroll
          for (int f = 0; f < G; ++f) {
            const int lo = max(0, q - half);   // align.py:64
            const int hi = min(lo + W, S);     // align.py:65
            q += qd; r += rd;
            if (r >= T) { r -= T; ++q; }
            const unsigned a = (unsigned)(vb - lo), wd = (unsigned)(hi - lo);
            float xb = ninf, x1 = ninf;
            if (f < nfr) {
              xb = *reinterpret_cast<const float *>(rowc + f * VB);
              x1 = *reinterpret_cast<const float *>(rowc + f * VB + c1);
            }
            tile[f * 32] = make_float2((a + 0u < wd) ? xb : ninf, (a + 1u < wd) ? x1 : ninf);
          }
        }
        if (!more) vbf[lane] = vb;  // final alias of this lane (the compute warp's forced end state)
        qg = qn; rg = rn;
        __syncwarp();
        kab_fence_cta();
        if (lane == 0) {
          kab_sts_relaxed_u32(ctrl + 4 * (KAB_BR_C_TILEFLG + t), flags);
          kab_sts_relaxed_u32(ctrl + 4 * (KAB_BR_C_TILESEQ + t), (uint32_t)(g + 1));
        }
        // ---- emission chunk finished?
        const bool next_crosses = fic + G == F;
        if (next_crosses || !more) {
          const uint32_t nst = st + 1 == NS ? 0 : st + 1;
          const uint32_t nph = nst == 0 ? ph ^ 1u : ph;
          __syncwarp();
          if (lane == 0) kab_mbar_arrive(&eempty[st]);
          st = nst; ph = nph;
          fic = 0;
          if (more) {
            kab_mbar_spin(&efull[st], ph);
            rowc = chunk_ptr(st);
          }
        } else {
          fic += G;
          rowc += G * VB;
        }
        service_bp();
      }
      while (bp_issued < n_blocks) service_bp();
      if (lane == 0) kab_bulk_wait0();  // the compute warp's backpointer blocks are in global memory
    } else {
      // ================= compute warp: the recurrence
      const uint32_t one = p.one;
      if (owned && slot0 == 0) s0 = 0.0f;  // virtual start state 0, score 0 (align.py:57-58)
      uint32_t bw = 0;  // backpointer nibbles of the current group
      auto frame = [&](const float xb, const float x1, const int sh) {
        const float h1 = __shfl_up_sync(KAB_FULL_MASK, s1, 1);  // state vb - 1
        const float h2 = __shfl_up_sync(KAB_FULL_MASK, s0, 1);  // state vb - 2
        const float h3 = __shfl_up_sync(KAB_FULL_MASK, s1, 2);  // state vb - 3
        float t0, th1, a0, a1, a2, a3;
        kab_add2(s0, h1, xb, t0, th1);        // blank <- vb (move 0), vb - 1 (move 1)
        const float th3 = __fadd_rn(h3, xb);  // vb - 3 (move 3)
        kab_add2(s0, s1, x1, a1, a0);         // label <- vb + 1 (move 0), vb (move 1)
        kab_add2(h2, h1, x1, a3, a2);         //       <- vb - 1 (move 2), vb - 2 (move 3)
        const float m0 = kab_blank_sel(t0, th1, th3, bw, 1u << (sh + 0), 2u << (sh + 0), one);
        const float m1 = kab_label_sel(a0, a1, a2, a3, bw, 1u << (sh + 2), 2u << (sh + 2), one);
        s0 = m0; s1 = m1;
      };
      // mailboxes: mine (messages of the warp below), and the one of the warp above -- in this CTA,
      // or in the next CTA of the cluster (distributed shared memory)
      const bool remote_up = cw == CW - 1;
      const uint32_t up_rank = remote_up ? (rank + 1 == NC ? 0u : rank + 1u) : rank;
      const int up_cw = remote_up ? 0 : cw + 1;
      const uint32_t up_mbox_local = smem0 + (uint32_t)geo.mbox_off + (uint32_t)up_cw * (MD * GH * 16u);
      const uint32_t up_ctrl_local = smem0 + (uint32_t)geo.ctrl_off + (uint32_t)up_cw * 128u;
      const uint32_t up_mbox = kab_mapa(up_mbox_local, up_rank) + (uint32_t)(lane >= 32 - GH ? lane - (32 - GH) : 0) * 16u;
      const uint32_t up_done = kab_mapa(up_ctrl_local + 4 * KAB_BR_C_MBOXDONE, up_rank);
      const uint32_t inbox = mbox + (uint32_t)(lane < GH ? lane : 0) * 16u;
      uint32_t cons_seen = 0;   // messages the warp above is known to be done with
      bool was_needed = false;  // the previous group read its message (the warp is inside the chain)
      uint2 pf0 = make_uint2(0, 0), pf1 = pf0;  // message g-1, loaded a group early
      int gib = 0, blk = 0;     // group inside the current backpointer block, block index
      for (int g = 0; g < n_groups; ++g) {
        const int i0 = g * G, nfr = min(G, T - i0);
        const bool more = i0 + G < T;
        const int t = g % TD;
        // ---- this group's tile
        for (;;) {  // (warp-uniform poll)
          uint32_t sq = lane == 0 ? kab_lds_relaxed_u32(ctrl + 4 * (KAB_BR_C_TILESEQ + t)) : 0u;
          sq = __shfl_sync(KAB_FULL_MASK, sq, 0);
          if (sq == (uint32_t)(g + 1)) break;
        }
        kab_fence_cta();
        const bool need = (kab_lds_relaxed_u32(ctrl + 4 * (KAB_BR_C_TILEFLG + t)) & 1u) != 0u;
        // ---- ghost lanes: the lower neighbour's top 24 states after its group g-1 (message g-1)
        if (g > 0) {
          if (need) {
            if (!owned) {
              if (!was_needed) {  // (re)joining the chain: let the warp below get KAB_BR_LAG groups ahead
                const int mt = min(g - 1 + KAB_BR_LAG - 1, n_groups - 2);
                const uint32_t ls = inbox + (uint32_t)(mt % MD) * (GH * 16u);
                while (kab_lds_relaxed_b64(ls + 8).y != (uint32_t)(mt + 1)) {
                }
              }
              const uint32_t slot = inbox + (uint32_t)((g - 1) % MD) * (GH * 16u);
              const uint32_t seq = (uint32_t)g;
              while (pf0.y != seq || pf1.y != seq) {
                pf0 = kab_lds_relaxed_b64(slot);
                pf1 = kab_lds_relaxed_b64(slot + 8);
              }
              s0 = __uint_as_float(pf0.x);
              s1 = __uint_as_float(pf1.x);
            }
          } else if (!owned) {
            s0 = ninf; s1 = ninf;
          }
          was_needed = need;
          __syncwarp();
          if (lane == 0) kab_sts_relaxed_u32(ctrl + 4 * KAB_BR_C_MBOXDONE, (uint32_t)g);  // done with messages 0 .. g-1
        }
        // message g (for the next group) may already be there: load it now, check it then
        if (more && !owned) {
          const uint32_t slot = inbox + (uint32_t)(g % MD) * (GH * 16u);
          pf0 = kab_lds_relaxed_b64(slot);
          pf1 = kab_lds_relaxed_b64(slot + 8);
        }
        // ---- the frames
        const float2 *tile = reinterpret_cast<const float2 *>(kab_smem + geo.tile_off + (size_t)cw * (TD * 2048) + (size_t)t * 2048) + lane;
        float2 e[G];  // (plain loads between the two fences: the compiler schedules them ahead of the frames)
#pragma unroll
        for (int f = 0; f < G; ++f) e[f] = tile[f * 32];
        bw = 0;
        if (nfr == G) {
#pragma unroll
          for (int f = 0; f < G; ++f) frame(e[f].x, e[f].y, 4 * f);
        } else {
#pragma unroll
          for (int f = 0; f < G; ++f)
            if (f < nfr) frame(e[f].x, e[f].y, 4 * f);
        }
        kab_fence_cta();  // the tile has been read (its values are in the scores) before the slot is given back
        if (lane == 0) kab_sts_relaxed_u32(ctrl + 4 * KAB_BR_C_COMPDONE, (uint32_t)(g + 1));
        // ---- hand the top twelve lanes to the warp above (message g)
        if (more) {
          if (g >= MD && (uint32_t)(g - MD) >= cons_seen) {  // about to lap the consumer: read its progress
            do {
              cons_seen = kab_ld_cluster_u32(up_done);
            } while ((uint32_t)(g - MD) >= cons_seen);
          }
          if (lane >= 32 - GH) {
            const uint32_t slot = up_mbox + (uint32_t)(g % MD) * (GH * 16u);
            const uint32_t seq = (uint32_t)(g + 1);
            kab_st_cluster_b64(slot, __float_as_uint(s0), seq);
            kab_st_cluster_b64(slot + 8, __float_as_uint(s1), seq);
          }
        }
        // ---- backpointer word of this group -> staging; block finished?
        if (gib == 0 && blk >= 2) {  // the buffer of block blk - 2 must have left shared memory
          for (;;) {
            int fr = lane == 0 ? (int)kab_lds_relaxed_u32(ctrl + 4 * KAB_BR_C_BPFREE) : 0;
            fr = __shfl_sync(KAB_FULL_MASK, fr, 0);
            if (fr >= blk - 1) break;
          }
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(bpst + (uint32_t)((blk & 1) * BG + gib) * 128u + (uint32_t)lane * 4u), "r"(bw) : "memory");
        ++gib;
        if (gib == BG || !more) {
          kab_fence_proxy_async_smem();  // every lane's words -> visible to the bulk store
          kab_fence_cta();
          __syncwarp();
          ++blk;
          gib = 0;
          if (lane == 0) kab_sts_relaxed_u32(ctrl + 4 * KAB_BR_C_BPREADY, (uint32_t)blk);
        }
      }
      // ---- end of the forward pass: cluster-wide forced end state (align.py:99-101)
      kab_fence_cta();
      vb = vbf[lane];  // (written by the prep warp before it published the last tile)
      int cand = -1;
      if (owned) {
        if (vb + 0 < S && s0 > ninf) cand = vb + 0;
        if (vb + 1 < S && s1 > ninf) cand = vb + 1;
      }
      cand = __reduce_max_sync(KAB_FULL_MASK, cand);
      if (lane == 0 && cand >= 0)
        for (uint32_t rr = 0; rr < NC; ++rr) kab_red_max_cluster_s32(kab_mapa(kab_smem_u32(s_vmax), rr), cand);
    }
    echunks = ec0 + (uint32_t)n_chunks;
    __syncthreads();
    kab_cluster_sync();

    const int v = *s_vmax;
    const int status = *s_bad ? 3 : (v < 0 ? 1 : 0);
    if (!is_prod && !is_prep && owned && status == 0 && p.final_score) {
      if (vb + 0 == v) p.final_score[lat.index] = s0;
      if (vb + 1 == v) p.final_score[lat.index] = s1;
    }
    if (rank == 0 && tid == 0) {
      p.status[lat.index] = status;
      if (status != 0 && p.final_score) p.final_score[lat.index] = __int_as_float(0x7fc00000);
      if (status == 0) p.end_state[lat.index] = v;  // traceback by kab_bt_maps_kernel / kab_bt_stitch_kernel
    }
    __syncthreads();  // everybody has read s_vmax / s_bad before they are reset for the next lattice
  }
}
