/*
 * oracle/ctc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, per-cell CPU restatement of the reference's banded, max_move-transition
 * CTC best-path alignment (kaiidams/Kokoro-Align, kokoro_align/align.py:43-109 and the
 * traceback in align.py:21-40).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.  The product
 * path (kokoro-align_b200/csrc) never links, loads or calls it.
 *
 * Parity is PINNED: tests/test_oracle_golden.py checks this restatement bit-for-bit
 * against tests/golden/, whose vectors were produced by running the reference's own
 * ctc_best_path (imported from /root/reference by tests/golden/make_golden.py).
 *
 * The recurrence, with S = 2L+1, W = beam_size, M = max_move:
 *
 *   active_{-1} = {0}, score_{-1}[0] = 0.0f                          align.py:57-58
 *   for i in 0..T-1:
 *     lo = max(0, (S*i)/T - W/2);  hi = min(lo + W, S)               align.py:64-65
 *     for v in lo..hi-1:
 *       e = lp[i, ext[v]]                                            align.py:77
 *       for j = 0..M-1 ascending:                                    align.py:70
 *         u = v - j; skip if u < 0 or u inactive at frame i-1        align.py:71-76
 *         skip if j > 0, j even and ext[v] == 0 (value test)         align.py:80-81
 *         val = fl32(score_{i-1}[u] + e)                             align.py:77
 *         strict '>' keeps the smallest j on ties                    align.py:83-85
 *   end state = highest active state of frame T-1                    align.py:99-101
 *   walk back v -= move[i][v]                                        align.py:21-40
 *   best_labels = ext[path]; best_scores[i] = lp[i, best_labels[i]]  align.py:105-107
 *
 * Status codes (shared with include/kokoro_align_b200.h):
 *   0 ok, 1 dead band (reference: ValueError from np.argmax([]), align.py:101),
 *   2 label outside [-V, V) (reference: IndexError, align.py:77),
 *   3 non-finite log-prob anywhere in the [T,V] array (rejected instead of emulating
 *     the reference's all--inf-column quirk).
 * Negative labels in [-V, -1] index from the end of the row exactly as numpy does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_DEAD_BAND 1
#define ORACLE_BAD_LABEL 2
#define ORACLE_NONFINITE 3
#define ORACLE_NOMEM (-1)
#define ORACLE_BAD_ARG (-2)

static int64_t band_lo(int64_t S, int64_t i, int64_t T, int64_t W) {
  /* Python floor division on non-negative operands == C division. align.py:64 */
  int64_t lo = (S * i) / T - W / 2;
  return lo < 0 ? 0 : lo;
}

/* Number of cells the reference evaluates: sum_i (hi_i - lo_i).  SURVEY.md 8(d). */
int64_t oracle_cells_eval(int64_t T, int64_t L, int32_t beam_size) {
  int64_t S = 2 * L + 1, W = beam_size, cells = 0;
  for (int64_t i = 0; i < T; ++i) {
    int64_t lo = band_lo(S, i, T, W);
    int64_t hi = lo + W < S ? lo + W : S;
    if (hi > lo) cells += hi - lo;
  }
  return cells;
}

int oracle_ctc_best_path(const float *lp, int64_t T, int32_t V, const int32_t *labels,
                         int64_t L, int32_t beam_size, int32_t max_move,
                         int32_t *best_path, int32_t *best_labels, float *best_scores,
                         float *final_score) {
  if (T <= 0 || V <= 0 || L < 0 || max_move < 1 || max_move > 255) return ORACLE_BAD_ARG;
  const int64_t S = 2 * L + 1, W = beam_size, M = max_move;

  for (int64_t l = 0; l < L; ++l)
    if (labels[l] < -V || labels[l] >= V) return ORACLE_BAD_LABEL;
  for (int64_t n = 0; n < T * (int64_t)V; ++n)
    if (!isfinite(lp[n])) return ORACLE_NONFINITE;

  /* ext[v]: raw label value (0 on even v); col[v]: numpy-style column index. align.py:46-48 */
  int32_t *ext = (int32_t *)calloc((size_t)S, sizeof(int32_t));
  int32_t *col = (int32_t *)calloc((size_t)S, sizeof(int32_t));
  /* scores padded with M-1 permanently inactive states below state 0 */
  const int64_t pad = M - 1;
  float *prev = (float *)malloc((size_t)(S + pad) * sizeof(float));
  float *cur = (float *)malloc((size_t)(S + pad) * sizeof(float));
  int64_t Wc = W < S ? W : S;
  if (Wc < 1) Wc = 1;
  uint8_t *move = (uint8_t *)malloc((size_t)T * (size_t)Wc);
  int64_t *los = (int64_t *)malloc((size_t)T * sizeof(int64_t));
  if (!ext || !col || !prev || !cur || !move || !los) {
    free(ext); free(col); free(prev); free(cur); free(move); free(los);
    return ORACLE_NOMEM;
  }
  for (int64_t l = 0; l < L; ++l) {
    ext[2 * l + 1] = labels[l];
    col[2 * l + 1] = labels[l] < 0 ? labels[l] + V : labels[l];
  }
  for (int64_t v = 0; v < S + pad; ++v) prev[v] = cur[v] = -INFINITY;
  prev[pad + 0] = 0.0f; /* virtual start, align.py:57-58 */

  int64_t plo = 0, phi = 1;     /* range of prev that may be active */
  int64_t clo = 0, chi = 0;     /* range of cur last written (two frames ago) */
  for (int64_t i = 0; i < T; ++i) {
    int64_t lo = band_lo(S, i, T, W);
    int64_t hi = lo + W < S ? lo + W : S;
    if (hi < lo) hi = lo;
    los[i] = lo;
    for (int64_t v = clo; v < chi; ++v) cur[pad + v] = -INFINITY;
    const float *row = lp + i * (int64_t)V;
    uint8_t *mrow = move + i * Wc;
    for (int64_t v = lo; v < hi; ++v) {
      const float e = row[col[v]];
      float best = -INFINITY;
      int bj = 0;
      for (int64_t j = 0; j < M; ++j) {
        if (j > 0 && (j & 1) == 0 && ext[v] == 0) continue; /* align.py:80-81 */
        /* volatile forces one IEEE binary32 add, no excess precision */
        volatile float val = prev[pad + v - j] + e;
        if (val > best) { best = val; bj = (int)j; }
      }
      cur[pad + v] = best; /* -inf <=> inactive (log-probs are finite) */
      mrow[v - lo] = (uint8_t)bj;
    }
    /* swap; what was prev becomes the buffer to clear next time */
    float *t = prev; prev = cur; cur = t;
    clo = plo; chi = phi;
    plo = lo; phi = hi;
  }

  /* forced end: highest active state of the last frame, align.py:99-101 */
  int64_t v = -1;
  for (int64_t u = phi - 1; u >= plo; --u)
    if (prev[pad + u] > -INFINITY) { v = u; break; }
  int status = ORACLE_OK;
  if (v < 0) {
    status = ORACLE_DEAD_BAND;
  } else {
    if (final_score) *final_score = prev[pad + v];
    for (int64_t i = T - 1; i >= 0; --i) {
      best_path[i] = (int32_t)v;
      best_labels[i] = ext[v];
      best_scores[i] = lp[i * (int64_t)V + col[v]];
      v -= move[i * Wc + (v - los[i])];
    }
  }
  free(ext); free(col); free(prev); free(cur); free(move); free(los);
  return status;
}

/*
 * Batched form with the same flat layout as the product C-ABI (offset arrays), so the
 * tests and the CPU-baseline leg of bench.py can run many lattices on n_threads host
 * threads (pthreads; lattices are handed out through an atomic counter).
 */
#include <pthread.h>
#include <stdatomic.h>

typedef struct {
  const float *lp; const int64_t *t_off; const int32_t *labels; const int64_t *l_off;
  int64_t B; int32_t V, beam_size, max_move;
  int32_t *best_path, *best_labels; float *best_scores, *final_score; int32_t *status;
  atomic_llong next;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *jb = (batch_job *)arg;
  for (;;) {
    int64_t b = (int64_t)atomic_fetch_add(&jb->next, 1);
    if (b >= jb->B) break;
    int64_t t0 = jb->t_off[b], T = jb->t_off[b + 1] - t0;
    int64_t l0 = jb->l_off[b], L = jb->l_off[b + 1] - l0;
    float fs = 0.0f;
    int st = oracle_ctc_best_path(jb->lp + t0 * (int64_t)jb->V, T, jb->V, jb->labels + l0, L,
                                  jb->beam_size, jb->max_move, jb->best_path + t0,
                                  jb->best_labels + t0, jb->best_scores + t0, &fs);
    jb->status[b] = st;
    jb->final_score[b] = st == ORACLE_OK ? fs : NAN;
  }
  return NULL;
}

int oracle_ctc_best_path_batch(const float *lp, const int64_t *t_off, const int32_t *labels,
                               const int64_t *l_off, int64_t B, int32_t V, int32_t beam_size,
                               int32_t max_move, int32_t *best_path, int32_t *best_labels,
                               float *best_scores, float *final_score, int32_t *status,
                               int32_t n_threads) {
  batch_job jb = {lp, t_off, labels, l_off, B, V, beam_size, max_move,
                  best_path, best_labels, best_scores, final_score, status, 0};
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 1024) n_threads = 1024;
  pthread_t *th = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
  if (!th) return ORACLE_NOMEM;
  int started = 0;
  for (int k = 0; k < n_threads - 1; ++k)
    if (pthread_create(&th[started], NULL, batch_worker, &jb) == 0) ++started;
  batch_worker(&jb);
  for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
  free(th);
  return 0;
}
