/* CPU ORACLE (plain C) -- TEST INFRASTRUCTURE ONLY, never linked into the product library.
 *
 * Scalar restatement of the sequential parts of the reference hot path, used by tests/ and
 * bench.py's cpu_baseline to check the CUDA kernels at sizes where Python loops are too slow:
 *
 *   mtso_crf_viterbi   <- models/CRF.py:172-216  (fp32 max-plus recurrence, first max wins, back-trace)
 *   mtso_crf_forward   <- models/CRF.py:218-240 + :17-21 (log-sum-exp forward algorithm)
 *   mtso_crf_gold      <- models/CRF.py:148-170
 *   mtso_lstm_dir      <- models/NeuralArchitectures.py:113 (torch.nn.LSTM, gate rows i,f,g,o; packed
 *                         variable-length semantics: reverse starts at len-1, padded steps are zero)
 *   mtso_pk / mtso_wd  <- models/lightning_model.py:26-55 via segeval 2.0.11 defaults (UNPINNED: segeval
 *                         is not available offline; algorithm restated from its published definition)
 *
 * Pinned (except pk/wd) by tests/test_oracle_golden.py against tests/golden/*.npz, which hold outputs
 * of the unmodified reference.  Build: `make -C oracle` (done by __graft_entry__.build()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IMPOSSIBLE (-1e4f)

/* emis [B,L,C] fp32, lengths [B] i64, trans [C,C] (trans[i*C+j] = j -> i).
 * out: best_score [B], paths [B,L] i32 (-1 beyond len_b). */
void mtso_crf_viterbi(const float *emis, const int64_t *lengths, const float *trans, int B, int L, int C,
                      float *best_score, int32_t *paths) {
  const int start = C - 2, stop = C - 1;
  int32_t *bp = (int32_t *)malloc((size_t)L * C * sizeof(int32_t));
  float *s = (float *)malloc(C * sizeof(float)), *n = (float *)malloc(C * sizeof(float));
  for (int b = 0; b < B; ++b) {
    int len = (int)lengths[b];
    for (int i = 0; i < C; ++i) s[i] = IMPOSSIBLE;
    s[start] = 0.f;
    for (int t = 0; t < L; ++t) {
      for (int i = 0; i < C; ++i) {
        float best = s[0] + trans[i * C + 0];
        int arg = 0;
        for (int j = 1; j < C; ++j) {
          float v = s[j] + trans[i * C + j];
          if (v > best) { best = v; arg = j; }
        }
        bp[t * C + i] = arg;
        n[i] = best + emis[((size_t)b * L + t) * C + i];
      }
      if (t < len) memcpy(s, n, C * sizeof(float)); /* mask blend x*1 + old*0 is exact */
    }
    int tag = 0;
    float best = s[0] + trans[stop * C + 0];
    for (int i = 1; i < C; ++i) {
      float v = s[i] + trans[stop * C + i];
      if (v > best) { best = v; tag = i; }
    }
    best_score[b] = best;
    for (int t = 0; t < L; ++t) paths[(size_t)b * L + t] = -1;
    for (int t = len - 1; t >= 0; --t) {
      paths[(size_t)b * L + t] = tag;
      tag = bp[t * C + tag];
    }
  }
  free(bp); free(s); free(n);
}

static float lse(const float *x, int n) {
  float m = x[0];
  for (int i = 1; i < n; ++i) if (x[i] > m) m = x[i];
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += expf(x[i] - m);
  return m + logf(acc);
}

void mtso_crf_forward(const float *emis, const int64_t *lengths, const float *trans, int B, int L, int C, float *logz) {
  const int start = C - 2, stop = C - 1;
  float s[16], n[16], tmp[16];
  for (int b = 0; b < B; ++b) {
    int len = (int)lengths[b];
    for (int i = 0; i < C; ++i) s[i] = IMPOSSIBLE;
    s[start] = 0.f;
    for (int t = 0; t < L && t < len; ++t) {
      for (int i = 0; i < C; ++i) {
        for (int j = 0; j < C; ++j) tmp[j] = (s[j] + trans[i * C + j]) + emis[((size_t)b * L + t) * C + i];
        n[i] = lse(tmp, C);
      }
      memcpy(s, n, C * sizeof(float));
    }
    for (int j = 0; j < C; ++j) tmp[j] = s[j] + trans[stop * C + j];
    logz[b] = lse(tmp, C);
  }
}

void mtso_crf_gold(const float *emis, const int64_t *tags, const int64_t *lengths, const float *trans, int B, int L,
                   int C, float *gold) {
  const int start = C - 2, stop = C - 1;
  for (int b = 0; b < B; ++b) {
    int len = (int)lengths[b];
    float acc = 0.f;
    int prev = start;
    for (int t = 0; t < len; ++t) {
      int y = (int)tags[(size_t)b * L + t];
      acc += trans[y * C + prev] + emis[((size_t)b * L + t) * C + y];
      prev = y;
    }
    gold[b] = acc + trans[stop * C + prev];
  }
}

static float sigm(float x) { return 1.f / (1.f + expf(-x)); }

/* One direction of one LSTM layer.  x [B,T,D], out [B,T,H] (zero beyond len_b). */
void mtso_lstm_dir(const float *x, const int64_t *lengths, const float *w_ih, const float *w_hh, const float *b_ih,
                   const float *b_hh, int B, int T, int D, int H, int reverse, float *out) {
  float *h = (float *)malloc(H * sizeof(float)), *c = (float *)malloc(H * sizeof(float));
  float *pre = (float *)malloc(4 * H * sizeof(float));
  memset(out, 0, (size_t)B * T * H * sizeof(float));
  for (int b = 0; b < B; ++b) {
    int len = (int)lengths[b];
    memset(h, 0, H * sizeof(float));
    memset(c, 0, H * sizeof(float));
    for (int s = 0; s < len; ++s) {
      int t = reverse ? len - 1 - s : s;
      const float *xt = x + ((size_t)b * T + t) * D;
      for (int r = 0; r < 4 * H; ++r) {
        float a = b_ih[r] + b_hh[r];
        const float *wi = w_ih + (size_t)r * D;
        for (int k = 0; k < D; ++k) a += wi[k] * xt[k];
        const float *wh = w_hh + (size_t)r * H;
        for (int k = 0; k < H; ++k) a += wh[k] * h[k];
        pre[r] = a;
      }
      float *o = out + ((size_t)b * T + t) * H;
      for (int u = 0; u < H; ++u) {
        float ig = sigm(pre[u]), fg = sigm(pre[H + u]), gg = tanhf(pre[2 * H + u]), og = sigm(pre[3 * H + u]);
        c[u] = fg * c[u] + ig * gg;
        o[u] = og * tanhf(c[u]);
      }
      memcpy(h, o, H * sizeof(float));
    }
  }
  free(h); free(c); free(pre);
}

/* position labels: segment index of every unit, after forcing the last unit to be a boundary */
static int *positions(const uint8_t *bnd, int n) {
  int *p = (int *)malloc(n * sizeof(int));
  int seg = 1;
  for (int i = 0; i < n; ++i) {
    p[i] = seg;
    if (bnd[i] || i == n - 1) ++seg;
  }
  return p;
}

static int window_size(const uint8_t *ref, int n) {
  int nseg = 0;
  for (int i = 0; i < n; ++i) if (ref[i] || i == n - 1) ++nseg;
  /* k = round_half_even(n / nseg / 2) with exact rational arithmetic, min 2 */
  long num = n, den = 2L * nseg;
  long q = num / den, r = num % den;
  if (2 * r > den || (2 * r == den && (q & 1))) ++q;
  return q > 1 ? (int)q : 2;
}

/* returns numerator; *den receives the number of windows (0 windows -> value 0) */
int mtso_pk(const uint8_t *hyp, const uint8_t *ref, int n, int *den) {
  int *r = positions(ref, n), *h = positions(hyp, n);
  int k = window_size(ref, n), diff = 0, meas = 0;
  for (int i = 0; i + k < n; ++i) {
    if ((r[i] == r[i + k]) != (h[i] == h[i + k])) ++diff;
    ++meas;
  }
  free(r); free(h);
  *den = meas;
  return diff;
}

/* returns numerator, or -1 when segeval would raise (no window fits) */
int mtso_wd(const uint8_t *hyp, const uint8_t *ref, int n, int *den) {
  int *r = positions(ref, n), *h = positions(hyp, n);
  int k = window_size(ref, n), diff = 0;
  *den = n - k;
  if (n - k <= 0) { free(r); free(h); return -1; }
  for (int i = 0; i + k < n; ++i) {
    int rb = 0, hb = 0;
    for (int j = i; j < i + k; ++j) { rb += r[j] != r[j + 1]; hb += h[j] != h[j + 1]; }
    diff += rb != hb;
  }
  free(r); free(h);
  return diff;
}
