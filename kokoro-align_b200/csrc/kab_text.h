// kab_text.h -- host-only text helpers of the C ABI (no CUDA): the transcript scanner and the
// regular expression of merge_repeated.  Included by kab_api.cu at file scope (KAB_* codes come
// from the public header).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

extern "C" {

int kab_encode_transcript(const uint8_t *text, int64_t n_bytes, const int16_t *token_ids, int8_t *labels,
                          int64_t *n_labels) {
  if (n_bytes < 0 || !token_ids || !n_labels || (n_bytes > 0 && (!text || !labels))) return KAB_E_BAD_ARG;
  int64_t n = 0, i = 0;
  while (i < n_bytes) {
    // one line [i, e): up to '\n'; a '\r' right before it is stripped (rstrip('\r\n'), transcript.py:64)
    int64_t e = i;
    while (e < n_bytes && text[e] != '\n') {
      if (text[e] == '\r' && !(e + 1 < n_bytes && text[e + 1] == '\n')) return KAB_E_UNSUPPORTED;  // universal newlines
      ++e;
    }
    int64_t le = e;
    if (le > i && text[le - 1] == '\r') --le;
    int64_t b = i;
    while (b < le && text[b] != '|') ++b;
    if (b >= le) return KAB_E_UNSUPPORTED;  // no second field: the reference raises IndexError (parts[1])
    int64_t f = b + 1, fe = f;
    while (fe < le && text[fe] != '|') ++fe;
    // tokens of the voca field [f, fe): str.split() on runs of spaces (any other whitespace or a
    // non-ASCII byte in the field leaves the plain case), ids by table, unknown tokens dropped
    while (f < fe) {
      while (f < fe && text[f] == ' ') ++f;
      int64_t t = f;
      while (t < fe && text[t] != ' ') {
        if (text[t] < 32 || text[t] > 126) return KAB_E_UNSUPPORTED;
        ++t;
      }
      const int64_t len = t - f;
      if (len == 1 || len == 2) {
        const int16_t id = token_ids[text[f] | (len == 2 ? (unsigned)text[f + 1] << 8 : 0u)];
        if (id >= 0) labels[n++] = (int8_t)id;
      }
      f = t;
    }
    i = e + 1;
  }
  *n_labels = n;
  return KAB_OK;
}

int kab_merge_repeated(const uint8_t *text, int64_t n, uint8_t *out, int64_t *n_out) {
  // re.sub(r'(.+)( \1)+', r'\1', text), encoder.py:28, for text without '\n' (where '.' is any
  // character): leftmost match, group 1 greedy (longest first), then as many " \1" as fit; a match
  // is replaced by group 1 and the scan continues behind it.
  if (n < 0 || !n_out || (n > 0 && (!text || !out))) return KAB_E_BAD_ARG;
  for (int64_t k = 0; k < n; ++k)
    if (text[k] == '\n' || text[k] >= 0x80) return KAB_E_UNSUPPORTED;  // '.' stops at newlines; bytes != characters
  std::vector<int64_t> spaces;  // positions of ' ', ascending: the only places group 1 can end
  for (int64_t k = 0; k < n; ++k)
    if (text[k] == ' ') spaces.push_back(k);
  int64_t i = 0, o = 0;
  size_t s_lo = 0;  // first space position > i
  while (i < n) {
    while (s_lo < spaces.size() && spaces[s_lo] <= i) ++s_lo;
    // group 1 = text[i, i+m), followed by ' ' at i+m and the same m bytes: longest m first
    const int64_t m_max = (n - i - 1) / 2;
    size_t k = std::upper_bound(spaces.begin() + (std::ptrdiff_t)s_lo, spaces.end(), i + m_max) - spaces.begin();
    int64_t m = 0;
    while (k > s_lo) {
      const int64_t cand = spaces[--k] - i;
      if (text[i] == text[i + cand + 1] && memcmp(text + i, text + i + cand + 1, (size_t)cand) == 0) { m = cand; break; }
    }
    if (m < 1) {
      out[o++] = text[i++];
      continue;
    }
    int64_t e = i + 2 * m + 1;
    while (e + m + 1 <= n && text[e] == ' ' && memcmp(text + i, text + e + 1, (size_t)m) == 0) e += m + 1;
    memcpy(out + o, text + i, (size_t)m);
    o += m;
    i = e;
  }
  *n_out = o;
  return KAB_OK;
}

}  // extern "C"
