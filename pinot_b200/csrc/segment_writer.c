/*
 * segment_writer.c — CPU-side writers for the Pinot on-disk index layouts the
 * B200 executor consumes.  This is *tooling* (the synthetic-segment generator
 * and the test fixtures use it); it is not on the query path.
 *
 * Layouts restated from the reference (no code copied):
 *   - fixed-bit forward index: big-endian, MSB-first bitstream, value i at bit
 *     offset i*w  (pinot-segment-local/.../io/writer/impl/FixedBitSVForwardIndexWriter.java:40-46,
 *     .../io/util/PinotDataBitSet.java:61-102)
 *   - RoaringBitmap portable serialisation (third-party spec, RoaringBitmap 1.3.0;
 *     used by .../creator/impl/inv/BitmapInvertedIndexWriter.java:35-50,90-97)
 *
 * Build: gcc -O3 -shared -fPIC -o libpinot_b200_segwriter.so segment_writer.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Pack n values of w bits (1..32) into a big-endian MSB-first bitstream.
 * out must hold ceil(n*w/8) bytes and be zero-initialised. */
void pbw_pack_bits_be(const uint32_t* values, int64_t n, int w, uint8_t* out) {
  uint64_t acc = 0;   /* bits waiting to be flushed, right-aligned */
  int nacc = 0;
  int64_t o = 0;
  const uint64_t mask = (w == 32) ? 0xffffffffull : ((1ull << w) - 1ull);
  for (int64_t i = 0; i < n; i++) {
    acc = (acc << w) | ((uint64_t)values[i] & mask);
    nacc += w;
    while (nacc >= 8) {
      out[o++] = (uint8_t)(acc >> (nacc - 8));
      nacc -= 8;
    }
    acc &= (1ull << nacc) - 1ull;
  }
  if (nacc > 0) out[o++] = (uint8_t)(acc << (8 - nacc));
}

/* Inverse of the above (used only by the writer's self-check). */
void pbw_unpack_bits_be(const uint8_t* in, int64_t n, int w, uint32_t* values) {
  for (int64_t i = 0; i < n; i++) {
    uint64_t bit = (uint64_t)i * (uint64_t)w;
    uint64_t v = 0;
    for (int b = 0; b < w; b++) {
      uint64_t p = bit + (uint64_t)b;
      v = (v << 1) | ((in[p >> 3] >> (7 - (p & 7))) & 1u);
    }
    values[i] = (uint32_t)v;
  }
}

/* ---- RoaringBitmap portable format ------------------------------------- */

#define RB_COOKIE_NO_RUN 12346u
#define RB_COOKIE_RUN 12347u
#define RB_NO_OFFSET_THRESHOLD 4

static void put_u16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put_u32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

/* Serialise a strictly ascending docId list.  Returns the number of bytes the
 * blob needs; writes it when out != NULL (out_cap is checked).  run_optimize
 * mirrors RoaringBitmap.runOptimize(): a container becomes a run container when
 * that is strictly smaller than its array/bitmap form. */
int64_t pbw_roaring_serialize(const uint32_t* docs, int64_t n, int run_optimize,
                              uint8_t* out, int64_t out_cap) {
  /* pass 1: container boundaries */
  int64_t ncont = 0;
  for (int64_t i = 0; i < n;) {
    uint32_t hi = docs[i] >> 16;
    int64_t j = i;
    while (j < n && (docs[j] >> 16) == hi) j++;
    ncont++;
    i = j;
  }
  if (ncont > 65536) return -1;
  int64_t* starts = (int64_t*)malloc(sizeof(int64_t) * (size_t)(ncont + 1));
  uint8_t* kinds = (uint8_t*)malloc((size_t)ncont + 1); /* 0 array, 1 bitmap, 2 run */
  int32_t* nruns = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ncont + 1));
  int64_t c = 0;
  int any_run = 0;
  for (int64_t i = 0; i < n;) {
    uint32_t hi = docs[i] >> 16;
    int64_t j = i;
    int32_t runs = 0;
    while (j < n && (docs[j] >> 16) == hi) {
      if (j == i || docs[j] != docs[j - 1] + 1) runs++;
      j++;
    }
    int64_t card = j - i;
    int64_t sz_plain = card <= 4096 ? 2 * card : 8192;
    int64_t sz_run = 2 + 4 * (int64_t)runs;
    starts[c] = i;
    nruns[c] = runs;
    if (run_optimize && sz_run < sz_plain) { kinds[c] = 2; any_run = 1; }
    else kinds[c] = card <= 4096 ? 0 : 1;
    c++;
    i = j;
  }
  starts[ncont] = n;

  int64_t header;
  int has_offsets;
  if (any_run) {
    header = 4 + (ncont + 7) / 8 + 4 * ncont;
    has_offsets = ncont >= RB_NO_OFFSET_THRESHOLD;
  } else {
    header = 8 + 4 * ncont;
    has_offsets = 1;
  }
  if (has_offsets) header += 4 * ncont;
  int64_t total = header;
  for (c = 0; c < ncont; c++) {
    int64_t card = starts[c + 1] - starts[c];
    total += kinds[c] == 0 ? 2 * card : (kinds[c] == 1 ? 8192 : 2 + 4 * (int64_t)nruns[c]);
  }
  if (out == NULL) { free(starts); free(kinds); free(nruns); return total; }
  if (out_cap < total) { free(starts); free(kinds); free(nruns); return -2; }
  memset(out, 0, (size_t)total);

  int64_t p = 0;
  if (any_run) {
    put_u32(out + p, RB_COOKIE_RUN | ((uint32_t)(ncont - 1) << 16)); p += 4;
    for (c = 0; c < ncont; c++) if (kinds[c] == 2) out[p + c / 8] |= (uint8_t)(1u << (c % 8));
    p += (ncont + 7) / 8;
  } else {
    put_u32(out + p, RB_COOKIE_NO_RUN); p += 4;
    put_u32(out + p, (uint32_t)ncont); p += 4;
  }
  for (c = 0; c < ncont; c++) {
    put_u16(out + p, docs[starts[c]] >> 16); p += 2;
    put_u16(out + p, (uint32_t)(starts[c + 1] - starts[c] - 1)); p += 2;
  }
  int64_t off_pos = p;
  if (has_offsets) p += 4 * ncont;
  for (c = 0; c < ncont; c++) {
    if (has_offsets) put_u32(out + off_pos + 4 * c, (uint32_t)p);
    int64_t s = starts[c], e = starts[c + 1];
    if (kinds[c] == 0) {
      for (int64_t k = s; k < e; k++) { put_u16(out + p, docs[k] & 0xffffu); p += 2; }
    } else if (kinds[c] == 1) {
      for (int64_t k = s; k < e; k++) {
        uint32_t lo = docs[k] & 0xffffu;
        out[p + (lo >> 3)] |= (uint8_t)(1u << (lo & 7)); /* LE u64 words == LE bytes */
      }
      p += 8192;
    } else {
      put_u16(out + p, (uint32_t)nruns[c]); p += 2;
      int64_t k = s;
      while (k < e) {
        int64_t r = k;
        while (r + 1 < e && docs[r + 1] == docs[r] + 1) r++;
        put_u16(out + p, docs[k] & 0xffffu); p += 2;
        put_u16(out + p, (uint32_t)(r - k)); p += 2;
        k = r + 1;
      }
    }
  }
  free(starts); free(kinds); free(nruns);
  return total;
}

/* Build a complete .bitmap.inv buffer for a single-value dictionary column:
 * (card+1) big-endian uint32 offsets followed by one Roaring blob per dictId.
 * Two-call protocol: out == NULL returns the required size. */
int64_t pbw_build_inverted_index(const uint32_t* dict_ids, int64_t num_docs, int32_t card,
                                 int run_optimize, uint8_t* out, int64_t out_cap) {
  /* counting sort of docIds by dictId (stable => ascending docIds per dictId) */
  int64_t* cnt = (int64_t*)calloc((size_t)card + 1, sizeof(int64_t));
  for (int64_t i = 0; i < num_docs; i++) cnt[dict_ids[i] + 1]++;
  for (int32_t d = 0; d < card; d++) cnt[d + 1] += cnt[d];
  uint32_t* sorted = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(num_docs > 0 ? num_docs : 1));
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(card + 1));
  memcpy(cur, cnt, sizeof(int64_t) * (size_t)(card + 1));
  for (int64_t i = 0; i < num_docs; i++) sorted[cur[dict_ids[i]]++] = (uint32_t)i;

  int64_t pos = 4 * ((int64_t)card + 1);
  int64_t rc = 0;
  for (int32_t d = 0; d < card; d++) {
    int64_t n = cnt[d + 1] - cnt[d];
    if (out != NULL) {
      uint32_t o = (uint32_t)pos;
      out[4 * d + 0] = (uint8_t)(o >> 24); out[4 * d + 1] = (uint8_t)(o >> 16);
      out[4 * d + 2] = (uint8_t)(o >> 8);  out[4 * d + 3] = (uint8_t)o;
    }
    int64_t sz = pbw_roaring_serialize(sorted + cnt[d], n, run_optimize,
                                       out ? out + pos : NULL, out ? out_cap - pos : 0);
    if (sz < 0) { rc = sz; break; }
    pos += sz;
  }
  if (rc == 0 && out != NULL) {
    uint32_t o = (uint32_t)pos;
    out[4 * card + 0] = (uint8_t)(o >> 24); out[4 * card + 1] = (uint8_t)(o >> 16);
    out[4 * card + 2] = (uint8_t)(o >> 8);  out[4 * card + 3] = (uint8_t)o;
  }
  free(cnt); free(sorted); free(cur);
  return rc < 0 ? rc : pos;
}
