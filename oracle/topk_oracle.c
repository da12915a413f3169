/*
 * CPU ORACLE (test infrastructure, not product): exact brute-force top-k with the semantics of
 * ItemProcessor.search (xfmr_rec/data/lightning.py:247-258):
 *   score(q, r) = <q, item_r>            (cosine for unit-norm embeddings; "score = 1 - _distance", :257)
 *   excluded ids are removed BEFORE ranking ("prefilter=True", :252)
 *   the k best by (score descending, item id ascending)   (tie rule from the north star; LanceDB's own
 *   tie order is unspecified).
 * The arithmetic of the reference lives in lancedb 0.25.0 / pylance 0.36.0 (uv.lock:1298, :2429), an
 * APPROXIMATE IVF_HNSW_PQ index that is neither in /root/reference nor installable here, and none of
 * the reference's tests pin its results: parity of this restatement is UNPINNED by the reference.
 *
 * Summation order is part of the definition so that a GPU re-scoring pass can match bit for bit:
 * products and the running sum are in double, sequential in the embedding index, rounded to float once.
 *
 * Also holds the XXH32 restatement used by the hashed-embedding oracle (xxhash 0.8 algorithm, XXH32 for
 * an 8-byte input; README.md:32-36 cites the technique, the reference has no code for it).
 *
 * Build: make -C oracle   (gcc -O2 -fPIC -shared -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float score;
  int64_t id;
} entry_t;

static int better(const entry_t* a, const entry_t* b) {
  if (a->score != b->score) return a->score > b->score;
  return a->id < b->id;
}

/* queries [Q, d], items [N, d] float32 row-major; item_ids [N] or NULL (id = id_base + row);
 * excl [Q, E] int64 (entries equal to INT64_MIN are padding) or NULL;
 * out_scores [Q, k], out_ids [Q, k] (unused slots: -inf / -1). */
void xbo_topk(const float* queries, const float* items, const int64_t* item_ids, int64_t id_base,
              const int64_t* excl, int E, int Q, int N, int d, int k, float* out_scores, int64_t* out_ids) {
  entry_t* heap = (entry_t*)malloc(sizeof(entry_t) * (size_t)(k + 1));
  for (int q = 0; q < Q; ++q) {
    int n = 0; /* sorted insertion list, best first */
    const float* qv = queries + (size_t)q * d;
    for (int r = 0; r < N; ++r) {
      const int64_t id = item_ids ? item_ids[r] : id_base + r;
      int skip = 0;
      for (int e = 0; e < E && excl; ++e) {
        const int64_t x = excl[(size_t)q * E + e];
        if (x != INT64_MIN && x == id) { skip = 1; break; }
      }
      if (skip) continue;
      const float* iv = items + (size_t)r * d;
      double acc = 0.0;
      for (int j = 0; j < d; ++j) acc += (double)qv[j] * (double)iv[j];
      entry_t cur = {(float)acc, id};
      if (n == k && !better(&cur, &heap[n - 1])) continue;
      int pos = n < k ? n : k - 1;
      while (pos > 0 && better(&cur, &heap[pos - 1])) {
        heap[pos] = heap[pos - 1];
        --pos;
      }
      heap[pos] = cur;
      if (n < k) ++n;
    }
    for (int j = 0; j < k; ++j) {
      out_scores[(size_t)q * k + j] = j < n ? heap[j].score : -INFINITY;
      out_ids[(size_t)q * k + j] = j < n ? heap[j].id : -1;
    }
  }
  free(heap);
}

/* XXH32 of the 8 little-endian bytes of `id` with `seed` (xxhash.h: len < 16 path). */
uint32_t xbo_xxh32_i64(int64_t id, uint32_t seed) {
  const uint32_t P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  uint8_t bytes[8];
  uint64_t x = (uint64_t)id;
  for (int i = 0; i < 8; ++i) bytes[i] = (uint8_t)(x >> (8 * i));
  uint32_t h = seed + P5 + 8u;
  for (int w = 0; w < 2; ++w) {
    uint32_t lane = (uint32_t)bytes[4 * w] | ((uint32_t)bytes[4 * w + 1] << 8) | ((uint32_t)bytes[4 * w + 2] << 16) |
                    ((uint32_t)bytes[4 * w + 3] << 24);
    h += lane * P3;
    h = ((h << 17) | (h >> 15)) * P4;
  }
  h ^= h >> 15;
  h *= P2;
  h ^= h >> 13;
  h *= P3;
  h ^= h >> 16;
  return h;
}

void xbo_hash_indices(const int64_t* ids, int64_t n, int num_hashes, uint32_t seed0, int log2_rows, int32_t* out) {
  const uint32_t mask = log2_rows >= 32 ? 0xffffffffu : ((1u << log2_rows) - 1u);
  for (int64_t i = 0; i < n; ++i)
    for (int h = 0; h < num_hashes; ++h) out[i * num_hashes + h] = (int32_t)(xbo_xxh32_i64(ids[i], seed0 + (uint32_t)h) & mask);
}
