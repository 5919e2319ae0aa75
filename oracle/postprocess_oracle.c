/* ORACLE (test infrastructure, not product).  Plain-C restatement of the reference's
 * post-processing for the hot path:
 *
 *   top_k_predictions + sigmoid + ScoreEntry ordering   src/postprocess.rs:8-93
 *   filter_predictions_impl                              src/rangefilter.rs:333-386
 *   calculate_week                                       src/rangefilter.rs:77-81
 *   random_logits / mock_embeddings LCG                  src/testutil.rs:110-147
 *
 * std::collections::BinaryHeap is re-stated from the Rust standard library's algorithm
 * (push = sift_up; pop = swap-remove root + sift_down_to_bottom + sift_up; into_iter walks the
 * backing Vec in order), so the survivor order *before* the final stable sort is reproduced
 * too.  PINNED against the reference's own known-answer tests (postprocess.rs:102-331,
 * rangefilter.rs:589-916) in tests/test_oracle_postprocess.py.  The order among entries with
 * equal score is whatever this emulation yields; the reference itself leaves it unspecified
 * (postprocess.rs:207).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t index; float score; } entry_t;
typedef struct { uint32_t index; float confidence; } pred_t;

/* f32::total_cmp: compares the bit patterns as sign-magnitude integers */
static int total_cmp(float a, float b) {
    int32_t x, y;
    memcpy(&x, &a, 4);
    memcpy(&y, &b, 4);
    x ^= (int32_t)(((uint32_t)(x >> 31)) >> 1);
    y ^= (int32_t)(((uint32_t)(y >> 31)) >> 1);
    return (x > y) - (x < y);
}

/* ScoreEntry::cmp = other.score.total_cmp(&self.score)   (postprocess.rs:29-35) */
static int entry_cmp(const entry_t* self, const entry_t* other) { return total_cmp(other->score, self->score); }
static int entry_le(const entry_t* a, const entry_t* b) { return entry_cmp(a, b) <= 0; }

static void sift_up(entry_t* d, size_t start, size_t pos) {
    entry_t hole = d[pos];
    while (pos > start) {
        size_t parent = (pos - 1) / 2;
        if (entry_le(&hole, &d[parent])) break;
        d[pos] = d[parent];
        pos = parent;
    }
    d[pos] = hole;
}

static void sift_down_to_bottom(entry_t* d, size_t end) {
    size_t pos = 0, start = 0;
    entry_t hole = d[0];
    size_t child = 1;
    while (end >= 2 && child <= end - 2) {
        if (entry_le(&d[child], &d[child + 1])) child += 1;
        d[pos] = d[child];
        pos = child;
        child = 2 * pos + 1;
    }
    if (child == end - 1) {
        d[pos] = d[child];
        pos = child;
    }
    d[pos] = hole;
    sift_up(d, start, pos);
}

float oracle_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }   /* postprocess.rs:91-93 */

/* Returns the number of predictions written (<= min(top_k, n)).  has_min: Option<f32>. */
uint64_t oracle_top_k(const float* logits, uint64_t n, uint64_t top_k, int has_min, float min_conf,
                      pred_t* out) {
    if (n == 0 || top_k == 0) return 0;                       /* postprocess.rs:46-48 */
    uint64_t k = top_k < n ? top_k : n;                       /* :50 */
    entry_t* heap = (entry_t*)malloc((k + 1) * sizeof(entry_t));
    size_t len = 0;
    for (uint64_t i = 0; i < n; ++i) {                        /* :55-60 */
        heap[len].index = i;
        heap[len].score = logits[i];
        sift_up(heap, 0, len);
        len++;
        if (len > k) {                                        /* pop(): remove the smallest */
            entry_t last = heap[--len];
            if (len > 0) {
                heap[0] = last;                               /* swap(&mut item, &mut data[0]) */
                sift_down_to_bottom(heap, len);
            }
        }
    }
    /* into_iter (Vec order) -> sigmoid -> filter -> stable sort by confidence desc (:63-84) */
    uint64_t m = 0;
    for (size_t i = 0; i < len; ++i) {
        float c = oracle_sigmoid(heap[i].score);
        if (has_min && !(c >= min_conf)) continue;
        out[m].index = (uint32_t)heap[i].index;
        out[m].confidence = c;
        m++;
    }
    free(heap);
    /* stable insertion sort; comparator b.partial_cmp(a).unwrap_or(Equal) */
    for (uint64_t i = 1; i < m; ++i) {
        pred_t x = out[i];
        uint64_t j = i;
        while (j > 0) {
            float a = out[j - 1].confidence, b = x.confidence;
            int less = (b > a);                /* x sorts before out[j-1] iff b.partial_cmp(a) == Greater */
            if (!less) break;
            out[j] = out[j - 1];
            j--;
        }
        out[j] = x;
    }
    return m;
}

void oracle_top_k_batch(const float* logits, uint64_t rows, uint64_t n, uint64_t top_k, int has_min,
                        float min_conf, pred_t* out, uint32_t* counts) {
    uint64_t k = top_k < n ? top_k : n;
    for (uint64_t r = 0; r < rows; ++r)
        counts[r] = (uint32_t)oracle_top_k(logits + r * n, n, top_k, has_min, min_conf, out + r * k);
}

/* filter_predictions_impl with the location map given densely per prediction:
 * state[j]: 0 = species not in the map, 1 = in the map (score[j]).   rangefilter.rs:346-378.
 * The rerank sort (sort_unstable_by total_cmp desc) is done stably here; ties are unpinned. */
uint64_t oracle_filter(const pred_t* in, uint64_t n, const uint8_t* in_map, const float* score,
                       float threshold, int rerank, pred_t* out) {
    uint64_t m = 0;
    for (uint64_t j = 0; j < n; ++j) {
        if (in_map[j]) {
            if (score[j] >= threshold) {
                out[m] = in[j];
                if (rerank) out[m].confidence = in[j].confidence * score[j];
                m++;
            }
        } else {
            out[m++] = in[j];
        }
    }
    if (rerank) {
        for (uint64_t i = 1; i < m; ++i) {
            pred_t x = out[i];
            uint64_t j = i;
            while (j > 0 && total_cmp(x.confidence, out[j - 1].confidence) > 0) {
                out[j] = out[j - 1];
                j--;
            }
            out[j] = x;
        }
    }
    return m;
}

float oracle_calculate_week(uint32_t month, uint32_t day) {
    return (float)((month - 1) * 4 + (day - 1) / 7 + 1);
}

void oracle_random_logits(uint64_t count, uint64_t seed, float* out) {     /* testutil.rs:110-121 */
    uint64_t state = seed;
    for (uint64_t i = 0; i < count; ++i) {
        state = state * 1103515245ULL + 12345ULL;
        float bits = (float)((state >> 16) & 0xFFFF);
        out[i] = fmaf(bits, 10.0f / 65535.0f, -5.0f);
    }
}

void oracle_mock_embeddings(uint64_t dim, uint64_t seed, float* out) {     /* testutil.rs:137-147 */
    uint64_t state = seed;
    for (uint64_t i = 0; i < dim; ++i) {
        state = state * 1103515245ULL + 12345ULL;
        float bits = (float)((state >> 16) & 0xFFFF);
        out[i] = bits / 65535.0f;
    }
}
