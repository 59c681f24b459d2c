/*
 * knn_oracle.c -- stage-1 oracle: exhaustive exact L2 kNN in the reference's
 * float32 operation order.  TEST INFRASTRUCTURE ONLY (see gloc_oracle.h).
 *
 * Compile WITHOUT -march / -ffast-math and with -ffp-contract=off so that no
 * FMA is formed: the reference is built Release for plain x86-64
 * (registration/CMakeLists.txt:5-7), where g++ emits mulss/addss only.
 */
#include "gloc_oracle.h"

#include <float.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* nanoflann.hpp:453-487 */
float gloc_oracle_l2(const float* a, const float* b, size_t size) {
  float result = 0.0f;
  const float* last = a + size;
  const float* lastgroup = last - 3;
  while (a < lastgroup) { /* :463-478 */
    const float diff0 = a[0] - b[0];
    const float diff1 = a[1] - b[1];
    const float diff2 = a[2] - b[2];
    const float diff3 = a[3] - b[3];
    result += diff0 * diff0 + diff1 * diff1 + diff2 * diff2 + diff3 * diff3;
    a += 4;
    b += 4;
  }
  while (a < last) { /* :481-485 */
    const float diff0 = *a++ - *b++;
    result += diff0 * diff0;
  }
  return result;
}

/* Sorted insertion, the (d2, idx)-ordered variant of KNNResultSet::addPoint
 * (nanoflann.hpp:200-233; equal to its NANOFLANN_FIRST_MATCH branch :207-209). */
static void insert_sorted(uint64_t* idx, float* d2, size_t k, size_t* count,
                          float dist, uint64_t index) {
  size_t i;
  for (i = *count; i > 0; --i) {
    if (d2[i - 1] > dist || (d2[i - 1] == dist && idx[i - 1] > index)) {
      if (i < k) {
        d2[i] = d2[i - 1];
        idx[i] = idx[i - 1];
      }
    } else {
      break;
    }
  }
  if (i < k) {
    d2[i] = dist;
    idx[i] = index;
  }
  if (*count < k) (*count)++;
}

static void knn_range(const float* db, size_t n, size_t dim, const float* q,
                      size_t q0, size_t q1, size_t k, uint64_t* out_idx,
                      float* out_d2) {
  for (size_t qi = q0; qi < q1; ++qi) {
    uint64_t* idx = out_idx + qi * k;
    float* d2 = out_d2 + qi * k;
    size_t count = 0;
    for (size_t j = 0; j < k; ++j) {
      idx[j] = UINT64_MAX;
      d2[j] = FLT_MAX;
    }
    const float* qv = q + qi * dim;
    for (size_t r = 0; r < n; ++r) {
      const float d = gloc_oracle_l2(qv, db + r * dim, dim);
      if (count < k || d < d2[k - 1] ||
          (d == d2[k - 1] && (uint64_t)r < idx[k - 1])) {
        insert_sorted(idx, d2, k, &count, d, (uint64_t)r);
      }
    }
  }
}

void gloc_oracle_knn(const float* db, size_t n, size_t dim, const float* q,
                     size_t nq, size_t k, uint64_t* out_idx, float* out_d2) {
  if (k == 0) return;
  knn_range(db, n, dim, q, 0, nq, k, out_idx, out_d2);
}

typedef struct {
  const float* db;
  size_t n, dim;
  const float* q;
  size_t q0, q1, k;
  uint64_t* out_idx;
  float* out_d2;
} knn_job;

static void* knn_worker(void* p) {
  knn_job* j = (knn_job*)p;
  knn_range(j->db, j->n, j->dim, j->q, j->q0, j->q1, j->k, j->out_idx,
            j->out_d2);
  return NULL;
}

void gloc_oracle_knn_mt(const float* db, size_t n, size_t dim, const float* q,
                        size_t nq, size_t k, uint64_t* out_idx, float* out_d2,
                        int nthreads) {
  if (k == 0 || nq == 0) return;
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > nq) nthreads = (int)nq;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  knn_job* jobs = (knn_job*)malloc(sizeof(knn_job) * (size_t)nthreads);
  for (int t = 0; t < nthreads; ++t) {
    jobs[t] = (knn_job){db, n, dim, q, nq * (size_t)t / (size_t)nthreads,
                        nq * (size_t)(t + 1) / (size_t)nthreads, k, out_idx,
                        out_d2};
    pthread_create(&th[t], NULL, knn_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}

void gloc_oracle_topk_merge(const uint64_t* idx, const float* d2, size_t g,
                            size_t nq, size_t k, uint64_t* out_idx,
                            float* out_d2) {
  for (size_t qi = 0; qi < nq; ++qi) {
    uint64_t* oi = out_idx + qi * k;
    float* od = out_d2 + qi * k;
    size_t count = 0;
    for (size_t j = 0; j < k; ++j) {
      oi[j] = UINT64_MAX;
      od[j] = FLT_MAX;
    }
    for (size_t s = 0; s < g; ++s) {
      const uint64_t* si = idx + (s * nq + qi) * k;
      const float* sd = d2 + (s * nq + qi) * k;
      for (size_t j = 0; j < k; ++j) {
        if (si[j] == UINT64_MAX) continue; /* empty slot of a short shard */
        if (count < k || sd[j] < od[k - 1] ||
            (sd[j] == od[k - 1] && si[j] < oi[k - 1])) {
          insert_sorted(oi, od, k, &count, sd[j], si[j]);
        }
      }
    }
  }
}
