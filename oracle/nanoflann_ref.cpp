// nanoflann_ref.cpp -- C wrapper around the REFERENCE's own nanoflann +
// KDTreeVectorOfVectorsAdaptor, compiled from the sources where they lie under
// /root/reference/registration (never copied into this repo) into
// oracle/_ref/libnanoflann_ref.so by oracle/Makefile.
//
// TEST INFRASTRUCTURE ONLY: used (a) to pin oracle/knn_oracle.c against the real
// reference, (b) to mint tests/golden/knn_*.npz, (c) as the CPU baseline of
// bench.py (cpu_baseline.kind = "reference").  Mirrors exactly how the reference
// uses the tree: InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>
// (registration/loop_detector.h:31-32), built with leaf_max_size 10
// (loop_detector.cpp:36) and queried one descriptor at a time (:45).
#include <cstddef>
#include <cstdint>
#include <memory>
#include <thread>
#include <vector>

#include "KDTreeVectorOfVectorsAdaptor.h"  // from -I/root/reference/registration

using KeyMat = std::vector<std::vector<float>>;
using InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>;

struct RefTree {
  KeyMat mat;
  std::unique_ptr<InvKeyTree> tree;
  size_t dim;
};

extern "C" {

void* gloc_ref_knn_build(const float* db, size_t n, size_t dim, int leaf_max_size) {
  RefTree* t = new RefTree;
  t->dim = dim;
  t->mat.resize(n);
  for (size_t i = 0; i < n; ++i) t->mat[i].assign(db + i * dim, db + (i + 1) * dim);
  t->tree.reset(new InvKeyTree(dim, t->mat, leaf_max_size));
  return t;
}

void gloc_ref_knn_query(const void* h, const float* q, size_t k, uint64_t* idx,
                        float* d2) {
  const RefTree* t = static_cast<const RefTree*>(h);
  std::vector<size_t> ids(k);
  t->tree->query(q, k, ids.data(), d2);
  for (size_t i = 0; i < k; ++i) idx[i] = ids[i];
}

// queries partitioned over nthreads (the tree is read-only, query() is const)
void gloc_ref_knn_query_batch(const void* h, const float* q, size_t nq, size_t k,
                              uint64_t* idx, float* d2, int nthreads) {
  const RefTree* t = static_cast<const RefTree*>(h);
  if (nthreads < 1) nthreads = 1;
  auto work = [&](size_t q0, size_t q1) {
    for (size_t i = q0; i < q1; ++i)
      gloc_ref_knn_query(h, q + i * t->dim, k, idx + i * k, d2 + i * k);
  };
  if (nthreads == 1) {
    work(0, nq);
    return;
  }
  std::vector<std::thread> th;
  for (int i = 0; i < nthreads; ++i)
    th.emplace_back(work, nq * i / nthreads, nq * (i + 1) / nthreads);
  for (auto& x : th) x.join();
}

void gloc_ref_knn_free(void* h) { delete static_cast<RefTree*>(h); }

}  // extern "C"
