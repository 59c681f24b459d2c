// gloc_inv_key_tree.hpp -- C++ host mirror of the reference's kNN adaptor over the C ABI.
//
// Same class name, template parameters, constructor and query() signature as
//   KDTreeVectorOfVectorsAdaptor   (/root/reference/registration/KDTreeVectorOfVectorsAdaptor.h:52-132)
// so that `using InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>;`
// (registration/loop_detector.h:31-32) keeps compiling unchanged; the nanoflann KD-tree is
// replaced by the exhaustive GPU search in libgloc3d.so.  Header-only; link with -lgloc3d.
// There is no CPU fallback: construction throws when no B200 is usable.
#ifndef GLOC_INV_KEY_TREE_HPP_
#define GLOC_INV_KEY_TREE_HPP_

#include <cassert>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gloc3d.h"

namespace nanoflann {
struct metric_L2;  // tag only: the metric is always squared L2 in evalMetric's operation order
}

template <class VectorOfVectorsType, typename num_t = double, int DIM = -1,
          class Distance = nanoflann::metric_L2, typename IndexType = size_t>
struct KDTreeVectorOfVectorsAdaptor {
  static_assert(sizeof(num_t) == sizeof(float), "the GPU path stores float32 descriptors");
  static_assert(sizeof(IndexType) == sizeof(uint64_t), "indices cross the C ABI as uint64_t");

  /// Constructor: takes a const ref to the vector of vectors object with the data points
  /// (KDTreeVectorOfVectorsAdaptor.h:70-84).  leaf_max_size is accepted and ignored.
  KDTreeVectorOfVectorsAdaptor(const size_t /* dimensionality */, const VectorOfVectorsType& mat,
                               const int /* leaf_max_size */ = 10, const int device = 0)
      : m_data(mat) {
    assert(mat.size() != 0 && mat[0].size() != 0);
    const size_t dims = mat[0].size();
    if (DIM > 0 && static_cast<int>(dims) != DIM)
      throw std::runtime_error("Data set dimensionality does not match the 'DIM' template argument");
    check(gloc_knn_create(&index, dims, device));
    std::vector<float> flat(mat.size() * dims);  // KeyMat rows are separate heap blocks
    for (size_t i = 0; i < mat.size(); ++i) {
      if (mat[i].size() != dims) throw std::runtime_error("ragged data set");
      for (size_t d = 0; d < dims; ++d) flat[i * dims + d] = static_cast<float>(mat[i][d]);
    }
    check(gloc_knn_set_db(index, flat.data(), mat.size()));
  }
  ~KDTreeVectorOfVectorsAdaptor() { gloc_knn_destroy(index); }
  KDTreeVectorOfVectorsAdaptor(const KDTreeVectorOfVectorsAdaptor&) = delete;
  KDTreeVectorOfVectorsAdaptor& operator=(const KDTreeVectorOfVectorsAdaptor&) = delete;

  /// The handle that replaces nanoflann's index_t* (KDTreeVectorOfVectorsAdaptor.h:66).
  gloc_knn_index* index = nullptr;
  const VectorOfVectorsType& m_data;

  /// Query for the num_closest closest points to a given point (entered as
  /// query_point[0:dim-1]) -- KDTreeVectorOfVectorsAdaptor.h:95-102.  Outputs are
  /// caller-allocated, ascending squared distances, exactly as nanoflann fills them.
  inline void query(const num_t* query_point, const size_t num_closest, IndexType* out_indices,
                    num_t* out_distances_sq) const {
    query_batch(query_point, 1, num_closest, out_indices, out_distances_sq);
  }

  /// Batch extension (new surface: the reference issues one query per call).
  inline void query_batch(const num_t* query_points, const size_t nq, const size_t num_closest,
                          IndexType* out_indices, num_t* out_distances_sq) const {
    check(gloc_knn_query(index, reinterpret_cast<const float*>(query_points), nq, num_closest,
                         reinterpret_cast<uint64_t*>(out_indices),
                         reinterpret_cast<float*>(out_distances_sq)));
  }

  /// SLAM mode (loop_detector.cpp:66-72): search all but the most recent keyframes without
  /// rebuilding anything -- append new rows, move the limit.
  inline void append(const std::vector<num_t>& row) {
    check(gloc_knn_append(index, reinterpret_cast<const float*>(row.data()), 1));
  }
  inline void set_search_limit(size_t n_search) { check(gloc_knn_set_search_limit(index, n_search)); }

  inline size_t kdtree_get_point_count() const { return gloc_knn_size(index); }

 private:
  static void check(int rc) {
    if (rc != GLOC_OK)  // nanoflann throws std::runtime_error too (nanoflann.hpp:1454-1457)
      throw std::runtime_error(std::string("libgloc3d: ") + gloc_last_error());
  }
};

using KeyMat = std::vector<std::vector<float>>;
using InvKeyTree = KDTreeVectorOfVectorsAdaptor<KeyMat, float>;  // loop_detector.h:31-32

#endif  // GLOC_INV_KEY_TREE_HPP_
