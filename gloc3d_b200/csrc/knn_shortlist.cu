// knn_shortlist.cu -- placeholder until the tcgen05 shortlist lands: reports "not
// applicable" so GLOC_KNN_AUTO uses the exact scan.
#include "knn_shortlist.cuh"

namespace gloc {
struct ShortlistState {};
bool shortlist_supported(size_t, size_t) { return false; }
bool shortlist_applicable(size_t, size_t, size_t, size_t) { return false; }
void shortlist_invalidate(ShortlistState*, size_t) {}
void shortlist_destroy(ShortlistState*) {}
int shortlist_query(ShortlistState**, const ShortlistArgs&, uint64_t*, uint64_t*, uint64_t*) {
  return fail(GLOC_ERR_RANGE, "tensor shortlist not built");
}
}  // namespace gloc
