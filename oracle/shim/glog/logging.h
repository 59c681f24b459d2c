// glog/logging.h -- TEST INFRASTRUCTURE ONLY (oracle/).  CHECK / DCHECK / LOG with glog's
// semantics where the reference's registration/2d code depends on them: a failed CHECK prints
// and aborts, DCHECKs compile to nothing (Release), LOG(INFO) is silent.
#ifndef GLOC_ORACLE_GLOG_SHIM_H_
#define GLOC_ORACLE_GLOG_SHIM_H_
#include <cstdlib>
#include <iostream>
#include <sstream>

namespace gloc_glog_shim {
struct NullStream {
  template <typename T>
  NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  std::ostringstream os;
  FatalStream(const char* file, int line, const char* what) { os << file << ":" << line << " Check failed: " << what << " "; }
  [[noreturn]] ~FatalStream() {
    std::cerr << os.str() << std::endl;
    std::abort();
  }
  template <typename T>
  FatalStream& operator<<(const T& v) {
    os << v;
    return *this;
  }
};
template <typename T>
T* check_notnull(const char* file, int line, const char* what, T* p) {
  if (p == nullptr) FatalStream(file, line, what);
  return p;
}
}  // namespace gloc_glog_shim

#define CHECK(c) while (!(c)) gloc_glog_shim::FatalStream(__FILE__, __LINE__, #c)
#define GLOC_CHECK_OP(a, op, b) while (!((a)op(b))) gloc_glog_shim::FatalStream(__FILE__, __LINE__, #a " " #op " " #b)
#define CHECK_EQ(a, b) GLOC_CHECK_OP(a, ==, b)
#define CHECK_NE(a, b) GLOC_CHECK_OP(a, !=, b)
#define CHECK_LE(a, b) GLOC_CHECK_OP(a, <=, b)
#define CHECK_LT(a, b) GLOC_CHECK_OP(a, <, b)
#define CHECK_GE(a, b) GLOC_CHECK_OP(a, >=, b)
#define CHECK_GT(a, b) GLOC_CHECK_OP(a, >, b)
#define CHECK_NOTNULL(p) gloc_glog_shim::check_notnull(__FILE__, __LINE__, #p " must be non-null", (p))
#define DCHECK(c) while (false) gloc_glog_shim::NullStream()
#define DCHECK_EQ(a, b) while (false) gloc_glog_shim::NullStream()
#define DCHECK_NE(a, b) while (false) gloc_glog_shim::NullStream()
#define DCHECK_LE(a, b) while (false) gloc_glog_shim::NullStream()
#define DCHECK_LT(a, b) while (false) gloc_glog_shim::NullStream()
#define DCHECK_GE(a, b) while (false) gloc_glog_shim::NullStream()
#define DCHECK_GT(a, b) while (false) gloc_glog_shim::NullStream()
#define LOG(severity) gloc_glog_shim::NullStream()
#endif
