// boost iostreams -- TEST INFRASTRUCTURE ONLY (oracle/).  3d/port.h defines two inline gzip
// helpers that nothing on the scan-match path calls; these stubs let the header parse.
#ifndef GLOC_ORACLE_BOOST_SHIM_H_
#define GLOC_ORACLE_BOOST_SHIM_H_
#include <cstddef>
namespace boost { namespace iostreams {
namespace zlib { const int best_speed = 1; }
struct gzip_compressor { gzip_compressor(int = 0) {} };
struct gzip_decompressor {};
struct filtering_ostream { template <typename T> void push(const T&) {} };
template <typename C> int back_inserter(C&) { return 0; }
template <typename S> void write(S&, const char*, std::size_t) {}
} }
#endif
