// Build shim for the oracle only (test infrastructure, not product code).
// The real Boost is absent from this image; karto_sdk only needs these names to
// exist so its serialize() member templates parse. Nothing here does any work.
#pragma once
#include <cstddef>
#include <shared_mutex>
#include <mutex>
#include <thread>

#define BOOST_VERSION 107400
#define BOOST_SERIALIZATION_NVP(x) x
#define BOOST_SERIALIZATION_BASE_OBJECT_NVP(B) (*static_cast<B *>(this))
#define BOOST_SERIALIZATION_ASSUME_ABSTRACT(T)
#define BOOST_CLASS_EXPORT(T)
#define BOOST_CLASS_EXPORT_KEY(T)
#define BOOST_CLASS_EXPORT_IMPLEMENT(T)

namespace boost {
namespace serialization {
class access {};
template <class T> inline T & make_nvp(const char *, T & t) { return t; }
template <class T> inline int make_array(T *, std::size_t) { return 0; }
}  // namespace serialization
namespace archive {
enum archive_flags { no_header = 1, no_codecvt = 2 };
struct shim_archive_base {
  template <class T> shim_archive_base & operator&(const T &) { return *this; }
  template <class T> shim_archive_base & operator<<(const T &) { return *this; }
  template <class T> shim_archive_base & operator>>(T &) { return *this; }
};
struct binary_oarchive : shim_archive_base {
  struct is_loading { static const bool value = false; };
  struct is_saving { static const bool value = true; };
  template <class S> explicit binary_oarchive(S &, unsigned = 0) {}
};
struct binary_iarchive : shim_archive_base {
  struct is_loading { static const bool value = true; };
  struct is_saving { static const bool value = false; };
  template <class S> explicit binary_iarchive(S &, unsigned = 0) {}
};
}  // namespace archive
typedef std::mutex mutex;
}  // namespace boost
