// TEST INFRASTRUCTURE (oracle/_ref MapPoint build).  Boost is not in this image; these stand-ins supply just the names the
// reference's include/MapPoint.h and src/MapPoint.cc mention so that both compile UNMODIFIED.  The archives are recording sinks /
// sources of raw bytes (no Boost framing): enough to run MapPoint::save / load, not a statement of Boost's file format.
#pragma once
#include <cstddef>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>
#define BOOST_SERIALIZATION_SPLIT_FREE(T)
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
namespace boost { namespace serialization {
class access {
public:
    template <class Ar, class T> static void save(Ar& ar, const T& t, unsigned v) { t.save(ar, v); }
    template <class Ar, class T> static void load(Ar& ar, T& t, unsigned v) { t.load(ar, v); }
};
template <class T> struct array_wrapper { T* p; std::size_t n; };
template <class T> inline array_wrapper<T> make_array(T* p, std::size_t n) { array_wrapper<T> a = {p, n}; return a; }
template <class Ar, class T> inline void split_member(Ar& ar, T& t, const unsigned int v);
}}
