// TEST INFRASTRUCTURE: see ../serialization/serialization.hpp.  binary_iarchive = the reader of what binary_oarchive recorded.
#pragma once
#include "binary_oarchive.hpp"
namespace boost { namespace archive {
class binary_iarchive {
public:
    const unsigned char* cur = nullptr;
    const unsigned char* end = nullptr;
    bool failed = false;
    void take(void* dst, std::size_t n) {
        if ((std::size_t)(end - cur) < n) { failed = true; std::memset(dst, 0, n); return; }
        std::memcpy(dst, cur, n);
        cur += n;
    }
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value, binary_iarchive&>::type operator&(T& v) { take(&v, sizeof(T)); return *this; }
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value, binary_iarchive&>::type operator&(const T& v) { take(const_cast<T*>(&v), sizeof(T)); return *this; }
    template <class T> binary_iarchive& operator&(const boost::serialization::array_wrapper<T>& a) { take((void*)a.p, a.n * sizeof(T)); return *this; }
    binary_iarchive& operator&(::cv::Mat& m) { boost::serialization::load(*this, m, 0u); return *this; }
    binary_iarchive& operator&(const ::cv::Mat& m) { boost::serialization::load(*this, const_cast< ::cv::Mat&>(m), 0u); return *this; }
    template <class T> binary_iarchive& operator>>(T& v) { return *this & v; }
};
}}
