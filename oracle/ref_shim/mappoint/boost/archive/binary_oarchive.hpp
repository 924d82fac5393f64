// TEST INFRASTRUCTURE: see ../serialization/serialization.hpp.  binary_oarchive = a byte recorder: arithmetic values and arrays are
// appended raw, class types go through their own save() (member, via access) or the free boost::serialization::save overload
// (cv::Mat, include/MapPoint.h:216-231), cv::KeyPoint through its free serialize() (:203-213).
#pragma once
#include "../serialization/serialization.hpp"
namespace cv { class Mat; struct KeyPoint; }
namespace boost { namespace serialization {
template <class Ar> void save(Ar& ar, const ::cv::Mat& m, const unsigned int version);
template <class Ar> void load(Ar& ar, ::cv::Mat& m, const unsigned int version);
template <class Ar> void serialize(Ar& ar, ::cv::KeyPoint& k, const unsigned int version);
}}
namespace boost { namespace archive {
class binary_oarchive {
public:
    std::vector<unsigned char> bytes;
    std::vector<std::string> fields;                // one tag per appended item: "u8" (bytes of an arithmetic value), "a<N>" (raw array)
    template <class T> typename std::enable_if<std::is_arithmetic<T>::value, binary_oarchive&>::type operator&(const T& v) {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(&v);
        bytes.insert(bytes.end(), p, p + sizeof(T));
        fields.push_back((std::is_floating_point<T>::value ? "f" : std::is_same<T, bool>::value ? "b" : std::is_signed<T>::value ? "i" : "u") +
                         std::to_string(sizeof(T)));
        return *this;
    }
    template <class T> binary_oarchive& operator&(const boost::serialization::array_wrapper<T>& a) {
        const unsigned char* p = reinterpret_cast<const unsigned char*>(a.p);
        bytes.insert(bytes.end(), p, p + a.n * sizeof(T));
        fields.push_back("a" + std::to_string(a.n * sizeof(T)));
        return *this;
    }
    binary_oarchive& operator&(const ::cv::Mat& m) { fields.push_back("Mat{"); boost::serialization::save(*this, m, 0u); fields.push_back("}"); return *this; }
    template <class T> binary_oarchive& operator<<(const T& v) { return *this & v; }
};
}}
