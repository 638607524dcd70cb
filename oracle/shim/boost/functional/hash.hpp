// TEST INFRASTRUCTURE ONLY (oracle build shim).
// Boost is not installed in this image; the reference's Node2D.h / Node3D.h include
// <boost/functional/hash.hpp> only for boost::hash_combine (Node2D.h:79-80, Node3D.h:94-96).
// The hash value only decides unordered_set bucket placement, never a numeric result
// ("parity unpinned, immaterial"), so the classic golden-ratio combiner is used.
#ifndef PP_ORACLE_BOOST_HASH_SHIM
#define PP_ORACLE_BOOST_HASH_SHIM
#include <cstddef>
#include <functional>
namespace boost
{
    template <class T>
    inline void hash_combine(std::size_t& seed, const T& v)
    {
        seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    }
}
#endif
