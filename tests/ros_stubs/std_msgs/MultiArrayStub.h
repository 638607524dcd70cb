#ifndef PP_STUB_MULTIARRAY_H
#define PP_STUB_MULTIARRAY_H
#include <string>
#include <vector>
#include "boost/shared_ptr.hpp"
namespace std_msgs
{
    struct MultiArrayDimension { std::string label; unsigned size = 0, stride = 0; };
    struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; unsigned data_offset = 0; };
    template <class T> struct MultiArrayT
    {
        MultiArrayLayout layout; std::vector<T> data;
        typedef boost::shared_ptr<const MultiArrayT<T>> ConstPtr;
    };
}
#endif
