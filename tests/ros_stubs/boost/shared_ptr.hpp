// TEST INFRASTRUCTURE stub: boost::shared_ptr spelled with the standard one (ROS message ConstPtr typedefs).
#ifndef PP_STUB_BOOST_SHARED_PTR
#define PP_STUB_BOOST_SHARED_PTR
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
#endif
