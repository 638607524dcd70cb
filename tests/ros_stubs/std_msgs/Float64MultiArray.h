#ifndef PP_STUB_F64MA_H
#define PP_STUB_F64MA_H
#include "MultiArrayStub.h"
namespace std_msgs { typedef MultiArrayT<double> Float64MultiArray; }
#endif
