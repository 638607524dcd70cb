#ifndef PP_STUB_F32MA_H
#define PP_STUB_F32MA_H
#include "MultiArrayStub.h"
namespace std_msgs { typedef MultiArrayT<float> Float32MultiArray; }
#endif
