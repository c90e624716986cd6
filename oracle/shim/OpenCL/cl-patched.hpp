// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for the reference's vendored
// <OpenCL/cl-patched.hpp>, which src/utils.cpp:5 includes solely for the
// `cl_uint` typedef used by a few std::vector<cl_uint> helpers.  With this
// header on the include path the reference's src/utils.cpp compiles
// unmodified without an OpenCL SDK or Boost.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>
typedef uint32_t cl_uint;
