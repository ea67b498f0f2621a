// Force-included (-include) ahead of every reference translation unit built for
// oracle/_ref.  Supplies only what MSVC's lax two-phase lookup let the reference omit.
#pragma once
// Model.cpp:29-44 defines operator< for tinyobj::index_t inside namespace std AFTER <map> has
// been parsed; g++ needs the declaration to be visible before std::less is defined.
namespace tinyobj { struct index_t; }
namespace std { inline bool operator<(const tinyobj::index_t&, const tinyobj::index_t&); }
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <algorithm>
#include <string>
// nvcc makes the float overloads of the <cmath> functions visible at global scope for device
// code (crt/math_functions.hpp), so the reference's unqualified sqrt(x)/fmax(a,b)/cos(x)/... on
// float arguments are float operations on the GPU.  Plain g++ would bind them to the C double
// functions instead; these using-declarations restore the device's overload resolution.
using std::isfinite;
using std::abs;
using std::fabs;
using std::sqrt;
using std::fmax;
using std::fmin;
using std::sin;
using std::cos;
using std::tan;
using std::acos;
using std::atan2;
using std::pow;
using std::exp;
using std::log;
using std::floor;
inline float Sqr(float x) { return x * x; }   // sample.h:143 uses an undeclared Sqr
