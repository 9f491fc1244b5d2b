#pragma once
#include <cmath>
namespace boost { namespace math { template <class T> inline T nextafter(const T& a, const T& b) { return std::nextafter(a, b); } } }
