#pragma once
#include <random>
namespace boost { template <class T = int> class uniform_int : public std::uniform_int_distribution<T> { public: using std::uniform_int_distribution<T>::uniform_int_distribution; }; }
