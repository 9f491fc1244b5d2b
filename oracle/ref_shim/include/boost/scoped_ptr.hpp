#pragma once
#include <memory>
namespace boost { template <class T> class scoped_ptr : public std::unique_ptr<T> { public: using std::unique_ptr<T>::unique_ptr; }; }
