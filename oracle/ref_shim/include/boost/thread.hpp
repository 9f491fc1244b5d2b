#pragma once
#include <thread>
namespace boost { class thread : public std::thread { public: using std::thread::thread; }; }
