// Stand-in for gtest: the reference's net.cpp includes caffe/test/test_caffe_main.hpp, which only needs these names to parse.
#pragma once
namespace testing {
class Test { public: virtual ~Test() {} };
template <typename... T> struct Types {};
}  // namespace testing
