#pragma once
#include <vector>
namespace google { namespace protobuf {
template <class T> class RepeatedField : public std::vector<T> {
 public:
  const T& Get(int i) const { return (*this)[i]; }
  void Add(const T& v) { this->push_back(v); }
  T* mutable_data() { return this->data(); }
  void Reserve(int n) { this->reserve(n); }
  void Clear() { this->clear(); }
};
template <class T> using RepeatedPtrField = RepeatedField<T>;
} }
