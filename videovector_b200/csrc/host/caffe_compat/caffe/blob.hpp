// Blob<Dtype> / SyncedMemory (ref: include/caffe/blob.hpp:19-141, include/caffe/syncedmem.hpp:40-68,
// src/caffe/syncedmem.cpp:21-109): NCHW contiguous data + diff, lazily allocated, zero-filled on first
// touch, host<->device mirror with a HEAD state machine; mutable_* marks the side dirty.
#pragma once
#include "caffe/common.hpp"

namespace caffe {

class SyncedMemory {
 public:
  enum SyncedHead { UNINITIALIZED, HEAD_AT_CPU, HEAD_AT_GPU, SYNCED };
  explicit SyncedMemory(size_t size) : cpu_ptr_(nullptr), gpu_ptr_(nullptr), size_(size), head_(UNINITIALIZED), own_gpu_(true) {}
  ~SyncedMemory();
  const void* cpu_data();
  const void* gpu_data();
  void* mutable_cpu_data();
  void* mutable_gpu_data();
  // wrap device memory owned by someone else (the fused trainer's parameter / gradient buffers)
  void set_gpu_data(void* data);
  SyncedHead head() const { return head_; }
  size_t size() const { return size_; }
 private:
  void to_cpu();
  void to_gpu();
  void* cpu_ptr_; void* gpu_ptr_; size_t size_; SyncedHead head_; bool own_gpu_;
  SyncedMemory(const SyncedMemory&) = delete;
  SyncedMemory& operator=(const SyncedMemory&) = delete;
};

template <typename Dtype>
class Blob {
 public:
  Blob() : num_(0), channels_(0), height_(0), width_(0), count_(0), capacity_(0) {}
  Blob(const int num, const int channels, const int height, const int width) : capacity_(0) { Reshape(num, channels, height, width); }
  void Reshape(const int num, const int channels, const int height, const int width);
  void ReshapeLike(const Blob& other) { Reshape(other.num(), other.channels(), other.height(), other.width()); }
  inline int num() const { return num_; }
  inline int channels() const { return channels_; }
  inline int height() const { return height_; }
  inline int width() const { return width_; }
  inline int count() const { return count_; }
  inline int offset(const int n, const int c = 0, const int h = 0, const int w = 0) const {
    return ((n * channels_ + c) * height_ + h) * width_ + w;
  }
  const Dtype* cpu_data() const;
  const Dtype* gpu_data() const;
  const Dtype* cpu_diff() const;
  const Dtype* gpu_diff() const;
  Dtype* mutable_cpu_data();
  Dtype* mutable_gpu_data();
  Dtype* mutable_cpu_diff();
  Dtype* mutable_gpu_diff();
  inline Dtype data_at(const int n, const int c, const int h, const int w) const { return cpu_data()[offset(n, c, h, w)]; }
  inline Dtype diff_at(const int n, const int c, const int h, const int w) const { return cpu_diff()[offset(n, c, h, w)]; }
  void Update();                                   // data -= diff (ref: blob.cpp:113-136)
  void ShareData(const Blob& other) { CHECK_EQ(count_, other.count()); data_ = other.data_; }
  void ShareDiff(const Blob& other) { CHECK_EQ(count_, other.count()); diff_ = other.diff_; }
  void CopyFrom(const Blob<Dtype>& source, bool copy_diff = false, bool reshape = false);
  // (de)serialisation as a BlobProto message tree (ref: blob.cpp:243-256, 305-322)
  void FromProto(const class PbMsg& proto);
  void ToProto(class PbMsg* proto, bool write_diff = false) const;
  Dtype asum_data() const;
  Dtype asum_diff() const;
  const shared_ptr<SyncedMemory>& data() const { return data_; }
  const shared_ptr<SyncedMemory>& diff() const { return diff_; }
  // alias externally owned device memory (count must match)
  void set_gpu_data(Dtype* p) { CHECK(data_); data_->set_gpu_data(p); }
  void set_gpu_diff(Dtype* p) { CHECK(diff_); diff_->set_gpu_data(p); }
 protected:
  shared_ptr<SyncedMemory> data_, diff_;
  int num_, channels_, height_, width_, count_, capacity_;
};

}  // namespace caffe
