#pragma once
// Per-DEVICE bookkeeping of function attributes.  cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (function,
// device) pair, so a process that drives several GPUs (or creates a detector on cuda:1 after cuda:0) has to opt in on
// each of them; two detector handles per GPU also call the launch helpers from different host threads.
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>

namespace vgh {

constexpr int kMaxDevices = 64;

struct SmemOptIn {
  std::atomic<size_t> bytes[kMaxDevices];
  std::mutex mu;
  SmemOptIn() {
    for (auto& b : bytes) b.store(0);
  }
};

// Makes sure `kernel` may be launched with `bytes` of dynamic shared memory on the CURRENT device.
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, SmemOptIn& st, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (bytes <= st.bytes[dev].load(std::memory_order_acquire)) return cudaSuccess;
  std::lock_guard<std::mutex> lock(st.mu);
  if (bytes <= st.bytes[dev].load(std::memory_order_relaxed)) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e == cudaSuccess) st.bytes[dev].store(bytes, std::memory_order_release);
  return e;
}

// SM count of the current device (cached per device).
inline int device_sm_count() {
  static std::atomic<int> sms[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n > 0) return n;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  sms[dev].store(n, std::memory_order_relaxed);
  return n;
}

}  // namespace vgh
