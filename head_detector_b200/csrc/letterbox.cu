// Device letterbox: cv2.resize(INTER_LANCZOS4) to longest side S + centred
// constant border (cv2's Scalar(127) = (127,0,0) per RGB pixel, as the reference produces it),
// bit-exact with the reference's host path (head_detector/detector.py:40-52), for a batch of
// differently sized RGB images in one launch.  The arithmetic lives in
// letterbox_core.h (shared with the CPU test harness); this file is the batching around it.
//
// HBM-bound by design: algorithmic bytes = source image bytes read once + S*S*3 written.  One thread
// per destination pixel evaluates the 8x8 taps of its three channels from L1/L2 (neighbouring threads
// share 7/8 of their source columns); the per-column / per-row int16 weight tables are built on the
// host (sin/cos in double precision like OpenCV; cached per (source, destination) extent pair) and
// staged through stream-ordered scratch memory.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "letterbox.cuh"
#include "letterbox_core.h"

namespace vgh {

constexpr int kLbTileX = 32, kLbTileY = 8;

__global__ void __launch_bounds__(kLbTileX* kLbTileY) letterbox_kernel(const uint8_t* __restrict__ src, const LetterboxImage* __restrict__ imgs,
                                                                       const int32_t* __restrict__ ofs, const int16_t* __restrict__ coef, int S,
                                                                       uint8_t* __restrict__ out) {
  const LetterboxImage g = imgs[blockIdx.z];
  const int x = blockIdx.x * kLbTileX + threadIdx.x;
  const int y = blockIdx.y * kLbTileY + threadIdx.y;
  if (x >= S || y >= S) return;
  uint8_t px[3] = {kPadR, kPadG, kPadB};
  const int rx = x - g.pad_x, ry = y - g.pad_y;
  if (rx >= 0 && rx < g.new_w && ry >= 0 && ry < g.new_h) {
    // 8 int16 weights = one 16-byte load per axis
    const int4 wx = __ldg(reinterpret_cast<const int4*>(coef) + g.xtab + rx);
    const int4 wy = __ldg(reinterpret_cast<const int4*>(coef) + g.ytab + ry);
    lanczos4_pixel_rgb(src + g.src_off, g.h, g.w, __ldg(ofs + g.xtab + rx), reinterpret_cast<const int16_t*>(&wx),
                       __ldg(ofs + g.ytab + ry), reinterpret_cast<const int16_t*>(&wy), px);
  }
  uint8_t* o = out + ((static_cast<size_t>(blockIdx.z) * S + y) * S + x) * 3;
  o[0] = px[0];
  o[1] = px[1];
  o[2] = px[2];
}

// Axis tables are pure functions of (source extent, destination extent): built once per pair (video
// frames and photo batches repeat a handful of sizes) and shared by every image / axis that uses the pair.
struct AxisTable {
  std::vector<int32_t> ofs;
  std::vector<int16_t> coef;
};
static AxisTable axis_table(int src, int dst) {  // by value: copied under the lock, the cache may be reset by another thread
  static std::mutex mu;
  static std::map<std::pair<int, int>, AxisTable> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find({src, dst});
  if (it == cache.end()) {
    if (cache.size() > 256) cache.clear();  // bounded
    AxisTable t;
    t.ofs.resize(dst);
    t.coef.resize(static_cast<size_t>(dst) * 8);
    lanczos4_axis_tables(src, dst, t.ofs.data(), t.coef.data());
    it = cache.emplace(std::make_pair(src, dst), std::move(t)).first;
  }
  return it->second;
}

int letterbox_launch(const uint8_t* src_dev, const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n, int S,
                     uint8_t* out_dev, float* xform_host, cudaStream_t stream, char* err, size_t errlen) {
  if (n == 0) return 0;
  std::vector<LetterboxImage> imgs(n);
  std::map<std::pair<int, int>, int> placed;  // (src, dst) -> first table entry of this call
  std::vector<int32_t> ofs;
  std::vector<int16_t> coef;
  auto place = [&](int src, int dst) {
    auto it = placed.find({src, dst});
    if (it != placed.end()) return it->second;
    const int at = static_cast<int>(ofs.size());
    const AxisTable t = axis_table(src, dst);
    ofs.insert(ofs.end(), t.ofs.begin(), t.ofs.end());
    coef.insert(coef.end(), t.coef.begin(), t.coef.end());
    placed[{src, dst}] = at;
    return at;
  };
  for (int i = 0; i < n; ++i) {
    LetterboxImage& g = imgs[i];
    if (!letterbox_geometry(heights[i], widths[i], S, &g)) {
      snprintf(err, errlen, "letterbox: image %d (%dx%d) has an empty resized extent at size %d", i, heights[i], widths[i], S);
      return 1;
    }
    g.src_off = offsets[i];
    g.xtab = place(g.w, g.new_w);
    g.ytab = place(g.h, g.new_h);
    if (xform_host) {
      xform_host[i * 3 + 0] = static_cast<float>(g.pad_x);
      xform_host[i * 3 + 1] = static_cast<float>(g.pad_y);
      const int longest = g.h > g.w ? g.h : g.w;
      xform_host[i * 3 + 2] = static_cast<float>(static_cast<double>(S) / longest);  // detector.py:46
    }
  }
  // one stream-ordered device scratch block [coef (16 B per entry) | ofs | image records], filled by ONE copy
  // from a per-thread pinned staging buffer (an event guards its reuse by the next call of this thread)
  const size_t coef_bytes = coef.size() * sizeof(int16_t);
  const size_t ofs_bytes = (ofs.size() * sizeof(int32_t) + 15) & ~static_cast<size_t>(15);
  const size_t img_bytes = imgs.size() * sizeof(LetterboxImage);
  const size_t total_bytes = coef_bytes + ofs_bytes + img_bytes;
  struct Staging {
    uint8_t* host = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
  };
  static thread_local Staging st;
  cudaError_t e = cudaSuccess;
  if (st.done) e = cudaEventSynchronize(st.done);  // the previous call's copy has left the buffer
  if (e == cudaSuccess && st.cap < total_bytes) {
    if (st.host) cudaFreeHost(st.host);
    st.host = nullptr;
    st.cap = 0;
    e = cudaMallocHost(reinterpret_cast<void**>(&st.host), total_bytes * 2);
    if (e == cudaSuccess) st.cap = total_bytes * 2;
  }
  if (e == cudaSuccess && !st.done) e = cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming);
  uint8_t* scratch = nullptr;
  if (e == cudaSuccess) {
    memcpy(st.host, coef.data(), coef_bytes);
    memcpy(st.host + coef_bytes, ofs.data(), ofs.size() * sizeof(int32_t));
    memcpy(st.host + coef_bytes + ofs_bytes, imgs.data(), img_bytes);
    e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), total_bytes, stream);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(scratch, st.host, total_bytes, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaEventRecord(st.done, stream);
  if (e == cudaSuccess) {
    const dim3 grid((S + kLbTileX - 1) / kLbTileX, (S + kLbTileY - 1) / kLbTileY, n);
    letterbox_kernel<<<grid, dim3(kLbTileX, kLbTileY), 0, stream>>>(
        src_dev, reinterpret_cast<const LetterboxImage*>(scratch + coef_bytes + ofs_bytes),
        reinterpret_cast<const int32_t*>(scratch + coef_bytes), reinterpret_cast<const int16_t*>(scratch), S, out_dev);
    e = cudaGetLastError();
  }
  if (scratch) {
    const cudaError_t e2 = cudaFreeAsync(scratch, stream);
    if (e == cudaSuccess) e = e2;
  }
  if (e != cudaSuccess) {
    snprintf(err, errlen, "letterbox: %s", cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

}  // namespace vgh
