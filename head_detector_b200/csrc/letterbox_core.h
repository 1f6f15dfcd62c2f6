// Letterbox arithmetic shared by the CUDA kernel (letterbox.cu) and the CPU test harness
// (tests/harness/letterbox_host.cpp): integer-exact restatement of what the reference's
// HeadDetector._transform_image does through OpenCV (head_detector/detector.py:40-52):
//
//   cv2.resize(image, (new_w, new_h), interpolation=cv2.INTER_LANCZOS4)          detector.py:47
//   cv2.copyMakeBorder(..., pad_h//2, ..., pad_w//2, ..., BORDER_CONSTANT, value=127)  detector.py:48-50
//
// OpenCV (third party; the reference pins opencv-contrib-python-headless==4.9.0.80) resizes 8-bit
// images in fixed point: per destination column / row an 8-tap Lanczos-4 kernel whose float
// coefficients are rounded to int16 with scale 2^11, a horizontal pass in exact int32, a vertical
// pass in exact int32, then (v + 2^21) >> 22 saturated to uint8; source taps outside the image
// replicate the border pixel.  Because every step after the coefficient table is integer
// arithmetic without overflow, evaluating a destination pixel directly (64 taps) gives the same
// bits as OpenCV's two-pass evaluation.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define VGH_HD __host__ __device__ __forceinline__
#define VGH_UNROLL _Pragma("unroll")
#else
#define VGH_HD inline
#define VGH_UNROLL
#endif

namespace vgh {

constexpr int kLanczosTaps = 8;
constexpr int kResizeCoefBits = 11;  // OpenCV INTER_RESIZE_COEF_BITS
// detector.py:50 passes `value=127` to cv2.copyMakeBorder: a Python scalar becomes cv::Scalar(127, 0, 0, 0),
// so the border of the RGB frame is (127, 0, 0), not grey.  Reproduced as the reference has it.
constexpr uint8_t kPadR = 127, kPadG = 0, kPadB = 0;

// Geometry of one image inside the letterboxed S x S frame (detector.py:41-52).
struct LetterboxImage {
  int64_t src_off;    // byte offset of the image in the packed source buffer (rows are width*3 bytes)
  int32_t h, w;       // source size
  int32_t new_h, new_w;
  int32_t pad_x, pad_y;  // pad_w // 2, pad_h // 2
  int32_t xtab, ytab;    // first entry of this image's column / row tables
};

// detector.py:41-46.  Python: int(w * S / h) = truncation of the double quotient.
inline bool letterbox_geometry(int h, int w, int S, LetterboxImage* g) {
  if (h <= 0 || w <= 0 || S <= 0) return false;
  if (h > w) {
    g->new_h = S;
    g->new_w = static_cast<int>(static_cast<double>(static_cast<int64_t>(w) * S) / h);
  } else {
    g->new_h = static_cast<int>(static_cast<double>(static_cast<int64_t>(h) * S) / w);
    g->new_w = S;
  }
  g->h = h;
  g->w = w;
  g->pad_x = (S - g->new_w) / 2;
  g->pad_y = (S - g->new_h) / 2;
  return g->new_h > 0 && g->new_w > 0;  // cv2.resize raises on an empty destination
}

// OpenCV interpolateLanczos4: 8 float weights for fractional position x in [0,1).
inline void lanczos4_weights(float x, float* coeffs) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[8][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  static const double kPi = 3.1415926535897932384626433832795;
  float sum = 0.f;
  const double y0 = -(x + 3) * kPi * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
  for (int i = 0; i < 8; ++i) {
    const float y0_ = (x + 3 - i);
    if (std::fabs(y0_) >= 1e-6f) {
      const double y = -y0_ * kPi * 0.25;
      coeffs[i] = static_cast<float>((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    } else {
      coeffs[i] = 1e30f;  // x ~ 0 or ~ 1: the tap that sits on the sample takes all the weight
    }
    sum += coeffs[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; ++i) coeffs[i] *= sum;
}

// Tables of one axis of cv::resize(INTER_LANCZOS4) for 8-bit data: first tap index (may be < 0 or
// reach past the end: taps are clamped when used) and the 8 int16 weights per destination index.
inline void lanczos4_axis_tables(int src, int dst, int32_t* ofs, int16_t* coef) {
  const double inv_scale = static_cast<double>(dst) / src;
  const double scale = 1. / inv_scale;
  for (int d = 0; d < dst; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    const int s = static_cast<int>(std::floor(f));
    f -= s;
    float c[8];
    lanczos4_weights(f, c);
    ofs[d] = s - 3;
    for (int k = 0; k < 8; ++k) {
      // saturate_cast<short>(float): round half to even (cvRound), then clamp
      long r = std::lrintf(c[k] * static_cast<float>(1 << kResizeCoefBits));
      coef[d * 8 + k] = static_cast<int16_t>(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
    }
  }
}

VGH_HD int lb_clamp(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// One destination pixel (3 channels) of the resized image.  x0 / y0 = first tap (ofs tables),
// ax / ay = the 8 weights of this column / row.
VGH_HD void lanczos4_pixel_rgb(const uint8_t* src, int h, int w, int x0, const int16_t* ax, int y0, const int16_t* ay,
                               uint8_t* out) {
  int xi[8];
VGH_UNROLL
  for (int k = 0; k < 8; ++k) xi[k] = lb_clamp(x0 + k, w - 1) * 3;
  int acc0 = 0, acc1 = 0, acc2 = 0;
VGH_UNROLL
  for (int ky = 0; ky < 8; ++ky) {
    const uint8_t* row = src + static_cast<int64_t>(lb_clamp(y0 + ky, h - 1)) * w * 3;
    int h0 = 0, h1 = 0, h2 = 0;
VGH_UNROLL
    for (int kx = 0; kx < 8; ++kx) {
      const uint8_t* p = row + xi[kx];
      const int a = ax[kx];
      h0 += p[0] * a;
      h1 += p[1] * a;
      h2 += p[2] * a;
    }
    const int b = ay[ky];
    acc0 += h0 * b;
    acc1 += h1 * b;
    acc2 += h2 * b;
  }
  constexpr int kShift = 2 * kResizeCoefBits, kDelta = 1 << (kShift - 1);
  const int v0 = (acc0 + kDelta) >> kShift, v1 = (acc1 + kDelta) >> kShift, v2 = (acc2 + kDelta) >> kShift;
  out[0] = static_cast<uint8_t>(v0 < 0 ? 0 : (v0 > 255 ? 255 : v0));
  out[1] = static_cast<uint8_t>(v1 < 0 ? 0 : (v1 > 255 ? 255 : v1));
  out[2] = static_cast<uint8_t>(v2 < 0 ? 0 : (v2 > 255 ? 255 : v2));
}

}  // namespace vgh
