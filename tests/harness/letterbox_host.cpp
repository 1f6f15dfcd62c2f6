// TEST INFRASTRUCTURE: runs the letterbox arithmetic of head_detector_b200/csrc/letterbox_core.h (the
// exact functions the CUDA kernel calls) on the CPU, pixel by pixel, so that it can be compared with
// cv2 / the oracle in the GPU-less build container.  Not part of the product library.
#include <cstdint>
#include <vector>

#include "../../head_detector_b200/csrc/letterbox_core.h"

extern "C" int lb_host_letterbox(const uint8_t* src, int h, int w, int S, uint8_t* out, int* geom /*new_h,new_w,pad_x,pad_y*/) {
  vgh::LetterboxImage g;
  if (!vgh::letterbox_geometry(h, w, S, &g)) return 1;
  std::vector<int32_t> xo(g.new_w), yo(g.new_h);
  std::vector<int16_t> xa(g.new_w * 8), ya(g.new_h * 8);
  vgh::lanczos4_axis_tables(w, g.new_w, xo.data(), xa.data());
  vgh::lanczos4_axis_tables(h, g.new_h, yo.data(), ya.data());
  for (int y = 0; y < S; ++y)
    for (int x = 0; x < S; ++x) {
      uint8_t* o = out + (static_cast<int64_t>(y) * S + x) * 3;
      const int rx = x - g.pad_x, ry = y - g.pad_y;
      if (rx >= 0 && rx < g.new_w && ry >= 0 && ry < g.new_h)
        vgh::lanczos4_pixel_rgb(src, h, w, xo[rx], &xa[rx * 8], yo[ry], &ya[ry * 8], o);
      else
        o[0] = vgh::kPadR, o[1] = vgh::kPadG, o[2] = vgh::kPadB;
    }
  geom[0] = g.new_h; geom[1] = g.new_w; geom[2] = g.pad_x; geom[3] = g.pad_y;
  return 0;
}

// axis tables alone (coefficient parity with the oracle / worst-case accumulator bound)
extern "C" void lb_host_axis_tables(int src, int dst, int32_t* ofs, int16_t* coef) { vgh::lanczos4_axis_tables(src, dst, ofs, coef); }
