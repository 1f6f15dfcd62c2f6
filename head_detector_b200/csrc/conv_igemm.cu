#include "conv_igemm.cuh"
#include "device_attr.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ptx.cuh"

namespace vgh {

static thread_local char g_conv_err[256] = "";
const char* conv_last_error() { return g_conv_err; }

constexpr int kBlockM = 128;
constexpr int kEpiWarps = 8;                   // 2 per TMEM lane quarter, each takes every other 16-column chunk
constexpr int kThreads = 64 + 32 * kEpiWarps;  // warp0 TMA, warp1 MMA/TMEM, warps 2..9 epilogue

// epilogue helpers: 16 channels of one pixel, activations stored as bf16 or fp16 (ConvLaunch::f16)
template <bool F16>
__device__ __forceinline__ void add_residual16(float (&f)[16], const uint4 (&rq)[2], float alpha) {
  const uint32_t w[8] = {rq[0].x, rq[0].y, rq[0].z, rq[0].w, rq[1].x, rq[1].y, rq[1].z, rq[1].w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 r = unpack16x2<F16>(w[i]);
    f[2 * i] = fmaf(alpha, r.x, f[2 * i]);
    f[2 * i + 1] = fmaf(alpha, r.y, f[2 * i + 1]);
  }
}
template <bool F16>
__device__ __forceinline__ void store_row16(const float (&f)[16], uint4* op) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack16x2<F16>(f[2 * i], f[2 * i + 1]);
  op[0] = make_uint4(w[0], w[1], w[2], w[3]);
  op[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Persistent, warp-specialised implicit-GEMM conv.  Each CTA (one per SM) walks work items
// item = blockIdx.x + i*gridDim.x, item -> (M group of `mt` 128-pixel tiles, N tile).  The shared-memory
// ring (TMA -> MMA) runs continuously across items; accumulators are double-buffered in TMEM when
// 2*mt*block_n <= 512 columns so the epilogue of item i overlaps the main loop of item i+1.
template <int BK>
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvLaunch p) {
  constexpr uint32_t kRowBytes = BK * 2;            // bytes of one K-slab row == swizzle span
  constexpr uint32_t kABytes = kBlockM * kRowBytes;  // one 128-pixel A tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * kRowBytes;
  const uint32_t a_stage_bytes = static_cast<uint32_t>(p.mt) * kABytes;  // mt A tiles per stage
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + p.stages * a_stage_bytes;
  const uint32_t bar_base = b_base + p.stages * b_bytes;  // 8-byte aligned (multiples of 1024)
  const uint32_t full_bar = bar_base;                     // [stages]
  const uint32_t empty_bar = bar_base + 8 * p.stages;     // [stages]
  const uint32_t tmem_full_bar = bar_base + 16 * p.stages;  // [2]
  const uint32_t tmem_empty_bar = tmem_full_bar + 16;       // [2]
  const uint32_t tmem_slot = tmem_empty_bar + 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Tiles: per image tiles_x * tiles_y rectangles of th rows.  When the map height is not a multiple of
  // th (p.tail_rows > 0), tiles_y counts only the FULL row groups and the left-over rows of
  // p.tail_imgs consecutive images are packed into one extra tile (4-D box [BK, tw, tail_rows,
  // tail_imgs]), which removes the mostly-empty last tile of every image (wave quantisation).
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n_full = tiles_per_img * p.B;
  const int total_tiles = n_full + p.n_tail_tiles;
  auto tile_coords = [&](int t, int& b_img, int& h0, int& w0) -> bool {  // returns true for a tail tile
    t = min(t, total_tiles - 1);  // tail of the last M group: reload the last tile, its epilogue is skipped
    if (t >= n_full) {
      const int u = t - n_full;
      const int bg = u / p.tiles_x;
      b_img = bg * p.tail_imgs;
      h0 = p.tiles_y * p.th;
      w0 = (u - bg * p.tiles_x) * p.tw;
      return true;
    }
    b_img = t / tiles_per_img;
    const int t_in = t - b_img * tiles_per_img;
    const int tyi = t_in / p.tiles_x;
    h0 = tyi * p.th;
    w0 = (t_in - tyi * p.tiles_x) * p.tw;
    return false;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    if (p.n_tail_tiles) tma_prefetch_desc(&p.tmOut);  // (normal kernel: tmOut holds the tail-tile input map)
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar + 8 * a, 1);
      mbar_init(tmem_empty_bar + 8 * a, kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_dyn(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));
  // everything above overlapped the previous kernel's tail (PDL); its outputs are needed from here on
  pdl_wait();
  pdl_launch_dependents();

  const int cblks = p.cin / BK;
  const int num_kb = p.ntaps * cblks;
  const uint32_t acc_cols = static_cast<uint32_t>(p.mt * p.block_n);  // TMEM columns of one accumulator stage

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      const uint32_t full_tile_bytes = static_cast<uint32_t>(p.tw * p.th) * kRowBytes;
      const uint32_t tail_tile_bytes = static_cast<uint32_t>(p.tw * p.tail_rows * p.tail_imgs) * kRowBytes;
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int mg = item / p.n_tiles;
        const int n0 = (item - mg * p.n_tiles) * p.block_n;
        int tb[4], th0[4], tw0[4];
        bool tail[4];
        uint32_t tx_bytes = b_bytes;
        for (int i = 0; i < p.mt; ++i) {
          tail[i] = tile_coords(mg * p.mt + i, tb[i], th0[i], tw0[i]);
          tx_bytes += tail[i] ? tail_tile_bytes : full_tile_bytes;
        }
        int tap = 0, cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          mbar_arrive_expect_tx(fb, tx_bytes);
          const int ty = tap / p.kw;
          const int tx = tap - ty * p.kw;
          for (int i = 0; i < p.mt; ++i)
            tma_load_4d(a_base + stage * a_stage_bytes + i * kABytes, tail[i] ? &p.tmOut : &p.tmA, fb, p.cin_off + cb * BK,
                        tw0[i] * p.stride + tx - p.pad, th0[i] * p.stride + ty - p.pad, tb[i]);
          tma_load_2d(b_base + stage * b_bytes, &p.tmB, fb, tap * p.cin + cb * BK, n0);
          if (++cb == cblks) {
            cb = 0;
            ++tap;
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer (one thread) =====================
      const uint32_t idesc = umma_idesc_16(kBlockM, static_cast<uint32_t>(p.block_n), p.f16 != 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        mbar_wait(tmem_empty_bar + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_base = tmem_acc + acc * acc_cols;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          const uint64_t b_desc = umma_smem_desc(b_base + stage * b_bytes, kRowBytes);
          for (int i = 0; i < p.mt; ++i) {
            const uint64_t a_desc = umma_smem_desc(a_base + stage * a_stage_bytes + i * kABytes, kRowBytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advance 16 elements (32 B) along K inside the swizzle span: +2 in 16-byte units
              umma_bf16(d_base + i * p.block_n, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(empty_bar + 8 * stage);  // frees the smem slot once these MMAs retire
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tmem_full_bar + 8 * acc);  // accumulators of this item complete
        if (++acc == p.acc_stages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are the ones this warp may read
    const int chunk0 = ew >> 2;    // two warps share a quarter: even / odd 16-column chunks
    const int r = quarter * 32 + lane;
    const int n_chunks = p.block_n >> 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int mg = item / p.n_tiles;
      const int n0 = (item - mg * p.n_tiles) * p.block_n;
      mbar_wait(tmem_full_bar + 8 * acc, acc_phase);
      tc_fence_after();
      for (int ti = 0; ti < p.mt; ++ti) {
        const int tile = mg * p.mt + ti;
        if (tile >= total_tiles) break;
        int b_img, h0, w0;
        const bool is_tail = tile_coords(tile, b_img, h0, w0);
        int ly, lx;
        bool row_ok;
        if (is_tail) {  // rows = (image, row, col) of the left-over rows of tail_imgs images
          const int per_img = p.tw * p.tail_rows;
          const int il = r / per_img;
          const int rr = r - il * per_img;
          ly = rr / p.tw;
          lx = rr - ly * p.tw;
          b_img += il;
          row_ok = (il < p.tail_imgs) && (b_img < p.B);
        } else {
          ly = r / p.tw;
          lx = r - ly * p.tw;
          row_ok = r < p.tw * p.th;
        }
        const int oh = h0 + ly, ow = w0 + lx;
        row_ok = row_ok && (oh < p.Ho) && (ow < p.Wo);
        int n_shift = 0;  // channel shift when the n-tile addresses a sub-pixel of the 2x2 transpose conv
        int ph = oh, pw = ow;
        if (p.up) {
          const int sub = n0 / p.up_cout;
          n_shift = sub * p.up_cout;
          ph = 2 * oh + (sub >> 1);
          pw = 2 * ow + (sub & 1);
        }
        const size_t out_pix = (static_cast<size_t>(b_img) * p.out_H + ph) * p.out_W + pw;
        const size_t out_off = out_pix * p.out_cstride + p.out_coff + (n0 - n_shift);
        const size_t res_off =
            ((static_cast<size_t>(b_img) * p.Ho + oh) * p.Wo + ow) * p.res_cstride + p.res_coff + n0;
        const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + acc * acc_cols + ti * p.block_n;
        const bool use_res = (p.res != nullptr) && row_ok && !p.split;
        // chunks of this warp: chunk0, chunk0+2, ...; handled four at a time so that the residual
        // loads of a whole group are in flight before the first one is consumed
        for (int cg = chunk0; cg < n_chunks; cg += 8) {
          uint4 rq[4][2];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int c = cg + 2 * g;
            rq[g][0] = make_uint4(0, 0, 0, 0);
            rq[g][1] = rq[g][0];
            if (use_res && c < n_chunks && n0 + c * 16 < p.n_total) {
              const uint4* rp = reinterpret_cast<const uint4*>(p.res + res_off + c * 16);
              rq[g][0] = __ldg(rp);
              rq[g][1] = __ldg(rp + 1);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int c = cg + 2 * g;
            if (c >= n_chunks) break;
            const int n = n0 + c * 16;
            uint32_t v[16];
            tmem_ld16(taddr + c * 16, v);
            tmem_ld_wait();
            if (!(row_ok && n < p.n_total)) continue;
            float f[16];
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(bp + i);
              f[4 * i] = __uint_as_float(v[4 * i]) + b4.x;
              f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
              f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z;
              f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            if (p.split) {
              // Parity mode (fp32-class arithmetic on the bf16 tensor pipe): an activation y is stored as three bf16 terms
              // h = bf16(y), m = bf16(y - h), l = bf16(y - h - m) (y == h + m + l exactly), laid out per 32-channel granule
              // as six planes [h|m|h|m|h|l]; the consumer's weights are packed [w_h|w_h|w_m|w_m|w_l|w_h] along K, so its
              // MMA accumulates x_h w_h + x_m w_h + x_h w_m + x_m w_m + x_h w_l + x_l w_h in fp32.
              const int nl = (n0 - n_shift) + c * 16;  // logical channel inside the producer's slice
              if (p.res != nullptr) {
                const int nr = n0 + c * 16;
                const __nv_bfloat16* rp = p.res + ((static_cast<size_t>(b_img) * p.Ho + oh) * p.Wo + ow) * p.res_cstride + p.res_coff +
                                          192 * (nr >> 5) + (nr & 31);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const float rh = __bfloat162float(rp[i]), rm = __bfloat162float(rp[32 + i]), rl = __bfloat162float(rp[160 + i]);
                  f[i] = fmaf(p.res_alpha, (rl + rm) + rh, f[i]);
                }
              }
              __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out) + out_pix * p.out_cstride + p.out_coff + 192 * (nl >> 5) + (nl & 31);
              uint32_t wh[8], wm[8], wl[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float y0 = f[2 * i], y1 = f[2 * i + 1];
                const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                y0 -= __bfloat162float(h.x);
                y1 -= __bfloat162float(h.y);
                const __nv_bfloat162 m = __floats2bfloat162_rn(y0, y1);
                y0 -= __bfloat162float(m.x);
                y1 -= __bfloat162float(m.y);
                const __nv_bfloat162 l = __floats2bfloat162_rn(y0, y1);
                wh[i] = *reinterpret_cast<const uint32_t*>(&h);
                wm[i] = *reinterpret_cast<const uint32_t*>(&m);
                wl[i] = *reinterpret_cast<const uint32_t*>(&l);
              }
              const uint4 h0 = make_uint4(wh[0], wh[1], wh[2], wh[3]), h1 = make_uint4(wh[4], wh[5], wh[6], wh[7]);
              const uint4 m0 = make_uint4(wm[0], wm[1], wm[2], wm[3]), m1 = make_uint4(wm[4], wm[5], wm[6], wm[7]);
              uint4* o4 = reinterpret_cast<uint4*>(op);  // planes are 32 channels = 4 uint4 apart
              o4[0] = h0;  o4[1] = h1;
              o4[4] = m0;  o4[5] = m1;
              o4[8] = h0;  o4[9] = h1;
              o4[12] = m0; o4[13] = m1;
              o4[16] = h0; o4[17] = h1;
              o4[20] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
              o4[21] = make_uint4(wl[4], wl[5], wl[6], wl[7]);
              continue;
            }
            if (p.res != nullptr) {
              if (p.f16) add_residual16<true>(f, rq[g], p.res_alpha);
              else add_residual16<false>(f, rq[g], p.res_alpha);
            }
            if (p.out_fp32) {
              float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + out_off + c * 16);
#pragma unroll
              for (int i = 0; i < 4; ++i) op[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            } else {
              uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + out_off + c * 16);
              if (p.f16) store_row16<true>(f, op);
              else store_row16<false>(f, op);
            }
          }
        }
      }
      // this warp is done reading the accumulator stage: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar + 8 * acc);
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_dyn(tmem_acc, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || ptr == nullptr) {
    snprintf(g_conv_err, sizeof(g_conv_err), "cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// PDL can be switched off with VGGHEADS_B200_NO_PDL=1 (debugging aid)
bool conv_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VGGHEADS_B200_NO_PDL");
    on = (e && e[0] == '1') ? 0 : 1;
  }
  return on == 1;
}

static int num_sms() { return device_sm_count(); }

int conv_default_mt(int block_n) {
  // M tiles per CTA sharing one weight stream, such that two accumulator stages still fit in the
  // 512 TMEM columns (epilogue/main-loop overlap): N<=64 -> 4, N<=128 -> 2, else 1
  int mt = 256 / block_n;
  if (mt > 4) mt = 4;
  if (mt < 1) mt = 1;
  if (mt == 3) mt = 2;
  return mt;
}

int conv_pick_stages(int block_n, int bk, int mt) {
  const int stage = (kBlockM * mt + block_n) * bk * 2;
  int s = (200 * 1024) / stage;  // one persistent CTA per SM owns the shared memory
  if (s < 2) s = 2;
  if (s > 8) s = 8;
  return s;
}

size_t conv_smem_bytes(const ConvLaunch& L, int bk) {
  return 1024 + static_cast<size_t>(L.stages) * (kBlockM * L.mt + L.block_n) * bk * 2 + 16 * L.stages + 64;
}

// fills the derived scheduling fields (call after tiles / block_n / mt are set)
void conv_finalize(ConvLaunch& L) {
  if (L.swap) {  // one item = one (tw x th)-pixel tile x one channel group; clusters take n_cl tiles of a group at a time
    L.mt = 1;
    L.n_tiles = 1;
    if (L.cluster < 1) L.cluster = 1;
    const int tiles = L.tiles_x * L.tiles_y * L.B;
    L.num_items = ((tiles + L.cluster - 1) / L.cluster) * L.ngroups * L.cluster;
    const int npix = L.tw * L.th;
    L.acc_stages = (2 * npix <= 512) ? 2 : 1;
    int cols = 32;
    while (cols < L.acc_stages * npix) cols <<= 1;
    L.tmem_cols = cols;
    return;
  }
  L.n_tiles = (L.n_total + L.block_n - 1) / L.block_n;
  const int m_groups = (L.tiles_x * L.tiles_y * L.B + L.n_tail_tiles + L.mt - 1) / L.mt;
  L.num_items = m_groups * L.n_tiles;
  L.acc_stages = (2 * L.mt * L.block_n <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < L.acc_stages * L.mt * L.block_n) cols <<= 1;
  L.tmem_cols = cols;
}

int conv_make_tensor_maps(ConvLaunch& L, const void* in_base, int in_C, int in_H, int in_W, const void* w_base,
                          int k_total, int n_pad, int bk) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  const CUtensorMapSwizzle sw = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  {
    cuuint64_t dims[4] = {(cuuint64_t)in_C, (cuuint64_t)in_W, (cuuint64_t)in_H, (cuuint64_t)L.B};
    cuuint64_t strides[3] = {(cuuint64_t)in_C * 2, (cuuint64_t)in_W * in_C * 2, (cuuint64_t)in_H * in_W * in_C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(L.tw * L.stride), (cuuint32_t)(L.th * L.stride), 1};
    if (L.swap && L.xr) box[2] = (cuuint32_t)(L.pair ? L.th / 2 + 2 : L.th + 2);  // tap-reuse variant: one halo row above and below (CTA pairs: half a tile each)
    cuuint32_t estr[4] = {1, (cuuint32_t)L.stride, (cuuint32_t)L.stride, 1};
    CUresult r = enc(&L.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in_base), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(g_conv_err, sizeof(g_conv_err), "encode A map failed: %d (C=%d W=%d H=%d B=%d box %d,%d,%d s=%d)", (int)r,
               in_C, in_W, in_H, L.B, bk, L.tw, L.th, L.stride);
      return 2;
    }
  }
  if (!L.swap && L.n_tail_tiles > 0) {  // tail tiles: left-over rows of tail_imgs consecutive images
    cuuint64_t dims[4] = {(cuuint64_t)in_C, (cuuint64_t)in_W, (cuuint64_t)in_H, (cuuint64_t)L.B};
    cuuint64_t strides[3] = {(cuuint64_t)in_C * 2, (cuuint64_t)in_W * in_C * 2, (cuuint64_t)in_H * in_W * in_C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(L.tw * L.stride), (cuuint32_t)(L.tail_rows * L.stride), (cuuint32_t)L.tail_imgs};
    cuuint32_t estr[4] = {1, (cuuint32_t)L.stride, (cuuint32_t)L.stride, 1};
    CUresult r = enc(&L.tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in_base), dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(g_conv_err, sizeof(g_conv_err), "encode tail map failed: %d (box %d,%d,%d,%d)", (int)r, bk, L.tw, L.tail_rows, L.tail_imgs);
      return 2;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)n_pad};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    // swapped: the group's own rows only (cluster pairs: each CTA fetches half of them)
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)(L.swap ? (L.cluster > 1 ? L.gw / L.cluster : L.gw) : L.block_n)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&L.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(g_conv_err, sizeof(g_conv_err), "encode B map failed: %d (K=%d N=%d box %d,%d)", (int)r, k_total, n_pad,
               bk, L.block_n);
      return 3;
    }
  }
  return 0;
}

int conv_make_io_maps(ConvLaunch& L, void* out_base, const void* res_base) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? out_base : res_base;
    if (!base) continue;
    const int C = which == 0 ? L.out_cstride : L.res_cstride;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)L.Wo, (cuuint64_t)L.Ho, (cuuint64_t)L.B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)L.Wo * C * 2, (cuuint64_t)L.Ho * L.Wo * C * 2};
    const bool f32 = which == 0 && L.out_fp32;
    const cuuint64_t es = f32 ? 4 : 2;
    strides[0] = (cuuint64_t)C * es; strides[1] = (cuuint64_t)L.Wo * C * es; strides[2] = (cuuint64_t)L.Ho * L.Wo * C * es;
    cuuint32_t box[4] = {(cuuint32_t)L.gw, (cuuint32_t)L.tw, (cuuint32_t)L.th, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(which == 0 ? &L.tmOut : &L.tmRes, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(g_conv_err, sizeof(g_conv_err), "encode io map %d failed: %d (C=%d W=%d H=%d box %d,%d,%d)", which, (int)r, C,
               L.Wo, L.Ho, L.n_total, L.tw, L.th);
      return 2;
    }
  }
  return 0;
}

template <int BK>
static int launch_t(const ConvLaunch& L, cudaStream_t stream) {
  static SmemOptIn opt_in;
  const size_t smem = conv_smem_bytes(L, BK);
  {
    cudaError_t e = ensure_dynamic_smem(conv_igemm_kernel<BK>, opt_in, smem);
    if (e != cudaSuccess) {
      snprintf(g_conv_err, sizeof(g_conv_err), "set smem %zu failed: %s", smem, cudaGetErrorString(e));
      return 4;
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(L.num_items < num_sms() ? L.num_items : num_sms());
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = conv_pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_igemm_kernel<BK>, L);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_conv_err, sizeof(g_conv_err), "conv launch failed: %s", cudaGetErrorString(e));
    return 5;
  }
  return 0;
}

int conv_launch(const ConvLaunch& L, int bk, cudaStream_t stream) {
  if (L.swap) return conv_swap_launch(L, bk, num_sms(), stream, g_conv_err, sizeof(g_conv_err));
  if (bk == 64) return launch_t<64>(L, stream);
  if (bk == 32) return launch_t<32>(L, stream);
  snprintf(g_conv_err, sizeof(g_conv_err), "unsupported BK %d", bk);
  return 6;
}

}  // namespace vgh
