// "Swapped" implicit-GEMM conv: output channels on the MMA M side, in groups of <= 128 (sm_100a).
//
// With Cout <= 128 the natural orientation (pixels = M, Cout = N <= 128) makes every tcgen05.mma
// read as many shared-memory bytes as an N=256 MMA for half the math, and the kernel saturates the
// 128 B/clk shared-memory port at ~50 % of tensor peak (measured: profiles/r1_ops_*).  Here the
// operands swap roles:  D^T[cout, pixel] = W[cout, K] * A[pixel, K]^T, i.e. M = 128 weight rows
// (zero-padded), N = a 256-pixel (tw x th) tile.  Same smem bytes per MMA as the N=256 case, twice
// the math.  The accumulator then has channels on TMEM lanes and pixels on columns; the epilogue
// transposes through a shared-memory tile [pixel][channel] (conflict-free 2-byte stores: a warp
// covers 32 consecutive channels of one pixel) which one thread hands to the TMA unit as a single
// 4-D box store into the consumer's channel slice (image-border clipping is done by the TMA unit).
// A residual tile, when present, is TMA-loaded into the same staging tile ahead of time and added
// in place.
//
// Layers with more than 128 output channels run as G equal channel groups (work item = pixel tile x
// group; the groups of one tile run on neighbouring CTAs at the same time, so the pixel tile is read
// from HBM once).  fp32 outputs (raw head tensors) use the same path with a 4-byte staging tile.
//
// XR variant (3x3 stride-1 layers, `p.xr`): the nine taps of a 3x3 filter read nine overlapping pixel
// tiles; fetched tap by tap that is 9x the unique L2->SM traffic of the tile, and with Cout <= 128 per
// group that traffic - not the tensor pipe - bounds the layer (~6-7 TB/s of distinct pixel bytes,
// profiles/r1_final_ncu_*).  Here the pixel tile is fetched ONCE per column shift dx with a one-row
// halo above and below ([BK, tw, th+2] box, OOB rows/columns zero-filled = padding) and the three
// row taps dy = 0,1,2 are three MMAs whose pixel-operand descriptor starts dy*tw rows further down
// the same shared-memory tile.  With tw a multiple of 8 that offset is a whole number of 8-row swizzle
// atoms, so the descriptors stay canonical.  3*(th+2) tile rows instead of 9*th: 2.4-2.7x less pixel
// traffic.  Weights keep streaming tap by tap through their own, deeper ring; the pixel ring has 2-3
// slots.  K order: (cin block, dx, dy).
//
// Cluster variant (`p.cluster == 2`, off by default): after the tap reuse the weight k-blocks - the same
// for every CTA - are more than half of the L2->SM bytes of a layer, and every conv launch plateaus at
// 7-9 TB/s of lts__t_bytes (profiles/r1_xr_ncu_launches_metrics.csv).  Two CTAs of a cluster work on two
// pixel tiles of the SAME channel group in lock step; each fetches one half of every weight k-block and
// TMA-multicasts it into both CTAs' rings, so the pair reads every weight byte from L2 once.  A slot is
// released to both producers by a multicast tcgen05.commit from each consumer (empty barriers count 2).
// Pixel tiles, TMEM, epilogue stay per CTA.  Measured neutral (profiles/r1_xr_cluster_ab.txt): the bound
// is the bytes each SM ingests, which a multicast does not reduce.  Kept, parity-tested, as the cluster
// plumbing (work-item mapping per pair, cluster barriers, multicast commits) cta_group::2 MMA pairs need.
//
// Pair variant (`p.pair`, template PAIR; 3x3 stride-1 layers with an even number of 128-channel groups): two CTAs of a
// cluster run ONE tcgen05.mma.cta_group::2 with M = 256 - CTA r holds the 128 weight rows of channel group 2g + r and gets
// their accumulator rows in its own TMEM - over ONE pixel tile whose N-side operand is split between them: CTA r stages
// only the upper / lower half of the tile (th/2 + 2 rows with halo) and the hardware reads N/2 rows from each CTA's
// shared memory.  Per SM that halves the pixel-tile fills and the MMA's pixel-operand reads: 160 -> 119 B/clk through the
// shared-memory port for a 256-channel layer, which is what bounded these layers at 78-80 % tensor-active.
// Both CTAs issue their own TMA loads but credit the byte counts to the LEADER's full barriers; the leader issues the
// MMAs and releases slots / publishes accumulators with multicast commits to both CTAs; the peer's epilogue warps
// hand the accumulator stage back by arriving on the leader's barrier through the cluster address space.
//
// Everything else matches conv_igemm.cu: persistent CTAs, TMA (4-D pixel box with OOB zero fill =
// padding, traversal stride = conv stride; 2-D weight box), mbarrier ring, double-buffered TMEM.
#include <cstdio>

#include "conv_igemm.cuh"
#include "device_attr.cuh"
#include "ptx.cuh"

namespace vgh {

constexpr int kSwapEpiWarps = 8;
constexpr int kSwapThreads = 64 + 32 * kSwapEpiWarps;

// epilogue: 16 pixels of one channel (one TMEM lane) -> + bias, ReLU, + alpha * residual (already in the staging tile) ->
// 16-bit activation (bf16 or fp16, ConvLaunch::f16) into the [pixel][channel] staging tile
template <bool F16>
__device__ __forceinline__ void stage_column16(const uint32_t (&v)[16], unsigned short* sp, int gw, float bias, bool relu, bool has_res,
                                               float alpha) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float x = __uint_as_float(v[j]) + bias;
    if (relu) x = fmaxf(x, 0.f);
    if (has_res) x = fmaf(alpha, unpack16<F16>(sp[j * gw]), x);
    sp[j * gw] = pack16<F16>(x);
  }
}

template <int BK, bool XR, bool PAIR = false>
__global__ void __launch_bounds__(kSwapThreads, 1) conv_igemm_swap_kernel(const __grid_constant__ ConvLaunch p) {
  static_assert(!PAIR || XR, "CTA pairs are built for the tap-reuse variant");
  constexpr uint32_t kRowBytes = BK * 2;
  constexpr uint32_t kWBytes = 128 * kRowBytes;  // weight slot: 128 cout rows (the MMA M)
  // only the group's own gw rows are fetched; the MMA still reads 128 rows, the rest is whatever the slot held -
  // those accumulator lanes are never stored
  const uint32_t w_tx = static_cast<uint32_t>(p.gw) * kRowBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int npix = p.tw * p.th;                                   // UMMA N (multiple of 16, <= 256)
  // pixel tile bytes; XR: the tile carries one halo row above and below
  // (PAIR: this CTA's half of the tile, th/2 rows + halo)
  const uint32_t x_bytes = static_cast<uint32_t>(PAIR ? p.tw * (p.th / 2 + 2) : XR ? p.tw * (p.th + 2) : npix) * kRowBytes;
  const uint32_t x_slot = (x_bytes + 1023u) & ~1023u;
  const int ks = p.ks;                                              // k-blocks per pipeline stage (one barrier round trip); XR: 1, or 3 = the three row taps of a column shift share a stage
  const int x_slots = XR ? p.xslots : p.stages * ks;               // XR: the pixel tiles have their own (shallower) ring
  const uint32_t w_base = smem_base;
  const uint32_t x_base = w_base + p.stages * ks * kWBytes;
  const uint32_t stage_base = x_base + x_slots * x_slot;                   // epilogue tile [npix][n_total] bf16
  const int gw = p.gw;                                                     // channels per group (<= 128)
  const uint32_t esz = p.out_fp32 ? 4u : 2u;
  const uint32_t stage_bytes = static_cast<uint32_t>(npix * gw) * esz;
  const uint32_t stage_slot = (stage_bytes + 1023u) & ~1023u;
  const int stg_bufs = p.stg_bufs;                                         // 1 or 2 epilogue tiles
  const uint32_t bar_base = stage_base + stg_bufs * stage_slot;
  const uint32_t full_bar = bar_base;
  const uint32_t empty_bar = bar_base + 8 * p.stages;
  const uint32_t tmem_full_bar = bar_base + 16 * p.stages;
  const uint32_t tmem_empty_bar = tmem_full_bar + 16;
  const uint32_t res_full_bar = tmem_empty_bar + 16;  // [2]
  const uint32_t xfull_bar = res_full_bar + 16;       // [8]  XR: pixel-tile ring
  const uint32_t xempty_bar = xfull_bar + 64;         // [8]
  const uint32_t tmem_slot = xempty_bar + 64;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Work items: a cluster (1 or 2 CTAs) walks (tile group of n_cl pixel tiles, channel group); CTA r of the
  // cluster takes tile r of the group.  With n_cl == 1 this is item = tile * ngroups + group as before.
  // The counts are set after the PDL wait: survivor-patch launches read their live height from device memory.
  const int n_cl = PAIR ? 1 : p.cluster;   // weight-multicast clusters (PAIR uses the cluster for the MMA instead)
  constexpr int kCtas = PAIR ? 2 : 1;
  const uint32_t cta_rank = (n_cl > 1 || PAIR) ? cluster_ctarank() : 0u;
  const bool leader = !PAIR || cta_rank == 0u;
  const int cl_id = blockIdx.x / (PAIR ? 2 : n_cl), n_cls = gridDim.x / (PAIR ? 2 : n_cl);
  int tiles_per_img = p.tiles_x * p.tiles_y;
  int total_tiles = tiles_per_img * p.B;
  const int item_groups = PAIR ? p.ngroups / 2 : p.ngroups;   // PAIR: one item = pixel tile x PAIR of channel groups
  int n_citems = ((total_tiles + n_cl - 1) / n_cl) * item_groups;
  // -> tile (clamped; `valid` false for the filler tile of an odd tile count: computed, never stored), group base
  auto item_tile = [&](int ci, int& tile, int& n_base) -> bool {
    const int tg = ci / item_groups;
    const int g = ci - tg * item_groups;
    n_base = (PAIR ? 2 * g + static_cast<int>(cta_rank) : g) * p.gw;
    tile = PAIR ? tg : tg * n_cl + static_cast<int>(cta_rank);
    const bool valid = tile < total_tiles;
    if (!valid) tile = total_tiles - 1;
    return valid;
  };
  const uint16_t mc_mask = static_cast<uint16_t>((1u << n_cl) - 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, n_cl);  // cluster: a weight slot is free once BOTH consumers have released it
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar + 8 * a, 1);
      mbar_init(tmem_empty_bar + 8 * a, kSwapEpiWarps * kCtas);  // PAIR: the leader's barrier also collects the peer's epilogue warps
    }
    mbar_init(res_full_bar, 1);
    mbar_init(res_full_bar + 8, 1);
    if (XR) {
      for (int s = 0; s < p.xslots; ++s) {
        mbar_init(xfull_bar + 8 * s, 1);
        mbar_init(xempty_bar + 8 * s, 1);
      }
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc_dyn_2cta(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish_2cta();
    } else {
      tmem_alloc_dyn(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (n_cl > 1 || PAIR) cluster_sync_all();  // the peer's barriers exist before anything is multicast to / arrives on them
  uint32_t tmem_acc;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_acc) : "r"(tmem_slot));
  // everything above overlapped the previous kernel's tail (PDL); its outputs are needed from here on
  pdl_wait();
  pdl_launch_dependents();
  if (p.dyn_rows != nullptr) {  // stacked survivor patches (one tall image): only the first *dyn_rows rows are live
    const int rows = min(__ldg(p.dyn_rows), p.Ho);
    tiles_per_img = p.tiles_x * ((rows + p.th - 1) / p.th);
    total_tiles = tiles_per_img * p.B;
    n_citems = ((total_tiles + n_cl - 1) / n_cl) * item_groups;
  }

  const int cblks = p.cin / BK;
  const int num_kb = p.ntaps * cblks;
  const uint32_t acc_cols = static_cast<uint32_t>(npix);
  // weight k-block fetch: the whole group (n_cl == 1) or this CTA's half of it, multicast to the pair
  const uint32_t w_half = (static_cast<uint32_t>(p.gw) / 2u) * kRowBytes;
  auto load_w = [&](uint32_t slot_addr, uint32_t bar, int k0, int n_base) {
    if (n_cl > 1)
      tma_load_2d_mc(slot_addr + cta_rank * w_half, &p.tmB, bar, k0, n_base + static_cast<int>(cta_rank) * (p.gw / 2), mc_mask);
    else
      tma_load_2d(slot_addr, &p.tmB, bar, k0, n_base);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      const uint32_t tx_bytes = w_tx + x_bytes;
      int stage = 0, xs = 0;
      uint32_t phase = 0, xphase = 0;
      for (int item = cl_id; item < n_citems; item += n_cls) {
        int tile, n_base;
        item_tile(item, tile, n_base);
        const int b_img = tile / tiles_per_img;
        const int t_in = tile - b_img * tiles_per_img;
        const int tyi = t_in / p.tiles_x;
        const int h0 = tyi * p.th, w0 = (t_in - tyi * p.tiles_x) * p.tw;
        if constexpr (PAIR) {
          // both CTAs fetch their own operand halves; the LEADER's full barriers count the bytes of both
          const int hh = h0 + static_cast<int>(cta_rank) * (p.th / 2) - 1;
          for (int cb = 0; cb < cblks; ++cb) {
            for (int dx = 0; dx < 3; ++dx) {
              mbar_wait(xempty_bar + 8 * xs, xphase ^ 1);
              if (leader) mbar_arrive_expect_tx(xfull_bar + 8 * xs, 2 * x_bytes);
              tma_load_4d_2cta(x_base + xs * x_slot, &p.tmA, map_to_cta(xfull_bar + 8 * xs, 0), p.cin_off + cb * BK, w0 + dx - 1, hh, b_img);
              if (++xs == p.xslots) {
                xs = 0;
                xphase ^= 1;
              }
              for (int dy = 0; dy < 3; ++dy) {
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                if (leader) mbar_arrive_expect_tx(full_bar + 8 * stage, 2 * w_tx);
                tma_load_2d_2cta(w_base + stage * kWBytes, &p.tmB, map_to_cta(full_bar + 8 * stage, 0), (dy * 3 + dx) * p.cin + cb * BK, n_base);
                if (++stage == p.stages) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
        } else if constexpr (XR) {
          // pixel tile (with halo rows) once per (cin block, dx); weights tap by tap
          for (int cb = 0; cb < cblks; ++cb) {
            for (int dx = 0; dx < 3; ++dx) {
              mbar_wait(xempty_bar + 8 * xs, xphase ^ 1);
              mbar_arrive_expect_tx(xfull_bar + 8 * xs, x_bytes);
              tma_load_4d(x_base + xs * x_slot, &p.tmA, xfull_bar + 8 * xs, p.cin_off + cb * BK, w0 + dx - 1, h0 - 1, b_img);
              if (++xs == p.xslots) {
                xs = 0;
                xphase ^= 1;
              }
              if (ks == 3) {   // one stage = the weight k-blocks of the three row taps: a third of the barrier round trips
                mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                mbar_arrive_expect_tx(full_bar + 8 * stage, 3 * w_tx);
                for (int dy = 0; dy < 3; ++dy)
                  load_w(w_base + (stage * 3 + dy) * kWBytes, full_bar + 8 * stage, (dy * 3 + dx) * p.cin + cb * BK, n_base);
                if (++stage == p.stages) {
                  stage = 0;
                  phase ^= 1;
                }
              } else {
                for (int dy = 0; dy < 3; ++dy) {
                  mbar_wait(empty_bar + 8 * stage, phase ^ 1);
                  mbar_arrive_expect_tx(full_bar + 8 * stage, w_tx);
                  load_w(w_base + stage * kWBytes, full_bar + 8 * stage, (dy * 3 + dx) * p.cin + cb * BK, n_base);
                  if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                  }
                }
              }
            }
          }
        } else {
          int tap = 0, cb = 0;
          for (int kb = 0; kb < num_kb; kb += ks) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1);
            const uint32_t fb = full_bar + 8 * stage;
            mbar_arrive_expect_tx(fb, tx_bytes * ks);
            for (int q = 0; q < ks; ++q) {
              const int ty = tap / p.kw;
              const int tx = tap - ty * p.kw;
              tma_load_4d(x_base + (stage * ks + q) * x_slot, &p.tmA, fb, p.cin_off + cb * BK, w0 * p.stride + tx - p.pad,
                          h0 * p.stride + ty - p.pad, b_img);
              load_w(w_base + (stage * ks + q) * kWBytes, fb, tap * p.cin + cb * BK, n_base);
              if (++cb == cblks) {
                cb = 0;
                ++tap;
              }
            }
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ===================== MMA issuer (PAIR: the leader CTA, for both) =====================
      const uint32_t idesc = umma_idesc_16(PAIR ? 256 : 128, static_cast<uint32_t>(npix), p.f16 != 0);
      int stage = 0, xs = 0;
      uint32_t phase = 0, xphase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = cl_id; item < n_citems; item += n_cls) {
        mbar_wait(tmem_empty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_acc + acc * acc_cols;
        if constexpr (XR) {
          uint32_t accumulate = 0;
          const uint32_t dy_bytes = static_cast<uint32_t>(p.tw) * kRowBytes;  // one tile row; tw % 8 == 0: whole swizzle atoms
          for (int g = 0; g < 3 * cblks; ++g) {
            mbar_wait(xfull_bar + 8 * xs, xphase);
            tc_fence_after();
            if (!PAIR && ks == 3) {   // the three row taps share one weight stage
              mbar_wait(full_bar + 8 * stage, phase);
              tc_fence_after();
              for (int dy = 0; dy < 3; ++dy) {
                const uint64_t w_desc = umma_smem_desc(w_base + (stage * 3 + dy) * kWBytes, kRowBytes);
                const uint64_t x_desc = umma_smem_desc(x_base + xs * x_slot + dy * dy_bytes, kRowBytes);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  umma_bf16(d, w_desc + 2 * k, x_desc + 2 * k, idesc, accumulate);
                  accumulate = 1u;
                }
              }
              if (n_cl > 1) umma_commit_mc(empty_bar + 8 * stage, mc_mask);
              else umma_commit(empty_bar + 8 * stage);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            } else {
            for (int dy = 0; dy < 3; ++dy) {
              mbar_wait(full_bar + 8 * stage, phase);
              tc_fence_after();
              const uint64_t w_desc = umma_smem_desc(w_base + stage * kWBytes, kRowBytes);
              const uint64_t x_desc = umma_smem_desc(x_base + xs * x_slot + dy * dy_bytes, kRowBytes);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                if (PAIR) umma_bf16_2cta(d, w_desc + 2 * k, x_desc + 2 * k, idesc, accumulate);
                else umma_bf16(d, w_desc + 2 * k, x_desc + 2 * k, idesc, accumulate);
                accumulate = 1u;
              }
              if (PAIR) umma_commit_2cta_mc(empty_bar + 8 * stage, 3);
              else if (n_cl > 1) umma_commit_mc(empty_bar + 8 * stage, mc_mask);
              else umma_commit(empty_bar + 8 * stage);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
            }
            // all three row taps of this pixel tile have been issued
            if (PAIR) umma_commit_2cta_mc(xempty_bar + 8 * xs, 3); else umma_commit(xempty_bar + 8 * xs);
            if (++xs == p.xslots) {
              xs = 0;
              xphase ^= 1;
            }
          }
        } else {
          for (int kb = 0; kb < num_kb; kb += ks) {
            mbar_wait(full_bar + 8 * stage, phase);
            tc_fence_after();
            for (int q = 0; q < ks; ++q) {
              const uint64_t w_desc = umma_smem_desc(w_base + (stage * ks + q) * kWBytes, kRowBytes);  // M side: weights
              const uint64_t x_desc = umma_smem_desc(x_base + (stage * ks + q) * x_slot, kRowBytes);   // N side: pixels
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16(d, w_desc + 2 * k, x_desc + 2 * k, idesc, (kb | q | k) != 0 ? 1u : 0u);
            }
            if (n_cl > 1) umma_commit_mc(empty_bar + 8 * stage, mc_mask); else umma_commit(empty_bar + 8 * stage);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (PAIR) umma_commit_2cta_mc(tmem_full_bar + 8 * acc, 3); else umma_commit(tmem_full_bar + 8 * acc);
        if (++acc == p.acc_stages) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue: lanes = channels, columns = pixels =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int chunk0 = ew >> 2;
    const int ch = quarter * 32 + lane;  // output channel of this thread inside its group
    const bool ch_ok = ch < gw;
    const int n_chunks = npix >> 4;
    const bool has_res = p.res != nullptr;
    const bool epi_lead = (ew == 0 && lane == 0);
    uint8_t* stile0 = smem_raw + (stage_base - smem_u32(smem_raw));
    // cluster work item -> tile origin, channel-group base; false for a filler tile (computed, not stored)
    auto item_coords = [&](int item, int& b_img, int& h0, int& w0, int& n_base) -> bool {
      int tile;
      const bool valid = item_tile(item, tile, n_base);
      b_img = tile / tiles_per_img;
      const int t_in = tile - b_img * tiles_per_img;
      const int tyi = t_in / p.tiles_x;
      h0 = tyi * p.th;
      w0 = (t_in - tyi * p.tiles_x) * p.tw;
      return valid;
    };
    if (has_res && epi_lead && cl_id < n_citems) {  // residual tile of the first item -> buffer 0
      int b_img, h0, w0, nb0;
      item_coords(cl_id, b_img, h0, w0, nb0);
      tma_prefetch_desc(&p.tmRes);
      mbar_arrive_expect_tx(res_full_bar, stage_bytes);
      tma_load_4d(stage_base, &p.tmRes, res_full_bar, p.res_coff + nb0, w0, h0, b_img);
    }
    if (epi_lead) tma_prefetch_desc(&p.tmOut);
    int acc = 0, sb = 0;
    uint32_t acc_phase = 0, res_phase[2] = {0, 0};
    for (int item = cl_id; item < n_citems; item += n_cls) {
      int b_img, h0, w0, n_base;
      const bool valid = item_coords(item, b_img, h0, w0, n_base);
      const float bias = ch_ok ? __ldg(p.bias + n_base + ch) : 0.f;
      // two staging tiles: the residual tile of the NEXT item is fetched into the other tile while this item is still being
      // multiplied (its previous user, the store of item i-1, only has to have been read out)
      if (has_res && stg_bufs == 2 && epi_lead) {
        const int next = item + n_cls;
        if (next < n_citems) {
          tma_store_wait_read();
          int nb, nh, nw, nn;
          item_coords(next, nb, nh, nw, nn);
          mbar_arrive_expect_tx(res_full_bar + 8 * (sb ^ 1), stage_bytes);
          tma_load_4d(stage_base + (sb ^ 1) * stage_slot, &p.tmRes, res_full_bar + 8 * (sb ^ 1), p.res_coff + nn, nw, nh, nb);
        }
      }
      mbar_wait(tmem_full_bar + 8 * acc, acc_phase);
      tc_fence_after();
      if (has_res) {
        mbar_wait(res_full_bar + 8 * sb, res_phase[sb]);
        res_phase[sb] ^= 1;
      }
      uint8_t* stile = stile0 + sb * stage_slot;
      const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quarter * 32) << 16) + acc * acc_cols;
      for (int c = chunk0; c < n_chunks; c += 2) {
        uint32_t v[16];
        tmem_ld16(taddr + c * 16, v);
        tmem_ld_wait();
        if (ch_ok) {
          if (p.out_fp32) {
            float* sp = reinterpret_cast<float*>(stile) + (c * 16) * gw + ch;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = __uint_as_float(v[j]) + bias;
              if (p.relu) x = fmaxf(x, 0.f);
              sp[j * gw] = x;
            }
          } else {
            unsigned short* sp = reinterpret_cast<unsigned short*>(stile) + (c * 16) * gw + ch;
            if (p.f16) stage_column16<true>(v, sp, gw, bias, p.relu != 0, has_res, p.res_alpha);
            else stage_column16<false>(v, sp, gw, bias, p.relu != 0, has_res, p.res_alpha);
          }
        }
      }
      // TMEM drained by this warp -> MMA may reuse the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && !leader) mbar_arrive_cluster(map_to_cta(tmem_empty_bar + 8 * acc, 0));   // the leader's MMA warp waits for both CTAs
        else mbar_arrive(tmem_empty_bar + 8 * acc);
      }
      // hand the finished tile to the TMA unit
      fence_proxy_async();
      named_bar_sync(1, 32 * kSwapEpiWarps);
      const int nsb = stg_bufs == 2 ? (sb ^ 1) : sb;
      if (epi_lead) {
        if (valid) tma_store_4d(&p.tmOut, stage_base + sb * stage_slot, p.out_coff + n_base, w0, h0, b_img);
        tma_store_commit();
        // the tile the NEXT item writes must have been read out by its previous store
        if (stg_bufs == 2) tma_store_wait_read_keep1(); else tma_store_wait_read();
        const int next = item + n_cls;
        if (has_res && stg_bufs == 1 && next < n_citems) {
          int nb, nh, nw, nn;
          item_coords(next, nb, nh, nw, nn);
          mbar_arrive_expect_tx(res_full_bar + 8 * nsb, stage_bytes);
          tma_load_4d(stage_base + nsb * stage_slot, &p.tmRes, res_full_bar + 8 * nsb, p.res_coff + nn, nw, nh, nb);
        }
      }
      named_bar_sync(1, 32 * kSwapEpiWarps);
      sb = nsb;
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (epi_lead) tma_store_wait_read();  // shared memory must outlive the last store's read
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs are done with the pair's MMAs / TMEM / barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_dyn_2cta(tmem_acc, static_cast<uint32_t>(p.tmem_cols));
    else tmem_dealloc_dyn(tmem_acc, static_cast<uint32_t>(p.tmem_cols));
  }
  // the peer's last slot releases land on this CTA's barriers: shared memory must outlive them
  if (n_cl > 1) cluster_sync_all();
}

size_t conv_swap_smem_bytes(const ConvLaunch& L, int bk) {
  const size_t staging = (static_cast<size_t>(L.tw * L.th) * L.gw * (L.out_fp32 ? 4 : 2) + 1023) & ~static_cast<size_t>(1023);
  if (L.xr) {
    const size_t x_slot = (static_cast<size_t>(L.tw * (L.pair ? L.th / 2 + 2 : L.th + 2)) * bk * 2 + 1023) & ~static_cast<size_t>(1023);
    return 1024 + static_cast<size_t>(L.stages) * L.ks * 128 * bk * 2 + L.xslots * x_slot + L.stg_bufs * staging + 16 * L.stages + 256;
  }
  const size_t x_slot = (static_cast<size_t>(L.tw * L.th) * bk * 2 + 1023) & ~static_cast<size_t>(1023);
  return 1024 + static_cast<size_t>(L.stages) * L.ks * (128 * bk * 2 + x_slot) + L.stg_bufs * staging + 16 * L.stages + 256;
}

template <int BK, bool XR, bool PAIR = false>
static int launch_swap_t(const ConvLaunch& L, int sms, cudaStream_t stream, char* err, size_t errlen) {
  static SmemOptIn opt_in;
  const size_t smem = conv_swap_smem_bytes(L, BK);
  {
    cudaError_t e = ensure_dynamic_smem(conv_igemm_swap_kernel<BK, XR, PAIR>, opt_in, smem);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "swap conv: set smem %zu failed: %s", smem, cudaGetErrorString(e));
      return 4;
    }
  }
  cudaLaunchConfig_t cfg{};
  const int n_cl = PAIR ? 2 : (L.cluster > 1 ? L.cluster : 1);
  int grid = L.num_items < sms ? L.num_items : sms;  // num_items counts CTA-level items (cluster items x cluster size)
  grid -= grid % n_cl;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kSwapThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (n_cl > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = n_cl;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (conv_pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_igemm_swap_kernel<BK, XR, PAIR>, L);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "swap conv launch failed: %s", cudaGetErrorString(e));
    return 5;
  }
  return 0;
}

int conv_swap_launch(const ConvLaunch& L, int bk, int sms, cudaStream_t stream, char* err, size_t errlen) {
  if (L.cluster != 1 && (L.cluster != 2 || L.gw % 16 || sms < 2)) {
    snprintf(err, errlen, "swap conv: bad cluster configuration (cluster %d, group width %d)", L.cluster, L.gw);
    return 7;
  }
  if (L.xr) {
    if (L.ntaps != 9 || L.stride != 1 || L.tw % 8 || L.xslots < 2 || L.xslots > 8 || L.stages < 2 || (L.ks != 1 && (L.ks != 3 || L.pair))) {
      snprintf(err, errlen, "swap conv: launch not eligible for the tap-reuse variant (taps %d stride %d tile %dx%d xslots %d stages %d)",
               L.ntaps, L.stride, L.tw, L.th, L.xslots, L.stages);
      return 7;
    }
    if (L.pair) {
      if (bk != 64 || L.gw != 128 || L.ngroups % 2 || L.th % 2 || L.cluster != 1 || L.dyn_rows != nullptr || (L.tw * L.th) % 16) {
        snprintf(err, errlen, "swap conv: launch not eligible for CTA pairs (bk %d, group width %d, groups %d, tile %dx%d)", bk, L.gw, L.ngroups, L.tw, L.th);
        return 7;
      }
      return launch_swap_t<64, true, true>(L, sms, stream, err, errlen);
    }
    if (bk == 64) return launch_swap_t<64, true>(L, sms, stream, err, errlen);
    if (bk == 32) return launch_swap_t<32, true>(L, sms, stream, err, errlen);
  }
  if (bk == 64) return launch_swap_t<64, false>(L, sms, stream, err, errlen);
  if (bk == 32) return launch_swap_t<32, false>(L, sms, stream, err, errlen);
  snprintf(err, errlen, "unsupported BK %d", bk);
  return 6;
}

}  // namespace vgh
