// C-ABI of libvggheads_b200.so (see include/vggheads_b200.h): plan executor for the conv network,
// CUDA-graph capture of the whole device-side pipeline, and the stand-alone NMS / FLAME entry points.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/vggheads_b200.h"
#include "aux_kernels.cuh"
#include "conv_igemm.cuh"
#include "flame_decode.cuh"
#include "gather.cuh"
#include "letterbox.cuh"
#include "mesh_kernels.cuh"
#include "select_nms.cuh"

using namespace vgh;

static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_OK(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) return fail(100, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

extern "C" int vgh_version(void) { return 1; }
extern "C" const char* vgh_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------ FLAME
struct vgh_flame {
  FlameModel* model;
};

extern "C" int vgh_flame_create(const float* v_template, const float* shapedirs, const float* posedirs,
                                const float* j_regressor, const float* lbs_weights, vgh_flame** out) {
  if (!v_template || !shapedirs || !posedirs || !j_regressor || !lbs_weights || !out) return fail(1, "null argument");
  FlameModel* m = nullptr;
  int rc = flame_model_create(v_template, shapedirs, posedirs, j_regressor, lbs_weights, &m, g_err, sizeof(g_err));
  if (rc) return rc;
  *out = new vgh_flame{m};
  return 0;
}
extern "C" void vgh_flame_destroy(vgh_flame* f) {
  if (!f) return;
  flame_model_destroy(f->model);
  delete f;
}
extern "C" int vgh_flame_decode(const vgh_flame* f, const float* params_dev, int n, int n_shape_live, int n_expr_live,
                                const float* xform_dev, float* verts_dev, float* rot_dev, float* proj_dev,
                                void* stream) {
  if (!f) return fail(1, "null flame handle");
  if (n < 0) return fail(1, "negative head count");
  if (n == 0) return 0;
  if (!params_dev || !proj_dev) return fail(1, "null params/proj pointer");
  return flame_decode_launch(f->model, params_dev, n, nullptr, n_shape_live, n_expr_live, xform_dev, verts_dev, rot_dev,
                             proj_dev, static_cast<cudaStream_t>(stream), g_err, sizeof(g_err));
}

// ------------------------------------------------------------------------------------------ NMS
extern "C" int vgh_select_nms(const float* boxes_dev, const float* scores_dev, int B, int A, float conf_thr,
                              float iou_thr, int top_k, int keep_k, int32_t* keep_idx_dev, int32_t* keep_cnt_dev,
                              float* keep_boxes_dev, float* keep_scores_dev, void* stream) {
  if (!boxes_dev || !scores_dev || !keep_idx_dev || !keep_cnt_dev) return fail(1, "null argument");
  if (B < 0 || A <= 0) return fail(1, "bad batch/anchor count");
  return select_nms_launch(boxes_dev, scores_dev, B, A, conf_thr, iou_thr, top_k, keep_k, keep_idx_dev, keep_cnt_dev,
                           keep_boxes_dev, keep_scores_dev, static_cast<cudaStream_t>(stream), g_err, sizeof(g_err));
}

// ------------------------------------------------------------------------------------------ letterbox
extern "C" int vgh_letterbox(const uint8_t* src_dev, const int64_t* offsets, const int32_t* heights, const int32_t* widths,
                             int n, int image_size, uint8_t* out_dev, float* xform_host, void* stream) {
  if (n < 0) return fail(1, "negative image count");
  if (n == 0) return 0;
  if (!src_dev || !offsets || !heights || !widths || !out_dev) return fail(1, "null argument");
  if (image_size < 1) return fail(1, "bad image size");
  return letterbox_launch(src_dev, offsets, heights, widths, n, image_size, out_dev, xform_host,
                          static_cast<cudaStream_t>(stream), g_err, sizeof(g_err));
}

// ------------------------------------------------------------------------------------------ mesh consumers
extern "C" int vgh_pncc_render(const float* verts_dev, int n, const int32_t* tris_dev, int ntri, const float* colors_dev, int H, int W,
                               uint8_t* image_dev, uint64_t* keys_dev, void* stream) {
  if (n < 0 || ntri < 0 || H < 1 || W < 1) return fail(1, "bad argument");
  if (n == 0 || ntri == 0) return 0;
  if (!verts_dev || !tris_dev || !colors_dev || !image_dev || !keys_dev) return fail(1, "null argument");
  const int rc = pncc_render_launch(verts_dev, n, VGH_NUM_VERTS, tris_dev, ntri, colors_dev, H, W, image_dev,
                                    reinterpret_cast<unsigned long long*>(keys_dev), static_cast<cudaStream_t>(stream));
  if (rc == 2) return fail(1, "pncc: at most 4094 heads and 2^20-1 triangles per call");
  return rc ? fail(5, "pncc launch failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
}
extern "C" int vgh_head_bbox(const float* verts_dev, int n, const int32_t* idx_dev, int n_idx, int32_t* out_xywh_dev, void* stream) {
  if (n < 0 || n_idx < 1) return fail(1, "bad argument");
  if (n == 0) return 0;
  if (!verts_dev || !idx_dev || !out_xywh_dev) return fail(1, "null argument");
  return head_bbox_launch(verts_dev, n, VGH_NUM_VERTS, idx_dev, n_idx, out_xywh_dev, static_cast<cudaStream_t>(stream))
             ? fail(5, "head bbox launch failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
}

// ------------------------------------------------------------------------------------------ detector
struct OpRt {
  vgh_op_desc d;
  ConvLaunch L;
  int bk;
  int cfg_mt = 0, cfg_stages = 0, cfg_tw = 0, cfg_th = 0;  // 0 = heuristic; set by vgh_detector_autotune
  int cfg_swap = -1;                                          // -1 = heuristic
  int cfg_ks = 0;                                             // swapped kernel: k-blocks per stage (0 = 1)
  int cfg_tail = 1;                                           // normal kernel: pack left-over rows across images
  int cfg_xr = -1, cfg_xslots = 0;                            // swapped kernel: tap-reuse variant (-1 = default), pixel ring depth
  int cfg_cluster = -1;                                       // swapped kernel: CTA pairs sharing the weight stream (-1 = default)
  int cfg_pair = -1;                                          // swapped tap-reuse kernel: cta_group::2 MMA pairs (-1 = default)
};

struct vgh_detector {
  int B = 0, S = 0, A = 0, keep_k = 100;
  std::vector<vgh_buf_desc> bufs;
  std::vector<void*> buf_ptr;
  std::vector<size_t> buf_bytes;
  std::vector<OpRt> ops;
  __nv_bfloat16* weights = nullptr;
  float* bias = nullptr;
  DecodeLevels lv;
  const vgh_flame* flame = nullptr;
  uint8_t* input = nullptr;
  float *boxes = nullptr, *scores = nullptr, *keep_boxes = nullptr, *keep_scores = nullptr;
  int *keep_idx = nullptr, *keep_cnt = nullptr, *offsets = nullptr, *head_img = nullptr;
  float *params = nullptr, *head_xform = nullptr, *verts = nullptr, *rot = nullptr, *img_xform = nullptr;
  const float *ovr_boxes = nullptr, *ovr_scores = nullptr;
  // sparse heads: ops [n_dense_ops, n) run after select/NMS on survivor patches (one stack per head level)
  int n_dense_ops = 0, patch_cap = 0;
  bool sparse = false;
  bool split = false;  // parity mode: activations as split bf16 terms, normal (pixels-on-M) conv kernel only
  bool f16 = false;    // 16-bit activations / weights are IEEE fp16 (vgh_net_desc.act_f16) instead of bf16
  int *head_level = nullptr, *head_patch = nullptr, *patch_src = nullptr, *level_rows = nullptr;
  cudaGraphExec_t graph = nullptr;
  cudaStream_t cap_stream = nullptr;  // capture needs a non-legacy stream; the graph then replays anywhere
  // two-deep host pipeline (vgh_detector_submit_host / collect_host)
  struct Slot {
    uint8_t* in_dev = nullptr;
    float *params = nullptr, *verts = nullptr, *kboxes = nullptr, *kscores = nullptr;
    int *cnt = nullptr, *total = nullptr;
    cudaEvent_t h2d_done = nullptr, in_consumed = nullptr, compute_done = nullptr, d2h_done = nullptr;
    bool busy = false;
  } slots[2];
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr, pipe_stream = nullptr;
  int submit_idx = 0, collect_idx = 0;  // host pipeline only: submit_host advances one, collect_host the other
  // multi-GPU record push (vgh_detector_arm_push): destination of the NEXT submit's packed result record
  struct Push {
    bool armed = false;
    float* dst = nullptr;
    const unsigned long long* wait_flag = nullptr;
    unsigned long long wait_val = 0;
    unsigned long long* done_flag = nullptr;
    unsigned long long done_val = 0;
  } push;
  int *push_counter = nullptr, *push_status = nullptr;  // device: block counter of the pack kernel / sticky time-out bits
  uint32_t push_seq = 0;
  std::vector<cudaStream_t> lane_streams;  // side streams for independent graph branches (lanes 1..)
  std::vector<cudaEvent_t> op_events;
  cudaEvent_t phase_event = nullptr;
  bool multi_lane = true;
  float g_conf = -1.f, g_iou = -1.f;
  int g_topk = -1;
  int launches = 0;
};

// number of 128-row tiles of a (tw x th) tiling of B images, with the left-over rows of several
// images packed into shared tiles when that is possible (see build_conv)
static long count_tiles(int Ho, int Wo, int B, int tw, int th, bool pack_tails) {
  const int tx = (Wo + tw - 1) / tw;
  const int tr = Ho % th;
  if (pack_tails && tr != 0 && 128 / (tw * tr) >= 2) {
    int imgs = 128 / (tw * tr);
    if (imgs > B) imgs = B;
    return static_cast<long>(tx) * (Ho / th) * B + static_cast<long>(tx) * ((B + imgs - 1) / imgs);
  }
  return static_cast<long>(tx) * ((Ho + th - 1) / th) * B;
}

static void pick_tile(int Ho, int Wo, int B, bool pack_tails, int& tw, int& th) {
  long best = -1;
  tw = 1; th = 1;
  for (int h = 1; h <= 128 && h <= Ho; ++h) {
    int w = 128 / h;
    if (w > Wo) w = Wo;
    if (w < 1) continue;
    const long n = count_tiles(Ho, Wo, B, w, h, pack_tails);
    // fewest tiles wins; prefer wide tiles on ties (longer contiguous runs per TMA box row)
    if (best < 0 || n < best || (n == best && w > tw)) { best = n; tw = w; th = h; }
  }
}

// swapped mode: tile of up to 256 pixels whose pixel count is a multiple of 16 (it is the UMMA N)
static void pick_tile_swap(int Ho, int Wo, int& tw, int& th, int max_px = 256) {
  double best = -1.0;
  tw = 16; th = 1;
  for (int h = 1; h <= max_px && h <= Ho; ++h) {
    for (int w = (max_px / h < Wo ? max_px / h : Wo); w >= 1; --w) {
      if ((w * h) % 16) continue;
      const int tx = (Wo + w - 1) / w, ty = (Ho + h - 1) / h;
      const double eff = static_cast<double>(Ho) * Wo / (static_cast<double>(tx) * ty * max_px);
      if (eff > best + 1e-9) { best = eff; tw = w; th = h; }
      break;  // widest admissible w for this h only
    }
  }
}

// tap-reuse variant of the swapped kernel (3x3 stride 1): tile width a multiple of 8 pixels so that a row
// of the tile is a whole number of 8-row swizzle atoms; the tile may overhang the map (TMA zero-fills the
// load and clips the store).  Fewest MMA columns wasted wins; ties go to the taller tile (smaller halo share).
static void pick_tile_swap_xr(int Ho, int Wo, int& tw, int& th, int max_px = 256) {
  double best = -1.0;
  tw = 8; th = 2;
  const int wo8 = (Wo + 7) / 8 * 8;
  for (int h = 1; h <= Ho && h <= 32 && h * 8 <= max_px; ++h) {
    int w = max_px / h / 8 * 8;
    if (w > wo8) w = wo8;
    if (w < 8 || (w * h) % 16) continue;
    const int tx = (Wo + w - 1) / w, ty = (Ho + h - 1) / h;
    const double eff = static_cast<double>(Ho) * Wo / (static_cast<double>(tx) * ty * max_px);
    if (eff > best - 1e-9) { best = eff > best ? eff : best; tw = w; th = h; }
  }
}

// swapped kernel: G equal output-channel groups of gw <= 128 channels; a group's TMA box must be a
// multiple of 16 bytes and the weight matrix must have 128 rows from every group start
static bool swap_groups(const vgh_op_desc& q, const vgh_buf_desc& ob, int& G, int& gw) {
  if (q.up) return false;
  G = (q.cout + 127) / 128;
  if (q.cout % G) return false;
  gw = q.cout / G;
  if ((gw * (ob.fp32 ? 4 : 2)) % 16) return false;
  return q.n_pad >= gw * (G - 1) + 128;
}
static bool swap_eligible(const vgh_op_desc& q, const vgh_buf_desc& ob) {
  int G, gw;
  if (!swap_groups(q, ob, G, gw)) return false;
  return !(q.res_buf >= 0 && ob.fp32);
}

static bool xr_eligible(const vgh_op_desc& q, const vgh_buf_desc& ob) {
  return q.ksize == 3 && q.stride == 1 && !q.up && swap_eligible(q, ob);
}
// default for eligible ops: VGGHEADS_B200_XR = 0 (never) | 1 (whenever the swapped kernel runs a 3x3 stride-1 layer)
static int xr_default() {
  const char* e = getenv("VGGHEADS_B200_XR");
  return (e && e[0] == '0') ? 0 : 1;
}
// CTA-pair weight multicast: VGGHEADS_B200_CLUSTER = 1 turns it on for every eligible swapped launch (and makes it an
// autotune candidate).  Off by default: measured neutral (same layer times to 0.5 %, profiles/r1_xr_cluster_ab.txt) -
// what bounds these layers is the bytes each SM has to ingest, not the reads at the L2 slices, and a multicast
// still delivers every byte to every SM.  Kept (and parity-tested) as the cluster plumbing cta_group::2 pairs need.
static int cluster_default() {
  const char* e = getenv("VGGHEADS_B200_CLUSTER");
  return (e && e[0] == '1') ? 1 : 0;
}
// cta_group::2 MMA pairs (conv_igemm_swap.cu, PAIR): VGGHEADS_B200_PAIR = 0 (never) | 1 (default: wherever eligible; the
// autotuner keeps whichever of pair / single measures faster per layer)
static int pair_default() {
  const char* e = getenv("VGGHEADS_B200_PAIR");
  return (e && e[0] == '0') ? 0 : 1;
}
// deeper TMA rings for the 32-channel K blocks of the tap-reuse kernel (up to 16 weight stages, a fourth pixel slot as an
// autotune candidate): VGGHEADS_B200_DEEP_RINGS = 0 restores the 8-stage cap (A/B aid)
static bool deep_rings() {
  const char* e = getenv("VGGHEADS_B200_DEEP_RINGS");
  return !(e && e[0] == '0');
}
// tap-reuse kernel, 32-channel K blocks: the three row taps of a column shift share one weight stage (default and
// autotune candidate); VGGHEADS_B200_WGROUP = 0: one tap per stage as for the 64-channel K blocks (A/B aid)
static bool wgroup_default() {
  const char* e = getenv("VGGHEADS_B200_WGROUP");
  return !(e && e[0] == '0');
}
// tap-reuse kernel, residual layers with 32-channel K blocks: second staging tile = residual prefetch one item ahead;
// VGGHEADS_B200_STG2 = 0 keeps one tile (A/B aid)
static bool stg2_default() {
  const char* e = getenv("VGGHEADS_B200_STG2");
  return !(e && e[0] == '0');
}
// test aid: VGGHEADS_B200_SWAP=1 makes the un-tuned heuristic pick the swapped kernel for every eligible op
// (by default only large maps with Cout <= 128 do), so that small parity cases exercise it everywhere
static bool swap_forced() {
  const char* e = getenv("VGGHEADS_B200_SWAP");
  return e && e[0] == '1';
}

static int auto_block_n(int cout, int up, int up_cout) {
  const int lim = up ? up_cout : cout;
  if (lim <= 256) return lim;
  for (int n = 256; n >= 16; n -= 16)
    if (lim % n == 0) return n;
  return 16;
}

static int build_conv(vgh_detector* d, OpRt& o) {
  const vgh_op_desc& q = o.d;
  if (q.in_buf < 0 || q.in_buf >= (int)d->bufs.size() || q.out_buf < 0 || q.out_buf >= (int)d->bufs.size())
    return fail(2, "op references unknown buffer");
  const vgh_buf_desc& ib = d->bufs[q.in_buf];
  const vgh_buf_desc& ob = d->bufs[q.out_buf];
  ConvLaunch& L = o.L;
  memset(&L, 0, sizeof(L));
  if (q.cin % 32) return fail(2, "cin %d not a multiple of 32", q.cin);
  o.bk = q.cin % 64 == 0 ? 64 : 32;
  L.B = d->B;
  const bool patch_op = q.level > 0;  // survivor patches: one stacked image whose live height is known on the device only
  if (patch_op) {
    if (!ib.stack || !ob.stack || q.level > 3 || q.stride != 1 || q.up || !swap_eligible(q, ob))
      return fail(2, "patch-level conv must map stacked buffers, stride 1, and be eligible for the swapped kernel");
    L.B = 1;
  } else if (ib.stack || ob.stack) {
    return fail(2, "dense op on a stacked buffer");
  }
  L.stride = q.stride;
  L.Ho = q.up ? ib.H : (ib.H + q.stride - 1) / q.stride;
  L.Wo = q.up ? ib.W : (ib.W + q.stride - 1) / q.stride;
  // un-tuned default: the swapped kernel for Cout <= 128 on large maps and - with the tap reuse, which measured
  // faster than every other variant on every 3x3 stride-1 layer of the network - for all of those
  L.swap = d->split ? 0 : patch_op ? 1 : o.cfg_swap >= 0 ? o.cfg_swap
                           : (swap_eligible(q, ob) && (swap_forced() || (xr_eligible(q, ob) && xr_default()) ||
                                                       (q.cout <= 128 && !ob.fp32 && L.Ho * L.Wo >= 1024)) ? 1 : 0);
  if (L.swap) swap_groups(q, ob, L.ngroups, L.gw);
  if (L.swap && !swap_eligible(q, ob)) return fail(2, "op not eligible for the swapped kernel");
  L.xr = (L.swap && xr_eligible(q, ob)) ? (o.cfg_xr >= 0 ? o.cfg_xr : xr_default()) : 0;
  // pairs need a group width whose halves are whole 8-row swizzle atoms
  L.cluster = (L.swap && L.gw % 16 == 0 && (o.cfg_cluster >= 0 ? o.cfg_cluster : cluster_default())) ? 2 : 1;
  // CTA pairs: two 128-channel groups per MMA; needs whole 64-channel K blocks and an even tile height (half a tile per CTA)
  const bool pair_ok = L.xr && !patch_op && L.gw == 128 && L.ngroups % 2 == 0 && o.bk == 64 && L.cluster == 1;
  L.pair = (pair_ok && (o.cfg_pair >= 0 ? o.cfg_pair : pair_default())) ? 1 : 0;
  if (o.cfg_tw > 0) { L.tw = o.cfg_tw; L.th = o.cfg_th; }
  else if (L.xr) pick_tile_swap_xr(L.Ho, L.Wo, L.tw, L.th);
  else if (L.swap) pick_tile_swap(L.Ho, L.Wo, L.tw, L.th);
  else pick_tile(L.Ho, L.Wo, d->B, o.cfg_tail != 0, L.tw, L.th);
  if (L.pair && o.cfg_tw <= 0 && L.th % 2 && L.th > 1) --L.th;   // un-tuned default: an even tile height (half a tile per CTA)
  if (L.pair && (L.th % 2 || (L.tw * L.th) % 16)) L.pair = 0;
  if (L.xr && L.tw % 8) return fail(2, "tap-reuse tiles must be a multiple of 8 pixels wide");
  L.tiles_x = (L.Wo + L.tw - 1) / L.tw;
  L.tiles_y = (L.Ho + L.th - 1) / L.th;
  // left-over rows: when the last row group of every image is mostly empty, pack those rows of
  // several images into shared tiles instead (fewer, fuller work items)
  L.tail_rows = 0; L.tail_imgs = 0; L.n_tail_tiles = 0;
  if (!L.swap && o.cfg_tail != 0 && L.Ho % L.th != 0) {
    const int tr = L.Ho % L.th;
    const int imgs = 128 / (L.tw * tr);
    if (imgs >= 2) {
      L.tail_rows = tr;
      L.tail_imgs = imgs > d->B ? d->B : imgs;
      L.tiles_y = L.Ho / L.th;  // full row groups only
      L.n_tail_tiles = L.tiles_x * ((d->B + L.tail_imgs - 1) / L.tail_imgs);
    }
  }
  L.cin_off = q.in_coff;
  L.cin = q.cin;
  L.kw = q.ksize;
  L.ntaps = q.ksize * q.ksize;
  L.pad = q.ksize / 2;
  L.n_total = q.cout;
  L.block_n = q.block_n > 0 ? q.block_n : auto_block_n(q.cout, q.up, q.up_cout);
  if (L.block_n % 16 || L.block_n > 256) return fail(2, "bad block_n %d", L.block_n);
  if (q.n_pad < ((q.cout + L.block_n - 1) / L.block_n) * L.block_n) return fail(2, "n_pad %d too small for block_n %d", q.n_pad, L.block_n);
  if (q.k_total != L.ntaps * q.cin) return fail(2, "k_total mismatch");
  L.out_cstride = ob.C;
  L.out_coff = q.out_coff;
  L.out_H = ob.H;
  L.out_W = ob.W;
  L.up = q.up;
  L.up_cout = q.up_cout;
  L.relu = q.relu;
  L.out_fp32 = ob.fp32;
  L.split = (d->split && !ob.fp32) ? 1 : 0;
  L.f16 = d->f16 ? 1 : 0;
  if (L.split && (q.cout % 32 || q.out_coff % 192 || (q.up && q.up_cout % 32))) return fail(2, "parity mode: channel slices must be whole 32-channel granules");
  if (q.up) {
    if (ob.H != 2 * ib.H || ob.W != 2 * ib.W) return fail(2, "transpose conv output buffer must be 2x the input");
  } else if (ob.H != L.Ho || ob.W != L.Wo) {
    return fail(2, "conv output buffer spatial mismatch (%dx%d vs %dx%d)", ob.H, ob.W, L.Ho, L.Wo);
  }
  L.bias = d->bias + q.b_off;
  L.out = d->buf_ptr[q.out_buf];
  L.dyn_rows = patch_op ? d->level_rows + (q.level - 1) : nullptr;
  if (q.res_buf >= 0) {
    const vgh_buf_desc& rb = d->bufs[q.res_buf];
    if (rb.H != L.Ho || rb.W != L.Wo || rb.fp32) return fail(2, "residual buffer mismatch");
    L.res = static_cast<const __nv_bfloat16*>(d->buf_ptr[q.res_buf]);
    L.res_cstride = rb.C;
    L.res_coff = q.res_coff;
    L.res_alpha = q.res_alpha;
  }
  L.mt = L.swap ? 1 : (o.cfg_mt > 0 ? o.cfg_mt : conv_default_mt(L.block_n));
  if (L.swap && L.xr) {
    // two rings: pixel tiles with halo rows (xslots deep) and weight k-blocks (`stages` deep)
    const int x_slot = (L.tw * (L.pair ? L.th / 2 + 2 : L.th + 2) * o.bk * 2 + 1023) & ~1023;
    // 32-channel K blocks: one weight stage holds the k-blocks of all three row taps of a column shift (24 KB, 6 MMAs per
    // barrier round trip instead of 2) - with 2 MMAs per stage the single MMA-issuing thread is the critical path
    L.ks = (o.bk == 32 && !L.pair && (o.cfg_ks > 0 ? o.cfg_ks == 3 : wgroup_default())) ? 3 : 1;
    const int w_bytes = L.ks * 128 * o.bk * 2;
    const int staging = (L.tw * L.th * L.gw * (ob.fp32 ? 4 : 2) + 1023) & ~1023;
    // residual layers on 32-channel K blocks: a second staging tile lets the next item's residual tile arrive while this
    // item is multiplied (the small K blocks leave the room); falls back to one tile when the rings would not fit
    L.stg_bufs = (q.res_buf >= 0 && o.bk == 32 && !L.pair && stg2_default() &&
                  224 * 1024 - 2 * staging - 2 * x_slot >= (L.ks == 3 ? 2 : 4) * w_bytes) ? 2 : 1;
    const int budget = 224 * 1024 - L.stg_bufs * staging;
    int xs = o.cfg_xslots > 0 ? o.cfg_xslots : 3;
    while (xs > 2 && budget - xs * x_slot < (L.ks == 3 ? 2 : 4) * w_bytes) --xs;
    int ws = (budget - xs * x_slot) / w_bytes;
    if (ws < (L.ks == 3 ? 2 : 3)) return fail(2, "tap-reuse variant does not fit shared memory (tile %dx%d, bk %d)", L.tw, L.th, o.bk);
    L.xslots = xs;
    // 32-channel K blocks (96-channel layers) have 8 KB weight stages: 8 of them leave half of the shared memory - i.e. of
    // the bytes in flight that hide the L2 latency - unused, so the ring may go 16 deep
    const int ring_cap = deep_rings() ? 16 : 8;
    L.stages = ws > ring_cap ? ring_cap : ws;
  } else if (L.swap) {
    const int stage_bytes = 128 * o.bk * 2 + ((L.tw * L.th * o.bk * 2 + 1023) & ~1023);
    const int staging = (L.tw * L.th * L.gw * (ob.fp32 ? 4 : 2) + 1023) & ~1023;  // epilogue tile [pixels][channels]
    const int num_kb = L.ntaps * (q.cin / o.bk);
    // shallow-K layers are epilogue-bound: give them a second staging tile; deep-K layers need the stages
    L.stg_bufs = (num_kb <= 8 && 2 * staging + 3 * stage_bytes <= 220 * 1024) ? 2 : 1;
    L.ks = (o.cfg_ks > 1 && num_kb % o.cfg_ks == 0) ? o.cfg_ks : 1;
    int st = o.cfg_stages > 0 ? o.cfg_stages : (222 * 1024 - L.stg_bufs * staging) / (stage_bytes * L.ks);
    if (st < 2) { L.ks = 1; st = (222 * 1024 - L.stg_bufs * staging) / stage_bytes; }
    L.stages = st > 8 ? 8 : (st < 2 ? 2 : st);
  } else {
    L.stages = o.cfg_stages > 0 ? o.cfg_stages : conv_pick_stages(L.block_n, o.bk, L.mt);
  }
  if (L.mt * L.block_n > 512) return fail(2, "mt %d x block_n %d exceeds TMEM", L.mt, L.block_n);
  conv_finalize(L);
  if (ib.fp32) return fail(2, "conv input must be bf16");
  int rc = conv_make_tensor_maps(L, d->buf_ptr[q.in_buf], ib.C, ib.H, ib.W, d->weights + q.w_off, q.k_total, q.n_pad,
                                 o.bk);
  if (rc) return fail(3, "%s", conv_last_error());
  if (L.swap) {
    rc = conv_make_io_maps(L, d->buf_ptr[q.out_buf], q.res_buf >= 0 ? d->buf_ptr[q.res_buf] : nullptr);
    if (rc) return fail(3, "%s", conv_last_error());
  }
  return 0;
}

extern "C" void vgh_detector_destroy(vgh_detector* d) {
  if (!d) return;
  if (d->graph) cudaGraphExecDestroy(d->graph);
  if (d->cap_stream) cudaStreamDestroy(d->cap_stream);
  for (auto& sl : d->slots) {
    void* sp[] = {sl.in_dev, sl.params, sl.verts, sl.kboxes, sl.kscores, sl.cnt, sl.total};
    for (void* q : sp) cudaFree(q);
    cudaEvent_t ev[] = {sl.h2d_done, sl.in_consumed, sl.compute_done, sl.d2h_done};
    for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
  }
  for (cudaStream_t t : {d->h2d_stream, d->d2h_stream, d->pipe_stream}) if (t) cudaStreamDestroy(t);
  for (cudaStream_t t : d->lane_streams) cudaStreamDestroy(t);
  for (cudaEvent_t e : d->op_events) if (e) cudaEventDestroy(e);
  if (d->phase_event) cudaEventDestroy(d->phase_event);
  for (void* p : d->buf_ptr) cudaFree(p);
  void* ptrs[] = {d->weights, d->bias, d->input, d->boxes, d->scores, d->keep_boxes,
                  d->keep_scores, d->keep_idx, d->keep_cnt, d->offsets, d->head_img, d->params, d->head_xform,
                  d->verts, d->rot, d->img_xform, d->head_level, d->head_patch, d->patch_src, d->level_rows, d->push_counter,
                  d->push_status};
  for (void* p : ptrs) cudaFree(p);
  delete d;
}

template <typename T>
static cudaError_t dmalloc(T** p, size_t count) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, count * sizeof(T));
  return e;
}

extern "C" int vgh_detector_create(const vgh_net_desc* n, const vgh_flame* flame, vgh_detector** out) {
  if (!n || !out) return fail(1, "null argument");
  if (n->batch < 1 || n->image_size < 64 || n->image_size % 32) return fail(1, "bad batch / image size");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) return fail(9, "no CUDA device (no CPU fallback)");
  vgh_detector* d = new vgh_detector();
  d->B = n->batch;
  d->S = n->image_size;
  d->keep_k = n->keep_k > 0 ? n->keep_k : 100;
  d->flame = flame;
  {
    const char* e = getenv("VGGHEADS_B200_SINGLE_LANE");
    d->multi_lane = !(e && e[0] == '1');
  }
  int rc = 0;
  auto bail = [&](int code) { vgh_detector_destroy(d); return code; };

  d->bufs.assign(n->bufs, n->bufs + n->n_bufs);
  d->n_dense_ops = (n->n_dense_ops > 0 && n->n_dense_ops < n->n_ops) ? n->n_dense_ops : n->n_ops;
  d->sparse = d->n_dense_ops < n->n_ops;
  d->split = n->split != 0;
  d->f16 = n->act_f16 != 0;
  if (d->split && d->sparse) return bail(fail(2, "parity (split) mode needs the dense-heads plan"));
  if (d->split && d->f16) return bail(fail(2, "parity (split) mode stores bf16 terms: act_f16 must be 0"));
  // survivor patches are addressed as image << 20 | y << 10 | x (aux_kernels.cu: patch_src)
  if (d->sparse && (d->B > 4096 || d->S / 8 > 1024)) return bail(fail(1, "sparse heads: batch <= 4096 and image size <= 8192"));
  d->patch_cap = d->B * d->keep_k;
  if (d->sparse && (dmalloc(&d->head_level, (size_t)d->patch_cap) != cudaSuccess || dmalloc(&d->head_patch, (size_t)d->patch_cap) != cudaSuccess ||
                    dmalloc(&d->patch_src, (size_t)3 * d->patch_cap) != cudaSuccess || dmalloc(&d->level_rows, 4) != cudaSuccess))
    return bail(fail(4, "patch table allocation failed"));
  for (const vgh_buf_desc& b : d->bufs) {
    if (b.stack && (b.W != kPatch || b.H != d->patch_cap * kPatch)) return bail(fail(2, "stacked buffer must be [batch*keep_k*%d, %d, C]", kPatch, kPatch));
    const size_t bytes = static_cast<size_t>(b.stack ? 1 : d->B) * b.H * b.W * b.C * (b.fp32 ? 4 : 2);
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess)
      return bail(fail(4, "activation buffer allocation failed (%zu bytes)", bytes));
    d->buf_ptr.push_back(p);
    d->buf_bytes.push_back(bytes);
  }
  if (dmalloc(&d->weights, (size_t)n->n_weights) != cudaSuccess || dmalloc(&d->bias, (size_t)n->n_bias) != cudaSuccess)
    return bail(fail(4, "weight allocation failed"));
  cudaMemcpy(d->weights, n->weights_host, n->n_weights * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d->bias, n->bias_host, n->n_bias * 4, cudaMemcpyHostToDevice);

  for (int i = 0; i < n->n_ops; ++i) {
    OpRt o;
    o.d = n->ops[i];
    o.bk = 0;
    if (o.d.kind == VGH_OP_CONV) {
      rc = build_conv(d, o);
      if (rc) return bail(rc);
    }
    if ((o.d.level > 0) != (i >= d->n_dense_ops)) return bail(fail(2, "op %d: patch-level ops must be exactly the ops after n_dense_ops", i));
    d->ops.push_back(o);
  }
  // decode bookkeeping
  int a_off = 0;
  for (int l = 0; l < 3; ++l) {
    const vgh_buf_desc& rb = d->bufs[n->reg_buf[l]];
    const vgh_buf_desc& fb = d->bufs[n->flame_buf[l]];
    if (!rb.fp32 || !fb.fp32 || (rb.H != fb.H && !fb.stack) || (fb.stack != 0) != d->sparse) return bail(fail(2, "raw head buffers must be fp32 (flame: dense map, or patch stack in a sparse plan)"));
    d->lv.reg[l] = static_cast<const float*>(d->buf_ptr[n->reg_buf[l]]);
    d->lv.flame[l] = static_cast<const float*>(d->buf_ptr[n->flame_buf[l]]);
    d->lv.a_off[l] = a_off;
    d->lv.W[l] = rb.W;
    d->lv.hw[l] = rb.H * rb.W;
    d->lv.stride[l] = static_cast<float>(d->S / rb.H);
    d->lv.reg_cstride = rb.C;
    d->lv.flame_cstride = fb.C;
    a_off += rb.H * rb.W;
  }
  d->lv.a_off[3] = a_off;
  d->lv.head_level = d->sparse ? d->head_level : nullptr;
  d->lv.head_patch = d->sparse ? d->head_patch : nullptr;
  d->A = a_off;
  const size_t B = d->B, A = d->A, K = d->keep_k, cap = B * K;
  if (dmalloc(&d->input, B * d->S * d->S * 3) != cudaSuccess || dmalloc(&d->boxes, B * A * 4) != cudaSuccess ||
      dmalloc(&d->scores, B * A) != cudaSuccess || dmalloc(&d->keep_boxes, cap * 4) != cudaSuccess ||
      dmalloc(&d->keep_scores, cap) != cudaSuccess || dmalloc(&d->keep_idx, cap) != cudaSuccess ||
      dmalloc(&d->keep_cnt, B) != cudaSuccess || dmalloc(&d->offsets, B + 1) != cudaSuccess ||
      dmalloc(&d->head_img, cap) != cudaSuccess || dmalloc(&d->params, cap * VGH_NUM_PARAMS) != cudaSuccess ||
      dmalloc(&d->head_xform, cap * 3) != cudaSuccess || dmalloc(&d->verts, cap * VGH_NUM_VERTS * 3) != cudaSuccess ||
      dmalloc(&d->rot, cap * 9) != cudaSuccess || dmalloc(&d->img_xform, B * 3) != cudaSuccess ||
      dmalloc(&d->push_counter, 1) != cudaSuccess || dmalloc(&d->push_status, 1) != cudaSuccess)
    return bail(fail(4, "result buffer allocation failed"));
  {
    std::vector<float> xf(B * 3, 0.f);
    for (size_t i = 0; i < B; ++i) xf[i * 3 + 2] = 1.f;
    cudaMemcpy(d->img_xform, xf.data(), xf.size() * 4, cudaMemcpyHostToDevice);
  }
  CUDA_OK(cudaDeviceSynchronize());
  *out = d;
  return 0;
}

static int launch_op(vgh_detector* d, OpRt& o, const uint8_t* images, cudaStream_t s) {
  int rc = 0;
  switch (o.d.kind) {
    case VGH_OP_STEM:
      if (d->f16) return fail(2, "the two-op (im2col) stem writes bf16; fp16 activations use the fused stem (VGH_OP_STEM_CONV)");
      rc = stem_pack_launch(images, static_cast<__nv_bfloat16*>(d->buf_ptr[o.d.out_buf]), d->B, d->S, d->split ? 1 : 0, s);
      if (rc) return fail(5, "stem launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    case VGH_OP_STEM_CONV: {
      const vgh_buf_desc& ob = d->bufs[o.d.out_buf];
      if (ob.fp32 || ob.H != d->S / 2 || o.d.k_total != 32 || o.d.cout != 64 || o.d.out_coff != 0)
        return fail(2, "fused stem: expects a bf16 [S/2,S/2,>=64] output at channel 0 and 32-wide packed taps");
      rc = stem_conv_launch(images, d->weights + o.d.w_off, d->bias + o.d.b_off, static_cast<__nv_bfloat16*>(d->buf_ptr[o.d.out_buf]), d->B, d->S,
                            ob.C, o.d.relu, d->f16 ? 1 : 0, s);
      if (rc) return fail(5, "fused stem launch failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    case VGH_OP_CONV:
      rc = conv_launch(o.L, o.bk, s);
      if (rc) return fail(5, "%s", conv_last_error());
      break;
    case VGH_OP_SPP: {
      const vgh_buf_desc& b = d->bufs[o.d.in_buf];
      rc = d->split ? spp_pool_split_launch(static_cast<__nv_bfloat16*>(d->buf_ptr[o.d.in_buf]), d->B, b.H, b.W, o.d.cin, s)
                    : spp_pool_launch(static_cast<__nv_bfloat16*>(d->buf_ptr[o.d.in_buf]), d->B, b.H, b.W, o.d.cin, d->f16 ? 1 : 0, s);
      if (rc) return fail(5, "spp launch failed");
      break;
    }
    case VGH_OP_PATCH_GATHER: {
      const vgh_buf_desc& fb = d->bufs[o.d.in_buf];
      const vgh_buf_desc& pb = d->bufs[o.d.out_buf];
      const int l = o.d.level - 1;
      rc = patch_gather_launch(static_cast<const __nv_bfloat16*>(d->buf_ptr[o.d.in_buf]), fb.H, fb.W, fb.C, o.d.in_coff, o.d.cin,
                               static_cast<__nv_bfloat16*>(d->buf_ptr[o.d.out_buf]), pb.C, o.d.out_coff,
                               d->patch_src + static_cast<size_t>(l) * d->patch_cap, d->level_rows + l, d->patch_cap, s);
      if (rc) return fail(5, "patch gather launch failed");
      break;
    }
    case VGH_OP_PATCH_MASK: {
      const vgh_buf_desc& pb = d->bufs[o.d.out_buf];
      const int l = o.d.level - 1;
      rc = patch_mask_launch(static_cast<__nv_bfloat16*>(d->buf_ptr[o.d.out_buf]), pb.C, o.d.out_coff, o.d.cin, d->lv.hw[l] / d->lv.W[l],
                             d->lv.W[l], d->patch_src + static_cast<size_t>(l) * d->patch_cap, d->level_rows + l, d->patch_cap, s);
      if (rc) return fail(5, "patch mask launch failed");
      break;
    }
    default:
      return fail(5, "unknown op kind %d", o.d.kind);
  }
  return 0;
}

// Executes the plan on `s` (lane 0) plus side streams for the other lanes.  Independent branches of
// the graph (the three head levels, the parallel tower chains) run on their own lanes so that the
// pipeline fill/drain of one kernel is covered by CTAs of another; cross-lane ordering comes from
// buffer-level dependency tracking (RAW / WAR / WAW per activation buffer), expressed as events -
// which CUDA-graph capture turns into graph edges.
static int run_ops(vgh_detector* d, int op_begin, int op_end, const uint8_t* images, cudaStream_t s, int* launches) {
  const int n_ops = static_cast<int>(d->ops.size());
  const int n_bufs = static_cast<int>(d->bufs.size());
  if (d->op_events.size() < static_cast<size_t>(n_ops)) {
    d->op_events.resize(n_ops, nullptr);
    for (auto& e : d->op_events) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  int max_lane = 0;
  for (const OpRt& o : d->ops) max_lane = o.d.lane > max_lane ? o.d.lane : max_lane;
  if (!d->multi_lane) max_lane = 0;
  while (static_cast<int>(d->lane_streams.size()) < max_lane) {
    cudaStream_t t;
    CUDA_OK(cudaStreamCreateWithFlags(&t, cudaStreamNonBlocking));
    d->lane_streams.push_back(t);
  }
  auto stream_of = [&](int lane) { return lane == 0 ? s : d->lane_streams[lane - 1]; };
  const int L = max_lane + 1;
  std::vector<int> last_w(static_cast<size_t>(n_bufs) * L, -1), last_r(static_cast<size_t>(n_bufs) * L, -1);
  std::vector<char> recorded(n_ops, 0);
  std::vector<int> last_op_on_lane(L, -1);
  // later phase: whatever lane 0 has queued so far (select/NMS, patch tables) precedes every lane; a side lane joins
  // when its first op of the phase is issued (lanes without ops in this phase stay out of the capture)
  std::vector<char> lane_joined(L, 0);
  if (L > 1 && op_begin > 0) {
    if (!d->phase_event) CUDA_OK(cudaEventCreateWithFlags(&d->phase_event, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(d->phase_event, s));
  }
  auto depend = [&](int lane, int op_j) -> int {  // lane must wait for op_j (which ran on another lane)
    if (op_j < 0) return 0;
    const int lj = d->multi_lane ? d->ops[op_j].d.lane : 0;
    if (lj == lane) return 0;
    if (!recorded[op_j]) {
      CUDA_OK(cudaEventRecord(d->op_events[op_j], stream_of(lj)));
      recorded[op_j] = 1;
    }
    CUDA_OK(cudaStreamWaitEvent(stream_of(lane), d->op_events[op_j], 0));
    return 0;
  };
  for (int i = op_begin; i < op_end; ++i) {
    OpRt& o = d->ops[i];
    const int lane = d->multi_lane ? o.d.lane : 0;
    if (lane > 0 && op_begin > 0 && !lane_joined[lane]) {
      CUDA_OK(cudaStreamWaitEvent(stream_of(lane), d->phase_event, 0));
      lane_joined[lane] = 1;
    }
    if (L > 1) {
      int reads[2] = {(o.d.kind == VGH_OP_STEM || o.d.kind == VGH_OP_STEM_CONV) ? -1 : o.d.in_buf, o.d.kind == VGH_OP_CONV ? o.d.res_buf : -1};
      for (int b : reads) {
        if (b < 0) continue;
        for (int l2 = 0; l2 < L; ++l2) { int rc = depend(lane, last_w[b * L + l2]); if (rc) return rc; }
      }
      const int wb = o.d.kind == VGH_OP_SPP ? o.d.in_buf : o.d.out_buf;
      for (int l2 = 0; l2 < L; ++l2) {
        int rc = depend(lane, last_w[wb * L + l2]);
        if (!rc) rc = depend(lane, last_r[wb * L + l2]);
        if (rc) return rc;
      }
      for (int b : reads)
        if (b >= 0) last_r[b * L + lane] = i;
      last_w[wb * L + lane] = i;
    }
    int rc = launch_op(d, o, images, stream_of(lane));
    if (rc) return rc;
    last_op_on_lane[lane] = i;
    if (launches) ++*launches;
  }
  for (int l2 = 1; l2 < L; ++l2) {  // join every side lane back into lane 0
    int rc = depend(0, last_op_on_lane[l2]);
    if (rc) return rc;
  }
  return 0;
}

static int run_forward(vgh_detector* d, const uint8_t* images, cudaStream_t s, int* launches) {
  int rc = run_ops(d, 0, d->n_dense_ops, images, s, launches);
  if (rc) return rc;
  if (box_decode_launch(d->lv, d->boxes, d->scores, d->B, d->A, s)) return fail(5, "box decode launch failed");
  if (launches) ++*launches;
  if (d->ovr_boxes && d->ovr_scores) {
    CUDA_OK(cudaMemcpyAsync(d->boxes, d->ovr_boxes, sizeof(float) * 4 * d->B * d->A, cudaMemcpyDeviceToDevice, s));
    CUDA_OK(cudaMemcpyAsync(d->scores, d->ovr_scores, sizeof(float) * d->B * d->A, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

static int run_post(vgh_detector* d, float conf, float iou, int top_k, const float* xform, cudaStream_t s, int* launches) {
  int rc = select_nms_launch(d->boxes, d->scores, d->B, d->A, conf, iou, top_k, d->keep_k, d->keep_idx, d->keep_cnt,
                             d->keep_boxes, d->keep_scores, s, g_err, sizeof(g_err));
  if (rc) return rc;
  if (head_offsets_launch(d->keep_cnt, d->B, d->offsets, d->offsets + d->B, s)) return fail(5, "offsets launch failed");
  if (d->sparse) {  // FLAME branch of the heads on survivor patches only
    if (patch_assign_launch(d->lv, d->keep_idx, d->keep_cnt, d->offsets, d->B, d->keep_k, d->patch_cap, d->head_level,
                            d->head_patch, d->patch_src, d->level_rows, s))
      return fail(5, "patch assign launch failed");
    if (launches) ++*launches;
    rc = run_ops(d, d->n_dense_ops, static_cast<int>(d->ops.size()), nullptr, s, launches);
    if (rc) return rc;
  }
  if (flame_gather_launch(d->lv, d->keep_idx, d->keep_cnt, d->B, d->keep_k, xform, d->offsets, d->params, d->head_xform,
                          d->head_img, s))
    return fail(5, "gather launch failed");
  if (launches) *launches += 3;
  if (d->flame) {
    rc = flame_decode_launch(d->flame->model, d->params, d->B * d->keep_k, d->offsets + d->B, 128, 64, d->head_xform,
                             nullptr, d->rot, d->verts, s, g_err, sizeof(g_err));
    if (rc) return rc;
    if (launches) ++*launches;
  }
  return 0;
}

// Per-op configuration search on the device: every conv op is timed under a few (mt, stages)
// candidates (inputs are whatever the buffers hold - timing does not depend on values) and the
// fastest is kept.  Invalidates the captured graph.
extern "C" int vgh_detector_autotune(vgh_detector* d, int iters, void* stream) {
  if (!d) return fail(1, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d->graph) { cudaGraphExecDestroy(d->graph); d->graph = nullptr; }
  if (d->split) return 0;  // parity mode runs one fixed kernel configuration
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  if (iters < 1) iters = 3;
  // patch-level convs are tuned at a nominal load of 8 survivors per image spread over the three head levels
  const int nominal_rows[4] = {d->B * 4 * kPatch, d->B * 3 * kPatch, d->B * 2 * kPatch, 0};
  int saved_rows[4] = {0, 0, 0, 0};
  if (d->sparse) {
    CUDA_OK(cudaMemcpy(saved_rows, d->level_rows, sizeof(saved_rows), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(d->level_rows, nominal_rows, sizeof(nominal_rows), cudaMemcpyHostToDevice));
  }
  for (OpRt& o : d->ops) {
    if (o.d.kind != VGH_OP_CONV) continue;
    float best = 1e30f;
    int best_mt = 0, best_st = 0, best_swap = 0, best_tw = 0, best_th = 0, best_ks = 1;
    const int bn = o.L.block_n;
    int best_xr = 0, best_xs = 0, best_cl = 0, best_pr = 0;
    if (xr_eligible(o.d, d->bufs[o.d.out_buf]) && xr_default()) {
      // candidate tiles: every admissible shape ranked by a coarse cost model - waves of work items over the SMs
      // x (MMA columns + a share for the pixel rows each item ingests + a fixed per-item cost) - and the best six
      // are MEASURED; the shape the un-tuned default would take is always among them
      struct Cand { int tw, th; double est; };
      Cand cands[8];
      int n_cand = 0;
      {
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int Ho = o.L.Ho, Wo = o.L.Wo, wo8 = (Wo + 7) / 8 * 8;
        int G = 1, gw = 0;
        swap_groups(o.d, d->bufs[o.d.out_buf], G, gw);
        int dtw, dth;
        pick_tile_swap_xr(Ho, Wo, dtw, dth);
        cands[n_cand++] = {dtw, dth, -1.0};
        for (int th = 1; th <= Ho && th <= 32; ++th)
          for (int tw = 8; tw <= wo8 && tw * th <= 256; tw += 8) {
            if ((tw * th) % 16 || tw * th < 64 || (tw == dtw && th == dth)) continue;
            const long items = static_cast<long>((Wo + tw - 1) / tw) * ((Ho + th - 1) / th) * d->B * G;
            const double waves = static_cast<double>((items + sms - 1) / sms);
            const double est = waves * (tw * th + 0.35 * 3.0 * tw * (th + 2) + 40.0);
            if (n_cand < 7) cands[n_cand++] = {tw, th, est};  // slots 1..6: the cheapest so far, ascending
            else if (est < cands[6].est) cands[6] = {tw, th, est};
            else continue;
            for (int i = n_cand - 1; i > 1 && cands[i].est < cands[i - 1].est; --i) { Cand c = cands[i]; cands[i] = cands[i - 1]; cands[i - 1] = c; }
          }
      }
      for (int ci = 0; ci < n_cand; ++ci) {
        const int ptw = cands[ci].tw, pth = cands[ci].th;
        for (int xs : {2, 3, 4}) for (int cl = 0; cl <= cluster_default(); ++cl) for (int pr = 0; pr <= pair_default(); ++pr) for (int wg : {1, 3}) {
          if (xs == 4 && (o.bk != 32 || !deep_rings())) continue;   // a fourth pixel slot only fits next to the small 32-channel K blocks
          if (wg == 3 && (o.bk != 32 || pr || !wgroup_default())) continue;
          OpRt t = o;
          t.cfg_ks = wg; t.cfg_swap = 1; t.cfg_mt = 0; t.cfg_stages = 0; t.cfg_xr = 1; t.cfg_xslots = xs; t.cfg_cluster = cl; t.cfg_pair = pr;
          t.cfg_tw = ptw; t.cfg_th = pth;
          if (build_conv(d, t) || t.L.xslots != xs || t.L.cluster != cl + 1 || t.L.pair != pr || t.L.ks != wg || conv_launch(t.L, t.bk, s)) continue;
          float ms = 1e30f;
          for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0, s);
            for (int i = 0; i < iters; ++i) conv_launch(t.L, t.bk, s);
            cudaEventRecord(e1, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) return fail(7, "autotune (tap-reuse) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            float m = 0.f;
            cudaEventElapsedTime(&m, e0, e1);
            if (m < ms) ms = m;
          }
          if (ms < best) { best = ms; best_swap = 1; best_mt = 1; best_st = 0; best_tw = t.cfg_tw; best_th = t.cfg_th; best_ks = wg; best_xr = 1; best_xs = xs; best_cl = cl; best_pr = pr; }
        }
      }
    }
    if (swap_eligible(o.d, d->bufs[o.d.out_buf])) {
      // pixel-tile size trades MMA width against pipeline depth; 32-channel K blocks: three k-blocks per stage as well (two
      // MMAs per barrier round trip leave the MMA-issuing thread on the critical path)
      for (int max_px : {256, 192, 128}) for (int cl = 0; cl <= cluster_default(); ++cl) for (int kg : {1, 3}) {
        if (kg == 3 && (o.bk != 32 || !wgroup_default() || (o.d.ksize * o.d.ksize * (o.d.cin / 32)) % 3)) continue;
        OpRt t = o;
        t.cfg_ks = kg; t.cfg_xr = 0; t.cfg_cluster = cl; t.cfg_pair = 0;
        t.cfg_swap = 1; t.cfg_mt = 0; t.cfg_stages = 0;
        pick_tile_swap(o.L.Ho, o.L.Wo, t.cfg_tw, t.cfg_th, max_px);
        if (max_px != 256 && t.cfg_tw * t.cfg_th > max_px) continue;
        if (build_conv(d, t) || t.L.cluster != cl + 1 || t.L.ks != kg || conv_launch(t.L, t.bk, s)) continue;
        float ms = 1e30f;
        for (int rep = 0; rep < 2; ++rep) {  // best of two timed bursts: the clock / power state is noisy
          cudaEventRecord(e0, s);
          for (int i = 0; i < iters; ++i) conv_launch(t.L, t.bk, s);
          cudaEventRecord(e1, s);
          if (cudaStreamSynchronize(s) != cudaSuccess) return fail(7, "autotune (swap) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
          float m = 0.f;
          cudaEventElapsedTime(&m, e0, e1);
          if (m < ms) ms = m;
        }
        if (ms < best) { best = ms; best_swap = 1; best_mt = 1; best_st = 0; best_tw = t.cfg_tw; best_th = t.cfg_th; best_ks = t.cfg_ks; best_xr = 0; best_cl = cl; }
      }
    }
    for (int mt : {1, 2, 4}) {
      if (mt * bn > 512 || o.d.level > 0) continue;  // patch-level convs: swapped kernel only
      for (int variant = 0; variant < 2; ++variant) {
        OpRt t = o;
        t.cfg_swap = 0; t.cfg_tw = 0; t.cfg_th = 0;
        t.cfg_mt = mt;
        const int stage_bytes = (128 * mt + bn) * t.bk * 2;
        int st = (variant == 0 ? 200 * 1024 : 100 * 1024) / stage_bytes;  // 1 CTA/SM deep vs 2 CTAs/SM
        if (st > 8) st = 8;
        if (st < 2) { if (variant == 1) continue; st = 2; }
        if ((size_t)st * stage_bytes > 210 * 1024) continue;
        t.cfg_stages = st;
        if (build_conv(d, t)) continue;
        if (conv_launch(t.L, t.bk, s)) continue;  // warm-up (also sets the smem attribute)
        float ms = 1e30f;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0, s);
          for (int i = 0; i < iters; ++i) conv_launch(t.L, t.bk, s);
          cudaEventRecord(e1, s);
          if (cudaStreamSynchronize(s) != cudaSuccess) return fail(7, "autotune launch failed: %s", cudaGetErrorString(cudaGetLastError()));
          float m = 0.f;
          cudaEventElapsedTime(&m, e0, e1);
          if (m < ms) ms = m;
        }
        if (ms < best) { best = ms; best_mt = mt; best_st = st; best_swap = 0; best_xr = 0; best_pr = 0; }
      }
    }
    if (best_mt) {
      o.cfg_swap = best_swap;
      o.cfg_ks = best_swap ? best_ks : 0;
      o.cfg_tw = best_swap ? best_tw : 0;
      o.cfg_th = best_swap ? best_th : 0;
      o.cfg_mt = best_mt;
      o.cfg_stages = best_st;
      o.cfg_xr = best_swap ? best_xr : 0;
      o.cfg_cluster = best_swap ? best_cl : 0;
      o.cfg_pair = best_xr ? best_pr : 0;
      o.cfg_xslots = best_xr ? best_xs : 0;
      int rc = build_conv(d, o);
      if (rc) return rc;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (d->sparse) CUDA_OK(cudaMemcpy(d->level_rows, saved_rows, sizeof(saved_rows), cudaMemcpyHostToDevice));
  return 0;
}
// Reports the configuration of conv op i: out[0..5] = mt, stages, block_n, bk, tw, th.
extern "C" int vgh_detector_op_config(const vgh_detector* d, int op, int32_t* out6) {
  if (!d || op < 0 || op >= (int)d->ops.size() || !out6) return fail(1, "bad argument");
  const OpRt& o = d->ops[op];
  out6[0] = o.L.swap ? -(o.L.ks + 10 * (o.L.cluster - 1) + 20 * o.L.pair) : o.L.mt;  // swapped: -ks, -1x = weight-multicast pairs, -2x = cta_group::2 MMA pairs
  out6[1] = (o.L.swap && o.L.xr) ? 100 * o.L.xslots + o.L.stages : o.L.stages;  // tap reuse: 100*pixel slots + weight slots
  out6[2] = o.L.block_n; out6[3] = o.bk; out6[4] = o.L.tw; out6[5] = o.L.th;
  return 0;
}

// Eager execution with one CUDA-event pair around every plan op and every post-processing stage (in execution
// order: dense ops, box decode, select/NMS, [sparse heads: patch tables + patch ops], survivor rows, FLAME decode).
extern "C" int vgh_detector_profile(vgh_detector* d, int iters, float conf_thr, float iou_thr, int top_k,
                                    float* ms_out, int capacity, void* stream) {
  if (!d || !ms_out || iters < 1) return fail(1, "bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int n_ops = static_cast<int>(d->ops.size());
  const int n = n_ops + 4;
  if (capacity < n) return fail(1, "ms_out capacity %d < %d", capacity, n);
  std::vector<cudaEvent_t> e0(n), e1(n);
  for (auto& e : e0) CUDA_OK(cudaEventCreate(&e));
  for (auto& e : e1) CUDA_OK(cudaEventCreate(&e));
  std::vector<double> acc(n, 0.0);
  int rc = 0;
  auto timed_ops = [&](int begin, int end) {
    for (int i = begin; i < end && !rc; ++i) {
      cudaEventRecord(e0[i], s);
      rc = launch_op(d, d->ops[i], d->input, s);
      cudaEventRecord(e1[i], s);
    }
  };
  for (int it = 0; it < iters && !rc; ++it) {
    timed_ops(0, d->n_dense_ops);
    if (rc) break;
    cudaEventRecord(e0[n_ops], s);
    rc = box_decode_launch(d->lv, d->boxes, d->scores, d->B, d->A, s);
    if (d->ovr_boxes && d->ovr_scores) {
      cudaMemcpyAsync(d->boxes, d->ovr_boxes, sizeof(float) * 4 * d->B * d->A, cudaMemcpyDeviceToDevice, s);
      cudaMemcpyAsync(d->scores, d->ovr_scores, sizeof(float) * d->B * d->A, cudaMemcpyDeviceToDevice, s);
    }
    cudaEventRecord(e1[n_ops], s);
    cudaEventRecord(e0[n_ops + 1], s);
    if (!rc) rc = select_nms_launch(d->boxes, d->scores, d->B, d->A, conf_thr, iou_thr, top_k, d->keep_k, d->keep_idx, d->keep_cnt,
                                    d->keep_boxes, d->keep_scores, s, g_err, sizeof(g_err));
    cudaEventRecord(e1[n_ops + 1], s);
    cudaEventRecord(e0[n_ops + 2], s);  // "gather" = head offsets (+ patch tables) ... survivor rows, without the patch ops
    if (!rc) rc = head_offsets_launch(d->keep_cnt, d->B, d->offsets, d->offsets + d->B, s);
    if (!rc && d->sparse)
      rc = patch_assign_launch(d->lv, d->keep_idx, d->keep_cnt, d->offsets, d->B, d->keep_k, d->patch_cap, d->head_level, d->head_patch,
                               d->patch_src, d->level_rows, s);
    cudaEventRecord(e1[n_ops + 2], s);
    if (!rc) timed_ops(d->n_dense_ops, n_ops);
    cudaEvent_t g0, g1;
    CUDA_OK(cudaEventCreate(&g0));
    CUDA_OK(cudaEventCreate(&g1));
    cudaEventRecord(g0, s);
    if (!rc) rc = flame_gather_launch(d->lv, d->keep_idx, d->keep_cnt, d->B, d->keep_k, d->img_xform, d->offsets, d->params,
                                      d->head_xform, d->head_img, s);
    cudaEventRecord(g1, s);
    cudaEventRecord(e0[n_ops + 3], s);
    if (!rc && d->flame)
      rc = flame_decode_launch(d->flame->model, d->params, d->B * d->keep_k, d->offsets + d->B, 128, 64, d->head_xform, nullptr,
                               d->rot, d->verts, s, g_err, sizeof(g_err));
    cudaEventRecord(e1[n_ops + 3], s);
    CUDA_OK(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0[i], e1[i]);
      acc[i] += ms;
    }
    float gms = 0.f;
    cudaEventElapsedTime(&gms, g0, g1);
    acc[n_ops + 2] += gms;
    cudaEventDestroy(g0);
    cudaEventDestroy(g1);
  }
  for (auto& e : e0) cudaEventDestroy(e);
  for (auto& e : e1) cudaEventDestroy(e);
  if (rc) return rc > 0 && g_err[0] ? rc : fail(5, "profile launch failed");
  for (int i = 0; i < n; ++i) ms_out[i] = static_cast<float>(acc[i] / iters);
  return 0;
}

extern "C" int vgh_detector_forward(vgh_detector* d, const uint8_t* images_dev, void* stream) {
  if (!d || !images_dev) return fail(1, "null argument");
  return run_forward(d, images_dev, static_cast<cudaStream_t>(stream), nullptr);
}
extern "C" int vgh_detector_postprocess(vgh_detector* d, float conf_thr, float iou_thr, int top_k,
                                        const float* img_xform_dev, void* stream) {
  if (!d) return fail(1, "null argument");
  return run_post(d, conf_thr, iou_thr, top_k, img_xform_dev ? img_xform_dev : d->img_xform,
                  static_cast<cudaStream_t>(stream), nullptr);
}
extern "C" int vgh_detector_dense_flame(vgh_detector* d, float* flame_dev, void* stream) {
  if (!d || !flame_dev) return fail(1, "null argument");
  if (d->sparse) return fail(2, "the dense [B,A,413] tensor does not exist in a sparse-heads plan (FLAME branch runs on survivors only)");
  return flame_dense_launch(d->lv, flame_dev, d->B, d->A, static_cast<cudaStream_t>(stream)) ? fail(5, "dense flame launch failed") : 0;
}
extern "C" void* vgh_detector_output(vgh_detector* d, int which) {
  if (!d) return nullptr;
  switch (which) {
    case VGH_OUT_BOXES: return d->boxes;
    case VGH_OUT_SCORES: return d->scores;
    case VGH_OUT_KEEP_IDX: return d->keep_idx;
    case VGH_OUT_KEEP_CNT: return d->keep_cnt;
    case VGH_OUT_KEEP_BOXES: return d->keep_boxes;
    case VGH_OUT_KEEP_SCORES: return d->keep_scores;
    case VGH_OUT_HEAD_OFFSETS: return d->offsets;
    case VGH_OUT_HEAD_PARAMS: return d->params;
    case VGH_OUT_HEAD_VERTS: return d->verts;
    case VGH_OUT_HEAD_ROT: return d->rot;
    case VGH_OUT_INPUT: return d->input;
  }
  return nullptr;
}
extern "C" int vgh_detector_num_anchors(const vgh_detector* d) { return d ? d->A : 0; }
extern "C" int vgh_detector_launch_count(const vgh_detector* d) { return d ? d->launches : 0; }
extern "C" int vgh_detector_read_buffer(vgh_detector* d, int buf, void* host_dst, size_t bytes) {
  if (!d || buf < 0 || buf >= (int)d->buf_ptr.size() || !host_dst) return fail(1, "bad buffer id");
  if (bytes > d->buf_bytes[buf]) return fail(1, "read of %zu bytes exceeds buffer (%zu)", bytes, d->buf_bytes[buf]);
  CUDA_OK(cudaMemcpy(host_dst, d->buf_ptr[buf], bytes, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int vgh_detector_write_buffer(vgh_detector* d, int buf, const void* host_src, size_t bytes) {
  if (!d || buf < 0 || buf >= (int)d->buf_ptr.size() || !host_src) return fail(1, "bad buffer id");
  if (bytes > d->buf_bytes[buf]) return fail(1, "write of %zu bytes exceeds buffer (%zu)", bytes, d->buf_bytes[buf]);
  CUDA_OK(cudaMemcpy(d->buf_ptr[buf], host_src, bytes, cudaMemcpyHostToDevice));
  return 0;
}
// Stage-wise parity aid: runs dense plan ops [first_op, n_dense_ops) over whatever the activation buffers hold (e.g.
// feature maps written with vgh_detector_write_buffer), then box decode.  first_op == n_dense_ops: decode only.
extern "C" int vgh_detector_forward_from(vgh_detector* d, int first_op, void* stream) {
  if (!d || first_op < 1 || first_op > d->n_dense_ops) return fail(1, "first_op must be in [1, n_dense_ops]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool ml = d->multi_lane;
  d->multi_lane = false;  // a partial plan has no complete dependency history: run it on one stream
  int rc = run_ops(d, first_op, d->n_dense_ops, nullptr, s, nullptr);
  d->multi_lane = ml;
  if (rc) return rc;
  if (box_decode_launch(d->lv, d->boxes, d->scores, d->B, d->A, s)) return fail(5, "box decode launch failed");
  return 0;
}
// Synthetic-workload hook (bench / tests): after box decode, overwrite boxes/scores with caller data
// (random weights never produce detections).  Pass NULLs to disable.
extern "C" int vgh_detector_set_override(vgh_detector* d, const float* boxes_dev, const float* scores_dev) {
  if (!d) return fail(1, "null argument");
  d->ovr_boxes = boxes_dev;
  d->ovr_scores = scores_dev;
  if (d->graph) { cudaGraphExecDestroy(d->graph); d->graph = nullptr; }
  return 0;
}

static int ensure_graph(vgh_detector* d, float conf, float iou, int top_k, cudaStream_t user) {
  if (d->graph && d->g_conf == conf && d->g_iou == iou && d->g_topk == top_k) return 0;
  if (d->graph) { cudaGraphExecDestroy(d->graph); d->graph = nullptr; }
  if (!d->cap_stream) CUDA_OK(cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking));
  cudaStream_t s = d->cap_stream;
  CUDA_OK(cudaStreamSynchronize(user));  // order the (re)build after whatever the caller queued
  // one eager pass first: sets the max-dynamic-smem attributes outside of capture and surfaces errors
  int launches = 0;
  int rc = run_forward(d, d->input, s, &launches);
  if (!rc) rc = run_post(d, conf, iou, top_k, d->img_xform, s, &launches);
  if (rc) return rc;
  CUDA_OK(cudaStreamSynchronize(s));
  d->launches = launches;
  cudaGraph_t g = nullptr;
  CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  rc = run_forward(d, d->input, s, nullptr);
  if (!rc) rc = run_post(d, conf, iou, top_k, d->img_xform, s, nullptr);
  cudaError_t e = cudaStreamEndCapture(s, &g);
  if (rc) { if (g) cudaGraphDestroy(g); return rc; }
  if (e != cudaSuccess) return fail(6, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&d->graph, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(6, "graph instantiate failed: %s", cudaGetErrorString(e));
  d->g_conf = conf; d->g_iou = iou; d->g_topk = top_k;
  return 0;
}

extern "C" int vgh_detector_run_device(vgh_detector* d, float conf_thr, float iou_thr, int top_k, void* stream) {
  if (!d) return fail(1, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = ensure_graph(d, conf_thr, iou_thr, top_k, s);
  if (rc) return rc;
  CUDA_OK(cudaGraphLaunch(d->graph, s));
  return 0;
}

extern "C" int vgh_detector_run_host(vgh_detector* d, const uint8_t* images_host, const float* img_xform_host,
                                     float conf_thr, float iou_thr, int top_k, int32_t* keep_cnt_host,
                                     float* keep_boxes_host, float* keep_scores_host, float* params_host,
                                     float* verts_host, int max_heads, int32_t* total_heads, void* stream) {
  if (!d || !images_host || !keep_cnt_host || !total_heads) return fail(1, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = ensure_graph(d, conf_thr, iou_thr, top_k, s);
  if (rc) return rc;
  const size_t B = d->B, K = d->keep_k;
  CUDA_OK(cudaMemcpyAsync(d->input, images_host, B * d->S * d->S * 3, cudaMemcpyHostToDevice, s));
  if (img_xform_host) CUDA_OK(cudaMemcpyAsync(d->img_xform, img_xform_host, B * 3 * 4, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaGraphLaunch(d->graph, s));
  CUDA_OK(cudaMemcpyAsync(keep_cnt_host, d->keep_cnt, B * 4, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaMemcpyAsync(total_heads, d->offsets + B, 4, cudaMemcpyDeviceToHost, s));
  if (keep_boxes_host) CUDA_OK(cudaMemcpyAsync(keep_boxes_host, d->keep_boxes, B * K * 16, cudaMemcpyDeviceToHost, s));
  if (keep_scores_host) CUDA_OK(cudaMemcpyAsync(keep_scores_host, d->keep_scores, B * K * 4, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  int n = *total_heads;
  if (n > max_heads) n = max_heads;
  if (n > 0) {
    if (params_host) CUDA_OK(cudaMemcpyAsync(params_host, d->params, (size_t)n * VGH_NUM_PARAMS * 4, cudaMemcpyDeviceToHost, s));
    if (verts_host) CUDA_OK(cudaMemcpyAsync(verts_host, d->verts, (size_t)n * VGH_NUM_VERTS * 12, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ host pipeline
// Two-deep software pipeline over HOST buffers: while step i computes, the images of step i+1 are
// uploaded and the results of step i-1 are downloaded (separate copy streams, double-buffered staging
// on the device).  Every step still performs its own H2D and D2H copies; they just overlap compute.
static int push_if_armed(vgh_detector* d, cudaStream_t s);
static int ensure_pipeline(vgh_detector* d) {
  if (d->pipe_stream) return 0;
  CUDA_OK(cudaStreamCreateWithFlags(&d->pipe_stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&d->h2d_stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&d->d2h_stream, cudaStreamNonBlocking));
  const size_t B = d->B, K = d->keep_k, cap = B * K;
  for (auto& sl : d->slots) {
    if (dmalloc(&sl.in_dev, B * d->S * d->S * 3) != cudaSuccess || dmalloc(&sl.params, cap * VGH_NUM_PARAMS) != cudaSuccess ||
        dmalloc(&sl.verts, cap * VGH_NUM_VERTS * 3) != cudaSuccess || dmalloc(&sl.kboxes, cap * 4) != cudaSuccess ||
        dmalloc(&sl.kscores, cap) != cudaSuccess || dmalloc(&sl.cnt, B) != cudaSuccess || dmalloc(&sl.total, 1) != cudaSuccess)
      return fail(4, "pipeline staging allocation failed");
    for (cudaEvent_t* e : {&sl.h2d_done, &sl.in_consumed, &sl.compute_done, &sl.d2h_done})
      CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  return 0;
}

extern "C" int vgh_detector_submit_host(vgh_detector* d, const uint8_t* images_host, float conf_thr, float iou_thr, int top_k) {
  if (!d || !images_host) return fail(1, "null argument");
  int rc = ensure_pipeline(d);
  if (rc) return rc;
  cudaStream_t s = d->pipe_stream;
  rc = ensure_graph(d, conf_thr, iou_thr, top_k, s);
  if (rc) return rc;
  vgh_detector::Slot& sl = d->slots[d->submit_idx & 1];
  if (sl.busy) return fail(8, "pipeline full: collect a result before submitting a third step");
  const size_t B = d->B, K = d->keep_k;
  // upload (copy stream); the slot's device image was last read by the D2D of its previous use
  CUDA_OK(cudaStreamWaitEvent(d->h2d_stream, sl.in_consumed, 0));
  CUDA_OK(cudaMemcpyAsync(sl.in_dev, images_host, B * d->S * d->S * 3, cudaMemcpyHostToDevice, d->h2d_stream));
  CUDA_OK(cudaEventRecord(sl.h2d_done, d->h2d_stream));
  // compute
  CUDA_OK(cudaStreamWaitEvent(s, sl.h2d_done, 0));
  CUDA_OK(cudaMemcpyAsync(d->input, sl.in_dev, B * d->S * d->S * 3, cudaMemcpyDeviceToDevice, s));
  CUDA_OK(cudaEventRecord(sl.in_consumed, s));
  CUDA_OK(cudaGraphLaunch(d->graph, s));
  // stage the results of this step so that the next step may overwrite the live buffers
  CUDA_OK(cudaStreamWaitEvent(s, sl.d2h_done, 0));
  CUDA_OK(cudaMemcpyAsync(sl.cnt, d->keep_cnt, B * 4, cudaMemcpyDeviceToDevice, s));
  CUDA_OK(cudaMemcpyAsync(sl.total, d->offsets + B, 4, cudaMemcpyDeviceToDevice, s));
  CUDA_OK(cudaMemcpyAsync(sl.kboxes, d->keep_boxes, B * K * 16, cudaMemcpyDeviceToDevice, s));
  CUDA_OK(cudaMemcpyAsync(sl.kscores, d->keep_scores, B * K * 4, cudaMemcpyDeviceToDevice, s));
  if (copy_rows_launch(d->params, sl.params, d->offsets + B, VGH_NUM_PARAMS, (int)(B * K), s) ||
      copy_rows_launch(d->verts, sl.verts, d->offsets + B, VGH_NUM_VERTS * 3, (int)(B * K), s))
    return fail(5, "result staging launch failed");
  rc = push_if_armed(d, s);  // multi-GPU: this step's record goes to rank 0 from the live buffers (before the next replay)
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(sl.compute_done, s));
  sl.busy = true;
  ++d->submit_idx;
  return 0;
}

extern "C" int vgh_detector_collect_host(vgh_detector* d, int32_t* keep_cnt_host, float* keep_boxes_host,
                                         float* keep_scores_host, float* params_host, float* verts_host, int max_heads,
                                         int32_t* total_heads) {
  if (!d || !keep_cnt_host || !total_heads) return fail(1, "null argument");
  if (!d->pipe_stream) return fail(8, "nothing submitted");
  vgh_detector::Slot& sl = d->slots[d->collect_idx & 1];
  if (!sl.busy) return fail(8, "nothing to collect");
  cudaStream_t c = d->d2h_stream;
  const size_t B = d->B, K = d->keep_k;
  CUDA_OK(cudaStreamWaitEvent(c, sl.compute_done, 0));
  CUDA_OK(cudaMemcpyAsync(keep_cnt_host, sl.cnt, B * 4, cudaMemcpyDeviceToHost, c));
  CUDA_OK(cudaMemcpyAsync(total_heads, sl.total, 4, cudaMemcpyDeviceToHost, c));
  if (keep_boxes_host) CUDA_OK(cudaMemcpyAsync(keep_boxes_host, sl.kboxes, B * K * 16, cudaMemcpyDeviceToHost, c));
  if (keep_scores_host) CUDA_OK(cudaMemcpyAsync(keep_scores_host, sl.kscores, B * K * 4, cudaMemcpyDeviceToHost, c));
  CUDA_OK(cudaStreamSynchronize(c));
  int n = *total_heads;
  if (n > max_heads) n = max_heads;
  if (n > 0) {
    if (params_host) CUDA_OK(cudaMemcpyAsync(params_host, sl.params, (size_t)n * VGH_NUM_PARAMS * 4, cudaMemcpyDeviceToHost, c));
    if (verts_host) CUDA_OK(cudaMemcpyAsync(verts_host, sl.verts, (size_t)n * VGH_NUM_VERTS * 12, cudaMemcpyDeviceToHost, c));
  }
  CUDA_OK(cudaEventRecord(sl.d2h_done, c));
  CUDA_OK(cudaStreamSynchronize(c));
  sl.busy = false;
  ++d->collect_idx;
  return 0;
}

// ------------------------------------------------------------------------------------------ multi-GPU record push
static int push_timeout_ms() {
  const char* e = getenv("VGGHEADS_B200_PUSH_TIMEOUT_MS");
  const int v = e ? atoi(e) : 0;
  return v > 0 ? v : 4000;
}
// Packs the live results into the armed destination on stream s (after the graph replay queued there) and disarms.
static int push_if_armed(vgh_detector* d, cudaStream_t s) {
  if (!d->push.armed) return 0;
  d->push.armed = false;
  RecordSrc src{d->keep_cnt, d->offsets + d->B, d->keep_boxes, d->keep_scores, d->params, d->verts};
  if (record_push_launch(src, record_layout(d->B, d->keep_k), d->push_seq++, d->push.dst, d->push.wait_flag, d->push.wait_val,
                         d->push.done_flag, d->push.done_val, d->push_counter, d->push_status, push_timeout_ms(), s))
    return fail(5, "record push launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

extern "C" int vgh_detector_record_layout(const vgh_detector* d, int64_t* out4) {
  if (!d || !out4) return fail(1, "null argument");
  const RecordLayout l = record_layout(d->B, d->keep_k);
  out4[0] = l.fixed_words; out4[1] = VGH_NUM_PARAMS; out4[2] = VGH_NUM_VERTS * 3; out4[3] = l.capacity_words;
  return 0;
}
extern "C" int vgh_detector_arm_push(vgh_detector* d, float* dst_record_dev, const uint64_t* wait_flag_dev, uint64_t wait_val,
                                     uint64_t* done_flag_dev, uint64_t done_val) {
  if (!d || !dst_record_dev) return fail(1, "null argument");
  if (reinterpret_cast<uintptr_t>(dst_record_dev) % 16) return fail(1, "record destination must be 16-byte aligned");
  d->push.armed = true;
  d->push.dst = dst_record_dev;
  d->push.wait_flag = reinterpret_cast<const unsigned long long*>(wait_flag_dev);
  d->push.wait_val = wait_val;
  d->push.done_flag = reinterpret_cast<unsigned long long*>(done_flag_dev);
  d->push.done_val = done_val;
  return 0;
}
extern "C" int vgh_detector_push_status(vgh_detector* d, int32_t* status_host) {
  if (!d || !status_host) return fail(1, "null argument");
  CUDA_OK(cudaMemcpy(status_host, d->push_status, 4, cudaMemcpyDeviceToHost));
  return 0;
}
// Device-resident step for multi-GPU runs: graph replay over the staging input on `stream`, then - if a push is armed -
// ONE kernel packs the step's result record straight into its destination (local snapshot or rank 0's peer ring).
extern "C" int vgh_detector_submit_device(vgh_detector* d, float conf_thr, float iou_thr, int top_k, void* stream) {
  if (!d) return fail(1, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = ensure_graph(d, conf_thr, iou_thr, top_k, s);
  if (rc) return rc;
  CUDA_OK(cudaGraphLaunch(d->graph, s));
  return push_if_armed(d, s);
}

// ------------------------------------------------------------------------------------------ peer memory (CUDA IPC)
extern "C" int vgh_peer_alloc(size_t bytes, void** dev_ptr, uint8_t* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) return fail(1, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return fail(100, "peer alloc: %s", cudaGetErrorString(e)); }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}
extern "C" int vgh_peer_free(void* dev_ptr) {
  if (dev_ptr) CUDA_OK(cudaFree(dev_ptr));
  return 0;
}
extern "C" int vgh_peer_open(const uint8_t* handle64, void** dev_ptr) {
  if (!handle64 || !dev_ptr) return fail(1, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(101, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); }
  return 0;
}
extern "C" int vgh_peer_close(void* dev_ptr) {
  if (dev_ptr) CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}
extern "C" int vgh_gather_wait(const uint64_t* ready_dev, int n, int stride, uint64_t value, const float* records_dev,
                               int64_t record_stride_words, uint64_t* const* ack_ptrs_dev, uint64_t ack_val, int32_t* total_out_dev,
                               int32_t* status_dev, int timeout_ms, void* stream) {
  if (!ready_dev || n < 1 || n > 32) return fail(1, "bad argument");
  if (gather_wait_launch(reinterpret_cast<const unsigned long long*>(ready_dev), n, stride, value, records_dev, record_stride_words,
                         reinterpret_cast<unsigned long long* const*>(ack_ptrs_dev), ack_val, total_out_dev, status_dev,
                         timeout_ms > 0 ? timeout_ms : push_timeout_ms(), static_cast<cudaStream_t>(stream)))
    return fail(5, "gather wait launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
