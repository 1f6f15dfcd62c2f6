// Fused FLAME decode for sm_100a: 413-float head parameters -> 5023x3 vertices.
//
// Replaces the reference's ~25 library launches per call (head_detector/flame.py:122-169 ->
// smplx.lbs.lbs; flame.py:179-208; utils.py:120-128; detector.py:66-69) with ONE kernel:
//   blendshapes (shape+expression) -> jaw pose correctives -> joint regression -> LBS skinning
//   -> +0.05 z -> 6D rotation -> scale -> translate -> un-letterbox.
//
// Precision contract (DESIGN.md "FLAME decode"): everything up to the model-space vertex is
// accumulated in fp64 (the reference's own fp32 rounding is the only difference left), the vertex
// is rounded once to fp32, and the remaining ops (R*v, *scale, +t, -pad, /scale) are done in fp32
// in the reference's operation order with non-contracted intrinsics so that roundings coincide.
//
// Work decomposition: item = 128 vertices x (8*HG heads), persistent CTAs walk the items.  The blendshape sum is a small
// fp64 GEMM per item - [8*HG heads x K] (betas) x [K x 384 vertex coordinates] (basis slab), K = live coefficients + 9 pose
// rows - and runs on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4), heads on M, vertex coordinates on N.
// Round 1/2 used one DFMA per (head, coordinate, coefficient): 24 DFMA + 7 LDS + 3 F2F per thread and coefficient issue at
// most 44 FMA/clk/SM (profiles/r2b_fp64_rate_probe.txt) and the kernel reached 25; DMMA sustains 62-66 FMA/clk/SM (= the
// B200's 37 TFLOP/s FP64 peak) with a fifth of the instructions.  Same arithmetic class: fp64 products and sums.
// Each warp owns 48/(4*HG) n-tiles x HG m-tiles = 12 accumulator tiles (24 doubles per thread) and streams ITS OWN slice of
// the basis slab (its 48 or 96 coordinates x 8 coefficients per stage) through a private 3-stage cp.async pipeline: the B
// operand is not shared between warps, so the main loop needs no CTA-wide barrier at all (cp.async.wait_group + __syncwarp)
// and the warps of the three resident CTAs drift apart instead of hitting the FP64 tensor pipe in lock step.  The basis
// stays fp32 - the precision the reference holds it in - and is widened exactly to fp64 when a B fragment is loaded; so are
// the betas (fp32 parameters).  Row pitches are padded so that the fragment loads (4 rows x 8 consecutive elements per warp)
// are bank-conflict free.  For the per-vertex epilogue (skinning needs x, y, z of a vertex in one thread) the accumulators
// cross shared memory once per m-tile.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "flame_decode.cuh"
#include "device_attr.cuh"

namespace vgh {

constexpr int kV = 5023;
constexpr int kVPad = 5120;  // 40 tiles of 128 vertices; padded rows are zero
constexpr int kL = 400;
constexpr int kTileV = 128;
constexpr int kHPT = 8;   // heads per thread
constexpr int kLc = 8;    // coefficients per pipeline stage
constexpr int kNS = 3;    // pipeline stages of the basis stream (3 x 12 KB: three CTAs per SM fit, ncu: the kernel is latency-, not bandwidth-bound)
constexpr int kNPose = 9; // live pose-corrective rows (jaw joint only)
constexpr int kNC = kTileV * 3;       // vertex coordinates per item = GEMM N
constexpr int kXPitch = kNC + 8;      // fp64 row pitch of the accumulator exchange tile: rows 16 banks apart
constexpr int kParams = 413;

struct FlameDev {
  float* sdt;    // [400][kVPad][3]  shape basis, coefficient-major (fp32 as the reference holds it; widened
                 //                  exactly to fp64 in registers - half the L2 / shared-memory traffic of an fp64 copy)
  float* pd;     // [9][kVPad][3]    posedirs rows 9..17 (jaw)
  double* vt;    // [kVPad][3]       template
  double* wI;    // [kVPad]          W0+W1+W3+W4
  double* w2;    // [kVPad]          W2 (jaw)
  double* js2;   // [3][400]         J_regressor[2] . shapedirs
  double j2t[3]; //                  J_regressor[2] . v_template
};

struct FlameArgs {
  FlameDev c;
  const float* params;  // [N,413]
  const float* xform;   // [N,3] (pad_x, pad_y, img_scale) or null
  const int* n_dev;     // optional device-side head count (overrides n when non-null)
  float* verts;         // optional [N,5023,3] model space (incl. +0.05 z)
  float* rot;           // optional [N,9]
  float* proj;          // [N,5023,3]
  int n;
  int ns, ne;           // live shape / expression coefficient counts
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// shared-memory geometry, used by the kernel and by the host-side size computation
__host__ __device__ constexpr int flame_beta_pitch(int heads) { return heads == 8 ? 8 : heads + 8; }
__host__ __device__ constexpr int flame_stage_floats(int heads) {   // all warps' private basis stages
  return (heads / 2) * kNS * kLc * ((kNC / 8) / (heads / 2) * 8 + 8);
}

// D[8x8] += A[8x4] * B[4x8] in fp64 (DMMA.8x8x4).  Lane l holds a[l>>2][l&3], b[l&3][l>>2], d[l>>2][2*(l&3) + {0,1}].
__device__ __forceinline__ void dmma_884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// rot6d -> rotation matrix, reference op order (utils.py:120-128; F.normalize eps 1e-12), fp32.
__device__ void rot6d_to_mat(const float* v, float* R /*row-major 3x3*/) {
  float n1 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
  n1 = fmaxf(n1, 1e-12f);
  float b1[3] = {__fdiv_rn(v[0], n1), __fdiv_rn(v[1], n1), __fdiv_rn(v[2], n1)};
  const float* vy = v + 3;
  float c[3] = {__fsub_rn(__fmul_rn(b1[1], vy[2]), __fmul_rn(b1[2], vy[1])),
                __fsub_rn(__fmul_rn(b1[2], vy[0]), __fmul_rn(b1[0], vy[2])),
                __fsub_rn(__fmul_rn(b1[0], vy[1]), __fmul_rn(b1[1], vy[0]))};
  float n3 = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(c[0], c[0]), __fmul_rn(c[1], c[1])), __fmul_rn(c[2], c[2])));
  n3 = fmaxf(n3, 1e-12f);
  float b3[3] = {__fdiv_rn(c[0], n3), __fdiv_rn(c[1], n3), __fdiv_rn(c[2], n3)};
  float b2[3] = {-__fsub_rn(__fmul_rn(b1[1], b3[2]), __fmul_rn(b1[2], b3[1])),
                 -__fsub_rn(__fmul_rn(b1[2], b3[0]), __fmul_rn(b1[0], b3[2])),
                 -__fsub_rn(__fmul_rn(b1[0], b3[1]), __fmul_rn(b1[1], b3[0]))};
  // columns are (b1, b2, b3)
  for (int i = 0; i < 3; ++i) {
    R[i * 3 + 0] = b1[i];
    R[i * 3 + 1] = b2[i];
    R[i * 3 + 2] = b3[i];
  }
}

// Rodrigues as smplx.lbs.batch_rodrigues does it in fp32: angle = |r + 1e-8|, axis = r / angle.
__device__ void rodrigues_f32(const float* r, float* R) {
  float a0 = __fadd_rn(r[0], 1e-8f), a1 = __fadd_rn(r[1], 1e-8f), a2 = __fadd_rn(r[2], 1e-8f);
  float ang = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)), __fmul_rn(a2, a2)));
  float kx = __fdiv_rn(r[0], ang), ky = __fdiv_rn(r[1], ang), kz = __fdiv_rn(r[2], ang);
  float s = sinf(ang), c = cosf(ang);
  float K[9] = {0.f, -kz, ky, kz, 0.f, -kx, -ky, kx, 0.f};
  float KK[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float acc = __fmul_rn(K[i * 3 + 0], K[0 * 3 + j]);
      acc = __fmaf_rn(K[i * 3 + 1], K[1 * 3 + j], acc);
      acc = __fmaf_rn(K[i * 3 + 2], K[2 * 3 + j], acc);
      KK[i * 3 + j] = acc;
    }
  float omc = __fsub_rn(1.f, c);
  for (int i = 0; i < 9; ++i) {
    float id = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
    R[i] = __fadd_rn(__fadd_rn(id, __fmul_rn(s, K[i])), __fmul_rn(omc, KK[i]));
  }
}

template <int HG>
__global__ void __launch_bounds__(kTileV* HG, HG == 1 ? 4 : 3) flame_decode_kernel(const FlameArgs a) {
  constexpr int kHeads = kHPT * HG;       // heads per item = HG m-tiles of 8
  constexpr int kBP = flame_beta_pitch(kHeads);   // fp32 row pitch of the beta table: rows 8 banks apart -> conflict-free A fragments
  constexpr int kWarps = 4 * HG;
  constexpr int kNT = (kNC / 8) / kWarps; // n-tiles (8 coordinates) per warp: 6 (HG = 2) or 12 (HG = 1)
  constexpr int kWP = kNT * 8 + 8;        // fp32 row pitch of a warp's basis stage: rows 8 / 24 banks apart -> conflict-free B fragments
  constexpr int kWStage = kLc * kWP;      // floats per stage of one warp
  static_assert(kNT * kWarps * 8 == kNC, "n-tiles must split evenly over the warps");
  static_assert(kWP % 32 == 8 || kWP % 32 == 24, "B-fragment rows must be 8 banks apart");
  static_assert(kBP % 32 == 8 || kBP % 32 == 24, "A-fragment rows must be 8 banks apart");
  static_assert(kWarps * kNS * kWStage == flame_stage_floats(kHeads), "host-side shared-memory size out of step");
  const int n_heads = a.n_dev ? min(*a.n_dev, a.n) : a.n;
  // work items = (vertex tile, head group); CTAs walk them with a grid stride so that the device-side
  // head count decides the amount of work, not the (capacity-sized) launch
  constexpr int kVTiles = kVPad / kTileV;
  const int n_items = kVTiles * ((n_heads + kHeads - 1) / kHeads);
  const int tid = threadIdx.x;
  const int vl = tid % kTileV;
  const int grp = tid / kTileV;
  const int warp = tid >> 5, lane = tid & 31;
  const int fr = lane >> 2, fc = lane & 3;   // fragment coordinates: A[fr][fc], B[fc][fr], D[fr][2*fc + {0,1}]
  const int lb = a.ns + a.ne;          // live blendshape coefficients
  const int lt = lb + kNPose;          // + pose-corrective rows
  const int n_chunks = (lt + kLc - 1) / kLc;
  const int lt_pad = n_chunks * kLc;

  extern __shared__ __align__(16) uint8_t smem[];
  float* sd_s = reinterpret_cast<float*>(smem);                         // [kWarps][kNS][kLc][kWP] fp32: private stages of every warp
  double* tj_s = reinterpret_cast<double*>(sd_s + kWarps * kNS * kWStage);    // [kHeads][3]
  double* r2_s = tj_s + kHeads * 3;                                     // [kHeads][9]
  float* R_s = reinterpret_cast<float*>(r2_s + kHeads * 9);             // [kHeads][9]
  float* st_s = R_s + kHeads * 9;                                       // [kHeads][8]: scale,tx,ty,tz,padx,pady,iscale
  float* beta_s = st_s + kHeads * 8;                                    // [lt_pad][kBP] fp32 (the parameters' own precision)
  // the accumulator exchange tile [8 heads][kXPitch] of the epilogue borrows the (by then drained) basis stage buffers
  double* x_s = reinterpret_cast<double*>(sd_s);
  static_assert(8 * kXPitch * sizeof(double) <= kWarps * kNS * kWStage * sizeof(float), "the exchange tile must fit the stage buffers");
  // items are ordered head-group-major and every CTA takes a CONTIGUOUS range of them, so that the per-head-group
  // prologue (betas to fp64, Rodrigues, 6D rotation, jaw-joint regression: a 192..400-term dependent chain) runs once per
  // head group a CTA touches instead of once per (vertex tile, head group)
  const int per_cta = (n_items + gridDim.x - 1) / gridDim.x;
  const int item_begin = blockIdx.x * per_cta, item_end = min(n_items, item_begin + per_cta);
  int prologue_head0 = -1;
  for (int item = item_begin; item < item_end; ++item) {
  const int head0 = (item / kVTiles) * kHeads;
  const int v0 = (item % kVTiles) * kTileV;
  const bool fresh = head0 != prologue_head0;   // uniform across the CTA
  prologue_head0 = head0;

  // ---- the basis stream of this item starts first: its latency overlaps the prologue
  const int n_warp0 = warp * (kNT * 8);   // first coordinate of this warp's n-tiles
  float* wst = sd_s + warp * (kNS * kWStage);          // this warp's stages
  const uint32_t wst_smem = static_cast<uint32_t>(__cvta_generic_to_shared(wst));
  constexpr int kVecPerRow = kNT * 8 * 4 / 16;          // 16-byte vectors of one coefficient row of the warp's slice (12 | 24)
  constexpr int kVecPerStage = kLc * kVecPerRow;        // 96 | 192: 3 | 6 per lane
  auto issue = [&](int chunk) {
    if (chunk < n_chunks) {
      const int buf = chunk % kNS;
#pragma unroll
      for (int q = lane; q < kVecPerStage; q += 32) {
        const int row = q / kVecPerRow;
        const int off = q - row * kVecPerRow;
        const int i = chunk * kLc + row;
        const float* src;
        if (i < lb) {
          const int l = i < a.ns ? i : 300 + (i - a.ns);
          src = a.c.sdt + (static_cast<size_t>(l) * kVPad + v0) * 3;
        } else {
          const int m = min(i - lb, kNPose - 1);  // rows past the end multiply a zero beta
          src = a.c.pd + (static_cast<size_t>(m) * kVPad + v0) * 3;
        }
        cp_async16(wst_smem + (buf * kWStage + row * kWP) * 4 + off * 16, src + n_warp0 + off * 4);
      }
    }
    cp_async_commit();  // (possibly empty) one group per chunk keeps the wait arithmetic uniform
  };
  for (int c = 0; c < kNS - 1; ++c) issue(c);   // (the previous item's reads of the exchange tile are behind its closing barrier)

  // ---- prologue: betas, per-head rotations
  if (fresh) {
  // (every load below is independent of the previous one and unrolled, so a phase costs ~one L2 round trip: at kernel start
  // all resident CTAs run this prologue at the same time and nothing else can hide it)
  for (int i = tid; i < lb; i += blockDim.x) {
    const int l = i < a.ns ? i : 300 + (i - a.ns);
    float v[kHeads];
#pragma unroll
    for (int h = 0; h < kHeads; ++h) v[h] = head0 + h < n_heads ? __ldg(a.params + static_cast<size_t>(head0 + h) * kParams + l) : 0.f;
#pragma unroll
    for (int h = 0; h < kHeads; ++h) beta_s[i * kBP + h] = v[h];
  }
  for (int idx = tid; idx < (lt_pad - lt) * kHeads; idx += blockDim.x) beta_s[(lt + idx / kHeads) * kBP + idx % kHeads] = 0.f;
  if (tid < kHeads) {
    const int hg = head0 + tid;
    float R2[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float sc = 1.f, t[3] = {0, 0, 0}, xf[3] = {0.f, 0.f, 1.f};
    if (hg < n_heads) {
      float q[13];   // jaw (3), rot6d (6), translation (3), scale
#pragma unroll
      for (int i = 0; i < 13; ++i) q[i] = __ldg(a.params + static_cast<size_t>(hg) * kParams + 400 + i);
      if (a.xform) { xf[0] = __ldg(a.xform + hg * 3); xf[1] = __ldg(a.xform + hg * 3 + 1); xf[2] = __ldg(a.xform + hg * 3 + 2); }
      rodrigues_f32(q, R2);
      rot6d_to_mat(q + 3, R);
      t[0] = q[9]; t[1] = q[10]; t[2] = q[11];
      sc = fmaxf(q[12], 1e-8f);
      if (a.rot)   // (every CTA that starts on this head group writes the same values)
        for (int i = 0; i < 9; ++i) a.rot[static_cast<size_t>(hg) * 9 + i] = R[i];
    }
    for (int i = 0; i < 9; ++i) {
      r2_s[tid * 9 + i] = static_cast<double>(R2[i]);
      R_s[tid * 9 + i] = R[i];
      // pose feature = (R2 - I) in fp32, as the reference forms it, acts as 9 extra "betas"
      const float id = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
      beta_s[(lb + i) * kBP + tid] = __fsub_rn(R2[i], id);
    }
    st_s[tid * 8 + 0] = sc; st_s[tid * 8 + 1] = t[0]; st_s[tid * 8 + 2] = t[1]; st_s[tid * 8 + 3] = t[2];
    st_s[tid * 8 + 4] = xf[0]; st_s[tid * 8 + 5] = xf[1]; st_s[tid * 8 + 6] = xf[2];
  }
  __syncthreads();
  // jaw joint J2 = J2_template + JS2 . beta ; tJ = J2 - R2 . J2.  The 3 * kHeads (head, coordinate) dot products of
  // 192..400 terms are spread over 16-lane groups, three outputs per group; the [3][400] regressor x basis table is read
  // straight from L2, twelve independent loads per lane in flight.
  {
    constexpr int kGroups = 4 * HG * 2;          // 16-lane groups of the CTA: 16 | 8 -> exactly three outputs each
    static_assert(3 * kGroups == 3 * kHeads, "three (head, coordinate) outputs per 16-lane group");
    const int g16 = tid >> 4, sub = tid & 15;
    double a3[3] = {0.0, 0.0, 0.0};
    for (int i0 = 0; i0 < lb; i0 += 64) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 16 + sub;
        const bool ok = i < lb;
        const int l = i < a.ns ? i : 300 + (i - a.ns);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int o = g16 + j * kGroups, h = o / 3, k = o - h * 3;
          const double js = ok ? __ldg(a.c.js2 + k * kL + l) : 0.0;
          const double b = ok ? static_cast<double>(beta_s[i * kBP + h]) : 0.0;
          a3[j] = fma(js, b, a3[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
      for (int sft = 1; sft < 16; sft <<= 1) a3[j] += __shfl_xor_sync(0xffffffffu, a3[j], sft);
      const int o = g16 + j * kGroups;
      if (sub == 0) tj_s[o] = a.c.j2t[o % 3] + a3[j];  // temporarily J2
    }
  }
  __syncthreads();
  if (tid < kHeads) {
    const double j0 = tj_s[tid * 3], j1 = tj_s[tid * 3 + 1], j2 = tj_s[tid * 3 + 2];
    const double* r = r2_s + tid * 9;
    const double o0 = j0 - (r[0] * j0 + r[1] * j1 + r[2] * j2);
    const double o1 = j1 - (r[3] * j0 + r[4] * j1 + r[5] * j2);
    const double o2 = j2 - (r[6] * j0 + r[7] * j1 + r[8] * j2);
    tj_s[tid * 3] = o0; tj_s[tid * 3 + 1] = o1; tj_s[tid * 3 + 2] = o2;
  }
  }  // fresh head group
  __syncthreads();   // tj_s / beta_s of a fresh head group are complete for every warp (the main loop has no CTA-wide barrier)

  // ---- main loop: acc[head][coord] = template[coord] + sum_i beta[head][i] * basis[i][coord], as D += A * B tiles
  double acc[HG][kNT][2];
#pragma unroll
  for (int j = 0; j < kNT; ++j) {
    const double2 t = *reinterpret_cast<const double2*>(a.c.vt + static_cast<size_t>(v0) * 3 + n_warp0 + j * 8 + 2 * fc);
#pragma unroll
    for (int m = 0; m < HG; ++m) { acc[m][j][0] = t.x; acc[m][j][1] = t.y; }
  }
  for (int c = 0; c < n_chunks; ++c) {
    cp_async_wait<kNS - 2>();  // this lane's copies of chunk c have landed (groups complete in order)
    __syncwarp();              // ... every lane's; and every lane is done reading the buffer of chunk c-1, which is refilled next
    issue(c + kNS - 1);
    const float* sd = wst + (c % kNS) * kWStage + fc * kWP + fr;    // B[fc][fr] of n-tile 0, k-step 0
    const float* bt = beta_s + (c * kLc + fc) * kBP + fr;           // A[fr][fc] of m-tile 0, k-step 0
#pragma unroll
    for (int ks = 0; ks < kLc / 4; ++ks) {
      double af[HG];
#pragma unroll
      for (int m = 0; m < HG; ++m) af[m] = static_cast<double>(bt[ks * 4 * kBP + m * 8]);
#pragma unroll
      for (int j = 0; j < kNT; ++j) {
        const double bf = static_cast<double>(sd[ks * 4 * kWP + j * 8]);
#pragma unroll
        for (int m = 0; m < HG; ++m) dmma_884(acc[m][j], af[m], bf);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // every warp is done with its stage buffers: together they become the exchange tile

  // ---- epilogue: skinning + rigid transform.  Per m-tile the accumulators cross shared memory ([8 heads][coordinate]) so
  // that thread (vertex vl, group grp) holds x, y, z of its vertex for 8 / HG heads
  const int v = v0 + vl;
  const double wi = a.c.wI[v], w2 = a.c.w2[v];   // (padded vertices: zeros)
#pragma unroll
  for (int m = 0; m < HG; ++m) {
    if (m) __syncthreads();   // the previous m-tile has been read
#pragma unroll
    for (int j = 0; j < kNT; ++j)
      *reinterpret_cast<double2*>(x_s + fr * kXPitch + n_warp0 + j * 8 + 2 * fc) = make_double2(acc[m][j][0], acc[m][j][1]);
    __syncthreads();
    if (v < kV) {
#pragma unroll
      for (int hs = 0; hs < kHPT / HG; ++hs) {
        const int hx = grp * (kHPT / HG) + hs;   // head inside the m-tile
        const int hl = m * 8 + hx;
        const int hg = head0 + hl;
        if (hg >= n_heads) break;
        const double* r2 = r2_s + hl * 9;
        const double x = x_s[hx * kXPitch + vl * 3], y = x_s[hx * kXPitch + vl * 3 + 1], z = x_s[hx * kXPitch + vl * 3 + 2];
        const double rx = r2[0] * x + r2[1] * y + r2[2] * z + tj_s[hl * 3];
        const double ry = r2[3] * x + r2[4] * y + r2[5] * z + tj_s[hl * 3 + 1];
        const double rz = r2[6] * x + r2[7] * y + r2[8] * z + tj_s[hl * 3 + 2];
        const float mx = static_cast<float>(wi * x + w2 * rx);
        const float my = static_cast<float>(wi * y + w2 * ry);
        const float mz = __fadd_rn(static_cast<float>(wi * z + w2 * rz), 0.05f);  // flame.py:164
        const size_t o = (static_cast<size_t>(hg) * kV + v) * 3;
        if (a.verts) { a.verts[o] = mx; a.verts[o + 1] = my; a.verts[o + 2] = mz; }
        const float* R = R_s + hl * 9;
        const float* st = st_s + hl * 8;
        float out[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float rv = __fmul_rn(R[i * 3], mx);
          rv = __fmaf_rn(R[i * 3 + 1], my, rv);
          rv = __fmaf_rn(R[i * 3 + 2], mz, rv);
          out[i] = __fadd_rn(__fmul_rn(rv, st[0]), st[1 + i]);  // flame.py:198-199
        }
        out[0] = __fsub_rn(out[0], st[4]);  // detector.py:67-69
        out[1] = __fsub_rn(out[1], st[5]);
        a.proj[o] = __fdiv_rn(out[0], st[6]);
        a.proj[o + 1] = __fdiv_rn(out[1], st[6]);
        a.proj[o + 2] = __fdiv_rn(out[2], st[6]);
      }
    }
  }
  __syncthreads();  // shared memory is reused by the next work item
  }  // item loop
}

// ------------------------------------------------------------------------------------------- host
struct FlameModel {
  FlameDev dev;
  void* blob = nullptr;
};

static size_t flame_smem_bytes(int heads, int lt_pad) {
  return sizeof(float) * flame_stage_floats(heads) + sizeof(double) * (heads * 3 + heads * 9) + sizeof(float) * (heads * 9 + heads * 8) +
         sizeof(float) * static_cast<size_t>(lt_pad) * flame_beta_pitch(heads);
}

int flame_model_create(const float* v_template, const float* shapedirs, const float* posedirs, const float* j_regressor,
                       const float* lbs_weights, FlameModel** out, char* err, size_t errlen) {
  // host re-layout: the two big tables stay fp32 (coefficient-major); the small per-vertex / per-joint
  // tables are widened to fp64 once (exact)
  const size_t n_sdt = static_cast<size_t>(kL) * kVPad * 3, n_pd = static_cast<size_t>(kNPose) * kVPad * 3;
  const size_t n_vt = static_cast<size_t>(kVPad) * 3, n_w = kVPad, n_js = 3 * kL;
  const size_t total_d = n_vt + 2 * n_w + n_js;
  std::vector<float> hostf(n_sdt + n_pd, 0.f);
  std::vector<double> host(total_d, 0.0);
  float* sdt = hostf.data();
  float* pd = sdt + n_sdt;
  double* vt = host.data();
  double* wI = vt + n_vt;
  double* w2 = wI + n_w;
  double* js2 = w2 + n_w;
  for (int v = 0; v < kV; ++v)
    for (int k = 0; k < 3; ++k) {
      const float* src = shapedirs + (static_cast<size_t>(v) * 3 + k) * kL;
      for (int l = 0; l < kL; ++l) sdt[(static_cast<size_t>(l) * kVPad + v) * 3 + k] = src[l];
      vt[v * 3 + k] = v_template[v * 3 + k];
      for (int m = 0; m < kNPose; ++m)  // posedirs is [36][15069]; jaw joint (index 2) -> rows 9..17
        pd[(static_cast<size_t>(m) * kVPad + v) * 3 + k] = posedirs[static_cast<size_t>(9 + m) * (kV * 3) + v * 3 + k];
    }
  for (int v = 0; v < kV; ++v) {
    const float* w = lbs_weights + v * 5;
    wI[v] = static_cast<double>(w[0]) + w[1] + w[3] + w[4];
    w2[v] = w[2];
  }
  FlameModel* m = new FlameModel();
  for (int k = 0; k < 3; ++k) {
    double jt = 0.0;
    for (int v = 0; v < kV; ++v) jt += static_cast<double>(j_regressor[2 * kV + v]) * v_template[v * 3 + k];
    m->dev.j2t[k] = jt;
    for (int l = 0; l < kL; ++l) {
      double s = 0.0;
      for (int v = 0; v < kV; ++v) {
        const float jr = j_regressor[2 * kV + v];
        if (jr != 0.f) s += static_cast<double>(jr) * shapedirs[(static_cast<size_t>(v) * 3 + k) * kL + l];
      }
      js2[k * kL + l] = s;
    }
  }
  const size_t bytes_d = total_d * sizeof(double), bytes_f = hostf.size() * sizeof(float);
  cudaError_t e = cudaMalloc(&m->blob, bytes_d + bytes_f);
  if (e == cudaSuccess) e = cudaMemcpy(m->blob, host.data(), bytes_d, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(static_cast<uint8_t*>(m->blob) + bytes_d, hostf.data(), bytes_f, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    snprintf(err, errlen, "flame_model_create: %s", cudaGetErrorString(e));
    if (m->blob) cudaFree(m->blob);
    delete m;
    return 1;
  }
  double* d = static_cast<double*>(m->blob);
  m->dev.vt = d; d += n_vt;
  m->dev.wI = d; d += n_w;
  m->dev.w2 = d; d += n_w;
  m->dev.js2 = d; d += n_js;
  float* f = reinterpret_cast<float*>(d);  // bytes_d is a multiple of 16: the fp32 tables stay 16-byte aligned
  m->dev.sdt = f;
  m->dev.pd = f + n_sdt;
  *out = m;
  return 0;
}

void flame_model_destroy(FlameModel* m) {
  if (!m) return;
  if (m->blob) cudaFree(m->blob);
  delete m;
}

template <int HG>
static int launch_flame(const FlameArgs& a, cudaStream_t stream, char* err, size_t errlen) {
  const int heads = kHPT * HG;
  const int lt = a.ns + a.ne + kNPose;
  const int lt_pad = (lt + kLc - 1) / kLc * kLc;
  const size_t smem = flame_smem_bytes(heads, lt_pad);
  static SmemOptIn opt_in;
  {
    cudaError_t e = ensure_dynamic_smem(flame_decode_kernel<HG>, opt_in, smem);
    if (e != cudaSuccess) {
      snprintf(err, errlen, "flame smem %zu: %s", smem, cudaGetErrorString(e));
      return 2;
    }
  }
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flame_decode_kernel<HG>, kTileV * HG, smem);
  if (per_sm < 1) per_sm = 1;
  const int max_items = (kVPad / kTileV) * ((a.n + heads - 1) / heads);
  int grid = sms * per_sm;
  if (grid > max_items) grid = max_items;
  flame_decode_kernel<HG><<<grid, kTileV * HG, smem, stream>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(err, errlen, "flame launch: %s", cudaGetErrorString(e));
    return 3;
  }
  return 0;
}

int flame_decode_launch(const FlameModel* m, const float* params, int n, const int* n_dev, int ns, int ne,
                        const float* xform, float* verts, float* rot, float* proj, cudaStream_t stream, char* err,
                        size_t errlen) {
  if (n <= 0) return 0;
  if (ns < 0 || ns > 300 || ne < 0 || ne > 100) {
    snprintf(err, errlen, "flame_decode: live coefficient counts out of range (%d,%d)", ns, ne);
    return 1;
  }
  FlameArgs a;
  a.c = m->dev;
  a.params = params; a.xform = xform; a.n_dev = n_dev; a.verts = verts; a.rot = rot; a.proj = proj;
  a.n = n; a.ns = ns; a.ne = ne;
  // 16-head items: small enough to balance over the SMs whatever the (device-side) head count is,
  // large enough to amortise the basis stream; tiny batches use 8-head items
  if (n > 8) return launch_flame<2>(a, stream, err, errlen);
  return launch_flame<1>(a, stream, err, errlen);
}

}  // namespace vgh
