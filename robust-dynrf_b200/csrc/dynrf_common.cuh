// Shared device helpers for the dynrf_b200 kernels (sm_100a).
//
// Numerics contract: coordinates that feed floor() or a high-frequency positional encoding
// are computed with explicitly rounded (non-contracted) fp32 ops in the same order as the
// reference's PyTorch expressions, so that the same texel / the same sin argument is selected.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dynrf_b200.h"

#define DYNRF_CHECK_LAUNCH()                          \
  do {                                                \
    cudaError_t _e = cudaGetLastError();              \
    if (_e != cudaSuccess) return (int)_e;            \
  } while (0)

// host-side: every plane must be channel-last with one of the two texel orders (see dynrf_vm_t.sv / .su)
static inline int dynrf_vm_layout_ok(const dynrf_vm_t* vm) {
  for (int i = 0; i < 3; ++i) {
    const bool xfast = vm->sv[i] == vm->W[i] && vm->su[i] == 1;
    const bool zfast = vm->sv[i] == 1 && vm->su[i] == vm->H[i];
    if (!xfast && !zfast && !(vm->H[i] == 1 || vm->W[i] == 1)) return 0;
  }
  return 1;
}
static inline int dynrf_vm_same_layout(const dynrf_vm_t* a, const dynrf_vm_t* b) {
  for (int i = 0; i < 3; ++i)
    if (a->sv[i] != b->sv[i] || a->su[i] != b->su[i]) return 0;
  return 1;
}

namespace dynrf {

constexpr int kTileM = 128;        // samples per MLP tile
constexpr int kMlpThreads = 256;   // threads per MLP-tile CTA
constexpr int kRayThreads = 128;   // threads per ray CTA in the scan kernels
constexpr int kTFeat = 48;         // per-ray time feature row: [t, PE(t,8)](17) | t_out(30) | pad

struct AabbNorm {
  float mn[3];
  float inv[3];  // 2 / (max - min), computed on device exactly like torch: 2.0f / (mx - mn)
};

__device__ __forceinline__ AabbNorm make_norm(const dynrf_field_cfg_t& c) {
  AabbNorm a;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    a.mn[i] = c.aabb_min[i];
    a.inv[i] = __fdiv_rn(2.0f, __fsub_rn(c.aabb_max[i], c.aabb_min[i]));
  }
  return a;
}

// models/tensorBase.py:425-428  (x - aabb0) * invaabbSize - 1
__device__ __forceinline__ float norm_coord(float x, float mn, float inv) {
  return __fsub_rn(__fmul_rn(__fsub_rn(x, mn), inv), 1.0f);
}
// models/tensorBase.py:430-433  (xn + 1) / invaabbSize + aabb0
__device__ __forceinline__ float unnorm_coord(float xn, float mn, float inv) {
  return __fadd_rn(__fdiv_rn(__fadd_rn(xn, 1.0f), inv), mn);
}

// ---- bilinear / linear taps: F.grid_sample(bilinear, zeros, align_corners=True) on the
// stride-s sub-sampled view (SURVEY 8a "lookup index math").
struct Tap1 {
  int i0, i1;     // texel indices in the FULL-resolution axis (already multiplied by stride); valid only if w != 0
  float w0, w1;   // weights, 0 when the tap is out of range
  float r0, r1;   // raw weights (1 - frac, frac) regardless of range      (backward)
  float scale;    // d(index)/d(coord) = (n_s - 1) / 2                      (backward)
  bool v0, v1;    // tap in range                                           (backward)
};

__device__ __forceinline__ Tap1 make_tap(float x, int n, int s) {
  const int ns = (n + s - 1) / s;
  // grid_sampler_unnormalize(align_corners=True): ((x + 1) / 2) * (size - 1)
  const float f = __fmul_rn(__fmul_rn(__fadd_rn(x, 1.0f), 0.5f), (float)(ns - 1));
  Tap1 t;
  t.scale = 0.5f * (float)(ns - 1);
  if (!(f > -1.0f && f < (float)ns)) {  // also catches NaN
    t.i0 = t.i1 = 0;
    t.w0 = t.w1 = t.r0 = t.r1 = 0.f;
    t.v0 = t.v1 = false;
    return t;
  }
  const float fl = floorf(f);
  const int i0 = (int)fl;
  const float w1 = __fsub_rn(f, fl);
  const float w0 = __fsub_rn(1.0f, w1);  // ATen CPU kernel: e = 1 - w
  const bool v0 = (i0 >= 0) && (i0 < ns);
  const bool v1 = (i0 + 1 >= 0) && (i0 + 1 < ns);
  t.i0 = v0 ? i0 * s : 0;
  t.i1 = v1 ? (i0 + 1) * s : 0;
  t.w0 = v0 ? w0 : 0.f;
  t.w1 = v1 ? w1 : 0.f;
  t.r0 = w0;
  t.r1 = w1;
  t.v0 = v0;
  t.v1 = v1;
  return t;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 fma4(float4 v, float w, float4 acc) {
  acc.x = fmaf(v.x, w, acc.x);
  acc.y = fmaf(v.y, w, acc.y);
  acc.z = fmaf(v.z, w, acc.z);
  acc.w = fmaf(v.w, w, acc.w);
  return acc;
}
__device__ __forceinline__ float4 mul4(float4 v, float w) { return make_float4(v.x * w, v.y * w, v.z * w, v.w * w); }
__device__ __forceinline__ float4 mul44(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }

// 4 channels [c..c+3] of plane (bilinear) x line (linear): one "channel group" product.
// plane: channel-last texels with strides st; tu indexes W, tv indexes H.  Weight order nw, ne, sw, se as ATen.
// texel strides of one plane (dynrf_vm_t.sv / .su): element (v, u) lives at (v * sv + u * su) * C
struct PlaneSt {
  int sv, su;
};
__device__ __forceinline__ PlaneSt plane_st(const dynrf_vm_t& vm, int pl) { return PlaneSt{vm.sv[pl], vm.su[pl]}; }

__device__ __forceinline__ float4 vm_group(const float* __restrict__ plane, const float* __restrict__ line, int C, PlaneSt st,
                                           const Tap1& tu, const Tap1& tv, const Tap1& tl, int c) {
  const float wnw = tu.w0 * tv.w0, wne = tu.w1 * tv.w0, wsw = tu.w0 * tv.w1, wse = tu.w1 * tv.w1;
  const size_t r0 = (size_t)tv.i0 * st.sv, r1 = (size_t)tv.i1 * st.sv;
  const size_t c0 = (size_t)tu.i0 * st.su, c1 = (size_t)tu.i1 * st.su;
  float4 p = mul4(ldg4(plane + (r0 + c0) * C + c), wnw);
  p = fma4(ldg4(plane + (r0 + c1) * C + c), wne, p);
  p = fma4(ldg4(plane + (r1 + c0) * C + c), wsw, p);
  p = fma4(ldg4(plane + (r1 + c1) * C + c), wse, p);
  float4 l = mul4(ldg4(line + (size_t)tl.i0 * C + c), tl.w0);
  l = fma4(ldg4(line + (size_t)tl.i1 * C + c), tl.w1, l);
  return mul44(p, l);
}

// matMode / vecMode (models/tensorBase.py:326-327)
__device__ __forceinline__ void plane_axes(int i, int& u, int& v, int& l) {
  u = (i == 2) ? 1 : 0;
  v = (i == 0) ? 1 : 2;
  l = 2 - i;
}

// ---- warp / block scans of products
__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float u = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= u;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float density_act(float f, int act, float shift) {
  if (act == DYNRF_ACT_RELU) return fmaxf(f, 0.f);
  const float x = f + shift;  // F.softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}

// ---- cp.async (LDGSTS) helpers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace dynrf
