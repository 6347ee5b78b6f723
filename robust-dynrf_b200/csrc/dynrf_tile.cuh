// Tile helpers shared by the forward (TM = 128, LD = 128) and backward (TM = 64, LD = 68) sample-tile
// kernels: per-tile scratch, VM feature gather into feature-major rows, its backward (scatter with
// vector atomics + coordinate gradients), and the shared [xn, PE(xn,10), t, PE(t,8)] input rows.
#pragma once
#include "dynrf_mlp.cuh"

namespace dynrf {

template <int TM>
struct TileScratch {
  float xn[3][TM];  // normalize(xyz)                 (un-warped, "xyz_sampled" of the reference's feature fns)
  float xq[3][TM];  // second coordinate set (kernel specific: xn2 / normalize(xyz_prime) / normalize(xyz))
  float xu[3][TM];  // unnormalize(xn) (warp) / view direction (static appearance)
  float gx[3][TM];  // coordinate-gradient accumulator (backward)
  float go[4][TM];  // upstream gradient rows of the narrow output layer (backward)
  int ray[TM];      // ray index of the sample, -1 = padding row
  int idx[TM];      // flat sample index
  int on[TM];       // sample takes part in gathers (valid / active)
};

// rows [row_base .. row_base + sumC) of A <- plane*line products at coordinate set xq, stride s.
// Work item = (sample m, channel group of 4); the 256/TM threads that share a sample split the groups.
template <int TM, int LD>
__device__ __forceinline__ void gather_vm_rows(const dynrf_vm_t& vm, const float (*xq)[TM], int stride, int row_base,
                                               float* __restrict__ A, int tid, const int* __restrict__ on) {
  constexpr int NPART = kMlpThreads / TM;
  const int m = tid % TM, part = tid / TM;
  const bool act = on[m] != 0;
  const float x3[3] = {xq[0][m], xq[1][m], xq[2][m]};
  int q = 0, coff = 0;
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    const int C = vm.C[pl], ng = C >> 2;
    int u, v, l;
    plane_axes(pl, u, v, l);
    Tap1 tu, tv, tl;
    if (act) {
      tu = make_tap(x3[u], vm.W[pl], stride);
      tv = make_tap(x3[v], vm.H[pl], stride);
      tl = make_tap(x3[l], vm.L[pl], stride);
    }
    for (int g = (part - q % NPART + NPART) % NPART; g < ng; g += NPART) {
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act) r = vm_group(vm.plane[pl], vm.line[pl], C, plane_st(vm, pl), tu, tv, tl, g * 4);
      float* a = A + (size_t)(row_base + coff + g * 4) * LD + m;
      a[0] = r.x;
      a[LD] = r.y;
      a[2 * LD] = r.z;
      a[3 * LD] = r.w;
    }
    q += ng;
    coff += C;
  }
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// backward of vm_group for one channel group: g = d loss / d (plane*line) for channels c..c+3.
// Adds into the plane / line gradients (vector reductions) and returns d loss / d (xu, xv, xl) in
// normalised-coordinate units.  Out-of-range taps carry value 0 (zeros padding) and get no gradient.
__device__ __forceinline__ float3 vm_group_bwd(const float* __restrict__ plane, const float* __restrict__ line,
                                               float* __restrict__ gplane, float* __restrict__ gline, int C, PlaneSt st,
                                               const Tap1& tu, const Tap1& tv, const Tap1& tl, int c, float4 g,
                                               float4* feat = nullptr /* optional: the forward value plane*line of the group */,
                                               bool need_gx = true /* false: skip d/d(coords), returns zeros */) {
  const size_t r0_ = (size_t)tv.i0 * st.sv, r1_ = (size_t)tv.i1 * st.sv, c0_ = (size_t)tu.i0 * st.su, c1_ = (size_t)tu.i1 * st.su;
  const size_t o00 = (r0_ + c0_) * C + c, o01 = (r0_ + c1_) * C + c;
  const size_t o10 = (r1_ + c0_) * C + c, o11 = (r1_ + c1_) * C + c;
  const bool b00 = tu.v0 && tv.v0, b01 = tu.v1 && tv.v0, b10 = tu.v0 && tv.v1, b11 = tu.v1 && tv.v1;
  const float4 P00 = b00 ? ldg4(plane + o00) : zero4(), P01 = b01 ? ldg4(plane + o01) : zero4();
  const float4 P10 = b10 ? ldg4(plane + o10) : zero4(), P11 = b11 ? ldg4(plane + o11) : zero4();
  const size_t l0 = (size_t)tl.i0 * C + c, l1 = (size_t)tl.i1 * C + c;
  const float4 L0 = tl.v0 ? ldg4(line + l0) : zero4(), L1 = tl.v1 ? ldg4(line + l1) : zero4();
  const float w00 = tu.r0 * tv.r0, w01 = tu.r1 * tv.r0, w10 = tu.r0 * tv.r1, w11 = tu.r1 * tv.r1;
  float4 p = mul4(P00, w00);
  p = fma4(P01, w01, p);
  p = fma4(P10, w10, p);
  p = fma4(P11, w11, p);
  float4 l = mul4(L0, tl.r0);
  l = fma4(L1, tl.r1, l);
  const float4 dp = mul44(g, l), dl = mul44(g, p);
  if (feat) *feat = mul44(p, l);
  if (b00) red_add4(gplane + o00, dp.x * w00, dp.y * w00, dp.z * w00, dp.w * w00);
  if (b01) red_add4(gplane + o01, dp.x * w01, dp.y * w01, dp.z * w01, dp.w * w01);
  if (b10) red_add4(gplane + o10, dp.x * w10, dp.y * w10, dp.z * w10, dp.w * w10);
  if (b11) red_add4(gplane + o11, dp.x * w11, dp.y * w11, dp.z * w11, dp.w * w11);
  if (tl.v0) red_add4(gline + l0, dl.x * tl.r0, dl.y * tl.r0, dl.z * tl.r0, dl.w * tl.r0);
  if (tl.v1) red_add4(gline + l1, dl.x * tl.r1, dl.y * tl.r1, dl.z * tl.r1, dl.w * tl.r1);
  float3 gx = make_float3(0.f, 0.f, 0.f);
  if (need_gx) {
    const float4 dpu = fma4(sub4(P11, P10), tv.r1, mul4(sub4(P01, P00), tv.r0));
    const float4 dpv = fma4(sub4(P11, P01), tu.r1, mul4(sub4(P10, P00), tu.r0));
    gx.x = dot4(dp, dpu) * tu.scale;
    gx.y = dot4(dp, dpv) * tv.scale;
    gx.z = dot4(dl, sub4(L1, L0)) * tl.scale;
  }
  return gx;
}

// ---- warp-aggregated variant ---------------------------------------------------------------------------------------
// Lanes of a warp hold consecutive samples of one ray, so neighbouring lanes very often hit the SAME texels (NDC rays run
// almost parallel to z: the 16/48-channel xy plane and the x / y lines see runs of 10-30 identical taps).  Identical
// addresses serialise in the L2 atomic unit, so each run of adjacent lanes with identical taps is summed with a segmented
// shuffle reduction and only the run head issues the reduction.  Non-adjacent duplicates simply issue separately.
struct WarpRuns {
  bool head;      // first lane of a run of identical keys
  int remaining;  // lanes in [lane, end of run)
  int maxrem;     // warp-wide maximum (uniform): 1 = no aggregation possible
};
__device__ __forceinline__ WarpRuns warp_runs(int k0, int k1, int lane) {
  const int p0 = __shfl_up_sync(0xffffffffu, k0, 1), p1 = __shfl_up_sync(0xffffffffu, k1, 1);
  const bool newrun = lane == 0 || p0 != k0 || p1 != k1;
  const uint32_t heads = __ballot_sync(0xffffffffu, newrun);
  const uint32_t above = lane == 31 ? 0u : (heads >> (lane + 1));
  WarpRuns r;
  r.head = newrun;
  r.remaining = above ? __ffs(above) : 32 - lane;
  r.maxrem = __reduce_max_sync(0xffffffffu, r.remaining);
  return r;
}
__device__ __forceinline__ float4 seg_reduce4(float4 v, const WarpRuns& r) {
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    if (off >= r.maxrem) break;  // warp-uniform
    const float ox = __shfl_down_sync(0xffffffffu, v.x, off), oy = __shfl_down_sync(0xffffffffu, v.y, off);
    const float oz = __shfl_down_sync(0xffffffffu, v.z, off), ow = __shfl_down_sync(0xffffffffu, v.w, off);
    if (off < r.remaining) {
      v.x += ox;
      v.y += oy;
      v.z += oz;
      v.w += ow;
    }
  }
  return v;
}
__device__ __forceinline__ int tap_key(const Tap1& t, bool act) {
  return act ? (t.i0 | ((int)t.v0 << 29) | ((int)t.v1 << 30)) : -1;
}

// Same contract as vm_group_bwd, but must be called by all 32 lanes of the warp (act = false for idle lanes).
__device__ __forceinline__ float3 vm_group_bwd_agg(const float* __restrict__ plane, const float* __restrict__ line,
                                                   float* __restrict__ gplane, float* __restrict__ gline, int C, PlaneSt st,
                                                   const Tap1& tu, const Tap1& tv, const Tap1& tl, int c, float4 g, bool act,
                                                   int lane, float4* feat = nullptr, bool need_gx = true) {
  const size_t r0_ = (size_t)tv.i0 * st.sv, r1_ = (size_t)tv.i1 * st.sv, c0_ = (size_t)tu.i0 * st.su, c1_ = (size_t)tu.i1 * st.su;
  const size_t o00 = (r0_ + c0_) * C + c, o01 = (r0_ + c1_) * C + c;
  const size_t o10 = (r1_ + c0_) * C + c, o11 = (r1_ + c1_) * C + c;
  const bool b00 = act && tu.v0 && tv.v0, b01 = act && tu.v1 && tv.v0, b10 = act && tu.v0 && tv.v1, b11 = act && tu.v1 && tv.v1;
  const float4 P00 = b00 ? ldg4(plane + o00) : zero4(), P01 = b01 ? ldg4(plane + o01) : zero4();
  const float4 P10 = b10 ? ldg4(plane + o10) : zero4(), P11 = b11 ? ldg4(plane + o11) : zero4();
  const size_t l0 = (size_t)tl.i0 * C + c, l1 = (size_t)tl.i1 * C + c;
  const bool c0 = act && tl.v0, c1 = act && tl.v1;
  const float4 L0 = c0 ? ldg4(line + l0) : zero4(), L1 = c1 ? ldg4(line + l1) : zero4();
  const float w00 = tu.r0 * tv.r0, w01 = tu.r1 * tv.r0, w10 = tu.r0 * tv.r1, w11 = tu.r1 * tv.r1;
  float4 p = mul4(P00, w00);
  p = fma4(P01, w01, p);
  p = fma4(P10, w10, p);
  p = fma4(P11, w11, p);
  float4 l = mul4(L0, tl.r0);
  l = fma4(L1, tl.r1, l);
  const float4 dp = act ? mul44(g, l) : zero4(), dl = act ? mul44(g, p) : zero4();
  if (feat) *feat = act ? mul44(p, l) : zero4();
  {
    const WarpRuns ru = warp_runs(tap_key(tu, act), tap_key(tv, act), lane);
    const float4 d00 = seg_reduce4(b00 ? mul4(dp, w00) : zero4(), ru), d01 = seg_reduce4(b01 ? mul4(dp, w01) : zero4(), ru);
    const float4 d10 = seg_reduce4(b10 ? mul4(dp, w10) : zero4(), ru), d11 = seg_reduce4(b11 ? mul4(dp, w11) : zero4(), ru);
    if (ru.head) {
      if (b00) red_add4(gplane + o00, d00.x, d00.y, d00.z, d00.w);
      if (b01) red_add4(gplane + o01, d01.x, d01.y, d01.z, d01.w);
      if (b10) red_add4(gplane + o10, d10.x, d10.y, d10.z, d10.w);
      if (b11) red_add4(gplane + o11, d11.x, d11.y, d11.z, d11.w);
    }
  }
  {
    const WarpRuns ru = warp_runs(tap_key(tl, act), 0, lane);
    const float4 e0 = seg_reduce4(c0 ? mul4(dl, tl.r0) : zero4(), ru), e1 = seg_reduce4(c1 ? mul4(dl, tl.r1) : zero4(), ru);
    if (ru.head) {
      if (c0) red_add4(gline + l0, e0.x, e0.y, e0.z, e0.w);
      if (c1) red_add4(gline + l1, e1.x, e1.y, e1.z, e1.w);
    }
  }
  float3 gx = make_float3(0.f, 0.f, 0.f);
  if (need_gx) {
    const float4 dpu = fma4(sub4(P11, P10), tv.r1, mul4(sub4(P01, P00), tv.r0));
    const float4 dpv = fma4(sub4(P11, P01), tu.r1, mul4(sub4(P10, P00), tu.r0));
    gx.x = dot4(dp, dpu) * tu.scale;
    gx.y = dot4(dp, dpv) * tv.scale;
    gx.z = dot4(dl, sub4(L1, L0)) * tl.scale;
  }
  return gx;
}

// Compile-time choice between the two scatters.  Measured on B200 (cfg 2): once layer 1 is no longer recomputed the
// backward tile kernels are dominated by this scatter and the L2 reduction rate (~45 red sectors / clk chip-wide) becomes the
// limiter: aggregation takes k_dyn_mlp_bwd_tc from 2.77 to 2.52 ms and k_static_density_bwd from 0.29 to 0.24 ms even with
// the incoherent taps of a randomly initialised warp field (with layer-1 recompute in the same kernel it was a 3 % loss).
constexpr bool kAggScatter = true;
__device__ __forceinline__ float3 vm_group_bwd_w(const float* __restrict__ plane, const float* __restrict__ line,
                                                 float* __restrict__ gplane, float* __restrict__ gline, int C, PlaneSt st, const Tap1& tu,
                                                 const Tap1& tv, const Tap1& tl, int c, float4 g, bool act, int lane,
                                                 float4* feat = nullptr, bool need_gx = true) {
  if constexpr (kAggScatter) {
    return vm_group_bwd_agg(plane, line, gplane, gline, C, st, tu, tv, tl, c, g, act, lane, feat, need_gx);
  } else {
    if (!act) {
      if (feat) *feat = zero4();
      return make_float3(0.f, 0.f, 0.f);
    }
    return vm_group_bwd(plane, line, gplane, gline, C, st, tu, tv, tl, c, g, feat, need_gx);
  }
}

// backward of gather_vm_rows: dA rows -> plane/line gradients; coordinate gradients are added to gxs[3][TM].
template <int TM, int LD>
__device__ __forceinline__ void scatter_vm_rows(const dynrf_vm_t& vm, const dynrf_vm_grad_t& gvm, const float (*xq)[TM],
                                                int stride, int row_base, const float* __restrict__ dA, int tid,
                                                const int* __restrict__ on, float (*gxs)[TM],
                                                float* __restrict__ Afeat = nullptr /* optional: forward rows (gather_vm_rows) */) {
  constexpr int NPART = kMlpThreads / TM;
  static_assert(TM % 32 == 0, "a warp must hold 32 consecutive samples");
  const int m = tid % TM, part = tid / TM;
  const bool act = on[m] != 0;
  const float x3[3] = {xq[0][m], xq[1][m], xq[2][m]};
  float gacc[3] = {0.f, 0.f, 0.f};
  int q = 0, coff = 0;
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    const int C = vm.C[pl], ng = C >> 2;
    int u, v, l;
    plane_axes(pl, u, v, l);
    const Tap1 tu = make_tap(x3[u], vm.W[pl], stride);
    const Tap1 tv = make_tap(x3[v], vm.H[pl], stride);
    const Tap1 tl = make_tap(x3[l], vm.L[pl], stride);
    for (int g = (part - q % NPART + NPART) % NPART; g < ng; g += NPART) {
      const float* a = dA + (size_t)(row_base + coff + g * 4) * LD + m;
      const float4 gg = make_float4(a[0], a[LD], a[2 * LD], a[3 * LD]);
      float3 r;
      if (Afeat) {   // fused recompute: the taps are loaded once for the scatter and for the forward value
        float4 fv;
        r = vm_group_bwd_w(vm.plane[pl], vm.line[pl], gvm.plane[pl], gvm.line[pl], C, plane_st(vm, pl), tu, tv, tl, g * 4, gg, act, tid & 31, &fv,
                           gxs != nullptr);
        float* fa = Afeat + (size_t)(row_base + coff + g * 4) * LD + m;
        fa[0] = fv.x;
        fa[LD] = fv.y;
        fa[2 * LD] = fv.z;
        fa[3 * LD] = fv.w;
      } else {
        r = vm_group_bwd_w(vm.plane[pl], vm.line[pl], gvm.plane[pl], gvm.line[pl], C, plane_st(vm, pl), tu, tv, tl, g * 4, gg, act, tid & 31);
      }
      gacc[u] += r.x;
      gacc[v] += r.y;
      gacc[l] += r.z;
    }
    q += ng;
    coff += C;
  }
  if (gxs && act) {
#pragma unroll
    for (int a = 0; a < 3; ++a) atomicAdd(&gxs[a][m], gacc[a]);
  }
}

// rows [r0, r0+80) <- [xn(3), sin PE(xn,10) (30), cos (30), t, sin PE(t,8), cos PE(t,8)]  (needs scr.xn, scr.ray)
template <int TM, int LD>
__device__ __forceinline__ void fill_pos_time_rows(const TileScratch<TM>& scr, const float* __restrict__ tfeat,
                                                   float* __restrict__ A, int r0, int tid) {
  for (int item = tid; item < 3 * TM; item += kMlpThreads) {
    const int m = item % TM, d = item / TM;
    const float x = scr.xn[d][m];
    A[(size_t)(r0 + d) * LD + m] = x;
    pe_rows<10, LD>(x, A, r0 + 3 + d * 10, r0 + 33 + d * 10, m);
  }
  for (int item = tid; item < 17 * TM; item += kMlpThreads) {
    const int m = item % TM, k = item / TM;
    const int ry = scr.ray[m];
    A[(size_t)(r0 + 63 + k) * LD + m] = ry >= 0 ? __ldg(tfeat + (size_t)ry * kTFeat + k) : 0.f;
  }
}

// warp-MLP input rows: 0..2 xn2 | 3..32 sin | 33..62 cos | 63..92 t_out | 93..95 zero   (needs scr.xq, scr.ray)
template <int TM, int LD>
__device__ __forceinline__ void fill_warp_rows(const TileScratch<TM>& scr, const float* __restrict__ tfeat, float* __restrict__ A,
                                               int tid) {
  for (int item = tid; item < 3 * TM; item += kMlpThreads) {
    const int m = item % TM, d = item / TM;
    const float x = scr.xq[d][m];
    A[(size_t)d * LD + m] = x;
    pe_rows<10, LD>(x, A, 3 + d * 10, 33 + d * 10, m);
  }
  for (int item = tid; item < 33 * TM; item += kMlpThreads) {
    const int m = item % TM, k = item / TM;
    const int ry = scr.ray[m];
    A[(size_t)(63 + k) * LD + m] = (k < 30 && ry >= 0) ? __ldg(tfeat + (size_t)ry * kTFeat + 17 + k) : 0.f;
  }
}

__device__ __forceinline__ void ray_viewdir(const float* __restrict__ rays, int ry, float (&vd)[3]) {
  // viewdirs / torch.norm(viewdirs)  (models/tensorBase.py:723-725)
  const float dx = __ldg(rays + (size_t)ry * 6 + 3), dy = __ldg(rays + (size_t)ry * 6 + 4), dz = __ldg(rays + (size_t)ry * 6 + 5);
  const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  vd[0] = __fdiv_rn(dx, nrm);
  vd[1] = __fdiv_rn(dy, nrm);
  vd[2] = __fdiv_rn(dz, nrm);
}

static int g_sm_count = 0;
static inline int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count = n > 0 ? n : 148;
  }
  return g_sm_count;
}

}  // namespace dynrf
