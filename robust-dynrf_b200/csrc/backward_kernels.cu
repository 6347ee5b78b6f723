// Backward kernels of the ray-march path (the gradients the reference obtains from autograd through
// grid_sampler_2d_backward / addmm / cumprod, SURVEY.md 7 hard part 2).
//
//   per ray : alpha->weight scan backward, two-field compositing backward (reverse suffix scans)
//   per sample : static density lookup backward (plane/line reductions + d/dxyz)
//   per 64-sample tile (recompute forward in shared memory, then back-propagate):
//       static appearance, dynamic appearance, dynamic density / blending MLP, coordinate warp
//   per ray block : time-embedding MLP
//
// Gradients of planes/lines and of packed weights are accumulated with vector reductions
// (REDG.E.ADD.F32x4); weight gradients are first summed over a CTA's tiles in registers where the
// register budget allows, otherwise flushed per tile.
#include "dynrf_tile.cuh"

// read-once activation streams: evict-first (see dynrf_tc.cuh)
#ifdef DYNRF_STREAM_HINTS
#define DYNRF_LD_STREAM_F(ptr) __ldcs(ptr)
#else
#define DYNRF_LD_STREAM_F(ptr) __ldg(ptr)
#endif

namespace dynrf {

constexpr int kBT = 64;       // backward tile
constexpr int kBL = kBT + 4;  // row stride: keeps row-strided 16-byte reads conflict-free (wgrad_acc)
using BScratch = TileScratch<kBT>;

// ------------------------------------------------------------------------------------------------
// packing of nn.Linear weights for the data-gradient GEMM: wb[n][j] = w[n][k0 + j]  (n < N, j < k1-k0)

__global__ void k_pack_linear_bwd(const float* __restrict__ w, int N, int K, int k0, int k1, int Nrows, int Jpad,
                                  float* __restrict__ wb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Nrows * Jpad) {
    const int n = i / Jpad, j = i % Jpad;
    wb[i] = (n < N && k0 + j < k1) ? __ldg(w + (size_t)n * K + k0 + j) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// per-ray reverse scans

// exclusive suffix sum over the block's 128 values (thread order = sample order) with carry from later chunks
__device__ __forceinline__ float block_excl_suffix_sum(float v, float& carry, float* wtot /*[4]*/, int lane, int wid) {
  float incl = v;  // inclusive suffix within the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float u = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += u;
  }
  float excl = __shfl_down_sync(0xffffffffu, incl, 1);
  if (lane == 31) excl = 0.f;
  if (lane == 0) wtot[wid] = incl;
  __syncthreads();
  float post = carry, total = carry;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float t = wtot[w];
    if (w > wid) post += t;
    total += t;
  }
  __syncthreads();
  carry = total;
  return post + excl;
}

__device__ __forceinline__ float block_excl_prod_b(float f, float& carry, float* wtot, int lane, int wid) {
  const float incl = warp_incl_prod(f, lane);
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  float pre = carry, total = carry;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float t = wtot[w];
    if (w < wid) pre *= t;
    total *= t;
  }
  __syncthreads();
  carry = total;
  return pre * excl;
}

__device__ __forceinline__ float block_sum1(float v, float* red /*[4]*/, int lane, int wid) {
  v = warp_sum(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const float r = (red[0] + red[1]) + (red[2] + red[3]);
  __syncthreads();
  return r;
}

// weight = alpha * T backward (raw2alpha, models/tensorBase.py:22-34).  dists are the scaled distances.
// gsigma += d/dsigma, gdists += d/ddists
__global__ void __launch_bounds__(kRayThreads)
k_alpha_weight_bwd(int N, int S, const float* __restrict__ sigma, const float* __restrict__ dists, const float* __restrict__ gweight,
                   float* __restrict__ gsigma, float* __restrict__ gdists) {
  extern __shared__ float Tbuf[];  // [S]
  __shared__ float wt[4];
  const int n = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float carry = 1.f;
  const int nch = (S + kRayThreads - 1) / kRayThreads;
  for (int c = 0; c < nch; ++c) {
    const int s = c * kRayThreads + threadIdx.x;
    const bool in = s < S;
    const size_t i = (size_t)n * S + (in ? s : 0);
    const float a = in ? __fsub_rn(1.0f, expf(-__fmul_rn(__ldg(sigma + i), __ldg(dists + i)))) : 0.f;
    const float f = in ? __fadd_rn(__fsub_rn(1.0f, a), 1e-10f) : 1.f;
    const float T = block_excl_prod_b(f, carry, wt, lane, wid);
    if (in) Tbuf[s] = T;
  }
  __syncthreads();
  float sc = 0.f;
  for (int c = nch - 1; c >= 0; --c) {
    const int s = c * kRayThreads + threadIdx.x;
    const bool in = s < S;
    const size_t i = (size_t)n * S + (in ? s : 0);
    float sg = 0.f, ds = 0.f, a = 0.f, e = 1.f, T = 0.f, g = 0.f;
    if (in) {
      sg = __ldg(sigma + i);
      ds = __ldg(dists + i);
      e = expf(-__fmul_rn(sg, ds));
      a = __fsub_rn(1.0f, e);
      T = Tbuf[s];
      g = __ldg(gweight + i);
    }
    const float suf = block_excl_suffix_sum(g * a * T, sc, wt, lane, wid);
    if (in) {
      const float f = __fadd_rn(__fsub_rn(1.0f, a), 1e-10f);
      const float da = g * T - suf / f;
      gsigma[i] += da * ds * e;
      gdists[i] += da * sg * e;
    }
  }
}

// renderer.raw2outputs backward (renderer.py:173-315).  See DESIGN.md "compositing backward" for the algebra.
// go [N][16] upstream of out_ray, gwf/gws/gwd [N][S] upstream of the three weight tensors (may be NULL).
// Outputs (overwritten): g_rgb_s, g_rgb_d [N][S][3]; g_sigma_s, g_sigma_d, g_dists, g_blend [N][S]; g_rays [N][6].
__global__ void __launch_bounds__(kRayThreads)
k_composite_bwd(int ray_type, int white_bg, const float* __restrict__ rgb_s, const float* __restrict__ sigma_s,
                const float* __restrict__ rgb_d, const float* __restrict__ sigma_d, const float* __restrict__ dists,
                const float* __restrict__ blending, const float* __restrict__ z_vals, const float* __restrict__ rays, int N, int S,
                const float* __restrict__ go, const float* __restrict__ gwf, const float* __restrict__ gws, const float* __restrict__ gwd,
                float* __restrict__ g_rgb_s, float* __restrict__ g_sigma_s, float* __restrict__ g_rgb_d, float* __restrict__ g_sigma_d,
                float* __restrict__ g_dists, float* __restrict__ g_blend, float* __restrict__ g_rays) {
  extern __shared__ float Tb[];  // Td[S] Ts[S] Tf[S]
  float* Td = Tb;
  float* Ts = Tb + S;
  float* Tf = Tb + 2 * S;
  __shared__ float wt[3][4];
  __shared__ float red[4];
  const int n = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nch = (S + kRayThreads - 1) / kRayThreads;
  float cd = 1.f, cs = 1.f, cf = 1.f;
  float a_rf[3] = {0, 0, 0}, a_rs[3] = {0, 0, 0}, a_accf = 0.f, a_accs = 0.f, a_totd = 0.f;
  // ---- pass A: transmittances + forward sums
  for (int c = 0; c < nch; ++c) {
    const int s = c * kRayThreads + threadIdx.x;
    const bool in = s < S;
    const size_t i = (size_t)n * S + (in ? s : 0);
    float ad = 0.f, as = 0.f, b = 0.f;
    if (in) {
      const float dist = __ldg(dists + i);
      ad = __fsub_rn(1.0f, expf(-__fmul_rn(__ldg(sigma_d + i), dist)));
      as = __fsub_rn(1.0f, expf(-__fmul_rn(__ldg(sigma_s + i), dist)));
      b = __ldg(blending + i);
    }
    const float omb = __fsub_rn(1.0f, b);
    const float fd = in ? __fadd_rn(__fsub_rn(1.0f, ad), 1e-10f) : 1.f;
    const float fs = in ? __fadd_rn(__fsub_rn(1.0f, as), 1e-10f) : 1.f;
    const float ff = in ? __fadd_rn(__fmul_rn(__fsub_rn(1.0f, __fmul_rn(ad, b)), __fsub_rn(1.0f, __fmul_rn(as, omb))), 1e-10f) : 1.f;
    const float tD = block_excl_prod_b(fd, cd, wt[0], lane, wid);
    const float tS = block_excl_prod_b(fs, cs, wt[1], lane, wid);
    const float tF = block_excl_prod_b(ff, cf, wt[2], lane, wid);
    if (in) {
      Td[s] = tD;
      Ts[s] = tS;
      Tf[s] = tF;
      const float ws = as * tS, kd = tF * ad * b, ks = tF * as * omb;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float cS = __ldg(rgb_s + i * 3 + ch), cD = __ldg(rgb_d + i * 3 + ch);
        a_rf[ch] += kd * cD + ks * cS;
        a_rs[ch] += ws * cS;
      }
      a_accf += (ad * b + as * omb) * tF;
      a_accs += ws;
      a_totd += ad * tD;
    }
  }
  float rf[3], rs[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    rf[ch] = block_sum1(a_rf[ch], red, lane, wid);
    rs[ch] = block_sum1(a_rs[ch], red, lane, wid);
  }
  const float acc_f = block_sum1(a_accf, red, lane, wid);
  const float acc_s = block_sum1(a_accs, red, lane, wid);
  const float tot_d = block_sum1(a_totd, red, lane, wid);
  const float den = __fadd_rn(tot_d, 1e-10f);
  // ---- pass B: normalised dynamic maps (needed for the clamp masks and the normalisation backward)
  float a_rd[3] = {0, 0, 0}, a_accd = 0.f;
  for (int s = threadIdx.x; s < S; s += kRayThreads) {
    const size_t i = (size_t)n * S + s;
    const float ad = __fsub_rn(1.0f, expf(-__fmul_rn(__ldg(sigma_d + i), __ldg(dists + i))));
    const float wd = __fdiv_rn(ad * Td[s], den);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) a_rd[ch] += wd * __ldg(rgb_d + i * 3 + ch);
    a_accd += wd;
  }
  float rd[3];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) rd[ch] = block_sum1(a_rd[ch], red, lane, wid);
  const float acc_d = block_sum1(a_accd, red, lane, wid);

  const float* g = go + (size_t)n * DYNRF_RAY_OUT;
  const float relu_f = (1.0f - acc_f) > 0.f ? 1.f : 0.f;
  float GRf[3], GRs[3], GRd[3];
  float sGRf = 0.f, sGRs = 0.f, sGRd = 0.f;
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    const float pf = rf[ch] + (white_bg ? fmaxf(1.0f - acc_f, 0.f) : 0.f);
    const float ps = rs[ch] + (white_bg ? (1.0f - acc_s) : 0.f);
    const float pd = rd[ch] + (white_bg ? (1.0f - acc_d) : 0.f);
    GRf[ch] = (pf >= 0.f && pf <= 1.f) ? __ldg(g + ch) : 0.f;  // torch.clamp backward: pass inside [min,max]
    GRs[ch] = (ps >= 0.f && ps <= 1.f) ? __ldg(g + 5 + ch) : 0.f;
    GRd[ch] = (pd >= 0.f && pd <= 1.f) ? __ldg(g + 10 + ch) : 0.f;
    sGRf += GRf[ch];
    sGRs += GRs[ch];
    sGRd += GRd[ch];
  }
  const float gDf = __ldg(g + 3), gAf = __ldg(g + 4), gDs = __ldg(g + 8), gAs = __ldg(g + 9), gDd = __ldg(g + 13),
              gAd = __ldg(g + 14), gDyn = __ldg(g + 15);
  float bg = 0.f;
  if (ray_type == DYNRF_RAY_NDC) bg = __fadd_rn(__ldg(rays + (size_t)n * 6 + 2), __ldg(rays + (size_t)n * 6 + 5));
  else if (ray_type == DYNRF_RAY_CONTRACT) bg = 256.f;
  const float wsum = white_bg ? 1.f : 0.f;
  const float GAf = gAf - relu_f * (gDf * bg + wsum * sGRf);
  const float GAs = gAs - (gDs * bg + wsum * sGRs);
  const float GAd = gAd - (gDd * bg + wsum * sGRd);
  // ---- pass C: sum_k Gwd_k * w_d,k  (normalisation backward)
  float a_nd = 0.f;
  for (int s = threadIdx.x; s < S; s += kRayThreads) {
    const size_t i = (size_t)n * S + s;
    const float ad = __fsub_rn(1.0f, expf(-__fmul_rn(__ldg(sigma_d + i), __ldg(dists + i))));
    const float wd = __fdiv_rn(ad * Td[s], den);
    float G = (gwd ? __ldg(gwd + i) : 0.f) + gDd * __ldg(z_vals + i) + GAd;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) G += GRd[ch] * __ldg(rgb_d + i * 3 + ch);
    a_nd += G * wd;
  }
  const float nd = block_sum1(a_nd, red, lane, wid);
  // ---- pass D: reverse chunks with the three suffix sums
  float sd = 0.f, ss = 0.f, sf = 0.f;
  for (int c = nch - 1; c >= 0; --c) {
    const int s = c * kRayThreads + threadIdx.x;
    const bool in = s < S;
    const size_t i = (size_t)n * S + (in ? s : 0);
    float sgd = 0.f, sgs = 0.f, dist = 0.f, ed = 1.f, es = 1.f, ad = 0.f, as = 0.f, b = 0.f, z = 0.f;
    float cS[3] = {0, 0, 0}, cD[3] = {0, 0, 0};
    float tD = 0.f, tS = 0.f, tF = 0.f, Gwf = 0.f, Gws = 0.f, Gu = 0.f, dRd = 0.f, dRs = 0.f;
    if (in) {
      sgd = __ldg(sigma_d + i);
      sgs = __ldg(sigma_s + i);
      dist = __ldg(dists + i);
      ed = expf(-__fmul_rn(sgd, dist));
      es = expf(-__fmul_rn(sgs, dist));
      ad = __fsub_rn(1.0f, ed);
      as = __fsub_rn(1.0f, es);
      b = __ldg(blending + i);
      z = __ldg(z_vals + i);
      tD = Td[s];
      tS = Ts[s];
      tF = Tf[s];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        cS[ch] = __ldg(rgb_s + i * 3 + ch);
        cD[ch] = __ldg(rgb_d + i * 3 + ch);
        dRd += GRf[ch] * cD[ch];
        dRs += GRf[ch] * cS[ch];
      }
      Gwf = (gwf ? __ldg(gwf + i) : 0.f) + gDf * z + GAf + gDyn * b;
      Gws = (gws ? __ldg(gws + i) : 0.f) + gDs * z + GAs + GRs[0] * cS[0] + GRs[1] * cS[1] + GRs[2] * cS[2];
      const float Gwd = (gwd ? __ldg(gwd + i) : 0.f) + gDd * z + GAd + GRd[0] * cD[0] + GRd[1] * cD[1] + GRd[2] * cD[2];
      Gu = (Gwd - nd) / den;
    }
    const float omb = __fsub_rn(1.0f, b);
    const float u = ad * tD, ws = as * tS;
    const float V = (ad * b + as * omb) * Gwf + ad * b * dRd + as * omb * dRs;
    const float sufd = block_excl_suffix_sum(Gu * u, sd, wt[0], lane, wid);
    const float sufs = block_excl_suffix_sum(Gws * ws, ss, wt[1], lane, wid);
    const float suff = block_excl_suffix_sum(tF * V, sf, wt[2], lane, wid);
    if (in) {
      const float fd = __fadd_rn(__fsub_rn(1.0f, ad), 1e-10f);
      const float fs = __fadd_rn(__fsub_rn(1.0f, as), 1e-10f);
      const float q1 = __fsub_rn(1.0f, __fmul_rn(ad, b)), q2 = __fsub_rn(1.0f, __fmul_rn(as, omb));
      const float ff = __fadd_rn(__fmul_rn(q1, q2), 1e-10f);
      const float dff = suff / ff;
      float dad = Gu * tD - sufd / fd + tF * b * (Gwf + dRd) + dff * (-b * q2);
      float das = Gws * tS - sufs / fs + tF * omb * (Gwf + dRs) + dff * (-omb * q1);
      const float wfv = (ad * b + as * omb) * tF;
      const float db = tF * ((ad - as) * Gwf + ad * dRd - as * dRs) + wfv * gDyn + dff * (-ad * q2 + as * q1);
      g_sigma_d[i] = dad * dist * ed;
      g_sigma_s[i] = das * dist * es;
      g_dists[i] = dad * sgd * ed + das * sgs * es;
      g_blend[i] = db;
      const float wdn = __fdiv_rn(u, den);
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        g_rgb_s[i * 3 + ch] = ws * GRs[ch] + tF * as * omb * GRf[ch];
        g_rgb_d[i * 3 + ch] = wdn * GRd[ch] + tF * ad * b * GRf[ch];
      }
    }
  }
  if (threadIdx.x < 6) {
    float v = 0.f;
    if (ray_type == DYNRF_RAY_NDC && (threadIdx.x == 2 || threadIdx.x == 5))
      v = gDd * (1.0f - acc_d) + gDs * (1.0f - acc_s) + gDf * fmaxf(1.0f - acc_f, 0.f);
    g_rays[(size_t)n * 6 + threadIdx.x] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// static density lookup backward: gsig = d loss / d sigma [P]; relu mask via sigma > 0 (softplus via recompute)

__global__ void __launch_bounds__(128)
k_static_density_bwd(dynrf_field_cfg_t cfg, dynrf_vm_t vm, dynrf_vm_grad_t gvm, const float* __restrict__ xyz,
                     const uint8_t* __restrict__ valid, const float* __restrict__ sigma, const float* __restrict__ gsig, int P,
                     float* __restrict__ gxyz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool act = i < P && valid[i] != 0;   // no early exits: the scatter below is warp-cooperative
  float g = act ? __ldg(gsig + i) : 0.f;
  const AabbNorm nb = make_norm(cfg);
  float xn[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) xn[a] = act ? norm_coord(__ldg(xyz + (size_t)i * 3 + a), nb.mn[a], nb.inv[a]) : 0.f;
  if (cfg.act == DYNRF_ACT_RELU) {
    if (act && !(__ldg(sigma + i) > 0.f)) act = false;
  } else if (act) {
    // softplus'(x) = sigmoid(x): recompute the feature
    float feat = 0.f;
    for (int pl = 0; pl < 3; ++pl) {
      int u, v, l;
      plane_axes(pl, u, v, l);
      const Tap1 tu = make_tap(xn[u], vm.W[pl], 1), tv = make_tap(xn[v], vm.H[pl], 1), tl = make_tap(xn[l], vm.L[pl], 1);
      for (int c = 0; c < vm.C[pl]; c += 4) {
        const float4 r = vm_group(vm.plane[pl], vm.line[pl], vm.C[pl], plane_st(vm, pl), tu, tv, tl, c);
        feat += (r.x + r.y) + (r.z + r.w);
      }
    }
    g *= sigmoidf_(feat + cfg.density_shift);
  }
  if (g == 0.f) act = false;
  if (!__any_sync(0xffffffffu, act)) return;   // warp-uniform
  float gacc[3] = {0.f, 0.f, 0.f};
  const float4 g4 = make_float4(g, g, g, g);
#pragma unroll
  for (int pl = 0; pl < 3; ++pl) {
    int u, v, l;
    plane_axes(pl, u, v, l);
    const Tap1 tu = make_tap(xn[u], vm.W[pl], 1), tv = make_tap(xn[v], vm.H[pl], 1), tl = make_tap(xn[l], vm.L[pl], 1);
    for (int c = 0; c < vm.C[pl]; c += 4) {
      const float3 r = vm_group_bwd_w(vm.plane[pl], vm.line[pl], gvm.plane[pl], gvm.line[pl], vm.C[pl], plane_st(vm, pl), tu, tv, tl, c, g4, act,
                                      threadIdx.x & 31, nullptr, gxyz != nullptr);
      gacc[u] += r.x;
      gacc[v] += r.y;
      gacc[l] += r.z;
    }
  }
  if (gxyz && act) {
#pragma unroll
    for (int a = 0; a < 3; ++a) gxyz[(size_t)i * 3 + a] += gacc[a] * nb.inv[a];
  }
}

// ------------------------------------------------------------------------------------------------
// tile helpers

// db[n] += sum_m dY[n][m]  for n < nrows, one thread per row, accumulated in a register across tiles
__device__ __forceinline__ float row_sum(const float* __restrict__ dY, int n) {
  float s = 0.f;
  const float* r = dY + (size_t)n * kBL;
#pragma unroll 4
  for (int m = 0; m < kBT; m += 4) {
    const float4 t = *reinterpret_cast<const float4*>(r + m);
    s += (t.x + t.y) + (t.z + t.w);
  }
  return s;
}

struct LayerGrad {
  float* dwt;    // [Kpad][Npad] packed
  float* dbias;  // [Npad]
};

// ------------------------------------------------------------------------------------------------
// coordinate warp backward (a15).  gxp [P][3] = d loss / d xyz_prime (total).  Accumulates packed
// gradients of layer3/4/5 and gtout [N][32] (d loss / d t_out per ray).

constexpr size_t kWarpBwdSmem = (size_t)(96 * kBL + 64 * kBL + 64 * kBL + 2 * kChunkK * 64) * sizeof(float) + sizeof(BScratch);

__global__ void __launch_bounds__(kMlpThreads, 2)
k_warp_bwd(dynrf_field_cfg_t cfg, const float* __restrict__ xyz, int P, int S, const float* __restrict__ tfeat, dynrf_layer_t l3,
           dynrf_layer_t l4, dynrf_layer_t l5, const float* __restrict__ wb4 /*[64][64]*/, const float* __restrict__ wb3t /*[64][32]*/,
           const float* __restrict__ gxp, LayerGrad g3, LayerGrad g4, LayerGrad g5, float* __restrict__ gtout,
           const float* __restrict__ H1T /*[64][ld] relu(layer 3) kept by the forward, optional*/, const float* __restrict__ H2T, int ld) {
  extern __shared__ __align__(16) float smem[];
  float* A = smem;
  float* H1 = A + 96 * kBL;
  float* H2 = H1 + 64 * kBL;
  float* wbuf = H2 + 64 * kBL;
  BScratch& scr = *reinterpret_cast<BScratch*>(wbuf + 2 * kChunkK * 64);
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int ntiles = (P + kBT - 1) / kBT;
  float accW3[6][4], accW4[4][4];
  wgrad_zero(accW3);
  wgrad_zero(accW4);
  float accW5[3] = {0.f, 0.f, 0.f}, accB = 0.f;  // thread k<64: dW5[k][0..2]; bias roles below
  float accB3 = 0.f, accB4 = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kBT) {
      const int i = tile * kBT + tid;
      const bool in = i < P;
      scr.ray[tid] = in ? i / S : -1;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = in ? __ldg(xyz + (size_t)i * 3 + a) : 0.f;
        const float xn = norm_coord(x, nb.mn[a], nb.inv[a]);
        const float xu = unnorm_coord(xn, nb.mn[a], nb.inv[a]);
        scr.xq[a][tid] = norm_coord(xu, nb.mn[a], nb.inv[a]);
        scr.go[a][tid] = in ? __ldg(gxp + (size_t)i * 3 + a) : 0.f;
      }
    }
    __syncthreads();
    fill_warp_rows<kBT, kBL>(scr, tfeat, A, tid);
    __syncthreads();
    if (H1T) {
      // hidden activations kept by the forward kernel: 10 of the 26 kMAC per sample of this kernel were their recomputation
      for (int e = tid; e < 64 * kBT; e += kMlpThreads) {
        const int k = e / kBT, m = e % kBT;
        const size_t src = (size_t)k * ld + (size_t)tile * kBT + m;
        H1[(size_t)k * kBL + m] = DYNRF_LD_STREAM_F(H1T + src);
        H2[(size_t)k * kBL + m] = DYNRF_LD_STREAM_F(H2T + src);
      }
      __syncthreads();
    } else {
      dense_tile<64, kBT, kBL, true>(A, l3, H1, wbuf, tid);
      dense_tile<64, kBT, kBL, true>(H1, l4, H2, wbuf, tid);
    }
    // layer5: dW5[k][o] += sum_m H2[k][m] g[o][m]; db5[o] += sum_m g[o][m]
    if (tid < 64) {
      const float* h = H2 + (size_t)tid * kBL;
      for (int m = 0; m < kBT; ++m) {
        const float hv = h[m];
        accW5[0] = fmaf(hv, scr.go[0][m], accW5[0]);
        accW5[1] = fmaf(hv, scr.go[1][m], accW5[1]);
        accW5[2] = fmaf(hv, scr.go[2][m], accW5[2]);
      }
    } else if (tid < 67) {
      for (int m = 0; m < kBT; ++m) accB += scr.go[tid - 64][m];
    }
    __syncthreads();
    // dH2 = (W5 g) * [H2 > 0], in place
    for (int e = tid; e < 64 * kBT; e += kMlpThreads) {
      const int k = e / kBT, m = e % kBT;
      float v = 0.f;
#pragma unroll
      for (int o = 0; o < 3; ++o) v = fmaf(__ldg(l5.wt + (size_t)k * l5.Npad + o), scr.go[o][m], v);
      float* h = H2 + (size_t)k * kBL + m;
      *h = (*h > 0.f) ? v : 0.f;
    }
    __syncthreads();
    wgrad_acc<4, 4, kBT, kBL>(accW4, H1, 64, H2, tid);
    if (tid < 64) accB4 += row_sum(H2, tid);
    {
      Acc<64, kBT> acc;
      acc_init<64, kBT>(acc, nullptr, tid >> 4);
      acc_gemm<64, kBT, kBL>(acc, H2, 64, wb4, wbuf, tid);
      acc_store<64, kBT, kBL, false, true>(acc, H1, 64, tid, H1);  // dH1 over H1 (masked by H1 > 0)
    }
    wgrad_acc<6, 4, kBT, kBL>(accW3, A, 96, H1, tid);
    if (tid < 64) accB3 += row_sum(H1, tid);
    {
      Acc<32, kBT> acc;
      acc_init<32, kBT>(acc, nullptr, tid >> 4);
      acc_gemm<32, kBT, kBL>(acc, H1, 64, wb3t, wbuf, tid);
      acc_store<32, kBT, kBL, false>(acc, H2, 32, tid);  // d t_out rows in H2[0..32)
    }
    if (tid < 30) {
      int cur = -1;
      float sacc = 0.f;
      for (int m = 0; m < kBT; ++m) {
        const int ry = scr.ray[m];
        if (ry != cur) {
          if (cur >= 0) atomicAdd(gtout + (size_t)cur * 32 + tid, sacc);
          cur = ry;
          sacc = 0.f;
        }
        if (ry >= 0) sacc += H2[(size_t)tid * kBL + m];
      }
      if (cur >= 0) atomicAdd(gtout + (size_t)cur * 32 + tid, sacc);
    }
    __syncthreads();
  }
  wgrad_flush(accW3, g3.dwt, 96, 64, tid);
  wgrad_flush(accW4, g4.dwt, 64, 64, tid);
  if (tid < 64) {
    atomicAdd(g3.dbias + tid, accB3);
    atomicAdd(g4.dbias + tid, accB4);
#pragma unroll
    for (int o = 0; o < 3; ++o) atomicAdd(g5.dwt + (size_t)tid * 16 + o, accW5[o]);
  } else if (tid < 67) {
    atomicAdd(g5.dbias + (tid - 64), accB);
  }
}

// time-embedding MLP backward: gtout [N][32] -> packed grads of layer1 (17->64) and layer2 (64->30)
__global__ void __launch_bounds__(64)
k_time_embed_bwd(const float* __restrict__ ts, int N, dynrf_layer_t l1, dynrf_layer_t l2, const float* __restrict__ gtout,
                 LayerGrad g1, LayerGrad g2) {
  __shared__ float tin[20];
  __shared__ float hid[64];
  __shared__ float go[32];
  __shared__ float dh[64];
  const int j = threadIdx.x;
  float aW1[17], aB1 = 0.f, aB2 = 0.f;
  float aW2[64];
#pragma unroll
  for (int k = 0; k < 17; ++k) aW1[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 64; ++k) aW2[k] = 0.f;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const float t = __ldg(ts + n);
    if (j == 0) tin[0] = t;
    if (j < 8) {
      float s, c;
      sincosf(t * (float)(1 << j), &s, &c);
      tin[1 + j] = s;
      tin[9 + j] = c;
    }
    if (j < 32) go[j] = j < 30 ? __ldg(gtout + (size_t)n * 32 + j) : 0.f;
    __syncthreads();
    float h = __ldg(l1.bias + j);
#pragma unroll
    for (int k = 0; k < 17; ++k) h = fmaf(tin[k], __ldg(l1.wt + (size_t)k * l1.Npad + j), h);
    hid[j] = fmaxf(h, 0.f);
    float d = 0.f;
#pragma unroll 6
    for (int o = 0; o < 30; ++o) d = fmaf(__ldg(l2.wt + (size_t)j * l2.Npad + o), go[o], d);
    d = h > 0.f ? d : 0.f;
    dh[j] = d;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 17; ++k) aW1[k] = fmaf(tin[k], d, aW1[k]);
    aB1 += d;
    if (j < 30) {
      const float gj = go[j];
#pragma unroll
      for (int k = 0; k < 64; ++k) aW2[k] = fmaf(hid[k], gj, aW2[k]);
      aB2 += gj;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 17; ++k) atomicAdd(g1.dwt + (size_t)k * 64 + j, aW1[k]);
  atomicAdd(g1.dbias + j, aB1);
  if (j < 30) {
#pragma unroll
    for (int k = 0; k < 64; ++k) atomicAdd(g2.dwt + (size_t)k * 32 + j, aW2[k]);
    atomicAdd(g2.dbias + j, aB2);
  }
}

// ------------------------------------------------------------------------------------------------
// dynamic density / blending MLP backward (a16 / a17), one table set + MLP per launch.
// gout [P] = d loss / d (sigma | blending).  mode 0: density (act), mode 1: blending (sigmoid).
// Outputs: vm grads, packed grads of layer1 (152->64) / layer2 (64->1), gxp [P][3] += d loss / d xyz_prime.

constexpr size_t kDmSmem = (size_t)(152 * kBL + 64 * kBL + 80 * kBL + 2 * kChunkK * 80) * sizeof(float) + sizeof(BScratch);

__global__ void __launch_bounds__(kMlpThreads, 2)
k_dyn_mlp_bwd(dynrf_field_cfg_t cfg, int mode, dynrf_vm_t vm, dynrf_vm_grad_t gvm, dynrf_layer_t l1, dynrf_layer_t l2,
              const float* __restrict__ wb1f /*[64][80]*/, const float* __restrict__ xyz, const float* __restrict__ xyz_prime,
              const uint8_t* __restrict__ valid, const float* __restrict__ tfeat, int P, int S, const float* __restrict__ gout,
              LayerGrad g1, LayerGrad g2, float* __restrict__ gxp) {
  extern __shared__ __align__(16) float smem[];
  float* A = smem;
  float* H = A + 152 * kBL;
  float* dA = H + 64 * kBL;
  float* wbuf = dA + 80 * kBL;
  BScratch& scr = *reinterpret_cast<BScratch*>(wbuf + 2 * kChunkK * 80);
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int ntiles = (P + kBT - 1) / kBT;
  float accW1[10][4];
  wgrad_zero(accW1);
  float accW2 = 0.f, accB1 = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kBT) {
      const int i = tile * kBT + tid;
      const bool in = i < P;
      scr.ray[tid] = in ? i / S : -1;
      scr.idx[tid] = i;
      scr.on[tid] = (in && valid[i] != 0) ? 1 : 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = in ? __ldg(xyz + (size_t)i * 3 + a) : 0.f;
        const float xp = in ? __ldg(xyz_prime + (size_t)i * 3 + a) : 0.f;
        scr.xn[a][tid] = norm_coord(x, nb.mn[a], nb.inv[a]);
        scr.xq[a][tid] = norm_coord(xp, nb.mn[a], nb.inv[a]);
        scr.gx[a][tid] = 0.f;
      }
    }
    __syncthreads();
    fill_pos_time_rows<kBT, kBL>(scr, tfeat, A, 72, tid);
#pragma unroll 1
    for (int si = 0; si < 3; ++si) gather_vm_rows<kBT, kBL>(vm, scr.xq, 1 << si, si * 24, A, tid, scr.on);
    __syncthreads();
    dense_tile<64, kBT, kBL, true>(A, l1, H, wbuf, tid);
    if (tid < kBT) {
      float o[1];
      small_dense<1, kBL>(H, 64, l2, tid, o);
      float g = 0.f;
      if (scr.on[tid]) {
        g = __ldg(gout + scr.idx[tid]);
        if (mode == 0) {
          if (cfg.act == DYNRF_ACT_RELU) g = o[0] > 0.f ? g : 0.f;
          else g *= sigmoidf_(o[0] + cfg.density_shift);
        } else {
          const float sg = sigmoidf_(o[0]);
          g *= sg * (1.0f - sg);
        }
      }
      scr.go[0][tid] = g;
    }
    __syncthreads();
    if (tid < 64) {
      const float* h = H + (size_t)tid * kBL;
      float s = 0.f;
      for (int m = 0; m < kBT; ++m) s = fmaf(h[m], scr.go[0][m], s);
      accW2 += s;
    } else if (tid == 64) {
      float s = 0.f;
      for (int m = 0; m < kBT; ++m) s += scr.go[0][m];
      accW2 += s;  // thread 64 holds db2
    }
    __syncthreads();
    for (int e = tid; e < 64 * kBT; e += kMlpThreads) {
      const int k = e / kBT, m = e % kBT;
      float* h = H + (size_t)k * kBL + m;
      *h = (*h > 0.f) ? __ldg(l2.wt + (size_t)k * l2.Npad) * scr.go[0][m] : 0.f;
    }
    __syncthreads();
    wgrad_acc<10, 4, kBT, kBL>(accW1, A, 152, H, tid);
    if (tid < 64) accB1 += row_sum(H, tid);
    {
      Acc<80, kBT> acc;
      acc_init<80, kBT>(acc, nullptr, tid >> 4);
      acc_gemm<80, kBT, kBL>(acc, H, 64, wb1f, wbuf, tid);
      acc_store<80, kBT, kBL, false>(acc, dA, 72, tid);
    }
#pragma unroll 1
    for (int si = 0; si < 3; ++si) scatter_vm_rows<kBT, kBL>(vm, gvm, scr.xq, 1 << si, si * 24, dA, tid, scr.on, scr.gx);
    __syncthreads();
    if (tid < kBT && scr.on[tid]) {
      const int i = scr.idx[tid];
#pragma unroll
      for (int a = 0; a < 3; ++a) gxp[(size_t)i * 3 + a] += scr.gx[a][tid] * nb.inv[a];
    }
    __syncthreads();
  }
  wgrad_flush(accW1, g1.dwt, 152, 64, tid);
  if (tid < 64) {
    atomicAdd(g1.dbias + tid, accB1);
    atomicAdd(g2.dwt + (size_t)tid * 16, accW2);
  } else if (tid == 64) {
    atomicAdd(g2.dbias, accW2);
  }
}

// ------------------------------------------------------------------------------------------------
// appearance backward (static a7+a12/a13, dynamic a18+a14) on the active list.
// DYN = true : 3 strides, head MLP_Fea_late_view, coordinates = normalize(xyz_prime), gx -> gxp
// DYN = false: 1 stride, head MLP_Fea | MLP_Fea_TimeEmbedding, coordinates = normalize(xyz), gx -> gxyz (optional),
//              gvd [N][3] += d loss / d viewdir (normalised direction)
// smem rows (LD 68): Ac 72 | X 144 | H1 128 | H2 128 | dAc 80 | dF 32 | wbuf | scratch.   dX aliases H2..dAc (208 rows)

constexpr int kAppXRows = 144;
constexpr size_t kAppBwdSmem =
    (size_t)((72 + kAppXRows + 128 + 128 + 80 + 32) * kBL + 2 * kChunkK * 144) * sizeof(float) + sizeof(BScratch);

template <bool DYN, int HEAD>
__global__ void __launch_bounds__(kMlpThreads)
k_app_bwd(dynrf_field_cfg_t cfg, dynrf_vm_t vm, dynrf_vm_grad_t gvm, dynrf_layer_t basis, dynrf_layer_t l1, dynrf_layer_t l2,
          dynrf_layer_t l3, const float* __restrict__ wb2 /*[128][128]*/, const float* __restrict__ wb1 /*[128][Jpad]*/,
          const float* __restrict__ wbB /*[nstr][32][80]*/, const float* __restrict__ rays, const float* __restrict__ xyz,
          const float* __restrict__ xyz_prime, const float* __restrict__ tfeat, int S, const int32_t* __restrict__ active_idx,
          const int32_t* __restrict__ active_count, const float* __restrict__ grgb, LayerGrad gB, LayerGrad g1, LayerGrad g2,
          LayerGrad g3, float* __restrict__ gcoord /*gxp | gxyz | NULL*/, float* __restrict__ gvd /*[N][3] | NULL*/) {
  constexpr int NSTR = DYN ? 3 : 1;
  constexpr int XR = DYN ? 108 : (HEAD == DYNRF_HEAD_FEA ? 140 : 136);       // Kpad of l1
  constexpr int XJ = DYN ? 32 : 144;                                           // data-gradient columns of l1 (padded)
  constexpr int kSin0 = HEAD == DYNRF_HEAD_FEA ? 30 : 27;
  constexpr int kCos0 = kSin0 + 54;
  constexpr int KT1 = (XR + 15) / 16;
  constexpr bool VIEW_LAST = DYN || HEAD != DYNRF_HEAD_FEA;  // last layer input = [h(128), viewdir(3)]
  extern __shared__ __align__(16) float smem[];
  float* Ac = smem;
  float* X = Ac + 72 * kBL;
  float* H1 = X + kAppXRows * kBL;
  float* H2 = H1 + 128 * kBL;
  float* dAc = H2 + 128 * kBL;
  float* dF = dAc + 80 * kBL;
  float* wbuf = dF + 32 * kBL;
  BScratch& scr = *reinterpret_cast<BScratch*>(wbuf + 2 * kChunkK * 144);
  float* dX = H2;  // [XJ][kBL], valid after dH2 is consumed
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int count = *active_count;
  const int ntiles = (count + kBT - 1) / kBT;
  float accW2[8][8], accW1[KT1][8];
  wgrad_zero(accW2);
  wgrad_zero(accW1);
  float accW3[3] = {0.f, 0.f, 0.f}, accB1 = 0.f, accB2 = 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kBT) {
      const int j = tile * kBT + tid;
      const bool in = j < count;
      const int i = in ? __ldg(active_idx + j) : 0;
      scr.idx[tid] = i;
      scr.ray[tid] = in ? i / S : -1;
      scr.on[tid] = in ? 1 : 0;
      float vd[3] = {0.f, 0.f, 0.f};
      if (in) ray_viewdir(rays, i / S, vd);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = in ? __ldg(xyz + (size_t)i * 3 + a) : 0.f;
        scr.xn[a][tid] = norm_coord(x, nb.mn[a], nb.inv[a]);
        if (DYN) {
          const float xp = in ? __ldg(xyz_prime + (size_t)i * 3 + a) : 0.f;
          scr.xq[a][tid] = norm_coord(xp, nb.mn[a], nb.inv[a]);
        } else {
          scr.xq[a][tid] = scr.xn[a][tid];
        }
        scr.xu[a][tid] = vd[a];
        scr.gx[a][tid] = 0.f;
        if (!DYN && HEAD == DYNRF_HEAD_FEA) X[(size_t)(27 + a) * kBL + tid] = vd[a];
      }
      for (int r = (DYN ? 107 : (HEAD == DYNRF_HEAD_FEA ? 138 : 135)); r < XR; ++r) X[(size_t)r * kBL + tid] = 0.f;
    }
    __syncthreads();
    if (DYN) fill_pos_time_rows<kBT, kBL>(scr, tfeat, X, 27, tid);
    // ---- forward recompute
    {
      Acc<32, kBT> fa;
      acc_init<32, kBT>(fa, nullptr, tid >> 4);
#pragma unroll 1
      for (int si = 0; si < NSTR; ++si) {
        gather_vm_rows<kBT, kBL>(vm, scr.xq, 1 << si, 0, Ac, tid, scr.on);
        __syncthreads();
        acc_gemm<32, kBT, kBL>(fa, Ac, 72, basis.wt + (size_t)si * 72 * 32, wbuf, tid);
      }
      acc_store<32, kBT, kBL, false>(fa, X, 27, tid);
    }
    if (!DYN) {
      for (int item = tid; item < 27 * kBT; item += kMlpThreads) {
        const int m = item % kBT, d = item / kBT;
        pe_rows<2, kBL>(X[(size_t)d * kBL + m], X, kSin0 + d * 2, kCos0 + d * 2, m);
      }
      __syncthreads();
    }
    dense_tile<128, kBT, kBL, true>(X, l1, H1, wbuf, tid);
    dense_tile<128, kBT, kBL, true>(H1, l2, H2, wbuf, tid);
    if (tid < kBT) {
      float o[3];
      small_dense<3, kBL>(H2, 128, l3, tid, o);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = o[c];
        if (VIEW_LAST) {
#pragma unroll
          for (int k = 0; k < 3; ++k) v = fmaf(scr.xu[k][tid], __ldg(l3.wt + (size_t)(128 + k) * l3.Npad + c), v);
        }
        const float r = sigmoidf_(v);
        scr.go[c][tid] = scr.on[tid] ? __ldg(grgb + (size_t)scr.idx[tid] * 3 + c) * r * (1.0f - r) : 0.f;
      }
    }
    __syncthreads();
    // ---- last layer
    if (tid < 128) {
      const float* h = H2 + (size_t)tid * kBL;
      for (int m = 0; m < kBT; ++m) {
        const float hv = h[m];
        accW3[0] = fmaf(hv, scr.go[0][m], accW3[0]);
        accW3[1] = fmaf(hv, scr.go[1][m], accW3[1]);
        accW3[2] = fmaf(hv, scr.go[2][m], accW3[2]);
      }
    } else if (tid < 131) {  // viewdir rows (VIEW_LAST) -- harmless extra work otherwise
      if (VIEW_LAST) {
        const int k = tid - 128;
        for (int m = 0; m < kBT; ++m) {
          const float hv = scr.xu[k][m];
          accW3[0] = fmaf(hv, scr.go[0][m], accW3[0]);
          accW3[1] = fmaf(hv, scr.go[1][m], accW3[1]);
          accW3[2] = fmaf(hv, scr.go[2][m], accW3[2]);
        }
      }
    } else if (tid < 134) {  // bias
      for (int m = 0; m < kBT; ++m) accW3[0] += scr.go[tid - 131][m];
    } else if (!DYN && HEAD != DYNRF_HEAD_FEA && gvd && tid >= 192 && tid < 192 + kBT) {
      // d loss / d viewdir through the last layer (MLP_Fea_TimeEmbedding); late_view detaches viewdirs
      const int m = tid - 192;
      if (scr.on[m]) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float v = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) v = fmaf(__ldg(l3.wt + (size_t)(128 + k) * l3.Npad + c), scr.go[c][m], v);
          atomicAdd(gvd + (size_t)scr.ray[m] * 3 + k, v);
        }
      }
    }
    __syncthreads();
    for (int e = tid; e < 128 * kBT; e += kMlpThreads) {
      const int k = e / kBT, m = e % kBT;
      float v = 0.f;
#pragma unroll
      for (int o = 0; o < 3; ++o) v = fmaf(__ldg(l3.wt + (size_t)k * l3.Npad + o), scr.go[o][m], v);
      float* h = H2 + (size_t)k * kBL + m;
      *h = (*h > 0.f) ? v : 0.f;
    }
    __syncthreads();
    // ---- layer 2
    wgrad_acc<8, 8, kBT, kBL>(accW2, H1, 128, H2, tid);
    if (tid < 128) accB2 += row_sum(H2, tid);
    {
      Acc<128, kBT> acc;
      acc_init<128, kBT>(acc, nullptr, tid >> 4);
      acc_gemm<128, kBT, kBL>(acc, H2, 128, wb2, wbuf, tid);
      acc_store<128, kBT, kBL, false, true>(acc, H1, 128, tid, H1);
    }
    // ---- layer 1
    wgrad_acc<KT1, 8, kBT, kBL>(accW1, X, XR, H1, tid);
    if (tid < 128) accB1 += row_sum(H1, tid);
    {
      Acc<XJ, kBT> acc;
      acc_init<XJ, kBT>(acc, nullptr, tid >> 4);
      acc_gemm<XJ, kBT, kBL>(acc, H1, 128, wb1, wbuf, tid);
      acc_store<XJ, kBT, kBL, false>(acc, dX, XJ, tid);
    }
    // ---- d feat (27 rows) : direct rows + positional-encoding chain (static heads)
    for (int item = tid; item < 27 * kBT; item += kMlpThreads) {
      const int m = item % kBT, d = item / kBT;
      float gacc = dX[(size_t)d * kBL + m];
      if (!DYN) {
        float sc = 1.0f;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          const float sn = X[(size_t)(kSin0 + d * 2 + f) * kBL + m], cn = X[(size_t)(kCos0 + d * 2 + f) * kBL + m];
          gacc += sc * (cn * dX[(size_t)(kSin0 + d * 2 + f) * kBL + m] - sn * dX[(size_t)(kCos0 + d * 2 + f) * kBL + m]);
          sc *= 2.0f;
        }
      }
      dF[(size_t)d * kBL + m] = gacc;
    }
    for (int item = tid; item < 5 * kBT; item += kMlpThreads) dF[(size_t)(27 + item / kBT) * kBL + item % kBT] = 0.f;
    if (!DYN && HEAD == DYNRF_HEAD_FEA && gvd && tid < kBT && scr.on[tid]) {
#pragma unroll
      for (int k = 0; k < 3; ++k) atomicAdd(gvd + (size_t)scr.ray[tid] * 3 + k, dX[(size_t)(27 + k) * kBL + tid]);
    }
    __syncthreads();
    // ---- basis + lookups, one stride at a time
#pragma unroll 1
    for (int si = 0; si < NSTR; ++si) {
      gather_vm_rows<kBT, kBL>(vm, scr.xq, 1 << si, 0, Ac, tid, scr.on);
      __syncthreads();
      {
        float accB[5][2];
        wgrad_zero(accB);
        wgrad_acc<5, 2, kBT, kBL>(accB, Ac, 72, dF, tid);
        wgrad_flush(accB, gB.dwt + (size_t)si * 72 * 32, 72, 32, tid);
      }
      {
        Acc<80, kBT> acc;
        acc_init<80, kBT>(acc, nullptr, tid >> 4);
        acc_gemm<80, kBT, kBL>(acc, dF, 32, wbB + (size_t)si * 32 * 80, wbuf, tid);
        acc_store<80, kBT, kBL, false>(acc, dAc, 72, tid);
      }
      scatter_vm_rows<kBT, kBL>(vm, gvm, scr.xq, 1 << si, 0, dAc, tid, scr.on, scr.gx);
      __syncthreads();
    }
    if (gcoord && tid < kBT && scr.on[tid]) {
      const int i = scr.idx[tid];
#pragma unroll
      for (int a = 0; a < 3; ++a) gcoord[(size_t)i * 3 + a] += scr.gx[a][tid] * nb.inv[a];
    }
    __syncthreads();
  }
  wgrad_flush(accW2, g2.dwt, 128, 128, tid);
  wgrad_flush(accW1, g1.dwt, XR, 128, tid);
  if (tid < 128) {
    atomicAdd(g1.dbias + tid, accB1);
    atomicAdd(g2.dbias + tid, accB2);
  }
  if (tid < 131) {
    if (tid < 128 || VIEW_LAST) {
#pragma unroll
      for (int o = 0; o < 3; ++o) atomicAdd(g3.dwt + (size_t)tid * 16 + o, accW3[o]);
    }
  } else if (tid < 134) {
    atomicAdd(g3.dbias + (tid - 131), accW3[0]);
  }
}

}  // namespace dynrf

using namespace dynrf;

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define DYNRF_TRY(x)             \
  do {                           \
    if (int _e = (x)) return _e; \
  } while (0)

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

static inline LayerGrad LG(const dynrf_layer_grad_t* g) { return LayerGrad{g->dwt, g->dbias}; }

extern "C" {

int dynrf_pack_linear_bwd(const float* w, int N, int K, int k0, int k1, int nrows_pad, float* wb, void* stream) {
  if (!w || !wb || N <= 0 || K <= 0 || k0 < 0 || k1 > K || k1 <= k0 || nrows_pad < N || (nrows_pad & 3)) return DYNRF_E_ARG;
  const int Nrows = nrows_pad, Jpad = (k1 - k0 + 15) & ~15;
  k_pack_linear_bwd<<<(Nrows * Jpad + 255) / 256, 256, 0, S_(stream)>>>(w, N, K, k0, k1, Nrows, Jpad, wb);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_alpha_weight_bwd(int N, int S, const float* sigma, const float* dists, const float* gweight, float* gsigma,
                           float* gdists, void* stream) {
  if (!sigma || !dists || !gweight || !gsigma || !gdists || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if ((size_t)S * sizeof(float) > 200 * 1024) return DYNRF_E_UNSUPPORTED;
  const size_t sm = (size_t)S * sizeof(float);
  if (sm > 48 * 1024) DYNRF_TRY(set_smem(k_alpha_weight_bwd, sm));
  k_alpha_weight_bwd<<<N, kRayThreads, sm, S_(stream)>>>(N, S, sigma, dists, gweight, gsigma, gdists);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_composite_bwd(int ray_type, int white_bg, const float* rgb_s, const float* sigma_s, const float* rgb_d,
                        const float* sigma_d, const float* dists, const float* blending, const float* z_vals, const float* rays,
                        int N, int S, const float* g_out_ray, const float* g_w_full, const float* g_w_s, const float* g_w_d,
                        float* g_rgb_s, float* g_sigma_s, float* g_rgb_d, float* g_sigma_d, float* g_dists, float* g_blend,
                        float* g_rays, void* stream) {
  if (!rgb_s || !sigma_s || !rgb_d || !sigma_d || !dists || !blending || !z_vals || !rays || !g_out_ray || !g_rgb_s ||
      !g_sigma_s || !g_rgb_d || !g_sigma_d || !g_dists || !g_blend || !g_rays || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  const size_t sm = (size_t)3 * S * sizeof(float);
  if (sm > 200 * 1024) return DYNRF_E_UNSUPPORTED;
  if (sm > 48 * 1024) DYNRF_TRY(set_smem(k_composite_bwd, sm));
  k_composite_bwd<<<N, kRayThreads, sm, S_(stream)>>>(ray_type, white_bg, rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals,
                                                     rays, N, S, g_out_ray, g_w_full, g_w_s, g_w_d, g_rgb_s, g_sigma_s, g_rgb_d,
                                                     g_sigma_d, g_dists, g_blend, g_rays);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_static_density_bwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_grad_t* gdensity,
                             const float* xyz, const uint8_t* valid, const float* sigma, const float* gsigma, int N, int S,
                             float* gxyz, void* stream) {
  if (density && !dynrf_vm_layout_ok(density)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !density || !gdensity || !xyz || !valid || !sigma || !gsigma || N <= 0 || S <= 0) return DYNRF_E_ARG;
  const int P = N * S;
  k_static_density_bwd<<<(P + 127) / 128, 128, 0, S_(stream)>>>(*cfg, *density, *gdensity, xyz, valid, sigma, gsigma, P, gxyz);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_warp_bwd(const dynrf_field_cfg_t* cfg, const float* xyz, int N, int S, const float* tfeat, const dynrf_layer_t* l3,
                   const dynrf_layer_t* l4, const dynrf_layer_t* l5, const float* wb4, const float* wb3t, const float* gxp,
                   const dynrf_layer_grad_t* g3, const dynrf_layer_grad_t* g4, const dynrf_layer_grad_t* g5, float* gtout,
                   const float* H1T, const float* H2T, int ld, void* stream) {
  if (!cfg || !xyz || !tfeat || !l3 || !l4 || !l5 || !wb4 || !wb3t || !gxp || !g3 || !g4 || !g5 || !gtout || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  if ((H1T != nullptr) != (H2T != nullptr)) return DYNRF_E_ARG;
  if (H1T && ld < ((N * S + kBT - 1) / kBT) * kBT) return DYNRF_E_ARG;
  DYNRF_TRY(set_smem(k_warp_bwd, kWarpBwdSmem));
  const int P = N * S;
  const int grid = min((P + kBT - 1) / kBT, 2 * sm_count());
  k_warp_bwd<<<grid, kMlpThreads, kWarpBwdSmem, S_(stream)>>>(*cfg, xyz, P, S, tfeat, *l3, *l4, *l5, wb4, wb3t, gxp, LG(g3), LG(g4),
                                                              LG(g5), gtout, H1T, H2T, ld);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_time_embed_bwd(const float* ts, int N, const dynrf_layer_t* l1, const dynrf_layer_t* l2, const float* gtout,
                         const dynrf_layer_grad_t* g1, const dynrf_layer_grad_t* g2, void* stream) {
  if (!ts || !l1 || !l2 || !gtout || !g1 || !g2 || N <= 0) return DYNRF_E_ARG;
  // one 64-thread block per ~14 rays (296 blocks = 2 per SM): with 64 blocks the kernel was a 160 us latency chain (ncu: 1.5 % SM busy)
  k_time_embed_bwd<<<min(N, 296), 64, 0, S_(stream)>>>(ts, N, *l1, *l2, gtout, LG(g1), LG(g2));
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_mlp_bwd(const dynrf_field_cfg_t* cfg, int mode, const dynrf_vm_t* vm, const dynrf_vm_grad_t* gvm,
                      const dynrf_layer_t* l1, const dynrf_layer_t* l2, const float* wb1f, const float* xyz, const float* xyz_prime,
                      const uint8_t* valid, const float* tfeat, int N, int S, const float* gout, const dynrf_layer_grad_t* g1,
                      const dynrf_layer_grad_t* g2, float* gxp, void* stream) {
  if (vm && !dynrf_vm_layout_ok(vm)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !vm || !gvm || !l1 || !l2 || !wb1f || !xyz || !xyz_prime || !valid || !tfeat || !gout || !g1 || !g2 || !gxp ||
      N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  if (vm->C[0] != 16 || vm->C[1] != 4 || vm->C[2] != 4) return DYNRF_E_UNSUPPORTED;
  DYNRF_TRY(set_smem(k_dyn_mlp_bwd, kDmSmem));
  const int P = N * S;
  const int grid = min((P + kBT - 1) / kBT, 2 * sm_count());
  k_dyn_mlp_bwd<<<grid, kMlpThreads, kDmSmem, S_(stream)>>>(*cfg, mode, *vm, *gvm, *l1, *l2, wb1f, xyz, xyz_prime, valid, tfeat, P, S,
                                                            gout, LG(g1), LG(g2), gxp);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_app_bwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_vm_grad_t* gapp, const dynrf_layer_t* basis,
                      const dynrf_layer_t* l1, const dynrf_layer_t* l2, const dynrf_layer_t* l3, const float* wb2, const float* wb1f,
                      const float* wbB, const float* rays, const float* xyz, const float* xyz_prime, const float* tfeat, int N, int S,
                      const int32_t* active_idx, const int32_t* active_count, const float* grgb, const dynrf_layer_grad_t* gB,
                      const dynrf_layer_grad_t* g1, const dynrf_layer_grad_t* g2, const dynrf_layer_grad_t* g3, float* gxp,
                      void* stream) {
  if (app && !dynrf_vm_layout_ok(app)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !app || !gapp || !basis || !l1 || !l2 || !l3 || !wb2 || !wb1f || !wbB || !rays || !xyz || !xyz_prime || !tfeat ||
      !active_idx || !active_count || !grgb || !gB || !g1 || !g2 || !g3 || !gxp || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  if (app->C[0] != 48 || app->C[1] != 12 || app->C[2] != 12) return DYNRF_E_UNSUPPORTED;
  auto kern = k_app_bwd<true, DYNRF_HEAD_FEA_LATE_VIEW>;
  DYNRF_TRY(set_smem(kern, kAppBwdSmem));
  const int grid = min((N * S + kBT - 1) / kBT, sm_count());
  kern<<<grid, kMlpThreads, kAppBwdSmem, S_(stream)>>>(*cfg, *app, *gapp, *basis, *l1, *l2, *l3, wb2, wb1f, wbB, rays, xyz, xyz_prime,
                                                       tfeat, S, active_idx, active_count, grgb, LG(gB), LG(g1), LG(g2), LG(g3), gxp,
                                                       nullptr);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_static_app_bwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_vm_grad_t* gapp, const dynrf_layer_t* basis,
                         const dynrf_layer_t* l1, const dynrf_layer_t* l2, const dynrf_layer_t* l3, int head_mode, const float* wb2,
                         const float* wb1, const float* wbB, const float* rays, const float* xyz, int N, int S,
                         const int32_t* active_idx, const int32_t* active_count, const float* grgb, const dynrf_layer_grad_t* gB,
                         const dynrf_layer_grad_t* g1, const dynrf_layer_grad_t* g2, const dynrf_layer_grad_t* g3, float* gxyz,
                         float* gviewdir, void* stream) {
  if (app && !dynrf_vm_layout_ok(app)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !app || !gapp || !basis || !l1 || !l2 || !l3 || !wb2 || !wb1 || !wbB || !rays || !xyz || !active_idx ||
      !active_count || !grgb || !gB || !g1 || !g2 || !g3 || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  if (app->C[0] != 48 || app->C[1] != 12 || app->C[2] != 12) return DYNRF_E_UNSUPPORTED;
  const int grid = min((N * S + kBT - 1) / kBT, sm_count());
  if (head_mode == DYNRF_HEAD_FEA) {
    auto kern = k_app_bwd<false, DYNRF_HEAD_FEA>;
    DYNRF_TRY(set_smem(kern, kAppBwdSmem));
    kern<<<grid, kMlpThreads, kAppBwdSmem, S_(stream)>>>(*cfg, *app, *gapp, *basis, *l1, *l2, *l3, wb2, wb1, wbB, rays, xyz, nullptr,
                                                         nullptr, S, active_idx, active_count, grgb, LG(gB), LG(g1), LG(g2), LG(g3),
                                                         gxyz, gviewdir);
  } else if (head_mode == DYNRF_HEAD_FEA_TIMEEMB) {
    auto kern = k_app_bwd<false, DYNRF_HEAD_FEA_TIMEEMB>;
    DYNRF_TRY(set_smem(kern, kAppBwdSmem));
    kern<<<grid, kMlpThreads, kAppBwdSmem, S_(stream)>>>(*cfg, *app, *gapp, *basis, *l1, *l2, *l3, wb2, wb1, wbB, rays, xyz, nullptr,
                                                         nullptr, S, active_idx, active_count, grgb, LG(gB), LG(g1), LG(g2), LG(g3),
                                                         gxyz, gviewdir);
  } else {
    return DYNRF_E_UNSUPPORTED;
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
