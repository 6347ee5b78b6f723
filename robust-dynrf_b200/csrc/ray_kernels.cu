// Per-ray kernels: sample generation (a1-a3), static density + sigma->alpha + transmittance
// scan (a6+a8+a9, the HBM-roofline kernel), sigma->weight for the dynamic field, and the
// two-field compositing of renderer.raw2outputs (a19).
//
// Mapping: one CTA of 128 threads per ray, threads over samples in chunks of 128; the
// transmittance scans are warp shuffles + a 4-entry shared-memory carry.
#include "dynrf_common.cuh"

namespace dynrf {

// ------------------------------------------------------------------------------------------------
// depth vectors (one per batch, shared by all rays -- models/tensorBase.py:490-492, 530-548)

// torch.linspace (CPU and CUDA kernels): symmetric evaluation around the midpoint.
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
  if (steps == 1) return start;
  const float step = __fdiv_rn(__fsub_rn(end, start), (float)(steps - 1));
  if (i < steps / 2) return __fadd_rn(start, __fmul_rn(step, (float)i));
  return __fsub_rn(end, __fmul_rn(step, (float)(steps - i - 1)));
}

__global__ void k_zvec_ndc(float near_, float far_, int S, const float* __restrict__ jitter, float* __restrict__ zvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  float z = linspace_at(near_, far_, S, i);
  // Python computes (far - near) / N_samples in double, then the tensor op rounds the product to fp32
  if (jitter) z = __fadd_rn(z, __fmul_rn(jitter[i], (float)(((double)far_ - (double)near_) / (double)S)));
  zvec[i] = z;
}

__global__ void k_zvec_contract(float near_, float far_, int S, const float* __restrict__ jin, const float* __restrict__ jout,
                                float* __restrict__ zvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  const int n_in = S - S / 2, n_out = S / 2;
  if (i < n_in) {
    // midpoints of linspace(near, 2, n_in+1), jitter on every edge but the last
    const float w = (float)((2.0 - (double)near_) / (double)n_in);
    float e0 = linspace_at(near_, 2.0f, n_in + 1, i);
    float e1 = linspace_at(near_, 2.0f, n_in + 1, i + 1);
    if (jin) {
      e0 = __fadd_rn(e0, __fmul_rn(jin[i], w));
      if (i + 1 < n_in) e1 = __fadd_rn(e1, __fmul_rn(jin[i + 1], w));
    }
    zvec[i] = __fmul_rn(__fadd_rn(e1, e0), 0.5f);
  } else {
    // r = flip(arange(n_out+1) (+jitter on all but last)); mid = (r[1:] + r[:-1]) / 2
    const int j = i - n_in;          // index into the flipped midpoints
    const int a = n_out - j;         // r_flipped[j]   = rng[n_out - j]
    const int b = n_out - j - 1;     // r_flipped[j+1] = rng[n_out - j - 1]
    float ra = (float)a, rb = (float)b;
    if (jout) {
      if (a < n_out) ra = __fadd_rn(ra, jout[a]);
      rb = __fadd_rn(rb, jout[b]);
    }
    const float mid = __fmul_rn(__fadd_rn(rb, ra), 0.5f);
    // 1 / (1/far + (1/2 - 1/far) * r / n_out): the scalar coefficients are Python doubles
    const float c0 = (float)(1.0 / (double)far_);
    const float c1 = (float)(0.5 - 1.0 / (double)far_);
    const float den = __fadd_rn(c0, __fdiv_rn(__fmul_rn(c1, mid), (float)n_out));
    zvec[i] = __fdiv_rn(1.0f, den);
  }
}

// ------------------------------------------------------------------------------------------------
// sample point generation shared by k_sample_points and the generate-in-kernel density path

struct RayCtx {
  float o[3], d[3];
  float norm;  // |d|
};

__device__ __forceinline__ RayCtx load_ray(const float* __restrict__ rays, int n) {
  RayCtx r;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    r.o[i] = __ldg(rays + (size_t)n * 6 + i);
    r.d[i] = __ldg(rays + (size_t)n * 6 + 3 + i);
  }
  // torch.norm(viewdirs, dim=-1): sqrt(x^2 + y^2 + z^2)
  r.norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.d[0], r.d[0]), __fmul_rn(r.d[1], r.d[1])), __fmul_rn(r.d[2], r.d[2])));
  return r;
}

__device__ __forceinline__ bool gen_point(const dynrf_field_cfg_t& c, const RayCtx& r, float z, float (&p)[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) p[i] = __fadd_rn(r.o[i], __fmul_rn(r.d[i], z));
  if (c.ray_type == DYNRF_RAY_NDC) {
    bool out = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) out |= (c.aabb_min[i] > p[i]) | (p[i] > c.aabb_max[i]);
    return !out;
  }
  // L-inf contraction, models/tensorBase.py:552-556
  const float linf = fmaxf(fmaxf(fabsf(p[0]), fabsf(p[1])), fabsf(p[2]));
  if (linf > 1.0f) {
    const float k = __fsub_rn(2.0f, __fdiv_rn(1.0f, linf));
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = __fmul_rn(k, __fdiv_rn(p[i], linf));
  }
  return true;
}

__global__ void k_sample_points(dynrf_field_cfg_t cfg, const float* __restrict__ rays, int N, int S,
                                const float* __restrict__ zvec, float* __restrict__ xyz, float* __restrict__ z_vals,
                                uint8_t* __restrict__ valid) {
  const int n = blockIdx.x;
  const RayCtx r = load_ray(rays, n);
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float z = __ldg(zvec + s);
    float p[3];
    const bool v = gen_point(cfg, r, z, p);
    const size_t i = (size_t)n * S + s;
    xyz[i * 3 + 0] = p[0];
    xyz[i * 3 + 1] = p[1];
    xyz[i * 3 + 2] = p[2];
    z_vals[i] = z;
    valid[i] = v ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------
// block-level exclusive running product with carry across chunks (128 threads = 4 warps)

struct ScanSmem {
  float wtot[3][4];
};

// returns the exclusive prefix product for this thread's factor f; updates carry (all threads agree)
__device__ __forceinline__ float block_excl_prod(float f, float& carry, float* wtot /*[4]*/, int lane, int wid) {
  const float incl = warp_incl_prod(f, lane);
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
  if (lane == 31) wtot[wid] = incl;
  __syncthreads();
  float pre = carry;
  float total = carry;
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

// append the active samples of this chunk to the global list; one atomic per warp
__device__ __forceinline__ void append_active(bool active, int flat_idx, int32_t* __restrict__ list, int32_t* __restrict__ count,
                                              int lane) {
  const unsigned bal = __ballot_sync(0xffffffffu, active);
  if (bal == 0) return;
  int base = 0;
  if (lane == 0) base = atomicAdd(count, __popc(bal));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (active) list[base + __popc(bal & ((1u << lane) - 1u))] = flat_idx;
}

// ------------------------------------------------------------------------------------------------
// a6+a8+a9: static density lookup + act + alpha + T scan + weights + app-mask compaction

template <bool GEN>
__global__ void __launch_bounds__(kRayThreads)
k_static_sigma_weight(dynrf_field_cfg_t cfg, dynrf_vm_t vm, const float* __restrict__ rays, int N, int S,
                      const float* __restrict__ xyz, const float* __restrict__ z_vals, const uint8_t* __restrict__ valid,
                      const float* __restrict__ zvec, float* __restrict__ sigma, float* __restrict__ weight,
                      float* __restrict__ dists, int32_t* __restrict__ active_idx, int32_t* __restrict__ active_count,
                      float* __restrict__ rgb_zero) {
  __shared__ ScanSmem sm;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const RayCtx r = load_ray(rays, n);
  const AabbNorm nb = make_norm(cfg);
  float carry = 1.0f;
  for (int s0 = 0; s0 < S; s0 += kRayThreads) {
    const int s = s0 + threadIdx.x;
    const bool in = s < S;
    const size_t i = (size_t)n * S + (in ? s : 0);
    float p[3] = {0.f, 0.f, 0.f};
    float z = 0.f, znext = 0.f;
    bool v = false;
    if (in) {
      if (GEN) {
        z = __ldg(zvec + s);
        znext = (s + 1 < S) ? __ldg(zvec + s + 1) : z;
        v = gen_point(cfg, r, z, p);
      } else {
        z = __ldg(z_vals + i);
        znext = (s + 1 < S) ? __ldg(z_vals + i + 1) : z;
        p[0] = __ldg(xyz + i * 3 + 0);
        p[1] = __ldg(xyz + i * 3 + 1);
        p[2] = __ldg(xyz + i * 3 + 2);
        v = valid[i] != 0;
      }
    }
    float sg = 0.f;
    if (v) {
      float xn[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) xn[a] = norm_coord(p[a], nb.mn[a], nb.inv[a]);
      float feat = 0.f;
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        int u, w, l;
        plane_axes(pl, u, w, l);
        const Tap1 tu = make_tap(xn[u], vm.W[pl], 1);
        const Tap1 tv = make_tap(xn[w], vm.H[pl], 1);
        const Tap1 tl = make_tap(xn[l], vm.L[pl], 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < vm.C[pl]; c += 4) {
          const float4 g = vm_group(vm.plane[pl], vm.line[pl], vm.C[pl], plane_st(vm, pl), tu, tv, tl, c);
          acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
        }
        feat += (acc.x + acc.y) + (acc.z + acc.w);
      }
      sg = density_act(feat, cfg.act, cfg.density_shift);
    }
    // dists = cat(z[1:] - z[:-1], 0) * |d|, then * distance_scale (models/tensorBase.py:719-724, 772)
    const float dz = (in && s + 1 < S) ? __fsub_rn(znext, z) : 0.f;
    const float dist = __fmul_rn(__fmul_rn(dz, r.norm), cfg.distance_scale);
    const float alpha = in ? __fsub_rn(1.0f, expf(-__fmul_rn(sg, dist))) : 0.f;
    const float f = in ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
    const float T = block_excl_prod(f, carry, sm.wtot[0], lane, wid);
    const float wgt = alpha * T;
    const bool active = in && (wgt > cfg.weight_thres);
    if (in) {
      sigma[i] = sg;
      weight[i] = wgt;
      dists[i] = dist;
      if (rgb_zero && !active) {
        rgb_zero[i * 3 + 0] = 0.f;
        rgb_zero[i * 3 + 1] = 0.f;
        rgb_zero[i * 3 + 2] = 0.f;
      }
    }
    append_active(active, (int)i, active_idx, active_count, lane);
  }
}

// sigma known (dynamic field): weights + compaction only
__global__ void __launch_bounds__(kRayThreads)
k_alpha_weight(dynrf_field_cfg_t cfg, const float* __restrict__ rays, int N, int S, const float* __restrict__ sigma,
               const float* __restrict__ z_vals, float* __restrict__ weight, float* __restrict__ dists,
               int32_t* __restrict__ active_idx, int32_t* __restrict__ active_count, float* __restrict__ rgb_zero) {
  __shared__ ScanSmem sm;
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const RayCtx r = load_ray(rays, n);
  float carry = 1.0f;
  for (int s0 = 0; s0 < S; s0 += kRayThreads) {
    const int s = s0 + threadIdx.x;
    const bool in = s < S;
    const size_t i = (size_t)n * S + (in ? s : 0);
    const float z = in ? __ldg(z_vals + i) : 0.f;
    const float znext = (in && s + 1 < S) ? __ldg(z_vals + i + 1) : z;
    const float sg = in ? __ldg(sigma + i) : 0.f;
    const float dz = (in && s + 1 < S) ? __fsub_rn(znext, z) : 0.f;
    const float dist = __fmul_rn(__fmul_rn(dz, r.norm), cfg.distance_scale);
    const float alpha = in ? __fsub_rn(1.0f, expf(-__fmul_rn(sg, dist))) : 0.f;
    const float f = in ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
    const float T = block_excl_prod(f, carry, sm.wtot[0], lane, wid);
    const float wgt = alpha * T;
    const bool active = in && (wgt > cfg.weight_thres);
    if (in) {
      weight[i] = wgt;
      dists[i] = dist;
      if (rgb_zero && !active) {
        rgb_zero[i * 3 + 0] = 0.f;
        rgb_zero[i * 3 + 1] = 0.f;
        rgb_zero[i * 3 + 2] = 0.f;
      }
    }
    append_active(active, (int)i, active_idx, active_count, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// a19: renderer.raw2outputs

constexpr int kNumSums = 13;  // rgb_f3 depth_f acc_f | rgb_s3 depth_s acc_s | dyn | tot_d | (pad)

__device__ __forceinline__ void block_sum_vec(float (&v)[kNumSums], float (*red)[kNumSums], int lane, int wid) {
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) red[wid][k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) v[k] = (red[0][k] + red[1][k]) + (red[2][k] + red[3][k]);
  __syncthreads();
}

__global__ void __launch_bounds__(kRayThreads)
k_composite(int ray_type, int white_bg, const float* __restrict__ rgb_s, const float* __restrict__ sigma_s,
            const float* __restrict__ rgb_d, const float* __restrict__ sigma_d, const float* __restrict__ dists,
            const float* __restrict__ blending, const float* __restrict__ z_vals, const float* __restrict__ rays, int N, int S,
            float* __restrict__ out_ray, float* __restrict__ w_full, float* __restrict__ w_s, float* __restrict__ w_d) {
  __shared__ ScanSmem sm;
  __shared__ float red[4][kNumSums];
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float cd = 1.f, cs = 1.f, cf = 1.f;
  float acc[kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) acc[k] = 0.f;

  for (int s0 = 0; s0 < S; s0 += kRayThreads) {
    const int s = s0 + threadIdx.x;
    const bool in = s < S;
    const size_t i = (size_t)n * S + (in ? s : 0);
    float ad = 0.f, as = 0.f, b = 0.f, z = 0.f;
    float cs3[3] = {0.f, 0.f, 0.f}, cd3[3] = {0.f, 0.f, 0.f};
    if (in) {
      const float dist = __ldg(dists + i);
      ad = __fsub_rn(1.0f, expf(-__fmul_rn(__ldg(sigma_d + i), dist)));
      as = __fsub_rn(1.0f, expf(-__fmul_rn(__ldg(sigma_s + i), dist)));
      b = __ldg(blending + i);
      z = __ldg(z_vals + i);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        cs3[c] = __ldg(rgb_s + i * 3 + c);
        cd3[c] = __ldg(rgb_d + i * 3 + c);
      }
    }
    const float omb = __fsub_rn(1.0f, b);
    const float fd = in ? __fadd_rn(__fsub_rn(1.0f, ad), 1e-10f) : 1.f;
    const float fs = in ? __fadd_rn(__fsub_rn(1.0f, as), 1e-10f) : 1.f;
    const float ff = in ? __fadd_rn(__fmul_rn(__fsub_rn(1.0f, __fmul_rn(ad, b)), __fsub_rn(1.0f, __fmul_rn(as, omb))), 1e-10f) : 1.f;
    const float Td = block_excl_prod(fd, cd, sm.wtot[0], lane, wid);
    const float Ts = block_excl_prod(fs, cs, sm.wtot[1], lane, wid);
    const float Tf = block_excl_prod(ff, cf, sm.wtot[2], lane, wid);
    if (in) {
      const float wd = ad * Td;
      const float ws = as * Ts;
      const float kd = Tf * ad * b;       // (T_full * alpha_d * blending)
      const float ks = Tf * as * omb;     // (T_full * alpha_s * (1 - blending))
      const float wf = (ad * b + as * omb) * Tf;
      w_d[i] = wd;  // un-normalised; rescaled in pass 2
      w_s[i] = ws;
      w_full[i] = wf;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        acc[c] += kd * cd3[c] + ks * cs3[c];
        acc[5 + c] += ws * cs3[c];
      }
      acc[3] += wf * z;
      acc[4] += wf;
      acc[8] += ws * z;
      acc[9] += ws;
      acc[10] += wf * b;
      acc[11] += wd;
    }
  }
  block_sum_vec(acc, red, lane, wid);
  const float denom = __fadd_rn(acc[11], 1e-10f);

  // pass 2: normalise w_d (renderer.py:252) and accumulate the dynamic-only maps
  float d5[kNumSums];
#pragma unroll
  for (int k = 0; k < kNumSums; ++k) d5[k] = 0.f;
  for (int s = threadIdx.x; s < S; s += kRayThreads) {
    const size_t i = (size_t)n * S + s;
    const float wd = __fdiv_rn(w_d[i], denom);
    w_d[i] = wd;
#pragma unroll
    for (int c = 0; c < 3; ++c) d5[c] += wd * __ldg(rgb_d + i * 3 + c);
    d5[3] += wd * __ldg(z_vals + i);
    d5[4] += wd;
  }
  block_sum_vec(d5, red, lane, wid);

  if (threadIdx.x == 0) {
    float rf[3] = {acc[0], acc[1], acc[2]}, rs[3] = {acc[5], acc[6], acc[7]}, rd[3] = {d5[0], d5[1], d5[2]};
    float depth_f = acc[3], acc_f = acc[4], depth_s = acc[8], acc_s = acc[9], depth_d = d5[3], acc_d = d5[4];
    if (white_bg) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        rd[c] += (1.0f - acc_d);
        rs[c] += (1.0f - acc_s);
        rf[c] += fmaxf(1.0f - acc_f, 0.f);
      }
    }
    float bg = 0.f;
    bool has_bg = false;
    if (ray_type == DYNRF_RAY_NDC) {
      bg = __fadd_rn(__ldg(rays + (size_t)n * 6 + 2), __ldg(rays + (size_t)n * 6 + 5));
      has_bg = true;
    } else if (ray_type == DYNRF_RAY_CONTRACT) {
      bg = 256.0f;
      has_bg = true;
    }
    if (has_bg) {
      depth_d += (1.0f - acc_d) * bg;
      depth_s += (1.0f - acc_s) * bg;
      depth_f += fmaxf(1.0f - acc_f, 0.f) * bg;
    }
    float* o = out_ray + (size_t)n * DYNRF_RAY_OUT;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = fminf(fmaxf(rf[c], 0.f), 1.f);
      o[5 + c] = fminf(fmaxf(rs[c], 0.f), 1.f);
      o[10 + c] = fminf(fmaxf(rd[c], 0.f), 1.f);
    }
    o[3] = depth_f;
    o[4] = acc_f;
    o[8] = depth_s;
    o[9] = acc_s;
    o[13] = depth_d;
    o[14] = acc_d;
    o[15] = acc[10];  // dynamicness (+ relu(1-acc)*0, renderer.py:298-299)
  }
}

}  // namespace dynrf

using namespace dynrf;

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int check_vm(const dynrf_vm_t* vm) {
  for (int i = 0; i < 3; ++i) {
    if (!vm->plane[i] || !vm->line[i]) return DYNRF_E_ARG;
    if (vm->C[i] <= 0 || (vm->C[i] & 3)) return DYNRF_E_UNSUPPORTED;
    if (vm->H[i] < 1 || vm->W[i] < 1 || vm->L[i] < 1) return DYNRF_E_ARG;
    if (((uintptr_t)vm->plane[i] & 15) || ((uintptr_t)vm->line[i] & 15)) return DYNRF_E_ALIGN;
  }
  return 0;
}

extern "C" {

int dynrf_zvec_ndc(float near_, float far_, int S, const float* jitter, float* zvec, void* stream) {
  if (!zvec || S <= 0) return DYNRF_E_ARG;
  k_zvec_ndc<<<(S + 127) / 128, 128, 0, S_(stream)>>>(near_, far_, S, jitter, zvec);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_zvec_contract(float near_, float far_, int S, const float* jin, const float* jout, float* zvec, void* stream) {
  if (!zvec || S <= 1) return DYNRF_E_ARG;
  k_zvec_contract<<<(S + 127) / 128, 128, 0, S_(stream)>>>(near_, far_, S, jin, jout, zvec);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_sample_points(const dynrf_field_cfg_t* cfg, const float* rays, int N, int S, const float* zvec, float* xyz,
                        float* z_vals, uint8_t* valid, void* stream) {
  if (!cfg || !rays || !zvec || !xyz || !z_vals || !valid || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if (cfg->ray_type != DYNRF_RAY_NDC && cfg->ray_type != DYNRF_RAY_CONTRACT) return DYNRF_E_UNSUPPORTED;
  k_sample_points<<<N, kRayThreads, 0, S_(stream)>>>(*cfg, rays, N, S, zvec, xyz, z_vals, valid);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_static_sigma_weight_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const float* rays, int N, int S,
                                  const float* xyz, const float* z_vals, const uint8_t* valid, const float* zvec,
                                  float* sigma, float* weight, float* dists, int32_t* active_idx, int32_t* active_count,
                                  float* rgb_zero, void* stream) {
  if (density && !dynrf_vm_layout_ok(density)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !density || !rays || !sigma || !weight || !dists || !active_idx || !active_count || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  if (int e = check_vm(density)) return e;
  if (xyz) {
    if (!z_vals || !valid) return DYNRF_E_ARG;
    k_static_sigma_weight<false><<<N, kRayThreads, 0, S_(stream)>>>(*cfg, *density, rays, N, S, xyz, z_vals, valid, zvec, sigma,
                                                                    weight, dists, active_idx, active_count, rgb_zero);
  } else {
    if (!zvec) return DYNRF_E_ARG;
    k_static_sigma_weight<true><<<N, kRayThreads, 0, S_(stream)>>>(*cfg, *density, rays, N, S, nullptr, nullptr, nullptr, zvec,
                                                                   sigma, weight, dists, active_idx, active_count, rgb_zero);
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_alpha_weight_fwd(const dynrf_field_cfg_t* cfg, const float* rays, int N, int S, const float* sigma,
                           const float* z_vals, float* weight, float* dists, int32_t* active_idx, int32_t* active_count,
                           float* rgb_zero, void* stream) {
  if (!cfg || !rays || !sigma || !z_vals || !weight || !dists || !active_idx || !active_count || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  k_alpha_weight<<<N, kRayThreads, 0, S_(stream)>>>(*cfg, rays, N, S, sigma, z_vals, weight, dists, active_idx, active_count,
                                                    rgb_zero);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_composite_fwd(int ray_type, int white_bg, const float* rgb_s, const float* sigma_s, const float* rgb_d,
                        const float* sigma_d, const float* dists, const float* blending, const float* z_vals,
                        const float* rays, int N, int S, float* out_ray, float* w_full, float* w_s, float* w_d,
                        void* stream) {
  if (!rgb_s || !sigma_s || !rgb_d || !sigma_d || !dists || !blending || !z_vals || !rays || !out_ray || !w_full || !w_s ||
      !w_d || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  k_composite<<<N, kRayThreads, 0, S_(stream)>>>(ray_type, white_bg, rgb_s, sigma_s, rgb_d, sigma_d, dists, blending, z_vals,
                                                 rays, N, S, out_ray, w_full, w_s, w_d);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
