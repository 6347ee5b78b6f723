// Kernels for the rows next to the ray-march path:
//   a4  TensorBase.sample_ray                      models/tensorBase.py:501-522
//   f1  fused ray generation + its backward        train.py:96-103, 1062-1077; dataLoader/ray_utils.py:53-90, 115-140
//   f2  expected 3-D point along a ray (the S-reduction of renderer.render_3d_point, renderer.py:1335-1338) fwd + bwd
//   f3  total-variation of every plane / line of a VM set in ONE launch (utils.py:157-181, models/tensoRF.py:100-116,
//       419-444) and density_L1 / blending_L1 WITHOUT the dense C x Gy x Gx x Gz product (models/tensoRF.py:80-98, 378-416);
//       both produce value and gradient in a single streaming pass.
#include "dynrf_common.cuh"

namespace dynrf {

// ------------------------------------------------------------------------------------------------ helpers

__device__ __forceinline__ void red_add4g(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// block-wide sum of a double (all threads call; result valid in thread 0)
template <int NT>
__device__ __forceinline__ double block_sum_d(double v, double* sm /*[NT/32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < NT / 32; ++i) t += sm[i];
  __syncthreads();
  return t;
}

// ------------------------------------------------------------------------------------------------ a4: sample_ray

__global__ void k_sample_ray(dynrf_field_cfg_t cfg, const float* __restrict__ rays, int N, int S, float step,
                             const float* __restrict__ jitter /*[N] or null*/, float* __restrict__ xyz, float* __restrict__ z_vals,
                             uint8_t* __restrict__ valid) {
  const int n = blockIdx.x;
  float o[3], d[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o[i] = __ldg(rays + (size_t)n * 6 + i);
    d[i] = __ldg(rays + (size_t)n * 6 + 3 + i);
  }
  // t_min = max_i min((aabb1 - o) / vec, (aabb0 - o) / vec), clamped to [near, far]   (vec: d with zeros replaced by 1e-6)
  float tmin = -INFINITY;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float vec = d[i] == 0.f ? 1e-6f : d[i];
    const float a = __fdiv_rn(__fsub_rn(cfg.aabb_max[i], o[i]), vec), b = __fdiv_rn(__fsub_rn(cfg.aabb_min[i], o[i]), vec);
    tmin = fmaxf(tmin, fminf(a, b));
  }
  tmin = fminf(fmaxf(tmin, cfg.near_), cfg.far_);
  const float jit = jitter ? __ldg(jitter + n) : 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float rng = jitter ? __fadd_rn((float)s, jit) : (float)s;
    const float z = __fadd_rn(tmin, __fmul_rn(step, rng));
    bool out = false;
    const size_t i = (size_t)n * S + s;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float p = __fadd_rn(o[a], __fmul_rn(d[a], z));
      out |= (cfg.aabb_min[a] > p) | (p > cfg.aabb_max[a]);
      xyz[i * 3 + a] = p;
    }
    z_vals[i] = z;
    valid[i] = out ? 0 : 1;
  }
}

// ------------------------------------------------------------------------------------------------ f1: ray generation

struct RayGenCam {
  float R[3][3], t[3];
};

__device__ __forceinline__ RayGenCam load_cam(const float* __restrict__ mtx, int view) {
  RayGenCam c;
  const float* m = mtx + (size_t)view * 12;  // [3][4] row-major
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int j = 0; j < 3; ++j) c.R[k][j] = __ldg(m + k * 4 + j);
    c.t[k] = __ldg(m + k * 4 + 3);
  }
  return c;
}

// ids [B] (int64) -> rays [B][6].  ndc != 0: ndc_rays_blender2 with `near`.
// pixel of ray b: either decoded from the flat id (ids2pixel, train.py:96-103; pixel centre = index + 0.5) or given explicitly as
// sub-pixel coordinates uv[b] = (x, y) with the camera index view[b] (the flow-displaced rays of train.py:1433-1455, 1960-1984)
__device__ __forceinline__ void raygen_pixel(const long long* __restrict__ ids, const float* __restrict__ uv, const long long* __restrict__ view_ids,
                                             int b, int H, int W, float& i, float& j, int& view) {
  if (uv) {
    i = __ldg(uv + (size_t)b * 2);
    j = __ldg(uv + (size_t)b * 2 + 1);
    view = (int)view_ids[b];
  } else {
    const long long id = ids[b];
    i = (float)(int)(id % W) + 0.5f;
    j = (float)(int)((id / W) % H) + 0.5f;
    view = (int)(id / ((long long)W * H));
  }
}

__global__ void k_raygen_fwd(const long long* __restrict__ ids, const float* __restrict__ uv, const long long* __restrict__ view_ids, int B,
                             const float* __restrict__ mtx /*[F][3][4]*/, const float* __restrict__ focal_p, int H, int W, int ndc,
                             float near_, float* __restrict__ rays) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float i, j;
  int view;
  raygen_pixel(ids, uv, view_ids, b, H, W, i, j, view);
  const float f = __ldg(focal_p);
  const RayGenCam c = load_cam(mtx, view);
  // get_ray_directions_lean: ((i - W/2) / f, -(j - H/2) / f, -1)
  const float dir[3] = {__fdiv_rn(__fsub_rn(i, (float)(W / 2.0)), f), -__fdiv_rn(__fsub_rn(j, (float)(H / 2.0)), f), -1.0f};
  float rd[3], ro[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    rd[k] = c.R[k][0] * dir[0] + c.R[k][1] * dir[1] + c.R[k][2] * dir[2];
    ro[k] = c.t[k];
  }
  float* out = rays + (size_t)b * 6;
  if (!ndc) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      out[k] = ro[k];
      out[3 + k] = rd[k];
    }
    return;
  }
  const float tt = -__fdiv_rn(__fadd_rn(near_, ro[2]), rd[2]);
  float op[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) op[k] = __fadd_rn(ro[k], __fmul_rn(tt, rd[k]));
  const float kw = __fdiv_rn(-1.0f, __fdiv_rn((float)W, __fmul_rn(2.0f, f)));
  const float kh = __fdiv_rn(-1.0f, __fdiv_rn((float)H, __fmul_rn(2.0f, f)));
  out[0] = __fdiv_rn(__fmul_rn(kw, op[0]), op[2]);
  out[1] = __fdiv_rn(__fmul_rn(kh, op[1]), op[2]);
  out[2] = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(2.0f, near_), op[2]));
  out[3] = __fmul_rn(kw, __fsub_rn(__fdiv_rn(rd[0], rd[2]), __fdiv_rn(op[0], op[2])));
  out[4] = __fmul_rn(kh, __fsub_rn(__fdiv_rn(rd[1], rd[2]), __fdiv_rn(op[1], op[2])));
  out[5] = __fdiv_rn(__fmul_rn(-2.0f, near_), op[2]);
}

// reverse mode of k_raygen_fwd: grays [B][6] -> gmtx [F][3][4] (atomics), gfocal [1] (block-reduced atomics)
__global__ void __launch_bounds__(128)
k_raygen_bwd(const long long* __restrict__ ids, const float* __restrict__ uv, const long long* __restrict__ view_ids, int B,
             const float* __restrict__ mtx, const float* __restrict__ focal_p, int H, int W, int ndc, float near_,
             const float* __restrict__ grays, float* __restrict__ gmtx, float* __restrict__ gfocal) {
  __shared__ double red[4];
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float gf = 0.f;
  if (b < B) {
    float pi, pj;
    int view;
    raygen_pixel(ids, uv, view_ids, b, H, W, pi, pj, view);
    const float f = __ldg(focal_p);
    const RayGenCam c = load_cam(mtx, view);
    const float ci = pi - (float)(W / 2.0), cj = pj - (float)(H / 2.0);
    const float dir[3] = {ci / f, -cj / f, -1.0f};
    float rd[3], ro[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rd[k] = c.R[k][0] * dir[0] + c.R[k][1] * dir[1] + c.R[k][2] * dir[2];
      ro[k] = c.t[k];
    }
    const float* g = grays + (size_t)b * 6;
    float g_ro[3] = {0.f, 0.f, 0.f}, g_rd[3] = {0.f, 0.f, 0.f};
    if (!ndc) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        g_ro[k] = g[k];
        g_rd[k] = g[3 + k];
      }
    } else {
      const float tt = -(near_ + ro[2]) / rd[2];
      float op[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) op[k] = ro[k] + tt * rd[k];
      const float kw = -2.0f * f / (float)W, kh = -2.0f * f / (float)H;   // -1 / (W / (2 f))
      const float iz = 1.0f / op[2], irz = 1.0f / rd[2];
      const float ax = op[0] * iz, ay = op[1] * iz;        // o'_x / o'_z, o'_y / o'_z
      const float bx = rd[0] * irz, by = rd[1] * irz;      // rd_x / rd_z, rd_y / rd_z
      // outputs: o0 = kw ax, o1 = kh ay, o2 = 1 + 2 near iz, d0 = kw (bx - ax), d1 = kh (by - ay), d2 = -2 near iz
      const float g_ax = kw * (g[0] - g[3]), g_ay = kh * (g[1] - g[4]);
      const float g_bx = kw * g[3], g_by = kh * g[4];
      const float g_iz = 2.0f * near_ * (g[2] - g[5]);
      const float g_kw = g[0] * ax + g[3] * (bx - ax), g_kh = g[1] * ay + g[4] * (by - ay);
      gf += g_kw * (-2.0f / (float)W) + g_kh * (-2.0f / (float)H);
      float g_op[3];
      g_op[0] = g_ax * iz;
      g_op[1] = g_ay * iz;
      g_op[2] = -(g_ax * ax + g_ay * ay + g_iz * iz) * iz;
      g_rd[0] = g_bx * irz;
      g_rd[1] = g_by * irz;
      g_rd[2] = -(g_bx * bx + g_by * by) * irz;
      // op = ro + tt rd
      float g_tt = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        g_ro[k] = g_op[k];
        g_rd[k] += tt * g_op[k];
        g_tt += rd[k] * g_op[k];
      }
      // tt = -(near + ro_z) / rd_z
      g_ro[2] += -g_tt * irz;
      g_rd[2] += -g_tt * tt * irz;
    }
    // rd_k = sum_m R[k][m] dir[m];  ro = t
    float g_dir[3] = {0.f, 0.f, 0.f};
    float* gm = gmtx + (size_t)view * 12;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        atomicAdd(gm + k * 4 + m, g_rd[k] * dir[m]);
        g_dir[m] += g_rd[k] * c.R[k][m];
      }
      atomicAdd(gm + k * 4 + 3, g_ro[k]);
    }
    // dir = (ci / f, -cj / f, -1)
    gf += -(g_dir[0] * dir[0] + g_dir[1] * dir[1]) / f;
  }
  const double tot = block_sum_d<128>((double)gf, red);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(gfocal, (float)tot);
}

// ------------------------------------------------------------------------------------------------ f2: expected point

// moved = pts (+ sf) (clamped to [lo, hi] when clamp != 0);  pm[n] = sum_s w[n,s] moved[n,s],  acc[n] = sum_s w[n,s]
__global__ void __launch_bounds__(128)
k_weighted_point_fwd(const float* __restrict__ w, const float* __restrict__ pts, const float* __restrict__ sf, int clamp, float lo,
                     float hi, int N, int S, float* __restrict__ pm, float* __restrict__ acc) {
  __shared__ float red[4][4];
  const int n = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = threadIdx.x; s < S; s += 128) {
    const size_t i = (size_t)n * S + s;
    const float ww = __ldg(w + i);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float p = __ldg(pts + i * 3 + c);
      if (sf) p += __ldg(sf + i * 3 + c);
      if (clamp) p = fminf(fmaxf(p, lo), hi);
      a[c] = fmaf(ww, p, a[c]);
    }
    a[3] += ww;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) a[c] = warp_sum(a[c]);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) red[wid][c] = a[c];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const float v = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
    if (threadIdx.x < 3) pm[(size_t)n * 3 + threadIdx.x] = v;
    else acc[n] = v;
  }
}

// gw[n,s] = gpm[n] . moved[n,s] + gacc[n];  gmoved[n,s] = w[n,s] gpm[n] (0 where the clamp is active)
__global__ void k_weighted_point_bwd(const float* __restrict__ w, const float* __restrict__ pts, const float* __restrict__ sf, int clamp,
                                     float lo, float hi, int N, int S, const float* __restrict__ gpm, const float* __restrict__ gacc,
                                     float* __restrict__ gw, float* __restrict__ gmoved) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * S) return;
  const int n = (int)(i / S);
  const float ww = __ldg(w + i);
  float dot = gacc ? __ldg(gacc + n) : 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p = __ldg(pts + i * 3 + c);
    if (sf) p += __ldg(sf + i * 3 + c);
    bool pass = true;
    if (clamp) {
      pass = (p >= lo) && (p <= hi);   // torch.clamp: gradient passes where lo <= x <= hi
      p = fminf(fmaxf(p, lo), hi);
    }
    const float g = __ldg(gpm + (size_t)n * 3 + c);
    dot = fmaf(g, p, dot);
    if (gmoved) gmoved[i * 3 + c] = pass ? ww * g : 0.f;
  }
  if (gw) gw[i] = dot;
}

// ------------------------------------------------------------------------------------------------ f3: total variation

constexpr int kTvMax = 6;
struct TvArgs {
  const float* x[kTvMax];   // channel-last [H][W][C]
  float* g[kTvMax];         // same layout, overwritten with d(total)/dx, or null
  int C[kTvMax], H[kTvMax], W[kTvMax];
  float ch[kTvMax], cw[kTvMax];   // d(total)/d(h_tv), d(total)/d(w_tv)
  int blk0[kTvMax + 1];     // first block of each tensor
  int n;
};

// one thread per texel (all C channels); sums[t*2+0] += sum (x[h+1]-x[h])^2, sums[t*2+1] += sum (x[w+1]-x[w])^2
__global__ void __launch_bounds__(256)
k_tv_fused(TvArgs a, double* __restrict__ sums) {
  __shared__ double red[8];
  int t = 0;
  while (t + 1 < a.n && (int)blockIdx.x >= a.blk0[t + 1]) ++t;
  const int C = a.C[t], H = a.H[t], W = a.W[t];
  const float* __restrict__ x = a.x[t];
  float* __restrict__ g = a.g[t];
  const long long texel = (long long)(blockIdx.x - a.blk0[t]) * 256 + threadIdx.x;
  float hs = 0.f, ws = 0.f;
  if (texel < (long long)H * W) {
    const int h = (int)(texel / W), w = (int)(texel % W);
    const size_t base = (size_t)texel * C;
    const bool hn = h + 1 < H, wn = w + 1 < W, hp = h > 0, wp = w > 0;
    const float ch2 = 2.f * a.ch[t], cw2 = 2.f * a.cw[t];
    for (int c = 0; c < C; c += 4) {
      const float4 v = ldg4(x + base + c);
      float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hn) {
        const float4 u = ldg4(x + base + (size_t)W * C + c);
        const float4 d = make_float4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w);
        hs += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
        gr = fma4(d, -ch2, gr);
      }
      if (wn) {
        const float4 u = ldg4(x + base + C + c);
        const float4 d = make_float4(u.x - v.x, u.y - v.y, u.z - v.z, u.w - v.w);
        ws += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
        gr = fma4(d, -cw2, gr);
      }
      if (g) {
        if (hp) {
          const float4 u = ldg4(x + base - (size_t)W * C + c);
          gr = fma4(make_float4(v.x - u.x, v.y - u.y, v.z - u.z, v.w - u.w), ch2, gr);
        }
        if (wp) {
          const float4 u = ldg4(x + base - C + c);
          gr = fma4(make_float4(v.x - u.x, v.y - u.y, v.z - u.z, v.w - u.w), cw2, gr);
        }
        *reinterpret_cast<float4*>(g + base + c) = gr;
      }
    }
  }
  const double th = block_sum_d<256>((double)hs, red);
  const double tw = block_sum_d<256>((double)ws, red);
  if (threadIdx.x == 0) {
    if (th != 0.0) atomicAdd(sums + t * 2, th);
    if (tw != 0.0) atomicAdd(sums + t * 2 + 1, tw);
  }
}

// ------------------------------------------------------------------------------------------------ f3: density_L1 / blending_L1

// mean_{x,y,z} | act( sum_c p0[y][x][c] l0[z][c] + p1[z][x][c] l1[y][c] + p2[z][y][c] l2[x][c] ) |, C = [16,4,4], never
// materialising the volume.  Thread = one (x,y) column, loops over its z chunk.  Warp tile 8(x) x 4(y), block = 4 x 2 warps.
// With gradients (gp0 != null) the same pass accumulates d(sum)/d(every table): register accumulators for the tables indexed by
// the thread's own (x,y), warp transposing reductions + atomics for the others.  Gradients are of the SUM; the caller scales
// by 1 / (Gx Gy Gz).
constexpr int kL1Threads = 256;

__device__ __forceinline__ float l1_act_abs(float v, int act, float shift, float& dv) {
  if (act == DYNRF_ACT_RELU) {
    dv = v > 0.f ? 1.f : 0.f;
    return fmaxf(v, 0.f);
  }
  const float xs = v + shift;
  dv = 1.0f / (1.0f + expf(-xs));          // softplus' = sigmoid; softplus > 0 so |.| is the identity
  return xs > 20.f ? xs : log1pf(expf(xs));
}

__global__ void __launch_bounds__(kL1Threads)
k_dense_l1(dynrf_vm_t vm, int act, float shift, int Gx, int Gy, int Gz, int zchunk, double* __restrict__ sum_out,
           float* __restrict__ gp0, float* __restrict__ gp1, float* __restrict__ gp2, float* __restrict__ gl0, float* __restrict__ gl1,
           float* __restrict__ gl2) {
  extern __shared__ float gl0_s[];   // [zchunk][16] when gradients are requested
  __shared__ double red[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tx = lane & 7, ty = lane >> 3;
  const int x = blockIdx.x * 32 + (wid & 3) * 8 + tx;
  const int y = blockIdx.y * 8 + (wid >> 2) * 4 + ty;
  const int z0 = blockIdx.z * zchunk, z1 = min(Gz, z0 + zchunk);
  const bool in = x < Gx && y < Gy;
  const bool grad = gp0 != nullptr;
  if (grad)
    for (int i = threadIdx.x; i < zchunk * 16; i += kL1Threads) gl0_s[i] = 0.f;
  float P0[16], L1v[4] = {0.f, 0.f, 0.f, 0.f}, L2v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 16; ++c) P0[c] = 0.f;
  if (in) {
    const float* p = vm.plane[0] + ((size_t)y * vm.sv[0] + (size_t)x * vm.su[0]) * 16;              // plane 0: (h, w) = (y, x)
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 v = ldg4(p + c);
      P0[c] = v.x; P0[c + 1] = v.y; P0[c + 2] = v.z; P0[c + 3] = v.w;
    }
    const float4 a = ldg4(vm.line[1] + (size_t)y * 4), b = ldg4(vm.line[2] + (size_t)x * 4);
    L1v[0] = a.x; L1v[1] = a.y; L1v[2] = a.z; L1v[3] = a.w;
    L2v[0] = b.x; L2v[1] = b.y; L2v[2] = b.z; L2v[3] = b.w;
  }
  float gP0[16], gL1[4] = {0.f, 0.f, 0.f, 0.f}, gL2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 16; ++c) gP0[c] = 0.f;
  if (grad) __syncthreads();
  float total = 0.f;
  for (int z = z0; z < z1; ++z) {
    float L0[16];
    const float* l0 = vm.line[0] + (size_t)z * 16;
#pragma unroll
    for (int c = 0; c < 16; c += 4) {
      const float4 v = ldg4(l0 + c);
      L0[c] = v.x; L0[c + 1] = v.y; L0[c + 2] = v.z; L0[c + 3] = v.w;
    }
    float4 P1 = make_float4(0.f, 0.f, 0.f, 0.f), P2 = P1;
    if (in) {
      P1 = ldg4(vm.plane[1] + ((size_t)z * vm.sv[1] + (size_t)x * vm.su[1]) * 4);     // plane 1: (h, w) = (z, x)
      P2 = ldg4(vm.plane[2] + ((size_t)z * vm.sv[2] + (size_t)y * vm.su[2]) * 4);     // plane 2: (h, w) = (z, y)
    }
    float v = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) v = fmaf(P0[c], L0[c], v);
    v += P1.x * L1v[0] + P1.y * L1v[1] + P1.z * L1v[2] + P1.w * L1v[3];
    v += P2.x * L2v[0] + P2.y * L2v[1] + P2.z * L2v[2] + P2.w * L2v[3];
    float dv;
    const float av = l1_act_abs(v, act, shift, dv);
    if (in) total += av;
    if (!grad) continue;
    const float g = in ? dv : 0.f;
    // own-column tables: registers
#pragma unroll
    for (int c = 0; c < 16; ++c) gP0[c] = fmaf(g, L0[c], gP0[c]);
    gL1[0] = fmaf(g, P1.x, gL1[0]); gL1[1] = fmaf(g, P1.y, gL1[1]); gL1[2] = fmaf(g, P1.z, gL1[2]); gL1[3] = fmaf(g, P1.w, gL1[3]);
    gL2[0] = fmaf(g, P2.x, gL2[0]); gL2[1] = fmaf(g, P2.y, gL2[1]); gL2[2] = fmaf(g, P2.z, gL2[2]); gL2[3] = fmaf(g, P2.w, gL2[3]);
    // d/d line0[z][c] = sum over the whole (x,y) tile of g * P0[c]: transposing warp reduction of 16 columns (16 shuffles)
    {
      float r[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) r[c] = g * P0[c];
#pragma unroll
      for (int s = 8; s >= 1; s >>= 1) {      // lane bits 4..1 select the column
        const bool up = (lane & (s << 1)) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
          const float send = up ? r[i] : r[i + s];
          const float keep = up ? r[i + s] : r[i];
          r[i] = keep + __shfl_xor_sync(0xffffffffu, send, s << 1);
        }
      }
      r[0] += __shfl_xor_sync(0xffffffffu, r[0], 1);
      if ((lane & 1) == 0) atomicAdd(&gl0_s[(z - z0) * 16 + (lane >> 1)], r[0]);
    }
    // d/d plane1[z][x][c] = sum_y g * l1[y][c]: reduce over ty (lane bits 3,4)
    {
      float4 r = make_float4(g * L1v[0], g * L1v[1], g * L1v[2], g * L1v[3]);
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        r.x += __shfl_xor_sync(0xffffffffu, r.x, o); r.y += __shfl_xor_sync(0xffffffffu, r.y, o);
        r.z += __shfl_xor_sync(0xffffffffu, r.z, o); r.w += __shfl_xor_sync(0xffffffffu, r.w, o);
      }
      if (ty == 0 && x < Gx) red_add4g(gp1 + ((size_t)z * vm.sv[1] + (size_t)x * vm.su[1]) * 4, r);
    }
    // d/d plane2[z][y][c] = sum_x g * l2[x][c]: reduce over tx (lane bits 0..2)
    {
      float4 r = make_float4(g * L2v[0], g * L2v[1], g * L2v[2], g * L2v[3]);
#pragma unroll
      for (int o = 1; o <= 4; o <<= 1) {
        r.x += __shfl_xor_sync(0xffffffffu, r.x, o); r.y += __shfl_xor_sync(0xffffffffu, r.y, o);
        r.z += __shfl_xor_sync(0xffffffffu, r.z, o); r.w += __shfl_xor_sync(0xffffffffu, r.w, o);
      }
      if (tx == 0 && y < Gy) red_add4g(gp2 + ((size_t)z * vm.sv[2] + (size_t)y * vm.su[2]) * 4, r);
    }
  }
  const double tot = block_sum_d<kL1Threads>((double)total, red);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(sum_out, tot);
  if (!grad) return;
  if (in) {
    float* p = gp0 + ((size_t)y * vm.sv[0] + (size_t)x * vm.su[0]) * 16;
#pragma unroll
    for (int c = 0; c < 16; c += 4) red_add4g(p + c, make_float4(gP0[c], gP0[c + 1], gP0[c + 2], gP0[c + 3]));   // z chunks add up
  }
  // line1[y]: sum over x within the warp (tx), then one reduction per (warp, y); line2[x]: sum over y within the warp (ty)
  {
    float4 r = make_float4(gL1[0], gL1[1], gL1[2], gL1[3]);
#pragma unroll
    for (int o = 1; o <= 4; o <<= 1) {
      r.x += __shfl_xor_sync(0xffffffffu, r.x, o); r.y += __shfl_xor_sync(0xffffffffu, r.y, o);
      r.z += __shfl_xor_sync(0xffffffffu, r.z, o); r.w += __shfl_xor_sync(0xffffffffu, r.w, o);
    }
    if (tx == 0 && y < Gy) red_add4g(gl1 + (size_t)y * 4, r);
    float4 q = make_float4(gL2[0], gL2[1], gL2[2], gL2[3]);
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      q.x += __shfl_xor_sync(0xffffffffu, q.x, o); q.y += __shfl_xor_sync(0xffffffffu, q.y, o);
      q.z += __shfl_xor_sync(0xffffffffu, q.z, o); q.w += __shfl_xor_sync(0xffffffffu, q.w, o);
    }
    if (ty == 0 && x < Gx) red_add4g(gl2 + (size_t)x * 4, q);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (z1 - z0) * 16; i += kL1Threads) {
    const float v = gl0_s[i];
    if (v != 0.f) atomicAdd(gl0 + (size_t)z0 * 16 + i, v);
  }
}

}  // namespace dynrf

using namespace dynrf;

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int dynrf_sample_ray(const dynrf_field_cfg_t* cfg, const float* rays, int N, int S, float step_size, const float* jitter, float* xyz,
                     float* z_vals, uint8_t* valid, void* stream) {
  if (!cfg || !rays || !xyz || !z_vals || !valid || N <= 0 || S <= 0) return DYNRF_E_ARG;
  k_sample_ray<<<N, 128, 0, S_(stream)>>>(*cfg, rays, N, S, step_size, jitter, xyz, z_vals, valid);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_raygen_fwd(const int64_t* ids, const float* uv, const int64_t* view, int B, const float* mtx, int F, const float* focal, int H,
                     int W, int ndc, float near_, float* rays, void* stream) {
  if ((!ids && !(uv && view)) || !mtx || !focal || !rays || B <= 0 || F <= 0 || H <= 0 || W <= 0) return DYNRF_E_ARG;
  k_raygen_fwd<<<(B + 127) / 128, 128, 0, S_(stream)>>>(reinterpret_cast<const long long*>(ids), uv, reinterpret_cast<const long long*>(view),
                                                        B, mtx, focal, H, W, ndc, near_, rays);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_raygen_bwd(const int64_t* ids, const float* uv, const int64_t* view, int B, const float* mtx, int F, const float* focal, int H,
                     int W, int ndc, float near_, const float* grays, float* gmtx, float* gfocal, void* stream) {
  if ((!ids && !(uv && view)) || !mtx || !focal || !grays || !gmtx || !gfocal || B <= 0 || F <= 0) return DYNRF_E_ARG;
  k_raygen_bwd<<<(B + 127) / 128, 128, 0, S_(stream)>>>(reinterpret_cast<const long long*>(ids), uv, reinterpret_cast<const long long*>(view),
                                                        B, mtx, focal, H, W, ndc, near_, grays, gmtx, gfocal);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_weighted_point_fwd(const float* w, const float* pts, const float* sf, int clamp, float lo, float hi, int N, int S, float* pm,
                             float* acc, void* stream) {
  if (!w || !pts || !pm || !acc || N <= 0 || S <= 0) return DYNRF_E_ARG;
  k_weighted_point_fwd<<<N, 128, 0, S_(stream)>>>(w, pts, sf, clamp, lo, hi, N, S, pm, acc);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_weighted_point_bwd(const float* w, const float* pts, const float* sf, int clamp, float lo, float hi, int N, int S,
                             const float* gpm, const float* gacc, float* gw, float* gmoved, void* stream) {
  if (!w || !pts || !gpm || N <= 0 || S <= 0 || (!gw && !gmoved)) return DYNRF_E_ARG;
  const size_t P = (size_t)N * S;
  k_weighted_point_bwd<<<(unsigned)((P + 255) / 256), 256, 0, S_(stream)>>>(w, pts, sf, clamp, lo, hi, N, S, gpm, gacc, gw, gmoved);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_tv_fused(int n, const float* const* x, float* const* g, const int32_t* C, const int32_t* H, const int32_t* W, const float* ch,
                   const float* cw, double* sums, void* stream) {
  if (n <= 0 || n > kTvMax || !x || !C || !H || !W || !ch || !cw || !sums) return DYNRF_E_ARG;
  TvArgs a;
  a.n = n;
  int blk = 0;
  for (int t = 0; t < n; ++t) {
    if (!x[t] || C[t] <= 0 || H[t] <= 0 || W[t] <= 0) return DYNRF_E_ARG;
    if (C[t] & 3) return DYNRF_E_UNSUPPORTED;
    if (((uintptr_t)x[t] & 15) || (g && g[t] && ((uintptr_t)g[t] & 15))) return DYNRF_E_ALIGN;
    a.x[t] = x[t];
    a.g[t] = g ? g[t] : nullptr;
    a.C[t] = C[t]; a.H[t] = H[t]; a.W[t] = W[t];
    a.ch[t] = ch[t]; a.cw[t] = cw[t];
    a.blk0[t] = blk;
    blk += (int)(((long long)H[t] * W[t] + 255) / 256);
  }
  a.blk0[n] = blk;
  for (int t = n; t < kTvMax; ++t) a.blk0[t + 1] = blk;
  k_tv_fused<<<blk, 256, 0, S_(stream)>>>(a, sums);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dense_l1(const dynrf_vm_t* vm, int act, float shift, double* sum_out, const dynrf_vm_grad_t* g, void* stream) {
  if (vm && !dynrf_vm_layout_ok(vm)) return DYNRF_E_UNSUPPORTED;
  if (!vm || !sum_out) return DYNRF_E_ARG;
  if (vm->C[0] != 16 || vm->C[1] != 4 || vm->C[2] != 4) return DYNRF_E_UNSUPPORTED;
  const int Gx = vm->W[0], Gy = vm->H[0], Gz = vm->L[0];
  if (vm->W[1] != Gx || vm->H[1] != Gz || vm->W[2] != Gy || vm->H[2] != Gz || vm->L[1] != Gy || vm->L[2] != Gx) return DYNRF_E_ARG;
  for (int i = 0; i < 3; ++i)
    if (!vm->plane[i] || !vm->line[i] || ((uintptr_t)vm->plane[i] & 15) || ((uintptr_t)vm->line[i] & 15)) return DYNRF_E_ALIGN;
  const int bx = (Gx + 31) / 32, by = (Gy + 7) / 8;
  // enough z chunks for ~4 waves of 148 SMs x 2 CTAs, at most 64 planes of shared-memory line gradients per chunk
  int nz = (4 * 296 + bx * by - 1) / (bx * by);
  nz = nz < 1 ? 1 : (nz > Gz ? Gz : nz);
  int zchunk = (Gz + nz - 1) / nz;
  if (zchunk > 256) zchunk = 256;
  nz = (Gz + zchunk - 1) / zchunk;
  const bool grad = g != nullptr;
  const size_t smem = grad ? (size_t)zchunk * 16 * sizeof(float) : 0;
  dim3 grid(bx, by, nz);
  k_dense_l1<<<grid, kL1Threads, smem, S_(stream)>>>(*vm, act, shift, Gx, Gy, Gz, zchunk, sum_out, grad ? g->plane[0] : nullptr,
                                                     grad ? g->plane[1] : nullptr, grad ? g->plane[2] : nullptr, grad ? g->line[0] : nullptr,
                                                     grad ? g->line[1] : nullptr, grad ? g->line[2] : nullptr);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
