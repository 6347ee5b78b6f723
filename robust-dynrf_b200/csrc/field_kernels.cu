// Sample-tile kernels (MLP-bearing stages): weight packing, per-ray time embedding,
// coordinate warp (a15), dynamic density + blending (a16, a17), dynamic appearance (a18 + a14)
// and static appearance (a7 + a11 + a12/a13).
//
// All of them push tiles of 128 samples through the tile MLP engine (dynrf_mlp.cuh):
// gather / encode the layer input feature-major into shared memory, run the dense layers,
// apply the narrow output layer per sample.  Grids are persistent (a multiple of the SM count).
#include "dynrf_tile.cuh"

namespace dynrf {

// ------------------------------------------------------------------------------------------------
// packing

__global__ void k_pack_linear(const float* __restrict__ w, const float* __restrict__ bias, int N, int K, int Kpad, int Npad,
                              float* __restrict__ wt, float* __restrict__ bias_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Kpad * Npad) {
    const int k = i / Npad, n = i % Npad;
    wt[i] = (k < K && n < N) ? __ldg(w + (size_t)n * K + k) : 0.f;
  }
  if (i < Npad) bias_out[i] = (bias && i < N) ? __ldg(bias + i) : 0.f;
}

__global__ void k_unpack_linear_grad(const float* __restrict__ dwt, int N, int K, int Npad, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * K) {
    const int n = i / K, k = i % K;
    dw[i] = dwt[(size_t)k * Npad + n];
  }
}

// ------------------------------------------------------------------------------------------------
// a15 (time branch): tfeat[n] = [t, sin(t 2^f) f<8, cos(t 2^f) f<8 | layer2(relu(layer1(.))) | 0]

__global__ void __launch_bounds__(64) k_time_embed(const float* __restrict__ ts, int N, dynrf_layer_t l1, dynrf_layer_t l2,
                                                   float* __restrict__ tfeat) {
  __shared__ float tin[20];
  __shared__ float hid[64];
  const int n = blockIdx.x, j = threadIdx.x;
  const float t = __ldg(ts + n);
  if (j == 0) tin[0] = t;
  if (j < 8) {
    float s, c;
    sincosf(t * (float)(1 << j), &s, &c);
    tin[1 + j] = s;
    tin[9 + j] = c;
  }
  if (j >= 17 && j < 20) tin[j] = 0.f;
  __syncthreads();
  float h = __ldg(l1.bias + j);
#pragma unroll
  for (int k = 0; k < 17; ++k) h = fmaf(tin[k], __ldg(l1.wt + (size_t)k * l1.Npad + j), h);
  hid[j] = fmaxf(h, 0.f);
  __syncthreads();
  float* o = tfeat + (size_t)n * kTFeat;
  if (j < 17) o[j] = tin[j];
  if (j < 30) {
    float v = __ldg(l2.bias + j);
#pragma unroll 8
    for (int k = 0; k < 64; ++k) v = fmaf(hid[k], __ldg(l2.wt + (size_t)k * l2.Npad + j), v);
    o[17 + j] = v;
  }
  if (j == 47) o[47] = 0.f;
}

using Scratch = TileScratch<kTileM>;

// ------------------------------------------------------------------------------------------------
// a15 (space branch): xyz -> xyz_prime.   smem: A[96][128] | H1[64][128] | wbuf | scratch ; H2 aliases A

constexpr int kWarpK = 96;  // 93 padded
constexpr size_t kWarpSmem = (size_t)(kWarpK * kTileM + 64 * kTileM + 2 * kChunkK * 64) * sizeof(float) + sizeof(Scratch);

__global__ void __launch_bounds__(kMlpThreads)
k_warp(dynrf_field_cfg_t cfg, const float* __restrict__ xyz, int P, int S, const float* __restrict__ tfeat, dynrf_layer_t l3,
       dynrf_layer_t l4, dynrf_layer_t l5, float* __restrict__ xyz_prime) {
  extern __shared__ __align__(16) float smem[];
  float* A = smem;
  float* H1 = A + kWarpK * kTileM;
  float* wbuf = H1 + 64 * kTileM;
  Scratch& scr = *reinterpret_cast<Scratch*>(wbuf + 2 * kChunkK * 64);
  float* H2 = A;
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int ntiles = (P + kTileM - 1) / kTileM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kTileM) {
      const int i = tile * kTileM + tid;
      const bool in = i < P;
      scr.ray[tid] = in ? i / S : -1;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = in ? __ldg(xyz + (size_t)i * 3 + a) : 0.f;
        // the feature functions see normalize(unnormalize(normalize(xyz))) -- models/tensoRF.py:526,544-546
        const float xn = norm_coord(x, nb.mn[a], nb.inv[a]);
        const float xu = unnorm_coord(xn, nb.mn[a], nb.inv[a]);
        scr.xu[a][tid] = xu;
        scr.xq[a][tid] = norm_coord(xu, nb.mn[a], nb.inv[a]);
      }
    }
    __syncthreads();
    fill_warp_rows<kTileM, kTileM>(scr, tfeat, A, tid);
    __syncthreads();
    dense_tile<64, kTileM, kTileM, true>(A, l3, H1, wbuf, tid);
    dense_tile<64, kTileM, kTileM, true>(H1, l4, H2, wbuf, tid);
    if (tid < kTileM) {
      float d3[3];
      small_dense<3, kTileM>(H2, 64, l5, tid, d3);
      const int i = tile * kTileM + tid;
      if (i < P) {
#pragma unroll
        for (int a = 0; a < 3; ++a) xyz_prime[(size_t)i * 3 + a] = __fadd_rn(scr.xu[a][tid], d3[a]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// a16 + a17: dynamic density and blending.  A rows: 0..71 features (3 strides x 24) | 72..151 pos/time
// smem: A[152][128] | H[64][128] | wbuf[2][32][64] | scratch

constexpr int kSbK = 152;
constexpr size_t kSbSmem = (size_t)(kSbK * kTileM + 64 * kTileM + 2 * kChunkK * 64) * sizeof(float) + sizeof(Scratch);

__global__ void __launch_bounds__(kMlpThreads)
k_dyn_sigma_blend(dynrf_field_cfg_t cfg, dynrf_vm_t vden, dynrf_vm_t vbl, dynrf_layer_t d1, dynrf_layer_t d2, dynrf_layer_t b1,
                  dynrf_layer_t b2, const float* __restrict__ xyz, const float* __restrict__ xyz_prime,
                  const uint8_t* __restrict__ valid, const float* __restrict__ tfeat, int P, int S, float* __restrict__ sigma,
                  float* __restrict__ blend) {
  extern __shared__ __align__(16) float smem[];
  float* A = smem;
  float* H = A + kSbK * kTileM;
  float* wbuf = H + 64 * kTileM;
  Scratch& scr = *reinterpret_cast<Scratch*>(wbuf + 2 * kChunkK * 64);
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int ntiles = (P + kTileM - 1) / kTileM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kTileM) {
      const int i = tile * kTileM + tid;
      const bool in = i < P;
      scr.ray[tid] = in ? i / S : -1;
      scr.on[tid] = (in && valid[i] != 0) ? 1 : 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = in ? __ldg(xyz + (size_t)i * 3 + a) : 0.f;
        const float xp = in ? __ldg(xyz_prime + (size_t)i * 3 + a) : 0.f;
        scr.xn[a][tid] = norm_coord(x, nb.mn[a], nb.inv[a]);
        scr.xq[a][tid] = norm_coord(xp, nb.mn[a], nb.inv[a]);
      }
    }
    __syncthreads();
    fill_pos_time_rows<kTileM, kTileM>(scr, tfeat, A, 72, tid);
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const dynrf_vm_t& vm = which == 0 ? vden : vbl;
#pragma unroll 1
      for (int si = 0; si < 3; ++si) gather_vm_rows<kTileM, kTileM>(vm, scr.xq, 1 << si, si * 24, A, tid, scr.on);
      __syncthreads();
      dense_tile<64, kTileM, kTileM, true>(A, which == 0 ? d1 : b1, H, wbuf, tid);
      if (tid < kTileM) {
        float o[1];
        small_dense<1, kTileM>(H, 64, which == 0 ? d2 : b2, tid, o);
        const int i = tile * kTileM + tid;
        if (i < P) {
          const bool v = scr.on[tid] != 0;
          if (which == 0)
            sigma[i] = v ? density_act(o[0], cfg.act, cfg.density_shift) : 0.f;
          else
            blend[i] = v ? sigmoidf_(o[0]) : 0.f;
        }
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// a18 + a14: dynamic appearance on the active list.
// smem: Ac[72][128] | X[108][128] | H1[128][128] | wbuf[2][32][128] | scratch ; H2 aliases Ac..X
// X rows: 0..26 app feature | 27..29 xn | 30..59 sin | 60..89 cos | 90 t | 91..98 sin t | 99..106 cos t | 107 zero

constexpr int kDaK = 108;
constexpr size_t kDaSmem = (size_t)(72 * kTileM + kDaK * kTileM + 128 * kTileM + 2 * kChunkK * 128) * sizeof(float) + sizeof(Scratch);

__global__ void __launch_bounds__(kMlpThreads)
k_dyn_app(dynrf_field_cfg_t cfg, dynrf_vm_t vapp, dynrf_layer_t basis, dynrf_layer_t l1, dynrf_layer_t l2, dynrf_layer_t l3,
          const float* __restrict__ rays, const float* __restrict__ xyz, const float* __restrict__ xyz_prime,
          const float* __restrict__ tfeat, int S, const int32_t* __restrict__ active_idx, const int32_t* __restrict__ active_count,
          float* __restrict__ rgb, float* __restrict__ featT, int ldf) {
  extern __shared__ __align__(16) float smem[];
  float* Ac = smem;
  float* X = Ac + 72 * kTileM;
  float* H1 = X + kDaK * kTileM;
  float* wbuf = H1 + 128 * kTileM;
  Scratch& scr = *reinterpret_cast<Scratch*>(wbuf + 2 * kChunkK * 128);
  float* H2 = Ac;
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int count = *active_count;
  const int ntiles = (count + kTileM - 1) / kTileM;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kTileM) {
      const int j = tile * kTileM + tid;
      const bool in = j < count;
      const int i = in ? __ldg(active_idx + j) : 0;
      scr.idx[tid] = i;
      scr.ray[tid] = in ? i / S : -1;
      scr.on[tid] = in ? 1 : 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = in ? __ldg(xyz + (size_t)i * 3 + a) : 0.f;
        const float xp = in ? __ldg(xyz_prime + (size_t)i * 3 + a) : 0.f;
        scr.xn[a][tid] = norm_coord(x, nb.mn[a], nb.inv[a]);
        scr.xq[a][tid] = norm_coord(xp, nb.mn[a], nb.inv[a]);
      }
    }
    __syncthreads();
    fill_pos_time_rows<kTileM, kTileM>(scr, tfeat, X, 27, tid);
    if (tid < kTileM) X[(size_t)107 * kTileM + tid] = 0.f;
    // basis_mat over the 3 strides: 216 -> 27, accumulated 72 rows at a time
    Acc<32, kTileM> fa;
    acc_init<32, kTileM>(fa, nullptr, tid >> 4);
#pragma unroll 1
    for (int si = 0; si < 3; ++si) {
      gather_vm_rows<kTileM, kTileM>(vapp, scr.xq, 1 << si, 0, Ac, tid, scr.on);
      __syncthreads();
      acc_gemm<32, kTileM, kTileM>(fa, Ac, 72, basis.wt + (size_t)si * 72 * 32, wbuf, tid);
    }
    acc_store<32, kTileM, kTileM, false>(fa, X, 27, tid);
    if (featT) {  // basis output, feature-major, for the tensor-core backward
      for (int item = tid; item < 32 * kTileM; item += kMlpThreads) {
        const int mm = item & (kTileM - 1), k = item >> 7;
        featT[(size_t)k * ldf + tile * kTileM + mm] = (k < 27 && scr.on[mm]) ? X[(size_t)k * kTileM + mm] : 0.f;
      }
    }
    dense_tile<128, kTileM, kTileM, true>(X, l1, H1, wbuf, tid);
    dense_tile<128, kTileM, kTileM, true>(H1, l2, H2, wbuf, tid);
    if (tid < kTileM) {
      float o[3];
      small_dense<3, kTileM>(H2, 128, l3, tid, o);
      if (scr.on[tid]) {
        // mlp_view input = [h(128), viewdir(3)], viewdir = d / |d|  (models/tensorBase.py:723-725,174-180)
        float vd[3];
        ray_viewdir(rays, scr.ray[tid], vd);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = o[c];
#pragma unroll
          for (int k = 0; k < 3; ++k) v = fmaf(vd[k], __ldg(l3.wt + (size_t)(128 + k) * l3.Npad + c), v);
          rgb[(size_t)scr.idx[tid] * 3 + c] = sigmoidf_(v);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// a7 + a11 + a12/a13: static appearance on the active list.
// smem: Ac[72][128] | X[140][128] | H1[128][128] | wbuf | scratch ; H2 aliases Ac..X
// MLP_Fea           X rows: 0..26 feat | 27..29 viewdir | 30..83 sin PE(feat,2) | 84..137 cos | 138,139 zero
// MLP_Fea_TimeEmb   X rows: 0..26 feat | 27..80 sin PE(feat,2) | 81..134 cos | 135 zero ; viewdir enters the last layer

constexpr int kSaK = 140;
constexpr size_t kSaSmem = (size_t)(72 * kTileM + kSaK * kTileM + 128 * kTileM + 2 * kChunkK * 128) * sizeof(float) + sizeof(Scratch);

template <int HEAD>
__global__ void __launch_bounds__(kMlpThreads)
k_static_app(dynrf_field_cfg_t cfg, dynrf_vm_t vapp, dynrf_layer_t basis, dynrf_layer_t l1, dynrf_layer_t l2, dynrf_layer_t l3,
             const float* __restrict__ rays, const float* __restrict__ xyz, int S, const int32_t* __restrict__ active_idx,
             const int32_t* __restrict__ active_count, float* __restrict__ rgb) {
  extern __shared__ __align__(16) float smem[];
  float* Ac = smem;
  float* X = Ac + 72 * kTileM;
  float* H1 = X + kSaK * kTileM;
  float* wbuf = H1 + 128 * kTileM;
  Scratch& scr = *reinterpret_cast<Scratch*>(wbuf + 2 * kChunkK * 128);
  float* H2 = Ac;
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int count = *active_count;
  const int ntiles = (count + kTileM - 1) / kTileM;
  constexpr int kSin0 = HEAD == DYNRF_HEAD_FEA ? 30 : 27;
  constexpr int kCos0 = kSin0 + 54;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kTileM) {
      const int j = tile * kTileM + tid;
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
        scr.xq[a][tid] = norm_coord(x, nb.mn[a], nb.inv[a]);
        scr.xu[a][tid] = vd[a];  // viewdir
        if (HEAD == DYNRF_HEAD_FEA) X[(size_t)(27 + a) * kTileM + tid] = vd[a];
      }
      // zero padding rows
      if (HEAD == DYNRF_HEAD_FEA) {
        X[(size_t)138 * kTileM + tid] = 0.f;
        X[(size_t)139 * kTileM + tid] = 0.f;
      } else {
        X[(size_t)135 * kTileM + tid] = 0.f;
      }
    }
    __syncthreads();
    gather_vm_rows<kTileM, kTileM>(vapp, scr.xq, 1, 0, Ac, tid, scr.on);
    __syncthreads();
    Acc<32, kTileM> fa;
    acc_init<32, kTileM>(fa, nullptr, tid >> 4);
    acc_gemm<32, kTileM, kTileM>(fa, Ac, 72, basis.wt, wbuf, tid);
    acc_store<32, kTileM, kTileM, false>(fa, X, 27, tid);
    // PE(feat, 2): dim-major, freq-minor (models/tensorBase.py:13-19)
    for (int item = tid; item < 27 * kTileM; item += kMlpThreads) {
      const int m = item & (kTileM - 1), d = item >> 7;
      pe_rows<2, kTileM>(X[(size_t)d * kTileM + m], X, kSin0 + d * 2, kCos0 + d * 2, m);
    }
    __syncthreads();
    dense_tile<128, kTileM, kTileM, true>(X, l1, H1, wbuf, tid);
    dense_tile<128, kTileM, kTileM, true>(H1, l2, H2, wbuf, tid);
    if (tid < kTileM) {
      float o[3];
      small_dense<3, kTileM>(H2, 128, l3, tid, o);
      if (scr.on[tid]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v = o[c];
          if (HEAD != DYNRF_HEAD_FEA) {
#pragma unroll
            for (int k = 0; k < 3; ++k) v = fmaf(scr.xu[k][tid], __ldg(l3.wt + (size_t)(128 + k) * l3.Npad + c), v);
          }
          rgb[(size_t)scr.idx[tid] * 3 + c] = sigmoidf_(v);
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------

}  // namespace dynrf

using namespace dynrf;

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int check_layer(const dynrf_layer_t* L, int K, int N, int npad_expect) {
  if (!L || !L->wt || !L->bias) return DYNRF_E_ARG;
  if (L->K != K || L->N != N) return DYNRF_E_UNSUPPORTED;
  if (L->Kpad != ((K + 3) & ~3) || L->Npad != npad_expect) return DYNRF_E_UNSUPPORTED;
  if ((uintptr_t)L->wt & 15) return DYNRF_E_ALIGN;
  return 0;
}

static int check_vm_counts(const dynrf_vm_t* vm, int c0, int c1, int c2) {
  if (!vm) return DYNRF_E_ARG;
  for (int i = 0; i < 3; ++i) {
    if (!vm->plane[i] || !vm->line[i]) return DYNRF_E_ARG;
    if (((uintptr_t)vm->plane[i] & 15) || ((uintptr_t)vm->line[i] & 15)) return DYNRF_E_ALIGN;
    if (vm->H[i] < 1 || vm->W[i] < 1 || vm->L[i] < 1) return DYNRF_E_ARG;
  }
  if (vm->C[0] != c0 || vm->C[1] != c1 || vm->C[2] != c2) return DYNRF_E_UNSUPPORTED;
  return 0;
}

#define DYNRF_TRY(x)            \
  do {                          \
    if (int _e = (x)) return _e; \
  } while (0)

template <typename Kern>
static int set_smem(Kern k, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? 0 : (int)e;
}

extern "C" {

int dynrf_abi_version(void) { return DYNRF_ABI_VERSION; }

const char* dynrf_strerror(int code) {
  switch (code) {
    case 0: return "ok";
    case DYNRF_E_ARG: return "dynrf: bad argument (null pointer or non-positive size)";
    case DYNRF_E_UNSUPPORTED: return "dynrf: configuration not supported by the compiled kernels";
    case DYNRF_E_ALIGN: return "dynrf: pointer not 16-byte aligned";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "dynrf: unknown error";
  }
}

int dynrf_pad_k(int K) { return (K + 3) & ~3; }
int dynrf_pad_n(int N) { return (N + 15) & ~15; }

int dynrf_pack_linear(const float* w, const float* bias, int N, int K, float* wt, float* bias_out, void* stream) {
  if (!w || !wt || !bias_out || N <= 0 || K <= 0) return DYNRF_E_ARG;
  const int Kpad = dynrf_pad_k(K), Npad = dynrf_pad_n(N);
  const int total = Kpad * Npad;
  k_pack_linear<<<(total + 255) / 256, 256, 0, S_(stream)>>>(w, bias, N, K, Kpad, Npad, wt, bias_out);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_unpack_linear_grad(const float* dwt, int N, int K, float* dw, void* stream) {
  if (!dwt || !dw || N <= 0 || K <= 0) return DYNRF_E_ARG;
  k_unpack_linear_grad<<<(N * K + 255) / 256, 256, 0, S_(stream)>>>(dwt, N, K, dynrf_pad_n(N), dw);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_time_embed_fwd(const float* ts, int N, const dynrf_layer_t* l1, const dynrf_layer_t* l2, float* tfeat, void* stream) {
  if (!ts || !tfeat || N <= 0) return DYNRF_E_ARG;
  DYNRF_TRY(check_layer(l1, 17, 64, 64));
  DYNRF_TRY(check_layer(l2, 64, 30, 32));
  k_time_embed<<<N, 64, 0, S_(stream)>>>(ts, N, *l1, *l2, tfeat);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_warp_fwd(const dynrf_field_cfg_t* cfg, const float* xyz, int N, int S, const float* tfeat, const dynrf_layer_t* l3,
                   const dynrf_layer_t* l4, const dynrf_layer_t* l5, float* xyz_prime, void* stream) {
  if (!cfg || !xyz || !tfeat || !xyz_prime || N <= 0 || S <= 0) return DYNRF_E_ARG;
  DYNRF_TRY(check_layer(l3, 93, 64, 64));
  DYNRF_TRY(check_layer(l4, 64, 64, 64));
  DYNRF_TRY(check_layer(l5, 64, 3, 16));
  DYNRF_TRY(set_smem(k_warp, kWarpSmem));
  const int P = N * S;
  const int ntiles = (P + kTileM - 1) / kTileM;
  const int grid = min(ntiles, 2 * sm_count());
  k_warp<<<grid, kMlpThreads, kWarpSmem, S_(stream)>>>(*cfg, xyz, P, S, tfeat, *l3, *l4, *l5, xyz_prime);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_sigma_blend_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                              const dynrf_layer_t* d1, const dynrf_layer_t* d2, const dynrf_layer_t* b1, const dynrf_layer_t* b2,
                              const float* xyz, const float* xyz_prime, const uint8_t* valid, const float* tfeat, int N, int S,
                              float* sigma, float* blend, void* stream) {
  if (density && blending && (!dynrf_vm_layout_ok(density) || !dynrf_vm_same_layout(density, blending))) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !xyz || !xyz_prime || !valid || !tfeat || !sigma || !blend || N <= 0 || S <= 0) return DYNRF_E_ARG;
  DYNRF_TRY(check_vm_counts(density, 16, 4, 4));
  DYNRF_TRY(check_vm_counts(blending, 16, 4, 4));
  DYNRF_TRY(check_layer(d1, 152, 64, 64));
  DYNRF_TRY(check_layer(d2, 64, 1, 16));
  DYNRF_TRY(check_layer(b1, 152, 64, 64));
  DYNRF_TRY(check_layer(b2, 64, 1, 16));
  DYNRF_TRY(set_smem(k_dyn_sigma_blend, kSbSmem));
  const int P = N * S;
  const int ntiles = (P + kTileM - 1) / kTileM;
  const int grid = min(ntiles, sm_count());
  k_dyn_sigma_blend<<<grid, kMlpThreads, kSbSmem, S_(stream)>>>(*cfg, *density, *blending, *d1, *d2, *b1, *b2, xyz, xyz_prime, valid,
                                                                tfeat, P, S, sigma, blend);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_app_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_layer_t* basis, const dynrf_layer_t* l1,
                      const dynrf_layer_t* l2, const dynrf_layer_t* l3, const float* rays, const float* xyz,
                      const float* xyz_prime, const float* tfeat, int N, int S, const int32_t* active_idx,
                      const int32_t* active_count, float* rgb, float* featT, int ldf, void* stream) {
  if (app && !dynrf_vm_layout_ok(app)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !rays || !xyz || !xyz_prime || !tfeat || !active_idx || !active_count || !rgb || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if (featT && ldf < ((N * S + kTileM - 1) / kTileM) * kTileM) return DYNRF_E_ARG;
  DYNRF_TRY(check_vm_counts(app, 48, 12, 12));
  DYNRF_TRY(check_layer(basis, 216, 27, 32));
  DYNRF_TRY(check_layer(l1, 107, 128, 128));
  DYNRF_TRY(check_layer(l2, 128, 128, 128));
  DYNRF_TRY(check_layer(l3, 131, 3, 16));
  DYNRF_TRY(set_smem(k_dyn_app, kDaSmem));
  const int ntiles = (N * S + kTileM - 1) / kTileM;
  const int grid = min(ntiles, sm_count());
  k_dyn_app<<<grid, kMlpThreads, kDaSmem, S_(stream)>>>(*cfg, *app, *basis, *l1, *l2, *l3, rays, xyz, xyz_prime, tfeat, S, active_idx,
                                                        active_count, rgb, featT, ldf);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_static_app_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_layer_t* basis, const dynrf_layer_t* l1,
                         const dynrf_layer_t* l2, const dynrf_layer_t* l3, int head_mode, const float* rays, const float* xyz, int N,
                         int S, const int32_t* active_idx, const int32_t* active_count, float* rgb, void* stream) {
  if (app && !dynrf_vm_layout_ok(app)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !rays || !xyz || !active_idx || !active_count || !rgb || N <= 0 || S <= 0) return DYNRF_E_ARG;
  DYNRF_TRY(check_vm_counts(app, 48, 12, 12));
  DYNRF_TRY(check_layer(basis, 72, 27, 32));
  DYNRF_TRY(check_layer(l2, 128, 128, 128));
  const int ntiles = (N * S + kTileM - 1) / kTileM;
  const int grid = min(ntiles, sm_count());
  if (head_mode == DYNRF_HEAD_FEA) {
    DYNRF_TRY(check_layer(l1, 138, 128, 128));
    DYNRF_TRY(check_layer(l3, 128, 3, 16));
    DYNRF_TRY(set_smem(k_static_app<DYNRF_HEAD_FEA>, kSaSmem));
    k_static_app<DYNRF_HEAD_FEA><<<grid, kMlpThreads, kSaSmem, S_(stream)>>>(*cfg, *app, *basis, *l1, *l2, *l3, rays, xyz, S,
                                                                             active_idx, active_count, rgb);
  } else if (head_mode == DYNRF_HEAD_FEA_TIMEEMB) {
    DYNRF_TRY(check_layer(l1, 135, 128, 128));
    DYNRF_TRY(check_layer(l3, 131, 3, 16));
    DYNRF_TRY(set_smem(k_static_app<DYNRF_HEAD_FEA_TIMEEMB>, kSaSmem));
    k_static_app<DYNRF_HEAD_FEA_TIMEEMB><<<grid, kMlpThreads, kSaSmem, S_(stream)>>>(*cfg, *app, *basis, *l1, *l2, *l3, rays, xyz, S,
                                                                                     active_idx, active_count, rgb);
  } else {
    return DYNRF_E_UNSUPPORTED;
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
