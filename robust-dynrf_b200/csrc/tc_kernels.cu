// Tensor-core (tcgen05) versions of the MLP-bearing forward stages: coordinate warp (a15) and dynamic
// density + blending (a16, a17).  Same inputs/outputs as k_warp / k_dyn_sigma_blend in field_kernels.cu.
//
// Per 128-sample tile: the layer input is encoded feature-major in shared memory exactly as in the SIMT
// engine, converted 32 input rows at a time into the K-major UMMA operand layout as a (hi, lo) split-TF32
// pair, multiplied against pre-split, pre-laid-out weight chunks fetched by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx), and accumulated in TMEM by 3 x tcgen05.mma kind::tf32
// (hi*hi + hi*lo + lo*hi).  The epilogue reads the accumulator with tcgen05.ld, applies bias + ReLU and
// either writes the next layer's operand (again split, K-major) or finishes the narrow output layer as a
// per-row dot product in registers.
#include "dynrf_tile.cuh"
#include "dynrf_umma.cuh"

namespace dynrf {

constexpr int kKC = 32;  // input rows per MMA chunk

// packed UMMA weights of one layer: chunk c at wq + c * 2*N*32 floats: [hi N x 32 K-major][lo N x 32 K-major]
struct UmmaLayer {
  const float* wq;
  const float* bias;  // [N]
  int nchunks;
};

__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(umma::smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}

// S rows [k0, k0+32) (feature-major fp32, row stride kTileM; rows >= Kvalid read as 0) -> a_hi / a_lo chunk (K-major, 128 rows)
__device__ __forceinline__ void convert_chunk(const float* __restrict__ S, int k0, int Kvalid, float* __restrict__ a_hi,
                                              float* __restrict__ a_lo, int tid) {
  for (int item = tid; item < kTileM * (kKC / 4); item += kMlpThreads) {
    const int m = item & (kTileM - 1), kq = item >> 7;
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + kq * 4 + j;
      const float v = k < Kvalid ? S[(size_t)k * kTileM + m] : 0.f;
      umma::split_tf32(v, hi[j], lo[j]);
    }
    const uint32_t off = umma::kmajor_off(m, kq * 4, kTileM) >> 2;
    *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct Pipe {
  uint64_t* bar_mma;  // MMAs of the last committed chunk finished
  uint64_t* bar_w;    // weight chunk landed
  uint32_t par_mma, par_w;
  uint32_t tmem;
};

// One dense layer D[128][N] (TMEM columns [col0, col0+N)) = A[128][K] x W^T, A taken from the fp32 staging buffer S.
// All 256 threads call this; ends with the accumulator complete and visible (after tcgen05.fence::after_thread_sync).
template <int N>
__device__ __forceinline__ void umma_layer_from_staging(Pipe& p, const float* S, int Kvalid, const UmmaLayer& L, float* a_hi,
                                                        float* a_lo, float* w_buf, uint32_t col0, int tid) {
  const uint32_t idesc = umma::make_idesc_tf32(kTileM, N);
  constexpr uint32_t kWBytes = 2u * N * kKC * 4u;
  for (int c = 0; c < L.nchunks; ++c) {
    if (tid == 0) {
      mbar_expect_tx(p.bar_w, kWBytes);
      tma_load_1d(w_buf, L.wq + (size_t)c * 2 * N * kKC, kWBytes, p.bar_w);
    }
    convert_chunk(S, c * kKC, Kvalid, a_hi, a_lo, tid);
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::mbar_wait(p.bar_w, p.par_w);
      umma::fence_after_sync();
      const uint32_t sa_hi = umma::smem_u32(a_hi), sa_lo = umma::smem_u32(a_lo);
      const uint32_t sb_hi = umma::smem_u32(w_buf), sb_lo = sb_hi + N * kKC * 4;
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {
        const uint32_t sa = (ps == 2) ? sa_lo : sa_hi;
        const uint32_t sb = (ps == 1) ? sb_lo : sb_hi;
#pragma unroll
        for (int j = 0; j < kKC / 8; ++j) {
          const uint64_t da = umma::make_desc(sa + 2 * j * (kTileM * 16), kTileM * 16, 128);
          const uint64_t db = umma::make_desc(sb + 2 * j * (N * 16), N * 16, 128);
          umma::mma_tf32(p.tmem + col0, da, db, idesc, (c > 0 || ps > 0 || j > 0) ? 1u : 0u);
        }
      }
      umma::commit(p.bar_mma);
    }
    p.par_w ^= 1;
    // the operand buffers are reused by the next chunk: wait for this chunk's MMAs
    umma::mbar_wait(p.bar_mma, p.par_mma);
    p.par_mma ^= 1;
    umma::fence_after_sync();
  }
}

// Same, A operand already resident as split K-major buffers h_hi / h_lo with K = 32 * L.nchunks columns.
template <int N>
__device__ __forceinline__ void umma_layer_from_operand(Pipe& p, const float* h_hi, const float* h_lo, const UmmaLayer& L,
                                                        float* w_buf, uint32_t col0, int tid) {
  const uint32_t idesc = umma::make_idesc_tf32(kTileM, N);
  constexpr uint32_t kWBytes = 2u * N * kKC * 4u;
  for (int c = 0; c < L.nchunks; ++c) {
    if (tid == 0) {
      mbar_expect_tx(p.bar_w, kWBytes);
      tma_load_1d(w_buf, L.wq + (size_t)c * 2 * N * kKC, kWBytes, p.bar_w);
      umma::mbar_wait(p.bar_w, p.par_w);
      umma::fence_after_sync();
      const uint32_t sa_hi = umma::smem_u32(h_hi) + c * (kKC / 4) * (kTileM * 16);
      const uint32_t sa_lo = umma::smem_u32(h_lo) + c * (kKC / 4) * (kTileM * 16);
      const uint32_t sb_hi = umma::smem_u32(w_buf), sb_lo = sb_hi + N * kKC * 4;
#pragma unroll
      for (int ps = 0; ps < 3; ++ps) {
        const uint32_t sa = (ps == 2) ? sa_lo : sa_hi;
        const uint32_t sb = (ps == 1) ? sb_lo : sb_hi;
#pragma unroll
        for (int j = 0; j < kKC / 8; ++j) {
          const uint64_t da = umma::make_desc(sa + 2 * j * (kTileM * 16), kTileM * 16, 128);
          const uint64_t db = umma::make_desc(sb + 2 * j * (N * 16), N * 16, 128);
          umma::mma_tf32(p.tmem + col0, da, db, idesc, (c > 0 || ps > 0 || j > 0) ? 1u : 0u);
        }
      }
      umma::commit(p.bar_mma);
    }
    p.par_w ^= 1;
    umma::mbar_wait(p.bar_mma, p.par_mma);
    p.par_mma ^= 1;
    umma::fence_after_sync();
  }
}

// ------------------------------------------------------------------------------------------------
// a16 + a17 on tensor cores.  smem: S[160][128] | a_hi[32x128] | a_lo | w_buf[2*64*32] | scratch | barriers

using Scratch = TileScratch<kTileM>;
constexpr int kSbRows = 160;
constexpr size_t kSbTcSmem = (size_t)(kSbRows * kTileM + 2 * kKC * kTileM + 2 * 64 * kKC) * sizeof(float) + sizeof(Scratch) + 64;

__global__ void __launch_bounds__(kMlpThreads)
k_dyn_sigma_blend_tc(dynrf_field_cfg_t cfg, dynrf_vm_t vden, dynrf_vm_t vbl, UmmaLayer d1, dynrf_layer_t d2, UmmaLayer b1,
                     dynrf_layer_t b2, const float* __restrict__ xyz, const float* __restrict__ xyz_prime,
                     const uint8_t* __restrict__ valid, const float* __restrict__ tfeat, int P, int S, float* __restrict__ sigma,
                     float* __restrict__ blend) {
  extern __shared__ __align__(128) float smem[];
  float* Sg = smem;
  float* a_hi = Sg + kSbRows * kTileM;
  float* a_lo = a_hi + kKC * kTileM;
  float* w_buf = a_lo + kKC * kTileM;
  Scratch& scr = *reinterpret_cast<Scratch*>(w_buf + 2 * 64 * kKC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(&scr) + sizeof(Scratch));
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, 64);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Pipe pipe{&bars[0], &bars[1], 0u, 0u, *tslot};
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
    fill_pos_time_rows<kTileM, kTileM>(scr, tfeat, Sg, 72, tid);
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      const dynrf_vm_t& vm = which == 0 ? vden : vbl;
#pragma unroll 1
      for (int si = 0; si < 3; ++si) gather_vm_rows<kTileM, kTileM>(vm, scr.xq, 1 << si, si * 24, Sg, tid, scr.on);
      __syncthreads();
      const UmmaLayer& L1 = which == 0 ? d1 : b1;
      const dynrf_layer_t& L2 = which == 0 ? d2 : b2;
      umma_layer_from_staging<64>(pipe, Sg, 152, L1, a_hi, a_lo, w_buf, 0, tid);
      if (warp < 4) {
        const int m = warp * 32 + lane;
        float o = __ldg(L2.bias);
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          float r[32];
          umma::tmem_ld32(pipe.tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = cb * 32 + j;
            const float h = fmaxf(r[j] + __ldg(L1.bias + n), 0.f);
            o = fmaf(h, __ldg(L2.wt + (size_t)n * L2.Npad), o);
          }
        }
        const int i = tile * kTileM + m;
        if (i < P) {
          const bool v = scr.on[m] != 0;
          if (which == 0) sigma[i] = v ? density_act(o, cfg.act, cfg.density_shift) : 0.f;
          else blend[i] = v ? sigmoidf_(o) : 0.f;
        }
      }
      umma::fence_before_sync();
      __syncthreads();  // TMEM accumulator and the feature rows of S are free again
      umma::fence_after_sync();
    }
  }
  if (warp == 0) umma::tmem_dealloc(pipe.tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// a15 on tensor cores: 93 -> 64 -> 64 -> 3.   smem: S[96][128] | a_hi | a_lo | h_hi[64x128] | h_lo | w_buf | scratch | bars

constexpr size_t kWarpTcSmem =
    (size_t)(96 * kTileM + 2 * kKC * kTileM + 2 * 64 * kTileM + 2 * 64 * kKC) * sizeof(float) + sizeof(Scratch) + 64;

__global__ void __launch_bounds__(kMlpThreads)
k_warp_tc(dynrf_field_cfg_t cfg, const float* __restrict__ xyz, int P, int S, const float* __restrict__ tfeat, UmmaLayer l3,
          UmmaLayer l4, dynrf_layer_t l5, float* __restrict__ xyz_prime) {
  extern __shared__ __align__(128) float smem[];
  float* Sg = smem;
  float* a_hi = Sg + 96 * kTileM;
  float* a_lo = a_hi + kKC * kTileM;
  float* h_hi = a_lo + kKC * kTileM;
  float* h_lo = h_hi + 64 * kTileM;
  float* w_buf = h_lo + 64 * kTileM;
  Scratch& scr = *reinterpret_cast<Scratch*>(w_buf + 2 * 64 * kKC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(&scr) + sizeof(Scratch));
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    umma::mbar_init(&bars[0], 1);
    umma::mbar_init(&bars[1], 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, 128);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Pipe pipe{&bars[0], &bars[1], 0u, 0u, *tslot};
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
        const float xn = norm_coord(x, nb.mn[a], nb.inv[a]);
        const float xu = unnorm_coord(xn, nb.mn[a], nb.inv[a]);
        scr.xu[a][tid] = xu;
        scr.xq[a][tid] = norm_coord(xu, nb.mn[a], nb.inv[a]);
      }
    }
    __syncthreads();
    fill_warp_rows<kTileM, kTileM>(scr, tfeat, Sg, tid);
    __syncthreads();
    umma_layer_from_staging<64>(pipe, Sg, 96, l3, a_hi, a_lo, w_buf, 0, tid);
    // epilogue 1: h1 = relu(acc + b3) -> split K-major operand of layer 4
    if (warp < 4) {
      const int m = warp * 32 + lane;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        float r[32];
        umma::tmem_ld32(pipe.tmem + ((uint32_t)(warp * 32) << 16) + cb * 32, r);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = cb * 32 + q * 4 + j;
            umma::split_tf32(fmaxf(r[q * 4 + j] + __ldg(l3.bias + n), 0.f), hi[j], lo[j]);
          }
          const uint32_t off = umma::kmajor_off(m, cb * 32 + q * 4, kTileM) >> 2;
          *reinterpret_cast<float4*>(h_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(h_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    umma_layer_from_operand<64>(pipe, h_hi, h_lo, l4, w_buf, 64, tid);
    // epilogue 2: h2 = relu(acc + b4); delta = W5 h2 + b5; xyz_prime = xu + delta
    if (warp < 4) {
      const int m = warp * 32 + lane;
      float d3[3] = {__ldg(l5.bias + 0), __ldg(l5.bias + 1), __ldg(l5.bias + 2)};
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        float r[32];
        umma::tmem_ld32(pipe.tmem + ((uint32_t)(warp * 32) << 16) + 64 + cb * 32, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = cb * 32 + j;
          const float h = fmaxf(r[j] + __ldg(l4.bias + n), 0.f);
#pragma unroll
          for (int o = 0; o < 3; ++o) d3[o] = fmaf(h, __ldg(l5.wt + (size_t)n * l5.Npad + o), d3[o]);
        }
      }
      const int i = tile * kTileM + m;
      if (i < P) {
#pragma unroll
        for (int a = 0; a < 3; ++a) xyz_prime[(size_t)i * 3 + a] = __fadd_rn(scr.xu[a][m], d3[a]);
      }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (warp == 0) umma::tmem_dealloc(pipe.tmem, 128);
}

// ------------------------------------------------------------------------------------------------
// weight packing for the UMMA path: nn.Linear weight [N][K] -> chunks of 32 input columns,
// each chunk = [hi: Npad x 32 K-major][lo: Npad x 32 K-major]; bias -> [Npad]

__global__ void k_pack_linear_umma(const float* __restrict__ w, const float* __restrict__ bias, int N, int K, int Npad, int nchunks,
                                   float* __restrict__ wq, float* __restrict__ bias_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_chunk = Npad * kKC;
  if (i < nchunks * per_chunk) {
    const int c = i / per_chunk, r = i % per_chunk;
    const int n = r / kKC, kk = r % kKC;
    const int k = c * kKC + kk;
    const float v = (n < N && k < K) ? __ldg(w + (size_t)n * K + k) : 0.f;
    float hi, lo;
    umma::split_tf32(v, hi, lo);
    const uint32_t off = umma::kmajor_off(n, kk, Npad) >> 2;
    wq[(size_t)c * 2 * per_chunk + off] = hi;
    wq[(size_t)c * 2 * per_chunk + per_chunk + off] = lo;
  }
  if (i < Npad) bias_out[i] = (bias && i < N) ? __ldg(bias + i) : 0.f;
}

}  // namespace dynrf

using namespace dynrf;

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define DYNRF_TRY(x)             \
  do {                           \
    if (int _e = (x)) return _e; \
  } while (0)

static inline UmmaLayer UL(const dynrf_umma_layer_t* l) { return UmmaLayer{l->wq, l->bias, l->nchunks}; }

static int check_ul(const dynrf_umma_layer_t* l, int K, int N) {
  if (!l || !l->wq || !l->bias) return DYNRF_E_ARG;
  if (l->K != K || l->N != N || l->nchunks != (K + kKC - 1) / kKC) return DYNRF_E_UNSUPPORTED;
  if ((uintptr_t)l->wq & 15) return DYNRF_E_ALIGN;
  return 0;
}

extern "C" {

int dynrf_pack_linear_umma(const float* w, const float* bias, int N, int K, float* wq, float* bias_out, void* stream) {
  if (!w || !wq || !bias_out || N <= 0 || K <= 0) return DYNRF_E_ARG;
  const int Npad = (N + 7) & ~7, nchunks = (K + kKC - 1) / kKC;
  const int total = nchunks * Npad * kKC;
  k_pack_linear_umma<<<(total + 255) / 256, 256, 0, S_(stream)>>>(w, bias, N, K, Npad, nchunks, wq, bias_out);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_warp_fwd_tc(const dynrf_field_cfg_t* cfg, const float* xyz, int N, int S, const float* tfeat,
                      const dynrf_umma_layer_t* l3, const dynrf_umma_layer_t* l4, const dynrf_layer_t* l5, float* xyz_prime,
                      void* stream) {
  if (!cfg || !xyz || !tfeat || !xyz_prime || !l5 || N <= 0 || S <= 0) return DYNRF_E_ARG;
  DYNRF_TRY(check_ul(l3, 93, 64));
  DYNRF_TRY(check_ul(l4, 64, 64));
  if (l5->K != 64 || l5->N != 3) return DYNRF_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(k_warp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWarpTcSmem);
  if (e != cudaSuccess) return (int)e;
  const int P = N * S;
  const int grid = min((P + kTileM - 1) / kTileM, sm_count());
  k_warp_tc<<<grid, kMlpThreads, kWarpTcSmem, S_(stream)>>>(*cfg, xyz, P, S, tfeat, UL(l3), UL(l4), *l5, xyz_prime);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_sigma_blend_fwd_tc(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                                 const dynrf_umma_layer_t* d1, const dynrf_layer_t* d2, const dynrf_umma_layer_t* b1,
                                 const dynrf_layer_t* b2, const float* xyz, const float* xyz_prime, const uint8_t* valid,
                                 const float* tfeat, int N, int S, float* sigma, float* blend, void* stream) {
  if (!cfg || !density || !blending || !d2 || !b2 || !xyz || !xyz_prime || !valid || !tfeat || !sigma || !blend || N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  DYNRF_TRY(check_ul(d1, 152, 64));
  DYNRF_TRY(check_ul(b1, 152, 64));
  if (d2->K != 64 || d2->N != 1 || b2->K != 64 || b2->N != 1) return DYNRF_E_UNSUPPORTED;
  for (int i = 0; i < 3; ++i) {
    const int c = i == 0 ? 16 : 4;
    if (density->C[i] != c || blending->C[i] != c) return DYNRF_E_UNSUPPORTED;
  }
  cudaError_t e = cudaFuncSetAttribute(k_dyn_sigma_blend_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSbTcSmem);
  if (e != cudaSuccess) return (int)e;
  const int P = N * S;
  const int grid = min((P + kTileM - 1) / kTileM, sm_count());
  k_dyn_sigma_blend_tc<<<grid, kMlpThreads, kSbTcSmem, S_(stream)>>>(*cfg, *density, *blending, UL(d1), *d2, UL(b1), *b2, xyz, xyz_prime,
                                                                     valid, tfeat, P, S, sigma, blend);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
