// Weight-gradient GEMM on tensor cores:  dW[k][n] += sum_{p < P} XT[k][p] * DYT[n][p]
//
// XT [K][ldx] and DYT [N][ldy] are feature-major fp32 activations / output-gradients that the fused backward
// tile kernels stream to global memory (sample index contiguous).  The reduction runs over samples, so both
// operands are K-major UMMA operands with "K" = samples: rows = features.  Each CTA walks sample chunks of 64,
// stages the chunk as split-TF32 (hi, lo) operands in shared memory (one stage, register prefetch of the next), issues 3-pass
// tcgen05.mma (hi*hi + hi*lo + lo*hi) into a TMEM accumulator that lives for the whole kernel, and finally
// adds its partial dW into the nn.Linear-layout gradient dw[n * dw_ld + k] with atomics.
#include "dynrf_common.cuh"
#include "dynrf_umma.cuh"

namespace dynrf {

constexpr int kWgS = 32;          // samples per chunk (4 MMA k-steps)
constexpr int kWgQ = kWgS / 4;    // sample quads per chunk
constexpr int kWgThreads = 256;

// per-thread prefetch registers for one chunk: rows [0, ROWS) x 64 samples, ITEMS = ROWS*16/256 float4 each
template <int ROWS>
struct Pref {
  static constexpr int ITEMS = (ROWS * kWgQ + kWgThreads - 1) / kWgThreads;
  float4 v[ITEMS];
};

// item -> (row, sample quad): lanes r8 fastest (8 rows x 16 B contiguous in the operand), then the quads, then row octets.
// A warp touches 8 rows x 64 contiguous bytes of global memory (full sectors) and 4 x 128 contiguous bytes of shared memory.
__device__ __forceinline__ void item_rc(int item, int& row, int& sq) {
  sq = (item >> 3) & (kWgQ - 1);
  row = (item / (8 * kWgQ)) * 8 + (item & 7);
}

template <int ROWS>
__device__ __forceinline__ void prefetch_rows(Pref<ROWS>& pf, const float* __restrict__ src, int ld, int nrows, int p0, int P, int tid) {
#pragma unroll
  for (int i = 0; i < Pref<ROWS>::ITEMS; ++i) {
    const int item = tid + i * kWgThreads;
    int row, sq;
    item_rc(item, row, sq);
    const int p = p0 + sq * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (item < ROWS * kWgQ && row < nrows && p < P) {
      const float* s = src + (size_t)row * ld + p;
#ifdef DYNRF_STREAM_HINTS
      if (p + 3 < P) v = __ldcs(reinterpret_cast<const float4*>(s));   // read-once stream: evict-first
#else
      if (p + 3 < P) v = __ldg(reinterpret_cast<const float4*>(s));
#endif
      else {
        v.x = __ldg(s);
        if (p + 1 < P) v.y = __ldg(s + 1);
        if (p + 2 < P) v.z = __ldg(s + 2);
      }
    }
    pf.v[i] = v;
  }
}

// registers -> split K-major operand blocks of RB rows each (block b holds rows [b*RB, (b+1)*RB))
template <int ROWS, int RB>
__device__ __forceinline__ void store_rows(const Pref<ROWS>& pf, float* __restrict__ hi, float* __restrict__ lo, int tid) {
#pragma unroll
  for (int i = 0; i < Pref<ROWS>::ITEMS; ++i) {
    const int item = tid + i * kWgThreads;
    if (item < ROWS * kWgQ) {
      int row, sq;
      item_rc(item, row, sq);
      const float4 v = pf.v[i];
      float4 h, l;
      umma::split_tf32(v.x, h.x, l.x);
      umma::split_tf32(v.y, h.y, l.y);
      umma::split_tf32(v.z, h.z, l.z);
      umma::split_tf32(v.w, h.w, l.w);
      const int b = row / RB, r = row % RB;
      const uint32_t off = (uint32_t)b * (RB * kWgS) + (umma::kmajor_off(r, sq * 4, RB) >> 2);
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  }
}

// NB = number of 128-row blocks of XT (1 or 2), N = padded output columns (16..128, multiple of 16).
// One shared-memory stage (<= 96 KB) + register prefetch, two CTAs per SM: the global loads of chunk c+1 are in flight while
// the tensor core consumes chunk c, and one CTA converts / issues while the other CTA's loads land.
template <int NB, int N>
__global__ void __launch_bounds__(kWgThreads, 2)
k_wgrad_tc(const float* __restrict__ XT, int ldx, int K, const float* __restrict__ DYT, int ldy, int Nvalid, int Pmax,
           const int32_t* __restrict__ count_ptr, float* __restrict__ dw, int dw_ld) {
  constexpr int kAFloats = 128 * kWgS;  // one 128-row block, hi or lo
  constexpr int kBFloats = N * kWgS;
  constexpr int kStageFloats = 2 * NB * kAFloats + 2 * kBFloats;
  extern __shared__ __align__(128) float smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kStageFloats);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kCols = NB * N < 32 ? 32 : (NB * N <= 64 ? 64 : (NB * N <= 128 ? 128 : 256));
  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, kCols);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tslot;
  const int P = count_ptr ? min(*count_ptr, Pmax) : Pmax;
  const int nchunks = (P + kWgS - 1) / kWgS;
  const uint32_t idesc = umma::make_idesc_tf32(128, N);
  float* a_hi = smem;
  float* a_lo = smem + NB * kAFloats;
  float* b_hi = smem + 2 * NB * kAFloats;
  float* b_lo = b_hi + kBFloats;
  uint32_t it = 0;
  Pref<NB * 128> pa;
  Pref<N> pb;
  if ((int)blockIdx.x < nchunks) {
    prefetch_rows<NB * 128>(pa, XT, ldx, K, blockIdx.x * kWgS, P, tid);
    prefetch_rows<N>(pb, DYT, ldy, Nvalid, blockIdx.x * kWgS, P, tid);
  }
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
    if (it > 0) umma::mbar_wait(bar, (it - 1) & 1);  // the MMAs that read the stage are done
    store_rows<NB * 128, 128>(pa, a_hi, a_lo, tid);
    store_rows<N, N>(pb, b_hi, b_lo, tid);
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint32_t sa_hi = umma::smem_u32(a_hi + b * kAFloats), sa_lo = umma::smem_u32(a_lo + b * kAFloats);
        const uint32_t sb_hi = umma::smem_u32(b_hi), sb_lo = umma::smem_u32(b_lo);
#pragma unroll
        for (int ps = 0; ps < 3; ++ps) {
          const uint32_t sa = (ps == 2) ? sa_lo : sa_hi;
          const uint32_t sb = (ps == 1) ? sb_lo : sb_hi;
#pragma unroll
          for (int j = 0; j < kWgS / 8; ++j) {
            const uint64_t da = umma::make_desc(sa + 2 * j * (128 * 16), 128 * 16, 128);
            const uint64_t db = umma::make_desc(sb + 2 * j * (N * 16), N * 16, 128);
            umma::mma_tf32(tmem + b * N, da, db, idesc, (it > 0 || ps > 0 || j > 0) ? 1u : 0u);
          }
        }
      }
      umma::commit(bar);
    }
    // the next chunk's global loads are in flight while the tensor core works on this one
    if (c + (int)gridDim.x < nchunks) {
      prefetch_rows<NB * 128>(pa, XT, ldx, K, (c + gridDim.x) * kWgS, P, tid);
      prefetch_rows<N>(pb, DYT, ldy, Nvalid, (c + gridDim.x) * kWgS, P, tid);
    }
  }
  if (it > 0) {
    umma::mbar_wait(bar, (it - 1) & 1);
    umma::fence_after_sync();
    // epilogue: warp w -> TMEM lanes 32*(w%4).., column halves split over the two warp groups
    const int q = warp & 3;
    for (int cb = (warp >> 2); cb * 32 < NB * N; cb += 2) {
      float r[32];
      umma::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + cb * 32, r);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int col = cb * 32 + j;
        const int b = col / N, n = col % N;
        const int k = b * 128 + q * 32 + lane;
        if (k < K && n < Nvalid) atomicAdd(dw + (size_t)n * dw_ld + k, r[j]);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, kCols);
}

static int g_sm = 0;
static inline int smc() {
  if (!g_sm) {
    int d = 0;
    cudaGetDevice(&d);
    cudaDeviceGetAttribute(&g_sm, cudaDevAttrMultiProcessorCount, d);
    if (g_sm <= 0) g_sm = 148;
  }
  return g_sm;
}

template <int NB, int N>
static int launch_wgrad(const float* XT, int ldx, int K, const float* DYT, int ldy, int Nvalid, int Pmax, const int32_t* cnt, float* dw,
                        int dw_ld, cudaStream_t st) {
  const size_t sm = (size_t)(2 * NB * 128 * kWgS + 2 * N * kWgS) * sizeof(float) + 64;
  cudaError_t e = cudaFuncSetAttribute(k_wgrad_tc<NB, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return (int)e;
  const int nchunks = (Pmax + kWgS - 1) / kWgS;
  const int grid = nchunks < 2 * smc() ? nchunks : 2 * smc();   // two CTAs per SM: one converts / issues while the other's loads are in flight
  k_wgrad_tc<NB, N><<<grid, kWgThreads, sm, st>>>(XT, ldx, K, DYT, ldy, Nvalid, Pmax, cnt, dw, dw_ld);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace dynrf

using namespace dynrf;

extern "C" int dynrf_wgrad_gemm(const float* XT, int ldx, int K, const float* DYT, int ldy, int N, int Pmax, const int32_t* count_ptr,
                                float* dw, int dw_ld, void* stream) {
  if (!XT || !DYT || !dw || K <= 0 || K > 256 || N <= 0 || N > 128 || Pmax <= 0 || ldx < Pmax || ldy < Pmax) return DYNRF_E_ARG;
  if (((uintptr_t)XT & 15) || ((uintptr_t)DYT & 15) || (ldx & 3) || (ldy & 3)) return DYNRF_E_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int nb = K <= 128 ? 1 : 2;
  const int np = N <= 16 ? 16 : (N <= 32 ? 32 : (N <= 64 ? 64 : 128));
#define WG(NB_, N_) return launch_wgrad<NB_, N_>(XT, ldx, K, DYT, ldy, N, Pmax, count_ptr, dw, dw_ld, st)
  if (nb == 1) {
    if (np == 16) WG(1, 16);
    if (np == 32) WG(1, 32);
    if (np == 64) WG(1, 64);
    WG(1, 128);
  } else {
    if (np == 16) WG(2, 16);
    if (np == 32) WG(2, 32);
    if (np == 64) WG(2, 64);
    WG(2, 128);
  }
#undef WG
}
