// Tile MLP engine: a CTA of 256 threads pushes a tile of 128 samples through dense layers.
//
// Activations live in shared memory feature-major: A[k][m], m = sample in tile (LD = 128).
// Weights are the packed [Kpad][Npad] matrices (include/dynrf_b200.h) streamed from global/L2
// into a double-buffered shared-memory ring with cp.async in chunks of kChunkK input rows,
// so layers of any size work with a fixed smem budget.  fp32 FFMA, fp32 accumulate:
// the reference runs true-fp32 nn.Linear (TF32 off), tolerance budget 1e-5 on alpha.
//
// Thread tile: thread (tm = tid & 15, tn = tid >> 4) owns rows {4tm..4tm+3, 64+4tm..64+4tm+3}
// and columns {tn*NC .. tn*NC+NC-1}, NC = N/16.  A reads are conflict-free 16-byte loads
// (16 lanes cover 256 contiguous bytes), W reads are 2-address warp broadcasts.
#pragma once
#include "dynrf_common.cuh"

namespace dynrf {

constexpr int kChunkK = 32;

template <int N>
struct Acc {
  static constexpr int NC = N / 16;
  float v[8][NC];
};

template <int N>
__device__ __forceinline__ void acc_init(Acc<N>& a, const float* __restrict__ bias_g, int tn) {
#pragma unroll
  for (int j = 0; j < Acc<N>::NC; ++j) {
    const float b = bias_g ? __ldg(bias_g + tn * Acc<N>::NC + j) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a.v[i][j] = b;
  }
}

template <int N>
__device__ __forceinline__ void load_w_chunk(float* dst, const float* __restrict__ src, int rows, int tid) {
  // rows*N floats, contiguous in both; 16 B per cp.async
  const int n16 = rows * (N / 4);
  for (int i = tid; i < n16; i += kMlpThreads) cp_async16(dst + i * 4, src + i * 4);
}

// acc += A[0..K) x Wt[0..K)   (K multiple of 4).  wbuf: 2 * kChunkK * N floats of shared memory.
// Ends with __syncthreads(): A and wbuf may be overwritten by the caller afterwards.
template <int N>
__device__ __forceinline__ void acc_gemm(Acc<N>& a, const float* __restrict__ A, int K, const float* __restrict__ Wt_g,
                                         float* wbuf, int tid) {
  constexpr int NC = Acc<N>::NC;
  const int tm = tid & 15, tn = tid >> 4;
  const int nchunks = (K + kChunkK - 1) / kChunkK;
  load_w_chunk<N>(wbuf, Wt_g, min(kChunkK, K), tid);
  cp_async_commit();
  for (int c = 0; c < nchunks; ++c) {
    const int k0 = c * kChunkK;
    const int rows = min(kChunkK, K - k0);
    if (c + 1 < nchunks) {
      load_w_chunk<N>(wbuf + ((c + 1) & 1) * kChunkK * N, Wt_g + (size_t)(k0 + kChunkK) * N, min(kChunkK, K - k0 - kChunkK), tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* __restrict__ W = wbuf + (c & 1) * kChunkK * N + tn * NC;
    const float* __restrict__ Ak = A + (size_t)k0 * kTileM + tm * 4;
#pragma unroll 4
    for (int k = 0; k < rows; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(Ak + k * kTileM);
      const float4 a1 = *reinterpret_cast<const float4*>(Ak + k * kTileM + 64);
      float b[NC];
      if constexpr (NC % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NC; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(W + k * N + j);
          b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
        }
      } else if constexpr (NC == 2) {
        const float2 t = *reinterpret_cast<const float2*>(W + k * N);
        b[0] = t.x; b[1] = t.y;
      } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) b[j] = W[k * N + j];
      }
      const float am[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) a.v[i][j] = fmaf(am[i], b[j], a.v[i][j]);
    }
    __syncthreads();
  }
}

// Out[n][m] = act(acc) for n < n_valid.  Ends with __syncthreads().
template <int N, bool RELU>
__device__ __forceinline__ void acc_store(const Acc<N>& a, float* __restrict__ Out, int n_valid, int tid) {
  constexpr int NC = Acc<N>::NC;
  const int tm = tid & 15, tn = tid >> 4;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int n = tn * NC + j;
    if (n < n_valid) {
      float4 lo = make_float4(a.v[0][j], a.v[1][j], a.v[2][j], a.v[3][j]);
      float4 hi = make_float4(a.v[4][j], a.v[5][j], a.v[6][j], a.v[7][j]);
      if (RELU) {
        lo.x = fmaxf(lo.x, 0.f); lo.y = fmaxf(lo.y, 0.f); lo.z = fmaxf(lo.z, 0.f); lo.w = fmaxf(lo.w, 0.f);
        hi.x = fmaxf(hi.x, 0.f); hi.y = fmaxf(hi.y, 0.f); hi.z = fmaxf(hi.z, 0.f); hi.w = fmaxf(hi.w, 0.f);
      }
      *reinterpret_cast<float4*>(Out + (size_t)n * kTileM + tm * 4) = lo;
      *reinterpret_cast<float4*>(Out + (size_t)n * kTileM + 64 + tm * 4) = hi;
    }
  }
  __syncthreads();
}

// Full layer: Out = act(A x Wt + bias).
template <int N, bool RELU>
__device__ __forceinline__ void dense_tile(const float* A, const dynrf_layer_t& L, float* Out, float* wbuf, int tid) {
  Acc<N> acc;
  acc_init<N>(acc, L.bias, tid >> 4);
  acc_gemm<N>(acc, A, L.Kpad, L.wt, wbuf, tid);
  acc_store<N, RELU>(acc, Out, N, tid);
}

// Narrow output layer (NOUT <= 4): thread m (< 128) computes out[o] = bias[o] + sum_k A[k][m] * Wt[k][o].
template <int NOUT>
__device__ __forceinline__ void small_dense(const float* __restrict__ A, int K, const dynrf_layer_t& L, int m, float (&out)[NOUT]) {
#pragma unroll
  for (int o = 0; o < NOUT; ++o) out[o] = __ldg(L.bias + o);
  const float* __restrict__ w = L.wt;
  const int ld = L.Npad;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = A[(size_t)k * kTileM + m];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) out[o] = fmaf(a, __ldg(w + (size_t)k * ld + o), out[o]);
  }
}

// positional encoding rows for one coordinate: sin at rows [sin0 .. sin0+F), cos at [cos0 .. cos0+F)
template <int F>
__device__ __forceinline__ void pe_rows(float x, float* __restrict__ A, int sin0, int cos0, int m) {
  float sc = 1.0f;
#pragma unroll
  for (int f = 0; f < F; ++f) {
    float s, c;
    sincosf(x * sc, &s, &c);  // x * 2^f is exact
    A[(size_t)(sin0 + f) * kTileM + m] = s;
    A[(size_t)(cos0 + f) * kTileM + m] = c;
    sc *= 2.0f;
  }
}

}  // namespace dynrf
