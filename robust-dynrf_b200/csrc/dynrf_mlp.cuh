// Tile MLP engine: a CTA of 256 threads pushes a tile of TM (128 forward, 64 backward) samples
// through dense layers.
//
// Activations live in shared memory feature-major: A[k][m], m = sample in tile, row stride LD.
// Weights are the packed [Kpad][Npad] matrices (include/dynrf_b200.h) streamed from global/L2
// into a double-buffered shared-memory ring with cp.async in chunks of kChunkK input rows,
// so layers of any size work with a fixed smem budget.  fp32 FFMA, fp32 accumulate:
// the reference runs true-fp32 nn.Linear (TF32 off), tolerance budget 1e-5 on alpha.
//
// Thread tile: thread (tm = tid & 15, tn = tid >> 4) owns rows {4tm..4tm+3} (+64 when TM = 128)
// and columns {tn*NC .. tn*NC+NC-1}, NC = N/16.  A reads are conflict-free 16-byte loads
// (16 lanes cover 256 contiguous bytes), W reads are 2-address warp broadcasts.
//
// Backward additions: wgrad_acc (dW[k][n] += sum_m X[k][m] dY[n][m], LD = TM+4 keeps the
// row-strided 16-byte reads conflict-free) and vector-atomic flushes into packed gradients.
#pragma once
#include "dynrf_common.cuh"

namespace dynrf {

constexpr int kChunkK = 32;

template <int N, int TM>
struct Acc {
  static constexpr int NC = N / 16;
  static constexpr int MR = TM / 16;
  float v[MR][NC];
};

template <int N, int TM>
__device__ __forceinline__ void acc_init(Acc<N, TM>& a, const float* __restrict__ bias_g, int tn) {
#pragma unroll
  for (int j = 0; j < Acc<N, TM>::NC; ++j) {
    const float b = bias_g ? __ldg(bias_g + tn * Acc<N, TM>::NC + j) : 0.f;
#pragma unroll
    for (int i = 0; i < Acc<N, TM>::MR; ++i) a.v[i][j] = b;
  }
}

template <int N>
__device__ __forceinline__ void load_w_chunk(float* dst, const float* __restrict__ src, int rows, int tid) {
  const int n16 = rows * (N / 4);  // rows*N floats, contiguous in both; 16 B per cp.async
  for (int i = tid; i < n16; i += kMlpThreads) cp_async16(dst + i * 4, src + i * 4);
}

// acc += A[0..K) x Wt[0..K)   (K multiple of 4).  wbuf: 2 * kChunkK * N floats of shared memory.
// Ends with __syncthreads(): A and wbuf may be overwritten by the caller afterwards.
template <int N, int TM, int LD>
__device__ __forceinline__ void acc_gemm(Acc<N, TM>& a, const float* __restrict__ A, int K, const float* __restrict__ Wt_g,
                                         float* wbuf, int tid) {
  constexpr int NC = Acc<N, TM>::NC;
  constexpr int MR = Acc<N, TM>::MR;
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
    const float* __restrict__ Ak = A + (size_t)k0 * LD + tm * 4;
#pragma unroll 4
    for (int k = 0; k < rows; ++k) {
      float am[MR];
#pragma unroll
      for (int h = 0; h < MR / 4; ++h) {
        const float4 t = *reinterpret_cast<const float4*>(Ak + k * LD + 64 * h);
        am[4 * h] = t.x; am[4 * h + 1] = t.y; am[4 * h + 2] = t.z; am[4 * h + 3] = t.w;
      }
      float b[NC];
      if constexpr (NC % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NC; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(W + k * N + j);
          b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
        }
      } else if constexpr (NC % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NC; j += 2) {
          const float2 t = *reinterpret_cast<const float2*>(W + k * N + j);
          b[j] = t.x; b[j + 1] = t.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) b[j] = W[k * N + j];
      }
#pragma unroll
      for (int i = 0; i < MR; ++i)
#pragma unroll
        for (int j = 0; j < NC; ++j) a.v[i][j] = fmaf(am[i], b[j], a.v[i][j]);
    }
    __syncthreads();
  }
}

// Out[n][m] = act(acc) for n < n_valid.  MASK (optional, same layout as Out): multiply by [MASK > 0]
// (relu backward).  Ends with __syncthreads().
template <int N, int TM, int LD, bool RELU, bool MASKED = false>
__device__ __forceinline__ void acc_store(const Acc<N, TM>& a, float* Out, int n_valid, int tid, const float* Mask = nullptr) {
  constexpr int NC = Acc<N, TM>::NC;
  constexpr int MR = Acc<N, TM>::MR;
  const int tm = tid & 15, tn = tid >> 4;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    const int n = tn * NC + j;
    if (n < n_valid) {
#pragma unroll
      for (int h = 0; h < MR / 4; ++h) {
        float4 v = make_float4(a.v[4 * h][j], a.v[4 * h + 1][j], a.v[4 * h + 2][j], a.v[4 * h + 3][j]);
        if (RELU) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        float* dst = Out + (size_t)n * LD + tm * 4 + 64 * h;
        if (MASKED) {
          const float4 mk = *reinterpret_cast<const float4*>(Mask + (size_t)n * LD + tm * 4 + 64 * h);
          v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f; v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
        }
        *reinterpret_cast<float4*>(dst) = v;
      }
    }
  }
  __syncthreads();
}

// Full layer: Out = act(A x Wt + bias).
template <int N, int TM, int LD, bool RELU>
__device__ __forceinline__ void dense_tile(const float* A, const dynrf_layer_t& L, float* Out, float* wbuf, int tid) {
  Acc<N, TM> acc;
  acc_init<N, TM>(acc, L.bias, tid >> 4);
  acc_gemm<N, TM, LD>(acc, A, L.Kpad, L.wt, wbuf, tid);
  acc_store<N, TM, LD, RELU>(acc, Out, N, tid);
}

// Narrow output layer (NOUT <= 4): thread m computes out[o] = bias[o] + sum_k A[k][m] * Wt[k][o].
template <int NOUT, int LD>
__device__ __forceinline__ void small_dense(const float* __restrict__ A, int K, const dynrf_layer_t& L, int m, float (&out)[NOUT]) {
#pragma unroll
  for (int o = 0; o < NOUT; ++o) out[o] = __ldg(L.bias + o);
  const float* __restrict__ w = L.wt;
  const int ld = L.Npad;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float a = A[(size_t)k * LD + m];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) out[o] = fmaf(a, __ldg(w + (size_t)k * ld + o), out[o]);
  }
}

// positional encoding rows for one coordinate: sin at rows [sin0 .. sin0+F), cos at [cos0 .. cos0+F)
template <int F, int LD>
__device__ __forceinline__ void pe_rows(float x, float* __restrict__ A, int sin0, int cos0, int m) {
  float sc = 1.0f;
#pragma unroll
  for (int f = 0; f < F; ++f) {
    float s, c;
    sincosf(x * sc, &s, &c);  // x * 2^f is exact
    A[(size_t)(sin0 + f) * LD + m] = s;
    A[(size_t)(cos0 + f) * LD + m] = c;
    sc *= 2.0f;
  }
}

// ------------------------------------------------------------------------------------------------
// backward helpers (TM = 64, LD = 68)

// dW[k][n] += sum_m X[k][m] * dY[n][m].  Thread (kg = tid & 15, ng = tid >> 4) owns rows k = kg + 16 i
// (i < KT) and columns n = ng*NT .. ng*NT+NT-1 of the packed [Kpad][Npad] gradient.
template <int KT, int NT, int TM, int LD>
__device__ __forceinline__ void wgrad_acc(float (&acc)[KT][NT], const float* __restrict__ X, int Krows,
                                          const float* __restrict__ dY, int tid) {
  const int kg = tid & 15, ng = tid >> 4;
#pragma unroll 2
  for (int m = 0; m < TM; m += 4) {
    float4 y[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) y[j] = *reinterpret_cast<const float4*>(dY + (size_t)(ng * NT + j) * LD + m);
#pragma unroll
    for (int i = 0; i < KT; ++i) {
      const int k = kg + 16 * i;
      if (k < Krows) {
        const float4 x = *reinterpret_cast<const float4*>(X + (size_t)k * LD + m);
#pragma unroll
        for (int j = 0; j < NT; ++j)
          acc[i][j] = fmaf(x.x, y[j].x, fmaf(x.y, y[j].y, fmaf(x.z, y[j].z, fmaf(x.w, y[j].w, acc[i][j]))));
      }
    }
  }
}

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add2(float* addr, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// flush a thread's wgrad block into the packed gradient dWt[Kpad][Npad] and clear it
template <int KT, int NT>
__device__ __forceinline__ void wgrad_flush(float (&acc)[KT][NT], float* __restrict__ dWt, int Krows, int Npad, int tid) {
  const int kg = tid & 15, ng = tid >> 4;
#pragma unroll
  for (int i = 0; i < KT; ++i) {
    const int k = kg + 16 * i;
    if (k < Krows) {
      float* p = dWt + (size_t)k * Npad + ng * NT;
      if constexpr (NT % 4 == 0) {
#pragma unroll
        for (int j = 0; j < NT; j += 4) red_add4(p + j, acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
      } else if constexpr (NT % 2 == 0) {
#pragma unroll
        for (int j = 0; j < NT; j += 2) red_add2(p + j, acc[i][j], acc[i][j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < NT; ++j) atomicAdd(p + j, acc[i][j]);
      }
    }
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
  }
}

template <int KT, int NT>
__device__ __forceinline__ void wgrad_zero(float (&acc)[KT][NT]) {
#pragma unroll
  for (int i = 0; i < KT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[i][j] = 0.f;
}

}  // namespace dynrf
