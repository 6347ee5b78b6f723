// TEST INFRASTRUCTURE (built into tests/lib/libdynrf_selftest.so, not into the product library).
// Self-test entry point for the tcgen05 tile GEMM building block (used by tests/test_gpu_umma.py):
// C[128][N] = A[128][K] x B[N][K]^T with operands staged in the K-major no-swizzle UMMA layout,
// 1-pass TF32 or 3-pass split-TF32, accumulation in TMEM, K processed in chunks of 32.
#include "../../robust-dynrf_b200/csrc/dynrf_common.cuh"
#include "../../robust-dynrf_b200/csrc/dynrf_umma.cuh"

namespace dynrf {

constexpr int kSelfM = 128;
constexpr int kSelfKC = 32;

template <int N>
__global__ void __launch_bounds__(256) k_selftest_umma(const float* __restrict__ A, const float* __restrict__ B, int K, int passes,
                                                       float* __restrict__ C) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* a_hi = reinterpret_cast<float*>(smem);
  float* a_lo = a_hi + kSelfM * kSelfKC;
  float* b_hi = a_lo + kSelfM * kSelfKC;
  float* b_lo = b_hi + N * kSelfKC;
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + N * kSelfKC);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, N < 32 ? 32 : N);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tslot;
  const uint32_t idesc = umma::make_idesc_tf32(kSelfM, N);
  uint32_t parity = 0;
  const int nchunks = (K + kSelfKC - 1) / kSelfKC;
  for (int c = 0; c < nchunks; ++c) {
    const int k0 = c * kSelfKC;
    // stage A and B chunk (zero padded beyond K), 4 consecutive k per store
    for (int item = tid; item < kSelfM * (kSelfKC / 4); item += 256) {
      const int row = item % kSelfM, kq = item / kSelfM;
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq * 4 + j;
        const float v = k < K ? A[(size_t)row * K + k] : 0.f;
        umma::split_tf32(v, hi[j], lo[j]);
      }
      const uint32_t off = umma::kmajor_off(row, kq * 4, kSelfM) / 4;
      *reinterpret_cast<float4*>(a_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(a_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    for (int item = tid; item < N * (kSelfKC / 4); item += 256) {
      const int row = item % N, kq = item / N;
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq * 4 + j;
        const float v = k < K ? B[(size_t)row * K + k] : 0.f;
        umma::split_tf32(v, hi[j], lo[j]);
      }
      const uint32_t off = umma::kmajor_off(row, kq * 4, N) / 4;
      *reinterpret_cast<float4*>(b_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(b_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t sa_hi = umma::smem_u32(a_hi), sa_lo = umma::smem_u32(a_lo);
      const uint32_t sb_hi = umma::smem_u32(b_hi), sb_lo = umma::smem_u32(b_lo);
      for (int p = 0; p < passes; ++p) {
        const uint32_t sa = (p == 2) ? sa_lo : sa_hi;
        const uint32_t sb = (p == 1) ? sb_lo : sb_hi;
#pragma unroll
        for (int j = 0; j < kSelfKC / 8; ++j) {
          const uint64_t da = umma::make_desc(sa + 2 * j * (kSelfM * 16), kSelfM * 16, 128);
          const uint64_t db = umma::make_desc(sb + 2 * j * (N * 16), N * 16, 128);
          umma::mma_tf32(tmem, da, db, idesc, (c > 0 || p > 0 || j > 0) ? 1u : 0u);
        }
      }
      umma::commit(bar);
    }
    umma::mbar_wait(bar, parity);
    parity ^= 1;
    umma::fence_after_sync();
  }
  // epilogue: warp w reads TMEM lanes 32*(w%4).., column blocks of 32 split between the two warp groups
  const int lane_base = 32 * (warp & 3);
  for (int cb = (warp >> 2); cb * 32 < N; cb += 2) {
    float r[32];
    umma::tmem_ld32(tmem + ((uint32_t)lane_base << 16) + cb * 32, r);
    const int row = lane_base + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) C[(size_t)row * N + cb * 32 + j] = r[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, N < 32 ? 32 : N);
}

}  // namespace dynrf

using namespace dynrf;

extern "C" int dynrf_selftest_umma_gemm(const float* A, const float* B, int K, int N, int passes, float* C, void* stream) {
  if (!A || !B || !C || K <= 0 || (passes != 1 && passes != 3)) return DYNRF_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (N == 64) {
    const size_t sm = (size_t)(2 * kSelfM * kSelfKC + 2 * 64 * kSelfKC) * 4 + 64;
    cudaFuncSetAttribute(k_selftest_umma<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_selftest_umma<64><<<1, 256, sm, st>>>(A, B, K, passes, C);
  } else if (N == 128) {
    const size_t sm = (size_t)(2 * kSelfM * kSelfKC + 2 * 128 * kSelfKC) * 4 + 64;
    cudaFuncSetAttribute(k_selftest_umma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_selftest_umma<128><<<1, 256, sm, st>>>(A, B, K, passes, C);
  } else if (N == 32) {
    const size_t sm = (size_t)(2 * kSelfM * kSelfKC + 2 * 32 * kSelfKC) * 4 + 64;
    cudaFuncSetAttribute(k_selftest_umma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    k_selftest_umma<32><<<1, 256, sm, st>>>(A, B, K, passes, C);
  } else {
    return DYNRF_E_UNSUPPORTED;
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}
