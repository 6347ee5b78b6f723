// tcgen05 (5th-gen tensor core) helpers for the tile-MLP engine: TMEM allocation, shared-memory
// matrix descriptors for the K-major no-swizzle canonical layout, kind::tf32 MMA issue, commit /
// mbarrier waits and TMEM loads.  sm_100a only.
//
// Precision: the reference computes its MLPs in true fp32 (TF32 off) and the parity budget is 1e-5
// on alpha, so single-pass TF32 (10-bit mantissa) is not enough for the sigma / blending / warp MLPs.
// Operands are split a = hi + lo with hi = rn_tf32(a) and lo = rn_tf32(a - hi) (cvt.rna.tf32.f32); three MMAs hi*hi + hi*lo + lo*hi accumulate in fp32 TMEM (error ~2^-20).
//
// Canonical K-major INTERLEAVE layout (cute::UMMA, mma_traits_sm100.hpp): a core matrix is 8 rows x
// 16 bytes, rows 16 B apart; 8-row groups are SBO bytes apart; the two 16-byte K chunks of one MMA
// K-step (K = 8 tf32) are LBO bytes apart.  We store an operand tile of R rows as
//   off(row, k) = (k/4) * R*16 + (row/8) * 128 + (row%8) * 16 + (k%4) * 4          [bytes]
// i.e. SBO = 128, LBO = R*16: a thread that owns one row writes 4 consecutive k as one 16-byte
// store, and 32 consecutive rows cover 512 contiguous bytes (bank-conflict free).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dynrf {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t kmajor_off(int row, int k, int R) {
  return (uint32_t)((k >> 2) * (R * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

// hi / lo split of an fp32 value into two tf32-representable parts
__device__ __forceinline__ float rn_tf32(float a) {
  // == cvt.rna.tf32.f32 for finite inputs (round to nearest, ties away from zero; low 13 mantissa bits cleared).  ptxas
  // expands the cvt into IADD + LOP3 + an inf/nan guard (FSETP + SEL); activations and weights here are finite, and the
  // guard doubles the cost of the split, which is ~7 % of the instructions of the tile kernels and most of k_wgrad_tc's.
  return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = rn_tf32(a);        // |a - hi| <= 2^-11 |a|
  lo = rn_tf32(a - hi);   // residual <= 2^-22 |a|, unbiased
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell); layout_type 0 = SWIZZLE_NONE, base_offset 0
  return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, dense, no negate
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// one warp allocates `ncols` (power of two >= 32) TMEM columns; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

// arrive on `bar` when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// producer-side wait: polls with a short sleep between attempts so that 16 spinning warps do not take issue slots
// from the tensor-core warp and from the warps that still have work
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  while (!done) {
    __nanosleep(40);
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns starting at taddr (lane field = warp's lane base)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&r)[32]) {
  uint32_t u[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}

// 16-column variant
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
  uint32_t u[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}

}  // namespace umma
}  // namespace dynrf
