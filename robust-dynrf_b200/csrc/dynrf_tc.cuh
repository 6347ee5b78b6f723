// Shared pieces of the warp-specialised tcgen05 tile kernels: split-TF32 operand production in the K-major
// UMMA layout, the operand ring (3 buffers) + TMA weight slots (2), their mbarrier protocol, and the
// position/time operand units.  See tc_kernels.cu for the protocol description.
#pragma once
#include "dynrf_tile.cuh"
#include "dynrf_umma.cuh"

namespace dynrf {

constexpr int kKC = 32;  // input rows per MMA chunk

__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(umma::smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}

constexpr int kTcThreads = 512;
constexpr int kOpFloats = 2 * kKC * kTileM;  // hi + lo, 32 rows x 128 samples
constexpr int kOpLo = kKC * kTileM;          // float offset of the lo half
constexpr int kWSlotFloats = 2 * 128 * kKC;  // largest weight chunk (N = 128, 32 rows)
struct TapF {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ TapF tapf(float x, int n, int s) {
  const Tap1 t = make_tap(x, n, s);
  return TapF{t.i0, t.i1, t.w0, t.w1};
}
__device__ __forceinline__ float4 vm_group_f(const float* __restrict__ plane, const float* __restrict__ line, int C, PlaneSt st,
                                             const TapF& tu, const TapF& tv, const TapF& tl, int c) {
  const float wnw = tu.w0 * tv.w0, wne = tu.w1 * tv.w0, wsw = tu.w0 * tv.w1, wse = tu.w1 * tv.w1;
  const size_t r0 = (size_t)tv.i0 * st.sv, r1 = (size_t)tv.i1 * st.sv;
  const size_t c0 = (size_t)tu.i0 * st.su, c1 = (size_t)tu.i1 * st.su;
  float4 p = mul4(ldg4(plane + (r0 + c0) * C + c), wnw);
  p = fma4(ldg4(plane + (r0 + c1) * C + c), wne, p);
  p = fma4(ldg4(plane + (r1 + c0) * C + c), wsw, p);
  p = fma4(ldg4(plane + (r1 + c1) * C + c), wse, p);
  float4 l = mul4(ldg4(line + (size_t)tl.i0 * C + c), tl.w0);
  l = fma4(ldg4(line + (size_t)tl.i1 * C + c), tl.w1, l);
  return mul44(p, l);
}

// write 4 consecutive operand rows (quad q of the chunk) for sample m as split tf32, one 16-byte store each
// feature-major global copy of one quad: rows r0..r0+3 of gT (row stride ld) at column `col`
// The feature-major activation streams (0.3 - 1.5 GB per launch at cfg 2) are written once and read once by a later kernel; with
// DYNRF_STREAM_HINTS they use the streaming (evict-first) cache policy so that they do not push the parameter tables and the
// gradient tables -- the targets of the gathers and of the L2 reductions -- out of the 126 MB L2.
#ifdef DYNRF_STREAM_HINTS
#define DYNRF_ST_STREAM(ptr, val) __stcs((ptr), (val))
#define DYNRF_LD_STREAM(ptr) __ldcs(ptr)
#else
#define DYNRF_ST_STREAM(ptr, val) (*(ptr) = (val))
#define DYNRF_LD_STREAM(ptr) __ldg(ptr)
#endif

__device__ __forceinline__ void store_quad_g(float* __restrict__ gT, int ld, int r0, int col, float4 v) {
  float* p = gT + (size_t)r0 * ld + col;
  DYNRF_ST_STREAM(p, v.x);
  DYNRF_ST_STREAM(p + ld, v.y);
  DYNRF_ST_STREAM(p + 2 * (size_t)ld, v.z);
  DYNRF_ST_STREAM(p + 3 * (size_t)ld, v.w);
}

__device__ __forceinline__ void store_quad(float* __restrict__ buf, int m, int q, float4 v) {
  float4 hi, lo;
  umma::split_tf32(v.x, hi.x, lo.x);
  umma::split_tf32(v.y, hi.y, lo.y);
  umma::split_tf32(v.z, hi.z, lo.z);
  umma::split_tf32(v.w, hi.w, lo.w);
  const uint32_t off = umma::kmajor_off(m, q * 4, kTileM) >> 2;
  *reinterpret_cast<float4*>(buf + off) = hi;
  *reinterpret_cast<float4*>(buf + kOpLo + off) = lo;
}

__device__ __forceinline__ void sincos4(float x, float base, float4& s, float4& c) {
  sincosf(x * base, &s.x, &c.x);
  sincosf(x * (base * 2.f), &s.y, &c.y);
  sincosf(x * (base * 4.f), &s.z, &c.z);
  sincosf(x * (base * 8.f), &s.w, &c.w);
}

// position/time operand rows, unit u of 14 (see K order above); writes into the chunk buffer that holds its quads
// gT != nullptr: additionally stream the quads feature-major to gT (rows = this chunk's rows, row stride ld) at column gcol
template <class Scr>
__device__ __forceinline__ void produce_pos_unit(int u, const Scr& scr, const float* __restrict__ tfeat, int m,
                                                 float* __restrict__ buf, float* __restrict__ gT = nullptr, int ld = 0, int gcol = 0) {
#define DYNRF_PUT(q_, v_)                                        \
  do {                                                           \
    const float4 _v = (v_);                                      \
    store_quad(buf, m, (q_), _v);                                \
    if (gT) store_quad_g(gT, ld, (q_) * 4, gcol, _v);             \
  } while (0)
  // chunk 0: u0 quad0 | u1 (d0,fq0) quads1,2 | u2 (d0,fq1) quads3,4 | u3 (d0,tail) quad5 | u4 (d1,fq0) quads6,7
  // chunk 1: u5 (d1,fq1) quads0,1 | u6 (d1,tail) quad2 | u7 (d2,fq0) quads3,4 | u8 (d2,fq1) quads5,6 | u9 (d2,tail) quad7
  // chunk 2: u10..u13 time quads 0..3
  const int ry = scr.ray[m];
  if (u == 0) {
    const float t = ry >= 0 ? __ldg(tfeat + (size_t)ry * kTFeat) : 0.f;
    DYNRF_PUT(0, make_float4(scr.xn[0][m], scr.xn[1][m], scr.xn[2][m], t));
    return;
  }
  if (u >= 10) {
    const int tq = u - 10;  // 0: sin f0-3, 1: sin f4-7, 2: cos f0-3, 3: cos f4-7   (tfeat cols 1..8 sin, 9..16 cos)
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ry >= 0) {
      const float* p = tfeat + (size_t)ry * kTFeat + 1 + tq * 4;
      v = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
    }
    DYNRF_PUT(tq, v);
    return;
  }
  // (dim, kind) and destination quad
  int d, kind, q;  // kind 0: fq0, 1: fq1, 2: tail
  switch (u) {
    case 1: d = 0; kind = 0; q = 1; break;
    case 2: d = 0; kind = 1; q = 3; break;
    case 3: d = 0; kind = 2; q = 5; break;
    case 4: d = 1; kind = 0; q = 6; break;
    case 5: d = 1; kind = 1; q = 0; break;
    case 6: d = 1; kind = 2; q = 2; break;
    case 7: d = 2; kind = 0; q = 3; break;
    case 8: d = 2; kind = 1; q = 5; break;
    default: d = 2; kind = 2; q = 7; break;
  }
  const float x = scr.xn[d][m];
  if (kind == 2) {
    float s8, c8, s9, c9;
    sincosf(x * 256.f, &s8, &c8);
    sincosf(x * 512.f, &s9, &c9);
    DYNRF_PUT(q, make_float4(s8, s9, c8, c9));
  } else {
    float4 s, c;
    sincos4(x, kind == 0 ? 1.f : 16.f, s, c);
    DYNRF_PUT(q, s);
    DYNRF_PUT(q + 1, c);
  }
}
#undef DYNRF_PUT

struct Ring {
  uint64_t* bar_free;  // [nbuf] operand buffer consumed (MMAs of the chunk that used it completed)
  uint64_t* bar_w;     // [nw] weight slot landed
  float* abuf;         // nbuf x kOpFloats
  float* wbuf;         // nw x wslot
  uint32_t chunk;      // running chunk counter, identical in every thread
  uint32_t tmem;
  uint32_t wslot = kWSlotFloats;  // floats per weight slot
  uint32_t nbuf = 3;              // operand buffers
  uint32_t nw = 2;                // weight slots: chunk c's weights are requested while chunk c - (nw - 1) is being issued
  // "the MMAs of chunk c have completed" is signalled on bar_free[c % nfree].  nfree == nbuf (default) ties the barrier to the
  // operand buffer: correct only if every warp waits at every chunk, because a parity wait on a barrier that is two phases behind
  // succeeds at once.  The kernels whose warps wait only for the chunks they write (two-tile kernels) use nfree = kNFull.
  uint32_t nfree = 3;
};

// wait until the operand buffer of the current chunk is free, i.e. until chunk (c - nbuf) has been consumed
__device__ __forceinline__ void ring_wait_free(const Ring& r) {
  if (r.chunk >= r.nbuf) {
    const uint32_t pc = r.chunk - r.nbuf;
    umma::mbar_wait_backoff(&r.bar_free[pc % r.nfree], (pc / r.nfree) & 1);
  }
}
__device__ __forceinline__ float* ring_abuf(const Ring& r) { return r.abuf + (size_t)(r.chunk % r.nbuf) * kOpFloats; }

// tid 0: weights of chunk `c` (global blob) -> slot c % nw.  The slot was last read by chunk c - nw.
__device__ __forceinline__ void ring_prefetch_w(const Ring& r, uint32_t c, const float* blob, uint32_t bytes) {
  if (c >= r.nw) {
    const uint32_t pc = c - r.nw;
    umma::mbar_wait(&r.bar_free[pc % r.nfree], (pc / r.nfree) & 1);
  }
  const uint32_t slot = c % r.nw;
  mbar_expect_tx(&r.bar_w[slot], bytes);
  tma_load_1d(r.wbuf + (size_t)slot * r.wslot, blob, bytes, &r.bar_w[slot]);
}

// tid 0: wait for this chunk's weights, issue the 3-pass MMAs, commit to the operand buffer's barrier
template <int N>
__device__ __forceinline__ void ring_issue(const Ring& r, int rows, uint32_t col0, bool first) {
  const uint32_t c = r.chunk;
  const uint32_t slot = c % r.nw, buf = c % r.nbuf;
  umma::mbar_wait(&r.bar_w[slot], (c / r.nw) & 1);
  umma::fence_after_sync();
  const uint32_t idesc = umma::make_idesc_tf32(kTileM, N);
  const uint32_t sa_hi = umma::smem_u32(r.abuf + (size_t)buf * kOpFloats), sa_lo = sa_hi + kOpLo * 4;
  const uint32_t sb_hi = umma::smem_u32(r.wbuf + (size_t)slot * r.wslot), sb_lo = sb_hi + N * rows * 4;
#pragma unroll
  for (int ps = 0; ps < 3; ++ps) {
    const uint32_t sa = (ps == 2) ? sa_lo : sa_hi;
    const uint32_t sb = (ps == 1) ? sb_lo : sb_hi;
    for (int j = 0; j < rows / 8; ++j) {
      const uint64_t da = umma::make_desc(sa + 2 * j * (kTileM * 16), kTileM * 16, 128);
      const uint64_t db = umma::make_desc(sb + 2 * j * (N * 16), N * 16, 128);
      umma::mma_tf32(r.tmem + col0, da, db, idesc, (!first || ps > 0 || j > 0) ? 1u : 0u);
    }
  }
  umma::commit(&r.bar_free[c % r.nfree]);
}

// row 0..79 of the position/time operand block (this engine's K order, see tc_kernels.cu) -> index 0..79 in the
// reference's [xn(3), sin PE(30), cos PE(30), t, sin PE(t)(8), cos PE(t)(8)]
__device__ __forceinline__ int pos_src_index(int r) {  // row 0..79 of the position/time block -> reference index 0..79
  const int qd = r >> 2, e = r & 3;
  if (qd == 0) return e < 3 ? e : 63;
  if (qd >= 16) {
    const int tq = qd - 16;
    return (tq < 2 ? 64 : 72) + (tq & 1) * 4 + e;
  }
  const int d = (qd - 1) / 5, w = (qd - 1) % 5;
  if (w == 4) return (e < 2 ? 3 : 33) + d * 10 + 8 + (e & 1);
  return ((w & 1) ? 33 : 3) + d * 10 + (w >> 1) * 4 + e;
}


// v[c] holds this lane's value of column c; returns (in lane L) the sum over the 32 lanes of column L
__device__ __forceinline__ float warp_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// number of "chunk written" mbarriers of the two-tile kernels (chunk c arrives on barrier c % kNFull): decoupled from the number
// of operand buffers so that a warp which does not write a chunk need not wait for the chunk's buffer before arriving
constexpr int kNFull = 8;
constexpr int kProdThreads = 512;
constexpr int kTc3Threads = kProdThreads + 32;

struct TapsSmem {
  int4 t[9][kTileM];  // [stride*3 + axis][sample]: (i0, i1, w0 bits, w1 bits)
};

__device__ __forceinline__ TapF load_tap(const TapsSmem& ts, int si, int axis, int m) {
  const int4 v = ts.t[si * 3 + axis][m];
  return TapF{v.x, v.y, __int_as_float(v.z), __int_as_float(v.w)};
}

__device__ __forceinline__ void prod_bar() { asm volatile("bar.sync 1, %0;\n" ::"n"(kProdThreads) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(umma::smem_u32(bar)) : "memory");
}


}  // namespace dynrf
