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
#include "dynrf_tc.cuh"

namespace dynrf {

// packed UMMA weights of one layer: chunk c at wq + c * 2*N*32 floats: [hi N x 32 K-major][lo N x 32 K-major]
struct UmmaLayer {
  const float* wq;
  const float* bias;  // [N]
  int nchunks;
};

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
          UmmaLayer l4, dynrf_layer_t l5, float* __restrict__ xyz_prime, float* __restrict__ H1T /*[64][ld] optional*/,
          float* __restrict__ H2T /*[64][ld] optional*/, int ld) {
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
            const float h = fmaxf(r[q * 4 + j] + __ldg(l3.bias + n), 0.f);
            umma::split_tf32(h, hi[j], lo[j]);
            if (H1T) DYNRF_ST_STREAM(H1T + (size_t)n * ld + tile * kTileM + m, h);   // kept for the backward (k_warp_bwd)
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
          if (H2T) DYNRF_ST_STREAM(H2T + (size_t)n * ld + tile * kTileM + m, h);
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

// ================================================================================================
// v2: dynamic density + blending, producers write the split K-major operand directly (no fp32 staging),
// the shared position/time rows are multiplied once against the concatenated [density | blending] weights
// (N = 128), 512 threads, ring of 3 operand buffers + 2 TMA weight slots so the asynchronous MMAs and
// weight copies overlap the gathers of the next chunk.
//
// K order (this kernel's own; the pack kernel permutes the reference's columns accordingly):
//   chunk 0..2  position/time rows, 20 quads of 4:
//       quad 0            : xn0 xn1 xn2 t
//       quads 1+5d .. 5+5d: dim d: [sin f0-3] [cos f0-3] [sin f4-7] [cos f4-7] [sin f8, sin f9, cos f8, cos f9]
//       quads 16..19      : [sin t f0-3] [sin t f4-7] [cos t f0-3] [cos t f4-7]
//   chunk 3+2si   density  features of stride 2^si (24 rows, reference order)   -> TMEM columns   0..63
//   chunk 4+2si   blending features of stride 2^si                              -> TMEM columns  64..127

constexpr int kSbChunks = 9;

__host__ __device__ constexpr int sb_rows(int j) { return j < 2 ? 32 : (j == 2 ? 16 : 24); }
__host__ __device__ constexpr int sb_n(int j) { return j < 3 ? 128 : 64; }
__host__ __device__ constexpr int sb_woff(int j) {  // float offset of chunk j's weight blob
  int o = 0;
  for (int i = 0; i < j; ++i) o += 2 * sb_n(i) * sb_rows(i);
  return o;
}
constexpr int kSbWFloats = sb_woff(kSbChunks);

// ---- v3: warp-specialised.  16 producer warps (512 threads, 4 per sample) fill the operand ring and signal a
// per-buffer "full" mbarrier; one extra warp owns the tensor core: it waits for "full" + the TMA'd weights, issues
// the 3-pass MMAs, commits to the buffer's "free" barrier and prefetches the next weight chunk.  No CTA-wide
// barrier in the steady state; producers use a named barrier (id 1) among themselves for the per-tile scratch.
// The bilinear taps depend only on (axis, stride) -- 9 per sample -- and are computed once per tile into shared memory.

// Two tiles in flight (see tc_app_fwd.cu): iteration k produces the nine operand chunks of tile Y = k (context k & 1: its own
// scratch, taps and 128 TMEM columns) and then reads back / finishes tile X = k - 1, whose last MMAs have long retired by then
// (the read-back wait was 9.6 % of the stall samples of the one-tile kernel).  The chunk order in the ring is unchanged.
struct SbScratch {
  float xn[3][kTileM];  // normalize(xyz)
  float xq[3][kTileM];  // normalize(xyz_prime)
  int ray[kTileM], on[kTileM];
};
constexpr size_t kSb2Smem = (size_t)(3 * kOpFloats + 2 * kWSlotFloats + 2 * 4 * kTileM + 256) * sizeof(float) + 2 * sizeof(TapsSmem) +
                            2 * sizeof(SbScratch) + 128;
static_assert(kSb2Smem <= 232448, "k_dyn_sigma_blend_tc2 shared memory");

__global__ void __launch_bounds__(kTc3Threads, 1)
k_dyn_sigma_blend_tc2(dynrf_field_cfg_t cfg, dynrf_vm_t vden, dynrf_vm_t vbl, const float* __restrict__ wq /*kSbWFloats*/,
                      const float* __restrict__ b1cat /*[128]*/, const float* __restrict__ w2cat /*[128]*/, const float* __restrict__ b2cat /*[2]*/,
                      const float* __restrict__ xyz, const float* __restrict__ xyz_prime, const uint8_t* __restrict__ valid,
                      const float* __restrict__ tfeat, int P, int S, float* __restrict__ sigma, float* __restrict__ blend,
                      float* __restrict__ XT /*[224][ld] layer-1 inputs, optional*/, float* __restrict__ HT /*[128][ld] relu(layer 1)*/,
                      int ld) {
  extern __shared__ __align__(128) float smem[];
  float* abuf = smem;
  float* wbuf = abuf + 3 * kOpFloats;
  float* part0 = wbuf + 2 * kWSlotFloats;  // [2][4][128] partial dots
  float* cst = part0 + 2 * 4 * kTileM;     // b1cat[128] | w2cat[128]
  TapsSmem* tapsv = reinterpret_cast<TapsSmem*>(cst + 256);
  SbScratch* scrv = reinterpret_cast<SbScratch*>(tapsv + 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scrv + 2);
  uint64_t* bar_free = bars;       // [3]
  uint64_t* bar_w = bars + 3;      // [2]
  uint64_t* bar_full = bars + 5;   // [3]
  uint64_t* bar_stage = bars + 8;  // [2]: the tile of context s is complete in TMEM
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) umma::mbar_init(&bars[i], 1);
    for (int i = 0; i < 3; ++i) umma::mbar_init(&bar_full[i], kProdThreads / 32);
    for (int i = 0; i < 2; ++i) umma::mbar_init(&bar_stage[i], 1);
    umma::fence_mbar_init();
  }
  if (tid < 128) cst[tid] = __ldg(b1cat + tid);
  else if (tid < 256) cst[tid] = __ldg(w2cat + tid - 128);
  if (warp == 0) umma::tmem_alloc(tslot, 256);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Ring ring{bar_free, bar_w, abuf, wbuf, 0u, *tslot};
  const int ntiles = (P + kTileM - 1) / kTileM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total_chunks = (uint32_t)my_tiles * kSbChunks;

  if (warp == kProdThreads / 32) {
    // ===================== tensor-core warp =====================
    if (lane == 0 && total_chunks > 0) {
      ring_prefetch_w(ring, 0, wq + sb_woff(0), 2u * sb_n(0) * sb_rows(0) * 4u);
      for (uint32_t c = 0; c < total_chunks; ++c) {
        const int j = (int)(c % kSbChunks);
        const uint32_t ctx = (c / kSbChunks) & 1u, colb = ctx * 128u;
        ring.chunk = c;
        umma::mbar_wait(&bar_full[c % 3], (c / 3) & 1);
        const uint32_t col0 = (j >= 3 && ((j - 3) & 1)) ? 64u : 0u;
        if (j < 3) ring_issue<128>(ring, sb_rows(j), colb, j == 0);
        else ring_issue<64>(ring, sb_rows(j), colb + col0, false);
        if (j == kSbChunks - 1) umma::commit(&bar_stage[ctx]);
        if (c + 1 < total_chunks) {
          const int nj = (j + 1) % kSbChunks;
          ring_prefetch_w(ring, c + 1, wq + sb_woff(nj), 2u * sb_n(nj) * sb_rows(nj) * 4u);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== producer / epilogue warps =====================
    const int m = tid & (kTileM - 1), prt = tid >> 7;
    const AabbNorm nb = make_norm(cfg);
    const int G[3] = {vden.W[0], vden.H[0], vden.L[0]};
    auto arrive = [&]() {
      umma::fence_before_sync();
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[ring.chunk % 3]);
      ring.chunk++;
    };
#pragma unroll 1
    for (int k = 0; k <= my_tiles; ++k) {
      const bool hasY = k < my_tiles, hasX = k >= 1;
      const int sy = k & 1, sx = (k - 1) & 1;
      const int tileY = (int)blockIdx.x + k * (int)gridDim.x, tileX = tileY - (int)gridDim.x;
      if (hasY) {
        SbScratch& scr = scrv[sy];
        TapsSmem& taps = tapsv[sy];
        if (tid < kTileM) {
          const int i = tileY * kTileM + tid;
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
        prod_bar();
        // 9 taps per sample: (stride, axis)
        for (int item = tid; item < 9 * kTileM; item += kProdThreads) {
          const int mm = item & (kTileM - 1), t = item >> 7;
          const int si = t / 3, ax = t % 3;
          const Tap1 tp = make_tap(scr.xq[ax][mm], ax == 0 ? G[0] : (ax == 1 ? G[1] : G[2]), 1 << si);
          taps.t[t][mm] = make_int4(tp.i0, tp.i1, __float_as_int(tp.w0), __float_as_int(tp.w1));
        }
        prod_bar();
        const bool act = scr.on[m] != 0;
        const int tile = tileY;
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
          ring_wait_free(ring);
          float* buf = ring_abuf(ring);
          const int u0 = j * 5, nu = (j == 2) ? 4 : 5;
          for (int u = prt; u < nu; u += 4)
            produce_pos_unit(u0 + u, scr, tfeat, m, buf, XT ? XT + (size_t)(j * 32) * ld : nullptr, ld, tile * kTileM + m);
          arrive();
        }
        // feature chunks, one (density, blending) pair per stride: both tables are read at the same taps, so the lookups of the
        // pair are issued together (twice the loads in flight per thread: the first use of a gathered value was 27 % of the stall
        // samples) and the ring waits come after them.  The two 4-channel planes of a table go to threads 0,1 (density) / 2,3
        // (blending).
#pragma unroll 1
        for (int si = 0; si < 3; ++si) {
          const TapF tx = load_tap(taps, si, 0, m), ty = load_tap(taps, si, 1, m), tz = load_tap(taps, si, 2, m);
          float4 rd = make_float4(0.f, 0.f, 0.f, 0.f), rb = rd, rs = rd;
          const int pw = prt & 1;
          const bool sbl = (prt >> 1) != 0;   // this thread's small plane belongs to the blending table
          if (act) {
            rd = vm_group_f(vden.plane[0], vden.line[0], 16, plane_st(vden, 0), tx, ty, tz, prt * 4);
            rb = vm_group_f(vbl.plane[0], vbl.line[0], 16, plane_st(vden, 0), tx, ty, tz, prt * 4);
            const float* pp = pw == 0 ? (sbl ? vbl.plane[1] : vden.plane[1]) : (sbl ? vbl.plane[2] : vden.plane[2]);
            const float* ll = pw == 0 ? (sbl ? vbl.line[1] : vden.line[1]) : (sbl ? vbl.line[2] : vden.line[2]);
            // plane 1: (u,v,l) = (x,z,y), W = Gx; plane 2: (y,z,x), W = Gy
            rs = vm_group_f(pp, ll, 4, plane_st(vden, pw == 0 ? 1 : 2), pw == 0 ? tx : ty, tz, pw == 0 ? ty : tx, 0);
          }
#pragma unroll
          for (int isbl = 0; isbl < 2; ++isbl) {
            ring_wait_free(ring);
            float* buf = ring_abuf(ring);
            const int xrow = (isbl ? 152 : 80) + si * 24;   // XT rows: 80 position/time | 72 density | 72 blending
            const float4 r = isbl ? rb : rd;
            store_quad(buf, m, prt, r);
            if (XT) store_quad_g(XT, ld, xrow + prt * 4, tile * kTileM + m, r);
            if (sbl == (isbl != 0)) {
              store_quad(buf, m, 4 + pw, rs);
              if (XT) store_quad_g(XT, ld, xrow + 16 + pw * 4, tile * kTileM + m, rs);
            }
            arrive();
          }
        }
      }
      if (hasX) {
        const SbScratch& scr = scrv[sx];
        float* part = part0 + sx * (4 * kTileM);
        const int tile = tileX;
        umma::mbar_wait_backoff(&bar_stage[sx], (uint32_t)(((k - 1) >> 1) & 1));
        umma::fence_after_sync();
        {
          const int q = warp & 3, cb = warp >> 2;  // lane quadrant, 32-column block (0,1: density  2,3: blending)
          float r[32];
          umma::tmem_ld32(ring.tmem + ((uint32_t)(q * 32) << 16) + sx * 128 + cb * 32, r);
          float acc = 0.f;
          const bool on = scr.on[q * 32 + lane] != 0;
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int n = cb * 32 + jj;
            r[jj] = fmaxf(r[jj] + cst[n], 0.f);
            acc = fmaf(r[jj], cst[128 + n], acc);
          }
          part[cb * kTileM + q * 32 + lane] = acc;
          if (HT) {   // hidden activations for the backward (zero on samples outside the bounding box)
#pragma unroll
            for (int qq = 0; qq < 8; ++qq)
              store_quad_g(HT, ld, cb * 32 + qq * 4, tile * kTileM + q * 32 + lane,
                           on ? make_float4(r[qq * 4], r[qq * 4 + 1], r[qq * 4 + 2], r[qq * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
        umma::fence_before_sync();
        prod_bar();
        if (tid < 2 * kTileM) {
          const int which = tid >> 7, mm = tid & (kTileM - 1);
          const int i = tile * kTileM + mm;
          if (i < P) {
            const bool v = scr.on[mm] != 0;
            const float o = part[(2 * which) * kTileM + mm] + part[(2 * which + 1) * kTileM + mm] + __ldg(b2cat + which);
            if (which == 0) sigma[i] = v ? density_act(o, cfg.act, cfg.density_shift) : 0.f;
            else blend[i] = v ? sigmoidf_(o) : 0.f;
          }
        }
        prod_bar();
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(ring.tmem, 256);
}

// reference column of density/blending layer-1 for (chunk j, row k) of the K order above
__device__ __forceinline__ int sb_src_col(int j, int k) {
  if (j >= 3) return ((j - 3) >> 1) * 24 + k;  // features: stride-major, reference order
  const int r = j * 32 + k;                    // position/time row 0..79
  const int q = r >> 2, e = r & 3;
  int p;                                       // index in the reference's [xn3, sin30, cos30, t, tsin8, tcos8]
  if (q == 0) p = e < 3 ? e : 63;
  else if (q >= 16) {
    const int tq = q - 16;
    p = (tq < 2 ? 64 : 72) + (tq & 1) * 4 + e;
  } else {
    const int d = (q - 1) / 5, w = (q - 1) % 5;
    if (w == 4) p = (e < 2 ? 3 : 33) + d * 10 + 8 + (e & 1);
    else p = ((w & 1) ? 33 : 3) + d * 10 + (w >> 1) * 4 + e;
  }
  return 72 + p;
}

__global__ void k_pack_sb_umma(const float* __restrict__ w_den /*[64][152]*/, const float* __restrict__ w_bl, float* __restrict__ wq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kSbWFloats / 2) return;
  // locate chunk
  int j = 0, base = 0;
  for (; j < kSbChunks; ++j) {
    const int sz = sb_n(j) * sb_rows(j);
    if (i < base + sz) break;
    base += sz;
  }
  const int N = sb_n(j), rows = sb_rows(j);
  const int e = i - base;
  const int n = e / rows, k = e % rows;
  const int col = sb_src_col(j, k);
  float v;
  if (j < 3) v = n < 64 ? __ldg(w_den + (size_t)n * 152 + col) : __ldg(w_bl + (size_t)(n - 64) * 152 + col);
  else v = ((j - 3) & 1) ? __ldg(w_bl + (size_t)n * 152 + col) : __ldg(w_den + (size_t)n * 152 + col);
  float hi, lo;
  umma::split_tf32(v, hi, lo);
  const uint32_t off = umma::kmajor_off(n, k, N) >> 2;
  float* blob = wq + sb_woff(j);
  blob[off] = hi;
  blob[N * rows + off] = lo;
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

int dynrf_sb_umma_floats(void) { return kSbWFloats; }

int dynrf_pack_sb_umma(const float* w_den, const float* w_bl, float* wq, void* stream) {
  if (!w_den || !w_bl || !wq) return DYNRF_E_ARG;
  k_pack_sb_umma<<<(kSbWFloats / 2 + 255) / 256, 256, 0, S_(stream)>>>(w_den, w_bl, wq);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_sigma_blend_fwd_tc2(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                                  const float* wq, const float* b1cat, const float* w2cat, const float* b2cat, const float* xyz,
                                  const float* xyz_prime, const uint8_t* valid, const float* tfeat, int N, int S, float* sigma,
                                  float* blend, float* XT, float* HT, int ld, void* stream) {
  if (density && blending && (!dynrf_vm_layout_ok(density) || !dynrf_vm_same_layout(density, blending))) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !density || !blending || !wq || !b1cat || !w2cat || !b2cat || !xyz || !xyz_prime || !valid || !tfeat || !sigma || !blend ||
      N <= 0 || S <= 0)
    return DYNRF_E_ARG;
  if ((XT != nullptr) != (HT != nullptr)) return DYNRF_E_ARG;
  if (XT && ((ld & 3) || ld < ((N * S + kTileM - 1) / kTileM) * kTileM)) return DYNRF_E_ARG;
  for (int i = 0; i < 3; ++i) {
    const int c = i == 0 ? 16 : 4;
    if (density->C[i] != c || blending->C[i] != c) return DYNRF_E_UNSUPPORTED;
  }
  for (int i = 0; i < 3; ++i)
    if (density->H[i] != blending->H[i] || density->W[i] != blending->W[i] || density->L[i] != blending->L[i]) return DYNRF_E_UNSUPPORTED;
  if (((uintptr_t)wq & 15)) return DYNRF_E_ALIGN;
  cudaError_t e = cudaFuncSetAttribute(k_dyn_sigma_blend_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSb2Smem);
  if (e != cudaSuccess) return (int)e;
  const int P = N * S;
  const int grid = min((P + kTileM - 1) / kTileM, sm_count());
  k_dyn_sigma_blend_tc2<<<grid, kTc3Threads, kSb2Smem, S_(stream)>>>(*cfg, *density, *blending, wq, b1cat, w2cat, b2cat, xyz,
                                                                  xyz_prime, valid, tfeat, P, S, sigma, blend, XT, HT, ld);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_warp_fwd_tc(const dynrf_field_cfg_t* cfg, const float* xyz, int N, int S, const float* tfeat,
                      const dynrf_umma_layer_t* l3, const dynrf_umma_layer_t* l4, const dynrf_layer_t* l5, float* xyz_prime,
                      float* H1T, float* H2T, int ld, void* stream) {
  if (!cfg || !xyz || !tfeat || !xyz_prime || !l5 || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if ((H1T != nullptr) != (H2T != nullptr)) return DYNRF_E_ARG;
  if (H1T && ld < ((N * S + kTileM - 1) / kTileM) * kTileM) return DYNRF_E_ARG;
  DYNRF_TRY(check_ul(l3, 93, 64));
  DYNRF_TRY(check_ul(l4, 64, 64));
  if (l5->K != 64 || l5->N != 3) return DYNRF_E_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(k_warp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWarpTcSmem);
  if (e != cudaSuccess) return (int)e;
  const int P = N * S;
  const int grid = min((P + kTileM - 1) / kTileM, sm_count());
  k_warp_tc<<<grid, kMlpThreads, kWarpTcSmem, S_(stream)>>>(*cfg, xyz, P, S, tfeat, UL(l3), UL(l4), *l5, xyz_prime, H1T, H2T, ld);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_sigma_blend_fwd_tc(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                                 const dynrf_umma_layer_t* d1, const dynrf_layer_t* d2, const dynrf_umma_layer_t* b1,
                                 const dynrf_layer_t* b2, const float* xyz, const float* xyz_prime, const uint8_t* valid,
                                 const float* tfeat, int N, int S, float* sigma, float* blend, void* stream) {
  if (density && blending && (!dynrf_vm_layout_ok(density) || !dynrf_vm_same_layout(density, blending))) return DYNRF_E_UNSUPPORTED;
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
