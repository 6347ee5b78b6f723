// Tensor-core backward of the dynamic appearance head (MLPRender_Fea_late_view on the app-masked samples):
//
//   k_head_bwd_tc   : per 128-sample tile, warp-specialised (16 producer/epilogue warps + 1 tcgen05 warp):
//                     recompute L1 (107->128) and L2 (128->128) on tensor cores, finish the 131->3 layer and the
//                     sigmoid in the epilogue, back-propagate dH2 -> dH1 -> dfeat with two more tensor-core GEMMs
//                     (data gradients), and stream every operand a weight gradient needs (layer inputs and
//                     output gradients, feature-major) to global memory.
//   dynrf_wgrad_gemm (wgrad_gemm.cu) then forms dW = X^T dY over all samples on tensor cores, and
//   k_basis_bwd     : per 64-sample tile (SIMT): basis_mat gradient, dfeat -> d(plane*line) -> plane / line
//                     reductions and d/d(xyz_prime).
//
// TMEM columns: acc1 [0,128) L1 pre-activation | acc2 [128,256) L2 pre-activation |
//               acc3 [256,384) dH1 before the ReLU mask | acc4 [384,416) dfeat.
#include "dynrf_tc.cuh"

namespace dynrf {

constexpr int kHbChunks = 16;
__host__ __device__ constexpr int hb_rows(int j) { return j == 2 ? 16 : 32; }
__host__ __device__ constexpr int hb_n(int j) { return j < 12 ? 128 : 32; }
__host__ __device__ constexpr int hb_woff(int j) {
  int o = 0;
  for (int i = 0; i < j; ++i) o += 2 * hb_n(i) * hb_rows(i);
  return o;
}
constexpr int kHbWFloats = hb_woff(kHbChunks);

// constants in shared memory: b1[128] | b2[128] | W3t[128][3] | W3v[3][3] | b3[3]
constexpr int kHbCst = 128 + 128 + 384 + 9 + 3;

struct HeadBwdArgs {
  const float* wq;      // kHbWFloats: packed split weights for the 16 chunks
  const float* cst;     // kHbCst
  const float* rays;    // [N][6]
  const float* xyz;     // [N*S][3]
  const float* tfeat;   // [N][48]
  const float* featT;   // [32][ld] basis output saved by the forward (rows 27..31 zero)
  const float* grgb;    // [N*S][3]
  const int32_t* active_idx;
  const int32_t* active_count;
  int S, ld;
  // feature-major outputs, row stride ld
  float* X1T;   // [113]: 80 position/time rows (this kernel's K order) | 32 feat rows | 1 row of ones
  float* H1T;   // [128]
  float* H2T;   // [128]
  float* DH2T;  // [128]
  float* DH1T;  // [128]
  float* DOT;   // [3]
  float* DFT;   // [32]
  float* gw3v;  // [3][3] += d W3[:, 128+k]   (row c, col k)
  float* gb3;   // [3]    += d b3
};

constexpr size_t kHbSmem = (size_t)(3 * kOpFloats + 2 * kWSlotFloats + kHbCst + 4 * 3 * kTileM + 3 * kTileM) * sizeof(float) +
                           sizeof(TileScratch<kTileM>) + 128;

__global__ void __launch_bounds__(kTc3Threads, 1) k_head_bwd_tc(dynrf_field_cfg_t cfg, HeadBwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  float* abuf = smem;
  float* wbuf = abuf + 3 * kOpFloats;
  float* cst = wbuf + 2 * kWSlotFloats;
  float* part = cst + kHbCst;       // [4][3][128] partial output dots
  float* go = part + 4 * 3 * kTileM;  // [3][128] d loss / d (pre-sigmoid output)
  TileScratch<kTileM>& scr = *reinterpret_cast<TileScratch<kTileM>*>(go + 3 * kTileM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(&scr) + sizeof(TileScratch<kTileM>));
  uint64_t* bar_free = bars;
  uint64_t* bar_w = bars + 3;
  uint64_t* bar_full = bars + 5;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) umma::mbar_init(&bars[i], 1);
    for (int i = 0; i < 3; ++i) umma::mbar_init(&bar_full[i], kProdThreads / 32);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < kHbCst; i += kTc3Threads) cst[i] = __ldg(a.cst + i);
  if (warp == 0) umma::tmem_alloc(tslot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Ring ring{bar_free, bar_w, abuf, wbuf, 0u, *tslot};
  const int count = *a.active_count;
  const int ntiles = (count + kTileM - 1) / kTileM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total_chunks = (uint32_t)my_tiles * kHbChunks;
  const float* b1 = cst;
  const float* b2 = cst + 128;
  const float* W3t = cst + 256;   // [128][3]
  const float* W3v = cst + 640;   // [k][c]
  const float* b3 = cst + 649;

  if (warp == kProdThreads / 32) {
    // ===================== tensor-core warp =====================
    if (lane == 0 && total_chunks > 0) {
      ring_prefetch_w(ring, 0, a.wq + hb_woff(0), 2u * hb_n(0) * hb_rows(0) * 4u);
      for (uint32_t c = 0; c < total_chunks; ++c) {
        const int j = (int)(c % kHbChunks);
        ring.chunk = c;
        umma::mbar_wait(&bar_full[c % 3], (c / 3) & 1);
        const uint32_t col0 = j < 4 ? 0u : (j < 8 ? 128u : (j < 12 ? 256u : 384u));
        const bool first = (j == 0) || (j == 4) || (j == 8) || (j == 12);
        if (j < 12) ring_issue<128>(ring, hb_rows(j), col0, first);
        else ring_issue<32>(ring, hb_rows(j), col0, first);
        if (c + 1 < total_chunks) {
          const int nj = (j + 1) % kHbChunks;
          ring_prefetch_w(ring, c + 1, a.wq + hb_woff(nj), 2u * hb_n(nj) * hb_rows(nj) * 4u);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== producer / epilogue warps =====================
    const int m = tid & (kTileM - 1), prt = tid >> 7;
    const int q = warp & 3, cb = warp >> 2;  // TMEM lane quadrant, 32-column block
    const int em = q * 32 + lane;            // sample row this thread owns in the epilogues
    const AabbNorm nb = make_norm(cfg);
    const int ld = a.ld;
    float acc_vd[9], acc_b3[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) acc_vd[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) acc_b3[i] = 0.f;

    auto wait_chunk = [&](uint32_t c) {
      umma::mbar_wait(&bar_free[c % 3], (c / 3) & 1);
      umma::fence_after_sync();
    };
    auto arrive = [&]() {
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[ring.chunk % 3]);
      ring.chunk++;
    };

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int jg0 = tile * kTileM;
      if (tid < kTileM) {
        const int jg = jg0 + tid;
        const bool in = jg < count;
        const int i = in ? __ldg(a.active_idx + jg) : 0;
        scr.idx[tid] = i;
        scr.ray[tid] = in ? i / a.S : -1;
        scr.on[tid] = in ? 1 : 0;
        float vd[3] = {0.f, 0.f, 0.f};
        if (in) ray_viewdir(a.rays, i / a.S, vd);
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          const float x = in ? __ldg(a.xyz + (size_t)i * 3 + ax) : 0.f;
          scr.xn[ax][tid] = in ? norm_coord(x, nb.mn[ax], nb.inv[ax]) : 0.f;
          scr.xu[ax][tid] = vd[ax];
        }
      }
      prod_bar();
      const uint32_t base = ring.chunk;
      // ---- L1 operand: 3 position/time chunks (also streamed to X1T rows 0..79) + the feat chunk (X1T rows 80..111 <- featT)
#pragma unroll 1
      for (int j = 0; j < 3; ++j) {
        ring_wait_free(ring);
        float* buf = ring_abuf(ring);
        const int u0 = j * 5, nu = (j == 2) ? 4 : 5;
        for (int u = prt; u < nu; u += 4) produce_pos_unit(u0 + u, scr, a.tfeat, m, buf, a.X1T + (size_t)(j * 32) * ld, ld, jg0 + m);
        arrive();
      }
      {
        ring_wait_free(ring);
        float* buf = ring_abuf(ring);
        const bool on = scr.on[m] != 0;
        for (int qq = prt; qq < 8; qq += 4) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (on) {
            const float* f = a.featT + (size_t)(qq * 4) * ld + jg0 + m;
            v = make_float4(__ldg(f), __ldg(f + ld), __ldg(f + 2 * (size_t)ld), __ldg(f + 3 * (size_t)ld));
          }
          store_quad(buf, m, qq, v);
          store_quad_g(a.X1T, ld, 80 + qq * 4, jg0 + m, v);
        }
        if (prt == 0) a.X1T[(size_t)112 * ld + jg0 + m] = on ? 1.f : 0.f;  // ones row: its weight-gradient column is d b1
        arrive();
      }
      // ---- epilogue 1: h1 = relu(acc1 + b1) -> L2 operand chunk cb, H1T
      wait_chunk(base + 3);
      {
        float r[32];
        umma::tmem_ld32(ring.tmem + ((uint32_t)(q * 32) << 16) + cb * 32, r);
        const bool on = scr.on[em] != 0;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) r[jj] = on ? fmaxf(r[jj] + b1[cb * 32 + jj], 0.f) : 0.f;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          // every warp waits for the slot's previous use before arriving: an early arrival meant for chunk c would
          // otherwise be counted in the still-open phase of chunk c-3 on the same barrier
          ring_wait_free(ring);
          if (c4 == cb) {
            float* buf = ring_abuf(ring);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const float4 v = make_float4(r[qq * 4], r[qq * 4 + 1], r[qq * 4 + 2], r[qq * 4 + 3]);
              store_quad(buf, em, qq, v);
              store_quad_g(a.H1T, ld, cb * 32 + qq * 4, jg0 + em, v);
            }
          }
          arrive();
        }
      }
      // ---- epilogue 2: h2 = relu(acc2 + b2); output layer; d(out); dH2 -> dgrad-2 operand chunk cb
      wait_chunk(base + 7);
      {
        float h2[32];
        umma::tmem_ld32(ring.tmem + ((uint32_t)(q * 32) << 16) + 128 + cb * 32, h2);
        const bool on = scr.on[em] != 0;
        float po[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int n = cb * 32 + jj;
          h2[jj] = on ? fmaxf(h2[jj] + b2[n], 0.f) : 0.f;
          po[0] = fmaf(h2[jj], W3t[n * 3 + 0], po[0]);
          po[1] = fmaf(h2[jj], W3t[n * 3 + 1], po[1]);
          po[2] = fmaf(h2[jj], W3t[n * 3 + 2], po[2]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) part[(cb * 3 + c) * kTileM + em] = po[c];
#pragma unroll
        for (int qq = 0; qq < 8; ++qq)
          store_quad_g(a.H2T, ld, cb * 32 + qq * 4, jg0 + em, make_float4(h2[qq * 4], h2[qq * 4 + 1], h2[qq * 4 + 2], h2[qq * 4 + 3]));
        prod_bar();
        if (tid < kTileM) {
          const bool onm = scr.on[tid] != 0;
          float g3[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float o = b3[c] + part[(0 * 3 + c) * kTileM + tid] + part[(1 * 3 + c) * kTileM + tid] + part[(2 * 3 + c) * kTileM + tid] +
                      part[(3 * 3 + c) * kTileM + tid];
#pragma unroll
            for (int k = 0; k < 3; ++k) o = fmaf(scr.xu[k][tid], W3v[k * 3 + c], o);
            const float rgbv = sigmoidf_(o);
            g3[c] = onm ? __ldg(a.grgb + (size_t)scr.idx[tid] * 3 + c) * rgbv * (1.0f - rgbv) : 0.f;
            go[c * kTileM + tid] = g3[c];
            a.DOT[(size_t)c * ld + jg0 + tid] = g3[c];
            acc_b3[c] += g3[c];
          }
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc_vd[c * 3 + k] = fmaf(scr.xu[k][tid], g3[c], acc_vd[c * 3 + k]);
        }
        prod_bar();
        const float g0 = go[em], g1 = go[kTileM + em], g2 = go[2 * kTileM + em];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int n = cb * 32 + jj;
          const float d = fmaf(W3t[n * 3 + 0], g0, fmaf(W3t[n * 3 + 1], g1, W3t[n * 3 + 2] * g2));
          h2[jj] = h2[jj] > 0.f ? d : 0.f;
        }
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          // every warp waits for the slot's previous use before arriving: an early arrival meant for chunk c would
          // otherwise be counted in the still-open phase of chunk c-3 on the same barrier
          ring_wait_free(ring);
          if (c4 == cb) {
            float* buf = ring_abuf(ring);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const float4 v = make_float4(h2[qq * 4], h2[qq * 4 + 1], h2[qq * 4 + 2], h2[qq * 4 + 3]);
              store_quad(buf, em, qq, v);
              store_quad_g(a.DH2T, ld, cb * 32 + qq * 4, jg0 + em, v);
            }
          }
          arrive();
        }
      }
      // ---- epilogue 3: dH1 = acc3 * [h1 > 0] -> dgrad-1 operand chunk cb, DH1T
      wait_chunk(base + 11);
      {
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          // every warp waits for the slot's previous use before arriving: an early arrival meant for chunk c would
          // otherwise be counted in the still-open phase of chunk c-3 on the same barrier
          ring_wait_free(ring);
          if (c4 == cb) {
            float* buf = ring_abuf(ring);
            const bool on = scr.on[em] != 0;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              float r3[16], r1[16];
              umma::tmem_ld16(ring.tmem + ((uint32_t)(q * 32) << 16) + 256 + cb * 32 + hf * 16, r3);
              umma::tmem_ld16(ring.tmem + ((uint32_t)(q * 32) << 16) + cb * 32 + hf * 16, r1);
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) r3[jj] = (on && (r1[jj] + b1[cb * 32 + hf * 16 + jj]) > 0.f) ? r3[jj] : 0.f;
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                const float4 v = make_float4(r3[qq * 4], r3[qq * 4 + 1], r3[qq * 4 + 2], r3[qq * 4 + 3]);
                store_quad(buf, em, hf * 4 + qq, v);
                store_quad_g(a.DH1T, ld, cb * 32 + hf * 16 + qq * 4, jg0 + em, v);
              }
            }
          }
          arrive();
        }
      }
      // ---- epilogue 4: dfeat
      wait_chunk(base + 15);
      if (cb == 0) {
        float r4[32];
        umma::tmem_ld32(ring.tmem + ((uint32_t)(q * 32) << 16) + 384, r4);
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) {
          float4 v = make_float4(r4[qq * 4], r4[qq * 4 + 1], r4[qq * 4 + 2], r4[qq * 4 + 3]);
          if (qq == 6) v.w = 0.f;            // only 27 feature rows carry a gradient
          if (qq == 7) v = make_float4(0.f, 0.f, 0.f, 0.f);
          store_quad_g(a.DFT, ld, qq * 4, jg0 + em, v);
        }
      }
      umma::fence_before_sync();
      prod_bar();
    }
    if (tid < kTileM) {
#pragma unroll
      for (int i = 0; i < 9; ++i) atomicAdd(a.gw3v + i, warp_sum(acc_vd[i]) * (lane == 0 ? 1.f : 0.f));
#pragma unroll
      for (int i = 0; i < 3; ++i) atomicAdd(a.gb3 + i, warp_sum(acc_b3[i]) * (lane == 0 ? 1.f : 0.f));
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(ring.tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// weight packing for k_head_bwd_tc: chunk j holds B[row r < N_j][k < rows_j] (split hi | lo, K-major):
//   j 0..2   L1, position/time inputs : B[n][k] = W1[n][27 + p(j,k)]         (p: this kernel's K order -> reference order)
//   j 3      L1, feat inputs          : B[n][k] = W1[n][k] (k < 27)
//   j 4..7   L2                       : B[n][k] = W2[n][32 (j-4) + k]
//   j 8..11  dH1 = dH2 W2             : B[kin][nn] = W2[32 (j-8) + nn][kin]
//   j 12..15 dfeat = dH1 W1[:, feat]  : B[f][nn] = W1[32 (j-12) + nn][f] (f < 27)

__global__ void k_pack_head_bwd(const float* __restrict__ w1 /*[128][107]*/, const float* __restrict__ w2 /*[128][128]*/,
                                float* __restrict__ wq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kHbWFloats / 2) return;
  int j = 0, base = 0;
  for (; j < kHbChunks; ++j) {
    const int sz = hb_n(j) * hb_rows(j);
    if (i < base + sz) break;
    base += sz;
  }
  const int N = hb_n(j), rows = hb_rows(j);
  const int e = i - base;
  const int r = e / rows, k = e % rows;
  float v = 0.f;
  if (j < 3) v = __ldg(w1 + (size_t)r * 107 + 27 + pos_src_index(j * 32 + k));
  else if (j == 3) v = k < 27 ? __ldg(w1 + (size_t)r * 107 + k) : 0.f;
  else if (j < 8) v = __ldg(w2 + (size_t)r * 128 + 32 * (j - 4) + k);
  else if (j < 12) v = __ldg(w2 + (size_t)(32 * (j - 8) + k) * 128 + r);
  else v = r < 27 ? __ldg(w1 + (size_t)(32 * (j - 12) + k) * 107 + r) : 0.f;
  float hi, lo;
  umma::split_tf32(v, hi, lo);
  const uint32_t off = umma::kmajor_off(r, k, N) >> 2;
  float* blob = wq + hb_woff(j);
  blob[off] = hi;
  blob[N * rows + off] = lo;
}

// column permutation of the position/time block (X1T rows 0..79 -> nn.Linear input column 27 + p)
__global__ void k_head_bwd_colmap(int32_t* __restrict__ map /*[113]*/) {
  const int r = threadIdx.x;
  if (r < 80) map[r] = 27 + pos_src_index(r);
  else if (r < 112) map[r] = (r - 80) < 27 ? (r - 80) : -1;
  else if (r == 112) map[r] = -2;  // bias
}

// ------------------------------------------------------------------------------------------------
// basis_mat + lookup backward on the active list (SIMT, 64-sample tiles): DFT -> d basis, plane / line grads, d coords

constexpr int kBT2 = 64;
constexpr int kBL2 = kBT2 + 4;
using BScratch2 = TileScratch<kBT2>;
constexpr size_t kBasisBwdSmem = (size_t)((72 + 32 + 80) * kBL2 + 2 * kChunkK * 80) * sizeof(float) + sizeof(BScratch2);

template <bool DYN>
__global__ void __launch_bounds__(kMlpThreads, 2)
k_basis_bwd(dynrf_field_cfg_t cfg, dynrf_vm_t vm, dynrf_vm_grad_t gvm, const float* __restrict__ wbB /*[nstr][32][80]*/,
            const float* __restrict__ coords /*xyz_prime (DYN) | xyz*/, const float* __restrict__ DFT, int ld,
            const int32_t* __restrict__ active_idx, const int32_t* __restrict__ active_count, float* __restrict__ gbasis_wt /*[K][32]*/,
            float* __restrict__ gcoord) {
  constexpr int NSTR = DYN ? 3 : 1;
  extern __shared__ __align__(128) float smem[];
  float* Ac = smem;
  float* dF = Ac + 72 * kBL2;
  float* dAc = dF + 32 * kBL2;
  float* wbuf = dAc + 80 * kBL2;
  BScratch2& scr = *reinterpret_cast<BScratch2*>(wbuf + 2 * kChunkK * 80);
  const int tid = threadIdx.x;
  const AabbNorm nb = make_norm(cfg);
  const int count = *active_count;
  const int ntiles = (count + kBT2 - 1) / kBT2;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    if (tid < kBT2) {
      const int j = tile * kBT2 + tid;
      const bool in = j < count;
      const int i = in ? __ldg(active_idx + j) : 0;
      scr.idx[tid] = i;
      scr.on[tid] = in ? 1 : 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float x = in ? __ldg(coords + (size_t)i * 3 + a) : 0.f;
        scr.xq[a][tid] = norm_coord(x, nb.mn[a], nb.inv[a]);
        scr.gx[a][tid] = 0.f;
      }
    }
    for (int item = tid; item < 32 * kBT2; item += kMlpThreads) {
      const int mm = item % kBT2, r = item / kBT2;
      const int j = tile * kBT2 + mm;
      dF[(size_t)r * kBL2 + mm] = (j < count && r < 27) ? __ldg(DFT + (size_t)r * ld + j) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int si = 0; si < NSTR; ++si) {
      gather_vm_rows<kBT2, kBL2>(vm, scr.xq, 1 << si, 0, Ac, tid, scr.on);
      __syncthreads();
      {
        float accB[5][2];
        wgrad_zero(accB);
        wgrad_acc<5, 2, kBT2, kBL2>(accB, Ac, 72, dF, tid);
        wgrad_flush(accB, gbasis_wt + (size_t)si * 72 * 32, 72, 32, tid);
      }
      {
        Acc<80, kBT2> acc;
        acc_init<80, kBT2>(acc, nullptr, tid >> 4);
        acc_gemm<80, kBT2, kBL2>(acc, dF, 32, wbB + (size_t)si * 32 * 80, wbuf, tid);
        acc_store<80, kBT2, kBL2, false>(acc, dAc, 72, tid);
      }
      scatter_vm_rows<kBT2, kBL2>(vm, gvm, scr.xq, 1 << si, 0, dAc, tid, scr.on, scr.gx);
      __syncthreads();
    }
    if (gcoord && tid < kBT2 && scr.on[tid]) {
      const int i = scr.idx[tid];
#pragma unroll
      for (int a = 0; a < 3; ++a) gcoord[(size_t)i * 3 + a] += scr.gx[a][tid] * nb.inv[a];
    }
    __syncthreads();
  }
}

}  // namespace dynrf

using namespace dynrf;

static inline cudaStream_t S_(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int dynrf_head_bwd_wfloats(void) { return kHbWFloats; }
int dynrf_head_bwd_cst_floats(void) { return kHbCst; }

int dynrf_pack_head_bwd(const float* w1, const float* w2, float* wq, int32_t* colmap, void* stream) {
  if (!w1 || !w2 || !wq || !colmap) return DYNRF_E_ARG;
  k_pack_head_bwd<<<(kHbWFloats / 2 + 255) / 256, 256, 0, S_(stream)>>>(w1, w2, wq);
  k_head_bwd_colmap<<<1, 128, 0, S_(stream)>>>(colmap);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_head_bwd_tc(const dynrf_field_cfg_t* cfg, const float* wq, const float* cst, const float* rays, const float* xyz,
                          const float* tfeat, const float* featT, const float* grgb, const int32_t* active_idx,
                          const int32_t* active_count, int N, int S, int ld, float* X1T, float* H1T, float* H2T, float* DH2T,
                          float* DH1T, float* DOT, float* DFT, float* gw3v, float* gb3, void* stream) {
  if (!cfg || !wq || !cst || !rays || !xyz || !tfeat || !featT || !grgb || !active_idx || !active_count || !X1T || !H1T || !H2T ||
      !DH2T || !DH1T || !DOT || !DFT || !gw3v || !gb3 || N <= 0 || S <= 0 || ld < N * S || (ld & 3))
    return DYNRF_E_ARG;
  if ((uintptr_t)wq & 15) return DYNRF_E_ALIGN;
  cudaError_t e = cudaFuncSetAttribute(k_head_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHbSmem);
  if (e != cudaSuccess) return (int)e;
  HeadBwdArgs a{wq, cst, rays, xyz, tfeat, featT, grgb, active_idx, active_count, S, ld, X1T, H1T, H2T, DH2T, DH1T, DOT, DFT, gw3v, gb3};
  const int grid = min((N * S + kTileM - 1) / kTileM, sm_count());
  k_head_bwd_tc<<<grid, kTc3Threads, kHbSmem, S_(stream)>>>(*cfg, a);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_basis_bwd(const dynrf_field_cfg_t* cfg, int dynamic, const dynrf_vm_t* app, const dynrf_vm_grad_t* gapp, const float* wbB,
                    const float* coords, const float* DFT, int ld, int N, int S, const int32_t* active_idx,
                    const int32_t* active_count, float* gbasis_wt, float* gcoord, void* stream) {
  if (!cfg || !app || !gapp || !wbB || !coords || !DFT || !active_idx || !active_count || !gbasis_wt || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if (app->C[0] != 48 || app->C[1] != 12 || app->C[2] != 12) return DYNRF_E_UNSUPPORTED;
  const int grid = min((N * S + kBT2 - 1) / kBT2, 2 * sm_count());
  cudaError_t e;
  if (dynamic) {
    e = cudaFuncSetAttribute(k_basis_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBasisBwdSmem);
    if (e != cudaSuccess) return (int)e;
    k_basis_bwd<true><<<grid, kMlpThreads, kBasisBwdSmem, S_(stream)>>>(*cfg, *app, *gapp, wbB, coords, DFT, ld, active_idx, active_count,
                                                                        gbasis_wt, gcoord);
  } else {
    e = cudaFuncSetAttribute(k_basis_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBasisBwdSmem);
    if (e != cudaSuccess) return (int)e;
    k_basis_bwd<false><<<grid, kMlpThreads, kBasisBwdSmem, S_(stream)>>>(*cfg, *app, *gapp, wbB, coords, DFT, ld, active_idx,
                                                                         active_count, gbasis_wt, gcoord);
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
