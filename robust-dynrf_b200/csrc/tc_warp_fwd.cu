// Tensor-core forward of the coordinate warp (a15, models/tensoRF.py:521-541 warp_coordinate): 93 -> 64 -> 64 -> 3 on the
// samples of a ray batch, xyz_prime = unnormalize(normalize(xyz)) + MLP(...).
//
// Warp-specialised like the other tile kernels (16 producer / epilogue warps + 1 tcgen05 warp, operand ring + TMA weight slots of
// dynrf_tc.cuh), two tiles in flight: iteration k reads back layer 1 of tile X = k - 1 and submits its layer-2 chunks, produces
// the three layer-1 operand chunks of tile Y = k (30 sincosf pairs per sample), then finishes X.  The first-generation kernel
// (k_warp_tc in tc_kernels.cu: 256 threads, fp32 staging tile, CTA-wide barriers between every step, four epilogue warps) ran
// with 8 warps per SM at 24 % issue utilisation.
//
// K order of layer 3 (this kernel's own; the pack kernel permutes the reference's columns):
//   chunks 0, 1 : the position rows in produce_pos_unit order, 16 quads: [xq0 xq1 xq2 -] then per dimension
//                 [sin f0-3][cos f0-3][sin f4-7][cos f4-7][sin f8, sin f9, cos f8, cos f9]      (row 3 has zero weights)
//   chunk 2     : t_out (30 values, 2 zero rows)
// TMEM columns of context s (base 128 s): [0,64) layer-3 pre-activation | [64,128) layer-4 pre-activation.
#include "dynrf_tc.cuh"

namespace dynrf {

constexpr int kWfChunks = 5;                      // 3 layer-3 chunks, 2 layer-4 chunks, 32 rows x N = 64 each
constexpr int kWfChunkFloats = 2 * 64 * kKC;      // hi + lo
constexpr int kWfWFloats = kWfChunks * kWfChunkFloats;
constexpr int kWfCst = 64 + 64 + 192 + 3;         // b3 | b4 | W5 [64][3] | b5

struct WfScratch {
  float xn[3][kTileM];  // normalize(unnormalize(normalize(xyz))): the coordinate the encodings see (produce_pos_unit reads .xn)
  float xu[3][kTileM];  // unnormalize(normalize(xyz))
  int ray[kTileM];
};

struct WarpFwdArgs {
  const float* wq;    // kWfWFloats
  const float* cst;   // kWfCst
  const float* xyz;   // [P][3]
  const float* tfeat; // [N][48]
  int P, S, ld;
  float* xyz_prime;   // [P][3]
  float* H1T;         // [64][ld] relu(layer 3), optional (kept for k_warp_bwd)
  float* H2T;         // [64][ld] relu(layer 4), optional
};

// chunk sequence of a CTA with T tiles: iteration k = 0..T: [k >= 1: the two layer-4 chunks of tile k - 1] [k < T: the three layer-3
// chunks of tile k]
struct WarpSeq {
  int k, i, T;   // i: 0,1 = layer-4 chunks of X, 2..4 = layer-3 chunks of Y
  __device__ __forceinline__ void norm() {
    for (;;) {
      if (k > T) return;
      if (i < 2 && k < 1) { i = 2; continue; }
      if (i >= 2 && i < 5 && k >= T) { i = 5; }
      if (i >= 5) { i = 0; ++k; continue; }
      return;
    }
  }
  __device__ __forceinline__ bool valid() const { return k <= T; }
  __device__ __forceinline__ int j() const { return i < 2 ? 3 + i : i - 2; }
  __device__ __forceinline__ int s() const { return (i < 2 ? k - 1 : k) & 1; }
};

constexpr size_t kWfSmem = (size_t)(3 * kOpFloats + 2 * kWfChunkFloats + kWfCst + 1 + 2 * 4 * 3 * kTileM) * sizeof(float) + 2 * sizeof(WfScratch) + 256;

__global__ void __launch_bounds__(kTc3Threads, 1) k_warp_fwd_tc2(dynrf_field_cfg_t cfg, WarpFwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  float* abuf = smem;
  float* wbuf = abuf + 3 * kOpFloats;
  float* cst = wbuf + 2 * kWfChunkFloats;
  float* part0 = cst + kWfCst + 1;  // [2][4][3][128]
  WfScratch* scrv = reinterpret_cast<WfScratch*>(part0 + 2 * 4 * 3 * kTileM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scrv + 2);
  uint64_t* bar_free = bars;                  // [kNFull] chunk c consumed (see dynrf_tc.cuh)
  uint64_t* bar_w = bars + kNFull;            // [2]
  uint64_t* bar_stage = bars + kNFull + 2;    // [2 contexts][2: layer 3 complete, layer 4 complete]
  uint64_t* bar_full = bars + kNFull + 6;     // [kNFull] chunk c written
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * kNFull + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kNFull + 2; ++i) umma::mbar_init(&bars[i], 1);
    for (int i = 0; i < 4; ++i) umma::mbar_init(&bar_stage[i], 1);
    for (int i = 0; i < kNFull; ++i) umma::mbar_init(&bar_full[i], kProdThreads / 32);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < kWfCst; i += kTc3Threads) cst[i] = __ldg(a.cst + i);
  if (warp == 0) umma::tmem_alloc(tslot, 256);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Ring ring{bar_free, bar_w, abuf, wbuf, 0u, *tslot, (uint32_t)kWfChunkFloats, 3u, 2u, (uint32_t)kNFull};
  const int P = a.P;
  const int ntiles = (P + kTileM - 1) / kTileM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const float* b3 = cst;
  const float* b4 = cst + 64;
  const float* W5 = cst + 128;   // [64][3]
  const float* b5 = cst + 320;
  constexpr uint32_t kWBytes = kWfChunkFloats * 4u;

  if (warp == kProdThreads / 32) {
    if (lane == 0 && my_tiles > 0) {
      WarpSeq sq{0, 0, my_tiles};
      sq.norm();
      ring_prefetch_w(ring, 0, a.wq + (size_t)sq.j() * kWfChunkFloats, kWBytes);
      for (uint32_t c = 0; sq.valid(); ++c) {
        const int j = sq.j();
        const uint32_t col = (uint32_t)(sq.s() * 128);
        ring.chunk = c;
        umma::mbar_wait(&bar_full[c % kNFull], (c / kNFull) & 1);
        if (j < 3) ring_issue<64>(ring, 32, col, j == 0);
        else ring_issue<64>(ring, 32, col + 64u, j == 3);
        if (j == 2) umma::commit(&bar_stage[sq.s() * 2 + 0]);
        else if (j == 4) umma::commit(&bar_stage[sq.s() * 2 + 1]);
        ++sq.i;
        sq.norm();
        if (sq.valid()) ring_prefetch_w(ring, c + 1, a.wq + (size_t)sq.j() * kWfChunkFloats, kWBytes);
      }
    }
    __syncwarp();
  } else {
    const int m = tid & (kTileM - 1), prt = tid >> 7;
    const int q = warp & 3, cb = warp >> 2;   // TMEM lane quadrant; 16-column block of the 64-wide accumulators
    const int em = q * 32 + lane;
    const AabbNorm nb = make_norm(cfg);
    const int ld = a.ld;
    auto arrive = [&]() {
      umma::fence_before_sync();
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[ring.chunk % kNFull]);
      ring.chunk++;
    };
#pragma unroll 1
    for (int k = 0; k <= my_tiles; ++k) {
      const bool hasY = k < my_tiles, hasX = k >= 1;
      const int sy = k & 1, sx = (k - 1) & 1;
      const int tileX = (int)blockIdx.x + (k - 1) * (int)gridDim.x;
      const uint32_t tcolX = ring.tmem + (uint32_t)(sx * 128) + ((uint32_t)(q * 32) << 16);
      const uint32_t parx = (uint32_t)(((k - 1) >> 1) & 1);
      if (hasX) {
        // ---- E1(X): h1 = relu(acc1 + b3) -> the two layer-4 operand chunks (chunk c2 <- the warps with cb >> 1 == c2)
        umma::mbar_wait_backoff(&bar_stage[sx * 2 + 0], parx);
        umma::fence_after_sync();
        float h[16];
        umma::tmem_ld16(tcolX + cb * 16, h);
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          h[jj] = fmaxf(h[jj] + b3[cb * 16 + jj], 0.f);
          if (a.H1T) DYNRF_ST_STREAM(a.H1T + (size_t)(cb * 16 + jj) * ld + tileX * kTileM + em, h[jj]);
        }
#pragma unroll 1
        for (int c2 = 0; c2 < 2; ++c2) {
          if (c2 == (cb >> 1)) {
            ring_wait_free(ring);
            float* buf = ring_abuf(ring);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq)
              store_quad(buf, em, (cb & 1) * 4 + qq, make_float4(h[qq * 4], h[qq * 4 + 1], h[qq * 4 + 2], h[qq * 4 + 3]));
          }
          arrive();
        }
      }
      if (hasY) {
        WfScratch& scr = scrv[sy];
        const int tileY = (int)blockIdx.x + k * (int)gridDim.x;
        // ---- S0(Y): per-sample scratch
        if (tid < kTileM) {
          const int i = tileY * kTileM + tid;
          const bool in = i < P;
          scr.ray[tid] = in ? i / a.S : -1;
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            const float x = in ? __ldg(a.xyz + (size_t)i * 3 + ax) : 0.f;
            const float xn = norm_coord(x, nb.mn[ax], nb.inv[ax]);
            const float xu = unnorm_coord(xn, nb.mn[ax], nb.inv[ax]);
            scr.xu[ax][tid] = xu;
            scr.xn[ax][tid] = norm_coord(xu, nb.mn[ax], nb.inv[ax]);
          }
        }
        prod_bar();
        // ---- S1(Y): layer-3 operand chunks
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          ring_wait_free(ring);
          float* buf = ring_abuf(ring);
          for (int u = prt; u < 5; u += 4) produce_pos_unit(j * 5 + u, scr, a.tfeat, m, buf);
          arrive();
        }
        {
          ring_wait_free(ring);
          float* buf = ring_abuf(ring);
          const int ry = scr.ray[m];
          for (int qq = prt; qq < 8; qq += 4) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = (ry >= 0 && qq * 4 + e < 30) ? __ldg(a.tfeat + (size_t)ry * kTFeat + 17 + qq * 4 + e) : 0.f;
            store_quad(buf, m, qq, make_float4(v[0], v[1], v[2], v[3]));
          }
          arrive();
        }
      }
      if (hasX) {
        // ---- E2(X): h2 = relu(acc2 + b4); delta = W5 h2 + b5; xyz_prime = xu + delta
        const WfScratch& scr = scrv[sx];
        float* part = part0 + sx * (4 * 3 * kTileM);
        umma::mbar_wait_backoff(&bar_stage[sx * 2 + 1], parx);
        umma::fence_after_sync();
        {
          float h[16];
          umma::tmem_ld16(tcolX + 64 + cb * 16, h);
          float d3[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const int n = cb * 16 + jj;
            const float hv = fmaxf(h[jj] + b4[n], 0.f);
            if (a.H2T) DYNRF_ST_STREAM(a.H2T + (size_t)n * ld + tileX * kTileM + em, hv);
#pragma unroll
            for (int o = 0; o < 3; ++o) d3[o] = fmaf(hv, W5[n * 3 + o], d3[o]);
          }
#pragma unroll
          for (int o = 0; o < 3; ++o) part[(cb * 3 + o) * kTileM + em] = d3[o];
        }
        umma::fence_before_sync();
        prod_bar();
        if (tid < 3 * kTileM) {
          const int o = tid >> 7, mm = tid & (kTileM - 1);
          const int i = tileX * kTileM + mm;
          if (i < P) {
            const float d = b5[o] + part[(0 * 3 + o) * kTileM + mm] + part[(1 * 3 + o) * kTileM + mm] + part[(2 * 3 + o) * kTileM + mm] +
                            part[(3 * 3 + o) * kTileM + mm];
            a.xyz_prime[(size_t)i * 3 + o] = __fadd_rn(scr.xu[o][mm], d);
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

// packing: chunk j < 3: B[n][k] = W3[n][col(j, k)]; chunk 3 + jj: B[n][k] = W4[n][32 jj + k]   (split hi | lo, K-major, N = 64)
__global__ void k_pack_warp_fwd(const float* __restrict__ w3 /*[64][93]*/, const float* __restrict__ w4 /*[64][64]*/, float* __restrict__ wq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kWfChunks * 64 * kKC) return;
  const int j = i / (64 * kKC), e = i % (64 * kKC);
  const int n = e / kKC, k = e % kKC;
  float v = 0.f;
  if (j < 2) {
    const int p = pos_src_index(j * 32 + k);   // index in [x(3), sin(30), cos(30), t, ...]: the first 63 are the warp MLP's own columns
    if (p < 63) v = __ldg(w3 + (size_t)n * 93 + p);
  } else if (j == 2) {
    if (k < 30) v = __ldg(w3 + (size_t)n * 93 + 63 + k);
  } else {
    v = __ldg(w4 + (size_t)n * 64 + (j - 3) * 32 + k);
  }
  float hi, lo;
  umma::split_tf32(v, hi, lo);
  const uint32_t off = umma::kmajor_off(n, k, 64) >> 2;
  float* blob = wq + (size_t)j * kWfChunkFloats;
  blob[off] = hi;
  blob[64 * kKC + off] = lo;
}

}  // namespace dynrf

using namespace dynrf;

extern "C" {

int dynrf_warp_fwd_wfloats(void) { return kWfWFloats; }
int dynrf_warp_fwd_cst_floats(void) { return kWfCst; }

int dynrf_pack_warp_fwd(const float* w3, const float* w4, float* wq, void* stream) {
  if (!w3 || !w4 || !wq) return DYNRF_E_ARG;
  k_pack_warp_fwd<<<(kWfChunks * 64 * kKC + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(w3, w4, wq);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_warp_fwd_tc2(const dynrf_field_cfg_t* cfg, const float* wq, const float* cst, const float* xyz, int N, int S, const float* tfeat,
                       float* xyz_prime, float* H1T, float* H2T, int ld, void* stream) {
  if (!cfg || !wq || !cst || !xyz || !tfeat || !xyz_prime || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if ((H1T != nullptr) != (H2T != nullptr)) return DYNRF_E_ARG;
  const int P = N * S;
  if (H1T && ld < ((P + kTileM - 1) / kTileM) * kTileM) return DYNRF_E_ARG;
  if ((uintptr_t)wq & 15) return DYNRF_E_ALIGN;
  cudaError_t e = cudaFuncSetAttribute(k_warp_fwd_tc2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWfSmem);
  if (e != cudaSuccess) return (int)e;
  WarpFwdArgs a{wq, cst, xyz, tfeat, P, S, ld, xyz_prime, H1T, H2T};
  const int grid = min((P + kTileM - 1) / kTileM, sm_count());
  k_warp_fwd_tc2<<<grid, kTc3Threads, kWfSmem, reinterpret_cast<cudaStream_t>(stream)>>>(*cfg, a);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
