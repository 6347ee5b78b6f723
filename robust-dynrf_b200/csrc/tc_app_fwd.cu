// Tensor-core forward of the appearance stage on the active (app-masked) samples:
//   DYN = false : static field   (a7 + a11 + a12/a13)  plane*line features at xyz, stride 1  -> basis_mat (72 -> 27)
//                 -> MLPRender_Fea / MLPRender_Fea_TimeEmbedding  (reference models/tensorBase.py MLPRender_Fea*)
//   DYN = true  : dynamic field  (a18 + a14)           features at xyz_prime, strides 1,2,4 -> basis_mat (216 -> 27)
//                 -> MLPRender_Fea_late_view
// One warp-specialised kernel per 128-sample tile (16 producer/epilogue warps + 1 tcgen05 warp, the operand ring of
// dynrf_tc.cuh): the gathered features are written straight into split-TF32 K-major operand chunks (24 rows each), the
// basis GEMM (N = 32) lands in TMEM, its epilogue builds the head's input operand (features, sin/cos encodings or the
// position/time rows), then L1 (-> 128) and L2 (128 -> 128) run on the tensor cores and the 3-wide output layer +
// sigmoid finish in the epilogue.  3-pass split TF32 (hi*hi + hi*lo + lo*hi), fp32 accumulation in TMEM.
//
// TMEM columns: [0,32) basis output | [32,160) L1 pre-activation | [160,288) L2 pre-activation
//
// K order of the head's first layer
//   DYN    : 3 position/time chunks (rows 32,32,16; produce_pos_unit order) | 1 feature chunk (27 features + 5 zero rows)
//   static : 4 encoding chunks -- chunk c holds features d = 8c..8c+7: rows [sin(f) x8 | sin(2f) x8 | cos(f) x8 | cos(2f) x8]
//            | 1 chunk [27 features, viewdir(3), 2 zero rows]
#include "dynrf_tc.cuh"

namespace dynrf {

template <bool DYN>
struct AppFwd {
  static constexpr int kNS = DYN ? 3 : 1;
  static constexpr int kNB = 3 * kNS;          // basis chunks
  static constexpr int kNL1 = DYN ? 4 : 5;     // layer-1 chunks
  static constexpr int kChunks = kNB + kNL1 + 4;
  __host__ __device__ static constexpr int rows(int j) { return j < kNB ? 24 : ((DYN && j == kNB + 2) ? 16 : 32); }
  __host__ __device__ static constexpr int n(int j) { return j < kNB ? 32 : 128; }
  __host__ __device__ static constexpr int woff(int j) {
    int o = 0;
    for (int i = 0; i < j; ++i) o += 2 * n(i) * rows(i);
    return o;
  }
  static constexpr int kWFloats = woff(kChunks);
};

// constants in shared memory: b1[128] | b2[128] | W3t[128][3] | W3v[3][3] (zero when the head has no late view input) | b3[3]
constexpr int kAfCst = 128 + 128 + 384 + 9 + 3;

struct AppFwdArgs {
  const float* wq;
  const float* cst;
  const float* rays;    // [N][6]
  const float* xyz;     // [N*S][3] sample positions (position rows of the dynamic head; lookup coordinates of the static field)
  const float* coords;  // [N*S][3] lookup coordinates (xyz_prime for the dynamic field, xyz for the static one)
  const float* tfeat;   // [N][48] (DYN)
  const int32_t* active_idx;
  const int32_t* active_count;
  int S, ld;
  float* rgb;    // [N*S][3]
  float* featT;  // [32][ld] basis output per active sample (optional, for the backward)
};

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

constexpr size_t kAfSmem = (size_t)(3 * kOpFloats + 2 * kWSlotFloats + kAfCst + 4 * 3 * kTileM) * sizeof(float) + sizeof(TapsSmem) +
                           sizeof(TileScratch<kTileM>) + 128;

template <bool DYN>
__global__ void __launch_bounds__(kTc3Threads, 1) k_app_fwd_tc(dynrf_field_cfg_t cfg, dynrf_vm_t vapp, AppFwdArgs a) {
  using AF = AppFwd<DYN>;
  extern __shared__ __align__(128) float smem[];
  float* abuf = smem;
  float* wbuf = abuf + 3 * kOpFloats;
  float* cst = wbuf + 2 * kWSlotFloats;
  float* part = cst + kAfCst;  // [4][3][128]
  TapsSmem& taps = *reinterpret_cast<TapsSmem*>(part + 4 * 3 * kTileM);
  TileScratch<kTileM>& scr = *reinterpret_cast<TileScratch<kTileM>*>(reinterpret_cast<uint8_t*>(&taps) + sizeof(TapsSmem));
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
  for (int i = tid; i < kAfCst; i += kTc3Threads) cst[i] = __ldg(a.cst + i);
  if (warp == 0) umma::tmem_alloc(tslot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Ring ring{bar_free, bar_w, abuf, wbuf, 0u, *tslot};
  const int count = *a.active_count;
  const int ntiles = (count + kTileM - 1) / kTileM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total_chunks = (uint32_t)my_tiles * AF::kChunks;
  const float* b1 = cst;
  const float* b2 = cst + 128;
  const float* W3t = cst + 256;
  const float* W3v = cst + 640;
  const float* b3 = cst + 649;

  if (warp == kProdThreads / 32) {
    if (lane == 0 && total_chunks > 0) {
      ring_prefetch_w(ring, 0, a.wq + AF::woff(0), 2u * AF::n(0) * AF::rows(0) * 4u);
      for (uint32_t c = 0; c < total_chunks; ++c) {
        const int j = (int)(c % AF::kChunks);
        ring.chunk = c;
        umma::mbar_wait(&bar_full[c % 3], (c / 3) & 1);
        if (j < AF::kNB) ring_issue<32>(ring, 24, 0u, j == 0);
        else if (j < AF::kNB + AF::kNL1) ring_issue<128>(ring, AF::rows(j), 32u, j == AF::kNB);
        else ring_issue<128>(ring, 32, 160u, j == AF::kNB + AF::kNL1);
        if (c + 1 < total_chunks) {
          const int nj = (j + 1) % AF::kChunks;
          ring_prefetch_w(ring, c + 1, a.wq + AF::woff(nj), 2u * AF::n(nj) * AF::rows(nj) * 4u);
        }
      }
    }
    __syncwarp();
  } else {
    const int m = tid & (kTileM - 1), prt = tid >> 7;
    const int q = warp & 3, cb = warp >> 2;
    const int em = q * 32 + lane;
    const AabbNorm nb = make_norm(cfg);
    const int G[3] = {vapp.W[0], vapp.H[0], vapp.L[0]};
    const int ld = a.ld;

    auto wait_chunk = [&](uint32_t c) {
      umma::mbar_wait_backoff(&bar_free[c % 3], (c / 3) & 1);
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
          const float xc = in ? __ldg(a.coords + (size_t)i * 3 + ax) : 0.f;
          scr.xn[ax][tid] = in ? norm_coord(x, nb.mn[ax], nb.inv[ax]) : 0.f;
          scr.xq[ax][tid] = norm_coord(xc, nb.mn[ax], nb.inv[ax]);
          scr.xu[ax][tid] = vd[ax];
        }
      }
      prod_bar();
      for (int item = tid; item < AF::kNS * 3 * kTileM; item += kProdThreads) {
        const int mm = item & (kTileM - 1), t = item >> 7;
        const int si = t / 3, ax = t % 3;
        const Tap1 tp = make_tap(scr.xq[ax][mm], ax == 0 ? G[0] : (ax == 1 ? G[1] : G[2]), 1 << si);
        taps.t[t][mm] = make_int4(tp.i0, tp.i1, __float_as_int(tp.w0), __float_as_int(tp.w1));
      }
      prod_bar();
      const bool act = scr.on[m] != 0;
      const uint32_t base = ring.chunk;
      // ---- basis operand chunks: (stride si, third c): c = 0,1 halves of the 48-channel xy plane; c = 2 the two 12-channel planes
#pragma unroll 1
      for (int j = 0; j < AF::kNB; ++j) {
        ring_wait_free(ring);
        float* buf = ring_abuf(ring);
        const int si = j / 3, c = j % 3;
        if (c < 2) {
          const TapF tx = load_tap(taps, si, 0, m), ty = load_tap(taps, si, 1, m), tz = load_tap(taps, si, 2, m);
          {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (act) r = vm_group_f(vapp.plane[0], vapp.line[0], 48, plane_st(vapp, 0), tx, ty, tz, (c * 6 + prt) * 4);
            store_quad(buf, m, prt, r);
          }
          if ((c == 0) == (prt < 2)) {  // c = 0: threads 0,1 take groups 4,5; c = 1: threads 2,3 take groups 10,11
            const int e = c == 0 ? prt : prt - 2;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (act) r = vm_group_f(vapp.plane[0], vapp.line[0], 48, plane_st(vapp, 0), tx, ty, tz, (c * 6 + 4 + e) * 4);
            store_quad(buf, m, 4 + e, r);
          }
        } else {
          // quads 0..2: plane 1 (x,z | line y) groups 0..2; quads 3..5: plane 2 (y,z | line x) groups 0..2
          for (int qd = prt; qd < 6; qd += 4) {
            const bool p1 = qd < 3;
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (act)
              r = vm_group_f(p1 ? vapp.plane[1] : vapp.plane[2], p1 ? vapp.line[1] : vapp.line[2], 12, plane_st(vapp, p1 ? 1 : 2),
                             load_tap(taps, si, p1 ? 0 : 1, m), load_tap(taps, si, 2, m), load_tap(taps, si, p1 ? 1 : 0, m),
                             (p1 ? qd : qd - 3) * 4);
            store_quad(buf, m, qd, r);
          }
        }
        arrive();
      }
      if (DYN) {
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {
          ring_wait_free(ring);
          float* buf = ring_abuf(ring);
          const int u0 = j * 5, nu = (j == 2) ? 4 : 5;
          for (int u = prt; u < nu; u += 4) produce_pos_unit(u0 + u, scr, a.tfeat, m, buf);
          arrive();
        }
      }
      // ---- epilogue A: basis output (27 features) -> head input operand
      // wait on the most recently produced chunk (for the dynamic head: the last position chunk, which shares its operand buffer
      // with the last basis chunk): tcgen05 operations retire in order, so this also covers the basis GEMM, and a wait on an
      // older chunk of the same buffer could alias with the next phase of its barrier
      wait_chunk(base + AF::kNB - 1 + (DYN ? 3 : 0));
      {
        float f[8];
        tmem_ld8(ring.tmem + ((uint32_t)(q * 32) << 16) + cb * 8, f);
        const bool on = scr.on[em] != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = on ? f[i] : 0.f;
        if (a.featT) {
          store_quad_g(a.featT, ld, cb * 8, jg0 + em, make_float4(f[0], f[1], f[2], f[3]));
          store_quad_g(a.featT, ld, cb * 8 + 4, jg0 + em, make_float4(f[4], f[5], f[6], f[7]));
        }
        if (!DYN) {
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            ring_wait_free(ring);
            if (c4 == cb) {
              float* buf = ring_abuf(ring);
              float s1[8], c1[8], s2[8], c2[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                sincosf(f[i], &s1[i], &c1[i]);
                sincosf(f[i] * 2.0f, &s2[i], &c2[i]);
              }
              store_quad(buf, em, 0, make_float4(s1[0], s1[1], s1[2], s1[3]));
              store_quad(buf, em, 1, make_float4(s1[4], s1[5], s1[6], s1[7]));
              store_quad(buf, em, 2, make_float4(s2[0], s2[1], s2[2], s2[3]));
              store_quad(buf, em, 3, make_float4(s2[4], s2[5], s2[6], s2[7]));
              store_quad(buf, em, 4, make_float4(c1[0], c1[1], c1[2], c1[3]));
              store_quad(buf, em, 5, make_float4(c1[4], c1[5], c1[6], c1[7]));
              store_quad(buf, em, 6, make_float4(c2[0], c2[1], c2[2], c2[3]));
              store_quad(buf, em, 7, make_float4(c2[4], c2[5], c2[6], c2[7]));
            }
            arrive();
          }
          if (cb == 3) {  // rows 27..29 of the feature chunk carry the view direction (zero weights for the time-embedding head)
            f[3] = scr.xu[0][em];
            f[4] = scr.xu[1][em];
            f[5] = scr.xu[2][em];
          }
        }
        ring_wait_free(ring);
        {
          float* buf = ring_abuf(ring);
          store_quad(buf, em, cb * 2, make_float4(f[0], f[1], f[2], f[3]));
          store_quad(buf, em, cb * 2 + 1, make_float4(f[4], f[5], f[6], f[7]));
        }
        arrive();
      }
      // ---- epilogue B: h1 = relu(acc1 + b1) -> L2 operand chunk cb
      wait_chunk(base + AF::kNB + AF::kNL1 - 1);
      {
        float r[32];
        umma::tmem_ld32(ring.tmem + ((uint32_t)(q * 32) << 16) + 32 + cb * 32, r);
        const bool on = scr.on[em] != 0;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) r[jj] = on ? fmaxf(r[jj] + b1[cb * 32 + jj], 0.f) : 0.f;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          ring_wait_free(ring);
          if (c4 == cb) {
            float* buf = ring_abuf(ring);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) store_quad(buf, em, qq, make_float4(r[qq * 4], r[qq * 4 + 1], r[qq * 4 + 2], r[qq * 4 + 3]));
          }
          arrive();
        }
      }
      // ---- epilogue C: h2 = relu(acc2 + b2); 3-wide output layer (+ late view direction); sigmoid
      wait_chunk(base + AF::kChunks - 1);
      {
        float h2[32];
        umma::tmem_ld32(ring.tmem + ((uint32_t)(q * 32) << 16) + 160 + cb * 32, h2);
        float po[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const int n = cb * 32 + jj;
          const float h = fmaxf(h2[jj] + b2[n], 0.f);
          po[0] = fmaf(h, W3t[n * 3 + 0], po[0]);
          po[1] = fmaf(h, W3t[n * 3 + 1], po[1]);
          po[2] = fmaf(h, W3t[n * 3 + 2], po[2]);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) part[(cb * 3 + c) * kTileM + em] = po[c];
      }
      umma::fence_before_sync();
      prod_bar();
      if (tid < 3 * kTileM) {
        const int c = tid >> 7, mm = tid & (kTileM - 1);
        if (scr.on[mm]) {
          float o = b3[c] + part[(0 * 3 + c) * kTileM + mm] + part[(1 * 3 + c) * kTileM + mm] + part[(2 * 3 + c) * kTileM + mm] +
                    part[(3 * 3 + c) * kTileM + mm];
#pragma unroll
          for (int k = 0; k < 3; ++k) o = fmaf(scr.xu[k][mm], W3v[k * 3 + c], o);
          a.rgb[(size_t)scr.idx[mm] * 3 + c] = sigmoidf_(o);
        }
      }
      prod_bar();
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(ring.tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// packing.  head: DYNRF_HEAD_FEA (l1 [128][138]) | DYNRF_HEAD_FEA_TIMEEMB (l1 [128][135]) | DYNRF_HEAD_FEA_LATE_VIEW (l1 [128][107])

template <bool DYN>
__global__ void k_pack_app_fwd(const float* __restrict__ wb /*[27][72 NS]*/, const float* __restrict__ w1, int K1,
                               const float* __restrict__ w2 /*[128][128]*/, int head, float* __restrict__ wq) {
  using AF = AppFwd<DYN>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= AF::kWFloats / 2) return;
  int j = 0, base = 0;
  for (; j < AF::kChunks; ++j) {
    const int sz = AF::n(j) * AF::rows(j);
    if (i < base + sz) break;
    base += sz;
  }
  const int N = AF::n(j), rows = AF::rows(j);
  const int e = i - base;
  const int n = e / rows, k = e % rows;
  float v = 0.f;
  if (j < AF::kNB) {
    const int si = j / 3, c = j % 3;
    if (n < 27) v = __ldg(wb + (size_t)n * (72 * AF::kNS) + si * 72 + c * 24 + k);
  } else if (j < AF::kNB + AF::kNL1) {
    const int jj = j - AF::kNB;
    int col = -1;
    if (DYN) {
      if (jj < 3) col = 27 + pos_src_index(jj * 32 + k);
      else col = k < 27 ? k : -1;
    } else {
      const int sin0 = head == DYNRF_HEAD_FEA ? 30 : 27, cos0 = sin0 + 54;
      if (jj < 4) {
        const int qd = k >> 2, el = k & 3;
        const int trig = qd >> 2, fq = (qd >> 1) & 1, d = jj * 8 + (qd & 1) * 4 + el;
        if (d < 27) col = (trig ? cos0 : sin0) + d * 2 + fq;
      } else {
        if (k < 27) col = k;
        else if (k < 30 && head == DYNRF_HEAD_FEA) col = k;
      }
    }
    if (col >= 0) v = __ldg(w1 + (size_t)n * K1 + col);
  } else {
    const int jj = j - AF::kNB - AF::kNL1;
    v = __ldg(w2 + (size_t)n * 128 + jj * 32 + k);
  }
  float hi, lo;
  umma::split_tf32(v, hi, lo);
  const uint32_t off = umma::kmajor_off(n, k, N) >> 2;
  float* blob = wq + AF::woff(j);
  blob[off] = hi;
  blob[N * rows + off] = lo;
}

}  // namespace dynrf

using namespace dynrf;

extern "C" {

int dynrf_app_fwd_wfloats(int dynamic) { return dynamic ? AppFwd<true>::kWFloats : AppFwd<false>::kWFloats; }
int dynrf_app_fwd_cst_floats(void) { return kAfCst; }

int dynrf_pack_app_fwd(int dynamic, int head, const float* basis_w, const float* w1, const float* w2, float* wq, void* stream) {
  if (!basis_w || !w1 || !w2 || !wq) return DYNRF_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dynamic) {
    if (head != DYNRF_HEAD_FEA_LATE_VIEW) return DYNRF_E_UNSUPPORTED;
    k_pack_app_fwd<true><<<(AppFwd<true>::kWFloats / 2 + 255) / 256, 256, 0, st>>>(basis_w, w1, 107, w2, head, wq);
  } else {
    if (head != DYNRF_HEAD_FEA && head != DYNRF_HEAD_FEA_TIMEEMB) return DYNRF_E_UNSUPPORTED;
    k_pack_app_fwd<false><<<(AppFwd<false>::kWFloats / 2 + 255) / 256, 256, 0, st>>>(basis_w, w1, head == DYNRF_HEAD_FEA ? 138 : 135, w2,
                                                                                    head, wq);
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_app_fwd_tc(const dynrf_field_cfg_t* cfg, int dynamic, const dynrf_vm_t* app, const float* wq, const float* cst,
                     const float* rays, const float* xyz, const float* coords, const float* tfeat, int N, int S,
                     const int32_t* active_idx, const int32_t* active_count, float* rgb, float* featT, int ld, void* stream) {
  if (app && !dynrf_vm_layout_ok(app)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !app || !wq || !cst || !rays || !xyz || !coords || !active_idx || !active_count || !rgb || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if (dynamic && !tfeat) return DYNRF_E_ARG;
  if (featT && ((ld & 3) || ld < ((N * S + kTileM - 1) / kTileM) * kTileM)) return DYNRF_E_ARG;
  if (app->C[0] != 48 || app->C[1] != 12 || app->C[2] != 12) return DYNRF_E_UNSUPPORTED;
  if ((uintptr_t)wq & 15) return DYNRF_E_ALIGN;
  AppFwdArgs a{wq, cst, rays, xyz, coords, tfeat, active_idx, active_count, S, ld, rgb, featT};
  const int grid = min((N * S + kTileM - 1) / kTileM, sm_count());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (dynamic) {
    e = cudaFuncSetAttribute(k_app_fwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAfSmem);
    if (e != cudaSuccess) return (int)e;
    k_app_fwd_tc<true><<<grid, kTc3Threads, kAfSmem, st>>>(*cfg, *app, a);
  } else {
    e = cudaFuncSetAttribute(k_app_fwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAfSmem);
    if (e != cudaSuccess) return (int)e;
    k_app_fwd_tc<false><<<grid, kTc3Threads, kAfSmem, st>>>(*cfg, *app, a);
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
