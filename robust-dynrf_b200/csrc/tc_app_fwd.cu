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

// ---- two tiles in flight per CTA --------------------------------------------------------------------------------------------
// A tile is a chain of dependent stages (gathers -> basis GEMM -> head input -> L1 -> L2 -> output layer).  With one tile per CTA
// the producer warps sit at an mbarrier while the tensor core finishes each GEMM group (29 % of the stall samples of the
// one-tile kernel: 12.7 % before the L1 read-back, 12.0 % before the L2 read-back, 3.9 % before the basis read-back) and the
// tensor core idles while they gather.  The CTA therefore keeps TWO tiles in flight in a software pipeline skewed by half a
// tile: iteration k starts tile Y = k (context k & 1) and finishes tile X = k - 1 (the other context), in the order
//
//     S0(Y) scratch + taps | S1a(Y) basis operand chunks | S3(X) L1 read-back -> L2 operand chunks |
//     S1b(Y) position/time chunks (dynamic head) | S2(Y) basis read-back -> L1 operand chunks | S4(X) L2 read-back -> rgb
//
// so that every read-back of a GEMM group is separated from the production of its last operand chunk by a long SIMT stage of the
// OTHER tile: L1(X) runs on the tensor core under S0 + S1a(Y), L2(X) under S1b + S2(Y), basis(Y) under S3(X).  The operand ring
// and the weight slots are shared; the ring order is the static sequence Y.basis, X.L2, Y.(rest of L1), which the tensor-core
// warp follows with the same generator (AppSeq).  Every (context, GEMM group) has its own completion mbarrier (tcgen05.commit
// after the group's last chunk, one phase per tile of that context): a wait on an operand-buffer barrier several chunks back
// would alias with a later phase of that barrier.
//
// TMEM columns of context s (base 160 s): [0,32) basis output | [32,160) L1 pre-activation, later L2 pre-activation (every warp
// has read its L1 columns before it arrives on the first L2 chunk, and that chunk's MMA needs all 16 arrivals).
struct FwdScratch {
  float xn[3][kTileM];  // normalize(xyz)
  float xq[3][kTileM];  // normalize(lookup coordinates)
  float xu[3][kTileM];  // view direction
  int ray[kTileM], idx[kTileM], on[kTileM];
};

// position in the CTA's chunk sequence: iteration k, phase ph (0: Y.basis, 1: X.L2, 2: Y.rest of L1), index i inside the phase
template <bool DYN>
struct AppSeq {
  using AF = AppFwd<DYN>;
  int k, ph, i, T;  // T = number of tiles of this CTA; iterations k = 0..T
  __device__ __forceinline__ int len() const {
    if (ph == 0) return k < T ? AF::kNB : 0;
    if (ph == 1) return k >= 1 ? 4 : 0;
    return k < T ? AF::kNL1 : 0;
  }
  __device__ __forceinline__ void norm() {
    while (k <= T && i >= len()) {
      i = 0;
      if (++ph == 3) {
        ph = 0;
        ++k;
      }
    }
  }
  __device__ __forceinline__ bool valid() const { return k <= T; }
  __device__ __forceinline__ int j() const { return ph == 0 ? i : (ph == 1 ? AF::kNB + AF::kNL1 + i : AF::kNB + i); }
  __device__ __forceinline__ int s() const { return (ph == 1 ? k - 1 : k) & 1; }
};

constexpr int kAfCtxCols = 160;
constexpr size_t kAfSmem = (size_t)(3 * kOpFloats + 2 * kWSlotFloats + kAfCst + 2 * 4 * 3 * kTileM) * sizeof(float) + 2 * sizeof(TapsSmem) +
                           2 * sizeof(FwdScratch) + 256;
static_assert(kAfSmem <= 232448, "k_app_fwd_tc shared memory");

template <bool DYN>
__global__ void __launch_bounds__(kTc3Threads, 1) k_app_fwd_tc(dynrf_field_cfg_t cfg, dynrf_vm_t vapp, AppFwdArgs a) {
  using AF = AppFwd<DYN>;
  extern __shared__ __align__(128) float smem[];
  float* abuf = smem;
  float* wbuf = abuf + 3 * kOpFloats;
  float* cst = wbuf + 2 * kWSlotFloats;
  float* part0 = cst + kAfCst;  // [2][4][3][128]
  TapsSmem* tapsv = reinterpret_cast<TapsSmem*>(part0 + 2 * 4 * 3 * kTileM);
  FwdScratch* scrv = reinterpret_cast<FwdScratch*>(tapsv + 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scrv + 2);
  uint64_t* bar_free = bars;                    // [kNFull] chunk c is consumed: barrier c % kNFull (see dynrf_tc.cuh)
  uint64_t* bar_w = bars + kNFull;              // [2]
  uint64_t* bar_stage = bars + kNFull + 2;      // [2 contexts][3 GEMM groups: basis, L1, L2]
  uint64_t* bar_full = bars + kNFull + 8;       // [kNFull] chunk c is written: barrier c % kNFull
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * kNFull + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kNFull + 2; ++i) umma::mbar_init(&bars[i], 1);
    for (int i = 0; i < kNFull; ++i) umma::mbar_init(&bar_full[i], kProdThreads / 32);
    for (int i = 0; i < 6; ++i) umma::mbar_init(&bar_stage[i], 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < kAfCst; i += kTc3Threads) cst[i] = __ldg(a.cst + i);
  if (warp == 0) umma::tmem_alloc(tslot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Ring ring{bar_free, bar_w, abuf, wbuf, 0u, *tslot, (uint32_t)kWSlotFloats, 3u, 2u, (uint32_t)kNFull};
  const int count = *a.active_count;
  const int ntiles = (count + kTileM - 1) / kTileM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const float* b1 = cst;
  const float* b2 = cst + 128;
  const float* W3t = cst + 256;
  const float* W3v = cst + 640;
  const float* b3 = cst + 649;

  if (warp == kProdThreads / 32) {
    // ===================== tensor-core warp: follows the chunk sequence =====================
    if (lane == 0 && my_tiles > 0) {
      AppSeq<DYN> sq{0, 0, 0, my_tiles};
      sq.norm();
      ring_prefetch_w(ring, 0, a.wq + AF::woff(sq.j()), 2u * AF::n(sq.j()) * AF::rows(sq.j()) * 4u);
      for (uint32_t c = 0; sq.valid(); ++c) {
        const int j = sq.j(), s = sq.s();
        ring.chunk = c;
        umma::mbar_wait(&bar_full[c % kNFull], (c / kNFull) & 1);
        const uint32_t col = (uint32_t)(s * kAfCtxCols);
        if (j < AF::kNB) ring_issue<32>(ring, 24, col, j == 0);
        else if (j < AF::kNB + AF::kNL1) ring_issue<128>(ring, AF::rows(j), col + 32u, j == AF::kNB);
        else ring_issue<128>(ring, 32, col + 32u, j == AF::kNB + AF::kNL1);
        if (j == AF::kNB - 1) umma::commit(&bar_stage[s * 3 + 0]);
        else if (j == AF::kNB + AF::kNL1 - 1) umma::commit(&bar_stage[s * 3 + 1]);
        else if (j == AF::kChunks - 1) umma::commit(&bar_stage[s * 3 + 2]);
        ++sq.i;
        sq.norm();
        if (sq.valid()) ring_prefetch_w(ring, c + 1, a.wq + AF::woff(sq.j()), 2u * AF::n(sq.j()) * AF::rows(sq.j()) * 4u);
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
    auto arrive = [&]() {
      umma::fence_before_sync();   // orders this warp's tcgen05.ld before the MMAs that follow the arrival
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[ring.chunk % kNFull]);
      ring.chunk++;
    };

    // the active-list entry of the NEXT tile is fetched one iteration ahead (threads 0..127): it heads the dependent-load chain
    // of S0 (index -> position / direction), during which the other twelve warps wait at the stage's barrier
    int i_pref = 0;
    bool in_pref = false;
    if (tid < kTileM && my_tiles > 0) {
      const int jg = (int)blockIdx.x * kTileM + tid;
      in_pref = jg < count;
      i_pref = in_pref ? __ldg(a.active_idx + jg) : 0;
    }
#pragma unroll 1
    for (int k = 0; k <= my_tiles; ++k) {
      const bool hasY = k < my_tiles, hasX = k >= 1;
      const int sy = k & 1, sx = (k - 1) & 1;
      const uint32_t pary = (uint32_t)((k >> 1) & 1), parx = (uint32_t)(((k - 1) >> 1) & 1);
      const int jgY = ((int)blockIdx.x + k * (int)gridDim.x) * kTileM;
      FwdScratch& scrY = scrv[sy];
      FwdScratch& scrX = scrv[sx];
      const TapsSmem& tapsY = tapsv[sy];
      const uint32_t tcolY = ring.tmem + (uint32_t)(sy * kAfCtxCols), tcolX = ring.tmem + (uint32_t)(sx * kAfCtxCols);
      if (hasY) {
        // ---- S0(Y): per-sample scratch, taps
        if (tid < kTileM) {
          const bool in = in_pref;
          const int i = i_pref;
          if (k + 1 < my_tiles) {
            const int jgn = jgY + (int)gridDim.x * kTileM + tid;
            in_pref = jgn < count;
            i_pref = in_pref ? __ldg(a.active_idx + jgn) : 0;
          }
          scrY.idx[tid] = i;
          scrY.ray[tid] = in ? i / a.S : -1;
          scrY.on[tid] = in ? 1 : 0;
          float vd[3] = {0.f, 0.f, 0.f};
          if (in) ray_viewdir(a.rays, i / a.S, vd);
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            const float x = in ? __ldg(a.xyz + (size_t)i * 3 + ax) : 0.f;
            const float xc = in ? __ldg(a.coords + (size_t)i * 3 + ax) : 0.f;
            scrY.xn[ax][tid] = in ? norm_coord(x, nb.mn[ax], nb.inv[ax]) : 0.f;
            scrY.xq[ax][tid] = norm_coord(xc, nb.mn[ax], nb.inv[ax]);
            scrY.xu[ax][tid] = vd[ax];
          }
        }
        prod_bar();
        for (int item = tid; item < AF::kNS * 3 * kTileM; item += kProdThreads) {
          const int mm = item & (kTileM - 1), t = item >> 7;
          const int si = t / 3, ax = t % 3;
          const Tap1 tp = make_tap(scrY.xq[ax][mm], ax == 0 ? G[0] : (ax == 1 ? G[1] : G[2]), 1 << si);
          tapsv[sy].t[t][mm] = make_int4(tp.i0, tp.i1, __float_as_int(tp.w0), __float_as_int(tp.w1));
        }
        prod_bar();
        // ---- S1a(Y): basis operand chunks (stride si, third c): c = 0,1 halves of the 48-channel xy plane; c = 2 the two
        //      12-channel planes
        // (issuing the lookups of the three chunks of a stride together -- 24-30 loads in flight per thread -- was measured and
        //  lost 7 %: the 96-register budget of a 17-warp CTA turns the extra live values into spills)
        const bool act = scrY.on[m] != 0;
#pragma unroll 1
        for (int j = 0; j < AF::kNB; ++j) {
          ring_wait_free(ring);
          float* buf = ring_abuf(ring);
          const int si = j / 3, c = j % 3;
          if (c < 2) {
            const TapF tx = load_tap(tapsY, si, 0, m), ty = load_tap(tapsY, si, 1, m), tz = load_tap(tapsY, si, 2, m);
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
                               load_tap(tapsY, si, p1 ? 0 : 1, m), load_tap(tapsY, si, 2, m), load_tap(tapsY, si, p1 ? 1 : 0, m),
                               (p1 ? qd : qd - 3) * 4);
              store_quad(buf, m, qd, r);
            }
          }
          arrive();
        }
      }
      if (hasX) {
        // ---- S3(X): h1 = relu(acc1 + b1) -> L2 operand chunk cb
        umma::mbar_wait_backoff(&bar_stage[sx * 3 + 1], parx);
        umma::fence_after_sync();
        float r[32];
        umma::tmem_ld32(tcolX + ((uint32_t)(q * 32) << 16) + 32 + cb * 32, r);
        const bool on = scrX.on[em] != 0;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) r[jj] = on ? fmaxf(r[jj] + b1[cb * 32 + jj], 0.f) : 0.f;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          if (c4 == cb) {
            ring_wait_free(ring);
            float* buf = ring_abuf(ring);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) store_quad(buf, em, qq, make_float4(r[qq * 4], r[qq * 4 + 1], r[qq * 4 + 2], r[qq * 4 + 3]));
          }
          arrive();
        }
      }
      if (hasY) {
        // ---- S1b(Y): the three position/time chunks of the dynamic head's first layer
        if (DYN) {
#pragma unroll 1
          for (int j = 0; j < 3; ++j) {
            ring_wait_free(ring);
            float* buf = ring_abuf(ring);
            const int u0 = j * 5, nu = (j == 2) ? 4 : 5;
            for (int u = prt; u < nu; u += 4) produce_pos_unit(u0 + u, scrY, a.tfeat, m, buf);
            arrive();
          }
        }
        // ---- S2(Y): basis output (27 features) -> head input operand
        umma::mbar_wait_backoff(&bar_stage[sy * 3 + 0], pary);
        umma::fence_after_sync();
        float f[8];
        tmem_ld8(tcolY + ((uint32_t)(q * 32) << 16) + cb * 8, f);
        const bool on = scrY.on[em] != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = on ? f[i] : 0.f;
        if (a.featT) {
          store_quad_g(a.featT, ld, cb * 8, jgY + em, make_float4(f[0], f[1], f[2], f[3]));
          store_quad_g(a.featT, ld, cb * 8 + 4, jgY + em, make_float4(f[4], f[5], f[6], f[7]));
        }
        if (!DYN) {
          // the encodings are computed BEFORE the ring wait: chunk c4 is written by the warps with cb == c4 only, and with three
          // operand buffers the cb == 3 warps otherwise start their 16 sincosf after chunk 0's MMAs (14.8 % of the stall samples)
          float s1[8], c1[8], s2[8], c2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            sincosf(f[i], &s1[i], &c1[i]);
            sincosf(f[i] * 2.0f, &s2[i], &c2[i]);
          }
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            if (c4 == cb) {
              ring_wait_free(ring);
              float* buf = ring_abuf(ring);
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
            f[3] = scrY.xu[0][em];
            f[4] = scrY.xu[1][em];
            f[5] = scrY.xu[2][em];
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
      if (hasX) {
        // ---- S4(X): h2 = relu(acc2 + b2); 3-wide output layer (+ late view direction); sigmoid
        float* part = part0 + sx * (4 * 3 * kTileM);
        umma::mbar_wait_backoff(&bar_stage[sx * 3 + 2], parx);
        umma::fence_after_sync();
        {
          float h2[32];
          umma::tmem_ld32(tcolX + ((uint32_t)(q * 32) << 16) + 32 + cb * 32, h2);
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
          if (scrX.on[mm]) {
            float o = b3[c] + part[(0 * 3 + c) * kTileM + mm] + part[(1 * 3 + c) * kTileM + mm] + part[(2 * 3 + c) * kTileM + mm] +
                      part[(3 * 3 + c) * kTileM + mm];
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) o = fmaf(scrX.xu[kk][mm], W3v[kk * 3 + c], o);
            a.rgb[(size_t)scrX.idx[mm] * 3 + c] = sigmoidf_(o);
          }
        }
        prod_bar();
      }
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
