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

__device__ __forceinline__ void tmem_ld8_(uint32_t taddr, float (&r)[8]) {
  uint32_t u[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}

// chunk schedule.  L1 chunks (DYN: 3 position/time + 1 feature; static: 4 encoding + 1 feature/viewdir), 4 L2 chunks,
// 4 chunks of dH1 = dH2 W2 and 4 chunks of the layer-1 data gradient (DYN: the 27 feature columns, N = 32; static: all
// 160 operand rows, N = 160, because the gradient also flows through the sin / cos encodings of the features).
template <bool DYN>
struct HeadBwd {
  static constexpr int kNL1 = DYN ? 4 : 5;
  static constexpr int kChunks = kNL1 + 12;
  static constexpr int kN4 = DYN ? 32 : 160;
  static constexpr int kXRows = DYN ? 113 : 161;  // X1T rows (last = ones)
  __host__ __device__ static constexpr int rows(int j) { return (DYN && j == 2) ? 16 : 32; }
  __host__ __device__ static constexpr int n(int j) { return j < kNL1 + 8 ? 128 : kN4; }
  __host__ __device__ static constexpr int woff(int j) {
    int o = 0;
    for (int i = 0; i < j; ++i) o += 2 * n(i) * rows(i);
    return o;
  }
  static constexpr int kWFloats = woff(kChunks);
};
constexpr int kHbWSlot = 2 * 160 * kKC;  // floats per weight slot (largest chunk: N = 160)

// constants in shared memory: b1[128] | b2[128] | W3t[128][3] | W3v[3][3] | b3[3]
constexpr int kHbCst = 128 + 128 + 384 + 9 + 3;

struct HeadBwdArgs {
  const float* wq;      // packed split weights for the chunks
  const float* cst;     // kHbCst
  const float* rays;    // [N][6]
  const float* xyz;     // [N*S][3]
  const float* tfeat;   // [N][48] (dynamic head)
  const float* featT;   // [32][ld] basis output saved by the forward (rows 27..31 zero)
  const float* grgb;    // [N*S][3]
  const int32_t* active_idx;
  const int32_t* active_count;
  int S, ld;
  // feature-major outputs, row stride ld
  float* X1T;   // dynamic [113]: 80 position/time rows (this kernel's K order) | 32 feat rows | ones
                // static  [161]: 4 x 32 encoding rows | 27 feat + 3 viewdir + 2 zero rows | ones
  float* H1T;   // [128]
  float* DH2T;  // [128]
  float* DH1T;  // [128]
  float* DFT;   // [32]
  float* gw3v;  // [3][3] += d W3[:, 128+k]   (row c, col k)
  float* gb3;   // [3]    += d b3
  float* gw3h;  // [3][128] += d W3[:, :128]
  float* gb2;   // [128]  += d b2
  float* gvd;   // [N][3] += d loss / d viewdir through layer 1 (static MLP_Fea head, optional)
};

// The tensor-core warp spends ~0.2 us on the MMAs of a chunk but a 32-40 KB weight blob takes ~1 us to arrive from L2, and the
// 16-17 chunks of a tile are consumed back to back: with one chunk of weight look-ahead the kernel ran at the weight latency.
// Three weight slots (two chunks of look-ahead) and two operand buffers fit the same shared memory.
constexpr uint32_t kHbBufs = 2, kHbSlots = 3;

// ---- two tiles in flight per CTA --------------------------------------------------------------------------------------------
// A tile is a chain of four dependent GEMM groups (L1 -> L2 -> dH1 -> dX) with an epilogue between each pair; with one tile per
// CTA the 16 producer warps spent 32 % of their stall samples at the group-completion mbarriers and 21 % at the barriers behind
// them, while the tensor core idled during every epilogue.  The CTA therefore runs TWO tiles (contexts A, B; tiles dealt
// alternately) through the stages  S (scratch + L1 operand), E1, E2, E3, E4  in strict alternation, B two stages behind A, with
// E4 of a context's previous tile executed at the head of the S slot of its next tile (see HeadStageSeq below):
//     ES(A) . E1(A) . E2(A) ES(B) E3(A) E1(B) ES(A') E2(B) E1(A') E3(B) E2(A') ES(B') ...
// so that between the production of a group's last operand chunk and the read-back of its accumulator the warps always run one
// long stage of the other tile.  Ring, weight slots and the tensor-core warp are shared; the ring order of the chunks follows from the
// same generator (HeadStageSeq) on both sides.  Every (context, group) has its own completion mbarrier (one phase per tile).
//
// TMEM columns of context s (base 256 s): acc1 [0,128) L1 pre-activation | acc2 [128,256) L2 pre-activation, then acc3 = dH1
// before the ReLU mask (every warp has read acc2 before it arrives on the first dH1-group chunk) | acc4 = dX over acc1 (and, for
// the static head's N = 160, the first 32 columns of acc3): all reads of acc1 / acc3 precede the barrier in E3.
struct BwdScratch {
  float xn[3][kTileM];  // normalize(xyz)
  float xu[3][kTileM];  // view direction
  int ray[kTileM], idx[kTileM], on[kTileM];
};

// Stage sequence of a CTA.  Per context the stages of a tile are  ES = [E4 of the context's PREVIOUS tile] + [S of this tile],
// E1, E2, E3 -- E4 (the read-back of the last GEMM group, short for the dynamic head) is merged with the next tile's S so that every
// slot of the alternation is a long stage: with E4 as a slot of its own, the GEMM group submitted by the stage before it had only
// that short slot to complete in, and the following read-back waited (10.8 % + 9.4 % of the stall samples of the 5-stage schedule).
// Contexts alternate strictly, B two stages behind A:  ES(A) . E1(A) . E2(A) ES(B) E3(A) E1(B) ES(A') E2(B) E1(A') E3(B) ...
struct HeadStageSeq {
  int T, step, st0, st1, tl0, tl1;   // scalars (no indexed arrays: they would live in local memory)
  bool fin0, fin1;
  int ctx, stage, tile;   // the stage to execute now: context, stage 0..3, tile ordinal (of this CTA) of its S / E1..E3 part
  bool do_e4, do_s;       // stage 0: read back tile - 2 of the context / start tile `tile`
  __device__ __forceinline__ void init(int T_) {
    T = T_;
    step = -1;
    st0 = st1 = 0;
    tl0 = 0;
    tl1 = 1;
    fin0 = fin1 = false;
  }
  // advance to the next executed stage; false when every tile of the CTA is finished
  __device__ __forceinline__ bool next() {
    for (;;) {
      if (fin0 && fin1) return false;
      ++step;
      const int c = step & 1;
      if (c == 1 && step < 5) continue;   // context B starts two stages behind A
      if (c ? fin1 : fin0) continue;
      const int st = c ? st1 : st0, tl = c ? tl1 : tl0;
      ctx = c;
      stage = st;
      tile = tl;
      int nst = st, ntl = tl;
      bool fin = false;
      if (st == 0) {
        do_e4 = tl >= 2;
        do_s = tl < T;
        fin = !do_s;
        if (do_s) nst = 1;
      } else {
        do_e4 = do_s = false;
        if (++nst == 4) {
          nst = 0;
          ntl = tl + 2;
        }
      }
      if (c) {
        st1 = nst;
        tl1 = ntl;
        fin1 = fin;
      } else {
        st0 = nst;
        tl0 = ntl;
        fin0 = fin;
      }
      if (st == 0 && !do_e4 && !do_s) continue;   // a context without tiles
      return true;
    }
  }
};

constexpr int kHbCtxCols = 256;
constexpr size_t kHbSmem = (size_t)(kHbBufs * kOpFloats + kHbSlots * kHbWSlot + kHbCst + 2 * (4 * 3 * kTileM + 3 * kTileM)) * sizeof(float) +
                           2 * sizeof(BwdScratch) + 256;
static_assert(kHbSmem <= 232448, "k_head_bwd_tc shared memory");

template <bool DYN>
__global__ void __launch_bounds__(kTc3Threads, 1) k_head_bwd_tc(dynrf_field_cfg_t cfg, HeadBwdArgs a) {
  using HB = HeadBwd<DYN>;
  constexpr int kL2 = HB::kNL1, kD2 = HB::kNL1 + 4, kD1 = HB::kNL1 + 8;  // first chunk of each group
  extern __shared__ __align__(128) float smem[];
  float* abuf = smem;
  float* wbuf = abuf + kHbBufs * kOpFloats;
  float* cst = wbuf + kHbSlots * kHbWSlot;
  float* part0 = cst + kHbCst;            // [2][4][3][128] partial output dots
  float* go0 = part0 + 2 * 4 * 3 * kTileM;  // [2][3][128] d loss / d (pre-sigmoid output)
  BwdScratch* scrv = reinterpret_cast<BwdScratch*>(go0 + 2 * 3 * kTileM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scrv + 2);
  uint64_t* bar_free = bars;                   // [kNFull]: "chunk c is consumed" lives on barrier c % kNFull, not on the operand buffer's index
  uint64_t* bar_w = bars + kNFull;             // [kHbSlots]
  uint64_t* bar_stage = bars + kNFull + 3;     // [2 contexts][4 groups]
  uint64_t* bar_full = bars + kNFull + 11;     // [kNFull]: "chunk c is written"
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * kNFull + 11);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kNFull + 3; ++i) umma::mbar_init(&bars[i], 1);
    for (int i = 0; i < kNFull; ++i) umma::mbar_init(&bar_full[i], kProdThreads / 32);
    for (int i = 0; i < 8; ++i) umma::mbar_init(&bar_stage[i], 1);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < kHbCst; i += kTc3Threads) cst[i] = __ldg(a.cst + i);
  if (warp == 0) umma::tmem_alloc(tslot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Ring ring{bar_free, bar_w, abuf, wbuf, 0u, *tslot, (uint32_t)kHbWSlot, kHbBufs, kHbSlots, (uint32_t)kNFull};
  const int count = *a.active_count;
  const int ntiles = (count + kTileM - 1) / kTileM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const float* b1 = cst;
  const float* b2 = cst + 128;
  const float* W3t = cst + 256;   // [128][3]
  const float* W3v = cst + 640;   // [k][c]
  const float* b3 = cst + 649;

  if (warp == kProdThreads / 32) {
    // ===================== tensor-core warp: follows the stage sequence chunk by chunk =====================
    if (lane == 0 && my_tiles > 0) {
      // chunk cursor = (stage sequence position, index inside the stage); `pf` runs kHbSlots - 1 chunks ahead for the weights
      struct Cur {
        HeadStageSeq sq;
        int i, n, j0;
        bool ok;
        __device__ __forceinline__ void load() {   // first chunk / chunk count of the current stage
          n = sq.stage == 0 ? (sq.do_s ? HB::kNL1 : 0) : 4;
          j0 = sq.stage == 0 ? 0 : HB::kNL1 + 4 * (sq.stage - 1);
        }
        __device__ __forceinline__ void start(int T) {
          sq.init(T);
          i = 0;
          n = 0;
          ok = true;
          adv();
        }
        __device__ __forceinline__ void adv() {   // move to the next chunk of the sequence
          ++i;
          while (ok && i >= n) {
            ok = sq.next();
            if (ok) load();
            i = 0;
          }
        }
        __device__ __forceinline__ int j() const { return j0 + i; }
      };
      Cur cu, pf;
      cu.start(my_tiles);
      pf.start(my_tiles);
      for (uint32_t c = 0; c + 1 < kHbSlots && pf.ok; ++c) {
        const int pj = pf.j();
        ring_prefetch_w(ring, c, a.wq + HB::woff(pj), 2u * HB::n(pj) * HB::rows(pj) * 4u);
        pf.adv();
      }
      for (uint32_t c = 0; cu.ok; ++c) {
        const int j = cu.j(), s = cu.sq.ctx;
        const uint32_t col = (uint32_t)(s * kHbCtxCols);
        ring.chunk = c;
        umma::mbar_wait(&bar_full[c % kNFull], (c / kNFull) & 1);
        if (j < kL2) ring_issue<128>(ring, HB::rows(j), col, j == 0);
        else if (j < kD2) ring_issue<128>(ring, 32, col + 128u, j == kL2);
        else if (j < kD1) ring_issue<128>(ring, 32, col + 128u, j == kD2);   // acc3 over acc2
        else if (DYN) ring_issue<32>(ring, 32, col, j == kD1);               // acc4 over acc1
        else ring_issue<160>(ring, 32, col, j == kD1);
        if (j == kL2 - 1) umma::commit(&bar_stage[s * 4 + 0]);
        else if (j == kD2 - 1) umma::commit(&bar_stage[s * 4 + 1]);
        else if (j == kD1 - 1) umma::commit(&bar_stage[s * 4 + 2]);
        else if (j == HB::kChunks - 1) umma::commit(&bar_stage[s * 4 + 3]);
        if (pf.ok) {
          const int pj = pf.j();
          ring_prefetch_w(ring, c + kHbSlots - 1, a.wq + HB::woff(pj), 2u * HB::n(pj) * HB::rows(pj) * 4u);
          pf.adv();
        }
        cu.adv();
      }
    }
    __syncwarp();
  } else {
    // ===================== producer / epilogue warps =====================
    const int m = tid & (kTileM - 1), prt = tid >> 7;
    const int q = warp & 3, cb = warp >> 2;  // TMEM lane quadrant, 32-column block   (m == q * 32 + lane, prt == cb)
    const int em = q * 32 + lane;            // sample row this thread owns in the epilogues
    const AabbNorm nb = make_norm(cfg);
    const int ld = a.ld;
    float acc_vd[9], acc_b3[3], acc_w3[3] = {0.f, 0.f, 0.f}, acc_b2 = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) acc_vd[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) acc_b3[i] = 0.f;

    auto arrive = [&]() {
      umma::fence_before_sync();   // orders this warp's tcgen05.ld before the MMAs that follow the arrival
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[ring.chunk % kNFull]);
      ring.chunk++;
    };

    HeadStageSeq sq;
    sq.init(my_tiles);
#pragma unroll 1
    while (sq.next()) {
      const int s = sq.ctx;
      BwdScratch& scr = scrv[s];
      float* part = part0 + s * (4 * 3 * kTileM);
      float* go = go0 + s * (3 * kTileM);
      const uint32_t tcol = ring.tmem + (uint32_t)(s * kHbCtxCols) + ((uint32_t)(q * 32) << 16);
      const uint32_t par = (uint32_t)((sq.tile >> 1) & 1);
      const int jg0 = ((int)blockIdx.x + sq.tile * (int)gridDim.x) * kTileM;
      if (sq.stage == 0 && sq.do_e4) {
        // ---- epilogue 4 of the context's previous tile (tile - 2): dfeat
        const uint32_t par4 = (uint32_t)(((sq.tile - 2) >> 1) & 1);
        const int jg4 = ((int)blockIdx.x + (sq.tile - 2) * (int)gridDim.x) * kTileM;
        umma::mbar_wait_backoff(&bar_stage[s * 4 + 3], par4);
        umma::fence_after_sync();
        if (DYN) {
          if (cb == 0) {
            float r4[32];
            umma::tmem_ld32(tcol, r4);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              float4 v = make_float4(r4[qq * 4], r4[qq * 4 + 1], r4[qq * 4 + 2], r4[qq * 4 + 3]);
              if (qq == 6) v.w = 0.f;            // only 27 feature rows carry a gradient
              if (qq == 7) v = make_float4(0.f, 0.f, 0.f, 0.f);
              store_quad_g(a.DFT, ld, qq * 4, jg4 + em, v);
            }
          }
        } else {
          // dX columns: [32 cb, 32 cb + 32) = d(sin f, sin 2f, cos f, cos 2f) of this thread's 8 features, [128 + 8 cb, +8) = d f
          float dx[32], df[8];
          umma::tmem_ld32(tcol + cb * 32, dx);
          tmem_ld8_(tcol + 128 + cb * 8, df);
          const bool on = scr.on[em] != 0;
          if (cb == 3 && on && a.gvd) {
            const int ry = scr.ray[em];
            atomicAdd(a.gvd + (size_t)ry * 3 + 0, df[3]);
            atomicAdd(a.gvd + (size_t)ry * 3 + 1, df[4]);
            atomicAdd(a.gvd + (size_t)ry * 3 + 2, df[5]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float fi = on ? __ldg(a.featT + (size_t)(cb * 8 + i) * ld + jg4 + em) : 0.f;   // the basis output again (L2 hit)
            float s1, c1, s2, c2;
            sincosf(fi, &s1, &c1);
            sincosf(fi * 2.0f, &s2, &c2);
            // d/df [sin f, sin 2f, cos f, cos 2f] = [cos f, 2 cos 2f, -sin f, -2 sin 2f]
            float g = df[i];
            g = fmaf(dx[i], c1, g);
            g = fmaf(dx[8 + i], 2.0f * c2, g);
            g = fmaf(dx[16 + i], -s1, g);
            g = fmaf(dx[24 + i], -2.0f * s2, g);
            df[i] = (on && cb * 8 + i < 27) ? g : 0.f;
          }
          store_quad_g(a.DFT, ld, cb * 8, jg4 + em, make_float4(df[0], df[1], df[2], df[3]));
          store_quad_g(a.DFT, ld, cb * 8 + 4, jg4 + em, make_float4(df[4], df[5], df[6], df[7]));
        }
        umma::fence_before_sync();
      }
      if (sq.stage == 0 && sq.do_s) {
        prod_bar();   // every warp is done with this context's scratch (E4 of its previous tile)
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
        if (DYN) {
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
                const float* fp = a.featT + (size_t)(qq * 4) * ld + jg0 + m;
                v = make_float4(__ldg(fp), __ldg(fp + ld), __ldg(fp + 2 * (size_t)ld), __ldg(fp + 3 * (size_t)ld));
              }
              store_quad(buf, m, qq, v);
              store_quad_g(a.X1T, ld, 80 + qq * 4, jg0 + m, v);
            }
            if (prt == 0) a.X1T[(size_t)112 * ld + jg0 + m] = on ? 1.f : 0.f;  // ones row: its weight-gradient column is d b1
            arrive();
          }
        } else {
          // ---- L1 operand: 4 encoding chunks (chunk c <- warps with cb == c) + the [features | viewdir] chunk; X1T rows 0..159
          const bool on = scr.on[em] != 0;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = on ? __ldg(a.featT + (size_t)(cb * 8 + i) * ld + jg0 + em) : 0.f;
          // encodings first (and their X1T stream), ring wait afterwards: the warps with cb == 3 would otherwise start their
          // sincosf only when chunk 0's MMAs have released an operand buffer
          float sv[2][8], cv[2][8];
#pragma unroll
          for (int h = 0; h < 2; ++h) {   // frequency 1, 2
#pragma unroll
            for (int i = 0; i < 8; ++i) sincosf(h ? f[i] * 2.0f : f[i], &sv[h][i], &cv[h][i]);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              store_quad_g(a.X1T, ld, cb * 32 + (h * 2 + half) * 4, jg0 + em,
                           make_float4(sv[h][half * 4], sv[h][half * 4 + 1], sv[h][half * 4 + 2], sv[h][half * 4 + 3]));
              store_quad_g(a.X1T, ld, cb * 32 + (4 + h * 2 + half) * 4, jg0 + em,
                           make_float4(cv[h][half * 4], cv[h][half * 4 + 1], cv[h][half * 4 + 2], cv[h][half * 4 + 3]));
            }
          }
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            if (c4 == cb) {
              ring_wait_free(ring);
              float* buf = ring_abuf(ring);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  store_quad(buf, em, h * 2 + half, make_float4(sv[h][half * 4], sv[h][half * 4 + 1], sv[h][half * 4 + 2], sv[h][half * 4 + 3]));
                  store_quad(buf, em, 4 + h * 2 + half, make_float4(cv[h][half * 4], cv[h][half * 4 + 1], cv[h][half * 4 + 2], cv[h][half * 4 + 3]));
                }
              }
            }
            arrive();
          }
          ring_wait_free(ring);
          {
            float* buf = ring_abuf(ring);
            float4 v0 = make_float4(f[0], f[1], f[2], f[3]), v1 = make_float4(f[4], f[5], f[6], f[7]);
            if (cb == 3) {  // rows 27..29: view direction (zero weights for the time-embedding head)
              v0.w = on ? scr.xu[0][em] : 0.f;
              v1.x = on ? scr.xu[1][em] : 0.f;
              v1.y = on ? scr.xu[2][em] : 0.f;
            }
            store_quad(buf, em, cb * 2, v0);
            store_quad(buf, em, cb * 2 + 1, v1);
            store_quad_g(a.X1T, ld, 128 + cb * 8, jg0 + em, v0);
            store_quad_g(a.X1T, ld, 128 + cb * 8 + 4, jg0 + em, v1);
            if (cb == 0) a.X1T[(size_t)160 * ld + jg0 + em] = on ? 1.f : 0.f;
          }
          arrive();
        }
      } else if (sq.stage == 1) {
        // ---- epilogue 1: h1 = relu(acc1 + b1) -> L2 operand chunk cb, H1T
        umma::mbar_wait_backoff(&bar_stage[s * 4 + 0], par);
        umma::fence_after_sync();
        float r[32];
        umma::tmem_ld32(tcol + cb * 32, r);
        const bool on = scr.on[em] != 0;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) r[jj] = on ? fmaxf(r[jj] + b1[cb * 32 + jj], 0.f) : 0.f;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          // only the warps that write the chunk wait for its operand buffer; the others arrive at once (the "written" barriers
          // are indexed by c % kNFull, so an early arrival cannot fall into the still-open phase of an older chunk: between two
          // chunks kNFull apart every warp passes a CTA-wide barrier of the producers)
          if (c4 == cb) {
            ring_wait_free(ring);
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
      } else if (sq.stage == 2) {
        // ---- epilogue 2: h2 = relu(acc2 + b2); output layer; d(out); dH2 -> dgrad-2 operand chunk cb
        umma::mbar_wait_backoff(&bar_stage[s * 4 + 1], par);
        umma::fence_after_sync();
        float h2[32];
        umma::tmem_ld32(tcol + 128 + cb * 32, h2);
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
            acc_b3[c] += g3[c];
          }
#pragma unroll
          for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) acc_vd[c * 3 + k] = fmaf(scr.xu[k][tid], g3[c], acc_vd[c * 3 + k]);
        }
        prod_bar();
        const float g0 = go[em], g1 = go[kTileM + em], g2 = go[2 * kTileM + em];
        {
          // d W3[c][n] += sum_m h2[m][n] g_c[m]: column sums over this warp's 32 samples, kept in registers (lane <-> column)
          float red[32];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float gc = c == 0 ? g0 : (c == 1 ? g1 : g2);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) red[jj] = h2[jj] * gc;
            acc_w3[c] += warp_col_reduce32(red, lane);
          }
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int n = cb * 32 + jj;
            const float d = fmaf(W3t[n * 3 + 0], g0, fmaf(W3t[n * 3 + 1], g1, W3t[n * 3 + 2] * g2));
            h2[jj] = h2[jj] > 0.f ? d : 0.f;
            red[jj] = h2[jj];
          }
          acc_b2 += warp_col_reduce32(red, lane);
        }
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          if (c4 == cb) {
            ring_wait_free(ring);
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
      } else if (sq.stage == 3) {
        // ---- epilogue 3: dH1 = acc3 * [h1 > 0] -> dgrad-1 operand chunk cb, DH1T.  All TMEM reads (acc3 and the sign of
        // acc1) finish before any warp arrives on the first dgrad-1 chunk: the dgrad-1 GEMM overwrites acc1 (and acc3's first columns).
        umma::mbar_wait_backoff(&bar_stage[s * 4 + 2], par);
        umma::fence_after_sync();
        float r3[32];
        const bool on = scr.on[em] != 0;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float t3[16], t1[16];
          umma::tmem_ld16(tcol + 128 + cb * 32 + hf * 16, t3);
          umma::tmem_ld16(tcol + cb * 32 + hf * 16, t1);
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) r3[hf * 16 + jj] = (on && (t1[jj] + b1[cb * 32 + hf * 16 + jj]) > 0.f) ? t3[jj] : 0.f;
        }
        umma::fence_before_sync();
        prod_bar();
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          if (c4 == cb) {
            ring_wait_free(ring);
            float* buf = ring_abuf(ring);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const float4 v = make_float4(r3[qq * 4], r3[qq * 4 + 1], r3[qq * 4 + 2], r3[qq * 4 + 3]);
              store_quad(buf, em, qq, v);
              store_quad_g(a.DH1T, ld, cb * 32 + qq * 4, jg0 + em, v);
            }
          }
          arrive();
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(a.gw3h + c * 128 + cb * 32 + lane, acc_w3[c]);
    atomicAdd(a.gb2 + cb * 32 + lane, acc_b2);
    if (tid < kTileM) {
#pragma unroll
      for (int i = 0; i < 9; ++i) {   // one reduction per warp: 32 same-address reductions per warp (31 of them adding zero) from
        const float v = warp_sum(acc_vd[i]);   // every CTA at the end of the kernel serialised in L2 -- 9-14 % of the stall samples
        if (lane == 0) atomicAdd(a.gw3v + i, v);   // sat at the closing barrier behind them
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float v = warp_sum(acc_b3[i]);
        if (lane == 0) atomicAdd(a.gb3 + i, v);
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(ring.tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// weight packing for k_head_bwd_tc: chunk j holds B[row r < N_j][k < rows_j] (split hi | lo, K-major):
//   L1 chunks   : B[n][k] = W1[n][col(j, k)]   (col: this kernel's K order -> nn.Linear input column, -1 = zero row)
//   L2 chunks   : B[n][k] = W2[n][32 jj + k]
//   dgrad-2     : B[kin][nn] = W2[32 jj + nn][kin]                     (dH1 = dH2 W2)
//   dgrad-1     : B[r][nn] = W1[32 jj + nn][col of operand row r]      (dynamic: r < 27 feature columns; static: all 160 rows)

template <bool DYN>
__device__ __forceinline__ int head_l1_col(int r, int head) {  // operand row r of layer 1 -> nn.Linear input column (-1: none)
  if (DYN) {
    if (r < 80) return 27 + pos_src_index(r);
    return (r - 80) < 27 ? (r - 80) : -1;
  }
  const int sin0 = head == DYNRF_HEAD_FEA ? 30 : 27, cos0 = sin0 + 54;
  if (r < 128) {
    const int c = r >> 5, k = r & 31;
    const int qd = k >> 2, el = k & 3;
    const int trig = qd >> 2, fq = (qd >> 1) & 1, d = c * 8 + (qd & 1) * 4 + el;
    return d < 27 ? (trig ? cos0 : sin0) + d * 2 + fq : -1;
  }
  const int k = r - 128;
  if (k < 27) return k;
  if (k < 30 && head == DYNRF_HEAD_FEA) return k;
  return -1;
}

template <bool DYN>
__global__ void k_pack_head_bwd(const float* __restrict__ w1, int K1, const float* __restrict__ w2 /*[128][128]*/, int head,
                                float* __restrict__ wq) {
  using HB = HeadBwd<DYN>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HB::kWFloats / 2) return;
  int j = 0, base = 0;
  for (; j < HB::kChunks; ++j) {
    const int sz = HB::n(j) * HB::rows(j);
    if (i < base + sz) break;
    base += sz;
  }
  const int N = HB::n(j), rows = HB::rows(j);
  const int e = i - base;
  const int r = e / rows, k = e % rows;
  float v = 0.f;
  if (j < HB::kNL1) {
    const int col = head_l1_col<DYN>(((DYN && j == 3) ? 80 : j * 32) + k, head);   // the dynamic head's third chunk has 16 rows
    if (col >= 0) v = __ldg(w1 + (size_t)r * K1 + col);
  } else if (j < HB::kNL1 + 4) {
    v = __ldg(w2 + (size_t)r * 128 + 32 * (j - HB::kNL1) + k);
  } else if (j < HB::kNL1 + 8) {
    v = __ldg(w2 + (size_t)(32 * (j - HB::kNL1 - 4) + k) * 128 + r);
  } else {
    const int jj = j - HB::kNL1 - 8;
    const int col = DYN ? (r < 27 ? r : -1) : head_l1_col<DYN>(r, head);
    if (col >= 0) v = __ldg(w1 + (size_t)(32 * jj + k) * K1 + col);
  }
  float hi, lo;
  umma::split_tf32(v, hi, lo);
  const uint32_t off = umma::kmajor_off(r, k, N) >> 2;
  float* blob = wq + HB::woff(j);
  blob[off] = hi;
  blob[N * rows + off] = lo;
}

// X1T row -> nn.Linear input column of layer 1 (-1: padding row, -2: the ones row whose gradient column is d bias)
template <bool DYN>
__global__ void k_head_bwd_colmap(int head, int32_t* __restrict__ map) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int kX = HeadBwd<DYN>::kXRows;
  if (r < kX - 1) map[r] = head_l1_col<DYN>(r, head);
  else if (r == kX - 1) map[r] = -2;
}

// ------------------------------------------------------------------------------------------------
// basis_mat + lookup backward on the active list (SIMT, 64-sample tiles): DFT -> d basis, plane / line grads, d coords

constexpr int kBT2 = 64;
constexpr int kBL2 = kBT2 + 4;
using BScratch2 = TileScratch<kBT2>;
constexpr size_t kBasisBwdSmem = (size_t)((72 + 32 + 80) * kBL2 + 2 * kChunkK * 80) * sizeof(float) + sizeof(BScratch2);

#ifndef DYNRF_BASIS_CTAS
#define DYNRF_BASIS_CTAS 2   // CTAs per SM the kernel is compiled for (3: 80 registers + ~300 B of spills per thread, 24 warps per SM)
#endif
template <bool DYN>
__global__ void __launch_bounds__(kMlpThreads, DYNRF_BASIS_CTAS)
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
      {
        Acc<80, kBT2> acc;
        acc_init<80, kBT2>(acc, nullptr, tid >> 4);
        acc_gemm<80, kBT2, kBL2>(acc, dF, 32, wbB + (size_t)si * 32 * 80, wbuf, tid);
        acc_store<80, kBT2, kBL2, false>(acc, dAc, 72, tid);
      }
      // one pass over the taps: scatter d(plane*line) and recompute the forward features for the basis weight gradient
      scatter_vm_rows<kBT2, kBL2>(vm, gvm, scr.xq, 1 << si, 0, dAc, tid, scr.on, gcoord ? scr.gx : nullptr, Ac);
      __syncthreads();
      {
        float accB[5][2];
        wgrad_zero(accB);
        wgrad_acc<5, 2, kBT2, kBL2>(accB, Ac, 72, dF, tid);
        wgrad_flush(accB, gbasis_wt + (size_t)si * 72 * 32, 72, 32, tid);
      }
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

int dynrf_head_bwd_wfloats(int dynamic) { return dynamic ? HeadBwd<true>::kWFloats : HeadBwd<false>::kWFloats; }
int dynrf_head_bwd_cst_floats(void) { return kHbCst; }
int dynrf_head_bwd_xrows(int dynamic) { return dynamic ? HeadBwd<true>::kXRows : HeadBwd<false>::kXRows; }

int dynrf_pack_head_bwd(int dynamic, int head, const float* w1, const float* w2, float* wq, int32_t* colmap, void* stream) {
  if (!w1 || !w2 || !wq || !colmap) return DYNRF_E_ARG;
  if (dynamic) {
    if (head != DYNRF_HEAD_FEA_LATE_VIEW) return DYNRF_E_UNSUPPORTED;
    k_pack_head_bwd<true><<<(HeadBwd<true>::kWFloats / 2 + 255) / 256, 256, 0, S_(stream)>>>(w1, 107, w2, head, wq);
    k_head_bwd_colmap<true><<<1, 192, 0, S_(stream)>>>(head, colmap);
  } else {
    if (head != DYNRF_HEAD_FEA && head != DYNRF_HEAD_FEA_TIMEEMB) return DYNRF_E_UNSUPPORTED;
    k_pack_head_bwd<false><<<(HeadBwd<false>::kWFloats / 2 + 255) / 256, 256, 0, S_(stream)>>>(w1, head == DYNRF_HEAD_FEA ? 138 : 135, w2,
                                                                                                head, wq);
    k_head_bwd_colmap<false><<<1, 192, 0, S_(stream)>>>(head, colmap);
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_head_bwd_tc(const dynrf_field_cfg_t* cfg, int dynamic, const float* wq, const float* cst, const float* rays, const float* xyz,
                      const float* tfeat, const float* featT, const float* grgb, const int32_t* active_idx,
                      const int32_t* active_count, int N, int S, int ld, float* X1T, float* H1T, float* DH2T, float* DH1T,
                      float* DFT, float* gw3h, float* gw3v, float* gb3, float* gb2, float* gviewdir, void* stream) {
  if (!cfg || !wq || !cst || !rays || !xyz || (dynamic && !tfeat) || !featT || !grgb || !active_idx || !active_count || !X1T || !H1T ||
      !DH2T || !DH1T || !DFT || !gw3h || !gw3v || !gb3 || !gb2 || N <= 0 || S <= 0 || ld < N * S || (ld & 3))
    return DYNRF_E_ARG;
  if ((uintptr_t)wq & 15) return DYNRF_E_ALIGN;
  HeadBwdArgs a{wq, cst, rays, xyz, tfeat, featT, grgb, active_idx, active_count, S, ld, X1T, H1T, DH2T, DH1T, DFT, gw3v, gb3, gw3h, gb2,
                dynamic ? nullptr : gviewdir};
  const int grid = min((N * S + kTileM - 1) / kTileM, sm_count());
  cudaError_t e;
  if (dynamic) {
    e = cudaFuncSetAttribute(k_head_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHbSmem);
    if (e != cudaSuccess) return (int)e;
    k_head_bwd_tc<true><<<grid, kTc3Threads, kHbSmem, S_(stream)>>>(*cfg, a);
  } else {
    e = cudaFuncSetAttribute(k_head_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHbSmem);
    if (e != cudaSuccess) return (int)e;
    k_head_bwd_tc<false><<<grid, kTc3Threads, kHbSmem, S_(stream)>>>(*cfg, a);
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_basis_bwd(const dynrf_field_cfg_t* cfg, int dynamic, const dynrf_vm_t* app, const dynrf_vm_grad_t* gapp, const float* wbB,
                    const float* coords, const float* DFT, int ld, int N, int S, const int32_t* active_idx,
                    const int32_t* active_count, float* gbasis_wt, float* gcoord, void* stream) {
  if (app && !dynrf_vm_layout_ok(app)) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !app || !gapp || !wbB || !coords || !DFT || !active_idx || !active_count || !gbasis_wt || N <= 0 || S <= 0) return DYNRF_E_ARG;
  if (app->C[0] != 48 || app->C[1] != 12 || app->C[2] != 12) return DYNRF_E_UNSUPPORTED;
  // planes (x,y) (x,z) (y,z) of one grid
  if (app->W[1] != app->W[0] || app->W[2] != app->H[0] || app->H[1] != app->L[0] || app->H[2] != app->L[0] || app->L[1] != app->H[0] ||
      app->L[2] != app->W[0])
    return DYNRF_E_UNSUPPORTED;
  const int grid = min((N * S + kBT2 - 1) / kBT2, DYNRF_BASIS_CTAS * sm_count());
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
