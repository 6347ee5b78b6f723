// Tensor-core backward of the dynamic density + blending stage (a16 + a17), both tables in one launch.
//
// Per 128-sample tile (warp-specialised like k_dyn_sigma_blend_tc2): recompute layer 1 of both MLPs on tensor cores
// (shared position/time rows against the concatenated weights, then the per-table 3-scale features), finish layer 2 and
// the activations in the epilogue, form dH = w2 * g * [h > 0] for both MLPs, back-propagate it to the 72 + 72 feature
// rows with two more tensor-core GEMMs (N = 80), and scatter d(plane*line) to the plane / line gradients and to
// d/d(xyz_prime) straight from TMEM.  Layer inputs and dH are streamed feature-major to global memory for the
// weight-gradient GEMMs (dynrf_wgrad_gemm); layer-2 weight / bias gradients and the layer-1 bias gradient are reduced
// in registers with a warp transpose-reduction.
//
// TMEM columns: [0,128) layer-1 pre-activations (density | blending) | [128,208) d feat density | [208,288) d feat blending
#include "dynrf_tc.cuh"

namespace dynrf {

constexpr int kSbbChunks = 13;
__host__ __device__ constexpr int sbb_rows(int j) { return j < 2 ? 32 : (j == 2 ? 16 : (j < 9 ? 24 : 32)); }
__host__ __device__ constexpr int sbb_n(int j) { return j < 3 ? 128 : (j < 9 ? 64 : 80); }
__host__ __device__ constexpr int sbb_woff(int j) {
  int o = 0;
  for (int i = 0; i < j; ++i) o += 2 * sbb_n(i) * sbb_rows(i);
  return o;
}
constexpr int kSbbWFloats = sbb_woff(kSbbChunks);
constexpr int kSbbCst = 128 + 128 + 2;  // b1cat | w2cat | b2cat
// XT rows: 80 position/time | 72 density feat | 72 blending feat | ones (225)

struct SbBwdArgs {
  const float* wq;
  const float* cst;
  const float* xyz;
  const float* xyz_prime;
  const uint8_t* valid;
  const float* tfeat;
  const float* gsigma;  // [P] d loss / d sigma   (or NULL)
  const float* gblend;  // [P] d loss / d blending (or NULL)
  int P, S, ld;
  float* XT;    // [225][ld]   (recompute variant: written here; saved variant: written by the forward, untouched)
  const float* HT;  // [128][ld] relu(layer 1) saved by the forward (saved variant)
  float* DHT;   // [128][ld]
  float* gxp;   // [P][3] +=
  float* gw2;   // [128] += (density | blending layer-2 weight gradients)
  float* gb1;   // [128] +=
  float* gb2;   // [2]   +=
};

struct TapsSmemB {
  int4 t[9][kTileM];  // (i0 | v0 << 30, i1 | v1 << 30, raw w0 bits, raw w1 bits)
};

__device__ __forceinline__ Tap1 load_tap1(const TapsSmemB& ts, int si, int axis, int m, int G) {
  const int4 v = ts.t[si * 3 + axis][m];
  Tap1 t;
  t.v0 = (v.x >> 30) & 1;
  t.v1 = (v.y >> 30) & 1;
  t.i0 = v.x & 0x3fffffff;
  t.i1 = v.y & 0x3fffffff;
  t.r0 = __int_as_float(v.z);
  t.r1 = __int_as_float(v.w);
  t.w0 = t.v0 ? t.r0 : 0.f;
  t.w1 = t.v1 ? t.r1 : 0.f;
  const int s = 1 << si;
  t.scale = 0.5f * (float)((G + s - 1) / s - 1);
  return t;
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float4& r) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
  r = make_float4(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c), __uint_as_float(d));
}

struct SbbScratch {
  float xn[3][kTileM];  // normalize(xyz)
  float xq[3][kTileM];  // normalize(xyz_prime)
  float gx[3][kTileM];  // coordinate-gradient accumulator
  int ray[kTileM], on[kTileM];
};
constexpr int kSbbCtxCols = 160;
constexpr size_t kSbbSmem = (size_t)(3 * kOpFloats + 2 * kWSlotFloats + kSbbCst + 2 + 2 * (4 * kTileM + 2 * kTileM)) * sizeof(float) +
                            2 * sizeof(TapsSmemB) + 2 * sizeof(SbbScratch) + 256;
static_assert(kSbbSmem <= 232448, "k_dyn_mlp_bwd_tc shared memory");

// SAVED = false: layer 1 is recomputed (gathers, encodings, 9 operand chunks) and X^T is streamed for the weight-gradient GEMM.
// SAVED = true : the forward kept X^T and H^T (1.4 GB at cfg 2); the tile only forms dH, runs the 4 data-gradient chunks and
//                scatters.  The path is bound by instruction issue, not by HBM, so trading recompute for 0.5 GB of reads pays.
//
// Two tiles in flight (see tc_app_fwd.cu): iteration k forms dH of tile Y = k and submits its data-gradient chunks, then
// scatters tile X = k - 1 from TMEM -- the data-gradient GEMM of Y runs under the scatter of X, and the wait for X's GEMM under
// the epilogue of Y.  TMEM columns of context s (base 160 s): [0,128) layer-1 pre-activations (recompute variant), then
// d feat density [0,80) | d feat blending [80,160) over them (every warp has read layer 1 before it arrives on the first
// data-gradient chunk).
template <bool SAVED>
__global__ void __launch_bounds__(kTc3Threads, 1)
k_dyn_mlp_bwd_tc(dynrf_field_cfg_t cfg, dynrf_vm_t vden, dynrf_vm_t vbl, dynrf_vm_grad_t gden, dynrf_vm_grad_t gbl, SbBwdArgs a) {
  extern __shared__ __align__(128) float smem[];
  float* abuf = smem;
  float* wbuf = abuf + 3 * kOpFloats;
  float* cst = wbuf + 2 * kWSlotFloats;
  float* part0 = cst + kSbbCst + 2;       // [2][4][128]
  float* go0 = part0 + 2 * 4 * kTileM;    // [2][2][128]
  TapsSmemB* tapsv = reinterpret_cast<TapsSmemB*>(go0 + 2 * 2 * kTileM);
  SbbScratch* scrv = reinterpret_cast<SbbScratch*>(tapsv + 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(scrv + 2);
  uint64_t* bar_free = bars;                  // [kNFull] chunk c consumed (see dynrf_tc.cuh)
  uint64_t* bar_w = bars + kNFull;            // [2]
  uint64_t* bar_stage = bars + kNFull + 2;    // [2 contexts][2: layer 1 complete, data gradient complete]
  uint64_t* bar_full = bars + kNFull + 6;     // [kNFull] chunk c written
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 2 * kNFull + 6);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kNFull + 2; ++i) umma::mbar_init(&bars[i], 1);
    for (int i = 0; i < 4; ++i) umma::mbar_init(&bar_stage[i], 1);
    for (int i = 0; i < kNFull; ++i) umma::mbar_init(&bar_full[i], kProdThreads / 32);
    umma::fence_mbar_init();
  }
  for (int i = tid; i < kSbbCst; i += kTc3Threads) cst[i] = __ldg(a.cst + i);
  if (warp == 0) umma::tmem_alloc(tslot, 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  Ring ring{bar_free, bar_w, abuf, wbuf, 0u, *tslot, (uint32_t)kWSlotFloats, 3u, 2u, (uint32_t)kNFull};
  const int P = a.P;
  const int ntiles = (P + kTileM - 1) / kTileM;
  const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  constexpr int kC0 = SAVED ? 9 : 0;                 // first chunk of the schedule that this variant runs
  constexpr int kNC = kSbbChunks - kC0;
  const uint32_t total_chunks = (uint32_t)my_tiles * kNC;
  const float* b1cat = cst;
  const float* w2cat = cst + 128;
  const float* b2cat = cst + 256;

  if (warp == kProdThreads / 32) {
    if (lane == 0 && total_chunks > 0) {
      ring_prefetch_w(ring, 0, a.wq + sbb_woff(kC0), 2u * sbb_n(kC0) * sbb_rows(kC0) * 4u);
      for (uint32_t c = 0; c < total_chunks; ++c) {
        const int j = kC0 + (int)(c % kNC);
        const uint32_t ctx = (c / kNC) & 1u, colb = ctx * kSbbCtxCols;
        ring.chunk = c;
        umma::mbar_wait(&bar_full[c % kNFull], (c / kNFull) & 1);
        if (j < 3) ring_issue<128>(ring, sbb_rows(j), colb, j == 0);
        else if (j < 9) ring_issue<64>(ring, sbb_rows(j), colb + (((j - 3) & 1) ? 64u : 0u), false);
        else ring_issue<80>(ring, sbb_rows(j), colb + (j < 11 ? 0u : 80u), j == 9 || j == 11);
        if (j == 8) umma::commit(&bar_stage[ctx * 2 + 0]);
        else if (j == kSbbChunks - 1) umma::commit(&bar_stage[ctx * 2 + 1]);
        if (c + 1 < total_chunks) {
          const int nj = kC0 + (int)((c + 1) % kNC);
          ring_prefetch_w(ring, c + 1, a.wq + sbb_woff(nj), 2u * sbb_n(nj) * sbb_rows(nj) * 4u);
        }
      }
    }
    __syncwarp();
  } else {
    const int m = tid & (kTileM - 1), prt = tid >> 7;
    const int q = warp & 3, cb = warp >> 2;
    const int em = q * 32 + lane;
    const AabbNorm nb = make_norm(cfg);
    const int G[3] = {vden.W[0], vden.H[0], vden.L[0]};
    const int ld = a.ld;
    float accW2 = 0.f, accB1 = 0.f, accB2 = 0.f;

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
      if (hasY) {
        const int sy = k & 1;
        SbbScratch& scr = scrv[sy];
        TapsSmemB& taps = tapsv[sy];
        float* part = part0 + sy * (4 * kTileM);
        float* go = go0 + sy * (2 * kTileM);
        const uint32_t tcol = ring.tmem + (uint32_t)(sy * kSbbCtxCols) + ((uint32_t)(q * 32) << 16);
        const int i0g = ((int)blockIdx.x + k * (int)gridDim.x) * kTileM;
        if (tid < kTileM) {
          const int i = i0g + tid;
          const bool in = i < P;
          scr.ray[tid] = in ? i / a.S : -1;
          scr.on[tid] = (in && a.valid[i] != 0) ? 1 : 0;
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            const float x = in ? __ldg(a.xyz + (size_t)i * 3 + ax) : 0.f;
            const float xp = in ? __ldg(a.xyz_prime + (size_t)i * 3 + ax) : 0.f;
            scr.xn[ax][tid] = in ? norm_coord(x, nb.mn[ax], nb.inv[ax]) : 0.f;
            scr.xq[ax][tid] = norm_coord(xp, nb.mn[ax], nb.inv[ax]);
            scr.gx[ax][tid] = 0.f;
          }
        }
        prod_bar();
        for (int item = tid; item < 9 * kTileM; item += kProdThreads) {
          const int mm = item & (kTileM - 1), t = item >> 7;
          const int si = t / 3, ax = t % 3;
          const Tap1 tp = make_tap(scr.xq[ax][mm], ax == 0 ? G[0] : (ax == 1 ? G[1] : G[2]), 1 << si);
          taps.t[t][mm] = make_int4(tp.i0 | ((int)tp.v0 << 30), tp.i1 | ((int)tp.v1 << 30), __float_as_int(tp.r0), __float_as_int(tp.r1));
        }
        prod_bar();
        const bool act = scr.on[m] != 0;
        // ---- layer-1 operand chunks (same K order as the forward kernel), streamed to XT
#pragma unroll 1
        for (int j = 0; j < (SAVED ? 0 : 9); ++j) {
          ring_wait_free(ring);
          float* buf = ring_abuf(ring);
          if (j < 3) {
            const int u0 = j * 5, nu = (j == 2) ? 4 : 5;
            for (int u = prt; u < nu; u += 4) produce_pos_unit(u0 + u, scr, a.tfeat, m, buf, a.XT + (size_t)(j * 32) * ld, ld, i0g + m);
            if (j == 0 && prt == 1) a.XT[(size_t)224 * ld + i0g + m] = act ? 1.f : 0.f;
          } else {
            const int si = (j - 3) >> 1;
            const bool isbl = ((j - 3) & 1) != 0;
            const int xrow = (isbl ? 152 : 80) + si * 24;
            {
              const float* pp = isbl ? vbl.plane[0] : vden.plane[0];
              const float* ll = isbl ? vbl.line[0] : vden.line[0];
              float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
              if (act) r = vm_group(pp, ll, 16, plane_st(vden, 0), load_tap1(taps, si, 0, m, G[0]), load_tap1(taps, si, 1, m, G[1]), load_tap1(taps, si, 2, m, G[2]), prt * 4);
              store_quad(buf, m, prt, r);
              store_quad_g(a.XT, ld, xrow + prt * 4, i0g + m, r);
            }
            if (prt < 2) {
              const float* pp = prt == 0 ? (isbl ? vbl.plane[1] : vden.plane[1]) : (isbl ? vbl.plane[2] : vden.plane[2]);
              const float* ll = prt == 0 ? (isbl ? vbl.line[1] : vden.line[1]) : (isbl ? vbl.line[2] : vden.line[2]);
              float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
              const int au = prt == 0 ? 0 : 1, al = prt == 0 ? 1 : 0;
              if (act) r = vm_group(pp, ll, 4, plane_st(vden, prt == 0 ? 1 : 2), load_tap1(taps, si, au, m, G[au]), load_tap1(taps, si, 2, m, G[2]), load_tap1(taps, si, al, m, G[al]), 0);
              store_quad(buf, m, 4 + prt, r);
              store_quad_g(a.XT, ld, xrow + 16 + prt * 4, i0g + m, r);
            }
          }
          arrive();
        }
        // ---- epilogue A: layer 2, activation derivatives, dH -> data-gradient operand chunks, DHT, register reductions
        if (!SAVED) {
          umma::mbar_wait_backoff(&bar_stage[sy * 2 + 0], (uint32_t)((k >> 1) & 1));
          umma::fence_after_sync();
        }
        {
          float h[32];
          const bool on = scr.on[em] != 0;
          float po = 0.f;
          if (SAVED) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) h[jj] = DYNRF_LD_STREAM(a.HT + (size_t)(cb * 32 + jj) * ld + i0g + em);   // already relu'd and masked
          } else {
            umma::tmem_ld32(tcol + cb * 32, h);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) h[jj] = on ? fmaxf(h[jj] + b1cat[cb * 32 + jj], 0.f) : 0.f;
          }
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) po = fmaf(h[jj], w2cat[cb * 32 + jj], po);
          part[cb * kTileM + em] = po;
          prod_bar();
          if (tid < 2 * kTileM) {
            const int which = tid >> 7, mm = tid & (kTileM - 1);
            const int i = i0g + mm;
            float g = 0.f;
            if (scr.on[mm]) {
              const float o = part[(2 * which) * kTileM + mm] + part[(2 * which + 1) * kTileM + mm] + b2cat[which];
              if (which == 0) {
                g = a.gsigma ? __ldg(a.gsigma + i) : 0.f;
                if (cfg.act == DYNRF_ACT_RELU) g = o > 0.f ? g : 0.f;
                else g *= sigmoidf_(o + cfg.density_shift);
              } else {
                const float sg = sigmoidf_(o);
                g = (a.gblend ? __ldg(a.gblend + i) : 0.f) * sg * (1.0f - sg);
              }
            }
            go[which * kTileM + mm] = g;
            accB2 += g;
          }
          prod_bar();
          const float gm = go[(cb >> 1) * kTileM + em];
          float red[32];
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) red[jj] = h[jj] * gm;  // d w2[n] contributions
          accW2 += warp_col_reduce32(red, lane);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            h[jj] = h[jj] > 0.f ? w2cat[cb * 32 + jj] * gm : 0.f;  // dH
            red[jj] = h[jj];
          }
          accB1 += warp_col_reduce32(red, lane);
#pragma unroll 1
          for (int c4 = 0; c4 < 4; ++c4) {
            if (c4 == cb) {   // only the warps that write the chunk wait for its operand buffer (see dynrf_tc.cuh, kNFull)
              ring_wait_free(ring);
              float* buf = ring_abuf(ring);
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                const float4 v = make_float4(h[qq * 4], h[qq * 4 + 1], h[qq * 4 + 2], h[qq * 4 + 3]);
                store_quad(buf, em, qq, v);
                store_quad_g(a.DHT, ld, cb * 32 + qq * 4, i0g + em, v);
              }
            }
            arrive();
          }
        }
      }
      if (hasX) {
        // ---- epilogue B of tile X: d feat (TMEM) -> plane / line gradients + coordinate gradients
        const int sx = (k - 1) & 1;
        SbbScratch& scr = scrv[sx];
        const TapsSmemB& taps = tapsv[sx];
        const float* go = go0 + sx * (2 * kTileM);
        const uint32_t tcol = ring.tmem + (uint32_t)(sx * kSbbCtxCols) + ((uint32_t)(q * 32) << 16);
        const int i0g = ((int)blockIdx.x + (k - 1) * (int)gridDim.x) * kTileM;
        const bool act_m = scr.on[m] != 0;
        umma::mbar_wait_backoff(&bar_stage[sx * 2 + 1], (uint32_t)(((k - 1) >> 1) & 1));
        umma::fence_after_sync();
        // a sample whose upstream gradient is exactly zero (relu'd density <= 0: half the samples at initialisation, all of
        // empty space in a trained scene) has dH = 0, hence d feat = 0: its lookups and reductions are skipped.
        // The 16-channel (x, y) planes of both tables are split over the four threads of a sample (one channel group each); the
        // four 4-channel planes {density, blending} x {(x, z), (y, z)} go one to each thread, so that the 16 warps carry equal
        // work (with both small planes of a table on threads 0 and 1, half the warps waited at the tile's closing barrier:
        // 14.8 % of the stall samples).
#pragma unroll 1
        for (int tb = 0; tb < 2; ++tb) {
          const uint32_t colbase = 80u * tb;
          const bool act = act_m && go[tb * kTileM + m] != 0.f;
          const float* pp = tb ? vbl.plane[0] : vden.plane[0];
          const float* ll = tb ? vbl.line[0] : vden.line[0];
          float* gp = tb ? gbl.plane[0] : gden.plane[0];
          float* gl = tb ? gbl.line[0] : gden.line[0];
          float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll 1
          for (int si = 0; si < 3; ++si) {
            float4 g4;
            tmem_ld4(tcol + colbase + si * 24 + prt * 4, g4);
            const float3 r = vm_group_bwd_w(pp, ll, gp, gl, 16, plane_st(vden, 0), load_tap1(taps, si, 0, m, G[0]), load_tap1(taps, si, 1, m, G[1]),
                                              load_tap1(taps, si, 2, m, G[2]), prt * 4, g4, act, lane);
            g0 += r.x;
            g1 += r.y;
            g2 += r.z;
          }
          if (act) {
            atomicAdd(&scr.gx[0][m], g0);
            atomicAdd(&scr.gx[1][m], g1);
            atomicAdd(&scr.gx[2][m], g2);
          }
        }
        {
          const int tb = prt >> 1, pw = prt & 1;  // warp-uniform: table, which of its two 4-channel planes
          const uint32_t colbase = 80u * tb;
          const bool act = act_m && go[tb * kTileM + m] != 0.f;
          const float* pp = pw == 0 ? (tb ? vbl.plane[1] : vden.plane[1]) : (tb ? vbl.plane[2] : vden.plane[2]);
          const float* ll = pw == 0 ? (tb ? vbl.line[1] : vden.line[1]) : (tb ? vbl.line[2] : vden.line[2]);
          float* gp = pw == 0 ? (tb ? gbl.plane[1] : gden.plane[1]) : (tb ? gbl.plane[2] : gden.plane[2]);
          float* gl = pw == 0 ? (tb ? gbl.line[1] : gden.line[1]) : (tb ? gbl.line[2] : gden.line[2]);
          float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll 1
          for (int si = 0; si < 3; ++si) {
            float4 g4;
            tmem_ld4(tcol + colbase + si * 24 + 16 + pw * 4, g4);
            if (pw == 0) {  // plane (x, z), line y
              const float3 r = vm_group_bwd_w(pp, ll, gp, gl, 4, plane_st(vden, 1), load_tap1(taps, si, 0, m, G[0]), load_tap1(taps, si, 2, m, G[2]),
                                                load_tap1(taps, si, 1, m, G[1]), 0, g4, act, lane);
              g0 += r.x;
              g2 += r.y;
              g1 += r.z;
            } else {  // plane (y, z), line x
              const float3 r = vm_group_bwd_w(pp, ll, gp, gl, 4, plane_st(vden, 2), load_tap1(taps, si, 1, m, G[1]), load_tap1(taps, si, 2, m, G[2]),
                                                load_tap1(taps, si, 0, m, G[0]), 0, g4, act, lane);
              g1 += r.x;
              g2 += r.y;
              g0 += r.z;
            }
          }
          if (act) {
            atomicAdd(&scr.gx[0][m], g0);
            atomicAdd(&scr.gx[1][m], g1);
            atomicAdd(&scr.gx[2][m], g2);
          }
        }
        umma::fence_before_sync();
        prod_bar();
        if (tid < kTileM && scr.on[tid]) {
          const int i = i0g + tid;
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) a.gxp[(size_t)i * 3 + ax] += scr.gx[ax][tid] * nb.inv[ax];
        }
        prod_bar();
      }
    }
    atomicAdd(a.gw2 + cb * 32 + lane, accW2);
    atomicAdd(a.gb1 + cb * 32 + lane, accB1);
    if (tid < 2 * kTileM) {
      const float s2 = warp_sum(accB2);
      if (lane == 0) atomicAdd(a.gb2 + (tid >> 7), s2);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(ring.tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// packing: chunks 0..8 exactly as the forward kernel (k_pack_sb_umma), chunks 9..12 the data-gradient operands
//   B[k < 80][nn < 32] = W1x[32 c + nn][feat k]  (k < 72), x = density for chunks 9,10, blending for 11,12

__global__ void k_pack_sb_bwd(const float* __restrict__ w_den /*[64][152]*/, const float* __restrict__ w_bl, float* __restrict__ wq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kSbbWFloats / 2) return;
  int j = 0, base = 0;
  for (; j < kSbbChunks; ++j) {
    const int sz = sbb_n(j) * sbb_rows(j);
    if (i < base + sz) break;
    base += sz;
  }
  const int N = sbb_n(j), rows = sbb_rows(j);
  const int e = i - base;
  const int r = e / rows, k = e % rows;
  float v = 0.f;
  if (j < 3) {
    const int col = 72 + pos_src_index(j * 32 + k);
    v = r < 64 ? __ldg(w_den + (size_t)r * 152 + col) : __ldg(w_bl + (size_t)(r - 64) * 152 + col);
  } else if (j < 9) {
    const int col = ((j - 3) >> 1) * 24 + k;
    v = ((j - 3) & 1) ? __ldg(w_bl + (size_t)r * 152 + col) : __ldg(w_den + (size_t)r * 152 + col);
  } else {
    const float* w = j < 11 ? w_den : w_bl;
    const int nn = 32 * ((j - 9) & 1) + k;
    v = r < 72 ? __ldg(w + (size_t)nn * 152 + r) : 0.f;
  }
  float hi, lo;
  umma::split_tf32(v, hi, lo);
  const uint32_t off = umma::kmajor_off(r, k, N) >> 2;
  float* blob = wq + sbb_woff(j);
  blob[off] = hi;
  blob[N * rows + off] = lo;
}

__global__ void k_sb_bwd_colmap(int32_t* __restrict__ map /*[80]*/) {
  const int r = threadIdx.x;
  if (r < 80) map[r] = 72 + pos_src_index(r);
}

}  // namespace dynrf

using namespace dynrf;

extern "C" {

int dynrf_sb_bwd_wfloats(void) { return kSbbWFloats; }

int dynrf_pack_sb_bwd(const float* w_den, const float* w_bl, float* wq, int32_t* colmap80, void* stream) {
  if (!w_den || !w_bl || !wq || !colmap80) return DYNRF_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  k_pack_sb_bwd<<<(kSbbWFloats / 2 + 255) / 256, 256, 0, st>>>(w_den, w_bl, wq);
  k_sb_bwd_colmap<<<1, 96, 0, st>>>(colmap80);
  DYNRF_CHECK_LAUNCH();
  return 0;
}

int dynrf_dyn_mlp_bwd_tc(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                         const dynrf_vm_grad_t* gdensity, const dynrf_vm_grad_t* gblending, const float* wq, const float* cst,
                         const float* xyz, const float* xyz_prime, const uint8_t* valid, const float* tfeat, const float* gsigma,
                         const float* gblend, int N, int S, int ld, float* XT, const float* HT_saved, float* DHT, float* gxp,
                         float* gw2, float* gb1, float* gb2, void* stream) {
  if (density && blending && (!dynrf_vm_layout_ok(density) || !dynrf_vm_same_layout(density, blending))) return DYNRF_E_UNSUPPORTED;
  if (!cfg || !density || !blending || !gdensity || !gblending || !wq || !cst || !xyz || !xyz_prime || !valid || !tfeat || !XT || !DHT ||
      !gxp || !gw2 || !gb1 || !gb2 || N <= 0 || S <= 0 || (ld & 3) || ld < ((N * S + kTileM - 1) / kTileM) * kTileM)
    return DYNRF_E_ARG;
  for (int i = 0; i < 3; ++i) {
    const int c = i == 0 ? 16 : 4;
    if (density->C[i] != c || blending->C[i] != c) return DYNRF_E_UNSUPPORTED;
    if (density->H[i] != blending->H[i] || density->W[i] != blending->W[i] || density->L[i] != blending->L[i]) return DYNRF_E_UNSUPPORTED;
  }
  if ((uintptr_t)wq & 15) return DYNRF_E_ALIGN;
  SbBwdArgs a{wq, cst, xyz, xyz_prime, valid, tfeat, gsigma, gblend, N * S, S, ld, XT, HT_saved, DHT, gxp, gw2, gb1, gb2};
  const int grid = min((N * S + kTileM - 1) / kTileM, sm_count());
  cudaError_t e;
  if (HT_saved) {
    e = cudaFuncSetAttribute(k_dyn_mlp_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSbbSmem);
    if (e != cudaSuccess) return (int)e;
    k_dyn_mlp_bwd_tc<true><<<grid, kTc3Threads, kSbbSmem, reinterpret_cast<cudaStream_t>(stream)>>>(*cfg, *density, *blending, *gdensity,
                                                                                                   *gblending, a);
  } else {
    e = cudaFuncSetAttribute(k_dyn_mlp_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSbbSmem);
    if (e != cudaSuccess) return (int)e;
    k_dyn_mlp_bwd_tc<false><<<grid, kTc3Threads, kSbbSmem, reinterpret_cast<cudaStream_t>(stream)>>>(*cfg, *density, *blending, *gdensity,
                                                                                                    *gblending, a);
  }
  DYNRF_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
