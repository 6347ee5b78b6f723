/*
 * dynrf_b200.h -- C ABI of libdynrf_b200.so: the B200 (sm_100a) ray-march hot path of
 * RoDynRF's static + dynamic TensoRF fields.
 *
 * The reference (facebookresearch/robust-dynrf) has no FFI: its boundary is Python call
 * signatures.  Each entry point below replaces the ATen kernel
 * chain behind one reference function; the file:line it replaces is cited per function.
 * The host-side mirror of those Python signatures lives in robust-dynrf_b200/models and
 * robust-dynrf_b200/renderer.py and binds this library with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; fp32 unless stated; masks are uint8.
 *   - the library allocates nothing persistent, never synchronises the device, never
 *     throws, never exits.  All work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 = OK; >0 = cudaError_t from the launch; <0 = DYNRF_E_* argument error.
 *   - planes are channel-LAST: the C values of a texel are contiguous; the order of the two texel dimensions is given per
 *     plane by dynrf_vm_t.sv / .su (the drop-in modules store plane 0 as [H][W][C] and planes 1, 2 as [W][H][C]);
 *     line i is [L][C] for (1,C,L,1).
 *     matMode = {{0,1},{0,2},{1,2}}, vecMode = {2,1,0} (models/tensorBase.py:326-327):
 *     plane i has W = G[matMode[i][0]], H = G[matMode[i][1]]; line i has L = G[vecMode[i]].
 *   - MLP layers are "packed": wt is [Kpad][Npad] (input-major, zero padded) built by
 *     dynrf_pack_linear from nn.Linear's [N][K] weight; bias is [Npad] (zeros if none).
 */
#ifndef DYNRF_B200_H
#define DYNRF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DYNRF_ABI_VERSION 3

#define DYNRF_E_ARG (-1)         /* null pointer / non-positive size */
#define DYNRF_E_UNSUPPORTED (-2) /* configuration outside the compiled kernels (component counts, widths) */
#define DYNRF_E_ALIGN (-3)       /* pointer not 16-byte aligned where vector loads need it */

#define DYNRF_RAY_NDC 0
#define DYNRF_RAY_CONTRACT 1

#define DYNRF_ACT_RELU 0
#define DYNRF_ACT_SOFTPLUS 1

#define DYNRF_HEAD_FEA 0            /* MLPRender_Fea               models/tensorBase.py:81-111  */
#define DYNRF_HEAD_FEA_TIMEEMB 1    /* MLPRender_Fea_TimeEmbedding models/tensorBase.py:114-145 */
#define DYNRF_HEAD_FEA_LATE_VIEW 2  /* MLPRender_Fea_late_view     models/tensorBase.py:148-183 */

/* One vector-matrix decomposed quantity (density, blending or appearance). */
typedef struct {
  const float* plane[3]; /* channel-last texels, see sv / su */
  const float* line[3];  /* [L][C]    */
  int32_t C[3], H[3], W[3], L[3];
  /* texel strides of plane i: element (h, w, c) lives at ((h * sv[i] + w * su[i]) * C[i] + c).
   * [H][W][C] ("x-fastest") is sv = W, su = 1; [W][H][C] ("z-fastest") is sv = 1, su = H.  The host stores plane 0 (x, y)
   * x-fastest and planes 1, 2 (x, z) / (y, z) z-fastest: consecutive samples of a ray are consecutive lanes of a warp and
   * advance mostly in z, so their taps in the two z planes fall into a few contiguous 128-byte lines instead of 32 rows. */
  int32_t sv[3], su[3];
} dynrf_vm_t;

/* Gradient buffers with the layout of dynrf_vm_t (accumulated into with atomics; caller zeroes). */
typedef struct {
  float* plane[3];
  float* line[3];
} dynrf_vm_grad_t;

/* A packed nn.Linear. */
typedef struct {
  const float* wt;   /* [Kpad][Npad] */
  const float* bias; /* [Npad]       */
  int32_t K, N, Kpad, Npad;
} dynrf_layer_t;

/* Gradient of a packed nn.Linear (same [Kpad][Npad] / [Npad] layout; accumulated into, caller zeroes). */
typedef struct {
  float* dwt;
  float* dbias;
} dynrf_layer_grad_t;

/* Per-field constants (TensorBase.__init__, models/tensorBase.py:282-339). */
typedef struct {
  float aabb_min[3], aabb_max[3];
  float near_, far_;
  float distance_scale;   /* 25 in every config */
  float weight_thres;     /* rayMarch_weight_thres, 1e-4 */
  float density_shift;    /* softplus only */
  int32_t act;            /* DYNRF_ACT_* (feature2density, models/tensorBase.py:678-682) */
  int32_t ray_type;       /* DYNRF_RAY_* */
} dynrf_field_cfg_t;

int dynrf_abi_version(void);
/* Human-readable name for a return code (static storage). */
const char* dynrf_strerror(int code);

/* ---- weight packing ------------------------------------------------------------------
 * nn.Linear weight [N][K] (+ optional bias [N]) -> wt [Kpad][Npad], bias [Npad].
 * Kpad = round_up(K,4), Npad = round_up(N,16) are the caller's (see dynrf_pad_k/_n). */
int dynrf_pad_k(int K);
int dynrf_pad_n(int N);
int dynrf_pack_linear(const float* w, const float* bias_or_null, int N, int K, float* wt, float* bias_out, void* stream);
/* inverse for gradients: dwt [Kpad][Npad] -> dw [N][K] (overwrite) */
int dynrf_unpack_linear_grad(const float* dwt, int N, int K, float* dw, void* stream);

/* ---- a1-a3: sample generation (renderer.py:147-170; models/tensorBase.py:487-499, 524-559)
 * zvec[S] is the batch-wide depth vector (one per call, shared by all rays, as in the reference).
 * jitter pointers are [S] (ndc), [Sin+1] / [Sout+1] (contract) uniforms or NULL for is_train=False. */
int dynrf_zvec_ndc(float near_, float far_, int S, const float* jitter, float* zvec, void* stream);
int dynrf_zvec_contract(float near_, float far_, int S, const float* jitter_inner, const float* jitter_outer,
                        float* zvec, void* stream);
/* rays [N][6] -> xyz [N][S][3], z_vals [N][S], valid [N][S] (uint8). */
int dynrf_sample_points(const dynrf_field_cfg_t* cfg, const float* rays, int N, int S, const float* zvec,
                        float* xyz, float* z_vals, uint8_t* valid, void* stream);

/* ---- a6+a8+a9(+a5): static density lookup + sigma->alpha + transmittance scan -----------
 * (TensorVMSplit.compute_densityfeature models/tensoRF.py:118-154; feature2density; raw2alpha
 *  models/tensorBase.py:22-34; the sigma/weight part of TensorBase.forward :717-774).
 * If xyz == NULL the samples are generated in-kernel from rays + zvec (nothing [N][S][3] is read).
 * If xyz != NULL then z_vals and valid must be given too ([N][S]).
 * Outputs: sigma, weight, dists (= delta * |d| * distance_scale) all [N][S];
 *          active_idx [N*S] int32 + active_count [1] int32 (caller zeroes the counter): flat
 *          indices n*S+s of samples with weight > weight_thres (the reference's app_mask);
 *          rgb_zero [N][S][3] or NULL: written 0 for samples NOT in the active list. */
int dynrf_static_sigma_weight_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const float* rays, int N, int S,
                                  const float* xyz, const float* z_vals, const uint8_t* valid, const float* zvec,
                                  float* sigma, float* weight, float* dists, int32_t* active_idx, int32_t* active_count,
                                  float* rgb_zero, void* stream);

/* sigma -> weight for a field whose sigma is already known (dynamic field): same outputs as above. */
int dynrf_alpha_weight_fwd(const dynrf_field_cfg_t* cfg, const float* rays, int N, int S, const float* sigma,
                           const float* z_vals, float* weight, float* dists, int32_t* active_idx,
                           int32_t* active_count, float* rgb_zero, void* stream);

/* ---- a7+a11+a12/a13: static appearance for the active samples -----------------------------
 * (compute_appfeature models/tensoRF.py:156-196 + basis_mat; heads models/tensorBase.py:81-145).
 * layers: basis (72->27), l1 (138|135 ->128), l2 (128->128), l3 (128|131 -> 3).  rgb [N][S][3]. */
int dynrf_static_app_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_layer_t* basis,
                         const dynrf_layer_t* l1, const dynrf_layer_t* l2, const dynrf_layer_t* l3, int head_mode,
                         const float* rays, const float* xyz, int N, int S, const int32_t* active_idx,
                         const int32_t* active_count, float* rgb, void* stream);

/* ---- a15: time embedding + coordinate warp (models/tensoRF.py:521-541) ---------------------
 * tfeat [N][48]: cols 0..16 = [t, PE(t,8)], cols 17..46 = layer2(relu(layer1(.))), col 47 = 0. */
int dynrf_time_embed_fwd(const float* ts, int N, const dynrf_layer_t* layer1, const dynrf_layer_t* layer2, float* tfeat,
                         void* stream);
/* xyz [P][3] (P = N*S samples of N rays, S per ray) -> xyz_prime [P][3] = unnorm(norm(xyz)) + delta. */
int dynrf_warp_fwd(const dynrf_field_cfg_t* cfg, const float* xyz, int N, int S, const float* tfeat,
                   const dynrf_layer_t* layer3, const dynrf_layer_t* layer4, const dynrf_layer_t* layer5,
                   float* xyz_prime, void* stream);

/* ---- a16+a17: dynamic density + blending (models/tensoRF.py:543-732) ----------------------
 * 3-scale lookups at normalize(xyz_prime), 152->64->1 MLPs; sigma = act(.), blending = sigmoid(.)
 * where valid, 0 elsewhere. */
int dynrf_dyn_sigma_blend_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                              const dynrf_layer_t* d1, const dynrf_layer_t* d2, const dynrf_layer_t* b1,
                              const dynrf_layer_t* b2, const float* xyz, const float* xyz_prime, const uint8_t* valid,
                              const float* tfeat, int N, int S, float* sigma, float* blend, void* stream);

/* ---- a18+a14: dynamic appearance for the active samples (models/tensoRF.py:734-811;
 *      MLPRender_Fea_late_view models/tensorBase.py:148-183). */
int dynrf_dyn_app_fwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_layer_t* basis,
                      const dynrf_layer_t* l1, const dynrf_layer_t* l2, const dynrf_layer_t* l3, const float* rays,
                      const float* xyz, const float* xyz_prime, const float* tfeat, int N, int S,
                      const int32_t* active_idx, const int32_t* active_count, float* rgb, float* featT_or_null, int ldf,
                      void* stream);
/* featT [32][ldf] (optional): the 27 basis outputs of active sample j at column j, feature-major, rows 27..31 zero;
 * consumed by dynrf_head_bwd_tc.  ldf >= round_up(N*S, 128). */

/* ---- a19: two-field compositing (renderer.raw2outputs, renderer.py:173-315) ----------------
 * white_bg: the reference's host-side coin (is_train && rand<0.5), decided by the caller.
 * out_ray [N][16]: rgb_full3, depth_full, acc_full, rgb_s3, depth_s, acc_s, rgb_d3, depth_d, acc_d, dynamicness.
 * w_full, w_s, w_d [N][S]. */
#define DYNRF_RAY_OUT 16
int dynrf_composite_fwd(int ray_type, int white_bg, const float* rgb_s, const float* sigma_s, const float* rgb_d,
                        const float* sigma_d, const float* dists, const float* blending, const float* z_vals,
                        const float* rays, int N, int S, float* out_ray, float* w_full, float* w_s, float* w_d,
                        void* stream);

/* ======================================================================================
 * Backward entry points (what autograd through grid_sampler_2d_backward / addmm / cumprod computes in
 * the reference; SURVEY.md 7 hard part 2).  All gradient outputs are ACCUMULATED into unless stated.
 * ====================================================================================== */

/* nn.Linear weight [N][K] -> wb [nrows_pad][Jpad], wb[n][j] = w[n][k0+j] (Jpad = round_up(k1-k0,16)): the operand of
 * the data-gradient GEMM dX = dY x W restricted to input columns k0..k1. */
int dynrf_pack_linear_bwd(const float* w, int N, int K, int k0, int k1, int nrows_pad, float* wb, void* stream);

/* weight = alpha*T backward (models/tensorBase.py:22-34): gsigma += , gdists += . */
int dynrf_alpha_weight_bwd(int N, int S, const float* sigma, const float* dists, const float* gweight, float* gsigma,
                           float* gdists, void* stream);

/* renderer.raw2outputs backward.  g_out_ray [N][16]; g_w_* [N][S] or NULL.  Outputs are OVERWRITTEN:
 * g_rgb_s/g_rgb_d [N][S][3], g_sigma_s/g_sigma_d/g_dists/g_blend [N][S], g_rays [N][6]. */
int dynrf_composite_bwd(int ray_type, int white_bg, const float* rgb_s, const float* sigma_s, const float* rgb_d,
                        const float* sigma_d, const float* dists, const float* blending, const float* z_vals,
                        const float* rays, int N, int S, const float* g_out_ray, const float* g_w_full, const float* g_w_s,
                        const float* g_w_d, float* g_rgb_s, float* g_sigma_s, float* g_rgb_d, float* g_sigma_d,
                        float* g_dists, float* g_blend, float* g_rays, void* stream);

/* static density lookup backward (grid_sampler_2d_backward of models/tensoRF.py:118-154): gsigma [N][S] is
 * d loss / d sigma; gxyz [N][S][3] (+=) may be NULL. */
int dynrf_static_density_bwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_grad_t* gdensity,
                             const float* xyz, const uint8_t* valid, const float* sigma, const float* gsigma, int N, int S,
                             float* gxyz, void* stream);

/* static appearance backward.  wb2 = pack_bwd(l2.w,0,128,128 rows); wb1 = pack_bwd(l1.w,0,K,128 rows);
 * wbB = pack_bwd(basis.w,0,72,32 rows).  grgb [N][S][3].  gxyz (+=) and gviewdir [N][3] (+=) may be NULL. */
int dynrf_static_app_bwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_vm_grad_t* gapp,
                         const dynrf_layer_t* basis, const dynrf_layer_t* l1, const dynrf_layer_t* l2, const dynrf_layer_t* l3,
                         int head_mode, const float* wb2, const float* wb1, const float* wbB, const float* rays,
                         const float* xyz, int N, int S, const int32_t* active_idx, const int32_t* active_count,
                         const float* grgb, const dynrf_layer_grad_t* gB, const dynrf_layer_grad_t* g1,
                         const dynrf_layer_grad_t* g2, const dynrf_layer_grad_t* g3, float* gxyz, float* gviewdir, void* stream);

/* dynamic appearance backward.  wb1f = pack_bwd(l1.w,0,27,128 rows); wbB = 3 x pack_bwd(basis.w, 72 si, 72 si+72, 32 rows)
 * concatenated.  gxp [N][S][3] += d loss / d xyz_prime. */
int dynrf_dyn_app_bwd(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* app, const dynrf_vm_grad_t* gapp,
                      const dynrf_layer_t* basis, const dynrf_layer_t* l1, const dynrf_layer_t* l2, const dynrf_layer_t* l3,
                      const float* wb2, const float* wb1f, const float* wbB, const float* rays, const float* xyz,
                      const float* xyz_prime, const float* tfeat, int N, int S, const int32_t* active_idx,
                      const int32_t* active_count, const float* grgb, const dynrf_layer_grad_t* gB,
                      const dynrf_layer_grad_t* g1, const dynrf_layer_grad_t* g2, const dynrf_layer_grad_t* g3, float* gxp,
                      void* stream);

/* dynamic density (mode 0) / blending (mode 1) MLP + lookup backward.  wb1f = pack_bwd(l1.w,0,72,64 rows).
 * gout [N][S] = d loss / d sigma | d blending.  gxp [N][S][3] += . */
int dynrf_dyn_mlp_bwd(const dynrf_field_cfg_t* cfg, int mode, const dynrf_vm_t* vm, const dynrf_vm_grad_t* gvm,
                      const dynrf_layer_t* l1, const dynrf_layer_t* l2, const float* wb1f, const float* xyz,
                      const float* xyz_prime, const uint8_t* valid, const float* tfeat, int N, int S, const float* gout,
                      const dynrf_layer_grad_t* g1, const dynrf_layer_grad_t* g2, float* gxp, void* stream);

/* coordinate warp backward.  wb4 = pack_bwd(layer4.w,0,64,64 rows); wb3t = pack_bwd(layer3.w,63,93,64 rows).
 * gxp [N][S][3] = total d loss / d xyz_prime.  gtout [N][32] += d loss / d t_out (per ray).
 * H1T / H2T [64][ld] (both or neither): relu(layer3), relu(layer4) as kept by dynrf_warp_fwd_tc; NULL = recompute them. */
int dynrf_warp_bwd(const dynrf_field_cfg_t* cfg, const float* xyz, int N, int S, const float* tfeat, const dynrf_layer_t* l3,
                   const dynrf_layer_t* l4, const dynrf_layer_t* l5, const float* wb4, const float* wb3t, const float* gxp,
                   const dynrf_layer_grad_t* g3, const dynrf_layer_grad_t* g4, const dynrf_layer_grad_t* g5, float* gtout,
                   const float* H1T, const float* H2T, int ld, void* stream);

/* time-embedding MLP backward: gtout [N][32] -> packed gradients of layer1 / layer2. */
int dynrf_time_embed_bwd(const float* ts, int N, const dynrf_layer_t* l1, const dynrf_layer_t* l2, const float* gtout,
                         const dynrf_layer_grad_t* g1, const dynrf_layer_grad_t* g2, void* stream);

/* ---- tensor-core (tcgen05 / TMEM / TMA) variants of the MLP-bearing forward stages ----------------
 * A "UMMA-packed" nn.Linear: weight [N][K] pre-split into tf32 hi/lo parts and laid out per 32-column chunk as
 * [hi: Npad x 32][lo: Npad x 32] in the K-major no-swizzle operand layout (dynrf_pack_linear_umma);
 * Npad = round_up(N,8), nchunks = ceil(K/32); wq holds nchunks*2*Npad*32 floats, bias Npad floats. */
typedef struct {
  const float* wq;
  const float* bias;
  int32_t K, N, Npad, nchunks;
} dynrf_umma_layer_t;
int dynrf_pack_linear_umma(const float* w, const float* bias_or_null, int N, int K, float* wq, float* bias_out, void* stream);
/* same contracts as dynrf_warp_fwd / dynrf_dyn_sigma_blend_fwd; 3-pass split-TF32 MMAs, fp32 accumulation in TMEM */
int dynrf_warp_fwd_tc(const dynrf_field_cfg_t* cfg, const float* xyz, int N, int S, const float* tfeat,
                      const dynrf_umma_layer_t* layer3, const dynrf_umma_layer_t* layer4, const dynrf_layer_t* layer5,
                      float* xyz_prime, float* H1T, float* H2T, int ld, void* stream);
/* v2 of the tensor-core coordinate warp (models/tensoRF.py:521-541): warp-specialised, two tiles in flight.
 * wq: dynrf_warp_fwd_wfloats() floats packed by dynrf_pack_warp_fwd from layer3.weight [64][93] and layer4.weight [64][64];
 * cst: dynrf_warp_fwd_cst_floats() floats = layer3.bias (64) | layer4.bias (64) | layer5.weight^T [64][3] | layer5.bias (3).
 * H1T / H2T [64][ld] (both or neither): relu(layer3) / relu(layer4) kept for dynrf_warp_bwd. */
int dynrf_warp_fwd_wfloats(void);
int dynrf_warp_fwd_cst_floats(void);
int dynrf_pack_warp_fwd(const float* w3, const float* w4, float* wq, void* stream);
int dynrf_warp_fwd_tc2(const dynrf_field_cfg_t* cfg, const float* wq, const float* cst, const float* xyz, int N, int S,
                       const float* tfeat, float* xyz_prime, float* H1T, float* H2T, int ld, void* stream);
int dynrf_dyn_sigma_blend_fwd_tc(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                                 const dynrf_umma_layer_t* d1, const dynrf_layer_t* d2, const dynrf_umma_layer_t* b1,
                                 const dynrf_layer_t* b2, const float* xyz, const float* xyz_prime, const uint8_t* valid,
                                 const float* tfeat, int N, int S, float* sigma, float* blend, void* stream);

/* v2 of the tensor-core density + blending stage: operands produced directly in the split K-major layout, the
 * shared position/time rows multiplied once against [density | blending] layer-1 weights (N = 128), 512 threads,
 * 3-deep operand ring + 2 TMA weight slots.  wq: dynrf_sb_umma_floats() floats built by dynrf_pack_sb_umma from the two
 * layer-1 weights [64][152]; b1cat [128] = [b1_density | b1_blending]; w2cat [128] = layer-2 weights; b2cat [2]. */
int dynrf_sb_umma_floats(void);
int dynrf_pack_sb_umma(const float* w1_density, const float* w1_blending, float* wq, void* stream);
int dynrf_dyn_sigma_blend_fwd_tc2(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                                  const float* wq, const float* b1cat, const float* w2cat, const float* b2cat,
                                  const float* xyz, const float* xyz_prime, const uint8_t* valid, const float* tfeat, int N,
                                  int S, float* sigma, float* blend, float* XT_or_null, float* HT_or_null, int ld, void* stream);
/* XT_or_null / HT_or_null (both or neither): keep the layer-1 inputs X^T [224][ld] (80 position/time rows in engine order,
 * 72 density features, 72 blending features) and the hidden activations H^T [128][ld] (density | blending, relu'd, zero on
 * invalid samples) for dynrf_dyn_mlp_bwd_tc; ld >= round_up(N*S,128), multiple of 4. */

/* ---- tensor-core backward of the appearance head on the active samples -----------------------------------
 * dynamic = 1: MLPRender_Fea_late_view of the dynamic field; dynamic = 0: MLPRender_Fea / MLPRender_Fea_TimeEmbedding of the
 *   static field (reference models/tensorBase.py MLPRender_Fea*, gradients reached through loss.backward()).
 * dynrf_pack_head_bwd: layer-1 [128][107 | 138 | 135] / layer-2 [128][128] weights -> dynrf_head_bwd_wfloats(dynamic) packed
 *   floats + the column map [dynrf_head_bwd_xrows(dynamic)] of X1T rows to nn.Linear input columns (-1: padding row, -2: the
 *   ones row whose gradient column is d bias).
 * dynrf_head_bwd_tc: cst = [b1(128) | b2(128) | W3[:, :128]^T (128x3) | W3[:, 128:131]^T (3x3, zeros for DYNRF_HEAD_FEA) | b3(3)];
 *   featT [32][ld] = basis outputs saved by dynrf_app_fwd_tc; writes feature-major [rows][ld] tensors X1T[xrows] H1T[128]
 *   DH2T[128] DH1T[128] DFT[32] (column j = j-th active sample) for the weight-gradient GEMMs and accumulates gw3h [3][128]
 *   (d W3[:, :128]), gw3v [3][3] (d W3[:, 128+k]), gb3 [3], gb2 [128] and (static MLP_Fea head, optional) gviewdir [N][3].
 *   ld >= round_up(N*S,128), multiple of 4.
 * dynrf_basis_bwd: DFT -> packed basis gradient [K][32] (+=), plane / line gradients (+=), d loss / d coords (+=). */
int dynrf_head_bwd_wfloats(int dynamic);
int dynrf_head_bwd_cst_floats(void);
int dynrf_head_bwd_xrows(int dynamic);
int dynrf_pack_head_bwd(int dynamic, int head, const float* w1, const float* w2, float* wq, int32_t* colmap, void* stream);
int dynrf_head_bwd_tc(const dynrf_field_cfg_t* cfg, int dynamic, const float* wq, const float* cst, const float* rays, const float* xyz,
                      const float* tfeat, const float* featT, const float* grgb, const int32_t* active_idx,
                      const int32_t* active_count, int N, int S, int ld, float* X1T, float* H1T, float* DH2T, float* DH1T,
                      float* DFT, float* gw3h, float* gw3v, float* gb3, float* gb2, float* gviewdir, void* stream);
int dynrf_basis_bwd(const dynrf_field_cfg_t* cfg, int dynamic, const dynrf_vm_t* app, const dynrf_vm_grad_t* gapp, const float* wbB,
                    const float* coords, const float* DFT, int ld, int N, int S, const int32_t* active_idx,
                    const int32_t* active_count, float* gbasis_wt, float* gcoord, void* stream);

/* ---- tensor-core forward of the appearance stage on the active samples (static: a7+a11+a12/a13, reference
 * models/tensorBase.py compute_appfeature + basis_mat + MLPRender_Fea / MLPRender_Fea_TimeEmbedding; dynamic: a18+a14,
 * compute_appfeature at strides 1,2,4 + MLPRender_Fea_late_view).  Same results as dynrf_static_app_fwd / dynrf_dyn_app_fwd
 * (3-pass split TF32, fp32 accumulate).
 * dynrf_pack_app_fwd: basis_mat weight [27][72 or 216], head layer-1 weight [128][138 | 135 | 107], layer-2 weight
 *   [128][128] -> dynrf_app_fwd_wfloats(dynamic) packed floats.  head = DYNRF_HEAD_*.
 * dynrf_app_fwd_tc: cst [dynrf_app_fwd_cst_floats()] = [b1(128) | b2(128) | W3[:, :128]^T (128x3) | W3[:, 128:131]^T (3x3, zeros
 *   for DYNRF_HEAD_FEA) | b3(3)]; coords = lookup coordinates (xyz for the static field, xyz_prime for the dynamic one);
 *   tfeat only for the dynamic field; featT (optional) [32][ld] receives the basis output of active sample j in column j. */
int dynrf_app_fwd_wfloats(int dynamic);
int dynrf_app_fwd_cst_floats(void);
int dynrf_pack_app_fwd(int dynamic, int head, const float* basis_w, const float* w1, const float* w2, float* wq, void* stream);
int dynrf_app_fwd_tc(const dynrf_field_cfg_t* cfg, int dynamic, const dynrf_vm_t* app, const float* wq, const float* cst,
                     const float* rays, const float* xyz, const float* coords, const float* tfeat, int N, int S,
                     const int32_t* active_idx, const int32_t* active_count, float* rgb, float* featT, int ld, void* stream);

/* ---- tensor-core backward of the dynamic density + blending stage (reference: TensorBase.compute_densityfeature /
 * blending MLP backward reached through loss.backward() from train.py; forward at models/tensorBase.py density_layer*,
 * blending_layer*) -- both tables in one launch.
 * dynrf_pack_sb_bwd: layer-1 weights [64][152] of the two MLPs -> dynrf_sb_bwd_wfloats() packed floats + colmap80 (row r of the
 *   position/time block of XT -> nn.Linear input column).
 * dynrf_dyn_mlp_bwd_tc: cst = [b1 density(64) | b1 blending(64) | w2 density(64) | w2 blending(64) | b2 density | b2 blending];
 *   gsigma / gblend [N*S] (either may be NULL); writes XT [225][ld] (80 position/time rows in engine order, 72 density features,
 *   72 blending features, ones row) and DHT [128][ld] (d pre-activation of layer 1, density | blending) for dynrf_wgrad_gemm;
 *   accumulates plane / line gradients, gxp [N*S][3], gw2 [128], gb1 [128], gb2 [2].  ld >= round_up(N*S,128), multiple of 4. */
int dynrf_sb_bwd_wfloats(void);
int dynrf_pack_sb_bwd(const float* w1_density, const float* w1_blending, float* wq, int32_t* colmap80, void* stream);
int dynrf_dyn_mlp_bwd_tc(const dynrf_field_cfg_t* cfg, const dynrf_vm_t* density, const dynrf_vm_t* blending,
                         const dynrf_vm_grad_t* gdensity, const dynrf_vm_grad_t* gblending, const float* wq, const float* cst,
                         const float* xyz, const float* xyz_prime, const uint8_t* valid, const float* tfeat, const float* gsigma,
                         const float* gblend, int N, int S, int ld, float* XT, const float* HT_saved_or_null, float* DHT, float* gxp,
                         float* gw2, float* gb1, float* gb2, void* stream);
/* HT_saved_or_null != NULL: X^T and H^T were kept by dynrf_dyn_sigma_blend_fwd_tc2 -- layer 1 is not recomputed and XT is only
 * read by the weight-gradient GEMM afterwards; NULL: layer 1 is recomputed and X^T written here. */

/* Weight-gradient GEMM on tensor cores (3-pass split TF32, fp32 TMEM accumulation):
 *   dw[n * dw_ld + k] += sum_{p < P} XT[k * ldx + p] * DYT[n * ldy + p],  k < K <= 256, n < N <= 128,
 * P = min(*count_ptr, Pmax) (count_ptr may be NULL).  XT / DYT: feature-major activations / output gradients written by the
 * fused backward kernels; 16-byte aligned, ldx / ldy multiples of 4. */
int dynrf_wgrad_gemm(const float* XT, int ldx, int K, const float* DYT, int ldy, int N, int Pmax, const int32_t* count_ptr,
                     float* dw, int dw_ld, void* stream);

/* ---- rows next to the ray-march path -----------------------------------------------------------
 * a4: TensorBase.sample_ray (models/tensorBase.py:501-522): aabb entry distance + fixed-step marching.
 *     jitter: [N] uniforms (is_train) or NULL.  Outputs xyz [N][S][3], z_vals [N][S] (per ray), valid [N][S]. */
int dynrf_sample_ray(const dynrf_field_cfg_t* cfg, const float* rays, int N, int S, float step_size, const float* jitter,
                     float* xyz, float* z_vals, uint8_t* valid, void* stream);
/* f1: fused ray generation (train.py:96-103, 1062-1077; dataLoader/ray_utils.py:53-90, 115-140): flat pixel ids [B] (int64,
 *     id = view*H*W + row*W + col) + camera-to-world matrices mtx [F][3][4] (camera.pose_to_mtx of the 9-D poses) + focal [1]
 *     (device scalar) -> rays [B][6]; ndc != 0 applies ndc_rays_blender2 with `near`.  The backward accumulates into
 *     gmtx [F][3][4] and gfocal [1] (caller zeroes both).
 *     Instead of ids, explicit sub-pixel coordinates uv [B][2] = (x, y) and camera indices view [B] (int64) may be given
 *     (ids == NULL): the flow-displaced rays of train.py:1433-1455. */
int dynrf_raygen_fwd(const int64_t* ids_or_null, const float* uv_or_null, const int64_t* view_or_null, int B, const float* mtx, int F,
                     const float* focal, int H, int W, int ndc, float near_, float* rays, void* stream);
int dynrf_raygen_bwd(const int64_t* ids_or_null, const float* uv_or_null, const int64_t* view_or_null, int B, const float* mtx, int F,
                     const float* focal, int H, int W, int ndc, float near_, const float* grays, float* gmtx, float* gfocal,
                     void* stream);
/* f2: the S-reduction of renderer.render_3d_point (renderer.py:1335-1338): pm[n] = sum_s w[n][s] * moved[n][s],
 *     acc[n] = sum_s w[n][s] with moved = pts (+ sf, the scene flow) (clamped to [lo,hi] when clamp != 0, train.py:1374-1378).
 *     Backward: gw [N][S] = gpm[n] . moved + gacc[n] (gacc may be NULL), gmoved [N][S][3] = w * gpm[n] (0 where clamped). */
int dynrf_weighted_point_fwd(const float* w, const float* pts, const float* sf_or_null, int clamp, float lo, float hi, int N, int S,
                             float* pm, float* acc, void* stream);
int dynrf_weighted_point_bwd(const float* w, const float* pts, const float* sf_or_null, int clamp, float lo, float hi, int N, int S,
                             const float* gpm, const float* gacc_or_null, float* gw_or_null, float* gmoved_or_null, void* stream);
/* f3: total variation (utils.TVLoss, utils.py:157-181) of up to 6 channel-last tensors [H][W][C] in one launch:
 *     sums[2t] += sum (x[h+1]-x[h])^2, sums[2t+1] += sum (x[w+1]-x[w])^2 (double, caller zeroes); if g[t] != NULL it is
 *     OVERWRITTEN with ch[t] * d(sums[2t])/dx + cw[t] * d(sums[2t+1])/dx.  Host pointers to arrays of length n. */
int dynrf_tv_fused(int n, const float* const* x, float* const* g_or_null, const int32_t* C, const int32_t* H, const int32_t* W,
                   const float* ch, const float* cw, double* sums, void* stream);
/* f3: density_L1 / blending_L1 (models/tensoRF.py:80-98, 378-416) without the dense C x Gy x Gx x Gz product:
 *     sum_out[0] += sum_{x,y,z} |act(sum_c plane_i (x) line_i)| (double, caller zeroes; the caller divides by Gx*Gy*Gz);
 *     if g != NULL the gradient of that SUM is accumulated into g (caller zeroes).  C must be [16,4,4]. */
int dynrf_dense_l1(const dynrf_vm_t* vm, int act, float shift, double* sum_out, const dynrf_vm_grad_t* g_or_null, void* stream);


/* ---- optimiser: one-launch Adam over the flat gradient buffer of the train step (torch.optim.Adam semantics, no weight
 * decay / amsgrad; train.py:934, 2313-2325).  param_ptr [T]: device array of the parameters' storage pointers (uint64);
 * offset [2T] (int64): [start, end) of every parameter in the flat grad / m / v buffers (starts may be padded for
 * alignment); blk0 [T+1] (int32) with
 * blk0[t+1] - blk0[t] = dynrf_adam_blocks(numel_t); lr [T]; step [2]: step[0] float update count (incremented by the call),
 * step[1] uint32 scratch that the caller zero-initialises once. */
int dynrf_adam_blocks(long long numel);
int dynrf_adam_flat(const uint64_t* param_ptr, const int64_t* offset, const int32_t* blk0, int T, int total_blocks, const float* grad,
                    float* m, float* v, float* step, const float* lr, float beta1, float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNRF_B200_H */
