/* brdfnerf_b200 — C ABI of the B200-native (sm_100a) BRDF-NeRF ray-rendering hot path.
 *
 * Drop-in boundary for the reference's `rendering.render_rays` (rendering.py:168-334) and the
 * `models/spsbrdfnerf.py` forward/backward.  The reference is pure Python/PyTorch and has no FFI of
 * its own; each entry point below names the reference function (file:line under /root/reference)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory
 *     (inputs, outputs, workspaces); the library never allocates or frees caller tensors;
 *   - all work is enqueued on `stream`; no entry point synchronises or reads device memory on the
 *     host, so a whole render / training step is CUDA-graph capturable;
 *   - return value: BN_OK (0) or a negative bn_status; `bn_last_error()` returns a thread-local
 *     human readable message for the last failure on the calling thread;
 *   - the library refuses to run on anything but compute capability 10.x (no fallback path).
 */
#ifndef BRDFNERF_B200_H_
#define BRDFNERF_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define BN_ABI_VERSION 1

typedef enum bn_status {
  BN_OK = 0,
  BN_ERR_ARG = -1,      /* invalid argument (null pointer, size out of range, unsupported config) */
  BN_ERR_CUDA = -2,     /* a CUDA runtime / driver call failed */
  BN_ERR_DEVICE = -3,   /* not a compute-capability 10.x device */
  BN_ERR_STATE = -4     /* handle used before weights were set, workspace too small, ... */
} bn_status;

int bn_abi_version(void);
const char* bn_last_error(void);
/* BN_OK iff device `device` is sm_100-class. Host-only query. */
int bn_device_check(int device);

/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
unsigned long long bn_launch_count(void);
/* per-launch CUDA-event timing of the GEMM family (bench.py roofline leg); off by default */
int bn_profile_enable(int on);
/* HOST pointers: per kind (0 = TN fwd/dgrad GEMM, 1 = NT wgrad GEMM, 2 = fused PE + trunk chain kernel) launch count,
 * milliseconds, flops */
int bn_profile_collect(int n_kinds, long long* count, double* ms, double* work);

/* ------------------------------------------------------------------ K-A  sample generators */

/* get_z_vals (rendering.py:149-166, perturb == 1): z = lower + (upper-lower)*u over the strata of
 * near*(1-t)+far*t.  near/far are read as near_p[r*stride], far_p[r*stride] (pass &rays[0][6],
 * &rays[0][7], 11 for a ray record).  t_vals = torch.linspace(0,1,S) (S floats), u (N,S) uniform
 * draws, z_out (N,S).  Bit-exact against the reference's torch-CPU result. */
int bn_sample_stratified(const float* near_p, const float* far_p, int stride,
                         const float* t_vals, const float* u, float* z_out,
                         int n_rays, int n_samples, cudaStream_t stream);

/* GenerateGuidedSamples (rendering.py:132-147) = calc_depth_std (train_utils.py:35-39) +
 * compute_samples_around_depth (rendering.py:116-130) + sample_3sigma_asym (:76-91) +
 * sample_3sigma (:54-74) + sample_pdf (:13-52) + sort, one warp per ray.
 *   z1, weights (N,S1), depth (N): pass-1 results; t_vals (G) and gauss_w (G-1): host-built tables
 *   (torch.linspace / Gaussian window of rendering.py:63,68-69); u_pred (N,G) uniform draws;
 *   near0/far0: device pointers to the chunk clamp scalars (reference uses ray 0's near/far);
 *   valid_depth (N) int64 or NULL: rows > 0 sample around gt_depth[r*gt_depth_stride] +-
 *   d_range*gt_std[r] with draws u_gt[r] instead (mask form of the reference's np.where scatter);
 *   z2_out (N,G) ascending; std_out (N) optional sampling std of pass 1. */
int bn_sample_guided(const float* z1, const float* depth, const float* weights,
                     const float* t_vals, const float* gauss_w, const float* u_pred,
                     const float* near0, const float* far0, float d_range,
                     const int64_t* valid_depth, const float* gt_depth, int gt_depth_stride,
                     const float* gt_std, const float* u_gt,
                     float* z2_out, float* std_out,
                     int n_rays, int n_samples, int n_guided, cudaStream_t stream);

/* cat + sort with indices (rendering.py:271-272): z_out (N,S1+G) ascending, idx_out int64 source
 * positions in [z1|z2] (stable on ties), unsort_out = [z1|z2] (both optional). */
int bn_merge_samples(const float* z1, const float* z2, float* z_out, int64_t* idx_out,
                     float* unsort_out, int n_rays, int n_samples, int n_guided, cudaStream_t stream);

/* Everything between the two network passes of render_rays in ONE launch (rendering.py:262-273): compositing of the
 * stratified densities (cal_weight, models/spsbrdfnerf.py:50-69 -> weights1 (N,S1), depth1 (N)), the depth-guided samples
 * (bn_sample_guided -> z2 (N,G), std1 (N) optional) and the merge (bn_merge_samples -> z_out, idx_out, unsort_out (N,S1+G);
 * the last two optional).  Arguments as in bn_composite_sigma / bn_sample_guided / bn_merge_samples; results bit-identical
 * to calling the three. */
int bn_coarse_to_fine(const float* z1, const float* sigma1, const float* noise1, float noise_std,
                      const float* t_vals, const float* gauss_w, const float* u_pred, const float* near0, const float* far0,
                      float d_range, const int64_t* valid_depth, const float* gt_depth, int gt_depth_stride,
                      const float* gt_std, const float* u_gt, float* weights1, float* depth1, float* std1, float* z2,
                      float* z_out, int64_t* idx_out, float* unsort_out, int n_rays, int n_samples, int n_guided,
                      cudaStream_t stream);

/* Applies sort_idx (rendering.py:271-273) to per-point rows of `pitch` floats.  The MLP may evaluate the
 * points of a call in generation order - all stratified samples ([N][S1] rows), then all guided samples
 * ([N][G] rows) - so that the stratified points' trunk is evaluated once for the density pass and the full
 * pass (bn_mlp_trunk_forward); compositing wants depth order.  scatter = 0: dst (N,S1+G,pitch) depth-ordered
 * <- src in MLP row order; scatter = 1: the inverse (gradients).  src != dst. */
int bn_permute_samples(const float* src, const int64_t* sort_idx, float* dst, int n_rays, int n_samples,
                       int n_guided, int pitch, int scatter, cudaStream_t stream);

/* per-row ascending sort (rendering.py:263). */
int bn_sort_rows(const float* in, float* out, int n_rays, int n, cudaStream_t stream);

/* ------------------------------------------------------------------ K-C  compositing */

/* cal_weight (spsbrdfnerf.py:50-69) for the sigma-only pass: alpha/trans optional outputs,
 * weights (N,S), depth (N), std_out optional = calc_depth_std. noise may be NULL. */
int bn_composite_sigma(const float* z, const float* sigma, const float* noise, float noise_std,
                       float* alpha, float* trans, float* weights, float* depth, float* std_out,
                       int n_rays, int n_samples, cudaStream_t stream);

/* cal_weight + every weighted accumulation of `inference` (spsbrdfnerf.py:196-199,242,270-275,
 * 292,314-317,326-338).  packed (N,S,C) is the MLP's packed per-sample output in the reference's
 * channel order [albedo(3), sigma, normal_an(3)?, normal_lr(3)?, brdf params...]; acc (N,C) receives
 * sum_s w*packed (channel 3 is meaningless), wsum (N) = sum_s w.  irr (N,S) optional per-sample
 * irradiance scalar (sun visibility) -> acc_irr (N,3) = sum_s w*irr*albedo.
 * sort_idx (N,S) int64, optional: `packed` is then in the MLP's generation order (all stratified samples [N][n_stratified]
 * rows, then all guided samples [N][S - n_stratified] rows, see bn_permute_samples) and sample s of ray r is read from the
 * row sort_idx[r][s] points at (rendering.py:271-273 applied on load instead of by a separate pass). */
int bn_composite_forward(const float* z, const float* packed, int n_channels, int sigma_channel,
                         const float* noise, float noise_std, const float* irr,
                         float* alpha, float* trans, float* weights,
                         float* depth, float* wsum, float* acc, float* acc_irr,
                         int n_rays, int n_samples, const int64_t* sort_idx, int n_stratified, cudaStream_t stream);

/* backward of bn_composite_forward: given grads w.r.t. acc / acc_irr / depth / wsum (per ray),
 * weights (per sample, optional) and optionally direct grads w.r.t. packed, writes g_packed
 * (N,S,C) (channel 3 = d loss / d sigma).  With sort_idx (as in bn_composite_forward) packed is read from, and g_packed
 * written to, the MLP's row order; g_packed_direct stays in depth order. */
int bn_composite_backward(const float* z, const float* packed, int n_channels, int sigma_channel,
                          const float* noise, float noise_std, const float* irr,
                          const float* alpha, const float* trans, const float* weights,
                          const float* g_acc, const float* g_acc_irr, const float* g_depth,
                          const float* g_wsum, const float* g_weights, const float* g_packed_direct,
                          float* g_packed, int n_rays, int n_samples, const int64_t* sort_idx, int n_stratified,
                          cudaStream_t stream);

/* K-C of the Lambertian training step in one launch, both directions (csrc/render_loss.cu): compositing of the 128-sample ray
 * (cal_weight + weighted sums, models/spsbrdfnerf.py:50-69,196-199), Lambertian colour with rgb_padding (:281-283,459),
 * SNerfLoss + DepthLoss(subset) (metrics.py:39-61,82-161) and the backward of all three down to g_rows, the gradient of every
 * packed MLP row.  rows / g_rows: [N*128, 4] (albedo rgb, sigma) in the MLP's generation order when sort_idx is given (see
 * bn_composite_forward), else in depth order.  Same arithmetic as bn_composite_forward -> bn_shade_rays_forward ->
 * bn_loss_color_depth -> bn_shade_rays_backward -> bn_composite_backward; loss (1) is zeroed by the call; rgb (N,3) and depth (N)
 * are optional per-ray outputs.  No normals, no BRDF, irradiance 1, noise_std 0, n_samples == 128. */
int bn_lambertian_render_loss(const float* z, const float* rows, const int64_t* sort_idx, int n_stratified,
                              const float* target_rgb, const int64_t* valid_depth, const float* target_depth,
                              const float* target_weight, int td_stride, const float* target_std, float lambda_rgb,
                              float lambda_ds, int use_all_depth, float* loss, float* g_rows, float* rgb, float* depth,
                              int n_rays, int n_samples, cudaStream_t stream);

/* ------------------------------------------------------------------ K-C  shading / BRDF */

enum { BN_BRDF_NONE = 0, BN_BRDF_MICROFACET = 1, BN_BRDF_RPV = 2, BN_BRDF_HAPKE = 3 };
enum { BN_IRR_ONES = 0, BN_IRR_COS = 1, BN_IRR_SUNVIS = 2 };

/* Which channels of a packed row mean what, and which BRDF is active for this call
 * (mirrors the switches `inference` reads: spsbrdfnerf.py:104-116,144-190,286-346). */
typedef struct bn_shade_cfg {
  int n_channels;      /* C of packed / acc rows */
  int normal_ch;       /* first channel of the normal used for shading (learned wins), -1 = none */
  int param_ch;        /* first BRDF-parameter channel, -1 = none */
  int brdf_ch;         /* MultiBRDF: first of 3 per-sample brdf channels appended to the row, else -1 */
  int brdf_type;       /* BN_BRDF_* active for this call (NONE when apply_brdf is off) */
  int funcM, funcF, funcH;                         /* RPV factors; funcH == 2: rhoc = albedo */
  int hapke_b, hapke_c, hapke_theta, shell_hapke;  /* Hapke parameters present in the row */
  int multi_brdf;      /* one BRDF per sample instead of one per ray */
  int irr_mode;        /* BN_IRR_* */
  float hpk_scl, fresnel_f0;
} bn_shade_cfg;

/* Per-ray epilogue of `inference` (spsbrdfnerf.py:198-199,241-275,284-357): from the accumulations
 * of bn_composite_forward to albedo_accu, rgb, normal_s, nr_vw, nr_sun, hpk_scl and the per-ray BRDF
 * (calc_angles + RPV.py / Hapke.py / microfacet.py).  Optional outputs may be NULL.
 * brdf (N,3); aux (N,3,8) per colour channel: RPV [M1,G,H,ci,cv], Hapke [P,Hi,Hv,ci,cv,S],
 * microfacet [glossy,f,g,d,l.n,v.n,n.h]. irr_last (N): irradiance of the last sample (sun-vis). */
int bn_shade_rays_forward(const bn_shade_cfg* cfg, const float* rays, const float* acc,
                          const float* wsum, const float* acc_irr, const float* irr_last,
                          float* rgb, float* albedo_accu, float* normal_s, float* nr_vw,
                          float* nr_sun, float* hpk_scl, float* brdf, float* aux,
                          int n_rays, cudaStream_t stream);

/* d loss / d rgb (N,3)  ->  d loss / d {acc (N,C), wsum (N), acc_irr (N,4)}. */
int bn_shade_rays_backward(const bn_shade_cfg* cfg, const float* rays, const float* acc,
                           const float* wsum, const float* acc_irr, const float* irr_last,
                           const float* g_rgb, float* g_acc, float* g_wsum, float* g_acc_irr,
                           int n_rays, cudaStream_t stream);

/* MultiBRDF (spsbrdfnerf.py:289-290,297-307,343-344): one BRDF per sample, written into channels
 * brdf_ch..brdf_ch+2 of the packed rows; backward folds d/d brdf into the other channels' grads. */
int bn_brdf_points_forward(const bn_shade_cfg* cfg, const float* rays, float* packed, float* aux,
                           int n_rays, int n_samples, cudaStream_t stream);
int bn_brdf_points_backward(const bn_shade_cfg* cfg, const float* rays, const float* packed,
                            float* g_packed, int n_rays, int n_samples, cudaStream_t stream);

/* ------------------------------------------------------------------ losses */

/* SNerfLoss (metrics.py:39-61, lambda_sc == 0) + DepthLoss(subset=True, GNLL=False) (metrics.py:82-161) of one
 * batch, fused with their gradients w.r.t. the rendered rgb (N,3) and depth (N):
 *   loss = lambda_rgb * mean((rgb - target)^2) + (lambda_ds / 3 / N) * sum_sel w* (d - d*)^2
 * where a ray is selected when valid_depth > 0 and (use_all_depth or |d - d*| > std* or calc_depth_std(z,d,w) > std*).
 * valid_depth == NULL drops the depth term.  target_depth / target_weight are read as ptr[r * td_stride]
 * (target_weight NULL = 1).  loss (1 float) is overwritten; g_rgb (N,3) and g_depth (N) receive d loss / d . */
int bn_loss_color_depth(const float* rgb, const float* target_rgb, const float* depth, const float* z,
                        const float* weights, const int64_t* valid_depth, const float* target_depth,
                        const float* target_weight, int td_stride, const float* target_std,
                        float lambda_rgb, float lambda_ds, int use_all_depth,
                        float* loss, float* g_rgb, float* g_depth, int n_rays, int n_samples, cudaStream_t stream);

/* Optional regularisers of the BRDF stage (main.py:269-299), fused with their gradients:
 *   NormalRegLoss   (metrics.py:179-216)  lambda_nr * sum_{r,s} w_rs min(0, n_rs . v_r)^2, v_r = -d_r, once per normal
 *                   kind (analytic: channel normal_an_ch, learned: normal_lr_ch; channel < 0 or lambda == 0 = off);
 *   HardSurfaceLoss (metrics.py:263-290, train_utils.py:38-39)  lambda_hs / N * sum_r sum_s (z_rs - depth_r)^2 w_rs.
 * loss (1 float) and g_depth (N) are ACCUMULATED (call after bn_loss_color_depth, or zero them first); g_weights (N,S)
 * and g_packed (N,S,pitch: normal channels, zeros elsewhere; only touched when a normal term is on) are overwritten and
 * are the g_weights / g_packed_direct operands of bn_composite_backward.  bad_count (2 floats, nullable, ACCUMULATED):
 * number of samples with n . v < 0 per normal kind (the reference's `perc_ng_nr` numerator, metrics.py:196-198). */
int bn_loss_regularizers(const float* weights, const float* z, const float* depth, const float* packed, int pitch,
                         int normal_an_ch, float lambda_nr_an, int normal_lr_ch, float lambda_nr_lr,
                         const float* rays, float lambda_hs, float* loss, float* g_weights, float* g_packed,
                         float* g_depth, float* bad_count, int n_rays, int n_samples, cudaStream_t stream);

/* Device-side NaN counter: *counter (DEVICE int, ACCUMULATED: zero it first) += number of NaNs in x[0..n).  Stands in for
 * the host-synchronising torch.isnan(x).sum() of train_utils.check_nan (train_utils.py:61-78; three calls per render_rays,
 * rendering.py:121-123): nothing here synchronises, the caller reads the counters when it wants to report. */
int bn_count_nan(const float* x, long long n, int* counter, cudaStream_t stream);

/* ------------------------------------------------------------------ K-B  PE + SIREN MLP */

enum { BN_PREC_FP32 = 0,   /* CUDA-core fp32 everywhere: parity mode (<= 1e-3 against the fp32 oracle) */
       BN_PREC_BF16 = 1 }; /* tcgen05 bf16 operands / fp32 TMEM accumulation: throughput mode */

/* linear-layer ids used to address weights/biases inside the caller's flat fp32 parameter buffer
 * (state_dict names of the reference: spsbrdfnerf.py:513-613) */
enum {
  BN_LIN_TRUNK0 = 0,        /* fc_net.{2l}           l = 0..15 */
  BN_LIN_SIGMA = 16,        /* sigma_from_xyz.0 */
  BN_LIN_FEATS = 17,        /* feats_from_xyz */
  BN_LIN_RGB0 = 18,         /* rgb_from_xyzdir.0 */
  BN_LIN_RGB2 = 19,         /* rgb_from_xyzdir.2 */
  BN_LIN_GRAD = 20,         /* grad_from_xyz (learned normal) */
  BN_LIN_HEAD0 = 21,        /* head h: 21+2h = {name}.0, 22+2h = {name}.2 */
  BN_NUM_LINEAR = 37
};
/* optional heads on the feature vector: the BRDF heads in the reference's output-channel order, then the transient-
 * uncertainty head beta_from_xyz (spsbrdfnerf.py:571-575), whose first layer also reads the ray's time embedding */
enum { BN_HEAD_ROUGH = 0, BN_HEAD_K = 1, BN_HEAD_THETA_RPV = 2, BN_HEAD_RHOC = 3,
       BN_HEAD_B = 4, BN_HEAD_C = 5, BN_HEAD_THETA = 6, BN_HEAD_BETA = 7, BN_NUM_HEADS = 8 };

typedef struct bn_mlp_cfg {
  int feat;               /* fc_feat (multiple of 64; 512 in the reference recipe) */
  int layers;             /* fc_layers (<= 16) */
  int skip_layer;         /* layer whose input is [PE | h], -1 = none (reference: 4) */
  int n_freq_xyz;         /* positional-encoding frequencies, 0 = raw xyz (no --mapping) */
  int normal_lr;          /* grad_from_xyz head exists */
  int viewdir;            /* --input_viewdir: the colour head reads [features | Mapping(ray direction)] (spsbrdfnerf.py:689-690) */
  int n_freq_dir;         /* frequencies of the direction encoding (mapping_sizes[1] = 4), 0 = raw direction (no --mapping) */
  int t_dims;             /* width of the per-ray time embedding read by the beta head (args.t_embbeding_tau, <= 32); 0 without beta */
  int head_dim[BN_NUM_HEADS];   /* output width of each head that exists (0 = absent) */
  int precision;          /* BN_PREC_* */
  int64_t w_off[BN_NUM_LINEAR]; /* element offset of each weight / bias in the flat buffer, -1 = absent */
  int64_t b_off[BN_NUM_LINEAR];
  int64_t n_params;       /* total elements of the flat buffer */
} bn_mlp_cfg;

typedef struct bn_mlp bn_mlp;   /* opaque: config + packed (bf16 / transposed) weight copies */

/* flags of one MLP evaluation (what `SpSBRDFNeRF.forward` is asked for, spsbrdfnerf.py:662) */
enum {
  BN_MLP_SIGMA_ONLY = 1,   /* sigma_only=True: out is (P) densities */
  BN_MLP_TRAIN = 2,        /* keep activations for bn_mlp_backward */
  BN_MLP_NORMAL_AN = 4,    /* nr_an_on: analytic normal -l2n(d sigma / d x) */
  BN_MLP_NORMAL_LR = 8,    /* nr_lr_on */
  BN_MLP_ROUGH = 16,       /* evaluate the roughness head (microfacet, apply_brdf) */
  BN_MLP_RPV = 32,         /* evaluate k / theta_rpv / rhoc heads that exist */
  BN_MLP_HAPKE = 64,       /* evaluate b / c heads that exist */
  BN_MLP_HAPKE_THETA = 128,/* ... and the Hapke theta head (apply_theta) */
  BN_MLP_BETA = 256        /* model.beta: channel 4 = softplus(beta_from_xyz([features | t])) (spsbrdfnerf.py:708-711); the time
                            * embedding of the rows must have been written with bn_mlp_write_t */
};

int bn_mlp_create(const bn_mlp_cfg* cfg, bn_mlp** out);
void bn_mlp_destroy(bn_mlp* h);
/* refresh the packed weight copies from the flat fp32 master (after every optimizer step) */
int bn_mlp_sync_weights(bn_mlp* h, const float* params, cudaStream_t stream);
/* number of packed output channels for `flags` (4 + normals + BRDF parameters) */
int bn_mlp_out_channels(const bn_mlp* h, int flags);
size_t bn_mlp_workspace_bytes(const bn_mlp* h, int64_t n_points, int flags);

/* PE (nerf.py:53-70) + trunk (spsbrdfnerf.py:636-646) + heads (:682-755) for the points
 * x = origin[r] + dir[r] * z[r][s].  origins/dirs are read as origins[r*o_stride + {0,1,2}].
 * out: (P) sigma when BN_MLP_SIGMA_ONLY, else packed rows (P, out_pitch >= out_channels) in the
 * reference's channel order. */
int bn_mlp_forward(bn_mlp* h, const float* params, const float* origins, int o_stride,
                   const float* dirs, int d_stride, const float* z, int n_rays, int n_samples,
                   int flags, float* out, int out_pitch, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream);

/* The same forward in two steps, for callers that evaluate the points of a call in several batches
 * (render_rays: stratified samples first, guided samples once the density of the first batch is known,
 * rendering.py:225-274) without evaluating any point twice:
 *   bn_mlp_trunk_forward  PE + trunk of n_rays*n_samples points; their activations go to rows
 *                         [row0, row0 + n_rays*n_samples) of a workspace sized for total_points (row0 % 128 == 0);
 *                         sigma_out (nullable, one float per point) = softplus(sigma_from_xyz(h)), spsbrdfnerf.py:682;
 *   bn_mlp_heads_forward  feature layer + every head + sigma for all total_points rows -> packed rows.
 * bn_mlp_backward / bn_mlp_normals_* then run on the whole workspace (n_rays*n_samples = total_points). */
int bn_mlp_trunk_forward(bn_mlp* h, const float* params, const float* origins, int o_stride,
                         const float* dirs, int d_stride, const float* z, int n_rays, int n_samples, int flags,
                         int64_t total_points, int64_t row0, float* sigma_out, void* workspace,
                         size_t workspace_bytes, cudaStream_t stream);
/* Time embedding of the rays (rays_t = models['t'](ts), rendering.py:208-209, repeated per sample spsbrdfnerf.py:98) into the
 * workspace rows [row0, row0 + n_rays*n_samples) of a forward sized for total_points: t_rows (n_rays, t_dims) fp32, row r
 * serves the n_samples points of ray r.  Call it before bn_mlp_heads_forward / bn_mlp_forward with BN_MLP_BETA. */
int bn_mlp_write_t(bn_mlp* h, const float* t_rows, int n_rays, int n_samples, int flags, int64_t total_points, int64_t row0,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream);
int bn_mlp_heads_forward(bn_mlp* h, const float* params, int64_t total_points, int flags, float* out,
                         int out_pitch, void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* backward of the last BN_MLP_TRAIN forward on `workspace`: `out` is that forward's packed output,
 * g_out (P, out_pitch) its gradient -> g_params (flat fp32, ACCUMULATED: zero it at the start of a
 * step; it is the allreduce bucket). */
int bn_mlp_backward(bn_mlp* h, const float* params, const float* out, const float* g_out, int out_pitch,
                    int n_rays, int n_samples, int flags, float* g_params,
                    void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* Analytic normals (SpSBRDFNeRF.calc_normals, spsbrdfnerf.py:648-660,713-716): reverse sweep
 * d sigma / d x through the trunk using the cosines stored by the preceding bn_mlp_forward on the
 * same workspace (flags must carry BN_MLP_NORMAL_AN); writes -l2n(grad) into channels
 * normal_channel..+2 of the packed rows. */
int bn_mlp_normals_forward(bn_mlp* h, const float* params, float* out, int out_pitch, int n_rays,
                           int n_samples, int flags, int normal_channel, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream);
/* Second-order backward of the sweep (the double backward autograd runs through calc_normals):
 * consumes d loss / d normal from g_out, accumulates weight gradients of the sweep into g_params,
 * ADDS the induced d loss / d sigma into channel 3 of g_out and leaves the per-layer second-order
 * terms in the workspace for the bn_mlp_backward call that must follow. */
int bn_mlp_normals_backward(bn_mlp* h, const float* params, const float* out, float* g_out, int out_pitch,
                            int n_rays, int n_samples, int flags, int normal_channel, float* g_params,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* fused Adam on the flat buffers (torch.optim.Adam semantics, main.py:150): one launch per step */
int bn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                 float grad_scale, cudaStream_t stream);

/* The same update for a CUDA-graph-captured training step: the learning rate and the step counter live in DEVICE memory,
 * state = {lr, step (number of updates done so far, as a float), scratch, scratch}; the call increments step, derives the
 * bias corrections on the device and applies the update, so replaying the captured graph advances the optimizer.  One
 * launch; the scratch words must be zero when the state is created (a block counter lives there); buffers 16-byte aligned. */
int bn_adam_step_graph(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                       float* state, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                       cudaStream_t stream);

/* ------------------------------------------------------------------ tile inference -> DSM */

/* get_latlonalt_from_nerf_prediction (datasets/satellite_rgb_dep.py:601-634; callers eval.py:170, main.py:621):
 * point = ((double)o + (double)d * (double)depth) * scene_range + center, every operation rounded separately in float64 —
 * bit-exact against the reference's torch/numpy result.  cs = 1 ('utm', the reference default): the point is
 * (east, north, alt).  cs = 0 ('ecef'): the point is geocentric and goes through ecef_to_latlon_custom (sat_utils.py:127-146)
 * and the UTM projection of zone utm_zone (sat_utils.py:148-162: the zone of the FIRST point, which the caller supplies).  rays: n_rays records of ray_stride floats
 * [o(3), d(3), ...] (6 <= ray_stride <= 16); depth (n_rays); scene_range / center_*: the dataset's float32 `range` /
 * `center` values (satellite_rgb_dep.py:164-165) widened to double.  cloud (n_rays,3) float64 [east, north, alt];
 * points_f32 (n_rays,3) nullable: the same points rounded to float32 (the `pts3d` operand of calc_normal_from_depth_v2,
 * satellite_rgb_dep.py:578-585); bounds nullable: 4 doubles [xmin, xmax, ymin, ymax] of the finite points (the operands
 * of the grid derivation :666-671) with bounds_scratch = 4 uint64 of device scratch (both or neither). */
int bn_dsm_points(const float* rays, int ray_stride, const float* depth, long long n_rays, double scene_range,
                  double center_x, double center_y, double center_z, int cs, int utm_zone, double* cloud, float* points_f32,
                  double* bounds, unsigned long long* bounds_scratch, cudaStream_t stream);

/* Host-only query: bytes of device workspace bn_dsm_rasterize needs for this raster. */
size_t bn_dsm_workspace_bytes(int xsize, int ysize, int radius, float sigma);

/* plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma) as called at satellite_rgb_dep.py:680 (radius=1,
 * sigma=inf; third-party plyflatten==0.2.0, requirements.txt:11): point p falls into cell i = floor((x-xoff)/resolution),
 * j = floor((yoff-y)/resolution) and adds its value with weight w (1 when sigma is +inf, else exp(-dist^2/(2 sigma^2)) of
 * its distance to the cell centre) to every cell of the (2 radius+1)^2 window that lies inside the raster; raster
 * (ysize,xsize) float32 = weighted mean, NaN where no point contributed; count (ysize,xsize) nullable = weight sums.
 * cloud: n_points rows of cloud_stride doubles [x, y, ...], value_col (>= 2) selects the rasterised column.  Sums are
 * float64 atomics (the reference keeps an order-dependent float32 running mean: rasters agree to ~1e-4, counts exactly). */
int bn_dsm_rasterize(const double* cloud, int cloud_stride, int value_col, long long n_points, double xoff, double yoff,
                     double resolution, int xsize, int ysize, int radius, float sigma, float* raster, float* count,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);

/* The two halves of bn_dsm_rasterize, for a tile whose rays are sharded over several GPUs (pixel blocks per
 * rank): every rank accumulates its own points into its workspace (zero_first = 1 on the first call), the workspaces are
 * summed across ranks (one all-reduce of workspace_bytes viewed as cells float64 followed by cells float32,
 * cells = bn_dsm_workspace_bytes / 12), then bn_dsm_finalize turns the accumulators into the raster on every rank. */
int bn_dsm_accumulate(const double* cloud, int cloud_stride, int value_col, long long n_points, double xoff, double yoff,
                      double resolution, int xsize, int ysize, int radius, float sigma, int zero_first,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream);
int bn_dsm_finalize(int xsize, int ysize, int radius, float sigma, const void* workspace, size_t workspace_bytes,
                    float* raster, float* count, cudaStream_t stream);

/* calc_normal_from_pts3d with valid_depth=None (sat_utils.py:16-50): normals (height,width,3) of a float32 point image
 * (height,width,3) from the four normalised cross products of its neighbour differences; zero on the border. */
int bn_dsm_normals_from_points(const float* points, int height, int width, float* normals, cudaStream_t stream);

/* ------------------------------------------------------------------ ray feed: RPC camera -> ray records */

/* The attributes of rpcm.RPCModel the path reads (dict_format "rpcm": scene JSON "rpc" entries, satellite_rgb_dep.py:246);
 * rescale_rpc (sat_utils.py:90-108) multiplies row/col scale and offset by alpha.  HOST struct, passed by pointer. */
typedef struct bn_rpc {
  double row_offset, col_offset, lat_offset, lon_offset, alt_offset;
  double row_scale, col_scale, lat_scale, lon_scale, alt_scale;
  double row_num[20], row_den[20], col_num[20], col_den[20];
} bn_rpc;

/* get_rays (datasets/satellite_rgb_dep.py:23-78) [+ normalize_rays :550-559 when `normalize`, + the sun-direction
 * columns of get_sun_dirs :561-576 / hstack :390 when sun_dir != NULL]: for every pixel, rpc.localization at max_alt
 * (ray origin) and at min_alt (far point) by the iterative RPC inversion of the third-party `rpcm` package, mapped to
 * ECEF (cs = 0, sat_utils.py:110-125) or UTM (cs = 1, sat_utils.py:148-162 via pyproj: zone utm_zone, GRS80, northern
 * formula), o = near, d = unit(far - near), bounds [0, |far - near|], cast to float32 like the reference, then
 * (o - center) / scene_range, near / scene_range, far / scene_range in float32.
 * cols / rows: DEVICE (n_rays) float64 pixel coordinates, or both NULL = the row-major pixel grid of an image `width`
 * pixels wide (np.meshgrid(arange(w), arange(h)) flattened, :353).  sun_dir: HOST 3 floats or NULL.  rays_out: DEVICE
 * (n_rays, out_stride) float32, out_stride = 8, or 11 with sun_dir.  fail_count: DEVICE int (nullable), incremented for
 * every ray whose localisation did not converge in 100 iterations (the reference raises).  iterations: DEVICE 2 ints of
 * scratch / output: the iteration counts of the two localisation calls (max_alt, min_alt).  The reference's loop runs until
 * the slowest pixel of a call has converged and applies that many iterations to every pixel; a first launch finds the two
 * counts, a second one builds the rays with exactly those counts (no host round trip in between). */
int bn_rays_from_rpc(const bn_rpc* rpc, const double* cols, const double* rows, long long n_rays, int width,
                     double min_alt, double max_alt, int cs, int utm_zone, int normalize, float center_x, float center_y,
                     float center_z, float scene_range, const float* sun_dir, float* rays_out, int out_stride,
                     int* fail_count, int* iterations, cudaStream_t stream);

/* Unit-test hook: one GEMM of the MLP engine in isolation (kind 0: out[M,N] = A[M,K] B[N,K]^T;
 * kind 1: out[M,N] += A[K,M]^T B[K,N], fp32 atomics). precision selects tcgen05 (bf16 operands)
 * or CUDA cores (fp32 operands). */
int bn_debug_gemm(int kind, int precision, const void* A, long long lda, const void* B, long long ldb,
                  float* out, long long ldo, long long M, int N, long long K, cudaStream_t stream);

/* Diagnostics: device buffer (>= 32 * layers int64) that the fused density pass fills with clock64() stamps of the
 * first 256-point block of CTA pair 0 (MMA issuer and first epilogue warp); NULL switches it off. */
int bn_debug_chain_trace(bn_mlp* h, long long* device_buf);

/* Unit-test hook for the TMA-staged epilogues of the tcgen05 GEMM (bf16 operands only):
 *   kind 0: out_bf16[M,N] = ((A[M,K] B[N,K]^T) + add[M,N]) * mul[M,N]   (add / mul nullable, bf16, pitch ldo);
 *           colsum[N] (nullable, fp32) += column sums of out
 *   kind 1: out_f32[M, N - (pad_hi - pad_lo)] += A[K,M]^T B[K,N] through 32x32 fp32 TMA reduce-add boxes,
 *           dropping the packed columns [pad_lo, pad_hi) (pitch ldo). */
int bn_debug_gemm_epi(int kind, const void* A, long long lda, const void* B, long long ldb, void* out, long long ldo,
                      const void* add, const void* mul, float* colsum, int pad_lo, int pad_hi,
                      long long M, int N, long long K, cudaStream_t stream);

/* Data-parallel gradient exchange over NVLink peer memory (csrc/ddp.cu): in-place two-shot all-reduce (sum) of the flat fp32
 * gradient bucket, one kernel, CUDA-graph capturable.  Replaces the DDP all-reduce that Lightning runs for the reference
 * (main.py:720-731).  peer_bufs[p] / peer_flags[p]: rank p's bucket / flag array as mapped into THIS process (symmetric
 * memory: e.g. torch.distributed._symmetric_memory rendezvous), host arrays of `world` pointers; every rank's flag array
 * holds bn_allreduce_p2p_flag_words() uint32, zero-initialised once; epoch: local device array of 128 uint32, zero-
 * initialised once; n: bucket length in floats (multiple of 4); n_blocks (<= 128) must be equal on all ranks.  Every rank
 * must launch the call once per step; the sum is formed in rank order, so all ranks end with bit-identical buckets. */
int bn_allreduce_p2p(void* const* peer_bufs, void* const* peer_flags, uint32_t* epoch, int64_t n, int rank, int world,
                     int n_blocks, cudaStream_t stream);
int bn_allreduce_p2p_flag_words(void);
/* Peer-mapped device memory for bn_allreduce_p2p (CUDA IPC): every rank allocates its bucket with bn_peer_alloc (zeroed),
 * exports a 64-byte handle (bn_peer_export), sends it to the other ranks (any host channel: torch.distributed
 * all_gather), and opens theirs (bn_peer_open: the mapping is usable by kernels of this process, peer access over NVLink). */
int bn_peer_alloc(size_t bytes, void** ptr);
int bn_peer_free(void* ptr);
int bn_peer_export(void* ptr, void* handle64);
int bn_peer_open(const void* handle64, void** ptr);
int bn_peer_close(void* ptr);

/* Unit-test hook: where a per-point activation tensor of a forward call lives inside the caller's MLP workspace
 * (the layout bn_mlp_workspace_bytes sizes for `n_points`, `flags`).  which 0: the encoding X3 (64 columns),
 * 1: h_l = sin(.) of trunk layer `layer`, 2: c_l = w0 cos(.) of trunk layer `layer` (training / analytic normals only).
 * Writes the byte offset from the workspace base and the row pitch in elements (bf16 in BN_PREC_BF16 mode, else fp32);
 * BN_ERR_ARG when the tensor does not exist for these flags.  Lets tests compare the fused trunk kernel's stored
 * activations with the reference's per-layer values (models/spsbrdfnerf.py:636-646). */
int bn_debug_ws_tensor(const bn_mlp* h, int64_t n_points, int flags, int which, int layer, int64_t* offset_bytes,
                       int64_t* pitch_elems);

#ifdef __cplusplus
}
#endif
#endif /* BRDFNERF_B200_H_ */
