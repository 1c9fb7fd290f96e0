/*
 * neusky_b200 -- C ABI of the B200-native NeuSky render-and-shade hot path.
 *
 * The reference (JADGardner/neusky) is 100% Python and has no FFI of its own; its boundary
 * for this path is the nerfstudio Field/Model plugin API.  These are the
 * entry points a drop-in plugin binds instead of the reference's torch/tcnn calls.  Each
 * function names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *     tensors are dense, row-major, fp32 unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - the callee never allocates, frees or retains caller memory; outputs are caller-allocated;
 *   - return value: 0 on success, non-zero on error; nsk_last_error() then returns a
 *     thread-local, NUL-terminated description.  There is no CPU fallback.
 *   - all entry points are re-entrant and keep no global mutable state (the nerfstudio viewer
 *     thread may render while the trainer thread trains: neusky/models/neusky_model.py:1388-1403).
 */
#ifndef NEUSKY_B200_H
#define NEUSKY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSK_ABI_VERSION 1

int nsk_version(void);
const char* nsk_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * K1  multiresolution hash-grid encode (nerfstudio HashEncoding.pytorch_fwd semantics)
 * replaces tcnn.Encoding at neusky/fields/sdf_albedo_field.py:117-130 and
 * neusky/fields/directional_distance_field.py:139-156.
 *   x        [n,3]        positions (any range; negative coordinates are legal)
 *   table    [L*T, 2]     T = 1<<log2_T, level-major
 *   scalings [L]          per-level scale (floor(min_res*growth^l), computed by the host)
 *   out      [n, 2L]      level-major features
 * ------------------------------------------------------------------------------------------- */
int nsk_hash_encode_fwd(const float* x, int64_t n, const float* table, const float* scalings,
                        int num_levels, int log2_T, float* out, void* stream);

/* d(table) += scatter of grad_out through the trilinear weights.  grad_table [L*T,2] must be
 * zero-filled (or hold a running sum) by the caller. */
int nsk_hash_encode_bwd(const float* x, int64_t n, const float* scalings, int num_levels, int log2_T,
                        const float* grad_out, float* grad_table, void* stream);

/* d L / d x through the trilinear interpolation (what autograd returns for the encode's input): grad_x [n,3] =
 * sum_l scale_l * sum_ch grad_out[.,l,ch] * d feat / d offset.  The reference gets its normals this way
 * (torch.autograd.grad(sdf, x, create_graph=True), neusky/fields/sdf_albedo_field.py:235-238).
 * nsk_hash_encode_grad_x_bwd is ITS backward for a cotangent cot_x [n,3] (double backward of the encode, needed because the
 * normals feed the losses): d_grad_out [n,2L] (overwritten; NULL = skip) and d_table [L*T,2] (ACCUMULATED INTO; NULL = skip). */
int nsk_hash_encode_grad_x(const float* x, int64_t n, const float* table, const float* scalings, int num_levels,
                           int log2_T, const float* grad_out, float* grad_x, void* stream);
int nsk_hash_encode_grad_x_bwd(const float* x, int64_t n, const float* table, const float* scalings, int num_levels,
                               int log2_T, const float* grad_out, const float* cot_x, float* d_grad_out, float* d_table,
                               void* stream);

/* tcnn ("tiny-cuda-nn") grid semantics for imported reference checkpoints: the reference's encodings are tcnn.Encoding modules
 * (neusky/fields/sdf_albedo_field.py:117-130, neusky/fields/directional_distance_field.py:139-156).  x [n,3] in [0,1];
 * table = fp32 [L*T,2] filled by neusky_b200.tcnn_import.tcnn_params_to_table; level_meta int32 [L,4] = (float bits of the
 * level scale, resolution, entries in the level, dense flag), 16-byte aligned; smoothstep 0/1 -> out [n, 2L]. */
int nsk_hash_encode_tcnn_fwd(const float* x, int64_t n, const float* table, const int32_t* level_meta, int num_levels, int log2_T,
                             int smoothstep, float* out, void* stream);
/* Integer part only, for bit-exact index parity tests: idx [n,L,8] int64 (including the l*T level
 * offset, corner order of SURVEY A.3), offsets [n,L,3]. */
int nsk_hash_indices(const float* x, int64_t n, const float* scalings, int num_levels, int log2_T,
                     int64_t* idx, float* offsets, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  SDF / albedo field of a sample, fused with its hash encode and the ANALYTIC gradient d sdf / d x.
 * replaces SDFAlbedoField.get_outputs and get_sdf_at_pos (neusky/fields/sdf_albedo_field.py:169-174, 211-269):
 * nerfstudio SDFField.forward_geonetwork [SURVEY A.4] (L-inf scene contraction, (p+2)/4, hash grid, [x | PE6(x) | feat]
 * -> 256 -> 256 -> 1+256, softplus beta=100), torch.autograd.grad(sdf, x) (:235-238) and get_colors (:185-209).
 *   x [n,3]; sdf_weights = packed fp32 blob (python neusky_b200.packing.pack_sdf_simt, weight_norm folded);
 *   hash_table [L*T,2]; outputs sdf [n], grad [n,3] (NULL = skip the reverse pass), albedo [n,3] (NULL = skip the
 *   colour network), geo [n,256] (NULL = do not write the geometry feature).
 * nsk_sdf_field_simt_fwd: exact fp32 CUDA-core path.
 * nsk_sdf_field_tc_fwd  : tcgen05/TMEM tensor-core path (fp16 operands, fp32 accumulate; x enters as fp16 hi+lo, the sdf
 *                         row of the last layer is an fp32 dot product).  sdf_weights = nsk pack "sdf tc" blob
 *                         (neusky_b200.packing.pack_sdf_tc); always writes sdf, grad and albedo.
 * ------------------------------------------------------------------------------------------- */
int64_t nsk_sdf_tc_weights_bytes(void);
int nsk_sdf_field_tc_fwd(const float* x, int64_t n, const void* sdf_weights, const float* hash_table,
                         const float* scalings, int num_levels, int log2_T, float* sdf, float* grad, float* albedo,
                         void* stream);
int64_t nsk_sdf_simt_weights_floats(void);
int nsk_sdf_field_simt_fwd(const float* x, int64_t n, const float* sdf_weights, const float* hash_table,
                           const float* scalings, int num_levels, int log2_T, float* sdf, float* grad, float* albedo,
                           float* geo, void* stream);
/* The same two kernels on an IMPORTED tiny-cuda-nn grid (the reference's encodings are tcnn.Encoding modules, sdf_albedo_field.py:117-130;
 * neusky_b200/tcnn_import.py): grid_meta [L][4] int32 = (float bits of scale, resolution, size, dense) per level, 16-byte aligned, NULL =
 * the nerfstudio torch grid of the plain entry points; smoothstep = tcnn's interpolation flag (the analytic normal carries its derivative).
 * tcnn semantics are restated from memory and pinned by self-consistency tests only. */
int nsk_sdf_field_tc_fwd_ex(const float* x, int64_t n, const void* sdf_weights, const float* hash_table,
                            const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                            float* sdf, float* grad, float* albedo, void* stream);
int nsk_sdf_field_simt_fwd_ex(const float* x, int64_t n, const float* sdf_weights, const float* hash_table,
                              const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                              float* sdf, float* grad, float* albedo, float* geo, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  NeuS logistic-CDF alpha + transmittance + compositing, one warp per ray.
 * replaces SDFField.get_alpha (called neusky/fields/sdf_albedo_field.py:266),
 * RaySamples.get_weights_and_transmittance_from_alphas (neusky/models/neusky_model.py:565) and the
 * depth / accumulation / normal / albedo renderers (neusky_model.py:591-595, 806-813).
 *   sdf [R,S], grad [R,S,3], albedo [R,S,3], ray_dirs [R,3], starts/ends/deltas [R,S], dnorm [R]
 *   outputs: weights [R,S], wa [R,S,3] = weights*albedo, normals [R,S,3] = normalize(grad),
 *            acc [R], p2p_raw [R] (unclipped expected distance), normal_out [R,3],
 *            albedo_out [R,3] (white background, clamped to [0,1] unless training), bg_T [R],
 *            steps_minmax [2] (global min / max of (start+end)/2; caller pre-fills {+inf,-inf})
 * nsk_neus_finalize_depth clips p2p to the batch-global [min,max] (nerfstudio DepthRenderer) and
 * divides by directions_norm:  p2p [R], depth [R].
 * ------------------------------------------------------------------------------------------- */
int nsk_neus_composite_fwd(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                           const float* starts, const float* ends, const float* deltas,
                           int64_t R, int S, float inv_s, float cos_anneal_ratio, int training,
                           float* weights, float* wa, float* normals, float* acc, float* p2p_raw,
                           float* normal_out, float* albedo_out, float* bg_T, float* steps_minmax, void* stream);
/* Backward of nsk_neus_composite_fwd (training mode: no output clamps): cotangents of weights [R,S], wa [R,S,3], normals
 * [R,S,3], acc [R], p2p_raw [R], normal_out [R,3], albedo_out [R,3], bg_T [R] (any may be NULL = zero) ->
 * d_sdf [R,S], d_grad [R,S,3], d_albedo [R,S,3] (OVERWRITTEN) and d_inv_s [1] (ACCUMULATED INTO; caller zero-fills).
 * What torch autograd computes through SDFField.get_alpha, get_weights_and_transmittance_from_alphas and the renderers in
 * the reference (neusky/fields/sdf_albedo_field.py:266, neusky/models/neusky_model.py:565, 591-595, 806-813). */
int nsk_neus_composite_bwd(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                           const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                           float inv_s, float cos_anneal_ratio, const float* g_weights, const float* g_wa,
                           const float* g_normals, const float* g_acc, const float* g_p2p_raw,
                           const float* g_normal_out, const float* g_albedo_out, const float* g_bg_T,
                           float* d_sdf, float* d_grad, float* d_albedo, float* d_inv_s, void* stream);
/* Device-scalar variants of the two calls above: inv_s is read from inv_s_dev [1] ON THE DEVICE.  The training step computes
 * inv_s = exp(10 * variance) from a parameter the optimizer updates in HBM (nerfstudio LearnedVariance); reading it
 * back for a by-value argument costs a host synchronisation per call and cannot be captured in a CUDA graph. */
int nsk_neus_composite_fwd_dv(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                              const float* starts, const float* ends, const float* deltas,
                              int64_t R, int S, const float* inv_s_dev, float cos_anneal_ratio, int training,
                              float* weights, float* wa, float* normals, float* acc, float* p2p_raw,
                              float* normal_out, float* albedo_out, float* bg_T, float* steps_minmax, void* stream);
int nsk_neus_composite_bwd_dv(const float* sdf, const float* grad, const float* albedo, const float* ray_dirs,
                              const float* starts, const float* ends, const float* deltas, int64_t R, int S,
                              const float* inv_s_dev, float cos_anneal_ratio, const float* g_weights, const float* g_wa,
                              const float* g_normals, const float* g_acc, const float* g_p2p_raw,
                              const float* g_normal_out, const float* g_albedo_out, const float* g_bg_T,
                              float* d_sdf, float* d_grad, float* d_albedo, float* d_inv_s, void* stream);
int nsk_neus_finalize_depth(const float* p2p_raw, const float* dnorm, const float* steps_minmax, int64_t R,
                            float* p2p, float* depth, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RENI++ radiance table: HDR radiance of K latent codes in D directions.
 * replaces RENIField.get_outputs + unnormalise as driven by NeuSkyFactoModel.sample_illumination
 * (ns_reni/reni/illumination_fields/reni_illumination_field.py:493-573, base_spherical_field.py:143-154,
 *  ns_reni/reni/field_components/transformer_decoder.py:21-155, vn_layers.py:191-246,404-419;
 *  neusky/models/neusky_model.py:445-551).
 *   dirs [D,3], latents [K,Ld,3], scale [K] (NULL = no scale), rotation [3,3] (NULL = none; applied
 *   to the latent, Z@R), weights = packed fp32 blob (layout: nsk_reni_weights_floats / python
 *   neusky_b200.packing.pack_reni), workspace [K*6*H] floats, out [K,D,3] = exp(log-HDR) when log_domain == 1;
 *   log_domain == 2: log-domain model, out = the raw log-HDR value (what RENIField.forward returns before unnormalise()).
 * ------------------------------------------------------------------------------------------- */
int64_t nsk_reni_weights_floats(int latent_dim, int hidden, int num_layers);
int nsk_reni_decode_fwd(const float* dirs, int64_t D, const float* latents, const float* scale, int64_t K,
                        const float* rotation, const float* weights, int latent_dim, int hidden, int num_layers,
                        int log_domain, float* workspace, float* out, void* stream);
/* Per-row variant: direction n is decoded with latent code row_cam[n] (int32) -> out [N,3]; the background radiance along
 * each camera ray of a mixed-camera batch (neusky/models/neusky_model.py:535-549).  Same workspace as above. */
int nsk_reni_decode_rows_fwd(const float* dirs, const int* row_cam, int64_t N, const float* latents, const float* scale, int64_t K,
                             const float* rotation, const float* weights, int latent_dim, int hidden, int num_layers,
                             int log_domain, float* workspace, float* out, void* stream);
/* Fused tensor-core row decode (csrc/reni_fused_tc.cu): the whole decoder of the rows above in ONE tcgen05 kernel, fp16 operands, fp32
 * accumulate and LayerNorm; for frame-sized batches (per-ray background of a render, relighting sweeps).  zxy [K,Ld,2] and attn [K,6,128]
 * are the two halves of the nsk_reni_prep workspace; fused_weights = neusky_b200.packing.pack_reni_fused (nsk_reni_fused_weights_bytes()
 * bytes, 16-byte aligned); row_cam NULL = one latent code; log_domain as for nsk_reni_decode_fwd. */
int64_t nsk_reni_fused_weights_bytes(void);
int nsk_reni_rows_fused_fwd(const float* dirs, const int* row_cam, int64_t N, const float* zxy, const float* attn, const float* scale,
                            const void* fused_weights, int latent_dim, int log_domain, float* out, void* stream);
/* Backward of both variants w.r.t. the latent codes and scales, decoder frozen: what torch autograd computes for the
 * per-image `illumination_latents` / `scale` parameters (neusky_model.py:261-269, 488-504) under
 * RENIField.hold_decoder_fixed (reni_illumination_field.py:157-196).  row_cam NULL = table mode (out, g_out [K,D,3]), else
 * per-row mode (out, g_out [D,3]).  out = the forward result; weights_bwd = the [out][in] copy of the linear weights
 * (nsk_reni_bwd_weights_floats / neusky_b200.packing.pack_reni_bwd); workspace nsk_reni_bwd_workspace_floats floats;
 * d_latents [K,Ld,3] and d_scale [K] are ACCUMULATED into (d_scale NULL iff scale NULL). */
int64_t nsk_reni_bwd_weights_floats(int latent_dim, int hidden, int num_layers);
int64_t nsk_reni_bwd_workspace_floats(int64_t K, int latent_dim, int hidden, int num_layers);
int nsk_reni_decode_bwd(const float* dirs, const int* row_cam, int64_t D, const float* latents, const float* scale, int64_t K,
                        const float* rotation, const float* weights, const float* weights_bwd, int latent_dim, int hidden,
                        int num_layers, int log_domain, const float* out, const float* g_out, float* workspace,
                        float* d_latents, float* d_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Lambertian pre-pass and finalisation.
 * replaces the count / un-occluded part of RGBLambertianRendererWithVisibility.render_and_combine_rgb
 * (neusky/model_components/renderers.py:93-113) and its background blend + sRGB (:122-128, :173-174).
 *   normals [R,S,3], wa [R,S,3] (= weight*albedo), dirs [D,3], ddf_mask [D] (uint8: 1 = this
 *   direction is shaded through the DDF by nsk_sky_shade_*, 0 = visibility is the constant
 *   `unoccluded_vis`), radiance [K,D,3], cam [R] int32 row of the radiance table (NULL = row 0).
 *   outputs: inv_count [R,S] = 1/max(1,#{j: n.l_j>0}); rgb_lin [R,3] = sum over samples and over
 *   un-masked directions of wa * clamp(n.l) * inv_count * unoccluded_vis * radiance (OVERWRITTEN).
 * nsk_shade_finalize: rgb = clamp(sRGB(rgb_lin + bg*(1-acc))) (no final clamp when training).
 * ------------------------------------------------------------------------------------------- */
int nsk_lambert_prep(const float* normals, const float* wa, int64_t R, int S, const float* dirs,
                     const uint8_t* ddf_mask, int D, const float* radiance, const int32_t* cam,
                     float unoccluded_vis, float* inv_count, float* rgb_lin, void* stream);
/* Relighting with cached visibility (fixed geometry, new illumination; the reference's illumination animation
 * re-renders geometry AND visibility per frame, neusky/models/neusky_model.py:1896-1980): the same Lambertian sum as
 * nsk_lambert_prep + nsk_sky_shade_* with vis_sel [R,Dp] read from the cache instead of evaluating the DDF.
 * sel_index [D] int32: position of direction j inside the masked set, or -1 (visibility = unoccluded_vis).
 * rgb_lin [R,3] is OVERWRITTEN. */
int nsk_lambert_relight(const float* normals, const float* wa, const float* inv_count, int64_t R, int S,
                        const float* dirs, const int32_t* sel_index, int D, int Dp, const float* radiance,
                        const int32_t* cam, const float* vis_sel, float unoccluded_vis, float* rgb_lin, void* stream);
/* RENI++ decode of many rows on the tensor cores (csrc/reni_rows_tc.cu): the same arithmetic as nsk_reni_decode_rows_fwd
 * (RENIField.get_outputs, ns_reni/reni/illumination_fields/reni_illumination_field.py:493-573; Decoder, transformer_decoder.py:21-155)
 * with the 13 dense layers on nsk_gemm_tf32_nt (3xTF32).  These entry points are the pieces between the contractions:
 * nsk_reni_prep: workspace [K*NL*H + K*L*2] <- attention vectors [K,NL,H] then rotated latent xy [K,L,2] (per latent code).
 * nsk_reni_pe_rows: pe [N,512] <- decoder input rows (510 features, zero padded); row_cam [N] int32 = code of each row (NULL = 0);
 *     zxy = workspace + K*NL*H.
 * nsk_reni_ln_rows: x [N,128] <- LayerNorm(x + add[code(row) * add_stride : +128]) * ln_weight + ln_bias, in place (add NULL = 0); when
 *     ln_weight2 is given, followed in the same pass by x <- LayerNorm(x + add2[...]) * ln_weight2 + ln_bias2 (norm2 of one layer and
 *     norm1 of the next). */
int nsk_reni_prep(const float* latents, const float* rotation, int64_t K, const float* weights, int latent_dim, int hidden,
                  int num_layers, float* workspace, void* stream);
int nsk_reni_pe_rows(const float* dirs, const int* row_cam, int64_t N, const float* zxy, int latent_dim, float* pe, void* stream);
int nsk_reni_ln_rows(float* x, int64_t N, const float* add, int add_stride, const int* row_cam, const float* ln_weight,
                     const float* ln_bias, const float* add2, const float* ln_weight2, const float* ln_bias2, void* stream);
/* Collapsed relighting cache (config 5; replaces the per-frame re-render of the reference's illumination animation,
 * neusky/models/neusky_model.py:1896-1980, publication/render_animation.py:188-221):
 * nsk_lambert_collapse: H [R,D,3] = vis * sum_s wa * clamp01(n.l_j) * inv_count   (everything but the light colours; OVERWRITTEN)
 * nsk_relight_collapsed: rgb_lin [R,3] = sum_j H[r,j,:] * radiance[cam(r)][j,:]   (one streaming pass per new illumination) */
int nsk_lambert_collapse(const float* normals, const float* wa, const float* inv_count, int64_t R, int S, const float* dirs,
                         const int32_t* sel_index, int D, int Dp, const float* vis_sel, float unoccluded_vis, float* H,
                         void* stream);
int nsk_relight_collapsed(const float* H, int64_t R, int D, const float* radiance, const int32_t* cam, float* rgb_lin,
                          void* stream);
/* nsk_relight_collapsed_multi: NL illuminations per pass over H: radiance [NL,D,3] -> rgb_lin [NL,R,3] (an illumination sweep is bound
 * by streaming H; NL = 4 reads it once for four latent codes). */
int nsk_relight_collapsed_multi(const float* H, int64_t R, int D, const float* radiance, int NL, float* rgb_lin, void* stream);
/* Compact relighting cache (csrc/relight_compact.cu): only rays with accumulation > 0 own a row (`rows` [Rs] int32 ->
 * ray index), a row is fp16, channel-planar [3][DP] (DP = D rounded up to 16, directions in mma A-fragment order inside each block of 16), normalised by its own maximum (hscale [Rs]).
 * nsk_relight_pack_h16 builds it from the fp32 coefficients H [R,D,3]; nsk_relight_h16_multi streams it once per 32 illuminations:
 * radiance [NL,D,3] -> rgb_lin [NL,R,3] (zero for rays without a row).  H16 must be 16-byte aligned. */
int nsk_relight_pack_h16(const float* H, const int32_t* rows, int64_t Rs, int D, void* H16, float* hscale, void* stream);
int nsk_relight_h16_multi(const void* H16, const float* hscale, const int32_t* rows, int64_t Rs, int64_t R, int D, const float* radiance,
                          int NL, float* rgb_lin, void* stream);
/* nsk_lambert_collapse_sel: G [R,Dp,3] = sum_s wa * clamp01(n.l_j) * inv_count over the Dp directions that go through the DDF
 * (dirs_sel [Dp,3]).  nsk_sky_shade_tc2_fwd called with S = 0 takes this table in its `wa` argument (normals / inv_count may be
 * NULL) and skips the per-pair loop over the ray's samples -- the form full renders (S = 48..128) use. */
int nsk_lambert_collapse_sel(const float* normals, const float* wa, const float* inv_count, int64_t R, int S, const float* dirs_sel,
                             int Dp, float* G, void* stream);
/* Backward of nsk_lambert_relight for a cotangent g_rgb_lin [R,3]: d_wa [R,S,3], d_normals [R,S,3] (overwritten),
 * d_vis_sel [R,Dp] (overwritten; NULL = skip), d_radiance [K,D,3] (ACCUMULATED INTO with atomics; NULL = skip).  The
 * positively-lit count is piecewise constant, as in torch autograd through renderers.py:93-113. */
int nsk_lambert_relight_bwd(const float* normals, const float* wa, const float* inv_count, int64_t R, int S,
                            const float* dirs, const int32_t* sel_index, int D, int Dp, const float* radiance,
                            const int32_t* cam, const float* vis_sel, float unoccluded_vis, const float* g_rgb_lin,
                            float* d_wa, float* d_normals, float* d_vis_sel, float* d_radiance, void* stream);
/* Backward of nsk_shade_finalize: d_rgb_lin [R,3], d_bg [R,3], d_acc [R]. */
int nsk_shade_finalize_bwd(const float* rgb_lin, const float* bg, const float* acc, const float* g_rgb, int64_t R,
                           float* d_rgb_lin, float* d_bg, float* d_acc, void* stream);
int nsk_shade_finalize(const float* rgb_lin, const float* bg, const float* acc, int64_t R, int training,
                       float* rgb, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  outside-in sky visibility + cosine-weighted Lambertian sum, fused.
 * replaces NeuSkyFactoModel.compute_visibility (neusky/models/neusky_model.py:1624-1778),
 * DDFModel.get_outputs (neusky/models/ddf_model.py:158-219), DirectionalDistanceField.get_outputs
 * (neusky/fields/directional_distance_field.py:261-306), FiLMSiren (ns_reni/reni/field_components/
 * film_siren.py:45-156) and the visibility-weighted part of the Lambertian renderer
 * (neusky/model_components/renderers.py:93-113) for every (ray, direction) pair.
 *   points   [R,3]     surface points (already pulled inside the sphere, neusky_model.py:1667-1683)
 *   normals  [R,S,3], wa [R,S,3], inv_count [R,S]   per-sample shading inputs (from K3 / lambert_prep)
 *   dirs     [Dp,3]    the directions with ddf_mask==1, radiance [K,Dp,3] their HDR radiance
 *   cam      [R] int32 radiance row per ray (NULL = row 0)
 *   rgb_lin  [R,3]     ACCUMULATED INTO (atomics): += sum_s wa*clamp(n.l)*inv_count*vis*radiance
 *   vis_out  [R,Dp]    optional (NULL to skip): the visibility tensor; never written when NULL
 *   ddf_out  [R*Dp]    optional: expected termination distance; term_out [R*Dp] optional: |p-q|
 * nsk_sky_shade_simt_fwd : exact fp32 CUDA-core path.   ddf_weights = nsk pack "simt" blob (fp32).
 * nsk_sky_shade_tc_fwd   : tcgen05/TMEM tensor-core path, fp16 operands, fp32 accumulate.
 *                          ddf_weights = nsk pack "tc" blob (pre-tiled fp16 operand images + fp32 biases).
 * ------------------------------------------------------------------------------------------- */
int nsk_sky_shade_simt_fwd(const float* points, int64_t R, const float* normals, const float* wa,
                           const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                           const int32_t* cam, const float* ddf_weights, const float* hash_table,
                           const float* scalings, int num_levels, int log2_T, float radius, float threshold,
                           float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out, float* term_out,
                           void* stream);
int nsk_sky_shade_tc_fwd(const float* points, int64_t R, const float* normals, const float* wa,
                         const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                         const int32_t* cam, const void* ddf_weights, const float* hash_table,
                         const float* scalings, int num_levels, int log2_T, float radius, float threshold,
                         float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out, float* term_out,
                         void* stream);
/* CTA-pair variant of the tensor-core path (tcgen05.mma.cta_group::2, M = 256 across two SMs of a TPC; each CTA streams
 * half of every weight stage).  Same arguments and numerics as nsk_sky_shade_tc_fwd; ddf_weights = nsk pack "tc2" blob
 * (neusky_b200.packing.pack_ddf_tc2). */
int nsk_sky_shade_tc2_fwd(const float* points, int64_t R, const float* normals, const float* wa,
                          const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                          const int32_t* cam, const void* ddf_weights, const float* hash_table,
                          const float* scalings, int num_levels, int log2_T, float radius, float threshold,
                          float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out, float* term_out,
                          void* stream);
/* K4 on an imported tiny-cuda-nn position grid (directional_distance_field.py:139-156 builds a tcnn.Encoding): grid_meta / smoothstep as for
 * nsk_sdf_field_tc_fwd_ex; NULL = the plain entry points above. */
int nsk_sky_shade_tc2_fwd_ex(const float* points, int64_t R, const float* normals, const float* wa,
                             const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                             const int32_t* cam, const void* ddf_weights, const float* hash_table,
                             const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                             float radius, float threshold, float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out,
                             float* term_out, void* stream);
int nsk_sky_shade_simt_fwd_ex(const float* points, int64_t R, const float* normals, const float* wa,
                              const float* inv_count, int S, const float* dirs, int Dp, const float* radiance,
                              const int32_t* cam, const float* ddf_weights, const float* hash_table,
                              const float* scalings, int num_levels, int log2_T, const int32_t* grid_meta, int smoothstep,
                              float radius, float threshold, float sigmoid_scale, float* rgb_lin, float* vis_out, float* ddf_out,
                              float* term_out, void* stream);
int64_t nsk_ddf_tc2_weights_bytes(void);
int64_t nsk_ddf_simt_weights_floats(void);
int64_t nsk_ddf_tc_weights_bytes(void);

/* Surface points for visibility incl. the outside-sphere replacement (neusky_model.py:1667-1683):
 * origins [R,3], ray_dirs [R,3], p2p [R] -> points [R,3]. */
int nsk_surface_points(const float* origins, const float* ray_dirs, const float* p2p, int64_t R, float radius,
                       float* points, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training step (BASELINE config 4): the layer contractions and the pointwise stages of the reference's
 * autograd graph, forward and backward.  The reference runs these as torch.nn.Linear / cuBLAS calls recorded
 * by autograd (film_siren.py:45-156 for the DDF, sdf_albedo_field.py:185-269 + nerfstudio SDFField for the
 * SDF/colour MLP); here every contraction is one tcgen05 tf32 launch with fp32 operands read straight from HBM
 * and the pointwise stage fused into its epilogue where it is a function of the output element only.
 *
 * nsk_gemm_tf32_nt:  C[M,N] = dact'(aux) * act(A[M,K] . B[N,K]^T + bias[N])  (+ C when accumulate)
 *     A, B row-major with leading dimensions lda, ldb (multiples of 4, 16-byte aligned bases), K a multiple of 8
 *     (pad with zero columns).  act: 0 none, 1 relu, 2 leaky-relu(0.2), 3 softplus(beta=100), 4 sigmoid.
 *     dact != 0 multiplies by the derivative of that activation expressed through its forward OUTPUT aux[M, ldaux]
 *     (the backward "dZ = dA * act'(z)" step fused into dA = dZ_next . W).
 * nsk_gemm_tf32_tn:  C[P,Q] += sum_m A[m,P]^T B[m,Q]   (weight gradients; the m range is split across CTAs and
 *     reduced with red.global.add, so C must hold zeros or a running sum).
 * split: 1 = one tf32 pass (10-bit mantissa operands, fp32 accumulate); 3 = 3xTF32 (hi/lo operand split, three MMAs
 *     per step), fp32-accurate -- the mode the parity tests pin against the fp32 oracle.
 * ------------------------------------------------------------------------------------------- */
int nsk_gemm_tf32_nt(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                     const float* bias, int act, const float* aux, int ldaux, int dact, int accumulate, int split,
                     void* stream);
int nsk_gemm_tf32_tn(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int P, int Q,
                     int split, void* stream);

/* DDF visibility network, training path.  Row i = r * D + j of every [N, .] tensor is the pair (surface point r,
 * light direction j), N = R * D (neusky_model.py:1685-1690).
 * nsk_ddf_pairs_fwd: points [R,3], dirs [D,3] (unit, already masked to the upper hemisphere) ->
 *     cond [N,40] = (q, hash(q), 0-pad)   q = sphere exit point (neusky_model.py:1693-1695; directional_distance_field.py:267-268)
 *     xin  [N,16] = (d_local, PE2(d_local), 0-pad) for d = -l in the local frame of q (ddf_model.py:158-200; :270-271)
 *     q    [N,3], term_dist [N] = |q - p| (neusky_model.py:1697-1699)
 * nsk_ddf_rows_fwd: the same cond / xin for rows that carry their own sphere point and world direction, origins [N,3],
 *     directions [N,3] -- DDFModel.get_outputs on a sampled ray bundle and its multi-view / sky-ray batches (the DDF fitting
 *     pass, ddf_model.py:193-217, 279-363).
 * nsk_film_sin_fwd / _bwd: a = sin((15 f + 30) z + phase) with f, phase = columns [layer*256, +256) of the two halves of the
 *     mapping output film [N, ldf] (film_siren.py:66-67, 74-81, 140); bwd writes d z [N,256] and the two column blocks of d film.
 * nsk_ddf_head_fwd: that = 2 r sigmoid(a5 . w + b) (directional_distance_field.py:297-299), vis = 1 - sigmoid(scale *
 *     (min(term_dist, 2r) - that - threshold)) (neusky_model.py:1724-1740); threshold is a device scalar (learnable).
 * nsk_ddf_head_bwd: cotangents d_vis [N] (NULL = 0) and d_that_extra [N] (NULL = 0; the sdf_at_termination branch) ->
 *     d a5 [N,256] (overwritten), d w [256], d b [1], d threshold [1] (NULL = skip) accumulated.
 * nsk_colsum: out[c] += sum_r X[r, c]  (bias gradients). */
int nsk_ddf_pairs_fwd(const float* points, int64_t R, const float* dirs, int D, const float* table, const float* scalings,
                      int num_levels, int log2_T, float radius, float* cond, float* xin, float* q, float* term_dist,
                      void* stream);
int nsk_ddf_rows_fwd(const float* origins, const float* directions, int64_t N, const float* table, const float* scalings,
                     int num_levels, int log2_T, float* cond, float* xin, void* stream);
int nsk_film_sin_fwd(const float* z, const float* film, int ldf, int layer, int64_t N, float* a, void* stream);
int nsk_film_sin_bwd(const float* da, const float* z, const float* film, int ldf, int layer, int64_t N, float* dz,
                     float* dfilm, void* stream);
/* nsk_film_sin_bwd that ALSO accumulates (atomics; caller zero-fills) the column sums the bias gradients need:
 * sum_dz [256] += sum_r dz[r, :] (trunk layer `layer`), sum_dfilm [ldf] += sum_r dfilm[r, :] on this layer's two column blocks
 * (the last mapping layer's bias gradient): saves the separate column-sum passes over [N,256] per layer and [N,ldf]. */
int nsk_film_sin_bwd_sums(const float* da, const float* z, const float* film, int ldf, int layer, int64_t N, float* dz,
                          float* dfilm, float* sum_dz, float* sum_dfilm, void* stream);
int nsk_ddf_head_fwd(const float* a5, const float* w_final, const float* b_final, const float* term_dist, int64_t N,
                     float radius, const float* threshold, float sigmoid_scale, float* that, float* vis, void* stream);
int nsk_ddf_head_bwd(const float* a5, const float* w_final, const float* that, const float* term_dist, const float* d_vis,
                     const float* d_that_extra, int64_t N, float radius, const float* threshold, float sigmoid_scale,
                     float* da5, float* d_w_final, float* d_b_final, float* d_threshold, void* stream);
int nsk_colsum(const float* X, int ld, int64_t M, int ncols, float* out, void* stream);

/* SDF / albedo field, training path (sdf_albedo_field.py:211-269; nerfstudio SDFField.forward_geonetwork).
 * nsk_sdf_inputs_fwd: x [n,3] -> H0 [n,72] = (x | PE6(x) | hash((contract_inf(x)+2)/4) | 0) (the geo-network input in the
 *     reference's concatenation order), tail [n, ld_tail] columns [0,40) = (x | PE6(x) | 0) (colour-network input tail; NULL =
 *     skip), pos [n,3] (hash-grid position) and J [n,9] = d pos / d x.
 * nsk_sdf_grad_assemble: input stage of the reverse pass that replaces torch.autograd.grad(sdf, x) (sdf_albedo_field.py:235-238):
 *     grad_x [n,3] = G[:,0:3] + PE6'(x)^T G[:,3:39] + J^T gpos, with G [n,72] = d . / d H0 and gpos [n,3] = nsk_hash_encode_grad_x
 *     of G[:,39:71].  nsk_sdf_grad_assemble_bwd is its transpose for a cotangent c [n,3] (the normals' double backward): writes
 *     columns [0,39) and 71 of dG [n,72] and cpos [n,3] = J c; the hash columns come from nsk_hash_encode_grad_x_bwd.
 * nsk_ew256: pointwise family over [n,256] tensors with s' / s'' of softplus(beta=100) taken from the activation OUTPUT:
 *     0: out = w[col] s'(a)   1: out = a s'(b)   2: out = a s'(b) + c d s''(b)   3: out = a s'(b) + c w[col] s''(b)
 *     4: out = a + s[row] w[col]      (NULL a / c drop that term).
 * nsk_rowdot256: out[r] = X[r,:256] . w + b[0] (the sdf head, exact fp32).  nsk_colsum_w: out[c] += sum_r v[r] X[r,c]. */
int nsk_sdf_inputs_fwd(const float* x, int64_t n, const float* table, const float* scalings, int num_levels, int log2_T,
                       float* H0, float* tail, int ld_tail, float* pos, float* J, void* stream);
int nsk_sdf_grad_assemble(const float* x, const float* G, const float* gpos, const float* J, int64_t n, float* grad,
                          void* stream);
int nsk_sdf_grad_assemble_bwd(const float* x, const float* c, const float* J, int64_t n, float* dG, float* cpos,
                              void* stream);
int nsk_ew256(int op, int64_t n, const float* a, const float* b, const float* c, const float* d, const float* w,
              const float* s, float* out, void* stream);
int nsk_rowdot256(const float* X, const float* w, const float* b, int64_t n, float* out, void* stream);
int nsk_colsum_w(const float* X, int ld, const float* v, int64_t M, int ncols, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Proposal-network sampler: the sample placement in front of the path.
 * replaces nerfstudio ProposalNetworkSampler.generate_ray_samples as called at neusky/models/neusky_model.py:561
 * (UniformSampler -> HashMLPDensityField.density_fn -> RaySamples.get_weights -> PDFSampler, NeuSFactoModelConfig defaults,
 * SURVEY A.6) and the interlevel loss that trains the proposal networks (neusky_model.py:575-576, 987-988).
 *
 * Sample placement lives in the SPACING domain: bins [R,S+1] in [0,1]; the euclidean edge is bin*far + (1-bin)*near.
 *   nsk_uniform_bins      base [S+1] = linspace(0,1,S+1) (host), jitter [R] in [0,1) or NULL (eval) -> bins [R,S+1]
 *   nsk_proposal_density_fwd   HashMLPDensityField: density [R,S] at the bin mid-points of rays (origins, dirs [R,3], near, far
 *       [R]); dirs == NULL: positions mode, `origins` is [R*S,3] world positions (density_fn(positions)).
 *       table [L*T,2], scalings [L]; mlp = packed blob W0 [2L][16] (input-major) | b0 [16] | W1 [16] | b1 (nsk_proposal_mlp_floats).
 *   nsk_proposal_density_bwd   g_density [R,S] -> d_table [L*T,2], d_mlp (both ACCUMULATED INTO).
 *   nsk_pdf_resample      density [R,S] (get_weights is fused) OR weights_in [R,S]  -> weights^anneal + histogram_padding -> pdf ->
 *       cdf -> N+1 new bin edges at u = u_base[j] + (jitter ? jitter[r]/(N+1) : u_half); u_base [N+1] =
 *       linspace(0, 1-1/(N+1), N+1) (host).  Writes weights_out [R,S] (NULL = skip), new_bins [R,N+1], new_euclid [R,N+1] (NULL =
 *       skip).  Bit-exact vs the oracle given the weights (fp64 scans, see proposal_sampler.cu).
 *   nsk_density_weights_bwd    g_weights [R,S] -> g_density [R,S] through RaySamples.get_weights.
 *   nsk_interlevel_loss   c [R,Sf+1], w [R,Sf] (fine histogram, detached), cp [R,Sp+1], wp [R,Sp] (proposal) -> loss_ray [R] =
 *       sum_i relu(w_i - outer_i)^2 / (w_i + eps), g_wp [R,Sp] = d loss_ray / d wp (NULL = skip).
 * ------------------------------------------------------------------------------------------- */
int64_t nsk_proposal_mlp_floats(int num_levels, int hidden);
int nsk_uniform_bins(const float* base, const float* jitter, int64_t R, int S, float* bins, void* stream);
int nsk_proposal_density_fwd(const float* origins, const float* dirs, const float* near, const float* far, const float* bins,
                             int64_t R, int S, const float* table, const float* scalings, int num_levels, int log2_T,
                             const float* mlp, int hidden, float* density, void* stream);
int nsk_proposal_density_bwd(const float* origins, const float* dirs, const float* near, const float* far, const float* bins,
                             int64_t R, int S, const float* table, const float* scalings, int num_levels, int log2_T,
                             const float* mlp, int hidden, const float* g_density, float* d_table, float* d_mlp, void* stream);
int nsk_pdf_resample(const float* bins, const float* density, const float* weights_in, const float* near, const float* far,
                     int64_t R, int S, int N, float anneal, float histogram_padding, float eps, const float* u_base,
                     float u_half, const float* jitter, float* weights_out, float* new_bins, float* new_euclid, void* stream);
int nsk_density_weights_bwd(const float* bins, const float* density, const float* near, const float* far, int64_t R, int S,
                            const float* g_weights, float* g_density, void* stream);
int nsk_interlevel_loss(const float* c, const float* w, int Sf, const float* cp, const float* wp, int Sp, int64_t R,
                        float* loss_ray, float* g_wp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Light-sum shaders (csrc/shaders.cu): ns_reni's inverse-rendering shaders and NeuSky's Blinn-Phong renderer branch.
 * Replaces  reni.model_components.shaders.LambertianShader.forward   (ns_reni/reni/model_components/shaders.py:25-70)   mode 0
 *           reni.model_components.shaders.BlinnPhongShader.forward   (ns_reni/reni/model_components/shaders.py:73-161)  mode 1
 *           RGBBlinnPhongRendererWithVisibility.render_and_combine_rgb light sum (neusky/model_components/renderers.py:199-241) mode 2
 * Rows i < N are pixels (modes 0, 1) or ray samples (mode 2).  albedo, normals, specular, view_dirs [N,3]; shininess [N];
 * dirs [M,3] shared by all rows (dirs_per_row = 0) or [N,M,3]; radiance [K,M,3] with cam [N] -> K (NULL = row 0 for all);
 * vis [ceil(N / rows_per_vis), M] (NULL = 1; mode 2).
 *   mode 0: out_a = sum_j max(n.l_j, 0) L_j              out_b = albedo * out_a
 *   mode 1: out_a = max(albedo * sum_j max(n.l_j,0) L_j + specular * (s+2)/(4(2-exp(-s/2))) * sum_j max(n.h_j,0)^s L_j, 1e-3),
 *           h_j = (l_j + v) / (|l_j + v| + 1e-8)
 *   mode 2: out_a = sum_j L_j vis_j (albedo * clamp01(n.l_j) + clamp01(n.h_j)^s),  h_j = (l_j + v) / |l_j + v|;
 *           with weights [N] and rgb_lin [N/S,3] (zero-filled): rgb_lin[i / S] += weights[i] * out_a[i]  (renderers.py:247).
 * nsk_shade_lights_bwd: cotangents g_a (on out_a), g_b (on out_b, mode 0; either may be NULL) -> d_albedo, d_normals, d_specular
 *   [N,3], d_shininess [N] (overwritten; NULL = skip) and d_radiance [K,M,3], d_vis (ACCUMULATED INTO; NULL = skip).
 * ------------------------------------------------------------------------------------------- */
int nsk_shade_lights_fwd(int mode, const float* albedo, const float* normals, const float* specular, const float* shininess,
                         const float* view_dirs, const float* dirs, int dirs_per_row, int normalize_dirs, const float* radiance,
                         const int* cam, const float* vis, int rows_per_vis, int64_t N, int M, float* out_a, float* out_b,
                         const float* weights, float* rgb_lin, int S, void* stream);
int nsk_shade_lights_bwd(int mode, const float* albedo, const float* normals, const float* specular, const float* shininess,
                         const float* view_dirs, const float* dirs, int dirs_per_row, int normalize_dirs, const float* radiance,
                         const int* cam, const float* vis, int rows_per_vis, int64_t N, int M, const float* g_a, const float* g_b,
                         float* d_albedo, float* d_normals, float* d_specular, float* d_shininess, float* d_radiance,
                         float* d_vis, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEUSKY_B200_H */
