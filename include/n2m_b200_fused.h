/* n2m_b200_fused.h -- C ABI of the FUSED stage-0 train path of libn2m_b200.so.
 *
 * The operator-level entry points in n2m_b200.h are the drop-in replacements for the reference's
 * pybind functions.  The entry points here implement the same arithmetic as ONE pipeline with a
 * B200-native data layout, no host synchronisation and no per-step allocation, so the whole train
 * step can be captured in a CUDA graph.  What each stage replaces in the reference:
 *
 *   n2m_s0_march          raymarching.near_far_from_aabb + march_rays_train (raymarching.py:19-49,181-245;
 *                          raymarching.cu:92-145,338-475) without the .item() sync (raymarching.py:232)
 *   n2m_s0_encode_fwd     GridEncoder.forward x2 (grid.py:151-168; gridencoder.cu:88-196) + cat + safe_normalize
 *                          + grad_total_variation (gridencoder.cu:506-609; utils.py:801-823)
 *   n2m_s0_mlp_fwd        sigma_net / color_net / specular_net + trunc_exp / sigmoid / clamp
 *                          (nerf/network.py:81-108,159-189) on tcgen05 tensor cores
 *   n2m_s0_composite_loss composite_rays_train fwd+bwd (raymarching.cu:501-694), background mix
 *                          (renderer.py:804) and the MSE(+mask, +specular) loss (utils.py:660-738)
 *   n2m_s0_mlp_bwd        autograd of the three MLPs (dgrad + wgrad) on tcgen05
 *   n2m_s0_encode_bwd     grid_encode backward x2 (gridencoder.cu:248-339)
 *   n2m_s0_adam           GradScaler.unscale_/step/update + Adam(eps 1e-15) + the per-step fp32->fp16
 *                          table cast (grid.py:45-46) + zero_grad  (utils.py:549,1163,1176-1177)
 *
 * Data layout (all buffers allocated by the caller, see nerf2mesh_b200/stage0.py):
 *   table      [rows]  8-byte entries {float density_feature; half2 colour_features}: the density table
 *              (fp32, C=1) and the fp16 working copy of the colour table (C=2) interleaved so that one
 *              64-bit access serves both encoders.  fp32 colour masters live in `color_master [rows] float2`.
 *   gtable     [rows]  float4 {g_density, g_colour0, g_colour1, 0}, loss-scaled, reduced with one
 *              red.global.add.v4.f32 per lattice corner.
 *   enc_tiles  one 16 KiB image per 128 samples: fp16 [128 x 64] in the UMMA no-swizzle core-matrix layout
 *              (chunk-major, tc05.cuh): cols 0-2 xyz, 3-18 density features, 19-50 colour features,
 *              51-53 unit view direction, 54-63 zero.  It is the A operand of every first-layer GEMM and is
 *              staged global->shared with a single bulk async copy.
 *   recs       [Mcap] float4 {t_before, dt, t_after, ray_id (bits)} per sample, ray order.
 *   counters   int32 [16]: [0] M (total samples marched), [1] min(M, Mcap), [2] overflow flag, [3] / [15] samples inside / outside the
 *              unit cube (counted by the TV pass, read by n2m_s0_tv_random),
 *              [4..12] sample offset of the first ray of every eighth of the batch (ray N*e/8, e = 0..8; [4] = 0,
 *              [12] = [1]): the boundaries of the ray-range parts of the *_part entry points below.
 *              [13] += 1 for every march whose M exceeded Mcap, [14] = largest M seen (persistent capacity accounting: the rays
 *              that do not fit -- always the last rays of the batch -- are rendered as background and get no gradient, which
 *              the reference never does (it allocates exactly M, raymarching.py:232-238); hosts must watch [13] and grow Mcap).
 *              Entry points without parts (and nparts == 1) read only [1], so hand-filled 4-entry arrays keep working.
 *   wpack      packed fp16 MLP weights in tensor-core tile layout (n2m_s0_pack_weights).
 */
#ifndef N2M_B200_FUSED_H
#define N2M_B200_FUSED_H

#include "n2m_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    float bound;            /* marching bound (renderer.real_bound) */
    float grid_bound;       /* bound that normalises positions for the hash grid (renderer.bound) */
    float inv_2gb;          /* float32(1) / float32(2 * grid_bound) -- torch divides by a scalar this way */
    float dt_gamma;
    float min_near;
    float T_thresh;
    float S;                /* log2(per_level_scale) as float32 (grid.py:38) */
    float lambda_mask;
    float lambda_specular;
    float lambda_tv;
    float lambda_entropy;   /* utils.py:728-733: entropy of the per-sample weights and of weights_sum (0 = off) */
    uint32_t contract;
    uint32_t max_steps;
    uint32_t cascades;
    uint32_t grid_size;
    uint32_t num_levels;    /* must be 16 */
    uint32_t base_res;
    uint32_t shading_full;  /* 0 = 'diffuse' (first diffuse_step iterations), 1 = 'full' */
    uint32_t gt_has_alpha;  /* gt is rgba: blend with bg and add the mask loss (utils.py:662-667,681-683) */
} n2m_s0_params;

/* one-time per-process setup (kernel attributes); call before the first fused launch / graph capture */
int n2m_s0_init(void);

/* test hook: 1 = sequential one-thread-per-ray marcher, 0 = warp-per-ray marcher (default); same results */
int n2m_s0_set_serial_march(int on);

/* where the TV gradient is evaluated: 0 = inside the backward scatter kernel, 2 = own launch n2m_s0_tv (default of the host code) */
int n2m_s0_set_tv_mode(int mode);
/* the stand-alone TV launch (tv mode 2): reads recs/table, adds into gtable; independent of the MLP kernels */
int n2m_s0_tv(const n2m_s0_params* p, const void* recs, const int32_t* counters, uint32_t Mcap, const float* rays_o,
              const float* rays_d, const void* table, const int32_t* offsets, void* gtable, const float* loss_scale,
              n2m_stream_t stream);
/* GridEncoder.grad_total_variation's fallback (grid.py:181-183): a TV call of post_train_step (utils.py:815-823) that received no sample
 * position evaluates the TV gradient at `num_points` (reference: 10^6) uniformly random points instead.  The TV pass of the step counts the
 * samples inside / outside the unit cube into counters[3] / counters[15]; this launch (after it, same stream) adds the fallback of every
 * group that stayed empty and exits at once otherwise.  Points: counter-based hash of (optimizer step, index).  `dump` (nullable,
 * [num_points,3]): test hook -- run unconditionally with weight lambda_tv and store the points. */
int n2m_s0_tv_random(const n2m_s0_params* p, const int32_t* counters, const void* table, const int32_t* offsets, void* gtable,
                     const float* loss_scale, uint32_t num_points, float* dump, n2m_stream_t stream);

/* sizes of the packed weight blob (bytes) and of the flat fp32 MLP parameter / gradient vector (floats) */
uint32_t n2m_s0_wpack_bytes(void);
uint32_t n2m_s0_mlp_param_count(void);      /* 7648 = 608+32 + 2240+4096+384 + 192+96 */

/* fp32 masters (reference nn.Linear layouts [out,in], concatenated: sigma0, sigma1, color0, color1, color2,
 * spec0, spec1) -> packed fp16 tiles */
int n2m_s0_pack_weights(const float* mlp_params, void* wpack, n2m_stream_t stream);

/* interleave / de-interleave the hash tables (import / export of reference-format tensors) */
int n2m_s0_pack_tables(const float* emb_density, const float* emb_color, uint32_t rows,
                       void* table, void* color_master, n2m_stream_t stream);
int n2m_s0_unpack_tables(const void* table, const void* color_master, uint32_t rows,
                         float* emb_density, float* emb_color, n2m_stream_t stream);
/* gtable (loss-scaled float4) -> reference-format gradients, divided by *loss_scale */
int n2m_s0_unpack_grads(const void* gtable, uint32_t rows, const float* loss_scale,
                        float* g_density, float* g_color, n2m_stream_t stream);

/* march: near/far + count + scan + sample records.  cam_near_far [N,2] nullable (renderer.py:689-691). */
int n2m_s0_march(const n2m_s0_params* p, const float* rays_o, const float* rays_d, const float* aabb,
                 const float* cam_near_far, const uint8_t* bitfield, const float* noises, uint32_t N,
                 int32_t* rays, int32_t* counters, float* tbuf, void* recs, uint32_t Mcap,
                 n2m_stream_t stream);

/* gtable / loss_scale: unused (kept for ABI stability; the TV gradient is evaluated by n2m_s0_tv or by the scatter) */
int n2m_s0_encode_fwd(const n2m_s0_params* p, const void* recs, const int32_t* counters, uint32_t Mcap,
                      const float* rays_o, const float* rays_d, const void* table, const int32_t* offsets,
                      void* enc_tiles, void* gtable, const float* loss_scale, n2m_stream_t stream);

/* same gather for explicit positions xyz [P,3] in [-bound, bound] (dirs [P,3] nullable); counters[1] = P.  Used by the
 * density-grid update and by tests. */
int n2m_s0_encode_points(const n2m_s0_params* p, const float* xyz, const float* dirs, const int32_t* counters, uint32_t Pcap,
                         const void* table, const int32_t* offsets, void* enc_tiles, n2m_stream_t stream);

/* density-grid update pieces (NeRFRenderer.update_extra_state, renderer.py:1074-1149):
 *   grid_points : jittered centres of cells [first_cell, first_cell+count) of one cascade, Morton order; noise [H^3,3] in [0,1) is the
 *                 cascade's whole draw in the reference's meshgrid order (row x*H*H + y*H + z), i.e. torch.rand_like(cas_xyzs)
 *   grid_update : cells[i] = max(cells[i] * decay, sigma_i) where both >= 0 (sigma_i = out[i].x)
 *   packbits_dev: bit = grid > min(*mean_density, density_thresh), threshold read on the device */
int n2m_s0_grid_points(uint32_t H, uint32_t first_cell, uint32_t count, float cas_bound, const float* noise, float* xyz,
                       n2m_stream_t stream);
int n2m_s0_grid_update(const void* out, uint32_t count, float decay, float* grid_cells, n2m_stream_t stream);
int n2m_s0_packbits_dev(const float* grid, uint32_t nbytes, const float* mean_density, float density_thresh, uint8_t* bitfield,
                        n2m_stream_t stream);

/* NeRFRenderer.mark_untrained_grid (renderer.py:985-1071; called once by Trainer.train, utils.py:925): density_grid [cascades, H^3]
 * cells (Morton order) that no camera sees (camera-space z > near, |x| < cx/fx * z + 2 * half_cell, same for y) or that lie outside
 * aabb (+- half a cell) are set to -1.  poses [num_poses,4,4] camera-to-world; intrinsics [intr_count,4] = (fx, fy, cx, cy) on the
 * DEVICE with intr_count 1 or num_poses; cam_near [num_poses] nullable (else min_near); *count_out (nullable) = marked cells. */
int n2m_mark_untrained_grid(const float* poses, uint32_t num_poses, const float* intrinsics, uint32_t intr_count,
                            const float* cam_near, float min_near, const float* aabb, float bound, uint32_t cascades,
                            uint32_t H, float* density_grid, int32_t* count_out, n2m_stream_t stream);

/* batch sampling on the device = get_rays (nerf/utils.py:236-290) + the stage-0 training collate (nerf/provider.py:300-331)
 * for N random (image, pixel) pairs: poses [num_poses,4,4] (device), intrinsics_host float[4] {fx, fy, cx, cy} (HOST),
 * img_idx / pix_idx int32 [N] (device; pix = j * W + i), images uint8 [num_poses, H, W, C] (device, nullable with gt).
 * Writes rays_o, rays_d [N,3] (unnormalised directions) and gt [N,C] = pixel / 255.  Indices are not range-checked. */
int n2m_s0_gen_rays(const float* poses, uint32_t num_poses, const float* intrinsics_host, uint32_t H, uint32_t W,
                    const int32_t* img_idx, const int32_t* pix_idx, const uint8_t* images, uint32_t C, uint32_t N,
                    float* rays_o, float* rays_d, float* gt, n2m_stream_t stream);

/* out [Mcap] float4 {sigma, r, g, b}; spec_sq_sum: += sum over samples of |specular|^2 (for the loss value) */
int n2m_s0_mlp_fwd(const n2m_s0_params* p, const void* enc_tiles, const int32_t* counters, uint32_t Mcap,
                   const void* wpack, void* out, float* spec_sq_sum, n2m_stream_t stream);

/* per ray: composite, loss, composite backward.  gt [N,4] (rgba) or [N,3]; bg [N,3].
 * dout [Mcap] float4 {dL/dsigma, dL/dr, dL/dg, dL/db} * loss_scale (zero beyond each ray's break).
 * loss_out float[4]: [0] += sum_rays per-ray loss / N (rgb + mask + ray-level entropy), [1] is the MLP forward's sum |spec|^2,
 * [2] += sum of H(weights_k) over the weights the compositor touched, [3] += their count (lambda_entropy > 0 only);
 * image/ws/depth [N] outputs. */
int n2m_s0_composite_loss(const n2m_s0_params* p, const void* out, const void* recs, const int32_t* rays,
                          const int32_t* counters, uint32_t N, uint32_t Mcap, const float* gt, const float* bg,
                          const float* loss_scale, void* dout, float* image, float* weights_sum, float* depth,
                          float* loss_out, n2m_stream_t stream);

/* MLP backward: denc_tiles (same tile layout as enc_tiles, fp16, loss-scaled), g_mlp flat fp32 [7648] += */
int n2m_s0_mlp_bwd(const n2m_s0_params* p, const void* enc_tiles, const void* dout, const int32_t* counters,
                   uint32_t Mcap, const void* wpack, void* denc_tiles, float* g_mlp, const float* loss_scale,
                   n2m_stream_t stream);

/* scatter denc into gtable */
int n2m_s0_encode_bwd(const n2m_s0_params* p, const void* recs, const int32_t* counters, uint32_t Mcap,
                      const float* rays_o, const float* rays_d, const void* denc_tiles, const void* table,
                      const int32_t* offsets, void* gtable, const float* loss_scale, n2m_stream_t stream);

/* Ray-range parts.  The stages between march and optimizer can be run on `nparts` (1, 2, 4 or 8) contiguous ray ranges
 * of the batch -- part k covers rays [N*k/nparts, N*(k+1)/nparts) and their (contiguous, ray-ordered) samples -- so that
 * independent chains  gather -> MLP -> composite -> MLP backward -> scatter  of different parts can be in flight on
 * different streams: the latency-bound tensor-core MLP kernels of one part then share the SMs with the memory-bound
 * gather / scatter kernels of another.  A 128-sample tile that straddles a part boundary is computed by both parts,
 * each one reading / writing only its own rows (rows are masked by the [lo, hi) sample range taken from counters[4..12]);
 * losses, weight gradients and table gradients accumulate atomically, so the union of all parts equals the un-split call
 * up to fp32 summation order.  The entry points above are the nparts == 1 case. */
int n2m_s0_encode_fwd_part(const n2m_s0_params* p, const void* recs, const int32_t* counters, uint32_t Mcap,
                           const float* rays_o, const float* rays_d, const void* table, const int32_t* offsets,
                           void* enc_tiles, void* gtable, const float* loss_scale, uint32_t part, uint32_t nparts,
                           n2m_stream_t stream);
int n2m_s0_mlp_fwd_part(const n2m_s0_params* p, const void* enc_tiles, const int32_t* counters, uint32_t Mcap,
                        const void* wpack, void* out, float* spec_sq_sum, uint32_t part, uint32_t nparts, n2m_stream_t stream);
int n2m_s0_composite_loss_part(const n2m_s0_params* p, const void* out, const void* recs, const int32_t* rays,
                               const int32_t* counters, uint32_t N, uint32_t Mcap, const float* gt, const float* bg,
                               const float* loss_scale, void* dout, float* image, float* weights_sum, float* depth,
                               float* loss_out, uint32_t part, uint32_t nparts, n2m_stream_t stream);
int n2m_s0_mlp_bwd_part(const n2m_s0_params* p, const void* enc_tiles, const void* dout, const int32_t* counters,
                        uint32_t Mcap, const void* wpack, void* denc_tiles, float* g_mlp, const float* loss_scale,
                        uint32_t part, uint32_t nparts, n2m_stream_t stream);
int n2m_s0_encode_bwd_part(const n2m_s0_params* p, const void* recs, const int32_t* counters, uint32_t Mcap,
                           const float* rays_o, const float* rays_d, const void* denc_tiles, const void* table,
                           const int32_t* offsets, void* gtable, const float* loss_scale, uint32_t part, uint32_t nparts,
                           n2m_stream_t stream);

/* Fused backward (csrc/fused.cu): MLP backward (tcgen05) + hash-grid scatter of one part in ONE persistent, warp-specialised launch --
 * warps 0-3 run the MLP backward of a 128-sample tile, warps 4-19 scatter the previous tile's feature gradients, which are handed over
 * through a double-buffered shared-memory image instead of `denc_tiles` in HBM.  Same arithmetic as n2m_s0_mlp_bwd_part followed by
 * n2m_s0_encode_bwd_part (gradients equal up to fp32 atomic order); the TV gradient stays with n2m_s0_tv.  n2m_s0_fused_init sets the
 * kernel attributes once per process. */
int n2m_s0_fused_init(void);
/* profiling hook: bit 0 = scatter warps skip their REDs, bit 1 = MLP warps skip the tensor-core rounds (results meaningless) */
int n2m_s0_set_fused_debug(int mode);
int n2m_s0_bwd_fused_part(const n2m_s0_params* p, const void* enc_tiles, const void* dout, const void* recs, const int32_t* counters,
                          uint32_t Mcap, const float* rays_o, const float* rays_d, const void* wpack, const int32_t* offsets,
                          void* gtable, float* g_mlp, float* loss_scale, uint32_t part, uint32_t nparts, n2m_stream_t stream);

/* optimizer state block (device, float[8]): [0] loss_scale, [1] growth_tracker, [2] adam step t,
 * [3] found_inf, [4] lr (host-written each step), [5] 1-beta1^t, [6] sqrt(1-beta2^t), [7] 1/loss_scale.
 * Every `loss_scale` pointer argument above is the base of this block: the kernels read [0] and set [3]
 * when they see a non-finite gradient (instead of GradScaler's separate unscale_/inf-check pass). */

/* Adam (betas 0.9/0.999, eps) on tables + MLP; unscales by loss_scale, skips everything when found_inf,
 * refreshes the fp16 working copies (table, wpack), zeroes gtable / g_mlp, then updates the scaler. */
int n2m_s0_adam(void* table, void* color_master, void* gtable, float* m_table, float* v_table, uint32_t rows,
                float* mlp_params, float* g_mlp, float* m_mlp, float* v_mlp, void* wpack,
                float* opt_state, float eps, n2m_stream_t stream);

/* the same stage as four launches: head (MLP-gradient inf scan + step constants), tables, mlp (+ weight repack), post
 * (GradScaler update).  `tables` and `mlp` only read opt_state and touch disjoint buffers: a host may run them on two
 * streams between head and post (nerf2mesh_b200/stage0.py does). */
int n2m_s0_adam_head(const float* g_mlp, float* opt_state, n2m_stream_t stream);
int n2m_s0_adam_tables(void* table, void* color_master, void* gtable, float* m_table, float* v_table, uint32_t rows,
                       const float* opt_state, float eps, n2m_stream_t stream);
/* `tables` without zeroing the gradient rows: the host zeroes that gradient table on a side stream under the next step's forward pass
 * (which accumulates into the other parity table) */
int n2m_s0_adam_tables_keep(void* table, void* color_master, const void* gtable, float* m_table, float* v_table, uint32_t rows,
                            const float* opt_state, float eps, n2m_stream_t stream);
int n2m_s0_adam_mlp(float* mlp_params, float* g_mlp, float* m_mlp, float* v_mlp, void* wpack, const float* opt_state, float eps,
                    n2m_stream_t stream);
int n2m_s0_adam_post(float* opt_state, n2m_stream_t stream);

/* Fused forward (csrc/fused.cu): hash-grid gather + MLP forward of the WHOLE batch (nparts == 1) in one persistent, warp-specialised
 * launch -- two gather groups of four warps fill double-buffered tile images in shared memory, warps 0-3 run the tensor-core MLP rounds on
 * them; a copy of every image is stored to enc_tiles by the TMA unit for the backward pass.  Same arithmetic as n2m_s0_encode_fwd followed
 * by n2m_s0_mlp_fwd (bit-identical enc_tiles / out). */
int n2m_s0_fwd_fused(const n2m_s0_params* p, const void* recs, const int32_t* counters, uint32_t Mcap, const float* rays_o,
                     const float* rays_d, const void* table, const int32_t* offsets, const void* wpack, void* enc_tiles, void* out,
                     float* spec_sq_sum, n2m_stream_t stream);

/* Evaluation renderer with the alive-ray bookkeeping on the device (csrc/render.cu) = NeRFRenderer.render, inference branch
 * (nerf/renderer.py:749-802: per round march_rays -> model -> composite_rays -> mask compaction, one host read-back per round).
 *   render_begin : near / far (+ per-ray camera clamp, nullable), zeroed weights_sum / depth / image [N], rays_t, the first alive list;
 *                  alive [2 N] i32, ctl [16] i32 (control block: [1] sample rows of the round -- the counters the forward kernels read --,
 *                  [8] alive rays, [9] slab width, [10] survivors appended so far, [11] list parity, [12] rounds run, [13] sample rows evaluated)
 *   render_rounds: `num_rounds` rounds of plan -> march (records {t, dt, t + dt, ray} in slab order) -> n2m_s0_encode_fwd -> n2m_s0_mlp_fwd
 *                  -> slab compositor (raymarching.cu:842-924) + survivor compaction, all sizes read on the device; `schedule` (HOST
 *                  array) = slab widths, clipped on the device to Mcap / alive; recs [Mcap] float4, enc_tiles [Mcap*64] half,
 *                  out [Mcap] float4; Mcap a multiple of 128 and >= N.  Afterwards ctl[10] = rays still alive.
 *   render_finish: image += (1 - weights_sum) * bg (renderer.py:804), bg [N,3] or NULL (then bg_scalar). */
int n2m_s0_render_begin(const n2m_s0_params* p, const float* rays_o, const float* rays_d, const float* aabb, const float* cam_near_far,
                        uint32_t N, float* rays_t, float* rays_far, int32_t* alive, int32_t* ctl, float* weights_sum, float* depth,
                        float* image, n2m_stream_t stream);
int n2m_s0_render_rounds(const n2m_s0_params* p, const float* rays_o, const float* rays_d, const uint8_t* bitfield, uint32_t N,
                         const uint32_t* schedule, uint32_t num_rounds, float* rays_t, float* rays_far, int32_t* alive, int32_t* ctl,
                         void* recs, void* enc_tiles, void* out, uint32_t Mcap, const void* table, const int32_t* offsets, const void* wpack,
                         float* weights_sum, float* depth, float* image, n2m_stream_t stream);
int n2m_s0_render_finish(float* image, const float* weights_sum, const float* bg, float bg_scalar, uint32_t N, n2m_stream_t stream);

/* EMA of the parameters = torch_ema.ExponentialMovingAverage as the reference's Trainer holds it (nerf/utils.py:544-545, decay 0.95
 * from main.py:241): `update` once per EPOCH (utils.py:1213-1214), parameters swapped with the shadow for evaluation
 * (utils.py:1250-1252,1340-1341) and for the 'best' checkpoint (utils.py:1389-1401).  shadow_density [rows] f32, shadow_color [rows] float2,
 * shadow_mlp [7648] f32.  ema_update: shadow -= one_minus_decay * (shadow - param).  ema_swap: params <-> shadow in place, fp16 working
 * copies (table colour half2, wpack) refreshed. */
int n2m_s0_ema_update(const void* table, const void* color_master, const float* mlp_params, float* shadow_density, void* shadow_color,
                      float* shadow_mlp, uint32_t rows, float one_minus_decay, n2m_stream_t stream);
int n2m_s0_ema_swap(void* table, void* color_master, float* mlp_params, float* shadow_density, void* shadow_color, float* shadow_mlp,
                    uint32_t rows, void* wpack, n2m_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel optimizer fused with its collective over NVLink peer memory (csrc/dp.cu).
 * One process per GPU; buffers of the other ranks are mapped with CUDA IPC.
 * ---------------------------------------------------------------------------------------- */
int n2m_ipc_export(const void* ptr, void* handle_out /* 64 bytes */, uint64_t* offset_out);
int n2m_ipc_open(const void* handle, void** base_out);
int n2m_ipc_close(void* base);
uint32_t n2m_dp_ctx_bytes(void);
/* host image of the device context: per-peer pointers to gradient tables (two parities), working tables, MLP
 * gradient vectors, optimizer state blocks and flag arrays (>= 16 uint32 each, zero-initialised) */
int n2m_dp_ctx_fill(void* host_ctx, uint32_t world, uint32_t rank, uint32_t rows, uint32_t n_mlp,
                    void* const* gtab0, void* const* gtab1, void* const* table, void* const* gmlp0, void* const* gmlp1,
                    void* const* opt, void* const* flags, void* epoch);
int n2m_dp_barrier(const void* ctx, n2m_stream_t stream);
/* barrier -> reduce-scatter + Adam + all-gather on this rank's row slice -> MLP -> zero next-parity grads -> barrier.
 * color_master / m / v are slice-sized: ceil(rows / world) rounded up to a multiple of 4 rows. */
int n2m_dp_adam(const void* ctx, uint32_t parity, uint32_t world, uint32_t rows, uint32_t n_mlp, void* color_master_slice,
                float* m_slice, float* v_slice, float* mlp_params, float* m_mlp, float* v_mlp, void* wpack,
                void* gtab_next, float* gmlp_next, float* opt_state, float eps, n2m_stream_t stream);

/* NVLS variant: the table gradients are reduced INSIDE the NVSwitch (multimem.ld_reduce on a multicast address that maps the gradient
 * table of every rank) and the refreshed 8-byte entries are broadcast with one multimem.st; mc_gtab / mc_table are the multicast addresses
 * (torch.distributed._symmetric_memory) of this parity's gradient table and of the working table.  Per rank and step the links carry 16 B x rows out
 * (the switch pulls every replica of a row once) against 16 B x rows x (W-1)/W each way with P2P loads, and the all-gather 8 B x rows / W
 * out instead of 8 B x rows x (W-1)/W: slower than n2m_dp_adam at W = 2, faster at W = 8 (profiles/r2_scaling.md).  In both entry points gtab_next / gmlp_next may be NULL: the caller then zeroes the next-parity gradient buffers itself. */
int n2m_dp_adam_nvls(const void* ctx, const void* mc_gtab, void* mc_table, uint32_t parity, uint32_t world, uint32_t rows, uint32_t n_mlp,
                     void* color_master_slice, float* m_slice, float* v_slice, float* mlp_params, float* m_mlp, float* v_mlp, void* wpack,
                     void* gtab_next, float* gmlp_next, float* opt_state, float eps, n2m_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* N2M_B200_FUSED_H */
