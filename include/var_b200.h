/* var_b200.h -- C ABI of libvar_b200.so: the B200 (sm_100a) implementation of the
 * VoiceControlledRobot-VAR hot path (triplet training step + batched reward query).
 *
 * The reference (pure Python over torch / torchaudio, no native code, no FFI) exposes
 * Python-level hooks only; each entry point below names the reference code it replaces
 * (paths relative to the reference repository root).  Conventions:
 *   - every pointer named d_* / documented "device" is a CUDA device pointer;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream ordered, nothing
 *     synchronises, nothing allocates device memory except *_create;
 *   - return value 0 = ok, negative = error (VAR_ERR_*); var_last_error() gives the text;
 *   - no function ever falls back to a CPU implementation.
 */
#ifndef VAR_B200_H_
#define VAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAR_B200_VERSION 100

#define VAR_ERR_ARG -1
#define VAR_ERR_CUDA -2
#define VAR_ERR_UNSUPPORTED -3
#define VAR_ERR_WORKSPACE -4

int var_version(void);
const char* var_last_error(void);

/* ------------------------------------------------------------------------------------
 * Audio front-end.  Replaces audioLoader.get_mfcc (Envs/audioLoader.py:147-164) followed
 * by processSoundFeat (Envs/audioLoader.py:241-252) for a whole batch of clips.
 *   flavour 0: torchaudio.transforms.MFCC as configured at Envs/audioLoader.py:150-157
 *              (int16 -> /32768, centre reflect pad, periodic Hamming, |rFFT|^2, HTK mel,
 *              log(x + 1e-6), ortho DCT-II), fp32.
 *   flavour 1: python_speech_features.mfcc as called at Envs/audioLoader.py:159-161
 *              (pre-emphasis on raw int16 scale, symmetric Hamming, |rFFT|^2/nfft, bin-index
 *              triangles, ln, ortho DCT, lifter 22, c0 := ln(energy)), fp32 output.
 * d_offsets[b] < 0 marks the "empty" class: the row block is all zeros (dataset.py:37-43).
 * Output [B, F, 40] fp32: frames beyond the clip are zero (pad), frames beyond F are dropped
 * (crop).
 * ---------------------------------------------------------------------------------- */
int var_mfcc_plan_create(int flavour, int fs, int n_fft, int win_length, int hop, void** plan);
int var_mfcc_plan_destroy(void* plan);
int var_mfcc_num_frames(void* plan, int n_samples);
int var_mfcc_fwd(void* plan, const int16_t* d_wav, const int64_t* d_offsets, const int32_t* d_lengths,
                 int B, int F, float* d_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Triplet sampler.  Replaces the integer sampling of VARDataset.__getitem__
 * (dataset.py:64-89, :34-62), audioLoader.getAudioSamples (Envs/audioLoader.py:166-177) and
 * DataLoader(shuffle=True, num_workers=0) batch order (dataset.py:157-162), bit-exactly, on a
 * device-resident mt19937 that advances like torch's global CPU generator.
 * d_state: uint32[625].
 * ---------------------------------------------------------------------------------- */
int var_sampler_seed(uint32_t* d_state, uint64_t seed, void* stream);
int var_sampler_epoch(uint32_t* d_state, int n_items, int32_t* d_perm, void* stream);
int var_sampler_batch(uint32_t* d_state, int B, int task_num, const int32_t* d_items,
                      const int32_t* d_gt, const int32_t* d_stored_sn, const int32_t* d_nds,
                      const int32_t* d_nclips, const int32_t* d_clip_base, int max_ds,
                      const int64_t* d_clip_off, const int32_t* d_clip_len, int32_t* d_scratch,
                      int32_t* d_out_item, int32_t* d_out_gt, int32_t* d_out_sn, int32_t* d_out_rec,
                      int64_t* d_out_off, int32_t* d_out_len, void* stream);

/* The same batch draw for the iTHOR configuration: VARDataset resolves ground truth / negative class
 * through its task list (dataset.py:17-29, :37-53) and audioLoader.getAudioFromTask draws a location
 * synonym, an object synonym (Envs/audioLoader.py:223-228) and then a clip of the resolved
 * (location, object, action) list (genSoundFeatFromTask, Envs/audioLoader.py:203-209): three draws per
 * sound.  d_n_loc_syn / d_n_obj_syn: [task_num] synonym counts of each task's location / object
 * (config.synonym, Envs/ai2thor/env_config.py:35-42); d_nclips / d_clip_base:
 * [task_num, max_loc_syn, max_obj_syn] clip count / first clip id of the resolved list.
 * out_rec rows are (task, li * max_obj_syn + oi, clip) for the positive then the negative. */
int var_sampler_batch_tasks(uint32_t* d_state, int B, int task_num, const int32_t* d_items,
                            const int32_t* d_gt, const int32_t* d_stored_sn, const int32_t* d_n_loc_syn,
                            const int32_t* d_n_obj_syn, const int32_t* d_nclips, const int32_t* d_clip_base,
                            int max_loc_syn, int max_obj_syn, const int64_t* d_clip_off,
                            const int32_t* d_clip_len, int32_t* d_scratch, int32_t* d_out_item,
                            int32_t* d_out_gt, int32_t* d_out_sn, int32_t* d_out_rec, int64_t* d_out_off,
                            int32_t* d_out_len, void* stream);
/* Continue a HOST mt19937 stream on the device: words[624] + read position (624 = twist before the
 * next draw), i.e. torch.get_rng_state() of the global CPU generator the reference's
 * DataLoader / __getitem__ draw from (dataset.py:76, :157-162). */
int var_sampler_set_state(uint32_t* d_state, const uint32_t* host_words, int pos, void* stream);

/* Host-side staging helpers of the streaming triplet loader (datasets kept in pinned HOST memory, the
 * counterpart of the reference's DataLoader workers, dataset.py:157-162): gather the rows (uint8 frames) /
 * clips (int16, packed back to back and 4-byte aligned; offset < 0 = empty class) a batch draws into a
 * pinned staging buffer with `nthreads` memcpy threads.  Host pointers only; no device work.
 * var_host_gather_clips returns the number of int16 written (or a negative VAR_ERR_*). */
int var_host_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, int n, void* dst, int nthreads);
int64_t var_host_gather_clips(const int16_t* arena, const int64_t* offsets, const int64_t* lengths, int n, int16_t* dst,
                              int64_t* new_offsets, int nthreads);

/* ------------------------------------------------------------------------------------
 * Encoders.  A net object holds the layer plan of one VARPretextNet
 *   kind 0: models/pretext/arm_pretext_model.py:37-59   (Kuka)
 *   kind 1: models/pretext/ai2thor_pretext_model.py:41-64 (iTHOR, incl. the bidirectional GRU)
 * with all parameters in ONE flat fp32 buffer in the engine's packed layout (caller-allocated,
 * var_net_param_floats() floats each for: master weights, tf32 operand copy, gradients).
 * var_net_tensor_* describe / convert the reference state_dict tensors (checkpoint compat,
 * pretext.py:102-111, VAR/pretext_VAR.py:75-80).
 * ---------------------------------------------------------------------------------- */
int var_net_create(int kind, int sound_frames, int rep_dim, void** net);
int var_net_destroy(void* net);
int64_t var_net_param_floats(void* net);
int var_net_num_tensors(void* net);
/* name: state_dict key; shape[4]/ndim: reference shape; offset/packed: location in the flat buffer */
int var_net_tensor_info(void* net, int index, char* name, int name_cap, int* ndim, int* shape,
                        int64_t* offset, int64_t* packed_floats);
/* (also allocates the net's own device scratch: f16 copies of the weights of the 16-bit conv region, 0.6 MB) */
int var_net_bind(void* net, float* d_params, float* d_params_mma, float* d_grads);
/* reference layout (contiguous fp32, device) -> packed master + tf32 copy */
int var_net_load_tensor(void* net, int index, const float* d_src, void* stream);
/* packed (which = 0 params, 1 grads) -> reference layout */
int var_net_store_tensor(void* net, int index, int which, float* d_dst, void* stream);
int var_net_refresh_mma(void* net, void* stream);
int64_t var_net_workspace_bytes(void* net, int n_images, int n_sounds, int train);
int var_net_raw_dims(void* net, int* img_raw, int* snd_raw);
/* on = 0 serialises the image and sound branches on the caller's stream (default: the sound branch runs on a
 * high-priority side stream so the branches overlap); used by bench.py for uncontended per-kernel timings. */
int var_net_set_overlap(void* net, int on);

/* Data-parallel gradient buckets (no reference counterpart: the reference is single-GPU).  Bucket 0 is
 * the contiguous range of the flat gradient buffer that holds every recurrent-layer gradient (rnn.*, iTHOR
 * net only: 77 % of the gradient bytes); it is final half way through the backward pass.
 * var_net_set_bucket_event registers a caller-owned cudaEvent_t (null = none) that the backward pass
 * records at that point, on the stream that produced the gradients, so the caller's all-reduce of the
 * range can run under the rest of the backward pass.  VAR_ERR_UNSUPPORTED for nets without the bucket. */
int var_net_grad_bucket(void* net, int bucket, int64_t* offset_floats, int64_t* count_floats);
int var_net_set_bucket_event(void* net, int bucket, void* cuda_event);

/* VARPretextNet.forward (models/pretext/pretext_base.py:10-42) for a batch.
 *   d_images : [n_images, 3, 96, 96] NCHW; image_kind 0 = uint8 (scaled by 1/255 in the first
 *              conv's loader, dataset.py:67-68 / vec_pretext_normalize.py:85), 1 = fp32.
 *   d_sounds : [n_sounds, F, 40] fp32 MFCC features (positives then negatives when both).
 * Either may be null with count 0.  Outputs (nullable): L2-normalised embeddings
 * [n, rep_dim]; raw features in the reference's NCHW-flatten order (image_feat_raw,
 * pos_sound_raw).  train != 0 keeps activations in the workspace for var_net_backward. */
int var_net_forward(void* net, const void* d_images, int image_kind, int n_images,
                    const float* d_sounds, int n_sounds, void* d_ws, int64_t ws_bytes, int train,
                    float* d_img_feat, float* d_img_raw, float* d_snd_feat, float* d_snd_raw,
                    void* stream);
/* Backward of the last var_net_forward(train=1) on the same workspace; gradients are
 * ACCUMULATED into the bound gradient buffer (zero it first). d_*_dfeat: [n, rep_dim] or null. */
int var_net_backward(void* net, const float* d_img_dfeat, const float* d_snd_dfeat, void* d_ws,
                     int64_t ws_bytes, void* stream);
/* One fused triplet step (VAR/pretext_VAR.py:56-68 minus the optimizer): forward of B images and
 * 2B sounds, fused head + F.normalize + TripletMarginLoss(margin, p=2) + its gradient
 * (one warp per triplet), full backward.  d_loss (+=) receives sum(hinge)/loss_denominator;
 * gradients are scaled by 1/loss_denominator (use the GLOBAL batch under data parallelism). */
int var_net_triplet_step(void* net, const void* d_images, int image_kind, const float* d_sounds,
                         int B, float margin, float loss_denominator, void* d_ws, int64_t ws_bytes,
                         float* d_loss, float* d_feats /* [3, B, rep_dim] or null */, void* stream);
/* Batched VAR reward query (Envs/vec_env/vec_pretext_normalize.py:82-101) for N envs:
 * image branch (+ goal-sound branch when d_goal_sounds != null, else d_goal_feat_cached is
 * the cached embedding, pretext_base.py:29-32), then ONE launch computing both embeddings'
 * normalisation, img_sound_dot and reward = dot + env_reward. */
int var_net_reward(void* net, const void* d_images, int image_kind, const float* d_goal_sounds,
                   const float* d_goal_feat_cached, const float* d_env_reward, int N, void* d_ws,
                   int64_t ws_bytes, float* d_img_feat, float* d_goal_feat, float* d_dot,
                   float* d_reward, void* stream);

/* Reward post-processing of VecPretextNormalize.step_wait (vec_pretext_normalize.py:53-59,
 * running_mean_std.py:16-35) on the device, in float64 like the reference: orig = rew;
 * ret = ret*gamma + rew; RunningMeanStd update with the batch moments of ret; out =
 * clip(rew / sqrt(var + eps), +-cliprew); ret[done] = 0.  d_ret: double[N]; d_rms: double[3] =
 * {mean, var, count} (initialise to {0, 1, 1e-4}).  update_rms = 0 passes rew through unscaled
 * (the wrapper's ret=False mode). */
int var_reward_normalize(const float* d_rew, const uint8_t* d_done, int N, double* d_ret, double* d_rms,
                         double gamma, double eps, double cliprew, int update_rms, float* d_orig,
                         float* d_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Optimizer.  torch.optim.Adam(lr, weight_decay) exactly as configured at
 * VAR/pretext_VAR.py:33-35 (L2 folded into the gradient, default betas/eps), on the flat
 * buffer; also refreshes the tf32 operand copy.  grad_scale multiplies the gradient first
 * (1.0, or 1/world_size when averaging).  n must be a multiple of 4.
 * ---------------------------------------------------------------------------------- */
int var_adam_step(float* d_params, const float* d_grads, float* d_m, float* d_v, float* d_params_mma,
                  int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                  int64_t step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------
 * Stand-alone operators (used by the parity tests; the net calls the same launchers).
 * Activations NHWC fp32 holding tf32-representable values; weights packed [Cout, Kpad],
 * k = (r*S + s)*Cin + c, Kpad = round_up(K, 32).  Replace torch.nn.Conv2d / Linear / MaxPool2d
 * forward+backward (cuDNN/cuBLAS) of models/pretext/{arm,ai2thor}_pretext_model.py.
 * ---------------------------------------------------------------------------------- */
int var_pack_weight(const float* d_ref_oihw, float* d_packed, float* d_packed_mma, int Cout, int Cin,
                    int R, int S, int kpad, void* stream);
int var_unpack_weight(const float* d_packed, float* d_ref_oihw, int Cout, int Cin, int R, int S,
                      int kpad, void* stream);
/* src_kind 0: NHWC fp32 (Cin % 32 == 0); 1: strided fp32; 2: strided uint8 (x scale).
 * strides[4] = element strides (N, H, W, C) for kinds 1/2. */
int var_conv2d_fwd(const void* d_x, int src_kind, const int64_t* strides, float scale, int N, int H,
                   int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw,
                   const float* d_w_packed, const float* d_bias, float* d_y, int relu, int round_out,
                   void* stream);
int var_conv2d_dgrad(const float* d_dy, const float* d_w_packed, float* d_dx, const float* d_mask,
                     int N, int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph,
                     int pw, int round_out, void* stream);
int var_conv2d_wgrad(const void* d_x, int src_kind, const int64_t* strides, float scale,
                     const float* d_dy, float* d_dw_packed, float* d_db, int N, int H, int W, int Cin,
                     int Cout, int R, int S, int sh, int sw, int ph, int pw, void* stream);
/* 16-bit operand forms of the same three conv passes (Cin % 64 == 0, Cout % 64 == 0, not 1x1): activations
 * and packed weights stored as IEEE f16 (the same 10-bit mantissa as tf32 at half the bytes per MAC --
 * what bounds the N = 64 convs), tcgen05.mma kind::f16, fp32 accumulation.  Gradients travel through the
 * region as f16 times a power-of-two scale picked on the device from the largest incoming magnitude
 * (var_grad_to_f16_scaled: d_scale[0] = S, d_scale[1] = 1/S); kernels multiply by d_out_scale /
 * d_inv_scale where a gradient leaves the region.  out_kind: 0 fp32, 1 f16.  mask_kind: 0 fp32, 1 f16.
 * var_cvt_f16: packed fp32 weights (or any fp32 array, n % 4 == 0) -> f16 copy. */
int var_cvt_f16(const float* d_src, void* d_dst, int64_t n, void* stream);
int var_grad_to_f16_scaled(const float* d_src, void* d_dst, int64_t n, float* d_scale, uint32_t* d_amax_scratch,
                           void* stream);
int var_conv2d_fwd_h16(const void* d_x_f16, int N, int H, int W, int Cin, int Cout, int R, int S, int sh, int sw,
                       int ph, int pw, const void* d_w_f16_packed, const float* d_bias, void* d_y, int out_kind,
                       int relu, int round_out, void* stream);
int var_conv2d_dgrad_h16(const void* d_dy_f16, const void* d_w_f16_packed, void* d_dx, int out_kind,
                         const void* d_mask, int mask_kind, const float* d_out_scale, int N, int H, int W, int Cin,
                         int Cout, int R, int S, int sh, int sw, int ph, int pw, int round_out, void* stream);
int var_conv2d_wgrad_h16(const void* d_x_f16, const void* d_dy_f16, float* d_dw_packed, float* d_db,
                         const float* d_inv_scale, int N, int H, int W, int Cin, int Cout, int R, int S, int sh,
                         int sw, int ph, int pw, void* stream);
/* Plain GEMMs of the 16-bit region (GRU input projection, replaces nn.GRU's x W_ih^T of models/pretext/
 * ai2thor_pretext_model.py:33 and its backward): out[M, N] (fp32, pitch ldo) = A[M, K] (f16, pitch lda) x W[N, K]^T (f16)
 * (+ bias) (* *d_out_scale) (* (mask > 0)); and d_dw[N][kpad] += *d_inv_scale * dY^T X over the M rows
 * (X f16 [M, K] pitch ldx, dY f16 [M, N] pitch ldy).  K % 64 == 0. */
int var_linear_h16(const void* d_a_f16, int64_t lda, const void* d_w_f16, const float* d_bias, float* d_out, int64_t ldo,
                   const void* d_mask, int mask_kind, int64_t ldm, const float* d_out_scale, int M, int K, int N,
                   int round_out, void* stream);
int var_linear_wgrad_h16(const void* d_x_f16, int64_t ldx, const void* d_dy_f16, int64_t ldy, float* d_dw, int kpad,
                         const float* d_inv_scale, int M, int K, int N, void* stream);
int var_maxpool2x2_fwd(const float* d_x, float* d_y, int N, int H, int W, int C, void* stream);
int var_maxpool2x2_bwd(const float* d_x, const float* d_dy, float* d_dx, int N, int H, int W, int C,
                       void* stream);
/* Fused head + triplet kernel on its own: h_* are the inputs of each head's last Linear. */
int var_triplet_fwd_bwd(const float* d_h_img, const float* d_h_pos, const float* d_h_neg, int B, int D,
                        int Kh_img, int Kh_snd, const float* d_W_img, const float* d_b_img,
                        const float* d_W_snd, const float* d_b_snd, float margin,
                        float loss_denominator, float* d_feats, float* d_loss, float* d_loss_rows,
                        float* d_dh_img, float* d_dh_pos, float* d_dh_neg, float* d_dW_img,
                        float* d_db_img, float* d_dW_snd, float* d_db_snd, void* stream);

/* ------------------------------------------------------------------------------------
 * Launch accounting (bench.py): number of kernels this library has launched so far, and a
 * per-kernel-family CUDA-event profiler (events recorded on the launching stream).
 * var_prof_end fills arrays of var_prof_num_tags() entries: total ms, algorithmic FLOPs of
 * the GEMM families, launch count.  Tag order: gemm_fwd, gemm_dgrad, gemm_scalar, gru_step,
 * wgrad, colsum, mfcc, tail, pool, adam, gru_cell_bwd, sampler, misc, gemm_fwd16, gemm_dgrad16, wgrad16
 * (the last three: the kind::f16 conv kernels).  var_h16_flags: bit 0 = 16-bit conv region on, bit 1 = the
 * recurrent kernels use f16 operands (decides which tensor peak a family is compared with).
 * ---------------------------------------------------------------------------------- */
long long var_launch_count(void);
/* Kernels launched by replaying a CUDA graph that was captured from this library's launches (the host mirror's
 * graphed training step / reward query): the capture counted them once, every replay reports them here. */
int var_launch_count_add(long long n);
int var_prof_begin(void);
int var_prof_end(double* ms, double* flops, long long* count, int ntags);
int var_prof_num_tags(void);
int var_h16_flags(void);

#ifdef __cplusplus
}
#endif
#endif /* VAR_B200_H_ */
