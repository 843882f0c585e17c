// Launchers of the sampler and the fused head / triplet kernel (triplet.cu).
#pragma once
#include "common.cuh"

namespace var {

constexpr int kSamplerStateWords = 625;  // mt19937 state[624] + pos

struct SamplerArgs {
  uint32_t* state;          // [625] device-resident generator (advanced in place)
  int B;                    // triplets in this batch
  int task_num;             // config.taskNum; class task_num is the "empty" class
  const int* items;         // [B] dataset indices of this batch (epoch permutation slice) or null = 0..B-1
  const int* gt;            // [n_items] ground-truth class per dataset item, in 0..task_num
  const int* stored_sn;     // [n_items] stored sound_negative_id or null (draw it)
  const int* nds;           // [task_num] number of sound datasets holding this intent
  const int* nclips;        // [task_num * max_ds] clips per (intent, dataset)
  const int* clip_base;     // [task_num * max_ds] first clip id of (intent, dataset)
  int max_ds;
  const int* nds2;          // iTHOR task tables: [task_num] object-synonym count (null = pybullet tables)
  int max_ds2;              //   row pitch of the object-synonym index inside a task's max_ds slots
  const long long* clip_off;  // [n_clips] sample offset of each clip in the int16 arena
  const int* clip_len;        // [n_clips] samples
  int* scratch_off;         // [B] workspace
  // outputs
  int* out_item;            // [B]
  int* out_gt;              // [B]
  int* out_sn;              // [B] negative class actually used
  int* out_rec;             // [B, 6] (intent, dataset, clip) of positive then negative, -1 = zero feature
  long long* out_off;       // [2B] arena offsets: positives then negatives (-1 = zero feature)
  int* out_len;             // [2B]
  int chunk;                // set by the launcher
};

int sampler_seed(uint32_t* state, unsigned long long seed, cudaStream_t st);
int sampler_set_state(uint32_t* state, const uint32_t* host_words, int pos, cudaStream_t st);
int sampler_epoch(uint32_t* state, int n, int* perm, cudaStream_t st);
int sampler_batch(const SamplerArgs& a, cudaStream_t st);

enum TailMode : int { TAIL_FWD = 0, TAIL_TRIPLET = 1, TAIL_BWD = 2, TAIL_REWARD = 3 };

struct TailArgs {
  int mode, B, D, Kh_img, Kh_snd;
  const float* h_img;   // [B, Kh_img] input of the last image-head Linear (post ReLU) or null
  const float* h_pos;   // [B, Kh_snd]
  const float* h_neg;   // [B, Kh_snd]
  const float* W_img; const float* b_img;  // [D, Kh_img], [D]
  const float* W_snd; const float* b_snd;  // [D, Kh_snd], [D]
  float* feat_img; float* feat_pos; float* feat_neg;  // [B, D] normalised embeddings (nullable)
  // TAIL_TRIPLET
  float margin, grad_scale, loss_scale;
  float* loss;          // scalar accumulator (+= sum * loss_scale)
  float* loss_rows;     // [B] per-triplet hinge (nullable)
  // TAIL_BWD
  const float* dfeat_img; const float* dfeat_pos; const float* dfeat_neg;  // [B, D] (nullable)
  // backward outputs
  float* dh_img; float* dh_pos; float* dh_neg;  // ReLU-masked, tf32-rounded
  float* dW_img; float* db_img; float* dW_snd; float* db_snd;  // accumulated (+=)
  // TAIL_REWARD
  const float* goal_feat_in;  // [B, D] cached goal-sound embedding (used when h_pos is null)
  const float* env_reward;    // [B] or null
  float* dot_out;             // [B] img_sound_dot
  float* reward_out;          // [B]
};
int tail_launch(const TailArgs& a, cudaStream_t st);

}  // namespace var
