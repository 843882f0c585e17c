// extern "C" surface of libvar_b200.so that is not tied to a net object
// (include/var_b200.h).  Thin argument marshalling only.
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../../include/var_b200.h"
#include "aux_kernels.cuh"
#include "engine_host.cuh"
#include "mfcc.cuh"
#include "triplet.cuh"

#define ST(s) (reinterpret_cast<cudaStream_t>(s))

using namespace var;

static ConvShape make_shape(int N, int H, int W, int Cin, int Cout, int R, int S, int sh, int sw,
                            int ph, int pw) {
  ConvShape c{N, H, W, Cin, Cout, R, S, sh, sw, ph, pw, 0, 0};
  c.P = (H + 2 * ph - R) / sh + 1;
  c.Q = (W + 2 * pw - S) / sw + 1;
  return c;
}
static SrcLayout make_layout(const int64_t* strides, float scale) {
  SrcLayout sl{0, 0, 0, 0, scale};
  if (strides) { sl.sN = strides[0]; sl.sH = strides[1]; sl.sW = strides[2]; sl.sC = strides[3]; }
  return sl;
}

extern "C" {

int var_version(void) { return VAR_B200_VERSION; }

int var_mfcc_plan_create(int flavour, int fs, int n_fft, int win_length, int hop, void** plan) {
  if (!plan) return VAR_ERR_ARG;
  MfccPlan* p = nullptr;
  const int rc = mfcc_plan_create(flavour, fs, n_fft, win_length, hop, &p);
  if (rc) return rc;
  *plan = p;
  return VAR_OK;
}
int var_mfcc_plan_destroy(void* plan) {
  mfcc_plan_destroy(reinterpret_cast<MfccPlan*>(plan));
  return VAR_OK;
}
int var_mfcc_num_frames(void* plan, int n_samples) {
  return mfcc_num_frames(reinterpret_cast<MfccPlan*>(plan), n_samples);
}
int var_mfcc_fwd(void* plan, const int16_t* wav, const int64_t* offsets, const int32_t* lengths, int B,
                 int F, float* out, void* stream) {
  if (!plan || !offsets || !lengths || !out) return VAR_ERR_ARG;
  return mfcc_fwd(reinterpret_cast<MfccPlan*>(plan), wav,
                  reinterpret_cast<const long long*>(offsets), lengths, B, F, out, ST(stream));
}

int var_sampler_seed(uint32_t* state, uint64_t seed, void* stream) {
  if (!state) return VAR_ERR_ARG;
  return sampler_seed(state, seed, ST(stream));
}
int var_sampler_epoch(uint32_t* state, int n_items, int32_t* perm, void* stream) {
  if (!state || !perm) return VAR_ERR_ARG;
  return sampler_epoch(state, n_items, perm, ST(stream));
}
int var_sampler_batch(uint32_t* state, int B, int task_num, const int32_t* items, const int32_t* gt,
                      const int32_t* stored_sn, const int32_t* nds, const int32_t* nclips,
                      const int32_t* clip_base, int max_ds, const int64_t* clip_off,
                      const int32_t* clip_len, int32_t* scratch, int32_t* out_item, int32_t* out_gt,
                      int32_t* out_sn, int32_t* out_rec, int64_t* out_off, int32_t* out_len,
                      void* stream) {
  if (!state || !gt || !nds || !nclips || !clip_base || !clip_off || !clip_len || !scratch ||
      !out_item || !out_gt || !out_sn || !out_rec || !out_off || !out_len)
    return VAR_ERR_ARG;
  SamplerArgs a;
  memset(&a, 0, sizeof(a));
  a.state = state; a.B = B; a.task_num = task_num; a.items = items; a.gt = gt;
  a.stored_sn = stored_sn; a.nds = nds; a.nclips = nclips; a.clip_base = clip_base;
  a.max_ds = max_ds; a.clip_off = reinterpret_cast<const long long*>(clip_off); a.clip_len = clip_len;
  a.scratch_off = scratch; a.out_item = out_item; a.out_gt = out_gt; a.out_sn = out_sn;
  a.out_rec = out_rec; a.out_off = reinterpret_cast<long long*>(out_off); a.out_len = out_len;
  return sampler_batch(a, ST(stream));
}

int var_sampler_batch_tasks(uint32_t* state, int B, int task_num, const int32_t* items, const int32_t* gt,
                            const int32_t* stored_sn, const int32_t* n_loc_syn, const int32_t* n_obj_syn,
                            const int32_t* nclips, const int32_t* clip_base, int max_loc_syn, int max_obj_syn,
                            const int64_t* clip_off, const int32_t* clip_len, int32_t* scratch,
                            int32_t* out_item, int32_t* out_gt, int32_t* out_sn, int32_t* out_rec,
                            int64_t* out_off, int32_t* out_len, void* stream) {
  if (!state || !gt || !n_loc_syn || !n_obj_syn || !nclips || !clip_base || !clip_off || !clip_len || !scratch ||
      !out_item || !out_gt || !out_sn || !out_rec || !out_off || !out_len || max_loc_syn <= 0 || max_obj_syn <= 0)
    return VAR_ERR_ARG;
  SamplerArgs a;
  memset(&a, 0, sizeof(a));
  a.state = state; a.B = B; a.task_num = task_num; a.items = items; a.gt = gt;
  a.stored_sn = stored_sn; a.nds = n_loc_syn; a.nds2 = n_obj_syn; a.nclips = nclips; a.clip_base = clip_base;
  a.max_ds = max_loc_syn * max_obj_syn; a.max_ds2 = max_obj_syn;
  a.clip_off = reinterpret_cast<const long long*>(clip_off); a.clip_len = clip_len;
  a.scratch_off = scratch; a.out_item = out_item; a.out_gt = out_gt; a.out_sn = out_sn;
  a.out_rec = out_rec; a.out_off = reinterpret_cast<long long*>(out_off); a.out_len = out_len;
  return sampler_batch(a, ST(stream));
}
int var_sampler_set_state(uint32_t* state, const uint32_t* host_words, int pos, void* stream) {
  if (!state || !host_words || pos < 0 || pos > 624) return VAR_ERR_ARG;
  return sampler_set_state(state, host_words, pos, ST(stream));
}

// ---- host-side staging helpers of the streaming triplet loader (plain memcpy work, no device access) ----
static void run_parallel(int n, int nthreads, const std::function<void(int, int)>& fn) {
  if (nthreads <= 1 || n < 2 * nthreads) { fn(0, n); return; }
  std::vector<std::thread> th;
  const int per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    const int lo = t * per, hi = lo + per < n ? lo + per : n;
    if (lo < hi) th.emplace_back(fn, lo, hi);
  }
  for (auto& x : th) x.join();
}
int var_host_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, int n, void* dst, int nthreads) {
  if (!src || !idx || !dst || row_bytes <= 0 || n < 0) return VAR_ERR_ARG;
  const char* s_ = reinterpret_cast<const char*>(src);
  char* d_ = reinterpret_cast<char*>(dst);
  run_parallel(n, nthreads, [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) memcpy(d_ + (int64_t)i * row_bytes, s_ + idx[i] * row_bytes, (size_t)row_bytes);
  });
  return VAR_OK;
}
int64_t var_host_gather_clips(const int16_t* arena, const int64_t* offsets, const int64_t* lengths, int n, int16_t* dst,
                              int64_t* new_offsets, int nthreads) {
  if (!arena || !offsets || !lengths || !dst || !new_offsets || n < 0) return VAR_ERR_ARG;
  int64_t cur = 0;
  for (int i = 0; i < n; ++i) {  // destination offsets: clips packed back to back, 4-byte aligned
    if (offsets[i] < 0) { new_offsets[i] = -1; continue; }
    new_offsets[i] = cur;
    cur += lengths[i] + (lengths[i] & 1);
  }
  run_parallel(n, nthreads, [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i)
      if (offsets[i] >= 0) memcpy(dst + new_offsets[i], arena + offsets[i], (size_t)lengths[i] * 2);
  });
  return cur;
}

int var_adam_step(float* p, const float* g, float* m, float* v, float* p_mma, int64_t n, float lr,
                  float beta1, float beta2, float eps, float wd, int64_t step, float grad_scale,
                  void* stream) {
  if (!p || !g || !m || !v || step < 1) return VAR_ERR_ARG;
  return adam_step(p, g, m, v, p_mma, n, lr, beta1, beta2, eps, wd, step, grad_scale, ST(stream));
}

int var_reward_normalize(const float* rew, const uint8_t* done, int N, double* ret, double* rms, double gamma,
                         double eps, double cliprew, int update_rms, float* orig, float* out, void* stream) {
  if (!rew || !done || !ret || !rms || !out) return VAR_ERR_ARG;
  return reward_normalize(rew, done, N, ret, rms, gamma, eps, cliprew, update_rms, orig, out, ST(stream));
}

int var_pack_weight(const float* ref, float* packed, float* packed_mma, int Cout, int Cin, int R,
                    int S, int kpad, void* stream) {
  return pack_weight(ref, packed, packed_mma, Cout, Cin, R, S, kpad, ST(stream));
}
int var_unpack_weight(const float* packed, float* ref, int Cout, int Cin, int R, int S, int kpad,
                      void* stream) {
  return unpack_weight(packed, ref, Cout, Cin, R, S, kpad, ST(stream));
}

int var_conv2d_fwd(const void* x, int src_kind, const int64_t* strides, float scale, int N, int H,
                   int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw,
                   const float* w, const float* bias, float* y, int relu, int round_out, void* stream) {
  const ConvShape cs = make_shape(N, H, W, Cin, Cout, R, S, sh, sw, ph, pw);
  const SrcLayout sl = make_layout(strides, scale);
  return conv_fwd(cs, x, src_kind, &sl, w, bias, y, relu, round_out, ST(stream));
}
int var_conv2d_dgrad(const float* dy, const float* w, float* dx, const float* mask, int N, int H, int W,
                     int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw, int round_out,
                     void* stream) {
  const ConvShape cs = make_shape(N, H, W, Cin, Cout, R, S, sh, sw, ph, pw);
  return conv_dgrad(cs, dy, w, dx, mask, nullptr, round_out, ST(stream));
}
int var_conv2d_wgrad(const void* x, int src_kind, const int64_t* strides, float scale, const float* dy,
                     float* dw, float* db, int N, int H, int W, int Cin, int Cout, int R, int S, int sh,
                     int sw, int ph, int pw, void* stream) {
  const ConvShape cs = make_shape(N, H, W, Cin, Cout, R, S, sh, sw, ph, pw);
  const SrcLayout sl = make_layout(strides, scale);
  return conv_wgrad(cs, x, src_kind, &sl, dy, dw, db, ST(stream));
}
int var_cvt_f16(const float* src, void* dst, int64_t n, void* stream) {
  if (!src || !dst) return VAR_ERR_ARG;
  return cvt_f16(src, dst, n, ST(stream));
}
int var_grad_to_f16_scaled(const float* src, void* dst, int64_t n, float* scale, uint32_t* amax_scratch, void* stream) {
  if (!src || !dst || !scale || !amax_scratch) return VAR_ERR_ARG;
  return grad_to_f16_scaled(src, dst, n, scale, amax_scratch, ST(stream));
}
int var_conv2d_fwd_h16(const void* x, int N, int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph,
                       int pw, const void* w, const float* bias, void* y, int out_kind, int relu, int round_out,
                       void* stream) {
  const ConvShape cs = make_shape(N, H, W, Cin, Cout, R, S, sh, sw, ph, pw);
  return conv_fwd_h16(cs, x, w, bias, y, out_kind, relu, round_out, ST(stream));
}
int var_conv2d_dgrad_h16(const void* dy, const void* w, void* dx, int out_kind, const void* mask, int mask_kind,
                         const float* out_scale, int N, int H, int W, int Cin, int Cout, int R, int S, int sh, int sw,
                         int ph, int pw, int round_out, void* stream) {
  const ConvShape cs = make_shape(N, H, W, Cin, Cout, R, S, sh, sw, ph, pw);
  return conv_dgrad_h16(cs, dy, w, dx, out_kind, mask, mask_kind, out_scale, round_out, ST(stream));
}
int var_conv2d_wgrad_h16(const void* x, const void* dy, float* dw, float* db, const float* inv_scale, int N, int H,
                         int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw, void* stream) {
  const ConvShape cs = make_shape(N, H, W, Cin, Cout, R, S, sh, sw, ph, pw);
  return conv_wgrad_h16(cs, x, dy, dw, db, inv_scale, nullptr, ST(stream));
}
int var_linear_h16(const void* a, int64_t lda, const void* w, const float* bias, float* out, int64_t ldo, const void* mask,
                   int mask_kind, int64_t ldm, const float* out_scale, int M, int K, int N, int round_out, void* stream) {
  return linear_h16(M, K, N, a, lda, w, bias, out, ldo, mask, mask_kind, ldm, out_scale, round_out, 0, ST(stream));
}
int var_linear_wgrad_h16(const void* x, int64_t ldx, const void* dy, int64_t ldy, float* dw, int kpad,
                         const float* inv_scale, int M, int K, int N, void* stream) {
  return linear_wgrad_h16(M, K, N, x, ldx, dy, ldy, dw, kpad, inv_scale, ST(stream));
}
int var_maxpool2x2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
  return maxpool_fwd(x, y, N, H, W, C, ST(stream));
}
int var_maxpool2x2_bwd(const float* x, const float* dy, float* dx, int N, int H, int W, int C,
                       void* stream) {
  return maxpool_bwd(x, dy, dx, N, H, W, C, ST(stream));
}

int var_triplet_fwd_bwd(const float* h_img, const float* h_pos, const float* h_neg, int B, int D,
                        int Kh_img, int Kh_snd, const float* W_img, const float* b_img,
                        const float* W_snd, const float* b_snd, float margin, float loss_denominator,
                        float* feats, float* loss, float* loss_rows, float* dh_img, float* dh_pos,
                        float* dh_neg, float* dW_img, float* db_img, float* dW_snd, float* db_snd,
                        void* stream) {
  TailArgs a;
  memset(&a, 0, sizeof(a));
  a.mode = TAIL_TRIPLET; a.B = B; a.D = D; a.Kh_img = Kh_img; a.Kh_snd = Kh_snd;
  a.h_img = h_img; a.h_pos = h_pos; a.h_neg = h_neg;
  a.W_img = W_img; a.b_img = b_img; a.W_snd = W_snd; a.b_snd = b_snd;
  a.margin = margin; a.grad_scale = 1.f / loss_denominator; a.loss_scale = 1.f / loss_denominator;
  a.loss = loss; a.loss_rows = loss_rows;
  if (feats) { a.feat_img = feats; a.feat_pos = feats + (long long)B * D; a.feat_neg = feats + 2LL * B * D; }
  a.dh_img = dh_img; a.dh_pos = dh_pos; a.dh_neg = dh_neg;
  a.dW_img = dW_img; a.db_img = db_img; a.dW_snd = dW_snd; a.db_snd = db_snd;
  return tail_launch(a, ST(stream));
}

}  // extern "C"
