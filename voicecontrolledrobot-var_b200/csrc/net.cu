// Layer plans and the forward / backward executors of the two VAR encoders:
//   kind 0  Kuka   models/pretext/arm_pretext_model.py:9-59
//   kind 1  iTHOR  models/pretext/ai2thor_pretext_model.py:5-64
// orchestrated like PretextNetBase.VAR_forward (models/pretext/pretext_base.py:10-42).
// Activations are NHWC fp32 (tf32-representable values) in a caller-provided
// workspace; parameters live in one flat packed buffer (see include/var_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/var_b200.h"
#include "aux_kernels.cuh"
#include "engine_host.cuh"
#include "kuka_sound.cuh"
#include "triplet.cuh"

namespace var {

namespace {

struct TensorDesc {
  std::string name;
  int ndim;
  int shape[4];
  int O, I, R, S, kpad;  // packed [O][kpad], k = (r*S+s)*I + c
  long long off, packed;
};

enum LayerType : int { LT_CONV = 0, LT_POOL = 1 };

struct Layer {
  int type;
  int H, W, Cin, Cout, R, S, sh, sw, ph, pw, P, Q;
  int relu, round_out, mask_in;
  int tw, tb;  // tensor indices of weight / bias
  // 16-bit operand region: h16 = this conv runs kind::f16 MMAs on f16 input / weights (w16 = f16 copy of the
  // packed weights, refreshed from the fp32 master at every forward); out_h16 = its output is stored as f16
  int h16, out_h16;
  void* w16;
  // run state
  const void* in;
  int in_kind;
  SrcLayout in_sl;
  float* out;
  int N;
};

struct Arena {
  uint8_t* base;
  long long cap, used;
  bool overflow;
  Arena(void* b, long long c) : base(reinterpret_cast<uint8_t*>(b)), cap(c), used(0), overflow(false) {}
  float* alloc(long long floats) {
    const long long bytes = ((floats * 4 + 255) / 256) * 256;
    const long long o = used;
    used += bytes;
    if (!base) return nullptr;
    if (used > cap) { overflow = true; return nullptr; }
    return reinterpret_cast<float*>(base + o);
  }
};

constexpr int kGruT = 73, kGruH = 512, kGruI = 448;

struct GruState {
  int B;
  const float* x;       // [B*T, 448]
  float* xproj[2];      // [B*T, 1536]
  float* gates[2];      // [T][B, 1536]
  float* hn_save[2];    // [T][B, 512]
  float* h_r[2];        // [T+1][B, 512] tf32-rounded hidden states (slot 0 = zeros)
  float* h32[2][2];     // ping-pong fp32 hidden state
  float* out;           // [B, 1024]
  float* out_r;
  unsigned int* counters;  // [row tiles * 2] persistent-kernel group barriers
  void* h_h[2];         // [T+1][B, 512] f16 hidden states: A operand of the 16-bit recurrent GEMM (nullptr: tf32)
  void* x16;            // [B*T, 448] f16 copy of x: A operand of the 16-bit input projection (nullptr: tf32)
  long long ldx, xts;   // xproj element (b, t, c) at b * ldx + t * xts + c
  bool x16_on;          // the forward pass took the 16-bit input projection
};

}  // namespace

struct Net {
  int kind, F, D;
  std::vector<TensorDesc> tensors;
  std::vector<Layer> img_trunk, img_head, snd_trunk, snd_head;
  bool has_gru = false;
  int t_gru[2][4];  // [dir][wih, whh, bih, bhh]
  int t_tail_img_w, t_tail_img_b, t_tail_snd_w, t_tail_snd_b;
  int Kh_img, Kh_snd;
  int img_raw_dim, snd_raw_dim;
  long long nparams = 0;
  float *P = nullptr, *PR = nullptr, *G = nullptr;
  // run state of the last forward
  int n_img = 0, n_snd = 0;
  long long fwd_used = 0;
  GruState gru;
  float *h_img = nullptr, *h_snd = nullptr;  // inputs of the tail
  float *img_raw_nhwc = nullptr, *snd_raw = nullptr;
  // The image and sound branches are independent until the tail.  The sound branch (the critical
  // path: big convs, then 73 + 72 small dependent GRU launches) runs on a HIGH-priority side stream,
  // the image branch stays on the caller's stream: its large-grid kernels fill whatever SMs the
  // sound branch leaves idle without delaying the GRU chain.  VAR_OVERLAP=0 serialises them.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap = true;
  // Data-parallel overlap: the recurrent-layer gradients (rnn.*: 77 % of the iTHOR gradient bytes, one
  // contiguous range of the flat buffer) are complete half way through the backward pass.  The caller
  // may hand in an event that is recorded at that point (on whichever stream produced them), so its
  // all-reduce of that range runs under the remaining conv backward kernels.
  cudaEvent_t ev_rnn_grads = nullptr;  // not owned
  float* h16_scale = nullptr;          // device: {S, 1/S} of the gradient entering the 16-bit region
  unsigned int* h16_amax = nullptr;
  void* whh16[2] = {nullptr, nullptr}; // f16 copies of W_hh (both recurrent kernels), refreshed every forward
  float* gru_scale = nullptr;          // device: [2 directions][S, 1/S] of the BPTT gate gradients
  unsigned int* gru_amax = nullptr;    // [2]
  // 16-bit input projection (x W_ih^T for both directions as ONE [B*T, 448] x [448, 3072] GEMM, its weight gradients
  // and dX = dgi W_ih as ONE GEMM over K = 3072): f16 W_ih of both directions stacked [3072][448], the same matrix
  // transposed [448][3072] (B operand of the dX GEMM), b_ih of both directions [3072]
  void* wih16 = nullptr;
  void* wih16_t = nullptr;
  float* bih_cat = nullptr;

  ~Net() {
    for (auto* L : {&img_trunk, &img_head, &snd_trunk, &snd_head})
      for (Layer& l : *L)
        if (l.w16) cudaFree(l.w16);
    if (h16_scale) cudaFree(h16_scale);
    if (h16_amax) cudaFree(h16_amax);
    for (int d = 0; d < 2; ++d) if (whh16[d]) cudaFree(whh16[d]);
    if (gru_scale) cudaFree(gru_scale);
    if (gru_amax) cudaFree(gru_amax);
    if (wih16) cudaFree(wih16);
    if (wih16_t) cudaFree(wih16_t);
    if (bih_cat) cudaFree(bih_cat);
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
  }
  // Returns the stream the sound branch should use (forked from st) or st itself.
  cudaStream_t fork_side(bool both, cudaStream_t st) {
    if (!both || !overlap) return st;
    if (!side) {
      const char* e = getenv("VAR_OVERLAP");
      if (e && e[0] == '0') { overlap = false; return st; }
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically smallest = greatest priority
      if (cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) {
        overlap = false;
        return st;
      }
    }
    cudaEventRecord(ev_fork, st);
    cudaStreamWaitEvent(side, ev_fork, 0);
    return side;
  }
  void join_side(cudaStream_t used, cudaStream_t st) {
    if (used == st) return;
    cudaEventRecord(ev_join, used);
    cudaStreamWaitEvent(st, ev_join, 0);
  }

  int add_tensor(const std::string& name, int ndim, const int* shape, int O, int I, int R, int S,
                 int kpad) {
    TensorDesc t;
    t.name = name; t.ndim = ndim;
    for (int i = 0; i < 4; ++i) t.shape[i] = i < ndim ? shape[i] : 1;
    t.O = O; t.I = I; t.R = R; t.S = S; t.kpad = kpad;
    t.off = nparams;
    t.packed = (((long long)O * kpad) + 3) / 4 * 4;
    nparams += t.packed;
    tensors.push_back(t);
    return (int)tensors.size() - 1;
  }
  // conv weight + bias named `<name>.weight/.bias`
  void add_conv(std::vector<Layer>& v, const std::string& name, int H, int W, int Cin, int Cout,
                int R, int S, int sh, int sw, int ph, int pw, int relu, int round_out, int mask_in) {
    Layer l;
    memset(&l, 0, sizeof(l));
    l.type = LT_CONV;
    l.H = H; l.W = W; l.Cin = Cin; l.Cout = Cout; l.R = R; l.S = S;
    l.sh = sh; l.sw = sw; l.ph = ph; l.pw = pw;
    l.P = (H + 2 * ph - R) / sh + 1;
    l.Q = (W + 2 * pw - S) / sw + 1;
    l.relu = relu; l.round_out = round_out; l.mask_in = mask_in;
    const int shp[4] = {Cout, Cin, R, S};
    l.tw = add_tensor(name + ".weight", 4, shp, Cout, Cin, R, S, round_up32(R * S * Cin));
    const int bs[1] = {Cout};
    l.tb = add_tensor(name + ".bias", 1, bs, 1, Cout, 1, 1, Cout);
    v.push_back(l);
  }
  // Linear whose input is the NHWC tensor [C, Hh, Ww] flattened (reference flattens NCHW)
  void add_linear(std::vector<Layer>& v, const std::string& name, int C, int Hh, int Ww, int Cout,
                  int relu, int round_out, int mask_in) {
    Layer l;
    memset(&l, 0, sizeof(l));
    l.type = LT_CONV;
    const int In = C * Hh * Ww;
    l.H = 1; l.W = 1; l.Cin = In; l.Cout = Cout; l.R = 1; l.S = 1;
    l.sh = l.sw = 1; l.P = l.Q = 1;
    l.relu = relu; l.round_out = round_out; l.mask_in = mask_in;
    const int shp[2] = {Cout, In};
    l.tw = add_tensor(name + ".weight", 2, shp, Cout, C, Hh, Ww, round_up32(In));
    const int bs[1] = {Cout};
    l.tb = add_tensor(name + ".bias", 1, bs, 1, Cout, 1, 1, Cout);
    v.push_back(l);
  }
  void add_pool(std::vector<Layer>& v, int H, int W, int C) {
    Layer l;
    memset(&l, 0, sizeof(l));
    l.type = LT_POOL;
    l.H = H; l.W = W; l.Cin = C; l.Cout = C; l.P = H / 2; l.Q = W / 2;
    v.push_back(l);
  }
  void add_tail(const std::string& name, int Kh, int* tw, int* tb) {
    const int shp[2] = {D, Kh};
    *tw = add_tensor(name + ".weight", 2, shp, D, Kh, 1, 1, Kh);
    const int bs[1] = {D};
    *tb = add_tensor(name + ".bias", 1, bs, 1, D, 1, 1, (D + 3) / 4 * 4);
  }

  int build() {
    if (kind == 0) {
      // registration order of arm_pretext_model.py:37-56: imgBranch, soundCNN, imgTriplet, soundTriplet
      if (F != 100) return VAR_ERR_UNSUPPORTED;
      const int ch[6] = {3, 32, 32, 64, 64, 64};
      int hw = 96;
      for (int i = 0; i < 5; ++i) {
        add_conv(img_trunk, "imgBranch." + std::to_string(2 * i), hw, hw, ch[i], ch[i + 1], 3, 3, 2,
                 2, 1, 1, 1, 1, i > 0);
        hw = (hw + 2 - 3) / 2 + 1;
      }
      img_raw_dim = 64 * hw * hw;  // 576
      add_conv(snd_trunk, "soundCNN.0", F, 40, 1, 32, 5, 40, 2, 1, 0, 0, 1, 1, 0);
      int h = snd_trunk.back().P;
      for (int i = 1; i < 4; ++i) {
        add_conv(snd_trunk, "soundCNN." + std::to_string(2 * i), h, 1, 32, 32, 3, 1, 2, 1, 0, 0, 1, 1, 1);
        h = snd_trunk.back().P;
      }
      snd_raw_dim = 32 * h;  // 160
      add_linear(img_head, "imgTriplet.0", 64, hw, hw, 128, 1, 0, 1);
      add_tail("imgTriplet.2", 128, &t_tail_img_w, &t_tail_img_b);
      add_linear(snd_head, "soundTriplet.0", 32, h, 1, 128, 1, 0, 1);
      add_tail("soundTriplet.2", 128, &t_tail_snd_w, &t_tail_snd_b);
      Kh_img = 128; Kh_snd = 128;
    } else if (kind == 1) {
      // registration order of ai2thor_pretext_model.py:41-61: imgBranch, rnn, cnn, imgTriplet, soundTriplet
      if (F != 600) return VAR_ERR_UNSUPPORTED;
      add_conv(img_trunk, "imgBranch.0", 96, 96, 3, 32, 3, 3, 1, 1, 1, 1, 1, 1, 0);
      add_conv(img_trunk, "imgBranch.2", 96, 96, 32, 32, 3, 3, 1, 1, 1, 1, 1, 1, 1);
      add_pool(img_trunk, 96, 96, 32);
      add_conv(img_trunk, "imgBranch.5", 48, 48, 32, 64, 3, 3, 1, 1, 1, 1, 1, 1, 1);
      add_pool(img_trunk, 48, 48, 64);
      add_conv(img_trunk, "imgBranch.8", 24, 24, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1, 1);
      add_pool(img_trunk, 24, 24, 64);
      add_conv(img_trunk, "imgBranch.11", 12, 12, 64, 128, 3, 3, 1, 1, 1, 1, 1, 1, 1);
      add_pool(img_trunk, 12, 12, 128);
      add_conv(img_trunk, "imgBranch.14", 6, 6, 128, 128, 3, 3, 2, 2, 1, 1, 1, 1, 1);
      img_raw_dim = 128 * 3 * 3;
      has_gru = true;
      for (int d = 0; d < 2; ++d) {
        const std::string sfx = d ? "_reverse" : "";
        // torch order: weight_ih, weight_hh, bias_ih, bias_hh (per direction)
        const int s_ih[2] = {3 * kGruH, kGruI};
        // W_ih columns follow x = transpose(conv_out, 1, 2).reshape(B, 73, 64*7): col = c*7 + w
        t_gru[d][0] = add_tensor("rnn.weight_ih_l0" + sfx, 2, s_ih, 3 * kGruH, 64, 1, 7, kGruI);
        const int s_hh[2] = {3 * kGruH, kGruH};
        t_gru[d][1] = add_tensor("rnn.weight_hh_l0" + sfx, 2, s_hh, 3 * kGruH, kGruH, 1, 1, kGruH);
        const int s_b[1] = {3 * kGruH};
        t_gru[d][2] = add_tensor("rnn.bias_ih_l0" + sfx, 1, s_b, 1, 3 * kGruH, 1, 1, 3 * kGruH);
        t_gru[d][3] = add_tensor("rnn.bias_hh_l0" + sfx, 1, s_b, 1, 3 * kGruH, 1, 1, 3 * kGruH);
      }
      add_conv(snd_trunk, "cnn.0", F, 40, 1, 64, 11, 11, 2, 2, 5, 5, 1, 1, 0);
      add_conv(snd_trunk, "cnn.2", 300, 20, 64, 64, 11, 5, 2, 2, 5, 5, 1, 1, 1);
      add_conv(snd_trunk, "cnn.4", 150, 13, 64, 64, 7, 3, 2, 2, 1, 1, 1, 1, 1);
      if (snd_trunk.back().P != kGruT || snd_trunk.back().Q != 7) return VAR_ERR_UNSUPPORTED;
      {  // cnn.2 / cnn.4 (Cin = Cout = 64: bound by the operand bytes per MAC in tf32) run in the 16-bit region;
        // cnn.0 feeds them f16 activations, cnn.4 hands tf32-rounded fp32 to the GRU input projection
        Layer& c0 = snd_trunk[0]; Layer& c2 = snd_trunk[1]; Layer& c4 = snd_trunk[2];
        const ConvShape s2{1, c2.H, c2.W, c2.Cin, c2.Cout, c2.R, c2.S, c2.sh, c2.sw, c2.ph, c2.pw, c2.P, c2.Q};
        const ConvShape s4{1, c4.H, c4.W, c4.Cin, c4.Cout, c4.R, c4.S, c4.sh, c4.sw, c4.ph, c4.pw, c4.P, c4.Q};
        if (conv_h16_ok(s2) && conv_h16_ok(s4)) {
          c0.out_h16 = 1;
          c2.h16 = 1; c2.out_h16 = 1;
          c4.h16 = 1;
        }
      }
      snd_raw_dim = 2 * kGruH;
      add_linear(img_head, "imgTriplet.0", 128, 3, 3, 128, 1, 0, 1);
      add_tail("imgTriplet.2", 128, &t_tail_img_w, &t_tail_img_b);
      add_linear(snd_head, "soundTriplet.0", 2 * kGruH, 1, 1, 128, 1, 1, 0);
      add_linear(snd_head, "soundTriplet.2", 128, 1, 1, 64, 1, 0, 1);
      add_tail("soundTriplet.4", 64, &t_tail_snd_w, &t_tail_snd_b);
      Kh_img = 128; Kh_snd = 64;
    } else {
      return VAR_ERR_UNSUPPORTED;
    }
    return VAR_OK;
  }

  // Device-side scratch of the 16-bit region (f16 weight copies: 0.6 MB for the iTHOR net; gradient scale).
  // Called from var_net_bind, the first call that is given device memory (var_net_create also runs on
  // hosts without a GPU, for plan inspection).
  int alloc_h16() {
    for (auto* L : {&img_trunk, &img_head, &snd_trunk, &snd_head})
      for (Layer& l : *L)
        if (l.h16 && !l.w16) {
          const TensorDesc& t = tensors[l.tw];
          VAR_CUDA_CHECK(cudaMalloc(&l.w16, (size_t)t.O * t.kpad * 2));
          if (!h16_scale) {
            VAR_CUDA_CHECK(cudaMalloc(&h16_scale, 2 * sizeof(float)));
            VAR_CUDA_CHECK(cudaMalloc(&h16_amax, sizeof(unsigned int)));
          }
        }
    if (has_gru && gru_h16_enabled() && !whh16[0]) {
      for (int d = 0; d < 2; ++d) VAR_CUDA_CHECK(cudaMalloc(&whh16[d], (size_t)3 * kGruH * kGruH * 2));
      VAR_CUDA_CHECK(cudaMalloc(&gru_scale, 4 * sizeof(float)));
      VAR_CUDA_CHECK(cudaMalloc(&gru_amax, 2 * sizeof(unsigned int)));
    }
    if (has_gru && gru_x16_enabled() && !wih16) {
      VAR_CUDA_CHECK(cudaMalloc(&wih16, (size_t)6 * kGruH * kGruI * 2));
      VAR_CUDA_CHECK(cudaMalloc(&wih16_t, (size_t)6 * kGruH * kGruI * 2));
      VAR_CUDA_CHECK(cudaMalloc(&bih_cat, (size_t)6 * kGruH * sizeof(float)));
    }
    return VAR_OK;
  }

  const float* wr(int t) const { return PR + tensors[t].off; }  // tf32 operand copy
  const float* wm(int t) const { return P + tensors[t].off; }   // fp32 master
  float* gr(int t) const { return G + tensors[t].off; }

  // ------------------------------------------------------------------ forward
  int run_layers_fwd(std::vector<Layer>& L, const void* input, int in_kind, const SrcLayout& in_sl,
                     int N, Arena& ar, cudaStream_t st, float** last_out) {
    const void* cur = input;
    int kind_ = in_kind;
    SrcLayout sl = in_sl;
    for (size_t i = 0; i < L.size(); ++i) {
      Layer& l = L[i];
      l.in = cur; l.in_kind = kind_; l.in_sl = sl; l.N = N;
      const long long out_elems = (long long)N * l.P * l.Q * l.Cout;
      l.out = ar.alloc(l.out_h16 ? (out_elems + 1) / 2 : out_elems);
      if (ar.base) {
        if (ar.overflow) return VAR_ERR_WORKSPACE;
        int rc;
        if (l.type == LT_CONV) {
          ConvShape cs{N, l.H, l.W, l.Cin, l.Cout, l.R, l.S, l.sh, l.sw, l.ph, l.pw, l.P, l.Q};
          if (l.h16) {
            const TensorDesc& t = tensors[l.tw];
            rc = cvt_f16(wm(l.tw), l.w16, (long long)t.O * t.kpad, st);  // weights may have moved (Adam, load)
            if (rc) return rc;
            rc = conv_fwd_h16(cs, cur, l.w16, wm(l.tb), l.out, l.out_h16, l.relu, l.round_out, st);
          } else {
            rc = conv_fwd(cs, cur, kind_, &sl, wr(l.tw), wm(l.tb), l.out, l.relu, l.round_out, st, l.out_h16);
          }
        } else {
          rc = maxpool_fwd(reinterpret_cast<const float*>(cur), l.out, N, l.H, l.W, l.Cin, st);
        }
        if (rc) return rc;
      }
      cur = l.out;
      kind_ = SRC_NHWC_F32;
    }
    *last_out = const_cast<float*>(reinterpret_cast<const float*>(cur));
    return VAR_OK;
  }

  int gru_forward(const float* x, int B, bool train, Arena& ar, cudaStream_t st) {
    GruState& g = gru;
    g.B = B; g.x = x;
    const long long BT = (long long)B * kGruT;
    // 16-bit input projection: both directions in one [B*T, 6H] matrix (direction d at column d * 3H)
    const bool x16 = gru_x16_enabled() && (wih16 || !ar.base);
    float* xproj_all = ar.alloc(BT * 6 * kGruH);
    g.x16 = x16 ? ar.alloc((BT * kGruI + 1) / 2) : nullptr;
    g.x16_on = x16;
    g.ldx = x16 ? (long long)kGruT * 6 * kGruH : (long long)kGruT * 3 * kGruH;
    g.xts = x16 ? 6LL * kGruH : 3LL * kGruH;
    for (int d = 0; d < 2; ++d) {
      g.xproj[d] = x16 ? xproj_all + (long long)d * 3 * kGruH : xproj_all + (long long)d * BT * 3 * kGruH;
      g.gates[d] = train ? ar.alloc(BT * 3 * kGruH) : nullptr;
      g.hn_save[d] = train ? ar.alloc(BT * kGruH) : nullptr;
      g.h_r[d] = ar.alloc((long long)(kGruT + 1) * B * kGruH);
      g.h_h[d] = (gru_h16_enabled() && (whh16[0] || !ar.base)) ? ar.alloc(((long long)(kGruT + 1) * B * kGruH + 1) / 2) : nullptr;
      g.h32[d][0] = ar.alloc((long long)B * kGruH);
      g.h32[d][1] = ar.alloc((long long)B * kGruH);
    }
    g.out = ar.alloc((long long)B * 2 * kGruH);
    g.out_r = ar.alloc((long long)B * 2 * kGruH);
    g.counters = reinterpret_cast<unsigned int*>(ar.alloc(256));
    if (!ar.base) return VAR_OK;
    if (ar.overflow) return VAR_ERR_WORKSPACE;
    if (x16) {
      // x-projection of both directions for all steps as one f16 GEMM: [B*T, 448] x [448, 6H] + (b_ih_fwd | b_ih_bwd)
      int rc = cvt_f16(x, g.x16, BT * kGruI, st);
      for (int d = 0; d < 2 && !rc; ++d) {
        rc = cvt_f16(wm(t_gru[d][0]), reinterpret_cast<uint16_t*>(wih16) + (size_t)d * 3 * kGruH * kGruI,
                     (long long)3 * kGruH * kGruI, st);  // weights may have moved
        if (!rc && cudaMemcpyAsync(bih_cat + (size_t)d * 3 * kGruH, wm(t_gru[d][2]), (size_t)3 * kGruH * 4,
                                   cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = VAR_ERR_CUDA;
      }
      if (!rc) rc = linear_h16((int)BT, kGruI, 6 * kGruH, g.x16, kGruI, wih16, bih_cat, xproj_all, 6LL * kGruH, nullptr, 0,
                               0, nullptr, 0, 0, st);
      if (rc) return rc;
    }
    for (int d = 0; d < 2; ++d) {
      int rc = VAR_OK;
      if (!x16) {
        // x-projection for all steps: [B*T, 448] x W_ih^T + b_ih
        ConvShape cs{(int)BT, 1, 1, kGruI, 3 * kGruH, 1, 1, 1, 1, 0, 0, 1, 1};
        rc = conv_fwd(cs, x, SRC_NHWC_F32, nullptr, wr(t_gru[d][0]), wm(t_gru[d][2]), g.xproj[d], 0, 0, st);
        if (rc) return rc;
      }
      VAR_CUDA_CHECK(cudaMemsetAsync(g.h_r[d], 0, (size_t)B * kGruH * 4, st));
      VAR_CUDA_CHECK(cudaMemsetAsync(g.h32[d][0], 0, (size_t)B * kGruH * 4, st));
      if (g.h_h[d]) {
        VAR_CUDA_CHECK(cudaMemsetAsync(g.h_h[d], 0, (size_t)B * kGruH * 2, st));
        rc = cvt_f16(wm(t_gru[d][1]), whh16[d], (long long)3 * kGruH * kGruH, st);  // weights may have moved
        if (rc) return rc;
      }
    }
    const long long slot = (long long)B * kGruH;
    {  // all 73 steps in one cooperative launch when the grid fits the device
      const float* xp[2] = {g.xproj[0], g.xproj[1]};
      const float* whh[2] = {wr(t_gru[0][1]), wr(t_gru[1][1])};
      const float* bhh[2] = {wm(t_gru[0][3]), wm(t_gru[1][3])};
      float* h32[2][2] = {{g.h32[0][0], g.h32[0][1]}, {g.h32[1][0], g.h32[1][1]}};
      float* hr[2] = {g.h_r[0], g.h_r[1]};
      float* gt[2] = {g.gates[0], g.gates[1]};
      float* hs[2] = {g.hn_save[0], g.hn_save[1]};
      const void* w16[2] = {whh16[0], whh16[1]};
      void* hh[2] = {g.h_h[0], g.h_h[1]};
      const int rc = gru_persist_fwd(B, kGruH, kGruT, xp, g.ldx, whh, bhh, h32, hr, gt, hs,
                                     g.counters, st, g.h_h[0] ? w16 : nullptr, g.h_h[0] ? hh : nullptr, g.xts);
      if (rc == VAR_OK)
        return concat2(g.h32[0][kGruT & 1], g.h32[1][kGruT & 1], g.out, g.out_r, B, kGruH, st);
      if (rc != VAR_ERR_UNSUPPORTED) return rc;
    }
    const bool keep_all = true;  // h_r always holds all T+1 slots
    for (int s = 0; s < kGruT; ++s) {
      const float* hp[2];
      const float* whh[2];
      GruEpiParams q[2];
      for (int d = 0; d < 2; ++d) {
        const int t = d == 0 ? s : kGruT - 1 - s;
        const int cur_slot = keep_all ? s : (s & 1), nxt_slot = keep_all ? s + 1 : ((s + 1) & 1);
        hp[d] = g.h_r[d] + cur_slot * slot;
        whh[d] = wr(t_gru[d][1]);
        memset(&q[d], 0, sizeof(GruEpiParams));
        q[d].xproj = g.xproj[d] + (long long)t * g.xts;
        q[d].ldx = g.ldx;
        q[d].bhh = wm(t_gru[d][3]);
        q[d].hprev = g.h32[d][s & 1];
        q[d].hnew = g.h32[d][(s + 1) & 1];
        q[d].hnew_r = g.h_r[d] + nxt_slot * slot;
        q[d].gates = train ? g.gates[d] + (long long)s * B * 3 * kGruH : nullptr;
        q[d].hn_save = train ? g.hn_save[d] + (long long)s * B * kGruH : nullptr;
        q[d].Hdim = kGruH;
      }
      int rc = gru_step_fwd(2, B, kGruH, hp, whh, q, st);
      if (rc) return rc;
    }
    return concat2(g.h32[0][kGruT & 1], g.h32[1][kGruT & 1], g.out, g.out_r, B, kGruH, st);
  }

  // Kuka sound branch: one fused fp32 kernel (kuka_sound.cu).  Buffers are laid out and the
  // per-layer run state is filled exactly as run_layers_fwd would, so the tcgen05 backward
  // path works on them unchanged.
  bool fp32_sound = true;
  int kuka_sound_forward(const float* sounds, const SrcLayout& sl, int N, bool train, Arena& ar,
                         cudaStream_t st) {
    const void* cur = sounds;
    int kind_ = SRC_STRIDED_F32;
    for (auto* L : {&snd_trunk, &snd_head})
      for (Layer& l : *L) {
        l.in = cur; l.in_kind = kind_; l.in_sl = sl; l.N = N;
        l.out = ar.alloc((long long)N * l.P * l.Q * l.Cout);
        cur = l.out;
        kind_ = SRC_NHWC_F32;
      }
    snd_raw = snd_trunk.back().out;
    h_snd = snd_head.back().out;
    if (!ar.base) return VAR_OK;
    if (ar.overflow) return VAR_ERR_WORKSPACE;
    KukaSoundArgs a;
    a.x = sounds;
    a.w1 = wm(snd_trunk[0].tw); a.b1 = wm(snd_trunk[0].tb);
    a.w2 = wm(snd_trunk[1].tw); a.b2 = wm(snd_trunk[1].tb);
    a.w3 = wm(snd_trunk[2].tw); a.b3 = wm(snd_trunk[2].tb);
    a.w4 = wm(snd_trunk[3].tw); a.b4 = wm(snd_trunk[3].tb);
    a.wl = wm(snd_head[0].tw); a.bl = wm(snd_head[0].tb);
    a.act1 = train ? snd_trunk[0].out : nullptr;
    a.act2 = train ? snd_trunk[1].out : nullptr;
    a.act3 = train ? snd_trunk[2].out : nullptr;
    a.act4 = snd_trunk[3].out;
    a.hidden = snd_head[0].out;
    return kuka_sound_fwd(a, N, st);
  }

  int forward(const void* images, int image_kind, int n_images, const float* sounds, int n_sounds,
              void* ws, long long ws_bytes, bool train, cudaStream_t st) {
    Arena ar(ws, ws_bytes);
    n_img = n_images; n_snd = n_sounds;
    h_img = h_snd = img_raw_nhwc = snd_raw = nullptr;
    const bool both = n_images > 0 && n_sounds > 0;
    cudaStream_t st_snd = (ar.base && both) ? fork_side(true, st) : st;
    struct Joiner {  // joins the side stream on every exit path
      Net* n; cudaStream_t a, b;
      ~Joiner() { n->join_side(a, b); }
    } joiner{this, st_snd, st};
    if (n_sounds > 0) {
      SrcLayout sl;
      sl.sC = 0; sl.sW = 1; sl.sH = 40; sl.sN = (long long)F * 40;
      sl.scale = 1.f;
      float* t;
      int rc;
      if (kind == 0 && fp32_sound) {
        rc = kuka_sound_forward(sounds, sl, n_sounds, train, ar, st_snd);
        if (rc) return rc;
      } else {
        rc = run_layers_fwd(snd_trunk, sounds, SRC_STRIDED_F32, sl, n_sounds, ar, st_snd, &t);
        if (rc) return rc;
        if (has_gru) {
          rc = gru_forward(t, n_sounds, train, ar, st_snd);
          if (rc) return rc;
          snd_raw = gru.out;
          t = gru.out_r;
        } else {
          snd_raw = t;
        }
        rc = run_layers_fwd(snd_head, t, SRC_NHWC_F32, sl, n_sounds, ar, st_snd, &h_snd);
        if (rc) return rc;
      }
    }
    if (n_images > 0) {
      SrcLayout sl;
      sl.sW = 1; sl.sH = 96; sl.sC = 96 * 96; sl.sN = 3 * 96 * 96;  // NCHW
      sl.scale = image_kind == 0 ? 1.f / 255.f : 1.f;
      float* t;
      int rc = run_layers_fwd(img_trunk, images, image_kind == 0 ? SRC_STRIDED_U8 : SRC_STRIDED_F32,
                              sl, n_images, ar, st, &t);
      if (rc) return rc;
      img_raw_nhwc = t;
      rc = run_layers_fwd(img_head, t, SRC_NHWC_F32, sl, n_images, ar, st, &h_img);
      if (rc) return rc;
    }
    fwd_used = ar.used;
    return VAR_OK;
  }

  // ----------------------------------------------------------------- backward
  // dY: grad wrt the output of the last layer of L (ReLU-masked, tf32-rounded).
  // Returns in *dx_out the grad wrt the input of L[0] (nullptr when L[0] reads the
  // branch input, which needs no gradient).
  int run_layers_bwd(std::vector<Layer>& L, float* dY, Arena& ar, cudaStream_t st, float** dx_out) {
    float* dy = dY;
    bool dy_h16 = false;  // dy is f16 times the region's gradient scale (produced by an h16 dgrad below)
    for (int i = (int)L.size() - 1; i >= 0; --i) {
      Layer& l = L[i];
      const int N = l.N;
      const bool need_dx = l.in_kind == SRC_NHWC_F32;
      if (l.type == LT_CONV && l.h16) {
        // ---- 16-bit region: f16 activations (l.in), f16 weights, f16 scaled gradients
        const long long dy_elems = (long long)N * l.P * l.Q * l.Cout, dx_elems = (long long)N * l.H * l.W * l.Cin;
        const bool prev_h16 = i > 0 && L[i - 1].type == LT_CONV && L[i - 1].h16;
        float* dy16 = dy_h16 ? dy : ar.alloc((dy_elems + 1) / 2);
        float* dx = need_dx ? ar.alloc(prev_h16 ? (dx_elems + 1) / 2 : dx_elems) : nullptr;
        if (ar.base) {
          if (ar.overflow) return VAR_ERR_WORKSPACE;
          ConvShape cs{N, l.H, l.W, l.Cin, l.Cout, l.R, l.S, l.sh, l.sw, l.ph, l.pw, l.P, l.Q};
          int rc = VAR_OK;
          if (!dy_h16) {  // gradient enters the region: pick the power-of-two scale from its largest magnitude
            rc = grad_to_f16_scaled(dy, dy16, dy_elems, h16_scale, h16_amax, st);
            if (rc) return rc;
          }
          rc = conv_wgrad_h16(cs, l.in, dy16, gr(l.tw), gr(l.tb), h16_scale + 1, nullptr, st);
          if (rc) return rc;
          if (need_dx) {
            // towards another h16 layer the gradient stays f16 and scaled; leaving the region it is
            // unscaled and stored as tf32-rounded fp32
            rc = conv_dgrad_h16(cs, dy16, l.w16, dx, prev_h16 ? 1 : 0, l.mask_in ? l.in : nullptr, 1,
                                prev_h16 ? nullptr : h16_scale + 1, 1, st);
            if (rc) return rc;
          }
        }
        dy = dx;
        dy_h16 = prev_h16;
        continue;
      }
      float* dx = need_dx ? ar.alloc((long long)N * l.H * l.W * l.Cin) : nullptr;
      if (ar.base) {
        if (ar.overflow) return VAR_ERR_WORKSPACE;
        int rc = VAR_OK;
        if (l.type == LT_CONV) {
          ConvShape cs{N, l.H, l.W, l.Cin, l.Cout, l.R, l.S, l.sh, l.sw, l.ph, l.pw, l.P, l.Q};
          rc = conv_wgrad(cs, l.in, l.in_kind, &l.in_sl, dy, gr(l.tw), gr(l.tb), st);
          if (rc) return rc;
          if (need_dx)
            rc = conv_dgrad(cs, dy, wr(l.tw), dx,
                            l.mask_in ? reinterpret_cast<const float*>(l.in) : nullptr, nullptr, 1, st);
        } else {
          rc = maxpool_bwd(reinterpret_cast<const float*>(l.in), dy, dx, N, l.H, l.W, l.Cin, st);
        }
        if (rc) return rc;
      }
      dy = dx;
    }
    *dx_out = dy;
    return VAR_OK;
  }

  // d_out: [B, 1024] grad wrt cat(h_fwd_T, h_bwd_T).  Produces dX [B*T, 448] masked by X > 0.
  int gru_backward(const float* d_out, Arena& ar, cudaStream_t st, float** dx_out) {
    GruState& g = gru;
    const int B = g.B;
    const long long BT = (long long)B * kGruT, slot = (long long)B * kGruH;
    float* dgi[2]; float* dgh[2]; float* dh[2][2]; float* dhd[2];
    for (int d = 0; d < 2; ++d) {
      dgi[d] = ar.alloc(BT * 3 * kGruH);   // batch-major [B, T, 3H]
      dgh[d] = ar.alloc(BT * 3 * kGruH);   // step-major  [T][B, 3H]
      dh[d][0] = ar.alloc(slot);
      dh[d][1] = ar.alloc(slot);
      dhd[d] = ar.alloc(slot);
    }
    float* dx0 = ar.alloc(BT * kGruI);
    float* dx = ar.alloc(BT * kGruI);
    const bool h16 = gru_h16_enabled() && (whh16[0] || !ar.base);
    void* dgh_h[2] = {nullptr, nullptr};
    if (h16) for (int d = 0; d < 2; ++d) dgh_h[d] = ar.alloc((BT * 3 * kGruH + 1) / 2);
    // 16-bit input projection backward: scaled f16 dgi of both directions in one [B*T, 6H] matrix
    const bool x16 = h16 && g.x16_on && (wih16 || !ar.base);
    void* dgi_h[2] = {nullptr, nullptr};
    if (x16) {
      dgi_h[0] = ar.alloc(BT * 3 * kGruH);  // (B*T * 6H halves)
      dgi_h[1] = ar.base ? reinterpret_cast<uint16_t*>(dgi_h[0]) + 3 * kGruH : nullptr;
    }
    const long long ldgi_h = (long long)kGruT * 6 * kGruH, tsgi_h = 6LL * kGruH;
    if (!ar.base) { *dx_out = nullptr; return VAR_OK; }
    if (ar.overflow) return VAR_ERR_WORKSPACE;
    int rc = split2(d_out, dh[0][0], dh[1][0], B, kGruH, st);
    if (rc) return rc;
    auto slot_t = [&](int d, int s) { return d == 0 ? s : kGruT - 1 - s; };
    {  // last step: stand-alone cell backward
      const int s = kGruT - 1;
      GruBwdArgs a[2];
      for (int d = 0; d < 2; ++d) {
        a[d].dh = dh[d][0];
        a[d].gates = g.gates[d] + (long long)s * B * 3 * kGruH;
        a[d].hn_save = g.hn_save[d] + (long long)s * slot;
        a[d].hprev = g.h_r[d] + (long long)s * slot;
        a[d].dgi = dgi[d] + (long long)slot_t(d, s) * 3 * kGruH;
        a[d].ldgi = (long long)kGruT * 3 * kGruH;
        a[d].dgh = dgh[d] + (long long)s * B * 3 * kGruH;
        a[d].dhd = dhd[d];
      }
      rc = gru_cell_bwd(a[0], a[1], 2, B, kGruH, st);
      if (rc) return rc;
    }
    // steps T-1 .. 1: recurrence GEMM with the cell backward of the previous step fused in
    float* dhd_pp[2][2] = {{dhd[0], dh[0][1]}, {dhd[1], dh[1][1]}};
    bool persistent_done = false;
    int bias_done = 0;  // the BPTT kernel also produced db_ih / db_hh
    GruBwdExtra ex;
    memset(&ex, 0, sizeof(ex));
    ex.bias_done = &bias_done;
    int dgi_h_done = 0;  // the BPTT kernel wrote the scaled f16 dgi (and skipped the fp32 dgi / dgh stores)
    ex.dgi_h_done = &dgi_h_done;
    if (x16) {
      // one gradient scale for both directions (their dgi are the K dimension of ONE dX GEMM), chosen from the last
      // step's gate gradients; that step's dgh / dgi slices are converted here, the BPTT kernel writes the rest
      const long long lo = (long long)(kGruT - 1) * B * 3 * kGruH;
      const float* gh[2] = {dgh[0] + lo, dgh[1] + lo};
      void* ghh[2] = {reinterpret_cast<uint16_t*>(dgh_h[0]) + lo, reinterpret_cast<uint16_t*>(dgh_h[1]) + lo};
      const float* gi[2] = {dgi[0] + (long long)slot_t(0, kGruT - 1) * 3 * kGruH, dgi[1] + (long long)slot_t(1, kGruT - 1) * 3 * kGruH};
      void* gih[2] = {reinterpret_cast<uint16_t*>(dgi_h[0]) + (long long)slot_t(0, kGruT - 1) * tsgi_h,
                      reinterpret_cast<uint16_t*>(dgi_h[1]) + (long long)slot_t(1, kGruT - 1) * tsgi_h};
      rc = gru_last_to_f16(gh, ghh, gi, (long long)kGruT * 3 * kGruH, gih, ldgi_h, B, 3 * kGruH, gru_scale, gru_amax, st);
      if (rc) return rc;
    }
    for (int d = 0; d < 2; ++d) {
      ex.db_ih[d] = gr(t_gru[d][2]); ex.db_hh[d] = gr(t_gru[d][3]);
      if (h16) {
        if (!x16) {
          // the gate gradients of the last step enter the 16-bit recurrence: scale from their largest magnitude
          const long long lo = (long long)(kGruT - 1) * B * 3 * kGruH;
          rc = grad_to_f16_scaled(dgh[d] + lo, reinterpret_cast<uint16_t*>(dgh_h[d]) + lo, (long long)B * 3 * kGruH,
                                  gru_scale + 2 * d, gru_amax + d, st);
          if (rc) return rc;
        }
        ex.whh16[d] = whh16[d]; ex.dgh_h[d] = dgh_h[d]; ex.gscale[d] = x16 ? gru_scale : gru_scale + 2 * d;
        if (x16) { ex.dgi_h[d] = dgi_h[d]; ex.dgi_h_ld = ldgi_h; ex.dgi_h_ts = tsgi_h; }
      }
    }
    {
      const float* whh[2] = {wr(t_gru[0][1]), wr(t_gru[1][1])};
      const float* gt[2] = {g.gates[0], g.gates[1]};
      const float* hs[2] = {g.hn_save[0], g.hn_save[1]};
      const float* hr[2] = {g.h_r[0], g.h_r[1]};
      float* dgh2[2] = {dgh[0], dgh[1]};
      float* dgi2[2] = {dgi[0], dgi[1]};
      rc = gru_persist_bwd(B, kGruH, kGruT, whh, gt, hs, hr, dgh2, dgi2, dhd_pp, g.counters, st, &ex);
      if (rc == VAR_OK) persistent_done = true;
      else if (rc != VAR_ERR_UNSUPPORTED) return rc;
    }
    for (int s = kGruT - 1; s >= 1 && !persistent_done; --s) {
      const int cur = (kGruT - 1 - s) & 1;
      const float* dy2[2];
      const float* w2[2] = {wr(t_gru[0][1]), wr(t_gru[1][1])};
      GruBwdEpiParams q[2];
      for (int d = 0; d < 2; ++d) {
        dy2[d] = dgh[d] + (long long)s * B * 3 * kGruH;
        q[d].dhd_in = dhd_pp[d][cur];
        q[d].gates = g.gates[d] + (long long)(s - 1) * B * 3 * kGruH;
        q[d].hn_save = g.hn_save[d] + (long long)(s - 1) * slot;
        q[d].hprev = g.h_r[d] + (long long)(s - 1) * slot;
        q[d].dgi = dgi[d] + (long long)slot_t(d, s - 1) * 3 * kGruH;
        q[d].ldgi = (long long)kGruT * 3 * kGruH;
        q[d].dgh = dgh[d] + (long long)(s - 1) * B * 3 * kGruH;
        q[d].dhd_out = dhd_pp[d][cur ^ 1];
        q[d].Hdim = kGruH;
      }
      rc = gru_step_bwd(2, B, kGruH, dy2, w2, q, st);
      if (rc == VAR_ERR_UNSUPPORTED) {  // cp.async gather mode: unfused GEMM + cell kernel
        float* dx2[2] = {dh[0][0], dh[1][0]};
        const float* add2[2] = {q[0].dhd_in, q[1].dhd_in};
        rc = linear_dgrad2(2, B, kGruH, 3 * kGruH, dy2, w2, dx2, add2, 0, st);
        if (rc) return rc;
        GruBwdArgs a[2];
        for (int d = 0; d < 2; ++d) {
          a[d].dh = dh[d][0]; a[d].gates = q[d].gates; a[d].hn_save = q[d].hn_save; a[d].hprev = q[d].hprev;
          a[d].dgi = q[d].dgi; a[d].ldgi = q[d].ldgi; a[d].dgh = q[d].dgh; a[d].dhd = q[d].dhd_out;
        }
        rc = gru_cell_bwd(a[0], a[1], 2, B, kGruH, st);
      }
      if (rc) return rc;
    }
    if (dgi_h_done) {
      // every gate gradient exists as scaled f16: weight gradients and dX on 16-bit operands
      for (int d = 0; d < 2; ++d) {
        // dW_hh += dgh^T h_prev (rows step-major; h_h slots 0 .. T-1 are the previous states)
        rc = linear_wgrad_h16((int)BT, kGruH, 3 * kGruH, g.h_h[d], kGruH, dgh_h[d], 3 * kGruH, gr(t_gru[d][1]), kGruH,
                              gru_scale + 1, st);
        if (rc) return rc;
        // dW_ih += dgi^T x (rows batch-major; direction d = columns d * 3H of the [B*T, 6H] matrix)
        rc = linear_wgrad_h16((int)BT, kGruI, 3 * kGruH, g.x16, kGruI, dgi_h[d], 6 * kGruH, gr(t_gru[d][0]), kGruI,
                              gru_scale + 1, st);
        if (rc) return rc;
      }
      if (ev_rnn_grads) VAR_CUDA_CHECK(cudaEventRecord(ev_rnn_grads, st));  // every rnn.* gradient is final
      // dX = [dgi_fwd | dgi_bwd] [W_ih_fwd ; W_ih_bwd] / S, ReLU mask of the conv output: one GEMM over K = 6H
      for (int d = 0; d < 2; ++d) {
        rc = cvt_f16_transpose(wm(t_gru[d][0]), wih16_t, 3 * kGruH, kGruI, 6LL * kGruH, d * 3 * kGruH, st);
        if (rc) return rc;
      }
      rc = linear_h16((int)BT, 6 * kGruH, kGruI, dgi_h[0], 6LL * kGruH, wih16_t, nullptr, dx, kGruI, g.x, 0, kGruI,
                      gru_scale + 1, 1, 1, st);
      if (rc) return rc;
      *dx_out = dx;
      return VAR_OK;
    }
    ConvShape cih{(int)BT, 1, 1, kGruI, 3 * kGruH, 1, 1, 1, 1, 0, 0, 1, 1};
    for (int d = 0; d < 2; ++d) {
      // dW_hh += dgh^T h_prev ; db_hh += colsum(dgh)      (rows: step-major)
      ConvShape chh{(int)BT, 1, 1, kGruH, 3 * kGruH, 1, 1, 1, 1, 0, 0, 1, 1};
      // (bias gradients: accumulated inside the BPTT kernel when it ran; else column sums here)
      rc = conv_wgrad(chh, g.h_r[d], SRC_NHWC_F32, nullptr, dgh[d], gr(t_gru[d][1]), bias_done ? nullptr : gr(t_gru[d][3]), st);
      if (rc) return rc;
      // dW_ih += dgi^T x ; db_ih += colsum(dgi)           (rows: batch-major)
      rc = conv_wgrad(cih, g.x, SRC_NHWC_F32, nullptr, dgi[d], gr(t_gru[d][0]), bias_done ? nullptr : gr(t_gru[d][2]), st);
      if (rc) return rc;
    }
    if (ev_rnn_grads) VAR_CUDA_CHECK(cudaEventRecord(ev_rnn_grads, st));  // every rnn.* gradient is final
    for (int d = 0; d < 2; ++d) {
      // dX = dgi_fwd W_ih_fwd + dgi_bwd W_ih_bwd, ReLU mask of the conv output
      rc = conv_dgrad(cih, dgi[d], wr(t_gru[d][0]), d == 0 ? dx0 : dx, d == 0 ? nullptr : g.x,
                      d == 0 ? nullptr : dx0, d == 0 ? 0 : 1, st);
      if (rc) return rc;
    }
    *dx_out = dx;
    return VAR_OK;
  }

  int backward_from_dh(float* dh_img, float* dh_snd, void* ws, long long ws_bytes, cudaStream_t st) {
    Arena ar(ws, ws_bytes);
    ar.used = fwd_used;
    cudaStream_t st_snd = fork_side(n_img > 0 && dh_img && n_snd > 0 && dh_snd, st);
    struct Joiner {
      Net* n; cudaStream_t a, b;
      ~Joiner() { n->join_side(a, b); }
    } joiner{this, st_snd, st};
    // the branches run concurrently: each keeps its own gradient scratch region
    if (n_snd > 0 && dh_snd) {
      float* d;
      int rc = run_layers_bwd(snd_head, dh_snd, ar, st_snd, &d);
      if (rc) return rc;
      if (has_gru) {
        rc = gru_backward(d, ar, st_snd, &d);
        if (rc) return rc;
      }
      rc = run_layers_bwd(snd_trunk, d, ar, st_snd, &d);
      if (rc) return rc;
    }
    if (n_img > 0 && dh_img) {
      float* d;
      int rc = run_layers_bwd(img_head, dh_img, ar, st, &d);
      if (rc) return rc;
      rc = run_layers_bwd(img_trunk, d, ar, st, &d);
      if (rc) return rc;
    }
    return VAR_OK;
  }

  long long bwd_scratch_bytes(int n_images, int n_sounds) {
    // dry run with the per-layer shapes of a forward of the same sizes
    Arena ar(nullptr, 0);
    long long best = 0;
    float* d = nullptr;
    if (n_images > 0) {
      for (auto& l : img_head) { l.N = n_images; l.in_kind = SRC_NHWC_F32; }
      for (size_t i = 0; i < img_trunk.size(); ++i) {
        img_trunk[i].N = n_images;
        img_trunk[i].in_kind = i == 0 ? SRC_STRIDED_U8 : SRC_NHWC_F32;
      }
      run_layers_bwd(img_head, nullptr, ar, 0, &d);
      run_layers_bwd(img_trunk, nullptr, ar, 0, &d);
      best = ar.used;
    }
    if (n_sounds > 0) {
      for (auto& l : snd_head) { l.N = n_sounds; l.in_kind = SRC_NHWC_F32; }
      for (size_t i = 0; i < snd_trunk.size(); ++i) {
        snd_trunk[i].N = n_sounds;
        snd_trunk[i].in_kind = i == 0 ? SRC_STRIDED_F32 : SRC_NHWC_F32;
      }
      run_layers_bwd(snd_head, nullptr, ar, 0, &d);
      if (has_gru) { gru.B = n_sounds; gru_backward(nullptr, ar, 0, &d); }
      run_layers_bwd(snd_trunk, nullptr, ar, 0, &d);
      best = ar.used;  // image + sound regions (disjoint: the branches overlap in time)
    }
    return best;
  }

  void fill_tail_weights(TailArgs& a, bool grads) const {
    a.D = D; a.Kh_img = Kh_img; a.Kh_snd = Kh_snd;
    a.W_img = wm(t_tail_img_w); a.b_img = wm(t_tail_img_b);
    a.W_snd = wm(t_tail_snd_w); a.b_snd = wm(t_tail_snd_b);
    if (grads) {
      a.dW_img = gr(t_tail_img_w); a.db_img = gr(t_tail_img_b);
      a.dW_snd = gr(t_tail_snd_w); a.db_snd = gr(t_tail_snd_b);
    }
  }
};

}  // namespace var

// ===========================================================================
// C ABI
// ===========================================================================
using var::Net;
using var::TailArgs;

#define NET(p) (reinterpret_cast<Net*>(p))
#define ST(s) (reinterpret_cast<cudaStream_t>(s))

extern "C" {

int var_net_create(int kind, int sound_frames, int rep_dim, void** out) {
  if (!out || rep_dim < 1 || rep_dim > 8) return VAR_ERR_ARG;
  Net* n = new Net();
  n->kind = kind; n->F = sound_frames; n->D = rep_dim;
  const int rc = n->build();
  if (rc) { delete n; return rc; }
  *out = n;
  return VAR_OK;
}
int var_net_destroy(void* net) { delete NET(net); return VAR_OK; }
int64_t var_net_param_floats(void* net) { return NET(net)->nparams; }
int var_net_num_tensors(void* net) { return (int)NET(net)->tensors.size(); }
int var_net_tensor_info(void* net, int index, char* name, int name_cap, int* ndim, int* shape,
                        int64_t* offset, int64_t* packed_floats) {
  Net* n = NET(net);
  if (index < 0 || index >= (int)n->tensors.size()) return VAR_ERR_ARG;
  const auto& t = n->tensors[index];
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", t.name.c_str());
  if (ndim) *ndim = t.ndim;
  if (shape) for (int i = 0; i < 4; ++i) shape[i] = t.shape[i];
  if (offset) *offset = t.off;
  if (packed_floats) *packed_floats = t.packed;
  return VAR_OK;
}
int var_net_bind(void* net, float* p, float* pr, float* g) {
  Net* n = NET(net);
  n->P = p; n->PR = pr; n->G = g;
  return n->alloc_h16();
}
int var_net_load_tensor(void* net, int index, const float* src, void* stream) {
  Net* n = NET(net);
  if (index < 0 || index >= (int)n->tensors.size() || !n->P || !n->PR) return VAR_ERR_ARG;
  const auto& t = n->tensors[index];
  return var::pack_weight(src, n->P + t.off, n->PR + t.off, t.O, t.I, t.R, t.S, t.kpad, ST(stream));
}
int var_net_store_tensor(void* net, int index, int which, float* dst, void* stream) {
  Net* n = NET(net);
  if (index < 0 || index >= (int)n->tensors.size()) return VAR_ERR_ARG;
  const float* src = which == 0 ? n->P : n->G;
  if (!src) return VAR_ERR_ARG;
  const auto& t = n->tensors[index];
  return var::unpack_weight(src + t.off, dst, t.O, t.I, t.R, t.S, t.kpad, ST(stream));
}
int var_net_refresh_mma(void* net, void* stream) {
  Net* n = NET(net);
  if (!n->P || !n->PR) return VAR_ERR_ARG;
  return var::round_copy(n->P, n->PR, n->nparams, ST(stream));
}
int var_net_grad_bucket(void* net, int bucket, int64_t* offset, int64_t* count) {
  Net* n = NET(net);
  if (!offset || !count) return VAR_ERR_ARG;
  if (bucket != 0 || !n->has_gru) return VAR_ERR_UNSUPPORTED;
  const auto& first = n->tensors[n->t_gru[0][0]];
  const auto& last = n->tensors[n->t_gru[1][3]];
  *offset = first.off;
  *count = last.off + last.packed - first.off;
  return VAR_OK;
}
int var_net_set_bucket_event(void* net, int bucket, void* cuda_event) {
  Net* n = NET(net);
  if (bucket != 0 || !n->has_gru) return VAR_ERR_UNSUPPORTED;
  n->ev_rnn_grads = reinterpret_cast<cudaEvent_t>(cuda_event);
  return VAR_OK;
}
int var_net_set_overlap(void* net, int on) {
  NET(net)->overlap = on != 0;
  return VAR_OK;
}
int var_net_raw_dims(void* net, int* img_raw, int* snd_raw) {
  if (img_raw) *img_raw = NET(net)->img_raw_dim;
  if (snd_raw) *snd_raw = NET(net)->snd_raw_dim;
  return VAR_OK;
}
int64_t var_net_workspace_bytes(void* net, int n_images, int n_sounds, int train) {
  Net* n = NET(net);
  int rc = n->forward(nullptr, 0, n_images, nullptr, n_sounds, nullptr, 0, train != 0, 0);
  if (rc) return rc;
  long long total = n->fwd_used;
  // tail scratch: feats / dh buffers
  total += 4096 + (long long)(n_images + n_sounds) * (n->Kh_img + n->Kh_snd + 64) * 4;
  if (train) total += n->bwd_scratch_bytes(n_images, n_sounds);
  return total + 4096;
}

int var_net_forward(void* net, const void* images, int image_kind, int n_images, const float* sounds,
                    int n_sounds, void* ws, int64_t ws_bytes, int train, float* img_feat,
                    float* img_raw, float* snd_feat, float* snd_raw, void* stream) {
  Net* n = NET(net);
  if (!n->P || !n->PR || !ws) return VAR_ERR_ARG;
  cudaStream_t st = ST(stream);
  int rc = n->forward(images, image_kind, n_images, sounds, n_sounds, ws, ws_bytes, train != 0, st);
  if (rc) return rc;
  if (n_images > 0 && img_feat) {
    TailArgs a;
    memset(&a, 0, sizeof(a));
    n->fill_tail_weights(a, false);
    a.mode = var::TAIL_FWD; a.B = n_images; a.h_img = n->h_img; a.feat_img = img_feat;
    rc = var::tail_launch(a, st);
    if (rc) return rc;
  }
  if (n_sounds > 0 && snd_feat) {
    TailArgs a;
    memset(&a, 0, sizeof(a));
    n->fill_tail_weights(a, false);
    a.mode = var::TAIL_FWD; a.B = n_sounds; a.h_pos = n->h_snd; a.feat_pos = snd_feat;
    rc = var::tail_launch(a, st);
    if (rc) return rc;
  }
  if (n_images > 0 && img_raw) {
    const var::Layer& l = n->img_trunk.back();
    rc = var::nhwc_to_nchw(n->img_raw_nhwc, img_raw, n_images, l.P * l.Q, l.Cout, st);
    if (rc) return rc;
  }
  if (n_sounds > 0 && snd_raw) {
    if (n->has_gru) {
      if (cudaMemcpyAsync(snd_raw, n->snd_raw, (size_t)n_sounds * n->snd_raw_dim * 4,
                          cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return VAR_ERR_CUDA;
    } else {
      const var::Layer& l = n->snd_trunk.back();
      rc = var::nhwc_to_nchw(n->snd_raw, snd_raw, n_sounds, l.P * l.Q, l.Cout, st);
      if (rc) return rc;
    }
  }
  return VAR_OK;
}

int var_net_backward(void* net, const float* img_dfeat, const float* snd_dfeat, void* ws,
                     int64_t ws_bytes, void* stream) {
  Net* n = NET(net);
  if (!n->G || !ws) return VAR_ERR_ARG;
  cudaStream_t st = ST(stream);
  var::Arena ar(ws, ws_bytes);
  ar.used = n->fwd_used;
  float* dh_img = (n->n_img > 0 && img_dfeat) ? ar.alloc((long long)n->n_img * n->Kh_img) : nullptr;
  float* dh_snd = (n->n_snd > 0 && snd_dfeat) ? ar.alloc((long long)n->n_snd * n->Kh_snd) : nullptr;
  if (ar.overflow) return VAR_ERR_WORKSPACE;
  int rc;
  if (dh_img) {
    TailArgs a;
    memset(&a, 0, sizeof(a));
    n->fill_tail_weights(a, true);
    a.dW_snd = nullptr; a.db_snd = nullptr;
    a.mode = var::TAIL_BWD; a.B = n->n_img; a.h_img = n->h_img; a.dfeat_img = img_dfeat;
    a.dh_img = dh_img;
    rc = var::tail_launch(a, st);
    if (rc) return rc;
  }
  if (dh_snd) {
    TailArgs a;
    memset(&a, 0, sizeof(a));
    n->fill_tail_weights(a, true);
    a.dW_img = nullptr; a.db_img = nullptr;
    a.mode = var::TAIL_BWD; a.B = n->n_snd; a.h_pos = n->h_snd; a.dfeat_pos = snd_dfeat;
    a.dh_pos = dh_snd;
    rc = var::tail_launch(a, st);
    if (rc) return rc;
  }
  const long long keep = n->fwd_used;
  n->fwd_used = ar.used;
  rc = n->backward_from_dh(dh_img, dh_snd, ws, ws_bytes, st);
  n->fwd_used = keep;
  return rc;
}

int var_net_triplet_step(void* net, const void* images, int image_kind, const float* sounds, int B,
                         float margin, float loss_denominator, void* ws, int64_t ws_bytes,
                         float* loss, float* feats, void* stream) {
  Net* n = NET(net);
  if (!n->P || !n->PR || !n->G || !ws || B <= 0 || !loss) return VAR_ERR_ARG;
  cudaStream_t st = ST(stream);
  int rc = n->forward(images, image_kind, B, sounds, 2 * B, ws, ws_bytes, true, st);
  if (rc) return rc;
  var::Arena ar(ws, ws_bytes);
  ar.used = n->fwd_used;
  float* dh_img = ar.alloc((long long)B * n->Kh_img);
  float* dh_snd = ar.alloc((long long)2 * B * n->Kh_snd);
  if (ar.overflow) return VAR_ERR_WORKSPACE;
  TailArgs a;
  memset(&a, 0, sizeof(a));
  n->fill_tail_weights(a, true);
  a.mode = var::TAIL_TRIPLET; a.B = B;
  a.h_img = n->h_img; a.h_pos = n->h_snd; a.h_neg = n->h_snd + (long long)B * n->Kh_snd;
  a.margin = margin; a.grad_scale = 1.f / loss_denominator; a.loss_scale = 1.f / loss_denominator;
  a.loss = loss;
  if (feats) {
    a.feat_img = feats; a.feat_pos = feats + (long long)B * n->D;
    a.feat_neg = feats + (long long)2 * B * n->D;
  }
  a.dh_img = dh_img; a.dh_pos = dh_snd; a.dh_neg = dh_snd + (long long)B * n->Kh_snd;
  rc = var::tail_launch(a, st);
  if (rc) return rc;
  const long long keep = n->fwd_used;
  n->fwd_used = ar.used;
  rc = n->backward_from_dh(dh_img, dh_snd, ws, ws_bytes, st);
  n->fwd_used = keep;
  return rc;
}

int var_net_reward(void* net, const void* images, int image_kind, const float* goal_sounds,
                   const float* goal_feat_cached, const float* env_reward, int N, void* ws,
                   int64_t ws_bytes, float* img_feat, float* goal_feat, float* dot, float* reward,
                   void* stream) {
  Net* n = NET(net);
  if (!n->P || !n->PR || !ws || N <= 0 || (!goal_sounds && !goal_feat_cached)) return VAR_ERR_ARG;
  cudaStream_t st = ST(stream);
  int rc = n->forward(images, image_kind, N, goal_sounds, goal_sounds ? N : 0, ws, ws_bytes, false, st);
  if (rc) return rc;
  TailArgs a;
  memset(&a, 0, sizeof(a));
  n->fill_tail_weights(a, false);
  a.mode = var::TAIL_REWARD; a.B = N;
  a.h_img = n->h_img; a.h_pos = goal_sounds ? n->h_snd : nullptr;
  a.goal_feat_in = goal_feat_cached;
  a.feat_img = img_feat; a.feat_pos = goal_sounds ? goal_feat : nullptr;
  a.env_reward = env_reward; a.dot_out = dot; a.reward_out = reward;
  rc = var::tail_launch(a, st);
  if (rc) return rc;
  if (!goal_sounds && goal_feat && goal_feat != goal_feat_cached) {
    if (cudaMemcpyAsync(goal_feat, goal_feat_cached, (size_t)N * n->D * 4, cudaMemcpyDeviceToDevice,
                        st) != cudaSuccess)
      return VAR_ERR_CUDA;
  }
  return VAR_OK;
}

}  // extern "C"
