// engine.cu -- host runtime of the VITS2 inference engine + the C ABI declared in include/vtts.h.
//
// Replaces the one call `self.model.onnx.run(None, args)` (vosk_tts/synth.py:123-126), i.e. the trace
// of SynthesizerTrn.infer (training/vits2/models.py:1679-1704).  The launch sequence below follows that
// function stage by stage; every kernel is in kernels.cuh (fp32 FFMA) or conv_tc.cuh (tcgen05).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vtts.h"
#include "kernels.cuh"

using namespace vtts;

namespace {

struct Tensor {
  const float* p = nullptr;
  size_t n = 0;
};
struct ConvW {
  const float* w = nullptr;
  const float* b = nullptr;
  int Cin = 0, Cout = 0, k = 0, ldw = 0;
};
struct LnW {
  const float* g = nullptr;
  const float* b = nullptr;
};
struct EncLayerW {
  ConvW qkv, o, ffn1, ffn2;
  LnW ln1, ln2;
  const float* relk = nullptr;
  const float* relv = nullptr;
};
struct DdsW {
  const float *sep_w, *sep_b;
  LnW ln1, ln2;
  ConvW pw;
};
struct CfW {
  const float *pre_w, *pre_b;
  DdsW dds[3];
  ConvW proj;
};
struct FlowW {
  ConvW pre, post;
  EncLayerW tr;
  std::vector<ConvW> in, rsx, rss;
};
struct UpW {
  std::vector<ConvW> phase;
  std::vector<int> pad;
};
struct RbW {
  std::vector<ConvW> c1, c2;
};

template <typename T>
struct Buf {
  T* p = nullptr;
  size_t cap = 0;
};

struct Err {
  int code;
  std::string msg;
};

#define CK(call)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      std::ostringstream os__;                                                            \
      os__ << "CUDA error '" << cudaGetErrorString(e__) << "' at " << __FILE__ << ":" << __LINE__ << " in " #call; \
      throw Err{VTTS_ERR_CUDA, os__.str()};                                               \
    }                                                                                     \
  } while (0)

#define REQUIRE(cond, code, text)                     \
  do {                                                \
    if (!(cond)) throw Err{(code), std::string(text)}; \
  } while (0)

}  // namespace

struct vtts_engine {
  vtts_config cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::string err;
  uint64_t launches = 0;

  float* d_blob = nullptr;
  size_t blob_floats = 0;
  std::unordered_map<std::string, Tensor> tensors;

  // ---- weights (views into d_blob)
  bool has_g = false;
  const float *emb_g = nullptr, *cond_w = nullptr, *cond_b = nullptr, *enc_emb = nullptr, *dp_ea = nullptr;
  const float *istft_basis = nullptr, *pqmf = nullptr;
  int condR = 0, r_spk = -1, r_dp = 0, r_flow = 0;
  std::vector<EncLayerW> enc;
  ConvW enc_proj, dp_pre, dp_proj, dec_pre, dec_post;
  DdsW dp_dds[3];
  std::vector<CfW> cf;   // index n-2 for n = 2..dp_n_flows
  std::vector<FlowW> flow;
  std::vector<UpW> ups;
  std::vector<RbW> rbs;
  int hop = 0, up_total = 1;

  // ---- per-call state
  int B = 0, Ttok = 0, maxTok = 0, Tfrm = 0, maxFrm = 0;
  bool have_durations = false;
  float scales[3] = {0.f, 1.f, 0.f};
  uint64_t seed = 0;
  std::vector<int> h_tok_len, h_tok_off, h_frm_len, h_frm_off;

  // ---- workspace
  Buf<int> d_ids, d_tok_len, d_tok_off, d_sid, d_wceil, d_cum, d_frm_len, d_frm_off, d_ftok;
  Buf<float> d_condv, d_x, d_xb, d_qkv, d_ao, d_y, d_ffh, d_stats, d_dA, d_dB, d_dx, d_h29, d_za, d_zb, d_eps_dp;
  Buf<float> d_z, d_h, d_h1, d_wx, d_acts, d_skip, d_fqkv, d_fao, d_fy, d_ffh2, d_eps_z, d_d0, d_post, d_wav;
  std::vector<Buf<float>> d_stage;               // X_i
  std::vector<std::vector<Buf<float>>> d_xj, d_tmp;
  Buf<float> d_zp_dbg;                           // copy of z_p kept when debug_flags & 1
  int debug_flags = 0;
  Buf<char> h_pin;                               // pinned staging (host)
  cudaEvent_t ev[8] = {};
  float stage_ms[8] = {};
  bool ev_valid = false;

  // -------------------------------------------------------------------------------------------
  template <typename T>
  T* ensure(Buf<T>& b, size_t n) {
    if (n > b.cap) {
      if (b.p) CK(cudaFree(b.p));
      size_t cap = n + n / 4 + 256;
      CK(cudaMalloc(&b.p, cap * sizeof(T)));
      b.cap = cap;
    }
    return b.p;
  }
  char* ensure_pinned(size_t n) {
    if (n > h_pin.cap) {
      if (h_pin.p) {
        CK(cudaStreamSynchronize(stream));   // copies staged through the old buffer may still be in flight
        CK(cudaFreeHost(h_pin.p));
        h_pin.p = nullptr;
      }
      size_t cap = n + n / 4 + 4096;
      CK(cudaMallocHost(&h_pin.p, cap));
      h_pin.cap = cap;
    }
    return h_pin.p;
  }

  Tensor tensor(const std::string& name) {
    auto it = tensors.find(name);
    if (it == tensors.end()) throw Err{VTTS_ERR_WEIGHTS, "weight blob has no tensor '" + name + "'"};
    return it->second;
  }
  const float* vec(const std::string& name, size_t n) {
    Tensor t = tensor(name);
    if (t.n != n) {
      std::ostringstream os;
      os << "tensor '" << name << "' has " << t.n << " elements, expected " << n;
      throw Err{VTTS_ERR_WEIGHTS, os.str()};
    }
    return t.p;
  }
  ConvW conv(const std::string& name, int Cin, int Cout, int k) {
    ConvW c;
    c.Cin = Cin; c.Cout = Cout; c.k = k; c.ldw = (Cout + 3) / 4 * 4;
    c.w = vec(name + ".w", (size_t)k * Cin * c.ldw);
    c.b = vec(name + ".b", (size_t)c.ldw);
    REQUIRE(Cin % CV_CK == 0, VTTS_ERR_INVALID, "conv input channels must be a multiple of 16");
    return c;
  }
  LnW ln(const std::string& name, int C) { return LnW{vec(name + ".g", C), vec(name + ".b", C)}; }
  EncLayerW enc_layer(const std::string& p, int Hc, int Fc, int ks) {
    EncLayerW L;
    const int dk = Hc / cfg.n_heads, nrel = 2 * cfg.window_size + 1;
    L.qkv = conv(p + ".qkv", Hc, 3 * Hc, 1);
    L.o = conv(p + ".o", Hc, Hc, 1);
    L.relk = vec(p + ".relk", (size_t)nrel * dk);
    L.relv = vec(p + ".relv", (size_t)nrel * dk);
    L.ln1 = ln(p + ".ln1", Hc);
    L.ffn1 = conv(p + ".ffn1", Hc, Fc, ks);
    L.ffn2 = conv(p + ".ffn2", Fc, Hc, ks);
    L.ln2 = ln(p + ".ln2", Hc);
    return L;
  }
  DdsW dds(const std::string& p, int C, int k) {
    DdsW d;
    d.sep_w = vec(p + ".sep_w", (size_t)k * C);
    d.sep_b = vec(p + ".sep_b", C);
    d.ln1 = ln(p + ".ln1", C);
    d.pw = conv(p + ".pw", C, C, 1);
    d.ln2 = ln(p + ".ln2", C);
    return d;
  }

  void bind_weights();
  void launch_conv(const std::vector<ConvP>& ps, int rmul, const int* lens, const int* offs, int maxLen, int nB);
  void encoder_layer(const EncLayerW& L, float*& x, float*& xb, float* qkv, float* ao, float* y, float* ffh, int Hc, int Fc,
                     int ks, const int* lens, const int* offs, int maxLen, const float* vec_after, int vec_ld,
                     const float* cadd_after);
  void dds_stack(const DdsW* d, int C, int k, float*& a, float*& b, const int* lens, const int* offs, int maxLen);
  void phase1(const int* ids_packed_host, const int64_t* d_ids64, int t_max, const int64_t* d_sid64, const int* sid_host,
              const float* noise_dp, bool noise_on_device);
  void phase2(const float* noise_z, int z_ld, bool noise_on_device);
};

namespace {

ConvP mk(const ConvW& W, const float* x, int ldx, int xoff, float* y, int ldy, int yoff, int dil, int pad) {
  ConvP p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.w = W.w; p.bias = W.b; p.y = y;
  p.ldx = ldx; p.xoff = xoff; p.ldw = W.ldw; p.ldy = ldy; p.yoff = yoff;
  p.Cin = W.Cin; p.Cout = W.Cout; p.k = W.k; p.dil = dil; p.pad = pad;
  p.out_mul = 1; p.alpha = 1.f; p.pro = PRO_NONE;
  return p;
}

}  // namespace

void vtts_engine::bind_weights() {
  const vtts_config& c = cfg;
  const int H = c.hidden_channels, I = c.inter_channels, G = c.gin_channels, D = c.dp_filter_channels;
  REQUIRE(H % 32 == 0 && H <= 256 && D % 32 == 0 && D <= 256, VTTS_ERR_INVALID, "hidden/dp channels must be multiples of 32, <= 256");
  REQUIRE((H / c.n_heads) % 32 == 0 && H / c.n_heads <= 128, VTTS_ERR_INVALID, "head dim must be 32/64/96/128");
  REQUIRE(c.dp_num_bins <= SPL_MAXB, VTTS_ERR_INVALID, "too many spline bins");
  REQUIRE(c.dp_kernel_size % 2 == 1 && c.flow_kernel_size % 2 == 1, VTTS_ERR_INVALID, "odd kernels expected");
  REQUIRE(c.n_resblock_kernels <= CV_MAXP, VTTS_ERR_INVALID, "at most 4 resblocks per stage");
  has_g = c.n_speakers > 0 && G > 0;
  const int nl = c.flow_wn_layers, nf = c.flow_n_flows;
  if (has_g) {
    emb_g = vec("emb_g", (size_t)c.n_speakers * G);
    int r = 0;
    if (c.spk_cond_encoder) { r_spk = r; r += H; }
    r_dp = r; r += D;
    r_flow = r; r += nf * nl * 2 * H;
    condR = r;
    cond_w = vec("cond.w", (size_t)condR * G);
    cond_b = vec("cond.b", condR);
  }
  enc_emb = vec("enc.emb", (size_t)c.n_vocab * H);
  enc.clear();
  for (int i = 0; i < c.n_layers; ++i) enc.push_back(enc_layer("enc." + std::to_string(i), H, c.filter_channels, c.kernel_size));
  enc_proj = conv("enc.proj", H, 2 * I, 1);
  dp_pre = conv("dp.pre", H, D, 1);
  dp_proj = conv("dp.proj", D, D, 1);
  for (int i = 0; i < 3; ++i) dp_dds[i] = dds("dp.convs." + std::to_string(i), D, c.dp_kernel_size);
  cf.clear();
  for (int n = 2; n <= c.dp_n_flows; ++n) {
    CfW f;
    const std::string p = "dp.cf" + std::to_string(n);
    f.pre_w = vec(p + ".pre_w", D);
    f.pre_b = vec(p + ".pre_b", D);
    for (int i = 0; i < 3; ++i) f.dds[i] = dds(p + ".convs." + std::to_string(i), D, c.dp_kernel_size);
    f.proj = conv(p + ".proj", D, 3 * c.dp_num_bins - 1, 1);
    cf.push_back(f);
  }
  dp_ea = vec("dp.ea", 4);
  flow.clear();
  for (int f = 0; f < nf; ++f) {
    FlowW F;
    const std::string p = "flow." + std::to_string(f);
    F.pre = conv(p + ".pre", I / 2, H, 1);
    if (c.use_transformer_flows) F.tr = enc_layer(p + ".tr", H, H, c.flow_kernel_size);
    for (int i = 0; i < nl; ++i) {
      F.in.push_back(conv(p + ".in" + std::to_string(i), H, 2 * H, c.flow_kernel_size));
      if (i < nl - 1) F.rsx.push_back(conv(p + ".rsx" + std::to_string(i), H, H, 1));
      F.rss.push_back(conv(p + ".rss" + std::to_string(i), H, H, 1));
    }
    F.post = conv(p + ".post", H, I / 2, 1);
    flow.push_back(F);
  }
  dec_pre = conv("dec.pre", I, c.upsample_initial_channel, 7);
  ups.clear();
  rbs.clear();
  int ch = c.upsample_initial_channel;
  up_total = 1;
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int u = c.upsample_rates[i], K = c.upsample_kernel_sizes[i], p = (K - u) / 2;
    UpW U;
    for (int r = 0; r < u; ++r) {
      // polyphase split of ConvTranspose1d (see weights.convt_phases): taps per phase and left padding
      int d_min = -((r + p) / u);
      int d_max = (K - 1 - r - p) / u;
      U.phase.push_back(conv("dec.up" + std::to_string(i) + ".p" + std::to_string(r), ch, ch / 2, d_max - d_min + 1));
      U.pad.push_back(d_max);
    }
    ups.push_back(U);
    ch /= 2;
    up_total *= u;
    for (int j = 0; j < c.n_resblock_kernels; ++j) {
      RbW R;
      const std::string p2 = "dec.rb" + std::to_string(i * c.n_resblock_kernels + j);
      for (int d = 0; d < c.n_resblock_dilations; ++d) {
        if (c.resblock_type == 1) {
          R.c1.push_back(conv(p2 + ".c1." + std::to_string(d), ch, ch, c.resblock_kernel_sizes[j]));
          R.c2.push_back(conv(p2 + ".c2." + std::to_string(d), ch, ch, c.resblock_kernel_sizes[j]));
        } else {
          R.c1.push_back(conv(p2 + ".c." + std::to_string(d), ch, ch, c.resblock_kernel_sizes[j]));
        }
        REQUIRE(CV_TT + (c.resblock_kernel_sizes[j] - 1) * c.resblock_dilations[j][d] <= 32 * CV_XR, VTTS_ERR_INVALID,
                "resblock receptive field too wide for the conv tile");
      }
      rbs.push_back(R);
    }
  }
  if (c.decoder_type == 0) {
    const int cps = c.istft_n_fft + 2;
    dec_post = conv("dec.post", ch, c.subbands * cps, 7);
    istft_basis = vec("dec.istft", (size_t)cps * c.istft_n_fft);
    pqmf = vec("dec.pqmf", (size_t)c.subbands * 63);
    hop = up_total * c.istft_hop * c.subbands;
  } else {
    dec_post = conv("dec.post", ch, 1, 7);
    hop = up_total;
  }
}

void vtts_engine::launch_conv(const std::vector<ConvP>& ps, int rmul, const int* lens, const int* offs, int maxLen, int nB) {
  ConvBatch cb;
  memset(&cb, 0, sizeof(cb));
  REQUIRE(!ps.empty() && (int)ps.size() <= CV_MAXP, VTTS_ERR_INVALID, "bad grouped conv");
  int maxCout = 0, maxHalo = 0, maxL = 0;
  for (size_t i = 0; i < ps.size(); ++i) {
    cb.p[i] = ps[i];
    maxCout = std::max(maxCout, ps[i].Cout);
    maxHalo = std::max(maxHalo, (ps[i].k - 1) * ps[i].dil);
    maxL = std::max(maxL, maxLen * rmul + ps[i].in_extra);
    REQUIRE(CV_TT + (ps[i].k - 1) * ps[i].dil <= 32 * CV_XR, VTTS_ERR_INVALID, "conv tile halo too large");
  }
  cb.n = (int)ps.size();
  cb.rmul = rmul;
  int xw = (CV_TT + maxHalo + 7) / 8 * 8 + 2;
  cb.xw = xw;
  const size_t smem = (size_t)(CV_CK * xw + 2 * CV_CK * CV_TC) * sizeof(float);
  dim3 grid((maxL + CV_TT - 1) / CV_TT, (maxCout + CV_TC - 1) / CV_TC, nB * cb.n);
  if (grid.x == 0) return;
  conv_kernel<<<grid, CV_THREADS, smem, stream>>>(cb, lens, offs);
  CK(cudaGetLastError());
  ++launches;
}

// One relative-attention encoder layer (attentions.py:57-63): x <- LN2(x1 + FFN(x1)), x1 = LN1(x + MHA(x)).
void vtts_engine::encoder_layer(const EncLayerW& L, float*& x, float*& xb, float* qkv, float* ao, float* y, float* ffh, int Hc,
                                int Fc, int ks, const int* lens, const int* offs, int maxLen, const float* vec_after, int vec_ld,
                                const float* cadd_after) {
  const int nB = B;
  launch_conv({mk(L.qkv, x, Hc, 0, qkv, 3 * Hc, 0, 1, 0)}, 1, lens, offs, maxLen, nB);
  {
    const int dk = Hc / cfg.n_heads, nrel = 2 * cfg.window_size + 1;
    dim3 grid((maxLen + AT_QT - 1) / AT_QT, cfg.n_heads, nB);
    const size_t smem = (size_t)(2 * AT_KT * (dk + 1) + AT_QT * dk + nrel * dk + AT_QT * nrel + AT_QT * AT_KT) * sizeof(float);
    switch (dk / 32) {
      case 1: attn_kernel<1><<<grid, AT_THREADS, smem, stream>>>(qkv, 3 * Hc, ao, Hc, L.relk, L.relv, cfg.n_heads, cfg.window_size, lens, offs); break;
      case 2: attn_kernel<2><<<grid, AT_THREADS, smem, stream>>>(qkv, 3 * Hc, ao, Hc, L.relk, L.relv, cfg.n_heads, cfg.window_size, lens, offs); break;
      case 3: attn_kernel<3><<<grid, AT_THREADS, smem, stream>>>(qkv, 3 * Hc, ao, Hc, L.relk, L.relv, cfg.n_heads, cfg.window_size, lens, offs); break;
      default: attn_kernel<4><<<grid, AT_THREADS, smem, stream>>>(qkv, 3 * Hc, ao, Hc, L.relk, L.relv, cfg.n_heads, cfg.window_size, lens, offs); break;
    }
    CK(cudaGetLastError());
    ++launches;
  }
  launch_conv({mk(L.o, ao, Hc, 0, y, Hc, 0, 1, 0)}, 1, lens, offs, maxLen, nB);
  dim3 lg((maxLen + 3) / 4, nB);
  add_ln_kernel<<<lg, 128, 0, stream>>>(x, y, L.ln1.g, L.ln1.b, nullptr, nullptr, 0, xb, lens, offs, Hc);
  CK(cudaGetLastError());
  ++launches;
  {
    ConvP p = mk(L.ffn1, xb, Hc, 0, ffh, Fc, 0, 1, (ks - 1) / 2);
    p.epi = EPI_RELU;
    launch_conv({p}, 1, lens, offs, maxLen, nB);
  }
  launch_conv({mk(L.ffn2, ffh, Fc, 0, y, Hc, 0, 1, (ks - 1) / 2)}, 1, lens, offs, maxLen, nB);
  add_ln_kernel<<<lg, 128, 0, stream>>>(xb, y, L.ln2.g, L.ln2.b, cadd_after, vec_after, vec_ld, x, lens, offs, Hc);
  CK(cudaGetLastError());
  ++launches;
}

void vtts_engine::dds_stack(const DdsW* d, int C, int k, float*& a, float*& b, const int* lens, const int* offs, int maxLen) {
  int dil = 1;
  for (int i = 0; i < 3; ++i) {
    DdsP P;
    P.x = a; P.y = b;
    P.sep_w = d[i].sep_w; P.sep_b = d[i].sep_b;
    P.ln1g = d[i].ln1.g; P.ln1b = d[i].ln1.b;
    P.pw_w = d[i].pw.w; P.pw_b = d[i].pw.b; P.ldw = d[i].pw.ldw;
    P.ln2g = d[i].ln2.g; P.ln2b = d[i].ln2.b;
    P.C = C; P.k = k; P.dil = dil;
    dim3 grid((maxLen + DDS_TT - 1) / DDS_TT, B);
    dds_layer_kernel<<<grid, C, 0, stream>>>(P, lens, offs);
    CK(cudaGetLastError());
    ++launches;
    std::swap(a, b);
    dil *= k;
  }
}

// ---------------------------------------------------------------------------------------------------
// Phase 1: speaker vector, TextEncoder, StochasticDurationPredictor(reverse), durations.
// ---------------------------------------------------------------------------------------------------
void vtts_engine::phase1(const int* ids_packed_host, const int64_t* d_ids64, int t_max, const int64_t* d_sid64,
                         const int* sid_host, const float* noise_dp, bool noise_on_device) {
  const vtts_config& c = cfg;
  const int H = c.hidden_channels, I = c.inter_channels, D = c.dp_filter_channels, Fc = c.filter_channels;
  const size_t T = (size_t)Ttok;
  CK(cudaEventRecord(ev[0], stream));
  // ---- inputs
  int* tl = ensure(d_tok_len, B);
  int* to = ensure(d_tok_off, B + 1);
  int* ids = ensure(d_ids, T);
  int* sid = ensure(d_sid, B);
  {
    char* pin = ensure_pinned((size_t)(2 * B + 1 + B) * sizeof(int) + T * sizeof(int) + (size_t)B * 2 * t_max * sizeof(float) + 64);
    int* p_len = reinterpret_cast<int*>(pin);
    int* p_off = p_len + B;
    int* p_sid = p_off + B + 1;
    int* p_ids = p_sid + B;
    memcpy(p_len, h_tok_len.data(), B * sizeof(int));
    memcpy(p_off, h_tok_off.data(), (B + 1) * sizeof(int));
    CK(cudaMemcpyAsync(tl, p_len, B * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(to, p_off, (B + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    if (ids_packed_host) {
      memcpy(p_ids, ids_packed_host, T * sizeof(int));
      memcpy(p_sid, sid_host, B * sizeof(int));
      CK(cudaMemcpyAsync(ids, p_ids, T * sizeof(int), cudaMemcpyHostToDevice, stream));
      CK(cudaMemcpyAsync(sid, p_sid, B * sizeof(int), cudaMemcpyHostToDevice, stream));
    } else {
      dim3 g((maxTok + 127) / 128, B);
      pack_ids_kernel<<<g, 128, 0, stream>>>(d_ids64, t_max, ids, tl, to);
      CK(cudaGetLastError());
      cast_sid_kernel<<<(B + 127) / 128, 128, 0, stream>>>(d_sid64, sid, B);
      CK(cudaGetLastError());
      launches += 2;
    }
    if (noise_dp && !noise_on_device) {
      float* p_eps = reinterpret_cast<float*>(p_ids + T);
      memcpy(p_eps, noise_dp, (size_t)B * 2 * t_max * sizeof(float));
      float* de = ensure(d_eps_dp, (size_t)B * 2 * t_max);
      CK(cudaMemcpyAsync(de, p_eps, (size_t)B * 2 * t_max * sizeof(float), cudaMemcpyHostToDevice, stream));
      noise_dp = de;
    }
  }
  CK(cudaEventRecord(ev[1], stream));

  // ---- speaker conditioning (models.py:1680-1683)
  float* condv = nullptr;
  if (has_g) {
    condv = ensure(d_condv, (size_t)B * condR);
    dim3 g((condR + 7) / 8, B);
    cond_kernel<<<g, 256, c.gin_channels * sizeof(float), stream>>>(emb_g, sid, cond_w, cond_b, condv, c.gin_channels, condR, c.n_speakers);
    CK(cudaGetLastError());
    ++launches;
  }
  const float* spk_vec = (has_g && r_spk >= 0) ? condv + r_spk : nullptr;

  // ---- text encoder (models.py:317-326)
  float* x = ensure(d_x, T * H);
  float* xb = ensure(d_xb, T * H);
  float* qkv = ensure(d_qkv, T * 3 * H);
  float* ao = ensure(d_ao, T * H);
  float* y = ensure(d_y, T * H);
  float* ffh = ensure(d_ffh, T * Fc);
  float* stats = ensure(d_stats, T * 2 * I);
  {
    dim3 g(maxTok, B);
    embed_kernel<<<g, 64, 0, stream>>>(ids, enc_emb, x, tl, to, H, sqrtf((float)H), c.n_vocab,
                                       (spk_vec && c.cond_layer_idx == 0) ? spk_vec : nullptr, condR);
    CK(cudaGetLastError());
    ++launches;
  }
  for (int i = 0; i < c.n_layers; ++i) {
    const float* va = (spk_vec && c.cond_layer_idx == i + 1) ? spk_vec : nullptr;
    encoder_layer(enc[i], x, xb, qkv, ao, y, ffh, H, Fc, c.kernel_size, tl, to, maxTok, va, condR, nullptr);
  }
  launch_conv({mk(enc_proj, x, H, 0, stats, 2 * I, 0, 1, 0)}, 1, tl, to, maxTok, B);
  CK(cudaEventRecord(ev[2], stream));

  // ---- stochastic duration predictor, reverse (models.py:56-63, 93-101)
  float* dA = ensure(d_dA, T * D);
  float* dB = ensure(d_dB, T * D);
  float* dx = ensure(d_dx, T * D);
  float* h29 = ensure(d_h29, T * 32);
  float* za = ensure(d_za, T);
  float* zb = ensure(d_zb, T);
  {
    ConvP p = mk(dp_pre, x, H, 0, dA, D, 0, 1, 0);
    if (has_g) { p.cond = condv + r_dp; p.cond_ld = condR; }
    launch_conv({p}, 1, tl, to, maxTok, B);
  }
  {
    float *a = dA, *b = dB;
    dds_stack(dp_dds, D, c.dp_kernel_size, a, b, tl, to, maxTok);
    launch_conv({mk(dp_proj, a, D, 0, dx, D, 0, 1, 0)}, 1, tl, to, maxTok, B);
  }
  {
    dim3 g((maxTok + 127) / 128, B);
    dp_noise_kernel<<<g, 128, 0, stream>>>(noise_dp, t_max, seed, scales[2], za, zb, tl, to);
    CK(cudaGetLastError());
    ++launches;
  }
  float* cvar = zb;   // conditioning half (x0 after the Flip)
  float* tvar = za;   // transformed half (x1)
  const int nbins = c.dp_num_bins;
  REQUIRE(3 * nbins - 1 <= 32, VTTS_ERR_INVALID, "spline parameter row too wide");
  for (int n = c.dp_n_flows; n >= 2; --n) {
    const CfW& F = cf[n - 2];
    {
      dim3 g(maxTok, B);
      cf_pre_kernel<<<g, 128, 0, stream>>>(cvar, F.pre_w, F.pre_b, dx, dA, tl, to, D);
      CK(cudaGetLastError());
      ++launches;
    }
    float *a = dA, *b = dB;
    dds_stack(F.dds, D, c.dp_kernel_size, a, b, tl, to, maxTok);
    launch_conv({mk(F.proj, a, D, 0, h29, 32, 0, 1, 0)}, 1, tl, to, maxTok, B);
    {
      dim3 g((maxTok + 127) / 128, B);
      spline_inverse_kernel<<<g, 128, 0, stream>>>(h29, 32, tvar, nbins, c.dp_tail_bound, sqrtf((float)D), tl, to);
      CK(cudaGetLastError());
      ++launches;
    }
    std::swap(cvar, tvar);
  }
  // after the last Flip channel 0 is the half transformed last (== cvar after the swap)
  const float* zlast = (c.dp_n_flows >= 2) ? cvar : za;
  int* wceil = ensure(d_wceil, T);
  int* cum = ensure(d_cum, T);
  int* fl = ensure(d_frm_len, B);
  int* fo = ensure(d_frm_off, B + 1);
  duration_kernel<<<B, 256, 0, stream>>>(zlast, dp_ea, 0, 2, scales[1], wceil, cum, fl, tl, to);
  CK(cudaGetLastError());
  frame_offsets_kernel<<<1, 32, 0, stream>>>(fl, fo, B);
  CK(cudaGetLastError());
  launches += 2;
  CK(cudaEventRecord(ev[3], stream));
  {
    char* pin = ensure_pinned((size_t)(2 * B + 2) * sizeof(int) + T * sizeof(int) + 64);
    int* p_len = reinterpret_cast<int*>(pin);
    int* p_off = p_len + B;
    CK(cudaMemcpyAsync(p_len, fl, B * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(p_off, fo, (B + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    h_frm_len.assign(p_len, p_len + B);
    h_frm_off.assign(p_off, p_off + B + 1);
  }
  Tfrm = h_frm_off[B];
  maxFrm = 0;
  for (int b = 0; b < B; ++b) maxFrm = std::max(maxFrm, h_frm_len[b]);
  have_durations = true;
}

// ---------------------------------------------------------------------------------------------------
// Phase 2: alignment + prior sampling, flow^-1, decoder.
// ---------------------------------------------------------------------------------------------------
void vtts_engine::phase2(const float* noise_z, int z_ld, bool noise_on_device) {
  const vtts_config& c = cfg;
  const int H = c.hidden_channels, I = c.inter_channels, half = I / 2;
  const size_t F = (size_t)Tfrm;
  const int* tl = d_tok_len.p;
  const int* to = d_tok_off.p;
  const int* fl = d_frm_len.p;
  const int* fo = d_frm_off.p;
  CK(cudaEventRecord(ev[4], stream));
  if (noise_z && !noise_on_device) {
    const size_t n = (size_t)B * I * z_ld;
    float* de = ensure(d_eps_z, n);
    // caller memory may be pageable: stage through the pinned buffer
    char* pin = ensure_pinned(n * sizeof(float));
    memcpy(pin, noise_z, n * sizeof(float));
    CK(cudaMemcpyAsync(de, pin, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    noise_z = de;
  }
  float* z = ensure(d_z, F * I);
  int* ftok = ensure(d_ftok, F);
  {
    dim3 g(maxFrm, B);
    sample_prior_kernel<<<g, 64, 0, stream>>>(d_stats.p, I, d_cum.p, tl, to, fl, fo, noise_z, z_ld, seed, scales[0], z, ftok);
    CK(cudaGetLastError());
    ++launches;
  }
  if (debug_flags & 1) {
    float* zp = ensure(d_zp_dbg, F * I);
    CK(cudaMemcpyAsync(zp, z, F * I * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  }
  // ---- flow, reverse (models.py:750-757).  Flip (modules.py:272-279) is folded into the packed pre/post
  // weights: for a "flipped" layer x0 lives in physical channels [half, 2*half), x1 in [0, half).
  float* h = ensure(d_h, F * H);
  float* h1 = ensure(d_h1, F * H);
  float* wx = ensure(d_wx, F * H);
  float* acts = ensure(d_acts, F * H);
  float* skip = ensure(d_skip, F * H);
  float* fy = ensure(d_fy, F * H);
  float* fqkv = nullptr; float* fao = nullptr; float* ffh2 = nullptr;
  if (c.use_transformer_flows) {
    fqkv = ensure(d_fqkv, F * 3 * H);
    fao = ensure(d_fao, F * H);
    ffh2 = ensure(d_ffh2, F * H);
  }
  const int nf = c.flow_n_flows, nl = c.flow_wn_layers, fk = c.flow_kernel_size;
  for (int f = nf - 1; f >= 0; --f) {
    const FlowW& W = flow[f];
    const bool flipped = ((nf - f) % 2) == 1;
    const int x0off = flipped ? half : 0, x1off = flipped ? 0 : half;
    launch_conv({mk(W.pre, z, I, x0off, h, H, 0, 1, 0)}, 1, fl, fo, maxFrm, B);
    float* wn_in = h;
    if (c.use_transformer_flows) {
      // h = h + Encoder(h)  (models.py:377): the layer's last LN adds `h` back and lands in wx
      float* xa = h; float* xb2 = h1;
      // encoder_layer writes its result into `xa` (== h) -- we need h preserved for the residual, so run the
      // layer on explicit buffers instead of the ping-pong helper:
      launch_conv({mk(W.tr.qkv, h, H, 0, fqkv, 3 * H, 0, 1, 0)}, 1, fl, fo, maxFrm, B);
      {
        const int dk = H / c.n_heads, nrel = 2 * c.window_size + 1;
        dim3 grid((maxFrm + AT_QT - 1) / AT_QT, c.n_heads, B);
        const size_t smem = (size_t)(2 * AT_KT * (dk + 1) + AT_QT * dk + nrel * dk + AT_QT * nrel + AT_QT * AT_KT) * sizeof(float);
        switch (dk / 32) {
          case 1: attn_kernel<1><<<grid, AT_THREADS, smem, stream>>>(fqkv, 3 * H, fao, H, W.tr.relk, W.tr.relv, c.n_heads, c.window_size, fl, fo); break;
          case 2: attn_kernel<2><<<grid, AT_THREADS, smem, stream>>>(fqkv, 3 * H, fao, H, W.tr.relk, W.tr.relv, c.n_heads, c.window_size, fl, fo); break;
          case 3: attn_kernel<3><<<grid, AT_THREADS, smem, stream>>>(fqkv, 3 * H, fao, H, W.tr.relk, W.tr.relv, c.n_heads, c.window_size, fl, fo); break;
          default: attn_kernel<4><<<grid, AT_THREADS, smem, stream>>>(fqkv, 3 * H, fao, H, W.tr.relk, W.tr.relv, c.n_heads, c.window_size, fl, fo); break;
        }
        CK(cudaGetLastError());
        ++launches;
      }
      launch_conv({mk(W.tr.o, fao, H, 0, fy, H, 0, 1, 0)}, 1, fl, fo, maxFrm, B);
      dim3 lg((maxFrm + 3) / 4, B);
      add_ln_kernel<<<lg, 128, 0, stream>>>(xa, fy, W.tr.ln1.g, W.tr.ln1.b, nullptr, nullptr, 0, xb2, fl, fo, H);
      CK(cudaGetLastError());
      ++launches;
      {
        ConvP p = mk(W.tr.ffn1, xb2, H, 0, ffh2, H, 0, 1, (fk - 1) / 2);
        p.epi = EPI_RELU;
        launch_conv({p}, 1, fl, fo, maxFrm, B);
      }
      launch_conv({mk(W.tr.ffn2, ffh2, H, 0, fy, H, 0, 1, (fk - 1) / 2)}, 1, fl, fo, maxFrm, B);
      add_ln_kernel<<<lg, 128, 0, stream>>>(xb2, fy, W.tr.ln2.g, W.tr.ln2.b, h, nullptr, 0, wx, fl, fo, H);
      CK(cudaGetLastError());
      ++launches;
      wn_in = wx;
    }
    // WN (modules.py:148-176).  The hidden state is updated in place in `wn_in`.
    int dil = 1;
    for (int i = 0; i < nl; ++i) {
      {
        ConvP p = mk(W.in[i], wn_in, H, 0, acts, H, 0, dil, dil * (fk - 1) / 2);
        p.epi = EPI_GATE;
        if (has_g) { p.cond = d_condv.p + r_flow + (f * nl + i) * 2 * H; p.cond_ld = condR; }
        launch_conv({p}, 1, fl, fo, maxFrm, B);
      }
      ConvP ps = mk(W.rss[i], acts, H, 0, skip, H, 0, 1, 0);
      if (i > 0) { ps.res = skip; ps.ldr = H; ps.roff = 0; }
      if (i < nl - 1) {
        ConvP px = mk(W.rsx[i], acts, H, 0, wn_in, H, 0, 1, 0);
        px.res = wn_in; px.ldr = H; px.roff = 0;
        launch_conv({px, ps}, 1, fl, fo, maxFrm, B);
      } else {
        launch_conv({ps}, 1, fl, fo, maxFrm, B);
      }
      dil *= c.flow_dilation_rate;
    }
    {
      // x1 <- (x1 - post(h)) (mean_only; models.py:381-391)
      ConvP p = mk(W.post, skip, H, 0, z, I, x1off, 1, 0);
      p.alpha = -1.f;
      p.res = z; p.ldr = I; p.roff = x1off;
      launch_conv({p}, 1, fl, fo, maxFrm, B);
    }
  }
  CK(cudaEventRecord(ev[5], stream));

  // ---- decoder (models.py:1016-1054 / 872-891)
  int ch = c.upsample_initial_channel;
  float* cur = ensure(d_d0, F * ch);
  launch_conv({mk(dec_pre, z, I, 0, cur, ch, 0, 1, 3)}, 1, fl, fo, maxFrm, B);
  int rm = 1;
  const int nk = c.n_resblock_kernels, nd = c.n_resblock_dilations;
  if ((int)d_stage.size() < c.n_upsamples) {
    d_stage.resize(c.n_upsamples);
    d_xj.resize(c.n_upsamples);
    d_tmp.resize(c.n_upsamples);
    for (int i = 0; i < c.n_upsamples; ++i) { d_xj[i].resize(nk); d_tmp[i].resize(nk); }
  }
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int u = c.upsample_rates[i], ch2 = ch / 2;
    const size_t rows = F * rm * u;
    float* X = ensure(d_stage[i], rows * ch2);
    for (int r0 = 0; r0 < u; r0 += CV_MAXP) {
      std::vector<ConvP> ps;
      for (int r = r0; r < std::min(u, r0 + CV_MAXP); ++r) {
        ConvP p = mk(ups[i].phase[r], cur, ch, 0, X, ch2, 0, 1, ups[i].pad[r]);
        p.pro = PRO_LRELU; p.slope = 0.1f;
        p.out_mul = u; p.out_add = r;
        ps.push_back(p);
      }
      launch_conv(ps, rm, fl, fo, maxFrm, B);
    }
    rm *= u;
    ch = ch2;
    std::vector<float*> xj(nk), tmp(nk);
    for (int j = 0; j < nk; ++j) {
      xj[j] = ensure(d_xj[i][j], rows * ch);
      tmp[j] = ensure(d_tmp[i][j], rows * ch);
    }
    for (int d = 0; d < nd; ++d) {
      std::vector<ConvP> p1, p2;
      for (int j = 0; j < nk; ++j) {
        const RbW& R = rbs[i * nk + j];
        const int k = c.resblock_kernel_sizes[j], dl = c.resblock_dilations[j][d];
        const float* src = (d == 0) ? X : xj[j];
        if (c.resblock_type == 1) {
          ConvP a = mk(R.c1[d], src, ch, 0, tmp[j], ch, 0, dl, dl * (k - 1) / 2);
          a.pro = PRO_LRELU; a.slope = 0.1f;
          ConvP b2 = mk(R.c2[d], tmp[j], ch, 0, xj[j], ch, 0, 1, (k - 1) / 2);
          b2.pro = PRO_LRELU; b2.slope = 0.1f;
          b2.res = src; b2.ldr = ch; b2.roff = 0;
          p1.push_back(a);
          p2.push_back(b2);
        } else {
          ConvP a = mk(R.c1[d], src, ch, 0, (d == 0) ? xj[j] : tmp[j], ch, 0, dl, dl * (k - 1) / 2);
          a.pro = PRO_LRELU; a.slope = 0.1f;
          a.res = src; a.ldr = ch; a.roff = 0;
          p1.push_back(a);
        }
      }
      launch_conv(p1, rm, fl, fo, maxFrm, B);
      if (c.resblock_type == 1) {
        launch_conv(p2, rm, fl, fo, maxFrm, B);
      } else if (d > 0) {
        for (int j = 0; j < nk; ++j) std::swap(xj[j], tmp[j]);   // ResBlock2 ping-pong (halo reads forbid in-place)
      }
    }
    {
      REQUIRE(nk <= 3, VTTS_ERR_INVALID, "more than 3 resblocks per stage not supported");
      const long total4 = (long)(rows * ch / 4);
      mrf_mean_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, stream>>>(xj[0], nk > 1 ? xj[1] : nullptr, nk > 2 ? xj[2] : nullptr,
                                                                           std::min(nk, 3), X, total4);
      CK(cudaGetLastError());
      ++launches;
    }
    cur = X;
  }
  float* wav = ensure(d_wav, F * hop + 16);
  if (c.decoder_type == 0) {
    const int cps = c.istft_n_fft + 2, pc = c.subbands * cps;
    float* post = ensure(d_post, (F * rm + B) * pc);
    ConvP p = mk(dec_post, cur, ch, 0, post, pc, 0, 1, 3);
    p.pro = PRO_LRELU; p.slope = 0.01f;
    p.reflect = 1; p.in_extra = 1; p.out_seq_extra = 1;
    launch_conv({p}, rm, fl, fo, maxFrm, B);
    const int M = maxFrm * rm * c.istft_hop;
    dim3 g((M + TL_M - 1) / TL_M, B);
    const size_t smem = ((size_t)(TL_M / 4 + 16) * pc + (size_t)c.subbands * (TL_M + 2 * (62 / 2 / c.subbands + 1))) * sizeof(float);
    REQUIRE(c.istft_hop == 4 && c.istft_n_fft == 16, VTTS_ERR_INVALID, "iSTFT tail kernel is sized for n_fft=16, hop=4");
    istft_pqmf_kernel<<<g, TL_THREADS, smem, stream>>>(post, pc, istft_basis, pqmf, c.subbands, c.istft_n_fft, c.istft_hop, 63, rm, fl, fo, wav, 0, 1);
    CK(cudaGetLastError());
    ++launches;
  } else {
    ConvP p = mk(dec_post, cur, ch, 0, wav, 1, 0, 1, 3);
    p.pro = PRO_LRELU; p.slope = 0.01f;
    p.epi = EPI_TANH;
    launch_conv({p}, rm, fl, fo, maxFrm, B);
  }
  CK(cudaEventRecord(ev[6], stream));
}

// ===================================================================================================
// C ABI
// ===================================================================================================
namespace {

template <typename Fn>
int guarded(vtts_handle h, Fn fn) {
  if (!h) return VTTS_ERR_INVALID;
  std::lock_guard<std::mutex> lk(h->mu);
  try {
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) throw Err{VTTS_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)};
    fn();
    return VTTS_OK;
  } catch (const Err& e) {
    h->err = e.msg;
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    return VTTS_ERR_INVALID;
  }
}

void collect_timings(vtts_handle h) {
  // ev: 0 start, 1 after H2D, 2 after encoder, 3 after dp, 4 phase2 start, 5 after flow, 6 after decoder, 7 after D2H
  float t;
  cudaEventElapsedTime(&t, h->ev[1], h->ev[2]); h->stage_ms[0] = t;
  cudaEventElapsedTime(&t, h->ev[2], h->ev[3]); h->stage_ms[1] = t;
  cudaEventElapsedTime(&t, h->ev[4], h->ev[5]); h->stage_ms[2] = t;
  cudaEventElapsedTime(&t, h->ev[5], h->ev[6]); h->stage_ms[3] = t;
  cudaEventElapsedTime(&t, h->ev[0], h->ev[1]); h->stage_ms[4] = t;
  cudaEventElapsedTime(&t, h->ev[6], h->ev[7]); h->stage_ms[5] = t;
}

void setup_lengths(vtts_handle h, const int64_t* lengths, int B, int t_max) {
  REQUIRE(B >= 1 && B <= 16384 && t_max >= 1, VTTS_ERR_INVALID, "bad batch size / t_max");
  h->B = B;
  h->h_tok_len.resize(B);
  h->h_tok_off.resize(B + 1);
  int off = 0, mx = 0;
  for (int b = 0; b < B; ++b) {
    REQUIRE(lengths[b] >= 1 && lengths[b] <= t_max, VTTS_ERR_INVALID, "input_lengths must be in [1, t_max]");
    h->h_tok_len[b] = (int)lengths[b];
    h->h_tok_off[b] = off;
    off += (int)lengths[b];
    mx = std::max(mx, (int)lengths[b]);
  }
  h->h_tok_off[B] = off;
  h->Ttok = off;
  h->maxTok = mx;
  h->have_durations = false;
}

}  // namespace

extern "C" {

int vtts_create(const vtts_config* cfg, const float* blob, size_t blob_floats, const char* manifest, int blob_is_device,
                int device, vtts_handle* out) {
  if (!cfg || !blob || !manifest || !out) return VTTS_ERR_INVALID;
  *out = nullptr;
  vtts_engine* h = new vtts_engine();
  h->cfg = *cfg;
  h->device = device;
  *out = h;   // returned even on failure so that vtts_last_error() is readable; caller destroys it
  return guarded(h, [&] {
    REQUIRE(cfg->precision == 0, VTTS_ERR_INVALID, "precision mode not built into this library");
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) CK(cudaEventCreate(&e));
    h->blob_floats = blob_floats;
    CK(cudaMalloc(&h->d_blob, blob_floats * sizeof(float)));
    CK(cudaMemcpyAsync(h->d_blob, blob, blob_floats * sizeof(float), blob_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    std::istringstream is(manifest);
    std::string name;
    unsigned long long off, n;
    while (is >> name >> off >> n) {
      REQUIRE(off + n <= blob_floats, VTTS_ERR_WEIGHTS, "manifest entry exceeds the blob");
      h->tensors[name] = Tensor{h->d_blob + off, (size_t)n};
    }
    h->bind_weights();
    CK(cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaStreamSynchronize(h->stream));
  });
}

void vtts_destroy(vtts_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(h->d_blob);
  Buf<int>* ib[] = {&h->d_ids, &h->d_tok_len, &h->d_tok_off, &h->d_sid, &h->d_wceil, &h->d_cum, &h->d_frm_len, &h->d_frm_off, &h->d_ftok};
  for (auto* b : ib) fr(b->p);
  Buf<float>* fb[] = {&h->d_condv, &h->d_x, &h->d_xb, &h->d_qkv, &h->d_ao, &h->d_y, &h->d_ffh, &h->d_stats, &h->d_dA, &h->d_dB, &h->d_dx,
                      &h->d_h29, &h->d_za, &h->d_zb, &h->d_eps_dp, &h->d_z, &h->d_h, &h->d_h1, &h->d_wx, &h->d_acts, &h->d_skip, &h->d_fqkv,
                      &h->d_fao, &h->d_fy, &h->d_ffh2, &h->d_eps_z, &h->d_d0, &h->d_post, &h->d_wav};
  for (auto* b : fb) fr(b->p);
  for (auto& b : h->d_stage) fr(b.p);
  for (auto& v : h->d_xj) for (auto& b : v) fr(b.p);
  for (auto& v : h->d_tmp) for (auto& b : v) fr(b.p);
  if (h->h_pin.p) cudaFreeHost(h->h_pin.p);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* vtts_last_error(vtts_handle h) { return h ? h->err.c_str() : "null handle"; }

int vtts_durations(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid, int B, int t_max,
                   const float* scales, const float* noise_dp, uint64_t seed, int64_t* y_lengths, int32_t* durations) {
  if (!ids || !lengths || !sid || !scales || !y_lengths) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    setup_lengths(h, lengths, B, t_max);
    memcpy(h->scales, scales, 3 * sizeof(float));
    h->seed = seed;
    std::vector<int> packed(h->Ttok), sid32(B);
    for (int b = 0; b < B; ++b) {
      for (int t = 0; t < h->h_tok_len[b]; ++t) packed[h->h_tok_off[b] + t] = (int)ids[(size_t)b * t_max + t];
      sid32[b] = (int)sid[b];
    }
    h->phase1(packed.data(), nullptr, t_max, nullptr, sid32.data(), noise_dp, false);
    for (int b = 0; b < B; ++b) y_lengths[b] = h->h_frm_len[b];
    if (durations) {
      std::vector<int> wc(h->Ttok);
      CK(cudaMemcpyAsync(wc.data(), h->d_wceil.p, h->Ttok * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      for (int b = 0; b < B; ++b) {
        for (int t = 0; t < t_max; ++t)
          durations[(size_t)b * t_max + t] = t < h->h_tok_len[b] ? wc[h->h_tok_off[b] + t] : 0;
      }
    }
  });
}

int vtts_synthesize(vtts_handle h, const float* noise_z, int z_ld, float* wav, int64_t wav_ld, int32_t* frame_token, int idx_ld) {
  if (!wav) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    REQUIRE(h->have_durations, VTTS_ERR_STATE, "vtts_synthesize called without vtts_durations");
    REQUIRE((int64_t)h->maxFrm * h->hop <= wav_ld, VTTS_ERR_CAPACITY, "wav_ld is smaller than hop * max(y_lengths)");
    REQUIRE(!noise_z || z_ld >= h->maxFrm, VTTS_ERR_CAPACITY, "noise_z has fewer columns than max(y_lengths)");
    REQUIRE(!frame_token || idx_ld >= h->maxFrm, VTTS_ERR_CAPACITY, "frame_token has fewer columns than max(y_lengths)");
    h->phase2(noise_z, z_ld, false);
    const size_t nw = (size_t)h->Tfrm * h->hop;
    char* pin = h->ensure_pinned(nw * sizeof(float) + (size_t)h->Tfrm * sizeof(int) + 64);
    float* pw = reinterpret_cast<float*>(pin);
    int* pi = reinterpret_cast<int*>(pw + nw);
    CK(cudaMemcpyAsync(pw, h->d_wav.p, nw * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (frame_token) CK(cudaMemcpyAsync(pi, h->d_ftok.p, (size_t)h->Tfrm * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->ev[7], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < h->B; ++b) {
      memcpy(wav + (size_t)b * wav_ld, pw + (size_t)h->h_frm_off[b] * h->hop, (size_t)h->h_frm_len[b] * h->hop * sizeof(float));
      if (frame_token) memcpy(frame_token + (size_t)b * idx_ld, pi + h->h_frm_off[b], (size_t)h->h_frm_len[b] * sizeof(int));
    }
    collect_timings(h);
    h->have_durations = false;
  });
}

int vtts_durations_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid, int B, int t_max,
                       const float* scales, const float* d_noise_dp, uint64_t seed, int64_t* y_lengths_host) {
  if (!d_ids || !lengths_host || !d_sid || !scales || !y_lengths_host) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    setup_lengths(h, lengths_host, B, t_max);
    memcpy(h->scales, scales, 3 * sizeof(float));
    h->seed = seed;
    h->phase1(nullptr, d_ids, t_max, d_sid, nullptr, d_noise_dp, true);
    for (int b = 0; b < B; ++b) y_lengths_host[b] = h->h_frm_len[b];
  });
}

int vtts_synthesize_dev(vtts_handle h, const float* d_noise_z, int z_ld, float* d_wav, int64_t wav_ld) {
  if (!d_wav) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    REQUIRE(h->have_durations, VTTS_ERR_STATE, "vtts_synthesize_dev called without vtts_durations_dev");
    REQUIRE((int64_t)h->maxFrm * h->hop <= wav_ld, VTTS_ERR_CAPACITY, "wav_ld is smaller than hop * max(y_lengths)");
    REQUIRE(!d_noise_z || z_ld >= h->maxFrm, VTTS_ERR_CAPACITY, "noise_z has fewer columns than max(y_lengths)");
    h->phase2(d_noise_z, z_ld, true);
    for (int b = 0; b < h->B; ++b)
      CK(cudaMemcpyAsync(d_wav + (size_t)b * wav_ld, h->d_wav.p + (size_t)h->h_frm_off[b] * h->hop,
                         (size_t)h->h_frm_len[b] * h->hop * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaEventRecord(h->ev[7], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    collect_timings(h);
    h->have_durations = false;
  });
}

int vtts_hop(vtts_handle h) { return h ? h->hop : 0; }

int vtts_stage_timings(vtts_handle h, float* ms, int n) {
  if (!h || !ms) return VTTS_ERR_INVALID;
  for (int i = 0; i < n && i < 8; ++i) ms[i] = h->stage_ms[i];
  return VTTS_OK;
}

uint64_t vtts_kernel_launches(vtts_handle h) { return h ? h->launches : 0; }
void* vtts_stream(vtts_handle h) { return h ? (void*)h->stream : nullptr; }

int vtts_debug_flags(vtts_handle h, int flags) {
  if (!h) return VTTS_ERR_INVALID;
  h->debug_flags = flags;
  return VTTS_OK;
}

int vtts_debug_read(vtts_handle h, const char* name, float* out, size_t max_floats, size_t* n_out) {
  if (!name || !out || !n_out) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    const vtts_config& c = h->cfg;
    const std::string nm(name);
    const float* src = nullptr;
    size_t n = 0;
    const size_t T = h->Ttok, F = h->Tfrm;
    if (nm == "x") { src = h->d_x.p; n = T * c.hidden_channels; }
    else if (nm == "stats") { src = h->d_stats.p; n = T * 2 * c.inter_channels; }
    else if (nm == "dx") { src = h->d_dx.p; n = T * c.dp_filter_channels; }
    else if (nm == "za") { src = h->d_za.p; n = T; }
    else if (nm == "zb") { src = h->d_zb.p; n = T; }
    else if (nm == "condv") { src = h->d_condv.p; n = (size_t)h->B * h->condR; }
    else if (nm == "z_p") { src = h->d_zp_dbg.p; n = F * c.inter_channels; }
    else if (nm == "z") { src = h->d_z.p; n = F * c.inter_channels; }
    else if (nm == "d0") { src = h->d_d0.p; n = F * c.upsample_initial_channel; }
    else if (nm == "post") { src = h->d_post.p; n = (F * h->up_total + h->B) * c.subbands * (c.istft_n_fft + 2); }
    else if (nm.rfind("stage", 0) == 0) {
      const int i = atoi(nm.c_str() + 5);
      REQUIRE(i >= 0 && i < (int)h->d_stage.size(), VTTS_ERR_INVALID, "no such stage");
      int rm = 1, ch = c.upsample_initial_channel;
      for (int j = 0; j <= i; ++j) { rm *= c.upsample_rates[j]; ch /= 2; }
      src = h->d_stage[i].p; n = F * rm * ch;
    }
    REQUIRE(src != nullptr, VTTS_ERR_INVALID, "unknown or unallocated debug tensor");
    REQUIRE(n <= max_floats, VTTS_ERR_CAPACITY, "debug buffer too small");
    CK(cudaMemcpyAsync(out, src, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *n_out = n;
  });
}

float vtts_microbench(vtts_handle h, const char* what, int iters) {
  (void)h; (void)what; (void)iters;
  return -1.f;
}

}  // extern "C"
