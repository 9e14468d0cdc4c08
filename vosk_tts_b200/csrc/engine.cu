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
#include <chrono>
#include <condition_variable>
#include <thread>
#include <initializer_list>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vtts.h"
#include "kernels.cuh"
#include "conv_tc.cuh"
#include "mas.cuh"
#include "attn_tc.cuh"
#include "wn_tc.cuh"

using namespace vtts;

namespace {
constexpr int CONV_SMEM_MAX = 160 * 1024;   // dynamic shared memory opt-in of conv_kernel<G>


struct Tensor {
  const float* p = nullptr;
  size_t n = 0;
};
struct ConvW {
  const float* w = nullptr;
  const float* b = nullptr;
  int Cin = 0, Cout = 0, k = 0, ldw = 0;
};
struct TcW {     // split-bf16 copy of a conv weight: [k][Cout][Cin]
  const __nv_bfloat16* hi = nullptr;
  const __nv_bfloat16* lo = nullptr;
  const __nv_bfloat16* mid = nullptr;   // third plane of the exact 3-way split (precision mode 3, text encoder)
};
struct Planes {  // split-bf16 activation planes [rows][C]
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  __nv_bfloat16* mid = nullptr;         // third plane (exact 3-way split) or null
  int C = 0;
  long rows = 0;
};
struct TcSpec {  // one problem of a grouped tensor-core conv launch
  Planes in;
  TcW w;
  const float* bias = nullptr;
  int Cin = 0, Cout = 0, k = 1, dil = 1, pad = 0;
  float* y = nullptr; int ldy = 0;
  const float* res = nullptr; int ldr = 0;
  Planes out; float pl_slope = 1.f;
  int out_mul = 1, out_add = 0, in_extra = 0, out_seq_extra = 0;
  int yoff = 0, roff = 0, epi = 0, poff = 0;
  float alpha = 1.f;
  const float* cond = nullptr; int cond_ld = 0;
};
struct LnW {
  const float* g = nullptr;
  const float* b = nullptr;
};
struct EncLayerW {
  ConvW qkv, o, ffn1, ffn2;
  TcW t_qkv, t_o, t_ffn1, t_ffn2;
  LnW ln1, ln2;
  const float* relk = nullptr;
  const float* relv = nullptr;
  int heads = 1;
  // split-bf16 [16][128] tiles of the relative-position tables for the tcgen05 attention (null: FFMA attention only)
  const __nv_bfloat16 *rk_hi = nullptr, *rk_lo = nullptr, *rv_hi = nullptr, *rv_lo = nullptr;
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
  TcW t_qkv, t_o, t_ffn1, t_ffn2, t_post;
  std::vector<TcW> t_in, t_rsx, t_rss;
};
struct UpW {
  std::vector<ConvW> phase;
  std::vector<TcW> tphase;
  std::vector<int> pad;
};
struct RbW {
  std::vector<ConvW> c1, c2;
  std::vector<TcW> t1, t2;
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
  // The two-phase API (vtts_durations -> vtts_synthesize / vtts_flow) keeps per-handle state between two calls.  A thread
  // that has run vtts_durations owns the handle until its vtts_synthesize / vtts_flow has succeeded (or failed for good);
  // entry points called by OTHER threads wait for that instead of overwriting the pending durations.
  bool two_phase = false;
  std::thread::id owner;
  std::condition_variable cv;
  std::string err;
  uint64_t launches = 0;

  float* d_blob = nullptr;
  size_t blob_floats = 0;
  std::unordered_map<std::string, Tensor> tensors;

  // ---- weights (views into d_blob)
  bool has_g = false;
  const float *emb_g = nullptr, *cond_w = nullptr, *cond_b = nullptr, *enc_emb = nullptr, *dp_ea = nullptr;
  const float *istft_basis = nullptr, *pqmf = nullptr;
  int condR = 0, r_spk = -1, r_dp = 0, r_flow = 0, r_dec = -1;
  std::vector<EncLayerW> enc;
  ConvW enc_proj, dp_pre, dp_proj, dec_pre, dec_post;
  DdsW dp_dds[3];
  std::vector<CfW> cf;   // index n-2 for n = 2..dp_n_flows
  std::vector<FlowW> flow;
  std::vector<UpW> ups;
  std::vector<RbW> rbs;
  int hop = 0, up_total = 1;
  bool tc = false;                      // precision mode 1: tcgen05 path for the decoder convs
  TcW tc_pre, tc_post, tc_encproj;
  bool enc_on_tc = false;               // precision modes 2 / 3: the text encoder's convs on tcgen05 as well
  bool enc_three = false;               // mode 3: with the exact 3-way operand split
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn encode_tiled = nullptr;
  Buf<__nv_bfloat16> pl_pool[128];            // plane buffers (hi/lo pairs), indexed by the decoder code
  unsigned long long* tc_dbg = nullptr;     // device stamps buffer (microbench)
  TcBatch tc_batch;                         // launch_tc's parameter block (calls on one handle are serialised by `mu`)
  double tc_prof_flops = 0.0;
  uint64_t tc_prof_launches = 0;
  std::vector<cudaEvent_t> tc_prof_ev;
  size_t tc_prof_used = 0;

  // ---- per-call state
  // Ttok / maxTok / Tfrm / maxFrm are the BUCKETED row counts that size buffers, grids, tensor maps and captured graphs
  // (== the real ones with VTTS_BUCKETS=0); the kernels themselves read the true lengths / offsets from device memory.
  // real_* hold the true values for host-side copies; v_*_len are virtual per-utterance lengths (functions of the buckets
  // only) for the launch heuristics, so that a graph captured for a bucket is valid for every call that maps to it.
  int B = 0, Ttok = 0, maxTok = 0, Tfrm = 0, maxFrm = 0;
  int real_Ttok = 0, real_maxTok = 0, real_Tfrm = 0, real_maxFrm = 0;
  std::vector<int> v_tok_len, v_frm_len;
  bool use_buckets = true;
  // Speculative second phase (single utterances): the frame count is data dependent, but the kernels read it from device
  // memory and the host only needs an upper bound -- the length bucket -- to size grids and pick the graph.  The bound is
  // predicted from the frames-per-(token x length_scale) ratio of recent calls, phase 2 is enqueued for that bucket right
  // behind phase 1 WITHOUT the host waiting for the lengths, and repeated with the right bucket in the rare case the
  // prediction was too small.  A whole utterance is then two graph launches and one synchronisation; the phase-1 ->
  // host -> phase-2 round trip (~50 us, and the part of a step most exposed to host jitter) is gone.
  // OFF by default (VTTS_SPEC=1): measured on never-seen utterances with random speakers, whose frames-per-token ratio is
  // heavy tailed (1.3 .. 4), the predicted bucket is usually far too large (the predictor must cover the recent maximum or pay a
  // repeat), the launch heuristics then size split-K / tiles / attention for 3x the rows, and the GPU part of a call grows
  // from 1.7 to 2.5 ms -- while the round trip it removes was already hidden behind the prior projection (value: 1.665 vs
  // 1.666 ms with and without).  It only pays for repeated or very regular texts.
  bool use_spec = false;
  float spec_hist[16] = {};
  int spec_n = 0;
  float spec_ratio = 0.f;
  uint64_t spec_hits = 0, spec_misses = 0;
  double spec_units() const { return (double)real_maxTok * (double)std::max(0.05f, scales[1]); }
  double host_us[8] = {};             // host-side wall clock of the last vtts_infer call: [0] phase-1 enqueue, [1] phase-2 enqueue
                                      //   (+ copy-back enqueue), [2] wait for the stream, [3] copy-out, [4] total, [5] 1 = speculative hit, 2 = miss
  int spec_cap = 0;                   // length bucket of the speculative second phase of the current call (0: not speculating)
  float spec_margin = 1.08f;
  int spec_predict() const { return (int)std::ceil((double)spec_ratio * spec_margin * spec_units()) + 8; }
  void spec_learn() {
    spec_hist[spec_n++ % 16] = (float)((double)real_maxFrm / spec_units());
    float m = 0.f;
    for (int i = 0; i < std::min(spec_n, 16); ++i) m = std::max(m, spec_hist[i]);
    spec_ratio = m;
  }
  void assume_frames(int frames) {      // host-side frame shape from a prediction (B == 1)
    h_frm_len.assign(1, frames);
    h_frm_off = {0, frames};
    set_frame_shape();
  }
  bool read_published_lengths() {       // what frame_offsets_kernel wrote into mapped host memory; false: not there (yet)
    if (!h_map || h_map[0] != call_seq) return false;
    h_frm_len.assign(h_map + 1, h_map + 1 + B);
    h_frm_off.assign(h_map + 1 + B, h_map + 1 + 2 * B + 1);
    return true;
  }
  int eps_dp_ld = 0;                 // row pitch of the duration-predictor noise the phase-1 kernels read
  static int bucket_tok(int n) { return n <= 256 ? (n + 15) / 16 * 16 : (n + 63) / 64 * 64; }
  static int bucket_frm(int n) {
    return n <= 256 ? (n + 31) / 32 * 32 : (n <= 1024 ? (n + 63) / 64 * 64 : (n <= 4096 ? (n + 127) / 128 * 128 : (n + 511) / 512 * 512));
  }
  static void virtual_lens(std::vector<int>& v, int nB, int total_cap, int max_cap) {
    const int per = std::max(1, std::min(max_cap, (total_cap - (nB - 1) * SEQ_GAP + nB - 1) / nB));
    v.assign(nB, per);
  }
  void set_token_shape() {      // from h_tok_len / h_tok_off (real)
    real_Ttok = h_tok_off[B];
    real_maxTok = 0;
    for (int b = 0; b < B; ++b) real_maxTok = std::max(real_maxTok, h_tok_len[b]);
    if (use_buckets) {
      maxTok = bucket_tok(real_maxTok);
      Ttok = B == 1 ? maxTok : (real_Ttok + 63) / 64 * 64;
    } else { maxTok = real_maxTok; Ttok = real_Ttok; }
    virtual_lens(v_tok_len, B, Ttok, maxTok);
    if (!use_buckets) v_tok_len = h_tok_len;
  }
  void set_frame_shape() {      // from h_frm_len / h_frm_off (real)
    real_Tfrm = h_frm_off[B];
    real_maxFrm = 0;
    for (int b = 0; b < B; ++b) real_maxFrm = std::max(real_maxFrm, h_frm_len[b]);
    if (use_buckets) {
      maxFrm = bucket_frm(real_maxFrm);
      Tfrm = B == 1 ? std::max(maxFrm, real_Tfrm) : (real_Tfrm <= 8192 ? (real_Tfrm + 255) / 256 * 256 : (real_Tfrm + 1023) / 1024 * 1024);
    } else { maxFrm = real_maxFrm; Tfrm = real_Tfrm; }
    virtual_lens(v_frm_len, B, Tfrm, maxFrm);
    if (!use_buckets) v_frm_len = h_frm_len;
  }
  // plane buffers whose rows behind each utterance must be zeroed for this phase (see zero_tails_kernel)
  TailList tail;
  bool collecting = false;
  void begin_planes() { tail.n = 0; collecting = true; }
  void flush_tails(const int* lens, const int* offs) {
    collecting = false;
    if (tail.n == 0) return;
    klaunch(zero_tails_kernel, dim3(tail.n, B), dim3(128), (size_t)0, tail, lens, offs, B);
    CK(cudaGetLastError());
    ++launches;
  }
  bool have_durations = false;
  float scales[3] = {0.f, 1.f, 0.f};
  uint64_t seed = 0;
  std::vector<int> h_tok_len, h_tok_off, h_frm_len, h_frm_off;

  // ---- workspace
  Buf<int> d_ids, d_tok_len, d_tok_off, d_sid, d_wceil, d_cum, d_frm_len, d_frm_off, d_ftok, d_done_ctr, d_frm_len_real;
  Buf<float> d_condv, d_x, d_xb, d_qkv, d_ao, d_y, d_ffh, d_stats, d_dA, d_dB, d_dx, d_h29, d_za, d_zb, d_eps_dp;
  Buf<float> d_z, d_h, d_h1, d_wx, d_acts, d_skip, d_fqkv, d_fao, d_fy, d_ffh2, d_eps_z, d_d0, d_post, d_wav;
  std::vector<Buf<float>> d_stage;               // X_i
  std::vector<std::vector<Buf<float>>> d_xj, d_tmp;
  // per-launch profiling of the conv kernel family (bench.py roofline): event pairs around each launch
  bool profiling = false;
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  double prof_flops = 0.0;
  uint64_t prof_launches = 0;
  Buf<float> d_zp_dbg;                           // copy of z_p kept when debug_flags & 1
  int debug_flags = 0;
  Buf<char> h_pin, h_pin_in, h_pin_len, h_pin_z;  // pinned staging (host): outputs, phase-1 inputs, lengths, noise_z
  Buf<float> d_prm;                              // per-call scalars (see kernels.cuh prm_seed)
  int* h_map = nullptr;                          // mapped pinned host memory: [0] sequence flag, lengths, offsets
  int* d_map = nullptr;                          // its device alias
  size_t map_cap = 0;
  int call_seq = 0;
  bool use_poll = true;
  // CUDA graphs: a call shape seen before is captured once and replayed (launch-bound at batch 1)
  struct GraphEntry { cudaGraphExec_t exec = nullptr; uint64_t gen = 0; uint64_t used = 0; uint64_t nlaunch = 0; int seen = 0; };
  std::map<std::vector<long long>, GraphEntry> graphs;
  uint64_t ws_gen = 0, graph_clock = 0, graph_replays = 0;
  bool capture_on_first = true;
  bool capturing = false, use_graphs = true, last_graphed = false, use_pdl = true;    // programmatic dependent launch (VTTS_PDL=0 turns it off)
  int conv_max_s = 8, conv_target = 120, conv_max_g = 4, conv_big_g = 1, tc_tall = 0, tc_baseoff = 0, tc_bn = 0, attn_rows = 0, tc_mc = 0, tc_split = 0, conv_min_g = 1, tc_min_steps = 2, conv_auto_g = 4, attn_split = 1, tc_persist = 1, tc_persist_min = 1, n_sm = 148, tc_coal = 0, tc_dbgskip = 0, tc_wmc = 0;
  int tc_cluster_cap[2][3] = {{0, 0, 0}, {0, 0, 0}};   // co-resident clusters of 2/4/8 conv_tc CTAs, [BN 64/128][log2(S)-1]   // multicast measured slower   // tuning knobs (env VTTS_CONV_MAXS / _TARGET / _MAXG)
  cudaEvent_t ev[8] = {};
  cudaStream_t side[3] = {};               // branch streams of the decoder's independent resblock chains (forked / joined with events)
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {};
  int attn_tc_mode = 1;                    // tcgen05 attention where the qkv conv runs on tensor cores: 0 never, 1 when throughput bound, 2 always
  int mrf_heavy_first = 1;
  int mrf_branch = 0;                      // VTTS_MRF_BRANCH=1: one stream per resblock chain (measured slower: 1.74 vs 1.62 ms)
  float stage_ms[8] = {};
  bool ev_valid = false;

  // -------------------------------------------------------------------------------------------
  template <typename T>
  T* ensure(Buf<T>& b, size_t n) {
    if (n > b.cap) {
      REQUIRE(!capturing, VTTS_ERR_STATE, "workspace growth during graph capture");
      if (b.p) CK(cudaFree(b.p));
      size_t cap = n + n / 4 + 256;
      CK(cudaMalloc(&b.p, cap * sizeof(T)));
      b.cap = cap;
      ++ws_gen;                                // captured graphs hold the old pointers
    }
    return b.p;
  }
  char* ensure_pinned(Buf<char>& hb, size_t n) {
    if (n > hb.cap) {
      REQUIRE(!capturing, VTTS_ERR_STATE, "staging growth during graph capture");
      if (hb.p) {
        CK(cudaStreamSynchronize(stream));   // copies staged through the old buffer may still be in flight
        CK(cudaFreeHost(hb.p));
        hb.p = nullptr;
      }
      size_t cap = n + n / 4 + 4096;
      CK(cudaMallocHost(&hb.p, cap));
      hb.cap = cap;
      ++ws_gen;
    }
    return hb.p;
  }
  char* ensure_pinned(size_t n) { return ensure_pinned(h_pin, n); }

  // Runs `enqueue` (which only enqueues work on `stream`) eagerly the first time a shape key is seen -- that run also
  // performs every workspace growth -- and right behind it records the same work into a CUDA graph (capture only, no second
  // execution), so that the SECOND call of a bucket already replays.  The key is the full tuple of everything the
  // enqueued work depends on besides device-resident data (phase tag, batch, length buckets, noise mode, raw pointers of
  // the *_dev entry points): entries are compared on the tuple itself, so there is no hash collision to replay a wrong
  // graph on.  The cache is bounded (LRU, checked on every insertion).
  static constexpr size_t GRAPH_CACHE_MAX = 48;
  template <typename Fn>
  void run_graphed(std::initializer_list<long long> key_il, Fn&& enqueue) {
    last_graphed = false;
    if (!use_graphs || profiling || debug_flags) { enqueue(); return; }
    const std::vector<long long> key(key_il);
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      if (graphs.size() >= GRAPH_CACHE_MAX) {       // evict the least recently used entry before inserting
        auto old = graphs.begin();
        for (auto k = graphs.begin(); k != graphs.end(); ++k) if (k->second.used < old->second.used) old = k;
        if (old->second.exec) cudaGraphExecDestroy(old->second.exec);
        graphs.erase(old);
      }
      it = graphs.emplace(key, GraphEntry{}).first;
    }
    GraphEntry& g = it->second;
    if (g.exec && g.gen != ws_gen) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; g.seen = 0; }
    g.used = ++graph_clock;
    if (g.exec) {
      CK(cudaGraphLaunch(g.exec, stream));
      ++graph_replays;
      launches += g.nlaunch;
      last_graphed = true;
      return;
    }
    const bool first = (g.seen++ == 0);
    if (first) {                                // first sighting: eager (also performs any workspace growth) ...
      enqueue();
      if (!capture_on_first) return;
    }
    const uint64_t gen0 = ws_gen, l0 = launches;      // ... then capture
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    capturing = true;
    try {
      enqueue();
    } catch (...) {
      capturing = false;
      cudaStreamEndCapture(stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    capturing = false;
    CK(cudaStreamEndCapture(stream, &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) throw Err{VTTS_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)};
    g.exec = exec;
    g.gen = gen0;
    g.nlaunch = launches - l0;
    if (first) { launches = l0; return; }       // (the eager run above already did the work)
    CK(cudaGraphLaunch(g.exec, stream));
    ++graph_replays;
    last_graphed = true;
  }
  // Every kernel goes through here: programmatic dependent launch lets the next kernel's prologue overlap this one's tail.
  template <typename... KArgs, typename... Args>
  void klaunch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = use_pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(args)...));
  }
  static uint64_t mix(uint64_t h, uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); return h; }

  Tensor tensor(const std::string& name) {
    auto it = tensors.find(name);
    if (it == tensors.end()) throw Err{VTTS_ERR_WEIGHTS, "weight blob has no tensor '" + name + "'"};
    looked_up.push_back(name);
    return it->second;
  }
  std::vector<std::string> looked_up;          // tensors the engine bound, in first-use order
  Buf<PrefRange> d_pref;                       // L2 prefetch list for this precision mode
  int n_pref = 0, n_pref_phase1 = 0;
  bool use_prefetch = false;                   // measured: no gain on B200 (weights are not the latency bottleneck)
  void build_prefetch_list();
  const float* vec(const std::string& name, size_t n) {
    Tensor t = tensor(name);
    if (t.n != n) {
      std::ostringstream os;
      os << "tensor '" << name << "' has " << t.n << " elements, expected " << n;
      throw Err{VTTS_ERR_WEIGHTS, os.str()};
    }
    return t.p;
  }
  // need_w = false: the conv runs on tcgen05 from its split-bf16 copy in this precision mode; the fp32 copy is bound only if
  // the blob happens to carry it (weights.pack(precision=...) leaves it out of the one-time weight broadcast)
  ConvW conv(const std::string& name, int Cin, int Cout, int k, bool need_w = true) {
    ConvW c;
    c.Cin = Cin; c.Cout = Cout; c.k = k; c.ldw = (Cout + 3) / 4 * 4;
    if (need_w || tensors.count(name + ".w")) c.w = vec(name + ".w", (size_t)k * Cin * c.ldw);
    c.b = vec(name + ".b", (size_t)c.ldw);
    REQUIRE(Cin % CV_CK == 0, VTTS_ERR_INVALID, "conv input channels must be a multiple of 16");
    return c;
  }
  LnW ln(const std::string& name, int C) { return LnW{vec(name + ".g", C), vec(name + ".b", C)}; }
  EncLayerW enc_layer(const std::string& p, int Hc, int Fc, int ks, int heads, bool need_w = true) {
    EncLayerW L;
    L.heads = heads;
    const int dk = Hc / heads, nrel = 2 * cfg.window_size + 1;
    L.qkv = conv(p + ".qkv", Hc, 3 * Hc, 1, need_w);
    L.o = conv(p + ".o", Hc, Hc, 1, need_w);
    L.relk = vec(p + ".relk", (size_t)nrel * dk);
    L.relv = vec(p + ".relv", (size_t)nrel * dk);
    L.ln1 = ln(p + ".ln1", Hc);
    L.ffn1 = conv(p + ".ffn1", Hc, Fc, ks, need_w);
    L.ffn2 = conv(p + ".ffn2", Fc, Hc, ks, need_w);
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

  void bind_rel_tc(EncLayerW& L, const std::string& p) {
    if (!tensors.count(p + ".rkh")) return;       // (blob packed without the tables: the FFMA attention is used)
    L.rk_hi = reinterpret_cast<const __nv_bfloat16*>(vec(p + ".rkh", 16 * 128 / 2));
    L.rk_lo = reinterpret_cast<const __nv_bfloat16*>(vec(p + ".rkl", 16 * 128 / 2));
    L.rv_hi = reinterpret_cast<const __nv_bfloat16*>(vec(p + ".rvh", 16 * 128 / 2));
    L.rv_lo = reinterpret_cast<const __nv_bfloat16*>(vec(p + ".rvl", 16 * 128 / 2));
  }
  TcW tcw(const std::string& name, int Cin, int Cout, int k, bool three = false) {
    TcW t;
    const size_t n = (size_t)k * Cout * Cin / 2;
    t.hi = reinterpret_cast<const __nv_bfloat16*>(vec(name + (three ? ".t3h" : ".th"), n));
    t.lo = reinterpret_cast<const __nv_bfloat16*>(vec(name + (three ? ".t3l" : ".tl"), n));
    if (three) t.mid = reinterpret_cast<const __nv_bfloat16*>(vec(name + ".t3m", n));
    REQUIRE(Cin % TC_BK == 0, VTTS_ERR_INVALID, "tensor-core conv needs input channels in multiples of 64");
    return t;
  }
  Buf<__nv_bfloat16> pl_pool_mid[64];
  Planes planes(int slot, long units, int rm, int C, int extra = 0, bool three = false) {
    Planes p;
    const long rows = units * rm + (extra ? (long)B * extra : 0);
    p.C = C; p.rows = rows;
    const size_t n = (size_t)rows * C + 64;
    REQUIRE(slot >= 0 && 2 * slot + 1 < 128, VTTS_ERR_INVALID, "plane slot out of range");
    const size_t cap_hi = pl_pool[2 * slot].cap, cap_lo = pl_pool[2 * slot + 1].cap;
    p.hi = ensure(pl_pool[2 * slot], n);
    p.lo = ensure(pl_pool[2 * slot + 1], n);
    // fresh device memory may hold NaN bit patterns: rows a kernel never writes (beyond an utterance's end inside the last
    // tile) are multiplied by exact zeros in the attention's P V product, so they must at least be finite
    if (pl_pool[2 * slot].cap != cap_hi) CK(cudaMemsetAsync(p.hi, 0, pl_pool[2 * slot].cap * sizeof(__nv_bfloat16), stream));
    if (pl_pool[2 * slot + 1].cap != cap_lo) CK(cudaMemsetAsync(p.lo, 0, pl_pool[2 * slot + 1].cap * sizeof(__nv_bfloat16), stream));
    if (three) {
      const size_t cap_mid = pl_pool_mid[slot].cap;
      p.mid = ensure(pl_pool_mid[slot], n);
      if (pl_pool_mid[slot].cap != cap_mid) CK(cudaMemsetAsync(p.mid, 0, pl_pool_mid[slot].cap * sizeof(__nv_bfloat16), stream));
    }
    if (collecting) {   // rows behind each utterance must read as zero through TMA: zeroed by one zero_tails launch per phase
      REQUIRE(tail.n < ZT_MAXP && C % 8 == 0, VTTS_ERR_INVALID, "too many plane buffers in one phase");
      TailList::E& e = tail.e[tail.n++];
      e.hi = p.hi; e.lo = p.lo; e.mid = p.mid; e.C = C; e.rm = rm; e.extra = extra; e.rows_cap = (int)rows;
    }
    return p;
  }
  CUtensorMap make_map(const void* base, int C, long rows, int box_rows);
  bool attn_tc_ok(const EncLayerW& L, int Hc) const;
  bool attn_use_tc(const EncLayerW& L, int Hc, const int* lens, int maxLen) const;
  void launch_attn_tc(const Planes& qkv, float* ao, Planes* pl, const EncLayerW& L, int Hc, const int* lens, const int* offs, int maxLen);
  void launch_tc(const std::vector<TcSpec>& ps, int rmul, const int* lens, const int* offs, int maxLen, int nB);
  // plane buffers of the frame-resolution stages: allocated (and their tails zeroed by ONE zero_tails launch) before the
  // first kernel of the phase
  struct FlowPl { Planes ph, pao, ph1, pff, pwx, pwx2, pacts, pskip, pqkv; } flp;
  // One cluster kernel per WN layer (wn_tc.cuh).  Correct (all goldens), but measured SLOWER than the two launches it
  // replaces: 36 vs 22 us per layer at batch 1 (1.94 vs 1.69 ms per utterance) -- without split-K the 15 k-steps of the gated
  // conv run serially in each CTA (8.7 us), the gate for 48 channels per thread costs 4.3 us, the DSMEM all-gather of the
  // 96 KB acts tile 4.1 us plus two cluster barriers, and the second epilogue 10.8 us (profiles/r2_timeline_wn_fused_nopdl.txt).
  // Kept behind VTTS_WN_FUSED=1.
  int wn_fused = 0;
  bool wn_fused_ok() const {
    const int H = cfg.hidden_channels;
    return wn_fused && (H == 64 || H == 128 || H == 192) && cfg.flow_kernel_size % 2 == 1;
  }
  void launch_wn_fused(const FlowW& W, int f, int i, const Planes& xin, const Planes& xout, float* x, float* skip, const Planes& pskip,
                       int dil, const int* fl, const int* fo);
  struct DecPl { Planes pz, cur; std::vector<Planes> px, nxt; std::vector<std::vector<Planes>> pj, pt; } dcp;
  void alloc_flow_planes();
  void alloc_decoder_planes();
  void decoder_tc(float* z, const int* fl, const int* fo, bool pz_ready);
  void flow_tc(float* z, const int* fl, const int* fo, bool emit_pz);
  void launch_attn(const float* qkv, float* ao, const EncLayerW& L, int Hc, const int* lens, const int* offs, int maxLen, Planes* pl);
  void bind_weights();
  void launch_conv(const std::vector<ConvP>& ps, int rmul, const int* lens, const int* offs, int maxLen, int nB);
  void encoder_layer(const EncLayerW& L, float*& x, float*& xb, float* qkv, float* ao, float* y, float* ffh, int Hc, int Fc,
                     int ks, const int* lens, const int* offs, int maxLen, const float* vec_after, int vec_ld,
                     const float* cadd_after);
  void dds_stack(const DdsW* d, int C, int k, float*& a, float*& b, const int* lens, const int* offs, int maxLen,
                 const float* x0 = nullptr, const float* pre_w = nullptr, const float* pre_b = nullptr, const float* cond = nullptr);
  struct P1Pin { int *len, *off, *sid, *ids; float *prm, *eps; };
  P1Pin p1_layout(int t_max, bool eps);
  void stage1(const int* ids_packed_host, const int* sid_host, int t_max, const float* noise_dp_host);
  void finish1();
  void phase1(const int* ids_packed_host, const int64_t* d_ids64, int t_max, const int64_t* d_sid64, const int* sid_host,
              const float* noise_dp, bool noise_on_device);
  void phase2(const float* noise_z, int z_ld, bool noise_on_device, bool run_decoder = true);
  void stage_noise_z(const float* noise_z, int z_ld) {       // caller memory (maybe pageable) -> pinned [B][I][maxFrm]
    const int I = cfg.inter_channels;
    float* pin = reinterpret_cast<float*>(ensure_pinned(h_pin_z, (size_t)B * I * maxFrm * sizeof(float)));
    const size_t ncopy = (size_t)std::min(z_ld, maxFrm);
    for (long r = 0; r < (long)B * I; ++r) memcpy(pin + r * maxFrm, noise_z + r * (long)z_ld, ncopy * sizeof(float));
  }
  void decode(float* z, const int* fl, const int* fo, bool planes_ready = false, bool pz_ready = false);
  bool have_latent = false;
  Buf<int> d_chunk;                              // [len, off, off_end] of the chunk being decoded
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

// Which bound tensors does a call actually read in this precision mode?  (fp32 `.w` copies of convs that run on
// tcgen05 are skipped, and so are the bf16 `.th/.tl` copies of convs that stay on the FFMA pipe.)
void vtts_engine::build_prefetch_list() {
  auto on_tc = [&](const std::string& nm) {
    if (nm.rfind("enc.", 0) == 0) return cfg.precision >= 2;
    if (nm.rfind("flow.", 0) == 0 || nm.rfind("dec.", 0) == 0) return cfg.precision >= 1;
    return false;
  };
  std::vector<PrefRange> p1, p2;
  std::unordered_map<std::string, bool> seen;
  for (const std::string& nm : looked_up) {
    if (seen[nm]) continue;
    seen[nm] = true;
    const bool is_w = nm.size() > 2 && nm.compare(nm.size() - 2, 2, ".w") == 0;
    const bool is_t = nm.size() > 3 && (nm.compare(nm.size() - 3, 3, ".th") == 0 || nm.compare(nm.size() - 3, 3, ".tl") == 0);
    if (is_w && tensors.count(nm.substr(0, nm.size() - 2) + ".th") && on_tc(nm)) continue;
    if (is_t && !on_tc(nm)) continue;
    if (nm == "emb_g") continue;                 // one row is read
    const Tensor& t = tensors[nm];
    PrefRange r{reinterpret_cast<const char*>(t.p), (unsigned long long)t.n * sizeof(float)};
    const bool phase1 = nm.rfind("enc.", 0) == 0 || nm.rfind("dp.", 0) == 0 || nm.rfind("cond.", 0) == 0;
    (phase1 ? p1 : p2).push_back(r);
  }
  n_pref_phase1 = (int)p1.size();
  p1.insert(p1.end(), p2.begin(), p2.end());
  n_pref = (int)p1.size();
  PrefRange* d = ensure(d_pref, p1.size() + 1);
  CK(cudaMemcpyAsync(d, p1.data(), p1.size() * sizeof(PrefRange), cudaMemcpyHostToDevice, stream));
  CK(cudaStreamSynchronize(stream));
}

void vtts_engine::bind_weights() {
  const vtts_config& c = cfg;
  const int H = c.hidden_channels, I = c.inter_channels, G = c.gin_channels, D = c.dp_filter_channels;
  REQUIRE(H % 32 == 0 && H <= 256 && D % 32 == 0 && D <= 256, VTTS_ERR_INVALID, "hidden/dp channels must be multiples of 32, <= 256");
  REQUIRE((H / c.n_heads) % 32 == 0 && H / c.n_heads <= 128, VTTS_ERR_INVALID, "head dim must be 32/64/96/128");
  const int fheads = c.flow_n_heads > 0 ? c.flow_n_heads : 2;      // models.py:355: the flow's pre_transformer always has 2 heads
  REQUIRE(!c.use_transformer_flows || ((H / fheads) % 32 == 0 && H / fheads <= 128 && H % fheads == 0), VTTS_ERR_INVALID,
          "flow head dim must be 32/64/96/128");
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
    if (c.decoder_type == 1 && tensors.count("cond.w") && tensors["cond.w"].n == (size_t)(r + c.upsample_initial_channel) * G) {
      r_dec = r; r += c.upsample_initial_channel;      // plain Generator: x = conv_pre(z) + cond(g)  (models.py:874-875)
    }
    condR = r;
    cond_w = vec("cond.w", (size_t)condR * G);
    cond_b = vec("cond.b", condR);
  }
  enc_emb = vec("enc.emb", (size_t)c.n_vocab * H);
  enc.clear();
  enc_on_tc = c.precision >= 2 && H % TC_BK == 0 && c.filter_channels % TC_BK == 0;
  enc_three = enc_on_tc && c.precision == 3;     // exact 3-way split: durations as exact as the fp32 FFMA path
  for (int i = 0; i < c.n_layers; ++i) enc.push_back(enc_layer("enc." + std::to_string(i), H, c.filter_channels, c.kernel_size, c.n_heads, !enc_on_tc));
  enc_proj = conv("enc.proj", H, 2 * I, 1, !enc_on_tc);
  if (enc_on_tc) {
    for (int i = 0; i < c.n_layers; ++i) {
      const std::string p = "enc." + std::to_string(i);
      enc[i].t_qkv = tcw(p + ".qkv", H, 3 * H, 1, enc_three);
      enc[i].t_o = tcw(p + ".o", H, H, 1, enc_three);
      enc[i].t_ffn1 = tcw(p + ".ffn1", H, c.filter_channels, c.kernel_size, enc_three);
      enc[i].t_ffn2 = tcw(p + ".ffn2", c.filter_channels, H, c.kernel_size, enc_three);
      if (!enc_three) bind_rel_tc(enc[i], p);      // (mode 3 keeps the encoder's attention on the fp32 pipe)
    }
    tc_encproj = tcw("enc.proj", H, 2 * I, 1, enc_three);
  }
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
    const bool fw = !(c.precision >= 1 && H % TC_BK == 0);       // fp32 copies needed? (no: the whole flow but its pre conv is on tcgen05)
    F.pre = conv(p + ".pre", I / 2, H, 1);
    if (c.use_transformer_flows) F.tr = enc_layer(p + ".tr", H, H, c.flow_kernel_size, fheads, fw);
    for (int i = 0; i < nl; ++i) {
      F.in.push_back(conv(p + ".in" + std::to_string(i), H, 2 * H, c.flow_kernel_size, fw));
      if (i < nl - 1) F.rsx.push_back(conv(p + ".rsx" + std::to_string(i), H, H, 1, fw));
      F.rss.push_back(conv(p + ".rss" + std::to_string(i), H, H, 1, fw));
    }
    F.post = conv(p + ".post", H, I / 2, 1, fw);
    if (c.precision >= 1 && H % TC_BK == 0) {
      const int fk = c.flow_kernel_size;
      if (c.use_transformer_flows) {
        F.t_qkv = tcw(p + ".tr.qkv", H, 3 * H, 1);
        F.t_o = tcw(p + ".tr.o", H, H, 1);
        F.t_ffn1 = tcw(p + ".tr.ffn1", H, H, fk);
        F.t_ffn2 = tcw(p + ".tr.ffn2", H, H, fk);
        bind_rel_tc(F.tr, p + ".tr");
      }
      for (int i = 0; i < nl; ++i) {
        F.t_in.push_back(tcw(p + ".in" + std::to_string(i), H, 2 * H, fk));
        if (i < nl - 1) F.t_rsx.push_back(tcw(p + ".rsx" + std::to_string(i), H, H, 1));
        F.t_rss.push_back(tcw(p + ".rss" + std::to_string(i), H, H, 1));
      }
      F.t_post = tcw(p + ".post", H, I / 2, 1);
    }
    flow.push_back(F);
  }
  tc = c.precision >= 1 && c.precision <= 3;
  dec_pre = conv("dec.pre", I, c.upsample_initial_channel, 7, !tc);
  if (tc) {
    REQUIRE(c.decoder_type == 0 && c.resblock_type == 1, VTTS_ERR_INVALID, "tensor-core mode supports the MB-iSTFT / ResBlock1 decoder");
    tc_pre = tcw("dec.pre", I, c.upsample_initial_channel, 7);
  }
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
      U.phase.push_back(conv("dec.up" + std::to_string(i) + ".p" + std::to_string(r), ch, ch / 2, d_max - d_min + 1, !tc));
      if (tc) U.tphase.push_back(tcw("dec.up" + std::to_string(i) + ".p" + std::to_string(r), ch, ch / 2, d_max - d_min + 1));
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
          R.c1.push_back(conv(p2 + ".c1." + std::to_string(d), ch, ch, c.resblock_kernel_sizes[j], !tc));
          R.c2.push_back(conv(p2 + ".c2." + std::to_string(d), ch, ch, c.resblock_kernel_sizes[j], !tc));
          if (tc) {
            R.t1.push_back(tcw(p2 + ".c1." + std::to_string(d), ch, ch, c.resblock_kernel_sizes[j]));
            R.t2.push_back(tcw(p2 + ".c2." + std::to_string(d), ch, ch, c.resblock_kernel_sizes[j]));
          }
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
    dec_post = conv("dec.post", ch, c.subbands * cps, 7, !tc);
    if (tc) tc_post = tcw("dec.post", ch, c.subbands * cps, 7);
    istft_basis = vec("dec.istft", (size_t)cps * c.istft_n_fft);
    pqmf = vec("dec.pqmf", (size_t)c.subbands * 63);
    hop = up_total * c.istft_hop * c.subbands;
  } else {
    dec_post = conv("dec.post", ch, 1, 7);
    hop = up_total;
  }
}

CUtensorMap vtts_engine::make_map(const void* base, int C, long rows, int box_rows) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Err{VTTS_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"};
  return m;
}

// Grouped tensor-core conv launch (conv_tc.cuh).  One CTA = 128 rows x 64 output channels of one problem.
void vtts_engine::launch_tc(const std::vector<TcSpec>& ps, int rmul, const int* lens, const int* offs, int maxLen, int nB) {
  // 128-wide channel tiles halve the activation traffic and run the MMA at its smem-operand optimum, but halve the CTA
  // count: used once the launch still fills the machine (batched calls), or when forced (VTTS_TC_BN).
  int BN = 64;
  {
    const std::vector<int>& hl = (lens == d_tok_len.p) ? v_tok_len : v_frm_len;
    long ctas128 = 0;
    bool wide = true;
    for (const TcSpec& q : ps) {
      if (q.Cout < 128) wide = false;
      for (int b = 0; b < nB; ++b) ctas128 += (long)((hl[b] * rmul + q.in_extra + TC_BM - 1) / TC_BM) * ((q.Cout + 127) / 128);
    }
    if (wide && ctas128 >= 2 * 148) BN = 128;
    if (tc_bn == 64 || tc_bn == 128) BN = tc_bn;
  }
  REQUIRE(!ps.empty() && (int)ps.size() <= TC_MAXP, VTTS_ERR_INVALID, "bad grouped tensor-core conv");
  TcBatch& tb = tc_batch;   // per-engine scratch (2.6 KB: kept off the stack frame of every caller)
  memset(&tb, 0, sizeof(tb));
  int maxCout = 0, maxL = 0, maxNR = TC_BM;
  // "Tall" activation tiles (one TMA box of 128 + (k-1)*dil rows per channel chunk, the taps read it through row-shifted
  // descriptors) cut the L2 -> shared-memory traffic of a k-step from A + W to A/k + W.  On machine-filling launches the
  // mainloop is bound by exactly that traffic (64 KB per k-step and SM at 128-wide tiles against 768 tensor-pipe cycles),
  // so they use it by default (VTTS_TC_TALL: 0 auto, 1 always, -1 never); single-utterance launches prefer split-K.
  bool np3 = false;
  long grid_tiles = 0;
  for (const TcSpec& q : ps) {
    if (q.w.mid && q.in.mid) np3 = true;
    grid_tiles += (long)((maxLen * rmul + q.in_extra + TC_BM - 1) / TC_BM) * ((q.Cout + BN - 1) / BN) * nB;
  }
  const bool big = tc_persist == 2 || (tc_persist && grid_tiles > (long)tc_persist_min * n_sm);   // (2: forced, for tests)
  bool tall = tc_tall > 0 || (tc_tall == 0 && big);
  for (const TcSpec& q : ps) {
    const int nr = TC_BM + (q.k - 1) * q.dil;
    if (nr > 192) tall = false;              // shared-memory budget of the activation ring (and TMA box <= 256)
    maxNR = std::max(maxNR, nr);
  }
  if (np3) tall = false;
  if (tall) {
    const int ab = (maxNR * 128 + 1023) / 1024 * 1024;
    const int need = BN == 128 ? tc_smem_bytes<128>(ab, 2, 2, tc_coal ? 3 : tc_wst<128>(), tc_coal ? TC_STAGE_BYTES : 0)
                               : tc_smem_bytes<64>(ab, 2, 2, tc_wst<64>(), tc_coal ? TC_STAGE_BYTES : 0);
    if (need > 227 * 1024) tall = false;
  }
  if (!tall) maxNR = TC_BM;
  // TMA multicast of the activation tile across the channel-tile CTAs of a cluster: only when every problem of the
  // launch has the same number of channel tiles (no CTA of a cluster may drop out) and the tile is not "tall"
  int cn = 1;
  if (tc_mc && !tall) {
    int ny = -1;
    bool same = true;
    for (const TcSpec& q : ps) {
      const int n = (q.Cout + BN - 1) / BN;
      if (ny < 0) ny = n; else if (n != ny) same = false;
    }
    if (same) cn = (ny % 4 == 0) ? 4 : (ny % 2 == 0 ? 2 : 1);
    if (tc_mc == 2 && same && ny % 2 == 0) cn = 2;          // VTTS_TC_MULTICAST=2: pairs only
  }
  // Cluster split-K for launches that leave most SMs idle (single utterances): S CTAs share the k-steps of one tile.
  // The widest split whose clusters are all co-resident wins; 128-wide channel tiles are taken when they allow a
  // wider split than 64-wide ones (same k-steps per CTA-step, half the CTAs per tile row).
  int split = 1;
  if (!tall && tc_split != 1) {
    const std::vector<int>& hl = (lens == d_tok_len.p) ? v_tok_len : v_frm_len;
    long active[2] = {0, 0};
    int minsteps = 1 << 30;
    bool wide = true;
    for (const TcSpec& q : ps) {
      minsteps = std::min(minsteps, q.Cin / TC_BK * q.k);
      if (q.Cout < 128) wide = false;
      for (int b = 0; b < nB; ++b) {
        const long rt = (hl[b] * rmul + q.in_extra + TC_BM - 1) / TC_BM;
        active[0] += rt * ((q.Cout + 63) / 64);
        active[1] += rt * ((q.Cout + 127) / 128);
      }
    }
    const int cap = tc_split > 1 ? tc_split : 8;
    auto best = [&](int wi) {
      for (int S = 8, si = 2; S >= 2; S >>= 1, --si)
        if (S <= cap && minsteps >= tc_min_steps * S && active[wi] * S <= (long)tc_cluster_cap[wi][si] * S) return S;
      return 1;
    };
    if (tc_bn == 64 || tc_bn == 128) split = best(tc_bn == 128);
    else if (BN == 64) {
      const int s64 = best(0), s128 = wide ? best(1) : 1;
      if (s128 > s64) { BN = 128; split = s128; } else split = s64;
    }
  }
  if (split > 1) cn = 1;
  int np = 0;
  for (const TcSpec& q : ps) {
    const int n = (q.w.mid && q.in.mid) ? 3 : 2;
    REQUIRE(np == 0 || np == n, VTTS_ERR_INVALID, "grouped tensor-core conv mixes 2- and 3-plane problems");
    np = n;
  }
  if (np == 3) { tall = false; cn = 1; }
  tb.np = np;
  tb.ast = (np == 3 || tall) ? 2 : (BN == 128 ? tc_ast<128>() : tc_ast<64>());
  // launches without split-K finish their tiles through a 32 KB transposition buffer (coalesced epilogue, conv_tc.cuh); with
  // 128-wide channel tiles it takes the place of the fourth weight stage
  long tiles_all = 0;                        // the launch's tile space (one wave or less: the one-tile-per-CTA kernel)
  {
    int mc = 0, ml = 0;
    for (const TcSpec& q : ps) { mc = std::max(mc, q.Cout); ml = std::max(ml, maxLen * rmul + q.in_extra); }
    tiles_all = (long)((ml + TC_BM - 1) / TC_BM) * ((mc + BN - 1) / BN) * nB * (long)ps.size() * split;
  }
  const bool one_wave = tiles_all <= 148 && !(tc_persist == 2 && split == 1 && cn == 1);
  tb.coal = (split == 1 && !one_wave && tc_coal) ? 1 : 0;
  const int stage_bytes = tb.coal ? TC_STAGE_BYTES : 0;
  tb.wst = np == 3 ? (BN == 128 ? 2 : 3) : (BN == 128 ? (tb.coal ? 3 : tc_wst<128>()) : tc_wst<64>());
  tb.split = split;
  tb.cn = cn;
  tb.tall = tall ? 1 : 0;
  // persistent launches: CTA pairs share every weight tile through TMA multicast (conv_tc.cuh)
  const int wmc = (big && tc_wmc && split == 1 && cn == 1 && np == 2 && n_sm % 2 == 0) ? 2 : 1;
  tb.baseoff = tc_baseoff;
  tb.dbgskip = tc_dbgskip;
  tb.a_bytes = (maxNR * 128 + 1023) / 1024 * 1024;
  for (size_t i = 0; i < ps.size(); ++i) {
    const TcSpec& q = ps[i];
    TcProblem& P = tb.p[i];
    const int box_rows = tall ? TC_BM + (q.k - 1) * q.dil : TC_BM / cn;
    P.a_hi = make_map(q.in.hi, q.in.C, q.in.rows, box_rows);
    P.a_lo = make_map(q.in.lo, q.in.C, q.in.rows, box_rows);
    P.w_hi = make_map(q.w.hi, q.Cin, (long)q.k * q.Cout, BN / wmc);
    P.w_lo = make_map(q.w.lo, q.Cin, (long)q.k * q.Cout, BN / wmc);
    if (np == 3) {
      P.a_mid = make_map(q.in.mid, q.in.C, q.in.rows, box_rows);
      P.w_mid = make_map(q.w.mid, q.Cin, (long)q.k * q.Cout, BN);
    }
    P.p_mid = q.out.mid;
    P.bias = q.bias;
    P.res = q.res; P.ldr = q.ldr; P.roff = q.roff;
    P.y = q.y; P.ldy = q.ldy; P.yoff = q.yoff;
    P.cond = q.cond; P.cond_ld = q.cond_ld; P.epi = q.epi;
    P.p_hi = q.out.hi; P.p_lo = q.out.lo; P.ldp = q.out.C; P.poff = q.poff;
    P.Cin = q.Cin; P.Cout = q.Cout; P.k = q.k; P.dil = q.dil; P.pad = q.pad;
    P.out_mul = q.out_mul; P.out_add = q.out_add; P.in_extra = q.in_extra; P.out_seq_extra = q.out_seq_extra;
    P.alpha = q.alpha; P.pl_slope = q.pl_slope;
    REQUIRE(q.in.C == q.Cin, VTTS_ERR_INVALID, "plane width must equal the conv input channels");
    maxCout = std::max(maxCout, q.Cout);
    maxL = std::max(maxL, maxLen * rmul + q.in_extra);
  }
  tb.n = (int)ps.size();
  tb.rmul = rmul;
  tb.dbg = tc_dbg;
  dim3 grid((maxL + TC_BM - 1) / TC_BM, (maxCout + BN - 1) / BN, nB * tb.n * split);
  if (grid.x == 0) return;
  tb.wpre = one_wave ? 1 : 0;
  // machine-filling launches: one resident CTA per SM walks the tile space (conv_tc.cuh)
  tb.gx = (int)grid.x; tb.gy = (int)grid.y; tb.gz = (int)grid.z;
  tb.persist = 0;
  tb.wmc = 1;
  if (split == 1 && cn == 1 && !tb.wpre && (tc_persist == 2 || (tc_persist && (long)grid.x * grid.y * grid.z > (long)tc_persist_min * n_sm))) {
    tb.persist = 1;
    tb.wmc = wmc;
    grid = dim3((unsigned)n_sm, 1, 1);
  }
  REQUIRE(tb.persist || wmc == 1, VTTS_ERR_INVALID, "tensor-core conv: weight multicast planned for a launch that is not persistent");
  if (profiling) {
    if (tc_prof_used + 2 > tc_prof_ev.size()) {
      tc_prof_ev.resize(tc_prof_used + 2);
      CK(cudaEventCreate(&tc_prof_ev[tc_prof_used]));
      CK(cudaEventCreate(&tc_prof_ev[tc_prof_used + 1]));
    }
    for (const TcSpec& q : ps)
      for (int b = 0; b < nB; ++b)
        tc_prof_flops += 2.0 * ((double)((lens == d_tok_len.p) ? h_tok_len[b] : h_frm_len[b]) * rmul + q.in_extra) * q.Cout * q.Cin * q.k;   // (true lengths)
    ++tc_prof_launches;
    CK(cudaEventRecord(tc_prof_ev[tc_prof_used], stream));
  }
  {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = grid; lc.blockDim = dim3(TC_THREADS); lc.stream = stream;
    lc.dynamicSmemBytes = BN == 128 ? tc_smem_bytes<128>(tb.a_bytes, tb.np, tb.ast, tb.wst, stage_bytes) : tc_smem_bytes<64>(tb.a_bytes, tb.np, tb.ast, tb.wst, stage_bytes);
    REQUIRE(lc.dynamicSmemBytes <= 227 * 1024, VTTS_ERR_INVALID, "tensor-core conv: shared-memory budget exceeded");
    cudaLaunchAttribute at[2];
    int na = 0;
    if (cn > 1 || split > 1 || tb.wmc > 1) {
      at[na].id = cudaLaunchAttributeClusterDimension;
      at[na].val.clusterDim.x = tb.wmc; at[na].val.clusterDim.y = cn; at[na].val.clusterDim.z = split;
      ++na;
    }
    if (use_pdl) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    lc.attrs = at; lc.numAttrs = na;
    // single-wave launches all use the split-capable instantiation (also with split == 1): alternating between two kernel
    // images costs instruction-cache misses on every launch of a latency-bound chain
    if (split > 1 || tb.wpre) {
      const bool dyn = tb.np != 2 || tb.ast != (BN == 128 ? tc_ast<128>() : tc_ast<64>()) || tb.wst != (BN == 128 ? tc_wst<128>() : tc_wst<64>());
      if (dyn) {
        if (BN == 128) CK(cudaLaunchKernelEx(&lc, conv_tc_kernel<128, true, true>, tb, lens, offs));
        else CK(cudaLaunchKernelEx(&lc, conv_tc_kernel<64, true, true>, tb, lens, offs));
      } else {                    // the default launches carry the descriptor without the third-plane tensor maps
        // (a one-slot descriptor for the single-conv launches was tried as well: alternating between two kernel images on the
        //  chain cost more than the 2 KB of parameters saved -- conv_tc 723 -> 765 us in-graph)
        const auto tl = tc_lite<TC_MAXP>(tb);
        if (BN == 128) CK(cudaLaunchKernelEx(&lc, conv_tc_kernel<128, true, false, TC_MAXP>, tl, lens, offs));
        else CK(cudaLaunchKernelEx(&lc, conv_tc_kernel<64, true, false, TC_MAXP>, tl, lens, offs));
      }
    } else {                      // more than one wave of tiles: the persistent kernel (also runs them one per CTA when tb.persist == 0)
      if (BN == 128) CK(cudaLaunchKernelEx(&lc, conv_tc_persist_kernel<128>, tb, lens, offs));
      else CK(cudaLaunchKernelEx(&lc, conv_tc_persist_kernel<64>, tb, lens, offs));
    }
  }
  CK(cudaGetLastError());
  if (profiling) {
    CK(cudaEventRecord(tc_prof_ev[tc_prof_used + 1], stream));
    tc_prof_used += 2;
  }
  ++launches;
}

// Attention on the tensor cores (attn_tc.cuh): q, k, v come as the split-bf16 planes the qkv conv's epilogue wrote.
bool vtts_engine::attn_tc_ok(const EncLayerW& L, int Hc) const {
  const int dk = Hc / L.heads;
  return attn_tc_mode > 0 && L.rk_hi != nullptr && dk % 32 == 0 && dk <= 128 && 2 * cfg.window_size + 1 <= ATC_RS && (3 * Hc) % 8 == 0;
}
// Which attention kernel for this launch?  The tensor-core kernel wins as soon as the launch is throughput bound (batches,
// long utterances: 8.8x at 4765 frames, 2.8x on the flow of a 64-utterance batch).  A single short utterance is latency
// bound -- a 128-row tcgen05 tile walks its 2-4 key tiles serially while the split-KV FFMA kernel spreads 4 query rows x 4
// key segments over 16 warps of ~80 CTAs -- and keeps the FFMA kernel (measured at 162 frames: 6 vs 19 us per launch in the
// graph).  VTTS_ATTN_TC = 0 never / 1 this rule / 2 always.
bool vtts_engine::attn_use_tc(const EncLayerW& L, int Hc, const int* lens, int maxLen) const {
  if (!attn_tc_ok(L, Hc)) return false;
  if (attn_tc_mode >= 2) return true;
  const std::vector<int>& hl = (lens == d_tok_len.p) ? v_tok_len : v_frm_len;
  long ctas = 0;
  for (int b = 0; b < B; ++b) ctas += (long)((hl[b] + ATS_ROWS - 1) / ATS_ROWS) * L.heads;
  const int mt = (maxLen + AT_KT - 1) / AT_KT;
  const bool split_kv_fits = attn_split && (attn_rows == 0 || attn_rows == 1) && mt <= ATS_MAXT && ctas <= 148;
  return !split_kv_fits;
}

void vtts_engine::launch_attn_tc(const Planes& qkv, float* ao, Planes* pl, const EncLayerW& L, int Hc, const int* lens, const int* offs, int maxLen) {
  const int dk = Hc / L.heads;
  AttnTcParams ap;
  memset(&ap, 0, sizeof(ap));
  REQUIRE(qkv.C == 3 * Hc, VTTS_ERR_INVALID, "qkv planes must hold 3*H channels");
  ap.q_hi = make_map(qkv.hi, qkv.C, qkv.rows, ATC_BM);
  ap.q_lo = make_map(qkv.lo, qkv.C, qkv.rows, ATC_BM);
  ap.kv_hi = make_map(qkv.hi, qkv.C, qkv.rows, ATC_KT);
  ap.kv_lo = make_map(qkv.lo, qkv.C, qkv.rows, ATC_KT);
  ap.rk_hi = make_map(L.rk_hi, 128, ATC_RELP, ATC_RELP);
  ap.rk_lo = make_map(L.rk_lo, 128, ATC_RELP, ATC_RELP);
  ap.rv_hi = make_map(L.rv_hi, 128, ATC_RELP, ATC_RELP);
  ap.rv_lo = make_map(L.rv_lo, 128, ATC_RELP, ATC_RELP);
  ap.out = ao; ap.ldo = Hc;
  ap.p_hi = pl ? pl->hi : nullptr; ap.p_lo = pl ? pl->lo : nullptr; ap.ldp = pl ? pl->C : 0;
  ap.n_heads = L.heads; ap.window = cfg.window_size;
  ap.koff = Hc; ap.voff = 2 * Hc;
  dim3 grid((maxLen + ATC_BM - 1) / ATC_BM, L.heads, B);
  if (grid.x == 0) return;
  switch (dk / 32) {
    case 1: klaunch(attn_tc_kernel<32>, grid, dim3(ATC_THREADS), (size_t)atc_smem_bytes(32), ap, lens, offs); break;
    case 2: klaunch(attn_tc_kernel<64>, grid, dim3(ATC_THREADS), (size_t)atc_smem_bytes(64), ap, lens, offs); break;
    case 3: klaunch(attn_tc_kernel<96>, grid, dim3(ATC_THREADS), (size_t)atc_smem_bytes(96), ap, lens, offs); break;
    default: klaunch(attn_tc_kernel<128>, grid, dim3(ATC_THREADS), (size_t)atc_smem_bytes(128), ap, lens, offs); break;
  }
  CK(cudaGetLastError());
  ++launches;
}

void vtts_engine::launch_attn(const float* qkv, float* ao, const EncLayerW& L, int Hc, const int* lens, const int* offs, int maxLen, Planes* pl) {
  const int n_heads = L.heads;
  const int dk = Hc / n_heads, nrel = 2 * cfg.window_size + 1;
  __nv_bfloat16* ph = pl ? pl->hi : nullptr;
  __nv_bfloat16* plo = pl ? pl->lo : nullptr;
  __nv_bfloat16* pmi = pl ? pl->mid : nullptr;
  // register-blocked variant (4 query rows per warp) once the launch is throughput bound
  const std::vector<int>& hl = (lens == d_tok_len.p) ? v_tok_len : v_frm_len;
  long rows = 0;
  for (int b = 0; b < B; ++b) rows += hl[b];
  const int R = (attn_rows == 1 || attn_rows == 4) ? attn_rows : (rows * n_heads >= 8L * 2 * 148 * 4 ? 4 : 1);
  // single short utterances: split-KV variant (all K/V tiles resident, 4 warps per query row) when it fits one wave
  {
    long ctas = 0;
    for (int b = 0; b < B; ++b) ctas += (long)((hl[b] + ATS_ROWS - 1) / ATS_ROWS) * n_heads;
    const int mt = (maxLen + AT_KT - 1) / AT_KT;
    if (attn_split && R == 1 && mt <= ATS_MAXT && ctas <= 148 && dk % 32 == 0 && dk <= 128) {
      dim3 grid((maxLen + ATS_ROWS - 1) / ATS_ROWS, n_heads, B);
      const size_t smem = (size_t)attn_split_smem_floats(dk, nrel, mt) * sizeof(float);
#define ATTN_SPLIT(D) klaunch(attn_split_kernel<D>, grid, dim3(ATS_THREADS), smem, qkv, 3 * Hc, ao, Hc, L.relk, L.relv, n_heads, cfg.window_size, mt, lens, offs, ph, plo, pmi)
      switch (dk / 32) { case 1: ATTN_SPLIT(1); break; case 2: ATTN_SPLIT(2); break; case 3: ATTN_SPLIT(3); break; default: ATTN_SPLIT(4); break; }
#undef ATTN_SPLIT
      CK(cudaGetLastError());
      ++launches;
      return;
    }
  }
  const int QT = 8 * R;
  dim3 grid((maxLen + QT - 1) / QT, n_heads, B);
  const size_t smem = (size_t)attn_smem_floats(dk, nrel, R) * sizeof(float);
#define ATTN_CASE(D, RR) klaunch(attn_kernel<D, RR>, grid, dim3(AT_THREADS), smem, qkv, 3 * Hc, ao, Hc, L.relk, L.relv, n_heads, cfg.window_size, lens, offs, ph, plo, pmi)
  if (R == 4) {
    switch (dk / 32) { case 1: ATTN_CASE(1, 4); break; case 2: ATTN_CASE(2, 4); break; case 3: ATTN_CASE(3, 4); break; default: ATTN_CASE(4, 4); break; }
  } else {
    switch (dk / 32) { case 1: ATTN_CASE(1, 1); break; case 2: ATTN_CASE(2, 1); break; case 3: ATTN_CASE(3, 1); break; default: ATTN_CASE(4, 1); break; }
  }
#undef ATTN_CASE
  CK(cudaGetLastError());
  ++launches;
}

// Flow (reverse, models.py:750-757) with every dense conv on the tensor cores.  Producers emit the split-bf16
// planes their consumer needs: FFMA pre-conv, attention and LayerNorm kernels through an extra epilogue output,
// tensor-core convs through theirs.  The Flip folding is the same as in the fp32 path.
void vtts_engine::alloc_flow_planes() {
  const int H = cfg.hidden_channels;
  const long F = Tfrm;
  int slot = 40;    // plane slots 40.. are the flow's (the decoder uses 0..)
  flp.ph = planes(slot++, F, 1, H); flp.pao = planes(slot++, F, 1, H); flp.ph1 = planes(slot++, F, 1, H);
  flp.pff = planes(slot++, F, 1, H); flp.pwx = planes(slot++, F, 1, H); flp.pacts = planes(slot++, F, 1, H);
  flp.pskip = planes(slot++, F, 1, H);
  flp.pqkv = planes(slot++, F, 1, 3 * H);
  flp.pwx2 = planes(slot++, F, 1, H);
}

// One WaveNet layer as a single cluster kernel (wn_tc.cuh): reads the planes `xin`, writes the updated hidden state to
// x (fp32, in place) and to the OTHER plane set `xout` (neighbouring row tiles still read `xin` for their conv halos).
void vtts_engine::launch_wn_fused(const FlowW& W, int f, int i, const Planes& xin, const Planes& xout, float* x, float* skip,
                                  const Planes& pskip, int dil, const int* fl, const int* fo) {
  const vtts_config& c = cfg;
  const int H = c.hidden_channels, nl = c.flow_wn_layers, fk = c.flow_kernel_size;
  const bool last = (i == nl - 1);
  const int BN1 = 2 * H / WN_NCL, BN2 = (last ? H : 2 * H) / WN_NCL;
  WnParams wp;
  memset(&wp, 0, sizeof(wp));
  wp.a_hi = make_map(xin.hi, H, xin.rows, 128);
  wp.a_lo = make_map(xin.lo, H, xin.rows, 128);
  wp.win_hi = make_map(W.t_in[i].hi, H, (long)fk * 2 * H, BN1);
  wp.win_lo = make_map(W.t_in[i].lo, H, (long)fk * 2 * H, BN1);
  if (!last) {
    wp.wrx_hi = make_map(W.t_rsx[i].hi, H, H, BN2);
    wp.wrx_lo = make_map(W.t_rsx[i].lo, H, H, BN2);
    wp.bias_rx = W.rsx[i].b;
  }
  wp.wrs_hi = make_map(W.t_rss[i].hi, H, H, BN2);
  wp.wrs_lo = make_map(W.t_rss[i].lo, H, H, BN2);
  wp.bias_rs = W.rss[i].b;
  wp.bias_in = W.in[i].b;
  if (has_g) { wp.cond = d_condv.p + r_flow + (f * nl + i) * 2 * H; wp.cond_ld = condR; }
  wp.x = x; wp.xp_hi = xout.hi; wp.xp_lo = xout.lo;
  wp.skip = skip;
  if (last) { wp.sp_hi = pskip.hi; wp.sp_lo = pskip.lo; }
  wp.k = fk; wp.dil = dil; wp.pad = dil * (fk - 1) / 2;
  wp.first = (i == 0) ? 1 : 0;
  dim3 grid((maxFrm + 127) / 128, WN_NCL, B);
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = grid; lc.blockDim = dim3(WN_THREADS); lc.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  at[na].id = cudaLaunchAttributeClusterDimension;
  at[na].val.clusterDim.x = 1; at[na].val.clusterDim.y = WN_NCL; at[na].val.clusterDim.z = 1;
  ++na;
  if (use_pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  lc.attrs = at; lc.numAttrs = na;
#define WN_LAUNCH(HH)                                                                                                   \
  do {                                                                                                                  \
    lc.dynamicSmemBytes = wn_smem_bytes<HH>();                                                                          \
    if (last) CK(cudaLaunchKernelEx(&lc, wn_layer_tc_kernel<HH, true>, wp, fl, fo));                                    \
    else CK(cudaLaunchKernelEx(&lc, wn_layer_tc_kernel<HH, false>, wp, fl, fo));                                        \
  } while (0)
  if (H == 192) WN_LAUNCH(192); else if (H == 128) WN_LAUNCH(128); else WN_LAUNCH(64);
#undef WN_LAUNCH
  CK(cudaGetLastError());
  if (profiling) {   // counted with the tcgen05 conv family (algorithmic FLOPs of both GEMMs)
    for (int b = 0; b < B; ++b) tc_prof_flops += 2.0 * h_frm_len[b] * ((double)2 * H * H * fk + (double)(last ? H : 2 * H) * H);
  }
  ++launches;
}
void vtts_engine::alloc_decoder_planes() {
  const vtts_config& c = cfg;
  const long F = Tfrm;
  const int nk = c.n_resblock_kernels;
  int slot = 0;
  dcp.pz = planes(slot++, F, 1, c.inter_channels);
  dcp.cur = planes(slot++, F, 1, c.upsample_initial_channel);
  dcp.px.resize(c.n_upsamples); dcp.nxt.resize(c.n_upsamples); dcp.pj.resize(c.n_upsamples); dcp.pt.resize(c.n_upsamples);
  int rmp = 1, chp = c.upsample_initial_channel;
  for (int i = 0; i < c.n_upsamples; ++i) {
    rmp *= c.upsample_rates[i];
    chp /= 2;
    dcp.px[i] = planes(slot++, F, rmp, chp);
    dcp.pj[i].resize(nk); dcp.pt[i].resize(nk);
    for (int j = 0; j < nk; ++j) { dcp.pj[i][j] = planes(slot++, F, rmp, chp); dcp.pt[i][j] = planes(slot++, F, rmp, chp); }
    dcp.nxt[i] = planes(slot++, F, rmp, chp, (i + 1 == c.n_upsamples) ? 1 : 0);
  }
}

// emit_pz: the post convs of the last two coupling layers also write the split-bf16 planes of their half of z for the
// decoder's conv_pre (dcp.pz), which saves the separate fp32 -> planes pass
void vtts_engine::flow_tc(float* z, const int* fl, const int* fo, bool emit_pz) {
  const vtts_config& c = cfg;
  const int H = c.hidden_channels, I = c.inter_channels, half = I / 2;
  const long F = Tfrm;
  const int nf = c.flow_n_flows, nl = c.flow_wn_layers, fk = c.flow_kernel_size;
  float* h = ensure(d_h, (size_t)F * H);
  float* h1 = ensure(d_h1, (size_t)F * H);
  float* wx = ensure(d_wx, (size_t)F * H);
  float* skip = ensure(d_skip, (size_t)F * H);
  float* fy = ensure(d_fy, (size_t)F * H);
  float* fqkv = ensure(d_fqkv, (size_t)F * 3 * H);
  float* fao = ensure(d_fao, (size_t)F * H);
  Planes ph = flp.ph, pao = flp.pao, ph1 = flp.ph1, pff = flp.pff, pwx = flp.pwx, pacts = flp.pacts, pskip = flp.pskip, pqkv = flp.pqkv;
  const bool fused = wn_fused_ok();
  dim3 lg((maxFrm + 3) / 4, B);
  for (int f = nf - 1; f >= 0; --f) {
    const FlowW& W = flow[f];
    const bool flipped = ((nf - f) % 2) == 1;
    const int x0off = flipped ? half : 0, x1off = flipped ? 0 : half;
    {
      ConvP p = mk(W.pre, z, I, x0off, h, H, 0, 1, 0);
      Planes& dst = c.use_transformer_flows ? ph : pwx;
      p.p_hi = dst.hi; p.p_lo = dst.lo; p.ldp = H; p.pl_slope = 1.f;
      launch_conv({p}, 1, fl, fo, maxFrm, B);
    }
    float* wn_x = h;
    if (c.use_transformer_flows) {
      if (attn_use_tc(W.tr, H, fl, maxFrm)) {
        // q, k, v leave the qkv conv as split-bf16 planes only; attention runs on tcgen05 (attn_tc.cuh)
        { TcSpec q; q.in = ph; q.w = W.t_qkv; q.bias = W.tr.qkv.b; q.Cin = H; q.Cout = 3 * H; q.out = pqkv; q.pl_slope = 1.f;
          launch_tc({q}, 1, fl, fo, maxFrm, B); }
        launch_attn_tc(pqkv, nullptr, &pao, W.tr, H, fl, fo, maxFrm);
      } else {
        { TcSpec q; q.in = ph; q.w = W.t_qkv; q.bias = W.tr.qkv.b; q.Cin = H; q.Cout = 3 * H; q.y = fqkv; q.ldy = 3 * H;
          launch_tc({q}, 1, fl, fo, maxFrm, B); }
        launch_attn(fqkv, fao, W.tr, H, fl, fo, maxFrm, &pao);
      }
      { TcSpec q; q.in = pao; q.w = W.t_o; q.bias = W.tr.o.b; q.Cin = H; q.Cout = H; q.y = fy; q.ldy = H;
        launch_tc({q}, 1, fl, fo, maxFrm, B); }
      klaunch(add_ln_kernel, dim3(lg), dim3(128), (size_t)(0), h, fy, W.tr.ln1.g, W.tr.ln1.b, nullptr, nullptr, 0, h1, fl, fo, H, ph1.hi, ph1.lo, (__nv_bfloat16*)nullptr);
      CK(cudaGetLastError());
      ++launches;
      { TcSpec q; q.in = ph1; q.w = W.t_ffn1; q.bias = W.tr.ffn1.b; q.Cin = H; q.Cout = H; q.k = fk; q.pad = (fk - 1) / 2;
        q.epi = TCE_RELU; q.out = pff; q.pl_slope = 1.f;
        launch_tc({q}, 1, fl, fo, maxFrm, B); }
      { TcSpec q; q.in = pff; q.w = W.t_ffn2; q.bias = W.tr.ffn2.b; q.Cin = H; q.Cout = H; q.k = fk; q.pad = (fk - 1) / 2;
        q.y = fy; q.ldy = H;
        launch_tc({q}, 1, fl, fo, maxFrm, B); }
      klaunch(add_ln_kernel, dim3(lg), dim3(128), (size_t)(0), h1, fy, W.tr.ln2.g, W.tr.ln2.b, h, nullptr, 0, wx, fl, fo, H, pwx.hi, pwx.lo, (__nv_bfloat16*)nullptr);
      CK(cudaGetLastError());
      ++launches;
      wn_x = wx;
    }
    int dil = 1;
    for (int i = 0; i < nl && fused; ++i) {
      // layer i reads the hidden state's planes from one buffer and writes the updated ones to the other
      launch_wn_fused(W, f, i, (i % 2 == 0) ? pwx : flp.pwx2, (i % 2 == 0) ? flp.pwx2 : pwx, wn_x, skip, pskip, dil, fl, fo);
      dil *= c.flow_dilation_rate;
    }
    for (int i = 0; i < nl && !fused; ++i) {
      { TcSpec q; q.in = pwx; q.w = W.t_in[i]; q.bias = W.in[i].b; q.Cin = H; q.Cout = 2 * H; q.k = fk; q.dil = dil;
        q.pad = dil * (fk - 1) / 2; q.epi = TCE_GATE; q.out = pacts; q.pl_slope = 1.f;
        if (has_g) { q.cond = d_condv.p + r_flow + (f * nl + i) * 2 * H; q.cond_ld = condR; }
        launch_tc({q}, 1, fl, fo, maxFrm, B); }
      TcSpec qs; qs.in = pacts; qs.w = W.t_rss[i]; qs.bias = W.rss[i].b; qs.Cin = H; qs.Cout = H; qs.y = skip; qs.ldy = H;
      if (i > 0) { qs.res = skip; qs.ldr = H; }
      if (i < nl - 1) {
        TcSpec qx; qx.in = pacts; qx.w = W.t_rsx[i]; qx.bias = W.rsx[i].b; qx.Cin = H; qx.Cout = H; qx.y = wn_x; qx.ldy = H;
        qx.res = wn_x; qx.ldr = H; qx.out = pwx; qx.pl_slope = 1.f;
        launch_tc({qx, qs}, 1, fl, fo, maxFrm, B);
      } else {
        qs.out = pskip; qs.pl_slope = 1.f;
        launch_tc({qs}, 1, fl, fo, maxFrm, B);
      }
      dil *= c.flow_dilation_rate;
    }
    { TcSpec q; q.in = pskip; q.w = W.t_post; q.bias = W.post.b; q.Cin = H; q.Cout = half; q.alpha = -1.f;
      q.y = z; q.ldy = I; q.yoff = x1off; q.res = z; q.ldr = I; q.roff = x1off;
      if (emit_pz && f <= 1) { q.out = dcp.pz; q.poff = x1off; q.pl_slope = 1.f; }     // this half of z is final now
      launch_tc({q}, 1, fl, fo, maxFrm, B); }
  }
}

// Decoder on the tensor cores (models.py:1016-1040): every conv consumes the split-bf16 planes written by its
// producer's epilogue; fp32 copies exist only where a residual or the MRF mean needs them.
void vtts_engine::decoder_tc(float* z, const int* fl, const int* fo, bool pz_ready) {
  const vtts_config& c = cfg;
  const int I = c.inter_channels;
  const long F = Tfrm;
  const int nk = c.n_resblock_kernels, nd = c.n_resblock_dilations;
  Planes pz = dcp.pz, cur = dcp.cur;
  const std::vector<Planes>&st_px = dcp.px, &st_nxt = dcp.nxt;
  const std::vector<std::vector<Planes>>&st_pj = dcp.pj, &st_pt = dcp.pt;
  if (!pz_ready) {
    dim3 g((maxFrm + EW_ROWS - 1) / EW_ROWS, B);
    klaunch(split_planes_kernel, dim3(g), dim3(EW_THREADS), (size_t)(0), z, I, pz.hi, pz.lo, I, I, 1.f, 0, 1, fl, fo);
    CK(cudaGetLastError());
    ++launches;
  }
  int ch = c.upsample_initial_channel;
  {
    TcSpec q;
    q.in = pz; q.w = tc_pre; q.bias = dec_pre.b; q.Cin = I; q.Cout = ch; q.k = 7; q.dil = 1; q.pad = 3;
    q.out = cur; q.pl_slope = 0.1f;
    if (debug_flags & 1) { q.y = ensure(d_d0, (size_t)F * ch); q.ldy = ch; }
    launch_tc({q}, 1, fl, fo, maxFrm, B);
  }
  int rm = 1;
  if ((int)d_stage.size() < c.n_upsamples) {
    d_stage.resize(c.n_upsamples);
    d_xj.resize(c.n_upsamples);
    d_tmp.resize(c.n_upsamples);
    for (int i = 0; i < c.n_upsamples; ++i) { d_xj[i].resize(nk); d_tmp[i].resize(nk); }
  }
  float* lastX = nullptr;
  for (int i = 0; i < c.n_upsamples; ++i) {
    const int u = c.upsample_rates[i], ch2 = ch / 2;
    const long rows = F * rm * u;
    float* X = ensure(d_stage[i], (size_t)rows * ch2);
    Planes px = st_px[i];
    for (int r0 = 0; r0 < u; r0 += TC_MAXP) {
      std::vector<TcSpec> ps;
      for (int r = r0; r < std::min(u, r0 + TC_MAXP); ++r) {
        TcSpec q;
        q.in = cur; q.w = ups[i].tphase[r]; q.bias = ups[i].phase[r].b; q.Cin = ch; q.Cout = ch2;
        q.k = ups[i].phase[r].k; q.dil = 1; q.pad = ups[i].pad[r];
        q.y = X; q.ldy = ch2; q.out = px; q.pl_slope = 0.1f; q.out_mul = u; q.out_add = r;
        ps.push_back(q);
      }
      launch_tc(ps, rm, fl, fo, maxFrm, B);
    }
    rm *= u;
    ch = ch2;
    std::vector<float*> xj(nk);
    std::vector<Planes> pj(nk), pt(nk);
    for (int j = 0; j < nk; ++j) {
      xj[j] = ensure(d_xj[i][j], (size_t)rows * ch);
      pj[j] = st_pj[i][j];
      pt[j] = st_pt[i][j];
    }
    auto rb_pair = [&](int j, int d, TcSpec& a, TcSpec& b2) {
      const RbW& R = rbs[i * nk + j];
      const int k = c.resblock_kernel_sizes[j], dl = c.resblock_dilations[j][d];
      a.in = (d == 0) ? px : pj[j]; a.w = R.t1[d]; a.bias = R.c1[d].b; a.Cin = ch; a.Cout = ch; a.k = k; a.dil = dl;
      a.pad = dl * (k - 1) / 2; a.out = pt[j]; a.pl_slope = 0.1f;
      b2.in = pt[j]; b2.w = R.t2[d]; b2.bias = R.c2[d].b; b2.Cin = ch; b2.Cout = ch; b2.k = k; b2.dil = 1; b2.pad = (k - 1) / 2;
      b2.res = (d == 0) ? X : xj[j]; b2.ldr = ch; b2.y = xj[j]; b2.ldy = ch;
      if (d + 1 < nd) { b2.out = pj[j]; b2.pl_slope = 0.1f; }
    };
    // The nk resblocks of the MRF (models.py:1030-1036) are independent chains of 2*nd convs.  Grouped launches keep them
    // in lock-step, so every step lasts as long as its largest kernel size.  Experiment (VTTS_MRF_BRANCH=1, off): for
    // single utterances each chain runs on its own stream (fork / join with events, also inside the captured graph) with
    // the split-K width that suits its own k-loop -- correct, but the concurrent cluster launches of three streams
    // contend and the step gets slower (1.74 vs 1.62 ms).
    long group_tiles = 0;
    for (int b = 0; b < B; ++b) group_tiles += (long)nk * ((v_frm_len[b] * rm + TC_BM - 1) / TC_BM) * ((ch + 63) / 64);
    const bool branch = mrf_branch && !profiling && nk > 1 && nk - 1 <= 3 && group_tiles <= 148;
    if (branch) {
      struct Restore { cudaStream_t& s; cudaStream_t v; ~Restore() { s = v; } } restore{stream, stream};
      cudaStream_t main_stream = stream;
      CK(cudaEventRecord(ev_fork, main_stream));
      for (int j = nk - 1; j >= 0; --j) {              // largest kernel size first
        if (j > 0) {
          CK(cudaStreamWaitEvent(side[j - 1], ev_fork, 0));
          stream = side[j - 1];
        } else {
          stream = main_stream;
        }
        for (int d = 0; d < nd; ++d) {
          TcSpec a, b2;
          rb_pair(j, d, a, b2);
          launch_tc({a}, rm, fl, fo, maxFrm, B);
          launch_tc({b2}, rm, fl, fo, maxFrm, B);
        }
        if (j > 0) CK(cudaEventRecord(ev_join[j - 1], stream));
      }
      stream = main_stream;
      for (int j = 1; j < nk; ++j) CK(cudaStreamWaitEvent(main_stream, ev_join[j - 1], 0));
    } else {
      for (int d = 0; d < nd; ++d) {
        std::vector<TcSpec> p1(nk), p2(nk);
        // CTAs are dispatched in blockIdx.z order = problem order: the resblock with the longest k-loop (largest kernel
        // size) goes first, so that a second wave holds the short ones
        for (int j = 0; j < nk; ++j) rb_pair(j, d, p1[mrf_heavy_first ? nk - 1 - j : j], p2[mrf_heavy_first ? nk - 1 - j : j]);
        launch_tc(p1, rm, fl, fo, maxFrm, B);
        launch_tc(p2, rm, fl, fo, maxFrm, B);
      }
    }
    const bool last = (i + 1 == c.n_upsamples);
    Planes nxt = st_nxt[i];
    {
      dim3 g((maxFrm * rm + (last ? 1 : 0) + EW_ROWS - 1) / EW_ROWS, B);
      klaunch(mrf_mean_planes_kernel, dim3(g), dim3(EW_THREADS), (size_t)(0), xj[0], nk > 1 ? xj[1] : nullptr, nk > 2 ? xj[2] : nullptr, std::min(nk, 3),
                                                   (debug_flags & 1) ? X : nullptr, nxt.hi, nxt.lo, ch, last ? 0.01f : 0.1f,
                                                   last ? 1 : 0, rm, fl, fo);
      CK(cudaGetLastError());
      ++launches;
    }
    cur = nxt;
    lastX = X;
  }
  (void)lastX;
  const int cps = c.istft_n_fft + 2, pc = c.subbands * cps;
  float* post = ensure(d_post, ((size_t)F * rm + B) * pc);
  {
    TcSpec q;
    q.in = cur; q.w = tc_post; q.bias = dec_post.b; q.Cin = ch; q.Cout = pc; q.k = 7; q.dil = 1; q.pad = 3;
    q.y = post; q.ldy = pc; q.in_extra = 1; q.out_seq_extra = 1;
    launch_tc({q}, rm, fl, fo, maxFrm, B);
  }
  float* wav = ensure(d_wav, (size_t)F * hop + 16);
  const int M = maxFrm * rm * c.istft_hop;
  dim3 g((M + TL_M - 1) / TL_M, B);
  const size_t smem = ((size_t)tl_rec_frames(63, c.subbands, c.istft_n_fft, c.istft_hop) * pc + (size_t)c.subbands * (TL_M + 2 * tl_halo(63, c.subbands))) * sizeof(float);
  REQUIRE(c.istft_hop == 4 && c.istft_n_fft == 16, VTTS_ERR_INVALID, "iSTFT tail kernel is sized for n_fft=16, hop=4");
  klaunch(istft_pqmf_kernel, dim3(g), dim3(TL_THREADS), (size_t)(smem), post, pc, istft_basis, pqmf, c.subbands, c.istft_n_fft, c.istft_hop, 63, rm, fl, fo, wav, 0, 1);
  CK(cudaGetLastError());
  ++launches;
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
  const std::vector<int>& hl0 = (lens == d_tok_len.p) ? v_tok_len : v_frm_len;
  const std::vector<int>& hl_true = (lens == d_tok_len.p) ? h_tok_len : h_frm_len;
  long base0 = 0;
  for (const ConvP& q : ps)
    for (int b = 0; b < nB; ++b) base0 += (long)((hl0[b] * rmul + q.in_extra + CV_TT - 1) / CV_TT) * ((q.Cout + CV_TC - 1) / CV_TC);
  // many tiles (batched calls): one thread group per CTA (several CTAs per SM; r2, batch 64: 6.98 ms per step against 8.01 / 8.21
  // with 2 / 4 groups), no cluster.  Few tiles (batch 1): the k-steps of a tile are
  // spread over a cluster of S CTAs until the launch fills ~1 wave of SMs; ranks that still have long k-loops then get
  // 2-4 thread groups each (a lone warp per scheduler issues an FFMA only every other cycle).
  int G = base0 >= 2 * 148 ? conv_big_g : conv_min_g;
  for (const ConvP& q : ps)
    while (G > 1 && q.Cin % (CV_CK * G) != 0) G >>= 1;
  auto min_steps = [&](int g) {
    int ms = 1 << 30;
    for (const ConvP& q : ps) ms = (q.Cin % (CV_CK * g) != 0) ? 0 : std::min(ms, q.Cin / (CV_CK * g) * q.k);
    return ms;
  };
  int S = 1;
  while (S < conv_max_s && base0 * S < conv_target && S * 2 <= min_steps(G)) S *= 2;
  const int xw = (CV_TT + maxHalo + 7) / 8 * 8 + 1;
  auto smem_floats = [&](int g) {
    const size_t pipe = (size_t)2 * CV_CK * g * xw + (size_t)CV_NS * CV_CK * g * CV_TC;
    return (std::max(pipe, (size_t)(g - 1) * 32 * CV_THREADS) + 3) / 4 * 4 + (size_t)32 * CV_THREADS;
  };
  if (conv_auto_g && base0 < 2 * 148)
    while (G * 2 <= conv_max_g && min_steps(G * 2) / S >= conv_auto_g && smem_floats(G * 2) * sizeof(float) <= (size_t)CONV_SMEM_MAX) G *= 2;
  REQUIRE(smem_floats(G) * sizeof(float) <= (size_t)CONV_SMEM_MAX, VTTS_ERR_INVALID, "conv tile does not fit in shared memory");
  cb.S = S;
  cb.xw = xw;
  const size_t pipe_floats = (size_t)2 * CV_CK * G * xw + (size_t)CV_NS * CV_CK * G * CV_TC;
  const size_t red_floats = (size_t)(G - 1) * 32 * CV_THREADS;
  const size_t stage_off = (std::max(pipe_floats, red_floats) + 3) / 4 * 4;
  cb.stage_off = (int)stage_off;
  const size_t smem = (stage_off + (S > 1 ? (size_t)32 * CV_THREADS : 0)) * sizeof(float);
  const std::vector<int>& hl = hl_true;      // (profiling FLOP count below)
  dim3 grid(((maxL + CV_TT - 1) / CV_TT) * S, (maxCout + CV_TC - 1) / CV_TC, nB * cb.n);
  if (grid.x == 0) return;
  if (profiling) {
    if (prof_used + 2 > prof_ev.size()) {
      prof_ev.resize(prof_used + 2);
      CK(cudaEventCreate(&prof_ev[prof_used]));
      CK(cudaEventCreate(&prof_ev[prof_used + 1]));
    }
    for (const ConvP& q : ps)
      for (int b = 0; b < nB; ++b)
        prof_flops += 2.0 * ((double)hl[b] * rmul + q.in_extra) * q.Cout * q.Cin * q.k;
    ++prof_launches;
    CK(cudaEventRecord(prof_ev[prof_used], stream));
  }
  {
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = grid;
    lc.blockDim = dim3(CV_THREADS * G);
    lc.dynamicSmemBytes = smem;
    lc.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (S > 1) {
      at[na].id = cudaLaunchAttributeClusterDimension;
      at[na].val.clusterDim.x = S;
      at[na].val.clusterDim.y = 1;
      at[na].val.clusterDim.z = 1;
      ++na;
    }
    if (use_pdl) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    lc.attrs = at;
    lc.numAttrs = na;
    switch (G) {
      case 4: CK(cudaLaunchKernelEx(&lc, conv_kernel<4>, cb, lens, offs)); break;
      case 2: CK(cudaLaunchKernelEx(&lc, conv_kernel<2>, cb, lens, offs)); break;
      default: CK(cudaLaunchKernelEx(&lc, conv_kernel<1>, cb, lens, offs)); break;
    }
  }
  CK(cudaGetLastError());
  if (profiling) {
    CK(cudaEventRecord(prof_ev[prof_used + 1], stream));
    prof_used += 2;
  }
  ++launches;
}

// One relative-attention encoder layer (attentions.py:57-63): x <- LN2(x1 + FFN(x1)), x1 = LN1(x + MHA(x)).
void vtts_engine::encoder_layer(const EncLayerW& L, float*& x, float*& xb, float* qkv, float* ao, float* y, float* ffh, int Hc,
                                int Fc, int ks, const int* lens, const int* offs, int maxLen, const float* vec_after, int vec_ld,
                                const float* cadd_after) {
  const int nB = B;
  launch_conv({mk(L.qkv, x, Hc, 0, qkv, 3 * Hc, 0, 1, 0)}, 1, lens, offs, maxLen, nB);
  launch_attn(qkv, ao, L, Hc, lens, offs, maxLen, nullptr);
  launch_conv({mk(L.o, ao, Hc, 0, y, Hc, 0, 1, 0)}, 1, lens, offs, maxLen, nB);
  dim3 lg((maxLen + 3) / 4, nB);
  klaunch(add_ln_kernel, dim3(lg), dim3(128), (size_t)(0), x, y, L.ln1.g, L.ln1.b, nullptr, nullptr, 0, xb, lens, offs, Hc, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr);
  CK(cudaGetLastError());
  ++launches;
  {
    ConvP p = mk(L.ffn1, xb, Hc, 0, ffh, Fc, 0, 1, (ks - 1) / 2);
    p.epi = EPI_RELU;
    launch_conv({p}, 1, lens, offs, maxLen, nB);
  }
  launch_conv({mk(L.ffn2, ffh, Fc, 0, y, Hc, 0, 1, (ks - 1) / 2)}, 1, lens, offs, maxLen, nB);
  klaunch(add_ln_kernel, dim3(lg), dim3(128), (size_t)(0), xb, y, L.ln2.g, L.ln2.b, cadd_after, vec_after, vec_ld, x, lens, offs, Hc, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr);
  CK(cudaGetLastError());
  ++launches;
}

void vtts_engine::dds_stack(const DdsW* d, int C, int k, float*& a, float*& b, const int* lens, const int* offs, int maxLen,
                            const float* x0, const float* pre_w, const float* pre_b, const float* cond) {
  int dil = 1;
  for (int i = 0; i < 3; ++i) {
    DdsP P;
    P.x = a; P.y = b;
    P.x0 = (i == 0) ? x0 : nullptr; P.pre_w = pre_w; P.pre_b = pre_b; P.cond = cond;
    P.sep_w = d[i].sep_w; P.sep_b = d[i].sep_b;
    P.ln1g = d[i].ln1.g; P.ln1b = d[i].ln1.b;
    P.pw_w = d[i].pw.w; P.pw_b = d[i].pw.b; P.ldw = d[i].pw.ldw;
    P.ln2g = d[i].ln2.g; P.ln2b = d[i].ln2.b;
    P.C = C; P.k = k; P.dil = dil;
    // (16 positions per CTA were tried for batched calls -- 4x less weight streaming per position -- and measured slower:
    //  duration stage 3.20 vs 2.95 ms at batch 64)
    dim3 grid((maxLen + DDS_TT - 1) / DDS_TT, B);
    const size_t smem = ((size_t)DDS_NS * DDS_CH * C + (size_t)C * DDS_TT + 8 * DDS_TT) * sizeof(float);
    klaunch(dds_layer_kernel<DDS_TT>, dim3(grid), dim3(C), (size_t)(smem), P, lens, offs);
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
  if (!capturing) CK(cudaEventRecord(ev[0], stream));
  // ---- inputs (host data was staged into h_pin_in by stage1(); only device work is enqueued here)
  int* tl = ensure(d_tok_len, B);
  int* to = ensure(d_tok_off, B + 1);
  int* ids = ensure(d_ids, T);
  int* sid = ensure(d_sid, B);
  float* prm = ensure(d_prm, 8);
  {
    P1Pin pp = p1_layout(t_max, noise_dp && !noise_on_device);
    CK(cudaMemcpyAsync(tl, pp.len, B * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(to, pp.off, (B + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(prm, pp.prm, 8 * sizeof(float), cudaMemcpyHostToDevice, stream));
    if (ids_packed_host) {
      CK(cudaMemcpyAsync(ids, pp.ids, T * sizeof(int), cudaMemcpyHostToDevice, stream));
      CK(cudaMemcpyAsync(sid, pp.sid, B * sizeof(int), cudaMemcpyHostToDevice, stream));
    } else {
      dim3 g((maxTok + 127) / 128, B);
      klaunch(pack_ids_kernel, dim3(g), dim3(128), (size_t)(0), d_ids64, t_max, ids, tl, to);
      CK(cudaGetLastError());
      klaunch(cast_sid_kernel, dim3((B + 127) / 128), dim3(128), (size_t)(0), d_sid64, sid, B);
      CK(cudaGetLastError());
      launches += 2;
    }
    if (noise_dp && !noise_on_device) {      // staged as [B][2][maxTok] (stage1)
      float* de = ensure(d_eps_dp, (size_t)B * 2 * maxTok);
      CK(cudaMemcpyAsync(de, pp.eps, (size_t)B * 2 * maxTok * sizeof(float), cudaMemcpyHostToDevice, stream));
      noise_dp = de;
    }
  }
  if (!capturing) CK(cudaEventRecord(ev[1], stream));

  if (use_prefetch && n_pref > 0) {
    // encoder / duration-predictor weights first, then flow + decoder (needed ~1 ms later)
    klaunch(l2_prefetch_kernel, dim3(8, 64), dim3(256), (size_t)0, (const PrefRange*)d_pref.p, n_pref, 0, n_pref_phase1);
    klaunch(l2_prefetch_kernel, dim3(8, 64), dim3(256), (size_t)0, (const PrefRange*)d_pref.p, n_pref, n_pref_phase1, n_pref);
    launches += 2;
  }
  // ---- speaker conditioning (models.py:1680-1683)
  float* condv = nullptr;
  if (has_g) {
    condv = ensure(d_condv, (size_t)B * condR);
    dim3 g((condR + 7) / 8, B);
    klaunch(cond_kernel, dim3(g), dim3(256), (size_t)(c.gin_channels * sizeof(float)), emb_g, sid, cond_w, cond_b, condv, c.gin_channels, condR, c.n_speakers);
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
  Planes px, px1, pao, pff, pqkv;
  if (enc_on_tc) {
    begin_planes();
    px = planes(60, (long)T, 1, H, 0, enc_three); px1 = planes(61, (long)T, 1, H, 0, enc_three);
    pao = planes(62, (long)T, 1, H, 0, enc_three); pff = planes(63, (long)T, 1, Fc, 0, enc_three);
    pqkv = planes(59, (long)T, 1, 3 * H);
    flush_tails(tl, to);
  }
  {
    dim3 g(maxTok, B);
    klaunch(embed_kernel, dim3(g), dim3(64), (size_t)(0), ids, enc_emb, x, tl, to, H, sqrtf((float)H), c.n_vocab,
            (spk_vec && c.cond_layer_idx == 0) ? spk_vec : (const float*)nullptr, condR, px.hi, px.lo, px.mid);
    CK(cudaGetLastError());
    ++launches;
  }
  for (int i = 0; i < c.n_layers; ++i) {
    const float* va = (spk_vec && c.cond_layer_idx == i + 1) ? spk_vec : nullptr;
    if (!enc_on_tc) {
      encoder_layer(enc[i], x, xb, qkv, ao, y, ffh, H, Fc, c.kernel_size, tl, to, maxTok, va, condR, nullptr);
      continue;
    }
    // precision mode 2: same layer with the four convs on tcgen05 (attentions.py:57-63)
    const EncLayerW& L = enc[i];
    const int ks = c.kernel_size;
    dim3 lg((maxTok + 3) / 4, B);
    if (attn_use_tc(L, H, tl, maxTok)) {
      { TcSpec q; q.in = px; q.w = L.t_qkv; q.bias = L.qkv.b; q.Cin = H; q.Cout = 3 * H; q.out = pqkv; q.pl_slope = 1.f;
        launch_tc({q}, 1, tl, to, maxTok, B); }
      launch_attn_tc(pqkv, nullptr, &pao, L, H, tl, to, maxTok);
    } else {
      { TcSpec q; q.in = px; q.w = L.t_qkv; q.bias = L.qkv.b; q.Cin = H; q.Cout = 3 * H; q.y = qkv; q.ldy = 3 * H;
        launch_tc({q}, 1, tl, to, maxTok, B); }
      launch_attn(qkv, ao, L, H, tl, to, maxTok, &pao);
    }
    { TcSpec q; q.in = pao; q.w = L.t_o; q.bias = L.o.b; q.Cin = H; q.Cout = H; q.y = y; q.ldy = H;
      launch_tc({q}, 1, tl, to, maxTok, B); }
    klaunch(add_ln_kernel, lg, dim3(128), (size_t)0, x, y, L.ln1.g, L.ln1.b, (const float*)nullptr, (const float*)nullptr, 0, xb, tl, to, H, px1.hi, px1.lo, px1.mid);
    ++launches;
    { TcSpec q; q.in = px1; q.w = L.t_ffn1; q.bias = L.ffn1.b; q.Cin = H; q.Cout = Fc; q.k = ks; q.pad = (ks - 1) / 2;
      q.epi = TCE_RELU; q.out = pff; q.pl_slope = 1.f;
      launch_tc({q}, 1, tl, to, maxTok, B); }
    { TcSpec q; q.in = pff; q.w = L.t_ffn2; q.bias = L.ffn2.b; q.Cin = Fc; q.Cout = H; q.k = ks; q.pad = (ks - 1) / 2;
      q.y = y; q.ldy = H;
      launch_tc({q}, 1, tl, to, maxTok, B); }
    klaunch(add_ln_kernel, lg, dim3(128), (size_t)0, xb, y, L.ln2.g, L.ln2.b, (const float*)nullptr, va, condR, x, tl, to, H, px.hi, px.lo, px.mid);
    ++launches;
  }
  // (the prior projection enc_p.proj, models.py:323, is only needed by phase 2: it is enqueued at the end of this phase so
  //  that it runs while the host picks up the utterance lengths)
  if (!capturing) CK(cudaEventRecord(ev[2], stream));

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
    klaunch(dp_noise_kernel, dim3(g), dim3(128), (size_t)(0), noise_dp, eps_dp_ld, prm, za, zb, tl, to);
    CK(cudaGetLastError());
    ++launches;
  }
  float* cvar = zb;   // conditioning half (x0 after the Flip)
  float* tvar = za;   // transformed half (x1)
  const int nbins = c.dp_num_bins;
  REQUIRE(3 * nbins - 1 <= 32, VTTS_ERR_INVALID, "spline parameter row too wide");
  for (int n = c.dp_n_flows; n >= 2; --n) {
    const CfW& F = cf[n - 2];
    // (the ConvFlow front h = pre(x0) + cond, modules.py:366-367, is computed inside the first DDS layer)
    float *a = dA, *b = dB;
    dds_stack(F.dds, D, c.dp_kernel_size, a, b, tl, to, maxTok, cvar, F.pre_w, F.pre_b, dx);
    launch_conv({mk(F.proj, a, D, 0, h29, 32, 0, 1, 0)}, 1, tl, to, maxTok, B);
    {
      dim3 g((maxTok + 127) / 128, B);
      klaunch(spline_inverse_kernel, dim3(g), dim3(128), (size_t)(0), h29, 32, tvar, nbins, c.dp_tail_bound, sqrtf((float)D), tl, to);
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
  unsigned int* dctr = reinterpret_cast<unsigned int*>(ensure(d_done_ctr, 4));
  int* fl_real = ensure(d_frm_len_real, B);
  klaunch(duration_kernel, dim3(B), dim3(256), (size_t)(0), zlast, dp_ea, 0, 2, prm, wceil, cum, fl, tl, to, fo, B,
          (volatile int*)(use_poll ? d_map : nullptr), dctr, fl_real);
  CK(cudaGetLastError());
  ++launches;
  if (!capturing) CK(cudaEventRecord(ev[3], stream));
  if (!use_poll) {
    int* p_len = reinterpret_cast<int*>(ensure_pinned(h_pin_len, (size_t)(2 * B + 2) * sizeof(int)));
    CK(cudaMemcpyAsync(p_len, fl, B * sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(p_len + B, fo, (B + 1) * sizeof(int), cudaMemcpyDeviceToHost, stream));
  }
  // prior statistics m_p, logs_p (models.py:323-325): overlaps the host's round trip between the two phases
  if (enc_on_tc) {
    TcSpec q; q.in = px; q.w = tc_encproj; q.bias = enc_proj.b; q.Cin = H; q.Cout = 2 * I; q.y = stats; q.ldy = 2 * I;
    launch_tc({q}, 1, tl, to, maxTok, B);
  } else {
    launch_conv({mk(enc_proj, x, H, 0, stats, 2 * I, 0, 1, 0)}, 1, tl, to, maxTok, B);
  }
}

// Host side of phase 1: stage the call's inputs in pinned memory (fixed layout, so a captured graph can re-read it).
vtts_engine::P1Pin vtts_engine::p1_layout(int t_max, bool eps) {
  const size_t T = (size_t)Ttok;
  (void)t_max;
  const size_t bytes = (size_t)(3 * B + 1) * sizeof(int) + T * sizeof(int) + 8 * sizeof(float) +
                       (eps ? (size_t)B * 2 * maxTok * sizeof(float) : 0) + 64;
  char* pin = ensure_pinned(h_pin_in, bytes);
  P1Pin pp;
  pp.len = reinterpret_cast<int*>(pin);
  pp.off = pp.len + B;
  pp.sid = pp.off + B + 1;
  pp.ids = pp.sid + B;
  pp.prm = reinterpret_cast<float*>(pp.ids + T);
  pp.eps = pp.prm + 8;
  return pp;
}

void vtts_engine::stage1(const int* ids_packed_host, const int* sid_host, int t_max, const float* noise_dp_host) {
  P1Pin pp = p1_layout(t_max, noise_dp_host != nullptr);
  memcpy(pp.len, h_tok_len.data(), B * sizeof(int));
  memcpy(pp.off, h_tok_off.data(), (B + 1) * sizeof(int));
  if (ids_packed_host) {
    memcpy(pp.ids, ids_packed_host, (size_t)real_Ttok * sizeof(int));
    memcpy(pp.sid, sid_host, B * sizeof(int));
  }
  pp.prm[0] = scales[0]; pp.prm[1] = scales[1]; pp.prm[2] = scales[2]; pp.prm[3] = 0.f;
  const uint32_t lo = (uint32_t)seed, hi = (uint32_t)(seed >> 32);
  memcpy(&pp.prm[4], &lo, 4);
  memcpy(&pp.prm[5], &hi, 4);
  if (use_poll) {
    const size_t need = (size_t)(2 * B + 4);
    if (need > map_cap) {
      REQUIRE(!capturing, VTTS_ERR_STATE, "mapped buffer growth during capture");
      if (h_map) { CK(cudaStreamSynchronize(stream)); CK(cudaFreeHost(h_map)); h_map = nullptr; }
      map_cap = need + 256;
      CK(cudaHostAlloc(reinterpret_cast<void**>(&h_map), map_cap * sizeof(int), cudaHostAllocMapped));
      CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d_map), h_map, 0));
      h_map[0] = 0;
      ++ws_gen;
    }
    call_seq = (call_seq % 1000000) + 1;
    memcpy(&pp.prm[6], &call_seq, 4);
  } else {
    pp.prm[6] = 0.f;
  }
  memcpy(&pp.prm[7], &spec_cap, 4);        // frames the speculative second phase is sized for (0: none), see duration_kernel
  if (noise_dp_host) {        // [B][2][t_max] -> [B][2][maxTok]: the device layout depends on the length bucket only
    for (int r = 0; r < 2 * B; ++r)
      memcpy(pp.eps + (size_t)r * maxTok, noise_dp_host + (size_t)r * t_max, (size_t)std::min(t_max, maxTok) * sizeof(float));
    eps_dp_ld = maxTok;
  }
}

void vtts_engine::finish1() {
  if (use_poll) {
    // spin on the flag the last phase-1 kernel writes into mapped host memory (bounded; then fall back to a sync)
    volatile int* flag = h_map;
    bool seen = false;
    for (long spin = 0; spin < 40000000L; ++spin) {
      if (*flag == call_seq) { seen = true; break; }
      if ((spin & 0xFFFFF) == 0xFFFFF && cudaStreamQuery(stream) != cudaErrorNotReady) break;   // finished or failed
    }
    if (!seen) {
      CK(cudaStreamSynchronize(stream));
      REQUIRE(*flag == call_seq, VTTS_ERR_CUDA, "phase 1 finished without publishing the utterance lengths");
    }
    h_frm_len.assign(h_map + 1, h_map + 1 + B);
    h_frm_off.assign(h_map + 1 + B, h_map + 1 + 2 * B + 1);
    set_frame_shape();
    have_durations = true;
    return;
  }
  CK(cudaStreamSynchronize(stream));
  const int* p_len = reinterpret_cast<const int*>(h_pin_len.p);
  h_frm_len.assign(p_len, p_len + B);
  h_frm_off.assign(p_len + B, p_len + 2 * B + 1);
  set_frame_shape();
  have_durations = true;
}

// ---------------------------------------------------------------------------------------------------
// Phase 2: alignment + prior sampling, flow^-1, decoder.
// ---------------------------------------------------------------------------------------------------
void vtts_engine::phase2(const float* noise_z, int z_ld, bool noise_on_device, bool run_decoder) {
  const vtts_config& c = cfg;
  const int H = c.hidden_channels, I = c.inter_channels, half = I / 2;
  const size_t F = (size_t)Tfrm;
  const int* tl = d_tok_len.p;
  const int* to = d_tok_off.p;
  const int* fl = d_frm_len.p;
  const int* fo = d_frm_off.p;
  if (!capturing) CK(cudaEventRecord(ev[4], stream));
  if (noise_z && !noise_on_device) {
    // host noise was staged into h_pin_z by the caller (stage_noise_z) as [B][I][maxFrm]: the layout depends on the bucket only
    const size_t n = (size_t)B * I * maxFrm;
    float* de = ensure(d_eps_z, n);
    CK(cudaMemcpyAsync(de, h_pin_z.p, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    noise_z = de;
    z_ld = maxFrm;
  }
  float* z = ensure(d_z, F * I);
  int* ftok = ensure(d_ftok, F);
  {
    dim3 g(maxFrm, B);
    klaunch(sample_prior_kernel, dim3(g), dim3(64), (size_t)(0), d_stats.p, I, d_cum.p, tl, to, fl, fo, noise_z, z_ld, d_prm.p, z, ftok);
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
  const bool flow_on_tc = tc && !flow.empty() && !flow[0].t_in.empty();
  const bool dec_now = tc && run_decoder;
  const bool emit_pz = flow_on_tc && dec_now && nf >= 2 && !(debug_flags & 2);
  if (flow_on_tc || dec_now) {
    begin_planes();
    if (flow_on_tc) alloc_flow_planes();
    if (dec_now) alloc_decoder_planes();
    flush_tails(fl, fo);
  }
  if (flow_on_tc) flow_tc(z, fl, fo, emit_pz);
  for (int f = nf - 1; f >= 0 && !flow_on_tc; --f) {
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
      launch_attn(fqkv, fao, W.tr, H, fl, fo, maxFrm, nullptr);
      launch_conv({mk(W.tr.o, fao, H, 0, fy, H, 0, 1, 0)}, 1, fl, fo, maxFrm, B);
      dim3 lg((maxFrm + 3) / 4, B);
      klaunch(add_ln_kernel, dim3(lg), dim3(128), (size_t)(0), xa, fy, W.tr.ln1.g, W.tr.ln1.b, nullptr, nullptr, 0, xb2, fl, fo, H, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr);
      CK(cudaGetLastError());
      ++launches;
      {
        ConvP p = mk(W.tr.ffn1, xb2, H, 0, ffh2, H, 0, 1, (fk - 1) / 2);
        p.epi = EPI_RELU;
        launch_conv({p}, 1, fl, fo, maxFrm, B);
      }
      launch_conv({mk(W.tr.ffn2, ffh2, H, 0, fy, H, 0, 1, (fk - 1) / 2)}, 1, fl, fo, maxFrm, B);
      klaunch(add_ln_kernel, dim3(lg), dim3(128), (size_t)(0), xb2, fy, W.tr.ln2.g, W.tr.ln2.b, h, nullptr, 0, wx, fl, fo, H, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr, (__nv_bfloat16*)nullptr);
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
  if (!capturing) CK(cudaEventRecord(ev[5], stream));

  if (!run_decoder) return;
  decode(z, fl, fo, /*planes_ready=*/dec_now, /*pz_ready=*/emit_pz);
}

// Decoder over the utterance rows described by (fl, fo) -- the whole batch, or one halo-extended chunk of a single
// utterance (vtts_decode_chunk).
void vtts_engine::decode(float* z, const int* fl, const int* fo, bool planes_ready, bool pz_ready) {
  const vtts_config& c = cfg;
  const int I = c.inter_channels;
  const size_t F = (size_t)Tfrm;
  // ---- decoder (models.py:1016-1054 / 872-891)
  if (tc) {
    if (!planes_ready) {          // (chunked decoding: the decoder runs on its own)
      begin_planes();
      alloc_decoder_planes();
      flush_tails(fl, fo);
    }
    decoder_tc(z, fl, fo, pz_ready);
    if (!capturing) CK(cudaEventRecord(ev[6], stream));
    return;
  }
  int ch = c.upsample_initial_channel;
  float* cur = ensure(d_d0, F * ch);
  {
    ConvP p = mk(dec_pre, z, I, 0, cur, ch, 0, 1, 3);
    if (has_g && r_dec >= 0) { p.cond = d_condv.p + r_dec; p.cond_ld = condR; }
    launch_conv({p}, 1, fl, fo, maxFrm, B);
  }
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
      klaunch(mrf_mean_kernel, dim3((unsigned)((total4 + 255) / 256)), dim3(256), (size_t)(0), xj[0], nk > 1 ? xj[1] : nullptr, nk > 2 ? xj[2] : nullptr,
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
    const size_t smem = ((size_t)tl_rec_frames(63, c.subbands, c.istft_n_fft, c.istft_hop) * pc + (size_t)c.subbands * (TL_M + 2 * tl_halo(63, c.subbands))) * sizeof(float);
    REQUIRE(c.istft_hop == 4 && c.istft_n_fft == 16, VTTS_ERR_INVALID, "iSTFT tail kernel is sized for n_fft=16, hop=4");
    klaunch(istft_pqmf_kernel, dim3(g), dim3(TL_THREADS), (size_t)(smem), post, pc, istft_basis, pqmf, c.subbands, c.istft_n_fft, c.istft_hop, 63, rm, fl, fo, wav, 0, 1);
    CK(cudaGetLastError());
    ++launches;
  } else {
    ConvP p = mk(dec_post, cur, ch, 0, wav, 1, 0, 1, 3);
    p.pro = PRO_LRELU; p.slope = 0.01f;
    p.epi = EPI_TANH;
    launch_conv({p}, rm, fl, fo, maxFrm, B);
  }
  if (!capturing) CK(cudaEventRecord(ev[6], stream));
}

// ===================================================================================================
// C ABI
// ===================================================================================================
namespace {

enum : int { G_ATOMIC = 0, G_BEGIN = 1, G_CONT = 2 };    // one-shot call | opens a two-phase section | continues / closes it

template <typename Fn>
int guarded(vtts_handle h, Fn fn, int mode = G_ATOMIC) {
  if (!h) return VTTS_ERR_INVALID;
  std::unique_lock<std::mutex> lk(h->mu);
  const std::thread::id me = std::this_thread::get_id();
  if (mode == G_CONT) {
    if (!h->two_phase || h->owner != me) {
      h->err = "second phase called without a preceding vtts_durations by the same thread";
      return VTTS_ERR_STATE;
    }
  } else if (h->two_phase && h->owner != me) {
    if (!h->cv.wait_for(lk, std::chrono::seconds(60), [&] { return !h->two_phase; })) {
      h->err = "another thread has held this handle between vtts_durations and vtts_synthesize for 60 s";
      return VTTS_ERR_STATE;
    }
  }
  auto close = [&] { if (h->two_phase) { h->two_phase = false; h->cv.notify_all(); } };
  if (mode != G_CONT) close();          // (the owner itself starting over)
  try {
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) throw Err{VTTS_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)};
    fn();
    if (mode == G_BEGIN) { h->two_phase = true; h->owner = me; }
    if (mode == G_CONT) close();
    return VTTS_OK;
  } catch (const Err& e) {
    h->err = e.msg;
    cudaGetLastError();
    if (mode == G_CONT && e.code != VTTS_ERR_CAPACITY) close();     // a capacity error keeps the durations for a retry
    return e.code;
  } catch (const std::exception& e) {
    h->err = e.what();
    if (mode == G_CONT) close();
    return VTTS_ERR_INVALID;
  }
}

void collect_timings(vtts_handle h) {
  // ev: 0 start, 1 after H2D, 2 after encoder, 3 after dp, 4 phase2 start, 5 after flow, 6 after decoder, 7 after D2H
  float t;
  if (h->last_graphed) {      // stage events are not recorded inside captured graphs
    for (float& v : h->stage_ms) v = 0.f;
    cudaGetLastError();
    return;
  }
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
    off += (int)lengths[b] + (b + 1 < B ? SEQ_GAP : 0);
    mx = std::max(mx, (int)lengths[b]);
  }
  h->h_tok_off[B] = off;
  (void)mx;
  h->set_token_shape();
  h->have_durations = false;
  h->have_latent = false;
}

}  // namespace

namespace {

static bool spec_ok(vtts_handle h, int B);
static void enqueue_phase1_host(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid, int B, int t_max,
                               const float* scales, const float* noise_dp, uint64_t seed, bool may_speculate) {
  setup_lengths(h, lengths, B, t_max);
  memcpy(h->scales, scales, 3 * sizeof(float));
  h->seed = seed;
  h->spec_cap = (may_speculate && spec_ok(h, B)) ? vtts_engine::bucket_frm(h->spec_predict()) : 0;
  std::vector<int> packed(h->Ttok), sid32(B);
  for (int b = 0; b < B; ++b) {
    for (int t = 0; t < h->h_tok_len[b]; ++t) {
      const int64_t id = ids[(size_t)b * t_max + t];
      REQUIRE(id >= 0 && id < h->cfg.n_vocab, VTTS_ERR_INVALID, "phoneme id out of range [0, n_vocab)");
      packed[h->h_tok_off[b] + t] = (int)id;
    }
    REQUIRE(!h->has_g || (sid[b] >= 0 && sid[b] < h->cfg.n_speakers), VTTS_ERR_INVALID, "speaker id out of range [0, n_speakers)");
    sid32[b] = (int)sid[b];
  }
  h->stage1(packed.data(), sid32.data(), t_max, noise_dp);
  h->run_graphed({0x11, B, h->maxTok, h->Ttok, noise_dp ? 1 : 0}, [&] { h->phase1(packed.data(), nullptr, t_max, nullptr, sid32.data(), noise_dp, false); });
}

static void impl_durations(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid, int B, int t_max,
                   const float* scales, const float* noise_dp, uint64_t seed, int64_t* y_lengths, int32_t* durations) {
  enqueue_phase1_host(h, ids, lengths, sid, B, t_max, scales, noise_dp, seed, false);
  h->finish1();
  if (B == 1) h->spec_learn();
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
}

static void impl_synthesize(vtts_handle h, const float* noise_z, int z_ld, float* wav, int64_t wav_ld, int32_t* frame_token, int idx_ld) {
  REQUIRE(h->have_durations, VTTS_ERR_STATE, "vtts_synthesize called without vtts_durations");
  REQUIRE((int64_t)h->real_maxFrm * h->hop <= wav_ld, VTTS_ERR_CAPACITY, "wav_ld is smaller than hop * max(y_lengths)");
  REQUIRE(!noise_z || z_ld >= h->real_maxFrm, VTTS_ERR_CAPACITY, "noise_z has fewer columns than max(y_lengths)");
  REQUIRE(!frame_token || idx_ld >= h->real_maxFrm, VTTS_ERR_CAPACITY, "frame_token has fewer columns than max(y_lengths)");
  if (noise_z) h->stage_noise_z(noise_z, z_ld);
  // graph key = the length BUCKETS (token rows, frame rows), not the lengths: kernels read the true lengths on the device
  h->run_graphed({0x22, h->B, h->maxFrm, h->Tfrm, noise_z ? 1 : 0}, [&] { h->phase2(noise_z, z_ld, false); });
  const size_t nw = (size_t)h->real_Tfrm * h->hop;
  char* pin = h->ensure_pinned((size_t)h->Tfrm * h->hop * sizeof(float) + (size_t)h->Tfrm * sizeof(int) + 64);
  float* pw = reinterpret_cast<float*>(pin);
  int* pi = reinterpret_cast<int*>(pw + (size_t)h->Tfrm * h->hop);
  CK(cudaMemcpyAsync(pw, h->d_wav.p, nw * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (frame_token) CK(cudaMemcpyAsync(pi, h->d_ftok.p, (size_t)h->real_Tfrm * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaEventRecord(h->ev[7], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int b = 0; b < h->B; ++b) {
    memcpy(wav + (size_t)b * wav_ld, pw + (size_t)h->h_frm_off[b] * h->hop, (size_t)h->h_frm_len[b] * h->hop * sizeof(float));
    if (frame_token) memcpy(frame_token + (size_t)b * idx_ld, pi + h->h_frm_off[b], (size_t)h->h_frm_len[b] * sizeof(int));
  }
  collect_timings(h);
  h->have_durations = false;
}

static void enqueue_phase1_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid, int B, int t_max,
                              const float* scales, const float* d_noise_dp, uint64_t seed, bool may_speculate) {
  setup_lengths(h, lengths_host, B, t_max);
  memcpy(h->scales, scales, 3 * sizeof(float));
  h->seed = seed;
  h->spec_cap = (may_speculate && spec_ok(h, B)) ? vtts_engine::bucket_frm(h->spec_predict()) : 0;
  h->stage1(nullptr, nullptr, t_max, nullptr);
  h->eps_dp_ld = t_max;      // device noise is read in the caller's [B][2][t_max] layout
  h->run_graphed({0x33, B, h->maxTok, h->Ttok, t_max, (long long)(uintptr_t)d_ids, (long long)(uintptr_t)d_sid, (long long)(uintptr_t)d_noise_dp},
                 [&] { h->phase1(nullptr, d_ids, t_max, d_sid, nullptr, d_noise_dp, true); });
}

static void impl_durations_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid, int B, int t_max,
                       const float* scales, const float* d_noise_dp, uint64_t seed, int64_t* y_lengths_host) {
  enqueue_phase1_dev(h, d_ids, lengths_host, d_sid, B, t_max, scales, d_noise_dp, seed, false);
  h->finish1();
  if (B == 1) h->spec_learn();
  for (int b = 0; b < B; ++b) y_lengths_host[b] = h->h_frm_len[b];
}

static void impl_synthesize_dev(vtts_handle h, const float* d_noise_z, int z_ld, float* d_wav, int64_t wav_ld);

static bool spec_ok(vtts_handle h, int B) {
  return B == 1 && h->use_spec && h->use_poll && h->spec_ratio > 0.f && !h->profiling && !h->debug_flags;
}

// Both phases of one call through host buffers (vtts_infer).
static void impl_infer(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid, int B, int t_max, const float* scales,
                       const float* noise_dp, const float* noise_z, int z_ld, uint64_t seed, int64_t* y_lengths, float* wav, int64_t wav_ld,
                       int32_t* frame_token, int idx_ld, int* phase) {
  auto now_us = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_0 = now_us();
  enqueue_phase1_host(h, ids, lengths, sid, B, t_max, scales, noise_dp, seed, true);
  const double t_1 = now_us();
  for (double& v : h->host_us) v = 0.0;
  h->host_us[0] = t_1 - t_0;
  if (h->spec_cap > 0) {
    h->assume_frames(h->spec_cap);
    const int cap = h->maxFrm;
    if (noise_z) h->stage_noise_z(noise_z, z_ld);
    h->run_graphed({0x22, h->B, h->maxFrm, h->Tfrm, noise_z ? 1 : 0}, [&] { h->phase2(noise_z, z_ld, false); });
    // the true length is not known on the host yet: the bucket's worth of samples comes back
    const size_t ncap = (size_t)cap * h->hop;
    char* pin = h->ensure_pinned(ncap * sizeof(float) + (size_t)cap * sizeof(int) + 64);
    float* pw = reinterpret_cast<float*>(pin);
    int* pi = reinterpret_cast<int*>(pw + ncap);
    CK(cudaMemcpyAsync(pw, h->d_wav.p, ncap * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (frame_token) CK(cudaMemcpyAsync(pi, h->d_ftok.p, (size_t)cap * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    const double t_2 = now_us();
    CK(cudaStreamSynchronize(h->stream));
    const double t_3 = now_us();
    h->host_us[1] = t_2 - t_1; h->host_us[2] = t_3 - t_2;
    REQUIRE(h->read_published_lengths(), VTTS_ERR_CUDA, "phase 1 finished without publishing the utterance lengths");
    const int real = h->h_frm_len[0];
    h->real_Tfrm = real; h->real_maxFrm = real;
    h->spec_learn();
    y_lengths[0] = real;
    *phase = 1;
    h->host_us[5] = real <= cap ? 1.0 : 2.0;
    if (real <= cap) {
      ++h->spec_hits;
      h->last_graphed = true;
      h->have_durations = true;          // (kept if one of the capacity checks below fails)
      REQUIRE((int64_t)real * h->hop <= wav_ld, VTTS_ERR_CAPACITY, "wav_ld is smaller than hop * max(y_lengths)");
      REQUIRE(!noise_z || z_ld >= real, VTTS_ERR_CAPACITY, "noise_z has fewer columns than max(y_lengths)");
      REQUIRE(!frame_token || idx_ld >= real, VTTS_ERR_CAPACITY, "frame_token has fewer columns than max(y_lengths)");
      memcpy(wav, pw, (size_t)real * h->hop * sizeof(float));
      if (frame_token) memcpy(frame_token, pi, (size_t)real * sizeof(int));
      h->have_durations = false;
      h->host_us[3] = now_us() - t_3; h->host_us[4] = now_us() - t_0;
      return;
    }
    ++h->spec_misses;                    // predicted bucket too small (the device-side lengths were clamped to it): true
    h->set_frame_shape();                //   lengths back, second phase again with the real shape
    h->klaunch(restore_lengths_kernel, dim3(1), dim3(32), (size_t)0, (const int*)h->d_frm_len_real.p, h->d_frm_len.p, h->d_frm_off.p, h->B);
    h->have_durations = true;
  } else {
    h->finish1();
    if (B == 1) h->spec_learn();
    for (int b = 0; b < B; ++b) y_lengths[b] = h->h_frm_len[b];
    *phase = 1;
  }
  impl_synthesize(h, noise_z, z_ld, wav, wav_ld, frame_token, idx_ld);
}

// Both phases through device buffers (vtts_infer_dev).
static void impl_infer_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid, int B, int t_max,
                           const float* scales, const float* d_noise_dp, const float* d_noise_z, int z_ld, uint64_t seed,
                           int64_t* y_lengths_host, float* d_wav, int64_t wav_ld, int* phase) {
  enqueue_phase1_dev(h, d_ids, lengths_host, d_sid, B, t_max, scales, d_noise_dp, seed, true);
  if (h->spec_cap > 0) {
    h->assume_frames(h->spec_cap);
    const int cap = h->maxFrm;
    h->run_graphed({0x44, h->B, h->maxFrm, h->Tfrm, z_ld, (long long)(uintptr_t)d_noise_z}, [&] { h->phase2(d_noise_z, z_ld, true); });
    const size_t ncopy = (size_t)std::min<int64_t>((int64_t)cap * h->hop, wav_ld);
    CK(cudaMemcpyAsync(d_wav, h->d_wav.p, ncopy * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    REQUIRE(h->read_published_lengths(), VTTS_ERR_CUDA, "phase 1 finished without publishing the utterance lengths");
    const int real = h->h_frm_len[0];
    h->real_Tfrm = real; h->real_maxFrm = real;
    h->spec_learn();
    y_lengths_host[0] = real;
    *phase = 1;
    if (real <= cap) {
      ++h->spec_hits;
      h->last_graphed = true;
      h->have_durations = true;
      REQUIRE((int64_t)real * h->hop <= wav_ld, VTTS_ERR_CAPACITY, "wav_ld is smaller than hop * max(y_lengths)");
      REQUIRE(!d_noise_z || z_ld >= real, VTTS_ERR_CAPACITY, "noise_z has fewer columns than max(y_lengths)");
      h->have_durations = false;
      return;
    }
    ++h->spec_misses;
    h->set_frame_shape();
    h->klaunch(restore_lengths_kernel, dim3(1), dim3(32), (size_t)0, (const int*)h->d_frm_len_real.p, h->d_frm_len.p, h->d_frm_off.p, h->B);
    h->have_durations = true;
  } else {
    h->finish1();
    if (B == 1) h->spec_learn();
    for (int b = 0; b < B; ++b) y_lengths_host[b] = h->h_frm_len[b];
    *phase = 1;
  }
  impl_synthesize_dev(h, d_noise_z, z_ld, d_wav, wav_ld);
}

static void impl_synthesize_dev(vtts_handle h, const float* d_noise_z, int z_ld, float* d_wav, int64_t wav_ld) {
  REQUIRE(h->have_durations, VTTS_ERR_STATE, "vtts_synthesize_dev called without vtts_durations_dev");
  REQUIRE((int64_t)h->real_maxFrm * h->hop <= wav_ld, VTTS_ERR_CAPACITY, "wav_ld is smaller than hop * max(y_lengths)");
  REQUIRE(!d_noise_z || z_ld >= h->real_maxFrm, VTTS_ERR_CAPACITY, "noise_z has fewer columns than max(y_lengths)");
  h->run_graphed({0x44, h->B, h->maxFrm, h->Tfrm, z_ld, (long long)(uintptr_t)d_noise_z}, [&] { h->phase2(d_noise_z, z_ld, true); });
  for (int b = 0; b < h->B; ++b)
    CK(cudaMemcpyAsync(d_wav + (size_t)b * wav_ld, h->d_wav.p + (size_t)h->h_frm_off[b] * h->hop,
                       (size_t)h->h_frm_len[b] * h->hop * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaEventRecord(h->ev[7], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  collect_timings(h);
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
    REQUIRE(cfg->precision >= 0 && cfg->precision <= 3, VTTS_ERR_INVALID, "unknown precision mode");
    if (cfg->precision >= 1) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, VTTS_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
      h->encode_tiled = reinterpret_cast<vtts_engine::EncodeFn>(fn);
      CK(cudaFuncSetAttribute(conv_tc_persist_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      CK(cudaFuncSetAttribute(conv_tc_persist_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      CK(cudaFuncSetAttribute(conv_tc_kernel<64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      CK(cudaFuncSetAttribute(conv_tc_kernel<128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      CK(cudaFuncSetAttribute(conv_tc_kernel<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      CK(cudaFuncSetAttribute(conv_tc_kernel<128, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      for (int wi = 0; wi < 2; ++wi)
        for (int si = 0; si < 3; ++si) {
          const int S = 2 << si;
          cudaLaunchConfig_t lc;
          memset(&lc, 0, sizeof(lc));
          lc.gridDim = dim3(1, 1, S * 64); lc.blockDim = dim3(TC_THREADS);
          lc.dynamicSmemBytes = wi ? tc_smem_bytes<128>(16 * 1024) : tc_smem_bytes<64>(16 * 1024);
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = S;
          lc.attrs = at; lc.numAttrs = 1;
          int nc = 0;
          cudaError_t e = wi ? cudaOccupancyMaxActiveClusters(&nc, conv_tc_kernel<128, true, false>, &lc) : cudaOccupancyMaxActiveClusters(&nc, conv_tc_kernel<64, true, false>, &lc);
          if (e != cudaSuccess) { nc = 0; cudaGetLastError(); }
          h->tc_cluster_cap[wi][si] = nc;
        }
      if (getenv("VTTS_VERBOSE"))
        fprintf(stderr, "[vtts] co-resident conv_tc clusters: BN64 %d/%d/%d  BN128 %d/%d/%d (S=2/4/8)\n", h->tc_cluster_cap[0][0], h->tc_cluster_cap[0][1],
                h->tc_cluster_cap[0][2], h->tc_cluster_cap[1][0], h->tc_cluster_cap[1][1], h->tc_cluster_cap[1][2]);
    }
    CK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& st : h->side) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    for (auto& e : h->ev_join) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
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
    if (const char* e = getenv("VTTS_CONV_MAXS")) h->conv_max_s = std::max(1, atoi(e));
    if (const char* e = getenv("VTTS_CONV_TARGET")) h->conv_target = std::max(1, atoi(e));
    if (const char* e = getenv("VTTS_CONV_MAXG")) h->conv_max_g = std::max(1, std::min(4, atoi(e)));
    if (const char* e = getenv("VTTS_CONV_BIGG")) h->conv_big_g = std::max(1, std::min(4, atoi(e)));   // thread groups per CTA on machine-filling FFMA launches
    if (const char* e = getenv("VTTS_TC_TALL")) h->tc_tall = atoi(e);
    if (const char* e = getenv("VTTS_TC_BASEOFF")) h->tc_baseoff = atoi(e);
    if (const char* e = getenv("VTTS_TC_BN")) h->tc_bn = atoi(e);
    if (const char* e = getenv("VTTS_TC_MULTICAST")) h->tc_mc = atoi(e);
    if (const char* e = getenv("VTTS_TC_SPLIT")) h->tc_split = atoi(e);
    if (const char* e = getenv("VTTS_TC_MINSTEPS")) h->tc_min_steps = std::max(1, atoi(e));   // k-steps per CTA below which split-K stops
    if (const char* e = getenv("VTTS_TC_PERSIST")) h->tc_persist = atoi(e);                 // 0: one tile per CTA also on machine-filling launches; 2: persistent grid on every launch without split-K (tests)
    if (const char* e = getenv("VTTS_TC_DBGSKIP")) h->tc_dbgskip = atoi(e);                 // timing experiments only (wrong results)
    if (const char* e = getenv("VTTS_TC_WMC")) h->tc_wmc = atoi(e);                         // 1: weight-tile multicast between CTA pairs of persistent launches (measured neutral, off)
    if (const char* e = getenv("VTTS_TC_COAL")) h->tc_coal = atoi(e);                       // 1: coalesced (transposed) epilogue on launches without split-K
    if (const char* e = getenv("VTTS_TC_PERSIST_MIN")) h->tc_persist_min = std::max(1, atoi(e));   // tiles per SM from which the persistent grid is used
    if (const char* e = getenv("VTTS_MRF_BRANCH")) h->mrf_branch = atoi(e);
    if (const char* e = getenv("VTTS_MRF_HEAVY_FIRST")) h->mrf_heavy_first = atoi(e);
    if (const char* e = getenv("VTTS_WN_FUSED")) h->wn_fused = atoi(e);
    if (const char* e = getenv("VTTS_ATTN_SPLIT")) h->attn_split = atoi(e);       // 0: never use the split-KV attention
    if (const char* e = getenv("VTTS_CONV_AUTOG")) h->conv_auto_g = std::max(0, atoi(e));   // k-steps per rank needed to add thread groups; 0 = never
    if (const char* e = getenv("VTTS_CONV_MING")) h->conv_min_g = std::max(1, std::min(4, atoi(e)));      // 0 auto, 1 off, 2/4/8 cap
    if (const char* e = getenv("VTTS_ATTN_ROWS")) h->attn_rows = atoi(e);
    if (const char* e = getenv("VTTS_ATTN_TC")) h->attn_tc_mode = atoi(e);
    if (const char* e = getenv("VTTS_PDL")) h->use_pdl = atoi(e) != 0;
    if (const char* e = getenv("VTTS_NO_POLL")) h->use_poll = atoi(e) == 0;
    if (const char* e = getenv("VTTS_NO_GRAPHS")) h->use_graphs = atoi(e) == 0;
    if (const char* e = getenv("VTTS_BUCKETS")) h->use_buckets = atoi(e) != 0;
    if (const char* e = getenv("VTTS_SPEC")) h->use_spec = atoi(e) != 0;
    if (const char* e = getenv("VTTS_SPEC_MARGIN")) h->spec_margin = (float)atof(e);      // (< 1 forces mispredictions: tests)
    if (const char* e = getenv("VTTS_CAPTURE_FIRST")) h->capture_on_first = atoi(e) != 0;   // 0: capture a bucket's graph on its second call          // 0: never enqueue phase 2 before the lengths are known    // 0: size everything by the exact lengths
    if (const char* e = getenv("VTTS_PREFETCH")) h->use_prefetch = atoi(e) != 0;
    h->bind_weights();
    h->build_prefetch_list();
    CK(cudaMemsetAsync(h->ensure(h->d_done_ctr, 4), 0, 4 * sizeof(int), h->stream));     // ticket counter of duration_kernel (self-resetting)
    CK(cudaFuncSetAttribute(dds_layer_kernel<DDS_TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    CK(cudaFuncSetAttribute(wn_layer_tc_kernel<192, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wn_smem_bytes<192>()));
    CK(cudaFuncSetAttribute(wn_layer_tc_kernel<192, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wn_smem_bytes<192>()));
    CK(cudaFuncSetAttribute(wn_layer_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wn_smem_bytes<128>()));
    CK(cudaFuncSetAttribute(wn_layer_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wn_smem_bytes<128>()));
    CK(cudaFuncSetAttribute(wn_layer_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wn_smem_bytes<64>()));
    CK(cudaFuncSetAttribute(wn_layer_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wn_smem_bytes<64>()));
    CK(cudaFuncSetAttribute(attn_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc_smem_bytes(32)));
    CK(cudaFuncSetAttribute(attn_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc_smem_bytes(64)));
    CK(cudaFuncSetAttribute(attn_tc_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc_smem_bytes(96)));
    CK(cudaFuncSetAttribute(attn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, atc_smem_bytes(128)));
    CK(cudaFuncSetAttribute(attn_split_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(attn_split_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(attn_split_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(attn_split_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(attn_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_MAX));
    CK(cudaFuncSetAttribute(conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_MAX));
    CK(cudaFuncSetAttribute(conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM_MAX));
    CK(cudaStreamSynchronize(h->stream));
  });
}

void vtts_destroy(vtts_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(h->d_blob);
  Buf<int>* ib[] = {&h->d_ids, &h->d_tok_len, &h->d_tok_off, &h->d_sid, &h->d_wceil, &h->d_cum, &h->d_frm_len, &h->d_frm_off, &h->d_ftok, &h->d_done_ctr, &h->d_frm_len_real};
  for (auto* b : ib) fr(b->p);
  Buf<float>* fb[] = {&h->d_condv, &h->d_x, &h->d_xb, &h->d_qkv, &h->d_ao, &h->d_y, &h->d_ffh, &h->d_stats, &h->d_dA, &h->d_dB, &h->d_dx,
                      &h->d_h29, &h->d_za, &h->d_zb, &h->d_eps_dp, &h->d_z, &h->d_h, &h->d_h1, &h->d_wx, &h->d_acts, &h->d_skip, &h->d_fqkv,
                      &h->d_fao, &h->d_fy, &h->d_ffh2, &h->d_eps_z, &h->d_d0, &h->d_post, &h->d_wav};
  for (auto* b : fb) fr(b->p);
  for (auto& b : h->d_stage) fr(b.p);
  for (auto& v : h->d_xj) for (auto& b : v) fr(b.p);
  for (auto& v : h->d_tmp) for (auto& b : v) fr(b.p);
  for (Buf<char>* hb : {&h->h_pin, &h->h_pin_in, &h->h_pin_len, &h->h_pin_z}) if (hb->p) cudaFreeHost(hb->p);
  for (auto& kv : h->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  fr(h->d_prm.p);
  fr(h->d_pref.p);
  fr(h->d_chunk.p);
  fr(h->d_zp_dbg.p);
  if (h->h_map) cudaFreeHost(h->h_map);
  for (auto& e : h->ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->prof_ev) if (e) cudaEventDestroy(e);
  for (auto& e : h->tc_prof_ev) if (e) cudaEventDestroy(e);
  for (auto& b : h->pl_pool) fr(b.p);
  for (auto& st : h->side) if (st) cudaStreamDestroy(st);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (auto& e : h->ev_join) if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

static thread_local std::string g_free_err;      // last error of the handle-free entry points (vtts_maximum_path*), per thread
const char* vtts_last_error(vtts_handle h) { return h ? h->err.c_str() : g_free_err.c_str(); }

// Monotonic Alignment Search (mas.cuh).  No engine state is involved: the functions run on the current (or given) device.
static int mas_launch(float* d_value, const int* d_ty, const int* d_tx, int B, int Ty, int Tx, int* d_path, cudaStream_t st) {
  if (Tx > MAS_THREADS * MAS_MAXPT) { g_free_err = "vtts_maximum_path: T_x above " + std::to_string(MAS_THREADS * MAS_MAXPT); return VTTS_ERR_INVALID; }
  const size_t smem = (size_t)2 * Tx * sizeof(float);
  cudaError_t e = cudaMemsetAsync(d_path, 0, (size_t)B * Ty * Tx * sizeof(int), st);
  if (e == cudaSuccess) {
    mas_kernel<<<B, MAS_THREADS, smem, st>>>(d_value, d_path, d_ty, d_tx, Ty, Tx);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { g_free_err = std::string("vtts_maximum_path: ") + cudaGetErrorString(e); return VTTS_ERR_CUDA; }
  return VTTS_OK;
}

int vtts_maximum_path_dev(float* d_value, const int32_t* d_t_ys, const int32_t* d_t_xs, int B, int T_y, int T_x, int32_t* d_path, void* stream) {
  if (!d_value || !d_t_ys || !d_t_xs || !d_path || B <= 0 || T_y <= 0 || T_x <= 0) { g_free_err = "vtts_maximum_path_dev: bad argument"; return VTTS_ERR_INVALID; }
  return mas_launch(d_value, d_t_ys, d_t_xs, B, T_y, T_x, d_path, (cudaStream_t)stream);
}

int vtts_maximum_path(const float* neg_cent, const int32_t* t_ys, const int32_t* t_xs, int B, int T_y, int T_x, int32_t* path, int device) {
  if (!neg_cent || !t_ys || !t_xs || !path || B <= 0 || T_y <= 0 || T_x <= 0) { g_free_err = "vtts_maximum_path: bad argument"; return VTTS_ERR_INVALID; }
  for (int b = 0; b < B; ++b)
    if (t_ys[b] < 0 || t_ys[b] > T_y || t_xs[b] < 0 || t_xs[b] > T_x || t_xs[b] > t_ys[b]) {
      g_free_err = "vtts_maximum_path: lengths must satisfy 0 <= t_x <= t_y <= T_y, t_x <= T_x (utterance " + std::to_string(b) + ")";
      return VTTS_ERR_INVALID;
    }
  float* dv = nullptr; int *dp = nullptr, *dl = nullptr;
  const size_t n = (size_t)B * T_y * T_x;
  int rc = VTTS_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&dv, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&dp, n * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&dl, (size_t)2 * B * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(dv, neg_cent, n * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dl, t_ys, (size_t)B * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dl + B, t_xs, (size_t)B * sizeof(int), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) rc = mas_launch(dv, dl, dl + B, B, T_y, T_x, dp, 0);
  if (e == cudaSuccess && rc == VTTS_OK) e = cudaMemcpy(path, dp, n * sizeof(int), cudaMemcpyDeviceToHost);
  if (dv) cudaFree(dv);
  if (dp) cudaFree(dp);
  if (dl) cudaFree(dl);
  if (e != cudaSuccess) { g_free_err = std::string("vtts_maximum_path: ") + cudaGetErrorString(e); return VTTS_ERR_CUDA; }
  return rc;
}

int vtts_durations(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid, int B, int t_max,
                   const float* scales, const float* noise_dp, uint64_t seed, int64_t* y_lengths, int32_t* durations) {
  if (!ids || !lengths || !sid || !scales || !y_lengths) return VTTS_ERR_INVALID;
  return guarded(h, [&] { impl_durations(h, ids, lengths, sid, B, t_max, scales, noise_dp, seed, y_lengths, durations); }, G_BEGIN);
}

int vtts_synthesize(vtts_handle h, const float* noise_z, int z_ld, float* wav, int64_t wav_ld, int32_t* frame_token, int idx_ld) {
  if (!wav) return VTTS_ERR_INVALID;
  return guarded(h, [&] { impl_synthesize(h, noise_z, z_ld, wav, wav_ld, frame_token, idx_ld); }, G_CONT);
}

int vtts_durations_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid, int B, int t_max,
                       const float* scales, const float* d_noise_dp, uint64_t seed, int64_t* y_lengths_host) {
  if (!d_ids || !lengths_host || !d_sid || !scales || !y_lengths_host) return VTTS_ERR_INVALID;
  return guarded(h, [&] { impl_durations_dev(h, d_ids, lengths_host, d_sid, B, t_max, scales, d_noise_dp, seed, y_lengths_host); }, G_BEGIN);
}

int vtts_synthesize_dev(vtts_handle h, const float* d_noise_z, int z_ld, float* d_wav, int64_t wav_ld) {
  if (!d_wav) return VTTS_ERR_INVALID;
  return guarded(h, [&] { impl_synthesize_dev(h, d_noise_z, z_ld, d_wav, wav_ld); }, G_CONT);
}

// ---- streaming: flow once, then decode halo-extended chunks (the decoder's receptive field is
// +-23.9 latent frames, so a chunk decoded with a 24-frame halo on each side reproduces the monolithic output).
int vtts_decoder_halo(vtts_handle h) { return h ? 24 : 0; }

int vtts_flow(vtts_handle h, const float* noise_z, int z_ld) {
  return guarded(h, [&] {
    REQUIRE(h->have_durations, VTTS_ERR_STATE, "vtts_flow called without vtts_durations");
    REQUIRE(!noise_z || z_ld >= h->real_maxFrm, VTTS_ERR_CAPACITY, "noise_z has fewer columns than max(y_lengths)");
    if (noise_z) h->stage_noise_z(noise_z, z_ld);
    h->last_graphed = false;
    h->phase2(noise_z, z_ld, false, /*run_decoder=*/false);
    CK(cudaStreamSynchronize(h->stream));
    h->have_durations = false;
    h->have_latent = true;
  }, G_CONT);
}

int vtts_decode_chunk(vtts_handle h, int f0, int f1, float* wav, int64_t wav_capacity) {
  if (!wav) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    REQUIRE(h->have_latent, VTTS_ERR_STATE, "vtts_decode_chunk called without vtts_flow");
    REQUIRE(h->B == 1, VTTS_ERR_INVALID, "chunked decoding handles one utterance per call");
    const int T = h->h_frm_len[0];
    REQUIRE(0 <= f0 && f0 < f1 && f1 <= T, VTTS_ERR_INVALID, "chunk out of range");
    REQUIRE((int64_t)(f1 - f0) * h->hop <= wav_capacity, VTTS_ERR_CAPACITY, "chunk does not fit the output buffer");
    const int halo = 24;
    const int lo = std::max(0, f0 - halo), hi = std::min(T, f1 + halo);
    int* dc = h->ensure(h->d_chunk, 4);
    int* pin = reinterpret_cast<int*>(h->ensure_pinned(h->h_pin_len, 64));
    pin[0] = hi - lo; pin[1] = lo; pin[2] = hi;
    CK(cudaMemcpyAsync(dc, pin, 3 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    // the launch helpers size grids and split-K from the host copies of the lengths: point them at the chunk
    // (the planes keep the full utterance's row count, Tfrm: the chunk is addressed by its absolute rows)
    const std::vector<int> len0 = h->h_frm_len, off0 = h->h_frm_off, v0 = h->v_frm_len;
    const int max0 = h->maxFrm;
    h->h_frm_len.assign(1, hi - lo);
    h->h_frm_off = {lo, hi};
    h->v_frm_len = h->h_frm_len;
    h->maxFrm = hi - lo;
    try {
      h->last_graphed = false;
      h->decode(h->d_z.p, dc, dc + 1);
    } catch (...) {
      h->h_frm_len = len0; h->h_frm_off = off0; h->maxFrm = max0; h->v_frm_len = v0;
      throw;
    }
    h->h_frm_len = len0; h->h_frm_off = off0; h->maxFrm = max0; h->v_frm_len = v0;
    const size_t n = (size_t)(f1 - f0) * h->hop;
    float* pw = reinterpret_cast<float*>(h->ensure_pinned(n * sizeof(float)));
    CK(cudaMemcpyAsync(pw, h->d_wav.p + (size_t)f0 * h->hop, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(wav, pw, n * sizeof(float));
  });
}

int vtts_infer(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid, int B, int t_max, const float* scales,
               const float* noise_dp, const float* noise_z, int z_ld, uint64_t seed, int64_t* y_lengths, float* wav, int64_t wav_ld,
               int32_t* frame_token, int idx_ld) {
  if (!ids || !lengths || !sid || !scales || !y_lengths || !wav) return VTTS_ERR_INVALID;
  int phase = 0;
  const int rc = guarded(h, [&] {
    impl_infer(h, ids, lengths, sid, B, t_max, scales, noise_dp, noise_z, z_ld, seed, y_lengths, wav, wav_ld, frame_token, idx_ld, &phase);
  });
  if (rc == VTTS_ERR_CAPACITY && phase == 1) {      // durations are kept: this thread may finish with vtts_synthesize
    std::lock_guard<std::mutex> lk(h->mu);
    h->two_phase = true;
    h->owner = std::this_thread::get_id();
  }
  return rc;
}

int vtts_infer_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid, int B, int t_max,
                   const float* scales, const float* d_noise_dp, const float* d_noise_z, int z_ld, uint64_t seed,
                   int64_t* y_lengths_host, float* d_wav, int64_t wav_ld) {
  if (!d_ids || !lengths_host || !d_sid || !scales || !y_lengths_host || !d_wav) return VTTS_ERR_INVALID;
  int phase = 0;
  const int rc = guarded(h, [&] {
    impl_infer_dev(h, d_ids, lengths_host, d_sid, B, t_max, scales, d_noise_dp, d_noise_z, z_ld, seed, y_lengths_host, d_wav, wav_ld, &phase);
  });
  if (rc == VTTS_ERR_CAPACITY && phase == 1) {
    std::lock_guard<std::mutex> lk(h->mu);
    h->two_phase = true;
    h->owner = std::this_thread::get_id();
  }
  return rc;
}

int vtts_hop(vtts_handle h) { return h ? h->hop : 0; }

int vtts_stage_timings(vtts_handle h, float* ms, int n) {
  if (!h || !ms) return VTTS_ERR_INVALID;
  for (int i = 0; i < n && i < 8; ++i) ms[i] = h->stage_ms[i];
  return VTTS_OK;
}

uint64_t vtts_kernel_launches(vtts_handle h) { return h ? h->launches : 0; }
void* vtts_stream(vtts_handle h) { return h ? (void*)h->stream : nullptr; }

int vtts_set_graphs(vtts_handle h, int enable) {
  return guarded(h, [&] { h->use_graphs = enable != 0; });
}

uint64_t vtts_graph_replays(vtts_handle h) { return h ? h->graph_replays : 0; }

int vtts_host_timings(vtts_handle h, double* us, int n) {
  if (!h || !us) return VTTS_ERR_INVALID;
  for (int i = 0; i < n && i < 8; ++i) us[i] = h->host_us[i];
  return VTTS_OK;
}

int vtts_speculation_stats(vtts_handle h, uint64_t* hits, uint64_t* misses) {
  if (!h || !hits || !misses) return VTTS_ERR_INVALID;
  *hits = h->spec_hits;
  *misses = h->spec_misses;
  return VTTS_OK;
}

int vtts_profile(vtts_handle h, int enable) {
  return guarded(h, [&] {
    h->profiling = enable != 0;
    h->prof_used = 0;
    h->prof_flops = 0.0;
    h->prof_launches = 0;
    h->tc_prof_used = 0;
    h->tc_prof_flops = 0.0;
    h->tc_prof_launches = 0;
  });
}

int vtts_profile_read(vtts_handle h, double* conv_ms, uint64_t* conv_launches, double* conv_flops) {
  if (!conv_ms || !conv_launches || !conv_flops) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    CK(cudaStreamSynchronize(h->stream));
    double ms = 0.0;
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
      ms += t;
    }
    *conv_ms = ms;
    *conv_launches = h->prof_launches;
    *conv_flops = h->prof_flops;
  });
}

// Timeline: enable -> every kernel's first CTA appends (source line, globaltimer) to a device buffer; read returns pairs.
int vtts_timeline(vtts_handle h, int enable, unsigned long long* out, size_t max_pairs, size_t* n_out) {
  return guarded(h, [&] {
    static unsigned long long* d_tl = nullptr;
    CK(cudaStreamSynchronize(h->stream));
    if (enable == 1) {
      if (!d_tl) CK(cudaMalloc(&d_tl, (1 + 2 * 4000) * 8));
      CK(cudaMemset(d_tl, 0, (1 + 2 * 4000) * 8));
      CK(cudaMemcpyToSymbol(g_timeline, &d_tl, sizeof(d_tl)));
    } else if (enable == 0) {
      unsigned long long* z = nullptr;
      CK(cudaMemcpyToSymbol(g_timeline, &z, sizeof(z)));
    } else if (d_tl && out && n_out) {
      std::vector<unsigned long long> hbuf(1 + 2 * 4000);
      CK(cudaMemcpy(hbuf.data(), d_tl, hbuf.size() * 8, cudaMemcpyDeviceToHost));
      size_t n = std::min<size_t>((size_t)hbuf[0], std::min<size_t>(4000, max_pairs));
      memcpy(out, hbuf.data() + 1, n * 16);
      *n_out = n;
    }
  });
}

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
    const size_t T = h->real_Ttok, F = h->real_Tfrm;
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

// Unit-test hook: runs ONE attention launch of the named encoder layer ("enc.<i>" or "flow.<f>.tr") on a caller-supplied
// fp32 qkv tensor [T][3H] of a single utterance and returns the fp32 output [T][H].  use_tc = 1: attn_tc_kernel on the
// split-bf16 planes of the input (needs a precision >= 1 engine); 0: the fp32 FFMA kernels.  iters > 0 also times the launch.
int vtts_debug_attention(vtts_handle h, const char* layer, const float* qkv_host, int T, int use_tc, float* out_host, int iters,
                         float* ms_out) {
  if (!layer || !qkv_host || !out_host || T < 1) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    const std::string nm(layer);
    const EncLayerW* L = nullptr;
    if (nm.rfind("enc.", 0) == 0) {
      const int i = atoi(nm.c_str() + 4);
      REQUIRE(i >= 0 && i < (int)h->enc.size(), VTTS_ERR_INVALID, "no such encoder layer");
      L = &h->enc[i];
    } else if (nm.rfind("flow.", 0) == 0) {
      const int f = atoi(nm.c_str() + 5);
      REQUIRE(f >= 0 && f < (int)h->flow.size() && h->cfg.use_transformer_flows, VTTS_ERR_INVALID, "no such flow layer");
      L = &h->flow[f].tr;
    }
    REQUIRE(L != nullptr, VTTS_ERR_INVALID, "layer must be enc.<i> or flow.<f>.tr");
    const int H = h->cfg.hidden_channels;
    REQUIRE(!use_tc || h->attn_tc_ok(*L, H), VTTS_ERR_INVALID, "tensor-core attention is not available for this layer / precision mode");
    const int B0 = h->B; const std::vector<int> fl0 = h->h_frm_len, tl0 = h->h_tok_len, vf0 = h->v_frm_len, vt0 = h->v_tok_len; const int mf0 = h->maxFrm;
    h->B = 1; h->h_frm_len.assign(1, T); h->h_tok_len.assign(1, T); h->maxFrm = T; h->v_frm_len = h->h_frm_len; h->v_tok_len = h->h_tok_len;
    int* dl = nullptr; float *dq = nullptr, *dout = nullptr;
    CK(cudaMalloc(&dl, 16));
    const int hl[3] = {T, 0, T};
    CK(cudaMemcpy(dl, hl, 12, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dq, (size_t)T * 3 * H * 4)); CK(cudaMalloc(&dout, (size_t)T * H * 4));
    CK(cudaMemcpy(dq, qkv_host, (size_t)T * 3 * H * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0, (size_t)T * H * 4));
    Planes pq;
    if (use_tc) {
      h->begin_planes();
      pq = h->planes(58, T, 1, 3 * H);
      h->flush_tails(dl, dl + 1);
      h->klaunch(split_planes_kernel, dim3((T + EW_ROWS - 1) / EW_ROWS, 1), dim3(EW_THREADS), (size_t)0, (const float*)dq, 3 * H, pq.hi, pq.lo, 3 * H, 3 * H, 1.f, 0, 1, (const int*)dl, (const int*)(dl + 1));
    }
    auto once = [&] {
      if (use_tc) h->launch_attn_tc(pq, dout, nullptr, *L, H, dl, dl + 1, T);
      else h->launch_attn(dq, dout, *L, H, dl, dl + 1, T, nullptr);
    };
    once();
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out_host, dout, (size_t)T * H * 4, cudaMemcpyDeviceToHost));
    if (iters > 0 && ms_out) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0, h->stream));
      for (int i = 0; i < iters; ++i) once();
      CK(cudaEventRecord(e1, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      *ms_out = ms / iters;
      cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(dl); cudaFree(dq); cudaFree(dout);
    h->B = B0; h->h_frm_len = fl0; h->h_tok_len = tl0; h->maxFrm = mf0; h->v_frm_len = vf0; h->v_tok_len = vt0;
  });
}

int vtts_profile_read_tc(vtts_handle h, double* ms, uint64_t* launches, double* flops) {
  if (!ms || !launches || !flops) return VTTS_ERR_INVALID;
  return guarded(h, [&] {
    CK(cudaStreamSynchronize(h->stream));
    double t = 0.0;
    for (size_t i = 0; i + 1 < h->tc_prof_used; i += 2) {
      float e = 0.f;
      CK(cudaEventElapsedTime(&e, h->tc_prof_ev[i], h->tc_prof_ev[i + 1]));
      t += e;
    }
    *ms = t;
    *launches = h->tc_prof_launches;
    *flops = h->tc_prof_flops;
  });
}

// Micro-benchmark of one conv kernel in isolation (back-to-back launches, CUDA events on the engine stream):
//   what = "tc:<Cin>:<Cout>:<k>:<dil>:<rows>"   tcgen05 kernel (precision mode 1 engines only)
//          "ffma:<Cin>:<Cout>:<k>:<dil>:<rows>" fp32 FFMA kernel
// Returns the average milliseconds per launch, or a negative status.  Used by bench.py for the roofline of the
// dominant kernel timed alone, and by tools/ for tuning.
float vtts_microbench(vtts_handle h, const char* what, int iters) {
  if (!h || !what || iters < 1) return -1.f;
  float result = -1.f;
  int rc = guarded(h, [&] {
    char kind[16] = {0};
    int Cin = 0, Cout = 0, k = 1, dil = 1, rows = 0;
    REQUIRE(sscanf(what, "%15[a-z]:%d:%d:%d:%d:%d", kind, &Cin, &Cout, &k, &dil, &rows) == 6, VTTS_ERR_INVALID, "bad microbench spec");
    REQUIRE(Cin > 0 && Cout > 0 && k > 0 && rows > 0 && Cin % 64 == 0, VTTS_ERR_INVALID, "bad microbench shape");
    const bool is_tc = strcmp(kind, "tc") == 0;
    REQUIRE(is_tc || strcmp(kind, "ffma") == 0, VTTS_ERR_INVALID, "unknown microbench kind");
    REQUIRE(!is_tc || h->tc, VTTS_ERR_INVALID, "tc microbench needs a precision-1 engine");
    // save state that the launch helpers read
    const int B0 = h->B; const std::vector<int> fl0 = h->h_frm_len, tl0 = h->h_tok_len, vf0 = h->v_frm_len, vt0 = h->v_tok_len;
    h->B = 1; h->h_frm_len.assign(1, rows); h->h_tok_len.assign(1, rows); h->v_frm_len = h->h_frm_len; h->v_tok_len = h->h_tok_len;
    int* dl = nullptr; int* dof = nullptr; float *x = nullptr, *y = nullptr, *w = nullptr, *bias = nullptr;
    __nv_bfloat16 *ph = nullptr, *pl = nullptr, *wh = nullptr, *wl = nullptr;
    const int ldw = (Cout + 3) / 4 * 4;
    CK(cudaMalloc(&dl, 8)); CK(cudaMalloc(&dof, 8));
    const int hl[2] = {rows, rows}, ho[2] = {0, rows};
    CK(cudaMemcpy(dl, hl, 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dof, ho, 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&y, (size_t)rows * Cout * 4)); CK(cudaMalloc(&bias, (size_t)ldw * 4)); CK(cudaMemset(bias, 0, (size_t)ldw * 4));
    if (is_tc) {
      CK(cudaMalloc(&ph, (size_t)rows * Cin * 2)); CK(cudaMalloc(&pl, (size_t)rows * Cin * 2));
      CK(cudaMalloc(&wh, (size_t)k * Cout * Cin * 2)); CK(cudaMalloc(&wl, (size_t)k * Cout * Cin * 2));
      CK(cudaMemset(ph, 0, (size_t)rows * Cin * 2)); CK(cudaMemset(pl, 0, (size_t)rows * Cin * 2));
      CK(cudaMemset(wh, 0, (size_t)k * Cout * Cin * 2)); CK(cudaMemset(wl, 0, (size_t)k * Cout * Cin * 2));
    } else {
      CK(cudaMalloc(&x, (size_t)rows * Cin * 4)); CK(cudaMalloc(&w, (size_t)k * Cin * ldw * 4));
      CK(cudaMemset(x, 0, (size_t)rows * Cin * 4)); CK(cudaMemset(w, 0, (size_t)k * Cin * ldw * 4));
    }
    auto once = [&] {
      if (is_tc) {
        TcSpec q;
        q.in.hi = ph; q.in.lo = pl; q.in.C = Cin; q.in.rows = rows;
        q.w.hi = wh; q.w.lo = wl; q.bias = bias; q.Cin = Cin; q.Cout = Cout; q.k = k; q.dil = dil; q.pad = dil * (k - 1) / 2;
        q.y = y; q.ldy = Cout;
        h->launch_tc({q}, 1, dl, dof, rows, 1);
      } else {
        ConvW W; W.w = w; W.b = bias; W.Cin = Cin; W.Cout = Cout; W.k = k; W.ldw = ldw;
        h->launch_conv({mk(W, x, Cin, 0, y, Cout, 0, dil, dil * (k - 1) / 2)}, 1, dl, dof, rows, 1);
      }
    };
    const bool prof0 = h->profiling; h->profiling = false;
    for (int i = 0; i < 3; ++i) once();
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // the launches are captured into one CUDA graph so that the host launch path is not what gets timed
    cudaGraph_t graph = nullptr; cudaGraphExec_t gexec = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < iters; ++i) once();
    CK(cudaStreamEndCapture(h->stream, &graph));
    CK(cudaGraphInstantiate(&gexec, graph, 0));
    CK(cudaGraphLaunch(gexec, h->stream));           // warm
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventRecord(e0, h->stream));
    CK(cudaGraphLaunch(gexec, h->stream));
    CK(cudaEventRecord(e1, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaGraphExecDestroy(gexec); cudaGraphDestroy(graph);
    if (is_tc && getenv("VTTS_TC_STAMPS")) {
      unsigned long long* d = nullptr;
      CK(cudaMalloc(&d, 16 * 8)); CK(cudaMemset(d, 0, 16 * 8));
      h->tc_dbg = d;
      cudaEvent_t a0, a1; CK(cudaEventCreate(&a0)); CK(cudaEventCreate(&a1));
      CK(cudaEventRecord(a0, h->stream));
      once();
      CK(cudaEventRecord(a1, h->stream));
      CK(cudaStreamSynchronize(h->stream));
      h->tc_dbg = nullptr;
      unsigned long long st[16];
      CK(cudaMemcpy(st, d, sizeof(st), cudaMemcpyDeviceToHost));
      float one = 0.f; CK(cudaEventElapsedTime(&one, a0, a1));
      fprintf(stderr, "[tc stamps %s] event %.2f us | entry->setup %.2f | ->first TMA issued %.2f | ->all TMA issued %.2f | ->first full %.2f | ->mma issued %.2f | ->acc ready %.2f | ->epi done %.2f | ->sync %.2f (us since entry)\n",
              what, one * 1e3, (st[1] - st[0]) / 1e3, (st[2] - st[0]) / 1e3, (st[3] - st[0]) / 1e3, (st[4] - st[0]) / 1e3,
              (st[5] - st[0]) / 1e3, (st[6] - st[0]) / 1e3, (st[7] - st[0]) / 1e3, (st[8] - st[0]) / 1e3);
      cudaFree(d); cudaEventDestroy(a0); cudaEventDestroy(a1);
    }
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    result = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    h->profiling = prof0;
    for (void* p2 : {(void*)dl, (void*)dof, (void*)x, (void*)y, (void*)w, (void*)bias, (void*)ph, (void*)pl, (void*)wh, (void*)wl}) if (p2) cudaFree(p2);
    h->B = B0; h->h_frm_len = fl0; h->h_tok_len = tl0; h->v_frm_len = vf0; h->v_tok_len = vt0;
  });
  return rc == VTTS_OK ? result : (float)rc;
}

}  // extern "C"
