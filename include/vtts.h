/*
 * vtts.h -- C ABI of the B200-native VITS2 inference engine (libvtts.so).
 *
 * Drop-in boundary: the single call the reference makes into its inference runtime,
 *
 *     audio = self.model.onnx.run(None, args)[0]        (vosk_tts/synth.py:123-126)
 *
 * on the session created at vosk_tts/model.py:46.  The graph behind that call is a trace of
 * SynthesizerTrn.infer (training/vits2/models.py:1679-1704, exported by
 * training/vits2/onnx_export.py:47-104) with feeds
 *     input int64[B,T], input_lengths int64[B], scales float32[3], sid int64[B]
 * and output float32[B,1,1,T_wav].  Every entry point below takes plain pointers and sizes
 * (no torch / numpy types); the Python facade in vosk_tts_b200/session.py binds them with
 * ctypes and exposes `run(None, feeds)`.
 *
 * All functions return VTTS_OK (0) or a negative status; the message of the last failure on a
 * handle is available through vtts_last_error().  Nothing throws across the ABI.  Calls on one
 * handle are serialised internally (InferenceSession.run is called concurrently by
 * server/tts_server.py:35,57), different handles are independent.
 */
#ifndef VTTS_H_
#define VTTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VTTS_OK 0
#define VTTS_ERR_INVALID (-1)   /* bad argument / unsupported configuration           */
#define VTTS_ERR_CUDA (-2)      /* CUDA runtime failure (see vtts_last_error)          */
#define VTTS_ERR_WEIGHTS (-3)   /* tensor missing from / mis-sized in the weight blob   */
#define VTTS_ERR_CAPACITY (-4)  /* output buffer too small for the predicted durations   */
#define VTTS_ERR_STATE (-5)     /* vtts_synthesize without a preceding vtts_durations    */

typedef struct vtts_engine* vtts_handle;

/* Model hyper-parameters = the "model" block of the reference training json
 * (training/vits2/configs/mb_istft_vits2_multi.json:42-79) + the constants hard-coded in
 * SynthesizerTrn.__init__ (training/vits2/models.py:1613-1625). */
typedef struct vtts_config {
  int32_t n_vocab, n_speakers, gin_channels;
  int32_t inter_channels, hidden_channels, filter_channels;
  int32_t n_heads, n_layers, kernel_size, window_size;
  int32_t spk_cond_encoder, cond_layer_idx;
  int32_t use_transformer_flows;
  int32_t flow_kernel_size, flow_dilation_rate, flow_wn_layers, flow_n_flows;
  int32_t dp_filter_channels, dp_kernel_size, dp_n_flows, dp_num_bins;
  float dp_tail_bound;
  int32_t decoder_type;            /* 0 = Multiband_iSTFT_Generator, 1 = HiFi-GAN Generator */
  int32_t resblock_type;           /* 1 = ResBlock1, 2 = ResBlock2 */
  int32_t n_resblock_kernels;
  int32_t resblock_kernel_sizes[8];
  int32_t n_resblock_dilations;    /* dilations per resblock (same count for all) */
  int32_t resblock_dilations[8][8];
  int32_t n_upsamples;
  int32_t upsample_rates[8];
  int32_t upsample_kernel_sizes[8];
  int32_t upsample_initial_channel;
  int32_t subbands, istft_n_fft, istft_hop;
  int32_t precision;               /* 0 = fp32 FFMA everywhere; 1 = split-bf16 tcgen05 for the flow + decoder (dense convs and
                                      attention); 2 = text encoder on tcgen05 as well */
  int32_t flow_n_heads;            /* heads of the flow's pre_transformer: the reference hard-codes 2 (models.py:355) */
} vtts_config;

/* Replaces onnxruntime.InferenceSession(model.onnx) (vosk_tts/model.py:46).
 * `blob` holds the packed fp32 tensors produced by vosk_tts_b200.weights.pack(); `manifest` is a
 * NUL-terminated text table "name offset_in_floats numel\n".  `blob_is_device` != 0 means `blob`
 * is already a device pointer on `device` (e.g. the destination of an NCCL broadcast); the engine
 * then copies device-to-device.  The engine keeps its own copy either way. */
int vtts_create(const vtts_config* cfg, const float* blob, size_t blob_floats, const char* manifest,
                int blob_is_device, int device, vtts_handle* out);
void vtts_destroy(vtts_handle h);
const char* vtts_last_error(vtts_handle h);

/* Phase 1 of InferenceSession.run (models.py:1680-1691): speaker lookup, text encoder,
 * stochastic duration predictor, ceil'd durations.
 *   ids        int64 [B, t_max]   phoneme ids ("input"), rows padded arbitrarily beyond lengths
 *   lengths    int64 [B]          ("input_lengths")
 *   sid        int64 [B]          ("sid")
 *   scales     float [3]          noise_scale, length_scale, noise_scale_w ("scales")
 *   noise_dp   float [B,2,t_max]  replaces torch.randn at models.py:96, or NULL -> Philox(seed)
 *   y_lengths  out int64 [B]      frames per utterance (models.py:1691)
 *   durations  out int32 [B,t_max] w_ceil per token, or NULL
 * Host pointers.  Blocks until the lengths are known (the one data-dependent shape of the path). */
int vtts_durations(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid,
                   int B, int t_max, const float* scales, const float* noise_dp, uint64_t seed,
                   int64_t* y_lengths, int32_t* durations);

/* Phase 2 (models.py:1692-1703): hard alignment, prior expansion + sampling, flow^-1, decoder.
 *   noise_z    float [B, inter_channels, z_ld]  replaces torch.randn_like at models.py:1700 (only the
 *              first y_lengths[b] columns of utterance b are read), or NULL -> Philox(seed)
 *   wav        out float [B, wav_ld]   utterance b occupies wav[b*wav_ld .. + hop*y_lengths[b])
 *   frame_token out int32 [B, idx_ld]  frame -> token index (the alignment `attn`, models.py:1694), or NULL
 * Host pointers.  Returns VTTS_ERR_CAPACITY if wav_ld < hop*max(y_lengths). */
int vtts_synthesize(vtts_handle h, const float* noise_z, int z_ld, float* wav, int64_t wav_ld,
                    int32_t* frame_token, int idx_ld);

/* Same two phases with device-resident inputs/outputs (pointers on the engine's device):
 * used to time the path without host<->device copies.  The layouts match the host variants. */
int vtts_durations_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid,
                       int B, int t_max, const float* scales, const float* d_noise_dp, uint64_t seed,
                       int64_t* y_lengths_host);
int vtts_synthesize_dev(vtts_handle h, const float* d_noise_z, int z_ld, float* d_wav, int64_t wav_ld);

/* Both phases in one call (== one InferenceSession.run).  The caller provides capacity instead of exact sizes:
 * z_ld columns of noise_z (if given) and wav_ld samples per utterance; VTTS_ERR_CAPACITY is returned after phase 1
 * (y_lengths filled in, durations kept) when they are too small, and vtts_synthesize can then be called with
 * larger buffers. */
int vtts_infer(vtts_handle h, const int64_t* ids, const int64_t* lengths, const int64_t* sid, int B, int t_max,
               const float* scales, const float* noise_dp, const float* noise_z, int z_ld, uint64_t seed,
               int64_t* y_lengths, float* wav, int64_t wav_ld, int32_t* frame_token, int idx_ld);
int vtts_infer_dev(vtts_handle h, const int64_t* d_ids, const int64_t* lengths_host, const int64_t* d_sid, int B, int t_max,
                   const float* scales, const float* d_noise_dp, const float* d_noise_z, int z_ld, uint64_t seed,
                   int64_t* y_lengths_host, float* d_wav, int64_t wav_ld);

/* Streaming synthesis of one long utterance (BASELINE.json configs[4]): after vtts_durations, vtts_flow runs the
 * alignment, sampling and the flow once; vtts_decode_chunk then vocodes latent frames [f0, f1) with a
 * vtts_decoder_halo()-frame halo on each side that is computed and discarded ("overlap-discard": exact, because the
 * halo covers the decoder's receptive field), so audio can be handed out chunk by chunk.  B must be 1.
 * wav receives hop*(f1-f0) samples. */
int vtts_decoder_halo(vtts_handle h);
int vtts_flow(vtts_handle h, const float* noise_z, int z_ld);
int vtts_decode_chunk(vtts_handle h, int f0, int f1, float* wav, int64_t wav_capacity);

/* Samples produced per latent frame (256 for the reference config). */
int vtts_hop(vtts_handle h);
/* CUDA-event time (ms) of each stage of the last call: [0] encoder, [1] duration predictor +
 * regulator, [2] prior sampling + flow, [3] decoder, [4] H2D, [5] D2H.  n <= 8. */
int vtts_stage_timings(vtts_handle h, float* ms, int n);
/* Number of kernels this engine launched since creation (bench.py's gpu_launches). */
uint64_t vtts_kernel_launches(vtts_handle h);
/* The CUDA stream (cudaStream_t as void*) the engine launches on, for external event timing. */
void* vtts_stream(vtts_handle h);
/* Runs one named kernel micro-benchmark on the engine's device; see csrc/engine.cu. Returns ms or <0. */
float vtts_microbench(vtts_handle h, const char* what, int iters);

/* CUDA graphs (default on): a call shape (batch, lengths) seen before is captured once and replayed, which removes
 * the ~160 per-call kernel-launch overheads at batch 1.  vtts_graph_replays counts graph launches so far. */
int vtts_set_graphs(vtts_handle h, int enable);
uint64_t vtts_graph_replays(vtts_handle h);
/* Single-utterance vtts_infer / vtts_infer_dev calls enqueue the second phase for a PREDICTED length bucket without waiting
 * for the durations (the kernels read the true lengths on the device) and repeat it only when the prediction was too small:
 * hits / misses since creation. */
int vtts_speculation_stats(vtts_handle h, uint64_t* hits, uint64_t* misses);
/* Host-side wall clock (microseconds) of the last single-utterance vtts_infer call: [0] phase-1 enqueue, [1] phase-2 +
 * copy-back enqueue, [2] wait for the stream, [3] copy-out, [4] total, [5] 1 = speculative hit, 2 = miss, 0 = not speculative. */
int vtts_host_timings(vtts_handle h, double* us, int n);

/* Per-launch profiling of the dense-conv kernel family (the dominant kernels): while enabled every launch is
 * bracketed by CUDA events on the engine's stream.  vtts_profile_read returns the summed device time, the
 * number of launches and the algorithmic FLOPs (2*Cin*k*Cout per output position) since vtts_profile(h,1). */
int vtts_profile(vtts_handle h, int enable);
int vtts_profile_read(vtts_handle h, double* conv_ms, uint64_t* conv_launches, double* conv_flops);
/* Same counters for the tcgen05 conv kernel (precision mode 1). */
int vtts_profile_read_tc(vtts_handle h, double* ms, uint64_t* launches, double* flops);

/* In-graph timeline for tuning: enable=1 arms it, enable=0 disarms, enable=2 reads up to max_pairs (source line,
 * %globaltimer ns) pairs -- one per kernel launch, stamped by the kernel's first CTA at entry. */
int vtts_timeline(vtts_handle h, int enable, unsigned long long* out, size_t max_pairs, size_t* n_out);

/* Test hooks: flags bit0 keeps a copy of z_p (models.py:1700); vtts_debug_read copies a named workspace
 * tensor of the last call ("x", "stats", "dx", "za", "zb", "condv", "z_p", "z", "d0", "stage<i>", "post") to
 * host memory in the engine's channels-last packed layout. */
int vtts_debug_flags(vtts_handle h, int flags);
int vtts_debug_read(vtts_handle h, const char* name, float* out, size_t max_floats, size_t* n_out);
/* Unit-test hook for the relative-position attention kernels (attentions.py:165-196): one launch of layer "enc.<i>" or
 * "flow.<f>.tr" on a host fp32 qkv tensor [T][3H] of one utterance; out receives fp32 [T][H].  use_tc = 1 selects the
 * tcgen05 kernel (attn_tc.cuh), 0 the fp32 FFMA kernels.  iters > 0: *ms_out = average device time of `iters` more launches. */
int vtts_debug_attention(vtts_handle h, const char* layer, const float* qkv_host, int T, int use_tc, float* out_host, int iters,
                         float* ms_out);

/* Monotonic Alignment Search on the GPU -- replaces monotonic_align.maximum_path (training/vits2/monotonic_align/__init__.py:6-22,
 * core.pyx:7-43; called from SynthesizerTrn.forward, models.py:1658).  Handle-free (no engine state); errors of these two are
 * read with vtts_last_error(NULL) on the calling thread.
 * neg_cent: float32 [B][T_y][T_x] scores (frames x tokens, as the reference passes them); t_ys / t_xs: valid frames / tokens
 * per item (the reference derives them from the mask sums), 0 <= t_x <= t_y required; path: int32 [B][T_y][T_x], 1 on the path.
 * Host pointers; runs on `device`.  The _dev variant takes device pointers, overwrites d_value with the accumulated scores
 * (like the reference's in-place update) and only enqueues on `stream` (a cudaStream_t, may be NULL). */
int vtts_maximum_path(const float* neg_cent, const int32_t* t_ys, const int32_t* t_xs, int B, int T_y, int T_x, int32_t* path,
                      int device);
int vtts_maximum_path_dev(float* d_value, const int32_t* d_t_ys, const int32_t* d_t_xs, int B, int T_y, int T_x, int32_t* d_path,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VTTS_H_ */
