"""ctypes binding of libvtts.so (include/vtts.h).  No CPU fallback: importing works without a GPU
(the library only needs libcudart), creating an Engine requires a CUDA device."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None


class VttsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vtts error %d: %s" % (code, msg))
        self.code = code


class VttsConfig(C.Structure):
    _fields_ = [
        ("n_vocab", C.c_int32), ("n_speakers", C.c_int32), ("gin_channels", C.c_int32),
        ("inter_channels", C.c_int32), ("hidden_channels", C.c_int32), ("filter_channels", C.c_int32),
        ("n_heads", C.c_int32), ("n_layers", C.c_int32), ("kernel_size", C.c_int32), ("window_size", C.c_int32),
        ("spk_cond_encoder", C.c_int32), ("cond_layer_idx", C.c_int32),
        ("use_transformer_flows", C.c_int32),
        ("flow_kernel_size", C.c_int32), ("flow_dilation_rate", C.c_int32), ("flow_wn_layers", C.c_int32),
        ("flow_n_flows", C.c_int32),
        ("dp_filter_channels", C.c_int32), ("dp_kernel_size", C.c_int32), ("dp_n_flows", C.c_int32),
        ("dp_num_bins", C.c_int32),
        ("dp_tail_bound", C.c_float),
        ("decoder_type", C.c_int32), ("resblock_type", C.c_int32),
        ("n_resblock_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * 8),
        ("n_resblock_dilations", C.c_int32), ("resblock_dilations", (C.c_int32 * 8) * 8),
        ("n_upsamples", C.c_int32), ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8),
        ("upsample_initial_channel", C.c_int32),
        ("subbands", C.c_int32), ("istft_n_fft", C.c_int32), ("istft_hop", C.c_int32),
        ("precision", C.c_int32), ("flow_n_heads", C.c_int32),
    ]


EXPORTS = ["vtts_create", "vtts_destroy", "vtts_last_error", "vtts_durations", "vtts_synthesize",
           "vtts_durations_dev", "vtts_synthesize_dev", "vtts_hop", "vtts_stage_timings",
           "vtts_kernel_launches", "vtts_stream", "vtts_microbench", "vtts_debug_flags", "vtts_debug_read",
           "vtts_profile", "vtts_profile_read", "vtts_set_graphs", "vtts_graph_replays",
           "vtts_profile_read_tc", "vtts_timeline", "vtts_infer", "vtts_infer_dev",
           "vtts_decoder_halo", "vtts_flow", "vtts_decode_chunk", "vtts_debug_attention", "vtts_speculation_stats", "vtts_host_timings",
           "vtts_maximum_path", "vtts_maximum_path_dev"]


class _Missing:
    """Stand-in for an entry point an alternative build (VTTS_LIB) does not export: accepts the argtypes / restype
    assignments of load_library and raises when called."""

    def __init__(self, name):
        self._name = name

    def __call__(self, *a):
        raise RuntimeError("%s is not exported by the library selected with VTTS_LIB" % self._name)


class _TolerantLib:
    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_missing", {})

    def __getattr__(self, name):
        try:
            return getattr(self._lib, name)
        except AttributeError:
            return self._missing.setdefault(name, _Missing(name))


def lib_path():
    return _build.LIB


def load_library(build_if_missing=True):
    """dlopen the in-tree libvtts.so; fails loudly if it is missing and cannot be built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("VTTS_LIB") or _build.LIB       # VTTS_LIB: load an alternative build (kernel A/B experiments)
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError("libvtts.so is missing: run `python -m vosk_tts_b200.build`")
        _build.build()
    lib = C.CDLL(path)
    if os.environ.get("VTTS_LIB"):
        lib = _TolerantLib(lib)                 # an older build loaded for an A/B may lack the newest entry points
    vp, i32, i64p, fp = C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_float)
    lib.vtts_create.argtypes = [C.POINTER(VttsConfig), vp, C.c_size_t, C.c_char_p, i32, i32, C.POINTER(vp)]
    lib.vtts_create.restype = i32
    lib.vtts_destroy.argtypes = [vp]
    lib.vtts_destroy.restype = None
    lib.vtts_last_error.argtypes = [vp]
    lib.vtts_last_error.restype = C.c_char_p
    lib.vtts_durations.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, C.c_uint64, vp, vp]
    lib.vtts_durations.restype = i32
    lib.vtts_synthesize.argtypes = [vp, vp, i32, vp, C.c_int64, vp, i32]
    lib.vtts_synthesize.restype = i32
    lib.vtts_durations_dev.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, C.c_uint64, vp]
    lib.vtts_durations_dev.restype = i32
    lib.vtts_synthesize_dev.argtypes = [vp, vp, i32, vp, C.c_int64]
    lib.vtts_synthesize_dev.restype = i32
    lib.vtts_infer.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, i32, C.c_uint64, vp, vp, C.c_int64, vp, i32]
    lib.vtts_infer.restype = i32
    lib.vtts_infer_dev.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, vp, i32, C.c_uint64, vp, vp, C.c_int64]
    lib.vtts_infer_dev.restype = i32
    lib.vtts_decoder_halo.argtypes = [vp]
    lib.vtts_decoder_halo.restype = i32
    lib.vtts_flow.argtypes = [vp, vp, i32]
    lib.vtts_flow.restype = i32
    lib.vtts_decode_chunk.argtypes = [vp, i32, i32, vp, C.c_int64]
    lib.vtts_decode_chunk.restype = i32
    lib.vtts_hop.argtypes = [vp]
    lib.vtts_hop.restype = i32
    lib.vtts_stage_timings.argtypes = [vp, vp, i32]
    lib.vtts_stage_timings.restype = i32
    lib.vtts_kernel_launches.argtypes = [vp]
    lib.vtts_kernel_launches.restype = C.c_uint64
    lib.vtts_stream.argtypes = [vp]
    lib.vtts_stream.restype = vp
    lib.vtts_microbench.argtypes = [vp, C.c_char_p, i32]
    lib.vtts_microbench.restype = C.c_float
    lib.vtts_debug_flags.argtypes = [vp, i32]
    lib.vtts_debug_flags.restype = i32
    lib.vtts_debug_read.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.vtts_debug_read.restype = i32
    lib.vtts_set_graphs.argtypes = [vp, i32]
    lib.vtts_set_graphs.restype = i32
    lib.vtts_graph_replays.argtypes = [vp]
    lib.vtts_graph_replays.restype = C.c_uint64
    lib.vtts_profile.argtypes = [vp, i32]
    lib.vtts_profile.restype = i32
    lib.vtts_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    lib.vtts_timeline.argtypes = [vp, i32, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.vtts_timeline.restype = i32
    lib.vtts_profile_read.restype = i32
    lib.vtts_profile_read_tc.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    lib.vtts_profile_read_tc.restype = i32
    lib.vtts_debug_attention.argtypes = [vp, C.c_char_p, vp, i32, i32, vp, i32, fp]
    lib.vtts_debug_attention.restype = i32
    lib.vtts_speculation_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.vtts_speculation_stats.restype = i32
    lib.vtts_host_timings.argtypes = [vp, C.POINTER(C.c_double), i32]
    lib.vtts_host_timings.restype = i32
    lib.vtts_maximum_path.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32]
    lib.vtts_maximum_path.restype = i32
    lib.vtts_maximum_path_dev.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    lib.vtts_maximum_path_dev.restype = i32
    _LIB = lib
    return lib


def make_c_config(cfg, precision=0):
    c = VttsConfig()
    for k in ("n_vocab", "n_speakers", "gin_channels", "inter_channels", "hidden_channels", "filter_channels",
              "n_heads", "n_layers", "kernel_size", "window_size", "cond_layer_idx", "flow_kernel_size",
              "flow_dilation_rate", "flow_wn_layers", "flow_n_flows", "dp_filter_channels", "dp_kernel_size",
              "dp_n_flows", "dp_num_bins", "upsample_initial_channel", "subbands"):
        setattr(c, k, int(cfg[k]))
    c.spk_cond_encoder = int(bool(cfg["use_spk_conditioned_encoder"]) and cfg["gin_channels"] > 0 and cfg["n_speakers"] > 0)
    c.use_transformer_flows = int(bool(cfg["use_transformer_flows"]))
    c.dp_tail_bound = float(cfg["dp_tail_bound"])
    # 0: conv_post -> exp / pi*sin -> inverse STFT -> 63-tap filter bank (Multiband_ / Multistream_ / plain iSTFT_Generator differ only
    #    in the bank: fixed PQMF, learned, unit impulse; models.py:901-971, 974-1063, 1066-1169); 1: HiFi-GAN Generator
    c.decoder_type = 0 if cfg["decoder"] in ("mb_istft", "ms_istft", "istft") else 1
    c.resblock_type = 1 if str(cfg["resblock"]) == "1" else 2
    rk, rd = cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]
    c.n_resblock_kernels = len(rk)
    c.n_resblock_dilations = len(rd[0])
    for j, k in enumerate(rk):
        c.resblock_kernel_sizes[j] = int(k)
        assert len(rd[j]) == len(rd[0])
        for d, v in enumerate(rd[j]):
            c.resblock_dilations[j][d] = int(v)
    c.n_upsamples = len(cfg["upsample_rates"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        c.upsample_rates[i] = int(u)
        c.upsample_kernel_sizes[i] = int(k)
    c.istft_n_fft = int(cfg["gen_istft_n_fft"])
    c.istft_hop = int(cfg["gen_istft_hop_size"])
    c.precision = int(precision)
    c.flow_n_heads = int(cfg.get("flow_n_heads", 2))
    return c


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    """One engine = one GPU.  `blob` may be a numpy float32 array (host) or an (int device_ptr, n_floats) tuple."""

    def __init__(self, cfg, blob, manifest, device=0, precision=0):
        self.lib = load_library()
        self.cfg = cfg
        self.h = C.c_void_p()
        ccfg = make_c_config(cfg, precision)
        if isinstance(blob, tuple):
            ptr, n, on_dev = C.c_void_p(int(blob[0])), int(blob[1]), 1
        else:
            blob = np.ascontiguousarray(blob, dtype=np.float32)
            ptr, n, on_dev = _ptr(blob), blob.size, 0
        rc = self.lib.vtts_create(C.byref(ccfg), ptr, n, manifest.encode(), on_dev, int(device), C.byref(self.h))
        if rc != 0:
            msg = self.lib.vtts_last_error(self.h).decode() if self.h else "allocation failed"
            if self.h:
                self.lib.vtts_destroy(self.h)
                self.h = C.c_void_p()
            raise VttsError(rc, msg)
        self.hop = self.lib.vtts_hop(self.h)
        self.device = device
        import threading
        self._tls = threading.local()

    def close(self):
        if getattr(self, "h", None):
            self.lib.vtts_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise VttsError(rc, self.lib.vtts_last_error(self.h).decode())

    # ---- host-buffer path (what the reference-facing session calls)
    def durations(self, ids, lengths, sid, scales, noise_dp=None, seed=0, want_durations=False):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        if ids.ndim == 1:
            ids = ids[None, :]
        B, t_max = ids.shape
        lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(B)
        sid = np.ascontiguousarray(sid, dtype=np.int64).reshape(B)
        scales = np.ascontiguousarray(scales, dtype=np.float32).reshape(3)
        if noise_dp is not None:
            noise_dp = np.ascontiguousarray(noise_dp, dtype=np.float32).reshape(B, 2, t_max)
        y_len = np.zeros(B, np.int64)
        dur = np.zeros((B, t_max), np.int32) if want_durations else None
        self._check(self.lib.vtts_durations(self.h, _ptr(ids), _ptr(lengths), _ptr(sid), B, t_max, _ptr(scales),
                                            _ptr(noise_dp), int(seed), _ptr(y_len), _ptr(dur)))
        self._B = B
        return (y_len, dur) if want_durations else y_len

    def synthesize(self, y_lengths, noise_z=None, want_alignment=False):
        B = self._B
        max_f = int(np.max(y_lengths))
        wav = np.zeros((B, max_f * self.hop), np.float32)
        z_ld = 0
        if noise_z is not None:
            noise_z = np.ascontiguousarray(noise_z, dtype=np.float32)
            assert noise_z.ndim == 3 and noise_z.shape[0] == B
            z_ld = noise_z.shape[2]
        idx = np.full((B, max_f), -1, np.int32) if want_alignment else None
        self._check(self.lib.vtts_synthesize(self.h, _ptr(noise_z), z_ld, _ptr(wav), wav.shape[1], _ptr(idx), max_f))
        return (wav, idx) if want_alignment else wav

    # Thread-safe variants of the two-phase pair: no per-Engine Python state (the batch size travels with the caller)
    def lib_durations_threadsafe(self, ids, lengths, sid, scales, noise_dp=None, seed=0):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        B, t_max = ids.shape
        lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(B)
        sid = np.ascontiguousarray(sid, dtype=np.int64).reshape(B)
        scales = np.ascontiguousarray(scales, dtype=np.float32).reshape(3)
        if noise_dp is not None:
            noise_dp = np.ascontiguousarray(noise_dp, dtype=np.float32).reshape(B, 2, t_max)
        y_len = np.zeros(B, np.int64)
        self._check(self.lib.vtts_durations(self.h, _ptr(ids), _ptr(lengths), _ptr(sid), B, t_max, _ptr(scales), _ptr(noise_dp), int(seed),
                                            _ptr(y_len), None))
        return y_len

    def lib_synthesize_threadsafe(self, B, y_lengths, noise_z=None):
        max_f = int(np.max(y_lengths))
        wav = np.zeros((B, max_f * self.hop), np.float32)
        z_ld = 0
        if noise_z is not None:
            noise_z = np.ascontiguousarray(noise_z, dtype=np.float32)
            z_ld = noise_z.shape[2]
        self._check(self.lib.vtts_synthesize(self.h, _ptr(noise_z), z_ld, _ptr(wav), wav.shape[1], None, 0))
        return wav

    def infer(self, ids, lengths, sid, scales, noise_dp=None, noise_z=None, seed=0, frames_hint=None):
        """Both phases.  noise_z may be a callable(max_frames)->[B,C,max_frames] (T_y is data dependent).
        With `frames_hint` (an upper bound on max(y_lengths), e.g. from a previous call) and array/None noise, the
        fused C entry point vtts_infer is used: one ctypes call, no Python between the two phases."""
        if frames_hint is not None and not callable(noise_z):
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            if ids.ndim == 1:
                ids = ids[None, :]
            B, t_max = ids.shape
            lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(B)
            sid = np.ascontiguousarray(sid, dtype=np.int64).reshape(B)
            scales = np.ascontiguousarray(scales, dtype=np.float32).reshape(3)
            if noise_dp is not None:
                noise_dp = np.ascontiguousarray(noise_dp, dtype=np.float32).reshape(B, 2, t_max)
            z_ld = 0
            if noise_z is not None:
                noise_z = np.ascontiguousarray(noise_z, dtype=np.float32)
                z_ld = noise_z.shape[2]
            y_len = np.zeros(B, np.int64)
            # capacity-sized scratch rows, reused by this thread's calls (a fresh 1 MB numpy buffer per call costs mmap / page
            # faults / munmap -- 0.1-0.3 ms of a 1.7 ms call); the engine writes y_len*hop samples per row
            tl = self._tls
            need = int(frames_hint) * self.hop
            wav = getattr(tl, "wav", None)
            if wav is None or wav.shape[0] < B or wav.shape[1] < need:
                wav = np.empty((B, need), np.float32)
                tl.wav = wav
            wav = wav[:B]
            rc = self.lib.vtts_infer(self.h, _ptr(ids), _ptr(lengths), _ptr(sid), B, t_max, _ptr(scales), _ptr(noise_dp),
                                     _ptr(noise_z), z_ld, int(seed), _ptr(y_len), _ptr(wav), wav.shape[1], None, 0)
            self._B = B
            if rc == -4:            # capacity: durations are kept, finish with exact buffers
                if noise_z is not None and z_ld < int(y_len.max()):
                    self._check(rc)
                return self.synthesize(y_len, noise_z), y_len
            self._check(rc)
            n = int(y_len.max()) * self.hop
            out = np.zeros((B, n), np.float32)                           # contiguous result sized to max(y_len); rows are zero beyond their length
            for b in range(B):
                m = int(y_len[b]) * self.hop
                out[b, :m] = wav[b, :m]
            return out, y_len
        y_len = self.durations(ids, lengths, sid, scales, noise_dp, seed)
        if callable(noise_z):
            noise_z = noise_z(int(y_len.max()))
        wav = self.synthesize(y_len, noise_z)
        return wav, y_len

    def infer_dev(self, d_ids, lengths, d_sid, B, t_max, scales, d_wav, wav_ld, d_noise_dp=0, d_noise_z=0, z_ld=0, seed=0):
        lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(B)
        scales = np.ascontiguousarray(scales, dtype=np.float32).reshape(3)
        y_len = np.zeros(B, np.int64)
        self._check(self.lib.vtts_infer_dev(self.h, C.c_void_p(d_ids), _ptr(lengths), C.c_void_p(d_sid), B, t_max, _ptr(scales),
                                            C.c_void_p(d_noise_dp) if d_noise_dp else None,
                                            C.c_void_p(d_noise_z) if d_noise_z else None, z_ld, int(seed), _ptr(y_len),
                                            C.c_void_p(d_wav), wav_ld))
        self._B = B
        return y_len

    def reserve(self, max_tokens=256, max_frames=1024, batch=1):
        """Sizes the engine's workspace (device buffers, plane pools, pinned staging) once for calls of up to `batch`
        utterances x `max_tokens` phonemes x `max_frames` frames, by synthesising one synthetic request of that size.
        Any LATER growth of the workspace moves buffers and therefore invalidates every captured CUDA graph (each length
        bucket then pays its capture again, ~15 ms); a service calls this once at start-up, like the reference server
        warms its ONNX session.  Returns the frame count that was reached."""
        T, B = int(max_tokens), int(batch)
        nv = int(self.cfg["n_vocab"])
        ids = (np.arange(B * T, dtype=np.int64).reshape(B, T) * 7 + 1) % nv
        lens, sid = [T] * B, [0] * B
        e1 = np.zeros((B, 2, T), np.float32)
        ls, F, yl = 1.0, 0, None
        for _ in range(5):
            yl = self.durations(ids, lens, sid, (0.667, ls, 0.8), e1)
            F = int(yl.max())
            if F >= max_frames:
                break
            ls *= max_frames / max(F, 1) * 1.05
        ez = np.zeros((B, int(self.cfg["inter_channels"]), F), np.float32)
        self.synthesize(yl, ez)
        self.infer(ids, lens, sid, (0.667, ls, 0.8), None, None, seed=1, frames_hint=F + 64)
        return F

    # ---- streaming (one utterance): flow once, then vocode chunk by chunk
    def synthesize_stream(self, ids, sid, scales, chunk_frames=64, noise_dp=None, noise_z=None, seed=0):
        """Generator of float32 chunks; concatenated they equal `infer(...)` of the same inputs."""
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(1, -1)
        y_len = self.durations(ids, [ids.shape[1]], [sid], scales, noise_dp, seed)
        T = int(y_len[0])
        z_ld = 0
        if noise_z is not None:
            noise_z = np.ascontiguousarray(noise_z, dtype=np.float32)
            z_ld = noise_z.shape[2]
        self._check(self.lib.vtts_flow(self.h, _ptr(noise_z), z_ld))
        for f0 in range(0, T, chunk_frames):
            f1 = min(T, f0 + chunk_frames)
            out = np.zeros((f1 - f0) * self.hop, np.float32)
            self._check(self.lib.vtts_decode_chunk(self.h, f0, f1, _ptr(out), out.size))
            yield out

    # ---- device-buffer path (raw pointers, e.g. torch tensors' data_ptr())
    def durations_dev(self, d_ids, lengths, d_sid, B, t_max, scales, d_noise_dp=0, seed=0):
        lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(B)
        scales = np.ascontiguousarray(scales, dtype=np.float32).reshape(3)
        y_len = np.zeros(B, np.int64)
        self._check(self.lib.vtts_durations_dev(self.h, C.c_void_p(d_ids), _ptr(lengths), C.c_void_p(d_sid), B, t_max,
                                                _ptr(scales), C.c_void_p(d_noise_dp) if d_noise_dp else None,
                                                int(seed), _ptr(y_len)))
        self._B = B
        return y_len

    def synthesize_dev(self, d_wav, wav_ld, d_noise_z=0, z_ld=0):
        self._check(self.lib.vtts_synthesize_dev(self.h, C.c_void_p(d_noise_z) if d_noise_z else None, z_ld,
                                                 C.c_void_p(d_wav), wav_ld))

    def stage_timings(self):
        ms = np.zeros(8, np.float32)
        self.lib.vtts_stage_timings(self.h, _ptr(ms), 8)
        return dict(encoder=float(ms[0]), duration=float(ms[1]), flow=float(ms[2]), decoder=float(ms[3]),
                    h2d=float(ms[4]), d2h=float(ms[5]))

    def kernel_launches(self):
        return int(self.lib.vtts_kernel_launches(self.h))

    def stream(self):
        return int(self.lib.vtts_stream(self.h) or 0)

    def microbench(self, what, iters=50):
        ms = float(self.lib.vtts_microbench(self.h, what.encode(), int(iters)))
        if ms < 0:
            raise VttsError(int(ms), self.lib.vtts_last_error(self.h).decode())
        return ms

    def timeline(self, mode):
        """mode 1: arm, 0: disarm, 2: read -> array [n,2] of (source line, globaltimer ns)."""
        if mode != 2:
            self._check(self.lib.vtts_timeline(self.h, int(mode), None, 0, None))
            return None
        out = np.zeros((4000, 2), np.uint64)
        n = C.c_size_t(0)
        self._check(self.lib.vtts_timeline(self.h, 2, _ptr(out), 4000, C.byref(n)))
        return out[: n.value].copy()

    def set_graphs(self, enable):
        self._check(self.lib.vtts_set_graphs(self.h, int(bool(enable))))

    def graph_replays(self):
        return int(self.lib.vtts_graph_replays(self.h))

    def speculation_stats(self):
        """(hits, misses) of the speculative second phase of single-utterance infer calls."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.vtts_speculation_stats(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def host_timings(self):
        """Host-side microseconds of the last single-utterance infer call (see vtts_host_timings)."""
        a = (C.c_double * 8)()
        self._check(self.lib.vtts_host_timings(self.h, a, 8))
        return [float(v) for v in a]

    def profile(self, enable):
        self._check(self.lib.vtts_profile(self.h, int(bool(enable))))

    def profile_read(self):
        ms, n, fl = C.c_double(0), C.c_uint64(0), C.c_double(0)
        self._check(self.lib.vtts_profile_read(self.h, C.byref(ms), C.byref(n), C.byref(fl)))
        out = dict(conv_ms=ms.value, conv_launches=int(n.value), conv_flops=fl.value)
        self._check(self.lib.vtts_profile_read_tc(self.h, C.byref(ms), C.byref(n), C.byref(fl)))
        out.update(tc_ms=ms.value, tc_launches=int(n.value), tc_flops=fl.value)
        return out

    def debug_attention(self, layer, qkv, use_tc, iters=0):
        """One attention launch of `layer` ("enc.<i>" / "flow.<f>.tr") on qkv float32 [T, 3H]; returns (out [T, H], ms or None)."""
        qkv = np.ascontiguousarray(qkv, dtype=np.float32)
        T, H = qkv.shape[0], qkv.shape[1] // 3
        out = np.zeros((T, H), np.float32)
        ms = C.c_float(0.0)
        self._check(self.lib.vtts_debug_attention(self.h, layer.encode(), _ptr(qkv), T, int(use_tc), _ptr(out), int(iters), C.byref(ms)))
        return out, (float(ms.value) if iters > 0 else None)

    def debug_flags(self, flags):
        self._check(self.lib.vtts_debug_flags(self.h, int(flags)))

    def debug_read(self, name, max_floats=1 << 26):
        out = np.zeros(max_floats, np.float32)
        n = C.c_size_t(0)
        self._check(self.lib.vtts_debug_read(self.h, name.encode(), _ptr(out), max_floats, C.byref(n)))
        return out[: n.value].copy()
