"""`onnxruntime.InferenceSession`-shaped facade over the CUDA engine.

The reference creates `self.onnx = onnxruntime.InferenceSession(model.onnx)` (vosk_tts/model.py:46) and
calls `self.model.onnx.run(None, args)[0]` (vosk_tts/synth.py:123-126).  `VitsSession` is assignable to
`Model.onnx`: `run(None, feeds)` takes the same feeds dict (keys `input`, `input_lengths`, `scales`,
`sid`; `bert` / `phone_duration_extra` must be None, synth.py:118-119) and returns
`[float32 [B,1,1,T_wav]]` like the exported graph (onnx_export.py:65-72).
"""
import logging
import threading

import numpy as np

from . import config as _config
from . import weights as _weights
from .engine import Engine


class VitsSession:
    def __init__(self, state_dict=None, cfg=None, device=0, seed=0, packed=None, precision=0, reserve=None):
        """state_dict: reference checkpoint `['model']` dict (weight_g/weight_v allowed) or already folded.
        packed: optional (blob, manifest) to skip packing (e.g. received through an NCCL broadcast).
        reserve: optional (max_tokens, max_frames[, batch]) -- size the workspace for such requests now (Engine.reserve)."""
        self.cfg = cfg or _config.DEFAULT_CONFIG
        if precision > 0 and not _weights.tc_supported(self.cfg):
            logging.warning("model widths are not multiples of 64: the tcgen05 conv path is unavailable, using the fp32 kernels")
            precision = 0
        if packed is None:
            folded = _weights.fold_weight_norm(state_dict)
            packed = _weights.pack(folded, self.cfg)
        self.engine = Engine(self.cfg, packed[0], packed[1], device=device, precision=precision)
        self._lock = threading.Lock()
        self._seed = int(seed)
        self._calls = 0
        self.last_y_lengths = None
        self.last_wav_lengths = None
        if reserve:
            self.reserve(*reserve)

    def reserve(self, max_tokens=256, max_frames=1024, batch=1):
        """Workspace reservation (see Engine.reserve): later calls within these bounds never move a buffer, so the CUDA graphs
        of the length buckets stay valid."""
        with self._lock:
            return self.engine.reserve(max_tokens, max_frames, batch)

    # -- onnxruntime-compatible surface ---------------------------------------------------------
    def get_providers(self):
        return ["B200VttsExecutionProvider"]

    def run(self, output_names, feeds, noise=None):
        """feeds as built at vosk_tts/synth.py:113-120.  `noise` (extension): dict(dp=[B,2,T], z=[B,C,>=T_y] or
        callable(max_frames)) to inject the two random draws; otherwise Philox with a per-call seed."""
        for k in ("bert", "phone_duration_extra"):
            if feeds.get(k) is not None:
                raise ValueError("feed %r is not None: model_type not supported by this engine (VITS2 graph only)" % k)
        ids = np.asarray(feeds["input"])
        if ids.ndim != 2:
            raise ValueError("multistream inputs ([1,5,T]) are not supported by this engine (VITS2 graph only)")
        lengths = np.asarray(feeds["input_lengths"]).reshape(-1)
        sid = feeds.get("sid")
        sid = np.zeros(ids.shape[0], np.int64) if sid is None else np.asarray(sid).reshape(-1)
        scales = np.asarray(feeds["scales"], dtype=np.float32).reshape(3)
        with self._lock:
            self._calls += 1
            seed = (self._seed * 0x9E3779B97F4A7C15 + self._calls) & 0xFFFFFFFFFFFFFFFF
            dp = z = None
            if noise is not None:
                dp, z = noise.get("dp"), noise.get("z")
            # one fused C call when the output capacity can be bounded up front (8 frames per phoneme covers the
            # duration predictor's range in practice; a CAPACITY status falls back to an exact second phase)
            hint = None if callable(z) else int(8 * ids.shape[1] + 16)
            if z is not None and not callable(z):
                hint = min(hint, int(np.asarray(z).shape[2]))
            wav, y_len = self.engine.infer(ids, lengths, sid, scales, dp, z, seed, frames_hint=hint)
            self.last_y_lengths = y_len
            self.last_wav_lengths = y_len * self.engine.hop
        return [wav[:, None, None, :]]

    def run_stream(self, feeds, chunk_frames=64):
        """Streaming variant for ONE utterance (no onnxruntime equivalent): the text encoder, duration predictor and flow
        run once, then the decoder is run over windows of `chunk_frames` frames (with the halo the engine reports) and every
        window's samples are yielded as float32 [n] as soon as they are on the host.  Concatenated, the chunks are what
        `run` returns for the same seed.  The handle keeps the flow output between the chunk calls, so the session lock is
        held for the whole stream (released when the generator is exhausted or closed)."""
        for k in ("bert", "phone_duration_extra"):
            if feeds.get(k) is not None:
                raise ValueError("feed %r is not None: model_type not supported by this engine (VITS2 graph only)" % k)
        ids = np.asarray(feeds["input"])
        if ids.ndim != 2 or ids.shape[0] != 1:
            raise ValueError("run_stream takes one utterance ([1, T] ids)")
        sid = feeds.get("sid")
        sid = 0 if sid is None else int(np.asarray(sid).reshape(-1)[0])
        scales = np.asarray(feeds["scales"], dtype=np.float32).reshape(3)
        n = int(np.asarray(feeds["input_lengths"]).reshape(-1)[0])
        with self._lock:
            self._calls += 1
            seed = (self._seed * 0x9E3779B97F4A7C15 + self._calls) & 0xFFFFFFFFFFFFFFFF
            total = 0
            stream = self.engine.synthesize_stream(ids[:, :n], sid, scales, chunk_frames=chunk_frames, seed=seed)
            try:
                for chunk in stream:
                    total += chunk.size
                    yield chunk
            finally:
                # closing THIS generator does not finalise a generator it iterates over (the frame keeps it alive): close it
                # explicitly so that an abandoned stream (client gone) ends here and the lock below is released now
                stream.close()
            self.last_wav_lengths = np.array([total], np.int64)
            self.last_y_lengths = self.last_wav_lengths // self.engine.hop

    def close(self):
        self.engine.close()
