"""`vosk_tts.Synth`-compatible front end (vosk_tts/synth.py:11-150) over the CUDA engine.

`synth_audio` / `synth` keep the reference signatures, defaults (config["inference"], synth.py:49-56), the
float->int16 conversion (:16-23), the RTF log line (:133-139) and the 22 050 Hz mono 16-bit WAV (:146-150).  Only the
VITS branch of the model_type dispatch (`g2p_noembed`, :100-103, :223-255) exists here; the other branches feed
graphs this engine does not implement and raise.
"""
import logging
import re
import time
import wave

import numpy as np

from .g2p import convert

_PUNCT = "([,.?!;:\"() ])"


class Synth:
    def __init__(self, model):
        self.model = model

    def audio_float_to_int16(self, audio, max_wav_value=32767.0):
        audio_norm = np.clip(audio * max_wav_value, -max_wav_value, max_wav_value)
        return audio_norm.astype("int16")

    def g2p_noembed(self, text):
        phonemes = ["^"]
        for word in re.split(_PUNCT, text.lower()):
            if word == "":
                continue
            if re.match(_PUNCT, word) or word == "-":
                phonemes.append(word)
            elif word in self.model.dic:
                phonemes.extend(self.model.dic[word].split())
            else:
                phonemes.extend(convert(word).split())
        phonemes.append("$")
        id_map = self.model.config["phoneme_id_map"]
        ids = []
        for i, p in enumerate(phonemes):            # intersperse the blank id 0 (synth.py:244-251)
            if i:
                ids.append(0)
            v = id_map[p]
            ids.extend(v if isinstance(v, list) else [v])
        logging.info(f"Text: {text}")
        logging.info(f"Phonemes: {phonemes}")
        return ids

    def synth_audio(self, text, speaker_id=0, noise_level=None, speech_rate=None, duration_noise_level=None, scale=None):
        inf = self.model.config.get("inference", {})
        noise_level = inf.get("noise_level", 0.8) if noise_level is None else noise_level
        speech_rate = inf.get("speech_rate", 1.0) if speech_rate is None else speech_rate
        duration_noise_level = inf.get("duration_noise_level", 0.8) if duration_noise_level is None else duration_noise_level
        scale = inf.get("scale", 1.0) if scale is None else scale
        if self.model.tokenizer is not None or str(self.model.config.get("model_type", "")).startswith("multistream"):
            raise ValueError("model_type %r is not a VITS2 graph: not supported by this engine" % self.model.config.get("model_type"))
        text = re.sub("—", "-", text.strip())
        ids = self.g2p_noembed(text)
        feeds = {"input": np.expand_dims(np.array(ids, dtype=np.int64), 0),
                 "input_lengths": np.array([len(ids)], dtype=np.int64),
                 "scales": np.array([noise_level, 1.0 / speech_rate, duration_noise_level], dtype=np.float32),
                 "sid": np.array([0 if speaker_id is None else speaker_id], dtype=np.int64),
                 "bert": None, "phone_duration_extra": None}
        t0 = time.perf_counter()
        audio = self.model.onnx.run(None, feeds)[0].squeeze() * scale
        audio = self.audio_float_to_int16(audio)
        infer_sec = time.perf_counter() - t0
        dur = audio.shape[-1] / 22050
        logging.info("Real-time factor: %0.2f (infer=%0.2f sec, audio=%0.2f sec)" % (infer_sec / dur if dur > 0 else 0.0, infer_sec, dur))
        return audio

    def synth_audio_stream(self, text, speaker_id=0, noise_level=None, speech_rate=None, duration_noise_level=None, scale=None,
                           chunk_frames=64):
        """Generator of int16 chunks of the same utterance `synth_audio` would return (extension; the reference server sends one
        message with the whole utterance, server/tts_server.py:55-56): first audio after encoder + flow + one decoder window."""
        inf = self.model.config.get("inference", {})
        noise_level = inf.get("noise_level", 0.8) if noise_level is None else noise_level
        speech_rate = inf.get("speech_rate", 1.0) if speech_rate is None else speech_rate
        duration_noise_level = inf.get("duration_noise_level", 0.8) if duration_noise_level is None else duration_noise_level
        scale = inf.get("scale", 1.0) if scale is None else scale
        if self.model.tokenizer is not None or str(self.model.config.get("model_type", "")).startswith("multistream"):
            raise ValueError("model_type %r is not a VITS2 graph: not supported by this engine" % self.model.config.get("model_type"))
        text = re.sub("—", "-", text.strip())
        ids = self.g2p_noembed(text)
        feeds = {"input": np.expand_dims(np.array(ids, dtype=np.int64), 0),
                 "input_lengths": np.array([len(ids)], dtype=np.int64),
                 "scales": np.array([noise_level, 1.0 / speech_rate, duration_noise_level], dtype=np.float32),
                 "sid": np.array([0 if speaker_id is None else speaker_id], dtype=np.int64),
                 "bert": None, "phone_duration_extra": None}
        t0 = time.perf_counter()
        n = 0
        stream = self.model.onnx.run_stream(feeds, chunk_frames=chunk_frames)
        try:
            for chunk in stream:
                n += chunk.size
                yield self.audio_float_to_int16(chunk * scale)
        finally:
            stream.close()              # releases the session (its lock) also when the consumer abandons this generator
        infer_sec = time.perf_counter() - t0
        dur = n / 22050
        logging.info("Real-time factor: %0.2f (infer=%0.2f sec, audio=%0.2f sec)" % (infer_sec / dur if dur > 0 else 0.0, infer_sec, dur))

    def synth(self, text, oname, speaker_id=0, noise_level=None, speech_rate=None, duration_noise_level=None, scale=None):
        audio = self.synth_audio(text, speaker_id, noise_level, speech_rate, duration_noise_level, scale)
        with wave.open(oname, "w") as f:
            f.setnchannels(1)
            f.setsampwidth(2)
            f.setframerate(22050)
            f.writeframes(audio.tobytes())
