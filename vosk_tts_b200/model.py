"""`vosk_tts.Model`-compatible loader (vosk_tts/model.py:33-63) whose `.onnx` attribute is the CUDA engine session.

Same constructor and attributes (`onnx`, `dic`, `config`, `tokenizer`); the model directory holds what the reference
ships (`config.json`, `dictionary`) plus the checkpoint the ONNX graph was exported from (`G_*.pth` / `model.pth`,
training/vits2/onnx_export.py:55) and, optionally, the training json (`vits_config.json`) when `config.json` has no
`model` block.  A directory that holds only what vosk-tts ships (`model.onnx`, `config.json`, `dictionary`) works too: the
weights AND the architecture are read from the graph's initializers (`onnx_weights.py`, no `onnx` package needed).
Downloading models needs a network and is out of scope -- a missing model is an error, never a silent fallback.
"""
import glob
import json
import logging
import os
import re
from pathlib import Path

from . import config as _config
from . import weights as _weights
from .session import VitsSession

MODEL_DIRS = [os.getenv("VOSK_MODEL_PATH"), Path("/usr/share/vosk"), Path.home() / "AppData/Local/vosk",
              Path.home() / ".cache/vosk"]


def list_models():
    raise RuntimeError("model listing needs network access to alphacephei.com; not available in this build")


def list_languages():
    raise RuntimeError("language listing needs network access to alphacephei.com; not available in this build")


def load_dictionary(path):
    """word -> phones, keeping the most probable pronunciation (vosk_tts/model.py:48-55)."""
    dic, probs = {}, {}
    with open(path, encoding="utf-8") as f:
        for line in f:
            items = line.split(maxsplit=2)
            if len(items) < 3:
                continue
            prob = float(items[1])
            if probs.get(items[0], 0) < prob:
                dic[items[0]] = items[2].strip()
                probs[items[0]] = prob
    return dic


def _reserve_from_env():
    """VTTS_RESERVE="tokens,frames[,batch]" (default "256,1024"; "0" = off): workspace reservation at load time, so that the
    first long sentence does not move buffers and invalidate the CUDA graphs captured for the short ones."""
    v = os.environ.get("VTTS_RESERVE", "256,1024")
    if v.strip() in ("", "0"):
        return None
    return tuple(int(x) for x in v.split(","))


class Model:
    def __init__(self, model_path=None, model_name=None, lang=None, device=0, precision=1, session=None):
        if model_path is None:
            model_path = self.get_model_path(model_name, lang)
        model_path = Path(model_path)
        logging.info(f"Loading model from {model_path}")
        self.config = json.load(open(model_path / "config.json"))
        self.dic = load_dictionary(model_path / "dictionary") if (model_path / "dictionary").exists() else {}
        self.tokenizer = None
        if (model_path / "bert" / "vocab.txt").exists() or str(self.config.get("model_type", "")).startswith("multistream"):
            raise ValueError("bert-conditioned / multistream models are not VITS2 graphs: not supported by this engine")
        if session is not None:
            self.onnx = session
            return
        cks = sorted(glob.glob(str(model_path / "G_*.pth")), key=lambda p: int(re.sub(r"\D", "", os.path.basename(p)) or 0))
        if (model_path / "model.pth").exists():
            cks.append(str(model_path / "model.pth"))
        if (model_path / "model.onnx").exists():
            # the deployed layout (vosk_tts/model.py:46): everything comes out of the graph.  Preferred over a checkpoint
            # lying next to it: model.onnx is what the reference itself would load, and it is not a pickle
            from . import onnx_weights as _onnx
            sr = int(self.config.get("audio", {}).get("sample_rate", 22050))
            cfg = _onnx.config_from_onnx(str(model_path / "model.onnx"), sampling_rate=sr)
            folded = _onnx.state_dict_from_onnx(str(model_path / "model.onnx"))
            self.onnx = VitsSession(state_dict=folded, cfg=cfg, device=device, precision=precision, reserve=_reserve_from_env())
            return
        if "model" in self.config and "data" in self.config:
            n_vocab = len(self.config.get("phoneme_id_map", {})) or 62
            cfg = _config.from_training_json(self.config, n_vocab=n_vocab)
        elif (model_path / "vits_config.json").exists():
            n_vocab = len(self.config.get("phoneme_id_map", {})) or 62
            cfg = _config.from_training_json(str(model_path / "vits_config.json"), n_vocab=n_vocab)
        else:
            cfg = _config.DEFAULT_CONFIG
        if not cks:
            raise FileNotFoundError("no weights in %s: expected model.onnx (deployed layout) or G_*.pth / model.pth" % model_path)
        folded = _weights.load_checkpoint(cks[-1])
        self.onnx = VitsSession(state_dict=folded, cfg=cfg, device=device, precision=precision, reserve=_reserve_from_env())

    def get_model_path(self, model_name, lang):
        for directory in MODEL_DIRS:
            if directory is None or not Path(directory).exists():
                continue
            for entry in os.listdir(directory):
                if (model_name is not None and entry == model_name) or \
                        (model_name is None and lang and re.match(r"vosk-model(-small)?-{}".format(lang), entry)):
                    return Path(directory, entry)
        raise FileNotFoundError("model %r (lang %r) not found in %s and cannot be downloaded (no network)"
                                % (model_name, lang, [str(d) for d in MODEL_DIRS if d]))
