"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference PyTorch modules.

Only usable in the build container (``/root/reference`` does not exist on the GPU
box).  Used by ``oracle/make_golden.py`` and by the CPU tests that pin the
restatement in ``oracle/vits_oracle.py`` against the real reference.

The deployed ONNX graph is a trace of ``SynthesizerTrn.infer``
(/root/reference/training/vits2/models.py:1679-1704, exported by
training/vits2/onnx_export.py:47-104); onnxruntime / model.onnx are absent in this
image, so the PyTorch module is the reference implementation we can run.

Shims: ``librosa.util`` stub (stft.py:32-33 imports
``pad_center``/``tiny``), ``monotonic_align`` stub (models.py:10; never called by
``infer``), and ``text`` is not imported (text/__init__.py:30 opens a dictionary file).
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("VTTS_REFERENCE_ROOT", "/root/reference")
REF_VITS2 = os.path.join(REF_ROOT, "training", "vits2")
REF_CONFIG = os.path.join(REF_VITS2, "configs", "mb_istft_vits2_multi.json")


def available():
    return os.path.isfile(os.path.join(REF_VITS2, "models.py"))


def _install_shims():
    if "librosa" not in sys.modules:
        librosa = types.ModuleType("librosa")
        util = types.ModuleType("librosa.util")

        def pad_center(data, *, size, axis=-1, **kw):
            n = data.shape[axis]
            lpad = (size - n) // 2
            pads = [(0, 0)] * data.ndim
            pads[axis] = (lpad, size - n - lpad)
            return np.pad(data, pads)

        def tiny(x):
            return np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating) else np.float32).tiny

        def normalize(S, norm=None, axis=0, **kw):
            return S

        util.pad_center, util.tiny, util.normalize = pad_center, tiny, normalize
        librosa.util = util
        sys.modules["librosa"] = librosa
        sys.modules["librosa.util"] = util
    if "monotonic_align" not in sys.modules:
        ma = types.ModuleType("monotonic_align")

        def maximum_path(*a, **k):
            raise RuntimeError("monotonic_align is training-only (stubbed)")

        ma.maximum_path = maximum_path
        sys.modules["monotonic_align"] = ma


def import_reference():
    """Returns the reference ``models`` module (training/vits2/models.py), unmodified."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _install_shims()
    if REF_VITS2 not in sys.path:
        sys.path.insert(0, REF_VITS2)
    import models  # noqa: F401  (the reference's module)
    return models


def load_ref_config(path=None):
    with open(path or REF_CONFIG) as f:
        return json.load(f)


def build_reference_model(state_dict=None, cfg=None, n_vocab=62, quiet=True):
    """Build SynthesizerTrn exactly as the exporter does (onnx_export.py:47-55,78-79):
    is_onnx=True, optional checkpoint load, weight-norm removed on dec and flow, eval()."""
    models = import_reference()
    cfg = cfg or load_ref_config()
    out = io.StringIO()
    with contextlib.redirect_stdout(out if quiet else sys.stdout):
        torch.manual_seed(1234)
        net = models.SynthesizerTrn(
            n_vocab, 80, cfg["train"]["segment_size"] // cfg["data"]["hop_length"],
            n_speakers=cfg["data"]["n_speakers"], is_onnx=True, **cfg["model"])
        if state_dict is not None:
            missing, unexpected = net.load_state_dict(state_dict, strict=False)
            assert not unexpected, unexpected
        net.eval()
        net.dec.remove_weight_norm()
        try:
            net.flow.remove_weight_norm()
        except AttributeError:
            # use_transformer_flows=False: the block holds plain modules.ResidualCouplingLayer (models.py:747-757 ->
            # modules.py:298-343), which has no remove_weight_norm of its own -- the reference exporter would fail here
            # (onnx_export.py:79).  Fold the same tensors through the WN module's own method (modules.py:178-184).
            for i, l in enumerate(net.flow.flows):
                if i % 2 == 0:
                    l.enc.remove_weight_norm()
    return net


def reference_infer(net, tokens, lengths, sid, scales, eps_dp, eps_z_fn):
    """Run the reference ``infer`` with injected noise.

    eps_dp: [B,2,T_x]; eps_z_fn(shape)->tensor supplies the second draw (its shape
    [B,192,T_y] is only known after the duration predictor ran)."""
    state = {"n": 0}
    orig_randn, orig_randn_like = torch.randn, torch.randn_like

    def randn(*size, **kw):
        state["n"] += 1
        assert state["n"] == 1
        return eps_dp.clone()

    def randn_like(x, **kw):
        state["n"] += 1
        assert state["n"] == 2
        return eps_z_fn(tuple(x.shape)).clone()

    torch.randn, torch.randn_like = randn, randn_like
    try:
        with torch.no_grad():
            o, o_mb, attn, y_mask, (z, z_p, m_p, logs_p) = net.infer(
                tokens, lengths, sid=sid, noise_scale=float(scales[0]),
                length_scale=float(scales[1]), noise_scale_w=float(scales[2]))
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_randn_like
    return dict(o=o, o_mb=o_mb, attn=attn, y_mask=y_mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p)


def export_reference_onnx(path, net, n_vocab=62, opset=15):
    """Write ``model.onnx`` for ``net`` the way the reference does (training/vits2/onnx_export.py:60-104): ``forward`` is
    replaced by the infer wrapper, inputs ``input, input_lengths, scales, sid``, output ``output``, dynamic batch/time
    axes, legacy TorchScript exporter.  The ``onnx`` python package is absent in this image; the exporter only imports
    it for a post-processing step that attaches onnxscript functions (there are none here), which is bypassed."""
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    def infer_forward(text, text_lengths, scales, sid=None):
        return net.infer(text, text_lengths, noise_scale=scales[0], length_scale=scales[1], noise_scale_w=scales[2],
                         sid=sid)[0].unsqueeze(1)

    orig_fn, orig_forward = onnx_proto_utils._add_onnxscript_fn, net.forward
    onnx_proto_utils._add_onnxscript_fn = lambda proto, custom_opsets: proto
    net.forward = infer_forward
    try:
        g = torch.Generator().manual_seed(0)
        text = torch.randint(0, n_vocab, (1, 50), dtype=torch.long, generator=g)
        args = (text, torch.LongTensor([50]), torch.FloatTensor([0.667, 1.0, 0.8]), torch.LongTensor([0]))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.onnx.export(model=net, args=args, f=str(path), verbose=False, opset_version=opset,
                              input_names=["input", "input_lengths", "scales", "sid"], output_names=["output"],
                              dynamic_axes={"input": {0: "batch_size", 1: "phonemes"}, "input_lengths": {0: "batch_size"},
                                            "output": {0: "batch_size", 1: "time"}}, dynamo=False)
    finally:
        onnx_proto_utils._add_onnxscript_fn = orig_fn
        net.forward = orig_forward
    return path
