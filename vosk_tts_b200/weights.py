"""Checkpoint -> engine weights: weight-norm folding and the packed device blob.

Contract (SURVEY.md section 5 "Checkpoint / resume"; /root/reference/training/vits2/utils.py:18-50,
onnx_export.py:55,78-79): a ``G_*.pth`` holds ``{'model': state_dict, ...}`` with weight-normed convs
stored as ``weight_g`` / ``weight_v``; the exporter loads it and removes weight norm on ``dec`` and
``flow`` before tracing.  ``fold_weight_norm`` performs the same fold; ``pack`` (below) lays every
tensor out the way the CUDA kernels consume it.
"""
import numpy as np
import torch


def fold_weight_norm(sd):
    """w = g * v / ||v||, the norm taken over all dims but 0 (torch.nn.utils.weight_norm dim=0;
    for ConvTranspose1d dim 0 is the *input* channel, e.g. dec.ups.0.weight_g is [512,1,1])."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_v"):
            base = k[: -len("_v")]
            g = sd[base + "_g"]
            t = v.float()
            nrm = t.reshape(t.shape[0], -1).norm(dim=1).reshape([-1] + [1] * (t.dim() - 1))
            out[base] = (t * (g.float() / nrm)).contiguous()
        elif k.endswith(".weight_g"):
            continue
        else:
            out[k] = v.float().contiguous() if torch.is_floating_point(v) else v
    return out


def load_checkpoint(path):
    """Reads a reference ``G_*.pth`` (utils.py:18-21) and returns the folded state dict."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    sd = ck["model"] if isinstance(ck, dict) and "model" in ck else ck
    sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
    return fold_weight_norm(sd)
