"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/make_golden.py

For every case the reference ``SynthesizerTrn.infer`` (training/vits2/models.py:1679-1704, built
as onnx_export.py:47-55,78-79 builds it) is run on CPU with the synthetic checkpoint
``vosk_tts_b200.synthetic.make_random_checkpoint(cfg, seed)`` loaded through ``load_state_dict``
and the two RNG draws (models.py:96, :1700) replaced by seeded tensors that are stored in the
fixture.  Weights are NOT stored (127 MB); they are regenerated from the seed, and a float64
checksum of the regenerated tensors is stored to detect generator drift.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from vosk_tts_b200 import config as C, synthetic  # noqa: E402

WEIGHT_SEED = 1234
OUT = os.path.join(ROOT, "tests", "golden")

# name, T_x list (B=1 runs each), sid list, scales, input seed
CASES = [
    ("t17_sid2", [17], [2], [0.8, 1.0, 0.8], 17),
    ("t128_sid2", [128], [2], [0.8, 1.0, 0.8], 0),          # BASELINE.json configs[0]/[1]
    ("t50_slow", [50], [7], [0.667, 1.3, 0.5], 50),
    ("t33_nonoise", [33], [0], [0.0, 1.0, 0.0], 33),
    ("ragged3", [9, 40, 23], [1, 2, 199], [0.8, 1.0, 0.8], 3),
    ("t1_single", [1], [5], [0.8, 1.0, 0.8], 1),
]


# Architecture variants: (fixture name, overrides of the training json's "model" block, T_x, sid, scales, input seed).
# "plainflow": use_transformer_flows=False with a transformer_flow_type other than the "mono_layer_post_residual" default
# selects the final else-branch of ResidualCouplingTransformersBlock.__init__ (models.py:735-747): plain
# modules.ResidualCouplingLayer + Flip, i.e. exactly ResidualCouplingBlock (models.py:765-810).
VARIANTS = [
    ("plainflow_t40", {"use_transformer_flows": False, "transformer_flow_type": "pre_conv2"}, 40, 3, [0.8, 1.1, 0.8], 40),
]


def weight_checksum(sd):
    return float(sum(v.double().sum().item() for k, v in sorted(sd.items())))


def main():
    torch.set_num_threads(1)   # fixed thread count: summation order of the fixture is reproducible
    cfg = C.from_training_json(rh.REF_CONFIG)
    sd = synthetic.make_random_checkpoint(cfg, WEIGHT_SEED)
    net = rh.build_reference_model(sd)
    os.makedirs(OUT, exist_ok=True)
    for name, txs, sids, scales, seed in CASES:
        g = torch.Generator().manual_seed(seed)
        out = {"weight_seed": WEIGHT_SEED, "weight_checksum": weight_checksum(sd),
               "scales": np.asarray(scales, np.float32), "n": len(txs)}
        for u, (T, sid) in enumerate(zip(txs, sids)):
            tok = torch.randint(0, cfg["n_vocab"], (1, T), generator=g)
            eps_dp = torch.randn(1, 2, T, generator=g)
            eps_z_full = torch.randn(1, cfg["inter_channels"], 24 * T + 8, generator=g)
            r = rh.reference_infer(net, tok, torch.tensor([T]), torch.tensor([sid]), scales, eps_dp,
                                   lambda s: eps_z_full[:, :, :s[2]])
            Ty = r["o"].shape[-1] // 256
            attn = r["attn"][0, 0]                      # [T_y, T_x] 0/1
            w_ceil = attn.sum(0)                        # frames per token (== w_ceil when y_len = sum)
            idx = attn.argmax(1)
            pre = "u%d_" % u
            out[pre + "tokens"] = tok[0].numpy().astype(np.int64)
            out[pre + "sid"] = np.int64(sid)
            out[pre + "eps_dp"] = eps_dp[0].numpy()
            out[pre + "eps_z"] = eps_z_full[0, :, :Ty].numpy().copy()
            out[pre + "w_ceil"] = w_ceil.numpy().astype(np.int32)
            out[pre + "idx"] = idx.numpy().astype(np.int32)
            out[pre + "y_length"] = np.int64(Ty)
            out[pre + "wav"] = r["o"][0, 0].numpy()
            out[pre + "z_p"] = r["z_p"][0].numpy()
            out[pre + "z"] = r["z"][0].numpy()
            out[pre + "o_mb"] = r["o_mb"][0].numpy()
            print(name, u, "T_x", T, "T_y", Ty, "wav absmax %.3f" % float(r["o"].abs().max()))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("weights checksum", weight_checksum(sd))
    import copy
    for name, over, T, sid, scales, seed in VARIANTS:
        jc = copy.deepcopy(rh.load_ref_config())
        jc["model"].update(over)
        vcfg = C.from_training_json(jc)
        vsd = synthetic.make_random_checkpoint(vcfg, WEIGHT_SEED)
        vnet = rh.build_reference_model(vsd, cfg=jc)
        g = torch.Generator().manual_seed(seed)
        tok = torch.randint(0, vcfg["n_vocab"], (1, T), generator=g)
        eps_dp = torch.randn(1, 2, T, generator=g)
        eps_z_full = torch.randn(1, vcfg["inter_channels"], 24 * T + 8, generator=g)
        r = rh.reference_infer(vnet, tok, torch.tensor([T]), torch.tensor([sid]), scales, eps_dp, lambda s: eps_z_full[:, :, :s[2]])
        Ty = r["o"].shape[-1] // 256
        attn = r["attn"][0, 0]
        out = {"weight_seed": WEIGHT_SEED, "weight_checksum": weight_checksum(vsd), "scales": np.asarray(scales, np.float32), "n": 1,
               "model_overrides": np.asarray(json.dumps(over)),
               "u0_tokens": tok[0].numpy().astype(np.int64), "u0_sid": np.int64(sid), "u0_eps_dp": eps_dp[0].numpy(),
               "u0_eps_z": eps_z_full[0, :, :Ty].numpy().copy(), "u0_w_ceil": attn.sum(0).numpy().astype(np.int32),
               "u0_idx": attn.argmax(1).numpy().astype(np.int32), "u0_y_length": np.int64(Ty), "u0_wav": r["o"][0, 0].numpy(),
               "u0_z_p": r["z_p"][0].numpy(), "u0_z": r["z"][0].numpy(), "u0_o_mb": r["o_mb"][0].numpy()}
        print(name, "T_x", T, "T_y", Ty, "wav absmax %.3f" % float(r["o"].abs().max()),
              "|z - z_p| max %.3f (the flow must not be an identity)" % float((r["z"] - r["z_p"]).abs().max()))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    main()
