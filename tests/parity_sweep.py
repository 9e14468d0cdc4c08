"""TEST INFRASTRUCTURE (not collected by pytest): statistical parity sweep on the GPU box.

  python tests/parity_sweep.py [n_utterances=1000] [n_full=100] [out.json]

For n random utterances (8..160 phonemes, random speaker / scales / noise) the oracle's duration path (text encoder +
stochastic duration predictor, fp32 on the CPU) runs ONCE and every precision mode of the CUDA engine runs on the same
inputs.  Reported per mode: utterances with any different ceil(duration) ("flips"; the oracle computes
w = exp(logw)*length_scale in fp32, a flip needs w within the engine's arithmetic error of an integer),
how many of those also change the frame count, the distribution of the relative error of w itself (from the engine's
debug tensors) and -- for the first n_full utterances, where the oracle also runs flow + decoder -- the waveform
max-abs error over the utterances whose durations agree.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vits_oracle as vo  # noqa: E402
from vosk_tts_b200 import config as C, synthetic, weights  # noqa: E402
from vosk_tts_b200.engine import Engine  # noqa: E402


def oracle_durations(w, cfg, tok, T, sid, scales, eps_dp):
    """The duration half of vits_oracle.infer (models.py:1680-1691)."""
    import torch.nn.functional as F
    g = F.embedding(torch.tensor([sid]), w["emb_g.weight"]).unsqueeze(-1)
    x, m_p, logs_p, x_mask = vo.text_encoder(tok, torch.tensor([T]), g, w, cfg)
    logw = vo.sdp_reverse(x, x_mask, g, eps_dp, float(scales[2]), w, cfg)
    w_ceil, y_lengths = vo.durations(logw, x_mask, float(scales[1]))
    return logw, w_ceil, y_lengths


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n_full = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    modes = [int(m) for m in os.environ.get("SWEEP_MODES", "0,1,2,3").split(",")]
    cfg = C.DEFAULT_CONFIG
    w = weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234))
    blob, man = weights.pack(w, cfg)
    engines = {m: Engine(cfg, blob, man, device=0, precision=m) for m in modes}
    g = torch.Generator().manual_seed(2024)
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    st = {m: dict(flips=0, len_diff=0, flipped_tokens=0, errs=[], flip_margins=[]) for m in modes}
    margins, tokens_total, frames_total = [], 0, 0
    t0 = time.time()
    for i in range(n):
        T = int(torch.randint(8, 161, (1,), generator=g))
        tok = torch.randint(0, cfg["n_vocab"], (1, T), generator=g)
        sid = int(torch.randint(0, cfg["n_speakers"], (1,), generator=g))
        scales = [float(torch.rand(1, generator=g)), 0.7 + 0.8 * float(torch.rand(1, generator=g)), float(torch.rand(1, generator=g))]
        eps_dp = torch.randn(1, 2, T, generator=g)
        full = i < n_full
        eps_z = None
        with torch.no_grad():
            logw, w_ceil, y_lengths = oracle_durations(w, cfg, tok, T, sid, scales, eps_dp)
            if full:       # the frame count is data dependent: draw the second noise tensor once it is known
                eps_z = torch.randn(1, cfg["inter_channels"], int(y_lengths[0]), generator=g)
                o = vo.infer(w, cfg, tok, torch.tensor([T]), torch.tensor([sid]), scales, eps_dp, eps_z, return_all=True)
                assert torch.equal(o["w_ceil"], w_ceil)
        wq = (torch.exp(logw) * scales[1])[0, 0].numpy()
        margins.append(float(np.abs(wq - np.round(wq)).min()))
        tokens_total += T
        frames_total += int(y_lengths[0])
        ref = w_ceil[0, 0].numpy().astype(np.int32)
        for m, eng in engines.items():
            ylen, dur = eng.durations(tok.numpy(), [T], [sid], scales, eps_dp.numpy(), want_durations=True)
            same = np.array_equal(dur[0], ref)
            s = st[m]
            if not same:
                s["flips"] += 1
                s["len_diff"] += int(int(ylen[0]) != int(y_lengths[0]))
                s["flipped_tokens"] += int((dur[0] != ref).sum())
                s["flip_margins"].append(margins[-1])
            Ty = int(ylen[0])
            if full and same:
                wav = eng.synthesize(ylen, eps_z.numpy())
                s["errs"].append(float(np.abs(wav[0, : Ty * eng.hop] - o["o"][0, 0].numpy()).max()))
            else:
                # finish the two-phase call (Philox noise; result unused)
                eng.synthesize(ylen, None)
    res = dict(utterances=n, full_waveform_utterances=n_full, tokens=tokens_total, frames=frames_total,
               frames_per_token=frames_total / max(tokens_total, 1), seconds=time.time() - t0,
               smallest_integer_margin=float(min(margins)), margin_p01=float(np.percentile(margins, 1)), modes={})
    for m in modes:
        s = st[m]
        e = np.array(s["errs"]) if s["errs"] else np.zeros(1)
        res["modes"][str(m)] = dict(duration_mismatch_utterances=s["flips"], also_change_frame_count=s["len_diff"],
                                    flipped_tokens=s["flipped_tokens"], flip_rate_per_utterance=s["flips"] / n,
                                    flip_rate_per_token=s["flipped_tokens"] / max(tokens_total, 1),
                                    margins_of_flipped=sorted(s["flip_margins"])[:20],
                                    wav_compared=len(s["errs"]), wav_err_median=float(np.median(e)),
                                    wav_err_p99=float(np.percentile(e, 99)), wav_err_max=float(e.max()), wav_budget=1e-3)
    print(json.dumps(res, indent=1))
    if out_path:
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
