"""TEST INFRASTRUCTURE (not collected by pytest): statistical parity sweep on the GPU box.

  python tests/parity_sweep.py [n_utterances=200] [precision=1]

For n random utterances (8..160 phonemes, random speaker / scales / noise) the CUDA engine and the oracle run on the same
inputs; reported: how many utterances have a different frame count or any different ceil(duration) (the oracle computes
w = exp(logw)*length_scale in fp32; a flip needs w within the engine's ~2e-6 error of an integer), the
smallest |w - round(w)| margin seen, and the waveform max-abs error distribution over the utterances whose durations agree.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vits_oracle as vo  # noqa: E402
from vosk_tts_b200 import config as C, synthetic, weights  # noqa: E402
from vosk_tts_b200.engine import Engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    precision = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    cfg = C.DEFAULT_CONFIG
    w = weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234))
    blob, man = weights.pack(w, cfg)
    eng = Engine(cfg, blob, man, device=0, precision=precision)
    g = torch.Generator().manual_seed(2024)
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    flips, len_diff, margins, errs = 0, 0, [], []
    for i in range(n):
        T = int(torch.randint(8, 161, (1,), generator=g))
        tok = torch.randint(0, cfg["n_vocab"], (1, T), generator=g)
        sid = int(torch.randint(0, cfg["n_speakers"], (1,), generator=g))
        scales = [float(torch.rand(1, generator=g)), 0.7 + 0.8 * float(torch.rand(1, generator=g)), float(torch.rand(1, generator=g))]
        eps_dp = torch.randn(1, 2, T, generator=g)
        eps_z = torch.randn(1, cfg["inter_channels"], 40 * T + 8, generator=g)
        with torch.no_grad():
            o = vo.infer(w, cfg, tok, torch.tensor([T]), torch.tensor([sid]), scales, eps_dp, eps_z, return_all=True)
        wq = (torch.exp(o["logw"]) * scales[1])[0, 0].numpy()
        margins.append(float(np.abs(wq - np.round(wq)).min()))
        ylen, dur = eng.durations(tok.numpy(), [T], [sid], scales, eps_dp.numpy(), want_durations=True)
        same = np.array_equal(dur[0], o["w_ceil"][0, 0].numpy().astype(np.int32))
        if not same:
            flips += 1
            len_diff += int(int(ylen[0]) != int(o["y_lengths"][0]))
            print("utterance %d: T_x %d, %d tokens with a different ceil, margin %.2e" % (i, T, int((dur[0] != o["w_ceil"][0, 0].numpy()).sum()), margins[-1]))
            eng.synthesize(ylen, eps_z[:, :, : int(ylen[0])].numpy())       # finish the call
            continue
        Ty = int(ylen[0])
        wav = eng.synthesize(ylen, eps_z[:, :, :Ty].numpy())
        errs.append(float(np.abs(wav[0, : Ty * eng.hop] - o["o"][0, 0].numpy()).max()))
    errs = np.array(errs) if errs else np.zeros(1)
    print("utterances %d  precision %d" % (n, precision))
    print("duration mismatches: %d (%d also change the frame count)   smallest integer margin seen %.3e" % (flips, len_diff, min(margins)))
    print("waveform max-abs error over %d matching utterances: median %.2e  p99 %.2e  max %.2e (budget 1e-3)"
          % (len(errs), float(np.median(errs)), float(np.percentile(errs, 99)), float(errs.max())))


if __name__ == "__main__":
    main()
