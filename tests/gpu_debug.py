"""TEST INFRASTRUCTURE (not collected by pytest): stage-by-stage comparison of the CUDA engine against the oracle on the
golden cases, for debugging on the GPU box:  python tests/gpu_debug.py t128_sid2 t17_sid2"""
import os, sys, time
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine
from oracle import vits_oracle as vo

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()), float(np.abs(b).max())

def main():
    cfg = C.DEFAULT_CONFIG
    sd = synthetic.make_random_checkpoint(cfg, 1234)
    w = weights.fold_weight_norm(sd)
    t0 = time.time(); blob, man = weights.pack(w, cfg); print("pack %.2fs" % (time.time() - t0))
    eng = Engine(cfg, blob, man, device=0, precision=int(os.environ.get('PRECISION', '0')))
    eng.debug_flags(1)
    cases = sys.argv[1:] or ["t17_sid2", "t128_sid2", "t1_single"]
    for name in cases:
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        print("==== case", name, "checksum ok" if abs(sum(v.double().sum().item() for k, v in sorted(sd.items())) - float(g["weight_checksum"])) < 1e-6 else "WEIGHT CHECKSUM MISMATCH")
        for u in range(int(g["n"])):
            p = "u%d_" % u
            tok = g[p + "tokens"]; T = len(tok); sid = int(g[p + "sid"]); scales = g["scales"]
            eps_dp = g[p + "eps_dp"]; eps_z = g[p + "eps_z"]
            with torch.no_grad():
                o = vo.infer(w, cfg, torch.as_tensor(tok)[None], torch.tensor([T]), torch.tensor([sid]), scales,
                             torch.as_tensor(eps_dp)[None], torch.as_tensor(eps_z)[None], return_all=True)
            ylen, dur = eng.durations(tok[None], [T], [sid], scales, eps_dp[None], want_durations=True)
            print(" T_x", T, "T_y engine", int(ylen[0]), "golden", int(g[p + "y_length"]), "oracle", int(o["y_lengths"][0]))
            x = eng.debug_read("x").reshape(T, -1)
            print("  x      ", rel(x, o["x"][0].T.numpy()))
            st = eng.debug_read("stats").reshape(T, -1)
            print("  m_p    ", rel(st[:, :192], o["m_p"][0].T.numpy()), " logs_p", rel(st[:, 192:], o["logs_p"][0].T.numpy()))
            print("  w_ceil equal:", bool((dur[0] == g[p + "w_ceil"]).all()), "n diff", int((dur[0] != g[p + "w_ceil"]).sum()))
            wq = (torch.exp(o["logw"]) * float(scales[1]))[0, 0].numpy()
            print("  min |w-round(w)| margin", float(np.abs(wq - np.round(wq)).min()))
            if int(ylen[0]) != int(g[p + "y_length"]):
                print("  !! length mismatch; skipping phase 2"); continue
            Ty = int(ylen[0])
            wav, idx = eng.synthesize(ylen, eps_z[None], want_alignment=True)
            print("  idx equal:", bool((idx[0, :Ty] == g[p + "idx"]).all()))
            zp = eng.debug_read("z_p").reshape(Ty, -1)
            print("  z_p    ", rel(zp, g[p + "z_p"].T))
            z = eng.debug_read("z").reshape(Ty, -1)
            print("  z      ", rel(z, g[p + "z"].T))
            # decoder intermediates from the oracle
            with torch.no_grad():
                zin = torch.as_tensor(g[p + "z"])[None]
                d0 = vo.conv(zin, w, "dec.conv_pre", padding=3)
                xs = d0
                stages = []
                nk = 3
                for i, (uu, ku) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
                    xs = F.leaky_relu(xs, 0.1)
                    xs = F.conv_transpose1d(xs, w["dec.ups.%d.weight" % i], w["dec.ups.%d.bias" % i], stride=uu, padding=(ku - uu) // 2)
                    up = xs
                    acc = None
                    for j in range(nk):
                        r = vo.resblock1(xs, w, "dec.resblocks.%d" % (i * nk + j), cfg["resblock_kernel_sizes"][j], cfg["resblock_dilation_sizes"][j])
                        acc = r if acc is None else acc + r
                    xs = acc / nk
                    stages.append(xs)
                xp = F.pad(F.leaky_relu(xs), (1, 0), mode="reflect")
                post = vo.conv(xp, w, "dec.subband_conv_post", padding=3)
            e_d0 = eng.debug_read("d0").reshape(Ty, -1)
            print("  d0     ", rel(e_d0, d0[0].T.numpy()))
            for i, s in enumerate(stages):
                e = eng.debug_read("stage%d" % i).reshape(s.shape[2], s.shape[1])
                print("  stage%d " % i, rel(e, s[0].T.numpy()))
            e_post = eng.debug_read("post").reshape(-1, 72)
            print("  post   ", rel(e_post[: post.shape[2]], post[0].T.numpy()))
            print("  wav vs golden", rel(wav[0, : Ty * 256], g[p + "wav"]), " vs oracle", rel(wav[0, : Ty * 256], o["o"][0, 0].numpy()))
            print("  stage ms", eng.stage_timings(), "launches", eng.kernel_launches())
    # quick timing of the headline case
    g = np.load(os.path.join(ROOT, "tests", "golden", "t128_sid2.npz"))
    tok = g["u0_tokens"]; T = len(tok)
    for it in range(8):
        if it == 5:
            eng.debug_flags(0)
        t0 = time.perf_counter()
        wav, yl = eng.infer(tok[None], [T], [2], g["scales"], g["u0_eps_dp"][None], g["u0_eps_z"][None])
        dt = time.perf_counter() - t0
        print("e2e infer %.3f ms, samples %d, %s replays=%d" % (dt * 1e3, int(yl[0]) * 256, eng.stage_timings(), eng.graph_replays()))
    ref = g["u0_wav"]
    print("graph-mode wav vs golden", float(np.abs(wav[0, :len(ref)] - ref).max()))

if __name__ == "__main__":
    main()
