"""Times single conv kernels in isolation on the GPU (engine.microbench)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosk_tts_b200 import config as C, synthetic, weights
from vosk_tts_b200.engine import Engine

cfg = C.DEFAULT_CONFIG
blob, man = weights.pack(weights.fold_weight_norm(synthetic.make_random_checkpoint(cfg, 1234)), cfg)
eng = Engine(cfg, blob, man, device=0, precision=1)
specs = sys.argv[1:] or ["tc:192:192:1:1:162", "tc:192:384:5:1:162", "tc:256:256:11:1:648", "tc:128:128:11:1:2592",
                         "tc:128:128:11:5:41472", "ffma:192:192:1:1:162", "ffma:192:384:5:1:162", "ffma:256:256:11:1:648",
                         "ffma:128:128:11:1:2592", "ffma:128:128:11:5:41472"]
for sp in specs:
    ms = eng.microbench(sp, 50)
    kind, cin, cout, k, dil, rows = sp.split(":")
    fl = 2.0 * int(cin) * int(cout) * int(k) * int(rows)
    print("%-28s %8.2f us  %8.2f TFLOP/s" % (sp, ms * 1e3, fl / (ms * 1e-3) / 1e12))
