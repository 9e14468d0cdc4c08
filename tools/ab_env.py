"""Same-box comparison of tuning switches: the configs[1] step (device-resident, CUDA events, L2 flush) under each environment
given on the command line, e.g.  python tools/ab_env.py "" "VTTS_CONV_AUTOG=0" "VTTS_TC_BN=64"  (two rounds, alternating)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ab_value import CHILD  # noqa: E402


def main():
    sets = sys.argv[1:] or [""]
    for r in range(2):
        for s in sets:
            env = dict(os.environ)
            env.pop("VTTS_LIB", None)
            for kv in s.split():
                k, v = kv.split("=", 1)
                env[k] = v
            out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
            print("%-40s %s" % (s or "(default)", (out.stdout.strip().splitlines() or [out.stderr[-300:]])[-1]))


if __name__ == "__main__":
    main()
