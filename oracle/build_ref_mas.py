"""Builds the reference's OWN Monotonic Alignment Search (training/vits2/monotonic_align/core.pyx, Cython) from the source where
it lies under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  Test infrastructure
only: it pins oracle/mas_oracle.py and generates tests/golden/mas_*.npz (oracle/make_golden_mas.py).  Nothing of the reference
is copied into the repository: Cython reads the .pyx in place, the generated C file and the .so go to oracle/_ref/.

    python oracle/build_ref_mas.py        # needs /root/reference, cython, gcc
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/training/vits2/monotonic_align/core.pyx"
OUT = os.path.join(HERE, "_ref")


def build():
    if not os.path.exists(SRC):
        raise SystemExit("reference tree not present: " + SRC)
    os.makedirs(OUT, exist_ok=True)
    c_file = os.path.join(OUT, "ref_mas_core.c")
    # module name must match the PyInit symbol Cython derives from the output file name
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--module-name", "ref_mas_core", SRC, "-o", c_file])
    import numpy as np
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "ref_mas_core" + ext)
    inc = [sysconfig.get_paths()["include"], np.get_include()]
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fopenmp", c_file, "-o", so] + ["-I" + i for i in inc]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build())
