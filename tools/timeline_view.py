"""Per-kernel view of one or two `VERBOSE=1 python tools/timeline.py` dumps (internal stamps folded into their kernel).

  VTTS_PDL=0 VERBOSE=1 python tools/timeline.py > a.txt      # entry stamps overlap under PDL: take the view without it
  python tools/timeline_view.py a.txt [b.txt] [first_kernel [last_kernel]]
"""
import re
import sys


def load(fn):
    rows = []
    for line in open(fn):
        m = re.match(r"\s*(\d+)\s+([\d.]+) us\s+\+\s*([\d.]+)\s+(\S+)", line)
        if m:
            rows.append((float(m.group(2)), float(m.group(3)), m.group(4)))
    kernels = []
    for t, d, name in rows:
        if name.startswith(".stamp") and kernels:
            kernels[-1][2] += d
            kernels[-1][3].append((name.strip(". "), d))
        else:
            kernels.append([t, name, d, []])
    return kernels


def main():
    files = [a for a in sys.argv[1:] if not a.isdigit()]
    nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
    a = load(files[0])
    b = load(files[1]) if len(files) > 1 else None
    lo = nums[0] if nums else 0
    hi = nums[1] if len(nums) > 1 else len(a)
    for i, k in enumerate(a):
        if not lo <= i < hi:
            continue
        extra = " ".join("%s=%.1f" % (n, d) for n, d in k[3])
        other = "  | %6.1f %s" % (b[i][2], b[i][1] if b[i][1] != k[1] else "") if b and i < len(b) else ""
        print("%4d %8.1f  %-24s %6.1f%s   %s" % (i, k[0], k[1], k[2], other, extra))
    tot = {}
    for k in a:
        tot[k[1]] = tot.get(k[1], 0.0) + k[2]
    print("total %.1f us over %d kernels" % (a[-1][0] + a[-1][2], len(a)))
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("  %-26s %8.1f" % (n, v))


if __name__ == "__main__":
    main()
