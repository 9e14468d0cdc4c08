#!/bin/bash
# A/B of runtime knobs: each argument is an env assignment list ("VTTS_TC_SPLIT=1"), "-" = defaults.
for cfg in "$@"; do
  echo "=== $cfg"
  ( [ "$cfg" != "-" ] && export $cfg
    timeout 300 python bench.py --steps 40 --warmup 5 --cpu-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('batch1 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'tc TF', r['achieved'], 'share', r['share_of_step'])"
    timeout 300 python tools/microbench.py tc:192:384:5:1:162 tc:768:192:3:1:162 tc:256:256:11:1:648 tc:128:128:11:1:2592 )
done
