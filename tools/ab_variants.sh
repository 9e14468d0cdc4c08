#!/bin/bash
# A/B of alternative builds of libvtts.so (VTTS_LIB): batch-1 bench value, batch-64, isolated conv_tc launches.
for so in "$@"; do
  echo "=== $so"
  export VTTS_LIB=$PWD/$so
  timeout 300 python bench.py --steps 40 --warmup 5 --cpu-steps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch1 ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline'])"
  timeout 300 python tools/bench_batch.py 64 5 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch64 ms', d['ms_per_step'], 'tc', d['conv_tc'])"
  timeout 300 python tools/microbench.py tc:192:384:5:1:162 tc:256:256:11:1:648 tc:128:128:11:1:2592 tc:128:128:11:5:41472 tc:512:512:7:1:10368
done
