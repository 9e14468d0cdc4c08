#!/bin/bash
# batch-64 A/B over alternative builds (VTTS_LIB); "-" = the in-tree library
for so in "$@"; do
  ( [ "$so" != "-" ] && export VTTS_LIB=$PWD/$so VTTS_TC_SPLIT=1
    echo "== $so"; python tools/bench_batch.py 64 4 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['conv_tc'], d['stage_ms']['flow'], d['stage_ms']['decoder'])" )
done
