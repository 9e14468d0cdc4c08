#!/bin/bash
# per-kernel in-graph timelines under different knob settings -> gpurun_out/tl_<n>.txt
i=0
for cfg in "$@"; do
  ( [ "$cfg" != "-" ] && export $cfg
    echo "=== $cfg" > gpurun_out/tl_$i.txt
    VERBOSE=1 timeout 300 python tools/timeline.py >> gpurun_out/tl_$i.txt 2>&1
    grep -E "^total|^conv_tc_kernel" gpurun_out/tl_$i.txt | sed "s/^/[$cfg] /" )
  i=$((i+1))
done
