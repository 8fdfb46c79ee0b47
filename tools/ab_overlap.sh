#!/bin/bash
# Same-box A/B of the encoder-stream overlap (DS2_ENC_OVERLAP=0/1), alternated: tools/ab_overlap.sh <tag>
tag=${1:-rX}
o=gpurun_out
for rep in 1 2 3; do
  for ov in 0 1; do
    echo "overlap=$ov rep=$rep: $(DS2_ENC_OVERLAP=$ov timeout 300 python bench.py --only-device --no-cpu-baseline --steps 60 2>/dev/null | cut -c1-200)"
  done
done | tee $o/${tag}_overlap_ab_60steps.txt
for ov in 0 1; do
  echo "stream overlap=$ov: $(DS2_ENC_OVERLAP=$ov timeout 300 python bench.py --mode stream --frames 300 2>/dev/null | cut -c1-300)"
done | tee -a $o/${tag}_overlap_ab_60steps.txt
