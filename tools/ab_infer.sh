#!/bin/bash
# inference-only same-box A/B: committed HEAD (ab_old/, tools/make_ab_old.sh) vs working tree
for i in 1 2 3; do
  for m in ${AB_MODES:-fp32 bf16}; do
    python ab_old/tools/quick_infer.py $m 20 > gpurun_out/ab_old_${m}_$i.log 2>&1
    python tools/quick_infer.py $m 20 > gpurun_out/ab_new_${m}_$i.log 2>&1
  done
done
