#!/bin/bash
# per-layer timing of the cfg1 step for several NSM_EW16_KB thresholds (longest K loop, in k-blocks, that gets sixteen epilogue warps)
for m in fp32 bf16; do
  for kb in default 0 4 9 18 36 72; do
    if [ $kb = default ]; then python tools/quick_infer.py $m 20 > gpurun_out/ew16_${m}_$kb.log 2>&1
    else NSM_EW16_KB=$kb python tools/quick_infer.py $m 20 > gpurun_out/ew16_${m}_$kb.log 2>&1; fi
  done
done
