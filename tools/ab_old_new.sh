#!/bin/bash
# alternate old (ab_old/, HEAD before this session) and new library on the same box
for i in 1 2 3; do
  for m in fp32 bf16; do
    python ab_old/tools/quick_infer.py $m 20 > gpurun_out/ab_old_${m}_$i.log 2>&1
    python tools/quick_infer.py $m 20 > gpurun_out/ab_new_${m}_$i.log 2>&1
  done
done
for i in 1 2; do
  (cd ab_old && python bench.py --workload train --no-perturb --no-stock --no-cpu-baseline --steps 10 --warmup 3) > gpurun_out/ab_old_train_$i.log 2>&1
  python bench.py --workload train --no-perturb --no-stock --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/ab_new_train_$i.log 2>&1
done
