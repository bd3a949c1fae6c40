#!/bin/bash
# re-test of design switches after the elected-lane MMA issue (the earlier A/B runs were made while the issuer was the bottleneck)
run() { name=$1; shift; env "$@" python tools/quick_infer.py ${MODE:-fp32} 20 | grep "step\|conv6.3x3\|conv7.3x3\|conv8\|conv9" | awk '{printf "%s ", $(NF-1)}' > gpurun_out/sw_${MODE:-fp32}_$name.log 2>&1; }
MODE=fp32
run default A=1
run nopair NSM_NO_WIDE_PAIR=1
run nopair_nomc NSM_NO_WIDE_PAIR=1 NSM_NO_WIDE_MC=1
run nowide NSM_NO_WIDE=1
run rings11 NSM_UB_RINGS=11
run rings12 NSM_UB_RINGS=12
run rings21 NSM_UB_RINGS=21
run rings22 NSM_UB_RINGS=22
run bst2 NSM_UB_BSTAGES=2
run nodb NSM_UB_NO_DB=1
run default2 A=1
MODE=bf16
run default A=1
run nodb NSM_UB_NO_DB=1
run bst2 NSM_UB_BSTAGES=2
run bst4 NSM_UB_BSTAGES=4
run default2 A=1
