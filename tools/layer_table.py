#!/usr/bin/env python
"""bench JSON line -> markdown per-layer roofline table (profiles/)."""
import json
import sys

d = json.load(open(sys.argv[1]))
r = d["roofline"]
print(f"# per-layer roofline, {d['config']['workload']}\n")
print(f"step {d['ms_per_step']:.3f} ms = {d['value']:.1f} {d['unit']}; sum of per-layer rooflines {r.get('step_roofline_ms', 0):.3f} ms "
      f"({100 * r.get('step_frac_of_roofline', 0):.0f} % of the event-timed kernel time per step); MMA work factor {r['mma_work_factor']}\n")
print("frac = roofline ms (slower of algorithmic FLOPs / measured bf16 peak and algorithmic bytes / measured HBM GB/s) / measured ms; "
      "'frac of issued' counts the MMA slots the fp32 mode issues per MAC (3, or 2 with 8-bit cross operands).\n")
print("| stage | ms | TFLOP/s (algorithmic) | GB/s (algorithmic) | bound | roofline ms | frac | MMA slots / MAC | frac of issued |")
print("|---|---|---|---|---|---|---|---|---|")
for k, v in r["per_layer"].items():
    print(f"| {k} | {v['ms']:.4f} | {v['tflops']:.1f} | {v['gbs']:.0f} | {v.get('bound', '')} | {v.get('roofline_ms', 0):.4f} | "
          f"{v.get('frac_of_roofline', 0):.2f} | {v.get('mma_slots_per_mac') or ''} | {v.get('frac_of_issued_roofline', 0):.2f} |")
