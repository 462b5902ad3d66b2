"""Pipe micro-benchmarks on the current GPU (roofline denominators / instruction-cost evidence)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import plonk_gadgets_b200 as pg

c = pg.StandardComposer(device=0)
props = torch.cuda.get_device_properties(0)
out = {"gpu": props.name, "sms": props.multi_processor_count}
for mode, name in enumerate(c.MICROBENCH_MODES):
    v = c.microbench(mode)
    out[name] = {"ops_per_s": v, "per_sm_per_clk_at_1965MHz": v / props.multi_processor_count / 1.965e9}
print(json.dumps(out, indent=1))
