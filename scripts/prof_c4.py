"""C4 alone (2^24 x (maybe_equal + is_non_zero)), a few steps: the workload for an ncu launch list of the inversion path."""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import plonk_gadgets_b200 as pg

SEED = 0x706C6F6E6B5F6732
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
mode = pg.CHECK_SPARSE if "--sparse" in sys.argv else pg.CHECK_GENERIC
c = pg.StandardComposer(device=0, check_mode=mode, timing=True, stream=stream.cuda_stream)
n = 1 << log2n
a = torch.empty((n, 4), dtype=torch.int64, device=dev); b = torch.empty_like(a)
c.synth(SEED, 41, 0, 0, a); c.synth(SEED, 42, 0, 0, b); b[0::2] = a[0::2]
for _ in range(3):
    c.reset(); va = c.add_input(a); vb = c.add_input(b)
    pg.maybe_equal(c, va, vb); pg.is_non_zero(c, va, a)
    bad, _ = c.check_circuit_satisfied(); assert bad == 0
torch.cuda.synchronize(dev)
print("ok", c.timing(reset=True))
