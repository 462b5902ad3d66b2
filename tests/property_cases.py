"""Random gadget compositions shared by the property tests: the host-backend ones (tests/test_emu_properties.py, run on CPU) and
their twins on the CUDA library (tests/test_gpu_properties.py).  A program is the list format of tests/programs.py."""
from __future__ import annotations

import numpy as np

from tests.programs import Q, hx

OPS = ["maybe_equal", "select_zero", "select_one", "is_non_zero", "max_bound", "constrain"]
OPS_WITH_RANGE_GATE = OPS + ["range_gate"]


def small_composition(n, vals, seq, seed):
    """Two input columns of n <= 5 values; later calls consume the columns earlier calls produced."""
    rng = np.random.default_rng(seed)
    prog = [dict(op="add_input", values=[hx(v) for v in vals[:n]]), dict(op="add_input", values=[hx(v) for v in vals[5:5 + n]])]
    cols = [0, 1]                                  # program indices that returned a column
    for op in seq:
        a, b = int(rng.choice(cols)), int(rng.choice(cols))
        if op == "maybe_equal":
            prog.append(dict(op=op, a=a, b=b)); cols.append(len(prog) - 1)
        elif op in ("select_zero", "select_one"):
            prog.append(dict(op=op, **({"x": a} if op == "select_zero" else {"y": a}), select=b)); cols.append(len(prog) - 1)
        elif op == "is_non_zero":
            prog.append(dict(op=op, var=a, assigned=[hx(v) for v in vals[:n]]))
        elif op == "max_bound":
            prog.append(dict(op=op, max=hx(2 ** int(rng.integers(1, 250))), witness=a)); cols.append(len(prog) - 1)
        elif op == "range_gate":
            prog.append(dict(op=op, witness=a, num_bits=2 * int(rng.integers(1, 129))))
        else:
            prog.append(dict(op="constrain_to_constant", a=a, constant=hx(int(rng.integers(0, 3))), pi=[hx(v) for v in vals[:n]] if seed % 2 else None))
    return prog


def batch_composition(vals, seq, seed, n=48):
    """Batches of n instances (48: large enough for the compiled structure-aware row program), satisfied or not."""
    rng = np.random.default_rng(seed)
    col_a = [vals[i % 8] if i % 3 else (vals[i % 8] + i) % Q for i in range(n)]
    col_b = [vals[(i + 3) % 8] if i % 2 else col_a[i] for i in range(n)]
    prog = [dict(op="add_input", values=[hx(v) for v in col_a]), dict(op="add_input", values=[hx(v) for v in col_b])]
    cols = [0, 1]
    for op in seq:
        a, b = int(rng.choice(cols)), int(rng.choice(cols))
        if op == "maybe_equal":
            prog.append(dict(op=op, a=a, b=b)); cols.append(len(prog) - 1)
        elif op in ("select_zero", "select_one"):
            prog.append(dict(op=op, **({"x": a} if op == "select_zero" else {"y": a}), select=b)); cols.append(len(prog) - 1)
        elif op == "is_non_zero":
            prog.append(dict(op=op, var=a, assigned=[hx(v if v else 1) for v in col_a]))
        elif op == "max_bound":
            bits = int(rng.integers(1, 250))      # per-instance bounds: max - 1 in [2^bits, 2^(bits+1)) for every instance (same num_bits)
            prog.append(dict(op=op, max=[hx(2 ** bits + 1 + int(rng.integers(0, 2 ** min(bits, 60)))) for _ in range(n)] if seed % 3 == 0 else hx(2 ** bits), witness=a))
            cols.append(len(prog) - 1)
        elif op == "range_gate":
            prog.append(dict(op=op, witness=a, num_bits=2 * int(rng.integers(1, 129))))
        else:
            prog.append(dict(op="constrain_to_constant", a=a, constant=hx(int(rng.integers(0, 3))), pi=[hx(v) for v in col_b] if seed % 2 else None))
    return prog
