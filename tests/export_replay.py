"""Replay of a composer export (pg_export_composer) into the CPU oracle's composer -- what the Rust import adapter does with a real
dusk-plonk StandardComposer (bindings/rust/plonk-gadgets-b200/src/import.rs): call by call, add_input for the call's Variables (checking
the numbering) and poly_gate for its rows; range_gate calls are replayed natively and their accumulators compared."""
from __future__ import annotations

import numpy as np

from plonk_gadgets_b200 import export_format


def replay(path: str, ob):
    ex = export_format.read(path)
    oc = ob.Composer()
    assert ex.calls[0].kind == "preamble" and (ex.calls[0].base_row, ex.calls[0].base_var) == (0, 0)
    assert oc.n == 3 and oc.n_vars == 5                      # StandardComposer::new()
    fr = lambda raw: ob.from_ints([int.from_bytes(bytes(r), "little") for r in raw])
    for call in ex.calls[1:]:
        nv, nr = call.n_inst * call.vars_per_inst, call.n_inst * call.rows_per_inst
        assert (call.base_row, call.base_var) == (oc.n, oc.n_vars)
        if call.kind == "range_gate":                        # native replay: composer.range_gate(witness_i, num_bits)
            wit = call.operand_first_var + call.operand_stride * np.arange(call.n_inst, dtype=np.uint64)
            oc.range_gate_batch(wit, call.num_bits)
            assert (oc.n, oc.n_vars) == (call.base_row + nr, call.base_var + nv)
            continue
        if nv:
            ids = oc.add_input_batch(fr(ex.variables[call.base_var: call.base_var + nv]))
            assert ids[0] == call.base_var and ids[-1] == call.base_var + nv - 1
        if nr:
            rows = slice(call.base_row, call.base_row + nr)
            assert (ex.w_idx[3, rows] == 0).all()            # poly_gate: the fourth wire is the zero variable
            one = np.zeros(32, dtype=np.uint8); one[0] = 1
            assert (ex.sel[6, rows] == one).all() and not ex.sel[7, rows].any()      # q_arith = 1, q_range = 0
            sel6 = np.stack([fr(ex.sel[k, rows]) for k in range(6)])
            oc.poly_gate_batch(ex.w_idx[0, rows], ex.w_idx[1, rows], ex.w_idx[2, rows], sel6, fr(ex.pi[rows]))
    return ex, oc
