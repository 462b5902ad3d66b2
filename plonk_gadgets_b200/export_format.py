"""Reader of the composer export written by pg_export_composer (include/pg_b200.h; layout: csrc/engine.hpp export_composer).

The file is the import adapter's input: a host that owns a real dusk-plonk StandardComposer replays it
(bindings/rust/plonk-gadgets-b200/src/import.rs).  This reader is what the tests use to replay it into the CPU oracle's composer."""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

MAGIC = b"PGB2EXP1"
KINDS = ("preamble", "add_input", "range_check", "max_bound", "maybe_equal", "is_non_zero", "is_non_zero_partial", "select_zero",
         "select_one", "constrain", "range_gate")          # GadgetKind of csrc/templates.hpp


@dataclass
class Call:
    kind: str
    num_bits: int
    n_inst: int
    base_row: int
    base_var: int
    rows_per_inst: int
    vars_per_inst: int
    operand_first_var: int
    operand_stride: int


@dataclass
class Export:
    n_rows: int
    n_vars: int
    calls: list
    variables: np.ndarray      # (n_vars, 32) uint8: BlsScalar::to_bytes
    w_idx: np.ndarray          # (4, n_rows) uint64: w_l, w_r, w_o, w_4
    sel: np.ndarray            # (8, n_rows, 32) uint8: q_m q_l q_r q_o q_4 q_c q_arith q_range, canonical bytes
    pi: np.ndarray             # (n_rows, 32) uint8: dense public inputs
    sigma: np.ndarray | None   # (4, n_rows) uint64 or None


def read(path: str) -> Export:
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("not a composer export")
        ver, flags = struct.unpack("<II", f.read(8))
        if ver != 1:
            raise ValueError(f"unknown export version {ver}")
        n_rows, n_vars, n_calls, chunk_rows, _, _ = struct.unpack("<6Q", f.read(48))
        calls = []
        for _ in range(n_calls):
            e = struct.unpack("<8Q", f.read(64))
            calls.append(Call(KINDS[e[0] & 0xFFFFFFFF], e[0] >> 32, e[1], e[2], e[3], e[4] & 0xFFFFFFFF, e[4] >> 32, e[5], e[6]))
        variables = np.frombuffer(f.read(32 * n_vars), dtype=np.uint8).reshape(n_vars, 32)
        w_idx = np.empty((4, n_rows), dtype=np.uint64)
        sel = np.empty((8, n_rows, 32), dtype=np.uint8)
        pi = np.empty((n_rows, 32), dtype=np.uint8)
        sigma = np.empty((4, n_rows), dtype=np.uint64) if flags & 1 else None
        done = 0
        while done < n_rows:
            r0, cnt = struct.unpack("<2Q", f.read(16))
            if r0 != done or cnt == 0 or cnt > chunk_rows:
                raise ValueError("row chunks out of order")
            w_idx[:, r0:r0 + cnt] = np.frombuffer(f.read(32 * cnt), dtype=np.uint64).reshape(4, cnt)
            sel[:, r0:r0 + cnt] = np.frombuffer(f.read(8 * 32 * cnt), dtype=np.uint8).reshape(8, cnt, 32)
            pi[r0:r0 + cnt] = np.frombuffer(f.read(32 * cnt), dtype=np.uint8).reshape(cnt, 32)
            if sigma is not None:
                sigma[:, r0:r0 + cnt] = np.frombuffer(f.read(32 * cnt), dtype=np.uint64).reshape(4, cnt)
            done += cnt
        if f.read(1):
            raise ValueError("trailing bytes")
    return Export(n_rows, n_vars, calls, variables, w_idx, sel, pi, sigma)
