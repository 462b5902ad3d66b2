//! SOURCE ONLY (never compiled in this repository's environment).
//!
//! Import adapter: brings a batch built on the GPU into a real `dusk_plonk::StandardComposer`, so that the steps
//! the reference's tests take next -- `prover.mut_cs()` -> gadget calls -> `preprocess` -> `prove`
//! (ref:tests/range_gadgets_tests.rs:82-91, ref:tests/scalar_gadgets_tests.rs:151) -- run on it unchanged.
//!
//! `Variable(pub(crate) usize)` cannot be forged outside dusk-plonk, so the adapter REPLAYS the export
//! (`BatchComposer::export` = `pg_export_composer`; file layout: csrc/engine.hpp `export_composer`, reader mirrored from
//! plonk_gadgets_b200/export_format.py) through public composer methods only:
//!   * every Variable a call appended: `composer.add_input(BlsScalar::from_bytes(..))`, checking that dusk-plonk hands out the index
//!     the export says (the numbering is sequential, so the target must be a FRESH composer: 3 rows, 5 variables);
//!   * every arithmetic row: `composer.poly_gate(a, b, c, q_m, q_l, q_r, q_o, q_c, pi)` -- all rows the six gadgets emit have
//!     `w_4 = zero_var`, `q_4 = 0`, `q_arith = 1`, which is exactly what `poly_gate` pushes;
//!   * `range_gate` calls: `composer.range_gate(witness_i, num_bits)` natively (their rows carry `q_range`, which no public
//!     method can set) -- the accumulators dusk-plonk allocates get the indices the export lists.
//! The replay is serial (it is bounded by dusk-plonk's own `HashMap` / `Vec` pushes: that is the cost structure SURVEY.md 3 measures
//! for the reference path); what it buys is that witness generation and the satisfaction check already happened on the GPU.
//! The tests replay the same file into the CPU oracle's composer (tests/export_replay.py, tests/test_export.py).
use dusk_bytes::Serializable;
use dusk_plonk::prelude::*;
use std::convert::TryInto;
use std::io::{self, Read};

#[derive(Debug)]
pub enum ImportError {
    Io(io::Error),
    /// not an export file, or a version this adapter does not know
    Format(&'static str),
    /// the target composer is not fresh, or dusk-plonk numbered a Variable differently from the export
    Numbering { expected: u64, got: u64 },
    /// a row no public composer method can express (q_4 != 0 or a fourth wire outside range_gate)
    Unsupported(u64),
}
impl From<io::Error> for ImportError {
    fn from(e: io::Error) -> Self {
        ImportError::Io(e)
    }
}

const KIND_RANGE_GATE: u32 = 10; // GadgetKind::G_RANGE_GATE, csrc/templates.hpp

struct Call {
    kind: u32,
    num_bits: u32,
    n_inst: u64,
    base_row: u64,
    base_var: u64,
    rows_per: u64,
    vars_per: u64,
    op_first_var: u64,
    op_stride: u64,
}

fn u64s<R: Read>(r: &mut R, n: usize) -> io::Result<Vec<u64>> {
    let mut buf = vec![0u8; 8 * n];
    r.read_exact(&mut buf)?;
    Ok(buf.chunks(8).map(|c| u64::from_le_bytes(c.try_into().unwrap())).collect())
}
fn scalars<R: Read>(r: &mut R, n: usize) -> Result<Vec<BlsScalar>, ImportError> {
    let mut buf = vec![0u8; 32 * n];
    r.read_exact(&mut buf)?;
    buf.chunks(32)
        .map(|c| BlsScalar::from_bytes(c.try_into().unwrap()).map_err(|_| ImportError::Format("non-canonical scalar")))
        .collect()
}

/// Replays an export into `composer` (which must be fresh).  Returns the `Variable`s in export order: `vars[i]` is Variable i of the
/// batch, so a column `first + k*stride` of `BatchComposer` maps to `vars[first + k*stride]` for further reference-gadget calls.
pub fn import_into<R: Read>(composer: &mut StandardComposer, mut src: R) -> Result<Vec<Variable>, ImportError> {
    let mut magic = [0u8; 8];
    src.read_exact(&mut magic)?;
    if &magic != b"PGB2EXP1" {
        return Err(ImportError::Format("magic"));
    }
    let head = u64s(&mut src, 7)?; // version | flags << 32, n_rows, n_vars, n_calls, chunk_rows, 0, 0
    if head[0] as u32 != 1 {
        return Err(ImportError::Format("version"));
    }
    let has_sigma = (head[0] >> 32) & 1 == 1;
    let (n_rows, n_vars, n_calls) = (head[1], head[2], head[3]);
    let mut calls = Vec::with_capacity(n_calls as usize);
    for _ in 0..n_calls {
        let e = u64s(&mut src, 8)?;
        calls.push(Call {
            kind: e[0] as u32, num_bits: (e[0] >> 32) as u32, n_inst: e[1], base_row: e[2], base_var: e[3],
            rows_per: e[4] & 0xffff_ffff, vars_per: e[4] >> 32, op_first_var: e[5], op_stride: e[6],
        });
    }
    if composer.circuit_size() != 3 {
        return Err(ImportError::Numbering { expected: 3, got: composer.circuit_size() as u64 });
    }
    let values = scalars(&mut src, n_vars as usize)?;
    // rows arrive in chunks; they are few bytes each compared with the replay work, so gather them first
    let mut w_idx: [Vec<u64>; 4] = Default::default();
    let mut sel: [Vec<BlsScalar>; 8] = Default::default();
    let mut pi: Vec<BlsScalar> = Vec::with_capacity(n_rows as usize);
    let mut done = 0u64;
    while done < n_rows {
        let h = u64s(&mut src, 2)?;
        let cnt = h[1] as usize;
        if h[0] != done || cnt == 0 {
            return Err(ImportError::Format("row chunks out of order"));
        }
        for w in w_idx.iter_mut() {
            w.extend(u64s(&mut src, cnt)?);
        }
        for s in sel.iter_mut() {
            s.extend(scalars(&mut src, cnt)?);
        }
        pi.extend(scalars(&mut src, cnt)?);
        if has_sigma {
            u64s(&mut src, 4 * cnt)?; // dusk-plonk rebuilds the permutation itself from the wires (perm.add_variables_to_map)
        }
        done += cnt as u64;
    }
    // ---- replay, call by call (call 0 is the fresh composer itself: zero variable + the two dummy rows)
    // (`zero_var()` / `circuit_size()` are dusk-plonk 0.8 accessors recalled from memory, like everything marked [dusk-plonk] in this
    // repository: this file has never met a compiler.)
    let mut vars: Vec<Variable> = Vec::with_capacity(n_vars as usize);
    vars.push(composer.zero_var());
    // Variables 1..=4 of StandardComposer::new() (6, 1, 7, -20) are not reachable through a public accessor; they never appear on a
    // gadget row, so their slots hold the zero variable as a placeholder.
    for _ in 1..5 {
        vars.push(composer.zero_var());
    }
    let mut next_var = 5u64;
    for call in calls.iter().skip(1) {
        if call.base_var != next_var || call.base_row != composer.circuit_size() as u64 {
            return Err(ImportError::Numbering { expected: call.base_var, got: next_var });
        }
        let (nv, nr) = (call.n_inst * call.vars_per, call.n_inst * call.rows_per);
        if call.kind == KIND_RANGE_GATE {
            for i in 0..call.n_inst {
                composer.range_gate(vars[(call.op_first_var + i * call.op_stride) as usize], call.num_bits as usize);
            }
            // the accumulators were allocated by dusk-plonk itself; they are only ever used by their own gate, so no handle is kept
            for _ in 0..nv {
                vars.push(composer.zero_var());
            }
        } else {
            for v in call.base_var..call.base_var + nv {
                vars.push(composer.add_input(values[v as usize]));
            }
            for r in call.base_row..call.base_row + nr {
                let r = r as usize;
                if w_idx[3][r] != 0 || sel[4][r] != BlsScalar::zero() {
                    return Err(ImportError::Unsupported(r as u64));
                }
                let p = if pi[r] == BlsScalar::zero() { None } else { Some(pi[r]) };
                composer.poly_gate(vars[w_idx[0][r] as usize], vars[w_idx[1][r] as usize], vars[w_idx[2][r] as usize],
                                   sel[0][r], sel[1][r], sel[2][r], sel[3][r], sel[5][r], p);
            }
        }
        next_var += nv;
        if composer.circuit_size() as u64 != call.base_row + nr {
            return Err(ImportError::Numbering { expected: call.base_row + nr, got: composer.circuit_size() as u64 });
        }
    }
    Ok(vars)
}
