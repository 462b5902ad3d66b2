//! Runs the programs of tests/golden/programs.json through the reference's gadgets on a real dusk-plonk StandardComposer and writes
//! tests/golden/ref_dump.json (schema of programs.json's `expected`).  Needs the `pg_dump` accessor appended to dusk-plonk
//! (README.md).  SOURCE ONLY in this repository: there is no Rust toolchain in its container.
use dusk_bytes::Serializable;
use dusk_plonk::prelude::*;
use plonk_gadgets::{AllocatedScalar, RangeGadgets, ScalarGadgets};
use serde_json::{json, Map, Value};
use sha2::{Digest, Sha256};
use std::collections::HashMap;

/// "0x..." canonical integer -> BlsScalar (values are < q in the golden programs)
fn scalar(hex_str: &str) -> BlsScalar {
    let digits = hex_str.trim_start_matches("0x");
    let mut bytes = [0u8; 32];
    let padded = format!("{:0>64}", digits);
    let be = hex::decode(padded).expect("hex");
    for (i, b) in be.iter().rev().enumerate() {
        bytes[i] = *b;
    }
    BlsScalar::from_bytes(&bytes).expect("canonical scalar")
}
fn hx(s: &BlsScalar) -> String {
    let mut be = s.to_bytes().to_vec();
    be.reverse();
    let h = hex::encode(be);
    let t = h.trim_start_matches('0');
    format!("0x{}", if t.is_empty() { "0" } else { t })
}
fn list(op: &Value, key: &str) -> Vec<BlsScalar> {
    match &op[key] {
        Value::Array(a) => a.iter().map(|v| scalar(v.as_str().unwrap())).collect(),
        Value::String(s) => vec![scalar(s)],
        _ => panic!("missing {}", key),
    }
}

fn run(program: &[Value]) -> Value {
    let mut composer = StandardComposer::new();                      // zero variable + two dummy rows: 3 rows, 5 variables
    let mut cols: HashMap<usize, Vec<AllocatedScalar>> = HashMap::new();
    let mut error = Value::Null;
    'ops: for (idx, op) in program.iter().enumerate() {
        let col = |k: &str| cols[&(op[k].as_u64().unwrap() as usize)].clone();
        match op["op"].as_str().unwrap() {
            "add_input" => {
                cols.insert(idx, list(op, "values").into_iter().map(|s| AllocatedScalar::allocate(&mut composer, s)).collect());
            }
            "range_check" => {
                let (wit, mn, mx) = (col("witness"), list(op, "min"), list(op, "max"));
                let out = wit.iter().enumerate().map(|(i, w)| {
                    let v = RangeGadgets::range_check(&mut composer, mn[i % mn.len()], mx[i % mx.len()], *w);
                    AllocatedScalar { var: v, scalar: BlsScalar::zero() }
                }).collect();
                cols.insert(idx, out);
            }
            "max_bound" => {
                let (wit, mx) = (col("witness"), list(op, "max"));
                let out = wit.iter().enumerate().map(|(i, w)| {
                    let (v, _bits) = RangeGadgets::max_bound(&mut composer, mx[i % mx.len()], *w);
                    AllocatedScalar { var: v, scalar: BlsScalar::zero() }
                }).collect();
                cols.insert(idx, out);
            }
            "maybe_equal" => {
                let (a, b) = (col("a"), col("b"));
                let out = a.iter().zip(b.iter()).map(|(x, y)| AllocatedScalar { var: ScalarGadgets::maybe_equal(&mut composer, *x, *y), scalar: BlsScalar::zero() }).collect();
                cols.insert(idx, out);
            }
            "is_non_zero" => {
                let (vars, assigned) = (col("var"), list(op, "assigned"));
                for (i, v) in vars.iter().enumerate() {
                    if ScalarGadgets::is_non_zero(&mut composer, v.var, assigned[i]).is_err() {
                        error = json!([idx, "NonExistingInverse", i]);
                        break 'ops;
                    }
                }
            }
            "select_zero" => {
                let (x, s) = (col("x"), col("select"));
                let out = x.iter().zip(s.iter()).map(|(x, s)| AllocatedScalar { var: ScalarGadgets::conditionally_select_zero(&mut composer, x.var, s.var), scalar: BlsScalar::zero() }).collect();
                cols.insert(idx, out);
            }
            "select_one" => {
                let (y, s) = (col("y"), col("select"));
                let out = y.iter().zip(s.iter()).map(|(y, s)| AllocatedScalar { var: ScalarGadgets::conditionally_select_one(&mut composer, y.var, s.var), scalar: BlsScalar::zero() }).collect();
                cols.insert(idx, out);
            }
            "constrain_to_constant" => {
                let (a, k) = (col("a"), list(op, "constant"));
                let pi = if op.get("pi").is_some() { Some(list(op, "pi")) } else { None };
                for (i, v) in a.iter().enumerate() {
                    composer.constrain_to_constant(v.var, k[i % k.len()], pi.as_ref().map(|p| p[i % p.len()]));
                }
            }
            "range_gate" => {
                let bits = op["num_bits"].as_u64().unwrap() as usize;
                for v in col("witness").iter() {
                    composer.range_gate(v.var, bits);
                }
            }
            other => panic!("unknown op {}", other),
        }
    }
    // ---- snapshot (the AllocatedScalar.scalar fields above are placeholders: values are read from the composer itself)
    let d = composer.pg_dump();
    let n_rows = d.wires[0].len();
    let mut h = Sha256::new();
    h.update((d.variables.len() as u64).to_le_bytes());
    h.update((n_rows as u64).to_le_bytes());
    for v in d.variables.iter() {
        h.update(v.to_bytes());
    }
    let mut unsat = Vec::new();
    for i in 0..n_rows {
        for w in 0..4 {
            h.update((d.wires[w][i] as u64).to_le_bytes());
        }
        for k in 0..11 {
            h.update(d.selectors[k][i].to_bytes());
        }
        h.update(d.dense_pi[i].to_bytes());
        let val = |w: usize, row: usize| d.variables[d.wires[w][row]];
        let (a, b, c, dd) = (val(0, i), val(1, i), val(2, i), val(3, i));
        let s = |k: usize| d.selectors[k][i];
        let arith = s(0) * a * b + s(1) * a + s(2) * b + s(3) * c + s(4) * dd + d.dense_pi[i] + s(5);
        let delta = |f: BlsScalar| f * (f - BlsScalar::one()) * (f - BlsScalar::from(2)) * (f - BlsScalar::from(3));
        let four = BlsScalar::from(4);
        let d_next = val(3, (i + 1) % n_rows);
        let range = delta(c - four * dd) + delta(b - four * c) + delta(a - four * b) + delta(d_next - four * a);
        if s(6) * arith + s(7) * range != BlsScalar::zero() {
            unsat.push(i);
        }
    }
    let mut results = Map::new();
    for (idx, col) in cols.iter() {
        results.insert(idx.to_string(), Value::Array(col.iter().map(|v| Value::String(hx(&d.variables[variable_index(&composer, v.var)]))).collect()));
    }
    json!({"n_rows": n_rows, "n_vars": d.variables.len(), "unsat": unsat, "error": error, "results": results, "digest": hex::encode(h.finalize())})
}

/// Variable -> index: `Variable(pub(crate) usize)`; the accessor lives next to pg_dump if the crate has no public one.
fn variable_index(_composer: &StandardComposer, v: Variable) -> usize {
    // dusk-plonk 0.8 derives Debug on Variable: "Variable(17)"
    let s = format!("{:?}", v);
    s.trim_start_matches("Variable(").trim_end_matches(')').parse().expect("Variable index")
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let (src, dst) = (&args[1], &args[2]);
    let programs: Map<String, Value> = serde_json::from_str(&std::fs::read_to_string(src).unwrap()).unwrap();
    let mut out = Map::new();
    for (name, spec) in programs.iter() {
        let expected = run(spec["program"].as_array().unwrap());
        out.insert(name.clone(), json!({"expected": expected}));
    }
    std::fs::write(dst, serde_json::to_string_pretty(&Value::Object(out)).unwrap()).unwrap();
    println!("{} programs dumped to {}", programs.len(), dst);
}
