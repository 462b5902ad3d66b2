
// ---- appended by plonk_gadgets_b200/oracle/ref_dumper (read-only accessor for parity dumps; changes no behaviour) ----------
/// Everything the composer holds, in plain types.
pub struct PgDump {
    /// variables[i] = value of Variable(i)
    pub variables: alloc::vec::Vec<BlsScalar>,
    /// w_l, w_r, w_o, w_4 as Variable indices
    pub wires: [alloc::vec::Vec<usize>; 4],
    /// q_m q_l q_r q_o q_4 q_c q_arith q_range q_logic q_fixed_group_add q_variable_group_add
    pub selectors: [alloc::vec::Vec<BlsScalar>; 11],
    /// construct_dense_pi_vec()
    pub dense_pi: alloc::vec::Vec<BlsScalar>,
}
impl StandardComposer {
    /// Copies the circuit state out (fields are pub(crate)).
    pub fn pg_dump(&self) -> PgDump {
        let mut variables = alloc::vec![BlsScalar::zero(); self.variables.len()];
        for (var, value) in self.variables.iter() {
            variables[var.0] = *value;
        }
        let idx = |w: &alloc::vec::Vec<Variable>| w.iter().map(|v| v.0).collect::<alloc::vec::Vec<usize>>();
        PgDump {
            variables,
            wires: [idx(&self.w_l), idx(&self.w_r), idx(&self.w_o), idx(&self.w_4)],
            selectors: [
                self.q_m.clone(), self.q_l.clone(), self.q_r.clone(), self.q_o.clone(), self.q_4.clone(), self.q_c.clone(),
                self.q_arith.clone(), self.q_range.clone(), self.q_logic.clone(), self.q_fixed_group_add.clone(),
                self.q_variable_group_add.clone(),
            ],
            dense_pi: self.construct_dense_pi_vec(),
        }
    }
}
