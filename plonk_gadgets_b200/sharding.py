"""Multi-GPU plumbing: one process per GPU, instances sharded contiguously, no collective inside the data path.

Gadget instances are independent (each call touches only its own new variables plus the shared zero variable,
/root/reference/src/range.rs:119-158, /root/reference/src/scalar.rs:41,:83), so a batch of n instances is split into
contiguous ranges, each rank runs the same kernels on its range, and the only exchanges are
  * an all-reduce of the verdict (sum of unsatisfied-row counts and error counts, min of the first bad row), and
  * an optional all-gather of the per-instance results / witness shards,
both through torch.distributed (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).
Row and Variable numbering of rank r's shard is offset by what ranks < r appended, so that the union of the shards is the
sequential composer of the whole batch.
"""
from __future__ import annotations

from dataclasses import dataclass

UINT64_MAX = 2 ** 64 - 1
FRESH_ROWS, FRESH_VARS = 3, 5     # StandardComposer::new(): every rank's local composer starts with these


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous instance range [lo, hi) of `rank`: floor(n*r/G) .. floor(n*(r+1)/G)."""
    return n * rank // world, n * (rank + 1) // world


@dataclass
class ShardOffsets:
    """Where a rank's gadget rows/variables sit in the sequential composer of the whole batch."""
    row_offset: int
    var_offset: int


def shard_offsets(n: int, rank: int, world: int, rows_per_instance: int, vars_per_instance: int) -> ShardOffsets:
    """Global id = local id - (fresh composer part) + offset, for one batched call of uniform instances."""
    lo, _ = shard_range(n, rank, world)
    return ShardOffsets(lo * rows_per_instance, lo * vars_per_instance)


def allreduce_verdict(n_unsat: int, first_bad_global: int | None, n_err: int, device=None):
    """Combine per-shard verdicts: (sum n_unsat, min first_bad, sum n_err).  `first_bad_global` must already be in global row
    numbering (or None).  Works on any initialised torch.distributed backend; returns the local values when not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return n_unsat, first_bad_global, n_err
    fb = UINT64_MAX if first_bad_global is None else first_bad_global
    # int64 tensors: counts are < 2^63; the first-bad row is sent as two 32-bit halves to keep the MIN reduction exact
    sums = torch.tensor([n_unsat, n_err], dtype=torch.int64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    hi = torch.tensor([fb >> 32], dtype=torch.int64, device=device)
    dist.all_reduce(hi, op=dist.ReduceOp.MIN)
    lo_val = (fb & 0xFFFFFFFF) if (fb >> 32) == int(hi.item()) else 0xFFFFFFFF
    lo = torch.tensor([lo_val], dtype=torch.int64, device=device)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    first = (int(hi.item()) << 32) | int(lo.item())
    return int(sums[0].item()), (None if first == UINT64_MAX else first), int(sums[1].item())


def allgather_scalars(local, device=None):
    """All-gather of per-instance scalar shards ((n_local, 4) int64/uint64 tensors of equal n_local) -> (world*n_local, 4)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def allgather_ragged(local, device=None):
    """All-gather of shards whose lengths differ between ranks (n not a multiple of the world size): the lengths are exchanged
    first, every shard is padded to the longest, gathered, and trimmed.  Returns the list of the ranks' shards, in rank order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [local]
    world = dist.get_world_size()
    lens = torch.zeros(world, dtype=torch.int64, device=local.device)
    lens[dist.get_rank()] = local.shape[0]
    dist.all_reduce(lens, op=dist.ReduceOp.SUM)
    longest = int(lens.max().item())
    padded = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return [out[r * longest: r * longest + int(lens[r].item())] for r in range(world)]


def allgather_witness_shards(local_vars, call_sizes):
    """Gather of witness shards: assemble the variable table of the SEQUENTIAL composer of the whole batch from the ranks' tables.

    `local_vars`: this rank's variable table in Variable order ((FRESH_VARS + sum(call_sizes), 4) int64 tensor, e.g.
    `composer.variables()`); `call_sizes`: how many variables each batched call (add_input, range_check, ...) appended on this
    rank, in call order -- every rank makes the same calls on its own instance range.  In the sequential composer call k's
    variables of all instances are contiguous, instance-major, so the result is
        [fresh variables] [call 0: rank 0 | rank 1 | ...] [call 1: rank 0 | rank 1 | ...] ...
    (the fresh composer's five variables are the same on every rank and are taken once)."""
    import torch
    assert local_vars.shape[0] == FRESH_VARS + sum(call_sizes)
    shards = allgather_ragged(local_vars)
    sizes = allgather_ragged(torch.tensor([list(call_sizes)], dtype=torch.int64, device=local_vars.device))
    parts = [shards[0][:FRESH_VARS]]
    offs = [FRESH_VARS] * len(shards)
    for k in range(len(call_sizes)):
        for r, sh in enumerate(shards):
            cnt = int(sizes[r][0, k].item())
            parts.append(sh[offs[r]: offs[r] + cnt])
            offs[r] += cnt
    return torch.cat(parts, dim=0)


# ---------------------------------------------------------------------------------------------------------------------------
# Mixed circuits over several GPUs through the C ABI (pg_shard_plan / pg_check_sharded / pg_gather_*; include/pg_b200.h).
# A circuit is a list of calls; call k is a dict
#     {"gadget": OP_*, "n": instances, "group": g, "num_bits": k (range gadgets), ...operands}
# with operands naming EARLIER calls by index ("witness", "a", "b", "x", "select", "var") or carrying data (n scalars:
# "values", "assigned", "constant"; one scalar: "min", "max").  Calls of one group share the instance index space.
def plan_of(circuit, world: int, policy: int = 0, _cdll=None):
    import plonk_gadgets_b200 as pg
    return pg.shard_plan([(c["gadget"], c.get("num_bits", 0), c["n"], c["group"]) for c in circuit], world, policy, _cdll=_cdll)


def run_circuit(composer, circuit, mine=None):
    """Replays `circuit` on `composer`; with `mine` (this rank's row of the plan) only instances [inst_lo, inst_hi) of every call.
    Returns (per-call results: Variables or None, local NonExistingInverse count)."""
    import plonk_gadgets_b200 as pg
    out, n_err = [], 0
    for k, c in enumerate(circuit):
        lo, hi = (0, c["n"]) if mine is None else (mine[k].inst_lo, mine[k].inst_hi)
        cut = lambda x: x[lo:hi]
        g = c["gadget"]
        if g == pg.OP_ADD_INPUT:
            out.append(composer.add_input(cut(c["values"])))
        elif g == pg.OP_RANGE_CHECK:
            out.append(pg.range_check(composer, c["min"], c["max"], out[c["witness"]]))
        elif g == pg.OP_MAX_BOUND:
            out.append(pg.max_bound(composer, c["max"], out[c["witness"]])[0])
        elif g == pg.OP_MAYBE_EQUAL:
            out.append(pg.maybe_equal(composer, out[c["a"]], out[c["b"]]))
        elif g == pg.OP_IS_NON_ZERO:
            pg.is_non_zero_flags(composer, out[c["var"]], cut(c["assigned"]), pg.NZ_UNIFORM)      # host or device scalars
            n_err += composer.last_n_err
            out.append(None)
        elif g == pg.OP_SELECT_ZERO:
            out.append(pg.conditionally_select_zero(composer, out[c["x"]], out[c["select"]]))
        elif g == pg.OP_SELECT_ONE:
            out.append(pg.conditionally_select_one(composer, out[c["x"]], out[c["select"]]))
        elif g == pg.OP_CONSTRAIN:
            composer.constrain_to_constant(out[c["a"]], cut(c["constant"]))
            out.append(None)
        elif g == pg.OP_RANGE_GATE:
            composer.range_gate(out[c["witness"]], c["num_bits"])
            out.append(None)
        else:
            raise ValueError(f"unknown gadget {g}")
    return out, n_err
