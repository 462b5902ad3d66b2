"""plonk_gadgets_b200 -- B200-native batched gadget engine for plonk_gadgets' hot path.

Only what the path needs: csrc/ (CUDA kernels + C ABI, built into libpg_b200.so), host/ (C++ mirror of the reference
API over the C ABI) and this Python mirror used by the tests and the benchmark.  Importing the package does not load the
CUDA library; creating a StandardComposer does, and fails loudly when it (or a B200) is missing -- there is no CPU path.
"""
from .api import (AllocatedScalar, CHECK_GENERIC, CHECK_SPARSE, DevicePtr, EngineError, Error, NZ_REFERENCE, NZ_UNIFORM,  # noqa: F401
                  NonExistingInverse, StandardComposer, Variables, conditionally_select_one, conditionally_select_zero,
                  is_non_zero, is_non_zero_flags, max_bound, maybe_equal, range_check, comm_unique_id, op_shape, shard_plan, template_get,
                  SHARD_EVEN, SHARD_ROWS, OP_ADD_INPUT, OP_RANGE_CHECK, OP_MAX_BOUND, OP_MAYBE_EQUAL, OP_IS_NON_ZERO, OP_SELECT_ZERO,
                  OP_SELECT_ONE, OP_CONSTRAIN, OP_RANGE_GATE)
from . import _lib  # noqa: F401


class RangeGadgets:      # re-export names of /root/reference/src/lib.rs:44
    range_check = staticmethod(range_check)
    max_bound = staticmethod(max_bound)


class ScalarGadgets:     # /root/reference/src/lib.rs:45
    conditionally_select_zero = staticmethod(conditionally_select_zero)
    conditionally_select_one = staticmethod(conditionally_select_one)
    is_non_zero = staticmethod(is_non_zero)
    is_non_zero_flags = staticmethod(is_non_zero_flags)
    maybe_equal = staticmethod(maybe_equal)
