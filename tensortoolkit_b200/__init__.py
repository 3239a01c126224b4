"""tensortoolkit_b200 -- B200-native block-sparse symmetric-tensor contraction
(the qlten::Contract hot path of QuantumLiquids/TensorToolkit), behind the C ABI in include/qlb200.h.
"""
from .tensor import (IN, OUT, QNKind, QNSector, Index, BlockSparseTensor, U1, fU1, U1U1, fU1U1, Z2, fZ2)
from . import qlten_io
from .contract import (Context, Match, ContractionPlan, RawPlan, contract, contract_1sector, contract_contiguous_axes, transpose,
                       default_context, AccumulateLayoutMismatch, contract_tail_head_contiguous_accumulate,
                       try_contract_tail_head_contiguous_accumulate)

from .axis_ops import AxisPlan, apply_rank2_to_axis_preserve_order, apply_two_rank2_to_axes_preserve_order  # noqa: E402

__all__ = ["AxisPlan", "apply_rank2_to_axis_preserve_order", "apply_two_rank2_to_axes_preserve_order", "IN", "OUT", "QNKind", "QNSector", "Index", "BlockSparseTensor", "U1", "fU1", "U1U1", "fU1U1", "Z2", "fZ2",
           "Context", "Match", "ContractionPlan", "RawPlan", "contract", "contract_1sector", "contract_contiguous_axes",
           "transpose", "default_context", "qlten_io", "AccumulateLayoutMismatch", "contract_tail_head_contiguous_accumulate",
           "try_contract_tail_head_contiguous_accumulate"]
