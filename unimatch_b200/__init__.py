"""unimatch_b200 -- B200-native (sm_100a) implementation of the UniMatch matching inference path.

    from unimatch_b200 import UniMatch      # drop-in for reference unimatch.unimatch.UniMatch (inference)

Importing the package loads libunimatch_sm100.so (building it with nvcc if absent) and registers the
`torch.ops.unimatch_sm100.*` custom ops; there is no CPU / eager fallback for them.
"""
from . import ops                      # noqa: F401  (loads the C-ABI library, registers the custom ops)
from .spec import BASELINE_CONFIGS, WORKLOADS, param_spec   # noqa: F401
from .unimatch import UniMatch         # noqa: F401
from .inference import InputPadder, forward_backward_consistency_check, infer_flow   # noqa: F401

__all__ = ["UniMatch", "ops", "WORKLOADS", "BASELINE_CONFIGS", "param_spec", "InputPadder", "infer_flow",
           "forward_backward_consistency_check"]
