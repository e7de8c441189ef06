"""unimatch_b200 -- B200-native (sm_100a) implementation of the UniMatch matching inference path.

    from unimatch_b200 import UniMatch      # drop-in for reference unimatch.unimatch.UniMatch (inference)

The C-ABI library (libunimatch_sm100.so) is loaded -- and built with nvcc when the in-tree copy is stale -- the first
time `ops`, `UniMatch` or one of the inference drivers is touched; `spec`, `synthetic` and `sharding` are plain Python and
import without it (the CPU reference arm of bench.py and the oracle tests never map the product library).
There is no CPU / eager fallback for the `torch.ops.unimatch_sm100.*` ops.
"""
import importlib

from .spec import BASELINE_CONFIGS, WORKLOADS, param_spec   # noqa: F401

_LAZY = {
    "ops": (".ops", None),
    "UniMatch": (".unimatch", "UniMatch"),
    "InputPadder": (".inference", "InputPadder"),
    "forward_backward_consistency_check": (".inference", "forward_backward_consistency_check"),
    "infer_flow": (".inference", "infer_flow"),
    "infer_stereo": (".inference", "infer_stereo"),
    "infer_depth": (".inference", "infer_depth"),
    "BatchedFlowRunner": (".inference", "BatchedFlowRunner"),
}

__all__ = ["UniMatch", "ops", "WORKLOADS", "BASELINE_CONFIGS", "param_spec", "InputPadder", "infer_flow", "infer_stereo",
           "infer_depth", "BatchedFlowRunner", "forward_backward_consistency_check"]


def __getattr__(name):
    if name in _LAZY:
        mod, attr = _LAZY[name]
        m = importlib.import_module(mod, __name__)
        val = m if attr is None else getattr(m, attr)
        globals()[name] = val
        return val
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
