"""xbitops_b200 -- B200-native (sm_100a) implementation of the XbitOps hot path:
group-wise 2..8-bit dequantisation to fp16 and the fused A16Wx GEMV / skinny GEMM.

    import xbitops_b200 as XbitOps          # drop-in for the reference's extension module
    w = XbitOps.dequant(qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias)
    y = XbitOps.gemv(x, qweight, scales, qzeros, groupsize, bits, in_features, add_zero_bias)

The compiled pybind11 twin (module name `XbitOps`, same as the reference) is built by
`python -m xbitops_b200._build --torch`; both go through the same C ABI (include/xbitops_b200.h).
There is no CPU fallback: the ops raise if the CUDA library is missing.
"""
from . import capi, synth  # noqa: F401
from .capi import GEMV_AUTO, GEMV_GENERIC, GEMV_MMA, GEMV_PERSIST, GEMV_SIMT  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch is imported lazily so that `import xbitops_b200.synth` stays light
    if name in ("dequant", "gemv", "gemv_multi", "set_static_weights", "get_static_weights", "set_native_bf16", "get_native_bf16"):
        from . import ops
        return getattr(ops, name)
    if name in ("QLinear", "pack_qweight", "pack_qzeros", "quantize_rtn"):
        from . import qlinear
        return getattr(qlinear, name)
    if name in ("ShardedQLinear", "shard_columns"):
        from . import sharded
        return getattr(sharded, name)
    raise AttributeError(name)
