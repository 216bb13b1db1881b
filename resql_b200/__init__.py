"""resql_b200 - B200-native execution engine for ReSQL query pipelines.

The product is the C-ABI library ``libresql_b200.so`` (include/resql_b200.h) plus the C++ host
shim ``resql_b200/host/gpu_executor.h`` that plugs it into the reference's
``executeSelectPlan`` (src/execute.h:213-247). This Python package is the test/bench harness
binding of the same C ABI; it contains no compute path of its own and raises when the CUDA
library is missing.
"""
from .native import Engine, EngineError, Result, Timings, lib_path  # noqa: F401
from .plan import Plan  # noqa: F401
from .serialize import serialize_value, serialize_result  # noqa: F401
