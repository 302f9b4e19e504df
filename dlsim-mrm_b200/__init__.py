"""B200-native DTALite traffic-assignment hot path (drop-in for asu-trans-ai-lab/DLSim-MRM's
``network_assignment`` loop).  The product is the native library ``lib/libdtalite_b200.so``
(hand-written sm_100a CUDA behind a C ABI, see ``include/dtalite_b200.h``); this package is the thin
host-side mirror used by tests and ``bench.py``."""
from .problem import Problem, subset_origins  # noqa: F401
from .engine import Engine, EngineError, load_library, plan_tasks, read_gmns  # noqa: F401
