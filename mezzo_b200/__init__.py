"""mezzo_b200 — B200-native engine for Mezzo's two data-parallel hot paths (route generation, link update).

The product is the C-ABI library (include/mezzo_b200.h, mezzo_b200/csrc/); this package is the thin host-side
mirror used by the tests and the benchmark. Nothing here computes on the CPU."""
from ._lib import (MZ_SIM_OPT_BLOCKS, MZ_SIM_OPT_KERNEL, MZ_SIM_OPT_MAX_EVENTS, MZ_SSSP_AUTO, MZ_SSSP_EXACT, SP_INFINITY, SP_START, SP_UNSET, MzError, lib)  # noqa: F401
from .sssp import Graph, LinkTimeInfo  # noqa: F401
from .flatten import FlatNet  # noqa: F401,E402
from .sim import Simulation  # noqa: F401,E402
