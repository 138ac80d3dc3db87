"""CPU oracle for the Galah stage-1/stage-2 hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``galah_b200`` (the product) must never
import it; ``tests/test_abi_symbols.py`` enforces that.

The arithmetic lives in ``finch_oracle.c`` / ``skani_oracle.c`` (plain C, built by
``oracle/Makefile`` into ``liboracle.so``); this module is the ctypes binding plus
``cluster_oracle.py`` (pure-Python restatement of ``src/clusterer.rs`` for small cases).
"""
from .binding import *  # noqa: F401,F403
