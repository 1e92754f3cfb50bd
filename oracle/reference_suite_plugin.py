"""pytest plugin that points the REFERENCE's own test modules at the drop-in package.

    PYTHONPATH=/root/repo python -m pytest -p oracle.reference_suite_plugin /root/reference/tests

(build container only: needs /root/reference; driven by tests/test_reference_suite_cpu.py).  Every module of ``theboss``
that ``theboss_b200`` rebuilds (same relative module path) is registered in ``sys.modules`` under its reference name, so
``from theboss.simulation_strategies.simulation_strategy_factory import ...`` in the reference's tests resolves to the
drop-in; modules outside the permanent hot path (mean-field strategies, TVD helpers, network simulation) stay the
reference's own.  There is no GPU in the build container, so the arithmetic underneath comes from the CPU oracle through
oracle/handle_standin.py: what this run checks is the API surface and host logic of the drop-in, not the kernels.
"""
import importlib
import os
import pkgutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("THEBOSS_REFERENCE", "/root/reference")
for p in (os.path.join(REPO, "oracle", "refshim"), REPO, REF):   # REF first: `tests` must be the reference's package
    if p not in sys.path:
        sys.path.insert(0, p)

import theboss  # noqa: E402  (the reference package: parent of the aliased modules)
import theboss_b200  # noqa: E402
from theboss_b200 import _native  # noqa: E402
from oracle import handle_standin as oracle_handle  # noqa: E402

ALIASED = []
for info in pkgutil.walk_packages(theboss_b200.__path__, "theboss_b200."):
    rel = info.name[len("theboss_b200."):]
    if info.ispkg or rel.split(".")[-1].startswith("_") or rel == "distributed":
        continue
    if not os.path.exists(os.path.join(REF, "theboss", *rel.split(".")) + ".py"):
        continue
    module = importlib.import_module(info.name)
    sys.modules["theboss." + rel] = module
    parent = importlib.import_module("theboss." + rel.rsplit(".", 1)[0]) if "." in rel else theboss
    setattr(parent, rel.rsplit(".", 1)[-1], module)
    ALIASED.append(rel)

# BOSSPERM_SUITE_HANDLE=cuda: keep the real handle (a machine that has both the reference checkout and a B200 then runs the
# reference's suite against the kernels themselves); default: the CPU stand-in.
_HANDLE = None
if os.environ.get("BOSSPERM_SUITE_HANDLE", "oracle") != "cuda":
    _HANDLE = oracle_handle.OracleHandle()
    _native.default_handle = lambda device=0: _HANDLE


# StrategyType members that never evaluate a permanent (FIXED_LOSS, UNIFORM_LOSS, CLIFFORD_R) are not part of the drop-in:
# its factory raises NotImplementedError for them.  The reference's tests of those mean-field strategies go through the
# same factory, so for THIS run the factory falls through to the reference's own classes, built as the reference factory
# builds them (simulation_strategy_factory.py:116-172 there); those tests then exercise reference code only.
from theboss_b200.simulation_strategies import simulation_strategy_factory as _ssf  # noqa: E402

_generate_drop_in = _ssf.SimulationStrategyFactory.generate_strategy
FELL_THROUGH = [0]


def _generate_with_reference_fall_through(self):
    try:
        return _generate_drop_in(self)
    except NotImplementedError:
        FELL_THROUGH[0] += 1
        kind, cfg = self.strategy_type, self.experiment_configuration
        if kind == _ssf.StrategyType.UNIFORM_LOSS:
            from theboss.simulation_strategies.uniform_loss_simulation_strategy import UniformLossSimulationStrategy
            return UniformLossSimulationStrategy(cfg.interferometer_matrix, cfg.number_of_modes, cfg.uniform_transmissivity)
        if kind == _ssf.StrategyType.CLIFFORD_R:
            from theboss.simulation_strategies.cliffords_r_simulation_strategy import CliffordsRSimulationStrategy
            return CliffordsRSimulationStrategy(cfg.interferometer_matrix)
        from theboss.simulation_strategies.fixed_loss_simulation_strategy import FixedLossSimulationStrategy
        return FixedLossSimulationStrategy(
            interferometer_matrix=cfg.interferometer_matrix, number_of_photons_left=cfg.number_of_particles_left,
            number_of_observed_modes=cfg.number_of_modes, network_simulation_strategy=cfg.network_simulation_strategy)


_ssf.SimulationStrategyFactory.generate_strategy = _generate_with_reference_fall_through


def pytest_runtest_setup(item):
    """The reference's tests are statistical and unseeded (each accepts with probability 1 - 1e-4 or so); seeding both
    generators from the test's name makes this run reproducible: once green, always green."""
    import random
    import zlib

    import numpy

    key = zlib.crc32(item.nodeid.encode())
    numpy.random.seed(key)
    random.seed(key)


def pytest_terminal_summary(terminalreporter):
    terminalreporter.write_line(f"factory requests for members outside the drop-in, served by reference classes: {FELL_THROUGH[0]}")
    terminalreporter.write_line(f"theboss -> theboss_b200 for {len(ALIASED)} modules: " + ", ".join(sorted(ALIASED)))
    if _HANDLE is not None:
        terminalreporter.write_line(f"oracle-backed handle served {_HANDLE.calls} calls from the drop-in package")
    else:
        terminalreporter.write_line(f"CUDA handle launched {_native.default_handle(0).launch_count()} kernels for the drop-in package")
