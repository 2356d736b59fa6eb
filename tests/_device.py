"""Device selection shared by the checkers (tests/check_*.py).

default            "cuda": the checkers drive the real kernels (the `-m gpu` suite, smoke())
CLB_EMU=1          "cpu":  every kernel wrapper of controllora_b200.ops is replaced by its torch restatement
                           (tests/emu_ops.py) - host-logic runs of the same checkers inside the `-m "not gpu"` suite
CLB_DRYRUN=1       "cpu":  plumbing dry run, kernels are no-ops (outputs are uninitialised memory; nothing is compared)

The environment is read once, when the first checker is imported; the CPU suite therefore runs the emulated cases in a
subprocess (tests/run_emulated.py) and never flips the mode of the pytest process that also collects the GPU tests."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def _select() -> str:
    if os.environ.get("CLB_EMU"):
        from tests import emu_ops

        emu_ops.install()
        return "cpu"
    if os.environ.get("CLB_DRYRUN"):
        from controllora_b200 import _lib, ops

        class _Dummy:
            def __getattr__(self, name):
                return lambda *a, **k: 0

        _lib.lib = lambda: _Dummy()
        ops._req = lambda *a, **k: None
        ops._stream = lambda: None
        return "cpu"
    return "cuda"


DEV = _select()
EMULATED = bool(os.environ.get("CLB_EMU"))


def sync() -> None:
    if DEV == "cuda":
        import torch

        torch.cuda.synchronize()
