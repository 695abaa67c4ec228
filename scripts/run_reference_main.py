#!/usr/bin/env python
"""Run the reference's own driver (`src/main.py`, unmodified) against the B200 engine.

    python scripts/run_reference_main.py [/path/to/reference/src/main.py]

`runpy.run_path` does not put the script's directory on sys.path, so `import maxent`, `solver`,
`optimizer`, `gridworld`, `trajectory`, `plot` resolve to irl-maxent_b200/.  matplotlib (imported by
main.py:11) is replaced by an inert stand-in when it is not installed.  Needs a CUDA device: the
engine has no CPU fallback.
"""
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "irl-maxent_b200"))


class _Inert(types.ModuleType):
    """Accepts any attribute access / call / iteration-free use: a headless matplotlib."""

    def __getattr__(self, name):
        if name.startswith("__"):           # keep inspect / importlib happy
            raise AttributeError(name)
        return _Inert(name)

    def __call__(self, *a, **k):
        return _Inert("call")


def _count_outer_steps():
    """IRLB200_MAIN_REPORT=1: report the outer gradient steps of every `irl` / `irl_causal` call (one optimizer
    step each, maxent.py:251 / :449).  The ENGINE's optimizer classes remember every instance handed to
    `reset`; an optimizer's step counter `k` (reference: optimizer.py:33,104) is read when the script ends --
    it counts the same whether the loop ran on the host or inside irlb200_irl_small.  main.py stays untouched."""
    import atexit
    import optimizer as O
    seen = []

    def wrap(cls):
        reset = cls.reset

        def recording_reset(self, parameters):
            seen.append(self)
            return reset(self, parameters)

        cls.reset = recording_reset

    for name in ("Sga", "ExpSga"):
        wrap(getattr(O, name))
    atexit.register(lambda: print("IRLB200_MAIN_OUTER_STEPS %s" % " ".join(str(o.k) for o in seen), flush=True))


def main(path):
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        sys.modules["matplotlib"] = _Inert("matplotlib")
        sys.modules["matplotlib.pyplot"] = _Inert("matplotlib.pyplot")
    if os.environ.get("IRLB200_MAIN_SEED"):
        import numpy as np
        np.random.seed(int(os.environ["IRLB200_MAIN_SEED"]))     # main.py draws its expert data from np.random
    if os.environ.get("IRLB200_MAIN_REPORT"):
        _count_outer_steps()
    return runpy.run_path(path, run_name="__main__")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/main.py")
