"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference (Python 2) in this container.

The reference at /root/reference is Python 2 (print statements, xrange, integer `/`).  This
module never copies its sources into the repo: it reads the three hot-path files where they
lie, applies a handful of *syntax-only* rewrite rules in memory, and executes the result as
throw-away modules.  It is used by `oracle/make_golden.py` (to write tests/golden/*.npz) and by
nothing else; it cannot run on the GPU box (no /root/reference there) and no product code,
test marked `gpu`, `smoke()` or `bench.py` imports it.

Rewrite rules:
  * `print X`            -> `print(X)`
  * `xrange`             -> `range`
  * Python-2 integer `/` -> `//` at MGCMTSolver.py:75-76,81,107-108,134-135,350,394,397
    (the places where the quotient is used as an array size / grid size)
  * `diags([1, -2, 1], ...)` integer literals: left alone (scipy only warns).
  * PotWellSolver.py:151-152,177-178: the float slice bounds `np.floor(...)` produce are wrapped in
    `int()` (old numpy accepted float indices); `from pylab import *` is served by a stub module that
    re-exports numpy (+ `np`, `math`, `numpy.linalg`'s eigh/eigvalsh) because matplotlib is absent.
"""
from __future__ import annotations

import os
import re
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("MGCMT_REFERENCE_ROOT", "/root/reference")

_PRINT_RE = re.compile(r"^(\s*)print (.*)$")

# (file, 1-based line) -> list of (old, new) substring replacements applied on that line only.
_INT_DIV_LINES = {
    "MGCMTSolver.py": {
        75: [("n / 2", "n // 2")],
        76: [("n / 2", "n // 2")],
        81: [("n / 2", "n // 2")],
        107: [("n / 2", "n // 2")],
        108: [("n / 2", "n // 2")],
        134: [("n / 2", "n // 2")],
        135: [("n / 2", "n // 2")],
        350: [("(n/2)", "(n//2)")],
        394: [("n / 2", "n // 2")],
        397: [("n / 4", "n // 4")],
    },
    "PotWellSolver.py": {
        151: [("0:self.potWellBoundary1", "0:int(self.potWellBoundary1)")],
        152: [("self.potWellBoundary2:", "int(self.potWellBoundary2):")],
        177: [("0:self.potWellBoundary1", "0:int(self.potWellBoundary1)")],
        178: [("self.potWellBoundary2:", "int(self.potWellBoundary2):")],
    },
}


def _translate(filename: str, text: str) -> str:
    out = []
    fixes = _INT_DIV_LINES.get(filename, {})
    for lineno, line in enumerate(text.splitlines(), start=1):
        m = _PRINT_RE.match(line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        line = line.replace("xrange(", "range(")
        for old, new in fixes.get(lineno, ()):  # Python-2 floor division of ints
            if old not in line:
                raise RuntimeError("reference drifted: %s:%d lacks %r" % (filename, lineno, old))
            line = line.replace(old, new)
        out.append(line)
    return "\n".join(out) + "\n"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "MGCMTSolver.py"))


def load_reference():
    """Return (MGCMTStencilMaker, MGCMTSolver, MGCMTProcessor) classes of the real reference."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    mods = {}
    for name in ("MGCMTStencilMaker", "MGCMTProcessor", "MGCMTSolver"):
        fn = name + ".py"
        with open(os.path.join(REFERENCE_ROOT, fn), "r") as fh:
            src = _translate(fn, fh.read())
        mod = types.ModuleType(name)
        mod.__file__ = os.path.join(REFERENCE_ROOT, fn)
        sys.modules[name] = mod  # MGCMTSolver does `from MGCMTStencilMaker import ...`
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            exec(compile(src, mod.__file__, "exec"), mod.__dict__)
        mods[name] = mod
    return (mods["MGCMTStencilMaker"].MGCMTStencilMaker,
            mods["MGCMTSolver"].MGCMTSolver,
            mods["MGCMTProcessor"].MGCMTProcessor)


def _exec_module(name):
    fn = name + ".py"
    with open(os.path.join(REFERENCE_ROOT, fn), "r") as fh:
        src = _translate(fn, fh.read())
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REFERENCE_ROOT, fn)
    sys.modules[name] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def load_potwell():
    """Return the reference's (baseCompounds, PotentialWell, PotWellSolver) modules -- the multiband
    Hamiltonian builder ThesisProblem.py:26-40 drives (PotWellSolver.makeMatrix, PotWellSolver.py:54-233)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if "pylab" not in sys.modules:
        import math
        import numpy
        stub = types.ModuleType("pylab")
        stub.__dict__.update({k: getattr(numpy, k) for k in dir(numpy) if not k.startswith("_")})
        stub.np = numpy
        stub.math = math
        stub.eigh = numpy.linalg.eigh
        stub.eigvalsh = numpy.linalg.eigvalsh
        sys.modules["pylab"] = stub
    return tuple(_exec_module(n) for n in ("baseCompounds", "PotentialWell", "PotWellSolver"))
