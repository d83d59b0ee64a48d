"""oracle/ref_shim.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Imports the UNMODIFIED reference (/root/reference, pure Python) in this build
container so its functions can be called directly to (a) pin the C oracle and
(b) generate the golden fixtures under tests/golden/.  /root/reference does not
exist on the GPU box: nothing executed there imports this module.

Two monkeypatches, no reference file is edited:
  1. open(..., "rU")  (scoary/methods.py:184,185,192,346) -> strip the "U",
     removed in Python 3.11;
  2. scipy.stats.binom_test (methods.py:1267,1271), removed in SciPy 1.12 ->
     binomtest(...).pvalue.
"""
import builtins
import os
import sys

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scoary"))


_methods = None


def load():
    """Return the reference's scoary.methods module (shimmed)."""
    global _methods
    if _methods is not None:
        return _methods
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    import scipy.stats as ss
    _open = builtins.open

    def open_no_U(f, mode="r", *a, **k):
        if isinstance(mode, str):
            mode = mode.replace("U", "")
        return _open(f, mode, *a, **k)

    builtins.open = open_no_U
    if not hasattr(ss, "binom_test"):
        ss.binom_test = lambda x, n=None, p=0.5: ss.binomtest(int(x), int(n), p).pvalue
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from scoary import methods  # noqa: E402  (the reference package)
    _methods = methods
    return methods


def run_cli(argv):
    """Run the reference CLI (methods.main) with argv; returns the exit status."""
    m = load()
    old = sys.argv
    sys.argv = ["scoary"] + list(argv)
    try:
        m.main()
    except SystemExit as e:
        return e.code
    finally:
        sys.argv = old
    return 0


def parse_newick_like(text):
    """Scoary-written trees (StoreUPGMAtreeToFile, methods.py:741-752) -> nested lists."""
    import ast
    return ast.literal_eval(text.strip().rstrip(";").replace("(", "[").replace(")", "]"))
