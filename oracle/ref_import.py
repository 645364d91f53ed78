"""Import the UNMODIFIED reference (hhcaz/CNS at /root/reference) on top of the PyG shim.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build
container); the GPU box never has it, so nothing under tests -m gpu / bench / smoke
may call this.  It is used by tests/golden/make_golden.py to generate the committed
fixtures and by CPU tests that cross-check oracle/cns_oracle.py against the
reference's own code.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("CNS_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cns", "models", "graph_vs.py"))


def setup():
    """Put shim, stubs and the reference on sys.path (idempotent)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, os.path.join(_HERE, "stubs"), os.path.join(_HERE, "pyg_shim")):
        if p not in sys.path:
            sys.path.insert(0, p)


def load():
    """Returns (graph_vs module, graph_gen module, functions module) of the reference."""
    setup()
    import cns.models.graph_vs as graph_vs
    import cns.midend.graph_gen as graph_gen
    import cns.models.functions as functions
    return graph_vs, graph_gen, functions
