"""Import shim for the live Python reference (TEST INFRASTRUCTURE ONLY).

The reference at /root/reference does not run as shipped on numpy >= 1.24:
  * BinaryPolarEncoderDecoder.py:279,285,299,305 call `.normalizeDistList()` which no
    VectorDistribution defines (only `.normalize()` exists, BinaryMemorylessVectorDistribution.py:79,
    BinaryTrellis.py:297, CollectionOfBinaryTrellises.py:99);
  * QaryMemorylessVectorDistribution.py:16,72 / QaryPolarEncoderDecoder.py:239,426,507 use np.float /
    np.int / np.product.
This module patches both WITHOUT editing the reference, and exposes its modules under `ref.*`.

It exists only in this build container: nothing under tests/ -m gpu, smoke() or bench.py imports it.
It is used by oracle/gen_golden.py to produce tests/golden/*.npz and by the optional live-reference
tests (skipped when /root/reference is absent).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("POLARCUB_REFERENCE", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "BinaryPolarEncoderDecoder.py"))


_ref = None


def load():
    """Return a namespace with the reference modules (imports once)."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    import numpy as np

    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "product"):
        np.product = np.prod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # import order matters (circular imports): vector distributions first
    from VectorDistributions import BinaryMemorylessVectorDistribution as BMVD
    from VectorDistributions import QaryMemorylessVectorDistribution as QMVD
    from VectorDistributions import BinaryTrellis as BT
    from VectorDistributions import CollectionOfBinaryTrellises as CBT
    from ScalarDistributions import BinaryMemorylessDistribution as BMD
    from ScalarDistributions import QaryMemorylessDistribution as QMD
    import BinaryPolarEncoderDecoder as BPED
    import QaryPolarEncoderDecoder as QPED
    import Guardbands

    for cls in (BMVD.BinaryMemorylessVectorDistribution, BT.BinaryTrellis, CBT.CollectionOfBinaryTrellises):
        if not hasattr(cls, "normalizeDistList"):
            cls.normalizeDistList = cls.normalize
    ns = types.SimpleNamespace(BMVD=BMVD, QMVD=QMVD, BT=BT, CBT=CBT, BMD=BMD, QMD=QMD, BPED=BPED, QPED=QPED,
                               Guardbands=Guardbands)
    _ref = ns
    return ns
