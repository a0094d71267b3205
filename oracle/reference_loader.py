"""ORACLE tooling (test infrastructure; never imported by the product path).

Imports the *unmodified* reference's post-classification code -- ``Predict.get_region_potential_svtypes``
(src/network/predict.py:29-145), ``write_results_to_vcf`` / ``refine_type`` (src/network/output.py:352-598)
and ``genotyper`` (src/network/genotype.py:17-73) -- from ``/root/reference`` in the build container,
stubbing only the third-party imports that are absent here (tensorflow, bs4: unused on this path;
pysam: ``oracle/pysam_stub`` or an in-memory fake, see ``FakePysam``).  Used by
``oracle/make_calls_golden.py`` and by the live cross-checks in ``tests/``."""
from __future__ import annotations

import contextlib
import os
import sys
import types
import warnings

import numpy as np

REF = os.environ.get("SVISION_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "src", "network"))


@contextlib.contextmanager
def reference_modules():
    """Yields a namespace with ``Predict``, ``write_results_to_vcf``, ``refine_type``, ``genotyper`` and
    the ``genotype`` module (whose ``pysam`` attribute a caller may replace)."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.dont_write_bytecode = True                      # the reference mount is read-only
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pysam_stub"))
    sys.path.insert(0, REF)
    tf = types.ModuleType("tensorflow")
    bs4 = types.ModuleType("bs4")
    bs4.BeautifulSoup = object
    el = types.ModuleType("bs4.element")
    el.NavigableString = str
    bs4.element = el
    sys.modules.update({"tensorflow": tf, "bs4": bs4, "bs4.element": el})
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from src.network.predict import Predict
            from src.network import output, genotype
        yield types.SimpleNamespace(Predict=Predict, write_results_to_vcf=output.write_results_to_vcf,
                                    refine_type=output.refine_type, genotyper=genotype.genotyper,
                                    genotype=genotype, output=output)
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):            # drop only the reference's modules and the stubs
            if k not in saved_mods and k.split(".")[0] in ("src", "pysam", "tensorflow", "bs4"):
                del sys.modules[k]


class _Aln:
    __slots__ = ("query_name", "is_unmapped", "is_secondary", "mapping_quality", "reference_start", "reference_end")


class FakePysam:
    """Stands in for the ``pysam`` module inside ``src.network.genotype``: ``AlignmentFile(path, 'r')``
    serves the synthetic alignments of ``svision_b200.sites.make_alignments`` with pysam's region
    semantics (records overlapping ``[start, stop)`` in coordinate order)."""

    def __init__(self, alignments: dict):
        self._a = alignments
        self.opens = 0

    def AlignmentFile(self, path, mode="r"):             # noqa: N802  (pysam's name)
        self.opens += 1
        return _FakeBam(self._a)


class _FakeBam:
    def __init__(self, a):
        self._a = a

    def get_reference_length(self, contig):
        return self._a["contig_length"]

    def fetch(self, contig=None, start=None, stop=None):
        a = self._a
        s, e = a["reference_start"], a["reference_end"]
        hit = np.flatnonzero((s < stop) & (e > start))
        for i in hit.tolist():
            r = _Aln()
            r.query_name = a["query_name"][i]
            r.is_unmapped = bool(a["is_unmapped"][i])
            r.is_secondary = bool(a["is_secondary"][i])
            r.mapping_quality = int(a["mapping_quality"][i])
            r.reference_start = int(s[i])
            r.reference_end = int(e[i])
            yield r
