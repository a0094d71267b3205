"""Signatures -> segment pairs as packed rows, skipping the text BED: the host
step immediately *before* the encode+classify path.

The reference writes one BED line per segment pair (``proc_one_cluster`` / ``proc_one_sig``,
``src/collection/output_clusters.py:93-210``) and the prediction stage parses the text back
(``src/network/create_batch.py:29-61``).  :func:`generate_pairs` produces the same rows directly as a
:class:`svision_b200.bed.SegmentsTable` -- the ``int32[N,12]`` block is what ``Classifier.classify``
takes -- from a :class:`SignatureTable` holding all signatures of a chromosome in flat arrays; the
geometry runs in one native pass (``svx_pairs_generate``, ``csrc/host_pairs.cpp``).
:func:`to_bed_lines` renders the reference's exact text when the file is still wanted
(``writer_cluster_to_file``, ``output_clusters.py:31-91``)."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np

from . import _lib
from .bed import SegmentsTable

ALN_FIELDS = ("ref_start", "ref_end", "q_start", "q_end", "is_reverse")     # classes.py:85-108


@dataclass
class SignatureTable:
    """All signatures of the clusters that passed the writer's filters, flattened.

    ``cluster_region[c]`` is the BED region string ``contig+int(cstart)+int(cend)+coverage``
    (``output_clusters.py:106-107``); signature ``s`` belongs to cluster ``sig_cluster[s]``
    (non-decreasing), owns alignments ``aln[sig_aln_off[s]:sig_aln_off[s+1]]`` (rows of
    :data:`ALN_FIELDS`, in ``sorted_aligns`` order, absolute coordinates) and breakpoints
    ``bkp[sig_bkp_off[s]:sig_bkp_off[s+1]]`` (start, end, length)."""
    cluster_region: list
    sig_cluster: np.ndarray
    sig_aln_off: np.ndarray
    aln: np.ndarray
    sig_bkp_off: np.ndarray
    bkp: np.ndarray
    qname: np.ndarray
    sig_type: np.ndarray
    mechanism: np.ndarray

    def __len__(self) -> int:
        return int(self.sig_cluster.shape[0])

    @classmethod
    def from_clusters(cls, clusters: Iterable, min_support: int = 1, max_sv_size: int = None) -> "SignatureTable":
        """From reference-style cluster objects (``contig, cstart, cend, coverage, read_num`` and
        ``get_signatures()`` yielding ``sorted_aligns, bkps, qname, type, mechanism``:
        ``src/collection/classes.py:8-24,122-175``), applying the writer's two cluster filters
        (``output_clusters.py:49-53``).  Does not modify the clusters (``get_segs_cords`` rewrites
        ``sorted_aligns`` in place; this does not)."""
        regions, sig_cluster, aln_off, aln, bkp_off, bkp, qn, ty, me = [], [], [0], [], [0], [], [], [], []
        for cl in clusters:
            if max_sv_size is not None and int(cl.cend) - int(cl.cstart) > max_sv_size:
                continue
            if cl.read_num < min_support:
                continue
            regions.append(f"{cl.contig}+{int(cl.cstart)}+{int(cl.cend)}+{cl.coverage}")
            for sig in cl.get_signatures():
                sig_cluster.append(len(regions) - 1)
                for a in sig.sorted_aligns:
                    aln.append([a["ref_start"], a["ref_end"], a["q_start"], a["q_end"], 1 if a["is_reverse"] else 0])
                aln_off.append(len(aln))
                bkp.extend([int(b[0]), int(b[1]), int(b[2])] for b in sig.bkps)
                bkp_off.append(len(bkp))
                qn.append(sig.qname)
                ty.append(sig.type)
                me.append(sig.mechanism)
        obj = lambda v: np.array(v, dtype=object) if v else np.empty(0, dtype=object)      # noqa: E731
        return cls(regions, np.array(sig_cluster, dtype=np.int64), np.array(aln_off, dtype=np.int64),
                   np.array(aln, dtype=np.int64).reshape(-1, len(ALN_FIELDS)), np.array(bkp_off, dtype=np.int64),
                   np.array(bkp, dtype=np.int64).reshape(-1, 3), obj(qn), obj(ty), obj(me))


def generate_pairs(sigs: SignatureTable) -> SegmentsTable:
    """Every non-linear segment pair of every signature, in the reference's file order.  The returned
    table carries an extra ``sub_id`` array (BED column 14, which the reference reader drops)."""
    lib = _lib.load()
    n_sig = len(sigs)
    aln_off = np.ascontiguousarray(sigs.sig_aln_off, dtype=np.int64)
    aln = np.ascontiguousarray(sigs.aln, dtype=np.int64)
    bkp_off = np.ascontiguousarray(sigs.sig_bkp_off, dtype=np.int64)
    if aln_off.shape[0] != n_sig + 1 or bkp_off.shape[0] != n_sig + 1:
        raise ValueError("offset arrays must have one entry per signature plus one")
    if n_sig and (aln_off[-1] > aln.shape[0] or bkp_off[-1] > sigs.bkp.shape[0] or np.any(np.diff(aln_off) < 0)
                  or np.any(np.diff(bkp_off) < 0) or aln_off[0] < 0 or bkp_off[0] < 0):
        raise ValueError("offset arrays do not describe the alignment / breakpoint arrays")
    need = ctypes.c_int64(0)
    args = (n_sig, aln_off.ctypes.data, aln.ctypes.data, bkp_off.ctypes.data)
    rc = lib.svx_pairs_generate(*args, 0, None, None, ctypes.byref(need))
    if rc != 0:
        raise ValueError(lib.svx_last_error().decode(errors="replace"))
    n = need.value
    rows = np.empty((n, 12), dtype=np.int32)
    meta = np.empty((n, 5), dtype=np.int64)
    if n:
        rc = lib.svx_pairs_generate(*args, n, rows.ctypes.data, meta.ctypes.data, ctypes.byref(need))
        if rc != 0:
            raise ValueError(lib.svx_last_error().decode(errors="replace"))
    sig = meta[:, 0]
    main = (meta[:, 2] & 1).astype(bool)
    forward = (meta[:, 2] & 2).astype(bool)
    # 1-based position of each signature inside its cluster (sig_cnt, output_clusters.py:111-113):
    # skipped signatures keep their number
    cluster = np.asarray(sigs.sig_cluster, dtype=np.int64)
    first_of_cluster = np.searchsorted(cluster, cluster, side="left") if n_sig else cluster
    sig_cnt = (np.arange(n_sig) - first_of_cluster + 1)[sig]
    num = sig_cnt.astype(str).astype(object)
    read_num = np.where(main, num + "m", num) if n else np.empty(0, dtype=object)
    region = np.array(sigs.cluster_region, dtype=object)[cluster[sig]] if n else np.empty(0, dtype=object)
    b = np.asarray(sigs.bkp, dtype=np.int64)[meta[:, 3]] if n else np.zeros((0, 3), np.int64)
    table = SegmentsTable(rows, b[:, 0].copy(), b[:, 1].copy(), b[:, 2].copy(), region=region, read_num=read_num,
                          read_name=sigs.qname[sig], sig_type=sigs.sig_type[sig],
                          sig_score=meta[:, 4].astype(str).astype(object),
                          forward=np.where(forward, "True", "False").astype(object), mechanism=sigs.mechanism[sig])
    table.sub_id = meta[:, 1].copy()
    return table


def to_bed_lines(table: SegmentsTable) -> list:
    """The reference's BED text for these rows (``output_clusters.py:171-173,197-199`` +
    ``Segment.toString``, ``src/segmentplot/classes.py:77-83``), without the trailing newline."""
    r = table.rows.tolist()
    tf = ("False", "True")
    sub = getattr(table, "sub_id", None)
    sub = sub.tolist() if sub is not None else [1] * len(table)
    cols = [getattr(table, k).tolist() for k in ("region", "read_num", "read_name", "sig_type", "sig_score", "forward",
                                                 "mechanism")]
    b0, b1, b2 = table.bkp_start.tolist(), table.bkp_end.tolist(), table.bkp_len.tolist()
    out = []
    for i, v in enumerate(r):
        out.append("\t".join([cols[0][i], str(v[0]), str(v[1]), str(v[2]), str(v[3]), tf[v[4]], str(v[5]), str(v[6]),
                              str(v[7]), str(v[8]), tf[v[9]], str(v[10]), str(v[11]), cols[1][i], str(sub[i]),
                              cols[2][i], cols[3][i], str(b0[i]), str(b1[i]), cols[4][i], cols[5][i], cols[6][i],
                              str(b2[i])]))
    return out
