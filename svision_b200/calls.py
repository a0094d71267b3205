"""Region aggregation, type refinement, VCF record assembly and a one-pass genotyper: the host step
immediately *after* the encode+classify path.

What the reference does per region, and what this module restates:

    aggregate_region   <- Predict.get_region_potential_svtypes   (src/network/predict.py:29-145)
    refine_types       <- refine_type                            (src/network/output.py:352-467)
    region_records     <- write_results_to_vcf                   (src/network/output.py:469-598)
    AlignmentTable     <- genotyper                              (src/network/genotype.py:17-73)
    call_chromosome    <- the per-row loop of Predict.run         (src/network/predict.py:213-300)

Differences in *how*, none in *what*: the per-row loop runs over plain Python lists taken from the
BED columns in one ``tolist()`` each (no per-row string splitting, no numpy scalar boxing); and the
genotyper reads the BAM once per chromosome into sorted arrays and answers every candidate with two
binary searches, instead of re-opening the BAM for every VCF record (``genotype.py:22``).

Numeric types are kept exactly as the reference has them, because they are visible in the text it
prints: class scores stay ``numpy.float32`` through ``round(…, 2)``, ``numpy.mean`` and
``(1 - round(mean, 2)) * 100`` (``predict.py:251``, ``output.py:473-474``), the signature-score
spread is a float64 ``numpy.std`` (``output.py:551``), and QUAL is printed with ``str()``.
"""
from __future__ import annotations

from collections import Counter
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import logging
import time

import numpy as np

TYPE_NAMES = ("DEL", "INS", "INV", "DUP", "tDUP")            # predict.py:133-142
USE_NATIVE = True          # svx_calls_aggregate for tables parsed from BED text (tests switch it off to compare)
MAX_GENOTYPE_ALIGNMENTS = 500                                 # genotype.py:34
MAX_PAIRS_PER_BLOCK = 8_000_000       # (candidate, alignment) pairs genotype_many expands at once (~0.5 GB)


# ------------------------------------------------------------------------------------------------
# aggregation
# ------------------------------------------------------------------------------------------------
def aggregate_region(reads: dict) -> list:
    """``reads``: ``{read_id: {class_id: [bkp_start, bkp_end, bkp_len]}}`` in insertion order.

    Reads that carry the same *set* of classes form one candidate; its breakpoints are a running
    integer mean taken read by read, ``int((new + old*n) / (n+1))`` with true division
    (predict.py:104-106) -- not the mean of all reads, so the order of reads matters.  Candidates
    come back ordered by support, ties in first-seen order (predict.py:114), as
    ``[("INS+tDUP", [read ids], [[start, end, len], ...]), ...]``."""
    groups: dict = {}
    for read_id, calls in reads.items():
        if len(calls) == 1:                                # the common case: one class per read
            (k, b), = calls.items()
            kinds, current = (k,), [b]
        else:
            kinds = tuple(sorted(calls))
            current = [calls[k] for k in kinds]
        g = groups.get(kinds)
        if g is None:
            groups[kinds] = [[read_id], current]
            continue
        n = len(g[0])
        g[1] = [[int((c[j] + o[j] * n) / (n + 1)) for j in range(3)] for c, o in zip(current, g[1])]
        g[0].append(read_id)
    ranked = sorted(groups.items(), key=lambda kv: -len(kv[1][0]))          # stable
    return [("+".join(TYPE_NAMES[k] for k in kinds), ids, bkps) for kinds, (ids, bkps) in ranked]


def refine_types(kinds: Sequence[str], bkps: Sequence[list], min_sv_size: int) -> Tuple[list, list]:
    """An INS next to a DUP/tDUP is usually the duplicated copy itself (output.py:352-467):

    * a DUP whose end lies within 10 bp of the INS position becomes a tDUP (only when the candidate
      has a plain DUP at all: output.py:388-405,429-441);
    * if the inserted length exceeds the duplicated length by more than ``min_sv_size`` the INS
      stays, shortened to the novel part; otherwise the INS is dropped.

    Class ids are sorted (DEL<INS<INV<DUP<tDUP) and unique within a candidate, so the INS always
    precedes the duplications it is compared with."""
    kinds, bkps = list(kinds), list(bkps)
    if "INS" not in kinds or not ("DUP" in kinds or "tDUP" in kinds):
        return kinds, bkps
    ins_len = dup_len = 0
    ins_pos = -1
    for i, k in enumerate(kinds):
        if k == "INS":
            ins_pos = int(bkps[i][0])
            ins_len += int(bkps[i][2])
        elif k == "DUP" or k == "tDUP":
            dup_len += int(bkps[i][2])
            if k == "DUP" and ins_pos != -1 and abs(ins_pos - int(bkps[i][1])) < 10:
                kinds[i] = "tDUP"
    if ins_len - dup_len > min_sv_size:
        out_b = [list(b) for b in bkps]
        out_b[kinds.index("INS")][2] = ins_len - dup_len
        return kinds, out_b
    keep = [i for i, k in enumerate(kinds) if k != "INS"]
    return [kinds[i] for i in keep], [bkps[i] for i in keep]


# ------------------------------------------------------------------------------------------------
# genotyping: one pass over the BAM per chromosome
# ------------------------------------------------------------------------------------------------
class AlignmentTable:
    """All alignments of one contig as arrays in BAM (coordinate) order.

    ``genotype`` answers what ``genotyper`` (genotype.py:17-73) answers, without touching the BAM
    again: the alignments overlapping ``[start-1000, end+1000)`` are found with two binary searches
    (``reference_start`` is sorted; a running maximum of ``reference_end`` bounds the left side),
    the first 500 usable ones that do not belong to a supporting read are taken, and the
    reference-supporting reads are counted by distinct query name."""

    def __init__(self, contig_length: int, reference_start, reference_end, mapping_quality,
                 is_unmapped, is_secondary, query_names: Sequence[str]):
        self.contig_length = int(contig_length)
        self.start = np.ascontiguousarray(reference_start, dtype=np.int64)
        self.end = np.ascontiguousarray(reference_end, dtype=np.int64)
        if self.start.size and np.any(np.diff(self.start) < 0):
            raise ValueError("alignments must be in coordinate order (the reference requires a sorted BAM: SVision:141-146)")
        self.mapq = np.ascontiguousarray(mapping_quality, dtype=np.int64)
        self.skip = np.asarray(is_unmapped, dtype=bool) | np.asarray(is_secondary, dtype=bool)
        names, self.name_id = np.unique(np.asarray(query_names, dtype=object).astype(str), return_inverse=True)
        self._name_to_id = {n: i for i, n in enumerate(names.tolist())}
        self._end_running_max = np.maximum.accumulate(self.end) if self.end.size else self.end

    @classmethod
    def from_bam(cls, bam_path: str, contig: str) -> "AlignmentTable":
        """One ``fetch`` over the whole contig.  Needs ``pysam`` (a dependency of the reference:
        ``setup.py:36``); raises ImportError where it is absent."""
        import pysam
        from array import array
        bam = pysam.AlignmentFile(bam_path, "r")
        # typed arrays (8 / 1 bytes per record, amortised growth) instead of lists of Python ints
        s, e, q, u, sec, names = array("q"), array("q"), array("q"), array("b"), array("b"), []
        for a in bam.fetch(contig=contig):
            s.append(a.reference_start)
            e.append(a.reference_end if a.reference_end is not None else a.reference_start)
            q.append(a.mapping_quality)
            u.append(1 if a.is_unmapped else 0)
            sec.append(1 if a.is_secondary else 0)
            names.append(a.query_name)
        return cls(bam.get_reference_length(contig), np.frombuffer(s, dtype=np.int64), np.frombuffer(e, dtype=np.int64),
                   np.frombuffer(q, dtype=np.int64), np.frombuffer(u, dtype=np.int8).astype(bool),
                   np.frombuffer(sec, dtype=np.int8).astype(bool), names)

    def __len__(self) -> int:
        return int(self.start.size)

    def genotype(self, candidate, support_reads: Iterable[str], options) -> Tuple[str, int, int]:
        contig, start, end, kinds = candidate[0], int(candidate[1]), int(candidate[2]), candidate[3]
        lo_q, hi_q = max(0, start - 1000), min(self.contig_length, end + 1000)
        alt = set(support_reads)
        ref_no = 0
        if len(self) and hi_q > lo_q:
            lo = int(np.searchsorted(self._end_running_max, lo_q, side="right"))
            hi = int(np.searchsorted(self.start, hi_q, side="left"))
            if hi > lo:
                sl = slice(lo, hi)
                s, e, nid = self.start[sl], self.end[sl], self.name_id[sl]
                usable = (e > lo_q) & ~self.skip[sl] & (self.mapq[sl] >= options.min_mapq)
                alt_ids = [self._name_to_id[n] for n in alt if n in self._name_to_id]
                if alt_ids:
                    usable &= ~np.isin(nid, alt_ids)
                pick = np.flatnonzero(usable)[:MAX_GENOTYPE_ALIGNMENTS]
                s, e, nid = s[pick], e[pick], nid[pick]
                if len(kinds) != 1:
                    votes = np.ones(pick.size, dtype=bool)                           # genotype.py:56-57
                elif kinds[0] in ("DEL", "INV"):                                     # genotype.py:46-50
                    ov = min((end - start) / 2, 2000)
                    votes = ((s < end - ov) & (e > end + 100)) | ((s < start - 100) & (e > start + ov))
                elif kinds[0] in ("INS", "DUP"):                                     # genotype.py:52-54
                    votes = (s < start - 100) & (e > end + 100)
                else:
                    votes = np.zeros(pick.size, dtype=bool)
                ref_no = int(np.unique(nid[votes]).size)
        alt_no = len(alt)
        gt = "./."
        if len(kinds) == 1 and alt_no + ref_no >= options.min_gt_depth:              # genotype.py:65-71
            ratio = alt_no / (alt_no + ref_no)
            if ratio >= options.homo_thresh:
                gt = "1/1"
            elif ratio >= options.hete_thresh:
                gt = "0/1"
            else:
                gt = "0/0"
        return gt, ref_no, alt_no


    def genotype_many(self, candidates: Sequence, supports: Sequence[Iterable[str]], options) -> list:
        """:meth:`genotype` for every candidate of a chromosome, vectorised in blocks of bounded size: the
        expansion below materialises one entry per (candidate, alignment in its window), which in deep
        pile-ups (rDNA, centromeres, amplicons) is far more than the 500 alignments per record the
        reference looks at.  Blocks hold at most MAX_PAIRS_PER_BLOCK such pairs; a candidate whose window
        alone exceeds that goes through :meth:`genotype` (slices, no expansion)."""
        m = len(candidates)
        if m == 0:
            return []
        if len(self) == 0:
            return self._genotype_block(candidates, supports, options)
        start = np.array([int(c[1]) for c in candidates], dtype=np.int64)
        end = np.array([int(c[2]) for c in candidates], dtype=np.int64)
        lo_q, hi_q = np.maximum(0, start - 1000), np.minimum(self.contig_length, end + 1000)
        cnt = np.where(hi_q > lo_q, np.maximum(np.searchsorted(self.start, hi_q, side="left") -
                                               np.searchsorted(self._end_running_max, lo_q, side="right"), 0), 0)
        if int(cnt.sum()) <= MAX_PAIRS_PER_BLOCK:
            return self._genotype_block(candidates, supports, options)
        out, a, load = [], 0, 0
        for j in range(m + 1):
            c = int(cnt[j]) if j < m else 0
            deep = j < m and c > MAX_PAIRS_PER_BLOCK
            if j == m or deep or load + c > MAX_PAIRS_PER_BLOCK:
                if j > a:
                    out.extend(self._genotype_block(candidates[a:j], supports[a:j], options))
                a, load = j, 0
                if deep:
                    out.append(self.genotype(candidates[j], supports[j], options))
                    a = j + 1
                    continue
            load += c
        return out

    def _genotype_block(self, candidates: Sequence, supports: Sequence[Iterable[str]], options) -> list:
        """One vectorised pass: the (candidate, alignment) pairs of all windows are expanded side by
        side, filtered, capped at 500 usable alignments per candidate by a segmented running count, and
        the distinct reference-supporting read names are counted per candidate."""
        m = len(candidates)
        if m == 0:
            return []
        start = np.array([int(c[1]) for c in candidates], dtype=np.int64)
        end = np.array([int(c[2]) for c in candidates], dtype=np.int64)
        # 0: several (or no) types -> every usable alignment votes; 1: DEL/INV; 2: INS/DUP; 3: no rule
        rule = np.array([0 if len(c[3]) != 1 else 1 if c[3][0] in ("DEL", "INV") else 2 if c[3][0] in ("INS", "DUP") else 3
                         for c in candidates], dtype=np.int64)
        alts = [set(sup) for sup in supports]
        ref_no = np.zeros(m, dtype=np.int64)
        if len(self):
            lo_q, hi_q = np.maximum(0, start - 1000), np.minimum(self.contig_length, end + 1000)
            lo = np.searchsorted(self._end_running_max, lo_q, side="right")
            hi = np.searchsorted(self.start, hi_q, side="left")
            cnt = np.where(hi_q > lo_q, np.maximum(hi - lo, 0), 0)
            total = int(cnt.sum())
            if total:
                first = np.cumsum(cnt) - cnt                               # pair index where each candidate begins
                cand = np.repeat(np.arange(m), cnt)
                pos = np.arange(total) - np.repeat(first, cnt) + np.repeat(lo, cnt)
                s, e, nid = self.start[pos], self.end[pos], self.name_id[pos]
                usable = (e > lo_q[cand]) & ~self.skip[pos] & (self.mapq[pos] >= options.min_mapq)
                n_names = len(self._name_to_id)
                key = cand * n_names + nid
                alt_key = [j * n_names + self._name_to_id[nm] for j, a in enumerate(alts) for nm in a
                           if nm in self._name_to_id]
                if alt_key:
                    usable &= ~np.isin(key, np.array(alt_key, dtype=np.int64))
                running = np.cumsum(usable)
                before = np.repeat(np.where(cnt > 0, running[np.minimum(first, total - 1)] - usable[np.minimum(first, total - 1)], 0), cnt)
                pick = usable & (running - before <= MAX_GENOTYPE_ALIGNMENTS)
                st, en, r = start[cand], end[cand], rule[cand]
                ov = np.minimum((en - st) / 2, 2000)
                votes = np.where(r == 0, True,
                                 np.where(r == 1, ((s < en - ov) & (e > en + 100)) | ((s < st - 100) & (e > st + ov)),
                                          np.where(r == 2, (s < st - 100) & (e > en + 100), False)))
                distinct = np.unique(key[pick & votes])
                ref_no = np.bincount(distinct // n_names, minlength=m)
        out = []
        for j in range(m):
            alt_no, rn = len(alts[j]), int(ref_no[j])
            gt = "./."
            if rule[j] != 0 and alt_no + rn >= options.min_gt_depth:
                ratio = alt_no / (alt_no + rn)
                gt = "1/1" if ratio >= options.homo_thresh else "0/1" if ratio >= options.hete_thresh else "0/0"
            out.append((gt, rn, alt_no))
        return out


# ------------------------------------------------------------------------------------------------
# VCF records of one region
# ------------------------------------------------------------------------------------------------
def region_records(candidates: list, region: str, read_names: dict, sig_types: Sequence[str],
                   sig_scores: dict, class_scores: Sequence, mechanisms: dict, options,
                   genotype: Callable) -> List[Tuple[object, str]]:
    """``[(qual, vcf_line), ...]`` for one region (output.py:469-598).  ``genotype(candidate,
    support_read_names, options) -> (GT, DR, DV)`` is :meth:`AlignmentTable.genotype` or the
    reference's ``genotyper``.  ``qual`` keeps its Python/numpy type so ``str(qual)`` prints what the
    reference prints into ``<chrom>.score.txt``."""
    pending = pending_records(candidates, region, read_names, sig_types, sig_scores, class_scores, options)
    return [(q, f"{head}\t{gt}:{dr}:{dv}") for (q, head, cand, names) in pending
            for gt, dr, dv in (genotype(cand, names, options),)]


def pending_records(candidates: list, region: str, read_names: dict, sig_types: Sequence[str],
                    sig_scores: dict, class_scores: Sequence, options) -> list:
    """Everything of a region's records except the genotype: ``[(qual, line_without_sample_column,
    candidate, supporting_read_names), ...]``."""
    if not candidates:
        return []
    mean_score = np.mean(class_scores)                                # float32 in, float32 out
    class_penalty = (1 - round(mean_score, 2)) * 100                  # output.py:473-474
    contig, start, end = region.split("+")[:3]
    start, end = int(start), int(end)
    counts = Counter(sig_types)                                       # output.py:525-529
    flt = "Uncovered" if counts.get("sigUncovered", 0) >= 0.75 * len(sig_types) and "sigUncovered" in counts else "PASS"
    out = []
    for kinds_str, read_ids, bkps in candidates:
        support = len(read_ids)
        if support < options.min_support:                             # output.py:495-496
            continue
        names = [read_names[r] for r in read_ids]
        spread = np.std([int(sig_scores[r]) for r in read_ids]) / support          # output.py:551
        qual = min(100, spread + class_penalty)
        out.append(_record(contig, start, end, kinds_str.split("+"), bkps, support, names, qual, flt, options))
    return out


def _record(contig: str, start: int, end: int, kinds: list, bkps: list, support: int, names: list, qual,
            flt: str, options) -> tuple:
    """One pending record ``(qual, line_without_sample_column, candidate, supporting_read_names)``
    (output.py:553-598: type refinement, INFO fields, ALT)."""
    kinds, kb = refine_types(kinds, bkps, options.min_sv_size)
    info = [f"END={end}", f"SVLEN={end - start}", "SVTYPE=" + "+".join(kinds), f"SUPPORT={support}",
            "BKPS=" + ",".join(f"{k}:{b[2]}-{b[0]}-{b[1]}" for k, b in zip(kinds, kb))]
    if options.qname:
        info.append("READS=" + ",".join(names))
    alt = "<CSV>" if len(kinds) >= 2 else "<SV>"                      # output.py:573-576
    return (qual, "\t".join([contig, str(start), "0", "N", alt, str(qual), flt, ";".join(info), "GT:DR:DV"]),
            (contig, start, end, kinds), names)


def pending_records_native(table, labels: np.ndarray, win: np.ndarray, options) -> Optional[list]:
    """All pending records of a chromosome through ``svx_calls_aggregate`` (``csrc/host_calls.cpp``):
    the per-row loop, the aggregation and the score arithmetic run natively over the parser's byte
    spans, and Python only formats the candidates that reached ``min_support``.  Returns ``None`` when
    the native step declines (e.g. a signature score that only Python's ``int()`` accepts), so the
    caller can take the pure-Python route."""
    from . import _lib
    lib = _lib.load()
    n = len(table)
    if n == 0:
        return []
    import ctypes
    text, spans = table._text, np.ascontiguousarray(table._spans, dtype=np.int64)
    flags = np.ascontiguousarray(table.flags, dtype=np.int32)
    b0 = np.ascontiguousarray(table.bkp_start, dtype=np.int64)
    b1 = np.ascontiguousarray(table.bkp_end, dtype=np.int64)
    b2 = np.ascontiguousarray(table.bkp_len, dtype=np.int64)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    win = np.ascontiguousarray(win, dtype=np.float32)
    cand = np.empty((n, 25), dtype=np.int64)
    qual = np.empty(n, dtype=np.float64)
    reads = np.empty(n, dtype=np.int64)
    nc, nr = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.svx_calls_aggregate(text, len(text), n, spans.ctypes.data, flags.ctypes.data, b0.ctypes.data,
                                 b1.ctypes.data, b2.ctypes.data, labels.ctypes.data, win.ctypes.data,
                                 int(options.min_support), cand.ctypes.data, qual.ctypes.data, reads.ctypes.data,
                                 ctypes.byref(nc), ctypes.byref(nr))
    if rc != 0:
        return None
    out = []
    region_cache = (-1, None)
    cand_l = cand[:nc.value].tolist()
    all_names = table.strings_at("read_name", reads[:nr.value])        # decoded once, for the emitted candidates only
    region_of = dict(zip(*(lambda r: (r.tolist(), table.strings_at("region", r)))(np.unique(cand[:nc.value, 0]))))
    for c, q in zip(cand_l, qual[:nc.value]):                           # q stays numpy.float64: str() prints like the reference
        row, support, nk, uncovered, roff = c[:5]
        if region_cache[0] != row:
            contig, start, end = region_of[row].split("+")[:3]
            region_cache = (row, (contig, int(start), int(end)))
        contig, start, end = region_cache[1]
        kinds = [TYPE_NAMES[k] for k in c[5:5 + nk]]
        bkps = [c[10 + 3 * k:13 + 3 * k] for k in range(nk)]
        names = all_names[roff:roff + support]
        out.append(_record(contig, start, end, kinds, bkps, support, names, min(100, q),
                           "Uncovered" if uncovered else "PASS", options))
    return out


# ------------------------------------------------------------------------------------------------
# whole chromosome
# ------------------------------------------------------------------------------------------------
def call_chromosome(table, labels: np.ndarray, probs: np.ndarray, options, genotype,
                    aggregate: Callable = aggregate_region) -> List[Tuple[object, str]]:
    """The per-row loop of ``Predict.run`` (predict.py:213-300) followed, region by region, by
    aggregation and record assembly.  Returns ``[(qual, vcf_line), ...]`` in file order.

    Row rules kept from the reference: a row whose label columns contain ``complement`` is skipped
    (predict.py:214); a forward signature classified INV is dropped before it can open a region
    (predict.py:229-231); a region is flushed when the *next kept* row names another
    region (predict.py:235-247); a row of a non-main segment pair (id without ``m``) classified DEL
    or INS still contributes its score and metadata but no call (predict.py:279-281); later rows
    overwrite earlier ones with the same read id and class."""
    assert probs.dtype == np.float32, "scores must stay numpy.float32 (SURVEY §8(b))"
    n = len(table)
    pred_arr = np.asarray(labels).astype(np.int64)
    # round(np.float32, 2) of the winning class, elementwise the same ufunc as numpy.round
    win = np.round(probs[np.arange(n), pred_arr], 2) if n else np.zeros(0, np.float32)
    if aggregate is aggregate_region and getattr(table, "has_text", lambda: False)() and USE_NATIVE:
        records = pending_records_native(table, pred_arr, win, options)
        if records is not None:
            return _genotyped(records, genotype, options)
    pred = pred_arr.tolist()
    fwd_inv = (((table.forward == "True") & (pred_arr == 2)) | ((np.asarray(table.flags) & 16) != 0)).tolist() if n else []
    read_num, region_col, read_name = table.read_num.tolist(), table.region.tolist(), table.read_name.tolist()
    sig_type, sig_score, mech = table.sig_type.tolist(), table.sig_score.tolist(), table.mechanism.tolist()
    b0, b1, b2 = table.bkp_start.tolist(), table.bkp_end.tolist(), table.bkp_len.tolist()

    records: list = []
    reads: dict = {}
    names: dict = {}
    scores: dict = {}
    mechs: dict = {}
    types: list = []
    wins: list = []
    last = ""

    def flush():
        records.extend(pending_records(aggregate(reads), last, names, types, scores, wins, options))

    for i in range(n):
        if fwd_inv[i]:
            continue
        region = region_col[i]
        if region != last:
            if last != "":
                flush()
            last = region
            reads, names, scores, mechs, types, wins = {}, {}, {}, {}, [], []
        rn = read_num[i]
        main = "m" in rn
        key = rn.replace("m", "") if main else rn
        names[key] = read_name[i]
        types.append(sig_type[i])
        wins.append(win[i])
        scores[key] = sig_score[i]
        mechs[key] = mech[i]
        p = pred[i]
        if not main and p < 2:
            continue
        slot = reads.get(key)
        if slot is None:
            reads[key] = {p: [b0[i], b1[i], b2[i]]}
        else:
            slot[p] = [b0[i], b1[i], b2[i]]
    flush()                                                           # predict.py:298-300
    return _genotyped(records, genotype, options)


def _genotyped(records: list, genotype, options) -> List[Tuple[object, str]]:
    if hasattr(genotype, "genotype_many"):
        gts = genotype.genotype_many([r[2] for r in records], [r[3] for r in records], options)
    else:
        gts = [genotype(r[2], r[3], options) for r in records]
    return [(q, f"{head}\t{gt}:{dr}:{dv}") for (q, head, _, _), (gt, dr, dv) in zip(records, gts)]


def region_cuts(table, chunk_rows: int) -> List[int]:
    """Chunk boundaries ``[0, c1, ..., N]``: about ``chunk_rows`` rows each, cut where the region column
    changes (so that little is held back at a seam; exactness does not depend on it, see
    :func:`call_chromosome_streamed`)."""
    n = len(table)
    if n == 0:
        return [0]
    flags = getattr(table, "flags", None)
    if flags is not None and len(flags) == n:
        starts = np.flatnonzero((np.asarray(flags) & 8) == 0)            # SVX_BED_FLAG_SAME_REGION unset
    else:
        reg = table.region
        starts = np.concatenate([[0], 1 + np.flatnonzero(reg[1:] != reg[:-1])])
    cuts, want = [0], chunk_rows
    while want < n:
        k = int(np.searchsorted(starts, want))                            # first region start >= want
        if k >= len(starts):
            break
        c = int(starts[k])
        if c > cuts[-1]:
            cuts.append(c)
        want = c + chunk_rows
    cuts.append(n)
    return cuts


def _open_region_start(table, labels: np.ndarray) -> int:
    """First row of the region that is still OPEN at the end of ``table``: the earliest kept row of the
    trailing run of kept rows naming one region (dropped rows in between do not close it,
    predict.py:214,229-247).  ``len(table)`` if no row is kept."""
    flags = np.asarray(table.flags)
    kept = np.flatnonzero(~((((flags & 2) != 0) & (np.asarray(labels) == 2)) | ((flags & 16) != 0)))
    if kept.size == 0:
        return len(table)
    region = table.region
    last = region[kept[-1]]
    j = kept.size - 1
    while j > 0 and region[kept[j - 1]] == last:
        j -= 1
    return int(kept[j])


def call_chromosome_streamed(table, classify: Callable, options, genotype, chunk_rows: int = 65536,
                             aggregate: Callable = aggregate_region) -> List[Tuple[object, str]]:
    """:func:`call_chromosome` with the GPU and the host working at the same time: ``classify(rows) ->
    (labels, probs)`` of chunk k+1 runs on a worker thread (``svx_classify`` releases the GIL) while
    this thread turns chunk k into records.  The reference closes a region only when the next KEPT row
    names another one, and which rows are kept depends on the labels: the region still open at the end
    of a chunk is therefore held back and processed with the next chunk, so the records are exactly
    those of one ``call_chromosome`` over the whole table wherever the cuts fall."""
    from concurrent.futures import ThreadPoolExecutor
    cuts = region_cuts(table, max(int(chunk_rows), 1))
    spans = list(zip(cuts[:-1], cuts[1:]))
    if len(spans) <= 1:
        labels, probs = classify(table.rows)
        return call_chromosome(table, labels, probs, options, genotype, aggregate)
    records: list = []
    start = 0                                    # rows [start, a) of earlier chunks are still pending
    held_l = np.zeros(0, np.int32)
    held_p = np.zeros((0, 5), np.float32)
    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = pool.submit(classify, np.ascontiguousarray(table.rows[spans[0][0]:spans[0][1]]))
        for k, (a, b) in enumerate(spans):
            t_wait = time.perf_counter()
            labels, probs = pending.result()
            t_host = time.perf_counter()
            if k + 1 < len(spans):
                na, nb = spans[k + 1]
                pending = pool.submit(classify, np.ascontiguousarray(table.rows[na:nb]))
            labels = np.concatenate([held_l, np.asarray(labels)])
            probs = np.concatenate([held_p, np.asarray(probs)])
            sub = table.take(slice(start, b))
            cut = len(sub) if k + 1 == len(spans) else _open_region_start(sub, labels)
            if cut > 0:
                records.extend(call_chromosome(sub.take(slice(0, cut)) if cut < len(sub) else sub,
                                               labels[:cut], probs[:cut], options, genotype, aggregate))
            held_l, held_p = labels[cut:], probs[cut:]
            start += cut
            logging.debug("chunk %d: rows [%d, %d): waited %.3f s for the classifier, host %.3f s", k, a, b,
                          t_host - t_wait, time.perf_counter() - t_host)
    return records


def write_chromosome(out_path_prefix: str, records: List[Tuple[object, str]]) -> None:
    """``<prefix>.vcf`` and ``<prefix>.score.txt`` as ``merge_split_vcfs`` (output.py:307) and
    ``cal_scores_max_min`` (output.py:601-612) expect them."""
    with open(out_path_prefix + ".vcf", "w") as vcf, open(out_path_prefix + ".score.txt", "w") as sc:
        for qual, line in records:
            sc.write(str(qual) + "\n")
            vcf.write(line + "\n")
