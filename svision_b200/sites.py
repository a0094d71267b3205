"""Synthetic candidate-site generators (sets P1 and P2).

A *site* is one row of SVision's ``<chrom>.segments.all.bed`` (reference writer
``src/collection/output_clusters.py:180-182,207-209``; reader
``src/network/create_batch.py:45,103-140``): a pair of read-vs-reference segments plus two
lengths.  The packed form used everywhere in this package is ``int32[N, 12]``::

    xS1, xE1, yS1, yE1, f1,  xS2, xE2, yS2, yE2, f2,  len_a, len_b

with ``f`` = 1 for the token ``'True'`` and 0 for anything else (``'False'`` and invalid tokens
both take the reverse branch: ``src/segmentplot/classes.py:50-53``,
``src/segmentplot/plot_segment.py:45-52``).  ``xE`` is parsed but never used by the reference
(``create_batch.py:106,121``).
"""
from __future__ import annotations

import numpy as np

ROW_FIELDS = 12
#: the reference's padding row ``'0_1_0_1_True_1_1_1_1_True_2_2'`` (create_batch.py:55)
PAD_ROW = np.array([0, 1, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2], dtype=np.int32)

SEED_CONFIG2 = 20261017   # 10 k sites, 1 GPU
SEED_CONFIG3 = 20261018   # 100 k sites, HiFi profile
SEED_CONFIG4 = 20261019   # 500 k sites stream
SEED_CONFIG5 = 20261020   # ONT profile
SEED_P2 = 7


def make_sites_p1(n: int, seed: int = SEED_CONFIG2, profile: str = "hifi") -> np.ndarray:
    """Set P1: realistic INS / DEL / minor-segment rows following the 2x-gap flank normalisation
    of ``src/collection/analyze_reads.py:92-98`` as observed on the demo rows."""
    rng = np.random.default_rng(seed)
    if profile == "hifi":
        u_hi, mix = 4.0, (0.40, 0.40, 0.07, 0.13)
    elif profile in ("ont", "contig"):
        u_hi, mix = 5.5, (0.30, 0.30, 0.15, 0.25)
    else:
        raise ValueError(f"unknown profile {profile!r}")
    t = rng.choice(4, size=n, p=mix)
    s = np.floor(10.0 ** rng.uniform(1.7, u_hi, size=n)).astype(np.int64)
    j = rng.integers(-3, 4, size=n)
    base_is_ins = np.where(t == 0, True, np.where(t == 1, False, rng.random(n) < 0.5))
    q = np.floor(rng.random(n) * (5 * s + 1)).astype(np.int64)
    a = np.floor(rng.random(n) * (6 * s + 1)).astype(np.int64) - s
    ell = np.floor(rng.random(n) * (2 * s + 1)).astype(np.int64)

    rows = np.zeros((n, ROW_FIELDS), dtype=np.int64)
    # seg1 is the same for INS and DEL
    rows[:, 0] = 0
    rows[:, 1] = 2 * s
    rows[:, 2] = 0
    rows[:, 3] = 2 * s
    rows[:, 4] = 1
    ins_seg2 = np.stack([3 * s + j, 5 * s, 2 * s + 1, 4 * s + 1, np.ones_like(s)], axis=1)
    del_seg2 = np.stack([2 * s + j, 4 * s, 3 * s + j, 5 * s, np.ones_like(s)], axis=1)
    rows[:, 5:10] = np.where(base_is_ins[:, None], ins_seg2, del_seg2)
    rows[:, 10] = np.where(base_is_ins, 5 * s + j, 4 * s)
    rows[:, 11] = np.where(base_is_ins, 4 * s, 5 * s + j)
    fwd_minor = t == 2
    rev_minor = t == 3
    rows[fwd_minor, 5:10] = np.stack([q, q + ell, a, a + ell, np.ones_like(s)], axis=1)[fwd_minor]
    rows[rev_minor, 5:10] = np.stack([q + ell, q, a, a + ell, np.zeros_like(s)], axis=1)[rev_minor]
    return rows.astype(np.int32)


def make_sites_p2(n: int, seed: int = SEED_P2) -> np.ndarray:
    """Set P2 (parity stress): half unconstrained integer fuzz (negatives, out-of-frame, L<=0,
    ratio<1), half P1 rows with a wider size range."""
    rng = np.random.default_rng(seed)
    n_fuzz = n // 2
    h = rng.choice(np.array([5, 300, 5000, 10 ** 6]), size=n_fuzz)
    lo = -(h // 4)
    rows = np.empty((n_fuzz, ROW_FIELDS), dtype=np.int64)
    for c in range(ROW_FIELDS):
        rows[:, c] = np.floor(rng.random(n_fuzz) * (h - lo + 1)).astype(np.int64) + lo
    rows[:, 4] = rng.integers(0, 2, size=n_fuzz)
    rows[:, 9] = rng.integers(0, 2, size=n_fuzz)
    rows[:, 10] = np.floor(rng.random(n_fuzz) * (h + 1)).astype(np.int64)
    rows[:, 11] = np.floor(rng.random(n_fuzz) * (h + 1)).astype(np.int64)
    n_p1 = n - n_fuzz
    # P1-style with U ~ Uniform(1.5, 5.5): reuse the ONT mix, then widen by rescaling sizes
    p1 = make_sites_p1(n_p1, seed=seed + 1, profile="ont").astype(np.int64)
    out = np.concatenate([rows, p1], axis=0)
    perm = rng.permutation(n)
    return out[perm].astype(np.int32)


def edge_case_sites() -> np.ndarray:
    """Hand-written rows the parity tests always include: the pad row, all-zero, L<=0, ratio<1,
    fully out-of-frame, and lines that cross every frame edge in both drawing orders."""
    r = [
        PAD_ROW.tolist(),
        [0] * 12,
        [0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0],
        [5, 0, 5, 5, 1, 9, 0, 9, 8, 0, 10, 10],             # L = 0 and L = -1
        [0, 0, 0, 227, 1, 226, 0, 0, 227, 0, 227, 227],       # full diagonals, ratio == 1
        [0, 0, 0, 228, 1, 227, 0, 0, 228, 0, 228, 228],       # ratio just above 1
        [0, 0, 0, 100, 1, 50, 0, 20, 60, 0, 100, 50],         # ratio < 1 -> clamped to 1
        [-50, 0, -80, 400, 1, 500, 0, -100, 600, 0, 300, 300],
        [1000, 0, 1000, 1500, 1, 2000, 0, 3000, 3500, 0, 300, 300],   # fully outside
        [-300, 0, 100, 900, 1, 900, 0, -300, 700, 0, 454, 454],
        [100, 0, -300, 900, 1, 100, 0, -300, 700, 0, 454, 454],
        [0, 0, 0, 250000000, 1, 250000000, 0, 0, 250000000, 0, 250000000, 250000000],
        [0, 0, 0, 2000, 1, 1000, 0, 1000, 3000, 1, 5000, 4000],
        [0, 0, 0, 2000, 1, 2000, 0, 0, 2000, 1, 2000, 2000],  # two identical-column segments
    ]
    return np.array(r, dtype=np.int32)


def rows_to_bed_lines(rows: np.ndarray, region_size: int = 19) -> list[str]:
    """Render packed rows as 23-column BED text in the reference's format
    (``src/collection/output_clusters.py:180-182``) so the reference reader can ingest them.
    Column meanings: SURVEY.md §8(f)#1."""
    lines = []
    for i, r in enumerate(np.asarray(rows)):
        reg = i // region_size
        region = f"chr1+{1000 * reg}+{1000 * reg + 500}+{region_size}"
        f1 = "True" if r[4] == 1 else "False"
        f2 = "True" if r[9] == 1 else "False"
        cols = [region,
                str(r[0]), str(r[1]), str(r[2]), str(r[3]), f1,
                str(r[5]), str(r[6]), str(r[7]), str(r[8]), f2,
                str(r[10]), str(r[11]),
                f"{i % region_size}m", "0", f"read{i}", "INS",
                str(1000 * reg + 100), str(1000 * reg + 200), "0.5", f2, "NA", "100"]
        lines.append("\t".join(cols))
    return lines


def make_region_table(n_rows: int, seed: int = SEED_CONFIG4, profile: str = "hifi", contig: str = "chr1"):
    """Config-4 style stream: P1 rows grouped into synthetic regions of 19±8 rows
    with the metadata columns the reference's per-row loop consumes (``src/network/predict.py:218-226``;
    BED columns 13-22, writer ``src/collection/output_clusters.py:180-182,207-209``).  Every read has a
    main row (id ``<k>m``); ≈12 % of reads add one or two rows of a main×minor pair (id ``<k>``),
    ≈8 % of regions are mostly ``sigUncovered``.  Returns a :class:`svision_b200.bed.SegmentsTable`."""
    from .bed import SegmentsTable
    rng = np.random.default_rng(seed + 1)
    rows = make_sites_p1(n_rows, seed=seed, profile=profile)
    col = {k: [] for k in ("read_num", "region", "read_name", "sig_type", "bkp_start", "bkp_end",
                           "sig_score", "forward", "mechanism", "bkp_len")}
    pos, region_no, n = 100_000, 0, 0
    mechanisms = ("None", "NHEJ+0", "NHEJ+1", "FoSTeS+2")
    while n < n_rows:
        size = int(np.clip(rng.normal(19, 8), 1, 60))
        width = int(10 ** rng.uniform(1.7, 4.0))
        start, end = pos, pos + width
        region = f"{contig}+{start}+{end}+{int(rng.integers(10, 60))}"
        uncovered = rng.random() < 0.08
        base_len = int(rng.integers(50, width + 51))
        k = 0
        emitted = 0
        while emitted < size and n < n_rows:
            k += 1
            name = f"m{region_no}/{k}/ccs"
            score = int(rng.integers(0, 40))
            n_minor = int(rng.choice([0, 1, 2], p=[0.88, 0.07, 0.05]))
            for sub in range(1 + n_minor):
                if emitted >= size or n >= n_rows:
                    break
                main = sub == 0
                col["read_num"].append(f"{k}m" if main else f"{k}")
                col["region"].append(region)
                col["read_name"].append(name)
                col["sig_type"].append("sigUncovered" if (uncovered and rng.random() < 0.85) else "sigGap")
                jit = int(rng.integers(-6, 7))
                col["bkp_start"].append(start + jit if main else start + int(rng.integers(0, width)))
                col["bkp_end"].append(start + jit + 1 if main else end + int(rng.integers(-8, 9)))
                col["sig_score"].append(str(score))
                col["forward"].append("True" if (main or rng.random() < 0.4) else "False")
                col["mechanism"].append(mechanisms[int(rng.integers(0, len(mechanisms)))])
                col["bkp_len"].append(base_len + int(rng.integers(-3, 4)))
                emitted += 1
                n += 1
        pos = end + int(rng.integers(2_000, 50_000))
        region_no += 1
    obj = lambda k: np.array(col[k], dtype=object)                                   # noqa: E731
    i64 = lambda k: np.array(col[k], dtype=np.int64)                                 # noqa: E731
    return SegmentsTable(rows, i64("bkp_start"), i64("bkp_end"), i64("bkp_len"), read_num=obj("read_num"),
                         region=obj("region"), read_name=obj("read_name"), sig_type=obj("sig_type"),
                         sig_score=obj("sig_score"), forward=obj("forward"), mechanism=obj("mechanism"))


def table_to_bed_lines(table) -> list[str]:
    """23-column BED text of a :class:`SegmentsTable` (inverse of ``bed.read_segments_bed``)."""
    r = table.rows
    tf = ("False", "True")
    return ["\t".join([str(table.region[i]), str(r[i, 0]), str(r[i, 1]), str(r[i, 2]), str(r[i, 3]), tf[int(r[i, 4] == 1)],
                       str(r[i, 5]), str(r[i, 6]), str(r[i, 7]), str(r[i, 8]), tf[int(r[i, 9] == 1)],
                       str(r[i, 10]), str(r[i, 11]), str(table.read_num[i]), "1", str(table.read_name[i]),
                       str(table.sig_type[i]), str(table.bkp_start[i]), str(table.bkp_end[i]),
                       str(table.sig_score[i]), str(table.forward[i]), str(table.mechanism[i]),
                       str(table.bkp_len[i])]) for i in range(len(table))]


def make_alignments(table, seed: int = 5, depth: int = 30, contig_length: int = 250_000_000):
    """Synthetic coordinate-sorted alignments around the regions of ``table`` for the genotyping step
    (reference ``src/network/genotype.py:17-73``): per region ``depth`` reads with random extents around
    the window, the region's own supporting read names among them, and a few secondary / unmapped /
    low-quality records.  Returns a dict of columns (``reference_start``, ``reference_end``,
    ``mapping_quality``, ``is_unmapped``, ``is_secondary``, ``query_name``, ``contig_length``)."""
    rng = np.random.default_rng(seed)
    regions, first = np.unique(np.asarray(table.region, dtype=object).astype(str), return_index=True)
    names_by_region: dict = {}
    for reg, nm in zip(table.region.tolist(), table.read_name.tolist()):
        names_by_region.setdefault(reg, []).append(nm)
    s, e, q, u, sec, names = [], [], [], [], [], []
    for reg in regions[np.argsort(first)].tolist():
        _, a, b = reg.split("+")[:3]
        a, b = int(a), int(b)
        own = list(dict.fromkeys(names_by_region[reg]))
        for j in range(depth):
            left = a - int(rng.integers(-200, 6000))
            right = b + int(rng.integers(-200, 6000))
            if right <= left:
                right = left + 50
            s.append(max(0, left))
            e.append(right)
            q.append(int(rng.choice([0, 5, 20, 60], p=[0.05, 0.05, 0.2, 0.7])))
            u.append(bool(rng.random() < 0.02))
            sec.append(bool(rng.random() < 0.05))
            # a third of the records belong to supporting reads (excluded from the reference count);
            # some names repeat (supplementary pieces of one read)
            names.append(own[int(rng.integers(0, len(own)))] if rng.random() < 0.33
                         else f"bg{a}/{int(rng.integers(0, depth * 2 // 3))}")
    order = np.argsort(np.array(s, dtype=np.int64), kind="stable")
    take = lambda x, dt: np.array(x, dtype=dt)[order]                                # noqa: E731
    return dict(reference_start=take(s, np.int64), reference_end=take(e, np.int64),
                mapping_quality=take(q, np.int64), is_unmapped=take(u, bool), is_secondary=take(sec, bool),
                query_name=take(names, object), contig_length=contig_length)
