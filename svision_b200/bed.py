"""Vectorised reader of SVision's ``<chrom>.segments.all.bed``.

Replaces ``BatchGenerator.read_class_list`` + the per-image token parsing of ``next_batch``
(reference ``src/network/create_batch.py:29-61,103-137``): 23 tab-separated columns

    0 region | 1-5 seg1 (xS xE yS yE fwd) | 6-10 seg2 | 11 read_len | 12 ref_len | 13 read id |
    14 sub id (dropped by the reference label, create_batch.py:48) | 15 qname | 16 sig type |
    17-18 bkp start/end | 19 non-linear score | 20 forward flag | 21 mechanism | 22 bkp len

(writer: ``src/collection/output_clusters.py:180-182,207-209``).  Columns 1-12 become the packed
``int32[N,12]`` rows the GPU path consumes; the remaining columns are kept as string columns for
the per-row replay of ``src/network/predict.py:213-300``.  No padding rows are appended: the
reference pads to a multiple of ``batch_size`` only because its TF placeholder has a fixed batch
dimension (``create_batch.py:54-59``, ``predict.py:167``)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

N_COLS = 23
_INT_COLS = (1, 2, 3, 4, 6, 7, 8, 9, 11, 12)


@dataclass
class SegmentsTable:
    rows: np.ndarray            # int32[N,12]
    read_num: np.ndarray        # col 13 (object array of str)
    region: np.ndarray          # col 0
    read_name: np.ndarray       # col 15
    sig_type: np.ndarray        # col 16
    bkp_start: np.ndarray       # col 17, int64
    bkp_end: np.ndarray         # col 18, int64
    sig_score: np.ndarray       # col 19 (str, passed through verbatim)
    forward: np.ndarray         # col 20 (str 'True'/'False')
    mechanism: np.ndarray       # col 21
    bkp_len: np.ndarray         # col 22, int64

    def __len__(self) -> int:
        return self.rows.shape[0]

    def label_strings(self) -> list:
        """The reference's per-row label strings (create_batch.py:48), for compatibility."""
        sep = "svision"
        return [sep.join([str(self.read_num[i]), str(self.region[i]), str(self.read_name[i]),
                          str(self.sig_type[i]), str(self.bkp_start[i]), str(self.bkp_end[i]),
                          str(self.sig_score[i]), str(self.forward[i]), str(self.mechanism[i]),
                          str(self.bkp_len[i])]) for i in range(len(self))]


def read_segments_bed(path: str) -> SegmentsTable:
    import pandas as pd
    try:
        df = pd.read_csv(path, sep="\t", header=None, dtype=str, keep_default_na=False,
                         quoting=3, engine="c")
    except pd.errors.EmptyDataError:
        df = pd.DataFrame({i: [] for i in range(N_COLS)}, dtype=str)
    if df.shape[1] < N_COLS:
        raise ValueError(f"{path}: expected {N_COLS} tab-separated columns, found {df.shape[1]}")
    n = df.shape[0]
    rows = np.empty((n, 12), dtype=np.int64)
    order = (1, 2, 3, 4, None, 6, 7, 8, 9, None, 11, 12)
    for j, c in enumerate(order):
        if c is not None:
            rows[:, j] = pd.to_numeric(df[c], downcast=None).to_numpy(dtype=np.int64)
    # 'True' -> forward; 'False' and anything else -> the reverse branch (create_batch.py:111-116)
    rows[:, 4] = (df[5].to_numpy() == "True").astype(np.int64)
    rows[:, 9] = (df[10].to_numpy() == "True").astype(np.int64)
    lim = np.iinfo(np.int32)
    if n and (rows.min() < lim.min or rows.max() > lim.max):
        raise OverflowError(f"{path}: coordinate does not fit int32")
    to_i64 = lambda c: pd.to_numeric(df[c]).to_numpy(dtype=np.int64)  # noqa: E731
    return SegmentsTable(
        rows=np.ascontiguousarray(rows.astype(np.int32)),
        read_num=df[13].to_numpy(dtype=object), region=df[0].to_numpy(dtype=object),
        read_name=df[15].to_numpy(dtype=object), sig_type=df[16].to_numpy(dtype=object),
        bkp_start=to_i64(17), bkp_end=to_i64(18), sig_score=df[19].to_numpy(dtype=object),
        forward=df[20].to_numpy(dtype=object), mechanism=df[21].to_numpy(dtype=object),
        bkp_len=to_i64(22))
