"""Reader of SVision's ``<chrom>.segments.all.bed``.

Replaces ``BatchGenerator.read_class_list`` + the per-image token parsing of ``next_batch``
(reference ``src/network/create_batch.py:29-61,103-137``): 23 tab-separated columns

    0 region | 1-5 seg1 (xS xE yS yE fwd) | 6-10 seg2 | 11 read_len | 12 ref_len | 13 read id |
    14 sub id (dropped by the reference label, create_batch.py:48) | 15 qname | 16 sig type |
    17-18 bkp start/end | 19 non-linear score | 20 forward flag | 21 mechanism | 22 bkp len

(writer: ``src/collection/output_clusters.py:180-182,207-209``).  The whole file is parsed in one
native pass (``svx_bed_parse`` in ``csrc/host_bed.cpp``, C-ABI in ``include/svx.h``): columns 1-12
become the packed ``int32[N,12]`` rows the GPU path consumes, columns 17/18/22 an ``int64[N,3]``,
and the string columns are kept as byte spans into the text and only turned into Python strings
when the per-row replay of ``src/network/predict.py:213-300`` asks for them.  No padding rows are
appended: the reference pads to a multiple of ``batch_size`` only because its TF placeholder has a
fixed batch dimension (``create_batch.py:54-59``, ``predict.py:167``)."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

N_COLS = 23
FLAG_MAIN, FLAG_FORWARD, FLAG_UNCOVERED, FLAG_SAME_REGION, FLAG_COMPLEMENT = 1, 2, 4, 8, 16     # include/svx.h
_SPAN_OF = {"region": 0, "read_num": 1, "read_name": 2, "sig_type": 3, "sig_score": 4, "forward": 5,
            "mechanism": 6}


class SegmentsTable:
    """Columns of a segments BED.  ``rows`` (int32[N,12]), ``bkp_start`` / ``bkp_end`` / ``bkp_len``
    (int64[N]) and ``flags`` (int32[N]) are arrays; the string columns ``region``, ``read_num``,
    ``read_name``, ``sig_type``, ``sig_score``, ``forward``, ``mechanism`` are object arrays of ``str``,
    either given directly or decoded on first use from ``(text, spans)``."""

    STRING_COLUMNS = tuple(_SPAN_OF)

    def __init__(self, rows, bkp_start, bkp_end, bkp_len, flags=None, text: bytes = None, spans=None, **strings):
        self.rows = rows
        self.bkp_start, self.bkp_end, self.bkp_len = bkp_start, bkp_end, bkp_len
        self._text, self._spans = text, spans
        self._cols = {k: np.asarray(v, dtype=object) for k, v in strings.items()}
        unknown = set(self._cols) - set(_SPAN_OF)
        if unknown:
            raise TypeError(f"unknown columns {sorted(unknown)}")
        if text is None and set(self._cols) != set(_SPAN_OF):
            raise TypeError("without (text, spans) every string column must be given")
        self.flags = flags if flags is not None else self._flags_from_strings()

    def _flags_from_strings(self) -> np.ndarray:
        n = len(self)
        fl = np.zeros(n, dtype=np.int32)
        if n:
            fl |= np.array(["m" in s for s in self.read_num.tolist()], dtype=np.int32) * FLAG_MAIN
            fl |= (self.forward == "True").astype(np.int32) * FLAG_FORWARD
            fl |= (self.sig_type == "sigUncovered").astype(np.int32) * FLAG_UNCOVERED
            fl[1:] |= (self.region[1:] == self.region[:-1]).astype(np.int32) * FLAG_SAME_REGION
            # predict.py:214 skips a row when its joined label contains 'complement' anywhere
            comp = np.zeros(n, dtype=bool)
            for col in (self.region, self.read_num, self.read_name, self.sig_type, self.sig_score,
                        self.forward, self.mechanism):
                comp |= np.char.find(np.asarray(col, dtype=str), "complement") >= 0
            fl |= comp.astype(np.int32) * FLAG_COMPLEMENT
        return fl

    def __len__(self) -> int:
        return self.rows.shape[0]

    def __getattr__(self, name):
        if name not in _SPAN_OF:
            raise AttributeError(name)
        col = self._cols.get(name)
        if col is None:
            col = self._cols[name] = self._decode(_SPAN_OF[name])
        return col

    def _decode(self, k: int) -> np.ndarray:
        off = self._spans[:, k, 0].tolist()
        end = (self._spans[:, k, 0] + self._spans[:, k, 1]).tolist()
        try:
            text = self._text.decode("ascii")                         # byte offsets == str offsets
            vals = [text[a:b] for a, b in zip(off, end)]
        except UnicodeDecodeError:
            raw = self._text
            vals = [raw[a:b].decode("utf-8", errors="replace") for a, b in zip(off, end)]
        out = np.empty(len(vals), dtype=object)
        out[:] = vals
        return out

    def take(self, index) -> "SegmentsTable":
        """Row subset (slice or index array).  A parsed table keeps pointing into its text (string
        columns stay lazy); otherwise the string columns are materialised."""
        if self._text is not None:
            return SegmentsTable(self.rows[index], self.bkp_start[index], self.bkp_end[index], self.bkp_len[index],
                                 flags=self.flags[index], text=self._text, spans=self._spans[index],
                                 **{k: v[index] for k, v in self._cols.items()})
        return SegmentsTable(self.rows[index], self.bkp_start[index], self.bkp_end[index], self.bkp_len[index],
                             **{k: getattr(self, k)[index] for k in _SPAN_OF})

    def has_text(self) -> bool:
        """True for a table parsed from BED text (byte spans available to the native host steps)."""
        return self._text is not None and self._spans is not None

    def strings_at(self, name: str, rows) -> list:
        """Cells of a string column at ``rows`` (index array) as a list, decoding only those."""
        col = self._cols.get(name)
        if col is not None:
            return [str(x) for x in col[rows].tolist()]
        sp = self._spans[rows, _SPAN_OF[name]]
        text = self._text
        return [text[o:o + n].decode() for o, n in zip(sp[:, 0].tolist(), sp[:, 1].tolist())]

    def string_at(self, name: str, row: int) -> str:
        """One cell of a string column without decoding the whole column."""
        col = self._cols.get(name)
        if col is not None:
            return str(col[row])
        off, ln = self._spans[row, _SPAN_OF[name]]
        return self._text[off:off + ln].decode()

    def label_strings(self) -> list:
        """The reference's per-row label strings (create_batch.py:48), for compatibility."""
        sep = "svision"
        return [sep.join([str(self.read_num[i]), str(self.region[i]), str(self.read_name[i]),
                          str(self.sig_type[i]), str(self.bkp_start[i]), str(self.bkp_end[i]),
                          str(self.sig_score[i]), str(self.forward[i]), str(self.mechanism[i]),
                          str(self.bkp_len[i])]) for i in range(len(self))]


def parse_segments_bed(text: bytes) -> SegmentsTable:
    """Parse the content of a segments BED.  Raises ``ValueError`` naming the offending line on
    malformed input (too few columns, a non-integer or out-of-int32 coordinate)."""
    lib = _lib.load()
    if not isinstance(text, (bytes, bytearray)):
        raise TypeError("parse_segments_bed takes the file content as bytes")
    text = bytes(text)
    n = ctypes.c_int64(0)
    _lib.check(lib.svx_bed_count_rows(text, len(text), ctypes.byref(n)), "svx_bed_count_rows")
    n = n.value
    rows = np.empty((n, 12), dtype=np.int32)
    bkp = np.empty((n, 3), dtype=np.int64)
    spans = np.empty((n, len(_SPAN_OF), 2), dtype=np.int64)
    flags = np.empty(n, dtype=np.int32)
    rc = lib.svx_bed_parse(text, len(text), n, rows.ctypes.data, bkp.ctypes.data, spans.ctypes.data, flags.ctypes.data)
    if rc != 0:
        raise ValueError(lib.svx_last_error().decode(errors="replace"))
    return SegmentsTable(rows, bkp[:, 0].copy(), bkp[:, 1].copy(), bkp[:, 2].copy(), flags=flags, text=text, spans=spans)


def read_segments_bed(path: str) -> SegmentsTable:
    with open(path, "rb") as f:
        return parse_segments_bed(f.read())


def read_segments_bed_pandas(path: str) -> SegmentsTable:
    """The earlier pandas reader, kept as an independent cross-check of the native parser (tests)."""
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
    rows[:, 4] = (df[5].to_numpy() == "True").astype(np.int64)
    rows[:, 9] = (df[10].to_numpy() == "True").astype(np.int64)
    lim = np.iinfo(np.int32)
    if n and (rows.min() < lim.min or rows.max() > lim.max):
        raise OverflowError(f"{path}: coordinate does not fit int32")
    to_i64 = lambda c: pd.to_numeric(df[c]).to_numpy(dtype=np.int64)  # noqa: E731
    return SegmentsTable(
        np.ascontiguousarray(rows.astype(np.int32)), to_i64(17), to_i64(18), to_i64(22),
        read_num=df[13].to_numpy(dtype=object), region=df[0].to_numpy(dtype=object),
        read_name=df[15].to_numpy(dtype=object), sig_type=df[16].to_numpy(dtype=object),
        sig_score=df[19].to_numpy(dtype=object), forward=df[20].to_numpy(dtype=object),
        mechanism=df[21].to_numpy(dtype=object))
