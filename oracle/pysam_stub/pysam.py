"""ORACLE tooling -- a minimal pure-Python stand-in for the parts of ``pysam`` that the reference's
collection stage touches, so that ``src/collection`` can run on the demo BAM
in the build container, where pysam/htslib are not installable.  It exists only to derive
*realistic encoder inputs* (``oracle/make_demo_rows.py``); it is not product code and it is not a
faithful pysam: FASTA access returns a deterministic pseudo-sequence (no GRCh38 here), so the
left-shifting of breakpoints differs from a real run and the rows are not a VCF golden."""
from __future__ import annotations

import gzip
import re
import struct

_CIGAR_OPS = "MIDNSHP=X"
_SEQ_CODE = "=ACMGRSVTWYHKDBN"


class AlignedSegment:
    def __init__(self):
        self.query_name = None
        self.flag = 0
        self.reference_id = -1
        self.reference_start = 0
        self.mapping_quality = 0
        self.cigarstring = None
        self.next_reference_id = -1
        self.next_reference_start = -1
        self.query_sequence = None
        self._reference_name = None

    # ---- flags -------------------------------------------------------------------------------
    def _getf(self, bit):
        return bool(self.flag & bit)

    def _setf(self, bit, v):
        self.flag = (self.flag | bit) if v else (self.flag & ~bit)

    is_unmapped = property(lambda s: s._getf(0x4), lambda s, v: s._setf(0x4, v))
    is_reverse = property(lambda s: s._getf(0x10), lambda s, v: s._setf(0x10, v))
    is_secondary = property(lambda s: s._getf(0x100), lambda s, v: s._setf(0x100, v))
    is_supplementary = property(lambda s: s._getf(0x800), lambda s, v: s._setf(0x800, v))

    # ---- aliases -----------------------------------------------------------------------------
    qname = property(lambda s: s.query_name, lambda s, v: setattr(s, "query_name", v))
    mapq = property(lambda s: s.mapping_quality, lambda s, v: setattr(s, "mapping_quality", v))

    @property
    def reference_name(self):
        return self._reference_name

    # ---- CIGAR-derived -------------------------------------------------------------------------
    def _ops(self):
        return [(int(n), op) for n, op in re.findall(r"(\d+)([MIDNSHP=X])", self.cigarstring or "")]

    @property
    def reference_end(self):
        return self.reference_start + sum(n for n, op in self._ops() if op in "MDN=X")

    @property
    def query_alignment_start(self):
        start = 0
        for n, op in self._ops():
            if op == "H":
                continue
            if op == "S":
                start += n
            else:
                break
        return start

    @property
    def query_alignment_end(self):
        return self.query_alignment_start + sum(n for n, op in self._ops() if op in "MI=X")

    @property
    def query_length(self):
        if self.query_sequence:
            return len(self.query_sequence)
        return sum(n for n, op in self._ops() if op in "MIS=X")


class AlignmentFile:
    def __init__(self, path, mode="rb"):
        with gzip.open(path, "rb") as f:          # BGZF is a series of gzip members
            data = f.read()
        assert data[:4] == b"BAM\x01", "not a BAM file"
        l_text = struct.unpack_from("<i", data, 4)[0]
        self.text = data[8:8 + l_text].decode(errors="replace")
        pos = 8 + l_text
        n_ref = struct.unpack_from("<i", data, pos)[0]
        pos += 4
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<i", data, pos)[0]
            pos += 4
            self.references.append(data[pos:pos + l_name - 1].decode())
            pos += l_name
            self.lengths.append(struct.unpack_from("<i", data, pos)[0])
            pos += 4
        self._records = []
        while pos + 4 <= len(data):
            block = struct.unpack_from("<i", data, pos)[0]
            pos += 4
            rec = data[pos:pos + block]
            pos += block
            self._records.append(self._parse(rec))
        so = re.search(r"@HD.*?SO:(\S+)", self.text)
        self.header = {"HD": {"SO": so.group(1) if so else "unknown"}}

    def _parse(self, rec):
        ref_id, p, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos, _tlen = struct.unpack_from(
            "<iiBBHHHiiii", rec, 0)
        off = 32
        a = AlignedSegment()
        a.query_name = rec[off:off + l_name - 1].decode()
        off += l_name
        cig = struct.unpack_from(f"<{n_cig}I", rec, off)
        off += 4 * n_cig
        a.cigarstring = "".join(f"{c >> 4}{_CIGAR_OPS[c & 15]}" for c in cig) if n_cig else None
        nb = (l_seq + 1) // 2
        sb = rec[off:off + nb]
        seq = []
        for i in range(l_seq):
            b = sb[i >> 1]
            seq.append(_SEQ_CODE[(b >> 4) if not (i & 1) else (b & 15)])
        a.query_sequence = "".join(seq) if l_seq else None
        a.flag, a.reference_id, a.reference_start, a.mapping_quality = flag, ref_id, p, mapq
        a.next_reference_id, a.next_reference_start = nref, npos
        a._reference_name = self.references[ref_id] if ref_id >= 0 else None
        return a

    def fetch(self, contig=None, start=None, stop=None):
        tid = self.references.index(contig) if contig is not None else None
        for a in self._records:
            if a.is_unmapped or a.reference_id < 0:
                continue
            if tid is not None and a.reference_id != tid:
                continue
            if start is not None and a.reference_end <= start:
                continue
            if stop is not None and a.reference_start >= stop:
                continue
            yield a

    def get_tid(self, name):
        return self.references.index(name) if name in self.references else -1

    def getrname(self, tid):
        return self.references[tid]

    get_reference_name = getrname

    def get_reference_length(self, name):
        return self.lengths[self.references.index(name)]

    def check_index(self):
        return True

    def close(self):
        pass


class FastaFile:
    """Deterministic pseudo-genome: base at position p of contig c depends only on (c, p)."""

    def __init__(self, path):
        self.references = []
        try:
            for line in open(path + ".fai"):
                self.references.append(line.split("\t")[0])
        except OSError:
            pass

    def fetch(self, contig, start, end):
        h = sum(map(ord, contig)) * 2654435761
        return "".join("ACGT"[((h ^ (p * 2246822519)) >> 13) & 3] for p in range(max(start, 0), max(end, 0)))

    def get_reference_length(self, name):
        return 250_000_000


class VariantFile:                                  # only referenced under --graph
    def __init__(self, *a, **k):
        raise NotImplementedError("VariantFile is not available in the pysam stand-in")
