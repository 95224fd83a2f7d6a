"""Pure-Python BGZF / BAM / BAI writer and reader (zlib only).

The reference never writes its own test inputs: its counters are tested in
external callers on real STAR BAMs (reference src/lib.rs:300-301) and no
samtools/pysam/htslib exists in this environment, so every input the parity
tests and the benchmark use is produced here, following the SAM/BAM
specification (layouts restated in SURVEY.md Appendix C):

* BGZF: gzip members with the `BC` extra subfield, payload <= 0xff00 bytes,
  htslib-style "a record never straddles two blocks unless it is larger than a
  block" (`split_records=False`), or plain streaming (`split_records=True`,
  what htsjdk does) to exercise the record framer's general path.
* BAM: header + records, little endian.
* BAI: binning index + 16 kb linear index + pseudo-bin 37450.

This module is host-side tooling (test inputs, the small-case generator); the
bench-scale generator is tools/bamgen.c.  Nothing here is on the product path.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

MAX_PAYLOAD = 0xFF00
BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
CIGAR_OPS = "MIDNSHP=X"
PSEUDO_BIN = 37450


# --------------------------------------------------------------------------- BGZF
def bgzf_block(payload: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY,
               sync_every: int = 0) -> bytes:
    """One BGZF member.  `sync_every` > 0 inserts Z_SYNC_FLUSH points so the member
    holds several deflate blocks including empty stored ones (BTYPE 0)."""
    assert len(payload) <= 65536
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    if sync_every:
        parts = []
        for i in range(0, len(payload), sync_every):
            parts.append(co.compress(payload[i:i + sync_every]))
            parts.append(co.flush(zlib.Z_SYNC_FLUSH))
        parts.append(co.flush())
        c = b"".join(parts)
    else:
        c = co.compress(payload) + co.flush()
    if len(c) + 26 > 65536:  # incompressible: store (what htslib does as well)
        co = zlib.compressobj(0, zlib.DEFLATED, -15)
        c = co.compress(payload) + co.flush()
    bsize = len(c) + 25
    assert bsize < 65536
    hdr = struct.pack("<4BI2BH2BHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 66, 67, 2, bsize)
    return hdr + c + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


class BgzfWriter:
    def __init__(self, fh, level: int = 6, split_records: bool = False,
                 payload_limit: int = MAX_PAYLOAD, strategy: int = zlib.Z_DEFAULT_STRATEGY,
                 sync_every: int = 0):
        self.fh = fh
        self.level = level
        self.split = split_records
        self.limit = payload_limit
        self.strategy = strategy
        self.sync_every = sync_every
        self.buf = bytearray()
        self.coffset = 0

    def tell(self) -> int:
        return (self.coffset << 16) | len(self.buf)

    def flush_block(self):
        if not self.buf:
            return
        blk = bgzf_block(bytes(self.buf), self.level, self.strategy, self.sync_every)
        self.fh.write(blk)
        self.coffset += len(blk)
        self.buf.clear()

    def write(self, data: bytes, atomic: bool = True):
        """atomic=True: keep `data` inside one block when it fits (htslib's bgzf_flush_try)."""
        if atomic and not self.split and self.buf and len(self.buf) + len(data) > self.limit:
            self.flush_block()
        mv = memoryview(data)
        while len(mv):
            room = self.limit - len(self.buf)
            take = min(room, len(mv))
            self.buf += mv[:take]
            mv = mv[take:]
            if len(self.buf) >= self.limit:
                self.flush_block()

    def close(self, eof: bool = True):
        self.flush_block()
        if eof:
            self.fh.write(BGZF_EOF)
            self.coffset += len(BGZF_EOF)


def iter_bgzf_blocks(data: bytes) -> Iterator[Tuple[int, int, bytes]]:
    """yield (coffset, csize, payload) for every member."""
    off = 0
    n = len(data)
    while off < n:
        if data[off:off + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF member at %d" % off)
        xlen = struct.unpack_from("<H", data, off + 10)[0]
        bsize = None
        p = off + 12
        while p < off + 12 + xlen:
            si1, si2, slen = struct.unpack_from("<BBH", data, p)
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", data, p + 4)[0]
            p += 4 + slen
        csize = bsize + 1
        raw = data[off + 12 + xlen: off + csize - 8]
        crc, isize = struct.unpack_from("<II", data, off + csize - 8)
        payload = zlib.decompress(raw, -15)
        assert len(payload) == isize and (zlib.crc32(payload) & 0xFFFFFFFF) == crc
        yield off, csize, payload
        off += csize


# --------------------------------------------------------------------------- BAM
def reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def parse_cigar(s: str) -> List[Tuple[int, int]]:
    out = []
    num = ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append((CIGAR_OPS.index(ch), int(num)))
            num = ""
    return out


def cigar_ref_len(cigar: Sequence[Tuple[int, int]]) -> int:
    return sum(l for op, l in cigar if op in (0, 2, 3, 7, 8))


def aux_int(tag: bytes, value: int, typ: str = "C") -> bytes:
    fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I"}[typ]
    return tag + typ.encode() + struct.pack("<" + fmt, value)


def aux_str(tag: bytes, value: bytes) -> bytes:
    return tag + b"Z" + value + b"\0"


def aux_array(tag: bytes, subtype: str, values: Sequence[int]) -> bytes:
    fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[subtype]
    return tag + b"B" + subtype.encode() + struct.pack("<i", len(values)) + struct.pack(
        "<%d%s" % (len(values), fmt), *values)


@dataclass
class Rec:
    tid: int
    pos: int
    qname: bytes = b"r"
    flag: int = 0
    mapq: int = 255
    cigar: List[Tuple[int, int]] = field(default_factory=list)
    l_seq: int = 0
    aux: bytes = b""
    next_tid: int = -1
    next_pos: int = -1
    tlen: int = 0
    seq: Optional[bytes] = None   # packed 4-bit, (l_seq+1)//2 bytes
    qual: Optional[bytes] = None

    def end(self) -> int:
        rl = cigar_ref_len(self.cigar)
        return self.pos + (rl if rl > 0 and not (self.flag & 4) else 1)

    def encode(self) -> bytes:
        qn = self.qname + b"\0"
        assert 1 <= len(qn) <= 255
        seq = self.seq if self.seq is not None else bytes((self.l_seq + 1) // 2)
        qual = self.qual if self.qual is not None else b"\xff" * self.l_seq
        assert len(seq) == (self.l_seq + 1) // 2 and len(qual) == self.l_seq
        b = reg2bin(self.pos, self.end()) if self.pos >= 0 else 4680
        body = struct.pack("<iiBBHHHiiii", self.tid, self.pos, len(qn), self.mapq, b,
                           len(self.cigar), self.flag, self.l_seq, self.next_tid,
                           self.next_pos, self.tlen)
        body += qn
        body += b"".join(struct.pack("<I", (l << 4) | op) for op, l in self.cigar)
        body += seq + qual + self.aux
        return struct.pack("<i", len(body)) + body


def bam_header_bytes(refs: Sequence[Tuple[str, int]], text: str = "") -> bytes:
    if not text:
        text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(
            "@SQ\tSN:%s\tLN:%d\n" % (n, l) for n, l in refs)
    t = text.encode()
    out = b"BAM\1" + struct.pack("<i", len(t)) + t + struct.pack("<i", len(refs))
    for n, l in refs:
        nb = n.encode() + b"\0"
        out += struct.pack("<i", len(nb)) + nb + struct.pack("<i", l)
    return out


class BaiBuilder:
    """Accumulates (tid, beg, end, voff_beg, voff_end) in file order -> .bai bytes."""

    def __init__(self, n_ref: int):
        self.n_ref = n_ref
        self.bins = [dict() for _ in range(n_ref)]        # bin -> [[beg,end],...]
        self.lin = [dict() for _ in range(n_ref)]         # window -> voff
        self.meta = [None] * n_ref                        # [off_beg, off_end, n_mapped, n_unmapped]
        self.last_bin = [None] * n_ref
        self.n_no_coor = 0

    def add(self, tid: int, beg: int, end: int, flag: int, v0: int, v1: int):
        if tid < 0:
            self.n_no_coor += 1
            return
        if beg < 0:
            beg, end = 0, 1  # htslib files unplaced-but-tid reads at bin 4680 / pos -1; keep it simple
        b = reg2bin(beg, end)
        chunks = self.bins[tid].setdefault(b, [])
        if chunks and self.last_bin[tid] == b and chunks[-1][1] == v0:
            chunks[-1][1] = v1
        else:
            chunks.append([v0, v1])
        self.last_bin[tid] = b
        for w in range(beg >> 14, ((end - 1) >> 14) + 1):
            if w not in self.lin[tid]:
                self.lin[tid][w] = v0
        m = self.meta[tid]
        if m is None:
            m = self.meta[tid] = [v0, v1, 0, 0]
        m[1] = v1
        if flag & 4:
            m[3] += 1
        else:
            m[2] += 1

    def tobytes(self) -> bytes:
        out = [b"BAI\1", struct.pack("<i", self.n_ref)]
        for tid in range(self.n_ref):
            bins = self.bins[tid]
            nb = len(bins) + (1 if self.meta[tid] else 0)
            out.append(struct.pack("<i", nb))
            for b in sorted(bins):
                out.append(struct.pack("<Ii", b, len(bins[b])))
                for c0, c1 in bins[b]:
                    out.append(struct.pack("<QQ", c0, c1))
            if self.meta[tid]:
                m = self.meta[tid]
                out.append(struct.pack("<IiQQQQ", PSEUDO_BIN, 2, m[0], m[1], m[2], m[3]))
            lin = self.lin[tid]
            n_intv = (max(lin) + 1) if lin else 0
            arr = [0] * n_intv
            nxt = 0
            for w in range(n_intv - 1, -1, -1):  # htslib fills unset windows from the right
                if w in lin:
                    nxt = lin[w]
                arr[w] = nxt
            out.append(struct.pack("<i", n_intv))
            out.append(struct.pack("<%dQ" % n_intv, *arr))
        out.append(struct.pack("<Q", self.n_no_coor))
        return b"".join(out)


def reg2bin_csi(beg: int, end: int, min_shift: int, depth: int) -> int:
    """CSIv1 specification, reg2bin for a (min_shift, depth) binning scheme."""
    l, s, t = depth, min_shift, ((1 << (depth * 3)) - 1) // 7
    end -= 1
    while l > 0:
        if beg >> s == end >> s:
            return t + (beg >> s)
        l -= 1
        s += 3
        t -= 1 << (l * 3)
    return 0


class CsiBuilder:
    """Accumulates records in file order -> .csi bytes (BGZF compressed), the way htslib's hts_idx_finish lays it out:
    per bin a `loffset` = linear-index entry of the bin's first window (unset windows filled from the right)."""

    def __init__(self, n_ref: int, min_shift: int = 14, depth: int = 5):
        self.n_ref, self.min_shift, self.depth = n_ref, min_shift, depth
        self.bins = [dict() for _ in range(n_ref)]
        self.lin = [dict() for _ in range(n_ref)]
        self.meta = [None] * n_ref
        self.last_bin = [None] * n_ref
        self.n_no_coor = 0

    def add(self, tid: int, beg: int, end: int, flag: int, v0: int, v1: int):
        if tid < 0:
            self.n_no_coor += 1
            return
        if beg < 0:
            beg, end = 0, 1
        b = reg2bin_csi(beg, end, self.min_shift, self.depth)
        chunks = self.bins[tid].setdefault(b, [])
        if chunks and self.last_bin[tid] == b and chunks[-1][1] == v0:
            chunks[-1][1] = v1
        else:
            chunks.append([v0, v1])
        self.last_bin[tid] = b
        for w in range(beg >> self.min_shift, ((end - 1) >> self.min_shift) + 1):
            if w not in self.lin[tid]:
                self.lin[tid][w] = v0
        m = self.meta[tid]
        if m is None:
            m = self.meta[tid] = [v0, v1, 0, 0]
        m[1] = v1
        if flag & 4:
            m[3] += 1
        else:
            m[2] += 1

    def tobytes(self) -> bytes:
        d, ms = self.depth, self.min_shift
        meta_bin = ((1 << (d * 3 + 3)) - 1) // 7 + 1
        out = [b"CSI\1", struct.pack("<iii", ms, d, 0), struct.pack("<i", self.n_ref)]
        for tid in range(self.n_ref):
            lin = self.lin[tid]
            n_intv = (max(lin) + 1) if lin else 0
            arr = [0] * n_intv
            nxt = 0
            for w in range(n_intv - 1, -1, -1):
                if w in lin:
                    nxt = lin[w]
                arr[w] = nxt
            bins = self.bins[tid]
            out.append(struct.pack("<i", len(bins) + (1 if self.meta[tid] else 0)))
            for b in sorted(bins):
                l, t = d, ((1 << (d * 3)) - 1) // 7
                while l > 0 and b < t:
                    l -= 1
                    t = ((1 << (l * 3)) - 1) // 7
                w = (b - t) << (3 * (d - l))
                loff = arr[w] if w < n_intv else 0
                out.append(struct.pack("<IQi", b, loff, len(bins[b])))
                for c0, c1 in bins[b]:
                    out.append(struct.pack("<QQ", c0, c1))
            if self.meta[tid]:
                m = self.meta[tid]
                out.append(struct.pack("<IQiQQQQ", meta_bin, 0, 2, m[0], m[1], m[2], m[3]))
        out.append(struct.pack("<Q", self.n_no_coor))
        raw = b"".join(out)
        comp = b"".join(bgzf_block(raw[i:i + MAX_PAYLOAD]) for i in range(0, len(raw), MAX_PAYLOAD))
        return comp + BGZF_EOF


def write_bam(path: str, refs: Sequence[Tuple[str, int]], records: Iterable[Rec],
              level: int = 6, split_records: bool = False, payload_limit: int = MAX_PAYLOAD,
              strategy: int = zlib.Z_DEFAULT_STRATEGY, sync_every: int = 0,
              index: bool = True, header_text: str = "", csi: Optional[Tuple[int, int]] = None) -> int:
    """Write `records` (already coordinate sorted) + `path`.bai -- or, with csi=(min_shift, depth), `path`.csi.
    Returns #records."""
    bai = CsiBuilder(len(refs), *csi) if csi else BaiBuilder(len(refs))
    n = 0
    with open(path, "wb") as fh:
        w = BgzfWriter(fh, level, split_records, payload_limit, strategy, sync_every)
        w.write(bam_header_bytes(refs, header_text), atomic=False)
        if not split_records:
            w.flush_block()  # samtools starts the records in a fresh block
        for r in records:
            enc = r.encode()
            if not split_records and w.buf and len(w.buf) + len(enc) > w.limit:
                w.flush_block()
            v0 = w.tell()
            w.write(enc, atomic=False)
            v1 = w.tell()
            bai.add(r.tid, r.pos, r.end(), r.flag, v0, v1)
            n += 1
        w.close()
    if index:
        with open(path + (".csi" if csi else ".bai"), "wb") as fh:
            fh.write(bai.tobytes())
    return n


# --------------------------------------------------------------------------- reading
@dataclass
class BamHeader:
    text: str
    refs: List[Tuple[str, int]]


def read_bam(path: str) -> Tuple[BamHeader, List[Rec]]:
    """Slow, obviously-correct reader (whole file in memory) for tests and py_oracle."""
    data = open(path, "rb").read()
    u = b"".join(p for _, _, p in iter_bgzf_blocks(data))
    assert u[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", u, 4)[0]
    text = u[8:8 + l_text].decode(errors="replace")
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", u, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", u, p)[0]
        name = u[p + 4:p + 4 + ln - 1].decode()
        p += 4 + ln
        refs.append((name, struct.unpack_from("<i", u, p)[0]))
        p += 4
    recs = []
    while p < len(u):
        bs = struct.unpack_from("<i", u, p)[0]
        (tid, pos, l_qn, mapq, _bin, n_cig, flag, l_seq, ntid, npos, tlen) = struct.unpack_from(
            "<iiBBHHHiiii", u, p + 4)
        q = p + 36
        qname = u[q:q + l_qn - 1]
        q += l_qn
        cig = [(v & 15, v >> 4) for v in struct.unpack_from("<%dI" % n_cig, u, q)]
        q += 4 * n_cig
        seq = u[q:q + (l_seq + 1) // 2]
        q += (l_seq + 1) // 2
        qual = u[q:q + l_seq]
        q += l_seq
        aux = u[q:p + 4 + bs]
        recs.append(Rec(tid, pos, qname, flag, mapq, cig, l_seq, aux, ntid, npos, tlen, seq, qual))
        p += 4 + bs
    return BamHeader(text, refs), recs


def aux_get_int(aux: bytes, tag: bytes):
    """(found, value_or_None_if_not_integer) -- htslib bam_aux_get + rust-htslib Aux::Integer."""
    p = 0
    sizes = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4, b"d": 8}
    fmts = {b"c": "b", b"C": "B", b"s": "h", b"S": "H", b"i": "i", b"I": "I"}
    while p + 3 <= len(aux):
        t = aux[p:p + 2]
        ty = aux[p + 2:p + 3]
        p += 3
        if ty in (b"Z", b"H"):
            e = aux.index(b"\0", p)
            if t == tag:
                return True, None
            p = e + 1
        elif ty == b"B":
            sub = aux[p:p + 1]
            cnt = struct.unpack_from("<i", aux, p + 1)[0]
            if t == tag:
                return True, None
            p += 5 + cnt * sizes[sub]
        else:
            if t == tag:
                if ty in fmts:
                    return True, struct.unpack_from("<" + fmts[ty], aux, p)[0]
                return True, None
            p += sizes[ty]
    return False, None
