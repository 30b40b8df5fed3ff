"""Minimal BGZF / BAM / BAI writer (and SAM -> sorted BAM converter) in pure Python + zlib.

The real workflow uses samtools for this step (reference README.md:33-38); samtools/htslib are not in the image,
so tests and the end-to-end bench build their BAM inputs here, from the SAM/BAM v1 specification.
Test/bench tooling only -- the product reads BAM through biscuit_b200/host/bq_bam.c.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
CIG_OPS = "MIDNSHP=XB"


def bgzf_block(data: bytes, level: int = 1) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    assert bsize < 65536
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


class BgzfWriter:
    """Tracks virtual offsets (coffset << 16 | uoffset) so that a BAI can be written alongside."""

    def __init__(self, path: str, block: int = 0xff00):
        self.fh = open(path, "wb")
        self.buf = bytearray()
        self.coff = 0
        self.block = block

    def tell(self) -> int:
        return (self.coff << 16) | len(self.buf)

    def write(self, data: bytes):
        self.buf += data
        while len(self.buf) >= self.block:
            self._flush(self.block)

    def _flush(self, n: int):
        blk = bgzf_block(bytes(self.buf[:n]))
        self.fh.write(blk)
        self.coff += len(blk)
        del self.buf[:n]

    def flush_block(self):
        if self.buf:
            self._flush(len(self.buf))

    def close(self):
        self.flush_block()
        self.fh.write(BGZF_EOF)
        self.fh.close()


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


def encode_record(tid, pos, mapq, flag, cigar, seq_nt16, qual, mtid, mpos, tlen, name: bytes, tags: bytes) -> bytes:
    """cigar: list of (len, op); seq_nt16: sequence of 4-bit codes; qual: bytes of phred values."""
    l_seq = len(seq_nt16)
    rlen = sum(ln for ln, op in cigar if op in (0, 2, 3, 7, 8))
    end = pos + (rlen if rlen > 0 else 1)
    s = list(seq_nt16) + [0]
    packed = bytes((s[i] << 4) | s[i + 1] for i in range(0, l_seq, 2))
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(name) + 1, mapq, reg2bin(pos, end), len(cigar), flag, l_seq, mtid, mpos, tlen)
    body += name + b"\0" + b"".join(struct.pack("<I", (ln << 4) | op) for ln, op in cigar) + packed + bytes(qual) + tags
    return struct.pack("<I", len(body)) + body, end


def tag_i(tag: str, v: int) -> bytes:
    return tag.encode() + b"i" + struct.pack("<i", v)


def tag_z(tag: str, v: str) -> bytes:
    return tag.encode() + b"Z" + v.encode() + b"\0"


def tag_a(tag: str, v: str) -> bytes:
    return tag.encode() + b"A" + v.encode()


class BamWriter:
    def __init__(self, path: str, contigs, header_text: str | None = None, block: int = 0xff00):
        """contigs: list of (name, length)."""
        self.path, self.contigs = path, list(contigs)
        self.w = BgzfWriter(path, block)
        if header_text is None:
            header_text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{ln}\n" for n, ln in self.contigs)
        t = header_text.encode()
        self.w.write(b"BAM\1" + struct.pack("<i", len(t)) + t + struct.pack("<i", len(self.contigs)))
        for n, ln in self.contigs:
            self.w.write(struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln))
        self.w.flush_block()
        self.bins = [dict() for _ in self.contigs]      # bin -> list of [beg, end] chunks
        self.lin = [dict() for _ in self.contigs]       # 16 kb window -> min voffset
        self.n_rec = 0

    def add(self, tid, pos, end, rec: bytes):
        v0 = self.w.tell()
        self.w.write(rec)
        v1 = self.w.tell()
        self.n_rec += 1
        if tid < 0:
            return
        b = reg2bin(pos, end)
        ch = self.bins[tid].setdefault(b, [])
        if ch and ch[-1][1] == v0:
            ch[-1][1] = v1
        else:
            ch.append([v0, v1])
        for w in range(pos >> 14, ((end - 1) >> 14) + 1):
            if w not in self.lin[tid] or v0 < self.lin[tid][w]:
                self.lin[tid][w] = v0

    def close(self, write_bai: bool = True):
        self.w.close()
        if not write_bai:
            return
        with open(self.path + ".bai", "wb") as fh:
            fh.write(b"BAI\1" + struct.pack("<i", len(self.contigs)))
            for tid in range(len(self.contigs)):
                bins = self.bins[tid]
                fh.write(struct.pack("<i", len(bins)))
                for b, chunks in bins.items():
                    fh.write(struct.pack("<Ii", b, len(chunks)))
                    for c0, c1 in chunks:
                        fh.write(struct.pack("<QQ", c0, c1))
                lin = self.lin[tid]
                n_intv = (max(lin) + 1) if lin else 0
                fh.write(struct.pack("<i", n_intv))
                last = 0
                for w in range(n_intv):
                    last = lin.get(w, last)
                    fh.write(struct.pack("<Q", last))
            fh.write(struct.pack("<Q", 0))


def write_bam_from_soa(path: str, contigs, per_contig_reads, sid: int | None = None, block: int = 0xff00, tag_style: str = "YD"):
    """per_contig_reads: list (one per contig, None allowed) of dicts as produced by tools/synth_plp.make_reads.
    Only reads of sample `sid` are written when sid is not None.  Returns the number of records."""
    w = BamWriter(path, contigs, block=block)
    I32MIN = np.iinfo(np.int32).min
    for tid, rd in enumerate(per_contig_reads):
        if rd is None:
            continue
        for i in range(int(rd["n_reads"])):
            if sid is not None and int(rd["sid"][i]) != sid:
                continue
            nc, co = int(rd["n_cigar"][i]), int(rd["cigar_off"][i])
            cigar = [(int(c) >> 4, int(c) & 0xf) for c in rd["cigar"][co:co + nc]]
            lq = int(rd["l_qseq"][i])
            so, qo = int(rd["seq_off"][i]), int(rd["qual_off"][i])
            packed = rd["seq"][so:so + (lq + 1) // 2]
            nt16 = np.empty(len(packed) * 2, np.uint8)
            nt16[0::2] = packed >> 4
            nt16[1::2] = packed & 0xf
            tags = b""
            if int(rd["nm"][i]) != I32MIN:
                tags += tag_i("NM", int(rd["nm"][i]))
            if int(rd["as_"][i]) != I32MIN:
                tags += tag_i("AS", int(rd["as_"][i]))
            if int(rd["mate_rlen"][i]) >= 0:
                tags += tag_z("MC", f"{int(rd['mate_rlen'][i])}M")
            b = int(rd["bss_tag"][i])
            if b >= 0:
                if tag_style == "YD":
                    tags += tag_a("YD", "fr"[b])
                elif tag_style == "ZS":
                    tags += tag_z("ZS", "+-"[b] + "+")
                elif tag_style == "none":  # no strand tag at all: the reader must infer (bisc_utils.c:163-205) ...
                    if cigar[0][1] == 5:  # ... except after a leading H, where the reference's inference indexes past SEQ (undefined)
                        tags += tag_a("YD", "fr"[b])
                else:
                    tags += tag_z("XG", ("CT", "GA")[b])
            rec, end = encode_record(tid, int(rd["pos"][i]), int(rd["mapq"][i]), int(rd["flag"][i]), cigar, nt16[:lq].tolist(),
                                     bytes(rd["qual"][qo:qo + lq]), tid, int(rd["mpos"][i]), 0, f"r{i}".encode(), tags)
            w.add(tid, int(rd["pos"][i]), end, rec)
    n = w.n_rec
    w.close()
    return n


def sam_to_sorted_bam(sam_path: str, bam_path: str) -> int:
    """Coordinate-sort a SAM file (header @SQ order) and write BAM + BAI.  Small inputs only (in-memory sort)."""
    contigs, hdr, recs = [], [], []
    with open(sam_path) as fh:
        for line in fh:
            if line.startswith("@"):
                hdr.append(line)
                if line.startswith("@SQ"):
                    f = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                    contigs.append((f["SN"], int(f["LN"])))
                continue
            recs.append(line.rstrip("\n").split("\t"))
    tid_of = {n: i for i, (n, _) in enumerate(contigs)}

    def key(f):
        t = tid_of.get(f[2], -1)
        return (t if t >= 0 else 1 << 30, int(f[3]))

    recs.sort(key=key)
    w = BamWriter(bam_path, contigs, header_text="".join(hdr))
    for f in recs:
        tid = tid_of.get(f[2], -1)
        pos = int(f[3]) - 1
        cigar = []
        if f[5] != "*":
            num = ""
            for ch in f[5]:
                if ch.isdigit():
                    num += ch
                else:
                    cigar.append((int(num), CIG_OPS.index(ch)))
                    num = ""
        seq = [] if f[9] == "*" else [NT16.get(c, 15) for c in f[9].upper()]
        qual = bytes([0xff] * len(seq)) if f[10] == "*" else bytes(ord(c) - 33 for c in f[10])
        mtid = tid if f[6] == "=" else tid_of.get(f[6], -1)
        tags = b""
        for t in f[11:]:
            tg, ty, val = t.split(":", 2)
            if ty == "i":
                tags += tag_i(tg, int(val))
            elif ty == "A":
                tags += tag_a(tg, val)
            elif ty == "f":
                tags += tg.encode() + b"f" + struct.pack("<f", float(val))
            else:
                tags += tag_z(tg, val)
        rec, end = encode_record(tid, pos, int(f[4]), int(f[1]), cigar, seq, qual, mtid, int(f[7]) - 1, int(f[8]), f[0].encode(), tags)
        w.add(tid, pos, end, rec)
    n = w.n_rec
    w.close()
    return n


def write_bam_fixed(path: str, contig: str, contig_len: int, rd: dict, level: int = 1) -> int:
    """Vectorised writer for the bench: every read has one M operation of its full length and the same tag set
    (NM:i AS:i MC:Z:<len>M YD:A:f|r), so all records have the same size and are laid out with numpy.  Records
    never straddle BGZF blocks.  Writes <path> and <path>.bai; returns the number of records."""
    n = int(rd["n_reads"])
    L = int(rd["l_qseq"][0])
    assert (rd["l_qseq"] == L).all() and (rd["n_cigar"] == 1).all()
    name_w = 10
    mc = f"{L}M".encode()
    tags_len = 7 + 7 + (3 + len(mc) + 1) + 4
    body = 32 + name_w + 4 + (L + 1) // 2 + L + tags_len
    rec = np.zeros((n, 4 + body), np.uint8)

    def put32(col, v):
        rec[:, col:col + 4] = np.ascontiguousarray(v.astype("<i4")).view(np.uint8).reshape(n, 4)

    def put16(col, v):
        rec[:, col:col + 2] = np.ascontiguousarray(v.astype("<u2")).view(np.uint8).reshape(n, 2)

    pos = rd["pos"].astype(np.int64)
    put32(0, np.full(n, body))
    put32(4, np.zeros(n))                       # refID
    put32(8, pos)
    rec[:, 12] = name_w
    rec[:, 13] = rd["mapq"]
    put16(14, np.full(n, 4680))                 # bin: not used by the reader
    put16(16, np.ones(n))                       # n_cigar_op
    put16(18, rd["flag"])
    put32(20, np.full(n, L))
    put32(24, np.zeros(n))                      # next refID
    put32(28, rd["mpos"])
    put32(32, np.zeros(n))                      # tlen
    o = 36
    ids = np.arange(n)
    digits = np.zeros((n, name_w - 1), np.uint8)
    for k in range(name_w - 2, 0, -1):
        digits[:, k] = 48 + ids % 10
        ids = ids // 10
    digits[:, 0] = ord("r")
    rec[:, o:o + name_w - 1] = digits
    o += name_w
    put32(o, np.full(n, (L << 4)))
    o += 4
    nb = (L + 1) // 2
    rec[:, o:o + nb] = rd["seq"].reshape(n, nb)
    o += nb
    rec[:, o:o + L] = rd["qual"].reshape(n, L)
    o += L
    rec[:, o:o + 3] = np.frombuffer(b"NMi", np.uint8); put32(o + 3, rd["nm"]); o += 7
    rec[:, o:o + 3] = np.frombuffer(b"ASi", np.uint8); put32(o + 3, rd["as_"]); o += 7
    t = b"MCZ" + mc + b"\0"
    rec[:, o:o + len(t)] = np.frombuffer(t, np.uint8); o += len(t)
    rec[:, o:o + 3] = np.frombuffer(b"YDA", np.uint8)
    rec[:, o + 3] = np.where(rd["bss_tag"] > 0, ord("r"), ord("f"))
    o += 4
    assert o == 4 + body
    per_block = max(1, 0xff00 // (4 + body))
    hdr_text = f"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:{contig}\tLN:{contig_len}\n".encode()
    hdr = (b"BAM\1" + struct.pack("<i", len(hdr_text)) + hdr_text + struct.pack("<i", 1) + struct.pack("<i", len(contig) + 1) +
           contig.encode() + b"\0" + struct.pack("<i", contig_len))
    flat = rec.reshape(-1)
    n_blocks = (n + per_block - 1) // per_block
    coffs = np.zeros(n_blocks + 1, np.int64)
    with open(path, "wb") as fh:
        blk = bgzf_block(hdr, level)
        fh.write(blk)
        off = len(blk)
        step = per_block * (4 + body)
        for b in range(n_blocks):
            coffs[b] = off
            blk = bgzf_block(flat[b * step:(b + 1) * step].tobytes(), level)
            fh.write(blk)
            off += len(blk)
        coffs[n_blocks] = off
        fh.write(BGZF_EOF)
    rec_idx = np.arange(n)
    voff = (coffs[rec_idx // per_block] << 16) | ((rec_idx % per_block) * (4 + body))
    end_v = int(coffs[n_blocks]) << 16
    n_intv = int((pos[-1] + L - 1) >> 14) + 1 if n else 0
    first_rec = np.searchsorted(pos + L, np.arange(n_intv, dtype=np.int64) << 14, side="right") if n else np.zeros(0, np.int64)
    with open(path + ".bai", "wb") as fh:
        fh.write(b"BAI\1" + struct.pack("<i", 1))
        if n:
            fh.write(struct.pack("<i", 1) + struct.pack("<Ii", 0, 1) + struct.pack("<QQ", int(voff[0]), end_v))
        else:
            fh.write(struct.pack("<i", 0))
        fh.write(struct.pack("<i", n_intv))
        fh.write(np.ascontiguousarray(voff[np.minimum(first_rec, n - 1)].astype("<u8")).tobytes() if n else b"")
        fh.write(struct.pack("<Q", 0))
    return n


def sam_to_soa(sam_path: str):
    """Parse a SAM file into the per-contig structure-of-arrays the pileup oracle takes (same field meanings as
    bsq_plp_reads), records in coordinate order (stable).  Returns (contigs [(name, len)], {name: dict})."""
    contigs, recs = [], {}
    with open(sam_path) as fh:
        for line in fh:
            if line.startswith("@"):
                if line.startswith("@SQ"):
                    f = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                    contigs.append((f["SN"], int(f["LN"])))
                continue
            f = line.rstrip("\n").split("\t")
            if f[2] == "*" or f[5] == "*":
                continue
            recs.setdefault(f[2], []).append(f)
    I32MIN = np.iinfo(np.int32).min
    out = {}
    for name, rows in recs.items():
        rows.sort(key=lambda f: int(f[3]))
        n = len(rows)
        d = dict(n_reads=n, pos=np.zeros(n, np.int32), mpos=np.zeros(n, np.int32), mate_rlen=np.full(n, -1, np.int32), l_qseq=np.zeros(n, np.int32),
                 nm=np.full(n, I32MIN, np.int32), as_=np.full(n, I32MIN, np.int32), flag=np.zeros(n, np.uint16), mapq=np.zeros(n, np.uint8),
                 bss_tag=np.full(n, -1, np.int8), sid=np.zeros(n, np.uint8), n_cigar=np.zeros(n, np.int32), cigar_off=np.zeros(n, np.int64),
                 seq_off=np.zeros(n, np.int64), qual_off=np.zeros(n, np.int64))
        cig, seq, qual = [], [], []
        so = qo = 0
        for i, f in enumerate(rows):
            d["pos"][i] = int(f[3]) - 1
            d["mpos"][i] = int(f[7]) - 1
            d["flag"][i] = int(f[1])
            d["mapq"][i] = int(f[4])
            ops, num = [], ""
            for ch in f[5]:
                if ch.isdigit():
                    num += ch
                else:
                    ops.append((int(num) << 4) | CIG_OPS.index(ch))
                    num = ""
            d["n_cigar"][i] = len(ops)
            d["cigar_off"][i] = len(cig)
            cig += ops
            s = [] if f[9] == "*" else [NT16.get(c, 15) for c in f[9].upper()] 
            d["l_qseq"][i] = len(s)
            s2 = s + [0]
            packed = [(s2[k] << 4) | s2[k + 1] for k in range(0, len(s), 2)]
            d["seq_off"][i] = so
            seq += packed
            so += len(packed)
            q = [0xff] * len(s) if f[10] == "*" else [ord(c) - 33 for c in f[10]]
            d["qual_off"][i] = qo
            qual += q
            qo += len(q)
            for t in f[11:]:
                tg, ty, val = t.split(":", 2)
                if tg == "NM":
                    d["nm"][i] = int(val)
                elif tg == "AS":
                    d["as_"][i] = int(val)
                elif tg == "MC":
                    rl, num = 0, ""
                    if val != "*":
                        for ch in val:
                            if ch.isdigit():
                                num += ch
                            else:
                                if ch in "MDN=X":
                                    rl += int(num)
                                num = ""
                    d["mate_rlen"][i] = rl
                elif tg == "YD" and d["bss_tag"][i] < 0:
                    d["bss_tag"][i] = 0 if val == "f" else 1 if val == "r" else -1
        d["cigar"] = np.array(cig, np.uint32)
        d["seq"] = np.array(seq, np.uint8)
        d["qual"] = np.array(qual, np.uint8)
        out[name] = d
    return contigs, out
