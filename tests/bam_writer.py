"""Test helper: SAM text -> BGZF-compressed BAM bytes (SAM spec section 4), so the C main's BAM input path can be
exercised without samtools (absent from this image).  Integer aux values are stored with the smallest type, as
samtools does; everything the writer emits is what htslib would print back as the original text."""
import re
import struct
import zlib

NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
CIG = {c: i for i, c in enumerate("MIDNSHP=X")}


def _bgzf(data: bytes) -> bytes:
    out = []
    for i in range(0, max(len(data), 1), 0xff00):
        chunk = data[i:i + 0xff00]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(comp) + 8
        out.append(struct.pack("<BBBBIBBHBBHH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6, ord("B"), ord("C"), 2, bsize - 1))
        out.append(comp)
        out.append(struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))     # EOF marker
    return b"".join(out)


def _aux(field: bytes) -> bytes:
    tag, ty, val = field[:2], field[3:4], field[5:]
    if ty == b"A":
        return tag + b"A" + val[:1]
    if ty == b"i":
        v = int(val)
        for code, fmt, lo, hi in ((b"C", "<B", 0, 255), (b"c", "<b", -128, 127), (b"S", "<H", 0, 65535), (b"s", "<h", -32768, 32767),
                                  (b"I", "<I", 0, 2**32 - 1), (b"i", "<i", -2**31, 2**31 - 1)):
            if lo <= v <= hi:
                return tag + code + struct.pack(fmt, v)
    if ty == b"Z" or ty == b"H":
        return tag + ty + val + b"\0"
    if ty == b"f":
        return tag + b"f" + struct.pack("<f", float(val))
    if ty == b"B":
        st = val[:1]
        nums = [x for x in val[2:].split(b",") if x]
        fmt = {b"c": "b", b"C": "B", b"s": "h", b"S": "H", b"i": "i", b"I": "I", b"f": "f"}[st]
        conv = float if st == b"f" else int
        return tag + b"B" + st + struct.pack("<I", len(nums)) + struct.pack("<%d%s" % (len(nums), fmt), *[conv(x) for x in nums])
    raise ValueError(field)


def sam_to_bam(sam: bytes) -> bytes:
    lines = sam.split(b"\n")
    hdr = b"".join(l + b"\n" for l in lines if l.startswith(b"@"))
    names, lens = [], []
    for l in lines:
        if l.startswith(b"@SQ"):
            f = dict(x.split(b":", 1) for x in l.split(b"\t")[1:])
            names.append(f[b"SN"]); lens.append(int(f[b"LN"]))
    tid = {n: i for i, n in enumerate(names)}
    out = [b"BAM\1", struct.pack("<i", len(hdr)), hdr, struct.pack("<i", len(names))]
    for n, ln in zip(names, lens):
        out.append(struct.pack("<i", len(n) + 1) + n + b"\0" + struct.pack("<i", ln))
    for l in lines:
        if not l or l.startswith(b"@"):
            continue
        f = l.split(b"\t")
        qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
        ref = tid.get(rname, -1) if rname != b"*" else -1
        nref = ref if rnext == b"=" else (tid.get(rnext, -1) if rnext != b"*" else -1)
        ops = [] if cigar == b"*" else [(int(n), CIG[chr(o[0])]) for n, o in re.findall(rb"(\d+)([MIDNSHP=X])", cigar)]
        l_seq = 0 if seq == b"*" else len(seq)
        sq = bytearray((l_seq + 1) // 2)
        for i in range(l_seq):
            sq[i >> 1] |= NT16.get(chr(seq[i]).upper(), 15) << (4 if i % 2 == 0 else 0)
        ql = (b"\xff" * l_seq) if qual == b"*" else bytes(c - 33 for c in qual)
        body = struct.pack("<iiBBHHHIiii", ref, int(pos) - 1, len(qname) + 1, int(mapq), 4680, len(ops), int(flag), l_seq, nref, int(pnext) - 1, int(tlen))
        body += qname + b"\0" + b"".join(struct.pack("<I", n << 4 | o) for n, o in ops) + bytes(sq) + ql + b"".join(_aux(x) for x in f[11:])
        out.append(struct.pack("<i", len(body)) + body)
    return _bgzf(b"".join(out))
