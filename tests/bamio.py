"""TEST INFRASTRUCTURE: minimal unaligned-BAM writer, BAM record reader and .sai writer, so the
tests can drive the UNMODIFIED reference binary (oracle/_ref/bwa bam2bam) end to end.

BAM per the SAM spec v1; the reference reads it through bamlite.c (zlib gzread, so a plain gzip
stream is accepted) and writes BGZF (a multi-member gzip, which gzip.decompress handles).
.sai = gap_opt_t (64 B) + per read: int32 n_aln + n_aln x bwt_aln1_t (bwtaln.c:242-246, 387)."""
from __future__ import annotations

import gzip
import struct

import numpy as np

NT16 = np.array([1, 2, 4, 8, 15], dtype=np.uint8)  # A C G T N -> 4-bit codes


def _record(name: bytes, seq: np.ndarray, flag: int, rg: bytes) -> bytes:
    l = int(seq.size)
    codes = NT16[seq]
    if l & 1:
        codes = np.append(codes, 0)
    packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).tobytes()
    qual = bytes([30]) * l
    tags = b"RGZ" + rg + b"\0"
    body = struct.pack("<iiIIiiii", -1, -1, (4680 << 16) | (len(name) + 1), flag << 16, l, -1, -1, 0)
    body += name + b"\0" + packed + qual + tags
    return struct.pack("<i", len(body)) + body


def write_unaligned_bam(path: str, reads, mates=None, rg: str = "rg1") -> None:
    """Single-end (flag 4) or paired (flags 77/141, mates adjacent, same name) unaligned BAM."""
    text = f"@HD\tVN:1.0\tSO:unsorted\n@RG\tID:{rg}\tSM:s\n".encode()
    out = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", 0)]
    for i in range(reads.n):
        name = f"r{i}".encode()
        if mates is None:
            out.append(_record(name, reads.read(i), 4, rg.encode()))
        else:
            out.append(_record(name, reads.read(i), 77, rg.encode()))
            out.append(_record(name, mates.read(i), 141, rg.encode()))
    with gzip.open(path, "wb", compresslevel=1) as f:
        f.write(b"".join(out))


def _bgzf_block(data: bytes, level: int) -> bytes:
    import zlib
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25) + body
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def write_unaligned_bam_fast(path: str, reads, mates=None, rg: str = "rg1", level: int = 1, threads: int = 8, qual_seed: int = 1) -> None:
    """The same file content as write_unaligned_bam for FIXED-LENGTH reads, laid out with numpy (millions of records in
    seconds) and written as BGZF (64 KB blocks, the framing `samtools`/sequencer pipelines produce; bamlite's gzread takes it
    as a multi-member gzip).  Names are r%09d; qualities are seeded noise around 30 (a constant would deflate to nothing)."""
    from concurrent.futures import ThreadPoolExecutor
    n = reads.n
    L = int(reads.offs[1] - reads.offs[0]) if n else 0
    assert n == 0 or (np.diff(reads.offs) == L).all(), "fixed-length reads only"
    per = 1 if mates is None else 2
    nrec = n * per
    name_w = 10  # 'r' + 9 digits
    packed_w = (L + 1) // 2
    tags = b"RGZ" + rg.encode() + b"\0"
    body_w = 32 + name_w + 1 + packed_w + L + len(tags)
    rec = np.zeros((nrec, 4 + body_w), dtype=np.uint8)
    u32 = np.zeros((nrec, 9), dtype="<u4")
    u32[:, 0] = body_w
    u32[:, 1] = 0xFFFFFFFF  # refID -1
    u32[:, 2] = 0xFFFFFFFF  # pos -1
    u32[:, 3] = (4680 << 16) | (name_w + 1)
    flags = np.full(nrec, 4, dtype=np.uint32)
    if per == 2:
        flags[0::2] = 77
        flags[1::2] = 141
    u32[:, 4] = flags << 16
    u32[:, 5] = L
    u32[:, 6] = 0xFFFFFFFF
    u32[:, 7] = 0xFFFFFFFF
    u32[:, 8] = 0
    rec[:, :36] = u32.view(np.uint8).reshape(nrec, 36)
    ids = np.repeat(np.arange(n, dtype=np.int64), per)
    rec[:, 36] = ord("r")
    for d in range(9):
        rec[:, 37 + d] = (ids // 10 ** (8 - d)) % 10 + 48
    o = 36 + name_w + 1
    bases = np.empty((nrec, L + (L & 1)), dtype=np.uint8)
    if per == 1:
        bases[:, :L] = reads.bases.reshape(n, L)
    else:
        bases[0::2, :L] = reads.bases.reshape(n, L)
        bases[1::2, :L] = mates.bases.reshape(n, L)
    codes = NT16[bases[:, :L]]
    if L & 1:
        codes = np.concatenate([codes, np.zeros((nrec, 1), dtype=np.uint8)], axis=1)
    rec[:, o:o + packed_w] = (codes[:, 0::2] << 4) | codes[:, 1::2]
    o += packed_w
    rng = np.random.default_rng(qual_seed)
    rec[:, o:o + L] = 30 + rng.integers(-4, 5, size=(nrec, L), dtype=np.int8).astype(np.uint8)
    o += L
    rec[:, o:o + len(tags)] = np.frombuffer(tags, dtype=np.uint8)
    text = f"@HD\tVN:1.0\tSO:unsorted\n@RG\tID:{rg}\tSM:s\n".encode()
    head = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 0)
    data = head + rec.tobytes()
    B = 65280
    pieces = [data[i:i + B] for i in range(0, len(data), B)]
    with ThreadPoolExecutor(max(1, threads)) as ex:
        blocks = list(ex.map(lambda b: _bgzf_block(b, level), pieces))
    with open(path, "wb") as f:
        for b in blocks:
            f.write(b)
        f.write(BGZF_EOF)


def read_bam_records(path: str) -> list:
    """-> list of raw record bodies (bytes, without the block_size prefix); the header is skipped
    (it embeds the command line, bam2bam.c:168-172)."""
    data = gzip.decompress(open(path, "rb").read())
    assert data[:4] == b"BAM\1"
    (l_text,) = struct.unpack_from("<i", data, 4)
    p = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, p)
    p += 4
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, p)
        p += 4 + l_name + 4
    recs = []
    while p < len(data):
        (bs,) = struct.unpack_from("<i", data, p)
        recs.append(data[p + 4:p + 4 + bs])
        p += 4 + bs
    return recs


def describe(rec: bytes) -> str:
    ref, pos, bmn, fnc, l_seq, mref, mpos, tlen = struct.unpack_from("<iiIIiiii", rec, 0)
    l_name = bmn & 0xff
    name = rec[32:32 + l_name - 1].decode()
    return f"{name} flag={fnc >> 16} ref={ref} pos={pos} mapq={(bmn >> 8) & 0xff} ncig={fnc & 0xffff} mref={mref} mpos={mpos} tlen={tlen}"


def write_sai(path: str, opt, n_aln: np.ndarray, aln_off: np.ndarray, aln: np.ndarray, index=None) -> None:
    """index: optional read indices (e.g. every mate 1) into the flat result arrays."""
    idx = range(n_aln.size) if index is None else index
    with open(path, "wb") as f:
        f.write(bytes(opt))
        for i in idx:
            f.write(struct.pack("<i", int(n_aln[i])))
            if n_aln[i]:
                f.write(aln[aln_off[i]:aln_off[i + 1]].tobytes())
