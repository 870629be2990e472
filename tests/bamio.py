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
