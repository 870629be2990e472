"""The device BGZF codec (network-aware-bwa_b200/csrc/bgzf.cuh; replaces bgzf.c:265-330 deflate_block + bgzf_write's blocking).

Deflate does not prescribe the compressed bytes, so parity here is what a BAM reader sees: every member this codec emits
must be a well-formed BGZF member (RFC 1952 header with the BC extra field and BSIZE, bgzf.c:274-291; CRC-32 and ISIZE,
bgzf.c:318-326) that zlib inflates back to exactly the input, member boundaries every 65280 input bytes.  The checker is
zlib (the library the reference itself uses), never this codec.

  * `-m "not gpu"`: the kernel body runs on a CPU block emulator (tests/host_emu/bgzf_emu.cpp: 256 coroutine threads,
    real barriers / ballots / match_any) -- edge sizes, incompressible input (stored blocks), degenerate alphabets that
    drive the Huffman length limiter, BAM-shaped records;
  * `-m gpu`: the same through the C-ABI (bwa_gpu_bgzf_deflate), equality of the device's bytes with the emulator's,
    determinism, and a multi-megabyte BAM-shaped stream.
"""
import ctypes as C
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import refload as R

ROOT = R.ROOT
EMU_SRC = os.path.join(ROOT, "tests", "host_emu", "bgzf_emu.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "host_emu", "libbgzf_emu.so")
KERNEL = os.path.join(ROOT, "network-aware-bwa_b200", "csrc", "bgzf.cuh")
BLOCK = 65280


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(EMU_LIB) or os.path.getmtime(EMU_LIB) < max(os.path.getmtime(EMU_SRC), os.path.getmtime(KERNEL)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", EMU_LIB, EMU_SRC], check=True)
    L = C.CDLL(EMU_LIB)
    L.bgzf_emu_deflate.restype = C.c_longlong
    L.bgzf_emu_deflate.argtypes = [C.c_char_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]

    L.bgzf_emu_inflate.argtypes = [C.c_char_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    L.bgzf_emu_inflate_warp.argtypes = L.bgzf_emu_inflate.argtypes

    def inflate(data: bytes, warp_form: bool = True):
        moff = R.bwa.api.bgzf_member_offsets(data)
        n = moff.size - 1
        out = np.zeros(65536 * max(1, n), dtype=np.uint8)
        ooff = np.zeros(n + 1, dtype=np.int64)
        status = np.zeros(max(1, n), dtype=np.int32)
        fn = L.bgzf_emu_inflate_warp if warp_form else L.bgzf_emu_inflate
        bad = fn(data, len(data), n, moff.ctypes.data, out.ctypes.data, ooff.ctypes.data, status.ctypes.data)
        return out[: ooff[n]].tobytes(), status[:n], bad

    def run(data: bytes, level: int = 2):
        nblk = max(1, (len(data) + BLOCK - 1) // BLOCK)
        out = np.zeros(nblk * 65536, dtype=np.uint8)
        clen = np.zeros(nblk, dtype=np.int32)
        tot = L.bgzf_emu_deflate(data, len(data), level, out.ctypes.data, clen.ctypes.data)
        return out[:tot].tobytes(), clen[: (len(data) + BLOCK - 1) // BLOCK]

    run.inflate = inflate
    return run


def check_members(data: bytes, packed: bytes, member_len) -> None:
    """Every member well-formed, inflates (zlib) to its 65280-byte slice of the input, nothing left over."""
    assert len(member_len) == (len(data) + BLOCK - 1) // BLOCK
    pos = 0
    for k, n in enumerate(member_len):
        m = packed[pos:pos + int(n)]
        pos += int(n)
        assert len(m) == n and 26 < n <= 65536
        assert m[:12] == bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0]) and m[12:16] == b"BC\x02\x00"  # bgzf.c:274-289
        assert struct.unpack("<H", m[16:18])[0] + 1 == n  # BSIZE
        d = zlib.decompressobj(-15)
        raw = d.decompress(m[18:-8])
        assert d.eof and not d.unused_data, "the deflate stream must end exactly at the trailer"
        want = data[k * BLOCK:(k + 1) * BLOCK]
        crc, isize = struct.unpack("<II", m[-8:])
        assert raw == want, f"member {k} inflates to something else"
        assert isize == len(want) and crc == zlib.crc32(want)
    assert pos == len(packed)


def bam_like(n_records: int, seed: int = 3) -> bytes:
    """Records shaped like bam2bam's output: fixed core, sequential names, packed bases, qualities, the usual tags."""
    rng = np.random.default_rng(seed)
    out = bytearray()
    for i in range(n_records):
        name = f"read{i // 2:08d}".encode() + b"\0"
        l = 100
        seq = rng.integers(0, 256, (l + 1) // 2, dtype=np.uint8).tobytes()
        qual = np.clip(rng.normal(34, 5, l), 2, 41).astype(np.uint8).tobytes()
        tags = b"XTAU" + b"NMC" + bytes([int(rng.integers(0, 3))]) + b"X0C\x01X1C\x00XMC\x01XOC\x00XGC\x00MDZ100\0"
        core = struct.pack("<iiIIiiii", 0, int(rng.integers(0, 1 << 28)), 4680 << 16 | 37 << 8 | len(name), (99 if i % 2 == 0 else 147) << 16 | 1, l,
                           0, int(rng.integers(0, 1 << 28)), 250)
        body = core + name + struct.pack("<I", l << 4) + seq + qual + tags
        out += struct.pack("<i", len(body)) + body
    return bytes(out)


def edge_cases():
    rng = np.random.default_rng(11)
    yield "1 byte", b"A"
    yield "3 bytes", b"abc"
    yield "4 bytes", b"abcd"
    yield "run", bytes(1000)
    yield "exactly one block", bam_like(400)[:BLOCK].ljust(BLOCK, b"x")
    yield "one block and a byte", (b"0123456789" * 7000)[:BLOCK + 1]
    yield "incompressible", rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    yield "two symbols", rng.integers(0, 2, 9000, dtype=np.uint8).tobytes()
    # Fibonacci-like counts: an unlimited Huffman code would be deeper than 15 bits -> the length limiter runs
    fib = [1, 1]
    while len(fib) < 22:
        fib.append(fib[-1] + fib[-2])
    skew = np.concatenate([np.full(c, 40 + s, dtype=np.uint8) for s, c in enumerate(fib)])
    rng.shuffle(skew)
    yield "deep code", skew.tobytes()


def test_emulated_kernel_round_trips(emu):
    for name, data in edge_cases():
        for level in (2, 0):
            packed, lens = emu(data, level)
            check_members(data, packed, lens)
            if level == 0:
                assert len(packed) == len(data) + 31 * len(lens), name  # stored: 26 bytes of BGZF + 5 of deflate per member


def test_emulated_kernel_on_bam_records(emu):
    data = bam_like(520)  # ~2 blocks
    packed, lens = emu(data, 2)
    check_members(data, packed, lens)
    z2 = sum(len(zlib.compress(data[i:i + BLOCK], 2)) for i in range(0, len(data), BLOCK))
    assert len(packed) < 1.05 * z2 + 26 * len(lens), "the codec should compress about as well as the reference's zlib level 2"


def test_length_limiter_is_exercised(emu):
    """A member whose literal code needed the 15-bit repair still decodes, and it really was deeper than 15."""
    fib = [1, 1]
    while len(fib) < 22:
        fib.append(fib[-1] + fib[-2])
    assert sum(fib) <= BLOCK and len(fib) - 1 > 15
    rng = np.random.default_rng(5)
    data = np.concatenate([np.full(c, 100 + s, dtype=np.uint8) for s, c in enumerate(fib)])
    rng.shuffle(data)
    packed, lens = emu(data.tobytes(), 2)
    check_members(data.tobytes(), packed, lens)


def bgzf_file(data: bytes, level, block: int = 65280, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    """What a BAM writer produces: zlib-deflated members of `block` input bytes each (the reference: bgzf.c:265-330)."""
    import bamio
    out = []
    for o in range(0, len(data), block):
        piece = data[o:o + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
        body = co.compress(piece) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25) + body
                   + struct.pack("<II", zlib.crc32(piece) & 0xFFFFFFFF, len(piece)))
    return b"".join(out) + bamio.BGZF_EOF


def inflate_cases():
    """Members as zlib writes them at every kind of setting: stored (level 0), fixed-code (Z_FIXED), dynamic at levels 1 / 6 / 9,
    run-length and Huffman-only strategies, tiny blocks, an empty member (the BGZF EOF block), incompressible and skewed data."""
    rng = np.random.default_rng(31)
    bam = bam_like(700)
    yield "bam level 1", bgzf_file(bam, 1), bam
    yield "bam level 6", bgzf_file(bam, 6), bam
    yield "bam level 9, 4 KB blocks", bgzf_file(bam[:60000], 9, 4096), bam[:60000]
    yield "stored", bgzf_file(bam[:100000], 0), bam[:100000]
    yield "fixed code", bgzf_file(bam[:100000], 6, strategy=zlib.Z_FIXED), bam[:100000]
    yield "huffman only", bgzf_file(bam[:100000], 6, strategy=zlib.Z_HUFFMAN_ONLY), bam[:100000]
    yield "rle", bgzf_file(bytes(70000) + b"ab" * 3000, 6, strategy=zlib.Z_RLE), bytes(70000) + b"ab" * 3000
    rnd = rng.integers(0, 256, 70000, dtype=np.uint8).tobytes()
    yield "incompressible", bgzf_file(rnd, 6), rnd
    yield "one byte", bgzf_file(b"A", 6), b"A"
    text = b"the quick brown fox jumps over the lazy dog. " * 3000
    yield "long matches", bgzf_file(text, 9), text


@pytest.mark.parametrize("warp_form", [True, False])
def test_emulated_inflate_matches_zlib_writers(emu, warp_form):
    """Both forms of the decoder: the warp-per-member kernel body (the product) and the plain thread-per-member statement."""
    for name, packed, want in inflate_cases():
        got, status, bad = emu.inflate(packed, warp_form)
        assert bad == 0 and not status.any(), (name, status[status != 0][:5])
        assert got == want, name


def test_emulated_inflate_rejects_damage(emu):
    bam = bam_like(300)
    good = bgzf_file(bam, 6)
    n = int.from_bytes(good[16:18], "little") + 1
    for where in (30, n // 2, n - 12):  # inside the first member's deflate stream
        broken = bytearray(good)
        broken[where] ^= 0x5a
        got, status, bad = emu.inflate(bytes(broken))
        assert bad >= 1 or got != bam, where  # a flipped byte either breaks the code or changes the bytes; never a crash
    truncated = bytearray(good)
    truncated[n - 4:n] = (12345).to_bytes(4, "little")  # ISIZE that does not match
    _, status, bad = emu.inflate(bytes(truncated))
    assert bad >= 1 and status[0] != 0


def test_emulated_codec_round_trip_through_itself(emu):
    """deflate by this codec, inflate by this codec."""
    data = bam_like(600)
    packed, lens = emu(data, 2)
    got, status, bad = emu.inflate(packed)
    assert bad == 0 and got == data


# ------------------------------------------------------------------ the device, through the C-ABI
@pytest.fixture(scope="module")
def gpu_api():
    api = R.bwa.api
    api.init()
    yield api
    api.destroy()


@pytest.mark.gpu
def test_device_round_trips_edge_cases(gpu_api, emu):
    for name, data in edge_cases():
        for level in (2, 0):
            packed, lens, _ = gpu_api.bgzf_deflate(data, level)
            check_members(data, packed, lens)
            want, _ = emu(data, level)
            assert packed == want, f"{name}: device bytes differ from the emulated kernel's"
    packed, lens, _ = gpu_api.bgzf_deflate(b"", 2)
    assert packed == b"" and len(lens) == 0


@pytest.mark.gpu
def test_device_on_a_bam_stream(gpu_api):
    data = bam_like(90000)  # ~24 MB, ~370 members: several waves of the persistent grid
    packed, lens, ms = gpu_api.bgzf_deflate(data, 2)
    check_members(data, packed, lens)
    again, lens2, _ = gpu_api.bgzf_deflate(data, 2)
    assert again == packed and (lens == lens2).all(), "the codec must be deterministic"
    z2 = sum(len(zlib.compress(data[i:i + BLOCK], 2)) for i in range(0, len(data), BLOCK))
    assert len(packed) < 1.05 * z2 + 26 * len(lens)
    # what a BAM reader does: the whole file is one multi-member gzip stream
    import gzip
    import io
    assert gzip.GzipFile(fileobj=io.BytesIO(packed)).read() == data
    assert ms > 0


@pytest.mark.gpu
def test_device_random_sizes(gpu_api):
    rng = np.random.default_rng(23)
    for _ in range(12):
        n = int(rng.integers(1, 400000))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 1:
            data = (rng.random(n) ** 6 * 255).astype(np.uint8).tobytes()
        else:
            data = bam_like(n // 260 + 1, seed=n)[:n]
        packed, lens, _ = gpu_api.bgzf_deflate(data, 2)
        check_members(data, packed, lens)


@pytest.mark.gpu
def test_device_inflate_matches_zlib_writers(gpu_api, emu):
    for name, packed, want in inflate_cases():
        got, ooff, _ = gpu_api.bgzf_inflate(packed)
        assert got == want, name
        assert ooff[-1] == len(want)


@pytest.mark.gpu
def test_device_inflate_large_and_own_members(gpu_api):
    data = bam_like(90000)  # ~24 MB: ~370 members of zlib's, then of this codec's
    got, ooff, ms = gpu_api.bgzf_inflate(bgzf_file(data, 1))
    assert got == data and ms > 0
    packed, lens, _ = gpu_api.bgzf_deflate(data, 2)
    got, _, _ = gpu_api.bgzf_inflate(packed)
    assert got == data


@pytest.mark.gpu
def test_device_inflate_fails_loudly_on_damage(gpu_api):
    good = bgzf_file(bam_like(300), 6)
    n = int.from_bytes(good[16:18], "little") + 1
    broken = bytearray(good)
    broken[n - 4:n] = (12345).to_bytes(4, "little")
    with pytest.raises(gpu_api.BwaGpuError, match="not a valid BGZF member"):
        gpu_api.bgzf_inflate(bytes(broken))
