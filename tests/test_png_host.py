"""CPU: host half of the device PNG encoder (mangatranslator_b200/png_device.py) — length-limited Huffman code, dynamic
block header, Adler-32 combination, PNG container — checked by emulating the kernels of csrc/png_kernels.cu in Python
(same tokenisation, same bit packing, same block framing) and decoding the result with zlib and Pillow."""
import io
import zlib

import numpy as np
import pytest
from PIL import Image

from mangatranslator_b200 import png_device as PD


def _filter_rows(img, oc, ftypes):
    """Filtered stream with a given filter type per row (any choice is a valid PNG)."""
    h, w, ic = img.shape
    a = np.full((h, w, oc), 255, np.uint8)
    a[:, :, :ic] = img
    rows = a.reshape(h, w * oc).astype(np.int32)
    out = bytearray()
    for y in range(h):
        f = ftypes[y % len(ftypes)]
        cur = rows[y]
        left = np.concatenate([np.zeros(oc, np.int32), cur[:-oc]])
        up = rows[y - 1] if y else np.zeros_like(cur)
        ul = np.concatenate([np.zeros(oc, np.int32), up[:-oc]])
        if f == 0:
            pred = 0
        elif f == 1:
            pred = left
        elif f == 2:
            pred = up
        elif f == 3:
            pred = (left + up) >> 1
        else:
            p = left + up - ul
            pa, pb, pc = abs(p - left), abs(p - up), abs(p - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, up, ul))
        out.append(f)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
    return bytes(out)


def _length_code(n):
    if n <= 10:
        return 254 + n, 0, 0
    if n == 258:
        return 285, 0, 0
    l = n - 3
    e = l.bit_length() - 3
    return 261 + 4 * e + ((l >> e) & 3), e, l & ((1 << e) - 1)


def _tokens(span):
    yield ("lit", span[0])
    i = 1
    while i < len(span):
        prev, r = span[i - 1], 0
        while i + r < len(span) and span[i + r] == prev:
            r += 1
        if r >= 3:
            yield ("match", r)
            i += r
        else:
            yield ("lit", span[i])
            i += 1


def _encode_like_the_kernels(stream: bytes) -> bytes:
    total = len(stream)
    segs = (total + PD.SEG - 1) // PD.SEG
    hist = np.zeros(288, np.int64)
    parts = np.zeros((segs, 2), np.int64)
    for s in range(segs):
        seg = stream[s * PD.SEG:(s + 1) * PD.SEG]
        for t in range(0, len(seg), 64):
            for kind, v in _tokens(seg[t:t + 64]):
                hist[v if kind == "lit" else _length_code(v)[0]] += 1
        hist[256] += 1
        arr = np.frombuffer(seg, np.uint8).astype(np.int64)
        parts[s] = (arr.sum(), ((len(seg) - np.arange(len(seg))) * arr).sum())
    codes, lens, header, hbits = PD.build_table(hist)
    body = bytearray()
    for s in range(segs):
        seg = stream[s * PD.SEG:(s + 1) * PD.SEG]
        b = PD._Bits()
        b.acc, b.n = int.from_bytes(header.tobytes(), "little") & ((1 << hbits) - 1), hbits
        last = s == segs - 1
        if last:
            b.acc |= 1
        for t in range(0, len(seg), 64):
            for kind, v in _tokens(seg[t:t + 64]):
                if kind == "lit":
                    b.put(int(codes[v]), int(lens[v]))
                else:
                    sym, eb, ex = _length_code(v)
                    b.put(int(codes[sym]), int(lens[sym]))
                    b.put(ex, eb + 1)
        b.put(int(codes[256]), int(lens[256]))
        if last:
            nbytes = (b.n + 7) // 8
            body += b.acc.to_bytes(nbytes, "little")
        else:
            b0 = (b.n + 3 + 7) // 8
            body += b.acc.to_bytes(b0, "little") + b"\x00\x00\xff\xff"
    adler = PD.adler32_from_parts(parts, total)
    assert adler == zlib.adler32(stream)
    return b"\x78\x01" + bytes(body) + adler.to_bytes(4, "big")


CASES = {
    "noise": lambda r: r.integers(0, 256, (37, 53, 3), dtype=np.uint8),
    "flat_white": lambda r: np.full((40, 300, 3), 255, np.uint8),
    "gradient": lambda r: np.stack([np.add.outer(np.arange(90), np.arange(200)) % 256] * 3, -1).astype(np.uint8),
    "one_pixel": lambda r: np.array([[[7, 8, 9]]], np.uint8),
    "page_like": lambda r: np.where(r.random((150, 260, 1)) < 0.9, 255, r.integers(0, 60, (150, 260, 3))).astype(np.uint8),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("oc", [3, 4])
def test_emulated_device_stream_is_a_valid_png(name, oc):
    img = CASES[name](np.random.default_rng(3))
    stream = _filter_rows(img, oc, ftypes=(1, 2, 4, 0, 3))
    z = _encode_like_the_kernels(stream)
    assert zlib.decompress(z) == stream                              # a conforming zlib stream of dynamic blocks
    h, w = img.shape[:2]
    import struct
    png = (b"\x89PNG\r\n\x1a\n" + PD._chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6 if oc == 4 else 2, 0, 0, 0)) +
           PD._chunk(b"IDAT", z) + PD._chunk(b"IEND", b""))
    back = Image.open(io.BytesIO(png))
    back.load()
    assert back.mode == ("RGBA" if oc == 4 else "RGB") and back.size == (w, h)
    got = np.asarray(back)
    assert np.array_equal(got[:, :, :3], img)
    if oc == 4:
        assert (got[:, :, 3] == 255).all()


def test_huffman_lengths_are_limited_and_complete():
    rng = np.random.default_rng(0)
    for trial in range(30):
        n = int(rng.integers(2, 286))
        freq = np.zeros(286, np.int64)
        idx = rng.choice(286, n, replace=False)
        freq[idx] = (rng.pareto(0.6, n) * 10 + 1).astype(np.int64)          # heavy-tailed: deep trees
        if trial % 3 == 0:
            freq[idx[0]] = 10 ** 9
        lens = PD.huffman_lengths(freq, 15)
        assert max(lens) <= 15 and all((l > 0) == (f > 0) for l, f in zip(lens, freq))
        assert abs(sum(2.0 ** -l for l in lens if l) - 1.0) < 1e-12        # Kraft equality: a complete prefix code
        codes = PD.canonical_codes(lens)
        seen = set()
        for s, l in enumerate(lens):
            if l:
                key = format(codes[s], f"0{l}b")[::-1]
                assert not any(key.startswith(k) or k.startswith(key) for k in seen)
                seen.add(key)
