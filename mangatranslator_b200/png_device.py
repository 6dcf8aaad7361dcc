"""PNG encoder whose heavy half runs on the device (csrc/png_kernels.cu): the finished page is filtered, tokenised and
Huffman-coded by CUDA kernels; the host only builds the (one per image) Huffman table from a 286-bin histogram and wraps
the deflate stream in the PNG container.

Replaces, for the batch path, the PIL PNG writer the reference calls per page (core/image/image_utils.py:59-170
`save_image_with_compression`, from core/pipeline.py:1996-2018).  The file is a standard PNG: 8-bit RGB or RGBA,
non-interlaced, one IDAT chunk holding a zlib stream of dynamic-Huffman blocks; any decoder returns the exact pixels."""
from __future__ import annotations

import ctypes as C
import heapq
import struct
import threading
import zlib
from typing import Dict, Tuple

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr

SEG, STRIDE = 16384, 32768
_CL_ORDER = (16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15)


def _declare(l) -> None:
    if getattr(l, "_png_declared", False):
        return
    vp, i32, ll = C.c_void_p, C.c_int, C.c_longlong
    l.mtb_png_filter.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    l.mtb_png_histogram.argtypes = [vp, ll, vp, vp, vp]
    l.mtb_png_deflate.argtypes = [vp, ll, vp, vp, vp, i32, vp, i32, vp, vp]
    l.mtb_png_compact.argtypes = [vp, i32, vp, vp, i32, vp, vp]
    for n in ("mtb_png_filter", "mtb_png_histogram", "mtb_png_deflate", "mtb_png_compact"):
        getattr(l, n).restype = i32
    l._png_declared = True


# ---- host: length-limited Huffman code + deflate dynamic-block header -------------------------------------------------
def huffman_lengths(freq, limit: int):
    """Code lengths (0 for unused symbols) of a Huffman code with no length above `limit`: plain Huffman on frequencies
    that are flattened (halved, floor 1) until the longest code fits — a few percent from optimal in the worst case, and
    the code is always complete."""
    freq = [int(f) for f in freq]
    used = [i for i, f in enumerate(freq) if f > 0]
    lengths = [0] * len(freq)
    if not used:
        return lengths
    if len(used) == 1:
        lengths[used[0]] = 1
        return lengths
    f = {i: freq[i] for i in used}
    while True:
        heap = [(w, i, (i,)) for i, w in f.items()]
        heapq.heapify(heap)
        depth = dict.fromkeys(f, 0)
        tie = len(freq)
        while len(heap) > 1:
            w1, _, s1 = heapq.heappop(heap)
            w2, _, s2 = heapq.heappop(heap)
            for s in s1 + s2:
                depth[s] += 1
            heapq.heappush(heap, (w1 + w2, tie, s1 + s2))
            tie += 1
        if max(depth.values()) <= limit:
            for i, d in depth.items():
                lengths[i] = d
            return lengths
        f = {i: max(1, w >> 1) for i, w in f.items()}


def canonical_codes(lengths):
    """RFC 1951 section 3.2.2: codes in symbol order within each length; returned BIT-REVERSED (deflate packs Huffman codes
    most-significant bit first into an LSB-first stream)."""
    max_len = max(lengths) if lengths else 0
    bl_count = [0] * (max_len + 2)
    for ln in lengths:
        if ln:
            bl_count[ln] += 1
    code, next_code = 0, [0] * (max_len + 2)
    for bits in range(1, max_len + 1):
        code = (code + bl_count[bits - 1]) << 1
        next_code[bits] = code
    out = [0] * len(lengths)
    for sym, ln in enumerate(lengths):
        if ln:
            c = next_code[ln]
            next_code[ln] += 1
            out[sym] = int(format(c, f"0{ln}b")[::-1], 2)
    return out


class _Bits:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, value: int, nbits: int) -> None:
        self.acc |= (value & ((1 << nbits) - 1)) << self.n
        self.n += nbits

    def words(self) -> np.ndarray:
        nw = (self.n + 31) // 32
        return np.frombuffer(self.acc.to_bytes(nw * 4, "little"), dtype=np.uint32).copy()


def build_table(hist: np.ndarray):
    """hist [288] -> (codes uint16 [288] bit-reversed, lengths uint8 [288], header words uint32, header bits): one dynamic
    block header (BFINAL = 0, BTYPE = 10) for every block of the image."""
    ll = huffman_lengths(hist[:286], 15)
    if sum(1 for v in ll if v) < 2:                       # a complete code needs two symbols
        for extra in (0, 1):
            if ll[extra] == 0:
                ll[extra] = 1
                break
        ll = [1 if v else 0 for v in ll]
    hlit = max(257, max(i for i, v in enumerate(ll) if v) + 1)
    dl = [1]                                              # one distance code (distance 1), one bit
    seq = ll[:hlit] + dl
    # run-length code the lengths: 18 = 11..138 zeros, 17 = 3..10 zeros, 16 = repeat previous 3..6
    toks, i = [], 0
    while i < len(seq):
        v, j = seq[i], i
        while j < len(seq) and seq[j] == v:
            j += 1
        run = j - i
        if v == 0:
            while run >= 11:
                r = min(run, 138)
                toks.append((18, r - 11, 7))
                run -= r
            if run >= 3:
                toks.append((17, run - 3, 3))
                run = 0
            toks += [(0, 0, 0)] * run
        else:
            toks.append((v, 0, 0))
            run -= 1
            while run >= 3:
                r = min(run, 6)
                toks.append((16, r - 3, 2))
                run -= r
            toks += [(v, 0, 0)] * run
        i = j
    cl_freq = [0] * 19
    for t, _, _ in toks:
        cl_freq[t] += 1
    cl_len = huffman_lengths(cl_freq, 7)
    if sum(1 for v in cl_len if v) < 2:
        cl_len[0 if cl_len[0] == 0 else 1] = 1
    cl_code = canonical_codes(cl_len)
    hclen = 19
    while hclen > 4 and cl_len[_CL_ORDER[hclen - 1]] == 0:
        hclen -= 1
    b = _Bits()
    b.put(0, 1)                                           # BFINAL (the kernel sets it in the last block)
    b.put(2, 2)                                           # BTYPE = dynamic
    b.put(hlit - 257, 5)
    b.put(len(dl) - 1, 5)
    b.put(hclen - 4, 4)
    for k in range(hclen):
        b.put(cl_len[_CL_ORDER[k]], 3)
    for t, extra, ebits in toks:
        b.put(cl_code[t], cl_len[t])
        if ebits:
            b.put(extra, ebits)
    codes = np.zeros(288, np.uint16)
    lens = np.zeros(288, np.uint8)
    codes[:286] = canonical_codes(ll)
    lens[:286] = ll
    return codes, lens, b.words(), b.n


def adler32_from_parts(parts: np.ndarray, total: int) -> int:
    """Combine per-segment (sum of bytes, sum of (n - i) * byte_i) into the Adler-32 of the whole stream."""
    a, bsum, mod = 1, 0, 65521
    for k in range(parts.shape[0]):
        n = min(SEG, total - k * SEG)
        s1, s2 = int(parts[k, 0]), int(parts[k, 1])
        bsum = (bsum + n * a + s2) % mod
        a = (a + s1) % mod
    return (bsum << 16) | a


def _chunk(kind: bytes, data) -> bytes:
    crc = zlib.crc32(data, zlib.crc32(kind))
    return struct.pack(">I", len(data)) + kind + bytes(data) + struct.pack(">I", crc)


class PngEncoderB200:
    """Device buffers per (H, W, channels); `encode(img)` returns the bytes of a PNG file."""

    def __init__(self, device: torch.device):
        self.device = device
        self.l = lib()
        _declare(self.l)
        self._bufs: Dict[Tuple[int, int, int], dict] = {}

    def _get(self, h: int, w: int, oc: int) -> dict:
        key = (h, w, oc)
        if key not in self._bufs:
            if len(self._bufs) >= 4:
                self._bufs.pop(next(iter(self._bufs)))
            total = h * (1 + w * oc)
            segs = (total + SEG - 1) // SEG
            dev = self.device
            self._bufs[key] = dict(
                total=total, segs=segs, stream=torch.empty(total + 64, dtype=torch.uint8, device=dev),
                hist=torch.zeros(288, dtype=torch.int32, device=dev), adler=torch.zeros((segs, 2), dtype=torch.int64, device=dev),
                staged=torch.empty(segs * STRIDE, dtype=torch.uint8, device=dev), sizes=torch.zeros(segs, dtype=torch.int32, device=dev),
                out=torch.empty(segs * STRIDE, dtype=torch.uint8, device=dev),
                # pinned landing buffers rotate: the writer thread that wraps page i's stream in the PNG container still reads
                # one while the device encodes page i + 1 (finalize=False); sized for a typical stream, grown on demand
                host=[torch.empty(max(total // 2, 1 << 16), dtype=torch.uint8, pin_memory=True) for _ in range(3)],
                free=[0, 1, 2], cv=threading.Condition())
        return self._bufs[key]

    @staticmethod
    def _acquire(b: dict, nbytes: int):
        with b["cv"]:
            while not b["free"]:
                b["cv"].wait()
            k = b["free"].pop()
        if b["host"][k].numel() < nbytes:
            b["host"][k] = torch.empty(int(nbytes * 1.25), dtype=torch.uint8, pin_memory=True)
        return k

    @staticmethod
    def _release(b: dict, k: int) -> None:
        with b["cv"]:
            b["free"].append(k)
            b["cv"].notify()

    def encode(self, img: torch.Tensor, out_channels: int = 0, finalize: bool = True):
        """img: device uint8 [H][W][3|4], RGB(A) order.  out_channels 4 with a 3-channel image writes an opaque RGBA file
        (the reference's target mode for PNG output, core/pipeline.py:702-712).  Returns the file's bytes, or with
        finalize=False the ingredients for `finalize_png` (so the checksums and the container can be made on another
        thread while the device moves on to the next page)."""
        assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous() and img.shape[2] in (3, 4)
        h, w, ic = int(img.shape[0]), int(img.shape[1]), int(img.shape[2])
        oc = out_channels or ic
        b = self._get(h, w, oc)
        l, st, total, segs = self.l, stream_ptr(), b["total"], b["segs"]
        b["hist"].zero_()
        check(l.mtb_png_filter(ptr(img), h, w, ic, oc, ptr(b["stream"]), st), "mtb_png_filter")
        check(l.mtb_png_histogram(ptr(b["stream"]), total, ptr(b["hist"]), ptr(b["adler"]), st), "mtb_png_histogram")
        hist = b["hist"].cpu().numpy().astype(np.int64)               # the one small round trip: 1.1 KB
        codes, lens, header, hbits = build_table(hist)
        tbl = torch.from_numpy(np.concatenate([codes.view(np.uint8), lens, header.view(np.uint8)])).to(self.device)
        p0 = tbl.data_ptr()
        check(l.mtb_png_deflate(ptr(b["stream"]), total, p0, p0 + 576, p0 + 576 + 288, hbits, ptr(b["staged"]), STRIDE,
                                ptr(b["sizes"]), st), "mtb_png_deflate")
        sizes64 = b["sizes"].to(torch.int64)
        offsets = torch.cumsum(sizes64, 0) - sizes64
        check(l.mtb_png_compact(ptr(b["staged"]), STRIDE, ptr(b["sizes"]), ptr(offsets), segs, ptr(b["out"]), st), "mtb_png_compact")
        nbytes = int(sizes64.sum().item())
        k = self._acquire(b, nbytes)
        b["host"][k][:nbytes].copy_(b["out"][:nbytes], non_blocking=True)
        parts = b["adler"].cpu().numpy()
        torch.cuda.current_stream().synchronize()
        raw = dict(deflate=b["host"][k][:nbytes].numpy(), release=lambda: self._release(b, k), adler_parts=parts, total=total,
                   w=w, h=h, oc=oc)
        return finalize_png(raw) if finalize else raw


def finalize_png(raw: dict) -> bytes:
    """Deflate stream + Adler-32 partial sums -> the bytes of the PNG file (zlib header / trailer, IHDR, IDAT, IEND)."""
    adler = adler32_from_parts(raw["adler_parts"], raw["total"])
    try:
        body = b"\x78\x01" + raw["deflate"].tobytes() + struct.pack(">I", adler)
    finally:
        if raw.get("release"):
            raw["release"]()                 # the pinned landing buffer goes back to the encoder
    ihdr = struct.pack(">IIBBBBB", raw["w"], raw["h"], 8, 6 if raw["oc"] == 4 else 2, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", body) + _chunk(b"IEND", b"")
