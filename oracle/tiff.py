"""Grayscale TIFF images for `parser_context.<name>.type = tiff` (test infrastructure, CPU restatement).

The reference reads images through libtiff (src/dune/copasi/common/tiff_file.cc:16-47) and evaluates them with
TIFFGrayscale::operator() (src/dune/copasi/common/tiff_grayscale.cc:35-105).  libtiff is not in this image: this is an
independent reader (struct-based, written apart from the product's csrc/tiff.cpp) for the baseline layouts -- one sample
per pixel, 8 / 16 / 32 / 64 bits, strips, uncompressed / PackBits / LZW / Deflate, optional horizontal predictor -- plus a
writer for the tests."""
from __future__ import annotations

import struct

import numpy as np

_TSIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8}


class TiffError(Exception):
    pass


def lzw_decode(data: bytes) -> bytes:
    """TIFF LZW (compression 5): MSB-first 9..12 bit codes, Clear = 256, EndOfInformation = 257, early change."""
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    width, prev, acc, bits, pos = 9, None, 0, 0, 0
    while True:
        while bits < width and pos < len(data):
            acc = (acc << 8) | data[pos]
            pos += 1
            bits += 8
        if bits < width:
            break
        code = (acc >> (bits - width)) & ((1 << width) - 1)
        bits -= width
        acc &= (1 << bits) - 1
        if code == 257:
            break
        if code == 256:
            table = table[:258]
            width, prev = 9, None
            continue
        if code < len(table):
            entry = table[code]
        elif prev is not None and code == len(table):
            entry = table[prev] + table[prev][:1]
        else:
            raise TiffError("corrupt LZW stream")
        out += entry
        if prev is not None:
            table.append(table[prev] + entry[:1])
        prev = code
        if len(table) + 1 >= (1 << width) and width < 12:
            width += 1
    return bytes(out)


def lzw_encode(data: bytes) -> bytes:
    """the matching encoder (tests only)"""
    out, acc, bits = bytearray(), 0, 0

    def put(code, width):
        nonlocal acc, bits
        acc = (acc << width) | code
        bits += width
        while bits >= 8:
            out.append((acc >> (bits - 8)) & 0xFF)
            bits -= 8
            acc &= (1 << bits) - 1

    table = {bytes([i]): i for i in range(256)}
    nxt, width = 258, 9
    put(256, width)
    w = b""
    for byte in data:
        wc = w + bytes([byte])
        if wc in table:
            w = wc
            continue
        put(table[w], width)
        table[wc] = nxt
        nxt += 1
        if nxt + 1 >= (1 << width) + 1 and width < 12:      # the decoder has one entry fewer at this point
            width += 1
        if nxt >= 4094:
            put(256, width)
            table = {bytes([i]): i for i in range(256)}
            nxt, width = 258, 9
        w = bytes([byte])
    if w:
        put(table[w], width)
        nxt += 1
        if nxt + 1 >= (1 << width) + 1 and width < 12:
            width += 1
    put(257, width)
    if bits:
        out.append((acc << (8 - bits)) & 0xFF)
    return bytes(out)


class Image:
    """info + scaled pixel values [rows, cols]: (zero ? raw : 2^bits - raw) / 2^bits (tiff_grayscale.cc:50-56)"""

    def __init__(self, values, x_res, y_res, x_off, y_off):
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.rows, self.cols = self.values.shape
        self.x_res, self.y_res = np.float32(x_res), np.float32(y_res)
        self.x_off, self.y_off = np.float32(x_off), np.float32(y_off)

    @staticmethod
    def _pixel(v):
        v = float(v)
        return 0 if v <= 0.0 else (0xFFFFFFFF if v >= 4294967040.0 else int(v))

    def __call__(self, x, y):
        """tiff_grayscale.cc:91-105: float arithmetic, truncation to uint32 (wrapping), clamped to the image"""
        px = self._pixel(self.x_res * (np.float32(x) - self.x_off))
        line = (self.rows - self._pixel(self.y_res * (np.float32(y) - self.y_off)) - 1) & 0xFFFFFFFF
        px = min(px, self.cols - 1)
        line = min(line, self.rows - 1)
        return float(self.values[line, px])

    def record(self):
        """what the C VM reads at an OP_TAB2 offset"""
        return [2.0, float(self.rows), float(self.cols), float(self.x_res), float(self.y_res), float(self.x_off),
                float(self.y_off)] + self.values.ravel().tolist()


def read(path: str) -> Image:
    try:
        data = open(path, "rb").read()
    except OSError:
        raise TiffError(f"File '{path}' does not exists.")
    if len(data) < 8 or data[:2] not in (b"II", b"MM"):
        raise TiffError(f"Error opening TIFF file '{path}'")
    bo = "<" if data[:2] == b"II" else ">"
    if struct.unpack(bo + "H", data[2:4])[0] != 42:
        raise TiffError(f"Error opening TIFF file '{path}'")
    ifd = struct.unpack(bo + "I", data[4:8])[0]
    nent = struct.unpack(bo + "H", data[ifd:ifd + 2])[0]
    tags = {}
    for e in range(nent):
        p = ifd + 2 + 12 * e
        tag, typ, count = struct.unpack(bo + "HHI", data[p:p + 8])
        if typ not in _TSIZE:
            continue
        size = _TSIZE[typ] * count
        at = p + 8 if size <= 4 else struct.unpack(bo + "I", data[p + 8:p + 12])[0]
        raw = data[at:at + size]
        if typ == 5:
            v = struct.unpack(bo + "II" * count, raw)
            vals = [np.float32(v[2 * k] / v[2 * k + 1]) if v[2 * k + 1] else np.float32(0) for k in range(count)]
        elif typ == 11:
            vals = list(struct.unpack(bo + "f" * count, raw))
        else:
            fmt = {1: "B", 3: "H", 4: "I", 6: "b", 8: "h", 9: "i"}.get(typ)
            vals = list(struct.unpack(bo + fmt * count, raw)) if fmt else []
        tags[tag] = vals
    photometric = tags.get(262, [99])[0]
    if photometric not in (0, 1):
        raise TiffError(f"TIFF file '{path}' must be in grayscale")
    bits = tags.get(258, [1])[0]
    rows, cols = tags[257][0], tags[256][0]
    x_res, y_res = tags.get(282, [0.0])[0], tags.get(283, [0.0])[0]
    if not x_res > 0 or not y_res > 0:
        raise TiffError(f"TIFF file '{path}' has negative resolution")
    if bits not in (8, 16, 32, 64):
        raise TiffError(f"Encoding with {bits} bits not implemented")
    if tags.get(277, [1])[0] != 1:
        raise TiffError("only one sample per pixel")
    comp = tags.get(259, [1])[0]
    if comp not in (1, 32773, 5, 8, 32946):
        raise TiffError(f"compression {comp} needs libtiff")
    predictor = tags.get(317, [1])[0]
    if predictor not in (1, 2):
        raise TiffError(f"predictor {predictor} needs libtiff")
    raw = bytearray()
    for off, n in zip(tags[273], tags[279]):
        chunk = data[off:off + n]
        if comp == 1:
            raw += chunk
        elif comp == 5:
            raw += lzw_decode(chunk)
        elif comp in (8, 32946):
            import zlib
            raw += zlib.decompress(chunk)
        else:                                   # PackBits
            i = 0
            while i < len(chunk):
                c = chunk[i] - 256 if chunk[i] > 127 else chunk[i]
                i += 1
                if c >= 0:
                    raw += chunk[i:i + c + 1]
                    i += c + 1
                elif c != -128:
                    raw += bytes([chunk[i]]) * (1 - c)
                    i += 1
    dt = np.dtype({8: "u1", 16: "u2", 32: "u4", 64: "u8"}[bits]).newbyteorder(bo)
    px = np.frombuffer(bytes(raw[:rows * cols * bits // 8]), dtype=dt).reshape(rows, cols)
    if predictor == 2:                          # horizontal differencing, modulo 2^bits
        px = np.cumsum(px.astype(np.uint64), axis=1, dtype=np.uint64) & np.uint64(2 ** bits - 1 if bits < 64 else 2 ** 64 - 1)
    px = px.astype(np.float64)
    maxv = float(2 ** bits)
    vals = (px if photometric != 0 else maxv - px) / maxv
    return Image(vals, x_res, y_res, tags.get(286, [0.0])[0], tags.get(287, [0.0])[0])


def write(path: str, pixels, bits=8, x_res=(10, 1), y_res=(10, 1), x_off=None, y_off=None, photometric=1,
          packbits=False, big_endian=False, rows_per_strip=None, compression=None, predictor=1):
    """Minimal baseline writer for the tests: `pixels` [rows, cols] unsigned integers, resolutions as rationals."""
    bo = ">" if big_endian else "<"
    px = np.asarray(pixels)
    rows, cols = px.shape
    dt = np.dtype({8: "u1", 16: "u2", 32: "u4", 64: "u8"}[bits]).newbyteorder(bo)
    rps = rows_per_strip or rows
    strips = []
    compression = compression or ("packbits" if packbits else "none")
    packbits = compression == "packbits"
    code = {"none": 1, "packbits": 32773, "lzw": 5, "deflate": 8}[compression]
    for r0 in range(0, rows, rps):
        block = px[r0:r0 + rps].astype(np.uint64)
        if predictor == 2:
            block = np.concatenate([block[:, :1], (block[:, 1:] - block[:, :-1]) & np.uint64(2 ** bits - 1 if bits < 64 else 2 ** 64 - 1)], axis=1)
        raw = block.astype(dt).tobytes()
        if compression == "lzw":
            raw = lzw_encode(raw)
        elif compression == "deflate":
            import zlib
            raw = zlib.compress(raw)
        if packbits:                            # literal runs only (valid PackBits), 128 bytes at a time
            out = bytearray()
            for i in range(0, len(raw), 128):
                chunk = raw[i:i + 128]
                out += bytes([len(chunk) - 1]) + chunk
            raw = bytes(out)
        strips.append(raw)
    entries = []                                # (tag, type, count, value bytes)
    extra = bytearray()

    def add(tag, typ, values):
        fmt = {3: "H", 4: "I"}.get(typ)
        if typ == 5:
            payload = b"".join(struct.pack(bo + "II", *v) for v in values)
        else:
            payload = struct.pack(bo + fmt * len(values), *values)
        entries.append((tag, typ, len(values), payload))

    add(256, 4, [cols]); add(257, 4, [rows]); add(258, 3, [bits]); add(259, 3, [code])
    add(262, 3, [photometric]); add(273, 4, [0] * len(strips)); add(277, 3, [1]); add(278, 4, [rps])
    add(279, 4, [len(s) for s in strips]); add(282, 5, [x_res]); add(283, 5, [y_res])
    if predictor != 1:
        add(317, 3, [predictor])
    if x_off is not None:
        add(286, 5, [x_off])
    if y_off is not None:
        add(287, 5, [y_off])
    entries.sort(key=lambda e: e[0])
    ifd_off = 8
    ifd_len = 2 + 12 * len(entries) + 4
    extra_off = ifd_off + ifd_len
    blobs, body = [], bytearray()
    for tag, typ, count, payload in entries:
        if len(payload) <= 4:
            blobs.append((tag, typ, count, payload.ljust(4, b"\0"), None))
        else:
            blobs.append((tag, typ, count, None, len(body)))
            body += payload + (b"\0" if len(payload) % 2 else b"")
    data_off = extra_off + len(body)
    strip_offsets, pos = [], data_off
    for s in strips:
        strip_offsets.append(pos)
        pos += len(s)
    out = bytearray((b"MM" if big_endian else b"II") + struct.pack(bo + "HI", 42, ifd_off))
    out += struct.pack(bo + "H", len(entries))
    for tag, typ, count, inline, where in blobs:
        if tag == 273:
            payload = struct.pack(bo + "I" * len(strip_offsets), *strip_offsets)
            if len(payload) <= 4:
                inline = payload.ljust(4, b"\0")
            else:
                body[where:where + len(payload)] = payload
        out += struct.pack(bo + "HHI", tag, typ, count)
        out += inline if inline is not None else struct.pack(bo + "I", extra_off + where)
    out += struct.pack(bo + "I", 0) + body
    for s in strips:
        out += s
    open(path, "wb").write(bytes(out))
