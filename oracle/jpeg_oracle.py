"""TEST INFRASTRUCTURE -- CPU restatement of the JPEG decode the reference's loader performs (`Image.open(...).convert('RGB')`,
/root/reference/src/utils/datasets.py:141-143), i.e. of libjpeg-turbo's default decompression path as Pillow drives it:
baseline / extended-sequential Huffman JPEG, 8 bit, one interleaved scan; integer "islow" inverse DCT (jidctint.c,
jpeg_idct_islow), "fancy" (triangle) chroma up-sampling for 2:1 horizontal and 2:1 x 2:1 factors (jdsample.c,
h2v1_fancy_upsample / h2v2_fancy_upsample), YCbCr -> RGB with the 16-bit fixed-point tables of jdcolor.c.  libjpeg-turbo is a
dependency of Pillow, not part of /root/reference; the algorithms below restate its published C code (version 3.x, the
integer code paths its SIMD kernels are bit-exact with).

Pin: bit-exact against Pillow's own decode of Pillow-encoded JPEGs over qualities, sub-samplings and odd sizes
(tests/test_jpeg.py; Pillow = the thing the reference calls).  Only tests/ may import this module; the product path is
vosb200/jpeg.py over csrc/jpeg.cu."""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
                   62, 63])      # natural (row-major) position of the k-th coefficient in zig-zag order


class Unsupported(ValueError):
    """A JPEG flavour outside the restated path (progressive, arithmetic, 12 bit, CMYK, several scans, other sampling factors)."""


class Component:
    def __init__(self, cid, h, v, tq):
        self.cid, self.h, self.v, self.tq = cid, h, v, tq
        self.td = self.ta = 0
        self.blocks_w = self.blocks_h = 0          # blocks in the padded plane (whole MCUs)
        self.coef = None                            # (blocks_h, blocks_w, 64) int16, natural order


class Header:
    def __init__(self):
        self.width = self.height = 0
        self.comps: List[Component] = []
        self.qt: Dict[int, np.ndarray] = {}        # natural order, int32
        self.huff: Dict[Tuple[int, int], Tuple[np.ndarray, np.ndarray]] = {}    # (class, id) -> (counts[16], symbols)
        self.restart = 0
        self.scan_start = 0                         # offset of the entropy-coded data
        self.hmax = self.vmax = 1
        self.mcus_x = self.mcus_y = 0
        self.adobe_transform = None
        self.jfif = False


def parse(data: bytes) -> Header:
    """Markers up to the start of the (single) scan.  ITU T.81 Annex B."""
    if data[:2] != b'\xff\xd8':
        raise Unsupported('not a JPEG')
    h = Header()
    i = 2
    while True:
        if data[i] != 0xFF:
            raise Unsupported('marker expected')
        while data[i] == 0xFF:
            i += 1
        m = data[i]
        i += 1
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            continue
        seg_len = (data[i] << 8) | data[i + 1]
        seg = data[i + 2:i + seg_len]
        i += seg_len
        if m == 0xDB:                                # DQT
            j = 0
            while j < len(seg):
                pq, tq = seg[j] >> 4, seg[j] & 15
                j += 1
                if pq == 0:
                    vals = np.frombuffer(seg[j:j + 64], dtype=np.uint8).astype(np.int32)
                    j += 64
                else:
                    vals = np.frombuffer(seg[j:j + 128], dtype='>u2').astype(np.int32)
                    j += 128
                q = np.zeros(64, dtype=np.int32)
                q[ZIGZAG] = vals
                h.qt[tq] = q
        elif m in (0xC0, 0xC1):                      # SOF0 / SOF1: baseline / extended sequential, Huffman
            if seg[0] != 8:
                raise Unsupported('sample precision %d' % seg[0])
            h.height, h.width = (seg[1] << 8) | seg[2], (seg[3] << 8) | seg[4]
            for c in range(seg[5]):
                cid, hv, tq = seg[6 + 3 * c: 9 + 3 * c]
                h.comps.append(Component(cid, hv >> 4, hv & 15, tq))
        elif 0xC2 <= m <= 0xCF and m not in (0xC4, 0xC8, 0xCC):
            raise Unsupported('SOF%d (progressive / lossless / arithmetic)' % (m - 0xC0))
        elif m == 0xC4:                              # DHT
            j = 0
            while j < len(seg):
                tc, th = seg[j] >> 4, seg[j] & 15
                counts = np.frombuffer(seg[j + 1:j + 17], dtype=np.uint8).astype(np.int32)
                n = int(counts.sum())
                h.huff[(tc, th)] = (counts, np.frombuffer(seg[j + 17:j + 17 + n], dtype=np.uint8).astype(np.int32))
                j += 17 + n
        elif m == 0xDD:
            h.restart = (seg[0] << 8) | seg[1]
        elif m == 0xE0 and seg[:5] == b'JFIF\0':
            h.jfif = True
        elif m == 0xEE and seg[:5] == b'Adobe':
            h.adobe_transform = seg[11]
        elif m == 0xDA:                              # SOS
            ns = seg[0]
            if ns != len(h.comps):
                raise Unsupported('scan with %d of %d components' % (ns, len(h.comps)))
            for c in range(ns):
                cs, tt = seg[1 + 2 * c], seg[2 + 2 * c]
                comp = next(k for k in h.comps if k.cid == cs)
                comp.td, comp.ta = tt >> 4, tt & 15
            if seg[1 + 2 * ns] != 0 or seg[2 + 2 * ns] != 63 or seg[3 + 2 * ns] != 0:
                raise Unsupported('spectral selection / successive approximation')
            h.scan_start = i
            break
        elif m == 0xD9:
            raise Unsupported('no scan')
    if len(h.comps) not in (1, 3):
        raise Unsupported('%d components' % len(h.comps))
    if len(h.comps) == 3 and h.adobe_transform not in (None, 1):
        raise Unsupported('Adobe colour transform %r' % h.adobe_transform)
    if len(h.comps) == 3 and not h.jfif and h.adobe_transform is None and [c.cid for c in h.comps] == [82, 71, 66]:
        raise Unsupported("components named 'R', 'G', 'B' without a JFIF / Adobe marker: libjpeg takes them as RGB data (jdapimin.c)")
    h.hmax, h.vmax = max(c.h for c in h.comps), max(c.v for c in h.comps)
    if len(h.comps) == 1:                            # a single-component scan is never interleaved: MCU = one block
        h.comps[0].h = h.comps[0].v = h.hmax = h.vmax = 1
    h.mcus_x = -(-h.width // (8 * h.hmax))
    h.mcus_y = -(-h.height // (8 * h.vmax))
    for c in h.comps:
        c.blocks_w, c.blocks_h = h.mcus_x * c.h, h.mcus_y * c.v
    return h


class _Bits:
    """Entropy-coded segment reader: byte stuffing (FF 00 -> FF), stops feeding at a marker."""

    def __init__(self, data: bytes, pos: int):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def _fill(self):
        b = 0
        if self.p < len(self.d):
            b = self.d[self.p]
            if b == 0xFF:
                if self.d[self.p + 1] == 0:
                    self.p += 2
                else:
                    b = 0                              # a marker: feed zeros (libjpeg's behaviour at the end of a segment)
            else:
                self.p += 1
        self.acc = ((self.acc << 8) | b) & 0xFFFFFFFF
        self.n += 8

    def get(self, k: int) -> int:
        while self.n < k:
            self._fill()
        self.n -= k
        return (self.acc >> self.n) & ((1 << k) - 1)

    def restart(self):
        self.acc = self.n = 0
        while not (self.d[self.p] == 0xFF and 0xD0 <= self.d[self.p + 1] <= 0xD7):
            self.p += 1
        self.p += 2


def _huff_table(counts, symbols):
    """code -> symbol by length (T.81 Annex C)."""
    table = {}
    code, k = 0, 0
    for length in range(1, 17):
        for _ in range(int(counts[length - 1])):
            table[(length, code)] = int(symbols[k])
            code += 1
            k += 1
        code <<= 1
    return table


def _decode_sym(bits: _Bits, table) -> int:
    code = 0
    for length in range(1, 17):
        code = (code << 1) | bits.get(1)
        s = table.get((length, code))
        if s is not None:
            return s
    raise ValueError('bad Huffman code')


def _extend(v: int, s: int) -> int:
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def entropy_decode(data: bytes, h: Header) -> None:
    """Fills comp.coef (quantised coefficients, natural order) for every component.  T.81 Annex F.2.2."""
    tables = {k: _huff_table(*v) for k, v in h.huff.items()}
    for c in h.comps:
        c.coef = np.zeros((c.blocks_h, c.blocks_w, 64), dtype=np.int16)
    bits = _Bits(data, h.scan_start)
    pred = [0] * len(h.comps)
    count = 0
    for my in range(h.mcus_y):
        for mx in range(h.mcus_x):
            if h.restart and count and count % h.restart == 0:
                bits.restart()
                pred = [0] * len(h.comps)
            count += 1
            for ci, c in enumerate(h.comps):
                dc_t, ac_t = tables[(0, c.td)], tables[(1, c.ta)]
                for by in range(c.v):
                    for bx in range(c.h):
                        blk = c.coef[my * c.v + by, mx * c.h + bx]
                        s = _decode_sym(bits, dc_t)
                        pred[ci] += _extend(bits.get(s), s) if s else 0
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = _decode_sym(bits, ac_t)
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r != 15:
                                    break
                                k += 16
                                continue
                            k += r
                            blk[ZIGZAG[k]] = _extend(bits.get(s), s)
                            k += 1


# ---- jidctint.c: jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2), vectorised over blocks ------------------------------------
_F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137, f1_961=16069,
          f2_053=16819, f2_562=20995, f3_072=25172)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _idct_1d(v0, v1, v2, v3, v4, v5, v6, v7, shift, dc_shift):
    z1 = (v2 + v6) * _F['f0_541']
    tmp2 = z1 + v6 * -_F['f1_847']
    tmp3 = z1 + v2 * _F['f0_765']
    tmp0 = (v0 + v4) << dc_shift
    tmp1 = (v0 - v4) << dc_shift
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    t0, t1, t2, t3 = v7, v5, v3, v1
    z1, z2, z3, z4 = t0 + t3, t1 + t2, t0 + t2, t1 + t3
    z5 = (z3 + z4) * _F['f1_175']
    t0, t1, t2, t3 = t0 * _F['f0_298'], t1 * _F['f2_053'], t2 * _F['f3_072'], t3 * _F['f1_501']
    z1, z2, z3, z4 = z1 * -_F['f0_899'], z2 * -_F['f2_562'], z3 * -_F['f1_961'] + z5, z4 * -_F['f0_390'] + z5
    t0, t1, t2, t3 = t0 + z1 + z3, t1 + z2 + z4, t2 + z2 + z3, t3 + z1 + z4
    return [_descale(tmp10 + t3, shift), _descale(tmp11 + t2, shift), _descale(tmp12 + t1, shift), _descale(tmp13 + t0, shift),
            _descale(tmp13 - t0, shift), _descale(tmp12 - t1, shift), _descale(tmp11 - t2, shift), _descale(tmp10 - t3, shift)]


def range_limit(x):
    """jdmaster.c prepare_range_limit_table as the IDCT indexes it: sample = table[(x & 1023)] centred on 128."""
    x = x & 1023
    return np.where(x < 128, x + 128, np.where(x < 512, 255, np.where(x < 896, 0, x - 896))).astype(np.uint8)


def idct_blocks(coef: np.ndarray, q: np.ndarray) -> np.ndarray:
    """coef (..., 64) int16 natural order, q (64,) -> samples (..., 8, 8) uint8."""
    w = (coef.astype(np.int32) * q.astype(np.int32)).reshape(coef.shape[:-1] + (8, 8)).astype(np.int64)
    cols = _idct_1d(*[w[..., r, :] for r in range(8)], shift=13 - 2, dc_shift=13)            # pass 1: columns -> rows of the workspace
    ws = np.stack(cols, axis=-2)
    rows = _idct_1d(*[ws[..., :, c] for c in range(8)], shift=13 + 2 + 3, dc_shift=13)      # pass 2: rows
    out = np.stack(rows, axis=-1)
    return range_limit(out)


def plane_from_blocks(samples: np.ndarray) -> np.ndarray:
    bh, bw = samples.shape[:2]
    return samples.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


# ---- jdsample.c --------------------------------------------------------------------------------------------------------------
def h2v1_fancy(p: np.ndarray) -> np.ndarray:
    """(h, w) -> (h, 2 w): 3/4 nearer + 1/4 further sample, rounding alternates (+1 left, +2 right); end columns copied."""
    a = p.astype(np.int32)
    h, w = a.shape
    out = np.empty((h, 2 * w), dtype=np.int32)
    left = np.concatenate([a[:, :1], a[:, :-1]], axis=1)
    right = np.concatenate([a[:, 1:], a[:, -1:]], axis=1)
    out[:, 0::2] = (a * 3 + left + 1) >> 2
    out[:, 1::2] = (a * 3 + right + 2) >> 2
    out[:, 0] = a[:, 0]
    out[:, -1] = a[:, -1]
    return out.astype(np.uint8)


def h2v2_fancy(p: np.ndarray) -> np.ndarray:
    """(h, w) -> (2 h, 2 w): vertical 3:1 blend with the nearer neighbouring row first (edge rows replicated), then the horizontal
    triangle on the column sums with rounding 8 (left) / 7 (right)."""
    a = p.astype(np.int32)
    h, w = a.shape
    up = np.concatenate([a[:1], a[:-1]], axis=0)
    down = np.concatenate([a[1:], a[-1:]], axis=0)
    out = np.empty((2 * h, 2 * w), dtype=np.int32)
    for v, other in ((0, up), (1, down)):
        s = a * 3 + other                                          # column sums
        last = np.concatenate([s[:, :1], s[:, :-1]], axis=1)
        nxt = np.concatenate([s[:, 1:], s[:, -1:]], axis=1)
        left = (s * 3 + last + 8) >> 4
        right = (s * 3 + nxt + 7) >> 4
        left[:, 0] = (s[:, 0] * 4 + 8) >> 4
        right[:, -1] = (s[:, -1] * 4 + 7) >> 4
        out[v::2, 0::2] = left
        out[v::2, 1::2] = right
    return out.astype(np.uint8)


# ---- jdcolor.c: build_ycc_rgb_table / ycc_rgb_convert ----------------------------------------------------------------------------
def _fix(x):
    return int(x * 65536 + 0.5)


_X = np.arange(256, dtype=np.int64) - 128
CR_R = (_fix(1.40200) * _X + 32768) >> 16
CB_B = (_fix(1.77200) * _X + 32768) >> 16
CR_G = -_fix(0.71414) * _X
CB_G = -_fix(0.34414) * _X + 32768


def ycc_to_rgb(y, cb, cr) -> np.ndarray:
    y = y.astype(np.int64)
    r = y + CR_R[cr]
    g = y + ((CB_G[cb] + CR_G[cr]) >> 16)
    b = y + CB_B[cb]
    return np.clip(np.stack([r, g, b], axis=-1), 0, 255).astype(np.uint8)


def reconstruct(h: Header) -> np.ndarray:
    """Coefficients -> (H, W, 3) uint8 RGB, as Image.open(...).convert('RGB') returns it."""
    planes = []
    for c in h.comps:
        full = plane_from_blocks(idct_blocks(c.coef, h.qt[c.tq]))
        ch = -(-h.height * c.v // h.vmax)                              # downsampled_height / _width (jdmaster.c)
        cw = -(-h.width * c.h // h.hmax)
        p = full[:ch, :cw]
        fh, fv = h.hmax // c.h, h.vmax // c.v
        if (fh, fv) == (1, 1):
            up = p
        elif (fh, fv) == (2, 1):
            # jdsample.c jinit_upsampler: the triangle filters need more than two samples per row, else plain replication
            up = h2v1_fancy(p) if cw > 2 else np.repeat(p, 2, axis=1)
        elif (fh, fv) == (2, 2):
            up = h2v2_fancy(p) if cw > 2 else np.repeat(np.repeat(p, 2, axis=0), 2, axis=1)
        else:
            raise Unsupported('sampling factors %dx%d of %dx%d' % (c.h, c.v, h.hmax, h.vmax))
        planes.append(up[:h.height, :h.width])
    if len(planes) == 1:
        return np.repeat(planes[0][..., None], 3, axis=-1)
    return ycc_to_rgb(*planes)


def decode(data: bytes) -> np.ndarray:
    h = parse(data)
    entropy_decode(data, h)
    return reconstruct(h)
