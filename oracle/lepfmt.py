"""TEST INFRASTRUCTURE ONLY -- python readers for the reference's on-disk formats.

Used by tests/ and by tests/golden/make_golden.py to take apart files written by the UNMODIFIED reference
binary (oracle/_ref/lepton):

* ``.lep`` container (reference src/lepton/jpgcoder.cc:3779-4097 write_ujpg):
  fixed 28-byte header, zlib'd header blob (HDR / P0D / HH handoffs / ...), "CMP", mux packets
  (src/io/MuxReader.hh:230-283), LE32 file-size trailer (src/lepton/vp8_encoder.cc:603-614).
* ``.ujg`` dump (reference ``-ujg`` mode, src/lepton/simple_encoder.cc:16-54): same header, then the raw
  coefficient planes in AlignedBlock order.

Nothing in the product imports this module.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Handoff:
    luma_y_start: int
    segment_size: int
    overhang_byte: int
    num_overhang_bits: int
    last_dc: tuple
    luma_y_end: int = 0


@dataclass
class Frame:
    """What the reference derives from the JPEG header (src/lepton/jpgcoder.cc setup_imginfo_jpg :4367-4540)."""
    ncmp: int
    width: int
    height: int
    sfh: list
    sfv: list
    qidx: list
    qtables: dict  # id -> 64 u16 zigzag order
    bch: list = field(default_factory=list)
    bcv: list = field(default_factory=list)
    mcuh: int = 0
    mcuv: int = 0
    progressive: bool = False


def parse_jpeg_header(hdr: bytes) -> Frame:
    """Parse DQT/SOF out of the header bytes stored in HDR (they start right after SOI)."""
    pos = 0
    qt = {}
    frame = None
    while pos + 4 <= len(hdr):
        if hdr[pos] != 0xFF:
            pos += 1
            continue
        m = hdr[pos + 1]
        if m in (0xD8, 0x01) or 0xD0 <= m <= 0xD7:
            pos += 2
            continue
        ln = struct.unpack(">H", hdr[pos + 2:pos + 4])[0]
        seg = hdr[pos + 4:pos + 2 + ln]
        if m == 0xDB:
            p = 0
            while p < len(seg):
                pq, tq = seg[p] >> 4, seg[p] & 15
                p += 1
                if pq:
                    qt[tq] = [struct.unpack(">H", seg[p + 2 * i:p + 2 * i + 2])[0] for i in range(64)]
                    p += 128
                else:
                    qt[tq] = list(seg[p:p + 64])
                    p += 64
        elif m in (0xC0, 0xC1, 0xC2):
            h, w, n = struct.unpack(">HHB", seg[1:6])
            sfh, sfv, qi = [], [], []
            for c in range(n):
                sfh.append(seg[6 + 3 * c + 1] >> 4)
                sfv.append(seg[6 + 3 * c + 1] & 15)
                qi.append(seg[6 + 3 * c + 2])
            frame = Frame(n, w, h, sfh, sfv, qi, qt, progressive=(m == 0xC2))
        pos += 2 + ln
    assert frame is not None, "no SOF in header"
    frame.qtables = qt
    sfhm, sfvm = max(frame.sfh), max(frame.sfv)
    frame.mcuv = -(-frame.height // (8 * sfvm))
    frame.mcuh = -(-frame.width // (8 * sfhm))
    frame.bch = [frame.mcuh * s for s in frame.sfh]
    frame.bcv = [frame.mcuv * s for s in frame.sfv]
    return frame


@dataclass
class LepFile:
    magic: bytes
    version: int
    flag: int
    nseg: int
    jpeg_size: int
    header_blob: bytes
    jpeg_header: bytes
    pad_bit: int
    handoffs: list
    sections: dict
    payload_off: int
    payload: bytes  # bytes after "CMP" (without the 4-byte trailer for .lep)
    trailer: int
    frame: Frame


def jpeg_header_of(jpeg: bytes) -> bytes:
    """The bytes the reference stores in HDR for a JPEG file: everything after SOI up to and including the first SOS
    segment (src/lepton/jpgcoder.cc:2270-2300 read_jpeg collects the header segments the same way)."""
    pos = 2
    while True:
        assert jpeg[pos] == 0xFF, "marker expected at %d" % pos
        m = jpeg[pos + 1]
        ln = struct.unpack(">H", jpeg[pos + 2:pos + 4])[0]
        pos += 2 + ln
        if m == 0xDA:
            return jpeg[2:pos]


def parse_container(data: bytes, jpeg_for_header: bytes = None) -> LepFile:
    """``jpeg_for_header``: for container versions whose header blob is brotli-coded (2..4; no brotli decoder in this
    image) the blob is skipped by its length field and the frame geometry / quantisation tables are taken from the
    JPEG file the container is known to decode to.  Only valid when the fixed header says one thread-segment (the
    handoff table inside the blob is then (luma_y_start 0) and nothing else on the coefficient path lives there)."""
    magic = data[:2]
    version, flag, nseg = data[2], data[3], data[4]
    jpeg_size, zlen = struct.unpack("<II", data[20:28])
    assert data[28 + zlen:31 + zlen] == b"CMP", "CMP marker missing"
    if version != 1:
        assert jpeg_for_header is not None and nseg == 1, "brotli header blob: need the JPEG and a single segment"
        jpeg_header = jpeg_header_of(jpeg_for_header)
        frame = parse_jpeg_header(jpeg_header)
        off = 31 + zlen
        return LepFile(magic, version, flag, nseg, jpeg_size, b"", jpeg_header, 0,
                       [Handoff(0, 0, 0, 0, (0, 0, 0, 0), frame.bcv[0])], {}, off, data[off:-4],
                       struct.unpack("<I", data[-4:])[0], frame)
    blob = zlib.decompress(data[28:28 + zlen])
    p = 0
    assert blob[:3] == b"HDR"
    n = struct.unpack("<I", blob[3:7])[0]
    jpeg_header = blob[7:7 + n]
    p = 7 + n
    sections = {}
    pad_bit = 0
    handoffs = []
    while p < len(blob):
        tag = blob[p:p + 3]
        if tag in (b"P0D", b"PAD"):
            pad_bit = blob[p + 3]
            p += 4
        elif blob[p:p + 1] == b"H" and tag not in (b"HDR",):
            # 'H' 'H' nseg, 16 bytes each (src/lepton/thread_handoff.cc:46-76); first 'H' is the luma marker
            assert blob[p + 1:p + 2] == b"H"
            k = blob[p + 2]
            q = p + 3
            for i in range(k):
                r = blob[q:q + 16]
                ys, ss, ob, nb = struct.unpack("<HIBB", r[:8])
                dcs = struct.unpack("<4h", r[8:16])
                handoffs.append(Handoff(ys, ss, ob, nb, dcs))
                q += 16
            p = q
        elif tag in (b"CRS",):
            k = struct.unpack("<I", blob[p + 3:p + 7])[0]
            sections["CRS"] = list(struct.unpack("<%dI" % k, blob[p + 7:p + 7 + 4 * k]))
            p += 7 + 4 * k
        elif tag in (b"FRS", b"GRB", b"PGR", b"PGE"):
            k = struct.unpack("<I", blob[p + 3:p + 7])[0]
            sections[tag.decode()] = blob[p + 7:p + 7 + k]
            p += 7 + k
        elif tag == b"EEE":
            sections["EEE"] = list(struct.unpack("<7I", blob[p + 3:p + 31]))
            p += 31
        else:
            raise ValueError("unknown header section %r at %d" % (tag, p))
    for i in range(len(handoffs) - 1):
        handoffs[i].luma_y_end = handoffs[i + 1].luma_y_start
    frame = parse_jpeg_header(jpeg_header)
    if handoffs:
        handoffs[-1].luma_y_end = frame.bcv[0]
    off = 31 + zlen
    is_lep = magic == b"\xcf\x84"
    payload = data[off:-4] if is_lep else data[off:]
    trailer = struct.unpack("<I", data[-4:])[0] if is_lep else 0
    if is_lep and not handoffs:
        # legacy files carry no handoff table: the payload opens with the segment count and the luma split rows
        # (src/lepton/vp8_decoder.cc:337-369), then the mux packets
        k = payload[0]
        assert k >= 1
        ends = list(struct.unpack("<%dH" % (k - 1), payload[1:1 + 2 * (k - 1)])) + [frame.bcv[0]]
        starts = [0] + ends[:-1]
        handoffs = [Handoff(starts[i], 0, 0, 0xFF, (0, 0, 0, 0), ends[i]) for i in range(k)]
        payload = payload[1 + 2 * (k - 1):]
        off += 1 + 2 * (k - 1)
    return LepFile(magic, version, flag, nseg, jpeg_size, blob, jpeg_header, pad_bit, handoffs, sections, off,
                   payload, trailer, frame)


def demux(payload: bytes, version: int = 1) -> list:
    """src/io/MuxReader.hh:230-283 -- returns the 16 logical streams."""
    streams = [bytearray() for _ in range(16)]
    p = 0
    n = len(payload)
    while p < n:
        if version > 1 and payload[p:p + 3] == b"\xff\xfe\xff":
            break
        h = payload[p]
        sid, flags = h & 15, (h >> 4) & 3
        if flags == 0:
            ln = payload[p + 1] + 256 * payload[p + 2] + 1
            streams[sid] += payload[p + 3:p + 3 + ln]
            p += 3 + ln
        else:
            ln = 1024 << (2 * flags)
            streams[sid] += payload[p + 1:p + 1 + ln]
            p += 1 + ln
    return [bytes(s) for s in streams]


def truncation(lep: LepFile):
    """(trunc_bcv, trunc_bc) per component -- uncompressed_components.hh:166-183 set_block_count_dpos."""
    f = lep.frame
    bc = [f.bch[c] * f.bcv[c] for c in range(f.ncmp)]
    trunc_bcv = list(f.bcv[:f.ncmp])
    trunc_bc = list(bc)
    if "EEE" in lep.sections:
        max_dpos = lep.sections["EEE"][3:7]
        for c in range(f.ncmp):
            tbc = max_dpos[c] + 1
            vs = min(tbc // f.bch[c] + (1 if tbc % f.bch[c] else 0), f.bcv[c])
            ratio = f.sfv[c]  # min_vertical_extcmp_multiple: luma rows per MCU for this component
            while vs % ratio != 0 and vs + 1 <= f.bcv[c]:
                vs += 1
            trunc_bcv[c] = vs
            trunc_bc[c] = tbc
    return trunc_bcv, trunc_bc


def parse_ujg_planes(data: bytes):
    """Coefficient planes (AlignedBlock order, int16 [bcv*bch, 64]) from a reference ``-ujg`` dump."""
    lep = parse_container(data)
    f = lep.frame
    pay = lep.payload
    batch = struct.unpack("<I", pay[:4])[0]
    target = [f.bch[c] * f.bcv[c] for c in range(f.ncmp)] + [0] * (4 - f.ncmp)
    if "EEE" in lep.sections:
        _, tbc = truncation(lep)
        for c in range(f.ncmp):
            target[c] = tbc[c]
    cur = [0, 0, 0]
    planes = [np.zeros((f.bch[c] * f.bcv[c], 64), dtype=np.int16) for c in range(f.ncmp)]
    p = 4
    while True:
        # bt_get_cmp, src/lepton/simple_decoder.cc:34-47
        cmp_ = 0
        prog = cur[0] / target[0] if target[0] else 0.0
        for ic in (1, 2):
            if target[cmp_] and cur[ic] != target[ic]:
                cp = cur[ic] / target[ic]
                if cp < prog:
                    cmp_, prog = ic, cp
        if cur[cmp_] == target[cmp_]:
            break
        nblk = min(batch, target[cmp_] - cur[cmp_])
        arr = np.frombuffer(pay, dtype="<i2", count=nblk * 64, offset=p).reshape(nblk, 64)
        planes[cmp_][cur[cmp_]:cur[cmp_] + nblk] = arr
        p += nblk * 128
        cur[cmp_] += nblk
    return lep, planes
