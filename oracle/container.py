"""oracle/container.py -- numpy restatement of the "LIMGB200" container (include/limgcu.h, limg_b200/csrc/kernels_container.cuh).

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under limg_b200/). The reference defines no bitstream, it only accounts for one
(limg.cpp:1629-1636: per-area header + rangeSize * ((8 - shiftA) + (8 - shiftB) + (8 - shiftC)) payload bits), so this file is the
independent second implementation of OUR layout: built from the reference's own per-area outputs (emission-ordered area table with
un-clamped int16 decompositions and shifts, right-aligned area-contiguous factor streams -- tests/golden, tools/make_golden.py) it yields
the bytes the GPU encoder must produce, and unpacked + decoded with the oracle's reconstruction (limg_decode.h:39-236) it must give the
reference's pDecoded.
"""
import struct

import numpy as np

MAGIC = b"LIMGB200"
HEADER = struct.Struct("<8sIIIIIIQQ")
FIELDS = ("dirA_min", "dirA_max", "dirB_offset", "dirB_mag", "dirC_offset", "dirC_mag")


def code_bits(shift: int, has_alpha: bool) -> int:
    """bits per code of one factor: 8 - shift; a dropped factor (shift 8) takes none for RGB and keeps its raw byte for RGBA (Q7)"""
    return (8 if has_alpha else 0) if shift > 7 else 8 - shift


def _pack_codes(codes: np.ndarray, pw: int, ph: int, bits: int) -> bytes:
    if bits == 0:
        return b""
    segs = (pw + 7) // 8
    grid = np.zeros((ph, segs * 8), np.uint64)
    grid[:, :pw] = codes.reshape(ph, pw) & np.uint64((1 << bits) - 1)
    grid = grid.reshape(ph * segs, 8)
    acc = np.zeros(ph * segs, np.uint64)
    for i in range(8):
        acc |= grid[:, i] << np.uint64(bits * i)
    return acc.astype("<u8").view(np.uint8).reshape(-1, 8)[:, :bits].tobytes()


def _unpack_codes(raw: np.ndarray, pw: int, ph: int, bits: int) -> np.ndarray:
    if bits == 0:
        return np.zeros(pw * ph, np.uint8)
    segs = (pw + 7) // 8
    b = np.zeros((ph * segs, 8), np.uint8)
    b[:, :bits] = raw.reshape(ph * segs, bits)
    acc = b.view("<u8").reshape(-1)
    out = np.zeros((ph * segs, 8), np.uint8)
    for i in range(8):
        out[:, i] = ((acc >> np.uint64(bits * i)) & np.uint64((1 << bits) - 1)).astype(np.uint8)
    return out.reshape(ph, segs * 8)[:, :pw].reshape(-1)


def pack(width: int, height: int, has_alpha: bool, areas: np.ndarray, fa: np.ndarray, fb: np.ndarray, fc: np.ndarray) -> bytes:
    """areas: emission-ordered area table (oracle.AREA_DTYPE / limg_b200.AREA_DTYPE); fa/fb/fc: right-aligned codes, area-contiguous."""
    ch = 4 if has_alpha else 3
    table, payload = [], []
    off = 0
    for a in areas:
        pw, ph = int(a["px_w"]), int(a["px_h"])
        n = pw * ph
        table.append(struct.pack("<HHHHBBBB", int(a["ox"]), int(a["oy"]), int(a["rx"]), int(a["ry"]), *[int(s) for s in a["shift"]], int(a["stage"])))
        for name in FIELDS:
            table.append(np.asarray(a["decomp"][name][:ch], "<i2").tobytes())
        for stream, s in zip((fa, fb, fc), a["shift"]):
            payload.append(_pack_codes(np.asarray(stream[off:off + n], np.uint64), pw, ph, code_bits(int(s), has_alpha)))
        off += n
    payload = b"".join(payload)
    header = HEADER.pack(MAGIC, 1, 1 if has_alpha else 0, width, height, len(areas), 12 + 12 * ch, len(payload), 0)
    return header + b"".join(table) + payload


def unpack(data: bytes, area_dtype) -> dict:
    magic, version, flags, w, h, count, record_bytes, payload_bytes, _ = HEADER.unpack_from(data, 0)
    assert magic == MAGIC and version == 1
    has_alpha = bool(flags & 1)
    ch = 4 if has_alpha else 3
    assert record_bytes == 12 + 12 * ch
    areas = np.zeros(count, dtype=area_dtype)
    pos = HEADER.size
    for k in range(count):
        ox, oy, rx, ry, s0, s1, s2, stage = struct.unpack_from("<HHHHBBBB", data, pos)
        pos += 12
        a = areas[k]
        a["ox"], a["oy"], a["rx"], a["ry"], a["stage"] = ox, oy, rx, ry, stage
        a["shift"] = (s0, s1, s2)
        a["px_x"], a["px_y"] = ox * 8, oy * 8
        a["px_w"], a["px_h"] = min(rx * 8, w - ox * 8), min(ry * 8, h - oy * 8)  # edge fit, limg.cpp:1722-1739
        for name in FIELDS:
            a["decomp"][name][:ch] = np.frombuffer(data, "<i2", ch, pos)
            pos += 2 * ch
    assert len(data) >= pos + payload_bytes
    raw = np.frombuffer(data, np.uint8, payload_bytes, pos)
    streams = [[], [], []]
    p = 0
    for a in areas:
        pw, ph = int(a["px_w"]), int(a["px_h"])
        segs = ((pw + 7) // 8) * ph
        for f in range(3):
            bits = code_bits(int(a["shift"][f]), has_alpha)
            streams[f].append(_unpack_codes(raw[p:p + segs * bits], pw, ph, bits))
            p += segs * bits
    assert p == payload_bytes
    return {"width": w, "height": h, "has_alpha": has_alpha, "areas": areas, "fa": np.concatenate(streams[0]), "fb": np.concatenate(streams[1]), "fc": np.concatenate(streams[2])}


def decode(data: bytes) -> np.ndarray:
    """container -> pixels through the oracle's reconstruction"""
    from oracle import oracle as lo
    u = unpack(data, lo.AREA_DTYPE)
    return lo.decode_areas(u["has_alpha"], u["areas"], u["fa"], u["fb"], u["fc"], u["height"], u["width"])
