"""The "LIMGB200" container (include/limgcu.h): the numpy restatement (oracle/container.py) against the reference's golden outputs on the
CPU, the header parser of the C ABI without a device, and -- on the GPU -- the encoder's bytes against the restatement's, bit for bit."""
import ctypes as C
import struct

import numpy as np
import pytest

from oracle import container as oc
from oracle import oracle as lo
from tests import helpers as H


def golden_area_table(g, dtype):
    n = g["area_rect"].shape[0]
    a = np.zeros(n, dtype=dtype)
    for i, k in enumerate(("ox", "oy", "rx", "ry", "stage")):
        a[k] = g["area_rect"][:, i]
    for i, k in enumerate(("px_x", "px_y", "px_w", "px_h")):
        a[k] = g["area_px"][:, i]
    a["shift"] = g["area_shift"]
    d = H.golden_area_decomps(g)
    for name in d.dtype.names:
        a["decomp"][name] = d[name]
    return a


def golden_container(g) -> bytes:
    h, w = g["img"].shape
    return oc.pack(w, h, bool(g["has_alpha"]), golden_area_table(g, lo.AREA_DTYPE), g["post"][0], g["post"][1], g["post"][2])


@pytest.mark.parametrize("name", H.golden_image_cases())
def test_container_of_reference_outputs_decodes_to_reference_pixels(name):
    """reference areas + factor streams -> container bytes -> unpack -> oracle reconstruction == the reference's pDecoded"""
    g = H.load_golden(name)
    data = golden_container(g)
    alpha = bool(g["has_alpha"])
    u = oc.unpack(data, lo.AREA_DTYPE)
    assert (u["width"], u["height"], u["has_alpha"]) == (g["img"].shape[1], g["img"].shape[0], alpha)
    shifts = np.repeat(g["area_shift"], g["area_px"][:, 2].astype(np.int64) * g["area_px"][:, 3], axis=0)
    for f, key in enumerate(("fa", "fb", "fc")):
        kept = (shifts[:, f] < 8) | alpha  # an RGB factor without bits reads back as code 0 (its normal is zero, Q7)
        assert np.array_equal(u[key][kept], g["post"][f][kept])
        assert not u[key][~kept].any()
    assert np.array_equal(oc.decode(data), g["plane_pDecoded"])
    # the payload is exactly the bits the reference accounts for (limg.cpp:1632), rounded up to whole 8-pixel runs
    bits = sum(((int(pw) + 7) // 8) * 8 * int(ph) * sum(oc.code_bits(int(s), alpha) for s in sh) for (_, _, pw, ph), sh in zip(g["area_px"], g["area_shift"]))
    assert struct.unpack_from("<Q", data, 32)[0] * 8 == bits


def test_ragged_runs_and_all_bit_widths_round_trip():
    rng = np.random.default_rng(5)
    for alpha in (False, True):
        w, h = 37, 21
        areas = np.zeros(15, dtype=lo.AREA_DTYPE)
        k = 0
        for by in range(3):
            for bx in range(5):
                a = areas[k]
                a["ox"], a["oy"], a["rx"], a["ry"], a["stage"] = bx, by, 1, 1, 2
                a["px_x"], a["px_y"], a["px_w"], a["px_h"] = bx * 8, by * 8, min(8, w - bx * 8), min(8, h - by * 8)
                a["shift"] = [(k + j * 4) % 9 for j in range(3)]
                for name in oc.FIELDS:
                    a["decomp"][name][: 4 if alpha else 3] = rng.integers(-32768, 32768, 4 if alpha else 3)
                k += 1
        n = int((areas["px_w"].astype(np.int64) * areas["px_h"]).sum())
        sh = np.repeat(areas["shift"], areas["px_w"].astype(np.int64) * areas["px_h"], axis=0)
        streams = [(rng.integers(0, 256, n) >> np.where(sh[:, f] > 7, 0, sh[:, f])).astype(np.uint8) for f in range(3)]
        data = oc.pack(w, h, alpha, areas, *streams)
        u = oc.unpack(data, lo.AREA_DTYPE)
        assert u["areas"].tobytes() == areas.tobytes()
        for f, key in enumerate(("fa", "fb", "fc")):
            kept = (sh[:, f] < 8) | alpha
            assert np.array_equal(u[key][kept], streams[f][kept])


def test_header_parser_of_the_c_abi_needs_no_device():
    from limg_b200 import _lib
    lib = _lib.load()
    g = H.load_golden("rgba_photo_64x64")
    data = golden_container(g)
    buf = np.frombuffer(data, np.uint8)
    w, h, a, n, pb = C.c_size_t(0), C.c_size_t(0), C.c_int(0), C.c_uint32(0), C.c_uint64(0)
    args = (C.byref(w), C.byref(h), C.byref(a), C.byref(n), C.byref(pb))
    assert lib.limgcu_container_info(buf.ctypes.data_as(C.c_void_p), buf.size, *args) == 0
    assert (w.value, h.value, a.value, n.value) == (64, 64, 1, g["area_rect"].shape[0])
    assert 48 + n.value * 60 + pb.value == len(data)
    assert lib.limgcu_container_bound(64, 64, 1) >= len(data)
    # truncated, wrong magic, wrong record size, area count beyond the block count
    assert lib.limgcu_container_info(buf.ctypes.data_as(C.c_void_p), len(data) - 1, *args) == 103
    for off, val in ((0, b"X"), (28, struct.pack("<I", 48)), (24, struct.pack("<I", 65))):
        bad = bytearray(data)
        bad[off:off + len(val)] = val
        b = np.frombuffer(bytes(bad), np.uint8)
        assert lib.limgcu_container_info(b.ctypes.data_as(C.c_void_p), b.size, *args) == 101
    assert lib.limgcu_container_info(None, 0, *args) == 102


# ---- GPU -------------------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def codec():
    from limg_b200 import Codec
    c = Codec(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", H.golden_image_cases())
def test_gpu_container_equals_the_restatement_built_from_reference_outputs(codec, name):
    g = H.load_golden(name)
    alpha = bool(g["has_alpha"])
    want = golden_container(g)
    # decoding a reference-produced container is bit exact (also for the AES-dithered one)
    assert np.array_equal(codec.decode_container(want), g["plane_pDecoded"])
    if bool(g["aes"]):
        return  # the GPU dithers with the LCG: same areas and shifts, different noise
    got = codec.encode_container(g["img"], alpha, int(g["error_factor"]), bool(g["fast"]))
    assert got == want
    assert codec.container_info(got)["area_count"] == g["area_rect"].shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(61, 37, False), (61, 37, True), (200, 136, False)])
def test_gpu_container_ragged_sizes_vs_oracle(codec, shape):
    from limg_b200 import synth
    w, h, alpha = shape
    img = synth.photo_like(w, h, 11, 4 if alpha else 3)
    o = lo.blocked_encode3d(img, alpha, 100, True)
    data = codec.encode_container(img, alpha)
    assert np.array_equal(codec.decode_container(data), o["planes"]["pDecoded"])
    assert np.array_equal(oc.decode(data), o["planes"]["pDecoded"])
    u = oc.unpack(data, lo.AREA_DTYPE)
    for k in ("ox", "oy", "rx", "ry", "stage", "px_x", "px_y", "px_w", "px_h", "shift"):
        assert np.array_equal(u["areas"][k], o["areas"][k]), k


@pytest.mark.gpu
def test_gpu_container_rejects_a_table_that_does_not_tile(codec):
    from limg_b200 import LimgError
    g = H.load_golden("rgb_photo_96x64")
    data = bytearray(golden_container(g))
    data[48 + 4] = 0  # rx of the first record
    with pytest.raises(LimgError):
        codec.decode_container(bytes(data))


@pytest.mark.gpu
def test_gpu_container_that_claims_a_huge_image_is_rejected_before_anything_is_allocated(codec):
    """A ~100-byte file whose header claims 65528 x 65528 pixels: the table is validated on the host first, nothing is allocated from the
    header's claim, and the context decodes a sound container right afterwards."""
    from limg_b200 import LimgError
    g = H.load_golden("rgb_photo_96x64")
    good = golden_container(g)
    magic, version, flags, w, h, count, rec, payload, rsv = oc.HEADER.unpack(good[:48])
    # one record, claims to cover a block rectangle that is not the whole (huge) grid: must fail in validation
    fake = oc.HEADER.pack(magic, version, flags, 65528, 65528, 1, rec, 0, rsv) + good[48:48 + rec]
    launches = codec.launch_count()
    with pytest.raises(LimgError):
        _decode_raw(codec, fake, 65528 * 65528)
    assert codec.launch_count() == launches  # no kernel ran
    assert np.array_equal(codec.decode_container(good), g["plane_pDecoded"])


def _decode_raw(codec, data, out_pixels):
    """limgcu_host_decode_container with a caller-declared output size but a tiny real buffer: valid only for inputs that must be rejected"""
    from limg_b200 import LimgError
    out = np.zeros(16, np.uint32)
    rc = codec.lib.limgcu_host_decode_container(codec.h, data, len(data), out.ctypes.data_as(C.c_void_p), C.c_size_t(out_pixels))
    if rc != 0:
        raise LimgError("limgcu_host_decode_container failed with %d: %s" % (rc, codec.lib.limgcu_last_error(codec.h).decode()))


@pytest.mark.gpu
@pytest.mark.parametrize("what", ["outside", "overlap", "gap", "shift", "pixels", "count"])
def test_gpu_stream_decode_rejects_a_bad_area_table(codec, what):
    """limgcu_host_decode validates a caller-supplied table before a kernel indexes the block map and the planes with it."""
    from limg_b200 import LimgError
    from limg_b200 import synth
    img = synth.photo_like(96, 64, 5, 3)
    st = codec.encode_stream(img, False, 100, True)
    a = st["areas"].copy()
    k = int(np.argmax(a["rx"] * a["ry"]))
    if what == "outside":
        a["ox"][k] = 4000
    elif what == "overlap":
        a["ox"][0], a["oy"][0], a["px_x"][0], a["px_y"][0] = a["ox"][1], a["oy"][1], a["px_x"][1], a["px_y"][1]
    elif what == "gap":
        a = a[:-1]
    elif what == "shift":
        a["shift"][k, 1] = 9
    elif what == "pixels":
        a["px_w"][k] += 8
    elif what == "count":
        a = a[:0]
    with pytest.raises(LimgError):
        codec.decode(a, st["codesA"], st["codesB"], st["codesC"], False)
    assert np.array_equal(codec.decode(st["areas"], st["codesA"], st["codesB"], st["codesC"], False), codec.encode_stream(img, False, 100, True, decoded=True)["decoded"])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["c2_4k_photo", "c3_8k_rgba"])
def test_gpu_container_full_size_round_trip(codec, cfg):
    from limg_b200 import synth
    img, alpha = synth.CONFIGS[cfg]()
    st = codec.encode_stream(img, alpha, 100, True, decoded=True)
    data = codec.encode_container(img, alpha)
    info = codec.container_info(data)
    a = st["areas"]
    assert info["area_count"] == len(a)
    bits = np.where(a["shift"] > 7, 8 if alpha else 0, 8 - a["shift"].astype(np.int64)).sum(axis=1)
    assert info["payload_bytes"] == int((((a["px_w"].astype(np.int64) + 7) // 8) * a["px_h"] * bits).sum())
    assert len(data) < img.size * (4 if alpha else 3)  # smaller than the raw pixels
    assert np.array_equal(codec.decode_container(data), st["decoded"])


@pytest.mark.gpu
def test_gpu_batch_of_frames_equals_one_at_a_time(codec):
    """Batch mode: frames on several contexts at once through the C batch entry points == the frames one at a time."""
    from limg_b200 import BatchCodec, synth
    frames = [synth.photo_like(320, 200, 60 + i, 3) for i in range(7)]
    want = [codec.encode_container(f, False) for f in frames]
    b = BatchCodec(0, lanes=3)
    try:
        got = b.encode_containers(frames, False)
        assert got == want
        decoded = b.decode_containers(got)
        for d, w_ in zip(decoded, want):
            assert np.array_equal(d, codec.decode_container(w_))
        streams = b.encode_streams(frames, False, decoded=True)
        for s, d in zip(streams, decoded):
            assert np.array_equal(s["decoded"], d)
    finally:
        b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(640, 360, False, 1), (640, 360, False, 3), (328, 200, True, 2), (1920, 1080, False, 4), (200, 36, False, 8)])
def test_gpu_exact_row_bands_equal_the_whole_image_encode(codec, case):
    """the phased, sharded encode (ranks simulated by contexts on one GPU, the all-reduces done by hand) gives the same area
    table (decompositions, shifts, dither chain) and the same codes as one encode of the whole image."""
    import torch
    from limg_b200 import Codec, shard, synth
    w, h, alpha, world = case
    img = synth.photo_like(w, h, 17, 4 if alpha else 3)
    want = codec.encode_stream(img, alpha, 100, True)
    d_src = torch.from_numpy(img.view(np.int32)).cuda()
    codecs = [Codec(0) for _ in range(world)]
    try:
        ranks = [shard.RowBandExact(c, d_src, w, h, alpha, r, world) for r, c in enumerate(codecs)]
        for r in ranks:
            r.pass1()                                   # stream ordered on the rank's own codec stream
            r.codec.sync()
        table = sum(r.table for r in ranks)             # the all-gather: every chunk of the table is zero except on the rank that owns it
        for r in ranks:
            r.table.copy_(table)
        torch.cuda.synchronize()
        for r in ranks:
            r.merge_and_encode()
            r.codec.sync()
        results = sum(r.results for r in ranks)         # the SUM all-reduce
        for r in ranks:
            r.results.copy_(results)
        torch.cuda.synchronize()
        codes = [np.zeros((h, w), np.uint8) for _ in range(3)]
        for r in ranks:
            r.finalize()
            for k in range(3):
                codes[k][r.y0:r.y1] = r.codes[k][r.y0:r.y1].cpu().numpy()
        got = ranks[0].area_table()
        assert len(got) == len(want["areas"])
        for name in got.dtype.names:
            assert got[name].tobytes() == want["areas"][name].tobytes(), name
        assert ranks[-1].area_table().tobytes() == got.tobytes()
        for k, name in enumerate(("codesA", "codesB", "codesC")):
            assert np.array_equal(codes[k], want[name]), name
    finally:
        for c in codecs:
            c.close()
