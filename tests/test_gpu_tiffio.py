"""GPU parity: the device side of raw-tile ingestion and prediction writing (csrc/tiffio.cu)
against the host codec ``read_geotiff`` -- itself pinned against libtiff (Pillow, OpenCV), handmade files and the
reference's raster fixtures in tests/test_geotiff.py.  Bar: bit-exact (16-bit integer samples)."""
import os

import numpy as np
import pytest
import torch

from instageo_b200.data import geotiff as G
from test_geotiff import _handmade_tiff

pytestmark = pytest.mark.gpu


def _chunky_tiff(path, a, *, endian="<", predictor=2, rows_per_strip=37, tile=None):
    """[bands, H, W] -> pixel-interleaved (PlanarConfiguration 1) classic TIFF, Deflate strips or tiles, assembled from
    the format rules (the repo's writer only emits planar strips)."""
    import struct
    import zlib
    bands, H, W = a.shape
    dt = a.dtype.newbyteorder(endian)
    px = np.ascontiguousarray(a.transpose(1, 2, 0))
    th, tw = tile if tile else (rows_per_strip, W)
    blocks = []
    for y0 in range(0, H, th):
        for x0 in range(0, W, tw):
            rows = th if tile else min(th, H - y0)
            blk = np.zeros((rows, tw, bands), dtype=dt)
            sub = px[y0:y0 + rows, x0:x0 + tw]
            blk[:sub.shape[0], :sub.shape[1]] = sub
            if predictor == 2:
                d = blk.copy()
                d[:, 1:] = blk[:, 1:] - blk[:, :-1]
                blk = d
            blocks.append(zlib.compress(blk.tobytes()))
    tags = [(256, 4, [W]), (257, 4, [H]), (258, 3, [16] * bands), (259, 3, [8]), (262, 3, [1]), (277, 3, [bands]),
            (284, 3, [1]), (317, 3, [predictor]), (339, 3, [{"u": 1, "i": 2}[a.dtype.kind]] * bands)]
    if tile:
        tags += [(322, 4, [tw]), (323, 4, [th]), (324, 4, None), (325, 4, [len(b) for b in blocks])]
    else:
        tags += [(278, 4, [th]), (273, 4, None), (279, 4, [len(b) for b in blocks])]
    tags.sort()
    fmt, size = {3: "H", 4: "I"}, {3: 2, 4: 4}
    pos = 8 + 2 + 12 * len(tags) + 4
    extra = {}
    for t, typ, val in tags:
        n = len(blocks) if val is None else len(val)
        if n * size[typ] > 4:
            extra[t] = pos
            pos += n * size[typ] + (n * size[typ] & 1)
    offs, p = [], pos
    for b in blocks:
        offs.append(p)
        p += len(b) + (len(b) & 1)
    with open(path, "wb") as fh:
        fh.write((b"II" if endian == "<" else b"MM") + struct.pack(endian + "HI", 42, 8))
        fh.write(struct.pack(endian + "H", len(tags)))
        for t, typ, val in tags:
            val = offs if val is None else val
            payload = struct.pack(endian + fmt[typ] * len(val), *val)
            fh.write(struct.pack(endian + "HHI", t, typ, len(val)))
            fh.write(payload.ljust(4, b"\x00") if t not in extra else struct.pack(endian + "I", extra[t]))
        fh.write(struct.pack(endian + "I", 0))
        for t, typ, val in tags:
            if t in extra:
                val = offs if val is None else val
                payload = struct.pack(endian + fmt[typ] * len(val), *val)
                fh.write(payload + (b"\x00" if len(payload) & 1 else b""))
        for b in blocks:
            fh.write(b + (b"\x00" if len(b) & 1 else b""))


@pytest.mark.parametrize("dtype,bands,H,W,pred,rps", [(np.int16, 6, 224, 224, 2, 64), (np.uint16, 18, 224, 224, 2, 256),
                                                      (np.int16, 1, 70, 53, 1, 16), (np.uint16, 3, 150, 211, 2, 37),
                                                      (np.int16, 6, 513, 1030, 2, 100)])
def test_planar_strips_match_host_codec(cuda_dev, tmp_path, dtype, bands, H, W, pred, rps):
    rng = np.random.default_rng(H + W)
    a = rng.integers(np.iinfo(dtype).min, int(np.iinfo(dtype).max) + 1, size=(bands, H, W)).astype(dtype)  # full range: wraps
    p = str(tmp_path / "a.tif")
    G.write_geotiff(p, a, {"geo_tags": {42113: "-9999"}}, compress="deflate", predictor=pred, rows_per_strip=rps)
    want, wprof = G.read_geotiff(p)
    got, prof = G.read_geotiff_device(p, cuda_dev)
    assert got.is_cuda and tuple(got.shape) == (bands, H, W)
    assert got.dtype == (torch.int16 if dtype == np.int16 else torch.uint16)
    assert np.array_equal(got.cpu().view(torch.int16).numpy().view(dtype), want) and np.array_equal(want, a)
    assert prof == wprof


@pytest.mark.parametrize("endian,pred,tile", [("<", 2, None), (">", 2, None), ("<", 1, (64, 128)), (">", 2, (32, 48)),
                                              ("<", 2, (128, 256))])
def test_chunky_tiles_and_byte_order_match_host_codec(cuda_dev, tmp_path, endian, pred, tile):
    """pixel-interleaved samples (what GDAL writes by default), strips and tiles with partial edge blocks, both byte
    orders: the device unpack equals the host codec (which test_geotiff.py pins against libtiff)"""
    a = np.random.default_rng(7).integers(-32768, 32768, size=(6, 301, 333)).astype(np.int16)
    p = str(tmp_path / "c.tif")
    _chunky_tiff(p, a, endian=endian, predictor=pred, tile=tile)
    want, _ = G.read_geotiff(p)
    assert np.array_equal(want, a)
    got, _ = G.read_geotiff_device(p, cuda_dev, threads=3)
    assert np.array_equal(got.cpu().numpy(), a)
    # single band, BigTIFF / tiles from the other handmade writer
    b = a[0]
    q = str(tmp_path / "h.tif")
    _handmade_tiff(q, b, tile=tile, big=True, endian=endian, predictor=pred)
    assert np.array_equal(G.read_geotiff_device(q, cuda_dev)[0][0].cpu().numpy(), b)


def test_device_reader_feeds_kernel_1_and_rejects_what_it_cannot_decode(cuda_dev, tmp_path):
    from instageo_b200 import ops
    from instageo_b200.model.dataloader import get_raster_data
    from conftest import FLOOD_MEAN, FLOOD_STD
    from oracle import preprocess as OP
    raw = OP.synth_chips(1, 3, seed=9)[0]
    p = str(tmp_path / "chip_000.tif")
    G.write_geotiff(p, raw, compress="deflate", predictor=2)
    bands = [5, 4, 3, 2, 1, 0] + list(range(6, 18))
    d = get_raster_data(p, is_label=False, bands=bands, device=cuda_dev)
    assert d.is_cuda and np.array_equal(d.cpu().numpy(), raw[bands])
    spec = ops.PreprocessSpec(FLOOD_MEAN, FLOOD_STD, 3, None, 1e-4, -9999, cuda_dev)
    got = ops.preprocess(d.contiguous(), spec, want_f32=True)["f32"][0].cpu().numpy()
    assert np.array_equal(got, OP.preprocess_chip(raw[bands], None, 1e-4, FLOOD_MEAN, FLOOD_STD, 3, -9999)[0])
    f = str(tmp_path / "f.tif")
    G.write_geotiff(f, np.zeros((1, 8, 8), np.float32))
    with pytest.raises(G.TiffError):
        G.read_geotiff_device(f, cuda_dev)
    with pytest.raises(RuntimeError, match="no CPU path"):
        G.read_geotiff_device(p, "cpu")


@pytest.mark.parametrize("dtype", [torch.int8, torch.uint8, torch.int16])
def test_batched_device_writer_round_trips(cuda_dev, tmp_path, dtype):
    """class maps [n, H, W] -> predictor-2 differencing on the GPU -> Deflate on host threads -> files that the host
    codec (and libtiff via OpenCV) decode to the same maps, georeferencing carried per file"""
    cv2 = pytest.importorskip("cv2")
    g = torch.Generator().manual_seed(3)
    lo, hi = (-1, 13) if dtype == torch.int8 else (0, 200)
    maps = torch.randint(lo, hi, (5, 224, 224), generator=g, dtype=torch.int32).to(dtype)
    geo = {33550: (30.0, 30.0, 0.0), 33922: (0.0, 0.0, 0.0, 318585.0, 4583115.0, 0.0),
           34735: (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32613)}
    paths = [str(tmp_path / f"prediction_{i}.tif") for i in range(5)]
    profs = [{"geo_tags": {**geo, 33922: (0.0, 0.0, 0.0, 318585.0 + 6720.0 * i, 4583115.0, 0.0)}} for i in range(5)]
    G.write_geotiffs_device(paths, maps.to(cuda_dev), profs, predictor=2, rows_per_strip=100, threads=4)
    for i, p in enumerate(paths):
        got, prof = G.read_geotiff(p)
        assert np.array_equal(got[0], maps[i].numpy()) and prof["crs_epsg"] == 32613
        assert prof["transform"][2] == 318585.0 + 6720.0 * i
        if dtype != torch.int8:
            assert np.array_equal(cv2.imread(p, cv2.IMREAD_UNCHANGED), maps[i].numpy())
    # no predictor: plain strips
    G.write_geotiffs_device(paths[:2], maps[:2].to(cuda_dev), None, predictor=1)
    assert np.array_equal(G.read_geotiff(paths[1])[0][0], maps[1].numpy())


def test_unpack_full_tile_properties(cuda_dev):
    """3660 x 3660 x 6 (BASELINE config 4's raster) straight through ig_tiff_unpack16: differencing on the host, then
    the device prefix sum must return the original bits; chunky and planar."""
    from instageo_b200 import _lib
    H = W = 3660
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(-32768, 32768, (6, H, W), generator=g, device=cuda_dev, dtype=torch.int32).to(torch.int16)
    for planar in (1, 0):
        src = a if planar else a.permute(1, 2, 0).contiguous()             # [6, H, W] planes | [H, W, 6] pixels
        diff = src.clone()
        if planar:
            diff[:, :, 1:] = src[:, :, 1:] - src[:, :, :-1]
        else:
            diff[:, 1:, :] = src[:, 1:, :] - src[:, :-1, :]
        out = torch.empty((6, H, W), dtype=torch.int16, device=cuda_dev)
        rps = 128   # strips of 128 rows (the last one short): blocks at a fixed stride, as read_geotiff_device lays them out
        nby = -(-H // rps)
        if planar:
            blocks = torch.zeros((6, nby * rps, W), dtype=torch.int16, device=cuda_dev)
            blocks[:, :H] = diff
        else:
            blocks = torch.zeros((nby * rps, W, 6), dtype=torch.int16, device=cuda_dev)
            blocks[:H] = diff
        _lib.call("ig_tiff_unpack16", cuda_dev, blocks.data_ptr(), out.data_ptr(), W, H, 6, W, rps, planar, 2, 0)
        assert torch.equal(out, a)
