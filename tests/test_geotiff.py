"""CPU: the in-repo GeoTIFF codec (instageo_b200/data/geotiff.py, the host side of SURVEY.md §8(f) row 2) against two
independent TIFF implementations that are installed -- Pillow (libtiff) and OpenCV -- in both directions, against the
reference's own raster fixtures when the reference tree is mounted, and through the two call sites it serves
(``get_raster_data``, ``save_prediction``)."""
import os

import numpy as np
import pytest

from instageo_b200.data import geotiff as G


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.uint16, np.int32, np.float32, np.float64])
@pytest.mark.parametrize("bands", [1, 6, 18])
def test_own_round_trip(tmp_path, dtype, bands):
    rng = np.random.default_rng(bands)
    a = rng.normal(size=(bands, 70, 53)).astype(dtype) if np.dtype(dtype).kind == "f" else \
        rng.integers(np.iinfo(dtype).min, int(np.iinfo(dtype).max) + 1, size=(bands, 70, 53)).astype(dtype)
    for comp in (None, "deflate"):
        for pred in ((1,) if np.dtype(dtype).kind == "f" else (1, 2)):
            p = str(tmp_path / "a.tif")
            G.write_geotiff(p, a, compress=comp, predictor=pred, rows_per_strip=16)
            b, prof = G.read_geotiff(p)
            assert b.dtype == a.dtype and np.array_equal(a, b)          # full integer range: the predictor wraps
            assert (prof["width"], prof["height"], prof["count"]) == (53, 70, bands)


def test_pillow_reads_what_we_write_and_back(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(1)
    for dtype in (np.uint8, np.uint16, np.int32, np.float32):
        a = rng.normal(size=(40, 31)).astype(dtype) if np.dtype(dtype).kind == "f" else rng.integers(0, 250, size=(40, 31)).astype(dtype)
        p = str(tmp_path / "b.tif")
        G.write_geotiff(p, a, compress="deflate")
        assert np.array_equal(np.array(Image.open(p)), a)
    a = rng.integers(0, 60000, size=(64, 48)).astype(np.uint16)
    for comp in ("raw", "tiff_lzw", "tiff_adobe_deflate", "tiff_deflate"):        # libtiff-written strips
        for pred in (1, 2):
            p = str(tmp_path / "c.tif")
            Image.fromarray(a).save(p, compression=comp, tiffinfo={317: pred} if comp != "raw" else {})
            b, _ = G.read_geotiff(p)
            assert b.shape == (1, 64, 48) and np.array_equal(b[0], a), (comp, pred)


def test_opencv_chunky_multiband(tmp_path):
    cv2 = pytest.importorskip("cv2")
    a = np.random.default_rng(2).integers(0, 60000, size=(50, 40, 3)).astype(np.uint16)
    p = str(tmp_path / "d.tif")
    cv2.imwrite(p, a)                                  # chunky RGB samples, LZW + predictor 2; OpenCV stores BGR as RGB
    b, _ = G.read_geotiff(p)
    assert np.array_equal(b.transpose(1, 2, 0)[..., ::-1], a)


def test_georeferencing_survives_a_prediction(tmp_path):
    """chip -> save_prediction: the single-band int8 prediction carries the source chip's tiepoint / scale / GeoKeys,
    and get_raster_data returns what rasterio's read() would ([bands, H, W], band gather for non-labels)."""
    from instageo_b200.model.dataloader import get_raster_data
    from instageo_b200.model.infer_utils import save_prediction
    geo = {33550: (30.0, 30.0, 0.0), 33922: (0.0, 0.0, 0.0, 318585.0, 4583115.0, 0.0),
           34735: (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 32613), 42113: "-9999"}
    chip = np.random.default_rng(3).integers(-100, 10000, size=(18, 224, 224)).astype(np.int16)
    src = str(tmp_path / "chip_001.tif")
    G.write_geotiff(src, chip, {"geo_tags": geo}, compress="deflate", predictor=2)
    back, prof = G.read_geotiff(src)
    assert np.array_equal(back, chip) and prof["crs_epsg"] == 32613 and prof["nodata"] == -9999.0
    assert prof["transform"] == (30.0, 0.0, 318585.0, 0.0, -30.0, 4583115.0)
    assert np.array_equal(get_raster_data(src, is_label=False, bands=[3, 0, 17]), chip[[3, 0, 17]])
    assert np.array_equal(get_raster_data(src, is_label=True, bands=[1]), chip)          # labels: no band gather
    out_dir = tmp_path / "out"
    out_dir.mkdir()
    pred = np.random.default_rng(4).integers(-1, 2, size=(224, 224)).astype(np.int8)
    save_prediction(pred, src, str(out_dir))
    got, gprof = G.read_geotiff(str(out_dir / "prediction_001.tif"))
    assert got.dtype == np.int8 and np.array_equal(got[0], pred)
    assert gprof["transform"] == prof["transform"] and gprof["crs_epsg"] == 32613
    save_prediction(pred, "chip_without_file.tif", str(out_dir))                        # nothing to inherit: .npy
    assert np.array_equal(np.load(out_dir / "prediction_without_file.npy"), pred)


def test_rejects_what_it_does_not_decode(tmp_path):
    p = tmp_path / "x.tif"
    p.write_bytes(b"not a tiff at all")
    with pytest.raises(G.TiffError):
        G.read_geotiff(str(p))
    with pytest.raises(ValueError):
        G.write_geotiff(str(p), np.zeros((2, 2), np.float32), predictor=2)
    with pytest.raises(TypeError):
        G.write_geotiff(str(p), np.zeros((2, 2), np.complex64))


@pytest.mark.skipif(not os.path.exists("/root/reference/tests/data/sample.tif"), reason="reference not mounted")
def test_reference_fixtures_decode_like_opencv():
    cv2 = pytest.importorskip("cv2")
    for name, epsg in (("sample.tif", 32613), ("fmask.tif", 32638)):
        path = os.path.join("/root/reference/tests/data", name)
        a, prof = G.read_geotiff(path)
        want = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        assert a.shape == (1, 224, 224) and a.dtype == want.dtype and np.array_equal(a[0], want, equal_nan=True)
        assert prof["crs_epsg"] == epsg and prof["transform"][0] == 30.0 and prof["transform"][4] == -30.0


def _handmade_tiff(path, a, *, tile=None, big=False, endian="<", deflate=True, predictor=1):
    """A single-band TIFF assembled straight from the TIFF 6.0 / BigTIFF layout rules (independent of the writer under
    test): tiles or one strip, chosen byte order, optional Deflate / predictor 2."""
    import struct
    import zlib
    H, W = a.shape
    dt = a.dtype.newbyteorder(endian)
    th, tw = tile if tile else (H, W)
    blocks = []
    for y0 in range(0, H, th):
        for x0 in range(0, W, tw):
            blk = np.zeros((th, tw), dtype=dt)
            sub = a[y0:y0 + th, x0:x0 + tw]
            blk[:sub.shape[0], :sub.shape[1]] = sub
            if predictor == 2:
                d = blk.copy()
                d[:, 1:] = blk[:, 1:] - blk[:, :-1]
                blk = d
            raw = blk.tobytes()
            blocks.append(zlib.compress(raw) if deflate else raw)
    o, c, hdr = ("Q", "Q", 16) if big else ("I", "I", 8)
    tags = [(256, 4, [W]), (257, 4, [H]), (258, 3, [a.dtype.itemsize * 8]), (259, 3, [8 if deflate else 1]), (262, 3, [1]),
            (277, 3, [1]), (317, 3, [predictor]), (339, 3, [{"u": 1, "i": 2, "f": 3}[a.dtype.kind]])]
    if tile:
        tags += [(322, 4, [tw]), (323, 4, [th]), (324, 16 if big else 4, None), (325, 16 if big else 4, [len(b) for b in blocks])]
    else:
        tags += [(278, 4, [H]), (273, 16 if big else 4, None), (279, 16 if big else 4, [len(b) for b in blocks])]
    tags.sort()
    esz, inline = (20, 8) if big else (12, 4)
    ifd_len = (8 if big else 2) + esz * len(tags) + (8 if big else 4)
    fmt = {3: "H", 4: "I", 16: "Q"}
    size = {3: 2, 4: 4, 16: 8}
    pos = hdr + ifd_len
    extra = {}
    for t, typ, val in tags:
        n = len(blocks) if val is None else len(val)
        if n * size[typ] > inline:
            extra[t] = pos
            pos += n * size[typ]
    offs, p = [], pos
    for b in blocks:
        offs.append(p)
        p += len(b)
    with open(path, "wb") as fh:
        fh.write((b"II" if endian == "<" else b"MM") + struct.pack(endian + "H", 43 if big else 42))
        fh.write(struct.pack(endian + "HHQ", 8, 0, 16) if big else struct.pack(endian + "I", 8))
        fh.write(struct.pack(endian + ("Q" if big else "H"), len(tags)))
        for t, typ, val in tags:
            val = offs if val is None else val
            fh.write(struct.pack(endian + "HH" + c, t, typ, len(val)))
            payload = struct.pack(endian + fmt[typ] * len(val), *val)
            fh.write(payload.ljust(inline, b"\x00") if t not in extra else struct.pack(endian + o, extra[t]))
        fh.write(struct.pack(endian + o, 0))
        for t, typ, val in tags:
            if t in extra:
                val = offs if val is None else val
                fh.write(struct.pack(endian + fmt[typ] * len(val), *val))
        for b in blocks:
            fh.write(b)


@pytest.mark.parametrize("tile,big,endian,predictor", [((64, 128), False, "<", 2), ((32, 48), False, ">", 1),
                                                       (None, True, "<", 2), ((64, 64), True, ">", 2), (None, False, ">", 2)])
def test_tiles_bigtiff_and_byte_order(tmp_path, tile, big, endian, predictor):
    """tiled layouts (COG-style), BigTIFF offsets and big-endian files, assembled by hand from the format rules;
    OpenCV (libtiff) must read the same pixels"""
    cv2 = pytest.importorskip("cv2")
    a = np.random.default_rng(5).integers(-30000, 30000, size=(150, 210)).astype(np.int16)
    p = str(tmp_path / "h.tif")
    _handmade_tiff(p, a, tile=tile, big=big, endian=endian, predictor=predictor)
    got, prof = G.read_geotiff(p)
    assert got.dtype == np.int16 and np.array_equal(got[0], a)
    want = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert want is not None and np.array_equal(want, a)        # the handmade file itself is a valid TIFF


def test_projected_crs_wins_over_geographic_key(tmp_path):
    """GeoKey directories of older GDAL versions carry GeographicType (2048 = 4326) next to ProjectedCSType (3072 = the
    UTM zone) for a projected CRS; keys are sorted by id, so 2048 comes first -- the profile must still report the UTM
    code, on the host reader and on the layout shared with the device reader."""
    a = np.arange(64, dtype=np.int16).reshape(8, 8)
    geo = {33550: (30.0, 30.0, 0.0), 33922: (0.0, 0.0, 0.0, 318585.0, 4583115.0, 0.0),
           34735: (1, 1, 0, 4, 1024, 0, 1, 1, 1025, 0, 1, 1, 2048, 0, 1, 4326, 3072, 0, 1, 32613)}
    p = str(tmp_path / "both.tif")
    G.write_geotiff(p, a, {"geo_tags": geo})
    assert G.read_geotiff(p)[1]["crs_epsg"] == 32613
    assert G._Layout(open(p, "rb").read()).profile("int16")["crs_epsg"] == 32613
    geo[34735] = (1, 1, 0, 2, 1024, 0, 1, 2, 2048, 0, 1, 4326)          # geographic only
    G.write_geotiff(p, a, {"geo_tags": geo})
    assert G.read_geotiff(p)[1]["crs_epsg"] == 4326
    geo[34735] = (1, 1, 0, 3, 1024, 0, 1, 1, 2048, 0, 1, 4326, 3072, 0, 1, 32767)   # user-defined projection: fall back
    G.write_geotiff(p, a, {"geo_tags": geo})
    assert G.read_geotiff(p)[1]["crs_epsg"] == 4326
