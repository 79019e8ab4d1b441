"""Minimal GeoTIFF codec for the two ends of the chip path when rasterio / GDAL are not installed
(host side only).

The reference reads chips with ``rasterio.open(path).read()`` (instageo/model/dataloader.py:672-704) and writes
predictions with ``rasterio.open(path, "w", **profile)`` (instageo/model/infer_utils.py:37-54).  rasterio is a
GDAL binding and is absent from this image, so this module restates the TIFF 6.0 / BigTIFF container and the
GeoTIFF 1.1 tags that matter for HLS / Sentinel-2 chips and COG tiles:

* classic TIFF and BigTIFF, little or big endian, strips or tiles, chunky or planar samples;
* compression none (1), Deflate (8 / 32946, zlib), LZW (5, MSB-first 9..12-bit codes with the TIFF "early change");
* predictor 1 (none) and 2 (horizontal differencing, undone per row and per sample with a wrapping cumulative sum);
* 8/16/32/64-bit unsigned, signed and IEEE samples;
* ModelPixelScale (33550), ModelTiepoint (33922), ModelTransformation (34264), GeoKeyDirectory (34735) with its
  double / ASCII parameter tags (34736, 34737) and GDAL_NODATA (42113), carried opaquely so that a prediction written
  with the profile of its source chip has the same georeferencing.

``read_geotiff`` returns ``(array [bands, H, W], profile)``; ``write_geotiff`` writes a little-endian classic TIFF
with one strip per band and block of rows (planar), optional Deflate and predictor 2: host code, pinned by
tests/test_geotiff.py against Pillow (libtiff) and OpenCV.

Device side (16-bit rasters, the HLS / Sentinel-2 case): ``read_geotiff_device`` inflates the blocks on host threads
(zlib releases the GIL) straight into one pinned buffer, ships it once, and ``ig_tiff_unpack16`` (csrc/tiffio.cu) undoes
the predictor, swaps bytes, de-interleaves and crops on the GPU, leaving the planar raster kernel 1 reads in place;
``write_geotiffs_device`` differences a batch of class maps on the GPU (``ig_tiff_predict``) and deflates the strips on
host threads.  ``read_geotiff`` is the oracle of both (tests/test_gpu_tiffio.py).
"""
from __future__ import annotations

import struct
import zlib
from typing import Optional

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 7: ("B", 1), 8: ("h", 2),
          9: ("i", 4), 10: ("ii", 8), 11: ("f", 4), 12: ("d", 8), 13: ("I", 4), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}
_GEO_TAGS = (33550, 33922, 34264, 34735, 34736, 34737, 42113)
_SAMPLE_KIND = {1: "u", 2: "i", 3: "f"}


class TiffError(ValueError):
    pass


def _lzw_decode(data: bytes, expected: int) -> bytes:
    """TIFF LZW (TIFF 6.0 §13): MSB-first variable-width codes, ClearCode 256, EoI 257, width bumped one code
    early ("early change")."""
    out = bytearray()
    table: list = []

    def reset():
        nonlocal table
        table = [bytes([i]) for i in range(256)] + [b"", b""]

    reset()
    bitbuf, nbits, width, prev = 0, 0, 9, None
    for byte in data:
        bitbuf = (bitbuf << 8) | byte
        nbits += 8
        while nbits >= width:
            code = (bitbuf >> (nbits - width)) & ((1 << width) - 1)
            nbits -= width
            if code == 256:
                reset()
                width, prev = 9, None
                continue
            if code == 257:
                return bytes(out[:expected])
            if prev is None:
                entry = table[code]
            elif code < len(table):
                entry = table[code]
                table.append(prev + entry[:1])
            elif code == len(table):
                entry = prev + prev[:1]
                table.append(entry)
            else:
                raise TiffError("corrupt LZW stream")
            out += entry
            prev = entry
            if len(table) >= (1 << width) - 1 and width < 12:
                width += 1
            if len(out) >= expected:
                return bytes(out[:expected])
    return bytes(out[:expected])


def _decompress(raw: bytes, compression: int, expected: int) -> bytes:
    if compression == 1:
        return raw[:expected]
    if compression in (8, 32946):
        return zlib.decompress(raw)[:expected]
    if compression == 5:
        return _lzw_decode(raw, expected)
    raise TiffError(f"unsupported TIFF compression {compression} (supported: none, Deflate, LZW)")


class _Ifd:
    def __init__(self, buf: bytes):
        if buf[:2] == b"II":
            self.e = "<"
        elif buf[:2] == b"MM":
            self.e = ">"
        else:
            raise TiffError("not a TIFF file")
        magic = struct.unpack(self.e + "H", buf[2:4])[0]
        self.big = magic == 43
        if magic not in (42, 43):
            raise TiffError(f"bad TIFF magic {magic}")
        self.buf = buf
        if self.big:
            off = struct.unpack(self.e + "Q", buf[8:16])[0]
            n = struct.unpack(self.e + "Q", buf[off:off + 8])[0]
            pos, esz, vsz, cfmt = off + 8, 20, 8, "Q"
        else:
            off = struct.unpack(self.e + "I", buf[4:8])[0]
            n = struct.unpack(self.e + "H", buf[off:off + 2])[0]
            pos, esz, vsz, cfmt = off + 2, 12, 4, "I"
        self.tags = {}
        for i in range(n):
            ent = buf[pos + i * esz: pos + (i + 1) * esz]
            tag, typ = struct.unpack(self.e + "HH", ent[:4])
            count = struct.unpack(self.e + cfmt, ent[4:4 + vsz])[0]
            if typ not in _TYPES:
                continue
            fmt, size = _TYPES[typ]
            nbytes = size * count
            if nbytes <= vsz:
                data = ent[4 + vsz: 4 + vsz + nbytes]
            else:
                o = struct.unpack(self.e + cfmt, ent[4 + vsz: 4 + 2 * vsz])[0]
                data = buf[o:o + nbytes]
            if typ == 2:
                self.tags[tag] = data.rstrip(b"\x00").decode("latin-1")
            elif typ in (5, 10):
                v = struct.unpack(self.e + fmt[0] * (2 * count), data)
                self.tags[tag] = tuple(v[2 * j] / v[2 * j + 1] if v[2 * j + 1] else 0.0 for j in range(count))
            else:
                self.tags[tag] = struct.unpack(self.e + fmt * count, data)

    def get(self, tag, default=None):
        return self.tags.get(tag, default)

    def one(self, tag, default=None):
        v = self.tags.get(tag)
        return default if v is None else (v[0] if isinstance(v, tuple) else v)


def _undo_predictor(block: np.ndarray) -> None:
    """horizontal differencing (TIFF 6.0 §14): every sample is the difference to its left neighbour of the same
    band; undone in place with a wrapping cumulative sum along x.  block [rows, width, samples]."""
    np.cumsum(block, axis=1, dtype=block.dtype, out=block)


class _Layout:
    """Block geometry of the first image of a TIFF (what both readers need after the IFD is parsed)."""

    def __init__(self, buf: bytes):
        self.ifd = ifd = _Ifd(buf)
        self.W, self.H = ifd.one(256), ifd.one(257)
        self.spp = ifd.one(277, 1)
        bits = ifd.get(258, (8,))
        fmt = ifd.get(339, (1,))
        if len(set(bits)) != 1 or len(set(fmt)) != 1 or bits[0] not in (8, 16, 32, 64) or fmt[0] not in _SAMPLE_KIND:
            raise TiffError(f"unsupported sample layout bits={bits} format={fmt}")
        self.dtype = np.dtype(f"{ifd.e}{_SAMPLE_KIND[fmt[0]]}{bits[0] // 8}")
        self.compression, self.predictor, self.planar = ifd.one(259, 1), ifd.one(317, 1), ifd.one(284, 1)
        if self.predictor not in (1, 2):
            raise TiffError(f"unsupported TIFF predictor {self.predictor}")
        if self.dtype.kind == "f" and self.predictor == 2:
            raise TiffError("predictor 2 on floating-point samples is not defined")
        self.tiled = 322 in ifd.tags
        if self.tiled:
            self.bw, self.bh = ifd.one(322), ifd.one(323)
            self.offsets, self.counts = ifd.get(324), ifd.get(325)
        else:
            self.bw, self.bh = self.W, min(ifd.one(278, self.H), self.H)
            self.offsets, self.counts = ifd.get(273), ifd.get(279)
        self.nbx, self.nby = -(-self.W // self.bw), -(-self.H // self.bh)
        self.planes = self.spp if self.planar == 2 else 1
        self.chunk_spp = 1 if self.planar == 2 else self.spp
        if len(self.offsets) != self.nbx * self.nby * self.planes:
            raise TiffError("block count does not match the image geometry")

    def profile(self, dtype_name: str) -> dict:
        ifd = self.ifd
        profile = {"width": self.W, "height": self.H, "count": self.spp, "dtype": dtype_name,
                   "geo_tags": {t: ifd.tags[t] for t in _GEO_TAGS if t in ifd.tags}}
        scale, tie = ifd.get(33550), ifd.get(33922)
        if scale and tie and len(tie) >= 6:
            # affine (a, b, c, d, e, f): x = a*col + b*row + c, y = d*col + e*row + f  (rasterio's Affine order)
            profile["transform"] = (scale[0], 0.0, tie[3] - tie[0] * scale[0], 0.0, -scale[1], tie[4] + tie[1] * scale[1])
        keys = ifd.get(34735)
        if keys and len(keys) >= 4:
            # ProjectedCSType (3072) wins over GeographicType (2048): files of older GDAL versions carry both for a
            # projected CRS (2048 = the datum's geographic CRS, e.g. 4326 next to 3072 = 326xx of a UTM HLS chip)
            found = {}
            for k in range(keys[3]):
                kid, loc, _cnt, val = keys[4 + 4 * k: 8 + 4 * k]
                if kid in (3072, 2048) and loc == 0 and val not in (0, 32767):
                    found[kid] = val
            if 3072 in found:
                profile["crs_epsg"] = found[3072]
            elif 2048 in found:
                profile["crs_epsg"] = found[2048]
        nod = ifd.get(42113)
        if nod:
            try:
                profile["nodata"] = float(nod)
            except ValueError:
                pass
        return profile


_PINNED: dict = {}


def read_geotiff_device(path: str, device="cuda", threads: int = 8):
    """-> (CUDA tensor [bands, H, W] int16 | uint16, profile): the device form of ``rasterio.open(path).read()``
    (instageo/model/dataloader.py:672-704) for 16-bit rasters.  Host: file read, IFD parse, zlib inflate of the blocks
    on ``threads`` threads into ONE pinned buffer (block-major, still differenced / byte-swapped / interleaved as in
    the file).  Device: a single H2D copy and ``ig_tiff_unpack16``.  Other sample types raise ``TiffError`` (use
    ``read_geotiff``): there is no silent host path behind this entry."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    from .. import _lib
    with open(path, "rb") as fh:
        buf = fh.read()
    lay = _Layout(buf)
    if lay.dtype.itemsize != 2 or lay.dtype.kind not in "ui":
        raise TiffError(f"read_geotiff_device handles 16-bit integer rasters; this file holds {lay.dtype.name} "
                        "(decode it with read_geotiff)")
    if lay.chunk_spp > 8:
        raise TiffError(f"{lay.chunk_spp} pixel-interleaved samples: the device kernel takes at most 8 (planar files are unlimited)")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("read_geotiff_device needs a CUDA device: there is no CPU path (use read_geotiff)")
    block_bytes = lay.bh * lay.bw * lay.chunk_spp * 2
    n_blocks = len(lay.offsets)
    key = (n_blocks * block_bytes, dev.index)
    pinned = _PINNED.get(key)
    if pinned is None:
        if len(_PINNED) >= 4:
            _PINNED.pop(next(iter(_PINNED)))
        pinned = _PINNED[key] = torch.empty(n_blocks * block_bytes, dtype=torch.uint8).pin_memory()
    host = pinned.numpy()

    def inflate(i):
        by = (i // lay.nbx) % lay.nby
        rows = lay.bh if lay.tiled else min(lay.bh, lay.H - by * lay.bh)
        expected = rows * lay.bw * lay.chunk_spp * 2
        raw = _decompress(buf[lay.offsets[i]: lay.offsets[i] + lay.counts[i]], lay.compression, expected)
        if len(raw) < expected:
            raise TiffError("truncated TIFF block")
        host[i * block_bytes: i * block_bytes + expected] = np.frombuffer(raw, dtype=np.uint8, count=expected)

    if n_blocks > 1 and threads > 1:
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(inflate, range(n_blocks)))
    else:
        for i in range(n_blocks):
            inflate(i)
    tdtype = torch.int16 if lay.dtype.kind == "i" else torch.uint16
    with torch.cuda.device(dev):
        d_blocks = pinned.to(dev, non_blocking=True)
        out = torch.empty((lay.spp, lay.H, lay.W), dtype=torch.int16, device=dev)
        _lib.call("ig_tiff_unpack16", dev, d_blocks.data_ptr(), out.data_ptr(), lay.W, lay.H, lay.spp, lay.bw, lay.bh,
                  int(lay.planar == 2), lay.predictor, int(lay.dtype.byteorder == ">"))
        # the pinned buffer is reused by the next read: its H2D copy must have left the host before we return
        torch.cuda.current_stream(dev).synchronize()
    return (out if tdtype == torch.int16 else out.view(torch.uint16)), lay.profile(lay.dtype.newbyteorder("=").name)


def write_geotiffs_device(paths, maps, profiles=None, compress: Optional[str] = "deflate", predictor: int = 2,
                          rows_per_strip: int = 256, threads: int = 8) -> None:
    """Batched writer of single-band predictions (instageo/model/infer_utils.py:37-54 for a whole batch): ``maps`` CUDA
    tensor [n, H, W] int8 | uint8 | int16 | uint16.  The horizontal differencing of predictor 2 runs on the GPU over
    the whole batch (``ig_tiff_predict``), one D2H copy brings the bytes back, the strips are deflated on ``threads``
    host threads and each map is written with its own profile (georeferencing of its source chip)."""
    import torch
    from concurrent.futures import ThreadPoolExecutor

    from .. import _lib
    if not maps.is_cuda or maps.dim() != 3:
        raise RuntimeError("write_geotiffs_device needs a CUDA tensor [n, H, W]: there is no CPU path (use write_geotiff)")
    if maps.dtype not in (torch.int8, torch.uint8, torch.int16, torch.uint16):
        raise TypeError(f"unsupported dtype {maps.dtype}")
    n, H, W = maps.shape
    if len(paths) != n or (profiles is not None and len(profiles) != n):
        raise ValueError("one path (and profile) per map")
    maps = maps.contiguous()
    src = maps
    if predictor == 2:
        src = torch.empty_like(maps)
        _lib.call("ig_tiff_predict", maps.device, maps.data_ptr(), src.data_ptr(), maps.element_size(), n * H, W)
    host = src.cpu().numpy()

    def one(i):
        _write_single_band(paths[i], host[i], (profiles[i] if profiles is not None else None), compress, predictor,
                           rows_per_strip, differenced=predictor == 2)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        list(pool.map(one, range(n)))


def read_geotiff(path: str):
    """-> (array [bands, H, W] in the file's sample type, profile dict)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    ifd = _Ifd(buf)
    W, H = ifd.one(256), ifd.one(257)
    spp = ifd.one(277, 1)
    bits = ifd.get(258, (8,))
    fmt = ifd.get(339, (1,))
    if len(set(bits)) != 1 or len(set(fmt)) != 1 or bits[0] not in (8, 16, 32, 64) or fmt[0] not in _SAMPLE_KIND:
        raise TiffError(f"unsupported sample layout bits={bits} format={fmt}")
    dtype = np.dtype(f"{ifd.e}{_SAMPLE_KIND[fmt[0]]}{bits[0] // 8}")
    compression, predictor, planar = ifd.one(259, 1), ifd.one(317, 1), ifd.one(284, 1)
    if predictor not in (1, 2):
        raise TiffError(f"unsupported TIFF predictor {predictor}")
    if dtype.kind == "f" and predictor == 2:
        raise TiffError("predictor 2 on floating-point samples is not defined")
    tiled = 322 in ifd.tags
    if tiled:
        bw, bh = ifd.one(322), ifd.one(323)
        offsets, counts = ifd.get(324), ifd.get(325)
    else:
        bw, bh = W, min(ifd.one(278, H), H)
        offsets, counts = ifd.get(273), ifd.get(279)
    nbx, nby = -(-W // bw), -(-H // bh)
    planes = spp if planar == 2 else 1
    chunk_spp = 1 if planar == 2 else spp
    if len(offsets) != nbx * nby * planes:
        raise TiffError("block count does not match the image geometry")
    out = np.empty((spp, H, W), dtype=dtype.newbyteorder("="))
    idx = 0
    for p in range(planes):
        for by in range(nby):
            for bx in range(nbx):
                rows = bh if tiled else min(bh, H - by * bh)
                expected = rows * bw * chunk_spp * dtype.itemsize
                raw = _decompress(buf[offsets[idx]: offsets[idx] + counts[idx]], compression, expected)
                idx += 1
                if len(raw) < expected:
                    raise TiffError("truncated TIFF block")
                blk = np.frombuffer(raw, dtype=dtype, count=rows * bw * chunk_spp).reshape(rows, bw, chunk_spp)
                blk = blk.astype(dtype.newbyteorder("="))
                if predictor == 2:
                    _undo_predictor(blk)
                y0, x0 = by * bh, bx * bw
                h, w = min(rows, H - y0), min(bw, W - x0)
                if planar == 2:
                    out[p, y0:y0 + h, x0:x0 + w] = blk[:h, :w, 0]
                else:
                    out[:, y0:y0 + h, x0:x0 + w] = blk[:h, :w, :].transpose(2, 0, 1)
    profile = {"width": W, "height": H, "count": spp, "dtype": out.dtype.name,
               "geo_tags": {t: ifd.tags[t] for t in _GEO_TAGS if t in ifd.tags}}
    scale, tie = ifd.get(33550), ifd.get(33922)
    if scale and tie and len(tie) >= 6:
        # affine (a, b, c, d, e, f): x = a*col + b*row + c, y = d*col + e*row + f  (rasterio's Affine order)
        profile["transform"] = (scale[0], 0.0, tie[3] - tie[0] * scale[0], 0.0, -scale[1], tie[4] + tie[1] * scale[1])
    keys = ifd.get(34735)
    if keys and len(keys) >= 4:
        # ProjectedCSType (3072) wins over GeographicType (2048): files written by older GDAL versions carry both for
        # a projected CRS (2048 = the datum's geographic CRS, e.g. 4326, next to 3072 = 326xx of a UTM HLS chip)
        found = {}
        for k in range(keys[3]):
            kid, loc, _cnt, val = keys[4 + 4 * k: 8 + 4 * k]
            if kid in (3072, 2048) and loc == 0 and val not in (0, 32767):
                found[kid] = val
        if 3072 in found:
            profile["crs_epsg"] = found[3072]
        elif 2048 in found:
            profile["crs_epsg"] = found[2048]
    nod = ifd.get(42113)
    if nod:
        try:
            profile["nodata"] = float(nod)
        except ValueError:
            pass
    return out, profile


def _write_single_band(path, plane: np.ndarray, profile, compress, predictor, rows_per_strip, differenced: bool) -> None:
    """write_geotiff for one [H, W] plane whose rows are ALREADY horizontally differenced (device writer)."""
    write_geotiff(path, plane, profile, compress=compress, predictor=predictor, rows_per_strip=rows_per_strip,
                  _differenced=differenced)


def write_geotiff(path: str, array: np.ndarray, profile: Optional[dict] = None, compress: Optional[str] = "deflate",
                  predictor: int = 1, rows_per_strip: int = 256, _differenced: bool = False) -> None:
    """array [bands, H, W] or [H, W] -> little-endian classic TIFF, planar strips.  ``profile["geo_tags"]`` (as
    returned by ``read_geotiff``) is copied, so a prediction inherits the georeferencing of its source chip."""
    a = np.asarray(array)
    if a.ndim == 2:
        a = a[None]
    if a.ndim != 3:
        raise ValueError("array must be [bands, H, W] or [H, W]")
    if a.dtype.kind not in "uif" or a.dtype.itemsize not in (1, 2, 4, 8):
        raise TypeError(f"unsupported dtype {a.dtype}")
    if compress not in (None, "none", "deflate"):
        raise ValueError("compress must be None or 'deflate'")
    if predictor == 2 and a.dtype.kind == "f":
        raise ValueError("predictor 2 needs integer samples")
    bands, H, W = a.shape
    le = a.astype(a.dtype.newbyteorder("<"), copy=False)
    rows_per_strip = max(1, min(rows_per_strip, H))
    strips = []
    for b in range(bands):
        for y0 in range(0, H, rows_per_strip):
            blk = np.ascontiguousarray(le[b, y0:y0 + rows_per_strip])
            if predictor == 2 and not _differenced:
                d = blk.copy()
                d[:, 1:] = blk[:, 1:] - blk[:, :-1]       # wraps for integer types, as the decoder expects
                blk = d
            raw = blk.tobytes()
            strips.append(zlib.compress(raw, 6) if compress == "deflate" else raw)
    kind = {"u": 1, "i": 2, "f": 3}[a.dtype.kind]
    entries = {256: (4, (W,)), 257: (4, (H,)), 258: (3, (a.dtype.itemsize * 8,) * bands),
               259: (3, (8 if compress == "deflate" else 1,)), 262: (3, (1,)), 277: (3, (bands,)),
               278: (4, (rows_per_strip,)), 284: (3, (2,)), 339: (3, (kind,) * bands)}
    if bands > 1:
        entries[338] = (3, (0,) * (bands - 1))    # ExtraSamples: unspecified
    if predictor == 2:
        entries[317] = (3, (2,))
    for tag, val in ((profile or {}).get("geo_tags") or {}).items():
        if tag in (33550, 33922, 34264, 34736):
            entries[tag] = (12, tuple(float(v) for v in val))
        elif tag == 34735:
            entries[tag] = (3, tuple(int(v) for v in val))
        elif tag in (34737, 42113):
            entries[tag] = (2, str(val))
    entries[273] = (4, (0,) * len(strips))
    entries[279] = (4, tuple(len(s) for s in strips))
    # layout: header (8) | IFD | out-of-line tag values | strips
    tags = sorted(entries)
    ifd_size = 2 + 12 * len(tags) + 4
    blobs, pos = {}, 8 + ifd_size

    def pack(typ, val):
        if typ == 2:
            return val.encode("latin-1") + b"\x00"
        fmt = {3: "H", 4: "I", 12: "d"}[typ]
        return struct.pack("<" + fmt * len(val), *val)

    sizes = {t: len(pack(*entries[t])) for t in tags}
    for t in tags:
        if sizes[t] > 4:
            blobs[t] = pos
            pos += sizes[t] + (sizes[t] & 1)
    offs, data_pos = [], pos
    for s in strips:
        offs.append(data_pos)
        data_pos += len(s) + (len(s) & 1)
    if data_pos >= 2 ** 32:
        raise TiffError("image too large for a classic TIFF")
    entries[273] = (4, tuple(offs))
    with open(path, "wb") as fh:
        fh.write(struct.pack("<2sHI", b"II", 42, 8))
        fh.write(struct.pack("<H", len(tags)))
        for t in tags:
            typ, val = entries[t]
            payload = pack(typ, val)
            count = len(val) + 1 if typ == 2 else len(val)
            fh.write(struct.pack("<HHI", t, typ, count))
            fh.write(payload.ljust(4, b"\x00") if len(payload) <= 4 else struct.pack("<I", blobs[t]))
        fh.write(struct.pack("<I", 0))
        for t in tags:
            if t in blobs:
                payload = pack(*entries[t])
                fh.write(payload + (b"\x00" if len(payload) & 1 else b""))
        for s in strips:
            fh.write(s + (b"\x00" if len(s) & 1 else b""))
