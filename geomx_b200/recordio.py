"""RecordIO containers (``mx.recordio``): sequential and indexed record files plus the image-record header.

Parity: ``python/mxnet/recordio.py`` (``MXRecordIO`` :36-210, ``MXIndexedRecordIO`` :213-330, ``IRHeader``/``pack``/``unpack``/``pack_img``/
``unpack_img`` :333-480) over dmlc-core's on-disk format (``3rdparty/dmlc-core/include/dmlc/recordio.h``): every record is
``uint32 magic 0xced7230a | uint32 (cflag << 29 | length) | payload | pad to 4 bytes``; payloads that contain the magic word are split into
continuation records (cflag 1 = first, 2 = middle, 3 = last), which the readers re-assemble.  Writing is Python file IO; reading has a native
path as well (``csrc/runtime/recordio.h``: the file is memory-mapped, ``scan_offsets`` finds the records, ``RecordReader`` fetches them by
offset, several at a time on worker threads) which ``mx.image.ImageIter`` uses for .rec files without an index."""
from __future__ import annotations

import io as _io
import numbers
import struct
from collections import namedtuple

import numpy as np

__all__ = ["MXRecordIO", "MXIndexedRecordIO", "IRHeader", "pack", "unpack", "pack_img", "unpack_img", "RecordReader", "scan_offsets"]

_MAGIC = 0xced7230a
_MAGIC_BYTES = struct.pack("<I", _MAGIC)


class MXRecordIO:
    def __init__(self, uri, flag):
        self.uri, self.flag = str(uri), flag
        self.fp = None
        self.open()

    def open(self):
        if self.flag == "w":
            self.fp = open(self.uri, "wb"); self.writable = True
        elif self.flag == "r":
            self.fp = open(self.uri, "rb"); self.writable = False
        else:
            raise ValueError("Invalid flag %s" % self.flag)
        self.is_open = True

    def close(self):
        if self.fp is not None and not self.fp.closed:
            self.fp.close()
        self.is_open = False

    def __del__(self):
        self.close()

    def __getstate__(self):
        d = dict(self.__dict__); d["fp"] = None; d["is_open"] = False
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        if not self.writable:
            self.open()

    def reset(self):
        self.close(); self.open()

    def tell(self):
        return self.fp.tell()

    def write(self, buf):
        assert self.writable
        buf = bytes(buf)
        # split at embedded magic words (dmlc recordio.h: the magic never appears inside a stored chunk)
        chunks, start = [], 0
        while True:
            pos = buf.find(_MAGIC_BYTES, start)
            while pos != -1 and (pos % 4) != 0:                  # only aligned occurrences matter to a reader scanning words
                pos = buf.find(_MAGIC_BYTES, pos + 1)
            if pos == -1:
                chunks.append(buf[start:]); break
            chunks.append(buf[start:pos]); start = pos + 4
        for i, c in enumerate(chunks):
            cflag = 0 if len(chunks) == 1 else (1 if i == 0 else (3 if i == len(chunks) - 1 else 2))
            self.fp.write(struct.pack("<II", _MAGIC, (cflag << 29) | len(c)))
            self.fp.write(c)
            self.fp.write(b"\x00" * ((4 - len(c) % 4) % 4))

    def read(self):
        assert not self.writable
        out, multi = b"", False
        while True:
            head = self.fp.read(8)
            if len(head) < 8:
                return out if multi else None
            magic, lrec = struct.unpack("<II", head)
            if magic != _MAGIC:
                raise IOError("invalid RecordIO file %s (bad magic at %d)" % (self.uri, self.fp.tell() - 8))
            cflag, length = lrec >> 29, lrec & ((1 << 29) - 1)
            data = self.fp.read(length)
            self.fp.read((4 - length % 4) % 4)
            if cflag == 0:
                return data
            out += data if cflag == 1 else _MAGIC_BYTES + data
            multi = True
            if cflag == 3:
                return out


class MXIndexedRecordIO(MXRecordIO):
    def __init__(self, idx_path, uri, flag, key_type=int):
        self.idx_path, self.idx, self.keys, self.key_type, self.fidx = idx_path, {}, [], key_type, None
        super().__init__(uri, flag)

    def open(self):
        super().open()
        self.idx, self.keys = {}, []
        self.fidx = open(self.idx_path, self.flag)
        if not self.writable:
            for line in iter(self.fidx.readline, ""):
                parts = line.strip().split("\t")
                if len(parts) >= 2:
                    k = self.key_type(parts[0]); self.idx[k] = int(parts[1]); self.keys.append(k)

    def close(self):
        if getattr(self, "fidx", None) is not None and not self.fidx.closed:
            self.fidx.close()
        super().close()

    def seek(self, idx):
        assert not self.writable
        self.fp.seek(self.idx[idx])

    def read_idx(self, idx):
        self.seek(idx)
        return self.read()

    def write_idx(self, idx, buf):
        key = self.key_type(idx)
        pos = self.tell()
        self.write(buf)
        self.fidx.write("%s\t%d\n" % (str(key), pos))
        self.idx[key] = pos; self.keys.append(key)


IRHeader = namedtuple("HEADER", ["flag", "label", "id", "id2"])
_IR_FORMAT = "<IfQQ"
_IR_SIZE = struct.calcsize(_IR_FORMAT)


def pack(header, s):
    header = IRHeader(*header)
    if isinstance(header.label, numbers.Number):
        header = header._replace(flag=0)
        payload = b""
    else:
        label = np.asarray(header.label, dtype=np.float32)
        header = header._replace(flag=label.size, label=0)
        payload = label.tobytes()
    return struct.pack(_IR_FORMAT, *header) + payload + bytes(s)


def unpack(s):
    header = IRHeader(*struct.unpack(_IR_FORMAT, s[:_IR_SIZE]))
    s = s[_IR_SIZE:]
    if header.flag > 0:
        header = header._replace(label=np.frombuffer(s, np.float32, header.flag))
        s = s[header.flag * 4:]
    return header, s


def pack_img(header, img, quality=95, img_fmt=".jpg"):
    """``img``: HxWxC uint8 (RGB) or HxW array, encoded with Pillow (the reference uses OpenCV, which is not in this image)."""
    from PIL import Image
    fmt = {".jpg": "JPEG", ".jpeg": "JPEG", ".png": "PNG"}[img_fmt.lower()]
    buf = _io.BytesIO()
    Image.fromarray(np.asarray(img)).save(buf, format=fmt, **({"quality": quality} if fmt == "JPEG" else {}))
    return pack(header, buf.getvalue())


def unpack_img(s, iscolor=-1):
    from PIL import Image
    header, s = unpack(s)
    img = Image.open(_io.BytesIO(s))
    if iscolor == 0:
        img = img.convert("L")
    elif iscolor == 1:
        img = img.convert("RGB")
    return header, np.asarray(img)


class RecordReader:
    """Random access to the records of a ``.rec`` file by byte offset (native, memory-mapped; falls back to ``MXRecordIO`` seeks when the
    native runtime is not built).  ``offsets`` lists every record; ``read(off)`` / ``read_many(offs)`` return payload bytes."""

    def __init__(self, uri):
        from . import runtime
        self.uri = str(uri)
        self._native = runtime.C().RecordFile(self.uri) if runtime.available() and hasattr(runtime.C(), "RecordFile") else None
        self._py = None
        self._offsets = None

    @property
    def offsets(self):
        if self._offsets is None:
            self._offsets = list(self._native.scan()) if self._native is not None else _scan_python(self.uri)
        return self._offsets

    def __len__(self):
        return len(self.offsets)

    def read(self, offset):
        if self._native is not None:
            return self._native.read(int(offset))
        if self._py is None:
            self._py = MXRecordIO(self.uri, "r")
        self._py.fp.seek(int(offset))
        return self._py.read()

    def read_many(self, offsets, threads=4):
        if self._native is not None:
            return list(self._native.read_many([int(o) for o in offsets], int(threads)))
        return [self.read(o) for o in offsets]


def _scan_python(uri):
    out = []
    with open(uri, "rb") as f:
        while True:
            pos = f.tell()
            head = f.read(8)
            if len(head) < 8:
                return out
            magic, lrec = struct.unpack("<II", head)
            if magic != _MAGIC:
                raise IOError("invalid RecordIO file %s (bad magic at %d)" % (uri, pos))
            cflag, length = lrec >> 29, lrec & ((1 << 29) - 1)
            if cflag in (0, 1):
                out.append(pos)
            f.seek((length + 3) & ~3, 1)


def scan_offsets(uri):
    """Byte offsets of all records of a ``.rec`` file (what ``tools/im2rec.py`` writes into the ``.idx`` file, recomputed from the data)."""
    return RecordReader(uri).offsets
