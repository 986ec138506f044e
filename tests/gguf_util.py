"""Minimal independent GGUF v3 writer/reader for the tests (not the product's parser)."""
import struct

import numpy as np

TYPE_SIZE = {0: (1, 4), 1: (1, 2), 2: (32, 18), 3: (32, 20), 6: (32, 22), 7: (32, 24), 8: (32, 34), 9: (32, 36), 10: (256, 84),
             11: (256, 110), 12: (256, 144), 13: (256, 176), 14: (256, 210), 15: (256, 290), 30: (1, 2), 26: (1, 4),
             16: (256, 66), 17: (256, 74), 18: (256, 194), 19: (256, 50), 20: (32, 34), 21: (256, 110), 22: (256, 82), 23: (256, 264),
             29: (256, 56)}
U32, STRING, ARRAY, U64 = 4, 8, 9, 10


def _s(x):
    b = x.encode()
    return struct.pack("<Q", len(b)) + b


def kv_bytes(key, ty, value):
    out = _s(key) + struct.pack("<I", ty)
    if ty == STRING:
        out += _s(value)
    elif ty == U32:
        out += struct.pack("<I", value)
    elif ty == U64:
        out += struct.pack("<Q", value)
    elif ty == ARRAY:  # (elem_type, list)
        et, items = value
        out += struct.pack("<IQ", et, len(items))
        for it in items:
            out += _s(it) if et == STRING else struct.pack("<I", it)
    else:
        raise ValueError(ty)
    return out


def write_gguf(path, kvs, tensors, alignment=32, alignment_first=False, version=3, magic=b"GGUF"):
    """kvs: list of (key, type, value); tensors: list of (name, shape, ggml_type, bytes)."""
    n_kv = len(kvs)
    body = b""
    for k, ty, v in kvs:
        body += kv_bytes(k, ty, v)
    infos, off, offsets = b"", 0, []
    for name, shape, ty, data in tensors:
        off += (alignment - off % alignment) % alignment
        offsets.append(off)
        infos += _s(name) + struct.pack("<I", len(shape)) + b"".join(struct.pack("<Q", d) for d in shape) + struct.pack("<IQ", ty, off)
        off += len(data)
    head = magic + struct.pack("<IQQ", version, len(tensors), n_kv) + body + infos
    blob = bytearray(head)
    if tensors:
        blob += b"\0" * ((alignment - len(blob) % alignment) % alignment)
    base = len(blob)
    for (name, shape, ty, data), o in zip(tensors, offsets):
        blob += b"\0" * (base + o - len(blob))
        blob += bytes(data)
    with open(path, "wb") as f:
        f.write(blob)


def index_gguf(path):
    """Like read_gguf but without copying tensor data: (tensors: dict name -> (shape, type, absolute offset, nbytes),
    alignment, uint8 memmap of the file)."""
    import mmap
    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    _, tensors, alignment, _ = read_gguf(path, _buf=mm, _index_only=True)
    return tensors, alignment, np.frombuffer(mm, np.uint8)


def read_gguf(path, _buf=None, _index_only=False):
    """Returns (kvs: list of (key, type, raw value bytes), tensors: dict name -> (shape, type, bytes), alignment)."""
    buf = open(path, "rb").read() if _buf is None else _buf
    assert buf[:4] == b"GGUF"
    ver, nt, nkv = struct.unpack_from("<IQQ", buf, 4)
    assert ver == 3
    p = 24

    def rs():
        nonlocal p
        (n,) = struct.unpack_from("<Q", buf, p)
        s = bytes(buf[p + 8:p + 8 + n]).decode()
        p += 8 + n
        return s

    SC = {0: 1, 1: 1, 2: 2, 3: 2, 4: 4, 5: 4, 6: 4, 7: 1, 10: 8, 11: 8, 12: 8}

    def skip(ty):
        nonlocal p
        if ty == STRING:
            rs()
        elif ty == ARRAY:
            et, n = struct.unpack_from("<IQ", buf, p)
            p += 12
            for _ in range(n):
                skip(et)
        else:
            p += SC[ty]
    kvs, alignment = [], 32
    for _ in range(nkv):
        k = rs()
        (ty,) = struct.unpack_from("<I", buf, p)
        p += 4
        v0 = p
        skip(ty)
        kvs.append((k, ty, buf[v0:p]))
        if k == "general.alignment":
            alignment = struct.unpack("<I", buf[v0:p])[0]
    infos = []
    for _ in range(nt):
        name = rs()
        (nd,) = struct.unpack_from("<I", buf, p)
        p += 4
        shape = struct.unpack_from("<%dQ" % nd, buf, p)
        p += 8 * nd
        ty, off = struct.unpack_from("<IQ", buf, p)
        p += 12
        infos.append((name, shape, ty, off))
    p += (alignment - p % alignment) % alignment if nt else 0
    tensors = {}
    for name, shape, ty, off in infos:
        e, b = TYPE_SIZE[ty]
        n = int(np.prod(shape)) // e * b
        tensors[name] = (shape, ty, p + off, n) if _index_only else (shape, ty, buf[p + off:p + off + n])
        assert off % alignment == 0
    return kvs, tensors, alignment, len(buf)
