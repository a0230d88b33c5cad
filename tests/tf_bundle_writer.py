"""Test-only writer of TensorFlow V2 checkpoints ("tensor bundles"), following the published on-disk layout
(tensorflow/core/util/tensor_bundle + LevelDB's table format): sorted keys, prefix compression with restart points,
several data blocks, an index block, a 48-byte footer ending in the table magic, block trailers = type byte + masked
crc32c.  Used to exercise ai-cv-automation-elect-micr_b200/tfckpt.py; no TensorFlow here, so this is format-level
verification only (the reader's header says the same)."""
import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
DTYPE_ENUM = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def _crc32c_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_T = _crc32c_table()


def crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c = _T[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def field(num, wt, payload):
    return varint((num << 3) | wt) + (varint(len(payload)) + payload if wt == 2 else payload)


def entry_proto(dtype, shape, shard, offset, size, crc):
    dims = b"".join(field(2, 2, field(1, 0, varint(d))) for d in shape)
    msg = field(1, 0, varint(dtype)) + field(2, 2, dims)
    if shard:
        msg += field(3, 0, varint(shard))
    if offset:
        msg += field(4, 0, varint(offset))
    msg += field(5, 0, varint(size)) + field(6, 5, struct.pack("<I", crc))
    return msg


def build_block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
        prev = k
    for r in restarts or [0]:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts) or 1)
    return bytes(out)


def write_checkpoint(prefix, variables, entries_per_block=7, num_shards=1):
    """variables: {name: ndarray}.  Writes <prefix>.index, <prefix>.data-0000S-of-0000N and the ``checkpoint`` state file."""
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    names = sorted(variables)
    shard_bytes = [bytearray() for _ in range(num_shards)]
    items = [(b"", field(1, 0, varint(num_shards)))]   # BundleHeaderProto{num_shards=1}; endianness/version default
    for i, n in enumerate(names):
        a = np.ascontiguousarray(variables[n])
        s = i % num_shards
        raw = a.tobytes()
        items.append((n.encode(), entry_proto(DTYPE_ENUM[a.dtype], a.shape, s, len(shard_bytes[s]), len(raw), masked_crc(raw[:64]))))
        shard_bytes[s] += raw
    for s in range(num_shards):
        with open(f"{prefix}.data-{s:05d}-of-{num_shards:05d}", "wb") as f:
            f.write(shard_bytes[s])

    table = bytearray()

    def emit(block):
        off = len(table)
        trailer = b"\x00"
        table.extend(block + trailer + struct.pack("<I", masked_crc(block + trailer)))
        return varint(off) + varint(len(block))

    index_items = []
    for i in range(0, len(items), entries_per_block):
        chunk = items[i:i + entries_per_block]
        index_items.append((chunk[-1][0] + b"\x00", emit(build_block(chunk, restart_interval=4))))
    meta = emit(build_block([]))
    index = emit(build_block(index_items, restart_interval=1))
    footer = meta + index
    table.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(table)
    with open(os.path.join(os.path.dirname(prefix), "checkpoint"), "w") as f:
        base = os.path.basename(prefix)
        f.write(f'model_checkpoint_path: "{base}"\nall_model_checkpoint_paths: "{base}"\n')
