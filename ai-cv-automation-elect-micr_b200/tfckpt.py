"""TensorFlow V2 ("tensor bundle") checkpoint reader without TensorFlow, and the mapping from the reference's TF-1
variable names to the parameter dict of ``weights.py``.

Replaces ``tf.train.latest_checkpoint(checkpoint_loc)`` + ``tf.train.Saver().restore(sess, ...)``
(machine_learning/denoiser.py:626-627; the trainer saves with ``tf.train.Saver().save``, misc_py/denoiser-multi-gpu.py:1149, 1218).

File formats (TensorFlow's tensor_bundle, public and stable since TF 0.12):
  ``<prefix>.index``                an SSTable in LevelDB's table format, uncompressed: key "" -> BundleHeaderProto,
                                    key <variable name> -> BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}
  ``<prefix>.data-SSSSS-of-NNNNN``  raw little-endian tensor bytes at [offset, offset+size) of shard shard_id
  ``checkpoint``                    text proto, ``model_checkpoint_path: "<prefix>"`` names the latest one

NOT VERIFIED AGAINST A REAL CHECKPOINT: no TensorFlow and no checkpoint of the reference exist in this environment; the
reader is tested against files produced by the format writer in tests/ (same published layout), and the variable naming
below is TF-1's documented auto-naming applied to the creation order of the graph -- ``strict=False``
reports what did not match instead of raising.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

from . import weights as _weights

_TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16}  # tensorflow DataType enum


# ---- protobuf wire format (varints, length-delimited) ------------------------------------------------------------
def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Yield (field number, wire type, value) of one protobuf message."""
    pos = 0
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield num, wt, val


def _parse_entry(buf):
    """BundleEntryProto -> (dtype enum, shape tuple, shard_id, offset, size)."""
    dtype, shape, shard, offset, size = 0, [], 0, 0, 0
    for num, wt, val in _fields(buf):
        if num == 1:
            dtype = val
        elif num == 2:  # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
            for n2, _, dimbuf in _fields(val):
                if n2 == 2:
                    d = 0
                    for n3, _, v3 in _fields(dimbuf):
                        if n3 == 1:
                            d = v3
                    shape.append(d)
        elif num == 3:
            shard = val
        elif num == 4:
            offset = val
        elif num == 5:
            size = val
    return dtype, tuple(shape), shard, offset, size


# ---- LevelDB table (SSTable) ---------------------------------------------------------------------------------------
def _block_handle(buf, pos):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def _read_block(data, off, size):
    block = data[off:off + size]
    ctype = data[off + size]  # 1-byte compression type, then a 4-byte masked crc32c
    if ctype != 0:
        raise ValueError("compressed table block (TensorFlow writes checkpoint indices uncompressed)")
    return block


def _block_entries(block):
    """Prefix-compressed key/value entries of one table block."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(path):
    """``<prefix>.index`` -> {variable name: (dtype enum, shape, shard_id, offset, size)}; the header entry (key "") is skipped."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != _TABLE_MAGIC:
        raise ValueError(f"{path} is not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _, _, pos = _block_handle(footer, 0)            # metaindex handle
    ioff, isize, _ = _block_handle(footer, pos)     # index handle
    entries = {}
    for _, handle in _block_entries(_read_block(data, ioff, isize)):
        boff, bsize, _ = _block_handle(handle, 0)
        for key, val in _block_entries(_read_block(data, boff, bsize)):
            if key:
                entries[key.decode()] = _parse_entry(val)
    return entries


def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: the prefix named by ``model_checkpoint_path`` in ``<dir>/checkpoint``, or None."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)


def read_checkpoint(prefix):
    """{variable name: ndarray} of every tensor in the checkpoint ``prefix`` (as returned by latest_checkpoint)."""
    index = read_index(prefix + ".index")
    n_shards = 1 + max((e[2] for e in index.values()), default=0)
    shards = {}
    out = {}
    for name, (dtype, shape, shard, offset, size) in index.items():
        if dtype not in _DTYPES:
            continue  # strings etc.: nothing the denoiser needs
        if shard not in shards:
            # the shard count in the file name is the checkpoint's, which may exceed the shards actually referenced
            cands = [f for f in os.listdir(os.path.dirname(prefix) or ".")
                     if f.startswith(os.path.basename(prefix) + f".data-{shard:05d}-of-")]
            if not cands:
                raise FileNotFoundError(f"{prefix}.data-{shard:05d}-of-{n_shards:05d}")
            shards[shard] = np.memmap(os.path.join(os.path.dirname(prefix) or ".", cands[0]), dtype=np.uint8, mode="r")
        raw = np.asarray(shards[shard][offset:offset + size])
        out[name] = raw.view(_DTYPES[dtype]).reshape(shape).copy()
    return out


# ---- TF-1 variable names of the reference graph -------------------------------------------------------------------
def tf_variable_names(variant="A", scope="nn"):
    """[(tf name, parameter name)] in creation order.  TF-1 auto-naming: the k-th use of a default scope gets the suffix
    ``_k`` (k >= 1).  Separable blocks: slim scope ``SeparableConv2d`` holding depthwise_weights / pointwise_weights and the
    normalizer's ``BatchNorm``; every block is followed by a stand-alone ``BatchNorm`` (batch_then_activ, DMG:220-223).
    Variant A dense convs are ``tf.layers`` (``conv2d[_k]/kernel|bias``, ``conv2d_transpose[_k]``) with the ASPP layers named
    ``1x1 lowRate mediumRate highRate imageLevel pellet`` (DMG:303-358); variant B uses slim (``Conv[_k]/weights|biases``,
    ``Conv2d_transpose[_k]``, DEN:91-145)."""
    counters = {}

    def uniq(base):
        k = counters.get(base, 0)
        counters[base] = k + 1
        return base if k == 0 else f"{base}_{k}"

    named = {"aspp_1x1": "1x1", "aspp_r6": "lowRate", "aspp_r12": "mediumRate", "aspp_r18": "highRate",
             "aspp_image": "imageLevel", "aspp_pellet": "pellet"}
    bn_vars = (("beta", "beta"), ("gamma", "gamma"), ("moving_mean", "mean"), ("moving_variance", "var"))
    pairs = []

    def bn(tf_scope, pname):
        for tfv, ours in bn_vars:
            pairs.append((f"{scope}/{tf_scope}/{tfv}", f"{pname}/{ours}"))

    for name, kind, cin, cout, k in _weights.layer_table(variant):
        if kind == "sep":
            s = uniq("SeparableConv2d")
            pairs.append((f"{scope}/{s}/depthwise_weights", f"{name}/dw"))
            pairs.append((f"{scope}/{s}/pointwise_weights", f"{name}/pw"))
            bn(f"{s}/BatchNorm", f"{name}/bn1")
            bn(uniq("BatchNorm"), f"{name}/bn2")
        elif kind == "bn":
            bn(uniq("BatchNorm"), f"{name}/bn")
        else:
            if variant == "A":
                s = named.get(name) if kind == "conv" and name in named else uniq("conv2d" if kind == "conv" else "conv2d_transpose")
                kern, bias = "kernel", "bias"
            else:
                s = uniq("Conv" if kind == "conv" else "Conv2d_transpose")
                kern, bias = "weights", "biases"
            pairs.append((f"{scope}/{s}/{kern}", f"{name}/{'kernel' if kind == 'conv' else 'tkernel'}"))
            pairs.append((f"{scope}/{s}/{bias}", f"{name}/bias"))
            bn(uniq("BatchNorm"), f"{name}/bn")
    return pairs


def params_from_variables(variables, variant="A", scope="nn", strict=True):
    """TF variables {name: ndarray} -> the parameter dict ``weights.pack`` takes.  Shapes are checked against the graph; with
    ``strict=False`` missing / mismatching variables are returned as a second value instead of raising."""
    shapes = {}
    for name, kind, cin, cout, k in _weights.layer_table(variant):
        if kind == "sep":
            shapes[f"{name}/dw"], shapes[f"{name}/pw"] = (3, 3, cin, 1), (1, 1, cin, cout)
        elif kind == "conv":
            shapes[f"{name}/kernel"], shapes[f"{name}/bias"] = (k, k, cin, cout), (cout,)
        elif kind == "deconv":
            shapes[f"{name}/tkernel"], shapes[f"{name}/bias"] = (3, 3, cout, cin), (cout,)
    params, problems = {}, []
    for tf_name, pname in tf_variable_names(variant, scope):
        if tf_name not in variables:
            problems.append(f"missing {tf_name} (for {pname})")
            continue
        v = np.asarray(variables[tf_name], np.float32)
        want = shapes.get(pname)
        if want is not None and tuple(v.shape) != want:
            problems.append(f"{tf_name}: shape {tuple(v.shape)}, expected {want} (for {pname})")
            continue
        params[pname] = v
    if problems and strict:
        raise ValueError("checkpoint does not match the graph: " + "; ".join(problems[:8]) + (" ..." if len(problems) > 8 else ""))
    return (params, problems) if not strict else params


def _prefix_of(checkpoint_loc):
    prefix = latest_checkpoint(checkpoint_loc) if os.path.isdir(checkpoint_loc) else checkpoint_loc
    if prefix is None or not os.path.exists(prefix + ".index"):
        raise FileNotFoundError(f"no TensorFlow checkpoint at {checkpoint_loc}")
    return prefix


def load_params(checkpoint_loc, variant="A", scope="nn", strict=True):
    """Denoiser(checkpoint_loc=<directory or prefix>) : latest checkpoint of a directory, or an explicit prefix.
    ``strict=False`` returns (params, problems) instead of raising on missing / mis-shaped variables."""
    return params_from_variables(read_checkpoint(_prefix_of(checkpoint_loc)), variant, scope, strict)


def diagnose(checkpoint_loc, scope="nn"):
    """For the first real checkpoint someone tries (the variable naming here is TF-1's documented scope uniquifier applied to
    the reference's creation order, unverified against a real file -- DESIGN.md): how the checkpoint's variables line up with
    each graph variant.  Returns {variant: {"problems": [...], "unused": [checkpoint variables the graph did not ask for]}}."""
    variables = read_checkpoint(_prefix_of(checkpoint_loc))
    report = {}
    for variant in ("A", "B"):
        _, problems = params_from_variables(variables, variant, scope, strict=False)
        wanted = {t for t, _ in tf_variable_names(variant, scope)}
        unused = sorted(v for v in variables if v not in wanted and "Momentum" not in v and v != "global_step")
        report[variant] = {"problems": problems, "unused": unused}
    return report
