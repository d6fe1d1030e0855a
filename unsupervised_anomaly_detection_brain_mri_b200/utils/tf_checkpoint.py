"""tf.train.Saver (V2, "tensor bundle") checkpoints without TensorFlow: read the reference's `<name>.model-<step>.index` +
`.data-00000-of-00001` files into {variable name: ndarray} and write the same format (reference trainers/DLMODEL.py:66-110
saves / restores through tf.train.Saver; the variable names and layouts are the ones this repo already uses).

Format restated from TensorFlow's sources (no TensorFlow-written file is available in this environment, so the reader is
validated against this writer, the protobuf runtime and the structural rules below - see tests/test_formats.py):

  <prefix>.data-00000-of-00001   raw little-endian tensor bytes, back to back, in key order
  <prefix>.index                 an SSTable (tensorflow/core/lib/io/table*.cc == LevelDB's table format):
      data blocks | metaindex block (empty) | index block | footer
      block   : entries [varint shared][varint non_shared][varint value_len][key suffix][value] ... ,
                restart offsets (uint32 LE each), restart count (uint32 LE); then a 5-byte trailer:
                compression type (0 = none; 1 = snappy is NOT supported here) + masked CRC-32C of block + type
      index   : one entry per data block: key >= last key of the block, value = BlockHandle (varint offset, varint size)
      footer  : metaindex BlockHandle, index BlockHandle, zero padding to 40 bytes, magic 0xdb4775248b80fb57 (LE)
    keys -> values (tensorflow/core/protobuf/tensor_bundle.proto):
      ""          -> BundleHeaderProto { int32 num_shards = 1; Endianness endianness = 2; VersionDef version = 3 {producer = 1} }
      <var name>  -> BundleEntryProto  { DataType dtype = 1; TensorShapeProto shape = 2 { repeated Dim dim = 2 { int64 size = 1 } };
                                         int32 shard_id = 3; int64 offset = 4; int64 size = 5; fixed32 crc32c = 6 (masked) }
  checkpoint                     text CheckpointState: model_checkpoint_path / all_model_checkpoint_paths
"""
import os
import re
import struct
from collections import OrderedDict

import numpy as np

from .tfrecord_utils import _fields, _len_delimited, _read_varint, _varint, crc32c

MAGIC = 0xDB4775248B80FB57
BLOCK_SIZE = 4096           # table::Options::block_size
RESTART_INTERVAL = 16       # table::Options::block_restart_interval
# tensorflow/core/framework/types.proto
DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
      17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DT_OF = {np.dtype(v): k for k, v in DT.items()}


def _mask(crc):
    return (((crc >> 15) | (crc << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- protobuf messages
def _tag_varint(field, value):
    return _varint(field << 3) + _varint(value)


def encode_header(num_shards=1):
    return _tag_varint(1, num_shards) + _len_delimited(3, _tag_varint(1, 1))      # endianness LITTLE = 0 is the default: omitted


def encode_entry(dtype, shape, offset, size, crc_masked, shard_id=0):
    dims = b''.join(_len_delimited(2, _tag_varint(1, int(d))) for d in shape)
    out = _tag_varint(1, DT_OF[np.dtype(dtype)]) + _len_delimited(2, dims)
    if shard_id:
        out += _tag_varint(3, shard_id)
    if offset:
        out += _tag_varint(4, offset)
    if size:
        out += _tag_varint(5, size)
    return out + _varint((6 << 3) | 5) + struct.pack('<I', crc_masked)


def decode_entry(buf):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None, 'slices': 0}
    for f, wt, v in _fields(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:
            for f2, _, dim in _fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, x in _fields(dim):
                        if f3 == 1:
                            size = x
                    e['shape'].append(size)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = struct.unpack('<I', v)[0]
        elif f == 7:
            e['slices'] += 1
    return e


def decode_header(buf):
    h = {'num_shards': 0, 'endianness': 0, 'producer': 0}
    for f, _, v in _fields(buf):
        if f == 1:
            h['num_shards'] = v
        elif f == 2:
            h['endianness'] = v
        elif f == 3:
            for f2, _, x in _fields(v):
                if f2 == 1:
                    h['producer'] = x
    return h


# ---------------------------------------------------------------------------------------------- SSTable
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.ri, self.buf, self.restarts, self.count, self.last = restart_interval, bytearray(), [0], 0, b''

    def add(self, key, value):
        shared = 0
        if self.count < self.ri:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def _handle(offset, size):
    return _varint(offset) + _varint(size)


def build_table(items):
    """items: list of (key bytes, value bytes) in strictly increasing key order -> SSTable bytes."""
    out = bytearray()
    index = _BlockBuilder(1)

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes + b'\x00' + struct.pack('<I', _mask(crc32c(block_bytes + b'\x00'))))
        return off, len(block_bytes)
    blk, last_key, prev = _BlockBuilder(RESTART_INTERVAL), None, None
    for key, value in items:
        assert prev is None or key > prev, 'keys must be strictly increasing'
        prev = key
        blk.add(key, value)
        last_key = key
        if blk.size() >= BLOCK_SIZE:
            off, size = emit(blk.finish())
            index.add(last_key, _handle(off, size))
            blk, last_key = _BlockBuilder(RESTART_INTERVAL), None
    if last_key is not None:
        off, size = emit(blk.finish())
        index.add(last_key, _handle(off, size))
    moff, msize = emit(_BlockBuilder(RESTART_INTERVAL).finish())
    ioff, isize = emit(index.finish())
    footer = _handle(moff, msize) + _handle(ioff, isize)
    out.extend(footer + bytes(40 - len(footer)) + struct.pack('<Q', MAGIC))
    return bytes(out)


def _read_block(buf, off, size, verify=True):
    block, ctype = buf[off:off + size], buf[off + size]
    if verify:
        want = struct.unpack('<I', buf[off + size + 1:off + size + 5])[0]
        if _mask(crc32c(buf[off:off + size + 1])) != want:
            raise IOError('tensor bundle index: block checksum mismatch')
    if ctype != 0:
        raise IOError(f'tensor bundle index: compressed block (type {ctype}) - snappy is not supported')
    n_restarts = struct.unpack('<I', block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key, items = 0, b'', []
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        items.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return items


def parse_table(buf, verify=True):
    if len(buf) < 48 or struct.unpack('<Q', buf[-8:])[0] != MAGIC:
        raise IOError('not a TensorFlow checkpoint index (bad table magic)')
    footer = buf[-48:-8]
    pos = 0
    _, pos = _read_varint(footer, pos)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    items = []
    for _, handle in _read_block(buf, ioff, isize, verify):
        off, p = _read_varint(handle, 0)
        size, _ = _read_varint(handle, p)
        items.extend(_read_block(buf, off, size, verify))
    return items


# ---------------------------------------------------------------------------------------------- bundles
def write_bundle(prefix, tensors):
    """tensors: {name: ndarray}.  Writes <prefix>.index and <prefix>.data-00000-of-00001 (one shard)."""
    names = sorted(tensors, key=lambda n: n.encode())
    items, offset = [(b'', encode_header(1))], 0
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    with open(prefix + '.data-00000-of-00001', 'wb') as f:
        for name in names:
            assert name, 'the empty key is reserved for the bundle header'
            a = np.asarray(tensors[name])
            if not a.flags.c_contiguous:
                a = np.ascontiguousarray(a)                  # (np.ascontiguousarray would turn a scalar into shape [1])
            if a.dtype.byteorder == '>':
                a = a.astype(a.dtype.newbyteorder('<'))
            raw = a.tobytes()
            f.write(raw)
            items.append((name.encode(), encode_entry(a.dtype, a.shape, offset, len(raw), _mask(crc32c(raw)))))
            offset += len(raw)
    with open(prefix + '.index', 'wb') as f:
        f.write(build_table(items))


def list_bundle(prefix, verify=True):
    """-> (header dict, OrderedDict name -> entry dict) of <prefix>.index"""
    items = parse_table(open(prefix + '.index', 'rb').read(), verify)
    if not items or items[0][0] != b'':
        raise IOError('tensor bundle index: missing header entry')
    header = decode_header(items[0][1])
    if header['endianness'] != 0:
        raise IOError('big-endian bundles are not supported')
    return header, OrderedDict((k.decode(), decode_entry(v)) for k, v in items[1:])


def read_bundle(prefix, names=None, verify=True):
    """-> OrderedDict name -> ndarray (all variables, or ``names``).  Checks every tensor's CRC-32C."""
    header, entries = list_bundle(prefix, verify)
    n = max(header['num_shards'], 1)
    shards = {}
    out = OrderedDict()
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e['slices']:
            raise IOError(f'{name}: partitioned (sliced) variables are not supported')
        if e['dtype'] not in DT:
            raise IOError(f'{name}: unsupported dtype enum {e["dtype"]}')
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = open(f'{prefix}.data-{sid:05d}-of-{n:05d}', 'rb')
        f = shards[sid]
        f.seek(e['offset'])
        raw = f.read(e['size'])
        if len(raw) != e['size']:
            raise IOError(f'{name}: truncated data shard')
        if verify and e['crc32c'] is not None and _mask(crc32c(raw)) != e['crc32c']:
            raise IOError(f'{name}: tensor checksum mismatch')
        out[name] = np.frombuffer(raw, dtype=np.dtype(DT[e['dtype']]).newbyteorder('<')).reshape(e['shape']).copy()
    for f in shards.values():
        f.close()
    return out


# ---------------------------------------------------------------------------------------------- CheckpointState file
def latest_checkpoint(checkpoint_dir):
    """tf.train.latest_checkpoint: the prefix named by <dir>/checkpoint, or None."""
    state = os.path.join(checkpoint_dir, 'checkpoint')
    if not os.path.isfile(state):
        return None
    m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', open(state).read())
    if not m:
        return None
    p = m.group(1)
    return p if os.path.isabs(p) else os.path.join(checkpoint_dir, p)


def update_checkpoint_state(checkpoint_dir, name, keep=5):
    state = os.path.join(checkpoint_dir, 'checkpoint')
    old = re.findall(r'all_model_checkpoint_paths:\s*"([^"]+)"', open(state).read()) if os.path.isfile(state) else []
    paths = [p for p in old if p != name][-(keep - 1):] + [name]
    with open(state, 'w') as f:
        f.write(f'model_checkpoint_path: "{name}"\n')
        for p in paths:
            f.write(f'all_model_checkpoint_paths: "{p}"\n')


# ---------------------------------------------------------------------------------------------- Saver-shaped variable sets
def saver_variables(weights, adam_m=None, adam_v=None, step=0, beta1=0.5, beta2=0.999):
    """The variable set tf.train.Saver() writes for the reference's graphs (every global variable): the trainable variables,
    the BatchNormalization moving statistics (never updated in the reference: the layers are called without training=True,
    so they stay at their initial 0 / 1 - SURVEY App. A.3), and, when given, the Adam slots `<var>/Adam`, `<var>/Adam_1`
    plus `beta1_power` / `beta2_power` (= beta^(step+1), tf.train.AdamOptimizer)."""
    out = OrderedDict()
    for name, w in weights.items():
        out[name] = np.asarray(w, np.float32)
        if name.endswith('/gamma') and 'batch_normalization' in name:
            base = name[:-len('/gamma')]
            out[base + '/moving_mean'] = np.zeros(w.shape, np.float32)
            out[base + '/moving_variance'] = np.ones(w.shape, np.float32)
    if adam_m is not None and adam_v is not None:
        for name in weights:
            out[name + '/Adam'] = np.asarray(adam_m[name], np.float32)
            out[name + '/Adam_1'] = np.asarray(adam_v[name], np.float32)
        out['beta1_power'] = np.float32(beta1 ** (step + 1))
        out['beta2_power'] = np.float32(beta2 ** (step + 1))
    return out


def split_saver_variables(variables, wanted):
    """Inverse of ``saver_variables`` for the names in ``wanted``: -> (weights, adam_m or None, adam_v or None, frozen_bn_ok).
    frozen_bn_ok is False when a checkpoint carries moving statistics other than 0 / 1 (the frozen-BN kernels assume them)."""
    weights = OrderedDict((n, variables[n]) for n in wanted)
    has_slots = all((n + '/Adam') in variables and (n + '/Adam_1') in variables for n in wanted)
    m = OrderedDict((n, variables[n + '/Adam']) for n in wanted) if has_slots else None
    v = OrderedDict((n, variables[n + '/Adam_1']) for n in wanted) if has_slots else None
    ok = True
    for n, a in variables.items():
        if n.endswith('/moving_mean') and np.any(a != 0):
            ok = False
        if n.endswith('/moving_variance') and np.any(a != 1):
            ok = False
    return weights, m, v, ok


# ---------------------------------------------------------------------------------------------- auto-suffix tolerant names
_SUFFIX = None


def _split_layer(prefix):
    """'Decoder/batch_normalization_7' -> ('Decoder', 'batch_normalization', 7); names without a numeric suffix -> index 0 ...
    unless the whole last component is an explicit layer name such as 'enc_conv2D_3' (kept verbatim, index None)."""
    import re
    global _SUFFIX
    if _SUFFIX is None:
        _SUFFIX = re.compile(r'^(batch_normalization|layer_normalization|dense|conv2d|conv2d_transpose)(?:_(\d+))?$')
    scope, _, layer = prefix.rpartition('/')
    m = _SUFFIX.match(layer)
    if not m:
        return scope, layer, None
    return scope, m.group(1), int(m.group(2) or 0)


def resolve_layer_names(wanted, available):
    """Map the layer prefixes this code uses onto the ones a checkpoint holds when only the AUTOMATIC numeric suffixes differ.

    TensorFlow numbers un-named layers automatically ('dense', 'dense_1', ...; 'batch_normalization_4', ...), and whether the
    counter runs per variable scope (tf.layers: the scope is made unique when the layer is first called) or over the whole
    graph (Keras object names) depends on the layer class and TF version - SURVEY App. A.10 could not confirm the suffixes
    without a TensorFlow runtime.  What IS determined by the model code is, per (scope, layer kind), the ORDER in which the
    layers are created; so within each (scope, kind) the wanted layers, sorted by suffix, are paired with the available ones,
    sorted by suffix.  wanted / available: iterables of variable names.  Returns {wanted_prefix: available_prefix}; raises
    KeyError when a (scope, kind) group has a different number of layers on the two sides."""
    def groups(names):
        g = OrderedDict()
        for n in names:
            prefix = n.rpartition('/')[0]
            scope, kind, idx = _split_layer(prefix)
            if idx is None:
                continue
            g.setdefault((scope, kind), {})[idx] = prefix
        return g
    leaves = ('/kernel', '/bias', '/gamma', '/beta')
    gw = groups(n for n in wanted if n.endswith(leaves))
    ga = groups(n for n in available if n.endswith(leaves))
    mapping = {}
    for key, layers in gw.items():
        have = ga.get(key, {})
        if len(have) != len(layers):
            raise KeyError(f'checkpoint holds {len(have)} {key[1]} layers in scope {key[0]!r}, the graph needs {len(layers)}')
        for w_idx, a_idx in zip(sorted(layers), sorted(have)):
            mapping[layers[w_idx]] = have[a_idx]
    return mapping


def per_scope_layer_names(names):
    """{layer prefix -> the prefix tf.compat.v1.layers gives it}: the legacy layer classes the reference imports
    (models/customlayers.py:4) open `variable_scope(None, default_name=<base name>)` when first called, so un-named layers are
    numbered PER ENCLOSING VARIABLE SCOPE in creation order ('Decoder/batch_normalization', 'Decoder/batch_normalization_1', ...),
    whereas engine.param_specs numbers them over the whole graph ('Decoder/batch_normalization_6', ...).  Explicitly named layers
    ('enc_conv2D_3') and layers whose kind appears once per scope keep their names."""
    order = OrderedDict()
    for n in names:
        prefix = n.rpartition('/')[0]
        scope, kind, idx = _split_layer(prefix)
        if idx is not None:
            order.setdefault((scope, kind), {})[idx] = prefix
    mapping = {}
    for (scope, kind), layers in order.items():
        for k, idx in enumerate(sorted(layers)):
            mapping[layers[idx]] = (scope + '/' if scope else '') + (kind if k == 0 else f'{kind}_{k}')
    return mapping


def rename_prefixes(variables, mapping):
    """Variables re-keyed by {old layer prefix -> new layer prefix}; slot variables and moving statistics follow their layer."""
    out = OrderedDict()
    for name, value in variables.items():
        hit = None
        for a in mapping:
            if name.startswith(a + '/') and (hit is None or len(a) > len(hit)):
                hit = a
        out[name if hit is None else mapping[hit] + name[len(hit):]] = value
    return out


def rename_layers(variables, mapping):
    """Checkpoint variables re-keyed to this code's layer prefixes (inverse of ``mapping``'s direction); slot variables and
    moving statistics follow their layer."""
    inverse = {a: w for w, a in mapping.items()}
    out = OrderedDict()
    for name, value in variables.items():
        hit = None
        for a in inverse:
            if name.startswith(a + '/') and (hit is None or len(a) > len(hit)):
                hit = a
        out[name if hit is None else inverse[hit] + name[len(hit):]] = value
    return out
