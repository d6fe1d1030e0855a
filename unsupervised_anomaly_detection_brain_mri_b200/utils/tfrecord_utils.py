"""TFRecord slice cache without TensorFlow (mirror of reference utils/tfrecord_utils.py: same function names, same files).

The reference caches the preprocessed slices of a dataset as one TFRecord of tf.train.Example messages
(dataloaders/BRAINWEB.py:69,197).  Both formats are small and public, so they are restated here byte for byte:

  TFRecord framing (tensorflow/core/lib/io/record_writer.cc):
      uint64 length (LE) | uint32 masked_crc32c(length bytes) | data | uint32 masked_crc32c(data)
      masked_crc = ((crc >> 15) | (crc << 17)) + 0xa282ead8   (mod 2^32),  crc = CRC-32C (Castagnoli)
  tf.train.Example (tensorflow/core/example/{example,feature}.proto), protobuf wire format:
      Example  { Features features = 1; }
      Features { map<string, Feature> feature = 1; }       -> repeated entry { string key = 1; Feature value = 2; }
      Feature  { oneof kind { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } }
      BytesList { repeated bytes value = 1; }   Int64List { repeated int64 value = 1 [packed = true]; }

`tests/test_formats.py` checks the hand-written encoder / decoder against the protobuf runtime (descriptors built from the
.proto definitions above) and the CRC against the published CRC-32C check value.
"""
import struct

import numpy

# ---------------------------------------------------------------------------------------------- CRC-32C
_CRC_TABLE = []


def _crc_table():
    if not _CRC_TABLE:
        poly = 0x82F63B78                       # reflected Castagnoli polynomial
        for n in range(256):
            c = n
            for _ in range(8):
                c = (c >> 1) ^ poly if c & 1 else c >> 1
            _CRC_TABLE.append(c)
    return _CRC_TABLE


def _crc_state_serial(mv, state):
    tab = _crc_table()
    for b in mv.tolist():
        state = tab[(state ^ b) & 0xFF] ^ (state >> 8)
    return state


def _zero_advance_matrix(nbytes):
    """Columns of the GF(2) operator 'shift the CRC register through nbytes zero bytes' (the register update is linear)."""
    tab = _crc_table()
    cols = []
    for bit in range(32):
        st = 1 << bit
        for _ in range(nbytes):
            st = tab[st & 0xFF] ^ (st >> 8)
        cols.append(st)
    return cols


def crc32c(data):
    """CRC-32C (Castagnoli) of ``data`` (bytes-like); check value crc32c(b'123456789') == 0xE3069283.
    Large buffers are cut into equal lanes whose registers advance together as one numpy vector (one table gather per
    byte position), then the lane results are folded with the zero-advance operator: crc(A||B) = Z_|B|(crc(A)) ^ crc_0(B)."""
    mv = memoryview(data).cast('B')
    n = len(mv)
    state = 0xFFFFFFFF
    if n >= (1 << 16):
        lanes = 4096
        L = n // lanes
        arr = numpy.frombuffer(mv[:lanes * L], dtype=numpy.uint8).reshape(lanes, L)
        tab = numpy.asarray(_crc_table(), dtype=numpy.uint32)
        st = numpy.zeros(lanes, dtype=numpy.uint32)
        st[0] = 0xFFFFFFFF
        for j in range(L):
            st = tab[(st ^ arr[:, j]) & 0xFF] ^ (st >> 8)
        cols = _zero_advance_matrix(L)
        state = int(st[0])
        for k in range(1, lanes):
            acc, bit, x = 0, 0, state
            while x:
                if x & 1:
                    acc ^= cols[bit]
                x >>= 1
                bit += 1
            state = acc ^ int(st[k])
        mv = mv[lanes * L:]
    return _crc_state_serial(mv, state) ^ 0xFFFFFFFF


def masked_crc32c(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- protobuf wire helpers
def _varint(n):
    n &= (1 << 64) - 1                           # int64 two's complement, as protobuf encodes negative values
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _len_delimited(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _fields(buf):
    """Yields (field number, wire type, value) of one serialized message; value is int or a bytes slice."""
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f'unsupported protobuf wire type {wt}')
        yield field, wt, v


def _bytes_feature(value):
    """Feature{bytes_list{value:[value]}} - same helper name as the reference (tfrecord_utils.py:6-7)."""
    return _len_delimited(1, _len_delimited(1, bytes(value)))


def _int64_feature(value):
    """Feature{int64_list{value:[value]}} (packed) - reference tfrecord_utils.py:10-11."""
    return _len_delimited(3, _len_delimited(1, _varint(int(value))))


def encode_example(features):
    """features: dict name -> serialized Feature.  Map entries are written in sorted key order (protobuf's deterministic
    order for string keys); readers accept any order."""
    body = b''.join(_len_delimited(1, _len_delimited(1, k.encode()) + _len_delimited(2, features[k])) for k in sorted(features))
    return _len_delimited(1, body)


def decode_example(record):
    """-> dict name -> ('bytes', [bytes...]) | ('int64', [int...]) | ('float', [float...])"""
    out = {}
    for f, _, feats in _fields(record):
        if f != 1:
            continue
        for f2, _, entry in _fields(feats):
            if f2 != 1:
                continue
            key, feature = None, b''
            for f3, _, v in _fields(entry):
                if f3 == 1:
                    key = v.decode()
                elif f3 == 2:
                    feature = v
            for kind, _, lst in _fields(feature):
                if kind == 1:
                    out[key] = ('bytes', [v for f4, _, v in _fields(lst) if f4 == 1])
                elif kind == 3:
                    vals = []
                    for f4, wt, v in _fields(lst):
                        if f4 != 1:
                            continue
                        if wt == 2:                           # packed
                            p = 0
                            while p < len(v):
                                x, p = _read_varint(v, p)
                                vals.append(x - (1 << 64) if x >> 63 else x)
                        else:
                            vals.append(v - (1 << 64) if v >> 63 else v)
                    out[key] = ('int64', vals)
                elif kind == 2:
                    vals = []
                    for f4, wt, v in _fields(lst):
                        if f4 == 1:
                            vals.extend(struct.unpack(f'<{len(v) // 4}f', v) if wt == 2 else struct.unpack('<f', v))
                    out[key] = ('float', vals)
    return out


# ---------------------------------------------------------------------------------------------- TFRecord framing
class TFRecordWriter:
    def __init__(self, filename):
        self._f = open(filename, 'wb')

    def write(self, record):
        header = struct.pack('<Q', len(record))
        self._f.write(header + struct.pack('<I', masked_crc32c(header)))
        self._f.write(record)
        self._f.write(struct.pack('<I', masked_crc32c(record)))

    def close(self):
        self._f.close()


def tf_record_iterator(path, verify=True):
    with open(path, 'rb') as f:
        while True:
            header = f.read(12)
            if not header:
                return
            if len(header) < 12:
                raise IOError(f'{path}: truncated TFRecord header')
            (length,), (hcrc,) = struct.unpack('<Q', header[:8]), struct.unpack('<I', header[8:])
            if verify and masked_crc32c(header[:8]) != hcrc:
                raise IOError(f'{path}: corrupted record length')
            data = f.read(length)
            footer = f.read(4)
            if len(data) < length or len(footer) < 4:
                raise IOError(f'{path}: truncated TFRecord')
            if verify and masked_crc32c(data) != struct.unpack('<I', footer)[0]:
                raise IOError(f'{path}: corrupted record data')
            yield data


# ---------------------------------------------------------------------------------------------- the reference's two functions
def write_tf_record(images, labels, sets, filename):
    """reference utils/tfrecord_utils.py:14-34 (features height, width, image, label, set per slice)."""
    writer = TFRecordWriter(filename)
    for i in range(0, images.shape[0]):
        img, label, set_ = images[i], labels[i], sets[i]
        example = encode_example({
            'height': _int64_feature(img.shape[0]),
            'width': _int64_feature(img.shape[1]),
            'image': _bytes_feature(numpy.ascontiguousarray(img).tobytes()),
            'label': _bytes_feature(numpy.ascontiguousarray(label).tobytes()),
            'set': _bytes_feature(numpy.ascontiguousarray(set_).tobytes())})
        writer.write(example)
    writer.close()


def read_tf_record(filename, verify=True):
    """reference utils/tfrecord_utils.py:37-52: float32 images / labels reshaped to [height, width, -1], int32 sets."""
    images, labels, sets = [], [], []
    for record in tf_record_iterator(filename, verify=verify):
        ex = decode_example(record)
        height, width = int(ex['height'][1][0]), int(ex['width'][1][0])
        images.append(numpy.frombuffer(ex['image'][1][0], dtype=numpy.float32).reshape(height, width, -1))
        labels.append(numpy.frombuffer(ex['label'][1][0], dtype=numpy.float32).reshape(height, width, -1))
        sets.append(numpy.frombuffer(ex['set'][1][0], dtype=numpy.int32))
    return numpy.array(images), numpy.array(labels), numpy.array(sets)
