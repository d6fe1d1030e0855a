"""CPU: the on-disk formats of the reference restated without TensorFlow: TFRecord framing + tf.train.Example
(utils/tfrecord_utils.py) and the tf.train.Saver V2 tensor bundle (utils/tf_checkpoint.py).  The hand-written protobuf
encoders / decoders are checked against the protobuf RUNTIME on descriptors built from the published .proto definitions,
the CRC against the published CRC-32C check value."""
import os
import struct

import numpy as np
import pytest

from unsupervised_anomaly_detection_brain_mri_b200.utils import tfrecord_utils as T


def _example_classes():
    """tensorflow/core/example/feature.proto + example.proto, rebuilt at run time for the protobuf runtime."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name='uad_test_example.proto', package='uadtest', syntax='proto3')
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields, nested=()):
        m = fd.message_type.add(name=name)
        for f in fields:
            m.field.add(**f)
        return m
    msg('BytesList', [dict(name='value', number=1, label=F.LABEL_REPEATED, type=F.TYPE_BYTES)])
    msg('FloatList', [dict(name='value', number=1, label=F.LABEL_REPEATED, type=F.TYPE_FLOAT)])
    msg('Int64List', [dict(name='value', number=1, label=F.LABEL_REPEATED, type=F.TYPE_INT64)])
    feat = msg('Feature', [dict(name='bytes_list', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.uadtest.BytesList', oneof_index=0),
                           dict(name='float_list', number=2, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.uadtest.FloatList', oneof_index=0),
                           dict(name='int64_list', number=3, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.uadtest.Int64List', oneof_index=0)])
    feat.oneof_decl.add(name='kind')
    feats = msg('Features', [dict(name='feature', number=1, label=F.LABEL_REPEATED, type=F.TYPE_MESSAGE,
                                  type_name='.uadtest.Features.FeatureEntry')])
    entry = feats.nested_type.add(name='FeatureEntry')
    entry.options.map_entry = True
    entry.field.add(name='key', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_STRING)
    entry.field.add(name='value', number=2, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.uadtest.Feature')
    msg('Example', [dict(name='features', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.uadtest.Features')])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName('uadtest.Example'))


def test_crc32c_check_values():
    assert T.crc32c(b'123456789') == 0xE3069283                       # the published CRC-32C check value
    assert T.crc32c(b'') == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA                          # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43                 # RFC 3720 B.4: 32 bytes of ones
    assert T.crc32c(bytes(range(32))) == 0x46DD794E                   # RFC 3720 B.4: incrementing bytes
    rng = np.random.default_rng(0)
    for n in (65535, 65536, 70001, 300007):                          # lane-parallel path == serial path
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(d) == (T._crc_state_serial(memoryview(d), 0xFFFFFFFF) ^ 0xFFFFFFFF)
    # TFRecord's mask (record_writer.cc): rotate right by 15, add the constant
    c = T.crc32c(b'abc')
    assert T.masked_crc32c(b'abc') == ((((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF)


def test_example_encoding_matches_protobuf_runtime():
    Example = _example_classes()
    img = np.arange(12, dtype=np.float32).reshape(3, 4).tobytes()
    mine = T.encode_example({'height': T._int64_feature(3), 'width': T._int64_feature(4), 'image': T._bytes_feature(img),
                             'neg': T._int64_feature(-7)})
    ex = Example()
    ex.ParseFromString(mine)                                          # the runtime parses the hand-written bytes ...
    f = ex.features.feature
    assert f['height'].int64_list.value[0] == 3 and f['width'].int64_list.value[0] == 4 and f['neg'].int64_list.value[0] == -7
    assert f['image'].bytes_list.value[0] == img
    assert ex.SerializeToString(deterministic=True) == mine           # ... and re-serialises them to the same bytes
    # and the hand-written decoder parses what the runtime writes (any field order, floats included)
    ex2 = Example()
    ex2.features.feature['set'].bytes_list.value.append(b'\\x01\\x00\\x00\\x00')
    ex2.features.feature['w'].float_list.value.extend([1.5, -2.0])
    ex2.features.feature['n'].int64_list.value.extend([5, -1, 1 << 40])
    got = T.decode_example(ex2.SerializeToString())
    assert got['set'] == ('bytes', [b'\\x01\\x00\\x00\\x00']) and got['w'] == ('float', [1.5, -2.0])
    assert got['n'] == ('int64', [5, -1, 1 << 40])


def test_tfrecord_roundtrip_and_framing(tmp_path):
    rng = np.random.default_rng(1)
    images = rng.random((5, 16, 12, 1), dtype=np.float32)
    labels = (rng.random((5, 16, 12, 1)) > 0.9).astype(np.float32)
    sets = np.array([[0], [0], [1], [2], [1]], dtype=np.int32)        # TRAIN / VAL / TEST ids (dataloaders/BRAINWEB.py)
    fn = str(tmp_path / 'cache.tfrecord')
    T.write_tf_record(images, labels, sets, fn)
    im2, lb2, st2 = T.read_tf_record(fn)
    assert im2.shape == images.shape and im2.dtype == np.float32 and np.array_equal(im2, images)
    assert np.array_equal(lb2, labels) and np.array_equal(st2, sets)
    raw = open(fn, 'rb').read()
    (length,) = struct.unpack('<Q', raw[:8])
    assert struct.unpack('<I', raw[8:12])[0] == T.masked_crc32c(raw[:8])
    rec = raw[12:12 + length]
    assert struct.unpack('<I', raw[12 + length:16 + length])[0] == T.masked_crc32c(rec)
    Example = _example_classes()
    ex = Example()
    ex.ParseFromString(rec)
    assert sorted(ex.features.feature) == ['height', 'image', 'label', 'set', 'width']
    assert ex.features.feature['height'].int64_list.value[0] == 16
    # corruption is detected
    bad = bytearray(raw)
    bad[40] ^= 0x10
    open(fn, 'wb').write(bytes(bad))
    with pytest.raises(IOError):
        T.read_tf_record(fn)
