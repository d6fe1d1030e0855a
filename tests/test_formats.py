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
    ex2.features.feature['set'].bytes_list.value.append(b'\x01\x00\x00\x00')
    ex2.features.feature['w'].float_list.value.extend([1.5, -2.0])
    ex2.features.feature['n'].int64_list.value.extend([5, -1, 1 << 40])
    got = T.decode_example(ex2.SerializeToString())
    assert got['set'] == ('bytes', [b'\x01\x00\x00\x00']) and got['w'] == ('float', [1.5, -2.0])
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


# ------------------------------------------------------------------------------------------------ tensor bundle (Saver V2)
def _bundle_classes():
    """tensorflow/core/protobuf/tensor_bundle.proto (+ the pieces of tensor_shape.proto / versions.proto it uses)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name='uad_test_bundle.proto', package='uadb', syntax='proto3')
    F = descriptor_pb2.FieldDescriptorProto
    ver = fd.message_type.add(name='VersionDef')
    ver.field.add(name='producer', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_INT32)
    ver.field.add(name='min_consumer', number=2, label=F.LABEL_OPTIONAL, type=F.TYPE_INT32)
    shp = fd.message_type.add(name='TensorShapeProto')
    dim = shp.nested_type.add(name='Dim')
    dim.field.add(name='size', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_INT64)
    dim.field.add(name='name', number=2, label=F.LABEL_OPTIONAL, type=F.TYPE_STRING)
    shp.field.add(name='dim', number=2, label=F.LABEL_REPEATED, type=F.TYPE_MESSAGE, type_name='.uadb.TensorShapeProto.Dim')
    shp.field.add(name='unknown_rank', number=3, label=F.LABEL_OPTIONAL, type=F.TYPE_BOOL)
    hdr = fd.message_type.add(name='BundleHeaderProto')
    hdr.field.add(name='num_shards', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_INT32)
    hdr.field.add(name='endianness', number=2, label=F.LABEL_OPTIONAL, type=F.TYPE_INT32)       # enum LITTLE = 0, BIG = 1
    hdr.field.add(name='version', number=3, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.uadb.VersionDef')
    ent = fd.message_type.add(name='BundleEntryProto')
    ent.field.add(name='dtype', number=1, label=F.LABEL_OPTIONAL, type=F.TYPE_INT32)            # enum DataType
    ent.field.add(name='shape', number=2, label=F.LABEL_OPTIONAL, type=F.TYPE_MESSAGE, type_name='.uadb.TensorShapeProto')
    ent.field.add(name='shard_id', number=3, label=F.LABEL_OPTIONAL, type=F.TYPE_INT32)
    ent.field.add(name='offset', number=4, label=F.LABEL_OPTIONAL, type=F.TYPE_INT64)
    ent.field.add(name='size', number=5, label=F.LABEL_OPTIONAL, type=F.TYPE_INT64)
    ent.field.add(name='crc32c', number=6, label=F.LABEL_OPTIONAL, type=F.TYPE_FIXED32)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName('uadb.' + n))   # noqa: E731
    return get('BundleHeaderProto'), get('BundleEntryProto')


def test_bundle_protos_match_protobuf_runtime():
    from unsupervised_anomaly_detection_brain_mri_b200.utils import tf_checkpoint as C
    Header, Entry = _bundle_classes()
    h = Header()
    h.ParseFromString(C.encode_header(1))
    assert (h.num_shards, h.endianness, h.version.producer) == (1, 0, 1)
    assert h.SerializeToString(deterministic=True) == C.encode_header(1)
    raw = C.encode_entry(np.float32, (5, 5, 32, 64), offset=1234567, size=204800, crc_masked=0xDEADBEEF)
    e = Entry()
    e.ParseFromString(raw)
    assert (e.dtype, [d.size for d in e.shape.dim], e.shard_id, e.offset, e.size, e.crc32c) == (1, [5, 5, 32, 64], 0, 1234567, 204800, 0xDEADBEEF)
    assert e.SerializeToString(deterministic=True) == raw
    e2 = Entry(dtype=9, offset=0, size=8, crc32c=7)                  # scalar int64 (e.g. a global_step): no dims at all
    d = C.decode_entry(e2.SerializeToString())
    assert (d['dtype'], d['shape'], d['offset'], d['size'], d['crc32c']) == (9, [], 0, 8, 7)


def test_bundle_roundtrip_table_structure(tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200 import engine as E
    from unsupervised_anomaly_detection_brain_mri_b200.utils import tf_checkpoint as C
    specs = E.param_specs(E.VAE, 256)
    weights = E.glorot_init(specs, seed=3)
    rng = np.random.default_rng(0)
    m = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in weights.items()}
    v = {k: rng.random(v.shape).astype(np.float32) for k, v in weights.items()}
    variables = C.saver_variables(weights, m, v, step=10, beta1=0.5, beta2=0.999)
    assert 'Encoder/batch_normalization/moving_mean' in variables and 'Decoder/dec_Conv2DT_4/kernel/Adam_1' in variables
    prefix = str(tmp_path / 'VAE_dSYNTH' / 'VAE.model-3')
    C.write_bundle(prefix, variables)
    C.update_checkpoint_state(os.path.dirname(prefix), 'VAE.model-3')
    assert C.latest_checkpoint(os.path.dirname(prefix)) == prefix
    header, entries = C.list_bundle(prefix)
    assert header == {'num_shards': 1, 'endianness': 0, 'producer': 1}
    assert list(entries) == sorted(variables, key=lambda n: n.encode())              # bytewise key order
    back = C.read_bundle(prefix)
    assert all(np.array_equal(back[k], variables[k]) and back[k].shape == np.shape(variables[k]) for k in variables)
    assert back['beta1_power'].shape == () and np.isclose(back['beta1_power'], 0.5 ** 11)
    w2, m2, v2, frozen = C.split_saver_variables(back, list(weights))
    assert frozen and all(np.array_equal(w2[k], weights[k]) and np.array_equal(m2[k], m[k]) for k in weights)
    # structure of the index file: magic, several data blocks (> 4 KiB of entries), checksummed blocks, prefix compression
    raw = open(prefix + '.index', 'rb').read()
    assert struct.unpack('<Q', raw[-8:])[0] == 0xDB4775248B80FB57
    items = C.parse_table(raw)
    assert items[0][0] == b'' and len(items) == len(variables) + 1
    assert len(raw) < sum(len(k) + len(val) for k, val in items) + 2048                # shared key prefixes were elided
    footer = raw[-48:-8]
    pos = 0
    for _ in range(2):
        _, pos = C._read_varint(footer, pos)
    ioff, pos = C._read_varint(footer, pos)
    isize, pos = C._read_varint(footer, pos)
    assert len(C._read_block(raw, ioff, isize)) >= 2                                   # more than one data block
    # data file = tensors back to back in key order
    sizes = [entries[k]['size'] for k in entries]
    assert [entries[k]['offset'] for k in entries] == list(np.cumsum([0] + sizes[:-1]))
    assert os.path.getsize(prefix + '.data-00000-of-00001') == sum(sizes)
    # corruption of a tensor and of an index block is detected
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[100] ^= 1
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    with pytest.raises(IOError):
        C.read_bundle(prefix)
    bad = bytearray(raw)
    bad[10] ^= 1
    with pytest.raises(IOError):
        C.parse_table(bytes(bad))
    # a bundle with non-trivial moving statistics is flagged (the kernels assume the reference's frozen 0 / 1)
    back['Encoder/batch_normalization/moving_mean'] = np.full_like(back['Encoder/batch_normalization/moving_mean'], 0.1)
    assert C.split_saver_variables(back, list(weights))[3] is False


def test_trainer_tf_checkpoint_export_import_on_cpu(tmp_path):
    """DLMODEL.export_tf_checkpoint / import_tf_checkpoint with a CPU-resident parameter buffer (no kernels involved)."""
    import types

    import torch

    from unsupervised_anomaly_detection_brain_mri_b200 import engine as E
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.DLMODEL import DLMODEL
    from unsupervised_anomaly_detection_brain_mri_b200.utils import tf_checkpoint as C

    class T(DLMODEL):
        def train(self, dataset):
            pass

    def make(seed):
        cfg = DLMODEL.Config()
        cfg.modelname, cfg.dataset, cfg.description = 'VAE', 'SYNTHETIC', 'cpu-test'
        t = T(None, cfg)
        specs = E.param_specs(E.VAE, 64)
        fp = E.FlatParams(specs, 'cpu')
        fp.load(E.glorot_init(specs, seed))
        t.engine = types.SimpleNamespace(fp=fp, specs=specs, t=0, step_dev=torch.zeros(1, dtype=torch.int64), adam_step=lambda *a, **k: None)
        return t
    a, b = make(1), make(2)
    rng = np.random.default_rng(0)
    a.engine.fp.load({k: rng.standard_normal(s).astype(np.float32) for k, s in a.engine.specs.items()}, buf=a.engine.fp.m)
    a.engine.fp.load({k: rng.random(s).astype(np.float32) for k, s in a.engine.specs.items()}, buf=a.engine.fp.v)
    a.engine.t = 37
    # the reference's variable set: no optimiser slots
    prefix = a.export_tf_checkpoint(str(tmp_path), 4)
    assert os.path.basename(prefix) == 'VAE.model-4' and os.path.isfile(prefix + '.index')
    names = list(C.list_bundle(prefix)[1])
    # default export: tf.compat.v1.layers' per-scope numbering of the un-named layers (the Decoder's first BN has no suffix)
    assert not any(n.endswith('/Adam') for n in names) and 'Decoder/batch_normalization/moving_variance' in names
    assert 'Decoder/batch_normalization_3/gamma' in names and 'Decoder/batch_normalization_6/gamma' not in names   # 64^2: 3 + 4 BN layers
    assert 'Encoder/batch_normalization_2/beta' in names and 'Bottleneck/dense_2/kernel' in names and 'Encoder/enc_conv2D_2/kernel' in names
    g_names = list(C.list_bundle(a.export_tf_checkpoint(str(tmp_path / 'graph'), 4, suffix_scheme='graph'))[1])
    assert 'Decoder/batch_normalization_6/moving_variance' in g_names
    assert b.import_tf_checkpoint(os.path.dirname(prefix)) == 4                      # via the `checkpoint` state file
    wa, wb = a._weights(), b._weights()
    assert all(np.array_equal(wa[k], wb[k]) for k in wa) and b.engine.t == 0
    # with optimiser state
    prefix = a.export_tf_checkpoint(str(tmp_path), 5, with_optimizer=True)
    c = make(3)
    assert c.import_tf_checkpoint(prefix) == 5
    assert c.engine.t == 37 and int(c.engine.step_dev) == 37
    ma, mc = a.engine.fp.to_numpy(a.engine.fp.m), c.engine.fp.to_numpy(c.engine.fp.m)
    assert all(np.array_equal(ma[k], mc[k]) for k in ma)
    # the exported bundles share the `checkpoint` state file with the native .npz checkpoints: a native save at step 3 followed by
    # the exports above (steps 4, 5 - no .npz) must not make load() restart from scratch
    d = make(4)
    d._load_weights = lambda w: d.engine.fp.load(w)
    a.curves = {}
    a.save(str(tmp_path), 3)
    a.export_tf_checkpoint(str(tmp_path), 6)
    ok, counter = d.load(str(tmp_path))
    assert ok and counter == 3
    assert all(np.array_equal(wa[k], v) for k, v in d._weights().items())
    # a checkpoint of another architecture is refused with a clear error
    ae = make(1)
    ae.engine.specs = E.param_specs(E.CEVAE, 128)
    with pytest.raises((KeyError, ValueError)):
        ae.import_tf_checkpoint(prefix)


def test_tf_checkpoint_import_pairs_layers_whose_automatic_suffixes_differ(tmp_path):
    """A bundle whose un-named layers are numbered PER VARIABLE SCOPE (tf.layers makes the layer scope unique when the layer is
    first called: the Decoder's first BatchNormalization is 'Decoder/batch_normalization') instead of over the whole graph (the
    convention of engine.param_specs: 'Decoder/batch_normalization_3') imports into the same variables, slots and moving
    statistics included; a bundle with a different number of layers is refused."""
    import re
    import types

    import torch

    from unsupervised_anomaly_detection_brain_mri_b200 import engine as E
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.DLMODEL import DLMODEL
    from unsupervised_anomaly_detection_brain_mri_b200.utils import tf_checkpoint as C

    class T(DLMODEL):
        def train(self, dataset):
            pass

    specs = E.param_specs(E.VAE, 64)
    rng = np.random.default_rng(3)
    weights = {k: rng.standard_normal(s).astype(np.float32) for k, s in specs.items()}
    m = {k: rng.standard_normal(s).astype(np.float32) for k, s in specs.items()}
    v = {k: rng.random(s).astype(np.float32) for k, s in specs.items()}
    saved = C.saver_variables(weights, m, v, step=9)
    n_enc = sum(1 for k in specs if re.match(r'Encoder/batch_normalization(_\d+)?/gamma', k))

    def per_scope(name):                      # Decoder/batch_normalization_{n_enc + j} -> Decoder/batch_normalization[_j]
        mm = re.match(r'^Decoder/batch_normalization_(\d+)(/.*)$', name)
        if not mm:
            return name
        j = int(mm.group(1)) - n_enc
        return f'Decoder/batch_normalization{"" if j == 0 else "_" + str(j)}{mm.group(2)}'

    renamed = {per_scope(k): a for k, a in saved.items()}
    assert 'Decoder/batch_normalization/gamma' in renamed and f'Decoder/batch_normalization_{2 * n_enc}/gamma' not in renamed
    mapping = C.resolve_layer_names(list(specs), renamed.keys())
    assert mapping[f'Decoder/batch_normalization_{n_enc}'] == 'Decoder/batch_normalization'
    assert mapping['Bottleneck/dense_2'] == 'Bottleneck/dense_2' and mapping['Encoder/batch_normalization_1'] == 'Encoder/batch_normalization_1'
    prefix = str(tmp_path / 'VAE.model-7')
    C.write_bundle(prefix, renamed)
    cfg = DLMODEL.Config()
    cfg.modelname = 'VAE'
    t = T(None, cfg)
    fp = E.FlatParams(specs, 'cpu')
    t.engine = types.SimpleNamespace(fp=fp, specs=specs, t=0, step_dev=torch.zeros(1, dtype=torch.int64), adam_step=lambda *a, **k: None)
    assert t.import_tf_checkpoint(prefix) == 7
    got, gm = t._weights(), fp.to_numpy(fp.m)
    assert all(np.array_equal(got[k], weights[k]) for k in specs) and all(np.array_equal(gm[k], m[k]) for k in specs)
    assert t.engine.t == 9
    fewer = {k: a for k, a in renamed.items() if not k.startswith('Decoder/batch_normalization_2/')}
    C.write_bundle(str(tmp_path / 'VAE.model-8'), fewer)
    with pytest.raises(KeyError, match='batch_normalization'):
        t.import_tf_checkpoint(str(tmp_path / 'VAE.model-8'))
