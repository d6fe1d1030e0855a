"""CPU: host-side logic of the drop-in boundary (no GPU): variable naming / flat layout, model protocol, dataset protocol,
the best-Dice recursion, checkpoint layout, config plumbing."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import scoring as oscore
from oracle import tf_graph_cpu as O
from unsupervised_anomaly_detection_brain_mri_b200 import engine as E
from unsupervised_anomaly_detection_brain_mri_b200.dataloaders.SYNTHETIC import SYNTHETIC
from unsupervised_anomaly_detection_brain_mri_b200.models import (autoencoder, context_encoder_variational_autoencoder,
                                                                  variational_autoencoder)
from unsupervised_anomaly_detection_brain_mri_b200.models.customlayers import Placeholder
from unsupervised_anomaly_detection_brain_mri_b200.trainers import CE, Metrics
from unsupervised_anomaly_detection_brain_mri_b200.utils import default_config_setup as dcs


@pytest.mark.parametrize('arch', O.ARCHS)
@pytest.mark.parametrize('S', [32, 128, 256])
def test_param_specs_match_oracle_variables(arch, S):
    specs = E.param_specs(arch, S)
    P = O.init_params(arch, S)
    assert list(specs.keys()) == list(P.keys())
    assert all(tuple(specs[k]) == P[k].shape for k in P)


def test_flat_params_roundtrip_and_alignment():
    specs = E.param_specs(O.VAE, 32)
    fp = E.FlatParams(specs, 'cpu')
    vals = O.perturb_params(O.init_params(O.VAE, 32))
    fp.load(vals)
    back = fp.to_numpy()
    assert all(np.array_equal(back[k], vals[k]) for k in vals)
    assert all(off % 64 == 0 for off in fp.offsets.values())         # 256-byte aligned slots
    assert fp.n_params == sum(v.size for v in vals.values())
    lo, hi = fp.subset_ranges('Encoder/')
    assert lo == 0 and hi == fp.offsets['Bottleneck/conv2d/kernel']


def test_glorot_init_statistics():
    v = E.glorot_init(E.param_specs(O.AE, 64))
    k = v['Encoder/enc_conv2D_1/kernel']
    lim = np.sqrt(6.0 / (25 * 32 + 25 * 64))
    assert np.abs(k).max() <= lim and np.abs(k).max() > 0.95 * lim
    assert np.all(v['Encoder/batch_normalization/gamma'] == 1) and np.all(v['Encoder/enc_conv2D_0/bias'] == 0)


def test_model_functions_keep_reference_protocol():
    class C:
        zDim, intermediateResolutions, outputWidth, numChannels = 128, [8, 8], 128, 1
    x = Placeholder([None, 128, 128, 1])
    assert set(autoencoder.autoencoder(x, 0.2, False, C)) == {'z', 'x_hat'}
    assert set(variational_autoencoder.variational_autoencoder(x, 0.2, False, C)) == {'z_mu', 'z_log_sigma', 'z_sigma', 'x_hat'}
    out = context_encoder_variational_autoencoder.context_encoder_variational_autoencoder(x, x, 0.2, False, C)
    assert {'x_hat', 'x_hat_ce', 'z_mu_ce'} <= set(out)
    g = out['x_hat'].graph
    assert (g.arch, g.S, g.res) == ('context_encoder_variational_autoencoder', 128, 8)
    assert autoencoder.autoencoder.__name__ == 'autoencoder'          # used for directory names (AEMODEL.py:32-33)
    assert [l['filters'] for l in g.encoder if l['op'] == 'conv2d'] == [32, 64, 128, 128]
    assert [l['filters'] for l in g.decoder if l['op'] == 'conv2d_transpose'] == [128, 64, 32, 32]


def test_synthetic_dataset_protocol():
    o = SYNTHETIC.Options()
    o.sliceResolution, o.numPatients, o.sliceStart, o.sliceEnd = [32, 32], 2, 20, 40
    ds = SYNTHETIC(o)
    assert ds.num_channels == 1
    assert ds.num_batches(8, set='TRAIN') == (28 // 8)
    b, l, m = ds.next_batch(8, set='TRAIN', return_brainmask=True)
    assert b.shape == (8, 32, 32, 1) and b.dtype == np.float32 and m.shape == b.shape
    assert 0.3 < (b == 0).mean() < 0.9 and b.max() <= 1.0
    for _ in range(10):                                               # wrap-around keeps serving full batches
        assert ds.next_batch(8, set='TRAIN')[0].shape[0] == 8
    vol, seg, skull = ds.load_volume_and_groundtruth(ds.patients[0]['filtered_files'])
    assert vol.num_slices_along_axis('axial') == 40 and vol.get_slice(25, 'axial').shape == (32, 32)


def test_ce_mask_quirk_last_sample_mask_is_broadcast():
    import random
    random.seed(0)
    batch = np.ones((3, 64, 64, 1), np.float32)
    bm = np.zeros((3, 64, 64, 1), np.uint8)
    bm[:, 8:56, 8:56] = 1
    out = CE.retrieve_masked_batch(batch, bm)
    assert out.shape == batch.shape and out.dtype == np.float32
    assert (out == 0).any()
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[1], out[2])     # one mask for the whole batch (CE.py:130-138)


class NumpyScorer:
    """Same interface as Metrics.DeviceScorer, computed with numpy (stands in for the GPU on the CPU box)."""

    def __init__(self, pred, lab):
        self.p, self.l = pred, lab

    def dice_scores(self, ths):
        return [Metrics.dice(np.where(self.p > t, 1, 0), self.l) for t in ths]


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_best_dice_recursion_matches_reference_semantics(seed):
    rng = np.random.default_rng(seed)
    lab = (rng.uniform(size=(4, 32, 32)) < 0.1).astype(int)
    pred = np.clip(0.25 * rng.standard_normal(lab.shape) + 0.35 * lab, 0, None).astype(np.float32).astype(np.float64)
    best, thr, ths, scs = oscore.best_dice_search(pred, lab, granularity=5)
    scores, threshs = Metrics.compute_dice_score(pred, lab, 5, scorer=NumpyScorer(pred, lab))
    assert list(threshs) == list(ths) and list(scores) == list(scs)
    b2, t2 = Metrics.compute_dice_curve_recursive(pred, lab, granularity=5, scorer=NumpyScorer(pred, lab))
    assert (b2, t2) == (best, thr)


def test_options_and_config_plumbing():
    opts = dcs.get_options(batchsize=8, learningrate=1e-4, numEpochs=2, zDim=128, outputWidth=128, outputHeight=128,
                           config={'CHECKPOINTDIR': '/tmp/c', 'SAMPLEDIR': '/tmp/s', 'BRAINWEBDIR': ''})
    assert opts['threshold'] == 'bestdice' and opts['keepOnlyPositiveResiduals'] and opts['sliceStart'] == 20
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.DLMODEL import DLMODEL

    class T:
        class Config(DLMODEL.Config):
            pass

    class DS:
        num_channels = 1
    cfg = dcs.get_config(T, opts, 'ADAM', [8, 8], 0.2, DS())
    assert (cfg.beta1, cfg.batchsize, cfg.dropout_rate, cfg.optimizer, cfg.dataset) == (0.5, 8, 0.2, 'ADAM', 'DS')
    json.dumps(cfg.__dict__)
    assert dcs.Dataset.Brainweb is dcs.Dataset.BRAINWEB      # run.py's spelling is accepted


def test_checkpoint_layout_roundtrip(tmp_path):
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.DLMODEL import DLMODEL

    class Eng:
        def __init__(self):
            self.specs = E.param_specs(O.AE, 32)
            self.fp = E.FlatParams(self.specs, 'cpu')
            self.fp.load(O.perturb_params(O.init_params(O.AE, 32)))
    m = DLMODEL(None)
    m.config.modelname, m.engine = 'AE', Eng()
    m.save(str(tmp_path), 3)
    d = os.path.join(str(tmp_path), m.model_dir)
    assert {'AE.model-3.npz', 'Config-3.json', 'Curves.npy', 'checkpoint'} <= set(os.listdir(d))
    want = m.engine.fp.to_numpy()
    m.engine.fp.params.zero_()
    ok, step = m.load(str(tmp_path))
    assert ok and step == 3
    got = m.engine.fp.to_numpy()
    assert all(np.array_equal(got[k], want[k]) for k in want)
    assert DLMODEL(None).load(str(tmp_path / 'nothing')) == (False, 0)


def test_early_stopping_rule():
    from unsupervised_anomaly_detection_brain_mri_b200.trainers.AEMODEL import indicate_early_stopping
    best, last = float('inf'), 0
    stops = []
    for c in [5, 4, 4.5, 4.6, 4.7, 4.8, 4.9]:
        best, last, stop = indicate_early_stopping(c, best, last)
        stops.append(stop)
    assert stops == [False, False, False, False, False, False, True]


def test_sibling_models_and_trainers_keep_reference_protocol():
    """models/autoencoder_spatial.py, models/constrained_autoencoder.py, trainers/ConstrainedAE.py, trainers/VAE_You.py:
    names, output keys and Config defaults of the reference (looked up BY NAME from run.py:21-24)."""
    import importlib
    PKG = 'unsupervised_anomaly_detection_brain_mri_b200'

    class C:
        zDim, intermediateResolutions, outputWidth, numChannels = 128, [8, 8], 128, 1
    x = Placeholder([None, 128, 128, 1])
    for mname, keys in (('autoencoder_spatial', {'z', 'x_hat'}), ('constrained_autoencoder', {'z', 'x_hat', 'z_rec'})):
        fn = getattr(importlib.import_module(f'{PKG}.models.{mname}'), mname)
        assert fn.__name__ == mname
        out = fn(x, 0.2, False, C)
        assert set(out) == keys and out['x_hat'].graph.arch == mname
    cae = getattr(importlib.import_module(f'{PKG}.trainers.ConstrainedAE'), 'ConstrainedAE')
    assert cae.Config().rho == 1 and cae.Config().modelname == 'ConstrainedAE'          # ConstrainedAE.py:13-16
    you = getattr(importlib.import_module(f'{PKG}.trainers.VAE_You'), 'VAE_You')
    c = you.Config()
    assert (c.modelname, c.restore_lr, c.restore_steps, c.tv_lambda) == ('VAE_You', 1e-3, 150, 1.8)   # VAE_You.py:13-18
    # the constrained AE shares the dense AE's variables (the re-encoding pass reuses the layers)
    assert list(E.param_specs(E.CAE, 128)) == list(E.param_specs(E.AE, 128))
    assert not any(k.startswith('Bottleneck/') for k in E.param_specs(E.AES, 128))
    for m in ('main_AE_spatial.py', 'main_constrainedAE.py', 'main_VAE_You.py'):
        assert os.path.isfile(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'mains', m))


def test_oracle_constrained_and_spatial_graphs_consistent():
    """The oracle's sibling graphs reduce to compositions of its own pieces (no dropout): z_rec == Enc(x_hat), and the
    spatial AE equals decoder(encoder(x))."""
    P = O.perturb_params(O.init_params(O.CAE, 32, seed=4))
    x = O.synthetic_slices(2, 32, seed=9)
    out = O.forward(O.CAE, P, x, dtype=torch.float64)
    again = O.forward(O.CAE, P, out['x_hat'].numpy(), dtype=torch.float64)
    assert torch.allclose(out['z_rec'], again['z'], rtol=1e-12, atol=1e-12)
    L = O.losses(O.CAE, out, x, dtype=torch.float64, rho=0.5)
    l2 = ((out['x_hat'].numpy() - x) ** 2).mean(axis=(1, 2, 3))
    rz = ((out['z'].numpy() - out['z_rec'].numpy()) ** 2).mean(axis=1)
    assert np.isclose(float(L['loss']), float((l2 + 0.5 * rz).mean()))
    Ps = O.perturb_params(O.init_params(O.AES, 32, seed=4))
    outs = O.forward(O.AES, Ps, x, dtype=torch.float64)
    Pt = {k: torch.from_numpy(v).double() for k, v in Ps.items()}
    ref = O.decoder(Pt, O.encoder(Pt, torch.from_numpy(x).double().permute(0, 3, 1, 2))).permute(0, 2, 3, 1)
    assert torch.allclose(outs['x_hat'], ref)
