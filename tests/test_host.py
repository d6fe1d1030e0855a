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


def test_lesion_wise_detection_rate_and_summary():
    """utils/Evaluation.compute_detection_rate / summarize_predictions (reference Evaluation.py:130-172, 463-500; skimage label /
    regionprops restated on scipy.ndimage): hand-built volumes with known lesion counts."""
    from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation as Ev
    gt = np.zeros((25, 16, 16), bool)
    pred = np.zeros_like(gt)
    gt[2:5, 2:5, 2:5] = True             # lesion A: detected
    pred[3:6, 3:6, 3:6] = True
    gt[10:12, 8:11, 8:11] = True         # lesion B: missed
    pred[8:10, 1:4, 12:15] = True        # 18-voxel false positive
    pred[14, 14, 14] = True              # 1-voxel false positive: dropped (< 8 voxels)
    gt[21:24, 4:8, 4:8] = True           # lesion C (second block of 20 slices): detected
    pred[22:24, 5:9, 5:9] = True
    pred[0, 0, 0] = pred[1, 1, 1] = True  # 26-connected with nothing else of size: a 2-voxel component, dropped
    assert Ev.compute_detection_rate(pred, gt) == (2, 1, 1)
    # a lesion straddling the block boundary is seen once per block, as in the reference's per-block labelling
    gt2 = np.zeros((25, 8, 8), bool)
    gt2[18:22, 2:5, 2:5] = True
    assert Ev.compute_detection_rate(gt2.copy(), gt2) == (2, 0, 0)
    assert Ev.compute_detection_rate(np.zeros_like(gt2), gt2) == (0, 0, 2)
    # summary over two "patients" stacked along the slice axis
    preds, gts = np.concatenate([pred, np.zeros_like(pred)]), np.concatenate([gt, gt])
    ev = Ev.summarize_predictions({}, gts.astype(np.uint8), preds, preds, 2, 25, 'bestdice')
    assert ev['thresholdType'] == 'bestdice' and (ev['TPCC'], ev['FPCC'], ev['FNCC']) == (2, 1, 1 + 3)
    assert ev['TPRCC'] == 2 / 6 and ev['PrecisionCC'] == 2 / 3
    tp = int((preds & gts).sum())
    assert (ev['TP'], ev['FP'], ev['FN'], ev['TN']) == (tp, int(preds.sum()) - tp, int(gts.sum()) - tp, gts.size - int((preds | gts).sum()))
    assert ev['DiceScore'] == pytest.approx(2 * tp / (preds.sum() + gts.sum()))
    assert ev['DiceScorePerPatient'][0] == pytest.approx(2 * tp / (pred.sum() + gt.sum())) and ev['DiceScorePerPatient'][1] == 0
    assert ev['RecallPerPatient'] == [pytest.approx(tp / gt.sum()), 0.0] and np.isnan(ev['PrecisionPerPatient'][1])
    assert ev['TPR'] == ev['FPR'] == pytest.approx(tp / gts.sum())            # (sic) the reference's FPR is Metrics.tpr
    assert ev['VD'] == pytest.approx((gts.sum() - tp) / gts.sum())
    assert ev['DiceScorePerPatientMean'] == pytest.approx(np.mean(ev['DiceScorePerPatient']))


def test_evaluate_orchestration_keys_and_files(tmp_path, monkeypatch):
    """utils/Evaluation.evaluate with the device pieces replaced by numpy stand-ins (a scorer that answers `diffs > t` counts, a
    canned _evaluate): the result carries the reference's evalPC keys (Evaluation.py:440-526), the files land in the reference's
    eval-<epoch>-<timestamp>-<description> directory, and the numbers agree with a direct numpy evaluation."""
    import types

    from unsupervised_anomaly_detection_brain_mri_b200.utils import Evaluation as Ev
    rng = np.random.default_rng(0)
    n_pat, Z, S = 2, 20, 24
    labels = np.zeros((n_pat * Z, S, S), np.uint8)
    labels[3:7, 5:10, 5:10] = 1
    labels[Z + 8:Z + 12, 12:18, 10:15] = 1
    diffs = (0.02 * rng.random((n_pat * Z, S, S))).astype(np.float32)
    diffs[labels > 0] += (0.05 + 0.1 * rng.random(int(labels.sum()))).astype(np.float32)
    diffs[Z + 2:Z + 4, 2:5, 2:5] = 0.2                                   # a false-positive blob
    diffs64 = diffs.astype(np.float64)

    class NumpyScorer:
        def __init__(self, predictions, lab, device=None, allreduce=None):
            self.d, self.l = np.asarray(predictions, np.float64).reshape(-1), np.asarray(lab).reshape(-1) != 0

        def dice_scores(self, thresholds):
            out = []
            for t in thresholds:
                p = self.d > t
                with np.errstate(divide='ignore', invalid='ignore'):
                    out.append(np.float64(2 * np.sum(p & self.l)) / np.float64(p.sum() + self.l.sum()))
            return out

        def threshold_mask(self, t):
            import torch
            return torch.from_numpy((self.d > t).astype(np.uint8))

    def fake_evaluate(datasetObj, modelObj, sampleDir, options, split='TEST', shard=None):
        ev = Ev.get_eval_dictionary()
        ev.update(x=diffs64 * 0, reconstructions=diffs64 * 0, diffs=diffs64, labelmaps=labels, l1reconstructionErrorMean=1.0,
                  reconstructionTimes=0.001)
        return ev, [{'name': 'p0'}, {'name': 'p1'}]

    monkeypatch.setattr(Ev, '_evaluate', fake_evaluate)
    monkeypatch.setattr(Ev.Metrics, 'DeviceScorer', NumpyScorer)
    model = types.SimpleNamespace(network=types.SimpleNamespace(__name__='variational_autoencoder'), model_dir='VAE_test', device='cpu', world=1)
    options = {'train': {'samplesDir': str(tmp_path)}, 'threshold': 'bestdice', 'exportROC': True, 'exportPRC': True,
               'sliceStart': 20, 'sliceEnd': 40}
    ev = Ev.evaluate(None, model, options, epoch='3', description='SYNTHETIC-bestdice')
    ref_keys = {'diff_AUC', 'diff_AUPRC', 'bestDiceScore', 'bestThreshold', 'thresholdType', 'DiceScore', 'DiceScorePerPatient',
                'PrecisionPerPatient', 'RecallPerPatient', 'DiceScorePerPatientMean', 'DiceScorePerPatientStd', 'PrecisionPerPatientMean',
                'PrecisionPerPatientStd', 'RecallPerPatientMean', 'RecallPerPatientStd', 'TP', 'FP', 'TN', 'FN', 'TPR', 'FPR', 'VD', 'TPCC',
                'FPCC', 'FNCC', 'TPRCC', 'PrecisionCC'}
    assert ref_keys <= set(ev), ref_keys - set(ev)
    d = ev['evalDir']
    assert os.path.basename(d).startswith('eval-3-') and d.endswith('-SYNTHETIC-bestdice')
    assert os.path.dirname(d) == os.path.join(str(tmp_path), 'variational_autoencoder', 'VAE_test')
    for f in ('rocPC.npy', 'prcPC.npy', 'evalPC.npy', 'evalPC.txt'):
        assert os.path.isfile(os.path.join(d, f)), f
    assert os.path.isdir(os.path.join(d, 'samples_test_PC'))
    stored = np.load(os.path.join(d, 'evalPC.npy'), allow_pickle=True).item()
    assert ref_keys <= set(stored) and 'diffs' not in stored and 'x' not in stored
    roc = np.load(os.path.join(d, 'rocPC.npy'), allow_pickle=True).item()
    assert set(roc) == {'fpr', 'tpr', 'threshs'}
    # numbers: the best-Dice threshold separates lesions (> 0.05) from background (< 0.02) and the blob (0.2) is a false positive
    assert 0.02 <= ev['bestThreshold'] < 0.05 and ev['thresholdType'] == 'bestdice'
    mask = Ev.filter_3d_connected_components((diffs64 > ev['bestThreshold']).copy())
    tp = int((mask & (labels > 0)).sum())
    assert ev['TP'] == tp and ev['FN'] == int(labels.sum()) - tp and ev['FP'] == 18
    assert ev['DiceScore'] == pytest.approx(2 * tp / (mask.sum() + labels.sum())) == ev['DICE']
    assert (ev['TPCC'], ev['FNCC']) == (2, 0) and ev['TPRCC'] == 1.0
    assert ev['diff_AUC'] > 0.99 and ev['AUC'] == ev['diff_AUC']
    # a fixed threshold: thresholdType records it and the lesion-wise rate is taken at that operating point
    options['threshold'] = 0.1
    ev2 = Ev.evaluate(None, model, options, epoch='3')
    assert ev2['thresholdType'] == 0.1 and ev2['threshold'] == 0.1 and not ev2['evalDir'].endswith('bestdice')
    assert ev2['FPCC'] == 1


def test_run_py_keeps_the_reference_command_line():
    """Flag set of reference run.py:119-152 (short + long spellings); --numPatients is the one addition (synthetic data size)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('uad_run', os.path.join(root, 'run.py'))
    run = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run)
    have = {tuple(a.option_strings) for a in run.build_parser()._actions if a.dest != 'help'}
    want = {('-c', '--config'), ('-b', '--batchsize'), ('-l', '--lr'), ('-E', '--numEpochs'), ('-z', '--zDim'), ('-w', '--outputWidth'),
            ('-g', '--outputHeight'), ('-o', '--optimizer'), ('-i', '--intermediateResolutions'), ('-s', '--slices_start'),
            ('-e', '--slices_end'), ('-t', '--trainer'), ('-m', '--model'), ('-O', '--threshold'), ('-d', '--ds'),
            ('-n', '--numMonteCarloSamples'), ('-G', '--use_gradient_based_restoration'), ('-K', '--kappa'), ('-M', '--scale'),
            ('-R', '--rho'), ('-C', '--dim_c'), ('-Z', '--dim_z'), ('-W', '--dim_w'), ('-A', '--c_lambda'), ('-L', '--restore_lr'),
            ('-S', '--restore_steps'), ('-T', '--tv_lambda')}
    assert have == want | {('--numPatients',)}
    args = run.build_parser().parse_args(['-t', 'VAE', '-m', 'variational_autoencoder', '-d', 'MSLUB', '-i', '16', '16'])
    assert args.trainer == 'VAE' and args.ds.name == 'MSLUB' and args.intermediateResolutions == [16, 16] and args.lr == 1e-4


def test_every_main_script_binds_to_existing_product_symbols():
    """mains/main_*.py (mirrors of the reference's mains/, which are scripts that train on import): parse, do not run - every
    `from <package>... import name` must resolve, none may touch the oracle, and the reference's hot-path mains all exist."""
    import ast
    import importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = 'unsupervised_anomaly_detection_brain_mri_b200'
    have = sorted(f for f in os.listdir(os.path.join(root, 'mains')) if f.endswith('.py'))
    for want in ('main_AE.py', 'main_AE_spatial.py', 'main_VAE.py', 'main_VAE_You.py', 'main_ceVAE.py', 'main_CE.py', 'main_constrainedAE.py',
                 'main_AAE.py', 'main_constrainedAAE.py', 'main_GMVAE.py', 'main_GMVAE_spatial.py', 'main_fAnoGAN.py'):
        assert want in have
    for f in have:
        tree = ast.parse(open(os.path.join(root, 'mains', f)).read(), f)
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom):
                assert not node.module.startswith('oracle'), f
                if node.module.startswith(pkg):
                    mod = importlib.import_module(node.module)
                    for alias in node.names:
                        assert hasattr(mod, alias.name), (f, node.module, alias.name)
            elif isinstance(node, ast.Import):
                assert not any(a.name.startswith(('oracle', 'tensorflow')) for a in node.names), f


def test_combined_predictive_uncertainty_formula():
    """trainers/Metrics.py:170-173: E[p^2] - E[p]^2 + E[sigma^2] over the Monte-Carlo axis; log_var=True exponentiates sigma first."""
    from unsupervised_anomaly_detection_brain_mri_b200.trainers import Metrics
    rng = np.random.default_rng(0)
    p = rng.uniform(size=(5, 3, 8, 8)).astype(np.float32)
    sig = rng.uniform(size=p.shape).astype(np.float32)
    got = Metrics.combined_predictive_uncertainty(p, sig, axis=0)
    ref = (p.astype(np.float64) ** 2).mean(0) - p.astype(np.float64).mean(0) ** 2 + sig.astype(np.float64).mean(0)
    assert got.shape == (3, 8, 8) and np.allclose(got, ref, atol=1e-6)
    got_lv = Metrics.combined_predictive_uncertainty(p, np.log(sig), axis=0, log_var=True)
    assert np.allclose(got_lv, ref, atol=1e-5)
    assert np.allclose(Metrics.combined_predictive_uncertainty(p, np.zeros_like(p), axis=0), p.var(axis=0), atol=1e-6)


def test_bench_reference_arm_plumbing():
    """bench.py --impl reference: c5 answers `unavailable` (the CPU arm times the VAE configs), and the argument table names the
    three configs BASELINE.json asks bench lines for."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--config', 'c5'], capture_output=True,
                         text=True, timeout=300)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert out.returncode == 0 and line['impl'] == 'reference' and 'unavailable' in line
    helptext = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--help'], capture_output=True, text=True, timeout=300).stdout
    assert all(k in helptext for k in ('c2', 'c4', 'c5', '--batch', 'tc1'))
