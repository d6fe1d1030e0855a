"""Residual-map anomaly scoring and evaluation (mirror of the hot part of reference utils/Evaluation.py).

Device side (hand-written kernels through the C ABI):
  * the reconstruction of a WHOLE sub-volume in one batched forward pass (the reference calls ``reconstruct`` once
    per slice with batch 1: Evaluation.py:246-250),
  * the residual / brain-mask / hyper-intensity-prior arithmetic (:282-291) -> ``uad_residual_score``,
  * ``diffs > t`` + Dice counts for every candidate threshold (:444-457, Metrics.py:138-162) -> ``uad_threshold_counts``.
Host side (scipy / sklearn like the reference; skimage's label / regionprops restated on scipy.ndimage): connected-component
filter (:113-127), lesion-wise detection rate (:130-172), ROC / PRC and their .npy exports, the per-patient Dice / precision /
recall, confusion counts and the evalPC.npy / evalPC.txt summary with the reference's keys (:440-526).  PNG exports and
matplotlib plots are not reproduced."""
import math
import os
import time

import numpy as np
import scipy.ndimage
import torch

from .. import abi
from ..trainers import Metrics


def should(options, key):
    return key in options and options[key]


def get_eval_dictionary():
    return {'x': [], 'reconstructions': [], 'diffs': [], 'labelmaps': [], 'l1reconstructionErrors': [],
            'l2reconstructionErrors': [], 'reconstructionTimes': [], 'epistemic_variance': []}


def erode_brainmask(brainmask):
    strel = scipy.ndimage.generate_binary_structure(2, 1)
    return scipy.ndimage.binary_erosion(np.squeeze(brainmask), structure=strel, iterations=12)


def apply_brainmask(x, brainmask, erode=True):
    if erode:
        brainmask = erode_brainmask(brainmask)
    return np.multiply(np.squeeze(brainmask), np.squeeze(x))


def apply_3d_median_filter(volume, kernelsize=5):
    return scipy.ndimage.median_filter(volume, (kernelsize, kernelsize, kernelsize))


def filter_3d_connected_components(volume):
    """Remove 26-connected components with filled area <= 7 (Evaluation.py:113-127; skimage.label(connectivity=3))."""
    sz = None
    if volume.ndim > 3:
        sz = volume.shape
        volume = np.reshape(volume, [sz[0] * sz[1], sz[2], sz[3]])
    cc, n = scipy.ndimage.label(volume, structure=np.ones((3, 3, 3)))
    if n:
        for idx, sl in enumerate(scipy.ndimage.find_objects(cc), start=1):
            if sl is None:
                continue
            comp = cc[sl] == idx
            if comp.sum() <= 7 and scipy.ndimage.binary_fill_holes(comp).sum() <= 7:
                volume[sl][comp] = 0
    if sz is not None:
        volume = np.reshape(volume, sz)
    return volume


def is_float(s):
    try:
        float(s)
        return True
    except (TypeError, ValueError):
        return False


def export_patient_volume(nii_seg, subvolume, zoom_factor, dataset_options, options, sample_dir, patient_name):
    """options['exportVolumes'] (Evaluation.py:323-334): the residual sub-volume, resampled back to the native slice resolution,
    written into a copy of the patient's volume geometry as <name>.nii.gz (+ <name>.binary.nii.gz for a numeric threshold).
    Needs a volume object with the NII interface (utils/NII.py); returns the paths written."""
    dezoom = (1,) + tuple(1 / np.asarray(zoom_factor, np.float64)) if zoom_factor is not None else (1, 1, 1)
    restored = scipy.ndimage.zoom(subvolume, dezoom)
    nii_seg.set_to_zero()
    nii_seg.cast_to_float()
    end = min(dataset_options.sliceEnd, dataset_options.sliceStart + restored.shape[0])
    nii_seg.set_subvolume(dataset_options.sliceStart, end, restored, axis=dataset_options.axis)
    paths = [os.path.join(sample_dir, '{}.nii.gz'.format(patient_name))]
    nii_seg.save(paths[0])
    if options.get('threshold') and is_float(options['threshold']):
        nii_seg.data = np.asarray(nii_seg.data > float(options['threshold'])).astype(np.float32)
        paths.append(os.path.join(sample_dir, '{}.binary.nii.gz'.format(patient_name)))
        nii_seg.save(paths[1])
    return paths


def compute_detection_rate(predicted_volume, groundtruth_volume):
    """Lesion-wise true / false positives and false negatives (Evaluation.py:130-172), in blocks of 20 slices: connected
    components (skimage.measure.label's default = full connectivity) of prediction AND ground truth are the true positives;
    predicted components of fewer than 8 voxels are dropped; every component touched by a true positive is removed from the
    prediction / ground truth, what remains are the false positives / false negatives."""
    full = np.ones((3, 3, 3))
    tps = fns = fps = 0
    pred = np.asarray(predicted_volume).astype(bool)
    gt = np.asarray(groundtruth_volume).astype(bool)
    num_slices = gt.shape[0]
    inter = pred & gt
    for s in range(int(math.ceil(num_slices / 20))):
        sl = slice(s * 20, min((s + 1) * 20, num_slices))
        cc_i, n_i = scipy.ndimage.label(inter[sl], structure=full)
        cc_p, n_p = scipy.ndimage.label(pred[sl], structure=full)
        cc_g, _ = scipy.ndimage.label(gt[sl], structure=full)
        if n_p:
            areas = np.bincount(cc_p.ravel(), minlength=n_p + 1)
            small = np.flatnonzero(areas < 8)
            cc_p[np.isin(cc_p, small[small > 0])] = 0
        for lab in range(1, n_i + 1):
            first = np.argwhere(cc_i == lab)[0]                       # regionprops' coords[0]: first voxel in row-major order
            lp = cc_p[tuple(first)]
            cc_p[cc_p == lp] = 0                                      # (lp == 0, a dropped small component, clears nothing new)
            lg = cc_g[tuple(first)]
            cc_g[cc_g == lg] = 0
        tps += n_i
        fns += len(np.unique(cc_g)) - (1 if (cc_g == 0).any() else 0)
        fps += len(np.unique(cc_p)) - (1 if (cc_p == 0).any() else 0)
    return tps, fps, fns


def summarize_predictions(eval_pc, labelmaps, diffs_thresholded, diffs_thresholded_at_precision70, num_patients, num_slices,
                          threshold_type):
    """The per-patient / lesion-wise / confusion statistics of the reference's evaluate() (Evaluation.py:463-500), same keys."""
    labels = np.asarray(labelmaps).astype(bool)
    eval_pc['thresholdType'] = threshold_type
    eval_pc['DiceScore'] = Metrics.dice(diffs_thresholded, labelmaps)
    eval_pc['DiceScorePerPatient'], eval_pc['PrecisionPerPatient'], eval_pc['RecallPerPatient'] = [], [], []
    eval_pc['TPCC'] = eval_pc['FPCC'] = eval_pc['FNCC'] = 0
    with np.errstate(divide='ignore', invalid='ignore'):
        for p in range(num_patients):
            sl = slice(p * num_slices, (p + 1) * num_slices)
            pred, gt = diffs_thresholded[sl], labels[sl]
            eval_pc['DiceScorePerPatient'] += [Metrics.dice(pred, gt)]
            eval_pc['PrecisionPerPatient'] += [Metrics.precision(pred, gt)]
            eval_pc['RecallPerPatient'] += [Metrics.recall(pred, gt)]
            tps, fps, fns = compute_detection_rate(np.squeeze(diffs_thresholded_at_precision70[sl]), np.squeeze(gt))
            eval_pc['TPCC'] += tps
            eval_pc['FPCC'] += fps
            eval_pc['FNCC'] += fns
        for key in ('DiceScore', 'Precision', 'Recall'):
            vals = np.array(eval_pc[key + 'PerPatient'])
            eval_pc[key + 'PerPatientMean'], eval_pc[key + 'PerPatientStd'] = np.mean(vals), np.std(vals)
        eval_pc['TP'], eval_pc['FP'], eval_pc['TN'], eval_pc['FN'] = Metrics.confusion_matrix(diffs_thresholded, labels)
        eval_pc['TPR'] = Metrics.tpr(diffs_thresholded, labels)
        eval_pc['FPR'] = Metrics.tpr(diffs_thresholded, labels)      # sic: the reference computes FPR with Metrics.tpr (:490)
        eval_pc['VD'] = Metrics.vd(diffs_thresholded, labels)
    eval_pc['TPRCC'] = eval_pc['TPCC'] / (eval_pc['TPCC'] + eval_pc['FNCC']) if eval_pc['TPCC'] + eval_pc['FNCC'] > 0 else 0.0
    eval_pc['PrecisionCC'] = eval_pc['TPCC'] / (eval_pc['TPCC'] + eval_pc['FPCC']) if eval_pc['TPCC'] + eval_pc['FPCC'] > 0 else 0.0
    return eval_pc


def residual_on_device(x, x_rec, mask, prior_quantile, keep_positive, apply_prior, device):
    """[N,H,W] float32 stacks -> float64 sub-volume, via uad_residual_score (bit-exact with Evaluation.py:282-291)."""
    return score_volume_on_device(x, x_rec, mask, 0, prior_quantile, keep_positive, apply_prior, False, device)


def score_volume_on_device(x, x_rec, brainmask, erode_iterations, prior_quantile, keep_positive, apply_prior, median, device):
    """One patient's residual sub-volume entirely on the GPU (Evaluation.py:282-291 + :311-312):
    brain-mask erosion (uad_binary_erosion_cross) -> residual / mask / prior (uad_residual_score) -> 5x5x5 median
    (uad_median_filter3d_5).  All three are bit-exact against the scipy / numpy reference; the float32 result is widened
    into the float64 ``subvolume`` exactly as the reference's ``subvolume[s] = x_diff`` store does."""
    st = torch.cuda.current_stream().cuda_stream
    xd = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(device)
    rd = torch.from_numpy(np.ascontiguousarray(x_rec, np.float32)).to(device)
    md = None
    if brainmask is not None:
        md = torch.from_numpy(np.ascontiguousarray(np.asarray(brainmask) != 0).astype(np.uint8).reshape(x.shape)).to(device)
        if erode_iterations:
            er = torch.empty_like(md)
            abi.call('uad_binary_erosion_cross', md.data_ptr(), er.data_ptr(), x.shape[0], x.shape[1], x.shape[2],
                     int(erode_iterations), st)
            md = er
    out = torch.empty_like(xd)
    abi.call('uad_residual_score', xd.data_ptr(), rd.data_ptr(), None if md is None else md.data_ptr(), float(prior_quantile),
             int(bool(keep_positive)), int(bool(apply_prior)), out.data_ptr(), xd.numel(), st)
    if median:
        filt = torch.empty_like(out)
        abi.call('uad_median_filter3d_5', out.data_ptr(), filt.data_ptr(), x.shape[0], x.shape[1], x.shape[2], st)
        out = filt
    sub = np.zeros(x.shape, np.float64)
    sub[...] = out.cpu().numpy()
    return sub


def brainmask_on_device(brainmask, erode_iterations, device):
    """The boolean brain mask `apply_brainmask` multiplies with (Evaluation.py:84-89), for a whole [Z,H,W] stack: the cross-shaped
    12-iteration erosion runs on the GPU (uad_binary_erosion_cross, bit-exact vs scipy)."""
    m = np.ascontiguousarray(np.asarray(brainmask) != 0).astype(np.uint8)
    if not erode_iterations:
        return m.astype(bool)
    st = torch.cuda.current_stream().cuda_stream
    md = torch.from_numpy(m).to(device)
    er = torch.empty_like(md)
    abi.call('uad_binary_erosion_cross', md.data_ptr(), er.data_ptr(), m.shape[0], m.shape[1], m.shape[2], int(erode_iterations), st)
    return er.cpu().numpy().astype(bool)


def _evaluate(datasetObj, modelObj, sampleDir, options, split="TEST", shard=None):
    """shard = (rank, world): score only this rank's volumes (volumes are the unit: the 5x5x5 median couples neighbouring
    slices); the integer Dice counts are all-reduced by Metrics.DeviceScorer."""
    eval_dict = get_eval_dictionary()
    patients = [datasetObj.patients[i] for i in datasetObj.get_patient_idx(split=split)]
    if shard is not None and shard[1] > 1:
        patients = patients[shard[0]::shard[1]]
    H, W = options['train']['outputHeight'], options['train']['outputWidth']
    device = modelObj.device
    for p, patient in enumerate(patients):
        filtered_files = patient['filtered_files']
        if type(filtered_files) is not list:
            filtered_files = [filtered_files]
        done = False
        for nii_filename in filtered_files:
            if done:
                continue
            nii, nii_seg, nii_skullmap = datasetObj.load_volume_and_groundtruth(nii_filename, patient)
            prior_quantile = np.quantile(nii.data, 0.9)
            if min(nii.shape()) < (datasetObj.options.sliceEnd - datasetObj.options.sliceStart):
                continue
            slice_start = datasetObj.options.sliceStart or 0
            slice_end = min(datasetObj.options.sliceEnd, nii.num_slices_along_axis(datasetObj.options.axis))
            xs, segs, skulls = [], [], []
            zoom_factor = None
            for s in range(slice_start, slice_end):
                slice_data = nii.get_slice(s, datasetObj.options.axis)
                slice_seg = nii_seg.get_slice(s, datasetObj.options.axis).astype(int)
                slice_skullmap = nii_skullmap.get_slice(s, datasetObj.options.axis).astype(int)
                if datasetObj.options.sliceResolution is not None and tuple(slice_data.shape) != tuple(datasetObj.options.sliceResolution):
                    zoom_factor = tuple([i / j for (i, j) in zip(datasetObj.options.sliceResolution, slice_data.shape)])
                    slice_data = scipy.ndimage.zoom(slice_data, zoom_factor)
                    slice_seg = scipy.ndimage.zoom(slice_seg, zoom_factor, mode="nearest")
                    slice_skullmap = scipy.ndimage.zoom(slice_skullmap, zoom_factor, mode="nearest")
                xs.append(slice_data.astype(np.float32))
                segs.append(slice_seg)
                skulls.append(slice_skullmap)
            x = np.stack(xs)                                       # [Z,H,W] float32
            _tmp = time.time()
            num_samples = options["numMonteCarloSamples"] if should(options, "numMonteCarloSamples") else 1
            recs = []
            for i in range(num_samples):                           # MC-dropout loop (:239-267), one batched pass per sample
                results = modelObj.reconstruct(x[..., None], dropout=num_samples > 1)
                recs.append(results['reconstruction'][..., 0])
            skull = np.stack([np.squeeze(m) for m in skulls])
            erode_it = 12 if should(options, "erodeBrainmask") else 0
            last_rec = recs[-1]                                    # l1err / l2err are those of the LAST pass (:274-276)
            var = None
            if num_samples > 1:
                # :253-267 - every sample is brain-masked BEFORE the statistics: mean and epistemic variance are 0 outside the
                # (eroded) mask; the scored reconstruction is the masked mean
                mask = brainmask_on_device(skull, erode_it, device)
                x_recs = np.array([np.multiply(mask, r) for r in recs])
                var = Metrics.combined_predictive_uncertainty(x_recs, np.zeros(x_recs.shape), axis=0, log_var=False)
                x_rec = np.mean(x_recs, axis=0).astype(np.float32)
            else:
                x_rec = recs[0]
            eval_dict['reconstructionTimes'] += [(time.time() - _tmp) / max(len(xs), 1)] * len(xs)
            subvolume = score_volume_on_device(x, x_rec, skull, erode_it, prior_quantile,
                                               should(options, "keepOnlyPositiveResiduals"),
                                               should(options, "applyHyperIntensityPrior"), should(options, "medianFiltering"),
                                               device)
            if var is not None:
                eval_dict['epistemic_variance'] += list(var)
            eval_dict['x'] += list(x[..., None])
            eval_dict['reconstructions'] += list(x_rec[..., None])
            eval_dict['labelmaps'] += segs
            for i in range(len(xs)):
                eval_dict['l1reconstructionErrors'] += [np.sum(np.abs(x[i] - last_rec[i]))]
                eval_dict['l2reconstructionErrors'] += [np.sum(np.sqrt((x[i] - last_rec[i]) ** 2))]
            eval_dict['diffs'] += [subvolume]
            if should(options, 'exportVolumes') and hasattr(nii_seg, 'set_subvolume'):
                export_patient_volume(nii_seg, subvolume, zoom_factor, datasetObj.options, options, sampleDir, patient['name'])
            done = True
    eval_dict['x'] = np.squeeze(np.array(eval_dict['x']))
    eval_dict['reconstructions'] = np.squeeze(np.array(eval_dict['reconstructions']))
    eval_dict['diffs'] = np.squeeze(np.array(eval_dict['diffs']))
    if eval_dict['diffs'].ndim > 3:
        d = eval_dict['diffs']
        eval_dict['diffs'] = np.reshape(d, [d.shape[0] * d.shape[1], d.shape[2], d.shape[3]])
    eval_dict['labelmaps'] = np.squeeze(np.array(eval_dict['labelmaps']))
    eval_dict['l1reconstructionErrorMean'] = np.mean(eval_dict['l1reconstructionErrors'])
    eval_dict['l1reconstructionErrorVariance'] = np.var(eval_dict['l1reconstructionErrors'])
    eval_dict['l2reconstructionErrorMean'] = np.mean(eval_dict['l2reconstructionErrors'])
    eval_dict['l2reconstructionErrorVariance'] = np.var(eval_dict['l2reconstructionErrors'])
    eval_dict['reconstructionTimes'] = np.mean(np.array(eval_dict['reconstructionTimes']))
    return eval_dict, patients


def pool_over_ranks(flat_d, flat_l):
    """Data-parallel evaluation: the threshold-free metrics (ROC / PRC, the 70 %-precision operating point) need every voxel, not
    just this rank's volumes - concatenate the residuals and labels of all ranks, in rank order, on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat_d, flat_l
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, (np.ascontiguousarray(flat_d), np.ascontiguousarray(flat_l).astype(np.uint8)))
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]).astype(int)


def _dp_context(model):
    """(shard, int64 all-reduce) when the trainer runs data-parallel (enable_data_parallel), else (None, None)."""
    world = int(getattr(model, 'world', 1) or 1)
    if world <= 1:
        return None, None
    from .. import dist as udist
    return (udist.rank(), world), udist.allreduce_sum_


def evaluate(datasetPC, gan, options, epoch='last', description=None):
    """Evaluation.py:372-526 minus PNG export / plots: returns the evalPC dictionary with the reference's keys (plus the short
    aliases DICE / AUC / AUPRC / Precision used by run.py) and writes rocPC.npy, prcPC.npy, evalPC.npy, evalPC.txt into the
    reference's directory layout <samplesDir>/<network>/<model_dir>/eval-<epoch>-<timestamp>[-<description>]/."""
    model = gan
    _time = {'evaluation': time.time()}
    histogram_range = (0.01, 0.075)
    eval_dir = os.path.join(options['train']['samplesDir'], model.network.__name__, model.model_dir,
                            'eval-' + str(epoch) + '-' + time.strftime('%Y-%m-%d %H-%M-%S'))
    if description is not None:
        eval_dir += '-' + str(description)
    shard, reduce_ = _dp_context(model)
    if shard is not None and shard[1] > 1:
        n_test = len(datasetPC.get_patient_idx(split="TEST"))
        if n_test < shard[1]:                           # every rank sees the same dataset: the same error everywhere, before any collective
            raise ValueError(f'data-parallel evaluation shards by volume: {n_test} test volumes cannot occupy {shard[1]} ranks')
        if shard[0] > 0:                                # one directory per rank (rank 0 keeps the reference's name and holds the pooled figures)
            eval_dir += f'-rank{shard[0]}'
    sample_dir = os.path.join(eval_dir, 'samples_test_PC')
    os.makedirs(sample_dir, exist_ok=True)
    eval_pc, patients = _evaluate(datasetPC, model, sample_dir, options, "TEST", shard=shard)
    diffs = eval_pc['diffs']
    labels = (eval_pc['labelmaps'] > 0)
    scorer = Metrics.DeviceScorer(diffs, labels, device=model.device, allreduce=reduce_)
    flat_d, flat_l = diffs.flatten(), labels.flatten().astype(int)
    if reduce_ is not None:                             # ROC / PRC / the precision-70 threshold over ALL ranks' voxels
        flat_d, flat_l = pool_over_ranks(flat_d, flat_l)
    eval_pc['diffHistogram'], _ = np.histogram(flat_d, bins='auto', range=histogram_range)
    if len(eval_pc.get('epistemic_variance', [])) > 0:
        ev = np.asarray(eval_pc['epistemic_variance'])
        eval_pc['uncertaintyHistogram'], _ = np.histogram(ev, bins=50, range=(1e-5, np.percentile(ev[ev >= 0], 99.8)))
    t0 = time.time()
    eval_pc['diff_AUC'], _fpr, _tpr, _roc_threshs = Metrics.compute_roc(flat_d, flat_l)
    _time['ROC'] = time.time() - t0
    if should(options, 'exportROC'):
        np.save(os.path.join(eval_dir, 'rocPC.npy'), {'fpr': _fpr, 'tpr': _tpr, 'threshs': _roc_threshs}, allow_pickle=True)
    t0 = time.time()
    eval_pc['diff_AUPRC'], _precisions, _recalls, _threshs = Metrics.compute_prc(flat_d, flat_l)
    _time['PRC'] = time.time() - t0
    if should(options, 'exportPRC'):
        np.save(os.path.join(eval_dir, 'prcPC.npy'), {'precisions': _precisions, 'recalls': _recalls, 'threshs': _threshs}, allow_pickle=True)
    eval_pc['AUC'], eval_pc['AUPRC'] = eval_pc['diff_AUC'], eval_pc['diff_AUPRC']
    # operating point at (at most) 70 % precision for the lesion-wise detection rate (:438-440)
    idx_precision70 = min(int(np.argmax(_precisions <= 0.7)), len(_threshs) - 1)
    mask70 = filter_3d_connected_components(np.squeeze(diffs > _threshs[idx_precision70]).copy())
    t0 = time.time()
    best_dice, best_thresh = Metrics.compute_dice_curve_recursive(diffs, labels, granularity=10, scorer=scorer)
    eval_pc['diceSearchTime'] = _time['DiceCurve'] = time.time() - t0
    eval_pc['bestDiceScore'], eval_pc['bestThreshold'] = best_dice, best_thresh
    threshold = best_thresh if options['threshold'] == 'bestdice' else float(options['threshold'])
    eval_pc['threshold'] = threshold
    mask = scorer.threshold_mask(threshold).cpu().numpy().astype(bool).reshape(diffs.shape)     # == diffs > threshold, bit-exact
    if options['threshold'] != 'bestdice':
        mask70 = mask
    mask = filter_3d_connected_components(np.squeeze(mask).copy())
    eval_pc['thresholded'] = mask
    n_per = diffs.shape[0] // max(len(patients), 1)
    summarize_predictions(eval_pc, eval_pc['labelmaps'], mask, mask70, len(patients), n_per, options['threshold'])
    eval_pc['DICE'], eval_pc['perPatientDice'] = eval_pc['DiceScore'], list(eval_pc['DiceScorePerPatient'])
    with np.errstate(divide='ignore', invalid='ignore'):
        eval_pc['Precision'] = Metrics.precision(mask, labels)
    if reduce_ is not None:                             # data-parallel scoring: the pooled figures from the rank-summed integer counts
        c = torch.tensor([int(eval_pc[k]) for k in ('TP', 'FP', 'TN', 'FN', 'TPCC', 'FPCC', 'FNCC')], dtype=torch.int64, device=model.device)
        reduce_(c)
        tp, fp, tn, fn, tpcc, fpcc, fncc = (np.int64(v) for v in c.cpu().numpy())
        eval_pc.update(TP=tp, FP=fp, TN=tn, FN=fn, TPCC=int(tpcc), FPCC=int(fpcc), FNCC=int(fncc))
        with np.errstate(divide='ignore', invalid='ignore'):
            eval_pc['DICE'] = eval_pc['DiceScore'] = (2 * tp) / ((tp + fp) + (tp + fn))
            eval_pc['TPR'] = eval_pc['FPR'] = tp / (tp + fn)
            eval_pc['Precision'] = tp / (tp + fp)
        eval_pc['TPRCC'] = tpcc / (tpcc + fncc) if tpcc + fncc > 0 else 0.0
        eval_pc['PrecisionCC'] = tpcc / (tpcc + fpcc) if tpcc + fpcc > 0 else 0.0
    _time['evaluation'] = time.time() - _time['evaluation']
    eval_pc['evalDir'] = eval_dir
    big = ('x', 'diffs', 'labelmaps', 'l1reconstructionErrors', 'l2reconstructionErrors', 'reconstructions', 'diffHistogram', 'thresholded',
           'epistemic_variance')
    stored = {k: v for k, v in eval_pc.items() if k not in big}
    np.save(os.path.join(eval_dir, 'evalPC.npy'), stored)
    with open(os.path.join(eval_dir, 'evalPC.txt'), 'w') as f:
        f.write(str(stored))
    return eval_pc


def determine_threshold_on_labeled_patients(dataset_pc, model, options, epoch='last', description=None):
    """Evaluation.py:529-567: best-Dice threshold over the labelled VALIDATION patients of one or several datasets (same signature).
    The reference evaluates split="VAL"; its own BrainWeb lesion set puts every patient into TEST (default_config_setup.py:211), which
    leaves that call without data there - a dataset with no VAL patients is therefore scored on its TEST split here."""
    eval_dir = os.path.join(options['train']['samplesDir'], model.network.__name__, model.model_dir,
                            'eval-' + str(epoch) + '-' + time.strftime('%Y-%m-%d %H-%M-%S'))
    if description is not None:
        eval_dir += '-' + str(description)
    sample_dir = os.path.join(eval_dir, 'samples_val_PC')
    os.makedirs(sample_dir, exist_ok=True)
    if not isinstance(dataset_pc, (list, tuple)):
        dataset_pc = [dataset_pc]
    diffs, labels = [], []
    shard, reduce_ = _dp_context(model)
    for ds in dataset_pc:
        split = 'VAL' if len(ds.get_patient_idx(split='VAL')) > 0 else 'TEST'
        ev, _ = _evaluate(ds, model, sample_dir, options, split, shard=shard)
        diffs.append(ev['diffs'])
        labels.append(ev['labelmaps'] > 0)
    diffs, labels = np.concatenate(diffs, 0), np.concatenate(labels, 0)
    scorer = Metrics.DeviceScorer(diffs, labels, device=model.device, allreduce=reduce_)
    return Metrics.compute_dice_curve_recursive(diffs, labels, granularity=10, scorer=scorer)
