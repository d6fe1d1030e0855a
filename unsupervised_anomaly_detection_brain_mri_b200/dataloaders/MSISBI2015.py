"""MSISBI2015 lesion dataset (interface of reference dataloaders/MSISBI2015.py: class MSISBI2015, layout
<dir>/training0{1..5}/preprocessed/<name>_{flair,mprage,pd,t2}_pp.nii | <name>_<protocol>.aligned.nii.gz, masks/<name>_mask1.nii)."""
import glob
import os

from ._lesion_dataset import LesionDataset


class MSISBI2015(LesionDataset):
    NAME = 'MSISBI2015'
    PROTOCOL_MAPPINGS = {'FLAIR': ['flair'], 'MPRAGE': ['mprage'], 'PD': ['pd'], 'T2': ['t2']}

    @staticmethod
    def get_patients(options):
        patients = []
        for folder in ('training01', 'training02', 'training03', 'training04', 'training05'):
            pre = os.path.join(options.dir, folder, 'preprocessed')
            for path in sorted(glob.glob(os.path.join(pre, folder + '_*_flair_pp.nii'))):
                name = os.path.basename(path).replace('_flair_pp.nii', '')
                patient = {'name': name, 'fullpath': pre, 'filtered_files': []}
                for protocol, aliases in MSISBI2015.PROTOCOL_MAPPINGS.items():
                    if len(options.filterProtocols) > 0 and protocol not in options.filterProtocols:
                        continue
                    fn = name + '_' + aliases[0] + ('_pp.nii' if options.format == 'raw' else '.aligned.nii.gz')
                    patient[protocol] = os.path.join(pre, fn)
                    patient['filtered_files'].append(patient[protocol])
                if options.format == 'raw':
                    patient['groundtruth'] = os.path.join(options.dir, folder, 'masks', name + '_mask1.nii')
                    patient['skullmap'] = os.path.join(pre, name + '_skullmap.nii.gz')
                else:
                    patient['groundtruth'] = os.path.join(pre, name + '_mask1.aligned.nii.gz')
                    patient['skullmap'] = os.path.join(pre, name + '_skullmap.aligned.nii.gz')
                patients.append(patient)
        return patients
