"""MSLUB lesion dataset (interface of reference dataloaders/MSLUB.py: class MSLUB, Options, get_patients layout
<dir>/data/<patient>/<patient>_{FLAIR,T1W,T1WKS,T2W}[.aligned].nii.gz, _consensus_gt, _brainmask); shared logic in _lesion_dataset."""
import os

from ._lesion_dataset import LesionDataset


class MSLUB(LesionDataset):
    NAME = 'MSLUB'
    PROTOCOL_MAPPINGS = {'FLAIR': ['FLAIR'], 'T1': ['T1W'], 'TWKS': ['T1WKS'], 'T2': ['T2W']}

    @staticmethod
    def get_patients(options):
        ext = '.nii.gz' if options.format == 'raw' else '.aligned.nii.gz'
        patients = []
        base = os.path.join(options.dir, 'data')
        for pname in sorted(e.name for e in os.scandir(base) if e.is_dir()):
            full = os.path.join(base, pname)
            patient = {'name': pname, 'fullpath': full, 'filtered_files': []}
            for protocol, aliases in MSLUB.PROTOCOL_MAPPINGS.items():
                patient[protocol] = os.path.join(full, pname + '_' + aliases[0] + ext)
                if len(options.filterProtocols) == 0 or protocol in options.filterProtocols:
                    patient['filtered_files'].append(patient[protocol])
            patient['groundtruth'] = os.path.join(full, pname + '_consensus_gt' + ext)
            patient['skullmap'] = os.path.join(full, pname + '_brainmask' + ext)
            patients.append(patient)
        return patients
