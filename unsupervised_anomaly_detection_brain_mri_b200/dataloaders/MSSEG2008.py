"""MSSEG2008 lesion dataset (interface of reference dataloaders/MSSEG2008.py: class MSSEG2008, folders {UNC,CHB}_{train,test}/<patient>/
<patient>_{FLAIR,T1,T2}.aligned.nii.gz, _lesion.aligned.nii.gz, _skullmap.nii.gz).  Only the `aligned` (NIfTI) format the reference's
default_config_setup selects is read.  The reference's `raw` path is dead code there: it calls NRRD.denoise() / apply_skullmap(),
which dataloaders/NRRD.py does not define (MSSEG2008.py:238-262, NRRD.py:7-76) - it is refused here with a clear error."""
import os

from ._lesion_dataset import LesionDataset


class MSSEG2008(LesionDataset):
    NAME = 'MSSEG2008'
    PROTOCOL_MAPPINGS = ['FLAIR', 'T1', 'T2']

    class Options(LesionDataset.Options):
        def __init__(self):
            super().__init__()
            self.folderTrainUNC = 'UNC_train'
            self.folderTestUNC = 'UNC_test'
            self.folderTrainCHB = 'CHB_train'
            self.folderTestCHB = 'CHB_test'
            self.filterScanner = 'UNC'           # UNC or CHB
            self.filterType = 'train'            # train or test

    @staticmethod
    def get_patients(options):
        if options.format != 'aligned':
            raise NotImplementedError('MSSEG2008: only format="aligned" (NIfTI) is supported (the raw NRRD path cannot run in the reference either)')
        patients = []
        for folder in (options.folderTrainUNC, options.folderTestUNC, options.folderTrainCHB, options.folderTestCHB):
            if options.filterScanner and options.filterScanner not in folder:
                continue
            if options.filterType and options.filterType not in folder:
                continue
            base = os.path.join(options.dir, folder)
            if not os.path.isdir(base):
                continue
            for pname in sorted(e.name for e in os.scandir(base) if e.is_dir()):
                full = os.path.join(base, pname)
                patient = {'name': pname, 'fullpath': full, 'type': 'train' if 'train' in folder else 'test', 'filtered_files': []}
                for protocol in MSSEG2008.PROTOCOL_MAPPINGS:
                    patient[protocol] = os.path.join(full, pname + '_' + protocol + '.aligned.nii.gz')
                    if len(options.filterProtocols) == 0 or protocol in options.filterProtocols:
                        patient['filtered_files'].append(patient[protocol])
                patient['groundtruth'] = os.path.join(full, pname + '_lesion.aligned.nii.gz')
                patient['skullmap'] = os.path.join(full, pname + '_skullmap.nii.gz')
                patients.append(patient)
        return patients
