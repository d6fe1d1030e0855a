"""B200-native hot path for unsupervised brain-MRI anomaly detection (conv AE/VAE/ceVAE train + score).

Host side mirrors the reference's ``models/`` / ``trainers/`` / ``utils/`` interface for this path; all arithmetic runs
in hand-written sm_100a CUDA behind the C ABI in ``include/uad_b200.h`` (``libuad_b200.so``).  There is no CPU fallback.
"""
__all__ = ['abi', 'engine']
