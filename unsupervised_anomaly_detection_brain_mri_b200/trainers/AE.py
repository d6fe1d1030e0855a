"""Dense-bottleneck AE trainer (mirror of reference trainers/AE.py): L1 loss (:28-29), train loop, reconstruct."""
from .AEMODEL import AEMODEL, Phase, indicate_early_stopping, update_log_dicts  # noqa: F401


class AE(AEMODEL):
    pass
