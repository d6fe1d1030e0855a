"""GPU (needs >= 2 devices): data-parallel equivalence through real NCCL, launched the way the driver launches bench.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_gpu_step_equals_single_gpu_step_on_concatenated_batch():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(here, 'dp_equiv_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert 'DP_EQUIV world=2' in out.stdout and 'DP_EQUIV_GRAPH world=2' in out.stdout and 'DP_EQUIV_PEER world=2' in out.stdout and 'DP_EQUIV_FANOGAN_PEER world=2' in out.stdout
    print(''.join(l + '\n' for l in out.stdout.splitlines() if l.startswith('DP_EQUIV')))
    assert 'DP_EQUIV_FANOGAN world=2' in out.stdout and 'DP_EQUIV_SCORING world=2' in out.stdout
