"""Multi-GPU path on real devices (skipped on single-GPU boxes): utterance-sharded synthesis whose final exchange is
fused into the last GEMM's epilogue -- each rank stores mel_post straight into rank 0's peer-mapped window over
NVLink -- must equal the plain NCCL gather bit for bit.  The CPU-side logic is covered by tests/test_batching.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs on one node')
def test_peer_window_gather_equals_nccl_gather():
    env = dict(os.environ, N_UTTS='48')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29541', str(ROOT / 'scripts' / 'p2p_gather_check.py')],
                       capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'p2p gather == nccl gather for 48 utterances' in r.stdout
