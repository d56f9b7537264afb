"""Driver-visible multi-rank parity (a15 MinAll, a16 halo exchange, cross-rank coarse-fine
transfers, cut-cell level over two ranks): spawns tests/multigpu_check.py under
torch.distributed.run with one rank per GPU -- NCCL halo exchange (ncclSend / ncclRecv), the
stable-dt all-reduce, the coarse-fine staging transfers -- and requires every rank's boxes and
the dt sequence to equal the CPU oracle of the whole level bit for bit.  Needs at least two
GPUs; with one the NCCL paths cannot run and the test is skipped (bench.py --gpus N repeats the
halo part at full size: "parity_check")."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    from finitevolumesolver_b200 import capi
    return capi.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_multirank_halo_dt_amr_eb_bitwise(nranks):
    n = _gpus()
    if n < nranks:
        pytest.skip(f"needs {nranks} GPUs, this box has {n}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1", "--master-port",
           str(_free_port()), os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"multigpu_check ok: {nranks} ranks" in r.stdout, r.stdout[-2000:]
