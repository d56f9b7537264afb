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


@pytest.mark.parametrize("nranks,halo", [(2, "nccl"), (2, "p2p"), (4, "nccl"), (8, "nccl")])
def test_multirank_halo_dt_amr_eb_bitwise(nranks, halo):
    """halo = p2p: the opt-in exchange by direct stores into the peers' ghost cells (CUDA IPC over
    NVLink, FVS_HALO_P2P=1) must give the same bits."""
    n = _gpus()
    if n < nranks:
        pytest.skip(f"needs {nranks} GPUs, this box has {n}")
    env = dict(os.environ, FVS_HALO_P2P="1" if halo == "p2p" else "0")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1", "--master-port",
           str(_free_port()), os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"multigpu_check ok: {nranks} ranks" in r.stdout, r.stdout[-2000:]
