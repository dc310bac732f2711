"""pkg/exch2 tile graph spread over several GPUs (needs >= 2 GPUs; the host side of it is covered without a GPU by
tests/test_exch2_dist_cpu.py).  Launches tests/dist_cs_worker.py under torchrun: exchanges bit-identical to the
literal exch2 algorithm, CG2D / CG2D_SR on the config-4 operator, adjustment.cs-32x32x1 against its golden output."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


NPROC = [int(n) for n in os.environ.get("MITGCM_B200_DIST_NPROC", "2").split(",")]      # "2,3,4" on a larger box


@pytest.mark.parametrize("nproc", NPROC)
def test_tile_graph_across_gpus(nproc):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs (torchrun --nproc-per-node {nproc} tests/dist_cs_worker.py)")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + nproc),
                        os.path.join(here, "dist_cs_worker.py")], capture_output=True, text=True, timeout=900)
    assert "DIST_CS PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
