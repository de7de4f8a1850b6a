"""GPU, two ranks (skipped on a single-GPU box): the row-partitioned iterated workload over torch.distributed / NCCL and
the NVLink exchanges.  Each case runs `bench.py --power-iter` under torchrun on 2 GPUs and compares the eigenvalue estimate
and the checksum of the final iterate with the single-GPU run of the same workload: the fused relabelled exchange (`perm`),
the 1.5-D partition that splits the long rows by columns (`hybrid`), the NCCL broadcast exchange (`bcast`), and the un-permute
passes through the multicast / peer mappings (`mcu`, `p2pu`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["--workload", "c5_spec", "--scale", "0.02", "--power-iter", "12", "--warmup", "3", "--no-secondary", "--no-cpu", "--no-others"]


def _run(nproc, exchange, port):
    if nproc == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--exchange", exchange] + COMMON
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(nproc), "--exchange", exchange] + COMMON
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout[-1500:]
    return json.loads(lines[0])


@pytest.fixture(scope="module")
def single(cuda_device):
    return _run(1, "bcast", 0)


@pytest.mark.parametrize("exchange", ["perm", "hybrid"])
def test_relabelled_and_hybrid_single_gpu_match_unpermuted(cuda_device, single, exchange):
    d = _run(1, exchange, 0)
    assert abs(d["eigenvalue_estimate"] - single["eigenvalue_estimate"]) <= 1e-10 * abs(single["eigenvalue_estimate"])
    assert abs(d["x_checksum"] - single["x_checksum"]) <= 1e-8 * max(1.0, abs(single["x_checksum"]))


@pytest.mark.parametrize("exchange", ["perm", "hybrid", "bcast", "mcu", "p2pu"])
def test_two_ranks_reproduce_the_single_gpu_iteration(cuda_device, single, exchange):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    d = _run(2, exchange, 29611 + ["perm", "hybrid", "bcast", "mcu", "p2pu"].index(exchange))
    assert d["n_gpus"] == 2 and len(d["config"]["slab_rows"]) == 2 and min(d["config"]["slab_rows"]) > 0
    assert abs(d["eigenvalue_estimate"] - single["eigenvalue_estimate"]) <= 1e-10 * abs(single["eigenvalue_estimate"])
    assert abs(d["x_checksum"] - single["x_checksum"]) <= 1e-8 * max(1.0, abs(single["x_checksum"]))
