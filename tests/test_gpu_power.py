"""GPU: the iterated workload's building blocks (y in original order written into a slab of the next x,
device-side norm and scaling) against a numpy power iteration on the oracle's CSR product."""
import numpy as np
import pytest

import oracle
from cases import get, x_for

pytestmark = pytest.mark.gpu


def test_power_iteration_matches_oracle(cuda_device):
    import torch

    import dasp_b200

    m, n, rp, ci, v = get("symmetric_like")
    x0 = x_for(n)
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, m, n, rp, ci, v)
    s = torch.cuda.current_stream().cuda_stream
    xa = torch.from_numpy(x0).to(cuda_device)
    xb = torch.zeros_like(xa)
    norm2 = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    x_ref = x0.copy()
    for step in range(6):
        h.spmv_unpermuted(xa, xb, s)
        dasp_b200.sumsq(xb, m, norm2, s)
        dasp_b200.scale_rsqrt(xb, m, norm2, s)
        xa, xb = xb, xa
        y = oracle.csr_spmv_f64(m, rp, ci, v, x_ref)
        nrm = np.sqrt(np.dot(y, y))
        x_ref = y / nrm
        torch.cuda.synchronize()
        assert abs(np.sqrt(norm2.item()) - nrm) <= 1e-12 * nrm
        got = xa.cpu().numpy()
        assert np.linalg.norm(got - x_ref) <= 1e-12 * np.linalg.norm(x_ref), f"step {step}"
    h.close()


def test_sumsq_large_and_empty(cuda_device):
    import torch

    import dasp_b200

    v = torch.arange(1, 300001, dtype=torch.float64, device=cuda_device) * 1e-3
    out = torch.zeros(1, dtype=torch.float64, device=cuda_device)
    dasp_b200.sumsq(v, v.numel(), out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = float((v.cpu().numpy() ** 2).sum())
    assert abs(out.item() - want) <= 1e-12 * want
    dasp_b200.sumsq(v, 0, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert out.item() == 0.0
