"""dasp_create timing on the C4 stencil: repeated creates in one process, with and without the slab allocator, plus one
traced create (DASP_TRACE_PREPROCESS=1: host time per phase, each mark synchronises)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dasp_b200  # noqa: E402
from dasp_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
dasp_b200.load()
g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
spec = synth.stencil27(g)
rp, ci, v, nnz = synth.generate(spec, 0, int(spec.m), dev)
for label, env in (("slab", {}), ("slab", {}), ("no slab", {"DASP_NO_SLAB": "1"}), ("no slab", {"DASP_NO_SLAB": "1"}), ("slab", {}),
                   ("slab, traced", {"DASP_TRACE_PREPROCESS": "1"})):
    for k in ("DASP_NO_SLAB", "DASP_TRACE_PREPROCESS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    torch.cuda.synchronize()
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, int(spec.m), int(spec.n), rp, ci, v, device=0, nnz=nnz)
    st = h.stats()
    print(f"{label:14s} preprocess_ms {st['preprocess_ms']:8.3f}  device_bytes {st['device_bytes']}", flush=True)
    h.close()
