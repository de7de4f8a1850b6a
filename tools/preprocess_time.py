"""dasp_create timing on the C4 stencil.  `python tools/preprocess_time.py 256 KEY=VAL ...`: sets the environment given on
the command line (A/B aids: DASP_NO_SLAB=1, DASP_SLAB_MAX_MB=n, DASP_TRACE_PREPROCESS=1), then creates the handle three times in
this process (the first create also pays CUDA's lazy module loading) and prints the event time of each analysis."""
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
for kv in sys.argv[2:]:
    k, v = kv.split("=", 1)
    os.environ[k] = v
spec = synth.stencil27(g)
rp, ci, v, nnz = synth.generate(spec, 0, int(spec.m), dev)
torch.cuda.synchronize()
for i in range(3):
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, int(spec.m), int(spec.n), rp, ci, v, device=0, nnz=nnz)
    st = h.stats()
    print(f"{' '.join(sys.argv[2:]) or 'default':40s} create #{i + 1}: preprocess_ms {st['preprocess_ms']:8.3f}  device_bytes {st['device_bytes']}", flush=True)
    h.close()
