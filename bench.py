#!/usr/bin/env python
"""bench.py — DASP SpMV throughput on B200 (metric of BASELINE.json: GFLOPS = 2*nnz/t and achieved
HBM GB/s on algorithmic bytes), one JSON line on rank 0.

  python bench.py [--gpus N --steps K --warmup W] [--workload c4|c1|c2|c3|c5] [--impl reference]

A step = one y = A*x over the whole (partitioned) matrix.  Default workload: C4 of BASELINE.json, the
27-point stencil on 256^3 (16.7 M rows, 449 M nnz, FP64) — the configuration the 1/2/4/8-GPU scaling
target is quoted on.  With N > 1 (launched by torchrun, one rank per GPU) the matrix is cut into
nnz-balanced contiguous row slabs, x is replicated, every rank multiplies its slab, no collective on
the data path (strong scaling); time = max over ranks of the device time of K back-to-back launches.

Keys beyond the base contract: roofline (dominant kernel vs MEASURED_PEAKS.json), cpu_baseline (serial
CSR loop of oracle/ on host cores, bounded sample), e2e (host buffers through dasp_spmv_host: H2D x,
kernel, D2H y every step), secondary (cuSPARSE CSR SpMV via torch, and the reference's own DASP kernels
recompiled for sm_100a from oracle/_ref, on a bounded sample).

--impl reference: the serial-CSR definition of the reference result (oracle/, row-parallel over all host
threads) on a bounded sample of the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5", "c3_spec", "c5_spec"])
    ap.add_argument("--grid", type=int, default=256, help="c4: stencil grid edge")
    ap.add_argument("--scale", type=float, default=1.0, help="c3/c5: fraction of the named size")
    ap.add_argument("--variant", default="auto", choices=["auto", "cuda", "mma", "split", "tma", "blocked", "banded", "nobands", "mband", "sband"])
    ap.add_argument("--no-secondary", action="store_true", help="skip cuSPARSE / reference-kernel comparison")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--cold", action="store_true", help="flush L2 before every timed launch (small workloads)")
    ap.add_argument("--no-others", action="store_true", help="skip the short measurements of the other BASELINE configs")
    ap.add_argument("--half", action="store_true", help="run the chosen workload in FP16 (values and x rounded to half)")
    ap.add_argument("--no-index-compression", action="store_true", help="A/B aid: kernels read reg_cid instead of the compact 16-bit indices")
    ap.add_argument("--categories", type=int, default=15, help="profiling aid: category mask (1 long, 2 medium, 4 short, 8 empty)")
    ap.add_argument("--breakdown", action="store_true", help="also time each row category alone (profiling aid)")
    ap.add_argument("--exchange", default="perm", choices=["perm", "hybrid", "bcast", "a2a", "p2p", "mc", "mcu", "p2pu", "mcc", "p2pc"],
                    help="power iteration: how the y slabs reach every rank (perm, default: relabelled P*A*P^T mode, the iterate lives "
                         "in permuted order, the SpMV kernel itself stores its slab of y, contiguous, into every rank's next x through one "
                         "NVSwitch multicast mapping (peer mappings if there is none); the all-reduce of the norm is the only collective; bcast: one NCCL broadcast per slab; a2a: all-to-all "
                         "into equal chunks + all-gather, measured slower: 3.18 vs 2.34 ms per step on 8 GPUs; p2p / mc: fused, the "
                         "SpMV kernel stores y straight into every peer's next x through NVLink peer mappings / one NVSwitch "
                         "multicast mapping, the all-reduce of the norm is the only collective)")
    ap.add_argument("--slab", default="", metavar="R/W", help="diagnostic: run rank R's slab of a W-way partition on one GPU, no process group")
    ap.add_argument("--no-iterated", action="store_true", help="skip the C5 100-step power-iteration leg (`iterated` key)")
    ap.add_argument("--power-iter", type=int, default=0, metavar="K",
                    help="iterated workload: K steps of x <- A x / ||A x|| with the y slabs gathered over NCCL every step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workloads

def make_spec(args):
    from dasp_b200 import synth

    w = args.workload
    if w == "c4":
        g = args.grid
        return synth.stencil27(g), f"C4 27-point stencil {g}^3 fp64", False
    if w in ("c1", "c2"):
        half = w == "c2"
        return synth.banded(), ("C2" if half else "C1") + " cop20k_A stand-in (banded symmetric, seed 7; real .mtx is a missing blob) " + ("fp16" if half else "fp64"), half
    if w == "c3":
        m = int(10_000_000 * args.scale)
        return synth.powerlaw(m=m), f"C3 power-law alpha=0.95 {m} rows, CSR-sorted variant (ascending windowed columns) fp64", False
    if w == "c3_spec":
        m = int(10_000_000 * args.scale)
        return synth.powerlaw_spec(m=m), f"C3 power-law alpha=0.95 {m} rows as SURVEY 8(d) words it (unsorted distinct columns: 90% +-4096 window, 10% global) fp64", False
    m = int(50_000_000 * args.scale)
    nl = max(1, int(1000 * args.scale))
    if w == "c5_spec":
        return synth.skewed_spec(n_long=nl, n_short=m), f"C5 skewed {nl}x1M long rows at seeded positions, columns uniform over n without replacement, + {m} short rows (unsorted, +-4096 window) as SURVEY 8(d) words it fp64", False
    return synth.skewed(n_long=nl, n_short=m), f"C5 skewed {nl}x1M long rows, CSR-sorted variant (ascending columns, 50% of a common 2^21 band) + {m} short rows fp64", False


def algorithmic_bytes(m, n, nnz, esz):
    """CSR bytes, x read once, y written once: the reference's data_origin1 (src/main_f64.cu:143)."""
    return nnz * (esz + 4) + (m + 1) * 4 + n * esz + m * esz


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def ncu_traffic(args):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload, or None."""
    p = os.path.join(ROOT, "profiles", "r02", "traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01", "traffic.json")
    if not os.path.exists(p):
        return None
    key = {"c4": f"c4:{args.grid}"}.get(args.workload, f"{args.workload}:{args.scale}")
    e = json.load(open(p)).get(key)
    return e["traffic_bytes"] if e and int(os.environ.get("WORLD_SIZE", "1")) == 1 else None


def device_check(rp, ci, v, x, y_orig, rows):
    """Relative L2 error of ALL `rows` entries of y (original row order) against an independent evaluation of the CSR
    definition on the device in float64 (torch index_add_ in row chunks of at most 2^27 entries; no code of this
    repository, no oracle)."""
    import torch

    num = torch.zeros((), dtype=torch.float64, device=rp.device)
    den = torch.zeros((), dtype=torch.float64, device=rp.device)
    xd = x.double()
    r0 = 0
    while r0 < rows:
        target = rp[r0:r0 + 1].long() + (1 << 27)
        r1 = int(torch.searchsorted(rp, target.to(rp.dtype), right=True).item()) - 1
        r1 = min(rows, max(r1, r0 + 1))
        k0, k1 = int(rp[r0].item()), int(rp[r1].item())
        ref = torch.zeros(r1 - r0, dtype=torch.float64, device=rp.device)
        if k1 > k0:
            row_of = torch.repeat_interleave(torch.arange(r1 - r0, device=rp.device), (rp[r0 + 1:r1 + 1] - rp[r0:r1]).long())
            ref.index_add_(0, row_of, v[k0:k1].double() * xd[ci[k0:k1].long()])
            del row_of
        num += ((y_orig[r0:r1].double() - ref) ** 2).sum()
        den += (ref ** 2).sum()
        r0 = r1
    return float((num.sqrt() / den.sqrt().clamp_min(1e-300)).item())


def config_dict(args, spec, wname, m, n, nnz_total, world, b_alg_rank, small, cold):
    """The `config` object of the JSON line; both arms (ours / --impl reference) print the same one."""
    return {"workload": wname, "m": m, "n": n, "nnz": nnz_total, "seed": int(spec.seed),
            "partition": "nnz-balanced contiguous row slabs, x replicated" if world > 1 else "single GPU",
            "l2": ("L2 flushed before every launch" if (small and cold) else
                   ("inputs exceed L2 (%.0f MB per rank)" % (b_alg_rank / 1e6) if not small else
                    "warm L2: back-to-back launches on an L2-resident matrix (reference protocol, src/dasp_f64.h:1301-1311)")),
            "variant": args.variant, "threshold": 0.75, "block_longest": 256}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, read+write)"
    return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------------
# reference arm: serial-CSR definition on host cores, bounded sample

def host_csr(args, spec, target_nnz=None):
    """The workload (or, with target_nnz, its first rows holding about that many entries) as host CSR in float64."""
    import torch

    from dasp_b200 import synth

    if torch.cuda.is_available():
        dev = torch.device("cuda:0")
        rows = int(spec.m)
        if target_nnz is not None:
            ln = synth.row_lengths(spec, 0, spec.m, dev)
            cs = torch.cumsum(ln, 0)
            rows = min(int(torch.searchsorted(cs, torch.tensor([target_nnz], device=dev)).item()) + 1, int(spec.m))
            del ln, cs
        rp, ci, v, nnz = synth.generate(spec, 0, rows, dev)
        out = (rows, int(spec.n), rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy())
        del rp, ci, v
        torch.cuda.empty_cache()
        return out
    # no GPU: numpy twin (stencil only)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import matrices

    return matrices.stencil27(min(args.grid, 96))


def cpu_leg(args, spec, threads, steps, warmup, target_nnz):
    """Serial CSR loop of oracle/ (per thread: contiguous row ranges) on host cores; every step is one full pass."""
    import oracle

    m, n, rp, ci, v = host_csr(args, spec, target_nnz)
    nnz = int(rp[m])
    x = np.random.default_rng(7).uniform(-1, 1, n)
    for _ in range(warmup):
        oracle.csr_spmv_f64(m, rp, ci, v, x, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.csr_spmv_f64(m, rp, ci, v, x, threads=threads)
    t = (time.perf_counter() - t0) / steps
    whole = m == int(spec.m)
    return {"value": 2.0 * nnz / t / 1e9, "unit": "GFLOP/s", "cores": threads, "kind": "port",
            "sample": (f"the whole workload ({m} rows, {nnz} nnz)" if whole else f"rows [0,{m}) of the workload ({nnz} nnz)")
                      + f", serial CSR loop per thread, mean of {steps} passes after {warmup} warm-up passes",
            "ms_per_step": t * 1e3, "m": m, "n": n, "nnz": nnz,
            "hbm_gbs": algorithmic_bytes(m, n, nnz, 8) / t / 1e9}


def run_reference(args):
    """--impl reference: the reference ships no CPU SpMV; its result is DEFINED by the serial CSR loop (oracle/), here
    row-parallel over all host threads, on the WHOLE workload of our arm, K timed passes after W warm-up passes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    spec, wname, half = make_spec(args)
    threads = os.cpu_count() or 1
    # C4 (5.4 GB of host CSR, ~40 ms per pass on 16 threads) and smaller run whole; the 1.1 G-entry C5 shapes are sampled
    big = args.workload in ("c5", "c5_spec") and args.scale > 0.3
    leg = cpu_leg(args, spec, threads, args.steps, args.warmup, 200_000_000 if big else None)
    m, n, nnz = leg["m"], leg["n"], leg["nnz"]
    b_alg_rank = algorithmic_bytes(m // world, n, nnz // world, 8)
    cfg = config_dict(args, spec, wname, int(spec.m), int(spec.n), nnz, world, b_alg_rank, b_alg_rank < 256e6, args.cold)
    line = {
        "impl": "reference", "metric": "spmv_gflops", "value": leg["value"], "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "note": "the reference ships no CPU SpMV; this is the serial CSR loop that defines its result (oracle/), row-parallel on all host threads",
        "hbm_gbs": leg["hbm_gbs"],
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm

def bind_host_memory_to_gpu_node(local):
    """Multi-GPU end-to-end runs move every rank's x and y over its own PCIe link, but the pinned host buffers of all ranks
    come from whatever NUMA node the process happens to run on: prefer the node the GPU hangs off (sysfs numa_node of its PCI
    device) for this process's allocations - set_mempolicy(MPOL_PREFERRED) - and, where the cgroup allows it, run on that
    node's cores.  Best effort, reported in the bench line; never fatal."""
    info = {"gpu_numa_node": None, "mempolicy_preferred": False, "cpu_affinity": False}
    try:
        import ctypes

        import pynvml

        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(hnd).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238  # x86-64
        if libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), ctypes.c_ulong(64)) == 0:
            info["mempolicy_preferred"] = True
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpu_affinity"] = True
        except OSError:
            pass
    except Exception as e:  # no sysfs / no permission / not Linux x86-64: leave the defaults
        info["error"] = repr(e)[:120]
    return info


def run_ours(args):
    import torch
    import torch.distributed as dist

    import dasp_b200
    from dasp_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    fake = None
    if args.slab:
        fake = tuple(int(t) for t in args.slab.split("/"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    numa = bind_host_memory_to_gpu_node(local) if (world > 1 and not os.environ.get("DASP_BENCH_NO_NUMA")) else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dasp_b200.load()

    spec, wname, half = make_spec(args)
    if args.half and not half:
        half, wname = True, wname.replace("fp64", "fp16")
    esz = 2 if half else 8
    tdt = torch.float16 if half else torch.float64
    dtype = dasp_b200.DASP_F16 if half else dasp_b200.DASP_F64
    m, n = int(spec.m), int(spec.n)
    stream = torch.cuda.current_stream(dev).cuda_stream

    # nnz-balanced contiguous row slabs: cut p = smallest i with rowptr[i] >= p*nnz/P
    ln = synth.row_lengths(spec, 0, m, dev)
    cs = torch.cumsum(ln, 0)  # cs[i] = rowptr[i+1]
    nnz_total = int(cs[-1].item())
    parts = fake[1] if fake else world
    targets = torch.tensor([nnz_total * p // parts for p in range(1, parts)], device=dev, dtype=torch.int64)
    cuts = [0] + [int(c) + 1 for c in torch.searchsorted(cs, targets, right=False).tolist()] + [m]
    cuts = [min(c, m) for c in cuts]
    del ln, cs
    r0, r1 = (cuts[fake[0]], cuts[fake[0] + 1]) if fake else (cuts[rank], cuts[rank + 1])

    rp, ci, v, nnz = synth.generate(spec, r0, r1, dev, half=half)
    t0 = time.perf_counter()
    h = dasp_b200.Dasp(dtype, r1 - r0, n, rp, ci, v, device=local, nnz=nnz)
    create_s = time.perf_counter() - t0
    st = h.stats()
    # the first dasp_create of a process also pays CUDA's lazy module loading and the driver's first large allocations
    # (measured 7 ms to 0.6 s on a fresh box): report it separately and time the analysis itself on a second create
    first_create_ms = st["preprocess_ms"]
    if world == 1 and not fake:
        h.close()
        t0 = time.perf_counter()
        h = dasp_b200.Dasp(dtype, r1 - r0, n, rp, ci, v, device=local, nnz=nnz)
        create_s = time.perf_counter() - t0
        st = h.stats()
    # (medium, long, short) variants; A/B aids for the band kernels: "banded" forces both, "mband" / "sband" one of them
    # (the other category in the fused kernel), "nobands" neither; the long rows stay AUTO in all four
    triples = {"auto": (0, 0, 0), "cuda": (1, 1, 1), "mma": (2, 2, 2), "split": (3, 0, 0), "tma": (0, 4, 0), "blocked": (0, 5, 0),
               "banded": (6, 0, 6), "nobands": (1, 0, 1), "mband": (6, 0, 1), "sband": (1, 0, 6)}
    h.set_variant(*triples[args.variant])

    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    x = (torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(tdt)
    y = torch.zeros(max(r1 - r0, 1), dtype=tdt, device=dev)

    # parity check inside the bench: ALL rows of this slab against an independent float64 evaluation of the CSR
    # definition on the device (the oracle itself is only used by tests/, smoke() and the CPU-baseline legs)
    chk_rows = r1 - r0
    chk = None
    if chk_rows > 0:
        yy = torch.empty_like(y)
        h.spmv_unpermuted(x, yy, stream)
        torch.cuda.synchronize(dev)
        chk = device_check(rp, ci, v, x, yy, chk_rows)
        del yy
        if chk > (2e-3 if half else 1e-12) and not os.environ.get("DASP_BENCH_NOCHECK"):
            raise SystemExit(f"bench.py: parity check failed, relative L2 {chk}")

    keep_csr = (world == 1 and not args.no_secondary and not half)
    if not keep_csr:
        del rp, ci, v
        torch.cuda.empty_cache()

    if args.categories != 15:
        h.set_category_mask(args.categories)
    if args.no_index_compression:
        h.set_index_compression(False)
    small = algorithmic_bytes(r1 - r0, n, nnz, esz) < 256e6
    flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device=dev) if (small and args.cold) else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(steps, warmup):
        for _ in range(warmup):
            h.spmv(x, y, stream)
        barrier()
        if flush is None:
            # K back-to-back launches issued from C (the reference's loop, src/dasp_f64.h:1301-1311),
            # CUDA events on the launching stream
            ms = h.spmv_timed(x, y, stream, 0, steps)
        else:
            ms = 0.0
            for _ in range(steps):
                synth.flush_l2(flush)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                h.spmv(x, y, stream)
                e1.record()
                torch.cuda.synchronize(dev)
                ms += e0.elapsed_time(e1)
        barrier()
        return ms

    if args.power_iter > 0 and args.exchange == "hybrid":
        h.close()
        res = power_iteration_hybrid(args, spec, wname, cuts, rank, world, dev, local, x, args.power_iter, args.warmup, nnz_total)
        if rank == 0:
            print(json.dumps(res), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.power_iter > 0:
        res = power_iteration(args, h, x, cuts, rank, world, dev, spec, wname, nnz_total, nnz, timed)
        if rank == 0:
            print(json.dumps(res), flush=True)
        h.close()
        if world > 1:
            dist.destroy_process_group()
        return

    with ClockSampler(local) as clk:
        ms_local = timed(args.steps, args.warmup)
    t = torch.tensor([ms_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps

    breakdown = None
    if args.breakdown:
        breakdown = {}
        for name, mask in (("long", 1), ("medium", 2), ("short", 4)):
            h.set_category_mask(mask)
            breakdown[name + "_ms"] = timed(max(5, args.steps // 3), 2) / max(5, args.steps // 3)
        h.set_category_mask(15)
        breakdown.update({"nnz_long": st["nnz_long"], "nnz_short": st["nnz_short"],
                          "nnz_medium": st["origin_nnz_reg"] + st["nnz_irreg"]})

    # end to end through the C ABI with host buffers: every step uploads ITS x from pinned host memory, multiplies,
    # and downloads ITS y to pinned host memory.  Two protocols: one blocking call per step (dasp_spmv_host, the
    # reference's data movement), and the batch entry that pipelines independent steps over three streams.
    hxs = [torch.empty(n, dtype=tdt).pin_memory() for _ in range(3)]
    for j, hx in enumerate(hxs):
        hx.copy_((x * (1.0 + 0.25 * j)).cpu())
    hys = [torch.empty(max(r1 - r0, 1), dtype=tdt).pin_memory() for _ in range(3)]
    e2e_steps = max(6, min(args.steps, 12))
    for _ in range(2):
        h.spmv_host(hxs[0], hys[0])
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        h.spmv_host(hxs[i % 3], hys[i % 3])
    torch.cuda.synchronize(dev)
    e2e_serial_local = (time.perf_counter() - t0) * 1e3
    h.spmv_host_batch([hxs[i % 3] for i in range(3)], [hys[i % 3] for i in range(3)])
    barrier()
    t0 = time.perf_counter()
    h.spmv_host_batch([hxs[i % 3] for i in range(e2e_steps)], [hys[i % 3] for i in range(e2e_steps)])
    torch.cuda.synchronize(dev)
    e2e_local = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_local, e2e_serial_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0].item()) / e2e_steps
    e2e_serial_ms = float(t[1].item()) / e2e_steps

    # BASELINE config 5 (every N, so the scaling record carries it): 100-step power iteration on the skewed matrix as
    # SURVEY 8(d) words it, row slabs over the N GPUs, the iterate exchanged every step
    iterated = None
    if args.workload == "c4" and not args.no_iterated and not args.no_others and not half:
        try:
            iterated = iterated_leg(args, dev, rank, world, local)
        except Exception as e:  # never take the headline number down
            iterated = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    b_alg_total = algorithmic_bytes(m, n, nnz_total, esz)
    # per-rank algorithmic bytes: the slab's CSR, its y, and the part of x the slab READS (its column range), not the
    # whole replicated x (a stencil slab reads 1/P of x plus a halo)
    x_read = (st["col_max"] - st["col_min"] + 1) if st["col_max"] >= st["col_min"] else 0
    b_alg_rank = algorithmic_bytes(r1 - r0, x_read, nnz, esz) if world > 1 else b_alg_total
    ach = b_alg_rank / (ms_local / args.steps * 1e-3) / 1e9
    line = {
        "metric": "spmv_gflops", "value": 2.0 * nnz_total / (ms_step * 1e-3) / 1e9, "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16" if half else "f64", "data": "synthetic",
        "config": config_dict(args, spec, wname, m, n, nnz_total, world,
                              algorithmic_bytes(m // world, n, nnz_total // world, 8), small, args.cold),
        "hbm_gbs": b_alg_total / (ms_step * 1e-3) / 1e9,
        "hbm_frac_of_8tbs": b_alg_total / (ms_step * 1e-3) / 8e12 / world,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": ncu_traffic(args), "kernel": "spmv_kernel (fused, rank 0 slab)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": b_alg_rank},
        "gpu_launches": args.steps * h.launches_per_spmv(),
        "clocks": clk.summary(),
        "e2e": {"value": 2.0 * nnz_total / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
                "h2d_bytes_per_step": x_read * esz, "d2h_bytes_per_step": (r1 - r0) * esz, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "path": "dasp_spmv_host_batch: every step uploads the part of its x this rank's slab reads (columns col_min..col_max) from pinned host memory, runs the fused kernel and downloads its y slab; independent steps pipelined over 3 streams (upload / kernel / download), per rank; bytes are rank 0's",
                "one_blocking_call_per_step": {"value": 2.0 * nnz_total / (e2e_serial_ms * 1e-3) / 1e9, "ms_per_step": e2e_serial_ms,
                                               "path": "dasp_spmv_host (H2D, kernel, D2H back to back, synchronous)"}},
        "preprocess": {"gpu_ms": st["preprocess_ms"], "first_create_in_process_gpu_ms": first_create_ms, "create_wall_s": create_s, "rate_fill0": st["rate_fill0"],
                       "row_long": st["row_long"], "row_block": st["row_block"], "short_rows": st["short_row_1"] + 2 * st["common_13"] + st["short_row_34"] + st["short_row_2"],
                       "device_bytes": st["device_bytes"], "long_gather_lines": st["long_gather_lines"],
                       "long_rows_column_blocked": bool(st["long_blocked"]), "short_rows_banded": bool(st["short_banded"]),
                       "short_band_hit_rate": st["short_band_hit_rate"], "medium_rows_banded": bool(st["medium_banded"]),
                       "medium_band_hit_rate": st["medium_band_hit_rate"], "medium_gather_lines": st["medium_gather_lines"],
                       "launches_per_spmv": h.launches_per_spmv()},
        "parity_check_rel_l2": chk,
    }
    if numa is not None:
        line["e2e"]["host_numa_binding_rank0"] = numa
    if breakdown:
        line["breakdown"] = breakdown
    if iterated is not None:
        line["iterated"] = iterated

    if world == 1 and not args.no_cpu and not half:
        try:
            leg = cpu_leg(args, spec, 1, 3, 1, 60_000_000)
            line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["host_cores_available"] = os.cpu_count()
        except Exception as e:  # the CPU leg must never take the GPU number down with it
            line["cpu_baseline"] = {"error": repr(e)}

    if keep_csr:
        sec = {}
        try:  # cuSPARSE CSR SpMV through torch (bench only, never on the product path)
            A = torch.sparse_csr_tensor(rp, ci, v, size=(m, n))
            for _ in range(3):
                yc = A @ x
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ks = max(3, min(args.steps, 10))
            e0.record()
            for _ in range(ks):
                yc = A @ x
            e1.record()
            torch.cuda.synchronize(dev)
            cms = e0.elapsed_time(e1) / ks
            sec["cusparse_csr"] = {"gflops": 2.0 * nnz_total / (cms * 1e-3) / 1e9, "ms": cms,
                                   "hbm_gbs": b_alg_total / (cms * 1e-3) / 1e9, "via": "torch.sparse_csr @ x"}
            del A, yc
        except Exception as e:
            sec["cusparse_csr"] = {"error": repr(e)}
        del rp, ci, v
        torch.cuda.empty_cache()
        try:
            sec["ref_dasp_sm100a"] = reference_kernels_leg(args, dev)
        except Exception as e:
            sec["ref_dasp_sm100a"] = {"error": repr(e)}
        line["secondary"] = sec
        rk = sec.get("ref_dasp_sm100a", {})
        line["vs_reference_kernels_sm100a"] = {k: {"speedup": e.get("speedup"), "ref_ms": e.get("ref_ms"), "ours_ms": e.get("ours_ms")}
                                               for k, e in rk.items() if isinstance(e, dict) and "speedup" in e}

    h.close()
    if world == 1 and args.workload == "c4" and not args.no_others and not args.no_secondary:
        del x, y
        torch.cuda.empty_cache()
        line["other_configs"] = other_configs(args, dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def power_iteration(args, h, x, cuts, rank, world, dev, spec, wname, nnz_total, nnz_rank, timed, steps=None, exchange=None):
    """x <- A x / ||A x||_2, K steps.  Every rank multiplies its row slab (y written straight into its slab of
    the next x, original row order), the squared norm is all-reduced, the slab is scaled on the device and the
    slabs are exchanged with NCCL broadcasts (slab sizes differ under the nnz-balanced partition)."""
    import torch
    import torch.distributed as dist

    import dasp_b200

    m, n = int(spec.m), int(spec.n)
    assert m == n, "power iteration needs a square matrix"
    steps = steps or args.power_iter
    exchange = exchange or args.exchange
    r0, r1 = cuts[rank], cuts[rank + 1]
    stream = torch.cuda.current_stream(dev).cuda_stream
    # The nnz-balanced slabs have very different row counts (C5: the rank holding the short rows owns almost all of
    # y).  Default exchange: one NCCL broadcast per slab.  Alternative (--exchange a2a): re-cut y into P EQUAL chunks by
    # global row index, all-to-all the slab pieces to the chunk owners, then all-gather the equal chunks.
    chunk = (m + world - 1) // world
    mp = chunk * world
    xa, xb = torch.zeros(mp, dtype=torch.float64, device=dev), torch.zeros(mp, dtype=torch.float64, device=dev)
    xa[:m].copy_(x)
    norm2 = torch.zeros(1, dtype=torch.float64, device=dev)
    esz = 8
    mine = torch.zeros(chunk, dtype=torch.float64, device=dev)

    def overlap(a0, a1, b0, b1):
        return max(0, min(a1, b1) - max(a0, b0))

    send_split = [overlap(r0, r1, j * chunk, min(m, (j + 1) * chunk)) for j in range(world)]
    recv_split = [overlap(cuts[p], cuts[p + 1], rank * chunk, min(m, (rank + 1) * chunk)) for p in range(world)]
    my_len = sum(recv_split)

    perm = exchange == "perm"
    if perm:
        # Relabelled mode: the global permuted index of original row j is (slab offset of its owner) + (that slab's inverse
        # permutation of j); every rank relabels its columns with that map, so the iterate lives in permuted order and the
        # slab product, left as the kernel produces it, IS this rank's contiguous slab of the next iterate.
        order = torch.from_numpy(h.export("order_rid")).to(dev).long()
        gmap = torch.empty(mp, dtype=torch.int32, device=dev)
        gmap[r0 + order] = torch.arange(r0, r1, device=dev, dtype=torch.int32)
        if world > 1:
            for p in range(world):
                if cuts[p + 1] > cuts[p]:
                    dist.broadcast(gmap[cuts[p]:cuts[p + 1]], src=p)
        h.relabel_columns(gmap, m)
        xperm = torch.zeros(mp, dtype=torch.float64, device=dev)
        xperm[gmap[:m].long()] = x
        x = xperm[:m]
        xa[:m].copy_(x)
        del order
    fused = (exchange in ("p2p", "mc", "mcu", "p2pu", "mcc", "p2pc") or perm) and world > 1
    if fused:
        # Fused product + exchange: both iterates live in symmetric memory; every rank's kernel stores its slab of y
        # into ALL copies of the next iterate (peer mappings or one multicast mapping).  The vector is kept
        # un-normalised; the next product scales by 1/sqrt(norm^2), read from device memory, so the only collective
        # per step is the all-reduce of the squared norm, which is also the barrier that orders the peer stores.
        import torch.distributed._symmetric_memory as symm

        sbuf = symm.empty(2 * mp, dtype=torch.float64, device=dev)
        hdl = symm.rendezvous(sbuf, dist.group.WORLD)
        sbuf.zero_()
        xa, xb = sbuf[:mp], sbuf[mp:]
        xa[:m].copy_(x)
        hdl.barrier()
        peer = [int(p) for p in hdl.buffer_ptrs]
        use_mc = exchange in ("mc", "mcu", "mcc") or (perm and int(hdl.multicast_ptr) != 0)
        two_pass = exchange in ("mcu", "p2pu", "mcc", "p2pc")
        copy_pass = exchange in ("mcc", "p2pc")
        yp = torch.zeros(max(r1 - r0, 1), dtype=torch.float64, device=dev) if two_pass else None
        token = torch.zeros(1, dtype=torch.float32, device=dev)
        mc_ptr = int(hdl.multicast_ptr) if use_mc else 0
        if use_mc and mc_ptr == 0:
            raise SystemExit("bench.py: --exchange mc: no multicast support on this system")
        nrm = [torch.ones(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.float64, device=dev)]
        fstate = {"k": 0}

        def dests_of(buf):
            off = (buf.data_ptr() - sbuf.data_ptr())
            if use_mc:
                return [mc_ptr + off]
            # own copy first, then the peers'
            return [peer[rank] + off] + [peer[p] + off for p in range(world) if p != rank]

        def step(src, dst):
            k = fstate["k"]
            if two_pass:
                if copy_pass:
                    # product scattered to original order LOCALLY, norm, then one coalesced scale + copy-to-all pass
                    h.spmv_unpermuted(src, yp, stream)
                    dasp_b200.sumsq(yp, r1 - r0, norm2, stream)
                    dist.all_reduce(norm2)
                    dasp_b200.scale_copy_to(yp, r1 - r0, dests_of(dst), r0, norm2, stream)
                else:
                    # permuted (coalesced) product, norm, then ONE coalesced un-permute + scale + store-to-all pass
                    h.spmv(src, yp, stream)
                    dasp_b200.sumsq(yp, r1 - r0, norm2, stream)
                    dist.all_reduce(norm2)
                    h.unpermute_to(yp, dests_of(dst), r0, norm2, stream)
                dist.all_reduce(token)  # orders every rank's stores into this rank's copy before the next product reads it
                return
            # y_k = A x_k / ||y_{k-1}||  (x_k is stored un-normalised); the first step scales by 1
            if perm:
                h.spmv_permuted_to(src, dests_of(dst), r0, nrm[k & 1], stream)
            else:
                h.spmv_scatter_to(src, dests_of(dst), r0, nrm[k & 1], stream)
            dasp_b200.sumsq(dst.data_ptr() + r0 * esz, r1 - r0, nrm[(k + 1) & 1], stream)
            dist.all_reduce(nrm[(k + 1) & 1])
            norm2.copy_(nrm[(k + 1) & 1])
            fstate["k"] = k + 1
    elif perm:  # one GPU, relabelled mode: the permuted product is the next iterate, scaled by the previous norm on the fly
        nrm = [torch.ones(1, dtype=torch.float64, device=dev), torch.ones(1, dtype=torch.float64, device=dev)]
        fstate = {"k": 0}
        fused = True

        def step(src, dst):
            k = fstate["k"]
            h.spmv_permuted_to(src, [dst], 0, nrm[k & 1], stream)
            dasp_b200.sumsq(dst, m, nrm[(k + 1) & 1], stream)
            norm2.copy_(nrm[(k + 1) & 1])
            fstate["k"] = k + 1
    else:
        def step(src, dst):
            h.spmv_unpermuted(src, dst.data_ptr() + r0 * esz, stream)
            dasp_b200.sumsq(dst.data_ptr() + r0 * esz, r1 - r0, norm2, stream)
            if world > 1:
                dist.all_reduce(norm2)
            dasp_b200.scale_rsqrt(dst.data_ptr() + r0 * esz, r1 - r0, norm2, stream)
            if world == 1:
                return
            if exchange == "bcast":
                works = [dist.broadcast(dst[cuts[p]:cuts[p + 1]], src=p, async_op=True)
                         for p in range(world) if cuts[p + 1] > cuts[p]]
                for w in works:
                    w.wait()
            else:
                dist.all_to_all_single(mine[:my_len], dst[r0:r1], output_split_sizes=recv_split, input_split_sizes=send_split)
                dist.all_gather_into_tensor(dst, mine)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(3, args.warmup)):
        step(xa, xb)
        xa, xb = xb, xa
    xa[:m].copy_(x)
    if fused:
        for t_ in nrm:
            t_.fill_(1.0)
        fstate["k"] = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(xa, xb)
        xa, xb = xb, xa
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    lam = float(torch.sqrt(norm2).item())
    chk = float(xa[:m].sum().item())
    if fused and exchange in ("p2p", "mc", "perm"):  # the iterate is stored un-normalised there
        chk /= lam
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    barrier()
    spmv_ms = torch.tensor([timed(10, 3) / 10], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(spmv_ms, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    step_ms = float(ms.item()) / steps
    return ({
        "metric": "power_iteration_gflops", "value": 2.0 * nnz_total / (step_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
        "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname + f", {steps}-step power iteration", "m": m, "nnz": nnz_total,
                   "partition": "nnz-balanced contiguous row slabs, x replicated",
                   "exchange": ("none (single GPU)" if world == 1 else
                                "relabelled P*A*P^T mode: the SpMV kernel stores its contiguous slab of the permuted product into every rank's next iterate through one NVSwitch multicast mapping (peer mappings without multicast); all_reduce(norm^2) is the only collective and the barrier" if exchange == "perm" else
                                "all_reduce(norm^2) + one NCCL broadcast per non-empty slab per step" if exchange == "bcast" else
                                "fused: SpMV kernel stores y into every peer copy over NVLink P2P mappings; all_reduce(norm^2) only" if exchange == "p2p" else
                                "fused: SpMV kernel stores y through one NVSwitch multicast mapping; all_reduce(norm^2) only" if exchange == "mc" else
                                "permuted SpMV + all_reduce(norm^2) + one coalesced un-permute/scale pass storing through an NVSwitch multicast mapping (dasp_unpermute_to) + token all_reduce" if exchange == "mcu" else
                                "permuted SpMV + all_reduce(norm^2) + one coalesced un-permute/scale pass storing to every peer mapping (dasp_unpermute_to) + token all_reduce" if exchange == "p2pu" else
                                "SpMV into a local original-order slab + all_reduce(norm^2) + one coalesced scale/copy pass storing through an NVSwitch multicast mapping (dasp_scale_copy_to) + token all_reduce" if exchange == "mcc" else
                                "SpMV into a local original-order slab + all_reduce(norm^2) + one coalesced scale/copy pass storing to every peer mapping (dasp_scale_copy_to) + token all_reduce" if exchange == "p2pc" else
                                "all_reduce(norm^2) + NCCL all_to_all of slab pieces into P equal chunks + all_gather of the chunks"),
                   "slab_rows": [cuts[p + 1] - cuts[p] for p in range(world)]},
        "spmv_only_ms": float(spmv_ms.item()), "exchange_and_vector_ms": step_ms - float(spmv_ms.item()),
        "eigenvalue_estimate": lam, "x_checksum": chk,
        "gpu_launches": steps * (h.launches_per_spmv() + 3),
    })


def power_iteration_hybrid(args, spec, wname, cuts, rank, world, dev, local, x0, steps, warmup, nnz_total, rp=None, ci=None, v=None):
    """The iterated workload with a 1.5-D partition: every rank owns a row slab AND the matching column slab of x.  The
    SHORT / MEDIUM rows of the slab are multiplied locally (they need x of the slab plus a halo from the neighbouring
    slabs); the LONG rows are split by COLUMNS — rank p multiplies, for every long row of the matrix, the entries whose
    columns lie in its slab — and their partial sums are merged by ONE all-reduce of row_long + 1 doubles per step (the
    squared norm of the short part rides in the same buffer).  This is DASP's split-row reduction of long rows carried across
    GPUs: no rank ever needs the whole iterate, the per-step traffic is 8 KB + the halos instead of an all-gather of x."""
    import torch
    import torch.distributed as dist

    import dasp_b200
    from dasp_b200 import partition, synth

    m = int(spec.m)
    r0, r1 = cuts[rank], cuts[rank + 1]
    rows = r1 - r0
    stream = torch.cuda.current_stream(dev).cuda_stream
    if rp is None:
        rp, ci, v, _ = synth.generate(spec, r0, r1, dev)
    lens = (rp[1:] - rp[:-1]).long()
    long_local = torch.nonzero(lens >= 256).flatten()  # block_longest
    # global list of long rows (ascending: slabs are contiguous and ordered)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = long_local.numel()
    if world > 1:
        dist.all_reduce(counts)
    nl = int(counts.sum().item())
    off = int(counts[:rank].sum().item())
    glong = torch.zeros(max(nl, 1), dtype=torch.int64, device=dev)
    glong[off:off + long_local.numel()] = long_local + r0
    if world > 1:
        dist.all_reduce(glong)
    glong = glong[:nl]
    # column pieces of ALL long rows that fall into this rank's slab [r0, r1)
    pieces = []
    for g in glong.tolist():
        _, gc, gv, _ = synth.generate(spec, g, g + 1, dev)
        keep = (gc >= r0) & (gc < r1)
        pieces.append((gc[keep], gv[keep]))
    # local matrix: the slab's rows with the long rows emptied, then one row per long-row piece (dasp_b200/partition.py,
    # the same code the two-rank gloo test runs on the CPU)
    rp_l, ci_l, v_l, cmin, cmax = partition.hybrid_local_matrix(rp, ci, v, r0, r1, long_local, pieces)
    nnz_l = int(rp_l[-1].item())
    del pieces, rp, ci, v
    torch.cuda.empty_cache()
    h = dasp_b200.Dasp(dasp_b200.DASP_F64, rows + nl, m, rp_l.to(torch.int32), ci_l.to(torch.int32), v_l, device=local, nnz=nnz_l)
    del rp_l, ci_l, v_l
    torch.cuda.empty_cache()
    # halo plan: rank q needs columns [cmin_q, cmax_q] of x; the owner of each part outside q's own slab sends it
    need = torch.zeros(world, 2, dtype=torch.int64, device=dev)
    need[rank, 0], need[rank, 1] = cmin, cmax + 1
    if world > 1:
        dist.all_reduce(need)
    need = need.tolist()
    sends, recvs = partition.halo_plan(need, cuts, rank)  # (peer, lo, hi) in global indices
    x = torch.zeros(m, dtype=torch.float64, device=dev)
    y = torch.zeros(rows + nl, dtype=torch.float64, device=dev)
    red = torch.zeros(nl + 1, dtype=torch.float64, device=dev)
    mine = (glong >= r0) & (glong < r1)
    my_pos = torch.nonzero(mine).flatten()           # positions in the long list of the long rows this rank owns
    my_rows = glong[my_pos] - r0                      # their local row index
    norm2 = torch.zeros(1, dtype=torch.float64, device=dev)

    def halo():
        if not sends and not recvs:
            return
        ops = [dist.P2POp(dist.isend, x[lo:hi], q) for q, lo, hi in sends] + [dist.P2POp(dist.irecv, x[lo:hi], q) for q, lo, hi in recvs]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def step():
        h.spmv_unpermuted(x, y, stream)                       # slab rows (long rows come out 0) + partial sums of the long pieces
        dasp_b200.sumsq(y, rows, red.data_ptr() + 8 * nl, stream)  # squared norm of the short part, next to the partials
        red[:nl].copy_(y[rows:])
        if world > 1:
            dist.all_reduce(red)                              # the ONE collective: row_long + 1 doubles
        norm2.copy_(red[nl:] + (red[:nl] * red[:nl]).sum())
        y[my_rows] = red[my_pos]                              # the long rows this rank owns, complete
        torch.mul(y[:rows], torch.rsqrt(norm2), out=x[r0:r1])  # next iterate, own slab
        halo()

    def reset():
        x.zero_()
        x[r0:r1].copy_(x0[r0:r1])
        lo, hi = need[rank]
        x[lo:hi].copy_(x0[lo:hi])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    reset()
    for _ in range(max(3, warmup)):
        step()
    reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    lam = float(torch.sqrt(norm2).item())
    chk = x[r0:r1].sum().reshape(1).clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(chk)
    barrier()
    sp = torch.tensor([h.spmv_timed(x, y, stream, 3, 10) / 10], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(sp, op=dist.ReduceOp.MAX)
    launches = h.launches_per_spmv()
    h.close()
    if rank != 0:
        return None
    step_ms = float(ms.item()) / steps
    halo_bytes = sum((hi - lo) * 8 for _, lo, hi in recvs)
    return {"metric": "power_iteration_gflops", "value": 2.0 * nnz_total / (step_ms * 1e-3) / 1e9, "unit": "GFLOP/s",
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "n_gpus": world, "steps": steps, "ms_per_step": step_ms, "spmv_only_ms": float(sp.item()),
            "exchange_and_vector_ms": step_ms - float(sp.item()), "eigenvalue_estimate": lam, "x_checksum": float(chk.item()),
            "long_rows": nl, "allreduce_bytes_per_step": (nl + 1) * 8, "halo_bytes_received_per_step_rank0": halo_bytes,
            "gpu_launches": steps * (launches + 1),
            "config": {"workload": wname + f", {steps}-step power iteration", "m": m,
                       "exchange": "1.5-D partition: short / medium rows by row slab with a halo exchange of x (NCCL send/recv with the neighbouring slabs), long rows split by COLUMNS across the ranks and merged by one all-reduce of row_long + 1 doubles per step; no rank holds the whole iterate",
                       "slab_rows": [cuts[p + 1] - cuts[p] for p in range(world)]}}


def iterated_leg(args, dev, rank, world, local):
    """C5 (skewed_spec, full size) as a 100-step power iteration x <- A x / ||A x|| on `world` GPUs: nnz-balanced contiguous
    row slabs (the long rows sit at seeded positions, so the slabs are balanced in rows as well), iterate replicated,
    exchanged every step.  Returns a compact dict: per-step time of the fused relabelled exchange and (N > 1) of the plain
    NCCL broadcast exchange, SpMV-only time, eigenvalue estimate and checksum (identical for every N and exchange)."""
    import copy

    import torch

    import dasp_b200
    from dasp_b200 import synth

    a = copy.copy(args)
    a.workload, a.scale, a.power_iter = "c5_spec", 1.0, 100
    spec, wname, _ = make_spec(a)
    m = int(spec.m)
    ln = synth.row_lengths(spec, 0, m, dev)
    cs = torch.cumsum(ln, 0)
    nnz_total = int(cs[-1].item())
    targets = torch.tensor([nnz_total * p // world for p in range(1, world)], device=dev, dtype=torch.int64)
    cuts = [0] + [min(int(c) + 1, m) for c in torch.searchsorted(cs, targets, right=False).tolist()] + [m]
    del ln, cs
    r0, r1 = cuts[rank], cuts[rank + 1]
    rp, ci, v, nnz = synth.generate(spec, r0, r1, dev)
    out = {"workload": wname + ", 100-step power iteration", "m": m, "nnz": nnz_total, "n_gpus": world,
           "slab_rows": [cuts[p + 1] - cuts[p] for p in range(world)]}
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    x = torch.rand(m, generator=gen, device=dev, dtype=torch.float64) * 2 - 1
    y = torch.zeros(max(r1 - r0, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    for exchange in (("perm", "bcast") if world > 1 else ("perm",)):
        h = dasp_b200.Dasp(dasp_b200.DASP_F64, r1 - r0, m, rp, ci, v, device=local, nnz=nnz)

        def timed(k, w, h=h):
            for _ in range(w):
                h.spmv(x, y, stream)
            torch.cuda.synchronize(dev)
            return h.spmv_timed(x, y, stream, 0, k)

        res = power_iteration(a, h, x, cuts, rank, world, dev, spec, wname, nnz_total, nnz, timed, steps=100, exchange=exchange)
        h.close()
        if res is not None:
            key = "fused_relabelled" if exchange == "perm" else "nccl_broadcast"
            out[key] = {"ms_per_step": res["ms_per_step"], "gflops": res["value"], "spmv_only_ms": res["spmv_only_ms"],
                        "exchange_and_vector_ms": res["exchange_and_vector_ms"], "eigenvalue_estimate": res["eigenvalue_estimate"],
                        "x_checksum": res["x_checksum"], "exchange": res["config"]["exchange"]}
    res = power_iteration_hybrid(a, spec, wname, cuts, rank, world, dev, local, x, 100, 3, nnz_total, rp, ci, v)
    if res is not None:
        out["hybrid_1p5d"] = {k: res[k] for k in ("ms_per_step", "spmv_only_ms", "exchange_and_vector_ms", "eigenvalue_estimate", "x_checksum",
                                                  "long_rows", "allreduce_bytes_per_step", "halo_bytes_received_per_step_rank0")}
        out["hybrid_1p5d"]["gflops"] = res["value"]
        out["hybrid_1p5d"]["exchange"] = res["config"]["exchange"]
    del rp, ci, v, x, y
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    best = min((out[k] for k in ("fused_relabelled", "nccl_broadcast", "hybrid_1p5d") if k in out), key=lambda e: e["ms_per_step"])
    out["best"] = [k for k in ("fused_relabelled", "nccl_broadcast", "hybrid_1p5d") if k in out and out[k] is best][0]
    out.update({"steps": 100, "ms_per_step": best["ms_per_step"], "gflops": best["gflops"],
                "limiter": ("one GPU: the product itself" if world == 1 else
                            "all-gather style exchanges (fused_relabelled, nccl_broadcast): every GPU must RECEIVE the other slabs of the iterate, %.0f MB per step, over NVLink (measured ~460 GB/s of ingress under 8-way multicast = 0.76 ms, twice the product); "
                            "hybrid_1p5d removes that volume (long rows split by columns: %d bytes all-reduced + halos), its step is the slab product plus ~10 small launches and one 8 KB all-reduce"
                            % ((m - (cuts[1] - cuts[0])) * 8 / 1e6, 8008))})
    return out


def other_configs(args, dev):
    """Short measurements of the remaining BASELINE.json configurations (1 GPU), same protocol as the headline:
    generate on the device, dasp_create, parity check of ALL rows against an independent device-side evaluation, K
    back-to-back launches timed with CUDA events from C.  C3 and C5 are reported on BOTH generators: `c3_spec` / `c5_spec`
    follow SURVEY.md 8(d) literally (unsorted columns; C5 long rows uniform over all n), `c3` / `c5` are the CSR-sorted
    variants (ascending windowed columns; C5 long rows in a common 2^21 band).  The L2-resident C1 / C2 are reported warm
    (reference protocol) AND cold (L2 flushed before every launch)."""
    import copy

    import torch

    import dasp_b200
    from dasp_b200 import synth

    peak, _ = measured_peak_gbs()
    out = {}
    for w in ("c1", "c2", "c3_spec", "c3", "c5_spec", "c5"):
        a = copy.copy(args)
        a.workload, a.scale = w, 1.0
        try:
            spec, wname, half = make_spec(a)
            esz = 2 if half else 8
            tdt = torch.float16 if half else torch.float64
            m, n = int(spec.m), int(spec.n)
            rp, ci, v, nnz = synth.generate(spec, 0, m, dev, half=half)
            h = dasp_b200.Dasp(dasp_b200.DASP_F16 if half else dasp_b200.DASP_F64, m, n, rp, ci, v, nnz=nnz)
            st = h.stats()
            gen = torch.Generator(device=dev)
            gen.manual_seed(7)
            x = (torch.rand(n, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(tdt)
            y = torch.zeros(m, dtype=tdt, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            h.spmv_unpermuted(x, y, stream)
            torch.cuda.synchronize(dev)
            err = device_check(rp, ci, v, x, y, m)
            if err > (2e-3 if half else 1e-12):
                raise RuntimeError(f"parity check failed: {err}")
            del rp, ci, v
            torch.cuda.empty_cache()
            b = algorithmic_bytes(m, n, nnz, esz)
            small = b < 256e6
            steps, warm = (2000, 200) if small else (20, 5)
            ms = h.spmv_timed(x, y, stream, warm, steps) / steps

            def entry(ms, l2):
                return {"ms_per_step": ms, "gflops": 2.0 * nnz / (ms * 1e-3) / 1e9, "hbm_gbs": b / (ms * 1e-3) / 1e9,
                        "hbm_frac_of_8tbs": b / (ms * 1e-3) / 8e12, "roofline_frac_of_measured_peak": b / (ms * 1e-3) / 1e9 / peak,
                        "l2": l2}

            out[w] = {"workload": wname, "m": m, "nnz": nnz, "dtype": "f16" if half else "f64", "steps": steps,
                      "parity_check_rel_l2_all_rows": err, "launches_per_spmv": h.launches_per_spmv(),
                      "preprocess_gpu_ms": st["preprocess_ms"]}
            out[w].update(entry(ms, "warm L2, back-to-back launches (reference protocol)" if small else "inputs exceed L2"))
            if small:  # cold: L2 flushed (512 MB written) before every timed launch
                flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.int32, device=dev)
                tot, k = 0.0, 200
                for _ in range(k):
                    synth.flush_l2(flush)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    h.spmv(x, y, stream)
                    e1.record()
                    torch.cuda.synchronize(dev)
                    tot += e0.elapsed_time(e1)
                out[w]["cold"] = entry(tot / k, "L2 flushed before every launch, one launch per event pair")
                del flush
            h.close()
            del x, y
            torch.cuda.empty_cache()
        except Exception as e:
            out[w] = {"error": repr(e)}
    return out


def reference_kernels_leg(args, dev):
    """The reference's own kernels (unmodified, recompiled for sm_100a: oracle/_ref) and ours on the same bounded
    samples: a 128^3 stencil (the reference preprocesses on one host thread) and the cop20k_A stand-in in FP64 (C1)
    and FP16 (C2).  The reference times itself (100 warm-up + 1000 launches, gettimeofday, src/dasp_f64.h:1285-1320);
    ours is timed with the same protocol from C (dasp_spmv_timed)."""
    import torch

    import dasp_b200
    import oracle
    from dasp_b200 import synth

    if not (oracle.ref_available(oracle.F64) and oracle.ref_available(oracle.F16)):
        return {"unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
    out = {}
    g = min(args.grid, 128)
    samples = [("stencil", synth.stencil27(g), False, f"27-point stencil {g}^3 fp64")]
    if args.grid > 128 and args.workload == "c4":  # the headline shape itself: ~7 s of single-threaded reference preprocessing
        samples.append(("c4", synth.stencil27(args.grid), False, f"27-point stencil {args.grid}^3 fp64 (C4)"))
    samples += [
("c1", synth.banded(), False, "cop20k_A stand-in fp64"),
                ("c2", synth.banded(), True, "cop20k_A stand-in fp16")]
    s = torch.cuda.current_stream(dev).cuda_stream
    for key, spec, half, label in samples:
        try:
            m = int(spec.m)
            dt = oracle.F16 if half else oracle.F64
            tdt = torch.float16 if half else torch.float64
            rp, ci, v, nnz = synth.generate(spec, 0, m, dev, half=half)
            rp_h, ci_h, v_h = rp.cpu().numpy(), ci.cpu().numpy(), v.cpu().numpy()
            h = dasp_b200.Dasp(dasp_b200.DASP_F16 if half else dasp_b200.DASP_F64, m, m, rp, ci, v, nnz=nnz)
            x = torch.ones(m, dtype=tdt, device=dev)
            y = torch.empty(m, dtype=tdt, device=dev)
            ours_ms = h.spmv_timed(x, y, s, 100, 1000) / 1000
            h.close()
            del rp, ci, v
            xs = np.ones(m, dtype=np.float16 if half else np.float64)
            r = oracle.ref_spmv_all(dt, m, m, rp_h, ci_h, v_h, x=xs)
            cols = r["csv"].split(",")
            ref_ms = float(cols[22 if half else 21])  # dasp_time of the CSV record (src/dasp_f64.h:1441, src/dasp_f16.h:1758)
            f = oracle.csr_spmv_f16 if half else oracle.csr_spmv_f64
            y_ref = f(m, rp_h, ci_h, v_h, xs)
            ref_err = (float(np.linalg.norm(r["y_perm"].astype(np.float64) - y_ref[r["order_rid"]]) / np.linalg.norm(y_ref))
                       if r["ran_on_gpu"] else None)
            out[key] = {"sample": f"{label} ({nnz} nnz); both timed as 100 warm-up + 1000 back-to-back launches",
                        "ref_ms": ref_ms, "ref_gflops": 2.0 * nnz / (ref_ms * 1e-3) / 1e9 if ref_ms > 0 else None,
                        "ours_ms": ours_ms, "ours_gflops": 2.0 * nnz / (ours_ms * 1e-3) / 1e9,
                        "speedup": ref_ms / ours_ms if ours_ms > 0 else None,
                        "ran_on_gpu": r["ran_on_gpu"], "ref_y_rel_l2_vs_serial_csr": ref_err}
        except Exception as e:
            out[key] = {"error": repr(e)}
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
