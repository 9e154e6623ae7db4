#!/usr/bin/env python
"""bench.py -- the JaggedArray hot-path sweep on B200 (BASELINE.json: "jagged elements/sec and HBM GB/s vs peak
for sum/argmax/mask/pairs at 1/2/4/8 GPUs").

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference itself (awkward0, oracle/_ref), host cores

Workload (config.workload): BASELINE.json configs[4] shard -- per GPU 125 M events (1 B events at 8 GPUs),
Poisson(5) counts, float32 content Exp(25); one STEP = counts2offsets scan + segmented sum + argmax +
(pt > 20) mask -> select -> flatten/counts, then the whole-array totals (NCCL all-reduce when N > 1).  Weak
scaling: every rank owns the same number of events, no data-path collective.

value  = content elements swept per second with inputs resident in HBM (CUDA events, max over ranks)
e2e    = the same sweep through the public API from pinned HOST buffers (H2D inside the timed region, totals read back)
roofline = dominant kernel: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json hbm_gbs
cpu_baseline = the UNMODIFIED reference (oracle/_ref, see oracle/fetch_ref.py; the numpy oracle port only if that copy
               is absent) on one host core, bounded sample; `--impl reference` runs it on all host cores (one process
               per core, equal shards); every API workflow also carries the reference's time on a stated sample
"""
import argparse
import re
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "awkward-0.x_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

THRESHOLD = 20.0
FALLBACK_HBM_GBS = 6650.0


# ----------------------------------------------------------------------------------------------------------------
# synthetic data
# ----------------------------------------------------------------------------------------------------------------
def make_shard(nevents, seed, mean=5.0, scale=25.0):
    rng = np.random.default_rng(seed)
    counts = rng.poisson(mean, nevents).astype(np.int64)
    total = int(counts.sum())
    pt = rng.standard_exponential(total, dtype=np.float32)
    pt *= np.float32(scale)
    return counts, pt


# ----------------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (numpy oracle port) on all host cores
# ----------------------------------------------------------------------------------------------------------------
_W = {}


def _worker_init(nevents, seed_base):
    import multiprocessing
    ident = multiprocessing.current_process()._identity
    seed = seed_base + (ident[0] if ident else 0)
    _W["data"] = make_shard(nevents, seed)


def reference_module():
    """The UNMODIFIED reference (awkward0): /root/reference in the build container, else the byte-identical copy
    oracle/_ref that oracle/fetch_ref.py made and the gpurun snapshot carried (None when neither exists)."""
    if "mod" not in _REF:
        from oracle import _refloader
        _REF["mod"] = _refloader.load() if _refloader.available() else None
    return _REF["mod"]


_REF = {}


def cpu_kind():
    return "reference" if reference_module() is not None else "port"


def cpu_sweep(counts, pt):
    """The same sweep on the CPU: through the reference's own API when the reference is here, else with the numpy
    oracle port of its algorithm.  Returns (elements, checksum tuple)."""
    ref = reference_module()
    if ref is not None:
        a = ref.JaggedArray.fromcounts(counts, pt)
        s = a.sum()
        lead = a.argmax()
        sel = a[a > THRESHOLD]
        flat = sel.flatten()
        c = sel.counts
        return len(pt), (float(s.sum(dtype=np.float64)), int(len(flat)), int(lead.counts.sum()), int(c.sum()))
    from oracle import jagged_oracle as O
    off = O.counts2offsets(counts)
    s = O.reduce(off, pt, "sum")
    aoff, aidx = O.argminmax(off, pt, False)
    moff, mask = O.ufunc_broadcast(np.greater, off, pt, np.float32(THRESHOLD))
    soff, sel = O.mask_select(off, pt, moff, mask)
    flat = O.flatten(soff, sel)
    c = O.offsets2counts(soff)
    return len(pt), (float(s.sum(dtype=np.float64)), int(len(flat)), int(len(aidx)), int(c.sum()))


def _worker_step(_):
    counts, pt = _W["data"]
    return cpu_sweep(counts, pt)[0]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_worker = args.ref_events_per_core
    ctx = multiprocessing.get_context("fork")
    pool = ctx.Pool(cores, initializer=_worker_init, initargs=(per_worker, 977))
    try:
        for _ in range(args.warmup):
            pool.map(_worker_step, range(cores), chunksize=1)
        t0 = time.perf_counter()
        elements = 0
        for _ in range(args.steps):
            elements += sum(pool.map(_worker_step, range(cores), chunksize=1))
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    value = elements / dt
    line = {
        "impl": "reference", "metric": "jagged elements/sec (offsets scan + sum + argmax + mask sweep)",
        "value": value, "unit": "elements/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "events_per_step": per_worker * cores, "threshold": THRESHOLD},
        "cpu_baseline": {"value": value, "unit": "elements/s", "cores": cores, "kind": cpu_kind(),
                         "sample": "%d events per core x %d cores per step (Poisson(5), f32), one process per core on equal "
                                   "event shards (the reference's ChunkedArray decomposition); %s"
                                   % (per_worker, cores, CPU_KIND_NOTE[cpu_kind()])},
        "e2e": {"value": value, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


CPU_KIND_NOTE = {
    "reference": "the UNMODIFIED reference awkward0 (oracle/_ref, copied by oracle/fetch_ref.py) through its own API",
    "port": "numpy oracle port of awkward0/array/jagged.py (oracle/_ref is absent on this box)",
}

WORKLOAD_NAME = ("C5 sweep shard: 125M events/GPU (1B events at 8 GPUs), Poisson(5) counts, float32 Exp(25) content; "
                 "step = counts2offsets + sum + argmax + (pt>20) mask-select + flatten/counts + global totals")


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import awkward0_b200 as ak
    from awkward0_b200 import _lib, sharded
    from awkward0_b200.device import DeviceArray, empty, runtime

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    binding = bind_host_to_gpu(local_rank)           # before any pinned allocation: first touch on the GPU's NUMA node
    strong = args.total_events > 0
    if strong:
        lo, hi = sharded.event_range(args.total_events, rank, world)     # BASELINE configs[4]: 1 B events over 1/2/4/8 GPUs
        nev = hi - lo
    else:
        nev = args.events_per_gpu
    device_rng = strong or args.device_rng
    if device_rng:
        # drawn in HBM (a 1 B-event shard would be a 28 GB host draw): same distributions, torch's generators
        g = torch.Generator(device=dev)
        g.manual_seed(12345 + rank)
        counts_t = torch.poisson(torch.full((nev,), 5.0, device=dev), generator=g).to(torch.int64)
        M = int(counts_t.sum())
        pt_t = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
        counts_dev, pt_dev = ak.DeviceArray(counts_t), ak.DeviceArray(pt_t)
        counts_pin = pt_pin = None
        del counts_t, pt_t
    else:
        counts_np, pt_np = make_shard(nev, 12345 + rank)
        M = len(pt_np)
        # pinned host staging (the e2e leg copies from here every step)
        counts_pin = torch.empty(nev, dtype=torch.int64, pin_memory=True)
        counts_pin.numpy()[:] = counts_np
        pt_pin = torch.empty(M, dtype=torch.float32, pin_memory=True)
        pt_pin.numpy()[:] = pt_np
        del counts_np, pt_np
        counts_dev = ak.asdevice(counts_pin)
        pt_dev = ak.asdevice(pt_pin)
    torch.cuda.synchronize()

    def sweep(counts, content):
        """One step through the public API.  Returns the whole-array totals (python numbers).  The partial totals stay in
        HBM where the kernels wrote them (the sum of sums from jg_flat_reduce, K and the number of leading objects as
        the last entry of the results' offsets): ONE collective and ONE device-to-host copy per step."""
        a = ak.JaggedArray.fromcounts(counts, content)          # scan (+ total/negative-count read-back)
        s = a.sum()
        acc = sharded.DeviceTotals()
        acc.add_sum(s.device_sum())                              # queued behind the reduction: no host round trip
        # the selection hands K back as soon as its counting pass has run (the compaction is still in flight), so the
        # arg-reduction is queued behind it without the device going idle; the arg-reduction needs the whole pass
        sel = a[a > THRESHOLD]
        lead = a.argmax()
        flat = sel.flatten()
        c = sel.counts
        acc.add_sum(sel.offsets[-1:])
        acc.add_sum(lead.offsets[-1:])
        gsum, gk, glead = acc.finish()
        assert len(flat) <= gk and len(c) == len(a)
        return gsum, gk, glead, len(c)

    def sweep_streamed(counts_host, content_host):
        """The same step from HOST buffers: the copy is queued once on a side stream and the sweep runs event chunk
        by event chunk on views of the resident array while later pieces are still on the bus (ingest.py)."""
        stream = ak.StreamedJaggedArray(counts_host, content_host, nchunks=args.e2e_chunks)   # H2D + scan
        acc = sharded.DeviceTotals(capacity=3 * max(args.e2e_chunks, 1))
        nev = 0
        for a in stream:
            s = a.sum()
            acc.add_sum(s.device_sum())
            sel = a[a > THRESHOLD]
            lead = a.argmax()
            flat = sel.flatten()
            c = sel.counts
            acc.add_sum(sel.offsets[-1:])
            acc.add_sum(lead.offsets[-1:])
            nev += len(c)
        parts = acc.finish()
        return sum(parts[0::3]), sum(parts[1::3]), sum(parts[2::3]), nev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms, out

    # ---- parity of the very step that is timed, at every N: every rank sweeps the same fixed 1 M-event shard through
    # the same code path (collective included) and the totals must equal the CPU arm's (reference / oracle) x world
    check_counts, check_pt = make_shard(1_000_000, 4242)
    _, (want_sum, want_k, want_lead, _) = cpu_sweep(check_counts, check_pt)
    got_sum, got_k, got_lead, got_n = sweep(ak.asdevice(check_counts), ak.asdevice(check_pt))
    totals_ok = (abs(got_sum - world * want_sum) <= 1e-5 * world * abs(want_sum) and got_k == world * want_k
                 and got_lead == world * want_lead and got_n == len(check_counts))
    if world > 1:
        flag = torch.tensor([1 if totals_ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        totals_ok = bool(flag.item())
    if not totals_ok:
        raise SystemExit("totals_check FAILED on rank %d: got %r, want %r x %d" % (
            rank, (got_sum, got_k, got_lead), (want_sum, want_k, want_lead), world))
    if world > 1:
        # a variable-size result replicated whole over NCCL (counts and content gathered separately, sizes first)
        chk = ak.JaggedArray.fromcounts(ak.asdevice(check_counts), ak.asdevice(check_pt))
        picked = chk[chk > THRESHOLD]
        gc, gt = sharded.gather_jagged(picked.counts, picked.flatten())
        assert gc.numel() == world * len(check_counts) and gt.numel() == world * want_k, (gc.numel(), gt.numel())
        assert int(gc.sum().item()) == world * want_k
        del chk, picked, gc, gt
    del check_counts, check_pt

    # ---- warm-up, then the device-resident timed region
    for _ in range(max(args.warmup, 3)):
        totals = sweep(counts_dev, pt_dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.LAUNCHES[0] = 0
    ms_dev, totals = timed(lambda: sweep(counts_dev, pt_dev), args.steps)
    launches = _lib.LAUNCHES[0]

    # ---- e2e: pinned host -> device inside the timed region, totals read back
    ms_e2e, h2d = None, None
    if counts_pin is not None:
        e2e_step = (lambda: sweep_streamed(counts_pin, pt_pin)) if args.e2e_chunks > 0 else (lambda: sweep(counts_pin, pt_pin))
        for _ in range(1):
            e2e_step()
        ms_e2e, totals_e2e = timed(e2e_step, args.steps)
        assert totals_e2e[1] == totals[1] and totals_e2e[2] == totals[2], (totals_e2e, totals)
        # the ceiling of that leg: the same bytes copied from the same pinned buffers with NO compute, all ranks at once
        stage_c, stage_p = torch.empty_like(counts_dev.tensor), torch.empty_like(pt_dev.tensor)

        def copy_only():
            stage_c.copy_(counts_pin, non_blocking=True)
            stage_p.copy_(pt_pin, non_blocking=True)
        copy_only()
        ms_copy, _ = timed(copy_only, max(2, args.steps // 4))
        ms_copy /= max(2, args.steps // 4)
        mine = torch.tensor([(nev * 8 + M * 4) / (ms_copy * 1e-3) / 1e9], dtype=torch.float64, device=dev)
        every = torch.empty(world, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_gather_into_tensor(every, mine)
        else:
            every = mine
        h2d = {"copy_only_ms": ms_copy, "gbs_per_rank": [round(float(x), 2) for x in every.cpu().tolist()]}
        del stage_c, stage_p
    clocks = sampler.stop() if rank == 0 else None

    # ---- result gather over NVLink: the per-event sums of every shard replicated on every rank
    gather_row = None
    if world > 1:
        a_full = ak.JaggedArray.fromcounts(counts_dev, pt_dev)
        per_event = a_full.sum()
        sizes = [sharded.event_range(args.total_events, r, world)[1] - sharded.event_range(args.total_events, r, world)[0]
                 for r in range(world)] if strong else [nev] * world
        sharded.gather_per_event(per_event, sizes=sizes)
        ms_g, gathered = timed(lambda: sharded.gather_per_event(per_event, sizes=sizes), 5)
        ms_g /= 5
        recv = (sum(sizes) - nev) * 4
        gather_row = {"ms": ms_g, "bytes_received_per_rank": recv, "gbs_per_rank": recv / (ms_g * 1e-3) / 1e9,
                      "elements": sum(sizes), "collective": "ncclAllGather (all_gather_into_tensor, one call, no staging copy)"}
        assert gathered.numel() == sum(sizes)
        del a_full, per_event, gathered

    Mt = torch.tensor([M], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(Mt)
    M_all = float(Mt.item())
    value = M_all * args.steps / (ms_dev * 1e-3)
    e2e_value = M_all * args.steps / (ms_e2e * 1e-3) if ms_e2e else None
    Nt = torch.tensor([nev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(Nt)
    N_all = int(Nt.item())

    # ---- per-kernel roofline table (rank 0, device-resident, CUDA events around each C-ABI call)
    ops = {}
    if rank == 0 and not args.no_tables:
        ops = kernel_table(ak, _lib, DeviceArray, empty, runtime, counts_dev, pt_dev, nev, M, args)
    peaks = load_peaks()
    roof = None
    if ops:
        step_ops = ["counts2offsets_tiles", "segreduce_sum_f32", "segargmax_dense_tiles_f32", "compare_select_f32"]       # a[a > 20] is one fused op; it writes its counts itself
        dom = max(step_ops, key=lambda k: ops[k]["ms"])
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
                traffic = float(sum(json.load(f)[dom]["kernels"].values()))
        except Exception:
            traffic = None
        roof = {"bound": "hbm", "kernel": dom, "achieved": ops[dom]["gbs"], "peak": peaks[0], "unit": "GB/s",
                "frac": ops[dom]["gbs"] / peaks[0], "traffic": traffic, "peak_source": peaks[1],
                "traffic_source": "ncu --set full dram bytes of the op's kernels at 125 M events, profiles/r2_traffic.json",
                "algorithmic_bytes": ops[dom]["bytes"], "ms": ops[dom]["ms"],
                "step_share": ops[dom]["ms"] / sum(ops[k]["ms"] for k in step_ops)}

    workflows = {}
    if rank == 0 and not args.no_tables:
        workflows = workflow_table(ak, counts_dev, pt_dev, nev, M, args)
        ops.update(distribution_table(ak, args))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args)
        if workflows:
            # the reference's time for the same API chains (per-unit rates are comparable: the CPU sample is smaller)
            for name, row in cpu_workflows(args).items():
                if name in workflows:
                    workflows[name].update(row)
                    unit = [k for k in row if k.endswith("_per_s")][0][4:]
                    if unit in workflows[name]:
                        workflows[name]["speedup_vs_cpu_1core"] = workflows[name][unit] / row["cpu_" + unit]

    if rank == 0:
        line = {
            "metric": "jagged elements/sec (offsets scan + sum + argmax + mask sweep)", "value": value,
            "unit": "elements/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "events_per_gpu": nev, "content_elements_per_gpu": M,
                       "total_events": N_all, "total_content_elements": int(M_all),
                       "threshold": THRESHOLD, "l2": "inputs (%.1f GB/GPU) are larger than L2; no flush needed" % ((nev * 8 + M * 4) / 1e9),
                       "rng": "torch generators on the device" if device_rng else "numpy default_rng(12345 + rank) on the host (SURVEY 8d)",
                       "parallelism": "event-range shards, %d rank(s)" % world, "host_binding": binding},
            "e2e": ({"value": e2e_value, "unit": "elements/s", "ms_per_step": ms_e2e / args.steps,
                     "h2d_bytes_per_step": int(nev * 8 + M * 4),
                     # the scan's 64-byte status / total word, per chunk those of argmax and the selection, the chunk
                     # boundaries, and ONE buffer with every partial total (3 x 8 bytes per chunk)
                     "d2h_bytes_per_step": (64 + max(args.e2e_chunks, 1) * (2 * 64 + 24) + (max(args.e2e_chunks, 1) + 1) * 8) if args.e2e_chunks > 0 else 3 * 64 + 24,
                     "h2d": h2d,
                     "note": ("public API from pinned host buffers (StreamedJaggedArray: ONE copy of counts + content per step on a side stream, "
                              "the sweep runs on %d event-range views while later pieces are still on the bus); per-event results stay in "
                              "HBM (none returns to the host inside the timed region), only the whole-array totals are read back; "
                              "h2d.copy_only_ms is the same copy with no compute, all ranks at once: the ceiling of this leg"
                              % args.e2e_chunks) if args.e2e_chunks > 0 else
                             "public API from pinned host buffers (copy, then compute); results stay in HBM, totals are read back"}
                    if ms_e2e else
                    {"value": None, "unit": "elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                     "note": "not measured in this mode: the shard is drawn in HBM (no host copy of it exists)"}),
            "gpu_launches": launches,
            "clocks": clocks,
            "totals": {"sum": totals[0], "selected": totals[1], "leading": totals[2]},
            "totals_check": totals_ok,
            "nccl_gather_per_event_f32": gather_row,
            "roofline": roof, "cpu_baseline": cpu, "ops": ops, "workflows": workflows,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def bind_host_to_gpu(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host memory is allocated, so
    that the staging buffers are first touched on the GPU's NUMA node (at 8 ranks the host-to-device copies of ranks
    whose buffers sit on the far socket fall to a fraction of PCIe line rate).  Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        numa = None
        try:
            numa = int(open("/sys/bus/pci/devices/%s/numa_node" % pynvml.nvmlDeviceGetPciInfo(h).busId.decode().lower()[4:]).read())
        except Exception:
            pass
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return {"bound": True, "cpus": len(cpus), "of": len(allowed), "numa_node": numa}
        return {"bound": False, "cpus": len(allowed), "numa_node": numa,
                "why": "NVML reports every allowed CPU as local to the GPU (no NUMA topology visible in this VM)"}
    except Exception as err:                                   # pragma: no cover
        return {"bound": False, "why": "%s: %s" % (type(err).__name__, err)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def kernel_table(ak, _lib, DeviceArray, empty, runtime, counts_dev, pt_dev, N, M, args):
    """CUDA-event time of every hot-path kernel on the resident shard; algorithmic bytes per SURVEY.md 8(d)."""
    import ctypes
    import torch
    from awkward0_b200.jagged import _vp
    reps = args.kernel_reps
    peak = load_peaks()[0]
    table = {}

    def time_call(fn):
        fn()
        torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1))
        return float(np.mean(best)), float(min(best))

    only = getattr(args, "only", None)

    def add(name, nbytes, elements, fn):
        if only and not re.search(only, name):
            fn()                          # (later rows use what this call leaves behind)
            return
        ms, best = time_call(fn)
        gbs = nbytes / (ms * 1e-3) / 1e9
        table[name] = {"ms": ms, "best_ms": best, "bytes": int(nbytes), "gbs": gbs, "frac": gbs / peak,
                       "elements_per_s": elements / (ms * 1e-3)}

    a = ak.JaggedArray.fromcounts(counts_dev, pt_dev)
    off = a._off64
    st = runtime.stream
    ws, wsb = runtime.workspace(N)
    scal = runtime.scalars()
    v = 4

    offsets_out = empty(N + 1, np.int64)
    add("counts2offsets", 16 * N, N, lambda: _lib.call(
        "jg_counts2offsets", counts_dev.data_ptr(), _lib.JG_I64, N, _vp(offsets_out), _vp(scal), _vp(scal, 1), ws, wsb, st()))
    cnt = empty(N, np.int64)
    add("offsets2counts", 16 * N, N, lambda: _lib.call("jg_offsets2counts", off.data_ptr(), N, _vp(cnt), st()))
    info = runtime.scalars()
    add("validate_offsets", 8 * N, N, lambda: _lib.call(
        "jg_validate_offsets", off.data_ptr(), N, M, _vp(info, 2), _vp(info, 1), st()))
    par = empty(M, np.int64)
    add("offsets2parents", 8 * N + 8 * M, M, lambda: _lib.call("jg_offsets2parents", off.data_ptr(), N, _vp(par), st()))
    add("localindex", 8 * N + 8 * M, M, lambda: _lib.call("jg_localindex", off.data_ptr(), N, _vp(par), st()))
    del par
    out32 = empty(N, np.float32)
    for name, op in (("sum", _lib.SUM), ("max", _lib.MAX), ("min", _lib.MIN), ("prod", _lib.PROD)):
        add("segreduce_%s_f32" % name, v * M + 8 * N + v * N, M, lambda op=op: _lib.call(
            "jg_segreduce", pt_dev.data_ptr(), _lib.JG_F32, off.data_ptr(), N, op, _vp(out32), M, st()))
    best = empty(N, np.int64)
    add("segargmax_f32", v * M + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_segargminmax", pt_dev.data_ptr(), _lib.JG_F32, off.data_ptr(), N, 0, _vp(best), ctypes.c_void_p(0), st()))
    aoff, aidx = empty(N + 1, np.int64), empty(N, np.int64)
    add("segargmax_dense_f32", v * M + 8 * N + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_segargminmax_dense", pt_dev.data_ptr(), _lib.JG_F32, off.data_ptr(), N, 0, _vp(aoff), _vp(aidx), ctypes.c_void_p(0), _vp(scal),
        _vp(scal, 1), ws, wsb, st()))
    # what the sweep runs: the array came from fromcounts, whose scan also counted the non-empty events per tile
    tiles = empty((N + 511) // 512, np.int64)
    add("counts2offsets_tiles", 8 * N + 8 * (N + 1), N, lambda: _lib.call(
        "jg_counts2offsets_tiles", counts_dev.data_ptr(), _lib.JG_I64, N, _vp(offsets_out), _vp(tiles), _vp(scal), _vp(scal, 1), ws, wsb, st()))
    add("segargmax_dense_tiles_f32", v * M + 8 * N + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_segargminmax_dense_tiles", pt_dev.data_ptr(), _lib.JG_F32, off.data_ptr(), N, 0, _vp(tiles), _vp(aoff), _vp(aidx),
        ctypes.c_void_p(0), _vp(scal), _vp(scal, 1), ws, wsb, st()))
    add("compact_nonneg", 8 * N + 8 * N + 8 * N, N, lambda: _lib.call(
        "jg_compact_nonneg", _vp(best), N, _vp(aoff), _vp(aidx), _vp(scal), ws, wsb, st()))
    mask = empty(M, np.bool_)
    add("greater_f32", v * M + M, M, lambda: _lib.call(
        "jg_binary", _lib.OP["GT"], pt_dev.data_ptr(), _lib.JG_F32, ctypes.c_void_p(0), 0, _lib.RHS_SCALAR,
        ctypes.c_double(THRESHOLD), ctypes.c_int64(0), 0, _lib.JG_F32, _vp(mask), _lib.JG_B8, M, ctypes.c_void_p(0), 0, st()))
    sel = empty(M, np.float32)
    noff = empty(N + 1, np.int64)
    ncnt = empty(N, np.int64)            # the result's counts, written by the same sweep
    _lib.call("jg_mask_select", pt_dev.data_ptr(), 4, _vp(mask), 0, off.data_ptr(), N, _vp(sel), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st())
    K = int(scal.cpu()[0])
    add("mask_select_f32", v * M + M + 8 * N + v * K + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_mask_select", pt_dev.data_ptr(), 4, _vp(mask), 0, off.data_ptr(), N, _vp(sel), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st()))
    # a[a > 20] as ONE op: the comparison is evaluated inside the counting and compaction sweeps (no mask).  Algorithmic
    # bytes = what this fused op must move (content twice is NOT counted: once read, K written, offsets in and out, counts)
    add("compare_select_f32", v * M + 8 * N + v * K + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_compare_select", pt_dev.data_ptr(), 4, pt_dev.data_ptr(), _lib.JG_F32, _lib.OP["GT"], ctypes.c_double(THRESHOLD), 0,
        off.data_ptr(), N, _vp(sel), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st()))
    del sel, mask
    # leading-object gather a[a.argmax()]
    lead = a.argmax()
    Kl = len(lead.content)
    gout = empty(Kl, np.float32)
    add("index_gather_f32", 8 * Kl + 16 * N + v * Kl + v * Kl, Kl, lambda: _lib.call(
        "jg_index_gather", pt_dev.data_ptr(), 4, off.data_ptr(), lead.content.data_ptr(), _lib.JG_I64,
        lead._off64.data_ptr(), N, _vp(gout), _vp(scal, 1), st()))
    del lead, gout
    flat = ak.DeviceArray(out32)
    bout = empty(M, np.float32)
    add("broadcast_f32", v * N + 8 * N + v * M, M, lambda: _lib.call(
        "jg_broadcast", flat.data_ptr(), 4, off.data_ptr(), N, _vp(bout), st()))
    add("add_flat_f32", v * M + v * N + 8 * N + v * M, M, lambda: _lib.call(
        "jg_binary", _lib.OP["ADD"], pt_dev.data_ptr(), _lib.JG_F32, flat.data_ptr(), _lib.JG_F32, _lib.RHS_PARENT,
        ctypes.c_double(0), ctypes.c_int64(0), 0, _lib.JG_F32, _vp(bout), _lib.JG_F32, M, off.data_ptr(), N, st()))
    del bout

    # per-event argsort (one pass: content + offsets in, one int64 local index per element out)
    sidx = empty(M, np.int64)
    add("argsort_f32", v * M + 8 * N + 8 * M, M, lambda: _lib.call(
        "jg_argsort", pt_dev.data_ptr(), _lib.JG_F32, off.data_ptr(), N, 0, _vp(sidx), _vp(scal, 1), _vp(scal, 2), 1, st()))
    del sidx
    # concatenate(axis=1) slot copy of one input (dense content -> its slot in the merged events)
    cat = empty(2 * M, np.float32)
    dst = ak.DeviceArray(off.tensor[:-1] * 2)
    add("segment_scatter_f32", v * M + 16 * N + v * M, M, lambda: _lib.call(
        "jg_segment_scatter", pt_dev.data_ptr(), 4, off.data_ptr(), dst.data_ptr(), N, _vp(cat), st()))
    del dst
    # the whole axis-1 concatenation of two inputs as ONE sweep over the output layout (coalesced loads and stores)
    pt2, off2 = ak.DeviceArray(pt_dev.tensor.clone()), ak.DeviceArray(off.tensor.clone())
    moff = ak.DeviceArray(off.tensor * 2)
    mptrs = (ctypes.c_void_p * 2)(pt_dev.data_ptr(), pt2.data_ptr())
    moffs = (ctypes.c_void_p * 2)(off.data_ptr(), off2.data_ptr())
    add("segment_merge2_f32", 2 * v * M + 3 * 8 * N + 2 * v * M, 2 * M, lambda: _lib.call(
        "jg_segment_merge", mptrs, moffs, 2, 4, moff.data_ptr(), N, _vp(cat), st()))
    del cat, pt2, off2, moff

    # float64 content (BASELINE configs[1] names float32/float64): same offsets, values cast up
    pt64 = ak.DeviceArray(pt_dev.tensor.double())
    out64 = empty(N, np.float64)
    add("segreduce_sum_f64", 8 * M + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_segreduce", pt64.data_ptr(), _lib.JG_F64, off.data_ptr(), N, _lib.SUM, _vp(out64), M, st()))
    aoff64, aidx64 = empty(N + 1, np.int64), empty(N, np.int64)
    add("segargmax_dense_f64", 8 * M + 8 * N + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_segargminmax_dense", pt64.data_ptr(), _lib.JG_F64, off.data_ptr(), N, 0, _vp(aoff64), _vp(aidx64), ctypes.c_void_p(0), _vp(scal),
        _vp(scal, 1), ws, wsb, st()))
    mask64 = empty(M, np.bool_)
    add("greater_f64", 8 * M + M, M, lambda: _lib.call(
        "jg_binary", _lib.OP["GT"], pt64.data_ptr(), _lib.JG_F64, ctypes.c_void_p(0), 0, _lib.RHS_SCALAR,
        ctypes.c_double(THRESHOLD), ctypes.c_int64(0), 0, _lib.JG_F64, _vp(mask64), _lib.JG_B8, M, ctypes.c_void_p(0), 0, st()))
    sel64 = empty(M, np.float64)
    add("mask_select_f64", 8 * M + M + 8 * N + 8 * K + 8 * N + 8 * N, M, lambda: _lib.call(
        "jg_mask_select", pt64.data_ptr(), 8, _vp(mask64), 0, off.data_ptr(), N, _vp(sel64), _vp(noff), _vp(ncnt), _vp(scal), ws, wsb, st()))
    del pt64, out64, aoff64, aidx64, mask64, sel64

    # combinatorics on a dimuon-like shard (BASELINE configs[3]: Poisson(2) counts)
    Nc = args.comb_events
    rng = np.random.default_rng(4242)
    mu_off = ak.JaggedArray.counts2offsets(ak.asdevice(rng.poisson(2.0, Nc).astype(np.int64)))
    el_off = ak.JaggedArray.counts2offsets(ak.asdevice(rng.poisson(2.0, Nc).astype(np.int64)))
    Mmu = int(mu_off[-1])
    mu_pt = ak.asdevice(rng.exponential(20.0, Mmu).astype(np.float32))
    Mel = int(el_off[-1])
    el_pt = ak.asdevice(rng.exponential(20.0, Mel).astype(np.float32))
    mu_eta, mu_phi = ak.asdevice(rng.uniform(-2.4, 2.4, Mmu).astype(np.float32)), ak.asdevice(rng.uniform(-np.pi, np.pi, Mmu).astype(np.float32))
    el_eta, el_phi = ak.asdevice(rng.uniform(-2.4, 2.4, Mel).astype(np.float32)), ak.asdevice(rng.uniform(-np.pi, np.pi, Mel).astype(np.float32))
    wsc, wscb = runtime.workspace(Nc)
    for kind, name, ob in ((_lib.COMB_PAIRS, "pairs", None), (_lib.COMB_DISTINCTS, "distincts", None), (_lib.COMB_CROSS, "cross", el_off)):
        poff = empty(Nc + 1, np.int64)
        obp = ob.data_ptr() if ob is not None else ctypes.c_void_p(0)
        nin = 8 * Nc * (2 if ob is not None else 1)
        add("comb_offsets_" + name, nin + 8 * Nc, Nc, lambda: _lib.call(
            "jg_comb_offsets", kind, 2, mu_off.data_ptr(), obp, Nc, _vp(poff), _vp(scal), wsc, wscb, st()))
        P = int(scal.cpu()[0])
        c0, c1 = empty(P, np.int64), empty(P, np.int64)
        arr = (ctypes.c_void_p * 2)(c0.data_ptr(), c1.data_ptr())
        add("arg" + name + "_fill", nin + 8 * Nc + 16 * P, P, lambda: _lib.call(
            "jg_comb_fill", kind, 2, mu_off.data_ptr(), obp, Nc, _vp(poff), arr, 0, st()))
        del c0, c1
        if kind == _lib.COMB_PAIRS:
            # argchoose(k), k = 2..5 (jagged.py:1128-1221): colexicographic k-combinations, k index columns.  Counted on
            # a Poisson(5) shard: Poisson(2) events hold almost no 4- or 5-combinations
            Nk = min(N, 20_000_000)
            for kk in (2, 3, 4, 5):
                koff = empty(Nk + 1, np.int64)
                add("comb_offsets_choose_k%d" % kk, 16 * Nk, Nk, lambda: _lib.call(
                    "jg_comb_offsets", _lib.COMB_CHOOSE, kk, off.data_ptr(), ctypes.c_void_p(0), Nk, _vp(koff), _vp(scal), wsc, wscb, st()))
                Pk = int(scal.cpu()[0])
                kcols = [empty(Pk, np.int64) for _ in range(kk)]
                karr = (ctypes.c_void_p * kk)(*[c.data_ptr() for c in kcols])
                add("argchoose_k%d_fill" % kk, 16 * Nk + 8 * kk * Pk, Pk, lambda: _lib.call(
                    "jg_comb_fill", _lib.COMB_CHOOSE, kk, off.data_ptr(), ctypes.c_void_p(0), Nk, _vp(koff), karr, 0, st()))
                if "argchoose_k%d_fill" % kk in table:
                    table["argchoose_k%d_fill" % kk]["combinations"] = Pk
                del kcols, koff
        l, r = empty(P, np.float32), empty(P, np.float32)
        cb = el_pt if ob is not None else mu_pt
        add(name + "_values_f32", nin + 8 * Nc + 4 * Mmu + (4 * Mel if ob is not None else 0) + 8 * P, P, lambda: _lib.call(
            "jg_comb_gather2", kind, mu_off.data_ptr(), obp, Nc, _vp(poff), mu_pt.data_ptr(), 4, cb.data_ptr(), 4,
            _vp(l), _vp(r), st()))
        del l, r
        if "arg" + name + "_fill" in table:
            table["arg" + name + "_fill"]["pairs"] = P
        # SURVEY 8d "pair mass fused from index generation": three columns per side, one result column; the program
        # is what awkward0_b200/lazy.py compiles for sqrt(2 pt1 pt2 (cosh(eta1 - eta2) - cos(phi1 - phi2)))
        if kind != _lib.COMB_DISTINCTS:
            from awkward0_b200 import lazy
            ctx = lazy.CombContext(kind, mu_off, ob, Nc, ak.DeviceArray(poff), P)
            other = (el_pt, el_eta, el_phi) if ob is not None else (mu_pt, mu_eta, mu_phi)
            (pt1, eta1, phi1), (pt2, eta2, phi2) = [ctx.leaf("L", c) for c in (mu_pt, mu_eta, mu_phi)], [ctx.leaf("R", c) for c in other]
            prog, _ = lazy._compile(np.sqrt(2 * pt1 * pt2 * (np.cosh(eta1 - eta2) - np.cos(phi1 - phi2))))
            mass_out = empty(P, np.float32)
            add(name + "_mass_fused_f32", nin + 8 * Nc + 3 * 4 * Mmu + (3 * 4 * Mel if ob is not None else 0) + 4 * P, P,
                lambda: _lib.call("jg_comb_eval", kind, mu_off.data_ptr(), obp, Nc, _vp(poff), _lib.JG_F32, ctypes.byref(prog),
                                  _vp(mass_out), st()))
            if name + "_mass_fused_f32" in table:
                table[name + "_mass_fused_f32"]["note"] = "issue-bound (14-step program incl. cosh, cos, sqrt per pair), not HBM-bound"
            del mass_out
    return table


def distribution_table(ak, args):
    """north_star: "variants chosen by count distribution".  The same four hot-path operations -- sum, argmax, a[a > 20],
    parents -- through the public API on count distributions other than the Poisson(5) the sweep is quoted on: mean
    counts of 10 (the event tile staged 32 KB at a time for the reductions), long events (Poisson(40)), a heavy tail (geometric, mean 4, plus events of 5 k - 70 k elements) and a few huge events
    (1000 events x 1 M elements).  CUDA events around the call, host synchronisations included; algorithmic bytes as
    in SURVEY 8d."""
    import torch
    peak = load_peaks()[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    out = {}

    def shapes():
        n = args.dist_elements // 10
        yield "poisson10", torch.poisson(torch.full((n,), 10.0, device=dev), generator=g).to(torch.int64)
        n = args.dist_elements // 40
        yield "poisson40", torch.poisson(torch.full((n,), 40.0, device=dev), generator=g).to(torch.int64)
        n = args.dist_elements // 4
        c = (torch.empty(n, device=dev, dtype=torch.float32).geometric_(0.2, generator=g).to(torch.int64) - 1).clamp_(min=0)
        where = torch.randint(0, n, (1000,), device=dev, generator=g)
        c[where] = torch.randint(5000, 70000, (1000,), device=dev, generator=g)
        yield "geometric_tail", c
        yield "huge_1k_x_1M", torch.full((1000,), max(1, args.dist_elements // 1000), device=dev, dtype=torch.int64)

    for name, counts_t in shapes():
        N = int(counts_t.numel())
        M = int(counts_t.sum())
        pt_t = torch.empty(M, device=dev, dtype=torch.float32).exponential_(1.0 / 25.0, generator=g)
        a = ak.JaggedArray.fromcounts(ak.DeviceArray(counts_t), ak.DeviceArray(pt_t))
        off = a.offsets
        K = len(a[a > THRESHOLD].content)
        npos = int((counts_t > 0).sum())
        rows = (("segreduce_sum_f32", 4 * M + 8 * N + 4 * N, lambda: a.sum()),
                ("segargmax_dense_f32", 4 * M + 8 * N + 8 * N + 8 * npos, lambda: a.argmax()),
                ("compare_select_f32", 4 * M + 8 * N + 4 * K + 16 * N, lambda: a[a > THRESHOLD]),
                ("offsets2parents", 8 * N + 8 * M, lambda: ak.JaggedArray.offsets2parents(off)))
        for op, nbytes, fn in rows:
            r = fn()
            del r
            torch.cuda.synchronize()
            ts = []
            for _ in range(args.kernel_reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
                del r
            ms = float(np.mean(ts))
            gbs = nbytes / (ms * 1e-3) / 1e9
            out["%s_%s" % (name, op)] = {"ms": ms, "best_ms": float(min(ts)), "bytes": int(nbytes), "gbs": gbs, "frac": gbs / peak,
                                         "elements_per_s": M / (ms * 1e-3), "events": N, "elements": M,
                                         "max_count": int(counts_t.max())}
        del a, off, pt_t, counts_t
        torch.cuda.empty_cache()
    return out


def workflow_table(ak, counts_dev, pt_dev, N, M, args):
    """BASELINE.json configs[1..3] through the PUBLIC API on resident data (CUDA events around the whole call
    chain, host synchronisations for the data-dependent sizes included)."""
    import torch
    out = {}

    def timed(name, fn, units, unit_name, reps=5):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            del r
        ms = float(np.mean(ts))
        out[name] = {"ms": ms, unit_name + "_per_s": units / (ms * 1e-3)}

    a = ak.JaggedArray.fromcounts(counts_dev, pt_dev)

    def c2(arr):
        sel = arr[arr > THRESHOLD]
        return sel.flatten(), sel.counts

    timed("C1_sum_f32", lambda: a.sum(), M, "elements")
    timed("C1_max_f32", lambda: a.max(), M, "elements")
    timed("C2_mask_flatten_counts_f32", lambda: c2(a), M, "elements")
    a64 = ak.JaggedArray.fromoffsets(a.offsets, ak.DeviceArray(pt_dev.tensor.double()))
    timed("C2_mask_flatten_counts_f64", lambda: c2(a64), M, "elements")
    del a64
    timed("C3_leading_object_argmax_getitem_f32", lambda: a[a.argmax()], M, "elements")
    timed("sort_by_value_argsort_getitem_f32", lambda: a[a.argsort()], M, "elements")
    timed("concatenate_axis0_f32", lambda: a.concatenate([a]), 2 * M, "elements", reps=2)
    timed("concatenate_axis1_f32", lambda: a.concatenate([a], axis=1), 2 * M, "elements", reps=2)

    # C4: dimuon combinatorics with the per-pair invariant mass written in ufuncs (records: pt, eta, phi)
    Nc = args.comb_events
    rng = np.random.default_rng(777)
    mc = rng.poisson(2.0, Nc).astype(np.int64)
    ec = rng.poisson(2.0, Nc).astype(np.int64)
    Mm, Me = int(mc.sum()), int(ec.sum())

    def collection(counts, m):
        cd = ak.asdevice(counts)
        first = ak.JaggedArray.fromcounts(cd, rng.exponential(20.0, m).astype(np.float32))
        mk = lambda x: ak.JaggedArray.fromoffsets(first.offsets, x)
        return ak.JaggedArray.zip(pt=first, eta=mk(rng.uniform(-2.4, 2.4, m).astype(np.float32)),
                                  phi=mk(rng.uniform(-np.pi, np.pi, m).astype(np.float32)))

    mu, el = collection(mc, Mm), collection(ec, Me)

    def mass(p):
        one, two = p.i0, p.i1
        return np.sqrt(2 * one.pt * two.pt * (np.cosh(one.eta - two.eta) - np.cos(one.phi - two.phi)))

    def force(j):
        """The columns of pairs()/cross() results and ufunc chains on them are deferred (awkward0_b200/lazy.py):
        touching the device buffer is what runs the gather / the fused kernel."""
        for leaf in _leaf_columns(j.content):
            leaf.tensor
        return j

    def unfused_mass(p):
        ak.lazy.ENABLED[0] = False
        try:
            return mass(force(p))
        finally:
            ak.lazy.ENABLED[0] = True

    P = int(((mc * (mc + 1)) // 2).sum())
    timed("C4_pairs_records", lambda: force(mu.pairs()), P, "pairs")
    timed("C4_pairs_plus_invariant_mass", lambda: force(mass(mu.pairs())), P, "pairs")
    timed("C4_pairs_plus_invariant_mass_unfused", lambda: unfused_mass(mu.pairs()), P, "pairs")
    timed("C4_argpairs", lambda: mu.argpairs(), P, "pairs")
    X = int((mc * ec).sum())
    timed("C4_cross_records", lambda: force(mu.cross(el)), X, "pairs")
    timed("C4_cross_plus_invariant_mass", lambda: force(mass(mu.cross(el))), X, "pairs")
    timed("C4_cross_plus_invariant_mass_unfused", lambda: unfused_mass(mu.cross(el)), X, "pairs")
    D = int(((mc * (mc - 1)) // 2).sum())
    timed("C4_distincts_plus_invariant_mass", lambda: force(mass(mu.distincts())), D, "pairs")
    out["C4_sizes"] = {"events": Nc, "muons": Mm, "electrons": Me, "pairs": P, "cross_pairs": X}
    return out


def cpu_workflows(args):
    """BASELINE configs[0..3] on the host CPU (one core: numpy's cumsum / repeat / reduceat / take are single
    threaded), the same API chains as workflow_table on a stated sample of the same synthetic data.  With the
    reference present (oracle/_ref) this is the reference itself; otherwise the oracle port covers what it restates."""
    ref = reference_module()
    out = {}

    def best_of(fn, reps=1):
        fn()                      # warm-up (page faults of the temporaries)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    def put(name, seconds, units, unit_name, sample):
        out[name] = {"cpu_ms": 1e3 * seconds, "cpu_" + unit_name + "_per_s": units / seconds, "cpu_sample": sample,
                     "cpu_kind": cpu_kind(), "cpu_cores": 1}

    n = args.cpu_workflow_events
    counts, pt = make_shard(n, 12345)
    M = len(pt)
    sample = "%d events, %d elements (Poisson(5), Exp(25))" % (n, M)
    if ref is not None:
        JA = ref.JaggedArray
        a = JA.fromcounts(counts, pt)
        put("C1_sum_f32", best_of(lambda: a.sum()), M, "elements", sample)
        put("C1_max_f32", best_of(lambda: a.max()), M, "elements", sample)

        def c2(arr):
            sel = arr[arr > THRESHOLD]
            return sel.flatten(), sel.counts
        put("C2_mask_flatten_counts_f32", best_of(lambda: c2(a)), M, "elements", sample)
        a64 = JA.fromoffsets(a.offsets, pt.astype(np.float64))
        put("C2_mask_flatten_counts_f64", best_of(lambda: c2(a64)), M, "elements", sample)
        del a64
        put("C3_leading_object_argmax_getitem_f32", best_of(lambda: a[a.argmax()]), M, "elements", sample)
        nc = args.cpu_comb_events
        rng = np.random.default_rng(777)
        mc, ec = rng.poisson(2.0, nc).astype(np.int64), rng.poisson(2.0, nc).astype(np.int64)
        Mm, Me = int(mc.sum()), int(ec.sum())

        def collection(c, m):
            first = JA.fromcounts(c, rng.exponential(20.0, m).astype(np.float32))
            mk = lambda x: JA.fromoffsets(first.offsets, x)
            return JA.zip(pt=first, eta=mk(rng.uniform(-2.4, 2.4, m).astype(np.float32)),
                          phi=mk(rng.uniform(-np.pi, np.pi, m).astype(np.float32)))

        mu, el = collection(mc, Mm), collection(ec, Me)

        def mass(p):
            one, two = p.i0, p.i1
            return np.sqrt(2 * one.pt * two.pt * (np.cosh(one.eta - two.eta) - np.cos(one.phi - two.phi)))

        csample = "%d events, %d muons, %d electrons (Poisson(2))" % (nc, Mm, Me)
        P, X, D = int(((mc * (mc + 1)) // 2).sum()), int((mc * ec).sum()), int(((mc * (mc - 1)) // 2).sum())
        put("C4_argpairs", best_of(lambda: mu.argpairs(), 1), P, "pairs", csample)
        put("C4_pairs_plus_invariant_mass", best_of(lambda: mass(mu.pairs()), 1), P, "pairs", csample)
        put("C4_cross_plus_invariant_mass", best_of(lambda: mass(mu.cross(el)), 1), X, "pairs", csample)
        put("C4_distincts_plus_invariant_mass", best_of(lambda: mass(mu.distincts()), 1), D, "pairs", csample)
    else:
        from oracle import jagged_oracle as O
        off = O.counts2offsets(counts)
        put("C1_sum_f32", best_of(lambda: O.reduce(off, pt, "sum")), M, "elements", sample)
        put("C1_max_f32", best_of(lambda: O.reduce(off, pt, "max")), M, "elements", sample)

        def c2(content):
            moff, mask = O.ufunc_broadcast(np.greater, off, content, content.dtype.type(THRESHOLD))
            soff, sel = O.mask_select(off, content, moff, mask)
            return O.flatten(soff, sel), O.offsets2counts(soff)
        put("C2_mask_flatten_counts_f32", best_of(lambda: c2(pt)), M, "elements", sample)
        pt64 = pt.astype(np.float64)
        put("C2_mask_flatten_counts_f64", best_of(lambda: c2(pt64)), M, "elements", sample)

        def c3():
            aoff, aidx = O.argminmax(off, pt, False)
            return O.index_gather(off, pt, aoff, aidx)
        put("C3_leading_object_argmax_getitem_f32", best_of(c3), M, "elements", sample)
        nc = args.cpu_comb_events
        mc = np.random.default_rng(777).poisson(2.0, nc).astype(np.int64)
        moff = O.counts2offsets(mc)
        put("C4_argpairs", best_of(lambda: O.argpairs(moff), 1), int(((mc * (mc + 1)) // 2).sum()), "pairs",
            "%d events (Poisson(2))" % nc)
    return out


def _leaf_columns(content):
    if hasattr(content, "columns") and not hasattr(content, "tensor"):
        for name in content.columns:
            for leaf in _leaf_columns(content[name]):
                yield leaf
    else:
        yield content


def cpu_baseline(args):
    """Oracle port (numpy, 1 core) on a bounded sample of the same workload."""
    n = args.cpu_events
    counts, pt = make_shard(n, 12345)
    cpu_sweep(counts[: n // 8], pt[: int(counts[: n // 8].sum())])     # warm numpy
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        elements, _ = cpu_sweep(counts, pt)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": elements / best, "unit": "elements/s", "cores": 1, "kind": cpu_kind(),
            "sample": "%d events (%d elements) of the same sweep, %s, best of 2; host has %d cores"
                      % (n, elements, CPU_KIND_NOTE[cpu_kind()], os.cpu_count() or 0),
            "seconds": best}


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, ...) write to fd 1; keep stdout for the ONE JSON line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events-per-gpu", type=int, default=125_000_000)
    ap.add_argument("--total-events", type=float, default=0,
                    help="STRONG scaling: this many events in total, split over the ranks (BASELINE configs[4]: 1e9); the "
                         "shards are drawn on the device and the e2e leg is skipped")
    ap.add_argument("--device-rng", action="store_true", help="draw the weak-scaling shard on the device as well")
    ap.add_argument("--comb-events", type=int, default=50_000_000)
    ap.add_argument("--cpu-events", type=int, default=4_000_000)
    ap.add_argument("--ref-events-per-core", type=int, default=1_000_000)
    ap.add_argument("--cpu-workflow-events", type=int, default=10_000_000,
                    help="events of the CPU sample behind every workflow's cpu_ms (C1 runs at its full size)")
    ap.add_argument("--cpu-comb-events", type=int, default=1_000_000)
    ap.add_argument("--kernel-reps", type=int, default=5)
    ap.add_argument("--dist-elements", type=int, default=600_000_000,
                    help="content elements of each shard of the count-distribution table (Poisson(40), heavy tail, huge events)")
    ap.add_argument("--e2e-chunks", type=int, default=8,
                    help="event chunks of the streamed host->HBM ingest used by the e2e leg (0: copy everything, then compute)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-tables", action="store_true",
                    help="skip the per-kernel roofline table and the API workflow table (scaling runs: only the sweep is wanted)")
    args = ap.parse_args()
    args.total_events = int(args.total_events)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
