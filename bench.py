#!/usr/bin/env python
"""bench.py -- gene-trait tests/sec (incl. permutations) of the hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

A "step" is one pass of the whole hot path over one batch of synthetic input
(SURVEY.md 8(d)): contingency + Fisher for every gene, the unpermuted
pairwise-comparison walk and P label permutations for every gene
(-p 1.0, exhaustive mode: no early stop), then -- for N > 1 -- one NCCL
all-gather of the per-gene records.  tests = G_tested * T * (1 + P).

  value : whole-job tests/s with inputs resident in HBM when the clock starts
  e2e   : the same through the host-buffer C-ABI calls (pinned host bitsets in,
          host result arrays out; H2D/D2H inside the timed region)
  roofline / roofline_int32 : the dominant kernel (K5, permutation walks)
  cpu_baseline : the oracle's C port on this box's host cores, bounded sample

Default workload = BASELINE.json configs[2] ("c3": 50k genes x 5k isolates x 1
trait, 1000 permutations + pairwise), the largest single-GPU configuration the
metric "tests/sec incl. permutations" is quoted on; configs[1] (Fisher only,
no permutations) is reported alongside as `fisher_pass`.  Under torchrun each
rank owns its own 50k-gene shard (weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

OPS_PER_NODE = 76          # int32 add/max operations per internal node, contract figure


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4", "c5", "north_star"])
    ap.add_argument("--genes", type=int, default=0, help="override genes per GPU")
    ap.add_argument("--isolates", type=int, default=0)
    ap.add_argument("--perms", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(a):
    from scoary_b200 import synth
    G, N, T, P, seed = synth.CONFIGS[a.workload]
    if a.workload in ("c4", "c5"):
        G = G // 8                      # per-GPU shard of the 8-GPU configurations
    if a.genes:
        G = a.genes
    if a.isolates:
        N = a.isolates
    if a.perms >= 0:
        P = a.perms
    return G, N, T, P, seed


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def usable_host_threads():
    """Host threads this process can really use: cpu_count, clipped by the affinity mask and by the
    container's CPU quota (cgroup cpu.max / cfs_quota).  With a quota of q CPUs, 2q threads measured
    best (more only thrash: 128 threads on a 16-CPU quota ran 2.5x slower than 32)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    if quota:
        n = min(n, max(1, int(round(2 * quota))))
    return n, quota


def cpu_sample(G, N, T, P, seed, budget_s):
    """Time the oracle's C port (OpenMP, all host threads) on a bounded sample of the
    same workload: the first `gs` genes x `ps` permutations (+ unpermuted walk + Fisher)."""
    from oracle import oracle as O
    from scoary_b200 import synth
    threads, quota = usable_host_threads()
    O.set_num_threads(threads)           # (torch's import / torchrun would otherwise pin OpenMP to fewer threads)
    traits = synth.make_traits(N, 1, seed)
    nested = synth.make_tree(N, seed)
    left, right, names = O.flatten_tree(nested)
    col = {n: j for j, n in enumerate(synth.isolate_names(N))}
    cols = np.asarray([col[n] for n in names])
    labels = traits[0][cols].astype(np.uint8)
    # bounded sample: ~200 genes per host thread x 32 permutations (a few seconds of wall time with the
    # -O3 port; dynamic scheduling over genes keeps every thread busy)
    ps = min(P, 32) if P > 0 else 0
    per_thread = max(1, int(round(192 * budget_s / 6.0)))
    gs = min(G, per_thread * threads) if P > 0 else min(G, 2000)
    bits = synth.make_genes_packed(gs, N, seed, traits=traits)
    m = synth.unpack_rows(bits, N)

    def run():
        """one pass over the sample; returns (seconds in contingency + Fisher, seconds in walks)"""
        t0 = time.perf_counter()
        counts = O.contingency(m, traits[0])
        O.fisher(counts)
        t1 = time.perf_counter()
        if P > 0:
            O.permute(left, right, m[:, cols], labels, P=ps, seed=seed)
        return t1 - t0, time.perf_counter() - t1

    def rate(t_stats, t_walks):
        """tests/s on the FULL workload shape: contingency + Fisher once per gene, 1 + P walks per gene
        (the sample walks 1 + ps labellings per gene; per-walk cost is constant)"""
        per_gene = t_stats / gs + ((t_walks / (gs * (1 + ps))) * (1 + P) if P > 0 else 0.0)
        return T * (1 + P) / (T * per_gene)

    sample = ("first %d genes x %d of %d permutations (+ unpermuted walk, contingency, Fisher), %d isolates, "
              "exhaustive mode; per-gene and per-walk costs combined at the full %d permutations per gene; %d OpenMP "
              "threads (os.cpu_count() = %s, container CPU quota = %s)"
              % (gs, ps, P, N, P, threads, os.cpu_count(), ("%.0f CPUs" % quota) if quota else "none"))
    return run, rate, threads, sample


def run_reference(a):
    G, N, T, P, seed = workload(a)
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget = max(1.0, min(6.0, 100.0 / max(1, a.steps + a.warmup)))
    run, rate, threads, sample = cpu_sample(G, N, T, P, seed, budget)
    for _ in range(a.warmup):
        run()
    t0 = time.perf_counter()
    ts = tw = 0.0
    for _ in range(a.steps):
        x, y = run()
        ts += x
        tw += y
    dt = time.perf_counter() - t0
    value = rate(ts / a.steps, tw / a.steps)
    line = {
        "impl": "reference", "metric": "gene-trait tests/sec (incl. permutations)", "value": value, "unit": "tests/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": "%s: %d genes x %d isolates x %d trait(s), %d permutations + pairwise" %
                               (a.workload, G, N, T, P)},
        "cpu_baseline": {"value": value, "unit": "tests/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of scoary/methods.py + classes.py (oracle/scoary_oracle.c, OpenMP); "
                                 "the Python reference itself measured 7 walks/s/core at N=5000 (BASELINE.md)"},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from scoary_b200 import synth
    from scoary_b200.engine import Engine, words_for

    G, N, T, P, seed = workload(a)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic inputs: every rank owns its own G-gene shard (weak scaling), same traits / tree
    traits = synth.make_traits(N, T, seed)
    bits_np = synth.make_genes_packed(G, N, seed + 1000003 * rank, traits=traits if rank == 0 else None)
    W = words_for(N)
    pinned = torch.empty((G, W), dtype=torch.int64, pin_memory=True)
    pinned.numpy().view(np.uint64)[:] = bits_np
    nested = synth.make_tree(N, seed)
    names = synth.isolate_names(N)
    col = {n: j for j, n in enumerate(names)}
    from scoary_b200 import tree as treemod
    left, right, leaf_names = treemod.flatten(nested)
    leaf_cols = np.asarray([col[n] for n in leaf_names], dtype=np.int32)

    e = Engine(local)
    # a dedicated (non-default) torch stream: the library launches on it and torch.cuda.Event
    # timing below sees the same stream (handle 0 would mean "the context's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    e.set_stream(stream.cuda_stream)
    int_peak = e.int32_peak(8192)

    # ---- device-resident state for `value`
    d_bits = pinned.to(dev, non_blocking=False)
    e.set_genes_device(d_bits.data_ptr(), G, N, W)
    for t in range(T):
        e.set_trait_vector(t, traits[t])
        e.set_tree(t, left, right, leaf_cols)
    d_counts = torch.empty((T, G, 4), dtype=torch.int32, device=dev)
    d_p = torch.empty((T, G), dtype=torch.float64, device=dev)
    d_pairs = torch.empty((T, G, 3), dtype=torch.int32, device=dev)
    d_r = torch.zeros((T, G), dtype=torch.int32, device=dev)
    d_nd = torch.zeros((T, G), dtype=torch.int32, device=dev)
    rec_w = 4 + 2 + 3 + 2
    d_rec = torch.empty((T, G, rec_w), dtype=torch.int32, device=dev)
    d_all = torch.empty((world * T, G, rec_w), dtype=torch.int32, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_device():
        for t in range(T):
            e.contingency_fisher_device(t, d_counts[t].data_ptr(), d_p[t].data_ptr())
            if P > 0:
                e.permute_device(t, G, P, seed, d_pairs[t].data_ptr(), d_r[t].data_ptr(), d_nd[t].data_ptr())
        if world > 1:   # one all-gather of the fixed-size per-gene records
            d_rec[..., 0:4] = d_counts
            d_rec[..., 4:6] = d_p.view(torch.int32).view(T, G, 2)
            d_rec[..., 6:9] = d_pairs
            d_rec[..., 9] = d_r
            d_rec[..., 10] = d_nd
            dist.all_gather_into_tensor(d_all, d_rec)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(a.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    counts_host = d_counts.cpu().numpy()
    g_tested = int(((counts_host[..., 0] + counts_host[..., 1] > 0) &
                    (counts_host[..., 2] + counts_host[..., 3] > 0)).sum())
    tests_per_step_rank = g_tested * (1 + P)

    e.stats_reset()
    e.set_profiling(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for k in range(a.steps):
        flush.zero_()                      # evict L2 between timed iterations (untimed)
        evs[k][0].record()
        step_device()
        evs[k][1].record()
    torch.cuda.synchronize()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(s.elapsed_time(t) for s, t in evs)
    st = e.stats()
    e.set_profiling(False)
    if world > 1:
        tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_ms = float(tmax.item())
        tt = torch.tensor([tests_per_step_rank], dtype=torch.int64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        tests_per_step = int(tt.item())
    else:
        tests_per_step = tests_per_step_rank
    value = tests_per_step * a.steps / (dev_ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI, copies inside the timed region
    counts_h = np.empty((G, 4), dtype=np.int32)
    bits_host = pinned.numpy().view(np.uint64)

    e2e_trace = []

    def step_e2e():
        h2d = d2h = 0
        t0 = time.perf_counter()
        e.set_genes(bits_host, N)
        h2d += bits_host.nbytes
        t1 = time.perf_counter()
        tf = tp = 0.0
        for t in range(T):
            e.set_trait_vector(t, traits[t])
            e.set_tree(t, left, right, leaf_cols)
            h2d += 2 * W * 8 + left.nbytes + right.nbytes + leaf_cols.nbytes
            ta = time.perf_counter()
            c, p, _ = e.contingency_fisher(t)
            d2h += c.nbytes + p.nbytes
            tb = time.perf_counter()
            if P > 0:
                pairs, r, nd = e.permute(t, P, seed=seed)
                d2h += pairs.nbytes + r.nbytes + nd.nbytes
            tc = time.perf_counter()
            tf += tb - ta
            tp += tc - tb
        e2e_trace.append({"set_genes_ms": (t1 - t0) * 1e3, "fisher_ms": tf * 1e3, "permute_ms": tp * 1e3,
                          "total_ms": (time.perf_counter() - t0) * 1e3})
        return h2d, d2h

    e.set_stream(0)
    step_e2e()
    barrier()
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    n_e2e = max(1, min(a.steps, 3))
    for _ in range(n_e2e):
        h2d_b, d2h_b = step_e2e()
    e.synchronize()
    e2e_s = time.perf_counter() - e0
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_value = tests_per_step * n_e2e / e2e_s

    # ---- reference-rule mode (second number): the reference's sequential early stop
    # (methods.py:1360-1363) applied between growing slices of permutations
    ref_rule = None
    if P >= 32:
        from scoary_b200.methods import early_stop_table
        e.set_stream(stream.cuda_stream)
        e.set_genes_device(d_bits.data_ptr(), G, N, W)
        for t in range(T):
            e.set_trait_vector(t, traits[t])
            e.set_tree(t, left, right, leaf_cols)
        d_rmin = torch.from_numpy(early_stop_table(P)).to(dev)
        def step_rule():
            for t in range(T):
                e.permute_device(t, G, P, seed, d_pairs[t].data_ptr(), d_r[t].data_ptr(), d_nd[t].data_ptr(),
                                 early_stop=True, rmin_ptr=d_rmin.data_ptr())
        step_rule()
        torch.cuda.synchronize()
        e.stats_reset()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        step_rule()
        r1.record()
        torch.cuda.synchronize()
        rule_ms = r0.elapsed_time(r1)
        walks = e.stats()["tests_walks"]
        nd = d_nd.cpu().numpy()
        ref_rule = {"ms_per_step": rule_ms, "walks_executed": int(walks), "walks_exhaustive": int(G * T * (1 + P)),
                    "genes_stopped_early": int((nd < P).sum()), "genes": int(G * T),
                    "equivalent_tests_per_s": G * T * (1 + P) / (rule_ms * 1e-3),
                    "note": "Permute's early stop on; pairwise walk + permutations only (no Fisher pass)"}

    # ---- Fisher-only pass (BASELINE configs[1] shape of work), device resident
    e.set_stream(stream.cuda_stream)
    e.set_genes_device(d_bits.data_ptr(), G, N, W)
    for t in range(T):
        e.set_trait_vector(t, traits[t])
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e.contingency_fisher_device(0, d_counts[0].data_ptr(), d_p[0].data_ptr())
    flush.zero_()
    f0.record()
    e.contingency_fisher_device(0, d_counts[0].data_ptr(), d_p[0].data_ptr())
    f1.record()
    torch.cuda.synchronize()
    fisher_ms = f0.elapsed_time(f1)

    # ---- BASELINE configs[1] itself (10k genes x 1k isolates, Fisher only, no tree): a second, small line
    c2 = None
    if rank == 0 and a.workload != "c2":
        G2, N2, _, _, seed2 = synth.CONFIGS["c2"]
        tr2 = synth.make_traits(N2, 1, seed2)
        bits2 = synth.make_genes_packed(G2, N2, seed2, traits=tr2)
        W2 = words_for(N2)
        pin2 = torch.empty((G2, W2), dtype=torch.int64, pin_memory=True)
        pin2.numpy().view(np.uint64)[:] = bits2
        d_bits2 = pin2.to(dev)
        e.set_stream(stream.cuda_stream)
        e.set_genes_device(d_bits2.data_ptr(), G2, N2, W2)
        e.set_trait_vector(0, tr2[0])
        dc2 = torch.empty((G2, 4), dtype=torch.int32, device=dev)
        dp2 = torch.empty(G2, dtype=torch.float64, device=dev)
        for _ in range(3):
            e.contingency_fisher_device(0, dc2.data_ptr(), dp2.data_ptr())
        reps, tot = 20, 0.0
        for _ in range(reps):
            flush.zero_()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            e.contingency_fisher_device(0, dc2.data_ptr(), dp2.data_ptr())
            c1.record()
            torch.cuda.synchronize()
            tot += c0.elapsed_time(c1)
        ms2 = tot / reps
        e.set_stream(0)
        host2 = pin2.numpy().view(np.uint64)
        e.set_genes(host2, N2); e.set_trait_vector(0, tr2[0]); e.contingency_fisher(0)
        t0 = time.perf_counter()
        for _ in range(reps):
            e.set_genes(host2, N2)
            e.set_trait_vector(0, tr2[0])
            cc2, pp2, _ = e.contingency_fisher(0)
        e2e2 = (time.perf_counter() - t0) / reps * 1e3
        tested2 = int(((cc2[:, 0] + cc2[:, 1] > 0) & (cc2[:, 2] + cc2[:, 3] > 0)).sum())
        bytes2 = G2 * (8 * W2 + 24)
        c2 = {"workload": "c2: %d genes x %d isolates x 1 trait, Fisher only" % (G2, N2), "tests_per_step": tested2,
              "value": tested2 / (ms2 * 1e-3), "unit": "tests/s", "ms_per_step": ms2,
              "e2e": {"value": tested2 / (e2e2 * 1e-3), "ms_per_step": e2e2, "h2d_bytes_per_step": int(host2.nbytes + 16 * W2),
                      "d2h_bytes_per_step": int(cc2.nbytes + pp2.nbytes)},
              "roofline": {"bound": "hbm", "achieved": bytes2 / (ms2 * 1e-3) / 1e9, "unit": "GB/s",
                           "note": "1.5 MB of input: launch-latency bound, not bandwidth bound"}}

    if rank != 0:
        if world > 1:
            dist.barrier()              # stay in the group until rank 0 has printed its line
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K5)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    k5_launches = max(1, int(st["launches_permute"]))
    k5_ms = st["ms_permute"] / k5_launches                       # average launch duration (CUDA events, this run)
    launches_per_step = k5_launches / float(a.steps)
    tests_per_launch = G * P / launches_per_step                 # (gene, labelling) walks one launch performs
    n_leaves = N
    bytes_per_test = (8 * W + 8) / float(1 + P) if P > 0 else 0  # compulsory HBM bytes per test
    ops_per_test = (n_leaves - 1) * OPS_PER_NODE                 # int32 add/max ops per test
    alg_bytes = tests_per_launch * bytes_per_test
    alg_ops = tests_per_launch * ops_per_test
    traffic = None
    prof = os.path.join(ROOT, "profiles", "k5_dram_bytes.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(a.workload)
        except Exception:
            traffic = None
    roofline = {"kernel": "walk_permute_kernel (K5 permutation walks)", "bound": "hbm",
                "achieved": alg_bytes / (k5_ms * 1e-3) / 1e9 if P > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                "frac": (alg_bytes / (k5_ms * 1e-3) / 1e9 / hbm_peak) if P > 0 else None, "traffic": traffic,
                "peak_source": peak_src, "ms_per_launch": k5_ms, "launches_per_step": launches_per_step,
                "note": "K5 is integer-issue bound by construction (%.2f algorithmic bytes per test): the meaningful "
                        "bound is roofline_int32.  traffic = ncu dram bytes of ONE launch with caches flushed (the 32 MB "
                        "walk-order gene matrix read once); within a step the %d launches find it in L2 (ncu "
                        "lts hit rate 96 %%)" % (bytes_per_test, int(round(launches_per_step)))}
    roofline_int = {"kernel": "walk_permute_kernel", "bound": "int32 add/max issue (DPX)",
                    "achieved": alg_ops / (k5_ms * 1e-3) / 1e12 if P > 0 else None, "peak": int_peak / 1e12,
                    "unit": "Top/s", "frac": (alg_ops / (k5_ms * 1e-3) / int_peak) if P > 0 else None,
                    "peak_packed16": 2 * int_peak / 1e12,
                    "frac_packed16": (alg_ops / (k5_ms * 1e-3) / (2 * int_peak)) if P > 0 else None,
                    "ops_per_node": OPS_PER_NODE,
                    "peak_source": "sb_int32_peak microbenchmark, this run: VIADDMNMX issue rate x (1 add + 1 max); "
                                   "peak_packed16 = the .S16x2 forms the kernel uses below 128 leaves (2 genes per "
                                   "instruction, same issue rate)",
                    "note": "achieved counts the SURVEY 8(d) contract ops (76 per internal node); the kernel executes "
                            "fewer (cheap leaf/cherry updates, two genes per instruction), so frac can exceed 1"}
    # issue-slot view of the same kernel: warp instructions actually executed (ncu count, profiles/) per second
    # against the SM's issue capacity (4 warp instructions per clock per SM) at the clock sampled above
    try:
        wi_file = json.load(open(os.path.join(ROOT, "profiles", "k5_warp_instructions.json")))
        wi, wi_src = wi_file.get(a.workload), wi_file.get("source", "ncu")
    except Exception:
        wi = wi_src = None
    if wi and P > 0 and clocks and clocks.get("sm_mhz"):
        issued = tests_per_launch / 64.0 * wi / (k5_ms * 1e-3)
        cap = 4.0 * st["sm_count"] * clocks["sm_mhz"] * 1e6
        roofline_int["issue"] = {"warp_instr_per_s": issued, "peak": cap, "frac": issued / cap,
                                 "warp_instr_per_64_walks": wi,
                                 "source": "instruction count from profiles/k5_warp_instructions.json (%s); time and "
                                           "clock from this run" % wi_src}
        alu = (wi_file.get("by_pipe_c3") or {}).get("alu_class")
        if alu:      # ALU-class instructions (DPX, LOP3, SHF, SEL ...) issue at 2 per clock per SM: the tighter pipe bound
            roofline_int["issue"]["alu_class"] = {
                "warp_instr_per_64_walks": alu, "peak_per_clock_per_sm": 2.0,
                "frac": tests_per_launch / 64.0 * alu / (k5_ms * 1e-3) / (2.0 * st["sm_count"] * clocks["sm_mhz"] * 1e6),
                "note": "upper estimate: assumes the DPX forms share the ALU pipe (ncu pipe_alu read 59.9 % where this "
                        "count gave 77 % on the profiled kernel)"}
    fisher_bytes = G * (8 * W + 24)
    fisher = {"kernel": "fisher_kernel (K2+K3)", "ms": fisher_ms, "tests_per_s": G / (fisher_ms * 1e-3),
              "bound": "hbm", "achieved": fisher_bytes / (fisher_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
              "frac": fisher_bytes / (fisher_ms * 1e-3) / 1e9 / hbm_peak}

    cpu = None
    if not a.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
        run, rate, threads, sample = cpu_sample(G, N, T, P, seed, 6.0)
        x, y = run()
        cpu = {"value": rate(x, y), "unit": "tests/s", "cores": threads, "kind": "port", "sample": sample}

    line = {
        "metric": "gene-trait tests/sec (incl. permutations)", "value": value, "unit": "tests/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32 (walk DP) + f64 (Fisher)", "data": "synthetic",
        "config": {"workload": "%s: %d genes/GPU x %d isolates x %d trait(s), %d permutations + pairwise, -p 1.0 "
                               "exhaustive" % (a.workload, G, N, T, P),
                   "parallelism": "gene-sharded x%d, one all-gather" % world if world > 1 else "single GPU",
                   "l2": "256 MiB buffer written between timed steps (inputs ~65 MB < 126 MB L2)",
                   "tests_per_step": tests_per_step, "seed": seed},
        "e2e": {"value": e2e_value, "unit": "tests/s", "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                "ms_per_step": e2e_s / n_e2e * 1e3, "calls_ms": e2e_trace[-1]},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks,
        "roofline": roofline, "roofline_int32": roofline_int, "fisher_pass": fisher, "reference_rule_mode": ref_rule,
        "config1_fisher_only": c2,
        "cpu_baseline": cpu,
        "kernel_ms": {k: st[k] for k in ("ms_fisher", "ms_shuffle", "ms_walk", "ms_permute", "ms_reduce")},
        "wall_s_timed_region": wall,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _guard_stdout():
    """Whatever libraries print (NCCL banners, warnings) goes to stderr; the process's real
    stdout receives exactly the one JSON line (emit())."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    a = parse_args()
    _guard_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
