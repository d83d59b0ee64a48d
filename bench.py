#!/usr/bin/env python
"""bench.py -- gene-trait tests/sec (incl. permutations) of the hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload north_star|c3|c4|c5|c2] [--impl reference]

A "step" is one pass of the whole hot path over one batch of synthetic input: contingency + Fisher
for every (gene, trait) -- every gene row read once for all traits --, the unpermuted pairwise-comparison walk and P
label permutations for every gene (-p 1.0, exhaustive mode: no early stop), then -- for N > 1 -- the one NCCL
all-gather of the per-gene records (scoary_b200.distributed).  tests = G_tested * T * (1 + P).

The job is FIXED and N GPUs split it ("scaling": "strong"): by genes -- contiguous shards, one all-gather of per-gene
records, the reference's own fan-out -- while a shard still fills a GPU, else by permutations -- every GPU walks all
genes under its own range of the labellings and the hit counts are summed in one all-reduce
(scoary_b200.distributed.split_for; --split forces one).  The default workload is BASELINE.json's north_star target,
50 000 genes x 5 000 isolates x 10 000 permutations, at 1 / 2 / 4 / 8 GPUs; --workload c3 / c4 / c5 are configs[2..4]
(c4 and c5 are the 8-GPU configurations; they also fit one GPU).

  value : whole-job tests/s with inputs resident in HBM when the clock starts (CUDA events, max over ranks)
  e2e   : the same through the host-buffer C-ABI calls (pinned host bitsets in, host result arrays out, H2D/D2H
          inside the timed region), the product's gather and rank 0's adjusted p-values + sort; all --steps steps
  roofline : the dominant kernel (K5, permutation walks): the HBM fraction the contract asks for (meaningless for a
          kernel that needs 0.07 bytes per test) and `issue`, its real utilisation (executed warp instructions from
          ncu / issue slots); contract_int32 is the SURVEY 8(d) op-count ratio, labelled as not a bound
  fisher_pass, config1_fisher_only, config2_c3, reference_rule_mode : the other kernels / configs, one GPU only
  cli_wall : the whole command line, file in -> results.csv out, at the C3 shape (tools/cli_wall.py), one GPU only
  cpu_baseline : the oracle's C port on this box's host cores, bounded sample (+ the Python reference's own timing,
          measured in the build container: it cannot travel to the GPU box)
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
    ap.add_argument("--workload", default="north_star", choices=["c2", "c3", "c4", "c5", "north_star"])
    ap.add_argument("--genes", type=int, default=0, help="override the job's gene count")
    ap.add_argument("--isolates", type=int, default=0)
    ap.add_argument("--perms", type=int, default=-1)
    ap.add_argument("--split", default="auto", choices=["auto", "genes", "permutations"],
                    help="how N GPUs divide the job (auto: scoary_b200.distributed.split_for)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli-wall", action="store_true", help="skip the command-line wall-clock sub-line (N = 1 only)")
    return ap.parse_args()


def workload(a):
    from scoary_b200 import synth
    G, N, T, P, seed = synth.CONFIGS[a.workload]       # the whole job; N GPUs split its genes (strong scaling)
    if a.genes:
        G = a.genes
    if a.isolates:
        N = a.isolates
    if a.perms >= 0:
        P = a.perms
    return G, N, T, P, seed


def workload_string(name, G, N, T, P):
    """config.workload -- the same text in both arms"""
    return "%s: %d genes x %d isolates x %d trait(s), %d permutations + pairwise, -p 1.0 exhaustive" % (name, G, N, T, P)


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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32+f64", "data": "synthetic",
        "config": {"workload": workload_string(a.workload, G, N, T, P)},
        "cpu_baseline": {"value": value, "unit": "tests/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of scoary/methods.py + classes.py (oracle/scoary_oracle.c, -O3, OpenMP); "
                                 "the Python reference itself: python_reference",
                         "python_reference": _profile_json("python_reference_timing.json") or None},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def rank0_epilogue(e, p, counts, device_min=None):
    """What rank 0 does with the gathered vector before results can be written: skip rule, Bonferroni,
    Benjamini-Hochberg, p-sort (methods.py:804-814, :900-925, :1448-1454) -- the product's own code path
    (scoary_b200.methods.adjust_pvalues: device epilogue for large vectors, NumPy below)."""
    from scoary_b200 import methods as M
    keep = ((counts[:, 0] + counts[:, 1]) > 0) & ((counts[:, 2] + counts[:, 3]) > 0)
    order, bonf, bh = M.adjust_pvalues(e, p, keep, device_min=device_min)
    return order, bonf, bh


def run_ours(a):
    import torch
    import torch.distributed as dist
    from scoary_b200 import distributed as sbd
    from scoary_b200 import synth
    from scoary_b200 import tree as treemod
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

    # ---- synthetic inputs: ONE fixed job (G genes), split into contiguous gene shards (strong scaling);
    # traits and tree are replicated
    # N GPUs split the job by genes (the reference's own fan-out) while a shard still fills a GPU, else by
    # permutations: every GPU walks all genes under its own range of the labellings and the hit counts are summed
    split = sbd.split_for(G, P, world) if a.split == "auto" else a.split
    if split == "permutations":
        bounds = [(0, G)] * world
        perm_first, perm_count = sbd.permutation_range(P, world, rank)
    else:
        bounds = sbd.shard_bounds(G, world)
        perm_first, perm_count = 0, P
    lo, hi = bounds[rank]
    g_loc = hi - lo
    traits = synth.make_traits(N, T, seed)
    bits_np = synth.make_genes_rows(lo, hi, G, N, seed, traits=traits)
    W = words_for(N)
    pinned = torch.empty((max(g_loc, 1), W), dtype=torch.int64, pin_memory=True)
    pinned.numpy().view(np.uint64)[:g_loc] = bits_np
    del bits_np
    nested = synth.make_tree(N, seed)
    names = synth.isolate_names(N)
    col = {n: j for j, n in enumerate(names)}
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
    e.set_genes_device(d_bits.data_ptr(), g_loc, N, W)
    for t in range(T):
        e.set_trait_vector(t, traits[t])
        if P > 0:
            e.set_tree(t, left, right, leaf_cols)
    d_counts = torch.empty((T, g_loc, 4), dtype=torch.int32, device=dev)
    d_p = torch.empty((T, g_loc), dtype=torch.float64, device=dev)
    d_pairs = torch.zeros((T, g_loc, 3), dtype=torch.int32, device=dev)
    d_r = torch.zeros((T, g_loc), dtype=torch.int32, device=dev)
    d_nd = torch.zeros((T, g_loc), dtype=torch.int32, device=dev)
    RW = sbd.RECORD_WORDS
    d_rec = torch.empty((g_loc, T * RW), dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2
    full_rec = [None]

    def step_device():
        # every gene row is read ONCE for all T traits (sb_contingency_fisher_multi), then walks + permutations
        e.contingency_fisher_multi_device(0, T, d_counts.data_ptr(), d_p.data_ptr())
        if split == "permutations":      # all genes, this rank's range of the labellings; one all-reduce of the hit counts
            for t in range(T):
                e.permute_range_device(t, g_loc, perm_first, perm_count, seed, d_pairs[t].data_ptr(), d_r[t].data_ptr())
            dist.all_reduce(d_r, op=dist.ReduceOp.SUM)
            return
        if P > 0:
            for t in range(T):
                e.permute_device(t, g_loc, P, seed, d_pairs[t].data_ptr(), d_r[t].data_ptr(), d_nd[t].data_ptr())
        if world > 1:   # the one collective of the path: an all-gather of fixed-size per-gene records (NCCL)
            for t in range(T):
                d_rec[:, t * RW + 0:t * RW + 4] = d_counts[t]
                d_rec[:, t * RW + 4:t * RW + 6] = d_p[t].view(torch.int32).view(g_loc, 2)
                d_rec[:, t * RW + 6:t * RW + 9] = d_pairs[t]
                d_rec[:, t * RW + 9] = d_r[t]
                d_rec[:, t * RW + 10] = d_nd[t]
            full_rec[0] = sbd.all_gather_records(d_rec, G, bounds)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(a.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    counts_host = d_counts.cpu().numpy()
    g_tested = int(((counts_host[..., 0] + counts_host[..., 1] > 0) &
                    (counts_host[..., 2] + counts_host[..., 3] > 0)).sum())
    # the job's tests are counted once: a rank that walks a range of the permutations repeats the (cheap) Fisher pass
    # and the unpermuted walk of every gene, which are not counted again
    tests_per_step_rank = g_tested * (1 + P) if split == "genes" else g_tested * perm_count + (g_tested if rank == 0 else 0)

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

    # ---- e2e: the calls a user of the C-ABI makes, host buffers in and out (copies inside the timed region), the
    # product's own gather (scoary_b200.distributed, NCCL) and rank 0's epilogue (adjusted p-values + sort)
    bits_host = pinned.numpy().view(np.uint64)[:g_loc]
    e2e_trace = []

    def step_e2e():
        h2d = d2h = 0
        t0 = time.perf_counter()
        e.set_genes(bits_host, N)
        h2d += bits_host.nbytes
        for t in range(T):
            e.set_trait_vector(t, traits[t])
            h2d += 2 * W * 8
            if P > 0:          # a Fisher-only job (C2) has no tree
                e.set_tree(t, left, right, leaf_cols)
                h2d += left.nbytes + right.nbytes + leaf_cols.nbytes
        t1 = time.perf_counter()
        c, p, _ = e.contingency_fisher_multi(0, T)
        d2h += c.nbytes + p.nbytes
        t2 = time.perf_counter()
        rec = np.zeros((g_loc, T * RW), dtype=np.int32)
        r_all = np.zeros((T, g_loc), dtype=np.int32)
        for t in range(T):
            pairs = r = nd = 0
            if split == "permutations":
                pairs, r = e.permute_range(t, perm_first, perm_count, seed=seed)
                d2h += pairs.nbytes + r.nbytes
                r_all[t], nd = r, P
            elif P > 0:
                pairs, r, nd = e.permute(t, P, seed=seed)
                d2h += pairs.nbytes + r.nbytes + nd.nbytes
            rec[:, t * RW:(t + 1) * RW] = sbd.pack_records(c[t], p[t], pairs, r, nd)
        t3 = time.perf_counter()
        if split == "permutations":
            r_all = sbd.all_reduce_sum(r_all)
            for t in range(T):
                rec[:, t * RW + 9] = r_all[t]
        elif world > 1:
            rec = sbd.gather_blocks(rec, G)
        t4 = time.perf_counter()
        checksum = 0.0
        if rank == 0:
            for t in range(T):
                u = sbd.unpack_records(rec[:, t * RW:(t + 1) * RW])
                order, bonf, bh = rank0_epilogue(e, u["p"], u["counts"])
                checksum += float(bh[order[0]]) if len(order) else 0.0
        t5 = time.perf_counter()
        e2e_trace.append({"set_inputs_ms": (t1 - t0) * 1e3, "fisher_ms": (t2 - t1) * 1e3, "permute_ms": (t3 - t2) * 1e3,
                          "gather_ms": (t4 - t3) * 1e3, "rank0_epilogue_ms": (t5 - t4) * 1e3,
                          "total_ms": (t5 - t0) * 1e3})
        return h2d, d2h

    e.set_stream(0)
    step_e2e()
    barrier()
    torch.cuda.synchronize()
    e0 = time.perf_counter()
    for _ in range(a.steps):
        h2d_b, d2h_b = step_e2e()
    e.synchronize()
    barrier()
    e2e_s = time.perf_counter() - e0
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    e2e_value = tests_per_step * a.steps / e2e_s

    # ---- sub-measurements that explain the line: one GPU only (they would only repeat per rank)
    ref_rule = fisher = c2 = c3 = None
    if world == 1:
        e.set_stream(stream.cuda_stream)
        e.set_genes_device(d_bits.data_ptr(), g_loc, N, W)
        for t in range(T):
            e.set_trait_vector(t, traits[t])
            if P > 0:
                e.set_tree(t, left, right, leaf_cols)
        # reference-rule mode (second number): the reference's sequential early stop (methods.py:1360-1363)
        if P >= 32:
            from scoary_b200.methods import early_stop_table
            d_rmin = torch.from_numpy(early_stop_table(P)).to(dev)

            def step_rule():
                for t in range(T):
                    e.permute_device(t, g_loc, P, seed, d_pairs[t].data_ptr(), d_r[t].data_ptr(), d_nd[t].data_ptr(),
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
            e.stats_reset()
            e.set_profiling(True)          # a third pass with per-kernel events: where the time of a pass goes
            step_rule()
            rst = e.stats()
            e.set_profiling(False)
            nd = d_nd.cpu().numpy()
            ref_rule = {"ms_per_step": rule_ms, "walks_executed": int(walks), "walks_exhaustive": int(g_loc * T * (1 + P)),
                        "genes_stopped_early": int((nd < P).sum()), "genes": int(g_loc * T),
                        "equivalent_tests_per_s": g_loc * T * (1 + P) / (rule_ms * 1e-3),
                        "kernel_ms": {k: rst[k] for k in ("ms_shuffle", "ms_walk", "ms_permute", "ms_reduce")},
                        "k5_launches": int(rst["launches_permute"]),
                        "note": "Permute's early stop on; pairwise walk + permutations only (no Fisher pass)"}
        # the Fisher pass alone (all T traits in one launch), device resident
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e.contingency_fisher_multi_device(0, T, d_counts.data_ptr(), d_p.data_ptr())
        flush.zero_()
        f0.record()
        e.contingency_fisher_multi_device(0, T, d_counts.data_ptr(), d_p.data_ptr())
        f1.record()
        torch.cuda.synchronize()
        fisher_ms = f0.elapsed_time(f1)
        fisher_bytes = g_loc * (8 * W + 24 * T) + 16 * W * T      # each row once, 24 B out per (gene, trait)
        peaks = _peaks()
        fisher = {"kernel": "fisher_kernel (K2+K3), %d trait(s) per row read" % T, "ms": fisher_ms,
                  "tests_per_s": g_loc * T / (fisher_ms * 1e-3), "bound": "hbm",
                  "achieved": fisher_bytes / (fisher_ms * 1e-3) / 1e9, "peak": peaks[0], "unit": "GB/s",
                  "frac": fisher_bytes / (fisher_ms * 1e-3) / 1e9 / peaks[0], "bytes_per_test": fisher_bytes / (g_loc * T)}
        if a.workload == "north_star":      # before the c2 line: that one loads another gene matrix (and drops the trees)
            c3 = c3_line(e, torch, dev, stream, flush, d_bits, g_loc, N, W, seed, d_counts, d_p, d_pairs, d_r, d_nd, g_tested)
        if a.workload != "c2":
            c2 = small_fisher_line(e, torch, dev, stream, flush, synth, words_for)

    if rank != 0:
        if world > 1:
            dist.barrier()              # stay in the group until rank 0 has printed its line
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (K5)
    hbm_peak, peak_src = _peaks()
    k5_launches = max(1, int(st["launches_permute"]))
    k5_ms = st["ms_permute"] / k5_launches                       # average launch duration (CUDA events, this run)
    launches_per_step = k5_launches / float(a.steps)
    tests_per_launch = g_loc * T * perm_count / launches_per_step    # (gene, labelling) walks one launch performs
    bytes_per_test = (8 * W + 8) / float(1 + P) if P > 0 else 0  # compulsory HBM bytes per test
    ops_per_test = (N - 1) * OPS_PER_NODE                        # int32 add/max ops per test
    alg_bytes = tests_per_launch * bytes_per_test
    alg_ops = tests_per_launch * ops_per_test
    prof = _profile_json("k5_dram_bytes.json")
    traffic = prof.get(a.workload) or prof.get("c3")
    roofline = {"kernel": "walk_permute_kernel (K5 permutation walks)", "bound": "hbm",
                "achieved": alg_bytes / (k5_ms * 1e-3) / 1e9 if P > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                "frac": (alg_bytes / (k5_ms * 1e-3) / 1e9 / hbm_peak) if P > 0 else None, "traffic": traffic,
                "peak_source": peak_src, "ms_per_launch": k5_ms, "launches_per_step": launches_per_step,
                "note": "K5 is integer-issue bound by construction (%.2f algorithmic bytes per test): this HBM fraction "
                        "says nothing about it, `issue` below is the utilisation figure.  traffic = dram read + write "
                        "bytes of one full-size launch inside a running step (ncu --cache-control none, %s).  It is far "
                        "above the algorithmic bytes and it is not re-reads of the input: the %d MB walk-order gene matrix "
                        "stays in L2; what goes to DRAM is the threads' local-memory DP stack (32-bit entries of the "
                        "tree's spine + register spills, ~0.7 KB per thread, rewritten by every block), which L1 / L2 write "
                        "back.  At ~40 GB/s it is 0.6 %% of the HBM peak and costs the kernel nothing"
                        % (bytes_per_test, prof.get("source", "profiles/"), int(g_loc * W * 8 / 1e6))}
    # the utilisation figure of K5: warp instructions it executes (ncu smsp__inst_executed.sum, profiles/) per second
    # against the SM's issue capacity (4 warp instructions per clock per SM) at the clock sampled in this run
    wi_file = _profile_json("k5_warp_instructions.json")
    key = a.workload if a.workload in wi_file else ("c3" if N == 5000 else None)
    wi = wi_file.get(key) if key else None
    issue = None
    if wi and P > 0 and clocks and clocks.get("sm_mhz"):
        issued = tests_per_launch / 64.0 * wi / (k5_ms * 1e-3)
        cap = 4.0 * st["sm_count"] * clocks["sm_mhz"] * 1e6
        issue = {"bound": "warp-instruction issue slots (4 per clock per SM)", "warp_instr_per_s": issued, "peak": cap,
                 "frac": issued / cap, "warp_instr_per_64_walks": wi,
                 "source": "%s; time and clock from this run" % wi_file.get("source", "ncu")}
        alu = (wi_file.get("by_pipe_" + key) or {}).get("alu_pct_of_peak")
        if alu:
            issue["alu_pipe_pct_of_peak_ncu"] = alu
    roofline["issue"] = issue
    contract = {"what": "SURVEY 8(d) contract figure: (N - 1) x 76 int32 add/max ops per walk / the DPX rate measured by "
                        "sb_int32_peak in this run.  NOT a bound for this kernel: it executes about 1/5.7 of those "
                        "operations (one key pass, two genes per .S16x2 instruction, fused add-max, cheap leaf and cherry "
                        "updates), so the ratio exceeds 1",
                "contract_ops_per_s": alg_ops / (k5_ms * 1e-3) if P > 0 else None, "dpx_peak_ops_per_s": int_peak,
                "ratio": (alg_ops / (k5_ms * 1e-3) / int_peak) if P > 0 else None, "ops_per_node": OPS_PER_NODE}

    cpu = None
    if not a.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
        run, rate, threads, sample = cpu_sample(G, N, T, P, seed, 6.0)
        x, y = run()
        cpu = {"value": rate(x, y), "unit": "tests/s", "cores": threads, "kind": "port", "sample": sample,
               "python_reference": _profile_json("python_reference_timing.json") or None}

    # the whole command line, file in -> results.csv out, at the C3 shape (tools/cli_wall.py in its own process, after
    # every timed region); the C5-shaped run (4 GB of input) is quoted from profiles/
    cli_wall = None
    if world == 1 and not a.no_cli_wall and not a.no_cpu_baseline:
        try:
            res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cli_wall.py"), "--shape", "c3"],
                                 capture_output=True, text=True, timeout=600)
            cli_wall = {"c3": json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]),
                        "c5": _profile_json("r2_cli_wall_c5.json") or None,
                        "note": "scoary_b200.methods.main on a synthetic Roary table, -p 1.0 -c I -e 1000: parse + pack, "
                                "UPGMA tree, statistics, pairwise + permutations (reference-rule mode), result file; c3 "
                                "measured in this run, c5 (1 000 000 x 2 000, 4 GB of CSV) from profiles/"}
        except Exception as ex:      # the sub-line is informative only
            cli_wall = {"error": str(ex)[:200]}

    line = {
        "metric": "gene-trait tests/sec (incl. permutations)", "value": value, "unit": "tests/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": dev_ms / a.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32 (walk DP) + f64 (Fisher)", "data": "synthetic",
        "config": {"workload": workload_string(a.workload, G, N, T, P),
                   "parallelism": "single GPU" if world == 1 else (
                       "the job's %d genes in %d contiguous shards (%d per GPU), traits and tree replicated, one NCCL "
                       "all-gather of %d-byte per-gene records (scoary_b200.distributed)"
                       % (G, world, bounds[0][1] - bounds[0][0], 4 * RW * T) if split == "genes" else
                       "the job's %d permutations in %d ranges (%d per GPU): every GPU walks all %d genes under its own "
                       "labellings (gene shards of %d would not fill a GPU), Fisher pass and unpermuted walk replicated, "
                       "one NCCL all-reduce of the %d-byte hit-count vector (scoary_b200.distributed.split_for)"
                       % (P, world, perm_count, G, G // world, 4 * G * T)),
                   "l2": "256 MiB buffer written between timed steps (inputs %d MB < 126 MB L2)" % (2 * g_loc * W * 8 // 1000000),
                   "value_includes": "Fisher pass, pairwise walk, label shuffles, permutation walks, hit bookkeeping%s; the "
                                     "gene matrix is already in walk order (K1 pack + host tree compile run once, in warm-up; "
                                     "e2e repeats them every step)" % ((", all-gather" if split == "genes" else ", all-reduce") if world > 1 else ""),
                   "tests_per_step": tests_per_step, "seed": seed},
        "e2e": {"value": e2e_value, "unit": "tests/s", "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                "ms_per_step": e2e_s / a.steps * 1e3, "steps": a.steps, "calls_ms": e2e_trace[-1],
                "includes": "host bitsets in, host result arrays out, the gather and rank 0's adjusted p-values + sort"},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks,
        "roofline": roofline, "contract_int32": contract, "fisher_pass": fisher, "reference_rule_mode": ref_rule,
        "config1_fisher_only": c2, "config2_c3": c3, "cli_wall": cli_wall,
        "cpu_baseline": cpu,
        "kernel_ms": {k: st[k] for k in ("ms_fisher", "ms_shuffle", "ms_walk", "ms_permute", "ms_reduce")},
        "wall_s_timed_region": wall,
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            return json.load(fh)
    except Exception:
        return {}


def _peaks():
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    return (float(peaks.get("hbm_gbs", 6650.0)),
            "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)")


def small_fisher_line(e, torch, dev, stream, flush, synth, words_for):
    """BASELINE configs[1] itself (10k genes x 1k isolates, Fisher only, no tree): a second, small line."""
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
    e.set_genes(host2, N2)
    e.set_trait_vector(0, tr2[0])
    e.contingency_fisher(0)
    t0 = time.perf_counter()
    for _ in range(reps):
        e.set_genes(host2, N2)
        e.set_trait_vector(0, tr2[0])
        cc2, pp2, _ = e.contingency_fisher(0)
    e2e2 = (time.perf_counter() - t0) / reps * 1e3
    tested2 = int(((cc2[:, 0] + cc2[:, 1] > 0) & (cc2[:, 2] + cc2[:, 3] > 0)).sum())
    bytes2 = G2 * (8 * W2 + 24)
    hbm_peak, _ = _peaks()
    return {"workload": "c2: %d genes x %d isolates x 1 trait, Fisher only" % (G2, N2), "tests_per_step": tested2,
            "value": tested2 / (ms2 * 1e-3), "unit": "tests/s", "ms_per_step": ms2,
            "e2e": {"value": tested2 / (e2e2 * 1e-3), "ms_per_step": e2e2, "h2d_bytes_per_step": int(host2.nbytes + 16 * W2),
                    "d2h_bytes_per_step": int(cc2.nbytes + pp2.nbytes)},
            "roofline": {"bound": "hbm", "achieved": bytes2 / (ms2 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes2 / (ms2 * 1e-3) / 1e9 / hbm_peak,
                         "note": "1.5 MB of input, one wave of one batch per warp: launch + FP64 latency, not bandwidth"}}


def c3_line(e, torch, dev, stream, flush, d_bits, G, N, W, seed, d_counts, d_p, d_pairs, d_r, d_nd, g_tested):
    """BASELINE configs[2] (the round-1 headline): the same matrix with 1 000 permutations, three timed steps."""
    P3 = 1000
    e.set_stream(stream.cuda_stream)

    def step():
        e.contingency_fisher_multi_device(0, 1, d_counts.data_ptr(), d_p.data_ptr())
        e.permute_device(0, G, P3, seed, d_pairs[0].data_ptr(), d_r[0].data_ptr(), d_nd[0].data_ptr())
    step()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(3):
        flush.zero_()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        step()
        c1.record()
        torch.cuda.synchronize()
        tot += c0.elapsed_time(c1)
    ms = tot / 3
    return {"workload": workload_string("c3", G, N, 1, P3), "ms_per_step": ms, "value": g_tested * (1 + P3) / (ms * 1e-3),
            "unit": "tests/s", "steps": 3}


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
