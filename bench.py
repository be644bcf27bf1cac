#!/usr/bin/env python3
"""bench.py -- read pairs/s of paired-end link inference on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path (record split -> 2-bit pack/lookup -> link keys ->
counted matrices, + one allreduce when N > 1) over one batch of synthetic reads of the named
config.  ``value`` times it with the FASTQ bytes already resident in HBM; ``e2e`` times the
same call from pinned HOST buffers (H2D inside) plus the D2H read of the matrices.
Weak scaling: every rank processes its own full-size batch (same graph, rank-specific reads).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vstrains_b200 import synth  # noqa: E402

METRIC = "read_pairs_per_s_pe_link_inference"
UNIT = "pairs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_workload(cfg_name: str, pairs: int, rank: int):
    """Graph from the config seed (identical on every rank); reads from a rank-specific stream."""
    cfg = synth.CONFIGS[cfg_name]
    rng = np.random.default_rng(cfg.seed)
    depth = pairs * 2.0 * cfg.read_len / cfg.genome_len / max(cfg.n_genomes, 1)
    g, genomes, ab = synth.make_graph(cfg, rng, depth)
    rrng = np.random.default_rng([cfg.seed, 7919, rank])
    fs, rs = [], []
    done = 0
    while done < pairs:
        n = min(250_000, pairs - done)
        f, r = synth.make_reads(genomes, ab, cfg.read_len, n, cfg.k, rrng, first_idx=done)
        fs.append(f)
        rs.append(r)
        done += n
    return cfg, g, np.concatenate(fs), np.concatenate(rs)


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arms (oracle port of the reference's Python hash-table path)
# ------------------------------------------------------------------------------------------
_W = {}


def _worker_init(gfa: bytes, k: int):
    from oracle import pe_oracle
    ids, seqs = pe_oracle.parse_gfa(gfa)
    _W["table"] = pe_oracle.build_index(seqs, k + 1)
    _W["lens"] = [len(s) for s in seqs]
    _W["k"] = k


def _worker_run(args):
    from oracle import pe_oracle
    fwd, rve = args
    node, short, stats = pe_oracle.count_pairs(pe_oracle.split_lines(fwd), pe_oracle.split_lines(rve),
                                               _W["table"], _W["lens"], _W["k"] + 1)
    return stats["total_pairs"], sum(node.values()) + sum(short.values())


def cpu_port_rate(gfa: bytes, f: np.ndarray, r: np.ndarray, k: int, n_proc: int, pairs_per_proc: int, steps: int, warmup: int):
    """Python port of the reference path on n_proc host cores: pairs/s over `steps` timed samples.
    The index is built once per worker before timing (the reference amortises it over the file)."""
    import multiprocessing as mp
    from vstrains_b200 import shard
    total = pairs_per_proc * n_proc
    rng_f = shard.line_ends(f[: min(f.size, (total * steps + total) * 700)])
    rng_r = shard.line_ends(r[: min(r.size, (total * steps + total) * 700)])
    have = min(rng_f.size, rng_r.size) // 4

    def pieces(step):
        out = []
        for p in range(n_proc):
            a = ((step * n_proc + p) * pairs_per_proc) % max(1, have - pairs_per_proc)
            b = a + pairs_per_proc
            lo_f = 0 if a == 0 else int(rng_f[4 * a - 1])
            lo_r = 0 if a == 0 else int(rng_r[4 * a - 1])
            out.append((f[lo_f:int(rng_f[4 * b - 1])].tobytes(), r[lo_r:int(rng_r[4 * b - 1])].tobytes()))
        return out

    ctx = mp.get_context("fork")
    with ctx.Pool(n_proc, initializer=_worker_init, initargs=(gfa, k)) as pool:
        for s in range(warmup):
            pool.map(_worker_run, pieces(s))
        times = []
        for s in range(steps):
            pc = pieces(warmup + s)
            t0 = time.perf_counter()
            res = pool.map(_worker_run, pc)
            times.append(time.perf_counter() - t0)
            assert sum(x[0] for x in res) == total
    return total / statistics.mean(times), statistics.mean(times)


def c_port_rate(gfa: bytes, f: np.ndarray, r: np.ndarray, k: int, pairs: int):
    from oracle import c_oracle
    from vstrains_b200 import shard
    ef, er = shard.line_ends(f[: pairs * 700]), shard.line_ends(r[: pairs * 700])
    fb, rb = f[: int(ef[4 * pairs - 1])], r[: int(er[4 * pairs - 1])]
    t0 = time.perf_counter()
    c_oracle.run(gfa, fb, rb, k, 0)
    dt = time.perf_counter() - t0
    return pairs / dt


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(synth.CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (default: the config's, capped)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-generic", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (experiments)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg0 = synth.CONFIGS[args.config]
    cap = {"C1": 100_000, "C2": 1_000_000}.get(args.config, 2_000_000)   # per-GPU batch of the config
    pairs = args.pairs or min(cfg0.pairs, cap)
    workload = "%s: %d-strain %d bp graph (k=%d), %d x 2x%d bp read pairs per GPU per step" % (
        cfg0.name, cfg0.strains, cfg0.genome_len, cfg0.k, pairs, cfg0.read_len)

    if args.impl == "reference":
        if rank != 0:
            return
        n_proc = os.cpu_count() or 1
        cfg, g, f, r = make_workload(args.config, min(pairs, 400_000), 0)
        gfa = g.to_gfa()
        ppp = 4000
        rate, dt = cpu_port_rate(gfa, f, r, cfg.k, n_proc, ppp, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": workload, "graph_nodes": len(g.ids)},
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": n_proc, "kind": "port",
                                 "sample": "python port of the reference hash-table path (oracle/pe_oracle.py), %d processes x %d pairs per step, index prebuilt" % (n_proc, ppp)},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    lib_path = os.path.join(ROOT, "vstrains_b200", "libvspe.so")
    if not os.path.exists(lib_path):               # clean checkout: compile the CUDA library first
        if local_rank == 0:
            import __graft_entry__
            __graft_entry__.build()
        else:
            t_wait = time.time()
            while not os.path.exists(lib_path) and time.time() - t_wait < 600:
                time.sleep(1.0)
            time.sleep(2.0)
    from vstrains_b200 import pe_inference

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    cfg, g, f, r = make_workload(args.config, pairs, rank)
    gfa = g.to_gfa()
    n_nodes = len(g.ids)
    bytes_step = int(f.size + r.size)
    b_pair = bytes_step / pairs
    ix = pe_inference.PEIndex([bytes(s) for s in g.seqs], cfg.k, device=local_rank)
    ix.set_option("force_generic", args.force_generic)
    for kv in args.opt:
        k_, v_ = kv.split("=")
        ix.set_option(k_, int(v_))

    class _Arr:                                    # torch view of the library's matrices
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
    sparse = ix.is_sparse                          # graphs too large for N*N matrices (C5) keep sorted runs
    if sparse:
        mats = None
    else:
        mptr, mn = ix.matrices_device()
        mats = torch.as_tensor(_Arr(mptr, mn), device=dev) if mn else torch.zeros(0, dtype=torch.int64, device=dev)

    def merge_sparse():
        """one exchange step: every rank's runs are gathered on rank 0 and merged there"""
        keys, counts = ix.sparse()
        sizes = [None] * world
        dist.all_gather_object(sizes, int(keys.size))
        cap = max(max(sizes), 1)
        buf = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
        buf[: keys.size] = torch.from_numpy(keys.view(np.int64)).to(dev)
        buf[cap: cap + keys.size] = torch.from_numpy(counts.view(np.int64)).to(dev)
        gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, gathered, dst=0)
        if rank == 0:
            for src in range(1, world):
                g_ = gathered[src].cpu().numpy()
                ix.sparse_merge(g_[: sizes[src]].view(np.uint64), g_[cap: cap + sizes[src]].view(np.uint64))

    d_f = torch.from_numpy(f).to(dev)
    d_r = torch.from_numpy(r).to(dev)
    torch.cuda.synchronize()

    def step_device():
        ix.reset()
        ix.count_device(d_f.data_ptr(), f.size, d_r.data_ptr(), r.size)
        if world > 1 and sparse:
            merge_sparse()
        elif world > 1:
            dist.all_reduce(mats)
            torch.cuda.synchronize()       # the library's stream does not order with torch's

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------
    sampler = ClockSampler(local_rank)           # samples across warm-up + timed region (steps are ms-short)
    sampler.start()
    for _ in range(warmup):
        step_device()
    barrier()
    # keep the GPU under load for >= 1 s before timing so the clock samples mean something; the
    # extra step count is decided on rank 0 and broadcast (every rank must run the same number
    # of collectives)
    t_w = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    extra = torch.tensor([max(0, min(2000, int(1.0 / max(time.perf_counter() - t_w, 1e-4))))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(extra, src=0)
    for _ in range(int(extra.item())):
        step_device()
    barrier()
    stage = {"ms_scan": 0.0, "ms_map": 0.0, "ms_count": 0.0, "ms_total": 0.0, "ms_k_scan_pack": 0.0, "ms_k_scan_count": 0.0}
    n_scan_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    launches = 0
    for _ in range(args.steps):
        step_device()
        st = ix.stats()
        for k in stage:
            stage[k] += st[k]
        n_scan_launches += st["n_k_scan_pack"]
        launches += st["kernel_launches"]
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * pairs * args.steps / (ms_max * 1e-3)
    st = ix.stats()

    # ---- end to end: pinned host buffers -> matrices on the host ------------------------
    h_f = torch.from_numpy(f).pin_memory()
    h_r = torch.from_numpy(r).pin_memory()

    def step_e2e():
        ix.reset()
        ix.count_host_ptr(h_f.data_ptr(), f.size, h_r.data_ptr(), r.size)
        if world > 1 and sparse:
            merge_sparse()
        elif world > 1:
            dist.all_reduce(mats)
            torch.cuda.synchronize()
        return ix.sparse() if sparse else ix.matrices()

    # the end-to-end ceiling: a plain pinned host->device copy of the same bytes (PCIe), for context
    h2d_gbs = None
    try:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            ev0.record()
            d_f.copy_(h_f, non_blocking=True)
            d_r.copy_(h_r, non_blocking=True)
            ev1.record()
            torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
        h2d_gbs = bytes_step / (best * 1e-3) / 1e9
    except Exception:
        pass

    for _ in range(2):
        step_e2e()
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * pairs * e2e_steps / float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # dominant kernel: the pack pass k_scan_pack<2> streams every algorithmic byte (one launch per
    # mate file) and is the longest single kernel of the step (profiles/).  Its duration comes from
    # CUDA events recorded around the launch on the library's own stream, over the timed region.
    # The count pass k_scan_pack<1> that precedes it reads the same bytes once more; `scan_both_passes`
    # reports the pair together so the figure stays comparable with the fused look-back scan
    # (--opt scan_mode=3), where ms_k_scan_count is 0.
    k_ms = stage["ms_k_scan_pack"] / max(1, n_scan_launches)            # average launch duration
    k_bytes = bytes_step / 2.0                                          # algorithmic bytes per launch (one mate)
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_scan_pack_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
        traffic = k_bytes * tj["dram_bytes_per_algorithmic_byte"]          # from the committed ncu --set full capture
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "graph_nodes": n_nodes, "bytes_per_pair": b_pair,
                   "l2": "inputs (%.0f MB per GPU) larger than the 126 MB L2; no flush needed" % (bytes_step / 1e6),
                   "index_build_ms": st["ms_index"], "keys_per_pair": st["n_keys"] / max(1, st["used_pairs"]),
                   "reads_fast": st["reads_fast"], "reads_generic": st["reads_generic"],
                   "counting": "sparse runs (LSD radix sort + RLE)" if sparse else "dense matrices (radix partition + counting sort)"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_step,
                "d2h_bytes_per_step": int(16 * ix.sparse()[0].size) if sparse else int(2 * n_nodes * n_nodes * 8),
                "pinned_h2d_copy_gbs": h2d_gbs,
                "frac_of_h2d_copy": (e2e_value / world * b_pair / 1e9 / h2d_gbs) if h2d_gbs else None},
        "stages_ms_per_step": {k: v / args.steps for k, v in stage.items()},
        "whole_job_hbm_frac": value / world * b_pair / 1e9 / peak,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_scan_pack (K1+K2 pack pass: TMA tile -> read table -> 2-bit rows)", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": k_bytes, "launch_ms": k_ms, "launches_per_step": n_scan_launches / args.steps,
                     "kernel_share_of_step": stage["ms_k_scan_pack"] / max(1e-9, stage["ms_total"]),
                     "scan_both_passes": {"launch_ms": k_ms + stage["ms_k_scan_count"] / max(1, n_scan_launches),
                                          "achieved": k_bytes / max(1e-9, (k_ms + stage["ms_k_scan_count"] / max(1, n_scan_launches)) * 1e-3) / 1e9}},
    }
    if world == 1 and not args.no_cpu_baseline:
        n = 40000
        rate, _ = cpu_port_rate(gfa, f, r, cfg.k, 1, n, 1, 0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "python port of the reference path (oracle/pe_oracle.py), first %d pairs of the same workload, 1 process, index prebuilt" % n,
                                "c_port_all_cores": {"value": c_port_rate(gfa, f, r, cfg.k, 100_000), "cores": os.cpu_count(),
                                                     "sample": "oracle/pe_oracle.c with OpenMP on the first 100000 pairs, index build included"}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
