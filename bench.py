#!/usr/bin/env python3
"""bench.py -- read pairs/s of paired-end link inference on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the whole hot path (record split -> 2-bit pack -> lookup / walk -> link
keys -> counted matrices, + one allreduce when N > 1) over the named config's read pairs.  The
default is C4, the largest single-GPU config of BASELINE.json (50 M 2x150 pairs, 31.8 GB of
FASTQ): a unique 10 M-pair block (6.36 GB, far larger than the 126 MB L2) is resident per GPU and
replayed 5 times per step -- the counts of a step are exactly 5x the block's.
``value`` times it with the FASTQ bytes already resident in HBM; ``e2e`` times the same work
through the host-buffer entry point (pinned H2D of every replay inside) plus the D2H read of the
matrices.  Weak scaling: every rank processes its own full-size batch (same graph, rank-specific
reads).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import synthgen as synth  # noqa: E402

METRIC = "read_pairs_per_s_pe_link_inference"
UNIT = "pairs/s"
REF_SCRIPT = os.path.join(ROOT, "baseline", "_ref", "VStrains_PE_Inference.py")
REF_PROCS = 16                     # fixed, so that the N=1 numbers of separate runs agree
BLOCK_PAIRS = {"C1": 100_000, "C2": 1_000_000, "C3": 10_000_000, "C4": 10_000_000, "C5": 12_500_000}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_graph(cfg_name: str, pairs: int):
    """Graph + strain genomes of a config (identical on every rank).  Under torchrun the multi-genome stress
    graph is built once, by local rank 0, and handed to the other ranks through /dev/shm."""
    import pickle
    cfg = synth.CONFIGS[cfg_name]
    depth = pairs * 2.0 * cfg.read_len / cfg.genome_len / max(cfg.n_genomes, 1)
    world, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    share = world > 1 and cfg.n_genomes > 1
    path = "/dev/shm/vspe_graph_%s_%d_%s.pkl" % (cfg_name, pairs, os.environ.get("MASTER_PORT", "0"))
    if share and local != 0:
        t0 = time.time()
        while not os.path.exists(path + ".done") and time.time() - t0 < 900:
            time.sleep(0.5)
        with open(path, "rb") as fh:
            g, genomes, ab = pickle.load(fh)
        return cfg, g, genomes, ab
    g, genomes, ab = synth.make_graph(cfg, np.random.default_rng(cfg.seed), depth)
    if share:
        with open(path, "wb") as fh:
            pickle.dump((g, genomes, ab), fh, protocol=4)
        open(path + ".done", "w").close()
    return cfg, g, genomes, ab


def make_reads(cfg, genomes, ab, pairs: int, rank: int, out=None):
    """Rank-specific read stream from the counter-based generator (synthgen/fastq_gen.c)."""
    return synth.make_reads_fast(genomes, ab, cfg.read_len, pairs, cfg.k, seed=cfg.seed * 1_000_003 + 7919 * rank, out=out)


def make_workload(cfg_name: str, pairs: int, rank: int):
    cfg, g, genomes, ab = make_graph(cfg_name, pairs)
    f, r = make_reads(cfg, genomes, ab, pairs, rank)
    return cfg, g, f, r


def prefix_pairs(f: np.ndarray, r: np.ndarray, first: int, n: int, rl: int):
    """Byte ranges of pairs [first, first + n) of a generated block (record-aligned)."""
    from vstrains_b200 import shard
    span = (first + n + 16) * (2 * rl + 18)
    ef, er = shard.line_ends(f[:span]), shard.line_ends(r[:span])
    lo_f = 0 if first == 0 else int(ef[4 * first - 1])
    lo_r = 0 if first == 0 else int(er[4 * first - 1])
    return f[lo_f:int(ef[4 * (first + n) - 1])], r[lo_r:int(er[4 * (first + n) - 1])]


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
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self, t_lo=None, t_hi=None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for t, r in self.rows:
            if t_lo is not None and not (t_lo <= t <= t_hi + 0.15):
                continue
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
# CPU arms: the UNMODIFIED reference script (baseline/_ref, copied there by __graft_entry__.build
# when /root/reference is present) and, for context, the oracle ports
# ------------------------------------------------------------------------------------------
def run_ref_shards(gfa_path: str, shards, k: int, work: str):
    """Run the unmodified script on every (fwd_bytes, rve_bytes) shard concurrently, one process
    (one core) each; returns the wall time of the slowest."""
    procs = []
    for i, (fb, rb) in enumerate(shards):
        d = os.path.join(work, "s%d" % i)
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "f.fq"), "wb") as fh:
            fh.write(fb)
        with open(os.path.join(d, "r.fq"), "wb") as fh:
            fh.write(rb)
    t0 = time.perf_counter()
    for i in range(len(shards)):
        d = os.path.join(work, "s%d" % i)
        procs.append(subprocess.Popen([sys.executable, REF_SCRIPT, "-g", gfa_path, "-o", os.path.join(d, "out"),
                                       "-f", os.path.join(d, "f.fq"), "-r", os.path.join(d, "r.fq"), "-k", str(k)],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
    for p in procs:
        _, err = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference script failed: %s" % err.decode()[-400:])
    dt = time.perf_counter() - t0
    for i in range(len(shards)):
        shutil.rmtree(os.path.join(work, "s%d" % i, "out"), ignore_errors=True)
    return dt


def ref_script_rate(cfg, gfa: bytes, f, r, n_proc: int, pairs_per_proc: int, steps: int):
    """pairs/s of the unmodified reference script on n_proc cores.  Every run pays the script's fixed
    cost (index build + two N*N-line files, reference :117-135, :190-207) -- about 25 s at C4's N --
    which a 50 M-pair job amortises to nothing, so it is measured by a 0-read run of the same
    n_proc processes and reported next to the raw wall time; `value` uses the per-pair part."""
    work = tempfile.mkdtemp(prefix="vspe_ref_")
    try:
        gfa_path = os.path.join(work, "g.gfa")
        with open(gfa_path, "wb") as fh:
            fh.write(gfa)
        t_fixed = run_ref_shards(gfa_path, [(b"", b"")] * n_proc, cfg.k, work)
        walls = []
        for s in range(steps):
            shards = []
            for p in range(n_proc):
                fb, rb = prefix_pairs(f, r, (s * n_proc + p) * pairs_per_proc, pairs_per_proc, cfg.read_len)
                shards.append((fb.tobytes(), rb.tobytes()))
            walls.append(run_ref_shards(gfa_path, shards, cfg.k, work))
    finally:
        shutil.rmtree(work, ignore_errors=True)
    total = n_proc * pairs_per_proc
    wall = statistics.mean(walls)
    per_pair = max(wall - t_fixed, 0.25 * wall)            # (guards the difference against timer noise)
    return {"value": total / per_pair, "as_shipped_value": total / wall, "wall_s": wall, "fixed_cost_s": t_fixed,
            "pairs_per_step": total, "steps": steps}


def ref_sample_pairs(n_nodes: int, seconds: float = 10.0) -> int:
    """Pairs per process that keep the reference script busy for about `seconds` beyond its fixed cost
    (it does O(N) work per read: reference :19-21, :36-47)."""
    per_pair = 0.3e-3 + 0.42e-6 * n_nodes
    return int(min(40000, max(1000, seconds / per_pair)))


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


def py_port_rate(cfg, gfa: bytes, f, r, n_proc: int, pairs_per_proc: int, steps: int, warmup: int):
    """Python port of the reference path (oracle/pe_oracle.py) on n_proc host cores, index prebuilt."""
    import multiprocessing as mp

    def pieces(step):
        out = []
        for p in range(n_proc):
            fb, rb = prefix_pairs(f, r, (step * n_proc + p) * pairs_per_proc, pairs_per_proc, cfg.read_len)
            out.append((fb.tobytes(), rb.tobytes()))
        return out

    ctx = mp.get_context("fork")
    with ctx.Pool(n_proc, initializer=_worker_init, initargs=(gfa, cfg.k)) as pool:
        for s in range(warmup):
            pool.map(_worker_run, pieces(s))
        times = []
        for s in range(steps):
            pc = pieces(warmup + s)
            t0 = time.perf_counter()
            res = pool.map(_worker_run, pc)
            times.append(time.perf_counter() - t0)
            assert sum(x[0] for x in res) == n_proc * pairs_per_proc
    return n_proc * pairs_per_proc / statistics.mean(times), statistics.mean(times)


def c_port_rate(cfg, gfa: bytes, f, r, pairs: int):
    from oracle import c_oracle
    fb, rb = prefix_pairs(f, r, 0, pairs, cfg.read_len)
    t0 = time.perf_counter()
    c_oracle.run(gfa, fb, rb, cfg.k, 0)
    return pairs / (time.perf_counter() - t0)


def bind_near_gpu(local_rank: int):
    """Run this rank (and allocate its pinned buffers, first touch) on the cores of the GPU's NUMA node."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=sorted(synth.CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs of the resident block per GPU (default: per config)")
    ap.add_argument("--replay", type=int, default=0, help="block passes per step (default: config pairs / block pairs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="experiments: skip the end-to-end leg")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (experiments)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg0 = synth.CONFIGS[args.config]
    block = args.pairs or min(cfg0.pairs, BLOCK_PAIRS[args.config])
    per_gpu_pairs = cfg0.pairs if args.config != "C5" else cfg0.pairs // 8       # C5 is quoted on 8 GPUs
    replay = args.replay or (1 if args.pairs else max(1, -(-per_gpu_pairs // block)))
    pairs = block * replay                                                       # pairs per GPU per step
    workload = "%s: %d-strain %d bp graph (k=%d), %d x 2x%d bp read pairs per GPU per step (unique %d-pair block replayed x%d)" % (
        cfg0.name, cfg0.strains, cfg0.genome_len, cfg0.k, pairs, cfg0.read_len, block, replay)

    if args.impl == "reference":
        if rank != 0:
            return
        n_proc = min(REF_PROCS, os.cpu_count() or 1)
        have_script = os.path.exists(REF_SCRIPT)
        steps = max(1, min(args.steps, 3))                     # every step pays ~25 s of fixed cost per process
        cfg, g, genomes, ab = make_graph(args.config, pairs)
        ppp = ref_sample_pairs(len(g.ids)) if have_script else 4000
        f, r = make_reads(cfg, genomes, ab, n_proc * ppp * (steps + max(0, args.warmup)) + 64, 0)
        gfa = g.to_gfa()
        if have_script:
            res = ref_script_rate(cfg, gfa, f, r, n_proc, ppp, steps)
            rate, dt = res["value"], res["wall_s"]
            base = {"value": rate, "unit": UNIT, "cores": n_proc, "kind": "reference",
                    "sample": "UNMODIFIED utils/VStrains_PE_Inference.py (baseline/_ref), %d processes x %d pairs per step, %d steps; "
                              "value = pairs / (wall - fixed cost), fixed cost (index build + 2 N*N-line files) from a 0-read run of the "
                              "same %d processes" % (n_proc, ppp, steps, n_proc),
                    "as_shipped_value": res["as_shipped_value"], "wall_s_per_step": res["wall_s"], "fixed_cost_s": res["fixed_cost_s"]}
        else:
            rate, dt = py_port_rate(cfg, gfa, f, r, n_proc, ppp, steps, max(0, min(args.warmup, 1)))
            base = {"value": rate, "unit": UNIT, "cores": n_proc, "kind": "port",
                    "sample": "python port of the reference hash-table path (oracle/pe_oracle.py; baseline/_ref is absent), "
                              "%d processes x %d pairs per step, index prebuilt" % (n_proc, ppp)}
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
                "steps": steps, "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": workload, "graph_nodes": len(g.ids)},
                "cpu_baseline": base,
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
    numa = bind_near_gpu(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    n_thr = max(1, len(os.sched_getaffinity(0)) // max(1, min(world, 8)))
    os.environ["OMP_NUM_THREADS"] = str(n_thr)     # the read generator's OpenMP team: ranks share the host

    # ---- workload: graph (same on every rank) + this rank's reads, generated into pinned host memory
    cfg, g, genomes, ab = make_graph(args.config, pairs)
    cap = block * (2 * cfg.read_len + 18)
    h_f_full = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_r_full = torch.empty(cap, dtype=torch.uint8).pin_memory()
    f, r = make_reads(cfg, genomes, ab, block, rank, out=(h_f_full.numpy(), h_r_full.numpy()))
    h_f, h_r = h_f_full[: f.size], h_r_full[: r.size]
    gfa = g.to_gfa()
    n_nodes = len(g.ids)
    block_bytes = int(f.size + r.size)
    bytes_step = block_bytes * replay
    b_pair = block_bytes / block
    ix = pe_inference.PEIndex([bytes(s) for s in g.seqs], cfg.k, device=local_rank)
    for kv in args.opt:
        k_, v_ = kv.split("=")
        ix.set_option(k_, int(v_))

    class _Arr:                                    # torch view of device memory owned by the library
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}
    sparse = ix.is_sparse                          # graphs too large for N*N matrices (C5) keep sorted runs
    if sparse:
        mats = None
    else:
        mptr, mn = ix.matrices_device()
        mats = torch.as_tensor(_Arr(mptr, mn), device=dev) if mn else torch.zeros(0, dtype=torch.int64, device=dev)
    # the library launches on its own stream; collectives are enqueued on the same stream, so a step
    # needs no host synchronisation between the count and the allreduce
    lib_stream = torch.cuda.ExternalStream(ix.stream(), device=dev)

    def merge_sparse():
        """One exchange step over NVLink: the key space is cut into `world` equal ranges; every rank sends the
        runs of range r to rank r (all-to-all of the counts, then of keys and values) and merges what it
        received (sort + run-length reduce).  The merged run list stays distributed by key range; the work per
        rank does not grow with the number of ranks."""
        n_own, kptr, cptr = ix.sparse_device()
        keys_t = torch.as_tensor(_Arr(kptr, n_own), device=dev) if n_own else torch.zeros(0, dtype=torch.int64, device=dev)
        vals_t = torch.as_tensor(_Arr(cptr, n_own), device=dev) if n_own else torch.zeros(0, dtype=torch.int64, device=dev)
        cells = 2 * n_nodes * n_nodes
        bounds = torch.tensor([cells * (q + 1) // world for q in range(world)], dtype=torch.int64, device=dev)
        ends = torch.searchsorted(keys_t, bounds, right=False)             # the runs are sorted by key
        send = torch.diff(ends, prepend=torch.zeros(1, dtype=torch.int64, device=dev))
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send)
        s_list, r_list = [int(x) for x in send.tolist()], [int(x) for x in recv.tolist()]
        rk = torch.empty(sum(r_list), dtype=torch.int64, device=dev)
        rv = torch.empty(sum(r_list), dtype=torch.int64, device=dev)
        dist.all_to_all_single(rk, keys_t.clone(), r_list, s_list)
        dist.all_to_all_single(rv, vals_t.clone(), r_list, s_list)
        torch.cuda.current_stream().synchronize()
        ix.sparse_clear()
        ix.sparse_merge_device(rk.data_ptr(), rv.data_ptr(), int(rk.numel()))

    d_f = torch.empty(f.size, dtype=torch.uint8, device=dev)
    d_r = torch.empty(r.size, dtype=torch.uint8, device=dev)
    d_f.copy_(h_f, non_blocking=True)
    d_r.copy_(h_r, non_blocking=True)
    torch.cuda.synchronize()

    def reduce_step():
        if world > 1 and sparse:
            with torch.cuda.stream(lib_stream):
                merge_sparse()
        elif world > 1:
            with torch.cuda.stream(lib_stream):
                dist.all_reduce(mats)

    def step_device():
        ix.reset()
        for _ in range(replay):
            ix.count_device(d_f.data_ptr(), f.size, d_r.data_ptr(), r.size)
        reduce_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warmup):
        step_device()
    barrier()
    # keep the GPU under load for >= 1 s before timing so the clock samples mean something; the
    # extra step count is decided on rank 0 and broadcast (every rank runs the same collectives)
    t_w = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    extra = torch.tensor([max(0, min(2000, int(1.0 / max(time.perf_counter() - t_w, 1e-4))))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(extra, src=0)
    for _ in range(int(extra.item())):
        step_device()
    barrier()
    stage = {"ms_scan": 0.0, "ms_map": 0.0, "ms_count": 0.0, "ms_total": 0.0, "ms_k_scan_rows": 0.0, "ms_k_walk": 0.0}
    n_scan_launches = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_lo = time.perf_counter()
    with torch.cuda.stream(lib_stream):
        e0.record()
    launches = 0
    for _ in range(args.steps):
        ix.reset()
        for _ in range(replay):
            ix.count_device(d_f.data_ptr(), f.size, d_r.data_ptr(), r.size)
        st = ix.stats()                            # stage times and launch counts accumulate since the reset
        for k in stage:
            stage[k] += st[k]
        n_scan_launches += st["n_k_scan_rows"]
        launches += st["kernel_launches"]
        reduce_step()
    with torch.cuda.stream(lib_stream):
        e1.record()
    barrier()
    t_hi = time.perf_counter()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_lo, t_hi)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * pairs * args.steps / (ms_max * 1e-3)
    st = ix.stats()

    # ---- the merged result against the pair counters (every rank, after the timed region) --------
    keys_total = torch.tensor([st["n_keys"], st["used_pairs"], st["total_pairs"], st["n_pairs"], st["short_pairs"]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(keys_total)
    kt = [int(x) for x in keys_total.tolist()]
    if not sparse:
        torch.cuda.synchronize()
        msum = int(mats.sum().item())
        check = {"matrix_sum": msum, "n_keys_all_ranks": kt[0], "equal": msum == kt[0]}
    else:
        keys_, counts_ = ix.sparse()                    # (N > 1: this rank's key range of the merged runs)
        rs = torch.tensor([int(counts_.sum()), int(keys_.size)], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(rs)
        check = {"run_sum": int(rs[0].item()), "n_keys_all_ranks": kt[0], "equal": int(rs[0].item()) == kt[0], "runs": int(rs[1].item())}
    assert check["equal"], "merged counts != link keys counted on all ranks: %r" % (check,)
    assert kt[2] == kt[1] + kt[3] + kt[4], "total != used + N + short"

    # ---- end to end: pinned host buffers -> matrices on the host (pinned too: the D2H read of the result then runs at
    # PCIe speed; into fresh pageable arrays it cost 45 ms of a 615 ms step) ------------------------
    h_out = None
    if not sparse:
        h_pin = torch.empty((2, n_nodes, n_nodes), dtype=torch.int64).pin_memory()
        h_out = (h_pin[0].numpy().view(np.uint64), h_pin[1].numpy().view(np.uint64))

    def step_e2e():
        ix.reset()
        for _ in range(replay):
            ix.count_host_ptr(h_f.data_ptr(), f.size, h_r.data_ptr(), r.size)
        reduce_step()
        if world > 1:
            lib_stream.synchronize()
        return ix.sparse() if sparse else ix.matrices(out=h_out)

    # the end-to-end ceiling: a plain pinned host->device copy of the same bytes (PCIe), all ranks at once
    h2d_gbs = None
    try:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(3):
            barrier()
            ev0.record()
            d_f.copy_(h_f, non_blocking=True)
            d_r.copy_(h_r, non_blocking=True)
            ev1.record()
            torch.cuda.synchronize()
            best = min(best, ev0.elapsed_time(ev1))
        tb = torch.tensor([best], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        h2d_gbs = world * block_bytes / (float(tb.item()) * 1e-3) / 1e9
    except Exception:
        pass

    out = step_e2e()
    barrier()
    e2e_steps = 1 if args.no_e2e else max(2, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(0 if args.no_e2e else e2e_steps):
        out = step_e2e()
    barrier()
    dt = max(time.perf_counter() - t0, 1e-9)
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * pairs * e2e_steps / float(t.item())
    d2h_bytes = int(16 * out[0].size) if sparse else int(2 * n_nodes * n_nodes * 8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # dominant kernel: k_scan_rows streams every algorithmic byte (one launch per mate file and replay): record
    # split + 2-bit pack in one pass.  Its duration comes from CUDA events recorded around each launch on the
    # library's own stream, over the timed region; k_walk (the second largest) is reported the same way.
    k_ms = stage["ms_k_scan_rows"] / max(1, n_scan_launches)            # average launch duration
    w_ms = stage["ms_k_walk"] / max(1, n_scan_launches)
    k_bytes = block_bytes / 2.0                                         # algorithmic bytes per launch (one mate of the block)
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "k_scan_rows_traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
        traffic = k_bytes * tj["dram_bytes_per_algorithmic_byte"]          # from the committed ncu --set full capture
    whole = value / world * b_pair / 1e9 / peak
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload, "graph_nodes": n_nodes, "bytes_per_pair": b_pair, "pairs_per_gpu_per_step": pairs,
                   "block_pairs": block, "replay_factor": replay,
                   "l2": "inputs (%.0f MB resident per GPU) larger than the 126 MB L2; no flush needed" % (block_bytes / 1e6),
                   "index_build_ms": st["ms_index"], "keys_per_pair": st["n_keys"] / max(1, st["used_pairs"]),
                   "reads_fast": st["reads_fast"], "reads_generic": st["reads_generic"], "numa_node": numa,
                   "read_memo_hit_rate": st["reads_memo"] / max(1, 2 * pairs), "scan_redo_tiles": st["scan_redo_tiles"],
                   "read_memo": "cleared with the matrices at every step start (the first block of a step runs cold); it only ever holds "
                                "error-free reads, which are bounded by the graph (2 x genome x strains) and which a real file of this "
                                "size repeats exactly as the replayed block does; option memo=0 walks every read (profiles/r02_experiments.txt)",
                   "counting": "sparse runs (LSD radix sort + RLE)" if sparse else "dense matrices (pair aggregation + radix partition + counting sort)"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "merged_result_check": check,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_step, "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_steps, "pinned_h2d_copy_gbs_all_gpus": h2d_gbs,
                "frac_of_h2d_copy": (e2e_value * b_pair / 1e9 / h2d_gbs) if h2d_gbs else None},
        "stages_ms_per_step": {k: v / args.steps for k, v in stage.items()},
        "whole_job_hbm_frac": whole,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "whole_path_frac": whole,
                     "kernel": "k_scan_rows (K1+K2 in one pass, tiles independent: TMA tile -> terminators -> guessed line phase -> 2-bit rows in tile-local slots)",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": k_bytes, "launch_ms": k_ms,
                     "launches_per_step": n_scan_launches / args.steps,
                     "kernel_share_of_step": stage["ms_k_scan_rows"] / max(1e-9, stage["ms_total"]),
                     "second_kernel": {"kernel": "k_memo + k_walk (K4 first tier: read memo lookup, then seed + flat walk + list interning for the reads it cannot answer)",
                                       "launch_ms": w_ms, "share_of_step": stage["ms_k_walk"] / max(1e-9, stage["ms_total"]),
                                       "reads_per_s": (block / (w_ms * 1e-3)) if w_ms > 0 else None}},
    }
    if world == 1 and not args.no_cpu_baseline:
        port_n, c_n = 20000, 100_000
        fs, rs = make_reads(cfg, genomes, ab, c_n + 64, 0) if block < c_n + 64 else (f, r)
        port_rate, _ = py_port_rate(cfg, gfa, fs, rs, 1, port_n, 1, 0)
        extra = {"python_port_1_core": {"value": port_rate, "sample": "oracle/pe_oracle.py, first %d pairs, index prebuilt" % port_n},
                 "c_port_all_cores": {"value": c_port_rate(cfg, gfa, fs, rs, c_n), "cores": os.cpu_count(),
                                      "sample": "oracle/pe_oracle.c with OpenMP on the first %d pairs, index build included" % c_n}}
        if os.path.exists(REF_SCRIPT):
            n = ref_sample_pairs(n_nodes, 6.0)
            res = ref_script_rate(cfg, gfa, fs, rs, 1, n, 1)
            line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": "UNMODIFIED utils/VStrains_PE_Inference.py (baseline/_ref) on the first %d pairs of the same workload, 1 process; "
                                              "value = pairs / (wall - fixed cost), fixed cost (index build + 2 N*N-line files) from a 0-read run" % n,
                                    "as_shipped_value": res["as_shipped_value"], "wall_s": res["wall_s"], "fixed_cost_s": res["fixed_cost_s"], **extra}
        else:
            line["cpu_baseline"] = {"value": port_rate, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "python port of the reference path (oracle/pe_oracle.py; baseline/_ref absent), first %d pairs, 1 process, index prebuilt" % port_n,
                                    **extra}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
