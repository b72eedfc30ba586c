#!/usr/bin/env python
"""bench.py — rate-limit decisions/s of the B200 evaluation path on BASELINE.json's headline workload.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path (default)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU worker-pool path (oracle port) on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...        # N GPUs: key space sharded by the replicated-hash ring

A "step" is one 65 536-request batch per GPU through the whole hot path (group -> probe -> bucket update -> response).
Workload at N = 1: BASELINE config 3 — 100 M resident keys, Zipf s = 1.1, TOKEN/LEAKY 50/50 by key.  At N > 1:
config 4 — the same key space sharded over the N GPUs by the 512-replica FNV-1 ring, requests routed to their owner
with NCCL all-to-all and responses routed back (weak scaling: 65 536 requests ingested per GPU per step).

Prints ONE JSON line (rank 0).  `value` = decisions/s with the request batches already resident in HBM (a pool of
pre-generated batches larger than L2 is cycled; the 12.8 GB table is far larger than L2).  `e2e` = the same through
the public host API (gub_submit_async with pinned host buffers: H2D of requests + D2H of responses inside the timed
region).  `roofline` = algorithmic bytes of the dominant kernel / its CUDA-event time, vs MEASURED_PEAKS.json.
`cpu_baseline` = the oracle's worker-pool port timed on this host's cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BATCH = 65536
T0 = 1_700_000_000_000
ALGO_BYTES_PER_DECISION = 224  # 64 slot read + 64 slot write-back + 64 request + 32 response


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--keys", type=int, default=int(os.environ.get("GUB_BENCH_KEYS", 100_000_000)))
    ap.add_argument("--zipf", type=float, default=1.1)
    ap.add_argument("--pool", type=int, default=32, help="distinct pre-generated batches cycled through (32 x 6 MiB > L2)")
    ap.add_argument("--cpu-keys", type=int, default=int(os.environ.get("GUB_BENCH_CPU_KEYS", 10_000_000)))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--route", default=os.environ.get("GUB_ROUTE", "p2p"), choices=["p2p", "nccl"],
                    help="N > 1: how request records reach their owning GPU (NVLink mailboxes written by the routing kernels, or NCCL all-to-all)")
    ap.add_argument("--no-route-overlap", action="store_true", help="p2p route: run routing and evaluation on one stream (no overlap of step e+1's routing with step e's evaluation)")
    ap.add_argument("--ncu-window", type=int, default=0,
                    help="profile this many extra steps between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)")
    return ap.parse_args()


# ---- helpers ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device_index):
        self.rows, self.proc, self.dev = [], None, device_index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if not self.rows:  # the timed region was shorter than one sampling period: take one reading now
            try:
                q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                self.rows = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                           capture_output=True, text=True, timeout=10).stdout.strip().splitlines()
            except Exception:
                pass
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def gen_batch(rng, n, n_keys, created_at, zipf_s, dtype):
    from workloads import bench_requests, zipf_ids
    ids = zipf_ids(rng, n, n_keys, zipf_s)
    return bench_requests(ids, created_at, mixed=True, dtype=dtype), ids


def batch_stats(ids):
    _, counts = np.unique(ids, return_counts=True)
    light = (counts > 1) & (counts <= 16)  # INLINE in gub_kernels.cuh
    heavy = counts > 16
    return dict(distinct=int(len(counts)), singles=int((counts == 1).sum()), light_groups=int(light.sum()),
                light_requests=int(counts[light].sum()), heavy_groups=int(heavy.sum()), heavy_requests=int(counts[heavy].sum()),
                top=int(counts.max()))


# ---- reference arm: the reference's CPU path (oracle port; the Go reference cannot be built in this image) ----------
def cpu_leg(n_keys, zipf_s, seconds, seed, steps=None, warmup=0, step_size=BATCH):
    import oracle_py as O
    from workloads import bench_requests
    cores = os.cpu_count() or 1
    workers = cores
    pool = O.Pool(workers=workers, cache_size=max(4 * n_keys, 1 << 20), now_ms=T0)
    rng = np.random.default_rng(seed)
    # warm pass: make every key resident, through the same worker-pool path
    t_fill = time.perf_counter()
    chunk = 1 << 20
    for lo in range(0, n_keys, chunk):
        ids = np.arange(lo, min(n_keys, lo + chunk), dtype=np.int64)
        pool.submit_hashed(bench_requests(ids, T0), threads=cores)
    t_fill = time.perf_counter() - t_fill
    batches = [gen_batch(rng, step_size, n_keys, T0 + 1 + b, zipf_s, O.HREQ_DTYPE)[0] for b in range(16)]
    for w in range(max(warmup, 1)):
        pool.set_now(T0 + 1 + w)
        pool.submit_hashed(batches[w % len(batches)], threads=cores)
    done, t_used, b = 0, 0.0, 0
    while (steps is None and t_used < seconds) or (steps is not None and b < steps):
        pool.set_now(T0 + 1 + b)
        pool.submit_hashed(batches[b % len(batches)], threads=cores)
        t_used += pool.last_mt_seconds
        done += step_size
        b += 1
    return dict(value=done / t_used, seconds=t_used, steps=b, cores=cores, fill_seconds=t_fill, keys=n_keys, step_size=step_size)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a step is one 65 536-request batch of the workload (~2.5 ms on 128 threads); only beyond 20 000 steps is the batch shrunk so
    # that the whole run stays within about a minute
    step_size = BATCH if args.steps <= 20000 else max(2048, int(BATCH * 20000 / args.steps) // 256 * 256)
    r = cpu_leg(args.cpu_keys, args.zipf, args.cpu_seconds, 0xB200 + 3, steps=args.steps, warmup=min(args.warmup, 20), step_size=step_size)
    sample = (f"{r['steps']} x {step_size}-request Zipf({args.zipf}) batches over {r['keys']:,} resident keys (scaled down from "
              f"{args.keys:,} to bound the warm pass), TOKEN/LEAKY 50/50, {r['cores']} worker threads")
    line = {
        "impl": "reference", "metric": "rate-limit decisions/sec", "value": r["value"], "unit": "decisions/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(r["steps"], 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config 3 shape on CPU: {sample}", "batch": step_size, "keys": r["keys"], "zipf_s": args.zipf},
        "cpu_baseline": {"value": r["value"], "unit": "decisions/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- this repo's arm --------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import gubernator_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N = world
    n_keys = args.keys
    seed = 0xB200 + (3 if N == 1 else 4)
    rng = np.random.default_rng(seed + 1000 * rank)

    # table sized at load factor <= 0.5 for this shard's share of the key space (+25 % for ring imbalance)
    shard_keys = n_keys if N == 1 else int(n_keys / N * 1.25)
    capacity = max(2 * shard_keys, 1 << 16)
    max_batch = BATCH if N == 1 else 262144
    tab = g.Table(capacity, max_batch=max_batch, device=local)
    stream = torch.cuda.current_stream().cuda_stream

    ring = None
    if N > 1:
        from gubernator_b200.sharded import shard_addresses
        ring = g.Ring(0, 512)
        for a in shard_addresses(N):
            ring.add(a)

    def t2np(t, dtype):
        return t.cpu().numpy().reshape(-1).view(dtype)

    # ---- one step of the sharded path (N > 1): route -> all-to-all -> evaluate -> all-to-all back -> unroute
    sharded = None
    if N > 1:
        from gubernator_b200.sharded import GpuBackend, P2PStep, ShardedStep
        if args.route == "p2p":
            sharded = P2PStep(tab, ring, N, rank, cap=BATCH)  # mailbox capacity = the largest batch a shard ingests per step
            sharded.connect(dist)
        else:
            sharded = ShardedStep(GpuBackend(tab, ring, N, dev, 262144), dist, N)

    # ---- warm pass: make every key resident through the real path
    t_fill = time.perf_counter()
    clk0 = g.clock_fill(T0)
    from workloads import bench_requests
    chunk = BATCH if N > 1 else 1 << 20
    d_chunk = torch.empty((chunk, 64), dtype=torch.uint8, device=dev)
    d_chunk_out = torch.empty((chunk, 32), dtype=torch.uint8, device=dev)
    my_lo = (n_keys * rank) // N
    my_hi = (n_keys * (rank + 1)) // N
    n_fill_steps = (n_keys // N + chunk - 1) // chunk  # identical on every rank (collectives inside)
    for s in range(n_fill_steps):
        lo = my_lo + s * chunk
        hi = min(my_hi, lo + chunk)
        ids = np.arange(lo, max(hi, lo), dtype=np.int64)
        reqs = bench_requests(ids, T0, dtype=g.REQ_DTYPE)
        n = len(reqs)
        if n:
            d_chunk[:n].copy_(torch.from_numpy(reqs.view(np.uint8).reshape(n, 64)), non_blocking=False)
        if N == 1:
            tab.submit_device(d_chunk.data_ptr(), n, clk0, d_chunk_out.data_ptr(), stream)
        else:
            sharded.step(d_chunk, n, clk0, d_chunk_out)
    torch.cuda.synchronize()
    t_fill = time.perf_counter() - t_fill
    c0 = tab.counters()

    # ---- pre-generated batch pool, resident in HBM
    pool_n = max(2, args.pool)
    host_batches, stats = [], []
    for b in range(pool_n):
        reqs, ids = gen_batch(rng, BATCH, n_keys, T0 + 1 + b, args.zipf, g.REQ_DTYPE)
        host_batches.append(reqs)
        stats.append(batch_stats(ids))
    d_batches = [torch.from_numpy(r.view(np.uint8).reshape(BATCH, 64)).to(dev) for r in host_batches]
    d_outs = [torch.empty((BATCH, 32), dtype=torch.uint8, device=dev) for _ in range(pool_n)]
    clocks = [g.clock_fill(T0 + 1 + b) for b in range(args.steps + args.warmup + 8)]

    # p2p route: the routing kernels run on their own (ingest) stream, so step e+1 is routed while step e is evaluated
    step_kw = {}
    ingest = None
    if N > 1 and args.route == "p2p" and not args.no_route_overlap:
        ingest = torch.cuda.Stream(device=dev)
        step_kw = {"ingest_stream": ingest.cuda_stream}

    def one_step(b):
        k = b % pool_n
        if N == 1:
            tab.submit_device(d_batches[k].data_ptr(), BATCH, clocks[min(b, len(clocks) - 1)], d_outs[k].data_ptr(), stream)
        else:
            sharded.step(d_batches[k], BATCH, clocks[min(b, len(clocks) - 1)], d_outs[k], **step_kw)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for b in range(args.warmup):
        one_step(b)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for b in range(args.steps):
        one_step(args.warmup + b)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks_info = sampler.stop() if rank == 0 else None
    if dist is not None:
        tms = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = N * BATCH * args.steps / (ms * 1e-3)

    # ---- informational: the same path with 4 x larger launches (262 144 requests; one C-ABI call, chunked by max_batch = 65 536
    # at N = 1): shows how much of the per-batch time is fixed cost rather than per-request work.  Not the headline.
    big = None
    if N == 1 and pool_n >= 4:
        d_big = torch.cat(d_batches[:4], dim=0).contiguous()
        d_big_out = torch.empty((4 * BATCH, 32), dtype=torch.uint8, device=dev)
        big_steps = max(10, args.steps // 8)
        for b in range(3):
            tab.submit_device(d_big.data_ptr(), 4 * BATCH, clocks[0], d_big_out.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in range(big_steps):
            tab.submit_device(d_big.data_ptr(), 4 * BATCH, clocks[min(b, len(clocks) - 1)], d_big_out.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        big = {"requests_per_call": 4 * BATCH, "value": 4 * BATCH * big_steps / (e0.elapsed_time(e1) * 1e-3), "unit": "decisions/s",
               "note": "same key pool every call: hot keys saturate; informational only"}

    # ---- per-kernel timing leg (separate from the number above: events between kernels perturb the pipeline)
    tab.set_profiling(True)
    prof_steps = min(args.steps, 100)
    for b in range(prof_steps):
        one_step(args.warmup + args.steps + b)
    barrier()
    prof = tab.get_profile()
    tab.set_profiling(False)

    if args.ncu_window > 0:  # the only launches an `ncu --profile-from-start off` run sees
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for b in range(args.ncu_window):
            one_step(args.warmup + args.steps + prof_steps + b)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # ---- phase trace of the batch kernel (diagnostic): per-CTA %globaltimer stamps of one launch
    phase_trace = None
    try:
        tab.set_trace(True)
        for b in range(3):
            one_step(args.warmup + args.steps + prof_steps + 64 + b)
        torch.cuda.synchronize()
        phase_trace = tab.get_trace()
        raw = tab.get_trace_raw()
        live = raw[:, 0] > 0
        t0 = raw[live, 0].min()
        rel = (raw[live].astype(np.int64) - np.int64(t0)) / 1e3
        slow = int(np.argmax(rel[:, 10]))
        phase_trace["slowest_cta"] = {"cta": int(np.nonzero(live)[0][slow]), "us": [round(float(v), 2) for v in rel[slow]]}
        phase_trace["median_cta_us"] = [round(float(v), 2) for v in np.median(rel, axis=0)]
        tab.set_trace(False)
    except Exception as ex:
        phase_trace = {"error": str(ex)}

    # ---- end-to-end leg through the host API with pinned buffers (N == 1 path; at N > 1 each rank ingests from its host)
    e2e = None
    if not args.no_e2e:
        depth = 4
        pin = [(g.native.PinnedArray(BATCH, g.REQ_DTYPE), g.native.PinnedArray(BATCH, g.RESP_DTYPE)) for _ in range(depth)]
        for k in range(depth):
            pin[k][0].array[:] = host_batches[k % pool_n]
        if N == 1:
            def e2e_run(submit):
                tickets = [None] * depth
                for b in range(min(args.warmup, 8)):
                    tab.wait(submit(b % depth, b))
                torch.cuda.synchronize()
                t_sub = t_wait = 0.0
                t0 = time.perf_counter()
                for b in range(e2e_steps):
                    k = b % depth
                    a = time.perf_counter()
                    if tickets[k] is not None:
                        tab.wait(tickets[k])  # the response buffer of this slot has been read back
                    c = time.perf_counter()
                    tickets[k] = submit(k, b)
                    t_wait += c - a
                    t_sub += time.perf_counter() - c
                for k in range(depth):
                    if tickets[k] is not None:
                        tab.wait(tickets[k])
                torch.cuda.synchronize()
                dt_ = time.perf_counter() - t0
                e2e_split.append({"cpu_us_in_submit": 1e6 * t_sub / e2e_steps, "cpu_us_in_wait": 1e6 * t_wait / e2e_steps,
                                  "us_per_step": 1e6 * dt_ / e2e_steps})
                return dt_
            e2e_steps = args.steps
            e2e_split = []
            # (a) compact records: 32 B per request + one small parameter table per batch (gub_submit_compact_async)
            cpin, ppin, bases, nparams = [], [], [], []
            for k in range(depth):
                c, prm, base = g.native.compact_batch(host_batches[k % pool_n])
                ca, pa = g.native.PinnedArray(BATCH, g.native.CREQ_DTYPE), g.native.PinnedArray(max(len(prm), 1), g.native.PARAMS_DTYPE)
                ca.array[:] = c; pa.array[:len(prm)] = prm
                cpin.append(ca); ppin.append(pa); bases.append(base); nparams.append(len(prm))
            dt = e2e_run(lambda k, b: tab.submit_compact_async(cpin[k].ptr, BATCH, ppin[k].ptr, nparams[k], bases[k],
                                                               clocks[min(b, len(clocks) - 1)], pin[k][1].ptr))
            # (b) full 64 B records (gub_submit_async), for comparison
            dt_full = e2e_run(lambda k, b: tab.submit_async(pin[k][0].ptr, BATCH, clocks[min(b, len(clocks) - 1)], pin[k][1].ptr))
            e2e_full = {"value": BATCH * e2e_steps / dt_full, "unit": "decisions/s", "h2d_bytes_per_step": BATCH * 64, "d2h_bytes_per_step": BATCH * 32,
                        "api": "gub_submit_async (64-byte records)"}
            h2d_compact = BATCH * 32 + int(np.mean(nparams)) * 32
            e2e_full["host_time"] = e2e_split[1]

            for a in cpin + ppin:
                a.free()
        else:
            # sharded e2e: per step copy the batch from pinned host memory, run the routed step, read responses back
            h_in = [torch.from_numpy(host_batches[k % pool_n].view(np.uint8).reshape(BATCH, 64)).pin_memory() for k in range(depth)]
            h_out = [torch.empty((BATCH, 32), dtype=torch.uint8).pin_memory() for _ in range(depth)]
            d_in = torch.empty((BATCH, 64), dtype=torch.uint8, device=dev)
            d_o = torch.empty((BATCH, 32), dtype=torch.uint8, device=dev)
            e2e_steps = args.steps
            barrier()
            t0 = time.perf_counter()
            d_ins = [torch.empty((BATCH, 64), dtype=torch.uint8, device=dev) for _ in range(2)]
            for b in range(e2e_steps):
                k = b % depth
                if ingest is not None:  # ingest copy + routing on the ingest stream, evaluation + read-back on the current one
                    with torch.cuda.stream(ingest):
                        d_ins[b & 1].copy_(h_in[k], non_blocking=True)
                    sharded.step(d_ins[b & 1], BATCH, clocks[min(b, len(clocks) - 1)], d_o, **step_kw)
                else:
                    d_in.copy_(h_in[k], non_blocking=True)
                    sharded.step(d_in, BATCH, clocks[min(b, len(clocks) - 1)], d_o)
                h_out[k].copy_(d_o, non_blocking=True)
            barrier()
            dt = time.perf_counter() - t0
            if dist is not None:
                tdt = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
                dt = float(tdt.item())
        e2e = {"value": N * BATCH * e2e_steps / dt, "unit": "decisions/s", "h2d_bytes_per_step": h2d_compact if N == 1 else N * BATCH * 64,
               "d2h_bytes_per_step": N * BATCH * 32,
               "api": "gub_submit_compact_async (pinned host buffers, 32-byte records + parameter table, depth 4)" if N == 1 else
               "pinned H2D + route/exchange/evaluate/return/unroute + D2H per step"}
        if N == 1:
            e2e["host_time"] = e2e_split[0]
            e2e["full_records"] = e2e_full
        for a, b_ in pin:
            a.free(); b_.free()

    c1 = tab.counters()
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peak, peak_src = load_peaks()
    kms = {"k_group": prof["k_group_ms"], "k_rank": prof["k_rank_ms"], "k_eval": prof["k_eval_ms"], "k_finish": prof["k_finish_ms"]}
    launches = max(prof["launches"], 1)
    dom = max(kms, key=kms.get)
    st = {k: float(np.mean([s[k] for s in stats])) for k in stats[0]}
    units = BATCH if N == 1 else BATCH  # per-GPU requests per launch (N > 1: expected share after routing)
    multi_req = st["light_requests"] + st["heavy_requests"]
    multi_grp = st["light_groups"] + st["heavy_groups"]
    alg = {  # algorithmic bytes per launch, by kernel (Kernels)
        "k_group": 16.0 * units,                                                        # 8 B key in, ent + meta out
        # this build evaluates keys seen once in k_rank (EARLY_SINGLES): the whole 224 B per singleton; every request reads its
        # ent/meta/entry (12 B); members of repeated keys read their request + the representative's (128 B) and write a rank;
        # per repeated key one slot probe + one 96 B snapshot
        "k_rank": ALGO_BYTES_PER_DECISION * st["singles"] + 12.0 * units + 132.0 * multi_req + 160.0 * multi_grp,
        "k_eval": (64 + 96 + 32 + 12.0) * multi_req + 64.0 * multi_grp,  # request + snapshot + response per member; one slot write-back per key
        "k_finish": 0.0,                                                   # only non-uniform groups (none in this workload)
    }
    dom_ms = kms[dom] / launches
    path_ms = sum(kms.values()) / launches
    achieved = alg[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch of that kernel, from the committed `ncu --set full` capture
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get("bytes_per_launch", {}).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": {k: v / launches for k, v in kms.items()},
                "algorithmic_bytes_per_launch": alg,
                "path": {"achieved": ALGO_BYTES_PER_DECISION * units / (path_ms * 1e-3) / 1e9 if path_ms > 0 else 0.0,
                         "bytes_per_decision": ALGO_BYTES_PER_DECISION, "ms_per_batch_kernels_only": path_ms}}
    roofline["path"]["frac"] = roofline["path"]["achieved"] / peak
    # secondary ceiling (north_star's "HBM-random-access roofline"): the measured random 64-byte read-modify-write rate of this
    # device over this very table, with the loads / stores the batch kernels use; the path's slot traffic (64 B read + 64 B
    # write-back per decision, algorithmic) as a fraction of it.  Runs last: nothing measured above can be disturbed by it.
    try:
        ra_gbs = tab.probe_random_access(1 << 26)
        slot_gbs = (value / N) * 128.0 / 1e9
        roofline["random_access"] = {"measured_gbs": ra_gbs, "accesses": 1 << 26, "bytes_per_access": 128,
                                     "path_slot_gbs": slot_gbs, "frac": slot_gbs / ra_gbs if ra_gbs > 0 else None}
    except Exception as ex:  # diagnostic only
        roofline["random_access"] = {"error": str(ex)}

    cpu = None
    if not args.no_cpu_baseline and N == 1:
        r = cpu_leg(args.cpu_keys, args.zipf, args.cpu_seconds, seed)
        cpu = {"value": r["value"], "unit": "decisions/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} x {BATCH}-request Zipf({args.zipf}) batches over {r['keys']:,} resident keys (scaled down from "
                         f"{n_keys:,} to bound the warm pass), oracle worker-pool port on {r['cores']} threads, {r['seconds']:.1f} s timed"}

    per_step_launches = 4 if N == 1 else (4 + 6 if args.route == "p2p" else 4 + 3 + 1)  # group/rank/eval/finish (+ routing kernels)
    line = {
        "metric": "rate-limit decisions/sec", "value": value, "unit": "decisions/s", "n_gpus": N, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64+f64", "data": "synthetic",
        "config": {"workload": ("BASELINE config 3: 100M keys, Zipf s=1.1, TOKEN/LEAKY 50/50, 1xB200" if N == 1 else
                                f"BASELINE config 4: 100M keys sharded over {N}xB200 by replicated_hash (fnv1, 512 replicas), Zipf s=1.1, routing: " +
                                ("NVLink peer-memory mailboxes written by the routing kernels" + ("" if args.no_route_overlap else ", routing of step e+1 overlapped with evaluation of step e (two streams)") if args.route == "p2p" else "NCCL all-to-all")),
                   "keys": n_keys, "batch_per_gpu": BATCH, "zipf_s": args.zipf, "table_slots_per_gpu": capacity,
                   "cache": f"inputs cycle through {pool_n} resident batches ({pool_n * 6} MiB > L2); table {capacity * 64 / 1e9:.1f} GB >> L2",
                   "batch_profile": st, "fill_seconds": t_fill, "resident_keys_after_fill": c0["inserts"]},
        "clocks": clocks_info, "e2e": e2e, "gpu_launches": per_step_launches * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "larger_calls": big, "phase_trace": phase_trace,
        "counters": {k: c1[k] - c0[k] for k in c1},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
