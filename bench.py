#!/usr/bin/env python
"""bench.py — haystacks/sec of the match_list hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (restated)

A "step" is one Matcher::match_list over one synthetic haystack list of BASELINE.json configs[2]
shape (needle 'deadbeef', 10M haystacks, len <= 64, mean 48, max_typos = 1) per GPU (weak scaling:
each rank holds its own 10M-item shard; for N > 1 a step ends with the NCCL all-gather of the
per-shard runs and the merge — Matcher::match_list_parallel).
  value : whole-job haystacks/s with the packed corpus already resident in HBM when the timed
          region starts and the ordered match list left in HBM (device in, device out; the same
          definition at every N).  `value_host_out` (N = 1) additionally copies the list to pinned
          host memory inside the timed region.
  e2e   : the same metric through frz_match_list_host — host Arrow buffers in pinned memory in,
          host matches out; pack + H2D + kernels + D2H all inside the timed region.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = dict(needle="deadbeef", n=10_000_000, mu=48, max_len=64, max_typos=1, seed=12345)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=WORKLOAD["n"], help="haystacks per GPU")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 5)")
    ap.add_argument("--e2e-offsets", type=int, default=32, choices=[32, 64],
                    help="Arrow offset width of the e2e input (32 = Utf8, 64 = LargeUtf8)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="haystacks in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, help="emulate_lanes (0 = what the reference picks on this CPU)")
    # the other BASELINE.json configs are parity-test cases; these flags let profiles/ record their numbers too
    ap.add_argument("--needle", default=WORKLOAD["needle"])
    ap.add_argument("--max-typos", type=int, default=WORKLOAD["max_typos"])
    ap.add_argument("--mu", type=int, default=WORKLOAD["mu"])
    ap.add_argument("--max-len", type=int, default=WORKLOAD["max_len"])
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_config(args):
    from frizbee_b200.types import Config
    return Config(max_typos=WORKLOAD["max_typos"], emulate_lanes=args.lanes)


def config_block(args, world, extra=None):
    default = (WORKLOAD["needle"], WORKLOAD["max_typos"], WORKLOAD["mu"], WORKLOAD["max_len"]) == ("deadbeef", 1, 48, 64)
    c = {"workload": f"needle '{WORKLOAD['needle']}' (len {len(WORKLOAD['needle'])}) vs {args.n} synthetic ASCII haystacks per GPU, "
                     f"len<={WORKLOAD['max_len']} (mean {WORKLOAD['mu']}, sd {WORKLOAD['mu'] // 4}), max_typos={WORKLOAD['max_typos']}, "
                     f"5% full / 20% partial matches, seed 12345"
                     + (" (BASELINE.json configs[2], the configuration the 10x target is quoted on)" if default else ""),
         "haystacks_per_gpu": args.n, "n_gpus": world, "sort": "ScoreThenIndexAsc",
         "l2": "inputs (~560 MB packed per GPU) are larger than the 126 MB L2; no explicit flush"}
    if extra:
        c.update(extra)
    return c


def run_reference(args):
    """The reference's own CPU implementation of the path (restated; no Rust toolchain here), all host threads,
    on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from frizbee_b200 import synth
    from oracle import cpu_baseline as cb
    cfg = workload_config(args)
    threads = cb.host_threads()
    sample = args.cpu_sample or args.n   # the whole single-GPU workload: large enough to amortise thread start-up
    data, off = synth.generate(WORKLOAD["needle"], sample, WORKLOAD["mu"], WORKLOAD["max_len"], WORKLOAD["seed"])
    lanes = args.lanes or detect_lanes(cfg)
    cfg = cfg.with_(emulate_lanes=lanes)
    # be generous to the CPU: the reference's final k-way merge is single-threaded, so more threads is not
    # always faster — use the thread count (all / half / quarter of the host threads) that runs fastest
    best = None
    for cand_threads in sorted({threads, max(1, threads // 2), max(1, threads // 4)}, reverse=True):
        dt, _ = cb.timed([WORKLOAD["needle"]], cfg, data, off, cand_threads, repeats=max(1, args.warmup))
        if best is None or dt < best[0]:
            best = (dt, cand_threads)
    threads = best[1]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = cb.match_list_parallel([WORKLOAD["needle"]], cfg, data, off, threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {"impl": "reference", "metric": "haystacks/sec", "value": value, "unit": "haystacks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_block(args, args.gpus, {"sample_haystacks_per_step": sample, "matches_per_step": int(len(res))}),
            "cpu_baseline": {"value": value, "unit": "haystacks/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} haystacks of the same workload per step; {cb.describe()}, "
                                       f"threaded like match_list_parallel (2048-item work claiming), emulating the "
                                       f"{lanes}-lane reference backend"},
            "e2e": {"value": value, "unit": "haystacks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def detect_lanes(cfg) -> int:
    import frizbee_b200 as F
    m = F.Matcher(WORKLOAD["needle"], cfg)
    lanes = m.backend_info()["prefilter_lanes"]
    m.close()
    return lanes


def run_ours(args):
    import torch
    import torch.distributed as dist
    import frizbee_b200 as F
    from frizbee_b200 import parallel, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload_config(args)
    n = args.n
    # each rank holds its own shard (weak scaling); shard r covers indices [r*n, (r+1)*n)
    data_np, off_np = synth.generate(WORKLOAD["needle"], n, WORKLOAD["mu"], WORKLOAD["max_len"], WORKLOAD["seed"] + rank)
    # pinned host buffers (inputs of the e2e call, and the output of every step)
    data_pin = torch.empty(data_np.size, dtype=torch.uint8, pin_memory=True)
    off_pin = torch.empty(off_np.size, dtype=torch.int64, pin_memory=True)
    out_pin = torch.empty(max(n * world, 1), dtype=torch.int64, pin_memory=True)
    data_h = data_pin.numpy(); data_h[:] = data_np
    off_h = off_pin.numpy().view(np.uint64); off_h[:] = off_np
    if args.e2e_offsets == 32 and int(off_np[-1]) < 2 ** 31:   # Arrow Utf8: int32 offsets
        off32_pin = torch.empty(off_np.size, dtype=torch.int32, pin_memory=True)
        off_e2e = off32_pin.numpy(); off_e2e[:] = off_np
    else:
        off_e2e = off_h
    out_h = out_pin.numpy().view(F.MATCH_DTYPE)

    corpus = F.Corpus.from_arrow(data_h, off_h, device=local)
    matcher = F.Matcher(WORKLOAD["needle"], cfg)
    info = matcher.backend_info()
    index_offset = rank * n

    runner = parallel.ShardRunner(matcher, corpus, index_offset, device=local)

    def step():
        # Matcher::match_list (N = 1) / match_list_parallel (N > 1): local pipeline, one all-gather of the
        # per-shard runs, device merge; the ordered list stays in HBM
        merged, total = runner.step()
        return total

    def step_host_out():
        return len(matcher.match_list_array(corpus, device=local, out=out_h))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        n_matches = step()
    if world == 1:
        n_matches = step_host_out()

    # ---- parity in the same run (outside the timed region): a prefix sample against the oracle
    parity = None
    if rank == 0 and world == 1:
        from oracle import pyoracle as O
        sub = min(n, 200_000)
        want = O.match_list_packed([WORKLOAD["needle"]], cfg.with_(emulate_lanes=info["prefilter_lanes"]),
                                   data_np[: int(off_np[sub])], off_np[: sub + 1])
        got = out_h[:n_matches]
        got = got[got["index"] < sub]
        want_s, got_s = np.sort(want, order=["index"]), np.sort(got, order=["index"])
        mism = int(len(want_s) != len(got_s)) if len(want_s) != len(got_s) else int(
            sum(int(np.count_nonzero(want_s[f] != got_s[f])) for f in ("index", "score", "exact")))
        parity = {"haystacks_checked": sub, "matches_checked": int(len(want_s)), "mismatches": mism}

    # ---- timed region: device value
    sampler = ClockSampler(local)
    stage_ms = np.zeros(4)
    launches = 0
    barrier()
    sampler.start()
    # nvidia-smi needs ~100 ms per sample: keep the GPU under the same load before (pre-roll) and after
    # (post-roll) the timed steps so that the samples describe the clocks the timed region ran at
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.4:
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        n_matches = step()
    ev1.record()
    barrier()
    t_post = time.perf_counter()
    while time.perf_counter() - t_post < 0.4:
        step()
    barrier()
    # per-stage CUDA-event timings (recorded inside the library on the launching stream) from separate,
    # identical steps, so that reading them back does not put a host sync inside the timed region
    stage_steps = 10
    for _ in range(stage_steps):
        step()
        t = matcher.last_timings()
        stage_ms += np.array([t["prefilter_ms"], t["sw_ms"], t["sort_ms"], t["total_ms"]])
        launches += t["launches"]
    stage_ms *= args.steps / stage_steps
    launches = int(launches * args.steps / stage_steps)
    clocks = sampler.stop()
    clocks["window"] = "0.4 s identical pre-roll + timed region + 0.4 s identical post-roll"
    if world == 1:
        n_matches = int(runner.count.item())
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = n * world * args.steps / (ms / 1e3)

    # ---- N = 1 only: the same call with the match list landing in pinned host memory
    host_out = None
    if world == 1:
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        h0.record()
        for _ in range(args.steps):
            step_host_out()
        h1.record()
        torch.cuda.synchronize(dev)
        hms = h0.elapsed_time(h1)
        host_out = {"value": n * args.steps / (hms / 1e3), "unit": "haystacks/s", "ms_per_step": hms / args.steps,
                    "d2h_bytes_per_step": int(n_matches * 8 + 64)}

    # ---- timed region: end to end (host buffers in, host matches out), single-GPU API per rank
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    matcher.match_list_host_array(data_h, off_e2e, device=local, out=out_h)  # warm
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        r = matcher.match_list_host_array(data_h, off_e2e, device=local, out=out_h)
    e1.record()
    barrier()
    e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
    e_ms = float(e_ms.item())
    e2e_value = n * world * e2e_steps / (e_ms / 1e3)
    h2d = int(data_h.nbytes + off_e2e.nbytes)
    d2h = int(len(r) * 8 + 64)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        pf_ms = stage_ms[0] / args.steps
        alg_bytes = int(corpus.total_bytes + 8 * n + 8 * n_matches / world)
        achieved = alg_bytes / (pf_ms / 1e3) / 1e9 if pf_ms > 0 else None
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if (WORKLOAD["needle"], WORKLOAD["max_typos"], WORKLOAD["mu"], WORKLOAD["max_len"], n) == ("deadbeef", 1, 48, 64, 10_000_000):
                traffic = int(tj["dram_bytes_read"] + tj["dram_bytes_write"])   # from the committed ncu capture
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_prefilter (+ tile rank/scan)", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": pf_ms,
                    "stage_ms_per_step": {"prefilter": pf_ms, "smith_waterman": stage_ms[1] / args.steps,
                                          "sort": stage_ms[2] / args.steps, "device_total": stage_ms[3] / args.steps}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline as cb
            threads = cb.host_threads()
            sample = args.cpu_sample or n
            ccfg = cfg.with_(emulate_lanes=info["prefilter_lanes"])
            sd, so = data_np[: int(off_np[sample])], off_np[: sample + 1]
            best = None
            for cand_threads in sorted({threads, max(1, threads // 2), max(1, threads // 4)}, reverse=True):
                dt_c, _ = cb.timed([WORKLOAD["needle"]], ccfg, sd, so, cand_threads, repeats=2)
                if best is None or dt_c < best[0]:
                    best = (dt_c, cand_threads)
            dt, threads = best
            cpu = {"value": sample / dt, "unit": "haystacks/s", "cores": threads, "kind": "port",
                   "sample": f"first {sample} haystacks of the same list, best of 2 at the fastest of all/half/quarter "
                             f"of the {cb.host_threads()} host threads; {cb.describe()}, threaded like "
                             f"match_list_parallel; emulating the {info['prefilter_lanes']}-lane reference backend"}
        line = {"metric": "haystacks/sec", "value": value, "unit": "haystacks/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": config_block(args, world, {"matches_per_step": int(n_matches),
                                                     "emulated_reference_backend": info}),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "haystacks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "ms_per_step": e_ms / e2e_steps,
                        "input": f"Arrow {'Utf8 (int32' if off_e2e.dtype.itemsize == 4 else 'LargeUtf8 (int64'} offsets) "
                                 "value+offset buffers in pinned host memory; H2D chunks overlap the pack kernels"},
                "value_host_out": host_out,
                "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu, "parity": parity}
        print(json.dumps(line))
    corpus.close()
    matcher.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    WORKLOAD.update(needle=args.needle, max_typos=args.max_typos, mu=args.mu, max_len=args.max_len)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
