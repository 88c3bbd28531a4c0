#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Groth16 hot path.

Metric (BASELINE.json): G1 MSM throughput in Mpairs/s on BN254, 2^20 random scalar/point pairs per
GPU (configs[1]); one "step" = one d_msm over one resident batch.  N > 1: every rank owns its own
2^20 pairs (weak scaling), the only exchange is d_msm's all-gather of the N XYZZ partials + point sum.

  value  : pairs processed by all ranks / device time of the step (inputs resident in HBM)
  e2e    : same metric through the reference-facing call with HOST buffers
           (pinned host -> H2D of bases+scalars, MSM, D2H of the affine result) inside the timed region
  roofline: msm_accumulate_g1 (dominant kernel): 96 B/pair algorithmic bytes / its CUDA-event duration
  cpu_baseline: oracle/bn254_ref.cpp (arkworks-equivalent CPU restatement, "port") on the host cores

`--impl reference` times that CPU restatement alone (the reference itself is Rust and cannot be built
in this image: no cargo/rustc, dependencies un-vendored -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = int(os.environ.get("B200ZK_BENCH_LOG_N", "20"))
METRIC = "G1 MSM throughput (BN254 Pippenger, 2^%d pairs per GPU)" % LOG_N
WORKLOAD = "BN254 G1 Pippenger MSM 2^%d random scalar/point pairs per GPU" % LOG_N     # same string in both arms
UNIT = "Mpairs/s"
ALG_BYTES_PER_PAIR = 96.0       # 32 B scalar + 64 B affine point, each read once


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """CPU arm: the arkworks-equivalent restatement on the host cores, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np  # noqa: F401
    from oracle import cref
    cref.build()
    n = 1 << LOG_N
    cores = os.cpu_count() or cref.num_threads()      # torchrun exports OMP_NUM_THREADS=1: ask for all host cores explicitly
    bases = cref.g1_generate(0xB2000002, n)
    scalars = cref.fr_generate(0xB2000002, n)
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        cref.msm_g1(bases, scalars, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cref.msm_g1(bases, scalars, cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e6
    sample = "full 2^%d-pair G1 MSM per step, %d OpenMP threads (windows in parallel, as arkworks+rayon)" % (LOG_N, cores)
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32x8 Montgomery (256-bit modular integers)", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "CPU restatement of arkworks VariableBaseMSM (oracle/bn254_ref.cpp); the Rust reference "
                           "cannot be built in this image"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def measure_ntt(net, hbm_peak, pipe_peak):
    """BASELINE config 3: Fr radix-2 NTT, 2^22 elements resident in HBM (forward, natural order in and out)."""
    import torch
    log_n = int(os.environ.get("B200ZK_BENCH_NTT_LOG_N", "22"))
    n = 1 << log_n
    x = net.generate_fr(3, n)
    y = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=x.device)
    for _ in range(3):
        net.ntt_dev(x, y)
    evs = []
    for _ in range(10):
        flush.fill_(1)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        net.ntt_dev(x, y)
        a1.record()
        evs.append((a0, a1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    ms = sum(ts) / len(ts)
    back = torch.empty_like(x)
    net.ntt_dev(y, back, inverse=True)
    passes = -(-log_n // 8)
    products = n * (log_n / 2.0 + 2.0 * (passes - 1))
    return {"metric": "Fr NTT 2^%d (BN254 scalar field)" % log_n, "ms": ms, "ms_min": ts[0], "gelem_s": n / ms / 1e6,
            "roofline": {"bound": "hbm", "achieved": 64.0 * n / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                         "frac": 64.0 * n / ms / 1e6 / hbm_peak, "algorithmic_bytes_per_element": 64,
                         "pipe": {"achieved": products / ms / 1e6, "peak": pipe_peak, "unit": "G modular products/s",
                                  "frac": products / ms / 1e6 / pipe_peak,
                                  "how": "n x (log n / 2 butterflies + 2 twiddle products per element per pass boundary)"}},
            "round_trip_exact": bool((back == x).all()),
            "timing": "CUDA events per transform, L2 flushed between transforms, 10 runs after 3 warm-ups"}


def measure_prove(net, args, with_cpu):
    """Secondary metric of BASELINE.json: Groth16 prove ms, BN254, 2^20 constraints (m = n_vars = 2^20, dummy CRS
    built like groth16/examples/local_groth_bench.rs:21-52; witness and QAP evaluations resident in HBM)."""
    import numpy as np
    import torch
    from distributed_groth16_b200.groth16 import ProvingKey, prove
    log_m = int(os.environ.get("B200ZK_BENCH_PROVE_LOG_M", "20"))
    m = 1 << log_m
    n_vars, n_inputs = m, 2
    aq, b1 = net.generate_g1(101, n_vars), net.generate_g1(102, n_vars)
    b2 = net.generate_g2(103, n_vars)
    lq, hq = net.generate_g1(104, n_vars - n_inputs), net.generate_g1(105, m)
    vk = np.concatenate([net.generate_g1(106, 3).cpu().numpy().view(np.uint64).reshape(-1),
                         net.generate_g2(107, 2).cpu().numpy().view(np.uint64).reshape(-1)])
    z = net.generate_fr(108, n_vars)
    from distributed_groth16_b200._constants import FR_ONE_MONT
    z[0] = torch.from_numpy(np.array(FR_ONE_MONT, dtype=np.uint64).view(np.int64)).to(z.device)
    a, b, c = (net.generate_fr(sd, m) for sd in (109, 110, 111))
    pk = ProvingKey.from_device(net, aq, b1, b2, lq, hq, n_inputs, vk)
    times = []
    proof = None
    for _ in range(2 + 5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = prove.create_proof_dev(pk, z, a, b, c)         # returns after the 128 proof bytes are on the host
        times.append((time.perf_counter() - t0) * 1e3)
    times = sorted(times[2:])
    res = {"metric": "Groth16 prove (BN254, 2^%d constraints, r = s = 0)" % log_m, "ms": times[len(times) // 2],
           "ms_min": times[0], "unit": "ms", "higher_is_better": False, "n_vars": n_vars, "domain": m,
           "msm_sizes": {"g1": [n_vars - 1, n_vars - n_inputs, m], "g2": [n_vars - 1]},
           "pk_table_gb": pk.table_bytes / 2**30,
           "timing": "host wall clock around b200zk_groth16_prove_dev (includes the D2H of the proof), 5 runs after 2 warm-ups"}
    if with_cpu:
        from oracle import cref
        ncores = os.cpu_count() or 1
        h2 = lambda t: t.cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        hh = cref.h_circom(h2(a), h2(b), h2(c), ncores)
        exp = cref.groth16_prove(h2(aq), h2(b1), h2(b2), h2(lq), h2(hq), vk, n_inputs, h2(z), hh, np.zeros(4, np.uint64),
                                 np.zeros(4, np.uint64), nthreads=ncores)
        res["cpu_baseline"] = {"ms": (time.perf_counter() - t0) * 1e3, "cores": ncores, "kind": "port",
                               "sample": "one full prove (h + 4 MSMs) with the CPU restatement"}
        res["bit_exact_vs_cpu"] = bool(exp == proof)
    pk.free()
    return res


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL's `NCCL version ...` banner comes from C
    code on fd 1): keep a private duplicate of fd 1 for the JSON line and point fd 1 at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        try:
            sys.stdout.flush()
            fd = os.dup(1)
            os.dup2(2, 1)
            _JSON_FD = fd
        except OSError:                 # no usable stderr: keep the plain stdout
            _JSON_FD = None


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prove", action="store_true", help="skip the secondary Groth16-prove measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from distributed_groth16_b200 import Net
    from distributed_groth16_b200.dist_primitives import d_msm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warm = max(args.warmup, 3)
    n = 1 << LOG_N
    net = Net(local)
    net.use_torch_stream(0)
    dev = torch.device("cuda", local)
    bases = net.generate_g1(0xB2000002 + rank * 0x1000000, n)
    scalars = net.generate_fr(0xB2000002 + rank, n)
    part = torch.empty(16, dtype=torch.int64, device=dev)
    gathered = torch.empty((world, 16), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        net.msm_dev(bases, scalars, part)
        if world > 1:
            dist.all_gather_into_tensor(gathered, part.reshape(1, -1))
            return net.sum_points_dev(gathered, world)
        return net.sum_points_dev(part, 1)

    # pinned host copies for the e2e arm
    h_bases = torch.empty((n, 8), dtype=torch.int64).pin_memory()
    h_scalars = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    h_bases.copy_(bases)
    h_scalars.copy_(scalars)
    hb_np, hs_np = h_bases.numpy().view(np.uint64), h_scalars.numpy().view(np.uint64)
    d_b2 = torch.empty_like(bases)
    d_s2 = torch.empty_like(scalars)

    def step_e2e():
        if world == 1:
            return net.msm(hb_np, hs_np)                       # the C-ABI host-buffer call (b200zk_msm_g1)
        d_b2.copy_(h_bases, non_blocking=True)
        d_s2.copy_(h_scalars, non_blocking=True)
        net.msm_dev(d_b2, d_s2, part)
        dist.all_gather_into_tensor(gathered, part.reshape(1, -1))
        return net.sum_points_dev(gathered, world)

    def timed(fn, steps):
        times = []
        for _ in range(steps):
            flush.fill_(1)                                      # evict L2 between timed iterations
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            res = fn()
            b.record()
            b.synchronize()
            times.append(a.elapsed_time(b))
        return times, res

    for _ in range(warm):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = net.launch_count()
    t_wall0 = time.perf_counter()
    times, res = timed(step_resident, args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = net.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = sum(times) / len(times)
    ms_step_sorted = sorted(times)

    for _ in range(2):
        step_e2e()
    barrier()
    times_e2e, res_e2e = timed(step_e2e, max(3, min(args.steps, 10)))
    barrier()
    ms_e2e = sum(times_e2e) / len(times_e2e)
    assert (res[0] == res_e2e[0]).all()

    # steady-state throughput with the steps issued round-robin over the three stream slots, i.e. the way the
    # reference itself issues concurrent d_msm calls (MultiplexedStreamID 0..2, groth16/src/prove.rs:119-125):
    # the latency-bound tail of one MSM (bucket reduction, Horner) overlaps the bucket accumulation of the next.
    pipelined = None
    if world == 1:
        parts3 = [torch.empty(16, dtype=torch.int64, device=dev) for _ in range(3)]
        kp = max(6, args.steps)
        for _ in range(3):
            for k in range(3):
                net.msm_dev(bases, scalars, parts3[k], sid=k)
        for k in range(3):
            net.sync(k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = []
        for k in range(kp):
            if k >= 3:
                outs.append(net.sum_points_dev(parts3[k % 3], 1, sid=k % 3))      # result of step k-3 (same slot) -> host
            net.msm_dev(bases, scalars, parts3[k % 3], sid=k % 3)
        for k in range(kp, kp + 3):
            outs.append(net.sum_points_dev(parts3[k % 3], 1, sid=k % 3))
        for k in range(3):
            net.sync(k)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3 / kp
        assert all((o[0] == res[0]).all() for o in outs)
        pipelined = {"value": n / dt / 1e3, "unit": UNIT, "ms_per_step": dt, "steps": kp,
                     "how": "steps issued round-robin on the 3 stream slots; host wall clock between device syncs; "
                            "every step's affine result is copied to the host"}

    # the same MSM over fixed-base window tables (b200zk_msm_table_*): what the proving path runs, since a proving key's
    # query vectors stay resident across proofs.  Reported beside `value`, which stays the generic d_msm (fresh bases).
    fixed_base = None
    if world == 1:
        c_tab = net.msm_table_auto_window(n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        table = net.msm_table_build(bases, c_tab)
        e1.record()
        torch.cuda.synchronize()
        build_ms = e0.elapsed_time(e1)
        part_t = torch.empty(16, dtype=torch.int64, device=dev)
        for _ in range(3):
            net.msm_table_dev(table, scalars, c_tab, part_t)
        evs = []
        for _ in range(max(args.steps, 5)):
            flush.fill_(1)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            net.msm_table_dev(table, scalars, c_tab, part_t)
            a1.record()
            evs.append((a0, a1))
        torch.cuda.synchronize()
        ts = sorted(x.elapsed_time(y) for x, y in evs)
        ms_t = sum(ts) / len(ts)
        got_t = net.sum_points_dev(part_t, 1)
        fixed_base = {"value": n / ms_t / 1e3, "unit": UNIT, "ms_per_step": ms_t, "ms_per_step_min": ts[0], "window": c_tab,
                      "windows": net.msm_table_windows(c_tab), "table_gb": table.numel() * 8 / 2**30, "table_build_ms": build_ms,
                      "bit_exact_vs_generic": bool((got_t[0] == res[0]).all()),
                      "how": "table[w*n+i] = 2^(c w) P_i resident in HBM; one bucket set, no Horner doublings; "
                             "CUDA events per step, L2 flushed between steps"}
        del table

    # per-kernel CUDA-event durations (separate short pass: the event pairs add launch gaps)
    net.profile(True)
    net.profile_reset()
    for _ in range(5):
        flush.fill_(1)
        step_resident()
    torch.cuda.synchronize()
    rep = net.profile_report()
    net.profile(False)

    if world > 1:
        t = torch.tensor([ms_step, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peaks()
    # the bucket kernel runs once per window group (msm.cu, window-group pipeline): each launch processes 1/G of every
    # pair's digits, i.e. 96 n / G algorithmic bytes, in 1/G of the step's bucket time -- same GB/s either way
    acc = rep.get("msm_accumulate_g1", {"launches": 5, "ms": float("nan")})
    acc_launches_per_step = max(acc["launches"], 1) / 5.0
    acc_ms_launch = acc["ms"] / max(acc["launches"], 1)
    acc_ms = acc["ms"] / 5.0                                                  # per step, all launches
    achieved = ALG_BYTES_PER_PAIR * n / acc_launches_per_step / (acc_ms_launch * 1e-3) / 1e9
    kernel_ms = {k: round(v["ms"] / 5.0, 4) for k, v in rep.items()}
    # what actually bounds the kernel: the multiplier pipe.  A 256-bit Montgomery product = 128 32-bit wide multiply-adds
    # with carry, which issue at 32 lanes/clk/SM (tools/microbench.cu, profiles/r1_microbench_pipes.txt).
    sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
    pipe_peak = 32.0 * torch.cuda.get_device_properties(dev).multi_processor_count * sm_mhz * 1e6 / 128.0 / 1e9          # G products/s
    adds = n * 16.0 * (1.0 - 2.0 ** -16)                                      # 16 signed 16-bit digits per scalar (GLV: 2 x 8)
    pipe_ach = adds * 10.0 / (acc_ms * 1e-3) / 1e9                            # XYZZ mixed addition = 8M + 2S
    pipe = {"bound": "fmaheavy (32-bit multiply-add with carry)", "achieved": pipe_ach, "peak": pipe_peak,
            "unit": "G modular products/s", "frac": pipe_ach / pipe_peak,
            "how": "bucket additions (n x 16 digits) x 10 products / kernel time; peak = 32 lanes/clk/SM x SMs x SM clock / 128"}
    out = {
        "metric": METRIC, "value": world * n / ms_step / 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "ms_per_step_median": ms_step_sorted[len(ms_step_sorted) // 2],
        "ms_per_step_min": ms_step_sorted[0], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 Montgomery (256-bit modular integers)", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "pairs_per_gpu": n, "l2": "256 MiB flush write between timed iterations; inputs+workspace > L2",
                   "parallelism": "length-sharded x%d, all-gather of XYZZ partials" % world},
        "e2e": {"value": world * n / ms_e2e / 1e3, "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": n * 96, "d2h_bytes_per_step": 72},
        "gpu_launches": int(launches),
        "wall_s_timed_region": t_wall,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "msm_accumulate_g1", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": acc_ms_launch,
                     "launches_per_step": acc_launches_per_step, "kernel_ms_per_step": acc_ms,
                     "algorithmic_bytes_per_launch": ALG_BYTES_PER_PAIR * n / acc_launches_per_step,
                     "note": "256-bit modular integer arithmetic: IMAD-bound by construction, HBM fraction is small",
                     "pipe": pipe},
        "kernel_ms_per_step": kernel_ms,
        "pipelined": pipelined,
        "fixed_base": fixed_base,
    }
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            out["roofline"]["traffic"] = tj.get("msm_accumulate_g1_bytes_per_launch")
            out["roofline"]["traffic_source"] = "static: ncu --set full capture recorded in profiles/traffic.json (%s), not measured in this run" % tj.get("source", "see profiles/README.md")
        except Exception:
            pass
    if not args.no_cpu_baseline:
        from oracle import cref                                  # cpu_baseline leg: the checker timed as a baseline
        cref.build()
        hb = bases.cpu().numpy().view(np.uint64)
        hs = scalars.cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        ncores = os.cpu_count() or cref.num_threads()
        exp, _ = cref.msm_g1(hb, hs, ncores)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt / 1e6, "unit": UNIT, "cores": ncores, "kind": "port",
                               "sample": "one full 2^%d-pair G1 MSM, all host threads (%.2f s)" % (LOG_N, dt),
                               "bit_exact_vs_gpu": bool(world == 1 and (exp == res[0]).all()) if world == 1 else None}
    if world == 1:
        out["ntt"] = measure_ntt(net, peak, pipe_peak)
    if world == 1 and not args.no_prove:
        out["prove"] = measure_prove(net, args, not args.no_cpu_baseline)
    emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
