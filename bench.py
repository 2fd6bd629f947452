#!/usr/bin/env python3
"""bench.py -- env steps/sec of the batched LOB + tile-coded TD hot path on B200.

One bench "step" = one launch of the fused tick kernel: TICKS_PER_LAUNCH market ticks for every
one of the B environments (about TICKS/K learner steps per env).  The metric is BASELINE.json's:
env steps/sec, one env step = one experiment::serial::Learner::_step (src/experiment/serial.cpp:53-70).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs B] [--algo ...]

N > 1 is launched by torchrun (one rank per GPU); envs are sharded by rank, independent policies
need no data-path collective ("scaling": "weak").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TICKS_PER_LAUNCH = 64
WORKLOADS = {
    # BASELINE.json configs[1]: 4096 parallel LOBs, Q-learning tile coding, synthetic Poisson flow, 1xB200
    "C1": dict(envs=4096, algo="q_learn", memory_size=65536, shared=False),
    # configs[2]: 65536 LOBs, SARSA(lambda) with eligibility traces, 1xB200
    "C2": dict(envs=65536, algo="sarsa", memory_size=16384, shared=False),
    # configs[3]: 262144 LOBs over 8 GPUs (32768 per GPU), shared policy, per-tick NCCL all-reduce of dtheta
    "C3": dict(envs=32768, algo="q_learn", memory_size=1 << 22, shared=True),
    # configs[4]: 1M LOBs over 8 GPUs (131072 per GPU), independent policies, no collective
    "C4": dict(envs=131072, algo="q_learn", memory_size=4096, shared=False),
}


def algorithmic_bytes_per_step(k_ticks, z_traces, double_q=False):
    """B_step = 2*S_env + K*(B_msg + B_ring) + B_q + 28*Z with S_env=640,
    B_msg=96, B_ring=112, B_q = 8*A*G*T*2 = 13824 (doubled for Double-Q)."""
    return 1280.0 + 208.0 * k_ticks + (27648.0 if double_q else 13824.0) + 28.0 * z_traces


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.03)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def make_cfg(workload, n_envs, env_index0, source, args):
    from rl_markets_b200 import abi, config
    w = WORKLOADS[workload]
    y = config.example_dict(**{"learning.memory_size": args.memory_size or w["memory_size"],
                               "learning.algorithm": args.algo or w["algo"]})
    # dt_ms = 1: 27e6 ticks per synthetic trading day, so no env reaches the close inside a bench run
    cfg = config.from_dict(y, n_envs=n_envs, env_index0=env_index0, source=source, flow_seed=2024, dt_ms=1,
                           shared_policy=bool(w.get("shared")))
    return y, cfg


def run_ours(args):
    import torch
    from rl_markets_b200 import abi, lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOADS[args.workload]
    B = args.envs or w["envs"]
    y, cfg = make_cfg(args.workload, B, rank * B, abi.SOURCE_GENERATOR, args)
    cfg.device = local
    m = lib.BatchedMarket(cfg)
    stream = torch.cuda.Stream()
    m.set_stream(stream.cuda_stream)
    ticks = args.ticks
    shared = bool(w.get("shared"))
    from rl_markets_b200 import parallel
    dist_mod = None
    if world > 1:
        import torch.distributed as dist_mod  # noqa: F811

    def run_chunk():
        if shared and world > 1:
            with torch.cuda.stream(stream):
                parallel.run_shared_policy(m, ticks, dist_mod)
        else:
            m.run_ticks(ticks)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- cold start (informational): the first steps of training from all-zero weight tables.  They are cheap -- the
    # learner kernel skips gathers of weights that were never written -- and NOT representative of a training run,
    # which keeps theta across ~1000 episodes (src/main.cpp:47-80): reported next to the headline, never as it.
    cold = None
    if args.pretrain_ticks > 0:
        for _ in range(3):
            run_chunk()
        m.sync()
        cc_a = m.counters()
        n_cold = min(args.steps, 20)
        ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            ev_a.record(stream)
            for _ in range(n_cold):
                run_chunk()
            ev_b.record(stream)
        barrier()
        m.sync()
        cold = {"ms": ev_a.elapsed_time(ev_b), "steps": m.counters().steps - cc_a.steps, "n": n_cold}
        # ---- pre-training (untimed): one reference trading day (LSE, 250 ms rows: 108 000 ticks) so that the timed
        # region sees weight tables in their long-run state
        left = args.pretrain_ticks
        while left > 0:
            if shared and world > 1:
                with torch.cuda.stream(stream):
                    parallel.run_shared_policy(m, min(left, 256), dist_mod)
            else:
                m.run_ticks(min(left, 256))
            left -= 256
        m.sync()
    occ_start = None
    if not shared:
        o = m.occupancy()
        occ_start = sum(o) / len(o) / float(args.memory_size or w["memory_size"])

    # ---- device-resident run (inputs = generator state, theta, traces: all in HBM)
    for _ in range(max(args.warmup, 3)):
        run_chunk()
    m.sync()
    c0 = m.counters()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with torch.cuda.stream(stream):
        evs[0].record(stream)
        for i in range(args.steps):
            run_chunk()
            evs[i + 1].record(stream)
    barrier()
    sampler.stop_flag = True
    m.sync()
    c1 = m.counters()
    total_ms = evs[0].elapsed_time(evs[-1])
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if rank == 0 and os.environ.get("RLM_BENCH_TRACE"):  # how the step time moves as the weight tables fill up
        sys.stderr.write("ms per bench step: " + " ".join("%.2f" % x for x in per_launch_ms[::max(1, args.steps // 40)]) + "\n")
    steps_done = c1.steps - c0.steps
    ticks_done = c1.ticks - c0.ticks
    z_sum = c1.sum_traces - c0.sum_traces
    launches_timed = c1.kernel_launches - c0.kernel_launches

    # ---- per-kernel durations (CUDA events around every launch, on the launching stream): a separate pass,
    # because the events serialise host launch and device execution; not used for `value`
    kt = None
    if not shared:
        m.set_profiling(True)
        cp0 = m.counters()
        for _ in range(min(args.steps, 5)):
            m.run_ticks(ticks)
        m.sync()
        cp1 = m.counters()
        kt = m.kernel_times()
        kt["steps"] = cp1.steps - cp0.steps
        kt["ticks"] = cp1.ticks - cp0.ticks
        kt["sum_traces"] = cp1.sum_traces - cp0.sum_traces
        m.set_profiling(False)

    # ---- end to end through the C ABI with HOST buffers: every launch uploads its tick messages from
    # pinned host memory (rlm_load_ticks) and reads the per-env rewards back (rlm_get_reward)
    e2e = None
    if not args.no_e2e and not shared:
        y2, cfg2 = make_cfg(args.workload, B, rank * B, abi.SOURCE_STREAM, args)
        cfg2.device = local
        m2 = lib.BatchedMarket(cfg2)
        m2.set_stream(stream.cuda_stream)
        m2.copy_theta_from(m)  # same long-run weight tables as the device-resident leg (a new data day, trained agent)
        n_warm = max(args.warmup, 3)
        e2e_steps_n = min(args.steps, 20)  # 32 MB of pinned host memory per step at C1: the e2e leg times at most 20 of them
        nbytes = ticks * B * C.sizeof(abi.TickMsg)
        gen_ticks = ticks * (n_warm + e2e_steps_n)
        # synthetic messages for all envs, generated once on the host cores and staged in pinned host memory
        # BEFORE the timed region (one pinned chunk per bench step; nothing host-side is excluded from the timing)
        import numpy as np
        per_env = [lib.flow_generate(cfg2.flow, rank * B + b, 0, gen_ticks) for b in range(min(B, args.e2e_distinct))]
        env_np = [np.frombuffer(pe, dtype=np.uint8).reshape(gen_ticks, 128) for pe in per_env]
        rew = (C.c_double * B)()

        def staged(chunk):
            host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            host_np = host.numpy().view(np.uint8).reshape(ticks, B, 128)
            for b in range(B):  # envs beyond e2e_distinct replay the stream of env (b mod distinct)
                host_np[:, b, :] = env_np[b % len(env_np)][chunk * ticks:(chunk + 1) * ticks]
            return host

        for ch in range(n_warm):
            host = staged(ch)
            m2.load_ticks(host.data_ptr(), ticks)
            m2.run_ticks(ticks)
            m2.sync()
        chunks = [staged(n_warm + i) for i in range(e2e_steps_n)]
        cc0 = m2.counters()
        barrier()
        t0 = time.perf_counter()
        # rlm_load_ticks double-buffers on a copy stream: the upload of step i+1 is issued before the result of
        # step i is read, so it overlaps step i's kernels; all K uploads and K read-backs are inside the region
        m2.load_ticks(chunks[0].data_ptr(), ticks)
        for i in range(e2e_steps_n):
            m2.run_ticks(ticks)
            if i + 1 < e2e_steps_n:
                m2.load_ticks(chunks[i + 1].data_ptr(), ticks)   # H2D inside the timed region
            lib.check(m2.L.rlm_get_reward(m2.h, rew))  # D2H inside the timed region (syncs)
        barrier()
        wall = time.perf_counter() - t0
        cc1 = m2.counters()
        e2e_steps = cc1.steps - cc0.steps
        e2e = {"steps": e2e_steps, "seconds": wall, "h2d": nbytes, "d2h": B * 8, "n": e2e_steps_n}
        m2.close()

    # ---- aggregate over ranks: max time, sum steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([total_ms, (e2e or {}).get("seconds", 0.0)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([steps_done, ticks_done, z_sum, (e2e or {}).get("steps", 0)], device="cuda", dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        total_ms = float(t[0]); steps_all, ticks_all, z_all = float(s[0]), float(s[1]), float(s[2])
        if e2e:
            e2e["seconds"] = float(t[1]); e2e["steps"] = float(s[3])
    else:
        steps_all, ticks_all, z_all = float(steps_done), float(ticks_done), float(z_sum)

    if cold is not None:
        if world > 1:
            import torch.distributed as dist
            tc = torch.tensor([cold["ms"]], device="cuda", dtype=torch.float64)
            sc = torch.tensor([float(cold["steps"])], device="cuda", dtype=torch.float64)
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            dist.all_reduce(sc, op=dist.ReduceOp.SUM)
            cold["ms"], cold["steps"] = float(tc[0]), float(sc[0])
    if rank == 0:
        k_bar = ticks_all / max(steps_all, 1.0)
        z_bar = z_all / max(steps_all, 1.0)
        is_dq = (args.algo or w["algo"]) == "double_q_learn"
        b_step = algorithmic_bytes_per_step(k_bar, z_bar, is_dq)
        value = steps_all / (total_ms * 1e-3)
        peak, peak_src = measured_peak_gbs()
        # dominant kernel = rlm_agent3_kernel (one launch per market tick); its algorithmic bytes are the agent
        # share of B_step: theta gathers + trace list (13824 + 28 Z per env step), SURVEY.md section 8d
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r1_summary.json")) as f:
                traffic = json.load(f).get(args.workload, {}).get("agent_kernel_dram_bytes_per_launch")
        except Exception:
            pass
        if kt and kt["agent_launches"]:
            avg_launch_ms = kt["agent_ms"] / kt["agent_launches"]
            launch_steps = kt["steps"] / kt["agent_launches"]
            z_k = kt["sum_traces"] / max(kt["steps"], 1)
            b_agent = (27648.0 if is_dq else 13824.0) + 28.0 * z_k
            achieved = launch_steps * b_agent / (avg_launch_ms * 1e-3) / 1e9
            env_avg_ms = kt["env_ms"] / max(kt["env_launches"], 1)
        else:
            avg_launch_ms = total_ms / max(launches_timed, 1)
            launch_steps = steps_done / max(launches_timed, 1)
            b_agent = b_step
            achieved = steps_done * b_step / (total_ms * 1e-3) / 1e9
            env_avg_ms = None
        line = {
            "metric": "env steps/sec (batched LOBs)", "value": value, "unit": "env_steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d parallel LOBs per GPU, %s + tile coding (32 tilings, memory_size %d per env), "
                                   "synthetic Poisson order flow (in-kernel generator), %d ticks per launch"
                                   % (args.workload, B, args.algo or w["algo"], args.memory_size or w["memory_size"], ticks),
                       "envs_per_gpu": B, "ticks_per_bench_step": ticks, "mean_ticks_per_step": k_bar, "mean_nonzero_traces": z_bar,
                       "policy": "shared theta, one SUM all-reduce of dtheta per tick" if shared else "independent theta per env, no collective",
                       "l2": ("working set (theta %.1f GB per GPU) is larger than L2; no flush needed" % (B * (args.memory_size or w["memory_size"]) * 8 / 1e9))
                             if not shared else ("shared theta %.0f MB (L2-resident) + %.1f GB of env records, traces and generator state" % ((args.memory_size or w["memory_size"]) * 8 / 1e6, B * 8.0e3 / 1e9)),
                       "ticks_per_s": ticks_all / (total_ms * 1e-3),
                       "pretrain_ticks": args.pretrain_ticks,
                       "theta_occupancy_at_start": occ_start,
                       "state": ("timed after %d ticks of training per env (one LSE trading day of the reference's 250 ms rows): "
                                 "long-run weight tables" % args.pretrain_ticks) if args.pretrain_ticks > 0
                                else "cold start: all-zero weight tables (gathers of never-written weights are skipped)"},
            "gpu_launches": int(launches_timed),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "kernel": "rlm_agent3_kernel (learner step, one launch per market tick)",
                         "algorithmic_bytes_per_env_step_agent_kernel": b_agent,
                         "algorithmic_bytes_per_env_step_whole_path": b_step,
                         "whole_path_achieved_GBps": steps_all * b_step / (total_ms * 1e-3) / 1e9 / max(world, 1),
                         "env_steps_per_launch": launch_steps, "avg_launch_ms": avg_launch_ms,
                         "env_kernel_avg_launch_ms": env_avg_ms,
                         "timing": "CUDA events around every kernel launch on the launching stream, separate pass"},
        }
        if cold is not None:
            line["cold_start"] = {"value": cold["steps"] / (cold["ms"] * 1e-3), "unit": "env_steps/s", "steps": cold["n"],
                                  "note": "first bench steps of training from all-zero weight tables; informational, not the headline"}
        if e2e:
            line["e2e"] = {"value": e2e["steps"] / e2e["seconds"], "unit": "env_steps/s",
                           "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "steps": e2e["n"],
                           "note": "STREAM source: rlm_load_ticks from pinned host memory (double-buffered upload) + rlm_run_ticks + rlm_get_reward per step"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_reference(args, n_procs=1, ticks=args.cpu_ticks)
        print(json.dumps(line))
    m.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
def _ref_paths():
    ref = os.path.join(ROOT, "oracle", "_ref")
    return os.path.join(ref, "ref_driver"), os.path.join(ref, "flow_csv")


def cpu_baseline_reference(args, n_procs, ticks):
    """Times the reference's own CPU loop: oracle/_ref/ref_driver = the UNMODIFIED reference sources
    compiled in the build container (oracle/Makefile), one single-threaded process per env (the
    reference's only deterministic mode, src/main.cpp:196-209), CSV parsing included, files on tmpfs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from rl_markets_b200 import config
    drv, flow = _ref_paths()
    w = WORKLOADS[args.workload]
    algo = args.algo or w["algo"]
    y = config.example_dict(**{"learning.memory_size": args.memory_size or w["memory_size"], "learning.algorithm": algo})
    if not os.path.exists(drv):
        # fall back to the CPU restatement ("port"), threads = n_procs
        from rl_markets_b200 import abi
        cfg = config.from_dict(y, n_envs=n_procs, flow_seed=2024, dt_ms=1)
        L = oracle_lib.lib()
        tt, secs = C.c_int64(0), C.c_double(0)
        steps = L.lobo_run_batch(C.byref(cfg), n_procs, ticks, n_procs, C.byref(tt), C.byref(secs))
        return {"value": steps / secs.value, "unit": "env_steps/s", "cores": n_procs, "kind": "port",
                "sample": "%d env(s) x %d ticks through oracle/liblob_oracle.so (in-memory ticks)" % (n_procs, ticks)}
    tmp = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=tmp) as d:
        cfgp = _ref_prepare(d, y, flow, n_procs, ticks)
        res = _ref_run(d, cfgp, drv, n_procs)
    return _ref_result(res, n_procs, ticks, algo)


def _ref_prepare(d, y, flow, n_procs, ticks):
    """Config + one synthetic CSV pair per process (written in parallel, on tmpfs, outside any timed region)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    cfgp = os.path.join(d, "cfg.yaml")
    oracle_lib.write_ref_yaml(cfgp, y)
    gens = []
    for i in range(n_procs):
        md, tas = os.path.join(d, "e%d_md_1.csv" % i), os.path.join(d, "e%d_tas_1.csv" % i)
        gens.append(subprocess.Popen([flow, "--seed", "2024", "--env", str(i), "--ticks", str(ticks), "--dt-ms", "1", "--md", md, "--tas", tas]))
    for g in gens:
        if g.wait() != 0:
            raise RuntimeError("flow_csv failed")
    return cfgp


def _ref_run(d, cfgp, drv, n_procs):
    t0 = time.perf_counter()
    procs = []
    for i in range(n_procs):
        md, tas = os.path.join(d, "e%d_md_1.csv" % i), os.path.join(d, "e%d_tas_1.csv" % i)
        procs.append(subprocess.Popen([drv, "--config", cfgp, "--md", md, "--tas", tas], stdout=subprocess.PIPE))
    outs = [p.communicate()[0] for p in procs]
    wall = time.perf_counter() - t0
    steps, inner = 0, 0.0
    for o in outs:
        r = json.loads(o.decode().strip().splitlines()[-1])
        steps += r["steps"]
        inner = max(inner, r["seconds"])
    return steps, inner, wall


def _ref_result(res, n_procs, ticks, algo):
    steps, inner, wall = res
    return {"value": steps / inner, "unit": "env_steps/s", "cores": n_procs, "kind": "reference",
            "sample": "%d process(es) x %d synthetic ticks (%d learner steps), %s, CSV parsing included, "
                      "timed inside ref_driver (max over processes %.2fs; wall incl. process start %.2fs)"
                      % (n_procs, ticks, steps, algo, inner, wall)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_procs = os.cpu_count() or 1
    drv, flow = _ref_paths()
    n_runs = max(args.warmup, 0) + args.steps
    # one bench step = every host thread runs one single-env reference process over `ticks` ticks of the workload;
    # the sample is sized so that the whole --steps/--warmup run stays within a few minutes (~1e5 ticks/s per process)
    ticks = min(args.ref_ticks, max(10000, int(6e6 / max(n_runs, 1))))
    vals = []
    last = None
    if not os.path.exists(drv):
        for i in range(n_runs):
            last = cpu_baseline_reference(args, n_procs=n_procs, ticks=ticks)  # oracle port fallback
            if i >= args.warmup:
                vals.append(last)
    else:
        from rl_markets_b200 import config
        w0 = WORKLOADS[args.workload]
        algo = args.algo or w0["algo"]
        y = config.example_dict(**{"learning.memory_size": args.memory_size or w0["memory_size"], "learning.algorithm": algo})
        tmp = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=tmp) as d:
            cfgp = _ref_prepare(d, y, flow, n_procs, ticks)
            for i in range(n_runs):
                last = _ref_result(_ref_run(d, cfgp, drv, n_procs), n_procs, ticks, algo)
                if i >= args.warmup:
                    vals.append(last)
    value = sum(v["value"] for v in vals) / len(vals)
    w = WORKLOADS[args.workload]
    line = {"impl": "reference", "metric": "env steps/sec (batched LOBs)", "value": value, "unit": "env_steps/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "%s: %s + tile coding (memory_size %d per env), synthetic Poisson order flow; reference arm = "
                                   "%d independent single-threaded reference processes on the host cores"
                                   % (args.workload, args.algo or w["algo"], args.memory_size or w["memory_size"], n_procs)},
            "cpu_baseline": {"value": value, "unit": "env_steps/s", "cores": n_procs, "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": value, "unit": "env_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C1", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0)
    ap.add_argument("--algo", default="")
    ap.add_argument("--memory-size", dest="memory_size", type=int, default=0)
    ap.add_argument("--ticks", type=int, default=TICKS_PER_LAUNCH)
    ap.add_argument("--pretrain-ticks", dest="pretrain_ticks", type=int, default=108000,
                    help="untimed training before the timed region (default: one LSE day of 250 ms rows); 0 = cold start")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-distinct", type=int, default=64, help="distinct host-generated streams replayed across envs in the e2e leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-ticks", type=int, default=400000)
    ap.add_argument("--ref-ticks", type=int, default=100000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
