#!/usr/bin/env python
"""SAC gradient steps/s on B200 (BASELINE.json metric), with roofline, CPU baseline and e2e.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload LL|VS|MS] [--replicas R]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference's CPU learner path (oracle port) on host cores

One "step" = one full SAC gradient step (target, twin-critic update, actor update, temperature
update, Polyak) of every learner replica on the GPU.  `value` = learner-steps/s with the replay
ring resident in HBM (sampling + gather + step all on device, one CUDA graph per step);
`e2e` = the same through the reference-shaped `Learner.update()` with the replay ring in pinned
HOST memory: per step a host-side sample/gather, an H2D copy of the minibatch and a D2H read of
the losses, synchronously, exactly like the reference's update() contract.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# algorithmic work per gradient step (streaming model)
WORK = {
    "LL": dict(gflop=0.557, mbytes=6.087),
    "VS": dict(gflop=10.970, mbytes=30.188),
    "MS": dict(gflop=13.846, mbytes=30.728),
    "C10": dict(gflop=17.088, mbytes=37.377),
    # CARE(O): same consumer MLPs and mixture as C10; the context path is per-task (10 rows) in both, so the per-step
    # streaming-model work differs only by the 89.6 k trainable context-encoder parameters
    "C10O": dict(gflop=17.088, mbytes=37.377 + 89600 * 24e-6),
}
WORKLOAD_DESC = {
    "LL": "LunarLanderContinuous-v2 SAC learner (obs 8, act 2, MLP 256-256, batch 256)",
    "VS": "MT1 VSAC-shape SAC learner (obs 39, act 4, MLP 400x3, batch 1024, twin-Q)",
    "MS": "MT10 MTSAC learner (mtobs 49, act 4, MLP 400x3, batch 1280, 10 tasks one-hot, weighted loss)",
    "C10": "MT10 CARE(M) learner (mtobs 49, act 4, K=6 mixture encoders 39-50-50, 768-d context, MLP 400x3 over the 100-d encoded state, batch 1280)",
    "C10O": "MT10 CARE(O) learner (use_modified_care=false: trainable context encoder 768-100-50-50-50-50 with its own Adam, "
            "K=6 mixture encoders, MLP 400x3, batch 1280, unweighted losses)",
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf=1590.0, src="fallback")


def core_config(workload, replicas, precision=1):
    from distributed_sac_b200.core import CoreConfig
    if workload == "LL":
        return CoreConfig(replicas=replicas, precision=precision)
    if workload == "VS":
        return CoreConfig(state_dim=39, act_dim=4, actor_hidden=[400] * 3, critic_hidden=[400] * 3, batch=1024,
                          replicas=replicas, precision=precision)
    if workload == "MS":
        return CoreConfig(state_dim=39, act_dim=4, actor_hidden=[400] * 3, critic_hidden=[400] * 3, batch=1280,
                          num_tasks=10, weighted_loss=True, replicas=replicas, precision=precision)
    if workload == "C10":
        return CoreConfig(state_dim=39, act_dim=4, actor_hidden=[400] * 3, critic_hidden=[400] * 3, batch=1280,
                          num_tasks=10, weighted_loss=True, replicas=replicas, precision=precision, care=True)
    if workload == "C10O":
        return CoreConfig(state_dim=39, act_dim=4, actor_hidden=[400] * 3, critic_hidden=[400] * 3, batch=1280,
                          num_tasks=10, weighted_loss=False, replicas=replicas, precision=precision, care=True,
                          care_original=True, emb_dim=50, lr_ctx=3e-4)
    raise ValueError(workload)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
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
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU side: the reference's learner path as restated by oracle/sac_port.py (kind "port")
# --------------------------------------------------------------------------------------------------
def cpu_learner(workload, n_buffer=20000, seed=0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sac_port as sp
    if workload in ("C10", "C10O"):
        import care_port as cp
        spec = cp.CareSpec() if workload == "C10" else cp.CareSpec(modified=False, weighted_loss=False)
        lrn = cp.CarePortLearner(spec, cp.init_params(spec, seed=seed))
        rspec = sp.ms_spec()                       # same replay layout as MTSAC: mtobs rows, B/T per task
        rb = sp.PortReplay(rspec, n_buffer, seed=seed)
        rb.push_many(*[t.numpy() for t in sp.synthetic_batch(rspec, seed=1234, batch=n_buffer)])
        return spec, lrn, rb
    spec = {"LL": sp.ll_spec, "VS": sp.vs_spec, "MS": sp.ms_spec}[workload]()
    lrn = sp.PortLearner(spec, sp.init_params(spec, seed=seed))
    rb = sp.PortReplay(spec, n_buffer, seed=seed)
    rb.push_many(*[t.numpy() for t in sp.synthetic_batch(spec, seed=1234, batch=n_buffer)])
    return spec, lrn, rb


def best_cpu_threads(workload):
    """Eager PyTorch on many cores is often slower than on a few for these tiny ops: probe a few thread
    counts briefly and use the fastest, so the CPU arm is the reference path at its best on this host."""
    ncpu = os.cpu_count() or 1
    cands = sorted({1, 4, 8, 16, min(32, ncpu), ncpu} & set(range(1, ncpu + 1)))
    best, best_v = 1, 0.0
    spec, lrn, rb = cpu_learner(workload)
    for t in cands:
        torch.set_num_threads(t)
        for _ in range(2):
            lrn.update_SAC(*rb.sample())
        t0 = time.perf_counter()
        n = 6
        for _ in range(n):
            lrn.update_SAC(*rb.sample())
        v = n / (time.perf_counter() - t0)
        if v > best_v:
            best, best_v = t, v
    return best


def time_cpu(workload, steps, warmup, threads=None):
    """steps/s of sample() + update_SAC() on the host cores."""
    torch.set_num_threads(threads or best_cpu_threads(workload))
    spec, lrn, rb = cpu_learner(workload)
    for _ in range(warmup):
        lrn.update_SAC(*rb.sample())
    t0 = time.perf_counter()
    for _ in range(steps):
        lrn.update_SAC(*rb.sample())   # (CarePortLearner.update_SAC is its update())
    dt = time.perf_counter() - t0
    return steps / dt, dt, torch.get_num_threads()


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, rank, world):
    """--impl reference: the CPU learner path on the host cores, rank 0 only."""
    if rank != 0:
        return
    steps, warmup = args.steps, max(args.warmup, 3)
    # bounded sample: keep the whole run within ~2 minutes of CPU time whatever K the driver passes
    threads = best_cpu_threads(args.workload)
    probe_v, _, _ = time_cpu(args.workload, 10, 3, threads)
    steps = max(10, min(steps, int(120 * probe_v)))
    warmup = min(warmup, max(3, int(10 * probe_v)))
    v, dt, cores = time_cpu(args.workload, steps, warmup, threads)
    line = {
        "impl": "reference", "metric": "SAC gradient steps/sec", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "replay": "20000-transition host buffer, uniform sampling w/o replacement",
                   "learners": 1},
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} update() calls (sample + update_SAC) after {warmup} warm-up, torch {torch.__version__} CPU, "
                                   f"{os.cpu_count()} logical CPUs, {cpu_model()}"},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------------
def write_cfg(workload, tmp):
    """cfg JSON in the reference's own format (cfg/*.json) for the drop-in Learner."""
    base = {"device": "cuda", "buffer_size": 1e6, "reward_scale": 1, "gamma": 0.99, "log_alpha": 0, "tau": 0.005,
            "start_memory_len": 5000, "random_step": 5000}
    if workload == "LL":
        cfg = dict(base, num_tasks=10, batch_size=256, lr_actor=3e-4, lr_critic=3e-4, num_learn=1, num_time_step=1)
    elif workload == "VS":
        cfg = dict(base, batch_size=1024, lr_actor=3e-4, lr_critic=3e-4, update_delay=5, print_period_player=2,
                   print_period_learner=5, actor_hidden_dim=[400] * 3, critic_hidden_dim=[400] * 3)
    elif workload in ("C10", "C10O"):
        names = [f"task-{i}" for i in range(10)]
        g = torch.Generator().manual_seed(7)
        emb = {n: (torch.randn(768, generator=g) * 0.3).tolist() for n in names}     # synthetic stand-in for the RoBERTa rows
        with open(os.path.join(tmp, "emb.json"), "w") as f:
            json.dump(emb, f)
        with open(os.path.join(tmp, "names.json"), "w") as f:
            json.dump(names, f)
        cfg = dict(base, use_modified_care=(workload == "C10"), num_tasks=10, batch_size=1280, update_delay=6, print_period_player=2,
                   print_period_learner=10, max_episode_time=500,
                   actor={"state_dim": 39, "action_dim": 4, "action_bound": [-1.0, 1.0], "lr_actor": 3e-4,
                          "actor_hidden_dim": [400] * 3},
                   critic={"state_dim": 39, "action_dim": 4, "lr_critic": 3e-4, "critic_hidden_dim": [400] * 3},
                   encoder={"state_dim": 39, "pretrained_embedding_json_path": os.path.join(tmp, "emb.json"),
                            "task_name_json_path": os.path.join(tmp, "names.json"), "hidden_dims_contextEnc": [50, 50],
                            "embedding_dim_contextEnc": 50, "output_dim_contextEnc": 50, "RoBERTa_embedding_dim": 768,
                            "lr_contextEnc": 3e-4, "hidden_dims_mixtureEnc": [50], "output_dim_mixtureEnc": 50,
                            "num_encoders": 6, "num_tasks": 10, "state_encoder_tau": 0.05})
    else:
        cfg = dict(base, use_weighted_loss=True, num_tasks=10, batch_size=1280, update_delay=6, print_period_player=2,
                   print_period_learner=10, max_episode_time=500,
                   actor={"state_dim": 39, "action_dim": 4, "action_bound": [-1.0, 1.0], "lr_actor": 3e-4,
                          "actor_hidden_dim": [400] * 3},
                   critic={"state_dim": 39, "action_dim": 4, "lr_critic": 3e-4, "critic_hidden_dim": [400] * 3})
    path = os.path.join(tmp, "cfg.json")
    with open(path, "w") as f:
        json.dump(cfg, f)
    return path


class _NullServer:
    """The Redis server is IPC, not part of the measured path; the bench feeds the ring directly."""

    def scan_iter(self): return []
    def delete(self, k): pass
    def set(self, k, v): pass
    def get(self, k): return None
    def rpush(self, k, v): pass

    def pipeline(self):
        class _P:
            def lrange(self, *a): return self
            def ltrim(self, *a): return self
            def execute(self): return [[], True]
        return _P()


def make_learner(workload, cfg_path, device_index, buffer_size, precision=1):
    from distributed_sac_b200 import learner as L
    srv = _NullServer()
    if workload == "LL":
        return L.LunarLanderLearner(cfg_path, write_mode=False, server=srv, device_index=device_index, buffer_size=buffer_size,
                                    precision=precision)
    if workload == "VS":
        return L.VSACLearner(cfg_path, write_mode=False, server=srv, device_index=device_index, precision=precision)
    if workload in ("C10", "C10O"):
        return L.CARELearner(None, None, cfg_path, write_mode=False, server=srv, device_index=device_index, precision=precision)
    return L.MTSACLearner(None, None, cfg_path, write_mode=False, server=srv, device_index=device_index, precision=precision)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="LL", choices=["LL", "VS", "MS", "C10", "C10O"])
    ap.add_argument("--replicas", type=int, default=1, help="independent learners co-scheduled per GPU")
    ap.add_argument("--ring", type=int, default=1 << 20, help="transitions in the device replay ring (per learner)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed update() calls for the e2e leg (default: min(steps, 2000))")
    ap.add_argument("--cpu-steps", type=int, default=300)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--precision", type=int, default=-1,
                    help="0 = fp32 FFMA GEMMs, 1 = 3xTF32 tcgen05 GEMMs (fp32-class accuracy), -1 = time both, headline = faster")
    ap.add_argument("--sweep", default="", help="comma list of extra replicas-per-GPU values to report (device-resident)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import __graft_entry__ as ge
    ge.build()
    from distributed_sac_b200 import _lib
    from distributed_sac_b200.core import Replay, SacCore

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(args.warmup, 3)
    K = args.steps
    R = args.replicas
    pk = peaks()
    stream = torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_resident(core, ring, steps, warm):
        with torch.cuda.stream(stream):
            core.step_sampled(ring, warm)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            core.step_sampled(ring, steps)
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident leg (value) -------------------------------------------------------------
    from distributed_sac_b200.replicas import broadcast_initial_params
    row_bytes = None
    by_precision = {}
    cands = [0, 1] if args.precision < 0 else [args.precision]
    probe = {}
    if len(cands) > 1:            # short probe of both GEMM back-ends; the full timed run uses the faster one
        for pr in cands:
            c0 = SacCore(core_config(args.workload, R, pr), local, seed=1234 + rank)
            r0 = Replay(c0, 1 << 16, where="device", seed=99 + rank)
            r0.fill_synthetic(1 << 16, seed=1234 + rank)
            probe[pr] = timed_resident(c0, r0, 300, 50)
            by_precision["fp32_ffma" if pr == 0 else "tc3xtf32"] = world * R * 300 / (probe[pr] * 1e-3)
            r0.close(); c0.close()
        args.precision = min(probe, key=probe.get)
    core = SacCore(core_config(args.workload, R, args.precision), local, seed=1234 + rank)
    if dist is not None:
        # independent replicas: ONE collective, the broadcast of the initial parameter arena
        broadcast_initial_params(core, src=0)
    ring = Replay(core, args.ring, where="device", seed=99 + rank)
    ring.fill_synthetic(args.ring, seed=1234 + rank)
    row_bytes = 4 * ((2 * core.cfg.obs_dim + core.cfg.act_dim + 2 + 31) // 32 * 32)
    ring_mib = args.ring * row_bytes / 2 ** 20

    clocks = ClockSampler(local)
    clocks.start()
    ms = timed_resident(core, ring, K, W)
    clk = clocks.stop()
    value = world * R * K / (ms * 1e-3)
    losses = core.read_losses(min(K, 64))
    assert torch.isfinite(losses).all(), "non-finite losses in the timed region"
    by_precision["fp32_ffma" if args.precision == 0 else "tc3xtf32"] = value

    # per-launch profile of one step (eager, CUDA events) for the roofline block
    with torch.cuda.stream(stream):
        prof = core.profile_step(ring, iters=30)
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        tl = core.graph_timeline(ring, iters=200)          # true in-graph start-to-start times (us)
    torch.cuda.synchronize()
    names = ["sample_indices", "ingest"] + [n for n, _ in prof][1:]
    prof = list(zip(names, [t * 1e-3 for t in tl]))      # ms, like the eager numbers it replaces
    tot = sum(t for _, t in prof)
    by_kernel = {}
    for name, t in prof:
        fam = name.split("(")[0] if "tcgen05" not in name else "gemm_tc"
        fam = "gemm_ffma" if fam.startswith("gemm_") and fam != "gemm_tc" else fam
        by_kernel[fam] = by_kernel.get(fam, 0.0) + t
    top = max(by_kernel.items(), key=lambda kv: kv[1])
    work = WORK[args.workload]
    step_s = ms * 1e-3 / K
    ach_gbs = R * work["mbytes"] * 1e6 / step_s / 1e9
    ach_tf = R * work["gflop"] * 1e9 / step_s / 1e12
    # DRAM traffic of one step from the committed ncu capture of this workload / back-end (cold caches under ncu: an upper
    # bound -- in the running step parameters, Adam state and activations stay in the 126 MB L2); null when not captured
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r1c_traffic.json")
    if os.path.exists(tpath) and R == 1 and (args.workload, args.precision) in (("LL", 0), ("VS", 1)):
        with open(tpath) as f:
            tj = json.load(f).get(args.workload)
        if tj:
            traffic = tj["dram_read_bytes_per_step"] + tj["dram_write_bytes_per_step"]
            traffic_src = "profiles/r1c_traffic.json: " + tj["source"]
    roofline = {
        "bound": "hbm", "achieved": ach_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": ach_gbs / pk["hbm"],
        "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per step (= per graph launch)", "traffic_source": traffic_src,
        "peak_source": pk["src"],
        "launch": f"one CUDA-graph launch = one gradient step of {R} learner(s) = {core.launches_per_step + 1} kernels",
        "algorithmic_bytes_per_step": work["mbytes"] * 1e6 * R, "algorithmic_flop_per_step": work["gflop"] * 1e9 * R,
        "tensor": {"achieved": ach_tf, "peak": pk["tf"], "unit": "TFLOP/s", "frac": ach_tf / pk["tf"],
                   "note": "3xTF32 issues 3 tensor-core MACs per algorithmic MAC; denominator = measured dense bf16 cuBLAS"},
        "dominant_kernel": {"name": top[0], "share_of_step": top[1] / tot, "us_per_step_in_graph": top[1] * 1e3,
                            "launches_per_step": sum(1 for n, _ in prof if (("tcgen05" in n) if top[0] == "gemm_tc" else
                                                     (n.startswith("gemm_") and "tcgen05" not in n) if top[0] == "gemm_ffma" else n.split("(")[0] == top[0]))},
        "per_launch_us_in_graph": [[n, round(t * 1e3, 2)] for n, t in prof],
        "note": "latency-bound: ~%d dependent launches per step; params+Adam state stay L2-resident between steps" % (core.launches_per_step + 1),
    }

    sweep = {}
    for r_extra in [int(x) for x in args.sweep.split(",") if x]:
        sweep[str(r_extra)] = {}
        n_ring = max(1 << 16, args.ring // max(1, r_extra // 2))
        for pr in (0, 1):
            c2 = SacCore(core_config(args.workload, r_extra, pr), local, seed=77 + rank)
            ring2 = Replay(c2, n_ring, where="device", seed=5)
            ring2.fill_synthetic(n_ring, seed=6)
            ms2 = timed_resident(c2, ring2, max(200, K // 4), W)
            sweep[str(r_extra)]["fp32_ffma" if pr == 0 else "tc3xtf32"] = world * r_extra * max(200, K // 4) / (ms2 * 1e-3)
            ring2.close(); c2.close()
    ring.close()
    core.close()

    # ---- e2e leg: reference-shaped Learner.update() with a pinned-host replay ring ------------------
    Ke = args.e2e_steps or min(K, 2000)
    with tempfile.TemporaryDirectory() as tmp:
        old = os.getcwd()
        os.chdir(tmp)
        try:
            lrn = make_learner(args.workload, write_cfg(args.workload, tmp), local, buffer_size=200000, precision=args.precision)
            host_ring_n = 200000
            lrn.memory.ring.fill_synthetic(host_ring_n, seed=4321 + rank)
            with torch.cuda.stream(stream):
                lrn.soft_update(None, None, 1.0)
                for _ in range(W):
                    lrn.update()
                barrier()
                t0 = time.perf_counter()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                last = None
                for _ in range(Ke):
                    last = lrn.update()
                e1.record()
                barrier()
                wall = time.perf_counter() - t0
            assert all(map(lambda x: x == x, last)), "NaN loss in e2e leg"
            t_e2e = max(wall, e0.elapsed_time(e1) * 1e-3)
            if dist is not None:
                t = torch.tensor([t_e2e], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                t_e2e = float(t.item())
            # pipelined variant: same per-step H2D + zero-copy D2H of the losses, no per-step host sync
            with torch.cuda.stream(stream):
                lrn.update_many(W)
                barrier()
                t0 = time.perf_counter()
                lrn.update_many(Ke)
                barrier()
                t_pipe = time.perf_counter() - t0
            # run()-loop body: update + publication of the actor blob (LL/learner.py:296-299), two ways:
            #   blocking  = update(); get_parameters()              (what the reference does, one after the other)
            #   overlapped = enqueue step; publish_begin(); read losses; publish_wait()   (what Learner.run() here does)
            import pickle as _pk
            Kp = max(50, min(Ke, 500))
            pub_bytes = 0
            with torch.cuda.stream(stream):
                t0 = time.perf_counter()
                for _ in range(Kp):
                    lrn.update()
                    blob = _pk.dumps(lrn.get_parameters())
                t_pub_block = time.perf_counter() - t0
                t0 = time.perf_counter()
                for _ in range(Kp):
                    lrn.memory.enqueue_step(lrn.core)
                    lrn.publish_begin()
                    lrn.core.read_losses(1)
                    blob = _pk.dumps(lrn.publish_wait())
                t_pub_over = time.perf_counter() - t0
                pub_bytes = 4 * sum(v.numel() for m in _pk.loads(blob).values() for v in m.values())
            h2d = lrn.core.cfg.batch * row_bytes
            d2h = 16
            lrn.memory.stop()
        finally:
            os.chdir(old)
    e2e = {"value": world * Ke / t_e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "Learner.update() per step (host sample+gather -> pinned staging -> cudaMemcpyAsync on a side stream -> "
                  "graph launch -> losses written to mapped pinned memory -> stream sync)",
           "steps": Ke, "pipelined_update_many": {"value": world * Ke / t_pipe, "unit": "steps/s"},
           "with_publication": {"unit": "steps/s", "d2h_bytes_per_step": pub_bytes + d2h, "steps": Kp,
                                "update_then_get_parameters": world * Kp / t_pub_block,
                                "overlapped_run_loop": world * Kp / t_pub_over,
                                "note": "per step: update + {'actor': state_dict} snapshot -> pinned host -> pickle.dumps, "
                                        "as Learner.run() publishes it for the players"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, cores = time_cpu(args.workload, args.cpu_steps, 10)
        cpu = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} update() calls (sample + update_SAC) of the CPU oracle port after 10 warm-up, "
                         f"{dt:.1f} s, {os.cpu_count()} logical CPUs, {cpu_model()}"}

    if rank == 0:
        line = {
            "metric": "SAC gradient steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "learners_per_gpu": R,
                       "replay": f"{args.ring}-transition device-resident ring per learner, {row_bytes} B rows = {ring_mib:.0f} MiB",
                       "l2": "inputs larger than L2: minibatches are gathered from the %.0f MiB ring (> 126 MB L2); no explicit flush" % (ring_mib * R),
                       "parallelism": f"{world} x independent learner replicas, NCCL broadcast of initial weights only",
                       "precision": ("hidden-layer GEMMs 3xTF32 on tcgen05 (fp32-class, <=2e-6 of fp64), rest fp32 FFMA" if args.precision == 1
                                     else "fp32 FFMA GEMMs") + ", fp64-evaluated transcendentals"},
            "clocks": clk, "e2e": e2e, "gpu_launches": K * (core.launches_per_step + 1),
            "roofline": roofline, "cpu_baseline": cpu,
        }
        line["by_precision"] = by_precision
        line["config"]["gemm_backend"] = "fp32 FFMA" if args.precision == 0 else "tcgen05 3xTF32"
        if sweep:
            line["replicas_per_gpu_sweep"] = sweep
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
