#!/usr/bin/env python
"""Benchmark of the shifu post-physics hot path (BASELINE.json metric: env-steps/s + % HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port)

A "step" = one control step of a1_conditional on resident simulator state:
PD torque x4 (+ action scale/clip), fused post-physics (body-frame carry, 187-point height scan,
termination, 6 reward terms + episode sums, reset with Philox draws, obs, history push, clips),
reset-id compaction, statistics collect (+ all-reduce over ranks when N>1) and extras publish.
Prints ONE JSON line on rank 0.  See DESIGN.md §Measurement for the definitions.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (obs+reward+term+reset)"
UNIT = "env-steps/s"
B_ALG_POST = 1819          # algorithmic HBM bytes / env-step of the fused post-physics kernel
B_ALG_PD = 192             # per PD substep: 96 dof_state + 48 action read, 48 torque write
ENVS_PER_GPU = 1 << 20     # BASELINE configs[2]/[4]: 1M envs per B200 (weak scaling -> 8M on 8 GPUs)
SWEEP = (4096, 16384, 65536, 262144)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# workload construction
# ---------------------------------------------------------------------------------------------

def _terrain():
    import numpy as np
    from shifu_b200.sim import fake_isaacgym
    fake_isaacgym.install("cpu")
    from shifu_b200.configs import TerrainEnvConfig
    from shifu_b200.utils.heightmap import Terrain
    np.random.seed(0)
    cfg = TerrainEnvConfig()
    return Terrain(cfg.terrain, 1), cfg


def build_a1(n_local, rank, world, device, terrain=None, seed=1234, carry=True, want_heights=False):
    """Resident synthetic state for n_local envs on `device` (global ids rank*n_local ...)."""
    import numpy as np
    import torch
    from shifu_b200 import hotpath
    from shifu_b200.sim.synthetic import a1_snapshot
    ter, cfg = terrain or _terrain()
    n_global = n_local * world
    gid = torch.arange(rank * n_local, (rank + 1) * n_local)
    # isaac_gym.py:342-344 on the GLOBAL env index, so results do not depend on the GPU count (§8e)
    types = torch.div(gid, (n_global / cfg.terrain.num_cols), rounding_mode='floor').to(torch.long)
    types.clamp_(max=cfg.terrain.num_cols - 1)
    g = torch.Generator().manual_seed(seed + rank)
    levels0 = torch.randint(0, cfg.terrain.max_init_terrain_level + 1, (n_local,), generator=g)
    origins = torch.from_numpy(ter.env_origins).float()
    env_origins = origins[levels0, types].to(device).contiguous()
    snap = a1_snapshot(seed + rank, 1, n_local, gen_device=device, p_base=0.01, p_leg=0.1, xy_range=3.0,
                       offmap=False)
    root = snap.root_offset.clone()
    root[:, :3] += env_origins
    dof = snap.dof[4].reshape(n_local * 12, 2).contiguous()
    contact = snap.contact.reshape(n_local * 17, 3).contiguous()
    desc = hotpath.a1_desc(n_local, env_offset=rank * n_local, rng_seed=seed)
    hp = hotpath.A1HotPath(desc, root_state=root.contiguous(), dof_state=dof, contact_state=contact,
                           height_samples=torch.from_numpy(ter.heightsamples), terrain_origins=origins,
                           terrain_types=types, env_origins=env_origins, carry_body_frame=carry,
                           want_measured_heights=want_heights)
    hp.ep_len.copy_(torch.randint(0, 500, (n_local,), generator=g).to(device))
    hp.command.copy_((torch.rand(n_local, 3, generator=g) * 2 - 1).to(device))
    hp.terrain_levels.copy_(levels0.to(device))
    hp.sync_level_sum()
    hp.body_frame()             # seed the carried body-frame velocities from the initial root rows
    raw_actions = snap.actions.contiguous()
    return hp, raw_actions


LAUNCHES_PER_STEP = 4 + 1 + 1 + 1 + 1      # pd x4, fused post-physics, compaction, collect, publish


def time_resident(hp, raw_actions, steps, warmup, allreduce=None, barrier=None, flush=None):
    """K steps on resident state.  Returns (total_ms, per-launch ms of the fused kernel)."""
    import torch
    for _ in range(warmup):
        hp.step_resident(raw_actions, allreduce=allreduce)
    torch.cuda.synchronize()
    if barrier is not None:
        barrier()
    torch.cuda.synchronize()
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    k1 = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total = 0.0
    if flush is None:
        t0.record()
    for i in range(steps):
        if flush is not None:
            flush.zero_()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
        hp.pd_torque(raw_actions)
        hp.pd_torque(); hp.pd_torque(); hp.pd_torque()
        k0[i].record()
        hp.post_physics()
        k1[i].record()
        hp.finalize(allreduce)
        if flush is not None:
            t1.record()
            t1.synchronize()
            total += t0.elapsed_time(t1)
    if flush is None:
        t1.record()
    torch.cuda.synchronize()
    if barrier is not None:
        barrier()
    if flush is None:
        total = t0.elapsed_time(t1)
    kern = [a.elapsed_time(b) for a, b in zip(k0, k1)]
    return total, kern


def time_graph(hp, raw_actions, steps, warmup):
    """Same step captured once into a CUDA graph (device-resident step counter) — the launch-bound
    small-N configurations."""
    import torch
    hp.step_dev.fill_(hp.step_counter + 1)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            hp.step_resident(raw_actions, use_step_dev=True)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        hp.step_resident(raw_actions, use_step_dev=True)
    for _ in range(warmup):
        g.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        g.replay()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1)


def time_abb(n, device, flush, steps=50, warmup=10):
    """BASELINE configs[3]: abb_pushbox_vision prior-stage obs/reward/reset, 65 536 envs.  ~90 B/env:
    launch-latency-bound, so env-steps/s is reported without a roofline fraction."""
    import torch
    from shifu_b200 import hotpath
    from shifu_b200.sim.synthetic import abb_snapshot
    snap = abb_snapshot(7, 1, n, gen_device=device)
    hp = hotpath.AbbHotPath(hotpath.abb_desc(n), root_state=snap.root.reshape(n * 4, 13).contiguous(),
                            body_state=snap.body.reshape(n * 10, 13).contiguous(),
                            dof_state=snap.dof.reshape(n * 6, 2).contiguous())
    hp.ep_len.copy_(torch.randint(0, 200, (n,), device=device))
    cube0 = hp.root_state.view(n, 4, 13)[:, 2].clone()
    for _ in range(warmup):
        hp.step_resident()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(steps):
        hp.root_state.view(n, 4, 13)[:, 2].copy_(cube0)      # undo the previous step's cube resets (untimed)
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        hp.step_resident()
        t1.record()
        t1.synchronize()
        total += t0.elapsed_time(t1)
    ms = total / steps
    # row N2 (pre-physics arm action path): action -> ee goal -> damped least-squares IK, one launch
    k = hotpath.EnvKernels(device, n)
    jac = (torch.rand(n, 9, 6, 6, device=device) * 2 - 1) * 0.6
    actions = torch.rand(n, 3, device=device) * 2 - 1
    ik = lambda: k.arm_ik(body_state=hp.body_state, num_bodies=10, ee_body=6, jacobian=jac, ee_link=5,
                          dof_state=hp.dof_state, num_dof=6, dof_targets=hp.dof_targets, actions=actions,
                          ee_velocity=0.2, dt=0.1, min_ee_pos=(-0.2, -0.2, 0.11), max_ee_pos=(0.2, 0.2, 0.14),
                          tar_quat=(0., 1., 0., 0.))
    for _ in range(warmup):
        ik()
    torch.cuda.synchronize()
    ik_total = 0.0
    for _ in range(steps):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        ik()
        t1.record()
        t1.synchronize()
        ik_total += t0.elapsed_time(t1)
    ik_ms = ik_total / steps
    ik_bytes = 28 + 144 + 24 + 12 + 24             # ee pose, 6x6 jacobian, dof_pos, action, dof_targets
    return {"envs": n, "ms_per_step_l2_flushed": ms, "env_steps_per_s": n / (ms * 1e-3),
            "reset_fraction": float(hp.n_reset.item()) / n, "launches_per_step": 4,
            "note": "post-physics + compaction + stats; launch-latency-bound (about 90 B/env)",
            "arm_ik_ms_l2_flushed": ik_ms, "arm_ik_algorithmic_gbs": ik_bytes * n / (ik_ms * 1e-3) / 1e9,
            "arm_ik_note": "pre-physics action path (SURVEY 8f N2): goal + clamp + 6x6 damped least squares, "
                           "232 B/env algorithmic; launch-latency-bound at this size"}


def time_camera(n, device, flush, peak, steps=20, warmup=5, h=128, w=128):
    """CameraSensor.refresh_image_tensors for the reference's perceptual-stage
    camera (task_config.py:124-145: 128x128, colour normalised + depth + segmentation) as one launch."""
    import torch
    from shifu_b200 import hotpath
    k = hotpath.EnvKernels(device, n)
    color = torch.randint(0, 256, (n, h, w, 4), dtype=torch.uint8, device=device)
    depth = -torch.rand(n, h, w, device=device)
    seg = torch.randint(0, 5, (n, h, w), dtype=torch.int32, device=device)
    table = lambda t: (t.data_ptr() + torch.arange(n, dtype=torch.int64) * t[0].numel() * t.element_size()).to(device)
    args = dict(height=h, width=w, normalize_color=True,
                color=(table(color), torch.empty(n, h, w, 3, device=device)),
                depth=(table(depth), torch.empty(n, h, w, device=device)),
                seg=(table(seg), torch.empty(n, h, w, dtype=torch.int32, device=device)))
    for _ in range(warmup):
        k.camera_gather(**args)
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(steps):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        k.camera_gather(**args)
        t1.record()
        t1.synchronize()
        total += t0.elapsed_time(t1)
    ms = total / steps
    nbytes = n * h * w * (4 + 12 + 8 + 8)
    return {"envs": n, "image": f"{h}x{w} rgba+depth+seg, colour normalised", "ms_l2_flushed": ms,
            "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
            "envs_per_s": n / (ms * 1e-3),
            "note": "one shifu_camera_gather launch replaces the reference's Python loop over the envs "
                    "(sensors.py:165-188); 32 B/pixel"}


def time_e2e(hp, raw_actions, steps, warmup):
    """Same step through the public host API with HOST buffers: every step copies the simulator
    state + actions from pinned host memory and reads obs / reward / reset flags back."""
    import torch
    n = hp.n
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t.cpu())
    h_root, h_dof, h_contact, h_act = pin(hp.root_state), pin(hp.dof_state), pin(hp.contact_state), pin(raw_actions)
    d_act = torch.empty_like(raw_actions)
    h_obs = torch.empty(hp.obs_buf.shape, dtype=torch.float, pin_memory=True)
    h_rew = torch.empty(n, dtype=torch.float, pin_memory=True)
    h_reset = torch.empty(n, dtype=torch.bool, pin_memory=True)
    h2d = sum(t.numel() * t.element_size() for t in (h_root, h_dof, h_contact, h_act))
    d2h = sum(t.numel() * t.element_size() for t in (h_obs, h_rew, h_reset))

    # Three streams so that the upload of step t+1 (host->device) overlaps the read-back of step t
    # (device->host) — PCIe is full duplex; events keep every buffer single-owner:
    #   in:      wait compute(t-1) done -> H2D state/actions of step t
    #   compute: wait in(t), wait out(t-1) done -> the step
    #   out:     wait compute(t) -> D2H obs / rew / reset
    s_in, s_cmp, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    ev_in, ev_cmp, ev_out = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    for s in (s_in, s_cmp, s_out):
        s.wait_stream(torch.cuda.current_stream())
    ev_cmp.record(s_cmp)
    ev_out.record(s_out)

    def one():
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_cmp)
            hp.root_state.copy_(h_root, non_blocking=True)
            hp.dof_state.copy_(h_dof, non_blocking=True)
            hp.contact_state.copy_(h_contact, non_blocking=True)
            d_act.copy_(h_act, non_blocking=True)
            ev_in.record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(ev_in)
            s_cmp.wait_event(ev_out)
            hp.step_resident(d_act)
            ev_cmp.record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_cmp)
            h_obs.copy_(hp.obs_buf, non_blocking=True)
            h_rew.copy_(hp.rew_buf, non_blocking=True)
            h_reset.copy_(hp.reset_buf, non_blocking=True)
            ev_out.record(s_out)

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.current_stream().wait_stream(s_cmp)
    return dt * 1e3, h2d, d2h, float(h_rew.mean())


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's torch-CPU path
# ---------------------------------------------------------------------------------------------

def cpu_oracle_rate(n, steps, warmup=1, seed=1234):
    """env-steps/s of oracle/shifu_oracle.py (torch CPU, all host threads) on the same synthetic
    workload, n envs.  bench.py's only use of oracle/: as the reported CPU baseline."""
    import numpy as np
    import torch
    from oracle import shifu_oracle as so
    from shifu_b200.sim.synthetic import a1_snapshot
    torch.set_num_threads(os.cpu_count() or 1)
    ter, cfg = _terrain()
    types = torch.div(torch.arange(n), (n / cfg.terrain.num_cols), rounding_mode='floor').to(torch.long)
    g = torch.Generator().manual_seed(seed)
    levels0 = torch.randint(0, cfg.terrain.max_init_terrain_level + 1, (n,), generator=g)
    origins = torch.from_numpy(ter.env_origins).float()
    p = so.A1Params(n=n, rng_seed=seed)
    st = so.a1_new_state(p, torch.from_numpy(ter.heightsamples), origins, types, origins[levels0, types])
    st.ep_len[:] = torch.randint(0, 500, (n,), generator=g)
    st.command[:] = torch.rand(n, 3, generator=g) * 2 - 1
    hpts = p.height_points()
    snap = a1_snapshot(seed, 1, n, p_base=0.01, p_leg=0.1, xy_range=3.0, offmap=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        so.a1_step(p, st, snap.actions, snap, hpts)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return n / statistics.median(times), statistics.median(times) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 65536
    rate, ms = cpu_oracle_rate(n, steps=max(args.steps, 1), warmup=min(args.warmup, 2))
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "a1_conditional post-physics step incl. 187-pt heightfield scan + PD torques, "
                               f"{ENVS_PER_GPU} envs/GPU", "sample_envs": n},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle/shifu_oracle.py a1_step (torch CPU restatement of the reference, "
                                   f"bit-identical to it), {n} envs x {args.steps} steps, median"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback; "
                         "use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    allreduce = barrier = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(device))
        allreduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)
        barrier = dist.barrier
    n = args.envs_per_gpu
    terrain = _terrain()
    hp, raw = build_a1(n, rank, world, device, terrain)
    peak, peak_src = _peaks()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, kern_ms = time_resident(hp, raw, args.steps, args.warmup, allreduce, barrier)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], device=device, dtype=torch.double)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3)
    kms = statistics.mean(kern_ms)
    achieved = B_ALG_POST * n / (kms * 1e-3) / 1e9
    reset_frac = float(hp.n_reset.item()) / n

    line = None
    if rank == 0:
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("a1_post_physics_kernel_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "a1_conditional post-physics step incl. 187-pt heightfield scan + PD torques, "
                                   f"{n} envs/GPU (BASELINE configs[2]; configs[4] when n_gpus=8)",
                       "envs_per_gpu": n, "global_envs": n * world, "decimation": 4,
                       "l2": "inputs larger than L2 (state + obs = %.0f MB per GPU vs 126 MB L2); no flush"
                             % ((B_ALG_POST + 4 * B_ALG_PD) * n / 1e6),
                       "reset_fraction_per_step": reset_frac, "launch": "direct (8 launches/step)",
                       "collective": "all_reduce(16 x f64) per step" if world > 1 else "none"},
            "roofline": {"bound": "hbm", "kernel": "a1_post_physics_tma_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_env": B_ALG_POST,
                         "kernel_ms": kms, "kernel_share_of_step": kms / ms_per_step,
                         "step_frac": (B_ALG_POST + 4 * B_ALG_PD) * n / (ms_per_step * 1e-3) / 1e9 / peak},
            "clocks": clocks, "gpu_launches": LAUNCHES_PER_STEP * args.steps,
        }
    # ---- N=1 extras: e2e through host buffers, the size sweep, the CPU baseline -----------------
    if world == 1 and args.quick:
        pass
    elif world == 1:
        e_ms, h2d, d2h, _ = time_e2e(hp, raw, min(40, max(3, args.steps // 4)), 2)
        e_steps = min(40, max(3, args.steps // 4))
        line["e2e"] = {"value": n / (e_ms / e_steps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "ms_per_step": e_ms / e_steps,
                       "note": "pinned host state+actions -> device, step, obs+rew+reset -> pinned host; upload of step t+1 overlaps read-back of step t (3 streams); PCIe-bound"}
        del hp
        torch.cuda.empty_cache()
        sweep = []
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        for m in SWEEP:
            h2, r2 = build_a1(m, 0, 1, device, terrain)
            tot, km = time_resident(h2, r2, 20, 5, flush=flush)
            gms = time_graph(h2, r2, 200, 20) / 200
            sweep.append({"envs": m, "ms_per_step_l2_flushed": tot / 20, "env_steps_per_s_l2_flushed": m / (tot / 20 * 1e-3),
                          "ms_per_step_graph_l2_warm": gms, "env_steps_per_s_graph_l2_warm": m / (gms * 1e-3),
                          "fused_kernel_ms": statistics.mean(km)})
            del h2
        line["sweep"] = sweep
        line["abb_prior_stage"] = time_abb(65536, device, flush)
        line["camera_gather"] = time_camera(2048, device, flush, peak)
        if not args.no_cpu:
            rate, ms = cpu_oracle_rate(65536, steps=5, warmup=1)
            rate4k, ms4k = cpu_oracle_rate(4096, steps=20, warmup=3)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "oracle a1_step (torch CPU, all cores), 65536 envs x 5 steps, median "
                                              f"{ms:.0f} ms/step; 4096 envs (BASELINE configs[0]): {rate4k:.3g} env-steps/s"}
    else:
        line_e2e = None
        e_steps = min(40, max(3, args.steps // 4))
        e_ms, h2d, d2h, _ = time_e2e(hp, raw, e_steps, 2)
        t = torch.tensor([e_ms], device=device, dtype=torch.double)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            line["e2e"] = {"value": n * world / (float(t.item()) / e_steps * 1e-3), "unit": UNIT,
                           "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world}
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries that print there (e.g. NCCL's version banner)
    are sent to stderr instead."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--quick", action="store_true", help="timed region only (for runs under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
