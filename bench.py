#!/usr/bin/env python
"""Benchmark of the batched drone-env step (BASELINE.json: agent-steps/s at 1/2/4/8 B200 vs the CPU env).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl reference]

One "step" = env.step + ss_reward for every env of the shard, followed by the auto-reset of the episodes that
finish: ONE call through the C ABI (aac_step_autoreset) = two launches of the env kernel (step, reset).  Prints ONE JSON line.
Workloads: c2 = one_model_att 4096 envs x 3 drones x 36 rays; c3 (default, the
configuration the 1/2/4/8-GPU metric and the north-star target are quoted on) = tdCPA_forV2 65536 envs x
10 drones x 36 rays per GPU; c5 = 131072 envs x 20 drones x 72 rays per GPU (the 8-GPU 1M-env sweep).
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
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, envs per GPU, drones, rays, description)
    "c1": ("att", 1, 3, 18, "one_model_att reference scenario: 1 env x 3 drones, 18-ray radar, seeded random actions, 1000 steps on ONE CPU core (BASELINE config 1)"),
    "c2": ("att", 4096, 3, 36, "one_model_att 4096 envs x 3 drones, 36-ray radar, single grid map"),
    "c2r18": ("att", 4096, 3, 18, "one_model_att 4096 envs x 3 drones, 18-ray radar, single grid map"),
    "c3": ("tdcpa_v2", 65536, 10, 36, "tdCPA_forV2 65536 envs x 10 drones per GPU, 36-ray radar, single grid map"),
    "c4": ("multimap", 65536, 3, 18, "radar_multipleMap 65536 envs x 3 drones per GPU, 18-ray radar, 14 heterogeneous maps, map drawn per episode"),
    "c5": ("tdcpa_v2", 131072, 20, 72, "tdCPA_forV2 131072 envs x 20 drones per GPU, 72-ray radar (1M envs on 8 GPUs)"),
}
W_REF = 4  # reference-line vertices assumed by SURVEY.md section 8d's byte count


def algorithmic_bytes(variant, n, r):
    """4 * (26 + 2W + obs words) per agent-step, W = 4."""
    obs = {"att": 6 + 4 * (n - 1) + r + 6 * (n - 1), "v2": 7 + 5 * (n - 1) + r, "mm": 6 + r}[variant]
    return 4 * (26 + 2 * W_REF + obs)


# committed `ncu --set full` summaries of the launches of one C3 step: the phased launch (step loop + reset loop in one
# kernel, what aac_step_autoreset runs for C3) or the step launch and the reset launch (--launches 2)
PROFILED = {1: ("r2_env_kernel_v2_phased_ncu_full_summary.csv",),
            2: ("r2_env_kernel_v2_step_ncu_full_summary.csv", "r2_env_kernel_v2_reset_ncu_full_summary.csv")}


def source_hash():
    """sha256 over the kernel sources: a committed ncu capture belongs to the build that is being timed only if it
    carries the same hash (tests/tools/ncu_summary.py writes it into the summary's `Source Hash` row)."""
    import hashlib
    h = hashlib.sha256()
    for f in ("aac_kernels.cu", "aac_kernels.cuh", "aac_radar.cuh", "aac_plan.cuh", "aac_capi.cu"):
        h.update(open(os.path.join(ROOT, "multi_agent_aac_b200", "csrc", f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "aac_env.h"), "rb").read())
    return h.hexdigest()[:16]


def _profiled_rows(kernel_names):
    """The committed `ncu --set full` captures of the two launches of one step (profiles/), or None unless each names
    exactly the kernel instantiation this run launches AND was taken from this very source."""
    import csv
    out = []
    for name, want in zip(PROFILED[len(kernel_names)], kernel_names):
        if not os.path.exists(os.path.join(ROOT, "profiles", name)):
            return None
        rows = {r[0]: (r[1], r[2]) for r in csv.reader(open(os.path.join(ROOT, "profiles", name))) if len(r) == 3}
        got = rows.get("Kernel Name", ("", ""))[1].replace("aac::", "").replace("(int)", "").replace("(bool)", "").replace(" ", "")
        if want.replace(" ", "") not in got or rows.get("Source Hash", ("", ""))[1] != source_hash():
            return None
        out.append(rows)
    return out


def profiled_traffic(kernel_names):
    """DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum, summed over the step launch and the reset
    launch) from the committed captures, or None."""
    try:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for rows in _profiled_rows(kernel_names):
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                unit, val = rows[k]
                tot += float(val) * scale[unit]
        return tot
    except Exception:
        return None


def profiled_metric(name, kernel_names):
    """One counter of the committed captures, summed over the two launches of a step, or None."""
    try:
        return sum(float(rows[name][1]) for rows in _profiled_rows(kernel_names))
    except Exception:
        return None


def launched_kernels(variant, n, r, radar_mode, launches_per_step):
    """The instantiations one aac_step_autoreset call launches for the specialised tdCPA_forV2 shapes (aac_kernels.cu
    launch_aux): env_kernel<VAR, AUX, LEAN, N, R, EVS, MT, RM, CS, PL>, MT = 4 phased (one launch: step loop, then reset loop) or
    3 step-only + 2 reset-only (two launches), RM = radar mode, CS = 0 (no later-fork sensors), PL = 0 (no per-episode search)."""
    if variant != "v2" or (n, r) not in ((10, 36), (20, 72)):
        return None
    if launches_per_step == 1:
        return ["env_kernel<1,0,1,%d,%d,0,4,%d,0,0>" % (n, r, radar_mode)]
    if launches_per_step == 2:
        return ["env_kernel<1,0,1,%d,%d,0,3,%d,0,0>" % (n, r, radar_mode), "env_kernel<1,0,1,%d,%d,0,2,%d,0,0>" % (n, r, radar_mode)]
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons sampled through NVML every 4 ms while the timed region runs (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; an `nvidia-smi -lms 25` stream is the fallback when
    the NVML binding cannot be loaded)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows, self.proc, self.nv, self.h, self.stop_flag = [], None, None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def run(self):
        if self.nv:
            nv = self.nv
            bits = [getattr(nv, n, 0) for n in ("nvmlClocksEventReasonHwSlowdown", "nvmlClocksEventReasonHwThermalSlowdown",
                                                "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksEventReasonSwPowerCap")]
            fallback = [getattr(nv, n, 0) for n in ("nvmlClocksThrottleReasonHwSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown",
                                                    "nvmlClocksThrottleReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwPowerCap")]
            bits = [b or f for b, f in zip(bits, fallback)]
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
            while not self.stop_flag:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    rs = int(reasons_fn(self.h))
                    self.rows.append((time.perf_counter(), [str(sm), str(self.max_sm), "%.2f" % pw] + ["Active" if (b and rs & b) else "Not Active" for b in bits]))
                except Exception:
                    pass
                time.sleep(0.004)
            return
        if not self.proc:
            return
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def summary(self, t0, t1):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=3)
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if r and num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if len(r) > 1 and num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if len(r) > 2 and num(r[2]) is not None]
        names = self.NAMES
        reasons = sorted({names[k] for r in rows for k in range(4) if len(r) > 3 + k and r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows),
                "source": "nvml" if self.nv else "nvidia-smi"}


def variant_of(preset_name):
    return {"att": "att", "tdcpa_v2": "v2", "multimap": "mm"}[preset_name]


def build_world(wl, n_scen, seed):
    from multi_agent_aac_b200.maps import multimap_set, synthetic_map
    from multi_agent_aac_b200.reset import MultiMapBank, ScenarioBank
    preset_name, envs, n, r, _ = WORKLOADS[wl]
    if preset_name == "multimap":
        maps = multimap_set(seed=0)
        return maps, MultiMapBank(maps, n, n_scen, w_max=32, seed=seed)
    gmap = synthetic_map(seed=0)
    bank = ScenarioBank(gmap, n, n_scen, w_max=32, seed=seed)
    return gmap, bank


def cpu_reference_run(wl, steps, warmup, sample_envs, threads=None):
    """The float64 oracle port (oracle/aac_oracle.c, OpenMP over envs) on the host cores."""
    from oracle.oracle import OracleEnv, RADAR_LAST_HIT, RADAR_MIN
    from multi_agent_aac_b200.reset import Episode
    preset_name, _, n, r, desc = WORKLOADS[wl]
    variant = variant_of(preset_name)
    cores = threads or os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    gmap, bank = build_world(wl, 64, seed=123)
    E = sample_envs
    orc = OracleEnv(variant, gmap, E, n, r, w_max=32, radar_mode=RADAR_LAST_HIT if variant == "v2" else RADAR_MIN)
    for e in range(E):
        s = e % bank.n_scenarios
        mid = int(bank.map_id[s]) if hasattr(bank, "map_id") else 0
        gm = gmap[mid] if isinstance(gmap, list) else gmap
        g = gm.grid_length
        lines = []
        for i in range(n):
            w = int(bank.w[s, i])
            c = bank.cells[s, i, :w].astype(np.int64)
            lines.append(np.stack([gm.x0c + (c >> 8) * g, gm.y0c + (c & 255) * g], -1).astype(np.float64))
        heads = [float(np.arctan2(l[1][1] - l[0][1], l[1][0] - l[0][0])) for l in lines]
        orc.set_episode(e, [l[0] for l in lines], lines, heads, map_id=mid)
    orc.observe()
    rng = np.random.default_rng(0)
    acts = rng.uniform(-1, 1, size=(4, E, n, 2))
    for k in range(warmup):
        orc.step(acts[k % 4])
    t0 = time.perf_counter()
    for k in range(steps):
        orc.step(acts[k % 4])
    dt = time.perf_counter() - t0
    return {"value": E * n * steps / dt, "unit": "agent-steps/s", "cores": cores, "kind": "port",
            "sample": "%d envs x %d drones x %d rays, %d steps of the float64 C oracle (OpenMP over envs), no auto-reset" % (E, n, r, steps),
            "seconds": dt}, dt / steps


def quick_device_rate(wl, dev, steps=300, warmup=10):
    """Device-resident throughput of another workload (reported next to the main line, N = 1 only)."""
    import torch
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset
    preset_name, envs, n, r, desc = WORKLOADS[wl]
    gmap, bank = build_world(wl, 256, seed=1000)
    env = BatchedDroneEnv(preset(preset_name, n_envs=envs, n_agents=n, n_rays=r, w_max=32, seed=1000), gmap, device=dev)
    from multi_agent_aac_b200.reset import OdTable
    env.set_od_tables([OdTable(m, w_max=32, planner="device") for m in (gmap if isinstance(gmap, list) else [gmap])])
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    acts = [(torch.rand((envs, n, 2), device=dev, generator=gen) * 2 - 1).contiguous() for _ in range(4)]
    for k in range(warmup):
        env.step(acts[k % 4], autoreset=True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    t0.record()
    for k in range(steps):
        env.step(acts[k % 4], autoreset=True)
    t1.record()
    torch.cuda.synchronize(dev)
    ms = t0.elapsed_time(t1) / steps
    env.close()
    return {"workload": desc, "value": envs * n / (ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms, "steps": steps}


def policy_rollout(env, dev, steps=200, warmup=5):
    """the caller side of the step.  The batched actor (ActorNetwork_allnei_wRadar, random-init
    parameters of the reference's architecture) reads the env's observation tensors where the env kernel left them and
    its actions drive the next step, so a rollout never leaves HBM.  Reports the actor kernel against the measured
    bf16 tensor peak (algorithmic flops = 2 * MACs of the six Linear layers) and the closed-loop rate."""
    import torch
    from multi_agent_aac_b200.actor import BatchedActor
    from multi_agent_aac_b200 import actor_params
    rows = env.E * env.N
    actor = BatchedActor.for_env(env)
    actor.load_state_dict(actor_params.reference_like_params(env.D, 5 * (env.N - 1), env.R, seed=0))
    act = torch.empty((env.E, env.N, 2), dtype=torch.float32, device=dev)
    obs = env.observe()
    for k in range(warmup):
        actor(obs, noise_scale=0.1, noise_seed=k, out=act)
        obs = env.step(act, autoreset=True)[0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize(dev)
    ev[0].record()
    for k in range(steps):
        actor(obs, noise_scale=0.1, noise_seed=100 + k, out=act)
    ev[1].record()
    for k in range(steps):
        actor(obs, noise_scale=0.1, noise_seed=k, out=act)
        obs = env.step(act, autoreset=True)[0]
    ev[2].record()
    torch.cuda.synchronize(dev)
    actor_ms, loop_ms = ev[0].elapsed_time(ev[1]) / steps, ev[1].elapsed_time(ev[2]) / steps
    flop = 2.0 * rows * (128 * (env.D + 5 * (env.N - 1) + env.R) + 384 * 512 + 512 * 256 + 256 * 2)
    peak = None
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    tf = flop / (actor_ms * 1e-3) / 1e12
    return {"what": "actor forward (choose_action for every drone) + env step + auto-reset, closed loop on the device",
            "actor_ms": actor_ms, "actor_rows_per_s": rows / (actor_ms * 1e-3), "loop_ms_per_step": loop_ms,
            "agent_steps_per_s": rows / (loop_ms * 1e-3), "steps": steps, "actor_launches": actor.launch_count,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": (tf / peak) if peak else None,
                         "peak_kind": "measured cuBLAS bf16 burst (MEASURED_PEAKS.json)" if peak else "unavailable",
                         "flop_per_row": flop / rows, "dtype": "bf16 operands, f32 accumulate", "kernel": "actor_kernel (tcgen05, TMEM accumulators)"}}


def policy_rollout_att(env, dev, steps=200, warmup=5):
    """The canonical variant's caller side: the attention actor (ActorNetwork_ATT_TwoPortion, fp32 CUDA-core kernel) on the
    att-preset env's observation tensors, closed loop with the step."""
    import torch
    from multi_agent_aac_b200.actor import BatchedAttActor
    from multi_agent_aac_b200 import actor_params
    rows = env.E * env.N
    actor = BatchedAttActor.for_env(env)
    actor.load_state_dict(actor_params.reference_like_params_att(env.D, env.R, seed=0))
    act = torch.empty((env.E, env.N, 2), dtype=torch.float32, device=dev)
    obs = env.observe()
    for k in range(warmup):
        obs = env.step(actor(obs, noise_scale=0.1, noise_seed=k, out=act), autoreset=True)[0]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize(dev)
    ev[0].record()
    for k in range(steps):
        actor(obs, noise_scale=0.1, noise_seed=100 + k, out=act)
    ev[1].record()
    for k in range(steps):
        obs = env.step(actor(obs, noise_scale=0.1, noise_seed=k, out=act), autoreset=True)[0]
    ev[2].record()
    torch.cuda.synchronize(dev)
    actor_ms, loop_ms = ev[0].elapsed_time(ev[1]) / steps, ev[1].elapsed_time(ev[2]) / steps
    M = env.N - 1
    flop = 2.0 * rows * (64 * (env.D + env.R) + 3 * 64 * 64 + 2 * M * (64 * 6 + 64) + 192 * 256 + 256 * 2)
    return {"what": "attention actor forward + env step + auto-reset, closed loop on the device", "actor_ms": actor_ms,
            "actor_rows_per_s": rows / (actor_ms * 1e-3), "actor_tflops_fp32": flop / (actor_ms * 1e-3) / 1e12, "loop_ms_per_step": loop_ms,
            "agent_steps_per_s": rows / (loop_ms * 1e-3), "steps": steps, "actor_launches": actor.launch_count,
            "kernel": "actor_att_kernel (fp32 GEMM chain on the CUDA cores)"}


def cpu_c1_run(steps=1000, seed=0):
    """BASELINE config 1: the reference's own CPU-runnable case - ONE env of the one_model_att scenario (3 drones, 18 rays),
    seeded random actions, `steps` steps with the caller's episode rule (reset on any done / all reached / 50-step cap,
    ATT/ma_main:448-462), on ONE core: the float64 C port, single-threaded as the reference's loop is."""
    from oracle.oracle import OracleEnv, RADAR_MIN
    os.environ["OMP_NUM_THREADS"] = "1"
    gmap, bank = build_world("c1", 64, seed=123)
    n, r = 3, 18
    orc = OracleEnv("att", gmap, 1, n, r, w_max=32, radar_mode=RADAR_MIN)
    g = gmap.grid_length

    def install(s):
        lines = []
        for i in range(n):
            w = int(bank.w[s, i])
            c = bank.cells[s, i, :w].astype(np.int64)
            lines.append(np.stack([gmap.x0c + (c >> 8) * g, gmap.y0c + (c & 255) * g], -1).astype(np.float64))
        heads = [float(np.arctan2(l[1][1] - l[0][1], l[1][0] - l[0][0])) for l in lines]
        orc.set_episode(0, [l[0] for l in lines], lines, heads)
        orc.observe()
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, size=(steps, 1, n, 2))
    install(0)
    episodes, ep_step = 0, 0
    t0 = time.perf_counter()
    for k in range(steps):
        o = orc.step(acts[k])
        ep_step += 1
        if o["done"].any() or orc.state["reach"].all() or ep_step > 50:
            episodes += 1
            ep_step = 0
            install(episodes % bank.n_scenarios)
    dt = time.perf_counter() - t0
    return {"value": n * steps / dt, "unit": "agent-steps/s", "cores": 1, "kind": "port",
            "sample": "1 env x %d drones x %d rays, %d steps of the float64 C oracle on one core, %d episodes (reset on done / all reached / 50-step cap)" % (n, r, steps, episodes),
            "seconds": dt}, dt / steps


def seek_actions(env, noise):
    """The goal-seeking scripted policy of tests/golden/ref_harness.py (steer at the next waypoint at 0.9 vmax, plus
    noise) as torch ops on the device state: episodes last tens of steps instead of the 2-3 of uniform random actions."""
    import torch
    st, gm, cfg = env.state, env.gmap, env.cfg
    cur = (st["meta"] & 0xFF).long()
    idx = torch.minimum(cur + 1, st["ref_w"].long() - 1)
    code = torch.gather(st["ref_cells"].long() & 0xFFFF, 2, idx[..., None])[..., 0]
    cell = float(gm.grid_length)
    wx = (gm.x0c - gm.origin[0]) + (code >> 8).float() * cell
    wy = (gm.y0c - gm.origin[1]) + (code & 255).float() * cell
    to = torch.stack([wx - st["px"], wy - st["py"]], -1)
    want = to / to.norm(dim=-1, keepdim=True).clamp_min(1e-9) * (0.9 * cfg.vmax)
    vel = torch.stack([st["vx"], st["vy"]], -1)
    return ((want - vel) / (0.5 * cfg.acc_max) + 0.25 * noise).clamp_(-1.0, 1.0).contiguous()


def low_reset_regime(env, dev, acts, steps=200, warmup=100):
    """The same step with a goal-seeking policy driving the drones: few envs finish per step, so the reset launch has
    little to do - the regime a trained policy produces.  Only the env calls are timed (CUDA events around each)."""
    import torch
    E, N = env.E, env.N
    for k in range(warmup):
        env.step(seek_actions(env, acts[k % len(acts)]), autoreset=True)
    env.read_stats(reset=True)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(steps)]
    stream = torch.cuda.current_stream(dev)
    for k in range(steps):
        a = seek_actions(env, acts[k % len(acts)])
        ev[k][0].record(stream)
        env.step(a, autoreset=True)
        ev[k][1].record(stream)
    torch.cuda.synchronize(dev)
    ms = float(np.median([e[0].elapsed_time(e[1]) for e in ev]))
    st = env.read_stats(reset=True)
    return {"what": "scripted goal-seeking policy (steer at the next waypoint at 0.9 vmax + noise) instead of uniform random actions; env calls timed alone",
            "ms_per_step": ms, "agent_steps_per_s": E * N / (ms * 1e-3), "steps": steps,
            "envs_finishing_per_step": float(st[0]) / (E * steps), "mean_episode_steps": float(st[1]) / max(float(st[0]), 1.0)}


_JSON_OUT = None


def keep_stdout_for_the_json_line():
    """Libraries print to file descriptor 1 (NCCL's version banner at the first collective, for one): send everything
    written there to stderr and keep the original stdout for the ONE JSON line the contract asks for."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def bind_to_gpu_cpus(local_rank):
    """Pin this rank to the CPUs NVML reports as local to its GPU (its NUMA node) BEFORE any pinned host memory is allocated,
    so that the pinned buffers of the end-to-end path live next to the GPU's PCIe root.  Returns a description for the line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus local to gpu %d (nvml)" % (len(cpus), local_rank)
    except Exception as e:
        return "unbound (%s)" % type(e).__name__
    return "unbound"


def self_launch(args):
    """`python bench.py --gpus N` outside torchrun: start the N ranks the way the driver does."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    return subprocess.call(cmd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="envs per GPU (default: the workload's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scenarios", type=int, default=512)
    ap.add_argument("--reset-source", default="od", choices=["od", "bank"],
                    help="od: origins/destinations drawn on the device from the map's OD table; bank: pre-planned scenario bank")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-envs", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the short runs of the other single-GPU workloads")
    ap.add_argument("--tile-envs", type=int, default=0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--launches", type=int, default=0, help="auto-reset as 1 fused launch or 2 launches (0: the library's rule)")
    ap.add_argument("--blocks", type=int, default=0, help="timed blocks of --steps steps each (0: at least 10, enough for 0.3 s)")
    args = ap.parse_args()

    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl != "reference" and args.workload != "c1":
        return self_launch(args)   # one line: python bench.py --gpus 8 --workload c5
    keep_stdout_for_the_json_line()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    preset_name, envs, n, r, desc = WORKLOADS[args.workload]
    if args.envs:
        envs = args.envs
    variant = variant_of(preset_name)
    bytes_per = algorithmic_bytes(variant, n, r)

    if args.impl == "reference" or args.workload == "c1":
        if rank != 0:
            return 0
        if args.workload == "c1":   # BASELINE config 1 is a CPU case whatever the arm: one env, one core
            steps, warm = max(1, args.steps), 0
            base, sec_per_step = cpu_c1_run(steps)
            sample = 1
        else:
            sample = args.cpu_envs or max(64, min(8192, 80000 // n))
            steps, warm = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
            base, sec_per_step = cpu_reference_run(args.workload, steps, warm, sample)
        line = {"impl": "reference", "metric": "agent_steps_per_sec", "value": base["value"], "unit": "agent-steps/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec_per_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "sample_envs": sample, "drones": n, "rays": r},
                "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": base["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from multi_agent_aac_b200 import _capi as K
    from multi_agent_aac_b200.env import BatchedDroneEnv, preset

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    affinity = bind_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K.lib()
    gmap, bank = build_world(args.workload, args.scenarios, seed=1000)
    cfg = preset(preset_name, n_envs=envs, n_agents=n, n_rays=r, w_max=32, seed=1000, env_id_base=rank * envs,
                 tile_envs=args.tile_envs, block_threads=args.threads, autoreset_launches=args.launches)
    env = BatchedDroneEnv(cfg, gmap, device=dev)
    if args.reset_source == "bank":
        env.set_bank(bank)
    else:   # origins / destinations drawn on the device at every reset from the maps' OD tables
        from multi_agent_aac_b200.reset import OdTable
        env.set_od_tables([OdTable(m, w_max=32, planner="device") for m in (gmap if isinstance(gmap, list) else [gmap])])
    env.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1 + rank)
    n_act = 8
    acts = [(torch.rand((envs, n, 2), device=dev, generator=gen) * 2 - 1).contiguous() for _ in range(n_act)]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K_ = args.steps
    step_no = [0]

    def timed_block():
        """EXACTLY K steps between two events, a barrier + synchronize on both sides; device time, max over ranks."""
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        h0 = time.perf_counter()
        t0.record(stream)
        for _ in range(K_):
            env.step(acts[step_no[0] % n_act], autoreset=True)   # step + reward launch, then the reset launch for the envs that terminated
            step_no[0] += 1
        t1.record(stream)
        host_ms = (time.perf_counter() - h0) * 1e3 / K_
        barrier()
        return max_over_ranks(t0.elapsed_time(t1)), host_ms

    for k in range(max(args.warmup, 3)):
        env.step(acts[k % n_act], autoreset=True)
    # The driver's K may be small (20 steps = 6 ms): the block is repeated - every block is exactly K steps, timed as the
    # contract says - and the line carries the MEDIAN block with the spread; enough blocks for >= 0.3 s and >= 10 clock samples
    probe_ms, _ = timed_block()
    n_blocks = args.blocks or int(min(400, max(10, np.ceil(300.0 / max(probe_ms, 1e-3)))))
    launches0 = env.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.05)   # let the first samples arrive
    wall_t0 = time.perf_counter()
    blocks = [timed_block() for _ in range(n_blocks)]
    wall_t1 = time.perf_counter()
    launches = (env.launch_count - launches0) // n_blocks
    clocks = sampler.summary(wall_t0, wall_t1) if sampler else None
    block_ms = np.array([b[0] for b in blocks])
    elapsed_ms = float(np.median(block_ms))
    host_issue_ms = float(np.median([b[1] for b in blocks]))
    step_kernel_ms = elapsed_ms / K_   # launches are issued back to back: the block IS the kernels (plus launch gaps on tiny batches)
    stats = torch.tensor(env.read_stats(), device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)  # episode statistics: the only collective of the path
    stats = stats.cpu().numpy()

    # end to end through the host-buffer entry point: pinned actions in, observations / reward / done out
    host = env.host_buffers()
    h_act = [a.cpu().pin_memory() for a in acts[:2]]
    for k in range(2):
        env.step_host(h_act[k % 2], autoreset=True)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.e2e_steps):
        env.step_host(h_act[k % 2], autoreset=True)
    torch.cuda.synchronize(dev)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = int(h_act[0].numel() * 4)
    d2h = int(sum(v.numel() * v.element_size() for v in host.values()))
    # the ceiling of that number: the bare copies of one step (same pinned buffers, same bytes, every rank at once, no kernel)
    def bare_copies(reps):
        barrier()
        c0 = time.perf_counter()
        for _ in range(reps):
            env.d_actions_probe.copy_(h_act[0], non_blocking=True)
            for k_, v in host.items():
                v.copy_(env.out[k_], non_blocking=True)
        torch.cuda.synchronize(dev)
        return max_over_ranks(time.perf_counter() - c0) / reps
    env.d_actions_probe = torch.empty_like(acts[0])
    bare_copies(2)
    copy_s = bare_copies(max(3, args.e2e_steps // 2))

    if rank == 0:
        agents_total = envs * n * world
        value = agents_total * K_ / (elapsed_ms * 1e-3)
        peak, peak_kind = measured_peak()
        achieved = envs * n * bytes_per / (step_kernel_ms * 1e-3) / 1e9
        names = launched_kernels(variant, n, r, cfg.radar_mode, int(round(launches / K_)))
        traffic = profiled_traffic(names) if names else None
        e2e_value = agents_total * args.e2e_steps / e2e_s
        line = {
            "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K_,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / K_, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "envs_per_gpu": envs, "drones": n, "rays": r, "variant": preset_name,
                       "radar_mode": "last_hit" if cfg.radar_mode else "min",
                       "reset": "device-side OD sampling from the map's origin/destination table" if args.reset_source == "od" else "scenario bank of %d" % args.scenarios,
                       "l2": "state+actions+outputs per step = %.0f MB > 126 MB L2, 8 rotating action buffers; no flush needed" % (envs * n * bytes_per / 1e6),
                       "tile_envs": args.tile_envs, "sharding": "envs by instance, no data-path collective"},
            "repeat": {"blocks": n_blocks, "steps_per_block": K_, "ms_per_step_median": elapsed_ms / K_, "ms_per_step_min": float(block_ms.min()) / K_,
                       "ms_per_step_max": float(block_ms.max()) / K_, "spread": float((block_ms.max() - block_ms.min()) / np.median(block_ms)),
                       "note": "every block = exactly --steps steps between barrier + synchronize, CUDA events, max over ranks; value / ms_per_step are the median block"},
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                    "roofline": {"bound": "host link (PCIe D2H into pinned memory)", "peak": agents_total / copy_s, "unit": "agent-steps/s", "frac": e2e_value * copy_s / agents_total,
                                 "peak_GBs": (h2d + d2h) * world / copy_s / 1e9, "achieved_GBs": (h2d + d2h) * world * args.e2e_steps / e2e_s / 1e9,
                                 "how": "the same pinned buffers and byte counts copied with no kernel in between, all ranks at once, timed the same way",
                                 "cpu_affinity": affinity}},
            "gpu_launches": int(launches) * n_blocks,
            "kernels": {"env_kernel(step+autoreset)_ms": step_kernel_ms, "host_issue_ms_per_step": host_issue_ms, "launches_per_step": launches / K_},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic,
                         "traffic_note": ("bytes per step, summed over the step's launch(es), profiles/%s: one ncu --set full capture of each, of this very source (hash %s)" % (" + ".join(PROFILED[len(names)]), source_hash()))
                                         if traffic else "no ncu capture of this build (source hash %s) and these kernels under profiles/: null" % source_hash(),
                         "peak_kind": peak_kind, "bytes_per_agent_step": bytes_per,
                         "kernel": " + ".join(names) if names else "env_kernel<%s>" % variant.upper()},
            "clocks": clocks,
            "issue_slots": None,
            "episode_stats": {k: float(v) for k, v in zip(K.STAT_NAMES, stats)},
        }
        # What bounds the kernel in practice (it is not HBM): warp-instruction issue.  This is a UTILISATION of the issue slots
        # by the kernel's own instruction count - not a roofline fraction: it would rise if the kernel executed more instructions.
        inst = profiled_metric("smsp__inst_executed.sum", names) if names else None
        if inst and clocks and clocks.get("sm_mhz"):
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            peak_issue = 4.0 * sms * clocks["sm_mhz"] * 1e6
            line["issue_slots"] = {"warp_inst_per_agent_step": inst / (envs * n), "utilisation": inst / (step_kernel_ms * 1e-3) / peak_issue,
                                   "note": "instructions per step (all launches of the step) from the same committed captures; 4 schedulers x SMs x sampled clock"}
        if world == 1 and not args.no_aux and preset_name == "tdcpa_v2":
            line["low_reset_regime"] = low_reset_regime(env, dev, acts)
            line["policy_rollout"] = policy_rollout(env, dev)
        if world == 1 and not args.no_aux and preset_name == "att":
            line["policy_rollout"] = policy_rollout_att(env, dev)
        if world == 1 and not args.no_aux:   # the other single-GPU configurations of BASELINE.json, device-resident
            line["other_workloads"] = {w: quick_device_rate(w, dev) for w in ("c2", "c4") if w != args.workload}
        if not args.no_cpu and world == 1:
            sample = args.cpu_envs or max(64, min(8192, 80000 // n))
            base, _ = cpu_reference_run(args.workload, 30, 1, sample)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
