#!/usr/bin/env python
"""bench.py -- env-steps/s of the GenNBV state-encoding + encoder hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus 1 --steps 30 --warmup 5                  # this framework (CUDA through the C ABI)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the reference's CPU PyTorch path on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...            # env-parallel, one rank per GPU (weak scaling)

A "step" is BASELINE.json configs[1] for 256 environments per GPU: one env.step() worth of state encoding (sensor
post-processing, depth un-projection, voxel scatter, Bresenham ray-cast, prob / tri-class / scanned-GT update, coverage
reward, termination, reset) on 128x128 synthetic depth and a 64^3 grid, followed by the 3D-CNN Hybrid_Encoder forward
(batch-statistics BatchNorm) and backward on the 256 fresh observations; with N > 1 GPUs the flat encoder gradient is
all-reduced over NCCL/NVLink every step.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

ENVS_PER_GPU, H, W, G, SCENES, FRAMES, STATE_DIM = 256, 128, 128, 64, 8, 4, 600
METRIC, UNIT = "env_steps_per_sec", "env-steps/s"
WORKLOAD = ("BASELINE configs[1]: 256 envs/GPU x 128x128 synthetic depth x 64^3 grid -- env.step() state encoding "
            "(post-process + unproject + voxel scatter + Bresenham ray-cast + prob/tri-class/scanned-GT update + coverage "
            "reward + termination/reset) + Hybrid_Encoder (3D-CNN) forward/backward on the 256 observations")


CONFIG = {"workload": WORKLOAD, "envs_per_gpu": ENVS_PER_GPU, "depth": [H, W], "grid": G, "frames_rotated": FRAMES,
          "l2": "per-step working set (prob+scanned+gt grids 0.8 GB, observations 0.28 GB, conv1 activations 0.49 GB) exceeds "
                "the 126 MB L2; no explicit flush"}          # identical in both arms (the driver compares it)


def algorithmic_bytes_per_env_step(P, V):
    """voxelize P*(4+4) + 64 + V*(4r+4w prob) + V*4 tri; coverage V*(4 gt + 4r + 4w scanned) + 4."""
    return P * 8 + 64 + V * 12, V * 12 + 4


def make_workload(num_envs, device, seed):
    """Synthetic scenes + FRAMES rendered views per env (raw sensor convention), rendered on `device`."""
    from gennbv_b200 import synth
    scenes = synth.make_house_scenes(SCENES, G, seed=seed)
    gen = torch.Generator().manual_seed(seed + 1)
    frames, actions = [], []
    for _ in range(FRAMES):
        a = synth.sample_lookat_actions(scenes.params, num_envs, gen)
        poses = synth.pose_from_action(a).to(device)
        depth, seg, rgb, c2w = synth.render(scenes.params, poses, H, W, with_rgb=True)
        frames.append(dict(depth=depth.contiguous(), seg=seg.contiguous(), rgba=rgb.contiguous(),
                           c2w=c2w.float().contiguous(), xyz=poses[:, :3].contiguous()))
        actions.append(a.to(device))
    return dict(scenes=scenes, frames=frames, actions=actions)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            time.sleep(0.3)          # let the first samples arrive before the timed region starts
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, ln in self.lines:
            if t_begin is not None and not (t_begin - 0.06 <= ts <= t_end + 0.12):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx = int(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_rate(wl, num_envs, threads, steps, warmup):
    """The reference's CPU PyTorch path for the same step: oracle/torch_ref.py restates update_occ_grid op for op
    (einsum, floor, unique, index_put ... in per-env Python loops; the PyCUDA Bresenham kernel served by the C
    restatement), oracle/encoder_ref.py is the reference encoder with the grid size parametrised; forward in train
    mode + backward.  torch intra-op threads = `threads`.  Returns (env-steps/s, seconds/step, stage seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_ref
    import encoder_ref
    from gennbv_b200 import synth
    torch.set_num_threads(threads)
    scenes = wl["scenes"]
    vs, nvalid, rg = synth.gt_metadata(scenes.grid_gt)
    idx = torch.arange(num_envs) % SCENES
    vs, rg, gt = vs[idx].contiguous(), rg[idx].contiguous(), scenes.grid_gt[..., 3][idx].contiguous()
    kinv = torch.linalg.inv(synth.camera_intrinsics(H, W)).float()
    pix = torch_ref.pixel_grid(H, W)
    prob = torch.zeros(num_envs, G, G, G); scan = torch.zeros_like(prob)
    frames = [{k: v[:num_envs].cpu() for k, v in f.items()} for f in wl["frames"]]
    enc = encoder_ref.HybridEncoderRef(G, STATE_DIM)
    enc.train()
    obs = torch.zeros(num_envs, STATE_DIM + G ** 3 + 8192)
    acc = {"vox": 0.0, "enc": 0.0}

    def one(i, timed):
        f = frames[i % FRAMES]
        t0 = time.perf_counter()
        tri, cov = torch_ref.voxelize_step(f["depth"], f["seg"], kinv, f["c2w"], rg, vs, f["xyz"], gt, prob, scan, pix)
        obs[:, STATE_DIM:STATE_DIM + G ** 3] = tri.view(num_envs, -1)
        t1 = time.perf_counter()
        enc.zero_grad()
        enc(obs).sum().backward()
        t2 = time.perf_counter()
        if timed:
            acc["vox"] += t1 - t0; acc["enc"] += t2 - t1

    for s in range(warmup):
        one(s, False)
    t0 = time.perf_counter()
    for s in range(steps):
        one(warmup + s, True)
    dt = time.perf_counter() - t0
    return num_envs * steps / dt, dt / steps, {"voxelize+coverage": acc["vox"] / steps, "encoder_fwd_bwd": acc["enc"] / steps}


def run_reference(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = make_workload(ENVS_PER_GPU, "cpu", seed=0)
    if args.cpu_envs is None:
        # bounded sample: size the per-step env count so that the whole run stays near two minutes
        probe, _, _ = cpu_reference_rate(wl, 8, threads, 1, 1)
        n = int(max(8, min(ENVS_PER_GPU, 120.0 * probe / (args.steps + args.warmup))))
    else:
        n = args.cpu_envs
    rate, sec, stages = cpu_reference_rate(wl, n, threads, args.steps, args.warmup)
    sample = (f"{n} of {ENVS_PER_GPU} envs per step, {args.steps} steps; PyTorch-CPU restatement of the reference's "
              f"update_occ_grid path (oracle/torch_ref.py) + reference encoder fwd/bwd (oracle/encoder_ref.py), "
              f"{threads} torch threads")
    _emit({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3 * ENVS_PER_GPU / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": CONFIG, "note": "host CPU arm; ms_per_step extrapolated to 256 envs",
        "stages_ms_per_sample_step": {k: v * 1e3 for k, v in stages.items()},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


# --------------------------------------------------------------------------------------------- GPU arm
def build_native(wl, dev, host_frames):
    from gennbv_b200.config import Config_GenNBV_Train
    from gennbv_b200.env import Env_Train_GenNBV
    from gennbv_b200.policy import ActorCriticPolicy_Train_Eval
    from gennbv_b200.sensors import FrameListSensor
    from gennbv_b200.wrapper import EnvWrapperGenNBVTrain

    class Cfg(Config_GenNBV_Train):
        class rewards(Config_GenNBV_Train.rewards):
            only_positive_rewards = False

    sensor = FrameListSensor(wl["frames"], dev, host=host_frames)
    env = EnvWrapperGenNBVTrain(Env_Train_GenNBV(Cfg(), sim_device=str(dev), sensor=sensor, grid_gt=wl["scenes"].grid_gt,
                                                 num_envs=ENVS_PER_GPU))
    kwargs = dict(encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
                  net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
                  state_input_shape=(STATE_DIM,), visual_input_shape=(100, H, W))
    torch.manual_seed(0)
    policy = ActorCriticPolicy_Train_Eval(env.observation_space, env.action_space, lambda _: 1e-4, net_arch=[],
                                          features_extractor_kwargs=kwargs, device=dev)
    policy.train(True)
    return env, sensor, policy


def ppo_iteration(wl, dev, world, n_steps, batch_size=128, n_epochs=5, target_kl=0.05, iters=1):
    """BASELINE configs[2] / [3]: one full PPO iteration through the public API -- `PPO_Grid_Obs.collect_rollouts()` (256 envs x
    n_steps env.step + policy) and `train()` (n_epochs x minibatches of `batch_size`, gradient all-reduce when world > 1) --
    after one untimed warm-up iteration.  Device time is bracketed by synchronisations; env-steps/s = envs * n_steps / total."""
    import torch.distributed as dist
    from gennbv_b200.config import Config_GenNBV_Train
    from gennbv_b200.env import Env_Train_GenNBV
    from gennbv_b200.ppo import PPO_Grid_Obs
    from gennbv_b200.sensors import FrameListSensor
    from gennbv_b200.wrapper import EnvWrapperGenNBVTrain

    class Cfg(Config_GenNBV_Train):
        class rewards(Config_GenNBV_Train.rewards):
            only_positive_rewards = False

    N = ENVS_PER_GPU
    env = EnvWrapperGenNBVTrain(Env_Train_GenNBV(Cfg(), sim_device=str(dev), sensor=FrameListSensor(wl["frames"], dev),
                                                 grid_gt=wl["scenes"].grid_gt, num_envs=N))
    kw = dict(net_arch=[], features_extractor_kwargs=dict(
        encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
        net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
        state_input_shape=(STATE_DIM,), visual_input_shape=(100, H, W)))
    algo = PPO_Grid_Obs(env=env, learning_rate=1e-4, n_steps=n_steps, batch_size=batch_size, n_epochs=n_epochs, gamma=0.99,
                        gae_lambda=0.95, clip_range=0.2, clip_range_vf=0.2, ent_coef=0.01, vf_coef=0.8, max_grad_norm=1,
                        target_kl=target_kl, policy_kwargs=kw, seed=1, device=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    algo._setup_learn()
    algo.collect_rollouts()                 # warm-up iteration: allocations, graph capture, first touch of the 35 GB buffer
    algo.train()
    roll, train, opt_steps, stopped = [], [], [], []
    for _ in range(iters):
        sync(); t0 = time.perf_counter()
        algo.collect_rollouts()
        sync(); t1 = time.perf_counter()
        before = algo._adam_step
        algo.train()
        sync(); t2 = time.perf_counter()
        roll.append(t1 - t0); train.append(t2 - t1)
        opt_steps.append(algo._adam_step - before); stopped.append(algo._last_train["stopped_epoch"])
    mb_run = [algo._last_train["minibatches_logged"]]
    t = torch.tensor([float(np.median(roll)), float(np.median(train))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rollout_s, train_s = t.tolist()
    n_mb = n_epochs * (-(-N * n_steps // batch_size))
    out = {"workload": f"BASELINE configs[{2 if world == 1 else 3}]: {N} envs/GPU x {n_steps} steps rollout + {n_epochs} epochs x "
                       f"minibatches of {batch_size} through PPO_Grid_Obs.collect_rollouts()/train(), target_kl={target_kl}",
           "rollout_s": rollout_s, "train_s": train_s, "optimizer_steps": opt_steps[-1], "minibatches_scheduled": n_mb,
           "minibatches_logged": mb_run[-1], "kl_stopped_epoch": stopped[-1],
           "ms_per_minibatch_update": train_s * 1e3 / max(1, n_mb if stopped[-1] is None else (stopped[-1] + 1) * (n_mb // n_epochs)),
           "env_steps_per_sec": world * N * n_steps / (rollout_s + train_s),
           "rollout_env_steps_per_sec": world * N * n_steps / rollout_s, "cuda_graph": bool(algo._graphs),
           "rollout_buffer_gb": algo.rollout_buffer.observations.numel() * 4 / 1e9}
    del algo, env
    torch.cuda.empty_cache()
    return out


def run_native(args, rank, world):
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # bounded collectives: a wedged peer makes the run fail after 3 minutes instead of hanging the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    from gennbv_b200 import dist as gdist
    N, V, P = ENVS_PER_GPU, G ** 3, H * W
    wl = make_workload(N, dev, seed=rank)
    K, Wm = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_step(host_frames):
        env, sensor, policy = build_native(wl, dev, host_frames)
        enc = policy.features_extractor
        grads = policy.encoder_grad_views()
        dfeat = torch.full((N, 256), 1.0 / N, device=dev)
        overlap = gdist.OverlappedGradAllreduce(policy.grad_bucket, 4 + policy.linear_slice_offset)
        env.reset()

        def step(i, ev=None):
            if ev is not None:
                env._gym_env.profile_events = ev[1:4]
                ev[0].record()
            obs, rew, done, info = env.step(wl["actions"][i % FRAMES])
            if ev is not None:
                ev[4].record()
            feats = enc._run_forward(obs, need_bwd=True, training=True)
            if ev is not None:
                ev[5].record()
            # the next action exists once the forward is done: the host sensor frame of step i+1 can cross PCIe under the backward
            sensor.prefetch_next()
            # PPO gradient all-reduce over NCCL/NVLink (configs[3]): the Linear-layer slice (99.9 % of the bytes) is final
            # after the first phase of the backward and travels under the convolution backward
            enc._run_backward(obs, feats, dfeat, N, True, enc._ws, grads=grads, phases=1)
            overlap.start_linear()
            enc._run_backward(obs, feats, dfeat, N, True, enc._ws, grads=grads, phases=2)
            overlap.finish()
            if ev is not None:
                ev[6].record()
            return rew, feats

        return step, env, sensor, policy

    # ---- device-resident arm: `value`
    step, env, sensor, policy = make_step(host_frames=False)
    for i in range(Wm):
        step(i)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(7)] for _ in range(K)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()           # before the barrier: its start-up sleep must not skew rank 0 against the others
    barrier()
    from gennbv_b200 import _lib
    _lib.check(_lib.lib().gnbv_profile_enable(1), "gnbv_profile_enable")      # stage events inside the encoder calls
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    torch.cuda.nvtx.range_push("timed")          # ncu --nvtx --nvtx-include "timed/" isolates the step's kernels
    t_start.record()
    for i in range(K):
        step(Wm + i, ev[i])
    t_end.record()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    barrier()
    wall1 = time.time()
    ms_total = t_start.elapsed_time(t_end)
    kernel_ms = {k: _lib.stage_ms(k) for k in _lib.ENCODER_STAGES}          # device time of each encoder stage, last timed step
    _lib.lib().gnbv_profile_enable(0)
    # bytes the sparse grid update actually moved in the last step, from its own masks (outside the timed region): every voxel
    # costs prob read + tri write; a 16-byte group with a touched bit adds the prob write, one with a target bit the scanned
    # read + write and the GT read
    from gennbv_b200 import ops as _ops
    genv = env._gym_env
    tmask, rmask = _ops.voxelize_masks(genv._workspace, N, G)
    grp = lambda m: int(m.view(N, -1, 4).any(dim=2).sum())
    g_t, g_r = grp(tmask), grp(rmask | tmask)
    grid_bytes_moved = N * V * 8 + g_r * 16 + g_t * 48 + N * 4
    grid_sparse = bool(getattr(genv, "_sparse_update", False))
    env._gym_env.profile_events = None
    stage = lambda a, b: float(np.median([e[a].elapsed_time(e[b]) for e in ev]))
    per_step = np.array([e[0].elapsed_time(e[6]) for e in ev])
    step_spread = {"min": float(per_step.min()), "median": float(np.median(per_step)), "max": float(per_step.max())}
    stages = {"env.step total": stage(0, 4), "scan_raycast": stage(1, 2), "grid_update+coverage": stage(2, 3),
              "encoder_forward": stage(4, 5), "encoder_backward(+allreduce)": stage(5, 6)}
    # kernel launches of one step, counted with the CUDA profiler (every kernel of libgennbv_b200 lives in namespace gnbv)
    # (every rank runs this extra step: it contains the gradient all-reduce, which must stay matched across ranks)
    launches = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(Wm + K)
            torch.cuda.synchronize()
        launches = sum(1 for e in prof.events() if "gnbv::" in e.name)
    except ImportError:
        step(Wm + K)
        torch.cuda.synchronize()
    del step, env, sensor, policy
    torch.cuda.empty_cache()

    # ---- end to end through the public API with HOST sensor buffers (pinned): H2D of every frame and D2H of the rewards
    #      and a feature checksum inside the timed region
    step, env, sensor, policy = make_step(host_frames=True)
    rew_host = torch.empty(N, dtype=torch.float32).pin_memory()
    chk_host = torch.empty(1, dtype=torch.float32).pin_memory()

    def e2e_step(i):
        rew, feats = step(i)
        rew_host.copy_(rew, non_blocking=True)
        chk_host.copy_(feats.sum().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the caller consumes rewards / loss every step

    for i in range(Wm):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        e2e_step(Wm + i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    h2d, d2h = sensor.bytes_per_frame, N * 4 + 4
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    del step, env, sensor, policy
    torch.cuda.empty_cache()

    keys = list(stages)
    kkeys = list(kernel_ms)
    t = torch.tensor([ms_total, ms_e2e] + [stages[k] for k in keys] + [kernel_ms[k] for k in kkeys], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vals = t.tolist()
    ms_total, ms_e2e = vals[0], vals[1]
    stages = dict(zip(keys, vals[2:2 + len(keys)]))
    kernel_ms = dict(zip(kkeys, vals[2 + len(keys):]))

    # ---- BASELINE configs[2] (1 GPU) / configs[3] (N GPUs): a full PPO iteration through PPO_Grid_Obs
    ppo = None
    if not args.no_ppo:
        ppo = ppo_iteration(wl, dev, world, n_steps=args.ppo_steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    bf16_peak, bf16_src = 2250.0, "nominal dense bf16 (B200_PROFILING.md)"
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        mp = json.load(open(pk))
        if mp.get("bf16_tflops_sustained"):
            bf16_peak, bf16_src = float(mp["bf16_tflops_sustained"]), "measured sustained cuBLAS bf16 (MEASURED_PEAKS.json)"
    b_vox, b_cov = algorithmic_bytes_per_env_step(P, V)
    G1 = (G - 3) // 2 + 1
    G2 = (G1 - 3) // 2 + 1
    D = STATE_DIM + V + 8192
    y1_b, y2_b, w_fc = N * G1 ** 3 * 16 * 4, N * G2 ** 3 * 16 * 4, 256 * 16 * G2 ** 3 * 4
    # per-kernel algorithmic work (SURVEY.md 8d style: the tensors a kernel must read and write once, reference dtypes) and
    # fp32 flops; the binding roofline of a kernel is the larger of bytes / HBM peak and 3 x flops / (bf16 peak / 2) -- the
    # contractions run as 3 TF32 MMAs per product (fp32-level accuracy, 1e-4 parity budget) and TF32 runs at half the bf16 rate
    work = {
        "scan_raycast": (N * P * 8 + N * 64, 0), "grid_update+coverage": (grid_bytes_moved if grid_sparse else N * (V * 24 + 4), 0),
        "fwd.conv1": (N * V * 4 + y1_b, 2 * N * G1 ** 3 * 16 * 27), "fwd.conv2": (y1_b + y2_b, 2 * N * G2 ** 3 * 16 * 432),
        "fwd.grid_fc": (y2_b + w_fc, 2 * N * 16 * G2 ** 3 * 256), "bwd.grid_fc": (2 * y2_b + 2 * w_fc, 4 * N * 16 * G2 ** 3 * 256),
        "bwd.conv2_wgrad": (y1_b + y2_b, 2 * N * G2 ** 3 * 16 * 432), "bwd.conv2_dgrad": (2 * y1_b + y2_b, 2 * N * G2 ** 3 * 16 * 432),
        "bwd.conv1_wgrad": (2 * y1_b + N * V * 4, 2 * N * G1 ** 3 * 16 * 27)}
    live = {"scan_raycast": stages["scan_raycast"], "grid_update+coverage": stages["grid_update+coverage"], **kernel_ms}
    kernels = {}
    for k, (nbytes, flops) in work.items():
        ms = live[k]
        t_hbm, t_tc = nbytes / (peak * 1e9) * 1e3, 3 * flops / (bf16_peak / 2 * 1e12) * 1e3
        kernels[k] = {"ms": ms, "share_of_step": ms / (ms_total / K), "algorithmic_mb": nbytes / 1e6, "gflop_fp32": flops / 1e9,
                      "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_hbm": nbytes / (ms * 1e-3) / 1e9 / peak,
                      "achieved_tflops_fp32_equiv": flops / (ms * 1e-3) / 1e12 if flops else None,
                      "frac_tensor_3xtf32": (t_tc / ms) if flops else None, "bound": "hbm" if t_hbm >= t_tc else "tensor",
                      "frac_of_binding_roofline": max(t_hbm, t_tc) / ms}
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    dk = kernels[dom]
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from ONE `ncu --set full` capture of this command at the default
    # workload (profiles/r02g_ncu_full_step.txt; a property of the kernel and the shapes, not of this run's timing)
    ncu_dram_mb = {"scan_raycast": 33.7, "grid_update+coverage": 640.7, "fwd.conv1": 718.6, "fwd.conv2": 545.5, "fwd.grid_fc": 116.7,
                   "bwd.grid_fc": 158.4, "bwd.conv2_wgrad": 557.9, "bwd.conv2_dgrad": 1078.5, "bwd.conv1_wgrad": 1263.9}
    default_shape = (N, G, P) == (256, 64, 128 * 128) and grid_sparse
    for k in kernels:
        kernels[k]["ncu_dram_mb"] = ncu_dram_mb.get(k) if default_shape else None
    traffic = ncu_dram_mb[dom] * 1e6 if default_shape and dom in ncu_dram_mb else None
    if dk["bound"] == "hbm":
        roofline = {"bound": "hbm", "kernel": dom, "achieved": dk["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dk["frac_hbm"],
                    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": work[dom][0]}
    else:
        roofline = {"bound": "tensor", "kernel": dom, "achieved": dk["achieved_tflops_fp32_equiv"], "peak": bf16_peak, "unit": "TFLOP/s",
                    "frac": dk["achieved_tflops_fp32_equiv"] / bf16_peak, "traffic": traffic, "peak_source": bf16_src,
                    "algorithmic_flops_per_launch": work[dom][1]}
    roofline["ms"] = dk["ms"]
    roofline["share_of_step"] = dk["share_of_step"]
    roofline["note"] = ("dominant kernel of the step by live device time (CUDA events on the launching stream inside the timed region); "
                        "algorithmic bytes/flops per launch in DESIGN.md section 4; frac_of_binding_roofline in `kernels` "
                        "compares with the larger of the HBM and 3xTF32 tensor floors")
    gk = kernels["grid_update+coverage"]
    out = {
        "metric": METRIC, "value": world * N * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": CONFIG,
        "stages_ms": stages, "step_ms_spread": step_spread, "encoder_kernel_ms": kernel_ms,
        "roofline": roofline,
        "roofline_hbm": {"bound": "hbm", "kernel": "grid_update_sparse_kernel" if grid_sparse else "grid_update_kernel",
                         "achieved": gk["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": gk["frac_hbm"],
                         "traffic": ncu_dram_mb["grid_update+coverage"] * 1e6 if default_shape else None, "ms": gk["ms"],
                         "bytes_moved_per_launch": work["grid_update+coverage"][0],
                         "dense_algorithmic_bytes_per_launch": N * (V * 24 + 4),
                         "speedup_vs_dense_figure": N * (V * 24 + 4) / work["grid_update+coverage"][0],
                         "note": "SURVEY 8d: a kernel that moves fewer bytes than the dense fp32 figure reports GB/s on the bytes it moved "
                                 "(counted from its own masks: 8 B per voxel + 16 B per touched 16-byte group + 48 B per target group) "
                                 "and the ratio to the dense figure; the dense kernel measured 0.247 ms = 0.995 of the copy peak",
                         "peak_source": peak_src, "voxelize_algorithmic_bytes_per_step": N * (b_vox + b_cov)},
        "kernels": kernels,
        "kernel_modes": {"GNBV_CONV2_TC": _lib.lib().gnbv_kernel_mode(0), "GNBV_CONV1_MMA": _lib.lib().gnbv_kernel_mode(1),
                         "GNBV_GEMM_MMA": _lib.lib().gnbv_kernel_mode(2)},
        "e2e": {"value": world * N * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K, "prefetch": "host frame i+1 copied on a side stream under step i's backward"},
        "gpu_launches": (launches or 0) * K, "gpu_launches_per_step": launches,
        "clocks": clocks,
    }
    if ppo is not None:
        out["ppo_iteration"] = ppo
    if world == 1 and not args.no_cpu_baseline:
        n, threads = 16, os.cpu_count() or 1
        wl_cpu = {"scenes": wl["scenes"], "frames": [{k: v[:n].cpu() for k, v in f.items()} for f in wl["frames"]]}
        rate, sec, st = cpu_reference_rate(wl_cpu, n, threads, 2, 1)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{n} of {N} envs x 2 steps; PyTorch-CPU restatement of the reference path "
                                         f"(oracle/torch_ref.py + oracle/encoder_ref.py), stages s/step: "
                                         + ", ".join(f"{k} {v:.2f}" for k, v in st.items())}
    _emit(out)
    if world > 1:
        dist.destroy_process_group()


_STDOUT_FD = None


def _emit(record):
    """Print the one JSON line on the real stdout (see main())."""
    sys.stdout.flush()
    if _STDOUT_FD is not None:
        os.dup2(_STDOUT_FD, 1)
    print(json.dumps(record), flush=True)


def main():
    if os.environ.get("GNBV_BENCH_FAULT_DUMP"):          # debugging aid: dump all Python stacks if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["GNBV_BENCH_FAULT_DUMP"]), exit=False)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--cpu-envs", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ppo", action="store_true", help="skip the full PPO iteration record (BASELINE configs[2]/[3])")
    ap.add_argument("--ppo-steps", type=int, default=128, help="n_steps of the PPO iteration record (reference default 128)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    # stdout carries exactly ONE line, the JSON record: until it is printed, file descriptor 1 points at stderr, so that
    # anything a library writes to stdout on the way (NCCL's version banner at the first collective) cannot precede it
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args, rank)
    else:
        args.steps = 50 if args.steps is None else args.steps
        args.warmup = 5 if args.warmup is None else max(args.warmup, 3)
        run_native(args, rank, world)


if __name__ == "__main__":
    main()
