"""Sensor sources: what Isaac Gym hands the reference env every step.

In the reference, `gym.render_all_camera_sensors` refreshes per-env GPU image tensors (depth / segmentation /
RGBA, env_train_gennbv.py:204-227,349-354) and `get_camera_view_matrix` returns the per-env view matrices as a
host numpy array (env_train_base.py:777-785).  Isaac Gym is closed source and absent here, so the env takes an
injectable source with the same outputs; `SyntheticHouseSensor` renders the analytic scenes of gennbv_b200.synth
on the device, `ReplaySensor` replays recorded frames (used by the parity tests)."""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import synth


@dataclass
class SensorFrame:
    depth: torch.Tensor                 # [N,H,W] f32 raw (negative z-depth, -inf = no hit), on the env's device
    seg: torch.Tensor                   # [N,H,W] i32
    rgba: Optional[torch.Tensor]        # [N,H,W,4] u8 or None
    c2w: Optional[torch.Tensor] = None          # [N,4,4] f32 env-local camera-to-world on the device, or
    view_matrix: Optional[np.ndarray] = None    # [N,4,4] f32 host array in Isaac's convention (one of the two)
    contact: Optional[torch.Tensor] = None      # [N] u8 collision flags (None = no contacts)


class SensorSource:
    height: int
    width: int

    def render(self, poses: torch.Tensor) -> SensorFrame:      # poses [N,6] f32 (x,y,z,roll,pitch,yaw)
        raise NotImplementedError


class SyntheticHouseSensor(SensorSource):
    """Analytic box+gable-roof scenes (gennbv_b200.synth), rendered on the device for the current poses."""

    def __init__(self, scene_params, height, width, fov_deg=90.0, with_rgb=True, device="cuda"):
        self.params = scene_params.to(device)
        self.height, self.width, self.fov, self.with_rgb = height, width, fov_deg, with_rgb

    def render(self, poses):
        depth, seg, rgb, c2w = synth.render(self.params, poses, self.height, self.width, self.fov, with_rgb=self.with_rgb)
        return SensorFrame(depth=depth.contiguous(), seg=seg.contiguous(), rgba=rgb, c2w=c2w.float().contiguous())


class ReplaySensor(SensorSource):
    """Replays recorded raw frames + Isaac-convention view matrices (host), ignoring the requested poses."""

    def __init__(self, depth, seg, rgba, view_matrix, device):
        self.depth, self.seg, self.rgba, self.view = depth, seg, rgba, view_matrix
        self.height, self.width = depth.shape[-2:]
        self.device, self.t = device, 0

    def render(self, poses):
        t = self.t
        self.t += 1
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a[t])).to(self.device)
        return SensorFrame(depth=d(self.depth), seg=d(self.seg), rgba=d(self.rgba) if self.rgba is not None else None,
                           view_matrix=np.ascontiguousarray(self.view[t]))


class FrameListSensor(SensorSource):
    """Cycles through pre-rendered frames (dicts with depth / seg / rgba / c2w).  With `host=True` the frames live in
    pinned host memory and every render() issues the host->device copies on the current stream -- the shape of a real
    deployment where the simulator hands over host or foreign-device buffers.

    `prefetch_next()` (host frames only) starts the copy of the NEXT frame on a side stream into the second of two device
    frame sets; the following render() then only waits for that copy.  Causality of the RL loop is the caller's business:
    the next frame exists once the next action has been chosen, i.e. after the policy forward on the current observation,
    so a training step calls it between its forward and its backward and the copy hides under the backward."""

    def __init__(self, frames, device, host=False):
        self.device, self.host, self.t = torch.device(device), host, 0
        keys = ("depth", "seg", "rgba", "c2w")
        if host:
            self.frames = [{k: f[k].cpu().pin_memory() for k in keys if f.get(k) is not None} for f in frames]
            self.dev = [{k: torch.empty_like(v, device=self.device) for k, v in self.frames[0].items()} for _ in range(2)]
            self._copy_stream = None
            self._ready = [None, None]             # event of an in-flight prefetch into set i, for frame index _ready_t[i]
            self._ready_t = [-1, -1]
        else:
            self.frames = [{k: f[k].to(self.device).contiguous() for k in keys if f.get(k) is not None} for f in frames]
        self.height, self.width = self.frames[0]["depth"].shape[-2:]
        self.bytes_per_frame = sum(v.numel() * v.element_size() for v in self.frames[0].values())

    def prefetch_next(self):
        if not self.host:
            return
        t, main = self.t, torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        s = t % 2
        # set s was last read by the kernels of step t-2, all enqueued on `main` before this call
        self._copy_stream.wait_stream(main)
        with torch.cuda.stream(self._copy_stream):
            for k, v in self.frames[t % len(self.frames)].items():
                self.dev[s][k].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._ready[s], self._ready_t[s] = ev, t

    def render(self, poses):
        t = self.t
        f = self.frames[t % len(self.frames)]
        self.t += 1
        if self.host:
            s = t % 2
            if self._ready_t[s] == t:
                torch.cuda.current_stream(self.device).wait_event(self._ready[s])
            else:
                for k, v in f.items():
                    self.dev[s][k].copy_(v, non_blocking=True)
            f = self.dev[s]
        return SensorFrame(depth=f["depth"], seg=f["seg"], rgba=f.get("rgba"), c2w=f["c2w"])
