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


def quat_from_euler_xyz(roll, pitch, yaw):
    """isaacgym.torch_utils.quat_from_euler_xyz: (x, y, z, w) from extrinsic XYZ Euler angles."""
    cy, sy = torch.cos(yaw * 0.5), torch.sin(yaw * 0.5)
    cr, sr = torch.cos(roll * 0.5), torch.sin(roll * 0.5)
    cp, sp = torch.cos(pitch * 0.5), torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qx, qy, qz, qw], dim=-1)


class IsaacGymSensor(SensorSource):
    """Adapter for a live Isaac Gym simulation: what the reference env does around its camera tensors, behind the
    `SensorSource` interface.

    `render(poses)` = `set_state(poses)` (env_train_base.py:686-714: root position = pose xyz + env origin, orientation from
    roll / pitch / yaw), `gym.simulate` + `fetch_results` (env_train_gennbv.py:257-261), then `step_graphics`,
    `render_all_camera_sensors`, `start_access_image_tensors` ... `end_access_image_tensors` (:349-354) around ONE stacked
    copy of the per-env wrapped image tensors (:204-227), and `get_camera_view_matrix` for every env (env_train_base.py:777-785,
    a host array in Isaac's row-vector convention; the env turns it into camera-to-world exactly as the reference does).

    The Isaac Gym modules are injected (`gym`, `gymapi`, `gymtorch`), so that this file imports without Isaac Gym and the
    adapter can be exercised against a mock (tests/test_isaac_sensor.py); nothing here touches the closed-source API beyond
    the calls listed above."""

    def __init__(self, gym, sim, envs, camera_handles, root_states, env_origins, height, width, gymapi, gymtorch, skip=1,
                 contact_forces=None, contact_threshold=1.0):
        self.gym, self.sim, self.envs, self.camera_handles = gym, sim, list(envs), list(camera_handles)
        assert len(self.envs) == len(self.camera_handles), "one camera per env (env_train_base.py:781)"
        self.root_states, self.env_origins, self.skip = root_states, env_origins, skip
        self.height, self.width, self.gymtorch = height, width, gymtorch
        self.contact_forces, self.contact_threshold = contact_forces, contact_threshold
        wrap = lambda kind: [gymtorch.wrap_tensor(gym.get_camera_image_gpu_tensor(sim, e, h, kind))
                             for e, h in zip(self.envs, self.camera_handles)]
        self.rgb_cam_tensors = wrap(gymapi.IMAGE_COLOR)               # [H,W,4] u8 each
        self.depth_cam_tensors = wrap(gymapi.IMAGE_DEPTH)             # [H,W] f32, negative z-depth, -inf = no hit
        self.seg_cam_tensors = wrap(gymapi.IMAGE_SEGMENTATION)        # [H,W] i32

    def set_state(self, poses):
        self.root_states[::self.skip, 0:3] = poses[..., 0:3] + self.env_origins.to(poses.device)
        self.root_states[::self.skip, 3:7] = quat_from_euler_xyz(poses[..., 3], poses[..., 4], poses[..., 5])
        self.gym.set_actor_root_state_tensor(self.sim, self.gymtorch.unwrap_tensor(self.root_states))

    def get_camera_view_matrix(self):
        return np.array([self.gym.get_camera_view_matrix(self.sim, e, h) for e, h in zip(self.envs, self.camera_handles)],
                        dtype=np.float32)

    def render(self, poses):
        gym, sim = self.gym, self.sim
        self.set_state(poses)
        gym.simulate(sim)
        gym.fetch_results(sim, True)
        gym.step_graphics(sim)
        gym.render_all_camera_sensors(sim)
        gym.start_access_image_tensors(sim)
        depth = torch.stack(self.depth_cam_tensors).float().contiguous()
        seg = torch.stack(self.seg_cam_tensors).to(torch.int32).contiguous()
        rgba = torch.stack(self.rgb_cam_tensors).to(torch.uint8).contiguous()
        gym.end_access_image_tensors(sim)
        contact = None
        if self.contact_forces is not None:                           # drone_robot.py check_termination: |F| > 1 on any body
            f = self.contact_forces
            contact = (torch.norm(f.view(len(self.envs), -1, 3), dim=-1) > self.contact_threshold).any(dim=1).to(torch.uint8)
        return SensorFrame(depth=depth, seg=seg, rgba=rgba, view_matrix=self.get_camera_view_matrix(), contact=contact)
