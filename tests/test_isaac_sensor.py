"""CPU: `IsaacGymSensor` against a mocked Isaac Gym -- the call sequence of env_train_gennbv.py:255-261,349-354 and
env_train_base.py:686-714,777-785, and the frame it hands to the env."""
import numpy as np
import torch

from gennbv_b200.sensors import IsaacGymSensor, SensorFrame, quat_from_euler_xyz


class FakeGymApi:
    IMAGE_COLOR, IMAGE_DEPTH, IMAGE_SEGMENTATION = 0, 1, 2


class FakeGymTorch:
    @staticmethod
    def wrap_tensor(x):
        return x

    @staticmethod
    def unwrap_tensor(x):
        return ("unwrapped", x)


class FakeGym:
    def __init__(self, n, h, w):
        self.calls, self.n = [], n
        g = torch.Generator().manual_seed(0)
        self.images = {0: [torch.randint(0, 255, (h, w, 4), generator=g).to(torch.uint8) for _ in range(n)],
                       1: [-torch.rand(h, w, generator=g) * 9 for _ in range(n)],
                       2: [torch.randint(0, 255, (h, w), generator=g).to(torch.int32) for _ in range(n)]}
        self.view = [np.eye(4, dtype=np.float32) * (k + 1) for k in range(n)]
        self.root = None

    def get_camera_image_gpu_tensor(self, sim, env, handle, kind):
        assert env == ("env", handle)
        return self.images[kind][handle]

    def set_actor_root_state_tensor(self, sim, t):
        assert t[0] == "unwrapped"
        self.root = t[1].clone()
        self.calls.append("set_actor_root_state_tensor")

    def get_camera_view_matrix(self, sim, env, handle):
        return self.view[handle]

    def __getattr__(self, name):
        def f(*a):
            self.calls.append(name)
        return f


def test_quat_from_euler_matches_scipy():
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(1)
    e = (torch.rand(50, 3, generator=g) - 0.5) * 6
    q = quat_from_euler_xyz(e[:, 0], e[:, 1], e[:, 2]).numpy()
    want = Rotation.from_euler("xyz", e.numpy()).as_quat()         # extrinsic xyz, (x, y, z, w)
    flip = np.sign((q * want).sum(1, keepdims=True))
    np.testing.assert_allclose(q * flip, want, atol=1e-6)


def test_render_follows_the_reference_call_sequence():
    n, h, w = 3, 8, 10
    gym = FakeGym(n, h, w)
    root = torch.zeros(n, 13)
    origins = torch.tensor([[0., 0, 0], [5, 0, 0], [0, 5, 0]])
    s = IsaacGymSensor(gym, "sim", [("env", k) for k in range(n)], list(range(n)), root, origins, h, w, FakeGymApi, FakeGymTorch)
    poses = torch.tensor([[1., 2, 3, 0, 0.1, 0.2], [0, 0, 1, 0, -0.3, 1.0], [-1, -2, 0.5, 0, 0, 3.0]])
    f = s.render(poses)
    assert isinstance(f, SensorFrame)
    assert gym.calls == ["set_actor_root_state_tensor", "simulate", "fetch_results", "step_graphics", "render_all_camera_sensors",
                         "start_access_image_tensors", "end_access_image_tensors"]
    np.testing.assert_allclose(gym.root[:, 0:3].numpy(), (poses[:, :3] + origins).numpy())
    np.testing.assert_allclose(gym.root[:, 3:7].numpy(), quat_from_euler_xyz(poses[:, 3], poses[:, 4], poses[:, 5]).numpy())
    assert f.depth.shape == (n, h, w) and f.depth.dtype == torch.float32 and torch.equal(f.depth[1], gym.images[1][1])
    assert f.seg.dtype == torch.int32 and torch.equal(f.seg[2], gym.images[2][2])
    assert f.rgba.shape == (n, h, w, 4) and f.rgba.dtype == torch.uint8 and torch.equal(f.rgba[0], gym.images[0][0])
    assert f.c2w is None and f.view_matrix.shape == (n, 4, 4) and f.view_matrix[2][0, 0] == 3
    # the frame is a COPY: Isaac may overwrite its image tensors after end_access_image_tensors
    gym.images[1][1].zero_()
    assert float(f.depth[1].abs().sum()) > 0
