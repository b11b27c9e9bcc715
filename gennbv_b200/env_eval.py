"""`Env_Eval_GenNBV` -- drop-in for the reference eval task env (gennbv/env/env_eval_gennbv.py, registered as
"eval_gennbv", gennbv/__init__.py:7): the training env plus the reconstruction-accuracy bookkeeping.

Differences from `Env_Train_GenNBV`, all taken from the reference class:
  * `reset()` and `step()` return a fifth element, the dict `ratios_accuracy` {str(env_idx): chamfer accuracy}
    (env_eval_gennbv.py:107-154);
  * every step's foreground world points are appended to a per-env history (`pts_target_list`, :160-164); when an env
    finishes an episode its history is reduced to the distinct 1 cm lattice points and compared with the env's GT point
    cloud by the chamfer distance (:250-264), once per env until the next `reset()`;
  * termination is collision or time-out only, OR-ed into `reset_buf` (:327-333), so a `reset()` reports every env as
    done; there is no coverage-ratio termination;
  * the config is `Config_GenNBV_Eval` (30-step episodes, reward = coverage gain, clipped at 0).

B200-first: the history holds packed 8-byte lattice keys written by `scan_points_kernel` (one launch per step for all
envs, capacity (max_episode_length + 1) * H * W per env, so it cannot overflow); at episode end the finished envs'
keys are tagged with an env rank in their spare high bits and deduplicated by ONE sort (`torch.unique` on int64:
plumbing), decoded by one `keys_to_points_kernel` launch, and all finished envs go through ONE batched exact-1-NN chamfer
launch group (`gennbv_b200.chamfer`, grid search for large clouds).
`pytorch3d.loss.chamfer_distance` is not vendored by the reference: parity of the accuracy value is unpinned
(SURVEY.md 8c); the point history itself is pinned bit for bit by tests/golden/env_eval_*.npz.
"""
import numpy as np
import torch

from . import _lib, chamfer, ops
from .config import Config_GenNBV_Eval
from .env import Env_Train_GenNBV


class Env_Eval_GenNBV(Env_Train_GenNBV):
    num_scene = 50                                              # env_eval_gennbv.py:16

    def __init__(self, cfg=None, sim_params=None, physics_engine=None, sim_device="cuda:0", headless=True, *,
                 sensor=None, grid_gt=None, pc_gt=None, num_envs=None):
        """pc_gt: list of [n_gt, 3] float tensors, entry e = GT surface cloud of env e (the reference loads
        data_gennbv/eval/gt/point_cloud/BAT12_SETA_HOUSE{e+1}_pc.pt, env_eval_gennbv.py:93-101)."""
        super().__init__(cfg if cfg is not None else Config_GenNBV_Eval(), sim_params, physics_engine, sim_device, headless,
                         sensor=sensor, grid_gt=grid_gt, num_envs=num_envs)
        if pc_gt is None or len(pc_gt) != self.num_envs:
            raise ValueError("Env_Eval_GenNBV needs one GT point cloud per env (pc_gt)")
        dev, N = self.device, self.num_envs
        self.pc_gt = [p.reshape(-1, 3).float().contiguous().to(dev) for p in pc_gt]
        self.ratio_threshold_term = float("inf")                # check_termination has no coverage threshold (:327-333)
        self._accumulate_reset = True                           # `reset_buf |= ...`
        self.ratios_accuracy = dict()                           # _init_buffers (:103-105)
        self._pts_capacity = (int(self.max_episode_length) + 1) * self.H * self.W
        self._pts_keys = torch.empty(N, self._pts_capacity, dtype=torch.int64, device=dev)
        self._pts_count = torch.zeros(N, dtype=torch.int32, device=dev)
        self._pts_overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self._len_sum_before = None
        self._gt_pack = None

    # ------------------------------------------------------------------ reference surface
    @property
    def pts_target_list(self):
        """The reference attribute (list of [n,3] f32 per env), reconstructed from the key history: the rounded points in
        append order are not kept, so this returns the *rounded* history (duplicates included) -- debugging aid."""
        counts = self._pts_count.tolist()
        return [self._decode(self._pts_keys[e, :c]) for e, c in enumerate(counts)]

    def reset(self, return_all=True):
        """env_eval_gennbv.py:107-132."""
        self._reset_u8.fill_(1)
        self._time_outs_extra_u8.copy_(self._time_out_u8)
        self._stats[204:207] = self.episode_sums_buf.double().mean(dim=1) / self.max_episode_length_s
        self._pts_count.zero_()                                 # reset_idx clears every env's list (:308-311)
        self._reset_flagged(clear=False)
        self.actions.copy_(torch.clip(self.actions, self.clip_pose_idx_low, self.clip_pose_idx_up))
        self.poses.copy_(self.get_pose_from_discrete_action(self.actions))
        obs, rew, dones, infos = self.post_physics_step()
        accuracies = self.ratios_accuracy
        self.ratios_accuracy = dict()                           # :126
        return (obs, rew, dones, infos, accuracies) if return_all else obs

    def step(self, actions):
        """env_eval_gennbv.py:134-154 (accepts numpy / lists like the reference)."""
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
            if actions.dim() == 1:
                actions = actions.unsqueeze(0)
        actions = actions.to(self.device, torch.int64)
        obs, rew, dones, infos = super().step(actions)
        return obs, rew, dones, infos, self.ratios_accuracy

    # ------------------------------------------------------------------ hooks called by Env_Train_GenNBV.post_physics_step
    def _after_occ_grid_update(self, frame, c2w):
        """update_occ_grid's history append (:160-164) for all envs, one launch."""
        # `cur_episode_length.sum() > 0` (:251) is evaluated before update_extra_episode_info runs: capture it here,
        # ahead of the reward / statistics kernel that advances the counters
        self._len_sum_before = self.cur_episode_length.sum()
        _lib.check(_lib.lib().gnbv_scan_points(
            ops._ptr(frame.depth, torch.float32, "depth", (self.num_envs, self.H, self.W)),
            ops._ptr(frame.seg, torch.int32, "seg", (self.num_envs, self.H, self.W)),
            self.inv_intri.data_ptr(), ops._ptr(c2w, torch.float32, "c2w", (self.num_envs, 4, 4)),
            self._pts_keys.data_ptr(), self._pts_count.data_ptr(), self._pts_overflow.data_ptr(), self.num_envs, self.H, self.W,
            self._pts_capacity, ops.GNBV_RAW_DEPTH, ops._stream()), "gnbv_scan_points")

    def _decode(self, keys):
        pts = torch.empty(keys.shape[0], 3, device=self.device)
        _lib.check(_lib.lib().gnbv_keys_to_points(keys.data_ptr(), keys.shape[0], pts.data_ptr(), ops._stream()),
                   "gnbv_keys_to_points")
        return pts

    def scanned_cloud(self, env_idx):
        """`torch.unique(torch.round(pts_target_list[env_idx], decimals=2), dim=0)` (:254-257): the distinct 1 cm lattice
        points scanned so far, in the reference's (lexicographic) row order."""
        c = int(self._pts_count[env_idx])
        return self._decode(ops.sort_unique(self._pts_keys[env_idx, :c].clone(), key_bits=self.ENV_KEY_SHIFT))

    ENV_KEY_SHIFT = 54                      # history keys use 54 bits (eval_points.cu); bits 54..62 carry an env rank
    ENVS_PER_SORT = 512

    def dedup_clouds(self, env_ids):
        """`torch.unique(torch.round(pts_target_list[e], decimals=2), dim=0)` (:254-257) for many envs at once: the valid key
        prefixes are gathered into one array, tagged with the env's rank in the spare high bits (gnbv_pack_env_keys),
        de-duplicated by ONE in-tree radix sort + compaction (gnbv_sort_unique_u64) and decoded by one kernel launch.
        Returns (points [sum n, 3] f32 packed in env order, sizes list)."""
        dev = self.device
        pts_parts, sizes = [], []
        for c0 in range(0, len(env_ids), self.ENVS_PER_SORT):
            ids = env_ids[c0:c0 + self.ENVS_PER_SORT]
            rows = torch.tensor(ids, device=dev)
            cnt = self._pts_count[rows].long()
            maxc = int(cnt.max())
            if maxc == 0:
                sizes += [0] * len(ids)
                continue
            offs = torch.zeros(len(ids) + 1, dtype=torch.int64, device=dev)
            offs[1:] = torch.cumsum(cnt, 0)
            total = int(offs[-1])
            packed = torch.empty(total, dtype=torch.int64, device=dev)
            _lib.check(_lib.lib().gnbv_pack_env_keys(self._pts_keys.data_ptr(), self._pts_keys.shape[1], rows.data_ptr(), offs.data_ptr(),
                                                     len(ids), self.ENV_KEY_SHIFT, packed.data_ptr(), ops._stream()), "gnbv_pack_env_keys")
            nbits = self.ENV_KEY_SHIFT + max(1, (len(ids) - 1).bit_length())
            uniq = ops.sort_unique(packed, key_bits=nbits)                                       # sorted: env rank major, key minor
            bounds = torch.searchsorted(uniq, torch.arange(len(ids) + 1, device=dev, dtype=torch.int64) << self.ENV_KEY_SHIFT)
            sizes += (bounds[1:] - bounds[:-1]).tolist()
            pts_parts.append(self._decode(uniq))
        pts = torch.cat(pts_parts, 0) if pts_parts else torch.zeros(0, 3, device=dev)
        return pts, sizes

    def _before_reset_idx(self):
        """Accuracy of the envs that finish on this step (:250-264), then reset_idx's `pts_target_list[env] = empty`."""
        dones = self._dones_u8
        done_ids = dones.nonzero().flatten().tolist()           # host read, as the reference's `.item()` per env
        if not done_ids:
            return
        if int(self._pts_overflow) != 0:
            raise RuntimeError("Env_Eval_GenNBV: point history overflow (an episode ran past max_episode_length + 1 steps)")
        if float(self._len_sum_before) > 0:
            todo = [e for e in done_ids if str(e) not in self.ratios_accuracy]
            if todo:
                pts, sizes = self.dedup_clouds(todo)
                ok = [i for i, n in enumerate(sizes) if n > 0]
                acc = [float("nan")] * len(todo)                 # an env that never saw the object has no scanned cloud
                if ok:
                    gt_key = tuple(todo[i] for i in ok)
                    if self._gt_pack is None or self._gt_pack[0] != gt_key:     # GT clouds packed once per set of envs
                        gts = [self.pc_gt[e] for e in gt_key]
                        self._gt_pack = (gt_key, torch.cat(gts, 0).contiguous(), [int(g.shape[0]) for g in gts])
                    cx, cy = chamfer.chamfer_terms_packed(pts, [sizes[i] for i in ok], self._gt_pack[1], self._gt_pack[2])
                    # `(chamfer_distance(...) * 100)[0]` (:258-259) repeats the returned tuple 100 times and takes its
                    # first element: the unscaled loss -- reproduced
                    for i, v in zip(ok, (cx + cy).tolist()):
                        acc[i] = v
                for e, v in zip(todo, acc):
                    self.ratios_accuracy[str(e)] = v
        self._pts_count.masked_fill_(dones.view(torch.bool), 0)
