"""TEST HARNESS ONLY -- writes tests/golden/*.npz from the reference's own Python.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py
The fixtures are committed; the GPU box (no /root/reference) only reads them.

env_*.npz : a roll-out of the reference's unmodified `Env_Train_GenNBV` on CPU
    (oracle/ref_driver.py: Isaac Gym replaced by the synthetic renderer, PyCUDA Bresenham
    replaced by the C restatement), recording for reset() and every step():
      inputs : actions[T,N,6] i64 (as handed to step), raw sensor images depth[T+1,N,H,W] f32
               (negative z-depth, -inf = no hit), seg[T+1,N,H,W] i32, rgb[T+1,N,H,W,4] u8,
               view matrices [T+1,N,4,4] f32 (Isaac convention) and the c2w the reference
               derives from them (env_train_gennbv.py:512-514)
      outputs: prob_grid / scanned_gt_grid / tri grid after the step (i8 codes + the distinct
               fp32 values, see `_pack_f32`), coverage ratio, rewards, dones, time_outs,
               episode_length_buf, poses, the obs dict's "state" and "state_rgb".
    index 0 of the [T+1] arrays is reset().
"""
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
sys.path.insert(0, os.path.dirname(_HERE))

import ref_driver  # noqa: E402
from gennbv_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden")


def _pack_f32(a):
    """Grids hold a handful of distinct fp32 values: store the palette (as raw bits) + u8 codes."""
    bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    pal, codes = np.unique(bits, return_inverse=True)
    assert pal.size <= 256, pal.size
    return pal, codes.reshape(a.shape).astype(np.uint8)


def unpack_f32(pal, codes):
    return pal[codes].view(np.float32)


def rollout(name, N, H, W, G, S, T, max_episode_length, seed):
    torch.manual_seed(seed)
    scenes = synth.make_house_scenes(S, G, seed=seed)
    env, ref = ref_driver.make_reference_env(N, H, W, scenes, buffer_size=100,
                                             max_episode_length=max_episode_length)
    gen = torch.Generator().manual_seed(seed + 1)
    rec = {k: [] for k in ("depth", "seg", "rgb", "view", "c2w", "prob", "scan", "tri", "ratio", "rew", "done",
                           "time_out", "ep_len", "poses", "state", "state_rgb", "rgb_gray", "applied_actions")}
    actions = []

    # reset_idx writes init_pose_buf through pose_buf[-1], which aliases env.poses (env_train_gennbv.py:268-270,
    # 397-399) and overwrites env.actions (:409-411): record what the step actually used, before the reset
    used = {}
    inner_update = env.update_occ_grid

    def update_occ_grid_spy():
        used["poses"] = env.poses.numpy().copy()
        used["actions"] = env.actions.numpy().copy()
        inner_update()

    env.update_occ_grid = update_occ_grid_spy

    def snap(obs, rew, done):
        rec["depth"].append(torch.stack(env.depth_cam_tensors).numpy().copy())
        rec["seg"].append(torch.stack(env.seg_cam_tensors).numpy().copy())
        rec["rgb"].append(torch.stack(env.rgb_cam_tensors).numpy().copy())
        rec["view"].append(env._view_matrix.copy())
        rec["c2w"].append(synth.c2w_from_view_matrix(torch.from_numpy(env._view_matrix), env.env_origins).numpy())
        rec["prob"].append(env.prob_grid.numpy().copy())
        rec["scan"].append(env.scanned_gt_grid.numpy().copy())
        rec["tri"].append(obs["grid"].numpy().astype(np.int8))
        rec["ratio"].append(env.reward_ratio_buf[-1].numpy().copy())
        rec["rew"].append(rew.numpy().copy())
        rec["done"].append(done.numpy().copy())
        rec["time_out"].append(env.extras["time_outs"].numpy().copy())
        rec["ep_len"].append(env.episode_length_buf.numpy().copy())
        rec["poses"].append(used["poses"])
        rec["state"].append(obs["state"].numpy().copy())
        rec["state_rgb"].append(obs["state_rgb"].numpy().copy())
        rec["rgb_gray"].append(env.rgb_grayscale.numpy().copy())
        rec["applied_actions"].append(used["actions"])

    obs = env.reset()
    snap(obs, env.rew_buf, env.reset_buf.clone())
    for t in range(T):
        a = synth.sample_lookat_actions(scenes.params, N, gen)
        if t % 3 == 2:      # some out-of-range requests to exercise the clip (env_train_gennbv.py:247)
            a[0, 0], a[-1, 2], a[0, 4] = 95, -4, 20
        actions.append(a.numpy().copy())
        obs, rew, done, info = env.step(a)
        snap(obs, rew, done)
    out = {k: np.stack(v) for k, v in rec.items()}
    out["actions"] = np.stack(actions)
    for k in ("prob", "scan"):
        out[k + "_pal"], out[k + "_codes"] = _pack_f32(out.pop(k))
    out["grid_gt_file"] = np.packbits(scenes.grid_gt[..., 3].numpy().astype(bool), axis=None)
    out["grid_centres_lohi"] = np.stack([scenes.grid_gt[:, 0, 0, 0, :3].numpy(), scenes.grid_gt[:, -1, -1, -1, :3].numpy()], 1)
    out["scene_params"] = scenes.params.numpy()
    out["range_gt"] = env.range_gt.numpy()
    out["voxel_size_gt"] = env.voxel_size_gt.numpy()
    out["num_valid_voxel_gt"] = env.num_valid_voxel_gt.numpy()
    out["env_origins"] = env.env_origins.numpy()
    out["inv_intri"] = env.inv_intri.numpy()
    out["reward_scales"] = np.array([env.reward_scales["surface_coverage"], env.reward_scales["short_path"],
                                     env.reward_scales["termination"]], np.float64)
    out["meta"] = np.array([N, H, W, G, S, T, max_episode_length, seed], np.int64)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB; dones={int(out['done'][1:].sum())} "
          f"timeouts={int(out['time_out'][1:].sum())} final ratio={out['ratio'][-1]}")


def main():
    # reference-native grid (20^3), short episodes so time-outs / resets / forced init_action occur
    rollout("env_g20", N=4, H=48, W=48, G=20, S=3, T=14, max_episode_length=5, seed=0)
    # BASELINE config 1 shape: one env, 128x128 depth, 64^3 grid
    rollout("env_g64", N=1, H=128, W=128, G=64, S=1, T=4, max_episode_length=100, seed=3)
    # non-square image, longer episode (coverage saturates; `ratio > 0.99` termination if reached)
    rollout("env_g20_long", N=3, H=40, W=56, G=20, S=2, T=40, max_episode_length=100, seed=5)




def eval_rollout(name, N, H, W, G, S, T, max_episode_length, seed, n_gt, reset_at=()):
    """Roll-out of the reference's unmodified `Env_Eval_GenNBV` on CPU (pytorch3d's chamfer_distance replaced by the float64
    restatement in ref_driver, everything else the reference's own lines).  Records, besides the train fixture's inputs,
    the five-element returns, the per-env point-history sizes after every call and -- whenever the env computes an
    accuracy -- the deduplicated 1 cm cloud it handed to chamfer_distance."""
    torch.manual_seed(seed)
    scenes = synth.make_house_scenes(S, G, seed=seed)
    pc_gt = synth.gt_point_clouds(scenes.params, N, n_gt, seed=seed)
    env, ref = ref_driver.make_reference_env(N, H, W, scenes, buffer_size=100, max_episode_length=max_episode_length,
                                             eval_env=True, pc_gt=pc_gt)
    env_eval_mod = sys.modules["gennbv.env.env_eval_gennbv"]
    gen = torch.Generator().manual_seed(seed + 1)
    keys = ("depth", "seg", "rgb", "view", "tri", "ratio", "rew", "done", "ep_len", "hist_sizes", "acc", "is_reset")
    rec = {k: [] for k in keys}
    clouds = []                                              # (call index, dedup cloud) in the order chamfer is called
    inner = env_eval_mod.chamfer_distance

    def chamfer_spy(x, y):
        clouds.append((len(rec["rew"]), x[0].numpy().copy()))
        return inner(x, y)

    env_eval_mod.chamfer_distance = chamfer_spy

    def snap(out, is_reset):
        obs, rew, done, info, acc = out
        rec["depth"].append(torch.stack(env.depth_cam_tensors).numpy().copy())
        rec["seg"].append(torch.stack(env.seg_cam_tensors).numpy().copy())
        rec["rgb"].append(torch.stack(env.rgb_cam_tensors).numpy().copy())
        rec["view"].append(env._view_matrix.copy())
        rec["tri"].append(obs["grid"].numpy().astype(np.int8))
        rec["ratio"].append(env.reward_ratio_buf[-1].numpy().copy())
        rec["rew"].append(rew.numpy().copy())
        rec["done"].append(done.numpy().copy())
        rec["ep_len"].append(env.episode_length_buf.numpy().copy())
        rec["hist_sizes"].append(np.array([p.shape[0] for p in env.pts_target_list], np.int64))
        rec["acc"].append(np.array([acc.get(str(e), np.nan) for e in range(N)], np.float64))
        rec["is_reset"].append(np.array(is_reset))

    snap(env.reset(), True)
    actions = []
    for t in range(T):
        if t in reset_at:
            snap(env.reset(), True)
            actions.append(np.zeros((N, 6), np.int64))
            continue
        a = synth.sample_lookat_actions(scenes.params, N, gen)
        actions.append(a.numpy().copy())
        snap(env.step(a), False)
    out = {k: np.stack(v) for k, v in rec.items()}
    out["actions"] = np.stack(actions)
    out["cloud_call"] = np.array([c for c, _ in clouds], np.int64)
    out["cloud_sizes"] = np.array([p.shape[0] for _, p in clouds], np.int64)
    out["cloud_points"] = np.concatenate([p for _, p in clouds], 0) if clouds else np.zeros((0, 3), np.float32)
    out["pc_gt"] = np.stack([p.numpy() for p in pc_gt])
    out["range_gt"] = env.range_gt.numpy()
    out["voxel_size_gt"] = env.voxel_size_gt.numpy()
    out["num_valid_voxel_gt"] = env.num_valid_voxel_gt.numpy()
    out["env_origins"] = env.env_origins.numpy()
    out["inv_intri"] = env.inv_intri.numpy()
    out["reward_scale_cov"] = np.array(env.reward_scales["surface_coverage"], np.float64)
    out["grid_gt_file"] = np.packbits(scenes.grid_gt[..., 3].numpy().astype(bool), axis=None)
    out["grid_centres_lohi"] = np.stack([scenes.grid_gt[:, 0, 0, 0, :3].numpy(), scenes.grid_gt[:, -1, -1, -1, :3].numpy()], 1)
    out["scene_params"] = scenes.params.numpy()
    out["meta"] = np.array([N, H, W, G, S, T, max_episode_length, seed], np.int64)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    env_eval_mod.chamfer_distance = inner
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB; dones={int(out['done'].sum())} chamfer calls={len(clouds)} "
          f"cloud sizes={out['cloud_sizes'].tolist()} acc(last)={out['acc'][-1]}")


# ------------------------------------------------------------------------------------------------ policy goldens
def policy_golden(name="policy_g20", seed=7):
    """Outputs of the reference's own Hybrid_Encoder / ActorCriticPolicy_Train_Eval / PPO loss lines at the
    reference-native 20^3 grid, on observations taken from the env_g20_long roll-out, with generator-seeded weights
    (oracle/encoder_ref.py::seeded_state_dict -- reproducible anywhere without the reference)."""
    import ref_loader
    import encoder_ref
    ref = ref_loader.load_reference()
    from gym import spaces
    d = np.load(os.path.join(GOLDEN_DIR, "env_g20_long.npz"))
    N, G = int(d["meta"][0]), int(d["meta"][3])
    rows = []
    for t in (0, 3, 9, 20, 33, 40):
        rows.append(np.concatenate([d["state"][t].reshape(N, -1), d["tri"][t].reshape(N, -1).astype(np.float32),
                                    d["state_rgb"][t].reshape(N, -1)], 1))
    obs = torch.from_numpy(np.concatenate(rows, 0))                      # [18, 16792]
    B, D = obs.shape
    obs_space = spaces.Box(low=-np.inf, high=np.inf, shape=(D,), dtype=np.float32)
    act_space = spaces.MultiDiscrete([81, 81, 51, 1, 13, 13])
    kwargs = dict(encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
                  net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
                  state_input_shape=(600,), visual_input_shape=(100, 48, 48))
    policy = ref.policies.ActorCriticPolicy_Train_Eval(obs_space, act_space, lambda _: 1e-4, net_arch=[],
                                                       features_extractor_class=ref.encoder.Hybrid_Encoder,
                                                       features_extractor_kwargs=kwargs)
    mirror = encoder_ref.PolicyRef(G, 600)
    sd = encoder_ref.seeded_state_dict(mirror, seed)
    policy.load_state_dict(sd)
    g = torch.Generator().manual_seed(seed + 1)
    actions = torch.stack([torch.randint(0, n, (B,), generator=g) for n in (81, 81, 51, 1, 13, 13)], 1)
    out = {"obs_rows": np.array([0, 3, 9, 20, 33, 40]), "actions": actions.numpy(), "seed": np.array(seed)}
    policy.set_training_mode(False)
    with torch.no_grad():
        out["features_eval"] = policy.extract_features(obs).numpy()
        v, lp, ent = policy.evaluate_actions(obs, actions)
        out["values_eval"], out["log_prob_eval"], out["entropy_eval"] = v.numpy(), lp.numpy(), ent.numpy()
    policy.set_training_mode(True)
    old_v = torch.randn(B, generator=g) * 0.5
    old_lp = lp.detach() + 0.3 * torch.randn(B, generator=g)
    adv = torch.randn(B, generator=g) * 2 + 0.3
    ret = torch.randn(B, generator=g)
    v, lp, ent = policy.evaluate_actions(obs, actions)
    out["features_train"] = policy.extract_features(obs).detach().numpy()      # second BN update; recorded below after it
    loss, parts = encoder_ref.ppo_loss(v, lp, ent, old_v, old_lp, adv, ret)
    policy.optimizer.zero_grad()
    loss.backward()
    out.update(values_train=v.detach().numpy(), log_prob_train=lp.detach().numpy(), entropy_train=ent.detach().numpy(),
               old_values=old_v.numpy(), old_log_prob=old_lp.numpy(), advantages=adv.numpy(), returns=ret.numpy(),
               loss=loss.detach().numpy(), **{k: x.detach().numpy() for k, x in parts.items()})
    for k, p in policy.named_parameters():
        out["grad." + k] = p.grad.numpy().copy()      # clip_grad_norm_ below rescales .grad in place
    for k, b in policy.named_buffers():
        out["buf." + k] = b.detach().numpy().copy()
    total_norm = torch.nn.utils.clip_grad_norm_(policy.parameters(), 1.0)
    policy.optimizer.step()
    out["grad_norm"] = total_norm.numpy()
    for k, p in policy.named_parameters():
        out["new." + k] = p.detach().numpy().copy()
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    keep = {k: v for k, v in out.items() if not (k.startswith("grad.") or k.startswith("new.")) or v.size <= 4096}
    # large tensors: keep a strided sample + norm (fixtures stay small)
    for k, v in out.items():
        if k not in keep:
            flat = v.reshape(-1)
            keep[k + ".sample"] = flat[:: max(1, flat.size // 2048)][:2048].copy()
            keep[k + ".norm"] = np.array(np.linalg.norm(flat.astype(np.float64)))
    np.savez_compressed(path, **keep)
    print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB; loss={float(loss):.6f} grad_norm={float(total_norm):.4f}")


if __name__ == "__main__":
    if "--eval" in sys.argv:
        # two 4-step episodes, a reset() in the middle of the third (accuracy of partial histories), then more steps
        eval_rollout("env_eval_g20", N=3, H=40, W=40, G=20, S=2, T=13, max_episode_length=4, seed=11, n_gt=1500, reset_at=(10,))
        sys.exit(0)
    if "--policy" not in sys.argv:
        main()
    policy_golden()
