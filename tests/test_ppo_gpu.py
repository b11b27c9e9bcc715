"""GPU: rollout buffer, collect_rollouts glue and the fused PPO update against a plain-PyTorch re-run of the
reference's train() lines (oracle/encoder_ref.py) with identical weights and data."""
import numpy as np
import pytest
import torch

import encoder_ref
import oracle as c_oracle
from gennbv_b200.buffers import TensorRolloutBuffer_Grid_Obs
from gennbv_b200.ppo import PPO_Grid_Obs
from gennbv_b200.spaces import Box, MultiDiscrete
from gennbv_b200.wrapper import EnvWrapperGenNBVTrain
from helpers import EnvGolden
from test_env_gpu import make_env

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
POLICY_KW = lambda: dict(net_arch=[], features_extractor_kwargs=dict(
    encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
    net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
    state_input_shape=(600,), visual_input_shape=(100, 48, 48)))


def make_algo(env, n_steps, batch_size, n_epochs, target_kl=None, seed=3, semantic=False):
    kw = POLICY_KW()
    if semantic:
        kw["features_extractor_kwargs"]["semantic_branch"] = True
    algo = PPO_Grid_Obs(env=env, learning_rate=1e-4, n_steps=n_steps, batch_size=batch_size, n_epochs=n_epochs, gamma=0.99,
                        gae_lambda=0.95, clip_range=0.2, clip_range_vf=0.2, ent_coef=0.01, vf_coef=0.8, max_grad_norm=1,
                        target_kl=target_kl, policy_kwargs=kw, seed=seed, device=DEV)
    ref = encoder_ref.PolicyRef(env.grid_size, 600, semantic=semantic)
    sd = encoder_ref.seeded_state_dict(ref, seed, scale=0.5)
    if semantic:
        sd["features_extractor.naive_encoder_rgb.0.weight"] /= 255.0        # frames are 0..255 gray levels
    ref.load_state_dict(sd)
    algo.policy.load_state_dict(sd)
    return algo, ref


def test_collect_rollouts_fills_buffer_like_the_reference_loop():
    g = EnvGolden("env_g20_long")
    env = EnvWrapperGenNBVTrain(make_env(g))
    T = 12
    algo, ref = make_algo(env, n_steps=T, batch_size=9, n_epochs=1)
    algo._setup_learn()
    first_obs = algo._last_obs.clone()
    seen = []

    def spy(loc):
        seen.append(dict(obs=loc["new_obs"].clone(), rew=loc["rewards"].clone(), done=loc["dones"].clone(),
                         tout=loc["infos"]["time_outs"].clone(), actions=loc["actions"].clone(), values=loc["values"].clone()))

    assert algo.collect_rollouts(callback=spy)
    buf = algo.rollout_buffer
    N = g.N
    assert buf.full and algo.num_timesteps == T * N
    np.testing.assert_array_equal(buf.observations[0].cpu().numpy(), first_obs.cpu().numpy())
    for t in range(1, T):          # obs stored at t is the obs returned by step t-1 (the ping-pong buffers keep it alive)
        np.testing.assert_array_equal(buf.observations[t].cpu().numpy(), seen[t - 1]["obs"].cpu().numpy())
        np.testing.assert_array_equal(buf.episode_starts[t, :, 0].cpu().numpy(), seen[t - 1]["done"].cpu().numpy())
    assert buf.episode_starts[0].all()
    ref.eval()
    for t in range(T):
        np.testing.assert_array_equal(buf.actions[t].cpu().numpy(), seen[t]["actions"].float().cpu().numpy())
        # time-out bootstrap: rewards += gamma * V(new_obs)[env 0] * time_outs  (on_policy_algorithm_grid_obs.py:205-208)
        with torch.no_grad():
            v0 = float(ref.value_net(ref.features_extractor(seen[t]["obs"].cpu()))[0, 0])
        want = seen[t]["rew"].cpu() + 0.99 * v0 * seen[t]["tout"].cpu().float()
        np.testing.assert_allclose(buf.rewards[t, :, 0].cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)
    # GAE of the stored columns (bit-exact against the oracle)
    with torch.no_grad():
        last_v = algo.policy.predict_values(seen[-1]["obs"]).flatten().cpu().numpy()
    adv, ret = c_oracle.gae(buf.rewards[..., 0].cpu().numpy(), buf.values[..., 0].cpu().numpy(),
                            buf.episode_starts[..., 0].cpu().numpy(), last_v, seen[-1]["done"].cpu().numpy().astype(np.uint8))
    np.testing.assert_array_equal(buf.advantages[..., 0].cpu().numpy(), adv)
    np.testing.assert_array_equal(buf.returns[..., 0].cpu().numpy(), ret)
    # get(): the reference's env-major flat index i = n*T + t under one permutation per rollout
    batch = next(buf.get(9))
    i = buf.indices[:9]
    n, t = i // T, i % T
    np.testing.assert_array_equal(batch.observations.cpu().numpy(), buf.observations[t, n].cpu().numpy())
    np.testing.assert_array_equal(batch.advantages.cpu().numpy(), buf.advantages[t, n, 0].cpu().numpy())


def test_rollout_observations_are_born_in_the_buffer_slots():
    """with `bind_rollout_slots` the env writes step t's observation straight into buffer slot t+1 (add() then
    skips its copy); the buffer content is identical to the copying path."""
    g = EnvGolden("env_g20_long")
    T = 10
    bufs = []
    for bind in (True, False):
        env = EnvWrapperGenNBVTrain(make_env(g))
        algo, _ = make_algo(env, n_steps=T, batch_size=9, n_epochs=1)
        algo.bind_rollout_slots = bind
        algo._setup_learn()
        ptrs = []
        assert algo.collect_rollouts(callback=lambda loc: ptrs.append(loc["new_obs"].data_ptr()))
        buf = algo.rollout_buffer
        slots = [buf.observations[t].data_ptr() for t in range(T)]
        if bind:
            assert ptrs[:T - 1] == slots[1:], "observations of steps 0..T-2 must live in slots 1..T-1"
            assert ptrs[T - 1] not in slots               # the last one stays in the env's own buffer for the next rollout
        else:
            assert not set(ptrs) & set(slots)
        bufs.append({k: getattr(buf, k).clone() for k in ("observations", "actions", "rewards", "episode_starts", "values",
                                                          "log_probs", "advantages", "returns")})
    for k in bufs[0]:
        assert torch.equal(bufs[0][k], bufs[1][k]), k


def _torch_rerun(ref, buf, N, T, B, E, target_kl):
    """Plain torch re-run of ppo_grid_obs.py:176-297 on the CPU with the same buffer content and permutation -- the loop that
    tests/test_ppo_ref_vs_reference.py pins bit for bit against the reference's own collect_rollouts() + train()."""
    flat = lambda x: x.transpose(0, 1).reshape(N * T, *x.shape[2:]).cpu()          # swap_and_flatten (buffers.py:56-69)
    obs, acts = flat(buf.observations), flat(buf.actions).long()
    vals, lps, advs, rets = (flat(x).flatten() for x in (buf.values, buf.log_probs, buf.advantages, buf.returns))
    opt = torch.optim.Adam(ref.parameters(), lr=1e-4, eps=1e-5)
    ref.train()
    logs, stop, steps, last_epoch_kl = [], False, 0, []
    for epoch in range(E):
        last_epoch_kl = []
        for start in range(0, N * T, B):
            idx = buf.indices[start:start + B]
            v, lp, ent = ref.evaluate_actions(obs[idx], acts[idx])
            loss, parts = encoder_ref.ppo_loss(v, lp, ent, vals[idx], lps[idx], advs[idx], rets[idx])
            logs.append([float(loss.detach())] + [float(parts[k].detach()) for k in
                                                  ("policy_loss", "value_loss", "entropy_loss", "approx_kl", "clip_fraction")])
            last_epoch_kl.append(logs[-1][4])
            if target_kl is not None and float(parts["approx_kl"]) > 1.5 * target_kl:
                stop = True
                break
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
            opt.step()
            steps += 1
        if stop:
            break
    return np.array(logs), steps, float(np.mean(last_epoch_kl)), stop


def _check_train_against_rerun(algo, ref, before, logs, steps, last_kl):
    rec = algo.logger.name_to_value
    close = lambda a, b: abs(a - b) <= 1e-4 * max(1.0, abs(b))           # BASELINE.json: <= 1e-4 relative on the loss terms
    for key, col in (("train/policy_gradient_loss", 1), ("train/value_loss", 2), ("train/entropy_loss", 3), ("train/clip_fraction", 5)):
        assert close(rec[key], logs[:, col].mean()), (key, rec[key], logs[:, col].mean())
    assert close(rec["train/approx_kl"], last_kl), (rec["train/approx_kl"], last_kl)
    assert close(rec["train/loss"], logs[-1, 0])
    assert algo._last_train["minibatches_logged"] == len(logs)
    assert algo._adam_step == steps
    sd_ref = ref.state_dict()
    for k, v in algo.policy.state_dict().items():
        a, b = v.detach().cpu().double(), sd_ref[k].double()
        if k.endswith("num_batches_tracked"):
            assert int(a) == int(b), k
            continue
        # every tensor of the updated policy within 1e-4 of the tensor's scale ...
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-12, k
        if k.endswith("naive_encoder_grid.0.bias") or k.endswith("naive_encoder_grid.3.bias") or "running" in k or steps == 0:
            continue          # zero-gradient parameters (conv bias under batch-stat BN): Adam amplifies rounding noise
        # ... and the DISPLACEMENT (what the update actually did, ~lr per step) in the 2-norm.  Error analysis: an Adam
        # step is lr * m / (sqrt(v) + eps); in the first steps |m| / sqrt(v) ~ 1 regardless of |g|, so an element whose
        # gradient is a cancellation residue (|g| below the fp32 rounding of its 10^4..10^5 summed terms) can move by up to
        # 2 lr in either implementation.  Those elements carry ~0 of the update's norm, so the norm of the displacement
        # difference stays a small fraction of the displacement norm even though single elements may differ by O(lr).
        da, db = a - before[k].cpu().double(), b - before[k].cpu().double()
        assert float((da - db).norm()) <= 0.02 * float(db.norm()) + 1e-9, (k, float((da - db).norm()), float(db.norm()))


@pytest.mark.parametrize("target_kl,graph", [(None, True), (None, False), (1e-7, True), ("mid", True), ("mid", False)])
def test_fused_train_matches_torch_rerun(target_kl, graph):
    g = EnvGolden("env_g20_long")
    env = EnvWrapperGenNBVTrain(make_env(g))
    T, B, E = 8, 12, 2
    N = g.N
    if target_kl == "mid":
        # a threshold the run crosses in the MIDDLE of an epoch: take it from an un-stopped re-run's 4th minibatch
        probe_algo, probe_ref = make_algo(env, n_steps=T, batch_size=B, n_epochs=E)
        probe_algo._setup_learn()
        probe_algo.collect_rollouts()
        kl = _torch_rerun(probe_ref, probe_algo.rollout_buffer, N, T, B, E, None)[0][:, 4]
        n_mb = -(-N * T // B)
        cand = [j for j in range(1, len(kl)) if kl[j] > kl[:j].max()]
        if not cand:
            pytest.skip("approx_kl never exceeds its running maximum after the first minibatch in this run")
        k = next((j for j in cand if j >= 2), cand[0])
        target_kl = float(0.5 * (kl[:k].max() + kl[k]) / 1.5)
        env = EnvWrapperGenNBVTrain(make_env(g))
    algo, ref = make_algo(env, n_steps=T, batch_size=B, n_epochs=E, target_kl=target_kl)
    algo.use_cuda_graph = graph
    algo._setup_learn()
    algo.collect_rollouts()
    buf = algo.rollout_buffer
    logs, steps, last_kl, stop = _torch_rerun(ref, buf, N, T, B, E, target_kl)
    assert stop == (target_kl is not None)
    before = {k: v.clone() for k, v in algo.policy.state_dict().items()}
    algo.train()
    assert (algo._last_train["stopped_epoch"] is not None) == stop
    _check_train_against_rerun(algo, ref, before, logs, steps, last_kl)
    if graph:
        assert len(algo._graphs) >= 1, "the minibatch update was expected to run as a captured CUDA graph"


def test_fused_train_with_the_semantic_branch_matches_torch_rerun():
    """the whole update (rollout, captured minibatch graph, clip + Adam over the 26-tensor arena) with the 2-D
    branch switched on, against the plain-PyTorch re-run of the reference's train() lines over the same three-branch network."""
    g = EnvGolden("env_g20_long")
    env = EnvWrapperGenNBVTrain(make_env(g))
    T, B, E = 8, 12, 2
    algo, ref = make_algo(env, n_steps=T, batch_size=B, n_epochs=E, semantic=True)
    assert len(algo.policy.arena_parameters()) == 26
    algo._setup_learn()
    algo.collect_rollouts()
    logs, steps, last_kl, stop = _torch_rerun(ref, algo.rollout_buffer, g.N, T, B, E, None)
    before = {k: v.clone() for k, v in algo.policy.state_dict().items()}
    algo.train()
    _check_train_against_rerun(algo, ref, before, logs, steps, last_kl)
    assert len(algo._graphs) >= 1


def test_graph_and_eager_updates_are_bit_identical():
    """The captured CUDA graph replays exactly the launches of the eager path: same parameters, moments and log, bit for bit."""
    g = EnvGolden("env_g20_long")
    out = []
    for graph in (True, False):
        env = EnvWrapperGenNBVTrain(make_env(g))
        algo, _ = make_algo(env, n_steps=8, batch_size=12, n_epochs=2, target_kl=None)
        algo.use_cuda_graph = graph
        algo._setup_learn()
        algo.collect_rollouts()
        algo.train()
        out.append((algo.policy.flat_params.clone(), algo._exp_avg.clone(), algo._exp_avg_sq.clone(),
                    torch.from_numpy(algo._last_train["scalars"]), [b.clone() for b in algo.policy.buffers()]))
    for a, b in zip(out[0][:4], out[1][:4]):
        assert torch.equal(a.cpu(), b.cpu())
    for a, b in zip(out[0][4], out[1][4]):
        assert torch.equal(a, b)


def test_golden_ppo_train_replay():
    """tests/golden/ppo_train_g20.npz: a rollout buffer filled by the REFERENCE's collect_rollouts(), its permutation, and what
    the reference's train() logged / left in the policy (oracle/ref_ppo_driver.py).  The fused update replays it."""
    import ppo_replay
    ppo_replay.replay_ppo_golden(DEV)
