"""`gennbv_b200.evaluation`: AUC_update against the reference's own function (CPU, needs /root/reference), and the whole
evaluation loop on the GPU eval env against a per-env Python restatement of the reference loop
(stable_baselines3/common/evaluation.py:262-335) fed with the same step outputs."""
import numpy as np
import pytest
import torch

from gennbv_b200.evaluation import AUC_update, evaluate_policy_grid_obs


@pytest.mark.reference
def test_auc_update_equals_reference_function():
    import ref_loader
    ref_loader.load_reference()
    from stable_baselines3.common.evaluation import AUC_update as ref_update
    g = torch.Generator().manual_seed(0)
    n, L = 7, 9
    mine, theirs = torch.zeros(n, L), torch.zeros(n, L)
    flag = torch.zeros(n)
    for step in range(1, L + 1):
        rew = torch.rand(n, generator=g)
        dones = (torch.rand(n, generator=g) < 0.25)
        mine = AUC_update(mine.clone(), rew.clone(), step, dones, flag)
        theirs = ref_update(theirs.clone(), rew.clone(), step, dones, flag)
        assert torch.equal(mine, theirs), step
        flag = flag + dones.float()


class _FixedPolicy:
    """predict() replays a fixed action table (what a deterministic policy would output)."""

    def __init__(self, actions):
        self.actions, self.t = actions, 0

    def predict(self, obs, state=None, episode_start=None, deterministic=True):
        a = self.actions[self.t]
        self.t += 1
        return a, state


@pytest.mark.gpu
def test_evaluation_loop_matches_reference_bookkeeping():
    from gennbv_b200 import synth
    from gennbv_b200.config import Config_GenNBV_Eval
    from gennbv_b200.env_eval import Env_Eval_GenNBV
    from gennbv_b200.sensors import SyntheticHouseSensor
    from gennbv_b200.wrapper import EnvWrapperGenNBVEval
    DEV = "cuda:0"
    N, H, W, G, S, L = 5, 48, 48, 20, 2, 6

    class Cfg(Config_GenNBV_Eval):
        max_episode_length = L

    scenes = synth.make_house_scenes(S, G, seed=2)
    pc_gt = synth.gt_point_clouds(scenes.params, N, 2000, seed=2)
    env = Env_Eval_GenNBV(Cfg(), sim_device=DEV, sensor=SyntheticHouseSensor(scenes.params, H, W, device=DEV),
                          grid_gt=scenes.grid_gt, pc_gt=pc_gt, num_envs=N)
    gen = torch.Generator().manual_seed(3)
    table = [synth.sample_lookat_actions(scenes.params, N, gen).to(DEV) for _ in range(L + 2)]
    # record what the env returns by running the same loop by hand first
    wrapped = EnvWrapperGenNBVEval(env)
    wrapped.reset()
    rews, dones = [], []
    for t in range(L):
        _, r, d, _, acc = wrapped.step(table[t])
        rews.append(r.cpu().clone()); dones.append(d.cpu().clone())
    want_acc = dict(acc)
    assert all(bool(d.all()) == (t == L - 1) for t, d in enumerate(dones))
    # reference bookkeeping, per env (evaluation.py:262-335)
    auc = torch.zeros(N, L)
    flag = torch.zeros(N)
    cur = torch.zeros(N)
    ep_rewards = []
    for t in range(L):
        for e in range(N):
            if flag[e]:
                auc[e, t] = auc[e, t - 1]
            elif dones[t][e] == 0:
                auc[e, t] = rews[t][e]
        cur += rews[t]
        for e in range(N):
            flag[e] += float(dones[t][e])
            if dones[t][e]:
                ep_rewards.append(cur[e].clone()); cur[e] = 0
    want_auc = sum(auc[:, i] * (L - i) for i in range(L)) / L
    # the loop under test, on a fresh pass over the same actions (the env is deterministic)
    er, el, mean_auc, accs = evaluate_policy_grid_obs(_FixedPolicy(table), wrapped, n_eval_episodes=N, deterministic=True)
    assert len(er) == N and [int(x) for x in el] == [L] * N
    np.testing.assert_allclose(torch.stack(er).numpy(), torch.stack(ep_rewards).numpy(), rtol=1e-6)
    np.testing.assert_allclose(mean_auc.cpu().numpy(), want_auc.numpy(), rtol=1e-6)
    assert sorted(accs) == sorted(want_acc.values()) and all(np.isfinite(a) and a > 0 for a in accs)


class _FakeEvalEnv:
    """Pure-torch stand-in for the wrapped eval env (50 envs x 30 steps, the constants the reference hard-codes): scripted
    rewards, a time-out at step 30 for every env plus a few scripted early terminations, accuracies filled in on done."""
    num_envs, max_episode_length, device = 50, 30, torch.device("cpu")

    def __init__(self, seed):
        g = torch.Generator().manual_seed(seed)
        self.rew = torch.rand(40, 50, generator=g)
        self.early = {5: [3, 17], 12: [8], 29: [44]}           # step -> envs that terminate early (collisions)
        self.t = 0
        self.len = torch.zeros(50, dtype=torch.long)
        self.acc = {}

    def env_is_wrapped(self, cls):
        return [False]

    def reset(self):
        self.t, self.acc = 0, {}
        self.len.zero_()
        return torch.zeros(50, 4), torch.zeros(50), torch.ones(50, dtype=torch.bool), {}, {}

    def step(self, actions):
        self.len += 1
        dones = self.len >= self.max_episode_length
        for e in self.early.get(self.t, []):
            dones[e] = True
        for e in dones.nonzero().flatten().tolist():
            self.acc.setdefault(str(e), 0.25 + 0.01 * e + 0.001 * self.t)
        self.len[dones] = 0
        r = self.rew[self.t].clone()
        self.t += 1
        return torch.zeros(50, 4), r, dones, {}, self.acc

    def render(self):
        pass


class _FakeModel:
    def predict(self, obs, state=None, episode_start=None, deterministic=True):
        return torch.zeros(50, 6, dtype=torch.long), state


@pytest.mark.reference
def test_evaluation_loop_equals_the_reference_function_on_a_scripted_env():
    """The reference's own evaluate_policy_grid_obs (stable_baselines3/common/evaluation.py:136-347) and ours, driven by the
    same scripted env and model on CPU tensors: episode rewards / lengths / accuracies in the same order, same mean AUC."""
    import warnings
    import ref_loader
    ref_loader.load_reference()
    from stable_baselines3.common.evaluation import evaluate_policy_grid_obs as ref_eval
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r_ref, l_ref, auc_ref, acc_ref = ref_eval(_FakeModel(), _FakeEvalEnv(7), n_eval_episodes=50, warn=False)
    r, l, auc, acc = evaluate_policy_grid_obs(_FakeModel(), _FakeEvalEnv(7), n_eval_episodes=50)
    assert len(r) == len(r_ref) == 50
    np.testing.assert_allclose([float(x) for x in r], [float(x) for x in r_ref], rtol=1e-6)
    assert [int(x) for x in l] == [int(x) for x in l_ref]
    assert acc == acc_ref
    np.testing.assert_allclose(auc.numpy(), auc_ref.numpy(), rtol=1e-6)
