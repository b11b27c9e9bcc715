"""CPU, build container only: pins the GAE and PPO-update restatements (the oracles of the GPU parity tests) against the
reference's OWN code -- `TensorRolloutBuffer_Grid_Obs.compute_returns_and_advantage` (buffers.py:706-724) and
`PPO_Grid_Obs.collect_rollouts()` + `train()` (on_policy_algorithm_grid_obs.py:128-221, ppo_grid_obs.py:176-297) run
unmodified on a scripted env behind the reference's wrapper class -- and checks that the committed fixture
tests/golden/ppo_train_g20.npz is what that run produces."""
import os

import numpy as np
import pytest
import torch

import oracle as c_oracle
from helpers import GOLDEN_DIR

pytestmark = pytest.mark.reference


def test_c_oracle_gae_equals_reference_buffer():
    import ref_loader
    ref = ref_loader.load_reference()
    from gym import spaces
    T, N = 128, 37
    buf = ref.buffers.TensorRolloutBuffer_Grid_Obs(T, spaces.Box(-np.inf, np.inf, (5,), np.float32),
                                                   spaces.MultiDiscrete([81, 81, 51, 1, 13, 13]), device="cpu", gamma=0.99,
                                                   gae_lambda=0.95, n_envs=N)
    g = torch.Generator().manual_seed(0)
    buf.rewards.copy_(torch.randn(T, N, 1, generator=g))
    buf.values.copy_(torch.randn(T, N, 1, generator=g))
    buf.episode_starts.copy_((torch.rand(T, N, 1, generator=g) < 0.1).byte())
    last_v, dones = torch.randn(N, 1, generator=g), torch.rand(N, generator=g) < 0.3
    buf.compute_returns_and_advantage(last_values=last_v, dones=dones)
    adv, ret = c_oracle.gae(buf.rewards[..., 0].numpy(), buf.values[..., 0].numpy(), buf.episode_starts[..., 0].numpy(),
                            last_v[:, 0].numpy(), dones.numpy().astype(np.uint8), 0.99, 0.95)
    np.testing.assert_array_equal(adv, buf.advantages[..., 0].numpy())
    np.testing.assert_array_equal(ret, buf.returns[..., 0].numpy())


@pytest.mark.parametrize("target_kl", [None, 1e-7])
def test_policyref_loss_adam_loop_equals_reference_train(target_kl):
    import ref_ppo_driver as drv
    r = drv.run_reference(target_kl=target_kl)
    m = drv.mirror_train(r["mirror"], r["cols"], r["indices"], r["cfg"])
    logs, rec = m["logs"], r["logs"]
    # the rollout's GAE columns are the C oracle's
    c = r["cols"]
    adv, ret = c_oracle.gae(c["rewards"][..., 0].numpy(), c["values"][..., 0].numpy(), c["episode_starts"][..., 0].numpy(),
                            r["last_values"][:, 0].numpy(), r["last_dones"].numpy().astype(np.uint8), 0.99, 0.95)
    np.testing.assert_array_equal(adv, c["advantages"][..., 0].numpy())
    np.testing.assert_array_equal(ret, c["returns"][..., 0].numpy())
    # logged scalars: means over all minibatches, approx_kl over the LAST epoch only (ppo_grid_obs.py:199,287)
    assert rec["train/policy_gradient_loss"] == np.mean(logs[:, 1])
    assert rec["train/value_loss"] == np.mean(logs[:, 2])
    assert rec["train/entropy_loss"] == np.mean(logs[:, 3])
    assert rec["train/clip_fraction"] == np.mean(logs[:, 5])
    assert float(rec["train/approx_kl"]) == pytest.approx(m["last_epoch_kl"], rel=1e-6)
    assert rec["train/loss"] == logs[-1, 0]
    # every parameter and buffer after the update: bit-equal
    sd = r["mirror"].state_dict()
    for k, v in r["after"].items():
        assert torch.equal(v, sd[k]), k
    assert m["steps"] == (0 if target_kl is not None else 8)


def test_committed_ppo_fixture_is_the_reference_run(tmp_path):
    import ref_ppo_driver as drv
    fresh = np.load(drv.write_golden(str(tmp_path / "ppo.npz")))
    have = np.load(os.path.join(GOLDEN_DIR, "ppo_train_g20.npz"))
    assert sorted(fresh.files) == sorted(have.files)
    for k in fresh.files:
        np.testing.assert_array_equal(fresh[k], have[k], err_msg=k)
