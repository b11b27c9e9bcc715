"""CPU, build container only: pins oracle/encoder_ref.py (the torch restatement used as the encoder oracle on the
GPU box) against the reference's own Hybrid_Encoder / ActorCriticPolicy_Train_Eval at the native 20^3 grid, and
checks that gennbv_b200's module mirrors the reference's state_dict keys and seeded initialisation."""
import numpy as np
import pytest
import torch

import encoder_ref

pytestmark = pytest.mark.reference


def _ref_policy():
    import ref_loader
    ref = ref_loader.load_reference()
    from gym import spaces
    D = 600 + 8000 + 8192
    kwargs = dict(encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
                  net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
                  state_input_shape=(600,), visual_input_shape=(100, 48, 48))
    make = lambda: ref.policies.ActorCriticPolicy_Train_Eval(
        spaces.Box(low=-np.inf, high=np.inf, shape=(D,), dtype=np.float32), spaces.MultiDiscrete([81, 81, 51, 1, 13, 13]),
        lambda _: 1e-4, net_arch=[], features_extractor_class=ref.encoder.Hybrid_Encoder,
        features_extractor_kwargs={k: (dict(v) if isinstance(v, dict) else v) for k, v in
                                   {**kwargs, "net_param": {"transformer_params": [[1, 256], [1, 256]],
                                                            "append_hidden_shapes": [256, 256]}}.items()})
    return make, D


def test_restatement_equals_reference_modules():
    make, D = _ref_policy()
    pol = make()
    mirror = encoder_ref.PolicyRef(20, 600)
    assert list(mirror.state_dict().keys()) == list(pol.state_dict().keys())
    sd = encoder_ref.seeded_state_dict(mirror, 3)
    pol.load_state_dict(sd); mirror.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    obs = torch.cat([torch.randn(7, 600, generator=g), torch.randint(-1, 2, (7, 8000), generator=g).float(),
                     torch.rand(7, 8192, generator=g)], 1)
    actions = torch.stack([torch.randint(0, n, (7,), generator=g) for n in (81, 81, 51, 1, 13, 13)], 1)
    for training in (False, True):
        pol.set_training_mode(training); mirror.train(training)
        a = pol.evaluate_actions(obs, actions)
        b = mirror.evaluate_actions(obs, actions)
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert sum(p.numel() for p in pol.parameters()) == 1_143_553


def test_product_module_has_reference_keys_and_seeded_init():
    """gennbv_b200's policy (built on the CPU here: construction only, no kernel call) consumes the torch RNG like the
    reference, so the same seed gives the same initial weights, and its state_dict keys are the checkpoint's."""
    make, D = _ref_policy()
    from gennbv_b200.policy import ActorCriticPolicy_Train_Eval
    from gennbv_b200.spaces import Box, MultiDiscrete
    torch.manual_seed(11)
    ref_pol = make()
    torch.manual_seed(11)
    kwargs = dict(encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
                  net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
                  state_input_shape=(600,), visual_input_shape=(100, 48, 48))
    mine = ActorCriticPolicy_Train_Eval(Box(-np.inf, np.inf, (D,), np.float32), MultiDiscrete([81, 81, 51, 1, 13, 13]),
                                        lambda _: 1e-4, net_arch=[], features_extractor_kwargs=kwargs, device="cpu")
    a, b = ref_pol.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert kwargs["net_param"]["append_hidden_shapes"] == [256]      # the pop of hybrid_encoder.py:27-28
