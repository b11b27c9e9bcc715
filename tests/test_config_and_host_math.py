"""CPU: constants mirrored from the reference's config classes, and the integer-division shortcut used by the tensor-core
kernels' row decode."""
import numpy as np
import pytest

from gennbv_b200.config import Config_GenNBV_Eval, Config_GenNBV_Train


@pytest.mark.reference
def test_config_values_equal_the_reference_classes():
    import ref_loader
    ref_loader.load_reference()
    from gennbv.env.config_gennbv_eval import Config_GenNBV_Eval as RefEval
    from gennbv.env.config_gennbv_train import Config_GenNBV_Train as RefTrain
    for mine, ref in ((Config_GenNBV_Train, RefTrain), (Config_GenNBV_Eval, RefEval)):
        assert mine.max_episode_length == ref.max_episode_length
        assert mine.rewards.only_positive_rewards == ref.rewards.only_positive_rewards
        for k in ("surface_coverage", "short_path", "termination"):
            assert getattr(mine.rewards.scales, k, None) == getattr(ref.rewards.scales, k, None), (mine.__name__, k)
        nz, rz = mine.normalization, ref.normalization
        for k in ("clip_pose_low", "clip_pose_idx_up", "clip_pose_idx_low", "init_pose_buf", "init_action", "action_unit"):
            np.testing.assert_allclose(getattr(nz, k), getattr(rz, k), rtol=0, atol=0, err_msg=k)
        assert mine.visual_input.stack == ref.visual_input.stack
        assert mine.visual_input.horizontal_fov == ref.visual_input.horizontal_fov
        assert (mine.visual_input.camera_height, mine.visual_input.camera_width) == \
               (ref.visual_input.camera_height, ref.visual_input.camera_width)
        assert mine.env.episode_length_s == ref.env.episode_length_s and mine.env.env_spacing == ref.env.env_spacing
        assert mine.termination.collision == ref.termination.collision
        assert mine.termination.max_step_done == ref.termination.max_step_done
    # the eval config REPLACES the rewards class: no short_path / termination terms (config_gennbv_eval.py:9-15)
    assert not hasattr(Config_GenNBV_Eval.rewards.scales, "short_path")


def test_fast_div_is_exact():
    """conv2_mma.cu::fast_div: floor((v + 0.5f) * (1.f / n)) == v // n in fp32 for every v < 2^22 and n < 200."""
    v = np.arange(0, 1 << 22, dtype=np.int64)
    vf = v.astype(np.float32) + np.float32(0.5)
    for n in list(range(1, 40)) + [63, 64, 65, 99, 127, 128, 160, 199]:
        inv = np.float32(1.0) / np.float32(n)
        assert np.array_equal(np.floor(vf * inv).astype(np.int64), v // n), n
