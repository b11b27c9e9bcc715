"""ctypes binding of libgennbv_b200.so (the C ABI declared in include/gennbv_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgennbv_b200.so")
ABI_VERSION = 3

_lib = None

c_void_p, c_int, c_int64, c_size_t, c_uint32, c_double, c_float, c_uint64 = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_double, ctypes.c_float,
    ctypes.c_uint64)


class EncoderParams(ctypes.Structure):
    """gnbv_encoder_params (include/gennbv_b200.h): 24 device pointers in declaration order."""
    FIELDS = ["conv1_w", "conv1_b", "bn1_w", "bn1_b", "bn1_rm", "bn1_rv", "bn1_nbt",
              "conv2_w", "conv2_b", "bn2_w", "bn2_b", "bn2_rm", "bn2_rv", "bn2_nbt",
              "grid_fc_w", "grid_fc_b", "act_fc1_w", "act_fc1_b", "act_fc2_w", "act_fc2_b", "out_fc_w", "out_fc_b",
              "rgb_conv1_w", "rgb_conv1_b", "rgb_conv2_w", "rgb_conv2_b", "rgb_fc_w", "rgb_fc_b"]     # last six: semantic branch or NULL
    _fields_ = [(f, ctypes.c_void_p) for f in FIELDS]


class EncoderGrads(ctypes.Structure):
    """gnbv_encoder_grads: 16 device pointers in declaration order (== Hybrid_Encoder._param_list())."""
    FIELDS = ["conv1_w", "conv1_b", "bn1_w", "bn1_b", "conv2_w", "conv2_b", "bn2_w", "bn2_b",
              "grid_fc_w", "grid_fc_b", "act_fc1_w", "act_fc1_b", "act_fc2_w", "act_fc2_b", "out_fc_w", "out_fc_b",
              "rgb_conv1_w", "rgb_conv1_b", "rgb_conv2_w", "rgb_conv2_b", "rgb_fc_w", "rgb_fc_b"]
    _fields_ = [(f, ctypes.c_void_p) for f in FIELDS]


class PpoMinibatch(ctypes.Structure):
    """gnbv_ppo_minibatch (include/gennbv_b200.h), field for field."""
    _fields_ = [("enc", ctypes.POINTER(EncoderParams)), ("enc_grads", ctypes.POINTER(EncoderGrads)),
                ("head_w", c_void_p), ("head_b", c_void_p), ("head_w_grad", c_void_p), ("head_b_grad", c_void_p),
                ("nvec", ctypes.POINTER(c_int)), ("num_sub", c_int), ("feat_dim", c_int),
                ("observations", c_void_p), ("obs_row_stride", c_int64),
                ("actions", c_void_p), ("values", c_void_p), ("log_probs", c_void_p), ("advantages", c_void_p),
                ("returns", c_void_p), ("storage_rows", c_void_p), ("rows_base", c_int64),
                ("batch", c_int), ("grid_size", c_int), ("state_dim", c_int), ("normalize_advantage", c_int),
                ("clip_range", c_double), ("clip_range_vf", c_double), ("ent_coef", c_double), ("vf_coef", c_double),
                ("pg_coef", c_double), ("target_kl", c_double),
                ("ctl", c_void_p), ("vote", c_void_p), ("log", c_void_p), ("log_capacity", c_int64),
                ("enc_workspace", c_void_p), ("enc_workspace_bytes", c_size_t),
                ("mb_workspace", c_void_p), ("mb_workspace_bytes", c_size_t)]


# name -> (restype, argtypes); mirrors include/gennbv_b200.h one to one
SIGNATURES = {
    "gnbv_abi_version": (c_int, []),
    "gnbv_kernel_mode": (c_int, [c_int]),
    "gnbv_last_error": (ctypes.c_char_p, []),
    "gnbv_profile_enable": (c_int, [c_int]),
    "gnbv_profile_elapsed_ms": (c_int, [c_int, c_int, ctypes.POINTER(c_float)]),
    "gnbv_voxelize_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gnbv_voxelize_step": (c_int, [c_void_p] * 11 + [c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                                      c_int, c_int, c_int, c_int, c_uint32, c_void_p]),
    "gnbv_scan_raycast": (c_int, [c_void_p] * 9 + [c_size_t, c_int, c_int, c_int, c_int, c_uint32, c_void_p]),
    "gnbv_grid_update": (c_int, [c_void_p] * 4 + [c_int64, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "gnbv_grid_update_sparse": (c_int, [c_void_p] * 4 + [c_int64, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "gnbv_voxelize_masks": (c_int, [c_void_p, c_int, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p),
                                    ctypes.POINTER(c_int64)]),
    "gnbv_points_to_voxel_mask": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "gnbv_bresenham_rays": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gnbv_reset_grids": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "gnbv_actions_to_poses": (c_int, [c_void_p] * 9 + [c_int, c_int, c_void_p]),
    "gnbv_obs_update": (c_int, [c_void_p] * 5 + [c_int64] * 3 + [c_int] * 8 + [c_void_p]),
    "gnbv_episode_stats_doubles": (c_size_t, []),
    "gnbv_reward_termination": (c_int, [c_void_p] * 14 + [c_double] * 3 + [c_int] * 3 + [c_int64, c_double, c_double,
                                                                                       c_int, c_int, c_void_p]),
    "gnbv_reset_envs": (c_int, [c_void_p] * 11 + [c_int] * 8 + [c_void_p]),
    "gnbv_encoder_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gnbv_encoder_forward": (c_int, [ctypes.POINTER(EncoderParams), c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "gnbv_encoder_backward": (c_int, [ctypes.POINTER(EncoderParams), c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, ctypes.POINTER(EncoderGrads), c_void_p, c_size_t, c_void_p]),
    "gnbv_encoder_workspace_view": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_int64), ctypes.POINTER(c_int64)]),
    "gnbv_debug_ts_profile": (c_int, [ctypes.POINTER(c_uint64)]),
    "gnbv_encoder_backward_phase": (c_int, [ctypes.POINTER(EncoderParams), c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int,
                                            c_void_p, c_void_p, ctypes.POINTER(EncoderGrads), c_void_p, c_size_t, c_int, c_void_p]),
    "gnbv_ppo_minibatch_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "gnbv_ppo_apply_workspace_bytes": (c_size_t, []),
    "gnbv_ppo_minibatch_grads": (c_int, [ctypes.POINTER(PpoMinibatch), c_int, c_void_p]),
    "gnbv_ppo_minibatch_scalars": (c_int, [ctypes.POINTER(PpoMinibatch), ctypes.POINTER(c_void_p)]),
    "gnbv_ppo_minibatch_apply": (c_int, [ctypes.POINTER(PpoMinibatch), c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                         c_double, c_double, c_double, c_double, c_double, c_double, c_void_p, c_void_p]),
    "gnbv_policy_heads_forward": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_void_p]),
    "gnbv_multicategorical_evaluate": (c_int, [c_void_p, c_int64, ctypes.POINTER(c_int), c_int, c_void_p, c_void_p,
                                               c_void_p, c_int, c_void_p]),
    "gnbv_multicategorical_sample": (c_int, [c_void_p, c_int64, ctypes.POINTER(c_int), c_int, c_uint64, c_uint64, c_int,
                                             c_void_p, c_void_p, c_int, c_void_p]),
    "gnbv_multicategorical_backward": (c_int, [c_void_p, c_int64, ctypes.POINTER(c_int), c_int, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "gnbv_ppo_loss": (c_int, [c_void_p] * 7 + [c_int] + [c_double] * 5 + [c_int] + [c_void_p] * 5),
    "gnbv_clip_adam_workspace_bytes": (c_size_t, []),
    "gnbv_grad_norm": (c_int, [c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "gnbv_adam_step": (c_int, [c_void_p] * 4 + [c_int64, c_void_p] + [c_double] * 4 + [c_int64, c_double, c_void_p]),
    "gnbv_sgemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gnbv_sgemm": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                           c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "gnbv_tc_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gnbv_tc_gemm": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                             c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "gnbv_chamfer_workspace_bytes": (c_size_t, [c_int]),
    "gnbv_chamfer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gnbv_chamfer_grid_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int]),
    "gnbv_chamfer_grid": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gnbv_nn_sqdist_brute": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gnbv_scan_points": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_int64, c_uint32, c_void_p]),
    "gnbv_keys_to_points": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gnbv_points_to_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gnbv_pack_env_keys": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "gnbv_sort_unique_workspace_bytes": (c_size_t, [c_int64]),
    "gnbv_sort_unique_u64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gnbv_gae": (c_int, [c_void_p] * 5 + [c_double, c_double, c_int, c_int, c_void_p, c_void_p, c_void_p]),
}


def lib():
    """The loaded library; raises if it has not been built (python -m gennbv_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: gennbv_b200 has no CPU fallback. Build it with "
                "`python -m gennbv_b200.build` (nvcc, sm_100a).")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)          # AttributeError if the export is missing
            fn.restype, fn.argtypes = res, args
        got = h.gnbv_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"libgennbv_b200.so ABI {got} != binding ABI {ABI_VERSION}: rebuild")
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().gnbv_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


# stage ids of include/gennbv_b200.h (each marks the START of the named stage)
ENCODER_STAGES = {"fwd.action_mlp": (0, 1), "fwd.conv1": (1, 2), "fwd.bn1_stats": (2, 3), "fwd.conv2": (3, 4),
                  "fwd.bn2_stats+apply": (4, 5), "fwd.grid_fc": (5, 6), "fwd.out_fc": (6, 7),
                  "bwd.linear_layers": (16, 17), "bwd.grid_fc": (17, 18), "bwd.bn2": (18, 19), "bwd.conv2_wgrad": (19, 20),
                  "bwd.conv2_dgrad": (20, 21), "bwd.bn1_stats": (21, 22), "bwd.conv1_wgrad": (22, 23)}


def stage_ms(name):
    a, b = ENCODER_STAGES[name]
    out = c_float()
    check(lib().gnbv_profile_elapsed_ms(a, b, ctypes.byref(out)), "gnbv_profile_elapsed_ms")
    return out.value
