"""Config values of the GenNBV training task, mirroring the attribute paths the env reads from the reference's
`Config_GenNBV_Train` (gennbv/env/config_gennbv_train.py:6-73, inherited defaults from config_legged.py and
legged_gym's base config).  Only values on the hot path are kept; Isaac Gym asset / terrain / PD-control settings
are out of scope."""
import math


class Config_GenNBV_Train:
    max_episode_length = 100                      # config_gennbv_train.py:11
    return_visual_observation = True              # config_legged.py:9

    class rewards:
        class scales:                             # config_gennbv_train.py:14-17 (multiplied by dt at start-up)
            surface_coverage = 1000
            short_path = 5
            termination = 50
        only_positive_rewards = True              # :20; train_gennbv.py:101-106 overrides it to False by default

    class visual_input:
        camera_width = 400                        # :24-26
        camera_height = 400
        horizontal_fov = 90.0
        stack = 100                               # :28
        normalization = True

    class env:
        num_envs = 256
        num_observations = 6
        episode_length_s = 20                     # :47
        num_actions = 6
        env_spacing = 5
        send_timeouts = True                      # legged_robot_config default

    class normalization:
        pi = 3.14159265359                        # :58
        clip_pose_low = [-8., -8., 0.1, 0., -1 / 2 * pi, 0.]          # :62-64
        clip_pose_idx_up = [80, 80, 50, 0, 12, 12]
        clip_pose_idx_low = [0, 0, 0, 0, 0, 0]
        init_pose_buf = [0., 0., 10.1, 0., 90 / 180 * math.pi, 0.]    # :67-69
        init_action = [40, 40, 50, 0, 12, 0]
        action_unit = [0.2, 0.2, 0.2, 0., 1 / 12 * math.pi, 1 / 6 * math.pi]

    class termination:
        collision = True                          # :72-73
        max_step_done = True

    class sim:
        dt = 0.005                                # legged_robot_config sim.dt (a C float inside gymapi.SimParams)

    class control:
        decimation = 4                            # dt = decimation * sim_params.dt  (drone_robot.py:874-875)


class Config_GenNBV_Eval(Config_GenNBV_Train):
    """gennbv/env/config_gennbv_eval.py:6-15.  `rewards` is redefined, not extended: only the coverage term remains
    (scale 50 x dt 0.02 = 1, "just easy to evaluate") and negative totals are clipped."""
    max_episode_length = 30

    class rewards:
        class scales:
            surface_coverage = 50
        only_positive_rewards = True
        max_contact_force = 100.
