"""BASELINE configs[2]: one full PPO iteration (256 envs x 128 steps rollout + 5 epochs x 256 minibatches of 128) on one
B200 through gennbv_b200.PPO_Grid_Obs, Houses3K-shape synthetic depth (128x128, 64^3).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gennbv_b200.config import Config_GenNBV_Train
from gennbv_b200.env import Env_Train_GenNBV
from gennbv_b200.ppo import PPO_Grid_Obs
from gennbv_b200.sensors import FrameListSensor
from gennbv_b200.wrapper import EnvWrapperGenNBVTrain

T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
wl = bench.make_workload(bench.ENVS_PER_GPU, dev, seed=0)

class Cfg(Config_GenNBV_Train):
    class rewards(Config_GenNBV_Train.rewards):
        only_positive_rewards = False

env = EnvWrapperGenNBVTrain(Env_Train_GenNBV(Cfg(), sim_device=str(dev), sensor=FrameListSensor(wl["frames"], dev),
                                             grid_gt=wl["scenes"].grid_gt, num_envs=bench.ENVS_PER_GPU))
kw = dict(net_arch=[], features_extractor_kwargs=dict(
    encoder_param={"hidden_shapes": [256, 256], "visual_dim": 256},
    net_param={"transformer_params": [[1, 256], [1, 256]], "append_hidden_shapes": [256, 256]},
    state_input_shape=(600,), visual_input_shape=(100, 128, 128)))
out = {}
for target_kl in ((None,) if os.environ.get("PPO_ITER_FAST") else (None, 0.05)):
    algo = PPO_Grid_Obs(env=env, learning_rate=1e-4, n_steps=T, batch_size=128, n_epochs=5, gamma=0.99, gae_lambda=0.95,
                        clip_range=0.2, clip_range_vf=0.2, ent_coef=0.01, vf_coef=0.8, max_grad_norm=1, target_kl=target_kl,
                        policy_kwargs=kw, seed=1, device=dev)
    algo._setup_learn()
    algo.collect_rollouts(n_rollout_steps=None)          # warm-up iteration (allocations, first-touch)
    algo.train()
    torch.cuda.synchronize()
    t0 = time.time(); algo.collect_rollouts(); torch.cuda.synchronize(); t1 = time.time()
    algo.train(); torch.cuda.synchronize(); t2 = time.time()
    steps = T * bench.ENVS_PER_GPU
    out["target_kl=%s" % target_kl] = {"rollout_s": t1 - t0, "train_s": t2 - t1, "optimizer_steps": algo._adam_step // 2,
                                        "env_steps_per_sec": steps / (t2 - t0), "rollout_env_steps_per_sec": steps / (t1 - t0),
                                        "logs": {k: v for k, v in algo.logger.name_to_value.items() if k.startswith("train/")}}
    del algo
    torch.cuda.empty_cache()
print(json.dumps({"workload": f"configs[2]: 256 envs x {T} steps, 5 epochs x minibatch 128, 128x128 depth, 64^3 grid", **out}))
