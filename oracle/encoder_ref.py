"""Plain-PyTorch fp32 restatement of the policy network: Hybrid_Encoder (gennbv/network/hybrid_encoder.py:12-91)
with the three hard-coded 20^3 constants (8000 / 20 / 1024, lines :40,:83-84) parametrised by the grid size, the
actor / critic heads (stable_baselines3/common/policies.py:984,994) and the MultiCategorical distribution
(distributions.py:299-352).

TEST INFRASTRUCTURE ONLY.  Pinned at G = 20 against the reference's own modules in the build container
(tests/test_encoder_ref_vs_reference.py); for G != 20 there is no reference to run (its forward cannot execute),
so this module *is* the oracle there -- "parity unpinned" in SURVEY.md section 8c's terms.
"""
import torch
from torch import nn

ACTION_NVEC = (81, 81, 51, 1, 13, 13)


class HybridEncoderRef(nn.Module):
    def __init__(self, grid_size=20, state_dim=600, features_dim=256, semantic=False):
        super().__init__()
        self.G, self.state_dim, self.features_dim, self.semantic = grid_size, state_dim, features_dim, semantic
        g1 = (grid_size - 3) // 2 + 1
        g2 = (g1 - 3) // 2 + 1
        self.naive_encoder_grid = nn.Sequential(
            nn.Conv3d(1, 16, kernel_size=3, stride=2, padding=0), nn.BatchNorm3d(16), nn.ReLU(inplace=True),
            nn.Conv3d(16, 16, kernel_size=3, stride=2, padding=0), nn.BatchNorm3d(16), nn.ReLU(inplace=True))
        self.output_layer_grid = nn.Sequential(nn.Linear(16 * g2 ** 3, 256), nn.ReLU(inplace=True))
        self.naive_encoder_action = nn.Sequential(nn.Linear(4 * state_dim, 256), nn.ReLU(inplace=True),
                                                  nn.Linear(256, 256), nn.ReLU(inplace=True))
        self.output_layer = nn.Sequential(nn.Linear(768 if semantic else 512, features_dim), nn.ReLU(inplace=True))
        if semantic:
            # 2-D branch over the k = 2 grayscale 64x64 frames behind the grid.  No reference forward exists
            # for it (hybrid_encoder.py:69-91 never reads the frames): PARITY UNPINNED, layer sizes are this build's choice.
            self.naive_encoder_rgb = nn.Sequential(nn.Conv2d(2, 16, kernel_size=3, stride=2, padding=0), nn.ReLU(inplace=True),
                                                   nn.Conv2d(16, 16, kernel_size=3, stride=2, padding=0), nn.ReLU(inplace=True))
            self.output_layer_rgb = nn.Sequential(nn.Linear(16 * 15 * 15, 256), nn.ReLU(inplace=True))

    @staticmethod
    def positional_encoding(positions, freqs=2):
        bands = (2 ** torch.arange(freqs).float()).to(positions.device)
        pts = (positions[..., None] * bands).reshape(positions.shape[:-1] + (freqs * positions.shape[-1],))
        return torch.cat([torch.sin(pts), torch.cos(pts)], dim=-1)

    def forward(self, obs):
        n, G = obs.shape[0], self.G
        a = obs[:, :self.state_dim].view(n, -1, 6)
        fa = self.naive_encoder_action(self.positional_encoding(a).view(n, -1))
        g = obs[:, self.state_dim:self.state_dim + G ** 3].reshape(n, 1, G, G, G)
        fg = self.output_layer_grid(self.naive_encoder_grid(g).reshape(n, -1))
        if not self.semantic:
            return self.output_layer(torch.cat((fa, fg), dim=-1))
        r = obs[:, self.state_dim + G ** 3:self.state_dim + G ** 3 + 2 * 64 * 64].reshape(n, 2, 64, 64)
        fr = self.output_layer_rgb(self.naive_encoder_rgb(r).reshape(n, -1))
        return self.output_layer(torch.cat((fa, fg, fr), dim=-1))


class PolicyRef(nn.Module):
    """ActorCriticPolicy_Train_Eval with net_arch=[] (policies.py:954-1090): heads directly on the encoder output."""

    def __init__(self, grid_size=20, state_dim=600, nvec=ACTION_NVEC, semantic=False):
        super().__init__()
        self.nvec = tuple(nvec)
        self.features_extractor = HybridEncoderRef(grid_size, state_dim, semantic=semantic)
        self.action_net = nn.Linear(256, sum(nvec))
        self.value_net = nn.Linear(256, 1)

    def dists(self, logits):
        return [torch.distributions.Categorical(logits=s) for s in torch.split(logits, self.nvec, dim=1)]

    def evaluate_actions(self, obs, actions):
        f = self.features_extractor(obs)
        d = self.dists(self.action_net(f))
        lp = torch.stack([di.log_prob(a) for di, a in zip(d, torch.unbind(actions, dim=1))], dim=1).sum(dim=1)
        ent = torch.stack([di.entropy() for di in d], dim=1).sum(dim=1)
        return self.value_net(f), lp, ent


def seeded_state_dict(module, seed, scale=1.0):
    """Deterministic, generator-driven weights (identical on every machine): N(0,1)*fan-in scaling for weights,
    small biases, BN affine near (1, 0) and non-trivial running statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in module.state_dict().items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif k.endswith("running_mean"):
            sd[k] = 0.2 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1 and ("naive_encoder_grid.1" in k or "naive_encoder_grid.4" in k) and k.endswith("weight"):
            sd[k] = 1.0 + 0.2 * torch.randn(v.shape, generator=g)
        elif v.dim() == 1:
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            sd[k] = scale * torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def ppo_loss(values, log_prob, entropy, old_values, old_log_prob, advantages, returns, clip_range=0.2, clip_range_vf=0.2,
             ent_coef=0.01, vf_coef=0.8, pg_coef=10.0, normalize_advantage=True):
    """The loss lines of PPO_Grid_Obs.train (stable_baselines3/ppo/ppo_grid_obs.py:213-262)."""
    values = values.flatten()
    if normalize_advantage:
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    ratio = torch.exp(log_prob - old_log_prob)
    policy_loss = -torch.min(advantages * ratio, advantages * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
    clip_fraction = torch.mean((torch.abs(ratio - 1) > clip_range).float())
    if clip_range_vf is None:
        values_pred = values
    else:
        values_pred = old_values + torch.clamp(values - old_values, -clip_range_vf, clip_range_vf)
    value_loss = torch.nn.functional.mse_loss(returns, values_pred)
    entropy_loss = -torch.mean(entropy)
    loss = policy_loss * pg_coef + ent_coef * entropy_loss + vf_coef * value_loss
    with torch.no_grad():
        log_ratio = log_prob - old_log_prob
        approx_kl = torch.mean((torch.exp(log_ratio) - 1) - log_ratio)
    return loss, dict(policy_loss=policy_loss, value_loss=value_loss, entropy_loss=entropy_loss, approx_kl=approx_kl,
                      clip_fraction=clip_fraction)
