"""GPU parity tests of the voxelize path (through the C ABI) against the goldens recorded from the
reference and against the C oracle on fresh seeded inputs; full-size checks use size-independent properties."""
import numpy as np
import pytest
import torch

import oracle as c_oracle
from gennbv_b200 import ops, synth
from helpers import ENV_GOLDENS, EnvGolden, replay_voxelize

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


def cuda_step(depth, seg, kinv, c2w, range_gt, vs, pose_xyz, grid_gt, prob, scan, raw_depth, want_masks=False):
    """numpy in / numpy out wrapper of one gnbv_voxelize_step call (prob / scan updated in place)."""
    N, G = prob.shape[0], prob.shape[1]
    p, s = cu(prob), cu(scan)
    tri = torch.full((N, G, G, G), 7.0, device=DEV)
    cov = torch.full((N,), -1.0, device=DEV)
    nt = torch.full((N,), -1, dtype=torch.int32, device=DEV)
    ws = ops.voxelize_step(cu(depth), cu(seg), cu(kinv), cu(c2w), cu(range_gt), cu(vs), cu(pose_xyz), cu(grid_gt), p, s,
                           tri, cov, nt, raw_depth=raw_depth)
    prob[...] = p.cpu().numpy()
    scan[...] = s.cpu().numpy()
    out = dict(tri=tri.cpu().numpy(), cov_sum=cov.cpu().numpy(), num_targets=nt.cpu().numpy())
    if want_masks:
        t, r = ops.voxelize_masks(ws, N, G)
        out.update(target_mask=t.cpu().numpy(), touched_mask=r.cpu().numpy())
    return out


@pytest.mark.parametrize("name", ENV_GOLDENS)
def test_golden_rollout(name):
    g = EnvGolden(name)

    def step(depth_raw, seg, c2w, pose_xyz, prob, scan):
        out = cuda_step(depth_raw, seg, g.inv_intri, c2w, g.range_gt, g.voxel_size_gt, pose_xyz, g.grid_gt, prob, scan, True)
        return out["tri"], out["cov_sum"]

    assert replay_voxelize(g, step) == g.T + 1


def synthetic_case(N, H, W, G, S, seed, steps):
    scenes = synth.make_house_scenes(S, G, seed=seed)
    vs, nvalid, rg = synth.gt_metadata(scenes.grid_gt)
    idx = torch.arange(N) % S
    gen = torch.Generator().manual_seed(seed + 100)
    kinv = torch.linalg.inv(synth.camera_intrinsics(H, W)).float()
    frames = []
    for _ in range(steps):
        a = synth.sample_lookat_actions(scenes.params, N, gen)
        poses = synth.pose_from_action(a)
        depth, seg, _, c2w = synth.render(scenes.params, poses, H, W)
        frames.append((depth.numpy(), seg.numpy(), c2w.float().numpy(), poses[:, :3].contiguous().numpy()))
    return dict(kinv=kinv.numpy(), range_gt=rg[idx].numpy(), vs=vs[idx].numpy(),
                grid_gt=scenes.grid_gt[..., 3][idx].contiguous().numpy(), frames=frames)


@pytest.mark.parametrize("N,H,W,G,S,seed", [(6, 128, 128, 64, 3, 11), (5, 33, 35, 21, 2, 12), (3, 64, 48, 32, 3, 13),
                                           (2, 400, 400, 20, 2, 14)])
def test_against_oracle(N, H, W, G, S, seed):
    c = synthetic_case(N, H, W, G, S, seed, steps=3)
    prob_o = np.zeros((N, G, G, G), np.float32); scan_o = np.zeros_like(prob_o)
    prob_c = np.zeros_like(prob_o); scan_c = np.zeros_like(prob_o)
    for depth, seg, c2w, xyz in c["frames"]:
        o = c_oracle.voxelize_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_o, scan_o,
                                   raw_depth=True, want_masks=True)
        k = cuda_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_c, scan_c, True, True)
        assert o["num_targets"].sum() > 0
        for key in ("num_targets", "target_mask", "touched_mask", "tri", "cov_sum"):
            np.testing.assert_array_equal(k[key], o[key], err_msg=key)
        np.testing.assert_array_equal(prob_c, prob_o)
        np.testing.assert_array_equal(scan_c, scan_o)


def test_more_distinct_targets_than_the_ray_list_holds(seed=31):
    """The kernel lists distinct targets while it scatters them (8192 slots); past that it rebuilds the list from the target
    mask in rounds.  Every pixel foreground with a random depth across the grid's bounding sphere: > 8192 distinct voxels."""
    N, H, W, G = 2, 512, 512, 64
    c = synthetic_case(N, H, W, G, 2, seed, steps=1)
    depth, seg, c2w, xyz = c["frames"][0]
    rng = np.random.default_rng(seed)
    rg = c["range_gt"]
    ctr = np.stack([(rg[:, 0] + rg[:, 1]) / 2, (rg[:, 2] + rg[:, 3]) / 2, (rg[:, 4] + rg[:, 5]) / 2], 1)
    R = np.abs(np.stack([rg[:, 0] - rg[:, 1], rg[:, 2] - rg[:, 3], rg[:, 4] - rg[:, 5]], 1)).max(1) / 2
    dist = np.linalg.norm(xyz - ctr, axis=1)
    depth = -np.abs((dist - R)[:, None, None] + rng.uniform(0, 1, depth.shape) * (2 * R)[:, None, None]).astype(np.float32)
    seg = np.full_like(seg, 255)
    prob_o = np.zeros((N, G, G, G), np.float32); scan_o = np.zeros_like(prob_o)
    prob_c = np.zeros_like(prob_o); scan_c = np.zeros_like(prob_o)
    o = c_oracle.voxelize_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_o, scan_o,
                               raw_depth=True, want_masks=True)
    k = cuda_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_c, scan_c, True, True)
    assert int(o["num_targets"].min()) > 8192, o["num_targets"]
    for key in ("num_targets", "target_mask", "touched_mask", "tri", "cov_sum"):
        np.testing.assert_array_equal(k[key], o[key], err_msg=key)
    np.testing.assert_array_equal(prob_c, prob_o)


def test_ray_sources_inside_behind_and_far(seed=21):
    """Bresenham entry logic: camera voxel inside the grid, on a face, far away (|idx| ~ 1e5), equal to a target."""
    N, H, W, G = 8, 64, 64, 32
    c = synthetic_case(N, H, W, G, 2, seed, steps=1)
    depth, seg, c2w, xyz = c["frames"][0]
    rng = np.random.default_rng(seed)
    lo = c["range_gt"][:, [1, 3, 5]] - 0.5 * c["vs"]
    src = np.stack([lo + c["vs"] * rng.uniform(0, G, 3).astype(np.float32) for _ in range(N)])[:, 0].astype(np.float32)
    src[1] = lo[1] + c["vs"][1] * np.array([0.5, 0.5, 0.5], np.float32)          # corner voxel
    src[2] = lo[2] + c["vs"][2] * np.array([-1e5, 3.0, 2e5], np.float32)         # very far
    src[3] = lo[3] + c["vs"][3] * np.array([G + 0.5, -0.5, G / 2], np.float32)   # just outside two faces
    prob_o = np.zeros((N, G, G, G), np.float32); scan_o = np.zeros_like(prob_o)
    prob_c = np.zeros_like(prob_o); scan_c = np.zeros_like(prob_o)
    o = c_oracle.voxelize_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], src, c["grid_gt"], prob_o, scan_o,
                               raw_depth=True, want_masks=True)
    k = cuda_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], src, c["grid_gt"], prob_c, scan_c, True, True)
    np.testing.assert_array_equal(k["touched_mask"], o["touched_mask"])
    np.testing.assert_array_equal(prob_c, prob_o)


def test_empty_and_degenerate_inputs():
    """No foreground at all / NaN and +-inf depth: no targets, prob untouched (the reference `continue`s,
    env_train_gennbv.py:298-299), tri still recomputed from the carried prob grid."""
    N, H, W, G = 3, 32, 32, 20
    c = synthetic_case(N, H, W, G, 2, 31, steps=1)
    depth, seg, c2w, xyz = c["frames"][0]
    seg0 = seg.copy(); seg0[0] = 0                               # env 0: nothing segmented
    depth0 = depth.copy(); depth0[1, ::2] = np.nan; depth0[1, 1::2] = np.inf    # env 1: garbage depth
    rng = np.random.default_rng(0)
    prob0 = rng.choice(np.array([0, 1, 0.5, -0.05, 0.55, 0.49999988], np.float32), (N, G, G, G)).astype(np.float32)
    scan0 = (rng.random((N, G, G, G)) < 0.1).astype(np.float32)
    prob_o, scan_o, prob_c, scan_c = prob0.copy(), scan0.copy(), prob0.copy(), scan0.copy()
    o = c_oracle.voxelize_step(depth0, seg0, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_o, scan_o,
                               raw_depth=True)
    k = cuda_step(depth0, seg0, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_c, scan_c, True)
    assert k["num_targets"][0] == 0 and (prob_c[0] == prob0[0]).all()
    for key in ("num_targets", "tri", "cov_sum"):
        np.testing.assert_array_equal(k[key], o[key], err_msg=key)
    np.testing.assert_array_equal(prob_c, prob_o)
    np.testing.assert_array_equal(scan_c, scan_o)


def test_tri_written_into_flat_observation_rows():
    """tri_row_stride: the tri-class grid lands in the `grid` columns of the flattened observation
    (env_wrapper_gennbv_train.py:27-56: state(600) | grid(G^3) | rgb(8192)) without touching its neighbours."""
    N, H, W, G = 4, 48, 48, 20
    c = synthetic_case(N, H, W, G, 2, 41, steps=1)
    depth, seg, c2w, xyz = c["frames"][0]
    D = 600 + G ** 3 + 8192
    obs = torch.full((N, D), 9.0, device=DEV)
    prob = torch.zeros(N, G, G, G, device=DEV); scan = torch.zeros_like(prob)
    dense = torch.zeros_like(prob); cov = torch.zeros(N, device=DEV); cov2 = torch.zeros(N, device=DEV)
    args = [cu(depth), cu(seg), cu(c["kinv"]), cu(c2w), cu(c["range_gt"]), cu(c["vs"]), cu(xyz), cu(c["grid_gt"])]
    ops.voxelize_step(*args, prob, scan, obs.view(-1)[600:], cov, raw_depth=True, tri_row_stride=D)
    ops.voxelize_step(*args, torch.zeros_like(prob), torch.zeros_like(scan), dense, cov2, raw_depth=True)
    assert torch.equal(obs[:, 600:600 + G ** 3], dense.view(N, -1)) and torch.equal(cov, cov2)
    assert (obs[:, :600] == 9).all() and (obs[:, 600 + G ** 3:] == 9).all()


def test_full_size_properties():
    """BASELINE config 2 size (256 envs x 128x128 depth x 64^3 grid): invariants that do not need the oracle."""
    N, H, W, G, S = 256, 128, 128, 64, 8
    c = synthetic_case(N, H, W, G, S, 51, steps=2)
    dev = [cu(c[k]) for k in ("kinv", "range_gt", "vs", "grid_gt")]
    kinv, rg, vs, gt = dev
    prob = torch.zeros(N, G, G, G, device=DEV); scan = torch.zeros_like(prob); tri = torch.empty_like(prob)
    cov = torch.zeros(N, device=DEV); nt = torch.zeros(N, dtype=torch.int32, device=DEV)
    ws = ops.voxelize_workspace(N, G, DEV)
    prev_scan = scan.clone()
    for depth, seg, c2w, xyz in c["frames"]:
        prev_prob = prob.clone()
        a = (cu(depth), cu(seg), kinv, cu(c2w), rg, vs, cu(xyz), gt)
        ops.voxelize_step(*a, prob, scan, tri, cov, nt, workspace=ws, raw_depth=True)
        tmask, rmask = ops.voxelize_masks(ws, N, G)
        # determinism: same inputs on a cloned state give the same bits
        p2, s2, t2, c2 = prev_prob.clone(), prev_scan.clone(), torch.empty_like(tri), torch.zeros_like(cov)
        ops.voxelize_step(*a, p2, s2, t2, c2, workspace=ops.voxelize_workspace(N, G, DEV), raw_depth=True)
        assert torch.equal(p2, prob) and torch.equal(s2, scan) and torch.equal(t2, tri) and torch.equal(c2, cov)
        assert torch.equal(nt.long(), tmask.flatten(1).sum(1)) and int(nt.min()) >= 0 and int(nt.max()) > 100
        assert bool((tmask & ~rmask).sum() == 0), "every target voxel is the end point of its own ray"
        assert torch.equal(prob[tmask], torch.ones_like(prob[tmask]))
        only_r = rmask & ~tmask
        assert torch.equal(prob[only_r], prev_prob[only_r] - 0.05)
        untouched = ~rmask
        assert torch.equal(prob[untouched], prev_prob[untouched])
        assert torch.equal(tri, (prob > 0.5).float() - (prob < 0).float())
        assert torch.equal(scan, torch.clamp(prev_scan + tmask.float() * gt, 0, 1)) and bool((scan >= prev_scan).all())
        assert torch.equal(cov, scan.flatten(1).sum(1))
        prev_scan = scan.clone()
    assert float((cov / gt.flatten(1).sum(1)).mean()) > 0.05


def test_full_size_against_the_c_oracle():
    """BASELINE config 2 size (256 envs x 128x128 depth x 64^3 grid), 3 consecutive steps on carried state: every output of the
    kernel pair -- target counts, tri-class grid, prob / scanned grids, coverage sums and both bit-masks -- equals the C oracle's,
    bit for bit (the oracle needs ~1 s per step here)."""
    N, H, W, G, S = 256, 128, 128, 64, 8
    c = synthetic_case(N, H, W, G, S, 77, steps=3)
    prob_o = np.zeros((N, G, G, G), np.float32); scan_o = np.zeros_like(prob_o)
    prob_c, scan_c = prob_o.copy(), scan_o.copy()
    for t, (depth, seg, c2w, xyz) in enumerate(c["frames"]):
        o = c_oracle.voxelize_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_o, scan_o,
                                   raw_depth=True, want_masks=True)
        k = cuda_step(depth, seg, c["kinv"], c2w, c["range_gt"], c["vs"], xyz, c["grid_gt"], prob_c, scan_c, True, want_masks=True)
        for key in ("num_targets", "tri", "cov_sum", "target_mask", "touched_mask"):
            np.testing.assert_array_equal(k[key], o[key], err_msg=f"{key}, step {t}")
        np.testing.assert_array_equal(prob_c, prob_o, err_msg=f"prob_grid, step {t}")
        np.testing.assert_array_equal(scan_c, scan_o, err_msg=f"scanned_gt_grid, step {t}")
    assert int(o["num_targets"].min()) >= 0 and int(o["num_targets"].max()) > 100


@pytest.mark.parametrize("N,H,W,G", [(6, 48, 48, 20), (16, 128, 128, 64)])
def test_sparse_update_equals_dense_update(N, H, W, G):
    """gnbv_grid_update_sparse (untouched 16-byte groups not rewritten, coverage carried incrementally) leaves prob / scanned /
    tri and the coverage sums bit-identical to the dense pass, step after step on carried state."""
    c = synthetic_case(N, H, W, G, 2, 91, steps=4)
    kinv, rg, vs, gt = (cu(c[k]) for k in ("kinv", "range_gt", "vs", "grid_gt"))
    ws = ops.voxelize_workspace(N, G, DEV)
    D = 600 + G ** 3 + 8192
    state = {}
    for mode in ("dense", "sparse"):
        prob = torch.zeros(N, G, G, G, device=DEV); scan = torch.zeros_like(prob)
        obs = torch.full((N, D), 5.0, device=DEV)
        cov = torch.zeros(N, device=DEV)
        hist = []
        for depth, seg, c2w, xyz in c["frames"]:
            ops.scan_raycast(cu(depth), cu(seg), kinv, cu(c2w), rg, vs, cu(xyz), G, ws, raw_depth=True)
            ops.grid_update(gt, prob, scan, obs.view(-1)[600:], cov, ws, tri_row_stride=D, sparse=(mode == "sparse"))
            hist.append((prob.clone(), scan.clone(), obs.clone(), cov.clone()))
        state[mode] = hist
    for t, (a, b) in enumerate(zip(state["dense"], state["sparse"])):
        for x, y, name in zip(a, b, ("prob", "scanned", "obs/tri", "cov_sum")):
            assert torch.equal(x, y), f"{name}, step {t}"
    assert float(state["sparse"][-1][3].min()) > 0


def test_reset_grids():
    N, G = 5, 20
    prob = torch.rand(N, G, G, G, device=DEV); scan = torch.rand_like(prob)
    p0, s0 = prob.clone(), scan.clone()
    flags = torch.tensor([1, 0, 0, 1, 0], dtype=torch.uint8, device=DEV)
    ops.reset_grids(prob, scan, flags)
    for n in range(N):
        if flags[n]:
            assert not prob[n].any() and not scan[n].any()
        else:
            assert torch.equal(prob[n], p0[n]) and torch.equal(scan[n], s0[n])
