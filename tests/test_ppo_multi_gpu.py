"""2 x B200 (skipped with fewer GPUs): `PPO_Grid_Obs.train()` under NCCL -- the overlapped two-piece gradient all-reduce, the
KL-stop vote carried by it, the BatchNorm buffer averaging -- exercised end to end (tests/multi_gpu_worker.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def run_worker(scenario, nproc, port):
    cmd = [sys.executable]
    if nproc > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
                "--master-port", str(port)]
    cmd += [os.path.join(HERE, "multi_gpu_worker.py"), scenario]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    return json.loads(line[7:])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_train_equals_single_gpu_on_identical_shards():
    one = run_worker("same", 1, 0)[0]
    two = run_worker("same", 2, 29611)
    assert one["adam_step"] == 8 and one["stopped_epoch"] is None
    for r in two:
        assert r == one, "a 2-rank run on identical shards must reproduce the single-GPU update bit for bit"


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_kl_stop_vote_of_one_rank_stops_every_rank():
    two = run_worker("vote", 2, 29613)
    for r in two:
        assert r["stopped_epoch"] == 0 and r["adam_step"] == 0 and r["logged"] == 1, r
    # weights untouched and equal; BN running buffers averaged over the ranks at the end of train() -> equal
    assert two[0]["params"] == two[1]["params"] and two[0]["bn"] == two[1]["bn"]
    # sanity: alone, rank 0's data does not trigger the stop at the first minibatch
    alone = run_worker("vote", 1, 0)[0]
    assert alone["adam_step"] >= 1
