"""CPU: the N>1 host logic (sharding, gathers, reductions) on a world_size-2 gloo group, plus scheduler/host logic."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_image_editing_with_generative_models_b200 import sweep
from fast_image_editing_with_generative_models_b200.scheduler import LCMSchedule

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = sweep.init_distributed("gloo")
    assert (r, w) == (rank, world)
    entries = list(range(n_items))
    mine = sweep.shard(entries, r, w)
    assert mine == entries[r::w]
    # each "image" is a tiny uint8 tensor tagged with its global index
    local = torch.stack([torch.full((2, 3), i, dtype=torch.uint8) for i in mine]) if mine else torch.zeros((0, 2, 3), dtype=torch.uint8)
    allv = sweep.gather_outputs(local)
    assert allv.shape[0] == n_items
    assert [int(allv[i, 0, 0]) for i in range(n_items)] == entries          # original order restored
    assert sweep.max_over_ranks(float(r + 1), torch.device("cpu")) == float(w)
    assert sweep.sum_over_ranks(float(len(mine)), torch.device("cpu")) == float(n_items)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_world2_gloo_shard_and_gather(n_items):
    mp.spawn(_worker, args=(2, _free_port(), n_items), nprocs=2, join=True)


def test_shard_balance_and_ownership():
    entries = list(range(700))
    shards = [sweep.shard(entries, r, 8) for r in range(8)]
    assert sorted(sum(shards, [])) == entries
    assert max(map(len, shards)) - min(map(len, shards)) <= 1          # 700 = 87.5 x 8 -> 88/87
    assert all(sweep.owner_of(i, 8) == r for r, sh in enumerate(shards) for i in sh)
    assert sweep.dist_env()[1] >= 1


def test_lcm_schedule_matches_oracle_and_known_answers():
    from oracle.diffusion_oracle import LCMSchedule as OracleSchedule
    s, o = LCMSchedule(), OracleSchedule()
    assert s.set_timesteps(4) == [999, 759, 499, 259] == o.set_timesteps(4)
    assert s.img2img_timesteps(4, 0.5) == ([499, 259], 2) and s.img2img_timesteps(4, 0.8) == ([759, 499, 259], 1)
    assert s.img2img_timesteps(4, 1.0)[0] == [999, 759, 499, 259] and s.img2img_timesteps(4, 0.1) == ([], 4)
    for t, a in [(999, 0.004660), (759, 0.052213), (499, 0.277669), (259, 0.658975)]:
        assert abs(float(s.alphas_cumprod[t]) - a) < 1e-5
        assert abs(float(s.alphas_cumprod[t]) - float(o.alphas_cumprod[t])) < 1e-6
    s.img2img_timesteps(4, 0.5); o.img2img_timesteps(4, 0.5)
    for i in (2, 3):
        a, b = s.step_coeffs(i), o.step_coeffs(i)
        for k in a:
            assert abs(float(a[k]) - float(b[k])) < 1e-6, (i, k)
    assert abs(s.add_noise_coeffs(499)[0] - 0.526944) < 1e-5 and abs(s.add_noise_coeffs(499)[1] - 0.849900) < 1e-5
