"""Whole-sequence sharding: the reference resets every piece of state at a video boundary
(src/utils/inference_utils.py:28-48), so the masks of a sequence must not depend on how many ranks share the job, on
which rank it lands or on what ran through that rank's engine before it.  Checked for world sizes 1, 2, 4 and 8 of the
bench's LPT assignment (each "rank" = its own engine, sequences in assignment order), and with two real processes that
propagate their shards on one GPU and gather through vosb200.shard.gather_results."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch

from oracle import propagation_oracle as O

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent

LENS = [9, 14, 6, 11, 8, 13, 7, 10, 12]
OBJS = [2, 1, 3, 2, 4, 1, 2, 3, 1]
SIZES = [(128, 288), (128, 288), (160, 320), (128, 288), (160, 320), (128, 288), (128, 288), (160, 320), (128, 288)]


def _clip(i):
    feats, first = O.synthetic_sequence(LENS[i], SIZES[i][0], SIZES[i][1], OBJS[i], seed=300 + i, feat_scale=0.30)
    return feats.half(), first


def _run_rank(indices, device='cuda'):
    """One rank's work: a fresh engine, its sequences one after the other."""
    from vosb200 import PropagationEngine
    from vosb200.sequence import propagate_clip
    eng = PropagationEngine(max_pixels=max((h // 8) * (w // 8) for h, w in SIZES), device=torch.device(device))
    out = {}
    for i in indices:
        feats, first = _clip(i)
        out[i] = propagate_clip(eng, feats.to(device), first).cpu()
    eng.close()
    return out


def test_masks_do_not_depend_on_the_world_size():
    from vosb200.shard import assign_lpt, imbalance, sequence_cost
    costs = [sequence_cost(LENS[i], (SIZES[i][0] // 8) * (SIZES[i][1] // 8)) for i in range(len(LENS))]
    want = _run_rank(range(len(LENS)))                         # world size 1
    oracle_masks, _ = O.propagate_sequence(_clip(3)[0].float(), _clip(3)[1])
    assert float((want[3].long() == oracle_masks).float().mean()) >= 0.999
    for world in (2, 4, 8):
        assignment = assign_lpt(costs, world)
        assert sorted(i for a in assignment for i in a) == list(range(len(LENS)))
        print(f'world {world}: sequences per rank {[len(a) for a in assignment]}, imbalance {imbalance(costs, assignment):.3f}')
        for rank in range(world):
            got = _run_rank(assignment[rank])
            for i, m in got.items():
                assert torch.equal(m, want[i]), f'sequence {i} differs on rank {rank} of {world}'


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / 'semi-supervised-vos_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vosb200.shard import assign_lpt, gather_results, sequence_cost
    costs = [sequence_cost(LENS[i], (SIZES[i][0] // 8) * (SIZES[i][1] // 8)) for i in range(len(LENS))]
    local = _run_rank(assign_lpt(costs, world)[rank])           # both processes share cuda:0; results come back on the host
    got = gather_results(local, dst=0)
    if rank == 0:
        torch.save(got, Path(out_dir) / 'gathered.pt')
    dist.barrier()
    dist.destroy_process_group()


def test_two_processes_shard_propagate_and_gather(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = torch.load(tmp_path / 'gathered.pt')
    want = _run_rank(range(len(LENS)))
    assert sorted(got) == list(range(len(LENS)))
    for i in want:
        assert torch.equal(got[i], want[i]), i
