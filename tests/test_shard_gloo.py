"""world_size-2 gloo test of the N>1 host path: LPT assignment + final per-sequence result gather
(the only collective of the job).  Runs on CPU."""
import os
import socket
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(REPO / 'semi-supervised-vos_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from vosb200.shard import assign_lpt, gather_results, sequence_cost
    lens = [12, 40, 7, 33, 21]
    costs = [sequence_cost(n, 240) for n in lens]
    mine = assign_lpt(costs, world)[rank]
    # stand-in for the propagation: a mask stack whose content identifies (sequence, frame)
    local = {i: (torch.arange(lens[i] - 1, dtype=torch.uint8).view(-1, 1, 1) + i).expand(-1, 4, 6).contiguous()
             for i in mine}
    got = gather_results(local, dst=0)
    if rank == 0:
        assert sorted(got) == list(range(len(lens)))
        for i, n in enumerate(lens):
            want = (torch.arange(n - 1, dtype=torch.uint8).view(-1, 1, 1) + i).expand(-1, 4, 6)
            assert torch.equal(got[i], want), i
        (Path(out_dir) / 'ok').write_text('ok')
    else:
        assert got == {}
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / 'ok').read_text() == 'ok'


def test_lpt_assignment_of_the_bench_workloads():
    """bench.py's sharded workloads: every sequence lands on exactly one rank, the assignment is the same on every rank
    (pure function of the costs) and its imbalance -- the scaling ceiling the bench reports -- stays small."""
    import types
    sys.path.insert(0, str(REPO))
    sys.path.insert(0, str(REPO / 'semi-supervised-vos_b200'))
    import bench
    from vosb200.shard import assign_lpt, imbalance, sequence_cost
    for workload, frames in (('davis30', 1999), ('ytvos', None)):
        for world in (1, 2, 4, 8):
            seqs = bench.workload_sequences(types.SimpleNamespace(workload=workload, clips=4, frames=70), world)
            if frames:
                assert sum(n for n, _ in seqs) == frames and all(34 <= n <= 104 for n, _ in seqs)
            costs = [sequence_cost(n, 6420, 9) for n, _ in seqs]
            a = assign_lpt(costs, world)
            assert a == assign_lpt(costs, world)
            assert sorted(i for r in a for i in r) == list(range(len(seqs)))
            assert imbalance(costs, a) < 1.08
