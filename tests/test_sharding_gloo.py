"""N > 1 path on CPU: two processes (torch.distributed, gloo, world_size 2) each align the batches
b % 2 == rank of the same input -- here through the CPU unit harness, on the GPU box through the product
library -- and rank 0 merges the parts; the merged SAM must equal the single-process reference output.
No collective touches the data path: the only exchange is the barrier before the merge."""
import gzip
import os
import subprocess
import sys

import pytest

from conftest import ROOT, first_diff, strip_pg

WORKER = r'''
import os, subprocess, sys
import torch.distributed as dist
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo')
out_dir, hostsim = sys.argv[1], sys.argv[2]
argv = sys.argv[3:]
sys.path.insert(0, os.environ['BSB_ROOT'])
from bsbolt_b200.shard import shard_env, merge_shards
env = shard_env(rank, world, f'{out_dir}/part{rank}.idx')
with open(f'{out_dir}/part{rank}.sam', 'w') as fo, open(f'{out_dir}/part{rank}.log', 'w') as fl:
    rc = subprocess.run([hostsim] + argv, stdout=fo, stderr=fl, env=env).returncode
assert rc == 0
dist.barrier()
if rank == 0:
    with open(f'{out_dir}/merged.sam', 'wb') as o:
        merge_shards([f'{out_dir}/part{r}.sam' for r in range(world)], [f'{out_dir}/part{r}.idx' for r in range(world)], o)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.parametrize('case', ['pe150_un', 'se100'])
def test_two_rank_sharded_run_equals_single_run(built, golden, tmp_path, case):
    worker = tmp_path / 'worker.py'
    worker.write_text(WORKER)
    hostsim = os.path.join(ROOT, 'tests', 'hostsim', 'hostsim')
    env = dict(os.environ, BSB_ROOT=ROOT)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29517', str(worker), str(tmp_path), hostsim] + golden.argv(case)
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    merged = strip_pg(open(tmp_path / 'merged.sam').read())
    want = golden.sam(case)
    assert merged == want, first_diff(want, merged)
    # both ranks really had work (the cases are cut into several batches by their -K)
    n0 = sum(1 for _ in open(tmp_path / 'part0.idx')); n1 = sum(1 for _ in open(tmp_path / 'part1.idx'))
    assert n0 >= 2 and n1 >= 1


FENCE_WORKER = r'''
import os, sys, time
sys.path.insert(0, os.environ['BSB_ROOT'])
os.environ['BSB_BENCH_BACKEND'] = 'gloo'
import bench
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
r = bench.Ranks(rank, world, int(os.environ['LOCAL_RANK']))
out = open(os.path.join(sys.argv[1], f'fence{rank}.log'), 'w')
for k in range(4):
    if rank == 0:
        time.sleep(0.4)                      # rank 0's "work" between two fences
        out.write(f'{k} {time.time()}\n')    # ... finished before it enters fence k
    r.fence()
    if rank != 0:
        out.write(f'{k} {time.time()}\n')    # a waiting rank leaves fence k
out.close()
r.close()
'''


def test_bench_ranks_wait_for_rank0_at_every_fence(tmp_path):
    """bench.py under torchrun: rank 0 drives every GPU through the product's one-process path, the other ranks only meet
    it at the fences around the timed regions (gloo here, NCCL on the GPU box). A waiting rank must never leave fence k
    before rank 0 has reached it."""
    worker = tmp_path / 'fence_worker.py'
    worker.write_text(FENCE_WORKER)
    env = dict(os.environ, BSB_ROOT=ROOT)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29519', str(worker), str(tmp_path)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    t0 = [float(l.split()[1]) for l in open(tmp_path / 'fence0.log')]
    t1 = [float(l.split()[1]) for l in open(tmp_path / 'fence1.log')]
    assert len(t0) == len(t1) == 4
    assert all(b >= a for a, b in zip(t0, t1))
