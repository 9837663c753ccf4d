"""Worker of tests/test_gpu_configs.py::test_two_gpus_bit_identical_to_one, launched with torch.distributed.run
(one process per GPU).  Shards 7 columns (the CaII/FALC base column and response-function columns, cycled) over
the ranks, solves each shard with the device-resident loop, gathers I and n over NCCL, and compares -- on rank 0 --
with the same batch solved on one GPU: np.array_equal, not a tolerance."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from helpers import load_golden  # noqa: E402
from lightspinner_b200.engine import MaliEngine  # noqa: E402
from lightspinner_b200.sharding import gather_columns, shard_range  # noqa: E402


def solve(problems, device):
    eng = MaliEngine(problems[0], len(problems), device=device)
    eng.upload(problems)
    eng.reset_iteration_state()
    for _ in range(8):
        eng.iterate_async(16)
        if bool((eng.t_done != 0).all().item()):
            break
    eng.raise_on_faults()
    n = len(problems)
    out = (eng.t_I.view(n, -1).clone(), eng.t_pops.view(n, -1).clone(), eng.t_J.view(n, -1).clone(),
           eng.t_iter.clone())
    eng.close()
    return out


def main():
    rank, world, local = (int(os.environ[k]) for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    names = ['c1_falc_ca', 'rf_k40p', 'rf_k10m']
    gold = [load_golden(nm)[0] for nm in names]
    ncol = 7                                         # uneven split on purpose
    problems = [gold[c % 3] for c in range(ncol)]
    lo, cnt = shard_range(ncol, world, rank)
    I, n, J, it = solve(problems[lo:lo + cnt], local)
    I_all, n_all, J_all = (gather_columns(t, ncol) for t in (I, n, J))
    it_all = gather_columns(it.view(-1, 1), ncol)
    ok = True
    if rank == 0:
        I1, n1, J1, it1 = solve(problems, local)
        ok = (np.array_equal(I_all.cpu().numpy(), I1.cpu().numpy()) and np.array_equal(n_all.cpu().numpy(), n1.cpu().numpy())
              and np.array_equal(J_all.cpu().numpy(), J1.cpu().numpy())
              and np.array_equal(it_all.cpu().numpy().ravel(), it1.cpu().numpy()))
        print('MGPU_BIT_IDENTICAL' if ok else 'MGPU_MISMATCH', 'world', world, 'iterations', it1.cpu().numpy().tolist())
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
