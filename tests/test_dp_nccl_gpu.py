"""The data-parallel train step on the HARDWARE path against the DP oracle of SURVEY.md 8(e): two NCCL ranks (one
process per GPU), two clips per rank, PlanModel.train_step(grad_allreduce=..., world=2) -- eager collective, collective
captured inside the step graph, and the LIPREAD_ALLREDUCE_IN_GRAPH=0 split-graph fallback -- compared with "the oracle
model on each shard separately, gradients averaged" and with torch's Adam applied to the averaged gradient.
Needs two GPUs (NCCL refuses two ranks on one device); skipped otherwise.  `bench.py --gpus N` prints the same check
as its `dp_parity` field."""
import json
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    import bench_checks
    res = {"in_graph": bench_checks.dp_parity(dev, rank, world)}
    os.environ["LIPREAD_ALLREDUCE_IN_GRAPH"] = "0"
    res["split_graph"] = bench_checks.dp_parity(dev, rank, world)
    if rank == 0:
        with open(out, "w") as f:
            json.dump(res, f)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)                                   # graphs with captured NCCL kernels: leave before teardown order matters


def test_two_rank_nccl_step_matches_sharded_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs: one NCCL rank per device")
    out = str(tmp_path / "dp.json")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = json.load(open(out))
    for variant in ("in_graph", "split_graph"):
        for mode in ("eager", "graph"):
            r = res[variant][mode]
            assert r["grad_max_rel"] <= 3e-3, (variant, mode, r)
            assert r["adam_weights_max_abs"] <= 2e-7 + 1e-5 * 3e-4, (variant, mode, r)
