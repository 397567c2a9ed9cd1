"""N > 1 host logic on CPU: world_size-2 gloo processes exercising multimodal_lipread_b200.dp (sharding, flat
gradient allreduce with 1/world applied by the consumer, parameter broadcast, max-over-ranks timing) against the
DP oracle of SURVEY.md 8(e): the reference model run on each rank's shard with the same weights, gradients averaged."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

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
    torch.set_num_threads(2)
    from multimodal_lipread_b200 import dp, synthetic
    from oracle.av_models import MidFusionFastOracle
    from oracle.frontend import AudioProcessorPort, lips_u8_to_model_input
    r, _, w = dp.init(backend="gloo")
    assert (r, w) == (rank, world)
    B = 4                                                     # global batch, 2 clips per rank
    lo, hi = dp.shard_range(B, rank, world)
    wav = synthetic.make_waveforms(B, seed=3)[lo:hi]
    lips = synthetic.make_lips_u8(B, size=44, seed=4)[lo:hi, :6].contiguous()
    labels = synthetic.make_labels(B, 40, seed=5)[lo:hi]
    torch.manual_seed(100 + rank)                             # replicas start different ...
    model = MidFusionFastOracle(40).train()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    dp.broadcast_parameters(flat, src=0)                      # ... and are made identical by one broadcast
    off = 0
    for p in model.parameters():
        p.data.copy_(flat[off:off + p.numel()].view_as(p)); off += p.numel()
    mel = AudioProcessorPort().batch_frontend_loop(wav)
    loss = torch.nn.functional.cross_entropy(model(mel, lips_u8_to_model_input(lips)), labels)
    loss.backward()
    g_local = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    ar = dp.GradAllReduce()
    g = ar(g_local.clone()) * (1.0 / world)                   # consumer applies 1/world (the Adam kernel does)
    gathered = [torch.zeros_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    t = dp.max_over_ranks([10.0 + rank, 5.0 - rank])
    if rank == 0:
        torch.save({"g": g, "mean": torch.stack(gathered).mean(0), "t": t, "calls": ar.calls, "flat": flat}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_sharded_oracle(tmp_path):
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    d = torch.load(out)
    assert torch.allclose(d["g"], d["mean"], rtol=1e-6, atol=1e-9)      # allreduce(sum)/world == mean of shard grads
    assert d["t"] == [11.0, 5.0] and d["calls"] == 1
    torch.manual_seed(100)
    from oracle.av_models import MidFusionFastOracle
    ref = torch.cat([p.detach().reshape(-1) for p in MidFusionFastOracle(40).parameters()])
    assert torch.equal(d["flat"], ref)                                   # every replica holds rank 0's weights


def test_shard_ranges_cover_the_batch():
    from multimodal_lipread_b200 import dp
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert dp.job_throughput(32, 10, 1000.0, 8) == 32 * 8 * 10
