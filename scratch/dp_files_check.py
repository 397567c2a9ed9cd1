"""2-rank sanity check of file-fed data-parallel training (run under torchrun on a 2-GPU box):
files -> DeviceBatchLoader(rank, world) -> MidFusionFast.train_step with the in-graph NCCL allreduce, two epochs.
Checks: both ranks take the same number of steps on disjoint clips, replicas stay bit-identical, the loss falls.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scratch/dp_files_check.py
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_lipread_b200 import data, dp, synthetic, train as T          # noqa: E402
from multimodal_lipread_b200.audio_video_models import create_mid_fusion_fast  # noqa: E402
from multimodal_lipread_b200.model_base import Cfg                            # noqa: E402


def main():
    rank, local_rank, world = dp.env_rank_world()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dp.init(device=dev)
    root = os.path.join(tempfile.gettempdir(), "lipread_dp_files", "GLips_4")
    if rank == 0:
        synthetic.write_dataset_tree(root, per_split={"train": 16}, T=8, size=44, missing_every=1000, audio_ext=".npy")
    dist.barrier()
    ds = data.GLipsMultimodalDataset(root, 117, "train", audio_ext=".npy")
    loader = data.DeviceBatchLoader(ds, 8, shuffle=True, device=dev, seed=7, rank=rank, world=world)
    torch.manual_seed(0)
    model = create_mid_fusion_fast(3, Cfg()).to(dev)
    model.configure_optimizer(lr=1e-3)
    seen, losses = [], []
    orig = loader._stage

    def spy(slot, idxs):
        seen.extend(idxs)
        return orig(slot, idxs)
    loader._stage = spy
    for epoch in range(3):
        loss, acc = T.train_epoch(model, loader, dev, grad_allreduce=dp.GradAllReduce(), world=world)
        losses.append(loss)
    torch.cuda.synchronize()
    flat = model._flat.flat
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, flat))
    counts = torch.tensor([len(seen)], device=dev)
    gathered = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(gathered, counts)
    mine = torch.zeros(len(ds), device=dev)
    mine[torch.tensor(seen[:len(seen) // 3], device=dev)] = 1                 # first epoch's clips of this rank
    total = mine.clone()
    dist.all_reduce(total)
    mean_losses = torch.tensor(losses, device=dev, dtype=torch.float64)
    dist.all_reduce(mean_losses)
    mean_losses = (mean_losses / world).tolist()                              # a rank's own shard loss is noisy
    ok = same and len({int(c) for c in gathered}) == 1 and float(total.max()) == 1.0 and mean_losses[-1] < mean_losses[0]
    line = {"rank": rank, "replicas_identical": same, "steps_per_rank": [int(c) // 8 for c in gathered],
            "first_epoch_overlap_max": float(total.max()), "first_epoch_clips": int(total.sum()), "n_clips": len(ds),
            "epoch_losses": losses, "epoch_losses_mean_over_ranks": mean_losses, "ok": ok}
    if rank == 0:
        print(json.dumps(line))
    model._graphs.clear()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
