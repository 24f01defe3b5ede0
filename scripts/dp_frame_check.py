"""Data-parallel check of the FRAME path on N GPUs (torchrun): 7 videos sharded over the ranks (padded to equal step counts),
two eager epochs then two CUDA-graph epochs; every rank must finish and hold bit-identical parameters (one gradient
all-reduce per video keeps the replicas in step).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_frame_check.py
"""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from multimodal_error_detection_b200 import parallel, synthetic  # noqa: E402
from multimodal_error_detection_b200.dataset.CustomFrameDataset import CustomFrameDataset, FrameLoader  # noqa: E402
from multimodal_error_detection_b200.modeling import modeling_utils as mu  # noqa: E402


def main():
    rank, local_rank, world = parallel.init_from_env()
    dev = torch.device("cuda", local_rank)
    path = os.path.join(tempfile.gettempdir(), f"b200med_dp_fold_{os.environ.get('MASTER_PORT', '0')}_{rank}")
    synthetic.write_fold(synthetic.make_fold(seed=42, n_train=7, n_test=2, t_lo=150, t_hi=320), path)   # same fold on every rank
    for graph in (False, True):
        kw = dict(cases.FRAME_EPOCH_CASES["tecno_multimodal"], cuda_graph=graph)
        ds = CustomFrameDataset(path + "/", csv_filename="train.csv", delete_ND=kw["delete_ND"])
        loader = FrameLoader(ds, shuffle=True, generator=torch.Generator().manual_seed(42), rank=rank, world_size=world)
        fe, model, crit, opt, sched = mu.define_model_objects(kw, cases.IN_FEATURES, dev, (0.4, 0.6), 0)
        losses = [mu.train_single_epoch(model, fe, loader, crit, opt, sched, dev, kw)[0] for _ in range(2)]
        flat = torch.cat([p.detach().reshape(-1) for p in list(fe.parameters()) + list(model.parameters())])
        digest = torch.stack([flat.double().sum(), flat.double().abs().sum()])
        got = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(got, digest)
        same = all(torch.equal(g, got[0]) for g in got)
        steps = len(list(loader.indices()))
        if graph:
            assert not opt._b200_frame_steps["failed"]
        if rank == 0:
            print(f"dp frame check: world {world}, graph={graph}, steps per rank {steps}, losses {losses}, replicas identical: {same}", flush=True)
        assert same, (graph, got)
        if graph:
            opt._b200_frame_steps = None          # graphs that captured NCCL work go before the communicator does
    torch.cuda.synchronize()
    parallel.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
