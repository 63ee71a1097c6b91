"""Worker of tests/test_multigpu_gpu.py: launched under torchrun on R ranks (one per GPU, NCCL). Every rank samples its
shard of the seeds through sharding.generate_sharded and takes part in the all_gather; rank 0 then regenerates every
shard on its own GPU with the same per-shard batch and checks that the gathered images are byte-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_stable_diffusion_b200 import pipeline, sharding, synthetic  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    seeds = [int(s) for s in sys.argv[1].split(",")]
    steps = int(sys.argv[2])
    models = synthetic.build_models(dev)
    kw = dict(models=models, n_inference_steps=steps, device=dev, tokenizer=synthetic.StubTokenizer())
    images = sharding.generate_sharded("a", "b", seeds, gather=True, **kw)
    assert images.shape == (len(seeds), 512, 512, 3) and images.dtype == np.uint8, images.shape
    mine = sharding.generate_sharded("a", "b", seeds, gather=False, **kw)
    b, e = sharding.shard_range(len(seeds), rank, world)
    assert (images[b:e] == mine).all(), "own shard differs from its slot in the gathered batch"
    ok = True
    if rank == 0:
        for r in range(world):
            b, e = sharding.shard_range(len(seeds), r, world)
            if e == b:
                continue
            ref = pipeline.generate("a", "b", seeds=seeds[b:e], batch_size=e - b, return_all=True, **kw)
            same = bool((images[b:e] == ref).all())
            print(f"[sharded x{world}] shard of rank {r} (seeds {seeds[b:e]}): "
                  f"{'byte-identical to' if same else 'DIFFERS from'} the single-GPU run", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and ok:
        print("SHARDED_OK", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
