"""torchrun check of picasso_b200.distributed.PeerGather against an NCCL all-gather:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peer_gather.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from picasso_b200 import _lib
    from picasso_b200.distributed import PeerGather

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    _lib.check(_lib.load().pb_set_device(local))
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 3_000_001                                     # odd size, 12 MB blocks
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    pg = PeerGather(dist, torch, n * 4, dev)
    ok = True
    for rep in range(3):
        block = torch.rand(n, generator=g, device=dev)
        want = torch.empty(n * world, device=dev)
        dist.all_gather_into_tensor(want, block)
        ev = pg.gather_async(block)
        PeerGather.wait(ev, torch.cuda.current_stream())
        pg.finish()
        ok = ok and bool(torch.equal(pg.to_tensor(torch.float32), want))
        # partial gather at an offset
        part = block[1000:5000].contiguous() + 1.0
        PeerGather.wait(pg.gather_async(part, offset_bytes=4000), torch.cuda.current_stream())
        pg.finish()
        got = pg.to_tensor(torch.float32).view(world, n)
        allparts = torch.empty(4000 * world, device=dev)
        dist.all_gather_into_tensor(allparts, part)
        ok = ok and bool(torch.equal(got[:, 1000:5000].contiguous().view(-1), allparts))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    pg.close()
    if rank == 0:
        print(json.dumps({"world": world, "peer_gather_equals_nccl": bool(flag.item() == 1.0)}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
