"""N-rank check (torchrun) of the NVSwitch multicast buffer and the fused fit + all-gather:
  * MulticastBuffer.copy_in: every rank pushes its block through the multicast mapping -> every rank's
    local copy equals an NCCL all-gather of the blocks;
  * pb_mle_fit_gather_dev: the gather buffer written by the fit kernel equals the NCCL all-gather of the
    packed outputs [thetas 6n | crlbs 6n | logliks n | iterations n], bit for bit, twice (buffer reuse).
One JSON line on rank 0; {"supported": false} when the box has no multicast support."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import bench
    from picasso_b200 import _lib
    from picasso_b200.distributed import MulticastBuffer

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    lib = _lib.load()
    _lib.check(lib.pb_set_device(local))
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}
    n = 200_000
    try:
        mcb = MulticastBuffer(dist, torch, 14 * n * 4, dev)
    except RuntimeError as exc:
        if rank == 0:
            print(json.dumps({"world": world, "supported": False, "why": str(exc)}), flush=True)
        dist.destroy_process_group()
        return
    out["supported"] = True

    def all_equal(flag):
        t = torch.tensor([1.0 if flag else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    # ---- plain multicast copy ----
    blk = torch.arange(14 * n, dtype=torch.float32, device=dev) + 1e6 * rank
    ref = torch.empty(14 * n * world, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(ref, blk)
    mcb.copy_in(blk)
    torch.cuda.synchronize(); dist.barrier()
    out["copy_equals_nccl"] = all_equal(torch.equal(mcb.local(torch.float32), ref))
    # ---- fused fit + gather ----
    lib.pb_mle_fit_gather_dev.argtypes = [C.c_size_t, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_int] + [C.c_void_p] * 7
    lib.pb_mle_fit_gather_dev.restype = C.c_int
    st = torch.cuda.current_stream()
    oks = []
    for rep in range(2):
        spots = bench.gen_spots_device(torch, n, 7, 50 + 10 * rep + rank, dev)
        flat = torch.empty(14 * n, dtype=torch.float32, device=dev)
        th, cr, ll, it = flat[:6 * n], flat[6 * n:12 * n], flat[12 * n:13 * n], flat[13 * n:].view(torch.int32)
        _lib.check(lib.pb_mle_fit_gather_dev(n, 7, spots.data_ptr(), 0.001, 100, 1, th.data_ptr(), cr.data_ptr(),
                                             ll.data_ptr(), it.data_ptr(), None, mcb.block_mc_ptr(), st.cuda_stream))
        torch.cuda.synchronize(); dist.barrier()
        dist.all_gather_into_tensor(ref, flat)
        oks.append(all_equal(torch.equal(mcb.local(torch.int32), ref.view(torch.int32))))
        dist.barrier()
    out["fused_fit_gather_equals_nccl"] = oks
    mcb.close()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
