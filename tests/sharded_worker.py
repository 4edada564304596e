"""One rank of the sharded-proof GPU test (tests/test_gpu_sharded.py): `world` processes share GPU 0 and talk over gloo
(TorchComm's host-staged mode — NCCL refuses two ranks on one device); every rank proves its column block with
wf_prove_fib_sharded and rank 0 compares the bytes with the single-GPU proof of the whole trace."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    k, log_n, ext, hash_id, resident = (int(x) for x in sys.argv[1:6])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import winterfell_b200 as wf
    from winterfell_b200 import dist as wd
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    ctx = wf.Context(0, stream.cuda_stream)
    n = 1 << log_n
    trace, results = wf.build_fib_trace(k, n)
    opts = np.array([24, 8, 6, ext, 4, 31, 0, 0, hash_id], dtype=np.uint32)
    cl = 2 * k // world
    local = np.ascontiguousarray(trace[rank * cl:(rank + 1) * cl])
    comm = wd.TorchComm(stream)
    stats = {}
    with torch.cuda.stream(stream):
        if resident:
            dev = torch.from_numpy(local.view(np.int64)).cuda()
            proof = wd.prove_fib_sharded(ctx, comm, None, k, log_n, results, opts, device_ptr=dev.data_ptr(), stats=stats)
        else:
            proof = wd.prove_fib_sharded(ctx, comm, local, k, log_n, results, opts, stats=stats)
        ok = True
        if rank == 0:
            want = ctx.prove_fib(trace, results, opts)
            ok = proof == want
            print(f"sharded proof {len(proof)} bytes, single-GPU {len(want)} bytes, equal={ok}, stats={stats}", flush=True)
        # every rank must hold the same bytes
        digest = torch.frombuffer(bytearray(__import__("hashlib").sha256(proof).digest()), dtype=torch.uint8)
        all_d = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(all_d, digest)
        ok = ok and all(bool((d == all_d[0]).all()) for d in all_d)
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
