"""Multi-GPU check, launched by torchrun (one process per GPU): the barcode-sharded run with the
per-step NCCL all-reduce must reproduce the single-GPU run with the same seed up to reduction
order.  Prints one JSON line on rank 0.  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_gpu_check.py [model]
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import barbay_b200 as bb  # noqa: E402

CFG = {"fitness_normal": 2, "replicate_fitness_normal": 3, "multienv_fitness_normal": 4, "genotype_fitness_normal": 5}


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "fitness_normal"
    dtype = sys.argv[2] if len(sys.argv) > 2 else "f64"
    wiring = sys.argv[3] if len(sys.argv) > 3 else "peer"        # peer: bb_peer_attach (no NCCL in the library) | nccl
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _, da, _ = bb.synth.config(CFG[model], scale=0.003)
    K, steps = 4, 6

    def run(r, w):
        eng = bb.Engine(da, model, n_samples=K, dtype=dtype, seed=11, device=local, rank=r, world=w)
        if w > 1 and wiring == "nccl":
            uid = [bb.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            eng.comm_init(uid[0])
        elif w > 1:
            hs = [None] * w
            dist.all_gather_object(hs, eng.peer_handle())
            eng.peer_attach(hs)
        eng.init_params(5)
        eng.set_optimizer("truncated", n=4)
        trace = eng.step(steps, elbo_trace=True)
        eng.step(steps)            # production path: fused step kernel + merged tail kernel (peer exchange inside)
        m, s = eng.get_posterior()
        eng.close()
        return trace, m, s

    trace_d, m_d, s_d = run(rank, world)
    # every rank reports only the latents it owns (zeros elsewhere): sum across ranks
    t = torch.from_numpy(np.stack([m_d, s_d])).cuda()
    dist.all_reduce(t)
    m_d, s_d = t[0].cpu().numpy(), t[1].cpu().numpy()
    if rank == 0:
        trace_1, m_1, s_1 = run(0, 1)
        out = {"model": model, "world": world, "dtype": dtype, "wiring": wiring,
               "elbo_rel": float(np.max(np.abs(trace_d - trace_1) / np.abs(trace_1))),
               "mean_rel": float(np.max(np.abs(m_d - m_1)) / np.max(np.abs(m_1))),
               "std_rel": float(np.max(np.abs(s_d - s_1)) / np.max(np.abs(s_1))),
               "all_owned": bool((s_d > 0).all())}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
