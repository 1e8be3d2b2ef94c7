"""Worker of tests/test_sharded.py::test_two_gpus_nccl_gather (launched with torchrun, one rank per GPU): the sharded
similarity call with gather=True -- work-balanced row ranges, the kernel writing into its slice of the gather buffer,
one in-place NCCL all-gather per slab array -- must reproduce the oracle's full matrix on EVERY rank."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    import similaripy_b200 as sim
    from similaripy_b200 import sharded
    from oracle import oracle
    from parity import assert_topk_parity, random_csr

    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    try:
        urm = random_csr(3000, 1200, 0.03, 31)
        # skewed work: the first items are ten times as popular, so equal-work ranges have unequal row counts
        boost = random_csr(3000, 1200, 0.3, 32)
        boost = boost[:, :100]
        import scipy.sparse as sp
        urm = sp.hstack([urm[:, :100] + boost, urm[:, 100:]]).tocsr().astype(np.float32)
        a = urm.T.tocsr()
        for name, kw, fmt, rows in (("cosine", {}, "csr", None), ("rp3beta", dict(alpha=0.9, beta=0.4), "coo", None),
                                    ("dot_product", {}, "csr", np.arange(1199, -1, -3, dtype=np.int32))):
            with sharded.shard_rows(gather=True):
                got = getattr(sim, name)(a, k=25, target_rows=rows, verbose=False, format_output=fmt, device=local, **kw)
            ref = oracle.similarity(name, a, k=25, target_rows=rows, verbose=False, format_output="csr", **kw)
            assert_topk_parity(ref, got, k=25, rtol=1e-5, what=f"rank {rank}: sharded {name}")
            with sharded.shard_rows(gather=False):
                part = getattr(sim, name)(a, k=25, target_rows=rows, verbose=False, format_output="csr", device=local, **kw)
            n_local = int((np.diff(part.indptr) > 0).sum())
            t = torch.tensor([n_local], device="cuda")
            dist.all_reduce(t)
            assert int(t.item()) == int((np.diff(ref.indptr) > 0).sum()), "the ranks' row ranges must partition the target rows"
            assert 0 < n_local < int(t.item())
        # on-device result, full matrix on every rank
        with sharded.shard_rows(gather=True):
            dm = sim.cosine(sim.to_device(a, device=local), k=25, verbose=False, on_device=True, device=local)
        ref = oracle.similarity("cosine", a, k=25, verbose=False, format_output="csr")
        assert_topk_parity(ref, sim.to_host(dm), k=25, rtol=1e-5, what=f"rank {rank}: sharded on-device")
        dist.barrier()
        print(f"SHARDED_OK rank={rank} world={world}", flush=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
