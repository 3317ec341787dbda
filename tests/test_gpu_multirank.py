"""Hardware multi-GPU parity: chains sharded over 2 ranks (one process per GPU, NCCL) produce, bit for bit, the run
of the same chains on one GPU -- across window closures, i.e. through the only collective of the path (the pooled
adaptation sums, klhr_b200/adaptation.py).  Skipped with fewer than 2 GPUs (`gpurun --gpus 2`)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, D, WARM, EXTRA = 8192, 24, 200, 40


def _make(kb, dev, chains, group=None, chain_offset=None, cls="KLHR"):
    model = kb.BSModel(stan_file="stan/ill-normal.stan", data={"D": D}, device=dev)
    kw = dict(seed=31, chains=chains, warmup=WARM, windowsize=25, device=dev, process_group=group, chain_offset=chain_offset)
    if cls == "KLHR_OR":
        return kb.KLHR(model, overrelaxed=True, **kw)
    return kb.KLHR(model, **kw)


def _rank_worker(rank, world, port, out, cls):
    import torch.distributed as dist
    import klhr_b200 as kb
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    s = _make(kb, dev, B // world, chain_offset=rank * (B // world), cls=cls)
    s.run(WARM + EXTRA)
    torch.save(dict(theta=s.theta.cpu(), cov=s._cov, eigvecs=s._eigvecs, eigvals=s._eigvals, mean=s._mean, K=s.K,
                    acc=s._accept_count.cpu()), f"{out}.{rank}")
    dist.destroy_process_group()


@pytest.mark.parametrize("cls", ["KLHR", "KLHR_OR"])
def test_two_ranks_nccl_equal_one_rank_bitwise(tmp_path, cls):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import klhr_b200 as kb
    out = str(tmp_path / "res")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_rank_worker, args=(2, port, out, cls), nprocs=2, join=True)
    r = [torch.load(f"{out}.{k}", weights_only=False) for k in range(2)]
    one = _make(kb, torch.device("cuda", 0), B, chain_offset=0, cls=cls)
    assert one._windowedadaptation.closures == [25, 75, 200]        # three closures = three collectives
    one.run(WARM + EXTRA)
    # adaptation state: identical on both ranks and identical to the unsharded run
    for k in ("cov", "eigvecs", "eigvals", "mean"):
        assert np.array_equal(r[0][k], r[1][k]), k
        assert np.array_equal(r[0][k], getattr(one, "_" + k)), k
    assert r[0]["K"] == r[1]["K"] == one.K
    # chains: bitwise
    assert torch.equal(torch.cat([r[0]["theta"], r[1]["theta"]]), one.theta.cpu())
    assert torch.equal(torch.cat([r[0]["acc"], r[1]["acc"]]), one._accept_count.cpu())


def test_sharded_start_points_and_adaptation_match_on_one_gpu():
    """The same invariance emulated on ONE GPU (always runs): two samplers that own halves of the chain range and
    exchange nothing start from the same points as the full sampler (start points are a function of the global chain
    id), and with the halves' adaptation sums added by hand the closure state equals the full sampler's bit for bit."""
    import klhr_b200 as kb
    from klhr_b200 import engine
    dev = torch.device("cuda", 0)
    full = _make(kb, dev, B, chain_offset=0)
    halves = [_make(kb, dev, B // 2, chain_offset=k * (B // 2)) for k in range(2)]
    assert torch.equal(torch.cat([h.theta for h in halves]), full.theta)
    full.run(12)
    for h in halves:
        h.run(12)
    assert torch.equal(torch.cat([h.theta for h in halves]), full.theta)       # before any closure: no communication needed
    planes = [h._outer_scratch for h in halves]
    outs = []
    for h, pl in zip(halves, planes):
        o = torch.zeros(D, D, dtype=torch.float64, device=dev)
        s1 = torch.zeros(D, dtype=torch.float64, device=dev)
        engine.outer_reduce(pl.clone(), o, s1, B // 2, D)
        outs.append((o, s1))
    o_full = torch.zeros(D, D, dtype=torch.float64, device=dev)
    s_full = torch.zeros(D, dtype=torch.float64, device=dev)
    engine.outer_reduce(full._outer_scratch.clone(), o_full, s_full, B, D)
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0] + outs[1][0], o_full) and torch.equal(outs[0][1] + outs[1][1], s_full)
