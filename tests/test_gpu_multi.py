"""Data-parallel training on >= 2 GPUs (skipped on a single-GPU box): N-rank gradients from the overlapped, bucketed NCCL
all-reduce equal the average of the per-shard single-process gradients (per-replica BatchNorm statistics, replayed
Dropout2d masks) -- SURVEY 8e."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _masks(seed, N):
    sys.path.insert(0, ROOT)
    import oracle
    torch.manual_seed(seed)
    return [torch.empty(N, cin, 1, 1).bernoulli_(1 - oracle.DROPOUT_P(name)).div_(1 - oracle.DROPOUT_P(name))
            for name, cin, _ in oracle.BLOCKS]


def _step(rank, x, t, masks, sync_cls=None):
    from Unetmodel import Unet
    from customLoss import CustomLoss
    torch.manual_seed(42)
    net = Unet(dropout_rate=0.2, precision="fp32").cuda().train()
    import nsm_train
    nsm_train.replay_masks(net, masks)
    sync = sync_cls(net) if sync_cls is not None else None
    out = net(x.cuda())
    CustomLoss("cuda")(out, t.cuda(), None).backward()
    if sync is not None:
        sync.finish()
    return {n: p.grad.detach().cpu() for n, p in net.named_parameters()}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    for p in (os.path.join(ROOT, "pcss-unet_b200"), ROOT):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from parallel import GradSync
    g = torch.Generator().manual_seed(10 + rank)
    x, t = torch.randn(2, 4, 64, 96, generator=g), torch.rand(2, 1, 64, 96, generator=g)
    grads = _step(rank, x, t, _masks(20 + rank, 2), GradSync)
    q.put((rank, {k: v.numpy() for k, v in grads.items()}))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp2_gradients_equal_mean_of_shards():
    import torch.multiprocessing as mp
    for p in (os.path.join(ROOT, "pcss-unet_b200"), ROOT):
        sys.path.insert(0, p)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-process evaluation of each shard, then the mean
    torch.cuda.set_device(0)
    shard = []
    for r in range(2):
        g = torch.Generator().manual_seed(10 + r)
        x, t = torch.randn(2, 4, 64, 96, generator=g), torch.rand(2, 1, 64, 96, generator=g)
        shard.append(_step(0, x, t, _masks(20 + r, 2)))
    for n in shard[0]:
        ref = (shard[0][n] + shard[1][n]) / 2
        for r in range(2):
            got = torch.from_numpy(res[r][n])
            assert torch.allclose(got, ref, rtol=1e-4, atol=1e-8), (n, r, (got - ref).abs().max().item())
