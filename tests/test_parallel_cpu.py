"""Host logic of the data-parallel path on CPU: world_size-2 gloo processes (no GPU, no kernels).  Checks the bucket
plan, the overlapped reduce_ready/flush protocol, the post-backward fallback and frame sharding."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pcss-unet_b200"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "pcss-unet_b200"))
    from Unetmodel import Unet
    from parallel import GradSync, bucket_plan
    torch.manual_seed(rank)              # deliberately different initial weights per rank
    net = Unet()
    sync = GradSync(net)
    sync.broadcast_state(0)
    w_sum = float(sum(p.double().sum() for p in net.parameters()))
    names = [n for n, _ in net.named_parameters()]
    plan = bucket_plan(names)
    # fake per-rank gradients handed over block by block in backward order, as nsm_train._backward does
    grads = {n: torch.full_like(p, float(rank + 1)) * (i + 1) for i, (n, p) in enumerate(net.named_parameters())}
    order = ["conv10", "conv9", "conv8", "conv7", "conv6", "conv5", "conv4", "conv3", "conv2"]
    for blk in order:
        sync.reduce_ready({n: g for n, g in grads.items() if n.split(".")[0] == blk})
    sync.flush()
    ok_overlap = all(torch.allclose(grads[n], torch.full_like(grads[n], 1.5 * (i + 1))) for i, n in enumerate(names))
    # fallback path on .grad
    for i, (n, p) in enumerate(net.named_parameters()):
        p.grad = torch.full_like(p, float(rank + 1)) * (i + 1)
    sync.synced_in_backward = False
    sync.finish()
    ok_finish = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, (n, p) in enumerate(net.named_parameters()))
    q.put((rank, w_sum, ok_overlap, ok_finish, [len(b) for b in plan]))
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, o0, f0, plan0), (r1, w1, o1, f1, plan1) = res
    assert w0 == w1                         # broadcast_state made the replicas identical
    assert o0 and o1 and f0 and f1          # mean of the two ranks' gradients everywhere, both paths
    assert plan0 == plan1 and sum(plan0) == 66 and len(plan0) == 6
    assert plan0[3] == 1                    # conv6.conv.0.weight travels alone (37.7 MB, 60 % of the payload)


def test_shard_frames():
    from parallel import shard_frames
    for n, w in ((16, 8), (5, 4), (3, 8), (1, 1)):
        got = [list(shard_frames(n, r, w)) for r in range(w)]
        assert sorted(i for g in got for i in g) == list(range(n))
        assert max(len(g) for g in got) - min(len(g) for g in got) <= 1
    assert list(shard_frames(16, 3, 8)) == [6, 7]
