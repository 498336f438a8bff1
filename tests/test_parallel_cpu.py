"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic (rotmv_b200/parallel.py): batch
sharding, gradient averaging == global-batch gradient, max-over-ranks timing, buffer broadcast.
The per-rank gradients come from the CPU oracle (the GPU engine is not involved here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "rot-mvgaze_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import rotmv_oracle as O
    from rotmv_b200 import parallel as P

    # identical replicas (same seed), eval-mode BatchNorm so that the loss is a plain batch mean
    model = O.build_model(num_iter=2, depth=18, seed=0).eval()
    images, pose, gt = O.synthetic_batch(5, 2, seed=3, size=64)   # 5 samples over 2 ranks: 3 + 2
    rot = O.pairwise_rotations(pose)
    b, e = P.shard_range(5, rank, world)
    out = model.forward_views(images[b:e], rot[b:e])
    # weight each rank's local mean by its share so that the average of ranks is the global mean
    loss = O.iteration_loss(out, [gt[b:e, 0], gt[b:e, 1]]) * ((e - b) * world / 5.0)
    loss.backward()
    params = [p for p in model.parameters() if p.grad is not None]
    flat = torch.cat([p.grad.flatten() for p in params])
    # bucketed, overlapped all-reduce in the order the training step completes its gradient buckets
    # (fusion stage, layer4, layer3, layer2..stem): later buckets are still being written while the
    # earlier collectives run
    names, offs, total = P.flat_layout(model.named_parameters())
    named = dict(model.named_parameters())
    flat2 = torch.zeros(total)
    for n, o in zip(names, offs):
        flat2[o:o + named[n].numel()] = named[n].grad.flatten()
    expect = flat2.clone()
    buckets = P.gradient_buckets(names, offs, total)
    ar = P.OverlappedAllReduce(flat2)
    assert ar.active
    for i, (_, b0, e0) in enumerate(buckets):
        flat2[b0:e0] += float(i)              # the "backward" finishing this bucket just before its reduce
        expect[b0:e0] += float(i)
        ar.start(b0, e0)
    ar.finish()
    flat2.mul_(1.0 / world)
    P.allreduce_mean_(expect)
    overlap_ok = torch.allclose(flat2, expect, rtol=1e-6, atol=1e-7)
    # replica synchronisation at start-up: rank 0's state everywhere
    state = [torch.full((7,), float(rank + 1)), torch.full((3,), 10.0 * (rank + 1), dtype=torch.float64)]
    P.broadcast_state_(state, None, src=0)
    overlap_ok = overlap_ok and float(state[0][0]) == 1.0 and float(state[1][0]) == 10.0
    P.allreduce_mean_(flat)
    t = P.max_over_ranks(1.0 + rank)
    bn = next(m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d))
    bn.running_mean.fill_(float(rank + 1))
    P.broadcast_buffers_(model, src=0)
    torch.save({"flat": flat, "range": (b, e), "tmax": t, "rm": bn.running_mean.clone(),
                "overlap_ok": overlap_ok},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_buckets_partition_the_flat_buffer():
    """parallel.flat_layout / gradient_buckets on the real parameter tree: the four buckets are
    contiguous, disjoint, cover [0, total) and come in backward-completion order."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "rot-mvgaze_b200"))
    from rotmv_b200 import parallel as P
    from rotmv_b200.module import FeatRotationSymm

    for depth in (50, 18):
        m = FeatRotationSymm(depth, 3)
        names, offs, total = P.flat_layout(m.named_parameters())
        assert not any(n.startswith("_feat_extractor.0.fc.") for n in names)           # Q4
        assert all(o % 64 == 0 for o in offs) and total % 64 == 0
        b = P.gradient_buckets(names, offs, total)
        assert [x[0] for x in b] == ["fusion", "layer4", "layer3", "layer2-stem"]
        assert b[0][2] == total and b[3][1] == 0
        assert all(b[i][1] == b[i + 1][2] for i in range(3))                            # contiguous, descending
        assert all(e > s for _, s, e in b)
        where = dict(zip(names, offs))
        assert b[1][1] == where["_feat_extractor.0.layer4.0.conv1.weight"]
        assert b[2][1] == where["_feat_extractor.0.layer3.0.conv1.weight"]
        assert b[0][1] == where["_lifter._lifter.blocks.0.0.weight"]
        if depth == 50:
            assert sum(p.numel() for n, p in m.named_parameters() if n in where) == 89591366


def test_shard_range_covers_batch():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "rot-mvgaze_b200"))
    from rotmv_b200 import parallel as P

    for batch in (0, 1, 5, 8, 128, 1000):
        for world in (1, 2, 3, 8):
            got = [P.shard_range(batch, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == batch
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_range(8, 2, 2)


@pytest.mark.timeout(600)
def test_gloo_two_ranks_gradient_average_equals_global_batch(tmp_path):
    from oracle import rotmv_oracle as O

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert r0["range"] == (0, 3) and r1["range"] == (3, 5)
    assert torch.equal(r0["flat"], r1["flat"])          # every rank holds the same averaged gradient
    assert r0["tmax"] == r1["tmax"] == 2.0              # slowest rank wins
    assert r0["overlap_ok"] and r1["overlap_ok"]        # bucketed overlapped all-reduce == one all-reduce; replica sync
    assert torch.equal(r0["rm"], r1["rm"]) and float(r1["rm"][0]) == 1.0   # rank 0's buffers
    # single-process gradient on the global batch
    model = O.build_model(num_iter=2, depth=18, seed=0).eval()
    images, pose, gt = O.synthetic_batch(5, 2, seed=3, size=64)
    rot = O.pairwise_rotations(pose)
    loss = O.iteration_loss(model.forward_views(images, rot), [gt[:, 0], gt[:, 1]])
    loss.backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None])
    rel = ((r0["flat"] - ref).norm() / ref.norm()).item()
    assert rel <= 1e-4, rel
