"""world_size-2 gloo (CPU) test of the multi-GPU host logic: pair sharding `rank::world`, the
ragged metric gather and the max-over-ranks timing reduction (c2m_b200/dist.py)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, os.path.join(ROOT, 'c2-matching_b200'))
    from c2m_b200.dist import gather_rows, max_over_ranks, shard_indices
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard_indices(n_pairs, rank, world)
    rows = torch.tensor([[i, 30.0 + i, 0.5] for i in mine], dtype=torch.float64).reshape(-1, 3)
    allrows = gather_rows(rows)
    tmax = max_over_ranks(1.0 + rank, 'cpu')
    if rank == 0:
        q.put((allrows.tolist(), tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    n_pairs, world, port = 7, 2, 29531 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, tmax = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(int(r[0]) for r in rows) == list(range(n_pairs))       # every pair exactly once
    assert [int(r[0]) for r in rows] == [0, 2, 4, 6, 1, 3, 5]           # rank order, ragged (4 + 3)
    assert all(abs(r[1] - (30.0 + r[0])) < 1e-12 for r in rows)
    assert tmax == 2.0


def _combine_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, 'c2-matching_b200'))
    from c2m_b200.dist import combine_argmax, ref_row_slab
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # a fake 9-row x 5-col Ref grid, 6 queries: global scores known to every rank, each rank owns a row slab
    g = torch.Generator().manual_seed(3)
    scores = torch.randn(6, 45, generator=g)
    scores[2, 7] = scores[2, 40] = 9.0            # exact tie across slabs -> lowest index (7) must win
    scores[4, 31] = scores[4, 33] = 8.0           # exact tie inside one slab
    r0, r1 = ref_row_slab(9, rank, world)
    part = scores[:, r0 * 5:r1 * 5]
    v, i = part.max(dim=1)
    # local first-max already resolves in-slab ties to the lowest index
    i = i + r0 * 5
    vg, ig = combine_argmax(v, i.to(torch.int64))
    if rank == 0:
        q.put((vg.tolist(), ig.tolist(), scores.max(dim=1).values.tolist(), [ref_row_slab(9, r, world) for r in range(world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_ref_sharded_argmax_combine_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29331 + os.getpid() % 200
    procs = [ctx.Process(target=_combine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    vg, ig, vmax, slabs = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert slabs == [(0, 5), (5, 9)]
    assert vg == vmax
    assert ig[2] == 7 and ig[4] == 31
    g = torch.Generator().manual_seed(3)
    scores = torch.randn(6, 45, generator=g)
    scores[2, 7] = scores[2, 40] = 9.0
    scores[4, 31] = scores[4, 33] = 8.0
    assert ig == scores.argmax(dim=1).tolist() or all(scores[k, ig[k]] == scores[k].max() for k in range(6))


def test_ref_row_slabs_partition():
    sys.path.insert(0, os.path.join(ROOT, 'c2-matching_b200'))
    from c2m_b200.dist import ref_row_slab
    for rh in (1, 7, 158):
        for w in (1, 2, 3, 8, 11):
            slabs = [ref_row_slab(rh, r, w) for r in range(w)]
            assert slabs[0][0] == 0 and slabs[-1][1] == rh
            assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
            assert max(b - a for a, b in slabs) - min(b - a for a, b in slabs) <= 1


def test_shard_indices_cover():
    sys.path.insert(0, os.path.join(ROOT, 'c2-matching_b200'))
    from c2m_b200.dist import shard_indices
    for n in (0, 1, 126):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in shard_indices(n, r, w))
            assert got == list(range(n))


def _loader_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, 'c2-matching_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from mmsr.data import create_dataloader, create_dataset
    opt = {'name': 'synth', 'type': 'SyntheticRefDataset', 'phase': 'test', 'num': 9, 'gt_size': 32, 'ref_size': 24,
           'scale': 4, 'batch_size': 2, 'num_workers': 1}
    loader = create_dataloader(create_dataset(opt), opt, dist=True)        # rank / world from the process group
    names, sizes = [], []
    for batch in loader:
        names += list(batch['lq_path'])
        sizes.append(int(batch['img_in_lq'].shape[0]))
    loader.close()
    got = [None] * world
    dist.all_gather_object(got, (names, sizes))
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_eval_loader_world2():
    """N1 under an initialised process group: every rank's loader (ShardedEvalSampler -> shape buckets -> PairBatcher with a
    worker process) decodes only its `rank::world` share; together the ranks cover the pair list exactly once."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_loader_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (n0, s0), (n1, s1) = got
    assert n0 == [f'synthetic_{i:04d}.png' for i in range(0, 9, 2)] and n1 == [f'synthetic_{i:04d}.png' for i in range(1, 9, 2)]
    assert s0 == [2, 2, 1] and s1 == [2, 2]
