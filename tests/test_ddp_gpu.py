"""Data-parallel gradient equivalence on real GPUs (-m gpu, needs >= 2 devices; run with `gpurun --gpus 2`).

SURVEY.md 8e: the path shards by the batch dimension, BatchNorm statistics stay per-GPU, the one exchange is the gradient mean.
Claim checked: after engine.TrainStep's bucketed, backward-overlapped NCCL all-reduce, every rank holds
(g(shard_0) + g(shard_1)) / 2 -- the single-GPU gradient of the concatenated batch with per-shard BatchNorm -- and the ranks' weights
stay identical after optimizer steps, eagerly and under CUDA-graph replay."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import ml_cvnets_b200 as m
    from oracle import cvnets_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        width, B, res = 0.5, 16, 128
        P = O.seeded_fill_(O.mobilevit_v2_shapes(width), 7)

        def fresh(train):
            model = m.MobileViTv2(m.default_opts(width_multiplier=width))
            model.load_state_dict(P, strict=True)
            model = model.cuda()
            return model.train() if train else model.eval()

        shards = [(O.seeded_input((B, 3, res, res), 50 + r).cuda(), ((torch.arange(B, device="cuda") * 37 + 11 * r) % 1000)) for r in range(world)]

        def reference(train):
            """every shard through the same kernels on THIS GPU, gradients averaged (BatchNorm per shard, as in the recipe)"""
            tot = None
            for xs, ys in shards:
                ts = m.TrainStep(fresh(train), lr=0.0, weight_decay=0.0, data_parallel=False)
                ts.step(xs, ys)
                tot = ts.ws.flat_g.clone() if tot is None else tot + ts.ws.flat_g
            return tot / world

        def rel(a, b):
            return float((a - b).norm() / b.norm())

        errs = {}
        x, y = shards[rank]
        # (a) eval-mode BatchNorm (running statistics): no batch-statistics amplification -> the exchange itself is checked tightly
        ref = reference(False)
        ts = m.TrainStep(fresh(False), lr=0.0, weight_decay=0.0, n_buckets=3)
        assert ts.world == world
        errs["eval"] = []
        for it in range(3):  # plan, descriptor tables, bucketed
            ts.step(x, y)
            errs["eval"].append(rel(ts.ws.flat_g / world, ref))  # flat_g holds the SUM over ranks (cvb_grad_norm divides by world)
        n_fire = len(ts.ws._bucket_range)
        # (b) train mode (per-GPU batch statistics, the recipe): atomics-order noise is amplified by BatchNorm through ~60 bf16 layers, so the
        # bound is the run-to-run difference of the single-GPU reference itself
        ref_t, ref_t2 = reference(True), reference(True)
        noise = rel(ref_t2, ref_t)
        model = fresh(True)
        ts = m.TrainStep(model, lr=0.0, weight_decay=0.0, n_buckets=3)
        errs["train"] = []
        for it in range(3):
            ts.step(x, y)
            errs["train"].append(rel(ts.ws.flat_g / world, ref_t))
        ts.capture(x, y)
        ts.step(x, y)
        errs["train_graph"] = rel(ts.ws.flat_g / world, ref_t)
        errs["train_noise"] = noise
        # weights stay in lock-step across ranks with a real learning rate (captured step, 3 replays)
        ts.set_lr(1e-3)
        for _ in range(3):
            ts.step(x, y)
        flat = ts.opt.flat_p.clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        drift = max(float((g - gathered[0]).abs().max()) for g in gathered)
        ts.ws.broadcast_buffers()  # what the next forward starts with (DDP syncs buffers at the START of each forward)
        bufs = torch.cat([b.flatten().float() for b in model.buffers()])
        gb = [torch.empty_like(bufs) for _ in range(world)]
        dist.all_gather(gb, bufs)
        bdrift = max(float((g - gb[0]).abs().max()) for g in gb)
        torch.cuda.synchronize()
        q.put((rank, errs, n_fire, drift, bdrift))
        q.close()
        q.join_thread()
    except BaseException as e:  # the parent must never wait for a result that will not come
        q.put((rank, repr(e)[:500], -1, -1.0, -1.0))
        q.close()
        q.join_thread()
    # a captured CUDA graph that contains NCCL kernels is still alive here; tearing the process group down under it can block, and the
    # result is already delivered: leave without the collective shutdown
    os._exit(0)


def test_allreduced_gradients_equal_single_gpu_gradients_of_the_concatenated_batch():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.terminate()
    for rank, errs, n_fire, drift, bdrift in res:
        assert not isinstance(errs, str), f"rank {rank} failed: {errs}"
        print(f"rank {rank}: rel-L2 of the all-reduced gradient vs the single-GPU gradient of the concatenated batch: {errs}; buckets {n_fire}; "
              f"weight drift across ranks {drift:.3g}; buffer drift {bdrift:.3g}")
        assert all(e <= 5e-3 for e in errs["eval"]), errs
        bound = 3.0 * errs["train_noise"] + 5e-3
        assert all(e <= bound for e in errs["train"]) and errs["train_graph"] <= bound, errs
        assert n_fire == 3, "the bucket plan did not form (gradients were exchanged in one piece)"
        assert drift == 0.0, "ranks diverged: every rank must apply the identical update to identical weights"
        assert bdrift == 0.0, "BatchNorm running statistics must follow rank 0 (DDP broadcast_buffers semantics)"


def _clip_worker(rank, world, port, q):
    import torch.distributed as dist
    import ml_cvnets_b200 as m
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        N, d = 16, 128
        feats = []
        for r in range(world):
            g = torch.Generator(device="cuda").manual_seed(900 + r)
            i = torch.nn.functional.normalize(torch.randn(N, d, device="cuda", generator=g), dim=-1).bfloat16()
            t = torch.nn.functional.normalize(torch.randn(N, d, device="cuda", generator=g), dim=-1).bfloat16()
            feats.append((i, t))
        ls = torch.tensor(2.0, device="cuda")
        # distributed: this rank's features, all-gather inside the loss, reduce-scatter in its backward
        img = feats[rank][0].clone().requires_grad_(True)
        txt = feats[rank][1].clone().requires_grad_(True)
        lsd = ls.clone().requires_grad_(True)
        loss = m.clip_contrastive_loss(img, txt, lsd)
        loss.backward()
        # reference on this GPU: fp32 torch on the concatenated global batch (the mean over all G rows = mean over ranks of the per-rank losses)
        I = torch.cat([f[0] for f in feats]).float().requires_grad_(True)
        T = torch.cat([f[1] for f in feats]).float().requires_grad_(True)
        lsr = ls.clone().requires_grad_(True)
        s = torch.clamp(lsr.exp(), 0, 100.0)
        lab = torch.arange(world * N, device="cuda")
        rows = slice(rank * N, (rank + 1) * N)
        per_rank = [0.5 * (torch.nn.functional.cross_entropy((s * I @ T.t())[r * N:(r + 1) * N], lab[r * N:(r + 1) * N]) +
                           torch.nn.functional.cross_entropy((s * T @ I.t())[r * N:(r + 1) * N], lab[r * N:(r + 1) * N])) for r in range(world)]
        sum(per_rank).backward()  # gather_all_features' backward SUMS the contributions of every rank's loss (ddp_functional_utils.py:334-357)
        e_loss = abs(float(loss) - float(per_rank[rank])) / abs(float(per_rank[rank]))
        e_i = float((img.grad.float() - I.grad[rows]).norm() / I.grad[rows].norm())
        e_t = float((txt.grad.float() - T.grad[rows]).norm() / T.grad[rows].norm())
        q.put((rank, (e_loss, e_i, e_t), 0, 0.0, 0.0))
        q.close()
        q.join_thread()
    except BaseException as e:
        q.put((rank, repr(e)[:500], -1, -1.0, -1.0))
        q.close()
        q.join_thread()
    os._exit(0)


def test_clip_loss_feature_gather_matches_the_global_batch():
    """CLIP's one exchange step (all-gather of the features in the loss, reduce-scatter in its backward) on real GPUs."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_clip_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        if p.is_alive():
            p.terminate()
    for rank, errs, *_ in res:
        assert not isinstance(errs, str), f"rank {rank} failed: {errs}"
        print(f"rank {rank}: contrastive loss / image-feature grad / text-feature grad relative errors vs the global-batch fp32 reference: {errs}")
        assert errs[0] <= 5e-3 and errs[1] <= 2e-2 and errs[2] <= 2e-2, errs
