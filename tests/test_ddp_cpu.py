"""Data-parallel host logic on CPU: world_size-2 gloo processes run the product's FlatParams / GradReducer / FusedAdamW
(kernels emulated by tests/ops_emulator.py): the bucketed, backward-overlapped all-reduce must deliver exactly the mean
of the per-rank gradients, with per-rank BatchNorm statistics (no SyncBN) like the reference's DDP."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
KW = dict(d_static_num=3, d_time_series_num=5, d_target=1, d_embedding=8, masked_transform_timesteps=4, max_len=4,
          n_duett_layers=2, d_feedforward=96)


def _install_emulator():
    sys.path[:0] = [HERE, os.path.dirname(HERE)]
    import ops_emulator
    from multimodal_edema_prediction_b200 import ops
    for name in ops_emulator.EMULATED:
        setattr(ops, name, getattr(ops_emulator, name))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_emulator()
        from oracle import duett_oracle as O
        from multimodal_edema_prediction_b200.ddp import FlatParams, FusedAdamW, GradReducer
        from multimodal_edema_prediction_b200.loss.losses_duett import StudentKDLoss
        from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
        torch.manual_seed(100 + rank)                                # replicas start DIFFERENT (per-rank seed / checkpoint)
        student = StudentModel(DuettFeatureExtractor(pretrain=False, **KW), head_hidden=16, head_dropout=0.0).train()
        with torch.no_grad():
            student.duett.tab_encoder[3].batch_norm.running_mean.add_(float(rank))     # ... including BatchNorm buffers
        cfg = O.DuettConfig(3, 5, 4, d_embedding=8, n_layers=2, d_feedforward=96)
        b = O.synth_batch(cfg, 6, seed=100 + rank)                   # each rank its own shard
        z_t = torch.randn(6, generator=torch.Generator().manual_seed(rank))
        loss_fn = StudentKDLoss()

        def local_step():
            z = student(b["x_ts"], b["x_static"], list(b["bin_ends"]))
            loss_fn(z, z_t, b["y"])["total"].backward()

        # 0) FlatParams broadcasts rank 0's parameters and buffers, like DDP does at construction
        flat = FlatParams(student)
        for t in [flat.data] + [bf.float() for bf in student.buffers()]:
            hi, lo = t.clone(), t.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            assert torch.equal(hi, lo), "replicas differ after FlatParams construction"
        # 1) this rank's own gradients (no reducer attached: nothing is communicated)
        flat.zero_grad()
        local_step()
        local = {n: p.grad.clone() for n, p in student.named_parameters()}
        # 2) overlapped bucketed all-reduce
        red = GradReducer(flat).attach()
        opt = FusedAdamW(flat, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
        opt.zero_grad()
        red.start_step()
        local_step()
        scale = red.finish()
        assert scale == 0.5 and red.launched >= 12, red.launched    # 3 weight-group buckets per encoder (2 layers x 2 axes) + rest
        errs = []
        for n, p in student.named_parameters():
            parts = [torch.zeros_like(local[n]) for _ in range(world)]
            dist.all_gather(parts, local[n])
            want = sum(parts) / world
            errs.append(float((p.grad * scale - want).abs().max()))
        # 3) fused AdamW on the flat buffers == torch.optim.AdamW on the averaged grads with global-norm clipping
        ref_params = [torch.nn.Parameter(p.detach().clone()) for p in flat.params]
        assert any(n.startswith("duett.head.") for n in flat.unused) and any("pretrain_value_proj" in n for n in flat.unused)
        for n, rp, p in zip(flat.names, ref_params, flat.params):
            # parameters the student never uses (the backbone's own supervised / SSL heads) have .grad None under torch
            # autograd: torch.optim.AdamW skips them (no weight decay either) and so must FusedAdamW
            rp.grad = None if n in flat.unused else p.grad.detach().clone() * scale
        torch.nn.utils.clip_grad_norm_([rp for rp in ref_params if rp.grad is not None], 1.0)
        ropt = torch.optim.AdamW(ref_params, lr=1e-3, weight_decay=0.01)
        ropt.step()
        opt.step(grad_scale=scale)
        perr = max(float((p - rp).abs().max()) for p, rp in zip(flat.params, ref_params))
        # replicas stay identical after the step
        chk = flat.data.clone()
        dist.all_reduce(chk, op=dist.ReduceOp.MAX)
        q.put((rank, max(errs), perr, float((chk - flat.data).abs().max())))
    except Exception as ex:      # surface the failure instead of letting the parent wait for the queue timeout
        import traceback
        q.put((rank, "error", traceback.format_exc(), repr(ex)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_and_fused_adamw_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, gerr, perr, derr in res:
        assert gerr < 1e-6, (rank, gerr)
        assert perr < 1e-6, (rank, perr)
        assert derr == 0.0, (rank, derr)


def test_flat_param_order_follows_backward_completion():
    sys.path[:0] = [HERE, os.path.dirname(HERE)]
    from multimodal_edema_prediction_b200.ddp import FlatParams
    from multimodal_edema_prediction_b200.models.main_architecture_duett import DuettFeatureExtractor, StudentModel
    student = StudentModel(DuettFeatureExtractor(pretrain=False, **KW), head_hidden=16)
    flat = FlatParams(student)
    first = lambda s: min(i for i, n in enumerate(flat.names) if s in n)
    assert first("head.") < first("time_transformers.1.") < first("event_transformers.1.") < first("time_transformers.0.") \
        < first("event_transformers.0.") < first("embedding_layers.")
    for p in flat.params:                       # parameters and grads are views into the flat buffers
        assert p.data_ptr() >= flat.data.data_ptr() and p.grad.data_ptr() >= flat.grad.data_ptr()
    assert flat.numel % 8 == 0


def _eval_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_emulator()
        from multimodal_edema_prediction_b200.training_duett import evaluator
        g = torch.Generator().manual_seed(11)
        z_all, y_all = torch.randn(37, generator=g), (torch.rand(37, generator=g) < 0.4).float()
        lo, hi = (0, 23) if rank == 0 else (23, 37)                 # uneven shards
        res = evaluator.binary_metrics(z_all[lo:hi], y_all[lo:hi])
        # a rank whose loader is empty still joins the gather and reports the metric of the other rank's shard
        loader = [] if rank == 1 else [{"logits": z_all, "y": y_all}]
        res2 = evaluator.evaluate_binary(torch.nn.Identity(), loader, torch.device("cpu"), lambda m, b, d: b)
        q.put((rank, res, res2))
    except Exception as ex:
        import traceback
        q.put((rank, "error", traceback.format_exc(), repr(ex)))
    finally:
        dist.destroy_process_group()


def test_evaluator_scores_the_whole_loader_on_every_rank_world2():
    """binary_metrics gathers every rank's (unevenly sized) shard before ranking: both ranks report the AUROC / AUPRC of the
    whole evaluation set (the reference scores rank 0's shard only)."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(11)
    z_all, y_all = torch.randn(37, generator=g), (torch.rand(37, generator=g) < 0.4).float()
    want_roc = roc_auc_score(y_all.numpy(), torch.sigmoid(z_all).numpy())
    want_pr = average_precision_score(y_all.numpy(), torch.sigmoid(z_all).numpy())
    for r in res:
        assert r[1] != "error", r
        rank, m, m2 = r
        assert m["n"] == 37 and abs(m["auroc"] - want_roc) < 1e-12 and abs(m["auprc"] - want_pr) < 1e-12, (rank, m)
        assert m2["n"] == 37 and abs(m2["auroc"] - want_roc) < 1e-12, (rank, m2)
