"""End-to-end convergence parity (role of the reference's loss-curve image, README.md:157-162):
a tiny LM trained with Ulysses2 x Ring2 sequence parallelism follows the single-process loss curve."""
import torch
import torch.distributed as dist

from dist_utils import run_distributed


def _train(rank, world, U, R, variant, steps=6):
    import lca_b200
    from lca_b200 import EXTRACT_FUNC_DICT, set_seq_parallel_pg
    from lca_b200.kernels import AttnType
    from lca_b200.models import SPTransformerConfig, SPTransformerLM, allreduce_sp_grads
    S, B = 64, 2
    torch.manual_seed(0)
    cfg = SPTransformerConfig(ring_impl_type=variant, attn_type=AttnType.TORCH)
    set_seq_parallel_pg(U, R, rank, world)
    model = SPTransformerLM(cfg)                    # same seed -> identical replicas
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(1)
    data = torch.randint(0, cfg.vocab_size, (steps, B, S + 1), generator=g)
    losses = []
    for t in range(steps):
        tok, lab = data[0, :, :-1], data[0, :, 1:]          # memorise one batch: the curve must fall
        if world > 1:
            tok = EXTRACT_FUNC_DICT[variant](tok, rank, world, rd=R, ud=U)
            lab = EXTRACT_FUNC_DICT[variant](lab, rank, world, rd=R, ud=U)
        logits = model(tok, S)
        loss_sum = torch.nn.functional.cross_entropy(logits.flatten(0, 1), lab.flatten(), reduction="sum")
        loss = loss_sum / (B * S)                   # global mean: local sums add up across ranks
        opt.zero_grad()
        loss.backward()
        allreduce_sp_grads(model)
        opt.step()
        tot = loss.detach().clone()
        if world > 1:
            dist.all_reduce(tot)
        losses.append(float(tot))
    return losses


def _worker(rank, world, U, R, variant, ref_losses):
    got = _train(rank, world, U, R, variant)
    for a, b in zip(got, ref_losses):
        assert abs(a - b) < 2e-4 * max(1.0, abs(b)), (got, ref_losses)


def test_loss_curve_matches_single_process():
    import lca_b200
    ref = _train(0, 1, 1, 1, "zigzag")
    assert ref[-1] < ref[0]
    run_distributed(_worker, 4, 2, 2, "zigzag", ref)
