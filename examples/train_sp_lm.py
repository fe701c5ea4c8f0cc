"""Train a tiny decoder-only LM with sequence-parallel attention (``lca_b200.models.SPTransformerLM``) on a synthetic
retrieval task (every position must output the FIRST token of its sequence, which only attention can deliver -- for most
ranks that token lives on another rank): every rank holds S/P tokens of the same sequences, attention runs through ``LongContextAttention``, the
gradients of the replicated weights are summed over the sequence-parallel group.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_sp_lm.py --device cpu
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_sp_lm.py --ulysses 2 --seq 8192 --dim 512 --heads 8

This is the integration contract of the reference's Megatron-DeepSpeed patch in 60 lines: call ``set_seq_parallel_pg``
where the model-parallel groups are built, shard tokens / labels / position ids with the layout of the ring variant,
use ``LongContextAttention`` as the attention core.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lca_b200 import EXTRACT_FUNC_DICT, set_seq_parallel_pg  # noqa: E402
from lca_b200.kernels import AttnType  # noqa: E402
from lca_b200.models import SPTransformerConfig, SPTransformerLM, allreduce_sp_grads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--ulysses", type=int, default=1)
    ap.add_argument("--ring-impl", default="zigzag", choices=["basic", "zigzag", "strip"])
    ap.add_argument("--seq", type=int, default=64)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--heads", type=int, default=4)
    ap.add_argument("--kv-heads", type=int, default=2)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--lr", type=float, default=1e-2)
    a = ap.parse_args()

    cpu = a.device == "cpu"
    dist.init_process_group("gloo" if cpu else "nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cpu") if cpu else torch.device("cuda", rank % torch.cuda.device_count())
    if not cpu:
        torch.cuda.set_device(dev)
    U, R = a.ulysses, world // a.ulysses
    set_seq_parallel_pg(U, R, rank, world)

    torch.manual_seed(0)                                     # replicated weights: same init on every rank
    cfg = SPTransformerConfig(vocab_size=64, dim=a.dim, n_layers=a.layers, n_heads=a.heads, n_kv_heads=a.kv_heads,
                              ring_impl_type=a.ring_impl, attn_type=AttnType.TORCH if cpu else AttnType.FA)
    model = SPTransformerLM(cfg).to(dev)
    if not cpu:
        model = model.to(torch.bfloat16)
    opt = torch.optim.AdamW(model.parameters(), lr=a.lr)
    shard = lambda t: EXTRACT_FUNC_DICT[a.ring_impl](t, rank, world, rd=R, ud=U)   # noqa: E731  (dim 1 = sequence)

    g = torch.Generator().manual_seed(1)
    first = last = None
    for step in range(a.steps):
        tokens = torch.randint(0, cfg.vocab_size, (a.batch, a.seq), generator=g).to(dev)
        labels = tokens[:, :1].expand(-1, a.seq).contiguous()          # retrieve the first token of the sequence
        weight = torch.ones(a.batch, a.seq, device=dev)
        logits = model(shard(tokens), a.seq).float()
        loss_tok = F.cross_entropy(logits.flatten(0, 1), shard(labels).flatten(), reduction="none")
        loss = (loss_tok * shard(weight).flatten()).sum() / weight.sum()     # local share of the global mean
        opt.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_sp_grads(model)                            # full-sequence gradient = sum of the per-rank gradients
        opt.step()
        total = loss.detach().clone()
        dist.all_reduce(total)
        first = total.item() if first is None else first
        last = total.item()
        if rank == 0 and (step % 10 == 0 or step == a.steps - 1):
            print(f"step {step:4d}  loss {last:.4f}")
    if rank == 0:
        print(f"loss {first:.3f} -> {last:.3f}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
