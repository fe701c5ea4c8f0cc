#!/usr/bin/env python
"""Wire lca_b200 into a Megatron-DeepSpeed checkout (role of the reference's ``patches/Megatron-DeepSpeed.patch``).

    python patches/apply_megatron_deepspeed.py /path/to/Megatron-DeepSpeed [--dry-run] [--revert]

A line-numbered diff against a moving third-party tree stops applying after a few upstream commits, so the edits are
expressed as ANCHORED insertions / replacements (each anchor must match exactly once, every edit is idempotent and
carries a ``# lca_b200`` marker so it can be reverted).  What changes, file by file:

* ``megatron/arguments.py``          ``--ds-ring-sequence-parallel-size`` (R of the U x R mesh; U = ds-sequence-parallel-size / R)
                                     and ``--ds-ring-impl-type`` (basic | zigzag | stripe).
* ``megatron/core/parallel_state.py`` ``initialize_model_parallel(..., ring_parallel_size=1)`` builds the lca_b200 mesh
                                     right after Megatron's own sequence-parallel groups exist, and
                                     ``get_ulysses_sequence_parallel_world_size()`` is added.
* ``megatron/initialize.py``          passes ``args.ds_ring_sequence_parallel_size`` through.
* ``megatron/model/transformer.py``   ``ParallelAttention`` uses ``lca_b200.integrations.megatron.DistributedAttention``
                                     (fused NVLink USP kernels) instead of DeepSpeed's all-to-all wrapper; the head
                                     divisibility assert is relaxed to the ULYSSES degree.

``patches/Megatron-DeepSpeed.patch`` is this script's output on the excerpt tree under ``tests/fixtures/megatron_ds``
(the unit test applies both).  With ``--ds-ring-impl-type zigzag|stripe`` the batch must be sharded with
``lca_b200.integrations.megatron.shard_batch`` in ``get_batch`` (token order is part of the load balancing).
"""
from __future__ import annotations

import argparse
import os
import re
import sys

MARK = "# lca_b200"


class Edit:
    def __init__(self, path, anchor, new, mode):
        self.path, self.anchor, self.new, self.mode = path, anchor, new, mode   # mode: after | before | replace


EDITS = [
    Edit("megatron/arguments.py",
         r"    group\.add_argument\('--force-ds-sequence-parallel', action='store_true',\n",
         "    group.add_argument('--ds-ring-sequence-parallel-size', type=int, default=1,  " + MARK + "\n"
         "                       help='Ring degree R of the Ulysses x Ring sequence-parallel mesh (U = ds-sequence-parallel-size / R).')\n"
         "    group.add_argument('--ds-ring-impl-type', type=str, default='basic', choices=['basic', 'zigzag', 'stripe'],  " + MARK + "\n"
         "                       help='Token layout of the ring dimension (zigzag/stripe need lca_b200 shard_batch in get_batch).')\n",
         "before"),
    Edit("megatron/core/parallel_state.py",
         r"from \.utils import GlobalMemoryBuffer\n",
         "\ntry:  " + MARK + "\n"
         "    from lca_b200.integrations.megatron import initialize_sequence_parallel as _lca_init_sp  " + MARK + "\n"
         "    from lca_b200.globals import PROCESS_GROUP as LCA_PROCESS_GROUP  " + MARK + "\n"
         "except ImportError:  " + MARK + "\n"
         "    _lca_init_sp = None  " + MARK + "\n",
         "after"),
    Edit("megatron/core/parallel_state.py",
         r"    use_distributed_optimizer: bool = False,\n\) -> None:\n",
         "    use_distributed_optimizer: bool = False,\n    ring_parallel_size: int = 1,  " + MARK + "\n) -> None:\n",
         "replace"),
    Edit("megatron/core/parallel_state.py",
         r"    # Build the sequence data parallel groups\.\n",
         "    if _lca_init_sp is not None and sequence_parallel_size > 1:  " + MARK + "\n"
         "        # U x R mesh of the fused NVLink attention: U * R = sequence_parallel_size  " + MARK + "\n"
         "        _lca_init_sp(sequence_parallel_size, ring_parallel_size)  " + MARK + "\n",
         "before"),
    Edit("megatron/core/parallel_state.py",
         r"def get_sequence_parallel_world_size\(\):\n",
         "def get_ulysses_sequence_parallel_world_size():  " + MARK + "\n"
         "    \"\"\"World size of the Ulysses (head-scatter) dimension of the lca_b200 mesh.\"\"\"  " + MARK + "\n"
         "    return LCA_PROCESS_GROUP.ulysses_degree  " + MARK + "\n"
         "\n\n",
         "before"),
    Edit("megatron/initialize.py",
         r"use_distributed_optimizer=args\.use_distributed_optimizer\)\n",
         "use_distributed_optimizer=args.use_distributed_optimizer,\n"
         "                                           ring_parallel_size=args.ds_ring_sequence_parallel_size)  " + MARK + "\n",
         "replace"),
    Edit("megatron/model/transformer.py",
         r"flash_attn_builder = None\n",
         "\ntry:  " + MARK + "\n"
         "    from lca_b200.integrations.megatron import DistributedAttention as LcaDistributedAttention  " + MARK + "\n"
         "except ImportError:  " + MARK + "\n"
         "    LcaDistributedAttention = None  " + MARK + "\n",
         "after"),
    Edit("megatron/model/transformer.py",
         r"            assert args\.num_attention_heads % parallel_state\.get_sequence_parallel_world_size\(\) == 0\n"
         r"            self\.dist_attn = DistributedAttention\(local_attn, parallel_state\.get_sequence_parallel_group\(\)\)\n",
         "            if LcaDistributedAttention is not None:  " + MARK + "\n"
         "                assert args.num_attention_heads % parallel_state.get_ulysses_sequence_parallel_world_size() == 0  " + MARK + "\n"
         "                self.dist_attn = LcaDistributedAttention(local_attn, ring_impl_type=args.ds_ring_impl_type,  " + MARK + "\n"
         "                                                         seq_first=False)  " + MARK + "\n"
         "            else:  " + MARK + "\n"
         "                assert args.num_attention_heads % parallel_state.get_sequence_parallel_world_size() == 0\n"
         "                self.dist_attn = DistributedAttention(local_attn, parallel_state.get_sequence_parallel_group())\n",
         "replace"),
]


def apply_edit(text: str, e: Edit):
    key = next(ln for ln in e.new.splitlines() if MARK in ln)      # first marked line identifies the edit
    if key in text:
        return text, "already applied"
    hits = list(re.finditer(e.anchor, text))
    if len(hits) != 1:
        raise SystemExit(f"{e.path}: anchor matched {len(hits)} times (expected 1): {e.anchor[:70]!r}")
    m = hits[0]
    if e.mode == "after":
        return text[:m.end()] + e.new + text[m.end():], "applied"
    if e.mode == "before":
        return text[:m.start()] + e.new + text[m.start():], "applied"
    return text[:m.start()] + e.new + text[m.end():], "applied"


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("tree")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args(argv)
    texts = {}
    for e in EDITS:
        path = os.path.join(a.tree, e.path)
        if e.path not in texts:
            if not os.path.exists(path):
                raise SystemExit(f"{path} not found: is {a.tree} a Megatron-DeepSpeed checkout?")
            texts[e.path] = open(path).read()
        texts[e.path], what = apply_edit(texts[e.path], e)
        print(f"{e.path}: {what}")
    for rel, text in texts.items():
        compile(text, rel, "exec")                      # the edited files must still be valid Python
        if not a.dry_run:
            with open(os.path.join(a.tree, rel), "w") as f:
                f.write(text)
    return 0


if __name__ == "__main__":
    sys.exit(main())
