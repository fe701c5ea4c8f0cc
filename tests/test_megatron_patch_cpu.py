"""The Megatron-DeepSpeed integration artefacts (``patches/``): the anchored applier must edit an excerpt-shaped tree,
be idempotent, leave valid Python behind, and the committed unified diff must be exactly its output."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "fixtures", "megatron_ds")
APPLY = os.path.join(ROOT, "patches", "apply_megatron_deepspeed.py")


def _run(tree):
    return subprocess.run([sys.executable, APPLY, tree], capture_output=True, text=True, check=True).stdout


def test_applier_edits_every_anchor_and_is_idempotent(tmp_path):
    tree = tmp_path / "mds"
    shutil.copytree(FIX, tree)
    first = _run(str(tree))
    assert first.count(": applied") == 8 and "already" not in first
    second = _run(str(tree))
    assert second.count("already applied") == 8
    ps = (tree / "megatron" / "core" / "parallel_state.py").read_text()
    assert "ring_parallel_size: int = 1" in ps and "_lca_init_sp(sequence_parallel_size, ring_parallel_size)" in ps
    assert "def get_ulysses_sequence_parallel_world_size" in ps
    tf = (tree / "megatron" / "model" / "transformer.py").read_text()
    assert "LcaDistributedAttention(local_attn, ring_impl_type=args.ds_ring_impl_type" in tf
    assert "--ds-ring-sequence-parallel-size" in (tree / "megatron" / "arguments.py").read_text()
    assert "ring_parallel_size=args.ds_ring_sequence_parallel_size" in (tree / "megatron" / "initialize.py").read_text()
    for rel in ("arguments.py", "core/parallel_state.py", "initialize.py", "model/transformer.py"):
        compile((tree / "megatron" / rel).read_text(), rel, "exec")


def test_committed_patch_is_the_appliers_output(tmp_path):
    a, b = tmp_path / "a", tmp_path / "b"
    shutil.copytree(FIX, a)
    shutil.copytree(FIX, b)
    _run(str(b))
    # GNU patch applies the committed diff to the pristine excerpt and must arrive at the applier's result
    subprocess.run(["patch", "-p1", "-s", "-d", str(a), "-i", os.path.join(ROOT, "patches", "Megatron-DeepSpeed.patch")],
                   check=True)
    for rel in ("arguments.py", "core/parallel_state.py", "initialize.py", "model/transformer.py"):
        assert (a / "megatron" / rel).read_text() == (b / "megatron" / rel).read_text(), rel


def test_shard_batch_matches_extract_functions():
    import torch
    from lca_b200 import EXTRACT_FUNC_DICT
    from lca_b200.integrations.megatron import shard_batch
    tokens = torch.arange(2 * 64).view(2, 64)
    for impl, key in (("basic", "basic"), ("zigzag", "zigzag"), ("stripe", "strip")):
        for rank in range(4):
            got = shard_batch(tokens, rank=rank, world_size=4, ulysses_degree=2, ring_degree=2, ring_impl_type=impl)
            want = EXTRACT_FUNC_DICT[key](tokens, rank, 4, rd=2, ud=2)
            assert torch.equal(got, want)
