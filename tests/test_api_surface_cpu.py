"""Drop-in check: every public module-level name of the reference's packages exists under the same name here
(``import lca_b200 as yunchang`` must keep working for user code).  The reference tree is parsed, never imported."""
import ast
import importlib
import os

import pytest

REF = os.environ.get("LCA_B200_REFERENCE_DIR", "/root/reference")

# names that are implementation details of the reference's optional third-party backends (symbols of flash_attn /
# flashinfer / aiter / sageattention that it merely imports at module level) or plain helper imports
THIRD_PARTY = {
    "yunchang.globals": {"flash3_attn_func", "flash_attn_forward_hopper", "flash_attn_func_hopper_backward",
                         "flash_attn_func_aiter", "single_prefill_with_kv_cache", "cuda_arch"},
    "yunchang.kernels": {"SparseAttentionMeansim", "flash3_attn_func", "partial", "auto"},
}


def _module_level_names(path):
    names = set()
    def visit(body):
        for n in body:
            if isinstance(n, ast.ImportFrom):
                names.update(a.asname or a.name for a in n.names)
            elif isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                names.add(n.name)
            elif isinstance(n, ast.Assign):
                names.update(t.id for t in n.targets if isinstance(t, ast.Name))
            elif isinstance(n, (ast.Try, ast.If)):          # guarded optional imports
                visit(n.body)
                for h in getattr(n, "handlers", []):
                    visit(h.body)
                visit(n.orelse)
    visit(ast.parse(open(path).read()).body)
    return {n for n in names if not n.startswith("_") and n != "*"}


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "yunchang")), reason="reference tree not available")
@pytest.mark.parametrize("mod", ["yunchang", "yunchang.ring", "yunchang.hybrid", "yunchang.ulysses", "yunchang.comm",
                                 "yunchang.kernels", "yunchang.globals"])
def test_every_reference_name_resolves(mod):
    rel = mod.replace(".", "/")
    path = os.path.join(REF, rel, "__init__.py") if os.path.isdir(os.path.join(REF, rel)) else os.path.join(REF, rel + ".py")
    ours = importlib.import_module(mod.replace("yunchang", "lca_b200", 1))
    wanted = _module_level_names(path) - THIRD_PARTY.get(mod, set())
    wanted -= {"torch", "dist", "os", "Enum", "Optional", "Tuple", "Any", "Tensor", "Function", "Module"}
    missing = sorted(n for n in wanted if not hasattr(ours, n))
    assert not missing, f"{mod}: {missing}"


def test_singleton_aliases():
    import lca_b200.globals as g
    assert g.ProcessGroupSingleton() is g.PROCESS_GROUP
    assert g.Singleton() is g.Singleton()
    assert g.get_cuda_arch().count(".") == 1


def _ref_modules():
    root = os.path.join(REF, "yunchang")
    if not os.path.isdir(root):
        return []
    mods = []
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py") and f != "__init__.py":
                rel = os.path.relpath(os.path.join(d, f), REF)[:-3].replace(os.sep, ".")
                mods.append(rel)
    return sorted(mods)


@pytest.mark.parametrize("mod", _ref_modules())
def test_every_reference_module_path_imports_with_its_public_defs(mod):
    """``from yunchang.ring.zigzag_ring_flash_attn import zigzag_ring_flash_attn_forward`` style imports of third-party
    code: same module path, same top-level function / class names."""
    tree = ast.parse(open(os.path.join(REF, mod.replace(".", os.sep) + ".py")).read())
    defs = {n.name for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and not n.name.startswith("_")}
    ours = importlib.import_module(mod.replace("yunchang", "lca_b200", 1))
    missing = sorted(n for n in defs if not hasattr(ours, n))
    assert not missing, f"{mod}: {missing}"
