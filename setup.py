"""`pip install -e .` / `python setup.py build_ext --inplace` build the sm_100a extension in-tree via lca_b200.ops.build."""
from setuptools import setup
from setuptools.command.build_ext import build_ext
from setuptools.command.build_py import build_py


def _build_native():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("lca_build", os.path.join(os.path.dirname(__file__), "lca_b200", "ops", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


class BuildExt(build_ext):
    def run(self):
        _build_native()


class BuildPy(build_py):
    def run(self):
        try:
            _build_native()
        except Exception as e:  # noqa: BLE001  (CPU-only installs still get the PyTorch engine)
            print(f"[lca_b200] native build skipped: {e}")
        super().run()


setup(cmdclass={"build_ext": BuildExt, "build_py": BuildPy})
