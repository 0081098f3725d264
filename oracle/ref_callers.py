"""TEST INFRASTRUCTURE — loads the reference's own CALLERS of the rasterizer, unchanged, with the module name
``diff_gaussian_rasterization`` bound to an implementation of the caller's choice.

oracle/build_ref.sh copies scripts/hierslam.py and utils/*.py of the reference to oracle/_ref/callers/ (git-ignored,
travels to the GPU box).  scripts/hierslam.py imports matplotlib, the dataset package and the evaluation helpers
(imgviz, torchmetrics, pytorch_msssim, ... — absent from this image) at module level although the
functions on the rasterizer's path (get_loss_semantic: scripts/hierslam.py:715-853, get_loss_semantic_mlp: :856-1107,
add_new_gaussians_semantic_newrender: :1307-1352, setup_camera: utils/recon_helpers.py:4-28, transform_to_frame /
transformed_params2rendervar_semantic: utils/slam_helpers.py:195-219,278-330) need none of them.  Missing third-party
modules are therefore replaced by permissive stubs for the duration of the import; nothing on the tested path touches
a stub.

Only tests/ may import this module.
"""
from __future__ import annotations

import contextlib
import importlib.abc
import importlib.machinery
import importlib.util
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CALLERS = os.path.join(HERE, "_ref", "callers")

# third-party packages the reference imports at module level but never uses on the rasterizer path
_STUB_IF_MISSING = ("matplotlib", "open3d", "imgviz", "torchmetrics", "pytorch_msssim", "kornia", "plyfile", "natsort",
                    "imageio", "faiss", "lpips", "wandb", "cv2", "tqdm", "PIL", "glob2")
# always stubbed: the reference's dataset package (not copied; an unrelated `datasets` distribution may be installed)
_STUB_ALWAYS = ("datasets",)
_OWNED_ROOTS = ("utils", "scripts", "datasets", "diff_gaussian_rasterization")


def available() -> bool:
    return os.path.exists(os.path.join(CALLERS, "scripts", "hierslam.py"))


class _Anything:
    """stands for any class / function / object of a stubbed module"""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self):
        self.roots = set(_STUB_ALWAYS)
        for r in _STUB_IF_MISSING:
            try:
                if importlib.util.find_spec(r) is None:
                    self.roots.add(r)
            except (ImportError, ValueError):
                self.roots.add(r)
        self.made = []

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        self.made.append(spec.name)
        return m

    def exec_module(self, module):
        pass


def load_callers(dgr_module, tag: str) -> types.SimpleNamespace:
    """Imports the reference's scripts/hierslam.py (and through it utils.slam_helpers / slam_external / recon_helpers)
    with ``import diff_gaussian_rasterization`` resolving to `dgr_module`.  Each call gives a fresh, independent set of
    module objects, so the same caller code can be bound to two implementations in one process."""
    if not available():
        raise FileNotFoundError(f"{CALLERS} missing: run oracle/build_ref.sh where /root/reference exists")
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _OWNED_ROOTS}
    for k in saved:
        del sys.modules[k]
    sys.modules["diff_gaussian_rasterization"] = dgr_module
    finder = _StubFinder()
    sys.meta_path.insert(0, finder)
    path_before = list(sys.path)
    sys.path.insert(0, CALLERS)
    try:
        with contextlib.redirect_stdout(io.StringIO()):      # hierslam.py prints sys.path at import
            spec = importlib.util.spec_from_file_location(f"hierslam_{tag}", os.path.join(CALLERS, "scripts", "hierslam.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        ns = types.SimpleNamespace(hierslam=mod, slam_helpers=sys.modules["utils.slam_helpers"],
                                   slam_external=sys.modules["utils.slam_external"],
                                   recon_helpers=sys.modules["utils.recon_helpers"],
                                   keyframe_selection=sys.modules["utils.keyframe_selection"],
                                   stubbed=sorted(set(n.split(".")[0] for n in finder.made)))
    finally:
        sys.meta_path.remove(finder)
        sys.path[:] = path_before
        for k in [k for k in sys.modules if k.split(".")[0] in _OWNED_ROOTS or k.split(".")[0] in finder.roots and
                  isinstance(sys.modules[k], _StubModule)]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ns


def load_reference_init_over(C_module, S: int, tag: str):
    """Executes the reference's OWN diff_gaussian_rasterization/__init__.py (the copy next to its built extension,
    oracle/_ref/S<S>/) with `from . import _C` (__init__.py:15) resolving to `C_module` instead of the reference's
    pybind extension: the A/B swap INTEGRATION.md section 2 describes."""
    pkg_dir = os.path.join(HERE, "_ref", f"S{S}", "diff_gaussian_rasterization")
    init = os.path.join(pkg_dir, "__init__.py")
    if not os.path.exists(init):
        raise FileNotFoundError(init)
    name = f"hsref_init_{tag}"
    spec = importlib.util.spec_from_file_location(name, init, submodule_search_locations=[])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    sys.modules[name + "._C"] = C_module
    try:
        spec.loader.exec_module(mod)
    except Exception:
        sys.modules.pop(name, None)
        sys.modules.pop(name + "._C", None)
        raise
    assert mod._C is C_module
    return mod
