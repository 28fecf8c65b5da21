"""The reference's OWN caller of the rasterizer (gs_renderer.py: GaussianModel, MiniCam, Renderer; cam_utils.py; sh_utils.py),
made importable on a machine that has no /root/reference.  TEST INFRASTRUCTURE ONLY (tests/, never the product).

A Python reference cannot be copied into the repo, and /root/reference does not exist on the GPU box.  What travels is
what `build_ref_pyc()` produces in THIS container: the three files compiled, UNMODIFIED and from where they lie under
/root/reference, to CPython byte code in oracle/_ref/pyc/*.pyc (git-ignored, not gpurun-ignored — the same treatment as
oracle/_ref/libsimple_knn_ref.so, which is the reference's simple_knn.cu compiled by nvcc).  `load()` imports the
reference modules from the sources when they are present, else from that byte code, with

  * `diff_gaussian_rasterization` and `simple_knn._C` resolving to THIS repo's drop-in packages (the point of the test), and
  * the third-party modules the caller imports but the rasterizer path never touches (plyfile, kiui, mesh, mesh_utils)
    replaced by empty stand-ins, exactly as tests/golden/make_golden.py does.
"""
import importlib.machinery
import importlib.util
import os
import py_compile
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
PYC_DIR = os.path.join(_HERE, "_ref", "pyc")
FILES = ("sh_utils", "cam_utils", "gs_renderer")


def build_ref_pyc():
    """Compile the reference's caller to byte code (only where /root/reference exists). Returns the directory or None."""
    if not os.path.exists(os.path.join(REF, "gs_renderer.py")):
        return PYC_DIR if available() else None
    os.makedirs(PYC_DIR, exist_ok=True)
    for name in FILES:
        py_compile.compile(os.path.join(REF, name + ".py"), cfile=os.path.join(PYC_DIR, name + ".pyc"), dfile=name + ".py",
                           doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    return PYC_DIR


def available():
    return os.path.exists(os.path.join(REF, "gs_renderer.py")) or all(
        os.path.exists(os.path.join(PYC_DIR, n + ".pyc")) for n in FILES)


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load():
    """Returns (cam_utils, gs_renderer, sh_utils) of the reference, bound to this repo's rasterizer and simple_knn."""
    root = os.path.dirname(_HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    import diff_gaussian_rasterization  # noqa: F401  (this repo's drop-in: must win over any stub)
    import simple_knn._C  # noqa: F401
    _stub("plyfile", PlyData=object, PlyElement=object)
    _stub("kiui")
    _stub("mesh", Mesh=object)
    _stub("mesh_utils", decimate_mesh=None, clean_mesh=None)
    mods = {}
    from_source = os.path.exists(os.path.join(REF, "gs_renderer.py"))
    for name in FILES:
        if name in sys.modules and getattr(sys.modules[name], "__dgr_reference__", False):
            mods[name] = sys.modules[name]
            continue
        if from_source:
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
        else:
            path = os.path.join(PYC_DIR, name + ".pyc")
            spec = importlib.util.spec_from_loader(name, importlib.machinery.SourcelessFileLoader(name, path))
        mod = importlib.util.module_from_spec(spec)
        mod.__dgr_reference__ = True
        sys.modules[name] = mod           # gs_renderer does `from sh_utils import ...`
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["cam_utils"], mods["gs_renderer"], mods["sh_utils"]
