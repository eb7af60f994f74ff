"""Import helper for the upstream reference (TEST INFRASTRUCTURE ONLY).

The reference (sihan-shao/QuantizationAwareTHzDOE) is pure Python and lives read-only at
/root/reference in the build container; it does NOT exist on the GPU box.  This helper is
used only by oracle/make_golden.py and by the `not gpu` tests that validate the oracle
restatement when the reference happens to be present.

DataType/ElectricField.py:8-12 -> utils/Visualization_Helper.py:1-7 import matplotlib,
pylab and imageio at module import, none of which are installed here, so we register
inert stub modules before importing.
"""
import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("THZ_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Props", "ASM_Prop.py"))


class _Anything(types.ModuleType):
    """A module whose every attribute is another inert stub (callable, subscriptable)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        stub = _Anything(self.__name__ + "." + name)
        setattr(self, name, stub)
        return stub

    def __call__(self, *a, **k):
        return self


_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "matplotlib.axes",
    "matplotlib.ticker", "pylab", "imageio", "mpl_toolkits", "mpl_toolkits.axes_grid1",
    "mpl_toolkits.mplot3d", "pytorch_msssim", "h5py", "seaborn",
]


def import_reference():
    """Put the reference on sys.path (with stubs) and return a namespace of its hot-path symbols."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()):
        from DataType.ElectricField import ElectricField
        from Props.ASM_Prop import ASM_prop
        from Props.CZT_Prop import CZT_prop
        from Props.RSC_Prop import RSC_prop, VRS_prop
        import Components.QuantizedDOE as QD
        import Components.quantization as QZ
        from Components.discrete_doe import DiscreteDOE
        import utils.Helper_Functions as HF
        from Components.Thin_Lens import Thin_LensElement
        from Components.Aperture import ApertureElement
        from LightSource.Gaussian_beam import Guassian_beam
    ns.ElectricField = ElectricField
    ns.ASM_prop = ASM_prop
    ns.CZT_prop = CZT_prop
    ns.RSC_prop = RSC_prop
    ns.VRS_prop = VRS_prop
    ns.QD = QD
    ns.QZ = QZ
    ns.DiscreteDOE = DiscreteDOE
    ns.HF = HF
    ns.Thin_LensElement = Thin_LensElement
    ns.ApertureElement = ApertureElement
    ns.Guassian_beam = Guassian_beam
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints from create_kernel (ASM_Prop.py:279-285) and CZT (CZT_Prop.py:167-176,217)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
