"""msnets_b200 -- B200-native matching-space (MS) hot path of ccj5351/MS-Nets.

Hand-written sm_100a CUDA kernels behind a C ABI (include/msnets_b200.h), with
Python mirrors of the reference's interfaces for this path:

    libmatchers      <- src.cpp.lib.libmatchers      (matchers.cpp)
    libfeatextract   <- src.cpp.lib.libfeatextract   (featextract.cpp)
    cbmv             <- src.dataloader.cbmv_generator (get_costs, extract_features_*)
    regression       <- softmax + disparityregression (gcnet_3dcnn.py:127-141)
    confidence       WTA / second-min / peak-ratio / left-right check
    volume           concat / difference 4D volume
    sharding         batch- and disparity-slab sharding over torch.distributed

`install_dropin()` registers the two native-module mirrors under the reference's
import names so its unmodified cbmv_generator.py picks them up.
"""
import sys

from . import _lib
from ._lib import MsnetsError, device_count

__all__ = ["libmatchers", "libfeatextract", "cbmv", "regression", "confidence", "volume", "sharding",
           "install_dropin", "set_aml_exact", "aml_exact", "MsnetsError", "device_count"]


def set_aml_exact(on=True):
    """Process-wide AML arithmetic mode (msn_set_aml_exact).  False (default): SFU exponential, within 2e-6 of the
    reference (1.2e-5 on rare degenerate rows).  True: the reference's own fp32 operations with glibc's expf
    replayed bit for bit -- extract_likelihood and feature channels 4-7 / 12-15 BIT-EXACT (featextract.cpp:435-453),
    about 6x slower on the fused path (fp64 pipe).  Returns the previous setting."""
    prev = bool(_lib.lib().msn_get_aml_exact())
    _lib.check(_lib.lib().msn_set_aml_exact(1 if on else 0))
    return prev


def aml_exact():
    return bool(_lib.lib().msn_get_aml_exact())


def __getattr__(name):
    if name in ("libmatchers", "libfeatextract", "cbmv", "regression", "confidence", "volume", "sharding"):
        import importlib
        mod = importlib.import_module(__name__ + "." + name)
        globals()[name] = mod
        return mod
    raise AttributeError(name)


def install_dropin():
    """Makes `import src.cpp.lib.libmatchers as mtc` / `...libfeatextract as fte`
    (cbmv_generator.py:16-17) resolve to the CUDA-backed mirrors.

    The reference's own `src` / `src.cpp` packages are left alone: when they are importable (its
    checkout is on sys.path, as it is when main_msnet.py runs) they are imported for real, so that
    `from src.dataloader import cbmv_generator` keeps working afterwards; only the two native
    modules under `src.cpp.lib` -- the directory the reference's CMake build writes its .so files to
    (CMakeLists.txt:73) -- are replaced.  Packages that do not exist are stubbed."""
    import importlib
    import importlib.util
    import types
    from . import libfeatextract, libmatchers

    def ensure(name):
        if name in sys.modules:
            return sys.modules[name]
        try:
            spec = importlib.util.find_spec(name)
        except (ImportError, ValueError, AttributeError):
            spec = None
        if spec is not None:
            try:
                return importlib.import_module(name)
            except Exception:   # e.g. a src/cpp/lib/__init__.py that needs the Boost-built .so files
                sys.modules.pop(name, None)
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], leaf, m)
        return m

    for pkg in ("src", "src.cpp", "src.cpp.lib"):
        ensure(pkg)
    lib = sys.modules["src.cpp.lib"]
    sys.modules["src.cpp.lib.libmatchers"] = libmatchers
    sys.modules["src.cpp.lib.libfeatextract"] = libfeatextract
    lib.libmatchers = libmatchers
    lib.libfeatextract = libfeatextract
    return libmatchers, libfeatextract
