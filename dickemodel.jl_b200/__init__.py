"""dickemodel_jl_b200 -- B200-native drop-in for DickeModel.jl's TWA ensemble hot path.

The directory is called `dickemodel.jl_b200`; because of the dot it is imported under the module
name `dickemodel_jl_b200` through `__graft_entry__.load_package()` (tests/conftest.py, bench.py).

    from dickemodel_jl_b200 import ClassicalDicke, ClassicalLMG, ClassicalSystems, TWA, PhaseSpaces

mirror the reference's modules of the same names (src/DickeModel.jl:1-19); everything numerical
runs in libdicketwa.so (csrc/, C ABI in include/dicketwa.h) on an sm_100a GPU.  There is no CPU
fallback: importing works anywhere, computing needs the built library and a B200.
"""
from . import _lib  # noqa: F401
from . import symbolic, PhaseSpaces, ClassicalSystems, ClassicalDicke, ClassicalLMG, TWA, distributed  # noqa: F401
from . import EnergyShellProjections, DickeBCE  # noqa: F401
from ._lib import Context, DtwaError, default_context, set_default_context, set_distributed, load as load_library  # noqa: F401

__all__ = ["ClassicalSystems", "ClassicalDicke", "ClassicalLMG", "PhaseSpaces", "TWA", "EnergyShellProjections", "DickeBCE", "symbolic", "Context",
           "DtwaError", "default_context", "set_default_context", "set_distributed", "load_library"]
