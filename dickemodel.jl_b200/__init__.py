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
from . import EnergyShellProjections, DickeBCE, UPOs  # noqa: F401
from ._lib import Context, DtwaError, default_context, set_default_context, set_distributed, load as load_library  # noqa: F401

__all__ = ["ClassicalSystems", "ClassicalDicke", "ClassicalLMG", "PhaseSpaces", "TWA", "EnergyShellProjections", "DickeBCE", "UPOs", "symbolic", "Context",
           "DtwaError", "default_context", "set_default_context", "set_distributed", "load_library"]


def _add_unnormalised_aliases():
    """Python NFKC-normalises identifiers in source code (ϵ U+03F5 -> ε U+03B5, ϕ U+03D5 -> φ U+03C6), so
    `ClassicalDicke.q_of_ϵ` written in a script finds `q_of_ε`.  getattr() with the reference's raw spelling does not
    normalise: register those spellings too."""
    import types
    for mod in (ClassicalDicke, ClassicalLMG, ClassicalSystems, PhaseSpaces, TWA, EnergyShellProjections, UPOs):
        for name, obj in list(vars(mod).items()):
            if name.startswith("_") or isinstance(obj, types.ModuleType):
                continue
            raw = name.replace("\u03b5", "\u03f5").replace("\u03c6", "\u03d5")
            if raw != name and not hasattr(mod, raw):
                setattr(mod, raw, obj)


_add_unnormalised_aliases()

