"""Host-side logic and the C-ABI surface, on a box without a GPU (no compute calls here)."""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.load_library()
    header = open(os.path.join(ROOT, "include", "dicketwa.h")).read()
    declared = set(re.findall(r"\b(dtwa_[a-z0-9_]+)\s*\(", header))
    assert declared == set(pkg._lib.EXPORTS), declared ^ set(pkg._lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.dtwa_version() == 200


# ---- struct layouts: include/dicketwa.h is the contract; the ctypes mirror and the Julia shim must both agree with it ----
_C_TYPES = {"int32_t": (4, 4), "uint32_t": (4, 4), "int64_t": (8, 8), "uint64_t": (8, 8), "double": (8, 8), "int": (4, 4)}


def _parse_header_structs():
    """{name: [(field, offset, size)], sizeof} for every `typedef struct {...} name;` of the header, natural alignment"""
    header = open(os.path.join(ROOT, "include", "dicketwa.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    structs = {}
    for body, name in re.findall(r"typedef\s+struct\s*\w*\s*\{(.*?)\}\s*(\w+)\s*;", header, flags=re.S):
        off, fields, maxal = 0, [], 1
        for decl in [d.strip() for d in body.split(";") if d.strip()]:
            m = re.match(r"^(?:const\s+)?(\w+)\s*((?:\*\s*(?:const\s*)?)*)\s*([\w, \[\]]+)$", decl)
            assert m, decl
            base, ptr, names = m.group(1), m.group(2), m.group(3)
            for nm in [x.strip() for x in names.split(",")]:
                arr = re.match(r"(\w+)\[(\d+)\]", nm)
                cnt = int(arr.group(2)) if arr else 1
                nm = arr.group(1) if arr else nm
                if "*" in ptr:
                    sz, al = 8, 8
                elif base in _C_TYPES:
                    sz, al = _C_TYPES[base]
                else:
                    sz, al = structs[base]["size"], structs[base]["align"]
                off = (off + al - 1) // al * al
                fields.append((nm, off, sz * cnt))
                off += sz * cnt
                maxal = max(maxal, al)
        structs[name] = {"fields": fields, "size": (off + maxal - 1) // maxal * maxal, "align": maxal}
    return structs


_JL_TYPES = {"Int32": (4, 4), "UInt32": (4, 4), "Cint": (4, 4), "Int64": (8, 8), "UInt64": (8, 8), "Float64": (8, 8), "Cdouble": (8, 8),
             "Cstring": (8, 8)}


def _parse_julia_structs():
    src = open(os.path.join(ROOT, "dickemodel.jl_b200", "julia", "DickeModelB200.jl")).read()
    structs = {}
    for name, body in re.findall(r"^(?:mutable )?struct (C\w+)\s*;?(.*?)\bend$", src, flags=re.S | re.M):
        if name not in {j for _, _, j in _PAIRS}:
            continue
        off, fields, maxal = 0, [], 1
        for nm, ty in re.findall(r"(\w+)::([\w{},]+)", body):
            t = re.match(r"NTuple\{(\d+),(\w+)\}", ty)
            if t:
                sz, al = _JL_TYPES[t.group(2)]
                sz *= int(t.group(1))
            elif ty.startswith("Ptr{"):
                sz, al = 8, 8
            elif ty in _JL_TYPES:
                sz, al = _JL_TYPES[ty]
            else:
                sz, al = structs[ty]["size"], structs[ty]["align"]
            off = (off + al - 1) // al * al
            fields.append((nm, off, sz))
            off += sz
            maxal = max(maxal, al)
        structs[name] = {"fields": fields, "size": (off + maxal - 1) // maxal * maxal, "align": maxal}
    return structs


_PAIRS = [("dtwa_system", "System", "CSystem"), ("dtwa_solver", "Solver", "CSolver"), ("dtwa_sampler", "Sampler", "CSampler"),
          ("dtwa_range", "Range", "CRange"), ("dtwa_reducer", "Reducer", "CReducer"), ("dtwa_reducer_out", "ReducerOut", "CReducerOut"),
          ("dtwa_result", "Result", "CResult"), ("dtwa_event", "Event", "CEvent")]


def test_struct_layouts_match_the_header(pkg):
    """field names, offsets and sizes of the ctypes mirror, parsed against include/dicketwa.h"""
    lib = pkg._lib
    H = _parse_header_structs()
    assert H["dtwa_solver"]["size"] == 104 and H["dtwa_system"]["size"] == 40 and H["dtwa_result"]["size"] == 88
    for cname, pyname, _ in _PAIRS:
        S = getattr(lib, pyname)
        got = [(n, getattr(S, n).offset, getattr(S, n).size) for n, _ in S._fields_]
        assert got == H[cname]["fields"], (cname, got, H[cname]["fields"])
        assert ctypes.sizeof(S) == H[cname]["size"], cname


def test_julia_shim_struct_layouts_match_the_header():
    """Julia cannot run here: at least keep the `struct` blocks of julia/DickeModelB200.jl in step with the header
    (same field names, offsets, sizes), and the positional constructor calls with the field counts"""
    H, J = _parse_header_structs(), _parse_julia_structs()
    for cname, _, jname in _PAIRS:
        assert jname in J, jname
        assert J[jname]["fields"] == H[cname]["fields"], (cname, J[jname]["fields"], H[cname]["fields"])
        assert J[jname]["size"] == H[cname]["size"], cname
    src = open(os.path.join(ROOT, "dickemodel.jl_b200", "julia", "DickeModelB200.jl")).read()
    call = src[src.index("    CSolver(alg,"):]
    call = call[:call.index("\nend")]
    depth, nargs = 0, 1
    for ch in call[call.index("(") + 1:]:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            if depth == 0:
                break
            depth -= 1
        elif ch == "," and depth == 0:
            nargs += 1
    assert nargs == len(J["CSolver"]["fields"]), nargs


def _parse_header_prototypes():
    """{name: (return type, [parameter type classes])} for every `int|const char * dtwa_*(...)` prototype of the header.
    Type classes: 'ptr', 'i32', 'i64', 'f64'."""
    header = open(os.path.join(ROOT, "include", "dicketwa.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = {}
    for ret, name, params in re.findall(r"^\s*((?:const\s+)?\w+\s*\*?)\s*(dtwa_\w+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S | re.M):
        cls = []
        params = " ".join(params.split())
        if params and params != "void":
            for prm in params.split(","):
                prm = prm.strip()
                if "*" in prm:
                    cls.append("ptr")
                else:
                    base = re.sub(r"\bconst\b", "", prm).split()[0]
                    cls.append({"int32_t": "i32", "uint32_t": "i32", "int": "i32", "int64_t": "i64", "uint64_t": "i64", "double": "f64"}[base])
        protos[name] = ("ptr" if "*" in ret else "i32", cls)
    return protos


def _split_top_level(text):
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_shim_ccall_signatures_match_the_header():
    """Julia cannot run here: every `ccall((:dtwa_x, libdicketwa), Ret, (ArgTypes...), args...)` of the shim must name an exported
    function and pass the return type, the number of arguments and the pointer / 32-bit / 64-bit / double class of each one that
    the prototype in include/dicketwa.h declares."""
    P = _parse_header_prototypes()
    assert len(P) >= 25 and "dtwa_ensemble" in P and P["dtwa_last_error"][0] == "ptr"
    src = open(os.path.join(ROOT, "dickemodel.jl_b200", "julia", "DickeModelB200.jl")).read()
    jl_cls = {"Cint": "i32", "Int32": "i32", "UInt32": "i32", "Int64": "i64", "UInt64": "i64", "Float64": "f64", "Cdouble": "f64", "Cstring": "ptr"}
    seen = set()
    for m in re.finditer(r"ccall\(\(:(dtwa_\w+), libdicketwa\),", src):
        name = m.group(1)
        seen.add(name)
        assert name in P, f"{name} is not declared in include/dicketwa.h"
        # the text of the call up to its matching parenthesis
        i, depth = m.start() + len("ccall"), 0
        for j in range(i, len(src)):
            depth += src[j] in "([{"
            depth -= src[j] in ")]}"
            if depth == 0:
                break
        parts = _split_top_level(src[i + 1:j])
        ret, argtypes, args = parts[1], parts[2], parts[3:]
        assert argtypes.startswith("(") and argtypes.endswith(")"), (name, argtypes)
        types = _split_top_level(argtypes[1:-1])
        got = ["ptr" if t.startswith(("Ptr{", "Ref{")) else jl_cls[t] for t in types]
        assert jl_cls.get(ret, "ptr" if ret.startswith("Ptr") else None) == P[name][0], (name, ret)
        assert got == P[name][1], (name, got, P[name][1])
        assert len(args) == len(types), (name, len(args), len(types))
    # the entry points a DickeModel.jl user reaches through the shim
    assert {"dtwa_create", "dtwa_destroy", "dtwa_integrate", "dtwa_ensemble", "dtwa_integrate_fm", "dtwa_lyapunov", "dtwa_integrate_events",
            "dtwa_shell_quadrature", "dtwa_last_error"} <= seen


def test_ctypes_prototypes_match_the_header(pkg):
    """the mirror's `argtypes` (count and class of every parameter) against the prototypes of include/dicketwa.h"""
    P = _parse_header_prototypes()
    L = pkg._lib.load()

    def cls_of(t):
        if t in (ctypes.c_int, ctypes.c_int32, ctypes.c_uint32):
            return "i32"
        if t in (ctypes.c_int64, ctypes.c_uint64, ctypes.c_longlong, ctypes.c_ulonglong):
            return "i64"
        if t is ctypes.c_double:
            return "f64"
        return "ptr"
    checked = 0
    for name, (ret, params) in P.items():
        fn = getattr(L, name)
        if fn.argtypes is None:
            continue
        got = [cls_of(t) for t in fn.argtypes]
        assert got == params, (name, got, params)
        checked += 1
    assert checked >= 20, checked


def test_no_cpu_fallback(pkg):
    """without a CUDA device the product fails loudly instead of computing on the host"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(pkg.DtwaError, match="no CUDA device"):
        pkg.Context([0])
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    with pytest.raises(pkg.DtwaError):
        TWA.average(system, observable="Q", N=10, ts=[0.0, 1.0], distribution=TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=10))


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under the package may import, link or load it"""
    pk = os.path.join(ROOT, "dickemodel.jl_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle|liboracle|oracle/_build|oracle\.c\b|\borc_[a-z]", re.M)
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".inc", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not bad.search(txt.replace("oracle/oracle.c orc_step_table", "")), f


def test_weyl_symbols_equal_the_oracles_strings(pkg, O):
    """two independent renderings of src/TWA.jl:798-883 evaluate to the same bits"""
    from dickemodel_jl_b200 import TWA
    rng = np.random.default_rng(0)
    u = np.column_stack([rng.uniform(-1.3, 1.3, 500), rng.normal(0, 1, 500), rng.uniform(-1.3, 1.3, 500), rng.normal(0, 1, 500)])
    names = O.varnames(O.SYS_DICKE)
    for j in (30, 100, 300.5):
        for ours, theirs in ((TWA.Weyl.Jz, O.Weyl.Jz), (TWA.Weyl.Jx, O.Weyl.Jx), (TWA.Weyl.Jy, O.Weyl.Jy), (TWA.Weyl.Jz2, O.Weyl.Jz2),
                             (TWA.Weyl.Jx2, O.Weyl.Jx2), (TWA.Weyl.Jy2, O.Weyl.Jy2), (TWA.Weyl.n, O.Weyl.n), (TWA.Weyl.n2, O.Weyl.n2)):
            assert np.array_equal(O.eval_expr(str(ours(j)), names, u), O.eval_expr(theirs(j), names, u))
    assert getattr(TWA.Weyl, "Jz²") is TWA.Weyl.Jz2
    # Jz^2 + Jx^2 + Jy^2 = j(j+1) (Casimir) holds for the Weyl symbols up to rounding
    j = 100
    tot = sum(O.eval_expr(str(f(j)), names, u) for f in (TWA.Weyl.Jx2, TWA.Weyl.Jy2, TWA.Weyl.Jz2))
    assert np.allclose(tot, j * (j + 1), rtol=1e-12)


def test_observable_forms(pkg, O):
    """src/TWA.jl:277-299: n-argument function, vector function, symbol, expression"""
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD, symbolic
    system = CD.ClassicalDickeSystem(w0=1.0, w=0.8, g=0.3)
    u = np.random.default_rng(1).uniform(-1, 1, (50, 4))
    names = system.varnames()
    forms = [":(q+p^2-Q)", lambda Q, q, P, p: q + p ** 2 - Q, lambda x: x[1] + x[3] ** 2 - x[0]]
    vals = [O.eval_expr(TWA._expr_of(system, f).lstrip(":"), names, u) for f in forms]
    assert np.array_equal(vals[0], vals[1]) and np.array_equal(vals[1], vals[2])
    H = CD.hamiltonian(system)
    assert np.allclose(O.eval_expr(TWA._expr_of(system, H), names, u), O.hamiltonian(O.SYS_DICKE, system.parameters(), u), rtol=1e-14)
    assert math.isclose(H(u[0]), O.hamiltonian(O.SYS_DICKE, system.parameters(), u[0])[0], rel_tol=1e-14)
    assert "sqrt" in TWA._expr_of(system, lambda Q, q, P, p: symbolic.sqrt(4 - Q ** 2 - P ** 2))
    with pytest.raises(ValueError, match="n-vector or n arguments"):
        TWA._expr_of(system, lambda a, b: a + b)


def test_julia_range_semantics(pkg):
    from dickemodel_jl_b200 import TWA
    r = TWA.jrange(-1.1, 0.01, 1.1)
    assert len(r) == 221 and r.linrange() == (-1.1, 1.1, 221)
    ts = np.asarray(TWA.jrange(0, 0.05, 100))
    assert len(ts) == 2001 and ts[-1] == 100.0 and ts[3] == 0.15 and ts[7] == 0.35   # correctly rounded k/20, not k*0.05
    assert len(TWA.jrange(0, 0)) == 1 and len(TWA.jrange(0, 0.3, 1)) == 4 and TWA.jrange(0, 0.3, 1).stop == 0.9
    assert TWA._as_linrange(np.linspace(-2.4, 2.4, 241), "ys") == (-2.4, 2.4, 241)
    with pytest.raises(ValueError):
        TWA._as_linrange([0.0, 0.1, 0.5], "xs")


def test_dicke_points_and_energies(pkg, O):
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalLMG as CL, PhaseSpaces as PS
    system = CD.ClassicalDickeSystem(**{"ω₀": 1.0, "ω": 1.0, "γ": 1.0})
    assert system.parameters() == [1.0, 1.0, 1.0] and system.varnames() == ["Q", "q", "P", "p"]
    x = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5)      # docs/src/TWAExamples.md:48
    assert math.isclose(x[1], 0.31783724519578224, rel_tol=1e-13)
    assert math.isclose(O.hamiltonian(O.SYS_DICKE, [1, 1, 1], x)[0], 0.5, abs_tol=1e-14)
    xm = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5, sgn="-")
    assert math.isclose(O.hamiltonian(O.SYS_DICKE, [1, 1, 1], xm)[0], 0.5, abs_tol=1e-14) and xm[1] < x[1]
    with pytest.raises(ValueError, match="There is no q"):
        CD.Point(system, Q=1, P=1, p=0, ϵ=-5)
    gs = CD.minimum_energy_point(system)
    assert math.isclose(O.hamiltonian(O.SYS_DICKE, [1, 1, 1], gs)[0], CD.minimum_energy(system), rel_tol=1e-13)
    assert CD.minimum_energy(CD.ClassicalDickeSystem(w0=1, w=1, g=0.3)) == -1.0
    lmg = CL.ClassicalLMGSystem(Ω=1, ξ=-1)
    y = CL.Point(lmg, Q=0.5, ϵ=-0.4)
    assert math.isclose(O.hamiltonian(O.SYS_LMG, [1, -1], y)[0], -0.4, abs_tol=1e-14)
    th, ph = PS.θ_of_QP(0.3, -0.4), PS.ϕ_of_QP(0.3, -0.4)
    assert math.isclose(PS.Q_of_θϕ(th, ph), 0.3, rel_tol=1e-12) and math.isclose(PS.P_of_θϕ(th, ph), -0.4, rel_tol=1e-12)
    assert system.out_of_bounds([1.5, 0, 1.5, 0]) and not system.out_of_bounds([1, 0, 1, 0])


def test_solver_keyword_mapping(pkg):
    from dickemodel_jl_b200 import ClassicalSystems as CS
    lib = pkg._lib
    s = CS.solver_from_kargs({})
    assert s.alg == lib.ALG_DOP853 and s.tol == 1e-12 and s.backwards == 0     # reference defaults (src/ClassicalSystems.jl:82-86)
    s = CS.solver_from_kargs({"integrator_alg": CS.DP5(), "tol": 1e-8, "integate_backwards": True, "maxiters": 500})
    assert (s.alg, s.tol, s.backwards, s.maxiters) == (lib.ALG_DP5, 1e-8, 1, 500)
    assert CS.solver_from_kargs({"integrator_alg": "RK4", "dt": 0.01}).alg == lib.ALG_RK4
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")     # Tsit5 is the real tableau: no substitution warning
        s = CS.solver_from_kargs({"integrator_alg": CS.Tsit5(), "tol": 1e-9})
    assert s.alg == lib.ALG_TSIT5 and CS.last_solver["algorithm"] == "Tsit5" and not CS.last_solver["substituted"]
    assert CS.last_solver["gains"]["beta1"] == pytest.approx(0.14) and CS.last_solver["gains"]["beta2"] == pytest.approx(0.08)   # generic order-5 gains
    assert CS.solver_from_kargs({"integrator_alg": CS.DP8(), "adaptive": False, "dt": 0.1}).alg == lib.ALG_RK8
    with pytest.raises(ValueError, match="dt > 0"):
        CS.solver_from_kargs({"integrator_alg": "RK4"})
    with pytest.raises(TypeError, match="unsupported solver keyword"):
        CS.solver_from_kargs({"abstol": 1e-3})
    with pytest.raises(NotImplementedError):
        CS.solver_from_kargs({"get_fundamental_matrix": True})


def test_monte_carlo_system_protocol(pkg):
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=10, seed=1)
    W2 = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=10, seed=2)
    ts = TWA.jrange(0, 1, 3)
    a = TWA.mcs_for_averaging(system, observable="Q", ts=ts, N=10, distribution=W)
    b = TWA.mcs_for_distributions(system, x="t", y="q", ys=TWA.jrange(-1, 0.5, 1), ts=ts, N=10, distribution=W)
    c = TWA.mcs_chain(a, b)
    assert len(c.f_loop) == 2 and c.N == 10
    with pytest.raises(ValueError, match="same N"):
        TWA.mcs_chain(a, TWA.mcs_for_averaging(system, observable="Q", ts=ts, N=11, distribution=W))
    with pytest.raises(ValueError, match="same ts"):
        TWA.mcs_chain(a, TWA.mcs_for_averaging(system, observable="Q", ts=TWA.jrange(0, 1, 4), N=10, distribution=W))
    with pytest.raises(ValueError, match="same phase space distribution"):
        TWA.mcs_chain(a, TWA.mcs_for_averaging(system, observable="Q", ts=ts, N=10, distribution=W2))
    with pytest.raises(ValueError, match="pass :ts"):
        TWA.mcs_for_distributions(system, x="q", xs=TWA.jrange(0, 1), y="p", ys=TWA.jrange(0, 1), N=10, distribution=W)
    d = TWA.mcs_for_distributions(system, x="q", xs=TWA.jrange(-1, 0.5, 1), N=10, distribution=W)   # y=:t, ts=0:0 default (:417-420)
    assert d.f_loop[0]["y_axis"] == pkg._lib.AXIS_TIME and len(d.ts) == 1
    # finalisers: /N, scalar unwrapping, orientation
    assert np.array_equal(a.f_final(np.full((4, 1), 20.0)), np.full(4, 2.0))
    assert b.f_final(np.ones((4, 5), dtype=np.uint64)).shape == (5, 4)
    v = TWA.mcs_for_variance(system, observable=["Q", "q"], ts=ts, N=10, distribution=W, return_average=True)
    assert v.f_loop[0]["exprs"] == ["Q", "q", "(Q)^2", "(q)^2"]
    raw = np.array([[10.0, 20.0, 30.0, 80.0]])
    var, mean = v.f_final(raw)
    assert np.allclose(mean, [[1.0, 2.0]]) and np.allclose(var, [[3.0 - 1.0, 8.0 - 4.0]])
    s = TWA.mcs_for_survival_probability(system, ts=ts, N=10, distribution=W)
    assert np.allclose(s.f_final(np.ones((4, 1))), (2 * np.pi / 10) ** 2 / 10)


def test_runtime_specialisation_compiles_without_a_gpu(pkg):
    """NVRTC needs no device: the generated source for every system/algorithm compiles to an sm_100a cubin"""
    from dickemodel_jl_b200 import TWA
    lib = pkg._lib
    for kind, exprs in ((lib.SYS_DICKE, [str(TWA.Weyl.Jz(100) / 100), "q", "sqrt(4-Q^2-P^2)*p"]), (lib.SYS_LMG, ["Q", "(Q)^2", "atan(P,Q)"])):
        for alg in (lib.ALG_RK4, lib.ALG_DOP853):
            s = lib.System()
            s.kind = kind
            s.params[0], s.params[1], s.params[2] = 1.0, 1.0, 1.0
            sol = lib.Solver()
            sol.alg, sol.dt, sol.tol = alg, 0.01, 1e-8
            src, nbytes = lib.jit_compile(s, sol, exprs)
            assert "dtwa_jit_ensemble" in src and "__dmul_rn" in src and nbytes > 10000
    s = lib.System()
    s.kind = lib.SYS_LMG
    sol = lib.Solver()
    with pytest.raises(pkg.DtwaError, match="unknown variable 'q'"):
        lib.jit_compile(s, sol, ["q"])


def test_specialised_kernels_keep_their_register_budgets(pkg, tmp_path, monkeypatch):
    """the occupancy the kernels are designed for is a property of the compiled code: RK4 <= 64 registers
    (4 CTAs of 256 threads per SM), DP5 <= 128 (16 one-warp CTAs), DOP853 <= 168 (12), no spill code beyond the samplers' stack
    frame -- checked on the cubin NVRTC produces here, without a GPU"""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump is not on PATH")
    from dickemodel_jl_b200 import TWA
    lib = pkg._lib
    out = str(tmp_path / "k.cubin")
    monkeypatch.setenv("DTWA_JIT_DUMP", out)
    s = lib.System()
    s.kind = lib.SYS_DICKE
    s.params[0], s.params[1], s.params[2] = 1.0, 1.0, 1.0
    for alg, max_regs in ((lib.ALG_RK4, 64), (lib.ALG_DP5, 128), (lib.ALG_TSIT5, 128), (lib.ALG_DOP853, 168)):
        sol = lib.Solver()
        sol.alg, sol.dt, sol.tol = alg, 2.0 ** -7, 1e-8
        lib.jit_compile(s, sol, [str(TWA.Weyl.Jz(100) / 100)])
        res = subprocess.run(["cuobjdump", "-res-usage", out], capture_output=True, text=True).stdout
        m = re.search(r"Function dtwa_jit_ensemble:\s*REG:(\d+) STACK:(\d+)", res)
        assert m, res
        assert int(m.group(1)) <= max_regs, (alg, res)
        assert int(m.group(2)) <= 128, (alg, res)   # (the out-of-line samplers pass the drawn point through the stack)


def test_runtime_specialisation_marks_unit_parameters(pkg):
    """the generated source drops the multiplications by w and g only when the (sign-folded) parameter is exactly +-1
    (x * 1.0 == x, so the specialised kernel keeps the bits of the generic one)"""
    lib = pkg._lib

    def defines(w, g, backwards=0):
        s = lib.System()
        s.kind = lib.SYS_DICKE
        s.params[0], s.params[1], s.params[2] = 1.0, w, g
        sol = lib.Solver()
        sol.alg, sol.dt, sol.tol, sol.backwards = lib.ALG_RK4, 0.01, 1e-8, backwards
        src, _ = lib.jit_compile(s, sol, ["Q"])
        return {d for d in ("DTWA_W_IS_ONE", "DTWA_W_IS_MINUS_ONE", "DTWA_G_IS_ONE", "DTWA_G_IS_MINUS_ONE") if f"#define {d} 1" in src}

    assert defines(1.0, 1.0) == {"DTWA_W_IS_ONE", "DTWA_G_IS_ONE"}
    assert defines(1.0, 1.0, backwards=1) == {"DTWA_W_IS_MINUS_ONE", "DTWA_G_IS_MINUS_ONE"}
    assert defines(0.8, 1.0) == {"DTWA_G_IS_ONE"}
    assert defines(1.0, 0.66) == {"DTWA_W_IS_ONE"}
    assert defines(0.8, -1.0) == {"DTWA_G_IS_MINUS_ONE"}
    assert defines(2.0, 0.5) == set()


GLOO_WORKER = r'''
import os, sys, json
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["DTWA_ROOT"])
import __graft_entry__ as graft
pkg = graft.load_package()
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
N, j = 600, 50
ts = np.arange(0, 9) * 0.25
lo, hi = pkg.distributed.shard_bounds(N, rank, world)
sp = O.make_sampler(O.SAMPLER_HWXSU2, [1.0, 0.31783724519578205, 1.0, 0.0], 1 / j, seed=99, offset=lo)
r = O.ensemble(O.SYS_DICKE, [1, 1, 1], sp, hi - lo, ts, values=[(2, 0, 1.0)], hists=[(0, -1.1, 1.1, 221)], alg=O.ALG_RK4, dt=2.0 ** -6)
arrays = [r["sums"], r["counts"][0], r["n_valid"]]
pkg.distributed.reduce_host_arrays(arrays)
if rank == 0:
    full = O.ensemble(O.SYS_DICKE, [1, 1, 1], O.make_sampler(O.SAMPLER_HWXSU2, [1.0, 0.31783724519578205, 1.0, 0.0], 1 / j, seed=99), N, ts,
                      values=[(2, 0, 1.0)], hists=[(0, -1.1, 1.1, 221)], alg=O.ALG_RK4, dt=2.0 ** -6)
    ok = bool(np.array_equal(arrays[1], full["counts"][0]) and np.array_equal(arrays[2], full["n_valid"])
              and np.allclose(arrays[0], full["sums"], rtol=1e-12, atol=1e-12))
    print(json.dumps({"ok": ok, "total": int(arrays[2][0])}))
dist.destroy_process_group()
'''


def test_sharded_ensemble_world_size_2_gloo(tmp_path, pkg):
    """N > 1 path: contiguous index blocks + one all-reduce (sum) reproduce the single-rank result;
    integer counts bit for bit.  Runs the CPU oracle per rank over gloo."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    env = dict(os.environ, DTWA_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    assert '"ok": true' in line and '"total": 600' in line
    sb, cp = pkg.distributed.shard_bounds, pkg.distributed.chunk_plan
    assert [sb(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]        # balanced, contiguous, complete
    assert [sb(3, r, 8) for r in range(8)] == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5   # more ranks than trajectories: empty blocks
    for N, world, mb in ((100, 8, 5), (9, 4, math.inf), (25, 8, 2), (1, 8, math.inf), (10 ** 7, 3, 10 ** 6)):
        chunk, nchunks = cp(N, world, mb)
        per_rank = []
        for r in range(world):       # what monte_carlo_integrate launches on rank r
            lo, hi = sb(N, r, world)
            ns = [max(0, min(chunk, hi - min(hi, lo + c * chunk))) for c in range(nchunks)]
            assert sum(ns) == hi - lo and len(ns) == nchunks      # same number of launches (= collectives) on every rank
            per_rank.append(ns)
        assert sum(map(sum, per_rank)) == N and chunk <= (mb if mb != math.inf else N)


CHUNK_WORKER = r'''
import json, os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["DTWA_ROOT"])
import __graft_entry__ as graft
pkg = graft.load_package()
from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")


class StubContext:
    """stands in for _lib.Context.ensemble: `sums[0][0]` = trajectories of this launch, then the caller's all-reduce"""
    calls = []

    def ensemble(self, csys, sol, smp, n, ts, reducers, tolerate_errors=True, allreduce=None):
        StubContext.calls.append((int(smp.offset), int(n)))
        a = np.zeros((len(ts), 1))
        a[:, 0] = n
        if allreduce is not None:
            allreduce([a])
        info = {"n_valid": np.zeros(len(ts), dtype=np.uint64), "n_failed": 0, "n_unfinished": 0, "steps_accepted": n, "steps_rejected": 0,
                "lane_steps": n, "resample_rounds": 0, "launches": 1 if n else 0, "jit": 0, "kernel_ms": 0.0, "total_ms": 0.0}
        return [a], info


pkg.set_default_context(StubContext())
pkg.set_distributed(rank, world, lambda arrays: pkg.distributed.reduce_host_arrays(arrays))
system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
out = {}
for N, mb in ((7, 2), (100, 5), (2, float("inf")), (9, float("inf")), (25, 3)):
    StubContext.calls.clear()
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=10, seed=5)
    res = TWA.average(system, observable="Q", N=N, ts=[0.0, 1.0], distribution=W, integrator_alg="RK4", dt=0.1, maxNBatch=mb)
    # every rank made the same number of launches (collectives); the all-reduced total covers N exactly once
    out[f"{N},{mb}"] = [len(StubContext.calls), float(res[0] * N), sorted(StubContext.calls)]
with open(os.path.join(os.environ["DTWA_OUT"], f"rank{rank}.json"), "w") as f:   # (stdout of several ranks interleaves)
    json.dump({"rank": rank, "out": out}, f)
dist.barrier()
dist.destroy_process_group()
'''


def test_uneven_shards_and_maxNBatch_issue_the_same_collectives_on_every_rank(tmp_path, pkg):
    """N not divisible by the world size, more ranks than trajectories, and maxNBatch chunks must not leave a rank
    with fewer all-reduces than the others (NCCL would hang).  3 gloo ranks drive TWA.monte_carlo_integrate over a stub context."""
    script = tmp_path / "chunk_worker.py"
    script.write_text(CHUNK_WORKER)
    env = dict(os.environ, DTWA_ROOT=ROOT, OMP_NUM_THREADS="1", DTWA_OUT=str(tmp_path))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    rows = [json.load(open(tmp_path / f"rank{k}.json")) for k in range(3)]
    for key in rows[0]["out"]:
        N = int(key.split(",")[0])
        launches = {row["out"][key][0] for row in rows}
        assert len(launches) == 1, (key, launches)                        # same collective count on every rank
        assert all(abs(row["out"][key][1] - N) < 1e-9 for row in rows)    # all-reduced total = N on every rank
        covered = sorted((o, n) for row in rows for o, n in row["out"][key][2] if n > 0)
        pos = 0
        for o, n in covered:                                              # blocks tile [0, N) without gap or overlap
            assert o == pos, (key, covered)
            pos += n
        assert pos == N


def test_quantum_dicke_system_delegates_to_the_classical_one(pkg):
    """src/DickeBCE.jl:51-72: the carrier every classical function accepts"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, DickeBCE
    cs = CD.ClassicalDickeSystem(w0=1, w=0.8, g=0.9)
    qs = DickeBCE.QuantumDickeSystem(cs, j=30, Nmax=120)
    assert qs.parameters() == cs.parameters() and qs.varnames() == cs.varnames() and qs.j == 30
    assert qs._c_system().kind == cs._c_system().kind and list(qs._c_system().params)[:3] == [1.0, 0.8, 0.9]
    assert CD.minimum_energy(qs) == CD.minimum_energy(cs)
    q2 = DickeBCE.QuantumDickeSystem(j=2.5, w0=1, w=1, g=1)
    assert q2.j == 2.5 and q2.parameters() == [1.0, 1.0, 1.0]
    with pytest.raises(ValueError, match="half-integer"):
        DickeBCE.QuantumDickeSystem(cs, j=0.3)
    with pytest.raises(ValueError, match="Nmax"):
        DickeBCE.QuantumDickeSystem(cs, j=3, Nmax=0)


def test_reference_spellings_resolve(pkg):
    """the reference's names use ϵ (U+03F5) and ϕ (U+03D5); Python source normalises them, getattr does not"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalLMG as CL, PhaseSpaces as PS
    for mod, names in ((CD, ["q_of_ϵ", "minimum_ϵ_for", "maximum_P_for_ϵ", "maximum_Q_for_ϵ", "minimum_nonnegative_Q_for_ϵ",
                              "Pointθϕ", "discriminant_of_q_solution", "energy_shell_volume", "classical_path_random_sampler"]),
                       (CL, ["P_of_ϵ"]), (PS, ["Q_of_θϕ", "P_of_θϕ", "ϕ_of_QP", "arc_between_θϕ"])):
        for n in names:
            assert callable(getattr(mod, n)), n
    assert CD.q_of_ϵ is getattr(CD, "q_of_ϵ")


def test_host_logic_of_the_widened_modules(pkg):
    """grids of matrix_QP∫∫dqdpδϵ (src/EnergyShellProjections.jl:100-118), the affine form of a ContinuousCallback
    condition, argument errors -- all decided on the host, before any kernel is launched"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS, EnergyShellProjections as ESP
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    Qs, Ps = ESP._generate_QPs(False, False, 0.1, system, -0.5)
    assert np.isclose(Qs[0], -Qs[-1]) and np.isclose(Ps[0], -Ps[-1]) and np.allclose(np.diff(Qs), 0.1)
    assert abs(Qs[-1]) >= CD.maximum_Q_for_ϵ(system, -0.5) and abs(Qs[-1]) < CD.maximum_Q_for_ϵ(system, -0.5) + 0.1
    Qh, Ph = ESP._generate_QPs(True, True, 0.1, system, -0.5)
    assert np.isclose(Qh[-1], 0) and np.isclose(Ph[-1], 0) and len(Qh) == (len(Qs) + 1) // 2
    Qp, Pp = ESP._generate_QPs(False, True, 0.1, system, -0.5)
    assert len(Qp) == len(Qs) and np.isclose(Pp[-1], 0)
    with pytest.raises(ValueError, match="integer fraction of 2"):
        ESP._generate_QPs(False, False, 0.3, system, -0.5)
    assert ESP._exprs(system, lambda x: [1, x[1] ** 2])[1] is False and ESP._exprs(system, lambda x: x[0])[1] is True
    assert ESP._root(None) == 0 and ESP._root("+") == 1 and ESP._root("-") == -1

    hits = []
    cb = CS.ContinuousCallback(lambda x, t, _: 2 * x[3] - 0.5 * x[0] + 0.25, hits.append)
    coef, offset, direction = cb.linear_form(4)
    assert np.allclose(coef, [-0.5, 0, 0, 2]) and offset == -0.25 and direction == 0
    assert CS.ContinuousCallback(lambda x, t, _: x[1], hits.append, None).linear_form(4)[2] == 1
    assert CS.ContinuousCallback(lambda x, t, _: x[1], None, hits.append).linear_form(4)[2] == -1
    with pytest.raises(NotImplementedError, match="affine"):
        CS.ContinuousCallback(lambda x, t, _: x[1] * x[2], hits.append).linear_form(4)
    with pytest.raises(NotImplementedError, match="save_positions"):
        CS.ContinuousCallback(lambda x, t, _: x[1], hits.append, save_positions=(True, False))
    with pytest.raises(NotImplementedError, match="ContinuousCallback"):
        CS.integrate(system, [1.0, 0.3, 1.0, 0.0], 1.0, callback=object())
    with pytest.raises(ValueError, match="out of bounds"):
        CS.lyapunov_spectrum(system, [1.9, 0.0, 1.9, 0.0])
    with pytest.raises(ValueError, match="should have length 4"):
        CS.lyapunov_spectrum(system, [1.0, 0.0])
    with pytest.raises(ValueError, match="3x3|4x4"):
        CS.integrate(system, [1.0, 0.3, 1.0, 0.0], 1.0, get_fundamental_matrix=True, fundamental_matrix_initial=np.eye(3))


def test_split_over_devices_interleaves_and_restores_item_order(pkg, monkeypatch):
    """the multi-GPU split of independent batch items (Lyapunov maps, sections, ...) is host logic: item k goes to GPU
    k mod ndev, per-item arrays come back in item order, times are reduced with max; more GPUs than items is fine"""
    L = pkg._lib
    monkeypatch.setattr(L, "device_context", lambda d: d)
    x = np.arange(11.0)
    seen = {}

    def call(ctx, items):
        seen[ctx] = x[items].copy()
        return x[items] * 2, np.stack([x[items], -x[items]], 1), float(ctx)

    a, b, ms = L.split_over_devices([0, 1, 2], 11, call)
    assert np.array_equal(a, 2 * x) and np.array_equal(b[:, 1], -x) and ms == 2.0
    assert np.array_equal(seen[1], [1.0, 4.0, 7.0, 10.0])
    a, ms = L.split_over_devices([0, 1, 2, 3], 2, lambda c, it: (x[:2][it] + 1, 0.5))
    assert np.array_equal(a, [1.0, 2.0])


def test_step_and_eye_protocol_functions(pkg, O):
    """ClassicalSystems.step(system)(du, u, p, t) and eye(n) (src/ClassicalSystems.jl:30): the host formula of the
    equations of motion equals the oracle's right-hand side, and Hamilton's equations of `hamiltonian`"""
    CS, CD, CL = pkg.ClassicalSystems, pkg.ClassicalDicke, pkg.ClassicalLMG
    assert np.array_equal(CS.eye(3), np.eye(3))
    for system, kind, u in ((CD.ClassicalDickeSystem(w0=1.1, w=0.9, g=0.7), O.SYS_DICKE, [0.4, -0.3, 0.8, 0.2]),
                            (CL.ClassicalLMGSystem(Ω=1.0, ξ=-1.0), O.SYS_LMG, [0.5, 0.3])):
        du = np.zeros(len(u))
        CS.step(system)(du, u, list(system.parameters()) + [1.0], 0.0)
        ref, bad = O.rhs(kind, list(system.parameters()), np.array(u))
        assert not bad and np.allclose(du, ref, rtol=1e-14, atol=1e-15)
        back = np.zeros(len(u))
        CS.step(system)(back, u, list(system.parameters()) + [-1.0], 0.0)
        assert np.allclose(back, -du)
        H = (CD if kind == O.SYS_DICKE else CL).hamiltonian(system)
        e, n = 1e-6, len(u) // 2
        g = np.array([(H(np.array(u) + e * d) - H(np.array(u) - e * d)) / (2 * e) for d in np.eye(len(u))])
        ham = np.concatenate([g[n:], -g[:n]]) if kind == O.SYS_DICKE else np.array([g[1], -g[0]])
        assert np.allclose(du, ham, atol=1e-8)


def test_user_defined_phase_space_distribution_host_side(pkg):
    """src/TWA.jl:22-26: PhaseSpaceDistribution(probability_density, sample, ħ) with user closures"""
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD, symbolic as S
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    rng = np.random.default_rng(0)
    D = TWA.PhaseSpaceDistribution(lambda u: S.exp(-(u[0] ** 2 + u[1] ** 2 + u[2] ** 2 + u[3] ** 2) / 0.02) / (math.pi * 0.02) ** 2,
                                   lambda: rng.normal(0, 0.1, 4), 0.02)
    assert D.ħ == 0.02 and D.nvar is None and D.sample().shape == (4,)
    assert math.isclose(D.probability_density([0, 0, 0, 0]), 1 / (math.pi * 0.02) ** 2)
    assert D.draw_host(5, 4).shape == (5, 4)
    with pytest.raises(ValueError, match="samples 4 variables"):
        D.draw_host(2, 2)
    mc = TWA.mcs_for_survival_probability(system, N=10, ts=[0.0, 1.0], distribution=D)
    assert mc.f_loop[0]["kind"] == pkg._lib.RED_AVERAGE and mc.f_loop[0]["exprs"][0].startswith("(exp(")   # the traced density is the observable
    with pytest.raises(ValueError, match="probability density"):
        TWA.mcs_for_survival_probability(system, N=10, ts=[0.0], distribution=TWA.PhaseSpaceDistribution(None, lambda: np.zeros(4), 0.02))
