"""The Fortran ISO_C_BINDING module (nekstab_b200/fortran/nekstab_b200_c.f90) against the C header (include/nekstab_b200.h):
every entry point bound, same arity, by-value exactly where C passes by value, matching kinds, scalars-by-reference exactly where the
header says NSB_SCALAR -- and the exported symbols of the built library.  No Fortran compiler exists in the build image,
so this is the check that keeps the binding honest (VERDICT r1 'boundary hardening')."""
import os
import re
import subprocess
import sys

from util import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
F90 = os.path.join(ROOT, "nekstab_b200", "fortran", "nekstab_b200_c.f90")
SHIM = os.path.join(ROOT, "nekstab_b200", "fortran", "nekstab_b200_shim.f")


def parse_f90(path=F90):
    """name -> (result kind, [(argname, base type, kind, by_value, is_array)]) from the interface block (independent parser)."""
    text = open(path).read()
    text = re.sub(r"&\s*\n\s*", " ", text)                       # join free-form continuations
    out = {}
    for m in re.finditer(r"^\s*(integer\((\w+)\)|type\((\w+)\))\s+function\s+(nsb_\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)(.*?)end function",
                         text, flags=re.M | re.S):
        rk = m.group(2) or m.group(3)
        name, bound = m.group(4), m.group(6)
        assert name == bound, (name, bound)
        args = [a.strip() for a in m.group(5).split(",") if a.strip()]
        decls = {}
        for line in m.group(7).splitlines():
            d = re.match(r"\s*(integer|real|character|type)\((?:kind=)?(\w+)\)\s*(,\s*value)?\s*::\s*(.*)", line)
            if not d:
                continue
            for v in d.group(4).split(","):
                v = v.strip()
                arr = v.endswith("(*)")
                decls[v.replace("(*)", "")] = (d.group(1), d.group(2), bool(d.group(3)), arr)
        assert set(decls) == set(args), (name, args, decls)
        out[name] = (rk, [(a,) + decls[a] for a in args])
    return out


def test_every_header_entry_point_is_bound_with_matching_arguments():
    import gen_fortran_bindings as gen
    hdr = gen.parse_header()
    f90 = parse_f90()
    assert len(hdr) >= 70
    assert {h[0] for h in hdr} == set(f90), sorted({h[0] for h in hdr} ^ set(f90))
    kinds = {"int": "c_int", "long long": "c_long_long", "double": "c_double", "char": "c_char", "nsb_stats": "nsb_stats",
             "nsb_step_callback": "c_funptr", "void": "c_ptr"}
    for name, ret, params in hdr:
        rk, fargs = f90[name]
        assert rk == {"int": "c_int", "long long": "c_long_long"}.get(ret, "c_ptr"), name
        assert len(fargs) == len(params), name
        for (ctype, pname, ptr, scalar), (fname, base, kind, by_value, is_array) in zip(params, fargs):
            assert kind == kinds[ctype], (name, pname, kind)
            if ctype in ("nsb_step_callback", "void"):
                assert by_value and not is_array, (name, pname)          # pointers handed over as c_funptr / c_ptr values
            else:
                assert by_value == (not ptr), (name, pname, "by value in Fortran <=> not a pointer in C")
                if ptr:
                    assert is_array == ((not scalar) and ctype != "nsb_stats"), (name, pname, "array <=> not NSB_SCALAR")


def test_generated_module_is_up_to_date_and_library_exports_every_symbol():
    import gen_fortran_bindings as gen
    assert open(F90).read() == gen.generate(), "run tools/gen_fortran_bindings.py"
    from nekstab_b200 import lib
    names = {h[0] for h in gen.parse_header()}
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    assert names <= exported, sorted(names - exported)
    assert names == set(lib.EXPORTED), sorted(names ^ set(lib.EXPORTED))          # the ctypes stand-in binds the same set


def test_shim_is_fixed_form_and_calls_only_bound_entry_points():
    """The replacement bodies include Nek5000's fixed-form SIZE / TOTAL, so the shim must be fixed-form itself: no
    free-form continuation, statements in columns 7+, continuation marks in column 6; and every nsb_* it calls is bound."""
    f90 = parse_f90()
    src = open(SHIM).read().splitlines()
    called = set()
    for i, line in enumerate(src, 1):
        if not line.strip() or line[0] in "cC*!":
            continue
        code = line.split("!")[0].rstrip()
        if not code.strip():
            continue
        assert not code.rstrip().endswith("&"), f"{SHIM}:{i}: free-form continuation"
        assert code[:5].strip() == "" or code[:5].strip().isdigit(), f"{SHIM}:{i}: text in columns 1-5: {line!r}"
        called |= set(re.findall(r"\b(nsb_(?!b200_)\w+)\s*\(", code))
    called -= {"nsb_error_message", "nsb_glo"}
    assert called and called <= set(f90), sorted(called - set(f90))
    text = "\n".join(src)
    for routine in ("krylov_inner_product", "krylov_norm", "krylov_normalize", "krylov_cmult", "krylov_add2", "krylov_sub2", "krylov_zero",
                    "krylov_copy", "krylov_matmul", "update_hessenberg_matrix", "matvec", "nonlinear_forward_map"):
        assert re.search(r"subroutine\s+" + routine + r"\s*\(", text), routine
    # uparam / param users include TOTAL (implicit none + uparam without TOTAL was the r1 bug)
    for blk in re.split(r"\n\s*end subroutine", text):
        if re.search(r"\bu?param\s*\(", blk.split("!")[0] if False else blk) and "subroutine" in blk:
            assert "include 'TOTAL'" in blk, blk[:200]
