"""The C-ABI library loads on a CPU-only box and exports every symbol include/autogpc_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "autogpc_b200.h")
LIB = os.path.join(ROOT, "autogpc_b200", "libautogpc_b200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(agpc_[a-z_]+)\s*\(", src)))


def _ensure_built():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()


def test_header_symbols_are_exported():
    _ensure_built()
    names = _declared()
    assert len(names) >= 14
    out = subprocess.run(["nm", "-D", "--defined-only", LIB], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (agpc_[a-z_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)


def test_library_loads_and_reports_version_without_gpu():
    _ensure_built()
    lib = ctypes.CDLL(LIB)
    assert lib.agpc_version() == 101
    ctx = ctypes.c_void_p()
    rc = lib.agpc_create(0, ctypes.byref(ctx))
    if rc == 0:            # a GPU is present (GPU box): clean up
        lib.agpc_destroy(ctx)
    else:                  # CPU box: must fail loudly, never fall back
        assert rc < 0 and not ctx.value


def test_python_structs_match_header_limits():
    from autogpc_b200 import engine
    src = open(HEADER).read()
    lim = {k: int(v) for k, v in re.findall(r"#define (AGPC_MAX_[A-Z_]+) (\d+)", src)}
    assert (engine.MAX_LEAVES, engine.MAX_TERMS, engine.MAX_TERM_LEN, engine.MAX_THETA) == \
        (lim["AGPC_MAX_LEAVES"], lim["AGPC_MAX_TERMS"], lim["AGPC_MAX_TERM_LEN"], lim["AGPC_MAX_THETA"])
    assert ctypes.sizeof(engine.Program) == 4 * (4 + 3 * engine.MAX_LEAVES + engine.MAX_TERMS + engine.MAX_TERMS * engine.MAX_TERM_LEN)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "autogpc_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in re.sub(r'""".*?"""', "", open(os.path.join(pkg, f)).read(), flags=re.S), f


def _header_struct_fields(name):
    """[(c type, field name)] of ``typedef struct <name> {...}`` in the public header, comments stripped."""
    import re
    text = open(HEADER).read()
    body = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            parts = decl.split()
            out.append((" ".join(parts[:-1]), parts[-1]))
    return out


def test_option_structs_mirror_the_header_field_by_field():
    """``EPOptions`` / ``LBFGSOptions`` are passed by pointer: a reordered or retyped field would be read as garbage by the
    library without any error, so the ctypes mirrors are checked against the header's declarations."""
    import ctypes
    from autogpc_b200 import engine
    ctype = {"double": ctypes.c_double, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64}
    for cname, mirror in (("agpc_ep_options", engine.EPOptions), ("agpc_lbfgs_options", engine.LBFGSOptions)):
        fields = _header_struct_fields(cname)
        assert [f for _, f in fields] == [f for f, _ in mirror._fields_], cname
        assert [ctype[t] for t, _ in fields] == [t for _, t in mirror._fields_], cname
