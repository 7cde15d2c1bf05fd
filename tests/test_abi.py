"""The C-ABI shared library loads without a GPU and exports every symbol include/bayesssm_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "bayesssm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bssm_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bayesssm_b200 import _native
    from bayesssm_b200.build import build_native
    build_native()
    lib = _native.load_library()
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _native.SYMBOLS, f"{n} has no ctypes prototype"
    assert lib.bssm_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bayesssm_b200 import _native
    with pytest.raises(_native.EngineError) as e:
        _native.Context(0)
    assert e.value.status == _native.ERR_NO_DEVICE


def test_peer_exchange_entry_points_refuse_bad_arguments_without_a_gpu():
    """bssm_shard_peer_* (the sharded filter's exchange through peer memory): argument errors are reported, not crashes."""
    import ctypes as C
    from bayesssm_b200 import _native
    lib = _native.load_library()
    buf = (C.c_ubyte * 64)()
    assert lib.bssm_shard_peer_export(None, buf) == _native.ERR_BAD_ARG
    assert "peer_export" in _native.last_error()
    assert lib.bssm_shard_peer_attach(None, buf) == _native.ERR_BAD_ARG
    assert lib.bssm_shard_peer_active(None) == 0
    assert lib.bssm_shard_peer_detach(None) == _native.OK


def test_host_side_transforms_match_oracle(orc):
    from bayesssm_b200 import _native
    lib = _native.load_library()
    for tr in (0, 1, 2):
        for th in (0.2, 0.5, 0.9):
            assert lib.bssm_transform(th, tr) == orc.transform(th, tr)
            z = orc.transform(th, tr)
            assert lib.bssm_back_transform(z, tr) == orc.back_transform(z, tr)
    for kind, a, b in ((0, 0, 0), (1, 0, 1), (1, 0, 10), (2, 1, 0), (3, 0, 1), (4, 1, 0), (4, 2, 0)):
        for x in (-0.5, 0.0, 0.3, 1.0, 2.5):
            u, v = lib.bssm_log_prior(kind, a, b, x), orc.log_prior(kind, a, b, x)
            assert u == v or abs(u - v) < 1e-15


def test_r_shim_compiles_against_stub_r_headers():
    """R is not installed here: the .Call shim is at least type-checked against the ABI header and a minimal
    restatement of the R C API declarations it uses."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "r_stub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r_shim", "src", "bssm_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every routine the shim registers exists in the file, and every bssm_* it calls is declared in the header
    src = open(os.path.join(ROOT, "r_shim", "src", "bssm_shim.c")).read()
    called = set(re.findall(r"\b(bssm_[a-z_0-9]+)\s*\(", src))
    assert called <= set(_header_symbols()), called - set(_header_symbols())


def test_nvrtc_compiles_a_snippet_with_all_model_kernels_without_a_gpu():
    """NVRTC needs no GPU to compile: the embedded header text (general kernels + streaming kernels) must stay free of
    host headers and compile for sm_100a together with a user snippet -- exactly what bssm_model_compile does."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "nvrtc_check.py")], capture_output=True, text=True, timeout=300)
    if "libnvrtc" in (r.stderr or "") and "cannot open shared object" in r.stderr:
        pytest.skip("libnvrtc not installed")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "cubin bytes" in r.stdout


def _strip_r(src):
    """R source with comments and string contents blanked (strings keep their quotes)."""
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c == "#":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if c in "\"'":
            q = c
            out.append(q)
            i += 1
            while i < n and src[i] != q:
                i += 2 if src[i] == "\\" else 1
            out.append(q)
            i += 1
            continue
        out.append(c)
        i += 1
    return "".join(out)


def test_r_frontends_are_well_formed_and_bind_registered_routines():
    """R is not installed here, so the R front-ends (r_shim/R/b200_frontends.R) get the checks that need no
    interpreter: brackets balance outside strings and comments, every .Call() target is a routine the shim registers
    (r_shim/src/bssm_shim.c CallEntries, the replacement of src/RcppExports.cpp:50-60) with the arity used, and every
    registered routine is reachable from R or is one of the reference's three resampler symbols (R/RcppExports.R:4-14)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    raw = open(os.path.join(root, "r_shim", "R", "b200_frontends.R")).read()
    src = _strip_r(raw)
    stack, pairs = [], {")": "(", "]": "[", "}": "{"}
    for ch in src:
        if ch in "([{":
            stack.append(ch)
        elif ch in ")]}":
            assert stack and stack.pop() == pairs[ch]
    assert not stack
    shim = open(os.path.join(root, "r_shim", "src", "bssm_shim.c")).read()
    registered = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(_bayesSSM_\w+)",\s*\(DL_FUNC\)&\1,\s*(\d+)\}', shim)}
    assert len(registered) >= 10
    used = {}
    for m in re.finditer(r'\.Call\("(_bayesSSM_\w+)"', raw):
        # count the top-level commas of this call's argument list
        i, depth, commas = m.end(), 1, 0
        while depth:
            ch = raw[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            commas += ch == "," and depth == 1
            i += 1
        used[m.group(1)] = commas
    for name, nargs in used.items():
        assert name in registered, name
        assert registered[name] == nargs, (name, registered[name], nargs)
    resamplers = {f"_bayesSSM_resample_{k}_cpp" for k in ("multinomial", "stratified", "systematic")}
    assert resamplers <= set(registered)
    assert set(registered) - set(used) <= resamplers


def test_r_config_lists_carry_the_names_the_shim_reads():
    """Each .Call passes a named list `cfg`; a name the shim looks up but R never sets would come back as NULL.
    Required names (read with list_get) must all be set by the matching front-end; optional ones (opt_int /
    opt_real, which have defaults) may be absent; R must not set a name nobody reads."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = open(os.path.join(root, "r_shim", "src", "bssm_shim.c")).read()
    raw = open(os.path.join(root, "r_shim", "R", "b200_frontends.R")).read()
    parts = re.split(r"\nSEXP (_bayesSSM_\w+)\(", shim)
    reads = {}
    for i in range(1, len(parts), 2):
        body = parts[i + 1]
        req = set(re.findall(r'list_get\(\s*cfg_?\s*,\s*"(\w+)"', body))
        opt = set(re.findall(r'opt_(?:int|real)\(\s*cfg_?\s*,\s*"(\w+)"', body))
        if req or opt:
            reads[parts[i]] = (req, opt)
    checked = 0
    for m in re.finditer(r"cfg <- list\(", raw):
        i, depth = m.end(), 1
        while depth:
            depth += raw[i] in "([{"
            depth -= raw[i] in ")]}"
            i += 1
        body = raw[m.end():i - 1]
        keys = set()
        for seg in re.finditer(r"(\w+)\s*=(?!=)", body):
            before = body[:seg.start()]
            if sum(ch in "([{" for ch in before) == sum(ch in ")]}" for ch in before):
                keys.add(seg.group(1))
        call = re.search(r'\.Call\("(_bayesSSM_\w+)",\s*cfg\b', raw[i:])
        req, opt = reads[call.group(1)]
        nullable = {"obs_times", "consts"}                     # the shim tests these for R_NilValue
        assert req - nullable <= keys, (call.group(1), req - keys)
        assert keys <= req | opt, (call.group(1), keys - req - opt)
        checked += 1
    assert checked == 3


def test_r_frontends_only_read_result_fields_the_shim_returns():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = open(os.path.join(root, "r_shim", "src", "bssm_shim.c")).read()
    raw = open(os.path.join(root, "r_shim", "R", "b200_frontends.R")).read()
    parts = re.split(r"\nSEXP (_bayesSSM_\w+)\(", shim)
    returned = {parts[i]: {x for lst in re.findall(r"names\w*\[\]\s*=\s*\{([^}]*)\}", parts[i + 1]) for x in re.findall(r'"(\w+)"', lst)}
                for i in range(1, len(parts), 2)}
    seen = 0
    for m in re.finditer(r'r <- \.Call\("(_bayesSSM_\w+)"', raw):
        nxt = re.search(r"\n[A-Za-z_.][\w.]* <- function", raw[m.end():])
        chunk = raw[m.end(): m.end() + (nxt.start() if nxt else len(raw))]
        used = set(re.findall(r"\br\$(\w+)", chunk))
        assert used and used <= returned[m.group(1)], (m.group(1), used - returned[m.group(1)])
        seen += 1
    assert seen >= 5


def _layout_from_c(tmp_path, header, structs, tag):
    """offsetof / sizeof of every field the ctypes mirror names, as the C compiler lays the header's struct out."""
    import subprocess
    lines = ["#include <stdio.h>", "#include <stddef.h>", f'#include "{header}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} * %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu %zu\\n", offsetof({cname}, {fname}), sizeof((({cname} *)0)->{fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / f"layout_{tag}.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / f"layout_{tag}"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    return [ln.split() for ln in out.strip().splitlines()]


def _check_layout(rows, structs):
    import ctypes as C
    n = 0
    for row in rows:
        cls = structs[row[0]]
        if row[1] == "*":
            assert C.sizeof(cls) == int(row[2]), row
        else:
            f = getattr(cls, row[1])
            assert (f.offset, f.size) == (int(row[2]), int(row[3])), row
            n += 1
    assert n == sum(len(c._fields_) for c in structs.values())
    for cname, cls in structs.items():                         # and no field of the C struct is missing from the mirror
        last = max(getattr(cls, f).offset + getattr(cls, f).size for f, _ in cls._fields_)
        assert C.sizeof(cls) - last < 8, cname


def test_ctypes_mirrors_have_the_layout_of_the_c_structs(tmp_path):
    """The Python side passes structs by pointer: every field of the ctypes mirrors (bayesssm_b200/_native.py for
    include/bayesssm_b200.h, tests/oracle.py for oracle/pf_oracle.h) must sit at the offset the C compiler gives it."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from bayesssm_b200 import _native as nat
    structs = {"bssm_noise_buffers": nat.NoiseBuffers, "bssm_filter_config": nat.FilterConfig,
               "bssm_filter_result": nat.FilterResult, "bssm_pmmh_config": nat.PmmhConfig, "bssm_pmmh_result": nat.PmmhResult}
    _check_layout(_layout_from_c(tmp_path, os.path.join(root, "include", "bayesssm_b200.h"), structs, "abi"), structs)
    import oracle
    ostructs = {"orc_noise_buffers": oracle.NoiseBuffers, "orc_filter_config": oracle.FilterConfig,
                "orc_filter_result": oracle.FilterResult, "orc_pmmh_config": oracle.PmmhConfig,
                "orc_pmmh_chain_result": oracle.PmmhChainResult}
    _check_layout(_layout_from_c(tmp_path, os.path.join(root, "oracle", "pf_oracle.h"), ostructs, "oracle"), ostructs)


def test_ctypes_prototypes_have_the_arity_and_scalar_kinds_of_the_header():
    """Every prototype in include/bayesssm_b200.h against bayesssm_b200/_native.py SYMBOLS: same number of
    parameters; pointers bound as pointers, integers as integers, doubles as doubles."""
    import ctypes as C
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from bayesssm_b200 import _native as nat
    text = re.sub(r"/\*.*?\*/", " ", open(os.path.join(root, "include", "bayesssm_b200.h")).read(), flags=re.S)
    protos = re.findall(r"\b[\w\s\*]+?\b(bssm_\w+)\s*\(([^;{}]*?)\)\s*;", text)
    assert len(protos) >= 30
    for name, params in protos:
        params = [p.strip() for p in params.split(",")] if params.strip() not in ("", "void") else []
        assert name in nat.SYMBOLS, name
        argtypes = nat.SYMBOLS[name][1]
        assert len(argtypes) == len(params), (name, params, argtypes)
        for p, a in zip(params, argtypes):
            is_ptr = "*" in p
            py_ptr = a in (C.c_void_p, C.c_char_p) or hasattr(a, "contents") or getattr(a, "_type_", None) == "P"
            assert is_ptr == bool(py_ptr), (name, p, a)
            if not is_ptr:
                kind = p.split()[-2] if len(p.split()) > 1 else p
                if kind in ("double",):
                    assert a is C.c_double, (name, p, a)
                elif kind in ("float",):
                    assert a is C.c_float, (name, p, a)
                else:
                    assert a in (C.c_int, C.c_int32, C.c_uint32, C.c_int64, C.c_uint64, C.c_size_t), (name, p, a)
                    assert C.sizeof(a) == {"int": 4, "int32_t": 4, "uint32_t": 4, "int64_t": 8, "uint64_t": 8, "size_t": 8}[kind], (name, p, a)
