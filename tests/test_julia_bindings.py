"""The Julia host layer (julia/ClassicalSpinMC) cannot be executed in the build image (no Julia toolchain), so
its ccall signatures are checked statically against include/csmc.h: every bound symbol exists, the argument
count matches and each argument has the right machine type (pointer / Int32 / Int64 / Float64 / UInt64), and
the isbits mirrors of the ABI structs have the header's fields in order."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    src = open(os.path.join(ROOT, "include", "csmc.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"(?:int32_t|const char \*)\s*(csmc_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        kinds = []
        for a in ([] if args in ("", "void") else args.split(",")):
            a = a.strip()
            if "*" in a or "[" in a:
                kinds.append("ptr")
            elif re.match(r"(const )?int32_t\b", a):
                kinds.append("i32")
            elif re.match(r"(const )?int64_t\b", a):
                kinds.append("i64")
            elif re.match(r"(const )?uint64_t\b", a):
                kinds.append("u64")
            elif re.match(r"(const )?double\b", a):
                kinds.append("f64")
            else:
                raise AssertionError(f"unclassified C argument {a!r} in {name}")
        protos[name] = kinds
    return protos


def _split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _julia_ccalls():
    calls = []
    for path in glob.glob(os.path.join(ROOT, "julia", "ClassicalSpinMC", "src", "*.jl")):
        src = open(path).read()
        for m in re.finditer(r"ccall\(\(:(csmc_\w+), libcsmc\),\s*(\w+),\s*\(", src):
            i, depth = m.end(), 1
            while depth:
                depth += {"(": 1, ")": -1}.get(src[i], 0)
                i += 1
            kinds = []
            for t in _split_top_level(src[m.end():i - 1]):
                if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
                    kinds.append("ptr")
                else:
                    kinds.append({"Int32": "i32", "Cint": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64"}[t])
            calls.append((os.path.basename(path), m.group(1), m.group(2), kinds))
    return calls


def test_every_julia_ccall_matches_the_header():
    protos = _header_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 25 and len(protos) >= 45
    for path, name, ret, kinds in calls:
        assert name in protos, f"{path}: {name} is not declared in include/csmc.h"
        assert kinds == protos[name], f"{path}: {name} bound as {kinds}, header says {protos[name]}"
        assert ret == ("Cstring" if name == "csmc_last_error" else "Int32"), f"{path}: {name} return type {ret}"
    bound = {c[1] for c in calls}
    # the drivers need at least these
    assert {"csmc_create", "csmc_destroy", "csmc_set_spins", "csmc_get_spins", "csmc_total_energy", "csmc_overrelax",
            "csmc_metropolis", "csmc_deterministic", "csmc_anneal_temperature", "csmc_pt_init", "csmc_pt_run",
            "csmc_comm_init", "csmc_pt_get_series"} <= bound


def _c_struct_fields(name):
    src = open(os.path.join(ROOT, "include", "csmc.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + r";", src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        fields.append(re.search(r"(\w+)\s*(?:\[[^\]]*\])?$", decl).group(1))
    return fields


def _julia_struct_fields(name):
    src = open(os.path.join(ROOT, "julia", "ClassicalSpinMC", "src", "libcsmc.jl")).read()
    body = re.search(r"struct " + name + r"\b(.*?)\bend", src, flags=re.S).group(1)
    return [m.group(1) for m in re.finditer(r"(\w+)::", body)]


def test_julia_struct_mirrors_follow_the_header_field_order():
    for c_name, j_name in (("csmc_model", "CsmcModel"), ("csmc_opts", "CsmcOpts"), ("csmc_pt_params", "CsmcPtParams")):
        assert _julia_struct_fields(j_name) == _c_struct_fields(c_name), (c_name, j_name)
