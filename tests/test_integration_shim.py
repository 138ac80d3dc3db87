"""The Rust shim (integration/rust, source only: no Rust toolchain in this image) against the C ABI it binds: the files
are the code blocks of INTEGRATION.md, every extern "C" function they declare exists in include/galah_b200.h and in the
built library, with the same number of arguments."""
import os
import re

import galah_b200 as gb
from conftest import ROOT

SHIM = os.path.join(ROOT, "integration", "rust")


def _strip_banner(text):
    lines = text.split("\n")
    while lines and lines[0].startswith("//"):
        lines.pop(0)
    return "\n".join(lines)


def test_shim_files_are_the_blocks_of_integration_md():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```rust\n(.*?)```", md, re.S)
    files = ["build.rs", "src/b200.rs", "src/b200_batched.rs", "src/b200_multi.rs"]
    assert len(blocks) == len(files)
    for block, name in zip(blocks, files):
        assert _strip_banner(open(os.path.join(SHIM, name)).read()) == block, name


def _c_prototypes():
    h = open(os.path.join(ROOT, "include", "galah_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(galah_b200_\w+)\s*\(([^;{]*?)\)\s*;", h, re.S):
        if "(*" in m.group(0).split(m.group(1))[0][-20:]:
            continue
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_args(args))
    return protos


def _split_args(args):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_every_declared_symbol_exists_with_the_same_arity():
    protos = _c_prototypes()
    exported = set(gb.exported_symbols())
    seen = 0
    for name in ("src/b200.rs", "src/b200_batched.rs", "src/b200_multi.rs"):
        text = open(os.path.join(SHIM, name)).read()
        for m in re.finditer(r"\bfn (galah_b200_\w+)\s*\((.*?)\)\s*(?:->\s*[\w:*\s]+)?;", text, re.S):
            fn, args = m.group(1), m.group(2).strip()
            n_args = 0 if not args else len(_split_args(args))
            assert fn in protos, f"{fn} is not declared in include/galah_b200.h"
            assert fn in exported, f"{fn} is not exported by the library"
            assert n_args == protos[fn], f"{fn}: {n_args} arguments in the shim, {protos[fn]} in the header"
            seen += 1
        for fn in set(re.findall(r"\b(galah_b200_\w+)\s*\(", text)):
            assert fn in protos, f"{fn} used in {name} but not in the header"
    assert seen >= 12
