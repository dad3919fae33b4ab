"""Checks bindings/rust/b200.rs against include/rt_b200.h without a Rust compiler.

For every #[repr(C)] struct of the .rs file: the C struct of the same name must exist and have the same fields, in the
same order, with equivalent types. For every function in its extern "C" block: the C prototype of the same name must
exist with the same number of parameters, equivalent parameter types and return type. Also: the struct sizes implied by
the Rust declarations (natural alignment, as #[repr(C)] lays them out) must equal the ctypes structures' sizes, which a
CPU test compares with the built library's view (tests/test_abi.py), and RT_B200_ABI_VERSION must match.

    python tools/check_rust_abi.py            # prints a report, exit code 1 on any mismatch
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C_TO_CANON = {"int32_t": "i32", "uint32_t": "u32", "int64_t": "i64", "uint64_t": "u64", "uint8_t": "u8", "int": "i32", "double": "f64",
              "float": "f32", "char": "i8", "void": "void", "size_t": "usize"}
RS_TO_CANON = {"i32": "i32", "u32": "u32", "i64": "i64", "u64": "u64", "u8": "u8", "c_int": "i32", "c_double": "f64", "c_float": "f32",
               "c_char": "i8", "c_void": "void", "f32": "f32", "f64": "f64", "usize": "usize"}
SIZES = {"i32": 4, "u32": 4, "i64": 8, "u64": 8, "u8": 1, "i8": 1, "f64": 8, "f32": 4, "usize": 8}


def strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def canon_c_type(t):
    """'const rt_scene_desc*' -> ('ptr', 'rt_scene_desc'); 'double' -> 'f64'; 'rt_context* const*' -> ptr to ptr"""
    t = t.replace("const", " ").strip()
    depth = t.count("*")
    base = t.replace("*", " ").split()
    base = base[-1] if base else "void"
    base = C_TO_CANON.get(base, base)
    out = base
    for _ in range(depth):
        out = ("ptr", out)
    return out


def canon_rs_type(t):
    t = t.strip()
    m = re.match(r"\*(?:const|mut)\s+(.*)", t)
    if m:
        return ("ptr", canon_rs_type(m.group(1)))
    m = re.match(r"\[(.*);\s*(\d+)\]$", t)
    if m:
        return ("arr", canon_rs_type(m.group(1)), int(m.group(2)))
    return RS_TO_CANON.get(t, t)


def parse_c(path):
    src = strip_c_comments(open(path).read())
    structs, funcs = {}, {}
    for m in re.finditer(r"typedef struct (\w+)\s*\{(.*?)\}\s*\w+;", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"(.*?)([\w\[\]\s,]+)$", decl)
            # split type and declarators: type is everything up to the last '*' or first identifier list
            tm = re.match(r"((?:const\s+)?\w+(?:\s*\*+)?)\s*(.*)$", decl)
            ctype, names = tm.group(1), tm.group(2)
            for name in names.split(","):
                name = name.strip()
                dims = [int(x) for x in re.findall(r"\[(\d+)\]", name)]
                name = re.sub(r"\[.*", "", name).replace("*", "").strip()
                t = canon_c_type(ctype + ("*" if "*" in names and "," not in names else ""))
                t = canon_c_type(ctype)
                for d in reversed(dims):
                    t = ("arr", t, d)
                fields.append((name, t))
        structs[m.group(1)] = fields
    for m in re.finditer(r"(?:^|\n)\s*((?:const\s+)?\w+(?:\s*\*)?)\s+\*?(rt_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, params = m.group(1), m.group(2), " ".join(m.group(3).split())
        if "typedef" in ret:
            continue
        if src[m.start(2) - 1] == "*":
            ret += "*"
        plist = []
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                arr = "[" in p
                p = re.sub(r"\[.*?\]", "", p)
                pm = re.match(r"(.*?)(\w+)$", p)
                t = canon_c_type(pm.group(1))
                if arr:
                    t = ("ptr", t)
                plist.append(t)
        funcs[name] = (canon_c_type(ret), plist)
    ver = int(re.search(r"#define RT_B200_ABI_VERSION (\d+)", src).group(1))
    return structs, funcs, ver


def parse_rs(path):
    src = re.sub(r"//.*", "", open(path).read())
    structs, funcs = {}, {}
    for m in re.finditer(r"#\[repr\(C\)\](?:\s*#\[derive\([^\]]*\)\])?\s*pub struct (\w+)\s*\{(.*?)\}", src, flags=re.S):
        fields = []
        for f in re.finditer(r"pub\s+(\w+)\s*:\s*([^,\n]+(?:;\s*\d+\]\s*;\s*\d+\])?[^,\n]*)", m.group(2)):
            fields.append((f.group(1), canon_rs_type(f.group(2).strip().rstrip(","))))
        structs[m.group(1)] = fields
    ext = re.search(r'extern\s+"C"\s*\{(.*?)\n\}', src, flags=re.S).group(1)
    for m in re.finditer(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", ext, flags=re.S):
        params = " ".join(m.group(2).split())
        plist = []
        depth, cur = 0, ""
        for ch in params + ",":
            if ch == "," and depth == 0:
                if cur.strip():
                    plist.append(canon_rs_type(cur.split(":", 1)[1]))
                cur = ""
            else:
                depth += ch in "[(<"
                depth -= ch in "])>"
                cur += ch
        ret = canon_rs_type(m.group(3)) if m.group(3) else "void"
        funcs[m.group(1)] = (ret, plist)
    ver = int(re.search(r"RT_B200_ABI_VERSION: i32 = (\d+)", src).group(1))
    return structs, funcs, ver


def size_align(t, structs):
    if isinstance(t, tuple):
        if t[0] == "ptr":
            return 8, 8
        if t[0] == "arr":
            s, a = size_align(t[1], structs)
            return s * t[2], a
    if t in SIZES:
        return SIZES[t], SIZES[t]
    off, al = 0, 1
    for _, ft in structs[t]:
        s, a = size_align(ft, structs)
        off = (off + a - 1) // a * a + s
        al = max(al, a)
    return (off + al - 1) // al * al, al


def equivalent(a, b):
    """pointer targets that are opaque structs compare by name; `void` pointers match any pointer"""
    if isinstance(a, tuple) and isinstance(b, tuple):
        if a[0] != b[0]:
            return False
        if a[0] == "ptr":
            return a[1] == "void" or b[1] == "void" or equivalent(a[1], b[1])
        return a[2] == b[2] and equivalent(a[1], b[1])
    return a == b


def main():
    cs, cf, cver = parse_c(os.path.join(ROOT, "include", "rt_b200.h"))
    rs, rf, rver = parse_rs(os.path.join(ROOT, "bindings", "rust", "b200.rs"))
    errors = []
    if cver != rver:
        errors.append(f"RT_B200_ABI_VERSION: header {cver}, Rust {rver}")
    opaque = {"rt_builder", "rt_context", "rt_scene"}
    n_struct = 0
    for name, fields in rs.items():
        if name in opaque:
            continue
        n_struct += 1
        if name not in cs:
            errors.append(f"struct {name}: not in the header")
            continue
        if [f for f, _ in fields] != [f for f, _ in cs[name]]:
            errors.append(f"struct {name}: field names differ: Rust {[f for f, _ in fields]} vs C {[f for f, _ in cs[name]]}")
            continue
        for (fn, rt_), (_, ct) in zip(fields, cs[name]):
            if not equivalent(rt_, ct):
                errors.append(f"struct {name}.{fn}: Rust {rt_} vs C {ct}")
        if size_align(name, rs) != size_align(name, cs):
            errors.append(f"struct {name}: size/alignment {size_align(name, rs)} vs {size_align(name, cs)}")
    for name, (ret, params) in rf.items():
        if name not in cf:
            errors.append(f"fn {name}: not in the header")
            continue
        cret, cparams = cf[name]
        if not equivalent(ret, cret):
            errors.append(f"fn {name}: return Rust {ret} vs C {cret}")
        if len(params) != len(cparams):
            errors.append(f"fn {name}: {len(params)} parameters in Rust, {len(cparams)} in C")
            continue
        for k, (a, b) in enumerate(zip(params, cparams)):
            if not equivalent(a, b):
                errors.append(f"fn {name} parameter {k}: Rust {a} vs C {b}")
    sizes = {n: size_align(n, rs)[0] for n in rs if n not in opaque}
    print(f"checked {n_struct} structs and {len(rf)} functions of bindings/rust/b200.rs against include/rt_b200.h (ABI version {cver})")
    print("struct sizes:", sizes)
    for e in errors:
        print("MISMATCH:", e)
    return (1 if errors else 0), sizes, rs, rf


if __name__ == "__main__":
    sys.exit(main()[0])
