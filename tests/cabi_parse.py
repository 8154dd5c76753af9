"""A small parser of include/chrom_cuda.h (structs, prototypes, constants) and of the Rust declarations in
bindings/rust/chrom-cuda/src/ffi.rs, used to tie the two together mechanically (no Rust toolchain exists here, so
this is the only executable evidence that the crate binds the header as written)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "chrom_cuda.h")
FFI_RS = os.path.join(ROOT, "bindings", "rust", "chrom-cuda", "src", "ffi.rs")

C_TO_RUST = {"double": "f64", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "size_t": "usize",
             "char": "c_char", "void": "c_void"}


def strip_c_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def rust_type(ctype: str) -> str:
    """'const chrom_run_cfg*' -> '*const chrom_run_cfg', 'chrom_cuda_ctx**' -> '*mut *mut chrom_cuda_ctx'."""
    t = ctype.strip()
    stars = t.count("*")
    const = t.startswith("const ")
    base = t.replace("const ", "").replace("*", "").strip()
    base = C_TO_RUST.get(base, base)
    if stars == 0:
        return base
    out = ("*const " if const else "*mut ") + base
    for _ in range(stars - 1):
        out = "*mut " + out
    return out


def parse_header(path: str = HEADER) -> dict:
    src = strip_c_comments(open(path).read())
    consts = {}
    for k, v in re.findall(r"#define\s+(CHROM_[A-Z0-9_]+)\s+(\(?-?\d+u?\)?)", src):
        consts[k] = int(v.strip("()").rstrip("u"))
    structs = {}
    for body, name in re.findall(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            m = re.match(r"((?:const\s+)?\w+\s*\**)\s*(.*)", decl)
            ctype, names = m.group(1).strip(), m.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                extra = nm.count("*")
                fields.append((nm.replace("*", "").strip(), rust_type(ctype + "*" * extra)))
        structs[name] = fields
    opaque = re.findall(r"typedef\s+struct\s+(\w+)\s+\w+\s*;", src)
    funcs = {}
    for ret, name, args in re.findall(r"\n\s*((?:const\s+)?\w+\s*\**)\s*(chrom_cuda_\w+)\s*\(([^)]*)\)\s*;", src):
        arglist = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"(.*?)(\w+)$", a)
                arglist.append((m.group(2), rust_type(m.group(1))))
        r = " ".join(ret.split())
        funcs[name] = (arglist, None if r == "void" else rust_type(r))
    return {"consts": consts, "structs": structs, "opaque": opaque, "funcs": funcs}


def parse_ffi_rs(path: str = FFI_RS) -> dict:
    src = re.sub(r"//[^\n]*", "", open(path).read())
    consts = {k: (t, int(v)) for k, t, v in re.findall(r"pub const (CHROM_[A-Z0-9_]+):\s*(\w+)\s*=\s*(-?\d+);", src)}
    structs, repr_c = {}, set()
    for attrs, name, body in re.findall(r"((?:#\[[^\]]*\]\s*)*)pub struct (\w+)\s*\{(.*?)\}", src, flags=re.S):
        if "repr(C)" in attrs:
            repr_c.add(name)
        fields = [(n, " ".join(t.split())) for n, t in re.findall(r"(?:pub\s+)?(\w+)\s*:\s*([^,\n]+),", body)]
        structs[name] = fields
    funcs = {}
    ext = re.search(r'extern\s+"C"\s*\{(.*)\}', src, flags=re.S).group(1)
    for name, args, ret in re.findall(r"pub fn (\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", ext, flags=re.S):
        arglist = [(n, " ".join(t.split())) for n, t in re.findall(r"(\w+)\s*:\s*([^,]+?)\s*(?:,|$)", " ".join(args.split()))]
        funcs[name] = (arglist, " ".join(ret.split()) if ret else None)
    return {"consts": consts, "structs": structs, "repr_c": repr_c, "funcs": funcs}
