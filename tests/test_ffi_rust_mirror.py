"""include/chrom_cuda.h <-> bindings/rust/chrom-cuda/src/ffi.rs <-> chromatography_b200/_ffi.py, mechanically.

The Rust crate cannot be compiled here (no rustc/cargo in the image), so the declarations it binds are compared
with the header by parsing both: every constant, every struct (field names, types, order, #[repr(C)]) and every
prototype (argument names, types, order, return type) must agree.  The ctypes mirror the tests and the bench call
through is held to the same header: struct fields, sizes computed from the C layout rules, and argument types."""
import ctypes as C

import cabi_parse as P
from chromatography_b200 import _ffi

H = P.parse_header()
R = P.parse_ffi_rs()


def test_every_constant_is_mirrored():
    assert len(H["consts"]) >= 20
    for k, v in H["consts"].items():
        assert k in R["consts"], f"{k} missing from ffi.rs"
        assert R["consts"][k][1] == v, (k, R["consts"][k], v)
    assert set(R["consts"]) == set(H["consts"]), set(R["consts"]) ^ set(H["consts"])


def test_every_struct_matches_field_by_field():
    assert set(H["structs"]) == {"chrom_single_col", "chrom_multi_col", "chrom_species", "chrom_run_cfg", "chrom_cuda_timing"}
    for name, fields in H["structs"].items():
        assert name in R["structs"], f"struct {name} missing from ffi.rs"
        assert name in R["repr_c"], f"struct {name} is not #[repr(C)]"
        assert R["structs"][name] == fields, (name, R["structs"][name], fields)
    for name in H["opaque"]:
        assert name in R["structs"] and name in R["repr_c"]


def test_every_prototype_matches():
    assert len(H["funcs"]) >= 30
    assert list(R["funcs"]) == list(H["funcs"]), "ffi.rs must declare every function of the header, in its order"
    for name, (args, ret) in H["funcs"].items():
        rargs, rret = R["funcs"][name]
        assert rargs == args, (name, rargs, args)
        assert rret == ret, (name, rret, ret)


RUST_TO_CTYPES = {"f64": C.c_double, "i32": C.c_int32, "u32": C.c_uint32, "u64": C.c_uint64, "i64": C.c_int64,
                  "usize": C.c_size_t}
SIZES = {"f64": 8, "i32": 4, "u32": 4, "u64": 8, "i64": 8, "usize": 8}
PY_STRUCTS = {"chrom_single_col": _ffi.SingleCol, "chrom_multi_col": _ffi.MultiCol, "chrom_species": _ffi.Species,
              "chrom_run_cfg": _ffi.RunCfg, "chrom_cuda_timing": _ffi.Timing}


def c_layout_size(fields):
    off, align = 0, 1
    for _, t in fields:
        sz = 8 if t.startswith("*") else SIZES[t]
        off = (off + sz - 1) // sz * sz + sz
        align = max(align, sz)
    return (off + align - 1) // align * align


def test_ctypes_structs_match_the_header():
    for name, fields in H["structs"].items():
        py = PY_STRUCTS[name]
        assert [n.rstrip("_") if n != "_reserved" else n for n, _ in py._fields_] == [n for n, _ in fields], name
        for (pn, pt), (hn, ht) in zip(py._fields_, fields):
            if ht.startswith("*"):
                assert C.sizeof(pt) == 8 and issubclass(pt, (C._Pointer, C.c_void_p)), (name, hn)
            else:
                assert pt is RUST_TO_CTYPES[ht], (name, hn, pt, ht)
        assert C.sizeof(py) == c_layout_size(fields), (name, C.sizeof(py), c_layout_size(fields))


def test_ctypes_prototypes_match_the_header():
    assert sorted(_ffi.SYMBOLS) == sorted(H["funcs"])
    for name, (args, ret) in H["funcs"].items():
        res, argtypes = _ffi.SYMBOLS[name]
        assert len(argtypes) == len(args), name
        for at, (an, t) in zip(argtypes, args):
            if t.startswith("*"):
                assert C.sizeof(at) == 8, (name, an)  # a pointer of some kind (struct / double / void)
            else:
                assert at is RUST_TO_CTYPES[t], (name, an, at, t)
        if ret is None:
            assert res is None, name
        elif ret.startswith("*"):
            assert res is C.c_char_p, name
        else:
            assert res is RUST_TO_CTYPES[ret], (name, res, ret)
