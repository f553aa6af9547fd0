"""ctypes binding of libcoopcap.so, generated from include/coopcap.h.

The struct layouts and function prototypes are parsed out of the header at import time, so the
Python side cannot drift from the C ABI; `load()` additionally checks every struct size against
`coopcap_sizeof`.  There is no CPU fallback and no alternative backend: if the library is missing
or a call fails the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcoopcap.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "coopcap.h")


class CoopcapError(RuntimeError):
    pass


_SCALARS = {
    "int": C.c_int, "float": C.c_float, "double": C.c_double, "int64_t": C.c_int64, "uint64_t": C.c_uint64,
    "coopcap_stream_t": C.c_void_p, "long long": C.c_longlong,
}


def _ctype_of(decl_type: str, structs):
    t = decl_type.replace("const", "").strip()
    if t.endswith("*"):
        base = t[:-1].strip()
        if base in structs:
            return C.POINTER(structs[base])
        if base == "int" and False:
            return C.POINTER(C.c_int)
        return C.c_void_p
    if t in _SCALARS:
        return _SCALARS[t]
    raise CoopcapError(f"coopcap.h: unsupported type {decl_type!r}")


def _parse_header(path):
    with open(path, "r") as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    structs, order = {}, []
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        name, body = m.group(3), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"^(.*?)([\w\s,\*]+)$", decl)
            # split "type a, b, c": the type is everything up to the first declarator
            parts = decl.split(",")
            first = parts[0].strip()
            fm = re.match(r"^(.*?[\s\*])(\w+)$", first)
            if not fm:
                raise CoopcapError(f"coopcap.h: cannot parse field {decl!r} of {name}")
            ftype = fm.group(1).strip()
            names = [fm.group(2)] + [p.strip() for p in parts[1:]]
            for n in names:
                t = ftype
                while n.startswith("*"):
                    t, n = t + "*", n[1:].strip()
                fields.append((n, _ctype_of(t, structs)))
        cls = type(name, (C.Structure,), {"_fields_": fields})
        structs[name] = cls
        order.append(name)
    funcs = {}
    body = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
    for m in re.finditer(r"\b(int|long long|uint64_t|const char\s*\*)\s+(coopcap_\w+)\s*\(([^)]*)\)\s*;", body):
        ret, fname, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                am = re.match(r"^(.*?[\s\*])(\w+)$", a)
                argtypes.append(_ctype_of(am.group(1).strip(), structs))
        restype = C.c_char_p if "char" in ret else (C.c_longlong if "long" in ret else
                                                    (C.c_uint64 if "uint64" in ret else C.c_int))
        funcs[fname] = (restype, argtypes)
    return structs, order, funcs


STRUCTS, STRUCT_ORDER, FUNCTIONS = _parse_header(HEADER_PATH)

GemmArgs = STRUCTS["coopcap_gemm_args"]
SpeakerPack = STRUCTS["coopcap_speaker_pack"]
Speaker = STRUCTS["coopcap_speaker"]
SpeakerGrads = STRUCTS["coopcap_speaker_grads"]
ListenerPack = STRUCTS["coopcap_listener_pack"]
Listener = STRUCTS["coopcap_listener"]
ListenerGrads = STRUCTS["coopcap_listener_grads"]
Cider = STRUCTS["coopcap_cider"]
Beam = STRUCTS["coopcap_beam"]

_SIZEOF_IDS = {"coopcap_gemm_args": 0, "coopcap_speaker_pack": 1, "coopcap_speaker": 2,
               "coopcap_speaker_grads": 3, "coopcap_listener_pack": 4, "coopcap_listener": 5,
               "coopcap_listener_grads": 6, "coopcap_cider": 7, "coopcap_beam": 8}

# kept for tests: name -> argtypes
SIGNATURES = {k: v[1] for k, v in FUNCTIONS.items()}

_lib = None


def load() -> C.CDLL:
    """Load libcoopcap.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CoopcapError(
            f"{LIB_PATH} not found: build it with `python -m cooperativeimagecaptioning_b200.build` "
            "(or __graft_entry__.build()). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in FUNCTIONS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise CoopcapError(f"libcoopcap.so does not export {name} (stale build?)") from e
        fn.restype = restype
        fn.argtypes = argtypes
    for sname, sid in _SIZEOF_IDS.items():
        got, want = C.sizeof(STRUCTS[sname]), lib.coopcap_sizeof(sid)
        if got != want:
            raise CoopcapError(f"struct {sname}: ctypes size {got} != C size {want}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().coopcap_last_error().decode("utf-8", "replace")
        raise CoopcapError(f"libcoopcap call failed (code {rc}): {msg}")
