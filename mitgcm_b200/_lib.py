"""ctypes loader of libmitgcm_b200.so (the C-ABI drop-in boundary, include/mitgcm_b200.h).

There is no CPU fallback: importing works anywhere (so the symbol table can be
checked on a CPU box), but every compute entry point needs a CUDA device and the
library reports an error otherwise."""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.environ.get("MITGCM_B200_LIB") or os.path.join(HERE, "lib", "libmitgcm_b200.so")
HEADER = os.path.join(HERE, "..", "include", "mitgcm_b200.h")

PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int)


def parse_enums(path=HEADER):
    """{NAME: value} for every enumerator of include/mitgcm_b200.h."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", txt, flags=re.S):
        v = -1
        for item in body.split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = [s.strip() for s in item.split("=")]
                v = int(val)
            else:
                name, v = item, v + 1
            out[name] = v
    return out


def declared_functions(path=HEADER):
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(\w+)\s*\(", " ".join(
        l for l in txt.splitlines() if not l.strip().startswith("#"))))
        - {"enum"})


ENUMS = parse_enums()
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} is missing: run `python -m mitgcm_b200.build` "
                               "(there is no CPU fallback for the B200 path)")
        L = C.CDLL(LIB)
        L.mitgcm_b200_last_error_string.restype = C.c_char_p
        L.mitgcm_b200_field_ptr.restype = C.c_void_p
        L.mitgcm_b200_field_ptr.argtypes = [C.c_int]
        _lib = L
    return _lib
