"""Parse include/mi_b200.h into {name: [param type strings]} (test helper)."""
import os
import re

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mi_b200.h")


def parse_header(path=HEADER):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(int|size_t|const char\*|unsigned long long)\s+(mi_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        plist = [] if params in ("void", "") else [" ".join(p.split()) for p in params.split(",")]
        out[name] = (ret, plist)
    return out
