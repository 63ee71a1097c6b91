"""Regenerates the ctypes struct snippets of INTEGRATION.md from pytorch_stable_diffusion_b200/_ext.py, so the
documented binding cannot drift from the one the package (and every -m gpu test) uses.

    python tools/gen_integration.py [--check]     # --check: exit 1 if INTEGRATION.md is stale
"""
import ctypes
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytorch_stable_diffusion_b200 import _ext  # noqa: E402

NAMES = {ctypes.c_int: "ctypes.c_int", ctypes.c_void_p: "ctypes.c_void_p", ctypes.c_longlong: "ctypes.c_longlong",
         ctypes.c_float: "ctypes.c_float"}


def render(struct, c_name):
    lines = [f"class {struct.__name__}(ctypes.Structure):            # mirrors {c_name} field for field "
             f"({ctypes.sizeof(struct)} bytes == sdb_args_size())", "    _fields_ = ["]
    row = "        "
    for name, typ in struct._fields_:
        item = f'("{name}", {NAMES[typ]}), '
        if len(row) + len(item) > 112:
            lines.append(row.rstrip())
            row = "        "
        row += item
    lines.append(row.rstrip().rstrip(",") + "]")
    return "\n".join(lines)


def main():
    path = os.path.join(ROOT, "INTEGRATION.md")
    text = open(path).read()
    new = text
    for struct, c_name in ((_ext.GemmArgs, "sdb_gemm_args"), (_ext.AttnArgs, "sdb_attn_args")):
        pat = re.compile(rf"(# BEGIN GENERATED {struct.__name__}\n).*?(# END GENERATED {struct.__name__})", re.S)
        if not pat.search(new):
            raise SystemExit(f"INTEGRATION.md has no generated block for {struct.__name__}")
        new = pat.sub(lambda m: m.group(1) + render(struct, c_name) + "\n" + m.group(2), new)
    if "--check" in sys.argv:
        if new != text:
            raise SystemExit("INTEGRATION.md is stale: run python tools/gen_integration.py")
        return
    if new != text:
        open(path, "w").write(new)
        print("INTEGRATION.md updated")


if __name__ == "__main__":
    main()
