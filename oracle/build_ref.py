"""Builds oracle/_ref/: the UNMODIFIED reference (/root/reference/sd/*.py) compiled to CPython bytecode.

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE. The reference is nine flat Python modules with no build
system of its own (`pip install /root/reference` has nothing to build), and its checkout does not travel to
the GPU box. This recipe compiles the modules the sampling path imports from the sources where they lie
(py_compile, no source copied) into oracle/_ref/sd/<name>.refbc - a build product, git-ignored, NOT
gpurun-ignored, so it ships with the snapshot like the repo's own built .so (the files are CPython .pyc images under
another suffix: the snapshot tool drops *.pyc) - plus a MANIFEST.json with the SHA-256 of every source it was
compiled from. `bench.py --impl reference` and the `cpu_baseline` leg import these modules through a small
meta-path finder (the GPU box runs the same image, hence the same CPython magic number) and run the reference's
own `pipeline.generate(device="cpu")`.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box: prebuilt files)
"""
import hashlib
import importlib.abc
import importlib.machinery
import importlib.util
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/sd"
OUT = os.path.join(HERE, "_ref", "sd")
SUFFIX = ".refbc"
MODULES = ("attention", "clip", "ddpm", "decoder", "diffusion", "encoder", "pipeline", "model_loader",
           "model_converter")


def build(force=False):
    """Returns the directory holding the compiled reference, or None when it cannot be built here and no
    prebuilt copy exists."""
    manifest_path = os.path.join(OUT, "MANIFEST.json")
    if not os.path.isdir(REF_SRC):
        return OUT if os.path.exists(manifest_path) else None
    digests = {}
    for name in MODULES:
        with open(os.path.join(REF_SRC, name + ".py"), "rb") as f:
            digests[name] = hashlib.sha256(f.read()).hexdigest()
    magic = importlib.util.MAGIC_NUMBER.hex()
    if not force and os.path.exists(manifest_path):
        old = json.load(open(manifest_path))
        if old.get("sha256") == digests and old.get("magic") == magic and \
                all(os.path.exists(os.path.join(OUT, n + SUFFIX)) for n in MODULES):
            return OUT
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:
        py_compile.compile(os.path.join(REF_SRC, name + ".py"), cfile=os.path.join(OUT, name + SUFFIX),
                           dfile=f"reference/sd/{name}.py", doraise=True, optimize=0)
    json.dump({"source": REF_SRC, "sha256": digests, "magic": magic, "python": sys.version.split()[0],
               "note": "bytecode of the unmodified reference; regenerate with python oracle/build_ref.py"},
              open(manifest_path, "w"), indent=1)
    return OUT


def available():
    return os.path.exists(os.path.join(OUT, "MANIFEST.json")) and \
        json.load(open(os.path.join(OUT, "MANIFEST.json"))).get("magic") == importlib.util.MAGIC_NUMBER.hex()


class _RefFinder(importlib.abc.MetaPathFinder):
    """Resolves the reference's flat module names (it imports its siblings as `from attention import ...`) to
    the compiled files under oracle/_ref/sd."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in MODULES:
            return None
        file = os.path.join(OUT, fullname + SUFFIX)
        if not os.path.exists(file):
            return None
        loader = importlib.machinery.SourcelessFileLoader(fullname, file)
        return importlib.util.spec_from_file_location(fullname, file, loader=loader)


def load():
    """Imports the compiled reference modules under their own flat names, isolated from the package under test
    (whose modules live in the pytorch_stable_diffusion_b200 namespace). Returns {name: module}."""
    import importlib
    if not available():
        raise RuntimeError("oracle/_ref is missing or was compiled by another CPython: run python oracle/build_ref.py "
                           "in the build container")
    clash = [n for n in MODULES if n in sys.modules and
             not str(getattr(sys.modules[n], "__file__", "")).startswith(OUT)]
    if clash:
        raise RuntimeError(f"modules {clash} are already imported from elsewhere")
    if not any(isinstance(f, _RefFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder())
    return {n: importlib.import_module(n) for n in MODULES if n not in ("model_loader", "model_converter")}


if __name__ == "__main__":
    d = build(force="--force" in sys.argv)
    print(d if d else "reference sources absent and no prebuilt oracle/_ref", flush=True)
