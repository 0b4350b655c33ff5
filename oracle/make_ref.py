"""Recipe that makes the UNMODIFIED reference importable on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY (same rule as gpfq_oracle.py: only tests/, __graft_entry__ and bench.py's CPU
legs may touch anything under oracle/).

The reference (YixuanSeanZhou/Quantized_Neural_Nets) is pure Python with no packaging, and /root/reference does not
exist on the GPU box.  This script copies the three source files of the hot path

    src/step_algorithm.py   src/quantize_neural_net.py   src/utils.py

byte for byte from /root/reference into oracle/_ref/ -- a directory that is git-ignored (the reference's sources never
enter this repository's history) but NOT gpurun-ignored, so it travels to the GPU box like the built .so files.
`__graft_entry__.build()` runs it whenever /root/reference is present; `bench.py --impl reference` and the
`cpu_baseline` leg import the reference from there (kind "reference") and fall back to the oracle port (kind "port")
when the directory is absent.

    python oracle/make_ref.py            # (re)creates oracle/_ref/ ; prints what it did
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("step_algorithm.py", "quantize_neural_net.py", "utils.py")


def make_ref(verbose=False):
    """-> True when oracle/_ref holds the three files (copied now or already there)."""
    have = all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)
    if not os.path.isdir(REF_SRC):
        if verbose:
            print(f"{REF_SRC} is absent; oracle/_ref {'kept as is' if have else 'not available'}")
        return have
    os.makedirs(REF_DIR, exist_ok=True)
    sums = []
    for f in FILES:
        shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DIR, f))
        with open(os.path.join(REF_DIR, f), "rb") as fh:
            sums.append(f"{hashlib.sha256(fh.read()).hexdigest()[:16]}  {f}")
    with open(os.path.join(REF_DIR, "SOURCE.txt"), "w") as fh:
        fh.write("Unmodified copies of /root/reference/src files made by oracle/make_ref.py (not tracked by git).\n"
                 + "\n".join(sums) + "\n")
    if verbose:
        print("oracle/_ref:\n  " + "\n  ".join(sums))
    return True


def import_reference():
    """-> (step_algorithm module, quantize_neural_net module, utils module) of the unmodified reference, imported from
    oracle/_ref (or straight from /root/reference/src when that exists and the copy does not); None when neither is
    there.  The reference's modules import each other by bare name, so its directory goes on sys.path."""
    for d in (REF_DIR, REF_SRC):
        if all(os.path.exists(os.path.join(d, f)) for f in FILES):
            os.environ.setdefault("TQDM_DISABLE", "1")      # the reference wraps its feature loop in tqdm
            if d not in sys.path:
                sys.path.insert(0, d)
            import importlib
            mods = tuple(importlib.import_module(n) for n in ("step_algorithm", "quantize_neural_net", "utils"))
            for mod in mods:
                if os.path.dirname(os.path.abspath(mod.__file__)) != d:
                    raise ImportError(f"{mod.__name__} resolved to {mod.__file__}, not to the reference in {d}")
            return mods
    return None


if __name__ == "__main__":
    ok = make_ref(verbose=True)
    sys.exit(0 if ok else 1)
