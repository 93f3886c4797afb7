"""Stage the reference's own model files for the CPU reference arm (bench.py --impl reference, cpu_baseline kind
"reference"): copies /root/reference/models/*.py UNMODIFIED into oracle/_ref/models/ (git-ignored build output, like a
compiled oracle/_ref binary: it travels to the GPU box with the snapshot but never enters history) plus a one-line
stand-in for the un-installed third-party `fast_pytorch_kmeans` (imported at modules.py:8, used only by the rare
re-initialisation path at :489-499, which the benchmark never reaches). TEST / BENCH INFRASTRUCTURE - the product never
imports anything under oracle/.

    python oracle/vendor_ref.py        # in the authoring container (the GPU box has no /root/reference)
"""
import importlib.util
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models/__init__.py", "models/vqvae.py", "models/modules.py", "models/transformer.py", "losses/loss_seg.py"]


def vendor(verbose=True):
    """Returns True when oracle/_ref is populated (freshly or from an earlier run)."""
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(DST, "models", "vqvae.py"))
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    with open(os.path.join(DST, "fast_pytorch_kmeans.py"), "w") as f:
        f.write("class KMeans:  # stand-in for the absent third-party package (SURVEY.md 8c); never called by the benchmark\n"
                "    def __init__(self, *a, **k):\n        raise RuntimeError('fast_pytorch_kmeans is not installed')\n")
    if verbose:
        print("oracle/_ref: staged", ", ".join(FILES))
    return True


def available():
    return os.path.exists(os.path.join(DST, "models", "vqvae.py"))


def load_models():
    """Import the staged reference `models` package under a private name (the drop-in package is also called `models`)."""
    if "_mas_ref_models" in sys.modules:
        return sys.modules["_mas_ref_models"]
    if not available():
        raise RuntimeError("oracle/_ref is empty: run `python oracle/vendor_ref.py` where /root/reference exists")
    if DST not in sys.path:
        sys.path.append(DST)          # for the fast_pytorch_kmeans stand-in only (appended: never shadows a real install)
    spec = importlib.util.spec_from_file_location("_mas_ref_models", os.path.join(DST, "models", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(DST, "models")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_mas_ref_models"] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ok = vendor()
    print("available:", ok)
