"""Stage the UNMODIFIED reference modules under oracle/_ref/ (git-ignored, NOT gpurun-ignored: like the built .so files
the directory travels to the GPU box, where /root/reference does not exist).

    python oracle/stage_ref.py            # run in the build container; __graft_entry__.build() calls it too

Test infrastructure only.  What is staged is a byte-for-byte copy of the reference's top-level Python modules (the
hot-path pair run_plnerf.py / run_nerf_helpers.py plus the sibling modules they import at load time); a SHA-256
manifest is written next to them so that tests can state which reference they ran against.  Used by
  * bench.py --impl reference / the cpu_baseline leg  (kind "reference": the reference's own render() on CPU),
  * tests/test_dropin_reference.py                     (install() into the real run_plnerf module),
  * tests/test_abi.py                                  (inspect.signature of the real callables).
Nothing under pl-nerf_b200/ imports it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("PLNERF_REFERENCE_ROOT", "/root/reference")
FILES = ["run_plnerf.py", "run_nerf_helpers.py", "load_llff.py", "load_dtu.py", "load_blender.py", "nerf_extract_mesh.py",
         "run_nerf_vanilla.py", "configs/blender_linear.txt", "configs/llff_linear.txt", "configs/llff_constant.txt"]


def stage(verbose=True):
    if not os.path.isfile(os.path.join(SRC, "run_plnerf.py")):
        if verbose:
            print(f"stage_ref: {SRC} not present, nothing staged (using what is already under {DST}, if anything)")
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"stage_ref: staged {len(manifest)} files from {SRC} under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
