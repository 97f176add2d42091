"""Import the UNMODIFIED reference (run_nerf_helpers, run_plnerf): from /root/reference in the build container, from
the byte-identical copies staged under oracle/_ref/ (oracle/stage_ref.py; git-ignored, travels to the GPU box) elsewhere.

Test infrastructure only (golden generation, the reference arm of bench.py, the drop-in tests).  The reference's
top-level imports need five packages that are not installed here; they are stubbed with empty modules (SURVEY.md
Appendix C) -- none of them is touched by the hot path.  Nothing under pl-nerf_b200/ imports this.
"""
import os
import sys
import types

STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _root():
    for r in (os.environ.get("PLNERF_REFERENCE_ROOT"), "/root/reference", STAGED):
        if r and os.path.isfile(os.path.join(r, "run_plnerf.py")):
            return r
    return None


REF_ROOT = _root()


def available():
    return REF_ROOT is not None


def load():
    """Returns (run_nerf_helpers, run_plnerf) modules of the reference."""
    if not available():
        raise RuntimeError("reference not found (neither /root/reference nor oracle/_ref; run oracle/stage_ref.py in the build container)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for n in ["imageio", "configargparse", "lpips", "natsort", "skimage", "skimage.metrics"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                sys.modules[n] = types.ModuleType(n)
    if not hasattr(sys.modules["lpips"], "LPIPS"):
        sys.modules["lpips"].LPIPS = object
    if not hasattr(sys.modules["natsort"], "natsorted"):
        sys.modules["natsort"].natsorted = sorted
    if not hasattr(sys.modules["skimage.metrics"], "structural_similarity"):
        sys.modules["skimage.metrics"].structural_similarity = lambda *a, **k: 0.0
    import run_nerf_helpers as H  # noqa: E402  (the reference's module)
    import run_plnerf as R        # noqa: E402
    return H, R
