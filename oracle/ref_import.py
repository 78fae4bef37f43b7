"""Import the UNMODIFIED reference transform.  TEST INFRASTRUCTURE ONLY.

Only usable where ``/root/reference`` exists (the build container); the GPU box
has no copy, so nothing executed there may call this.  The reference imports
``pytorchvideo`` at nexar_video_aug.py:16 for a symbol it never uses; the
package is not installed, so an empty stub module is pre-seeded.
"""
import os
import sys
import types

REFERENCE_DIR = os.environ.get("NEXAR_REFERENCE_DIR", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "nexar_video_aug.py"))


def import_reference_aug():
    if not reference_available():
        raise ImportError(f"reference not present at {REFERENCE_DIR}")
    if "pytorchvideo" not in sys.modules:
        pkg = types.ModuleType("pytorchvideo")
        sub = types.ModuleType("pytorchvideo.transforms")
        sub.create_video_transform = None
        pkg.transforms = sub
        sys.modules["pytorchvideo"] = pkg
        sys.modules["pytorchvideo.transforms"] = sub
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import nexar_video_aug  # noqa: E402
    return nexar_video_aug
