"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference files of the hot path into oracle/_ref/.

    python -m oracle.stage_ref          # in the build container, where /root/reference exists

oracle/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so the files travel to the GPU
box the way a compiled oracle/_ref/*.so would: there `bench.py --impl reference` runs the reference's OWN
`MultiAgentPPO._compute_adv_and_value_target` / `_compute_loss` / `step` (cpu_baseline.kind = "reference"), on the host
cores and -- as the "what SRL users get today on this GPU" line -- on cuda (the ATen-eager path).
The files are copied byte for byte (MANIFEST.json records their sha256); oracle/ref_loader.py execs them by path.
Nothing under srl_b200/ reads this directory."""
from __future__ import annotations

import hashlib
import json
import os
import shutil

SRC = os.environ.get("SRL_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# what oracle/ref_loader.py loads, plus what those files import at call time (mappo.py:220 -> base/timeutil.py)
FILES = [
    "base/namedarray.py",
    "base/numpy_utils.py",
    "base/gpu_utils.py",
    "base/timeutil.py",
    "api/policy.py",
    "api/trainer.py",
    "legacy/algorithm/modules/utils.py",
    "legacy/algorithm/modules/gae.py",
    "legacy/algorithm/modules/popart.py",
    "legacy/algorithm/modules/n_step_return.py",
    "legacy/algorithm/ppo/mappo.py",
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def stage(verbose: bool = True) -> bool:
    """Copies FILES from the reference checkout; returns False (and leaves oracle/_ref alone) when there is none."""
    if not os.path.isfile(os.path.join(SRC, FILES[-1])):
        return False
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, files=manifest), f, indent=1)
    if verbose:
        print(f"staged {len(FILES)} unmodified reference files into {DST}")
    return True


def verify() -> bool:
    """True when oracle/_ref holds every file and each matches its recorded hash."""
    try:
        with open(os.path.join(DST, "MANIFEST.json")) as f:
            files = json.load(f)["files"]
        return set(files) == set(FILES) and all(_sha(os.path.join(DST, rel)) == h for rel, h in files.items())
    except (OSError, ValueError, KeyError):
        return False


if __name__ == "__main__":
    if not stage():
        raise SystemExit(f"no reference checkout under {SRC}")
