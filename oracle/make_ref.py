"""oracle/_ref/: artefacts made FROM the reference tree where it is mounted (the build container), git-ignored, shipped to the
GPU box with the working-tree snapshot.  TEST INFRASTRUCTURE ONLY.

The reference has no native code to compile; the one artefact is data: the LPIPS network weights it vendors as TensorFlow
checkpoints (``lpips_tf2/models/{vgg,lin}/exported.*``, 58 MB), re-saved as a plain ``.npz`` under the oracle's variable names so
that the oracle -- and the GPU tests, which cannot see ``/root/reference`` -- can evaluate the reference's own LPIPS known
answers (``lpips_tf2/test.py:17-19``).  No reference SOURCE is copied.   python oracle/make_ref.py
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SNTC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LPIPS_WEIGHTS = os.path.join(OUT, "lpips_weights.npz")


def build(force=False):
  """Returns the list of artefacts that exist afterwards (possibly empty when the reference is not mounted)."""
  made = []
  src = os.path.join(REF, "lpips_tf2", "models")
  if os.path.isdir(src) and (force or not os.path.exists(LPIPS_WEIGHTS)):
    import numpy as np
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ntc_oracle as O
    w = O.lpips_weights_from_reference_checkpoints(os.path.join(src, "vgg", "exported"), os.path.join(src, "lin", "exported"))
    os.makedirs(OUT, exist_ok=True)
    np.savez(LPIPS_WEIGHTS, **w)
  if os.path.exists(LPIPS_WEIGHTS):
    made.append(LPIPS_WEIGHTS)
  return made


def load_lpips_weights():
  """name -> float32 array, or None when neither oracle/_ref nor the reference tree is available."""
  import numpy as np
  if not os.path.exists(LPIPS_WEIGHTS):
    build()
  if not os.path.exists(LPIPS_WEIGHTS):
    return None
  with np.load(LPIPS_WEIGHTS) as z:
    return {k: z[k] for k in z.files}


if __name__ == "__main__":
  print(build(force="--force" in sys.argv))
