"""tests/golden/lpips_reference.npz: the reference's OWN known-answer fixture for LPIPS (lpips_tf2/test.py:17-19: "official
pytorch model metric value  ex_ref.png <-> ex_p0.png: 0.569,  ex_ref.png <-> ex_p1.png: 0.422"), i.e. the three 64x64 test images it
ships (lpips_tf2/imgs/) and those two numbers, plus what the oracle computes for them with the vendored weights (full precision and
per layer) so that runs without the 58 MB of weights can still check the arithmetic path they have.
Run in the build container (needs /root/reference):  python tests/golden/make_golden_lpips.py"""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ntc_oracle as O  # noqa: E402
from oracle import make_ref  # noqa: E402

R = "/root/reference/lpips_tf2/"
load = lambda n: np.asarray(Image.open(R + f"imgs/{n}.png"))[..., :3].astype(np.uint8)   # test.py load_image: drop alpha
ref, p0, p1 = load("ex_ref"), load("ex_p0"), load("ex_p1")
w = make_ref.load_lpips_weights()
val, layers = O.lpips(w, np.stack([ref, ref]), np.stack([p0, p1]), return_layers=True)
assert [round(float(v), 3) for v in val] == [0.569, 0.422], val
np.savez_compressed(os.path.join(HERE, "lpips_reference.npz"), ex_ref=ref, ex_p0=p0, ex_p1=p1, official=np.array([0.569, 0.422]),
                    oracle_value=val, oracle_layers=layers)
print("oracle", val, "official", [0.569, 0.422])
