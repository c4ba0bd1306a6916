"""not-gpu: the C-ABI library loads, exports every symbol include/sntc.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
  src = open(os.path.join(ROOT, "include", "sntc.h")).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(sntc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
  import shallow_ntc_b200 as pkg
  lib = ctypes.CDLL(pkg.LIB_PATH)
  declared = header_symbols()
  assert len(declared) >= 30
  for name in declared:
    assert hasattr(lib, name), f"{name} declared in include/sntc.h but not exported by libsntc.so"
  assert sorted(pkg.EXPORTED_SYMBOLS) == declared, "ctypes prototypes and the header disagree"


def test_struct_layouts_match_the_header():
  from shallow_ntc_b200 import _lib
  assert ctypes.sizeof(_lib.Tensor) == 48                    # DLTensor on LP64
  assert _lib.Tensor.shape.offset == 24 and _lib.Tensor.byte_offset.offset == 40
  assert ctypes.sizeof(_lib.TransformDesc) == 48
  assert ctypes.sizeof(_lib.ModelDesc) == 4 + 2 * 48 + 16
  assert ctypes.sizeof(_lib.ImageMetrics) == 24
  assert ctypes.sizeof(_lib.ImageRate) == 16
  assert _lib.ModelDesc.prior.offset == 4 + 2 * 48 + 12
  hdr = open(os.path.join(ROOT, "include", "sntc.h")).read()
  assert _lib.lib.sntc_version() == int(re.search(r"#define SNTC_VERSION (\d+)", hdr).group(1))


def test_sass_contains_tcgen05_and_tma():
  """The shipped cubin is sm_100a and really contains the Blackwell tensor / TMA instructions."""
  import shutil
  import subprocess
  import shallow_ntc_b200 as pkg
  if not shutil.which("cuobjdump"):
    pytest.skip("cuobjdump not on PATH")
  sass = subprocess.run(["cuobjdump", "-sass", pkg.LIB_PATH], capture_output=True, text=True).stdout
  assert "sm_100a" in sass
  for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
    assert mnemonic in sass, mnemonic


def test_no_cpu_fallback():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  from shallow_ntc_b200 import Context, SntcError, build_config, synthetic
  with pytest.raises(SntcError, match="no CPU fallback"):
    Context(0)
  m = build_config("jpegl")
  m.load_weights(synthetic.make_weights(m.variable_shapes()))
  zs, ys = m.latent_shapes(1, 64, 64)
  z, q = synthetic.make_latents(zs, ys)
  with pytest.raises(SntcError):
    m.decompress(z, q, (64, 64))


def test_product_never_imports_the_oracle():
  pkg_dir = os.path.join(ROOT, "shallow_ntc_b200")
  for dirpath, _, files in os.walk(pkg_dir):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
        txt = open(os.path.join(dirpath, f)).read()
        assert "oracle" not in txt.lower().replace("# oracle", ""), f"{f} mentions the oracle"
        assert "import torch" not in txt, f          # no PyTorch in the product: the collective is NCCL through libsntc
        if f != "tf_glue.py":                          # the reference-side binding imports TensorFlow lazily, on the reference's side only
          assert "import tensorflow" not in txt, f
