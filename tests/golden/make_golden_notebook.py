"""Golden fixture from outputs the REFERENCE ITSELF recorded: the PNG images embedded in the output cells of
``notebooks/vis_syn_filters.ipynb`` (real TensorFlow 2.10 runs of the trained jpegl model, ``checkpoints/jpegl/wid=3-...``:
``JPEGLikeSynthesis(kernel_size=18, strides=16)``, ``mshyper/configs/jpegl.py:38``).

  cell 44:  ei = zeros([1, 2, 2, 320]); ei[0, 0, 0, i] = 30;  bi = floats_to_pixels(model._synthesis(ei))      (100 channels i)
            -> 100 uint8 images [32, 32, 3]: the response of ONE active latent pixel at (0, 0) of a 2 x 2 latent grid on top of
               the constant background pixel(bias).
  cell 41:  ei = zeros([1, 1, 1, 320]); ei[0, 0, 0, i] = 30;  bi = floats_to_pixels(model._synthesis(ei) - g0)  (same channels)
            -> 100 uint8 images [16, 16, 3].

The weights are not available, but the GEOMETRY of these outputs is weight-independent and pins assumption A1 of the oracle (the
alignment of ``tf.keras.layers.Conv2DTranspose(padding="SAME")``: out[o] += in[n] W[a], o = n s + a - p with p = max(k - s, 0) // 2):
the support of the response in cell 44 is rows / columns 0..16 for every one of the 100 channels -- p = 1, not 0 (0..17) or 2 (0..15).
They also pin that floats_to_pixels maps the constant bias to one pixel value everywhere outside the support, and (cell 41 vs 44,
same channel) that the 1 x 1 and 2 x 2 latent grids crop the same kernel taps a = 1..16.

Run in the build container (the reference tree is mounted there):  python tests/golden/make_golden_notebook.py
"""
import base64, io, json, os, re
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NB = "/root/reference/notebooks/vis_syn_filters.ipynb"


def cell_images(nb, i):
  from PIL import Image
  out = []
  for o in nb["cells"][i].get("outputs", []):
    html = "".join(o.get("data", {}).get("text/html", ""))
    for b in re.findall(r'src="data:image/png;base64,([^"]+)"', html):
      out.append(np.array(Image.open(io.BytesIO(base64.b64decode(b)))))
  return np.stack(out)


if __name__ == "__main__":
  nb = json.load(open(NB))
  src44, src41 = "".join(nb["cells"][44]["source"]), "".join(nb["cells"][41]["source"])
  assert "np.zeros([1, 2, 2, C]" in src44 and "ei[0, 0, 0, i] = 30.0" in src44 and "np.zeros([1, 1, 1, C]" in src41
  c44, c41 = cell_images(nb, 44), cell_images(nb, 41)
  assert c44.shape == (100, 32, 32, 3) and c41.shape == (100, 16, 16, 3) and c44.dtype == np.uint8
  np.savez_compressed(os.path.join(HERE, "notebook_jpegl_responses.npz"), cell44=c44, cell41=c41,
                      source=np.array("notebooks/vis_syn_filters.ipynb cells 41, 44 (embedded PNG outputs of the trained jpegl model)"))
  print("saved", c44.shape, c41.shape)
