"""ctypes binding of libsntc.so (include/sntc.h).  No PyTorch, no CPU fallback: importing works
without a GPU (so the registry and the symbol table can be inspected), but creating a context
raises unless an sm_100 device is present, and a missing shared library raises at import."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SNTC_LIB_PATH") or os.path.join(_HERE, "libsntc.so")   # SNTC_LIB_PATH: experimental builds (tools/)

SNTC_OK = 0
DL_CPU, DL_CUDA, DL_CUDA_HOST = 1, 2, 3
DL_INT, DL_UINT, DL_FLOAT = 0, 1, 2

# enum sntc_transform_kind
T_NONE = 0
T_HYPER_SYNTHESIS, T_JPEG_LIKE_HYPER, T_HYPER_SMALL = 1, 2, 3
T_JPEG_LIKE_SYNTHESIS, T_TWO_LAYER, T_TWO_LAYER_RES, T_MBT2018, T_BLS2017, T_CNN = 10, 11, 12, 13, 14, 15
T_TWO_LAYER_RES_D2S = 16
# enum sntc_activation
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU, ACT_IGDN1, ACT_GDN1, ACT_IGDN_CLASSIC = 0, 1, 2, 3, 4, 5
PRECISION_FP32, PRECISION_TC_F16X3, PRECISION_TC_F16X3_SYN2 = 0, 1, 2
LAUNCH_KINDS = ("total", "band_tc", "band_f32", "tail_mma", "tail_tc", "final_f32")   # SNTC_LAUNCH_*
INDEX_RINT, INDEX_TRUNC = 0, 1
PRIOR_NONE, PRIOR_DEEP_FACTORIZED = 0, 1
HOST_WRITE_COMBINED = 1
COMM_ID_BYTES, REDUCE_SUM, REDUCE_MAX = 128, 0, 1


class SntcError(RuntimeError):
  def __init__(self, code, msg):
    super().__init__(f"libsntc error {code}: {msg}")
    self.code = code


class Tensor(C.Structure):  # == DLTensor
  _fields_ = [("data", C.c_void_p), ("device_type", C.c_int32), ("device_id", C.c_int32), ("ndim", C.c_int32),
              ("dtype_code", C.c_uint8), ("dtype_bits", C.c_uint8), ("dtype_lanes", C.c_uint16),
              ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class TransformDesc(C.Structure):
  _fields_ = [("kind", C.c_int32), ("in_channels", C.c_int32), ("channels", C.c_int32 * 2),
              ("kernel_sizes", C.c_int32 * 2), ("strides", C.c_int32 * 2), ("activation", C.c_int32),
              ("n_layers", C.c_int32), ("use_bias", C.c_int32), ("use_offset", C.c_int32)]


class ModelDesc(C.Structure):
  _fields_ = [("struct_size", C.c_int32), ("hyper", TransformDesc), ("synthesis", TransformDesc),
              ("num_scales", C.c_int32), ("index_rounding", C.c_int32), ("precision", C.c_int32), ("prior", C.c_int32)]


class ImageMetrics(C.Structure):
  _fields_ = [("mse", C.c_double), ("psnr", C.c_double), ("ssd", C.c_uint64)]


class ImageRate(C.Structure):
  _fields_ = [("bits_y", C.c_double), ("bits_z", C.c_double)]


# name -> (restype, argtypes); every symbol include/sntc.h declares
_P = C.c_void_p
_PROTOS = {
  "sntc_version": (C.c_int, []),
  "sntc_last_error": (C.c_char_p, []),
  "sntc_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
  "sntc_destroy": (C.c_int, [_P]),
  "sntc_sync": (C.c_int, [_P]),
  "sntc_stream": (_P, [_P]),
  "sntc_device_name": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
  "sntc_device_pci_bus_id": (C.c_int, [_P, C.c_char_p, C.c_size_t]),
  "sntc_image_msssim": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(C.c_double), _P]),
  "sntc_lpips_create": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
  "sntc_lpips_destroy": (C.c_int, [_P]),
  "sntc_lpips_load_weights": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.c_int]),
  "sntc_lpips_finalize": (C.c_int, [_P]),
  "sntc_image_lpips": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
  "sntc_model_create": (C.c_int, [_P, C.POINTER(ModelDesc), C.POINTER(_P)]),
  "sntc_model_destroy": (C.c_int, [_P]),
  "sntc_model_num_variables": (C.c_int, [_P]),
  "sntc_model_variable": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
  "sntc_model_load_weights": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_float), C.POINTER(C.c_int64), C.c_int]),
  "sntc_model_finalize": (C.c_int, [_P]),
  "sntc_model_enable_graphs": (C.c_int, [_P, C.c_int]),
  "sntc_hyper_synthesis": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), _P]),
  "sntc_synthesis": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), _P]),
  "sntc_model_enable_vjp": (C.c_int, [_P, C.c_int]),
  "sntc_synthesis_vjp": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), _P]),
  "sntc_hyper_synthesis_vjp": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), _P]),
  "sntc_decode": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor),
                            C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(ImageMetrics), _P]),
  "sntc_decode_rd": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor),
                               C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(ImageMetrics), C.POINTER(ImageRate), _P]),
  "sntc_decode_hyper": (C.c_int, [_P, C.POINTER(Tensor), C.POINTER(Tensor), _P]),
  "sntc_decode_latents": (C.c_int, [_P, C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor),
                                    C.POINTER(Tensor), C.POINTER(ImageMetrics), _P]),
  "sntc_coder_create": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(_P)]),
  "sntc_coder_destroy": (C.c_int, [_P]),
  "sntc_coder_set_prior": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float)]),
  "sntc_coder_table": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.POINTER(C.c_uint32))]),
  "sntc_coder_encode": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.POINTER(C.c_uint8)), C.POINTER(C.c_size_t)]),
  "sntc_coder_decode": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_uint8), C.c_size_t, C.POINTER(C.c_int32)]),
  "sntc_coder_free": (None, [_P]),
  "sntc_last_stage_times_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
  "sntc_profile_enable": (C.c_int, [_P, C.c_int]),
  "sntc_profile_count": (C.c_int, [_P]),
  "sntc_profile_get": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
  "sntc_launch_count": (C.c_uint64, [_P]),
  "sntc_launch_counts": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
  "sntc_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
  "sntc_free": (C.c_int, [_P, _P]),
  "sntc_host_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
  "sntc_host_alloc_flags": (C.c_int, [_P, C.c_size_t, C.c_uint, C.POINTER(_P)]),
  "sntc_host_free": (C.c_int, [_P, _P]),
  "sntc_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
  "sntc_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t, _P]),
  "sntc_memset": (C.c_int, [_P, _P, C.c_int, C.c_size_t, _P]),
  "sntc_stream_create": (C.c_int, [_P, C.POINTER(_P)]),
  "sntc_stream_destroy": (C.c_int, [_P, _P]),
  "sntc_stream_wait_event": (C.c_int, [_P, _P, _P]),
  "sntc_stream_sync": (C.c_int, [_P, _P]),
  "sntc_comm_unique_id": (C.c_int, [_P]),
  "sntc_comm_create": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(_P)]),
  "sntc_comm_destroy": (C.c_int, [_P]),
  "sntc_comm_allreduce_f64": (C.c_int, [_P, C.POINTER(C.c_double), C.c_int, C.c_int]),
  "sntc_allreduce_metrics": (C.c_int, [_P, C.POINTER(C.c_double)]),
  "sntc_event_create": (C.c_int, [_P, C.POINTER(_P)]),
  "sntc_event_destroy": (C.c_int, [_P, _P]),
  "sntc_event_record": (C.c_int, [_P, _P, _P]),
  "sntc_event_elapsed_ms": (C.c_int, [_P, _P, _P, C.POINTER(C.c_float)]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)


def _load():
  if not os.path.exists(LIB_PATH):
    raise ImportError(
      f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
      "(nvcc -gencode arch=compute_100a,code=sm_100a). shallow_ntc_b200 has no CPU or PyTorch fallback.")
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in _PROTOS.items():
    fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
    fn.restype = res
    fn.argtypes = args
  return lib


lib = _load()


def check(code):
  if code != SNTC_OK:
    raise SntcError(code, lib.sntc_last_error().decode("utf-8", "replace"))
