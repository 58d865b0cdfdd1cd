"""Loader of the in-tree C-ABI shared library (foldcomp_b200/csrc/libfcz_engine.so).

The library is built by `__graft_entry__.build()` / `make -C foldcomp_b200/csrc`.  There is NO
fallback: if the library is missing the import fails loudly (the product path is CUDA only).
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# FCZ_ENGINE_LIB: another build of the SAME library (A/B experiments of tools/kernel_ab.py); there is still no fallback
LIB_PATH = os.environ.get("FCZ_ENGINE_LIB") or os.path.join(_HERE, "csrc", "libfcz_engine.so")

# every symbol include/fcz_engine.h declares
SYMBOLS = [
    "fcz_engine_create",
    "fcz_engine_destroy",
    "fcz_engine_set_opts",
    "fcz_encode_bound",
    "fcz_encode_batch",
    "fcz_decode_plan",
    "fcz_decode_batch",
    "fcz_pdb_text_plan",
    "fcz_pdb_text_batch",
    "fcz_decode_to_pdb_plan",
    "fcz_decode_to_pdb_batch",
    "fcz_extract_batch",
    "fcz_check_batch",
    "fcz_parse_pdb_plan",
    "fcz_parse_pdb_batch",
    "fcz_encode_pdb_text_batch",
    "fcz_unpack_angles_batch",
    "fcz_backbone_angles_batch",
    "fcz_host_alloc",
    "fcz_host_free",
    "fcz_engine_sync",
    "fcz_engine_launch_count",
    "fcz_engine_set_profiling",
    "fcz_engine_get_profile",
    "fcz_strerror",
    "fcz_last_error",
    "fcz_type_natoms",
    "fcz_type_name3",
    "fcz_type_atom_name",
    "fcz_type_alt_slot",
    "fcz_type_pred",
    "fcz_type_bond_length",
    "fcz_type_bond_angle",
]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA engine first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C foldcomp_b200/csrc). "
            "foldcomp_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.fcz_engine_create.restype = C.c_void_p
    lib.fcz_engine_create.argtypes = [C.c_int, P(abi.FczOpts)]
    lib.fcz_engine_destroy.restype = None
    lib.fcz_engine_destroy.argtypes = [C.c_void_p]
    lib.fcz_engine_set_opts.restype = C.c_int
    lib.fcz_engine_set_opts.argtypes = [C.c_void_p, P(abi.FczOpts)]
    lib.fcz_encode_bound.restype = C.c_uint64
    lib.fcz_encode_bound.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32]
    lib.fcz_encode_batch.restype = C.c_int
    lib.fcz_encode_batch.argtypes = [C.c_void_p, P(abi.FczChainBatch), P(abi.FczBlobBatch)]
    lib.fcz_decode_plan.restype = C.c_int
    lib.fcz_decode_plan.argtypes = [C.c_void_p, P(abi.FczBlobBatch), P(abi.FczChainBatch), P(abi.FczSizes)]
    lib.fcz_decode_batch.restype = C.c_int
    lib.fcz_decode_batch.argtypes = [C.c_void_p, P(abi.FczBlobBatch), P(abi.FczChainBatch)]
    lib.fcz_pdb_text_plan.restype = C.c_int
    lib.fcz_pdb_text_plan.argtypes = [C.c_void_p, P(abi.FczChainBatch), P(abi.FczTextBatch), P(C.c_uint64)]
    lib.fcz_pdb_text_batch.restype = C.c_int
    lib.fcz_pdb_text_batch.argtypes = [C.c_void_p, P(abi.FczChainBatch), P(abi.FczTextBatch)]
    lib.fcz_decode_to_pdb_plan.restype = C.c_int
    lib.fcz_decode_to_pdb_plan.argtypes = [C.c_void_p, P(abi.FczBlobBatch), P(abi.FczTextBatch), P(C.c_uint64)]
    lib.fcz_decode_to_pdb_batch.restype = C.c_int
    lib.fcz_decode_to_pdb_batch.argtypes = [C.c_void_p, P(abi.FczBlobBatch), P(abi.FczTextBatch)]
    lib.fcz_extract_batch.restype = C.c_int
    lib.fcz_extract_batch.argtypes = [C.c_void_p, P(abi.FczBlobBatch), C.c_int32, C.c_int32, P(abi.FczTextBatch), P(C.c_uint64)]
    lib.fcz_check_batch.restype = C.c_int
    lib.fcz_check_batch.argtypes = [C.c_void_p, P(abi.FczBlobBatch), C.c_void_p, C.c_void_p]
    lib.fcz_parse_pdb_plan.restype = C.c_int
    lib.fcz_parse_pdb_plan.argtypes = [C.c_void_p, P(abi.FczTextBatch), P(abi.FczChainBatch), P(abi.FczSizes)]
    lib.fcz_parse_pdb_batch.restype = C.c_int
    lib.fcz_parse_pdb_batch.argtypes = [C.c_void_p, P(abi.FczTextBatch), P(abi.FczChainBatch)]
    lib.fcz_encode_pdb_text_batch.restype = C.c_int
    lib.fcz_encode_pdb_text_batch.argtypes = [C.c_void_p, P(abi.FczTextBatch), C.c_void_p, C.c_void_p, P(abi.FczBlobBatch), P(C.c_uint64)]
    lib.fcz_unpack_angles_batch.restype = C.c_int
    lib.fcz_unpack_angles_batch.argtypes = [C.c_void_p, P(abi.FczBlobBatch), C.c_void_p, C.c_void_p, C.c_uint64, P(C.c_uint64)]
    lib.fcz_backbone_angles_batch.restype = C.c_int
    lib.fcz_backbone_angles_batch.argtypes = [C.c_void_p, P(abi.FczChainBatch), C.c_void_p]
    lib.fcz_host_alloc.restype = C.c_void_p
    lib.fcz_host_alloc.argtypes = [C.c_size_t]
    lib.fcz_host_free.restype = None
    lib.fcz_host_free.argtypes = [C.c_void_p]
    lib.fcz_engine_sync.restype = C.c_int
    lib.fcz_engine_sync.argtypes = [C.c_void_p]
    lib.fcz_engine_launch_count.restype = C.c_uint64
    lib.fcz_engine_launch_count.argtypes = [C.c_void_p]
    lib.fcz_engine_set_profiling.restype = C.c_int
    lib.fcz_engine_set_profiling.argtypes = [C.c_void_p, C.c_int]
    lib.fcz_engine_get_profile.restype = C.c_int
    lib.fcz_engine_get_profile.argtypes = [C.c_void_p, P(abi.FczProfile)]
    lib.fcz_strerror.restype = C.c_char_p
    lib.fcz_strerror.argtypes = [C.c_int]
    lib.fcz_last_error.restype = C.c_char_p
    lib.fcz_last_error.argtypes = [C.c_void_p]
    lib.fcz_type_natoms.restype = C.c_int
    lib.fcz_type_natoms.argtypes = [C.c_int]
    lib.fcz_type_name3.restype = C.c_char_p
    lib.fcz_type_name3.argtypes = [C.c_int]
    lib.fcz_type_atom_name.restype = C.c_char_p
    lib.fcz_type_atom_name.argtypes = [C.c_int, C.c_int]
    lib.fcz_type_alt_slot.restype = C.c_int
    lib.fcz_type_alt_slot.argtypes = [C.c_int, C.c_int]
    lib.fcz_type_pred.restype = C.c_int
    lib.fcz_type_pred.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.fcz_type_bond_length.restype = C.c_float
    lib.fcz_type_bond_length.argtypes = [C.c_int, C.c_int]
    lib.fcz_type_bond_angle.restype = C.c_float
    lib.fcz_type_bond_angle.argtypes = [C.c_int, C.c_int]
    _lib = lib
    return lib
