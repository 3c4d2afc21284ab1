"""ctypes binding of libspvd_b200.so (the C ABI declared in include/spvd_b200.h).

There is no CPU fallback: importing this module builds the library if the in-tree .so is missing or
stale and raises if that is impossible; every wrapper in spvd_b200.ops refuses non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_i, _f, _p, _sz, _ll = C.c_int, C.c_float, C.c_void_p, C.c_size_t, C.c_longlong

# name -> (restype, argtypes); mirrors include/spvd_b200.h one to one (tests/test_abi.py checks it)
SIGNATURES = {
    "spvd_last_error": (C.c_char_p, []),
    "spvd_abi_version": (_i, []),
    "spvd_point_voxel_coords": (_i, [_p, _i, _f, _f, _i, _p, _p, _p]),
    "spvd_unique_workspace_bytes": (_sz, [_i]),
    "spvd_unique_coords": (_i, [_p, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "spvd_hash_build": (_i, [_p, _i, _p, _i, _p, _p, _i, _p, _p]),
    "spvd_hash_query": (_i, [_p, _i, _i, _i, _p, _p, _i, _p, _p]),
    "spvd_kmap_subm3": (_i, [_p, _i, _p, _p, _p, _i, _p, _i, _p, _p]),
    "spvd_downsample_workspace_bytes": (_sz, [_i]),
    "spvd_downsample_coords": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "spvd_kmap_down2": (_i, [_p, _i, _p, _p, _p, _i, _p, _i, _p, _p, _i, _p, _p, _p, _p]),
    "spvd_batch_counts": (_i, [_p, _i, _p, _i, _p, _p]),
    "spvd_point_voxel_query": (_i, [_p, _i, _i, _p, _p, _i, _p, _p]),
    "spvd_devox_query": (_i, [_p, _i, _i, _p, _p, _i, _p, _p, _p]),
    "spvd_scatter_mean_fwd": (_i, [_p, _p, _i, _i, _p, _p, _i, _p]),
    "spvd_conv_packed_weight_count_f16": (_sz, [_i, _i, _i]),
    "spvd_conv_pack_weights_f16": (_i, [_p, _i, _i, _i, _p, _p]),
    "spvd_conv_fwd_pairs_simt": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "spvd_cat_affine_f16_fwd": (_i, [_p, _p, C.c_longlong, _i, _i, _p, _p, _i, _p, _p, _p]),
    "spvd_group_points_by_voxel": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "spvd_segment_mean_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _i, _p]),
    "spvd_point_mean_weights": (_i, [_p, _i, _i, _p, _p, _p]),
    "spvd_scatter_wmean_fwd": (_i, [_p, _p, _p, _i, _i, _p, _i, _p]),
    "spvd_chamfer_fwd": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "spvd_chamfer_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "spvd_emd_workspace_floats": (_sz, [_i, _i, _i]),
    "spvd_emd_approxmatch": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "spvd_emd_matchcost_fwd": (_i, [_p, _p, _p, _i, _i, _i, _p, _p]),
    "spvd_emd_matchcost_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "spvd_time_embed_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "spvd_scatter_mean_bwd": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "spvd_devoxelize_fwd": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "spvd_devoxelize_bwd": (_i, [_p, _p, _p, _i, _i, _p, _i, _p]),
    "spvd_round_tf32": (_i, [_p, _p, _ll, _p]),
    "spvd_conv_transpose_weights": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "spvd_conv_packed_weight_count": (_sz, [_i, _i, _i]),
    "spvd_conv_pack_weights": (_i, [_p, _i, _i, _i, _p, _p]),
    "spvd_conv_umma_supported": (_i, [_i, _i, _i]),
    "spvd_conv_fwd": (_i, [_p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _i, _p]),
    "spvd_conv_fwd_simt": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "spvd_conv_fwd_umma": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "spvd_conv_fwd_umma_rows": (_i, [_p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "spvd_kmap_sort_workspace_bytes": (_sz, [_i]),
    "spvd_kmap_sort": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "spvd_conv_sp_f16": (_i, [_p, _i, _i, _p]),
    "spvd_bn_train_fwd": (_i, [_p, _ll, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p]),
    "spvd_bn_train_bwd": (_i, [_p, _p, _ll, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p]),
    "spvd_colsum": (_i, [_p, _ll, _i, _p, _p, _p]),
    "spvd_film_bwd": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _p, _p, _i, _p]),
    "spvd_layernorm_train_fwd": (_i, [_p, _p, _p, _i, _i, _f, _p, _p, _p, _p]),
    "spvd_layernorm_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "spvd_attention_train_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "spvd_attention_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "spvd_cat_any_fwd": (_i, [_p, _p, _ll, _i, _i, _p, _p]),
    "spvd_split_fwd": (_i, [_p, _ll, _i, _i, _p, _p, _p]),
    "spvd_silu_fwd": (_i, [_p, _ll, _p, _p]),
    "spvd_silu_bwd": (_i, [_p, _p, _ll, _p, _p]),
    "spvd_add_fwd": (_i, [_p, _p, _ll, _p, _p]),
    "spvd_embedding_fwd": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "spvd_embedding_bwd": (_i, [_p, _p, _i, _i, _p, _p]),
    "spvd_timestep_embedding": (_i, [_p, _i, _i, _p, _p]),
    "spvd_mse_loss": (_i, [_p, _p, _ll, _f, _p, _p, _p, _p]),
    "spvd_sumsq": (_i, [_p, _ll, _p, _p]),
    "spvd_conv_pack_weights_train": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "spvd_adam_step": (_i, [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _f, _i, _p, _f, _f, _p, _p]),
    "spvd_conv_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "spvd_conv_bwd": (_i, [_p, _p]),
    "spvd_kmap_pairs": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "spvd_conv_fwd_pairs": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "spvd_conv_wgrad_umma": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "spvd_conv_wgrad": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "spvd_bn_silu_fwd": (_i, [_p, _p, _p, _ll, _i, _i, _p, _p]),
    "spvd_film_fwd": (_i, [_p, _p, _p, _ll, _i, _p, _p]),
    "spvd_film_affine_fwd": (_i, [_p, _p, _i, _p, _p, _p, _ll, _i, _i, _p, _p]),
    "spvd_cat_fwd": (_i, [_p, _p, _ll, _i, _i, _p, _p]),
    "spvd_layernorm_fwd": (_i, [_p, _p, _p, _i, _i, _f, _i, _p, _p]),
    "spvd_attention_fwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p]),
    "spvd_program_run": (_i, [_p, _i, _p, _i, _i, _i, _p, _sz, _p, _p]),
    "spvd_segmented_unique": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "spvd_ddpm_step": (_i, [_p, _p, _p, _f, _f, _f, _i, _i, _i, _f, _p, _p, _p, _p, _p]),
    "spvd_plan_build": (_i, [_p, _i, _i, _i, _f, _f, _i, _i, _p, _i, _i, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "spvd_plan_build_groups": (_i, [_p, _i, _i, _i, _f, _f, _i, _i, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _sz,
                                    _p, _p]),
    "spvd_plan_graph_create": (_i, [_i, _i, _i, _f, _f, _i, _i, _p, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i, _p, _sz, _p]),
    "spvd_plan_graph_launch": (_i, [_p, _p, _p, _p, _p]),
    "spvd_plan_graph_destroy": (None, [_p]),
    "spvd_plan_build_range": (_i, [_p, _i, _i, _i, _f, _f, _i, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p, _sz, _p]),
}


def _load():
    path = _build.LIB
    try:
        path = _build.build()
    except _build.NvccMissing as e:  # no nvcc on this machine: a prebuilt in-tree library is still fine (compile errors raise)
        if not os.path.exists(path):
            raise ImportError(f"libspvd_b200.so is missing and could not be built: {e}") from e
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    if lib.spvd_abi_version() != 1:
        raise ImportError("libspvd_b200.so ABI version mismatch")
    return lib


lib = _load()
LIB_PATH = _build.LIB


def check(rc: int):
    if rc != 0:
        raise RuntimeError(f"libspvd_b200 error {rc}: {lib.spvd_last_error().decode()}")
