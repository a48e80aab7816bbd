"""ctypes binding of libmarsb200.so (the C ABI declared in include/marsb200.h).

There is no CPU fallback: if the shared library is missing the import fails
loudly, and every op refuses tensors that are not on a CUDA device.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmarsb200.so")

MASK_F32, MASK_U8 = 0, 1
GEMM_TCGEN05, GEMM_SIMT = 0, 1
PAIR_POPC, PAIR_MMA, PAIR_FP4, PAIR_AUTO = 0, 1, 2, 3


class MarsB200Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        f"or `make -C {os.path.join(_HERE, 'csrc')}`. marsb200 has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

_p = ctypes.c_void_p
_i = ctypes.c_int
_l = ctypes.c_int64
_f = ctypes.c_float
_d = ctypes.c_double

# name -> (restype, argtypes); mirrors include/marsb200.h one to one
SIGNATURES = {
    "marsb200_version": (_i, []),
    "marsb200_last_error": (ctypes.c_char_p, []),
    "marsb200_stream_sm_count": (_i, [_p, ctypes.POINTER(_i)]),
    "marsb200_stream_set_sm_cap": (_i, [_p, _i]),
    "marsb200_words_per_mask": (_l, [_l]),
    "marsb200_pad_rows": (_l, [_l]),
    "marsb200_pad_k": (_l, [_l]),
    "marsb200_normalize_rows": (_i, [_p, _l, _i, _l, _l, _i, _p, _p, _p]),
    "marsb200_pool_mask": (_i, [_p, _i, _l, _i, _i, _i, _p, _p]),
    "marsb200_sim_contract": (_i, [_p, _p, _p, _p, _i, _l, _l, _l, _p, _p, _p, _p, _i, _p]),
    "marsb200_match_argmax": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "marsb200_lsap": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "marsb200_vva_finalize": (_i, [_p, _p, _i, _l, _l, _p, _p]),
    "marsb200_attn_mean": (_i, [ctypes.POINTER(_p), _i, _i, _i, _i, _i, _p, _l, _p]),
    "marsb200_pir_workspace_bytes": (_l, [_i, _l]),
    "marsb200_pir_refine": (_i, [_p, _p, _l, _i, _i, _d, _i, _p, _p, _p, _l, _i, _p]),
    "marsb200_pir_stages": (_i, [_p, _p, _l, _i, _i, _d, _i, _p, _p, _p, _l, _i, _i, _p]),
    "marsb200_scoremap_boxes": (_i, [_p, _i, _i, _d, _p, _p, _p]),
    "marsb200_resize_minmax": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "marsb200_pack_masks": (_i, [_p, _i, _l, _l, _p, _p]),
    "marsb200_pack_masks_slice": (_i, [_p, _i, _l, _l, _l, _l, _p, _p]),
    "marsb200_pool_packed": (_i, [_p, _l, _i, _i, _i, _p, _p, _p, _p]),
    "marsb200_pack_pool_masks": (_i, [_p, _i, _l, _i, _i, _i, _p, _p, _p, _p, _p]),
    "marsb200_region_sums": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "marsb200_pairwise_inter": (_i, [_p, _i, _i, _l, _p, _i, _p]),
    "marsb200_pairwise_inter_slice": (_i, [_p, _i, _i, _l, _l, _l, _i, _p, _i, _p]),
    "marsb200_pack_pairwise": (_i, [_p, _i, _i, _i, _l, _p, _p, _i, _p]),
    "marsb200_emd_workspace_bytes": (_l, [_i, _i, _i, _l, _i, _i]),
    "marsb200_emd_scores": (_i, [_p, _p, _p, _i, _i, _l, _i, _i, _i, _p, _l, _p, _p, _p]),
    "marsb200_clip_scores": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "marsb200_clip_scores_f16": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "marsb200_fuse_rank": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _d, _d, _d, _f, _i, _p, _p, _p, _p, _p, _l, _p]),
    "marsb200_nms_bitmask": (_i, [_p, _i, _i, _f, _p, _p]),
    "marsb200_host_pack_masks": (_i, [_p, _i, _l, _l, _p, _i]),
    "marsb200_record_bytes": (_l, [_i]),
    "marsb200_merge_masks": (_i, [_p, _p, _i, _i, _l, _p, _p, _p]),
    "marsb200_points_in_masks": (_i, [_p, _l, _i, _i, _p, _i, _p, _p]),
    "marsb200_matcher_scores": (_i, [_p, _p, _p, _l, _i, _f, _f, _f, _p, _p, _p, _p]),
    "marsb200_masked_feature_means_workspace_bytes": (_l, [_i, _i, _i, _i]),
    "marsb200_masked_feature_means": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _l, _i, _p]),
    "marsb200_masked_sim_stats": (_i, [_p, _p, _p, _i, _l, _l, _p, _p]),
    "marsb200_masked_row_mean": (_i, [_p, _p, _i, _l, _l, _p, _p]),
    "marsb200_eval_areas": (_i, [_p, _p, _p, _l, _l, _p, _p]),
    "marsb200_eval_accumulate": (_i, [_p, _p, _l, _i, _p, _p, _p, _p]),
    "marsb200_eval_iou": (_i, [_p, _p, _i, _p, _i, _p, _p]),
    "marsb200_rle_workspace_bytes": (_l, [_l, _i, _i]),
    "marsb200_rle_decode": (_i, [_p, _p, _l, _i, _i, _p, _p, _l, _p, _p]),
    "marsb200_mask_boxes": (_i, [_p, _l, _i, _i, _p, _p]),
    "marsb200_stability_score": (_i, [_p, _l, _l, _f, _f, _p, _p, _p]),
    "marsb200_box_nms_workspace_bytes": (_l, [_i]),
    "marsb200_box_nms": (_i, [_p, _p, _i, _f, _p, _p, _p, _p, _l, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int) -> None:
    if rc != 0:
        raise MarsB200Error(f"marsb200 error {rc}: {lib.marsb200_last_error().decode(errors='replace')}")
