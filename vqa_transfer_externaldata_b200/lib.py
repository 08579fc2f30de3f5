"""ctypes binding of the C ABI declared in include/vqa_answer.h.

The shared library is built in-tree by `build.py` (nvcc, sm_100a) and loaded with ctypes.CDLL, so it
shows up in the process' loaded-library list. There is no fallback: if the library is missing,
`load()` raises, and every compute entry point returns an error without a B200.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvqa_answer_b200.so")

# ---- status / enums (mirror include/vqa_answer.h) ----
VQA_OK = 0
VQA_ERR_BAD_ARG, VQA_ERR_BAD_SHAPE, VQA_ERR_WORKSPACE = -1, -2, -3
VQA_ERR_CUDA, VQA_ERR_NO_DEVICE, VQA_ERR_STATE = -4, -5, -6
VARIANT_VLMAP_ANSWER, VARIANT_STANDARD, VARIANT_VLMAP_ANSWER2, VARIANT_VLMAP_ANSWER_NO_NOISE, VARIANT_VLMAP_ANSWER_NOC = range(5)
VARIANTS = {"vlmap_answer": 0, "standard": 1, "vlmap_answer2": 2, "vlmap_answer_no_noise": 3,
            "vlmap_answer_noc": 4, "vlmap_answer_nocarch": 4,   # nocarch is the same graph (model_vlmap_answer_nocarch.py)
            "vlmap_answer_full": 5, "vlmap_answer_vqa_all": 6, "vlmap_answer_vqa_all2": 7, "vlmap_answer_adapt": 8,
            "vlmap_answer_ent": 9}
PREC_BF16, PREC_FP32 = 0, 1

REPORT_KEYS = [
    "answer_train_loss", "answer_report_loss", "answer_acc", "exist_acc", "test_acc",
    "normal_test_acc", "normal_test_object_acc", "normal_test_attribute_acc", "normal_exist_acc",
    "normal_train_exist_acc", "max_exist_acc", "test_max_acc", "test_max_exist_acc",
]
# report slots after the 13 common ones (vqa/model_vlmap_answer_full.py:221-223); 0 for the other variants
EXTRA_REPORT_KEYS = ["latent_loss", "train_latent_loss", "entropy", "weighted_entropy"]
# which of them a model_type reports (model_vlmap_answer_full.py:221-223, model_vlmap_answer_ent.py:292-294)
VARIANT_REPORT_KEYS = {"vlmap_answer_full": ["latent_loss", "train_latent_loss"],
                       "vlmap_answer_ent": ["entropy", "weighted_entropy"]}
NUM_REPORT = len(REPORT_KEYS) + len(EXTRA_REPORT_KEYS)   # VQA_NUM_REPORT
PER_SAMPLE_KEYS = [
    "all_score", "max_train_score", "test_obj_score", "test_obj_max_score", "test_attr_score",
    "test_attr_max_score",
]

NUM_PHASES = 14
ACT_HQ, ACT_HL, ACT_HP, ACT_JD, ACT_Z, ACT_JDL, ACT_VA = range(7)
SITE_ATT, SITE_JOINT, SITE_JOINT_L, SITE_ENT = 1, 2, 3, 5   # dropout sites of vqa_dropout_mask_site

PARAM_FIELDS = [
    "embed", "v_w", "v_b", "v_gamma", "v_beta", "gru_gates_w", "gru_gates_b", "gru_cand_w",
    "gru_cand_b", "qv_w", "qv_b", "qv_gamma", "qv_beta", "att_w", "att_b", "pl_w", "pl_b",
    "pl_gamma", "pl_beta", "ql_w", "ql_b", "ql_gamma", "ql_beta", "joint_w", "joint_b",
    "joint_gamma", "joint_beta", "ans_w", "ans_b",
]


class VqaError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"vqa_answer error {status}: {msg}")
        self.status = status


class VqaConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "K", "Dv", "D", "L", "J", "A", "T", "W", "Vq", "num_train_answer", "variant",
                 "precision")] + [("keep_att", C.c_float), ("keep_joint", C.c_float), ("num_marginal", C.c_int32)]


# the extra question layer of model_vlmap_answer2 (q_L_ft2: FC + LayerNorm + tanh) and model_vlmap_answer_no_noise
# (q_L_mean: FC only); NULL in the struct for the other variants
EXTRA_FIELDS = ["qp_w", "qp_b", "qp_gamma", "qp_beta", "jl_w", "jl_b", "jl_gamma", "jl_beta", "al_w", "al_b",
                "qs_w", "qs_b", "tw_w", "tw_b", "va_w", "va_b", "va_gamma", "va_beta"]


def param_fields(variant):
    """Parameter fields a model_type owns, in struct order."""
    if variant == "vlmap_answer2":
        return PARAM_FIELDS + EXTRA_FIELDS[:4]
    if variant == "vlmap_answer_no_noise":
        return PARAM_FIELDS + EXTRA_FIELDS[:2]
    if variant in ("vlmap_answer_noc", "vlmap_answer_nocarch"):   # joint_l + WordWeightAnswerL
        return PARAM_FIELDS + EXTRA_FIELDS[4:10]
    if variant == "vlmap_answer_full":        # q_L_mean (qp_*) + q_L_log_sigma_sq (qs_*)
        return PARAM_FIELDS + ["qp_w", "qp_b", "qs_w", "qs_b"]
    if variant in ("vlmap_answer_vqa_all", "vlmap_answer_vqa_all2"):   # TunedWordWeightAnswer
        return PARAM_FIELDS + ["tw_w", "tw_b"]
    if variant == "vlmap_answer_adapt":       # v_adapt
        return PARAM_FIELDS + ["va_w", "va_b", "va_gamma", "va_beta"]
    return list(PARAM_FIELDS)


class VqaParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS + EXTRA_FIELDS]


class VqaFeatureBank(C.Structure):
    _fields_ = [("features", C.c_void_p), ("num_boxes", C.c_void_p), ("num_images", C.c_int64),
                ("features_bf16", C.c_void_p)]


class VqaBatch(C.Structure):
    _fields_ = [("batch_size", C.c_int32), ("q_len_max", C.c_int32), ("image_idx", C.c_void_p),
                ("q_intseq", C.c_void_p), ("q_intseq_len", C.c_void_p), ("answer_target", C.c_void_p)]


class VqaAnswerMasks(C.Structure):
    _fields_ = [("is_object", C.c_void_p), ("is_attribute", C.c_void_p), ("answer_exist", C.c_void_p)]


class VqaOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("loss", "report", "att_score", "logit", "pred", "per_sample", "condition", "pooled")]


class VqaGemmDesc(C.Structure):
    _fields_ = [("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("b_hi", C.c_void_p), ("b_lo", C.c_void_p),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("a_mn_major", C.c_int32),
                ("b_mn_major", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("bias", C.c_void_p), ("addend", C.c_void_p), ("ld_addend", C.c_int64),
                ("out_f32", C.c_void_p), ("ld_f32", C.c_int64), ("out_hi", C.c_void_p),
                ("out_lo", C.c_void_p), ("ld_bf", C.c_int64), ("block_n", C.c_int32)]


class VqaAttnFwd(C.Structure):
    _fields_ = [("batch", C.c_int32), ("z", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("hq", C.c_void_p), ("att_w", C.c_void_p), ("att_b", C.c_void_p), ("nbox", C.c_void_p),
                ("v_hi", C.c_void_p), ("v_lo", C.c_void_p), ("seed", C.c_uint64), ("step", C.c_uint64),
                ("att", C.c_void_p), ("pooled", C.c_void_p), ("pooled_hi", C.c_void_p),
                ("pooled_lo", C.c_void_p), ("ln_mean", C.c_void_p), ("ln_rstd", C.c_void_p), ("keep_bits", C.c_void_p)]


class VqaAttnBwd(C.Structure):
    _fields_ = [("batch", C.c_int32), ("z", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("hq", C.c_void_p), ("att_w", C.c_void_p), ("nbox", C.c_void_p), ("v_hi", C.c_void_p),
                ("v_lo", C.c_void_p), ("seed", C.c_uint64), ("step", C.c_uint64), ("att", C.c_void_p),
                ("ln_mean", C.c_void_p), ("ln_rstd", C.c_void_p), ("d_pooled", C.c_void_p),
                ("dz_hi", C.c_void_p), ("dz_lo", C.c_void_p), ("d_hq", C.c_void_p),
                ("d_att_w", C.c_void_p), ("d_att_b", C.c_void_p), ("d_gamma", C.c_void_p),
                ("d_beta", C.c_void_p), ("d_bias", C.c_void_p), ("keep_bits", C.c_void_p)]


# every symbol include/vqa_answer.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "vqa_create": (C.c_int32, [C.POINTER(VqaConfig), C.POINTER(_P)]),
    "vqa_destroy": (C.c_int32, [_P]),
    "vqa_last_error": (C.c_char_p, []),
    "vqa_abi_version": (C.c_int32, []),
    "vqa_launch_count": (C.c_uint64, []),
    "vqa_gru_kernel_path": (C.c_int32, []),
    "vqa_crc32c": (C.c_uint32, [C.c_char_p, C.c_uint64]),
    "vqa_tfrecord_index_host": (C.c_int32, [_P, C.c_uint64, C.c_int32, _P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "vqa_parse_examples_host": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, C.POINTER(C.c_int32),
                                            _P, _P, _P, C.c_int32, C.POINTER(C.c_int32), _P, _P]),
    "vqa_densify_targets": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "vqa_workspace_bytes": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "vqa_set_workspace": (C.c_int32, [_P, _P, C.c_uint64]),
    "vqa_prepare_params": (C.c_int32, [_P, C.POINTER(VqaParams), _P]),
    "vqa_forward": (C.c_int32, [_P, C.POINTER(VqaParams), C.POINTER(VqaFeatureBank), C.POINTER(VqaBatch),
                                C.POINTER(VqaAnswerMasks), C.c_uint64, C.c_uint64, C.POINTER(VqaOutputs),
                                _P]),
    "vqa_prefetch_features": (C.c_int32, [_P, C.POINTER(VqaFeatureBank), C.POINTER(VqaBatch), _P]),
    "vqa_backward": (C.c_int32, [_P, C.POINTER(VqaParams), C.POINTER(VqaBatch), C.POINTER(VqaParams),
                                 C.c_float, _P]),
    "vqa_dropout_masks": (C.c_int32, [_P, C.c_int32, C.c_uint64, C.c_uint64, _P, _P, _P]),
    "vqa_peek_activation": (C.c_int32, [_P, C.c_int32, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "vqa_adam_step": (C.c_int32, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float,
                                  C.c_float, C.c_float, C.c_int64, _P, _P]),
    "vqa_adam_step_shadowed": (C.c_int32, [_P, C.POINTER(VqaParams), _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float,
                                           C.c_float, C.c_float, C.c_float, C.c_int64, _P, _P]),
    "vqa_profile_enable": (C.c_int32, [_P, C.c_int32]),
    "vqa_profile_read": (C.c_int32, [_P, C.POINTER(C.c_float)]),
    "vqa_phase_name": (C.c_char_p, [C.c_int32]),
    "vqa_gemm": (C.c_int32, [_P, C.POINTER(VqaGemmDesc), _P]),
    "vqa_split_bf16": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, C.c_int64, _P]),
    "vqa_dropout_mask_site": (C.c_int32, [_P, C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, _P, _P]),
    "vqa_reparam_noise": (C.c_int32, [_P, C.c_int32, C.c_uint64, C.c_uint64, _P, _P]),
    "vqa_multimem_all_reduce": (C.c_int32, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vqa_set_embedding_slice_norm": (C.c_int32, [_P, _P]),
    "vqa_set_optimizer_tail": (C.c_int32, [_P, C.c_int64]),
    "vqa_sync_params": (C.c_int32, [_P, _P]),
    "vqa_set_deferred_outputs": (C.c_int32, [_P, C.c_int32]),
    "vqa_sync_outputs": (C.c_int32, [_P, _P]),
    "vqa_input_error_count": (C.c_int32, [C.POINTER(C.c_uint32), C.c_int32]),
    "vqa_set_early_gradients": (C.c_int32, [_P, C.c_int32]),
    "vqa_stream_wait_early_gradients": (C.c_int32, [_P, _P]),
    "vqa_set_gradient_allreduce": (C.c_int32, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32]),
    "vqa_keep_bits": (C.c_int32, [_P, C.c_int32, C.c_uint64, C.c_uint64, _P, _P]),
    "vqa_attn_fwd": (C.c_int32, [_P, C.POINTER(VqaAttnFwd), _P]),
    "vqa_attn_bwd": (C.c_int32, [_P, C.POINTER(VqaAttnBwd), _P]),
    "vqa_bce_metrics": (C.c_int32, [_P, C.c_int32, _P, _P, C.POINTER(VqaAnswerMasks), C.c_float, _P, _P,
                                    _P, _P, _P, _P]),
}


# ---- include/vqa_memft.h: operators of the vlmap pre-training path ----
class VqaSlabLn(C.Structure):
    _fields_ = [("slabs", C.c_int32), ("n", C.c_int32), ("N", C.c_int32), ("act", C.c_int32), ("z", C.c_void_p),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("mul", C.c_void_p), ("mul_rows", C.c_int64),
                ("keep", C.c_float), ("seed", C.c_uint64), ("step", C.c_uint64), ("site0", C.c_uint32),
                ("rows_per_site", C.c_int64), ("mean", C.c_void_p), ("rstd", C.c_void_p), ("y", C.c_void_p),
                ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("dout", C.c_void_p),
                ("dout2", C.c_void_p), ("dz_f32", C.c_void_p), ("dz_hi", C.c_void_p), ("dz_lo", C.c_void_p),
                ("dmul", C.c_void_p), ("part", C.c_void_p)]


class VqaLinearLn(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("backward", C.c_int32), ("a", C.c_void_p),
                ("lda", C.c_int64), ("w", C.c_void_p), ("ldw", C.c_int64), ("bias", C.c_void_p), ("gamma", C.c_void_p),
                ("beta", C.c_void_p), ("mul", C.c_void_p), ("act", C.c_int32), ("keep", C.c_float), ("seed", C.c_uint64),
                ("step", C.c_uint64), ("site", C.c_uint32), ("z", C.c_void_p), ("mean", C.c_void_p), ("rstd", C.c_void_p),
                ("y", C.c_void_p), ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("raw", C.c_void_p),
                ("dz_f32", C.c_void_p), ("dz_hi", C.c_void_p), ("dgamma_part", C.c_void_p), ("dbeta_part", C.c_void_p)]


class VqaSpatAttn(C.Structure):
    _fields_ = [("B", C.c_int32), ("K", C.c_int32), ("n", C.c_int32), ("D", C.c_int32), ("Dv", C.c_int32),
                ("kinds", C.c_int32), ("hv_hi", C.c_void_p), ("hv_lo", C.c_void_p), ("hq", C.c_void_p),
                ("att_w", C.c_void_p), ("att_b", C.c_void_p), ("num_boxes", C.c_void_p), ("v", C.c_void_p),
                ("keep", C.c_float), ("seed", C.c_uint64), ("step", C.c_uint64), ("site0", C.c_uint32),
                ("att", C.c_void_p), ("pooled", C.c_void_p), ("pooled_hi", C.c_void_p), ("pooled_lo", C.c_void_p),
                ("d_pooled", C.c_void_p), ("d_hv", C.c_void_p), ("d_hq", C.c_void_p), ("part", C.c_void_p),
                ("keep_bits", C.c_void_p)]


class VqaSoftmaxCe(C.Structure):
    _fields_ = [("heads", C.c_int32), ("B", C.c_int32), ("n", C.c_int32), ("A", C.c_int32), ("top_k", C.c_int32),
                ("logit", C.c_void_p), ("fills", C.c_void_p), ("num", C.c_void_p * 8), ("loss_scale", C.c_float), ("count", C.c_float * 8),
                ("stats", C.c_void_p), ("report", C.c_void_p), ("d_logit", C.c_void_p), ("d_hi", C.c_void_p),
                ("d_lo", C.c_void_p)]


class VqaGruSeq(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("L", C.c_int32), ("W", C.c_int32), ("Vq", C.c_int32),
                ("precision", C.c_int32), ("embed", C.c_void_p), ("gates_w", C.c_void_p), ("gates_b", C.c_void_p),
                ("cand_w", C.c_void_p), ("cand_b", C.c_void_p), ("tokens", C.c_void_p), ("len", C.c_void_p),
                ("ws", C.c_void_p), ("ws_bytes", C.c_uint64), ("q", C.c_void_p), ("q_hi", C.c_void_p),
                ("q_lo", C.c_void_p), ("dq", C.c_void_p), ("d_embed", C.c_void_p), ("d_gates_w", C.c_void_p),
                ("d_gates_b", C.c_void_p), ("d_cand_w", C.c_void_p), ("d_cand_b", C.c_void_p)]


MEMFT_SYMBOLS = {
    "vqa_ops_create": (C.c_int32, [C.POINTER(_P)]),
    "vqa_ops_destroy": (C.c_int32, [_P]),
    "vqa_ops_gemm": (C.c_int32, [_P, C.POINTER(VqaGemmDesc), C.c_int32, _P]),
    "vqa_ops_colsum": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    "vqa_ops_adam": (C.c_int32, [_P, _P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                 C.c_int64, _P, _P]),
    "vqa_ops_split_bf16": (C.c_int32, [_P, C.c_int64, C.c_int64, C.c_int64, _P, _P, C.c_int64, _P]),
    "vqa_ops_dropout_mask": (C.c_int32, [_P, C.c_int64, C.c_float, C.c_uint64, C.c_uint64, C.c_uint32, _P]),
    "vqa_ops_slab_ln_fwd": (C.c_int32, [_P, C.POINTER(VqaSlabLn), _P]),
    "vqa_ops_slab_ln_bwd": (C.c_int32, [_P, C.POINTER(VqaSlabLn), _P]),
    "vqa_ops_linear_ln": (C.c_int32, [_P, C.POINTER(VqaLinearLn), _P]),
    "vqa_ops_pad_planes": (C.c_int32, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P]),
    "vqa_ops_feat_wgrad": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _P, _P, C.c_int64, C.c_int32, _P, C.c_int32, _P, _P]),
    "vqa_memft_spat_attn_fwd": (C.c_int32, [_P, C.POINTER(VqaSpatAttn), _P]),
    "vqa_memft_spat_attn_bwd": (C.c_int32, [_P, C.POINTER(VqaSpatAttn), _P]),
    "vqa_memft_softmax_ce": (C.c_int32, [_P, C.POINTER(VqaSoftmaxCe), _P]),
    "vqa_memft_wordset_fwd": (C.c_int32, [_P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P]),
    "vqa_memft_wordset_bwd": (C.c_int32, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "vqa_ops_gru_workspace_bytes": (C.c_int32, [C.POINTER(VqaGruSeq), C.POINTER(C.c_uint64)]),
    "vqa_ops_gru_fwd": (C.c_int32, [_P, C.POINTER(VqaGruSeq), _P]),
    "vqa_ops_gru_bwd": (C.c_int32, [_P, C.POINTER(VqaGruSeq), _P]),
}

_lib = None


def load(path=None):
    """Load the CUDA library. Raises if it has not been built -- there is no Python/CPU fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            f"{p} not found: build it with `python -m vqa_transfer_externaldata_b200.build` "
            "(or __graft_entry__.build()). This package has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in list(SYMBOLS.items()) + list(MEMFT_SYMBOLS.items()):
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(status):
    if status != VQA_OK:
        raise VqaError(status, load().vqa_last_error().decode("utf-8", "replace"))
