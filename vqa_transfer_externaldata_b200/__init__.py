"""B200-native (sm_100a) VQA answer-model hot path: forward + backward of the reference's
vqa/model_vlmap_answer* family behind its Model / importer interface. CUDA kernels live in
libvqa_answer_b200.so (csrc/, C ABI in include/vqa_answer.h); there is no CPU fallback."""
from .importer import get_model_class, get_model_types  # noqa: F401

__all__ = ["get_model_class", "get_model_types"]
