"""Model registry with the reference's model_type strings (vqa/importer.py:1-52).

Only the hot-path family is served; the other names of the reference raise NotImplementedError naming the
scope decision instead of silently mapping to something else."""

SUPPORTED = ("vlmap_answer", "vlmap_answer2", "vlmap_answer_no_noise", "vlmap_answer_noc", "vlmap_answer_nocarch",
             "vlmap_answer_full", "vlmap_answer_vqa_all", "vlmap_answer_vqa_all2", "vlmap_answer_adapt",
             "vlmap_answer_ent", "standard")
PLANNED = ()   # every vlmap_answer* member of vqa/importer.py:22-51 is served
OUT_OF_SCOPE = ("vqa", "standard_testmask", "standard_word2vec", "vlmap_only", "vlmap_finetune")


def get_model_types():
    return list(SUPPORTED)


def get_model_class(model_type="vlmap_answer"):
    from .model import (AdaptModel, Answer2Model, EntModel, FullModel, Model, NocArchModel, NocModel, NoNoiseModel,
                        StandardModel, VqaAll2Model, VqaAllModel)
    extra = {"vlmap_answer_full": FullModel, "vlmap_answer_vqa_all": VqaAllModel,
             "vlmap_answer_vqa_all2": VqaAll2Model, "vlmap_answer_adapt": AdaptModel, "vlmap_answer_ent": EntModel}
    if model_type in extra:
        return extra[model_type]
    if model_type == "vlmap_answer_noc":
        return NocModel
    if model_type == "vlmap_answer_nocarch":
        return NocArchModel
    if model_type == "vlmap_answer":
        return Model
    if model_type == "vlmap_answer2":
        return Answer2Model
    if model_type == "vlmap_answer_no_noise":
        return NoNoiseModel
    if model_type == "standard":
        return StandardModel
    if model_type in PLANNED:
        raise NotImplementedError(f"model_type {model_type!r}: variant of the answer-model family not built yet")
    if model_type in OUT_OF_SCOPE:
        raise NotImplementedError(f"model_type {model_type!r} is outside the hot path this package replaces")
    raise ValueError("Unknown model_type: {}".format(model_type))
