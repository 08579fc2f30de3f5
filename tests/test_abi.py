"""The C-ABI library loads without a GPU and exports every symbol include/vqa_answer.h declares; the Python
binding (lib.SYMBOLS) covers exactly that set; compute entry points fail loudly when no B200 is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vqa_answer.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"VQA_API\s+[\w\s\*]+?\b(vqa_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from vqa_transfer_externaldata_b200 import build, lib as L
    build.build()
    return L.load()


def test_header_symbols_all_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vqa_answer.h but not exported"


def test_binding_covers_header_exactly():
    from vqa_transfer_externaldata_b200 import lib as L
    assert sorted(L.SYMBOLS) == _declared()


def test_struct_layouts_match_header():
    """Field ORDER of the ctypes structs = field order of the C structs (names parsed from the header)."""
    from vqa_transfer_externaldata_b200 import lib as L
    src = open(HEADER).read()
    for cname, cls in (("VqaConfig", L.VqaConfig), ("VqaParams", L.VqaParams), ("VqaBatch", L.VqaBatch),
                       ("VqaFeatureBank", L.VqaFeatureBank), ("VqaAnswerMasks", L.VqaAnswerMasks),
                       ("VqaOutputs", L.VqaOutputs), ("VqaGemmDesc", L.VqaGemmDesc),
                       ("VqaAttnFwd", L.VqaAttnFwd), ("VqaAttnBwd", L.VqaAttnBwd)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:  # "int64_t lda, ldb" declares two fields
                fields.extend(re.findall(r"(\w+)\s*$", part.strip())[0] for part in decl.split(","))
        assert fields == [f[0] for f in cls._fields_], cname
    assert L.NUM_PHASES == len(re.findall(r"\bVQA_PH_\w+\s*(?:=\s*0)?,", src))
    assert len(L.REPORT_KEYS) == 13 and len(L.PER_SAMPLE_KEYS) == 6 and len(L.PARAM_FIELDS) == 29


def test_no_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vqa_transfer_externaldata_b200 import lib as L
    from vqa_transfer_externaldata_b200.engine import AnswerModelConfig, Engine
    cfg = AnswerModelConfig(B=4, K=4, Dv=64, D=64, L=64, J=128, A=16, T=3, W=8, Vq=10, num_train_answer=12)
    with pytest.raises(RuntimeError):
        Engine(cfg)  # the product path has no CPU fallback
    c = cfg.to_c()
    h = C.c_void_p()
    st = lib.vqa_create(C.byref(c), C.byref(h))
    assert st in (L.VQA_ERR_NO_DEVICE, L.VQA_ERR_CUDA)
    assert lib.vqa_last_error()
    assert lib.vqa_abi_version() >= 1


def test_product_never_imports_oracle():
    """The oracle is the checker: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "vqa_transfer_externaldata_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_enums_match_header():
    """model_type -> VQA_VARIANT_* values, the report array length and the dropout sites follow the header."""
    from vqa_transfer_externaldata_b200 import importer, lib as L
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    variants = dict((n.lower(), int(v)) for n, v in re.findall(r"VQA_VARIANT_(\w+)\s*=\s*(\d+)", src))
    for model_type, value in L.VARIANTS.items():
        key = "vlmap_answer_noc" if model_type == "vlmap_answer_nocarch" else model_type   # nocarch is the same graph
        assert variants[key] == value, model_type
    assert set(importer.get_model_types()) == set(L.VARIANTS)
    report = re.search(r"VQA_REPORT_ANSWER_TRAIN_LOSS = 0,(.*?)VQA_NUM_REPORT", src, flags=re.S).group(1)
    assert 1 + len(re.findall(r"VQA_REPORT_\w+", report)) == L.NUM_REPORT == len(L.REPORT_KEYS) + len(L.EXTRA_REPORT_KEYS)
    acts = re.search(r"VQA_ACT_HQ = 0,(.*?)VQA_NUM_ACT", src, flags=re.S).group(1)
    assert 1 + len(re.findall(r"VQA_ACT_\w+", acts)) == 7 == L.ACT_VA + 1


MEMFT_HEADER = os.path.join(ROOT, "include", "vqa_memft.h")


def test_memft_header_binding_and_layouts(lib):
    """include/vqa_memft.h (the pre-training path's operators): every declared symbol is exported and bound, the binding
    binds nothing else, and the ctypes structs list the fields in the header's order."""
    from vqa_transfer_externaldata_b200 import lib as L
    src = re.sub(r"/\*.*?\*/", "", open(MEMFT_HEADER).read(), flags=re.S)
    names = sorted(set(re.findall(r"VQA_API\s+[\w\s\*]+?\b(vqa_\w+)\s*\(", src)))
    assert len(names) >= 15 and sorted(L.MEMFT_SYMBOLS) == names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vqa_memft.h but not exported"
    for cname, cls in (("VqaSlabLn", L.VqaSlabLn), ("VqaSpatAttn", L.VqaSpatAttn), ("VqaSoftmaxCe", L.VqaSoftmaxCe),
                       ("VqaGruSeq", L.VqaGruSeq), ("VqaLinearLn", L.VqaLinearLn)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                fields.extend(re.findall(r"(\w+)\s*(?:\[\d+\])?\s*$", part.strip())[0] for part in decl.split(","))
        assert fields == [f[0] for f in cls._fields_], cname
