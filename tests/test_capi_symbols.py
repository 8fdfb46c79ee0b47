"""The C-ABI shared library loads and exports every entry point include/detsam2.h declares (no compute
calls here: this runs on the GPU-less build box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "detsam2.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return sorted(set(re.findall(r"\b(?:int|int64_t|const char\s*\*|void)\s+(ds2_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 30, syms
    for must in ("ds2_gemm", "ds2_flash_attn", "ds2_mha", "ds2_layernorm", "ds2_connected_components",
                 "ds2_fill_holes", "ds2_resize_bilinear", "ds2_threshold_pack", "ds2_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from detsam2_b200 import build, capi
    path = build.build()  # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    # the ctypes binding types exactly the declared set
    assert sorted(capi.SYMBOLS) == declared_symbols()
    capi.load(build_if_missing=False)
    assert capi.load().ds2_version() >= 1


def test_only_sm100a_code_is_embedded():
    import shutil
    import subprocess
    from detsam2_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", build.build()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_shipped_flash_kernels_do_not_spill():
    """Guards the regression of profiles/r2_s9_flash_variants_ab.txt: measurement code left in the flash kernel pushed the
    cross-attention instantiation to 255 registers with spills (+25 % per launch).  The default instantiations
    (<BM 64, BN 128, ..., IL 1, TP 0> for the memory bank, the 256-row ones for the windows) must keep the 48-byte frame
    that holds only the tensor-map parameters."""
    import shutil
    import subprocess
    from detsam2_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--dump-resource-usage", build.build()], capture_output=True, text=True).stdout
    usage = dict(re.findall(r"Function (\S*flash_d256_tcgen05_kernel\S*):\s*\n\s*REG:(\d+ STACK:\d+)", out))
    assert usage, out[:400]
    default = [k for k in usage if "ILi64ELi128ELi1ELi3ELi2ELi1ELi1ELi1ELi0E" in k]
    assert len(default) == 1, sorted(usage)
    for name, u in usage.items():
        reg, stack = (int(x) for x in re.findall(r"\d+", u))
        if "ELi2ELi1ELi256EEEv" in name or "ELi2ELi0ELi256EEEv" in name:   # IL=2 (opt-in): workspace variant spills its 128-column row;
            continue                                           # the setmaxnreg variant reports the 168-register launch size

        assert stack <= 48, (name, u)
        assert reg <= 240, (name, u)


def test_engine_refuses_to_run_without_cuda():
    """The product path must fail loudly, never fall back to the CPU oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from detsam2_b200.build_sam import build_sam2_video_predictor
    from detsam2_b200.capi import Ds2Error
    with pytest.raises(RuntimeError):
        build_sam2_video_predictor("configs/sam2.1/sam2.1_hiera_t.yaml", device="cpu")
    with pytest.raises((Ds2Error, RuntimeError)):
        build_sam2_video_predictor("configs/sam2.1/sam2.1_hiera_t.yaml", device="cuda")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "det-sam2_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_from_pretrained_maps_hub_ids_like_the_reference(monkeypatch):
    """svp:208-222 / build_sam.py:148-163: model id -> (config, checkpoint file) -> factory; unknown ids are refused."""
    import pytest
    from detsam2_b200 import build_sam
    from detsam2_b200.predictor import SAM2VideoPredictor
    seen = {}

    def fake_factory(config_file, ckpt_path=None, **kw):
        seen.update(config_file=config_file, ckpt_path=ckpt_path, kw=kw)
        return "predictor"

    import huggingface_hub
    monkeypatch.setattr(huggingface_hub, "hf_hub_download", lambda repo_id, filename: f"/cache/{repo_id}/{filename}")
    monkeypatch.setattr(build_sam, "build_sam2_video_predictor", fake_factory)
    assert SAM2VideoPredictor.from_pretrained("facebook/sam2.1-hiera-large", device="cuda") == "predictor"
    assert seen["config_file"] == "configs/sam2.1/sam2.1_hiera_l.yaml"
    assert seen["ckpt_path"] == "/cache/facebook/sam2.1-hiera-large/sam2.1_hiera_large.pt"
    assert seen["kw"] == {"device": "cuda"}
    with pytest.raises(KeyError):
        SAM2VideoPredictor.from_pretrained("facebook/sam2-hiera-large")


def test_factory_accepts_the_reference_style_hydra_overrides():
    """build_sam.py:111-146: callers pass hydra_overrides_extra such as "++model.fill_hole_area=0"; the factory maps the
    keys that exist on this path and refuses unknown ones by name (no hydra involved)."""
    from detsam2_b200.build_sam import build_sam2_video_predictor, parse_hydra_overrides
    from detsam2_b200.config import get_config
    from detsam2_b200.weights import synthetic_state_dict
    from oracle import sam2_oracle as O
    cfg_over, pred_kw = parse_hydra_overrides(["++model.fill_hole_area=0", "++model.non_overlap_masks=true",
                                               "++model.sam_mask_decoder_extra_args.dynamic_multimask_stability_thresh=0.95",
                                               "++model.clear_non_cond_mem_around_input=true"])
    assert cfg_over == {"fill_hole_area": 0, "non_overlap_masks": True, "dynamic_multimask_stability_thresh": 0.95}
    assert pred_kw == {"clear_non_cond_mem_around_input": True}
    cfg = get_config("tiny", image_size=256)
    eng = O.OracleEngine(cfg, synthetic_state_dict(cfg, 0), fill_holes=False)
    pred = build_sam2_video_predictor("configs/sam2.1/sam2.1_hiera_t.yaml", device="cpu", engine=eng,
                                      hydra_overrides_extra=["++model.fill_hole_area=0", "++model.non_overlap_masks=true",
                                                             "++model.clear_non_cond_mem_around_input=true"])
    assert pred.fill_hole_area == 0 and pred.non_overlap_masks is True and pred.clear_non_cond_mem_around_input is True
    import pytest
    with pytest.raises(ValueError):
        parse_hydra_overrides(["++model.compile_image_encoder=true"])
    with pytest.raises(ValueError):
        parse_hydra_overrides(["nonsense"])
