"""TEST INFRASTRUCTURE — imports the *unmodified* reference (/root/reference) in this container.

Only usable where /root/reference exists (the build container; NOT the GPU box).  Used by
``oracle/gen_golden.py`` to produce the committed fixtures under tests/golden/ and by the CPU tests
that pin ``oracle/sam2_oracle.py`` against the reference.  Nothing in the product imports this.

Three shims:
  1. ``sam2/__init__.py`` imports hydra (absent)  -> register a synthetic ``sam2`` package whose
     ``__path__`` is the reference directory, so sub-modules import without running ``__init__``;
  2. ``sam2/utils/misc.py:269`` calls ``os.path.isfile`` on ndarrays (TypeError on py3.12)
     -> wrapped to return False for non-path objects;
  3. hydra ``instantiate`` is absent -> the model is constructed here from the YAML values
     (sam2/configs/sam2.1/*.yaml, restated in detsam2_b200.config).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DS2_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "sam2", "modeling"))


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    pkg = types.ModuleType("sam2")
    pkg.__path__ = [os.path.join(REF_ROOT, "sam2")]
    sys.modules["sam2"] = pkg
    _orig_isfile = os.path.isfile

    def _isfile(p):
        if not isinstance(p, (str, bytes, os.PathLike, int)):
            return False
        return _orig_isfile(p)

    os.path.isfile = _isfile
    _installed = True


def build_reference_predictor(cfg, state_dict=None, device="cpu"):
    """Constructs the reference SAM2VideoPredictor from a detsam2_b200.config.ModelConfig.

    Mirrors sam2/build_sam.py:111-146 (eval overrides: dynamic multimask via stability,
    binarize_mask_from_pts_for_mem_enc, fill_hole_area=8) without hydra.
    """
    install()
    import torch
    from sam2.modeling.backbones.hieradet import Hiera
    from sam2.modeling.backbones.image_encoder import FpnNeck, ImageEncoder
    from sam2.modeling.memory_attention import MemoryAttention, MemoryAttentionLayer
    from sam2.modeling.memory_encoder import CXBlock, Fuser, MaskDownSampler, MemoryEncoder
    from sam2.modeling.position_encoding import PositionEmbeddingSine
    from sam2.modeling.sam.transformer import RoPEAttention
    from sam2.sam2_video_predictor import SAM2VideoPredictor

    trunk = Hiera(
        embed_dim=cfg.embed_dim, num_heads=cfg.num_heads, stages=tuple(cfg.stages),
        global_att_blocks=tuple(cfg.global_att_blocks), window_spec=tuple(cfg.window_spec),
        window_pos_embed_bkg_spatial_size=tuple(cfg.window_pos_embed_bkg_spatial_size),
    )
    neck = FpnNeck(
        position_encoding=PositionEmbeddingSine(num_pos_feats=256, normalize=True, scale=None, temperature=10000),
        d_model=256, backbone_channel_list=list(cfg.backbone_channel_list),
        fpn_top_down_levels=[2, 3], fpn_interp_model="nearest",
    )
    image_encoder = ImageEncoder(trunk=trunk, neck=neck, scalp=1)

    feat = cfg.image_size // 16 // 2  # yaml says [32,32] for 1024; recomputed at run time anyway

    def _layer():
        sa = RoPEAttention(rope_theta=10000.0, feat_sizes=[feat, feat], embedding_dim=256, num_heads=1,
                           downsample_rate=1, dropout=0.1)
        ca = RoPEAttention(rope_theta=10000.0, feat_sizes=[feat, feat], rope_k_repeat=True, embedding_dim=256,
                           num_heads=1, downsample_rate=1, dropout=0.1, kv_in_dim=64)
        return MemoryAttentionLayer(activation="relu", cross_attention=ca, d_model=256, dim_feedforward=2048,
                                    dropout=0.1, pos_enc_at_attn=False, pos_enc_at_cross_attn_keys=True,
                                    pos_enc_at_cross_attn_queries=False, self_attention=sa)

    memory_attention = MemoryAttention(d_model=256, pos_enc_at_input=True, layer=_layer(), num_layers=4)
    memory_encoder = MemoryEncoder(
        out_dim=64,
        position_encoding=PositionEmbeddingSine(num_pos_feats=64, normalize=True, scale=None, temperature=10000),
        mask_downsampler=MaskDownSampler(kernel_size=3, stride=2, padding=1),
        fuser=Fuser(layer=CXBlock(dim=256, kernel_size=7, padding=3, layer_scale_init_value=1e-6, use_dwconv=True),
                    num_layers=2),
    )
    model = SAM2VideoPredictor(
        image_encoder=image_encoder, memory_attention=memory_attention, memory_encoder=memory_encoder,
        num_maskmem=7, image_size=cfg.image_size,
        sigmoid_scale_for_mem_enc=20.0, sigmoid_bias_for_mem_enc=-10.0,
        use_mask_input_as_output_without_sam=True, directly_add_no_mem_embed=True, no_obj_embed_spatial=True,
        use_high_res_features_in_sam=True, multimask_output_in_sam=True, iou_prediction_use_sigmoid=True,
        use_obj_ptrs_in_encoder=True, add_tpos_enc_to_obj_ptrs=True, proj_tpos_enc_in_obj_ptrs=True,
        use_signed_tpos_enc_to_obj_ptrs=True, only_obj_ptrs_in_the_past_for_eval=True,
        pred_obj_scores=True, pred_obj_scores_mlp=True, fixed_no_obj_ptr=True,
        multimask_output_for_tracking=True, use_multimask_token_for_obj_ptr=True,
        multimask_min_pt_num=0, multimask_max_pt_num=1, use_mlp_for_obj_ptr_proj=True,
        compile_image_encoder=False,
        # eval-time overrides of build_sam.py:126-135
        sam_mask_decoder_extra_args=dict(dynamic_multimask_via_stability=True,
                                         dynamic_multimask_stability_delta=0.05,
                                         dynamic_multimask_stability_thresh=0.98),
        binarize_mask_from_pts_for_mem_enc=True,
        fill_hole_area=8,
    )
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=True)
        assert not missing and not unexpected
    model = model.to(device)
    model.eval()
    return model


def reference_video_processor_class(predictor, detector):
    """Imports the UNMODIFIED det_sam2_inference/det_sam2_RT.py with its absent third-party imports
    (ultralytics, IPython, matplotlib, pympler — SURVEY.md §0) replaced by inert stubs, and returns a
    subclass factory whose ``predictor`` is `predictor` and whose YOLO model is `detector`
    (``detector(frames_bgr) -> [[{"coordinates","class","confidence"}...]...]``), adapted to the
    ultralytics result interface the reference reads (det_sam2_RT.py:230-245)."""
    install()
    import numpy as np

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Val:
        def __init__(self, a):
            self._a = np.asarray(a)

        def cpu(self):
            return self

        def numpy(self):
            return self._a

        def __getitem__(self, i):
            return _Val(self._a[i])

    class _Box:
        def __init__(self, d):
            self.xyxy = _Val(np.asarray(d["coordinates"], np.float32)[None])
            self.cls = _Val(np.asarray(d["class"]).reshape(-1))
            self.conf = _Val(np.asarray(d["confidence"]).reshape(-1))

    class _Result:
        def __init__(self, dets):
            self.boxes = [_Box(d) for d in dets]

    class YOLO:
        def __init__(self, weights=None):
            pass

        def __call__(self, frames, **kw):
            return (_Result(d) for d in detector(frames))

    stub("pympler", asizeof=None)
    disp = stub("IPython.display", display=lambda *a, **k: None, Image=None, clear_output=lambda *a, **k: None)
    stub("IPython", display=disp)
    stub("ultralytics", checks=lambda: None, YOLO=YOLO)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            plt = stub("matplotlib.pyplot")
            stub("matplotlib", pyplot=plt)
    stub("sam2.build_sam", build_sam2_video_predictor=lambda *a, **k: predictor)
    inf = os.path.join(REF_ROOT, "det_sam2_inference")
    if inf not in sys.path:
        sys.path.insert(0, inf)
    sys.modules.pop("det_sam2_RT", None)
    import det_sam2_RT
    # the constructor enters a CUDA autocast context when a GPU is present (det_sam2_RT.py:101-107);
    # on this CPU-only container it is a no-op
    return det_sam2_RT.VideoProcessor
